#!/usr/bin/env python
"""Short digest of one kernel of an ncu --set full report: time, instructions, issue, stalls, SM-active balance.
usage: ncu_brief.py report.ncu-rep [row]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]; r = rows[2 + row]
g = lambda k: r[h.index(k)] if k in h else "n/a"
print(g("Kernel Name")[:80])
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_active.min",
          "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]:
    print(k.ljust(75), g(k))
st = [(float(r[i]), k.split("issue_stalled_")[1].split("_per_issue")[0]) for i, k in enumerate(h)
      if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
print("stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:9]))

#!/usr/bin/env python
"""Turns an `ncu --set full` report into the tracked summaries under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_r1a.ncu-rep r1a [--workload atari_peripheral]

Writes profiles/ncu_<tag>_summary.csv (one row per profiled launch, the metrics DESIGN.md quotes)
and refreshes profiles/traffic.json (DRAM bytes per launch of the ingest / observe kernels, which
bench.py reports as roofline.traffic).  Runs here, without a GPU (ncu -i only reads the report).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    workload = sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else "atari_peripheral"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ni = hdr.index("Kernel Name")
    out_rows, traffic = [], {}
    for r in body:
        name = r[ni]
        short = name.split("(")[0].split("::")[-1].strip()
        rec = {"kernel": short}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                rec[k + (" [" + units[i] + "]" if units[i] else "")] = r[i]
        out_rows.append(rec)
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            b += float(r[i]) * SCALE.get(units[i], 1.0)
        kind = "ingest" if "ingest" in short else ("observe" if "observe" in short else short)
        traffic.setdefault(kind, []).append(b)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    path = os.path.join(ROOT, "profiles", f"ncu_{tag}_summary.csv")
    cols = list(out_rows[0].keys())
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for rec in out_rows:
            w.writerow(rec)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    sys.path.insert(0, ROOT)
    from bench import sources_hash  # the same hash bench.py compares against: traffic is only reported for matching sources
    sha = sources_hash(workload)   # of the files this workload's kernels are built from
    try:
        tj = json.load(open(tpath))
    except Exception:
        tj = {}
    if not isinstance(tj.get("sources_sha16"), dict):
        tj = {"sources_sha16": {}, "workloads": {}, "source": {}}
    if tj["sources_sha16"].get(workload) != sha:  # captured on other kernel sources: this workload starts over
        tj["workloads"][workload] = {}
    tj["sources_sha16"][workload] = sha
    tj["workloads"].setdefault(workload, {})
    for kind, v in traffic.items():
        tj["workloads"][workload][kind] = sum(v) / len(v)
    tj["source"][workload] = f"ncu --set full, {os.path.basename(rep)} ({tag}); dram__bytes_read.sum + dram__bytes_write.sum per launch"
    json.dump(tj, open(tpath, "w"), indent=1, sort_keys=True)
    print(path, tpath, {k: sum(v) / len(v) for k, v in traffic.items()})


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Quick look at an ncu report (no GPU needed): headline metrics + stall breakdown per profiled launch."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_lsu.sum']
ni = hdr.index('Kernel Name')
for r in rows[2:]:
    print('=====', r[ni][:90])
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print('  %-75s %s %s' % (k, r[i], units[i]))
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
            v = float(r[i])
            if v >= 0.05:
                print('  stall %-40s %.3f' % (h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v))

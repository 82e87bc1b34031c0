#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -q -x -k "flexible or golden or fullsize or error_word or plain_bilinear" 2>&1 | tail -4
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "graph", round(d["ms_per_step"],4), "eager", round(d["ms_per_step_eager"],4), {k:(round(v["ms"],4),round(v["ms_best"],4)) for k,v in d["kernels"].items()})'
for i in 1 2; do python bench.py --workload atari_flexible --only --no-e2e --no-cpu-baseline --steps 48 2>$O/err_flex.log | python -c "$P" "flex"; done

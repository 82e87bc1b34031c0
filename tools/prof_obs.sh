#!/bin/bash
# ncu --set full capture of one launch of a kernel (regex $1) in the default bench; report -> gpurun_out/prof_$2.ncu-rep
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e ${3:-}"
$CMD > gpurun_out/plain_$2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$1 -s 8 -c 1 -f -o gpurun_out/prof_$2 $CMD > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log

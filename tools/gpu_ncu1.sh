#!/bin/bash
# One full ncu capture of a workload's kernel: tools/gpu_ncu1.sh <workload> <kernel regex> <tag>
O=gpurun_out
C2="python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --only"
$C2 > $O/plain_$3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$2" -s 8 -c 2 -f -o $O/prof_$3 $C2 > $O/ncu_f_$3.log 2>&1
echo "ncu $3 rc=$?"

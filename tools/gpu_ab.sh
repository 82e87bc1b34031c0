#!/bin/bash
# Same-box A/B of library builds: tools/gpu_ab.sh <workload> <steps> <lib> [<lib> ...] (interleaved, two rounds)
O=gpurun_out
W=$1; S=$2; shift 2
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "graph", round(d["ms_per_step"],4), "eager", round(d["ms_per_step_eager"],4), {k:(round(v["ms"],4),round(v["ms_best"],4)) for k,v in d["kernels"].items()})'
for i in 1 2; do for L in "$@"; do
  AGYM_LIB=$PWD/$L python bench.py --workload $W --only --no-e2e --no-cpu-baseline --steps $S 2>>$O/err_ab.log | python -c "$P" "$L"
done; done

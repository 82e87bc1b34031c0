#!/bin/bash
O=gpurun_out; T=r2n
echo skip-tests
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "graph", round(d["ms_per_step"],4), "eager", round(d["ms_per_step_eager"],4), d["launch"][:12], {k:(round(v["ms"],4)) for k,v in d["kernels"].items()})'
for w in dmc_fixed atari_fixed; do
  for v in "AGYM_X=0" "AGYM_NO_PDL=1" "AGYM_X=0" "AGYM_NO_PDL=1"; do
    env $v python bench.py --workload $w --only --no-e2e --no-cpu-baseline --steps 48 2>$O/err_$T.log | python -c "$P" "$w $v"
  done
done

#!/bin/bash
O=gpurun_out; T=r2j
python -m pytest tests -m gpu -q -x > $O/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$T.log
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],4), round(d["ms_per_step_eager"],4), {k:(round(v["ms"],4),round(v["ms_best"],4)) for k,v in d["kernels"].items()})'
for w in dmc_fixed atari_fixed; do
  for v in "AGYM_CROP_VER=2" "AGYM_CROP_WPE=1" "AGYM_CROP_WPE=2" "AGYM_CROP_WPE=4" "AGYM_CROP_WPE=0"; do
    env $v python bench.py --workload $w --only --no-e2e --no-cpu-baseline --steps 48 2>$O/err_$T.log | python -c "$P" "$w $v"
  done
done
python bench.py --no-e2e --no-cpu-baseline > $O/bench_${T}_noe2e.json 2>$O/err_$T.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2j_noe2e.json').read().strip().splitlines()[-1])
print('main', d['ms_per_step'], d['ms_per_step_eager'], d['launch'][:40], d['value'])
for k,v in d['workloads'].items(): print(k, round(v['ms_per_step'],4), round(v['ms_per_step_eager'],4), {a:round(b['ms'],4) for a,b in v['kernels'].items()})
print(d['strong_scaling']['ms_per_step'])
PY

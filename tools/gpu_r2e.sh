#!/bin/bash
O=gpurun_out; T=r2e
python -m pytest tests/test_gpu_persistent_paths.py -m gpu -q -x -k "fused" > $O/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_$T.log
for w in atari_fixed dmc_fixed; do
  python bench.py --workload $w --only --steps 30 --no-cpu-baseline --no-e2e > $O/bench_${T}_$w.json 2> $O/bench_${T}_$w.err; echo "$w rc=$?"; tail -2 $O/bench_${T}_$w.err
  AGYM_BENCH_UNFUSED=1 python bench.py --workload $w --only --steps 30 --no-cpu-baseline --no-e2e > $O/bench_${T}_${w}_unfused.json 2>/dev/null
done
python - <<PY
import json
for w in ('atari_fixed','dmc_fixed'):
    for s in ('','_unfused'):
        try:
            d=json.load(open('$O/bench_${T}_'+w+s+'.json')); print(w+s, d['value'], d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['roofline']['frac'])
        except Exception as e: print(w+s,'failed',e)
PY

#!/bin/bash
O=gpurun_out; T=${1:-r2b}
python -m pytest tests -m gpu -q -x > $O/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest_gpu_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke_$T.log
python bench.py --steps 20 --warmup 5 > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench rc=$?"; tail -5 $O/bench_$T.err
python - <<PY
import json
try:
    d=json.load(open('$O/bench_$T.json'))
    print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
    print('e2e',d['e2e'])
    print('cpu',d['cpu_baseline'])
    for k,v in d['workloads'].items(): print(k, v['value'], v['ms_per_step'], {a:b['ms'] for a,b in v['kernels'].items()}, v['roofline']['frac'], v.get('e2e',{}).get('value'), v.get('e2e',{}).get('error'), (v.get('cpu_baseline') or {}).get('value'))
    print('strong',d['strong_scaling'])
except Exception as e: print('parse failed',e)
PY

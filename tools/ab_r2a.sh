#!/bin/bash
# round-2 call A: parity of the new standard-geometry ingest kernel + A/B timings of its variants
O=gpurun_out; T=r2a
python -m pytest tests/test_gpu_parity.py tests/test_gpu_persistent_paths.py tests/test_gpu_golden.py -m gpu -q > $O/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu_$T.log
: > $O/ab_$T.jsonl
for v in "AGYM_INGEST_STD=0" "AGYM_INGEST_STD=160" "AGYM_INGEST_STD=176" "AGYM_INGEST_STD=176 AGYM_TM_L2=0" "AGYM_INGEST_STD=176 AGYM_TM_L2=64" "AGYM_INGEST_STD=176 AGYM_TM_L2=256" "AGYM_INGEST_STD=176"; do
  env $v timeout 120 python tools/ingest_ab.py 16384 40 1 >> $O/ab_$T.jsonl 2>> $O/ab_$T.err; echo "$v rc=$?"
done
env AGYM_INGEST_STD=0 timeout 120 python tools/ingest_ab.py 16384 40 0 >> $O/ab_$T.jsonl 2>> $O/ab_$T.err
env AGYM_INGEST_STD=176 timeout 120 python tools/ingest_ab.py 16384 40 0 >> $O/ab_$T.jsonl 2>> $O/ab_$T.err
cat $O/ab_$T.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['variant'], d['pcache'], round(d['ms_avg'],4), round(d['ms_min'],4), d['ring_hash'], d['head_sum'], d['pcache_sum'])"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/bench_$T.json')); print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()}, d['roofline']['frac'])"

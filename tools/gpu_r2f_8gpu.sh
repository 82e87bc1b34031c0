#!/bin/bash
# 8-GPU call: host-memory ceiling of the e2e path (all ranks copying at once), the env pipeline at 1 / 2 shards, the bench line
O=gpurun_out; T=r2f
nvidia-smi topo -m > $O/topo_$T.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > $O/lscpu_$T.txt
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/e2e_sweep.py atari_peripheral 6 1,2 > $O/e2e_sweep_${T}_${N}gpu.jsonl 2> $O/e2e_sweep_${T}_${N}gpu.err; echo "sweep rc=$?"
cut -c1-600 $O/e2e_sweep_${T}_${N}gpu.jsonl; tail -3 $O/e2e_sweep_${T}_${N}gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${T}_${N}gpu.json 2> $O/bench_${T}_${N}gpu.err; echo "bench rc=$?"; tail -3 $O/bench_${T}_${N}gpu.err
python - <<PY
import json
try:
    d=json.load(open('$O/bench_${T}_${N}gpu.json'))
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'].get('value'),d['e2e'].get('ms_per_step'),d['e2e'].get('error'))
    for k,v in d['workloads'].items(): print(k, v['value'], v['ms_per_step'], v.get('e2e',{}).get('value'), v.get('e2e',{}).get('error'))
    print('strong',d['strong_scaling'])
except Exception as e: print('parse failed',e)
PY

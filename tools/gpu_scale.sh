#!/bin/bash
# usage: gpu_scale.sh N TAG : the default bench line on N GPUs of one box under torch.distributed.run (what the driver's scaling run does)
N=$1; T=$2; O=gpurun_out
nvidia-smi topo -m > $O/topo_$T.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N \
  > $O/bench_${T}_${N}gpu.json 2> $O/bench_${T}_${N}gpu.err; echo "bench rc=$?"
tail -c 400 $O/bench_${T}_${N}gpu.json

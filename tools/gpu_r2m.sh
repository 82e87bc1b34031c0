#!/bin/bash
O=gpurun_out; T=r2m
nvidia-smi topo -p2p r > $O/p2p_$T.txt 2>&1
python -m pytest tests/test_vector_env.py -m gpu -q -x > $O/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_$T.log
timeout 300 python tools/p2p_gather_bench.py 16384 30 > $O/p2p_gather_$T.jsonl 2> $O/p2p_gather_$T.err; echo "p2p rc=$?"; cat $O/p2p_gather_$T.jsonl; tail -3 $O/p2p_gather_$T.err

#!/bin/bash
O=gpurun_out; T=r2k
python -m pytest tests -m gpu -q -x > $O/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$T.log
python tools/e2e_sweep.py atari_peripheral 8 1,2,4,8 > $O/e2e_sweep_$T.jsonl 2>$O/e2e_sweep_$T.err; cut -c1-330 $O/e2e_sweep_$T.jsonl
for w in atari_flexible dmc_fixed atari_fixed; do python tools/e2e_sweep.py $w 8 1,2,4 2>>$O/e2e_sweep_$T.err | tail -3 | cut -c1-200; done

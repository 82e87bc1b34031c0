#!/bin/bash
O=gpurun_out; T=r2l
python -m pytest tests -m gpu -q -x > $O/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$T.log
python tools/e2e_sweep.py atari_peripheral 10 1,2,1,2 > $O/e2e_sweep_$T.jsonl 2>$O/e2e_sweep_$T.err; cut -c1-330 $O/e2e_sweep_$T.jsonl
for w in atari_flexible dmc_fixed; do python tools/e2e_sweep.py $w 10 1,2 2>>$O/e2e_sweep_$T.err | tail -2 | cut -c1-200; done

#!/bin/bash
O=gpurun_out; T=r2i
python -m pytest tests -m gpu -q -x -k "peripheral or persistent or fullsize or golden" > $O/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$T.log
for i in 1 2; do
python bench.py --only --no-e2e --no-cpu-baseline --steps 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FS  ', d['ms_per_step'], {k:(round(v['ms'],4),round(v['ms_best'],4)) for k,v in d['kernels'].items()})"
AGYM_STD_NOFS=1 python bench.py --only --no-e2e --no-cpu-baseline --steps 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NOFS', d['ms_per_step'], {k:(round(v['ms'],4),round(v['ms_best'],4)) for k,v in d['kernels'].items()})"
done

#!/bin/bash
O=gpurun_out; T=r2g
python -m pytest tests -m gpu -q -x > $O/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu_$T.log
python - <<'PY' > gpurun_out/norm_timing_r2g.txt 2>&1
import torch, sys, time
sys.path.insert(0, '.')
from active_gym_b200 import ObservationPath
n=16384
p=ObservationPath(n,4,(84,84),(210,160,1),fov_size=(30,30),peripheral_res=(20,20),sensory_action_mode="relative",sensory_action_space=(-10.,10.))
f=torch.empty((n,210,160),dtype=torch.uint8,device='cuda'); p.synth_frames(f,1)
fl=torch.full((n,),5,dtype=torch.uint8,device='cuda')
for i in range(4): p.ingest_atari(f,f,fl)
a=torch.randint(-10,11,(n,2),device='cuda').double()
out=torch.empty((n,4,84,84),dtype=torch.uint8,device='cuda')
def t(fn,reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e=[(torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for x,y in e:
        x.record(); fn(); y.record()
    torch.cuda.synchronize()
    return sum(x.elapsed_time(y) for x,y in e)/reps
print('observe u8 only', t(lambda: p.observe_peripheral(a,out=out)))
for dt in (torch.float16, torch.bfloat16, torch.float32):
    no=torch.empty((n,4,84,84),dtype=dt,device='cuda')
    print(dt,'fused', t(lambda: p.observe_peripheral(a,out=out,norm_out=no)), 'separate', t(lambda: (p.observe_peripheral(a,out=out), p.normalize(out,dt,out=no))))
PY
cat gpurun_out/norm_timing_r2g.txt

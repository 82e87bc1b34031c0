#!/usr/bin/env python
"""A/B timing of the gray Atari ingest kernel alone (run once per variant; variants are picked by environment
variables read when libagym_b200.so is loaded: AGYM_INGEST_STD, AGYM_TM_L2, ...).  Prints one JSON line with the
average launch time (CUDA events around each launch) and checksums of ring / pcache for cross-variant comparison.
usage: ingest_ab.py [n_envs] [reps] [pcache 0|1]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_gym_b200 import ObservationPath

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
pc = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
p = ObservationPath(n, 4, (84, 84), (210, 160, 1), fov_size=(30, 30), peripheral_res=(20, 20) if pc else None,
                    sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0), device=dev)
frames = [torch.empty((n, 210, 160), dtype=torch.uint8, device=dev) for _ in range(6)]
for i, f in enumerate(frames):
    p.synth_frames(f, 1234 + 17 * i)
g = torch.Generator(device=dev).manual_seed(5)
flags_ragged = torch.randint(0, 16, (n,), dtype=torch.uint8, device=dev, generator=g)
flags_reset = torch.full((n,), 5, dtype=torch.uint8, device=dev)
flags_step = torch.full((n,), 3, dtype=torch.uint8, device=dev)
p.ingest_atari(frames[0], frames[0], flags_reset)
p.ingest_atari(frames[2], frames[3], flags_ragged)
t = 0
for _ in range(5):
    p.ingest_atari(frames[(2 * t) % 6], frames[(2 * t + 1) % 6], flags_step); t += 1
torch.cuda.synchronize()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for a, b in evs:
    a.record(); p.ingest_atari(frames[(2 * t) % 6], frames[(2 * t + 1) % 6], flags_step); b.record(); t += 1
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in evs)
ring_sum = int(p.ring.to(torch.int64).sum().item())
w = torch.arange(1, 1 + 7056 * 4, device=dev, dtype=torch.int64).reshape(1, 4, 84, 84) % 8191
ring_hash = int((p.ring.to(torch.int64) * w).sum().item())
pc_sum = float(p.pcache.double().sum().item()) if pc else 0.0
print(json.dumps({"variant": {k: v for k, v in os.environ.items() if k.startswith("AGYM_")}, "n": n, "pcache": pc,
                  "ms_avg": sum(ts) / len(ts), "ms_med": ts[len(ts) // 2], "ms_min": ts[0], "ring_sum": ring_sum,
                  "ring_hash": ring_hash, "head_sum": int(p.head.sum().item()), "pcache_sum": pc_sum}))

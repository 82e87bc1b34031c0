#!/usr/bin/env python
"""Learner-side gather of the observation shards of a multi-GPU env batch, two ways (one process, G GPUs):
  copy : every device writes its block locally, then ShardedVecEnv.gather copies the blocks to the learner GPU (peer copies)
  peer : ShardedVecEnv(learner_device=): the observe kernels store straight into the learner's tensor over NVLink
usage: p2p_gather_bench.py [envs_per_gpu] [steps]   -> JSON lines"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import active_gym_b200 as ag
from active_gym_b200.sources import SyntheticAtariSource

per = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
G = torch.cuda.device_count()
n = per * G
args = ag.AtariEnvArgs(game="boxing", seed=0, obs_size=(84, 84), fov_size=(30, 30), fov_init_loc=(0, 0),
                       sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0), peripheral_res=(20, 20))
devices = [f"cuda:{i}" for i in range(G)]
rng = np.random.default_rng(0)
acts = [{"motor_action": np.zeros(n, np.int64), "sensory_action": rng.integers(-10, 11, (n, 2)).astype(np.float64)} for _ in range(4)]


def make(learner):
    return ag.ShardedVecEnv(lambda m, d, lo, hi: ag.AtariFixedFovealPeripheralEnv(
        args, num_envs=m, source=SyntheticAtariSource(m, device=d, seed=1234 + lo), device=d), n, devices=devices, learner_device=learner)


def run(mode):
    sh = make("cuda:0" if mode == "peer" else None)
    sh.reset()
    last = None

    def step(i):
        nonlocal last
        obs = sh.step(acts[i % 4])[0]
        last = obs if mode == "peer" else ag.ShardedVecEnv.gather(obs, "cuda:0")
    for i in range(4):
        step(i)
    for d in devices:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    for d in devices:
        torch.cuda.synchronize(d)
    dt = (time.perf_counter() - t0) / steps
    chk = int(last.to(torch.int64).sum().item())
    sh.close()
    return dt, chk


res = {}
for mode in ("copy", "peer", "copy", "peer"):
    dt, chk = run(mode)
    res.setdefault(mode, []).append(dt)
    print(json.dumps({"mode": mode, "gpus": G, "envs_per_gpu": per, "ms_per_step": round(dt * 1e3, 4),
                      "obs_per_s": round(n / dt), "gathered_mb_per_step": round(n * 28224 / 1e6, 1), "checksum": chk}), flush=True)

# device-timed: the peripheral observe kernel of GPU 1 storing into its own memory vs into GPU 0's (CUDA events on GPU 1)
if G >= 2:
    from active_gym_b200 import ObservationPath
    from active_gym_b200.vector import enable_peer_access
    d1, d0 = torch.device("cuda:1"), torch.device("cuda:0")
    enable_peer_access(d1, d0)
    with torch.cuda.device(d1):
        p = ObservationPath(per, 4, (84, 84), (210, 160, 1), fov_size=(30, 30), peripheral_res=(20, 20),
                            sensory_action_mode="relative", sensory_action_space=(-10., 10.), device=d1)
        f = torch.empty((per, 210, 160), dtype=torch.uint8, device=d1)
        p.synth_frames(f, 1)
        fl = torch.full((per,), 5, dtype=torch.uint8, device=d1)
        for _ in range(4):
            p.ingest_atari(f, f, fl)
        a = torch.randint(-10, 11, (per, 2), device=d1).double()
        outs = {"local": torch.empty((per, 4, 84, 84), dtype=torch.uint8, device=d1),
                "peer": torch.empty((per, 4, 84, 84), dtype=torch.uint8, device=d0)}
        for name, out in outs.items():
            for _ in range(3):
                p.observe_peripheral(a, out=out)
            torch.cuda.synchronize(d1)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            for x, y in evs:
                x.record(); p.observe_peripheral(a, out=out); y.record()
            torch.cuda.synchronize(d1)
            ms = sum(x.elapsed_time(y) for x, y in evs) / len(evs)
            print(json.dumps({"kernel": "k_observe_peripheral_std", "out": name, "envs": per, "ms": round(ms, 4),
                              "store_gbs": round(per * 28224 / ms / 1e6, 1)}), flush=True)
        # (the two outputs are not compared here: every launch moves the foveae again; tests/test_vector_env.py checks
        # the peer-written batch against the oracle)

"""e2e probe: HostPipelinedEnv.step_host wall time per step for several shard counts, packed vs full H2D."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bench
from active_gym_b200.hostpipe import HostPipelinedEnv
w = bench.WORKLOADS["atari_peripheral"]; n = 16384
rng = np.random.default_rng(7)
for packed in (True, False):
    for shards in (8, 16, 32):
        env = HostPipelinedEnv.from_workload(w, n, torch.device("cuda:0"), shards=shards, obs_size=bench.S)
        if not packed:
            env = HostPipelinedEnv(n, w["K"], bench.S, w["raw"], kind=w["kind"], wrapper=w["wrapper"], variant=w["variant"],
                                   fov_size=w["fov"], sensory_action_mode=w["mode"], peripheral_res=w["periph"],
                                   device=torch.device("cuda:0"), shards=shards, packed_h2d=False)
        hf = [env.alloc_host_frames() for _ in range(2)]
        for f in hf:
            for t in f: t.numpy()[...] = rng.integers(0, 256, t.shape, dtype=np.uint8)
        act = rng.integers(-10, 11, (n, 2)).astype(np.float64); at = np.zeros(n, np.int32)
        env.reset_host(hf[0])
        for i in range(2): env.step_host(hf[i % 2], act, at)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(6): env.step_host(hf[i % 2], act, at)
        dt = (time.perf_counter() - t0) / 6
        print(f"packed={packed} shards={shards}: {dt*1e3:.2f} ms/step -> {n/dt/1e3:.0f} k obs/s  (h2d {env.h2d_bytes_per_step/1e6:.0f} MB)")
        del env, hf

#!/usr/bin/env python
"""End-to-end (host buffers in / out) sweep on one GPU: env-API pipeline at several shard counts next to the pure-copy
ceiling of the same byte volumes (pinned H2D of the packed raw rows, pinned D2H of the observations, both at once).
usage: e2e_sweep.py [workload] [steps]   -> JSON lines"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

wname = sys.argv[1] if len(sys.argv) > 1 else "atari_peripheral"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
w = bench.WORKLOADS[wname]
n = w["n"]
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
rank = int(os.environ.get("RANK", "0"))


def copy_ceiling(n_streams):
    rows = 168 if w["kind"] == "atari" else w["raw"][0]
    per_env_in = (2 if w["kind"] == "atari" else 1) * rows * w["raw"][1] * w["raw"][2]
    out_shape = (w["K"],) + (bench.S if w["wrapper"] != "fixed" else w["fov"])
    per_env_out = int(np.prod(out_shape))
    h_in = torch.empty(n * per_env_in, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n * per_env_in, dtype=torch.uint8, device=dev)
    h_out = torch.empty(n * per_env_out, dtype=torch.uint8).pin_memory()
    d_out = torch.empty(n * per_env_out, dtype=torch.uint8, device=dev)
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    res = {}
    for mode in ("h2d", "d2h", "both"):
        def go():
            for i, s in enumerate(streams):
                lo_i, hi_i = i * len(h_in) // n_streams, (i + 1) * len(h_in) // n_streams
                lo_o, hi_o = i * len(h_out) // n_streams, (i + 1) * len(h_out) // n_streams
                with torch.cuda.stream(s):
                    if mode in ("h2d", "both"):
                        d_in[lo_i:hi_i].copy_(h_in[lo_i:hi_i], non_blocking=True)
                    if mode in ("d2h", "both"):
                        h_out[lo_o:hi_o].copy_(d_out[lo_o:hi_o], non_blocking=True)
        go(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            go()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        res[mode] = {"ms": dt * 1e3, "gbs_in": n * per_env_in / dt / 1e9 if mode != "d2h" else 0, "gbs_out": n * per_env_out / dt / 1e9 if mode != "h2d" else 0}
    return res


print(json.dumps({"rank": rank, "workload": wname, "copy_ceiling_8_streams": copy_ceiling(8), "copy_ceiling_2_streams": copy_ceiling(2)}), flush=True)
for shards in (1, 2, 3, 4):
    dt, h2d, d2h = bench.measure_e2e(torch, wname, n, dev, steps, 2, shards)
    print(json.dumps({"rank": rank, "workload": wname, "shards": shards, "ms_per_step": dt / steps * 1e3, "obs_per_s": n * steps / dt,
                      "h2d_mb": h2d / 1e6, "d2h_mb": d2h / 1e6}), flush=True)

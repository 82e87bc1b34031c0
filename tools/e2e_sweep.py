#!/usr/bin/env python
"""End-to-end (host buffers in / out) probe: the env-API pipeline at several shard counts next to the pure-copy
ceiling of the same byte volumes (pinned H2D of the packed raw rows, pinned D2H of the observations, both at once).
Under torchrun every rank runs the same phase at the same time (barrier before each), so the numbers show what the
host's memory system delivers to N GPUs together.
usage: [torchrun --nproc-per-node N] e2e_sweep.py [workload] [steps] [shards,shards,...]   -> JSON lines (rank 0)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # sets CUDA_DEVICE_MAX_CONNECTIONS before the CUDA context exists
import numpy as np
import torch

wname = sys.argv[1] if len(sys.argv) > 1 else "atari_peripheral"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
shard_list = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4]
w = bench.WORKLOADS[wname]
n = w["n"]
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def gather(obj):
    if dist is None:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


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
        go()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            go()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        res[mode] = {"ms": round(dt * 1e3, 3), "gbs_in": round(n * per_env_in / dt / 1e9, 2) if mode != "d2h" else 0,
                     "gbs_out": round(n * per_env_out / dt / 1e9, 2) if mode != "h2d" else 0}
    return res


noceil = len(sys.argv) > 4 and sys.argv[4] == "NOCEIL"
c = gather(copy_ceiling(2)) if not noceil else []
if rank == 0 and not noceil:
    agg = {m: {"ms_max": max(r[m]["ms"] for r in c), "gbs_in_sum": round(sum(r[m]["gbs_in"] for r in c), 1),
               "gbs_out_sum": round(sum(r[m]["gbs_out"] for r in c), 1)} for m in ("h2d", "d2h", "both")}
    print(json.dumps({"n_gpus": world, "workload": wname, "envs_per_gpu": n, "copy_ceiling": agg, "per_rank": c}), flush=True)
for shards in shard_list:
    barrier()
    dt, h2d, d2h = bench.measure_e2e(torch, wname, n, dev, steps, 2, shards)
    r = gather({"ms_per_step": round(dt / steps * 1e3, 3), "obs_per_s": round(n * steps / dt)})
    if rank == 0:
        worst = max(x["ms_per_step"] for x in r)
        print(json.dumps({"n_gpus": world, "workload": wname, "shards": shards, "ms_per_step_max": worst,
                          "obs_per_s_total": round(n * world / worst * 1e3), "h2d_mb_per_gpu": h2d / 1e6, "d2h_mb_per_gpu": d2h / 1e6,
                          "per_rank_ms": [x["ms_per_step"] for x in r]}), flush=True)
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()

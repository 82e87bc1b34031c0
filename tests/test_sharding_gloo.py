"""CPU, world_size 2 over gloo: the env batch shards by env index with no data-path collective —
every rank transforms only its own contiguous block (here with the CPU oracle standing in for the
kernels) and the concatenation of the shards equals the unsharded result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_env_shard_partitions():
    from active_gym_b200.sharding import all_shards, env_shard
    for n in (0, 1, 7, 16384, 16385):
        for w in (1, 2, 3, 8):
            sh = all_shards(n, w)
            assert sh[0][0] == 0 and sh[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            assert max(h - l for l, h in sh) - min(h - l for l, h in sh) <= 1
    with pytest.raises(ValueError):
        env_shard(4, 2, 2)


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    from active_gym_b200.sharding import env_shard
    from oracle import agym_oracle as orc
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(5)  # every rank draws the same batch and keeps its slice
    K, fov, periph = 4, (30, 30), (20, 20)
    frames = rng.integers(0, 256, (3, 2, n, 210, 160), dtype=np.uint8)
    acts = rng.integers(-10, 11, (3, n, 2)).astype(np.float64)
    lo, hi = env_shard(n, rank, world)
    m = hi - lo
    ring, head = orc.new_state(m, K, (84, 84))
    loc = np.zeros((m, 2), np.int32)
    for t in range(3):
        flags = np.full(m, 5 if t == 0 else 3, np.uint8)
        orc.ingest_atari(frames[t, 0, lo:hi], frames[t, 1, lo:hi], flags, ring, head)
        orc.update_loc(acts[t, lo:hi], loc, obs_size=(84, 84), fov_size=fov, relative=True, lo=-10.0, hi=10.0)
    obs = torch.from_numpy(np.rint(orc.observe_peripheral(ring, head, loc, fov, periph)).astype(np.uint8))
    # gather only to CHECK the result; the data path itself has no collective
    sizes = [env_shard(n, r, world) for r in range(world)]
    bufs = [torch.empty((h - l,) + tuple(obs.shape[1:]), dtype=torch.uint8) for l, h in sizes]
    dist.all_gather(bufs, obs) if len({b.shape for b in bufs}) == 1 else dist.all_gather_object(bufs, obs)
    if rank == 0:
        np.save(out_path, torch.cat([torch.as_tensor(b) for b in bufs]).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_the_unsharded_batch(tmp_path):
    from oracle import agym_oracle as orc
    n, world = 6, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, port, n, out), nprocs=world, join=True)
    got = np.load(out)
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (3, 2, n, 210, 160), dtype=np.uint8)
    acts = rng.integers(-10, 11, (3, n, 2)).astype(np.float64)
    ring, head = orc.new_state(n, 4, (84, 84))
    loc = np.zeros((n, 2), np.int32)
    for t in range(3):
        orc.ingest_atari(frames[t, 0], frames[t, 1], np.full(n, 5 if t == 0 else 3, np.uint8), ring, head)
        orc.update_loc(acts[t], loc, obs_size=(84, 84), fov_size=(30, 30), relative=True, lo=-10.0, hi=10.0)
    want = np.rint(orc.observe_peripheral(ring, head, loc, (30, 30), (20, 20))).astype(np.uint8)
    assert np.array_equal(got, want)

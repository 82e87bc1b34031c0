"""Host-buffer front end of the observation path: raw frames and actions come from (pinned) host
memory, observations go back to (pinned) host memory — the interface a host-side simulator pool
or a CPU learner talks to.

The env batch is cut into env-index shards, each with its own ``ObservationPath`` and CUDA
stream, so that the H2D copy of shard i+1 overlaps the kernels and the D2H copy of shard i
(the two copy engines run in both directions at once).  Shards are independent — the same
property that lets the batch shard over GPUs with no collective.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import LUMA_DMC, LUMA_RGB, ObservationPath


class HostPipelinedEnv:
    def __init__(self, n_envs: int, frame_stack: int, obs_size, raw_shape, kind: str = "atari", wrapper: str = "fixed",
                 variant: str = "crop", fov_size=(30, 30), fov_init_loc=(0, 0), sensory_action_mode: str = "absolute",
                 sensory_action_space=(-10.0, 10.0), peripheral_res=None, device=None, shards: int = 16):
        self.kind, self.wrapper, self.variant = kind, wrapper, variant
        self.n_envs = int(n_envs)
        self.device = torch.device(device if device is not None else "cuda")
        shards = max(1, min(int(shards), self.n_envs))
        bounds = np.linspace(0, self.n_envs, shards + 1).astype(int)
        self.ranges: List[Tuple[int, int]] = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        luma = LUMA_RGB if kind == "atari" else LUMA_DMC
        self.paths = [ObservationPath(hi - lo, frame_stack, obs_size, raw_shape, luma=luma, fov_size=fov_size,
                                      fov_init_loc=fov_init_loc, sensory_action_mode=sensory_action_mode,
                                      sensory_action_space=sensory_action_space, peripheral_res=peripheral_res,
                                      device=self.device) for lo, hi in self.ranges]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.ranges]
        self.n_frames = 2 if kind == "atari" else 1
        self.d_frames = [[torch.empty(p.raw_frame_shape(), dtype=torch.uint8, device=self.device) for _ in range(self.n_frames)]
                         for p in self.paths]
        self.d_act = [torch.empty((p.n_envs, 2), dtype=torch.float64, device=self.device) for p in self.paths]
        self.d_atype = [torch.empty((p.n_envs,), dtype=torch.int32, device=self.device) for p in self.paths]
        self.d_out = [torch.empty(p.out_shape(wrapper, variant), dtype=torch.uint8, device=self.device) for p in self.paths]
        self.flags_step = [torch.full((p.n_envs,), 3 if kind == "atari" else 1, dtype=torch.uint8, device=self.device)
                           for p in self.paths]
        self.flags_reset = [torch.full((p.n_envs,), 5, dtype=torch.uint8, device=self.device) for p in self.paths]
        out_shape = (self.n_envs,) + tuple(self.d_out[0].shape[1:])
        self.h_obs = torch.empty(out_shape, dtype=torch.uint8).pin_memory()
        self.h_act = torch.empty((self.n_envs, 2), dtype=torch.float64).pin_memory()
        self.h_atype = torch.zeros((self.n_envs,), dtype=torch.int32).pin_memory()
        self.h_loc = torch.empty((self.n_envs, 2), dtype=torch.int32).pin_memory()
        frame_bytes = int(np.prod(self.paths[0].raw_frame_shape()[1:]))
        self.h2d_bytes_per_step = self.n_envs * (self.n_frames * frame_bytes + 16 + (4 if wrapper == "flexible" else 0))
        self.d2h_bytes_per_step = int(self.h_obs.numel()) + self.n_envs * 8

    @classmethod
    def from_workload(cls, w, n, device, shards=16, obs_size=(84, 84)):
        return cls(n, w["K"], obs_size, w["raw"], kind=w["kind"], wrapper=w["wrapper"], variant=w["variant"],
                   fov_size=w["fov"], sensory_action_mode=w["mode"], peripheral_res=w["periph"], device=device, shards=shards)

    def alloc_host_frames(self) -> Tuple[torch.Tensor, ...]:
        shape = (self.n_envs,) + tuple(self.paths[0].raw_frame_shape()[1:])
        return tuple(torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(self.n_frames))

    def _observe(self, i, action, atype, ctrl):
        p, out = self.paths[i], self.d_out[i]
        if self.wrapper == "peripheral":
            p.observe_peripheral(action, ctrl=ctrl, out=out)
        elif self.wrapper == "flexible":
            p.observe_flexible(action, atype, variant=self.variant, ctrl=ctrl, out=out)
        else:
            p.observe_fixed(action, variant=self.variant, ctrl=ctrl, out=out)

    def _run(self, frames: Sequence[torch.Tensor], reset: bool):
        for i, (lo, hi) in enumerate(self.ranges):
            with torch.cuda.stream(self.streams[i]):
                for d, h in zip(self.d_frames[i], frames):
                    d.copy_(h[lo:hi], non_blocking=True)
                p = self.paths[i]
                fl = self.flags_reset[i] if reset else self.flags_step[i]
                if self.kind == "atari":
                    fb = self.d_frames[i][1] if not reset else self.d_frames[i][0]
                    p.ingest_atari(self.d_frames[i][0], fb, fl)
                else:
                    p.ingest_dmc(self.d_frames[i][0], fl)
                if reset:
                    self._observe(i, None, None, "reset")
                else:
                    self.d_act[i].copy_(self.h_act[lo:hi], non_blocking=True)
                    self.d_atype[i].copy_(self.h_atype[lo:hi], non_blocking=True)
                    self._observe(i, self.d_act[i], self.d_atype[i], None)
                self.h_obs[lo:hi].copy_(self.d_out[i], non_blocking=True)
                self.h_loc[lo:hi].copy_(p.loc, non_blocking=True)
        for s in self.streams:
            s.synchronize()
        return self.h_obs, self.h_loc

    def reset_host(self, frames: Sequence[torch.Tensor]):
        """All envs: hard reset with one frame each (the Atari `frames[0]` is the un-pooled reset screen)."""
        return self._run(frames, reset=True)

    def step_host(self, frames: Sequence[torch.Tensor], sensory_action, sensory_action_type=None):
        """frames: pinned host tensors from ``alloc_host_frames``; sensory_action: (N,2) real array.
        Returns (observations, fov_loc) in pinned host memory, valid until the next call."""
        self.h_act.numpy()[...] = np.asarray(sensory_action, np.float64).reshape(self.n_envs, 2)
        if sensory_action_type is not None:
            self.h_atype.numpy()[...] = np.asarray(sensory_action_type, np.int32).reshape(self.n_envs)
        return self._run(frames, reset=False)

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

_cudart = None


def _rt():
    """libcudart through ctypes, for the strided (row-skipping) pinned copies torch has no call for."""
    global _cudart
    if _cudart is None:
        import ctypes as C
        last = None
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _cudart = C.CDLL(name)
                break
            except OSError as ex:
                last = ex
        if _cudart is None:
            raise RuntimeError(f"libcudart not found: {last}")
        vp, sz = C.c_void_p, C.c_size_t
        _cudart.cudaMemcpy2DAsync.argtypes = [vp, sz, vp, sz, sz, sz, C.c_int, vp]
        _cudart.cudaMemcpyAsync.argtypes = [vp, vp, sz, C.c_int, vp]
    return _cudart


def periodic_run(used: np.ndarray, n_rows: int):
    """If the sampled rows repeat with a period P that divides n_rows and form ONE cyclic run of L rows
    starting at offset o inside a period, returns (P, o, L); else None.  (210 -> 84: every row except
    g % 5 == 2, i.e. P = 5, o = 3, L = 4.)"""
    mask = np.zeros(n_rows, bool)
    mask[used] = True
    for P in range(1, n_rows + 1):
        if n_rows % P or not np.array_equal(mask, np.tile(mask[:P], n_rows // P)):
            continue
        m = mask[:P]
        if m.all():
            return None  # nothing to skip
        starts = [i for i in range(P) if m[i] and not m[i - 1]]
        if len(starts) != 1:
            return None
        return P, starts[0], int(m.sum())
    return None


class HostPipelinedEnv:
    def __init__(self, n_envs: int, frame_stack: int, obs_size, raw_shape, kind: str = "atari", wrapper: str = "fixed",
                 variant: str = "crop", fov_size=(30, 30), fov_init_loc=(0, 0), sensory_action_mode: str = "absolute",
                 sensory_action_space=(-10.0, 10.0), peripheral_res=None, device=None, shards: int = 16,
                 packed_h2d: bool = True):
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
        # packed H2D: ship only the raw rows the resize samples (transport only: the host buffers hold full frames)
        self.run = None
        rh, rw, rc = (int(v) for v in raw_shape)
        if packed_h2d and kind == "atari":
            self.run = periodic_run(self.paths[0].used_rows, rh)
        self.row_bytes = rw * rc
        if self.run:
            nu = len(self.paths[0].used_rows)
            shape = lambda p: (p.n_envs, nu, rw) if rc == 1 else (p.n_envs, nu, rw, rc)
            self.d_frames = [[torch.empty(shape(p), dtype=torch.uint8, device=self.device) for _ in range(self.n_frames)]
                             for p in self.paths]
        else:
            self.d_frames = [[torch.empty(p.raw_frame_shape(), dtype=torch.uint8, device=self.device)
                              for _ in range(self.n_frames)] for p in self.paths]
        self.raw_h = rh
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
        frame_bytes = int(np.prod(self.d_frames[0][0].shape[1:]))
        self.h2d_bytes_per_step = self.n_envs * (self.n_frames * frame_bytes + 16 + (4 if wrapper == "flexible" else 0))
        self.d2h_bytes_per_step = int(self.h_obs.numel()) + self.n_envs * 8

    @classmethod
    def from_workload(cls, w, n, device, shards=16, obs_size=(84, 84)):
        return cls(n, w["K"], obs_size, w["raw"], kind=w["kind"], wrapper=w["wrapper"], variant=w["variant"],
                   fov_size=w["fov"], sensory_action_mode=w["mode"], peripheral_res=w["periph"], device=device, shards=shards)

    def alloc_host_frames(self) -> Tuple[torch.Tensor, ...]:
        shape = (self.n_envs,) + tuple(self.paths[0].raw_frame_shape()[1:])
        return tuple(torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(self.n_frames))

    def _h2d_packed(self, dst: torch.Tensor, src: torch.Tensor, lo: int, hi: int):
        """Rows g of envs [lo, hi) with (g % P) inside the cyclic run [o, o + L) -> dst, in row order."""
        P, o, L = self.run
        rb, rt = self.row_bytes, _rt()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        periods = (hi - lo) * (self.raw_h // P)
        s0 = src.data_ptr() + lo * self.raw_h * rb
        d0 = dst.data_ptr()
        wrap = max(o + L - P, 0)
        H2D = 1
        if wrap == 0:
            err = rt.cudaMemcpy2DAsync(d0, L * rb, s0 + o * rb, P * rb, L * rb, periods, H2D, stream)
        else:
            err = rt.cudaMemcpyAsync(d0, s0, wrap * rb, H2D, stream)                       # rows 0 .. wrap-1 of the first period
            if not err and periods > 1:
                err = rt.cudaMemcpy2DAsync(d0 + wrap * rb, L * rb, s0 + o * rb, P * rb, L * rb, periods - 1, H2D, stream)
            if not err:                                                                     # rows o .. P-1 of the last period
                tail = (P - o) * rb
                err = rt.cudaMemcpyAsync(d0 + wrap * rb + (periods - 1) * L * rb, s0 + ((periods - 1) * P + o) * rb, tail, H2D, stream)
        if err:
            raise RuntimeError(f"packed H2D copy failed: cudaError {err}")

    def _observe(self, i, action, atype, ctrl):
        p, out = self.paths[i], self.d_out[i]
        if self.wrapper == "peripheral":
            p.observe_peripheral(action, ctrl=ctrl, out=out)
        elif self.wrapper == "flexible":
            p.observe_flexible(action, atype, variant=self.variant, ctrl=ctrl, out=out)
        else:
            p.observe_fixed(action, variant=self.variant, ctrl=ctrl, out=out)

    def _submit(self, frames: Sequence[torch.Tensor], reset: bool):
        """Enqueues one env step of every shard on its stream and returns at once."""
        for i, (lo, hi) in enumerate(self.ranges):
            with torch.cuda.stream(self.streams[i]):
                for d, h in zip(self.d_frames[i], frames):
                    if self.run:
                        self._h2d_packed(d, h, lo, hi)
                    else:
                        d.copy_(h[lo:hi], non_blocking=True)
                p = self.paths[i]
                fl = self.flags_reset[i] if reset else self.flags_step[i]
                if self.kind == "atari":
                    fb = self.d_frames[i][1] if not reset else self.d_frames[i][0]
                    (p.ingest_atari_packed if self.run else p.ingest_atari)(self.d_frames[i][0], fb, fl)
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

    def wait_host(self):
        """Blocks until the step submitted last is complete; returns (observations, fov_loc) in pinned host memory."""
        for s in self.streams:
            s.synchronize()
        return self.h_obs, self.h_loc

    def _run(self, frames: Sequence[torch.Tensor], reset: bool):
        self._submit(frames, reset)
        return self.wait_host()

    def submit_host(self, frames: Sequence[torch.Tensor], sensory_action, sensory_action_type=None) -> None:
        """Asynchronous ``step_host``: enqueue the step and return; ``wait_host`` collects the result.  Two env
        groups driven alternately (wait A, act, submit A, wait B, act, submit B, ...) keep the PCIe link busy
        in both directions across step boundaries, the way double-buffered samplers run host simulators."""
        self.h_act.numpy()[...] = np.asarray(sensory_action, np.float64).reshape(self.n_envs, 2)
        if sensory_action_type is not None:
            self.h_atype.numpy()[...] = np.asarray(sensory_action_type, np.int32).reshape(self.n_envs)
        self._submit(frames, reset=False)

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

"""TEST DOUBLE — ``OraclePath``: the surface of ``active_gym_b200.pipeline.PipelinedPath`` on CPU tensors, with the
C oracle standing in for the CUDA kernels.  It exists so that the host-side logic of the env classes (wrappers,
RecordWrapper flow, sources, VectorEnv adapter, multi-GPU front end) can be tested in the CPU-only container against
the fixtures recorded from the reference; the product never imports it (and fails loudly without a GPU).

Use: ``monkeypatch.setattr("active_gym_b200.atari_env.PipelinedPath", OraclePath)``.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import agym_oracle as orc


def _np(v, dtype=None):
    if v is None:
        return None
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    return np.ascontiguousarray(v, dtype) if dtype is not None else np.ascontiguousarray(v)


class OraclePath:
    streams = None

    def __init__(self, n_envs, frame_stack, obs_size, raw_shape, shards=1, device=None, luma=orc.LUMA_RGB, fov_size=None,
                 fov_init_loc=(0, 0), sensory_action_mode="absolute", sensory_action_space=(0.0, 0.0), peripheral_res=None,
                 cache_peripheral=True, packed_h2d=True, side_streams=False, antialias=True):
        self.antialias = bool(antialias)
        self.device = torch.device("cpu")
        self.n_envs, self.frame_stack = int(n_envs), int(frame_stack)
        self.obs_size = tuple(int(v) for v in obs_size)
        self.raw_shape = tuple(int(v) for v in raw_shape)
        self.fov_size = None if fov_size is None else tuple(int(v) for v in fov_size)
        self.peripheral_res = None if peripheral_res is None else tuple(int(v) for v in peripheral_res)
        self.relative = sensory_action_mode == "relative"
        self.lo, self.hi = (float(sensory_action_space[0]), float(sensory_action_space[1])) if self.relative else (0.0, 0.0)
        self.init_loc = np.rint(np.asarray(fov_init_loc, np.float64)).astype(np.int32)
        self.luma = tuple(luma)
        self._ring, self._head = orc.new_state(self.n_envs, self.frame_stack, self.obs_size)
        self.loc = torch.zeros((self.n_envs, 2), dtype=torch.int32)
        self.res = torch.zeros((self.n_envs, 2), dtype=torch.int32)
        if self.fov_size:
            self.res[:, 0], self.res[:, 1] = self.fov_size
        self.ep_len = torch.zeros((self.n_envs,), dtype=torch.int64)
        self.cum_reward = torch.zeros((self.n_envs,), dtype=torch.float64)
        self.shards = shards
        from active_gym_b200.pipeline import periodic_run
        self.run = periodic_run(self.used_rows, self.raw_shape[0]) if self.raw_shape[0] != self.obs_size[0] else None
        self.h2d_bytes = self.d2h_bytes = 0
        self.calls = []

    # ---- state as tensors
    @property
    def ring(self):
        return torch.from_numpy(self._ring)

    @property
    def head(self):
        return torch.from_numpy(self._head)

    @property
    def used_rows(self):
        from active_gym_b200 import _lib
        import ctypes as C
        h, S = self.raw_shape[0], self.obs_size[0]
        s0, s1, cf = (C.c_int32 * S)(), (C.c_int32 * S)(), (C.c_int32 * S)()
        _lib.lib().agym_table_cv2(h, S, 0, s0, s1, cf)
        return np.unique(np.concatenate([np.array(s0), np.array(s1)])).astype(np.int32)

    def out_shape(self, kind, variant, pad=None):
        N, K, S = self.n_envs, self.frame_stack, self.obs_size
        if kind == "peripheral" or variant in ("mask", "resize_full"):
            return (N, K) + S
        if kind == "flexible":
            return (N, K) + tuple(pad if pad is not None else S)
        return (N, K) + self.fov_size

    def join(self):
        pass

    def sync(self):
        pass

    def record_events(self):
        return []

    def prestage(self, action=None, action_type=None, raw_reward=None, done=None):
        self.calls.append(("prestage",))

    def read_errors(self):
        return 0

    def poll_errors(self):
        return 0

    # ---- ingest
    def _unpack(self, f, packed):
        f = _np(f)
        if f.ndim == 4 and f.shape[-1] == 1:
            f = f[..., 0]
        if not packed:
            return f
        h = self.raw_shape[0]
        full = np.zeros((f.shape[0], h) + f.shape[2:], np.uint8)
        full[:, self.used_rows] = f
        return full

    def ingest_atari(self, fa, fb, flags, packed=False):
        self.calls.append(("ingest_atari", packed))
        orc.ingest_atari(self._unpack(fa, packed), self._unpack(fb, packed), _np(flags, np.uint8), self._ring, self._head, self.luma)

    def ingest_dmc(self, frames, flags):
        self.calls.append(("ingest_dmc",))
        orc.ingest_dmc(_np(frames), _np(flags, np.uint8), self._ring, self._head, self.luma)

    def stack(self, out=None):
        return torch.from_numpy(orc.stack(self._ring, self._head))

    # ---- observe
    def _update(self, action, action_type, ctrl, flexible):
        n = self.n_envs
        loc, res = self.loc.numpy(), self.res.numpy()
        if isinstance(ctrl, str):
            ctrl = np.full(n, 1, np.uint8)
        ctrl = np.zeros(n, np.uint8) if ctrl is None else _np(ctrl, np.uint8)
        new_loc, new_res = loc.copy(), res.copy()
        if action is not None:
            a = _np(action, np.float64).reshape(n, 2)
            at = None if action_type is None else _np(action_type, np.int32).reshape(n)
            if flexible and at is not None:
                # FOV_RES: the oracle stores the action as the new res (fov_env.py:323)
                orc.update_loc(a, new_loc, obs_size=self.obs_size, fov_size=self.fov_size, relative=self.relative, lo=self.lo,
                               hi=self.hi, atype=at, res=new_res)
            elif flexible:
                orc.update_loc(a, new_loc, obs_size=self.obs_size, fov_size=self.fov_size, relative=self.relative, lo=self.lo,
                               hi=self.hi, atype=np.zeros(n, np.int32), res=new_res)
            else:
                orc.update_loc(a, new_loc, obs_size=self.obs_size, fov_size=self.fov_size, relative=self.relative, lo=self.lo, hi=self.hi)
        apply, reset = ctrl == 0, ctrl == 1
        loc[apply], res[apply] = new_loc[apply], new_res[apply]
        loc[reset] = self.init_loc
        res[reset] = np.array(self.fov_size, np.int32)

    def _finish(self, out, host_out, flexible=False, norm_out=None):
        out = torch.from_numpy(np.clip(np.rint(out), 0, 255).astype(np.uint8)) if out.dtype != np.uint8 else torch.from_numpy(out)
        if norm_out is not None:   # float32(u) / 255, rounded once more for f16 / bf16 (atari_env.py:75)
            norm_out.copy_((out.to(torch.float32) / 255.0).to(norm_out.dtype))
        if host_out:
            return out, (out, self.loc.clone(), self.res.clone() if flexible else None)
        return out

    def _oracle(self, fn, *a, **kw):
        """The oracle's antialias switch is process wide: set it for this call only."""
        orc.set_antialias(self.antialias)
        try:
            return fn(*a, **kw)
        finally:
            orc.set_antialias(True)

    def observe_fixed(self, action, variant="crop", ctrl=None, out=None, host_out=False, norm_out=None):
        self._update(action, None, ctrl, False)
        o = self._oracle(orc.observe_fixed, self._ring, self._head, self.loc.numpy(), self.fov_size, variant)
        return self._finish(o, host_out, norm_out=norm_out)

    def observe_peripheral(self, action, ctrl=None, out=None, use_cache=True, host_out=False, norm_out=None):
        self._update(action, None, ctrl, False)
        o = self._oracle(orc.observe_peripheral, self._ring, self._head, self.loc.numpy(), self.fov_size, self.peripheral_res)
        return self._finish(o, host_out, norm_out=norm_out)

    def observe_flexible(self, action, action_type=None, variant="mask", ctrl=None, pad=None, out=None, host_out=False, norm_out=None):
        self._update(action, action_type, ctrl, True)
        o = self._oracle(orc.observe_flexible, self._ring, self._head, self.loc.numpy(), self.res.numpy(), self.fov_size, variant,
                         pad=pad if pad is not None else self.obs_size)
        return self._finish(o, host_out, flexible=True, norm_out=norm_out)

    # ---- RecordWrapper counters (restates k_record_step)
    def record_step(self, raw_reward=None, done=None, reset_mask=None, is_reset=False, trace_row=None, with_res=False,
                    host_out=False):
        n = self.n_envs
        if is_reset:
            sel = np.ones(n, bool) if reset_mask is None else _np(reset_mask).astype(bool)
            self.ep_len[torch.from_numpy(sel)] = 0
            self.cum_reward[torch.from_numpy(sel)] = 0
            valid = sel.astype(np.int32)
        else:
            self.ep_len += 1
            if raw_reward is not None:
                self.cum_reward += torch.from_numpy(_np(raw_reward, np.float64).reshape(n))
            valid = np.ones(n, np.int32) if done is None else (~_np(done).astype(bool)).astype(np.int32)
        if trace_row is not None:
            trace_row[:, 0:2] = self.loc
            trace_row[:, 2:4] = self.res if with_res else 0
            trace_row[:, 4] = self.ep_len.to(torch.int32)
            trace_row[:, 5] = torch.from_numpy(valid)
        if host_out:
            return self.ep_len.clone(), self.cum_reward.clone()
        return None

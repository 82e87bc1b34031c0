"""Vector front ends over the batched foveal envs.

* ``FovealVectorEnv`` — the ``gymnasium.vector.VectorEnv`` surface the reference reaches through
  ``gym.vector.SyncVectorEnv`` of thunks (atari_env.py:221-241, 276): ``num_envs``, ``single_*_space`` / batched
  spaces, ``reset(seed=, options=)``, ``step`` 5-tuple with batched ``info``, ``step_async`` / ``step_wait``,
  SyncVectorEnv-style autoreset (``final_observation`` / ``final_info``), ``call`` / ``get_attr`` / ``set_attr`` and
  per-env views in ``.envs`` (``envs[0].fov_loc``, ``envs[0].save_record_to_file(...)``).  It subclasses
  ``gymnasium.vector.VectorEnv`` when gymnasium is importable and is duck-typed otherwise.
* ``ShardedVecEnv`` — one env batch over several GPUs of a box: the global env index is cut into contiguous
  blocks (``sharding.env_shard``), one batched env per device, no collective.  Steps are submitted
  to every device before any is waited for, so the devices (and their host->device copies) run concurrently.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from .sharding import all_shards
from .spaces import HAVE_GYMNASIUM, Box, Dict, Discrete

if HAVE_GYMNASIUM:  # pragma: no cover - this image has no gymnasium
    import gymnasium as _gym
    _VectorBase = _gym.vector.VectorEnv
else:
    class _VectorBase:  # the attributes gymnasium.vector.VectorEnv documents
        is_vector_env = True
        closed = False
        metadata = {"render_modes": []}
        render_mode = None
        spec = None


def batch_space(space, n: int):
    """``gymnasium.vector.utils.batch_space`` for the space types these envs use."""
    if isinstance(space, Dict):
        return Dict({k: batch_space(v, n) for k, v in space.items()})
    if isinstance(space, Discrete):
        return Box(low=0, high=space.n - 1, shape=(n,), dtype=np.int64)
    low, high = np.asarray(space.low), np.asarray(space.high)
    return Box(low=np.broadcast_to(low, (n,) + low.shape).copy(), high=np.broadcast_to(high, (n,) + high.shape).copy(),
               shape=(n,) + tuple(space.shape), dtype=space.dtype)


def _take(v, idx):
    """Rows `idx` of a batched value (device tensor, host tensor or array); anything unbatched is passed through."""
    if isinstance(v, torch.Tensor):
        return v[torch.as_tensor(idx, device=v.device)].clone() if v.dim() > 0 else v
    if isinstance(v, np.ndarray) and v.ndim > 0:
        return v[idx].copy()
    return v


class _EnvView:
    """What ``SyncVectorEnv.envs[i]`` is used for in the reference's scripts: one env's fovea state and record."""

    def __init__(self, vec, index: int):
        self._vec, self.index = vec, index

    @property
    def fov_loc(self):
        return self._vec.env.path.loc[self.index].cpu().numpy()

    @property
    def fov_res(self):
        return self._vec.env.path.res[self.index].cpu().numpy()

    def save_record_to_file(self, file_path: str):
        return self._vec.env.save_record_to_file(file_path, env_index=self.index)

    def episode_record(self):
        return self._vec.env.episode_record(self.index)


class FovealVectorEnv(_VectorBase):
    """``env``: a batched env from the factories (``AtariFixedFovealEnv(args, num_envs=N)`` ...).

    ``fix_sensory_space`` (default True): the reference declares ``sensory_action`` as
    ``Box(low=sas[0], high=sas[1], dtype=int)`` — shape (1,), and degenerate (low == high) in absolute mode
    (fov_env.py:125-129), so sampled actions are constant and one-dimensional.  With the flag the single-env space is
    the box the env actually accepts: shape (2,), ``[0, obs - fov]`` (absolute) or ``[lo, hi]`` (relative).  ``False``
    keeps the reference's declaration bit for bit."""

    def __init__(self, env, autoreset: bool = True, fix_sensory_space: bool = True):
        self.env = env
        self.num_envs = int(env.num_envs)
        self.autoreset = bool(autoreset)
        single_act = env.action_space
        if fix_sensory_space and isinstance(single_act, Dict) and "sensory_action" in single_act:
            single_act = Dict(dict(single_act))
            if env.sensory_action_mode == "absolute":
                lo, hi = np.zeros(2, np.int64), (np.array(env.obs_size) - np.array(env.fov_size)).astype(np.int64)
            else:
                lo = np.full(2, np.floor(env.sensory_action_space[0]), np.int64)
                hi = np.full(2, np.ceil(env.sensory_action_space[1]), np.int64)
            single_act["sensory_action"] = Box(low=lo, high=hi, shape=(2,), dtype=np.int64)
        self.single_action_space = single_act
        self.single_observation_space = env.observation_space
        self.action_space = batch_space(single_act, self.num_envs)
        self.observation_space = batch_space(env.observation_space, self.num_envs)
        self.envs = [_EnvView(self, i) for i in range(self.num_envs)]
        self.closed = False
        self._pending_reset = None

    # ---- reset
    def reset_async(self, seed=None, options=None):
        self._pending_reset = (seed, options)

    def reset_wait(self, seed=None, options=None):
        if self._pending_reset is not None:
            seed, options = self._pending_reset
            self._pending_reset = None
        mask = None if not options else options.get("reset_mask")
        if seed is not None:
            self.action_space.seed(seed) if hasattr(self.action_space, "seed") else None
        kind = getattr(self.env, "_kind", None)
        if kind is None:   # base env under RecordWrapper: reset(seed, options, mask=)
            return self.env.reset(seed, options, mask=mask)
        return self.env.reset(mask=mask)

    def reset(self, *, seed=None, options=None):
        self.reset_async(seed, options)
        return self.reset_wait()

    # ---- step
    def step_async(self, actions):
        self.env.step_async(actions)

    def step_wait(self):
        obs, reward, terminated, truncated, info = self.env.step_wait()
        terminated = np.asarray(terminated, bool)
        truncated = np.asarray(truncated, bool)
        finished = terminated | truncated
        if self.autoreset and finished.any():
            # SyncVectorEnv: the finished envs restart at once; their last observation / info travel in `info`
            idx = np.flatnonzero(finished)
            info["final_observation"] = _take(obs, idx)
            info["_final_observation"] = finished.copy()
            info["final_info"] = {k: _take(v, idx) for k, v in info.items() if not k.startswith(("final_", "_final_"))}
            info["_final_info"] = finished.copy()
            kind = getattr(self.env, "_kind", None)
            new_obs, new_info = (self.env.reset(None, None, mask=finished) if kind is None else self.env.reset(mask=finished))
            obs = new_obs   # the batched reset re-observes every env: unfinished envs keep their ring and fovea
            for k, v in new_info.items():
                info[k] = v
        return obs, reward, terminated, truncated, info

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    # ---- attribute plumbing of gymnasium.vector.VectorEnv
    def call(self, name: str, *args, **kwargs):
        attr = getattr(self.env, name)
        return attr(*args, **kwargs) if callable(attr) else attr

    def get_attr(self, name: str):
        return getattr(self.env, name)

    def set_attr(self, name: str, values):
        setattr(self.env, name, values)

    def close_extras(self, **kwargs):
        self.env.close()

    def close(self, **kwargs):
        if not self.closed:
            self.close_extras(**kwargs)
            self.closed = True

    @property
    def unwrapped(self):
        return self

    def __repr__(self):
        return f"FovealVectorEnv({type(self.env).__name__}, num_envs={self.num_envs})"


def enable_peer_access(src: torch.device, dst: torch.device) -> None:
    """Kernels running on ``src`` may dereference pointers into ``dst``'s memory afterwards (``cudaDeviceEnablePeerAccess``;
    over NVLink / NVSwitch on an HGX box).  Raises when the two GPUs cannot reach each other."""
    if src == dst:
        return
    if not torch.cuda.can_device_access_peer(src.index, dst.index):
        raise RuntimeError(f"{src} cannot access {dst} as a peer")
    from .pipeline import _rt
    rt = _rt()
    with torch.cuda.device(src):
        torch.cuda.current_stream(src).synchronize()      # make sure the context exists
        err = rt.cudaDeviceEnablePeerAccess(int(dst.index), 0)
        if err not in (0, 704):                           # 704 = cudaErrorPeerAccessAlreadyEnabled
            raise RuntimeError(f"cudaDeviceEnablePeerAccess({src} -> {dst}) failed: cudaError {err}")
        if err == 704:
            rt.cudaGetLastError()                         # clear the sticky "already enabled"


class ShardedVecEnv:
    """``num_envs`` environments over ``devices`` (default: every visible GPU): rank-ordered contiguous env blocks,
    one batched env per device, no collective.  ``make_env(n, device, lo, hi)`` builds the env of one block
    (e.g. ``lambda n, dev, lo, hi: AtariFixedFovealPeripheralEnv(args, num_envs=n, source=make_source(lo, hi), device=dev)``).

    Without ``learner_device`` observations and device-resident ``info`` entries come back as one tensor per device
    (``ShardedVecEnv.gather`` concatenates them on one device with peer copies when a learner wants them together);
    host arrays (reward, done) are concatenated in env order.

    ``learner_device``: the global observation batch ``(num_envs, K, h, w)`` lives on that GPU and EVERY device's observe
    kernel stores its env block straight into it through peer memory — the kernels' own stores (TMA bulk stores for
    the peripheral and flexible kernels) travel over NVLink / NVSwitch while the kernel is still computing the next
    envs, so there is no gather pass and no second copy of the observations.  ``step`` / ``reset`` then return that one
    tensor; the learner's current stream waits (device side) for every device's step."""

    def __init__(self, make_env: Callable, num_envs: int, devices: Optional[Sequence] = None, learner_device=None):
        if devices is None:
            devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
        if not devices:
            raise RuntimeError("ShardedVecEnv needs at least one CUDA device; there is no CPU fallback")
        self.num_envs = int(num_envs)
        self.devices = [torch.device(d) for d in devices]
        self.ranges = [r for r in all_shards(self.num_envs, len(self.devices)) if r[1] > r[0]]
        self.envs = [make_env(hi - lo, self.devices[i], lo, hi) for i, (lo, hi) in enumerate(self.ranges)]
        self.action_space = self.envs[0].action_space
        self.observation_space = self.envs[0].observation_space
        self.learner = None if learner_device is None else torch.device(learner_device)
        self.obs = None
        if self.learner is not None:
            if self.learner.type != "cuda":
                raise RuntimeError("learner_device must be a CUDA device")
            if self.learner.index is None:
                self.learner = torch.device("cuda", torch.cuda.current_device())
            if not all(hasattr(e, "set_output") for e in self.envs):
                raise TypeError("learner_device needs foveal envs (FixedFovealEnv and subclasses): they own the observation tensor")
            shape = (self.num_envs,) + tuple(self.envs[0].obs_shape[1:])
            self.obs = torch.zeros(shape, dtype=torch.uint8, device=self.learner)
            torch.cuda.current_stream(self.learner).synchronize()   # the fill is done before any peer writes into it
            for e, (lo, hi), d in zip(self.envs, self.ranges, self.devices):
                d = d if d.index is not None else torch.device("cuda", torch.cuda.current_device())
                enable_peer_access(d, self.learner)
                e.set_output(self.obs[lo:hi])
            self._evs = [torch.cuda.Event() for _ in self.envs]

    def _learner_waits(self):
        """The learner's current stream waits for what every device has enqueued so far (cross-device event waits)."""
        cur = torch.cuda.current_stream(self.learner)
        for e, ev in zip(self.envs, self._evs):
            dev = e.path.device
            ev.record(torch.cuda.current_stream(dev))
            cur.wait_event(ev)

    def _split(self, actions, i):
        lo, hi = self.ranges[i]
        if isinstance(actions, dict):
            return {k: v[lo:hi] for k, v in actions.items()}
        return actions[lo:hi]

    @staticmethod
    def _merge_info(infos):
        out = {}
        for k in infos[0]:
            vals = [inf[k] for inf in infos]
            out[k] = vals if isinstance(vals[0], torch.Tensor) and vals[0].is_cuda else np.concatenate([np.asarray(v) for v in vals])
        return out

    def reset(self, mask=None):
        res = [e.reset(mask=None if mask is None else np.asarray(mask)[lo:hi]) for e, (lo, hi) in zip(self.envs, self.ranges)]
        if self.learner is not None:
            self._learner_waits()
            return self.obs, self._merge_info([r[1] for r in res])
        return [r[0] for r in res], self._merge_info([r[1] for r in res])

    def step_async(self, actions):
        for i, e in enumerate(self.envs):
            e.step_async(self._split(actions, i))

    def step_wait(self):
        res = [e.step_wait() for e in self.envs]
        obs = [r[0] for r in res]
        if self.learner is not None:
            self._learner_waits()
            obs = self.obs
        reward = np.concatenate([np.asarray(r[1]) for r in res])
        done = np.concatenate([np.asarray(r[2]) for r in res])
        trunc = np.concatenate([np.asarray(r[3]) for r in res])
        return obs, reward, done, trunc, self._merge_info([r[4] for r in res])

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    @staticmethod
    def gather(parts: List[torch.Tensor], device) -> torch.Tensor:
        """Concatenates per-device tensors on ``device`` (peer copies over NVLink / NVSwitch)."""
        device = torch.device(device)
        return torch.cat([p.to(device, non_blocking=True) for p in parts], dim=0)

    def close(self):
        for e in self.envs:
            e.close()

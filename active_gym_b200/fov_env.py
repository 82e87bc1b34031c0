"""Active-vision wrappers over a batched base env: ``RecordWrapper``, ``FixedFovealEnv``,
``FlexibleFovealEnv``, ``FixedFovealPeripheralEnv`` — same names, constructor signature
``(env, args)``, ``reset`` / ``step`` protocol, Dict action space and ``info`` keys as
``active_gym/fov_env.py`` (citations refer to it), with a leading N axis on every array.

The pixel work of ``_fov_step`` / ``_get_fov_state`` runs on the GPU (``ObservationPath``);
these classes keep only the host-side bookkeeping of the reference.
"""
from __future__ import annotations

from enum import IntEnum

import numpy as np
import torch

from . import _lib
from .spaces import Box, Dict, Discrete, Wrapper


class RecordWrapper(Wrapper):
    """Episode counters of fov_env.py:15-67, batched: ``info["reward"]`` is the cumulative raw
    reward and ``info["ep_len"]`` the step count of every env.  Trajectory recording to
    mp4/.pt (fov_env.py:70-102) is a debugging aid outside the hot path and is not provided."""

    def __init__(self, env, args):
        super().__init__(env)
        self.args = args
        self.record = bool(args.record)
        if self.record:
            raise NotImplementedError("record=True (mp4/.pt trajectory dumps, fov_env.py:70-102) is out of scope")
        n = env.num_envs
        self.cumulative_reward = np.zeros(n, np.float64)
        self.ep_len = np.zeros(n, np.int64)

    def _add_info(self, info):
        info["reward"] = self.cumulative_reward.copy()
        info["ep_len"] = self.ep_len.copy()
        return info

    def reset(self, seed=None, options=None, mask=None, return_state=True):
        state, info = self.env.reset(seed, options, mask=mask, return_state=return_state)
        sel = slice(None) if mask is None else np.asarray(mask, bool)
        self.cumulative_reward[sel] = 0
        self.ep_len[sel] = 0
        return state, self._add_info(info)

    def step(self, action, return_state=True):
        state, return_reward, done, truncated, info = self.env.step(action, return_state=return_state)
        self.ep_len += 1
        self.cumulative_reward += np.asarray(info.get("raw_reward", return_reward), np.float64)
        return state, return_reward, done, truncated, self._add_info(info)


class FixedFovealEnv(Wrapper):
    """fov_env.py:107-234 for N envs.  Observation: uint8 CUDA tensor (N, K, f_h, f_w), or
    (N, K, S_h, S_w) with ``mask_out`` / ``resize_to_full``."""

    _kind = "fixed"

    def __init__(self, env, args):
        super().__init__(env)
        self.fov_size = tuple(args.fov_size)
        self.fov_init_loc = tuple(args.fov_init_loc)
        assert (np.array(self.fov_size) < np.array(self.obs_size)).all()  # fov_env.py:112
        self.sensory_action_mode = args.sensory_action_mode
        if self.sensory_action_mode == "relative":
            self.sensory_action_space = np.array(args.sensory_action_space)
        elif self.sensory_action_mode == "absolute":
            self.sensory_action_space = np.array(self.obs_size) - np.array(self.fov_size)
        else:
            raise ValueError(self.sensory_action_mode)
        self.resize = bool(getattr(args, "resize_to_full", False))
        self.mask_out = bool(args.mask_out)
        base = self.env.unwrapped
        self.path = base.path
        self.num_envs = base.num_envs
        if self.path.fov_size != self.fov_size:
            raise ValueError("the base env was built from different args (fov_size mismatch)")
        # fov_env.py:125-129 declares Box(low=sas[0], high=sas[1], dtype=int): shape (1,) and, in
        # absolute mode, degenerate.  Kept for compatibility; step() accepts any real (N,2) array.
        self.action_space = Dict({
            "motor_action": self.env.action_space,
            "sensory_action": Box(low=self.sensory_action_space[0], high=self.sensory_action_space[1], dtype=int),
        })
        shape = self.obs_size if (self.mask_out or self.resize) else self.fov_size
        self.observation_space = Box(low=-1., high=1., shape=(self.env.frame_stack,) + tuple(shape), dtype=np.float32)

    # ---- state the reference exposes as attributes
    @property
    def fov_loc(self) -> torch.Tensor:
        return self.path.loc

    @property
    def variant(self) -> str:
        return "mask" if self.mask_out else ("resize_full" if self.resize else "crop")

    def _observe(self, action, ctrl, action_type=None):
        return self.path.observe_fixed(action, variant=self.variant, ctrl=ctrl)

    def _fov_info(self, info):
        info["fov_loc"] = self.path.loc.clone()
        return info

    @staticmethod
    def _ctrl_for_reset(mask, n):
        if mask is None:
            return "reset"
        return np.where(np.asarray(mask, bool), _lib.FOV_RESET, _lib.FOV_KEEP).astype(np.uint8)

    def reset(self, mask=None):
        """fov_env.py:156-164 (takes no seed/options, like the reference)."""
        _, info = self.env.reset(mask=mask, return_state=False)
        obs = self._observe(None, self._ctrl_for_reset(mask, self.num_envs))
        return obs, self._fov_info(info)

    def step(self, action):
        """fov_env.py:209-221.  action = {"motor_action": (N,), "sensory_action": (N,2)}."""
        _, reward, done, truncated, info = self.env.step(action["motor_action"], return_state=False)
        obs = self._observe(action["sensory_action"], None, action.get("sensory_action_type"))
        return obs, reward, done, truncated, self._fov_info(info)


class FlexibleFovealEnvActionType(IntEnum):  # fov_env.py:236-238
    FOV_LOC = 0
    FOV_RES = 1


class FlexibleFovealEnv(FixedFovealEnv):
    """fov_env.py:240-355 for N envs.  Without mask_out / resize_to_full the reference returns a
    variable-shape (K, rh, rw) crop; the batched form writes it into the top-left corner of a
    zeroed (N, K, S_h, S_w) tensor and reports the shape in ``info["fov_res"]``."""

    _kind = "flexible"

    def __init__(self, env, args):
        super().__init__(env, args)
        self.action_space["sensory_action_type"] = Discrete(len(FlexibleFovealEnvActionType))
        self.fov_init_res = tuple(args.fov_size)
        self.validate_actions = bool(getattr(args, "validate_actions", True))

    @property
    def fov_res(self) -> torch.Tensor:
        return self.path.res

    def _check_res(self, action, action_type):
        """The reference raises for a window larger than the frame; do it before the launch
        when the action lives on the host (a device tensor would need a sync)."""
        if not self.validate_actions or action is None or isinstance(action, torch.Tensor):
            return
        a = np.asarray(action, np.float64).reshape(self.num_envs, 2)
        t = np.zeros(self.num_envs, np.int64) if action_type is None else \
            np.asarray(action_type.cpu() if isinstance(action_type, torch.Tensor) else action_type).reshape(-1)
        sel = t == int(FlexibleFovealEnvActionType.FOV_RES)
        if sel.any():
            r = a[sel]
            if (r < 1).any() or (r > np.array(self.obs_size)).any() or (r != np.floor(r)).any():
                raise ValueError("FOV_RES actions must be integer window sizes within [1, obs_size]")

    def _observe(self, action, ctrl, action_type=None):
        self._check_res(action, action_type)
        return self.path.observe_flexible(action, action_type, variant=self.variant, ctrl=ctrl)

    def _fov_info(self, info):
        info["fov_loc"] = self.path.loc.clone()
        info["fov_res"] = self.path.res.clone()
        return info


class FixedFovealPeripheralEnv(FixedFovealEnv):
    """fov_env.py:358-388 for N envs: blurred periphery (squeeze to ``peripheral_res``, expand
    back) with the full-resolution fovea pasted at ``fov_loc``; (N, K, S_h, S_w) uint8."""

    _kind = "peripheral"

    def __init__(self, env, args):
        super().__init__(env, args)
        self.mask_out = False
        self.resize_to_full = True
        self.peripheral_res = tuple(args.peripheral_res)
        self.observation_space = Box(low=-1., high=1., shape=(self.env.frame_stack,) + tuple(self.obs_size), dtype=np.float32)
        if self.path.peripheral_res != self.peripheral_res:
            raise ValueError("the base env was built from different args (peripheral_res mismatch)")

    def _observe(self, action, ctrl, action_type=None):
        return self.path.observe_peripheral(action, ctrl=ctrl)


class SingleEnvAdapter:
    """The N=1 drop-in: presents a batched env with the reference's single-env types — float64
    NumPy observations normalised to [0,1] (= float32(u8)/255), scalar reward / done, NumPy
    ``fov_loc`` — so code written against the reference runs unchanged."""

    def __init__(self, env):
        assert env.num_envs == 1
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def fov_loc(self):
        return self.env.path.loc[0].cpu().numpy()

    @property
    def fov_res(self):
        return self.env.path.res[0].cpu().numpy()

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def _obs(self, obs):
        o = obs[0].cpu().numpy()
        if getattr(self.env, "_kind", "") == "flexible" and self.env.variant == "crop":
            rh, rw = self.fov_res
            o = o[:, :rh, :rw]
        return (o.astype(np.float32) / np.float32(255.0)).astype(np.float64)

    @staticmethod
    def _unbatch(info):
        out = {}
        for k, v in info.items():
            if isinstance(v, torch.Tensor):
                v = v.cpu().numpy()
            v = np.asarray(v)
            v = v[0] if v.ndim >= 1 and v.shape[0] == 1 else v
            out[k] = v.item() if isinstance(v, np.ndarray) and v.ndim == 0 else v
        return out

    def reset(self, seed=None, options=None):
        wrapped_fov = isinstance(self.env, FixedFovealEnv)
        obs, info = self.env.reset() if wrapped_fov else self.env.reset(seed, options)
        return self._obs(obs), self._unbatch(info)

    def step(self, action):
        if isinstance(action, dict):
            action = dict(action)
            sa = action["sensory_action"]
            sa = sa.detach().cpu().numpy() if isinstance(sa, torch.Tensor) else np.asarray(sa)
            action["sensory_action"] = sa.reshape(1, 2)
            action["motor_action"] = np.asarray(action["motor_action"]).reshape(1, *np.shape(action["motor_action"]))
            if "sensory_action_type" in action:
                t = action["sensory_action_type"]
                t = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
                action["sensory_action_type"] = t.reshape(-1)[:1].astype(np.int64)
        else:
            action = np.asarray(action).reshape(1, *np.shape(action))
        obs, reward, done, truncated, info = self.env.step(action)
        return (self._obs(obs), np.asarray(reward).reshape(-1)[0].item(), bool(np.asarray(done).reshape(-1)[0]),
                False, self._unbatch(info))

    def train(self):
        self.env.train()

    def eval(self):
        self.env.eval()

    def close(self):
        self.env.close()

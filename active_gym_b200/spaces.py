"""gymnasium when it is installed, otherwise duck-typed stand-ins with the same surface.

The reference builds its envs on ``gymnasium.Env`` / ``gymnasium.Wrapper`` and the ``Box`` /
``Discrete`` / ``Dict`` spaces (fov_env.py:8-9, atari_env.py:10-11).  This image has no
gymnasium, and neither has the GPU box, so the env layer binds to whichever is available.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - exercised only where gymnasium exists
    import gymnasium as _gym
    from gymnasium.spaces import Box, Dict, Discrete  # noqa: F401

    Env, Wrapper, HAVE_GYMNASIUM = _gym.Env, _gym.Wrapper, True
except Exception:  # ImportError, or a broken install
    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self, shape=None, dtype=None):
            self.shape, self.dtype = shape, dtype
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low) if np.ndim(low) > 0 else (1,)
            super().__init__(tuple(shape), np.dtype(dtype))
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            if np.issubdtype(self.dtype, np.integer):
                return self._rng.integers(lo.astype(np.int64), hi.astype(np.int64) + 1).astype(self.dtype)
            return (lo + (hi - lo) * self._rng.random(self.shape)).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete(_Space):
        def __init__(self, n):
            super().__init__((), np.dtype(np.int64))
            self.n = int(n)

        def contains(self, x):
            return 0 <= int(x) < self.n

        def sample(self):
            return int(self._rng.integers(self.n))

        def __repr__(self):
            return f"Discrete({self.n})"

    class Dict(dict):
        def __init__(self, spaces=None, **kw):
            super().__init__(spaces or {}, **kw)

        def sample(self):
            return {k: v.sample() for k, v in self.items()}

        def seed(self, seed=None):
            for v in self.values():
                v.seed(seed)

        def contains(self, x):
            return all(k in x and v.contains(x[k]) for k, v in self.items())

    class Env:
        metadata = {"render_modes": []}
        render_mode = None
        spec = None

        def reset(self, seed=None, options=None):
            raise NotImplementedError

        def step(self, action):
            raise NotImplementedError

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_") or name == "env":
                raise AttributeError(name)
            return getattr(self.env, name)

        def reset(self, *a, **k):
            return self.env.reset(*a, **k)

        def step(self, action):
            return self.env.step(action)

        def close(self):
            return self.env.close()

        @property
        def unwrapped(self):
            return self.env.unwrapped

"""Frame sources: the seam BELOW the observation path.

The simulators stay on the host (``BASELINE.json:north_star``): ALE and MuJoCo step per env on
CPU cores and hand batched raw frames to the GPU through pinned staging buffers.  A source
produces, per call, exactly what the kernels consume: raw frames + per-env ingest flags, and
the simulator-side scalars (reward, done).

* ``ALEPool``      — N ``atari_py.ALEInterface`` objects driven with the reference's simulator-side
                     logic: no-op + fire reset, action repeat with frames taken at t==2/t==3,
                     episodic-life termination (atari_env.py:84-108, 123-131, 135-140).
* ``DMCPool``      — N ``dm_control`` tasks (dmc_env.py:100-106, 166-173, 202-224).
* ``SyntheticAtariSource`` / ``SyntheticDMCSource`` — device-generated frames for benchmarks.
* ``PinnedFrameSource`` — pre-generated frames in pinned HOST memory, cycled: the stand-in for a simulator pool
                     whose stepping cost is not what is being measured (end-to-end benchmark, pipeline tests).

Staging protocol (host sources): a source owns TWO sets of pinned staging buffers and alternates between them on
every ``reset`` / ``step`` call, and the env calls ``source.frames_consumed(events)`` with CUDA events recorded
(one per copy stream) after the asynchronous host->device copies of a call were enqueued; before a set is written
again the source waits for the events of the call that last used it.  So an in-flight copy is never overwritten.
Row packing: ``set_used_rows(rows)`` (called by the env with ``ObservationPath.used_rows``) makes an Atari source
write only the raw rows the resize samples — (N, 168, 160) instead of (N, 210, 160) — so that the copy is one
contiguous block per shard; ``packed_rows`` tells the env which layout it gets.
"""
from __future__ import annotations

import random
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from ._lib import FLAG_FRAME_A, FLAG_FRAME_B, FLAG_HARD_RESET, FLAG_IDLE


class _Workers:
    """Runs ``fn(i)`` for every env index on ``workers`` host threads (contiguous index blocks).  The
    simulators are C/C++ libraries behind ctypes / pybind calls that release the GIL, so env stepping
    scales with host cores; every env writes only its own rows of the pinned staging buffers."""

    def __init__(self, num_envs: int, workers: int):
        self.num_envs = int(num_envs)
        self.workers = max(1, min(int(workers), self.num_envs))
        self.pool = None
        if self.workers > 1:
            from concurrent.futures import ThreadPoolExecutor
            self.pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="agym-sim")
            b = np.linspace(0, self.num_envs, self.workers + 1).astype(int)
            self.blocks = [(int(lo), int(hi)) for lo, hi in zip(b[:-1], b[1:]) if hi > lo]

    def run(self, fn) -> None:
        if self.pool is None:
            for i in range(self.num_envs):
                fn(i)
            return

        def block(r):
            for i in range(*r):
                fn(i)
        for f in [self.pool.submit(block, r) for r in self.blocks]:
            f.result()  # re-raises a worker's exception

    def close(self) -> None:
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None


def _pinned(shape, dtype=torch.uint8) -> torch.Tensor:
    t = torch.zeros(shape, dtype=dtype)
    if torch.cuda.is_available():
        t = t.pin_memory()
    return t


class _Staging:
    """Two alternating sets of pinned buffers + the events that guard them (see the module docstring)."""

    def _init_staging(self, make_set):
        self._sets = [make_set(), make_set()]
        self._events = [None, None]
        self._cur = 0

    def _next_set(self):
        self._cur ^= 1
        for ev in self._events[self._cur] or ():
            ev.synchronize()
        self._events[self._cur] = None
        return self._sets[self._cur]

    def frames_consumed(self, events) -> None:
        """The env recorded `events` (one per copy stream) after enqueueing the copies that read the set handed out last."""
        self._events[self._cur] = list(events)


class ALEPool(_Staging):
    """Host pool of Arcade Learning Environment instances (the L0 simulators of SURVEY.md §1)."""

    raw_shape = (210, 160, 1)

    def __init__(self, args, num_envs: int, ale_factory: Optional[Callable[[int], object]] = None, workers: int = 1):
        self.num_envs = int(num_envs)
        self.action_repeat = int(args.action_repeat)
        self.training = True
        self._workers = _Workers(self.num_envs, workers)
        if ale_factory is None:
            import atari_py  # third-party; absent in this image

            def ale_factory(i):  # atari_env.py:44-50
                ale = atari_py.ALEInterface()
                ale.setInt("random_seed", args.seed + i)
                ale.setInt("max_num_frames_per_episode", int(args.max_episode_length))
                ale.setFloat("repeat_action_probability", 0)
                ale.setInt("frame_skip", 0)
                ale.setBool("color_averaging", False)
                ale.loadROM(atari_py.get_game_path(args.game))
                return ale
        self.ales = [ale_factory(i) for i in range(self.num_envs)]
        actions = self.ales[0].getMinimalActionSet()
        self.actions = dict(enumerate(actions))  # atari_env.py:51-52
        self.n_actions = len(actions)
        self.lives = [0] * self.num_envs
        self.life_termination = [False] * self.num_envs
        self.packed_rows = False
        self._rows = None
        self._alloc(210)
        self.flags = np.zeros(self.num_envs, np.uint8)

    def _alloc(self, rows: int):
        n = self.num_envs
        self._init_staging(lambda: (_pinned((n, rows, 160)), _pinned((n, rows, 160))))
        self.frames_a, self.frames_b = self._sets[0]
        self._fa, self._fb = self.frames_a.numpy(), self.frames_b.numpy()

    def set_used_rows(self, rows) -> None:
        """Write only these raw rows (ascending) of every screen: the rows cv2's resize samples (atari_env.py:74)."""
        self._rows = np.asarray(rows, np.int64)
        self.packed_rows = True
        self._alloc(len(self._rows))

    def _flip(self):
        self.frames_a, self.frames_b = self._next_set()
        self._fa, self._fb = self.frames_a.numpy(), self.frames_b.numpy()

    def _screen(self, ale, dst):
        g = np.asarray(ale.getScreenGrayscale()).reshape(210, 160)
        dst[...] = g if self._rows is None else g[self._rows]

    def reset(self, mask: Optional[np.ndarray] = None):
        """atari_env.py:84-113 minus the buffer work.  Returns (frames, frames, flags)."""
        # the reference draws its no-op count from the global `random` (atari_env.py:96); draw them here,
        # serially and in env order, so that worker threads do not change the stream
        self._flip()
        noops = [None] * self.num_envs
        for i in range(self.num_envs):
            if (mask is None or mask[i]) and not self.life_termination[i]:
                noops[i] = random.randrange(30)

        def one(i):
            ale = self.ales[i]
            if mask is not None and not mask[i]:
                self.flags[i] = FLAG_IDLE
                return
            if self.life_termination[i]:
                self.life_termination[i] = False
                ale.act(0)
                self.flags[i] = FLAG_FRAME_A
            else:
                ale.reset_game()
                for _ in range(noops[i]):
                    ale.act(0)
                    if ale.game_over():
                        ale.reset_game()
                self.flags[i] = FLAG_FRAME_A | FLAG_HARD_RESET
            if self.n_actions >= 3:
                ale.act(1)
                if ale.game_over():
                    ale.reset_game()
                    ale.act(2)
                if ale.game_over():
                    ale.reset_game()
            self._screen(ale, self._fa[i])
            self.lives[i] = ale.lives()
        self._workers.run(one)
        return self.frames_a, self.frames_a, self.flags

    def step(self, motor_action: Sequence[int]):
        """atari_env.py:119-140 minus the buffer work.  Returns (fa, fb, flags, reward, done)."""
        n = self.num_envs
        self._flip()
        reward, done = np.zeros(n, np.float64), np.zeros(n, bool)
        motor_action = np.asarray(motor_action).reshape(n)

        def one(i):
            ale = self.ales[i]
            fl, r, d = 0, 0, False
            for t in range(self.action_repeat):
                r += ale.act(self.actions.get(int(motor_action[i])))
                if t == 2:
                    self._screen(ale, self._fa[i]); fl |= FLAG_FRAME_A
                elif t == 3:
                    self._screen(ale, self._fb[i]); fl |= FLAG_FRAME_B
                d = ale.game_over()
                if d:
                    break
            if self.training:
                lives = ale.lives()
                if lives < self.lives[i] and lives > 0:
                    self.life_termination[i] = not d
                    d = True
                self.lives[i] = lives
            self.flags[i], reward[i], done[i] = fl, r, d
        self._workers.run(one)
        return self.frames_a, self.frames_b, self.flags, reward, done

    def close(self):
        self._workers.close()


class DMCPool(_Staging):
    """Host pool of dm_control tasks; renders at obs_size (dmc_env.py:175-180)."""

    def __init__(self, args, num_envs: int, env_factory: Optional[Callable[[int], object]] = None, workers: int = 1):
        self.num_envs, self.action_repeat = int(num_envs), int(args.action_repeat)
        self._workers = _Workers(self.num_envs, workers)
        self.obs_size, self.camera_id = tuple(args.obs_size), args.camera_id
        if env_factory is None:
            from dm_control import suite  # third-party; absent in this image

            def env_factory(i):  # dmc_env.py:92-106
                kw = dict(args.task_kwargs); kw["random"] = args.seed + i
                return suite.load(domain_name=args.domain_name, task_name=args.task_name, task_kwargs=kw,
                                  visualize_reward=args.visualize_reward, environment_kwargs=args.environment_kwargs)
        self.envs = [env_factory(i) for i in range(self.num_envs)]
        spec = self.envs[0].action_spec()
        self.true_low = np.asarray(spec.minimum, np.float32) + np.zeros(spec.shape, np.float32)
        self.true_high = np.asarray(spec.maximum, np.float32) + np.zeros(spec.shape, np.float32)
        h, w = self.obs_size
        self.raw_shape = (h, w, 3)
        self._init_staging(lambda: _pinned((self.num_envs, h, w, 3)))
        self.frames = self._sets[0]
        self._f = self.frames.numpy()
        self.flags = np.zeros(self.num_envs, np.uint8)
        self.last_time_steps = [None] * self.num_envs

    def _convert_action(self, a):  # dmc_env.py:166-173, norm space is [-1, 1]
        a = a.astype(np.float64)
        a = (a - (-1.0)) / 2.0
        return (a * (self.true_high - self.true_low) + self.true_low).astype(np.float32)

    def _render(self, i):
        h, w = self.obs_size
        self._f[i] = self.envs[i].physics.render(height=h, width=w, camera_id=self.camera_id)

    def _flip(self):
        self.frames = self._next_set()
        self._f = self.frames.numpy()

    def reset(self, mask=None):
        self._flip()

        def one(i):
            if mask is not None and not mask[i]:
                self.flags[i] = FLAG_IDLE
                return
            self.last_time_steps[i] = self.envs[i].reset()
            self._render(i)
            self.flags[i] = FLAG_FRAME_A | FLAG_HARD_RESET
        self._workers.run(one)
        return self.frames, self.flags

    def extra_info(self):
        """dmc_env.py:188-191: MuJoCo state and discount ride along in ``info``."""
        return {"internal_state": np.stack([np.asarray(e.physics.get_state()).copy() for e in self.envs]),
                "discount": np.array([getattr(ts, "discount", None) for ts in self.last_time_steps], dtype=object)}

    def step(self, motor_action):
        n = self.num_envs
        reward, done = np.zeros(n, np.float64), np.zeros(n, bool)
        motor_action = np.asarray(motor_action, np.float32).reshape(n, -1)
        self._flip()

        def one(i):
            env = self.envs[i]
            a = self._convert_action(motor_action[i])
            r = 0
            for _ in range(self.action_repeat):  # dmc_env.py:218-223
                ts = env.step(a)
                r += ts.reward or 0
                if ts.last():
                    done[i] = True
                    break
            self.last_time_steps[i] = ts
            self._render(i)
            reward[i], self.flags[i] = r, FLAG_FRAME_A
        self._workers.run(one)
        return self.frames, self.flags, reward, done

    def close(self):
        self._workers.close()


class SyntheticAtariSource:
    """Frames generated on the device (or given as a pool of pre-generated batches): the
    benchmark / test stand-in for ``ALEPool``.  Never terminates; reward is 0."""

    def __init__(self, num_envs: int, channels: int = 1, device=None, pool: int = 4, seed: int = 1234):
        self.num_envs, self.n_actions = int(num_envs), 18
        self.raw_shape = (210, 160, channels)
        self.device = torch.device(device if device is not None else "cuda")
        shape = (num_envs, 210, 160) if channels == 1 else (num_envs, 210, 160, 3)
        g = torch.Generator(device=self.device).manual_seed(seed)
        self.batches = [torch.randint(0, 256, shape, dtype=torch.uint8, device=self.device, generator=g)
                        for _ in range(max(2, pool))]
        self.t = 0
        self.flags_step = torch.full((num_envs,), FLAG_FRAME_A | FLAG_FRAME_B, dtype=torch.uint8, device=self.device)
        self.flags_reset = torch.full((num_envs,), FLAG_FRAME_A | FLAG_HARD_RESET, dtype=torch.uint8, device=self.device)
        self._zero = np.zeros(num_envs)
        self._false = np.zeros(num_envs, bool)

    def _next(self):
        b = self.batches[self.t % len(self.batches)]
        self.t += 1
        return b

    def reset(self, mask=None):
        f = self._next()
        flags = self.flags_reset
        if mask is not None:
            m = torch.as_tensor(np.asarray(mask, bool), device=self.device)
            flags = torch.where(m, flags, torch.full_like(flags, FLAG_IDLE))
        return f, f, flags

    def step(self, motor_action):
        return self._next(), self._next(), self.flags_step, self._zero, self._false


class SyntheticDMCSource:
    def __init__(self, num_envs: int, obs_size=(84, 84), device=None, pool: int = 4, seed: int = 1234, action_dim: int = 2):
        self.num_envs, self.action_dim = int(num_envs), action_dim
        self.raw_shape = (obs_size[0], obs_size[1], 3)
        self.device = torch.device(device if device is not None else "cuda")
        g = torch.Generator(device=self.device).manual_seed(seed)
        self.batches = [torch.randint(0, 256, (num_envs,) + self.raw_shape, dtype=torch.uint8, device=self.device, generator=g)
                        for _ in range(max(2, pool))]
        self.t = 0
        self.flags_step = torch.full((num_envs,), FLAG_FRAME_A, dtype=torch.uint8, device=self.device)
        self.flags_reset = torch.full((num_envs,), FLAG_FRAME_A | FLAG_HARD_RESET, dtype=torch.uint8, device=self.device)
        self.true_low, self.true_high = -np.ones(action_dim, np.float32), np.ones(action_dim, np.float32)
        self._zero, self._false = np.zeros(num_envs), np.zeros(num_envs, bool)

    def _next(self):
        b = self.batches[self.t % len(self.batches)]
        self.t += 1
        return b

    def reset(self, mask=None):
        flags = self.flags_reset
        if mask is not None:
            m = torch.as_tensor(np.asarray(mask, bool), device=self.device)
            flags = torch.where(m, flags, torch.full_like(flags, FLAG_IDLE))
        return self._next(), flags

    def step(self, motor_action):
        return self._next(), self.flags_step, self._zero, self._false


class PinnedFrameSource:
    """Pre-generated frames in pinned host memory, cycled (``pool`` batches): what a host simulator pool hands over,
    minus the simulation.  ``kind`` "atari" (two screens per step, gray or RGB; only the sampled rows once the env
    has called ``set_used_rows``) or "dmc" (one render per step).  Never terminates unless ``done_every`` is set
    (then env i reports done — and a raw reward of 1 — every ``done_every`` + i % 7 steps: exercises the counters)."""

    def __init__(self, num_envs: int, kind: str = "atari", channels: int = 1, obs_size=(84, 84), pool: int = 2,
                 seed: int = 7, done_every: int = 0, action_dim: int = 2):
        self.num_envs, self.kind, self.n_actions, self.action_dim = int(num_envs), kind, 18, action_dim
        self.raw_shape = (210, 160, channels) if kind == "atari" else (obs_size[0], obs_size[1], 3)
        self.pool, self.seed, self.done_every = max(2, int(pool)), seed, int(done_every)
        self.packed_rows, self._rows = False, None
        self.true_low, self.true_high = -np.ones(action_dim, np.float32), np.ones(action_dim, np.float32)
        self.t = 0
        self._steps = np.zeros(self.num_envs, np.int64)
        self._build()

    def _build(self):
        h, w, c = self.raw_shape
        rows = h if self._rows is None else len(self._rows)
        shape = (self.num_envs, rows, w) if c == 1 else (self.num_envs, rows, w, c)
        rng = np.random.default_rng(self.seed)
        per_step = 2 if self.kind == "atari" else 1
        self.batches = []
        for _ in range(self.pool * per_step):
            t = _pinned(shape)
            t.numpy()[...] = rng.integers(0, 256, shape, dtype=np.uint8)
            self.batches.append(t)

    def set_used_rows(self, rows) -> None:
        if self.kind != "atari":
            return
        self._rows = np.asarray(rows, np.int64)
        self.packed_rows = True
        self._build()

    def frames_consumed(self, events) -> None:  # the pool is read-only: nothing to guard
        pass

    def _next(self):
        b = self.batches[self.t % len(self.batches)]
        self.t += 1
        return b

    def _flags(self, value, mask):
        f = np.full(self.num_envs, value, np.uint8)
        if mask is not None:
            f[~np.asarray(mask, bool)] = FLAG_IDLE
        return f

    def reset(self, mask=None):
        sel = slice(None) if mask is None else np.asarray(mask, bool)
        self._steps[sel] = 0
        f = self._next()
        fl = self._flags(FLAG_FRAME_A | FLAG_HARD_RESET, mask)
        return (f, f, fl) if self.kind == "atari" else (f, fl)

    def step(self, motor_action):
        self._steps += 1
        n = self.num_envs
        if self.done_every:
            done = self._steps >= self.done_every + (np.arange(n) % 7)
            reward = done.astype(np.float64)
        else:
            done, reward = np.zeros(n, bool), np.zeros(n, np.float64)
        if self.kind == "atari":
            return self._next(), self._next(), self._flags(FLAG_FRAME_A | FLAG_FRAME_B, None), reward, done
        return self._next(), self._flags(FLAG_FRAME_A, None), reward, done

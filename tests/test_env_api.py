"""The env layer (host logic) — CPU tests drive the simulator pool with the scripted fake ALE /
fake physics of oracle/ref_harness.py and compare against what the unmodified reference did
on the same script (tests/golden); the drop-in env tests run twice: on the GPU box through the CUDA
path (backend "cuda"), and in the CPU-only container with the C oracle standing in for the kernels
(backend "oracle", tests/fake_path.py) so that the Python host logic is covered without a GPU."""
import random

import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from tests import golden_replay as gr

BACKENDS = ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)]


def _backend(monkeypatch, backend):
    """"oracle": the env classes build tests/fake_path.OraclePath instead of the CUDA engine."""
    if backend == "oracle":
        from tests.fake_path import OraclePath
        monkeypatch.setattr("active_gym_b200.atari_env.PipelinedPath", OraclePath)


ATARI = [s for s in gr.SCENARIOS if s.startswith("atari")]
DMC = [s for s in gr.SCENARIOS if s.startswith("dmc")]


def _atari_args(meta, **extra):
    from active_gym_b200 import AtariEnvArgs
    kw = dict(fov_size=tuple(meta["fov_size"]), fov_init_loc=tuple(meta["fov_init_loc"]),
              sensory_action_mode=meta["mode"], frame_stack=meta["frame_stack"], action_repeat=meta["action_repeat"],
              mask_out=meta["variant"] == "mask", resize_to_full=meta["variant"] == "resize_full")
    if meta["mode"] == "relative":
        kw["sensory_action_space"] = (meta["lo"], meta["hi"])
    if meta["peripheral_res"]:
        kw["peripheral_res"] = tuple(meta["peripheral_res"])
    kw.update(extra)
    return AtariEnvArgs(game="boxing", seed=0, obs_size=tuple(meta["obs_size"]), **kw)


def _script(meta):
    s = meta["script"]
    script = rh.ScreenScript(gr.screens("atari")[..., None], game_over_at=s["game_over_at"],
                             lives_at={int(k): v for k, v in s["lives_at"].items()})
    rh.ScreenScript.current = script
    return script


@pytest.mark.parametrize("name", ATARI)
def test_ale_pool_reproduces_reference_simulator_logic(name):
    """ALEPool (no-op/fire reset, t==2/t==3 frames, early game-over, episodic life) emits the same
    frames, flags and done signals as AtariEnv did in the reference run."""
    from active_gym_b200.sources import ALEPool
    z, meta = gr.load(name)
    script = _script(meta)
    pool = ALEPool(_atari_args(meta), 1, ale_factory=lambda i: rh._FakeALE())
    random.seed(meta["random_seed"])
    scr = gr.screens("atari")
    i = 0
    fa, fb, flags = pool.reset()
    while True:
        assert int(flags[0]) == int(z["flags"][i]), (name, i)
        if z["ia"][i] >= 0:
            assert np.array_equal(fa[0].numpy(), scr[z["ia"][i]]), (name, i)
        if z["ib"][i] >= 0:
            assert np.array_equal(fb[0].numpy(), scr[z["ib"][i]]), (name, i)
        if z["atype"][i] >= 0 and z["done"][i]:
            i += 1
            fa, fb, flags = pool.reset()
            continue
        i += 1
        if i >= len(z["flags"]):
            break
        fa, fb, flags, reward, done = pool.step([0])
        assert bool(done[0]) == bool(z["done"][i]), (name, i)
        assert reward[0] == min(meta["action_repeat"], 4 if not done[0] else reward[0])


def test_spaces_and_args_surface():
    import active_gym_b200 as ag
    for n in ("AtariBaseEnv", "AtariFixedFovealEnv", "AtariFlexibleFovealEnv", "AtariFixedFovealPeripheralEnv",
              "AtariEnvArgs", "DMCBaseEnv", "DMCFixedFovealEnv", "DMCFlexibleFovealEnv", "DMCFixedFovealPeripheralEnv",
              "DMCEnvArgs", "RecordWrapper", "FixedFovealEnv", "FlexibleFovealEnv", "FlexibleFovealEnvActionType",
              "FixedFovealPeripheralEnv"):
        assert hasattr(ag, n), n
    a = ag.AtariEnvArgs(game="boxing", seed=1, obs_size=(84, 84), fov_size=(30, 30), custom=5)
    assert (a.frame_stack, a.action_repeat, a.mask_out, a.record, a.clip_reward, a.custom) == (4, 4, False, False, False, 5)
    d = ag.DMCEnvArgs(domain_name="reacher", task_name="easy", seed=1, obs_size=(84, 84))
    assert (d.frame_stack, d.action_repeat, d.grey, d.from_pixels, d.camera_id) == (3, 4, True, True, 0)
    assert int(ag.FlexibleFovealEnvActionType.FOV_LOC) == 0 and int(ag.FlexibleFovealEnvActionType.FOV_RES) == 1


def test_product_fails_loudly_without_gpu_or_library(monkeypatch):
    from active_gym_b200 import ObservationPath, _lib
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ObservationPath(1, 4, (84, 84), (210, 160, 1))
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libagym_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


# ------------------------------------------------------------------------------------ GPU
def _drive_single_env(name, env, z, meta, script):
    """Replays the golden action sequence through a single-env drop-in, yields (call, obs, info)."""
    random.seed(meta.get("random_seed", 7))
    i = 0
    obs, info = env.reset()
    while True:
        yield i, obs, info
        if z["atype"][i] >= 0 and z["done"][i]:
            i += 1
            obs, info = env.reset()
            continue
        i += 1
        if i >= len(z["flags"]):
            return
        act = {"motor_action": 0 if meta["kind"] == "atari" else np.zeros(2, np.float32),
               "sensory_action": np.asarray(z["action"][i])}
        if meta["flexible"]:
            act["sensory_action_type"] = np.array([int(z["atype"][i])])
            if z["atype"][i] == 1:
                act["sensory_action"] = act["sensory_action"].astype(np.int64)
        obs, reward, done, truncated, info = env.step(act)
        assert truncated is False and bool(done) == bool(z["done"][i])
        assert info["ep_len"] >= 1 and "raw_reward" in info and "reward" in info


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", ATARI)
def test_atari_single_env_dropin_matches_reference(name, backend, monkeypatch):
    import active_gym_b200 as ag
    _backend(monkeypatch, backend)
    from active_gym_b200.sources import ALEPool
    z, meta = gr.load(name)
    script = _script(meta)
    args = _atari_args(meta)
    src = ALEPool(args, 1, ale_factory=lambda i: rh._FakeALE())
    make = {"FixedFovealEnv": ag.AtariFixedFovealEnv, "FlexibleFovealEnv": ag.AtariFlexibleFovealEnv,
            "FixedFovealPeripheralEnv": ag.AtariFixedFovealPeripheralEnv}[meta["env"].replace("Atari", "")]
    env = make(args, source=src)
    want_all = z["obs_u8"] if meta["exact"] else z["obs_f32"]
    for i, obs, info in _drive_single_env(name, env, z, meta, script):
        assert obs.dtype == np.float64
        assert np.array_equal(info["fov_loc"], z["loc"][i]), (name, i)
        want = want_all[i]
        if meta.get("ragged"):
            rh_, rw_ = (int(v) for v in z["res"][i])
            want = want[:, :rh_, :rw_]
            assert np.array_equal(info["fov_res"], z["res"][i])
        assert obs.shape == want.shape, (name, i, obs.shape, want.shape)
        if meta["exact"]:
            ref = (want.astype(np.float32) / np.float32(255)).astype(np.float64)
            assert np.array_equal(obs, ref), (name, i)  # the reference's float64 values, bit for bit
        else:
            assert np.abs(obs * 255.0 - want).max() <= 0.5 + 1e-2, (name, i)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", DMC)
def test_dmc_single_env_dropin_matches_reference(name, backend, monkeypatch):
    import active_gym_b200 as ag
    _backend(monkeypatch, backend)
    from active_gym_b200.sources import DMCPool
    z, meta = gr.load(name)
    script = rh.ScreenScript(gr.screens("dmc"))
    rh.ScreenScript.current = script
    kw = dict(fov_size=tuple(meta["fov_size"]), fov_init_loc=tuple(meta["fov_init_loc"]), sensory_action_mode=meta["mode"],
              frame_stack=meta["frame_stack"], action_repeat=meta["action_repeat"], mask_out=meta["variant"] == "mask")
    if meta["mode"] == "relative":
        kw["sensory_action_space"] = (meta["lo"], meta["hi"])
    if meta["peripheral_res"]:
        kw["peripheral_res"] = tuple(meta["peripheral_res"])
    args = ag.DMCEnvArgs(domain_name="reacher", task_name="easy", seed=0, obs_size=(84, 84), **kw)
    src = DMCPool(args, 1, env_factory=lambda i: rh._FakeDMC())
    make = ag.DMCFixedFovealPeripheralEnv if meta["peripheral_res"] else ag.DMCFixedFovealEnv
    env = make(args, source=src)
    want_all = z["obs_u8"] if meta["exact"] else z["obs_f32"]
    for i, obs, info in _drive_single_env(name, env, z, meta, script):
        assert np.array_equal(info["fov_loc"], z["loc"][i])
        if meta["exact"]:
            assert np.array_equal(obs, (want_all[i].astype(np.float32) / np.float32(255)).astype(np.float64))
        else:
            assert np.abs(obs * 255.0 - want_all[i]).max() <= 0.5 + 1e-2


@pytest.mark.gpu
def test_batched_env_equals_independent_envs_and_masked_reset():
    """N-env batch == N single envs, and reset(mask) touches only the masked envs."""
    import active_gym_b200 as ag
    from active_gym_b200.sources import SyntheticAtariSource
    n = 6
    args = ag.AtariEnvArgs(game="boxing", seed=0, obs_size=(84, 84), fov_size=(30, 30), fov_init_loc=(3, 5),
                           sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0), peripheral_res=(20, 20))
    src = SyntheticAtariSource(n, channels=1, device="cuda", pool=8, seed=5)
    env = ag.AtariFixedFovealPeripheralEnv(args, num_envs=n, source=src)
    obs, info = env.reset()
    assert obs.shape == (n, 4, 84, 84) and obs.dtype == torch.uint8 and obs.is_cuda
    assert torch.equal(info["fov_loc"].cpu(), torch.tensor([[3, 5]] * n, dtype=torch.int32))
    rng = np.random.default_rng(0)
    acts = [rng.integers(-10, 11, (n, 2)) for _ in range(3)]
    outs = []
    for a in acts:
        obs, reward, done, trunc, info = env.step({"motor_action": np.zeros(n, np.int64), "sensory_action": a})
        outs.append((obs.clone(), info["fov_loc"].clone()))
        assert info["ep_len"].tolist() == [len(outs)] * n
    # the same frames, one env at a time
    for e in range(n):
        class OneOf:
            raw_shape, n_actions = src.raw_shape, src.n_actions
            def __init__(s): s.t = 0
            def _next(s):
                b = src.batches[s.t % len(src.batches)][e:e + 1]; s.t += 1; return b
            def reset(s, mask=None):
                f = s._next(); return f, f, src.flags_reset[:1]
            def step(s, a):
                return s._next(), s._next(), src.flags_step[:1], np.zeros(1), np.zeros(1, bool)
        single = ag.AtariFixedFovealPeripheralEnv(args, num_envs=1, source=OneOf())
        single.reset()
        for k, a in enumerate(acts):
            o, *_rest, inf = single.step({"motor_action": np.zeros(1, np.int64), "sensory_action": a[e:e + 1]})
            assert torch.equal(o[0], outs[k][0][e]) and torch.equal(inf["fov_loc"][0], outs[k][1][e])
    # masked reset: env 1 and 4 restart, the others keep ring, loc and counters
    before_ring, before_loc = env.path.ring.clone(), env.path.loc.clone()
    mask = np.array([False, True, False, False, True, False])
    obs, info = env.reset(mask=mask)
    keep = torch.tensor(~mask)
    assert torch.equal(env.path.ring[keep], before_ring[keep]) and torch.equal(env.path.loc[keep], before_loc[keep])
    assert info["fov_loc"][1].tolist() == [3, 5] and info["ep_len"].tolist() == [3, 0, 3, 3, 0, 3]
    assert int(env.path.ring[1].ne(0).any(dim=-1).any(dim=-1).sum()) == 1, "hard reset leaves one non-zero frame"


def test_ale_pool_worker_threads_do_not_change_results():
    """N fake ALEs with their own scripts: 4 worker threads == serial, frame for frame (the no-op counts of
    the reference's reset come from the global `random` and are drawn serially in env order)."""
    from active_gym_b200 import AtariEnvArgs
    from active_gym_b200.sources import ALEPool
    scr = gr.screens("atari")
    n = 11

    def factory_for(offsets):
        def factory(i):
            s = rh.ScreenScript(scr[..., None], game_over_at=(40 + 7 * i, 90 + i), lives_at={25 + i: 2})
            s.acts = offsets[i]
            rh.ScreenScript.current = s
            return rh._FakeALE()
        return factory

    args = AtariEnvArgs(game="boxing", seed=0, obs_size=(84, 84), frame_stack=4, action_repeat=4)
    runs = []
    for workers in (1, 4):
        random.seed(123)
        pool = ALEPool(args, n, ale_factory=factory_for([3 * i for i in range(n)]), workers=workers)
        log = []
        fa, fb, fl = pool.reset()
        log.append((fa.numpy().copy(), fl.copy()))
        for step in range(12):
            fa, fb, fl, r, d = pool.step(np.zeros(n, np.int64))
            log.append((fa.numpy().copy(), fb.numpy().copy(), fl.copy(), r.copy(), d.copy()))
            if d.any():
                fa, fb, fl = pool.reset(mask=d)
                log.append((fa.numpy().copy(), fl.copy()))
        pool.close()
        runs.append(log)
    assert len(runs[0]) == len(runs[1])
    for a, b in zip(*runs):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)

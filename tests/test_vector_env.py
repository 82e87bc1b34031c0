"""gymnasium.vector.VectorEnv-shaped adapter and the multi-GPU front end.  Backend "oracle"
runs the Python logic in the CPU container (tests/fake_path.py), backend "cuda" the same through the kernels."""
import numpy as np
import pytest
import torch

from tests.test_env_api import BACKENDS, _backend


def _args(**kw):
    import active_gym_b200 as ag
    base = dict(game="boxing", seed=0, obs_size=(84, 84), fov_size=(30, 30), fov_init_loc=(4, 6),
                sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0), peripheral_res=(20, 20))
    base.update(kw)
    return ag.AtariEnvArgs(**base)


@pytest.mark.parametrize("backend", BACKENDS)
def test_vector_env_surface_spaces_and_autoreset(backend, monkeypatch):
    import active_gym_b200 as ag
    from active_gym_b200.sources import PinnedFrameSource
    _backend(monkeypatch, backend)
    n = 8
    env = ag.AtariFixedFovealPeripheralEnv(_args(), num_envs=n, source=PinnedFrameSource(n, pool=3, done_every=2))
    vec = ag.FovealVectorEnv(env)
    # the attributes gymnasium.vector.VectorEnv documents
    assert vec.num_envs == n and vec.is_vector_env and not vec.closed
    assert vec.single_observation_space.shape == (4, 84, 84) and vec.observation_space.shape == (n, 4, 84, 84)
    sa = vec.single_action_space["sensory_action"]
    assert sa.shape == (2,) and sa.low.tolist() == [-10, -10] and sa.high.tolist() == [10, 10]
    assert vec.action_space["sensory_action"].shape == (n, 2) and vec.action_space["motor_action"].shape == (n,)
    a = vec.action_space.sample()
    assert a["sensory_action"].shape == (n, 2) and a["motor_action"].shape == (n,)
    # the reference's own (degenerate, one-dimensional) declaration stays available (fov_env.py:125-129)
    ref_space = ag.FovealVectorEnv(env, fix_sensory_space=False).single_action_space["sensory_action"]
    assert ref_space.shape == (1,)
    obs, info = vec.reset(seed=3)
    assert tuple(obs.shape) == (n, 4, 84, 84) and [int(v) for v in info["fov_loc"][0]] == [4, 6]
    assert vec.get_attr("fov_size") == (30, 30) and vec.call("variant") == "crop"
    seen_final = False
    for step in range(6):
        obs, r, term, trunc, info = vec.step({"motor_action": np.zeros(n, np.int64), "sensory_action": np.full((n, 2), 3.0)})
        assert tuple(obs.shape) == (n, 4, 84, 84) and term.dtype == bool and trunc.dtype == bool and not trunc.any()
        ep_len = np.asarray(torch.as_tensor(info["ep_len"]).cpu())
        loc = np.asarray(torch.as_tensor(info["fov_loc"]).cpu())
        if term.any():
            seen_final = True
            k = int(term.sum())
            assert tuple(info["final_observation"].shape) == (k, 4, 84, 84) and info["_final_observation"].tolist() == term.tolist()
            assert np.asarray(torch.as_tensor(info["final_info"]["ep_len"]).cpu()).min() >= 2
            assert (ep_len[term] == 0).all() and (loc[term] == [4, 6]).all()       # restarted at once, like SyncVectorEnv
            assert (ep_len[~term] > 0).all()
        assert list(vec.envs[1].fov_loc) == loc[1].tolist()
    assert seen_final
    vec.close()
    assert vec.closed


@pytest.mark.parametrize("backend", BACKENDS)
def test_sharded_vec_env_equals_one_batch(backend, monkeypatch):
    """Global batch over several 'devices' (the same GPU twice on a 1-GPU box; env-index blocks from sharding.env_shard)
    == one unsharded env on the same frames."""
    import active_gym_b200 as ag
    from active_gym_b200.sources import PinnedFrameSource
    _backend(monkeypatch, backend)
    n = 11
    dev = "cuda:0" if backend == "cuda" else "cpu"
    full = PinnedFrameSource(n, pool=3, seed=21)

    class Block:   # rows [lo, hi) of the full source's batches
        def __init__(s, lo, hi):
            s.lo, s.hi, s.t, s.raw_shape, s.n_actions = lo, hi, 0, full.raw_shape, full.n_actions
        def _next(s):
            b = full.batches[s.t % len(full.batches)][s.lo:s.hi]; s.t += 1; return b
        def reset(s, mask=None):
            f = s._next(); return f, f, np.full(s.hi - s.lo, 5, np.uint8)
        def step(s, a):
            return s._next(), s._next(), np.full(s.hi - s.lo, 3, np.uint8), np.zeros(s.hi - s.lo), np.zeros(s.hi - s.lo, bool)

    args = _args(shards=2)
    sh = ag.ShardedVecEnv(lambda m, d, lo, hi: ag.AtariFixedFovealPeripheralEnv(args, num_envs=m, source=Block(lo, hi), device=d),
                          n, devices=[dev, dev, dev])
    assert sh.ranges == [(0, 4), (4, 8), (8, 11)]
    one = ag.AtariFixedFovealPeripheralEnv(args, num_envs=n, source=Block(0, n), device=dev)
    o_sh, i_sh = sh.reset()
    o_one, i_one = one.reset()
    rng = np.random.default_rng(1)
    for step in range(3):
        act = {"motor_action": np.zeros(n, np.int64), "sensory_action": rng.integers(-10, 11, (n, 2)).astype(np.float64)}
        o_sh, r, d, t, i_sh = sh.step(act)
        o_one, r1, d1, t1, i_one = one.step(act)
        assert torch.equal(ag.ShardedVecEnv.gather(o_sh, dev), o_one), step
        assert torch.equal(torch.cat([torch.as_tensor(v) for v in i_sh["fov_loc"]]) if isinstance(i_sh["fov_loc"], list)
                           else torch.as_tensor(i_sh["fov_loc"]), torch.as_tensor(i_one["fov_loc"])), step
        assert r.shape == (n,) and d.shape == (n,)
    sh.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_env_returns_normalised_observations_when_obs_dtype_is_set(backend, monkeypatch):
    """args.obs_dtype: the env hands out float32(u8) / 255 (the reference's values, atari_env.py:75) in the requested
    float type, the u8 tensor stays available in last_obs_u8."""
    import active_gym_b200 as ag
    from active_gym_b200.sources import PinnedFrameSource
    _backend(monkeypatch, backend)
    n = 6
    for make, kw in ((ag.AtariFixedFovealPeripheralEnv, {}), (ag.AtariFixedFovealEnv, {}), (ag.AtariFlexibleFovealEnv, dict(mask_out=True))):
        env = make(_args(obs_dtype=torch.float16, **kw), num_envs=n, source=PinnedFrameSource(n, pool=3))
        obs, _ = env.reset()
        act = {"motor_action": np.zeros(n, np.int64), "sensory_action": np.full((n, 2), 3.0), "sensory_action_type": np.zeros(n, np.int64)}
        obs, *_ = env.step(act)
        assert obs.dtype == torch.float16 and obs.shape == env.last_obs_u8.shape
        want = (env.last_obs_u8.to(torch.float32) / 255.0).to(torch.float16)
        assert torch.equal(obs.cpu(), want.cpu())


@pytest.mark.parametrize("backend", BACKENDS)
def test_async_sim_thread_gives_the_same_steps(backend, monkeypatch):
    """args.async_sim: step_async returns before the simulators have stepped (they run, with the enqueueing of copies and
    kernels, on the env's step thread); step_wait joins.  Same observations, counters and fovea as the synchronous env."""
    import time
    import active_gym_b200 as ag
    from active_gym_b200.sources import PinnedFrameSource
    _backend(monkeypatch, backend)
    n = 7

    class SlowSource(PinnedFrameSource):
        def step(self, a):
            time.sleep(0.05)
            return super().step(a)

    envs = [ag.AtariFixedFovealPeripheralEnv(_args(async_sim=flag, shards=2), num_envs=n, source=SlowSource(n, pool=3, done_every=4))
            for flag in (False, True)]
    for e in envs:
        e.reset()
    rng = np.random.default_rng(5)
    for step in range(5):
        act = {"motor_action": np.zeros(n, np.int64), "sensory_action": rng.integers(-10, 11, (n, 2)).astype(np.float64)}
        t0 = time.perf_counter()
        envs[1].step_async(act)
        assert time.perf_counter() - t0 < 0.04, "step_async must not wait for the simulators"
        want = envs[0].step(act)
        got = envs[1].step_wait()
        assert torch.equal(torch.as_tensor(got[0]).cpu(), torch.as_tensor(want[0]).cpu()), step
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
        for k in ("fov_loc", "ep_len", "reward"):
            assert torch.equal(torch.as_tensor(got[4][k]).cpu(), torch.as_tensor(want[4][k]).cpu()), (step, k)
    for e in envs:
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["peripheral", "fixed", "flexible"])
def test_sharded_vec_env_learner_device_output_is_written_by_the_kernels(kind):
    """ShardedVecEnv(learner_device=): every device's observe kernel stores its env block straight into ONE tensor on
    the learner's GPU (peer memory when the devices differ; all visible GPUs are used, the same GPU three times on a
    1-GPU box) — equal to the oracle's observation of every env, no gather copy."""
    import active_gym_b200 as ag
    from active_gym_b200.sources import PinnedFrameSource
    from oracle import agym_oracle as orc
    n, K, S, fov = 301, 4, (84, 84), (30, 30)
    ndev = torch.cuda.device_count()
    devices = [f"cuda:{i}" for i in range(ndev)] if ndev > 1 else ["cuda:0"] * 3
    full = PinnedFrameSource(n, pool=3, seed=33)

    class Block:   # rows [lo, hi) of the full source's batches
        def __init__(s, lo, hi):
            s.lo, s.hi, s.t, s.raw_shape, s.n_actions = lo, hi, 0, full.raw_shape, full.n_actions
        def _next(s):
            b = full.batches[s.t % len(full.batches)][s.lo:s.hi]; s.t += 1; return b
        def reset(s, mask=None):
            f = s._next(); return f, f, np.full(s.hi - s.lo, 5, np.uint8)
        def step(s, a):
            return s._next(), s._next(), np.full(s.hi - s.lo, 3, np.uint8), np.zeros(s.hi - s.lo), np.zeros(s.hi - s.lo, bool)

    make = {"peripheral": ag.AtariFixedFovealPeripheralEnv, "fixed": ag.AtariFixedFovealEnv, "flexible": ag.AtariFlexibleFovealEnv}[kind]
    args = _args(shards=2, mask_out=(kind == "flexible"))
    sh = ag.ShardedVecEnv(lambda m, d, lo, hi: make(args, num_envs=m, source=Block(lo, hi), device=d), n, devices=devices,
                          learner_device="cuda:0")
    ring, head = orc.new_state(n, K, S)
    loc = np.tile(np.array([4, 6], np.int32), (n, 1))
    res = np.tile(np.array(fov, np.int32), (n, 1))

    def want():
        if kind == "peripheral":
            return orc.observe_peripheral(ring, head, loc, fov, (20, 20))
        if kind == "flexible":
            return orc.observe_flexible(ring, head, loc, res, fov, variant="mask")
        return orc.observe_fixed(ring, head, loc, fov, variant="crop").astype(np.float64)

    tol = 0.0 if kind == "fixed" else 0.5 + 1e-2
    obs, info = sh.reset()
    assert obs is sh.obs and obs.device == torch.device("cuda:0") and tuple(obs.shape)[0] == n
    b = full.batches[0].numpy()
    orc.ingest_atari(b, b, np.full(n, 5, np.uint8), ring, head, orc.LUMA_RGB)
    torch.cuda.current_stream(obs.device).synchronize()
    assert np.abs(obs.cpu().numpy().astype(np.float64) - want()).max() <= tol
    rng = np.random.default_rng(5)
    t = 1
    for step in range(3):
        sa = rng.integers(-10, 11, (n, 2)).astype(np.float64)
        act = {"motor_action": np.zeros(n, np.int64), "sensory_action": sa, "sensory_action_type": np.zeros(n, np.int64)}
        obs, r, d, tr, info = sh.step(act)
        nb = len(full.batches)
        fa, fb = full.batches[t % nb].numpy(), full.batches[(t + 1) % nb].numpy()
        t += 2
        orc.ingest_atari(fa, fb, np.full(n, 3, np.uint8), ring, head, orc.LUMA_RGB)
        orc.update_loc(sa, loc, obs_size=S, fov_size=fov, relative=True, lo=-10.0, hi=10.0,
                       **(dict(atype=np.zeros(n, np.int32), res=res) if kind == "flexible" else {}))
        torch.cuda.current_stream(obs.device).synchronize()   # the learner's stream already waits for every device
        assert np.abs(obs.cpu().numpy().astype(np.float64) - want()).max() <= tol, (kind, step)
    sh.close()

"""GPU: the persistent / pipelined kernels at batch sizes where every CTA walks several envs
(TMA ingest, peripheral std kernel with its mbarrier ring and batched fov_loc, flexible fast path),
with ragged per-env flags and fovea controls, against the CPU oracle for EVERY env; and the host
pipeline's env shards against the unsharded path."""
import numpy as np
import pytest
import torch

from oracle import agym_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 0.5 + 1e-2
S = (84, 84)


def _path(n, K=4, raw=(210, 160, 1), luma=None, **kw):
    from active_gym_b200 import ObservationPath, LUMA_RGB
    return ObservationPath(n, K, S, raw, luma=luma or LUMA_RGB, **kw)


def _np(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def _flags(rng, n):
    choices = np.array([3, 3, 3, 3, 1, 0, 5, 1 | 4, 8, 3 | 4, 2], np.uint8)  # incl. frame B only, idle, resets
    return choices[rng.integers(0, len(choices), n)]


def _frames(rng, n):
    # cheap but structured: per-env random 8x8 blocks upsampled + noise, so that resize errors would show
    base = rng.integers(0, 256, (n, 27, 20), dtype=np.uint8).repeat(8, 1).repeat(8, 2)[:, :210, :160]
    return base ^ rng.integers(0, 32, (n, 210, 160), dtype=np.uint8)


@pytest.mark.parametrize("n", [445, 1500])
def test_tma_ingest_many_envs_per_cta_ragged(n):
    rng = np.random.default_rng(n)
    K = 4
    p = _path(n, K, fov_size=(30, 30), peripheral_res=(20, 20))  # with the squeeze cache
    ring, head = orc.new_state(n, K, S)
    for step in range(4):
        fa, fb = _frames(rng, n), _frames(rng, n)
        fl = np.full(n, 5, np.uint8) if step == 0 else _flags(rng, n)
        p.ingest_atari(fa, fb, fl)
        orc.ingest_atari(fa, fb, fl, ring, head)
        assert np.array_equal(_np(p.head), head), step
        assert np.array_equal(_np(p.ring), ring), step
    # the cache written by the ingest kernel == the squeeze recomputed from the ring (uncached kernel)
    a = rng.integers(0, 55, (n, 2)).astype(np.float64)
    cached = p.observe_peripheral(a)
    direct = p.observe_peripheral(None, ctrl=np.full(n, 2, np.uint8), use_cache=False)
    assert (cached.int() - direct.int()).abs().max().item() <= 1


@pytest.mark.parametrize("n", [449, 1400])
def test_tma_ingest_rgb_many_envs_per_cta_ragged(n):
    """3-channel frames through the persistent TMA ingest (luma fused into the horizontal pass): bit exact."""
    rng = np.random.default_rng(3 * n)
    K = 4
    p = _path(n, K, raw=(210, 160, 3), fov_size=(30, 30))
    ring, head = orc.new_state(n, K, S)
    for step in range(3):
        fa = rng.integers(0, 256, (n, 210, 160, 3), dtype=np.uint8)
        fb = rng.integers(0, 256, (n, 210, 160, 3), dtype=np.uint8)
        fl = np.full(n, 5, np.uint8) if step == 0 else _flags(rng, n)
        p.ingest_atari(fa, fb, fl)
        orc.ingest_atari(fa, fb, fl, ring, head)
        assert np.array_equal(_np(p.head), head), step
        assert np.array_equal(_np(p.ring), ring), step


@pytest.mark.parametrize("K,n", [(4, 700), (3, 611), (4, 297)])
def test_peripheral_std_kernel_all_envs_vs_oracle(K, n):
    rng = np.random.default_rng(100 + n)
    fov, periph = (30, 30), (20, 20)
    p = _path(n, K, fov_size=fov, peripheral_res=periph, fov_init_loc=(7, 11), sensory_action_mode="relative",
              sensory_action_space=(-10.0, 10.0))
    ring, head = orc.new_state(n, K, S)
    loc = np.zeros((n, 2), np.int32)
    for step in range(5):
        fa, fb = _frames(rng, n), _frames(rng, n)
        fl = np.full(n, 5, np.uint8) if step == 0 else _flags(rng, n)
        p.ingest_atari(fa, fb, fl)
        orc.ingest_atari(fa, fb, fl, ring, head)
        a = rng.uniform(-14, 14, (n, 2))
        a[::5] = rng.integers(-10, 11, (len(a[::5]), 2)) + 0.5  # ties: round half to even after the clip
        # fovea control: reset everywhere at step 0, afterwards a mix of apply / reset / keep
        ctrl = np.full(n, 1, np.uint8) if step == 0 else rng.choice(np.array([0, 0, 0, 1, 2], np.uint8), n)
        apply_, reset = ctrl == 0, ctrl == 1
        new = loc.copy()
        orc.update_loc(a, new, obs_size=S, fov_size=fov, relative=True, lo=-10.0, hi=10.0)
        loc[apply_] = new[apply_]
        loc[reset] = (7, 11)
        got = _np(p.observe_peripheral(a, ctrl=ctrl))
        assert np.array_equal(_np(p.loc), loc), step
        want = orc.observe_peripheral(ring, head, loc, fov, periph)
        assert np.abs(got.astype(np.float64) - want).max() <= TOL, step
        full = orc.stack(ring, head)
        for e in range(0, n, 13):  # the pasted fovea is bit exact
            r, c = loc[e]
            assert np.array_equal(got[e, :, r:r + 30, c:c + 30], full[e, :, r:r + 30, c:c + 30])


def test_peripheral_std_kernel_other_fovea_sizes():
    # the std kernel takes any fovea at the 84/20 geometry: non-multiple-of-4 widths, tall, tiny, huge
    rng = np.random.default_rng(5)
    n, K = 333, 4
    for fov in [(30, 30), (17, 45), (61, 9), (83, 83), (1, 1)]:
        p = _path(n, K, fov_size=fov, peripheral_res=(20, 20), sensory_action_mode="absolute")
        ring, head = orc.new_state(n, K, S)
        for step in range(K + 1):
            fa, fb = _frames(rng, n), _frames(rng, n)
            fl = np.full(n, 5 if step == 0 else 3, np.uint8)
            p.ingest_atari(fa, fb, fl)
            orc.ingest_atari(fa, fb, fl, ring, head)
        loc = np.zeros((n, 2), np.int32)
        a = rng.integers(-5, 90, (n, 2)).astype(np.float64)
        orc.update_loc(a, loc, obs_size=S, fov_size=fov)
        got = _np(p.observe_peripheral(a)).astype(np.float64)
        assert np.array_equal(_np(p.loc), loc), fov
        assert np.abs(got - orc.observe_peripheral(ring, head, loc, fov, (20, 20))).max() <= TOL, fov


@pytest.mark.parametrize("variant,pad,n", [("mask", None, 300), ("crop", None, 300), ("crop", (52, 56), 300),
                                           ("mask", None, 1500), ("crop", (52, 56), 1500)])
def test_flexible_fast_path_many_envs(variant, pad, n):
    # n = 1500: every persistent CTA (2 per SM) claims ~5 envs from the device counter, so the cp.async window
    # prefetch, the two-ahead fov update and the re-armed counter (several launches) are all exercised
    rng = np.random.default_rng(21)
    K, fov = 4, (30, 30)
    p = _path(n, K, fov_size=fov, sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0))
    ring, head = orc.new_state(n, K, S)
    for step in range(K + 1):
        fa, fb = _frames(rng, n), _frames(rng, n)
        fl = np.full(n, 5 if step == 0 else 3, np.uint8)
        p.ingest_atari(fa, fb, fl)
        orc.ingest_atari(fa, fb, fl, ring, head)
    loc = np.zeros((n, 2), np.int32)
    res = np.tile(np.array([fov], np.int32), (n, 1))
    p.observe_flexible(None, variant=variant, ctrl="reset", pad=pad)
    for step in range(4):
        atype = rng.integers(0, 2, n).astype(np.int32)
        hi = 51 if pad else 85
        a = np.where(atype[:, None] == 1, rng.integers(1, hi, (n, 2)), rng.uniform(-12, 12, (n, 2))).astype(np.float64)
        orc.update_loc(a, loc, obs_size=S, fov_size=fov, relative=True, lo=-10.0, hi=10.0, atype=atype, res=res)
        got = _np(p.observe_flexible(a, atype, variant=variant, pad=pad)).astype(np.float64)
        want = orc.observe_flexible(ring, head, loc, res, fov, variant=variant, pad=pad)
        assert np.array_equal(_np(p.loc), loc) and np.array_equal(_np(p.res), res), step
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= TOL, step
        sharp = res[:, 0] <= fov[0]  # not blurred: bit exact (fov_env.py:286 looks at rows only)
        assert np.array_equal(got[sharp], want[sharp])


@pytest.mark.parametrize("K,fov", [(3, (30, 30)), (2, (24, 40)), (5, (30, 30))])
def test_flexible_v3_other_stack_depths_and_fovea_shapes(K, fov):
    """DMC-style K = 3, K = 2 / 5 and a non-square fovea (blur decided by rows only, fov_env.py:286) through the
    persistent flexible kernel: the shared-memory budget, the frame groups of the W/H passes and the ring order
    depend on K."""
    rng = np.random.default_rng(100 + K)
    n = 333
    p = _path(n, K, fov_size=fov, sensory_action_mode="absolute")
    ring, head = orc.new_state(n, K, S)
    for step in range(K + 2):
        fa, fb = _frames(rng, n), _frames(rng, n)
        fl = np.full(n, 5 if step == 0 else 3, np.uint8)
        p.ingest_atari(fa, fb, fl)
        orc.ingest_atari(fa, fb, fl, ring, head)
    loc = np.zeros((n, 2), np.int32)
    res = np.tile(np.array([fov], np.int32), (n, 1))
    p.observe_flexible(None, variant="mask", ctrl="reset")
    for step in range(3):
        atype = rng.integers(0, 2, n).astype(np.int32)
        a = np.where(atype[:, None] == 1, rng.integers(1, 85, (n, 2)), rng.uniform(-5, 90, (n, 2))).astype(np.float64)
        orc.update_loc(a, loc, obs_size=S, fov_size=fov, atype=atype, res=res)
        got = _np(p.observe_flexible(a, atype, variant="mask")).astype(np.float64)
        want = orc.observe_flexible(ring, head, loc, res, fov, variant="mask")
        assert np.array_equal(_np(p.loc), loc) and np.array_equal(_np(p.res), res), step
        assert np.abs(got - want).max() <= TOL, step
        sharp = res[:, 0] <= fov[0]
        assert np.array_equal(got[sharp], want[sharp])


@pytest.mark.parametrize("shards,pinned", [(7, True), (3, False), (1, True)])
def test_pipelined_path_host_inputs_match_the_oracle(shards, pinned):
    """PipelinedPath (env-index shards on their own streams, host frames / actions in, pinned observations out) against
    the ORACLE: full host frames (strided packed copy when pinned), then frames that already hold only the sampled rows."""
    from active_gym_b200.pipeline import PipelinedPath
    rng = np.random.default_rng(9)
    n, K = 203, 4  # not a multiple of the shard count
    kw = dict(fov_size=(30, 30), sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0), peripheral_res=(20, 20))
    pp = PipelinedPath(n, K, S, (210, 160, 1), shards=shards, **kw)
    assert pp.run == (5, 3, 4)
    ring, head = orc.new_state(n, K, S)
    loc = np.zeros((n, 2), np.int32)
    rows = pp.used_rows

    def host(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.pin_memory() if pinned else t

    for step in range(5):
        fa, fb = _frames(rng, n), _frames(rng, n)
        fl = np.full(n, 5, np.uint8) if step == 0 else _flags(rng, n)
        if step % 2:
            pp.ingest_atari(host(fa[:, rows]), host(fb[:, rows]), fl, packed=True)
        else:
            pp.ingest_atari(host(fa), host(fb), fl)
        orc.ingest_atari(fa, fb, fl, ring, head)
        if step == 0:
            out, (h_obs, h_loc, _) = pp.observe_peripheral(None, ctrl="reset", host_out=True)
        else:
            a = rng.integers(-10, 11, (n, 2)).astype(np.float64)
            orc.update_loc(a, loc, obs_size=S, fov_size=(30, 30), relative=True, lo=-10.0, hi=10.0)
            out, (h_obs, h_loc, _) = pp.observe_peripheral(a, host_out=True)
        pp.sync()
        assert np.array_equal(_np(pp.ring), ring) and np.array_equal(_np(pp.head), head), step
        assert np.array_equal(h_loc.numpy(), loc), step
        want = orc.observe_peripheral(ring, head, loc, (30, 30), (20, 20))
        assert np.abs(h_obs.numpy().astype(np.float64) - want).max() <= TOL, step
        assert np.array_equal(h_obs.numpy(), _np(out)), step
    assert pp.h2d_bytes > 0 and pp.d2h_bytes > 0


def test_pipelined_path_device_inputs_equal_single_path_and_record_counters():
    """Device-resident inputs: the shards read slices of the caller's tensors (no copies); results equal one unsharded
    ObservationPath bit for bit; the RecordWrapper counters / trace row follow k_record_step's contract."""
    from active_gym_b200.pipeline import PipelinedPath
    rng = np.random.default_rng(10)
    n, K = 130, 4
    kw = dict(fov_size=(30, 30), sensory_action_mode="absolute")
    pp = PipelinedPath(n, K, S, (210, 160, 1), shards=4, **kw)
    ref = _path(n, K, **kw)
    trace = torch.zeros((n, 6), dtype=torch.int32, device="cuda")
    want_len, want_cum = np.zeros(n, np.int64), np.zeros(n)
    for step in range(4):
        fa = torch.from_numpy(_frames(rng, n)).cuda()
        fb = torch.from_numpy(_frames(rng, n)).cuda()
        fl = torch.from_numpy(np.full(n, 5, np.uint8) if step == 0 else _flags(rng, n)).cuda()
        a = torch.from_numpy(rng.uniform(-5, 60, (n, 2))).cuda()
        at = torch.from_numpy(rng.integers(0, 2, n).astype(np.int32)).cuda()
        a = torch.where(at[:, None] == 1, torch.randint(20, 51, (n, 2), device="cuda").double(), a)
        pp.ingest_atari(fa, fb, fl)
        ref.ingest_atari(fa, fb, fl)
        got = pp.observe_flexible(a, at, variant="mask")
        want = ref.observe_flexible(a, at, variant="mask")
        assert torch.equal(got, want) and torch.equal(pp.loc, ref.loc) and torch.equal(pp.res, ref.res), step
        assert torch.equal(pp.ring, ref.ring) and torch.equal(pp.head, ref.head), step
        rew = rng.uniform(-1, 1, n)
        done = rng.random(n) < 0.3
        pp.record_step(raw_reward=rew, done=done, trace_row=trace, with_res=True)
        want_len += 1
        want_cum += rew
        assert np.array_equal(_np(pp.ep_len), want_len) and np.array_equal(_np(pp.cum_reward), want_cum)
        t = _np(trace)
        assert np.array_equal(t[:, 0:2], _np(pp.loc)) and np.array_equal(t[:, 2:4], _np(pp.res))
        assert np.array_equal(t[:, 4], want_len) and np.array_equal(t[:, 5], (~done).astype(np.int32))
        mask = rng.random(n) < 0.2
        pp.record_step(reset_mask=mask, is_reset=True)
        want_len[mask], want_cum[mask] = 0, 0
        assert np.array_equal(_np(pp.ep_len), want_len) and np.array_equal(_np(pp.cum_reward), want_cum)
    assert pp.read_errors() == 0


def test_flexible_device_actions_report_invalid_windows_in_the_error_word():
    """FOV_RES actions the reference would fail on (fov_env.py:322-324: fractional or larger than the frame) are
    clamped by the kernels and reported: ObservationPath.read_errors, and the env raises like the reference."""
    import active_gym_b200 as ag
    from active_gym_b200 import _lib
    from active_gym_b200.sources import SyntheticAtariSource
    n = 40
    p = _path(n, 4, fov_size=(30, 30), sensory_action_mode="absolute")
    p.ingest_atari(np.zeros((n, 210, 160), np.uint8), np.zeros((n, 210, 160), np.uint8), np.full(n, 5, np.uint8))
    at = torch.ones(n, dtype=torch.int32, device="cuda")
    ok = torch.full((n, 2), 40.0, dtype=torch.float64, device="cuda")
    p.observe_flexible(ok, at, variant="mask")
    assert p.read_errors() == 0
    bad = ok.clone(); bad[3, 0] = 90.0
    p.observe_flexible(bad, at, variant="mask")
    assert p.read_errors() == _lib.ERR_RES_RANGE and p.read_errors() == 0
    assert _np(p.res)[3].tolist() == [84, 40]
    bad = ok.clone(); bad[7, 1] = 33.5; bad[9, 0] = float("nan")
    p.observe_flexible(bad, at, variant="crop", pad=(84, 84))
    assert p.read_errors() == (_lib.ERR_RES_RANGE | _lib.ERR_RES_FRACTION)
    assert _np(p.res)[7].tolist() == [40, 33] and _np(p.res)[9].tolist() == [1, 40]
    # through the env: device-tensor actions are validated after the fact (at the latest one step later)
    args = ag.AtariEnvArgs(game="boxing", seed=0, obs_size=(84, 84), fov_size=(30, 30), fov_init_loc=(0, 0),
                           sensory_action_mode="absolute", mask_out=True)
    env = ag.AtariFlexibleFovealEnv(args, num_envs=n, source=SyntheticAtariSource(n, device="cuda"))
    env.reset()
    act = {"motor_action": np.zeros(n, np.int64), "sensory_action": bad, "sensory_action_type": at}
    with pytest.raises(ValueError, match="FOV_RES"):
        for _ in range(3):
            env.step(act)
            torch.cuda.synchronize()
    with pytest.raises(ValueError, match="FOV_RES"):   # host actions are checked before the launch
        env.step({"motor_action": np.zeros(n, np.int64), "sensory_action": np.full((n, 2), 99.0), "sensory_action_type": np.ones(n, np.int64)})


@pytest.mark.parametrize("ch,n", [(1, 700), (3, 40)])
def test_packed_rows_ingest_is_bit_identical(ch, n):
    # frames that carry only the raw rows the resize samples (agym_ingest_atari_packed)
    rng = np.random.default_rng(31)
    K = 4
    a = _path(n, K, raw=(210, 160, ch), fov_size=(30, 30), peripheral_res=(20, 20))
    b = _path(n, K, raw=(210, 160, ch), fov_size=(30, 30), peripheral_res=(20, 20))
    rows = a.used_rows
    assert len(rows) == 168 and rows[0] == 0 and rows[-1] == 209 and 2 not in rows
    for step in range(3):
        shape = (n, 210, 160) if ch == 1 else (n, 210, 160, 3)
        fa = rng.integers(0, 256, shape, dtype=np.uint8)
        fb = rng.integers(0, 256, shape, dtype=np.uint8)
        fl = np.full(n, 5, np.uint8) if step == 0 else _flags(rng, n)
        a.ingest_atari(fa, fb, fl)
        b.ingest_atari_packed(np.ascontiguousarray(fa[:, rows]), np.ascontiguousarray(fb[:, rows]), fl)
        assert torch.equal(a.ring, b.ring) and torch.equal(a.head, b.head), step
        assert torch.equal(a.pcache, b.pcache), step


def test_periodic_run_detection():
    from active_gym_b200.pipeline import periodic_run
    p = _path(4, 4, fov_size=(30, 30))
    assert periodic_run(p.used_rows, 210) == (5, 3, 4)
    assert periodic_run(np.arange(84), 84) is None  # nothing to skip: no packed mode


@pytest.mark.parametrize("obs,fov,periph", [((64, 64), (20, 24), (16, 16)), ((96, 96), (30, 30), (24, 20)), ((84, 84), (30, 30), (21, 28))])
def test_other_geometries_use_the_table_driven_kernels(obs, fov, periph):
    # nothing here is the standard 84/20 geometry: generic / v2 kernels, other TMA unit counts
    from active_gym_b200 import ObservationPath
    rng = np.random.default_rng(obs[0] + periph[0])
    n, K = 450, 4
    p = ObservationPath(n, K, obs, (210, 160, 1), fov_size=fov, peripheral_res=periph, sensory_action_mode="relative",
                        sensory_action_space=(-6.0, 6.0))
    ring, head = orc.new_state(n, K, obs)
    loc = np.zeros((n, 2), np.int32)
    p.observe_peripheral(None, ctrl="reset")
    for step in range(K + 1):
        fa, fb = _frames(rng, n), _frames(rng, n)
        fl = np.full(n, 5, np.uint8) if step == 0 else _flags(rng, n)
        p.ingest_atari(fa, fb, fl)
        orc.ingest_atari(fa, fb, fl, ring, head)  # the oracle takes the obs size from the ring's shape
        assert np.array_equal(_np(p.ring), ring), step
        a = rng.uniform(-8, 8, (n, 2))
        orc.update_loc(a, loc, obs_size=obs, fov_size=fov, relative=True, lo=-6.0, hi=6.0)
        got = _np(p.observe_peripheral(a)).astype(np.float64)
        assert np.array_equal(_np(p.loc), loc), step
        assert np.abs(got - orc.observe_peripheral(ring, head, loc, fov, periph)).max() <= TOL, step
    crop = _np(p.observe_fixed(None, ctrl=np.full(n, 2, np.uint8)))
    assert np.array_equal(crop, orc.observe_fixed(ring, head, loc, fov))
    mask = _np(p.observe_fixed(None, variant="mask", ctrl=np.full(n, 2, np.uint8)))
    assert np.array_equal(mask, orc.observe_fixed(ring, head, loc, fov, variant="mask"))
    res = np.tile(np.array([fov], np.int32), (n, 1))
    p.res[:] = torch.from_numpy(res).cuda()
    atype = np.ones(n, np.int32)
    a = rng.integers(1, obs[0] + 1, (n, 2)).astype(np.float64)
    orc.update_loc(a, loc, obs_size=obs, fov_size=fov, relative=True, lo=-6.0, hi=6.0, atype=atype, res=res)
    got = _np(p.observe_flexible(a, atype, variant="mask")).astype(np.float64)
    assert np.array_equal(_np(p.res), res) and np.array_equal(_np(p.loc), loc)
    assert np.abs(got - orc.observe_flexible(ring, head, loc, res, fov, variant="mask")).max() <= TOL

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


def test_host_pipeline_shards_equal_unsharded_path():
    from active_gym_b200.hostpipe import HostPipelinedEnv
    rng = np.random.default_rng(9)
    n, K = 203, 4  # not a multiple of the shard count
    kw = dict(fov_size=(30, 30), sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0), peripheral_res=(20, 20))
    env = HostPipelinedEnv(n, K, S, (210, 160, 1), kind="atari", wrapper="peripheral", shards=7, **kw)
    ref = _path(n, K, **kw)
    hf = env.alloc_host_frames()
    for t in hf:
        t.numpy()[...] = _frames(rng, n)
    obs, loc = env.reset_host(hf)
    ref.ingest_atari(hf[0].numpy(), hf[0].numpy(), np.full(n, 5, np.uint8))
    want = ref.observe_peripheral(None, ctrl="reset")
    assert np.array_equal(obs.numpy(), _np(want)) and np.array_equal(loc.numpy(), _np(ref.loc))
    for step in range(3):
        for t in hf:
            t.numpy()[...] = _frames(rng, n)
        a = rng.integers(-10, 11, (n, 2)).astype(np.float64)
        obs, loc = env.step_host(hf, a)
        ref.ingest_atari(hf[0].numpy(), hf[1].numpy(), np.full(n, 3, np.uint8))
        want = ref.observe_peripheral(a)
        assert np.array_equal(obs.numpy(), _np(want)), step
        assert np.array_equal(loc.numpy(), _np(ref.loc)), step


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


def test_host_pipeline_packed_h2d_equals_full_frames():
    from active_gym_b200.hostpipe import HostPipelinedEnv, periodic_run
    rng = np.random.default_rng(13)
    n, K = 150, 4
    kw = dict(fov_size=(30, 30), sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0), peripheral_res=(20, 20))
    packed = HostPipelinedEnv(n, K, S, (210, 160, 1), kind="atari", wrapper="peripheral", shards=4, packed_h2d=True, **kw)
    full = HostPipelinedEnv(n, K, S, (210, 160, 1), kind="atari", wrapper="peripheral", shards=3, packed_h2d=False, **kw)
    assert packed.run == (5, 3, 4) and full.run is None
    assert packed.h2d_bytes_per_step < 0.81 * full.h2d_bytes_per_step
    assert periodic_run(np.arange(84), 84) is None  # nothing to skip: no packed mode
    hp, hf = packed.alloc_host_frames(), full.alloc_host_frames()
    for step in range(4):
        for tp, tf in zip(hp, hf):
            tf.numpy()[...] = _frames(rng, n)
            tp.numpy()[...] = tf.numpy()
        a = rng.integers(-10, 11, (n, 2)).astype(np.float64)
        (op, lp), (of, lf) = (packed.reset_host(hp), full.reset_host(hf)) if step == 0 else (packed.step_host(hp, a), full.step_host(hf, a))
        assert np.array_equal(op.numpy(), of.numpy()) and np.array_equal(lp.numpy(), lf.numpy()), step


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

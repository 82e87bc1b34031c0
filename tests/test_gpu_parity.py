"""GPU: seeded batches, CUDA path (through the C ABI) vs the CPU oracle on the same inputs, at
sizes the oracle finishes in seconds; plus size-independent properties at BASELINE.json sizes."""
import numpy as np
import pytest
import torch

from oracle import agym_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 0.5 + 1e-2  # half an LSB of rounding + evaluation error (fp32, or the 2^-8 grid of the biased lerp)


def _path(n, K=4, raw=(210, 160, 1), luma=None, **kw):
    from active_gym_b200 import ObservationPath, LUMA_RGB
    return ObservationPath(n, K, (84, 84), raw, luma=luma or LUMA_RGB, **kw)


def _np(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def _ragged_flags(rng, n, atari=True):
    # every combination the simulators can produce: both frames, A only, none, hard/soft reset, idle
    choices = np.array([3, 3, 3, 1, 0, 5, 1 | 4, 8, 3 | 4] if atari else [1, 1, 1, 5, 8], np.uint8)
    return choices[rng.integers(0, len(choices), n)]


@pytest.mark.parametrize("ch", [1, 3])
def test_ingest_atari_and_stack_bit_exact(ch):
    rng = np.random.default_rng(10 + ch)
    n, K = 96, 4
    p = _path(n, K, raw=(210, 160, ch))
    ring, head = orc.new_state(n, K, (84, 84))
    shape = (n, 210, 160) if ch == 1 else (n, 210, 160, 3)
    for step in range(6):
        fa = rng.integers(0, 256, shape, dtype=np.uint8)
        fb = rng.integers(0, 256, shape, dtype=np.uint8)
        flags = np.full(n, 5, np.uint8) if step == 0 else _ragged_flags(rng, n)
        p.ingest_atari(fa, fb, flags)
        orc.ingest_atari(fa, fb, flags, ring, head)
        assert np.array_equal(_np(p.head), head), step
        assert np.array_equal(_np(p.ring), ring), step
        assert np.array_equal(_np(p.stack()), orc.stack(ring, head)), step


def test_ingest_dmc_bit_exact():
    from active_gym_b200 import LUMA_DMC
    rng = np.random.default_rng(3)
    n, K = 128, 3
    p = _path(n, K, raw=(84, 84, 3), luma=LUMA_DMC)
    ring, head = orc.new_state(n, K, (84, 84))
    for step in range(5):
        f = rng.integers(0, 256, (n, 84, 84, 3), dtype=np.uint8)
        flags = np.full(n, 5, np.uint8) if step == 0 else _ragged_flags(rng, n, atari=False)
        p.ingest_dmc(f, flags)
        orc.ingest_dmc(f, flags, ring, head)
        assert np.array_equal(_np(p.head), head)
        assert np.array_equal(_np(p.ring), ring)


def _fill(p, rng, n, K, steps=None):
    """Random ring contents through the real ingest (gray frames), mirrored in the oracle."""
    ring, head = orc.new_state(n, K, (84, 84))
    for step in range(steps or K + 1):
        fa = rng.integers(0, 256, (n, 210, 160), dtype=np.uint8)
        fb = rng.integers(0, 256, (n, 210, 160), dtype=np.uint8)
        flags = np.full(n, 5, np.uint8) if step == 0 else _ragged_flags(rng, n)
        p.ingest_atari(fa, fb, flags)
        orc.ingest_atari(fa, fb, flags, ring, head)
    return ring, head


@pytest.mark.parametrize("mode", ["absolute", "relative"])
@pytest.mark.parametrize("fov,K", [((30, 30), 4), ((20, 36), 4), ((50, 50), 2), ((83, 1), 4), ((7, 83), 1), ((29, 31), 3)])
def test_observe_fixed_crop_mask_bit_exact(mode, fov, K):
    rng = np.random.default_rng(5)
    n = 64
    p = _path(n, K, fov_size=fov, fov_init_loc=(1, 0), sensory_action_mode=mode, sensory_action_space=(-10.0, 10.0))
    ring, head = _fill(p, rng, n, K)
    loc = np.tile(np.array([[1, 0]], np.int32), (n, 1))
    p.observe_fixed(None, ctrl="reset")
    for step in range(4):
        if mode == "absolute":
            a = rng.uniform(-5, 90, (n, 2))
            a[::3] = np.floor(a[::3]) + 0.5  # ties -> round half to even
        else:
            a = rng.uniform(-14, 14, (n, 2))
            a[::3] = np.floor(a[::3]) + 0.5
        orc.update_loc(a, loc, obs_size=(84, 84), fov_size=fov, relative=(mode == "relative"), lo=-10.0, hi=10.0)
        for variant in ("crop", "mask"):
            got = _np(p.observe_fixed(a if variant == "crop" else None, variant=variant,
                                      ctrl=None if variant == "crop" else np.full(n, 2, np.uint8)))
            assert np.array_equal(_np(p.loc), loc), (step, variant)
            assert np.array_equal(got, orc.observe_fixed(ring, head, loc, fov, variant=variant)), (step, variant)


def test_observe_fixed_resize_full_within_half_lsb():
    rng = np.random.default_rng(6)
    n, K, fov = 48, 4, (30, 30)
    p = _path(n, K, fov_size=fov)
    ring, head = _fill(p, rng, n, K)
    loc = np.zeros((n, 2), np.int32)
    a = rng.integers(0, 55, (n, 2)).astype(np.float64)
    orc.update_loc(a, loc, obs_size=(84, 84), fov_size=fov)
    got = _np(p.observe_fixed(a, variant="resize_full")).astype(np.float64)
    want = orc.observe_fixed(ring, head, loc, fov, variant="resize_full")
    assert np.abs(got - want).max() <= TOL


@pytest.mark.parametrize("use_cache", [True, False])
@pytest.mark.parametrize("periph,fov", [((20, 20), (30, 30)), ((16, 24), (24, 40)), ((42, 42), (10, 10))])
def test_observe_peripheral(use_cache, periph, fov):
    rng = np.random.default_rng(7)
    n, K = 48, 4
    p = _path(n, K, fov_size=fov, peripheral_res=periph, sensory_action_mode="relative",
              sensory_action_space=(-10.0, 10.0), cache_peripheral=use_cache)
    ring, head = _fill(p, rng, n, K, steps=7)
    loc = np.zeros((n, 2), np.int32)
    p.observe_peripheral(None, ctrl="reset", use_cache=use_cache)
    for step in range(3):
        a = rng.integers(-10, 11, (n, 2)).astype(np.float64)
        orc.update_loc(a, loc, obs_size=(84, 84), fov_size=fov, relative=True, lo=-10.0, hi=10.0)
        got = _np(p.observe_peripheral(a, use_cache=use_cache))
        want = orc.observe_peripheral(ring, head, loc, fov, periph)
        assert np.array_equal(_np(p.loc), loc)
        assert np.abs(got.astype(np.float64) - want).max() <= TOL
        # the pasted fovea is bit exact (fov_env.py:385-386)
        full = orc.stack(ring, head)
        for e in range(0, n, 7):
            r, c = loc[e]
            assert np.array_equal(got[e, :, r:r + fov[0], c:c + fov[1]], full[e, :, r:r + fov[0], c:c + fov[1]])


@pytest.mark.parametrize("variant", ["mask", "resize_full", "crop"])
def test_observe_flexible(variant):
    rng = np.random.default_rng(8)
    n, K, fov = 64, 4, (30, 30)
    p = _path(n, K, fov_size=fov, sensory_action_mode="absolute")
    ring, head = _fill(p, rng, n, K)
    loc = np.zeros((n, 2), np.int32)
    res = np.tile(np.array([fov], np.int32), (n, 1))
    p.observe_flexible(None, variant=variant, ctrl="reset")
    for step in range(4):
        atype = rng.integers(0, 2, n).astype(np.int32)
        a = np.where(atype[:, None] == 1, rng.integers(20, 51, (n, 2)), rng.integers(0, 65, (n, 2))).astype(np.float64)
        if step == 2:  # extreme windows: 1 px, full frame, tall/narrow
            a[:6] = [[1, 1], [84, 84], [84, 1], [1, 84], [31, 2], [2, 31]]
            atype[:6] = 1
        orc.update_loc(a, loc, obs_size=(84, 84), fov_size=fov, atype=atype, res=res)
        got = _np(p.observe_flexible(a, atype, variant=variant)).astype(np.float64)
        want = orc.observe_flexible(ring, head, loc, res, fov, variant=variant)
        assert np.array_equal(_np(p.loc), loc) and np.array_equal(_np(p.res), res), step
        assert np.abs(got - want).max() <= TOL, step
        # windows that are not blurred (rows <= fov rows) are bit exact
        sharp = res[:, 0] <= fov[0]
        if variant != "resize_full":
            assert np.array_equal(got[sharp], want[sharp])


def test_full_size_properties_config2_and_config4():
    """BASELINE sizes (N=4096 RGB ingest; N=16384 peripheral): properties that need no oracle."""
    n, K = 4096, 4
    p = _path(n, K, raw=(210, 160, 3), fov_size=(30, 30), sensory_action_mode="relative",
              sensory_action_space=(-10.0, 10.0))
    fa = torch.empty(p.raw_frame_shape(), dtype=torch.uint8, device="cuda")
    fb = torch.empty_like(fa)
    hard = torch.full((n,), 5, dtype=torch.uint8, device="cuda")
    both = torch.full((n,), 3, dtype=torch.uint8, device="cuda")
    p.synth_frames(fa, 1); p.synth_frames(fb, 2)
    p.ingest_atari(fa, fb, hard)
    first = p.stack().clone()
    assert int(first[:, :-1].max()) == 0, "hard reset leaves K-1 zero frames"
    # constant frames resize to the same constant; max-pool takes the brighter frame
    fa.fill_(17); fb.fill_(200)
    p.ingest_atari(fa, fb, both)
    s = p.stack()
    assert bool((s[:, -1] == 200).all()) and bool((s[:, -2] == first[:, -1]).all())
    # idempotence of KEEP + crop == slicing the stack at loc
    a = torch.randint(-10, 11, (n, 2), device="cuda").double()
    crop = p.observe_fixed(a)
    loc = p.loc.clone()
    again = p.observe_fixed(None, ctrl=torch.full((n,), 2, dtype=torch.uint8, device="cuda"))
    assert torch.equal(crop, again) and torch.equal(loc, p.loc)
    idx = torch.arange(0, n, 257, device="cuda")
    for e in idx.tolist():
        r, c = loc[e].tolist()
        assert torch.equal(crop[e], s[e, :, r:r + 30, c:c + 30])
    # oracle spot-check of 8 envs out of the full batch
    sel = np.arange(0, n, 512)
    p.synth_frames(fa, 3); p.synth_frames(fb, 4)
    ring_before = p.ring[sel].cpu().numpy().copy(); head_before = p.head[sel].cpu().numpy().copy()
    p.ingest_atari(fa, fb, both)
    orc.ingest_atari(fa[sel].cpu().numpy(), fb[sel].cpu().numpy(), np.full(len(sel), 3, np.uint8), ring_before, head_before)
    assert np.array_equal(p.ring[sel].cpu().numpy(), ring_before)

    n4 = 16384
    q = _path(n4, K, fov_size=(30, 30), peripheral_res=(20, 20))
    ga = torch.empty(q.raw_frame_shape(), dtype=torch.uint8, device="cuda")
    for i in range(K):
        q.synth_frames(ga, 10 + i)
        q.ingest_atari(ga, ga, torch.full((n4,), 5 if i == 0 else 3, dtype=torch.uint8, device="cuda"))
    act = torch.randint(0, 55, (n4, 2), device="cuda").double()
    cached = q.observe_peripheral(act)
    direct = q.observe_peripheral(None, ctrl=torch.full((n4,), 2, dtype=torch.uint8, device="cuda"), use_cache=False)
    assert (cached.int() - direct.int()).abs().max().item() <= 1
    assert (cached != direct).float().mean().item() < 1e-2, "cached and recomputed squeeze agree up to rare ties"
    sel = np.arange(0, n4, 2048)
    want = orc.observe_peripheral(q.ring[sel].cpu().numpy(), q.head[sel].cpu().numpy(), q.loc[sel].cpu().numpy(), (30, 30), (20, 20))
    assert np.abs(cached[sel].cpu().numpy().astype(np.float64) - want).max() <= TOL


def test_normalize_is_the_reference_float32_value():
    # atari_env.py:75 / dmc_env.py:183: float32(u) / 255, bit for bit; f16 / bf16 round that value once
    rng = np.random.default_rng(4)
    p = _path(8, 4, fov_size=(30, 30))
    u = torch.from_numpy(rng.integers(0, 256, (8, 4, 30, 30), dtype=np.uint8)).cuda()
    want = u.cpu().numpy().astype(np.float32) / np.float32(255.0)
    got = _np(p.normalize(u))
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert torch.equal(p.normalize(u, torch.float16).cpu(), torch.from_numpy(want).to(torch.float16))
    assert torch.equal(p.normalize(u, torch.bfloat16).cpu(), torch.from_numpy(want).to(torch.bfloat16))
    allv = torch.arange(256, dtype=torch.uint8).repeat(2).cuda()  # every value
    assert np.array_equal(_np(p.normalize(allv))[:256], np.arange(256, dtype=np.float32) / np.float32(255.0))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_normalised_second_output_of_the_observe_kernels_is_bit_identical_to_normalize(dtype):
    """the observe kernels write float32(u) / 255 (atari_env.py:75) as a second output from the tile /
    words they hold — the standard-geometry kernels in the same launch (peripheral_std, flexible_v3, crop_v2), the
    table-driven ones through a normalise pass behind them — bit-identical to agym_normalize of the u8 output."""
    rng = np.random.default_rng(8)
    n, K = 333, 4
    p = _path(n, K, fov_size=(30, 30), peripheral_res=(20, 20), sensory_action_mode="absolute")
    for step in range(K + 1):
        fa = rng.integers(0, 256, (n, 210, 160), dtype=np.uint8)
        p.ingest_atari(fa, fa, np.full(n, 5 if step == 0 else 3, np.uint8))
    a = rng.integers(0, 55, (n, 2)).astype(np.float64)
    at = rng.integers(0, 2, n).astype(np.int32)
    af = np.where(at[:, None] == 1, rng.integers(20, 51, (n, 2)), a).astype(np.float64)
    cases = [
        ("peripheral", lambda no: p.observe_peripheral(a, norm_out=no)),
        ("peripheral uncached (generic kernel + pass)", lambda no: p.observe_peripheral(a, use_cache=False, norm_out=no)),
        ("crop", lambda no: p.observe_fixed(a, variant="crop", norm_out=no)),
        ("mask (generic kernel + pass)", lambda no: p.observe_fixed(a, variant="mask", norm_out=no)),
        ("flexible mask", lambda no: p.observe_flexible(af, at, variant="mask", norm_out=no)),
        ("flexible crop", lambda no: p.observe_flexible(af, at, variant="crop", pad=(84, 84), norm_out=no)),
        ("flexible resize_full (generic kernel + pass)", lambda no: p.observe_flexible(af, at, variant="resize_full", norm_out=no)),
    ]
    for name, call in cases:
        shape = call(None).shape
        no = torch.full(shape, -1.0, dtype=dtype, device="cuda")
        out = call(no)
        want = p.normalize(out, dtype)
        assert torch.equal(no.view(torch.int16 if dtype != torch.float32 else torch.int32),
                           want.view(torch.int16 if dtype != torch.float32 else torch.int32)), name
    # every u8 value: the FMA-corrected quotient of the fused stores == the IEEE division, bit for bit
    q = _path(64, 4, fov_size=(30, 30))
    q.ring.copy_(torch.arange(64 * 4 * 7056, device="cuda").remainder(256).to(torch.uint8).view(64, 4, 84, 84))
    no = torch.empty((64, 4, 30, 30), dtype=torch.float32, device="cuda")
    out = q.observe_fixed(np.zeros((64, 2)), variant="crop", norm_out=no)
    ref = out.cpu().numpy().astype(np.float32) / np.float32(255.0)
    assert set(np.unique(out.cpu().numpy()).tolist()) == set(range(256))
    assert np.array_equal(no.cpu().numpy().view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("kind", ["peripheral_std", "peripheral_other", "peripheral_uncached", "flexible_mask", "flexible_resize_full"])
def test_plain_bilinear_mode_matches_the_oracle_without_antialias(kind):
    """antialias=False (older torchvision releases resized tensors without antialiasing): the same
    kernels on plain-bilinear tables against the oracle's restatement of F.interpolate(..., antialias=False), which a CPU
    test pins against torch itself.  The results must really differ from the antialiased mode."""
    rng = np.random.default_rng(31)
    n, K, fov = 40, 4, (30, 30)
    periph = (16, 24) if kind == "peripheral_other" else (20, 20)
    kw = dict(fov_size=fov, sensory_action_mode="absolute", antialias=False)
    if kind.startswith("peripheral"):
        kw.update(peripheral_res=periph, cache_peripheral=kind != "peripheral_uncached")
    p = _path(n, K, **kw)
    ring, head = _fill(p, rng, n, K, steps=6)
    loc = np.zeros((n, 2), np.int32)
    res = np.tile(np.array([fov], np.int32), (n, 1))
    orc.set_antialias(False)
    try:
        for step in range(3):
            if kind.startswith("peripheral"):
                a = rng.integers(0, 55, (n, 2)).astype(np.float64)
                orc.update_loc(a, loc, obs_size=(84, 84), fov_size=fov)
                got = _np(p.observe_peripheral(a, use_cache=kind != "peripheral_uncached")).astype(np.float64)
                want = orc.observe_peripheral(ring, head, loc, fov, periph)
                orc.set_antialias(True)
                other = orc.observe_peripheral(ring, head, loc, fov, periph)
                orc.set_antialias(False)
            else:
                variant = kind.split("_", 1)[1]
                atype = rng.integers(0, 2, n).astype(np.int32)
                a = np.where(atype[:, None] == 1, rng.integers(20, 51, (n, 2)), rng.integers(0, 65, (n, 2))).astype(np.float64)
                orc.update_loc(a, loc, obs_size=(84, 84), fov_size=fov, atype=atype, res=res)
                got = _np(p.observe_flexible(a, atype, variant=variant)).astype(np.float64)
                want = orc.observe_flexible(ring, head, loc, res, fov, variant=variant)
                orc.set_antialias(True)
                other = orc.observe_flexible(ring, head, loc, res, fov, variant=variant)
                orc.set_antialias(False)
            assert np.array_equal(_np(p.loc), loc)
            assert np.abs(got - want).max() <= TOL, (kind, step)
            if step == 2 and kind != "flexible_resize_full":
                assert np.abs(other - want).max() > 2.0, "the antialiased mode gives visibly different pixels"
    finally:
        orc.set_antialias(True)

"""GPU: oracle parity AT THE BASELINE BATCH SIZES (16,384 / 8,192 / 4,096 envs) for >= 1,024 envs per case — the first
and last envs, the envs on both sides of every persistent kernel's CTA stride (env = blockIdx + k * gridDim: 444 CTAs
for the TMA ingest kernels, 296 for the peripheral / flexible kernels, 8 envs per CTA in the crop kernel) and a random
sample — with ragged ingest flags and mixed fovea control (APPLY / RESET / KEEP) on every step.  The per-CTA env walk of
the persistent kernels is ~37 envs deep at these sizes; the small-batch tests never get past 4."""
import numpy as np
import pytest
import torch

from oracle import agym_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 0.5 + 1e-2
S = (84, 84)


def _select(n, rng, want=1100):
    pick = set(range(0, 160)) | set(range(n - 160, n))
    for stride in (444, 296, 148, 8):
        for k in range(0, n, stride):
            pick.update(v for v in (k - 1, k, k + 1) if 0 <= v < n)
        if len(pick) > want:
            break
    rest = np.setdiff1d(np.arange(n), np.fromiter(pick, int))
    extra = max(1024 - len(pick), 64)
    pick.update(rng.choice(rest, size=min(extra, len(rest)), replace=False).tolist())
    sel = np.array(sorted(pick))
    assert len(sel) >= 1024
    return sel


def _flags(rng, n, atari=True):
    choices = np.array([3, 3, 3, 3, 1, 0, 5, 1 | 4, 8, 3 | 4] if atari else [1, 1, 1, 1, 5, 8], np.uint8)
    return choices[rng.integers(0, len(choices), n)]


def _ctrl(rng, n):
    return rng.choice(np.array([0, 0, 0, 0, 1, 2], np.uint8), size=n)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_config3_gray_ingest_and_peripheral_16384_envs():
    from active_gym_b200 import ObservationPath
    rng = np.random.default_rng(1)
    n, K, fov, per = 16384, 4, (30, 30), (20, 20)
    p = ObservationPath(n, K, S, (210, 160, 1), fov_size=fov, peripheral_res=per, sensory_action_mode="relative",
                        sensory_action_space=(-10.0, 10.0), fov_init_loc=(7, 40))
    sel = _select(n, rng)
    tsel = _dev(sel)
    ring, head = orc.new_state(len(sel), K, S)
    loc = np.zeros((len(sel), 2), np.int32)
    fa = torch.empty(p.raw_frame_shape(), dtype=torch.uint8, device="cuda")
    fb = torch.empty_like(fa)
    for step in range(6):
        p.synth_frames(fa, 100 + 2 * step); p.synth_frames(fb, 101 + 2 * step)
        fl = np.full(n, 5, np.uint8) if step == 0 else _flags(rng, n)
        p.ingest_atari(fa, fb, _dev(fl))
        orc.ingest_atari(fa[tsel].cpu().numpy(), fb[tsel].cpu().numpy(), fl[sel], ring, head)
        assert np.array_equal(p.head[tsel].cpu().numpy(), head), step
        assert np.array_equal(p.ring[tsel].cpu().numpy(), ring), step
        ctrl = np.full(n, 1, np.uint8) if step == 0 else _ctrl(rng, n)
        a = rng.uniform(-14, 14, (n, 2))
        tie = rng.random(n) < 0.2
        a[tie] = np.round(a[tie]) + 0.5   # ties: round half to even
        got = p.observe_peripheral(_dev(a), ctrl=_dev(ctrl))
        new = loc.copy()
        orc.update_loc(a[sel], new, obs_size=S, fov_size=fov, relative=True, lo=-10.0, hi=10.0)
        c = ctrl[sel]
        loc[c == 0] = new[c == 0]
        loc[c == 1] = (7, 40)
        assert np.array_equal(p.loc[tsel].cpu().numpy(), loc), step
        want = orc.observe_peripheral(ring, head, loc, fov, per)
        assert np.abs(got[tsel].cpu().numpy().astype(np.float64) - want).max() <= TOL, step
        # the fovea itself is a bit-exact paste
        e = 17
        r, cc = loc[e]
        assert np.array_equal(got[int(sel[e])].cpu().numpy()[:, r:r + 30, cc:cc + 30],
                              orc.stack(ring, head)[e][:, r:r + 30, cc:cc + 30])


def test_config1_rgb_ingest_and_crop_4096_envs_config2_flexible():
    from active_gym_b200 import ObservationPath
    rng = np.random.default_rng(2)
    n, K, fov = 4096, 4, (30, 30)
    p = ObservationPath(n, K, S, (210, 160, 3), fov_size=fov, sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0))
    sel = _select(n, rng)
    tsel = _dev(sel)
    ring, head = orc.new_state(len(sel), K, S)
    loc = np.zeros((len(sel), 2), np.int32)
    res = np.tile(np.array([fov], np.int32), (len(sel), 1))
    fa = torch.empty(p.raw_frame_shape(), dtype=torch.uint8, device="cuda")
    fb = torch.empty_like(fa)
    for step in range(5):
        p.synth_frames(fa, 200 + 2 * step); p.synth_frames(fb, 201 + 2 * step)
        fl = np.full(n, 5, np.uint8) if step == 0 else _flags(rng, n)
        p.ingest_atari(fa, fb, _dev(fl))
        orc.ingest_atari(fa[tsel].cpu().numpy(), fb[tsel].cpu().numpy(), fl[sel], ring, head)
        assert np.array_equal(p.ring[tsel].cpu().numpy(), ring) and np.array_equal(p.head[tsel].cpu().numpy(), head), step
        ctrl = np.full(n, 1, np.uint8) if step == 0 else _ctrl(rng, n)
        a = rng.uniform(-14, 14, (n, 2))
        got = p.observe_fixed(_dev(a), variant="crop", ctrl=_dev(ctrl))
        new = loc.copy()
        orc.update_loc(a[sel], new, obs_size=S, fov_size=fov, relative=True, lo=-10.0, hi=10.0)
        c = ctrl[sel]
        loc[c == 0] = new[c == 0]
        loc[c == 1] = 0
        assert np.array_equal(p.loc[tsel].cpu().numpy(), loc), step
        assert np.array_equal(got[tsel].cpu().numpy(), orc.observe_fixed(ring, head, loc, fov)), step
    # configs[2]: the flexible kernel over the same ring, absolute mode semantics come from a second path object
    q = ObservationPath(n, K, S, (210, 160, 1), fov_size=fov, sensory_action_mode="absolute")
    q.ring.copy_(p.ring); q.head.copy_(p.head)
    loc[:] = 0
    q.observe_flexible(None, variant="mask", ctrl="reset")
    for step in range(4):
        ctrl = _ctrl(rng, n)
        atype = rng.integers(0, 2, n).astype(np.int32)
        a = np.where(atype[:, None] == 1, rng.integers(20, 51, (n, 2)), rng.uniform(-5, 70, (n, 2))).astype(np.float64)
        got = q.observe_flexible(_dev(a), _dev(atype), variant="mask", ctrl=_dev(ctrl))
        nl, nr = loc.copy(), res.copy()
        orc.update_loc(a[sel], nl, obs_size=S, fov_size=fov, atype=atype[sel], res=nr)
        c = ctrl[sel]
        loc[c == 0], res[c == 0] = nl[c == 0], nr[c == 0]
        loc[c == 1], res[c == 1] = 0, fov
        assert np.array_equal(q.loc[tsel].cpu().numpy(), loc) and np.array_equal(q.res[tsel].cpu().numpy(), res), step
        want = orc.observe_flexible(ring, head, loc, res, fov, variant="mask")
        g = got[tsel].cpu().numpy().astype(np.float64)
        assert np.abs(g - want).max() <= TOL, step
        sharp = res[:, 0] <= fov[0]
        assert np.array_equal(g[sharp], want[sharp]), step
    assert q.read_errors() == 0


def test_config4_dmc_ingest_and_crop_8192_envs():
    from active_gym_b200 import LUMA_DMC, ObservationPath
    rng = np.random.default_rng(3)
    n, K, fov = 8192, 3, (30, 30)
    p = ObservationPath(n, K, S, (84, 84, 3), luma=LUMA_DMC, fov_size=fov, sensory_action_mode="absolute")
    sel = _select(n, rng)
    tsel = _dev(sel)
    ring, head = orc.new_state(len(sel), K, S)
    loc = np.zeros((len(sel), 2), np.int32)
    f = torch.empty(p.raw_frame_shape(), dtype=torch.uint8, device="cuda")
    for step in range(5):
        p.synth_frames(f, 300 + step)
        fl = np.full(n, 5, np.uint8) if step == 0 else _flags(rng, n, atari=False)
        p.ingest_dmc(f, _dev(fl))
        orc.ingest_dmc(f[tsel].cpu().numpy(), fl[sel], ring, head)
        assert np.array_equal(p.ring[tsel].cpu().numpy(), ring) and np.array_equal(p.head[tsel].cpu().numpy(), head), step
        ctrl = np.full(n, 1, np.uint8) if step == 0 else _ctrl(rng, n)
        a = rng.uniform(-5, 60, (n, 2))
        for variant in ("crop", "mask"):
            got = p.observe_fixed(_dev(a), variant=variant, ctrl=_dev(ctrl if variant == "crop" else np.full(n, 2, np.uint8)))
            if variant == "crop":
                new = loc.copy()
                orc.update_loc(a[sel], new, obs_size=S, fov_size=fov)
                c = ctrl[sel]
                loc[c == 0] = new[c == 0]
                loc[c == 1] = 0
            assert np.array_equal(p.loc[tsel].cpu().numpy(), loc), (step, variant)
            assert np.array_equal(got[tsel].cpu().numpy(), orc.observe_fixed(ring, head, loc, fov, variant=variant)), (step, variant)

"""Out-of-bounds WRITES of every kernel family, caught with guard bands (compute-sanitizer is closed on the GPU pool).

Every buffer a kernel writes — ring, head, loc, res, squeeze cache, the u8 output and the normalised output — is the
interior of a larger allocation whose borders hold a canary pattern; after a few steps with ragged flags, windows at the
frame's edge and env counts that leave the last CTA / warp partly empty, the borders must be untouched and the results
must still equal the oracle's (so the interior was written where it should be).  Out-of-bounds READS cannot be seen this
way; the kernels' deliberate over-reads stay inside a plane's 16-byte hull.
"""
import numpy as np
import pytest
import torch

from oracle import agym_oracle as orc

pytestmark = pytest.mark.gpu

GUARD = 4096          # bytes on either side; a multiple of 256 keeps the interior as aligned as a fresh allocation
CANARY = 0xA5
S = (84, 84)


class Guarded:
    """A contiguous device tensor of `shape` / `dtype` between two canary bands."""

    def __init__(self, shape, dtype, dev, fill=0):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        self.pad = (-n) % 256
        self.raw = torch.full((GUARD + n + self.pad + GUARD,), CANARY, dtype=torch.uint8, device=dev)
        self.t = self.raw[GUARD:GUARD + n].view(dtype).view(shape)
        self.t.fill_(fill)
        self.n = n

    def intact(self):
        return bool((self.raw[:GUARD] == CANARY).all()) and bool((self.raw[GUARD + self.n:] == CANARY).all())


def _path(n, K, raw, fov, periph=None, mode="relative", dev="cuda:0"):
    from active_gym_b200 import LUMA_DMC, LUMA_RGB, ObservationPath
    g = dict(ring=Guarded((n, K) + S, torch.uint8, dev), head=Guarded((n,), torch.int32, dev, K - 1),
             loc=Guarded((n, 2), torch.int32, dev), res=Guarded((n, 2), torch.int32, dev))
    g["res"].t[:, 0], g["res"].t[:, 1] = fov
    if periph:
        g["pcache"] = Guarded((n, K) + periph, torch.float32, dev)
    p = ObservationPath(n, K, S, raw, luma=LUMA_DMC if raw[0] == S[0] else LUMA_RGB, fov_size=fov, peripheral_res=periph, sensory_action_mode=mode,
                        sensory_action_space=(-10.0, 10.0), device=dev, buffers={k: v.t for k, v in g.items()})
    return p, g


def _flags(n, step, rng):
    fl = np.full(n, 5 if step == 0 else 3, np.uint8)
    if step:
        fl[rng.random(n) < 0.1] = 8      # idle
        fl[rng.random(n) < 0.1] = 1      # early game over: one frame
        fl[rng.random(n) < 0.05] = 5     # hard reset: K zero frames, then one un-pooled frame
    return fl


def _edge_actions(n, rng, hi):
    """Mostly random targets, with the corners of the valid range and far overshoots mixed in."""
    a = rng.integers(0, hi + 1, (n, 2)).astype(np.float64)
    a[0::9] = (0, 0); a[1::9] = (hi, hi); a[2::9] = (hi, 0); a[3::9] = (-1e6, 1e6)
    return a


@pytest.mark.parametrize("n", [1, 37, 601])
def test_peripheral_step_stays_inside_its_buffers(n):
    rng = np.random.default_rng(n)
    fov, periph, K = (30, 30), (20, 20), 4
    p, g = _path(n, K, (210, 160, 1), fov, periph, mode="absolute")
    out, nrm = Guarded((n, K) + S, torch.uint8, "cuda:0"), Guarded((n, K) + S, torch.float16, "cuda:0")
    ring, head = orc.new_state(n, K, S)
    loc = np.zeros((n, 2), np.int32)
    p.observe_peripheral(None, ctrl="reset", out=out.t)
    for step in range(3):
        fa, fb = rng.integers(0, 256, (2, n, 210, 160), dtype=np.uint8)
        fl = _flags(n, step, rng)
        a = _edge_actions(n, rng, 54)
        p.ingest_atari(fa, fb, fl)
        orc.ingest_atari(fa, fb, fl, ring, head)
        p.observe_peripheral(a, out=out.t, norm_out=nrm.t)
        orc.update_loc(a, loc, obs_size=S, fov_size=fov, relative=False, lo=-10.0, hi=10.0)
    torch.cuda.synchronize()
    assert all(v.intact() for v in g.values()) and out.intact() and nrm.intact()
    assert np.array_equal(g["ring"].t.cpu().numpy(), ring) and np.array_equal(g["loc"].t.cpu().numpy(), loc)
    want = orc.observe_peripheral(ring, head, loc, fov, periph)
    assert np.abs(out.t.cpu().numpy().astype(np.float64) - want).max() <= 0.5 + 1e-2


@pytest.mark.parametrize("n,K,raw,fov", [(1, 4, (210, 160, 3), (30, 30)), (45, 4, (210, 160, 3), (30, 30)), (1001, 3, (84, 84, 3), (30, 30)),
                                         (13, 3, (84, 84, 3), (21, 33)), (70, 4, (210, 160, 1), (50, 28)), (70, 4, (210, 160, 1), (26, 22))])
def test_fixed_crop_and_mask_stay_inside_their_buffers(n, K, raw, fov):
    rng = np.random.default_rng(n + K)
    dmc = raw[0] == 84
    p, g = _path(n, K, raw, fov, mode="absolute")
    out = Guarded((n, K) + fov, torch.uint8, "cuda:0")
    nrm = Guarded((n, K) + fov, torch.float32, "cuda:0") if (n * K * fov[0] * fov[1]) % 16 == 0 else None
    msk = Guarded((n, K) + S, torch.uint8, "cuda:0")
    ring, head = orc.new_state(n, K, S)
    loc = np.zeros((n, 2), np.int32)
    p.observe_fixed(None, ctrl="reset", out=out.t)
    for step in range(3):
        fl = _flags(n, step, rng)
        if dmc:
            fl = np.where(fl == 8, 8, np.where(fl & 4, 5, 1)).astype(np.uint8)
            f = rng.integers(0, 256, (n,) + raw, dtype=np.uint8)
            p.ingest_dmc(f, fl)
            orc.ingest_dmc(f, fl, ring, head)
        else:
            fa, fb = rng.integers(0, 256, (2, n) + (raw if raw[2] == 3 else raw[:2]), dtype=np.uint8)
            p.ingest_atari(fa, fb, fl)
            orc.ingest_atari(fa, fb, fl, ring, head)
        a = np.stack([_edge_actions(n, rng, S[0] - fov[0])[:, 0], _edge_actions(n, rng, S[1] - fov[1])[:, 1]], 1)
        p.observe_fixed(a, out=out.t, norm_out=None if nrm is None else nrm.t)
        p.observe_fixed(None, variant="mask", ctrl=np.full(n, 2, np.uint8), out=msk.t)
        orc.update_loc(a, loc, obs_size=S, fov_size=fov, relative=False, lo=-10.0, hi=10.0)
    torch.cuda.synchronize()
    assert all(v.intact() for v in g.values()) and out.intact() and msk.intact() and (nrm is None or nrm.intact())
    assert np.array_equal(g["ring"].t.cpu().numpy(), ring) and np.array_equal(g["loc"].t.cpu().numpy(), loc)
    assert np.array_equal(out.t.cpu().numpy(), orc.observe_fixed(ring, head, loc, fov))
    assert np.array_equal(msk.t.cpu().numpy(), orc.observe_fixed(ring, head, loc, fov, variant="mask"))


@pytest.mark.parametrize("n,variant", [(1, "mask"), (333, "mask"), (333, "crop"), (29, "resize_full")])
def test_flexible_step_stays_inside_its_buffers(n, variant):
    rng = np.random.default_rng(n)
    fov, K = (30, 30), 4
    p, g = _path(n, K, (210, 160, 1), fov, mode="absolute")
    out = Guarded((n, K) + S, torch.uint8, "cuda:0")
    ring, head = orc.new_state(n, K, S)
    loc = np.zeros((n, 2), np.int32)
    res = np.tile(np.array([fov], np.int32), (n, 1))
    p.observe_flexible(None, variant=variant, ctrl="reset", out=out.t)
    for step in range(4):
        fa, fb = rng.integers(0, 256, (2, n, 210, 160), dtype=np.uint8)
        fl = _flags(n, step, rng)
        p.ingest_atari(fa, fb, fl)
        orc.ingest_atari(fa, fb, fl, ring, head)
        atype = rng.integers(0, 2, n).astype(np.int32)
        a = np.where(atype[:, None] == 1, rng.integers(1, 85, (n, 2)), _edge_actions(n, rng, 83)).astype(np.float64)
        if step == 2 and n >= 6:   # extreme windows: 1 px, full frame, tall / narrow
            a[:6] = [[1, 1], [84, 84], [84, 1], [1, 84], [31, 2], [2, 31]]
            atype[:6] = 1
        p.observe_flexible(a, atype, variant=variant, out=out.t)
        orc.update_loc(a, loc, obs_size=S, fov_size=fov, atype=atype, res=res)
    torch.cuda.synchronize()
    assert all(v.intact() for v in g.values()) and out.intact()
    assert np.array_equal(g["loc"].t.cpu().numpy(), loc) and np.array_equal(g["res"].t.cpu().numpy(), res)
    want = orc.observe_flexible(ring, head, loc, res, fov, variant=variant)
    assert np.abs(out.t.cpu().numpy().astype(np.float64) - want).max() <= 0.5 + 1e-2


@pytest.mark.parametrize("kind", ["peripheral", "flexible", "fixed_rgb", "dmc"])
def test_a_step_is_deterministic_whatever_the_scheduling(kind):
    """Racecheck substitute: the same step from the same state, six times, must give bit-identical state and outputs —
    the persistent kernels hand envs to CTAs dynamically (atomic claims, whichever CTA is free), so a shared-memory or
    pipeline race shows up as a run-to-run difference."""
    rng = np.random.default_rng(11)
    n, fov = 2500, (30, 30)
    K = 3 if kind == "dmc" else 4
    raw = {"peripheral": (210, 160, 1), "flexible": (210, 160, 1), "fixed_rgb": (210, 160, 3), "dmc": (84, 84, 3)}[kind]
    p, g = _path(n, K, raw, fov, (20, 20) if kind == "peripheral" else None, mode="relative" if kind != "flexible" else "absolute")

    def frames():
        if kind == "dmc":
            return (rng.integers(0, 256, (n,) + raw, dtype=np.uint8),)
        return tuple(rng.integers(0, 256, (2, n) + (raw if raw[2] == 3 else raw[:2]), dtype=np.uint8))

    def ingest(fr, fl):
        p.ingest_dmc(fr[0], fl) if kind == "dmc" else p.ingest_atari(fr[0], fr[1], fl)

    def observe(a, at, ctrl=None):
        if kind == "peripheral":
            return p.observe_peripheral(a, ctrl=ctrl)
        if kind == "flexible":
            return p.observe_flexible(a, at, variant="mask", ctrl=ctrl)
        return p.observe_fixed(a, ctrl=ctrl)

    first = np.full(n, 5, np.uint8)
    ingest(frames(), first)
    observe(None, None, ctrl="reset")
    for _ in range(K):
        ingest(frames(), np.full(n, 1 if kind == "dmc" else 3, np.uint8))
    torch.cuda.synchronize()
    state0 = {k: v.t.clone() for k, v in g.items()}
    fr = frames()
    fl = _flags(n, 1, rng)
    if kind == "dmc":
        fl = np.where(fl == 8, 8, np.where(fl & 4, 5, 1)).astype(np.uint8)
    at = rng.integers(0, 2, n).astype(np.int32)
    a = rng.integers(-10, 11, (n, 2)).astype(np.float64)
    if kind == "flexible":
        a = np.where(at[:, None] == 1, rng.integers(20, 51, (n, 2)), rng.integers(0, 55, (n, 2))).astype(np.float64)
    ref = None
    for rep in range(6):
        for k, v in g.items():
            v.t.copy_(state0[k])
        ingest(fr, fl)
        out = observe(a, at)
        torch.cuda.synchronize()
        snap = [out.clone()] + [v.t.clone() for v in g.values()]
        if ref is None:
            ref = snap
        else:
            assert all(torch.equal(x, y) for x, y in zip(ref, snap)), rep
    assert all(v.intact() for v in g.values())

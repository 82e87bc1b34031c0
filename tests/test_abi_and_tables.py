"""CPU: the C-ABI library loads without a GPU and exports every symbol the header declares; the
host-side coefficient tables it would upload reproduce the oracle's (= the reference libraries')
arithmetic when applied in NumPy."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from active_gym_b200 import _lib
from oracle import agym_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "agym_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(agym_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    L = C.CDLL(_lib.LIB_PATH)  # loads with no CUDA device present
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == declared
    assert _lib.lib().agym_abi_version() == _lib.ABI_VERSION
    assert _lib.lib().agym_status_string(-2).decode().startswith("geometry")


def test_plan_create_rejects_bad_geometry_without_touching_the_gpu():
    L = _lib.lib()
    cfg = _lib.Config()
    cfg.n_envs, cfg.frame_stack, cfg.obs_h, cfg.obs_w = 4, 4, 84, 84
    cfg.raw_h, cfg.raw_w, cfg.raw_c = 210, 160, 1
    cfg.fov_h, cfg.fov_w = 84, 30  # fov must be strictly smaller than obs (fov_env.py:112)
    plan = C.c_void_p()
    assert L.agym_plan_create(C.byref(cfg), C.byref(plan)) == -1
    cfg.fov_h = 30
    cfg.fov_init_loc[:] = [60.0, 0.0]  # out-of-range init: the reference would return a short crop
    assert L.agym_plan_create(C.byref(cfg), C.byref(plan)) == -2
    cfg.fov_init_loc[:] = [0.0, 0.0]
    cfg.obs_w = 83  # not a multiple of 4
    assert L.agym_plan_create(C.byref(cfg), C.byref(plan)) == -2
    assert L.agym_plan_create(None, C.byref(plan)) == -1


def _cv2_axis(n_src, n_dst, zero):
    s0, s1, cf = (np.zeros(n_dst, np.int32) for _ in range(3))
    assert _lib.lib().agym_table_cv2(n_src, n_dst, int(zero), s0.ctypes.data, s1.ctypes.data, cf.ctypes.data) == 0
    return s0, s1, cf & 0xffff, cf >> 16


@pytest.mark.parametrize("shape,out", [((210, 160), (84, 84)), ((210, 160), (64, 96)), ((84, 84), (30, 30)), ((100, 120), (84, 84))])
def test_cv2_tables_reproduce_the_fixed_point_resize(shape, out):
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, shape, dtype=np.uint8).astype(np.int64)
    xs0, xs1, c0, c1 = _cv2_axis(shape[1], out[1], True)
    ys0, ys1, b0, b1 = _cv2_axis(shape[0], out[0], False)
    h = src[:, xs0] * c0 + src[:, xs1] * c1                      # horizontal pass, scale 2^11
    v = (((b0[:, None] * (h[ys0] >> 4)) >> 16) + ((b1[:, None] * (h[ys1] >> 4)) >> 16) + 2) >> 2
    assert np.array_equal(v.astype(np.uint8), orc.cv2_resize_linear(src.astype(np.uint8), out))


def _aa_axis(n_in, n_out, antialias=True):
    xmin = np.zeros(n_out, np.int32)
    w = np.zeros(n_out * 128, np.float32)
    taps = C.c_int32()
    assert _lib.lib().agym_table_aa(n_in, n_out, int(antialias), xmin.ctypes.data, w.ctypes.data, w.size, C.byref(taps)) == 0
    return xmin, w[:n_out * taps.value].reshape(n_out, taps.value)


@pytest.mark.parametrize("ish,osh", [((84, 84), (20, 20)), ((20, 20), (84, 84)), ((30, 30), (84, 84)), ((44, 50), (30, 30)),
                                     ((30, 30), (31, 20)), ((1, 84), (30, 30)), ((84, 2), (84, 84)), ((30, 30), (30, 30))])
def test_aa_tables_reproduce_the_antialiased_resize(ish, osh):
    rng = np.random.default_rng(2)
    x = rng.integers(0, 256, ish).astype(np.float64)
    xm, ww = _aa_axis(ish[1], osh[1])
    ym, wh = _aa_axis(ish[0], osh[0])
    assert (xm >= 0).all() and (xm + ww.shape[1] <= ish[1]).all(), "windows stay inside the row (no bounds checks on device)"
    assert np.allclose(ww.sum(1), 1.0, atol=1e-6) and np.allclose(wh.sum(1), 1.0, atol=1e-6)
    t = np.stack([(x[:, xm[i]:xm[i] + ww.shape[1]] * ww[i]).sum(1) for i in range(osh[1])], 1)   # W pass
    y = np.stack([(t[ym[j]:ym[j] + wh.shape[1]] * wh[j][:, None]).sum(0) for j in range(osh[0])], 0)  # H pass
    assert np.abs(y - orc.aa_resize(x, osh)).max() <= 1e-4  # float32 weights vs the float64 oracle, u8 LSB


def _blur_axis(r, f, antialias=True):
    xmin = np.zeros(r, np.int32)
    w = np.zeros(r * 32, np.float32)
    q = np.zeros(r * 32, np.uint16)
    taps, halves = C.c_int32(), C.c_int32()
    assert _lib.lib().agym_table_blur(r, f, int(antialias), xmin.ctypes.data, w.ctypes.data, q.ctypes.data, w.size,
                                      C.byref(taps), C.byref(halves)) == 0
    return xmin, w[:r * taps.value].reshape(r, taps.value), q[:r * halves.value * 8].reshape(r, halves.value * 8)


@pytest.mark.parametrize("r", [1, 2, 7, 29, 30, 31, 35, 44, 50, 63, 84])
def test_blur_tables_compose_the_two_resamples_and_quantise_within_bound(r):
    """The flexible fovea's blur Resize(30) -> Resize(r) as one banded operator (fov_env.py:276-280): the float table
    reproduces the oracle's two-step resample, the 16-bit table (what the CUDA W pass multiplies with) sums to 2^16
    per row and stays within 255 * taps / 2^17 LSB of it on arbitrary u8 rows."""
    f = 30
    rng = np.random.default_rng(r)
    xm, w, q = _blur_axis(r, f)
    taps = w.shape[1]
    assert (xm >= 0).all() and (xm + taps <= r).all()
    x = rng.integers(0, 256, (5, r)).astype(np.float64)
    want = orc.aa_resize(orc.aa_resize(x, (5, f)), (5, r))          # W axis only: rows are kept (5 -> 5 is the identity)
    got = np.stack([(x[:, xm[i]:xm[i] + taps] * w[i]).sum(1) for i in range(r)], 1)
    assert np.abs(got - want).max() <= 1e-4
    assert (q.astype(np.int64).sum(1) == 65536).all() or (w.max() >= 1.0 - 1e-7)   # a lone 1.0 is stored as 65535
    assert (q[:, taps:] == 0).all(), "padding taps carry zero weight (they read past the window)"
    xp = np.concatenate([x, np.zeros((5, 16))], 1)
    gq = np.stack([(xp[:, xm[i]:xm[i] + q.shape[1]] * q[i].astype(np.float64)).sum(1) for i in range(r)], 1) / 65536.0
    assert np.abs(gq - want).max() <= 255.0 * max(taps, 2) / 2 ** 17 + 1e-4   # taps == 1: the lone 65535 / 65536


def test_standard_vertical_resize_never_samples_every_fifth_raw_row():
    """cv2.resize 210 -> 84 (atari_env.py:74): output row 2m reads raw rows {5m, 5m+1}, row 2m+1 reads {5m+3, 5m+4};
    row 5m+2 is never read.  The TMA ingest kernel's strided tensor copies rely on this pattern (the plan re-checks
    it on its own tables before enabling them) and on the two vertical weight pairs it implies."""
    s0, s1, b0, b1 = _cv2_axis(210, 84, False)
    m = np.arange(42)
    assert np.array_equal(s0[0::2], 5 * m) and np.array_equal(s1[0::2], 5 * m + 1)
    assert np.array_equal(s0[1::2], 5 * m + 3) and np.array_equal(s1[1::2], 5 * m + 4)
    used = np.union1d(s0, s1)
    assert len(used) == 168 and not np.any(used % 5 == 2)
    assert set(zip(b0[0::2], b1[0::2])) == {(512, 1536)} and set(zip(b0[1::2], b1[1::2])) == {(1536, 512)}



@pytest.mark.parametrize("ish,osh", [((84, 84), (20, 20)), ((44, 50), (30, 30)), ((30, 30), (44, 50)), ((84, 84), (84, 20))])
def test_plain_bilinear_tables_and_oracle_match_torch_without_antialias(ish, osh):
    """antialias=False (the default of older torchvision releases on tensors): the host tables and
    the oracle both restate F.interpolate(bilinear, align_corners=False, antialias=False), checked against torch itself."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, ish).astype(np.float64)
    want = F.interpolate(torch.from_numpy(x)[None, None], size=osh, mode="bilinear", align_corners=False, antialias=False)[0, 0].numpy()
    orc.set_antialias(False)
    try:
        got_oracle = orc.aa_resize(x, osh)
    finally:
        orc.set_antialias(True)
    assert np.abs(got_oracle - want).max() <= 1e-9
    xm, ww = _aa_axis(ish[1], osh[1], antialias=False)
    ym, wh = _aa_axis(ish[0], osh[0], antialias=False)
    assert ww.shape[1] <= 2 and wh.shape[1] <= 2 and (xm + ww.shape[1] <= ish[1]).all() and (ym + wh.shape[1] <= ish[0]).all()
    t = np.stack([(x[:, xm[i]:xm[i] + ww.shape[1]] * ww[i]).sum(1) for i in range(osh[1])], 1)
    y = np.stack([(t[ym[j]:ym[j] + wh.shape[1]] * wh[j][:, None]).sum(0) for j in range(osh[0])], 0)
    assert np.abs(y - want).max() <= 1e-4
    assert np.abs(orc.aa_resize(x, osh) - want).max() > 1.0 or ish[1] <= osh[1]   # the antialiased result really differs when downscaling


def test_blur_table_without_antialias_composes_the_two_plain_resamples():
    r, f = 44, 30
    rng = np.random.default_rng(9)
    x = rng.integers(0, 256, (3, r)).astype(np.float64)
    orc.set_antialias(False)
    try:
        want = orc.aa_resize(orc.aa_resize(x, (3, f)), (3, r))
    finally:
        orc.set_antialias(True)
    xm, w, q = _blur_axis(r, f, antialias=False)
    got = np.stack([(x[:, xm[i]:xm[i] + w.shape[1]] * w[i]).sum(1) for i in range(r)], 1)
    assert np.abs(got - want).max() <= 1e-4

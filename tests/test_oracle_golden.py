"""CPU: the restated oracle (oracle/agym_oracle.c) against the golden vectors recorded from the
unmodified reference (tests/golden, made by oracle/make_golden.py).  This is the 'pin' of the
oracle: bit-exact for crop / stack / max-pool / mask / cv2 resize / luma, <= 1e-4 u8-LSB for the
torchvision resamples (the fixture stores float32(255*ref))."""
import json

import numpy as np
import pytest

from oracle import agym_oracle as orc
from tests import golden_replay as gr


class OracleBackend:
    def start(self, meta):
        self.meta = meta
        self.K, self.S, self.f = meta["frame_stack"], tuple(meta["obs_size"]), tuple(meta["fov_size"])
        self.ring, self.head = orc.new_state(1, self.K, self.S)
        self.loc = np.zeros((1, 2), np.int32)
        self.res = np.array([self.f], np.int32)
        self.call = -1

    def use_f32(self):
        # DMCEnv hands float32 stacks to the wrappers once the float64 zero frames are evicted
        # (dmc_env.py:183,195); torchvision then resamples in float32.
        d = self.meta.get("obs_dtypes")
        return bool(d) and d[self.call] == "float32"

    def ingest(self, meta, fa, fb, flags):
        self.call += 1
        fl = np.array([flags], np.uint8)
        if meta["kind"] == "atari":
            orc.ingest_atari(fa[None], fb[None], fl, self.ring, self.head)
        else:
            orc.ingest_dmc(fa[None], fl, self.ring, self.head)

    def reset_fov(self):
        self.loc[:] = np.rint(np.array(self.meta["fov_init_loc"])).astype(np.int32)
        self.res[:] = self.f

    def observe(self, action, atype):
        m = self.meta
        if action is not None:
            orc.update_loc(action, self.loc, obs_size=self.S, fov_size=self.f, relative=(m["mode"] == "relative"),
                           lo=m["lo"], hi=m["hi"], atype=np.array([atype]) if m["flexible"] else None,
                           res=self.res if m["flexible"] else None)
        if m["flexible"]:
            out = orc.observe_flexible(self.ring, self.head, self.loc, self.res, self.f, variant=m["variant"])
        elif m["peripheral_res"]:
            out = orc.observe_peripheral(self.ring, self.head, self.loc, self.f, m["peripheral_res"],
                                         use_f32=self.use_f32())
        else:
            out = orc.observe_fixed(self.ring, self.head, self.loc, self.f, variant=m["variant"])
        return out[0], self.loc[0].copy(), self.res[0].copy()


@pytest.mark.parametrize("name", gr.SCENARIOS)
def test_oracle_replays_reference(name):
    n = 0
    for r in gr.replay(name, OracleBackend()):
        assert np.array_equal(r["loc_got"], r["loc_want"]), (name, r["call"], r["loc_got"], r["loc_want"])
        if r["flexible"]:
            assert np.array_equal(r["res_got"], r["res_want"]), (name, r["call"])
        assert r["got"].shape == r["want"].shape, (name, r["call"], r["got"].shape, r["want"].shape)
        if r["exact"]:
            assert r["got"].dtype == np.uint8
            assert np.array_equal(r["got"], r["want"]), (name, r["call"])
        else:
            err = np.abs(r["got"].astype(np.float64) - r["want"].astype(np.float64)).max()
            assert err <= 1e-4, (name, r["call"], err)
        n += 1
    assert n >= 4


def test_events_fixture_covers_early_done_and_soft_reset():
    z, meta = gr.load("atari_fixed_rel_crop_events")
    calls = list(zip(z["flags"].tolist(), z["atype"].tolist()))
    assert (1, -1) in calls, "a soft reset (life loss: frame A only, ring not cleared, atari_env.py:86-88)"
    assert (1, 0) in calls, "a step whose game-over came after the t==2 frame (atari_env.py:125-131)"
    assert (0, 0) in calls, "a step whose game-over came before any frame was read"
    assert calls.count((5, -1)) >= 3, "hard resets"


def test_primitives_cv2_resize_and_luma_bit_exact():
    z = np.load(gr.GOLD + "/primitives.npz")
    for i, s in enumerate(z["resize_src"]):
        assert np.array_equal(orc.cv2_resize_linear(s, (84, 84)), z["resize_84"][i]), i
    for i in range(2):
        assert np.array_equal(orc.cv2_resize_linear(z["resize_src"][i], (64, 96)), z["resize_64x96"][i]), i
    assert np.array_equal(orc.luma(z["luma_src"], orc.LUMA_DMC), z["luma_bgr2gray"])
    assert np.array_equal(orc.luma(z["luma_src"], orc.LUMA_RGB), z["luma_rgb2gray"])


def test_primitives_torchvision_resize():
    z = np.load(gr.GOLD + "/primitives.npz")
    cases = json.loads(str(z["aa_cases"]))
    for i, (ish, osh) in enumerate(cases):
        x = z[f"aa{i}_in"].astype(np.float64)
        e64 = np.abs(orc.aa_resize(x, osh) - z[f"aa{i}_out64"]).max()
        e32 = np.abs(orc.aa_resize(x, osh, use_f32=True) - z[f"aa{i}_out32"]).max()
        assert e64 <= 1e-10, (ish, osh, e64)
        assert e32 <= 1e-4, (ish, osh, e32)

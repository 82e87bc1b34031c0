"""CPU: oracle/ref_port.py (the per-env CPU path timed as the baseline) reproduces the golden
vectors of the unmodified reference — same library calls, so the values must be identical."""
import numpy as np
import pytest

from oracle.ref_port import RefPortEnv
from tests import golden_replay as gr


class PortBackend:
    def start(self, meta):
        wrapper = "flexible" if meta["flexible"] else ("peripheral" if meta["peripheral_res"] else "fixed")
        self.env = RefPortEnv(kind=meta["kind"], wrapper=wrapper, frame_stack=meta["frame_stack"],
                              obs_size=tuple(meta["obs_size"]), fov_size=tuple(meta["fov_size"]),
                              fov_init_loc=meta["fov_init_loc"], mode=meta["mode"], lo=meta["lo"], hi=meta["hi"],
                              variant=meta["variant"], peripheral_res=meta["peripheral_res"])
        self.full = None

    def ingest(self, meta, fa, fb, flags):
        self.flags, self.fa, self.fb = flags, fa, fb

    def reset_fov(self):
        self.full = self.env.push_reset(self.fa, bool(self.flags & 4))
        self.env.reset_fov()

    def observe(self, action, atype):
        if action is not None:
            self.full = self.env.push_step(self.fa, self.fb, self.flags)
            a = np.asarray(action)
            if atype == 1:
                a = a.astype(np.int64)  # the reference needs integer fov_res (it slices with it)
            self.env.move(a, atype)
        out = self.env.view(self.full)
        return np.asarray(out, np.float64), np.asarray(self.env.loc), np.asarray(self.env.res)


@pytest.mark.parametrize("name", gr.SCENARIOS)
def test_ref_port_matches_reference(name):
    for r in gr.replay(name, PortBackend()):
        assert np.array_equal(r["loc_got"], r["loc_want"]), (name, r["call"])
        if r["exact"]:  # the reference value of a u8 pixel is float64(float32(u)/255), exactly
            ref = (r["want"].astype(np.float32) / np.float32(255.0)).astype(np.float64)
            assert np.array_equal(r["got"], ref), (name, r["call"])
        else:  # the fixture stores float32(255*ref): compare at float32 resolution
            assert np.abs(r["got"] * 255.0 - r["want"].astype(np.float64)).max() <= 3e-5, (name, r["call"])

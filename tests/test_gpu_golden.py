"""GPU: the CUDA path, called through the C ABI, replays the golden scenarios recorded from the
unmodified reference (tests/golden).  Bit-exact for crop / stack / max-pool / mask / cv2 resize /
luma; |u8 - 255*ref| <= 0.5 + 1e-2 for the torchvision resamples (north_star allows +-1 LSB)."""
import numpy as np
import pytest
import torch

from tests import golden_replay as gr

pytestmark = pytest.mark.gpu

RESAMPLE_TOL = 0.5 + 1e-2  # half an LSB of rounding + evaluation error (fp32, or the 2^-8 grid of the biased lerp)  # u8 LSB: half an LSB of rounding + fp32 evaluation noise


class CudaBackend:
    def __init__(self, use_cache=True):
        self.use_cache = use_cache

    def start(self, meta):
        from active_gym_b200 import ObservationPath, LUMA_DMC, LUMA_RGB
        self.meta = meta
        raw = (210, 160, 1) if meta["kind"] == "atari" else (84, 84, 3)
        self.path = ObservationPath(
            1, meta["frame_stack"], tuple(meta["obs_size"]), raw, luma=LUMA_RGB if meta["kind"] == "atari" else LUMA_DMC,
            fov_size=tuple(meta["fov_size"]), fov_init_loc=meta["fov_init_loc"], sensory_action_mode=meta["mode"],
            sensory_action_space=(meta["lo"], meta["hi"]),
            peripheral_res=tuple(meta["peripheral_res"]) if meta["peripheral_res"] else None,
            cache_peripheral=self.use_cache)
        self.pending_reset = False

    def ingest(self, meta, fa, fb, flags):
        fl = np.array([flags], np.uint8)
        if meta["kind"] == "atari":
            self.path.ingest_atari(fa[None], fb[None], fl)
        else:
            self.path.ingest_dmc(fa[None], fl)

    def reset_fov(self):
        self.pending_reset = True

    def observe(self, action, atype):
        m, p = self.meta, self.path
        ctrl = "reset" if self.pending_reset else None
        self.pending_reset = False
        act = None if action is None else np.asarray(action, np.float64)[None]
        if m["flexible"]:
            out = p.observe_flexible(act, None if action is None else np.array([atype]), variant=m["variant"], ctrl=ctrl)
        elif m["peripheral_res"]:
            out = p.observe_peripheral(act, ctrl=ctrl, use_cache=self.use_cache)
        else:
            out = p.observe_fixed(act, variant=m["variant"], ctrl=ctrl)
        torch.cuda.synchronize()
        return out[0].cpu().numpy(), p.loc[0].cpu().numpy(), p.res[0].cpu().numpy()


@pytest.mark.parametrize("use_cache", [True, False])
@pytest.mark.parametrize("name", gr.SCENARIOS)
def test_cuda_replays_reference(name, use_cache):
    z, meta = gr.load(name)
    if not use_cache and not meta["peripheral_res"]:
        pytest.skip("cache only matters for the peripheral env")
    for r in gr.replay(name, CudaBackend(use_cache)):
        assert np.array_equal(r["loc_got"], r["loc_want"]), (name, r["call"], r["loc_got"], r["loc_want"])
        if r["flexible"]:
            assert np.array_equal(r["res_got"], r["res_want"]), (name, r["call"])
        assert r["got"].shape == r["want"].shape
        assert r["got"].dtype == np.uint8
        if r["exact"]:
            assert np.array_equal(r["got"], r["want"]), (name, r["call"], int((r["got"] != r["want"]).sum()))
        else:
            err = np.abs(r["got"].astype(np.float64) - r["want"].astype(np.float64)).max()
            assert err <= RESAMPLE_TOL, (name, r["call"], err)

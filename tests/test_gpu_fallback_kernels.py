"""The geometry-generic fallback kernels against the oracle at the STANDARD geometry.

The launchers pick a specialised kernel whenever the geometry allows (k_ingest_gray_std / k_ingest_atari_tma,
k_observe_peripheral_std, k_observe_flexible_v3, k_observe_fixed_crop_v3); the table-driven and generic kernels behind
them only run for other geometries, paddings or window sizes.  The library reads a few `AGYM_*` switches at load time
that force those fallbacks (csrc/agym_device.cuh); they are process-wide, so the oracle parity tests of
tests/test_gpu_parity.py are re-run here in a child process per switch set.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SWITCH_SETS = [
    # every specialised kernel off: table-driven peripheral, non-persistent ingest, one-CTA-per-env flexible, byte-gather crop
    {"AGYM_NO_STD": "1", "AGYM_NO_TMA": "1", "AGYM_FLEX_OLD": "1", "AGYM_CROP_OLD": "1"},
    # the previous crop kernel, per-thread fovea staging in the peripheral kernel, plain (non-programmatic) launches
    {"AGYM_CROP_V2": "1", "AGYM_STD_NOFS": "1", "AGYM_NO_PDL": "1", "AGYM_NO_TM": "1"},
]


@pytest.mark.gpu
@pytest.mark.parametrize("switches", SWITCH_SETS, ids=lambda s: "+".join(k[5:].lower() for k in s))
def test_fallback_kernels_match_the_oracle(switches):
    env = dict(os.environ, **switches)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]

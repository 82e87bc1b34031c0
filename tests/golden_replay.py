"""Replays a tests/golden/*.npz scenario (recorded from the unmodified reference) through a
backend — the CPU oracle or the CUDA path — and returns what the backend produced next to
what the reference produced.  Shared by the CPU and the GPU parity tests."""
from __future__ import annotations

import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SCENARIOS = sorted(f[:-4] for f in os.listdir(GOLD)
                   if f.endswith(".npz") and not f.startswith(("screens_", "primitives", "record_")))

_screens = {}


def screens(kind):
    if kind not in _screens:
        _screens[kind] = np.load(os.path.join(GOLD, f"screens_{kind}.npz"))["screens"]
    return _screens[kind]


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def frames_for(meta, idx):
    s = screens(meta["kind"])
    if idx < 0:
        return np.zeros_like(s[0])
    return s[idx]


def replay(name, backend):
    """backend: object with
         start(meta)                       -> allocate N=1 state
         ingest(meta, fa, fb, flags)       -> push one frame
         reset_fov()                       -> loc = rint(init), res = fov_size
         observe(action(2,), atype)        -> (obs ndarray (K,h,w) in u8 units, loc(2,), res(2,))
    Yields dict(call=i, got=obs, want=obs, exact=bool, loc_got, loc_want, res_got, res_want)."""
    z, meta = load(name)
    backend.start(meta)
    want_all = z["obs_u8"] if meta["exact"] else z["obs_f32"]
    for i in range(len(z["flags"])):
        fa, fb = frames_for(meta, int(z["ia"][i])), frames_for(meta, int(z["ib"][i]))
        backend.ingest(meta, fa, fb, int(z["flags"][i]))
        atype = int(z["atype"][i])
        if atype < 0:
            backend.reset_fov()
            got, loc, res = backend.observe(None, 0)
        else:
            got, loc, res = backend.observe(z["action"][i], atype)
        want = want_all[i]
        if meta.get("ragged"):
            rh, rw = (int(v) for v in z["res"][i])
            want = want[:, :rh, :rw]
            got = got[:, :rh, :rw]
        yield dict(call=i, got=np.asarray(got), want=want, exact=meta["exact"], loc_got=np.asarray(loc),
                   loc_want=z["loc"][i], res_got=np.asarray(res), res_want=z["res"][i], flexible=meta["flexible"],
                   meta=meta)

"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck): N chosen so that the
persistent kernels iterate (N > resident CTAs is impossible at sanitizer speed; 2 envs per CTA is forced
with few envs by... the grid being min(N, SMs*occ), so iteration needs N > 296: use 600 for the std kernel)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from active_gym_b200 import ObservationPath, LUMA_RGB, LUMA_DMC
rng = np.random.default_rng(0)
S = (84, 84)
def frames(n, c=1):
    shape = (n, 210, 160) if c == 1 else (n, 210, 160, 3)
    return rng.integers(0, 256, shape, dtype=np.uint8)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
p = ObservationPath(n, 4, S, (210, 160, 1), fov_size=(30, 30), peripheral_res=(20, 20), sensory_action_mode="relative",
                    sensory_action_space=(-10.0, 10.0))
for step in range(3):
    fl = np.full(n, 5 if step == 0 else 3, np.uint8)
    if step == 2: fl[::7] = 8; fl[1::7] = 1; fl[2::7] = 0
    p.ingest_atari(frames(n), frames(n), fl)
    p.observe_peripheral(rng.integers(-10, 11, (n, 2)).astype(np.float64), ctrl="reset" if step == 0 else None)
p.observe_fixed(rng.integers(-10, 11, (n, 2)).astype(np.float64))
p.observe_fixed(None, variant="mask", ctrl=np.full(n, 2, np.uint8))
p.observe_fixed(None, variant="resize_full", ctrl=np.full(n, 2, np.uint8))
at = rng.integers(0, 2, n).astype(np.int32)
a = np.where(at[:, None] == 1, rng.integers(1, 85, (n, 2)), rng.integers(-10, 11, (n, 2))).astype(np.float64)
for v in ("mask", "crop", "resize_full"):
    p.observe_flexible(a, at, variant=v)
p.stack()
m = 64
q = ObservationPath(m, 4, S, (210, 160, 3), luma=LUMA_RGB, fov_size=(30, 30))
q.ingest_atari(frames(m, 3), frames(m, 3), np.full(m, 5, np.uint8))
d = ObservationPath(m, 3, S, (84, 84, 3), luma=LUMA_DMC, fov_size=(30, 30), peripheral_res=(20, 20))
d.ingest_dmc(rng.integers(0, 256, (m, 84, 84, 3), dtype=np.uint8), np.full(m, 5, np.uint8))
d.observe_peripheral(None, ctrl="reset")
torch.cuda.synchronize()
print("sanitize smoke done")

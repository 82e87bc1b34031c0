"""Where do the e2e milliseconds go?  The host pipeline's copies alone: packed (strided) H2D of both frames in
16 shards, with and without the concurrent D2H of the observations, no kernels."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from active_gym_b200.hostpipe import _rt
rt = _rt()
n, rb, P, o, L, raw_h = 16384, 160, 5, 3, 4, 210
shards = 16
h = [torch.empty(n * raw_h * rb, dtype=torch.uint8).pin_memory() for _ in range(2)]
d = [torch.empty(n * 168 * rb, dtype=torch.uint8, device="cuda") for _ in range(2)]
ho = torch.empty(n * 28224, dtype=torch.uint8).pin_memory()
do = torch.empty(n * 28224, dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(shards)]
per = n // shards
def h2d(i, f, packed=True):
    lo, hi = i * per, (i + 1) * per
    s = torch.cuda.current_stream().cuda_stream
    if not packed:
        d[f][lo * 168 * rb:(lo * 168 + per * 210 * 0 + per * 168) * rb]  # no-op
        rt.cudaMemcpyAsync(d[f].data_ptr() + lo * 168 * rb, h[f].data_ptr() + lo * raw_h * rb, per * 168 * rb, 1, s)
        return
    periods = per * (raw_h // P)
    s0 = h[f].data_ptr() + lo * raw_h * rb; d0 = d[f].data_ptr() + lo * 168 * rb
    rt.cudaMemcpyAsync(d0, s0, 2 * rb, 1, s)
    rt.cudaMemcpy2DAsync(d0 + 2 * rb, L * rb, s0 + o * rb, P * rb, L * rb, periods - 1, 1, s)
    rt.cudaMemcpyAsync(d0 + 2 * rb + (periods - 1) * L * rb, s0 + ((periods - 1) * P + o) * rb, 2 * rb, 1, s)
def step(with_d2h, packed=True):
    for i in range(shards):
        with torch.cuda.stream(streams[i]):
            h2d(i, 0, packed); h2d(i, 1, packed)
            if with_d2h:
                ho[i * per * 28224:(i + 1) * per * 28224].copy_(do[i * per * 28224:(i + 1) * per * 28224], non_blocking=True)
    for s in streams: s.synchronize()
for name, args in (("packed H2D only", (False, True)), ("packed H2D + D2H", (True, True)), ("contiguous same bytes H2D + D2H", (True, False)), ("contiguous H2D only", (False, False))):
    step(*args); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): step(*args)
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt*1e3:.2f} ms/step")

#!/usr/bin/env python
"""Upper bound of what overlapping the observe launch with the ingest launch could buy a two-kernel step (DMC, fixed
fovea): the two kernels of a step on TWO streams with no dependency between them (NOT a valid step — the observe kernel
reads the ring the ingest kernel is writing — only a timing probe), captured in a CUDA graph, against the serial step.

    python tools/overlap_probe.py [workload] [steps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

wname = sys.argv[1] if len(sys.argv) > 1 else "dmc_fixed"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 48
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = bench.Workload(wname, dev)


def graph_time(body):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            wl.step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        body()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    return best


def serial():
    for _ in range(steps):
        wl.step()


def forked():
    cur = torch.cuda.current_stream(dev)
    other = torch.cuda.Stream(device=dev)
    for _ in range(steps):
        other.wait_stream(cur)          # step t's two kernels start together, after step t-1 has finished
        with torch.cuda.stream(other):
            wl.observe()
        wl.ingest(); wl.t += 1
        cur.wait_stream(other)


print(json.dumps({"workload": wname, "envs": wl.n, "serial_ms_per_step": graph_time(serial), "overlapped_ms_per_step": graph_time(forked),
                  "note": "overlapped = ingest and observe of a step on two streams without a dependency: an upper bound, not a valid step"}))

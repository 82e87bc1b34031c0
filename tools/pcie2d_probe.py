"""Does a strided (row-skipping) pinned H2D copy keep PCIe bandwidth?  width 640 B of every 800 B."""
import ctypes, time, torch
rt = ctypes.CDLL("libcudart.so.12") if True else None
n_rows5 = 16384 * 42  # 5-row groups in one frame batch (16384 envs x 210 rows)
h = torch.empty(n_rows5 * 800, dtype=torch.uint8).pin_memory()
d = torch.empty(n_rows5 * 640, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
def f2d():
    r = rt.cudaMemcpy2DAsync(d.data_ptr(), 640, h.data_ptr() + 480, 800, 640, n_rows5 - 1, 1, s)
    assert r == 0, r
def f1d():
    d.copy_(h[: d.numel()], non_blocking=True)
for name, fn, nbytes in (("2D 640/800", f2d, (n_rows5 - 1) * 640), ("1D same bytes", f1d, d.numel())):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {nbytes/dt/1e9:.1f} GB/s  ({dt*1e3:.2f} ms)")

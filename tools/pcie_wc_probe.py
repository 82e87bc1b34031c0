"""PCIe probe: H2D bandwidth from default pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined),
contiguous and with the host pipeline's strided row copies, alone and with the observation D2H running."""
import ctypes as C, time, torch
rt = C.CDLL("libcudart.so.12")
vp, sz = C.c_void_p, C.c_size_t
rt.cudaHostAlloc.argtypes = [C.POINTER(vp), sz, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [vp, vp, sz, C.c_int, vp]
rt.cudaMemcpy2DAsync.argtypes = [vp, sz, vp, sz, sz, sz, C.c_int, vp]
torch.cuda.init(); torch.zeros(1, device="cuda")
N = 16384
full = N * 210 * 160 * 2
d = torch.empty(N * 168 * 160 * 2, dtype=torch.uint8, device="cuda")
ho = torch.empty(N * 4 * 84 * 84, dtype=torch.uint8).pin_memory(); do = torch.empty_like(ho, device="cuda")
s2 = torch.cuda.Stream()
def alloc(flags):
    p = vp()
    assert rt.cudaHostAlloc(C.byref(p), full, flags) == 0
    C.memset(p, 7, full)
    return p
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
for name, flags in (("default", 0), ("write-combined", 4)):
    h = alloc(flags)
    st = torch.cuda.current_stream().cuda_stream
    contig = t(lambda: rt.cudaMemcpyAsync(d.data_ptr(), h, d.numel(), 1, st))
    # strided: 4 of every 5 rows (800-byte period, 640 bytes copied), as hostpipe does
    rows = full // 800
    strided = t(lambda: rt.cudaMemcpy2DAsync(d.data_ptr(), 640, h.value + 480, 800, 640, rows - 1, 1, st))
    def both():
        rt.cudaMemcpy2DAsync(d.data_ptr(), 640, h.value + 480, 800, 640, rows - 1, 1, st)
        with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    b = t(both)
    print(f"{name}: contiguous {d.numel()/contig/1e9:.1f} GB/s, strided {640*(rows-1)/strided/1e9:.1f} GB/s ({strided*1e3:.1f} ms), strided + D2H {b*1e3:.1f} ms")

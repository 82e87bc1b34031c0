"""PCIe probe on the GPU box: pinned H2D / D2H bandwidth alone and together (bounds the e2e number)."""
import time, torch
dev = torch.device("cuda:0")
h = torch.empty(1101004800, dtype=torch.uint8).pin_memory()
d = torch.empty_like(h, device=dev)
ho = torch.empty(462422016, dtype=torch.uint8).pin_memory()
do = torch.empty_like(ho, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
a = t(lambda: d.copy_(h, non_blocking=True))
b = t(lambda: ho.copy_(do, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
c = t(both)
print(f"H2D {h.numel()/a/1e9:.1f} GB/s ({a*1e3:.1f} ms)  D2H {ho.numel()/b/1e9:.1f} GB/s ({b*1e3:.1f} ms)  both {c*1e3:.1f} ms -> H2D {h.numel()/c/1e9:.1f} GB/s")
# chunked H2D in 4 / 16 pieces on one stream
for k in (4, 16):
    n = h.numel() // k
    def ch():
        for i in range(k): d[i*n:(i+1)*n].copy_(h[i*n:(i+1)*n], non_blocking=True)
    x = t(ch)
    print(f"H2D in {k} chunks: {h.numel()/x/1e9:.1f} GB/s")

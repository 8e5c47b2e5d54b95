import os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200")]
import numpy as np, torch

class Mapped:
    def __init__(self, t):
        self.t = t
        self.__cuda_array_interface__ = {"shape": tuple(t.shape), "typestr": "<f8", "data": (t.data_ptr(), False), "version": 2}

n = 3920032 // 8
h = torch.randn(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
hm = torch.as_tensor(Mapped(h), device="cuda")
print("mapped view:", hm.device, hm.shape, hm.data_ptr() == h.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def graph_of(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    return g

def timeit(fn, reps=20):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))

g_dma = graph_of(lambda: d.copy_(h, non_blocking=True))
g_ker = graph_of(lambda: d.copy_(hm))
print("H2D 3.92 MB  DMA eager      ", timeit(lambda: d.copy_(h, non_blocking=True)))
print("H2D 3.92 MB  DMA in graph   ", timeit(g_dma.replay))
print("H2D 3.92 MB  kernel eager   ", timeit(lambda: d.copy_(hm)))
print("H2D 3.92 MB  kernel in graph", timeit(g_ker.replay))
assert torch.equal(d.cpu(), h)
# D2H 0.72 MB
m = 720208 // 8
ho = torch.empty(m, dtype=torch.float64).pin_memory(); do = torch.randn(m, dtype=torch.float64, device="cuda")
hom = torch.as_tensor(Mapped(ho), device="cuda")
g_dma2 = graph_of(lambda: ho.copy_(do, non_blocking=True))
g_ker2 = graph_of(lambda: hom.copy_(do))
print("D2H 0.72 MB  DMA in graph   ", timeit(g_dma2.replay))
print("D2H 0.72 MB  kernel in graph", timeit(g_ker2.replay))
torch.cuda.synchronize(); assert torch.equal(ho, do.cpu())

"""Where does the end-to-end time of one host-memory pass go?  Times, with CUDA events and the L2 flushed
before every sample: the H2D copy alone, the resident pass alone, the D2H copy alone, and the whole graph.
usage: python profiles/e2e_probe.py [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200")]
import numpy as np
import torch
from ipoc_b200 import workloads
from ipoc_b200.runner import HostNewtonPass

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000
dev = torch.device("cuda", 0)
w = workloads.newton_inputs("cartpole", N, dev, seed=1, x0_noise=0.0)
hp = HostNewtonPass(w, dev)
hp.capture()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_of(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


g_h2d = graph_of(lambda: hp.d_in.copy_(hp.h_in, non_blocking=True))
g_d2h = graph_of(lambda: hp.h_out.copy_(hp.d_out, non_blocking=True))
g_pass = graph_of(hp.inner.run)
print(f"N={N} h2d {hp.h2d_bytes/1e6:.2f} MB, d2h {hp.d2h_bytes/1e6:.2f} MB   (median, min) ms")
print("  H2D graph      ", timeit(g_h2d.replay))
print("  H2D eager      ", timeit(lambda: hp.d_in.copy_(hp.h_in, non_blocking=True)))
print("  pass graph     ", timeit(g_pass.replay))
print("  D2H graph      ", timeit(g_d2h.replay))
print("  whole graph    ", timeit(hp.replay))
print("  three graphs   ", timeit(lambda: (g_h2d.replay(), g_pass.replay(), g_d2h.replay())))
print("  eager copies + pass graph", timeit(lambda: (hp.d_in.copy_(hp.h_in, non_blocking=True), g_pass.replay(),
                                                     hp.h_out.copy_(hp.d_out, non_blocking=True))))

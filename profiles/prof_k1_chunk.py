"""K1 alone (ipoc_costates_f64: costate scan + ||cu||) against the leaf chunk length and the level organisation:
graph replay with L2 flushed.  usage: python profiles/prof_k1_chunk.py N [N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from ipoc_b200 import _lib, noc

CHUNKS = [int(c) for c in os.environ.get("CHUNKS", "0,2,4,8,10,16,24,32,64").split(",")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for N in [int(float(a)) for a in sys.argv[1:]] or [10000]:
    g = torch.Generator(device="cuda").manual_seed(0)
    fx = torch.eye(4, dtype=torch.float64, device="cuda") + torch.randn(1, N, 4, 4, dtype=torch.float64, device="cuda", generator=g) / N
    cx = torch.randn(1, N, 4, dtype=torch.float64, device="cuda", generator=g)
    cu = torch.randn(1, N, 1, dtype=torch.float64, device="cuda", generator=g)
    lamT = torch.randn(1, 4, dtype=torch.float64, device="cuda", generator=g)
    for hier in (1, 0):
        for chunk in CHUNKS:
            _lib.lib().ipoc_set_hier(hier, 0, 0)
            _lib.lib().ipoc_set_tuning(chunk, 0, 0)
            out = noc.costates_fused(fx, cx, lamT, cu)
            for _ in range(3):
                noc.costates_fused(fx, cx, lamT, cu, out=out)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                noc.costates_fused(fx, cx, lamT, cu, out=out)
            ts = []
            for _ in range(20):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            print(f"N={N} hier={hier} leaf_chunk={chunk}: K1 {np.median(ts) * 1e3:.1f} us (min {np.min(ts) * 1e3:.1f})", flush=True)
_lib.lib().ipoc_set_hier(1, 0, 0)
_lib.lib().ipoc_set_tuning(0, 0, 0)

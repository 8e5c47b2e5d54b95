"""Per-kernel CUDA-event times of one hot-path pass for the scan organisations of the library:
fused (5 launches), in-kernel levels as separate calls, separate level kernels (round-1 sequence).
usage: python profiles/prof_variants.py N [N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from helpers import random_lq
from ipoc_b200 import _lib
from ipoc_b200.runner import NewtonPass

T = lambda a: torch.as_tensor(a, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for N in [int(float(a)) for a in sys.argv[1:]] or [10000]:
    rng = np.random.default_rng(0)
    fx, fu, ru, Q, R, M = random_lq(rng, N, 4, 1, dt=1.0 / N)
    shape = fx.shape[:-2]
    args = (T(fx), T(fu), T(rng.standard_normal(shape + (4,))), T(rng.standard_normal(shape + (1,))),
            T(rng.standard_normal((1, 4))), T(ru), T(Q), T(R), T(M), T(-np.ones(shape + (2,))))
    for label, hier, fused in (("fused", 1, True), ("fused, Riccati levels in-kernel", 2, True),
                               ("fused, no cooperative level kernel (r02 mid-round)", 3, True),
                               ("hier, separate calls", 1, False), ("level kernels", 0, False)):
        _lib.lib().ipoc_set_hier(hier, 0, 0)
        p = NewtonPass(*args)
        p.fused = fused
        for _ in range(3):
            p.run()
        torch.cuda.synchronize()
        acc = {}
        order = []
        reps = 5
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            for name, t in p.profile():
                if name not in acc:
                    order.append(name)
                acc[name] = acc.get(name, 0.0) + t / reps
        p.capture()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); p.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"N={N} [{label}] graph replay {np.mean(ts) * 1e3:.1f} us; eager per-kernel (us): " +
              ", ".join(f"{k} {acc[k] * 1e3:.1f}" for k in order), flush=True)
    _lib.lib().ipoc_set_hier(1, 0, 0)

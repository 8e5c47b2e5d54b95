"""K2 + K3 (ipoc_newton_attempt_f64, trial point as a side job) at the batched shape of BASELINE config 5 and at one long
horizon: event time of the call and of its kernels (library profiler).  usage: prof_attempt_batched.py [B N] ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from helpers import random_lq
from ipoc_b200 import _lib, noc

T = lambda a: torch.as_tensor(a, device="cuda")
shapes = [(int(sys.argv[i]), int(float(sys.argv[i + 1]))) for i in range(1, len(sys.argv) - 1, 2)] or [(8192, 1000), (1, 1000000)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, N in shapes:
    rng = np.random.default_rng(0)
    fx, fu, ru, Q, R, M = (T(a) for a in random_lq(rng, N, 4, 1, batch=B, dt=1.0 / N))
    x, u = torch.randn(B, N + 1, 4, dtype=torch.float64, device="cuda"), torch.randn(B, N, 1, dtype=torch.float64, device="cuda")
    tx, tu = torch.empty_like(x), torch.empty_like(u)
    rp, cn = torch.ones(B, dtype=torch.float64, device="cuda"), torch.full((B,), 0.1, dtype=torch.float64, device="cuda")
    act = torch.ones(B, dtype=torch.int32, device="cuda")
    buf = noc.AttemptBuffers(B, N, 4, 1, "cuda")
    for trial in (True, False):
        args = (x, u, tx, tu, act) if trial else ()
        for _ in range(3):
            noc.newton_attempt(buf, fx, fu, ru, Q, R, M, rp, cn, *args)
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); noc.newton_attempt(buf, fx, fu, ru, Q, R, M, rp, cn, *args); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"B={B} N={N} trial_point={trial}: K2+K3 {np.median(ts) * 1e3:.1f} us (min {np.min(ts) * 1e3:.1f})", flush=True)

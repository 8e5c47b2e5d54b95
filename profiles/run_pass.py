"""Tiny driver for ncu captures: runs a few hot-path passes on random well-conditioned LQ data
(no torch.func set-up kernels in the way).  usage: python profiles/run_pass.py N [nx] [batch] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from helpers import random_lq
from ipoc_b200.runner import NewtonPass

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 4
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rng = np.random.default_rng(0)
fx, fu, ru, Q, R, M = random_lq(rng, N, nx, 1, batch=batch or None, dt=1.0 / N)
shape = fx.shape[:-2]
T = lambda a: torch.as_tensor(a, device="cuda")
p = NewtonPass(T(fx), T(fu), T(rng.standard_normal(shape + (nx,))), T(rng.standard_normal(shape + (1,))),
               T(rng.standard_normal((max(batch, 1), nx))), T(ru), T(Q), T(R), T(M), T(-np.ones(shape + (2,))))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    flush.zero_()
    p.run()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
flush.zero_()
ev[0].record(); p.run(); ev[1].record(); torch.cuda.synchronize()
print(f"N={N} nx={nx} batch={batch}: one pass {ev[0].elapsed_time(ev[1]):.3f} ms, feasible={int(p.bwd_feas[0])}")

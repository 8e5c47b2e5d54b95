"""ncu driver: a few fused hot-path passes (ipoc_costates_f64 + ipoc_newton_attempt_f64) on random LQ data.
usage: python profiles/run_fused.py N [hier] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from helpers import random_lq
from ipoc_b200 import _lib
from ipoc_b200.runner import NewtonPass

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000
hier = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
_lib.lib().ipoc_set_hier(hier, 0, 0)
rng = np.random.default_rng(0)
fx, fu, ru, Q, R, M = random_lq(rng, N, 4, 1, dt=1.0 / N)
shape = fx.shape[:-2]
T = lambda a: torch.as_tensor(a, device="cuda")
p = NewtonPass(T(fx), T(fu), T(rng.standard_normal(shape + (4,))), T(rng.standard_normal(shape + (1,))),
               T(rng.standard_normal((1, 4))), T(ru), T(Q), T(R), T(M), T(-np.ones(shape + (2,))))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    flush.zero_()
    p.run()
torch.cuda.synchronize()
print("ok", int(p.bwd_feas[0]), float(p.hu[0]))

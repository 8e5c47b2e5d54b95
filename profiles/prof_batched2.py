"""Batched solves/s (BASELINE config 5 shape: N = 1000) for a few batch sizes.
usage: python profiles/prof_batched2.py problem B [B ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200")]
import numpy as np
import torch
from ipoc_b200 import problems, batched

prob = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
dev = torch.device("cuda")
N = 1000
ocp = problems.make_pendulum(1.0 / N) if prob == "pendulum" else problems.make_cartpole(1.0 / N)
x0b = (problems.pendulum_x0 if prob == "pendulum" else problems.cartpole_x0)(device=dev)
for B in [int(a) for a in sys.argv[2:]] or [512]:
    rng = np.random.default_rng(1)
    x0s = x0b[None] + torch.as_tensor(0.1 * rng.standard_normal((B, x0b.numel())), device=dev)
    u0s = torch.as_tensor(0.1 * rng.standard_normal((B, N, 1)), device=dev)
    for rep in range(2):      # first = warm-up (graph captures of the ladder)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ub, itb = batched.par_interior_point_optimal_control_batched(ocp, u0s, x0s)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{prob} N={N} B={B}: {dt:.3f} s = {B / dt:.0f} solves/s; iterations mean {float(itb.double().mean()):.1f} "
          f"max {int(itb.max())}; mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)

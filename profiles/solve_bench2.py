"""Full IP solves (reference timing protocol: 1 warm-up = graph capture, then timed calls) for the cartpole at a
few horizons.  usage: python profiles/solve_bench2.py [N ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200")]
import numpy as np
import torch
from ipoc_b200 import noc, problems

dev = torch.device("cuda")
for N in [int(float(a)) for a in sys.argv[1:]] or [1000, 10000, 100000]:
    ocp = problems.make_cartpole(1.0 / N)
    x0 = problems.cartpole_x0(device=dev)
    u0 = torch.as_tensor(0.1 * np.random.default_rng(1).standard_normal((N, 1)), device=dev)
    noc.par_interior_point_optimal_control(ocp, u0, x0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        u, it = noc.par_interior_point_optimal_control(ocp, u0, x0)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(f"cartpole N={N}: solve {np.median(ts) * 1e3:.1f} ms, {it} Newton iterations, max|u| {float(u.abs().max()):.6f}", flush=True)

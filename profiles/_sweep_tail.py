import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200")]
import numpy as np, torch
from ipoc_b200 import problems, batched
prob, B, N = sys.argv[1], int(sys.argv[2]), 1000
dev = "cuda"
rng = np.random.default_rng(1)
ocp = problems.make_pendulum(1.0 / N) if prob == "pendulum" else problems.make_cartpole(1.0 / N)
x0 = (problems.pendulum_x0 if prob == "pendulum" else problems.cartpole_x0)(device=dev)
x0s = x0[None] + torch.as_tensor(0.1 * rng.standard_normal((B, x0.numel())), device=dev)
u0s = torch.as_tensor(0.1 * rng.standard_normal((B, N, 1)), device=dev)
for K, burst in ((8, 4), (8, 8), (16, 8), (32, 8), (32, 16), (64, 16)):
    batched._TailGraph.K, batched._TailGraph.BURST = K, burst
    batched._tail_cache.clear()
    batched.par_interior_point_optimal_control_batched(ocp, u0s[:64], x0s[:64])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    u, its = batched.par_interior_point_optimal_control_batched(ocp, u0s, x0s)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{prob} B={B} K={K} burst={burst}: {dt:.2f} s  {B/dt:.0f} solves/s  its mean {float(its.double().mean()):.1f}")

import os, sys, time, cProfile, pstats
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200")]
import numpy as np, torch
from ipoc_b200 import problems, batched
dev = torch.device("cuda"); N = 1000; B = 8192
ocp = problems.make_pendulum(1.0 / N); x0b = problems.pendulum_x0(device=dev)
rng = np.random.default_rng(1)
x0s = x0b[None] + torch.as_tensor(0.1 * rng.standard_normal((B, x0b.numel())), device=dev)
u0s = torch.as_tensor(0.1 * rng.standard_normal((B, N, 1)), device=dev)
batched.par_interior_point_optimal_control_batched(ocp, u0s, x0s); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
batched.par_interior_point_optimal_control_batched(ocp, u0s, x0s); torch.cuda.synchronize()
dt = time.perf_counter() - t0
pr.disable()
print("solve", dt)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

"""ncu driver for the kernels added late in round 2: the lane-cooperative Riccati level kernel (fused pass, N = 1e6),
the grid form of the plant cost kernel and the parallel rollout's linearisation kernel (cartpole, N = 1e6).
usage: python profiles/run_new_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from helpers import random_lq
from ipoc_b200 import plants, problems
from ipoc_b200.runner import NewtonPass

N = 1000000
rng = np.random.default_rng(0)
fx, fu, ru, Q, R, M = random_lq(rng, N, 4, 1, dt=1.0 / N)
T = lambda a: torch.as_tensor(a, device="cuda")
p = NewtonPass(T(fx), T(fu), T(rng.standard_normal((N, 4))), T(rng.standard_normal((N, 1))), T(rng.standard_normal((1, 4))),
               T(ru), T(Q), T(R), T(M), T(-np.ones((N, 2))))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ocp = problems.make_cartpole(1.0 / N)
plant = plants.plant_of(ocp)
u = T(0.1 * rng.standard_normal((N, 1)))
x = plants.rollout(plant, u, problems.cartpole_x0(device="cuda"))
for _ in range(2):
    flush.zero_()
    p.run()
    flush.zero_()
    plants.cost(plant, x, u, 0.1)
    flush.zero_()
    plants.rollout_parallel(plant, u, problems.cartpole_x0(device="cuda"), x_guess=x, max_iter=1)
torch.cuda.synchronize()
print("ok")

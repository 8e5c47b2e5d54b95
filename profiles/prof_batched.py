import os, sys, time
sys.path[:0] = ["/root/repo", "/root/repo/ip-parallel-optimal-control_b200"]
import numpy as np, torch
from ipoc_b200 import noc, problems, batched, plants
B, N = 4096, 1000
dev = "cuda"
rng = np.random.default_rng(1)
ocp = problems.make_cartpole(1.0 / N); plant = plants.plant_of(ocp)
x0 = problems.cartpole_x0().to(dev)
x0s = x0[None] + torch.as_tensor(0.1 * rng.standard_normal((B, 4)), device=dev)
u = torch.as_tensor(0.1 * rng.standard_normal((B, N, 1)), device=dev)
def T(fn, name, reps=5):
    fn(); torch.cuda.synchronize(); t = time.time()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize(); print(f"{name:28s} {(time.time()-t)/reps*1e3:8.3f} ms"); return out
x = T(lambda: plants.rollout(plant, u, x0s), "rollout")
d, lamT = T(lambda: plants.derivatives(plant, x, u, 0.1), "derivatives")
cost, _ = T(lambda: plants.cost(plant, x, u, 0.1), "cost")
lam = T(lambda: noc.affine_scan(d.fx, d.cx, lamT, reverse=True, transpose=True), "K1 costates")
ru, Q, R, M = T(lambda: noc.compute_lqr_params(lam, d), "lqr params")
hu, cn, _ = T(lambda: noc.reductions(ru=ru, cu=d.cu), "reductions")
rp = torch.ones(B, dtype=torch.float64, device=dev)
dx, du, _, _, pred, bf = T(lambda: noc.newton_step(d.fx, d.fu, ru, Q, R, M, rp * cn), "K2+K3 newton_step")
T(lambda: (x + dx, u + du), "axpy")
act = torch.ones(B, dtype=torch.bool, device=dev)
T(lambda: torch.where(act.view(B,1,1), x+dx, x), "where")
T(lambda: bool(act.any()), "any() sync")

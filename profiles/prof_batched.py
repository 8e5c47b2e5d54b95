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
d, lamT = T(lambda: plants.derivatives(plant, x, u, 0.1), "derivatives (full record)")
T(lambda: plants.linearize(plant, x, u, 0.1), "linearize")
cost, _ = T(lambda: plants.cost(plant, x, u, 0.1), "cost")
lam = T(lambda: noc.affine_scan(d.fx, d.cx, lamT, reverse=True, transpose=True), "K1 costates")
ru, Q, R, M = T(lambda: noc.compute_lqr_params(lam, d), "lqr params (A3 kernel)")
T(lambda: plants.hamiltonian(plant, x, u, lam, 0.1), "hamiltonian")
T(lambda: noc.eval_iteration(ocp, x, u, 0.1), "eval_iteration total")
T(lambda: plants.cost(plant, x + 0.0, u + 0.0, 0.1), "cost (trial)")
cost_t = torch.ones(B, dtype=torch.float64, device=dev)
hu, cn, _ = T(lambda: noc.reductions(ru=ru, cu=d.cu), "reductions")
rp = torch.ones(B, dtype=torch.float64, device=dev)
dx, du, _, _, pred, bf = T(lambda: noc.newton_step(d.fx, d.fu, ru, Q, R, M, rp * cn), "K2+K3 newton_step")
T(lambda: (x + dx, u + du), "axpy")
act = torch.ones(B, dtype=torch.bool, device=dev)
T(lambda: torch.where(act.view(B,1,1), x+dx, x), "where")
T(lambda: bool(act.any()), "any() sync")
ri = torch.full((B,), 2.0, dtype=torch.float64, device=dev)
tf = torch.ones(B, dtype=torch.int32, device=dev)
T(lambda: noc.accept_update(cost_t, cost_t * 0.9, tf, pred, bf, rp, ri, active=act.to(torch.int32)), "accept_update")
import ipoc_b200.batched as bt
t = time.time(); xx, uu, its = bt.newton_oc_batched(ocp, u, x0s, 0.1); torch.cuda.synchronize(); dt = time.time() - t
print(f"newton_oc_batched stage bp=0.1: {dt:.2f}s, iterations max {int(its.max())} mean {float(its.double().mean()):.1f}")

"""Where does a batched solve spend its attempts?  Wraps the kernels the batched Newton loop calls and
histograms the calls by the number of members they were run on (compaction shrinks the working set).
usage: python profiles/prof_batched_trace.py [cartpole|pendulum] [B] [N]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200")]
import numpy as np, torch
from ipoc_b200 import problems, batched

prob = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
dev = "cuda"
rng = np.random.default_rng(1)
ocp = problems.make_pendulum(1.0 / N) if prob == "pendulum" else problems.make_cartpole(1.0 / N)
x0 = (problems.pendulum_x0 if prob == "pendulum" else problems.cartpole_x0)(device=dev)
x0s = x0[None] + torch.as_tensor(0.1 * rng.standard_normal((B, x0.numel())), device=dev)
u0s = torch.as_tensor(0.1 * rng.standard_normal((B, N, 1)), device=dev)
batched.par_interior_point_optimal_control_batched(ocp, u0s[:64], x0s[:64])
torch.cuda.synchronize()

log = {"newton_step": [], "eval_iteration": []}
def wrap(name, fn, arg_index):
    def f(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize()
        log[name].append((a[arg_index].shape[0], time.perf_counter() - t))
        return out
    return f
batched.newton_step = wrap("newton_step", batched.newton_step, 0)
batched.eval_iteration = wrap("eval_iteration", batched.eval_iteration, 1)
t0 = time.perf_counter()
u, its = batched.par_interior_point_optimal_control_batched(ocp, u0s, x0s)
torch.cuda.synchronize()
total = time.perf_counter() - t0
print(f"{prob} B={B} N={N}: {total:.2f} s (with per-call syncs), iterations mean {float(its.double().mean()):.1f} max {int(its.max())}")
for name, rows in log.items():
    rows = np.array(rows)
    print(f"  {name}: {len(rows)} calls, {rows[:,1].sum():.2f} s inside")
    for lo, hi in ((1, 1), (2, 8), (9, 64), (65, 512), (513, 1 << 30)):
        m = (rows[:, 0] >= lo) & (rows[:, 0] <= hi)
        if m.any():
            print(f"     members {lo:>4}..{min(hi, B):<5}: {int(m.sum()):6d} calls, {rows[m,1].sum():7.2f} s, "
                  f"{rows[m,1].mean()*1e3:7.3f} ms/call")

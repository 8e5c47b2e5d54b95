"""Full-solve timings through the reference-facing API (not only the hot-path pass):
   python profiles/solve_bench.py single cartpole 1000     # one OCP, ref protocol: 1 warm-up + reps
   python profiles/solve_bench.py batched cartpole 1000 1024"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200")]
import numpy as np, torch
from ipoc_b200 import noc, problems, batched

mode, problem, N = sys.argv[1], sys.argv[2], int(float(sys.argv[3]))
use_graphs = os.environ.get("IPOC_EAGER", "0") != "1"
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = "cuda"
rng = np.random.default_rng(1)
ocp = problems.make_cartpole(1.0 / N) if problem == "cartpole" else problems.make_pendulum(1.0 / N)
x0 = (problems.cartpole_x0() if problem == "cartpole" else problems.pendulum_x0()).to(dev)
if mode == "single":
    u0 = torch.as_tensor(0.1 * rng.standard_normal((N, 1)), device=dev)
    trace = []
    t = time.time(); u, it = noc.par_interior_point_optimal_control(ocp, u0, x0, trace=trace, use_graphs=use_graphs); torch.cuda.synchronize()
    print(f"warm-up solve: {time.time()-t:.2f}s iterations={it} attempts={len(trace)}")
    ts = []
    for _ in range(3):
        t = time.time(); u, it = noc.par_interior_point_optimal_control(ocp, u0, x0, use_graphs=use_graphs); torch.cuda.synchronize(); ts.append(time.time() - t)
    print(f"{problem} N={N}: solve mean {np.mean(ts)*1e3:.1f} ms, {it} Newton iterations, {np.mean(ts)*1e3/max(1,len(trace)):.2f} ms per attempt")
else:
    x0s = x0[None] + torch.as_tensor(0.1 * rng.standard_normal((B, x0.numel())), device=dev)
    u0s = torch.as_tensor(0.1 * rng.standard_normal((B, N, 1)), device=dev)
    t = time.time(); u, it = batched.par_interior_point_optimal_control_batched(ocp, u0s, x0s); torch.cuda.synchronize()
    dt = time.time() - t
    print(f"{problem} N={N} batch={B}: batched solve {dt:.2f}s -> {B/dt:.1f} solves/s; iterations min/mean/max {int(it.min())}/{float(it.double().mean()):.1f}/{int(it.max())}; peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")

"""Synthetic workloads of the shapes BASELINE.json names (bench.py, tests): the LQ tensors of one
Newton iteration of the reference's example problems at a given horizon.

Inputs follow BASELINE.md §3: N*Ts = 1 s, x0 from the example scripts, u0 = 0.1*N(0,1) drawn with
numpy.random.default_rng(seed) (JAX's PRNGKey stream cannot be reproduced without JAX)."""
import math
import numpy as np
import torch
from . import problems
from .noc import compute_derivatives, compute_lqr_params, par_costates
from torch.func import grad, vmap


def cartpole_rollout_np(u, x0, Ts):
    """Scalar-float Euler rollout of the cartpole (same arithmetic as problems.make_cartpole);
    the generic torch rollout is a Python loop of tensor ops and too slow for N >= 1e5."""
    g, l, mc, mp = 9.81, 0.5, 10.0, 1.0
    mt = mc + mp
    N = u.shape[0]
    xs = np.empty((N + 1, 4))
    p, th, v, w = (float(a) for a in x0)
    xs[0] = (p, th, v, w)
    uu = u.reshape(-1)
    for k in range(N):
        a = float(uu[k])
        s, c = math.sin(th), math.cos(th)
        acc_c = (a + mp * s * (l * w ** 2 + g * c)) / (mc + mp * s ** 2)
        acc_p = (-a * c - mp * l * w ** 2 * c * s - mt * g * s) / (l * mc + l * mp * s ** 2)
        p, th, v, w = p + Ts * v, th + Ts * w, v + Ts * acc_c, w + Ts * acc_p
        xs[k + 1] = (p, th, v, w)
    return xs


def pendulum_rollout_np(u, x0, Ts):
    g, l, m, damping = 9.81, 1.0, 1.0, 1e-3
    N = u.shape[0]
    xs = np.empty((N + 1, 2))
    th, w = float(x0[0]), float(x0[1])
    xs[0] = (th, w)
    uu = u.reshape(-1)
    for k in range(N):
        acc = -g / l * math.sin(th) + (float(uu[k]) - damping * w) / (m * l ** 2)
        th, w = th + Ts * w, w + Ts * acc
        xs[k + 1] = (th, w)
    return xs


def newton_inputs(problem, N, device, seed=1, bp=0.1, x0_noise=0.0, chunk=200000):
    """LQ tensors of the first Newton iteration: dict(fx, fu, cx, cu, lamT, ru, Q, R, M, x, u, cons,
    cost) on `device`.  Derivatives come from the host framework (torch.func), chunked over time
    to bound autodiff temporaries."""
    rng = np.random.default_rng(seed)
    Ts = 1.0 / N
    if problem == "cartpole":
        ocp, x0 = problems.make_cartpole(Ts), problems.cartpole_x0().numpy()
        roll = cartpole_rollout_np
    elif problem == "pendulum":
        ocp, x0 = problems.make_pendulum(Ts), problems.pendulum_x0().numpy()
        roll = pendulum_rollout_np
    else:
        raise ValueError(problem)
    u0 = 0.1 * rng.standard_normal((N, 1))
    if x0_noise:
        x0 = x0 + x0_noise * rng.standard_normal(x0.shape)
        x0[1 if problem == "cartpole" else 0] %= 2.0 * math.pi
    xs = roll(u0, x0, Ts)
    x = torch.as_tensor(xs, dtype=torch.float64, device=device)
    u = torch.as_tensor(u0, dtype=torch.float64, device=device)
    parts = []
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        parts.append(compute_derivatives(ocp, x[lo:hi + 1], u[lo:hi], bp))
    d = type(parts[0])(*(torch.cat([getattr(p, f) for p in parts]) for f in parts[0]._fields))
    lamT = grad(ocp.final_cost)(x[-1])
    lam = par_costates(ocp, x[-1], d)
    ru, Q, R, M = compute_lqr_params(lam, d)
    cons = vmap(ocp.constraints)(x[:-1], u).reshape(N, -1).contiguous()
    cost = ocp.total_cost(x, u, bp)
    return dict(ocp=ocp, d=d, fx=d.fx, fu=d.fu, cx=d.cx, cu=d.cu, lamT=lamT, lam=lam, ru=ru, Q=Q, R=R, M=M, x=x, u=u,
                cons=cons, cost=cost, x0=x0, u0=u0, Ts=Ts)


def synthetic_lq(N, nx, nu, device, seed=0, dt=None):
    """Well-conditioned random time-varying LQ data (fx, fu, ru, Q, R, M) generated on the device —
    for timing at horizons where the torch.func set-up would dominate (time-sharded sweeps)."""
    g = torch.Generator(device=device).manual_seed(seed)
    o = dict(dtype=torch.float64, device=device, generator=g)
    dt = dt if dt is not None else min(0.5, 10.0 / N)
    eye = lambda n: torch.eye(n, dtype=torch.float64, device=device)
    fx = eye(nx) + dt * torch.randn(N, nx, nx, **o)
    fu = dt * torch.randn(N, nx, nu, **o) + 0.5 * dt
    Lq = 0.3 * torch.randn(N, nx, nx, **o)
    Q = Lq @ Lq.transpose(1, 2) + eye(nx) * (0.5 + torch.rand(N, 1, 1, **o))
    Lr = 0.3 * torch.randn(N, nu, nu, **o)
    R = Lr @ Lr.transpose(1, 2) + eye(nu) * (0.5 + torch.rand(N, 1, 1, **o))
    M = 0.06 * torch.randn(N, nx, nu, **o)
    ru = torch.randn(N, nu, **o)
    return fx.contiguous(), fu.contiguous(), ru.contiguous(), Q.contiguous(), R.contiguous(), M.contiguous()

"""The reference's example workloads as torch callables (float64).

These are *problem definitions* (user code in the reference), restated on torch so that
`torch.func` can differentiate them:
  pendulum        ref examples/pendulum_runtime.py:19-72,88-90
  cartpole        ref examples/cartpole_runtime.py:18-81,99-101
  linear (LQ)     ref examples/linear_demo_cuda.py:19-55
  mpc LQT         ref examples/linear_mpc_parallel.py:24-64
"""
import math
import torch
from torch.func import vmap
from .optimal_control_problem import OCP
from .utils import wrap_angle, euler, discretize_dynamics


_const_cache = {}


def _const(kind, vals, like):
    """Constant vectors / diagonal matrices, created once per (device, dtype): building them with
    torch.tensor(...) inside the cost would issue a host-to-device copy per call (and cannot be
    captured in a CUDA graph)."""
    key = (kind, tuple(vals), like.dtype, like.device)
    t = _const_cache.get(key)
    if t is None:
        t = torch.tensor(vals, dtype=like.dtype, device=like.device)
        if kind == "diag":
            t = torch.diag(t)
        _const_cache[key] = t
    return t


def _diag(vals, like):
    return _const("diag", vals, like)


def _vec(vals, like):
    return _const("vec", vals, like)


# ------------------------------------------------------------------ pendulum (nx=2, nu=1, nc=2)
def make_pendulum(Ts: float, control_bound: float = 5.0) -> OCP:
    def constraints(state, control):
        return torch.hstack((control - control_bound, -control - control_bound))

    def _err(state):
        angle, ang_vel = state[0], state[1]
        return torch.stack((wrap_angle(angle), ang_vel)) - _vec([math.pi, 0.0], state)

    def final_cost(state):
        e = _err(state)
        return 0.5 * e @ _diag([1e0, 1e-1], state) @ e

    def stage_cost(state, action, bp):
        e = _err(state)
        c = 0.5 * e @ _diag([1e0, 1e-1], state) @ e
        c = c + 0.5 * action @ _diag([1e-3], state) @ action
        return c - bp * torch.sum(torch.log(-constraints(state, action)))

    def total_cost(states, controls, bp):
        ct = vmap(stage_cost, in_dims=(0, 0, None))(states[:-1], controls, bp)
        return final_cost(states[-1]) + torch.sum(ct)

    def ode(state, action):
        gravity, length, mass, damping = 9.81, 1.0, 1.0, 1e-3
        position, velocity = state[0], state[1]
        acc = -gravity / length * torch.sin(position) + (action - damping * velocity) / (mass * length ** 2)
        return torch.hstack((velocity, acc))

    from . import plants
    return plants.register(OCP(euler(ode, Ts), constraints, stage_cost, final_cost, total_cost), "pendulum", Ts,
                           control_bound)


def pendulum_x0(dtype=torch.float64, device="cpu"):
    return torch.tensor([0.1 % (2.0 * math.pi), -0.1], dtype=dtype, device=device)


# ------------------------------------------------------------------ cartpole (nx=4, nu=1, nc=2)
def make_cartpole(Ts: float, control_bound: float = 50.0) -> OCP:
    def constraints(state, control):
        return torch.stack((control[0] - control_bound, -control[0] - control_bound))

    def _err(state):
        w = torch.stack((state[0], wrap_angle(state[1]), state[2], state[3]))
        return w - _vec([0.0, math.pi, 0.0, 0.0], state)

    def final_cost(state):
        e = _err(state)
        return 0.5 * e @ _diag([1e0, 1e1, 1e-1, 1e-1], state) @ e

    def stage_cost(state, action, bp):
        e = _err(state)
        c = 0.5 * e @ _diag([1e0, 1e1, 1e-1, 1e-1], state) @ e
        c = c + 0.5 * action @ _diag([1e-3], state) @ action
        return c - bp * torch.sum(torch.log(-constraints(state, action)))

    def total_cost(states, controls, bp):
        ct = vmap(stage_cost, in_dims=(0, 0, None))(states[:-1], controls, bp)
        return final_cost(states[-1]) + torch.sum(ct)

    def ode(state, action):
        gravity, pole_length, cart_mass, pole_mass = 9.81, 0.5, 10.0, 1.0
        total_mass = cart_mass + pole_mass
        pole_position, cart_velocity, pole_velocity = state[1], state[2], state[3]
        sth, cth = torch.sin(pole_position), torch.cos(pole_position)
        cart_acc = (action + pole_mass * sth * (pole_length * pole_velocity ** 2 + gravity * cth)) / (
            cart_mass + pole_mass * sth ** 2)
        pole_acc = (-action * cth - pole_mass * pole_length * pole_velocity ** 2 * cth * sth
                    - total_mass * gravity * sth) / (pole_length * cart_mass + pole_length * pole_mass * sth ** 2)
        return torch.hstack((cart_velocity, pole_velocity, cart_acc, pole_acc))

    from . import plants
    return plants.register(OCP(euler(ode, Ts), constraints, stage_cost, final_cost, total_cost), "cartpole", Ts,
                           control_bound)


def cartpole_x0(dtype=torch.float64, device="cpu"):
    return torch.tensor([0.01, (-0.01) % (2.0 * math.pi), 0.01, -0.01], dtype=dtype, device=device)


# ------------------------------------------------------------------ double integrator
def _double_integrator_ode(state, control):
    # xdot = [[0,1],[0,0]] x + [[0],[1]] u   (ref examples/linear_demo_cuda.py:19-22)
    return torch.hstack((state[1:2], control[0:1]))


def make_linear_demo(step: float = 0.1, control_bound=None) -> OCP:
    """LQ problem through the IP API.  With control_bound=None the constraint is the
    reference's dummy `-1.0` (ref examples/linear_demo_cuda.py:30-31); with a bound it is the
    box-constrained extension named in BASELINE.json config 3 (no reference script)."""
    dynamics = discretize_dynamics(_double_integrator_ode, step, 1)

    def constraints(state, control):
        if control_bound is None:
            return -1.0 + 0.0 * control[:1]
        return torch.hstack((control - control_bound, -control - control_bound))

    def stage_cost(state, control, bp):
        c = 0.5 * state @ _diag([1e2, 1e0], state) @ state + 0.5 * 1e-1 * (control @ control)
        if control_bound is not None:
            c = c - bp * torch.sum(torch.log(-constraints(state, control)))
        return c

    def final_cost(state):
        return 0.5 * state @ _diag([1e2, 1e0], state) @ state

    def total_cost(states, controls, bp):
        ct = vmap(stage_cost, in_dims=(0, 0, None))(states[:-1], controls, bp)
        return final_cost(states[-1]) + torch.sum(ct)

    return OCP(dynamics, constraints, stage_cost, final_cost, total_cost)


def make_mpc_lqt_terms(T: int = 5, step: float = 1e-3, dtype=torch.float64, device="cpu"):
    """The 13 LQT fields of ref examples/linear_mpc_parallel.py:24-64, in `LQT` order."""
    from torch.func import jacfwd
    dynamics = discretize_dynamics(_double_integrator_ode, step, 1)
    x0 = torch.tensor([2.0, 1.0], dtype=torch.float64)
    u0 = torch.zeros(1, dtype=torch.float64)
    A = jacfwd(dynamics, 0)(x0, u0)
    B = jacfwd(dynamics, 1)(x0, u0)
    nx, nu = 2, 1
    rep = lambda m: m.unsqueeze(0).repeat(T, *([1] * m.dim())).contiguous()
    Q = torch.diag(torch.tensor([1e2, 1e0], dtype=torch.float64))
    R = 1e-1 * torch.eye(nu, dtype=torch.float64)
    P = torch.diag(torch.tensor([1e2, 1e0], dtype=torch.float64))
    fields = (rep(A), rep(B), torch.zeros(T, nx, dtype=torch.float64), P, torch.eye(nx, dtype=torch.float64),
              torch.zeros(nx, dtype=torch.float64), rep(Q), rep(torch.eye(nx, dtype=torch.float64)),
              torch.zeros(T, nx, dtype=torch.float64), rep(R), rep(torch.eye(nu, dtype=torch.float64)),
              torch.zeros(T, nu, dtype=torch.float64), torch.zeros(T, nx, nu, dtype=torch.float64))
    return tuple(f.to(device=device, dtype=dtype) for f in fields), x0.to(device=device, dtype=dtype)

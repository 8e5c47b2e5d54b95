"""Host-framework helpers with the reference's names (ref noc/utils.py:8-63), on torch."""
import math
from typing import Callable
import torch


def wrap_angle(x):
    """Wrap to [0, 2*pi) — ref noc/utils.py:8-10 (`%` is floor-mod, like torch.remainder)."""
    return x % (2.0 * math.pi)


def runge_kutta(state, action, ode: Callable, step: float):   # ref noc/utils.py:13-23
    k1 = ode(state, action)
    k2 = ode(state + 0.5 * step * k1, action)
    k3 = ode(state + 0.5 * step * k2, action)
    k4 = ode(state + step * k3, action)
    return state + step / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def discretize_dynamics(ode: Callable, simulation_step: float, downsampling: int):  # ref :26-47
    def dynamics(state, action):
        for _ in range(downsampling):
            state = runge_kutta(state, action, ode, simulation_step)
        return state

    return dynamics


def euler(ode: Callable, simulation_step: float):   # ref noc/utils.py:50-54
    def dynamics(state, control):
        return state + simulation_step * ode(state, control)

    return dynamics


def rollout(dynamics, controls, initial_state):
    """Serial nonlinear rollout, (N,nu),(nx,) -> (N+1,nx) — ref noc/utils.py:57-63.
    O(N) sequential, once per barrier stage and outside the Newton step.  The user's `dynamics` is evaluated by
    the host framework on the device of `controls` (callables that capture CUDA tensors work for every horizon);
    N tiny host-framework launches per step make this slow, which is why `noc.initial_rollout` switches to the
    parallel rollout below from a few hundred steps on."""
    us = controls.detach().to(torch.float64)
    x = initial_state.detach().to(device=us.device, dtype=torch.float64)
    xs = [x]
    with torch.no_grad():
        for k in range(us.shape[0]):
            x = dynamics(x, us[k])
            xs.append(x)
    return torch.stack(xs)


def rollout_parallel(dynamics, controls, initial_state, x_guess=None, tol=1e-14, max_iter=200):
    """Parallel-in-time nonlinear rollout (SURVEY.md §8f "next" #2; no reference counterpart — the
    reference's `rollout` is a serial `lax.scan`, ref noc/utils.py:57-63).

    Newton iteration on the rollout equations x_{k+1} = f(x_k, u_k): linearise along the current guess,
    F_k = df/dx(x_k, u_k), c_k = f(x_k, u_k) - F_k x_k, and solve the resulting affine recursion for all
    k at once with the forward affine scan kernel (ipoc_affine_scan_f64).  Every iteration fixes at
    least one more leading state exactly, convergence is quadratic in practice (a handful of
    iterations for the example plants), and the fixed point IS the serial rollout (same f evaluated at
    the same states, up to rounding).  Returns (states (N+1,nx), iterations); falls back to the serial
    rollout if it has not converged after max_iter iterations."""
    from torch.func import vmap, jacrev
    from .noc import affine_scan
    dev = controls.device
    N = controls.shape[0]
    x0 = initial_state.to(dev)
    X = x_guess if x_guess is not None else x0.unsqueeze(0).expand(N + 1, -1).contiguous()
    X = X.clone()
    X[0] = x0
    f = vmap(dynamics)
    fx = vmap(jacrev(dynamics, 0))
    with torch.no_grad():
        for it in range(1, max_iter + 1):
            fv = f(X[:-1], controls)
            F = fx(X[:-1], controls).contiguous()
            c = (fv - (F @ X[:-1].unsqueeze(-1)).squeeze(-1)).contiguous()
            Xn = affine_scan(F, c, x0, reverse=False, transpose=False)
            err = float((Xn - X).abs().max())
            scale = 1.0 + float(Xn.abs().max())
            X = Xn
            if not (err == err) or not bool(torch.isfinite(Xn).all()):   # NaN / overflow: diverged, use the serial loop
                break
            if err <= tol * scale:
                X = torch.cat((x0.unsqueeze(0), f(X[:-1], controls)))   # states are exactly f of their predecessor
                return X.contiguous(), it
    return rollout(dynamics, controls, initial_state), -1

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
    O(N) sequential and outside the Newton step; runs on the CPU in float64 (N tiny
    host-framework ops per step would be launch-bound on the GPU) and returns on the
    device of `controls`."""
    dev = controls.device
    us = controls.detach().to("cpu", torch.float64)
    x = initial_state.detach().to("cpu", torch.float64)
    xs = [x]
    with torch.no_grad():
        for k in range(us.shape[0]):
            x = dynamics(x, us[k])
            xs.append(x)
    return torch.stack(xs).to(dev)

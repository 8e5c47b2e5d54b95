"""Shared test helpers: golden -> oracle containers, error metrics, synthetic LQ data."""
import numpy as np
from oracle.noc_np import Derivatives

STEP_FIXTURES = ["step_pendulum_N64", "step_pendulum_N33_warm", "step_cartpole_N100", "step_cartpole_N257_warm"]


def derivs_from_golden(g):
    return Derivatives(*(g["d_" + f] for f in Derivatives._fields))


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def random_lq(rng, N, nx, nu, batch=None, coupling=0.3, dt=None):
    """Well-conditioned random time-varying LQ data in the Newton-step layout
    (fx, fu, ru, Q, R, M).  Q is SPD, R is SPD, M small, fx = I + dt*random."""
    shape = (N,) if batch is None else (batch, N)
    dt = dt if dt is not None else min(0.5, 10.0 / N)
    fx = np.eye(nx) + dt * rng.standard_normal(shape + (nx, nx))
    fu = dt * rng.standard_normal(shape + (nx, nu)) + dt * 0.5
    Lq = rng.standard_normal(shape + (nx, nx)) * 0.3
    Q = Lq @ np.swapaxes(Lq, -1, -2) + np.eye(nx) * (0.5 + rng.random(shape + (1, 1)))
    Lr = rng.standard_normal(shape + (nu, nu)) * 0.3
    R = Lr @ np.swapaxes(Lr, -1, -2) + np.eye(nu) * (0.5 + rng.random(shape + (1, 1)))
    M = coupling * 0.2 * rng.standard_normal(shape + (nx, nu))
    ru = rng.standard_normal(shape + (nu,))
    return fx, fu, ru, Q, R, M

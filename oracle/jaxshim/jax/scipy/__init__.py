"""Fake `jax.scipy` (only linalg.solve, ref noc/par_interior_point_newton.py:63-64)."""
import torch
from .._core import wrap


class _Linalg:
    @staticmethod
    def solve(a, b):
        return wrap(torch.linalg.solve(a, b))


linalg = _Linalg()

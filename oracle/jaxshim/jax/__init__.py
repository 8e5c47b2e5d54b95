"""Fake top-level `jax` package (see _core.py).  TEST INFRASTRUCTURE ONLY."""
import torch
from torch import func as _F
from . import _core
from ._core import wrap as _wrap
from . import numpy, lax, scipy, random  # noqa: F401


def _argnums(a):
    return a


def _tensorize(out):
    # JAX broadcasts Python scalars returned from a vmapped function
    # (ref examples/linear_demo_cuda.py:30-31 returns the literal -1.0)
    if isinstance(out, (tuple, list)):
        return type(out)(_tensorize(o) for o in out)
    if isinstance(out, (int, float)):
        return torch.tensor(float(out))
    return out


def vmap(f, in_axes=0, out_axes=0):
    def g(*args):
        return _wrap(_F.vmap(lambda *a: _tensorize(f(*a)), in_dims=in_axes, out_dims=out_axes)(*args))
    return g


def grad(f, argnums=0):
    def g(*args):
        return _wrap(_F.grad(f, argnums=argnums)(*args))
    return g


def jacrev(f, argnums=0):
    def g(*args):
        return _wrap(_F.jacrev(f, argnums=argnums)(*args))
    return g


def jacfwd(f, argnums=0):
    def g(*args):
        return _wrap(_F.jacfwd(f, argnums=argnums)(*args))
    return g


def hessian(f, argnums=0):
    def g(*args):
        return _wrap(_F.hessian(f, argnums=argnums)(*args))
    return g


def jit(f, backend=None, **kw):
    return f


def block_until_ready(x):
    return x


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


class _Debug:
    @staticmethod
    def print(*a, **k):
        pass

    @staticmethod
    def breakpoint(*a, **k):
        pass


debug = _Debug()

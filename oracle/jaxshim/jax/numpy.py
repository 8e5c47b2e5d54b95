"""Fake `jax.numpy` on float64 torch CPU (see _core.py).  TEST INFRASTRUCTURE ONLY."""
import math
import torch
from ._core import Arr, wrap, asarr

pi = math.pi
inf = math.inf
float64 = torch.float64
ndarray = torch.Tensor


def _t(x):
    return x if isinstance(x, torch.Tensor) else asarr(x)


def array(x, dtype=None):
    return asarr(x, dtype)


asarray = array


def zeros(shape, dtype=None):
    return wrap(torch.zeros(shape, dtype=dtype or torch.float64))


def ones(shape, dtype=None):
    return wrap(torch.ones(shape, dtype=dtype or torch.float64))


def eye(n):
    return wrap(torch.eye(n, dtype=torch.float64))


def zeros_like(x):
    return wrap(torch.zeros_like(x))


def diag(x):
    return wrap(torch.diag(_t(x)))


def kron(a, b):
    return wrap(torch.kron(_t(a), _t(b)))


def vstack(xs):
    return wrap(torch.vstack([_t(x) for x in xs]))


def hstack(xs):
    return wrap(torch.hstack([_t(x) for x in xs]))


def tensordot(a, b, axes=2):
    return wrap(torch.tensordot(_t(a), _t(b), dims=axes))


def transpose(x, axes=None):
    if axes is None:
        return _t(x).T
    return wrap(_t(x).permute(*axes))


def all(x):  # noqa: A001
    return wrap(torch.all(_t(x)))


def max(x):  # noqa: A001
    return wrap(torch.max(_t(x)))


def abs(x):  # noqa: A001
    return wrap(torch.abs(_t(x)))


def sum(x, axis=None):  # noqa: A001
    return wrap(torch.sum(_t(x)) if axis is None else torch.sum(_t(x), dim=axis))


def mean(x):
    return wrap(torch.mean(_t(x)))


def median(x):
    return wrap(torch.median(_t(x)))


def sin(x):
    return wrap(torch.sin(_t(x)))


def cos(x):
    return wrap(torch.cos(_t(x)))


def log(x):
    return wrap(torch.log(_t(x)))


def sqrt(x):
    return wrap(torch.sqrt(_t(x)))


def where(c, a, b):
    c = _t(c)
    if not isinstance(a, torch.Tensor) and not isinstance(b, torch.Tensor):
        a = _t(float(a))
    return wrap(torch.where(c, a, b))


def logical_and(a, b):
    return wrap(torch.logical_and(_t(a), _t(b)))


def logical_or(a, b):
    return wrap(torch.logical_or(_t(a), _t(b)))


def logical_not(a):
    return wrap(torch.logical_not(_t(a)))


def maximum(a, b):
    a, b = _t(a), _t(b)
    return wrap(torch.maximum(a.to(torch.float64), b.to(torch.float64)))


def clip(x, lo, hi):
    return wrap(torch.clamp(_t(x), lo, hi))


def bool_(x):
    return wrap(torch.tensor(builtins_bool(x)))


def builtins_bool(x):
    return True if x else False


bool = bool_  # noqa: A001  (ref noc/seq_interior_point_newton.py:176 uses jnp.bool)


class _Linalg:
    @staticmethod
    def norm(x):
        return wrap(torch.linalg.norm(_t(x).reshape(-1)))

    @staticmethod
    def eigh(x):
        w, v = torch.linalg.eigh(_t(x))
        return wrap(w), wrap(v)

    @staticmethod
    def inv(x):
        return wrap(torch.linalg.inv(_t(x)))

    @staticmethod
    def solve(a, b):
        return wrap(torch.linalg.solve(_t(a), _t(b)))


linalg = _Linalg()

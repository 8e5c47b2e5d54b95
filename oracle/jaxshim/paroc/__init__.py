"""Stand-in for the absent `paroc` package, used ONLY by tests/golden/gen_golden.py so that
the reference's own `noc/par_interior_point_newton.py` can be imported and executed.
TEST INFRASTRUCTURE ONLY.  The arithmetic is the NumPy restatement in oracle/paroc_np.py."""
import numpy as _np
import torch as _torch
from jax._core import wrap as _wrap
from oracle import paroc_np as _p
from .lqt_problem import LQT  # noqa: F401


def _n(x):
    return x.detach().numpy() if isinstance(x, _torch.Tensor) else _np.asarray(x, dtype=_np.float64)


def _lqt(lqt):
    return _p.LQT(*(_n(a) for a in lqt))


def _w(x):
    if isinstance(x, (bool, _np.bool_)):
        return _wrap(_torch.tensor(bool(x)))
    return _wrap(_torch.as_tensor(_np.ascontiguousarray(x, dtype=_np.float64)))


def par_bwd_pass(lqt):
    return tuple(_w(o) for o in _p.par_bwd_pass(_lqt(lqt)))


def par_fwd_pass(lqt, x0, Kx, d):
    return tuple(_w(o) for o in _p.par_fwd_pass(_lqt(lqt), _n(x0), _n(Kx), _n(d)))


def seq_bwd_pass(lqt):
    return tuple(_w(o) for o in _p.seq_bwd_pass(_lqt(lqt)))


def seq_fwd_pass(lqt, x0, Kx, d):
    return tuple(_w(o) for o in _p.seq_fwd_pass(_lqt(lqt), _n(x0), _n(Kx), _n(d)))

"""`paroc.lqt_problem.LQT`: 13 positional fields, order pinned by the reference's call sites
(ref noc/par_interior_point_newton.py:69-83, examples/linear_mpc_parallel.py:64)."""
from typing import NamedTuple, Any


class LQT(NamedTuple):
    A: Any
    B: Any
    c: Any
    XT: Any
    HT: Any
    rT: Any
    X: Any
    H: Any
    r: Any
    U: Any
    Z: Any
    s: Any
    M: Any

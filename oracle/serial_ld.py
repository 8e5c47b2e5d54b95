"""ctypes front-end of oracle/serial_ld.c: the reference's in-tree SEQUENTIAL Newton step
(ref noc/seq_interior_point_newton.py:42-90) restated in C, once in x87 `long double` (the arbiter for
N = 1e6 conditioning questions) and once in plain double (the serial O(N) CPU comparator "B2").

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The shared object is built by
`__graft_entry__.build()`; if it is missing it is compiled here with gcc (present in this image)."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libserial_ld.so")
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, "serial_ld.c"), os.path.join(_HERE, "serial_body.inc")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src[0], "-lm"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        P, I, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        for sfx in ("ld", "f64"):
            f = getattr(L, "ipoc_oracle_seq_newton_" + sfx)
            f.restype, f.argtypes = I, [I, I, I] + [P] * 7 + [D] + [P] * 7
            f = getattr(L, "ipoc_oracle_seq_costates_" + sfx)
            f.restype, f.argtypes = I, [I, I, P, P, P, P]
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def seq_newton(fx, fu, ru, Q, R, M, reg, VxxN=None, precision="ld"):
    """-> dx (N+1,nx), du (N,nu), K (N,nu,nx), k (N,nu), dV, convex.  `reg` = rp of ref :51 (the par path's
    reg_param*||cu||); VxxN defaults to Q[0] (the par path's terminal weight)."""
    fx, fu, ru, Q, R, M = (_c(a) for a in (fx, fu, ru, Q, R, M))
    N, nx, nu = fx.shape[0], fx.shape[1], fu.shape[2]
    V = _c(Q[0] if VxxN is None else VxxN)
    K, k = np.empty((N, nu, nx)), np.empty((N, nu))
    dx, du = np.empty((N + 1, nx)), np.empty((N, nu))
    dV, convex = ctypes.c_double(), ctypes.c_int()
    work = np.empty(N * nu * (nx + 1))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    fn = getattr(lib(), "ipoc_oracle_seq_newton_" + precision)
    rc = fn(N, nx, nu, p(fx), p(fu), p(ru), p(Q), p(R), p(M), p(V), float(reg), p(K), p(k), p(dx), p(du),
            ctypes.cast(ctypes.byref(dV), ctypes.c_void_p), ctypes.cast(ctypes.byref(convex), ctypes.c_void_p), p(work))
    if rc:
        raise ValueError("serial_ld: unsupported dimensions")
    return dx, du, K, k, dV.value, bool(convex.value)


def seq_costates(fx, cx, lamT, precision="ld"):
    """ref noc/costates.py:43-54 -> lam (N+1, nx)."""
    fx, cx, lamT = _c(fx), _c(cx), _c(lamT)
    N, nx = fx.shape[0], fx.shape[1]
    lam = np.empty((N + 1, nx))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    if getattr(lib(), "ipoc_oracle_seq_costates_" + precision)(N, nx, p(fx), p(cx), p(lamT), p(lam)):
        raise ValueError("serial_ld: unsupported dimensions")
    return lam

"""JAX front-end of the same C ABI: `jax.ffi` custom calls registered for CUDA only.

UNTESTED IN THIS IMAGE — JAX, jaxlib and the XLA FFI headers are not installed, so neither this module
nor csrc/ipoc_xla_ffi.cc can be exercised here; the PyTorch/ctypes binding (`_lib.py`, `noc.py`,
`paroc.py`) is the one the tests run.  With JAX available, `par_Newton` of the reference
(ref noc/par_interior_point_newton.py:107-124) becomes the body of `par_Newton` below and the rest of the
reference (its `lax.while_loop`s, `vmap`ped autodiff) runs unchanged.
"""
import ctypes
import os

try:  # import guard: the product package must import without JAX
    import jax
    import jax.numpy as jnp
    import numpy as np
    HAVE_JAX = True
except Exception:  # pragma: no cover
    HAVE_JAX = False

_HERE = os.path.dirname(os.path.abspath(__file__))
XLA_LIB_PATH = os.path.join(os.path.dirname(_HERE), "libipoc_xla.so")
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libipoc.so")
_registered = False
_TARGETS = {"ipoc_newton_step": "IpocNewtonStep", "ipoc_affine_scan": "IpocAffineScan", "ipoc_lqt_bwd": "IpocLqtBwd",
            "ipoc_lqt_fwd": "IpocLqtFwd", "ipoc_reductions": "IpocReductions", "ipoc_accept_update": "IpocAcceptUpdate",
            "ipoc_costates": "IpocCostates", "ipoc_newton_attempt": "IpocNewtonAttempt"}


def register():
    """Register the FFI targets (platform CUDA only — there is no CPU implementation to fall back to)."""
    global _registered
    if not HAVE_JAX:
        raise RuntimeError("jax is not installed; use the PyTorch binding (ipoc_b200.noc / ipoc_b200.paroc)")
    if _registered:
        return
    lib = ctypes.CDLL(XLA_LIB_PATH)
    for target, symbol in _TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")
    _registered = True


_core = None


def _ws_bytes(kind, N, nx, nu, batch):
    """Workspace size through plain ctypes on libipoc.so — no torch import on the JAX side."""
    global _core
    if _core is None:
        _core = ctypes.CDLL(LIB_PATH)
        _core.ipoc_workspace_bytes.restype = ctypes.c_size_t
        _core.ipoc_workspace_bytes.argtypes = [ctypes.c_int] * 5
    return int(_core.ipoc_workspace_bytes(kind, N, nx, nu, batch))


def newton_step(fx, fu, ru, Q, R, M, reg):
    """(dx, du, Kx, d, pred, feasible) for (N, ...) inputs; jit/while_loop compatible."""
    register()
    N, nx, nu = fx.shape[0], fx.shape[1], fu.shape[-1]
    f64 = jnp.float64
    out_types = (
        jax.ShapeDtypeStruct((1, N + 1, nx), f64), jax.ShapeDtypeStruct((1, N, nu), f64),
        jax.ShapeDtypeStruct((1, N, nu, nx), f64), jax.ShapeDtypeStruct((1, N, nu), f64),
        jax.ShapeDtypeStruct((1,), f64), jax.ShapeDtypeStruct((1,), jnp.int32),
        jax.ShapeDtypeStruct((_ws_bytes(0, N, nx, nu, 1),), jnp.uint8),
    )
    dx, du, Kx, d, pred, feas, _ = jax.ffi.ffi_call("ipoc_newton_step", out_types)(
        fx[None], fu[None], ru[None], Q[None], R[None], M[None], jnp.reshape(reg, (1,)))
    return dx[0], du[0], Kx[0], d[0], pred[0], feas[0] != 0


def par_Newton(nominal_states, d, reg_param, ru, Q, R, M):
    """Drop-in body for the reference's `par_Newton` (same arguments, same 5 results)."""
    reg = reg_param * jnp.linalg.norm(d.cu)
    dx, du, _, _, pred, feasible = newton_step(d.fx, d.fu, ru, Q, R, M, reg)
    return dx, du, pred, feasible, ru


def par_costates_scan(fx, cx, lamda_T):
    """lambda (N+1, nx) — replaces `par_scan` in ref noc/costates.py:15-16,34-40."""
    register()
    N, nx = fx.shape[0], fx.shape[1]
    out_types = (jax.ShapeDtypeStruct((1, N + 1, nx), jnp.float64),
                 jax.ShapeDtypeStruct((_ws_bytes(3, N, nx, 1, 1),), jnp.uint8))
    lam, _ = jax.ffi.ffi_call("ipoc_affine_scan", out_types)(fx[None], cx[None], lamda_T[None],
                                                             reverse=np.int32(1), transpose=np.int32(1))
    return lam[0]


def par_bwd_pass_effective(A, B, c, X, U, M, q, p, ST, vT):
    """`paroc.par_bwd_pass` on effective LQT terms (H, Z folded in by the caller) -> (Kx, d, S, v, pred, feasible);
    ref call sites noc/par_interior_point_newton.py:120, examples/linear_mpc_parallel.py:68."""
    register()
    T, nx, nu = A.shape[0], A.shape[1], B.shape[-1]
    f64 = jnp.float64
    out_types = (jax.ShapeDtypeStruct((1, T, nu, nx), f64), jax.ShapeDtypeStruct((1, T, nu), f64),
                 jax.ShapeDtypeStruct((1, T + 1, nx, nx), f64), jax.ShapeDtypeStruct((1, T + 1, nx), f64),
                 jax.ShapeDtypeStruct((1,), f64), jax.ShapeDtypeStruct((1,), jnp.int32),
                 jax.ShapeDtypeStruct((_ws_bytes(1, T, nx, nu, 1),), jnp.uint8))
    Kx, d, S, v, pred, feas, _ = jax.ffi.ffi_call("ipoc_lqt_bwd", out_types)(
        A[None], B[None], c[None], X[None], U[None], M[None], q[None], p[None], ST[None], vT[None])
    return Kx[0], d[0], S[0], v[0], pred[0], feas[0] != 0


def par_fwd_pass_effective(A, B, c, Kx, d, x0):
    """`paroc.par_fwd_pass` -> (u, x); ref noc/par_interior_point_newton.py:121-123, linear_mpc_parallel.py:69."""
    register()
    T, nx, nu = A.shape[0], A.shape[1], B.shape[-1]
    out_types = (jax.ShapeDtypeStruct((1, T, nu), jnp.float64), jax.ShapeDtypeStruct((1, T + 1, nx), jnp.float64),
                 jax.ShapeDtypeStruct((_ws_bytes(2, T, nx, nu, 1),), jnp.uint8))
    u, x, _ = jax.ffi.ffi_call("ipoc_lqt_fwd", out_types)(A[None], B[None], c[None], Kx[None], d[None], x0[None])
    return u[0], x[0]


def reductions(ru, cu, cons):
    """(max|ru|, ||cu||_F, all(cons <= 0)) — ref :158, :116, :45-47."""
    register()
    N, nu, nc = ru.shape[0], ru.shape[1], cons.shape[-1]
    out_types = (jax.ShapeDtypeStruct((1,), jnp.float64), jax.ShapeDtypeStruct((1,), jnp.float64),
                 jax.ShapeDtypeStruct((1,), jnp.int32),
                 jax.ShapeDtypeStruct((_ws_bytes(4, N, max(nu, nc), nu, 1),), jnp.uint8))
    hu, cn, feas, _ = jax.ffi.ffi_call("ipoc_reductions", out_types)(ru[None], cu[None], cons[None])
    return hu[0], cn[0], feas[0] != 0


def accept_update(cost, new_cost, traj_feasible, pred, bwd_feasible, rp, r_inc):
    """Functional A8 (ref :159-173) -> (rp', r_inc', success, gain_ratio); scalars as shape-(1,) arrays."""
    register()
    f1, i1 = jax.ShapeDtypeStruct((1,), jnp.float64), jax.ShapeDtypeStruct((1,), jnp.int32)
    r = lambda a, t: jnp.reshape(a, (1,)).astype(t)
    rp2, ri2, succ, gain = jax.ffi.ffi_call("ipoc_accept_update", (f1, f1, i1, f1))(
        r(cost, jnp.float64), r(new_cost, jnp.float64), r(traj_feasible, jnp.int32), r(pred, jnp.float64),
        r(bwd_feasible, jnp.int32), r(rp, jnp.float64), r(r_inc, jnp.float64))
    return rp2[0], ri2[0], succ[0] != 0, gain[0]

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
_registered = False


def register():
    """Register the FFI targets (platform CUDA only — there is no CPU implementation to fall back to)."""
    global _registered
    if not HAVE_JAX:
        raise RuntimeError("jax is not installed; use the PyTorch binding (ipoc_b200.noc / ipoc_b200.paroc)")
    if _registered:
        return
    lib = ctypes.CDLL(XLA_LIB_PATH)
    jax.ffi.register_ffi_target("ipoc_newton_step", jax.ffi.pycapsule(lib.IpocNewtonStep), platform="CUDA")
    jax.ffi.register_ffi_target("ipoc_affine_scan", jax.ffi.pycapsule(lib.IpocAffineScan), platform="CUDA")
    _registered = True


def _ws_bytes(kind, N, nx, nu, batch):
    from . import _lib
    return int(_lib.lib().ipoc_workspace_bytes(kind, N, nx, nu, batch))


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

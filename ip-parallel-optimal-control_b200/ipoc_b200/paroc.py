"""The `paroc` API surface the reference imports (`from paroc import par_bwd_pass, par_fwd_pass`,
`from paroc.lqt_problem import LQT` — ref noc/par_interior_point_newton.py:6-7,
examples/linear_mpc_parallel.py:6-8), served by the sm_100a kernels behind include/ipoc.h.

LQT problem (13 positional fields, order pinned by the reference's call sites):
    x+ = A x + B u + c,   cost = 1/2 (HT xT - rT)' XT (HT xT - rT)
       + sum 1/2 (Hx - r)' X (Hx - r) + 1/2 (Zu - s)' U (Zu - s) + (Hx - r)' M (Zu - s)
All tensors float64 on a CUDA device.  A leading batch axis on every field is accepted
(extension: the reference solves one problem at a time).
"""
from typing import NamedTuple
import ctypes
import torch
from . import _lib as L


class LQT(NamedTuple):
    A: torch.Tensor
    B: torch.Tensor
    c: torch.Tensor
    XT: torch.Tensor
    HT: torch.Tensor
    rT: torch.Tensor
    X: torch.Tensor
    H: torch.Tensor
    r: torch.Tensor
    U: torch.Tensor
    Z: torch.Tensor
    s: torch.Tensor
    M: torch.Tensor


def _is_identity_stack(H):
    n = H.shape[-1]
    if H.shape[-2] != n:
        return False
    return bool(torch.all(H == torch.eye(n, dtype=H.dtype, device=H.device)))


def _effective(lqt: LQT):
    """Fold H, Z and the references r, s into plain LQ terms (X, U, M, q, p, ST, vT).
    Cheap batched host-framework ops; identity H/Z (both reference call sites) skip the products."""
    A, B, c, XT, HT, rT, X, H, r, U, Z, s, M = lqt
    mv = lambda m, v: (m @ v.unsqueeze(-1)).squeeze(-1)
    tr = lambda m: m.transpose(-1, -2)
    idH, idZ = _is_identity_stack(H), _is_identity_stack(Z)
    qx = mv(X, r) + mv(M, s)
    pu = mv(U, s) + mv(tr(M), r)
    Xe, Ue, Me = X, U, M
    if not idH:
        Xe, Me, qx = tr(H) @ X @ H, tr(H) @ M, mv(tr(H), qx)
    if not idZ:
        Ue, Me, pu = tr(Z) @ U @ Z, Me @ Z, mv(tr(Z), pu)
    ST = tr(HT) @ XT @ HT
    vT = mv(tr(HT), mv(XT, rT))
    return Xe, Ue, Me, -qx, -pu, ST, vT


def par_bwd_pass(lqt: LQT):
    """-> (Kx (T,nu,nx), d (T,nu), S (T+1,nx,nx), v (T+1,nx), pred_reduction (), feasible () bool)
    Control law u = -Kx x + d; V_k(x) = 1/2 x'S_k x - v_k'x."""
    lqt = LQT(*(L.dev_f64(t) for t in lqt))
    batched = lqt.A.dim() == 4
    if not batched:
        lqt = LQT(*(t.unsqueeze(0) for t in lqt))
    dev = lqt.A.device
    Bn, T, nx = lqt.A.shape[0], lqt.A.shape[1], lqt.A.shape[2]
    nu = lqt.B.shape[-1]
    L.require_supported(nx, nu)
    Xe, Ue, Me, q, p, ST, vT = (L.dev_f64(t) for t in _effective(lqt))
    o = dict(dtype=torch.float64, device=dev)
    Kx = torch.empty(Bn, T, nu, nx, **o)
    d = torch.empty(Bn, T, nu, **o)
    S = torch.empty(Bn, T + 1, nx, nx, **o)
    v = torch.empty(Bn, T + 1, nx, **o)
    pred = torch.empty(Bn, **o)
    feas = torch.empty(Bn, dtype=torch.int32, device=dev)
    ws, nbytes = L.workspace(L.WS_LQT_BWD, T, nx, nu, Bn, dev)
    with torch.cuda.device(dev):
        L.check(L.lib().ipoc_lqt_bwd_f64(
            T, nx, nu, Bn, L.ptr(lqt.A), L.ptr(lqt.B), L.ptr(lqt.c), L.ptr(Xe), L.ptr(Ue), L.ptr(Me), L.ptr(q),
            L.ptr(p), L.ptr(ST), L.ptr(vT), L.ptr(Kx), L.ptr(d), L.ptr(S), L.ptr(v), L.ptr(pred), L.ptr(feas),
            L.ptr(ws), nbytes, L.stream_ptr()))
    feas = feas != 0
    if not batched:
        return Kx[0], d[0], S[0], v[0], pred[0], feas[0]
    return Kx, d, S, v, pred, feas


def par_fwd_pass(lqt: LQT, x0, Kx, d):
    """-> (u (T,nu), x (T+1,nx)) with u_k = -Kx_k x_k + d_k, x_{k+1} = A x + B u + c."""
    A, B, c = (L.dev_f64(t) for t in lqt[:3])
    dev = A.device
    batched = A.dim() == 4
    x0, Kx, d = L.dev_f64(x0, dev), L.dev_f64(Kx, dev), L.dev_f64(d, dev)
    if not batched:
        A, B, c, x0, Kx, d = (t.unsqueeze(0) for t in (A, B, c, x0, Kx, d))
    Bn, T, nx = A.shape[0], A.shape[1], A.shape[2]
    nu = B.shape[-1]
    L.require_supported(nx, nu)
    o = dict(dtype=torch.float64, device=dev)
    u = torch.empty(Bn, T, nu, **o)
    x = torch.empty(Bn, T + 1, nx, **o)
    ws, nbytes = L.workspace(L.WS_LQT_FWD, T, nx, nu, Bn, dev)
    with torch.cuda.device(dev):
        L.check(L.lib().ipoc_lqt_fwd_f64(T, nx, nu, Bn, L.ptr(A), L.ptr(B), L.ptr(c), L.ptr(Kx), L.ptr(d),
                                         L.ptr(x0.contiguous()), L.ptr(u), L.ptr(x), L.ptr(ws), nbytes,
                                         L.stream_ptr()))
    if not batched:
        return u[0], x[0]
    return u, x


def seq_bwd_pass(lqt: LQT):
    """Serial twin of `par_bwd_pass` (ref examples/linear_mpc_parallel.py:8,74): the plain Riccati
    recursion, one thread walking the whole horizon (the same kernels with a single chunk per problem).
    -> (Kx, d, S, v)"""
    T = lqt.A.shape[-3]
    with L.tuning(leaf_chunk=T):
        Kx, d, S, v, _, _ = par_bwd_pass(lqt)
    return Kx, d, S, v


def seq_fwd_pass(lqt: LQT, x0, Kx, d):
    """Serial twin of `par_fwd_pass` (ref examples/linear_mpc_parallel.py:8,75) -> (u, x)."""
    T = lqt.A.shape[-3]
    with L.tuning(leaf_chunk=T):
        out = par_fwd_pass(lqt, x0, Kx, d)
    return out

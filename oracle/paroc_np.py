"""NumPy restatement of the `paroc` API used by the reference.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`paroc` (github.com/casiacob/parallel-optimal-control, un-pinned HEAD) is NOT in
/root/reference and not installable here, so its internals are restated from
the published algorithm — Särkkä & García-Fernández, "Temporal Parallelization
of Dynamic Programming and Linear Quadratic Control" (IEEE TAC 2023):
conditional value-function elements (A, b, C, eta, J) and their associative
combination — anchored on the reference's own call sites:

  * `LQT(A, B, c, XT, HT, rT, X, H, r, U, Z, s, M)`   13 positional fields,
        ref noc/par_interior_point_newton.py:69-83, examples/linear_mpc_parallel.py:64
  * `par_bwd_pass(lqt) -> (Kx, d, S, v, pred_reduction, feasible)`
        ref noc/par_interior_point_newton.py:120, examples/linear_mpc_parallel.py:68
  * `par_fwd_pass(lqt, x0, Kx, d) -> (u, x)`
        ref noc/par_interior_point_newton.py:121-123, examples/linear_mpc_parallel.py:69
  * `seq_bwd_pass(lqt) -> (Kx, d, S, v)`, `seq_fwd_pass(lqt, x0, Kx, d) -> (u, x)`
        ref examples/linear_mpc_parallel.py:74-75

LQT problem:  x_{k+1} = A_k x_k + B_k u_k + c_k,
  cost = 1/2 (HT x_T - rT)' XT (HT x_T - rT)
       + sum_k 1/2 (H x - r)' X (H x - r) + 1/2 (Z u - s)' U (Z u - s) + (H x - r)' M (Z u - s).
Value function convention V_k(x) = 1/2 x' S_k x - v_k' x (+const); control law u = -Kx x + d.

`pred_reduction` and `feasible` are not observable from the call sites; they are
fixed by how the caller uses them (gain_ratio = (new_cost-cost)/pred_reduction > 0
accepts a DEcrease, ref noc/par_interior_point_newton.py:164-166) and by the in-tree
sequential twin (dV = k'Qu + 1/2 k'Quu k, convex = all(eigh(Quu) > 0),
ref noc/seq_interior_point_newton.py:51-53,63):
    pred_reduction = -1/2 sum_k d_k' G_k d_k,   feasible = all_k (G_k > 0),
    G_k = Z'UZ + B' S_{k+1} B.
"""
from typing import NamedTuple
import numpy as np
from .assoc_scan import associative_scan


class LQT(NamedTuple):
    A: np.ndarray
    B: np.ndarray
    c: np.ndarray
    XT: np.ndarray
    HT: np.ndarray
    rT: np.ndarray
    X: np.ndarray
    H: np.ndarray
    r: np.ndarray
    U: np.ndarray
    Z: np.ndarray
    s: np.ndarray
    M: np.ndarray


def _T(a):
    return np.swapaxes(a, -1, -2)


def _mv(a, x):
    return np.einsum("...ij,...j->...i", a, x)


def effective_terms(lqt: LQT):
    """Fold H, Z and the tracking references into plain LQ terms.

    stage cost = 1/2 x'Xe x + 1/2 u'Ue u + x'Me u + q'x + p'u + const
    """
    A, B, c, XT, HT, rT, X, H, r, U, Z, s, M = (np.asarray(a, dtype=np.float64) for a in lqt)
    Xe = _T(H) @ X @ H
    Ue = _T(Z) @ U @ Z
    Me = _T(H) @ M @ Z
    q = -_mv(_T(H), _mv(X, r) + _mv(M, s))
    p = -_mv(_T(Z), _mv(U, s) + _mv(_T(M), r))
    ST = HT.T @ XT @ HT
    vT = HT.T @ (XT @ rT)
    return A, B, c, Xe, Ue, Me, q, p, ST, vT


def bwd_elements(lqt: LQT):
    """Per-step conditional value-function elements, k = 0..T-1, plus the terminal one."""
    A, B, c, Xe, Ue, Me, q, p, ST, vT = effective_terms(lqt)
    T, nx = A.shape[0], A.shape[1]
    UinvMt = np.linalg.solve(Ue, _T(Me))          # U^-1 M'
    UinvBt = np.linalg.solve(Ue, _T(B))           # U^-1 B'
    Uinvp = np.linalg.solve(Ue, p[..., None])[..., 0]
    Ae = A - B @ UinvMt
    be = c - _mv(B, Uinvp)
    Ce = B @ UinvBt
    Je = Xe - Me @ UinvMt
    etae = -q + _mv(Me, Uinvp)
    z = np.zeros
    Ae = np.concatenate([Ae, z((1, nx, nx))])
    be = np.concatenate([be, z((1, nx))])
    Ce = np.concatenate([Ce, z((1, nx, nx))])
    Je = np.concatenate([Je, ST[None]])
    etae = np.concatenate([etae, vT[None]])
    return Ae, be, Ce, etae, Je


def combine(e1, e2):
    """e1 = (i -> j) earlier segment, e2 = (j -> k) later segment; batched on axis 0."""
    A1, b1, C1, eta1, J1 = e1
    A2, b2, C2, eta2, J2 = e2
    n = A1.shape[-1]
    eye = np.eye(n)
    W = eye + C1 @ J2
    rhs = np.concatenate([A1, (b1 + _mv(C1, eta2))[..., None], C1 @ _T(A2)], axis=-1)
    sol = np.linalg.solve(W, rhs)
    A = A2 @ sol[..., :n]
    b = _mv(A2, sol[..., n]) + b2
    C = A2 @ sol[..., n + 1:] + C2
    Wt = eye + J2 @ C1
    rhs2 = np.concatenate([(eta2 - _mv(J2, b1))[..., None], J2 @ A1], axis=-1)
    sol2 = np.linalg.solve(Wt, rhs2)
    eta = _mv(_T(A1), sol2[..., 0]) + eta1
    J = _T(A1) @ sol2[..., 1:] + J1
    return A, b, C, eta, J


def _combine_rev(later, earlier):
    # associative_scan(reverse=True) hands the already-accumulated (later-in-time)
    # partial result as first operand
    return combine(earlier, later)


def value_functions_par(lqt: LQT):
    """S_k, v_k for k = 0..T via the reverse associative scan (suffix aggregates)."""
    elems = bwd_elements(lqt)
    _, _, _, eta, J = associative_scan(_combine_rev, elems, reverse=True)
    return J, eta


def gains(lqt: LQT, S, v):
    A, B, c, Xe, Ue, Me, q, p, ST, vT = effective_terms(lqt)
    Sn, vn = S[1:], v[1:]
    BtS = _T(B) @ Sn
    G = Ue + BtS @ B
    Kx = np.linalg.solve(G, _T(Me) + BtS @ A)
    rhs = -p + _mv(_T(B), vn - _mv(Sn, c))
    d = np.linalg.solve(G, rhs[..., None])[..., 0]
    pred = -0.5 * np.sum(np.einsum("ti,tij,tj->t", d, G, d))
    with np.errstate(invalid="ignore"):
        if np.all(np.isfinite(G)):
            feasible = bool(np.all(np.linalg.eigvalsh(G) > 0))
        else:
            feasible = False
    return Kx, d, pred, feasible


def par_bwd_pass(lqt: LQT):
    S, v = value_functions_par(lqt)
    Kx, d, pred, feasible = gains(lqt, S, v)
    return Kx, d, S, v, pred, feasible


def seq_bwd_pass(lqt: LQT):
    """Plain Riccati recursion (the serial twin the MPC example compares against)."""
    A, B, c, Xe, Ue, Me, q, p, ST, vT = effective_terms(lqt)
    T, nx = A.shape[0], A.shape[1]
    S = np.zeros((T + 1, nx, nx))
    v = np.zeros((T + 1, nx))
    S[T], v[T] = ST, vT
    nu = B.shape[2]
    Kx = np.zeros((T, nu, nx))
    d = np.zeros((T, nu))
    for k in range(T - 1, -1, -1):
        Sn, vn = S[k + 1], v[k + 1]
        G = Ue[k] + B[k].T @ Sn @ B[k]
        Kx[k] = np.linalg.solve(G, Me[k].T + B[k].T @ Sn @ A[k])
        d[k] = np.linalg.solve(G, -p[k] + B[k].T @ (vn - Sn @ c[k]))
        Acl = A[k] - B[k] @ Kx[k]
        S[k] = Xe[k] + A[k].T @ Sn @ Acl - Me[k] @ Kx[k]
        # v_k: linear term of the closed-loop value function
        v[k] = -q[k] + Acl.T @ (vn - Sn @ c[k]) + Kx[k].T @ p[k]
    return Kx, d, S, v


def _affine_combine(e1, e2):
    F1, c1 = e1
    F2, c2 = e2
    return F2 @ F1, _mv(F2, c1) + c2


def par_fwd_pass(lqt: LQT, x0, Kx, d):
    """Closed-loop affine prefix scan; same pre-applied-first-element pattern as
    ref noc/costates.py:19-31."""
    A, B, c = (np.asarray(a, dtype=np.float64) for a in lqt[:3])
    x0 = np.asarray(x0, dtype=np.float64)
    Ft = A - B @ Kx
    ct = c + _mv(B, d)
    tF = Ft.copy()
    tc = ct.copy()
    tc[0] = Ft[0] @ x0 + ct[0]
    tF[0] = 0.0
    _, xs = associative_scan(_affine_combine, (tF, tc))
    x = np.concatenate([x0[None], xs])
    u = -_mv(Kx, x[:-1]) + d
    return u, x


def seq_fwd_pass(lqt: LQT, x0, Kx, d):
    A, B, c = (np.asarray(a, dtype=np.float64) for a in lqt[:3])
    T, nx = A.shape[0], A.shape[1]
    x = np.zeros((T + 1, nx))
    u = np.zeros((T, B.shape[2]))
    x[0] = x0
    for k in range(T):
        u[k] = -Kx[k] @ x[k] + d[k]
        x[k + 1] = A[k] @ x[k] + B[k] @ u[k] + c[k]
    return u, x

"""NumPy restatement of the reference's par IP-Newton path (and its sequential twin).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the
reference lines it follows.  User functions (dynamics, costs, constraints) are
evaluated through an `Evaluator` (oracle/autodiff.py: torch.func on CPU, float64),
which plays the role JAX autodiff plays in the reference.
"""
from typing import NamedTuple
import numpy as np
from .assoc_scan import associative_scan
from . import paroc_np
from .paroc_np import LQT, _mv, _T


class Derivatives(NamedTuple):  # ref noc/optimal_control_problem.py:13-23
    cx: np.ndarray
    cu: np.ndarray
    cxx: np.ndarray
    cuu: np.ndarray
    cxu: np.ndarray
    fx: np.ndarray
    fu: np.ndarray
    fxx: np.ndarray
    fuu: np.ndarray
    fxu: np.ndarray


# --------------------------------------------------------------------------- costates
def _combine_fc(e1, e2):  # ref noc/costates.py:6-12
    F1, c1 = e1
    F2, c2 = e2
    return F2 @ F1, _mv(F2, c1) + c2


def par_costates(lamda_T, d: Derivatives):
    """ref noc/costates.py:34-40 with par_init (:19-31) and par_scan (:15-16).
    `lamda_T` = grad(final_cost)(x_N) is supplied by the caller (:35)."""
    F = _T(d.fx[::-1])                      # :36
    c = d.cx[::-1]                          # :37
    tF = F.copy()
    tc = c.copy()
    tc[0] = F[0] @ lamda_T + c[0]           # :21
    tF[0] = 0.0                             # :20
    _, cs = associative_scan(_combine_fc, (tF, tc))   # :16
    return np.concatenate([lamda_T[None], cs])[::-1]  # :40


def seq_costates(lamda_T, d: Derivatives):
    """ref noc/costates.py:43-54."""
    N = d.cx.shape[0]
    lam = np.zeros((N + 1, d.cx.shape[1]))
    lam[N] = lamda_T
    for k in range(N - 1, -1, -1):
        lam[k] = d.cx[k] + d.fx[k].T @ lam[k + 1]     # :50
    return lam


# --------------------------------------------------------------------------- LQ assembly
def compute_lqr_params(lam, d: Derivatives):
    """ref noc/par_interior_point_newton.py:31-42 (tensordot contracts the OUTPUT index)."""
    l = lam[1:]
    ru = d.cu + np.einsum("tou,to->tu", d.fu, l)
    Q = d.cxx + np.einsum("to,toij->tij", l, d.fxx)
    R = d.cuu + np.einsum("to,toij->tij", l, d.fuu)
    M = d.cxu + np.einsum("to,toij->tij", l, d.fxu)
    return ru, Q, R, M


def noc_to_lqt(ru, Q, R, M, A, B):
    """ref noc/par_interior_point_newton.py:50-84."""
    T, nx, nu = Q.shape[0], Q.shape[1], R.shape[1]
    X_inv_M = np.linalg.solve(Q, M)                                    # :63
    s = -np.linalg.solve(R - _T(M) @ X_inv_M, ru[..., None])[..., 0]    # :64
    r = -_mv(X_inv_M, s)                                               # :65
    eyex = np.broadcast_to(np.eye(nx), (T, nx, nx)).copy()
    eyeu = np.broadcast_to(np.eye(nu), (T, nu, nu)).copy()
    return LQT(A, B, np.zeros((T, nx)), Q[0], np.eye(nx), np.zeros(nx),
               Q, eyex, r, R, eyeu, s, M)                              # :68-83


def par_Newton(nx, d: Derivatives, reg_param, ru, Q, R, M):
    """ref noc/par_interior_point_newton.py:107-124."""
    grad_cost_norm = np.linalg.norm(d.cu.reshape(-1))                  # :116
    reg = reg_param * grad_cost_norm                                   # :117
    nu = R.shape[1]
    R = R + reg * np.eye(nu)[None]                                     # :118
    lqt = noc_to_lqt(ru, Q, R, M, d.fx, d.fu)                          # :119
    Kx, dd, S, v, pred, feasible = paroc_np.par_bwd_pass(lqt)          # :120
    du, dx = paroc_np.par_fwd_pass(lqt, np.zeros(nx), Kx, dd)          # :121-123
    return dx, du, pred, feasible, ru


# --------------------------------------------------------------------------- sequential twin
def seq_bwd_pass(VxxN, ru, Q, R, M, fx, fu, rp):
    """ref noc/seq_interior_point_newton.py:42-75 (VxxN passed in; :66 takes the final-cost Hessian)."""
    N, nx, nu = Q.shape[0], Q.shape[1], R.shape[1]
    Vxx, Vx = VxxN.copy(), np.zeros(nx)                                # :66-67
    K = np.zeros((N, nu, nx))
    k = np.zeros((N, nu))
    dV = np.zeros(N)
    convex = True
    for t in range(N - 1, -1, -1):
        Qxx = Q[t] + fx[t].T @ Vxx @ fx[t]                             # :49
        Quu = R[t] + fu[t].T @ Vxx @ fu[t] + rp * np.eye(nu)           # :50-51
        convex = convex and bool(np.all(np.linalg.eigvalsh(Quu) > 0))  # :52-53
        Qxu = M[t] + fx[t].T @ Vxx @ fu[t]                             # :54
        Qu = ru[t] + fu[t].T @ Vx                                      # :55
        Qx = fx[t].T @ Vx                                              # :56
        Quu_inv = np.linalg.inv(Quu)
        k[t] = -Quu_inv @ Qu                                           # :58
        K[t] = -Quu_inv @ Qxu.T                                        # :59
        Vx = Qx - Qu @ Quu_inv @ Qxu.T                                 # :61
        Vxx = Qxx - Qxu @ Quu_inv @ Qxu.T                              # :62
        dV[t] = k[t] @ Qu + 0.5 * k[t] @ Quu @ k[t]                    # :63
    return K, k, float(np.sum(dV)), convex


def seq_fwd_pass(K, k, fx, fu):
    """ref noc/seq_interior_point_newton.py:78-90."""
    N, nx = K.shape[0], K.shape[2]
    dx = np.zeros((N + 1, nx))
    for t in range(N):
        dx[t + 1] = (fx[t] + fu[t] @ K[t]) @ dx[t] + fu[t] @ k[t]      # :84
    du = np.einsum("tux,tx->tu", K, dx[:-1]) + k                       # :89
    return du, dx


# --------------------------------------------------------------------------- driver loops
class Trace(NamedTuple):
    stage: int
    iteration: int
    attempt: int
    cost: float
    new_cost: float
    pred: float
    gain_ratio: float
    success: bool
    rp: float
    Hu_norm: float


def newton_oc(ev, controls, initial_state, barrier_param, trace=None, stage=0,
              newton_step=par_Newton, costates=par_costates):
    """ref noc/par_interior_point_newton.py:127-225.  `ev` is an oracle Evaluator."""
    u = np.array(controls, dtype=np.float64)
    x = ev.rollout(u, initial_state)                                   # :133
    nx = x.shape[1]
    reg_param, reg_inc = 1.0, 2.0                                      # :134-135
    iteration, Hu_norm = 0, 1.0
    while not (Hu_norm < 1e-4 or iteration > 1000):                    # :199-202
        cost = ev.total_cost(x, u, barrier_param)                      # :142
        d = ev.derivatives(x, u, barrier_param)                        # :145
        lam = costates(ev.final_cost_grad(x[-1]), d)                   # :147
        ru, Q, R, M = compute_lqr_params(lam, d)                       # :149
        success, inner = False, 0
        rp, r_inc = reg_param, reg_inc
        tx, tu = x, u
        while not (success or inner > 500):                            # :177-182
            dx, du, pred, bwd_feasible, Hu = newton_step(nx, d, rp, ru, Q, R, M)  # :153
            tu = u + du                                                # :156
            tx = x + dx                                                # :157
            Hu_norm = float(np.max(np.abs(Hu)))                        # :158
            if ev.feasible(tx, tu):                                    # :159-163
                new_cost = ev.total_cost(tx, tu, barrier_param)
            else:
                new_cost = np.inf
            actual = new_cost - cost                                   # :164
            with np.errstate(all="ignore"):
                gain_ratio = np.float64(actual) / np.float64(pred)     # :165
            success = bool(gain_ratio > 0.0) and bool(bwd_feasible)    # :166
            rp_before = rp
            if success:                                                # :167-172
                rp = rp * max(1.0 / 3.0, 1.0 - (2.0 * gain_ratio - 1.0) ** 3)
                r_inc = 2.0
            else:
                rp = rp * r_inc
                r_inc = 2 * r_inc
            rp = float(np.clip(rp, 1e-16, 1e16))                       # :173
            inner += 1                                                 # :174
            if trace is not None:
                trace.append(Trace(stage, iteration, inner, float(cost), float(new_cost),
                                   float(pred), float(gain_ratio), success, rp_before, Hu_norm))
        x, u = tx, tu                                                  # :184 (taken even if never successful)
        reg_param, reg_inc = rp, r_inc
        iteration += 1                                                 # :194
    return x, u, iteration


def par_interior_point_optimal_control(ev, controls, initial_state, trace=None, **kw):
    """ref noc/par_interior_point_newton.py:228-254."""
    u = np.array(controls, dtype=np.float64)
    bp, total, stage = 0.1, 0, 0                                       # :233
    while bp > 1e-4:                                                   # :243-245
        _, u, its = newton_oc(ev, u, initial_state, bp, trace, stage, **kw)   # :237
        bp = bp / 5                                                    # :238
        total += its                                                   # :239
        stage += 1
    return u, total


def constrained_mpc(ev, x0, horizon=40, sim_steps=20, u_init=None):
    """CPU twin of ipoc_b200.mpc.constrained_mpc (BASELINE config 3, box-constrained extension): the loop of
    ref examples/linear_mpc_parallel.py:67-81 around par_interior_point_optimal_control
    (ref noc/par_interior_point_newton.py:228-254) with a shifted warm start.  `ev` = Evaluator of the OCP."""
    import torch
    x = np.asarray(x0, dtype=np.float64)
    u = 0.1 * np.random.default_rng(1).standard_normal((horizon, 1)) if u_init is None else np.array(u_init, dtype=np.float64)
    xs, us, its = [x.copy()], [], []
    for _ in range(sim_steps):
        u_opt, n_it = par_interior_point_optimal_control(ev, u, x)
        x = ev.ocp.dynamics(torch.as_tensor(x), torch.as_tensor(u_opt[0])).detach().numpy().astype(np.float64)
        xs.append(x.copy())
        us.append(u_opt[0].copy())
        its.append(int(n_it))
        u = np.concatenate((u_opt[1:], u_opt[-1:]))
    return np.stack(xs), np.stack(us), its

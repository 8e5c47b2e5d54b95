"""Host-side mirror of the reference's par IP-Newton interface
(ref noc/par_interior_point_newton.py, noc/costates.py): same function names, argument meaning
and return values, with torch CUDA tensors in place of jnp arrays.

What runs where
  * user functions (dynamics, costs, constraints) and their derivatives: the host framework
    (`torch.func` vmapped autodiff on the GPU) — as in the reference, where JAX does it;
  * everything the reference does with `lax.associative_scan` (costates, Riccati scan + gains,
    forward scan) plus the accept/reject reductions: the sm_100a kernels behind include/ipoc.h.
There is no CPU fallback; CPU tensors raise.
"""
import ctypes
import torch
from torch.func import vmap, grad, hessian, jacrev
from . import _lib as L
from .optimal_control_problem import OCP, Derivatives
from .paroc import LQT, par_bwd_pass, par_fwd_pass  # noqa: F401  (re-exported, as the reference imports them)
from .utils import rollout, rollout_parallel
from . import plants


# ------------------------------------------------------------------ A1: derivatives (host framework)
def compute_derivatives(ocp: OCP, states: torch.Tensor, controls: torch.Tensor, bp: float) -> Derivatives:
    """ref noc/par_interior_point_newton.py:13-28.  Index order (output, wrt_1, wrt_2)."""
    def body(x, u):
        cx_k, cu_k = grad(ocp.stage_cost, (0, 1))(x, u, bp)
        cxx_k = hessian(ocp.stage_cost, 0)(x, u, bp)
        cuu_k = hessian(ocp.stage_cost, 1)(x, u, bp)
        cxu_k = jacrev(jacrev(ocp.stage_cost, 0), 1)(x, u, bp)
        fx_k, fu_k = jacrev(ocp.dynamics, (0, 1))(x, u)
        fxx_k = jacrev(jacrev(ocp.dynamics, 0), 0)(x, u)
        fuu_k = jacrev(jacrev(ocp.dynamics, 1), 1)(x, u)
        fxu_k = jacrev(jacrev(ocp.dynamics, 0), 1)(x, u)
        return cx_k, cu_k, cxx_k, cuu_k, cxu_k, fx_k, fu_k, fxx_k, fuu_k, fxu_k

    return Derivatives(*(t.contiguous() for t in vmap(body)(states[:-1], controls)))


# ------------------------------------------------------------------ A3: LQ parameters (host framework)
def compute_lqr_params(lagrange_multipliers: torch.Tensor, d: Derivatives):
    """ref noc/par_interior_point_newton.py:31-42 (tensordot contracts the OUTPUT index).
    CUDA tensors go through one streaming kernel (ipoc_lqr_params_f64); the einsum form below is the
    same arithmetic for anything else.  Accepts (N, ...) or (B, N, ...)."""
    if lagrange_multipliers.is_cuda and lagrange_multipliers.shape[-1] <= 8:   # kernel: nx <= 8; else the einsums
        lam = L.dev_f64(lagrange_multipliers)
        t = [L.dev_f64(a) for a in (d.cu, d.cxx, d.cuu, d.cxu, d.fu, d.fxx, d.fuu, d.fxu)]
        batched = lam.dim() == 3
        if not batched:
            lam = lam.unsqueeze(0)
            t = [a.unsqueeze(0) for a in t]
        Bn, N, nu = t[0].shape[0], t[0].shape[1], t[0].shape[2]
        nx = lam.shape[-1]
        o = dict(dtype=torch.float64, device=lam.device)
        ru, Q = torch.empty(Bn, N, nu, **o), torch.empty(Bn, N, nx, nx, **o)
        R, M = torch.empty(Bn, N, nu, nu, **o), torch.empty(Bn, N, nx, nu, **o)
        with torch.cuda.device(lam.device):
            L.check(L.lib().ipoc_lqr_params_f64(N, nx, nu, Bn, L.ptr(lam), *(L.ptr(a) for a in t), L.ptr(ru), L.ptr(Q),
                                                L.ptr(R), L.ptr(M), L.stream_ptr()))
        return (ru, Q, R, M) if batched else (ru[0], Q[0], R[0], M[0])
    l = lagrange_multipliers[1:]
    ru = d.cu + torch.einsum("tou,to->tu", d.fu, l)
    Q = d.cxx + torch.einsum("to,toij->tij", l, d.fxx)
    R = d.cuu + torch.einsum("to,toij->tij", l, d.fuu)
    M = d.cxu + torch.einsum("to,toij->tij", l, d.fxu)
    return ru.contiguous(), Q.contiguous(), R.contiguous(), M.contiguous()


def check_traj_feasibility(ocp: OCP, x: torch.Tensor, u: torch.Tensor):
    """ref noc/par_interior_point_newton.py:45-47 — `all(cons <= 0)` reduced by K4."""
    cons = vmap(ocp.constraints)(x[:-1], u)
    cons = L.dev_f64(cons.reshape(cons.shape[0], -1))
    _, _, feas = reductions(cons=cons)
    return feas[0] != 0


# ------------------------------------------------------------------ K1: costates
def affine_scan(F, c, seed, reverse=False, transpose=False):
    """out[0]=seed, out[k+1]=F_k out[k]+c_k  (reverse: out[N]=seed, out[k]=F_k(') out[k+1]+c_k)."""
    F, c = L.dev_f64(F), L.dev_f64(c)
    seed = L.dev_f64(seed, F.device)
    batched = F.dim() == 4
    if not batched:
        F, c, seed = F.unsqueeze(0), c.unsqueeze(0), seed.unsqueeze(0)
    Bn, N, nx = F.shape[0], F.shape[1], F.shape[2]
    out = torch.empty(Bn, N + 1, nx, dtype=torch.float64, device=F.device)
    ws, nbytes = L.workspace(L.WS_AFFINE_SCAN, N, nx, 1, Bn, F.device)
    with torch.cuda.device(F.device):
        L.check(L.lib().ipoc_affine_scan_f64(int(reverse), int(transpose), N, nx, Bn, L.ptr(F), L.ptr(c),
                                             L.ptr(seed.contiguous()), L.ptr(out), L.ptr(ws), nbytes, L.stream_ptr()))
    return out if batched else out[0]


def par_costates(ocp: OCP, final_state: torch.Tensor, d: Derivatives):
    """ref noc/costates.py:34-40: lambda_N = grad final_cost(x_N), lambda_k = cx_k + fx_k' lambda_{k+1}."""
    lamda_T = grad(ocp.final_cost)(final_state)
    return affine_scan(d.fx, d.cx, lamda_T, reverse=True, transpose=True)


# ------------------------------------------------------------------ A4 (API completeness)
def noc_to_lqt(ru, Q, R, M, A, B) -> LQT:
    """ref noc/par_interior_point_newton.py:50-84.  `par_Newton` below does NOT call this (the fused
    kernel forms r, s per step in registers and never materialises the identity stacks); it exists so
    that code written against the reference's helper keeps working."""
    T, nx, nu = Q.shape[0], Q.shape[1], R.shape[1]
    X_inv_M = torch.linalg.solve(Q, M)
    s = -torch.linalg.solve(R - M.transpose(1, 2) @ X_inv_M, ru.unsqueeze(-1)).squeeze(-1)
    r = -(X_inv_M @ s.unsqueeze(-1)).squeeze(-1)
    o = dict(dtype=Q.dtype, device=Q.device)
    eye = lambda n: torch.eye(n, **o).expand(T, n, n).contiguous()
    return LQT(A, B, torch.zeros(T, nx, **o), Q[0], torch.eye(nx, **o), torch.zeros(nx, **o),
               Q, eye(nx), r, R, eye(nu), s, M)


# ------------------------------------------------------------------ K4 / A8
_ones_cache = {}


def _ones_i32(n, dev):
    """Read-only all-ones int32 vector (the `feasible` result when no constraints were passed), cached."""
    key = (n, str(dev))
    t = _ones_cache.get(key)
    if t is None:
        t = _ones_cache[key] = torch.ones(n, dtype=torch.int32, device=dev)
    return t


def reductions(ru=None, cu=None, cons=None, rp=None, reg_out=None, hu_out=None):
    """(max|ru|, ||cu||_F, all(cons<=0)) per problem — ref :158, :116, :45-47.  Inputs (N,·) or (B,N,·).
    With `rp` and `reg_out` (device, one per problem) also writes reg_out = rp * ||cu||_F (:117)."""
    ref = next(t for t in (ru, cu, cons) if t is not None)
    dev = ref.device
    batched = ref.dim() == 3
    prep = lambda t: None if t is None else (L.dev_f64(t) if batched else L.dev_f64(t).unsqueeze(0))
    ru, cu, cons = prep(ru), prep(cu), prep(cons)
    ref = next(t for t in (ru, cu, cons) if t is not None)
    Bn, N = ref.shape[0], ref.shape[1]
    nu = (ru if ru is not None else cu).shape[2] if (ru is not None or cu is not None) else 1
    nc = cons.shape[2] if cons is not None else 1
    # outputs the kernel computes are left uninitialised (every fill is a launch on the critical path of a solve);
    # the others get their neutral values
    o = dict(dtype=torch.float64, device=dev)
    if hu_out is not None:                      # caller-owned result buffer (device-resident loops)
        hu = hu_out
    else:
        hu = torch.empty(Bn, **o) if ru is not None else torch.zeros(Bn, **o)
    cn = torch.empty(Bn, **o) if cu is not None else torch.zeros(Bn, **o)
    fe = torch.empty(Bn, dtype=torch.int32, device=dev) if cons is not None else _ones_i32(Bn, dev)
    ws, nbytes = L.workspace(L.WS_REDUCTIONS, N, max(nu, nc), nu, Bn, dev)
    with torch.cuda.device(dev):
        L.check(L.lib().ipoc_reductions_f64(N, nu, nc, Bn, L.ptr(ru), L.ptr(cu), L.ptr(cons), L.ptr(hu), L.ptr(cn),
                                            L.ptr(fe), L.ptr(rp), L.ptr(reg_out), L.ptr(ws), nbytes, L.stream_ptr()))
    return hu, cn, fe


def accept_update(cost, new_cost, traj_feasible, pred, bwd_feasible, rp, r_inc, active=None):
    """In-place device update of (rp, r_inc); returns (success int32, gain_ratio) — ref :159-173."""
    Bn = rp.numel()
    dev = rp.device
    success = torch.zeros(Bn, dtype=torch.int32, device=dev)
    gain = torch.zeros(Bn, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().ipoc_accept_update_f64(Bn, L.ptr(cost), L.ptr(new_cost), L.ptr(traj_feasible), L.ptr(pred),
                                               L.ptr(bwd_feasible), L.ptr(active), L.ptr(rp), L.ptr(r_inc),
                                               L.ptr(success), L.ptr(gain), L.stream_ptr()))
    return success, gain


# ------------------------------------------------------------------ attempt-loop glue (one launch each)
def attempt_begin(done, rp, cu_norm, active, reg):
    """active <- !done (int32), reg <- rp * ||cu|| (ref :117) for a batch whose attempt loops live on the device."""
    with torch.cuda.device(rp.device):
        L.check(L.lib().ipoc_attempt_begin_f64(rp.numel(), L.ptr(done), L.ptr(rp), L.ptr(cu_norm), L.ptr(active),
                                               L.ptr(reg), L.stream_ptr()))


def trial_point(x, dx, u, du, tx, tu):
    """tx <- x + dx, tu <- u + du (ref :156-157) into preallocated buffers; (B,N+1,nx) / (B,N,nu) tensors."""
    Bn, N, nx, nu = u.shape[0], u.shape[1], x.shape[-1], u.shape[-1]
    with torch.cuda.device(x.device):
        L.check(L.lib().ipoc_trial_point_f64(N, nx, nu, Bn, L.ptr(x), L.ptr(dx), L.ptr(u), L.ptr(du), L.ptr(tx),
                                             L.ptr(tu), L.stream_ptr()))


def attempt_commit(active, success, tx, tu, keep_x, keep_u, inner, done, max_attempts=500):
    """For active members: keep the trial point (ref :175), inner += 1 (:174), done |= success or inner >
    max_attempts (:180-181).  `done` is a torch.bool tensor, `inner` int64."""
    Bn, N, nx, nu = tu.shape[0], tu.shape[1], tx.shape[-1], tu.shape[-1]
    with torch.cuda.device(tx.device):
        L.check(L.lib().ipoc_attempt_commit_f64(N, nx, nu, Bn, L.ptr(active), L.ptr(success), L.ptr(tx), L.ptr(tu),
                                                L.ptr(keep_x), L.ptr(keep_u), L.ptr(inner), L.ptr(done),
                                                int(max_attempts), L.stream_ptr()))


def newton_advance(hu, inner_done, outer_done, inner, iteration, advanced, tx, tu, x, u, hu_tol=1e-4,
                   max_iterations=1000):
    """End-of-iteration bookkeeping of `newton_oc` on the device (ref :184-202): members whose attempt loop has
    ended take the step (x <- tx, u <- tu), count the iteration and test the exit condition."""
    Bn = hu.numel()
    N, nx, nu = tu.shape[-2], tx.shape[-1], tu.shape[-1]
    with torch.cuda.device(tx.device):
        L.check(L.lib().ipoc_newton_advance_f64(N, nx, nu, Bn, L.ptr(hu), L.ptr(inner_done), L.ptr(outer_done),
                                                L.ptr(inner), L.ptr(iteration), L.ptr(advanced), L.ptr(tx), L.ptr(tu),
                                                L.ptr(x), L.ptr(u), float(hu_tol), int(max_iterations),
                                                L.stream_ptr()))


def costates_fused(fx, cx, lamT, cu, fresh=None, out=None):
    """K1 with ||cu||_F folded into the up-sweep (ipoc_costates_f64; ref noc/costates.py:34-40 + :116 of the
    Newton step) -> (lam (B,N+1,nx), cu_norm (B,)).  Batched (B,N,...) tensors.  `fresh` (int32 per problem): members
    whose flag is 0 are skipped and keep what `out` = (lam, cu_norm) of an earlier call holds."""
    fx, cx, cu = L.dev_f64(fx), L.dev_f64(cx), L.dev_f64(cu)
    lamT = L.dev_f64(lamT, fx.device).reshape(fx.shape[0], -1).contiguous()
    Bn, N, nx, nu = fx.shape[0], fx.shape[1], fx.shape[2], cu.shape[-1]
    if out is not None:
        lam, cu_norm = out
    else:
        lam = torch.empty(Bn, N + 1, nx, dtype=torch.float64, device=fx.device)
        cu_norm = torch.empty(Bn, dtype=torch.float64, device=fx.device)
    ws, nbytes = L.workspace(L.WS_COSTATES, N, nx, nu, Bn, fx.device)
    with torch.cuda.device(fx.device):
        L.check(L.lib().ipoc_costates_f64(N, nx, nu, Bn, L.ptr(fx), L.ptr(cx), L.ptr(lamT), L.ptr(cu), L.ptr(lam),
                                          L.ptr(cu_norm), L.ptr(fresh), L.ptr(ws), nbytes, L.stream_ptr()))
    return lam, cu_norm


class AttemptBuffers:
    """Result buffers + private workspace of `newton_attempt` (allocated once per loop object)."""

    def __init__(self, B, N, nx, nu, dev):
        o = dict(dtype=torch.float64, device=dev)
        self.dx, self.du = torch.empty(B, N + 1, nx, **o), torch.empty(B, N, nu, **o)
        self.Kx, self.d = torch.empty(B, N, nu, nx, **o), torch.empty(B, N, nu, **o)
        self.pred, self.hu = torch.empty(B, **o), torch.ones(B, **o)
        self.bwd_feas = torch.empty(B, dtype=torch.int32, device=dev)
        self.nbytes = L.lib().ipoc_workspace_bytes(L.WS_NEWTON_ATTEMPT, N, nx, nu, B)
        if self.nbytes == 0:
            raise L.IpocError(f"unsupported (nx={nx}, nu={nu}): no kernel instantiated and there is no CPU fallback")
        self.ws = torch.zeros(self.nbytes, dtype=torch.uint8, device=dev)   # control block must start zeroed


def newton_attempt(buf: AttemptBuffers, fx, fu, ru, Q, R, M, rp, cu_norm, x=None, u=None, tx=None, tu=None, active=None):
    """ipoc_newton_attempt_f64: K2 + K3 with reg = rp*||cu|| formed in the kernel (ref :117), max|ru| (:158) folded
    into the up-sweep and the trial point tx = x + dx, tu = u + du (:156-157; only for members with active != 0)
    written by K3's leaf kernel.  Three launches.  Results land in `buf`."""
    Bn, N, nx, nu = fx.shape[0], fx.shape[1], fx.shape[2], fu.shape[-1]
    p = L.ptr
    with torch.cuda.device(fx.device):
        L.check(L.lib().ipoc_newton_attempt_f64(
            N, nx, nu, 1, Bn, p(fx), p(fu), p(ru), p(Q), p(R), p(M), p(rp), p(cu_norm), p(buf.dx), p(buf.du), p(buf.Kx),
            p(buf.d), p(buf.pred), p(buf.bwd_feas), p(buf.hu), p(x), p(u), p(tx), p(tu), None, None, None, None, None,
            p(active), None, None, None, None, p(buf.ws), buf.nbytes, L.stream_ptr()))
    return buf


# ------------------------------------------------------------------ K2 + K3: the Newton step
def newton_step(fx, fu, ru, Q, R, M, reg):
    """Fused device Newton step on LQ data; `reg` is a device tensor (one value per problem).
    -> dx, du, Kx, d, pred (B,), feasible (B,) int32.  Accepts (N,...) or (B,N,...)."""
    fx, fu, ru, Q, R, M = (L.dev_f64(t) for t in (fx, fu, ru, Q, R, M))
    dev = fx.device
    batched = fx.dim() == 4
    if not batched:
        fx, fu, ru, Q, R, M = (t.unsqueeze(0) for t in (fx, fu, ru, Q, R, M))
    Bn, N, nx = fx.shape[0], fx.shape[1], fx.shape[2]
    nu = fu.shape[-1]
    L.require_supported(nx, nu)
    reg = L.dev_f64(reg, dev).reshape(-1)
    if reg.numel() != Bn:
        reg = reg.expand(Bn).contiguous()
    o = dict(dtype=torch.float64, device=dev)
    dx = torch.empty(Bn, N + 1, nx, **o)
    du = torch.empty(Bn, N, nu, **o)
    Kx = torch.empty(Bn, N, nu, nx, **o)
    d = torch.empty(Bn, N, nu, **o)
    pred = torch.empty(Bn, **o)
    feas = torch.empty(Bn, dtype=torch.int32, device=dev)
    ws, nbytes = L.workspace(L.WS_NEWTON_STEP, N, nx, nu, Bn, dev)
    with torch.cuda.device(dev):
        L.check(L.lib().ipoc_newton_step_f64(N, nx, nu, Bn, L.ptr(fx), L.ptr(fu), L.ptr(ru), L.ptr(Q), L.ptr(R),
                                             L.ptr(M), L.ptr(reg), L.ptr(dx), L.ptr(du), L.ptr(Kx), L.ptr(d),
                                             L.ptr(pred), L.ptr(feas), L.ptr(ws), nbytes, L.stream_ptr()))
    if not batched:
        return dx[0], du[0], Kx[0], d[0], pred, feas
    return dx, du, Kx, d, pred, feas


def par_Newton(nominal_states, d: Derivatives, reg_param, ru, Q, R, M):
    """ref noc/par_interior_point_newton.py:107-124 -> (dx, du, pred_reduction, feasible, ru).
    `nominal_states` is only used for its shape in the reference (:122); the initial deviation is 0."""
    _, cu_norm, _ = reductions(cu=d.cu)                      # :116
    reg = torch.as_tensor(reg_param, dtype=torch.float64, device=cu_norm.device) * cu_norm   # :117
    dx, du, _, _, pred, feas = newton_step(d.fx, d.fu, ru, Q, R, M, reg)    # :118-123
    return dx, du, pred[0], feas[0] != 0, ru


# ------------------------------------------------------------------ user-function evaluation of one iteration / trial
def eval_iteration(ocp: OCP, x, u, bp):
    """Everything one Newton iteration needs from the user functions at the iterate (x, u):
    -> cost (1,), fx, fu, cu, ru, Q, R, M   (ref :142, :145, :147, :149; K1 runs inside).
    General OCPs: host-framework autodiff (`compute_derivatives`), costate scan, `compute_lqr_params`.
    Built-in plants (plants.py): first-order pass -> costate scan -> Hamiltonian pass; the full
    `Derivatives` record is never materialised.  Works for (N, ...) and (B, N, ...) iterates."""
    plant = plants.plant_of(ocp)
    if plant is not None:
        fx, fu, cx, cu, lamT = plants.linearize(plant, x, u, bp)
        cost, _ = plants.cost(plant, x, u, bp)
        lam = affine_scan(fx, cx, lamT, reverse=True, transpose=True)   # :147
        ru, Q, R, M = plants.hamiltonian(plant, x, u, lam, bp)          # :149
        return cost, fx, fu, cu, ru, Q, R, M
    if x.dim() == 3:
        from .batched import compute_derivatives_batched
        cost = vmap(ocp.total_cost, in_dims=(0, 0, None))(x, u, bp)
        d = compute_derivatives_batched(ocp, x, u, bp)
        lamT = vmap(grad(ocp.final_cost))(x[:, -1])
    else:
        cost = ocp.total_cost(x, u, bp).reshape(1)
        d = compute_derivatives(ocp, x, u, bp)
        lamT = grad(ocp.final_cost)(x[-1])
    lam = affine_scan(d.fx, d.cx, lamT, reverse=True, transpose=True)   # :147
    ru, Q, R, M = compute_lqr_params(lam, d)                            # :149
    return cost, d.fx, d.fu, d.cu, ru, Q, R, M


def eval_trial(ocp: OCP, tx, tu, bp):
    """new cost (1,) and trajectory feasibility (1,) int32 of a trial iterate — ref :159-163."""
    plant = plants.plant_of(ocp)
    if plant is not None:
        return plants.cost(plant, tx, tu, bp)
    if tx.dim() == 3:
        cons = vmap(vmap(ocp.constraints))(tx[:, :-1], tu)
        _, _, traj_feas = reductions(cons=cons.reshape(cons.shape[0], cons.shape[1], -1))
        return vmap(ocp.total_cost, in_dims=(0, 0, None))(tx, tu, bp), traj_feas
    cons = vmap(ocp.constraints)(tx[:-1], tu)                            # :160
    _, _, traj_feas = reductions(cons=cons.reshape(cons.shape[0], -1))
    return ocp.total_cost(tx, tu, bp).reshape(1), traj_feas              # :161 (masked to inf by A8)


PLANT_PARALLEL_ROLLOUT_FROM = 2048   # serial kernel: ~0.4 us per step; parallel: a handful of ~60 us iterations


def initial_rollout(ocp: OCP, u, initial_state, parallel_rollout_from=256, x_guess=None):
    """states of the nominal rollout (ref :133).  Built-in plants: the serial device kernel for short horizons, the
    parallel-in-time Newton-on-the-rollout iteration on the plant's own kernels for long ones (`x_guess`, e.g. the
    trajectory the previous barrier stage ended with, saves iterations).  User OCPs: the same iteration through
    the host framework's autodiff for long horizons, else the serial host loop."""
    plant = plants.plant_of(ocp)
    N = u.shape[0]
    if plant is not None:
        if N >= PLANT_PARALLEL_ROLLOUT_FROM:
            return plants.rollout_parallel(plant, u, initial_state.to(u.device), x_guess)[0]
        return plants.rollout(plant, u, initial_state.to(u.device))
    if N >= parallel_rollout_from:
        return rollout_parallel(ocp.dynamics, u, initial_state.to(u.device), x_guess)[0]
    return rollout(ocp.dynamics, u, initial_state.to(u.device))


# ------------------------------------------------------------------ driver loops
def newton_oc(ocp: OCP, controls: torch.Tensor, initial_state: torch.Tensor, barrier_param: float, trace=None,
              stage: int = 0, use_graphs: bool = True, parallel_rollout_from: int = 256, x_guess=None):
    """ref noc/par_interior_point_newton.py:127-225 -> (opt_x, opt_u, iterations).
    use_graphs=True (default): for the built-in plants the whole loop runs from one CUDA graph with its
    control flow on the device (graphed.DeviceLoopNewton; the host only watches the exit flag; "device"
    forces this for any OCP).  For user OCPs, with a `trace`, or with use_graphs="host", the two loop bodies
    are separate CUDA graphs and the host reads one small record per attempt to steer the loop
    (graphed.GraphedNewton).  use_graphs=False: the eager path below, the same sequence of statements."""
    dev = controls.device
    u = L.dev_f64(controls)
    x = initial_rollout(ocp, u, initial_state, parallel_rollout_from, x_guess)   # :133
    if use_graphs:
        from . import graphed
        # Whole loop on the device (no per-attempt host read) where re-evaluating the iterate after a rejected
        # attempt is cheap — the built-in plants (7 kernels).  With host-framework autodiff that evaluation is
        # hundreds of kernels, so user OCPs keep the host-steered graphs, which evaluate once per iteration
        # (measured: cartpole N=1e4 through autodiff 0.41 s host-steered vs 0.52 s device-resident).
        if trace is None and use_graphs != "host" and (plants.plant_of(ocp) is not None or use_graphs == "device"):
            loop = graphed.get_device_loop(ocp, u.shape[0], x.shape[1], u.shape[1], dev, x, u, barrier_param)
            if loop:
                return loop.run(x, u, barrier_param)
        g = graphed.get(ocp, u.shape[0], x.shape[1], u.shape[1], dev, x, u, barrier_param)
        if g:
            return _newton_oc_graphed(g, x, u, barrier_param, trace, stage)
    o = dict(dtype=torch.float64, device=dev)
    rp = torch.ones(1, **o)                                              # :134
    r_inc = torch.full((1,), 2.0, **o)                                   # :135
    iteration, Hu_norm = 0, 1.0
    while not (Hu_norm < 1e-4 or iteration > 1000):                      # :199-202
        cost, fx, fu, cu, ru, Q, R, M = eval_iteration(ocp, x, u, barrier_param)   # :142-149
        hu, cu_norm, _ = reductions(ru=ru, cu=cu)                        # :158 (pre-step ru), :116
        success, inner = False, 0
        tx, tu = x, u
        while not (success or inner > 500):                              # :177-182
            dx, du, _, _, pred, bwd_feas = newton_step(fx, fu, ru, Q, R, M, rp * cu_norm)   # :153
            tu = u + du                                                  # :156
            tx = x + dx                                                  # :157
            new_cost, traj_feas = eval_trial(ocp, tx, tu, barrier_param)  # :159-163
            rp_before = rp.clone() if trace is not None else None
            succ, gain = accept_update(cost, new_cost, traj_feas, pred, bwd_feas, rp, r_inc)    # :159-173
            inner += 1                                                   # :174
            rec = torch.stack((succ[0].to(torch.float64), hu[0])).cpu()  # the one host read per attempt
            success, Hu_norm = bool(rec[0] != 0), float(rec[1])
            if trace is not None:
                trace.append(dict(stage=stage, iteration=iteration, attempt=inner, cost=float(cost),
                                  new_cost=float(new_cost) if bool(traj_feas[0]) else float("inf"),
                                  pred=float(pred), gain_ratio=float(gain), success=success,
                                  rp=float(rp_before), Hu_norm=Hu_norm))
        x, u = tx, tu                                                    # :184 (taken even if never successful)
        iteration += 1                                                   # :194
    return x, u, iteration


def _newton_oc_graphed(g, x, u, barrier_param, trace, stage):
    """Same loop as above with the two bodies replayed from CUDA graphs."""
    g.x.copy_(x)
    g.u.copy_(u)
    g.bp.fill_(float(barrier_param))
    g.rp.fill_(1.0)                                                      # :134
    g.r_inc.fill_(2.0)                                                   # :135
    iteration, Hu_norm = 0, 1.0
    while not (Hu_norm < 1e-4 or iteration > 1000):                      # :199-202
        g.iteration()                                                    # :142-149
        success, inner = False, 0
        while not (success or inner > 500):                              # :177-182
            rp_before = float(g.rp) if trace is not None else None
            success, Hu_norm, new_cost, pred, gain = g.attempt()         # :153-173
            inner += 1                                                   # :174
            if trace is not None:
                trace.append(dict(stage=stage, iteration=iteration, attempt=inner, cost=float(g.cost),
                                  new_cost=new_cost, pred=pred, gain_ratio=gain, success=success,
                                  rp=rp_before, Hu_norm=Hu_norm))
        g.x.copy_(g.tx)                                                  # :184 (taken even if never successful)
        g.u.copy_(g.tu)
        iteration += 1                                                   # :194
    return g.x.clone(), g.u.clone(), iteration


def par_interior_point_optimal_control(ocp: OCP, controls: torch.Tensor, initial_state: torch.Tensor, trace=None,
                                       use_graphs: bool = True):
    """ref noc/par_interior_point_newton.py:228-254 -> (opt_u, N_iterations)."""
    if not controls.is_cuda:
        raise L.IpocError("par_interior_point_optimal_control needs CUDA tensors; there is no CPU fallback")
    u = controls
    bp, total, stage = 0.1, 0, 0                                         # :233
    x = None   # the reference keeps only u between stages (:237); the last trajectory merely seeds the parallel rollout
    while bp > 1e-4:                                                     # :243-245
        x, u, its = newton_oc(ocp, u, initial_state, bp, trace, stage, use_graphs, x_guess=x)   # :237
        bp = bp / 5                                                      # :238
        total += its                                                     # :239
        stage += 1
    return u, total

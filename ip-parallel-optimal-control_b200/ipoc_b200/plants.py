"""Optional fast path for the reference's two example plants (SURVEY.md §8f "next" #4).

OCPs built by `problems.make_pendulum` / `problems.make_cartpole` carry a tag on their dynamics callable;
for those the per-step `Derivatives`, the total cost / feasibility of a trajectory and the serial rollout
are single fused kernels of libipoc.so (csrc/ipoc_plants.cu: second-order forward-mode autodiff in
registers) instead of a few hundred host-framework kernels.  User-defined OCPs are untouched and keep going
through `torch.func`.  Set `ENABLED = False` to force the autodiff path everywhere (the tests run both and
compare them).
"""
import ctypes
import torch
from . import _lib as L
from .optimal_control_problem import Derivatives

ENABLED = True
PLANT_IDS = {"pendulum": 1, "cartpole": 2}


def register(ocp, name, Ts, bound):
    """Mark an OCP built by `problems.make_pendulum` / `make_cartpole` as a built-in plant.  The kernels of
    ipoc_plants.cu hard-code the dynamics, stage cost, final cost, goal state, weights AND box constraints, so the
    descriptor records the identity of ALL five callables: an OCP in which any of them was replaced
    (`ocp._replace(stage_cost=my_cost)`, `OCP(make_cartpole(Ts).dynamics, my_cons, ...)`) no longer matches and
    goes through the host framework's autodiff like every user-defined problem."""
    ocp.dynamics._ipoc_plant = {"name": name, "id": PLANT_IDS[name], "Ts": float(Ts), "bound": float(bound),
                                "callables": tuple(id(f) for f in ocp)}
    return ocp


def plant_of(ocp):
    """Plant descriptor of an UNMODIFIED built-in OCP, or None (also None when the fast path is disabled)."""
    if not ENABLED:
        return None
    d = getattr(ocp.dynamics, "_ipoc_plant", None)
    if d is None or d["callables"] != tuple(id(f) for f in ocp):
        return None
    return d


def _bp_tensor(bp, dev):
    if isinstance(bp, torch.Tensor):
        return bp.to(device=dev, dtype=torch.float64).reshape(1).contiguous()
    return torch.tensor([float(bp)], dtype=torch.float64, device=dev)


def _dims(plant):
    nx, nu, nc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    L.check(L.lib().ipoc_plant_dims(plant["id"], ctypes.byref(nx), ctypes.byref(nu), ctypes.byref(nc)))
    return nx.value, nu.value, nc.value


def derivatives(plant, states, controls, bp):
    """-> (Derivatives, lamT): same tensors (index order (output, wrt_1, wrt_2)) as
    `noc.compute_derivatives` + grad final_cost(x_N).  states (N+1,nx) / controls (N,nu), or batched."""
    x, u = L.dev_f64(states), L.dev_f64(controls)
    batched = x.dim() == 3
    if not batched:
        x, u = x.unsqueeze(0), u.unsqueeze(0)
    B, N = u.shape[0], u.shape[1]
    nx, nu, _ = _dims(plant)
    dev = x.device
    o = dict(dtype=torch.float64, device=dev)
    shapes = dict(cx=(nx,), cu=(nu,), cxx=(nx, nx), cuu=(nu, nu), cxu=(nx, nu), fx=(nx, nx), fu=(nx, nu),
                  fxx=(nx, nx, nx), fuu=(nx, nu, nu), fxu=(nx, nx, nu))
    outs = {k: torch.empty((B, N) + sh, **o) for k, sh in shapes.items()}
    lamT = torch.empty(B, nx, **o)
    bpt = _bp_tensor(bp, dev)
    with torch.cuda.device(dev):
        L.check(L.lib().ipoc_plant_derivatives_f64(
            plant["id"], N, B, plant["Ts"], plant["bound"], L.ptr(bpt), L.ptr(x), L.ptr(u),
            *(L.ptr(outs[k]) for k in Derivatives._fields), L.ptr(lamT), L.stream_ptr()))
    if not batched:
        return Derivatives(*(outs[k][0] for k in Derivatives._fields)), lamT[0]
    return Derivatives(*(outs[k] for k in Derivatives._fields)), lamT


def _prep(states, controls):
    x, u = L.dev_f64(states), L.dev_f64(controls)
    batched = x.dim() == 3
    if not batched:
        x, u = x.unsqueeze(0), u.unsqueeze(0)
    return x, u, batched


def linearize(plant, states, controls, bp, fresh=None, out=None, take=None):
    """First-order pass (before the costate scan) -> fx, fu, cx, cu, lamT.  `fresh` (int32 per problem): members
    whose flag is 0 are skipped and keep what `out` (the tuple returned by an earlier call) holds.  `take` =
    (tx, tu): the evaluated members first take their iterate from there (states, controls are then written: the
    masked copy of an accepted step in the same launch, ipoc_plant_take_linearize_f64)."""
    x, u, batched = _prep(states, controls)
    B, N = u.shape[0], u.shape[1]
    nx, nu, _ = _dims(plant)
    o = dict(dtype=torch.float64, device=x.device)
    if out is not None:
        fx, fu, cx, cu, lamT = out
    else:
        fx, fu = torch.empty(B, N, nx, nx, **o), torch.empty(B, N, nx, nu, **o)
        cx, cu, lamT = torch.empty(B, N, nx, **o), torch.empty(B, N, nu, **o), torch.empty(B, nx, **o)
    bpt = _bp_tensor(bp, x.device)
    with torch.cuda.device(x.device):
        if take is not None:
            if x.data_ptr() != states.data_ptr() or u.data_ptr() != controls.data_ptr():
                raise L.IpocError("linearize(take=...) writes states / controls in place: pass contiguous float64 CUDA tensors")
            L.check(L.lib().ipoc_plant_take_linearize_f64(plant["id"], N, B, plant["Ts"], plant["bound"], L.ptr(bpt),
                                                          L.ptr(take[0]), L.ptr(take[1]), L.ptr(x), L.ptr(u), L.ptr(fx),
                                                          L.ptr(fu), L.ptr(cx), L.ptr(cu), L.ptr(lamT), L.ptr(fresh),
                                                          L.stream_ptr()))
        else:
            L.check(L.lib().ipoc_plant_linearize_f64(plant["id"], N, B, plant["Ts"], plant["bound"], L.ptr(bpt), L.ptr(x),
                                                     L.ptr(u), L.ptr(fx), L.ptr(fu), L.ptr(cx), L.ptr(cu), L.ptr(lamT),
                                                     L.ptr(fresh), L.stream_ptr()))
    out = (fx, fu, cx, cu, lamT)
    return out if batched else tuple(t[0] for t in out)


def hamiltonian(plant, states, controls, lam, bp, fresh=None, out=None):
    """Second-order pass (after the costate scan): ru, Q, R, M = H_u, H_xx, H_uu, H_xu of
    H = stage_cost + lam[k+1]' f  (== compute_lqr_params, ref noc/par_interior_point_newton.py:31-42)."""
    x, u, batched = _prep(states, controls)
    lam = L.dev_f64(lam)
    if not batched:
        lam = lam.unsqueeze(0)
    B, N = u.shape[0], u.shape[1]
    nx, nu, _ = _dims(plant)
    o = dict(dtype=torch.float64, device=x.device)
    if out is not None:
        ru, Q, R, M = out
    else:
        ru, Q = torch.empty(B, N, nu, **o), torch.empty(B, N, nx, nx, **o)
        R, M = torch.empty(B, N, nu, nu, **o), torch.empty(B, N, nx, nu, **o)
    bpt = _bp_tensor(bp, x.device)
    with torch.cuda.device(x.device):
        L.check(L.lib().ipoc_plant_hamiltonian_f64(plant["id"], N, B, plant["Ts"], plant["bound"], L.ptr(bpt), L.ptr(x),
                                                   L.ptr(u), L.ptr(lam), L.ptr(ru), L.ptr(Q), L.ptr(R), L.ptr(M),
                                                   L.ptr(fresh), L.stream_ptr()))
    out = (ru, Q, R, M)
    return out if batched else tuple(t[0] for t in out)


_cost_ws = {}


def cost_scratch(N, B, dev, private=False):
    """(tensor or None, bytes) — zero-filled scratch for the grid form of the cost kernel (long horizons, small
    batches; ipoc_plant_cost_workspace_bytes).  Shared per (device, stream) unless `private` (graph captures)."""
    nbytes = int(L.lib().ipoc_plant_cost_workspace_bytes(N, B))
    if nbytes == 0:
        return None, 0
    if private:
        return torch.zeros(nbytes, dtype=torch.uint8, device=dev), nbytes
    dev = torch.device(dev)
    key = (dev.index or 0, torch.cuda.current_stream(dev).cuda_stream)
    buf = _cost_ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = _cost_ws[key] = torch.zeros(2 * nbytes, dtype=torch.uint8, device=dev)
    return buf, buf.numel()


def cost(plant, states, controls, bp, fresh=None, out=None, scratch=None):
    """-> (total_cost (B,), feasible (B,) int32) of trajectories: final cost + sum of stage costs (log barrier
    included; NaN where infeasible, as in the reference) and all(constraints <= 0)."""
    x, u = L.dev_f64(states), L.dev_f64(controls)
    if x.dim() == 2:
        x, u = x.unsqueeze(0), u.unsqueeze(0)
    B, N = u.shape[0], u.shape[1]
    dev = x.device
    if out is not None:
        total, feas = out
    else:
        total = torch.empty(B, dtype=torch.float64, device=dev)
        feas = torch.empty(B, dtype=torch.int32, device=dev)
    bpt = _bp_tensor(bp, dev)
    ws, nbytes = scratch if scratch is not None else cost_scratch(N, B, dev)
    with torch.cuda.device(dev):
        L.check(L.lib().ipoc_plant_cost_f64(plant["id"], N, B, plant["Ts"], plant["bound"], L.ptr(bpt), L.ptr(x),
                                            L.ptr(u), L.ptr(total), L.ptr(feas), L.ptr(fresh), L.ptr(ws), nbytes,
                                            L.stream_ptr()))
    return total, feas


def rollout(plant, controls, initial_state):
    """Serial rollout on the device, one thread per problem (ref noc/utils.py:57-63)."""
    u = L.dev_f64(controls)
    x0 = L.dev_f64(initial_state, u.device)
    batched = u.dim() == 3
    if not batched:
        u, x0 = u.unsqueeze(0), x0.unsqueeze(0)
    B, N = u.shape[0], u.shape[1]
    nx = x0.shape[-1]
    x = torch.empty(B, N + 1, nx, dtype=torch.float64, device=u.device)
    with torch.cuda.device(u.device):
        L.check(L.lib().ipoc_plant_rollout_f64(plant["id"], N, B, plant["Ts"], L.ptr(x0.contiguous()), L.ptr(u), L.ptr(x),
                                               L.stream_ptr()))
    return x if batched else x[0]


def rollout_parallel(plant, controls, initial_state, x_guess=None, tol=1e-13, max_iter=100):
    """Parallel-in-time rollout of a built-in plant (the reference's `rollout` is a serial
    `lax.scan`, ref noc/utils.py:57-63): Newton's method on the rollout equations x_{k+1} = f(x_k, u_k).  Per iteration
    one linearisation kernel along the current guess (ipoc_plant_rollout_lin_f64: F_k, c_k, f(x_k,u_k), the guess's
    defect) and one forward affine scan (ipoc_affine_scan_f64) for the next guess; the host reads 16 bytes per
    problem and iteration.  The iteration stops when the defect max_k |x_{k+1} - f(x_k,u_k)| is at rounding level
    (tol * (1 + max|x|)) or has stopped shrinking below 1e-10; the result is (x_0, f(x_k, u_k)) — every state is the
    serial rollout's update of its predecessor.  `x_guess` (e.g. the previous barrier stage's trajectory) saves
    iterations.  Falls back to the serial kernel on divergence.  -> (states, iterations or -1)."""
    from .noc import affine_scan
    u = L.dev_f64(controls)
    x0 = L.dev_f64(initial_state, u.device)
    batched = u.dim() == 3
    if not batched:
        u, x0 = u.unsqueeze(0), x0.unsqueeze(0)
        x_guess = None if x_guess is None else x_guess.unsqueeze(0)
    u = u.contiguous()
    B, N, nx = u.shape[0], u.shape[1], x0.shape[-1]
    o = dict(dtype=torch.float64, device=u.device)
    if x_guess is not None and tuple(x_guess.shape) == (B, N + 1, nx) and bool(torch.isfinite(x_guess).all()):
        X = L.dev_f64(x_guess, u.device).clone()
    else:
        X = x0.unsqueeze(1).expand(B, N + 1, nx).contiguous()
    X[:, 0] = x0
    F, c, fv = torch.empty(B, N, nx, nx, **o), torch.empty(B, N, nx, **o), torch.empty(B, N, nx, **o)
    stats = torch.empty(B, 2, **o)
    prev = float("inf")
    with torch.cuda.device(u.device):
        for it in range(1, max_iter + 1):
            L.check(L.lib().ipoc_plant_rollout_lin_f64(plant["id"], N, B, plant["Ts"], L.ptr(X), L.ptr(u), L.ptr(F),
                                                       L.ptr(c), L.ptr(fv), L.ptr(stats), L.stream_ptr()))
            st = stats.cpu()
            defect, scale = float(st[:, 0].max()), 1.0 + float(st[:, 1].max())
            if not (defect == defect) or defect == float("inf") or scale == float("inf"):
                break                                                     # diverged: serial kernel below
            if defect <= tol * scale or (defect <= 1e-10 * scale and defect >= 0.5 * prev):
                X[:, 1:] = fv
                return (X if batched else X[0]), it
            prev = defect
            X = affine_scan(F, c, x0, reverse=False, transpose=False)
    return rollout(plant, controls, initial_state), -1

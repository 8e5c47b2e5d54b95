"""Batched independent OCPs (BASELINE config 5: MPC over many initial states).

No reference counterpart as a script — the reference solves one OCP at a time — but the natural JAX
spelling would be `jax.vmap(par_interior_point_optimal_control, in_axes=(None, 0, 0))`, under which
every `lax.while_loop` runs until ALL members are done while finished members are frozen by
`select`: per-member iterates and iteration counts equal those of solving each problem alone
(SURVEY.md Appendix B).  This module reproduces exactly that: every OCP keeps its own regularisation
`rp`, `r_inc`, attempt counter, Newton counter and done-masks; the kernels run on the whole batch
(`batch` axis of the C ABI), the accept/update kernel only touches the active members.

Multi-GPU: split the batch with `sharded.shard_batch(batch, rank, world)` and call this on the local
slice — no data-path collective.
"""
import torch
from torch.func import vmap, grad, hessian, jacrev
from . import _lib as L
from .optimal_control_problem import OCP, Derivatives
from .noc import (reductions, newton_step, accept_update, eval_iteration, eval_trial, attempt_begin, trial_point,
                  attempt_commit)
from . import plants


def rollout_batched(dynamics, controls, initial_states):
    """(B,N,nu),(B,nx) -> (B,N+1,nx): serial in time, vectorised over the batch (ref noc/utils.py:57-63)."""
    step = vmap(dynamics)
    x = initial_states
    xs = [x]
    with torch.no_grad():
        for k in range(controls.shape[1]):
            x = step(x, controls[:, k])
            xs.append(x)
    return torch.stack(xs, dim=1).contiguous()


def compute_derivatives_batched(ocp: OCP, states, controls, bp) -> Derivatives:
    """ref noc/par_interior_point_newton.py:13-28, vmapped over (batch, time)."""
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

    B, N = controls.shape[0], controls.shape[1]
    flat = vmap(body)(states[:, :-1].reshape(B * N, -1), controls.reshape(B * N, -1))
    return Derivatives(*(t.reshape((B, N) + tuple(t.shape[1:])).contiguous() for t in flat))


class _TailGraph:
    """Static-shape accept/reject attempt body for the LAST few members of an attempt loop.

    A batched solve spends most of its attempts on a handful of hard members (cartpole, 2048 OCPs: 18.8 k of
    19.4 k attempt trips run on <= 8 members, 13.7 k on exactly one), where the ~25 small host-framework ops and
    the host sync of the eager trip cost 4x the kernels.  When <= K members are left, their rows are copied
    once into static buffers and the attempt body (ref noc/par_interior_point_newton.py:153-175) is replayed
    as one CUDA graph; finished members are frozen by their `done` flag exactly like the `select` of a
    vmapped `lax.while_loop`, so replaying in bursts between host checks cannot change any result.
    Two sizes are kept (SIZES): a session starts in the smallest graph that holds its members and migrates
    to the smaller one when enough of them have finished (a replay costs roughly in proportion to K)."""
    SIZES = (8, 32)
    BURST = 8

    def __init__(self, ocp, N, nx, nu, dev, K):
        self.ocp, self.K = ocp, K
        o = dict(dtype=torch.float64, device=dev)
        eye = torch.eye(nx, **o)
        self.fx = eye.repeat(K, N, 1, 1)
        self.fu = torch.zeros(K, N, nx, nu, **o)
        self.ru = torch.zeros(K, N, nu, **o)
        self.Q = eye.repeat(K, N, 1, 1)
        self.R = torch.eye(nu, **o).repeat(K, N, 1, 1)
        self.M = torch.zeros(K, N, nx, nu, **o)
        self.x, self.tx = torch.zeros(K, N + 1, nx, **o), torch.zeros(K, N + 1, nx, **o)
        self.u, self.tu = torch.zeros(K, N, nu, **o), torch.zeros(K, N, nu, **o)
        self.cost, self.cu_norm = torch.zeros(K, **o), torch.zeros(K, **o)
        self.rp, self.rinc = torch.ones(K, **o), torch.full((K,), 2.0, **o)
        self.inner = torch.zeros(K, dtype=torch.int64, device=dev)
        self.done = torch.ones(K, dtype=torch.bool, device=dev)
        self.bp = torch.zeros((), **o)
        self.cx, self.cu = torch.zeros(K, N + 1, nx, **o), torch.zeros(K, N, nu, **o)   # trial point
        self.act = torch.zeros(K, dtype=torch.int32, device=dev)
        self.succ = torch.zeros(K, dtype=torch.int32, device=dev)
        self.reg, self.gain = torch.zeros(K, **o), torch.zeros(K, **o)
        self.rows = (self.fx, self.fu, self.ru, self.Q, self.R, self.M, self.x, self.u, self.cost, self.cu_norm)
        self.state = (self.rp, self.rinc, self.tx, self.tu, self.inner)
        self.graph = None

    def _body(self):
        attempt_begin(self.done, self.rp, self.cu_norm, self.act, self.reg)                     # :117, :177-182
        dx, du, _, _, pred, bwd_feas = newton_step(self.fx, self.fu, self.ru, self.Q, self.R, self.M, self.reg)  # :153
        trial_point(self.x, dx, self.u, du, self.cx, self.cu)                                   # :156-157
        new_cost, traj_feas = eval_trial(self.ocp, self.cx, self.cu, self.bp)                   # :159-163
        with torch.cuda.device(self.rp.device):
            L.check(L.lib().ipoc_accept_update_f64(self.K, L.ptr(self.cost), L.ptr(new_cost.contiguous()),
                                                   L.ptr(traj_feas), L.ptr(pred), L.ptr(bwd_feas), L.ptr(self.act),
                                                   L.ptr(self.rp), L.ptr(self.rinc), L.ptr(self.succ),
                                                   L.ptr(self.gain), L.stream_ptr()))           # :159-173
        attempt_commit(self.act, self.succ, self.cx, self.cu, self.tx, self.tu, self.inner, self.done)   # :174-182

    def capture(self):
        dev = self.fx.device
        self.done.fill_(True)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self._body()
        self.graph = g

    def load(self, rows, state, bp):
        k = state[0].numel()
        for dst, src in zip(self.rows + self.state, tuple(rows) + tuple(state)):
            dst[:k].copy_(src)
        self.done.fill_(True)
        self.done[:k] = False
        self.bp.fill_(float(bp))
        return k


def _run_tail(graphs, rows, state, bp):
    """graphs: captured _TailGraph objects by ascending K.  rows = the k members' (fx, fu, ru, Q, R, M, x, u, cost,
    cu_norm), state = their (rp, rinc, tx, tu, inner).  Runs the members' attempt loops to completion and
    returns their final state."""
    k = state[0].numel()
    level = next(i for i, g in enumerate(graphs) if g.K >= k)
    g = graphs[level]
    g.load(rows, state, bp)
    while True:
        for _ in range(g.BURST):
            g.graph.replay()
        done = g.done[:k].cpu()
        left = int(k - int(done.sum()))
        if left == 0:
            break
        if level > 0 and left <= graphs[level - 1].K:      # migrate the survivors to the smaller graph
            idx = torch.nonzero(~done).reshape(-1).to(g.done.device)
            res = _run_tail(graphs[:level], tuple(r[:k].index_select(0, idx) for r in g.rows),
                            tuple(t[:k].index_select(0, idx) for t in g.state), bp)
            for dst, src in zip(g.state, res):
                dst[idx] = src
            break
    return tuple(t[:k].clone() for t in g.state)


_tail_cache = {}


def _tail_graphs(ocp, N, nx, nu, dev):
    """Captured tail bodies for this problem / horizon (ascending K), or None if the OCP's callables cannot
    be captured."""
    key = (id(ocp.dynamics), id(ocp.stage_cost), id(ocp.final_cost), id(ocp.constraints), id(ocp.total_cost),
           N, nx, nu, str(dev), plants.ENABLED)
    gs = _tail_cache.get(key)
    if gs is None:
        try:
            gs = []
            for K in _TailGraph.SIZES:
                g = _TailGraph(ocp, N, nx, nu, dev, K)
                g.capture()
                gs.append(g)
        except Exception as e:
            import warnings
            warnings.warn(f"ipoc_b200: CUDA-graph capture of the batched attempt body failed "
                          f"({type(e).__name__}: {str(e)[:120]}); using eager launches")
            torch.cuda.synchronize(dev)
            gs = False
        while len(_tail_cache) >= 4:
            _tail_cache.pop(next(iter(_tail_cache)))
        _tail_cache[key] = gs
    return gs or None


# ------------------------------------------------------------------ device-resident batched loops
class _DeviceLadder:
    """Device-resident Newton + attempt loops (graphed.DeviceLoopNewton with a batch axis) for a LADDER of batch
    sizes 8, 16, ..., 2^k >= B.  A stage starts in the smallest loop that holds its members; every replay advances
    ALL members still inside their loops by one accept/reject attempt (finished members are frozen on the device,
    the `select` of a vmapped `lax.while_loop`), the host looks at the exit flags once per burst and MIGRATES the
    survivors to the next smaller loop when at most half a loop is still alive — so the long tail of a few hard
    members (cartpole: iterations_max 340 against a mean of 127) runs in small graphs instead of dragging the whole
    batch through every attempt, and no eager host-framework op or per-attempt host sync is left on the path."""
    MIN = 8

    def __init__(self, ocp, N, nx, nu, dev):
        self.ocp, self.N, self.nx, self.nu, self.dev = ocp, N, nx, nu, dev
        self.loops = {}

    def loop(self, size, x, u, bp):
        from .graphed import DeviceLoopNewton
        lp = self.loops.get(size)
        if lp is None:
            lp = DeviceLoopNewton(self.ocp, self.N, self.nx, self.nu, self.dev, batch=size)
            k = x.shape[0]
            reps = (size + k - 1) // k
            lp.capture(x.repeat(reps, 1, 1)[:size], u.repeat(reps, 1, 1)[:size], bp)
            self.loops[size] = lp
        return lp

    @staticmethod
    def _size_for(k):
        s = _DeviceLadder.MIN
        while s < k:
            s *= 2
        return s

    @staticmethod
    def _load(lp, k, x, u, tx, tu, rp, rinc, inner, iteration, bp):
        """Members 0..k-1 <- the given state; the padding slots are frozen copies of member 0."""
        lp.bp.fill_(float(bp))
        for dst, src in ((lp.x, x), (lp.u, u), (lp.tx, tx), (lp.tu, tu), (lp.rp, rp), (lp.r_inc, rinc), (lp.inner, inner),
                         (lp.iteration, iteration)):
            dst[:k].copy_(src)
            if k < lp.B:
                dst[k:].copy_(src[:1].expand((lp.B - k,) + tuple(src.shape[1:])))
        # the trial buffers of an earlier attempt are dead once a new attempt starts (the next one overwrites them),
        # so tx := x makes "take the step" a no-op and adv = 1 marks every iterate as new in this loop's buffers
        lp.tx.copy_(lp.x)
        lp.tu.copy_(lp.u)
        lp.act.fill_(0)
        lp.act[:k] = 1
        lp.outer_done.fill_(True)
        lp.outer_done[:k] = False
        lp.adv.fill_(1)
        lp.need_cost.fill_(1)
        lp.buf.hu.fill_(1.0)

    def run_stage(self, x, u, bp):
        """`newton_oc` for every member from the iterate (x, u): -> (x, u, iterations (B,) int64)."""
        B = x.shape[0]
        dev = self.dev
        out_x, out_u = x.clone(), u.clone()
        out_it = torch.zeros(B, dtype=torch.int64, device=dev)
        ids = torch.arange(B, device=dev)                     # global index of the member in each live slot
        o = dict(dtype=torch.float64, device=dev)
        st = dict(x=x, u=u, tx=x.clone(), tu=u.clone(), rp=torch.ones(B, **o), rinc=torch.full((B,), 2.0, **o),
                  inner=torch.zeros(B, dtype=torch.int64, device=dev),
                  iteration=torch.zeros(B, dtype=torch.int64, device=dev))
        while ids.numel() > 0:
            k = ids.numel()
            size = self._size_for(k)
            lp = self.loop(size, st["x"], st["u"], bp)
            self._load(lp, k, st["x"], st["u"], st["tx"], st["tu"], st["rp"], st["rinc"], st["inner"], st["iteration"], bp)
            burst = 4 if size * self.N > 2_000_000 else 8
            while True:
                for _ in range(burst):
                    lp.graph.replay()
                done = lp.outer_done[:k].clone()
                alive = int(k - int(done.sum()))              # the one host sync per burst
                if alive == 0 or (size > self.MIN and alive <= size // 2):
                    break
            lp.take_last_step()                               # x <- tx for members whose last attempt ended an iteration
            fin = torch.nonzero(done).reshape(-1)
            if fin.numel() > 0:
                g = ids[fin]
                out_x[g], out_u[g], out_it[g] = lp.x[fin], lp.u[fin], lp.iteration[fin]
            keep = torch.nonzero(~done).reshape(-1)
            ids = ids[keep]
            if keep.numel() > 0:
                st = dict(x=lp.x[keep], u=lp.u[keep], tx=lp.tx[keep], tu=lp.tu[keep], rp=lp.rp[keep], rinc=lp.r_inc[keep],
                          inner=lp.inner[keep], iteration=lp.iteration[keep])
        return out_x, out_u, out_it


_ladder_cache = {}


def _device_ladder(ocp, N, nx, nu, dev):
    key = (id(ocp.dynamics), id(ocp.stage_cost), id(ocp.final_cost), id(ocp.constraints), id(ocp.total_cost),
           N, nx, nu, str(dev), plants.ENABLED)
    ld = _ladder_cache.get(key)
    if ld is None:
        while len(_ladder_cache) >= 2:
            _ladder_cache.pop(next(iter(_ladder_cache)))
        ld = _ladder_cache[key] = _DeviceLadder(ocp, N, nx, nu, dev)
        ld._keepalive = ocp
    return ld


def newton_oc_batched(ocp: OCP, controls, initial_states, barrier_param, use_graphs: bool = True):
    """Per-member semantics of ref noc/par_interior_point_newton.py:127-225 for a batch.
    -> (x (B,N+1,nx), u (B,N,nu), iterations (B,) int64)

    Members leave the two loops at different times; the work is COMPACTED to the members still inside
    (index_select of the active rows), so a few hard members with hundreds of rejected attempts do not
    drag the whole batch through every attempt."""
    dev = controls.device
    u_all = L.dev_f64(controls).clone()
    plant = plants.plant_of(ocp)
    if plant is not None:
        x_all = plants.rollout(plant, u_all, initial_states.to(dev))               # :133
    else:
        x_all = rollout_batched(ocp.dynamics, u_all, initial_states.to(dev))
    B = u_all.shape[0]
    if use_graphs and plant is not None and use_graphs != "host":
        # built-in plants: both loops live on the device for every member (graphed.DeviceLoopNewton, batch axis)
        return _device_ladder(ocp, u_all.shape[1], x_all.shape[-1], u_all.shape[-1], dev).run_stage(x_all, u_all,
                                                                                                     barrier_param)
    o = dict(dtype=torch.float64, device=dev)
    rp_all = torch.ones(B, **o)                                                    # :134
    rinc_all = torch.full((B,), 2.0, **o)                                          # :135
    iters = torch.zeros(B, dtype=torch.int64, device=dev)
    tail = _tail_graphs(ocp, u_all.shape[1], x_all.shape[-1], u_all.shape[-1], dev) if use_graphs else None
    act = torch.arange(B, device=dev)                                              # members still in the Newton loop
    while act.numel() > 0:                                                         # :199-202 (per member)
        full = act.numel() == B
        x, u = (x_all, u_all) if full else (x_all[act], u_all[act])
        cost, fx, fu, cu, ru, Q, R, M = eval_iteration(ocp, x, u, barrier_param)   # :142-149
        hu, cu_norm, _ = reductions(ru=ru, cu=cu)                                  # :158, :116
        na = act.numel()
        rp, rinc = rp_all[act], rinc_all[act]
        tx, tu = x.clone(), u.clone()
        inner = torch.zeros(na, dtype=torch.int64, device=dev)
        sub = torch.arange(na, device=dev)                                         # members still in the attempt loop
        while sub.numel() > 0:                                                     # :177-182 (per member)
            if tail is not None and sub.numel() <= tail[-1].K:
                rows = tuple(t.index_select(0, sub) for t in (fx, fu, ru, Q, R, M, x, u, cost, cu_norm))
                rp[sub], rinc[sub], tx[sub], tu[sub], inner[sub] = _run_tail(
                    tail, rows, (rp[sub], rinc[sub], tx[sub], tu[sub], inner[sub]), barrier_param)
                break
            if sub.numel() == na:
                a = (fx, fu, ru, Q, R, M, x, u, cost, cu_norm)
            else:
                a = tuple(t.index_select(0, sub) for t in (fx, fu, ru, Q, R, M, x, u, cost, cu_norm))
            rp_s, rinc_s = rp[sub].contiguous(), rinc[sub].contiguous()
            dx, du, _, _, pred, bwd_feas = newton_step(*a[:6], rp_s * a[9])        # :153
            cx_try, cu_try = a[6] + dx, a[7] + du                                  # :156-157
            new_cost, traj_feas = eval_trial(ocp, cx_try, cu_try, barrier_param)   # :159-163
            succ, _ = accept_update(a[8].contiguous(), new_cost.contiguous(), traj_feas, pred, bwd_feas, rp_s,
                                    rinc_s)                                        # :159-173
            rp[sub], rinc[sub] = rp_s, rinc_s
            tx[sub], tu[sub] = cx_try, cu_try                                      # :175 (kept regardless of success)
            inner[sub] += 1                                                        # :174
            sub = sub[~((succ != 0) | (inner[sub] > 500))]
        if full:
            x_all, u_all, rp_all, rinc_all = tx, tu, rp, rinc                      # :184
        else:
            x_all[act], u_all[act], rp_all[act], rinc_all[act] = tx, tu, rp, rinc
        iters[act] += 1                                                            # :194
        act = act[~((hu < 1e-4) | (iters[act] > 1000))]
    return x_all, u_all, iters


def par_interior_point_optimal_control_batched(ocp: OCP, controls, initial_states, use_graphs: bool = True):
    """Batched `par_interior_point_optimal_control` (ref :228-254): controls (B,N,nu), initial_states (B,nx)
    -> (opt_u (B,N,nu), N_iterations (B,) int64)."""
    if not controls.is_cuda:
        raise L.IpocError("the batched solver needs CUDA tensors; there is no CPU fallback")
    u = controls
    total = torch.zeros(controls.shape[0], dtype=torch.int64, device=controls.device)
    bp = 0.1                                                                       # :233
    while bp > 1e-4:                                                               # :243-245
        _, u, its = newton_oc_batched(ocp, u, initial_states, bp, use_graphs)      # :237
        bp = bp / 5                                                                # :238
        total = total + its                                                        # :239
    return u, total

"""CUDA-graph execution of the two loop bodies of `newton_oc` (ref noc/par_interior_point_newton.py:137-197).

In the reference the whole solver is one `jax.jit` executable; here the host framework's part of an
iteration (vmapped autodiff, `compute_lqr_params`, cost / constraint evaluation) is a few hundred small
eager kernels.  Both bodies are captured once per (problem, horizon) into CUDA graphs — "streams and
graphs instead of a tracing compiler" — and replayed:

  graph A (once per Newton iteration): total cost, derivatives, K1 costates, LQ parameters,
          K4 reductions (max|ru|, ||cu||)
  graph B (once per accept/reject attempt): reg = rp*||cu||, K2+K3 Newton step, trial trajectory,
          constraints + K4 feasibility, new cost, A8 accept / regularisation update

The barrier parameter is a device scalar, so one capture serves all five barrier stages and every later
solve of the same problem (the warm-up solve pays the capture, like jit compilation in the reference's
timing protocol, ref examples/cartpole_runtime.py:119-146).  Arithmetic and its order are those of the
eager path, so iterates and iteration counts are identical.
"""
import torch
from torch.func import vmap
from . import _lib as L
from .optimal_control_problem import OCP
from . import noc


class GraphedNewton:
    def __init__(self, ocp: OCP, N: int, nx: int, nu: int, device):
        self.ocp, self.N, self.nx, self.nu, self.dev = ocp, N, nx, nu, torch.device(device)
        o = dict(dtype=torch.float64, device=self.dev)
        self.x = torch.zeros(N + 1, nx, **o)
        self.u = torch.zeros(N, nu, **o)
        self.bp = torch.zeros((), **o)
        self.rp = torch.ones(1, **o)
        self.r_inc = torch.full((1,), 2.0, **o)
        self.graph_a = self.graph_b = None

    # ---- loop bodies (same statements as noc.newton_oc) ------------------------------------------------
    def _iteration(self):
        (self.cost, self.fx, self.fu, cu, self.ru, self.Q, self.R,
         self.M) = noc.eval_iteration(self.ocp, self.x, self.u, self.bp)           # :142-149
        self.hu, self.cu_norm, _ = noc.reductions(ru=self.ru, cu=cu)               # :158, :116

    def _attempt(self):
        ocp = self.ocp
        dx, du, _, _, pred, bwd_feas = noc.newton_step(self.fx, self.fu, self.ru, self.Q, self.R, self.M,
                                                       self.rp * self.cu_norm)     # :153
        self.tu = self.u + du                                                      # :156
        self.tx = self.x + dx                                                      # :157
        new_cost, traj_feas = noc.eval_trial(ocp, self.tx, self.tu, self.bp)       # :159-163
        succ, gain = noc.accept_update(self.cost, new_cost, traj_feas, pred, bwd_feas, self.rp, self.r_inc)
        self.rec = torch.stack((succ[0].to(torch.float64), self.hu[0], new_cost[0], pred[0], gain[0]))

    # ---- capture ---------------------------------------------------------------------------------------
    def capture(self, x, u, bp):
        self.x.copy_(x)
        self.u.copy_(u)
        self.bp.fill_(float(bp))
        rp0, ri0 = self.rp.clone(), self.r_inc.clone()
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):          # eager warm-up (allocator, workspaces, lazy inits)
            for _ in range(2):
                self._iteration()
                self._attempt()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.rp.copy_(rp0)
        self.r_inc.copy_(ri0)
        ga = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga, stream=side):
            self._iteration()
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, pool=ga.pool(), stream=side):
            self._attempt()
        self.rp.copy_(rp0)
        self.r_inc.copy_(ri0)
        self.graph_a, self.graph_b = ga, gb

    def iteration(self):
        self.graph_a.replay()

    def attempt(self):
        """-> (success, Hu_norm, new_cost, pred, gain_ratio): the one host read per attempt."""
        self.graph_b.replay()
        rec = self.rec.cpu()
        return bool(rec[0] != 0), float(rec[1]), float(rec[2]), float(rec[3]), float(rec[4])


# Captured graphs are cached per (identity of the OCP's five callables, horizon, dimensions, device) — the semantics
# of a jit cache: a captured graph bakes in every Python scalar and global the user functions read, so changing a
# closed-over constant WITHOUT creating new callables replays stale arithmetic; call `clear_cache()` (or build a new
# OCP) after such a change.  Failed captures are remembered too (with the OCP kept alive, so that its ids cannot be
# recycled by another object).
_cache = {}          # insertion-ordered; bounded so that throw-away OCP closures cannot pile up graphs
_CACHE_MAX = 8
_failed_keepalive = {}


def clear_cache():
    """Drop every captured loop body / device-resident loop / batched ladder (they are re-captured on next use)."""
    _cache.clear()
    _loop_cache.clear()
    _failed_keepalive.clear()
    from . import batched
    batched._tail_cache.clear()
    batched._ladder_cache.clear()


def get(ocp: OCP, N, nx, nu, device, x, u, bp):
    """Captured bodies for this problem/horizon (cached on the identity of the OCP's callables)."""
    from . import plants
    key = (id(ocp.dynamics), id(ocp.stage_cost), id(ocp.final_cost), id(ocp.constraints), id(ocp.total_cost),
           N, nx, nu, str(device), plants.ENABLED)
    g = _cache.get(key)
    if g is None:
        g = GraphedNewton(ocp, N, nx, nu, device)
        try:
            g.capture(x, u, bp)
        except Exception as e:   # user callables that cannot be captured (host syncs, H2D copies ...)
            import warnings
            warnings.warn(f"ipoc_b200: CUDA-graph capture of the loop bodies failed ({type(e).__name__}: "
                          f"{str(e)[:120]}); using eager launches for this problem")
            torch.cuda.synchronize(g.dev)
            g = False
            _failed_keepalive[key] = ocp
        else:
            g._keepalive = ocp
        while len(_cache) >= _CACHE_MAX:
            _failed_keepalive.pop(next(iter(_cache)), None)
            _cache.pop(next(iter(_cache)))
        _cache[key] = g
    return g


class DeviceLoopNewton:
    """The WHOLE `newton_oc` loop (ref noc/par_interior_point_newton.py:137-202) with its control flow on the
    device, for one problem or a batch of independent problems: one CUDA graph = one
    accept/reject attempt of every member,

        take step : x <- tx, u <- tu for members whose attempt loop ended in the previous replay (:184)
                                                                                    (ipoc_masked_copy_f64)
        iterate   : cost, linearisation, K1 costates (+ ||cu||), LQ parameters at (x, u)   (:142-149; recomputed
                    — idempotently — when the previous attempt was rejected)
        attempt   : K2 + K3 with reg = rp*||cu|| (:117, :153), max|ru| (:158) and the trial point (:156-157) as
                    side jobs of the scan kernels                                   (ipoc_newton_attempt_f64)
        finish    : trial cost / feasibility, accept rule and rp update (:159-173), attempt counter and inner
                    exit (:174-182), iteration counter and outer exit on the PRE-step max|ru| (:194-202)
                                        (ipoc_plant_attempt_finish_f64 / eval_trial + ipoc_attempt_finish_f64)

    Ten launches for the built-in plants.  Every member is frozen once its `outer_done` is set — exactly the
    `select` semantics of a (vmapped) `lax.while_loop` — so the host may run ahead: it keeps `depth` replays
    queued and looks at the exit flags of an OLDER replay (copied to pinned memory behind each replay), i.e. no
    host round trip sits on the critical path.  Iterates and iteration counts are those of the eager loop."""

    def __init__(self, ocp: OCP, N: int, nx: int, nu: int, device, batch: int = 1):
        self.ocp, self.N, self.nx, self.nu, self.dev, self.B = ocp, N, nx, nu, torch.device(device), batch
        B = batch
        o = dict(dtype=torch.float64, device=self.dev)
        i32 = dict(dtype=torch.int32, device=self.dev)
        self.x, self.tx = torch.zeros(B, N + 1, nx, **o), torch.zeros(B, N + 1, nx, **o)
        self.u, self.tu = torch.zeros(B, N, nu, **o), torch.zeros(B, N, nu, **o)
        self.bp = torch.zeros((), **o)
        self.rp, self.r_inc = torch.ones(B, **o), torch.full((B,), 2.0, **o)
        self.gain, self.new_cost = torch.zeros(B, **o), torch.zeros(B, **o)
        self.act, self.succ, self.adv = torch.ones(B, **i32), torch.zeros(B, **i32), torch.zeros(B, **i32)
        self.traj_feas = torch.zeros(B, **i32)
        self.inner = torch.zeros(B, dtype=torch.int64, device=self.dev)
        self.iteration = torch.zeros(B, dtype=torch.int64, device=self.dev)
        self.outer_done = torch.zeros(B, dtype=torch.bool, device=self.dev)
        self.buf = noc.AttemptBuffers(B, N, nx, nu, self.dev)
        self._eval = {}                            # iterate-evaluation buffers, kept across replays (see _step)
        self._cost_ws = None
        self.need_cost = torch.ones(B, dtype=torch.int32, device=self.dev)   # see _step: cost is carried between iterations
        self.depth = 3 if N * B <= 20000 else 1    # replays kept in flight (a replay is > 1 ms for long horizons)
        self.ring = 8
        self.flags = torch.zeros(self.ring, B, dtype=torch.bool).pin_memory()
        self.events = [torch.cuda.Event() for _ in range(self.ring)]
        self.graph = None

    def _step(self):
        from . import plants
        lib, p = L.lib(), L.ptr
        B, N, nx, nu = self.B, self.N, self.nx, self.nu
        plant = plants.plant_of(self.ocp)
        if plant is None:
            with torch.cuda.device(self.dev):
                L.check(lib.ipoc_masked_copy_f64(N, nx, nu, B, p(self.adv), p(self.tx), p(self.tu), p(self.x), p(self.u),
                                                 L.stream_ptr()))                           # :184
        if plant is not None:                                                              # :142-149
            # `adv` doubles as the "iterate changed" flag: after a REJECTED attempt the iterate is the same, every
            # member-wise kernel below skips the member and its buffers keep the previous evaluation
            fr, ev = self.adv, self._eval
            # the step of an accepted attempt (x <- tx, u <- tu where adv != 0, :184) is taken by the linearisation kernel
            ev["lin"] = plants.linearize(plant, self.x, self.u, self.bp, fresh=fr, out=ev.get("lin"),
                                         take=(self.tx, self.tu))
            if self._cost_ws is None:        # private zero-filled scratch of the two cost evaluations of an attempt
                self._cost_ws = (plants.cost_scratch(N, B, self.dev, private=True),
                                 plants.cost_scratch(N, B, self.dev, private=True))
            # the cost of an iterate that came from an accepted trial point was already computed as that attempt's
            # new_cost (same kernel, same data): the finish kernel carries it over and clears `need_cost`, which is
            # set again only where a new iterate is loaded (_reset, batched._load)
            ev["cost"] = plants.cost(plant, self.x, self.u, self.bp, fresh=self.need_cost, out=ev.get("cost"),
                                     scratch=self._cost_ws[0])
            fx, fu, cx, cu, lamT = ev["lin"]
            cost = ev["cost"][0]
            ev["cos"] = noc.costates_fused(fx, cx, lamT, cu, fresh=fr, out=ev.get("cos"))   # :147, :116
            lam, cu_norm = ev["cos"]
            ev["ham"] = plants.hamiltonian(plant, self.x, self.u, lam, self.bp, fresh=fr, out=ev.get("ham"))
            ru, Q, R, M = ev["ham"]
        else:
            cost, fx, fu, cu, ru, Q, R, M = noc.eval_iteration(self.ocp, self.x, self.u, self.bp)
            _, cu_norm, _ = noc.reductions(cu=cu)
        buf = noc.newton_attempt(self.buf, fx, fu, ru, Q, R, M, self.rp, cu_norm, self.x, self.u, self.tx, self.tu,
                                 self.act)                                                 # :153-158
        fin = (p(cost.contiguous()), p(buf.pred), p(buf.bwd_feas), p(buf.hu), p(self.act), p(self.rp), p(self.r_inc),
               p(self.succ), p(self.gain), p(self.inner), p(self.iteration), p(self.outer_done), p(self.adv), 1e-4,
               500, 1000)                                                                  # :159-202
        with torch.cuda.device(self.dev):
            if plant is not None:
                L.check(lib.ipoc_plant_attempt_finish_f64(plant["id"], N, B, plant["Ts"], plant["bound"],
                                                          p(plants._bp_tensor(self.bp, self.dev)), p(self.tx), p(self.tu),
                                                          p(self.new_cost), p(self.traj_feas), *fin, p(cost), p(self.need_cost),
                                                          p(self._cost_ws[1][0]), self._cost_ws[1][1], L.stream_ptr()))
            else:
                new_cost, traj_feas = noc.eval_trial(self.ocp, self.tx, self.tu, self.bp)  # :159-163
                L.check(lib.ipoc_attempt_finish_f64(B, fin[0], p(new_cost.contiguous()), p(traj_feas), *fin[1:],
                                                    L.stream_ptr()))

    def _reset(self, x, u, bp):
        self.x.copy_(x.reshape(self.x.shape))
        self.u.copy_(u.reshape(self.u.shape))
        self.tx.copy_(self.x)
        self.tu.copy_(self.u)
        self.bp.fill_(float(bp))
        self.rp.fill_(1.0)                                                                 # :134
        self.r_inc.fill_(2.0)                                                              # :135
        self.buf.hu.fill_(1.0)
        self.act.fill_(1)
        self.adv.fill_(1)      # "take the step tx -> x" is a no-op (tx == x) and marks every iterate as new
        self.need_cost.fill_(1)
        self.inner.zero_()
        self.iteration.zero_()
        self.outer_done.zero_()

    def capture(self, x, u, bp):
        self._reset(x, u, bp)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self._step()
        self.graph = g

    def launch(self, r):
        """Replay number r; its exit flags land in pinned memory behind it."""
        self.graph.replay()
        k = r % self.ring
        self.flags[k].copy_(self.outer_done, non_blocking=True)
        self.events[k].record()

    def finished(self, r):
        """Exit flags (B,) bool as of replay r (blocks until that replay is done)."""
        k = r % self.ring
        self.events[k].synchronize()
        return self.flags[k]

    def take_last_step(self):
        with torch.cuda.device(self.dev):
            L.check(L.lib().ipoc_masked_copy_f64(self.N, self.nx, self.nu, self.B, L.ptr(self.adv), L.ptr(self.tx),
                                                 L.ptr(self.tu), L.ptr(self.x), L.ptr(self.u), L.stream_ptr()))
        self.adv.zero_()

    def run(self, x, u, bp):
        """-> (x, u, iterations) of `newton_oc` started at (x, u) (single problem: unbatched results)."""
        self._reset(x, u, bp)
        r = 0
        while True:
            self.launch(r)
            r += 1
            if r >= self.depth and bool(self.finished(r - self.depth).all()):
                break
        self.take_last_step()
        if self.B == 1:
            return self.x[0].clone(), self.u[0].clone(), int(self.iteration)
        return self.x.clone(), self.u.clone(), self.iteration.clone()


_loop_cache = {}


def get_device_loop(ocp: OCP, N, nx, nu, device, x, u, bp):
    """Captured device-resident loop for this problem / horizon, or False if it cannot be captured."""
    from . import plants
    key = (id(ocp.dynamics), id(ocp.stage_cost), id(ocp.final_cost), id(ocp.constraints), id(ocp.total_cost),
           N, nx, nu, str(device), plants.ENABLED)
    g = _loop_cache.get(key)
    if g is None:
        g = DeviceLoopNewton(ocp, N, nx, nu, device)
        try:
            g.capture(x, u, bp)
        except Exception as e:
            import warnings
            warnings.warn(f"ipoc_b200: CUDA-graph capture of the Newton loop failed ({type(e).__name__}: "
                          f"{str(e)[:120]}); using the host-steered loop for this problem")
            torch.cuda.synchronize(g.dev)
            g = False
            _failed_keepalive[("loop",) + key] = ocp
        else:
            g._keepalive = ocp
        while len(_loop_cache) >= _CACHE_MAX:
            _failed_keepalive.pop(("loop",) + next(iter(_loop_cache)), None)
            _loop_cache.pop(next(iter(_loop_cache)))
        _loop_cache[key] = g
    return g

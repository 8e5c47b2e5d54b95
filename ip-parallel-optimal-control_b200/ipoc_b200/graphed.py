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
        with torch.cuda.graph(ga):
            self._iteration()
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, pool=ga.pool()):
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


_cache = {}          # insertion-ordered; bounded so that throw-away OCP closures cannot pile up graphs
_CACHE_MAX = 8


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
        else:
            g._keepalive = ocp
        while len(_cache) >= _CACHE_MAX:
            _cache.pop(next(iter(_cache)))
        _cache[key] = g
    return g

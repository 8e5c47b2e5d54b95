"""Receding-horizon loop of the reference's linear MPC example (BASELINE config 3,
ref examples/linear_mpc_parallel.py:67-81): every MPC step runs `par_bwd_pass(lqt)` and
`par_fwd_pass(lqt, x, Kx, d)` on a fixed LQT problem and keeps `x_par[1]`, `u_par[0]`.

In the reference the loop is a `jax.lax.scan` inside one XLA executable; here `unroll` consecutive MPC
steps are captured into ONE CUDA graph: step i reads its initial state straight from the trajectory
buffer of step i-1 (`x_{i-1}[1]`), so the graph contains nothing but the kernels of the two passes, and
a 5000-step simulation is 5000/unroll graph launches.  As in the reference, the backward pass is
recomputed every step even though the problem does not change ("as written").
"""
import torch
from . import _lib as L
from .paroc import LQT, _effective


class MpcLoop:
    def __init__(self, lqt: LQT, unroll: int = 100, serial: bool = False):
        """serial=True runs the `seq_bwd_pass` / `seq_fwd_pass` twins (one chunk per horizon)."""
        lqt = LQT(*(L.dev_f64(t) for t in lqt))
        if lqt.A.dim() != 3:
            raise L.IpocError("MpcLoop takes one (unbatched) LQT problem")
        self.dev = dev = lqt.A.device
        self.T, self.nx, self.nu = lqt.A.shape[0], lqt.A.shape[1], lqt.B.shape[-1]
        L.require_supported(self.nx, self.nu)
        self.unroll, self.serial = int(unroll), bool(serial)
        # H, Z, r, s folded once (host-framework glue of the wrapper; the passes themselves run every step)
        self.A, self.B, self.c = (t.unsqueeze(0).contiguous() for t in (lqt.A, lqt.B, lqt.c))
        self.eff = tuple(L.dev_f64(t.unsqueeze(0)).contiguous() for t in _effective(lqt))
        o = dict(dtype=torch.float64, device=dev)
        T, nx, nu, K = self.T, self.nx, self.nu, self.unroll
        self.Kx, self.d = torch.empty(1, T, nu, nx, **o), torch.empty(1, T, nu, **o)
        self.S, self.v = torch.empty(1, T + 1, nx, nx, **o), torch.empty(1, T + 1, nx, **o)
        self.pred, self.feas = torch.empty(1, **o), torch.empty(1, dtype=torch.int32, device=dev)
        # trajectory of every unrolled step; rows padded to an even number of doubles (the C ABI wants
        # 16-byte aligned pointers)
        even = lambda n: n + (n & 1)
        self._Xp = torch.zeros(K, even((T + 1) * nx), **o)
        self._Up = torch.zeros(K, even(T * nu), **o)
        self.X = self._Xp[:, :(T + 1) * nx].view(K, T + 1, nx)
        self.U = self._Up[:, :T * nu].view(K, T, nu)
        self.xstart = torch.zeros(nx, **o)
        self._xin = torch.zeros(K, even(nx), **o) if nx & 1 else None   # odd nx: x_{i-1}[1] is not 16-byte aligned
        self.ws_b, self.nb_b = L.workspace(L.WS_LQT_BWD, T, nx, nu, 1, dev)
        self.ws_f, self.nb_f = L.workspace(L.WS_LQT_FWD, T, nx, nu, 1, dev)
        self.ws_b, self.ws_f = self.ws_b.clone(), self.ws_f.clone()   # private: the graph keeps their addresses
        self.graph = None

    def _step(self, x0, i):
        """One MPC step: backward pass, forward pass from x0 -> trajectory i."""
        p, lib, s = L.ptr, L.lib(), L.stream_ptr()
        Xe, Ue, Me, q, pp, ST, vT = self.eff
        L.check(lib.ipoc_lqt_bwd_f64(self.T, self.nx, self.nu, 1, p(self.A), p(self.B), p(self.c), p(Xe), p(Ue),
                                     p(Me), p(q), p(pp), p(ST), p(vT), p(self.Kx), p(self.d), p(self.S), p(self.v),
                                     p(self.pred), p(self.feas), p(self.ws_b), self.nb_b, s))
        L.check(lib.ipoc_lqt_fwd_f64(self.T, self.nx, self.nu, 1, p(self.A), p(self.B), p(self.c), p(self.Kx),
                                     p(self.d), p(x0), p(self.U[i]), p(self.X[i]), p(self.ws_f), self.nb_f, s))

    def _chunk(self):
        for i in range(self.unroll):
            x0 = self.xstart if i == 0 else self.X[i - 1, 1]
            if i > 0 and self._xin is not None:
                x0 = self._xin[i, :self.nx].copy_(x0)
            self._step(x0, i)

    def _with_tuning(self, fn):
        if not self.serial:
            return fn()
        with L.tuning(leaf_chunk=self.T):     # one chunk per horizon = the serial twins; knobs restored, lock held
            fn()

    def capture(self):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self._with_tuning(lambda: self._step(self.xstart, 0))      # lazy initialisations outside the capture
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()

        def cap():
            with torch.cuda.graph(g, stream=side):
                self._chunk()
        self._with_tuning(cap)
        self.graph = g

    def run(self, x0, steps: int, use_graph: bool = True):
        """-> (xs (steps, nx), us (steps, nu)): the closed-loop states x_par[1] and controls u_par[0]."""
        if use_graph and self.graph is None:
            self.capture()
        xs, us = [], []
        self.xstart.copy_(L.dev_f64(x0, self.dev))
        done = 0
        while done < steps:
            if use_graph:
                self.graph.replay()
            else:
                self._with_tuning(self._chunk)
            k = min(self.unroll, steps - done)
            xs.append(self.X[:k, 1].clone())
            us.append(self.U[:k, 0].clone())
            self.xstart.copy_(self.X[k - 1, 1])
            done += k
        return torch.cat(xs), torch.cat(us)


def constrained_mpc(ocp, x0, horizon: int = 40, sim_steps: int = 20, u_init=None, use_graphs: bool = True):
    """Box-constrained receding-horizon MPC of REPEATED interior-point solves — BASELINE config 3 as
    BASELINE.json words it ("receding-horizon loop of repeated IP solves"; an extension: the reference has no
    script for it).  Anchors: the loop of ref examples/linear_mpc_parallel.py:67-81 (solve from the current
    state, apply u[0], move to x[1]) around `par_interior_point_optimal_control` on the OCP of
    ref examples/linear_demo_cuda.py:19-55 with the dummy constraint `-1` (:30-31) replaced by the box
    |u| <= u_max (problems.make_linear_demo(control_bound=...)), horizon 40 (:51).  The first solve starts
    from u = 0.1 N(0,1) (the start of the reference's constrained examples, ref examples/cartpole_runtime.py:102-103)
    instead of the demo's u = 0 (:54): at u = 0 a symmetric box makes cu = 0, hence reg = rp*||cu|| = 0 (ref
    noc/par_interior_point_newton.py:116-117) for every rp, the full Newton step is re-tried 501 times and the
    reference's own algorithm walks out of the feasible set for good.

    Every MPC step: solve the horizon-`horizon` OCP from the current state, warm-started with the previous
    solution shifted by one step (last control repeated), apply the first control through the OCP's own
    dynamics.  -> (xs (sim_steps+1, nx), us (sim_steps, nu), iterations per step (list))."""
    from . import noc
    dev = x0.device
    if not x0.is_cuda:
        raise L.IpocError("constrained_mpc needs CUDA tensors; there is no CPU fallback")
    x = L.dev_f64(x0)
    nu = 1 if u_init is None else u_init.shape[-1]
    if u_init is None:
        import numpy as np
        u_init = torch.as_tensor(0.1 * np.random.default_rng(1).standard_normal((horizon, nu)), device=dev)
    u = L.dev_f64(u_init, dev).clone()
    xs, us, its = [x.clone()], [], []
    for _ in range(sim_steps):
        u_opt, n_it = noc.par_interior_point_optimal_control(ocp, u, x, use_graphs=use_graphs)
        x = ocp.dynamics(x, u_opt[0]).detach()           # plant = model (linear_mpc_parallel.py:69: x_par[1])
        xs.append(x.clone())
        us.append(u_opt[0].clone())
        its.append(int(n_it))
        u = torch.cat((u_opt[1:], u_opt[-1:])).contiguous()   # shifted warm start
    return torch.stack(xs), torch.stack(us), its

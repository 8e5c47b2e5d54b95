"""Multi-GPU modes (no reference counterpart — the reference is single-device).

time-sharded: a long horizon is cut into P contiguous segments, one per rank.  Each scan is
    local reduce  ->  all-gather of the P segment aggregates (NCCL over NVLink; <= 448 B per rank
    for nx = 4, i.e. latency-bound)  ->  local seeded scan.
    The exchange is injected as a callable so that the same code runs with torch.distributed
    (NCCL on the GPU box, gloo in CPU tests of the plumbing) or with P "virtual ranks" on one GPU.
batch-sharded: independent OCPs are split across ranks with no communication at all
    (`shard_batch`).
"""
import torch
from . import _lib as L


def carry_doubles(kind, nx):
    return L.lib().ipoc_carry_doubles(kind, nx)


class SegmentNewton:
    """One rank's share of a time-sharded Newton step (K2 + K3) on its contiguous segment."""

    def __init__(self, fx, fu, ru, Q, R, M, rank, nranks):
        self.fx, self.fu, self.ru, self.Q, self.R, self.M = (L.dev_f64(t) for t in (fx, fu, ru, Q, R, M))
        self.rank, self.nranks = rank, nranks
        self.N, self.nx = self.fx.shape[0], self.fx.shape[1]
        self.nu = self.fu.shape[-1]
        L.require_supported(self.nx, self.nu)
        self.dev = self.fx.device
        need = L.lib().ipoc_workspace_bytes(L.WS_NEWTON_STEP, self.N, self.nx, self.nu, 1)
        # private workspace: the leaf aggregates of phase 1 are reused by phase 2
        self.ws = torch.zeros(need, dtype=torch.uint8, device=self.dev)   # control block must start zeroed
        self.nbytes = need
        o = dict(dtype=torch.float64, device=self.dev)
        self.Kx = torch.empty(self.N, self.nu, self.nx, **o)
        self.d = torch.empty(self.N, self.nu, **o)
        self.dx = torch.empty(self.N + 1, self.nx, **o)
        self.du = torch.empty(self.N, self.nu, **o)
        self.pred = torch.empty(1, **o)
        self.feas = torch.empty(1, dtype=torch.int32, device=self.dev)

    # ---- CUDA-graph variant: the three local phases become three graph launches, the two exchanges stay
    #      ordinary NCCL calls between them (static carry buffers) — removes ~25 kernel launches of Python
    #      overhead per step, which otherwise hides the benefit of sharding at N <= 1e6.
    def capture(self, reg, ST):
        o = dict(dtype=torch.float64, device=self.dev)
        self.g_reg = L.dev_f64(reg, self.dev).reshape(1).clone()
        self.g_ST = L.dev_f64(ST, self.dev).clone()
        self.g_carries = torch.zeros(self.nranks, carry_doubles(L.CARRY_RICCATI, self.nx), **o)
        na = carry_doubles(L.CARRY_AFFINE, self.nx)
        self.g_fwd = torch.zeros(self.nranks, na + 2, **o)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):      # warm-up outside capture
            c = self.bwd_reduce(self.g_reg)
            self.g_carries[self.rank].copy_(c)
            f = self.bwd_apply(self.g_carries, self.g_ST)
            self.fwd_apply(self.g_fwd[:, :na].contiguous())
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.graph1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph1, stream=side):
            self.g_carry = self.bwd_reduce(self.g_reg)
        self.graph2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph2, pool=self.graph1.pool(), stream=side):
            fc = self.bwd_apply(self.g_carries, self.g_ST)
            self.g_scal = torch.cat((fc, self.pred, self.feas.to(torch.float64)))
        self.graph3 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph3, pool=self.graph1.pool(), stream=side):
            self.fwd_apply(self.g_fwd[:, :na].contiguous())
        return self

    def capture_one(self, reg, ST, all_gather_into):
        """The whole time-sharded Newton step — three local phases AND the two all-gathers of the carries — as
        ONE CUDA graph (NCCL collectives are graph-capturable): one launch per step on every rank instead of
        three graph launches and two host-issued collectives.  `step_one()` replays it."""
        o = dict(dtype=torch.float64, device=self.dev)
        self.g_reg = L.dev_f64(reg, self.dev).reshape(1).clone()
        self.g_ST = L.dev_f64(ST, self.dev).clone()
        na = carry_doubles(L.CARRY_AFFINE, self.nx)
        self.g_carries = torch.zeros(self.nranks, carry_doubles(L.CARRY_RICCATI, self.nx), **o)
        self.g_fwd = torch.zeros(self.nranks, na + 2, **o)
        self.g_mine = torch.zeros(na + 2, **o)

        def body():
            carry = self.bwd_reduce(self.g_reg)
            all_gather_into(self.g_carries, carry)
            fc = self.bwd_apply(self.g_carries, self.g_ST)
            self.g_mine[:na].copy_(fc)
            self.g_mine[na:na + 1].copy_(self.pred)
            self.g_mine[na + 1:].copy_(self.feas.to(torch.float64))
            all_gather_into(self.g_fwd, self.g_mine)
            self.fwd_apply(self.g_fwd[:, :na].contiguous())

        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):          # warm-up outside the capture (communicator, allocator, lazy inits)
            for _ in range(2):
                body()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            body()
        self.graph_one = g
        return self

    def step_one(self):
        """Replay of `capture_one`: -> (dx, du, pred_total, feasible_all)."""
        self.graph_one.replay()
        na = self.g_fwd.shape[1] - 2
        return self.dx, self.du, self.g_fwd[:, na].sum(), (self.g_fwd[:, na + 1] != 0).all()

    def step_graphed(self, all_gather_into):
        """One time-sharded Newton step from the captured graphs; `all_gather_into(out, t)` fills the
        (P, len) tensor `out` with every rank's `t`.  Returns (dx, du, pred_total, feasible_all)."""
        self.graph1.replay()
        all_gather_into(self.g_carries, self.g_carry)
        self.graph2.replay()
        all_gather_into(self.g_fwd, self.g_scal)
        self.graph3.replay()
        na = self.g_fwd.shape[1] - 2
        return self.dx, self.du, self.g_fwd[:, na].sum(), (self.g_fwd[:, na + 1] != 0).all()   # 0-d bool tensor

    def bwd_reduce(self, reg):
        self.reg = L.dev_f64(reg, self.dev).reshape(1)
        carry = torch.empty(carry_doubles(L.CARRY_RICCATI, self.nx), dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(L.lib().ipoc_newton_bwd_reduce_f64(
                self.N, self.nx, self.nu, L.ptr(self.fx), L.ptr(self.fu), L.ptr(self.ru), L.ptr(self.Q), L.ptr(self.R),
                L.ptr(self.M), L.ptr(self.reg), L.ptr(carry), L.ptr(self.ws), self.nbytes, L.stream_ptr()))
        return carry

    def bwd_apply(self, carries, ST):
        """carries (P, ESZ) gathered aggregates; ST (nx,nx) terminal weight of the whole horizon."""
        fwd_carry = torch.empty(carry_doubles(L.CARRY_AFFINE, self.nx), dtype=torch.float64, device=self.dev)
        carries, ST = L.dev_f64(carries, self.dev), L.dev_f64(ST, self.dev)
        with torch.cuda.device(self.dev):
            L.check(L.lib().ipoc_newton_bwd_apply_f64(
                self.N, self.nx, self.nu, self.rank, self.nranks, L.ptr(self.fx), L.ptr(self.fu), L.ptr(self.ru),
                L.ptr(self.Q), L.ptr(self.R), L.ptr(self.M), L.ptr(self.reg), L.ptr(carries), L.ptr(ST),
                L.ptr(self.Kx), L.ptr(self.d), L.ptr(self.pred), L.ptr(self.feas), L.ptr(fwd_carry), L.ptr(self.ws),
                self.nbytes, L.stream_ptr()))
        return fwd_carry

    def fwd_apply(self, fwd_carries):
        fwd_carries = L.dev_f64(fwd_carries, self.dev)
        with torch.cuda.device(self.dev):
            L.check(L.lib().ipoc_newton_fwd_apply_f64(
                self.N, self.nx, self.nu, self.rank, self.nranks, L.ptr(self.fx), L.ptr(self.fu), L.ptr(self.Kx),
                L.ptr(self.d), L.ptr(fwd_carries), L.ptr(self.dx), L.ptr(self.du), L.ptr(self.ws), self.nbytes,
                L.stream_ptr()))
        return self.dx, self.du


class SegmentPass:
    """One rank's share of a WHOLE time-sharded hot-path pass (K1 + K4 + K2 + K3 + K4) on its contiguous segment:

        K1 reduce + local max|ru|, sum cu^2        -> exchange 1: affine carry, the two scalars, lamT
        K1 apply  (costates of the segment), reg = rp * ||cu|| from the gathered sums (fixed rank order)
        K2 reduce                                   -> exchange 2: Riccati carry
        K2 apply  (gains, pred / feasibility partials, K3's segment aggregate)
                                                    -> exchange 3: forward carry, pred, feasible[, constraints ok]
        K3 apply  (dx, du of the segment)

    i.e. the three dependent carry exchanges of SURVEY section 8(e), each a (P, len) all-gather of < 0.5 KB per
    rank.  Every local phase is one or two launches of libipoc.so (the scan levels complete inside the leaf
    kernels and the seeds are chained through the gathered carries inside the down-sweeps).  `step()` enqueues the
    whole pass on the current stream; `capture()` records it — collectives included — into ONE CUDA graph."""

    def __init__(self, fx, fu, cx, cu, lamT, ru, Q, R, M, rank, nranks, cons=None, rp=1.0):
        self.new = SegmentNewton(fx, fu, ru, Q, R, M, rank, nranks)
        n = self.new
        self.rank, self.nranks, self.dev, self.N, self.nx, self.nu = rank, nranks, n.dev, n.N, n.nx, n.nu
        self.cx, self.cu = L.dev_f64(cx), L.dev_f64(cu)
        self.cons = None if cons is None else L.dev_f64(cons)
        o = dict(dtype=torch.float64, device=self.dev)
        self.na = carry_doubles(L.CARRY_AFFINE, self.nx)
        self.nr = carry_doubles(L.CARRY_RICCATI, self.nx)
        P = nranks
        # exchange buffers: row r = rank r's contribution
        self.x1_mine, self.x1 = torch.zeros(self.na + 2 + self.nx, **o), torch.zeros(P, self.na + 2 + self.nx, **o)
        self.x2_mine, self.x2 = torch.zeros(self.nr, **o), torch.zeros(P, self.nr, **o)
        self.x3_mine, self.x3 = torch.zeros(self.na + 3, **o), torch.zeros(P, self.na + 3, **o)
        self.x1_mine[self.na + 2:].copy_(L.dev_f64(lamT, self.dev).reshape(-1))   # only the last rank's is used
        self.lam = torch.empty(self.N + 1, self.nx, **o)
        self.rp = torch.full((1,), float(rp), **o)
        self.reg = torch.empty(1, **o)
        self.hu, self.cu_norm = torch.empty(1, **o), torch.empty(1, **o)
        self.hu_l, self.cn_l = torch.empty(1, **o), torch.empty(1, **o)
        self.feas_l = torch.ones(1, dtype=torch.int32, device=self.dev)
        lib = L.lib()
        self.nb_aff = lib.ipoc_workspace_bytes(L.WS_AFFINE_SCAN, self.N, self.nx, self.nu, 1)
        self.ws_aff = torch.zeros(self.nb_aff, dtype=torch.uint8, device=self.dev)
        nc = 1 if self.cons is None else self.cons.shape[-1]
        self.nc = nc
        self.nb_red = lib.ipoc_workspace_bytes(L.WS_REDUCTIONS, self.N, max(self.nu, nc), self.nu, 1)
        self.ws_red = torch.zeros(self.nb_red, dtype=torch.uint8, device=self.dev)
        self.graph = None

    # ---- local phases (C ABI calls only) -----------------------------------------------------------------
    def _k1_reduce(self):
        p, lib, n = L.ptr, L.lib(), self.new
        with torch.cuda.device(self.dev):
            L.check(lib.ipoc_affine_reduce_f64(1, 1, self.N, self.nx, p(n.fx), p(self.cx), p(self.x1_mine),
                                               p(self.ws_aff), self.nb_aff, L.stream_ptr()))
            L.check(lib.ipoc_reductions_f64(self.N, self.nu, self.nc, 1, p(n.ru), p(self.cu), None, p(self.hu_l),
                                            p(self.cn_l), None, None, None, p(self.ws_red), self.nb_red,
                                            L.stream_ptr()))
        self.x1_mine[self.na:self.na + 1].copy_(self.cn_l * self.cn_l)      # sum of squares travels, not the norm
        self.x1_mine[self.na + 1:self.na + 2].copy_(self.hu_l)

    def _k1_apply(self):
        p, lib, n, na = L.ptr, L.lib(), self.new, self.na
        carries = self.x1[:, :na].contiguous()
        seed = self.x1[self.nranks - 1, na + 2:].contiguous()               # lamT lives on the last rank
        with torch.cuda.device(self.dev):
            L.check(lib.ipoc_affine_apply_f64(1, 1, self.N, self.nx, self.rank, self.nranks, p(n.fx), p(self.cx),
                                              p(carries), p(seed), p(self.lam), p(self.ws_aff), self.nb_aff,
                                              L.stream_ptr()))
        self.cu_norm.copy_(self.x1[:, na].sum().sqrt().reshape(1))          # fixed rank order (ref :116)
        self.hu.copy_(self.x1[:, na + 1].max().reshape(1))                  # (ref :158)
        self.reg.copy_(self.rp * self.cu_norm)                              # (ref :117)

    def _k2_reduce(self):
        self.x2_mine.copy_(self.new.bwd_reduce(self.reg))

    def _k2_apply(self, ST):
        n, na = self.new, self.na
        fc = n.bwd_apply(self.x2, ST)
        self.x3_mine[:na].copy_(fc)
        self.x3_mine[na:na + 1].copy_(n.pred)
        self.x3_mine[na + 1:na + 2].copy_(n.feas.to(torch.float64))
        if self.cons is not None:
            p, lib = L.ptr, L.lib()
            with torch.cuda.device(self.dev):
                L.check(lib.ipoc_reductions_f64(self.N, self.nu, self.nc, 1, None, None, p(self.cons), None, None,
                                                p(self.feas_l), None, None, p(self.ws_red), self.nb_red,
                                                L.stream_ptr()))
        self.x3_mine[na + 2:na + 3].copy_(self.feas_l.to(torch.float64))

    def _k3_apply(self):
        self.new.fwd_apply(self.x3[:, :self.na].contiguous())

    # ---- the collective pass -------------------------------------------------------------------------------
    def step(self, all_gather_into, ST):
        """Enqueue one pass.  `all_gather_into(out (P, len), mine (len))` fills `out` with every rank's `mine`.
        Results: lam, new.Kx, new.d, new.dx, new.du (segment), hu, cu_norm, and — after `scalars()` — the
        horizon-wide pred / feasibility flags."""
        self._k1_reduce()
        all_gather_into(self.x1, self.x1_mine)          # exchange 1
        self._k1_apply()
        self._k2_reduce()
        all_gather_into(self.x2, self.x2_mine)          # exchange 2
        self._k2_apply(ST)
        all_gather_into(self.x3, self.x3_mine)          # exchange 3
        self._k3_apply()

    def scalars(self):
        """-> (pred (sum over ranks, fixed order), bwd_feasible (all ranks), traj_feasible (all ranks))."""
        na = self.na
        return (self.x3[:, na].sum(), bool((self.x3[:, na + 1] != 0).all()), bool((self.x3[:, na + 2] != 0).all()))

    def capture(self, all_gather_into, ST):
        """Record the whole pass, its three all-gathers included, into ONE CUDA graph (NCCL collectives are
        graph-capturable); `replay()` then costs a single launch per pass on every rank."""
        self.g_ST = L.dev_f64(ST, self.dev).clone()
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):          # warm-up outside the capture (allocator, communicator, lazy inits)
            for _ in range(2):
                self.step(all_gather_into, self.g_ST)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self.step(all_gather_into, self.g_ST)
        self.graph = g
        return self

    def replay(self):
        self.graph.replay()


def pass_virtual_ranks(fx, fu, cx, cu, lamT, ru, Q, R, M, nranks, cons=None, rp=1.0):
    """The time-sharded PASS with `nranks` virtual ranks on one GPU (fake all-gather between the phases) —
    the single-GPU test of `SegmentPass`.  -> dict of horizon-wide results."""
    N = fx.shape[0]
    bounds = segment_bounds(N, nranks)
    segs = [SegmentPass(fx[lo:hi], fu[lo:hi], cx[lo:hi], cu[lo:hi], lamT, ru[lo:hi], Q[lo:hi], R[lo:hi], M[lo:hi], r,
                        nranks, None if cons is None else cons[lo:hi], rp) for r, (lo, hi) in enumerate(bounds)]
    ST = Q[0].contiguous()
    for s in segs:
        s._k1_reduce()
    x1 = torch.stack([s.x1_mine for s in segs])
    for s in segs:
        s.x1.copy_(x1)
        s._k1_apply()
        s._k2_reduce()
    x2 = torch.stack([s.x2_mine for s in segs])
    for s in segs:
        s.x2.copy_(x2)
        s._k2_apply(ST)
    x3 = torch.stack([s.x3_mine for s in segs])
    for s in segs:
        s.x3.copy_(x3)
        s._k3_apply()
    pred, bwd_feas, traj_feas = segs[0].scalars()
    return dict(lam=torch.cat([s.lam[:-1] for s in segs[:-1]] + [segs[-1].lam]),
                dx=torch.cat([s.new.dx[:-1] for s in segs[:-1]] + [segs[-1].new.dx]),
                du=torch.cat([s.new.du for s in segs]), Kx=torch.cat([s.new.Kx for s in segs]),
                d=torch.cat([s.new.d for s in segs]), pred=pred, bwd_feasible=bwd_feas, traj_feasible=traj_feas,
                hu=segs[0].hu.clone(), cu_norm=segs[0].cu_norm.clone())


def newton_step_time_sharded(seg: SegmentNewton, reg, ST, all_gather):
    """Collective Newton step.  `all_gather(t)` returns the (P, len) stack of every rank's `t`.
    `ST`: terminal weight (= Q[0] of the global horizon, ref noc/par_interior_point_newton.py:73),
    identical on all ranks.  Returns this rank's (dx, du, pred_total, feasible_all)."""
    carries = all_gather(seg.bwd_reduce(reg))                    # exchange 1: Riccati aggregates
    fwd_carry = seg.bwd_apply(carries, ST)
    scal = torch.cat((fwd_carry, seg.pred, seg.feas.to(torch.float64)))
    gathered = all_gather(scal)                                  # exchange 2: forward aggregates + scalars
    na = fwd_carry.numel()
    dx, du = seg.fwd_apply(gathered[:, :na].contiguous())
    pred = gathered[:, na].sum()                                 # fixed rank order -> deterministic
    feas = bool(torch.all(gathered[:, na + 1] != 0))
    return dx, du, pred, feas


def dist_all_gather(group=None):
    """all_gather callable on torch.distributed (NCCL on GPUs, gloo for CPU plumbing tests)."""
    import torch.distributed as dist

    def gather(t):
        P = dist.get_world_size(group)
        flat = t.contiguous().reshape(-1)
        out = torch.empty(P * flat.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, flat, group=group)
        return out.view((P,) + tuple(t.shape))

    return gather


def dist_all_gather_into(group=None):
    """`all_gather_into(out (P, len), t (len))` on torch.distributed with preallocated output."""
    import torch.distributed as dist

    def gather(out, t):
        dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=group)

    return gather


def segment_bounds(N, nranks, allow_empty=False):
    """Contiguous, near-equal segments [lo, hi) per rank.  Time segments must be non-empty (a rank without steps
    would reach the C ABI as N = 0); batch shards may be empty."""
    if nranks < 1 or (N < nranks and not allow_empty):
        raise ValueError(f"cannot cut a horizon of {N} steps into {nranks} non-empty contiguous segments")
    base, rem = divmod(N, nranks)
    bounds, lo = [], 0
    for r in range(nranks):
        hi = lo + base + (1 if r < rem else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def shard_batch(batch, rank, nranks):
    """[lo, hi) of the independent problems owned by `rank` (no data-path collective needed)."""
    return segment_bounds(batch, nranks, allow_empty=True)[rank]


def newton_step_virtual_ranks(fx, fu, ru, Q, R, M, reg, nranks):
    """Run the time-sharded algorithm with `nranks` virtual ranks on ONE GPU (sequentially, fake
    all-gather) — the single-GPU test of the multi-GPU path."""
    N = fx.shape[0]
    segs = [SegmentNewton(fx[lo:hi], fu[lo:hi], ru[lo:hi], Q[lo:hi], R[lo:hi], M[lo:hi], r, nranks)
            for r, (lo, hi) in enumerate(segment_bounds(N, nranks))]
    ST = Q[0]
    carries = torch.stack([s.bwd_reduce(reg) for s in segs])
    fwd = torch.stack([s.bwd_apply(carries, ST) for s in segs])
    outs = [s.fwd_apply(fwd) for s in segs]
    dx = torch.cat([o[0][:-1] for o in outs[:-1]] + [outs[-1][0]])
    du = torch.cat([o[1] for o in outs])
    pred = torch.stack([s.pred[0] for s in segs]).sum()
    feas = all(bool(s.feas[0] != 0) for s in segs)
    Kx = torch.cat([s.Kx for s in segs])
    d = torch.cat([s.d for s in segs])
    return dx, du, Kx, d, pred, feas

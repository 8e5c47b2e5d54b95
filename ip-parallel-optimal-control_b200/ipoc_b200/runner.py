"""Pre-allocated, allocation-free driver of one hot-path pass (K1 + K4 + K2 + K3 + K4 + A8) on fixed
device buffers — what `newton_oc` does per accept/reject attempt, minus the user functions.
Every call only enqueues kernels of libipoc.so on the current stream, so a pass can be captured
into a CUDA graph (`capture()`) and replayed with one launch."""
import ctypes
import torch
from . import _lib as L


class NewtonPass:
    def __init__(self, fx, fu, cx, cu, lamT, ru, Q, R, M, cons=None, rp=1.0, out=None):
        """Tensors (N,...) for one problem or (B,N,...) for a batch of independent problems.
        `out`: optional dict of preallocated result buffers (see HostNewtonPass)."""
        f = L.dev_f64
        batched = fx.dim() == 4
        up = (lambda t: f(t)) if batched else (lambda t: f(t).unsqueeze(0))
        self.fx, self.fu, self.cx, self.cu, self.ru, self.Q, self.R, self.M = (up(t) for t in
                                                                               (fx, fu, cx, cu, ru, Q, R, M))
        self.lamT = f(lamT).reshape(-1, self.fx.shape[2]).contiguous()
        self.cons = None if cons is None else up(cons)
        self.B, self.N, self.nx = self.fx.shape[0], self.fx.shape[1], self.fx.shape[2]
        self.nu = self.fu.shape[-1]
        self.nc = 1 if self.cons is None else self.cons.shape[-1]
        L.require_supported(self.nx, self.nu)
        dev = self.dev = self.fx.device
        o = dict(dtype=torch.float64, device=dev)
        B, N, nx, nu = self.B, self.N, self.nx, self.nu
        out = out or {}
        i32 = dict(dtype=torch.int32, device=dev)
        self.lam = out["lam"] if "lam" in out else torch.empty(B, N + 1, nx, **o)
        self.dx = out["dx"] if "dx" in out else torch.empty(B, N + 1, nx, **o)
        self.du = out["du"] if "du" in out else torch.empty(B, N, nu, **o)
        self.Kx = torch.empty(B, N, nu, nx, **o)
        self.d = torch.empty(B, N, nu, **o)
        self.pred = out["pred"] if "pred" in out else torch.empty(B, **o)
        self.bwd_feas = out["bwd_feas"] if "bwd_feas" in out else torch.empty(B, **i32)
        self.hu = out["hu"] if "hu" in out else torch.zeros(B, **o)
        self.cu_norm = out["cu_norm"] if "cu_norm" in out else torch.zeros(B, **o)
        self.traj_feas = out["traj_feas"] if "traj_feas" in out else torch.ones(B, **i32)
        self.rp = out["rp"] if "rp" in out else torch.empty(B, **o)
        self.rp.fill_(float(rp))
        self.r_inc = out["r_inc"] if "r_inc" in out else torch.empty(B, **o)
        self.r_inc.fill_(2.0)
        self.reg = torch.empty(B, **o)
        self.cost = torch.full((B,), 1.0, **o)
        self.new_cost = torch.full((B,), 0.5, **o)
        self.success = out["success"] if "success" in out else torch.zeros(B, **i32)
        self.gain = out["gain"] if "gain" in out else torch.zeros(B, **o)
        lib = L.lib()
        self.ws_aff_bytes = lib.ipoc_workspace_bytes(L.WS_AFFINE_SCAN, N, nx, nu, B)
        self.ws_new_bytes = lib.ipoc_workspace_bytes(L.WS_NEWTON_STEP, N, nx, nu, B)
        self.ws_red_bytes = lib.ipoc_workspace_bytes(L.WS_REDUCTIONS, N, max(nu, self.nc), nu, B)
        self.ws_red = torch.empty(self.ws_red_bytes, dtype=torch.uint8, device=dev)
        self.ws_aff = torch.zeros(self.ws_aff_bytes, dtype=torch.uint8, device=dev)
        self.ws_new = torch.zeros(self.ws_new_bytes, dtype=torch.uint8, device=dev)
        # fused entry points (K1 + ||cu||; K2 + K3 + max|ru| + constraints + accept): their own workspaces
        self.ws_cos_bytes = lib.ipoc_workspace_bytes(L.WS_COSTATES, N, nx, nu, B)
        self.ws_att_bytes = lib.ipoc_workspace_bytes(L.WS_NEWTON_ATTEMPT, N, nx, max(nu, self.nc), B)
        self.ws_cos = torch.zeros(self.ws_cos_bytes, dtype=torch.uint8, device=dev)
        self.ws_att = torch.zeros(self.ws_att_bytes, dtype=torch.uint8, device=dev)
        self.fused = True
        self.graph = None

    # algorithmic bytes per pass: each phase reads its inputs once, writes outputs once
    def algorithmic_bytes(self):
        nx, nu, n = self.nx, self.nu, self.N * self.B
        k1 = 8 * (nx * nx + 2 * nx) * n
        k2 = 8 * (2 * nx * nx + 3 * nx * nu + nu * nu + 2 * nu) * n
        k3 = 8 * (nx * nx + 2 * nx * nu + nx + 2 * nu) * n
        k4 = 8 * (2 * nu + self.nc) * n
        return dict(K1=k1, K2=k2, K3=k3, K4=k4, total=k1 + k2 + k3 + k4)

    def costates(self):
        p, lib = L.ptr, L.lib()
        L.check(lib.ipoc_affine_scan_f64(1, 1, self.N, self.nx, self.B, p(self.fx), p(self.cx), p(self.lamT),
                                         p(self.lam), p(self.ws_aff), self.ws_aff_bytes, L.stream_ptr()))

    def newton(self):
        p, lib = L.ptr, L.lib()
        L.check(lib.ipoc_newton_step_f64(self.N, self.nx, self.nu, self.B, p(self.fx), p(self.fu), p(self.ru),
                                         p(self.Q), p(self.R), p(self.M), p(self.reg), p(self.dx), p(self.du),
                                         p(self.Kx), p(self.d), p(self.pred), p(self.bwd_feas), p(self.ws_new),
                                         self.ws_new_bytes, L.stream_ptr()))

    def reductions_ru_cu(self):
        p, lib = L.ptr, L.lib()
        L.check(lib.ipoc_reductions_f64(self.N, self.nu, self.nc, self.B, p(self.ru), p(self.cu), None, p(self.hu),
                                        p(self.cu_norm), None, p(self.rp), p(self.reg), p(self.ws_red),
                                        self.ws_red_bytes, L.stream_ptr()))

    def feasibility_and_accept(self):
        p, lib, s = L.ptr, L.lib(), L.stream_ptr()
        if self.cons is not None:
            L.check(lib.ipoc_reductions_f64(self.N, self.nu, self.nc, self.B, None, None, p(self.cons), None, None,
                                            p(self.traj_feas), None, None, p(self.ws_red), self.ws_red_bytes, s))
        L.check(lib.ipoc_accept_update_f64(self.B, p(self.cost), p(self.new_cost), p(self.traj_feas), p(self.pred),
                                           p(self.bwd_feas), None, p(self.rp), p(self.r_inc), p(self.success),
                                           p(self.gain), s))

    def costates_fused(self):
        """K1 with ||cu||_F as a side job of the up-sweep (ipoc_costates_f64): two launches."""
        p, lib = L.ptr, L.lib()
        L.check(lib.ipoc_costates_f64(self.N, self.nx, self.nu, self.B, p(self.fx), p(self.cx), p(self.lamT), p(self.cu),
                                      p(self.lam), p(self.cu_norm), None, p(self.ws_cos), self.ws_cos_bytes, L.stream_ptr()))

    def attempt_fused(self):
        """K2 + K3 with reg = rp*||cu|| formed in the kernel, max|ru| folded into the up-sweep, the constraint
        reduction and the accept / regularisation update folded into K3's leaf kernel
        (ipoc_newton_attempt_f64): three launches."""
        p, lib = L.ptr, L.lib()
        L.check(lib.ipoc_newton_attempt_f64(
            self.N, self.nx, self.nu, self.nc, self.B, p(self.fx), p(self.fu), p(self.ru), p(self.Q), p(self.R),
            p(self.M), p(self.rp), p(self.cu_norm), p(self.dx), p(self.du), p(self.Kx), p(self.d), p(self.pred),
            p(self.bwd_feas), p(self.hu), None, None, None, None,
            p(self.cons), p(self.traj_feas) if self.cons is not None else None,
            p(self.cost), p(self.new_cost), p(self.traj_feas), None, p(self.rp), p(self.r_inc), p(self.success),
            p(self.gain), p(self.ws_att), self.ws_att_bytes, L.stream_ptr()))

    def run_unfused(self):
        """The same pass as four separate C-ABI calls (round-1 sequence, 11 launches): K1, K4, K2+K3, K4 + A8."""
        self.costates()
        self.reductions_ru_cu()
        self.newton()
        self.feasibility_and_accept()

    def run(self):
        """K1 costates, K4 (max|ru|, ||cu||), reg = rp*||cu||, K2+K3 Newton step, K4 (constraints of
        the stepped trajectory), A8 accept/update (rp, r_inc evolve on the device from pass to pass; the
        work per pass does not depend on their values).  Only kernels of libipoc.so are launched: five with
        the fused entry points (default), eleven as separate calls (`fused = False`)."""
        if self.fused:
            self.costates_fused()
            self.attempt_fused()
        else:
            self.run_unfused()

    def capture(self):
        """Capture one pass into a CUDA graph (every C-ABI call is enqueue-only)."""
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self.run()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self.run()
        self.graph = g
        return g

    def replay(self):
        self.graph.replay()

    def launches_per_pass(self):
        lib = L.lib()
        torch.cuda.synchronize(self.dev)
        a = lib.ipoc_launch_count()
        self.run()
        torch.cuda.synchronize(self.dev)
        return int(lib.ipoc_launch_count() - a)

    def profile(self, fn=None):
        """Per-launch device durations (ms) of one pass via the library's CUDA-event profiler."""
        lib = L.lib()
        L.check(lib.ipoc_profile_begin(L.stream_ptr()))
        (fn or self.run)()
        names = ctypes.create_string_buffer(16384)
        ms = (ctypes.c_float * 256)()
        n = lib.ipoc_profile_end(names, 16384, ms, 256)
        if n < 0:
            L.check(n)
        return list(zip(names.value.decode().rstrip(",").split(","), [float(ms[i]) for i in range(n)]))



class HostNewtonPass:
    """The same pass, end to end from HOST memory: every input of the pass lives in ONE pinned host arena
    and every result comes back into ONE pinned host arena, so a step is a single host->device copy, the
    kernels of `NewtonPass.run()` and a single device->host copy — all three captured into one CUDA graph
    (`capture()`), i.e. one launch per step.  The input arena is ordered by first use
    (fx, cx, lamT, cu | ru, fu, Q, R, M, cons); `views` exposes the named host views an integrator fills."""
    IN = ("fx", "cx", "lamT", "cu", "ru", "fu", "Q", "R", "M", "cons")
    OUT_F64 = ("lam", "dx", "du", "pred", "hu", "cu_norm", "gain", "rp", "r_inc")
    OUT_I32 = ("bwd_feas", "traj_feas", "success")

    @staticmethod
    def _arena(shapes, dev):
        offs, total = {}, 0
        for k, shp in shapes.items():
            n = 1
            for d in shp:
                n *= d
            offs[k] = (total, n, shp)
            total += n + (n & 1)                      # keep every view 16-byte aligned
        host = torch.empty(total, dtype=torch.float64).pin_memory()
        devb = torch.empty(total, dtype=torch.float64, device=dev)
        hv = {k: host[o:o + n].view(shp) for k, (o, n, shp) in offs.items()}
        dv = {k: devb[o:o + n].view(shp) for k, (o, n, shp) in offs.items()}
        return host, devb, hv, dv

    def __init__(self, inputs, device):
        """inputs: dict name -> tensor ((N,...) single problem), any device; copied once into the host arena."""
        dev = torch.device(device)
        shapes = {k: tuple(inputs[k].shape) for k in self.IN}
        self.h_in, self.d_in, self.views, dv = self._arena(shapes, dev)
        for k in self.IN:
            self.views[k].copy_(inputs[k].detach().to("cpu", torch.float64))
        N, nx = shapes["fx"][0], shapes["fx"][1]
        nu = shapes["fu"][-1]
        oshape = {"lam": (1, N + 1, nx), "dx": (1, N + 1, nx), "du": (1, N, nu)}
        oshape.update({k: (1,) for k in self.OUT_F64[3:]})
        oshape.update({k: (1,) for k in self.OUT_I32})           # one 8-byte slot each, used as int32
        self.h_out, self.d_out, self.results, do = self._arena(oshape, dev)
        for k in self.OUT_I32:
            do[k] = do[k].view(torch.int32)[:1]
            self.results[k] = self.results[k].view(torch.int32)[:1]
        self.d_in.copy_(self.h_in)                                # first fill (warm-up runs read it)
        self.inner = NewtonPass(dv["fx"], dv["fu"], dv["cx"], dv["cu"], dv["lamT"], dv["ru"], dv["Q"], dv["R"],
                                dv["M"], dv["cons"], out=do)
        for k in ("fx", "fu", "cx", "cu", "ru", "Q", "R", "M", "cons"):   # the pass must read the arena itself
            assert getattr(self.inner, k).data_ptr() == dv[k].data_ptr(), k
        off = lambda name, arena_views, base: (arena_views[name].data_ptr() - base.data_ptr()) // 8
        self.cut_a = off("ru", self.views, self.h_in)        # [fx cx lamT cu] | [ru fu Q R M cons]
        self.cut_lam = off("dx", self.results, self.h_out)   # [lam] | [dx du scalars]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ev_a, self.ev_b, self.ev_lam = (torch.cuda.Event() for _ in range(3))
        self.h2d_bytes = self.h_in.numel() * 8
        self.d2h_bytes = self.h_out.numel() * 8
        self.graph = None

    def run(self):
        """Copies on a second stream, ordered by first use, so that K1 runs while the inputs of K2 are still
        arriving and the costates go back while K2/K3 compute (PCIe is full duplex):
           copy stream : H2D [fx cx lamT cu] | H2D [ru fu Q R M cons] ............ D2H [lam]
           main stream :        wait A -> K1 (+ ||cu||) | wait B -> K2+K3 (+ K4, A8) | D2H [dx du scalars]"""
        main = torch.cuda.current_stream(self.d_in.device)
        cs = self.copy_stream
        a = self.cut_a
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            self.d_in[:a].copy_(self.h_in[:a], non_blocking=True)
            self.ev_a.record(cs)
            self.d_in[a:].copy_(self.h_in[a:], non_blocking=True)
            self.ev_b.record(cs)
        main.wait_event(self.ev_a)
        self.inner.costates_fused()
        self.ev_lam.record(main)
        with torch.cuda.stream(cs):
            cs.wait_event(self.ev_lam)
            self.h_out[:self.cut_lam].copy_(self.d_out[:self.cut_lam], non_blocking=True)
        main.wait_event(self.ev_b)
        self.inner.attempt_fused()
        self.h_out[self.cut_lam:].copy_(self.d_out[self.cut_lam:], non_blocking=True)
        main.wait_stream(cs)

    def capture(self):
        dev = self.d_in.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self.run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self.run()
        self.graph = g
        return g

    def replay(self):
        self.graph.replay()

"""Pre-allocated, allocation-free driver of one hot-path pass (K1 + K4 + K2 + K3 + K4 + A8) on fixed
device buffers — what `newton_oc` does per accept/reject attempt, minus the user functions.
Every call only enqueues kernels of libipoc.so on the current stream, so a pass can be captured
into a CUDA graph (`capture()`) and replayed with one launch."""
import ctypes
import torch
from . import _lib as L


class NewtonPass:
    def __init__(self, fx, fu, cx, cu, lamT, ru, Q, R, M, cons=None, rp=1.0):
        """Tensors (N,...) for one problem or (B,N,...) for a batch of independent problems."""
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
        self.lam = torch.empty(B, N + 1, nx, **o)
        self.dx = torch.empty(B, N + 1, nx, **o)
        self.du = torch.empty(B, N, nu, **o)
        self.Kx = torch.empty(B, N, nu, nx, **o)
        self.d = torch.empty(B, N, nu, **o)
        self.pred = torch.empty(B, **o)
        self.bwd_feas = torch.empty(B, dtype=torch.int32, device=dev)
        self.hu = torch.zeros(B, **o)
        self.cu_norm = torch.zeros(B, **o)
        self.traj_feas = torch.ones(B, dtype=torch.int32, device=dev)
        self.rp = torch.full((B,), float(rp), **o)
        self.r_inc = torch.full((B,), 2.0, **o)
        self.reg = torch.empty(B, **o)
        self.cost = torch.full((B,), 1.0, **o)
        self.new_cost = torch.full((B,), 0.5, **o)
        self.success = torch.zeros(B, dtype=torch.int32, device=dev)
        self.gain = torch.zeros(B, **o)
        lib = L.lib()
        self.ws_aff_bytes = lib.ipoc_workspace_bytes(L.WS_AFFINE_SCAN, N, nx, nu, B)
        self.ws_new_bytes = lib.ipoc_workspace_bytes(L.WS_NEWTON_STEP, N, nx, nu, B)
        self.ws_red_bytes = lib.ipoc_workspace_bytes(L.WS_REDUCTIONS, N, max(nu, self.nc), nu, B)
        self.ws_red = torch.empty(self.ws_red_bytes, dtype=torch.uint8, device=dev)
        self.ws_aff = torch.empty(self.ws_aff_bytes, dtype=torch.uint8, device=dev)
        self.ws_new = torch.empty(self.ws_new_bytes, dtype=torch.uint8, device=dev)
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

    def run(self):
        """K1 costates, K4 (max|ru|, ||cu||), reg = rp*||cu||, K2+K3 Newton step, K4 (constraints of
        the stepped trajectory), A8 accept/update (rp, r_inc evolve on the device from pass to pass; the
        work per pass does not depend on their values).  Only kernels of libipoc.so are launched."""
        p, lib, s = L.ptr, L.lib(), L.stream_ptr()
        self.costates()
        L.check(lib.ipoc_reductions_f64(self.N, self.nu, self.nc, self.B, p(self.ru), p(self.cu), None, p(self.hu),
                                        p(self.cu_norm), None, p(self.rp), p(self.reg), p(self.ws_red), self.ws_red_bytes, s))
        self.newton()
        if self.cons is not None:
            L.check(lib.ipoc_reductions_f64(self.N, self.nu, self.nc, self.B, None, None, p(self.cons), None, None,
                                            p(self.traj_feas), None, None, p(self.ws_red), self.ws_red_bytes, s))
        L.check(lib.ipoc_accept_update_f64(self.B, p(self.cost), p(self.new_cost), p(self.traj_feas), p(self.pred),
                                           p(self.bwd_feas), None, p(self.rp), p(self.r_inc), p(self.success),
                                           p(self.gain), s))

    def capture(self):
        """Capture one pass into a CUDA graph (every C-ABI call is enqueue-only)."""
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self.run()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
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

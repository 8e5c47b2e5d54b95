#!/usr/bin/env python
"""bench.py — headline benchmark of the par IP-Newton hot path.

A "step" is ONE pass of the hot path over one Newton iteration of the workload:
    K1 costate scan + K4 reductions (max|ru|, ||cu||, reg) + K2 Riccati scan/gains + K3 forward scan
    + K4 constraint reduction + A8 accept/regularisation update          (all kernels of libipoc.so)
on the LQ tensors of BASELINE.json configs[1]: constrained cartpole, nx=4, nu=1, N=10^4, float64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`value` : whole-job Newton steps/s with inputs resident in HBM (CUDA-graph replay of one pass, L2
          flushed between timed steps); N > 1: every rank runs its own independent OCP (batch-sharded
          mode, no data-path collective) -> weak scaling, value = sum over ranks.
`e2e`   : the same pass from HOST memory (runner.HostNewtonPass): all inputs in one pinned host arena,
          H2D + the pass's kernels + D2H of the results inside the timed region (one CUDA graph); the
          K2+K3-only host-buffer C call (ipoc_newton_step_host_f64) is reported next to it.
`--impl reference`: the reference path on the host CPU.  The reference itself (JAX + paroc) cannot be
          installed here, so this times the NumPy oracle port (oracle/), labelled kind="port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "ip-parallel-optimal-control_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

NX, NU, NC = 4, 1, 2
N_HEADLINE = 10_000
METRIC = "ip_newton_step_throughput_cartpole_N1e4"
UNIT = "newton_steps/s"


def alg_bytes(N, nx=NX, nu=NU, nc=NC):
    k1 = 8 * (nx * nx + 2 * nx) * N
    k2 = 8 * (2 * nx * nx + 3 * nx * nu + nu * nu + 2 * nu) * N
    k3 = 8 * (nx * nx + 2 * nx * nu + nx + 2 * nu) * N
    k4 = 8 * (2 * nu + nc) * N
    return dict(K1=k1, K2=k2, K3=k3, K4=k4, total=k1 + k2 + k3 + k4)


# DRAM bytes (read + write) per pass of each phase at the headline workload, from the committed ncu capture
# profiles/r02_launches_bench_ncu.csv (dram__bytes_read.sum + dram__bytes_write.sum, median per kernel, summed over
# the phase's launches; writes mostly stay in L2 at this size).  Offline measurement, as the bench contract asks.
NCU_TRAFFIC_N1E4 = {"K1": 3.80e6, "K2": 7.87e6, "K3": 2.62e6}

PHASE_OF = {
    "k_aff_seed": "K1", "k_aff_leaf_up": "K1", "k_aff_leaf_down": "K1",
    "k_ric_seed": "K2", "k_ric_leaf_up": "K2", "k_mid_up_ric": "K2", "k_top_ric": "K2", "k_mid_down_ric": "K2",
    "k_ric_leaf_down": "K2", "k_finalize_pred": "K2",
    "k_fwd_leaf_down": "K3", "k_fwd_leaf_up": "K3",
    "k_reduce_partial": "K4", "k_reduce_final": "K4", "k_reduce_single": "K4", "ipoc_reductions_f64": "K4", "ipoc_accept_update_f64": "K4",
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ host placement
def bind_to_gpu_numa(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index` BEFORE any pinned host memory is
    allocated (first touch then places the staging arenas on the GPU's NUMA node; a remote node costs
    the host<->device copies of the e2e number more than 2x on these boxes).  Returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return "nvml affinity empty or outside the cpuset; unchanged"
        os.sched_setaffinity(0, allowed)
        return f"bound to {len(allowed)} CPUs local to GPU {index} ({allowed[0]}..{allowed[-1]})"
    except Exception as e:   # measurement hygiene only
        return f"unchanged ({type(e).__name__})"


def h2d_probe(dev, mb=64):
    """Pinned host -> device copy bandwidth (GB/s) as seen by this process: context for the e2e number."""
    import torch
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(dev)
    a.record()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize(dev)
    return 4 * (mb << 20) / (a.elapsed_time(b) * 1e-3) / 1e9


# ------------------------------------------------------------------------------ clocks sampler
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------ reference arm (CPU)
def cpu_inputs(N, seed=1):
    """Same workload as the GPU arm, computed entirely on the CPU (oracle evaluator)."""
    from ipoc_b200 import problems, workloads
    from oracle.autodiff import Evaluator
    from oracle import noc_np
    Ts = 1.0 / N
    ocp = problems.make_cartpole(Ts)
    x0 = problems.cartpole_x0().numpy()
    u0 = 0.1 * np.random.default_rng(seed).standard_normal((N, 1))
    xs = workloads.cartpole_rollout_np(u0, x0, Ts)
    ev = Evaluator(ocp)
    d = ev.derivatives(xs, u0, 0.1)
    lamT = ev.final_cost_grad(xs[-1])
    lam = noc_np.par_costates(lamT, d)
    ru, Q, R, M = noc_np.compute_lqr_params(lam, d)
    return dict(d=d, lamT=lamT, ru=ru, Q=Q, R=R, M=M, cons=ev.constraints(xs, u0))


def cpu_pass(inp):
    """One hot-path pass with the NumPy oracle (ref noc/costates.py:34-40, par_Newton :107-124,
    reductions :45-47,:158)."""
    from oracle import noc_np
    d = inp["d"]
    noc_np.par_costates(inp["lamT"], d)
    hu = np.max(np.abs(inp["ru"]))
    out = noc_np.par_Newton(NX, d, 1.0, inp["ru"], inp["Q"], inp["R"], inp["M"])
    feas = bool(np.all(inp["cons"] <= 0))
    return out, hu, feas


def cpu_pass_serial(inp):
    """The same pass with the reference's SEQUENTIAL formulation restated in C, double precision
    (oracle/serial_ld.c: ref noc/costates.py:43-54 seq_costates, noc/seq_interior_point_newton.py:42-90 bwd/fwd
    pass with VxxN = Q[0] and rp = reg) — comparator "B2" of BASELINE.md section 3: O(N) work, one thread."""
    from oracle import serial_ld
    d = inp["d"]
    serial_ld.seq_costates(d.fx, d.cx, inp["lamT"], precision="f64")
    hu = np.max(np.abs(inp["ru"]))
    reg = 1.0 * np.linalg.norm(d.cu.reshape(-1))
    out = serial_ld.seq_newton(d.fx, d.fu, inp["ru"], inp["Q"], inp["R"], inp["M"], reg, precision="f64")
    feas = bool(np.all(inp["cons"] <= 0))
    return out, hu, feas


def time_cpu(inp, steps, warmup, fn=None):
    fn = fn or cpu_pass
    for _ in range(warmup):
        fn(inp)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(inp)
    return (time.perf_counter() - t0) / steps


def cpu_comparators(inp, steps, warmup):
    """Both CPU restatements of the pass on this host -> (best, dict of both).  The reference's own JAX paths
    cannot be run here (no jax / paroc); its par path is what B1 restates (tree scans, NumPy), its seq twin what
    B2 restates (serial recursion, C).  Both are single-threaded: the tree-scan port is bound by the Python-level
    recursion and batched small LAPACK solves, the serial recursion is sequential by construction."""
    dt_tree = time_cpu(inp, max(2, steps // 4), min(warmup, 2))
    dt_ser = time_cpu(inp, steps, warmup, cpu_pass_serial)
    alts = {"tree_scan_numpy_port_B1": {"ms_per_step": dt_tree * 1e3, "value": 1.0 / dt_tree, "cores": 1},
            "serial_riccati_C_port_B2": {"ms_per_step": dt_ser * 1e3, "value": 1.0 / dt_ser, "cores": 1}}
    best = "serial_riccati_C_port_B2" if dt_ser <= dt_tree else "tree_scan_numpy_port_B1"
    return best, alts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = 1   # both ports are single-threaded by construction (see cpu_comparators); host has os.cpu_count()
    inp = cpu_inputs(N_HEADLINE)
    best, alts = cpu_comparators(inp, args.steps, args.warmup)
    dt = alts[best]["ms_per_step"] * 1e-3
    val = 1.0 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cartpole nx=4 nu=1 N=10000 f64: one par IP-Newton hot-path pass (K1+K2+K3+K4)",
                   "note": "reference (JAX+paroc) not installable here: CPU ports of its two formulations of the step "
                           "on the host, the faster one reported; at --gpus N this arm is still ONE CPU process"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "which": best,
                         "alternatives": alts,
                         "sample": f"{args.steps} full passes at N=10000 after {args.warmup} warm-up; the faster of the "
                                   f"NumPy tree-scan port (B1) and the C serial-Riccati port (B2), both 1 thread; host "
                                   f"has {os.cpu_count()} cores"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm (GPU)
def run_ours(args):
    import torch
    import torch.distributed as dist
    from ipoc_b200 import _lib, workloads
    from ipoc_b200.runner import NewtonPass

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    numa = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    hbm_peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # ---- max over ranks, aggregate
    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])


    def make_pass(N, seed):
        w = workloads.newton_inputs("cartpole", N, dev, seed=seed, x0_noise=0.0 if seed == 1 else 0.01)
        return w, NewtonPass(w["fx"], w["fu"], w["cx"], w["cu"], w["lamT"], w["ru"], w["Q"], w["R"], w["M"], w["cons"])

    def time_steps(fn, steps, warmup, do_flush=True):
        for _ in range(warmup):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in evs:
            if do_flush:
                flush.zero_()
            a.record()
            fn()
            b.record()
        barrier()
        return [a.elapsed_time(b) for a, b in evs]   # ms

    # ---- headline workload: every rank its own OCP (rank 0 = the BASELINE config-2 inputs, seed 1)
    w, npass = make_pass(N_HEADLINE, seed=1 + rank)
    launches_per_pass = npass.launches_per_pass()
    npass.capture()
    clocks = Clocks(local) if rank == 0 else None
    ms = time_steps(npass.replay, args.steps, args.warmup, do_flush=True)
    ms_step = float(np.mean(ms))
    ms_plain = float(np.mean(time_steps(npass.run, args.steps, args.warmup, do_flush=True)))
    ms_warm = float(np.mean(time_steps(npass.replay, args.steps, args.warmup, do_flush=False)))
    # the same pass in the round-1 organisation (separate level kernels, K4 and A8 as their own launches: 11
    # launches), graph-replayed on the same box in the same run — the like-for-like "before" of the fused pass
    lib.ipoc_set_hier(0, 0, 0)
    _, npass_r1 = make_pass(N_HEADLINE, seed=1 + rank)
    npass_r1.fused = False
    launches_r1 = npass_r1.launches_per_pass()
    npass_r1.capture()
    ms_r1 = float(np.mean(time_steps(npass_r1.replay, args.steps, args.warmup, do_flush=True)))
    lib.ipoc_set_hier(1, 0, 0)
    del npass_r1

    # ---- per-launch profile (CUDA events after every launch, on the launching stream), L2 flushed
    def profile_pass(p, reps):
        acc = {}
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize(dev)
            for name, t in p.profile():
                acc.setdefault(name, []).append(t)
        return acc

    prof = profile_pass(npass, max(3, min(args.steps, 10)))
    kern = {k: (float(np.mean(v)) * (len(v) / max(3, min(args.steps, 10)))) for k, v in prof.items()}   # ms per pass
    phases = {}
    for k, t in kern.items():
        phases[PHASE_OF.get(k, "other")] = phases.get(PHASE_OF.get(k, "other"), 0.0) + t
    ab = alg_bytes(N_HEADLINE)
    dom = max(("K1", "K2", "K3"), key=lambda ph: phases.get(ph, 0.0))
    dom_ms = phases[dom]
    achieved = ab[dom] / (dom_ms * 1e-3) / 1e9

    # ---- end to end through the host-buffer C ABI call (pinned host memory)
    N = N_HEADLINE
    host = {k: w[k].detach().cpu().contiguous().pin_memory() for k in ("fx", "fu", "ru", "Q", "R", "M")}
    reg_h = torch.tensor([float(npass.cu_norm[0])], dtype=torch.float64).pin_memory()
    dx_h = torch.empty(N + 1, NX, dtype=torch.float64).pin_memory()
    du_h = torch.empty(N, NU, dtype=torch.float64).pin_memory()
    pred_h = torch.empty(1, dtype=torch.float64).pin_memory()
    feas_h = torch.empty(1, dtype=torch.int32).pin_memory()
    scratch_bytes = lib.ipoc_newton_step_host_scratch_bytes(N, NX, NU, 1)
    scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
    import ctypes
    hp = lambda t: ctypes.c_void_p(t.data_ptr())

    def e2e_step():
        _lib.check(lib.ipoc_newton_step_host_f64(N, NX, NU, 1, hp(host["fx"]), hp(host["fu"]), hp(host["ru"]),
                                                 hp(host["Q"]), hp(host["R"]), hp(host["M"]), hp(reg_h), hp(dx_h),
                                                 hp(du_h), hp(pred_h), hp(feas_h), hp(scratch), scratch_bytes,
                                                 _lib.stream_ptr()))

    ms_host_call = float(np.mean(time_steps(e2e_step, args.steps, args.warmup, do_flush=True)))
    h2d_call = sum(t.numel() * 8 for t in host.values()) + 8
    d2h_call = dx_h.numel() * 8 + du_h.numel() * 8 + 8 + 4
    torch.cuda.synchronize(dev)

    # ---- end to end of the WHOLE pass (same work as `value`): every input in one pinned host arena,
    #      one H2D copy + the pass + one D2H copy of the results, captured into one CUDA graph
    from ipoc_b200.runner import HostNewtonPass
    hpass = HostNewtonPass(w, dev)
    hpass.capture()
    ms_e2e = float(np.mean(time_steps(hpass.replay, args.steps, args.warmup, do_flush=True)))
    clk = clocks.stop() if clocks else None      # sampled across every timed region of the headline workload
    h2d_gbps = h2d_probe(dev)
    h2d, d2h = hpass.h2d_bytes, hpass.d2h_bytes
    for q in (npass, hpass.inner):       # rp / r_inc evolve from pass to pass: compare one pass from equal state
        q.rp.fill_(1.0)
        q.r_inc.fill_(2.0)
    npass.replay()
    hpass.replay()
    torch.cuda.synchronize(dev)
    e2e_check = max(float((hpass.results["dx"].to(dev) - npass.dx).abs().max()),
                    float((hpass.results["lam"].to(dev) - npass.lam).abs().max()),
                    float((hpass.results["du"].to(dev) - npass.du).abs().max()))
    from ipoc_b200 import noc
    dx_res = noc.newton_step(w["fx"], w["fu"], w["ru"], w["Q"], w["R"], w["M"], reg_h.to(dev))[0]
    dx_check = float((dx_h.to(dev) - dx_res).abs().max())   # host-buffer path == resident path on the same inputs

    ms_step_all, ms_e2e_all = allmax(ms_step), allmax(ms_e2e)
    value = world / (ms_step_all * 1e-3)
    e2e_value = world / (ms_e2e_all * 1e-3)

    # ---- time-sharded mode (BASELINE config 4): cartpole-shaped LQ data (nx=4, nu=1), horizon sweep
    #      N = 1e3 .. 1e6 (+1e7), ONE horizon cut into `world` contiguous segments.  Two measurements per N:
    #      (a) the Newton step K2+K3 (sharded.SegmentNewton: 3 graphs + 2 all-gathers), (b) the WHOLE hot-path pass
    #      K1+K4+K2+K3 (sharded.SegmentPass) with its 3 NCCL all-gathers captured inside ONE CUDA graph.
    #      Strong scaling, max over ranks; every rank also recomputes the step UNSHARDED (N <= 1e6) and reports the
    #      max relative error of its slice, so the NCCL path proves itself in the same run.
    time_sharded = []
    if not args.no_sweep:
        from ipoc_b200 import sharded, noc as _noc
        from ipoc_b200.runner import NewtonPass as _NP
        rel_ = lambda a_, b_: float((a_ - b_).abs().max() / b_.abs().max().clamp_min(1e-300))
        for Ns in (1_000, 10_000, 100_000, 1_000_000, 10_000_000):
            try:
                if Ns < world * 32:
                    continue
                bounds = sharded.segment_bounds(Ns, world)
                lo, hi = bounds[rank]
                check = Ns <= 1_000_000
                # segment r of the horizon is seeded by r: every rank can build any segment it needs
                gen = lambda r_: workloads.synthetic_lq(bounds[r_][1] - bounds[r_][0], NX, NU, dev, seed=100 + r_,
                                                        dt=min(0.5, 10.0 / Ns))
                g2 = torch.Generator(device=dev).manual_seed(7)
                aux = lambda n_, w_: torch.randn(n_, w_, dtype=torch.float64, device=dev, generator=g2)
                if check:
                    segs_ = [gen(r_) for r_ in range(world)]
                    full_ = tuple(torch.cat([sg[i] for sg in segs_]) for i in range(6))
                    cx_f, cu_f, cons_f = aux(Ns, NX), aux(Ns, NU), -aux(Ns, NC).abs()
                    mine = tuple(t[lo:hi].contiguous() for t in full_)
                    cx_, cu_, cons_ = cx_f[lo:hi].contiguous(), cu_f[lo:hi].contiguous(), cons_f[lo:hi].contiguous()
                    ST = full_[3][0].contiguous()
                else:
                    mine = gen(rank)
                    cx_, cu_, cons_ = aux(hi - lo, NX), aux(hi - lo, NU), -aux(hi - lo, NC).abs()
                    ST = workloads.synthetic_lq(4, NX, NU, dev, seed=100, dt=min(0.5, 10.0 / Ns))[3][0].contiguous()
                fx_, fu_, ru_, Q_, R_, M_ = mine
                lamT_ = torch.ones(NX, dtype=torch.float64, device=dev)
                regs = torch.tensor([0.25], dtype=torch.float64, device=dev)
                rec = {"N": Ns, "n_gpus": world}
                gather_into = sharded.dist_all_gather_into() if world > 1 else (lambda out, t: out.copy_(t.reshape(out.shape)))

                def timed(fn, reps=10):
                    for _ in range(3):
                        fn()
                    barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(reps):
                        fn()
                    e1.record()
                    barrier()
                    return allmax(e0.elapsed_time(e1) / reps)

                # (a) Newton step K2 + K3
                if world > 1:
                    seg = sharded.SegmentNewton(fx_, fu_, ru_, Q_, R_, M_, rank, world)
                    seg.capture_one(regs, ST, gather_into)   # 3 local phases + 2 NCCL all-gathers in ONE CUDA graph
                    fn_a = seg.step_one
                else:
                    fn_a = lambda: _noc.newton_step(fx_, fu_, ru_, Q_, R_, M_, regs)
                rec["ms_per_newton_step_K2K3"] = timed(fn_a)
                # (b) whole pass, one graph including the collectives
                sp = sharded.SegmentPass(fx_, fu_, cx_, cu_, lamT_, ru_, Q_, R_, M_, rank, world, cons_, rp=0.25)
                sp.capture(gather_into, ST)
                rec["ms_per_pass_K1K2K3K4_one_graph"] = timed(sp.replay)
                ab_ = alg_bytes(Ns)
                rec["hbm_frac_of_aggregate_peak_K2K3"] = ((ab_["K2"] + ab_["K3"]) / (rec["ms_per_newton_step_K2K3"] * 1e-3)
                                                          / 1e9 / (hbm_peak * world))
                rec["hbm_frac_of_aggregate_peak_pass"] = (ab_["total"] / (rec["ms_per_pass_K1K2K3K4_one_graph"] * 1e-3)
                                                          / 1e9 / (hbm_peak * world))
                rec["collectives_per_pass"] = 0 if world == 1 else 3
                rec["launch"] = ("K2+K3: ONE CUDA graph incl. 2 NCCL all-gathers; pass: ONE CUDA graph incl. 3 NCCL all-gathers"
                                 if world > 1 else "K2+K3: eager API call; pass: one CUDA graph") + ", max over ranks"
                if check:      # the sharded results of this rank's slice against the single-device scans
                    ref = _NP(full_[0], full_[1], cx_f, cu_f, lamT_, full_[2], full_[3], full_[4], full_[5], cons_f, rp=0.25)
                    ref.run()
                    sp.replay()
                    torch.cuda.synchronize(dev)
                    err = max(rel_(sp.lam, ref.lam[0, lo:hi + 1]), rel_(sp.new.dx, ref.dx[0, lo:hi + 1]),
                              rel_(sp.new.du, ref.du[0, lo:hi]), rel_(sp.new.Kx, ref.Kx[0, lo:hi]),
                              abs(float(sp.scalars()[0]) - float(ref.pred)) / abs(float(ref.pred)))
                    rec["max_rel_err_vs_single_device"] = allmax(err)
                    del ref, full_, segs_, cx_f, cu_f, cons_f
                time_sharded.append(rec)
                del fx_, fu_, ru_, Q_, R_, M_, mine, sp
                torch.cuda.empty_cache()
            except Exception as e:
                time_sharded.append({"N": Ns, "error": repr(e)[:300]})

    # ---- horizon sweep (single GPU): Newton-step time and HBM fraction at N = 1e5, 1e6
    sweep = []
    if rank == 0 and not args.no_sweep:
        for Ns in (100_000, 1_000_000):
            try:
                t0 = time.time()
                ws_, ps_ = make_pass(Ns, seed=1)
                ps_.capture()
                m = float(np.mean(_time_local(torch, ps_.replay, flush, 5, 3, dev)))
                pr = {}
                for _ in range(3):
                    flush.zero_()
                    torch.cuda.synchronize(dev)
                    for name, t in ps_.profile():
                        pr[name] = pr.get(name, 0.0) + t / 3
                ph = {}
                for k, t in pr.items():
                    ph[PHASE_OF.get(k, "other")] = ph.get(PHASE_OF.get(k, "other"), 0.0) + t
                abs_ = alg_bytes(Ns)
                sweep.append({"N": Ns, "ms_per_step": m,
                              "hbm_frac_step": abs_["total"] / (m * 1e-3) / 1e9 / hbm_peak,
                              "phase_ms": {k: round(v, 4) for k, v in ph.items()},
                              "hbm_frac_K2": abs_["K2"] / (ph.get("K2", 1e9) * 1e-3) / 1e9 / hbm_peak,
                              "hbm_frac_K1": abs_["K1"] / (ph.get("K1", 1e9) * 1e-3) / 1e9 / hbm_peak,
                              "hbm_frac_K3": abs_["K3"] / (ph.get("K3", 1e9) * 1e-3) / 1e9 / hbm_peak,
                              "setup_s": round(time.time() - t0, 1)})
                del ws_, ps_
                torch.cuda.empty_cache()
            except Exception as e:   # the sweep must never take the headline line down
                sweep.append({"N": Ns, "error": repr(e)[:200]})

    # ---- full solves through the reference-facing API (BASELINE metric "IP-Newton solve ms vs horizon N"):
    #      reference protocol (ref examples/cartpole_runtime.py:119-146): 1 warm-up (graph capture ~ jit), then
    #      timed calls; rollout + derivatives (host framework) + all five barrier stages included.
    solves, batched_solves = [], None
    plant_fast_path = None
    if rank == 0 and world == 1 and not args.no_solve:
        from ipoc_b200 import noc as _noc2, problems as _pb, batched as _bt
        for Ns in (1000, 10_000, 100_000, 1_000_000):
            try:
                ocp_ = _pb.make_cartpole(1.0 / Ns)
                x0_ = _pb.cartpole_x0(device=dev)
                u0_ = torch.as_tensor(0.1 * np.random.default_rng(1).standard_normal((Ns, 1)), device=dev)
                _noc2.par_interior_point_optimal_control(ocp_, u0_, x0_)
                torch.cuda.synchronize(dev)
                ts_ = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    tr_ = []
                    u_, it_ = _noc2.par_interior_point_optimal_control(ocp_, u0_, x0_)
                    torch.cuda.synchronize(dev)
                    ts_.append(time.perf_counter() - t0)
                from ipoc_b200 import plants as _pl
                plant_fast_path = bool(_pl.plant_of(ocp_) is not None)
                solves.append({"problem": "cartpole", "N": Ns, "solve_ms_mean": float(np.mean(ts_)) * 1e3,
                               "solve_ms_median": float(np.median(ts_)) * 1e3, "newton_iterations": int(it_),
                               "max_abs_u": float(u_.abs().max())})
            except Exception as e:
                solves.append({"N": Ns, "error": repr(e)[:200]})
        batched_solves = []
        for prob_, Bb in (("pendulum", 8192), ("cartpole", 8192)):   # BASELINE config 5: 8192 OCPs per GPU, N = 1000
            try:
                Nb = 1000
                ocp_ = _pb.make_pendulum(1.0 / Nb) if prob_ == "pendulum" else _pb.make_cartpole(1.0 / Nb)
                x0b = (_pb.pendulum_x0 if prob_ == "pendulum" else _pb.cartpole_x0)(device=dev)
                rng_ = np.random.default_rng(1)
                x0s_ = x0b[None] + torch.as_tensor(0.1 * rng_.standard_normal((Bb, x0b.numel())), device=dev)
                u0s_ = torch.as_tensor(0.1 * rng_.standard_normal((Bb, Nb, 1)), device=dev)
                # warm-up solve of the same batch: captures the ladder of device-resident loops (~ jit in the
                # reference's timing protocol, ref examples/cartpole_runtime.py:119-146), then one timed solve
                _bt.par_interior_point_optimal_control_batched(ocp_, u0s_, x0s_)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                ub_, itb_ = _bt.par_interior_point_optimal_control_batched(ocp_, u0s_, x0s_)
                torch.cuda.synchronize(dev)
                dtb = time.perf_counter() - t0
                # per-OCP check on a 64-problem subsample: iterate and Newton iteration count of solving it alone
                sub_ = list(range(0, Bb, Bb // 64))[:64]
                it1_, err_ = [], 0.0
                for b_ in sub_:
                    u1_, i1_ = _noc2.par_interior_point_optimal_control(ocp_, u0s_[b_], x0s_[b_])
                    it1_.append(int(i1_))
                    err_ = max(err_, float((u1_ - ub_[b_]).abs().max() / u1_.abs().max()))
                batched_solves.append({"problem": prob_, "N": Nb, "batch": Bb, "solves_per_s": Bb / dtb,
                                       "seconds": dtb, "iterations_mean": float(itb_.double().mean()),
                                       "iterations_max": int(itb_.max()),
                                       "subsample64_iteration_histogram_equal": bool(it1_ == [int(itb_[b_]) for b_ in sub_]),
                                       "subsample64_max_rel_u_err_vs_unbatched": err_})
            except Exception as e:
                batched_solves.append({"problem": prob_, "error": repr(e)[:200]})

    # ---- BASELINE config 5 on N > 1 GPUs: the batch of independent OCPs is split over the ranks
    #      (sharded.shard_batch, no data-path collective); solves/s = all problems / max-over-ranks time.
    if world > 1 and not args.no_solve:
        from ipoc_b200 import problems as _pbw, batched as _btw, sharded as _shw
        batched_solves = []
        for prob_, per_rank in (("pendulum", 8192), ("cartpole", 8192)):   # config 5 as stated: 65 536 OCPs on 8 GPUs
            try:
                Nb, Bt = 1000, per_rank * world
                ocp_ = _pbw.make_pendulum(1.0 / Nb) if prob_ == "pendulum" else _pbw.make_cartpole(1.0 / Nb)
                x0b = (_pbw.pendulum_x0 if prob_ == "pendulum" else _pbw.cartpole_x0)(device=dev)
                rng_ = np.random.default_rng(1)                      # same global batch on every rank
                x0_all = 0.1 * rng_.standard_normal((Bt, x0b.numel()))
                u0_all = 0.1 * rng_.standard_normal((Bt, Nb, 1))
                lo, hi = _shw.shard_batch(Bt, rank, world)
                x0s_ = x0b[None] + torch.as_tensor(x0_all[lo:hi], device=dev)
                u0s_ = torch.as_tensor(u0_all[lo:hi], device=dev)
                _btw.par_interior_point_optimal_control_batched(ocp_, u0s_, x0s_)   # warm-up: captures the loop ladder
                barrier()
                t0 = time.perf_counter()
                ub_, itb_ = _btw.par_interior_point_optimal_control_batched(ocp_, u0s_, x0s_)
                torch.cuda.synchronize(dev)
                dtb = allmax(time.perf_counter() - t0)
                its_sum = torch.tensor([float(itb_.double().sum())], dtype=torch.float64, device=dev)
                dist.all_reduce(its_sum)
                if rank == 0:
                    batched_solves.append({"problem": prob_, "N": Nb, "batch": Bt, "batch_per_gpu": per_rank,
                                           "n_gpus": world, "solves_per_s": Bt / dtb, "seconds": dtb,
                                           "iterations_mean": float(its_sum[0]) / Bt, "scaling": "weak"})
            except Exception as e:
                batched_solves.append({"problem": prob_, "error": repr(e)[:200]})

    # ---- BASELINE config 3 as written (ref examples/linear_mpc_parallel.py:67-81): 5000 receding-horizon steps,
    #      each = par_bwd_pass + par_fwd_pass on the T = 5 double-integrator LQT; the whole loop is timed, like the
    #      reference does (after one untimed run).  CPU side: the oracle's par passes on a 300-step sample.
    mpc = None
    if rank == 0 and world == 1 and not args.no_solve:
        try:
            from ipoc_b200 import problems as _pb3
            from ipoc_b200.paroc import LQT as _LQT
            from ipoc_b200.mpc import MpcLoop
            fields_, x0m = _pb3.make_mpc_lqt_terms(T=5, device=dev)
            mpc = {"steps": 5000, "horizon_T": 5, "unroll_per_graph": 100}
            for label, serial in (("par", False), ("seq", True)):
                loop = MpcLoop(_LQT(*fields_), unroll=100, serial=serial)
                loop.run(x0m, 5000)
                torch.cuda.synchronize(dev)
                tm = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    xs_m, us_m = loop.run(x0m, 5000)
                    torch.cuda.synchronize(dev)
                    tm.append(time.perf_counter() - t0)
                mpc[f"{label}_loop_seconds"] = float(np.median(tm))
                mpc[f"{label}_final_state"] = [float(v) for v in xs_m[-1]]
            if not args.no_cpu:
                from oracle import paroc_np as _pn
                lq_ = _pn.LQT(*(f.cpu().numpy() for f in fields_))
                xk = x0m.cpu().numpy()
                t0 = time.perf_counter()
                for _ in range(300):
                    Kx_, d_ = _pn.par_bwd_pass(lq_)[:2]
                    _, xo_ = _pn.par_fwd_pass(lq_, xk, Kx_, d_)
                    xk = xo_[1]
                mpc["cpu_port_seconds_per_5000_steps"] = (time.perf_counter() - t0) * 5000 / 300
                mpc["cpu_port_sample"] = "300 MPC steps of the NumPy oracle's par passes, scaled to 5000"
        except Exception as e:
            mpc = {"error": repr(e)[:200]}

    # ---- B3 proxy: the same Newton step as a log-depth tree scan of batched torch ops on this
    #      GPU — structurally what XLA:GPU emits for lax.associative_scan; NOT the reference (JAX is absent).
    proxy = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            sys.path.insert(0, os.path.join(ROOT, "baseline"))
            import torch_treescan as _tt
            regp = float(npass.cu_norm[0])
            fnp = lambda: _tt.par_newton(w["fx"], w["fu"], w["ru"], w["Q"], w["R"], w["M"], regp)
            dxp = fnp()[0]
            dxr = _noc2_step(w, regp, dev)
            ms_p = float(np.mean(_time_local(torch, fnp, flush, 5, 3, dev)))
            ms_ours = float(np.mean(_time_local(torch, lambda: _noc2_step(w, regp, dev), flush, 20, 3, dev)))
            proxy = {"kind": "torch batched odd/even tree scan + torch.linalg.solve on B200, eager (structural proxy "
                             "for XLA:GPU; NOT the reference)", "ms_per_newton_step_K2K3": ms_p,
                     "ours_ms_per_newton_step_K2K3_eager_api": ms_ours,
                     "max_rel_dx_diff_vs_ours": float((dxp - dxr).abs().max() / dxr.abs().max())}
        except Exception as e:
            proxy = {"error": repr(e)[:200]}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload with the oracle port
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        inp = cpu_inputs(N_HEADLINE)
        n_cpu = 40
        best, alts = cpu_comparators(inp, n_cpu, 2)
        dt = alts[best]["ms_per_step"] * 1e-3
        cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": "port", "which": best, "alternatives": alts,
               "sample": f"full passes at N=10000: the faster of the NumPy tree-scan port of the par path (B1, "
                         f"{max(2, n_cpu // 4)} passes) and the C serial-Riccati port of the seq twin (B2, {n_cpu} passes); "
                         f"both single-threaded; host has {os.cpu_count()} cores; ms_per_step={dt * 1e3:.2f}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step_all, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cartpole nx=4 nu=1 nc=2 N=10000 f64 (BASELINE configs[1]): one par IP-Newton "
                                   "hot-path pass K1+K2+K3+K4 per step; N>1: one independent OCP per GPU "
                                   "(batch-sharded, no collective)",
                       "l2": "L2 flushed (256 MiB memset) before every timed step",
                       "launch": "CUDA-graph replay of one pass", "seed": 1},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e_all,
                    "api": "runner.HostNewtonPass: one pinned host arena -> H2D, the same pass as `value` "
                           "(K1+K4+K2+K3+K4+A8 through the C ABI), D2H of lam/dx/du/scalars; one CUDA graph",
                    "check_max_abs_diff_vs_resident": e2e_check,
                    "host_placement": numa, "pinned_h2d_GBps_probe": h2d_gbps,
                    "newton_step_host_call": {
                        "api": "ipoc_newton_step_host_f64 (K2+K3 only, six pinned host buffers, eager)",
                        "ms_per_step": ms_host_call, "h2d_bytes_per_step": h2d_call,
                        "d2h_bytes_per_step": d2h_call, "check_max_abs_dx_diff_vs_resident": dx_check}},
            "gpu_launches": launches_per_pass * args.steps,
            "launches_per_step": launches_per_pass,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": NCU_TRAFFIC_N1E4.get(dom),
                         "traffic_source": "ncu dram__bytes_{read,write}.sum, profiles/r02_launches_bench_ncu.csv "
                                           "(K2 = k_ric_leaf_up 3.43 MB + k_top 0.10 MB + k_ric_leaf_down 4.33 MB: the "
                                           "inputs are read by the up- and by the down-sweep)",
                         "kernel": f"{dom} phase (all its launches)", "peak_source": peak_src,
                         "algorithmic_bytes": ab[dom], "duration_ms": dom_ms,
                         "note": "N=1e4 moves 8 MB: latency regime, see sweep for N>=1e5"},
            "kernels_ms_per_step": {k: round(v, 5) for k, v in sorted(kern.items(), key=lambda kv: -kv[1])},
            "phases_ms_per_step": {k: round(v, 5) for k, v in phases.items()},
            "ms_per_step_plain_launch": ms_plain, "ms_per_step_l2_warm": ms_warm,
            "ms_per_step_round1_organisation": {"ms_per_step": ms_r1, "launches_per_step": launches_r1,
                                                "what": "same pass, same box, same run: separate level kernels, "
                                                        "K4 / A8 as their own launches (graph replay, L2 flushed)"},
            "sweep": sweep,
            "time_sharded": time_sharded,
            "xla_proxy": proxy,
            "solves": solves,
            "solves_note": "built-in plant kernels (fused derivatives/cost/rollout) " + ("ON" if plant_fast_path else "off"),
            "batched_solves": batched_solves,
            "mpc": mpc,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs holding captured NCCL kernels are still alive (time-sharded records): tearing the process
        # group down under them hangs, so every rank synchronises and leaves without the collective teardown
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def _noc2_step(w, reg, dev):
    import torch
    from ipoc_b200 import noc
    return noc.newton_step(w["fx"], w["fu"], w["ru"], w["Q"], w["R"], w["M"],
                           torch.tensor([reg], dtype=torch.float64, device=dev))[0]


def _time_local(torch, fn, flush, steps, warmup, dev):
    for _ in range(warmup):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize(dev)
    for a, b in evs:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize(dev)
    return [a.elapsed_time(b) for a, b in evs]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-solve", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

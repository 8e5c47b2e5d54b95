#!/usr/bin/env python
"""Multi-GPU check of the time-sharded Newton step with real NCCL (run on the GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_time_sharded.py [N] [nx]

Every rank builds the same synthetic LQ data, takes its contiguous time segment, runs
reduce -> all-gather -> seeded scan with torch.distributed (NCCL), and compares its slice with the
single-device scan of the whole horizon computed locally.  Prints one OK line per rank + timings."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist
from helpers import random_lq
from ipoc_b200 import noc, sharded


def stage(msg):
    print(f"[rank {os.environ.get('RANK')}] {msg}", file=sys.stderr, flush=True)


def main():
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100003
    nx = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(5)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, 1, dt=min(0.5, 10.0 / N))
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    full = [T(a) for a in (fx, fu, ru, Q, R, M)]
    reg = torch.tensor([0.25], dtype=torch.float64, device=dev)
    stage("init done")
    dx1, du1, Kx1, d1, pred1, feas1 = noc.newton_step(*full, reg)
    lo, hi = sharded.segment_bounds(N, world)[rank]
    seg = sharded.SegmentNewton(*(t[lo:hi] for t in full), rank, world)
    gather = sharded.dist_all_gather()
    ST = full[3][0].contiguous()
    dx, du, pred, feas = sharded.newton_step_time_sharded(seg, reg, ST, gather)
    torch.cuda.synchronize()
    stage("eager sharded step done")
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
    errs = (rel(dx, dx1[lo:hi + 1]), rel(du, du1[lo:hi]), rel(seg.Kx, Kx1[lo:hi]), rel(seg.d, d1[lo:hi]),
            abs(float(pred) - float(pred1)) / abs(float(pred1)))
    ok = max(errs) < 1e-10 and feas == bool(feas1[0])
    # timing (max over ranks)
    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])
    ms_sharded = timed(lambda: sharded.newton_step_time_sharded(seg, reg, ST, gather))
    stage("eager timing done")
    seg.capture(reg, ST)
    stage("3-graph capture done")
    gather_into = sharded.dist_all_gather_into()
    dxg, dug, predg, _ = seg.step_graphed(gather_into)
    torch.cuda.synchronize()
    errg = max(rel(dxg, dx1[lo:hi + 1]), rel(dug, du1[lo:hi]), abs(float(predg) - float(pred1)) / abs(float(pred1)))
    ok = ok and errg < 1e-10
    ms_graphed = timed(lambda: seg.step_graphed(gather_into))
    stage("3-graph timing done")
    seg.capture_one(reg, ST, gather_into)
    stage("one-graph capture done")
    dxo, duo, predo, _ = seg.step_one()
    torch.cuda.synchronize()
    ok = ok and max(rel(dxo, dx1[lo:hi + 1]), rel(duo, du1[lo:hi]), abs(float(predo) - float(pred1)) / abs(float(pred1))) < 1e-10
    ms_one = timed(seg.step_one)
    stage("one-graph timing done")
    ms_single = timed(lambda: noc.newton_step(*full, reg))
    # ---- the WHOLE pass (K1 + K4 + K2 + K3) time-sharded, three all-gathers captured inside ONE CUDA graph
    from ipoc_b200.runner import NewtonPass
    cx, cu = T(rng.standard_normal((N, nx))), T(rng.standard_normal((N, 1)))
    lamT = T(rng.standard_normal(nx))
    cons = T(-rng.random((N, 2)))
    ref = NewtonPass(full[0], full[1], cx, cu, lamT, full[2], full[3], full[4], full[5], cons, rp=0.8)
    ref.run()
    sp = sharded.SegmentPass(full[0][lo:hi], full[1][lo:hi], cx[lo:hi], cu[lo:hi], lamT, full[2][lo:hi],
                             full[3][lo:hi], full[4][lo:hi], full[5][lo:hi], rank, world, cons[lo:hi], rp=0.8)
    stage("pass objects built")
    sp.capture(gather_into, ST)
    stage("pass captured")
    sp.replay()
    torch.cuda.synchronize()
    stage("pass replayed")
    pred_p, bf_p, tf_p = sp.scalars()
    errp = max(rel(sp.lam, ref.lam[0, lo:hi + 1]), rel(sp.new.dx, ref.dx[0, lo:hi + 1]), rel(sp.new.du, ref.du[0, lo:hi]),
               abs(float(pred_p) - float(ref.pred)) / abs(float(ref.pred)),
               abs(float(sp.cu_norm) - float(ref.cu_norm)) / float(ref.cu_norm))
    ok = ok and errp < 1e-10 and float(sp.hu) == float(ref.hu) and bf_p == bool(ref.bwd_feas[0]) and tf_p
    ms_pass = timed(sp.replay)
    ref.capture()
    ms_pass_single = timed(ref.replay)
    print(f"[rank {rank}/{world}] N={N} nx={nx} segment=[{lo},{hi}) max rel err {max(max(errs), errg, errp):.2e} "
          f"{'OK' if ok else 'FAIL'}  time-sharded K2+K3 {ms_sharded:.3f} ms (3 graphs {ms_graphed:.3f} ms, one graph incl. NCCL {ms_one:.3f} ms) "
          f"vs single-GPU {ms_single:.3f} ms;  whole pass, one graph incl. 3 NCCL all-gathers {ms_pass:.3f} ms "
          f"vs single-GPU graph {ms_pass_single:.3f} ms", flush=True)
    # CUDA graphs that captured NCCL kernels must be gone before the communicator is torn down (destroying the
    # process group first hangs); leave without the collective teardown altogether
    del sp, seg, ref
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()

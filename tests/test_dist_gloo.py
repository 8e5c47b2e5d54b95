"""CPU, world_size 2, gloo: the multi-GPU plumbing — torch.distributed all-gather wrapper, segment
bounds, and the time-sharded exchange logic (NumPy model of the per-rank kernels) against the
unsharded oracle.  The same algorithm runs on real GPUs with NCCL (tests/dist_time_sharded.py,
launched by torchrun under `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, nx, nu, out):
    for p in (ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import random_lq
    from ipoc_b200 import sharded
    from oracle import sharded_np
    rng = np.random.default_rng(11)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    lo, hi = sharded.segment_bounds(N, world)[rank]
    gather = sharded.dist_all_gather()
    reg = 0.3
    lqt, agg = sharded_np.bwd_reduce(fx[lo:hi], fu[lo:hi], ru[lo:hi], Q[lo:hi], R[lo:hi], M[lo:hi], reg)
    flat = torch.as_tensor(np.concatenate([a.ravel() for a in agg]))
    allc = gather(flat).numpy()                                        # exchange 1 (rank-major)
    sizes = [a.size for a in agg]
    shapes = [a.shape for a in agg]
    carries = []
    for r in range(world):
        parts, o = [], 0
        for sz, sh in zip(sizes, shapes):
            parts.append(allc[r, o:o + sz].reshape(sh))
            o += sz
        carries.append(tuple(parts))
    Kx, d, pred, feas, (F, c) = sharded_np.bwd_apply(lqt, rank, world, carries, Q[0])
    fl = torch.as_tensor(np.concatenate([F.ravel(), c, [pred, float(feas)]]))
    allf = gather(fl).numpy()                                          # exchange 2
    fwd = [(allf[r, :nx * nx].reshape(nx, nx), allf[r, nx * nx:nx * nx + nx]) for r in range(world)]
    x, u = sharded_np.fwd_apply(lqt, rank, fwd, Kx, d)
    res = dict(dx=x, du=u, Kx=Kx, d=d, pred=float(allf[:, -2].sum()), feas=bool(np.all(allf[:, -1] != 0)), lo=lo,
               hi=hi)
    torch.save(res, os.path.join(out, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nx,nu,N", [(4, 1, 101), (2, 1, 64)])
def test_time_sharded_exchange_logic_world2(tmp_path, nx, nu, N):
    from helpers import random_lq, relerr
    from oracle import noc_np, paroc_np
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), N, nx, nu, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(11)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    lqt = noc_np.noc_to_lqt(ru, Q, R + 0.3 * np.eye(nu)[None], M, fx, fu)
    Kx, d, S, v, pred, feas = paroc_np.par_bwd_pass(lqt)
    du, dx = paroc_np.par_fwd_pass(lqt, np.zeros(nx), Kx, d)
    parts = [torch.load(os.path.join(tmp_path, f"rank{r}.pt"), weights_only=False) for r in range(world)]
    Kx_s = np.concatenate([p["Kx"] for p in parts])
    du_s = np.concatenate([p["du"] for p in parts])
    dx_s = np.concatenate([p["dx"][:-1] for p in parts[:-1]] + [parts[-1]["dx"]])
    assert parts[0]["lo"] == 0 and parts[-1]["hi"] == N and parts[0]["hi"] == parts[1]["lo"]
    assert relerr(Kx_s, Kx) < 1e-10 and relerr(du_s, du) < 1e-10 and relerr(dx_s, dx) < 1e-10
    for p in parts:
        assert abs(p["pred"] - pred) <= 1e-10 * abs(pred) and p["feas"] == feas


def test_batch_sharding_covers_all_problems():
    from ipoc_b200 import sharded
    for batch, world in [(65536, 8), (10, 4), (7, 8)]:
        spans = [sharded.shard_batch(batch, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))

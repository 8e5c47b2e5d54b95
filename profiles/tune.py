"""Sweep the scan-plan knobs (leaf chunk, mid fan-in, top width) and print the time of one hot-path
pass (CUDA-graph replay, L2 flushed).  usage: python profiles/tune.py N [nx] [batch]"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from helpers import random_lq
from ipoc_b200 import _lib
from ipoc_b200.runner import NewtonPass

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 4
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rng = np.random.default_rng(0)
fx, fu, ru, Q, R, M = random_lq(rng, N, nx, 1, batch=batch or None, dt=1.0 / N)
shape = fx.shape[:-2]
T = lambda a: torch.as_tensor(a, device="cuda")
args = (T(fx), T(fu), T(rng.standard_normal(shape + (nx,))), T(rng.standard_normal(shape + (1,))),
        T(rng.standard_normal((max(batch, 1), nx))), T(ru), T(Q), T(R), T(M), T(-np.ones(shape + (2,))))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
lib = _lib.lib()


def timeit(p, reps=10):
    p.capture()
    for _ in range(3):
        p.replay()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in ev:
        flush.zero_()
        a.record(); p.replay(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


grid = [(0, 0, 0)] + list(itertools.product([2, 3, 4, 6, 8, 12, 16, 32], [2, 4, 8, 16], [64, 128, 256, 512, 1024]))
if len(sys.argv) > 4:
    grid = [tuple(int(v) for v in g.split(",")) for g in sys.argv[4:]]
res = []
for t0, mid, top in grid:
    lib.ipoc_set_tuning(t0, mid, top)
    try:
        p = NewtonPass(*args)
        ms = timeit(p)
        res.append((ms, t0, mid, top, p.launches_per_pass()))
    except Exception as e:
        print("fail", t0, mid, top, repr(e)[:100])
    del p
res.sort()
print(f"N={N} nx={nx} batch={batch}")
for r in res[:12]:
    print("  ms=%.4f leaf=%d mid=%d top=%d launches=%d" % r)
d = [r for r in res if r[1:4] == (0, 0, 0)]
if d:
    print("  default: ms=%.4f launches=%d" % (d[0][0], d[0][4]))

"""Graph-replay time of one fused hot-path pass against the leaf chunk length T0 (steps folded per lane before the
in-warp scan) and the Riccati level organisation.  usage: python profiles/prof_chunk_sweep.py N [N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from helpers import random_lq
from ipoc_b200 import _lib
from ipoc_b200.runner import NewtonPass

T = lambda a: torch.as_tensor(a, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
CHUNKS = [int(c) for c in os.environ.get("CHUNKS", "0,1,2,3,4,6,8,16").split(",")]
for N in [int(float(a)) for a in sys.argv[1:]] or [10000]:
    rng = np.random.default_rng(0)
    fx, fu, ru, Q, R, M = random_lq(rng, N, 4, 1, dt=1.0 / N)
    shape = fx.shape[:-2]
    args = (T(fx), T(fu), T(rng.standard_normal(shape + (4,))), T(rng.standard_normal(shape + (1,))),
            T(rng.standard_normal((1, 4))), T(ru), T(Q), T(R), T(M), T(-np.ones(shape + (2,))))
    for hier in (1, 3, 4):
        for chunk in CHUNKS:
            if chunk and N // chunk > 32 * 32 * 32:
                continue
            _lib.lib().ipoc_set_hier(hier, 0, 0)
            _lib.lib().ipoc_set_tuning(chunk, 0, 0)
            p = NewtonPass(*args)
            for _ in range(3):
                p.run()
            torch.cuda.synchronize()
            p.capture()
            ts = []
            for _ in range(20):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); p.replay(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            print(f"N={N} hier={hier} leaf_chunk={chunk}: {np.median(ts) * 1e3:.1f} us (min {np.min(ts) * 1e3:.1f})", flush=True)
_lib.lib().ipoc_set_hier(1, 0, 0)
_lib.lib().ipoc_set_tuning(0, 0, 0)

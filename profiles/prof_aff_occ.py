"""K1 (costate scan) time vs the leaf warps per SM it is planned for.  usage: python profiles/prof_aff_occ.py N [N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ip-parallel-optimal-control_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from ipoc_b200 import _lib, noc

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for N in [int(float(a)) for a in sys.argv[1:]] or [1000000]:
    g = torch.Generator(device=dev).manual_seed(0)
    F = torch.eye(4, dtype=torch.float64, device=dev) + (1.0 / N) * torch.randn(N, 4, 4, dtype=torch.float64, device=dev, generator=g)
    c = torch.randn(N, 4, dtype=torch.float64, device=dev, generator=g)
    seed = torch.ones(4, dtype=torch.float64, device=dev)
    ref = None
    for occ in (0, 6, 8, 10, 12, 16):
        _lib.lib().ipoc_set_affine_occupancy(occ)
        out = noc.affine_scan(F, c, seed, reverse=True, transpose=True)
        ts = []
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = noc.affine_scan(F, c, seed, reverse=True, transpose=True)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        if ref is None:
            ref = out.clone()
        err = float((out - ref).abs().max() / ref.abs().max())
        t = float(np.median(ts)) * 1e3
        print(f"N={N} warps/SM={occ}: K1 {t:.1f} us = {192.0 * N / t / 1e3 / 6537.3:.3f} of HBM peak (algorithmic), "
              f"rel diff vs default {err:.1e}", flush=True)
_lib.lib().ipoc_set_affine_occupancy(0)

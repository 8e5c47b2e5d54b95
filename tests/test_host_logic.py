"""CPU: the C-ABI library loads and exports every symbol include/ipoc.h declares; pure host logic."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from ipoc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ipoc.h")).read()
    declared = set(re.findall(r"\b(ipoc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(L, sym), f"{sym} declared in include/ipoc.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_pure_host_entry_points_without_gpu():
    from ipoc_b200 import _lib
    L = _lib.lib()
    assert L.ipoc_version() >= 100
    assert L.ipoc_supported(4, 1) == 1 and L.ipoc_supported(7, 1) == 1 and L.ipoc_supported(9, 1) == 0
    assert L.ipoc_supported(6, 3) == 1 and L.ipoc_supported(5, 3) == 1 and L.ipoc_supported(5, 5) == 0 and L.ipoc_supported(3, 4) == 0
    assert L.ipoc_carry_doubles(_lib.CARRY_RICCATI, 4) == 44 and L.ipoc_carry_doubles(_lib.CARRY_AFFINE, 4) == 20
    assert L.ipoc_carry_doubles(_lib.CARRY_RICCATI, 2) == 14
    small = L.ipoc_workspace_bytes(_lib.WS_NEWTON_STEP, 1000, 4, 1, 1)
    big = L.ipoc_workspace_bytes(_lib.WS_NEWTON_STEP, 1000000, 4, 1, 1)
    assert 0 < small < big
    assert L.ipoc_workspace_bytes(_lib.WS_NEWTON_STEP, 1000, 9, 1, 1) == 0   # unsupported nx
    assert L.ipoc_strerror(-2).decode().startswith("workspace")


def test_no_cpu_fallback_in_product_path():
    import torch
    from ipoc_b200 import noc, _lib
    with pytest.raises(_lib.IpocError):
        noc.par_interior_point_optimal_control(None, torch.zeros(4, 1), torch.zeros(2))
    # the product package must not import the oracle
    src_dir = os.path.join(ROOT, "ip-parallel-optimal-control_b200", "ipoc_b200")
    for fn in os.listdir(src_dir):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(src_dir, fn)).read().replace("no oracle", ""), fn


def test_segment_bounds_cover_horizon():
    from ipoc_b200 import sharded
    for N, P in [(10, 3), (1003, 8), (8, 8), (1000000, 8)]:
        b = sharded.segment_bounds(N, P)
        assert b[0][0] == 0 and b[-1][1] == N and all(b[i][1] == b[i + 1][0] for i in range(P - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` needs no GPU: it times the oracle port on the host and prints ONE JSON line
    with the keys the driver reads (metric/unit/config equal to the CUDA arm's, impl, cpu_baseline, e2e)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ip_newton_step_throughput_cartpole_N1e4"
    assert d["unit"] == "newton_steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(d["e2e"]["value"] - d["value"]) < 1e-9 * d["value"]

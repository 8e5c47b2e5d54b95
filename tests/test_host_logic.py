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


def test_workspace_sizes_include_the_control_block():
    """Every scan workspace starts with the IPOC_WS_CONTROL_BYTES control block (arrival counters of the in-kernel
    levels); the fused kinds add room for the stand-alone reductions they may have to run."""
    from ipoc_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "ipoc.h")).read()
    ctrl = int(re.search(r"#define IPOC_WS_CONTROL_BYTES (\d+)", hdr).group(1))
    for kind in (_lib.WS_NEWTON_STEP, _lib.WS_LQT_BWD, _lib.WS_LQT_FWD, _lib.WS_AFFINE_SCAN):
        assert L.ipoc_workspace_bytes(kind, 17, 4, 1, 1) > ctrl
    step = L.ipoc_workspace_bytes(_lib.WS_NEWTON_STEP, 100000, 4, 1, 1)
    att = L.ipoc_workspace_bytes(_lib.WS_NEWTON_ATTEMPT, 100000, 4, 2, 1)
    red = L.ipoc_workspace_bytes(_lib.WS_REDUCTIONS, 100000, 2, 2, 1)
    assert att == step + red
    assert L.ipoc_workspace_bytes(_lib.WS_COSTATES, 1000, 9, 1, 1) == 0        # unsupported nx stays 0


def test_plant_descriptor_requires_all_five_callables():
    """the fused plant kernels hard-code dynamics, costs, goal and box, so an OCP in which any callable of a
    built-in problem was replaced must NOT take the plant fast path."""
    import torch
    from ipoc_b200 import plants, problems
    from ipoc_b200.optimal_control_problem import OCP
    ocp = problems.make_cartpole(0.01)
    assert plants.plant_of(ocp) is not None and plants.plant_of(ocp)["name"] == "cartpole"
    mine = lambda x, u, bp: (x * x).sum() + (u * u).sum()
    assert plants.plant_of(ocp._replace(stage_cost=mine)) is None
    assert plants.plant_of(OCP(ocp.dynamics, lambda x, u: u - 1.0, ocp.stage_cost, ocp.final_cost, ocp.total_cost)) is None
    other = problems.make_cartpole(0.02)
    assert plants.plant_of(OCP(ocp.dynamics, *other[1:])) is None               # mixed instances: Ts / bound may differ
    assert plants.plant_of(other)["Ts"] == 0.02


def test_time_segments_must_be_non_empty_but_batch_shards_may_be():
    from ipoc_b200 import sharded
    with pytest.raises(ValueError):
        sharded.segment_bounds(3, 8)
    assert sharded.shard_batch(3, 7, 8) == (3, 3)

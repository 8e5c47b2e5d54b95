"""GPU parity tests proper: the sm_100a kernels (through the C ABI / ctypes) against the CPU
oracle on the same seeded inputs, against the golden fixtures produced by the reference's own
source, and — at full size — through size-independent properties (KKT residuals of the Newton
step, agreement of differently-chunked scans).

Tolerances: the north-star asks for <= 1e-9 relative on states, controls and cost and an identical
Newton iteration count.  Kernel-vs-oracle comparisons here are held to much tighter bounds
(1e-11 .. 1e-12) where the conditioning allows."""
import ctypes
import numpy as np
import pytest
import torch

from helpers import STEP_FIXTURES, derivs_from_golden, relerr, random_lq
from oracle import noc_np, paroc_np

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)


def N_(t):
    return t.detach().cpu().numpy()


@pytest.fixture(autouse=True)
def _reset_tuning():
    from ipoc_b200 import _lib, plants
    _lib.lib().ipoc_set_tuning(0, 0, 0)
    _lib.lib().ipoc_set_hier(1, 0, 0)
    _lib.lib().ipoc_set_literal_lqt(0)
    plants.ENABLED = True
    yield
    _lib.lib().ipoc_set_tuning(0, 0, 0)
    _lib.lib().ipoc_set_hier(1, 0, 0)
    _lib.lib().ipoc_set_literal_lqt(0)
    plants.ENABLED = True


def oracle_newton(fx, fu, ru, Q, R, M, reg):
    nu = R.shape[1]
    lqt = noc_np.noc_to_lqt(ru, Q, R + reg * np.eye(nu)[None], M, fx, fu)
    Kx, d, S, v, pred, feas = paroc_np.par_bwd_pass(lqt)
    du, dx = paroc_np.par_fwd_pass(lqt, np.zeros(fx.shape[1]), Kx, d)
    return dx, du, Kx, d, pred, feas


def test_library_loaded_and_supported():
    from ipoc_b200 import _lib
    L = _lib.lib()
    assert L.ipoc_version() >= 100
    assert L.ipoc_supported(2, 1) and L.ipoc_supported(4, 1) and L.ipoc_supported(5, 3) and L.ipoc_supported(8, 4)
    assert not L.ipoc_supported(3, 4) and not L.ipoc_supported(8, 5) and not L.ipoc_supported(9, 1)   # nu <= min(nx, 4)
    assert b"no CPU fallback" in L.ipoc_strerror(-1)


@pytest.mark.parametrize("nx,nu", [(nx, nu) for nx in range(1, 9) for nu in range(1, min(nx, 4) + 1)])
@pytest.mark.parametrize("N", [1, 2, 3, 31, 32, 33, 500])
def test_newton_step_vs_oracle(nx, nu, N):
    from ipoc_b200 import noc
    rng = np.random.default_rng(1000 * nx + 100 * nu + N)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    reg = 0.37
    dxo, duo, Kxo, do, predo, feaso = oracle_newton(fx, fu, ru, Q, R, M, reg)
    dx, du, Kx, d, pred, feas = noc.newton_step(T(fx), T(fu), T(ru), T(Q), T(R), T(M), T([reg]))
    tol = 1e-11 if nx <= 4 else 1e-10
    assert relerr(N_(Kx), Kxo) < tol and relerr(N_(d), do) < tol
    assert relerr(N_(dx), dxo) < tol and relerr(N_(du), duo) < tol
    assert abs(float(pred) - predo) <= tol * abs(predo)
    assert bool(feas[0]) == bool(feaso)


@pytest.mark.parametrize("tuning", [(1, 2, 4), (3, 3, 8), (4, 4, 32), (16, 8, 64), (0, 0, 0)])
@pytest.mark.parametrize("nx,nu,N", [(2, 1, 1000), (4, 1, 10000), (4, 2, 777)])
def test_newton_step_all_hierarchy_shapes(tuning, nx, nu, N):
    """Every way of cutting the horizon (leaf chunk / mid fan-in / top width) gives the same step."""
    from ipoc_b200 import noc, _lib
    rng = np.random.default_rng(7 * nx + N)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    dxo, duo, Kxo, do, predo, feaso = oracle_newton(fx, fu, ru, Q, R, M, 0.05)
    _lib.lib().ipoc_set_tuning(*tuning)
    dx, du, Kx, d, pred, feas = noc.newton_step(T(fx), T(fu), T(ru), T(Q), T(R), T(M), T([0.05]))
    assert relerr(N_(dx), dxo) < 1e-10 and relerr(N_(du), duo) < 1e-10
    assert relerr(N_(Kx), Kxo) < 1e-10 and relerr(N_(d), do) < 1e-10
    assert abs(float(pred) - predo) <= 1e-10 * abs(predo) and bool(feas[0]) == bool(feaso)


@pytest.mark.parametrize("literal", [0, 1])
@pytest.mark.parametrize("name", STEP_FIXTURES)
def test_newton_step_vs_reference_fixtures(golden, name, literal):
    """Real pendulum / cartpole linearisations; expected values computed by the reference's own
    sequential Newton step (ref noc/seq_interior_point_newton.py:42-90) run from its source.
    Both `noc_to_lqt` variants of the kernel (closed form / literal) are held to the same bound."""
    from ipoc_b200 import noc, _lib
    from ipoc_b200.optimal_control_problem import Derivatives
    _lib.lib().ipoc_set_literal_lqt(literal)
    g = golden(name)
    d = Derivatives(*(T(g["d_" + f]) for f in Derivatives._fields))
    dx, du, pred, feas, ru = noc.par_Newton(T(g["states"]), d, float(g["reg_param"]), T(g["ref_ru"]), T(g["ref_Q"]),
                                            T(g["ref_R"]), T(g["ref_M"]))
    assert relerr(N_(dx), g["ref_seq_dx"]) < 1e-9 and relerr(N_(du), g["ref_seq_du"]) < 1e-9
    assert relerr(N_(dx), g["refp_dx"]) < 1e-10 and relerr(N_(du), g["refp_du"]) < 1e-10
    assert abs(float(pred) - float(g["ref_seq_dV"])) <= 1e-10 * abs(float(g["ref_seq_dV"]))
    assert bool(feas) == bool(g["ref_seq_convex"])
    # costates (K1) and LQ parameters
    lam = noc.affine_scan(d.fx, d.cx, T(g["ref_lamT"]), reverse=True, transpose=True)
    assert relerr(N_(lam), g["ref_costates_par"]) < 1e-12
    ru2, Q2, R2, M2 = noc.compute_lqr_params(lam, d)
    assert relerr(N_(Q2), g["ref_Q"]) < 1e-12 and relerr(N_(ru2), g["ref_ru"]) < 1e-12
    # K4 reductions on the stepped trajectory
    hu, cn, fe = noc.reductions(ru=T(g["ref_ru"]), cu=d.cu, cons=T(g["ref_new_cons"]))
    assert float(hu) == float(np.max(np.abs(g["ref_ru"])))
    assert abs(float(cn) - np.linalg.norm(g["d_cu"].ravel())) < 1e-13 * float(cn)
    assert bool(fe[0]) == bool(g["ref_new_feasible"])


@pytest.mark.parametrize("reverse,transpose", [(False, False), (True, True), (True, False), (False, True)])
@pytest.mark.parametrize("nx,N", [(2, 1), (2, 500), (4, 33), (4, 10001), (3, 257), (8, 100)])
def test_affine_scan_vs_serial(reverse, transpose, nx, N):
    from ipoc_b200 import noc
    rng = np.random.default_rng(nx + N)
    F = np.eye(nx) + (2.0 / max(N, 8)) * rng.standard_normal((N, nx, nx))
    c = rng.standard_normal((N, nx))
    seed = rng.standard_normal(nx)
    out = np.zeros((N + 1, nx))
    Fe = np.swapaxes(F, 1, 2) if transpose else F
    if reverse:
        out[N] = seed
        for k in range(N - 1, -1, -1):
            out[k] = Fe[k] @ out[k + 1] + c[k]
    else:
        out[0] = seed
        for k in range(N):
            out[k + 1] = Fe[k] @ out[k] + c[k]
    got = noc.affine_scan(T(F), T(c), T(seed), reverse=reverse, transpose=transpose)
    assert relerr(N_(got), out) < 1e-11


@pytest.mark.parametrize("nx,nu,N", [(2, 1, 5), (2, 1, 600), (4, 1, 129), (4, 2, 64), (3, 1, 40)])
def test_raw_lqt_api_vs_oracle(nx, nu, N):
    """`par_bwd_pass` / `par_fwd_pass` with everything switched on: c != 0, rT != 0, r, s != 0,
    x0 != 0, non-identity H and Z."""
    from ipoc_b200.paroc import LQT, par_bwd_pass, par_fwd_pass
    rng = np.random.default_rng(nx * 31 + nu * 7 + N)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    H = np.eye(nx) + 0.1 * rng.standard_normal((N, nx, nx))
    Z = np.eye(nu) + 0.1 * rng.standard_normal((N, nu, nu))
    lq = paroc_np.LQT(fx, fu, 0.05 * rng.standard_normal((N, nx)), Q[0] + np.eye(nx),
                      np.eye(nx) + 0.1 * rng.standard_normal((nx, nx)), rng.standard_normal(nx), Q, H,
                      rng.standard_normal((N, nx)), R, Z, rng.standard_normal((N, nu)), M)
    Kxo, do, So, vo, predo, feaso = paroc_np.par_bwd_pass(lq)
    x0 = rng.standard_normal(nx)
    uo, xo = paroc_np.par_fwd_pass(lq, x0, Kxo, do)
    lqt = LQT(*(T(a) for a in lq))
    Kx, d, S, v, pred, feas = par_bwd_pass(lqt)
    u, x = par_fwd_pass(lqt, T(x0), Kx, d)
    for got, exp in ((Kx, Kxo), (d, do), (S, So), (v, vo), (u, uo), (x, xo)):
        assert relerr(N_(got), exp) < 1e-10
    assert abs(float(pred) - predo) <= 1e-10 * abs(predo) and bool(feas) == bool(feaso)


def test_mpc_example_par_equals_seq():
    """BASELINE config 3 as written (ref examples/linear_mpc_parallel.py:24-81): receding-horizon
    loop of par_bwd_pass + par_fwd_pass, T = 5; the reference's intent is par == seq."""
    from ipoc_b200 import problems
    from ipoc_b200.paroc import LQT, par_bwd_pass, par_fwd_pass
    fields, x0 = problems.make_mpc_lqt_terms(T=5, device=DEV)
    lqt = LQT(*fields)
    lq_np = paroc_np.LQT(*(N_(f) for f in fields))
    steps = 300
    x, xs_par, us_par = x0, [], []
    for _ in range(steps):
        Kx, d, _, _, _, _ = par_bwd_pass(lqt)
        u_par, x_par = par_fwd_pass(lqt, x, Kx, d)
        x = x_par[1]
        xs_par.append(x)
        us_par.append(u_par[0])
    xs_par, us_par = N_(torch.stack(xs_par)), N_(torch.stack(us_par))
    xk, xs_seq, us_seq = N_(x0), [], []
    Kxs, ds, _, _ = paroc_np.seq_bwd_pass(lq_np)
    for _ in range(steps):
        u_seq, x_seq = paroc_np.seq_fwd_pass(lq_np, xk, Kxs, ds)
        xk = x_seq[1]
        xs_seq.append(xk)
        us_seq.append(u_seq[0])
    assert relerr(xs_par, np.array(xs_seq)) < 1e-10 and relerr(us_par, np.array(us_seq)) < 1e-10
    # the graph-unrolled loop (chunks of 7 steps, so 300 steps cross many chunk boundaries) and its serial twin
    from ipoc_b200.mpc import MpcLoop
    for serial in (False, True):
        xs_g, us_g = MpcLoop(lqt, unroll=7, serial=serial).run(x0, steps)
        assert xs_g.shape == (steps, 2) and us_g.shape == (steps, 1)
        if not serial:
            assert np.array_equal(N_(xs_g), xs_par) and np.array_equal(N_(us_g), us_par)
        assert relerr(N_(xs_g), np.array(xs_seq)) < 1e-10 and relerr(N_(us_g), np.array(us_seq)) < 1e-10
    xs_e, us_e = MpcLoop(lqt, unroll=16).run(x0, 40, use_graph=False)
    assert np.array_equal(N_(xs_e), xs_par[:40]) and np.array_equal(N_(us_e), us_par[:40])


@pytest.mark.parametrize("nx,nu,N,B", [(2, 1, 100, 7), (4, 1, 64, 33), (4, 1, 1000, 300), (2, 1, 1000, 40000)])
def test_batched_equals_unbatched(nx, nu, N, B):
    """Independent OCPs stacked on a batch axis give each problem the result of solving it alone."""
    from ipoc_b200 import noc
    rng = np.random.default_rng(B)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu, batch=B)
    reg = 0.1 + rng.random(B)
    dx, du, Kx, d, pred, feas = noc.newton_step(T(fx), T(fu), T(ru), T(Q), T(R), T(M), T(reg))
    for b in list(range(min(B, 3))) + [B - 1]:
        dxo, duo, Kxo, do, predo, feaso = oracle_newton(fx[b], fu[b], ru[b], Q[b], R[b], M[b], reg[b])
        assert relerr(N_(dx[b]), dxo) < 1e-10 and relerr(N_(du[b]), duo) < 1e-10
        assert abs(float(pred[b]) - predo) <= 1e-10 * abs(predo) and bool(feas[b]) == bool(feaso)
        dx1, du1, _, _, pred1, _ = noc.newton_step(T(fx[b]), T(fu[b]), T(ru[b]), T(Q[b]), T(R[b]), T(M[b]),
                                                   T(reg[b:b + 1]))
        assert relerr(N_(dx[b]), N_(dx1)) < 1e-11 and relerr(N_(du[b]), N_(du1)) < 1e-11


@pytest.mark.parametrize("P", [2, 3, 8])
@pytest.mark.parametrize("nx,nu,N", [(4, 1, 1003), (2, 1, 64)])
def test_time_sharded_virtual_ranks(P, nx, nu, N):
    """The multi-GPU time-sharded algorithm (reduce -> exchange carries -> seeded scan) run with P
    virtual ranks on one GPU equals the single-device scan."""
    from ipoc_b200 import noc, sharded
    rng = np.random.default_rng(P + N)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    reg = T([0.2])
    args = [T(a) for a in (fx, fu, ru, Q, R, M)]
    dx1, du1, Kx1, d1, pred1, feas1 = noc.newton_step(*args, reg)
    dx, du, Kx, d, pred, feas = sharded.newton_step_virtual_ranks(*args, reg, P)
    assert relerr(N_(Kx), N_(Kx1)) < 1e-11 and relerr(N_(d), N_(d1)) < 1e-11
    assert relerr(N_(dx), N_(dx1)) < 1e-11 and relerr(N_(du), N_(du1)) < 1e-11
    assert abs(float(pred) - float(pred1)) <= 1e-11 * abs(float(pred1)) and feas == bool(feas1[0])


def _kkt_residuals(fx, fu, ru, Q, R, M, reg, dx, du):
    """Size-independent check: (dx, du) solves the regularised LQ Newton system
    (dynamics + stationarity with costates p_k = Q dx + M du + fx' p_{k+1}, p_N = Q[0] dx_N)."""
    N, nu = fx.shape[0], R.shape[1]
    dyn = dx[1:] - np.einsum("tij,tj->ti", fx, dx[:-1]) - np.einsum("tij,tj->ti", fu, du)
    p = Q[0] @ dx[N]
    stat = np.zeros((N, nu))
    for k in range(N - 1, -1, -1):
        stat[k] = ru[k] + (R[k] + reg * np.eye(nu)) @ du[k] + M[k].T @ dx[k] + fu[k].T @ p
        p = Q[k] @ dx[k] + M[k] @ du[k] + fx[k].T @ p
    scale = max(1.0, np.max(np.abs(ru)))
    return np.max(np.abs(dyn)), np.max(np.abs(stat)) / scale, np.max(np.abs(dx[0]))


@pytest.mark.parametrize("nx,nu,N", [(4, 1, 100000), (2, 1, 300000)])
def test_newton_step_kkt_at_full_size(nx, nu, N):
    from ipoc_b200 import noc
    rng = np.random.default_rng(N)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu, dt=1e-3)
    dx, du, Kx, d, pred, feas = noc.newton_step(T(fx), T(fu), T(ru), T(Q), T(R), T(M), T([0.3]))
    r_dyn, r_stat, r_x0 = _kkt_residuals(fx, fu, ru, Q, R, M, 0.3, N_(dx), N_(du))
    assert r_x0 == 0.0 and r_dyn < 1e-9 and r_stat < 1e-8
    assert bool(feas[0]) and float(pred) < 0


def test_nonconvex_and_nan_are_data_not_errors():
    from ipoc_b200 import noc
    rng = np.random.default_rng(3)
    fx, fu, ru, Q, R, M = random_lq(rng, 50, 2, 1)
    R[17] = -5.0   # G < 0 at one step -> infeasible flag, no exception
    _, _, _, _, pred, feas = noc.newton_step(T(fx), T(fu), T(ru), T(Q), T(R), T(M), T([0.0]))
    assert not bool(feas[0])
    ru[3] = np.nan
    _, du, _, _, pred, feas = noc.newton_step(T(fx), T(fu), T(ru), T(Q), T(R), T(M), T([0.0]))
    assert np.isnan(float(pred))


def test_accept_update_matches_reference_rule():
    from ipoc_b200 import noc
    cost = T([10.0, 10.0, 10.0, 10.0, 10.0])
    new = T([9.0, 11.0, 9.0, 9.5, np.nan])
    traj = torch.tensor([1, 1, 0, 1, 1], dtype=torch.int32, device=DEV)
    pred = T([-2.0, -2.0, -2.0, -0.4, -1.0])
    bwd = torch.tensor([1, 1, 1, 0, 1], dtype=torch.int32, device=DEV)
    rp = T([1.0, 1.0, 1.0, 1.0, 1.0])
    ri = T([2.0, 2.0, 4.0, 2.0, 2.0])
    succ, gain = noc.accept_update(cost, new, traj, pred, bwd, rp, ri)
    exp_rp, exp_ri, exp_s = [], [], []
    for c, n, t, p, b, r, i in zip([10.0] * 5, [9.0, 11.0, 9.0, 9.5, np.nan], [1, 1, 0, 1, 1],
                                   [-2.0, -2.0, -2.0, -0.4, -1.0], [1, 1, 1, 0, 1], [1.0] * 5, [2.0, 2.0, 4.0, 2.0, 2.0]):
        nc = n if t else np.inf
        rho = (nc - c) / p
        ok = bool(rho > 0) and bool(b)
        r2 = r * max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3) if ok else r * i
        exp_rp.append(min(max(r2, 1e-16), 1e16))
        exp_ri.append(2.0 if ok else 2 * i)
        exp_s.append(int(ok))
    assert N_(succ).tolist() == exp_s
    assert np.allclose(N_(rp), exp_rp, rtol=1e-15) and np.allclose(N_(ri), exp_ri, rtol=0)


@pytest.mark.parametrize("literal", [0, 1])
@pytest.mark.parametrize("nx,nu,N", [(4, 1, 300), (2, 2, 77)])
def test_literal_and_closed_form_lqt_agree(nx, nu, N, literal):
    from ipoc_b200 import noc, _lib
    rng = np.random.default_rng(N)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu, coupling=1.0)
    dxo, duo, Kxo, do, predo, feaso = oracle_newton(fx, fu, ru, Q, R, M, 0.2)
    _lib.lib().ipoc_set_literal_lqt(literal)
    dx, du, Kx, d, pred, feas = noc.newton_step(T(fx), T(fu), T(ru), T(Q), T(R), T(M), T([0.2]))
    assert relerr(N_(dx), dxo) < 1e-11 and relerr(N_(du), duo) < 1e-11
    assert abs(float(pred) - predo) <= 1e-11 * abs(predo) and bool(feas[0]) == bool(feaso)


@pytest.mark.parametrize("name", ["solve_pendulum_N20", "solve_linear_N40", "solve_cartpole_N40",
                                  "solve_pendulum_N100", "solve_pendulum_N500"])
def test_full_solve_matches_reference(golden, name):
    """End to end through the reference-facing API: same optimal controls (<= 1e-9 relative) and the
    SAME Newton iteration count as the reference's driver run from its own source (BASELINE config 1
    is solve_pendulum_N500)."""
    from ipoc_b200 import noc, problems
    g = golden(name)
    N = g["u0"].shape[0]
    if "pendulum" in name:
        ocp = problems.make_pendulum(1.0 / N)
    elif "cartpole" in name:
        ocp = problems.make_cartpole(1.0 / N)
    else:
        ocp = problems.make_linear_demo(0.1)
    u, its = noc.par_interior_point_optimal_control(ocp, T(g["u0"]), T(g["x0"]))
    assert its == int(g["refp_iterations"])
    assert relerr(N_(u), g["refp_opt_u"]) < 1e-9
    # the default above is the device-resident loop; the host-steered graphs must give the same iterate
    u3, its3 = noc.par_interior_point_optimal_control(ocp, T(g["u0"]), T(g["x0"]), use_graphs="host")
    assert its3 == its and relerr(N_(u3), N_(u)) < 1e-12
    if N <= 100:   # the eager (graph-free) driver is the same sequence of statements
        u2, its2 = noc.par_interior_point_optimal_control(ocp, T(g["u0"]), T(g["x0"]), use_graphs=False)
        assert its2 == its and relerr(N_(u2), N_(u)) < 1e-12


def test_seq_twins_equal_par():
    """`seq_bwd_pass` / `seq_fwd_pass` (serial comparators of the MPC example) vs the parallel scans."""
    from ipoc_b200.paroc import LQT, par_bwd_pass, par_fwd_pass, seq_bwd_pass, seq_fwd_pass
    rng = np.random.default_rng(8)
    N, nx, nu = 333, 4, 1
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    lq = noc_np.noc_to_lqt(ru, Q, R, M, fx, fu)._replace(c=0.01 * rng.standard_normal((N, nx)))
    lqt = LQT(*(T(a) for a in lq))
    Kx, d, S, v, _, _ = par_bwd_pass(lqt)
    Kx2, d2, S2, v2 = seq_bwd_pass(lqt)
    x0 = T(rng.standard_normal(nx))
    u, x = par_fwd_pass(lqt, x0, Kx, d)
    u2, x2 = seq_fwd_pass(lqt, x0, Kx2, d2)
    for a, b in ((Kx, Kx2), (d, d2), (S, S2), (v, v2), (u, u2), (x, x2)):
        assert relerr(N_(a), N_(b)) < 1e-10
    Kxo, do, So, vo = paroc_np.seq_bwd_pass(lq)
    assert relerr(N_(Kx2), Kxo) < 1e-10 and relerr(N_(S2), So) < 1e-10


@pytest.mark.parametrize("problem,N,B", [("pendulum", 30, 6), ("cartpole", 24, 5)])
def test_batched_solver_equals_per_problem_solves(problem, N, B):
    """Config 5 semantics: every member of a batched solve gets the iterate and the Newton iteration
    count of solving it alone (vmap-of-while_loop semantics)."""
    from ipoc_b200 import noc, problems, batched
    rng = np.random.default_rng(3)
    if problem == "pendulum":
        ocp, x0 = problems.make_pendulum(1.0 / N), problems.pendulum_x0().numpy()
    else:
        ocp, x0 = problems.make_cartpole(1.0 / N), problems.cartpole_x0().numpy()
    x0s = x0[None] + 0.1 * rng.standard_normal((B, x0.shape[0]))
    u0s = 0.1 * rng.standard_normal((B, N, 1))
    ub, itb = batched.par_interior_point_optimal_control_batched(ocp, T(u0s), T(x0s))
    for b in range(B):
        u1, it1 = noc.par_interior_point_optimal_control(ocp, T(u0s[b]), T(x0s[b]))
        assert int(itb[b]) == it1
        assert relerr(N_(ub[b]), N_(u1)) < 1e-9
    # the graph-replayed tail of the attempt loops (<= 8 members left) against the all-eager loop
    ue, ite = batched.par_interior_point_optimal_control_batched(ocp, T(u0s), T(x0s), use_graphs=False)
    assert torch.equal(ite, itb) and relerr(N_(ue), N_(ub)) < 1e-9


@pytest.mark.parametrize("problem,N", [("cartpole", 3000), ("pendulum", 2500)])
def test_parallel_rollout_equals_serial(problem, N):
    """Newton-on-the-rollout with the forward affine scan reproduces the serial rollout."""
    from ipoc_b200 import problems, utils
    rng = np.random.default_rng(N)
    ocp = problems.make_cartpole(1.0 / N) if problem == "cartpole" else problems.make_pendulum(1.0 / N)
    x0 = (problems.cartpole_x0 if problem == "cartpole" else problems.pendulum_x0)(device=DEV)
    u = T(2.0 * rng.standard_normal((N, 1)))
    xs = utils.rollout(ocp.dynamics, u, x0)
    xp, its = utils.rollout_parallel(ocp.dynamics, u, x0)
    assert 0 < its < 50
    assert relerr(N_(xp), N_(xs)) < 1e-11


def test_reference_helper_functions(golden):
    """`noc_to_lqt`, `check_traj_feasibility`, `par_costates`, `compute_derivatives` keep the reference's
    signatures and values (torch tensors for jnp arrays)."""
    from ipoc_b200 import noc, problems
    g = golden("step_cartpole_N100")
    N = g["controls"].shape[0]
    ocp = problems.make_cartpole(1.0 / N)
    x, u = T(g["states"]), T(g["controls"])
    d = noc.compute_derivatives(ocp, x, u, float(g["bp"]))
    for f in d._fields:
        assert np.max(np.abs(N_(getattr(d, f)) - g["d_" + f])) <= 1e-11 * max(1.0, np.max(np.abs(g["d_" + f]))), f
    lam = noc.par_costates(ocp, x[-1], d)
    assert relerr(N_(lam), g["ref_costates_par"]) < 1e-12
    ru, Q, R, M = noc.compute_lqr_params(lam, d)
    lqt = noc.noc_to_lqt(ru, Q, R + float(g["ref_reg"]) * torch.eye(1, dtype=torch.float64, device=DEV), M, d.fx, d.fu)
    assert relerr(N_(lqt.r), g["ref_lqt_r"]) < 1e-10 and relerr(N_(lqt.s), g["ref_lqt_s"]) < 1e-10
    assert lqt.H.shape == (N, 4, 4) and lqt.Z.shape == (N, 1, 1) and torch.equal(lqt.XT, Q[0])
    dx, du, pred, feas, _ = noc.par_Newton(x, d, float(g["reg_param"]), ru, Q, R, M)
    assert bool(noc.check_traj_feasibility(ocp, x + dx, u + du)) == bool(g["ref_new_feasible"])
    nc, ncr = float(ocp.total_cost(x + dx, u + du, float(g["bp"]))), float(g["ref_new_cost"])
    # an infeasible trial gives log(negative) = NaN in the reference too; the `where` of :159-163 masks it
    assert (np.isnan(nc) and np.isnan(ncr)) or abs(nc - ncr) <= 1e-9 * abs(ncr)


def test_host_buffer_entry_point_and_error_codes():
    """`ipoc_newton_step_host_f64` (pinned host buffers) equals the resident call; bad calls return codes."""
    import ctypes
    from ipoc_b200 import noc, _lib
    L = _lib.lib()
    rng = np.random.default_rng(4)
    N, nx, nu = 1500, 4, 1
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    host = [torch.as_tensor(np.ascontiguousarray(a)).pin_memory() for a in (fx, fu, ru, Q, R, M)]
    reg = torch.tensor([0.4], dtype=torch.float64).pin_memory()
    dx = torch.empty(N + 1, nx, dtype=torch.float64).pin_memory()
    du = torch.empty(N, nu, dtype=torch.float64).pin_memory()
    pred = torch.empty(1, dtype=torch.float64).pin_memory()
    feas = torch.empty(1, dtype=torch.int32).pin_memory()
    nbytes = L.ipoc_newton_step_host_scratch_bytes(N, nx, nu, 1)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    hp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = L.ipoc_newton_step_host_f64(N, nx, nu, 1, *(hp(t) for t in host), hp(reg), hp(dx), hp(du), hp(pred), hp(feas),
                                     hp(scratch), nbytes, _lib.stream_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    dx2, du2, _, _, pred2, feas2 = noc.newton_step(*(T(a) for a in (fx, fu, ru, Q, R, M)), T([0.4]))
    assert torch.equal(dx.to(DEV), dx2) and torch.equal(du.to(DEV), du2) and float(pred) == float(pred2)
    # workspace too small / unsupported dimension / misaligned pointer are error CODES, not crashes
    dev = [T(a) for a in (fx, fu, ru, Q, R, M)]
    o = dict(dtype=torch.float64, device=DEV)
    outs = [torch.empty(N + 1, nx, **o), torch.empty(N, nu, **o), torch.empty(N, nu, nx, **o), torch.empty(N, nu, **o),
            torch.empty(1, **o), torch.empty(1, dtype=torch.int32, device=DEV)]
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device=DEV)
    dp = lambda t: ctypes.c_void_p(t.data_ptr())
    regd = T([0.4])
    call = lambda nx_, wsb, fxp: L.ipoc_newton_step_f64(N, nx_, nu, 1, fxp, *(dp(t) for t in dev[1:]), dp(regd),
                                                        *(dp(t) for t in outs), dp(ws), wsb, _lib.stream_ptr())
    assert call(nx, 64, dp(dev[0])) == -2                      # IPOC_EWORKSPACE
    assert call(9, 1 << 20, dp(dev[0])) == -1                  # IPOC_EUNSUPPORTED_DIM
    assert call(nx, 1 << 20, ctypes.c_void_p(dev[0].data_ptr() + 8)) == -6   # IPOC_EALIGN
    assert b"workspace" in L.ipoc_strerror(-2)


@pytest.mark.parametrize("name", STEP_FIXTURES)
def test_plant_kernels_vs_reference_fixtures(golden, name):
    """Built-in plant kernels (in-register forward-mode autodiff, cost, rollout) against the values the
    reference's own definitions produced (golden fixtures) and against the host-framework autodiff."""
    from ipoc_b200 import noc, problems, plants
    g = golden(name)
    N = g["controls"].shape[0]
    ocp = problems.make_pendulum(1.0 / N) if "pendulum" in name else problems.make_cartpole(1.0 / N)
    plant = plants.plant_of(ocp)
    assert plant is not None
    x, u, bp = T(g["states"]), T(g["controls"]), float(g["bp"])
    d, lamT = plants.derivatives(plant, x, u, bp)
    for f in d._fields:
        ref = g["d_" + f]
        assert N_(getattr(d, f)).shape == ref.shape, f
        assert np.max(np.abs(N_(getattr(d, f)) - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))), f
    assert relerr(N_(lamT), g["ref_lamT"]) < 1e-13
    cost, feas = plants.cost(plant, x, u, bp)
    assert abs(float(cost) - float(g["ref_cost"])) <= 1e-12 * abs(float(g["ref_cost"])) and bool(feas[0])
    if "warm" not in name:
        xr = plants.rollout(plant, u, T(g["x0"]))
        assert relerr(N_(xr), g["states"]) < 1e-13
    # fused first-order + Hamiltonian passes == compute_lqr_params of the reference
    fx, fu, cx, cu, lamT2 = plants.linearize(plant, x, u, bp)
    assert torch.equal(fx, d.fx) and torch.equal(fu, d.fu) and torch.equal(cx, d.cx) and torch.equal(cu, d.cu)
    ru, Q, R, M = plants.hamiltonian(plant, x, u, T(g["ref_costates_par"]), bp)
    for got, key in ((ru, "ref_ru"), (Q, "ref_Q"), (R, "ref_R"), (M, "ref_M")):
        assert np.max(np.abs(N_(got) - g[key])) <= 1e-12 * max(1.0, np.max(np.abs(g[key]))), key
    ru2, Q2, R2, M2 = noc.compute_lqr_params(T(g["ref_costates_par"]), d)     # streaming A3 kernel on Derivatives
    assert relerr(N_(Q2), g["ref_Q"]) < 1e-13 and relerr(N_(ru2), g["ref_ru"]) < 1e-13
    # batched call == per-problem call
    xb, ub = torch.stack((x, x)), torch.stack((u, u * 0.5))
    db, _ = plants.derivatives(plant, xb, ub, bp)
    assert torch.equal(db.fxx[0], d.fxx) and torch.equal(db.cuu[0], d.cuu)


@pytest.mark.parametrize("name", ["solve_pendulum_N100", "solve_cartpole_N40"])
def test_full_solve_with_autodiff_path(golden, name):
    """The general (host-framework autodiff) path, with the plant kernels switched off, still reproduces the
    reference driver's iterate and iteration count."""
    from ipoc_b200 import noc, problems, plants
    plants.ENABLED = False
    g = golden(name)
    N = g["u0"].shape[0]
    ocp = problems.make_pendulum(1.0 / N) if "pendulum" in name else problems.make_cartpole(1.0 / N)
    assert plants.plant_of(ocp) is None
    u, its = noc.par_interior_point_optimal_control(ocp, T(g["u0"]), T(g["x0"]))
    assert its == int(g["refp_iterations"]) and relerr(N_(u), g["refp_opt_u"]) < 1e-9


def _custom_ocp(Ts, bad_constants=False):
    """A user-defined 3-state problem (not one of the built-ins): general autodiff path, nx = 3 kernels."""
    from ipoc_b200.optimal_control_problem import OCP
    from ipoc_b200.utils import euler
    from torch.func import vmap

    def ode(x, u):
        return torch.hstack((x[1], -x[0] - 0.1 * x[1] - 0.1 * x[0] ** 3 + u[0], torch.sin(x[0]) - x[2]))

    def constraints(x, u):
        return torch.hstack((u - 2.0, -u - 2.0))

    def stage_cost(x, u, bp):
        if bad_constants:   # builds a tensor from a Python list on every call: H2D copy, not graph-capturable
            w = torch.tensor([1.0, 0.1, 0.5], dtype=x.dtype, device=x.device)
        else:
            w = torch.stack((x[0] * 0 + 1.0, x[0] * 0 + 0.1, x[0] * 0 + 0.5))
        return 0.5 * torch.sum(w * x * x) + 0.5 * 1e-2 * (u @ u) - bp * torch.sum(torch.log(-constraints(x, u)))

    def final_cost(x):
        return 0.5 * (x @ x)

    def total_cost(xs, us, bp):
        return final_cost(xs[-1]) + torch.sum(vmap(stage_cost, in_dims=(0, 0, None))(xs[:-1], us, bp))

    return OCP(euler(ode, Ts), constraints, stage_cost, final_cost, total_cost)


def test_user_defined_ocp_matches_oracle_driver():
    """End to end on a problem that is NOT built in: host-framework autodiff + nx=3 kernels + graphs vs the
    NumPy oracle driver (same iterate, same Newton iteration count)."""
    from ipoc_b200 import noc, plants
    from oracle.autodiff import Evaluator
    N = 30
    ocp = _custom_ocp(0.05)
    assert plants.plant_of(ocp) is None
    rng = np.random.default_rng(2)
    u0 = 0.1 * rng.standard_normal((N, 1))
    x0 = np.array([1.0, -0.5, 0.3])
    uo, ito = noc_np.par_interior_point_optimal_control(Evaluator(ocp), u0, x0)
    ug, itg = noc.par_interior_point_optimal_control(ocp, T(u0), T(x0))
    assert itg == ito and relerr(N_(ug), uo) < 1e-9
    ud, itd = noc.par_interior_point_optimal_control(ocp, T(u0), T(x0), use_graphs="device")   # device-resident loop
    assert itd == ito and relerr(N_(ud), N_(ug)) < 1e-12


def test_graph_capture_failure_falls_back_to_eager():
    from ipoc_b200 import noc
    import warnings
    N = 12
    ocp = _custom_ocp(0.05, bad_constants=True)
    rng = np.random.default_rng(3)
    u0, x0 = T(0.1 * rng.standard_normal((N, 1))), T([0.5, 0.0, 0.1])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ug, itg = noc.par_interior_point_optimal_control(ocp, u0, x0)
    assert any("capture" in str(m.message) for m in w)
    ue, ite = noc.par_interior_point_optimal_control(ocp, u0, x0, use_graphs=False)
    assert itg == ite and relerr(N_(ug), N_(ue)) < 1e-12


def test_cpu_tensors_are_rejected():
    from ipoc_b200 import noc, _lib
    rng = np.random.default_rng(0)
    fx, fu, ru, Q, R, M = (torch.as_tensor(a) for a in random_lq(rng, 8, 2, 1))
    with pytest.raises(_lib.IpocError):
        noc.newton_step(fx, fu, ru, Q, R, M, torch.tensor([0.1]))


def test_attempt_glue_kernels_vs_torch():
    """ipoc_attempt_begin / trial_point / attempt_commit against the few torch statements they replace
    (ref noc/par_interior_point_newton.py:116-117, 156-157, 174-182)."""
    from ipoc_b200 import noc
    rng = np.random.default_rng(11)
    B, N, nx, nu = 13, 37, 3, 2
    done = torch.as_tensor(rng.random(B) < 0.4, device=DEV)
    rp, cun = T(rng.random(B) + 0.1), T(rng.random(B))
    act = torch.full((B,), -7, dtype=torch.int32, device=DEV)
    reg = torch.zeros(B, dtype=torch.float64, device=DEV)
    noc.attempt_begin(done, rp, cun, act, reg)
    assert torch.equal(act, (~done).to(torch.int32)) and torch.equal(reg, rp * cun)
    noc.attempt_begin(None, rp, cun, act, reg)
    assert bool((act == 1).all())
    x, dx = T(rng.standard_normal((B, N + 1, nx))), T(rng.standard_normal((B, N + 1, nx)))
    u, du = T(rng.standard_normal((B, N, nu))), T(rng.standard_normal((B, N, nu)))
    tx, tu = torch.zeros_like(x), torch.zeros_like(u)
    noc.trial_point(x, dx, u, du, tx, tu)
    assert torch.equal(tx, x + dx) and torch.equal(tu, u + du)
    act = (~done).to(torch.int32)
    succ = torch.as_tensor(rng.random(B) < 0.5, device=DEV).to(torch.int32)
    inner = torch.as_tensor(rng.integers(0, 3, B), device=DEV)
    inner[0], inner[1] = 500, 499
    keep_x, keep_u = T(rng.standard_normal((B, N + 1, nx))), T(rng.standard_normal((B, N, nu)))
    kx0, ku0, inner0, done0 = keep_x.clone(), keep_u.clone(), inner.clone(), done.clone()
    noc.attempt_commit(act, succ, tx, tu, keep_x, keep_u, inner, done, max_attempts=500)
    m = act.bool()
    assert torch.equal(keep_x, torch.where(m[:, None, None], tx, kx0))
    assert torch.equal(keep_u, torch.where(m[:, None, None], tu, ku0))
    assert torch.equal(inner, inner0 + m.to(torch.int64))
    assert torch.equal(done, done0 | (m & ((succ != 0) | (inner > 500))))


@pytest.mark.parametrize("N", [300, 10000])
def test_host_arena_pass_equals_resident_pass(N):
    """`runner.HostNewtonPass` (pinned host arena -> overlapped H2D, the pass, D2H; one CUDA graph) returns
    bit-identical results to the resident `NewtonPass` on the same inputs."""
    from ipoc_b200 import workloads
    from ipoc_b200.runner import NewtonPass, HostNewtonPass
    w = workloads.newton_inputs("cartpole", N, DEV, seed=2, x0_noise=0.01)
    res = NewtonPass(w["fx"], w["fu"], w["cx"], w["cu"], w["lamT"], w["ru"], w["Q"], w["R"], w["M"], w["cons"])
    res.run()
    hp = HostNewtonPass(w, DEV)
    hp.capture()
    for v in hp.views.values():          # the graph must read what the host arena holds at replay time
        assert v.is_pinned()
    hp.d_in.zero_()
    hp.inner.rp.fill_(1.0)               # rp / r_inc evolve from pass to pass (A8): start from the same state
    hp.inner.r_inc.fill_(2.0)
    hp.replay()
    torch.cuda.synchronize()
    for k in ("lam", "dx", "du", "pred", "hu", "cu_norm", "gain", "rp", "r_inc"):
        assert torch.equal(hp.results[k].to(DEV).reshape(-1), getattr(res, k).reshape(-1)), k
    assert int(hp.results["bwd_feas"][0]) == int(res.bwd_feas[0]) and int(hp.results["success"][0]) == int(res.success[0])
    hp.views["ru"].mul_(2.0)             # change an input on the host: the next replay must see it
    hp.replay()
    torch.cuda.synchronize()
    res2 = NewtonPass(w["fx"], w["fu"], w["cx"], w["cu"], w["lamT"], 2.0 * w["ru"], w["Q"], w["R"], w["M"], w["cons"])
    res2.run()
    torch.cuda.synchronize()
    assert torch.equal(hp.results["hu"].to(DEV), res2.hu)


# ====================================================================== BASELINE-size parity (round 2)
BIG_STEP_FIXTURES = ["step_cartpole_N10000", "step_cartpole_N10000_warm"]


@pytest.mark.parametrize("use_plant", [True, False])
@pytest.mark.parametrize("name", BIG_STEP_FIXTURES)
def test_config2_step_vs_reference(golden, name, use_plant):
    """BASELINE config 2 (cartpole nx=4 nu=1 N=1e4, Ts=1e-4): the WHOLE chain of one Newton step — derivatives
    at the recorded iterate (plant kernels / host-framework autodiff), K1 costates, LQ parameters, K4, K2+K3 —
    against the outputs of the reference's own source (tests/golden/gen_golden.py): costates, ru/Q/R/M, the
    in-tree sequential step (`ref_seq_*`, pure reference code) and par_Newton (`refp_*`).  The warm fixture is
    an interior iterate whose Q = cxx + lam.fxx is indefinite on 2114 of the 10^4 steps."""
    from ipoc_b200 import noc, problems, plants
    g = golden(name)
    N = g["controls"].shape[0]
    plants.ENABLED = use_plant
    ocp = problems.make_cartpole(1.0 / N)
    x, u, bp = T(g["states"]), T(g["controls"]), float(g["bp"])
    assert (plants.plant_of(ocp) is not None) == use_plant
    cost, fx, fu, cu, ru, Q, R, M = noc.eval_iteration(ocp, x, u, bp)
    assert relerr(N_(ru), g["ref_ru"]) < 1e-11 and relerr(N_(Q), g["ref_Q"]) < 1e-11
    assert relerr(N_(R), g["ref_R"]) < 1e-11 and relerr(N_(M), g["ref_M"]) < 1e-11
    assert abs(float(cost) - float(g["ref_cost"])) <= 1e-12 * abs(float(g["ref_cost"]))
    hu, cu_norm, _ = noc.reductions(ru=ru, cu=cu)
    assert abs(float(cu_norm) - float(g["ref_cu_norm"])) <= 1e-12 * float(g["ref_cu_norm"])
    assert abs(float(hu) - np.max(np.abs(g["ref_ru"]))) <= 1e-11 * np.max(np.abs(g["ref_ru"]))
    reg = float(g["reg_param"]) * cu_norm
    dx, du, Kx, d, pred, feas = noc.newton_step(fx, fu, ru, Q, R, M, reg)
    assert relerr(N_(dx), g["ref_seq_dx"]) < 1e-9 and relerr(N_(du), g["ref_seq_du"]) < 1e-9
    assert relerr(N_(dx), g["refp_dx"]) < 1e-9 and relerr(N_(du), g["refp_du"]) < 1e-9
    assert abs(float(pred) - float(g["ref_seq_dV"])) <= 1e-9 * abs(float(g["ref_seq_dV"]))
    assert abs(float(pred) - float(g["refp_pred"])) <= 1e-9 * abs(float(g["refp_pred"]))
    assert bool(feas[0]) == bool(g["ref_seq_convex"]) == bool(g["refp_feasible"])
    # the same step on the reference's OWN LQ data (isolates K2+K3), both noc_to_lqt variants, and the
    # extended-precision serial oracle as the arbiter
    from ipoc_b200 import _lib
    from oracle import serial_ld
    args = [T(g[k]) for k in ("ref_ru", "ref_Q", "ref_R", "ref_M")]
    dxl, dul, _, _, dVl, cvx = serial_ld.seq_newton(N_(fx), N_(fu), g["ref_ru"], g["ref_Q"], g["ref_R"], g["ref_M"],
                                                   float(g["ref_reg"]))
    for literal in (0, 1):
        _lib.lib().ipoc_set_literal_lqt(literal)
        dx2, du2, _, _, pred2, feas2 = noc.newton_step(fx, fu, args[0], args[1], args[2], args[3], T([float(g["ref_reg"])]))
        assert relerr(N_(dx2), g["ref_seq_dx"]) < 1e-9 and relerr(N_(du2), g["ref_seq_du"]) < 1e-9
        assert relerr(N_(dx2), dxl) < 1e-9 and relerr(N_(du2), dul) < 1e-9
        assert abs(float(pred2) - dVl) <= 1e-9 * abs(dVl) and bool(feas2[0]) == cvx
    # trial point: cost / feasibility of the stepped trajectory (A7)
    new_cost, traj_feas = noc.eval_trial(ocp, x + dx, u + du, bp)
    assert bool(traj_feas[0]) == bool(g["ref_new_feasible"])
    if bool(g["ref_new_feasible"]):
        assert abs(float(new_cost) - float(g["ref_new_cost"])) <= 1e-9 * abs(float(g["ref_new_cost"]))


@pytest.mark.parametrize("name", ["solve_cartpole_N1000", "solve_cartpole_N10000"])
def test_config2_full_solve_matches_reference(golden, name):
    """Full IP solves of the cartpole at config 5's horizon (N = 1000) and config 2's (N = 1e4) against the
    reference's driver run from its own source on the shim: same Newton iteration count (114 / 143), controls
    within 1e-9 relative."""
    from ipoc_b200 import noc, problems
    g = golden(name)
    N = g["u0"].shape[0]
    ocp = problems.make_cartpole(1.0 / N)
    u, its = noc.par_interior_point_optimal_control(ocp, T(g["u0"]), T(g["x0"]))
    assert its == int(g["refp_iterations"])
    assert relerr(N_(u), g["refp_opt_u"]) < 1e-9
    if N <= 1000:   # user-OCP path (host-framework autodiff, host-steered graphs): same count, same iterate
        from ipoc_b200 import plants
        plants.ENABLED = False
        u2, its2 = noc.par_interior_point_optimal_control(ocp, T(g["u0"]), T(g["x0"]))
        assert its2 == its and relerr(N_(u2), g["refp_opt_u"]) < 1e-9


def test_config4_step_N1e6_vs_longdouble_serial_oracle():
    """Large end of BASELINE config 4: cartpole N = 1e6, Ts = 1e-6 (fu ~ Ts, C = B U^-1 B' ~ 1e-12, 10^6-fold
    products).  The CUDA scans (K1, K2+K3) are held to 1e-9 against an extended-precision SERIAL restatement of
    the reference's in-tree sequential step (oracle/serial_ld.c, x87 long double)."""
    from ipoc_b200 import noc, workloads
    from oracle import serial_ld
    N = 1_000_000
    w = workloads.newton_inputs("cartpole", N, DEV, seed=1)
    lam = noc.affine_scan(w["fx"], w["cx"], w["lamT"], reverse=True, transpose=True)
    lam_ld = serial_ld.seq_costates(N_(w["fx"]), N_(w["cx"]), N_(w["lamT"]))
    assert relerr(N_(lam), lam_ld) < 1e-9
    _, cu_norm, _ = noc.reductions(cu=w["cu"])
    dx, du, Kx, d, pred, feas = noc.newton_step(w["fx"], w["fu"], w["ru"], w["Q"], w["R"], w["M"], cu_norm)
    dxl, dul, Kl, kl, dVl, cvx = serial_ld.seq_newton(*(N_(w[k]) for k in ("fx", "fu", "ru", "Q", "R", "M")),
                                                      float(cu_norm))
    assert relerr(N_(dx), dxl) < 1e-9 and relerr(N_(du), dul) < 1e-9
    assert relerr(-N_(Kx), Kl) < 1e-9 and relerr(N_(d), kl) < 1e-9     # reference sign convention: u = K x + k
    assert abs(float(pred) - dVl) <= 1e-9 * abs(dVl) and bool(feas[0]) == cvx
    # and the float64 serial recursion (what the reference's seq twin computes) sits just as close to it
    dxd, dud, _, _, dVd, _ = serial_ld.seq_newton(*(N_(w[k]) for k in ("fx", "fu", "ru", "Q", "R", "M")),
                                                  float(cu_norm), precision="f64")
    assert relerr(dxd, dxl) < 1e-9


@pytest.mark.parametrize("problem", ["pendulum", "cartpole"])
def test_config5_batched_subsample_histogram(problem):
    """BASELINE config 5 check: a 64-problem subsample at N = 1000 solved as ONE batch gives
    every member the iterate (<= 1e-9) and the Newton iteration count of solving it alone — equal iteration
    histograms."""
    from ipoc_b200 import noc, problems, batched
    N, B = 1000, 64
    rng = np.random.default_rng(1)
    if problem == "pendulum":
        ocp, x0 = problems.make_pendulum(1.0 / N), problems.pendulum_x0().numpy()
    else:
        ocp, x0 = problems.make_cartpole(1.0 / N), problems.cartpole_x0().numpy()
    x0s = x0[None] + 0.1 * rng.standard_normal((B, x0.shape[0]))
    u0s = 0.1 * rng.standard_normal((B, N, 1))
    ub, itb = batched.par_interior_point_optimal_control_batched(ocp, T(u0s), T(x0s))
    its1 = []
    for b in range(B):
        u1, it1 = noc.par_interior_point_optimal_control(ocp, T(u0s[b]), T(x0s[b]))
        its1.append(it1)
        assert int(itb[b]) == it1, (b, int(itb[b]), it1)
        assert relerr(N_(ub[b]), N_(u1)) < 1e-9
    assert np.array_equal(np.bincount(np.array(its1)), np.bincount(N_(itb).astype(np.int64)))


# ====================================================================== in-kernel levels (round 2)
# (1, 1, 0, 0): ten groups through the lane-cooperative level kernel; 2 = in-kernel levels everywhere; 3 = round-2 mix
HIER_SHAPES = [(0, 0, 0, 0), (1, 1, 2, 1), (1, 1, 5, 32), (1, 2, 32, 3), (1, 3, 7, 8), (1, 0, 3, 2),
               (1, 1, 0, 0), (4, 1, 0, 0), (4, 0, 0, 0), (2, 1, 0, 0), (3, 0, 0, 0)]


@pytest.mark.parametrize("enabled,leaf_chunk,group_warps,serial_top", HIER_SHAPES)
@pytest.mark.parametrize("nx,nu,N", [(2, 1, 1000), (4, 1, 10000), (4, 2, 777), (3, 1, 4097)])
def test_in_kernel_levels_all_shapes(enabled, leaf_chunk, group_warps, serial_top, nx, nu, N):
    """The levels above the warp scans completed by the last-arriving warps inside the leaf kernels (no top /
    mid kernels): every grouping (2 ... 32 warps per group; serial value chain or scan over the groups; more than
    32 groups folded q per lane) and the separate level kernels (enabled = 0) give the oracle's step, costates
    and forward pass."""
    from ipoc_b200 import noc, _lib
    from ipoc_b200.paroc import LQT, par_fwd_pass
    rng = np.random.default_rng(11 * nx + N)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    dxo, duo, Kxo, do, predo, feaso = oracle_newton(fx, fu, ru, Q, R, M, 0.05)
    _lib.lib().ipoc_set_tuning(leaf_chunk, 0, 0)
    _lib.lib().ipoc_set_hier(enabled, group_warps, serial_top)
    for rep in range(2):   # second call: the arrival counters must have been re-armed
        dx, du, Kx, d, pred, feas = noc.newton_step(T(fx), T(fu), T(ru), T(Q), T(R), T(M), T([0.05]))
        assert relerr(N_(dx), dxo) < 1e-10 and relerr(N_(du), duo) < 1e-10
        assert relerr(N_(Kx), Kxo) < 1e-10 and relerr(N_(d), do) < 1e-10
        assert abs(float(pred) - predo) <= 1e-10 * abs(predo) and bool(feas[0]) == bool(feaso)
    # K1 in both directions against a serial loop
    c = rng.standard_normal((N, nx))
    seed = rng.standard_normal(nx)
    lam = N_(noc.affine_scan(T(fx), T(c), T(seed), reverse=True, transpose=True))
    ref = np.zeros((N + 1, nx))
    ref[N] = seed
    for k in range(N - 1, -1, -1):
        ref[k] = fx[k].T @ ref[k + 1] + c[k]
    assert relerr(lam, ref) < 1e-11
    xs = N_(noc.affine_scan(T(fx), T(c), T(seed), reverse=False, transpose=False))
    ref = np.zeros((N + 1, nx))
    ref[0] = seed
    for k in range(N):
        ref[k + 1] = fx[k] @ ref[k] + c[k]
    assert relerr(xs, ref) < 1e-11
    # raw forward pass (its own up-sweep) with x0 != 0
    lq = noc_np.noc_to_lqt(ru, Q, R, M, fx, fu)
    x0 = rng.standard_normal(nx)
    uo, xo = paroc_np.par_fwd_pass(lq, x0, Kxo, do)
    u, x = par_fwd_pass(LQT(*(T(a) for a in lq)), T(x0), T(Kxo), T(do))
    assert relerr(N_(u), uo) < 1e-10 and relerr(N_(x), xo) < 1e-10


@pytest.mark.parametrize("nx", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("N,B,leaf_chunk", [(5000, 1, 1), (40000, 1, 0), (1500, 3, 1), (33, 2, 1)])
def test_cooperative_level_kernel_every_nx(nx, N, B, leaf_chunk):
    """Riccati levels by the lane-cooperative kernel (2, 4 or 8 lanes per combine, one CTA per group of 32 warp
    totals, last CTA scans the groups): single and several groups, ragged last group, batches, every nx."""
    from ipoc_b200 import noc, _lib
    rng = np.random.default_rng(100 * nx + N)
    nu = 1 + (nx % 2 if nx > 1 else 0)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu, batch=B)
    reg = 0.05 + 0.1 * rng.random(B)
    _lib.lib().ipoc_set_tuning(leaf_chunk, 0, 0)
    _lib.lib().ipoc_set_hier(4, 0, 0)   # the cooperative kernel for any number of groups (default: from 6)
    for rep in range(2):
        dx, du, Kx, d, pred, feas = noc.newton_step(T(fx), T(fu), T(ru), T(Q), T(R), T(M), T(reg))
    for b in range(B):
        dxo, duo, Kxo, do, predo, feaso = oracle_newton(fx[b], fu[b], ru[b], Q[b], R[b], M[b], reg[b])
        assert relerr(N_(dx[b]), dxo) < 1e-10 and relerr(N_(du[b]), duo) < 1e-10
        assert relerr(N_(Kx[b]), Kxo) < 1e-10 and relerr(N_(d[b]), do) < 1e-10
        assert abs(float(pred[b]) - predo) <= 1e-10 * abs(predo) and bool(feas[b]) == bool(feaso)


@pytest.mark.parametrize("enabled,leaf_chunk,group_warps,serial_top", HIER_SHAPES)
def test_in_kernel_levels_batched_and_sharded(enabled, leaf_chunk, group_warps, serial_top):
    from ipoc_b200 import noc, sharded, _lib
    rng = np.random.default_rng(5)
    _lib.lib().ipoc_set_tuning(leaf_chunk if leaf_chunk else 2, 0, 0)
    _lib.lib().ipoc_set_hier(enabled, group_warps, serial_top)
    N, nx, nu, B = 700, 4, 1, 5
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu, batch=B)
    reg = 0.1 + rng.random(B)
    dx, du, Kx, d, pred, feas = noc.newton_step(T(fx), T(fu), T(ru), T(Q), T(R), T(M), T(reg))
    for b in range(B):
        dxo, duo, Kxo, do, predo, feaso = oracle_newton(fx[b], fu[b], ru[b], Q[b], R[b], M[b], reg[b])
        assert relerr(N_(dx[b]), dxo) < 1e-10 and relerr(N_(du[b]), duo) < 1e-10
        assert abs(float(pred[b]) - predo) <= 1e-10 * abs(predo) and bool(feas[b]) == bool(feaso)
    args = [T(a[0]) for a in (fx, fu, ru, Q, R, M)]
    for P in (2, 3):
        dxs, dus, Kxs, ds, preds, feass = sharded.newton_step_virtual_ranks(*args, T(reg[:1]), P)
        assert relerr(N_(dxs), N_(dx[0])) < 1e-11 and relerr(N_(dus), N_(du[0])) < 1e-11
        assert abs(float(preds) - float(pred[0])) <= 1e-11 * abs(float(pred[0])) and feass == bool(feas[0])


@pytest.mark.parametrize("N,B", [(300, 1), (10000, 1), (1000, 7), (100000, 1), (64, 200)])
def test_fused_pass_equals_separate_calls(N, B):
    """ipoc_costates_f64 + ipoc_newton_attempt_f64 (5 launches: ||cu||, max|ru|, reg = rp*||cu||, constraint
    reduction and accept update as side jobs of the scan kernels) against the same pass as separate C-ABI calls
    (K1, K4, K2+K3, K4, A8).  Everything but the summation order of ||cu|| is the same arithmetic."""
    from ipoc_b200 import _lib
    from ipoc_b200.runner import NewtonPass
    rng = np.random.default_rng(N + B)
    nx, nu, nc = 4, 1, 2
    shape = (B, N) if B > 1 else (N,)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu, batch=B if B > 1 else None)
    cx, cu = rng.standard_normal(shape + (nx,)), rng.standard_normal(shape + (nu,))
    lamT = rng.standard_normal((B, nx))
    cons = -rng.random(shape + (nc,))
    if B > 1:
        cons[1, N // 2, 1] = 0.25          # one member infeasible
    mk = lambda: NewtonPass(T(fx), T(fu), T(cx), T(cu), T(lamT), T(ru), T(Q), T(R), T(M), T(cons), rp=0.7)
    a, b = mk(), mk()
    b.fused = False
    for p in (a, b):
        p.new_cost.copy_(T(rng.standard_normal(B)) * 0 + 0.5)
        p.cost.fill_(1.0)
    assert mk().launches_per_pass() <= 6 + (B > 1) or B >= 64     # (+ the Riccati top below two dozen groups; batch > 1: + the seed transposition of K1)
    for rep in range(3):      # rp / r_inc evolve from pass to pass on the device
        a.run()
        b.run()
        torch.cuda.synchronize()
        assert torch.equal(a.lam, b.lam) and torch.equal(a.hu, b.hu)
        assert float((a.cu_norm - b.cu_norm).abs().max()) <= 1e-14 * float(b.cu_norm.abs().max())
        assert relerr(N_(a.dx), N_(b.dx)) < 1e-12 and relerr(N_(a.du), N_(b.du)) < 1e-12
        assert relerr(N_(a.pred), N_(b.pred)) < 1e-12
        assert torch.equal(a.bwd_feas, b.bwd_feas) and torch.equal(a.traj_feas, b.traj_feas)
        assert torch.equal(a.success, b.success)
        assert relerr(N_(a.rp), N_(b.rp)) < 1e-12 and torch.equal(a.r_inc, b.r_inc)
    if B > 1:
        assert int(a.traj_feas[1]) == 0 and int(a.traj_feas[0]) == 1
    # a captured graph of the fused pass replays to the same numbers
    a2 = mk()
    a2.new_cost.fill_(0.5)
    a2.capture()
    a2.rp.fill_(0.7)          # the warm-up pass of the capture moved them
    a2.r_inc.fill_(2.0)
    for rep in range(3):
        a2.replay()
    torch.cuda.synchronize()
    assert torch.equal(a2.dx, a.dx) and torch.equal(a2.rp, a.rp) and torch.equal(a2.success, a.success)


def test_trial_point_side_job():
    """tx = x + dx, tu = u + du written by K3's leaf kernel (ipoc_newton_attempt_f64) == the stand-alone kernel."""
    from ipoc_b200 import _lib as L
    rng = np.random.default_rng(9)
    for N, B in ((257, 1), (5000, 1), (100, 40)):
        nx, nu = 4, 1
        fx, fu, ru, Q, R, M = (T(a) for a in random_lq(rng, N, nx, nu, batch=B))
        x, u = T(rng.standard_normal((B, N + 1, nx))), T(rng.standard_normal((B, N, nu)))
        o = dict(dtype=torch.float64, device=DEV)
        dx, du, tx, tu = torch.empty(B, N + 1, nx, **o), torch.empty(B, N, nu, **o), torch.empty(B, N + 1, nx, **o), torch.empty(B, N, nu, **o)
        Kx, d = torch.empty(B, N, nu, nx, **o), torch.empty(B, N, nu, **o)
        pred, hu = torch.empty(B, **o), torch.empty(B, **o)
        feas = torch.empty(B, dtype=torch.int32, device=DEV)
        rp, cn = T(0.5 + rng.random(B)), T(0.5 + rng.random(B))
        nbytes = L.lib().ipoc_workspace_bytes(L.WS_NEWTON_ATTEMPT, N, nx, nu, B)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
        L.check(L.lib().ipoc_workspace_init(L.ptr(ws), nbytes, L.stream_ptr()))
        p = L.ptr
        L.check(L.lib().ipoc_newton_attempt_f64(N, nx, nu, 1, B, p(fx), p(fu), p(ru), p(Q), p(R), p(M), p(rp), p(cn),
                                                p(dx), p(du), p(Kx), p(d), p(pred), p(feas), p(hu), p(x), p(u), p(tx),
                                                p(tu), *([None] * 10), p(ws), nbytes, L.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(tx, x + dx) and torch.equal(tu, u + du)
        assert torch.equal(hu, ru.abs().amax(dim=(1, 2)))
        from ipoc_b200 import noc
        dx2 = noc.newton_step(fx, fu, ru, Q, R, M, rp * cn)[0]
        assert torch.equal(dx2, dx)


def test_config3_constrained_receding_horizon_mpc():
    """BASELINE config 3, box-constrained extension (no reference script; anchors in mpc.constrained_mpc): the
    receding-horizon loop of repeated IP solves against its oracle twin — closed-loop states and controls within
    1e-9, the same Newton iteration count at every MPC step, the bound active at the start and never violated."""
    from ipoc_b200 import problems
    from ipoc_b200.mpc import constrained_mpc
    from oracle.autodiff import Evaluator
    ub = 15.0                                 # the unconstrained solution starts at u = -48
    ocp = problems.make_linear_demo(0.1, control_bound=ub)
    x0 = np.array([2.0, 1.0])
    xs_o, us_o, its_o = noc_np.constrained_mpc(Evaluator(ocp), x0, horizon=40, sim_steps=4)
    xs, us, its = constrained_mpc(ocp, T(x0), horizon=40, sim_steps=4)
    assert its == its_o and min(its) > 5
    assert relerr(N_(xs), xs_o) < 1e-9 and relerr(N_(us), us_o) < 1e-9
    assert float(us.abs().max()) < ub and float(us[0].abs()) > ub - 1e-3       # the box is active, never crossed


@pytest.mark.parametrize("P", [2, 3, 8])
@pytest.mark.parametrize("nx,nu,N", [(4, 1, 1003), (2, 1, 64), (4, 1, 40000)])
def test_time_sharded_whole_pass_virtual_ranks(P, nx, nu, N):
    """The WHOLE hot-path pass time-sharded (sharded.SegmentPass: K1 with its carry exchange, the K4 scalars
    folded over the ranks in fixed order, reg = rp*||cu||, K2, K3 — three exchanges) with P virtual ranks on one
    GPU equals the single-device pass."""
    from ipoc_b200 import sharded
    from ipoc_b200.runner import NewtonPass
    rng = np.random.default_rng(P * 7 + N)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    cx, cu = rng.standard_normal((N, nx)), rng.standard_normal((N, nu))
    lamT = rng.standard_normal(nx)
    cons = -rng.random((N, 2))
    cons[N // 3, 0] = 0.5
    ref = NewtonPass(T(fx), T(fu), T(cx), T(cu), T(lamT), T(ru), T(Q), T(R), T(M), T(cons), rp=0.8)
    ref.run()
    out = sharded.pass_virtual_ranks(*(T(a) for a in (fx, fu, cx, cu, lamT, ru, Q, R, M)), P, cons=T(cons), rp=0.8)
    torch.cuda.synchronize()
    assert relerr(N_(out["lam"]), N_(ref.lam[0])) < 1e-11
    assert relerr(N_(out["dx"]), N_(ref.dx[0])) < 1e-10 and relerr(N_(out["du"]), N_(ref.du[0])) < 1e-10
    assert relerr(N_(out["Kx"]), N_(ref.Kx[0])) < 1e-10 and relerr(N_(out["d"]), N_(ref.d[0])) < 1e-10
    assert abs(float(out["pred"]) - float(ref.pred)) <= 1e-10 * abs(float(ref.pred))
    assert float(out["hu"]) == float(ref.hu) and abs(float(out["cu_norm"]) - float(ref.cu_norm)) <= 1e-14 * float(ref.cu_norm)
    assert out["bwd_feasible"] == bool(ref.bwd_feas[0]) and out["traj_feasible"] is False


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL)")
def test_time_sharded_pass_nccl_two_gpus():
    """Real NCCL path (runs when the box has >= 2 GPUs): tests/dist_time_sharded.py under torchrun — every rank
    compares its slice of the sharded step AND of the whole sharded pass (three all-gathers inside one CUDA graph)
    with the single-device result."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "dist_time_sharded.py"), "100003", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(" OK ") == 2


# ====================================================================== parallel-in-time rollout of the built-in plants
@pytest.mark.parametrize("name,N,amp", [("cartpole", 10000, 0.1), ("cartpole", 4096, 20.0), ("pendulum", 5000, 0.1),
                                        ("pendulum", 3000, 4.0), ("cartpole", 100000, 0.1)])
def test_plant_parallel_rollout_equals_serial_kernel(name, N, amp):
    """Newton on the rollout equations (ipoc_plant_rollout_lin_f64 + forward affine scan per iteration) against the
    serial one-thread kernel (ref noc/utils.py:57-63): same trajectory to rounding, from a cold start and from a
    perturbed guess (what a barrier stage hands to the next)."""
    from ipoc_b200 import plants, problems
    ocp = (problems.make_cartpole if name == "cartpole" else problems.make_pendulum)(1.0 / N)
    x0 = (problems.cartpole_x0 if name == "cartpole" else problems.pendulum_x0)(device="cuda")
    plant = plants.plant_of(ocp)
    u = T(amp * np.random.default_rng(3).standard_normal((N, 1)))
    xs = plants.rollout(plant, u, x0)
    scale = 1.0 + float(xs.abs().max())
    xp, it = plants.rollout_parallel(plant, u, x0)
    assert 0 < it <= 40, it
    assert float((xp - xs).abs().max()) <= 1e-10 * scale
    guess = xs + 1e-3 * torch.randn_like(xs)
    xw, itw = plants.rollout_parallel(plant, u, x0, x_guess=guess)
    assert 0 < itw <= it and float((xw - xs).abs().max()) <= 1e-10 * scale
    # batched, members with different controls
    ub = torch.stack((u, 0.5 * u, -u))
    xb, itb = plants.rollout_parallel(plant, ub, x0.expand(3, -1).contiguous())
    xsb = plants.rollout(plant, ub, x0.expand(3, -1).contiguous())
    assert itb > 0 and float((xb - xsb).abs().max()) <= 1e-10 * (1.0 + float(xsb.abs().max()))


def test_plant_parallel_rollout_falls_back_on_divergence():
    from ipoc_b200 import plants, problems
    N = 4096
    plant = plants.plant_of(problems.make_cartpole(1.0 / N))
    u = T(np.full((N, 1), 1e300))
    x, it = plants.rollout_parallel(plant, u, problems.cartpole_x0(device="cuda"))
    assert it == -1 and tuple(x.shape) == (N + 1, 4)       # the serial kernel's (non-finite) result, no exception


@pytest.mark.parametrize("name,N,B", [("cartpole", 100000, 1), ("cartpole", 20000, 3), ("pendulum", 1000000, 1),
                                      ("cartpole", 9000, 2), ("cartpole", 3000, 1), ("pendulum", 700, 40)])
def test_plant_cost_forms_agree(name, N, B):
    """total_cost / feasibility of the built-in plants: single CTA, thread-block cluster (distributed shared memory)
    and grid form (caller scratch, last arriver folds) against the host framework evaluating the same total_cost, incl. an
    infeasible member, and twice in a row (the arrival counters must be back at zero)."""
    from ipoc_b200 import plants, problems, _lib
    Ts = 1.0 / N
    ocp = (problems.make_cartpole if name == "cartpole" else problems.make_pendulum)(Ts)
    plant = plants.plant_of(ocp)
    rng = np.random.default_rng(N + B)
    nx = 4 if name == "cartpole" else 2
    x = rng.standard_normal((B, N + 1, nx))
    u = 0.3 * plant["bound"] * rng.uniform(-1, 1, (B, N, 1))
    if B > 1:
        u[1, N // 2, 0] = 1.5 * plant["bound"]            # infeasible member: NaN cost, feasible = 0
    bp = 0.02
    xt, ut = T(x), T(u)
    want = np.array([float(ocp.total_cost(xt[b], ut[b], bp)) for b in range(B)])
    nbytes = int(_lib.lib().ipoc_plant_cost_workspace_bytes(N, B))
    assert (nbytes > 0) == (B < 32 and N > 8192)
    forms = [None] if nbytes == 0 else [None, (None, 0)]      # default (scratch if the shape uses it) / cluster form
    for scratch in forms:
        for rep in range(2):
            tot, feas = plants.cost(plant, xt, ut, bp, scratch=scratch)
            tot, feas = N_(tot), N_(feas)
            for b in range(B):
                bad = B > 1 and b == 1
                assert bool(feas[b]) == (not bad)
                if bad:
                    assert np.isnan(tot[b]) and np.isnan(want[b])
                else:
                    assert abs(tot[b] - want[b]) <= 1e-11 * abs(want[b]), (scratch, rep, b, tot[b], want[b])


@pytest.mark.parametrize("name,N", [("cartpole", 3000), ("pendulum", 6000)])
def test_mid_horizon_solve_device_loop_equals_host_steered(name, N):
    """Horizons between the fixture sizes: the device-resident loop here launches the plant cost kernel as a
    thread-block cluster INSIDE a captured graph (1024 < N <= 8192) and starts every stage with the parallel plant
    rollout (N >= 2048); the host-steered graphs of the same solver (one host read per attempt, plant kernels,
    shared cost scratch) must give the same iterate and Newton iteration count."""
    from ipoc_b200 import noc, problems
    ocp = (problems.make_cartpole if name == "cartpole" else problems.make_pendulum)(1.0 / N)
    x0 = (problems.cartpole_x0 if name == "cartpole" else problems.pendulum_x0)(device="cuda")
    u0 = T(0.1 * np.random.default_rng(1).standard_normal((N, 1)))
    ud, itd = noc.par_interior_point_optimal_control(ocp, u0, x0)
    uh, ith = noc.par_interior_point_optimal_control(ocp, u0, x0, use_graphs="host")
    assert itd == ith and itd > 20
    assert relerr(N_(ud), N_(uh)) < 1e-9

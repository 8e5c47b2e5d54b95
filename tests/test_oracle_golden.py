"""CPU: the NumPy oracle against fixtures produced by executing the reference's own source
(tests/golden/gen_golden.py).  `ref_*` keys = reference code alone; `refp_*` = reference code +
restated paroc."""
import numpy as np
import pytest

from oracle import noc_np, paroc_np
from oracle.assoc_scan import associative_scan, serial_scan
from helpers import STEP_FIXTURES, derivs_from_golden, relerr, random_lq


@pytest.mark.parametrize("name", STEP_FIXTURES)
def test_costates_match_reference(golden, name):
    g = golden(name)
    d = derivs_from_golden(g)
    lam_par = noc_np.par_costates(g["ref_lamT"], d)
    lam_seq = noc_np.seq_costates(g["ref_lamT"], d)
    assert relerr(lam_par, g["ref_costates_par"]) < 1e-13
    assert relerr(lam_seq, g["ref_costates_seq"]) < 1e-13
    assert relerr(lam_par, g["ref_costates_seq"]) < 1e-12


@pytest.mark.parametrize("name", STEP_FIXTURES)
def test_lqr_params_and_lqt_references(golden, name):
    g = golden(name)
    d = derivs_from_golden(g)
    ru, Q, R, M = noc_np.compute_lqr_params(g["ref_costates_par"], d)
    for mine, key in ((ru, "ref_ru"), (Q, "ref_Q"), (R, "ref_R"), (M, "ref_M")):
        assert relerr(mine, g[key]) < 1e-14
    nu = R.shape[1]
    lqt = noc_np.noc_to_lqt(ru, Q, R + float(g["ref_reg"]) * np.eye(nu)[None], M, d.fx, d.fu)
    assert relerr(lqt.r, g["ref_lqt_r"]) < 1e-12
    assert relerr(lqt.s, g["ref_lqt_s"]) < 1e-12
    # identities of noc_to_lqt: -X r - M s = 0, -U s - M' r = ru  (ref :62-66)
    res_x = np.einsum("tij,tj->ti", lqt.X, lqt.r) + np.einsum("tij,tj->ti", lqt.M, lqt.s)
    res_u = np.einsum("tij,tj->ti", lqt.U, lqt.s) + np.einsum("tji,tj->ti", lqt.M, lqt.r) + ru
    assert np.max(np.abs(res_x)) < 1e-10 and np.max(np.abs(res_u)) < 1e-10


@pytest.mark.parametrize("name", STEP_FIXTURES)
def test_newton_step_matches_reference_sequential_twin(golden, name):
    """The restated `paroc` scans vs the reference's in-tree serial Riccati (pure reference code)."""
    g = golden(name)
    d = derivs_from_golden(g)
    nx = d.fx.shape[1]
    dx, du, pred, feas, _ = noc_np.par_Newton(nx, d, float(g["reg_param"]), g["ref_ru"], g["ref_Q"], g["ref_R"],
                                              g["ref_M"])
    assert relerr(dx, g["ref_seq_dx"]) < 1e-10
    assert relerr(du, g["ref_seq_du"]) < 1e-10
    assert abs(pred - float(g["ref_seq_dV"])) <= 1e-11 * abs(float(g["ref_seq_dV"]))
    assert bool(feas) == bool(g["ref_seq_convex"])
    # and against the reference driver + restated paroc run through the shim
    assert relerr(dx, g["refp_dx"]) < 1e-12 and relerr(du, g["refp_du"]) < 1e-12
    # oracle's own serial restatement of ref seq bwd/fwd
    reg = float(g["ref_reg"])
    K, k, dV, convex = noc_np.seq_bwd_pass(g["ref_Q"][0], g["ref_ru"], g["ref_Q"], g["ref_R"], g["ref_M"], d.fx, d.fu,
                                           reg)
    assert relerr(K, g["ref_seq_K"]) < 1e-11 and relerr(k, g["ref_seq_k"]) < 1e-11
    du2, dx2 = noc_np.seq_fwd_pass(K, k, d.fx, d.fu)
    assert relerr(dx2, g["ref_seq_dx"]) < 1e-11 and relerr(du2, g["ref_seq_du"]) < 1e-11


@pytest.mark.parametrize("nx,nu,N", [(2, 1, 37), (4, 1, 64), (3, 2, 50), (8, 2, 33)])
def test_par_vs_seq_lqt_random(nx, nu, N):
    rng = np.random.default_rng(nx * 100 + nu * 10 + N)
    fx, fu, ru, Q, R, M = random_lq(rng, N, nx, nu)
    lqt = noc_np.noc_to_lqt(ru, Q, R, M, fx, fu)
    lqt = lqt._replace(c=0.01 * rng.standard_normal((N, nx)), rT=rng.standard_normal(nx))
    Kx, d, S, v, pred, feas = paroc_np.par_bwd_pass(lqt)
    Kx2, d2, S2, v2 = paroc_np.seq_bwd_pass(lqt)
    assert relerr(Kx, Kx2) < 1e-10 and relerr(d, d2) < 1e-10 and relerr(S, S2) < 1e-10 and relerr(v, v2) < 1e-10
    x0 = rng.standard_normal(nx)
    u, x = paroc_np.par_fwd_pass(lqt, x0, Kx, d)
    u2, x2 = paroc_np.seq_fwd_pass(lqt, x0, Kx, d)
    assert relerr(u, u2) < 1e-10 and relerr(x, x2) < 1e-10
    assert feas and pred < 0


def test_combine_is_associative_and_scan_orders_agree():
    rng = np.random.default_rng(5)
    fx, fu, ru, Q, R, M = random_lq(rng, 9, 3, 1)
    lqt = noc_np.noc_to_lqt(ru, Q, R, M, fx, fu)
    el = paroc_np.bwd_elements(lqt)
    pick = lambda i: tuple(e[i:i + 1] for e in el)
    a, b, c = pick(2), pick(3), pick(4)
    left = paroc_np.combine(paroc_np.combine(a, b), c)
    right = paroc_np.combine(a, paroc_np.combine(b, c))
    for l, r in zip(left, right):
        assert relerr(l, r) < 1e-12
    tree = associative_scan(paroc_np._combine_rev, el, reverse=True)
    fold = serial_scan(paroc_np._combine_rev, el, reverse=True)
    for t, f in zip(tree, fold):
        assert relerr(t, f) < 1e-11
    # terminal element (A=0,b=0,C=0) annihilates everything after it: suffix value function of the
    # last real step does not depend on what is appended on the right of the terminal element
    A, b_, C, eta, J = el
    assert np.all(A[-1] == 0) and np.all(C[-1] == 0) and np.all(b_[-1] == 0)


@pytest.mark.parametrize("name,tol_u", [("solve_pendulum_N20", 1e-9), ("solve_linear_N40", 1e-9),
                                        ("solve_cartpole_N40", 1e-9), ("solve_pendulum_N100", 1e-9)])
def test_full_solve_matches_reference_driver(golden, name, tol_u):
    """Driver loops (barrier / Newton / accept-reject + regularisation) vs the reference's own
    `par_interior_point_optimal_control` executed on the shim: same iterate, same iteration count."""
    import ipoc_b200.problems as P
    from oracle.autodiff import Evaluator
    g = golden(name)
    N = g["u0"].shape[0]
    if "pendulum" in name:
        ocp = P.make_pendulum(1.0 / N)
    elif "cartpole" in name:
        ocp = P.make_cartpole(1.0 / N)
    else:
        ocp = P.make_linear_demo(0.1)
    u, its = noc_np.par_interior_point_optimal_control(Evaluator(ocp), g["u0"], g["x0"])
    assert its == int(g["refp_iterations"])
    assert relerr(u, g["refp_opt_u"]) < tol_u


# ---------------------------------------------------------------------- BASELINE-size fixtures (round 2)
@pytest.mark.parametrize("name", ["step_cartpole_N10000", "step_cartpole_N10000_warm"])
def test_config2_oracle_vs_reference(golden, name):
    """Cartpole N = 1e4 (BASELINE config 2): oracle derivatives -> costates -> LQ parameters -> par_Newton against
    what the reference's own source produced; the C serial oracles (long double / double) against the
    reference's in-tree sequential step."""
    import ipoc_b200.problems as P
    from oracle.autodiff import Evaluator
    from oracle import serial_ld
    g = golden(name)
    N = g["controls"].shape[0]
    ev = Evaluator(P.make_cartpole(1.0 / N))
    d = ev.derivatives(g["states"], g["controls"], float(g["bp"]))
    lam = noc_np.par_costates(ev.final_cost_grad(g["states"][-1]), d)
    assert relerr(lam, g["ref_costates_par"]) < 1e-12
    ru, Q, R, M = noc_np.compute_lqr_params(lam, d)
    for mine, key in ((ru, "ref_ru"), (Q, "ref_Q"), (R, "ref_R"), (M, "ref_M")):
        assert relerr(mine, g[key]) < 1e-12
    dx, du, pred, feas, _ = noc_np.par_Newton(4, d, float(g["reg_param"]), g["ref_ru"], g["ref_Q"], g["ref_R"],
                                              g["ref_M"])
    assert relerr(dx, g["refp_dx"]) < 1e-11 and relerr(du, g["refp_du"]) < 1e-11
    assert relerr(dx, g["ref_seq_dx"]) < 1e-9 and relerr(du, g["ref_seq_du"]) < 1e-9
    assert abs(pred - float(g["ref_seq_dV"])) <= 1e-10 * abs(float(g["ref_seq_dV"]))
    assert bool(feas) == bool(g["ref_seq_convex"])
    for prec, tol in (("ld", 1e-10), ("f64", 1e-10)):
        dxs, dus, _, _, dV, cvx = serial_ld.seq_newton(d.fx, d.fu, g["ref_ru"], g["ref_Q"], g["ref_R"], g["ref_M"],
                                                       float(g["ref_reg"]), precision=prec)
        assert relerr(dxs, g["ref_seq_dx"]) < tol and relerr(dus, g["ref_seq_du"]) < tol
        assert abs(dV - float(g["ref_seq_dV"])) <= tol * abs(dV) and cvx == bool(g["ref_seq_convex"])


@pytest.mark.parametrize("name", STEP_FIXTURES)
def test_serial_c_oracles_match_reference_sequential_step(golden, name):
    """oracle/serial_ld.c (both precisions) is pinned by `ref_seq_*`: the reference's in-tree sequential Newton
    step and sequential costates executed from their source."""
    from oracle import serial_ld
    g = golden(name)
    d = derivs_from_golden(g)
    for prec in ("ld", "f64"):
        dx, du, K, k, dV, cvx = serial_ld.seq_newton(d.fx, d.fu, g["ref_ru"], g["ref_Q"], g["ref_R"], g["ref_M"],
                                                     float(g["ref_reg"]), precision=prec)
        assert relerr(dx, g["ref_seq_dx"]) < 1e-13 and relerr(du, g["ref_seq_du"]) < 1e-13
        assert relerr(K, g["ref_seq_K"]) < 1e-13 and relerr(k, g["ref_seq_k"]) < 1e-13
        assert abs(dV - float(g["ref_seq_dV"])) <= 1e-13 * abs(dV) and cvx == bool(g["ref_seq_convex"])
        lam = serial_ld.seq_costates(d.fx, d.cx, g["ref_lamT"], precision=prec)
        assert relerr(lam, g["ref_costates_seq"]) < 1e-13

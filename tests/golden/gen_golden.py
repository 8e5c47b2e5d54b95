#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

Run in the build container only (needs /root/reference; the GPU box has neither it nor
any use for this script):

    python tests/golden/gen_golden.py

How: JAX is not installed, so the reference's unmodified files
(/root/reference/noc/*.py, and the function definitions of
/root/reference/examples/{pendulum,cartpole}_runtime.py, linear_demo_cuda.py) are imported on
top of `oracle/jaxshim` (jax -> float64 torch-CPU/torch.func).  `paroc`, which the reference
imports but which is absent from this machine, is served by `oracle/jaxshim/paroc`
(= oracle/paroc_np.py).  Consequently:
  * everything whose key starts with `ref_` was computed by reference source code alone;
  * keys starting with `refp_` were computed by reference source code calling the restated
    `paroc` (so they pin the driver logic, not the scan arithmetic);
  * the scan arithmetic itself is pinned by `ref_seq_*`: the reference's in-tree sequential
    Newton step (noc/seq_interior_point_newton.py:42-90) on the same LQ data.
Random inputs come from numpy.random.default_rng (JAX's threefry PRNGKey(1) stream cannot be
reproduced without JAX).
"""
import ast
import os
import sys
import warnings
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "jaxshim"), ROOT, REF]
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
from noc.optimal_control_problem import OCP, LinearizedOCP  # noqa: E402  (reference)
import noc.par_interior_point_newton as refpar  # noqa: E402  (reference)
import noc.seq_interior_point_newton as refseq  # noqa: E402  (reference)
import noc.costates as refcos  # noqa: E402  (reference)
from noc.utils import wrap_angle, euler, rollout, discretize_dynamics  # noqa: E402  (reference)


def example_functions(script):
    """exec only the top-level `def`s of a reference example script (they are scripts with a
    timing loop at module level, so they cannot be imported)."""
    src = open(os.path.join(REF, "examples", script)).read()
    tree = ast.parse(src)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    ns = {"jnp": jnp, "jax": jax, "wrap_angle": wrap_angle}
    exec(compile(tree, script, "exec"), ns)
    return ns


def A(x):
    if isinstance(x, torch.Tensor):
        return x.detach().numpy().copy()
    return np.asarray(x)


def newton_step_fixture(name, ocp, x0, u, bp, reg_param, warm_iters=0):
    """All intermediate quantities of ONE Newton step at a given iterate."""
    x0 = jnp.array(x0)
    u = jnp.array(u)
    if warm_iters:
        # move to an interior iterate by a few accepted reference Newton iterations
        states = rollout(ocp.dynamics, u, x0)
        for _ in range(warm_iters):
            d = refpar.compute_derivatives(ocp, states, u, bp)
            lam = refcos.par_costates(ocp, states[-1], d)
            ru, Q, R, M = refpar.compute_lqr_params(lam, d)
            rp = 1.0
            while True:
                dx, du, pred, feas, _ = refpar.par_Newton(states, d, rp, ru, Q, R, M)
                tx, tu = states + dx, u + du
                ok = bool(refpar.check_traj_feasibility(ocp, tx, tu))
                if ok and bool(feas) and float(ocp.total_cost(tx, tu, bp)) < float(ocp.total_cost(states, u, bp)):
                    states, u = tx, tu
                    break
                rp *= 4.0
    else:
        states = rollout(ocp.dynamics, u, x0)
    d = refpar.compute_derivatives(ocp, states, u, bp)
    lamT = jax.grad(ocp.final_cost, 0)(states[-1])
    lam_par = refcos.par_costates(ocp, states[-1], d)
    lam_seq = refcos.seq_costates(ocp, states[-1], d)
    ru, Q, R, M = refpar.compute_lqr_params(lam_par, d)
    reg = reg_param * jnp.linalg.norm(d.cu)
    Rreg = R + jnp.kron(jnp.ones((R.shape[0], 1, 1)), reg * jnp.eye(R.shape[1]))
    lqt = refpar.noc_to_lqt(ru, Q, Rreg, M, d.fx, d.fu)
    # reference's in-tree sequential Newton step on the same LQ data, with the two documented
    # differences neutralised: terminal Hessian := Q[0] (par: XT = Q[0]) and rp := reg.
    Q0 = Q[0]
    fc = lambda xx: 0.5 * xx @ Q0 @ xx
    K, k, dV, convex = refseq.bwd_pass(fc, states[-1], LinearizedOCP(ru, Q, R, M), d, reg)
    du_seq, dx_seq = refseq.fwd_pass(K, k, d)
    # reference par_Newton (reference code + restated paroc)
    dx, du, pred, feas, _ = refpar.par_Newton(states, d, reg_param, ru, Q, R, M)
    tx, tu = states + dx, u + du
    cons = jax.vmap(ocp.constraints)(tx[:-1], tu)
    out = dict(
        bp=bp, reg_param=reg_param, states=A(states), controls=A(u), x0=A(x0),
        **{"d_" + f: A(getattr(d, f)) for f in d._fields},
        ref_lamT=A(lamT), ref_costates_par=A(lam_par), ref_costates_seq=A(lam_seq),
        ref_ru=A(ru), ref_Q=A(Q), ref_R=A(R), ref_M=A(M),
        ref_lqt_r=A(lqt.r), ref_lqt_s=A(lqt.s), ref_reg=A(reg),
        ref_seq_K=A(K), ref_seq_k=A(k), ref_seq_dV=A(dV), ref_seq_convex=A(convex),
        ref_seq_du=A(du_seq), ref_seq_dx=A(dx_seq),
        refp_dx=A(dx), refp_du=A(du), refp_pred=A(pred), refp_feasible=A(feas),
        ref_cost=A(ocp.total_cost(states, u, bp)),
        ref_new_cons=A(cons).reshape(cons.shape[0], -1),
        ref_new_feasible=A(refpar.check_traj_feasibility(ocp, tx, tu)),
        ref_new_cost=A(ocp.total_cost(tx, tu, bp)),
    )
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    err = max(float(abs(dx - dx_seq).max()), float(abs(du - du_seq).max()))
    print(f"{name}: N={u.shape[0]} |par-seq| step diff {err:.3e} pred {float(pred):.6e} dV {float(dV):.6e}")


def big_step_fixture(name, ocp, x0, u, bp, reg_param, warm_iters=0):
    """Compact fixture of ONE Newton step at a BASELINE-size horizon (config 2: cartpole N = 1e4).  The 130
    doubles per step of `Derivatives` are NOT stored (10 MB): a test re-derives them from (states, controls)
    and is then held against the reference's costates, LQ parameters (ru, Q diagonal-free summary: full Q, R, M
    are 21 doubles per step and are kept), the in-tree sequential step and the reference par_Newton outputs."""
    x0 = jnp.array(x0)
    u = jnp.array(u)
    states = rollout(ocp.dynamics, u, x0)
    for _ in range(warm_iters):   # accepted reference Newton iterations -> an interior iterate, indefinite Q
        d = refpar.compute_derivatives(ocp, states, u, bp)
        lam = refcos.par_costates(ocp, states[-1], d)
        ru, Q, R, M = refpar.compute_lqr_params(lam, d)
        rp = 1.0
        while True:
            dx, du, pred, feas, _ = refpar.par_Newton(states, d, rp, ru, Q, R, M)
            tx, tu = states + dx, u + du
            ok = bool(refpar.check_traj_feasibility(ocp, tx, tu))
            if ok and bool(feas) and float(ocp.total_cost(tx, tu, bp)) < float(ocp.total_cost(states, u, bp)):
                states, u = tx, tu
                break
            rp *= 4.0
    d = refpar.compute_derivatives(ocp, states, u, bp)
    lam_par = refcos.par_costates(ocp, states[-1], d)
    ru, Q, R, M = refpar.compute_lqr_params(lam_par, d)
    reg = reg_param * jnp.linalg.norm(d.cu)
    Q0 = Q[0]
    fc = lambda xx: 0.5 * xx @ Q0 @ xx
    t = time.time()
    K, k, dV, convex = refseq.bwd_pass(fc, states[-1], LinearizedOCP(ru, Q, R, M), d, reg)
    du_seq, dx_seq = refseq.fwd_pass(K, k, d)
    t1 = time.time() - t
    dx, du, pred, feas, _ = refpar.par_Newton(states, d, reg_param, ru, Q, R, M)
    tx, tu = states + dx, u + du
    eig = np.linalg.eigvalsh(0.5 * (A(Q) + A(Q).transpose(0, 2, 1)))
    out = dict(
        bp=bp, reg_param=reg_param, states=A(states), controls=A(u), x0=A(x0),
        ref_costates_par=A(lam_par), ref_ru=A(ru), ref_Q=A(Q), ref_R=A(R), ref_M=A(M), ref_reg=A(reg),
        ref_cu_norm=A(jnp.linalg.norm(d.cu)),
        ref_seq_du=A(du_seq), ref_seq_dx=A(dx_seq), ref_seq_dV=A(dV), ref_seq_convex=A(convex),
        refp_dx=A(dx), refp_du=A(du), refp_pred=A(pred), refp_feasible=A(feas),
        ref_cost=A(ocp.total_cost(states, u, bp)),
        ref_new_feasible=A(refpar.check_traj_feasibility(ocp, tx, tu)),
        ref_new_cost=A(ocp.total_cost(tx, tu, bp)),
        ref_Q_min_eig=float(eig.min()), ref_Q_indefinite_steps=int((eig.min(axis=1) < 0).sum()),
    )
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    err = max(float(abs(dx - dx_seq).max()), float(abs(du - du_seq).max()))
    print(f"{name}: N={u.shape[0]} |par-seq| step diff {err:.3e} pred {float(pred):.6e} dV {float(dV):.6e} "
          f"min eig(Q) {eig.min():.3e} ({out['ref_Q_indefinite_steps']} indefinite steps) seq step {t1:.1f}s", flush=True)


def solve_fixture(name, ocp, x0, u0, with_seq=True):
    x0 = jnp.array(x0)
    u0 = jnp.array(u0)
    t = time.time()
    u_par, it_par = refpar.par_interior_point_optimal_control(ocp, u0, x0)
    t1 = time.time() - t
    t = time.time()
    if with_seq:
        u_seq, it_seq = refseq.seq_interior_point_optimal_control(ocp, u0, x0)
    else:   # the sequential twin's Python-level scan is too slow through the shim at this size
        u_seq, it_seq = jnp.array(np.full(A(u_par).shape, np.nan)), -1
    t2 = time.time() - t
    xs = rollout(ocp.dynamics, u_par, x0)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), x0=A(x0), u0=A(u0),
        refp_opt_u=A(u_par), refp_iterations=int(it_par),
        ref_seq_opt_u=A(u_seq), ref_seq_iterations=int(it_seq),
        refp_final_cost0=A(ocp.total_cost(xs, u_par, 0.0)) if name.startswith("solve_linear") else np.nan,
    )
    print(f"{name}: par its {int(it_par)} ({t1:.1f}s)  seq its {int(it_seq)} ({t2:.1f}s)  "
          f"|u_par-u_seq| {float(abs(u_par - u_seq).max()):.3e}", flush=True)


def main():
    rng = lambda seed: np.random.default_rng(seed)
    pen = example_functions("pendulum_runtime.py")
    car = example_functions("cartpole_runtime.py")
    lin = example_functions("linear_demo_cuda.py")

    def pendulum_ocp(N):
        return OCP(euler(pen["pendulum"], 1.0 / N), pen["constraints"], pen["transient_cost"],
                   pen["final_cost"], pen["total_cost"])

    def cartpole_ocp(N):
        return OCP(euler(car["cartpole"], 1.0 / N), car["constraints"], car["transient_cost"],
                   car["final_cost"], car["total_cost"])

    pen_x0 = A(jnp.array([wrap_angle(0.1), -0.1]))
    car_x0 = A(jnp.array([0.01, wrap_angle(-0.01), 0.01, -0.01]))

    which = sys.argv[1:] or ["steps", "solves", "config1"]
    if "steps" in which:
        newton_step_fixture("step_pendulum_N64", pendulum_ocp(64), pen_x0,
                            0.1 * rng(1).standard_normal((64, 1)), 0.1, 1.0)
        newton_step_fixture("step_pendulum_N33_warm", pendulum_ocp(33), pen_x0,
                            0.1 * rng(2).standard_normal((33, 1)), 0.02, 0.37, warm_iters=3)
        newton_step_fixture("step_cartpole_N100", cartpole_ocp(100), car_x0,
                            0.1 * rng(1).standard_normal((100, 1)), 0.1, 1.0)
        newton_step_fixture("step_cartpole_N257_warm", cartpole_ocp(257), car_x0,
                            0.1 * rng(3).standard_normal((257, 1)), 0.004, 2.5, warm_iters=2)
    if "solves" in which:
        solve_fixture("solve_pendulum_N20", pendulum_ocp(20), pen_x0, 0.1 * rng(1).standard_normal((20, 1)))
        solve_fixture("solve_pendulum_N100", pendulum_ocp(100), pen_x0, 0.1 * rng(1).standard_normal((100, 1)))
        solve_fixture("solve_cartpole_N40", cartpole_ocp(40), car_x0, 0.1 * rng(1).standard_normal((40, 1)))
        lin_dyn = discretize_dynamics(lin["ode"], 0.1, 1)
        lin_ocp = OCP(lin_dyn, lin["constraints"], lin["stage_cost"], lin["final_cost"], lin["total_cost"])
        solve_fixture("solve_linear_N40", lin_ocp, np.array([2.0, 1.0]), np.zeros((40, 1)))
    if "config1" in which:
        # BASELINE.json config 1: pendulum N=500, Ts=1/500
        solve_fixture("solve_pendulum_N500", pendulum_ocp(500), pen_x0, 0.1 * rng(1).standard_normal((500, 1)))


    if "config2" in which:
        # BASELINE.json config 2 (cartpole N = 1e4, Ts = 1e-4): the first Newton step (bp 0.1, rp 1) and a warm
        # interior iterate; full solves at N = 1000 (config 5's horizon) and N = 1e4
        u1e4 = 0.1 * rng(1).standard_normal((10000, 1))
        big_step_fixture("step_cartpole_N10000", cartpole_ocp(10000), car_x0, u1e4, 0.1, 1.0)
        big_step_fixture("step_cartpole_N10000_warm", cartpole_ocp(10000), car_x0, u1e4, 0.1, 0.6, warm_iters=2)
    if "config2solve1000" in which or "config2" in which:
        solve_fixture("solve_cartpole_N1000", cartpole_ocp(1000), car_x0, 0.1 * rng(1).standard_normal((1000, 1)),
                      with_seq=False)
    if "config2solve" in which:
        solve_fixture("solve_cartpole_N10000", cartpole_ocp(10000), car_x0, 0.1 * rng(1).standard_normal((10000, 1)),
                      with_seq=False)


if __name__ == "__main__":
    main()

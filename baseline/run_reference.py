#!/usr/bin/env python
"""Run the TRUE reference (casiacob/ip-parallel-optimal-control on JAX + paroc) with its own timing protocol
(ref examples/cartpole_runtime.py:115-153: jit, 1 warm-up, 10 timed calls, mean/median) — iff `jax` and `paroc`
import.  In this image they do not (no wheels, no network), so this prints the reason and exits 0; nothing else
in the repo may be labelled "reference".

    python baseline/run_reference.py [--problem cartpole|pendulum] [--N 1000]
"""
import argparse
import json
import sys
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problem", default="cartpole")
    ap.add_argument("--N", type=int, default=1000)
    ap.add_argument("--reference-path", default="/root/reference")
    args = ap.parse_args()
    try:
        import jax  # noqa: F401
        import paroc  # noqa: F401
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {e} (jax/paroc not installed, no network)"}))
        return 0
    sys.path.insert(0, args.reference_path)
    import ast
    import os
    import numpy as np
    import jax.numpy as jnp
    from jax import config
    config.update("jax_enable_x64", True)
    from noc.optimal_control_problem import OCP
    from noc.par_interior_point_newton import par_interior_point_optimal_control
    from noc.utils import wrap_angle, euler
    src = open(os.path.join(args.reference_path, "examples", f"{args.problem}_runtime.py")).read()
    tree = ast.parse(src)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    ns = {"jnp": jnp, "jax": jax, "wrap_angle": wrap_angle}
    exec(compile(tree, args.problem, "exec"), ns)
    N = args.N
    ode = ns[args.problem]
    ocp = OCP(euler(ode, 1.0 / N), ns["constraints"], ns["transient_cost"], ns["final_cost"], ns["total_cost"])
    x0 = (jnp.array([0.01, wrap_angle(-0.01), 0.01, -0.01]) if args.problem == "cartpole"
          else jnp.array([wrap_angle(0.1), -0.1]))
    u0 = jnp.array(0.1 * np.random.default_rng(1).standard_normal((N, 1)))
    solve = jax.jit(lambda u, x: par_interior_point_optimal_control(ocp, u, x))
    u, it = solve(u0, x0)
    jax.block_until_ready(u)
    ts = []
    for _ in range(10):
        t = time.time()
        u, it = solve(u0, x0)
        jax.block_until_ready(u)
        ts.append(time.time() - t)
    print(json.dumps({"impl": "reference", "problem": args.problem, "N": N, "iterations": int(it),
                      "mean_s": float(np.mean(ts)), "median_s": float(np.median(ts)),
                      "backend": jax.default_backend()}))
    return 0


if __name__ == "__main__":
    sys.exit(main())

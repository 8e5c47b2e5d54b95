"""CPU: the torch restatements of the example workloads (ipoc_b200/problems.py) against the
reference's own definitions (whose derivatives were recorded in the golden step fixtures)."""
import numpy as np
import pytest
import torch

from helpers import STEP_FIXTURES, relerr
from oracle.autodiff import Evaluator
from oracle.noc_np import Derivatives


@pytest.mark.parametrize("name", STEP_FIXTURES)
def test_problem_definitions_match_reference(golden, name):
    import ipoc_b200.problems as P
    g = golden(name)
    N = g["controls"].shape[0]
    ocp = P.make_pendulum(1.0 / N) if "pendulum" in name else P.make_cartpole(1.0 / N)
    ev = Evaluator(ocp)
    # rollout of the recorded controls reproduces the recorded states only for the cold fixtures
    if "warm" not in name:
        assert relerr(ev.rollout(g["controls"], g["x0"]), g["states"]) < 1e-13
    d = ev.derivatives(g["states"], g["controls"], float(g["bp"]))
    for f in Derivatives._fields:
        ref = g["d_" + f]
        got = getattr(d, f)
        assert got.shape == ref.shape, f
        assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))), f
    assert relerr(ev.final_cost_grad(g["states"][-1]), g["ref_lamT"]) < 1e-13
    assert abs(ev.total_cost(g["states"], g["controls"], float(g["bp"])) - float(g["ref_cost"])) < 1e-10 * abs(
        float(g["ref_cost"]))


def test_x0_helpers():
    import ipoc_b200.problems as P
    assert torch.allclose(P.pendulum_x0(), torch.tensor([0.1, -0.1], dtype=torch.float64))
    x = P.cartpole_x0()
    assert abs(float(x[1]) - (2 * np.pi - 0.01)) < 1e-15

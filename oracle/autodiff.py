"""CPU evaluator of the user's OCP callables for the oracle driver loops.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Plays the part JAX autodiff plays in the reference
(ref noc/par_interior_point_newton.py:13-28, noc/utils.py:57-63): the user
functions are torch callables (float64) and are differentiated with
`torch.func` on the CPU; results are handed to the NumPy oracle as ndarrays.
The index order of every derivative tensor is (output, wrt_1, wrt_2), as in JAX.
"""
import numpy as np
import torch
from torch.func import vmap, grad, hessian, jacrev
from .noc_np import Derivatives


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64, copy=True)


class Evaluator:
    def __init__(self, ocp):
        self.ocp = ocp  # any 5-tuple (dynamics, constraints, stage_cost, final_cost, total_cost)

    @staticmethod
    def _t(a):
        return torch.as_tensor(np.asarray(a, dtype=np.float64))

    def rollout(self, controls, initial_state):           # ref noc/utils.py:57-63
        dyn = self.ocp.dynamics
        x = self._t(initial_state)
        xs = [x]
        for u in self._t(controls):
            x = dyn(x, u)
            xs.append(x)
        return _np(torch.stack(xs))

    def derivatives(self, states, controls, bp):          # ref noc/par_interior_point_newton.py:13-28
        o = self.ocp
        x, u = self._t(states)[:-1], self._t(controls)
        bp = float(bp)

        def body(xk, uk):
            cx, cu = grad(o.stage_cost, (0, 1))(xk, uk, bp)
            cxx = hessian(o.stage_cost, 0)(xk, uk, bp)
            cuu = hessian(o.stage_cost, 1)(xk, uk, bp)
            cxu = jacrev(jacrev(o.stage_cost, 0), 1)(xk, uk, bp)
            fx, fu = jacrev(o.dynamics, (0, 1))(xk, uk)
            fxx = jacrev(jacrev(o.dynamics, 0), 0)(xk, uk)
            fuu = jacrev(jacrev(o.dynamics, 1), 1)(xk, uk)
            fxu = jacrev(jacrev(o.dynamics, 0), 1)(xk, uk)
            return cx, cu, cxx, cuu, cxu, fx, fu, fxx, fuu, fxu

        return Derivatives(*(_np(t) for t in vmap(body)(x, u)))

    def final_cost_grad(self, xN):                        # ref noc/costates.py:35
        return _np(grad(self.ocp.final_cost)(self._t(xN)))

    def final_cost_hess(self, xN):                        # ref noc/seq_interior_point_newton.py:66
        return _np(hessian(self.ocp.final_cost)(self._t(xN)))

    def total_cost(self, states, controls, bp):
        return float(self.ocp.total_cost(self._t(states), self._t(controls), float(bp)))

    def constraints(self, states, controls):              # ref noc/par_interior_point_newton.py:45-46
        return _np(vmap(self.ocp.constraints)(self._t(states)[:-1], self._t(controls)))

    def feasible(self, states, controls):                 # ref noc/par_interior_point_newton.py:47
        return bool(np.all(self.constraints(states, controls) <= 0))

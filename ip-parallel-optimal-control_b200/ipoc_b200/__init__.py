"""ipoc_b200 — B200-native (sm_100a, FP64) Newton step of the parallel-in-time interior-point
optimal-control solver of casiacob/ip-parallel-optimal-control.

Reference-facing names (same signatures as the reference, torch CUDA tensors for jnp arrays):
    from ipoc_b200.noc import OCP, par_interior_point_optimal_control, newton_oc, par_Newton, par_costates
    from ipoc_b200.paroc import LQT, par_bwd_pass, par_fwd_pass
"""
from .optimal_control_problem import OCP, Derivatives, LinearizedOCP  # noqa: F401
from .paroc import LQT, par_bwd_pass, par_fwd_pass, seq_bwd_pass, seq_fwd_pass  # noqa: F401
from .batched import par_interior_point_optimal_control_batched, newton_oc_batched  # noqa: F401
from .noc import (compute_derivatives, compute_lqr_params, check_traj_feasibility, noc_to_lqt,  # noqa: F401
                  par_costates, par_Newton, newton_oc, par_interior_point_optimal_control,
                  newton_step, affine_scan, reductions, accept_update)
from .utils import wrap_angle, euler, discretize_dynamics, rollout, rollout_parallel, runge_kutta  # noqa: F401
from . import problems, plants  # noqa: F401

__version__ = "0.1.0"

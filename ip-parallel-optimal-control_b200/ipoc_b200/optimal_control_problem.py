"""Problem / derivative containers — same names and field order as the reference
(ref noc/optimal_control_problem.py:5-30), holding torch tensors instead of jnp arrays."""
from typing import NamedTuple, Callable
import torch


class OCP(NamedTuple):
    dynamics: Callable      # (x (nx,), u (nu,)) -> (nx,)
    constraints: Callable   # (x, u) -> (nc,)   feasible iff all <= 0
    stage_cost: Callable    # (x, u, bp) -> ()
    final_cost: Callable    # (x,) -> ()
    total_cost: Callable    # (xs (N+1,nx), us (N,nu), bp) -> ()


class Derivatives(NamedTuple):
    cx: torch.Tensor
    cu: torch.Tensor
    cxx: torch.Tensor
    cuu: torch.Tensor
    cxu: torch.Tensor
    fx: torch.Tensor
    fu: torch.Tensor
    fxx: torch.Tensor
    fuu: torch.Tensor
    fxu: torch.Tensor


class LinearizedOCP(NamedTuple):
    r: torch.Tensor
    Q: torch.Tensor
    R: torch.Tensor
    M: torch.Tensor

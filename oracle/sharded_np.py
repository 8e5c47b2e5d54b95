"""NumPy model of the time-sharded Newton step (reduce -> all-gather carries -> seeded local scan).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  No reference counterpart: the reference is
single-device.  Used by the world_size-2 gloo test to validate the exchange logic (which rank's
carry is applied in which order, what is seeded where) against the unsharded oracle.
"""
import numpy as np
from . import paroc_np, noc_np


def segment_elements(fx, fu, ru, Q, R, M, reg):
    """Per-step Riccati elements of one segment (no terminal element)."""
    nu = R.shape[1]
    lqt = noc_np.noc_to_lqt(ru, Q, R + reg * np.eye(nu)[None], M, fx, fu)
    A, b, C, eta, J = paroc_np.bwd_elements(lqt)
    return lqt, tuple(a[:-1] for a in (A, b, C, eta, J))


def bwd_reduce(fx, fu, ru, Q, R, M, reg):
    """Segment aggregate (A, b, C, eta, J): fold from the segment's end to its start."""
    lqt, el = segment_elements(fx, fu, ru, Q, R, M, reg)
    agg = tuple(e[-1:] for e in el)
    for k in range(el[0].shape[0] - 2, -1, -1):
        agg = paroc_np.combine(tuple(e[k:k + 1] for e in el), agg)
    return lqt, tuple(a[0] for a in agg)


def apply_elem(e, S, v):
    """Value function (S, v) at the end of segment e -> at its start."""
    A, b, C, eta, J = e
    n = A.shape[0]
    W = np.eye(n) + S @ C
    Y = np.linalg.solve(W, np.concatenate([(v - S @ b)[:, None], S @ A], axis=1))
    return A.T @ Y[:, 1:] + J, A.T @ Y[:, 0] + eta


def bwd_apply(lqt, rank, nranks, carries, ST):
    """Seed from the later ranks' carries, then the plain seeded recursion on this segment."""
    n = ST.shape[0]
    S, v = 0.5 * (ST + ST.T), np.zeros(n)
    for r in range(nranks - 1, rank, -1):
        S, v = apply_elem(carries[r], S, v)
    lq = lqt._replace(XT=S, HT=np.eye(n), rT=np.linalg.solve(S, v) if np.any(v) else np.zeros(n))
    # seq_bwd_pass takes the terminal value through (XT, HT, rT): S_T = XT, v_T = XT rT
    Kx, d, Sall, vall = paroc_np.seq_bwd_pass(lq)
    _, _, pred, feas = paroc_np.gains(lq, Sall, vall)
    Fcl = lqt.A - lqt.B @ Kx
    ccl = np.einsum("tij,tj->ti", lqt.B, d)
    F, c = np.eye(n), np.zeros(n)
    for k in range(Fcl.shape[0]):
        F, c = Fcl[k] @ F, Fcl[k] @ c + ccl[k]
    return Kx, d, pred, feas, (F, c)


def fwd_apply(lqt, rank, fwd_carries, Kx, d):
    n = lqt.A.shape[1]
    x0 = np.zeros(n)
    for r in range(rank):
        F, c = fwd_carries[r]
        x0 = F @ x0 + c
    u, x = paroc_np.seq_fwd_pass(lqt, x0, Kx, d)
    return x, u

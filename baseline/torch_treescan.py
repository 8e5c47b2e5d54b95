"""B3 of BASELINE.md §3 — a "library tree-scan on B200" PROXY for the reference's XLA-on-GPU path.

The reference's Newton step is `lax.associative_scan` over `vmap`ped combines; XLA lowers that to a
log-depth odd/even recursion in which every level is a handful of small batched kernels (batched LU /
triangular solves, batched matmuls, slices, concatenations).  JAX is not installed here, so this file
re-creates that *structure* with PyTorch batched ops on the GPU: the same recursion
(SURVEY.md Appendix B), `torch.linalg.solve` for the (I + C J) systems, eager launches.
It is NOT the reference and is labelled as a proxy wherever it is reported.  Test/bench
infrastructure only — the product path never imports it.
"""
import torch


def _interleave(even, odd):
    n = even.shape[0] + odd.shape[0]
    out = torch.empty((n,) + tuple(even.shape[1:]), dtype=even.dtype, device=even.device)
    out[0::2] = even
    out[1::2] = odd
    return out


def associative_scan(fn, elems, reverse=False):
    elems = tuple(elems)
    if reverse:
        elems = tuple(torch.flip(e, [0]) for e in elems)

    def rec(es):
        n = es[0].shape[0]
        if n < 2:
            return es
        red = fn(tuple(e[0:n - 1:2] for e in es), tuple(e[1:n:2] for e in es))
        odd = rec(red)
        tail = tuple(e[2:n:2] for e in es)
        if tail[0].shape[0] == 0:
            even = tuple(e[0:1] for e in es)
        else:
            even = fn(tuple(o[:-1] for o in odd), tail) if n % 2 == 0 else fn(odd, tail)
            even = tuple(torch.cat([e[0:1], r]) for e, r in zip(es, even))
        return tuple(_interleave(a, b) for a, b in zip(even, odd))

    res = rec(elems)
    if reverse:
        res = tuple(torch.flip(r, [0]) for r in res)
    return res


def _mv(a, x):
    return (a @ x.unsqueeze(-1)).squeeze(-1)


def ric_combine(e1, e2):
    A1, b1, C1, eta1, J1 = e1
    A2, b2, C2, eta2, J2 = e2
    n = A1.shape[-1]
    eye = torch.eye(n, dtype=A1.dtype, device=A1.device)
    W = eye + C1 @ J2
    rhs = torch.cat([A1, (b1 + _mv(C1, eta2)).unsqueeze(-1), C1 @ A2.transpose(-1, -2)], dim=-1)
    sol = torch.linalg.solve(W, rhs)
    A = A2 @ sol[..., :n]
    b = _mv(A2, sol[..., n]) + b2
    C = A2 @ sol[..., n + 1:] + C2
    Wt = eye + J2 @ C1
    rhs2 = torch.cat([(eta2 - _mv(J2, b1)).unsqueeze(-1), J2 @ A1], dim=-1)
    sol2 = torch.linalg.solve(Wt, rhs2)
    eta = _mv(A1.transpose(-1, -2), sol2[..., 0]) + eta1
    J = A1.transpose(-1, -2) @ sol2[..., 1:] + J1
    return A, b, C, eta, J


def aff_combine(e1, e2):
    F1, c1 = e1
    F2, c2 = e2
    return F2 @ F1, _mv(F2, c1) + c2


def par_newton(fx, fu, ru, Q, R, M, reg):
    """par_Newton (ref noc/par_interior_point_newton.py:107-124) with tree scans of batched torch ops."""
    N, nx, nu = fx.shape[0], fx.shape[1], fu.shape[-1]
    o = dict(dtype=fx.dtype, device=fx.device)
    U = R + reg * torch.eye(nu, **o)
    XiM = torch.linalg.solve(Q, M)
    s = -torch.linalg.solve(U - M.transpose(1, 2) @ XiM, ru.unsqueeze(-1)).squeeze(-1)
    r = -_mv(XiM, s)
    UiMt = torch.linalg.solve(U, M.transpose(1, 2))
    UiBt = torch.linalg.solve(U, fu.transpose(1, 2))
    p = -(_mv(U, s) + _mv(M.transpose(1, 2), r))
    q = -(_mv(Q, r) + _mv(M, s))
    Uip = torch.linalg.solve(U, p.unsqueeze(-1)).squeeze(-1)
    z = lambda *sh: torch.zeros(*sh, **o)
    A = torch.cat([fx - fu @ UiMt, z(1, nx, nx)])
    b = torch.cat([-_mv(fu, Uip), z(1, nx)])
    C = torch.cat([fu @ UiBt, z(1, nx, nx)])
    J = torch.cat([Q - M @ UiMt, Q[0:1]])
    eta = torch.cat([-q + _mv(M, Uip), z(1, nx)])
    _, _, _, v, S = associative_scan(lambda later, earlier: ric_combine(earlier, later), (A, b, C, eta, J), reverse=True)
    Sn, vn = S[1:], v[1:]
    BtS = fu.transpose(1, 2) @ Sn
    G = U + BtS @ fu
    Kx = torch.linalg.solve(G, M.transpose(1, 2) + BtS @ fx)
    d = torch.linalg.solve(G, (-p + _mv(fu.transpose(1, 2), vn)).unsqueeze(-1)).squeeze(-1)
    pred = -0.5 * torch.sum(d * _mv(G, d))
    feas = torch.all(torch.linalg.eigvalsh(G) > 0)
    Ft = fx - fu @ Kx
    ct = _mv(fu, d)
    tF = Ft.clone()
    tc = ct.clone()
    tF[0] = 0.0
    _, xs = associative_scan(aff_combine, (tF, tc))
    dx = torch.cat([z(1, nx), xs])
    du = -_mv(Kx, dx[:-1]) + d
    return dx, du, pred, feas

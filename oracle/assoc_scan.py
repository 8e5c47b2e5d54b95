"""NumPy restatement of `jax.lax.associative_scan`'s odd/even recursion.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference relies on it at noc/costates.py:15-16 (reverse=False) and, inside
the absent `paroc`, for the Riccati (reverse) and forward scans
(call sites noc/par_interior_point_newton.py:120-123).  JAX's source is not on
this machine; the recursion below is the published algorithm (SURVEY.md
Appendix B): pairwise-combine neighbours, recurse on the half-length sequence,
then fill in the even positions.  `fn(a, b)`: `a` is the earlier
(already-accumulated) operand, `b` the later one; both are tuples of arrays
batched along axis 0.  reverse=True == flip(scan(flip(elems))).
"""
import numpy as np


def _interleave(even, odd):
    n = even.shape[0] + odd.shape[0]
    out = np.empty((n,) + even.shape[1:], dtype=even.dtype)
    out[0::2] = even
    out[1::2] = odd
    return out


def associative_scan(fn, elems, reverse=False):
    elems = tuple(np.asarray(e) for e in elems)
    if reverse:
        elems = tuple(e[::-1] for e in elems)

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
            if n % 2 == 0:
                even = fn(tuple(o[:-1] for o in odd), tail)
            else:
                even = fn(odd, tail)
            even = tuple(np.concatenate([e[0:1], r]) for e, r in zip(es, even))
        return tuple(_interleave(a, b) for a, b in zip(even, odd))

    res = rec(elems)
    if reverse:
        res = tuple(r[::-1] for r in res)
    return res


def serial_scan(fn, elems, reverse=False):
    """Left fold with the same `fn` — the simplest possible truth for tests."""
    elems = tuple(np.asarray(e) for e in elems)
    if reverse:
        elems = tuple(e[::-1] for e in elems)
    n = elems[0].shape[0]
    outs = [tuple(e[0:1] for e in elems)]
    for i in range(1, n):
        outs.append(fn(outs[-1], tuple(e[i:i + 1] for e in elems)))
    res = tuple(np.concatenate([o[j] for o in outs]) for j in range(len(elems)))
    if reverse:
        res = tuple(r[::-1] for r in res)
    return res

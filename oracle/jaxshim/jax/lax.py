"""Fake `jax.lax`: eager Python loops with JAX's semantics (see _core.py).

`associative_scan` restates the odd/even recursion of JAX's implementation
 so that the combine order — and hence rounding — follows
JAX's, given that the real thing is unavailable.
"""
import torch
from ._core import wrap


def _leaves(tree):
    if isinstance(tree, torch.Tensor) or not isinstance(tree, (tuple, list)):
        return [tree], None
    return list(tree), type(tree)


def _rebuild(kind, leaves, proto=None):
    if kind is None:
        return leaves[0]
    if proto is not None and hasattr(proto, "_fields"):
        return type(proto)(*leaves)
    return kind(leaves)


def scan(f, init, xs=None, length=None, reverse=False):
    xl, xkind = _leaves(xs) if xs is not None else ([], None)
    n = length if xs is None else xl[0].shape[0]
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, outs = init, [None] * n
    for i in order:
        x_i = None if xs is None else _rebuild(xkind, [a[i] for a in xl], xs)
        carry, outs[i] = f(carry, x_i)
    if n == 0 or outs[0] is None:
        return carry, None
    ol, okind = _leaves(outs[0])
    if okind is None:
        stacked = wrap(torch.stack([torch.as_tensor(o) for o in outs]))
    else:
        stacked = _rebuild(okind, [wrap(torch.stack([torch.as_tensor(o[j]) for o in outs]))
                                   for j in range(len(ol))], outs[0])
    return carry, stacked


def while_loop(cond, body, init):
    val = init
    while bool(cond(val)):
        val = body(val)
    return val


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(lower, upper):
        val = body_fun(i, val)
    return val


def _interleave(a, b):
    # a: even positions (len ceil), b: odd positions
    n = a.shape[0] + b.shape[0]
    out = torch.empty((n,) + tuple(a.shape[1:]), dtype=a.dtype)
    out[0::2] = a
    out[1::2] = b
    return out


def associative_scan(fn, elems, reverse=False):
    leaves, kind = _leaves(elems)
    leaves = [torch.as_tensor(e) for e in leaves]
    if reverse:
        leaves = [torch.flip(e, [0]) for e in leaves]

    def call(a, b):
        out = fn(_rebuild(kind, a, elems), _rebuild(kind, b, elems))
        return [torch.as_tensor(o) for o in _leaves(out)[0]]

    def rec(es):
        n = es[0].shape[0]
        if n < 2:
            return es
        red = call([e[0:n - 1:2] for e in es], [e[1:n:2] for e in es])
        odd = rec(red)
        if es[0][2:n:2].shape[0] == 0:
            even = [e[0:1] for e in es]
        else:
            if n % 2 == 0:
                even = call([o[:-1] for o in odd], [e[2:n:2] for e in es])
            else:
                even = call(odd, [e[2:n:2] for e in es])
            even = [torch.cat([e[0:1], r]) for e, r in zip(es, even)]
        return [_interleave(a, b) for a, b in zip(even, odd)]

    res = rec(leaves)
    if reverse:
        res = [torch.flip(r, [0]) for r in res]
    return _rebuild(kind, [wrap(r) for r in res], elems)

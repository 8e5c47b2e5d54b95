"""Minimal stand-in for the parts of JAX that the reference's `noc/` modules use.

TEST INFRASTRUCTURE ONLY.  JAX is not installed in the build image, so the
reference (pure Python/JAX, /root/reference/noc/*.py) cannot be imported as is.
This shim maps the handful of `jax`, `jax.numpy`, `jax.lax`, `jax.scipy` names
those files touch onto float64 PyTorch-CPU + `torch.func`, so that the
reference's *own, unmodified source files* can be executed in this container
to generate golden vectors (tests/golden/gen_golden.py).  It is never imported
by the product package, by `bench.py`'s GPU arm or by anything that runs on the
GPU box.

`Arr` is a torch.Tensor subclass that adds the two NumPy/JAX indexing habits
the reference relies on and torch lacks: negative-step slices (`x[::-1]`,
ref noc/costates.py:36-37,40) and `.T` on 1-D arrays
(ref examples/pendulum_runtime.py:36).
"""
import torch

torch.set_default_dtype(torch.float64)


class Arr(torch.Tensor):
    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        flips, new, out_dim = [], [], 0
        for it in idx:
            if it is Ellipsis:
                raise NotImplementedError("shim: Ellipsis indexing")
            if isinstance(it, slice) and it.step is not None and it.step < 0:
                if not (it.start is None and it.stop is None and it.step == -1):
                    raise NotImplementedError("shim: only [::-1] negative slices")
                flips.append(out_dim)
                new.append(slice(None))
                out_dim += 1
            elif isinstance(it, int):
                new.append(it)  # removes a dim
            else:
                new.append(it)  # None adds a dim, slice keeps one
                out_dim += 1
        out = super().__getitem__(tuple(new))
        if flips:
            out = torch.flip(out, flips)
        return out

    @property
    def T(self):
        if self.dim() < 2:
            return self
        return self.permute(*range(self.dim() - 1, -1, -1))


def wrap(x):
    """Re-tag plain tensors (e.g. torch.func outputs) as Arr, recursively."""
    if isinstance(x, torch.Tensor):
        return x.as_subclass(Arr) if type(x) is torch.Tensor else x
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return type(x)(*(wrap(v) for v in x))
    if isinstance(x, (tuple, list)):
        return type(x)(wrap(v) for v in x)
    return x


def asarr(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, bool):
        return wrap(torch.tensor(x))
    if isinstance(x, (list, tuple)) and any(isinstance(v, torch.Tensor) for v in x):
        return wrap(torch.stack([asarr(v) for v in x]))
    return wrap(torch.as_tensor(x, dtype=dtype if dtype is not None else
                                (torch.float64 if _is_floaty(x) else None)))


def _is_floaty(x):
    if isinstance(x, float):
        return True
    if isinstance(x, (list, tuple)):
        return any(_is_floaty(v) for v in x)
    try:
        import numpy as np
        if isinstance(x, np.ndarray):
            return x.dtype.kind == "f"
    except Exception:
        pass
    return False

"""Fake `jax.random`: NOT threefry — inputs for goldens are drawn with NumPy instead."""
import numpy as _np
from ._core import asarr


def PRNGKey(seed):
    return int(seed)


def normal(key, shape=()):
    return asarr(_np.random.default_rng(key).standard_normal(shape))

"""Shared helpers of the parity tests: oracle fields <-> gpt_b200 lattices."""
import functools

import numpy as np

from oracle import qcd
from oracle.rng import random as oracle_random


def rel(a, b):
    a = np.asarray(a, dtype=np.complex128).ravel()
    b = np.asarray(b, dtype=np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def sites(arr, ncomp_dims):
    """oracle layout [T,Z,Y,X,(S),tensor...] -> GPT order [sites, tensor...]"""
    t = arr.shape[len(arr.shape) - ncomp_dims:]
    return np.ascontiguousarray(arr.reshape((-1,) + t))


@functools.lru_cache(maxsize=None)
def gauge(seed, dims, scale=1.0):
    rng = oracle_random(seed)
    U = qcd.gauge_random(rng, list(dims), scale=scale)
    return rng, U


def to_links(g, grid, U):
    return g.qcd.gauge.from_numpy(grid, [sites(u.astype(grid.precision.complex_dtype), 2) for u in U])


def to_spinor(g, grid, arr, cb=None):
    """oracle full-lattice spinor -> gpt_b200 lattice on `grid` (full, or the `cb` half if grid is red-black)"""
    five_d = grid.nd == 5
    l = g.vspincolor(grid)
    if grid.cb.n == 1:
        l[:] = sites(arr, 2).astype(grid.precision.complex_dtype)
    else:
        half = qcd.pick_checkerboard(arr, cb.tag, ls=five_d)
        l.checkerboard(cb)
        l[:] = half.reshape(-1, 4, 3).astype(grid.precision.complex_dtype)
    return l


def from_spinor(l, like):
    """gpt_b200 lattice -> oracle layout; half lattices are embedded into zeros"""
    a = l[:]
    five_d = l.grid.nd == 5
    if l.grid.cb.n == 1:
        return a.reshape(like.shape)
    out = np.zeros(like.shape, dtype=a.dtype)
    tail = like.shape[4:]
    qcd.set_checkerboard(out, a.reshape((-1,) + tail), l.checkerboard().tag, ls=five_d)
    return out

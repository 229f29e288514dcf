"""
CPU tests of the product's random number generator (gpt_b200/csrc/rng.cu, host code inside libcgpt_b200.so; no device
needed): the reference's known-answer vectors (/root/reference/tests/random/simple.py:33-144) and element-wise agreement
with the oracle restatement (oracle/rng.py) for every distribution, both engines, 4d and 5d grids.
"""
import numpy as np
import pytest

import gpt_b200 as g
from oracle.rng import random as oracle_random

NORMAL = {"distribution": "normal", "mu": 0.0, "sigma": 1.0}


def _at(a, dims, x, y, z, t):
    return a[x + dims[0] * (y + dims[1] * (z + dims[2] * t)), 0]


def test_product_rng_lattice_kat():
    dims = [8, 4, 4, 4]
    grid = g.grid(dims, g.double)
    rng = g.random("block_seed_string_13")
    v = rng.host_array(grid, 1, NORMAL)
    pts = [(0, 0, 0, 0), (2, 0, 0, 0), (0, 2, 0, 0), (1, 3, 1, 3), (3, 2, 1, 0)]
    got = np.array([_at(v, dims, *p).real for p in pts])
    ref = np.array([-0.29101665386129116, -1.4591269443435488, -0.3641310411719848, -0.9454532383815435, 0.4996115272362977])
    assert np.linalg.norm(got - ref) < 1e-14
    assert np.all(v.imag == 0.0)
    for _ in range(1000):
        v = rng.host_array(grid, 1, NORMAL)
    got = np.array([_at(v, dims, *p).real for p in pts])
    ref = np.array([1.473846437649123, 0.06134886004475955, -1.4849224560837744, -0.2316303634513769, -0.5309613807759392])
    assert np.linalg.norm(got - ref) < 1e-14


def test_product_rng_scalar_kat():
    # simple.py:33-66,132-144
    rng = g.random("block_seed_string_13")
    zs = {}
    for _ in range(10000):
        z = rng.zn()
        zs[z] = zs.get(z, 0) + 1
    assert len(zs) == 2 and all(abs(abs(z) - 1.0) < 1e-15 for z in zs)
    for _ in range(10000):
        rng.zn(n=3)
    m = np.array([rng.normal() for _ in range(10000)])
    assert abs(m.mean()) < 0.05 and abs((m**2).mean() - 1.0) < 0.05
    assert rng.choice(["A", "B", "C"], 10) == ["C", "C", "A", "C", "C", "B", "A", "B", "C", "A"]
    assert np.all(rng.choice(np.array([1, 2, 3]), 5) == np.array([2, 3, 3, 3, 1]))


@pytest.mark.parametrize("engine", ["vectorized_ranlux24_389_64", "vectorized_ranlux24_24_64"])
@pytest.mark.parametrize("dims,nel", [([4, 4, 4, 8], 12), ([6, 4, 4, 4, 4], 12), ([4, 6, 2, 4], 1)])
def test_product_rng_matches_oracle(engine, dims, nel):
    grid = g.grid(dims, g.double)
    rng = g.random("compare with the oracle", engine)
    orng = oracle_random("compare with the oracle", engine)
    shape = (nel,) if nel > 1 else ()
    for dist, p, okw in [
        ("cnormal", {"mu": 0.0, "sigma": 1.0}, {}),
        ("uniform_real", {"min": -0.5, "max": 0.5}, {"min": -0.5, "max": 0.5}),
        ("normal", {"mu": 1.5, "sigma": 2.0}, {"mu": 1.5, "sigma": 2.0}),
        ("cnormal", {"mu": 0.0, "sigma": 1.0}, {}),
    ]:
        got = rng.host_array(grid, nel, {"distribution": dist, **p})
        ref = getattr(orng, dist)(dims, shape, **okw).reshape(-1, nel)
        # the integer stream is identical; libm's log / sin / cos may differ from numpy's in the last bit
        assert np.max(np.abs(got - ref)) < 1e-14, dist


def test_product_rng_grid_objects_and_errors():
    rng = g.random("abc")
    g1, g2 = g.grid([4, 4, 4, 4], g.double), g.grid([4, 4, 4, 4], g.single)
    a = rng.host_array(g1, 1, NORMAL)
    b = rng.host_array(g2, 1, NORMAL)
    assert np.array_equal(a, b)  # a grid of another precision is another GridBase*: it starts from the seed (engine.h:82-99; simple.py:19-20)
    a2 = rng.host_array(g1, 1, NORMAL)
    assert not np.array_equal(a, a2)
    # cgpt interns grids by (fdimensions, simd, cb, mpi) (lib/cgpt/lib/grid.h:33-82): a second grid OBJECT of the same shape is the
    # same GridBase* and continues the stream instead of repeating it
    g3 = g.grid([4, 4, 4, 4], g.double)
    assert g3 == g1 and g3 is not g1
    a3 = rng.host_array(g3, 1, NORMAL)
    assert not np.array_equal(a3, a) and not np.array_equal(a3, a2)
    orng = oracle_random("abc")
    ref = [orng.normal([4, 4, 4, 4]).reshape(-1, 1) for _ in range(3)]
    for got, want in zip([a, a2, a3], ref):
        assert np.max(np.abs(got - want)) < 1e-14
    assert g1.converted(g.single).serial == g2.serial and g1.checkerboarded(g.redblack).serial != g1.serial
    with pytest.raises(RuntimeError):
        g.random("abc", "no such engine")
    with pytest.raises(ValueError):
        g.grid([3, 4, 4, 4], g.double)  # odd extent: neither checkerboards nor 2^4 rng blocks
    with pytest.raises(KeyError):
        rng.normal(None, {"sigmaa": 1.0})


@pytest.mark.parametrize("mpi", [[1, 1, 1, 2], [1, 1, 2, 2], [1, 2, 1, 1]])
def test_product_rng_is_decomposition_independent(mpi):
    """every rank draws the blocks of its part of the global lattice (block seeds carry the global block index,
    lib/cgpt/lib/random/parallel.h:74-89): stitched together, the ranks' fields are the single-rank field"""
    from gpt_b200 import cgpt

    gd = [4, 4, 4, 8]
    p = {"distribution": "cnormal", "mu": 0.0, "sigma": 1.0}
    for five_d in (False, True):
        g_dims = ([3] if five_d else []) + gd
        m = ([1] if five_d else []) + mpi
        rng = g.random("decomposition")
        ref = cgpt.random_sample_host(rng.obj, 1, g_dims, g_dims, [0] * len(g_dims), 12, p).reshape(g_dims[::-1] + [12])
        ld = [d // k for d, k in zip(g_dims, m)]
        out = np.zeros_like(ref)
        key = 100
        for rank in range(int(np.prod(m))):
            co, r = [], rank
            for k in m:
                co.append(r % k)
                r //= k
            ls = [c * n for c, n in zip(co, ld)]
            rr = g.random("decomposition")  # a fresh engine per rank, like one process per GPU
            key += 1
            loc = cgpt.random_sample_host(rr.obj, key, ld, g_dims, ls, 12, p).reshape(ld[::-1] + [12])
            sl = tuple(slice(s, s + n) for s, n in zip(ls[::-1], ld[::-1]))
            out[sl] = loc
        assert np.array_equal(out, ref)

"""g.qcd.gauge.{unit, random} (lib/gpt/qcd/gauge/create.py:24-80), g.qcd.gauge.plaquette (lib/gpt/qcd/gauge/stencil/plaquette.py:23-44)"""
import numpy as np

import gpt_b200 as g


def unit(grid):
    U = []
    n = grid.lsites
    a = np.zeros((n, 3, 3), dtype=grid.precision.complex_dtype)
    a[:, range(3), range(3)] = 1.0
    for mu in range(4):
        u = g.mcolor(grid)
        u[:] = a
        U.append(u)
    return U


def from_numpy(grid, arrays):
    """four [sites,3,3] arrays in GPT order -> list of link lattices"""
    U = []
    for mu in range(4):
        u = g.mcolor(grid)
        u[:] = np.asarray(arrays[mu]).reshape(grid.lsites, 3, 3)
        U.append(u)
    return U


def random(grid, rng, scale=1.0):
    """g.qcd.gauge.random(grid, rng, scale): needs a gpt_b200.random engine"""
    return rng.element_links(grid, scale)


def plaquette(U):
    """average of Re tr P_{mu nu} / Nc over the six planes and all sites, computed on the device"""
    return g.cgpt.gauge_plaquette([u.obj for u in U])[0]


def link_trace(U):
    """average of Re tr U_mu / Nc (the LINK_TRACE of a NERSC header)"""
    return g.cgpt.gauge_plaquette([u.obj for u in U])[1]

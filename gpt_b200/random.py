"""g.random (lib/gpt/core/random.py:23-165): GPT's parallel random number generator, same streams as the reference
(RANLUX24 lanes seeded by SHA-256, one generator per 2^4 block of sites; gpt_b200/csrc/rng.cu)."""
import numpy as np

import gpt_b200 as g
import cgpt
from gpt_b200 import capi
from gpt_b200.params import params_convention


class random:
    def __init__(self, first, second=None):
        if isinstance(first, dict) and second is None:
            s, engine = first["seed"], first["engine"]
        else:
            s, engine = first, second
        if engine is None:
            engine = "vectorized_ranlux24_389_64"
        self.seed, self.engine = s, engine
        self.obj = cgpt.create_random(engine, s)

    def __del__(self):
        if getattr(self, "obj", None) is not None:
            cgpt.delete_random(self.obj)
            self.obj = None

    def sample(self, t, p):
        if isinstance(t, list):
            for x in t:
                self.sample(x, p)
            return t
        if t is None:
            return cgpt.random_sample(self.obj, p)
        if isinstance(t, g.lattice):
            cgpt.random_sample(self.obj, {**p, **{"lattices": [t]}})
            return t
        raise TypeError(f"cannot sample into {type(t)}")

    @params_convention(mu=0.0, sigma=1.0)
    def normal(self, t=None, p={}):
        r = self.sample(t, {**{"distribution": "normal"}, **p})
        return r.real if t is None else r

    @params_convention(mu=0.0, sigma=1.0)
    def cnormal(self, t=None, p={}):
        return self.sample(t, {**{"distribution": "cnormal"}, **p})

    @params_convention(min=0.0, max=1.0)
    def uniform_real(self, t=None, p={}):
        r = self.sample(t, {**{"distribution": "uniform_real"}, **p})
        return r.real if t is None else r

    @params_convention(min=0, max=1)
    def uniform_int(self, t=None, p={}):
        r = self.sample(t, {**{"distribution": "uniform_int"}, **p})
        return int(r.real) if t is None else r

    @params_convention(n=2)
    def zn(self, t=None, p={}):
        return self.sample(t, {**{"distribution": "zn"}, **p})

    def choice(self, array, n):
        idx = [self.uniform_int(min=0, max=len(array) - 1) for i in range(n)]
        if isinstance(array, np.ndarray):
            return np.take(array, idx, axis=0)
        return [array[i] for i in idx]

    def element_links(self, grid, scale):
        """four SU(3) link lattices exp(i scale sum_a u_a T_a), u_a uniform in [-1/2, 1/2): what random.element() gives for
        the colour matrices of g.qcd.gauge.random (lib/gpt/core/random.py:110-148, lib/gpt/qcd/gauge/create.py:66-71)"""
        U = [g.mcolor(grid) for mu in range(4)]
        capi.random_su3_links(self.obj, grid.serial, [u.obj for u in U], scale)
        return U

    def host_array(self, grid, nel, p):
        """[sites, nel] complex128 array in GPT order drawn from this grid's generators without touching the device
        (same stream as sampling a lattice with nel complex components on `grid`)"""
        from gpt_b200 import parallel

        ld = list(grid.ldimensions)
        gd = list(grid.fdimensions)
        coor = ([0] if grid.nd == 5 else []) + list(parallel.processor_coor(parallel.rank, parallel.mpi))
        ls = [c * n for c, n in zip(coor, ld)]
        return capi.random_sample_host(self.obj, grid.serial, ld, gd, ls, nel, p)

"""
Host-side mirror of the slice of GPT's core objects that the fermion-operator hot path touches: precision,
grid, lattice, lazy linear expressions, matrix_operator and the vector transforms.  Names, argument meaning
and error behaviour follow the reference (file:line relative to /root/reference):

  precision        lib/gpt/core/precision.py
  grid             lib/gpt/core/grid.py:31-37,96-271
  lattice          lib/gpt/core/lattice.py:57-122,301-303
  expr / eval      lib/gpt/core/expr.py:126-143,316-412
  matrix_operator  lib/gpt/core/operator/matrix_operator.py:35-305
  transforms       lib/gpt/core/transform.py:101-153, core/checkerboard.py:56-81, core/basis.py:66-74

All arithmetic happens in libcgpt_b200.so.  Calls that exist in the reference's `cgpt` module go through the signature-exact
stand-in of that module (./cgpt/__init__.py: same names, argument lists and return values, so this file talks to the library
the way lib/gpt talks to Grid); calls the reference has no counterpart for (host import / export of whole fields, the fused
kernels) go through `capi`, the ctypes binding of include/cgpt_b200.h.  This file holds no numerics.
"""
import builtins
import numbers
import sys
import time as _time

import numpy as np

import cgpt
from gpt_b200 import capi


# ---- precision -----------------------------------------------------------------------------------------
class _precision:
    def __init__(self, name, real, cplx, nbytes, eps, code):
        self.__name__ = name
        self.cgpt_dtype = name
        self.real_dtype = real
        self.complex_dtype = cplx
        self.nbytes = nbytes
        self.eps = eps
        self.code = code

    def __repr__(self):
        return self.__name__


single = _precision("single", np.float32, np.complex64, 4, 1e-7, capi.SINGLE)
double = _precision("double", np.float64, np.complex128, 8, 1e-15, capi.DOUBLE)


# ---- checkerboards ---------------------------------------------------------------------------------------
class _cb:
    def __init__(self, name, tag):
        self.__name__ = name
        self.tag = tag

    def inv(self):
        return {capi.EVEN: odd, capi.ODD: even, capi.FULL: none}[self.tag]

    def __repr__(self):
        return self.__name__


even = _cb("even", capi.EVEN)
odd = _cb("odd", capi.ODD)
none = _cb("none", capi.FULL)
_cb_of = {capi.EVEN: even, capi.ODD: odd, capi.FULL: none}


class full:
    n = 1
    __name__ = "full"


class redblack:
    n = 2
    __name__ = "redblack"


# ---- grid --------------------------------------------------------------------------------------------------
class grid:
    """g.grid(fdimensions, precision, cb=full): 4d [x,y,z,t] or 5d [s,x,y,z,t] (s never checkerboarded)."""


    def __init__(self, fdimensions, precision, cb=None, parent=None, mpi=None):
        self.fdimensions = [int(x) for x in fdimensions]
        self.gdimensions = list(self.fdimensions)
        self.ldimensions = list(self.fdimensions)
        self.nd = len(self.fdimensions)
        if self.nd not in (4, 5):
            raise ValueError("only 4d and 5d grids are supported")
        self.precision = precision
        self.cb = full if cb is None else cb
        self.parent = parent
        # processor grid over (x,y,z,t); the fifth dimension is never split (lib/gpt/core/grid.py:83-87)
        from gpt_b200 import parallel

        mpi4 = list(parallel.mpi) if mpi is None else list(mpi)[-4:]
        self.mpi = ([1] if self.nd == 5 else []) + mpi4
        self.processor = parallel.rank
        self.Nprocessors = parallel.world
        self.ldimensions = self.fdimensions[:-4] + parallel.local_dims(self.fdimensions[-4:], mpi4)
        self.gsites = int(np.prod(self.fdimensions))
        # cgpt interns grids by (fdimensions, simd layout = precision, cb mask, mpi) (lib/cgpt/lib/grid.h:33-82) and a g.random keeps
        # one set of parallel generators per grid handle (random/engine.h:82-99): grids that compare equal share a stream, a second
        # g.grid(...) of the same shape continues it
        cb_mask = [0] * self.nd
        if self.cb.n != 1:
            cb_mask[self.nd - 4] = 1  # red-black grids are checkerboarded along x; the fifth dimension never is (grid.py:31-37)
        self.obj = cgpt.create_grid(self.fdimensions, precision.cgpt_dtype, cb_mask, [1] * self.nd, self.mpi, 0)
        self.serial = self.obj

    @property
    def dims4(self):
        """local x,y,z,t extents (what libcgpt_b200 sees)"""
        return self.ldimensions[-4:]

    @property
    def Ls(self):
        return self.fdimensions[0] if self.nd == 5 else 0

    def checkerboarded(self, cb):
        return grid(self.fdimensions, self.precision, cb)

    def converted(self, precision):
        return grid(self.fdimensions, precision, self.cb)

    def inserted_dimension(self, dimension, extent, cb_mask=None):
        assert dimension == 0 and self.nd == 4
        return grid([extent] + self.fdimensions, self.precision, self.cb)

    def removed_dimension(self, dimension):
        assert dimension == 0 and self.nd == 5
        return grid(self.fdimensions[1:], self.precision, self.cb)

    def __eq__(self, other):
        return (
            isinstance(other, grid)
            and self.fdimensions == other.fdimensions
            and self.precision is other.precision
            and self.cb.n == other.cb.n
        )

    def __hash__(self):
        return hash((tuple(self.fdimensions), self.precision.__name__, self.cb.n))

    def __str__(self):
        return f"{self.fdimensions};{self.precision.__name__};{self.cb.__name__}"

    def barrier(self):
        cgpt.accelerator_barrier()

    def globalsum(self, x):
        # lib/gpt/core/global_sum.py:23-31: numbers come back as new values, arrays are summed in place
        if isinstance(x, (list, tuple)):
            x = np.array(x)
        return cgpt.grid_globalsum(self.obj, x)

    @property
    def lsites(self):
        return int(np.prod(self.ldimensions))


def _local_site(gr, pos):
    """global coordinates -> lexicographic index (dimension 0 fastest) into this rank's block, or None if another rank owns the site"""
    from gpt_b200 import parallel

    coor = ([0] if gr.nd == 5 else []) + list(parallel.processor_coor(parallel.rank, parallel.mpi))
    idx = 0
    for d in reversed(range(gr.nd)):
        x = int(pos[d]) % gr.fdimensions[d] - coor[d] * gr.ldimensions[d]
        if x < 0 or x >= gr.ldimensions[d]:
            return None
        idx = idx * gr.ldimensions[d] + x
    return idx


# ---- object types ----------------------------------------------------------------------------------------------
class _otype:
    def __init__(self, name, shape, code):
        self.__name__ = name
        self.shape = shape
        self.nfloats = 2 * int(np.prod(shape)) if shape else 2
        self.code = code
        # name of the (single) cgpt lattice type that stores this object (lib/gpt/core/object_type: v_otype)
        self.v_otype = ["ot_matrix_color(3)" if name == "ot_matrix_su_n_fundamental_group(3)" else name]
        self.v_idx = range(1)
        if name == "ot_matrix_su_n_fundamental_group(3)":
            self.Nc = 3

    def __repr__(self):
        return self.__name__


ot_singlet = _otype("ot_singlet", (), capi.OT_SINGLET)
ot_matrix_su_n_fundamental_group_3 = _otype("ot_matrix_su_n_fundamental_group(3)", (3, 3), capi.OT_MCOLOR)
ot_vector_spin_color_4_3 = _otype("ot_vector_spin_color(4,3)", (4, 3), capi.OT_VSPINCOLOR)


def ot_vector_spin_color(ns, nc):
    assert (ns, nc) == (4, 3)
    return ot_vector_spin_color_4_3


def ot_matrix_su_n_fundamental_group(nc):
    assert nc == 3
    return ot_matrix_su_n_fundamental_group_3


# ---- lattice -----------------------------------------------------------------------------------------------------
class lattice:
    """g.lattice(grid, otype) or g.lattice(other) (same grid/otype/checkerboard, uninitialised)."""

    def __init__(self, first, otype=None, device_ptr=None):
        if isinstance(first, lattice):
            self.grid = first.grid
            self.otype = first.otype
            cb = first.checkerboard().tag
        else:
            self.grid = first
            self.otype = otype
            cb = capi.FULL if first.cb.n == 1 else capi.EVEN
        g = self.grid
        if device_ptr is None:
            self.obj = cgpt.create_lattice(g.obj, self.otype.v_otype[0], g.precision.cgpt_dtype)
            if cb == capi.ODD:
                cgpt.lattice_change_checkerboard(self.obj, cb)
        else:  # view of externally owned device memory (library extension)
            self.obj = capi.create_lattice(g.dims4, g.Ls, g.precision.code, self.otype.code, cb, device_ptr)
        self.v_obj = [self.obj]

    def __del__(self):
        if getattr(self, "obj", None) is not None:
            cgpt.delete_lattice(self.obj)
            self.obj = None

    # -- checkerboard label
    def checkerboard(self, val=None):
        if val is None:
            if self.grid.cb.n == 1:
                return none
            return _cb_of[cgpt.lattice_get_checkerboard(self.obj)]
        if val is not none:
            assert self.grid.cb.n != 1
            cgpt.lattice_change_checkerboard(self.obj, val.tag)
        return self

    # -- host views in GPT order
    def _host_shape(self):
        return (int(capi.lattice_bytes(self.obj) // (self.otype.nfloats * self.grid.precision.nbytes)),) + tuple(
            self.otype.shape
        )

    def __getitem__(self, key):
        a = np.empty(self._host_shape(), dtype=self.grid.precision.complex_dtype)
        capi.lattice_export(self.obj, a)
        if isinstance(key, builtins.slice) and key == builtins.slice(None):
            return a
        if isinstance(key, tuple) and len(key) == self.grid.nd and all(isinstance(k, (int, np.integer)) for k in key):
            assert self.grid.cb.n == 1
            idx = _local_site(self.grid, key)
            v = a[idx].astype(np.complex128) if idx is not None else np.zeros(a.shape[1:], np.complex128)
            return np.asarray(self.grid.globalsum(v)).astype(a.dtype).reshape(a.shape[1:])
        raise NotImplementedError("lattice[...] supports [:] and full-lattice point access")

    def __setitem__(self, key, value):
        if isinstance(key, builtins.slice) and key == builtins.slice(None):
            if isinstance(value, numbers.Number) and value == 0:
                cgpt.lattice_set_to_number(self.obj, 0.0)
                return
            a = np.asarray(value)
            if a.shape != self._host_shape() and a.size != int(np.prod(self._host_shape())):
                raise ValueError(f"shape mismatch: {a.shape} vs {self._host_shape()}")
            capi.lattice_import(self.obj, np.ascontiguousarray(a, dtype=self.grid.precision.complex_dtype))
            return
        if isinstance(key, tuple) and len(key) == self.grid.nd:
            idx = _local_site(self.grid, key)
            if idx is not None:  # the rank that owns the site
                a = self[:]
                a[idx] = np.asarray(value, dtype=a.dtype).reshape(a[idx].shape)
                capi.lattice_import(self.obj, a)
            return
        raise NotImplementedError("lattice[...] = supports [:] and full-lattice point access")

    # -- expression sugar
    def __mul__(self, other):
        return expr(self) * other

    def __rmul__(self, other):
        return other * expr(self)

    def __add__(self, other):
        return expr(self) + other

    def __sub__(self, other):
        return expr(self) - other

    def __neg__(self):
        return -1.0 * expr(self)

    def __truediv__(self, other):
        return expr(self) * (1.0 / other)

    def __imatmul__(self, other):
        eval(self, other)
        return self

    def __iadd__(self, other):
        eval(self, other, ac=True)
        return self

    def __isub__(self, other):
        eval(self, -1.0 * expr(other), ac=True)
        return self

    def __imul__(self, other):
        capi.lattice_scale(self.obj, other)
        return self

    def __itruediv__(self, other):
        capi.lattice_scale(self.obj, 1.0 / other)
        return self

    def global_bytes(self):
        return int(capi.lattice_bytes(self.obj))


def vspincolor(grid_):
    return lattice(grid_, ot_vector_spin_color_4_3)


def mcolor(grid_):
    return lattice(grid_, ot_matrix_su_n_fundamental_group_3)


def complex(grid_):  # noqa: A001  (GPT's name)
    return lattice(grid_, ot_singlet)


def real(grid_):
    """g.real(grid): GPT stores real fields as complex singlets as well (lib/gpt/core/object_type/__init__.py)"""
    return lattice(grid_, ot_singlet)


# ---- expressions ---------------------------------------------------------------------------------------------------
class expr:
    """
    Lazy linear combination  sum_i c_i * (op_i1 * op_i2 * ... * lattice_i), evaluated right-to-left like
    lib/gpt/core/expr.py:316-330; every operator factor allocates its result from its vector space.
    """

    def __init__(self, first=None, terms=None):
        if terms is not None:
            self.terms = terms
        elif isinstance(first, expr):
            self.terms = list(first.terms)
        elif isinstance(first, lattice):
            self.terms = [(1.0, [first])]
        elif isinstance(first, matrix_operator):
            self.terms = [(1.0, [first])]
        elif first is None:
            self.terms = []
        else:
            raise TypeError(f"cannot build an expression from {type(first)}")

    def __mul__(self, other):
        if isinstance(other, numbers.Number):
            return expr(terms=[(c * other, f) for c, f in self.terms])
        other = expr(other)
        return expr(terms=[(c1 * c2, f1 + f2) for c1, f1 in self.terms for c2, f2 in other.terms])

    def __rmul__(self, other):
        if isinstance(other, numbers.Number):
            return expr(terms=[(c * other, f) for c, f in self.terms])
        return expr(other) * self

    def __add__(self, other):
        return expr(terms=self.terms + expr(other).terms)

    def __sub__(self, other):
        return expr(terms=self.terms + [(-c, f) for c, f in expr(other).terms])

    def __neg__(self):
        return expr(terms=[(-c, f) for c, f in self.terms])

    def __truediv__(self, other):
        return self * (1.0 / other)


def _apply_chain(factors):
    cur = factors[-1]
    if not isinstance(cur, lattice):
        raise TypeError("right-most factor of a product must be a lattice")
    for op in reversed(factors[:-1]):
        if not isinstance(op, matrix_operator):
            raise TypeError("factors left of a lattice must be matrix operators")
        cur = op(cur)
    return cur


def eval(first, second=None, ac=False):  # noqa: A001  (GPT's name)
    """g.eval(expr) -> new lattice ; g.eval(dst, expr[, ac]) -> dst (+)= expr   (expr.py:333-412)"""
    if second is None:
        e, dst = first, None
        if isinstance(e, (lattice, mspincolor)):
            return e
        if isinstance(e, _deferred):
            return e.op(e.arg)
    else:
        dst, e = first, second
    e = expr(e)
    vals = [(c, _apply_chain(f)) for c, f in e.terms]
    if dst is None:
        dst = lattice(vals[0][1])
    # the reference's encoding of a linear combination (lib/gpt/core/expr.py:126-143): [(coefficient, [(factor_unary, [lattice])])]
    terms = [(builtins.complex(c), [(0, [v])]) for c, v in vals]
    cgpt.eval(dst.v_obj, terms, 0, bool(ac), 0)
    return dst


# ---- matrix_operator -------------------------------------------------------------------------------------------------
class vector_space:
    """where the result of an operator lives (lib/gpt/core/vector_space.py, explicit_grid_otype)"""

    def __init__(self, grid_, otype, cb=None):
        self.grid = grid_
        self.otype = otype
        self.cb = cb

    def lattice(self):
        l = lattice(self.grid, self.otype)
        if self.cb is not None and self.grid.cb.n != 1:
            l.checkerboard(self.cb)
        return l

    def clone(self):
        return vector_space(self.grid, self.otype, self.cb)

    def converted(self, precision):
        return vector_space(self.grid.converted(precision), self.otype, self.cb)


class matrix_operator:
    def __init__(self, mat, adj_mat=None, inv_mat=None, adj_inv_mat=None, vector_space=None,
                 accept_guess=(False, False), accept_list=False):
        self.mat = mat
        self.adj_mat = adj_mat
        self.inv_mat = inv_mat
        self.adj_inv_mat = adj_inv_mat
        self.vector_space = vector_space if isinstance(vector_space, tuple) else (vector_space, vector_space)
        self.accept_guess = accept_guess if isinstance(accept_guess, tuple) else (accept_guess, accept_guess)
        self.accept_list = accept_list

    def inv(self):
        return matrix_operator(self.inv_mat, self.adj_inv_mat, self.mat, self.adj_mat,
                               tuple(reversed(self.vector_space)), tuple(reversed(self.accept_guess)), self.accept_list)

    def adj(self):
        return matrix_operator(self.adj_mat, self.mat, self.adj_inv_mat, self.inv_mat,
                               tuple(reversed(self.vector_space)), tuple(reversed(self.accept_guess)), self.accept_list)

    def clone(self):
        vs = tuple(v.clone() if v is not None else None for v in self.vector_space)
        return matrix_operator(self.mat, self.adj_mat, self.inv_mat, self.adj_inv_mat, vs, self.accept_guess,
                               self.accept_list)

    def specialized_singlet_callable(self):
        return self.mat if not self.accept_list else self

    def packed(self):
        """operator on a 5d grid whose fifth dimension enumerates right-hand sides -> operator on lists of n_rhs 4d fields
        (lib/gpt/core/operator/matrix_operator.py:200-239); packing and unpacking are device copies"""
        vs5 = self.vector_space
        grid5 = vs5[1].grid
        assert grid5 is not None and grid5.nd == 5
        n_rhs = grid5.fdimensions[0]

        def _packed(dst, src, op5):
            assert len(src) == n_rhs and len(dst) == n_rhs
            t_src = lattice(grid5, src[0].otype)
            t_dst = lattice(grid5, dst[0].otype)
            if grid5.cb.n != 1:
                t_src.checkerboard(src[0].checkerboard())
            capi.lattice_pack_rhs(t_src.obj, [x.obj for x in src])
            op5(t_dst, t_src)
            capi.lattice_pack_rhs(t_dst.obj, [x.obj for x in dst], unpack=True)

        def wrap(get):
            return lambda dst, src: _packed(dst, src, get())

        grid4 = grid5.removed_dimension(0)
        vs4 = tuple(vector_space(grid4, v.otype, v.cb) if v is not None else None for v in vs5)
        return matrix_operator(
            mat=wrap(lambda: self), adj_mat=wrap(lambda: self.adj()),
            inv_mat=wrap(lambda: self.inv()) if self.inv_mat is not None else None,
            adj_inv_mat=wrap(lambda: self.adj().inv()) if self.adj_inv_mat is not None else None,
            vector_space=vs4, accept_guess=self.accept_guess, accept_list=True)

    def __mul__(self, other):
        if isinstance(other, matrix_operator):
            return matrix_operator_product([self, other])
        return expr(self) * expr(other)

    def __rmul__(self, other):
        return expr(self) * other

    def _new_dst(self, src):
        vs = self.vector_space[0]
        if vs is None:
            dst = lattice(src)
        else:
            dst = vs.lattice()
            if vs.cb is None and dst.grid.cb.n != 1 and src.grid.cb.n != 1:
                dst.checkerboard(src.checkerboard())
        if self.accept_guess[0]:
            dst[:] = 0
        return dst

    def __call__(self, first, second=None):
        """op(src) -> new dst ; op(dst, src) -> dst    (matrix_operator.py:256-305)"""
        if self.mat is None:
            raise NotImplementedError("matrix_operator has no matrix defined for this direction")
        if second is None:
            src = first
            if isinstance(src, expr):
                src = eval(src)
            if isinstance(src, list):
                dst = [self._new_dst(s) for s in src]
            else:
                dst = self._new_dst(src)
        else:
            dst, src = first, second
        if isinstance(src, list):
            if self.accept_list:
                self.mat(dst, src)
            else:
                for d, s in zip(dst, src):
                    self.mat(d, s)
        else:
            if self.accept_list:
                self.mat([dst], [src])
            else:
                self.mat(dst, src)
        return dst


class projected_matrix_operator:
    """lib/gpt/core/operator/projected_matrix_operator.py (gradients projected to the gauge algebra): outside the hot path; the
    class exists so that `isinstance(x, g.projected_matrix_operator)` in user code is False for every operator of this package"""


class matrix_operator_product(matrix_operator):
    def __init__(self, factors):
        self.factors = factors

        def _mat(dst, src):
            cur = src
            for f in reversed(factors[1:]):
                cur = f(cur)
            factors[0](dst, cur)

        super().__init__(_mat, vector_space=(factors[0].vector_space[0], factors[-1].vector_space[1]),
                         accept_guess=(factors[0].accept_guess[0], factors[-1].accept_guess[1]))

    def adj(self):
        return matrix_operator_product([f.adj() for f in reversed(self.factors)])

    def inv(self):
        return matrix_operator_product([f.inv() for f in reversed(self.factors)])

    def __mul__(self, other):
        if isinstance(other, matrix_operator_product):
            return matrix_operator_product(self.factors + other.factors)
        if isinstance(other, matrix_operator):
            return matrix_operator_product(self.factors + [other])
        return expr(self) * expr(other)


# ---- transforms ------------------------------------------------------------------------------------------------------
def copy(first, second=None):
    if second is None:
        dst = lattice(first)
        cgpt.copy(dst.obj, first.obj)
        return dst
    cgpt.copy(first.obj, second.obj)
    return first


def convert(first, second):
    """g.convert(dst, src) or g.convert(src, precision) (also for lists)"""
    if isinstance(first, list):
        return [convert(x, second) for x in first]
    if isinstance(second, _precision):
        src = first
        dst = lattice(src.grid.converted(second), src.otype)
        cgpt.convert(dst.obj, src.obj)
        return dst
    cgpt.convert(first.obj, second.obj)
    return first


def norm2(l):
    if isinstance(l, list):
        return [norm2(x) for x in l]
    l = eval(l)
    return l.grid.globalsum(cgpt.lattice_norm2(l.v_obj[0]))


def inner_product(a, b):
    a, b = eval(a), eval(b)
    return a.grid.globalsum(builtins.complex(cgpt.lattice_rank_inner_product([a], [b], 1, 1)[0, 0]))


def rank_inner_product(a, b, n_block=1, use_accelerator=True):
    """rank-local <a,b>; the caller does the global sum (lib/gpt/core/transform.py:91-98)"""
    a, b = eval(a), eval(b)
    return builtins.complex(cgpt.lattice_rank_inner_product([a], [b], 1, 1)[0, 0])


def inner_product_norm2(a, b):
    ip, n2 = cgpt.lattice_inner_product_norm2(a.obj, b.obj)
    return a.grid.globalsum(ip), a.grid.globalsum(n2)


def separate(x, dimension=0):
    """5d field(s) -> list of the 4d fields of the slices of dimension 0 (gpt.separate, lib/gpt/core/transform.py; device copies)"""
    assert dimension == 0
    if isinstance(x, list):
        return [y for f in x for y in separate(f, dimension)]
    if isinstance(x, expr):
        x = eval(x)
    grid4 = x.grid.removed_dimension(0)
    out = [lattice(grid4, x.otype) for _ in range(x.grid.fdimensions[0])]
    capi.lattice_pack_rhs(x.obj, [o.obj for o in out], unpack=True)
    return out


def merge(lst, dimension=0, N=-1):
    """list of 4d fields -> 5d field(s) with N slices each (all of them for N = -1); gpt.merge"""
    assert dimension == 0
    lst = [eval(y) if isinstance(y, expr) else y for y in lst]
    if N == -1:
        N = len(lst)
    assert len(lst) % N == 0
    out = []
    for i in range(0, len(lst), N):
        grid5 = lst[i].grid.inserted_dimension(0, N)
        l5 = lattice(grid5, lst[i].otype)
        capi.lattice_pack_rhs(l5.obj, [y.obj for y in lst[i:i + N]])
        out.append(l5)
    return out[0] if len(out) == 1 else out


def scale_per_coordinate(d, s, a, dim):
    """d = a[x_dim] * s (lib/gpt/core/transform.py:210-214)"""
    if isinstance(s, expr):
        s = eval(s)
    cgpt.lattice_scale_per_coordinate(d.obj, s.obj, a, dim)


def axpy(d, a, x, y):
    cgpt.lattice_axpy(d.v_obj[0], builtins.complex(a), x.v_obj[0], y.v_obj[0])


def axpy_norm2(d, a, x, y):
    return d.grid.globalsum(capi.lattice_axpy_norm2(d.obj, a, x.obj, y.obj))


def linear_combination(r, basis, Qt, n_block=8):
    rr = r if isinstance(r, list) else [r]
    q = np.atleast_2d(np.asarray(Qt))
    cgpt.linear_combination(rr, basis, q, n_block)
    return r


def pick_checkerboard(cb, dst, src):
    cgpt.lattice_pick_checkerboard(cb.tag, src.v_obj[0], dst.v_obj[0])  # cgpt's (cb, src, dst) order, checkerboard.py:70
    return dst


def set_checkerboard(dst, src):
    cgpt.lattice_set_checkerboard(src.v_obj[0], dst.v_obj[0])  # (src, dst), checkerboard.py:81
    return dst


# ---- logging / timing ------------------------------------------------------------------------------------------------
_t0 = _time.time()


def time():  # noqa: A001
    """seconds since start-up.  Every cgpt call of the reference returns with its work done; this library enqueues, so g.time()
    drains the stream first -- the `t0 = g.time(); ...; t1 = g.time()` idiom of the benchmarks keeps measuring the work."""
    if capi._initialized:
        capi.accelerator_barrier()
    return _time.time() - _t0


def message(*a):
    import os

    if int(os.environ.get("RANK", "0")) == 0:
        print("GPT_B200 : %14.6f s :" % time(), *a)
        sys.stdout.flush()


# ---- spin-colour matrix fields (propagators) as 12 spin-colour vector columns ----------------------------------------
class mspincolor:
    """
    g.mspincolor(grid): a 12x12 spin-colour matrix field held as its 12 columns (spin-colour vectors) -- the
    decomposition GPT applies before a solver sees a propagator source
    (lib/gpt/core/object_type/container.py:330-357, core/operator/matrix_operator.py:290-295).
    """

    def __init__(self, grid_):
        self.grid = grid_
        self.columns = [vspincolor(grid_) for _ in range(12)]
        for c in self.columns:
            c[:] = 0

    def __getitem__(self, key):
        if isinstance(key, builtins.slice) and key == builtins.slice(None):
            # [site, spin_i, spin_j, color_a, color_b] like GPT's mspincolor tensor
            cols = np.stack([c[:] for c in self.columns], axis=-1)  # [site, 4, 3, 12]
            n = cols.shape[0]
            return cols.reshape(n, 4, 3, 4, 3).transpose(0, 1, 3, 2, 4)
        raise NotImplementedError

    def __mul__(self, other):
        if isinstance(other, _adjoint) and isinstance(other.x, mspincolor):
            return _ab_dagger(self, other.x)
        return NotImplemented


class _adjoint:
    def __init__(self, x):
        self.x = x


class _ab_dagger:
    def __init__(self, a, b):
        self.a, self.b = a, b


class _trace_ab_dagger:
    def __init__(self, a, b):
        self.a, self.b = a, b


def adj(x):
    if isinstance(x, matrix_operator):
        return x.adj()
    return _adjoint(x)


def inv(x):
    return x.inv()


def trace(x):
    if isinstance(x, _ab_dagger):
        return _trace_ab_dagger(x.a, x.b)
    raise NotImplementedError("g.trace is implemented for prop * g.adj(prop)")


def slice(x, dim):  # noqa: A001  (GPT's name)
    """g.slice(g.trace(a * g.adj(b)), 3) -> list of complex per time slice (lib/gpt/core/transform.py:170-171)"""
    if isinstance(x, _trace_ab_dagger) and dim == 3:
        from gpt_b200 import parallel

        gr = x.a.grid
        nt, lt = gr.fdimensions[-1], gr.ldimensions[-1]
        t0 = parallel.processor_coor(parallel.rank, parallel.mpi)[3] * lt  # first global time slice of this rank's block
        acc = np.zeros(nt, dtype=np.complex128)
        for ca, cb_ in zip(x.a.columns, x.b.columns):
            acc[t0:t0 + lt] += capi.lattice_slice_inner_product(cb_.obj, ca.obj, lt)
        acc = np.asarray(gr.globalsum(acc))  # ranks split in y / z hold partial sums of the same slices
        return [builtins.complex(v) for v in acc]
    raise NotImplementedError("g.slice is implemented for g.trace(a * g.adj(b)) along time")


class _create:
    @staticmethod
    def point(src, pos):
        """g.create.point(src, pos): unit spin-colour matrix at `pos`, zero elsewhere (lib/gpt/create/point.py)"""
        assert isinstance(src, mspincolor)
        gr = src.grid
        idx = _local_site(gr, pos)  # None on the ranks that do not own the site
        for j, col in enumerate(src.columns):
            a = np.zeros((gr.lsites, 4, 3), dtype=gr.precision.complex_dtype)
            if idx is not None:
                a.reshape(gr.lsites, 12)[idx, j] = 1.0
            col[:] = a
        return src


create = _create()


class propagator_operator(matrix_operator):
    """a spin-colour-vector operator that also accepts mspincolor sources column by column"""

    def __init__(self, op):
        self.op = op
        super().__init__(op.mat, op.adj_mat, op.inv_mat, op.adj_inv_mat, op.vector_space, op.accept_guess, op.accept_list)

    def __call__(self, first, second=None):
        src = first if second is None else second
        if isinstance(src, mspincolor):
            dst = mspincolor(self.vector_space[0].grid) if second is None else first
            for d, s in zip(dst.columns, src.columns):
                matrix_operator.__call__(self, d, s)
            return dst
        return matrix_operator.__call__(self, first, second)

    def __mul__(self, other):
        if isinstance(other, mspincolor):
            return _deferred(self, other)
        return matrix_operator.__mul__(self, other)

    def adj(self):
        return propagator_operator(self.op.adj())


class _deferred:
    def __init__(self, op, arg):
        self.op, self.arg = op, arg

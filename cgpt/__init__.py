"""
cgpt -- drop-in stand-in for GPT's CPython extension module `cgpt` on the fermion-operator hot path.

The reference's Python layer (lib/gpt) talks to its C++/Grid back end exclusively through a module called `cgpt`
(/root/reference/lib/cgpt/lib/lib.cc:22-56).  This module exports the entry points of that boundary which the hot path uses,
with the reference's names, argument lists and return values (the PyArg_ParseTuple format of each export is quoted next to the
function), and implements them on libcgpt_b200.so -- the CUDA library behind the C ABI of include/cgpt_b200.h -- through the
ctypes binding gpt_b200/cgpt.py.  A reference checkout that puts this directory before its own build of cgpt on PYTHONPATH
gets its fermion operators (lib/gpt/qcd/fermion/operator/interface.py runs unchanged), inner products, axpy, linear
combinations and random fields from the B200 library; INTEGRATION.md lists what is and is not covered.

Handles are plain Python ints like the reference's (PyLong_FromVoidPtr): lattices and operators are the C library's pointers,
grids are small integers (the C library has no grid objects: a lattice carries its own geometry).  Errors raise RuntimeError with
the library's message (reference: lib/cgpt/lib/exception.h:23-39).  There is no CPU fallback: without the shared library or
without a CUDA device every compute entry point raises.

Deliberate differences from the reference (storage is owned by this library, not by Grid):
  * lattice_memory_view / copy plans are not provided; `lattice[:]` goes through lattice_import / lattice_export (GPT order);
  * eval() evaluates linear combinations of lattices of one type (what the CG and the solver stack need), not the general
    tensor-expression engine (lib/cgpt/lib/eval.cc, expression/*.h).
"""
import sys as _sys
import time as _time

import numpy as _np

from gpt_b200 import cgpt as _capi

# ---- runtime ---------------------------------------------------------------------------------------------------------------
_mpi_default = [1, 1, 1, 1]


def init(argv):
    """cgpt.init(sys.argv)  "O"  (lib/cgpt/lib/init.cc:26-103): Grid_init there, device selection here (LOCAL_RANK under torchrun)"""
    _capi.init()
    return 0


def exit():  # noqa: A001
    """cgpt.exit()  ""  (lib/cgpt/lib/init.cc:105-117)"""
    _capi.comm_finalize()
    return 0


def time():
    """cgpt.time() -> float seconds  ""  (lib/cgpt/lib/time.cc:79-81)"""
    return _time.time()


def accelerator_barrier():
    """cgpt.accelerator_barrier()  ""  (lib/cgpt/lib/util.cc:335-338)"""
    _capi.accelerator_barrier()
    return 0


def global_rank():
    """(lib/cgpt/lib/grid.cc, gpt/core/mpi.py:26)"""
    from gpt_b200 import parallel

    return parallel.rank


def global_ranks():
    from gpt_b200 import parallel

    return parallel.world


def barrier():
    _capi.accelerator_barrier()
    return 0


# ---- grids -----------------------------------------------------------------------------------------------------------------
class _grid_info:
    __slots__ = ("fdimensions", "precision", "cb_mask", "mpi", "ldimensions", "processor_coor", "ref")


_grids = {}
_grid_by_tag = {}


def _grid(h):
    try:
        return _grids[h]
    except KeyError:
        raise RuntimeError(f"cgpt: {h} is not a grid handle")


def create_grid(fdimensions, precision, cb_mask, simd_mask, mpi, parent_grid):
    """cgpt.create_grid(fdimensions, precision, cb_mask, simd_mask, mpi, parent) -> handle  "OOOOOl"  (lib/cgpt/lib/grid.cc:24-85).
    Grids are interned by (fdimensions, precision, cb_mask, mpi) like cgpt_grid_cache (lib/cgpt/lib/grid.h:33-82): equal grids
    share one handle, which is what keys the parallel random number generators."""
    from gpt_b200 import parallel

    if precision not in ("single", "double"):
        raise RuntimeError("Unknown precision")
    nd = len(fdimensions)
    if not (len(cb_mask) == nd and len(simd_mask) == nd and len(mpi) == nd):
        raise RuntimeError("Assert failed: nd == fdimensions.size()")
    if parent_grid:
        raise RuntimeError("cgpt_b200: split grids (a parent grid) are not supported")
    if nd not in (4, 5):
        raise RuntimeError("cgpt_b200: only 4d grids [x,y,z,t] and 5d grids [s,x,y,z,t] are supported")
    tag = (tuple(int(x) for x in fdimensions), precision, tuple(int(x) for x in cb_mask), tuple(int(x) for x in mpi))
    if tag in _grid_by_tag:
        h = _grid_by_tag[tag]
        _grids[h].ref += 1
        return h
    g = _grid_info()
    g.fdimensions, g.precision, g.cb_mask, g.mpi = list(tag[0]), precision, list(tag[2]), list(tag[3])
    mpi4 = g.mpi[-4:]
    if parallel.active and mpi4 != list(parallel.mpi):
        raise RuntimeError(f"cgpt_b200: grid mpi layout {mpi4} differs from the processor grid {parallel.mpi} set up at start")
    for d in range(nd):
        cb_factor = 2 if g.cb_mask[d] else 1
        if g.fdimensions[d] % (g.mpi[d] * cb_factor) != 0:
            raise RuntimeError(f"Dimension {d} is not consistent:\n fdimension = {g.fdimensions[d]}\n mpi = {g.mpi[d]}\n cb = {cb_factor}\n")
    g.ldimensions = [f // m for f, m in zip(g.fdimensions, g.mpi)]
    coor4 = parallel.processor_coor(parallel.rank, mpi4)
    g.processor_coor = ([0] if nd == 5 else []) + list(coor4)
    g.ref = 1
    h = len(_grids) + 1
    _grids[h] = g
    _grid_by_tag[tag] = h
    return h


def delete_grid(grid):
    """cgpt.delete_grid(grid)  "l"  (lib/cgpt/lib/grid.cc:89-99): grids stay interned (the reference keeps them cached as well)"""
    _grid(grid).ref -= 1
    return 0


def grid_get_processor(grid):
    """-> (rank, ranks, processor_coor, gdimensions, ldimensions, srank, sranks)  "l"  (lib/cgpt/lib/grid.cc:162-180)"""
    from gpt_b200 import parallel

    g = _grid(grid)
    # gdimensions of a checkerboarded grid: the checkerboarded dimension is halved (Grid's _gdimensions)
    gdims = [f // (2 if c else 1) for f, c in zip(g.fdimensions, g.cb_mask)]
    ldims = [l // (2 if c else 1) for l, c in zip(g.ldimensions, g.cb_mask)]
    return (parallel.rank, parallel.world, list(g.processor_coor), gdims, ldims, 0, 1)


def grid_get_simd(grid):
    """SIMT threads take the place of SIMD lanes: the layout is trivial"""
    return [1] * len(_grid(grid).fdimensions)


def grid_barrier(grid):
    _capi.accelerator_barrier()
    return 0


def grid_globalsum(grid, x):
    """cgpt.grid_globalsum(grid, x)  "lO"  (lib/cgpt/lib/grid.cc:119-160): complex / float / int -> new value; ndarray in place"""
    from gpt_b200 import parallel

    _grid(grid)
    if isinstance(x, _np.ndarray):
        if parallel.active:
            flat = x.reshape(-1)
            if _np.iscomplexobj(flat):
                r = _capi.comm_globalsum(flat.astype(_np.complex128).view(_np.float64)).view(_np.complex128)
            else:
                r = _capi.comm_globalsum(flat.astype(_np.float64))
            flat[:] = r.astype(x.dtype)
        return x
    if isinstance(x, (complex, float, int)):
        r = parallel.globalsum(x)
        return type(x)(r) if not isinstance(x, complex) else r
    raise RuntimeError("Unsupported object")


# ---- lattices ----------------------------------------------------------------------------------------------------------------
# v_otype strings of lib/gpt/core/object_type -> complex components per site
_otype_components = {
    "ot_singlet": 1, "ot_vector_color(3)": 3, "ot_vector_spin(4)": 4, "ot_matrix_color(3)": 9, "ot_vector_spin_color(4,3)": 12,
    "ot_matrix_spin(4)": 16, "ot_matrix_spin_color(4,3)": 144,
}
_prec = {"single": _capi.SINGLE, "double": _capi.DOUBLE}


def _handle(x, i=0):
    """a lattice handle, or the i-th handle of a gpt.lattice object (cgpt_basis_fill, lib/cgpt/lib/lattice/basis.h)"""
    return x.v_obj[i] if hasattr(x, "v_obj") else int(x)


def _handles(objs):
    out = []
    for x in objs:
        out.extend(x.v_obj if hasattr(x, "v_obj") else [int(x)])
    return out


def create_lattice(grid, otype, prec):
    """cgpt.create_lattice(grid, otype, prec) -> handle  "lOO"  (lib/cgpt/lib/lattice.cc:42-64)"""
    g = _grid(grid)
    if otype not in _otype_components or prec not in _prec:
        raise RuntimeError(f"Unknown field type: {otype}, {prec}")
    nd = len(g.fdimensions)
    cb = _capi.EVEN if any(g.cb_mask) else _capi.FULL
    return _capi.create_lattice(g.ldimensions[-4:], g.fdimensions[0] if nd == 5 else 0, _prec[prec], _otype_components[otype], cb)


def delete_lattice(lattice):
    """"l"  (lattice.cc:66-75)"""
    _capi.delete_lattice(lattice)
    return 0


def lattice_set_to_number(lattice, number):
    """cgpt.lattice_set_to_number(l, number)  "lO"  (lattice.cc:86-99): `lattice[:] = 0` is the case the hot path uses"""
    if complex(number) != 0:
        raise RuntimeError("cgpt_b200: lattice_set_to_number supports 0 (use lattice_import for other values)")
    _capi.lattice_set_to_zero(lattice)
    return 0


def lattice_get_checkerboard(lattice):
    """-> 0 (even) / 1 (odd)  "l"  (lattice.cc:189-198)"""
    return _capi.lattice_get_checkerboard(lattice)


def lattice_change_checkerboard(lattice, cb):
    """"ll"  (lattice.cc:200-210)"""
    _capi.lattice_change_checkerboard(lattice, int(cb))
    return 0


def lattice_pick_checkerboard(cb, src, dst):
    """cgpt.lattice_pick_checkerboard(cb, src(full), dst(half))  "lll"  (lattice.cc:164-175; gpt/core/checkerboard.py:70)"""
    _capi.lattice_pick_checkerboard(int(cb), dst, src)
    return 0


def lattice_set_checkerboard(src, dst):
    """cgpt.lattice_set_checkerboard(src(half), dst(full))  "ll"  (lattice.cc:177-187; gpt/core/checkerboard.py:81)"""
    _capi.lattice_set_checkerboard(dst, src)
    return 0


def copy(dst, src):
    """"ll"  (lib/cgpt/lib/transform.cc:41-54)"""
    _capi.copy(dst, src)
    return 0


def convert(dst, src):
    """"ll"  (transform.cc:128-141)"""
    _capi.convert(dst, src)
    return 0


def lattice_import(lattice, array):
    """replacement for the write direction of lattice_memory_view + copy plan: `lattice[:] = array`, array in GPT order
    (site lexicographic with dimension 0 fastest, tensor row-major; lib/cgpt/lib/lattice/implementation.h:246-281)"""
    _capi.lattice_import(lattice, array)
    return 0


def lattice_export(lattice, array):
    """`array[...] = lattice[:]`, see lattice_import"""
    _capi.lattice_export(lattice, array)
    return array


# ---- vector kernels ------------------------------------------------------------------------------------------------------------
def lattice_axpy(r, a, x, y):
    """cgpt.lattice_axpy(r, a, x, y): r = a x + y  "lOll"  (transform.cc:229-246); no barrier, like accelerator_forNB"""
    _capi.lattice_axpy(r, a, x, y)
    return 0


def lattice_rank_inner_product(left, right, n_block, use_accelerator):
    """cgpt.lattice_rank_inner_product(left, right, n_block, use_accelerator) -> ndarray complex128  "OOll"  (transform.cc:143-178):
    left / right are lists of gpt.lattice objects; shape [n_left, n_right] for n_block = 1, else [n_block, n_left/n_block,
    n_right/n_block] (block i pairs the i-th groups).  Rank local, accumulated in double (foundation/reduce.h:129).
    use_accelerator is accepted and ignored: there is no host path."""
    n_block = int(n_block)
    lv, rv = len(left[0].v_obj) if hasattr(left[0], "v_obj") else 1, len(right[0].v_obj) if hasattr(right[0], "v_obj") else 1
    if lv != rv or lv != 1:
        raise RuntimeError("Assert failed: n_virtual_left == n_virtual_right (cgpt_b200: one lattice per object)")
    L, R = _handles(left), _handles(right)
    if n_block == 1:
        return _capi.lattice_rank_inner_product(L, R)
    if len(L) % n_block or len(R) % n_block:
        raise RuntimeError("Assert failed: left.size() % (n_virtual_left * n_block) == 0")
    nl, nr = len(L) // n_block, len(R) // n_block
    out = _np.empty((n_block, nl, nr), dtype=_np.complex128)
    for b in range(n_block):
        out[b] = _capi.lattice_rank_inner_product(L[b * nl:(b + 1) * nl], R[b * nr:(b + 1) * nr])
    return out


def lattice_inner_product_norm2(a, b):
    """-> (<a,b>, |a|^2)  "ll"  (transform.cc:180-197)"""
    return _capi.lattice_inner_product_norm2(a, b)


def lattice_norm2(a):
    """"l"  (transform.cc:199-209)"""
    return _capi.lattice_norm2(a)


def lattice_scale_per_coordinate(d, s, a, dim):
    """"llOl"  (lib/cgpt/lib/lattice.cc; gpt/core/transform.py:210-214)"""
    _capi.lattice_scale_per_coordinate(d, s, a, dim)
    return 0


def linear_combination(r, basis, Qt, n_block):
    """cgpt.linear_combination(r, basis, Qt, n_block): r_i = sum_k Qt[i,k] basis_k  "OOOl"  (lib/cgpt/lib/basis.cc:146-173)"""
    _capi.linear_combination(_handles(r), _handles(basis), _np.asarray(Qt, dtype=_np.complex128))
    return 0


FACTOR_UNARY_NONE = 0  # lib/gpt/core/expr.py:27-31


def eval(dst, terms, unary, ac, idx):  # noqa: A001
    """cgpt.eval(dst | None, terms, unary, ac, idx)  "OOiOi"  (lib/cgpt/lib/eval.cc:323-365).
    terms = [(coefficient, [(factor_unary, [lattice objects])])]  (lib/gpt/core/expr.py:126-143); idx selects the idx-th
    lattice of every factor list.  Supported: sums of coefficient * lattice (no unary operators, one factor per term) -- the
    expressions of cg.py, defect_correcting.py, schur_complement_two.py.  Returns the list of destination handles for a given
    dst, or [(handle, otype, precision)] declarations for dst = None like cgpt_Lattice_base::to_decl."""
    if not isinstance(ac, bool):
        raise RuntimeError("Assert failed: PyBool_Check(_ac)")
    if unary != 0:
        raise RuntimeError("cgpt_b200.eval: traces are not part of the hot path")
    coefs, lats = [], []
    for coef, factors in terms:
        if len(factors) != 1 or factors[0][0] != FACTOR_UNARY_NONE:
            raise RuntimeError("cgpt_b200.eval: only linear combinations of lattices are supported")
        f = factors[0][1]
        f = f[idx] if isinstance(f, (list, tuple)) else f
        if not hasattr(f, "v_obj") and not isinstance(f, int):
            raise RuntimeError("cgpt_b200.eval: factors must be lattices")
        coefs.append(complex(coef))
        lats.append(f)
    if not lats:
        raise RuntimeError("cgpt_b200.eval: empty expression")
    nv = len(lats[0].v_obj) if hasattr(lats[0], "v_obj") else 1
    if dst is None:
        first = lats[0]
        out, decl = [], []
        for i in range(nv):
            h = _capi.create_lattice_like(_handle(first, i))
            _capi.lattice_lc(h, False, coefs, [_handle(x, i) for x in lats])
            out.append(h)
            decl.append((h, first.otype.v_otype[i] if hasattr(first, "otype") and hasattr(first.otype, "v_otype") else None,
                         first.grid.precision.cgpt_dtype if hasattr(first, "grid") else None))
        return decl
    dst = list(dst)
    for i, h in enumerate(dst):
        _capi.lattice_lc(h, ac, coefs, [_handle(x, i) for x in lats])
    return dst


# ---- random numbers ----------------------------------------------------------------------------------------------------------------
def create_random(engine, seed):
    """cgpt.create_random(engine, seed) -> handle  "OO"  (lib/cgpt/lib/random.cc:38-62)"""
    return _capi.create_random(engine, seed)


def delete_random(rng):
    """"l"  (random.cc:64-73)"""
    _capi.delete_random(rng)
    return 0


def random_sample(rng, params):
    """cgpt.random_sample(rng, params)  "lO"  (random.cc:75-101, random/engine.h:64-125): params = {"distribution": ..., mu / sigma |
    min / max | n, optionally "lattices": [gpt.lattice, ...]}; without lattices a scalar is drawn from the engine's own stream.
    Lattices are filled from the per-grid parallel generators (one per 2^4 block, keyed by the interned grid)."""
    lattices = params.get("lattices") if hasattr(params, "get") else None
    if not lattices:
        return _capi.random_sample_scalar(rng, params)
    for lat in lattices:
        key = lat.grid.obj if isinstance(lat.grid.obj, int) else lat.grid.serial
        for h in lat.v_obj:
            _capi.random_sample(rng, key, h, params)
    return 0


# ---- fermion operators -----------------------------------------------------------------------------------------------------------------
def create_fermion_operator(optype, prec, params):
    """cgpt.create_fermion_operator(optype, prec, params) -> handle  "OOO"  (lib/cgpt/lib/operators.cc:34-57): optype in
    {"wilson_clover", "wilson_twisted_mass", "mobius", "zmobius"}; params carries "U" (four lattice handles), the physics
    parameters of lib/cgpt/lib/operators/{wilson_clover,mobius,zmobius}.h and the grid handles U_grid / F_grid ... (unused here:
    lattices know their geometry)"""
    if prec not in _prec:
        raise RuntimeError(f"Unknown precision {prec}")
    return _capi.create_fermion_operator(optype, prec, params)


def update_fermion_operator(op, params):
    """"lO"  (operators.cc:59-70): re-import the gauge field params["U"]"""
    _capi.update_fermion_operator(op, params)
    return 0


def set_mass_fermion_operator(op, params):
    """"lO"  (operators.cc:72-83)"""
    _capi.set_mass_fermion_operator(op, params)
    return 0


def delete_fermion_operator(op):
    """"l"  (operators.cc:85-94)"""
    _capi.delete_fermion_operator(op)
    return 0


def apply_fermion_operator(op, opcode, src, dst):
    """cgpt.apply_fermion_operator(op, opcode, src, dst) -> 0.0  "llOO"  (operators.cc:96-107): src and dst are the v_obj LISTS of
    the two lattices, in Grid's (in, out) order (lib/gpt/qcd/fermion/operator/interface.py:88-91); opcodes of
    lib/cgpt/lib/operators/register.h:2-20"""
    if len(src) != 1 or len(dst) != 1:
        raise RuntimeError("Assert failed: src.size() == 1 && dst.size() == 1")
    return _capi.apply_fermion_operator(op, opcode, src[0], dst[0])


# ---- generic matrix-vector stencils ----------------------------------------------------------------------------------------
def stencil_matrix_vector_create(lattice_matrix, lattice_vector, grid, shifts, code, code_parallel_block_size, local, matrix_parity,
                                 vector_parity):
    """cgpt.stencil_matrix_vector_create(lattice_matrix, lattice_vector, grid, shifts, code, code_parallel_block_size, local,
    matrix_parity, vector_parity) -> handle  "lllOOllll"  (lib/cgpt/lib/stencil.cc:41-58).  code: list of dicts with the keys
    target, accumulate, source, source_point, weight, factor = [(field index, point, adjoint)]
    (lib/cgpt/lib/lattice/implementation.h stencil_matrix_vector, lib/gpt/core/local_stencil/matrix_vector.py:22-34)"""
    g = _grid(grid)
    if len(g.fdimensions) != 4:
        raise RuntimeError("cgpt_b200: matrix-vector stencils live on 4d grids")
    return _capi.stencil_matrix_vector_create(g.ldimensions, g.precision, [tuple(int(x) for x in p) for p in shifts], code,
                                              code_parallel_block_size, local, matrix_parity, vector_parity)


def stencil_matrix_vector_execute(stencil, matrix_fields, vector_fields, fast_osites):
    """"lOOl"  (lib/cgpt/lib/stencil.cc:101-121): the two lists hold gpt lattices (their first v_obj is used, cgpt_basis_fill)"""
    _capi.stencil_matrix_vector_execute(stencil, [_handle(x) for x in matrix_fields], [_handle(x) for x in vector_fields], fast_osites)
    return 0


def stencil_matrix_vector_delete(stencil):
    """"l"  (lib/cgpt/lib/stencil.cc)"""
    _capi.stencil_matrix_vector_delete(stencil)
    return 0

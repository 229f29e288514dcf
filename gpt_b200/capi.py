"""
ctypes binding of libcgpt_b200.so -- the stand-in for GPT's CPython module `cgpt`
(/root/reference/lib/cgpt/lib/lib.cc:22-56).  Function names and argument meaning follow the cgpt exports
they replace; handles are plain integers (cgpt: PyLong_FromVoidPtr, lib/cgpt/lib/operators.cc:56); a failing
call raises RuntimeError with the library's message (cgpt: lib/cgpt/lib/exception.h:23-39).

There is deliberately NO fallback: if the shared library is missing or no CUDA device is present every
entry point raises.
"""
import ctypes
import os

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
LIBRARY_PATH = os.environ.get("GPT_B200_LIBRARY", os.path.join(_here, "lib", "libcgpt_b200.so"))  # override: A/B builds

SINGLE, DOUBLE = 0, 1
EVEN, ODD, FULL = 0, 1, 2
OT_SINGLET, OT_MCOLOR, OT_VSPINCOLOR = 1, 9, 12
WILSON_CLOVER, MOBIUS = 0, 1

c_void_p, c_int, c_double, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
_pp = ctypes.POINTER(c_void_p)
_pd = ctypes.POINTER(c_double)
_pi = ctypes.POINTER(c_int)


class fermion_params(ctypes.Structure):
    _fields_ = [
        ("mass", c_double), ("csw_r", c_double), ("csw_t", c_double), ("cF", c_double), ("xi_0", c_double),
        ("nu", c_double), ("isAnisotropic", c_int),
        ("mass_plus", c_double), ("mass_minus", c_double), ("M5", c_double), ("b", c_double), ("c", c_double),
        ("Ls", c_int),
        ("boundary_phases", c_double * 8),
        ("mu", c_double),
        ("n_omega", c_int),
        ("omega", c_double * 128),
        ("link_compression", c_int),
    ]


# name -> (restype, argtypes); every symbol declared in include/cgpt_b200.h
SIGNATURES = {
    "cgptb_init": (c_int, [c_int]),
    "cgptb_last_error": (ctypes.c_char_p, []),
    "cgptb_accelerator_barrier": (c_int, []),
    "cgptb_set_stream": (c_int, [c_void_p]),
    "cgptb_get_stream": (c_void_p, []),
    "cgptb_timer_start": (c_int, []),
    "cgptb_timer_stop": (c_int, [_pd]),
    "cgptb_device_info": (c_int, [_pi, ctypes.POINTER(c_size_t), _pi, _pi]),
    "cgptb_launch_count": (ctypes.c_uint64, []),
    "cgptb_comm_unique_id": (c_int, [ctypes.c_char_p]),
    "cgptb_comm_init": (c_int, [c_int, c_int, _pi, ctypes.c_char_p]),
    "cgptb_comm_finalize": (c_int, []),
    "cgptb_comm_info": (c_int, [_pi, _pi, _pi, _pi]),
    "cgptb_comm_globalsum": (c_int, [_pd, c_int]),
    "cgptb_lattice_spin_matrix": (c_int, [c_void_p, c_void_p, _pd]),
    "cgptb_lattice_scale_per_coordinate": (c_int, [c_void_p, c_void_p, _pd, c_int, c_int]),
    "cgptb_lattice_pack_rhs": (c_int, [c_void_p, ctypes.POINTER(c_void_p), c_int, c_int]),
    "cgptb_gauge_plaquette": (c_int, [ctypes.POINTER(c_void_p), _pd]),
    "cgptb_nersc_munge": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_uint)]),
    "cgptb_create_random": (c_int, [ctypes.POINTER(c_void_p), ctypes.c_char_p, ctypes.c_char_p]),
    "cgptb_delete_random": (c_int, [c_void_p]),
    "cgptb_random_sample_scalar": (c_int, [c_void_p, c_int, c_double, c_double, _pd]),
    "cgptb_random_sample_host": (c_int, [c_void_p, ctypes.c_uint64, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                         ctypes.POINTER(c_int), c_int, c_int, c_double, c_double, _pd]),
    "cgptb_random_sample": (c_int, [c_void_p, ctypes.c_uint64, c_void_p, c_int, c_double, c_double]),
    "cgptb_random_su3_links": (c_int, [c_void_p, ctypes.c_uint64, ctypes.POINTER(c_void_p), c_double]),
    "cgptb_create_lattice": (c_int, [_pp, _pi, c_int, c_int, c_int, c_int]),
    "cgptb_create_lattice_view": (c_int, [_pp, _pi, c_int, c_int, c_int, c_int, c_void_p]),
    "cgptb_delete_lattice": (c_int, [c_void_p]),
    "cgptb_lattice_bytes": (c_size_t, [c_void_p]),
    "cgptb_lattice_sites": (c_size_t, [c_void_p]),
    "cgptb_lattice_device_ptr": (c_void_p, [c_void_p]),
    "cgptb_lattice_info": (c_int, [c_void_p, _pi, _pi, _pi, _pi, _pi]),
    "cgptb_lattice_get_checkerboard": (c_int, [c_void_p]),
    "cgptb_lattice_change_checkerboard": (c_int, [c_void_p, c_int]),
    "cgptb_lattice_set_to_zero": (c_int, [c_void_p]),
    "cgptb_lattice_import": (c_int, [c_void_p, c_void_p, c_size_t]),
    "cgptb_lattice_export": (c_int, [c_void_p, c_void_p, c_size_t]),
    "cgptb_lattice_import_device": (c_int, [c_void_p, c_void_p, c_size_t]),
    "cgptb_lattice_export_device": (c_int, [c_void_p, c_void_p, c_size_t]),
    "cgptb_lattice_copy": (c_int, [c_void_p, c_void_p]),
    "cgptb_lattice_convert": (c_int, [c_void_p, c_void_p]),
    "cgptb_lattice_pick_checkerboard": (c_int, [c_int, c_void_p, c_void_p]),
    "cgptb_lattice_set_checkerboard": (c_int, [c_void_p, c_void_p]),
    "cgptb_lattice_axpy": (c_int, [c_void_p, c_double, c_double, c_void_p, c_void_p]),
    "cgptb_lattice_axpy_norm2": (c_int, [c_void_p, c_double, c_double, c_void_p, c_void_p, _pd]),
    "cgptb_lattice_rank_inner_product": (c_int, [_pp, c_int, _pp, c_int, _pd]),
    "cgptb_lattice_norm2": (c_int, [c_void_p, _pd]),
    "cgptb_lattice_inner_product_norm2": (c_int, [c_void_p, c_void_p, _pd, _pd]),
    "cgptb_lattice_lc": (c_int, [c_void_p, c_int, c_int, _pd, _pp]),
    "cgptb_linear_combination": (c_int, [_pp, c_int, _pp, c_int, _pd]),
    "cgptb_lattice_scale": (c_int, [c_void_p, c_double, c_double]),
    "cgptb_lattice_slice_inner_product": (c_int, [c_void_p, c_void_p, _pd]),
    "cgptb_debug_tma_schedule": (c_int, [_pi, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _pi, c_int, _pi]),
    "cgptb_stencil_matrix_vector_create": (c_int, [_pp, _pi, c_int, c_int, _pi, c_int, _pi, _pd, _pi, c_int, c_int, c_int, c_int]),
    "cgptb_stencil_matrix_vector_execute": (c_int, [c_void_p, _pp, c_int, _pp, c_int, c_int]),
    "cgptb_stencil_matrix_vector_delete": (c_int, [c_void_p]),
    "cgptb_create_fermion_operator": (c_int, [_pp, c_int, c_int, ctypes.POINTER(fermion_params), _pp]),
    "cgptb_update_fermion_operator": (c_int, [c_void_p, _pp]),
    "cgptb_set_mass_fermion_operator": (c_int, [c_void_p, ctypes.POINTER(fermion_params)]),
    "cgptb_delete_fermion_operator": (c_int, [c_void_p]),
    "cgptb_apply_fermion_operator": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "cgptb_apply_fermion_operator_host": (c_int, [c_void_p, c_int, c_void_p, c_void_p, ctypes.c_size_t]),
    "cgptb_apply_schur_two": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "cgptb_cg_eo2_ne": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_int, _pd, _pi, _pi]),
}

_lib = None
_initialized = False


def library():
    """dlopen the C-ABI library and attach the signatures; no GPU needed for this step."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBRARY_PATH):
            raise RuntimeError(
                f"{LIBRARY_PATH} not found: build it with `python -c 'import __graft_entry__ as e; e.build()'` "
                "(gpt_b200 has no CPU fallback)"
            )
        lib = ctypes.CDLL(LIBRARY_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
        _lib = lib
    return _lib


def _check(status):
    if status != 0:
        raise RuntimeError(library().cgptb_last_error().decode())


def init(device=None):
    """cgpt.init: pick the CUDA device (LOCAL_RANK under torchrun) and create the library stream."""
    global _initialized
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    _check(library().cgptb_init(int(device)))
    _initialized = True


def _lib_ready():
    if not _initialized:
        init()
    return _lib


def _handles(objs):
    return (c_void_p * len(objs))(*objs)


# ---- runtime -----------------------------------------------------------------------------------------
def accelerator_barrier():
    _check(_lib_ready().cgptb_accelerator_barrier())


def set_stream(cuda_stream):
    _check(_lib_ready().cgptb_set_stream(c_void_p(cuda_stream)))


def get_stream():
    """the CUDA stream the library launches on (a cudaStream_t as an integer), e.g. for torch.cuda.ExternalStream"""
    return int(_lib_ready().cgptb_get_stream() or 0)


def timer_start():
    _check(_lib_ready().cgptb_timer_start())


def timer_stop():
    ms = c_double()
    _check(_lib_ready().cgptb_timer_stop(ctypes.byref(ms)))
    return ms.value


def device_info():
    sms, mem, ma, mi = c_int(), c_size_t(), c_int(), c_int()
    _check(_lib_ready().cgptb_device_info(ctypes.byref(sms), ctypes.byref(mem), ctypes.byref(ma), ctypes.byref(mi)))
    return {"sm_count": sms.value, "total_mem": mem.value, "cc": (ma.value, mi.value)}


def launch_count():
    return int(_lib_ready().cgptb_launch_count())


# ---- processor grid ------------------------------------------------------------------------------------
def comm_unique_id():
    buf = ctypes.create_string_buffer(128)
    _check(_lib_ready().cgptb_comm_unique_id(buf))
    return buf.raw


def comm_init(rank, world, mpi, unique_id):
    m = (c_int * 4)(*[int(x) for x in mpi])
    _check(_lib_ready().cgptb_comm_init(int(rank), int(world), m, ctypes.create_string_buffer(unique_id, 128)))


def comm_finalize():
    if _lib is not None:
        _lib.cgptb_comm_finalize()


def comm_globalsum(values):
    a = np.ascontiguousarray(np.array(values, dtype=np.float64).reshape(-1))
    _check(_lib_ready().cgptb_comm_globalsum(a.ctypes.data_as(_pd), a.size))
    return a


# ---- lattices ----------------------------------------------------------------------------------------
def create_lattice(dims4, Ls, precision, otype, cb, device_ptr=None):
    h = c_void_p()
    d = (c_int * 4)(*[int(x) for x in dims4])
    if device_ptr is None:
        _check(_lib_ready().cgptb_create_lattice(ctypes.byref(h), d, int(Ls), precision, otype, cb))
    else:
        _check(_lib_ready().cgptb_create_lattice_view(ctypes.byref(h), d, int(Ls), precision, otype, cb, c_void_p(device_ptr)))
    return h.value


def lattice_info(h):
    d = (c_int * 4)()
    ls, prec, otype, cb = c_int(), c_int(), c_int(), c_int()
    _check(_lib_ready().cgptb_lattice_info(c_void_p(h), d, ctypes.byref(ls), ctypes.byref(prec), ctypes.byref(otype), ctypes.byref(cb)))
    return {"dims4": list(d), "Ls": ls.value, "precision": prec.value, "otype": otype.value, "cb": cb.value}


def create_lattice_like(h):
    """a new lattice of the same grid, precision, object type and checkerboard label"""
    i = lattice_info(h)
    return create_lattice(i["dims4"], i["Ls"], i["precision"], i["otype"], i["cb"])


def delete_lattice(h):
    if _lib is not None:
        _lib.cgptb_delete_lattice(c_void_p(h))


def lattice_bytes(h):
    return _lib_ready().cgptb_lattice_bytes(c_void_p(h))


def lattice_device_ptr(h):
    return _lib_ready().cgptb_lattice_device_ptr(c_void_p(h))


def lattice_get_checkerboard(h):
    return _lib_ready().cgptb_lattice_get_checkerboard(c_void_p(h))


def lattice_change_checkerboard(h, cb):
    _check(_lib_ready().cgptb_lattice_change_checkerboard(c_void_p(h), cb))


def lattice_set_to_zero(h):
    _check(_lib_ready().cgptb_lattice_set_to_zero(c_void_p(h)))


def lattice_import(h, array):
    a = np.ascontiguousarray(array)
    _check(_lib_ready().cgptb_lattice_import(c_void_p(h), a.ctypes.data_as(c_void_p), a.nbytes))


def lattice_export(h, array):
    assert array.flags["C_CONTIGUOUS"]
    _check(_lib_ready().cgptb_lattice_export(c_void_p(h), array.ctypes.data_as(c_void_p), array.nbytes))


def lattice_import_ptr(h, host_ptr, nbytes):
    _check(_lib_ready().cgptb_lattice_import(c_void_p(h), c_void_p(host_ptr), nbytes))


def lattice_export_ptr(h, host_ptr, nbytes):
    _check(_lib_ready().cgptb_lattice_export(c_void_p(h), c_void_p(host_ptr), nbytes))


def lattice_import_device(h, dev_ptr, nbytes):
    _check(_lib_ready().cgptb_lattice_import_device(c_void_p(h), c_void_p(dev_ptr), nbytes))


def lattice_export_device(h, dev_ptr, nbytes):
    _check(_lib_ready().cgptb_lattice_export_device(c_void_p(h), c_void_p(dev_ptr), nbytes))


def copy(dst, src):
    _check(_lib_ready().cgptb_lattice_copy(c_void_p(dst), c_void_p(src)))


def convert(dst, src):
    _check(_lib_ready().cgptb_lattice_convert(c_void_p(dst), c_void_p(src)))


def lattice_pick_checkerboard(cb, half, full):
    _check(_lib_ready().cgptb_lattice_pick_checkerboard(cb, c_void_p(half), c_void_p(full)))


def lattice_set_checkerboard(full, half):
    _check(_lib_ready().cgptb_lattice_set_checkerboard(c_void_p(full), c_void_p(half)))


# ---- vector kernels ------------------------------------------------------------------------------------
def lattice_axpy(r, a, x, y):
    a = complex(a)
    _check(_lib_ready().cgptb_lattice_axpy(c_void_p(r), a.real, a.imag, c_void_p(x), c_void_p(y)))


def lattice_axpy_norm2(r, a, x, y):
    a = complex(a)
    n = c_double()
    _check(_lib_ready().cgptb_lattice_axpy_norm2(c_void_p(r), a.real, a.imag, c_void_p(x), c_void_p(y), ctypes.byref(n)))
    return n.value


def lattice_rank_inner_product(left, right):
    res = np.zeros((len(left), len(right)), dtype=np.complex128)
    _check(
        _lib_ready().cgptb_lattice_rank_inner_product(
            _handles(left), len(left), _handles(right), len(right), res.ctypes.data_as(_pd)
        )
    )
    return res


def lattice_norm2(a):
    n = c_double()
    _check(_lib_ready().cgptb_lattice_norm2(c_void_p(a), ctypes.byref(n)))
    return n.value


def lattice_inner_product_norm2(a, b):
    ip = (c_double * 2)()
    n = c_double()
    _check(_lib_ready().cgptb_lattice_inner_product_norm2(c_void_p(a), c_void_p(b), ip, ctypes.byref(n)))
    return complex(ip[0], ip[1]), n.value


def lattice_lc(dst, accumulate, coefs, lattices):
    c = np.array([complex(x) for x in coefs], dtype=np.complex128)
    _check(
        _lib_ready().cgptb_lattice_lc(
            c_void_p(dst), 1 if accumulate else 0, len(lattices), c.ctypes.data_as(_pd), _handles(lattices)
        )
    )


def linear_combination(r, basis, Qt):
    q = np.ascontiguousarray(np.array(Qt, dtype=np.complex128))
    assert q.shape == (len(r), len(basis))
    _check(_lib_ready().cgptb_linear_combination(_handles(r), len(r), _handles(basis), len(basis), q.ctypes.data_as(_pd)))


def lattice_scale(h, a):
    a = complex(a)
    _check(_lib_ready().cgptb_lattice_scale(c_void_p(h), a.real, a.imag))


def lattice_slice_inner_product(b, a, nt):
    out = np.zeros(nt, dtype=np.complex128)
    _check(_lib_ready().cgptb_lattice_slice_inner_product(c_void_p(b), c_void_p(a), out.ctypes.data_as(_pd)))
    return out


def debug_tma_schedule(dims4, Ls, grid, chunks_per_cta=1, sched=1, trl=16, t_begin=0, t_count=0):
    """work items of the TMA sweep kernel: array [n, 6] = (group, xh0, y0, z0, t0, trl); needs no GPU"""
    d = (c_int * 4)(*[int(x) for x in dims4])
    cap = 1 << 16
    out = np.zeros((cap, 6), dtype=np.int32)
    n = c_int()
    lib = library()
    if lib.cgptb_debug_tma_schedule(d, int(Ls), int(grid), int(chunks_per_cta), int(sched), int(trl), int(t_begin), int(t_count),
                                    out.ctypes.data_as(_pi), cap, ctypes.byref(n)):
        raise RuntimeError(lib.cgptb_last_error().decode())
    assert n.value <= cap
    return out[: n.value].copy()


# ---- generic matrix-vector stencil ---------------------------------------------------------------------------
def stencil_matrix_vector_create(dims4, prec, points, code, block_size, local, matrix_parity, vector_parity):
    """code: list of dicts target / source / source_point / accumulate / weight / factor = [(index, point, adj)]"""
    pts = np.ascontiguousarray(np.array(points, dtype=np.int32).reshape(len(points), 4))
    ci, w, fac = [], [], []
    for c in code:
        ci.append([c["target"], c["accumulate"], c["source"], c["source_point"], len(c["factor"])])
        z = complex(c["weight"])
        w.append([z.real, z.imag])
        fac.extend([int(f[0]), int(f[1]), int(f[2])] for f in c["factor"])
    ci = np.ascontiguousarray(np.array(ci, dtype=np.int32))
    w = np.ascontiguousarray(np.array(w, dtype=np.float64))
    fac = np.ascontiguousarray(np.array(fac if fac else [[0, 0, 0]], dtype=np.int32))
    d = (c_int * 4)(*[int(x) for x in dims4])
    h = c_void_p()
    _check(_lib_ready().cgptb_stencil_matrix_vector_create(
        ctypes.byref(h), d, _precisions[prec] if isinstance(prec, str) else prec, len(points), pts.ctypes.data_as(_pi), len(code),
        ci.ctypes.data_as(_pi), w.ctypes.data_as(_pd), fac.ctypes.data_as(_pi), int(block_size), int(local), int(matrix_parity),
        int(vector_parity)))
    return h.value


def stencil_matrix_vector_execute(h, matrix_fields, vector_fields, fast_osites=0):
    _check(_lib_ready().cgptb_stencil_matrix_vector_execute(c_void_p(h), _handles(matrix_fields), len(matrix_fields),
                                                            _handles(vector_fields), len(vector_fields), int(fast_osites)))


def stencil_matrix_vector_delete(h):
    if _lib is not None:
        _lib.cgptb_stencil_matrix_vector_delete(c_void_p(h))


# ---- fermion operators -----------------------------------------------------------------------------------
def _params(p):
    fp = fermion_params()
    for k in ["mass", "csw_r", "csw_t", "cF", "xi_0", "nu", "mass_plus", "mass_minus", "M5", "b", "c", "mu"]:
        v = p.get(k, None)
        setattr(fp, k, float(v) if v is not None else 0.0)
    fp.isAnisotropic = 1 if p.get("isAnisotropic", False) else 0
    omega = p.get("omega", None) or []
    if len(omega) > 64:
        raise RuntimeError("zmobius supports Ls <= 64")
    fp.n_omega = len(omega)
    for i, w in enumerate(omega):
        fp.omega[2 * i], fp.omega[2 * i + 1] = complex(w).real, complex(w).imag
    fp.Ls = int(p.get("Ls", 0) or 0)
    fp.link_compression = int(p.get("link_compression", 0) or 0)
    bp = p.get("boundary_phases", [1.0, 1.0, 1.0, 1.0])
    for i in range(4):
        z = complex(bp[i])
        fp.boundary_phases[2 * i] = z.real
        fp.boundary_phases[2 * i + 1] = z.imag
    return fp


# wilson_twisted_mass is the Wilson operator with the parameter mu set (no clover term)
# zmobius is the Moebius operator with the complex omega_s set
_optypes = {"wilson_clover": WILSON_CLOVER, "wilson_twisted_mass": WILSON_CLOVER, "mobius": MOBIUS, "zmobius": MOBIUS}
_precisions = {"single": SINGLE, "double": DOUBLE}


def create_fermion_operator(optype, prec, params):
    """cgpt.create_fermion_operator(optype, prec, params) -> handle (lib/cgpt/lib/operators.cc:34-57)"""
    if optype not in _optypes:
        raise RuntimeError(f"Unknown operator type {optype}")
    h = c_void_p()
    fp = _params(params)
    _check(
        _lib_ready().cgptb_create_fermion_operator(
            ctypes.byref(h), _optypes[optype], _precisions[prec], ctypes.byref(fp), _handles(params["U"])
        )
    )
    return h.value


def update_fermion_operator(h, params):
    _check(_lib_ready().cgptb_update_fermion_operator(c_void_p(h), _handles(params["U"])))


def set_mass_fermion_operator(h, params):
    fp = _params(params)
    _check(_lib_ready().cgptb_set_mass_fermion_operator(c_void_p(h), ctypes.byref(fp)))


def delete_fermion_operator(h):
    if _lib is not None:
        _lib.cgptb_delete_fermion_operator(c_void_p(h))


def apply_fermion_operator(h, opcode, src, dst):
    """note the (src, dst) order (lib/cgpt/lib/operators.cc:96-107)"""
    _check(_lib_ready().cgptb_apply_fermion_operator(c_void_p(h), int(opcode), c_void_p(src), c_void_p(dst)))
    return 0.0


# ---- random numbers (lib/cgpt/lib/random.cc:38-101) -------------------------------------------------------------
DISTRIBUTIONS = {"normal": 0, "cnormal": 1, "uniform_real": 2, "uniform_int": 3, "zn": 4}


def _dist_params(p):
    d = p["distribution"]
    if d in ("normal", "cnormal"):
        return DISTRIBUTIONS[d], float(p["mu"]), float(p["sigma"])
    if d in ("uniform_real", "uniform_int"):
        return DISTRIBUTIONS[d], float(p["min"]), float(p["max"])
    if d == "zn":
        return DISTRIBUTIONS[d], float(p["n"]), 0.0
    raise RuntimeError(f"Unknown distribution: {d}")


def create_random(engine, seed):
    """the library itself needs no device for this: the generators are host code"""
    h = c_void_p()
    _check(library().cgptb_create_random(ctypes.byref(h), engine.encode(), seed.encode()))
    return h.value


def delete_random(h):
    if _lib is not None:
        _lib.cgptb_delete_random(c_void_p(h))


def random_sample_scalar(h, p):
    out = (c_double * 2)()
    d, p0, p1 = _dist_params(p)
    _check(library().cgptb_random_sample_scalar(c_void_p(h), d, p0, p1, out))
    return complex(out[0], out[1])


def random_sample_host(h, grid_key, ldims, gdims, lstart, nel, p):
    """numpy array [sites, nel] complex128 in GPT order (dimension 0 fastest); no device needed"""
    import numpy as np

    nd = len(ldims)
    out = np.empty((int(np.prod(ldims)), nel), dtype=np.complex128)
    d, p0, p1 = _dist_params(p)
    arr = lambda v: (c_int * nd)(*[int(x) for x in v])  # noqa: E731
    _check(library().cgptb_random_sample_host(c_void_p(h), int(grid_key), nd, arr(ldims), arr(gdims), arr(lstart), int(nel), d, p0, p1,
                                              out.ctypes.data_as(_pd)))
    return out


def random_sample(h, grid_key, lattice, p):
    d, p0, p1 = _dist_params(p)
    _check(_lib_ready().cgptb_random_sample(c_void_p(h), int(grid_key), c_void_p(lattice), d, p0, p1))


def random_su3_links(h, grid_key, U, scale):
    arr = (c_void_p * 4)(*U)
    _check(_lib_ready().cgptb_random_su3_links(c_void_p(h), int(grid_key), arr, float(scale)))


def lattice_spin_matrix(d, s, m):
    import numpy as np

    m = np.ascontiguousarray(np.asarray(m, dtype=np.complex128).reshape(4, 4))
    _check(_lib_ready().cgptb_lattice_spin_matrix(c_void_p(d), c_void_p(s), m.view(np.float64).ctypes.data_as(_pd)))


def lattice_scale_per_coordinate(d, s, a, dim):
    import numpy as np

    a = np.ascontiguousarray(np.asarray(a, dtype=np.complex128))
    _check(_lib_ready().cgptb_lattice_scale_per_coordinate(c_void_p(d), c_void_p(s), a.view(np.float64).ctypes.data_as(_pd), len(a), int(dim)))


def lattice_pack_rhs(l5, l4, unpack=False):
    _check(_lib_ready().cgptb_lattice_pack_rhs(c_void_p(l5), (c_void_p * len(l4))(*l4), len(l4), 1 if unpack else 0))


def gauge_plaquette(U):
    """(plaquette, link trace) of four link lattices"""
    out = (c_double * 2)()
    _check(_lib_ready().cgptb_gauge_plaquette((c_void_p * 4)(*U), out))
    return out[0], out[1]


def nersc_munge(raw, float_size, big_endian, rows, U):
    """raw: C-contiguous uint8 numpy array with the data part of a NERSC file; returns the checksum"""
    cs = ctypes.c_uint(0)
    _check(_lib_ready().cgptb_nersc_munge(c_void_p(raw.ctypes.data), int(raw.nbytes), int(float_size), 1 if big_endian else 0, int(rows),
                                          (c_void_p * 4)(*U), ctypes.byref(cs)))
    return cs.value


def apply_fermion_operator_host(h, opcode, src_ptr, dst_ptr, nbytes):
    """same operator on host buffers (addresses of full fields in GPT order); upload / stencil / download are pipelined"""
    _check(_lib_ready().cgptb_apply_fermion_operator_host(c_void_p(h), int(opcode), c_void_p(src_ptr), c_void_p(dst_ptr), int(nbytes)))
    return 0.0


def apply_schur_two(h, dag, src, dst):
    _check(_lib_ready().cgptb_apply_schur_two(c_void_p(h), 1 if dag else 0, c_void_p(src), c_void_p(dst)))


def cg_eo2_ne(h, psi, src, eps, maxiter):
    hist = np.zeros(max(int(maxiter), 1), dtype=np.float64)
    it, conv = c_int(), c_int()
    _check(
        _lib_ready().cgptb_cg_eo2_ne(
            c_void_p(h), c_void_p(psi), c_void_p(src), float(eps), int(maxiter), hist.ctypes.data_as(_pd),
            ctypes.byref(it), ctypes.byref(conv)
        )
    )
    return list(hist[: it.value]), bool(conv.value)

"""
ORACLE (test infrastructure only): ctypes loader for oracle/dslash_ref.c (the C/OpenMP restatement of Dhop
used as cross-check on larger volumes and as bench.py's CPU baseline).  Never imported by gpt_b200/.
"""
import ctypes
import os
import subprocess

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _here])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_here, "liboracle_dslash.so")
        if not os.path.exists(path):
            build()
        _lib = ctypes.CDLL(path)
        for name, real in [("oracle_dhop_f", ctypes.c_float), ("oracle_dhop_d", ctypes.c_double)]:
            f = getattr(_lib, name)
            f.restype = None
            f.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        for name in ("oracle_dhop_sites_f", "oracle_dhop_sites_d"):
            f = getattr(_lib, name)
            f.restype = None
            f.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                          ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_set_num_threads.argtypes = [ctypes.c_int]
        _lib.oracle_set_num_threads.restype = None
    return _lib


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    """OpenMP threads of the C port (torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every host core)"""
    lib().oracle_set_num_threads(int(n))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def dhop_sites(dims4, Ls, V, psi, site_list, coef=(1.0, 1.0, 1.0, 1.0), dag=False):
    """Dhop on the listed 4d sites only (lexicographic indices, x fastest): returns [n*max(Ls,1), 4, 3]"""
    V = np.ascontiguousarray(V)
    psi = np.ascontiguousarray(psi)
    assert V.dtype == psi.dtype and V.dtype in (np.complex64, np.complex128)
    idx = np.ascontiguousarray(site_list, dtype=np.int64)
    ls = max(int(Ls), 1)
    out = np.empty((idx.size * ls, 4, 3), dtype=psi.dtype)
    d = (ctypes.c_int * 4)(*[int(x) for x in dims4])
    c = (ctypes.c_double * 4)(*[float(x) for x in coef])
    f = lib().oracle_dhop_sites_f if V.dtype == np.complex64 else lib().oracle_dhop_sites_d
    f(d, int(Ls), V.ctypes.data, c, psi.ctypes.data, idx.size, idx.ctypes.data, out.ctypes.data, 1 if dag else 0)
    return out


def dhop(dims4, Ls, V, psi, coef=(1.0, 1.0, 1.0, 1.0), dag=False):
    """V: [4, V4, 3, 3] complex (phases applied), psi: [V4*max(Ls,1), 4, 3] complex, both GPT order"""
    V = np.ascontiguousarray(V)
    psi = np.ascontiguousarray(psi)
    assert V.dtype == psi.dtype and V.dtype in (np.complex64, np.complex128)
    out = np.empty_like(psi)
    d = (ctypes.c_int * 4)(*[int(x) for x in dims4])
    c = (ctypes.c_double * 4)(*[float(x) for x in coef])
    f = lib().oracle_dhop_f if V.dtype == np.complex64 else lib().oracle_dhop_d
    f(d, int(Ls), V.ctypes.data, c, psi.ctypes.data, out.ctypes.data, 1 if dag else 0)
    return out

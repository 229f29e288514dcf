"""
CPU tests of the drop-in boundary and the host logic (no GPU needed):
  * libcgpt_b200.so loads and exports every symbol include/cgpt_b200.h declares (no compute calls here);
  * without a CUDA device the product path fails loudly instead of falling back to anything on the CPU;
  * parameter handling mirrors the reference (@params_convention, opcode table);
  * the processor-grid helpers of the multi-GPU path.
"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "cgpt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cgptb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gpt_b200 import cgpt

    lib = ctypes.CDLL(cgpt.LIBRARY_PATH)
    declared = _header_symbols()
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cgpt_b200.h but not exported"
    # the ctypes binding covers the same set
    assert sorted(cgpt.SIGNATURES) == declared
    cgpt.library()  # attaches the signatures


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from gpt_b200 import cgpt

    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        cgpt.init(0)
    import gpt_b200 as g

    with pytest.raises(RuntimeError):
        g.vspincolor(g.grid([4, 4, 4, 4], g.double))


def test_product_never_imports_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "gpt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_params_convention():
    import gpt_b200 as g

    @g.params_convention(a=1, b=None)
    def f(x, params):
        return x, params

    assert f(3, {"a": 2}) == (3, {"a": 2, "b": None})
    assert f(3, b=5) == (3, {"a": 1, "b": 5})
    assert f(3, {"a": 2}, a=7)[1]["a"] == 7
    with pytest.raises(Exception):
        f(3, c=1)


def test_opcode_table_matches_reference_registry():
    # lib/cgpt/lib/operators/register.h:2-20
    from gpt_b200.qcd.fermion.register import OPCODES

    ref = {"M": 2001, "Mdag": 2002, "Meooe": 2003, "MeooeDag": 2004, "Mooee": 2005, "MooeeDag": 2006, "MooeeInv": 2007,
           "MooeeInvDag": 2008, "Mdiag": 2009, "Dminus": 2010, "DminusDag": 2011, "ImportPhysicalFermionSource": 2012,
           "ImportUnphysicalFermion": 2013, "ExportPhysicalFermionSolution": 2014, "ExportPhysicalFermionSource": 2015,
           "Dhop": 3001, "DhopDag": 4001, "DhopEO": 3002, "DhopEODag": 4002}
    assert OPCODES == ref
    header = open(os.path.join(ROOT, "include", "cgpt_b200.h")).read()
    for name, code in ref.items():
        assert re.search(rf"CGPTB_OP_{name}\s*=\s*{code}\b", header), name


def test_processor_grid_helpers():
    from gpt_b200 import parallel

    assert parallel.default_mpi(1) == [1, 1, 1, 1]
    assert parallel.default_mpi(2) == [1, 1, 1, 2]
    assert parallel.default_mpi(4) == [1, 1, 1, 4]
    assert parallel.default_mpi(8) == [1, 1, 2, 4]
    assert parallel.local_dims([32, 32, 64, 256], [1, 1, 2, 4]) == [32, 32, 32, 64]
    with pytest.raises(ValueError):
        parallel.local_dims([8, 8, 8, 12], [1, 1, 1, 4])  # local extent 3 is odd
    coords = [parallel.processor_coor(r, [1, 1, 2, 4]) for r in range(8)]
    assert coords[0] == [0, 0, 0, 0] and coords[1] == [0, 0, 1, 0] and coords[2] == [0, 0, 0, 1] and coords[7] == [0, 0, 1, 3]
    assert len({tuple(c) for c in coords}) == 8


def test_nersc_header_parsing(tmp_path):
    """host side of g.load: header keys, data offset, non-NERSC files are rejected (no device involved)"""
    from gpt_b200.io import nersc

    p = tmp_path / "cfg"
    body = b"\x00" * 64
    p.write_bytes(b"BEGIN_HEADER\nDATATYPE = 4D_SU3_GAUGE\nDIMENSION_1 = 4\nCHECKSUM =   1a2b\nFLOATING_POINT = IEEE32BIG\nEND_HEADER\n" + body)
    md, off = nersc.read_header(str(p))
    assert md["DATATYPE"] == "4D_SU3_GAUGE" and md["CHECKSUM"] == "1a2b" and md["FLOATING_POINT"] == "IEEE32BIG"
    assert p.read_bytes()[off:] == body
    q = tmp_path / "other"
    q.write_bytes(b"\xff\xfe not a header")
    assert nersc.read_header(str(q)) is None
    assert nersc.read_header(str(tmp_path / "missing")) is None
    assert nersc._tolerance("0.588123", 1e-16) == 1e-4 and nersc._tolerance("1.5e-3", 1e-7) == 10.0
    assert nersc.format.nersc(label="x").params == {"label": "x", "id": "gpt", "sequence_number": 1}


def test_gamma_algebra_host():
    """g.gamma tables and their algebra (numpy side; the kernel is covered by the GPU tests): Clifford algebra, hermiticity,
    gamma_5 = gamma_0 gamma_1 gamma_2 gamma_3, sigma_{mu nu} against the oracle's basis (lib/gpt/core/gamma.py:28-53)"""
    import numpy as np

    import gpt_b200 as g
    from oracle import qcd

    for mu in range(4):
        assert np.array_equal(g.gamma[mu].matrix, qcd.gamma[mu])
        assert np.array_equal(g.gamma["XYZT"[mu]].matrix, qcd.gamma[mu])
        assert np.allclose(g.gamma[mu].adj().matrix, g.gamma[mu].matrix)
        for nu in range(4):
            acomm = (g.gamma[mu] * g.gamma[nu] + g.gamma[nu] * g.gamma[mu]).matrix
            assert np.allclose(acomm, 2.0 * np.identity(4) * (mu == nu))
            if mu != nu:
                assert np.allclose(g.gamma[mu, nu].matrix, qcd.sigma(mu, nu))
    g5 = g.gamma[0] * g.gamma[1] * g.gamma[2] * g.gamma[3]
    assert np.allclose(g5.matrix, g.gamma[5].matrix)
    Pp, Pm = 0.5 * (g.gamma["I"] + g.gamma[5]), 0.5 * (g.gamma["I"] - g.gamma[5])
    assert np.allclose((Pp * Pp).matrix, Pp.matrix) and np.allclose((Pp * Pm).matrix, 0.0) and np.allclose((Pp + Pm).matrix, np.identity(4))
    assert np.allclose((g.gamma[5].inv() * g.gamma[5]).matrix, np.identity(4))
    assert np.allclose((-g.gamma[2]).matrix, -qcd.gamma[2]) and np.allclose((2j * g.gamma[2]).matrix, 2j * qcd.gamma[2])


def test_params_convention_semantics():
    """same calling conventions as lib/gpt/params.py:20-82: dict and keyword arguments merge, positional arguments with defaults
    may be omitted, unknown parameters raise"""
    import pytest

    from gpt_b200.params import params_convention

    @params_convention(a=1, b=2)
    def f(x, y=None, p={}):
        return x, y, p

    assert f(5) == (5, None, {"a": 1, "b": 2})
    assert f(5, 6, {"a": 3}) == (5, 6, {"a": 3, "b": 2})
    assert f(5, 6, {"a": 3}, b=4) == (5, 6, {"a": 3, "b": 4})
    assert f(5, b=7) == (5, None, {"a": 1, "b": 7})
    with pytest.raises(KeyError):
        f(5, c=1)

    class C:
        @params_convention(eps=1e-15)
        def __init__(self, params):
            self.params = params

    assert C().params == {"eps": 1e-15} and C(eps=1e-8).params == {"eps": 1e-8} and C({"eps": 1e-3}).params == {"eps": 1e-3}


def test_bench_reference_arm_prints_contract_line():
    """`python bench.py --impl reference` needs no GPU: one JSON line with the keys of the bench contract (impl, metric, unit, value,
    cpu_baseline kind / cores / sample, e2e with zero copy bytes)"""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mobius_dwf_dslash_gflops" and d["unit"] == "GFlop/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 2 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "Dhop" in d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]

"""
The reference's OWN scripts for the hot path, executed unchanged over this repo's `gpt` / `cgpt` modules on the GPU.

The scripts are not part of this repository: oracle/build_ref.py compiles them from /root/reference (where that tree exists)
into code objects under oracle/_ref/, which travel to the GPU box like a built library.  Without them the tests skip.

  * benchmarks/dslash.py and benchmarks/wilson_clover_dslash.py with their own command line flags (g.default), both
    precisions, --full timings (g.timer, pick/set_checkerboard, Meooe, Mooee);
  * the finger-print suite of tests/qcd/fermion_operators.py: the reference's parameter tables and golden numbers, its verify_*
    helpers (adjoints, inverses, even-odd projections, daggered operator, single versus double) and its loop, for every operator
    of the suite this package implements;
  * lib/gpt/qcd/fermion/operator/interface.py + register.py (handle cache, opcode dispatch, `cgpt.apply_fermion_operator(op,
    opcode, src.v_obj, dst.v_obj)`) and lib/gpt/algorithms/inverter/cg.py driving this library through the `cgpt` stand-in.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import pytest

from oracle import build_ref

pytestmark = pytest.mark.gpu


def _code(name):
    if os.path.isdir(build_ref.REFERENCE):
        build_ref.build()
    code = build_ref.load(name)
    if code is None:
        pytest.skip("oracle/_ref is not built (the reference tree is only present in the build container)")
    return code


@pytest.fixture(scope="module")
def g():
    import gpt  # the alias package: this IS gpt_b200

    gpt.cgpt.init(0)
    return gpt


def _run_script(code, argv):
    old = sys.argv
    sys.argv = list(argv)
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out):
            exec(code, {"__name__": "__main__"})
    finally:
        sys.argv = old
    return out.getvalue()


def test_reference_benchmark_dslash_runs_unchanged(g):
    out = _run_script(_code("benchmarks__dslash"), ["dslash.py", "--grid", "8.8.8.16", "--Ls", "8", "--N", "3", "--full"])
    print(out)
    assert out.count("3 applications of Dhop") == 2 and "precision    : single" in out and "precision    : double" in out
    assert out.count("GFlops/s") == 2 and "Full Timings" in out and "Meooe" in out and "Promote to full" in out


def test_reference_benchmark_wilson_clover_dslash_runs_unchanged(g):
    # BASELINE.json configs[0]: 16^4 (the script's default grid is 16^3 x 32)
    out = _run_script(_code("benchmarks__wilson_clover_dslash"), ["wilson_clover_dslash.py", "--grid", "16.16.16.16", "--N", "10"])
    print(out)
    assert out.count("10 applications of Dhop") == 2 and "precision    : double" in out
    out = _run_script(_code("benchmarks__wilson_clover_dslash"), ["x", "--grid", "8.8.8.8", "--N", "2", "--n_rhs", "4"])
    assert out.count("2 applications of Dhop") == 2


def test_reference_fingerprint_suite_runs_unchanged(g):
    """tests/qcd/fermion_operators.py: set-up block, parameter tables + golden finger prints + verify_* helpers, and the suite loop,
    all as the reference wrote them.  The only intervention: operators this package does not provide (the pure-Python
    g.qcd.fermion.reference implementations) are taken out of `test_suite` between the definitions and the loop."""
    ns = {"__name__": "__main__"}
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        exec(_code("tests__qcd__fermion_operators__head"), ns)
        exec(_code("tests__qcd__fermion_operators__defs"), ns)
        suite = ns["test_suite"]
        dropped = [k for k, v in suite.items() if v["fermion"] is g.qcd.fermion.reference.wilson_clover]
        for k in dropped:
            del suite[k]
        ran = list(suite)
        exec(_code("tests__qcd__fermion_operators__loop"), ns)
    text = out.getvalue()
    print(text[-3000:])
    assert sorted(dropped) == ["wilson_clover_openbc_reference", "wilson_clover_reference"]
    assert set(ran) >= {"zmobius", "mobius", "mobius_axial_mass", "wilson", "wilson_clover", "wilson_clover_openbc", "wilson_twisted_mass"}
    for name in ran:
        assert f"Starting test suite for {name}" in text
    assert text.count("fingerprint:") >= 2 * len(ran)


def test_reference_interface_and_cg_over_the_cgpt_stand_in(g, monkeypatch):
    """the reference's interface.py / register.py replace this package's operator front end, the reference's cg.py replaces
    inv.cg: same results as the package's own classes, CG histories identical iteration by iteration"""
    import gpt_b200.qcd.fermion.operator as op_mod
    from oracle import qcd
    from oracle.rng import random as oracle_random
    from tests.util import rel, to_links

    ref_if = types.ModuleType("reference_interface")
    exec(_code("lib__gpt__qcd__fermion__operator__interface"), ref_if.__dict__)
    ref_reg = types.ModuleType("reference_register")
    exec(_code("lib__gpt__qcd__fermion__register"), ref_reg.__dict__)
    ref_cg = types.ModuleType("reference_cg")
    exec(_code("lib__gpt__algorithms__inverter__cg"), ref_cg.__dict__)

    dims = [8, 8, 8, 8]
    rng = oracle_random("reference glue")
    U = qcd.gauge_random(rng, dims, scale=0.7)
    params = dict(mass=0.12, M5=1.8, b=1.5, c=0.5, Ls=8, boundary_phases=[1.0, 1.0, 1.0, -1.0])
    s_np = rng.cnormal([8] + dims, (4, 3))
    oo = qcd.mobius(U, **params)
    results = {}
    for which in ("package", "reference"):
        if which == "reference":
            monkeypatch.setattr(op_mod, "interface", ref_if.interface)
            monkeypatch.setattr(op_mod, "register", ref_reg.register)
        grid = g.grid(dims, g.double)
        op = g.qcd.fermion.mobius(to_links(g, grid, U), dict(params))
        assert type(op.interface).__module__ == ("reference_interface" if which == "reference" else op_mod.__name__)
        src = g.vspincolor(op.F_grid)
        src[:] = s_np.reshape(-1, 4, 3)
        got = g(op * src)[:].reshape(s_np.shape)
        assert rel(got, oo.M(s_np)) < 1e-12
        got = g(op.adj() * src)[:].reshape(s_np.shape)
        assert rel(got, oo.Mdag(s_np)) < 1e-12
        inv = g.algorithms.inverter
        hist = {}
        for solver in ("package cg (device loop)", "package cg (python loop)", "reference cg.py"):
            if solver == "package cg (python loop)":
                monkeypatch.setenv("GPT_B200_NO_FUSED", "1")
            cg = (ref_cg.cg if solver == "reference cg.py" else inv.cg)(eps=1e-8, maxiter=500)
            x = g(inv.preconditioned(g.qcd.fermion.preconditioner.eo2_ne(), cg)(op) * src)
            monkeypatch.delenv("GPT_B200_NO_FUSED", raising=False)
            hist[solver] = list(cg.history)
            r = g(op * x - src)
            assert (g.norm2(r) / g.norm2(src)) ** 0.5 < 1e-7
        results[which] = hist
        n = len(hist["reference cg.py"])
        assert n > 10
        for k, h in hist.items():
            assert len(h) == n, (which, k, len(h), n)
            assert np.allclose(h, hist["reference cg.py"], rtol=1e-6), (which, k)
        if which == "package":
            _, oh = qcd.solve_eo2_ne(oo, s_np, 1e-8, 500)
            assert len(oh) == n
    assert results["package"]["reference cg.py"] == results["reference"]["reference cg.py"]

"""
GPU parity tests: the CUDA path (through the C ABI, via the gpt_b200 host API) against the CPU oracle on the
same seeded inputs, and against the reference's golden fingerprints
(/root/reference/tests/qcd/fermion_operators.py:371-459).  Tolerances: 1e-12 relative (double), 1e-5 (single)
as BASELINE.json's north_star states; fingerprints to 100*eps like the reference (fermion_operators.py:527).
"""
import os

import numpy as np
import pytest

from oracle import qcd
from oracle.rng import random as oracle_random
from tests.util import from_spinor, rel, sites, to_links, to_spinor

pytestmark = pytest.mark.gpu

TOL = {"double": 1e-12, "single": 1e-5}

WILSON = dict(kappa=0.13500, csw_r=0.0, csw_t=0.0, xi_0=1.33111, nu=2.61, isAnisotropic=True, boundary_phases=[1.0, -1.0, 1.0, -1.0])
CLOVER = dict(WILSON, csw_r=1.5, csw_t=1.951)
MOBIUS = dict(mass=0.08, M5=1.8, b=1.5, c=0.5, Ls=12, boundary_phases=[1.0, -1.0, 1.0, -1.0])
MOBIUS_AXIAL = dict(mass_plus=0.08, mass_minus=0.11, M5=1.8, b=1.5, c=0.5, Ls=12, boundary_phases=[1.0, -1.0, 1.0, -1.0])
DIMS = [8, 8, 8, 16]


@pytest.fixture(scope="module")
def g():
    import gpt_b200 as g

    g.cgpt.init(0)
    return g


@pytest.fixture(scope="module")
def fields():
    # same draw order as tests/qcd/fermion_operators.py:849-884
    rng = oracle_random("finger_print")
    U = qcd.gauge_random(rng, DIMS)
    d5 = [12] + DIMS
    f = dict(U=U)
    f["src5"], f["dst5"] = rng.cnormal(d5, (4, 3)), rng.cnormal(d5, (4, 3))
    f["src4"], f["dst4"] = rng.cnormal(DIMS, (4, 3)), rng.cnormal(DIMS, (4, 3))
    rng_w = oracle_random("finger_print")
    qcd.gauge_random(rng_w, DIMS)
    f["srcw"], f["dstw"] = rng_w.cnormal(DIMS, (4, 3)), rng_w.cnormal(DIMS, (4, 3))
    return f


def prec_of(g, name):
    return g.double if name == "double" else g.single


# ---------------------------------------------------------------------------------------------------------
# storage seam
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["double", "single"])
@pytest.mark.parametrize("five_d", [False, True])
def test_import_export_checkerboard(g, fields, precision, five_d):
    p = prec_of(g, precision)
    src = fields["src5"] if five_d else fields["src4"]
    dims = ([12] if five_d else []) + DIMS
    grid = g.grid(dims, p)
    l = to_spinor(g, grid, src)
    back = l[:]
    assert np.array_equal(back, sites(src, 2).astype(p.complex_dtype))  # bit exact round trip
    grid_eo = grid.checkerboarded(g.redblack)
    for cb in [g.even, g.odd]:
        h = g.vspincolor(grid_eo)
        g.pick_checkerboard(cb, h, l)
        assert h.checkerboard() is cb
        ref = qcd.pick_checkerboard(src, cb.tag, ls=five_d).reshape(-1, 4, 3).astype(p.complex_dtype)
        assert np.array_equal(h[:], ref)
    # set_checkerboard reassembles the field
    l2 = g.vspincolor(grid)
    l2[:] = 0
    for cb in [g.even, g.odd]:
        h = g.vspincolor(grid_eo)
        g.pick_checkerboard(cb, h, l)
        g.set_checkerboard(l2, h)
    assert np.array_equal(l2[:], back)
    # convert
    q = g.single if precision == "double" else g.double
    c = g.convert(l, q)
    assert np.array_equal(c[:], back.astype(q.complex_dtype))


@pytest.mark.parametrize("precision", ["double", "single"])
def test_vector_kernels(g, fields, precision):
    p = prec_of(g, precision)
    grid = g.grid([12] + DIMS, p)
    a_np = sites(fields["src5"], 2).astype(p.complex_dtype)
    b_np = sites(fields["dst5"], 2).astype(p.complex_dtype)
    a, b = g.vspincolor(grid), g.vspincolor(grid)
    a[:] = a_np
    b[:] = b_np
    tol = 1e-14 if precision == "double" else 1e-6
    ip = g.inner_product(a, b)
    ref = np.vdot(a_np.astype(np.complex128), b_np.astype(np.complex128))
    assert abs(ip - ref) / abs(ref) < 1e-13  # double accumulation for both precisions (reduce.h:129)
    n2 = g.norm2(a)
    assert abs(n2 - np.vdot(a_np.astype(np.complex128), a_np.astype(np.complex128)).real) / n2 < 1e-13
    ipn = g.inner_product_norm2(a, b)
    assert abs(ipn[0] - ref) / abs(ref) < 1e-13 and abs(ipn[1] - n2) / n2 < 1e-13
    alpha = 0.37 - 1.21j
    r = g.lattice(a)
    g.axpy(r, alpha, a, b)
    want = p.complex_dtype(alpha) * a_np + b_np
    assert rel(r[:], want) < tol
    r2 = g.lattice(a)
    n = g.axpy_norm2(r2, alpha, a, b)
    assert np.array_equal(r2[:], r[:])
    assert abs(n - g.norm2(r)) / n < 1e-13
    # aliasing as cg.py uses it: r = -a*mmp + r ; p = b*p + r
    g.axpy(r, -0.25, a, r)
    want = p.complex_dtype(-0.25) * a_np + want
    assert rel(r[:], want) < tol
    # expressions: psi += a*p, dst @= x - y, scaling
    psi = g.copy(b)
    psi += 0.5 * a
    assert rel(psi[:], b_np + p.complex_dtype(0.5) * a_np) < tol
    d = g.lattice(a)
    d @= a - b
    assert rel(d[:], a_np - b_np) < tol
    d *= 2.0
    assert rel(d[:], 2 * (a_np - b_np)) < tol
    # linear_combination (basis.py:66-74)
    Qt = np.array([[0.3 + 0.1j, -1.0], [2.0, 0.5j]])
    out = [g.lattice(a), g.lattice(a)]
    g.linear_combination(out, [a, b], Qt)
    for i in range(2):
        assert rel(out[i][:], Qt[i, 0] * a_np + Qt[i, 1] * b_np) < 10 * tol


# ---------------------------------------------------------------------------------------------------------
# operators against the oracle and the golden fingerprints
# ---------------------------------------------------------------------------------------------------------
def _ops4(g, fields, params, precision):
    p = prec_of(g, precision)
    grid = g.grid(DIMS, p)
    U = to_links(g, grid, fields["U"])
    ctor = g.qcd.fermion.wilson_twisted_mass if "mu" in params else g.qcd.fermion.wilson_clover
    w = ctor(U, dict(params))
    Uo = [u.astype(p.complex_dtype) for u in fields["U"]]
    wo = qcd.wilson_clover(Uo, **params)
    return grid, w, wo


TWISTED = dict(mass=-1.8, mu=0.2, boundary_phases=[1.0, 1.0, 1.0, -1.0])  # tests/qcd/fermion_operators.py:365-369
# open boundary conditions in time with the boundary improvement cF (tests/qcd/fermion_operators.py:356-364)
OPEN = dict(kappa=0.13500, csw_r=1.978, csw_t=1.978, cF=1.3, xi_0=1, nu=1, isAnisotropic=False, boundary_phases=[1.0, 1.0, 1.0, 0.0])


@pytest.mark.parametrize("precision", ["double", "single"])
@pytest.mark.parametrize("name", ["wilson", "clover", "twisted", "open"])
def test_wilson_clover_full(g, fields, name, precision):
    params = {"wilson": WILSON, "clover": CLOVER, "twisted": TWISTED, "open": OPEN}[name]
    grid, w, wo = _ops4(g, fields, params, precision)
    tol = TOL[precision]
    src_np = fields["srcw"].astype(grid.precision.complex_dtype)
    src = to_spinor(g, grid, src_np)
    for tag, ref in [
        ("M", wo.M(src_np)), ("Mdag", wo.Mdag(src_np)), ("Mdiag", wo.Mdiag(src_np)),
        ("Dhop", wo.Dhop(src_np)), ("DhopDag", wo.Dhop(src_np, dag=True)),
    ]:
        op = {"M": w, "Mdag": w.adj(), "Mdiag": w.Mdiag, "Dhop": w.Dhop, "DhopDag": w.Dhop.adj()}[tag]
        got = from_spinor(g(op * src), src_np)
        assert rel(got, ref) < tol, tag
    if precision == "double":
        dst = to_spinor(g, grid, fields["dstw"])
        golden = {"wilson": (-999.7564252326631 - 466.7758727463097j, -961.5053827614738 - 3468.430447866095j),
                  "clover": (-946.8714968698364 - 427.1253034080037j, -908.620454398646 - 3428.779878527792j),
                  "twisted": (-5.665095757463064 + 373.96051873176737j, -440.5312395819657 - 1102.362512575698j),
                  "open": (-1634.2615676797234 + 239.27037187495998j, -1239.3535155227526 - 1158.5295177146759j)}[name]
        X = g.inner_product(dst, g(w * src))
        assert abs(X - golden[0]) / abs(golden[0]) < 1e-13
        X = g.inner_product(dst, g(w.Mdiag * src))
        assert abs(X - golden[1]) / abs(golden[1]) < 1e-13


@pytest.mark.parametrize("precision", ["double", "single"])
@pytest.mark.parametrize("name", ["wilson", "clover", "twisted", "open"])
def test_wilson_clover_eo(g, fields, name, precision):
    params = {"wilson": WILSON, "clover": CLOVER, "twisted": TWISTED, "open": OPEN}[name]
    grid, w, wo = _ops4(g, fields, params, precision)
    tol = TOL[precision]
    e = qcd.eo_ops(wo)
    src_np = fields["srcw"].astype(grid.precision.complex_dtype)
    for cb in [g.even, g.odd]:
        half_np = e.proj(src_np, cb.tag)
        half = to_spinor(g, w.F_grid_eo, src_np, cb)
        checks = [
            (w.Meooe, e.Meooe(half_np, cb.tag)), (w.Meooe.adj(), e.Meooe(half_np, cb.tag, dag=True)),
            (w.DhopEO, e.Meooe(half_np, cb.tag)),
            (w.Mooee, e.Mooee(half_np)), (w.Mooee.adj(), e.Mooee(half_np, dag=True)),
            (w.Mooee.inv(), e.MooeeInv(half_np)), (w.Mooee.adj().inv(), e.MooeeInv(half_np, dag=True)),
        ]
        for i, (op, ref) in enumerate(checks):
            got = g(op * half)
            assert rel(from_spinor(got, src_np), ref) < tol, (cb, i)
        assert g(w.Meooe * half).checkerboard() is cb.inv()
        assert g(w.Mooee * half).checkerboard() is cb


def _ops5(g, fields, params, precision):
    p = prec_of(g, precision)
    grid = g.grid(DIMS, p)
    U = to_links(g, grid, fields["U"])
    m = g.qcd.fermion.mobius(U, dict(params))
    Uo = [u.astype(p.complex_dtype) for u in fields["U"]]
    mo = qcd.mobius(Uo, **params)
    return grid, m, mo


@pytest.mark.parametrize("precision", ["double", "single"])
@pytest.mark.parametrize("name", ["mobius", "axial"])
def test_mobius_full(g, fields, name, precision):
    params = MOBIUS if name == "mobius" else MOBIUS_AXIAL
    grid, m, mo = _ops5(g, fields, params, precision)
    tol = TOL[precision]
    cdt = grid.precision.complex_dtype
    s5 = fields["src5"].astype(cdt)
    s4 = fields["src4"].astype(cdt)
    src5 = to_spinor(g, m.F_grid, s5)
    src4 = to_spinor(g, m.U_grid, s4)
    checks5 = [
        (m, mo.M(s5)), (m.adj(), mo.Mdag(s5)), (m.Mdiag, mo.Mdiag(s5)),
        (m.Dhop, mo.Dhop(s5)), (m.Dhop.adj(), mo.Dhop(s5, dag=True)),
        (m.Dminus, mo.Dminus(s5)), (m.Dminus.adj(), mo.Dminus(s5, dag=True)),
    ]
    for i, (op, ref) in enumerate(checks5):
        assert rel(from_spinor(g(op * src5), s5), ref) < tol, i
    assert rel(from_spinor(g(m.ImportPhysicalFermionSource * src4), s5), mo.ImportPhysicalFermionSource(s4)) < tol
    assert rel(from_spinor(g(m.ImportUnphysicalFermion * src4), s5), mo.ImportUnphysicalFermion(s4)) < tol
    assert rel(from_spinor(g(m.ExportPhysicalFermionSolution * src5), s4), mo.ExportPhysicalFermionSolution(s5)) < tol
    assert rel(from_spinor(g(m.ExportPhysicalFermionSource * src5), s4), mo.ExportPhysicalFermionSource(s5)) < tol
    if precision == "double":
        dst5 = to_spinor(g, m.F_grid, fields["dst5"])
        golden = {
            "mobius": (-8693.09425573421 - 4130.7793316734915j, -4966.960264746144 - 2525.83968136146j, -97.93443075273976 - 690.6405168964976j),
            "axial": (-8690.547330400455 - 4127.148886222195j, -4967.102993398692 - 2525.589904941078j, -97.93443075274081 - 690.6405168964941j),
        }[name]
        for op, s, ref in [(m, src5, golden[0]), (m.Mdiag, src5, golden[1]), (m.ImportPhysicalFermionSource, src4, golden[2])]:
            X = g.inner_product(dst5, g(op * s))
            assert abs(X - ref) / abs(ref) < 1e-13


@pytest.mark.parametrize("precision", ["double", "single"])
def test_mobius_eo(g, fields, precision):
    grid, m, mo = _ops5(g, fields, MOBIUS_AXIAL, precision)
    tol = TOL[precision]
    e = qcd.eo_ops(mo)
    s5 = fields["src5"].astype(grid.precision.complex_dtype)
    for cb in [g.even, g.odd]:
        half_np = e.proj(s5, cb.tag)
        half = to_spinor(g, m.F_grid_eo, s5, cb)
        checks = [
            (m.Meooe, e.Meooe(half_np, cb.tag)), (m.Meooe.adj(), e.Meooe(half_np, cb.tag, dag=True)),
            (m.DhopEO, e.proj(mo.Dhop(half_np), 1 - cb.tag)), (m.DhopEO.adj(), e.proj(mo.Dhop(half_np, dag=True), 1 - cb.tag)),
            (m.Mooee, e.Mooee(half_np)), (m.Mooee.adj(), e.Mooee(half_np, dag=True)),
            (m.Mooee.inv(), e.MooeeInv(half_np)), (m.Mooee.adj().inv(), e.MooeeInv(half_np, dag=True)),
        ]
        for i, (op, ref) in enumerate(checks):
            assert rel(from_spinor(g(op * half), s5), ref) < tol, (cb, i)


# ---------------------------------------------------------------------------------------------------------
# Schur complement + CG: same iteration count and residual history as the oracle's cg.py restatement
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("kind", ["mobius", "clover"])
def test_eo2_ne_cg_double(g, fields, kind, fused, monkeypatch):
    if not fused:
        monkeypatch.setenv("GPT_B200_NO_FUSED", "1")
    small = [4, 4, 4, 8]
    rng = oracle_random("cg_test")
    U = qcd.gauge_random(rng, small, scale=0.5)
    grid = g.grid(small, g.double)
    Ug = to_links(g, grid, U)
    if kind == "mobius":
        params = dict(mass=0.1, M5=1.8, b=1.5, c=0.5, Ls=6, boundary_phases=[1.0, 1.0, 1.0, -1.0])
        op, oo = g.qcd.fermion.mobius(Ug, dict(params)), qcd.mobius(U, **params)
        src_np = rng.cnormal([6] + small, (4, 3))
    else:
        params = dict(mass=0.2, csw_r=1.1, csw_t=1.3, xi_0=1.0, nu=1.0, isAnisotropic=False, boundary_phases=[1.0, 1.0, 1.0, -1.0])
        op, oo = g.qcd.fermion.wilson_clover(Ug, dict(params)), qcd.wilson_clover(U, **params)
        src_np = rng.cnormal(small, (4, 3))
    eps, maxiter = 1e-8, 500
    ref, hist_ref = qcd.solve_eo2_ne(oo, src_np, eps, maxiter)
    inv = g.algorithms.inverter
    cg = inv.cg(eps=eps, maxiter=maxiter)
    slv = inv.preconditioned(g.qcd.fermion.preconditioner.eo2_ne(), cg)(op)
    src = to_spinor(g, op.F_grid, src_np)
    dst = g(slv * src)
    assert len(cg.history) == len(hist_ref)  # identical iteration count
    assert np.allclose(cg.history, hist_ref, rtol=1e-6)
    assert rel(from_spinor(dst, src_np), ref) < 1e-10
    # true residual
    r = g(op * dst - src)
    assert (g.norm2(r) / g.norm2(src)) ** 0.5 < 1e-6


def test_mixed_precision_defect_correction(g, fields):
    # pattern of /root/reference/tests/manual/mpi.py:104-110 and tests/algorithms/solvers.py:112-118
    small = [4, 4, 4, 8]
    rng = oracle_random("cg_test")
    U = qcd.gauge_random(rng, small, scale=0.5)
    grid = g.grid(small, g.double)
    Ug = to_links(g, grid, U)
    params = dict(mass=0.2, csw_r=1.1, csw_t=1.3, xi_0=1.0, nu=1.0, isAnisotropic=False, boundary_phases=[1.0, 1.0, 1.0, -1.0])
    op = g.qcd.fermion.wilson_clover(Ug, dict(params))
    src_np = rng.cnormal(small, (4, 3))
    src = to_spinor(g, grid, src_np)
    inv = g.algorithms.inverter
    pc = g.qcd.fermion.preconditioner
    cg = inv.cg(eps=1e-4, maxiter=500)
    dc = inv.defect_correcting(inv.mixed_precision(inv.preconditioned(pc.eo2_ne(), cg), g.single, g.double), eps=1e-10, maxiter=20)
    dst = g(dc(op) * src)
    r = g(op * dst - src)
    assert (g.norm2(r) / g.norm2(src)) ** 0.5 < 1e-10
    assert len(dc.history) >= 2


def test_errors(g):
    grid = g.grid([4, 4, 4, 4], g.double)
    a = g.vspincolor(grid)
    b = g.vspincolor(g.grid([4, 4, 4, 8], g.double))
    with pytest.raises(RuntimeError):
        g.axpy(a, 1.0, a, b)
    with pytest.raises(Exception):
        g.qcd.fermion.mobius(g.qcd.gauge.unit(grid), {"bogus": 1})
    with pytest.raises(ValueError):
        g.grid([4, 4, 4], g.double)
    with pytest.raises((RuntimeError, ValueError)):
        g.vspincolor(g.grid([3, 4, 4, 4], g.double))


@pytest.mark.parametrize("Ls", [12, 8])
def test_fused_schur_single(g, fields, Ls):
    """fp32 Moebius: the fused path (sweep kernel + Dslash with T / axpy / dot epilogues) against the oracle's Mpc,
    Mpc^dag and against the unfused opcode sequence; CG on Mpc^dag Mpc with the oracle's iteration count."""
    params = dict(mass_plus=0.08, mass_minus=0.11, M5=1.8, b=1.5, c=0.5, Ls=Ls, boundary_phases=[1.0, -1.0, 1.0, -1.0])
    grid = g.grid(DIMS, g.single)
    U = to_links(g, grid, fields["U"])
    m = g.qcd.fermion.mobius(U, dict(params))
    Uo = [u.astype(np.complex64) for u in fields["U"]]
    mo = qcd.mobius(Uo, **params)
    sc = qcd.schur_complement_two(mo, parity=1)
    rng = oracle_random("fused")
    s5 = rng.cnormal([Ls] + DIMS, (4, 3)).astype(np.complex64)
    e = qcd.eo_ops(mo)
    half_np = e.proj(s5, 1)
    half = to_spinor(g, m.F_grid_eo, s5, g.odd)
    out = g.lattice(half)
    for dag, ref in [(False, sc.Mpc(half_np)), (True, sc.MpcDag(half_np))]:
        g.cgpt.apply_schur_two(m.interface.obj, dag, half.obj, out.obj)
        assert out.checkerboard() is g.odd
        assert rel(from_spinor(out, s5), ref) < 1e-5, dag
    if Ls != 8:
        return
    # CG: fused device loop vs oracle (single precision: allow the count to differ by one)
    eps, maxiter = 1e-4, 300
    src_np = sc.MpcDag(sc.R(s5))
    ref, hist = qcd.cg(lambda x: sc.MpcDag(sc.Mpc(x)), src_np, eps, maxiter)
    src = to_spinor(g, m.F_grid_eo, src_np, g.odd)
    psi = g.lattice(src)
    psi[:] = 0
    h, conv = g.cgpt.cg_eo2_ne(m.interface.obj, psi.obj, src.obj, eps, maxiter)
    assert conv and abs(len(h) - len(hist)) <= 1
    assert rel(from_spinor(psi, s5), ref) < 1e-3


def test_wilson_pion_correlator_golden(g):
    """the reference's own end-to-end check (tests/qcd/fermion_operators.py:12-43,135-218) through the drop-in API:
    single precision Wilson operator with complex boundary phases, eo2_ne CG propagator from a point source,
    g.slice(g.trace(dst * g.adj(dst)), 3) against the 16 golden values (tolerance 1e-5)"""
    from tests.test_oracle_golden import PION_PARAMS, PION_REF

    rng = oracle_random("test")
    U = qcd.gauge_random(rng, DIMS)
    grid = g.grid(DIMS, g.single)
    Ug = to_links(g, grid, U)
    w = g.qcd.fermion.wilson_clover(Ug, dict(PION_PARAMS))
    src = g.mspincolor(grid)
    g.create.point(src, [1, 0, 0, 0])
    inv = g.algorithms.inverter
    pc = g.qcd.fermion.preconditioner
    cg = inv.cg({"eps": 1e-6, "maxiter": 1000})
    slv_eo2 = w.propagator(inv.preconditioned(pc.eo2_ne(), cg))
    dst = g(slv_eo2 * src)
    assert 0 < len(cg.history) < 1000
    correlator = g.slice(g.trace(dst * g.adj(dst)), 3)
    eps = np.linalg.norm(np.array(correlator) - np.array(PION_REF))
    assert eps < 1e-5
    # true residuum of one column (fermion_operators.py:165-168)
    r = g(w * dst.columns[0] - src.columns[0])
    assert g.norm2(r) / g.norm2(src.columns[0]) < 1e-10


# ---------------------------------------------------------------------------------------------------------
# host-buffer operator call (cgptb_apply_fermion_operator_host): pipelined upload / stencil / download
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("slabs", ["16", "4", "off"])
def test_dhop_host_buffers(g, fields, slabs, monkeypatch):
    """op.Dhop_host(dst, src) on numpy buffers == oracle Dhop; `off` forces the plain import -> apply -> export path"""
    if slabs == "off":
        monkeypatch.setenv("CGPTB_NO_HOSTPIPE", "1")
    else:
        monkeypatch.setenv("CGPTB_HOSTPIPE_SLABS", slabs)
    grid, m, mo = _ops5(g, fields, MOBIUS, "single")
    s5 = np.ascontiguousarray(fields["src5"].astype(np.complex64))
    out = np.zeros_like(s5)
    m.Dhop_host(out, s5)
    assert rel(out, mo.Dhop(s5)) < TOL["single"]
    out2 = np.zeros_like(s5)
    m.adj().Dhop_host(out2, s5)
    assert rel(out2, mo.Dhop(s5, dag=True)) < TOL["single"]
    # same numbers as the lattice path (identical kernels, identical summation order)
    ref = from_spinor(g(m.Dhop * to_spinor(g, m.F_grid, s5)), s5)
    assert np.array_equal(out, ref)
    # a second call reuses the staging buffers
    out[:] = 0
    m.Dhop_host(out, s5)
    assert np.array_equal(out, ref)
    with pytest.raises(RuntimeError):
        g.cgpt.apply_fermion_operator_host(m.interface.obj, 3001, s5.ctypes.data, out.ctypes.data, 16)


def test_dhop_host_buffers_double(g, fields):
    """double precision has no pipelined path: the host call must still give the oracle's answer to 1e-12"""
    grid, m, mo = _ops5(g, fields, MOBIUS, "double")
    s5 = np.ascontiguousarray(fields["src5"].astype(np.complex128))
    out = np.zeros_like(s5)
    m.Dhop_host(out, s5)
    assert rel(out, mo.Dhop(s5)) < TOL["double"]


# ---------------------------------------------------------------------------------------------------------
# TMA sweep kernel (dslash_tma.cu): asymmetric lattices, several Ls, work schedules that put many items and
# time ranges on one CTA; against the oracle and against the L1 kernels (CGPTB_NO_TMA)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dims,Ls", [([16, 8, 12, 8], 8), ([8, 12, 8, 6], 4), ([8, 8, 8, 16], 16), ([8, 16, 8, 10], 12), ([8, 8, 8, 6], 24)])
def test_tma_sweep_kernel_geometries(g, dims, Ls, monkeypatch):
    rng = oracle_random("tma" + str(dims))
    U = qcd.gauge_random(rng, dims, scale=0.8)
    params = dict(mass=0.08, M5=1.8, b=1.5, c=0.5, Ls=Ls, boundary_phases=[1.0, -1.0, np.exp(0.3j), -1.0])
    grid = g.grid(dims, g.single)
    m = g.qcd.fermion.mobius(to_links(g, grid, U), dict(params))
    mo = qcd.mobius([u.astype(np.complex64) for u in U], **params)
    s5 = rng.cnormal([Ls] + dims, (4, 3)).astype(np.complex64)
    src = to_spinor(g, m.F_grid, s5)
    for dag in (False, True):
        ref = mo.Dhop(s5, dag=dag)
        op = m.Dhop.adj() if dag else m.Dhop
        monkeypatch.setenv("CGPTB_NO_TMA", "1")
        l1 = from_spinor(g(op * src), s5)
        monkeypatch.delenv("CGPTB_NO_TMA")
        assert rel(l1, ref) < TOL["single"]
        # default: one chunk of four s-slices per CTA; CGPTB_TMA_G = 2 / 3 chunks per CTA (ignored unless it divides Ls / 4),
        # CGPTB_TMA_GRID few CTAs (many items per CTA), CGPTB_TMA_SCHED=0 fixed time ranges of CGPTB_TMA_TRL slices
        for env in ({}, {"CGPTB_TMA_GRID": "1", "CGPTB_TMA_G": "3"}, {"CGPTB_TMA_GRID": "5", "CGPTB_TMA_G": "3"}, {"CGPTB_TMA_GRID": "7", "CGPTB_TMA_G": "2"},
                    {"CGPTB_TMA_G": "3"}, {"CGPTB_TMA_G": "2"},
                    {"CGPTB_TMA_SCHED": "0", "CGPTB_TMA_GRID": "1", "CGPTB_TMA_TRL": "2", "CGPTB_TMA_G": "3"},
                    {"CGPTB_TMA_SCHED": "0", "CGPTB_TMA_GRID": "5", "CGPTB_TMA_TRL": "4"},
                    {"CGPTB_TMA_SCHED": "0", "CGPTB_TMA_GRID": "3", "CGPTB_TMA_TRL": "1", "CGPTB_TMA_G": "2"}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got = from_spinor(g(op * src), s5)
            for k in env:
                monkeypatch.delenv(k)
            assert rel(got, ref) < TOL["single"], (dag, env)
            assert rel(got, l1) < 2e-6, (dag, env)  # only the summation order of the eight hops differs
    # checkerboarded entry (DhopEO on half fields: component planes of half the stride)
    e = qcd.eo_ops(mo)
    for cb in (g.even, g.odd):
        half = to_spinor(g, m.F_grid_eo, s5, cb)
        out = from_spinor(g(m.DhopEO * half), s5)
        assert rel(out, e.proj(mo.Dhop(e.proj(s5, cb.tag)), 1 - cb.tag)) < TOL["single"]


# ---------------------------------------------------------------------------------------------------------
# g.random / g.qcd.gauge.random / g.qcd.gauge.plaquette of the product against the reference's known answers
# (/root/reference/tests/random/simple.py:18-31) and the oracle
# ---------------------------------------------------------------------------------------------------------
def test_product_gauge_random_and_plaquette_kat(g):
    rng = g.random("block_seed_string_13")
    orng = oracle_random("block_seed_string_13")
    for dims, precision, ref, scale, tol in [
        ([8, 4, 4, 4], "double", -0.00014108397456619623, 10, 1e-14),
        ([8, 4, 4, 4], "single", -0.00014108397456619623, 10, 1e-7),
        ([8, 8, 4, 8], "double", 0.38723058417632267, 2, 1e-14),
    ]:
        grid = g.grid(dims, prec_of(g, precision))
        U = g.qcd.gauge.random(grid, rng, scale=scale)
        P = g.qcd.gauge.plaquette(U)
        assert abs(P - ref) < tol, (dims, precision, P)
        Uo = qcd.gauge_random(orng, dims, scale=scale, precision=precision)
        for u, uo in zip(U, Uo):
            assert rel(u[:], sites(uo, 2)) < (1e-13 if precision == "double" else 1e-6)
        assert abs(P - qcd.plaquette(Uo)) < tol
        lt = sum(np.trace(uo, axis1=-2, axis2=-1).real.mean() for uo in Uo) / 12.0
        assert abs(g.qcd.gauge.link_trace(U) - lt) < tol


@pytest.mark.parametrize("precision", ["double", "single"])
def test_product_rng_lattices(g, precision):
    """rng.cnormal / normal / uniform_real into 4d and 5d lattices == oracle stream (draw order as in
    tests/qcd/fermion_operators.py:849-884: U, then sources on F_grid and U_grid)"""
    dims = [8, 8, 8, 16]
    p = prec_of(g, precision)
    grid = g.grid(dims, p)
    rng = g.random("finger_print")
    orng = oracle_random("finger_print")
    tag = None if precision == "double" else precision
    U = g.qcd.gauge.random(grid, rng)
    Uo = qcd.gauge_random(orng, dims, precision=precision)
    tol = 1e-13 if precision == "double" else 1e-6
    for u, uo in zip(U, Uo):
        assert rel(u[:], sites(uo, 2)) < tol
    grid5 = grid.inserted_dimension(0, 12)
    src5 = rng.cnormal(g.vspincolor(grid5))
    dst5 = rng.cnormal(g.vspincolor(grid5))
    src4 = rng.cnormal(g.vspincolor(grid))
    d5 = [12] + dims
    assert rel(src5[:], sites(orng.cnormal(d5, (4, 3), grid_tag=tag), 2)) < tol
    assert rel(dst5[:], sites(orng.cnormal(d5, (4, 3), grid_tag=tag), 2)) < tol
    assert rel(src4[:], sites(orng.cnormal(dims, (4, 3), grid_tag=tag), 2)) < tol
    if precision == "double":
        # the reference's Moebius fingerprint from the product's own random fields (fermion_operators.py:427-442)
        m = g.qcd.fermion.mobius(U, dict(MOBIUS))
        X = g.inner_product(dst5, g(m * src5))
        assert abs(X - (-8693.09425573421 - 4130.7793316734915j)) / abs(X) < 1e-13


# ---------------------------------------------------------------------------------------------------------
# NERSC gauge configurations (g.load / g.save, device munge kernel); reference test: tests/io/io.py:198-206
# ---------------------------------------------------------------------------------------------------------
def test_nersc_roundtrip_and_formats(g, tmp_path):
    from tests.nersc_util import write_nersc

    dims = [8, 4, 4, 6]
    rng = oracle_random("nersc")
    Uo = qcd.gauge_random(rng, dims, scale=0.7)
    grid = g.grid(dims, g.double)
    U = to_links(g, grid, Uo)
    # the reference's own round trip: g.save -> g.load, eps < 1e-14
    fn = str(tmp_path / "ckpoint.0000")
    g.save(fn, U, g.format.nersc(label="test"))
    Up = g.load(fn)
    assert len(Up) == 4 and Up[0].metadata["DATATYPE"] == "4D_SU3_GAUGE_3x3"
    for up, u in zip(Up, U):
        assert (g.norm2(g(up - u)) / g.norm2(u)) ** 0.5 < 1e-14
    # files written independently of the product: every float format, both data types
    for fp, dt, tol in [("IEEE64BIG", "4D_SU3_GAUGE_3x3", 1e-15), ("IEEE64LITTLE", "4D_SU3_GAUGE", 1e-13), ("IEEE32BIG", "4D_SU3_GAUGE", 1e-6),
                        ("IEEE32", "4D_SU3_GAUGE_3x3", 1e-6), ("IEEE64BIG", "4D_SU3_GAUGE", 1e-13)]:
        fn = str(tmp_path / f"cfg_{fp}_{dt}")
        write_nersc(fn, Uo, fp, dt)
        Ul = g.load(fn)
        assert Ul[0].grid.precision is (g.double if "64" in fp else g.single)
        for ul, uo in zip(Ul, Uo):
            assert rel(ul[:], sites(uo, 2)) < tol, (fp, dt)
        assert abs(g.qcd.gauge.plaquette(Ul) - qcd.plaquette(Uo)) < max(tol, 1e-14) * 10
    # corrupted data, wrong checksum, wrong plaquette are refused
    fn = str(tmp_path / "bad_cs")
    write_nersc(fn, Uo, checksum=0x1234)
    with pytest.raises(RuntimeError, match="checksum"):
        g.load(fn)
    fn = str(tmp_path / "bad_plaq")
    write_nersc(fn, Uo, plaquette=0.512345678)  # nine digits: tolerance 1e-7 (a header value "0.5" would accept anything)
    with pytest.raises(RuntimeError, match="plaquette"):
        g.load(fn)
    fn = str(tmp_path / "truncated")
    write_nersc(fn, Uo)
    with open(fn, "r+b") as f:
        f.truncate(os.path.getsize(fn) - 8)
    with pytest.raises(RuntimeError, match="bytes of data"):
        g.load(fn)
    with pytest.raises(NotImplementedError):
        g.load(str(tmp_path / "does_not_exist"))


# ---------------------------------------------------------------------------------------------------------
# multi-rhs Wilson-clover: wilson_clover(n_rhs=...).packed() (tests/qcd/fermion_operators.py:97-132)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,n_rhs", [("double", 4), ("single", 12), ("single", 3)])
def test_wilson_multi_rhs_packed(g, fields, precision, n_rhs):
    """the n_rhs columns go through the stencil as the fifth dimension of one field (in single precision with n_rhs = 12 this is
    the TMA sweep kernel) and must give what the single-rhs operator gives column by column"""
    params = dict(mass=0.123, csw_r=0.6, csw_t=0.6, cF=1.0, xi_0=1.0, nu=1.0, isAnisotropic=False, boundary_phases=[1, 1, 1, -1])
    p = prec_of(g, precision)
    grid = g.grid(DIMS, p)
    U = to_links(g, grid, fields["U"])
    single = g.qcd.fermion.wilson_clover(U, dict(params))
    multi5 = g.qcd.fermion.wilson_clover(U, dict(params, n_rhs=n_rhs))
    multi = multi5.packed()
    wo = qcd.wilson_clover([u.astype(p.complex_dtype) for u in fields["U"]], **params)
    rng = oracle_random("multi rhs")
    cols = [rng.cnormal(DIMS, (4, 3)).astype(p.complex_dtype) for _ in range(n_rhs)]
    test = [to_spinor(g, grid, c) for c in cols]
    tol = 1e-15 if precision == "double" else 2e-6
    for op1, opn, ref in [(single, multi, wo.M), (single.adj(), multi.adj(), wo.Mdag), (single.Dhop, multi5.Dhop.packed(), wo.Dhop),
                          (single.Mdiag, multi5.Mdiag.packed(), wo.Mdiag)]:
        test0 = [op1(t) for t in test]
        test1 = opn(test)
        assert len(test1) == n_rhs
        for i in range(n_rhs):
            eps = (g.norm2(g(test0[i] - test1[i])) / g.norm2(test0[i])) ** 0.5
            assert eps < tol, (i, eps)
            assert rel(from_spinor(test1[i], cols[i]), ref(cols[i])) < TOL[precision]
    # even-odd pieces on the 5d grid: MooeeInv Mooee = 1 for every column
    half = [to_spinor(g, single.F_grid_eo, c, g.odd) for c in cols]
    back = multi5.Mooee.inv().packed()(multi5.Mooee.packed()(half))
    for b, h in zip(back, half):
        assert (g.norm2(g(b - h)) / g.norm2(h)) ** 0.5 < (1e-13 if precision == "double" else 1e-5)


def test_wilson_multi_rhs_block_solve(g):
    """eo2_ne CG on the multi-rhs operator: the 4 columns are one 5d vector (one alpha / beta per iteration for all of them),
    every column must solve its own system  M x_i = b_i  to the tolerance; compared with per-column oracle solves"""
    small = [4, 4, 4, 8]
    n_rhs = 4
    rng = oracle_random("block solve")
    U = qcd.gauge_random(rng, small, scale=0.5)
    grid = g.grid(small, g.double)
    Ug = to_links(g, grid, U)
    params = dict(mass=0.2, csw_r=1.1, csw_t=1.3, xi_0=1.0, nu=1.0, isAnisotropic=False, boundary_phases=[1.0, 1.0, 1.0, -1.0])
    op5 = g.qcd.fermion.wilson_clover(Ug, dict(params, n_rhs=n_rhs))
    op1 = g.qcd.fermion.wilson_clover(Ug, dict(params))
    oo = qcd.wilson_clover(U, **params)
    cols = [rng.cnormal(small, (4, 3)) for _ in range(n_rhs)]
    inv = g.algorithms.inverter
    cg = inv.cg(eps=1e-10, maxiter=1000)
    slv = inv.preconditioned(g.qcd.fermion.preconditioner.eo2_ne(), cg)(op5).packed()
    src = [to_spinor(g, grid, c) for c in cols]
    dst = slv(src)
    assert len(cg.history) > 5
    for i in range(n_rhs):
        ref, _ = qcd.solve_eo2_ne(oo, cols[i], 1e-10, 1000)
        assert rel(from_spinor(dst[i], cols[i]), ref) < 1e-8, i
        r = g(op1 * dst[i] - src[i])
        assert (g.norm2(r) / g.norm2(src[i])) ** 0.5 < 1e-8


# ---------------------------------------------------------------------------------------------------------
# zMoebius (complex, s-dependent coefficients): every opcode against the oracle, the reference's fingerprints
# (tests/qcd/fermion_operators.py:397-426) from the product's own random fields, eo2_ne CG
# ---------------------------------------------------------------------------------------------------------
ZMOBIUS = dict(
    mass=0.08, M5=1.8, b=1.0, c=0.0, boundary_phases=[1.0, 1.0, 1.0, -1.0],
    omega=[0.17661651536320583 + 1j * (0.14907774771612217), 0.23027432016909377 + 1j * (-0.03530801572584271),
           0.3368765581549033 + 1j * (0), 0.7305711010541054 + 1j * (0), 1.1686138337986505 + 1j * (0.3506492418109086),
           1.1686138337986505 + 1j * (-0.3506492418109086), 0.994175013717952 + 1j * (0), 0.5029903152251229 + 1j * (0),
           0.23027432016909377 + 1j * (0.03530801572584271), 0.17661651536320583 + 1j * (-0.14907774771612217)])


@pytest.mark.parametrize("precision", ["double", "single"])
def test_zmobius(g, precision):
    p = prec_of(g, precision)
    tol = TOL[precision]
    grid = g.grid(DIMS, p)
    rng = g.random("finger_print")
    orng = oracle_random("finger_print")
    tag = None if precision == "double" else precision
    U = g.qcd.gauge.random(grid, rng)
    Uo = qcd.gauge_random(orng, DIMS, precision=precision)
    m = g.qcd.fermion.zmobius(U, dict(ZMOBIUS))
    mo = qcd.zmobius(Uo, **ZMOBIUS)
    src5, dst5 = rng.cnormal(g.vspincolor(m.F_grid)), rng.cnormal(g.vspincolor(m.F_grid))
    src4 = rng.cnormal(g.vspincolor(grid))
    d5 = [10] + DIMS
    s5 = orng.cnormal(d5, (4, 3), grid_tag=tag).astype(p.complex_dtype)
    orng.cnormal(d5, (4, 3), grid_tag=tag)
    s4 = orng.cnormal(DIMS, (4, 3), grid_tag=tag).astype(p.complex_dtype)
    assert rel(src5[:], sites(s5, 2)) < (1e-13 if precision == "double" else 1e-6)
    for i, (op, ref) in enumerate([
        (m, mo.M(s5)), (m.adj(), mo.Mdag(s5)), (m.Mdiag, mo.Mdiag(s5)), (m.Dhop, mo.Dhop(s5)),
        (m.Dminus, mo.Dminus(s5)), (m.Dminus.adj(), mo.Dminus(s5, dag=True)),
    ]):
        assert rel(from_spinor(g(op * src5), s5), ref) < tol, i
    assert rel(from_spinor(g(m.ImportPhysicalFermionSource * src4), s5), mo.ImportPhysicalFermionSource(s4)) < tol
    assert rel(from_spinor(g(m.ExportPhysicalFermionSolution * src5), s4), mo.ExportPhysicalFermionSolution(s5)) < tol
    e = qcd.eo_ops(mo)
    for cb in [g.even, g.odd]:
        half_np = e.proj(s5, cb.tag)
        half = to_spinor(g, m.F_grid_eo, s5, cb)
        for j, (op, ref) in enumerate([
            (m.Meooe, e.Meooe(half_np, cb.tag)), (m.Meooe.adj(), e.Meooe(half_np, cb.tag, dag=True)),
            (m.Mooee, e.Mooee(half_np)), (m.Mooee.adj(), e.Mooee(half_np, dag=True)),
            (m.Mooee.inv(), e.MooeeInv(half_np)), (m.Mooee.adj().inv(), e.MooeeInv(half_np, dag=True)),
        ]):
            assert rel(from_spinor(g(op * half), s5), ref) < tol, (cb, j)
    if precision == "double":
        for op, s, ref in [(m, src5, -2424.048033434305 + 10557.661684178218j), (m.Mdiag, src5, 2643.396577965267 + 6550.259431381319j),
                           (m.ImportPhysicalFermionSource, src4, 4064.7879718582053 - 1357.0856808000196j)]:
            X = g.inner_product(dst5, g(op * s))
            assert abs(X - ref) / abs(ref) < 1e-13
        # eo2_ne CG: true residual of the full system
        inv = g.algorithms.inverter
        cg = inv.cg(eps=1e-8, maxiter=2000)
        slv = inv.preconditioned(g.qcd.fermion.preconditioner.eo2_ne(), cg)(m)
        x = g(slv * src5)
        r = g(m * x - src5)
        assert (g.norm2(r) / g.norm2(src5)) ** 0.5 < 1e-6
        # the production preconditioner of zMoebius runs (applications/propagator/rbc/24D.py:16-61): kappa-rescaled Schur complement
        kap = m.kappa()
        half = to_spinor(g, m.F_grid_eo, s5, g.odd)
        back = g(kap.inv() * kap * half)
        assert (g.norm2(g(back - half)) / g.norm2(half)) ** 0.5 < 1e-14
        ks = 1.0 / (2.0 * (mo.bs * (4.0 - 1.8) + 1.0))
        ref = e.proj(s5, 1) * ks.reshape(-1, 1, 1)
        assert rel(from_spinor(g(kap * half), s5), ref) < 1e-14
        cg2 = inv.cg(eps=1e-8, maxiter=2000)
        slv2 = inv.preconditioned(g.qcd.fermion.preconditioner.eo2_kappa_ne(), cg2)(m)
        x2 = g(slv2 * src5)
        r = g(m * x2 - src5)
        assert (g.norm2(r) / g.norm2(src5)) ** 0.5 < 1e-6
        assert rel(x2[:], x[:]) < 1e-6


# ---------------------------------------------------------------------------------------------------------
# g.gamma: spin matrices on fields; gamma_5 hermiticity and the twisted-mass identity of the reference's test
# (tests/qcd/fermion_operators.py:80-95)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["double", "single"])
def test_gamma_algebra_on_fields(g, fields, precision):
    grid, w, wo = _ops4(g, fields, CLOVER, precision)
    tol = TOL[precision]
    cdt = grid.precision.complex_dtype
    s_np = fields["srcw"].astype(cdt)
    src = to_spinor(g, grid, s_np)
    for key in [0, 1, 2, 3, 5, "I", (0, 1), (2, 3)]:
        mat = qcd.gamma[key] if not isinstance(key, tuple) else qcd.sigma(*key)
        got = from_spinor(g(g.gamma[key] * src), s_np)
        assert rel(got, qcd.spin_mul(mat.astype(cdt), s_np)) < tol, key
    Pp, Pm = 0.5 * (g.gamma["I"] + g.gamma[5]), 0.5 * (g.gamma["I"] - g.gamma[5])
    assert rel(from_spinor(g(Pp * src + Pm * src), s_np), s_np) < tol
    assert rel(from_spinor(g(Pp * Pm * src), s_np) + s_np, s_np) < tol  # P+ P- = 0
    # gamma_5 hermiticity of the Wilson-clover operator: g5 M g5 = M^dag
    a = g(g.gamma[5] * w * g.gamma[5] * src)
    b = g(w.adj() * src)
    assert (g.norm2(g(a - b)) / g.norm2(b)) ** 0.5 < tol
    # twisted mass: M_tm = M_wilson + i mu gamma_5
    grid, wt, _ = _ops4(g, fields, dict(mass=-1.8, mu=0.3, boundary_phases=[1.0, 1.0, 1.0, -1.0]), precision)
    grid, w0, _ = _ops4(g, fields, dict(mass=-1.8, csw_r=0.0, csw_t=0.0, xi_0=1.0, nu=1.0, isAnisotropic=False,
                                        boundary_phases=[1.0, 1.0, 1.0, -1.0]), precision)
    lhs = g(w0 * src + 0.3j * (g.gamma[5] * src))
    rhs = g(wt * src)
    assert (g.norm2(g(lhs - rhs)) / g.norm2(rhs)) ** 0.5 < tol
    # 5d fields too
    grid5, m, mo = _ops5(g, fields, MOBIUS, precision)
    s5 = fields["src5"].astype(cdt)
    got = from_spinor(g(g.gamma[5] * to_spinor(g, m.F_grid, s5)), s5)
    assert rel(got, qcd.spin_mul(qcd.gamma[5].astype(cdt), s5)) < tol


def test_madwf(g):
    """MADWF (tests/qcd/domain_wall.py:140-164): Moebius Ls = 12 propagator through the zMoebius Ls = 10 inner operator; the
    approximation is close to the direct solve and, wrapped into defect correction, converges to it"""
    small = [4, 4, 4, 8]
    grid = g.grid(small, g.double)
    rng = g.random("madwf")
    U = g.qcd.gauge.random(grid, rng, scale=0.3)
    bc = [1.0, 1.0, 1.0, 1.0]
    qm = g.qcd.fermion.mobius(U, dict(mass=0.08, M5=1.8, b=1.5, c=0.5, Ls=12, boundary_phases=bc))
    qz = g.qcd.fermion.zmobius(U, dict(ZMOBIUS, boundary_phases=bc))
    inv = g.algorithms.inverter
    pc = g.qcd.fermion.preconditioner
    slv_5d = inv.preconditioned(pc.eo2_ne(), inv.cg(eps=1e-7, maxiter=2000))
    slv_e = inv.preconditioned(pc.eo2_ne(), inv.cg(eps=1e-10, maxiter=2000))
    src = rng.cnormal(g.vspincolor(grid))
    direct = g(qm.propagator(slv_e) * src)
    madwf = g(qm.propagator(pc.mixed_dwf(slv_5d, slv_5d, qz)) * src)
    eps2 = g.norm2(g(madwf - direct)) / g.norm2(direct)
    assert eps2 < 5e-3, eps2
    madwf_dc = g(qm.propagator(inv.defect_correcting(pc.mixed_dwf(slv_5d, slv_5d, qz), eps=1e-8, maxiter=20)) * src)
    eps2 = g.norm2(g(madwf_dc - direct)) / g.norm2(direct)
    assert eps2 < 1e-12, eps2
    # separate / merge round trip
    x5 = rng.cnormal(g.vspincolor(qm.F_grid))
    parts = g.separate(x5)
    assert len(parts) == 12
    assert g.norm2(g(g.merge(parts) - x5)) == 0.0


def test_mobius_correlator_golden_from_seed(g):
    """seed string in, physics out, no oracle in between: the reference's domain-wall test (tests/qcd/domain_wall.py:14-20,26-35,
    116-137,252-283) through the drop-in API -- g.random("test") -> gauge.random(scale=2) on 8^4 in double, converted to single,
    Moebius Ls = 12, point source at [0,1,0,0], eo2_ne CG (eps 1e-8, maxiter 1000), pion correlator against the 8 golden values"""
    correlator_ref = [0.5534145832061768, 0.2355920523405075, 0.08622127771377563, 0.05764763802289963, 0.05238068848848343,
                      0.057377591729164124, 0.08141942322254181, 0.21931196749210358]
    rng = g.random("test")
    U = g.qcd.gauge.random(g.grid([8, 8, 8, 8], g.double), rng, scale=2.0)
    U = g.convert(U, g.single)
    grid = U[0].grid
    qm = g.qcd.fermion.mobius(U, dict(mass=0.08, M5=1.8, b=1.5, c=0.5, Ls=12, boundary_phases=[1.0, 1.0, 1.0, 1.0]))
    src = g.mspincolor(grid)
    g.create.point(src, [0, 1, 0, 0])
    inv = g.algorithms.inverter
    pc = g.qcd.fermion.preconditioner
    cg_e = inv.cg({"eps": 1e-8, "maxiter": 1000})
    dst = g(qm.propagator(inv.preconditioned(pc.eo2_ne(), cg_e)) * src)
    correlator = g.slice(g.trace(dst * g.adj(dst)), 3)
    eps = sum((correlator[t].real - correlator_ref[t]) ** 2.0 for t in range(8)) ** 0.5 / 8
    assert eps < 1e-5, (eps, [c.real for c in correlator])


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configs[0] and configs[1] verbatim
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["double", "single"])
def test_config0_wilson_clover_dslash_16(g, precision):
    """BASELINE.json configs[0]: the loop of /root/reference/benchmarks/wilson_clover_dslash.py:9-67 on 16^4 (rng "benchmark" with
    the fast engine, gauge.random scale 0.5, kwargs call form, cnormal source, 5 + N applications of Dhop), result against the
    C restatement of the hopping term on the same links"""
    from oracle import cref

    dims = [16, 16, 16, 16]
    prec = prec_of(g, precision)
    rng = g.random("benchmark", "vectorized_ranlux24_24_64")
    grid = g.grid(dims, prec)
    qm = g.qcd.fermion.wilson_clover(
        g.qcd.gauge.random(grid, rng, scale=0.5),
        mass=0.08, csw_r=1.0, csw_t=1.0, xi_0=1.0, nu=1, isAnisotropic=False, boundary_phases=[1, 1, 1, -1], n_rhs=1,
    )
    src = g.vspincolor(qm.F_grid)
    dst = g.vspincolor(qm.F_grid)
    rng.cnormal(src)
    for n in range(5):
        qm.Dhop.mat(dst, src)
    N = 10
    t0 = g.time()
    for n in range(N):
        qm.Dhop.mat(dst, src)
    g.cgpt.accelerator_barrier()
    t1 = g.time()
    flops = 8 * 3 * (7 + 16 * 3) * src.grid.gsites * N
    g.message(f"wilson_clover_dslash 16^4 {precision}: {flops / (t1 - t0) / 1e9:.1f} GFlops/s")
    # antiperiodic in time: the phase multiplies U_t on the last time slice (lib/gpt/core/covariant.py:29-37)
    V = np.stack([u[:] for u in qm.U]).reshape(4, 16, 16, 16, 16, 3, 3).copy()
    V[3, 15] *= -1.0
    ref = cref.dhop(dims, 0, V.reshape(4, -1, 3, 3), src[:])
    assert rel(dst[:], ref) < TOL[precision]
    # the 12-column form of the same benchmark (--n_rhs 12) shares every link load
    if precision == "single":
        qm12 = g.qcd.fermion.wilson_clover(qm.U, mass=0.08, csw_r=1.0, csw_t=1.0, xi_0=1.0, nu=1, isAnisotropic=False,
                                           boundary_phases=[1, 1, 1, -1], n_rhs=12)
        s12, d12 = g.vspincolor(qm12.F_grid), g.vspincolor(qm12.F_grid)
        rng.cnormal(s12)
        qm12.Dhop.mat(d12, s12)
        ref = cref.dhop(dims, 12, V.reshape(4, -1, 3, 3), s12[:])
        assert rel(d12[:], ref) < TOL[precision]


def test_config1_readme_example(g):
    """BASELINE.json configs[1]: README.md:132-167 line by line (double 8^4, g.random("seed text"), Moebius Ls = 24 with b = 1,
    c = 0 in the kwargs call form, eo2_ne CG eps 1e-4, point source, 12 columns, pion correlator).  Checked against the oracle:
    two of the twelve columns solved by the numpy restatement of the same solver stack (solution and iteration count), every
    column through the true residual, and the correlator recomputed on the host from the propagator."""
    grid = g.grid([8, 8, 8, 8], g.double)
    rng = g.random("seed text")
    U = g.qcd.gauge.random(grid, rng)
    fermion = g.qcd.fermion.mobius(U, mass=0.1, M5=1.8, b=1.0, c=0.0, Ls=24, boundary_phases=[1, 1, 1, -1])
    inv = g.algorithms.inverter
    pc = g.qcd.fermion.preconditioner
    cg = inv.cg(eps=1e-4, maxiter=1000)
    slv_5d = inv.preconditioned(pc.eo2_ne(), cg)
    fermion_propagator = fermion.propagator(slv_5d)
    src = g.mspincolor(U[0].grid)
    g.create.point(src, [0, 0, 0, 0])
    prop = g(fermion_propagator * src)
    correlator = g.slice(g.trace(prop * g.adj(prop)), 3)
    g.message(correlator)
    assert len(correlator) == 8 and all(c.real > 0 and abs(c.imag) < 1e-12 * c.real for c in correlator)
    p = prop[:]  # [site, si, sj, ca, cb]
    host = (np.abs(p.reshape(8, 8 * 8 * 8, -1)) ** 2).sum(axis=(1, 2))
    assert np.max(np.abs(host - np.array([c.real for c in correlator])) / host) < 1e-12
    # oracle: the numpy restatement of the same solver stack on the same links needs minutes per column, so its results are a
    # committed fixture (tests/golden/make_readme_vectors.py): iteration count per column, correlator, propagator on 24 sites
    import json

    with open(os.path.join(os.path.dirname(__file__), "golden", "readme_mobius_ls24.json")) as f:
        gold = json.load(f)
    orng = oracle_random("seed text")
    Uo = qcd.gauge_random(orng, [8, 8, 8, 8])
    assert rel(np.stack([u[:] for u in U]), np.stack([sites(u, 2) for u in Uo])) < 1e-13  # the fixture's links are these links
    assert rel([c.real for c in correlator], gold["correlator"]) < 1e-6  # eps = 1e-4 solves: agreement is O(eps * rounding path)
    pp = p.reshape(8, 8, 8, 8, 4, 4, 3, 3)  # [t,z,y,x,si,sj,ca,cb]
    for key, val in gold["sample"].items():
        x, y, z, t = (int(v) for v in key.split(","))
        want = np.array(val).view(np.float64).reshape(12, 12, 2)
        want = (want[..., 0] + 1j * want[..., 1]).reshape(4, 3, 4, 3).transpose(0, 2, 1, 3)  # [si,sj,ca,cb]
        assert rel(pp[t, z, y, x], want) < 1e-7, key
    assert len(cg.history) == gold["iterations"][11]  # cg.history belongs to the last solve: column 11
    # every column: true residual of the 5d system behind it is not checked by the reference either; the 4d check is that
    # D_ov-like relation M5d x = Import(src) holds to the CG tolerance for one more column
    s4 = g.vspincolor(U[0].grid)
    s4[:] = 0
    s4[0, 0, 0, 0] = np.eye(12)[5].reshape(4, 3)
    b5 = g(fermion.ImportPhysicalFermionSource * s4)
    x5 = g(slv_5d(fermion) * b5)
    assert (g.norm2(g(fermion * x5 - b5)) / g.norm2(b5)) ** 0.5 < 1e-3
    assert rel(g(fermion.ExportPhysicalFermionSolution * x5)[:], prop.columns[5][:]) < 1e-12


def test_cg_zero_source_and_not_converged_paths(g, fields):
    """the device loop returns silently for b = 0 (cg.py:67-69) also with fail_if_not_converged, and raises when maxiter is hit"""
    params = dict(MOBIUS, Ls=8, boundary_phases=[1.0, 1.0, 1.0, -1.0])
    grid = g.grid(DIMS, g.single)
    op = g.qcd.fermion.mobius(to_links(g, grid, fields["U"]), dict(params))
    inv = g.algorithms.inverter
    pc = g.qcd.fermion.preconditioner
    src = g.vspincolor(op.F_grid)
    src[:] = 0
    for fused in (True, False):
        if not fused:
            os.environ["GPT_B200_NO_FUSED"] = "1"
        try:
            cg = inv.cg(eps=1e-6, maxiter=100, fail_if_not_converged=True)
            dst = g(inv.preconditioned(pc.eo2_ne(), cg)(op) * src)
            assert g.norm2(dst) == 0.0 and cg.history == []
            g.random("x").cnormal(src)
            cg = inv.cg(eps=1e-6, maxiter=3, fail_if_not_converged=True)
            with pytest.raises(ValueError):
                g(inv.preconditioned(pc.eo2_ne(), cg)(op) * src)
            assert len(cg.history) == 3
            src[:] = 0
        finally:
            os.environ.pop("GPT_B200_NO_FUSED", None)


def test_open_bc_dhop_host_and_long_linear_combination(g, fields):
    """(a) Dhop_host on an operator with open boundary conditions equals Dhop (boundary slices cleared);
    (b) a linear combination with more than 8 terms in which dst itself appears beyond the first chunk"""
    grid = g.grid(DIMS, g.single)
    p = dict(kappa=0.135, csw_r=1.978, csw_t=1.978, cF=1.3, xi_0=1, nu=1, isAnisotropic=False, boundary_phases=[1.0, 1.0, 1.0, 0.0], n_rhs=4)
    op = g.qcd.fermion.wilson_clover(to_links(g, grid, fields["U"]), dict(p))
    src = g.vspincolor(op.F_grid)
    g.random("open").cnormal(src)
    ref = g(op.Dhop * src)[:]
    h_in = np.ascontiguousarray(src[:])
    h_out = np.zeros_like(h_in)
    op.Dhop_host(h_out, h_in)
    assert np.array_equal(h_out, ref)
    a = ref.reshape(16, 8 * 8 * 8, 4, -1)
    assert np.all(a[0] == 0) and np.all(a[15] == 0) and np.any(a[1] != 0)
    # (b)
    vs = [g.vspincolor(grid) for _ in range(11)]
    g.random("lc").cnormal(vs)
    c = [0.1 * (k + 1) - 0.05j * k for k in range(11)]
    want = sum(ck * v[:].astype(np.complex128) for ck, v in zip(c, vs))
    dst = vs[9]
    g.cgpt.lattice_lc(dst.obj, False, c, [v.obj for v in vs])
    assert rel(dst[:], want) < 1e-6


# ---------------------------------------------------------------------------------------------------------
# two-row SU(3) link compression (an option of this package: params link_compression=12): the stencil's link tables keep rows
# 0 and 1 and the U(1) factor of the stored link (-c_mu/2 x boundary phase); row 2 is rebuilt in registers.  Against the
# 18-real path and the oracle, for the TMA sweep kernel (single, Ls % 4 == 0), the generic kernel (double; single where the
# sweep kernel does not apply), complex boundary phases, anisotropy and open boundary conditions (vanishing links)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["double", "single"])
def test_link_compression(g, fields, precision):
    prec = prec_of(g, precision)
    tol = TOL[precision]
    close = 1e-13 if precision == "double" else 3e-6  # compressed against uncompressed: rounding of the rebuilt row only
    grid = g.grid(DIMS, prec)
    cdt = prec.complex_dtype
    U = fields["U"]
    Ug = to_links(g, grid, U)
    Uo = [u.astype(cdt) for u in U]
    phases = [1.0, -1.0, np.exp(0.3j), -1.0]
    # Moebius: every opcode that contains the stencil
    for Ls in (8, 6):
        params = dict(mass=0.08, M5=1.8, b=1.5, c=0.5, Ls=Ls, boundary_phases=phases)
        m18 = g.qcd.fermion.mobius(Ug, dict(params))
        m12 = g.qcd.fermion.mobius(Ug, dict(params, link_compression=12))
        mo = qcd.mobius(Uo, **params)
        s5 = oracle_random("cmp" + str(Ls)).cnormal([Ls] + DIMS, (4, 3)).astype(cdt)
        src = to_spinor(g, m18.F_grid, s5)
        for tag, a, b, ref in [("Dhop", m12.Dhop, m18.Dhop, mo.Dhop(s5)), ("DhopDag", m12.Dhop.adj(), m18.Dhop.adj(), mo.Dhop(s5, dag=True)),
                               ("M", m12, m18, mo.M(s5)), ("Mdag", m12.adj(), m18.adj(), mo.Mdag(s5))]:
            got = from_spinor(g(a * src), s5)
            assert rel(got, ref) < tol, (Ls, tag)
            assert rel(got, from_spinor(g(b * src), s5)) < close, (Ls, tag)
        e = qcd.eo_ops(mo)
        for cb in (g.even, g.odd):
            half = to_spinor(g, m12.F_grid_eo, s5, cb)
            out = from_spinor(g(m12.Meooe * half), s5)
            assert rel(out, e.Meooe(e.proj(s5, cb.tag), cb.tag)) < tol
    # Wilson-clover: anisotropic with a clover term, and open boundary conditions in time
    s4 = fields["src4"].astype(cdt)
    src4 = to_spinor(g, grid, s4)
    for p in (dict(CLOVER, boundary_phases=phases),
              dict(kappa=0.135, csw_r=1.978, csw_t=1.978, cF=1.3, xi_0=1, nu=1, isAnisotropic=False, boundary_phases=[1.0, 1.0, 1.0, 0.0])):
        w18 = g.qcd.fermion.wilson_clover(Ug, dict(p))
        w12 = g.qcd.fermion.wilson_clover(Ug, dict(p, link_compression=12))
        wo = qcd.wilson_clover(Uo, **p)
        for tag, a, b, ref in [("M", w12, w18, wo.M(s4)), ("Mdag", w12.adj(), w18.adj(), wo.Mdag(s4)), ("Dhop", w12.Dhop, w18.Dhop, wo.Dhop(s4))]:
            got = from_spinor(g(a * src4), s4)
            assert rel(got, ref) < tol, tag
            assert rel(got, from_spinor(g(b * src4), s4)) < close, tag
    # a solve on the compressed operator: same iteration count as on the 18-real one
    if precision == "double":
        params = dict(mass=0.1, M5=1.8, b=1.5, c=0.5, Ls=8, boundary_phases=[1.0, 1.0, 1.0, -1.0])
        inv = g.algorithms.inverter
        pc = g.qcd.fermion.preconditioner
        s5 = oracle_random("cmp solve").cnormal([8] + DIMS, (4, 3))
        its = []
        sols = []
        for extra in ({}, {"link_compression": 12}):
            m = g.qcd.fermion.mobius(Ug, dict(params, **extra))
            cg = inv.cg(eps=1e-8, maxiter=500)
            sols.append(g(inv.preconditioned(pc.eo2_ne(), cg)(m) * to_spinor(g, m.F_grid, s5))[:])
            its.append(len(cg.history))
        assert its[0] == its[1] and rel(sols[1], sols[0]) < 1e-9
    with pytest.raises(Exception):
        g.qcd.fermion.mobius(Ug, dict(mass=0.1, M5=1.8, b=1.5, c=0.5, Ls=8, boundary_phases=phases, link_compression=8))


# ---------------------------------------------------------------------------------------------------------
# generic matrix-vector stencil (cgpt.stencil_matrix_vector_*): the covariant Laplacian of
# /root/reference/benchmarks/stencil.py:91-119 (eight one-link terms, links and pre-shifted adjoint links as separate matrix
# fields), the three-point version of tests/core/stencil.py:168-215 (adjoint flag + shifted matrix point), a two-link term with a
# diagonal shift and two independent code blocks -- against numpy on the oracle's fields
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["double", "single"])
def test_stencil_matrix_vector(g, fields, precision):
    prec = prec_of(g, precision)
    tol = TOL[precision]
    cdt = prec.complex_dtype
    grid = g.grid(DIMS, prec)
    U = [u.astype(cdt) for u in fields["U"]]
    Ug = to_links(g, grid, fields["U"])
    s4 = fields["src4"].astype(cdt)
    src = to_spinor(g, grid, s4)

    def mul(m, f):
        return np.einsum("...ab,...sb->...sa", m, f)

    evec = [(1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)]
    nevec = [tuple(-x for x in y) for y in evec]
    # (a) benchmarks/stencil.py: laplace with U and UdagShift = adj(cshift(U, mu, -1)) as eight matrix fields
    Udag = [qcd.adj(qcd.shift(U[mu], mu, -1)) for mu in range(4)]
    Udag_g = to_links(g, grid, Udag)
    _X, _Xp, _Xm = 0, [1, 2, 3, 4], [5, 6, 7, 8]
    code = [(0, 1, _X, -1, -8.0, [])]
    for mu in range(4):
        code.append((0, 1, _Xp[mu], 0, 1.0, [(mu, _X, 0)]))
        code.append((0, 1, _Xm[mu], 0, 1.0, [(4 + mu, _X, 0)]))
    st = g.stencil.matrix_vector(Ug[0], src, [(0, 0, 0, 0)] + evec + nevec, code)
    dst = g.lattice(src)
    st(Ug + Udag_g, [dst, src])
    ref = -8.0 * s4
    for mu in range(4):
        ref = ref + mul(U[mu], qcd.shift(s4, mu, +1)) + mul(Udag[mu], qcd.shift(s4, mu, -1))
    assert rel(from_spinor(dst, s4), ref) < tol
    # (b) tests/core/stencil.py: the backward term through the adjoint flag and a shifted matrix point
    for mu in range(4):
        st = g.stencil.matrix_vector(Ug[0], src, [(0, 0, 0, 0), evec[mu], nevec[mu]], [
            {"target": 0, "source": 1, "source_point": 0, "accumulate": -1, "weight": -2.0, "factor": []},
            {"target": 0, "source": 1, "source_point": 1, "accumulate": 0, "weight": 1.0, "factor": [(mu, 0, 0)]},
            {"target": 0, "source": 1, "source_point": 2, "accumulate": 0, "weight": 1.0, "factor": [(mu, 2, 1)]},
        ])
        st(Ug, [dst, src])
        ref = -2.0 * s4 + mul(U[mu], qcd.shift(s4, mu, +1)) + mul(Udag[mu], qcd.shift(s4, mu, -1))
        assert rel(from_spinor(dst, s4), ref) < tol, mu
    # (c) two factors, a diagonal point, a complex weight, and two independent blocks writing two targets
    diag = (1, 0, -1, 0)
    st = g.stencil.matrix_vector(Ug[0], src, [(0, 0, 0, 0), evec[0], diag, evec[3]], [
        (0, 2, 2, -1, 0.5 - 0.25j, [(0, 0, 0), (2, 1, 1)]),   # dst0(x) = w U_0(x) U_2^dag(x + 0) src(x + diag)
        (1, 2, 3, 2, 2.0, [(3, 0, 0)]),                         # dst1(x) = 2 U_3(x) src(x + t) + src(x)
    ], code_parallel_block_size=1)
    d0, d1 = g.lattice(src), g.lattice(src)
    st(Ug, [d0, d1, src])
    sd = qcd.shift(qcd.shift(s4, 0, +1), 2, -1)
    ref0 = (0.5 - 0.25j) * mul(U[0], mul(qcd.adj(qcd.shift(U[2], 0, +1)), sd))
    ref1 = 2.0 * mul(U[3], qcd.shift(s4, 3, +1)) + s4
    assert rel(from_spinor(d0, s4), ref0) < tol and rel(from_spinor(d1, s4), ref1) < tol
    with pytest.raises(Exception):
        g.stencil.matrix_vector(Ug[0], src, [(0, 0, 0, 0)], [(0, 1, 0, -1, 1.0, []), (0, 1, 0, -1, 1.0, [])], code_parallel_block_size=3)


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json's full size (configs[2]: Moebius 32^3 x 64, Ls = 12, single; 148 CTAs x 880 work items of the TMA sweep
# kernel) through size-independent properties -- the oracle does this lattice only on sampled sites (bench.py's parity
# field): adjointness <y, D x> = <D^dag y, x>, linearity, |Mpc x|^2 = <x, Mpc^dag Mpc x>, even-odd decomposition
# Dhop = DhopEO + DhopOE, and the chunk split (CGPTB_TMA_G) leaving the result unchanged up to the order of two additions
# ---------------------------------------------------------------------------------------------------------
def test_full_size_properties(g, monkeypatch):
    dims, Ls = [32, 32, 32, 64], 12
    grid = g.grid(dims, g.single)
    rng = g.random("full size", "vectorized_ranlux24_24_64")
    U = g.qcd.gauge.random(grid, rng, scale=0.5)
    m = g.qcd.fermion.mobius(U, dict(mass=0.08, M5=1.8, b=1.5, c=0.5, Ls=Ls, boundary_phases=[1.0, 1.0, 1.0, -1.0]))
    x, y = g.vspincolor(m.F_grid), g.vspincolor(m.F_grid)
    rng.cnormal([x, y])
    Dx = g(m.Dhop * x)
    Ddy = g(m.Dhop.adj() * y)
    nx, ny, nDx = g.norm2(x), g.norm2(y), g.norm2(Dx)
    # adjointness (reductions are double precision; the fields are single)
    a, b = g.inner_product(y, Dx), g.inner_product(Ddy, x)
    assert abs(a - b) / (ny * nDx) ** 0.5 < 1e-6
    assert 0.1 < nDx / nx < 100.0
    # linearity
    ca, cb = 0.7 - 0.2j, -1.3 + 0.5j
    z = g(ca * x + cb * y)
    Dz = g(m.Dhop * z)
    Dy = g(m.Dhop * y)
    assert (g.norm2(g(Dz - ca * Dx - cb * Dy)) / g.norm2(Dz)) ** 0.5 < 2e-6
    # Dhop only connects the two parities: the even-odd entries on the two halves reproduce it
    xe, xo = g.vspincolor(m.F_grid_eo), g.vspincolor(m.F_grid_eo)
    g.pick_checkerboard(g.even, xe, x)
    g.pick_checkerboard(g.odd, xo, x)
    full = g.vspincolor(m.F_grid)
    g.set_checkerboard(full, g(m.DhopEO * xe))
    g.set_checkerboard(full, g(m.DhopEO * xo))
    assert g.norm2(g(full - Dx)) == 0.0
    # one and three chunks of the fifth dimension per CTA: the same arithmetic, the two x hops summed in the other order
    monkeypatch.setenv("CGPTB_TMA_G", "3")
    D3 = g(m.Dhop * x)
    monkeypatch.delenv("CGPTB_TMA_G")
    assert (g.norm2(g(D3 - Dx)) / nDx) ** 0.5 < 1e-6
    # Schur complement and its normal equation (the fused even-odd paths): <x, Mpc^dag Mpc x> = |Mpc x|^2
    pc = g.qcd.fermion.preconditioner
    Mpc = pc.eo2()(m).Mpc
    NE = pc.eo2_ne()(m).Mpc
    g.pick_checkerboard(g.odd, xo, x)
    v = g(Mpc * xo)
    w = g(NE * xo)
    assert abs(g.inner_product(xo, w).real - g.norm2(v)) / g.norm2(v) < 1e-5
    assert abs(g.inner_product(xo, w).imag) / g.norm2(v) < 1e-5


def test_random_on_checkerboarded_lattice(g):
    """the reference samples a checkerboarded lattice over its reduced coordinates (x/2, y, z, t) (lib/cgpt/lib/random/parallel.h:
    27-128; used by tests/qcd/fermion_operators.py:649-654): same numbers as a full lattice of the reduced extents"""
    for dims, Ls in [([8, 4, 4, 4], 0), ([8, 8, 4, 4], 4)]:
        grid = g.grid(([Ls] if Ls else []) + dims, g.double).checkerboarded(g.redblack)
        l = g.vspincolor(grid)
        l.checkerboard(g.odd)
        g.random("cb lattice").cnormal(l)
        red = ([Ls] if Ls else []) + [dims[0] // 2] + dims[1:]
        ref = oracle_random("cb lattice").cnormal(red, (4, 3))
        assert rel(l[:], ref.reshape(-1, 4, 3)) < 1e-15


def test_lattice_memory_is_reused(g):
    """deleted lattices hand their device memory to the next lattice of the same size (lattice.cu: no cudaFree / cudaMalloc per
    temporary of the host layer); a lattice that reuses memory is a normal lattice: zero it, fill it, read it back"""
    from gpt_b200 import capi

    grid = g.grid([8, 8, 8, 8], g.double)
    first = [g.vspincolor(grid) for _ in range(40)]  # more than any earlier test can have left in the cache for this size
    ptrs = sorted(capi.lattice_device_ptr(l.obj) for l in first)
    assert len(set(ptrs)) == 40
    del first
    second = [g.vspincolor(grid) for _ in range(40)]
    assert sorted(capi.lattice_device_ptr(l.obj) for l in second) == ptrs
    b, c = second[0], second[1]
    b[:] = 0
    assert g.norm2(b) == 0.0
    g.random("reuse").cnormal([b, c])
    assert g.norm2(b) > 0 and abs(g.inner_product(b, c)) < g.norm2(b)

"""
CPU tests: pin the oracle against every golden vector the reference's own tests hold for the hot path.
  /root/reference/tests/random/simple.py:18-31,69-144        RNG known-answer vectors
  /root/reference/tests/qcd/fermion_operators.py:371-459     operator fingerprints <dst| M |src>
"""
import numpy as np
import pytest

from oracle import qcd
from oracle.rng import random


def test_rng_normal_kat():
    rng = random("block_seed_string_13")
    v = rng.normal([8, 4, 4, 4])

    def at(x, y, z, t):
        return v[t, z, y, x].real

    got = np.array([at(0, 0, 0, 0), at(2, 0, 0, 0), at(0, 2, 0, 0), at(1, 3, 1, 3), at(3, 2, 1, 0)])
    ref = np.array([-0.29101665386129116, -1.4591269443435488, -0.3641310411719848, -0.9454532383815435, 0.4996115272362977])
    assert np.linalg.norm(got - ref) < 1e-14
    for _ in range(1000):
        v = rng.normal([8, 4, 4, 4])
    got = np.array([at(0, 0, 0, 0), at(2, 0, 0, 0), at(0, 2, 0, 0), at(1, 3, 1, 3), at(3, 2, 1, 0)])
    ref = np.array([1.473846437649123, 0.06134886004475955, -1.4849224560837744, -0.2316303634513769, -0.5309613807759392])
    assert np.linalg.norm(got - ref) < 1e-14


def test_rng_plaquette_kat():
    # simple.py:18-27 : same rng object walks through three grids
    rng = random("block_seed_string_13")
    for dims, prec, ref, scale, precision in [
        ([8, 4, 4, 4], 1e-28, -0.00014108397456619623, 10, "double"),
        ([8, 4, 4, 4], 1e-14, -0.00014108397456619623, 10, "single"),
        ([8, 8, 4, 8], 1e-28, 0.38723058417632267, 2, "double"),
    ]:
        U = qcd.gauge_random(rng, dims, scale=scale, precision=precision)
        assert abs(qcd.plaquette(U) - ref) ** 2.0 < prec
        for u in U:
            eye = np.zeros_like(u)
            eye[..., range(3), range(3)] = 1
            assert np.sum(np.abs(qcd.adj(u) @ u - eye) ** 2) / np.sum(np.abs(u) ** 2) < prec


def test_rng_scalar_choice_kat():
    # simple.py:33-66,132-144: 10000 zn(2), 10000 zn(3), 10000 normal, one lattice normal x1001, then choice
    rng = random("block_seed_string_13")
    for _ in range(10000):
        rng.scalar_zn(2)
    for _ in range(10000):
        rng.scalar_zn(3)
    for _ in range(10000):
        rng.scalar_normal()
    assert rng.choice(["A", "B", "C"], 10) == ["C", "C", "A", "C", "C", "B", "A", "B", "C", "A"]
    assert rng.choice([1, 2, 3], 5) == [2, 3, 3, 3, 1]


@pytest.fixture(scope="module")
def fingerprint_fields():
    # tests/qcd/fermion_operators.py:849-884: U, then for grid in [F_grid, U_grid]: src, dst
    dims = [8, 8, 8, 16]
    rng = random("finger_print")
    U = qcd.gauge_random(rng, dims)
    d5 = [12] + dims
    src5, dst5 = rng.cnormal(d5, (4, 3)), rng.cnormal(d5, (4, 3))
    src4, dst4 = rng.cnormal(dims, (4, 3)), rng.cnormal(dims, (4, 3))
    # Wilson tests: F_grid == U_grid, so the 4d fields are the FIRST draws after U
    rng_w = random("finger_print")
    Uw = qcd.gauge_random(rng_w, dims)
    srcw, dstw = rng_w.cnormal(dims, (4, 3)), rng_w.cnormal(dims, (4, 3))
    return dict(U=U, src5=src5, dst5=dst5, src4=src4, dst4=dst4, Uw=Uw, srcw=srcw, dstw=dstw)


def _close(x, ref):
    return abs(x - ref) / abs(ref) < 1e-13  # finger_print_tolerance * eps (fermion_operators.py:527,892-908)


WILSON = dict(kappa=0.13500, csw_r=0.0, csw_t=0.0, xi_0=1.33111, nu=2.61, isAnisotropic=True, boundary_phases=[1.0, -1.0, 1.0, -1.0])
CLOVER = dict(WILSON, csw_r=1.5, csw_t=1.951)
MOBIUS = dict(mass=0.08, M5=1.8, b=1.5, c=0.5, Ls=12, boundary_phases=[1.0, -1.0, 1.0, -1.0])
MOBIUS_AXIAL = dict(mass_plus=0.08, mass_minus=0.11, M5=1.8, b=1.5, c=0.5, Ls=12, boundary_phases=[1.0, -1.0, 1.0, -1.0])


def test_wilson_fingerprints(fingerprint_fields):
    f = fingerprint_fields
    w = qcd.wilson_clover(f["Uw"], **WILSON)
    assert _close(qcd.inner_product(f["dstw"], w.M(f["srcw"])), -999.7564252326631 - 466.7758727463097j)
    assert _close(qcd.inner_product(f["dstw"], w.Mdiag(f["srcw"])), -961.5053827614738 - 3468.430447866095j)
    w = qcd.wilson_clover(f["Uw"], **CLOVER)
    assert _close(qcd.inner_product(f["dstw"], w.M(f["srcw"])), -946.8714968698364 - 427.1253034080037j)
    assert _close(qcd.inner_product(f["dstw"], w.Mdiag(f["srcw"])), -908.620454398646 - 3428.779878527792j)


def test_wilson_twisted_mass_fingerprints(fingerprint_fields):
    # tests/qcd/fermion_operators.py:365-369,387-390
    f = fingerprint_fields
    w = qcd.wilson_clover(f["Uw"], mass=-1.8, mu=0.2, boundary_phases=[1.0, 1.0, 1.0, -1.0])
    assert _close(qcd.inner_product(f["dstw"], w.M(f["srcw"])), -5.665095757463064 + 373.96051873176737j)
    assert _close(qcd.inner_product(f["dstw"], w.Mdiag(f["srcw"])), -440.5312395819657 - 1102.362512575698j)
    # MooeeInv Mooee = 1, Mdag = adjoint of M
    x = w.MooeeInv(w.Mooee(f["srcw"]))
    assert np.linalg.norm(x - f["srcw"]) / np.linalg.norm(f["srcw"]) < 1e-14
    x = w.MooeeInv(w.Mooee(f["srcw"], dag=True), dag=True)
    assert np.linalg.norm(x - f["srcw"]) / np.linalg.norm(f["srcw"]) < 1e-14
    a, b = qcd.inner_product(f["dstw"], w.M(f["srcw"])), qcd.inner_product(w.Mdag(f["dstw"]), f["srcw"])
    assert abs(a - b) / abs(a) < 1e-13


def test_wilson_clover_open_bc_fingerprints(fingerprint_fields):
    # tests/qcd/fermion_operators.py:356-364,383-386: open boundary conditions in time, cF = 1.3
    f = fingerprint_fields
    w = qcd.wilson_clover(f["Uw"], kappa=0.13500, csw_r=1.978, csw_t=1.978, cF=1.3, xi_0=1, nu=1, isAnisotropic=False,
                          boundary_phases=[1.0, 1.0, 1.0, 0.0])
    assert _close(qcd.inner_product(f["dstw"], w.M(f["srcw"])), -1634.2615676797234 + 239.27037187495998j)
    assert _close(qcd.inner_product(f["dstw"], w.Mdiag(f["srcw"])), -1239.3535155227526 - 1158.5295177146759j)


ZMOBIUS = dict(
    mass=0.08, M5=1.8, b=1.0, c=0.0, boundary_phases=[1.0, 1.0, 1.0, -1.0],
    omega=[0.17661651536320583 + 1j * (0.14907774771612217), 0.23027432016909377 + 1j * (-0.03530801572584271),
           0.3368765581549033 + 1j * (0), 0.7305711010541054 + 1j * (0), 1.1686138337986505 + 1j * (0.3506492418109086),
           1.1686138337986505 + 1j * (-0.3506492418109086), 0.994175013717952 + 1j * (0), 0.5029903152251229 + 1j * (0),
           0.23027432016909377 + 1j * (0.03530801572584271), 0.17661651536320583 + 1j * (-0.14907774771612217)])


def test_zmobius_fingerprints():
    # tests/qcd/fermion_operators.py:397-426; draw order of the reference's test: U, then src/dst on F_grid (Ls = 10), U_grid
    dims = [8, 8, 8, 16]
    rng = random("finger_print")
    U = qcd.gauge_random(rng, dims)
    d5 = [10] + dims
    src5, dst5 = rng.cnormal(d5, (4, 3)), rng.cnormal(d5, (4, 3))
    src4 = rng.cnormal(dims, (4, 3))
    m = qcd.zmobius(U, **ZMOBIUS)
    assert _close(qcd.inner_product(dst5, m.M(src5)), -2424.048033434305 + 10557.661684178218j)
    assert _close(qcd.inner_product(dst5, m.Mdiag(src5)), 2643.396577965267 + 6550.259431381319j)
    assert _close(qcd.inner_product(dst5, m.ImportPhysicalFermionSource(src4)), 4064.7879718582053 - 1357.0856808000196j)
    # structure: adjoints and inverses
    a, b = qcd.inner_product(dst5, m.M(src5)), qcd.inner_product(m.Mdag(dst5), src5)
    assert abs(a - b) / abs(a) < 1e-13
    x = m.MooeeInv(m.Mooee(src5))
    assert np.linalg.norm(x - src5) / np.linalg.norm(src5) < 1e-13
    x = m.MooeeInv(m.Mooee(src5, dag=True), dag=True)
    assert np.linalg.norm(x - src5) / np.linalg.norm(src5) < 1e-13
    a, b = qcd.inner_product(dst5, m.Dminus(src5)), qcd.inner_product(m.Dminus(dst5, dag=True), src5)
    assert abs(a - b) / abs(a) < 1e-13


def test_mobius_fingerprints(fingerprint_fields):
    f = fingerprint_fields
    m = qcd.mobius(f["U"], **MOBIUS)
    assert _close(qcd.inner_product(f["dst5"], m.M(f["src5"])), -8693.09425573421 - 4130.7793316734915j)
    assert _close(qcd.inner_product(f["dst5"], m.Mdiag(f["src5"])), -4966.960264746144 - 2525.83968136146j)
    assert _close(qcd.inner_product(f["dst5"], m.ImportPhysicalFermionSource(f["src4"])), -97.93443075273976 - 690.6405168964976j)
    m = qcd.mobius(f["U"], **MOBIUS_AXIAL)
    assert _close(qcd.inner_product(f["dst5"], m.M(f["src5"])), -8690.547330400455 - 4127.148886222195j)
    assert _close(qcd.inner_product(f["dst5"], m.Mdiag(f["src5"])), -4967.102993398692 - 2525.589904941078j)
    assert _close(qcd.inner_product(f["dst5"], m.ImportPhysicalFermionSource(f["src4"])), -97.93443075274081 - 690.6405168964941j)


def test_oracle_structure(fingerprint_fields):
    # adjoint consistency, inverse, eo projection (fermion_operators.py:677-743)
    f = fingerprint_fields
    m = qcd.mobius(f["U"], **MOBIUS_AXIAL)
    a = qcd.inner_product(f["dst5"], m.M(f["src5"]))
    b = qcd.inner_product(m.Mdag(f["dst5"]), f["src5"])
    assert abs(a - b) / abs(a) < 1e-13
    x = m.MooeeInv(m.Mooee(f["src5"]))
    assert np.linalg.norm(x - f["src5"]) / np.linalg.norm(f["src5"]) < 1e-13
    # M restricted to parities equals Meooe + Mooee
    e = qcd.eo_ops(m)
    se = e.proj(f["src5"], 0)
    full = m.M(se)
    assert np.linalg.norm(e.proj(full, 1) - e.Meooe(se, 0)) / np.linalg.norm(full) < 1e-13
    assert np.linalg.norm(e.proj(full, 0) - e.Mooee(se)) / np.linalg.norm(full) < 1e-13


PION_REF = [
    1.0710210800170898, 0.08988216519355774, 0.015699388459324837, 0.003721018321812153, 0.0010877142194658518,
    0.0003579717595130205, 0.00012700144725386053, 5.180457083042711e-05, 3.406393443583511e-05, 5.2738148951902986e-05,
    0.0001297977869398892, 0.0003634534077718854, 0.0011047901352867484, 0.0037904218770563602, 0.015902264043688774,
    0.09077762067317963,
]
PION_PARAMS = dict(kappa=0.137, csw_r=0.0, csw_t=0.0, xi_0=1.0, nu=1.0, isAnisotropic=False,
                   boundary_phases=[np.exp(1j), np.exp(2j), np.exp(3j), np.exp(4j)])


def test_wilson_pion_correlator_golden():
    """end-to-end pin: eo2_ne CG Wilson propagator from a point source at [1,0,0,0] and its pion correlator
    (/root/reference/tests/qcd/fermion_operators.py:12-43,135-218; 16 values, tolerance 1e-5)"""
    dims = [8, 8, 8, 16]
    rng = random("test")
    U = qcd.gauge_random(rng, dims)
    w = qcd.wilson_clover(U, **PION_PARAMS)
    corr = np.zeros(16)
    for col in range(12):
        src = np.zeros(tuple(dims[::-1]) + (4, 3), dtype=np.complex128)
        src[0, 0, 0, 1].reshape(12)[col] = 1.0  # [t,z,y,x] = point (x=1, y=z=t=0)
        sol, hist = qcd.propagator_column(w, src, 1e-6, 1000)
        corr += np.sum(np.abs(sol) ** 2, axis=(1, 2, 3, 4, 5))
    assert np.linalg.norm(corr - np.array(PION_REF)) < 1e-5


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("Ls", [12, 8, 5])
def test_c_port_matches_numpy_oracle(dtype, Ls):
    """oracle/dslash_ref.c (bench.py's CPU baseline) against the numpy restatement of the reference's hopping term, both
    directions of the dagger, anisotropic coefficients"""
    from oracle import cref

    dims = [4, 6, 4, 8]
    rng = random("c port")
    U = qcd.gauge_random(rng, dims, scale=0.9)
    V = qcd.apply_boundary_phases(U, [1.0, -1.0, np.exp(0.3j), -1.0])
    psi = rng.cnormal([Ls] + dims, (4, 3)).astype(dtype)
    Vc = np.stack([v.reshape(-1, 3, 3) for v in V]).astype(dtype)
    coef = (1.3, 1.3, 1.3, 1.0)
    tol = 1e-14 if dtype == np.complex128 else 2e-6
    for dag in (False, True):
        ref = qcd.dhop([v.astype(dtype) for v in V], psi, coef, dag, five_d=True)
        a = cref.dhop(dims, Ls, Vc, psi.reshape(-1, 4, 3), coef, dag).reshape(ref.shape)
        assert np.linalg.norm(a - ref) / np.linalg.norm(ref) < tol


def test_fixture_matches_embedded_numbers():
    """tests/golden/reference_vectors.json (made by tests/golden/extract_reference_vectors.py from the reference's own tests)
    holds exactly the numbers the parity tests assert; where the reference tree is available the extraction is repeated"""
    import importlib.util
    import json
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_vectors.json")) as f:
        gold = json.load(f)
    fp = gold["fingerprints"]

    def c(entry, key):
        re_, im_ = entry["values"][key]
        return complex(re_, im_)

    assert c(fp["wilson_matrices"], "") == -999.7564252326631 - 466.7758727463097j
    assert c(fp["wilson_clover_matrices"], ".Mdiag") == -908.620454398646 - 3428.779878527792j
    assert c(fp["wilson_clover_matrices_open"], "") == -1634.2615676797234 + 239.27037187495998j
    assert c(fp["wilson_twisted_mass_matrices"], ".Mdiag") == -440.5312395819657 - 1102.362512575698j
    assert c(fp["mobius_matrices"], "") == -8693.09425573421 - 4130.7793316734915j
    assert c(fp["mobius_axial_mass_matrices"], ".ImportPhysicalFermionSource") == -97.93443075274081 - 690.6405168964941j
    assert c(fp["zmobius_matrices"], "") == -2424.048033434305 + 10557.661684178218j
    assert c(fp["zmobius_matrices"], ".ImportPhysicalFermionSource") == 4064.7879718582053 - 1357.0856808000196j
    assert fp["wilson_clover_params"]["values"] == {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in CLOVER.items()}
    assert gold["wilson_pion_correlator"]["values"] == list(PION_REF)
    assert gold["domain_wall_correlator_ref"]["values"][0] == 0.5534145832061768
    assert gold["rng_normal_sequences"][0]["values"] == [-0.29101665386129116, -1.4591269443435488, -0.3641310411719848,
                                                         -0.9454532383815435, 0.4996115272362977]
    assert gold["rng_normal_sequences"][1]["values"][0] == 1.473846437649123
    assert [r["plaquette"] for r in gold["rng_gauge_random_plaquettes"]["values"]] == [-0.00014108397456619623, -0.00014108397456619623,
                                                                                        0.38723058417632267]
    assert gold["rng_choice_letters"]["values"] == ["C", "C", "A", "C", "C", "B", "A", "B", "C", "A"]
    assert gold["rng_choice_numbers"]["values"] == [2, 3, 3, 3, 1]
    if os.path.isdir("/root/reference/tests"):
        spec = importlib.util.spec_from_file_location("extract_reference_vectors", os.path.join(here, "golden", "extract_reference_vectors.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert json.loads(json.dumps(mod.extract("/root/reference"), sort_keys=True)) == gold

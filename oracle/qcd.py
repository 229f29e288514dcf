"""
ORACLE (test infrastructure only) -- CPU restatement of the fermion-operator hot path of GPT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product
(gpt_b200/) never does.

The arithmetic of the reference lives in a third-party dependency that is NOT in /root/reference:
lehner/Grid, branch feature/gpt (unpinned; /root/reference/README.md:27,
.github/workflows/github-build-test.yml:9).  This file restates the published algorithm and anchors it on
the reference's own Python statement of the same maths and its golden fingerprints:

  * Wilson-clover operator          lib/gpt/qcd/fermion/reference/wilson_clover.py:92-220
  * covariant shift / boundary      lib/gpt/core/covariant.py:23-61
  * gamma basis, sigma_{mu nu}      lib/gpt/core/gamma.py:28-53
  * field strength (clover leaves)  lib/gpt/qcd/gauge/loops.py:156-174
  * SU(3) generators                lib/gpt/core/object_type/su_n.py:208-240
  * matrix exp                      lib/gpt/core/foundation/lattice/matrix/exp.py:167-219
  * gauge.random                    lib/gpt/qcd/gauge/create.py:66-71, lib/gpt/core/random.py:110-148
  * Moebius operator D_DWF          tests/qcd/domain_wall.py:309-341
  * M5D / MooeeInv structure        lib/cgpt/lib/foundation/mobius_with_vector_field.h:36-96,176-302,500-513
  * opcode table                    lib/cgpt/lib/operators/register.h:2-20

Parity pin: tests/test_oracle_golden.py (this repo) reproduces the golden numbers of
/root/reference/tests/qcd/fermion_operators.py:371-459 and tests/random/simple.py:18-27.

Layout ("oracle layout"): a field on a grid with dims [L0,L1,L2,L3] (x,y,z,t) is a numpy array of shape
[L3,L2,L1,L0] + tensor shape; on a 5d grid [Ls,L0,L1,L2,L3] it is [L3,L2,L1,L0,Ls] + tensor.  The C-order
flat site index is therefore lexicographic with dimension 0 fastest, which is GPT's `lattice[:]` order.
"""

import numpy as np

# ------------------------------------------------------------------------------------------------
# gamma algebra (gamma.py:28-53)
# ------------------------------------------------------------------------------------------------
gamma = {
    0: np.array([[0, 0, 0, 1j], [0, 0, 1j, 0], [0, -1j, 0, 0], [-1j, 0, 0, 0]], dtype=np.complex128),
    1: np.array([[0, 0, 0, -1], [0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0]], dtype=np.complex128),
    2: np.array([[0, 0, 1j, 0], [0, 0, 0, -1j], [-1j, 0, 0, 0], [0, 1j, 0, 0]], dtype=np.complex128),
    3: np.array([[0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0]], dtype=np.complex128),
    5: np.diagflat([1, 1, -1, -1]).astype(np.complex128),
    "I": np.identity(4, dtype=np.complex128),
}


def sigma(mu, nu):
    return 0.5 * (gamma[mu] @ gamma[nu] - gamma[nu] @ gamma[mu])


Pplus = 0.5 * (gamma["I"] + gamma[5])
Pminus = 0.5 * (gamma["I"] - gamma[5])


def axis(mu):
    """numpy axis of direction mu (0..3 = x,y,z,t) in oracle layout (4d and 5d)."""
    return 3 - mu


def shift(f, mu, d=1):
    """cshift(f, mu, d)(x) = f(x + d mu)   (core/foundation/lattice/cshift_plan.py:53-55)"""
    return np.roll(f, -d, axis=axis(mu))


def spin_mul(g, f):
    """apply 4x4 spin matrix to a spinor field [..., 4, 3]"""
    return np.einsum("ab,...bc->...ac", g, f)


def parity_mask(dims4):
    L0, L1, L2, L3 = dims4
    t, z, y, x = np.meshgrid(np.arange(L3), np.arange(L2), np.arange(L1), np.arange(L0), indexing="ij")
    return (x + y + z + t) % 2


# ------------------------------------------------------------------------------------------------
# gauge field
# ------------------------------------------------------------------------------------------------
def su3_generators():
    # su_n.py:208-240
    Nc = 3
    r = []
    for i in range(Nc):
        for j in range(i + 1, Nc):
            a = np.zeros((Nc, Nc), dtype=np.complex128)
            a[i, j] = 1.0
            a[j, i] = 1.0
            r.append(a)
            a = np.zeros((Nc, Nc), dtype=np.complex128)
            a[i, j] = -1j
            a[j, i] = 1j
            r.append(a)
            if j == i + 1:
                a = np.zeros((Nc, Nc), dtype=np.complex128)
                for l in range(j):
                    a[l, l] = 1.0
                a[j, j] = -j
                r.append(a)
    for a in r:
        a /= (np.trace(a @ a) * 2.0) ** 0.5
    return r


def matrix_exp(x):
    # exp.py:167-219, double precision, lattice-global scaling
    x = x.astype(np.complex128)
    gsites = int(np.prod(x.shape[:-2]))
    n = float(np.sum(np.abs(x) ** 2)) ** 0.5 / gsites
    maxn = 0.01
    ns = 0
    if n > maxn:
        ns = int(np.log2(n / maxn))
        x = x / 2**ns
    o = np.zeros_like(x)
    o[..., range(3), range(3)] = 1.0
    xn = x.copy()
    o = o + xn
    nfac = 1.0
    for j in range(2, 20):
        nfac /= j
        xn = xn @ x
        o = o + xn * nfac
    for j in range(ns):
        o = o @ o
    return o


def gauge_random(rng, dims4, scale=1.0, precision="double"):
    """g.qcd.gauge.random(grid, rng, scale) -> list of 4 link fields [T,Z,Y,X,3,3]
    (fields drawn on a double grid use the rng's default generators, i.e. grid_tag None == "double")"""
    precision = None if precision == "double" else precision
    cdt = np.complex128 if precision is None else np.complex64
    gens = [t.astype(cdt) for t in su3_generators()]
    U = []
    for mu in range(4):
        A = np.zeros(tuple(dims4[::-1]) + (3, 3), dtype=cdt)
        for ta in gens:
            ca = rng.uniform_real(dims4, (), min=-0.5, max=0.5, grid_tag=precision).astype(cdt)
            A = A + (cdt(scale) * ca)[..., None, None] * ta
        U.append(matrix_exp(A * 1j).astype(cdt))
    return U


def gauge_unit(dims4, precision="double"):
    cdt = np.complex128 if precision == "double" else np.complex64
    u = np.zeros(tuple(dims4[::-1]) + (3, 3), dtype=cdt)
    u[..., range(3), range(3)] = 1.0
    return [u.copy() for _ in range(4)]


def adj(m):
    return np.conj(np.swapaxes(m, -1, -2))


def plaquette(U):
    # lib/gpt/qcd/gauge/loops.py plaquette: average of Re tr P_{mu nu} / Nc over 6 planes and sites
    tr = 0.0
    vol = int(np.prod(U[0].shape[:-2]))
    for mu in range(4):
        for nu in range(mu):
            p = U[mu] @ shift(U[nu], mu) @ adj(shift(U[mu], nu)) @ adj(U[nu])
            tr += np.sum(np.trace(p, axis1=-2, axis2=-1)).real
    return tr / vol / 6.0 / 3.0


def field_strength(U, mu, nu):
    # loops.py:156-174
    staple_up = shift(U[nu], mu, 1) @ adj(shift(U[mu], nu, 1)) @ adj(U[nu])
    staple_down = shift(adj(shift(U[nu], mu, 1)) @ adj(U[mu]) @ U[nu], nu, -1)
    v = staple_up - staple_down
    F = U[mu] @ v + shift(v @ U[mu], mu, -1)
    return 0.125 * (F - adj(F))


def apply_boundary_phases(U, phases):
    # covariant.py:29-37 : phase multiplies U_mu on the last slice x_mu = L-1
    V = []
    for mu in range(4):
        v = U[mu].copy()
        idx = [slice(None)] * v.ndim
        idx[axis(mu)] = v.shape[axis(mu)] - 1
        v[tuple(idx)] = v[tuple(idx)] * v.dtype.type(phases[mu])
        V.append(v)
    return V


# ------------------------------------------------------------------------------------------------
# checkerboards
# ------------------------------------------------------------------------------------------------
def pick_checkerboard(full, cb, ls=False):
    """full field -> flat [V/2 (*Ls), ...] in lexicographic order of the sites of parity cb."""
    dims4 = full.shape[:4][::-1]
    m = parity_mask(dims4) == cb
    return full[m].reshape((-1,) + full.shape[(5 if ls else 4):])


def set_checkerboard(full, half, cb, ls=False):
    dims4 = full.shape[:4][::-1]
    m = parity_mask(dims4) == cb
    full[m] = half.reshape((-1,) + full.shape[4:])
    return full


# ------------------------------------------------------------------------------------------------
# Wilson hopping term (shared by Wilson-clover and Moebius)
# ------------------------------------------------------------------------------------------------
def link_mul(Umu, f, five_d):
    if five_d:
        return np.einsum("...ab,...ksb->...ksa", Umu, f)
    return np.einsum("...ab,...sb->...sa", Umu, f)


def dhop(V, psi, coef=(1.0, 1.0, 1.0, 1.0), dag=False, five_d=False):
    """
    Grid's Dhop: -1/2 sum_mu c_mu [ (1-g_mu) V_mu(x) psi(x+mu) + (1+g_mu) V_mu^dag(x-mu) psi(x-mu) ];
    dag: g_mu -> -g_mu.   (reference/wilson_clover.py:182-200)
    """
    out = np.zeros_like(psi)
    sgn = -1.0 if dag else 1.0
    for mu in range(4):
        fwd = link_mul(V[mu], shift(psi, mu, +1), five_d)
        bwd = shift(link_mul(adj(V[mu]), psi, five_d), mu, -1)
        gm = sgn * gamma[mu]
        out += coef[mu] / 2.0 * (spin_mul(gm - gamma["I"], fwd) - spin_mul(gm + gamma["I"], bwd))
    return out


# ------------------------------------------------------------------------------------------------
# Wilson-clover
# ------------------------------------------------------------------------------------------------
class wilson_clover:
    """
    g.qcd.fermion.wilson_clover(U, ...)  (lib/gpt/qcd/fermion/wilson.py:115-148, cgpt operators/wilson_clover.h)
    periodic / phase boundary conditions, and open boundary conditions in time (boundary_phases[3] == 0) with the boundary
    improvement coefficient cF exactly as lib/gpt/qcd/fermion/reference/wilson_clover.py:78-125,179-220 does them: no hopping
    across the boundary, clover term replaced by -csw_t/2 on the time slices 0 and T-1, (cF - 1) added on the slices 1 and T-2,
    every operator result set to zero on the slices 0 and T-1 (lib/gpt/qcd/fermion/boundary_conditions.py:23-31).
    Grid's operator (the one the reference fingerprints, tests/qcd/fermion_operators.py:356-364,383-386) is P D P with P the
    projector on 0 < t < T-1: the hopping term also ignores what sits on the two boundary slices of its INPUT; this, not the
    Python reference's one-sided mask, reproduces the golden number.
    """

    def __init__(self, U, kappa=None, mass=None, csw_r=0.0, csw_t=0.0, xi_0=1.0, nu=1.0,
                 isAnisotropic=False, boundary_phases=(1, 1, 1, 1), cF=1.0, mu=0.0):
        if kappa is not None:
            assert mass is None
            mass = 1.0 / kappa / 2.0 - 4.0
        self.open_bc = boundary_phases[3] == 0.0
        if self.open_bc:
            assert xi_0 == 1.0 and nu == 1.0 and csw_r == csw_t  # reference/wilson_clover.py:80-88
        self.U = U
        self.dtype = U[0].dtype
        self.dims4 = U[0].shape[:4][::-1]
        self.V = apply_boundary_phases(U, boundary_phases)
        self.mass = mass
        self.coef = (nu / xi_0, nu / xi_0, nu / xi_0, 1.0)
        self.diag = mass + 1.0 + 3.0 * nu / xi_0
        # twisted mass (g.qcd.fermion.wilson_twisted_mass, lib/gpt/qcd/fermion/wilson.py:99-107; Grid's WilsonTMFermion):
        # Mooee = (4 + m) + i mu gamma_5, gamma_5 = diag(1, 1, -1, -1) in GPT's basis (lib/gpt/core/gamma.py:28-41)
        self.mu = mu
        assert mu == 0.0 or (csw_r == 0.0 and csw_t == 0.0)
        self.clover = None
        if csw_r != 0.0 or csw_t != 0.0:
            cl = np.zeros(U[0].shape[:4] + (4, 4, 3, 3), dtype=np.complex128)
            Ud = [u.astype(np.complex128) for u in U]
            for mu in range(4):
                for nu_ in range(mu + 1, 4):
                    cp = csw_t if nu_ == 3 else csw_r / xi_0
                    F = field_strength(Ud, mu, nu_)
                    cl += -0.5 * cp * np.einsum("ab,...ij->...abij", sigma(mu, nu_), F)
            if self.open_bc:
                T = cl.shape[0]
                for t in (0, T - 1):
                    cl[t] = 0.0
                    for a in range(4):
                        for i in range(3):
                            cl[t, ..., a, a, i, i] = -0.5 * csw_t
                if cF != 1.0:
                    for t in (1, T - 2):
                        for a in range(4):
                            for i in range(3):
                                cl[t, ..., a, a, i, i] += cF - 1.0
            for a in range(4):
                for i in range(3):
                    cl[..., a, a, i, i] += self.diag
            # as 12x12 matrices (spin*3+color)
            self.clover = cl.transpose(0, 1, 2, 3, 4, 6, 5, 7).reshape(U[0].shape[:4] + (12, 12))
            self.clover_inv = np.linalg.inv(self.clover)
            self.clover = self.clover.astype(self.dtype)
            self.clover_inv = self.clover_inv.astype(self.dtype)

    def _bc(self, psi):
        if self.open_bc:
            psi = psi.copy()
            psi[0] = 0.0
            psi[-1] = 0.0
        return psi

    def Dhop(self, psi, dag=False):
        return self._bc(dhop(self.V, self._bc(psi), self.coef, dag))

    def _site(self, mat, psi):
        sh = psi.shape
        return np.einsum("...ab,...b->...a", mat, psi.reshape(sh[:4] + (12,))).reshape(sh)

    def _twist(self, psi, a, b):
        g5 = np.array([1.0, 1.0, -1.0, -1.0]).reshape((1,) * (psi.ndim - 2) + (4, 1))
        return (psi.dtype.type(a) * psi + psi.dtype.type(1j * b) * (g5 * psi)).astype(psi.dtype)

    def Mooee(self, psi, dag=False):  # on a full-lattice field (== Mdiag)
        if self.mu != 0.0:
            return self._twist(psi, self.diag, -self.mu if dag else self.mu)
        if self.clover is None:
            return self._bc(psi.dtype.type(self.diag) * psi)
        return self._bc(self._site(adj(self.clover) if dag else self.clover, psi))

    def MooeeInv(self, psi, dag=False):
        if self.mu != 0.0:
            den = self.diag**2 + self.mu**2
            return self._twist(psi, self.diag / den, (self.mu if dag else -self.mu) / den)
        if self.clover is None:
            return self._bc(psi.dtype.type(1.0 / self.diag) * psi)
        return self._bc(self._site(adj(self.clover_inv) if dag else self.clover_inv, psi))

    Mdiag = Mooee

    def M(self, psi):
        return self.Dhop(psi) + self.Mooee(psi)

    def Mdag(self, psi):
        return self.Dhop(psi, dag=True) + self.Mooee(psi, dag=True)

    # identity maps (reference/wilson_clover.py:165-176)
    def ImportPhysicalFermionSource(self, psi):
        return psi.copy()

    def ExportPhysicalFermionSolution(self, psi):
        return psi.copy()

    def Dminus(self, psi):
        return psi.copy()

    five_d = False


# ------------------------------------------------------------------------------------------------
# Moebius domain wall
# ------------------------------------------------------------------------------------------------
class mobius:
    """
    g.qcd.fermion.mobius(U, mass|mass_plus/mass_minus, M5, b, c, Ls, boundary_phases)
    (lib/gpt/qcd/fermion/mobius.py:26-50,314-327; cgpt operators/mobius.h:34-87).

    M = D_W (b + c S5) + (1 - S5),   D_W = (4 - M5) + Dhop,
    (S5 psi)_s = P+ psi_{s-1} + P- psi_{s+1} - m+ P+ psi_{Ls-1} d_{s,0} - m- P- psi_0 d_{s,Ls-1}
    which is tests/qcd/domain_wall.py:309-341 (D_DWF) written as operators.
    """

    five_d = True

    def __init__(self, U, mass=None, mass_plus=None, mass_minus=None, M5=None, b=None, c=None, Ls=None,
                 boundary_phases=(1, 1, 1, 1)):
        if mass is not None:
            mass_plus = mass
            mass_minus = mass
        self.U = U
        self.dtype = U[0].dtype
        self.dims4 = U[0].shape[:4][::-1]
        self.V = apply_boundary_phases(U, boundary_phases)
        self.mp, self.mm, self.M5, self.b, self.c, self.Ls = mass_plus, mass_minus, M5, b, c, Ls
        self.bee = b * (4.0 - M5) + 1.0
        self.cee = 1.0 - c * (4.0 - M5)
        # dense Ls x Ls chirality blocks of S5 (acting on the s index)
        Ls = self.Ls
        Sp = np.zeros((Ls, Ls))  # multiplies P+ components: (S5 psi)_s += Sp[s,s'] P+ psi_s'
        Sm = np.zeros((Ls, Ls))
        for s in range(Ls):
            if s >= 1:
                Sp[s, s - 1] = 1.0
            if s + 1 < Ls:
                Sm[s, s + 1] = 1.0
        Sp[0, Ls - 1] = -mass_plus
        Sm[Ls - 1, 0] = -mass_minus
        self.Sp, self.Sm = Sp, Sm

    # ---- s-direction building blocks ---------------------------------------------------------
    def _sop(self, psi, Ap, Am):
        """(out)_s = sum_s' Ap[s,s'] P+ psi_s' + Am[s,s'] P- psi_s'  ; psi [T,Z,Y,X,Ls,4,3]"""
        out = np.empty_like(psi)
        Ap = Ap.astype(psi.dtype if np.iscomplexobj(Ap) else psi.real.dtype)
        Am = Am.astype(psi.dtype if np.iscomplexobj(Am) else psi.real.dtype)
        out[..., 0:2, :] = np.einsum("st,...tac->...sac", Ap, psi[..., 0:2, :])
        out[..., 2:4, :] = np.einsum("st,...tac->...sac", Am, psi[..., 2:4, :])
        return out

    def _AB(self, kind, dag):
        I = np.identity(self.Ls)
        if kind == "A":  # b + c S5
            Ap, Am = self.b * I + self.c * self.Sp, self.b * I + self.c * self.Sm
        elif kind == "B":  # 1 - S5
            Ap, Am = I - self.Sp, I - self.Sm
        elif kind == "ee":  # Mooee = bee - cee S5
            Ap, Am = self.bee * I - self.cee * self.Sp, self.bee * I - self.cee * self.Sm
        elif kind == "eeinv":
            Ap = np.linalg.inv(self.bee * I - self.cee * self.Sp)
            Am = np.linalg.inv(self.bee * I - self.cee * self.Sm)
        if dag:
            Ap, Am = np.conj(Ap.T).copy(), np.conj(Am.T).copy()
        return Ap, Am

    def S(self, kind, psi, dag=False):
        return self._sop(psi, *self._AB(kind, dag))

    # ---- opcodes ---------------------------------------------------------------------------------
    def Dhop(self, psi, dag=False):  # 3001 / 4001 ; on eo-projected input this is DhopEO/OE
        return dhop(self.V, psi, dag=dag, five_d=True)

    def DW(self, psi, dag=False):
        return psi.dtype.type(4.0 - self.M5) * psi + self.Dhop(psi, dag)

    def Meooe(self, psi, dag=False):  # 2003 / 2004 (full-lattice form; restrict with parity masks)
        if not dag:
            return self.Dhop(self.S("A", psi))
        return self.S("A", self.Dhop(psi, dag=True), dag=True)

    def Mooee(self, psi, dag=False):  # 2005 / 2006 / 2009
        return self.S("ee", psi, dag)

    def MooeeInv(self, psi, dag=False):  # 2007 / 2008
        return self.S("eeinv", psi, dag)

    Mdiag = Mooee

    def M(self, psi):  # 2001
        return self.DW(self.S("A", psi)) + self.S("B", psi)

    def Mdag(self, psi):  # 2002
        return self.S("A", self.DW(psi, dag=True), dag=True) + self.S("B", psi, dag=True)

    def Dminus(self, psi, dag=False):  # 2010 / 2011
        return psi - psi.dtype.type(self.c) * self.DW(psi, dag)

    def ImportUnphysicalFermion(self, src4):  # 2013
        out = np.zeros(src4.shape[:4] + (self.Ls,) + src4.shape[4:], dtype=src4.dtype)
        out[..., 0, :, :] = spin_mul(Pplus, src4)
        out[..., self.Ls - 1, :, :] = spin_mul(Pminus, src4)
        return out

    def ImportPhysicalFermionSource(self, src4):  # 2012
        return self.Dminus(self.ImportUnphysicalFermion(src4))

    def ExportPhysicalFermionSolution(self, psi):  # 2014
        return spin_mul(Pminus, psi[..., 0, :, :]) + spin_mul(Pplus, psi[..., self.Ls - 1, :, :])

    def ExportPhysicalFermionSource(self, psi):  # 2015
        return spin_mul(Pplus, psi[..., 0, :, :]) + spin_mul(Pminus, psi[..., self.Ls - 1, :, :])


class zmobius(mobius):
    """
    g.qcd.fermion.zmobius(U, mass, M5, b, c, omega, boundary_phases)  (lib/gpt/qcd/fermion/zmobius.py:24-74; cgpt
    operators/zmobius.h:20-56 -> Grid's ZMobiusFermion): the Moebius operator with complex, s-dependent coefficients
        b_s = 1/2 ((b + c) / omega_s + (b - c)),   c_s = 1/2 ((b + c) / omega_s - (b - c))      (zmobius.py:33),
    M = D_W (diag(b_s) + diag(c_s) S5) + (1 - S5); Mooee = (4 - M5)(diag(b_s) + diag(c_s) S5) + (1 - S5), i.e. Grid's
    bee_s = b_s (4 - M5) + 1, cee_s = 1 - c_s (4 - M5); Dminus psi_s = psi_s - c_s D_W psi_s.  Daggers are the conjugate
    transposes.  Pinned by the reference's fingerprints tests/qcd/fermion_operators.py:397-426.
    """

    def __init__(self, U, omega, mass=None, mass_plus=None, mass_minus=None, M5=None, b=None, c=None, boundary_phases=(1, 1, 1, 1)):
        super().__init__(U, mass=mass, mass_plus=mass_plus, mass_minus=mass_minus, M5=M5, b=b, c=c, Ls=len(omega),
                         boundary_phases=boundary_phases)
        om = np.asarray(omega, dtype=np.complex128)
        self.bs = 0.5 * ((b + c) / om + (b - c))
        self.cs = 0.5 * ((b + c) / om - (b - c))

    def _AB(self, kind, dag):
        I = np.identity(self.Ls, dtype=np.complex128)
        Db, Dc = np.diag(self.bs), np.diag(self.cs)
        A = [Db + Dc @ S for S in (self.Sp, self.Sm)]
        B = [I - S for S in (self.Sp, self.Sm)]
        if kind == "A":
            Ap, Am = A
        elif kind == "B":
            Ap, Am = B
        else:
            ee = [(4.0 - self.M5) * a + bb for a, bb in zip(A, B)]
            Ap, Am = ee if kind == "ee" else [np.linalg.inv(e) for e in ee]
        if dag:
            Ap, Am = np.conj(Ap.T).copy(), np.conj(Am.T).copy()
        return Ap, Am

    def Dminus(self, psi, dag=False):
        cs = (np.conj(self.cs) if dag else self.cs).astype(psi.dtype).reshape((self.Ls, 1, 1))
        return psi - cs * self.DW(psi, dag)


# ------------------------------------------------------------------------------------------------
# even-odd Schur complement, normal equation, CG -- on full-lattice arrays with parity masks
# ------------------------------------------------------------------------------------------------
class eo_ops:
    """Half-lattice operators expressed on full arrays that are zero on the other parity."""

    def __init__(self, op):
        self.op = op
        m = parity_mask(op.dims4)
        extra = (1,) * (3 if op.five_d else 2)
        self.mask = [(m == cb).reshape(m.shape + extra) for cb in (0, 1)]

    def proj(self, f, cb):
        return f * self.mask[cb].astype(f.real.dtype)

    def Meooe(self, f, cb_in, dag=False):
        if self.op.five_d:
            r = self.op.Meooe(f, dag)
        else:
            r = self.op.Dhop(f, dag)
        return self.proj(r, 1 - cb_in)

    def Mooee(self, f, dag=False):
        return self.op.Mooee(f, dag)

    def MooeeInv(self, f, dag=False):
        return self.op.MooeeInv(f, dag)


def inner_product(a, b):
    # conj on the left argument, accumulated in double (foundation/reduce.h:129)
    return complex(np.vdot(a.astype(np.complex128), b.astype(np.complex128)))


def norm2(a):
    return inner_product(a, a).real


class schur_complement_two:
    """lib/gpt/algorithms/preconditioner/schur_complement_two.py:56-160 with D = parity (default odd)."""

    def __init__(self, op, parity=1):
        self.eo = eo_ops(op)
        self.D, self.C = parity, 1 - parity

    def Mpc(self, i_d):
        e = self.eo
        t = e.MooeeInv(i_d)
        t = e.Meooe(t, self.D)
        t = e.MooeeInv(t)
        t = e.Meooe(t, self.C)
        return i_d - t

    def MpcDag(self, i_d):
        e = self.eo
        t = e.Meooe(i_d, self.D, dag=True)
        t = e.MooeeInv(t, dag=True)
        t = e.Meooe(t, self.C, dag=True)
        t = e.MooeeInv(t, dag=True)
        return i_d - t

    def R(self, i):
        e = self.eo
        return e.proj(i, self.D) - e.Meooe(e.MooeeInv(e.proj(i, self.C)), self.C)

    def L(self, i_d):
        e = self.eo
        t = e.MooeeInv(i_d)
        return t - e.MooeeInv(e.Meooe(t, self.D))

    def S(self, i):
        e = self.eo
        return e.MooeeInv(e.proj(i, self.C))


def cg(mat, src, eps, maxiter, psi=None):
    """lib/gpt/algorithms/inverter/cg.py:47-112 (no preconditioner); returns (psi, history)"""
    rdt = src.real.dtype.type
    if psi is None:
        psi = np.zeros_like(src)
    mmp = mat(psi)
    r = src - mmp
    p = r.copy()
    cp = norm2(p)
    ssq = norm2(src)
    history = []
    if ssq == 0.0:
        return np.zeros_like(src), history
    rsq = eps**2.0 * ssq
    for k in range(maxiter):
        c = cp
        mmp = mat(p)
        d = inner_product(p, mmp).real
        a = c / d
        r = src.dtype.type(-a) * mmp + r
        cp = norm2(r)
        b = cp / c
        psi = psi + src.dtype.type(a) * p
        p = src.dtype.type(b) * p + r
        history.append(abs(cp))
        if abs(cp) <= rsq:
            break
    return psi, history


def solve_eo2_ne(op, src, eps, maxiter, parity=1):
    """
    inv.preconditioned(pc.eo2_ne(), inv.cg(eps, maxiter))(op) applied to a full-lattice source with zero
    initial guess (algorithms/inverter/preconditioned.py:32-54, preconditioner/normal_equation.py:30-63).
    """
    sc = schur_complement_two(op, parity)
    pc_src = sc.MpcDag(sc.R(src))
    pc_dst, history = cg(lambda x: sc.MpcDag(sc.Mpc(x)), pc_src, eps, maxiter)
    return sc.L(pc_dst) + sc.S(src), history


def propagator_column(op, src4, eps, maxiter):
    """fermion.propagator(slv)(src) for one spin-colour column (operator/base.py:270-286)"""
    s5 = op.ImportPhysicalFermionSource(src4)
    sol, history = solve_eo2_ne(op, s5, eps, maxiter)
    return op.ExportPhysicalFermionSolution(sol), history

"""g.gamma (lib/gpt/core/gamma.py:28-80): the gamma matrices of GPT's basis as operators on spin-colour vector fields.

  g.gamma[mu] (mu = 0..3 or "X","Y","Z","T"), g.gamma[5], g.gamma["I"], g.gamma[mu, nu] = sigma_{mu nu} = 1/2 [gamma_mu, gamma_nu]

Linear combinations and products stay spin matrices (0.5 * (g.gamma["I"] + g.gamma[5]) is the chiral projector P+), and
`G * field` enters expressions like any matrix operator; the arithmetic is one kernel (cgptb_lattice_spin_matrix).
"""
import numpy as np

from gpt_b200 import capi
from gpt_b200.core import matrix_operator

_BASIS = {
    0: np.array([[0, 0, 0, 1j], [0, 0, 1j, 0], [0, -1j, 0, 0], [-1j, 0, 0, 0]], dtype=np.complex128),
    1: np.array([[0, 0, 0, -1], [0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0]], dtype=np.complex128),
    2: np.array([[0, 0, 1j, 0], [0, 0, 0, -1j], [-1j, 0, 0, 0], [0, 1j, 0, 0]], dtype=np.complex128),
    3: np.array([[0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0]], dtype=np.complex128),
    5: np.diagflat([1, 1, -1, -1]).astype(np.complex128),
    "I": np.identity(4, dtype=np.complex128),
}
_NAMES = {"X": 0, "Y": 1, "Z": 2, "T": 3}


class spin_matrix(matrix_operator):
    def __init__(self, m):
        self.matrix = np.asarray(m, dtype=np.complex128).reshape(4, 4)
        m0 = self.matrix

        def apply(mat):
            return lambda dst, src: capi.lattice_spin_matrix(dst.obj, src.obj, mat)

        try:
            minv = np.linalg.inv(m0)
        except np.linalg.LinAlgError:
            minv = None
        super().__init__(
            mat=apply(m0), adj_mat=apply(np.conj(m0.T)),
            inv_mat=apply(minv) if minv is not None else None,
            adj_inv_mat=apply(np.conj(minv.T)) if minv is not None else None,
        )

    def adj(self):
        return spin_matrix(np.conj(self.matrix.T))

    def inv(self):
        return spin_matrix(np.linalg.inv(self.matrix))

    def __add__(self, other):
        return spin_matrix(self.matrix + other.matrix) if isinstance(other, spin_matrix) else NotImplemented

    def __sub__(self, other):
        return spin_matrix(self.matrix - other.matrix) if isinstance(other, spin_matrix) else NotImplemented

    def __neg__(self):
        return spin_matrix(-self.matrix)

    def __mul__(self, other):
        if isinstance(other, spin_matrix):
            return spin_matrix(self.matrix @ other.matrix)
        if isinstance(other, (int, float, complex)):
            return spin_matrix(self.matrix * other)
        return super().__mul__(other)

    def __rmul__(self, other):
        if isinstance(other, (int, float, complex)):
            return spin_matrix(self.matrix * other)
        return super().__rmul__(other)


class _gamma_table:
    def __getitem__(self, key):
        if isinstance(key, tuple):
            mu, nu = (_NAMES.get(k, k) for k in key)
            return spin_matrix(0.5 * (_BASIS[mu] @ _BASIS[nu] - _BASIS[nu] @ _BASIS[mu]))
        return spin_matrix(_BASIS[_NAMES.get(key, key)])


gamma = _gamma_table()

"""
inv.defect_correcting(inner_inverter, eps, maxiter): iterative refinement.

    x_{i+1} = x_i + inner(b - A x_i)

with `inner` an approximate inverse of A -- in practice inv.mixed_precision(...) of an even-odd CG, so that the defect is
formed in double precision while the iterations run in single (the stack of /root/reference/tests/manual/mpi.py:104-110).
Observable behaviour follows lib/gpt/algorithms/inverter/defect_correcting.py:73-139: history[i] = |b - A x_i| / |b| measured
BEFORE the i-th correction, convergence when that drops below eps, the defect handed to the inner solver normalised by |b|
(single precision would underflow on small defects otherwise), and |A x_0| (or 1) standing in for |b| when b = 0.
"""
import gpt_b200 as g
from gpt_b200.algorithms.base import base_iterative


class defect_correcting(base_iterative):
    @g.params_convention(eps=1e-15, maxiter=1000000)
    def __init__(self, inner_inverter, params):
        super().__init__()
        self.params, self.inner_inverter = params, inner_inverter
        self.eps, self.maxiter = params["eps"], params["maxiter"]

    @staticmethod
    def _scale_of(src, outer_mat, psi):
        for candidate in (lambda: g.norm2(src), lambda: g.norm2(g(outer_mat * psi))):
            n2 = candidate()
            if n2 != 0.0:
                return n2**0.5
        return 1.0

    def __call__(self, outer_mat):
        approximate_inverse = self.inner_inverter(outer_mat)

        @self.timed_function
        def solve(psi, src, t):
            scale = self._scale_of(src, outer_mat, psi)
            defect = g.lattice(src)
            for i in range(self.maxiter):
                defect @= src - outer_mat * psi
                relative = g.norm2(defect) ** 0.5 / scale
                self.log_convergence(i, relative, self.eps)
                if relative < self.eps:
                    self.log(f"converged after {i} iterations")
                    return
                defect /= scale
                psi += approximate_inverse(defect) * scale

        vector_space = outer_mat.vector_space if isinstance(outer_mat, g.matrix_operator) else None
        return g.matrix_operator(mat=solve, inv_mat=outer_mat, vector_space=vector_space, accept_guess=(True, False))

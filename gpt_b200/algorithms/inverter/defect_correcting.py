"""inv.defect_correcting (lib/gpt/algorithms/inverter/defect_correcting.py:73-139)"""
import gpt_b200 as g
from gpt_b200.algorithms.base import base_iterative


class defect_correcting(base_iterative):
    @g.params_convention(eps=1e-15, maxiter=1000000)
    def __init__(self, inner_inverter, params):
        super().__init__()
        self.params = params
        self.eps = params["eps"]
        self.maxiter = params["maxiter"]
        self.inner_inverter = inner_inverter

    def __call__(self, outer_mat):
        inner_inv_mat = self.inner_inverter(outer_mat)

        @self.timed_function
        def inv(psi, src, t):
            _s = g.lattice(src)
            norm2_of_source = g.norm2(src)
            if norm2_of_source == 0.0:
                norm2_of_source = g.norm2(g(outer_mat * psi))
                if norm2_of_source == 0.0:
                    norm2_of_source = 1.0
            for i in range(self.maxiter):
                _s @= src - outer_mat * psi  # remaining src
                norm2_of_defect = g.norm2(_s)
                eps = (norm2_of_defect / norm2_of_source) ** 0.5
                self.log_convergence(i, eps, self.eps)
                if eps < self.eps:
                    self.log(f"converged after {i} iterations")
                    break
                # normalize _s to avoid floating-point underflow in inner_inv_mat
                _s /= norm2_of_source**0.5
                _d = inner_inv_mat(_s)
                psi += _d * norm2_of_source**0.5

        vector_space = outer_mat.vector_space if isinstance(outer_mat, g.matrix_operator) else None
        return g.matrix_operator(mat=inv, inv_mat=outer_mat, vector_space=vector_space, accept_guess=(True, False))

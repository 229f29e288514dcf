"""inv.preconditioned: M^-1 = L Mpc^-1 R + S (lib/gpt/algorithms/inverter/preconditioned.py:24-54)"""
import gpt_b200 as g
from gpt_b200.algorithms.base import base


class preconditioned(base):
    @g.params_convention()
    def __init__(self, preconditioner, inverter, params):
        super().__init__()
        self.params = params
        self.preconditioner = preconditioner
        self.inverter = inverter

    def __call__(self, mat):
        matrix = self.preconditioner(mat)
        inv_mat = self.inverter(matrix.Mpc)

        @self.timed_function
        def inv(dst, src, t):
            pc_src = g(matrix.R * src)
            pc_dst = g(matrix.L.inv() * dst)
            inv_mat(pc_dst, pc_src)
            g.eval(dst, matrix.L * pc_dst + matrix.S * src)

        return g.matrix_operator(
            mat=inv, inv_mat=mat, adj_inv_mat=mat.adj(), adj_mat=None,
            vector_space=mat.vector_space, accept_guess=(True, False),
        )

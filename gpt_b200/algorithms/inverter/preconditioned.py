"""
inv.preconditioned(preconditioner, inverter): solve M x = b through a preconditioned system.

A preconditioner (e.g. g.qcd.fermion.preconditioner.eo2_ne) turns the matrix M into an object with four operators,
    M^-1 = L Mpc^-1 R + S
(reference: lib/gpt/algorithms/inverter/preconditioned.py:24-54, lib/gpt/algorithms/preconditioner/schur_complement_two.py):
R maps the source to the preconditioned space, the inner inverter solves with Mpc there -- starting from L^-1 applied to
whatever the caller left in dst, which is how an initial guess survives the change of variables --, L maps the solution back
and S adds the part of the solution that needs no solve.
"""
import gpt_b200 as g
from gpt_b200.algorithms.base import base


class preconditioned(base):
    @g.params_convention()
    def __init__(self, preconditioner, inverter, params):
        super().__init__()
        self.params, self.preconditioner, self.inverter = params, preconditioner, inverter

    def __call__(self, mat):
        pc = self.preconditioner(mat)
        solve_pc = self.inverter(pc.Mpc)
        to_pc, from_pc, guess_to_pc, direct = pc.R, pc.L, pc.L.inv(), pc.S

        @self.timed_function
        def solve(dst, src, t):
            rhs = g(to_pc * src)
            x = g(guess_to_pc * dst)
            solve_pc(x, rhs)
            g.eval(dst, from_pc * x + direct * src)

        return g.matrix_operator(mat=solve, adj_mat=None, inv_mat=mat, adj_inv_mat=mat.adj(), vector_space=mat.vector_space,
                                 accept_guess=(True, False))

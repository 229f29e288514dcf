"""
Schur complement two of an even-odd decomposed operator
(lib/gpt/algorithms/preconditioner/schur_complement_two.py:56-160):

      ( DD DC )
  M = ( CD CC ) ,  Mpc = 1 - DC CC^-1 CD DD^-1 ,  M^-1 = L Mpc^-1 R + S
"""
import os

import gpt_b200 as g
from gpt_b200 import capi


class schur_complement_two:
    def __init__(self, op, domain_decomposition):
        dd_op = domain_decomposition(op)
        DD, CC, CD, DC = dd_op.DD, dd_op.CC, dd_op.CD, dd_op.DC
        CC_inv = CC.inv()
        CC_adj_inv = CC_inv.adj()
        DD_inv = DD.inv()
        DC_adj = DC.adj()
        D, C = dd_op.parity, dd_op.parity.inv()

        op_vector_space = op.vector_space[0]
        D_vector_space = DD.vector_space[0]
        C_vector_space = CC.vector_space[0]
        tmp_d = [D_vector_space.lattice() for i in range(2)]
        tmp_c = [C_vector_space.lattice() for i in range(2)]

        fused = hasattr(op, "interface") and not op.daggered and not os.environ.get("GPT_B200_NO_FUSED")

        def _N(o_d, i_d):
            if fused:
                capi.apply_schur_two(op.interface.obj, False, i_d.obj, o_d.obj)
                return
            DD.inv_mat(tmp_d[0], i_d)
            CD.mat(tmp_c[0], tmp_d[0])
            CC.inv_mat(tmp_c[1], tmp_c[0])
            DC.mat(o_d, tmp_c[1])
            g.axpy(o_d, -1.0, o_d, i_d)

        def _N_dag(o_d, i_d):
            if fused:
                capi.apply_schur_two(op.interface.obj, True, i_d.obj, o_d.obj)
                return
            DC.adj_mat(tmp_c[0], i_d)
            CC.adj_inv_mat(tmp_c[1], tmp_c[0])
            CD.adj_mat(tmp_d[0], tmp_c[1])
            DD.adj_inv_mat(o_d, tmp_d[0])
            g.axpy(o_d, -1.0, o_d, i_d)

        def _L(o, i_d):
            tmp = g(DD_inv * g.expr(i_d))
            dd_op.promote(o, tmp)
            tmp = g(-1.0 * (CC_inv * CD * g.expr(tmp)))
            dd_op.promote(o, tmp)

        def _L_pseudo_inverse(o_d, i):
            g.pick_checkerboard(D, o_d, i)
            t = g.copy(o_d)
            g.eval(o_d, DD * g.expr(t))

        self.L = g.matrix_operator(mat=_L, inv_mat=_L_pseudo_inverse, vector_space=(op_vector_space, D_vector_space))

        def _R(o_d, i):
            g.eval(o_d, g.expr(dd_op.project(D, i)) - DC * CC_inv * g.expr(dd_op.project(C, i)))

        def _R_dag(o, i_d):
            dd_op.promote(o, i_d)
            dd_op.promote(o, g(-1.0 * (CC_adj_inv * DC_adj * g.expr(i_d))))

        self.R = g.matrix_operator(mat=_R, adj_mat=_R_dag, vector_space=(D_vector_space, op_vector_space))

        def _S(o, i):
            dd_op.promote(o, g(CC_inv * g.expr(dd_op.project(C, i))))
            dd_op.promote(o, g(0.0 * g.expr(dd_op.project(D, i))))

        self.S = g.matrix_operator(mat=_S, vector_space=(op_vector_space, op_vector_space))

        self.Mpc = g.matrix_operator(mat=_N, adj_mat=_N_dag, vector_space=(D_vector_space, D_vector_space))
        self.op = op
        self.fused = fused

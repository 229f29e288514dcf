"""similarity transformation of a preconditioned system (lib/gpt/algorithms/preconditioner/similarity_transformation.py:22-84):

  M^-1 = L Mpc^-1 R + S  =  (L V^-1) (V Mpc V^-1)^-1 (V R) + S
"""
import gpt_b200 as g


class similarity_transformation:
    def __init__(self, pc, V):
        self.S = pc.S
        V_inv = V.inv()
        V_adj = V.adj()
        V_adj_inv = V_adj.inv()
        Mpc, L, R = pc.Mpc, pc.L, pc.R
        Mpc_adj = Mpc.adj()
        R_adj = R.adj()

        def _Mpc(o_d, i_d):
            g.eval(o_d, V * Mpc * V_inv * g.expr(i_d))

        def _Mpc_dag(o_d, i_d):
            g.eval(o_d, V_adj_inv * Mpc_adj * V_adj * g.expr(i_d))

        def _R(o_d, i):
            g.eval(o_d, V * R * g.expr(i))

        def _R_dag(o, i_d):
            g.eval(o, R_adj * V_adj * g.expr(i_d))

        def _L(o, i_d):
            g.eval(o, L * V_inv * g.expr(i_d))

        L_inv = L.inv()

        def _L_inv(o_d, i):
            g.eval(o_d, V * L_inv * g.expr(i))

        self.R = g.matrix_operator(mat=_R, adj_mat=_R_dag, vector_space=R.vector_space)
        self.L = g.matrix_operator(mat=_L, inv_mat=_L_inv, vector_space=L.vector_space)
        self.Mpc = g.matrix_operator(mat=_Mpc, adj_mat=_Mpc_dag, vector_space=Mpc.vector_space)

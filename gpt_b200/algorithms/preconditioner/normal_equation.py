"""normal equation of a preconditioned system: Mpc_ne = Mpc^dag Mpc, R_ne = Mpc^dag R
(lib/gpt/algorithms/preconditioner/normal_equation.py:30-68)"""
import gpt_b200 as g


class normal_equation:
    def __init__(self, pc):
        self.L = pc.L
        self.S = pc.S
        Mpc, R = pc.Mpc, pc.R
        Mpc_adj = Mpc.adj()
        R_adj = R.adj()

        def _N_dag_N(o_d, i_d):
            g.eval(o_d, Mpc_adj * Mpc * g.expr(i_d))

        def _R(o_d, i):
            g.eval(o_d, Mpc_adj * R * g.expr(i))

        def _R_dag(o, i_d):
            g.eval(o, R_adj * Mpc * g.expr(i_d))

        self.R = g.matrix_operator(mat=_R, adj_mat=_R_dag, vector_space=R.vector_space)
        self.Mpc = g.matrix_operator(mat=_N_dag_N, adj_mat=_N_dag_N, vector_space=Mpc.vector_space)
        if getattr(pc, "fused", False):
            op = pc.op
            self.Mpc.fused_eo2_ne = lambda: op

"""MADWF (arXiv:1111.5059): approximate inverse of an outer domain-wall operator through a cheaper inner one (e.g. zMoebius with
a smaller Ls).  Mirror of lib/gpt/qcd/fermion/preconditioner/mixed_dwf.py:29-104, written with explicit loops over the
right-hand sides; separate / merge are device copies, the chiral rotation P is built from g.gamma projectors."""
import gpt_b200 as g


class mixed_dwf:
    def __init__(self, solver, solver_pv, dwf_inner):
        self.solver = solver
        self.solver_pv = solver_pv
        self.dwf_inner = dwf_inner
        self.dwf_inner_pv = dwf_inner.modified(mass=1.0)

    def __call__(self, dwf_outer):
        dwf_inner, dwf_inner_pv = self.dwf_inner, self.dwf_inner_pv
        dwf_outer_pv = dwf_outer.modified(mass=1.0)
        inv_dwf_outer_pv = self.solver_pv(dwf_outer_pv)
        inv_dwf_inner = self.solver(dwf_inner)
        Pplus = 0.5 * (g.gamma["I"] + g.gamma[5])
        Pminus = 0.5 * (g.gamma["I"] - g.gamma[5])

        def P(src5, offset):
            # (P x)_s = P- x_s + P+ x_{s + offset}   (offset +1: P, offset -1: P^dag = P^-1)
            xs = g.separate(src5)
            Ls = len(xs)
            return g.merge([g(Pminus * xs[s] + Pplus * xs[(s + Ls + offset) % Ls]) for s in range(Ls)])

        def inv(dst_outer, src_outer):
            Ls_inner = dwf_inner.F_grid.fdimensions[0]
            Ls_outer = dwf_outer.F_grid.fdimensions[0]
            for dst, src in zip(dst_outer, src_outer):
                zero4d = g.lattice(dwf_outer.U_grid, src.otype)
                zero4d[:] = 0
                c_s = g.separate(P(inv_dwf_outer_pv(src), -1))
                wall = g.merge([c_s[0]] + [zero4d] * (Ls_inner - 1))
                y0prime = g.separate(P(inv_dwf_inner(g(dwf_inner_pv * P(wall, +1))), -1))[0]
                wall = g.merge([g(-1.0 * y0prime)] + c_s[1:])
                y1 = g.separate(P(inv_dwf_outer_pv(g(dwf_outer * P(wall, +1))), -1))
                g.copy(dst, P(g.merge([y0prime] + y1[1:]), +1))

        return g.matrix_operator(mat=inv, inv_mat=dwf_outer, accept_guess=(True, False), vector_space=dwf_outer.vector_space,
                                 accept_list=True)

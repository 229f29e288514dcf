"""
inv.cg (lib/gpt/algorithms/inverter/cg.py:24-117): same signature, update order, residual test and `history`.
When the matrix is the even-odd normal equation of one of our fermion operators the whole loop runs
device-side through cgptb_cg_eo2_ne (identical arithmetic; see gpt_b200/csrc/solver.cu).
"""
import os

import gpt_b200 as g
from gpt_b200 import cgpt
from gpt_b200.algorithms.base import base_iterative


class cg(base_iterative):
    @g.params_convention(eps=1e-15, maxiter=1000000, eps_abs=None, miniter=0, prec=None, fail_if_not_converged=False)
    def __init__(self, params):
        super().__init__()
        self.params = params
        self.eps = params["eps"]
        self.eps_abs = params["eps_abs"]
        self.maxiter = params["maxiter"]
        self.miniter = params["miniter"]
        self.prec = params["prec"]
        self.fail_if_not_converged = params["fail_if_not_converged"]

    def modified(self, **params):
        return cg({**self.params, **params})

    def __call__(self, mat):
        prec = self.prec(mat) if self.prec is not None else None
        vector_space = None
        fused = None
        if isinstance(mat, g.matrix_operator):
            vector_space = mat.vector_space
            fused = getattr(mat, "fused_eo2_ne", None)
            mat = mat.specialized_singlet_callable()
        if prec is not None or self.eps_abs is not None or self.miniter != 0 or os.environ.get("GPT_B200_NO_FUSED"):
            fused = None

        @self.timed_function
        def inv(psi, src, t):
            assert src != psi
            if fused is not None:
                hist, conv = cgpt.cg_eo2_ne(fused().interface.obj, psi.obj, src.obj, self.eps, self.maxiter)
                self.history.extend(hist)
                if conv:
                    self.log(f"converged in {len(hist)} iterations")
                elif self.fail_if_not_converged:
                    raise ValueError("FATAL error: CG not converged")
                return
            p, mmp, r = g.lattice(src), g.lattice(src), g.lattice(src)
            if prec is not None:
                z = g.lattice(src)
            mat(mmp, psi)  # in, out
            g.axpy(r, -1.0, mmp, src)
            if prec is not None:
                z[:] = 0
                prec(z, r)
                g.copy(p, z)
                cp = g.inner_product(r, z).real
            else:
                g.copy(p, r)
                cp = g.norm2(p)
            ssq = g.norm2(src)
            if ssq == 0.0:
                psi[:] = 0
                return
            rsq = self.eps**2.0 * ssq
            for k in range(self.maxiter):
                c = cp
                mat(mmp, p)
                d = g.inner_product(p, mmp).real
                a = c / d
                if prec is not None:
                    g.axpy(r, -a, mmp, r)
                    z[:] = 0
                    prec(z, r)
                    cp = g.inner_product(r, z).real
                else:
                    cp = g.axpy_norm2(r, -a, mmp, r)
                b = cp / c
                psi += a * p
                if prec is not None:
                    g.axpy(p, b, p, z)
                else:
                    g.axpy(p, b, p, r)
                res = abs(cp)
                self.log_convergence(k, res, rsq)
                if k + 1 >= self.miniter:
                    if self.eps_abs is not None and res <= self.eps_abs**2.0:
                        self.log(f"converged in {k + 1} iterations (absolute criterion)")
                        return
                    if res <= rsq:
                        self.log(f"converged in {k + 1} iterations")
                        return
            self.log(f"NOT converged in {k + 1} iterations;  squared residual {res:e} / {rsq:e}")
            if self.fail_if_not_converged:
                raise ValueError("FATAL error: CG not converged")

        return g.matrix_operator(mat=inv, inv_mat=mat, accept_guess=(True, False), vector_space=vector_space)

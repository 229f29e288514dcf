"""
inv.cg -- conjugate gradient for a hermitian positive matrix, the solver GPT's propagator stacks end in.

Behaviour is the reference's (lib/gpt/algorithms/inverter/cg.py:24-117), which the parity tests pin down to the iteration:
constructor parameters eps / maxiter / eps_abs / miniter / prec / fail_if_not_converged, `history` = |r|^2 after every
iteration, the stopping rule |r|^2 <= eps^2 |b|^2 evaluated after the update of iteration k, a zero right-hand side giving
psi = 0 without an iteration, the solution vector doubling as initial guess, and the reference's log lines.  The order of
the floating-point operations per iteration is the reference's as well (alpha from <p, A p>, r updated together with its
norm, then psi, then p), so that residual histories agree to rounding.

Two execution paths produce that behaviour:
  * the recurrence below, one library call per vector operation (any matrix, optional preconditioner);
  * for the even-odd normal equation of one of this package's fermion operators the whole solve is ONE library call,
    cgptb_cg_eo2_ne (gpt_b200/csrc/solver.cu), which runs the same recurrence with the Schur complement, the fifth
    dimension sweeps and the vector updates fused into a few kernels per iteration.
"""
import os

import gpt_b200 as g
from gpt_b200 import capi
from gpt_b200.algorithms.base import base_iterative


class _recurrence:
    """state of one solve: psi (solution / guess), r (residual), p (search direction), A p, optionally z = prec r"""

    def __init__(self, matrix, preconditioner, psi, src):
        self.A, self.prec, self.psi = matrix, preconditioner, psi
        self.p, self.Ap, self.r = (g.lattice(src) for _ in range(3))
        self.z = g.lattice(src) if preconditioner is not None else None
        # r = b - A psi, p = (prec) r, rho = <r, (prec) r>
        self.A(self.Ap, psi)
        g.axpy(self.r, -1.0, self.Ap, src)
        self.rho = self._direction_from_residual(first=True)

    def _direction_from_residual(self, first=False):
        if self.prec is None:
            if first:
                g.copy(self.p, self.r)
                return g.norm2(self.p)
            return None
        self.z[:] = 0
        self.prec(self.z, self.r)
        if first:
            g.copy(self.p, self.z)
        return g.inner_product(self.r, self.z).real

    def advance(self):
        """one iteration; returns |rho| (the squared residual for the unpreconditioned recurrence)"""
        rho_old = self.rho
        self.A(self.Ap, self.p)
        alpha = rho_old / g.inner_product(self.p, self.Ap).real
        if self.prec is None:
            self.rho = g.axpy_norm2(self.r, -alpha, self.Ap, self.r)
        else:
            g.axpy(self.r, -alpha, self.Ap, self.r)
            self.rho = self._direction_from_residual()
        beta = self.rho / rho_old
        self.psi += alpha * self.p
        g.axpy(self.p, beta, self.p, self.r if self.prec is None else self.z)
        return abs(self.rho)


class cg(base_iterative):
    @g.params_convention(eps=1e-15, maxiter=1000000, eps_abs=None, miniter=0, prec=None, fail_if_not_converged=False)
    def __init__(self, params):
        super().__init__()
        self.params = params
        for name in ("eps", "eps_abs", "maxiter", "miniter", "prec", "fail_if_not_converged"):
            setattr(self, name, params[name])

    def modified(self, **params):
        return cg({**self.params, **params})

    def _finish(self, iterations, residual, target, converged, how=""):
        if converged:
            self.log(f"converged in {iterations} iterations{how}")
            return
        self.log(f"NOT converged in {iterations} iterations;  squared residual {residual:e} / {target:e}")
        if self.fail_if_not_converged:
            raise ValueError("FATAL error: CG not converged")

    def __call__(self, mat):
        preconditioner = self.prec(mat) if self.prec is not None else None
        vector_space, device_loop = None, None
        if isinstance(mat, g.matrix_operator):
            vector_space = mat.vector_space
            device_loop = getattr(mat, "fused_eo2_ne", None)  # set by preconditioner.eo2_ne on Mpc^dag Mpc of our operators
            mat = mat.specialized_singlet_callable()
        plain = preconditioner is None and self.eps_abs is None and self.miniter == 0
        if not plain or os.environ.get("GPT_B200_NO_FUSED"):
            device_loop = None

        @self.timed_function
        def solve(psi, src, t):
            assert src != psi
            if device_loop is not None:
                history, converged = capi.cg_eo2_ne(device_loop().interface.obj, psi.obj, src.obj, self.eps, self.maxiter)
                self.history.extend(history)
                if history:  # no iteration at all: zero right-hand side, psi = 0 (silent, like the reference)
                    self._finish(len(history), history[-1], self.eps**2.0 * g.norm2(src) if not converged else 0.0, converged)
                return
            state = _recurrence(mat, preconditioner, psi, src)
            norm2_src = g.norm2(src)
            if norm2_src == 0.0:
                psi[:] = 0
                return
            target = self.eps**2.0 * norm2_src
            residual, k = None, -1
            for k in range(self.maxiter):
                residual = state.advance()
                self.log_convergence(k, residual, target)
                if k + 1 < self.miniter:
                    continue
                if self.eps_abs is not None and residual <= self.eps_abs**2.0:
                    return self._finish(k + 1, residual, target, True, " (absolute criterion)")
                if residual <= target:
                    return self._finish(k + 1, residual, target, True)
            self._finish(k + 1, residual, target, False)

        return g.matrix_operator(mat=solve, inv_mat=mat, accept_guess=(True, False), vector_space=vector_space)

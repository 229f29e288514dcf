"""even-odd preconditioners (lib/gpt/qcd/fermion/preconditioner/even_odd_sites.py:23-72)"""
import gpt_b200 as g
from gpt_b200.params import params_convention
from gpt_b200.qcd.fermion.mixed_dwf import mixed_dwf  # noqa: F401


@params_convention(parity=None)
def eo2(params):
    parity = params["parity"] if params["parity"] is not None else g.odd

    def instantiate(op):
        return g.algorithms.preconditioner.schur_complement_two(op, lambda op: op.even_odd_sites_decomposed(parity))

    return instantiate


@params_convention(parity=None)
def eo2_ne(params):
    parity = params["parity"] if params["parity"] is not None else g.odd

    def instantiate(op):
        return g.algorithms.preconditioner.normal_equation(
            g.algorithms.preconditioner.schur_complement_two(op, lambda op: op.even_odd_sites_decomposed(parity))
        )

    return instantiate


@params_convention(parity=None)
def eo2_kappa_ne(params):
    """eo2_ne of the kappa-rescaled Schur complement V Mpc V^-1, V = op.kappa() (zMoebius; even_odd_sites.py:75-89)"""
    parity = params["parity"] if params["parity"] is not None else g.odd

    def instantiate(op):
        return g.algorithms.preconditioner.normal_equation(
            g.algorithms.preconditioner.similarity_transformation(
                g.algorithms.preconditioner.schur_complement_two(op, lambda op: op.even_odd_sites_decomposed(parity)),
                op.kappa(),
            )
        )

    return instantiate

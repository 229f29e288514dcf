"""g.qcd.fermion.wilson_clover (lib/gpt/qcd/fermion/wilson.py:115-148)"""
import copy

import gpt_b200 as g
from gpt_b200.qcd.fermion.operator import fine_operator


@g.params_convention(
    kappa=None, mass=None, cF=1, use_legacy=False, boundary_phases=None, isAnisotropic=None,
    csw_r=None, csw_t=None, nu=None, xi_0=None, n_rhs=1,
    link_compression=None,  # extension of this package: 12 = two-row SU(3) link compression in the stencil's link tables
)
def wilson_clover(U, params):
    params = copy.deepcopy(params)
    if params["kappa"] is not None:
        assert params["mass"] is None
        params["mass"] = 1.0 / params["kappa"] / 2.0 - 4.0
        del params["kappa"]
    if params["n_rhs"] > 1:
        # multi-rhs operator: the right-hand sides are the fifth dimension of the fermion grid and share every link
        # (and clover block) load; use .packed() to apply it to a list of 4d fields (wilson.py:134-138)
        params["multi_rhs"] = True
        params["Ls"] = params["n_rhs"]
    else:
        params["multi_rhs"] = False
    if params["boundary_phases"][-1] != 0.0:
        assert params["cF"] == 1.0  # forbid usage of cF without open bc (wilson.py:142-143)
    return fine_operator("wilson_clover", U, params, otype=g.ot_vector_spin_color(4, 3))


@g.params_convention(mass=None, mu=None, boundary_phases=None, link_compression=None)
def wilson_twisted_mass(U, params):
    """g.qcd.fermion.wilson_twisted_mass (lib/gpt/qcd/fermion/wilson.py:99-107): Wilson hopping term with the site-diagonal
    term (4 + mass) + i mu gamma_5; isotropic, no clover term"""
    params = copy.deepcopy(params)
    params.update(csw_r=0.0, csw_t=0.0, cF=1.0, xi_0=1.0, nu=1.0, isAnisotropic=False)
    return fine_operator("wilson_twisted_mass", U, params, otype=g.ot_vector_spin_color(4, 3))

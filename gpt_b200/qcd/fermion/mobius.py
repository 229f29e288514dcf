"""g.qcd.fermion.mobius (lib/gpt/qcd/fermion/mobius.py:26-50,314-327)"""
import copy

import gpt_b200 as g
from gpt_b200.qcd.fermion.operator import fine_operator


class mobius_class_operator(fine_operator):
    def __init__(self, name, U, params, otype=None, daggered=False):
        if params["mass"] is not None:
            params["mass_plus"] = params["mass"]
            params["mass_minus"] = params["mass"]
        fine_operator.__init__(self, name, U, params, otype, daggered)
        self.bulk_propagator_to_propagator = self.ExportPhysicalFermionSolution

    def bulk_propagator(self, solver):
        imp = self.ImportPhysicalFermionSource
        inv_matrix = solver(self)

        def prop(dst_sc, src_sc):
            g.eval(dst_sc, inv_matrix * imp * g.expr(src_sc))

        op = g.matrix_operator(prop, vector_space=imp.vector_space)
        if self.daggered:
            op = op.adj()
        return g.propagator_operator(op)


@g.params_convention(mass=None, mass_plus=None, mass_minus=None, b=None, c=None, M5=None, boundary_phases=None, Ls=None,
                     link_compression=None)  # link_compression: extension of this package (12 = two-row SU(3) links)
def mobius(U, params):
    params = copy.deepcopy(params)
    return mobius_class_operator("mobius", U, params, otype=g.ot_vector_spin_color(4, 3))


class zmobius_class_operator(mobius_class_operator):
    def kappa(self):
        """diagonal rescaling kappa_s = 1 / (2 (b_s (4 - M5) + 1)) of the fifth dimension (lib/gpt/qcd/fermion/zmobius.py:28-59)"""
        import numpy as np

        b, c, M5 = self.params["b"], self.params["c"], self.params["M5"]
        bs = [0.5 * (1.0 / w * (b + c) + (b - c)) for w in self.params["omega"]]
        kappa = np.array([1.0 / (2.0 * (x * (4.0 - M5) + 1.0)) for x in bs], np.complex128)
        tabs = [kappa, np.conj(kappa), 1.0 / kappa, np.conj(1.0 / kappa)]
        mat, adj_mat, inv_mat, adj_inv_mat = [(lambda t: lambda dst, src: g.scale_per_coordinate(dst, src, t, 0))(t) for t in tabs]
        return g.matrix_operator(mat=mat, adj_mat=adj_mat, inv_mat=inv_mat, adj_inv_mat=adj_inv_mat, vector_space=self.vector_space_F_eo)


@g.params_convention(omega=None, mass=None, mass_plus=None, mass_minus=None, b=None, c=None, M5=None, boundary_phases=None,
                     link_compression=None)
def zmobius(U, params):
    """g.qcd.fermion.zmobius (lib/gpt/qcd/fermion/zmobius.py:62-74): Moebius with complex, s-dependent coefficients
    b_s, c_s = 1/2 ((b + c) / omega_s +- (b - c)); Ls = len(omega)"""
    params = copy.deepcopy(params)
    params["omega"] = [complex(w) for w in params["omega"]]
    params["Ls"] = len(params["omega"])
    return zmobius_class_operator("zmobius", U, params, otype=g.ot_vector_spin_color(4, 3))

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


@g.params_convention(mass=None, mass_plus=None, mass_minus=None, b=None, c=None, M5=None, boundary_phases=None, Ls=None)
def mobius(U, params):
    params = copy.deepcopy(params)
    return mobius_class_operator("mobius", U, params, otype=g.ot_vector_spin_color(4, 3))


@g.params_convention(omega=None, mass=None, mass_plus=None, mass_minus=None, b=None, c=None, M5=None, boundary_phases=None)
def zmobius(U, params):
    """g.qcd.fermion.zmobius (lib/gpt/qcd/fermion/zmobius.py:62-74): Moebius with complex, s-dependent coefficients
    b_s, c_s = 1/2 ((b + c) / omega_s +- (b - c)); Ls = len(omega)"""
    params = copy.deepcopy(params)
    params["omega"] = [complex(w) for w in params["omega"]]
    params["Ls"] = len(params["omega"])
    return mobius_class_operator("zmobius", U, params, otype=g.ot_vector_spin_color(4, 3))

from gpt_b200.qcd.fermion import operator
from gpt_b200.qcd.fermion.operator import fine_operator, interface, differentiable_fine_operator, gauge_independent_g5_hermitian
from gpt_b200.qcd.fermion.boundary_conditions import apply_open_boundaries
from gpt_b200.qcd.fermion import reference
from gpt_b200.qcd.fermion.wilson import wilson_clover, wilson_twisted_mass
from gpt_b200.qcd.fermion.mobius import mobius, zmobius
from gpt_b200.qcd.fermion import preconditioner

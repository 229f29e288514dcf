from gpt_b200.qcd.fermion.operator import fine_operator, interface
from gpt_b200.qcd.fermion.wilson import wilson_clover, wilson_twisted_mass
from gpt_b200.qcd.fermion.mobius import mobius, zmobius
from gpt_b200.qcd.fermion import preconditioner

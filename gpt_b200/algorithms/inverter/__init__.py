from gpt_b200.algorithms.inverter.cg import cg
from gpt_b200.algorithms.inverter.preconditioned import preconditioned
from gpt_b200.algorithms.inverter.mixed_precision import mixed_precision
from gpt_b200.algorithms.inverter.defect_correcting import defect_correcting

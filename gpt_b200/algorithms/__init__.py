from gpt_b200.algorithms import inverter, preconditioner
from gpt_b200.algorithms.base import base, base_iterative

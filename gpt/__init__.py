"""
`import gpt as g` -- the reference's package name, served by gpt_b200.

A script written for GPT (benchmarks/dslash.py, benchmarks/wilson_clover_dslash.py, the README example, the fermion-operator
tests) runs unchanged with this directory on PYTHONPATH: the module object it gets IS gpt_b200, and every sub-module is
reachable under both names (gpt.qcd.fermion, gpt.algorithms.inverter, gpt.default, ...).
"""
import sys

import gpt_b200

for _name, _mod in list(sys.modules.items()):
    if _name == "gpt_b200" or _name.startswith("gpt_b200."):
        sys.modules["gpt" + _name[len("gpt_b200"):]] = _mod
sys.modules[__name__] = gpt_b200

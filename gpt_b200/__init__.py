"""
gpt_b200 -- B200-native drop-in for the fermion-operator hot path of GPT (lehner/gpt).

Usage mirrors the reference (`import gpt as g`):  import gpt_b200 as g
The arithmetic lives in gpt_b200/lib/libcgpt_b200.so (CUDA, sm_100a) behind the C ABI of include/cgpt_b200.h.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:  # the signature-exact `cgpt` module lives next to this package (./cgpt)
    _sys.path.insert(0, _root)
from gpt_b200 import capi  # ctypes binding of the C ABI (include/cgpt_b200.h)

cgpt = capi  # g.cgpt.<extension>: entry points of the library beyond the reference's cgpt boundary (timers, host pipeline, ...)
_sys.modules[__name__ + ".cgpt"] = capi
from gpt_b200.params import params_convention
from gpt_b200.core import *  # noqa: F401,F403
from gpt_b200.core import eval, slice, time, complex  # noqa: F401,A004  (GPT's names shadow builtins on purpose)
from gpt_b200 import algorithms, qcd
from gpt_b200.random import random  # noqa: F401
from gpt_b200.gamma import gamma  # noqa: F401
from gpt_b200.io import load, save, format  # noqa: F401,A004
from gpt_b200 import default  # noqa: F401
from gpt_b200.timer import timer  # noqa: F401
from gpt_b200 import local_stencil, stencil  # noqa: F401
import sys as _sys


class _callable_module(_sys.modules[__name__].__class__):
    # g(expr) == g.eval(expr)   (lib/gpt/__init__.py)
    def __call__(self, first, second=None, ac=False):
        return eval(first, second, ac)


_sys.modules[__name__].__class__ = _callable_module

cgpt = capi  # (the star import brought core's reference-exact `cgpt` along; g.cgpt is the extension-level binding)

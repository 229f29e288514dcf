"""g.load / g.save for the gauge-configuration wire format of the hot path's callers (NERSC archive format)."""
from gpt_b200.io import nersc
from gpt_b200.io.nersc import load, save, format  # noqa: F401,A004

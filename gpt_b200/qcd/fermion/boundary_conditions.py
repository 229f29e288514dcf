"""g.qcd.fermion.apply_open_boundaries (lib/gpt/qcd/fermion/boundary_conditions.py:23-31): open boundary conditions in time
set a field to zero on the first and the last time slice."""
import numpy as np

import gpt_b200 as g


def apply_open_boundaries(field):
    nt = field.grid.fdimensions[-1]
    keep = np.ones(nt, dtype=np.complex128)
    keep[0] = keep[nt - 1] = 0.0
    g.scale_per_coordinate(field, g.copy(field), keep, field.grid.nd - 1)
    return field

#!/usr/bin/env python3
"""
Golden vectors for BASELINE.json configs[1] (the README example, /root/reference/README.md:132-167):
8^4 double, g.random("seed text"), gauge.random (scale 1), Moebius mass=0.1 M5=1.8 b=1 c=0 Ls=24, antiperiodic in time,
eo2_ne CG (eps 1e-4, maxiter 1000), point source at the origin, 12 spin-colour columns, pion correlator.

Produced by the CPU oracle (oracle/qcd.py: the numpy restatement of the reference's operator and solver stack, pinned to the
reference's own golden numbers by tests/test_oracle_golden.py).  The numpy CG needs ~4 minutes per column, so the result is
committed as a fixture instead of being recomputed by the GPU test:

    python tests/golden/make_readme_vectors.py        # ~45 min on 8 cores, writes tests/golden/readme_mobius_ls24.json

Content: CG iteration count per column, the correlator, and the propagator on a fixed sample of 24 sites (all 144 spin-colour
entries), enough to pin the solution of every column.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import qcd  # noqa: E402
from oracle.rng import random as oracle_random  # noqa: E402

DIMS = [8, 8, 8, 8]
PARAMS = dict(mass=0.1, M5=1.8, b=1.0, c=0.0, Ls=24, boundary_phases=[1.0, 1.0, 1.0, -1.0])
SAMPLE = [(0, 0, 0, 0), (1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1), (7, 7, 7, 7), (3, 5, 2, 6), (4, 4, 4, 4),
          (2, 0, 7, 3), (5, 1, 1, 0), (6, 3, 0, 7), (1, 2, 3, 4), (7, 0, 0, 0), (0, 7, 0, 0), (0, 0, 7, 0), (0, 0, 0, 7),
          (3, 3, 3, 3), (2, 6, 5, 1), (4, 0, 4, 0), (5, 5, 0, 2), (1, 7, 6, 5), (6, 6, 6, 2), (2, 2, 2, 6), (7, 1, 3, 5)]


def main():
    rng = oracle_random("seed text")
    U = qcd.gauge_random(rng, DIMS)
    op = qcd.mobius(U, **PARAMS)
    prop = np.zeros(tuple(reversed(DIMS)) + (4, 3, 12), dtype=np.complex128)  # [t,z,y,x,spin,color,column]
    iterations = []
    for col in range(12):
        s4 = np.zeros(tuple(reversed(DIMS)) + (4, 3), dtype=np.complex128)
        s4[0, 0, 0, 0].reshape(12)[col] = 1.0
        sol, hist = qcd.propagator_column(op, s4, 1e-4, 1000)
        prop[..., col] = sol
        iterations.append(len(hist))
        print(f"column {col}: {len(hist)} iterations", flush=True)
    corr = (np.abs(prop) ** 2).sum(axis=(1, 2, 3, 4, 5, 6))
    sample = {}
    for (x, y, z, t) in SAMPLE:
        v = prop[t, z, y, x].reshape(12, 12)  # [spin*3+color (row), column]
        sample[f"{x},{y},{z},{t}"] = [[float(v[i, j].real), float(v[i, j].imag)] for i in range(12) for j in range(12)]
    out = {"source": "oracle/qcd.py propagator_column on the README example (README.md:132-167); see this script's docstring",
           "dims": DIMS, "params": PARAMS, "seed": "seed text", "eps": 1e-4, "iterations": iterations,
           "correlator": [float(c) for c in corr], "sample_layout": "[row = spin*3+color][column], (re, im)", "sample": sample}
    with open(os.path.join(HERE, "readme_mobius_ls24.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()

"""Wilson-clover Dhop at 32^3 x 64 in single precision: 12 right-hand sides one after the other (L1 kernel, Ls = 1) against the
multi-rhs operator wilson_clover(n_rhs=12) (TMA sweep kernel).  Prints one JSON line."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpt_b200 as g
from gpt_b200 import cgpt

def main():
    cgpt.init(0)
    dims = [int(x) for x in os.environ.get("DIMS", "32.32.32.64").split(".")]
    n_rhs, steps = 12, int(os.environ.get("STEPS", "50"))
    grid = g.grid(dims, g.single)
    rng = g.random("benchmark", "vectorized_ranlux24_24_64")
    U = g.qcd.gauge.random(grid, rng, scale=0.5)
    params = dict(mass=0.08, csw_r=1.0, csw_t=1.0, xi_0=1.0, nu=1.0, isAnisotropic=False, boundary_phases=[1.0, 1.0, 1.0, -1.0])
    w1 = g.qcd.fermion.wilson_clover(U, dict(params))
    wn = g.qcd.fermion.wilson_clover(U, dict(params, n_rhs=n_rhs))
    src = g.vspincolor(grid); rng.cnormal(src); dst = g.vspincolor(grid)
    src5 = g.vspincolor(wn.F_grid); rng.cnormal(src5); dst5 = g.vspincolor(wn.F_grid)
    out = {}
    for name, op, d, s, reps in [("single_rhs_x12", w1.Dhop, dst, src, n_rhs), ("multi_rhs_12", wn.Dhop, dst5, src5, 1)]:
        for _ in range(5):
            op.mat(d, s)
        cgpt.accelerator_barrier()
        cgpt.timer_start()
        for _ in range(steps * reps):
            op.mat(d, s)
        ms = cgpt.timer_stop() / steps
        out[name] = {"ms_per_12_columns": ms, "gflops": 1320 * np.prod(dims) * n_rhs / (ms * 1e-3) / 1e9}
    out["speedup"] = out["single_rhs_x12"]["ms_per_12_columns"] / out["multi_rhs_12"]["ms_per_12_columns"]
    print(json.dumps(out))

main()

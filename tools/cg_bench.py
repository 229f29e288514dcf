"""eo2_ne CG time-to-solve on the bench lattice (Moebius 32^3x64 Ls=12 single): prints one JSON line."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpt_b200 as g
from gpt_b200 import cgpt
import bench

def main():
    dims = [int(x) for x in os.environ.get("CG_DIMS", "32.32.32.64").split(".")]
    Ls = int(os.environ.get("CG_LS", "12"))
    maxiter = int(os.environ.get("CG_MAXITER", "100"))
    cgpt.init(0)
    grid = g.grid(dims, g.single)
    rng = g.random("benchmark", "vectorized_ranlux24_24_64")
    U = g.qcd.gauge.random(grid, rng, scale=0.5)
    p = dict(bench.MOBIUS); p["Ls"] = Ls
    qm = g.qcd.fermion.mobius(U, p)
    src = g.vspincolor(qm.F_grid)
    rng.cnormal(src)
    half = g.vspincolor(qm.F_grid_eo)
    g.pick_checkerboard(g.odd, half, src)
    psi = g.lattice(half); psi[:] = 0
    # warm up
    cgpt.cg_eo2_ne(qm.interface.obj, psi.obj, half.obj, 1e-30, 3)
    psi[:] = 0
    cgpt.accelerator_barrier()
    l0 = cgpt.launch_count()
    cgpt.timer_start()
    hist, conv = cgpt.cg_eo2_ne(qm.interface.obj, psi.obj, half.obj, 1e-30, maxiter)
    ms = cgpt.timer_stop()
    print(json.dumps({"cg_iterations": len(hist), "ms_total": ms, "ms_per_iteration": ms / len(hist), "launches_per_iteration": (cgpt.launch_count() - l0) / len(hist), "res_first": hist[0], "res_last": hist[-1], "dims": dims, "Ls": Ls}))

main()

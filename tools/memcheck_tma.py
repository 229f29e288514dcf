"""small driver for compute-sanitizer: Moebius Dhop through the TMA sweep kernel and the pipelined host call on 8x8x8x8, Ls = 8
  compute-sanitizer --tool memcheck python tools/memcheck_tma.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpt_b200 as g

g.cgpt.init(0)
dims = [8, 8, 8, 8]
grid = g.grid(dims, g.single)
rng = g.random("memcheck", "vectorized_ranlux24_24_64")
U = g.qcd.gauge.random(grid, rng, scale=0.5)
m = g.qcd.fermion.mobius(U, dict(mass=0.08, M5=1.8, b=1.5, c=0.5, Ls=8, boundary_phases=[1.0, 1.0, 1.0, -1.0]))
src = rng.cnormal(g.vspincolor(m.F_grid))
os.environ["CGPTB_TMA_GRID"] = "3"
os.environ["CGPTB_TMA_TRL"] = "2"
a = g(m.Dhop * src)
b = g(m.Dhop.adj() * src)
h_in = np.ascontiguousarray(src[:])
h_out = np.zeros_like(h_in)
m.Dhop_host(h_out, h_in)
print("norms", g.norm2(a), g.norm2(b), float(np.linalg.norm(h_out) ** 2))

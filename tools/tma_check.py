"""A/B of the TMA sweep kernel (dslash_tma.cu) against the register/L1 kernels (dslash_f32.cu) on the GPU.

  python tools/tma_check.py            correctness on small lattices (several schedules), then timing at 32^3x64x12
  CHECK_ONLY=1 / TIME_ONLY=1           one of the two

Correctness: same synthetic fields through both kernels (CGPTB_NO_TMA switches per call), relative L2 difference must be
at fp32 rounding level (the summation order of the eight hops differs).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import gpt_b200 as g
from gpt_b200 import cgpt


def synthetic_fields_device(dims, ls, seed):
    """links = exp(i * 0.5 * sum_a u_a T_a), u_a ~ U[-1/2,1/2); source = N(0,1) + i N(0,1), generated on the device with torch's RNG
    (quick inputs for A/B runs; bench.py and the tests draw GPT's own RANLUX stream)"""
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    v4 = int(np.prod(dims))
    T = torch.tensor(bench.gell_mann_half(), dtype=torch.complex64, device="cuda")
    U = []
    for mu in range(4):
        u = torch.rand((v4, 8), generator=gen, device="cuda", dtype=torch.float32) - 0.5
        A = torch.einsum("na,aij->nij", (0.5 * u).to(torch.complex64), T)
        U.append(torch.linalg.matrix_exp(1j * A).contiguous())
    src = torch.randn((v4 * ls, 4, 3, 2), generator=gen, device="cuda", dtype=torch.float32)
    return U, torch.view_as_complex(src).contiguous()


def setup(dims, Ls, seed):
    grid = g.grid(dims, g.single)
    U_t, src_t = synthetic_fields_device(dims, Ls, seed)
    U = []
    for mu in range(4):
        u = g.mcolor(grid)
        cgpt.lattice_import_device(u.obj, U_t[mu].data_ptr(), U_t[mu].numel() * 8)
        U.append(u)
    p = dict(bench.MOBIUS)
    p["Ls"] = Ls
    p["boundary_phases"] = [1.0, -1.0, 1.0, -1.0]
    qm = g.qcd.fermion.mobius(U, p)
    src = g.vspincolor(qm.F_grid)
    cgpt.lattice_import_device(src.obj, src_t.data_ptr(), src_t.numel() * 8)
    return qm, src


def apply(qm, src, dag, env):
    for k in [k for k in os.environ if k.startswith("CGPTB_")]:
        os.environ.pop(k, None)
    os.environ.update(env)
    dst = g.vspincolor(qm.F_grid)
    dst[:] = 0
    op = qm.Dhop.adj() if dag else qm.Dhop
    op.mat(dst, src)
    cgpt.accelerator_barrier()
    out = np.array(dst[:])
    for k in env:
        os.environ.pop(k, None)
    return out


def check():
    ok = True
    for dims, Ls in [([8, 8, 8, 16], 12), ([16, 8, 12, 8], 8), ([8, 12, 8, 6], 4), ([16, 16, 16, 16], 12)]:
        qm, src = setup(dims, Ls, 11)
        for dag in (False, True):
            ref = apply(qm, src, dag, {"CGPTB_NO_TMA": "1"})
            for env in ({}, {"CGPTB_TMA_GRID": "1", "CGPTB_TMA_G": "3"}, {"CGPTB_TMA_GRID": "5", "CGPTB_TMA_G": "3"}, {"CGPTB_TMA_GRID": "7", "CGPTB_TMA_G": "2"},
                        {"CGPTB_TMA_G": "3"}, {"CGPTB_TMA_SCHED": "0", "CGPTB_TMA_GRID": "1", "CGPTB_TMA_TRL": "2"}, {"CGPTB_TMA_SCHED": "0", "CGPTB_TMA_GRID": "5", "CGPTB_TMA_TRL": "4", "CGPTB_TMA_G": "3"},
                        {"CGPTB_TMA_SCHED": "0", "CGPTB_TMA_GRID": "3", "CGPTB_TMA_TRL": "1", "CGPTB_TMA_G": "2"}, {"CGPTB_TMA_SCHED": "0", "CGPTB_TMA_TRL": "64"}):
                got = apply(qm, src, dag, dict(env))
                err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
                mx = np.abs(got - ref).max()
                good = err < 2e-6
                ok &= good
                print(f"CHECK dims={dims} Ls={Ls} dag={dag} env={env} rel={err:.3e} max={mx:.3e} {'ok' if good else 'FAIL'}", flush=True)
        # checkerboarded call (DhopEO on a half field: plane stride = half)
        half = g.vspincolor(qm.F_grid_eo)
        for cb in (g.even, g.odd):
            g.pick_checkerboard(cb, half, src)
            outs = []
            for env in ({"CGPTB_NO_TMA": "1"}, {}):
                os.environ.pop("CGPTB_NO_TMA", None)
                os.environ.update(env)
                o = g.lattice(half)
                qm.DhopEO.mat(o, half)
                cgpt.accelerator_barrier()
                outs.append(np.array(o[:]))
                os.environ.pop("CGPTB_NO_TMA", None)
            err = np.linalg.norm(outs[1] - outs[0]) / np.linalg.norm(outs[0])
            good = err < 2e-6
            ok &= good
            print(f"CHECK-EO dims={dims} Ls={Ls} cb={cb.tag if hasattr(cb, 'tag') else cb} rel={err:.3e} {'ok' if good else 'FAIL'}", flush=True)
        del qm, src
    print("CHECK RESULT", "PASS" if ok else "FAIL", flush=True)
    return ok


def timing():
    dims, Ls = bench.DIMS, bench.LS
    if os.environ.get("DIMS"):
        dims = [int(x) for x in os.environ["DIMS"].split(".")]
    qm, src = setup(dims, Ls, 5)
    dst = g.vspincolor(qm.F_grid)
    v5 = int(np.prod(dims)) * Ls
    v4 = int(np.prod(dims))
    bytes_per_launch = (v5 // 2) * 48 * 4 + (v4 // 2) * 8 * 18 * 4
    # VARIANTS="name:K=V,K=V;name2:..." overrides the default sweep
    variants = [("old", {"CGPTB_NO_TMA": "1"})]
    if os.environ.get("VARIANTS"):
        variants = []
        for item in os.environ["VARIANTS"].split(";"):
            name, _, kv = item.partition(":")
            variants.append((name, dict(x.split("=") for x in kv.split(",") if x)))
    else:
        if os.environ.get("ABLATE"):
            variants += [("tma compute-only", {"CGPTB_ABLATE": "1"}), ("tma memory-only", {"CGPTB_ABLATE": "2"}),
                         ("tma G=3 compute-only", {"CGPTB_ABLATE": "1", "CGPTB_TMA_G": "3"}), ("tma G=3 memory-only", {"CGPTB_ABLATE": "2", "CGPTB_TMA_G": "3"})]
        variants += [("tma G=1 (one chunk per CTA)", {}), ("tma G=3 (all Ls per CTA)", {"CGPTB_TMA_G": "3"}),
                     ("tma G=1 sched0 trl16", {"CGPTB_TMA_SCHED": "0"}), ("tma G=1 grid144", {"CGPTB_TMA_GRID": "144"})]
    steps = int(os.environ.get("STEPS", "200"))
    for rnd in range(int(os.environ.get("ROUNDS", "2"))):
        for name, env in variants:
            for k in list(os.environ):
                if k.startswith("CGPTB_"):
                    os.environ.pop(k)
            os.environ.update(env)
            for _ in range(10):
                qm.Dhop.mat(dst, src)
            cgpt.accelerator_barrier()
            cgpt.timer_start()
            for _ in range(steps):
                qm.Dhop.mat(dst, src)
            ms = cgpt.timer_stop() / steps
            gbs = bytes_per_launch / (ms * 1e-3 / 2) / 1e9
            print(f"TIME {name}: {ms:.4f} ms/step  {gbs:.0f} GB/s  frac {gbs / 6553:.3f}", flush=True)
    for k in [k for k in os.environ if k.startswith("CGPTB_")]:
        os.environ.pop(k, None)


if __name__ == "__main__":
    cgpt.init(0)
    ok = True
    if not os.environ.get("TIME_ONLY"):
        ok = check()
    if ok and not os.environ.get("CHECK_ONLY"):
        timing()

"""
Multi-GPU parity check, launched as
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/mgpu_check.py --mpi 1.1.1.N
Every rank builds the same global gauge field / source with the oracle RNG, keeps its local block, applies the
operators through gpt_b200 on its GPU (halo exchange inside libcgpt_b200) and compares with the matching block
of the oracle's global result.  Also checks that the eo2_ne CG takes exactly the oracle's iteration count.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def local_block(a, dims, mpi, coor, five_d):
    """global oracle-layout array [T,Z,Y,X,(S),...] -> this rank's block"""
    sl = []
    for mu in (3, 2, 1, 0):
        n = dims[mu] // mpi[mu]
        sl.append(slice(coor[mu] * n, (coor[mu] + 1) * n))
    return np.ascontiguousarray(a[tuple(sl)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mpi", default=None)
    ap.add_argument("--dims", default="8.8.8.16")
    ap.add_argument("--Ls", type=int, default=6, help="Ls = 8 routes the single-precision Dslash through the TMA sweep kernel")
    ap.add_argument("--only", default="", help="comma list of parts to run: mobius, clover, cg, solve (default: all)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import gpt_b200 as g
    from gpt_b200 import parallel
    from oracle import qcd
    from oracle.rng import random as oracle_random

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    g.cgpt.init(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mpi = [int(x) for x in args.mpi.split(".")] if args.mpi else None
    parallel.setup(dist, mpi)
    mpi = parallel.mpi
    rank = dist.get_rank()
    coor = parallel.processor_coor(rank, mpi)
    dims = [int(x) for x in args.dims.split(".")]

    rng = oracle_random("mgpu")
    U = qcd.gauge_random(rng, dims, scale=0.7)
    Ls = args.Ls
    phases = [1.0, -1.0, np.exp(0.4j), -1.0]
    failures = []
    parts = set(x for x in args.only.split(",") if x) or {"mobius", "clover", "cg", "solve"}

    def check(tag, got, ref, tol):
        err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        ok = err < tol
        if rank == 0 or not ok:
            print(f"[rank {rank}] {tag}: rel err {err:.3e} {'ok' if ok else 'FAIL'}", flush=True)
        if not ok:
            failures.append(tag)

    for prec, tol in [(g.double, 1e-12), (g.single, 1e-5)] if "mobius" in parts else []:
        cdt = prec.complex_dtype
        grid = g.grid(dims, prec)
        Ul = [local_block(u, dims, mpi, coor, False).astype(cdt) for u in U]
        Ug = g.qcd.gauge.from_numpy(grid, [u.reshape(-1, 3, 3) for u in Ul])
        Uo = [u.astype(cdt) for u in U]
        # Moebius
        params = dict(mass_plus=0.08, mass_minus=0.11, M5=1.8, b=1.5, c=0.5, Ls=Ls, boundary_phases=phases)
        op = g.qcd.fermion.mobius(Ug, dict(params))
        oo = qcd.mobius(Uo, **params)
        s5 = rng.cnormal([Ls] + dims, (4, 3)).astype(cdt)
        src = g.vspincolor(op.F_grid)
        src[:] = local_block(s5, dims, mpi, coor, True).reshape(-1, 4, 3)
        for tag, mat, ref in [("Dhop", op.Dhop, oo.Dhop(s5)), ("DhopDag", op.Dhop.adj(), oo.Dhop(s5, dag=True)),
                              ("M", op, oo.M(s5)), ("Mdag", op.adj(), oo.Mdag(s5))]:
            got = g(mat * src)[:]
            refl = local_block(ref, dims, mpi, coor, True).reshape(got.shape)
            check(f"mobius {prec.__name__} {tag}", got, refl, tol)
        if prec is g.single and Ls % 4 == 0:
            # host-buffer call: on a lattice split in t the slab pipeline runs with one halo exchange at the end; on a z split
            # it falls back to import -> apply -> export
            h_in = np.ascontiguousarray(local_block(s5, dims, mpi, coor, True))
            h_out = np.zeros_like(h_in)
            op.Dhop_host(h_out, h_in)
            check("mobius single Dhop_host", h_out, local_block(oo.Dhop(s5), dims, mpi, coor, True), tol)
        # eo: Meooe on both parities
        e = qcd.eo_ops(oo)
        for cb in [g.even, g.odd]:
            half = g.vspincolor(op.F_grid_eo)
            g.pick_checkerboard(cb, half, src)
            out = g(op.Meooe * half)
            full = g.vspincolor(op.F_grid)
            full[:] = 0
            g.set_checkerboard(full, out)
            ref = e.Meooe(e.proj(s5, cb.tag), cb.tag)
            check(f"mobius {prec.__name__} Meooe {cb.__name__}", full[:], local_block(ref, dims, mpi, coor, True).reshape(-1, 4, 3), tol)
        # global reductions
        n2 = g.norm2(src)
        ref = qcd.norm2(s5)
        if abs(n2 - ref) / ref > 1e-12:
            failures.append("norm2")
        # plain Wilson
        wp = dict(kappa=0.137, csw_r=0.0, csw_t=0.0, xi_0=1.0, nu=1.0, isAnisotropic=False, boundary_phases=phases)
        w = g.qcd.fermion.wilson_clover(Ug, dict(wp))
        wo = qcd.wilson_clover(Uo, **wp)
        s4 = rng.cnormal(dims, (4, 3)).astype(cdt)
        src4 = g.vspincolor(grid)
        src4[:] = local_block(s4, dims, mpi, coor, False).reshape(-1, 4, 3)
        got = g(w * src4)[:]
        check(f"wilson {prec.__name__} M", got, local_block(wo.M(s4), dims, mpi, coor, False).reshape(got.shape), tol)

    # Wilson-clover on the decomposed lattice (the field strength reaches over the corners of a rank's volume: the blocks are
    # built from the all-gathered links), anisotropic, and with open boundary conditions in the global time direction
    for prec, tol in [(g.double, 1e-12), (g.single, 1e-5)] if "clover" in parts else []:
        cdt = prec.complex_dtype
        grid = g.grid(dims, prec)
        Ug = g.qcd.gauge.from_numpy(grid, [local_block(u, dims, mpi, coor, False).astype(cdt).reshape(-1, 3, 3) for u in U])
        Uo = [u.astype(cdt) for u in U]
        s4 = rng.cnormal(dims, (4, 3)).astype(cdt)
        src4 = g.vspincolor(grid)
        src4[:] = local_block(s4, dims, mpi, coor, False).reshape(-1, 4, 3)
        for name, cp in [("clover", dict(kappa=0.13565, csw_r=2.0171 / 2.0, csw_t=2.0171 / 2.0, xi_0=1.0, nu=1.0, isAnisotropic=False, boundary_phases=phases)),
                         ("clover aniso", dict(mass=-0.05, csw_r=1.3, csw_t=0.9, xi_0=1.7, nu=1.2, isAnisotropic=True, boundary_phases=[1.0, 1.0, 1.0, -1.0])),
                         ("clover open", dict(kappa=0.135, csw_r=1.978, csw_t=1.978, cF=1.3, xi_0=1.0, nu=1.0, isAnisotropic=False, boundary_phases=[1.0, 1.0, 1.0, 0.0]))]:
            w = g.qcd.fermion.wilson_clover(Ug, dict(cp))
            wo = qcd.wilson_clover(Uo, **cp)
            for tag, mat, ref in [("M", w, wo.M(s4)), ("Mdag", w.adj(), wo.Mdag(s4)), ("Mdiag", w.Mdiag, wo.Mooee(s4)),
                                  ("MooeeInv", None, wo.MooeeInv(s4))]:
                if mat is None:
                    # the inverse blocks, through the even-odd entry on both parities
                    full = g.vspincolor(grid)
                    full[:] = 0
                    for cb in [g.even, g.odd]:
                        half = g.vspincolor(w.F_grid_eo)
                        g.pick_checkerboard(cb, half, src4)
                        g.set_checkerboard(full, g(w.Mooee.inv() * half))
                    got = full[:]
                else:
                    got = g(mat * src4)[:]
                check(f"{name} {prec.__name__} {tag}", got, local_block(ref, dims, mpi, coor, False).reshape(got.shape), tol)

    # BASELINE.json configs[3] in small: Wilson-clover, defect-correcting mixed-precision (double outer / single inner) even-odd CG
    # on the decomposed lattice (tests/manual/mpi.py:104-110), against the oracle's double-precision solution
    if "solve" in parts:
        cp = dict(kappa=0.137, csw_r=1.1, csw_t=1.1, xi_0=1.0, nu=1.0, isAnisotropic=False, boundary_phases=[1.0, 1.0, 1.0, -1.0])
        grid = g.grid(dims, g.double)
        Ug = g.qcd.gauge.from_numpy(grid, [local_block(u, dims, mpi, coor, False).reshape(-1, 3, 3) for u in U])
        w = g.qcd.fermion.wilson_clover(Ug, dict(cp))
        wo = qcd.wilson_clover(U, **cp)
        s4 = rng.cnormal(dims, (4, 3))
        src4 = g.vspincolor(grid)
        src4[:] = local_block(s4, dims, mpi, coor, False).reshape(-1, 4, 3)
        inv = g.algorithms.inverter
        pc = g.qcd.fermion.preconditioner
        slv = inv.defect_correcting(inv.mixed_precision(inv.preconditioned(pc.eo2_ne(), inv.cg(eps=1e-4, maxiter=1000)), g.single, g.double),
                                    eps=1e-10, maxiter=20)
        dst = g(slv(w) * src4)
        res = (g.norm2(g(w * dst - src4)) / g.norm2(src4)) ** 0.5
        ref, hist = qcd.solve_eo2_ne(wo, s4, 1e-11, 2000)
        check("clover mixed-precision solve", dst[:], local_block(ref, dims, mpi, coor, False).reshape(-1, 4, 3), 1e-8)
        if rank == 0:
            print(f"clover mixed-precision solve: true residual {res:.3e} (oracle CG {len(hist)} iterations)", flush=True)
        if not res < 1e-9:
            failures.append(f"solve residual {res}")

    # eo2_ne CG in double: identical iteration count, same solution
    grid = g.grid(dims, g.double)
    Ug = g.qcd.gauge.from_numpy(grid, [local_block(u, dims, mpi, coor, False).reshape(-1, 3, 3) for u in U])
    params = dict(mass=0.1, M5=1.8, b=1.5, c=0.5, Ls=Ls, boundary_phases=[1.0, 1.0, 1.0, -1.0])
    op = g.qcd.fermion.mobius(Ug, dict(params))
    oo = qcd.mobius(U, **params)
    s5 = rng.cnormal([Ls] + dims, (4, 3))
    src = g.vspincolor(op.F_grid)
    src[:] = local_block(s5, dims, mpi, coor, True).reshape(-1, 4, 3)
    inv = g.algorithms.inverter
    for fused in [True, False] if "cg" in parts else []:
        if not fused:
            os.environ["GPT_B200_NO_FUSED"] = "1"
        cg = inv.cg(eps=1e-8, maxiter=400)
        slv = inv.preconditioned(g.qcd.fermion.preconditioner.eo2_ne(), cg)(op)
        dst = g(slv * src)
        if fused:
            ref, hist = qcd.solve_eo2_ne(oo, s5, 1e-8, 400)
        check(f"cg fused={fused} solution", dst[:], local_block(ref, dims, mpi, coor, True).reshape(-1, 4, 3), 1e-9)
        if len(cg.history) != len(hist):
            failures.append(f"cg iterations {len(cg.history)} vs {len(hist)}")
        if rank == 0:
            print(f"cg fused={fused}: {len(cg.history)} iterations (oracle {len(hist)})", flush=True)

    t = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(t)
    g.cgpt.comm_finalize()
    dist.destroy_process_group()
    if int(t.item()) != 0:
        print(f"[rank {rank}] FAILURES: {failures}", flush=True)
        sys.exit(1)
    if rank == 0:
        print("MGPU CHECK PASSED", mpi, flush=True)


if __name__ == "__main__":
    main()

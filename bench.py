#!/usr/bin/env python3
"""
bench.py -- headline benchmark of the fermion-operator hot path (BASELINE.json):
Moebius domain-wall `Dhop` (opcode 3001, the loop of /root/reference/benchmarks/dslash.py:62-70) on a
32^3 x 64 lattice with Ls = 12 in single precision per GPU (T-split weak scaling: global T = 64 * n_gpus).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One "step" = one application of Dhop to the full 5d field.  Prints ONE JSON line (rank 0).

What the line holds (native arm):
  value / ms_per_step   K steps timed on the device AFTER the loop has already run for >= --preheat seconds, i.e. with
                        sw_power_cap engaged (the sustained number); `first_window` is the same K steps timed right after
                        the W warm-up steps of a cold GPU (what a short run would have reported)
  parity                the result of the timed operator on THIS lattice (every rank's block of the split lattice, halo slices
                        included) against oracle/dslash_ref.c on >= 1e5 sampled 5d sites; the run fails above 1e-5
  roofline              compulsory bytes of the even-odd kernel / measured launch time / measured HBM peak
  e2e                   the same Dhop through the host-buffer call (upload, stencil, download inside the timed region)
  e2e_solve             a propagator column: 4d host source -> upload -> eo2_ne CG -> 4d solution -> download
  eo_cg                 fused device CG: ms per iteration; time_to_solve: defect_correcting(mixed_precision(eo2_ne CG)) to 1e-8
                        with the true residual, and the same solver stack on a small lattice next to the CPU oracle
  kernels               GB/s of the CG's vector kernels and s-direction operators on the half lattice
  cpu_baseline          oracle/dslash_ref.c (OpenMP, all host cores) on the same lattice, a bounded number of applications

Flop / byte accounting: SURVEY.md 8(d) -- 1320 flop per 5d site; compulsory bytes per output site of the
even-odd kernel = (24 in + 24 out + 144/Ls links) reals; GPT's "effective" bytes (benchmarks/dslash.py:55-62)
are reported next to it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = [32, 32, 32, 64]
LS = 12
FLOPS_PER_SITE = 1320  # 8*Nc*(7+16*Nc), benchmarks/dslash.py:53
MOBIUS = dict(mass=0.08, M5=1.8, b=1.5, c=0.5, Ls=LS, boundary_phases=[1.0, 1.0, 1.0, 1.0])  # benchmarks/dslash.py:30-40
WORKLOAD = "Mobius DWF Dhop 32^3x64 Ls=12 single per GPU (BASELINE.json configs[2]; T-split, global T=64*n_gpus)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class clock_sampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, pw, reasons = [], [], [], set()
        for ts, line in list(self.lines):
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


def numa_bind(local_rank):
    """run this rank (and therefore its pinned host buffers: first touch) on the NUMA node of its GPU"""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def gell_mann_half():
    """T_a = lambda_a / 2 (tr T_a T_b = delta_ab / 2); the ordering is irrelevant for synthetic links"""
    lam = np.zeros((8, 3, 3), dtype=np.complex64)
    lam[0, 0, 1] = lam[0, 1, 0] = 1
    lam[1, 0, 1], lam[1, 1, 0] = -1j, 1j
    lam[2, 0, 0], lam[2, 1, 1] = 1, -1
    lam[3, 0, 2] = lam[3, 2, 0] = 1
    lam[4, 0, 2], lam[4, 2, 0] = -1j, 1j
    lam[5, 1, 2] = lam[5, 2, 1] = 1
    lam[6, 1, 2], lam[6, 2, 1] = -1j, 1j
    lam[7, 0, 0] = lam[7, 1, 1] = 1 / np.sqrt(3)
    lam[7, 2, 2] = -2 / np.sqrt(3)
    return lam / 2


def pad_with_neighbour_faces(flat, dims, ncomp, mpi, rank, gather):
    """A rank's block of a field (GPT order: x fastest, then y, z, t; ncomp complex numbers per site) -> [T + 2, Z + 2, Y X ncomp],
    the block padded by one slice in t and in z with the facing slices of its neighbours in the processor grid mpi = 1.1.Z.T
    (rank = cz + Z * ct; periodic: a direction that is not split wraps onto the block itself).  gather(face) returns the
    list of that face from every rank, in rank order.  Corners stay zero: a one-hop stencil does not read them."""
    X, Y, Z, T = dims
    pz, pt = mpi[2], mpi[3]
    cz, ct = rank % pz, rank // pz
    row = Y * X * ncomp
    P = np.zeros((T + 2, Z + 2, row), dtype=flat.dtype)
    P[1:-1, 1:-1] = flat.reshape(T, Z, row)
    every = {k: gather(np.ascontiguousarray(v)) for k, v in
             (("t_lo", P[1, 1:-1]), ("t_hi", P[T, 1:-1]), ("z_lo", P[1:-1, 1]), ("z_hi", P[1:-1, Z]))}

    def nb(key, dz, dt):
        return every[key][((cz + dz) % pz) + pz * ((ct + dt) % pt)]

    P[0, 1:-1] = nb("t_hi", 0, -1)
    P[T + 1, 1:-1] = nb("t_lo", 0, +1)
    P[1:-1, 0] = nb("z_hi", -1, 0)
    P[1:-1, Z + 1] = nb("z_lo", +1, 0)
    return P


# ---- parity of the timed operator on the bench lattice ------------------------------------------------------------------
def parity_check(torch, dist, cgpt, U, src, dst, dims, mpi, rank, n_sites4=9000):
    """
    Compare dst = Dhop src (already computed on the device) with oracle/dslash_ref.c on sampled sites of this rank's
    block.  The block is padded by one slice in t and in z with the faces of the neighbours in the processor grid mpi =
    1.1.Z.T (its own, periodically, where a direction is not split), so boundary sites check the halo exchange; a quarter
    of the sample lies on the boundary slices of each split direction.  Returns (rel_err, n_5d_sites); the caller takes the
    max over ranks.
    """
    from oracle import cref

    world = int(np.prod(mpi))
    cref.set_num_threads(max(1, cref.host_cores() // max(1, min(world, 8))))
    X, Y, Z, T = dims
    assert mpi[0] == 1 and mpi[1] == 1, "parity check: T and Z splits"

    def padded(lat, ncomp):
        """local field in GPT order -> [T + 2, Z + 2, Y * X * ncomp] with the neighbours' faces"""
        flat = np.empty(T * Z * Y * X * ncomp, dtype=np.complex64)
        cgpt.lattice_export_ptr(lat.obj, flat.ctypes.data, flat.nbytes)

        def gather(face):
            if world == 1:
                return [face]
            mine = torch.from_numpy(np.ascontiguousarray(face)).cuda()
            ev = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype, device="cuda")
            dist.all_gather_into_tensor(ev, mine)

            class _Faces:  # only the two neighbours' faces are copied to the host
                def __getitem__(self, r):
                    return ev[r].cpu().numpy()

            return _Faces()

        return pad_with_neighbour_faces(flat, dims, ncomp, mpi, rank, gather)

    psi = padded(src, LS * 12)
    V = np.stack([padded(U[mu], 9) for mu in range(4)])

    rs = np.random.default_rng(1234 + rank)
    nb = n_sites4 // 8
    t_of = np.concatenate([np.zeros(nb, np.int64), np.full(nb, T - 1, np.int64), rs.integers(0, T, n_sites4 - 2 * nb)])
    z_of = rs.integers(0, Z, n_sites4)
    z_of[2 * nb:3 * nb] = 0
    z_of[3 * nb:4 * nb] = Z - 1
    xy = rs.integers(0, X * Y, n_sites4)
    key = np.unique((t_of * Z + z_of) * (X * Y) + xy)
    t_of, z_of, xy = key // (Z * X * Y), (key // (X * Y)) % Z, key % (X * Y)
    idx_pad = ((t_of + 1) * (Z + 2) + (z_of + 1)) * (X * Y) + xy
    ref = cref.dhop_sites([X, Y, Z + 2, T + 2], LS, V.reshape(4, -1, 3, 3), psi.reshape(-1, 4, 3), idx_pad)
    del psi, V
    out = np.empty(T * Z * Y * X * LS * 12, dtype=np.complex64)
    cgpt.lattice_export_ptr(dst.obj, out.ctypes.data, out.nbytes)
    got = out.reshape(T * Z * Y * X, LS * 12)[key].reshape(-1)
    ref = ref.reshape(-1)
    err = float(np.linalg.norm(got.astype(np.complex128) - ref) / np.linalg.norm(ref.astype(np.complex128)))
    return err, int(key.size) * LS


def allmax(torch, dist, x):
    if dist is None:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_native(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    numa_node = numa_bind(local_rank)

    import gpt_b200 as g
    from gpt_b200 import cgpt

    cgpt.init(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from gpt_b200 import parallel

    mpi = [1, 1, 1, world]  # BASELINE.json configs[2]: "1 GPU then T-split across 2/4/8 B200" (weak scaling in T)
    if args.mpi:
        mpi = [int(x) for x in args.mpi.split(".")]  # e.g. 1.1.2.4: T x Z split (north_star: "T (and then Z)")
        assert len(mpi) == 4 and int(np.prod(mpi)) == world, (mpi, world)
    if world > 1:
        parallel.setup(dist, mpi)
    dims = list(DIMS)  # local extents; weak scaling: the global lattice grows with the processor grid
    gdims = [d * m for d, m in zip(dims, mpi)]
    grid = g.grid(gdims, g.single)
    # inputs exactly as /root/reference/benchmarks/dslash.py:10-49: GPT's own generator (fast engine), SU(3) links
    # exp(i 0.5 sum_a u_a T_a), complex normal source; every rank draws its block of the global lattice
    rng = g.random("benchmark", "vectorized_ranlux24_24_64")
    U = g.qcd.gauge.random(grid, rng, scale=0.5)
    qm = g.qcd.fermion.mobius(U, dict(MOBIUS, link_compression=12) if args.compress else dict(MOBIUS))
    src = g.vspincolor(qm.F_grid)
    dst = g.vspincolor(qm.F_grid)
    rng.cnormal(src)
    del rng
    cgpt.accelerator_barrier()

    v5 = int(np.prod(dims)) * LS
    v4 = int(np.prod(dims))

    def sync():
        cgpt.accelerator_barrier()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(fn, n):
        """n calls of fn, device time on the library stream (CUDA events), max over ranks"""
        sync()
        cgpt.timer_start()
        for _ in range(n):
            fn()
        ms = cgpt.timer_stop()
        sync()
        return allmax(torch, dist, ms)

    def step():
        qm.Dhop.mat(dst, src)

    # ---- parity of exactly this operator / lattice / decomposition -------------------------------------------------
    parity = None
    if not args.no_parity:
        step()
        sync()
        err, n5 = parity_check(torch, dist, cgpt, U, src, dst, dims, mpi, rank)
        err = allmax(torch, dist, err)
        parity = {"rel_err": err, "tolerance": 1e-5, "sites_checked_per_rank": n5, "ranks": world,
                  "against": "oracle/dslash_ref.c on sampled sites of every rank's block (boundary slices included)",
                  "lattice": gdims + [LS]}
        if not err < 1e-5:
            raise SystemExit(f"parity check failed: rel err {err} on {gdims} Ls={LS} ({world} ranks)")

    # ---- the timed region ------------------------------------------------------------------------------------------
    sampler = clock_sampler(local_rank)
    for _ in range(args.warmup):
        step()
    l0 = cgpt.launch_count()
    tw0 = time.time()
    ms_first = timed(step, args.steps)  # a cold GPU: ends before the power cap engages if K is small
    tw1 = time.time()
    launches = cgpt.launch_count() - l0
    # pre-heat: keep the same loop running until it has run for --preheat seconds, then time K steps again
    n_pre = 0
    t_pre = time.time()
    while time.time() - t_pre < args.preheat:
        for _ in range(50):
            step()
        cgpt.accelerator_barrier()
        n_pre += 50
    ts0 = time.time()
    ms = timed(step, args.steps)
    ts1 = time.time()
    if ts1 - ts0 < 0.5:
        time.sleep(0.3)  # let nvidia-smi (50 ms period) see the tail of the region
    clocks = sampler.window(t_pre, ts1 + 0.1)
    clocks_first = sampler.window(tw0, tw1 + 0.05)
    sampler.stop()
    ms_per_step = ms / args.steps
    gflops = FLOPS_PER_SITE * v5 * world / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant kernel: one launch per parity, two per step
    # (two-row link compression: 12 reals of the link + its U(1) factor (2 reals) instead of 18)
    bytes_per_launch = (v5 // 2) * 48 * 4 + (v4 // 2) * 8 * (14 if args.compress else 18) * 4
    launches_per_step = 2
    peak, peak_src = peaks()
    achieved = bytes_per_launch / (ms_per_step * 1e-3 / launches_per_step) / 1e9
    eff_bytes = (8 * 2 * 4 * 3 + 8 * 2 * 9 / LS + 2 * 4 * 3) * 4 * v5
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "dhop_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_launch")
        traffic_src = "not measured in this run: ncu --set full capture committed as " + tj.get("source", "profiles/dhop_traffic.json")
    except Exception:
        pass

    # ---- what a plain copy sustains in the same (power-capped) state ------------------------------------------------------
    # MEASURED_PEAKS.json's hbm_gbs is a best-of-10 burst; a copy that runs for seconds sits at the 1000 W cap like the stencil
    # does.  Same method (torch copy_ of 2 GiB, read + write bytes), but 1.5 s back to back right after the timed loop.
    sustained_copy = None
    if not args.no_copy_peak:
        try:
            n_el = 1 << 30
            a_buf = torch.empty(n_el, dtype=torch.bfloat16, device="cuda")
            b_buf = torch.empty(n_el, dtype=torch.bfloat16, device="cuda")
            a_buf.zero_()
            torch.cuda.synchronize()
            t_c = time.time()
            while time.time() - t_c < 1.0:
                for _ in range(20):
                    b_buf.copy_(a_buf)
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100):
                b_buf.copy_(a_buf)
            e1.record()
            torch.cuda.synchronize()
            sustained_copy = 100 * 2 * n_el * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9
            sustained_copy = allmax(torch, dist, -sustained_copy) * -1.0  # the slowest rank
            del a_buf, b_buf
        except Exception as ex:  # out of memory on a shared box: the number is optional
            sustained_copy = None
            print("sustained copy peak not measured:", ex, file=sys.stderr)

    # ---- end to end through the public API with HOST buffers (pinned): import -> Dhop -> export ------------------------
    e2e = None
    if not args.no_e2e:
        nbytes = v5 * 12 * 8
        h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        cgpt.lattice_export_ptr(src.obj, h_in.data_ptr(), nbytes)
        n_e2e = max(3, min(args.steps, 10))
        # the public host-buffer call: op.Dhop_host(dst, src) == lattice[:] = src ; dst_l = Dhop * src_l ; dst = dst_l[:]
        # with upload / stencil / download pipelined over slabs of time slices (gpt_b200/csrc/hostpipe.cu)
        for _ in range(2):
            qm.Dhop_host(h_out, h_in)
        ms_e = timed(lambda: qm.Dhop_host(h_out, h_in), n_e2e)
        e2e = {"value": FLOPS_PER_SITE * v5 * world / (ms_e / n_e2e * 1e-3) / 1e9, "unit": "GFlop/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": n_e2e, "ms_per_step": ms_e / n_e2e,
               "numa_node": numa_node,
               "note": "one Dhop per upload+download: PCIe bound by construction; e2e_solve is the path as users run it"}
        del h_in, h_out

    # ---- GB/s of the other kernels of the CG on the half lattice -------------------------------------------------------------
    kernels = None
    if not args.no_kernels:
        a, b, c = (g.vspincolor(qm.F_grid_eo) for _ in range(3))
        g.pick_checkerboard(g.odd, a, src)
        g.pick_checkerboard(g.odd, b, dst)
        c[:] = 0
        reals = v5 // 2 * 4  # bytes per real per half-lattice site
        table = [
            ("axpy", lambda: g.axpy(c, 0.3, a, b), 72), ("inner_product", lambda: g.rank_inner_product(a, b), 48),
            ("norm2", lambda: cgpt.lattice_norm2(a.obj), 24), ("psi_plus_a_p", lambda: cgpt.lattice_lc(c.obj, True, [0.3], [a.obj]), 72),
            ("Mooee", lambda: qm.Mooee.mat(c, a), 48), ("MooeeInv", lambda: qm.Mooee.inv_mat(c, a), 48),
            ("Meooe", lambda: qm.Meooe.mat(c, a), 60),
        ]
        kernels = {}
        for name, fn, nreal in table:
            for _ in range(3):
                fn()
            msk = timed(fn, 20) / 20
            gbs = nreal * reals / (msk * 1e-3) / 1e9
            kernels[name] = {"ms": msk, "GB/s": gbs, "frac_of_hbm_peak": gbs / peak, "reals_per_site": nreal}
        del a, b, c

    # ---- eo-CG (second half of BASELINE.json's metric) ------------------------------------------------------------------------------
    cg_info = None
    if not args.no_cg:
        half = g.vspincolor(qm.F_grid_eo)
        g.pick_checkerboard(g.odd, half, src)
        psi = g.lattice(half)
        psi[:] = 0
        cgpt.cg_eo2_ne(qm.interface.obj, psi.obj, half.obj, 1e-30, 10)  # warm-up
        ms_cg = None
        tries = []
        for _ in range(3):  # best of three identical solves; all three are reported (host jitter shows in the two syncs per iteration)
            psi[:] = 0
            sync()
            l1 = cgpt.launch_count()
            cgpt.timer_start()
            hist, conv = cgpt.cg_eo2_ne(qm.interface.obj, psi.obj, half.obj, 1e-30, args.cg_iterations)
            ms_try = cgpt.timer_stop()
            sync()
            ms_cg = ms_try if ms_cg is None else min(ms_cg, ms_try)
            tries.append(round(ms_try / max(len(hist), 1), 3))
        ms_cg = allmax(torch, dist, ms_cg)
        cg_info = {"solver": "inv.preconditioned(pc.eo2_ne(), inv.cg) on Mpc^dag Mpc, fused device loop", "iterations": len(hist),
                   "ms_total": ms_cg, "ms_per_iteration": ms_cg / max(len(hist), 1),
                   "residual_reduction": (hist[-1] / hist[0]) ** 0.5 if hist else None,
                   "launches_per_iteration": (cgpt.launch_count() - l1) / max(len(hist), 1), "ms_per_iteration_all_tries": tries}
        del half, psi

    # ---- time to solve: wall time, iterations, true residual ----------------------------------------------------------
    solve = None
    e2e_solve = None
    if not args.no_solve:
        inv = g.algorithms.inverter
        pc = g.qcd.fermion.preconditioner
        # (a) a propagator column through the public API from a HOST source: upload once, solve, download once
        src4 = g.vspincolor(qm.U_grid)
        rng4 = g.random("bench_source", "vectorized_ranlux24_24_64")
        rng4.cnormal(src4)
        nb4 = v4 * 12 * 8
        h_src = torch.empty(nb4, dtype=torch.uint8, pin_memory=True)
        h_dst = torch.empty(nb4, dtype=torch.uint8, pin_memory=True)
        cgpt.lattice_export_ptr(src4.obj, h_src.data_ptr(), nb4)
        cg_sp = inv.cg(eps=args.solve_eps_single, maxiter=args.solve_maxiter)
        prop = qm.propagator(inv.preconditioned(pc.eo2_ne(), cg_sp))
        dst4 = g.vspincolor(qm.U_grid)
        sync()
        t0 = time.time()
        cgpt.lattice_import_ptr(src4.obj, h_src.data_ptr(), nb4)
        prop(dst4, src4)
        cgpt.lattice_export_ptr(dst4.obj, h_dst.data_ptr(), nb4)
        sync()
        wall = allmax(torch, dist, time.time() - t0)
        e2e_solve = {"what": "propagator column: 4d host source -> Import -> eo2_ne CG (single) -> Export -> 4d host solution",
                     "seconds": wall, "iterations": len(cg_sp.history), "eps": args.solve_eps_single,
                     "h2d_bytes": nb4, "d2h_bytes": nb4, "ms_per_iteration": wall * 1e3 / max(len(cg_sp.history), 1)}
        del h_src, h_dst, prop
        # (b) the production stack (tests/manual/mpi.py:104-110): double outer defect correction, single inner eo2_ne CG
        qd = qm.converted(g.double)
        src_d = g.convert(src4, g.double)
        b5 = g(qd.ImportPhysicalFermionSource * src_d)
        cg_in = inv.cg(eps=1e-4, maxiter=args.solve_maxiter)
        dc = inv.defect_correcting(inv.mixed_precision(inv.preconditioned(pc.eo2_ne(), cg_in), g.single, g.double),
                                   eps=args.solve_eps, maxiter=40)
        x5 = g.lattice(b5)
        x5[:] = 0
        sync()
        t0 = time.time()
        dc(qd)(x5, b5)
        sync()
        wall = allmax(torch, dist, time.time() - t0)
        r = g(qd * x5 - b5)
        true_res = (g.norm2(r) / g.norm2(b5)) ** 0.5
        solve = {"solver": "defect_correcting(mixed_precision(preconditioned(eo2_ne, cg(eps=1e-4)), single, double), eps=%g)" % args.solve_eps,
                 "lattice": gdims + [LS], "seconds": wall, "outer_iterations": len(dc.history),
                 "true_residual": float(true_res), "converged": bool(true_res < 10 * args.solve_eps)}
        del qd, src_d, b5, x5, r, src4, dst4
        # (c) the same stack on a lattice the numpy oracle finishes in seconds: GPU and CPU side by side
        if rank == 0 and world == 1:
            solve["small_lattice"] = small_solve_side_by_side(g)

    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_baseline(DIMS, seconds=args.cpu_seconds)
    if dist is not None:
        dist.barrier()

    if rank == 0:
        out = {
            "metric": "mobius_dwf_dslash_gflops", "value": gflops, "unit": "GFlop/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (g.random(\"benchmark\", ranlux24_24) links scale 0.5 + cnormal source, as benchmarks/dslash.py)",
            "config": {"workload": WORKLOAD,
                       "local_dims": dims, "Ls": LS, "cache": "inputs (2.4 GB field + 0.6 GB links) larger than L2, no flush needed",
                       "global_dims": gdims, "parallelism": "mpi " + ".".join(str(m) for m in mpi) + " (x.y.z.t), halo exchange overlapped with the interior stencil",
                       "timing": f"K steps timed after {n_pre} untimed steps (>= {args.preheat} s of the same loop): sustained clocks; first_window = the K steps right after warm-up"},
            "first_window": {"ms_per_step": ms_first / args.steps, "value": FLOPS_PER_SITE * v5 * world / (ms_first / args.steps * 1e-3) / 1e9,
                             "clocks": clocks_first},
            "parity": parity,
            "gbs_effective_gpt_convention": eff_bytes * world / (ms_per_step * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "k_dhop_f32_tma (persistent t-sweep; centre / face / link rings fed by four TMA producer warps, packed FFMA2; one launch per parity)", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "launches_per_step": launches_per_step,
                         "peak_sustained_copy": sustained_copy, "frac_of_sustained_copy": (achieved / sustained_copy) if sustained_copy else None,
                         "peak_sustained_copy_how": "torch b.copy_(a) over 1 Gi bf16 elements (read + write bytes), 100 copies timed after 1 s of the same loop right after the stencil loop: the copy a power-capped chip sustains; `peak` above is the burst figure the contract prescribes",
                         "frac_first_window": bytes_per_launch / (ms_first / args.steps * 1e-3 / launches_per_step) / 1e9 / peak},
            "cpu_baseline": cpu, "e2e": e2e, "e2e_solve": e2e_solve, "gpu_launches": int(launches), "clocks": clocks,
            "eo_cg": cg_info, "time_to_solve": solve, "kernels": kernels,
        }
        print(json.dumps(out))
    if dist is not None:
        cgpt.comm_finalize()
        dist.destroy_process_group()


def small_solve_side_by_side(g):
    """BASELINE.md 3.3 on a lattice the numpy oracle (oracle/qcd.py: cg.py restated) finishes in seconds: same operator,
    source and eps on the GPU (double, through inv.preconditioned(eo2_ne, cg)) and on the CPU; iterations must agree"""
    from oracle import qcd
    from oracle.rng import random as oracle_random

    dims, ls, eps = [8, 8, 8, 8], 8, 1e-8
    rng = oracle_random("bench_small")
    U = qcd.gauge_random(rng, dims, scale=0.5)
    params = dict(mass=0.08, M5=1.8, b=1.5, c=0.5, Ls=ls, boundary_phases=[1.0, 1.0, 1.0, -1.0])
    s_np = rng.cnormal([ls] + dims, (4, 3))
    t0 = time.time()
    ref, hist = qcd.solve_eo2_ne(qcd.mobius(U, **params), s_np, eps, 1000)
    t_cpu = time.time() - t0
    grid = g.grid(dims, g.double)
    op = g.qcd.fermion.mobius(g.qcd.gauge.from_numpy(grid, [u.reshape(-1, 3, 3) for u in U]), dict(params))
    src = g.vspincolor(op.F_grid)
    src[:] = s_np.reshape(-1, 4, 3)
    inv = g.algorithms.inverter
    cg = inv.cg(eps=eps, maxiter=1000)
    slv = inv.preconditioned(g.qcd.fermion.preconditioner.eo2_ne(), cg)(op)
    g.cgpt.accelerator_barrier()
    t0 = time.time()
    dst = g(slv * src)
    g.cgpt.accelerator_barrier()
    t_gpu = time.time() - t0
    r = g(op * dst - src)
    x = dst[:].reshape(ref.shape)
    return {"lattice": dims + [ls], "eps": eps, "gpu_seconds": t_gpu, "gpu_iterations": len(cg.history),
            "gpu_true_residual": float((g.norm2(r) / g.norm2(src)) ** 0.5),
            "cpu_seconds": t_cpu, "cpu_iterations": len(hist), "cpu_kind": "numpy oracle (oracle/qcd.py), 1 process",
            "solution_rel_diff": float(np.linalg.norm(x - ref) / np.linalg.norm(ref))}


def cpu_baseline(sample_dims, seconds, steps=None, warmup=1):
    """C/OpenMP restatement of Dhop (oracle/dslash_ref.c) on ALL host cores (torchrun exports OMP_NUM_THREADS=1, so the
    thread count is set explicitly), a bounded number of applications on the bench lattice"""
    from oracle import cref

    cref.set_num_threads(cref.host_cores())
    rs = np.random.default_rng(7)
    v4 = int(np.prod(sample_dims))
    # cheap unitary links: second-order exp is enough for a throughput sample; normalisation does not matter
    T = gell_mann_half()
    V = np.empty((4, v4, 3, 3), dtype=np.complex64)
    for mu in range(4):
        u = (rs.random((v4, 8), dtype=np.float32) - 0.5) * 0.5
        A = np.einsum("na,aij->nij", u.astype(np.complex64), T)
        V[mu] = np.eye(3, dtype=np.complex64) + 1j * A - 0.5 * (A @ A)
    psi = rs.standard_normal((v4 * LS, 4, 3, 2), dtype=np.float32).view(np.complex64).reshape(v4 * LS, 4, 3)
    for _ in range(warmup):
        cref.dhop(sample_dims, LS, V, psi)
    t0 = time.time()
    n = 0
    while True:
        cref.dhop(sample_dims, LS, V, psi)
        n += 1
        if steps is not None and n >= steps:
            break
        if steps is None and time.time() - t0 > seconds:
            break
    dt = (time.time() - t0) / n
    gf = FLOPS_PER_SITE * v4 * LS / dt / 1e9
    cpu_model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    cpu_model = line.split(":", 1)[1].strip()
                    break
    except Exception:
        pass
    return {"value": gf, "unit": "GFlop/s", "cores": cref.num_threads(), "kind": "port",
            "sample": f"{n} applications of Dhop on {list(sample_dims)} Ls={LS} single (oracle/dslash_ref.c, OpenMP; Grid unavailable)",
            "ms_per_step": dt * 1e3, "cpu": cpu_model, "nproc": os.cpu_count()}


def run_config(args):
    """The other configurations of BASELINE.json (parity-test cases and multi-GPU solves, not the driver's bench line):
      --config wilson_clover_16   configs[0]: Wilson-clover Dhop 16^4 double, parameters of benchmarks/wilson_clover_dslash.py:29-40;
                                  the lattice (0.1 GB) lives in L2, so L2 is flushed before every timed application
      --config clover_solve       configs[3]: Wilson-clover 48^3 x 96, defect-correcting mixed-precision even-odd CG (double
                                  outer / single inner, tests/manual/mpi.py:100-110) on a T x Z processor grid (--mpi 1.1.2.4)
      --config mobius_prop        configs[4]: Moebius 64^3 x 128 Ls = 12, 12 spin-colour columns of a point-source propagator
                                  (README.md:150-160 stack: eo2_ne CG), T-split
    --grid x.y.z.t sets the GLOBAL lattice (smaller ones for a quick check)."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    numa_bind(local_rank)
    import gpt_b200 as g
    from gpt_b200 import cgpt, parallel

    cgpt.init(local_rank)
    dist = None
    mpi = [int(x) for x in args.mpi.split(".")] if args.mpi else [1, 1, 1, world]
    assert int(np.prod(mpi)) == world, (mpi, world)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        parallel.setup(dist, mpi)

    def sync():
        cgpt.accelerator_barrier()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    defaults = {"wilson_clover_16": [16, 16, 16, 16], "clover_solve": [48, 48, 48, 96], "mobius_prop": [64, 64, 64, 128]}
    gdims = [int(x) for x in args.grid.split(".")] if args.grid else defaults[args.config]
    rng = g.random("benchmark", "vectorized_ranlux24_24_64")
    inv = g.algorithms.inverter
    pc = g.qcd.fermion.preconditioner
    clover = dict(mass=0.08, csw_r=1.0, csw_t=1.0, xi_0=1.0, nu=1, isAnisotropic=False, boundary_phases=[1, 1, 1, -1])
    out = {"config": {"workload": args.config, "global_dims": gdims, "parallelism": "mpi " + ".".join(str(m) for m in mpi)}, "n_gpus": world,
           "data": "synthetic (g.random(\"benchmark\") links of scale 0.5)"}
    if args.config == "wilson_clover_16":
        prec = g.single if args.single else g.double
        grid = g.grid(gdims, prec)
        qm = g.qcd.fermion.wilson_clover(g.qcd.gauge.random(grid, rng, scale=0.5), dict(clover))
        src, dst = g.vspincolor(grid), g.vspincolor(grid)
        rng.cnormal(src)
        v4 = int(np.prod(gdims))
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        stream = torch.cuda.ExternalStream(cgpt.get_stream()) if hasattr(cgpt, "get_stream") else torch.cuda.current_stream()
        for _ in range(5):
            qm.Dhop.mat(dst, src)
        sync()
        ms = 0.0
        with torch.cuda.stream(stream):
            for _ in range(args.steps):
                flush.fill_(1)  # 256 MB written: nothing of the lattice is left in the 126 MB L2
                cgpt.timer_start()
                qm.Dhop.mat(dst, src)
                ms += cgpt.timer_stop()
        ms /= args.steps
        flops = 8 * 3 * (7 + 16 * 3) * v4  # benchmarks/wilson_clover_dslash.py:53
        bytes_alg = (24 + 24 + 8 * 18) * (4 if args.single else 8) * v4  # spinor in + out, 8 double-stored links per site
        # the same without flushing: what benchmarks/wilson_clover_dslash.py measures on a lattice that fits L2
        sync()
        cgpt.timer_start()
        for _ in range(args.steps):
            qm.Dhop.mat(dst, src)
        ms_resident = cgpt.timer_stop() / args.steps
        peak, peak_src = peaks()
        out.update({"metric": "wilson_clover_dslash_gflops", "value": flops / (ms * 1e-3) / 1e9, "unit": "GFlop/s", "steps": args.steps,
                    "ms_per_step": ms, "ms_per_step_without_l2_flush": ms_resident, "dtype": "f32" if args.single else "f64", "higher_is_better": True,
                    "roofline": {"bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": bytes_alg / (ms * 1e-3) / 1e9 / peak, "kernel": "k_dhop<double> / k_dhop_f32 (one thread per site; two launches)",
                                 "algorithmic_bytes_per_step": bytes_alg, "peak_source": peak_src,
                                 "note": "L2 flushed before every timed application when the problem (0.1 GB at 16^4 double) fits L2"}})
    elif args.config == "clover_solve":
        grid = g.grid(gdims, g.double)
        U = g.qcd.gauge.random(grid, rng, scale=0.5)
        qm = g.qcd.fermion.wilson_clover(U, dict(clover))
        src = g.vspincolor(grid)
        rng.cnormal(src)
        cg_inner = inv.cg({"eps": 1e-4, "maxiter": 40000})
        slv = inv.defect_correcting(inv.mixed_precision(inv.preconditioned(pc.eo2_ne(), cg_inner), g.single, g.double), eps=args.solve_eps, maxiter=100)
        prop = slv(qm)
        dst = g(prop * src)  # warm-up solve (operators in both precisions are built here)
        sync()
        t0 = time.time()
        dst = g(prop * src)
        sync()
        t1 = time.time()
        res = (g.norm2(g(qm * dst - src)) / g.norm2(src)) ** 0.5
        out.update({"metric": "wilson_clover_mixed_precision_solve_seconds", "value": t1 - t0, "unit": "s", "higher_is_better": False, "dtype": "f64 outer / f32 inner",
                    "solver": "defect_correcting(mixed_precision(preconditioned(eo2_ne, cg(eps=1e-4)), single, double), eps=%g)" % args.solve_eps,
                    "true_residual": res, "inner_iterations_last_cycle": len(cg_inner.history), "converged": bool(res < 10 * args.solve_eps)})
    else:
        grid = g.grid(gdims, g.single)
        U = g.qcd.gauge.random(grid, rng, scale=0.5)
        qm = g.qcd.fermion.mobius(U, dict(MOBIUS, boundary_phases=[1.0, 1.0, 1.0, -1.0]))
        cg = inv.cg({"eps": args.solve_eps_single, "maxiter": 20000})
        prop = qm.propagator(inv.preconditioned(pc.eo2_ne(), cg))
        src = g.mspincolor(grid)
        g.create.point(src, [0, 0, 0, 0])
        sync()
        t0 = time.time()
        dst = g(prop * src)
        sync()
        t1 = time.time()
        corr = g.slice(g.trace(dst * g.adj(dst)), 3)
        out.update({"metric": "mobius_12_column_propagator_seconds", "value": t1 - t0, "unit": "s", "higher_is_better": False, "dtype": "f32", "Ls": LS,
                    "solver": "propagator(preconditioned(eo2_ne, cg(eps=%g))), 12 columns" % args.solve_eps_single, "iterations_last_column": len(cg.history),
                    "correlator_t0_t1": [float(corr[0].real), float(corr[1].real)]})
    if world > 1:
        cgpt.comm_finalize()
        dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)


def run_reference(args):
    """reference arm: the CPU restatement of the reference's Dhop on all host cores (Grid cannot be built here), on the
    same lattice as the native arm; under torchrun rank 0 alone runs it"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = cpu_baseline(DIMS, seconds=None, steps=args.steps, warmup=max(min(args.warmup, 2), 1))
    out = {
        "impl": "reference", "metric": "mobius_dwf_dslash_gflops", "value": c["value"], "unit": "GFlop/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": c["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "local_dims": DIMS, "Ls": LS,
                   "note": "each step = one Dhop on the full 32^3x64 Ls=12 lattice on the host cores (one lattice, whatever --gpus says)"},
        "cpu_baseline": {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu", "nproc")},
        "e2e": {"value": c["value"], "unit": "GFlop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--preheat", type=float, default=2.0, help="seconds the loop runs untimed before the K timed steps")
    ap.add_argument("--mpi", default=None, help="processor grid x.y.z.t (default 1.1.1.N: T-split)")
    ap.add_argument("--config", default="dslash", choices=["dslash", "wilson_clover_16", "clover_solve", "mobius_prop"],
                    help="dslash: the bench line (BASELINE.json configs[2]); the others: see run_config")
    ap.add_argument("--grid", default=None, help="global lattice x.y.z.t of a --config workload")
    ap.add_argument("--single", action="store_true", help="--config wilson_clover_16 in single precision")
    ap.add_argument("--compress", action="store_true", help="two-row SU(3) link compression (link_compression=12)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-copy-peak", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-kernels", action="store_true")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--cg-iterations", type=int, default=50)
    ap.add_argument("--solve-eps", type=float, default=1e-8)
    ap.add_argument("--solve-eps-single", type=float, default=1e-6)
    ap.add_argument("--solve-maxiter", type=int, default=1500)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.config != "dslash":
        run_config(a)
    else:
        run_native(a)

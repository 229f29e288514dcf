#!/usr/bin/env python3
"""
bench.py -- headline benchmark of the fermion-operator hot path (BASELINE.json):
Moebius domain-wall `Dhop` (opcode 3001, the loop of /root/reference/benchmarks/dslash.py:62-70) on a
32^3 x 64 lattice with Ls = 12 in single precision per GPU (T-split weak scaling: global T = 64 * n_gpus).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One "step" = one application of Dhop to the full 5d field.  Prints ONE JSON line (rank 0).
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

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def synthetic_fields_device(torch, dims, ls, seed):
    """links = exp(i * 0.5 * sum_a u_a T_a), u_a ~ U[-1/2,1/2); source = N(0,1) + i N(0,1)   (SURVEY 8(d)),
    generated on the device with torch's RNG (the reference's RANLUX stream is only needed for parity tests)."""
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    v4 = int(np.prod(dims))
    T = torch.tensor(gell_mann_half(), dtype=torch.complex64, device="cuda")
    U = []
    for mu in range(4):
        u = torch.rand((v4, 8), generator=gen, device="cuda", dtype=torch.float32) - 0.5
        A = torch.einsum("na,aij->nij", (0.5 * u).to(torch.complex64), T)
        U.append(torch.linalg.matrix_exp(1j * A).contiguous())
    src = torch.randn((v4 * ls, 4, 3, 2), generator=gen, device="cuda", dtype=torch.float32)
    return U, torch.view_as_complex(src).contiguous()


def run_native(args):
    import torch

    import gpt_b200 as g
    from gpt_b200 import cgpt

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    cgpt.init(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from gpt_b200 import parallel

    mpi = [1, 1, 1, world]  # BASELINE.json configs[2]: "1 GPU then T-split across 2/4/8 B200" (weak scaling in T)
    if world > 1:
        parallel.setup(dist, mpi)
    dims = list(DIMS)  # local extents; weak scaling: the global lattice grows with the processor grid
    gdims = [d * m for d, m in zip(dims, mpi)]
    grid = g.grid(gdims, g.single)
    # inputs exactly as /root/reference/benchmarks/dslash.py:10-49: GPT's own generator (fast engine), SU(3) links
    # exp(i 0.5 sum_a u_a T_a), complex normal source; every rank draws its block of the global lattice
    rng = g.random("benchmark", "vectorized_ranlux24_24_64")
    U = g.qcd.gauge.random(grid, rng, scale=0.5)
    qm = g.qcd.fermion.mobius(U, dict(MOBIUS))
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

    sampler = clock_sampler(local_rank)
    for _ in range(args.warmup):
        qm.Dhop.mat(dst, src)
    sync()
    time.sleep(1.0)  # let nvidia-smi start sampling
    l0 = cgpt.launch_count()
    t0 = time.time()
    cgpt.timer_start()
    for _ in range(args.steps):
        qm.Dhop.mat(dst, src)
    ms = cgpt.timer_stop()
    sync()
    t1 = time.time()
    launches = cgpt.launch_count() - l0
    clocks = sampler.stop(t0, t1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    gflops = FLOPS_PER_SITE * v5 * world / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant kernel k_dhop<float>: one launch per parity, two per step
    bytes_per_launch = (v5 // 2) * 48 * 4 + (v4 // 2) * 8 * 18 * 4
    launches_per_step = 2
    peak, peak_src = peaks()
    achieved = bytes_per_launch / (ms_per_step * 1e-3 / launches_per_step) / 1e9
    eff_bytes = (8 * 2 * 4 * 3 + 8 * 2 * 9 / LS + 2 * 4 * 3) * 4 * v5
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "dhop_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    # end to end through the public API with HOST buffers (pinned): import -> Dhop -> export
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
        sync()
        cgpt.timer_start()
        for _ in range(n_e2e):
            qm.Dhop_host(h_out, h_in)
        ms_e = cgpt.timer_stop()
        sync()
        if dist is not None:
            t = torch.tensor([ms_e], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        e2e = {"value": FLOPS_PER_SITE * v5 * world / (ms_e / n_e2e * 1e-3) / 1e9, "unit": "GFlop/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": n_e2e, "ms_per_step": ms_e / n_e2e}

    # eo-CG time-to-solve (second half of BASELINE.json's metric): eo2_ne CG on the same operator and source,
    # fixed iteration count so that the number is comparable across runs (single precision, device-side loop)
    cg_info = None
    if not args.no_cg:
        half = g.vspincolor(qm.F_grid_eo)
        g.pick_checkerboard(g.odd, half, src)
        psi = g.lattice(half)
        psi[:] = 0
        cgpt.cg_eo2_ne(qm.interface.obj, psi.obj, half.obj, 1e-30, 10)  # warm-up
        ms_cg = None
        for _ in range(2):  # best of two identical solves (the first one after a cold start has been seen 2x slower)
            psi[:] = 0
            sync()
            l1 = cgpt.launch_count()
            cgpt.timer_start()
            hist, conv = cgpt.cg_eo2_ne(qm.interface.obj, psi.obj, half.obj, 1e-30, args.cg_iterations)
            ms_try = cgpt.timer_stop()
            sync()
            ms_cg = ms_try if ms_cg is None else min(ms_cg, ms_try)
        if dist is not None:
            t = torch.tensor([ms_cg], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_cg = float(t.item())
        cg_info = {"solver": "inv.preconditioned(pc.eo2_ne(), inv.cg) on Mpc^dag Mpc, fused device loop", "iterations": len(hist),
                   "ms_total": ms_cg, "ms_per_iteration": ms_cg / max(len(hist), 1),
                   "residual_reduction": (hist[-1] / hist[0]) ** 0.5 if hist else None,
                   "launches_per_iteration": (cgpt.launch_count() - l1) / max(len(hist), 1)}
        del half, psi

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(sample_dims=[16, 16, 16, 32], seconds=12.0)

    if rank == 0:
        out = {
            "metric": "mobius_dwf_dslash_gflops", "value": gflops, "unit": "GFlop/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (g.random(\"benchmark\", ranlux24_24) links scale 0.5 + cnormal source, as benchmarks/dslash.py)",
            "config": {"workload": "Mobius DWF Dhop 32^3x64 Ls=12 single per GPU (BASELINE.json configs[2]; T-split, global T=64*n_gpus)",
                       "local_dims": dims, "Ls": LS, "cache": "inputs (2.4 GB field + 0.6 GB links) larger than L2, no flush needed",
                       "global_dims": gdims, "parallelism": "mpi " + ".".join(str(m) for m in mpi) + " (x.y.z.t), halo exchange NCCL send/recv overlapped with the interior stencil"},
            "gbs_effective_gpt_convention": eff_bytes * world / (ms_per_step * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_dhop_f32_tma (TMA-fed persistent t-sweep, packed FFMA2, one launch per parity)", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "launches_per_step": launches_per_step},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "eo_cg": cg_info,
        }
        print(json.dumps(out))
    if dist is not None:
        cgpt.comm_finalize()
        dist.destroy_process_group()


def cpu_baseline(sample_dims, seconds, steps=None, warmup=1):
    """C/OpenMP restatement of Dhop (oracle/dslash_ref.c) on the host cores, bounded sample of the workload"""
    from oracle import cref

    rs = np.random.default_rng(7)
    v4 = int(np.prod(sample_dims))
    # cheap unitary links: first-order exp is enough for a throughput sample; normalisation does not matter
    T = gell_mann_half()
    V = np.empty((4, v4, 3, 3), dtype=np.complex64)
    for mu in range(4):
        u = (rs.random((v4, 8), dtype=np.float32) - 0.5) * 0.5
        A = np.einsum("na,aij->nij", u.astype(np.complex64), T)
        V[mu] = np.eye(3, dtype=np.complex64) + 1j * A - 0.5 * (A @ A)
    psi = (rs.standard_normal((v4 * LS, 4, 3), dtype=np.float32) + 1j * rs.standard_normal((v4 * LS, 4, 3), dtype=np.float32)).astype(np.complex64)
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
            "sample": f"{n} applications of Dhop on {sample_dims} Ls={LS} single (oracle/dslash_ref.c, OpenMP; Grid unavailable)",
            "ms_per_step": dt * 1e3, "cpu": cpu_model, "nproc": os.cpu_count()}


def run_reference(args):
    """reference arm: the CPU restatement of the reference's Dhop on the host cores (Grid cannot be built here)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dims = [16, 16, 16, 32]
    c = cpu_baseline(dims, seconds=None, steps=args.steps, warmup=max(args.warmup, 1))
    out = {
        "impl": "reference", "metric": "mobius_dwf_dslash_gflops", "value": c["value"], "unit": "GFlop/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": c["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Mobius DWF Dhop 32^3x64 Ls=12 single per GPU (BASELINE.json configs[2]); each step is a bounded "
                               f"sample: one Dhop on {dims} Ls={LS}", "Ls": LS},
        "cpu_baseline": {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": c["value"], "unit": "GFlop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--cg-iterations", type=int, default=50)
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)

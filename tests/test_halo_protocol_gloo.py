"""
World-size-2 test of the halo protocol on CPU (gloo): two processes each own half of the lattice in T (or Z), run
the oracle's projection / link multiply on their local block following exactly the data flow of
gpt_b200/csrc/halo.cu (pack low face for the neighbour's forward hop and high face for its backward hop, exchange,
exterior update with ghost links) and must reproduce the single-process oracle Dhop.  This pins the decomposition
logic (which face goes where, boundary phase on the global last slice, parity with even local offsets).
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mu, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import qcd
    from oracle.rng import random as oracle_random

    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    dims = [4, 4, 4, 8]
    rng = oracle_random("halo")
    U = qcd.gauge_random(rng, dims, scale=0.8)
    phases = [1.0, -1.0, np.exp(0.3j), -1.0]
    V = qcd.apply_boundary_phases(U, phases)  # phase sits on the GLOBAL last slice
    psi = rng.cnormal(dims, (4, 3))
    ref = qcd.dhop(V, psi)

    ax = qcd.axis(mu)
    n = dims[mu] // world
    sl = [slice(None)] * 4
    sl[ax] = slice(rank * n, (rank + 1) * n)
    sl = tuple(sl)
    Vl = [v[sl] for v in V]
    pl = psi[sl]

    def take(a, idx):
        s = [slice(None)] * a.ndim
        s[ax] = idx
        return a[tuple(s)]

    g = qcd.gamma
    one = g["I"]
    # pack: low face projected with (1 - gamma_mu) (neighbour's forward hop), high face with (1 + gamma_mu)
    to_lo = qcd.spin_mul(one - g[mu], take(pl, 0))
    to_hi = qcd.spin_mul(one + g[mu], take(pl, n - 1))
    link_hi = take(Vl[mu], n - 1)  # ghost links for the rank above

    def exchange(send_lo, send_hi):
        lo, hi = (rank - 1) % world, (rank + 1) % world
        r_hi = torch.zeros_like(torch.from_numpy(np.ascontiguousarray(send_lo)))
        r_lo = torch.zeros_like(r_hi)
        ops = [dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(send_lo)), lo, tag=1),
               dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(send_hi)), hi, tag=2),
               dist.P2POp(dist.irecv, r_hi, hi, tag=1), dist.P2POp(dist.irecv, r_lo, lo, tag=2)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return r_lo.numpy(), r_hi.numpy()

    from_lo, from_hi = exchange(to_lo, to_hi)
    ghost, _ = exchange(link_hi, link_hi)

    # interior: all hops that stay local (off-rank hops of direction mu skipped)
    out = np.zeros_like(pl)
    for nu in range(4):
        fwd = qcd.link_mul(Vl[nu], qcd.shift(pl, nu, +1), False)
        bwd = qcd.shift(qcd.link_mul(qcd.adj(Vl[nu]), pl, False), nu, -1)
        if nu == mu:
            s = [slice(None)] * pl.ndim
            s[ax] = n - 1
            fwd[tuple(s)] = 0
            s[ax] = 0
            bwd[tuple(s)] = 0
        out += 0.5 * (qcd.spin_mul(g[nu] - one, fwd) - qcd.spin_mul(g[nu] + one, bwd))
    # exterior: high face gets U_mu(x) h_fwd, low face gets U_mu(x-mu)^dag h_bwd (ghost link)
    s = [slice(None)] * pl.ndim
    s[ax] = n - 1
    out[tuple(s)] += -0.5 * qcd.link_mul(take(Vl[mu], n - 1), from_hi, False)
    s[ax] = 0
    out[tuple(s)] += -0.5 * qcd.link_mul(qcd.adj(ghost), from_lo, False)
    err = float(np.linalg.norm(out - ref[sl]) / np.linalg.norm(ref[sl]))
    q.put((rank, err))
    dist.destroy_process_group()


@pytest.mark.parametrize("mu", [3, 2])
def test_halo_protocol_two_ranks(mu):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mu, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err in res:
        assert err < 1e-13, (rank, err)

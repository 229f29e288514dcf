"""
World-size-4 test (gloo, CPU) of the decomposition logic behind bench.py's multi-GPU parity check: every rank pads its block of
a global field with the facing slices of its T and Z neighbours in the processor grid 1.1.Z.T (bench.pad_with_neighbour_faces)
and must find exactly the periodic neighbourhood of its block in the global field.  Processor grids 1.1.1.4 (the SCALE run's
T-split), 1.1.2.2 (T x Z) and 1.1.4.1.
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


def _worker(rank, world, port, mpi, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import bench

    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    X, Y, Z, T = 4, 2, 4, 6  # local extents
    ncomp = 3
    pz, pt = mpi[2], mpi[3]
    gZ, gT = Z * pz, T * pt
    rs = np.random.default_rng(7)
    glob = (rs.standard_normal((gT, gZ, Y, X, ncomp)) + 1j * rs.standard_normal((gT, gZ, Y, X, ncomp))).astype(np.complex64)
    cz, ct = rank % pz, rank // pz
    block = np.ascontiguousarray(glob[ct * T:(ct + 1) * T, cz * Z:(cz + 1) * Z])

    def gather(face):
        mine = torch.from_numpy(np.ascontiguousarray(face))
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        return [e.numpy() for e in every]

    P = bench.pad_with_neighbour_faces(block.reshape(-1), [X, Y, Z, T], ncomp, mpi, rank, gather)
    P = P.reshape(T + 2, Z + 2, Y, X, ncomp)
    t_idx = [(ct * T + t) % gT for t in range(-1, T + 1)]
    z_idx = [(cz * Z + z) % gZ for z in range(-1, Z + 1)]
    want = glob[np.ix_(t_idx, z_idx)]
    want[[0, 0, -1, -1], [0, -1, 0, -1]] = 0  # corners are not exchanged
    q.put((rank, bool(np.array_equal(P, want))))
    dist.destroy_process_group()


@pytest.mark.parametrize("mpi", [[1, 1, 1, 4], [1, 1, 2, 2], [1, 1, 4, 1]])
def test_bench_padding_four_ranks(mpi):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 4, port, mpi, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res

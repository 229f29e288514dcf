"""
Multi-GPU set-up: one process per GPU (torchrun), a Cartesian processor grid over (x,y,z,t) like GPT's
`--mpi X.Y.Z.T` (lib/gpt/core/grid.py:77-94; the fifth dimension is never split, grid.py:83-87).
torch.distributed is only the plumbing that broadcasts the NCCL id; halos and global sums then run inside
libcgpt_b200 (gpt_b200/csrc/comm.cu, halo.cu).
"""
import numpy as np

from gpt_b200 import capi

mpi = [1, 1, 1, 1]
rank = 0
world = 1
active = False


def processor_coor(rank_, mpi_):
    """rank -> processor coordinates, x fastest (gpt_b200/csrc/comm.cu:cgptb_comm_init)"""
    c = []
    r = rank_
    for m in mpi_:
        c.append(r % m)
        r //= m
    return c


def default_mpi(world_):
    """T first, then Z: 2 -> 1.1.1.2, 4 -> 1.1.1.4, 8 -> 1.1.2.4"""
    m = [1, 1, 1, 1]
    w = world_
    t = 1
    while w % 2 == 0 and t < 4:
        t *= 2
        w //= 2
    m[3] = t
    m[2] = w
    return m


def local_dims(global_dims4, mpi_=None):
    m = mpi if mpi_ is None else mpi_
    out = []
    for d, p in zip(global_dims4, m):
        if d % p or (d // p) % 2:
            raise ValueError(f"global extent {d} cannot be split evenly over {p} ranks (local extents must be even)")
        out.append(d // p)
    return out


def setup(dist, mpi_=None):
    """call once per process after torch.distributed.init_process_group and cgpt.init"""
    global mpi, rank, world, active
    rank = dist.get_rank()
    world = dist.get_world_size()
    mpi = list(mpi_) if mpi_ is not None else default_mpi(world)
    assert int(np.prod(mpi)) == world
    obj = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    capi.comm_init(rank, world, mpi, obj[0])
    active = world > 1


def globalsum(x):
    if not active:
        return x
    if isinstance(x, complex):
        r = capi.comm_globalsum([x.real, x.imag])
        return complex(r[0], r[1])
    if isinstance(x, (float, int)):
        return float(capi.comm_globalsum([float(x)])[0])
    a = np.asarray(x)
    if np.iscomplexobj(a):
        r = capi.comm_globalsum(a.astype(np.complex128).view(np.float64))
        return r.view(np.complex128).reshape(a.shape)
    return capi.comm_globalsum(a.astype(np.float64)).reshape(a.shape)

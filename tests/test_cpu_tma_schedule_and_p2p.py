"""
Host-side logic of round 2 that needs no GPU:
  * the work schedule of the TMA sweep kernel (gpt_b200/csrc/dslash_tma.cu build_schedule, exported through the C ABI as
    cgptb_debug_tma_schedule): every (tile, chunk group, time slice) is covered exactly once, whatever the number of CTAs, the
    number of chunks per CTA, the schedule type or the slab of time slices; the load imbalance of the default schedule stays small;
  * the peer-to-peer halo protocol of gpt_b200/csrc/halo.cu (pack into the neighbour's receive buffer of parity n & 1, flag = n,
    wait for the own flags, read) as a thread simulation with random delays on rings of 2, 3 and 4 ranks: no buffer is overwritten
    before it has been read although there are no acknowledgements -- the exchange is symmetric and the buffers alternate.
"""
import random
import threading
import time

import numpy as np
import pytest

from gpt_b200 import capi


@pytest.mark.parametrize("dims,Ls,G", [([32, 32, 32, 64], 12, 1), ([32, 32, 32, 64], 12, 3), ([16, 8, 12, 8], 8, 2), ([8, 12, 8, 6], 4, 1),
                                       ([8, 8, 8, 16], 24, 3), ([48, 48, 24, 24], 12, 3)])
@pytest.mark.parametrize("grid", [148, 140, 7, 1])
def test_tma_schedule_covers_everything_once(dims, Ls, G, grid):
    hx, Ly, Lz, T = dims[0] // 2, dims[1], dims[2], dims[3]
    ngroup = Ls // 4 // G
    for sched, trl, t_begin, t_count in [(1, 16, 0, 0), (0, 4, 0, 0), (0, 1, 0, 0), (1, 16, 2, max(2, T // 4))]:
        items = capi.debug_tma_schedule(dims, Ls, grid, G, sched, trl, t_begin, t_count)
        t0, tc = (t_begin, t_count) if t_count > 0 else (0, T)
        cover = np.zeros((ngroup, hx // 4, Ly // 4, Lz // 4, T), dtype=np.int32)
        for c, xh0, y0, z0, it0, itrl in items:
            assert xh0 % 4 == 0 and y0 % 4 == 0 and z0 % 4 == 0 and itrl >= 1 and 0 <= c < ngroup
            assert t0 <= it0 and it0 + itrl <= t0 + tc
            cover[c, xh0 // 4, y0 // 4, z0 // 4, it0:it0 + itrl] += 1
        assert np.all(cover[..., t0:t0 + tc] == 1), (sched, trl)
        assert np.all(cover[..., :t0] == 0) and np.all(cover[..., t0 + tc:] == 0)
        if sched == 1 and t_count == 0:
            # load balance: time steps (+ 2 partial steps per item) of the busiest CTA against the mean
            work = np.zeros(grid)
            for i, (_, _, _, _, _, itrl) in enumerate(items):
                work[i % grid] += itrl + 2
            if len(items) >= grid:
                assert work.max() <= 1.25 * work.mean() + 3, (work.max(), work.mean())


def test_tma_schedule_headline_numbers():
    """what DESIGN.md quotes for 32^3 x 64 x 12 on 148 CTAs"""
    it1 = capi.debug_tma_schedule([32, 32, 32, 64], 12, 148, 1)
    it3 = capi.debug_tma_schedule([32, 32, 32, 64], 12, 148, 3)
    assert len(it1) == 5 * 148 + 28 * 5 and len(it3) == 148 + 108 * 4
    w3 = np.zeros(148)
    for i, row in enumerate(it3):
        w3[i % 148] += row[5]
    assert w3.max() == 64 + 3 * 16  # one full sweep + three ranges of 16 slices: 112 time steps (110.7 if the work were divisible)
    with pytest.raises(RuntimeError):
        capi.debug_tma_schedule([32, 32, 30, 64], 12, 148, 1)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_p2p_halo_protocol_simulation(world):
    ncalls = 60
    rng = random.Random(world)
    delays = [[rng.random() * 2e-4 for _ in range(4 * ncalls)] for _ in range(world)]
    # arena of rank r: flags[side], recv[side][parity]; side 0 = "from_lo" (written by rank r-1), 1 = "from_hi" (by rank r+1)
    flags = [[0, 0] for _ in range(world)]
    recv = [[[None, None], [None, None]] for _ in range(world)]
    errors = []
    lock = threading.Lock()

    def rank_main(r):
        lo, hi = (r - 1) % world, (r + 1) % world
        d = iter(delays[r])
        for n in range(1, ncalls + 1):
            # pack + copy: my low face is the lo neighbour's from_hi, my high face the hi neighbour's from_lo
            time.sleep(next(d))
            recv[lo][1][n & 1] = ("lo-face", r, n)
            recv[hi][0][n & 1] = ("hi-face", r, n)
            with lock:
                flags[lo][1] = n
                flags[hi][0] = n
            time.sleep(next(d))  # interior stencil
            t_end = time.time() + 20
            while True:  # cuStreamWaitValue32(flag >= n)
                with lock:
                    ok = flags[r][0] >= n and flags[r][1] >= n
                if ok:
                    break
                if time.time() > t_end:
                    errors.append((r, n, "timeout"))
                    return
                time.sleep(1e-5)
            a = recv[r][0][n & 1]
            time.sleep(next(d))  # exterior kernel reading the buffers
            b = recv[r][1][n & 1]
            a2 = recv[r][0][n & 1]
            if a != ("hi-face", lo, n) or a2 != a or b != ("lo-face", hi, n):
                errors.append((r, n, a, a2, b))
                return

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not errors, errors[:3]


def test_stencil_program_lines():
    """g.local_stencil: tuple and dictionary forms of a code line are the same program; malformed lines are refused before
    anything reaches the library (no GPU needed)"""
    from gpt_b200.local_stencil import _program_line

    t = _program_line((0, 1, 2, -1, 0.5 - 1j, [(3, 0, 1), [1, 2, 0]]))
    d = _program_line({"target": 0, "source": 1, "source_point": 2, "accumulate": -1, "weight": 0.5 - 1j, "factor": [(3, 0, 1), (1, 2, 0)]})
    assert t == d and t["factor"] == [(3, 0, 1), (1, 2, 0)]
    with pytest.raises(ValueError):
        _program_line({"target": 0, "source": 1})
    with pytest.raises(ValueError):
        _program_line((0, 1, 2, -1, 1.0))

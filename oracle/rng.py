"""
ORACLE (test infrastructure only) -- CPU restatement of GPT's parallel random number generator.

This file is a *checker*: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import it.  The product (gpt_b200/) never does.

Restates (file:line relative to /root/reference):
  * RANLUX24 with luxury p            lib/cgpt/lib/random/ranlux.h:27-104
  * SHA-256 seeding + 64 virtual lanes lib/cgpt/lib/random/vector.h:19-107
  * bit reservoir / distributions      lib/cgpt/lib/random/distribution.h:19-175
  * block decomposition + fill order   lib/cgpt/lib/random/parallel.h:21-318, engine.h:37-131

Parity pin: tests/test_oracle_rng.py checks the known-answer vectors of
/root/reference/tests/random/simple.py:18-31,69-130 (five normals on an [8,4,4,4] grid before and after
1000 further draws, the gauge.random plaquette, and the scalar `choice` sequence).

Implementation note: every block generator of a lattice consumes exactly the same number of words in the
same pattern, so all blocks advance in lock-step and the state is held as numpy arrays [24, blocks, 64].
"""

import hashlib
import math

import numpy as np

_B = 1 << 24
_MASK = _B - 1
_R = 24
_S = 10


def _sha_words(seed_u64, idx):
    # vector.h:38-58 : sha256 over the little-endian uint64 array (seed ++ [idx]); hash[] are the 8
    # big-endian digest words (checksums/sha256.h).
    data = np.array(list(seed_u64) + [idx], dtype="<u8").tobytes()
    return np.frombuffer(hashlib.sha256(data).digest(), dtype=">u4").astype(np.int64)


class _ranlux:
    """ranlux.h:27-104; state arrays carry arbitrary trailing shape (lock-step generators)."""

    def __init__(self, seed, p):
        # seed: int64 array [24, ...]
        self.x = (seed & _MASK).astype(np.int64)
        self.c = (seed[0] == 0).astype(np.int64)
        self.offset = 0
        self.discard = _R - 1
        self.p = p

    def step(self):
        self.offset = (self.offset + 1) % _R
        o = self.offset
        xs = self.x[(_S - 1 - o + _R) % _R]
        xr = self.x[(_R - 1 - o + _R) % _R]
        d = xs - xr - self.c
        self.c = (d < 0).astype(np.int64)
        v = d & _MASK  # == (d + b) & (b-1) for both signs (two's complement)
        self.x[(-1 - o + _R) % _R] = v
        return v

    def __call__(self):
        self.discard += 1
        if self.discard == _R:
            self.discard = 0
            for _ in range(self.p - _R):
                self.step()
        return self.step()


class vector_rng:
    """vector.h:19-107 for a batch of `nb` independent generators advancing in lock-step."""

    def __init__(self, seeds_u64, p):
        # seeds_u64: list (len nb) of uint64 seed vectors
        nb = len(seeds_u64)
        words = np.empty((_R, nb), dtype=np.int64)
        for i, s in enumerate(seeds_u64):
            w = np.concatenate([_sha_words(s, k) for k in range(3)])
            words[:, i] = w[:_R]
        srng = _ranlux(words, p)
        # lane L of the vector generator is seeded with scalar outputs [24 L, 24 L + 24)
        vseed = np.empty((_R, nb, 64), dtype=np.int64)
        for lane in range(64):
            for j in range(_R):
                vseed[j, :, lane] = srng()
        self.vrng = _ranlux(vseed, p)
        self.nb = nb
        self.populate()

    def populate(self):
        self.buffer = self.vrng().copy()  # [nb, 64]
        self.nbuffer = 0

    def __call__(self):
        if self.nbuffer == 64:
            self.populate()
        r = self.buffer[:, self.nbuffer]
        self.nbuffer += 1
        return r

    def words(self, n):
        """next n words of every generator, shape [nb, n]"""
        out = np.empty((self.nb, n), dtype=np.int64)
        k = 0
        while k < n:
            if self.nbuffer == 64:
                self.populate()
            m = min(64 - self.nbuffer, n - k)
            out[:, k : k + m] = self.buffer[:, self.nbuffer : self.nbuffer + m]
            self.nbuffer += m
            k += m
        return out


class random_bits:
    """distribution.h:19-110 (cgpt_random<RNG, uint64_t>) for nb lock-step generators."""

    def __init__(self, seeds_u64, p):
        self.rng = vector_rng(seeds_u64, p)
        self.nb = self.rng.nb
        self.state = np.zeros(self.nb, dtype=np.uint64)
        self.nbits = 0
        self.stack_normal = []
        self._populate()

    def _populate(self):
        w = self.rng().astype(np.uint64)
        self.state = (self.state << np.uint64(24)) + w  # uint64 wrap == state * 2^24 + w
        self.nbits = min(self.nbits + 24, 64)

    def get_bits(self, bits):
        while bits > self.nbits:
            self._populate()
        res = self.state & np.uint64((1 << bits) - 1)
        self.state = self.state >> np.uint64(bits)
        self.nbits -= bits
        return res

    def get_double(self):
        return self.get_bits(53).astype(np.float64) / float(1 << 53)

    def get_doubles(self, n):
        """n successive get_double() of every generator -> [nb, n]; same bit stream as n calls."""
        out = np.empty((self.nb, n), dtype=np.float64)
        for i in range(n):
            out[:, i] = self.get_double()
        return out

    def get_uniform_int(self, mx):
        # scalar use only (nb == 1): rejection sampling is data dependent
        assert self.nb == 1
        if mx == 0:
            return 0
        bits = int(math.floor(math.log2(mx))) + 1
        while True:
            r = int(self.get_bits(bits)[0])
            if r <= mx:
                return r

    def get_normals(self, n):
        """n successive get_normal() -> [nb, n] (Box-Muller, z1 stacked; distribution.h:83-106)"""
        out = np.empty((self.nb, n), dtype=np.float64)
        i = 0
        while i < n:
            if self.stack_normal:
                out[:, i] = self.stack_normal.pop()
                i += 1
                continue
            u1 = self.get_double()
            u2 = self.get_double()
            # rejection u1 <= DBL_MIN would de-synchronise the lock-step generators
            assert np.all(u1 > np.finfo(np.float64).tiny)
            rad = np.sqrt(-2.0 * np.log(u1))
            two_pi = 2.0 * 3.14159265358979323846
            z0 = rad * np.cos(two_pi * u2)
            z1 = rad * np.sin(two_pi * u2)
            self.stack_normal.append(z1)
            out[:, i] = z0
            i += 1
        return out


_ENGINES = {"vectorized_ranlux24_389_64": 389, "vectorized_ranlux24_24_64": 24}


class random:
    """
    Mirror of gpt.random (lib/gpt/core/random.py:23-165) restricted to what the hot path's tests need.
    Lattices are numpy arrays in "oracle layout": shape dims[::-1] + tensor shape, C-order, i.e. the flat
    site index is lexicographic with dimension 0 fastest (for 5d grids dimension 0 is s).
    """

    def __init__(self, seed, engine="vectorized_ranlux24_389_64"):
        self.seed_str = seed
        self.p = _ENGINES[engine]
        self.srng = random_bits([[ord(ch) for ch in seed]], self.p)
        self.prng = {}

    # ---- parallel.h:21-141 --------------------------------------------------------------------
    def _setup(self, dims):
        dims = tuple(int(d) for d in dims)
        nd = len(dims)
        if nd <= 4:
            blocked = [True] * nd
        elif nd == 5:
            blocked = [False, True, True, True, True]
        else:
            raise ValueError("Nd not supported")
        block = 2
        block_dim = [dims[j] // block if blocked[j] else 1 for j in range(nd)]
        reduced_dim = [block if blocked[j] else dims[j] for j in range(nd)]
        for j in range(nd):
            assert not blocked[j] or dims[j] % block == 0
        nblocks = int(np.prod(block_dim))
        # seed: string ++ fdimensions ++ gdimensions ++ [t]   (engine.h:88-98, parallel.h:86-89)
        base = [ord(ch) for ch in self.seed_str] + list(dims) + list(dims)
        seeds = []
        bcoors = np.empty((nblocks, nd), dtype=np.int64)
        for idx in range(nblocks):
            # Lexicographic::CoorFromIndex : dimension 0 fastest
            r = idx
            bc = []
            for j in range(nd):
                bc.append(r % block_dim[j])
                r //= block_dim[j]
            bcoors[idx] = bc
            t = 0
            for j in range(nd):
                if blocked[j]:
                    t = t * (dims[j] // block) + bc[j]
            seeds.append(base + [t])
        # flat lattice index (dim 0 fastest) of sample i in block idx
        nred = int(np.prod(reduced_dim))
        site = np.empty((nblocks, nred), dtype=np.int64)
        for ridx in range(nred):
            r = ridx
            rc = []
            for j in range(nd):
                rc.append(r % reduced_dim[j])
                r //= reduced_dim[j]
            flat = np.zeros(nblocks, dtype=np.int64)
            stride = 1
            for j in range(nd):
                c = bcoors[:, j] * (block if blocked[j] else 1) + rc[j]
                flat += c * stride
                stride *= dims[j]
            site[:, ridx] = flat
        return {"gen": random_bits(seeds, self.p), "site": site, "nred": nred, "nblocks": nblocks}

    def _state(self, dims, grid_tag):
        # generators are kept per grid OBJECT (engine.h:82-99): two grids of equal dims (e.g. a single and a
        # double precision one) each start from the same seed; `grid_tag` tells such grids apart
        key = (tuple(int(d) for d in dims), grid_tag)
        if key not in self.prng:
            self.prng[key] = self._setup(key[0])
        return self.prng[key]

    def _sample(self, dims, tensor_shape, dist, dtype=np.complex128, grid_tag=None, **kw):
        st = self._state(dims, grid_tag)
        nel = int(np.prod(tensor_shape)) if len(tensor_shape) else 1
        n = st["nred"] * nel
        gen = st["gen"]
        if dist == "normal":
            v = gen.get_normals(n) * kw.get("sigma", 1.0) + kw.get("mu", 0.0)
            v = v.astype(np.complex128)
        elif dist == "cnormal":
            z = gen.get_normals(2 * n) * kw.get("sigma", 1.0) + kw.get("mu", 0.0)
            v = z[:, 1::2] + 1j * z[:, 0::2]  # imaginary part drawn first (distribution.h:133-137)
        elif dist == "uniform_real":
            lo, hi = kw.get("min", 0.0), kw.get("max", 1.0)
            v = (gen.get_doubles(n) * (hi - lo) + lo).astype(np.complex128)
        else:
            raise ValueError(dist)
        nsites = int(np.prod(dims))
        out = np.empty((nsites, nel), dtype=np.complex128)
        out[st["site"].reshape(-1)] = v.reshape(-1, nel)
        out = out.reshape(tuple(dims[::-1]) + tuple(tensor_shape))
        return out.astype(dtype)  # generated in double, then cast (engine.h:104-105)

    def normal(self, dims, tensor_shape=(), **kw):
        return self._sample(dims, tensor_shape, "normal", **kw)

    def cnormal(self, dims, tensor_shape=(), **kw):
        return self._sample(dims, tensor_shape, "cnormal", **kw)

    def uniform_real(self, dims, tensor_shape=(), **kw):
        return self._sample(dims, tensor_shape, "uniform_real", **kw)

    # ---- scalar interface (engine.h:324-327) ------------------------------------------------------
    def scalar_normal(self):
        return float(self.srng.get_normals(1)[0, 0])

    def scalar_uniform_int(self, lo, hi):
        return self.srng.get_uniform_int(hi - lo) + lo

    def scalar_zn(self, n=2):
        # distribution.h:112-114 ; only the consumed bits matter for the KAT
        k = self.srng.get_uniform_int(n - 1)
        return np.exp(1j * 2.0 * np.pi * k / n)

    def choice(self, array, n):
        return [array[self.scalar_uniform_int(0, len(array) - 1)] for _ in range(n)]

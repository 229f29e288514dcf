// Single-precision hopping term as a persistent, TMA-fed sweep along t ("3.5D blocking") for sm_100a.
//
// Why: the register/L1-based kernels in dslash_f32.cu re-fetch every neighbour spinor through L1/L2 (about 4.8 of the 8
// neighbour reads miss L1), spend ~45 % of their instructions on neighbour index arithmetic and can keep only one or two
// hops of loads in flight per thread.  Here the loads are issued by one producer warp with cp.async.bulk.tensor (TMA)
// into a shared-memory ring, completely decoupled from the arithmetic:
//
//   * work item = (4x4x4 tile of checkerboard sites in (x/2, y, z)) x (group of G chunks of SC = 4 fifth-dimension slices) x
//     (range of time slices); a persistent CTA (one per SM) walks its items and sweeps each along t.  A time step consists of
//     G sub-steps, one per chunk, that share the tile's links: with G = 3 at Ls = 12 a CTA owns ALL fifth-dimension slices of
//     its tile, so the 8 links of a 4d site come from L2 once instead of three times (4.8 -> 3.5 GB of L2->SM traffic per
//     launch), 148 instead of 49 tiles are swept concurrently (the z faces between two "rounds" of CTAs are what misses in L2:
//     5.2 -> 1.7 rounds), and everything a fifth-dimension epilogue needs is in one CTA;
//   * per time slice the producer loads, for each of the three 32-byte component planes, the centre box of the tile
//     (64 sites, as bottom / middle / top z layers) and its six (x/2, y, z) faces (16 sites each) -- 27 box loads with the
//     128-byte swizzle, each shared memory row being the 4 x 32 B of one site -- plus the tile's 64 x 8 links (two boxes
//     of four links each): 2.5 spinor loads per output site instead of 8, all address arithmetic done by the TMA unit;
//   * only two stages (2 x 98 KB) fit into shared memory, so each stage is handed back in two halves (after the fourth
//     and after the last hop of a step) and refilled in two halves: every byte is requested 1.5 steps ahead of its use;
//   * the +-t neighbours never travel twice: the thread that owns (site, s) reads its own entry of the centre box once
//     per slice, uses it for the forward-t hop of the output one slice back (kept open in registers with its U_t) and
//     carries its (1 +- gamma_t) projection (a half spinor) to the backward-t hop of the next slice;
//   * 256 compute threads = 64 sites x 4 s, each with the open accumulator + carried half spinor of its G chunks in
//     registers; a neighbour spinor is six conflict-free LDS.128 at (slot base + per-thread constant + immediate), the
//     site's links are broadcast LDS.128; arithmetic is the packed FFMA2 code of packed.cuh.
//
// Reference semantics: Grid's DhopEO/DhopOE behind cgpt opcodes 3002/4002 (lib/cgpt/lib/operators/register.h:2-20),
// restated in lib/gpt/qcd/fermion/reference/wilson_clover.py:182-200.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <utility>
#include <vector>
#include "dslash.cuh"
#include "operator.cuh"
#include "packed.cuh"

namespace cgptb {

namespace tma {
constexpr int TX = 4, TY = 4, TZ = 4;  // tile in checkerboard coordinates (x/2, y, z)
constexpr int SC = 4;                  // fifth-dimension slices per chunk (4 x 32 B = one 128-byte swizzle row)
constexpr int NS = TX * TY * TZ;       // 64 sites per time slice
constexpr int NCOMP = NS * SC;         // compute threads
constexpr int NTHREADS = NCOMP + 128;  // + producer warp group (one warp works; setmaxnreg hands its registers to the consumers)
constexpr int REGS_PRODUCER = 72, REGS_CONSUMER = 216;  // 128 x 72 + 256 x 216 = 64512 <= 65536
constexpr int ROW_B = 128;
// shared memory rings.  A sub-step (one chunk of one time slice) needs: the centre box of the tile (64 rows of 128 B per
// 32-byte component plane), the faces of its first half (A: x-, x+, y+, 16 rows each) and of its second half (B: y-, z+, z-); a
// time step needs the tile's links in two halves of four, in the order a step uses them: A = {t-, x+, x-, y+}, B = {y-, z+,
// z-, t+}.  Each class has its own ring with its own full / empty mbarriers, so a buffer is refilled as soon as ITS last reader
// is done: centre boxes three deep (requested 2 sub-steps before use), A faces two deep (1.5), B faces three deep (2.5: the z
// faces are what misses in L2), link halves three deep (G >= 2; a half lives for G sub-steps) or four deep (G = 1).
// Every box carries its three component planes (fifth tensor dimension), so a sub-step is 7 TMA instructions: the TMA unit
// spends ~100 cycles per instruction whatever the box size, which bounded the memory side at 28 instructions per sub-step.
constexpr int CENTER1_B = NS * ROW_B;           // 8192: one plane of a centre box, rows [z][y][x]
constexpr int FACE1_B = (NS / TX) * ROW_B;      // 2048: one plane of a face
constexpr int CENTER_B = 3 * CENTER1_B;         // 24576: a centre box [plane][z][y][x]
constexpr int FACE_B = 3 * FACE1_B;             // 6144: a face [plane][16 rows]
constexpr int FACES_B = 3 * FACE_B;             // 18432: a face set (three faces)
// a link row is 4 x 72 B + 16 B padding (bank shift of 12 words per site: the 8 sites a warp reads are conflict free)
constexpr int LINK_ROW_F = 76;
constexpr int LINK_ROW_B = LINK_ROW_F * 4;    // 304
constexpr int LINK_HALF_B = NS * LINK_ROW_B;  // 19456
// two-row compression: a row is 4 x (rows 0 and 1 of the link: 48 B) + 4 x 8 B of U(1) factors = 224 B (bank shift 24 words)
constexpr int LINKC_ROW_F = 56;
constexpr int LINKC_ROW_B = LINKC_ROW_F * 4;    // 224
constexpr int LINKC_HALF_B = NS * LINKC_ROW_B;  // 14336 (the buffers keep the size of the uncompressed halves)
template <int G>
struct Ring {
  static constexpr int NC = 3, NFA = 2, NFB = G == 1 ? 2 : 3, NLH = G == 1 ? 4 : 3;
  static constexpr int OFF_C = 0, OFF_FA = NC * CENTER_B, OFF_FB = OFF_FA + NFA * FACES_B;
  static constexpr int OFF_LINKS = OFF_FB + NFB * FACES_B;
  static constexpr int OFF_BARS = OFF_LINKS + NLH * LINK_HALF_B;  // G = 1: 225280, else 224256
  // barriers: full then empty of each class
  static constexpr int B_FC = 0, B_EC = NC, B_FA = 2 * NC, B_EA = B_FA + NFA, B_FB = B_EA + NFA, B_EB = B_FB + NFB, B_FL = B_EB + NFB,
                       B_EL = B_FL + NLH, NBARS = B_EL + NLH;
  static constexpr int SMEM_B = OFF_BARS + NBARS * 8 + 1024;  // + alignment slack
};
constexpr int GMAX = 3;  // chunks per CTA (register budget: G x 36 registers of open state per thread)

// One unit of work: a tile x a group of G chunks of the fifth dimension x a range of time slices.  The host lays the items out
// so that item i runs on CTA i % gridDim.x ("rounds" of gridDim.x concurrent items that walk t in lock-step, see build_schedule).
struct __align__(16) ItemDesc {
  int c, xh0, y0, z0;   // group index (chunks c*G .. c*G + G - 1), tile origin
  int t0, trl;          // first output time slice, number of output time slices
  int pad0, pad1;
};

struct Geo {
  int hx, Ly, Lz, T;
  int nbx, nby, nbz, ngroup;
  int nitems;
  int p_out;
  int ls;
  int comm_mask;  // split directions (bit 2: z, bit 3: t): hops that leave the local volume are left to the halo kernels
  int skip;  // debug (CGPTB_TMA_SKIP): bit 0 x faces, 1 y faces, 2 z faces, 3 links are not loaded (traffic attribution)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait for the phase with the given parity; a wait that never completes (a lost transaction count) traps instead of
// hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// spinor of one (site, s) from a box in shared memory: a = byte address of the low 16-byte chunk in plane 0 (swizzle
// applied), ps = distance of the component planes (centre box 8192, face 2048)
__device__ __forceinline__ void lds_spinor(uint32_t a, uint32_t ps, c32 (&p)[12]) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const uint32_t ak = a + k * ps;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(p[4 * k]), "=l"(p[4 * k + 1]) : "r"(ak));
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(p[4 * k + 2]), "=l"(p[4 * k + 3]) : "r"(ak ^ 16u));
  }
}

// link at position P (0..3) of a half row (row = shared address of the site's 304-byte row): 72 B each, 16-byte aligned
// for even P, 8 mod 16 for odd P
template <int P>
__device__ __forceinline__ void lds_link(uint32_t row, float (&wr)[9], float (&wi)[9]) {
  const uint32_t a = row + P * 72;
  float v[18];
  if (P % 2 == 0) {
#pragma unroll
    for (int k = 0; k < 4; k++)
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4 * k]), "=f"(v[4 * k + 1]), "=f"(v[4 * k + 2]), "=f"(v[4 * k + 3]) : "r"(a + 16 * k));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[16]), "=f"(v[17]) : "r"(a + 64));
  } else {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(a));
#pragma unroll
    for (int k = 0; k < 4; k++)
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[2 + 4 * k]), "=f"(v[3 + 4 * k]), "=f"(v[4 + 4 * k]), "=f"(v[5 + 4 * k]) : "r"(a + 8 + 16 * k));
  }
#pragma unroll
  for (int k = 0; k < 9; k++) {
    wr[k] = v[2 * k];
    wi[k] = v[2 * k + 1];
  }
}

// the same from a compressed row: rows 0 and 1 (three LDS.128) and the U(1) factor f (LDS.64); row 2 = f conj(row 0 x row 1)
// in packed arithmetic, 18 FFMA2 / FMUL2 per link
template <int P>
__device__ __forceinline__ void lds_link_c(uint32_t row, float (&wr)[9], float (&wi)[9]) {
  const uint32_t a = row + P * 48;
  c32 u[6], f;
#pragma unroll
  for (int k = 0; k < 3; k++) asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(u[2 * k]), "=l"(u[2 * k + 1]) : "r"(a + 16 * k));
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(f) : "r"(row + 192 + P * 8));
#pragma unroll
  for (int k = 0; k < 6; k++) upk(u[k], wr[k], wi[k]);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const int i = (c + 1) % 3, j = (c + 2) % 3;
    // d = u0[i] u1[j] - u0[j] u1[i]
    c32 d = mul2(u[3 + j], bcast(wr[i]));
    d = fma2(times_iph<1>(u[3 + j]), bcast(wi[i]), d);
    d = fma2(u[3 + i], bcast(-wr[j]), d);
    d = fma2(times_iph<1>(u[3 + i]), bcast(-wi[j]), d);
    // f conj(d) = f d.re + (-i f) d.im
    float dr, di;
    upk(d, dr, di);
    c32 r = mul2(f, bcast(dr));
    r = fma2(times_iph<3>(f), bcast(di), r);
    upk(r, wr[6 + c], wi[6 + c]);
  }
}
template <int P, bool CMP>
__device__ __forceinline__ void lds_link_any(uint32_t row, float (&wr)[9], float (&wi)[9]) {
  if (CMP)
    lds_link_c<P>(row, wr, wi);
  else
    lds_link<P>(row, wr, wi);
}

// acc += recon( W(^dag) proj psi ), same arithmetic as hop_core of dslash_f32.cu, in two pieces: the spin projection (the
// backward-t hop carries the projected half spinor from one time slice to the next) and SU(3) x half spinor + reconstruction
template <int MU, bool FWD, bool DAG>
__device__ __forceinline__ void hop_proj(c32 (&h)[6], const c32 (&psi)[12]) {
  const int SGN = (FWD != DAG) ? -1 : +1;
  typedef Proj<MU, SGN> P;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    h[c] = add2(psi[c], times_iph<P::A>(psi[P::J0 * 3 + c]));
    h[3 + c] = add2(psi[3 + c], times_iph<P::B>(psi[P::J1 * 3 + c]));
  }
}
// INIT: acc = ... instead of acc += ... (first hop of a site: saves clearing the accumulator)
template <int MU, bool FWD, bool DAG, bool INIT = false>
__device__ __forceinline__ void hop_mulrecon(c32 (&acc)[12], const c32 (&h)[6], const float (&wr)[9], const float (&wi)[9]) {
  const int SGN = (FWD != DAG) ? -1 : +1;
  typedef Proj<MU, SGN> P;
  c32 chi[6];
#pragma unroll
  for (int sp = 0; sp < 2; sp++) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
      c32 a;
      if (FWD) {
        a = cmul<false>(wr[r * 3 + 0], wi[r * 3 + 0], h[sp * 3 + 0]);
        a = cmac<false>(a, wr[r * 3 + 1], wi[r * 3 + 1], h[sp * 3 + 1]);
        a = cmac<false>(a, wr[r * 3 + 2], wi[r * 3 + 2], h[sp * 3 + 2]);
      } else {
        a = cmul<true>(wr[0 * 3 + r], wi[0 * 3 + r], h[sp * 3 + 0]);
        a = cmac<true>(a, wr[1 * 3 + r], wi[1 * 3 + r], h[sp * 3 + 1]);
        a = cmac<true>(a, wr[2 * 3 + r], wi[2 * 3 + r], h[sp * 3 + 2]);
      }
      chi[sp * 3 + r] = a;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    acc[c] = INIT ? chi[c] : add2(acc[c], chi[c]);
    acc[3 + c] = INIT ? chi[3 + c] : add2(acc[3 + c], chi[3 + c]);
    acc[6 + c] = add2(INIT ? 0ull : acc[6 + c], times_iph<P::C2>(chi[P::K2 * 3 + c]));
    acc[9 + c] = add2(INIT ? 0ull : acc[9 + c], times_iph<P::C3>(chi[P::K3 * 3 + c]));
  }
}
template <int MU, bool FWD, bool DAG>
__device__ __forceinline__ void hop_math(c32 (&acc)[12], const c32 (&psi)[12], const float (&wr)[9], const float (&wi)[9]) {
  c32 h[6];
  hop_proj<MU, FWD, DAG>(h, psi);
  hop_mulrecon<MU, FWD, DAG>(acc, h, wr, wi);
}

template <int MU, bool FWD, bool DAG, int D, bool CMP>
__device__ __forceinline__ void hop_smem(c32 (&acc)[12], uint32_t spinor_addr, uint32_t plane_stride, uint32_t link_row) {
  c32 psi[12];
  float wr[9], wi[9];
  lds_spinor(spinor_addr, plane_stride, psi);
  lds_link_any<D, CMP>(link_row, wr, wi);
  hop_math<MU, FWD, DAG>(acc, psi, wr, wi);
}

// Position in a CTA's sequence of sub-steps: item (stride gridDim.x through the item table), step st of the item (0 and
// trl + 1 are the two partial steps that only feed the t hops), chunk c of the group.  The producer keeps one cursor per
// ring because the rings run different numbers of sub-steps ahead of the consumers.
struct Cursor {
  ItemDesc it;
  int item, st, c;
  __device__ __forceinline__ void start(const ItemDesc* items, int nitems) {
    item = blockIdx.x;
    st = 0;
    c = 0;
    if (item < nitems) it = items[item];
  }
  __device__ __forceinline__ bool valid(int nitems) const { return item < nitems; }
  __device__ __forceinline__ bool full() const { return st >= 1 && st <= it.trl; }
  __device__ __forceinline__ int tau(int T) const {
    int t = it.t0 - 1 + st;
    if (t < 0) t += T;
    if (t >= T) t -= T;
    return t;
  }
  template <int G>
  __device__ __forceinline__ void next(const ItemDesc* items, int nitems) {
    if (++c < G) return;
    c = 0;
    if (++st <= it.trl + 1) return;
    st = 0;
    item += gridDim.x;
    if (item < nitems) it = items[item];
  }
  // the same over full time steps only (the steps that have links)
  __device__ __forceinline__ void start_steps(const ItemDesc* items, int nitems) {
    item = blockIdx.x;
    st = 1;
    c = 0;
    if (item < nitems) it = items[item];
  }
  __device__ __forceinline__ void next_step(const ItemDesc* items, int nitems) {
    if (++st <= it.trl) return;
    st = 1;
    item += gridDim.x;
    if (item < nitems) it = items[item];
  }
};

// ABL (ablation, CGPTB_ABLATE): 0 production; 1 compute only (no TMA loads, the rings are signalled empty-handed);
// 2 memory only (all loads and stores, no hop arithmetic)
// COMM: the lattice is split across GPUs in z and/or t (Geo::comm_mask); off-rank hops are skipped here and added by
// k_exterior (halo.cu) once the faces have arrived
// G: chunks of the fifth dimension per CTA (sub-steps per time step)
// CMP: two-row link compression (the link table holds rows of LINKC_ROW_B bytes)
template <bool DAG, int ABL, bool COMM, int G, bool CMP>
__global__ void __launch_bounds__(NTHREADS, 1)
    k_dhop_f32_tma(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                   const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmC,
                   const __grid_constant__ CUtensorMap tmL, const Geo geo, const ItemDesc* __restrict__ items,
                   float* __restrict__ out, size_t out_stride) {
  typedef Ring<G> R;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_bar = sbase + R::OFF_BARS;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  auto bar = [&](int which, uint32_t i) -> uint32_t { return s_bar + 8u * ((uint32_t)which + i); };

  if (tid == 0) {
    for (int i = 0; i < R::NBARS; i++) {
      const bool full = i < R::B_EC || (i >= R::B_FA && i < R::B_EA) || (i >= R::B_FB && i < R::B_EB) || (i >= R::B_FL && i < R::B_EL);
      mbar_init(s_bar + 8 * i, full ? 1 : NCOMP / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= NCOMP / 32) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
    // ---------------- producers: one warp (one elected lane) per ring ----------------------------------------
    // warp 0: A faces, 1: centre boxes, 2: B faces, 3: link halves.  Each walks the CTA's sequence of sub-steps (time steps)
    // on its own: wait until the consumers have released the buffer that comes next in its ring, refill it.  (L2
    // eviction-priority hints on the z faces were measured in rounds 1 and 2: no effect on the DRAM traffic; removed.)
    const int role = warp - NCOMP / 32;
    if (lane != 0) return;
    Cursor cur;
    if (role == 3) {
      const bool load_links = ABL != 1 && !(geo.skip & 8);
      cur.start_steps(items, geo.nitems);
      for (uint32_t n = 0; cur.valid(geo.nitems); n++) {
        const uint32_t i = n % R::NLH, par = (n / R::NLH) & 1u;
        mbar_wait(bar(R::B_EL, i), par ^ 1u);
        if (load_links) {
          const int tau = cur.tau(geo.T);
          mbar_expect_tx(bar(R::B_FL, i), (uint32_t)(CMP ? LINKC_HALF_B : LINK_HALF_B));
          tma_load_5d(sbase + R::OFF_LINKS + i * LINK_HALF_B, &tmL, bar(R::B_FL, i), 0, cur.it.xh0, cur.it.y0, cur.it.z0,
                      (n & 1u) ? geo.T + tau : tau);
        } else {
          mbar_arrive(bar(R::B_FL, i));
        }
        if (n & 1u) cur.next_step(items, geo.nitems);
      }
      return;
    }
    const uint32_t bytesA = ((geo.skip & 1) ? 0 : 2 * FACE_B) + ((geo.skip & 2) ? 0 : FACE_B);
    const uint32_t bytesB = ((geo.skip & 4) ? 0 : 2 * FACE_B) + ((geo.skip & 2) ? 0 : FACE_B);
    const uint32_t depth = role == 0 ? R::NFA : role == 1 ? R::NC : R::NFB;
    const int b_full = role == 0 ? R::B_FA : role == 1 ? R::B_FC : R::B_FB, b_empty = role == 0 ? R::B_EA : role == 1 ? R::B_EC : R::B_EB;
    const uint32_t bytes = role == 0 ? bytesA : role == 1 ? (uint32_t)CENTER_B : bytesB;
    cur.start(items, geo.nitems);
    uint32_t i = 0, par = 0;
    for (; cur.valid(geo.nitems); cur.next<G>(items, geo.nitems)) {
      const uint32_t fb = bar(b_full, i);
      mbar_wait(bar(b_empty, i), par ^ 1u);
      if (ABL == 1 || bytes == 0 || (role != 1 && !cur.full())) {
        mbar_arrive(fb);
      } else {
        mbar_expect_tx(fb, bytes);
        const ItemDesc& it = cur.it;
        const int s0f = (it.c * G + cur.c) * SC * 8, zt = geo.Lz * cur.tau(geo.T);  // fourth coordinate = z + Lz * t
        if (role == 0) {
          const int xm = (it.xh0 == 0 ? geo.hx : it.xh0) - 1, xp = it.xh0 + TX == geo.hx ? 0 : it.xh0 + TX;
          const int yp = it.y0 + TY == geo.Ly ? 0 : it.y0 + TY;
          const uint32_t d = sbase + R::OFF_FA + i * FACES_B;
          if (!(geo.skip & 1)) {
            tma_load_5d(d, &tmX, fb, s0f, xm, it.y0, zt + it.z0, 0);
            tma_load_5d(d + FACE_B, &tmX, fb, s0f, xp, it.y0, zt + it.z0, 0);
          }
          if (!(geo.skip & 2)) tma_load_5d(d + 2 * FACE_B, &tmY, fb, s0f, it.xh0, yp, zt + it.z0, 0);
        } else if (role == 1) {
          tma_load_5d(sbase + R::OFF_C + i * CENTER_B, &tmC, fb, s0f, it.xh0, it.y0, zt + it.z0, 0);
        } else {
          const int ym = (it.y0 == 0 ? geo.Ly : it.y0) - 1;
          const int zm = (it.z0 == 0 ? geo.Lz : it.z0) - 1, zp = it.z0 + TZ == geo.Lz ? 0 : it.z0 + TZ;
          const uint32_t d = sbase + R::OFF_FB + i * FACES_B;
          if (!(geo.skip & 2)) tma_load_5d(d, &tmY, fb, s0f, it.xh0, ym, zt + it.z0, 0);
          if (!(geo.skip & 4)) {
            tma_load_5d(d + FACE_B, &tmZ, fb, s0f, it.xh0, it.y0, zt + zp, 0);
            tma_load_5d(d + 2 * FACE_B, &tmZ, fb, s0f, it.xh0, it.y0, zt + zm, 0);
          }
        }
      }
      if (++i == depth) {
        i = 0;
        par ^= 1u;
      }
    }
    return;
  }

  // ---------------- consumers: thread = (site l of the tile, slice j of each of the G chunks) -------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_CONSUMER));
  // a warp owns two x rows of the tile, (y, z) and (y + 1, z ^ 1): the parity of y + z, and with it which of the two x
  // neighbours of a checkerboard site is the site's own x/2 index, is then uniform in the warp, so that hop can take the
  // thread's own spinor from registers instead of reading it from shared memory a second time
  const int j = tid & 3, lx = (tid >> 2) & 3, row = (tid >> 4) & 1;
  const int ly = 2 * (warp & 1) + row, lz = (warp >> 1) ^ row;
  const int l = lx + TX * (ly + TY * lz);
  // offset of this thread's 16-byte chunk in row `row` of a buffer (128-byte swizzle; buffers are 1 KB aligned)
  auto rowoff = [&](int row) -> uint32_t { return (uint32_t)(row * ROW_B + (((2 * j) ^ (row & 7)) << 4)); };
  // neighbours: inside the centre box (planes 8192 B apart), or a row of a face (A: [x-][x+][y+], B: [y-][z+][z-]; 16 rows per
  // plane, planes 2048 B apart)
  const uint32_t o_own = rowoff(l);
  const bool f_left = lx == 0, f_right = lx == TX - 1, f_yp = ly == TY - 1, f_ym = ly == 0, f_zp = lz == TZ - 1, f_zm = lz == 0;
  const uint32_t o_left = f_left ? rowoff(lz * TY + ly) : rowoff(l - 1);
  const uint32_t o_right = f_right ? FACE_B + rowoff(lz * TY + ly) : rowoff(l + 1);
  const uint32_t o_yp = f_yp ? 2 * FACE_B + rowoff(lz * TX + lx) : rowoff(l + TX);
  const uint32_t o_ym = f_ym ? rowoff(lz * TX + lx) : rowoff(l - TX);
  const uint32_t o_zp = f_zp ? FACE_B + rowoff(ly * TX + lx) : rowoff(l + TX * TY);
  const uint32_t o_zm = f_zm ? 2 * FACE_B + rowoff(ly * TX + lx) : rowoff(l - TX * TY);
  auto pstride = [](bool face) -> uint32_t { return face ? (uint32_t)FACE1_B : (uint32_t)CENTER1_B; };
  const int b0 = (ly + lz + geo.p_out) & 1;  // tile origins are even
  const int slice_sites = geo.hx * geo.Ly * geo.Lz;

  // open state per chunk: the accumulator of the previous slice's output (waits for its forward-t hop) and the
  // (1 +- gamma_t) projection of the previous slice's own spinor (for the backward-t hop of this slice's output)
  c32 acc[G][12], hc[G][6];
  float utr[9], uti[9];
#pragma unroll
  for (int c = 0; c < G; c++) {
#pragma unroll
    for (int k = 0; k < 12; k++) acc[c][k] = 0ull;
#pragma unroll
    for (int k = 0; k < 6; k++) hc[c][k] = 0ull;
  }
#pragma unroll
  for (int k = 0; k < 9; k++) utr[k] = uti[k] = 0.f;

  uint32_t q = 0;  // full time steps so far
  uint32_t iC = 0, pC = 0, iA = 0, pA = 0, iB = 0, pB = 0;  // ring positions and phase parities of the current sub-step
  for (int item = blockIdx.x; item < geo.nitems; item += gridDim.x) {
    const ItemDesc it = items[item];
    const int site0 = (it.xh0 + lx) + geo.hx * ((it.y0 + ly) + geo.Ly * (it.z0 + lz));
    int tau_prev = 0;
    for (int st = 0; st <= it.trl + 1; st++) {
      int tau = it.t0 - 1 + st;
      if (tau < 0) tau += geo.T;
      if (tau >= geo.T) tau -= geo.T;
      const bool full_step = st >= 1 && st <= it.trl;
      const uint32_t hA = 2 * q, hB = 2 * q + 1;
      const uint32_t iLA = hA % R::NLH, pLA = (hA / R::NLH) & 1u, iLB = hB % R::NLH, pLB = (hB / R::NLH) & 1u;
      const uint32_t lrowA = sbase + R::OFF_LINKS + iLA * LINK_HALF_B + l * (CMP ? LINKC_ROW_B : LINK_ROW_B);
      const uint32_t lrowB = sbase + R::OFF_LINKS + iLB * LINK_HALF_B + l * (CMP ? LINKC_ROW_B : LINK_ROW_B);
      const int b = (b0 + tau) & 1;
#pragma unroll
      for (int c = 0; c < G; c++) {
        const int s = (it.c * G + c) * SC + j;
        const uint32_t cb = sbase + R::OFF_C + iC * CENTER_B, fa = sbase + R::OFF_FA + iA * FACES_B, fb = sbase + R::OFF_FB + iB * FACES_B;
        mbar_wait(bar(R::B_FC, iC), pC);
        c32 own[12];
        lds_spinor(cb + o_own, CENTER1_B, own);
        if (st >= 2) {
          // forward-t hop closes the output of the previous slice
          if (ABL != 2 && !(COMM && (geo.comm_mask & 8) && tau_prev == geo.T - 1)) hop_math<3, true, DAG>(acc[c], own, utr, uti);
          const size_t site = ((size_t)site0 + (size_t)slice_sites * tau_prev) * geo.ls + s;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; e++) upk(acc[c][4 * k + e], v[2 * e], v[2 * e + 1]);
            st256_cs(out + (k * out_stride + site) * 8, v);
          }
        }
        // ---- first half: t-, x+, x-, y+
        mbar_wait(bar(R::B_FA, iA), pA);
        if (c == 0 && full_step) mbar_wait(bar(R::B_FL, iLA), pLA);
        if (ABL == 2) {
#pragma unroll
          for (int k = 0; k < 12; k++) acc[c][k] = own[k];
        } else if (full_step) {
          if (COMM && (geo.comm_mask & 8) && tau == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++) acc[c][k] = 0ull;
          } else {
            float wr[9], wi[9];
            lds_link_any<0, CMP>(lrowA, wr, wi);
            hop_mulrecon<3, false, DAG, true>(acc[c], hc[c], wr, wi);
          }
        }
        hop_proj<3, false, DAG>(hc[c], own);
        if (ABL != 2 && full_step) {
          // b (uniform in the warp): the forward x neighbour has x/2 + 1 and the backward one the site's own x/2, or the
          // forward one the own x/2 and the backward one x/2 - 1.  The hop to the own x/2 takes the spinor from registers
          // (G = 1; with more chunks the second copy of two hops costs more in instruction fetch than the 6 LDS.128 save).
          if (G == 1) {
            float wr[9], wi[9];
            if (b) {
              lds_link_any<2, CMP>(lrowA, wr, wi);
              hop_math<0, false, DAG>(acc[c], own, wr, wi);
              hop_smem<0, true, DAG, 1, CMP>(acc[c], (f_right ? fa : cb) + o_right, pstride(f_right), lrowA);
            } else {
              lds_link_any<1, CMP>(lrowA, wr, wi);
              hop_math<0, true, DAG>(acc[c], own, wr, wi);
              hop_smem<0, false, DAG, 2, CMP>(acc[c], (f_left ? fa : cb) + o_left, pstride(f_left), lrowA);
            }
          } else {
            const uint32_t a_right = (f_right ? fa : cb) + o_right, a_left = (f_left ? fa : cb) + o_left;
            hop_smem<0, true, DAG, 1, CMP>(acc[c], b ? a_right : cb + o_own, b ? pstride(f_right) : (uint32_t)CENTER1_B, lrowA);
            hop_smem<0, false, DAG, 2, CMP>(acc[c], b ? cb + o_own : a_left, b ? (uint32_t)CENTER1_B : pstride(f_left), lrowA);
          }
          hop_smem<1, true, DAG, 3, CMP>(acc[c], (f_yp ? fa : cb) + o_yp, pstride(f_yp), lrowA);
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(R::B_EA, iA));
          if (c == G - 1 && full_step) mbar_arrive(bar(R::B_EL, iLA));
        }
        // ---- second half: y-, z+, z-, and U_t for the forward-t hops of the next step
        mbar_wait(bar(R::B_FB, iB), pB);
        if (c == 0 && full_step) mbar_wait(bar(R::B_FL, iLB), pLB);
        if (ABL != 2 && full_step) {
          hop_smem<1, false, DAG, 0, CMP>(acc[c], (f_ym ? fb : cb) + o_ym, pstride(f_ym), lrowB);
          // (z split: the two x rows of a warp have different z, the boundary predicates are per row)
          if (!(COMM && (geo.comm_mask & 4) && it.z0 + lz == geo.Lz - 1)) hop_smem<2, true, DAG, 1, CMP>(acc[c], (f_zp ? fb : cb) + o_zp, pstride(f_zp), lrowB);
          if (!(COMM && (geo.comm_mask & 4) && it.z0 + lz == 0)) hop_smem<2, false, DAG, 2, CMP>(acc[c], (f_zm ? fb : cb) + o_zm, pstride(f_zm), lrowB);
          if (c == G - 1) lds_link_any<3, CMP>(lrowB, utr, uti);
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(R::B_EC, iC));
          mbar_arrive(bar(R::B_EB, iB));
          if (c == G - 1 && full_step) mbar_arrive(bar(R::B_EL, iLB));
        }
        if (++iC == R::NC) {
          iC = 0;
          pC ^= 1u;
        }
        if (++iA == R::NFA) {
          iA = 0;
          pA ^= 1u;
        }
        if (++iB == R::NFB) {
          iB = 0;
          pB ^= 1u;
        }
      }
      if (full_step) q++;
      tau_prev = tau;
    }
  }
}

// links [site][8][9] complex -> [half][site][38] complex: half A = {t-, x+, x-, y+} = entries {7, 0, 4, 1}, half B = {y-, z+, z-,
// t+} = entries {5, 2, 6, 3}, two complex of padding per row
__global__ void k_pad_links(size_t n4, const float2* __restrict__ links, float2* __restrict__ padded) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4 * 2 * 38) return;
  const int k = (int)(i % 38);
  const size_t r = i / 38;
  const size_t site = r % n4;
  const int half = (int)(r / n4);
  float2 v = make_float2(0.f, 0.f);
  if (k < 36) {
    const int pos = k / 9, e = k - 9 * pos;
    const int d = half == 0 ? (pos == 0 ? 7 : pos == 1 ? 0 : pos == 2 ? 4 : 1) : (pos == 0 ? 5 : pos == 1 ? 2 : pos == 2 ? 6 : 3);
    v = links[(site * 8 + d) * 9 + e];
  }
  padded[i] = v;
}

// compressed links [site][8][14 reals] -> [half][site][4 x 12 reals, 4 x f]: same order of the directions as k_pad_links
__global__ void k_pad_links_c(size_t n4, const float* __restrict__ links_c, float* __restrict__ padded) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4 * 2 * LINKC_ROW_F) return;
  const int k = (int)(i % LINKC_ROW_F);
  const size_t r = i / LINKC_ROW_F;
  const size_t site = r % n4;
  const int half = (int)(r / n4);
  const int pos = k < 48 ? k / 12 : (k - 48) / 2, e = k < 48 ? k - 12 * pos : 12 + (k - 48) % 2;
  const int d = half == 0 ? (pos == 0 ? 7 : pos == 1 ? 0 : pos == 2 ? 4 : 1) : (pos == 0 ? 5 : pos == 1 ? 2 : pos == 2 ? 6 : 3);
  padded[i] = links_c[(site * 8 + d) * 14 + e];
}

static PFN_cuTensorMapEncodeTiled_v12000 encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = 0;
  if (!fn) {
    void* p = 0;
    cudaDriverEntryPointQueryResult q;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) CGPTB_ERR("cuTensorMapEncodeTiled is not available from this driver");
    fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

static void encode5(CUtensorMap* m, const void* base, const cuuint64_t (&dims)[5], const cuuint64_t (&strides)[4],
                    const cuuint32_t (&box)[5], CUtensorMapSwizzle sw) {
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = encoder()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) CGPTB_ERR("cuTensorMapEncodeTiled failed with code %d", (int)r);
}

static int env_i(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Work schedule of one launch (item i runs on CTA i % grid; the CTAs of a "round" walk t in lock-step, so the x / y / z
// neighbours inside a round's set of tiles hit in L2).
//
//   sched 1 (default): every (tile, chunk group) pair is swept over the WHOLE time range by one CTA -- floor(pairs / grid) full
//     rounds --, and the pairs that are left are cut into m time ranges each, m chosen to minimise the number of time steps the
//     left-over rounds take (ranges cost two extra centre loads each).  32^3 x 64 x 12 with G = 3: 256 pairs on 148 CTAs = one
//     full round + 108 pairs x 4 ranges in three short rounds (116 time steps per CTA, 111 if the work were divisible).
//   sched 0: fixed ranges of CGPTB_TMA_TRL slices for every pair (the round-1 schedule), kept for A/B runs.
static std::vector<ItemDesc> build_schedule(const Geo& geo, int grid, int t_begin, int t_count, int sched, int trl_max) {
  std::vector<ItemDesc> items;
  const int ntile = geo.nbx * geo.nby * geo.nbz, npairs = ntile * geo.ngroup;
  auto make = [&](int pair, int t0, int trl) {
    ItemDesc it;
    it.c = pair % geo.ngroup;
    int r = pair / geo.ngroup;
    it.xh0 = (r % geo.nbx) * TX;
    r /= geo.nbx;
    it.y0 = (r % geo.nby) * TY;
    it.z0 = (r / geo.nby) * TZ;
    it.t0 = t0;
    it.trl = trl;
    it.pad0 = it.pad1 = 0;
    return it;
  };
  if (sched == 0) {
    int trl = 1;
    for (int d = 1; d <= t_count && d <= trl_max; d++)
      if (t_count % d == 0) trl = d;
    for (int tr = 0; tr < t_count / trl; tr++)
      for (int pair = 0; pair < npairs; pair++) items.push_back(make(pair, t_begin + tr * trl, trl));
    return items;
  }
  const int rounds = npairs / grid, left = npairs % grid;
  for (int pair = 0; pair < rounds * grid; pair++) items.push_back(make(pair, t_begin, t_count));
  if (left) {
    // ranges shorter than 4 slices pay too much for their two extra loads
    const int mmax = t_count >= 8 ? t_count / 4 : 1;
    int m = 1;
    double best = 1e30;
    for (int k = 1; k <= mmax; k++) {
      const int r = (left * k + grid - 1) / grid;
      const double cost = r * ((t_count + k - 1) / k + 1.0);
      if (cost < best - 1e-9) {
        best = cost;
        m = k;
      }
    }
    // range-major: the pairs of one range are spatial neighbours and run concurrently
    int t0 = t_begin;
    for (int j = 0; j < m; j++) {
      const int trl = t_count / m + (j < t_count % m ? 1 : 0);
      for (int q = 0; q < left; q++) items.push_back(make(rounds * grid + q, t0, trl));
      t0 += trl;
    }
  }
  return items;
}

struct SchedKey {
  int v[12];
  bool operator<(const SchedKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
static std::map<SchedKey, std::pair<ItemDesc*, int>> g_sched;

static const ItemDesc* schedule_on_device(const Geo& G, int grid, int t_begin, int t_count, int sched, int trl_max, int* nitems) {
  SchedKey key = {{G.hx, G.Ly, G.Lz, G.T, G.ngroup, grid, t_begin, t_count, sched, trl_max, G.nbx, G.nby}};
  auto f = g_sched.find(key);
  if (f == g_sched.end()) {
    std::vector<ItemDesc> items = build_schedule(G, grid, t_begin, t_count, sched, trl_max);
    ItemDesc* d = 0;
    CUDA_CHECK(cudaMalloc(&d, items.size() * sizeof(ItemDesc)));
    // pageable source: the call returns once the host data has been staged; ordered before the launch on the same stream
    CUDA_CHECK(cudaMemcpyAsync(d, items.data(), items.size() * sizeof(ItemDesc), cudaMemcpyHostToDevice, g_stream));
    CUDA_CHECK(cudaStreamSynchronize(g_stream));
    f = g_sched.emplace(key, std::make_pair(d, (int)items.size())).first;
  }
  *nitems = f->second.second;
  return f->second.first;
}

}  // namespace tma

// true if the TMA sweep kernel handles this operator / lattice (single GPU, Ls a multiple of 4, extents divisible by
// the tile); the caller falls back to the kernels of dslash_f32.cu otherwise
bool dhop_tma_usable(const cgptb_fermion_operator* op) {
  if (tma::env_i("CGPTB_NO_TMA", 0) || op->prec != CGPTB_SINGLE) return false;
  if (op->g.comm_mask & 3) return false;  // x / y splits: the boundary predicates would diverge inside a warp
  const Geom& g = op->g;
  if (op->ls() % tma::SC) return false;
  if (g.hx % tma::TX || g.L[1] % tma::TY || g.L[2] % tma::TZ || g.L[3] < 2) return false;
  return true;
}

void dhop_tma_release(cgptb_fermion_operator* op) {
  for (int p = 0; p < 2; p++)
    if (op->links_pad[p]) {
      cudaFree(op->links_pad[p]);
      op->links_pad[p] = 0;
    }
}

void dhop_half_f32_tma(cgptb_fermion_operator* op, bool dag, const float* pin, size_t in_stride, float* pout, size_t out_stride,
                       int p_out, int t_begin, int t_count) {
  // t_count <= 0: the whole lattice; otherwise output time slices [t_begin, t_begin + t_count) only
  using namespace tma;
  const Geom& g = op->g;
  const int ls = op->ls();
  const bool cmp = op->compress;
  if (!op->links_pad_valid) {
    size_t n4 = (size_t)g.half4;
    for (int p = 0; p < 2; p++) {
      if (!op->links_pad[p]) CUDA_CHECK(cudaMalloc(&op->links_pad[p], n4 * 2 * LINK_ROW_B));
      if (cmp) {
        size_t n = n4 * 2 * LINKC_ROW_F;
        k_pad_links_c<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(n4, (const float*)op->links_c[p], (float*)op->links_pad[p]);
      } else {
        size_t n = n4 * 2 * 38;
        k_pad_links<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(n4, (const float2*)op->links[p], (float2*)op->links_pad[p]);
      }
      LAUNCH_CHECK();
    }
    op->links_pad_valid = true;
  }
  Geo G;
  G.hx = g.hx;
  G.Ly = g.L[1];
  G.Lz = g.L[2];
  G.T = g.L[3];
  G.nbx = g.hx / TX;
  G.nby = g.L[1] / TY;
  G.nbz = g.L[2] / TZ;
  // chunks per CTA.  Default 1: measured on B200 at 32^3 x 64 x 12 (profiles/ablation_r2.txt) G = 1 and G = 3 are within 3 % of
  // each other -- G = 3 moves 7 % less DRAM and 20 % less L2->SM traffic, G = 1 has a third of the code and the own-spinor
  // reuse --, G = 1 ahead.  CGPTB_TMA_G selects another divisor of Ls / 4 up to GMAX.
  const int nchunk = ls / SC;
  int ng = 1;
  const int ng_env = env_i("CGPTB_TMA_G", 0);
  if (ng_env >= 1 && ng_env <= GMAX && nchunk % ng_env == 0) ng = ng_env;
  G.ngroup = nchunk / ng;
  if (t_count <= 0) {
    t_begin = 0;
    t_count = G.T;
  }
  CGPTB_ASSERT(t_begin >= 0 && t_begin + t_count <= G.T);
  G.p_out = p_out;
  G.ls = ls;
  G.skip = env_i("CGPTB_TMA_SKIP", 0);
  G.comm_mask = g.comm_mask;
  // the field as a 5-d tensor (s * 8 floats, x/2, y, z + Lz * t, component plane); boxes with all three planes: one x column,
  // one y row, one z layer, the tile
  CUtensorMap tmX, tmY, tmZ, tmC, tmL;
  {
    const cuuint64_t dims[5] = {(cuuint64_t)ls * 8, (cuuint64_t)g.hx, (cuuint64_t)g.L[1], (cuuint64_t)g.L[2] * g.L[3], 3};
    const cuuint64_t row = (cuuint64_t)ls * 32;
    const cuuint64_t strides[4] = {row, row * g.hx, row * g.hx * g.L[1], (cuuint64_t)in_stride * 32};
    const cuuint32_t bx[5] = {SC * 8, 1, TY, TZ, 3}, by[5] = {SC * 8, TX, 1, TZ, 3}, bz[5] = {SC * 8, TX, TY, 1, 3}, bc[5] = {SC * 8, TX, TY, TZ, 3};
    encode5(&tmX, pin, dims, strides, bx, CU_TENSOR_MAP_SWIZZLE_128B);
    encode5(&tmY, pin, dims, strides, by, CU_TENSOR_MAP_SWIZZLE_128B);
    encode5(&tmZ, pin, dims, strides, bz, CU_TENSOR_MAP_SWIZZLE_128B);
    encode5(&tmC, pin, dims, strides, bc, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  {
    const cuuint64_t rowf = cmp ? LINKC_ROW_F : LINK_ROW_F;
    const cuuint64_t dims[5] = {rowf, (cuuint64_t)g.hx, (cuuint64_t)g.L[1], (cuuint64_t)g.L[2], (cuuint64_t)2 * g.L[3]};  // last index = half * T + t
    const cuuint64_t row = rowf * 4;
    const cuuint64_t strides[4] = {row, row * g.hx, row * g.hx * g.L[1], row * g.hx * g.L[1] * g.L[2]};
    const cuuint32_t bl[5] = {(cuuint32_t)rowf, TX, TY, TZ, 1};
    encode5(&tmL, op->links_pad[p_out], dims, strides, bl, CU_TENSOR_MAP_SWIZZLE_NONE);
  }
  const int grid_env = env_i("CGPTB_TMA_GRID", 0);
  // a persistent grid on every SM would keep the NCCL send/recv kernels of the halo exchange from being scheduled until
  // the stencil is done, so a split lattice leaves a few SMs free for them (CGPTB_TMA_COMM_SMS) -- unless the halo goes
  // through peer memory (halo.cu), which needs no communication kernel
  int grid = grid_env > 0 ? grid_env : sm_count() - (g.comm_mask && !halo_is_p2p(op) ? env_i("CGPTB_TMA_COMM_SMS", 8) : 0);
  if (grid < 1) grid = 1;
  const ItemDesc* items = schedule_on_device(G, grid, t_begin, t_count, env_i("CGPTB_TMA_SCHED", 1), env_i("CGPTB_TMA_TRL", 16), &G.nitems);
  if (grid > G.nitems) grid = G.nitems;
  const int abl = g.comm_mask ? 0 : env_i("CGPTB_ABLATE", 0);
  typedef void (*kernel_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const Geo,
                           const ItemDesc*, float*, size_t);
  kernel_t kern = 0;
  int smem_b = 0;
#define TMA_PICK(G_)                                                                                          \
  if (ng == G_) {                                                                                             \
    smem_b = Ring<G_>::SMEM_B;                                                                                \
    if (cmp) {                                                                                                \
      if (g.comm_mask)                                                                                        \
        kern = dag ? k_dhop_f32_tma<true, 0, true, G_, true> : k_dhop_f32_tma<false, 0, true, G_, true>;      \
      else                                                                                                    \
        kern = dag ? k_dhop_f32_tma<true, 0, false, G_, true> : k_dhop_f32_tma<false, 0, false, G_, true>;    \
    } else if (g.comm_mask)                                                                                   \
      kern = dag ? k_dhop_f32_tma<true, 0, true, G_, false> : k_dhop_f32_tma<false, 0, true, G_, false>;      \
    else if (abl == 1)                                                                                        \
      kern = k_dhop_f32_tma<false, 1, false, G_, false>;                                                      \
    else if (abl == 2)                                                                                        \
      kern = k_dhop_f32_tma<false, 2, false, G_, false>;                                                      \
    else                                                                                                      \
      kern = dag ? k_dhop_f32_tma<true, 0, false, G_, false> : k_dhop_f32_tma<false, 0, false, G_, false>;    \
  }
  TMA_PICK(1)
  TMA_PICK(2)
  TMA_PICK(3)
#undef TMA_PICK
  CGPTB_ASSERT(kern != 0);
  static std::map<kernel_t, bool> configured;
  if (!configured[kern]) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
    configured[kern] = true;
  }
  kern<<<grid, NTHREADS, smem_b, g_stream>>>(tmX, tmY, tmZ, tmC, tmL, G, items, pout, out_stride);
  LAUNCH_CHECK();
}

}  // namespace cgptb

// the work schedule of the TMA sweep kernel as a table (host only; tests/test_cpu_boundary.py checks that every (tile, chunk
// group, time slice) is covered exactly once): item i = (group, x/2 origin, y origin, z origin, first slice, number of slices)
extern "C" int cgptb_debug_tma_schedule(const int dims4[4], int Ls, int grid, int chunks_per_cta, int sched, int trl, int t_begin,
                                        int t_count, int* out6, int max_items, int* n_items) {
  using namespace cgptb;
  using namespace cgptb::tma;
  CGPTB_API_BEGIN
  if (chunks_per_cta < 1 || dims4[0] % (2 * TX) || dims4[1] % TY || dims4[2] % TZ || Ls % SC || (Ls / SC) % chunks_per_cta || grid < 1)
    CGPTB_ERR("lattice %d.%d.%d.%d Ls=%d is not tiled by the sweep kernel with %d chunks per CTA", dims4[0], dims4[1], dims4[2], dims4[3], Ls, chunks_per_cta);
  Geo G;
  memset(&G, 0, sizeof(G));
  G.hx = dims4[0] / 2;
  G.Ly = dims4[1];
  G.Lz = dims4[2];
  G.T = dims4[3];
  G.nbx = G.hx / TX;
  G.nby = G.Ly / TY;
  G.nbz = G.Lz / TZ;
  G.ngroup = Ls / SC / chunks_per_cta;
  if (t_count <= 0) {
    t_begin = 0;
    t_count = G.T;
  }
  std::vector<ItemDesc> items = build_schedule(G, grid, t_begin, t_count, sched, trl);
  *n_items = (int)items.size();
  for (int i = 0; i < (int)items.size() && i < max_items; i++) {
    const ItemDesc& it = items[i];
    int* o = out6 + 6 * i;
    o[0] = it.c;
    o[1] = it.xh0;
    o[2] = it.y0;
    o[3] = it.z0;
    o[4] = it.t0;
    o[5] = it.trl;
  }
  CGPTB_API_END
}

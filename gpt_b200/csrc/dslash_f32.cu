// Single-precision Wilson / Moebius hopping kernels on the packed FFMA2 pipe (sm_100a).
//
// All complex arithmetic is done on (re, im) register pairs: a complex multiply-add is two FFMA2, spin
// projection and reconstruction are FADD2 with swap / negate operand modifiers, so the kernel issues ~430
// packed FP instructions per site instead of ~770 scalar ones (SURVEY.md section 7 "hard parts").
//
// k_dhop_f32_tiled (the production kernel):
//   * one CTA = one 4d tile of NS checkerboard sites (default 4x4x2x2 in x,y,z,t = 32 sites of one parity)
//     times all Ls slices, one thread per output (site, s); the neighbours of a tile overlap heavily, so the
//     unified L1 serves about half of the 8 neighbour reads;
//   * the tile's 8 x NS gauge links are staged once into shared memory (padded so the 3-4 distinct link
//     addresses a warp reads land in different banks) and broadcast from there -- in the untiled kernel
//     link loads through L1 were 45 % of all L1 wavefronts;
//   * CTAs walk the lattice in z-slabs (x, y, z-in-slab, t, slab): the +-t neighbour reuse distance shrinks
//     from a whole time slice to a slab of it and stays resident in L2; outputs are written with streaming
//     stores so they do not evict the input window.
// k_dhop_f32 (fallback for extents the tile does not divide): same arithmetic, linear site order.
#include <stdio.h>
#include <stdlib.h>
#include "dslash.cuh"
#include "operator.cuh"
#include "packed.cuh"
#include "sweep.cuh"
#include <string.h>

namespace cgptb {

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// a spinor is 3 blocks of 32 bytes (4 complex each): 3 x LDG.E.ENL2.256 per neighbour
__device__ __forceinline__ void load_spinor_c32(const float* __restrict__ base, size_t nsites, size_t site, c32 (&p)[12]) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float v[8];
    ld256(base + (k * nsites + site) * 8, v);
#pragma unroll
    for (int e = 0; e < 4; e++) p[4 * k + e] = pk(v[2 * e], v[2 * e + 1]);
  }
}

// same with the three plane base pointers precomputed (uniform) and a 32-bit site index: two integer
// instructions per plane instead of a 64-bit multiply-add chain
__device__ __forceinline__ void load_spinor_c32_planes(const float* __restrict__ p0, const float* __restrict__ p1,
                                                       const float* __restrict__ p2, unsigned site, c32 (&p)[12]) {
  const size_t off = (size_t)site * 8;
  float v[8];
  ld256(p0 + off, v);
#pragma unroll
  for (int e = 0; e < 4; e++) p[e] = pk(v[2 * e], v[2 * e + 1]);
  ld256(p1 + off, v);
#pragma unroll
  for (int e = 0; e < 4; e++) p[4 + e] = pk(v[2 * e], v[2 * e + 1]);
  ld256(p2 + off, v);
#pragma unroll
  for (int e = 0; e < 4; e++) p[8 + e] = pk(v[2 * e], v[2 * e + 1]);
}

template <bool STREAM>
__device__ __forceinline__ void store_spinor_c32(float* __restrict__ base, size_t nsites, size_t site, const c32 (&p)[12]) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 4; e++) upk(p[4 * k + e], v[2 * e], v[2 * e + 1]);
    if (STREAM)
      st256_cs(base + (k * nsites + site) * 8, v);
    else
      st256(base + (k * nsites + site) * 8, v);
  }
}

// one direction: acc += recon( W(^dag) proj psi(neighbour) ), link given as 9 (re,im) pairs
template <int MU, bool FWD, bool DAG>
__device__ __forceinline__ void hop_core(c32 (&acc)[12], const c32 (&psi)[12], const float (&wr)[9], const float (&wi)[9]) {
  const int SGN = (FWD != DAG) ? -1 : +1;
  typedef Proj<MU, SGN> P;
  c32 h[6];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    h[c] = add2(psi[c], times_iph<P::A>(psi[P::J0 * 3 + c]));
    h[3 + c] = add2(psi[3 + c], times_iph<P::B>(psi[P::J1 * 3 + c]));
  }
  c32 chi[6];
#pragma unroll
  for (int sp = 0; sp < 2; sp++) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
      c32 a;
      if (FWD) {  // W h
        a = cmul<false>(wr[r * 3 + 0], wi[r * 3 + 0], h[sp * 3 + 0]);
        a = cmac<false>(a, wr[r * 3 + 1], wi[r * 3 + 1], h[sp * 3 + 1]);
        a = cmac<false>(a, wr[r * 3 + 2], wi[r * 3 + 2], h[sp * 3 + 2]);
      } else {  // W^dag h
        a = cmul<true>(wr[0 * 3 + r], wi[0 * 3 + r], h[sp * 3 + 0]);
        a = cmac<true>(a, wr[1 * 3 + r], wi[1 * 3 + r], h[sp * 3 + 1]);
        a = cmac<true>(a, wr[2 * 3 + r], wi[2 * 3 + r], h[sp * 3 + 2]);
      }
      chi[sp * 3 + r] = a;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    acc[c] = add2(acc[c], chi[c]);
    acc[3 + c] = add2(acc[3 + c], chi[3 + c]);
    acc[6 + c] = add2(acc[6 + c], times_iph<P::C2>(chi[P::K2 * 3 + c]));
    acc[9 + c] = add2(acc[9 + c], times_iph<P::C3>(chi[P::K3 * 3 + c]));
  }
}

template <int MU, bool FWD, bool DAG, int ABL = 0>
__device__ __forceinline__ void hop_global_links(c32 (&acc)[12], const Geom& g, int x, int y, int z, int t, int i4, int s, int ls,
                                                 const float* __restrict__ in, size_t in_stride, const float* __restrict__ links) {
  if (off_rank<MU, FWD>(g, x, y, z, t)) return;
  int n4 = neighbor<MU, FWD>(g, x, y, z, t);
  if (ABL == 1) n4 = i4;  // ablation: no neighbour traffic (every direction reads the site itself)
  c32 psi[12];
  load_spinor_c32(in, in_stride, (size_t)n4 * ls + s, psi);
  if (ABL == 2) {  // ablation: no arithmetic
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = add2(acc[k], psi[k]);
    return;
  }
  float wr[9], wi[9];
  const float2* lb = reinterpret_cast<const float2*>(links) + ((size_t)i4 * 8 + (FWD ? MU : MU + 4)) * 9;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    float2 v = __ldg(lb + k);
    wr[k] = v.x;
    wi[k] = v.y;
  }
  hop_core<MU, FWD, DAG>(acc, psi, wr, wi);
}

template <bool DAG, int LS, int ABL = 0>
__global__ void __launch_bounds__(128) k_dhop_f32(Geom g, int ls_rt, int p_out, const float* __restrict__ in, size_t in_stride,
                                                  float* __restrict__ out, size_t out_stride, const float* __restrict__ links) {
  const int ls = LS > 0 ? LS : ls_rt;
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (size_t)g.half4 * ls) return;
  int i4 = (int)(tid / ls);
  int s = (int)(tid - (size_t)i4 * ls);
  int x, y, z, t;
  cb_coords(g, p_out, i4, x, y, z, t);
  c32 acc[12];
#pragma unroll
  for (int k = 0; k < 12; k++) acc[k] = 0ull;
  hop_global_links<0, true, DAG, ABL>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_global_links<0, false, DAG, ABL>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_global_links<1, true, DAG, ABL>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_global_links<1, false, DAG, ABL>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_global_links<2, true, DAG, ABL>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_global_links<2, false, DAG, ABL>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_global_links<3, true, DAG, ABL>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_global_links<3, false, DAG, ABL>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  store_spinor_c32<false>(out, out_stride, tid, acc);
}

// ---- tiled kernel ------------------------------------------------------------------------------------
struct TileGeom {
  int txh, ty, tz, tt;   // tile extents in checkerboard coordinates (xh = x/2)
  int nxh, ny, nz, nt;   // tiles per dimension
  int zslab;             // z-tiles per slab
  int stream_stores;     // write the output with st.global.cs
  const int2* zt_table;  // blockIdx.z -> (z tile, t tile), slab order (device memory)
};

static const int LINK_F4 = 37;  // float4 per site in shared memory: 36 (8 links x 72 B) + 1 pad -> 148 words, bank shift 20

// link d of site l from shared memory; 72 B blocks are 16-byte aligned for even d, 8 mod 16 for odd d
template <int D>
__device__ __forceinline__ void load_link_smem(const float4* __restrict__ slinks, int l, float (&wr)[9], float (&wi)[9]) {
  const float* base = reinterpret_cast<const float*>(slinks + l * LINK_F4) + D * 18;
  float v[18];
  if (D % 2 == 0) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float4 q = *reinterpret_cast<const float4*>(base + 4 * k);
      v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
    float2 q = *reinterpret_cast<const float2*>(base + 16);
    v[16] = q.x; v[17] = q.y;
  } else {
    float2 q = *reinterpret_cast<const float2*>(base);
    v[0] = q.x; v[1] = q.y;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float4 q4 = *reinterpret_cast<const float4*>(base + 2 + 4 * k);
      v[2 + 4 * k] = q4.x; v[3 + 4 * k] = q4.y; v[4 + 4 * k] = q4.z; v[5 + 4 * k] = q4.w;
    }
  }
#pragma unroll
  for (int k = 0; k < 9; k++) {
    wr[k] = v[2 * k];
    wi[k] = v[2 * k + 1];
  }
}

// one direction for the SPER s-slices a thread owns: the link is read once, the spinors are independent loads
template <int MU, bool FWD, bool DAG, int LS, int SPER, int ABL, bool COMM>
__device__ __forceinline__ void hop_tile(c32 (&acc)[SPER][12], const Geom& g, int x, int y, int z, int t, int l, int j, int i4,
                                         const float* __restrict__ in, size_t in_stride, const float4* __restrict__ slinks) {
  constexpr int TPS = LS / SPER;
  if (COMM && off_rank<MU, FWD>(g, x, y, z, t)) return;  // COMM = false: single GPU, no boundary predicates at all
  int n4 = neighbor<MU, FWD>(g, x, y, z, t);
  if (ABL == 1) n4 = i4;
  c32 psi[SPER][12];
#pragma unroll
  for (int r = 0; r < SPER; r++)
    load_spinor_c32_planes(in, in + in_stride * 8, in + in_stride * 16, (unsigned)(n4 * LS + j + r * TPS), psi[r]);
  if (ABL == 2) {
#pragma unroll
    for (int r = 0; r < SPER; r++)
#pragma unroll
      for (int k = 0; k < 12; k++) acc[r][k] = add2(acc[r][k], psi[r][k]);
    return;
  }
  float wr[9], wi[9];
  load_link_smem<FWD ? MU : MU + 4>(slinks, l, wr, wi);
#pragma unroll
  for (int r = 0; r < SPER; r++) hop_core<MU, FWD, DAG>(acc[r], psi[r], wr, wi);
}

// optional fused epilogue of the tiled kernel (all of it after the 8 hops, before the store):
//   1. SWEEP : result <- T result, T a fifth-dimension sweep (sweep.cuh), e.g. Meooe5D o MooeeInv of the NEXT factor
//   2. AXPY  : result <- z - result            (the "o = i - ..." of schur_complement_two._N/_N_dag)
//   3. DOT   : partial[cta] = sum conj(dotp) * result, |result|^2 in double   (cg.py's <p, A p>)
struct EpiArgs {
  const float* z;
  size_t z_stride;
  const float* dotp;
  size_t dot_stride;
  double* partial;
  SweepParams<float> P;
};

template <bool DAG, int LS, int SPER, int NS, int MINB, int ABL, bool EPI, bool COMM>
__global__ void __launch_bounds__(NS* LS / SPER, MINB)
    k_dhop_f32_tile(Geom g, TileGeom tg, int p_out, const float* __restrict__ in, size_t in_stride, float* __restrict__ out,
                    size_t out_stride, const float* __restrict__ links, EpiArgs epi) {
  constexpr int TPS = LS / SPER;  // threads per 4d site
  constexpr int NT = NS * LS / SPER;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* slinks = reinterpret_cast<float4*>(smem_raw);  // [NS][LINK_F4]
  // tile coordinates: grid = (x tiles, y tiles, (z in slab, t, slab)); the last index is decoded with a small table so
  // that no thread executes an integer division by a run-time value
  constexpr int TXH = NS == 32 ? 4 : 2, TY = 2, TZ = 2;  // tile = TXH x 2 x 2 x 2 checkerboard sites
  const int bx = blockIdx.x, by = blockIdx.y;
  const int2 zt = tg.zt_table[blockIdx.z];
  const int bz = zt.x, bt = zt.y;

  const int l = threadIdx.x / TPS;
  const int j = threadIdx.x - l * TPS;
  const int lx = l % TXH, ly = (l / TXH) % TY, lz = (l / (TXH * TY)) % TZ, lt = l / (TXH * TY * TZ);
  const int xh = bx * TXH + lx, y = by * TY + ly, z = bz * TZ + lz, t = bt * 2 + lt;
  const int i4 = xh + g.hx * (y + g.L[1] * (z + g.L[2] * t));
  const int x = 2 * xh + ((y + z + t + p_out) & 1);

  // stage this site's 8 links (576 B) with the TPS threads that own it: 16-byte cp.async, L2 only
  {
    const float4* gl = reinterpret_cast<const float4*>(links) + (size_t)i4 * 36;
    unsigned sbase = (unsigned)__cvta_generic_to_shared(slinks + l * LINK_F4);
    if (36 % TPS == 0) {
#pragma unroll
      for (int i = 0; i < 36 / TPS; i++) {
        int m = j + i * TPS;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + m * 16), "l"(gl + m));
      }
    } else {
      for (int m = j; m < 36; m += TPS)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + m * 16), "l"(gl + m));
    }
    asm volatile("cp.async.commit_group;");
  }
  c32 acc[SPER][12];
#pragma unroll
  for (int r = 0; r < SPER; r++)
#pragma unroll
    for (int k = 0; k < 12; k++) acc[r][k] = 0ull;
  asm volatile("cp.async.wait_group 0;");
  __syncthreads();
  hop_tile<0, true, DAG, LS, SPER, ABL, COMM>(acc, g, x, y, z, t, l, j, i4, in, in_stride, slinks);
  hop_tile<0, false, DAG, LS, SPER, ABL, COMM>(acc, g, x, y, z, t, l, j, i4, in, in_stride, slinks);
  hop_tile<1, true, DAG, LS, SPER, ABL, COMM>(acc, g, x, y, z, t, l, j, i4, in, in_stride, slinks);
  hop_tile<1, false, DAG, LS, SPER, ABL, COMM>(acc, g, x, y, z, t, l, j, i4, in, in_stride, slinks);
  hop_tile<2, true, DAG, LS, SPER, ABL, COMM>(acc, g, x, y, z, t, l, j, i4, in, in_stride, slinks);
  hop_tile<2, false, DAG, LS, SPER, ABL, COMM>(acc, g, x, y, z, t, l, j, i4, in, in_stride, slinks);
  hop_tile<3, true, DAG, LS, SPER, ABL, COMM>(acc, g, x, y, z, t, l, j, i4, in, in_stride, slinks);
  hop_tile<3, false, DAG, LS, SPER, ABL, COMM>(acc, g, x, y, z, t, l, j, i4, in, in_stride, slinks);

  if (EPI) {
    constexpr int PITCH = LS + 1;
    if (epi.P.nstages > 0) {
      // transpose through shared memory: row (k, l) holds the Ls values of one 16-byte component block
      float4* se = slinks + NS * LINK_F4;  // [6][NS][PITCH]
#pragma unroll
      for (int r = 0; r < SPER; r++) {
        int s = j + r * TPS;
#pragma unroll
        for (int k = 0; k < 6; k++) {
          float4 v;
          upk(acc[r][2 * k], v.x, v.y);
          upk(acc[r][2 * k + 1], v.z, v.w);
          se[(k * NS + l) * PITCH + s] = v;
        }
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < NS * 6; idx += NT) {
        int l2 = idx % NS, k = idx / NS;
        sweep_row<float, LS>(epi.P, k, se + (k * NS + l2) * PITCH);
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < SPER; r++) {
        int s = j + r * TPS;
#pragma unroll
        for (int k = 0; k < 6; k++) {
          float4 v = se[(k * NS + l) * PITCH + s];
          acc[r][2 * k] = pk(v.x, v.y);
          acc[r][2 * k + 1] = pk(v.z, v.w);
        }
      }
    }
    if (epi.z) {
#pragma unroll
      for (int r = 0; r < SPER; r++) {
        c32 zz[12];
        load_spinor_c32(epi.z, epi.z_stride, (size_t)i4 * LS + j + r * TPS, zz);
#pragma unroll
        for (int k = 0; k < 12; k++) acc[r][k] = add2(zz[k], times_iph<2>(acc[r][k]));
      }
    }
    if (epi.partial) {
      double v[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int r = 0; r < SPER; r++) {
        c32 pp[12];
        load_spinor_c32(epi.dotp, epi.dot_stride, (size_t)i4 * LS + j + r * TPS, pp);
#pragma unroll
        for (int k = 0; k < 12; k++) {
          float ar, ai, br, bi;
          upk(pp[k], ar, ai);
          upk(acc[r][k], br, bi);
          v[0] += (double)ar * br + (double)ai * bi;
          v[1] += (double)ar * bi - (double)ai * br;
          v[2] += (double)br * br + (double)bi * bi;
        }
      }
      __shared__ double red[96];
      block_reduce<3>(v, red);
      if (threadIdx.x == 0) {
        size_t cta = blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z);
        epi.partial[cta * 3 + 0] = v[0];
        epi.partial[cta * 3 + 1] = v[1];
        epi.partial[cta * 3 + 2] = v[2];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < SPER; r++) {
    if (tg.stream_stores)
      store_spinor_c32<true>(out, out_stride, (size_t)i4 * LS + j + r * TPS, acc[r]);
    else
      store_spinor_c32<false>(out, out_stride, (size_t)i4 * LS + j + r * TPS, acc[r]);
  }
}

template <bool DAG>
static void launch_linear(int ls, unsigned blocks, int threads, const Geom& g, int p_out, const float* in, size_t is, float* out,
                          size_t os, const float* links) {
  static int abl = env_int("CGPTB_ABLATE", 0);
  if (ls == 12 && abl == 1) { k_dhop_f32<DAG, 12, 1><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); return; }
  if (ls == 12 && abl == 2) { k_dhop_f32<DAG, 12, 2><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); return; }
  switch (ls) {
    case 1: k_dhop_f32<DAG, 1><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
    case 8: k_dhop_f32<DAG, 8><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
    case 12: k_dhop_f32<DAG, 12><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
    case 16: k_dhop_f32<DAG, 16><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
    case 24: k_dhop_f32<DAG, 24><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
    default: k_dhop_f32<DAG, 0><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
  }
}

// tile of NS checkerboard sites (NS = 32: 4 x 2 x 2 x 2 in xh,y,z,t ; NS = 16: 2 x 2 x 2 x 2); returns false if the
// lattice is not divisible.  The (z,t) tile order table is cached per geometry.
static bool make_tiles(const Geom& g, int ns, TileGeom& tg) {
  tg.txh = ns == 32 ? 4 : 2;
  tg.ty = 2;
  tg.tz = 2;
  tg.tt = 2;
  if (ns != 32 && ns != 16) return false;
  if (g.hx % tg.txh || g.L[1] % tg.ty || g.L[2] % tg.tz || g.L[3] % tg.tt) return false;
  tg.nxh = g.hx / tg.txh;
  tg.ny = g.L[1] / tg.ty;
  tg.nz = g.L[2] / tg.tz;
  tg.nt = g.L[3] / tg.tt;
  if (tg.ny > 65535 || tg.nz * tg.nt > 65535) return false;
  static int zslab_sites = env_int("CGPTB_ZSLAB", 16);  // slab width in z (sites)
  static int stcs = env_int("CGPTB_STCS", 1);
  tg.stream_stores = stcs;
  tg.zslab = zslab_sites / tg.tz;
  if (tg.zslab < 1) tg.zslab = 1;
  if (tg.zslab > tg.nz) tg.zslab = tg.nz;
  // order: z inside the slab fastest, then t, then the slab
  struct Cache {
    int nz, nt, zslab;
    int2* dev;
  };
  static std::vector<Cache> cache;
  for (auto& c : cache)
    if (c.nz == tg.nz && c.nt == tg.nt && c.zslab == tg.zslab) {
      tg.zt_table = c.dev;
      return true;
    }
  std::vector<int2> h;
  for (int slab0 = 0; slab0 < tg.nz; slab0 += tg.zslab) {
    int zw = slab0 + tg.zslab > tg.nz ? tg.nz - slab0 : tg.zslab;
    for (int t = 0; t < tg.nt; t++)
      for (int zi = 0; zi < zw; zi++) h.push_back(make_int2(slab0 + zi, t));
  }
  int2* dev;
  CUDA_CHECK(cudaMalloc(&dev, h.size() * sizeof(int2)));
  CUDA_CHECK(cudaMemcpy(dev, h.data(), h.size() * sizeof(int2), cudaMemcpyHostToDevice));
  cache.push_back({tg.nz, tg.nt, tg.zslab, dev});
  tg.zt_table = dev;
  return true;
}

template <bool DAG, int LS, int SPER, int NS, int MINB, int ABL, bool EPI, bool COMM = false>
static void launch_tile_k(const Geom& g, const TileGeom& tg, int p_out, const float* in, size_t is, float* out, size_t os,
                          const float* links, const EpiArgs& epi) {
  dim3 blocks(tg.nxh, tg.ny, tg.nz * tg.nt);
  size_t smem = (size_t)NS * LINK_F4 * 16 + (EPI ? (size_t)6 * NS * (LS + 1) * 16 : 0);
  auto kern = k_dhop_f32_tile<DAG, LS, SPER, NS, MINB, ABL, EPI, COMM>;
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  kern<<<blocks, NS * LS / SPER, smem, g_stream>>>(g, tg, p_out, in, is, out, os, links, epi);
}

template <bool DAG, int LS, int SPER, int NS, int MINB>
static bool launch_tile_t(const Geom& g, int p_out, const float* in, size_t is, float* out, size_t os, const float* links,
                          const EpiArgs* epi) {
  TileGeom tg;
  if (!make_tiles(g, NS, tg)) return false;
  static int abl = env_int("CGPTB_ABLATE", 0);
  if (epi) {
    launch_tile_k<DAG, LS, SPER, NS, MINB, 0, true>(g, tg, p_out, in, is, out, os, links, *epi);
    return true;
  }
  EpiArgs none;
  memset(&none, 0, sizeof(none));
  if (abl == 1)
    launch_tile_k<DAG, LS, SPER, NS, MINB, 1, false>(g, tg, p_out, in, is, out, os, links, none);
  else if (abl == 2)
    launch_tile_k<DAG, LS, SPER, NS, MINB, 2, false>(g, tg, p_out, in, is, out, os, links, none);
  else if (g.comm_mask)
    launch_tile_k<DAG, LS, SPER, NS, MINB, 0, false, true>(g, tg, p_out, in, is, out, os, links, none);
  else
    launch_tile_k<DAG, LS, SPER, NS, MINB, 0, false>(g, tg, p_out, in, is, out, os, links, none);
  return true;
}

template <bool DAG>
static bool launch_tiled(int ls, const Geom& g, int p_out, const float* in, size_t is, float* out, size_t os, const float* links,
                         const EpiArgs* epi) {
  static int variant = env_int("CGPTB_DHOP_VARIANT", 0);
  switch (ls) {
    case 8: return launch_tile_t<DAG, 8, 1, 32, 2>(g, p_out, in, is, out, os, links, epi);
    case 12:
      if (variant == 7 && !epi) return launch_tile_t<DAG, 12, 1, 16, 4>(g, p_out, in, is, out, os, links, epi);
      if (variant == 3 && !epi) return launch_tile_t<DAG, 12, 1, 16, 5>(g, p_out, in, is, out, os, links, epi);
      return launch_tile_t<DAG, 12, 1, 32, 2>(g, p_out, in, is, out, os, links, epi);
    case 16: return launch_tile_t<DAG, 16, 1, 16, 2>(g, p_out, in, is, out, os, links, epi);
    case 24: return launch_tile_t<DAG, 24, 1, 16, 2>(g, p_out, in, is, out, os, links, epi);
    default: return false;
  }
}

int dhop_tile_blocks(const cgptb_fermion_operator* op) {
  TileGeom tg;
  int ns = (op->Ls == 16 || op->Ls == 24) ? 16 : 32;
  if (!make_tiles(op->g, ns, tg)) return 0;
  return tg.nxh * tg.ny * tg.nz * tg.nt;
}

// true if the fused-epilogue kernel exists for this operator / lattice
bool dhop_fusable(const cgptb_fermion_operator* op) {
  static int no_tiles = env_int("CGPTB_NO_TILES", 0);
  static int no_fuse = env_int("CGPTB_NO_EPILOGUE", 0);
  if (no_tiles || no_fuse || op->prec != CGPTB_SINGLE || op->type != CGPTB_MOBIUS || op->zmobius || op->g.comm_mask || op->compress) return false;
  if (!(op->Ls == 8 || op->Ls == 12 || op->Ls == 16 || op->Ls == 24)) return false;
  return dhop_tile_blocks(op) > 0;
}

void dhop_half_f32(cgptb_fermion_operator* op, bool dag, const float* pin, size_t in_stride, float* pout, size_t out_stride,
                   int p_out) {
  int ls = op->ls();
  static int no_tiles = env_int("CGPTB_NO_TILES", 0);
  const float* links = (const float*)op->links[p_out];
  bool done = false;
  if (dhop_tma_usable(op)) {
    dhop_half_f32_tma(op, dag, pin, in_stride, pout, out_stride, p_out);
    return;
  }
  if (!no_tiles) done = dag ? launch_tiled<true>(ls, op->g, p_out, pin, in_stride, pout, out_stride, links, 0)
                            : launch_tiled<false>(ls, op->g, p_out, pin, in_stride, pout, out_stride, links, 0);
  if (!done) {
    size_t half = (size_t)op->g.half4 * ls;
    int threads = 128;
    unsigned blocks = (unsigned)((half + threads - 1) / threads);
    if (dag)
      launch_linear<true>(ls, blocks, threads, op->g, p_out, pin, in_stride, pout, out_stride, links);
    else
      launch_linear<false>(ls, blocks, threads, op->g, p_out, pin, in_stride, pout, out_stride, links);
  }
  LAUNCH_CHECK();
}

// out(p_out) = [z -] [S] Dhop(^dag) in, S = fifth-dimension sweep `sweep_mode` (-1: none); optional partial sums
// of <dotp, out> and |out|^2 per CTA into `partial` (3 doubles per CTA, dhop_tile_blocks(op) CTAs)
void dhop_half_f32_fused(cgptb_fermion_operator* op, bool dag, const float* pin, size_t in_stride, float* pout, size_t out_stride,
                         int p_out, int sweep_mode, const float* z, size_t z_stride, const float* dotp, size_t dot_stride,
                         double* partial) {
  EpiArgs epi;
  memset(&epi, 0, sizeof(epi));
  if (sweep_mode >= 0) {
    if (!make_sweep_params<float>(op, sweep_mode, epi.P)) CGPTB_ERR("unknown sweep mode %d", sweep_mode);
  } else
    epi.P.nstages = 0;
  epi.z = z;
  epi.z_stride = z_stride;
  epi.dotp = dotp;
  epi.dot_stride = dot_stride;
  epi.partial = partial;
  const float* links = (const float*)op->links[p_out];
  bool done = dag ? launch_tiled<true>(op->ls(), op->g, p_out, pin, in_stride, pout, out_stride, links, &epi)
                  : launch_tiled<false>(op->ls(), op->g, p_out, pin, in_stride, pout, out_stride, links, &epi);
  if (!done) CGPTB_ERR("fused Dslash is not available for this lattice (check dhop_fusable first)");
  LAUNCH_CHECK();
}

}  // namespace cgptb

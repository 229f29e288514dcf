// Fermion operators of libcgpt_b200: Wilson-clover and Moebius domain wall.
// Replaces the cgpt fermion-operator entry points (lib/cgpt/lib/operators.cc:34-107) and the Grid kernels
// they dispatch to (lib/cgpt/lib/operators/unary.h:20-48, register.h:2-20):
//   Dhop/DhopEO(+Dag), Meooe(+Dag), Mooee(+Dag), MooeeInv(+Dag), M, Mdag, Mdiag, Dminus(+Dag),
//   Import/Export{Physical,Unphysical}Fermion{Source,Solution}.
#include <stdlib.h>
#include <complex>
#include "operator.cuh"
#include "dslash.cuh"

namespace cgptb {

// ----------------------------------------------------------------------------------------------------
// link import: fold -c_mu/2 and the boundary phases, double-store U_mu(x) and U_mu(x-mu) per parity
// (what Grid's ImportGauge / DoubleStore does; phase convention lib/gpt/core/covariant.py:29-37)
// ----------------------------------------------------------------------------------------------------
template <typename TU, typename T>
__global__ void k_build_links(Geom g, int p, size_t nsitesU, const TU* U0, const TU* U1, const TU* U2, const TU* U3,
                              LinkCoef lc, T* __restrict__ links, const double* gh1, const double* gh2, const double* gh3) {
  int i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= g.half4) return;
  int x, y, z, t;
  cb_coords(g, p, i4, x, y, z, t);
  const TU* U[4] = {U0, U1, U2, U3};
  for (int d = 0; d < 8; d++) {
    int mu = d & 3;
    int c[4] = {x, y, z, t};
    int pp = p;
    const double* ghost = 0;  // link owned by the rank-mu neighbour (split direction, low face)
    int gc = c[mu] + lc.goff[mu];  // global coordinate of the link's site in direction mu
    if (d >= 4) {
      if (c[mu] == 0 && ((g.comm_mask >> mu) & 1)) ghost = mu == 1 ? gh1 : (mu == 2 ? gh2 : gh3);
      c[mu] = c[mu] == 0 ? g.L[mu] - 1 : c[mu] - 1;
      gc = (gc - 1 + lc.gL[mu]) % lc.gL[mu];
      pp = 1 - p;
    }
    size_t site = (size_t)pp * g.half4 + cb_index(g, c[0], c[1], c[2], c[3]);
    size_t ft = 0;
    if (ghost) {  // transverse lexicographic index on the face (x fastest), as written by k_pack_links
      int o[2], n = 0;
      for (int e = 1; e < 4; e++)
        if (e != mu) o[n++] = e;
      ft = c[0] + (size_t)g.L[0] * (c[o[0]] + (size_t)g.L[o[0]] * c[o[1]]);
    }
    // phase multiplies U_mu on the last slice of direction mu (global coordinate)
    bool last = gc == lc.gL[mu] - 1;
    double fr = lc.w[mu] * (last ? lc.ph[2 * mu] : 1.0);
    double fi = lc.w[mu] * (last ? lc.ph[2 * mu + 1] : 0.0);
    // open boundary conditions: the operator is P D P with P the projector on 0 < t < T-1, i.e. nothing hops out of the
    // slices 0 and T-1 either: U_t(t = 0) and U_t(t = T-2) vanish like U_t(t = T-1) (phase 0) does
    if (lc.open_bc && mu == 3 && (gc == 0 || gc == lc.gL[3] - 2)) fr = fi = 0.0;
    for (int k = 0; k < 9; k++) {
      size_t o = elem_offset<TU>(nsitesU, site, k, 1);
      double ur = ghost ? ghost[(ft * 9 + k) * 2] : (double)U[mu][o];
      double ui = ghost ? ghost[(ft * 9 + k) * 2 + 1] : (double)U[mu][o + 1];
      links[((size_t)(i4 * 8 + d) * 9 + k) * 2] = (T)(fr * ur - fi * ui);
      links[((size_t)(i4 * 8 + d) * 9 + k) * 2 + 1] = (T)(fr * ui + fi * ur);
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// the 8-point stencil.  One thread per output (4d site, s); consecutive threads walk s then the 4d site,
// so the Ls threads of one 4d site broadcast-share its 8 links and spinor reads are 16-byte coalesced.
// ----------------------------------------------------------------------------------------------------
// [i4][8][9 complex] -> [i4][8][rows 0, 1, f]: f = <d, row 2> / <d, d> with d = conj(row 0 x row 1) (exact for a multiple of an
// SU(3) matrix, which is what the stencil's link tables hold)
template <typename T>
__global__ void k_compress_links(size_t n, const T* __restrict__ links, T* __restrict__ links_c) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // (site, direction)
  if (i >= n) return;
  const T* W = links + i * 18;
  T* o = links_c + i * 14;
  double w[18];
  for (int k = 0; k < 18; k++) w[k] = W[k];
  double nr = 0, ni = 0, dd = 0;
  for (int c = 0; c < 3; c++) {
    const int a = (c + 1) % 3, b = (c + 2) % 3;
    double xr = w[2 * a] * w[6 + 2 * b] - w[2 * a + 1] * w[6 + 2 * b + 1] - (w[2 * b] * w[6 + 2 * a] - w[2 * b + 1] * w[6 + 2 * a + 1]);
    double xi = w[2 * a] * w[6 + 2 * b + 1] + w[2 * a + 1] * w[6 + 2 * b] - (w[2 * b] * w[6 + 2 * a + 1] + w[2 * b + 1] * w[6 + 2 * a]);
    // d = conj(x) = (xr, -xi); conj(d) row2 = (xr + i xi)(r + i s)
    nr += xr * w[12 + 2 * c] - xi * w[12 + 2 * c + 1];
    ni += xr * w[12 + 2 * c + 1] + xi * w[12 + 2 * c];
    dd += xr * xr + xi * xi;
  }
  for (int k = 0; k < 12; k++) o[k] = W[k];
  o[12] = dd > 0 ? (T)(nr / dd) : (T)0;
  o[13] = dd > 0 ? (T)(ni / dd) : (T)0;
}

template <typename T, bool DAG, bool CMP = false>
__global__ void __launch_bounds__(128) k_dhop(Geom g, int ls, int p_out, const T* __restrict__ in, size_t in_stride,
                                             T* __restrict__ out, size_t out_stride, const T* __restrict__ links) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (size_t)g.half4 * ls) return;
  int i4 = (int)(tid / ls);
  int s = (int)(tid - (size_t)i4 * ls);
  int x, y, z, t;
  cb_coords(g, p_out, i4, x, y, z, t);
  T acc[24];
#pragma unroll
  for (int k = 0; k < 24; k++) acc[k] = 0;
  hop<0, true, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop<0, false, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop<1, true, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop<1, false, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop<2, true, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop<2, false, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop<3, true, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop<3, false, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  store_spinor(out, out_stride, tid, acc);
}

// The same stencil with one WARP per hop direction: a CTA of 8 warps handles 32 consecutive output sites, warp d computes the
// hop in direction d for all of them (no divergence: the direction is uniform in a warp; neighbour spinors of consecutive sites
// are consecutive), the eight partial spinors meet in shared memory and 3 (fp32) / 6 (fp64) x 32 threads add them up and store
// aligned 32-byte blocks.  Eight times the threads of k_dhop and eight independent loads in flight per site: what a small or
// single-rhs lattice (Wilson Ls = 1: 32 k sites per parity at 16^4) needs to hide the latency of its loads.
template <typename T, bool DAG, bool CMP>
__global__ void __launch_bounds__(256) k_dhop_dir(Geom g, int ls, int p_out, const T* __restrict__ in, size_t in_stride,
                                                  T* __restrict__ out, size_t out_stride, const T* __restrict__ links) {
  __shared__ T part[8 * 24 * 32];  // [direction][component][site of the CTA]
  const int lane = threadIdx.x & 31, d = threadIdx.x >> 5;
  const size_t nsite = (size_t)g.half4 * ls;
  const size_t tid = (size_t)blockIdx.x * 32 + lane;
  T acc[24];
#pragma unroll
  for (int k = 0; k < 24; k++) acc[k] = 0;
  if (tid < nsite) {
    const int i4 = (int)(tid / ls);
    const int s = (int)(tid - (size_t)i4 * ls);
    int x, y, z, t;
    cb_coords(g, p_out, i4, x, y, z, t);
    switch (d) {  // uniform in the warp
      case 0: hop<0, true, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links); break;
      case 1: hop<0, false, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links); break;
      case 2: hop<1, true, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links); break;
      case 3: hop<1, false, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links); break;
      case 4: hop<2, true, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links); break;
      case 5: hop<2, false, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links); break;
      case 6: hop<3, true, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links); break;
      default: hop<3, false, DAG, CMP>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links); break;
    }
  }
#pragma unroll
  for (int k = 0; k < 24; k++) part[(d * 24 + k) * 32 + lane] = acc[k];
  __syncthreads();
  constexpr int RB = 32 / sizeof(T);      // reals per 32-byte block: 8 / 4
  constexpr int NBLK = 24 / RB;           // blocks per spinor: 3 / 6
  if (threadIdx.x < NBLK * 32 && tid < nsite) {
    const int kb = threadIdx.x >> 5;      // (kb, lane): warps 0 .. NBLK-1
    T v[RB];
#pragma unroll
    for (int e = 0; e < RB; e++) {
      T sum = part[(0 * 24 + kb * RB + e) * 32 + lane];
#pragma unroll
      for (int dd = 1; dd < 8; dd++) sum += part[(dd * 24 + kb * RB + e) * 32 + lane];
      v[e] = sum;
    }
    st256(out + ((size_t)kb * out_stride + tid) * RB, v);
  }
}

void dhop_half_f32(cgptb_fermion_operator* op, bool dag, const float* pin, size_t in_stride, float* pout, size_t out_stride,
                   int p_out);  // dslash_f32.cu

static bool use_generic_dhop() {
  static int v = -1;
  if (v < 0) v = getenv("CGPTB_GENERIC_DHOP") ? 1 : 0;
  return v == 1;
}

// out(parity p_out) = Dhop in(parity 1-p_out); in/out given as (lattice, parity half to use)
template <typename T>
static void dhop_half(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out, int p_out) {
  int ls = op->ls();
  size_t half = (size_t)op->g.half4 * ls;
  const T* pin = (const T*)in->data;
  T* pout = (T*)out->data;
  const size_t blk_reals = 32 / sizeof(T);
  if (in->cb == CGPTB_FULL) pin += (size_t)(1 - p_out) * half * blk_reals;
  if (out->cb == CGPTB_FULL) pout += (size_t)p_out * half * blk_reals;
  // CGPTB_HALO_TIMING=1: device time of the three phases of a decomposed Dslash (pack + signal | interior | wait + exterior),
  // accumulated over the calls and printed every 200 calls -- a measurement aid, it synchronises the stream
  static int timing = -1;
  if (timing < 0) timing = getenv("CGPTB_HALO_TIMING") ? 1 : 0;
  static cudaEvent_t tev[4];
  static double tacc[3] = {0, 0, 0};
  static int tcalls = 0;
  const bool timed = timing == 1 && op->g.comm_mask;
  if (timed && tcalls == 0 && tacc[0] == 0)
    for (int i = 0; i < 4; i++) cudaEventCreate(&tev[i]);
  if (timed) cudaEventRecord(tev[0], g_stream);
  // split lattice: faces go out first, the interior stencil hides the transfer, then the boundary update
  if (op->g.comm_mask) halo_begin(op, dag, p_out, pin, in->sites);
  if (timed) cudaEventRecord(tev[1], g_stream);
  // two-row link compression: the TMA sweep kernel (single precision) and the generic kernel have a compressed instance
  static int dir_env = getenv("CGPTB_DHOP_DIR") ? atoi(getenv("CGPTB_DHOP_DIR")) : -1;
  // one warp per hop direction (k_dhop_dir): measured SLOWER than one thread per site for single-rhs Wilson-clover at 16^4 and at
  // 32^3 x 64 in both precisions (profiles/ablation_r2.txt), so it only runs on request (CGPTB_DHOP_DIR=1)
  const bool use_dir = dir_env == 1 && !dhop_tma_usable(op);
  if (sizeof(T) == 4 && !use_generic_dhop() && !use_dir && !(op->compress && !dhop_tma_usable(op))) {
    dhop_half_f32(op, dag, (const float*)pin, in->sites, (float*)pout, out->sites, p_out);
  } else {
    int threads = 128;
    unsigned blocks = (unsigned)((half + threads - 1) / threads);
    if (use_dir) {
      const unsigned db = (unsigned)((half + 31) / 32);
      const T* lk = (const T*)(op->compress ? op->links_c[p_out] : op->links[p_out]);
#define DIR_LAUNCH(DAG_, CMP_) k_dhop_dir<T, DAG_, CMP_><<<db, 256, 0, g_stream>>>(op->g, ls, p_out, pin, in->sites, pout, out->sites, lk)
      if (op->compress) {
        if (dag)
          DIR_LAUNCH(true, true);
        else
          DIR_LAUNCH(false, true);
      } else if (dag)
        DIR_LAUNCH(true, false);
      else
        DIR_LAUNCH(false, false);
#undef DIR_LAUNCH
    } else if (op->compress) {
      if (dag)
        k_dhop<T, true, true><<<blocks, threads, 0, g_stream>>>(op->g, ls, p_out, pin, in->sites, pout, out->sites, (const T*)op->links_c[p_out]);
      else
        k_dhop<T, false, true><<<blocks, threads, 0, g_stream>>>(op->g, ls, p_out, pin, in->sites, pout, out->sites, (const T*)op->links_c[p_out]);
    } else if (dag)
      k_dhop<T, true><<<blocks, threads, 0, g_stream>>>(op->g, ls, p_out, pin, in->sites, pout, out->sites, (const T*)op->links[p_out]);
    else
      k_dhop<T, false><<<blocks, threads, 0, g_stream>>>(op->g, ls, p_out, pin, in->sites, pout, out->sites, (const T*)op->links[p_out]);
    LAUNCH_CHECK();
  }
  if (timed) cudaEventRecord(tev[2], g_stream);
  if (op->g.comm_mask) halo_end(op, dag, p_out, pout, out->sites);
  if (timed) {
    cudaEventRecord(tev[3], g_stream);
    cudaEventSynchronize(tev[3]);
    for (int i = 0; i < 3; i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, tev[i], tev[i + 1]);
      tacc[i] += ms;
    }
    if (++tcalls % 200 == 0) {
      fprintf(stderr, "[halo timing rank %d] per half Dslash: pack+signal %.1f us, interior %.1f us, wait+exterior %.1f us (%d calls)\n", g_comm.rank,
              1e3 * tacc[0] / tcalls, 1e3 * tacc[1] / tcalls, 1e3 * tacc[2] / tcalls, tcalls);
    }
  }
}

// results vanish on the global time slices 0 and T-1 (lib/gpt/qcd/fermion/boundary_conditions.py:23-31); in the device layout
// a time slice of one parity is contiguous in every component plane
void apply_open_boundaries(const cgptb_fermion_operator* op, cgptb_lattice* l) {
  if (!op->open_bc) return;
  const Geom& g = op->g;
  const int* goff = op->goff;
  const int* gL = op->gL;
  const int* dims4 = op->dims4;
  const size_t rs = l->real_size(), blk = 32 / rs;
  const size_t slice = (size_t)g.hx * g.L[1] * g.L[2] * op->ls();  // blocks of one parity per time slice
  const size_t half = slice * g.L[3];
  const int nplanes = (int)(24 / blk);
  const int nhalves = l->cb == CGPTB_FULL ? 2 : 1;
  for (int side = 0; side < 2; side++) {
    if (side == 0 && goff[3] != 0) continue;
    if (side == 1 && goff[3] + dims4[3] != gL[3]) continue;
    const size_t t = side == 0 ? 0 : (size_t)g.L[3] - 1;
    for (int k = 0; k < nplanes; k++)
      for (int h = 0; h < nhalves; h++) {
        char* p = (char*)l->data + ((size_t)k * l->sites + (size_t)h * half + t * slice) * 32;
        CUDA_CHECK(cudaMemsetAsync(p, 0, slice * 32, g_stream));
      }
  }
}

void op_dhop(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out) {
  op->check_field(in);
  op->check_field(out);
  CGPTB_ASSERT(in->data != out->data);
  if (in->cb == CGPTB_FULL) {
    CGPTB_ASSERT(out->cb == CGPTB_FULL);
    for (int p = 0; p < 2; p++) {
      if (op->prec == CGPTB_SINGLE)
        dhop_half<float>(op, dag, in, out, p);
      else
        dhop_half<double>(op, dag, in, out, p);
    }
  } else {
    CGPTB_ASSERT(out->cb != CGPTB_FULL);
    out->cb = 1 - in->cb;
    if (op->prec == CGPTB_SINGLE)
      dhop_half<float>(op, dag, in, out, out->cb);
    else
      dhop_half<double>(op, dag, in, out, out->cb);
  }
  apply_open_boundaries(op, out);
}

// ----------------------------------------------------------------------------------------------------
// s-direction operators of the Moebius action (M5D shape: lib/cgpt/lib/foundation/mobius_with_vector_field.h:36-96)
//   out_s (+)= d psi_s + P+ (lp[s] psi_{s-1} + up[s] psi_{s+1}) + P- (lm[s] psi_{s-1} + um[s] psi_{s+1}), cyclic in s
// coef table: [5][ls] = d, lp, up, lm, um
// ----------------------------------------------------------------------------------------------------
template <typename T, bool ACC>
__global__ void __launch_bounds__(128) k_s_tridiag(size_t n, int ls, const T* __restrict__ in, T* __restrict__ out,
                                                   size_t stride, const T* __restrict__ coef) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  size_t i4 = tid / ls;
  int s = (int)(tid - i4 * ls);
  int sm = s == 0 ? ls - 1 : s - 1;
  int sp = s == ls - 1 ? 0 : s + 1;
  T c[24], a[24], b[24], r[24];
  load_spinor(in, stride, tid, c);
  load_spinor(in, stride, i4 * ls + sm, a);
  load_spinor(in, stride, i4 * ls + sp, b);
  T d = coef[s], lp = coef[ls + s], up = coef[2 * ls + s], lm = coef[3 * ls + s], um = coef[4 * ls + s];
  if (ACC) load_spinor_rw(out, stride, tid, r);
#pragma unroll
  for (int k = 0; k < 12; k++) {
    T v = d * c[k] + lp * a[k] + up * b[k];
    r[k] = ACC ? r[k] + v : v;
  }
#pragma unroll
  for (int k = 12; k < 24; k++) {
    T v = d * c[k] + lm * a[k] + um * b[k];
    r[k] = ACC ? r[k] + v : v;
  }
  store_spinor(out, stride, tid, r);
}

// dense Ls x Ls chirality blocks (MooeeInv): out_s = sum_s' Mp[s][s'] P+ psi_s' + Mm[s][s'] P- psi_s'
template <typename T>
__global__ void __launch_bounds__(128) k_s_dense(size_t n, int ls, const T* __restrict__ in, T* __restrict__ out,
                                                 size_t stride, const T* __restrict__ Mp, const T* __restrict__ Mm) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  size_t i4 = tid / ls;
  int s = (int)(tid - i4 * ls);
  T r[24];
#pragma unroll
  for (int k = 0; k < 24; k++) r[k] = 0;
  for (int sp = 0; sp < ls; sp++) {
    T v[24];
    load_spinor(in, stride, i4 * ls + sp, v);
    T mp = Mp[s * ls + sp], mm = Mm[s * ls + sp];
#pragma unroll
    for (int k = 0; k < 12; k++) r[k] += mp * v[k];
#pragma unroll
    for (int k = 12; k < 24; k++) r[k] += mm * v[k];
  }
  store_spinor(out, stride, tid, r);
}

// zMoebius: dense complex Ls x Ls chirality blocks, out_s (+)= sum_s' Mp[s][s'] P+ psi_s' + Mm[s][s'] P- psi_s'.  Every
// s-operator of the zMoebius action goes through this one kernel (the tridiagonal structure is not exploited yet).
template <typename T, bool ACC>
__global__ void __launch_bounds__(128) k_s_dense_c(size_t n, int ls, const T* __restrict__ in, T* __restrict__ out, size_t stride,
                                                   const T* __restrict__ Mp, const T* __restrict__ Mm) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  size_t i4 = tid / ls;
  int s = (int)(tid - i4 * ls);
  T r[24];
  if (ACC)
    load_spinor_rw(out, stride, tid, r);
  else {
#pragma unroll
    for (int k = 0; k < 24; k++) r[k] = 0;
  }
  for (int sp = 0; sp < ls; sp++) {
    const T pr = Mp[2 * (s * ls + sp)], pi = Mp[2 * (s * ls + sp) + 1];
    const T mr = Mm[2 * (s * ls + sp)], mi = Mm[2 * (s * ls + sp) + 1];
    if (pr == 0 && pi == 0 && mr == 0 && mi == 0) continue;  // uniform over the threads that share s
    T v[24];
    load_spinor(in, stride, i4 * ls + sp, v);
#pragma unroll
    for (int c = 0; c < 6; c++) {
      r[2 * c] += pr * v[2 * c] - pi * v[2 * c + 1];
      r[2 * c + 1] += pr * v[2 * c + 1] + pi * v[2 * c];
    }
#pragma unroll
    for (int c = 6; c < 12; c++) {
      r[2 * c] += mr * v[2 * c] - mi * v[2 * c + 1];
      r[2 * c + 1] += mr * v[2 * c + 1] + mi * v[2 * c];
    }
  }
  store_spinor(out, stride, tid, r);
}

void op_s_z(cgptb_fermion_operator* op, int zkind, bool dag, bool acc, const cgptb_lattice* in, cgptb_lattice* out) {
  op->check_field(in);
  op->check_field(out);
  CGPTB_ASSERT(op->zmobius && op->z_tab && in->sites == out->sites && in->data != out->data);
  if (!acc) out->cb = in->cb;
  const int ls = op->Ls;
  const size_t n = in->sites;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  const size_t blk = (size_t)ls * ls * 2;
  const size_t off = (size_t)(zkind * 2 + (dag ? 1 : 0)) * 2 * blk;
  if (op->prec == CGPTB_SINGLE) {
    const float* m = (const float*)op->z_tab + off;
    if (acc)
      k_s_dense_c<float, true><<<blocks, 128, 0, g_stream>>>(n, ls, (const float*)in->data, (float*)out->data, n, m, m + blk);
    else
      k_s_dense_c<float, false><<<blocks, 128, 0, g_stream>>>(n, ls, (const float*)in->data, (float*)out->data, n, m, m + blk);
  } else {
    const double* m = (const double*)op->z_tab + off;
    if (acc)
      k_s_dense_c<double, true><<<blocks, 128, 0, g_stream>>>(n, ls, (const double*)in->data, (double*)out->data, n, m, m + blk);
    else
      k_s_dense_c<double, false><<<blocks, 128, 0, g_stream>>>(n, ls, (const double*)in->data, (double*)out->data, n, m, m + blk);
  }
  LAUNCH_CHECK();
}

void op_s_tridiag(cgptb_fermion_operator* op, int kind, bool dag, bool acc, const cgptb_lattice* in, cgptb_lattice* out) {
  if (op->zmobius) {
    CGPTB_ASSERT(kind == 0 || kind == 2);
    op_s_z(op, kind == 0 ? ZK_A : ZK_EE, dag, acc, in, out);
    return;
  }
  op->check_field(in);
  op->check_field(out);
  CGPTB_ASSERT(in->sites == out->sites && in->data != out->data);
  if (!acc) out->cb = in->cb;
  int ls = op->Ls;
  size_t n = in->sites;
  int threads = 128;
  unsigned blocks = (unsigned)((n + threads - 1) / threads);
  size_t off = (size_t)(kind * 2 + (dag ? 1 : 0)) * 5 * ls;
  if (op->prec == CGPTB_SINGLE) {
    const float* coef = (const float*)op->s_coef + off;
    if (acc)
      k_s_tridiag<float, true><<<blocks, threads, 0, g_stream>>>(n, ls, (const float*)in->data, (float*)out->data, n, coef);
    else
      k_s_tridiag<float, false><<<blocks, threads, 0, g_stream>>>(n, ls, (const float*)in->data, (float*)out->data, n, coef);
  } else {
    const double* coef = (const double*)op->s_coef + off;
    if (acc)
      k_s_tridiag<double, true><<<blocks, threads, 0, g_stream>>>(n, ls, (const double*)in->data, (double*)out->data, n, coef);
    else
      k_s_tridiag<double, false><<<blocks, threads, 0, g_stream>>>(n, ls, (const double*)in->data, (double*)out->data, n, coef);
  }
  LAUNCH_CHECK();
}

void op_s_dense(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out) {
  if (op->zmobius) {
    op_s_z(op, ZK_EEINV, dag, false, in, out);
    return;
  }
  op->check_field(in);
  op->check_field(out);
  CGPTB_ASSERT(in->sites == out->sites && in->data != out->data);
  out->cb = in->cb;
  int ls = op->Ls;
  size_t n = in->sites;
  int threads = 128;
  unsigned blocks = (unsigned)((n + threads - 1) / threads);
  size_t off = (size_t)(dag ? 2 : 0) * ls * ls;
  if (op->prec == CGPTB_SINGLE) {
    const float* m = (const float*)op->s_inv + off;
    k_s_dense<float><<<blocks, threads, 0, g_stream>>>(n, ls, (const float*)in->data, (float*)out->data, n, m, m + ls * ls);
  } else {
    const double* m = (const double*)op->s_inv + off;
    k_s_dense<double><<<blocks, threads, 0, g_stream>>>(n, ls, (const double*)in->data, (double*)out->data, n, m, m + ls * ls);
  }
  LAUNCH_CHECK();
}

// 4d <-> 5d maps (Grid CayleyFermion5D Import/Export)
//   import (unphysical): out_0 = P+ src, out_{Ls-1} = P- src, rest 0
//   export solution: P- psi_0 + P+ psi_{Ls-1} ; export source: P+ psi_0 + P- psi_{Ls-1}
template <typename T>
__global__ void k_import5(size_t n4, int ls, const T* __restrict__ in4, T* __restrict__ out5) {
  size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= n4) return;
  T v[24], z[24], a[24], b[24];
  load_spinor(in4, n4, i4, v);
#pragma unroll
  for (int k = 0; k < 24; k++) {
    z[k] = 0;
    a[k] = k < 12 ? v[k] : (T)0;
    b[k] = k < 12 ? (T)0 : v[k];
  }
  for (int s = 0; s < ls; s++) {
    if (s == 0 && s == ls - 1) {
      store_spinor(out5, n4 * ls, i4 * ls + s, v);
    } else if (s == 0)
      store_spinor(out5, n4 * ls, i4 * ls + s, a);
    else if (s == ls - 1)
      store_spinor(out5, n4 * ls, i4 * ls + s, b);
    else
      store_spinor(out5, n4 * ls, i4 * ls + s, z);
  }
}

template <typename T, bool SOLUTION>
__global__ void k_export5(size_t n4, int ls, const T* __restrict__ in5, T* __restrict__ out4) {
  size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= n4) return;
  T a[24], b[24], r[24];
  load_spinor(in5, n4 * ls, i4 * ls, a);
  load_spinor(in5, n4 * ls, i4 * ls + ls - 1, b);
#pragma unroll
  for (int k = 0; k < 24; k++) {
    bool upper = k < 12;
    r[k] = SOLUTION ? (upper ? b[k] : a[k]) : (upper ? a[k] : b[k]);
  }
  store_spinor(out4, n4, i4, r);
}

// ----------------------------------------------------------------------------------------------------
// clover term: two Hermitian 6x6 chirality blocks per site, compact storage 2 x (6 real diag + 15 complex
// lower triangle) = 72 reals / site (Grid CompactWilsonClover), SoA over the sites of one parity.
// (lib/gpt/qcd/fermion/reference/wilson_clover.py:92-139,202-220)
// ----------------------------------------------------------------------------------------------------
template <typename T, bool ACC>
__global__ void __launch_bounds__(128) k_clover_apply(size_t n, int ls, const T* __restrict__ in, size_t in_stride, T* __restrict__ out,
                                                      size_t out_stride, const T* __restrict__ clov) {
  // n: 4d sites of this parity; ls > 1: multi-rhs field, the ls right-hand sides of a site share its clover blocks
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n * ls) return;
  const size_t i = j / ls;
  T psi[24], r[24];
  load_spinor(in, in_stride, j, psi);
  if (ACC) load_spinor_rw(out, out_stride, j, r);
#pragma unroll
  for (int blk = 0; blk < 2; blk++) {
    T o[12];
    // diagonal
#pragma unroll
    for (int a = 0; a < 6; a++) {
      T d = clov[(size_t)(blk * 36 + a) * n + i];
      o[2 * a] = d * psi[blk * 12 + 2 * a];
      o[2 * a + 1] = d * psi[blk * 12 + 2 * a + 1];
    }
    int e = 0;
#pragma unroll
    for (int a = 1; a < 6; a++) {
#pragma unroll
      for (int b = 0; b < a; b++) {
        T lr = clov[(size_t)(blk * 36 + 6 + 2 * e) * n + i];
        T li = clov[(size_t)(blk * 36 + 6 + 2 * e + 1) * n + i];
        e++;
        T pr = psi[blk * 12 + 2 * b], pi = psi[blk * 12 + 2 * b + 1];
        o[2 * a] += lr * pr - li * pi;
        o[2 * a + 1] += lr * pi + li * pr;
        pr = psi[blk * 12 + 2 * a];
        pi = psi[blk * 12 + 2 * a + 1];
        o[2 * b] += lr * pr + li * pi;
        o[2 * b + 1] += lr * pi - li * pr;
      }
    }
#pragma unroll
    for (int k = 0; k < 12; k++) r[blk * 12 + k] = ACC ? r[blk * 12 + k] + o[k] : o[k];
  }
  store_spinor(out, out_stride, j, r);
}

struct M3 {
  double re[9], im[9];
};
__device__ inline M3 m3_mul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double sr = 0, si = 0;
      for (int k = 0; k < 3; k++) {
        sr += a.re[i * 3 + k] * b.re[k * 3 + j] - a.im[i * 3 + k] * b.im[k * 3 + j];
        si += a.re[i * 3 + k] * b.im[k * 3 + j] + a.im[i * 3 + k] * b.re[k * 3 + j];
      }
      r.re[i * 3 + j] = sr;
      r.im[i * 3 + j] = si;
    }
  return r;
}
__device__ inline M3 m3_adj(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      r.re[i * 3 + j] = a.re[j * 3 + i];
      r.im[i * 3 + j] = -a.im[j * 3 + i];
    }
  return r;
}
__device__ inline M3 m3_axpy(double s, const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 9; i++) {
    r.re[i] = s * a.re[i] + b.re[i];
    r.im[i] = s * a.im[i] + b.im[i];
  }
  return r;
}

template <typename TU>
__device__ inline M3 load_U(const Geom& g, size_t nsites, const TU* U, int x, int y, int z, int t) {
  x = (x + g.L[0]) % g.L[0];
  y = (y + g.L[1]) % g.L[1];
  z = (z + g.L[2]) % g.L[2];
  t = (t + g.L[3]) % g.L[3];
  int p = (x + y + z + t) & 1;
  size_t site = (size_t)p * g.half4 + cb_index(g, x, y, z, t);
  M3 r;
  for (int k = 0; k < 9; k++) {
    size_t o = elem_offset<TU>(nsites, site, k, 1);
    r.re[k] = U[o];
    r.im[k] = U[o + 1];
  }
  return r;
}

// v(x) = staple_up - staple_down of lib/gpt/qcd/gauge/loops.py:156-166 at site c
template <typename TU>
__device__ inline M3 clover_v(const Geom& g, size_t ns, const TU* Umu, const TU* Unu, int mu, int nu, const int c[4]) {
  int e[4];
  // staple_up = U_nu(x+mu) U_mu^dag(x+nu) U_nu^dag(x)
  for (int i = 0; i < 4; i++) e[i] = c[i];
  e[mu]++;
  M3 a = load_U(g, ns, Unu, e[0], e[1], e[2], e[3]);
  e[mu]--;
  e[nu]++;
  M3 b = m3_adj(load_U(g, ns, Umu, e[0], e[1], e[2], e[3]));
  e[nu]--;
  M3 d = m3_adj(load_U(g, ns, Unu, e[0], e[1], e[2], e[3]));
  M3 up = m3_mul(m3_mul(a, b), d);
  // staple_down = U_nu^dag(x+mu-nu) U_mu^dag(x-nu) U_nu(x-nu)
  e[mu]++;
  e[nu]--;
  a = m3_adj(load_U(g, ns, Unu, e[0], e[1], e[2], e[3]));
  e[mu]--;
  b = m3_adj(load_U(g, ns, Umu, e[0], e[1], e[2], e[3]));
  d = load_U(g, ns, Unu, e[0], e[1], e[2], e[3]);
  M3 dn = m3_mul(m3_mul(a, b), d);
  return m3_axpy(-1.0, dn, up);
}

struct CloverCoef {
  int open_bc, T;      // open boundary conditions: clover term -> -csw_t/2 on t = 0, T-1; += cF - 1 on t = 1, T-2
  double edge, cF1;    // -csw_t / 2, cF - 1     (lib/gpt/qcd/fermion/reference/wilson_clover.py:106-125)
  double diag;
  double c[6];         // -1/2 * c_{mu nu} for planes (0,1),(0,2),(0,3),(1,2),(1,3),(2,3)
  double sre[6][16];   // sigma_{mu nu} as 4x4 complex
  double sim[6][16];
};

// in-place inverse of a 6x6 complex matrix (Gauss-Jordan, partial pivoting)
__device__ inline void inv6(double (&ar)[36], double (&ai)[36]) {
  double br[36], bi[36];
  for (int i = 0; i < 36; i++) {
    br[i] = (i / 6 == i % 6) ? 1.0 : 0.0;
    bi[i] = 0.0;
  }
  for (int col = 0; col < 6; col++) {
    int piv = col;
    double best = ar[col * 6 + col] * ar[col * 6 + col] + ai[col * 6 + col] * ai[col * 6 + col];
    for (int r = col + 1; r < 6; r++) {
      double m = ar[r * 6 + col] * ar[r * 6 + col] + ai[r * 6 + col] * ai[r * 6 + col];
      if (m > best) {
        best = m;
        piv = r;
      }
    }
    if (piv != col)
      for (int k = 0; k < 6; k++) {
        double t;
        t = ar[col * 6 + k]; ar[col * 6 + k] = ar[piv * 6 + k]; ar[piv * 6 + k] = t;
        t = ai[col * 6 + k]; ai[col * 6 + k] = ai[piv * 6 + k]; ai[piv * 6 + k] = t;
        t = br[col * 6 + k]; br[col * 6 + k] = br[piv * 6 + k]; br[piv * 6 + k] = t;
        t = bi[col * 6 + k]; bi[col * 6 + k] = bi[piv * 6 + k]; bi[piv * 6 + k] = t;
      }
    double pr = ar[col * 6 + col], pi = ai[col * 6 + col];
    double nn = pr * pr + pi * pi;
    double ir = pr / nn, ii = -pi / nn;
    for (int k = 0; k < 6; k++) {
      double xr = ar[col * 6 + k], xi = ai[col * 6 + k];
      ar[col * 6 + k] = xr * ir - xi * ii;
      ai[col * 6 + k] = xr * ii + xi * ir;
      xr = br[col * 6 + k];
      xi = bi[col * 6 + k];
      br[col * 6 + k] = xr * ir - xi * ii;
      bi[col * 6 + k] = xr * ii + xi * ir;
    }
    for (int r = 0; r < 6; r++) {
      if (r == col) continue;
      double fr = ar[r * 6 + col], fi = ai[r * 6 + col];
      for (int k = 0; k < 6; k++) {
        ar[r * 6 + k] -= fr * ar[col * 6 + k] - fi * ai[col * 6 + k];
        ai[r * 6 + k] -= fr * ai[col * 6 + k] + fi * ar[col * 6 + k];
        br[r * 6 + k] -= fr * br[col * 6 + k] - fi * bi[col * 6 + k];
        bi[r * 6 + k] -= fr * bi[col * 6 + k] + fi * br[col * 6 + k];
      }
    }
  }
  for (int i = 0; i < 36; i++) {
    ar[i] = br[i];
    ai[i] = bi[i];
  }
}

template <typename T>
__device__ inline void store_compact(T* dst, size_t n, size_t i, int blk, const double (&ar)[36], const double (&ai)[36]) {
  for (int a = 0; a < 6; a++) dst[(size_t)(blk * 36 + a) * n + i] = (T)ar[a * 6 + a];
  int e = 0;
  for (int a = 1; a < 6; a++)
    for (int b = 0; b < a; b++) {
      dst[(size_t)(blk * 36 + 6 + 2 * e) * n + i] = (T)ar[a * 6 + b];
      dst[(size_t)(blk * 36 + 6 + 2 * e + 1) * n + i] = (T)ai[a * 6 + b];
      e++;
    }
}

// one thread per site of parity p: field strength of the six planes from the un-phased links
// (reference/wilson_clover.py:96-104), the two 6x6 blocks and their inverses, always in double.
template <typename TU, typename T>
__global__ void __launch_bounds__(64) k_build_clover(Geom gl, Geom g, int off0, int off1, int off2, int off3, int p, size_t nsU,
                                                     const TU* U0, const TU* U1, const TU* U2, const TU* U3, CloverCoef cc,
                                                     T* __restrict__ clov, T* __restrict__ clov_inv) {
  // gl: the local lattice the blocks are stored for; g, nsU, U*: the lattice the links live on -- the same, or the global
  // lattice of a decomposed run (the field strength reaches over the corners of the local volume), off = local origin in it
  int i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= gl.half4) return;
  int c[4];
  cb_coords(gl, p, i4, c[0], c[1], c[2], c[3]);
  c[0] += off0;
  c[1] += off1;
  c[2] += off2;
  c[3] += off3;
  const TU* U[4] = {U0, U1, U2, U3};
  double br[2][36], bi[2][36];
  for (int b = 0; b < 2; b++)
    for (int k = 0; k < 36; k++) {
      br[b][k] = (k / 6 == k % 6) ? cc.diag : 0.0;
      bi[b][k] = 0.0;
    }
  int plane = 0;
  for (int mu = 0; mu < 4; mu++)
    for (int nu = mu + 1; nu < 4; nu++, plane++) {
      // F = U_mu(x) v(x) + v(x-mu) U_mu(x-mu) ; F = 1/8 (F - F^dag)    (loops.py:168-174)
      M3 v = clover_v(g, nsU, U[mu], U[nu], mu, nu, c);
      M3 F = m3_mul(load_U(g, nsU, U[mu], c[0], c[1], c[2], c[3]), v);
      int e[4] = {c[0], c[1], c[2], c[3]};
      e[mu]--;
      M3 v2 = clover_v(g, nsU, U[mu], U[nu], mu, nu, e);
      M3 F2 = m3_mul(v2, load_U(g, nsU, U[mu], e[0], e[1], e[2], e[3]));
      F = m3_axpy(1.0, F, F2);
      M3 Fd = m3_adj(F);
      F = m3_axpy(-1.0, Fd, F);
      double coef = cc.c[plane] * 0.125;
      for (int b = 0; b < 2; b++)
        for (int s1 = 0; s1 < 2; s1++)
          for (int s2 = 0; s2 < 2; s2++) {
            double sr = cc.sre[plane][(2 * b + s1) * 4 + 2 * b + s2], si = cc.sim[plane][(2 * b + s1) * 4 + 2 * b + s2];
            if (sr == 0.0 && si == 0.0) continue;
            for (int a1 = 0; a1 < 3; a1++)
              for (int a2 = 0; a2 < 3; a2++) {
                double fr = F.re[a1 * 3 + a2], fi = F.im[a1 * 3 + a2];
                br[b][(s1 * 3 + a1) * 6 + s2 * 3 + a2] += coef * (sr * fr - si * fi);
                bi[b][(s1 * 3 + a1) * 6 + s2 * 3 + a2] += coef * (sr * fi + si * fr);
              }
          }
    }
  if (cc.open_bc) {
    const int t = c[3];
    for (int b = 0; b < 2; b++)
      for (int k = 0; k < 36; k++) {
        const bool dg = k / 6 == k % 6;
        if (t == 0 || t == cc.T - 1) {
          br[b][k] = dg ? cc.diag + cc.edge : 0.0;
          bi[b][k] = 0.0;
        } else if ((t == 1 || t == cc.T - 2) && dg)
          br[b][k] += cc.cF1;
      }
  }
  for (int b = 0; b < 2; b++) {
    store_compact<T>(clov, gl.half4, i4, b, br[b], bi[b]);
    inv6(br[b], bi[b]);
    store_compact<T>(clov_inv, gl.half4, i4, b, br[b], bi[b]);
  }
}

}  // namespace cgptb

using namespace cgptb;

// ----------------------------------------------------------------------------------------------------
// host side of the operator object
// ----------------------------------------------------------------------------------------------------
void cgptb_fermion_operator::check_field(const cgptb_lattice* l) const {
  if (l->prec != prec) CGPTB_ERR("fermion operator is %s precision, field is not", prec == CGPTB_SINGLE ? "single" : "double");
  if (l->otype != CGPTB_OT_VSPINCOLOR) CGPTB_ERR("fermion operator needs spin-colour vector fields");
  for (int i = 0; i < 4; i++)
    if (l->dims4[i] != dims4[i]) CGPTB_ERR("field lives on a different grid than the operator");
  if (l->Ls != Ls) CGPTB_ERR("field has Ls=%d, operator has Ls=%d", l->Ls, Ls);
}

cgptb_lattice* cgptb_fermion_operator::tmp(int i, int cb) {
  CGPTB_ASSERT(i >= 0 && i < 4);
  int want_full = cb == CGPTB_FULL;
  cgptb_lattice*& t = want_full ? tmp_full[i] : tmp_half[i];
  if (!t) {
    if (cgptb_create_lattice(&t, dims4, Ls, prec, CGPTB_OT_VSPINCOLOR, want_full ? CGPTB_FULL : CGPTB_EVEN))
      CGPTB_ERR("%s", cgptb_last_error());
  }
  if (!want_full) t->cb = cb;
  return t;
}

static void invert_dense(int n, std::vector<double>& a) {
  std::vector<double> b(n * n, 0.0);
  for (int i = 0; i < n; i++) b[i * n + i] = 1.0;
  for (int col = 0; col < n; col++) {
    int piv = col;
    for (int r = col + 1; r < n; r++)
      if (fabs(a[r * n + col]) > fabs(a[piv * n + col])) piv = r;
    if (a[piv * n + col] == 0.0) CGPTB_ERR("Mooee is singular");
    for (int k = 0; k < n; k++) {
      std::swap(a[col * n + k], a[piv * n + k]);
      std::swap(b[col * n + k], b[piv * n + k]);
    }
    double inv = 1.0 / a[col * n + col];
    for (int k = 0; k < n; k++) {
      a[col * n + k] *= inv;
      b[col * n + k] *= inv;
    }
    for (int r = 0; r < n; r++) {
      if (r == col) continue;
      double f = a[r * n + col];
      if (f == 0.0) continue;
      for (int k = 0; k < n; k++) {
        a[r * n + k] -= f * a[col * n + k];
        b[r * n + k] -= f * b[col * n + k];
      }
    }
  }
  a = b;
}

template <typename T>
static void upload(void** dev, const std::vector<double>& h) {
  std::vector<T> t(h.size());
  for (size_t i = 0; i < h.size(); i++) t[i] = (T)h[i];
  if (!*dev) CUDA_CHECK(cudaMalloc(dev, t.size() * sizeof(T)));
  CUDA_CHECK(cudaMemcpyAsync(*dev, t.data(), t.size() * sizeof(T), cudaMemcpyHostToDevice, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
}

// coefficient tables of the s-direction operators; S5 as in SURVEY Appendix A.2 / tests/qcd/domain_wall.py:309-341
// zMoebius: all s-operators as dense complex chirality blocks, built like the oracle does (oracle/qcd.py zmobius._AB)
static void setup_zmobius_tables(cgptb_fermion_operator* op) {
  typedef std::complex<double> cd;
  const int ls = op->Ls;
  const cgptb_fermion_params& p = op->p;
  std::vector<cd> bs(ls), cs(ls);
  for (int s = 0; s < ls; s++) {
    const cd om(p.omega[2 * s], p.omega[2 * s + 1]);
    bs[s] = 0.5 * ((p.b + p.c) / om + (p.b - p.c));  // lib/gpt/qcd/fermion/zmobius.py:33
    cs[s] = 0.5 * ((p.b + p.c) / om - (p.b - p.c));
  }
  const size_t nn = (size_t)ls * ls;
  auto S5 = [&](int chir) {  // chir 0: P+ block, 1: P- block
    std::vector<cd> S(nn, cd(0));
    for (int s = 0; s < ls; s++) {
      if (chir == 0 && s >= 1) S[s * ls + s - 1] = 1.0;
      if (chir == 1 && s + 1 < ls) S[s * ls + s + 1] = 1.0;
    }
    if (chir == 0)
      S[0 * ls + ls - 1] += -p.mass_plus;
    else
      S[(ls - 1) * ls + 0] += -p.mass_minus;
    return S;
  };
  auto inverse = [&](std::vector<cd> a) {
    std::vector<cd> b(nn, cd(0));
    for (int i = 0; i < ls; i++) b[i * ls + i] = 1.0;
    for (int col = 0; col < ls; col++) {
      int piv = col;
      for (int r = col + 1; r < ls; r++)
        if (std::abs(a[r * ls + col]) > std::abs(a[piv * ls + col])) piv = r;
      if (std::abs(a[piv * ls + col]) == 0.0) CGPTB_ERR("zmobius: Mooee is singular");
      for (int k = 0; k < ls; k++) {
        std::swap(a[col * ls + k], a[piv * ls + k]);
        std::swap(b[col * ls + k], b[piv * ls + k]);
      }
      const cd d = 1.0 / a[col * ls + col];
      for (int k = 0; k < ls; k++) {
        a[col * ls + k] *= d;
        b[col * ls + k] *= d;
      }
      for (int r = 0; r < ls; r++) {
        if (r == col) continue;
        const cd f = a[r * ls + col];
        if (f == cd(0)) continue;
        for (int k = 0; k < ls; k++) {
          a[r * ls + k] -= f * a[col * ls + k];
          b[r * ls + k] -= f * b[col * ls + k];
        }
      }
    }
    return b;
  };
  std::vector<double> tab((size_t)ZK_COUNT * 2 * 2 * nn * 2, 0.0);
  for (int chir = 0; chir < 2; chir++) {
    const std::vector<cd> S = S5(chir);
    std::vector<cd> M[ZK_COUNT];
    for (int k = 0; k < ZK_COUNT; k++) M[k].assign(nn, cd(0));
    for (int s = 0; s < ls; s++)
      for (int t = 0; t < ls; t++) {
        const cd id = s == t ? 1.0 : 0.0;
        const cd A = bs[s] * id + cs[s] * S[s * ls + t];
        M[ZK_A][s * ls + t] = A;
        M[ZK_EE][s * ls + t] = (4.0 - p.M5) * A + id - S[s * ls + t];
        M[ZK_DM1][s * ls + t] = (1.0 - cs[s] * (4.0 - p.M5)) * id;
        M[ZK_DM2][s * ls + t] = -cs[s] * id;
      }
    M[ZK_EEINV] = inverse(M[ZK_EE]);
    for (int k = 0; k < ZK_COUNT; k++)
      for (int dag = 0; dag < 2; dag++) {
        double* o = &tab[((size_t)(k * 2 + dag) * 2 + chir) * nn * 2];
        for (int s = 0; s < ls; s++)
          for (int t = 0; t < ls; t++) {
            const cd v = dag ? std::conj(M[k][t * ls + s]) : M[k][s * ls + t];
            o[2 * (s * ls + t)] = v.real();
            o[2 * (s * ls + t) + 1] = v.imag();
          }
      }
  }
  if (op->prec == CGPTB_SINGLE)
    upload<float>(&op->z_tab, tab);
  else
    upload<double>(&op->z_tab, tab);
}

void cgptb_fermion_operator::setup_mobius_tables() {
  if (zmobius) {
    setup_zmobius_tables(this);
    return;
  }
  int ls = Ls;
  double b = p.b, c = p.c, mp = p.mass_plus, mm = p.mass_minus;
  double bee = b * (4.0 - p.M5) + 1.0, cee = 1.0 - c * (4.0 - p.M5);
  // kind: 0 = A = b + c S5 ; 1 = B = 1 - S5 ; 2 = ee = bee - cee S5 ; each non-dag / dag ; rows d, lp, up, lm, um
  std::vector<double> tab(3 * 2 * 5 * ls, 0.0);
  for (int kind = 0; kind < 3; kind++) {
    double d = kind == 0 ? b : (kind == 1 ? 1.0 : bee);
    double f = kind == 0 ? c : (kind == 1 ? -1.0 : -cee);  // coefficient of S5
    for (int dag = 0; dag < 2; dag++) {
      double* t = &tab[(size_t)(kind * 2 + dag) * 5 * ls];
      for (int s = 0; s < ls; s++) {
        t[s] = d;
        double* lpp = t + ls;
        double* upp = t + 2 * ls;
        double* lmm = t + 3 * ls;
        double* umm = t + 4 * ls;
        if (!dag) {
          // S5: P+ psi_{s-1} (s>=1), -m+ P+ psi_{Ls-1} (s=0) ; P- psi_{s+1} (s<Ls-1), -m- P- psi_0 (s=Ls-1)
          lpp[s] += f * (s == 0 ? -mp : 1.0);
          umm[s] += f * (s == ls - 1 ? -mm : 1.0);
        } else {
          // S5^dag: P+ psi_{s+1} (s<Ls-1), -m+ P+ psi_0 (s=Ls-1) ; P- psi_{s-1} (s>=1), -m- P- psi_{Ls-1} (s=0)
          upp[s] += f * (s == ls - 1 ? -mp : 1.0);
          lmm[s] += f * (s == 0 ? -mm : 1.0);
        }
      }
    }
  }
  // dense inverses of Mooee per chirality (exact; replaces Grid's LDU sweep,
  // lib/cgpt/lib/foundation/mobius_with_vector_field.h:185-302)
  std::vector<double> Ap(ls * ls, 0.0), Am(ls * ls, 0.0);
  for (int s = 0; s < ls; s++) {
    Ap[s * ls + s] += bee;
    Am[s * ls + s] += bee;
    int sm1 = (s + ls - 1) % ls, sp1 = (s + 1) % ls;
    Ap[s * ls + sm1] += -cee * (s == 0 ? -mp : 1.0);
    Am[s * ls + sp1] += -cee * (s == ls - 1 ? -mm : 1.0);
  }
  invert_dense(ls, Ap);
  invert_dense(ls, Am);
  std::vector<double> inv(4 * ls * ls);
  for (int i = 0; i < ls; i++)
    for (int j = 0; j < ls; j++) {
      inv[0 * ls * ls + i * ls + j] = Ap[i * ls + j];
      inv[1 * ls * ls + i * ls + j] = Am[i * ls + j];
      inv[2 * ls * ls + i * ls + j] = Ap[j * ls + i];
      inv[3 * ls * ls + i * ls + j] = Am[j * ls + i];
    }
  if (prec == CGPTB_SINGLE) {
    upload<float>(&s_coef, tab);
    upload<float>(&s_inv, inv);
  } else {
    upload<double>(&s_coef, tab);
    upload<double>(&s_inv, inv);
  }
}

// local blocks of a colour-matrix field, one per rank as ncclAllGather delivers them, -> the field on the global lattice
template <typename TU>
__global__ void k_assemble_global(Geom gg, Geom gl, int pg0, int pg1, int pg2, int pg3, size_t ns_local, const TU* __restrict__ blocks,
                                  TU* __restrict__ global) {
  size_t site = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= (size_t)2 * gg.half4) return;
  int p = site >= (size_t)gg.half4 ? 1 : 0;
  int c[4];
  cb_coords(gg, p, (int)(site - (size_t)p * gg.half4), c[0], c[1], c[2], c[3]);
  const int pc[4] = {c[0] / gl.L[0], c[1] / gl.L[1], c[2] / gl.L[2], c[3] / gl.L[3]};
  const int rank = pc[0] + pg0 * (pc[1] + pg1 * (pc[2] + pg2 * pc[3]));
  (void)pg3;
  const int x = c[0] % gl.L[0], y = c[1] % gl.L[1], z = c[2] % gl.L[2], t = c[3] % gl.L[3];
  const size_t ls = (size_t)((x + y + z + t) & 1) * gl.half4 + cb_index(gl, x, y, z, t);
  const TU* src = blocks + (size_t)rank * ns_local * 18;
  for (int k = 0; k < 9; k++) {
    size_t o = elem_offset<TU>(ns_local, ls, k, 1), og = elem_offset<TU>((size_t)2 * gg.half4, site, k, 1);
    global[og] = src[o];
    global[og + 1] = src[o + 1];
  }
}

template <typename TU, typename T>
static void import_gauge_t(cgptb_fermion_operator* op, const cgptb_lattice* const U[4]) {
  LinkCoef lc;
  double ani = op->type == CGPTB_WILSON_CLOVER ? op->p.nu / op->p.xi_0 : 1.0;
  for (int mu = 0; mu < 4; mu++) {
    lc.w[mu] = -0.5 * (mu < 3 ? ani : 1.0);
    lc.ph[2 * mu] = op->p.boundary_phases[2 * mu];
    lc.ph[2 * mu + 1] = op->p.boundary_phases[2 * mu + 1];
    lc.goff[mu] = op->goff[mu];
    lc.gL[mu] = op->gL[mu];
  }
  lc.open_bc = op->open_bc ? 1 : 0;
  size_t link_bytes = (size_t)op->g.half4 * 8 * 18 * sizeof(T);
  op->links_pad_valid = false;
  int threads = 128;
  unsigned blocks = (unsigned)((op->g.half4 + threads - 1) / threads);
  for (int p = 0; p < 2; p++) {
    if (!op->links[p]) CUDA_CHECK(cudaMalloc(&op->links[p], link_bytes));
    k_build_links<TU, T><<<blocks, threads, 0, g_stream>>>(op->g, p, U[0]->sites, (const TU*)U[0]->data, (const TU*)U[1]->data,
                                                           (const TU*)U[2]->data, (const TU*)U[3]->data, lc, (T*)op->links[p],
                                                           (const double*)op->ghost_links[1], (const double*)op->ghost_links[2],
                                                           (const double*)op->ghost_links[3]);
    LAUNCH_CHECK();
    if (op->compress) {
      const size_t n = (size_t)op->g.half4 * 8;
      if (!op->links_c[p]) CUDA_CHECK(cudaMalloc(&op->links_c[p], n * 14 * sizeof(T)));
      k_compress_links<T><<<(unsigned)((n + 127) / 128), 128, 0, g_stream>>>(n, (const T*)op->links[p], (T*)op->links_c[p]);
      LAUNCH_CHECK();
    }
  }
  op->has_clover = op->type == CGPTB_WILSON_CLOVER && (op->p.csw_r != 0.0 || op->p.csw_t != 0.0);
  if (op->has_clover) {
    CloverCoef cc;
    cc.diag = op->p.mass + 1.0 + 3.0 * op->p.nu / op->p.xi_0;
    cc.open_bc = op->open_bc ? 1 : 0;
    cc.T = op->gL[3];
    cc.edge = -0.5 * op->p.csw_t;
    cc.cF1 = op->p.cF - 1.0;
    // gamma matrices (lib/gpt/core/gamma.py:28-41) as (re,im) 4x4
    static const double gre[4][16] = {{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                                      {0, 0, 0, -1, 0, 0, 1, 0, 0, 1, 0, 0, -1, 0, 0, 0},
                                      {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                                      {0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 0, 0}};
    static const double gim[4][16] = {{0, 0, 0, 1, 0, 0, 1, 0, 0, -1, 0, 0, -1, 0, 0, 0},
                                      {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
                                      {0, 0, 1, 0, 0, 0, 0, -1, -1, 0, 0, 0, 0, 1, 0, 0},
                                      {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};
    int plane = 0;
    for (int mu = 0; mu < 4; mu++)
      for (int nu = mu + 1; nu < 4; nu++, plane++) {
        double cp = nu == 3 ? op->p.csw_t : op->p.csw_r / op->p.xi_0;
        cc.c[plane] = -0.5 * cp;
        // sigma = 1/2 (g_mu g_nu - g_nu g_mu)
        for (int i = 0; i < 4; i++)
          for (int j = 0; j < 4; j++) {
            double sr = 0, si = 0;
            for (int k = 0; k < 4; k++) {
              double ar = gre[mu][i * 4 + k], ai = gim[mu][i * 4 + k], br = gre[nu][k * 4 + j], bi = gim[nu][k * 4 + j];
              double cr = gre[nu][i * 4 + k], ci = gim[nu][i * 4 + k], dr = gre[mu][k * 4 + j], di = gim[mu][k * 4 + j];
              sr += (ar * br - ai * bi) - (cr * dr - ci * di);
              si += (ar * bi + ai * br) - (cr * di + ci * dr);
            }
            cc.sre[plane][i * 4 + j] = 0.5 * sr;
            cc.sim[plane][i * 4 + j] = 0.5 * si;
          }
      }
    // The field strength at a site needs links up to one step away in two directions, i.e. across the faces AND the corners
    // of a rank's volume.  On a decomposed lattice every rank therefore assembles the global (un-phased) links once -- an
    // all-gather of the local blocks over NVLink, 4 x 18 reals per global site, one box has the memory -- and builds its own
    // blocks from them; nothing of this is on the per-Dslash path.
    Geom gU = op->g;
    size_t nsU = U[0]->sites;
    const TU* Ug[4] = {(const TU*)U[0]->data, (const TU*)U[1]->data, (const TU*)U[2]->data, (const TU*)U[3]->data};
    TU* global[4] = {0, 0, 0, 0};
    TU* blocks = 0;
    int off[4] = {0, 0, 0, 0};
    if (op->g.comm_mask) {
      gU = make_geom(op->gL);
      nsU = (size_t)2 * gU.half4;
      const size_t local_reals = U[0]->sites * 18;
      CUDA_CHECK(cudaMalloc(&blocks, local_reals * g_comm.world * sizeof(TU)));
      for (int mu = 0; mu < 4; mu++) {
        CUDA_CHECK(cudaMalloc(&global[mu], nsU * 18 * sizeof(TU)));
        comm_allgather_device(U[mu]->data, blocks, local_reals * sizeof(TU), g_stream);
        k_assemble_global<TU><<<(unsigned)((nsU + 127) / 128), 128, 0, g_stream>>>(gU, op->g, g_comm.pgrid[0], g_comm.pgrid[1], g_comm.pgrid[2],
                                                                                  g_comm.pgrid[3], U[0]->sites, blocks, global[mu]);
        LAUNCH_CHECK();
        Ug[mu] = global[mu];
        off[mu] = op->goff[mu];
      }
    }
    size_t cl_bytes = (size_t)op->g.half4 * 72 * sizeof(T);
    unsigned cblocks = (unsigned)((op->g.half4 + 63) / 64);
    for (int p = 0; p < 2; p++) {
      if (!op->clov[p]) CUDA_CHECK(cudaMalloc(&op->clov[p], cl_bytes));
      if (!op->clov_inv[p]) CUDA_CHECK(cudaMalloc(&op->clov_inv[p], cl_bytes));
      k_build_clover<TU, T><<<cblocks, 64, 0, g_stream>>>(op->g, gU, off[0], off[1], off[2], off[3], p, nsU, Ug[0], Ug[1], Ug[2], Ug[3], cc,
                                                          (T*)op->clov[p], (T*)op->clov_inv[p]);
      LAUNCH_CHECK();
    }
    if (blocks) {
      CUDA_CHECK(cudaStreamSynchronize(g_stream));
      CUDA_CHECK(cudaFree(blocks));
      for (int mu = 0; mu < 4; mu++) CUDA_CHECK(cudaFree(global[mu]));
    }
  }
}

void cgptb_fermion_operator::import_gauge(const cgptb_lattice* const U[4]) {
  for (int mu = 0; mu < 4; mu++) {
    CGPTB_ASSERT(U[mu] != 0);
    if (U[mu]->otype != CGPTB_OT_MCOLOR || U[mu]->cb != CGPTB_FULL || U[mu]->Ls != 0)
      CGPTB_ERR("U[%d] must be a colour-matrix field on the full 4d grid", mu);
    for (int i = 0; i < 4; i++)
      if (U[mu]->dims4[i] != dims4[i]) CGPTB_ERR("U[%d] lives on a different grid", mu);
    if (U[mu]->prec != U[0]->prec) CGPTB_ERR("gauge links have mixed precision");
  }
  halo_setup(this, U);
  if (U[0]->prec == CGPTB_SINGLE) {
    if (prec == CGPTB_SINGLE)
      import_gauge_t<float, float>(this, U);
    else
      import_gauge_t<float, double>(this, U);
  } else {
    if (prec == CGPTB_SINGLE)
      import_gauge_t<double, float>(this, U);
    else
      import_gauge_t<double, double>(this, U);
  }
}

namespace cgptb {

static void clover_apply(cgptb_fermion_operator* op, bool inverse, bool acc, const cgptb_lattice* in, cgptb_lattice* out) {
  op->check_field(in);
  op->check_field(out);
  CGPTB_ASSERT(in->sites == out->sites);
  if (!acc) out->cb = in->cb;
  size_t half = op->g.half4;
  const int ls = op->ls();
  int threads = 128;
  unsigned blocks = (unsigned)((half * ls + threads - 1) / threads);
  for (int p = 0; p < 2; p++) {
    if (in->cb != CGPTB_FULL && in->cb != p) continue;
    size_t off = in->cb == CGPTB_FULL ? (size_t)p * half * ls : 0;
    void* cl = inverse ? op->clov_inv[p] : op->clov[p];
    if (op->prec == CGPTB_SINGLE) {
      const float* pin = (const float*)in->data + off * 8;
      float* pout = (float*)out->data + off * 8;
      if (acc)
        k_clover_apply<float, true><<<blocks, threads, 0, g_stream>>>(half, ls, pin, in->sites, pout, out->sites, (const float*)cl);
      else
        k_clover_apply<float, false><<<blocks, threads, 0, g_stream>>>(half, ls, pin, in->sites, pout, out->sites, (const float*)cl);
    } else {
      const double* pin = (const double*)in->data + off * 4;
      double* pout = (double*)out->data + off * 4;
      if (acc)
        k_clover_apply<double, true><<<blocks, threads, 0, g_stream>>>(half, ls, pin, in->sites, pout, out->sites, (const double*)cl);
      else
        k_clover_apply<double, false><<<blocks, threads, 0, g_stream>>>(half, ls, pin, in->sites, pout, out->sites, (const double*)cl);
    }
    LAUNCH_CHECK();
  }
}

// Mooee / MooeeInv (+Dag) on a half or full field
// out (+)= (a + i b gamma_5) in, gamma_5 = diag(1, 1, -1, -1): the site-diagonal term of the twisted-mass operator and its
// inverse (Grid's axpibg5x; lib/cgpt/lib/operators/wilson_twisted_mass.h)
template <typename T, bool ACC>
__global__ void __launch_bounds__(128) k_twist(size_t n, const T* __restrict__ in, size_t in_stride, T* __restrict__ out,
                                               size_t out_stride, T a, T b) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T psi[24], r[24];
  load_spinor(in, in_stride, i, psi);
  if (ACC) load_spinor_rw(out, out_stride, i, r);
#pragma unroll
  for (int c = 0; c < 12; c++) {
    const T bb = c < 6 ? b : -b;
    const T re = a * psi[2 * c] - bb * psi[2 * c + 1], im = a * psi[2 * c + 1] + bb * psi[2 * c];
    r[2 * c] = ACC ? r[2 * c] + re : re;
    r[2 * c + 1] = ACC ? r[2 * c + 1] + im : im;
  }
  store_spinor(out, out_stride, i, r);
}

template <typename T>
static void twist_apply(const cgptb_lattice* in, cgptb_lattice* out, bool acc, double a, double b) {
  unsigned blocks = (unsigned)((in->sites + 127) / 128);
  if (acc)
    k_twist<T, true><<<blocks, 128, 0, g_stream>>>(in->sites, (const T*)in->data, in->sites, (T*)out->data, out->sites, (T)a, (T)b);
  else
    k_twist<T, false><<<blocks, 128, 0, g_stream>>>(in->sites, (const T*)in->data, in->sites, (T*)out->data, out->sites, (T)a, (T)b);
  LAUNCH_CHECK();
}

static void op_mooee_impl(cgptb_fermion_operator* op, bool inverse, bool dag, bool acc, const cgptb_lattice* in, cgptb_lattice* out);

void op_mooee(cgptb_fermion_operator* op, bool inverse, bool dag, bool acc, const cgptb_lattice* in, cgptb_lattice* out) {
  op_mooee_impl(op, inverse, dag, acc, in, out);
  apply_open_boundaries(op, out);
}

static void op_mooee_impl(cgptb_fermion_operator* op, bool inverse, bool dag, bool acc, const cgptb_lattice* in, cgptb_lattice* out) {
  if (op->type == CGPTB_MOBIUS) {
    if (inverse) {
      CGPTB_ASSERT(!acc);
      static int no_sweep = getenv("CGPTB_NO_SWEEP") ? 1 : 0;
      if (no_sweep || !op_s_sweep(op, dag ? SWEEP_MINVDAG : SWEEP_MINV, in, out)) op_s_dense(op, dag, in, out);
    } else
      op_s_tridiag(op, 2, dag, acc, in, out);
    return;
  }
  if (op->has_clover) {
    clover_apply(op, inverse, acc, in, out);  // Hermitian: dag == non-dag
    return;
  }
  double diag = op->p.mass + 1.0 + 3.0 * op->p.nu / op->p.xi_0;
  if (op->p.mu != 0.0) {
    // twisted mass: (diag + i mu gamma_5), inverse (diag - i mu gamma_5) / (diag^2 + mu^2), dagger: mu -> -mu
    op->check_field(in);
    op->check_field(out);
    CGPTB_ASSERT(in->data != out->data || !acc);
    if (!acc) out->cb = in->cb;
    double a = diag, b = dag ? -op->p.mu : op->p.mu;
    if (inverse) {
      const double den = diag * diag + op->p.mu * op->p.mu;
      a = diag / den;
      b = -b / den;
    }
    if (op->prec == CGPTB_SINGLE)
      twist_apply<float>(in, out, acc, a, b);
    else
      twist_apply<double>(in, out, acc, a, b);
    return;
  }
  // plain Wilson: (m0 + 1 + 3 nu/xi_0) psi   (lib/cgpt/lib/operators/implementation.h:21-29)
  double coef[2] = {inverse ? 1.0 / diag : diag, 0.0};
  const cgptb_lattice* a[1] = {in};
  op->check_field(in);
  op->check_field(out);
  blas_lc(out, acc ? 1 : 0, 1, coef, a);
}

static void export5(cgptb_fermion_operator* op, bool solution, const cgptb_lattice* in, cgptb_lattice* out) {
  op->check_field(in);
  CGPTB_ASSERT(out->Ls == 0 && out->prec == op->prec && out->otype == CGPTB_OT_VSPINCOLOR && out->cb == in->cb &&
               out->sites * op->Ls == in->sites);
  size_t n4 = out->sites;
  unsigned blocks = (unsigned)((n4 + 127) / 128);
  if (op->prec == CGPTB_SINGLE) {
    if (solution)
      k_export5<float, true><<<blocks, 128, 0, g_stream>>>(n4, op->Ls, (const float*)in->data, (float*)out->data);
    else
      k_export5<float, false><<<blocks, 128, 0, g_stream>>>(n4, op->Ls, (const float*)in->data, (float*)out->data);
  } else {
    if (solution)
      k_export5<double, true><<<blocks, 128, 0, g_stream>>>(n4, op->Ls, (const double*)in->data, (double*)out->data);
    else
      k_export5<double, false><<<blocks, 128, 0, g_stream>>>(n4, op->Ls, (const double*)in->data, (double*)out->data);
  }
  LAUNCH_CHECK();
}

static void import5(cgptb_fermion_operator* op, const cgptb_lattice* in, cgptb_lattice* out) {
  op->check_field(out);
  CGPTB_ASSERT(in->Ls == 0 && in->prec == op->prec && in->otype == CGPTB_OT_VSPINCOLOR && in->sites * op->Ls == out->sites);
  out->cb = in->cb;
  size_t n4 = in->sites;
  unsigned blocks = (unsigned)((n4 + 127) / 128);
  if (op->prec == CGPTB_SINGLE)
    k_import5<float><<<blocks, 128, 0, g_stream>>>(n4, op->Ls, (const float*)in->data, (float*)out->data);
  else
    k_import5<double><<<blocks, 128, 0, g_stream>>>(n4, op->Ls, (const double*)in->data, (double*)out->data);
  LAUNCH_CHECK();
}

// Meooe / MeooeDag on half fields (also valid on full fields: hopping part of M)
void op_meooe(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out) {
  if (op->type != CGPTB_MOBIUS) {
    op_dhop(op, dag, in, out);
    return;
  }
  // Moebius: Meooe = Dhop (b + c S5) ; MeooeDag = (b + c S5)^dag Dhop^dag   
  if (!dag) {
    cgptb_lattice* t = op->tmp(0, in->cb);
    op_s_tridiag(op, 0, false, false, in, t);
    op_dhop(op, false, t, out);
  } else {
    cgptb_lattice* t = op->tmp(0, in->cb == CGPTB_FULL ? CGPTB_FULL : 1 - in->cb);
    op_dhop(op, true, in, t);
    op_s_tridiag(op, 0, true, false, t, out);
  }
}

static void op_dminus(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out) {
  if (op->type != CGPTB_MOBIUS) {
    blas_copy(out, in);
    return;
  }
  // Dminus psi = psi - c D_W psi, D_W = (4 - M5) + Dhop    
  CGPTB_ASSERT(in->cb == CGPTB_FULL && out->cb == CGPTB_FULL);
  cgptb_lattice* t = op->tmp(1, CGPTB_FULL);
  op_dhop(op, dag, in, t);
  if (op->zmobius) {  // psi_s - c_s D_W psi_s with complex c_s (conjugated for the adjoint)
    op_s_z(op, ZK_DM1, dag, false, in, out);
    op_s_z(op, ZK_DM2, dag, true, t, out);
    return;
  }
  double coef[4] = {1.0 - op->p.c * (4.0 - op->p.M5), 0.0, -op->p.c, 0.0};
  const cgptb_lattice* a[2] = {in, t};
  blas_lc(out, 0, 2, coef, a);
}

void op_apply(cgptb_fermion_operator* op, int opcode, const cgptb_lattice* src, cgptb_lattice* dst) {
  CGPTB_ASSERT(src != dst && src->data != dst->data);
  bool mob = op->type == CGPTB_MOBIUS;
  switch (opcode) {
    case CGPTB_OP_Dhop:
    case CGPTB_OP_DhopDag:
      CGPTB_ASSERT(src->cb == CGPTB_FULL);
      op_dhop(op, opcode == CGPTB_OP_DhopDag, src, dst);
      break;
    case CGPTB_OP_DhopEO:
    case CGPTB_OP_DhopEODag:
      CGPTB_ASSERT(src->cb != CGPTB_FULL);
      op_dhop(op, opcode == CGPTB_OP_DhopEODag, src, dst);
      break;
    case CGPTB_OP_Meooe:
    case CGPTB_OP_MeooeDag:
      CGPTB_ASSERT(src->cb != CGPTB_FULL);
      op_meooe(op, opcode == CGPTB_OP_MeooeDag, src, dst);
      break;
    case CGPTB_OP_Mooee:
    case CGPTB_OP_MooeeDag:
      op_mooee(op, false, opcode == CGPTB_OP_MooeeDag, false, src, dst);
      break;
    case CGPTB_OP_Mdiag:
      CGPTB_ASSERT(src->cb == CGPTB_FULL);
      op_mooee(op, false, false, false, src, dst);
      break;
    case CGPTB_OP_MooeeInv:
    case CGPTB_OP_MooeeInvDag:
      op_mooee(op, true, opcode == CGPTB_OP_MooeeInvDag, false, src, dst);
      break;
    case CGPTB_OP_M:
    case CGPTB_OP_Mdag: {
      // M = Meooe + Mooee on the full lattice
      CGPTB_ASSERT(src->cb == CGPTB_FULL && dst->cb == CGPTB_FULL);
      bool dag = opcode == CGPTB_OP_Mdag;
      op_meooe(op, dag, src, dst);
      op_mooee(op, false, dag, true, src, dst);
      break;
    }
    case CGPTB_OP_Dminus:
    case CGPTB_OP_DminusDag:
      op_dminus(op, opcode == CGPTB_OP_DminusDag, src, dst);
      break;
    case CGPTB_OP_ImportPhysicalFermionSource:
      if (!mob)
        blas_copy(dst, src);
      else {
        cgptb_lattice* t = op->tmp(2, CGPTB_FULL);
        CGPTB_ASSERT(src->cb == CGPTB_FULL);
        import5(op, src, t);
        op_dminus(op, false, t, dst);
      }
      break;
    case CGPTB_OP_ImportUnphysicalFermion:
      if (!mob)
        blas_copy(dst, src);
      else
        import5(op, src, dst);
      break;
    case CGPTB_OP_ExportPhysicalFermionSolution:
    case CGPTB_OP_ExportPhysicalFermionSource:
      if (!mob)
        blas_copy(dst, src);
      else
        export5(op, opcode == CGPTB_OP_ExportPhysicalFermionSolution, src, dst);
      break;
    default:
      CGPTB_ERR("Unknown opcode %d", opcode);
  }
}

}  // namespace cgptb

extern "C" {

int cgptb_create_fermion_operator(cgptb_fermion_operator** out, int optype, int precision, const cgptb_fermion_params* params,
                                  const cgptb_lattice* const U[4]) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(optype == CGPTB_WILSON_CLOVER || optype == CGPTB_MOBIUS);
  CGPTB_ASSERT(precision == CGPTB_SINGLE || precision == CGPTB_DOUBLE);
  CGPTB_ASSERT(U && U[0]);
  cgptb_fermion_operator* op = new cgptb_fermion_operator();
  op->type = optype;
  op->prec = precision;
  for (int i = 0; i < 4; i++) op->dims4[i] = U[0]->dims4[i];
  op->g = make_geom(op->dims4);
  op->p = *params;
  // Wilson-clover with Ls > 0: multi-rhs operator, the fifth dimension enumerates right-hand sides that share the links
  // (wilson_clover(n_rhs=...), lib/gpt/qcd/fermion/wilson.py:134-138)
  op->Ls = params->Ls > 0 ? params->Ls : 0;
  if (params->link_compression != 0 && params->link_compression != 12) {
    delete op;
    CGPTB_ERR("link_compression must be 0 (18 reals per link) or 12 (two rows), got %d", params->link_compression);
  }
  op->compress = params->link_compression == 12;
  try {
    if (optype == CGPTB_MOBIUS) {
      if (params->Ls < 1) CGPTB_ERR("mobius needs Ls >= 1");
      if (params->n_omega > 0) {
        if (params->n_omega != params->Ls || params->Ls > 64) CGPTB_ERR("zmobius needs Ls = len(omega) <= 64");
        op->zmobius = true;
      }
      op->setup_mobius_tables();
    } else {
      if (params->boundary_phases[6] == 0.0 && params->boundary_phases[7] == 0.0) {
        // open boundary conditions in time (reference/wilson_clover.py:78-88: isotropic, csw_r == csw_t, cF given)
        if (params->xi_0 != 1.0 || params->nu != 1.0 || params->csw_r != params->csw_t)
          CGPTB_ERR("open boundary conditions need xi_0 = nu = 1 and csw_r = csw_t");
        op->open_bc = true;
      }
      if (params->xi_0 == 0.0) CGPTB_ERR("xi_0 must be non-zero");
      if (params->mu != 0.0 && (params->csw_r != 0.0 || params->csw_t != 0.0)) CGPTB_ERR("twisted mass with a clover term is not a GPT operator");
    }
    op->import_gauge(U);
  } catch (...) {
    cgptb_delete_fermion_operator(op);
    throw;
  }
  *out = op;
  CGPTB_API_END
}

int cgptb_update_fermion_operator(cgptb_fermion_operator* op, const cgptb_lattice* const U[4]) {
  CGPTB_API_BEGIN
  op->import_gauge(U);
  CGPTB_API_END
}

int cgptb_set_mass_fermion_operator(cgptb_fermion_operator* op, const cgptb_fermion_params* params) {
  CGPTB_API_BEGIN
  if (op->type == CGPTB_MOBIUS) {
    op->p.mass_plus = params->mass_plus;
    op->p.mass_minus = params->mass_minus;
    op->setup_mobius_tables();
  } else {
    op->p.mass = params->mass;  // clover blocks are rebuilt by the update that follows (interface.py:43-45)
  }
  CGPTB_API_END
}

int cgptb_delete_fermion_operator(cgptb_fermion_operator* op) {
  CGPTB_API_BEGIN
  if (op) {
    dhop_tma_release(op);
    halo_release(op);
    for (int p = 0; p < 2; p++) {
      if (op->links[p]) cudaFree(op->links[p]);
      if (op->links_c[p]) cudaFree(op->links_c[p]);
      if (op->clov[p]) cudaFree(op->clov[p]);
      if (op->clov_inv[p]) cudaFree(op->clov_inv[p]);
    }
    for (int mu = 0; mu < 4; mu++) {
      for (int side = 0; side < 2; side++) {
        if (op->halo_send[mu][side]) cudaFree(op->halo_send[mu][side]);
        if (op->halo_recv[mu][side]) cudaFree(op->halo_recv[mu][side]);
      }
      if (op->ghost_links[mu]) cudaFree(op->ghost_links[mu]);
    }
    if (op->s_coef) cudaFree(op->s_coef);
    if (op->s_inv) cudaFree(op->s_inv);
    if (op->z_tab) cudaFree(op->z_tab);
    for (int i = 0; i < 4; i++) {
      if (op->tmp_full[i]) cgptb_delete_lattice(op->tmp_full[i]);
      if (op->tmp_half[i]) cgptb_delete_lattice(op->tmp_half[i]);
    }
    for (int i = 0; i < 5; i++)
      if (op->cg_half[i]) cgptb_delete_lattice(op->cg_half[i]);
    delete op;
  }
  CGPTB_API_END
}

int cgptb_apply_fermion_operator(cgptb_fermion_operator* op, int opcode, const cgptb_lattice* src, cgptb_lattice* dst) {
  CGPTB_API_BEGIN
  op_apply(op, opcode, src, dst);
  CGPTB_API_END
}
}

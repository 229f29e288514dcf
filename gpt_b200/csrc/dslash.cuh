// Wilson hopping term building blocks: spin projection / reconstruction in GPT's gamma basis
// (lib/gpt/core/gamma.py:28-41) and the SU(3) x half-spinor product, all in registers.
//
//   Dhop psi(x) = -1/2 sum_mu c_mu [ (1 - g_mu) U_mu(x) psi(x+mu) + (1 + g_mu) U_mu^dag(x-mu) psi(x-mu) ]
//   (lib/gpt/qcd/fermion/reference/wilson_clover.py:182-200; dag: g_mu -> -g_mu)
//
// The factor -c_mu/2 and the boundary phases are folded into the stored links (as Grid does at ImportGauge).
#pragma once
#include "common.cuh"

namespace cgptb {

// multiply (re,im) by i^PH
template <int PH, typename T>
__device__ __forceinline__ void mul_iph(T re, T im, T& ore, T& oim) {
  if (PH == 0) {
    ore = re;
    oim = im;
  } else if (PH == 1) {
    ore = -im;
    oim = re;
  } else if (PH == 2) {
    ore = -re;
    oim = -im;
  } else {
    ore = im;
    oim = -re;
  }
}

// Projector table for (1 + SGN*gamma_MU), SGN = -1 or +1:
//   h0 = psi_0 + i^A psi_J0 ; h1 = psi_1 + i^B psi_J1 ; (result)_2 = i^C2 h_K2 ; (result)_3 = i^C3 h_K3
// (1 - gamma_mu) rows derived from gamma.py:28-41; (1 + gamma_mu) flips every phase by i^2.
template <int MU, int SGN>
struct Proj;
template <>
struct Proj<0, -1> { enum { J0 = 3, A = 3, J1 = 2, B = 3, K2 = 1, C2 = 1, K3 = 0, C3 = 1 }; };
template <>
struct Proj<1, -1> { enum { J0 = 3, A = 0, J1 = 2, B = 2, K2 = 1, C2 = 2, K3 = 0, C3 = 0 }; };
template <>
struct Proj<2, -1> { enum { J0 = 2, A = 3, J1 = 3, B = 1, K2 = 0, C2 = 1, K3 = 1, C3 = 3 }; };
template <>
struct Proj<3, -1> { enum { J0 = 2, A = 2, J1 = 3, B = 2, K2 = 0, C2 = 2, K3 = 1, C3 = 2 }; };
template <int MU>
struct Proj<MU, +1> {
  typedef Proj<MU, -1> M;
  enum { J0 = M::J0, A = (M::A + 2) % 4, J1 = M::J1, B = (M::B + 2) % 4, K2 = M::K2, C2 = (M::C2 + 2) % 4, K3 = M::K3, C3 = (M::C3 + 2) % 4 };
};

// h[12]: (hspin*3+color)*2+reim
template <int MU, int SGN, typename T>
__device__ __forceinline__ void project(const T (&psi)[24], T (&h)[12]) {
  typedef Proj<MU, SGN> P;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    T ar, ai;
    mul_iph<P::A>(psi[(P::J0 * 3 + c) * 2], psi[(P::J0 * 3 + c) * 2 + 1], ar, ai);
    h[(0 * 3 + c) * 2] = psi[(0 * 3 + c) * 2] + ar;
    h[(0 * 3 + c) * 2 + 1] = psi[(0 * 3 + c) * 2 + 1] + ai;
    mul_iph<P::B>(psi[(P::J1 * 3 + c) * 2], psi[(P::J1 * 3 + c) * 2 + 1], ar, ai);
    h[(1 * 3 + c) * 2] = psi[(1 * 3 + c) * 2] + ar;
    h[(1 * 3 + c) * 2 + 1] = psi[(1 * 3 + c) * 2 + 1] + ai;
  }
}

template <int MU, int SGN, typename T>
__device__ __forceinline__ void reconstruct_add(T (&acc)[24], const T (&chi)[12]) {
  typedef Proj<MU, SGN> P;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    acc[(0 * 3 + c) * 2] += chi[(0 * 3 + c) * 2];
    acc[(0 * 3 + c) * 2 + 1] += chi[(0 * 3 + c) * 2 + 1];
    acc[(1 * 3 + c) * 2] += chi[(1 * 3 + c) * 2];
    acc[(1 * 3 + c) * 2 + 1] += chi[(1 * 3 + c) * 2 + 1];
    T r, i;
    mul_iph<P::C2>(chi[(P::K2 * 3 + c) * 2], chi[(P::K2 * 3 + c) * 2 + 1], r, i);
    acc[(2 * 3 + c) * 2] += r;
    acc[(2 * 3 + c) * 2 + 1] += i;
    mul_iph<P::C3>(chi[(P::K3 * 3 + c) * 2], chi[(P::K3 * 3 + c) * 2 + 1], r, i);
    acc[(3 * 3 + c) * 2] += r;
    acc[(3 * 3 + c) * 2 + 1] += i;
  }
}

// chi = W h (ADJ = false) or W^dag h (ADJ = true); W[18]: (row*3+col)*2+reim
template <bool ADJ, typename T>
__device__ __forceinline__ void su3_mul(const T (&W)[18], const T (&h)[12], T (&chi)[12]) {
#pragma unroll
  for (int sp = 0; sp < 2; sp++) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
      T re = 0, im = 0;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        T wr, wi;
        if (!ADJ) {
          wr = W[(r * 3 + c) * 2];
          wi = W[(r * 3 + c) * 2 + 1];
        } else {
          wr = W[(c * 3 + r) * 2];
          wi = -W[(c * 3 + r) * 2 + 1];
        }
        T hr = h[(sp * 3 + c) * 2], hi = h[(sp * 3 + c) * 2 + 1];
        re += wr * hr - wi * hi;
        im += wr * hi + wi * hr;
      }
      chi[(sp * 3 + r) * 2] = re;
      chi[(sp * 3 + r) * 2 + 1] = im;
    }
  }
}

// links: [i4][8 dirs][9 complex], dir d<4: w U_d(x), d>=4: w U_{d-4}(x - mu) (dagger applied here)
template <typename T>
__device__ __forceinline__ void load_link(const T* __restrict__ links, size_t i4, int dir, T (&W)[18]);
template <>
__device__ __forceinline__ void load_link<float>(const float* __restrict__ links, size_t i4, int dir, float (&W)[18]) {
  const float2* b = reinterpret_cast<const float2*>(links) + (i4 * 8 + dir) * 9;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    float2 v = __ldg(b + k);
    W[2 * k] = v.x;
    W[2 * k + 1] = v.y;
  }
}
template <>
__device__ __forceinline__ void load_link<double>(const double* __restrict__ links, size_t i4, int dir, double (&W)[18]) {
  const double2* b = reinterpret_cast<const double2*>(links) + (i4 * 8 + dir) * 9;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    double2 v = __ldg(b + k);
    W[2 * k] = v.x;
    W[2 * k + 1] = v.y;
  }
}

// two-row compression: [i4][8 dirs][row 0, row 1 (6 complex), f (1 complex)]; row 2 = f conj(row 0 x row 1).  f is the U(1)
// factor of the stored link w U (w = -c_mu/2 x boundary phase): f = w / conj(w)^2, fitted from the link itself when the table
// is built (k_compress_links), 0 for the vanishing links of open boundary conditions
template <typename T>
__device__ __forceinline__ void load_link_c(const T* __restrict__ links_c, size_t i4, int dir, T (&W)[18]) {
  const T* b = links_c + (i4 * 8 + dir) * 14;
#pragma unroll
  for (int k = 0; k < 12; k++) W[k] = __ldg(b + k);
  const T fr = __ldg(b + 12), fi = __ldg(b + 13);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const int a = (c + 1) % 3, bb = (c + 2) % 3;
    // d = row0[a] row1[bb] - row0[bb] row1[a]
    T dr = W[2 * a] * W[6 + 2 * bb] - W[2 * a + 1] * W[6 + 2 * bb + 1] - (W[2 * bb] * W[6 + 2 * a] - W[2 * bb + 1] * W[6 + 2 * a + 1]);
    T di = W[2 * a] * W[6 + 2 * bb + 1] + W[2 * a + 1] * W[6 + 2 * bb] - (W[2 * bb] * W[6 + 2 * a + 1] + W[2 * bb + 1] * W[6 + 2 * a]);
    // f conj(d)
    W[12 + 2 * c] = fr * dr + fi * di;
    W[12 + 2 * c + 1] = fi * dr - fr * di;
  }
}

// neighbour checkerboard index of output site (x,y,z,t) in direction MU, forward (FWD) or backward
template <int MU, bool FWD>
__device__ __forceinline__ int neighbor(const Geom& g, int x, int y, int z, int t) {
  int c[4] = {x, y, z, t};
  if (FWD)
    c[MU] = c[MU] + 1 == g.L[MU] ? 0 : c[MU] + 1;
  else
    c[MU] = c[MU] == 0 ? g.L[MU] - 1 : c[MU] - 1;
  return cb_index(g, c[0], c[1], c[2], c[3]);
}

// true if the hop in direction MU (forward / backward) leaves the local volume of a split direction
template <int MU, bool FWD>
__device__ __forceinline__ bool off_rank(const Geom& g, int x, int y, int z, int t) {
  if (!((g.comm_mask >> MU) & 1)) return false;
  int c = MU == 0 ? x : (MU == 1 ? y : (MU == 2 ? z : t));
  return FWD ? c == g.L[MU] - 1 : c == 0;
}

// one direction of the stencil: acc += recon( W(^dag) proj psi(neighbour) )
template <int MU, bool FWD, bool DAG, bool CMP = false, typename T>
__device__ __forceinline__ void hop(T (&acc)[24], const Geom& g, int x, int y, int z, int t, int i4, int s, int ls,
                                    const T* __restrict__ in, size_t in_stride, const T* __restrict__ links) {
  // forward hop uses (1 - g_mu), backward (1 + g_mu); daggered operator swaps them
  const int SGN = (FWD != DAG) ? -1 : +1;
  if (off_rank<MU, FWD>(g, x, y, z, t)) return;
  int n4 = neighbor<MU, FWD>(g, x, y, z, t);
  T psi[24], h[12], chi[12], W[18];
  load_spinor(in, in_stride, (size_t)n4 * ls + s, psi);
  project<MU, SGN>(psi, h);
  if (CMP)
    load_link_c<T>(links, (size_t)i4, FWD ? MU : MU + 4, W);
  else
    load_link<T>(links, (size_t)i4, FWD ? MU : MU + 4, W);
  su3_mul<!FWD>(W, h, chi);
  reconstruct_add<MU, SGN>(acc, chi);
}

}  // namespace cgptb

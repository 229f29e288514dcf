// Gauge observables the hot path's callers use to validate a configuration: g.qcd.gauge.plaquette
// (lib/gpt/qcd/gauge/stencil/plaquette.py:23-44: average of Re tr U_mu(x) U_nu(x+mu) U_mu^dag(x+nu) U_nu^dag(x) / Nc over the six
// planes and all sites) and the link trace that NERSC headers carry (lib/gpt/core/io/nersc_io.py:238-247).
// One thread per site, double accumulation, block reduction, one atomicAdd per block.
#include "common.cuh"

namespace cgptb {

template <typename T>
__device__ __forceinline__ void load_u(const T* __restrict__ U, size_t nsites, size_t site, double (&m)[18]) {
#pragma unroll
  for (int k = 0; k < 9; k++) {
    size_t o = elem_offset<T>(nsites, site, k, 1);
    m[2 * k] = (double)U[o];
    m[2 * k + 1] = (double)U[o + 1];
  }
}

// c = a b (ADJB: c = a b^dag), 3x3 complex as 18 doubles
template <bool ADJB>
__device__ __forceinline__ void mul33(const double (&a)[18], const double (&b)[18], double (&c)[18]) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double re = 0, im = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double ar = a[2 * (3 * i + k)], ai = a[2 * (3 * i + k) + 1];
        const double br = ADJB ? b[2 * (3 * j + k)] : b[2 * (3 * k + j)];
        const double bi = ADJB ? -b[2 * (3 * j + k) + 1] : b[2 * (3 * k + j) + 1];
        re += ar * br - ai * bi;
        im += ar * bi + ai * br;
      }
      c[2 * (3 * i + j)] = re;
      c[2 * (3 * i + j) + 1] = im;
    }
}

// out[0] += sum Re tr plaquettes, out[1] += sum Re tr U_mu
template <typename T>
__global__ void __launch_bounds__(128) k_plaquette(Geom g, size_t nsites, const T* U0, const T* U1, const T* U2, const T* U3,
                                                   double* __restrict__ out) {
  const T* U[4] = {U0, U1, U2, U3};
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double v[2] = {0.0, 0.0};
  if (idx < 2 * (size_t)g.half4) {
    const int p = idx >= (size_t)g.half4 ? 1 : 0;
    const int i4 = (int)(idx - (size_t)p * g.half4);
    int c[4];
    cb_coords(g, p, i4, c[0], c[1], c[2], c[3]);
    auto site_of = [&](int mu, int nu) -> size_t {  // x + mu (mu >= 0) + nu (nu >= 0, optional -1)
      int d[4] = {c[0], c[1], c[2], c[3]};
      int par = p;
      if (mu >= 0) {
        d[mu] = d[mu] + 1 == g.L[mu] ? 0 : d[mu] + 1;
        par ^= 1;
      }
      if (nu >= 0) {
        d[nu] = d[nu] + 1 == g.L[nu] ? 0 : d[nu] + 1;
        par ^= 1;
      }
      return (size_t)par * g.half4 + cb_index(g, d[0], d[1], d[2], d[3]);
    };
    double um[4][18];
    for (int mu = 0; mu < 4; mu++) {
      load_u(U[mu], nsites, idx, um[mu]);
      v[1] += um[mu][0] + um[mu][8] + um[mu][16];
    }
    for (int mu = 1; mu < 4; mu++)
      for (int nu = 0; nu < mu; nu++) {
        double a[18], b[18], t1[18], t2[18];
        load_u(U[nu], nsites, site_of(mu, -1), a);  // U_nu(x+mu)
        load_u(U[mu], nsites, site_of(nu, -1), b);  // U_mu(x+nu)
        mul33<false>(um[mu], a, t1);
        mul33<true>(t1, b, t2);
        mul33<true>(t2, um[nu], t1);
        v[0] += t1[0] + t1[8] + t1[16];
      }
  }
  __shared__ double red[64];
  block_reduce<2>(v, red);
  if (threadIdx.x == 0) {
    atomicAdd(out, v[0]);
    atomicAdd(out + 1, v[1]);
  }
}

}  // namespace cgptb

using namespace cgptb;

extern "C" {

// out[0] = plaquette (g.qcd.gauge.plaquette), out[1] = link trace sum_mu <Re tr U_mu> / (4 Nc); global over all ranks.
// The lattice must not be split across GPUs in this version (the staples would need the neighbours' links).
int cgptb_gauge_plaquette(const cgptb_lattice* const U[4], double out[2]) {
  CGPTB_API_BEGIN
  for (int mu = 0; mu < 4; mu++) CGPTB_ASSERT(U[mu] && U[mu]->otype == 9 && U[mu]->Ls == 0 && U[mu]->cb == CGPTB_FULL && same_shape(U[mu], U[0]));
  if (g_comm.active) CGPTB_ERR("gauge_plaquette on a split lattice is not implemented (needs the neighbours' links)");
  Geom g = make_geom(U[0]->dims4);
  double* d = reduce_scratch(8);
  CUDA_CHECK(cudaMemsetAsync(d, 0, 2 * sizeof(double), g_stream));
  const size_t n = U[0]->sites;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (U[0]->prec == CGPTB_SINGLE)
    k_plaquette<float><<<blocks, 128, 0, g_stream>>>(g, n, (const float*)U[0]->data, (const float*)U[1]->data, (const float*)U[2]->data,
                                                     (const float*)U[3]->data, d);
  else
    k_plaquette<double><<<blocks, 128, 0, g_stream>>>(g, n, (const double*)U[0]->data, (const double*)U[1]->data,
                                                      (const double*)U[2]->data, (const double*)U[3]->data, d);
  LAUNCH_CHECK();
  double h[2];
  CUDA_CHECK(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  out[0] = h[0] / (double)n / 6.0 / 3.0;
  out[1] = h[1] / (double)n / 4.0 / 3.0;
  CGPTB_API_END
}
}

// ---------------------------------------------------------------------------------------------------------
// NERSC gauge configurations: the byte work between the file and the link lattices, on the device.
// File order (lib/gpt/core/io/nersc_io.py:146-199): sites lexicographic with x fastest, per site the four directions,
// per link 2 x 3 (4D_SU3_GAUGE, third row = conj(row0 x row1), lib/cgpt/lib/munge.h:21-39) or 3 x 3 complex numbers,
// IEEE32/64, big or little endian.  Checksum = sum of the data as native-endian 32-bit words
// (lib/cgpt/lib/checksums/nersc.h:19-33).  One thread per (site, direction).
// ---------------------------------------------------------------------------------------------------------
namespace cgptb {

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

template <typename TF, typename TL>
__global__ void __launch_bounds__(128) k_nersc_munge(Geom g, size_t nsites, int big_endian, int rows, const unsigned char* __restrict__ raw,
                                                     TL* U0, TL* U1, TL* U2, TL* U3, unsigned int* __restrict__ checksum) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // = lex site * 4 + mu
  uint32_t cs = 0;
  if (idx < nsites * 4) {
    const int mu = (int)(idx & 3);
    size_t lex = idx >> 2;
    const int x = (int)(lex % g.L[0]);
    lex /= g.L[0];
    const int y = (int)(lex % g.L[1]);
    lex /= g.L[1];
    const int z = (int)(lex % g.L[2]);
    const int t = (int)(lex / g.L[2]);
    const int nval = rows * 3 * 2;
    const unsigned char* p = raw + idx * (size_t)nval * sizeof(TF);
    double m[18];
    for (int k = 0; k < nval; k++) {
      if (sizeof(TF) == 4) {
        uint32_t w = reinterpret_cast<const uint32_t*>(p)[k];
        if (big_endian) w = bswap32(w);
        cs += w;
        m[k] = (double)__uint_as_float(w);
      } else {
        uint32_t a = reinterpret_cast<const uint32_t*>(p)[2 * k], b = reinterpret_cast<const uint32_t*>(p)[2 * k + 1];
        uint32_t lo = a, hi = b;
        if (big_endian) {
          lo = bswap32(b);
          hi = bswap32(a);
        }
        cs += lo;
        cs += hi;
        m[k] = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
      }
    }
    if (rows == 2) {
      // third row = conj(row0 x row1)
      for (int c = 0; c < 3; c++) {
        const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
        const double ar = m[2 * c1], ai = m[2 * c1 + 1], br = m[2 * (3 + c2)], bi = m[2 * (3 + c2) + 1];
        const double cr = m[2 * c2], ci = m[2 * c2 + 1], dr = m[2 * (3 + c1)], di = m[2 * (3 + c1) + 1];
        const double re = (ar * br - ai * bi) - (cr * dr - ci * di);
        const double im = (ar * bi + ai * br) - (cr * di + ci * dr);
        m[2 * (6 + c)] = re;
        m[2 * (6 + c) + 1] = -im;
      }
    }
    const int par = (x + y + z + t) & 1;
    const size_t site = (size_t)par * g.half4 + cb_index(g, x, y, z, t);
    TL* U = mu == 0 ? U0 : (mu == 1 ? U1 : (mu == 2 ? U2 : U3));
    for (int k = 0; k < 9; k++) {
      const size_t o = elem_offset<TL>(nsites, site, k, 1);
      U[o] = (TL)m[2 * k];
      U[o + 1] = (TL)m[2 * k + 1];
    }
  }
  // block sum of the checksum (wraps mod 2^32 by construction)
  for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(checksum, cs);
}

}  // namespace cgptb

extern "C" {

// raw_host: the data part of a NERSC file (nbytes = sites * 4 * rows * 3 * 2 * float_size); fills the four link lattices
// (any precision) and returns the NERSC checksum of the data
int cgptb_nersc_munge(const void* raw_host, size_t nbytes, int float_size, int big_endian, int rows, cgptb_lattice* const U[4],
                      unsigned int* checksum) {
  CGPTB_API_BEGIN
  for (int mu = 0; mu < 4; mu++) CGPTB_ASSERT(U[mu] && U[mu]->otype == 9 && U[mu]->Ls == 0 && U[mu]->cb == CGPTB_FULL && same_shape(U[mu], U[0]));
  CGPTB_ASSERT((float_size == 4 || float_size == 8) && (rows == 2 || rows == 3));
  if (g_comm.active) CGPTB_ERR("nersc_munge on a split lattice is not implemented (every rank would read its own block)");
  const size_t n = U[0]->sites;
  if (nbytes != n * 4 * (size_t)rows * 6 * (size_t)float_size)
    CGPTB_ERR("nersc_munge: %zu bytes of data do not match %zu sites of %d x 3 matrices of %d-byte floats", nbytes, n, rows, float_size);
  unsigned char* raw = 0;
  CUDA_CHECK(cudaMalloc(&raw, nbytes + 4));
  unsigned int* dcs = reinterpret_cast<unsigned int*>(reduce_scratch(8));
  CUDA_CHECK(cudaMemsetAsync(dcs, 0, sizeof(unsigned int), g_stream));
  CUDA_CHECK(cudaMemcpyAsync(raw, raw_host, nbytes, cudaMemcpyHostToDevice, g_stream));
  Geom g = make_geom(U[0]->dims4);
  const unsigned blocks = (unsigned)((n * 4 + 127) / 128);
  const bool ld = U[0]->prec == CGPTB_DOUBLE;
#define MUNGE(TF, TL)                                                                                                        \
  k_nersc_munge<TF, TL><<<blocks, 128, 0, g_stream>>>(g, n, big_endian, rows, raw, (TL*)U[0]->data, (TL*)U[1]->data, (TL*)U[2]->data, \
                                                      (TL*)U[3]->data, dcs)
  if (float_size == 4 && ld) MUNGE(float, double);
  else if (float_size == 4) MUNGE(float, float);
  else if (ld) MUNGE(double, double);
  else MUNGE(double, float);
#undef MUNGE
  LAUNCH_CHECK();
  CUDA_CHECK(cudaMemcpyAsync(checksum, dcs, sizeof(unsigned int), cudaMemcpyDeviceToHost, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  CUDA_CHECK(cudaFree(raw));
  CGPTB_API_END
}
}

// Shared declarations of libcgpt_b200: lattice object, HBM layout accessors, error handling.
//
// HBM layout (DESIGN.md "Data layout"):
//   * a full-lattice field is stored as [even half][odd half]; a checkerboarded field is one half.
//     Within a half the 4d sites are in checkerboard order i4 = (x/2) + Lx/2*(y + Ly*(z + Lz*t)); a 5d
//     field has the s index fastest: site = i4*Ls + s (Grid's order, evidence:
//     lib/cgpt/lib/foundation/mobius_with_vector_field.h:79-81).
//   * site-major SoA in 32-byte vector blocks for spin-colour vectors: block k of site i lives at
//     data32[k*nsites + i]; fp32 spinor = 3 blocks of 4 complex, fp64 spinor = 6 blocks of 2 complex.  One
//     LDG.E.ENL2.256 (ld.global.v8.f32 / v4.f64, sm_100) moves a block, so a neighbour spinor costs 3 (fp32) or
//     6 (fp64) load instructions and address computations instead of 6 / 12 with 16-byte blocks, and a warp
//     reading block k of 32 consecutive sites issues one fully coalesced 1 KB request.
//     Other objects (links, singlets): one complex per block ([component][site]).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/cgpt_b200.h"

namespace cgptb {

extern thread_local std::string g_error;
extern cudaStream_t g_stream;
extern uint64_t g_launches;

// processor grid (comm.cu); inactive for a single GPU
struct Comm {
  bool active = false;
  int rank = 0, world = 1;
  int pgrid[4] = {1, 1, 1, 1};
  int pcoor[4] = {0, 0, 0, 0};
  void* nccl = 0;
  cudaStream_t stream = 0;
  cudaEvent_t ev_pack = 0, ev_comm = 0;
};
extern Comm g_comm;
int comm_neighbor_rank(int mu, int dir);
void comm_exchange_begin();
void comm_exchange_dir(int mu, const void* to_lo, const void* to_hi, void* from_lo, void* from_hi, size_t bytes, cudaStream_t s);
void comm_exchange_end();
void comm_allreduce_device(double* dev, int n, cudaStream_t s);
struct CommExport {  // what a rank publishes about a device allocation its neighbours may write into
  unsigned char handle[64];
  unsigned long long offset;
};
void comm_allgather_host(const void* in, void* out, size_t bytes);
void comm_allgather_device(const void* in, void* out, size_t bytes, cudaStream_t s);
bool comm_p2p_available();
void comm_export(void* ptr, CommExport* e);
void* comm_import(int rank, const CommExport* e);  // 0 if the allocation cannot be mapped
void comm_stream_write32(cudaStream_t s, void* addr, unsigned value);
void comm_stream_wait_geq32(cudaStream_t s, void* addr, unsigned value);
extern bool g_reduce_global;

struct Error {
  std::string msg;
};

#define CGPTB_ERR(...)                                   \
  do {                                                   \
    char _buf[1024];                                     \
    snprintf(_buf, sizeof(_buf), __VA_ARGS__);           \
    throw cgptb::Error{std::string(_buf)};               \
  } while (0)

#define CGPTB_ASSERT(x)                                                              \
  do {                                                                               \
    if (!(x)) CGPTB_ERR("Assert failed: %s (%s:%d)", #x, __FILE__, __LINE__);        \
  } while (0)

#define CUDA_CHECK(x)                                                                                  \
  do {                                                                                                 \
    cudaError_t _e = (x);                                                                              \
    if (_e != cudaSuccess) CGPTB_ERR("CUDA error %s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// wrap every C-ABI body: exceptions become a status code + message (cgpt: exception.h:23-39)
#define CGPTB_API_BEGIN try {
#define CGPTB_API_END                                    \
  }                                                      \
  catch (const cgptb::Error& e) {                        \
    cgptb::g_error = e.msg;                              \
    fprintf(stderr, "cgpt_b200: %s\n", e.msg.c_str());   \
    return 1;                                            \
  }                                                      \
  return 0;

#define LAUNCH_CHECK()              \
  do {                              \
    cgptb::g_launches++;            \
    CUDA_CHECK(cudaGetLastError()); \
  } while (0)

}  // namespace cgptb

struct cgptb_lattice {
  int prec;     // CGPTB_SINGLE / CGPTB_DOUBLE
  int otype;    // complex components per site
  int dims4[4];
  int Ls;       // 0: 4d
  int cb;       // CGPTB_EVEN / ODD / FULL
  size_t sites4;  // stored 4d sites (V4 or V4/2)
  size_t sites;   // stored sites (sites4 * max(Ls,1))
  void* data;
  bool owns;
  void* alloc = 0;  // what cudaMalloc returned (data = alloc + skew, see create_lattice)
  size_t alloc_bytes = 0;

  int ls() const { return Ls > 0 ? Ls : 1; }
  size_t real_size() const { return prec == CGPTB_DOUBLE ? 8 : 4; }
  size_t nreals() const { return sites * (size_t)otype * 2; }
  size_t bytes() const { return nreals() * real_size(); }
  size_t half4() const { return (size_t)dims4[0] * dims4[1] * dims4[2] * dims4[3] / 2; }
  // complex components per 16-byte (or 8-byte) block
  int cpb() const { return otype % 4 == 0 ? (prec == CGPTB_SINGLE ? 4 : 2) : 1; }
  // bytes of one vector block
  size_t block_bytes() const { return (size_t)cpb() * 2 * real_size(); }
};

namespace cgptb {

inline bool same_shape(const cgptb_lattice* a, const cgptb_lattice* b) {
  return a->prec == b->prec && a->otype == b->otype && a->Ls == b->Ls && a->sites == b->sites &&
         a->dims4[0] == b->dims4[0] && a->dims4[1] == b->dims4[1] && a->dims4[2] == b->dims4[2] &&
         a->dims4[3] == b->dims4[3];
}

template <typename T>
struct vec_of;
template <>
struct vec_of<float> {
  typedef float4 type;
};
template <>
struct vec_of<double> {
  typedef double2 type;
};

// ---- 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256) ------------------------------------
__device__ __forceinline__ void ld256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void ld256(const double* p, double (&v)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
// coherent variant for fields that are read and written by the same kernel
__device__ __forceinline__ void ld256_rw(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void ld256_rw(const double* p, double (&v)[4]) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st256(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void st256(double* p, const double (&v)[4]) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void st256_cs(float* p, const float (&v)[8]) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// ---- spinor accessors (24 reals: index (spin*3+color)*2 + reim) -----------------------------------
// base: first block of the (half) field, nsites: sites per component plane, site: index within the plane
__device__ __forceinline__ void load_spinor(const float* __restrict__ base, size_t nsites, size_t site, float (&p)[24]) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float v[8];
    ld256(base + (k * nsites + site) * 8, v);
#pragma unroll
    for (int e = 0; e < 8; e++) p[8 * k + e] = v[e];
  }
}
__device__ __forceinline__ void load_spinor(const double* __restrict__ base, size_t nsites, size_t site, double (&p)[24]) {
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double v[4];
    ld256(base + (k * nsites + site) * 4, v);
#pragma unroll
    for (int e = 0; e < 4; e++) p[4 * k + e] = v[e];
  }
}
// same, for a field the kernel also writes (no read-only path)
__device__ __forceinline__ void load_spinor_rw(const float* base, size_t nsites, size_t site, float (&p)[24]) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float v[8];
    ld256_rw(base + (k * nsites + site) * 8, v);
#pragma unroll
    for (int e = 0; e < 8; e++) p[8 * k + e] = v[e];
  }
}
__device__ __forceinline__ void load_spinor_rw(const double* base, size_t nsites, size_t site, double (&p)[24]) {
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double v[4];
    ld256_rw(base + (k * nsites + site) * 4, v);
#pragma unroll
    for (int e = 0; e < 4; e++) p[4 * k + e] = v[e];
  }
}
__device__ __forceinline__ void store_spinor(float* __restrict__ base, size_t nsites, size_t site, const float (&p)[24]) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = p[8 * k + e];
    st256(base + (k * nsites + site) * 8, v);
  }
}
__device__ __forceinline__ void store_spinor(double* __restrict__ base, size_t nsites, size_t site, const double (&p)[24]) {
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double v[4];
#pragma unroll
    for (int e = 0; e < 4; e++) v[e] = p[4 * k + e];
    st256(base + (k * nsites + site) * 4, v);
  }
}

// generic complex-element accessor for any object type (used by import/export and setup kernels)
template <typename T>
__device__ __forceinline__ size_t elem_offset(size_t nsites, size_t site, int c, int cpb) {
  // offset in units of T of the real part of complex component c of `site`
  return ((size_t)(c / cpb) * nsites + site) * (2 * cpb) + (size_t)(c % cpb) * 2;
}

// geometry of one parity of the local 4d lattice
struct Geom {
  int L[4];       // x,y,z,t (local extents)
  int hx;         // L[0]/2
  int half4;      // sites per parity
  int comm_mask;  // bit mu set: direction mu is split across GPUs, hops leaving the local volume are skipped
};

inline Geom make_geom(const int dims4[4]) {
  Geom g;
  for (int i = 0; i < 4; i++) g.L[i] = dims4[i];
  g.hx = dims4[0] / 2;
  g.half4 = dims4[0] / 2 * dims4[1] * dims4[2] * dims4[3];
  g.comm_mask = 0;
  return g;
}

__host__ __device__ __forceinline__ void cb_coords(const Geom& g, int p, int i, int& x, int& y, int& z, int& t) {
  int xh = i % g.hx;
  int r = i / g.hx;
  y = r % g.L[1];
  r /= g.L[1];
  z = r % g.L[2];
  t = r / g.L[2];
  x = 2 * xh + ((y + z + t + p) & 1);
}

__host__ __device__ __forceinline__ int cb_index(const Geom& g, int x, int y, int z, int t) {
  return (x >> 1) + g.hx * (y + g.L[1] * (z + g.L[2] * t));
}

// warp + block reduction of doubles (n values per thread), result valid in thread 0
template <int N>
__device__ __forceinline__ void block_reduce(double (&v)[N], double* smem /* [32*N] */) {
#pragma unroll
  for (int i = 0; i < N; i++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; i++) smem[w * N + i] = v[i];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      double x = lane < nw ? smem[lane * N + i] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      v[i] = x;
    }
  }
}

// scratch for reductions
double* reduce_scratch(size_t n_doubles);  // device
double* reduce_host(size_t n_doubles);     // pinned host
int sm_count();

// internal entry points shared between translation units
void blas_axpy(cgptb_lattice* r, double are, double aim, const cgptb_lattice* x, const cgptb_lattice* y);
void blas_copy(cgptb_lattice* d, const cgptb_lattice* s);
void blas_zero(cgptb_lattice* d);
void blas_lc(cgptb_lattice* dst, int accumulate, int n, const double* coef, const cgptb_lattice* const* a);

}  // namespace cgptb

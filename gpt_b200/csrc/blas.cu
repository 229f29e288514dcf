// CG vector kernels of libcgpt_b200: axpy, fused axpy+norm2, inner products / norms with double
// accumulation, linear combinations.  Replaces lib/cgpt/lib/transform.cc:143-246,
// foundation/transform.h:233-248, foundation/reduce.h:76-327,
// expression/linear_combination_implementation.h:154-192 and foundation/basis.h:21-97.
//
// All of these are pure HBM streaming kernels.  Every field is a flat array of interleaved complex
// numbers (see common.cuh), processed in 16-byte vectors by a persistent grid (sm_count * 8 CTAs of 256
// threads, grid-stride loop, 4 independent vectors in flight per thread).  Reductions: per-thread double
// accumulators -> warp shuffle -> one partial per CTA -> a single-CTA second stage that adds the
// partials in a fixed order (bit-reproducible, no atomics), result copied to pinned host memory.
#include "common.cuh"

namespace cgptb {

bool g_reduce_global = false;  // solver.cu: sum reduction results over all ranks before they reach the host
// solver.cu, device-resident CG: reduction results stay on the device (no copy, no synchronisation); the pointer to the last
// result is left in g_reduce_dev_ptr
bool g_reduce_to_device = false;
double* g_reduce_dev_ptr = 0;

static const int BT = 256;
static const int UNROLL = 4;

static inline unsigned grid_for(size_t nvec) {
  size_t want = (nvec + (size_t)BT * UNROLL - 1) / ((size_t)BT * UNROLL);
  size_t cap = (size_t)sm_count() * 8;
  if (want < 1) want = 1;
  return (unsigned)(want < cap ? want : cap);
}

template <typename T>
struct V;
template <>
struct V<float> {
  typedef float4 vec;
  static const int NC = 2;  // complex per vector
};
template <>
struct V<double> {
  typedef double2 vec;
  static const int NC = 1;
};

__device__ __forceinline__ float4 caxpy(float ar, float ai, float4 x, float4 y) {
  float4 r;
  r.x = fmaf(ar, x.x, fmaf(-ai, x.y, y.x));
  r.y = fmaf(ar, x.y, fmaf(ai, x.x, y.y));
  r.z = fmaf(ar, x.z, fmaf(-ai, x.w, y.z));
  r.w = fmaf(ar, x.w, fmaf(ai, x.z, y.w));
  return r;
}
__device__ __forceinline__ double2 caxpy(double ar, double ai, double2 x, double2 y) {
  double2 r;
  r.x = fma(ar, x.x, fma(-ai, x.y, y.x));
  r.y = fma(ar, x.y, fma(ai, x.x, y.y));
  return r;
}
__device__ __forceinline__ float4 vzero(float4*) { return make_float4(0, 0, 0, 0); }
__device__ __forceinline__ double2 vzero(double2*) { return make_double2(0, 0); }

__device__ __forceinline__ void acc_norm2(double& s, float4 v) {
  s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
}
__device__ __forceinline__ void acc_norm2(double& s, double2 v) { s += v.x * v.x + v.y * v.y; }
// conj(a) * b
__device__ __forceinline__ void acc_dot(double& re, double& im, float4 a, float4 b) {
  re += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
  im += (double)a.x * b.y - (double)a.y * b.x + (double)a.z * b.w - (double)a.w * b.z;
}
__device__ __forceinline__ void acc_dot(double& re, double& im, double2 a, double2 b) {
  re += a.x * b.x + a.y * b.y;
  im += a.x * b.y - a.y * b.x;
}

// r = a x + y, optionally |r|^2
template <typename T, bool NORM>
// r may alias x and / or y (cg.py: axpy(r, -a, mmp, r), axpy(p, b, p, r)): element-wise, so that is well defined as long as
// no pointer carries __restrict__ (the loads must not be moved to the read-only path)
__global__ void __launch_bounds__(BT) k_axpy(size_t nvec, T ar, T ai, const typename V<T>::vec* x,
                                             const typename V<T>::vec* y, typename V<T>::vec* r,
                                             double* __restrict__ partial, const double* scal_dev = 0, double scal_mult = 0.0) {
  typedef typename V<T>::vec vec;
  if (scal_dev) {  // real scalar computed on the device (CG: a = c / d)
    ar = (T)(scal_mult * *scal_dev);
    ai = (T)0;
  }
  double s = 0.0;
  size_t stride = (size_t)gridDim.x * BT;
  for (size_t i = (size_t)blockIdx.x * BT + threadIdx.x; i < nvec; i += stride * UNROLL) {
    vec xv[UNROLL], yv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      size_t j = i + u * stride;
      if (j < nvec) {
        xv[u] = x[j];
        yv[u] = y[j];
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      size_t j = i + u * stride;
      if (j < nvec) {
        vec rv = caxpy(ar, ai, xv[u], yv[u]);
        r[j] = rv;
        if (NORM) acc_norm2(s, rv);
      }
    }
  }
  if (NORM) {
    __shared__ double sm[32];
    double v[1] = {s};
    block_reduce<1>(v, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = v[0];
  }
}

// partial[b*3 + {0,1,2}] = re<a,b>, im<a,b>, |a|^2
template <typename T, bool DOT, bool NRM>
__global__ void __launch_bounds__(BT) k_reduce(size_t nvec, const typename V<T>::vec* __restrict__ a,
                                               const typename V<T>::vec* __restrict__ b, double* __restrict__ partial) {
  typedef typename V<T>::vec vec;
  double re = 0.0, im = 0.0, n2 = 0.0;
  size_t stride = (size_t)gridDim.x * BT;
  for (size_t i = (size_t)blockIdx.x * BT + threadIdx.x; i < nvec; i += stride * UNROLL) {
    vec av[UNROLL], bv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      size_t j = i + u * stride;
      av[u] = vzero((vec*)0);
      bv[u] = vzero((vec*)0);
      if (j < nvec) {
        av[u] = a[j];
        if (DOT) bv[u] = b[j];
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (DOT) acc_dot(re, im, av[u], bv[u]);
      if (NRM) acc_norm2(n2, av[u]);
    }
  }
  __shared__ double sm[96];
  double v[3] = {re, im, n2};
  block_reduce<3>(v, sm);
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 3 + 0] = v[0];
    partial[blockIdx.x * 3 + 1] = v[1];
    partial[blockIdx.x * 3 + 2] = v[2];
  }
}

// second stage: out[c] = sum_b partial[b*ncomp + c], fixed order
__global__ void __launch_bounds__(BT) k_final(int nblocks, int ncomp, const double* __restrict__ partial,
                                              double* __restrict__ out) {
  __shared__ double sm[96];
  double v[3] = {0, 0, 0};
  for (int b = threadIdx.x; b < nblocks; b += BT)
    for (int c = 0; c < ncomp; c++) v[c] += partial[b * ncomp + c];
  block_reduce<3>(v, sm);
  if (threadIdx.x == 0)
    for (int c = 0; c < ncomp; c++) out[c] = v[c];
}

static const int MAXT = 8;
template <typename T>
struct LcArgs {
  const typename V<T>::vec* a[MAXT];
  T cr[MAXT], ci[MAXT];
  int n;
};

// dst may be one of the inputs (psi += a p; lattice *= a): no __restrict__
template <typename T, bool ACC>
__global__ void __launch_bounds__(BT) k_lc(size_t nvec, LcArgs<T> args, typename V<T>::vec* dst) {
  typedef typename V<T>::vec vec;
  size_t stride = (size_t)gridDim.x * BT;
  for (size_t i = (size_t)blockIdx.x * BT + threadIdx.x; i < nvec; i += stride) {
    vec r = ACC ? dst[i] : vzero((vec*)0);
    for (int t = 0; t < args.n; t++) r = caxpy(args.cr[t], args.ci[t], args.a[t][i], r);
    dst[i] = r;
  }
}

template <typename T>
static size_t nvec_of(const cgptb_lattice* l) {
  return l->nreals() * sizeof(T) / 16;
}

static void check_vec(const cgptb_lattice* l) {
  if ((l->nreals() * l->real_size()) % 16) CGPTB_ERR("field size %zu bytes is not a multiple of 16", l->bytes());
}

template <typename T>
static void axpy_t(cgptb_lattice* r, double are, double aim, const cgptb_lattice* x, const cgptb_lattice* y, double* norm2) {
  typedef typename V<T>::vec vec;
  size_t nvec = nvec_of<T>(x);
  unsigned g = grid_for(nvec);
  if (!norm2) {
    k_axpy<T, false><<<g, BT, 0, g_stream>>>(nvec, (T)are, (T)aim, (const vec*)x->data, (const vec*)y->data, (vec*)r->data, 0);
    LAUNCH_CHECK();
    return;
  }
  double* part = reduce_scratch((size_t)sm_count() * 8 * 3 + 8);
  double* out = part + (size_t)sm_count() * 8 * 3;
  k_axpy<T, true><<<g, BT, 0, g_stream>>>(nvec, (T)are, (T)aim, (const vec*)x->data, (const vec*)y->data, (vec*)r->data, part);
  LAUNCH_CHECK();
  k_final<<<1, BT, 0, g_stream>>>((int)g, 1, part, out);
  LAUNCH_CHECK();
  if (g_reduce_global) comm_allreduce_device(out, 1, g_stream);
  double* h = reduce_host(8);
  CUDA_CHECK(cudaMemcpyAsync(h, out, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  *norm2 = h[0];
}

// r = (mult * *scal_dev) x + y and |r|^2, everything on the device: returns where the (globally summed) |r|^2 is
double* blas_axpy_norm2_dev(cgptb_lattice* r, const double* scal_dev, double mult, const cgptb_lattice* x, const cgptb_lattice* y) {
  CGPTB_ASSERT(same_shape(r, x) && same_shape(r, y));
  check_vec(r);
  r->cb = x->cb;
  double* part = reduce_scratch((size_t)sm_count() * 8 * 3 + 8);
  double* out = part + (size_t)sm_count() * 8 * 3;
  if (r->prec == CGPTB_SINGLE) {
    typedef V<float>::vec vec;
    size_t nvec = nvec_of<float>(x);
    unsigned g = grid_for(nvec);
    k_axpy<float, true><<<g, BT, 0, g_stream>>>(nvec, 0.f, 0.f, (const vec*)x->data, (const vec*)y->data, (vec*)r->data, part, scal_dev, mult);
    LAUNCH_CHECK();
    k_final<<<1, BT, 0, g_stream>>>((int)g, 1, part, out);
  } else {
    typedef V<double>::vec vec;
    size_t nvec = nvec_of<double>(x);
    unsigned g = grid_for(nvec);
    k_axpy<double, true><<<g, BT, 0, g_stream>>>(nvec, 0.0, 0.0, (const vec*)x->data, (const vec*)y->data, (vec*)r->data, part, scal_dev, mult);
    LAUNCH_CHECK();
    k_final<<<1, BT, 0, g_stream>>>((int)g, 1, part, out);
  }
  LAUNCH_CHECK();
  if (g_reduce_global) comm_allreduce_device(out, 1, g_stream);
  return out;
}

void blas_axpy(cgptb_lattice* r, double are, double aim, const cgptb_lattice* x, const cgptb_lattice* y) {
  CGPTB_ASSERT(same_shape(r, x) && same_shape(r, y));
  check_vec(r);
  r->cb = x->cb;
  if (r->prec == CGPTB_SINGLE)
    axpy_t<float>(r, are, aim, x, y, 0);
  else
    axpy_t<double>(r, are, aim, x, y, 0);
}

// res = {re<a,b>, im<a,b>, |a|^2}
template <typename T>
static void reduce_t(const cgptb_lattice* a, const cgptb_lattice* b, bool dot, bool nrm, double res[3]) {
  typedef typename V<T>::vec vec;
  size_t nvec = nvec_of<T>(a);
  unsigned g = grid_for(nvec);
  double* part = reduce_scratch((size_t)sm_count() * 8 * 3 + 8);
  double* out = part + (size_t)sm_count() * 8 * 3;
  const vec* pa = (const vec*)a->data;
  const vec* pb = b ? (const vec*)b->data : pa;
  if (dot && nrm)
    k_reduce<T, true, true><<<g, BT, 0, g_stream>>>(nvec, pa, pb, part);
  else if (dot)
    k_reduce<T, true, false><<<g, BT, 0, g_stream>>>(nvec, pa, pb, part);
  else
    k_reduce<T, false, true><<<g, BT, 0, g_stream>>>(nvec, pa, pb, part);
  LAUNCH_CHECK();
  k_final<<<1, BT, 0, g_stream>>>((int)g, 3, part, out);
  LAUNCH_CHECK();
  if (g_reduce_global) comm_allreduce_device(out, 3, g_stream);
  double* h = reduce_host(8);
  CUDA_CHECK(cudaMemcpyAsync(h, out, 3 * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  res[0] = h[0];
  res[1] = h[1];
  res[2] = h[2];
}

// per-CTA partial sums written by another kernel (e.g. the fused Dslash epilogue) -> host, fixed order
static double* g_partial = 0;
static size_t g_partial_n = 0;
double* blas_partial_scratch(int nblocks) {
  size_t n = (size_t)nblocks * 3 + 8;
  if (n > g_partial_n) {
    if (g_partial) CUDA_CHECK(cudaFree(g_partial));
    CUDA_CHECK(cudaMalloc(&g_partial, n * sizeof(double)));
    g_partial_n = n;
  }
  return g_partial;
}

void blas_finalize(int nblocks, int ncomp, const double* partial, double* host_out) {
  CGPTB_ASSERT(ncomp <= 3);
  double* out = const_cast<double*>(partial) + (size_t)nblocks * 3;
  k_final<<<1, BT, 0, g_stream>>>(nblocks, ncomp, partial, out);
  LAUNCH_CHECK();
  if (g_reduce_global) comm_allreduce_device(out, ncomp, g_stream);
  if (g_reduce_to_device) {
    g_reduce_dev_ptr = out;
    return;
  }
  double* h = reduce_host(8);
  CUDA_CHECK(cudaMemcpyAsync(h, out, ncomp * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  for (int i = 0; i < ncomp; i++) host_out[i] = h[i];
}

static void reduce_any(const cgptb_lattice* a, const cgptb_lattice* b, bool dot, bool nrm, double res[3]) {
  check_vec(a);
  if (b) CGPTB_ASSERT(same_shape(a, b));
  if (a->prec == CGPTB_SINGLE)
    reduce_t<float>(a, b, dot, nrm, res);
  else
    reduce_t<double>(a, b, dot, nrm, res);
}

template <typename T>
static void lc_t(cgptb_lattice* dst, int accumulate, int n, const double* coef, const cgptb_lattice* const* a) {
  typedef typename V<T>::vec vec;
  size_t nvec = nvec_of<T>(dst);
  size_t want = (nvec + BT - 1) / BT, cap = (size_t)sm_count() * 16;
  unsigned g = (unsigned)(want < cap ? want : cap);
  int done = 0;
  bool acc = accumulate != 0;
  if (n == 0 && !acc) {
    blas_zero(dst);
    return;
  }
  // more than MAXT terms go in chunks that accumulate into dst: a term beyond the first chunk that IS dst must see the
  // value dst had on entry, so it is read from a copy
  void* snapshot = 0;
  for (int t = MAXT; t < n && !snapshot; t++)
    if (a[t]->data == dst->data) {
      CUDA_CHECK(cudaMallocAsync(&snapshot, dst->bytes(), g_stream));
      CUDA_CHECK(cudaMemcpyAsync(snapshot, dst->data, dst->bytes(), cudaMemcpyDeviceToDevice, g_stream));
    }
  while (done < n) {
    LcArgs<T> args;
    args.n = (n - done) < MAXT ? (n - done) : MAXT;
    for (int t = 0; t < args.n; t++) {
      args.a[t] = (const vec*)((snapshot && done > 0 && a[done + t]->data == dst->data) ? snapshot : a[done + t]->data);
      args.cr[t] = (T)coef[2 * (done + t)];
      args.ci[t] = (T)coef[2 * (done + t) + 1];
    }
    if (acc)
      k_lc<T, true><<<g, BT, 0, g_stream>>>(nvec, args, (vec*)dst->data);
    else
      k_lc<T, false><<<g, BT, 0, g_stream>>>(nvec, args, (vec*)dst->data);
    LAUNCH_CHECK();
    done += args.n;
    acc = true;
  }
  if (snapshot) CUDA_CHECK(cudaFreeAsync(snapshot, g_stream));
}

void blas_lc(cgptb_lattice* dst, int accumulate, int n, const double* coef, const cgptb_lattice* const* a) {
  check_vec(dst);
  for (int i = 0; i < n; i++) CGPTB_ASSERT(same_shape(dst, a[i]));
  if (n > 0 && !accumulate) dst->cb = a[0]->cb;
  if (dst->prec == CGPTB_SINGLE)
    lc_t<float>(dst, accumulate, n, coef, a);
  else
    lc_t<double>(dst, accumulate, n, coef, a);
}

// out[t] = sum over the sites of time slice t of conj(b) a  (spin-colour vectors, 4d)
template <typename T>
__global__ void __launch_bounds__(BT) k_slice_dot(Geom g, size_t nsites, const T* __restrict__ a, const T* __restrict__ b,
                                                  int otype, int cpb, double* __restrict__ out) {
  int t = blockIdx.x;
  int per_t = g.half4 / g.L[3];
  double re = 0.0, im = 0.0;
  for (int p = 0; p < 2; p++)
    for (int i = threadIdx.x; i < per_t; i += BT) {
      size_t site = (size_t)p * g.half4 + (size_t)t * per_t + i;
      for (int c = 0; c < otype; c++) {
        size_t o = elem_offset<T>(nsites, site, c, cpb);
        double ar = a[o], ai = a[o + 1], br = b[o], bi = b[o + 1];
        re += br * ar + bi * ai;
        im += br * ai - bi * ar;
      }
    }
  __shared__ double sm[64];
  double v[2] = {re, im};
  block_reduce<2>(v, sm);
  if (threadIdx.x == 0) {
    out[2 * t] = v[0];
    out[2 * t + 1] = v[1];
  }
}


// d = a[x_dim] * s for every complex component (gpt.scale_per_coordinate, lib/gpt/core/transform.py:210-214 ->
// cgpt.lattice_scale_per_coordinate): one thread per (component, site), the coordinate is recovered from the device site index
template <typename T>
__global__ void k_scale_per_coordinate(Geom g, int cb, int ls, int nd, int dim, int otype, int cpb, size_t nsites, const T* __restrict__ src,
                                       T* __restrict__ dst, const double* __restrict__ a) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nsites * otype) return;
  const int c = (int)(idx / nsites);
  const size_t j = idx - (size_t)c * nsites;
  const size_t per = (size_t)g.half4 * ls;
  int p = cb;
  size_t rem = j;
  if (cb == CGPTB_FULL) {
    p = j >= per ? 1 : 0;
    rem = j - (size_t)p * per;
  }
  const int i4 = (int)(rem / ls);
  const int s = (int)(rem - (size_t)i4 * ls);
  int co[5];
  co[0] = s;
  cb_coords(g, p, i4, co[1], co[2], co[3], co[4]);
  const int x = nd == 5 ? co[dim] : co[dim + 1];
  const T ar = (T)a[2 * x], ai = (T)a[2 * x + 1];
  const size_t o = elem_offset<T>(nsites, j, c, cpb);
  const T re = src[o], im = src[o + 1];
  dst[o] = ar * re - ai * im;
  dst[o + 1] = ar * im + ai * re;
}


// dst = G src for a constant 4 x 4 complex spin matrix G (gamma matrices, chiral projectors, sigma_{mu nu}: g.gamma[...] * field,
// lib/gpt/core/gamma.py:28-80); one thread per site, zero entries of G are skipped (uniform branch)
struct SpinMatrix {
  double re[16], im[16];
};
template <typename T>
__global__ void __launch_bounds__(128) k_spin_matrix(size_t n, const T* __restrict__ in, T* __restrict__ out, SpinMatrix G) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T psi[24], r[24];
  load_spinor(in, n, i, psi);
#pragma unroll
  for (int a = 0; a < 4; a++) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      T re = 0, im = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const T gr = (T)G.re[4 * a + b], gi = (T)G.im[4 * a + b];
        if (gr == 0 && gi == 0) continue;
        re += gr * psi[(3 * b + c) * 2] - gi * psi[(3 * b + c) * 2 + 1];
        im += gr * psi[(3 * b + c) * 2 + 1] + gi * psi[(3 * b + c) * 2];
      }
      r[(3 * a + c) * 2] = re;
      r[(3 * a + c) * 2 + 1] = im;
    }
  }
  store_spinor(out, n, i, r);
}

}  // namespace cgptb

using namespace cgptb;

extern "C" {

int cgptb_lattice_spin_matrix(cgptb_lattice* d, const cgptb_lattice* s, const double* m_re_im) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(d && s && m_re_im && same_shape(d, s) && d->data != s->data);
  if (s->otype != CGPTB_OT_VSPINCOLOR) CGPTB_ERR("spin matrices act on spin-colour vector fields");
  d->cb = s->cb;
  SpinMatrix G;
  for (int k = 0; k < 16; k++) {
    G.re[k] = m_re_im[2 * k];
    G.im[k] = m_re_im[2 * k + 1];
  }
  const unsigned blocks = (unsigned)((s->sites + 127) / 128);
  if (s->prec == CGPTB_SINGLE)
    k_spin_matrix<float><<<blocks, 128, 0, g_stream>>>(s->sites, (const float*)s->data, (float*)d->data, G);
  else
    k_spin_matrix<double><<<blocks, 128, 0, g_stream>>>(s->sites, (const double*)s->data, (double*)d->data, G);
  LAUNCH_CHECK();
  CGPTB_API_END
}

int cgptb_lattice_scale_per_coordinate(cgptb_lattice* d, const cgptb_lattice* s, const double* a_re_im, int n, int dim) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(d && s && a_re_im && same_shape(d, s));
  const int nd = s->Ls > 0 ? 5 : 4;
  CGPTB_ASSERT(dim >= 0 && dim < nd);
  const int ext = (nd == 5 && dim == 0) ? s->Ls : s->dims4[dim - (nd - 4)];
  if (n != ext) CGPTB_ERR("scale_per_coordinate: %d factors for a dimension of extent %d", n, ext);
  d->cb = s->cb;
  double* tab = reduce_scratch((size_t)sm_count() * 8 * 3 + 8 + 2 * (size_t)n + 16) + (size_t)sm_count() * 8 * 3 + 16;
  CUDA_CHECK(cudaMemcpyAsync(tab, a_re_im, 2 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));  // the host array may be a temporary of the caller
  Geom g = make_geom(s->dims4);
  const size_t total = s->sites * (size_t)s->otype;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (s->prec == CGPTB_SINGLE)
    k_scale_per_coordinate<float><<<blocks, 256, 0, g_stream>>>(g, s->cb, s->ls(), nd, dim, s->otype, s->cpb(), s->sites, (const float*)s->data,
                                                                (float*)d->data, tab);
  else
    k_scale_per_coordinate<double><<<blocks, 256, 0, g_stream>>>(g, s->cb, s->ls(), nd, dim, s->otype, s->cpb(), s->sites,
                                                                 (const double*)s->data, (double*)d->data, tab);
  LAUNCH_CHECK();
  CGPTB_API_END
}

int cgptb_lattice_axpy(cgptb_lattice* r, double a_re, double a_im, const cgptb_lattice* x, const cgptb_lattice* y) {
  CGPTB_API_BEGIN
  blas_axpy(r, a_re, a_im, x, y);
  CGPTB_API_END
}

int cgptb_lattice_axpy_norm2(cgptb_lattice* r, double a_re, double a_im, const cgptb_lattice* x, const cgptb_lattice* y,
                             double* norm2) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(same_shape(r, x) && same_shape(r, y));
  check_vec(r);
  r->cb = x->cb;
  if (r->prec == CGPTB_SINGLE)
    axpy_t<float>(r, a_re, a_im, x, y, norm2);
  else
    axpy_t<double>(r, a_re, a_im, x, y, norm2);
  CGPTB_API_END
}

int cgptb_lattice_rank_inner_product(const cgptb_lattice* const* left, int n_left, const cgptb_lattice* const* right,
                                     int n_right, double* result) {
  CGPTB_API_BEGIN
  for (int i = 0; i < n_left; i++)
    for (int j = 0; j < n_right; j++) {
      double res[3];
      if (left[i] == right[j]) {
        reduce_any(left[i], 0, false, true, res);
        result[2 * (i * n_right + j)] = res[2];
        result[2 * (i * n_right + j) + 1] = 0.0;
      } else {
        reduce_any(left[i], right[j], true, false, res);
        result[2 * (i * n_right + j)] = res[0];
        result[2 * (i * n_right + j) + 1] = res[1];
      }
    }
  CGPTB_API_END
}

int cgptb_lattice_norm2(const cgptb_lattice* a, double* norm2) {
  CGPTB_API_BEGIN
  double res[3];
  reduce_any(a, 0, false, true, res);
  *norm2 = res[2];
  CGPTB_API_END
}

int cgptb_lattice_inner_product_norm2(const cgptb_lattice* a, const cgptb_lattice* b, double* ip, double* a2) {
  CGPTB_API_BEGIN
  double res[3];
  reduce_any(a, b, true, true, res);
  ip[0] = res[0];
  ip[1] = res[1];
  *a2 = res[2];
  CGPTB_API_END
}

int cgptb_lattice_lc(cgptb_lattice* dst, int accumulate, int n, const double* coef, const cgptb_lattice* const* a) {
  CGPTB_API_BEGIN
  blas_lc(dst, accumulate, n, coef, a);
  CGPTB_API_END
}

int cgptb_linear_combination(cgptb_lattice* const* r, int n_r, const cgptb_lattice* const* basis, int n_basis,
                             const double* Qt) {
  CGPTB_API_BEGIN
  for (int i = 0; i < n_r; i++) blas_lc(r[i], 0, n_basis, Qt + 2 * (size_t)i * n_basis, basis);
  CGPTB_API_END
}

int cgptb_lattice_scale(cgptb_lattice* l, double a_re, double a_im) {
  CGPTB_API_BEGIN
  double coef[2] = {a_re, a_im};
  const cgptb_lattice* a[1] = {l};
  blas_lc(l, 0, 1, coef, a);
  CGPTB_API_END
}

int cgptb_lattice_slice_inner_product(const cgptb_lattice* b, const cgptb_lattice* a, double* out) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(same_shape(a, b) && a->cb == CGPTB_FULL && a->Ls == 0);
  Geom g = make_geom(a->dims4);
  int T = a->dims4[3];
  double* d = reduce_scratch((size_t)sm_count() * 8 * 3 + 8 + 2 * T);
  if (a->prec == CGPTB_SINGLE)
    k_slice_dot<float><<<T, BT, 0, g_stream>>>(g, a->sites, (const float*)a->data, (const float*)b->data, a->otype, a->cpb(), d);
  else
    k_slice_dot<double><<<T, BT, 0, g_stream>>>(g, a->sites, (const double*)a->data, (const double*)b->data, a->otype, a->cpb(), d);
  LAUNCH_CHECK();
  double* h = reduce_host(2 * T > 8 ? 2 * T : 8);
  CUDA_CHECK(cudaMemcpyAsync(h, d, 2 * T * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  for (int i = 0; i < 2 * T; i++) out[i] = h[i];
  CGPTB_API_END
}
}

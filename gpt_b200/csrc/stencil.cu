// Generic matrix-vector stencil executor: cgpt.stencil_matrix_vector_create / _execute / _delete
// (lib/cgpt/lib/stencil.cc:41-121, lib/cgpt/lib/stencil/matrix_vector.h:20-290; Python side lib/gpt/core/local_stencil/
// matrix_vector.py, benchmarks/stencil.py:91-145: the covariant Laplacian written as a code list).
//
// A stencil is a list of shifts ("points") and a program ("code"): every line
//     vector[target](x) = weight * M_1 ... M_n vector[source](x + p_source)  [+ vector[accumulate](x)]
// with M_j = matrix[index_j](x + p_j) or its adjoint, the factors applied right to left (the reference loops j from size-1
// down to 0).  The code is cut into blocks of code_parallel_block_size lines; lines of a block run one after the other, blocks
// are independent (the reference runs them as one more parallel index).  Here: matrices are colour matrices, vectors are
// spin-colour or colour vectors, on the full 4d lattice; one thread per (site, block) keeps the vector in registers, the
// links come through L1/L2, neighbour sites through the checkerboard index arithmetic of common.cuh -- HBM bound like the
// hopping term (8 x 18 + 2 x 24 reals per site for the Laplacian).  One rank only (the reference's padded / halo variants
// belong to its copy-plan machinery).
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace cgptb {

struct StCode {
  int target, accumulate, source, source_point;
  int nfac, fac0;
  double wr, wi;
};
struct StFactor {
  int index, point, adj;
};

}  // namespace cgptb

struct cgptb_stencil_mv {
  int dims4[4];
  int prec;
  int npoints, ncode, nblocks, block_size;
  int max_m, max_v;  // highest field index the code refers to
  int* d_points = 0;
  cgptb::StCode* d_code = 0;
  cgptb::StFactor* d_fac = 0;
  void** d_fields = 0;  // [matrix pointers..., vector pointers...], refreshed per execute
  int cap_fields = 0;
};

namespace cgptb {

__device__ __forceinline__ size_t shifted_site(const Geom& g, int x, int y, int z, int t, const int* p) {
  int c[4] = {x + p[0], y + p[1], z + p[2], t + p[3]};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    c[i] %= g.L[i];
    if (c[i] < 0) c[i] += g.L[i];
  }
  return (size_t)((c[0] + c[1] + c[2] + c[3]) & 1) * g.half4 + cb_index(g, c[0], c[1], c[2], c[3]);
}

// NV complex components of a vector field (12: spin-colour, 3: colour), NV / 3 colour vectors per site
template <typename T, int NV>
__device__ __forceinline__ void load_vec(const T* base, size_t ns, size_t site, int cpb, T (&v)[2 * NV]) {
#pragma unroll
  for (int c = 0; c < NV; c++) {
    size_t o = elem_offset<T>(ns, site, c, cpb);
    v[2 * c] = base[o];
    v[2 * c + 1] = base[o + 1];
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(128) k_stencil_mv(Geom g, int nblocks, int block_size, const int* __restrict__ points,
                                                   const StCode* __restrict__ code, const StFactor* __restrict__ fac, size_t ns,
                                                   int cpb_v, void* const* __restrict__ fields, int n_m) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= ns * nblocks) return;
  const size_t site = tid % ns;
  const int blk = (int)(tid / ns);
  const int par = site >= (size_t)g.half4 ? 1 : 0;
  int x, y, z, t;
  cb_coords(g, par, (int)(site - (size_t)par * g.half4), x, y, z, t);
  for (int line = blk * block_size; line < (blk + 1) * block_size; line++) {
    const StCode c = code[line];
    T v[2 * NV];
    load_vec<T, NV>((const T*)fields[n_m + c.source], ns, shifted_site(g, x, y, z, t, points + 4 * c.source_point), cpb_v, v);
    for (int j = c.nfac - 1; j >= 0; j--) {
      const StFactor f = fac[c.fac0 + j];
      const T* M = (const T*)fields[f.index];
      const size_t ms = shifted_site(g, x, y, z, t, points + 4 * f.point);
      T m[18];
#pragma unroll
      for (int k = 0; k < 9; k++) {
        size_t o = elem_offset<T>(ns, ms, k, 1);
        m[2 * k] = M[o];
        m[2 * k + 1] = M[o + 1];
      }
      T w[2 * NV];
#pragma unroll
      for (int sp = 0; sp < NV / 3; sp++)
#pragma unroll
        for (int r = 0; r < 3; r++) {
          T re = 0, im = 0;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const T mr = f.adj ? m[2 * (3 * k + r)] : m[2 * (3 * r + k)];
            const T mi = f.adj ? -m[2 * (3 * k + r) + 1] : m[2 * (3 * r + k) + 1];
            const T vr = v[2 * (3 * sp + k)], vi = v[2 * (3 * sp + k) + 1];
            re += mr * vr - mi * vi;
            im += mr * vi + mi * vr;
          }
          w[2 * (3 * sp + r)] = re;
          w[2 * (3 * sp + r) + 1] = im;
        }
#pragma unroll
      for (int k = 0; k < 2 * NV; k++) v[k] = w[k];
    }
    const T wr = (T)c.wr, wi = (T)c.wi;
    T* dst = (T*)fields[n_m + c.target];
    const T* acc = c.accumulate >= 0 ? (const T*)fields[n_m + c.accumulate] : 0;
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const size_t o = elem_offset<T>(ns, site, k, cpb_v);
      T re = wr * v[2 * k] - wi * v[2 * k + 1], im = wr * v[2 * k + 1] + wi * v[2 * k];
      if (acc) {
        re += acc[o];
        im += acc[o + 1];
      }
      dst[o] = re;
      dst[o + 1] = im;
    }
  }
}

}  // namespace cgptb

using namespace cgptb;

extern "C" {

int cgptb_stencil_matrix_vector_create(cgptb_stencil_mv** out, const int dims4[4], int precision, int n_points, const int* points,
                                       int n_code, const int* code_ints, const double* weights_re_im, const int* factors,
                                       int code_parallel_block_size, int local, int matrix_parity, int vector_parity) {
  CGPTB_API_BEGIN
  (void)local;
  if (matrix_parity != 0 || vector_parity != 0) CGPTB_ERR("stencil_matrix_vector: checkerboarded fields are not supported (parity %d, %d)", matrix_parity, vector_parity);
  if (g_comm.active) CGPTB_ERR("stencil_matrix_vector runs on one rank only");
  if (n_code < 1 || code_parallel_block_size < 1 || n_code % code_parallel_block_size) CGPTB_ERR("stencil_matrix_vector: %d code lines are not a multiple of the block size %d", n_code, code_parallel_block_size);
  cgptb_stencil_mv* s = new cgptb_stencil_mv();
  for (int i = 0; i < 4; i++) s->dims4[i] = dims4[i];
  s->prec = precision;
  s->npoints = n_points;
  s->ncode = n_code;
  s->block_size = code_parallel_block_size;
  s->nblocks = n_code / code_parallel_block_size;
  std::vector<StCode> code(n_code);
  std::vector<StFactor> fac;
  s->max_m = -1;
  s->max_v = -1;
  for (int i = 0; i < n_code; i++) {
    const int* c = code_ints + 5 * i;  // target, accumulate, source, source_point, number of factors
    code[i].target = c[0];
    code[i].accumulate = c[1];
    code[i].source = c[2];
    code[i].source_point = c[3];
    code[i].nfac = c[4];
    code[i].fac0 = (int)fac.size();
    code[i].wr = weights_re_im[2 * i];
    code[i].wi = weights_re_im[2 * i + 1];
    if (c[0] < 0 || c[2] < 0 || c[3] < 0 || c[3] >= n_points || c[4] < 0) {
      delete s;
      CGPTB_ERR("stencil_matrix_vector: bad code line %d", i);
    }
    s->max_v = std::max(s->max_v, std::max(c[0], std::max(c[1], c[2])));
    for (int j = 0; j < c[4]; j++) {
      StFactor f = {factors[3 * fac.size()], factors[3 * fac.size() + 1], factors[3 * fac.size() + 2]};
      if (f.index < 0 || f.point < 0 || f.point >= n_points) {
        delete s;
        CGPTB_ERR("stencil_matrix_vector: bad factor in code line %d", i);
      }
      s->max_m = std::max(s->max_m, f.index);
      fac.push_back(f);
    }
  }
  CUDA_CHECK(cudaMalloc(&s->d_points, sizeof(int) * 4 * n_points));
  CUDA_CHECK(cudaMemcpy(s->d_points, points, sizeof(int) * 4 * n_points, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMalloc(&s->d_code, sizeof(StCode) * n_code));
  CUDA_CHECK(cudaMemcpy(s->d_code, code.data(), sizeof(StCode) * n_code, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMalloc(&s->d_fac, sizeof(StFactor) * (fac.size() + 1)));
  if (!fac.empty()) CUDA_CHECK(cudaMemcpy(s->d_fac, fac.data(), sizeof(StFactor) * fac.size(), cudaMemcpyHostToDevice));
  *out = s;
  CGPTB_API_END
}

int cgptb_stencil_matrix_vector_execute(cgptb_stencil_mv* s, const cgptb_lattice* const* matrix_fields, int n_m,
                                        cgptb_lattice* const* vector_fields, int n_v, int fast_osites) {
  CGPTB_API_BEGIN
  (void)fast_osites;  // loop-order hint of the reference (MAP_INDEXING); the site index is the fast one here
  if (n_m <= s->max_m || n_v <= s->max_v) CGPTB_ERR("stencil_matrix_vector: the code refers to matrix field %d / vector field %d, got %d / %d fields", s->max_m, s->max_v, n_m, n_v);
  if (n_v < 1) CGPTB_ERR("stencil_matrix_vector: no vector fields");
  const cgptb_lattice* v0 = vector_fields[0];
  for (int i = 0; i < n_m + n_v; i++) {
    const cgptb_lattice* l = i < n_m ? matrix_fields[i] : vector_fields[i - n_m];
    if (!l) CGPTB_ERR("stencil_matrix_vector: field %d is missing", i);
    if (l->cb != CGPTB_FULL || l->Ls != 0 || l->prec != s->prec) CGPTB_ERR("stencil_matrix_vector: fields must live on the full 4d grid in the stencil's precision");
    for (int d = 0; d < 4; d++)
      if (l->dims4[d] != s->dims4[d]) CGPTB_ERR("stencil_matrix_vector: field %d lives on a different grid", i);
    if (i < n_m && l->otype != CGPTB_OT_MCOLOR) CGPTB_ERR("stencil_matrix_vector: matrix fields must be colour matrices");
    if (i >= n_m && l->otype != v0->otype) CGPTB_ERR("stencil_matrix_vector: vector fields of different types");
  }
  if (v0->otype != 12 && v0->otype != 3) CGPTB_ERR("stencil_matrix_vector: vector fields must be spin-colour or colour vectors");
  if (s->cap_fields < n_m + n_v) {
    if (s->d_fields) CUDA_CHECK(cudaFree(s->d_fields));
    CUDA_CHECK(cudaMalloc(&s->d_fields, sizeof(void*) * (n_m + n_v)));
    s->cap_fields = n_m + n_v;
  }
  std::vector<void*> ptrs(n_m + n_v);
  for (int i = 0; i < n_m; i++) ptrs[i] = matrix_fields[i]->data;
  for (int i = 0; i < n_v; i++) ptrs[n_m + i] = vector_fields[i]->data;
  CUDA_CHECK(cudaMemcpyAsync(s->d_fields, ptrs.data(), sizeof(void*) * ptrs.size(), cudaMemcpyHostToDevice, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));  // ptrs is a stack object
  Geom g = make_geom(s->dims4);
  const size_t ns = v0->sites;
  const size_t nthreads = ns * s->nblocks;
  const unsigned blocks = (unsigned)((nthreads + 127) / 128);
#define ST_LAUNCH(T_, NV_)                                                                                                      \
  k_stencil_mv<T_, NV_><<<blocks, 128, 0, g_stream>>>(g, s->nblocks, s->block_size, s->d_points, s->d_code, s->d_fac, ns, v0->cpb(), \
                                                     s->d_fields, n_m)
  if (s->prec == CGPTB_SINGLE) {
    if (v0->otype == 12)
      ST_LAUNCH(float, 12);
    else
      ST_LAUNCH(float, 3);
  } else {
    if (v0->otype == 12)
      ST_LAUNCH(double, 12);
    else
      ST_LAUNCH(double, 3);
  }
#undef ST_LAUNCH
  LAUNCH_CHECK();
  CGPTB_API_END
}

int cgptb_stencil_matrix_vector_delete(cgptb_stencil_mv* s) {
  CGPTB_API_BEGIN
  if (s) {
    if (s->d_points) cudaFree(s->d_points);
    if (s->d_code) cudaFree(s->d_code);
    if (s->d_fac) cudaFree(s->d_fac);
    if (s->d_fields) cudaFree(s->d_fields);
    delete s;
  }
  CGPTB_API_END
}
}

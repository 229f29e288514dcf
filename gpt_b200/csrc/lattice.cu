// Runtime + lattice storage seam of libcgpt_b200 (replaces lib/cgpt/lib/lattice.cc:42-210,
// lib/cgpt/lib/transform.cc:41-54,128-141 and lattice/implementation.h:246-281 for the hot path's objects).
#include <map>
#include <stdlib.h>
#include "common.cuh"

namespace cgptb {

// Device memory of deleted lattices is kept for the next lattice of the same size instead of going back to the driver: the host
// layer creates and drops temporaries all the time (every g(expr), every outer iteration of a defect-correcting solve), and
// cudaFree of a GB-sized block costs tens to hundreds of milliseconds and synchronises the device (the reference relies on Grid's
// memory manager cache for the same reason).  Reuse is stream-ordered like everything else in the library.  CGPTB_CACHE_GB
// bounds what is kept (default 24; 0 switches the cache off).
static std::multimap<size_t, void*> g_lat_cache;
static size_t g_lat_cached = 0;
static size_t lattice_cache_limit() {
  static double gb = getenv("CGPTB_CACHE_GB") ? atof(getenv("CGPTB_CACHE_GB")) : 24.0;
  return (size_t)(gb * 1e9);
}
static void* lattice_cache_get(size_t bytes) {
  auto it = g_lat_cache.find(bytes);
  if (it == g_lat_cache.end()) return 0;
  void* p = it->second;
  g_lat_cache.erase(it);
  g_lat_cached -= bytes;
  return p;
}
static void lattice_cache_flush() {
  for (auto& kv : g_lat_cache) cudaFree(kv.second);
  g_lat_cache.clear();
  g_lat_cached = 0;
}
static void lattice_cache_put(void* p, size_t bytes) {
  if (g_lat_cached + bytes <= lattice_cache_limit()) {
    g_lat_cache.emplace(bytes, p);
    g_lat_cached += bytes;
  } else {
    CUDA_CHECK(cudaFree(p));
  }
}

thread_local std::string g_error;
cudaStream_t g_stream = 0;
static cudaStream_t g_own_stream = 0;
uint64_t g_launches = 0;
static int g_device = -1;
static int g_sm_count = 0;
static cudaEvent_t g_ev0 = 0, g_ev1 = 0;
static double* g_scratch = 0;
static size_t g_scratch_n = 0;
static double* g_hscratch = 0;
static size_t g_hscratch_n = 0;
static void* g_stage = 0;
static size_t g_stage_n = 0;

int sm_count() { return g_sm_count; }

double* reduce_scratch(size_t n) {
  if (n > g_scratch_n) {
    if (g_scratch) CUDA_CHECK(cudaFree(g_scratch));
    CUDA_CHECK(cudaMalloc(&g_scratch, n * sizeof(double)));
    g_scratch_n = n;
  }
  return g_scratch;
}

double* reduce_host(size_t n) {
  if (n > g_hscratch_n) {
    if (g_hscratch) CUDA_CHECK(cudaFreeHost(g_hscratch));
    CUDA_CHECK(cudaMallocHost(&g_hscratch, n * sizeof(double)));
    g_hscratch_n = n;
  }
  return g_hscratch;
}

static void* stage(size_t bytes) {
  if (bytes > g_stage_n) {
    if (g_stage) CUDA_CHECK(cudaFree(g_stage));
    CUDA_CHECK(cudaMalloc(&g_stage, bytes));
    g_stage_n = bytes;
  }
  return g_stage;
}

static void require_init() {
  if (g_device < 0) CGPTB_ERR("cgptb_init() has not been called (no CUDA device selected)");
}

// device stored site j  ->  host (GPT order) site index
__device__ __forceinline__ size_t host_site(const Geom& g, int cb, int ls, size_t j) {
  if (cb != CGPTB_FULL) return j;
  size_t per = (size_t)g.half4 * ls;
  int p = j >= per ? 1 : 0;
  size_t rem = j - (size_t)p * per;
  int i4 = (int)(rem / ls);
  int s = (int)(rem - (size_t)i4 * ls);
  int x, y, z, t;
  cb_coords(g, p, i4, x, y, z, t);
  size_t lex4 = x + (size_t)g.L[0] * (y + (size_t)g.L[1] * (z + (size_t)g.L[2] * t));
  return s + (size_t)ls * lex4;
}

// TH: host real type, TD: device real type
template <typename TH, typename TD, bool IMPORT>
__global__ void k_layout(Geom g, int cb, int ls, int otype, int cpb, size_t nsites, TD* dev, TH* host) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = nsites * otype;
  if (idx >= n) return;
  // consecutive threads walk consecutive device sites of one component block -> coalesced device side
  int c = (int)(idx / nsites);
  size_t j = idx - (size_t)c * nsites;
  size_t hs = host_site(g, cb, ls, j);
  size_t ho = (hs * otype + c) * 2;
  size_t d = elem_offset<TD>(nsites, j, c, cpb);
  if (IMPORT) {
    dev[d] = (TD)host[ho];
    dev[d + 1] = (TD)host[ho + 1];
  } else {
    host[ho] = (TH)dev[d];
    host[ho + 1] = (TH)dev[d + 1];
  }
}

template <typename TS, typename TD>
__global__ void k_convert(int otype, int cpb_s, int cpb_d, size_t nsites, const TS* src, TD* dst) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = nsites * otype;
  if (idx >= n) return;
  int c = (int)(idx / nsites);
  size_t j = idx - (size_t)c * nsites;
  size_t so = elem_offset<TS>(nsites, j, c, cpb_s);
  size_t d = elem_offset<TD>(nsites, j, c, cpb_d);
  dst[d] = (TD)src[so];
  dst[d + 1] = (TD)src[so + 1];
}

template <bool IMPORT>
static void layout_device(cgptb_lattice* l, void* devbuf) {
  Geom g = make_geom(l->dims4);
  size_t n = l->sites * l->otype;
  int threads = 256;
  unsigned blocks = (unsigned)((n + threads - 1) / threads);
  if (l->prec == CGPTB_SINGLE)
    k_layout<float, float, IMPORT><<<blocks, threads, 0, g_stream>>>(g, l->cb, l->ls(), l->otype, l->cpb(), l->sites,
                                                                      (float*)l->data, (float*)devbuf);
  else
    k_layout<double, double, IMPORT><<<blocks, threads, 0, g_stream>>>(g, l->cb, l->ls(), l->otype, l->cpb(), l->sites,
                                                                        (double*)l->data, (double*)devbuf);
  LAUNCH_CHECK();
}

void blas_copy(cgptb_lattice* d, const cgptb_lattice* s) {
  CGPTB_ASSERT(same_shape(d, s));
  d->cb = s->cb;
  CUDA_CHECK(cudaMemcpyAsync(d->data, s->data, s->bytes(), cudaMemcpyDeviceToDevice, g_stream));
}

void blas_zero(cgptb_lattice* d) { CUDA_CHECK(cudaMemsetAsync(d->data, 0, d->bytes(), g_stream)); }

static void new_lattice(cgptb_lattice** out, const int dims4[4], int Ls, int precision, int otype, int cb, void* ptr) {
  require_init();
  CGPTB_ASSERT(precision == CGPTB_SINGLE || precision == CGPTB_DOUBLE);
  CGPTB_ASSERT(cb == CGPTB_EVEN || cb == CGPTB_ODD || cb == CGPTB_FULL);
  CGPTB_ASSERT(otype >= 1 && Ls >= 0);
  for (int i = 0; i < 4; i++)
    if (dims4[i] < 2 || dims4[i] % 2) CGPTB_ERR("lattice extent %d of dimension %d must be even and >= 2", dims4[i], i);
  cgptb_lattice* l = new cgptb_lattice;
  l->prec = precision;
  l->otype = otype;
  for (int i = 0; i < 4; i++) l->dims4[i] = dims4[i];
  l->Ls = Ls;
  l->cb = cb;
  size_t v4 = (size_t)dims4[0] * dims4[1] * dims4[2] * dims4[3];
  l->sites4 = cb == CGPTB_FULL ? v4 : v4 / 2;
  l->sites = l->sites4 * (size_t)(Ls > 0 ? Ls : 1);
  l->owns = ptr == 0;
  l->data = ptr;
  if (!ptr) {
    // Optional skew of successive allocations against each other (CGPTB_LATTICE_SKEW bytes, a multiple of 256, times a counter
    // that cycles over 16 positions; default 0 = off).  Tried against the bimodal CG of round 1 / 2 on the theory that equal-shape
    // fields 2 MB-multiples apart alias in L2 / DRAM; the cause turned out to be the per-solve cudaMalloc / cudaFree of the
    // solver's work fields (solver.cu), so this stays a tuning knob only.
    static size_t skew_unit = getenv("CGPTB_LATTICE_SKEW") ? (size_t)atol(getenv("CGPTB_LATTICE_SKEW")) / 256 * 256 : 0;
    static unsigned counter = 0;
    const size_t skew = skew_unit * (counter++ % 16);
    l->alloc_bytes = l->bytes() + skew_unit * 16;
    l->alloc = lattice_cache_get(l->alloc_bytes);
    cudaError_t e = l->alloc ? cudaSuccess : cudaMalloc(&l->alloc, l->alloc_bytes);
    if (e != cudaSuccess) {  // give the cached blocks back and try once more
      cudaGetLastError();
      lattice_cache_flush();
      e = cudaMalloc(&l->alloc, l->alloc_bytes);
    }
    if (e != cudaSuccess) {
      size_t b = l->bytes();
      delete l;
      CGPTB_ERR("cudaMalloc of %zu bytes failed: %s", b, cudaGetErrorString(e));
    }
    l->data = (char*)l->alloc + skew;
  }
  *out = l;
}

}  // namespace cgptb

using namespace cgptb;

namespace cgptb {
// gpt.pack of n 4d fields into one 5d field with the list index as the fifth dimension and back
// (matrix_operator.packed(), lib/gpt/core/operator/matrix_operator.py:200-239): 5d site = 4d site * n + s in both halves
template <bool PACK>
__global__ void k_pack_rhs(size_t n4, int n, int s, int nblk, size_t blk16, const uint4* __restrict__ a4, uint4* __restrict__ a5,
                           const uint4* __restrict__ a5c, uint4* __restrict__ a4w) {
  // one thread per (component block, 4d site, 16 bytes of the block)
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)nblk * n4 * blk16) return;
  const size_t q = idx % blk16;
  const size_t r = idx / blk16;
  const size_t i = r % n4;
  const size_t k = r / n4;
  const size_t o4 = (k * n4 + i) * blk16 + q;
  const size_t o5 = (k * n4 * n + i * n + s) * blk16 + q;
  if (PACK)
    a5[o5] = a4[o4];
  else
    a4w[o4] = a5c[o5];
}

}  // namespace cgptb

extern "C" {

int cgptb_init(int device) {
  CGPTB_API_BEGIN
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    CGPTB_ERR("no CUDA device available (%s); libcgpt_b200 has no CPU fallback", cudaGetErrorString(e));
  CGPTB_ASSERT(device >= 0 && device < n);
  CUDA_CHECK(cudaSetDevice(device));
  if (g_device != device) {
    g_device = device;
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    g_sm_count = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&g_own_stream, cudaStreamNonBlocking));
    g_stream = g_own_stream;
    CUDA_CHECK(cudaEventCreate(&g_ev0));
    CUDA_CHECK(cudaEventCreate(&g_ev1));
  }
  CGPTB_API_END
}

const char* cgptb_last_error(void) { return g_error.c_str(); }

int cgptb_accelerator_barrier(void) {
  CGPTB_API_BEGIN
  require_init();
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  CGPTB_API_END
}

int cgptb_set_stream(void* s) {
  CGPTB_API_BEGIN
  require_init();
  g_stream = s ? (cudaStream_t)s : g_own_stream;
  CGPTB_API_END
}

void* cgptb_get_stream(void) { return (void*)g_stream; }

int cgptb_timer_start(void) {
  CGPTB_API_BEGIN
  require_init();
  CUDA_CHECK(cudaEventRecord(g_ev0, g_stream));
  CGPTB_API_END
}

int cgptb_timer_stop(double* ms) {
  CGPTB_API_BEGIN
  require_init();
  CUDA_CHECK(cudaEventRecord(g_ev1, g_stream));
  CUDA_CHECK(cudaEventSynchronize(g_ev1));
  float f = 0;
  CUDA_CHECK(cudaEventElapsedTime(&f, g_ev0, g_ev1));
  *ms = f;
  CGPTB_API_END
}

int cgptb_device_info(int* sms, size_t* total_mem, int* cc_major, int* cc_minor) {
  CGPTB_API_BEGIN
  require_init();
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, g_device));
  *sms = prop.multiProcessorCount;
  *total_mem = prop.totalGlobalMem;
  *cc_major = prop.major;
  *cc_minor = prop.minor;
  CGPTB_API_END
}

uint64_t cgptb_launch_count(void) { return g_launches; }

int cgptb_create_lattice(cgptb_lattice** out, const int dims4[4], int Ls, int precision, int otype, int cb) {
  CGPTB_API_BEGIN
  new_lattice(out, dims4, Ls, precision, otype, cb, 0);
  CGPTB_API_END
}

int cgptb_create_lattice_view(cgptb_lattice** out, const int dims4[4], int Ls, int precision, int otype, int cb,
                              void* device_ptr) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(device_ptr != 0);
  CGPTB_ASSERT(((uintptr_t)device_ptr & 31) == 0);
  new_lattice(out, dims4, Ls, precision, otype, cb, device_ptr);
  CGPTB_API_END
}

int cgptb_delete_lattice(cgptb_lattice* l) {
  CGPTB_API_BEGIN
  if (l) {
    if (l->owns && l->alloc) lattice_cache_put(l->alloc, l->alloc_bytes);
    delete l;
  }
  CGPTB_API_END
}

size_t cgptb_lattice_bytes(const cgptb_lattice* l) { return l->bytes(); }
size_t cgptb_lattice_sites(const cgptb_lattice* l) { return l->sites; }
void* cgptb_lattice_device_ptr(cgptb_lattice* l) { return l->data; }
int cgptb_lattice_info(const cgptb_lattice* l, int dims4[4], int* Ls, int* precision, int* otype, int* cb) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(l);
  for (int i = 0; i < 4; i++) dims4[i] = l->dims4[i];
  *Ls = l->Ls;
  *precision = l->prec;
  *otype = l->otype;
  *cb = l->cb;
  CGPTB_API_END
}
int cgptb_lattice_get_checkerboard(const cgptb_lattice* l) { return l->cb; }

int cgptb_lattice_change_checkerboard(cgptb_lattice* l, int cb) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(l->cb != CGPTB_FULL && (cb == CGPTB_EVEN || cb == CGPTB_ODD));
  l->cb = cb;
  CGPTB_API_END
}

int cgptb_lattice_set_to_zero(cgptb_lattice* l) {
  CGPTB_API_BEGIN
  blas_zero(l);
  CGPTB_API_END
}

int cgptb_lattice_import(cgptb_lattice* l, const void* host, size_t nbytes) {
  CGPTB_API_BEGIN
  if (nbytes != l->bytes()) CGPTB_ERR("import: buffer has %zu bytes, lattice needs %zu", nbytes, l->bytes());
  void* st = stage(nbytes);
  CUDA_CHECK(cudaMemcpyAsync(st, host, nbytes, cudaMemcpyHostToDevice, g_stream));
  layout_device<true>(l, st);
  CGPTB_API_END
}

int cgptb_lattice_export(const cgptb_lattice* l, void* host, size_t nbytes) {
  CGPTB_API_BEGIN
  if (nbytes != l->bytes()) CGPTB_ERR("export: buffer has %zu bytes, lattice needs %zu", nbytes, l->bytes());
  void* st = stage(nbytes);
  layout_device<false>(const_cast<cgptb_lattice*>(l), st);
  CUDA_CHECK(cudaMemcpyAsync(host, st, nbytes, cudaMemcpyDeviceToHost, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  CGPTB_API_END
}

int cgptb_lattice_import_device(cgptb_lattice* l, const void* dev, size_t nbytes) {
  CGPTB_API_BEGIN
  if (nbytes != l->bytes()) CGPTB_ERR("import: buffer has %zu bytes, lattice needs %zu", nbytes, l->bytes());
  layout_device<true>(l, const_cast<void*>(dev));
  CGPTB_API_END
}

int cgptb_lattice_export_device(const cgptb_lattice* l, void* dev, size_t nbytes) {
  CGPTB_API_BEGIN
  if (nbytes != l->bytes()) CGPTB_ERR("export: buffer has %zu bytes, lattice needs %zu", nbytes, l->bytes());
  layout_device<false>(const_cast<cgptb_lattice*>(l), dev);
  CGPTB_API_END
}

int cgptb_lattice_pack_rhs(cgptb_lattice* l5, cgptb_lattice* const* l4, int n, int unpack) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(l5 && l4 && n > 0 && l5->Ls == n);
  for (int s = 0; s < n; s++) {
    const cgptb_lattice* a = l4[s];
    CGPTB_ASSERT(a && a->Ls == 0 && a->prec == l5->prec && a->otype == l5->otype && a->sites * n == l5->sites);
    for (int i = 0; i < 4; i++) CGPTB_ASSERT(a->dims4[i] == l5->dims4[i]);
    if (unpack) {
      l4[s]->cb = l5->cb;
    } else {
      if (s == 0) l5->cb = a->cb;
      CGPTB_ASSERT(a->cb == l5->cb);
    }
    const size_t blk16 = a->block_bytes() / 16;
    CGPTB_ASSERT(a->block_bytes() % 16 == 0);
    const int nblk = a->otype / a->cpb();
    const size_t nthr = (size_t)nblk * a->sites * blk16;
    const unsigned blocks = (unsigned)((nthr + 255) / 256);
    if (unpack)
      k_pack_rhs<false><<<blocks, 256, 0, g_stream>>>(a->sites, n, s, nblk, blk16, 0, 0, (const uint4*)l5->data, (uint4*)l4[s]->data);
    else
      k_pack_rhs<true><<<blocks, 256, 0, g_stream>>>(a->sites, n, s, nblk, blk16, (const uint4*)a->data, (uint4*)l5->data, 0, 0);
    LAUNCH_CHECK();
  }
  CGPTB_API_END
}

int cgptb_lattice_copy(cgptb_lattice* dst, const cgptb_lattice* src) {
  CGPTB_API_BEGIN
  blas_copy(dst, src);
  CGPTB_API_END
}

int cgptb_lattice_convert(cgptb_lattice* dst, const cgptb_lattice* src) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(dst->otype == src->otype && dst->sites == src->sites && dst->Ls == src->Ls);
  dst->cb = src->cb;
  if (dst->prec == src->prec) {
    blas_copy(dst, src);
  } else {
    size_t n = src->sites * src->otype;
    int threads = 256;
    unsigned blocks = (unsigned)((n + threads - 1) / threads);
    if (src->prec == CGPTB_DOUBLE)
      k_convert<double, float><<<blocks, threads, 0, g_stream>>>(src->otype, src->cpb(), dst->cpb(), src->sites,
                                                                 (const double*)src->data, (float*)dst->data);
    else
      k_convert<float, double><<<blocks, threads, 0, g_stream>>>(src->otype, src->cpb(), dst->cpb(), src->sites,
                                                                 (const float*)src->data, (double*)dst->data);
    LAUNCH_CHECK();
  }
  CGPTB_API_END
}

// a half of a full lattice is, per 16-byte component plane, one contiguous run of `half` sites
static void copy_half(void* full, void* half, const cgptb_lattice* lf, int cb, bool to_full) {
  size_t blk = 2 * lf->cpb() * lf->real_size();  // bytes per block
  size_t nblocks = (size_t)lf->otype / lf->cpb();
  size_t hs = lf->sites / 2;
  char* f = (char*)full + (size_t)cb * hs * blk;
  if (to_full)
    CUDA_CHECK(cudaMemcpy2DAsync(f, lf->sites * blk, half, hs * blk, hs * blk, nblocks, cudaMemcpyDeviceToDevice, g_stream));
  else
    CUDA_CHECK(cudaMemcpy2DAsync(half, hs * blk, f, lf->sites * blk, hs * blk, nblocks, cudaMemcpyDeviceToDevice, g_stream));
}

int cgptb_lattice_pick_checkerboard(int cb, cgptb_lattice* half, const cgptb_lattice* full) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(full->cb == CGPTB_FULL && half->cb != CGPTB_FULL && (cb == CGPTB_EVEN || cb == CGPTB_ODD));
  CGPTB_ASSERT(half->prec == full->prec && half->otype == full->otype && half->Ls == full->Ls &&
               half->sites * 2 == full->sites);
  half->cb = cb;
  copy_half(full->data, half->data, full, cb, false);
  CGPTB_API_END
}

int cgptb_lattice_set_checkerboard(cgptb_lattice* full, const cgptb_lattice* half) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(full->cb == CGPTB_FULL && half->cb != CGPTB_FULL);
  CGPTB_ASSERT(half->prec == full->prec && half->otype == full->otype && half->Ls == full->Ls &&
               half->sites * 2 == full->sites);
  copy_half(full->data, half->data, full, half->cb, true);
  CGPTB_API_END
}
}

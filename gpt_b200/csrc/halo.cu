// Halo exchange of the hopping term across the processor grid: the 4d lattice is split along
// T, then Z (then Y); per Dslash each split direction exchanges one face of spin-projected half-spinors
// (12 reals per 5d site) with each neighbour:
//
// Default on one box (all ranks can map each other's memory, CUDA IPC): everything on the compute stream,
//
//   pack faces, storing them INTO the neighbours' receive buffers over NVLink -> cuStreamWriteValue32 of the call number into
//   the neighbours' flag words -> interior stencil (all local hops, all SMs) -> cuStreamWaitValue32 on the own flag words ->
//   exterior update
//
// so the transfer is the pack kernel's own stores, no communication kernel takes SMs from the persistent stencil and no
// second stream or event is involved.  Receive buffers alternate with the call number; a buffer is safe to overwrite two
// calls later because the exchange is symmetric: the neighbour's data of call n+1 -- which this rank waits for before it
// can pack call n+2 -- was packed after the neighbour's exterior kernel of call n in its stream.
// Fallback (CGPTB_HALO=nccl, or memory that cannot be mapped):
//
//   compute stream : pack faces -> [event] -> interior stencil (all local hops) -> wait -> exterior update
//   comm stream    :               wait -> NCCL send/recv group over NVLink -> [event]
//
// Either way the transfer is hidden behind the interior kernel.  Links on the low face (U_mu(x-mu) of the
// neighbour) are fetched once at gauge import, so the receiver does the SU(3) multiply for both faces.
// This is what Grid's CartesianStencil::HaloExchange + overlapCommsCompute do for the reference
// (lib/cgpt/lib/operators/mobius.h:52, wilson_clover.h:45).
#include <string.h>
#include <utility>
#include <vector>
#include "dslash.cuh"
#include "operator.cuh"

namespace cgptb {

// transverse geometry of the face orthogonal to mu (mu = 1,2,3): coordinates (xh | x, a, b)
struct FaceGeom {
  int mu;
  int da, db;  // which of (y,z,t) are a and b: indices into L
  int A, B;    // extents
};

static FaceGeom make_face(const Geom& g, int mu) {
  FaceGeom f;
  f.mu = mu;
  int o[2], n = 0;
  for (int d = 1; d < 4; d++)
    if (d != mu) o[n++] = d;
  f.da = o[0];
  f.db = o[1];
  f.A = g.L[o[0]];
  f.B = g.L[o[1]];
  return f;
}

// face index f of parity q at x_mu = c  ->  coordinates and checkerboard index
__device__ __forceinline__ int face_site(const Geom& g, const FaceGeom& fg, int q, int c, int f, int& x, int& y, int& z, int& t) {
  int xh = f % g.hx;
  int r = f / g.hx;
  int a = r % fg.A, b = r / fg.A;
  int co[4] = {0, 0, 0, 0};
  co[fg.mu] = c;
  co[fg.da] = a;
  co[fg.db] = b;
  y = co[1];
  z = co[2];
  t = co[3];
  x = 2 * xh + ((y + z + t + q) & 1);
  return cb_index(g, x, y, z, t);
}

template <typename T>
struct HS;  // half-spinor buffer element access: SoA of 16-byte blocks, 12 reals per site
template <>
struct HS<float> {
  static __device__ __forceinline__ void store(float* b, size_t n, size_t i, const float (&h)[12]) {
    float4* p = reinterpret_cast<float4*>(b);
#pragma unroll
    for (int k = 0; k < 3; k++) p[k * n + i] = make_float4(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
  }
  static __device__ __forceinline__ void load(const float* b, size_t n, size_t i, float (&h)[12]) {
    const float4* p = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float4 v = p[k * n + i];
      h[4 * k] = v.x; h[4 * k + 1] = v.y; h[4 * k + 2] = v.z; h[4 * k + 3] = v.w;
    }
  }
};
template <>
struct HS<double> {
  static __device__ __forceinline__ void store(double* b, size_t n, size_t i, const double (&h)[12]) {
    double2* p = reinterpret_cast<double2*>(b);
#pragma unroll
    for (int k = 0; k < 6; k++) p[k * n + i] = make_double2(h[2 * k], h[2 * k + 1]);
  }
  static __device__ __forceinline__ void load(const double* b, size_t n, size_t i, double (&h)[12]) {
    const double2* p = reinterpret_cast<const double2*>(b);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double2 v = p[k * n + i];
      h[2 * k] = v.x; h[2 * k + 1] = v.y;
    }
  }
};

// pack both faces of direction MU: lo face (x_mu = 0) projected for the neighbour's forward hop,
// hi face (x_mu = L-1) projected for the neighbour's backward hop
template <typename T, int MU, bool DAG>
__global__ void __launch_bounds__(128) k_pack(Geom g, FaceGeom fg, int ls, int q_in, const T* __restrict__ in, size_t in_stride,
                                             T* __restrict__ to_lo, T* __restrict__ to_hi, size_t nface) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= 2 * nface) return;
  int side = tid >= nface ? 1 : 0;
  size_t idx = tid - (size_t)side * nface;
  int f = (int)(idx / ls), s = (int)(idx - (size_t)f * ls);
  int x, y, z, t;
  int i4 = face_site(g, fg, q_in, side ? g.L[MU] - 1 : 0, f, x, y, z, t);
  T psi[24], h[12];
  load_spinor(in, in_stride, (size_t)i4 * ls + s, psi);
  if (side == 0) {
    project<MU, DAG ? +1 : -1>(psi, h);  // forward hop of the receiver
    HS<T>::store(to_lo, nface, idx, h);
  } else {
    project<MU, DAG ? -1 : +1>(psi, h);  // backward hop of the receiver
    HS<T>::store(to_hi, nface, idx, h);
  }
}

// add the two off-rank hops of direction MU to the boundary sites of the output parity
template <typename T, int MU, bool DAG>
__global__ void __launch_bounds__(128) k_exterior(Geom g, FaceGeom fg, int ls, int p_out, T* __restrict__ out, size_t out_stride,
                                                 const T* __restrict__ from_lo, const T* __restrict__ from_hi, size_t nface,
                                                 const T* __restrict__ links) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= 2 * nface) return;
  int side = tid >= nface ? 1 : 0;
  size_t idx = tid - (size_t)side * nface;
  int f = (int)(idx / ls), s = (int)(idx - (size_t)f * ls);
  int x, y, z, t;
  int i4 = face_site(g, fg, p_out, side ? g.L[MU] - 1 : 0, f, x, y, z, t);
  T h[12], chi[12], W[18], acc[24];
  load_spinor_rw(out, out_stride, (size_t)i4 * ls + s, acc);
  if (side == 1) {  // forward neighbour lives on rank+mu
    HS<T>::load(from_hi, nface, idx, h);
    load_link<T>(links, (size_t)i4, MU, W);
    su3_mul<false>(W, h, chi);
    reconstruct_add<MU, DAG ? +1 : -1>(acc, chi);
  } else {  // backward neighbour lives on rank-mu
    HS<T>::load(from_lo, nface, idx, h);
    load_link<T>(links, (size_t)i4, MU + 4, W);
    su3_mul<true>(W, h, chi);
    reconstruct_add<MU, DAG ? -1 : +1>(acc, chi);
  }
  store_spinor(out, out_stride, (size_t)i4 * ls + s, acc);
}

// high-face links U_mu(x_mu = L-1) of the full lattice -> [transverse lex index][9 complex] double
template <typename TU>
__global__ void k_pack_links(Geom g, FaceGeom fg, size_t nsU, const TU* __restrict__ U, double* __restrict__ buf) {
  int ft = blockIdx.x * blockDim.x + threadIdx.x;
  int n = g.L[0] * fg.A * fg.B;
  if (ft >= n) return;
  int x = ft % g.L[0], r = ft / g.L[0];
  int co[4] = {x, 0, 0, 0};
  co[fg.mu] = g.L[fg.mu] - 1;
  co[fg.da] = r % fg.A;
  co[fg.db] = r / fg.A;
  int p = (co[0] + co[1] + co[2] + co[3]) & 1;
  size_t site = (size_t)p * g.half4 + cb_index(g, co[0], co[1], co[2], co[3]);
  for (int k = 0; k < 9; k++) {
    size_t o = elem_offset<TU>(nsU, site, k, 1);
    buf[((size_t)ft * 9 + k) * 2] = U[o];
    buf[((size_t)ft * 9 + k) * 2 + 1] = U[o + 1];
  }
}

// Peer-to-peer state of an operator: an arena in this rank's memory that the neighbours write into -- flag words [mu][side]
// followed by the receive buffers [mu][side][call parity] --, and where the same things live in the neighbours' arenas.
// Arenas are pooled and never freed (the neighbours keep them mapped).
struct HaloP2P {
  char* arena = 0;
  size_t arena_bytes = 0;
  size_t off_recv[4][2][2];
  char* peer[4][2];  // arena of the neighbour at -mu (0) / +mu (1), mapped into this process
  unsigned seq = 0;  // number of exchanges so far
  static size_t flag_off(int mu, int side) { return (size_t)(mu * 2 + side) * 128; }
};
struct ArenaPool {
  std::vector<std::pair<char*, size_t>> free_list;
};
static ArenaPool g_arenas;

bool halo_is_p2p(const cgptb_fermion_operator* op) { return op->p2p != 0; }

// CGPTB_HALO_PACK=direct: the pack kernel stores the faces straight into the neighbours' memory (one stream, no copies; the
// kernel then runs at NVLink speed, 46 us for 2 x 9.4 MB to one peer); default: pack locally, copy engines move the faces
static bool p2p_direct_stores() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CGPTB_HALO_PACK");
    v = e && !strcmp(e, "direct") ? 1 : 0;
  }
  return v == 1;
}

template <typename T, int MU>
static void pack_t(cgptb_fermion_operator* op, bool dag, int q_in, const T* in, size_t in_stride) {
  FaceGeom fg = make_face(op->g, MU);
  int ls = op->ls();
  size_t nface = (size_t)(op->g.half4 / op->g.L[MU]) * ls;
  unsigned blocks = (unsigned)((2 * nface + 127) / 128);
  T *to_lo = (T*)op->halo_send[MU][0], *to_hi = (T*)op->halo_send[MU][1];
  if (op->p2p && p2p_direct_stores()) {
    // my low face is the -mu neighbour's "from_hi" (side 1), my high face the +mu neighbour's "from_lo" (side 0)
    HaloP2P* h = (HaloP2P*)op->p2p;
    to_lo = (T*)(h->peer[MU][0] + h->off_recv[MU][1][h->seq & 1]);
    to_hi = (T*)(h->peer[MU][1] + h->off_recv[MU][0][h->seq & 1]);
  }
  if (dag)
    k_pack<T, MU, true><<<blocks, 128, 0, g_stream>>>(op->g, fg, ls, q_in, in, in_stride, to_lo, to_hi, nface);
  else
    k_pack<T, MU, false><<<blocks, 128, 0, g_stream>>>(op->g, fg, ls, q_in, in, in_stride, to_lo, to_hi, nface);
  LAUNCH_CHECK();
}

template <typename T, int MU>
static void exterior_t(cgptb_fermion_operator* op, bool dag, int p_out, T* out, size_t out_stride) {
  FaceGeom fg = make_face(op->g, MU);
  int ls = op->ls();
  size_t nface = (size_t)(op->g.half4 / op->g.L[MU]) * ls;
  unsigned blocks = (unsigned)((2 * nface + 127) / 128);
  const T* links = (const T*)op->links[p_out];
  const T *from_lo = (const T*)op->halo_recv[MU][0], *from_hi = (const T*)op->halo_recv[MU][1];
  if (op->p2p) {
    HaloP2P* h = (HaloP2P*)op->p2p;
    from_lo = (const T*)(h->arena + h->off_recv[MU][0][h->seq & 1]);
    from_hi = (const T*)(h->arena + h->off_recv[MU][1][h->seq & 1]);
  }
  if (dag)
    k_exterior<T, MU, true><<<blocks, 128, 0, g_stream>>>(op->g, fg, ls, p_out, out, out_stride, from_lo, from_hi, nface, links);
  else
    k_exterior<T, MU, false><<<blocks, 128, 0, g_stream>>>(op->g, fg, ls, p_out, out, out_stride, from_lo, from_hi, nface, links);
  LAUNCH_CHECK();
}

template <typename T>
static void halo_begin_t(cgptb_fermion_operator* op, bool dag, int p_out, const T* in, size_t in_stride) {
  if (op->p2p) {
    HaloP2P* h = (HaloP2P*)op->p2p;
    h->seq++;
    const bool direct = p2p_direct_stores();
    // copy-engine variant: the faces are packed into local buffers (HBM speed) and a copy engine pushes them over NVLink
    // while the interior stencil runs; the send buffers are free again when the previous call's copies are done
    if (!direct) CUDA_CHECK(cudaStreamWaitEvent(g_stream, g_comm.ev_comm, 0));
    for (int mu = 1; mu < 4; mu++) {
      if (!((op->g.comm_mask >> mu) & 1)) continue;
      if (mu == 1) pack_t<T, 1>(op, dag, 1 - p_out, in, in_stride);
      if (mu == 2) pack_t<T, 2>(op, dag, 1 - p_out, in, in_stride);
      if (mu == 3) pack_t<T, 3>(op, dag, 1 - p_out, in, in_stride);
    }
    cudaStream_t cs = g_stream;
    if (!direct) {
      cs = g_comm.stream;
      CUDA_CHECK(cudaEventRecord(g_comm.ev_pack, g_stream));
      CUDA_CHECK(cudaStreamWaitEvent(cs, g_comm.ev_pack, 0));
      for (int mu = 1; mu < 4; mu++) {
        if (!((op->g.comm_mask >> mu) & 1)) continue;
        size_t bytes = (size_t)(op->g.half4 / op->g.L[mu]) * op->ls() * 12 * sizeof(T);
        CUDA_CHECK(cudaMemcpyAsync(h->peer[mu][0] + h->off_recv[mu][1][h->seq & 1], op->halo_send[mu][0], bytes, cudaMemcpyDefault, cs));
        CUDA_CHECK(cudaMemcpyAsync(h->peer[mu][1] + h->off_recv[mu][0][h->seq & 1], op->halo_send[mu][1], bytes, cudaMemcpyDefault, cs));
      }
    }
    for (int mu = 1; mu < 4; mu++) {
      if (!((op->g.comm_mask >> mu) & 1)) continue;
      comm_stream_write32(cs, h->peer[mu][0] + HaloP2P::flag_off(mu, 1), h->seq);
      comm_stream_write32(cs, h->peer[mu][1] + HaloP2P::flag_off(mu, 0), h->seq);
    }
    if (!direct) CUDA_CHECK(cudaEventRecord(g_comm.ev_comm, cs));
    return;
  }
  for (int mu = 1; mu < 4; mu++) {
    if (!((op->g.comm_mask >> mu) & 1)) continue;
    if (mu == 1) pack_t<T, 1>(op, dag, 1 - p_out, in, in_stride);
    if (mu == 2) pack_t<T, 2>(op, dag, 1 - p_out, in, in_stride);
    if (mu == 3) pack_t<T, 3>(op, dag, 1 - p_out, in, in_stride);
  }
  CUDA_CHECK(cudaEventRecord(g_comm.ev_pack, g_stream));
  CUDA_CHECK(cudaStreamWaitEvent(g_comm.stream, g_comm.ev_pack, 0));
  comm_exchange_begin();
  for (int mu = 1; mu < 4; mu++) {
    if (!((op->g.comm_mask >> mu) & 1)) continue;
    size_t bytes = (size_t)(op->g.half4 / op->g.L[mu]) * op->ls() * 12 * sizeof(T);
    comm_exchange_dir(mu, op->halo_send[mu][0], op->halo_send[mu][1], op->halo_recv[mu][0], op->halo_recv[mu][1], bytes, g_comm.stream);
  }
  comm_exchange_end();
  CUDA_CHECK(cudaEventRecord(g_comm.ev_comm, g_comm.stream));
}

template <typename T>
static void halo_end_t(cgptb_fermion_operator* op, bool dag, int p_out, T* out, size_t out_stride) {
  if (op->p2p) {
    HaloP2P* h = (HaloP2P*)op->p2p;
    for (int mu = 1; mu < 4; mu++) {
      if (!((op->g.comm_mask >> mu) & 1)) continue;
      comm_stream_wait_geq32(g_stream, h->arena + HaloP2P::flag_off(mu, 0), h->seq);
      comm_stream_wait_geq32(g_stream, h->arena + HaloP2P::flag_off(mu, 1), h->seq);
    }
  } else {
    CUDA_CHECK(cudaStreamWaitEvent(g_stream, g_comm.ev_comm, 0));
  }
  for (int mu = 1; mu < 4; mu++) {
    if (!((op->g.comm_mask >> mu) & 1)) continue;
    if (mu == 1) exterior_t<T, 1>(op, dag, p_out, out, out_stride);
    if (mu == 2) exterior_t<T, 2>(op, dag, p_out, out, out_stride);
    if (mu == 3) exterior_t<T, 3>(op, dag, p_out, out, out_stride);
  }
}

void halo_begin(cgptb_fermion_operator* op, bool dag, int p_out, const void* in, size_t in_stride) {
  if (op->prec == CGPTB_SINGLE)
    halo_begin_t<float>(op, dag, p_out, (const float*)in, in_stride);
  else
    halo_begin_t<double>(op, dag, p_out, (const double*)in, in_stride);
}

void halo_end(cgptb_fermion_operator* op, bool dag, int p_out, void* out, size_t out_stride) {
  if (op->prec == CGPTB_SINGLE)
    halo_end_t<float>(op, dag, p_out, (float*)out, out_stride);
  else
    halo_end_t<double>(op, dag, p_out, (double*)out, out_stride);
}

// decomposition set-up of an operator: comm mask, global offsets, halo buffers, ghost links
void halo_setup(cgptb_fermion_operator* op, const cgptb_lattice* const U[4]) {
  for (int mu = 0; mu < 4; mu++) {
    op->goff[mu] = g_comm.pcoor[mu] * op->dims4[mu];
    op->gL[mu] = g_comm.pgrid[mu] * op->dims4[mu];
  }
  op->g.comm_mask = 0;
  if (!g_comm.active) return;
  size_t real = op->prec == CGPTB_SINGLE ? 4 : 8;
  const bool p2p = comm_p2p_available();
  if (p2p && !op->p2p) {
    // arena: flag words, then the receive buffers; taken from the pool of released arenas if one is large enough
    HaloP2P* h = new HaloP2P;
    size_t off = 4096;
    for (int mu = 1; mu < 4; mu++) {
      if (g_comm.pgrid[mu] == 1) continue;
      size_t bytes = (((size_t)(op->g.half4 / op->g.L[mu]) * op->ls() * 12 * real) + 255) & ~(size_t)255;
      for (int side = 0; side < 2; side++)
        for (int par = 0; par < 2; par++) {
          h->off_recv[mu][side][par] = off;
          off += bytes;
        }
    }
    h->arena_bytes = off;
    for (size_t i = 0; i < g_arenas.free_list.size(); i++)
      if (g_arenas.free_list[i].second >= off) {
        h->arena = g_arenas.free_list[i].first;
        h->arena_bytes = g_arenas.free_list[i].second;
        g_arenas.free_list.erase(g_arenas.free_list.begin() + i);
        break;
      }
    if (!h->arena) CUDA_CHECK(cudaMalloc(&h->arena, h->arena_bytes));
    // flags start at zero BEFORE any neighbour learns about this arena (the all-gather below)
    CUDA_CHECK(cudaMemsetAsync(h->arena, 0, 4096, g_stream));
    CUDA_CHECK(cudaStreamSynchronize(g_stream));
    CommExport mine;
    comm_export(h->arena, &mine);
    std::vector<CommExport> all(g_comm.world);
    comm_allgather_host(&mine, all.data(), sizeof(CommExport));
    for (int mu = 1; mu < 4; mu++) {
      if (g_comm.pgrid[mu] == 1) continue;
      for (int side = 0; side < 2; side++) {
        int nb = comm_neighbor_rank(mu, side ? +1 : -1);
        h->peer[mu][side] = (char*)comm_import(nb, &all[nb]);
        if (!h->peer[mu][side]) CGPTB_ERR("cannot map the halo arena of rank %d (set CGPTB_HALO=nccl)", nb);
      }
    }
    op->p2p = h;
  }
  for (int mu = 1; mu < 4; mu++) {
    if (g_comm.pgrid[mu] == 1) continue;
    op->g.comm_mask |= 1 << mu;
    size_t bytes = (size_t)(op->g.half4 / op->g.L[mu]) * op->ls() * 12 * real;
    for (int side = 0; side < 2; side++) {
      if (!op->halo_send[mu][side]) CUDA_CHECK(cudaMalloc(&op->halo_send[mu][side], bytes));
      if (!p2p && !op->halo_recv[mu][side]) CUDA_CHECK(cudaMalloc(&op->halo_recv[mu][side], bytes));
    }
    // ghost links: my rank-mu neighbour's high-face U_mu
    FaceGeom fg = make_face(op->g, mu);
    int nft = op->g.L[0] * fg.A * fg.B;
    size_t lbytes = (size_t)nft * 18 * sizeof(double);
    double *snd, *dummy;
    CUDA_CHECK(cudaMalloc(&snd, lbytes));
    CUDA_CHECK(cudaMalloc(&dummy, lbytes));
    if (!op->ghost_links[mu]) CUDA_CHECK(cudaMalloc(&op->ghost_links[mu], lbytes));
    unsigned blocks = (unsigned)((nft + 127) / 128);
    if (U[mu]->prec == CGPTB_SINGLE)
      k_pack_links<float><<<blocks, 128, 0, g_stream>>>(op->g, fg, U[mu]->sites, (const float*)U[mu]->data, snd);
    else
      k_pack_links<double><<<blocks, 128, 0, g_stream>>>(op->g, fg, U[mu]->sites, (const double*)U[mu]->data, snd);
    LAUNCH_CHECK();
    comm_exchange_begin();
    // to_hi = my high face; what arrives in from_lo is the high face of rank-mu
    comm_exchange_dir(mu, snd, snd, op->ghost_links[mu], dummy, lbytes, g_stream);
    comm_exchange_end();
    CUDA_CHECK(cudaStreamSynchronize(g_stream));
    CUDA_CHECK(cudaFree(snd));
    CUDA_CHECK(cudaFree(dummy));
  }
}

// the arena goes back to the pool: the neighbours have it mapped, and none of them writes into it any more (every
// store into it belonged to an exchange this rank has waited for)
void halo_release(cgptb_fermion_operator* op) {
  if (!op->p2p) return;
  HaloP2P* h = (HaloP2P*)op->p2p;
  cudaStreamSynchronize(g_stream);
  if (g_comm.stream) cudaStreamSynchronize(g_comm.stream);
  g_arenas.free_list.push_back(std::make_pair(h->arena, h->arena_bytes));
  delete h;
  op->p2p = 0;
}

}  // namespace cgptb

// Fermion operator applied to HOST buffers: cgptb_apply_fermion_operator_host.
//
// What the reference does for data that lives in host memory is  lattice[:] = array ; dst = op * src ; array = dst[:]
// (lib/gpt/core/lattice.py:213-260 -> cgpt.lattice_import / lattice_export around cgpt.apply_fermion_operator): one
// host->device copy, the operator, one device->host copy, strictly one after the other.  On a PCIe-attached B200 the two
// copies are > 95 % of that time, so this entry point pipelines them for the hopping term:
//
//   * the host array (GPT order: s fastest, then x, y, z, t) is cut into slabs of consecutive time slices;
//   * copy engine 1 uploads slab after slab (pinned or pageable host memory), the compute stream reorders each slab into
//     the device layout (k_slab_layout) and, as soon as the slabs j-1, j, j+1 are resident, runs the hopping term on the
//     time slices of slab j only (the TMA sweep kernel takes a time range), reorders the result back into GPT order
//     and copy engine 2 downloads it while later slabs are still being uploaded;
//   * the slabs 0 and N-1 need each other (periodic lattice) and are done last; on a lattice split in t across GPUs they
//     are the only ones that need the neighbours' faces, so the halo exchange happens once, at the end.
//
// Everything else (other opcodes, double precision, split lattices, lattices the sweep kernel does not tile) goes through
// the plain import -> apply -> export sequence, so the call is valid for every opcode of register.h:2-20.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "operator.cuh"

using namespace cgptb;

namespace cgptb {

// one thread per (parity, device site of the slab): moves the three 32-byte blocks of a spinor
//   raw : slab in GPT order  [t - t0][z][y][x][s][12 complex]   (device staging copy of the host array)
//   dev : field in device layout, planes of `stride` 32-byte blocks: [k][parity][i4 * ls + s]
template <bool IMPORT>
__global__ void __launch_bounds__(256) k_slab_layout(Geom g, int ls, int t0, int nt, float* __restrict__ raw, float* __restrict__ dev,
                                                     size_t stride) {
  const size_t slice = (size_t)g.hx * g.L[1] * g.L[2] * ls;  // device sites of one parity per time slice
  const size_t n = 2 * slice * nt;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int p = idx >= slice * nt ? 1 : 0;
  const size_t r = idx - (size_t)p * slice * nt;  // = (i4 - t0 * slice4) * ls + s within the slab
  const int s = (int)(r % ls);
  const size_t q = r / ls;
  const int xh = (int)(q % g.hx);
  size_t w = q / g.hx;
  const int y = (int)(w % g.L[1]);
  w /= g.L[1];
  const int z = (int)(w % g.L[2]);
  const int tl = (int)(w / g.L[2]);
  const int t = t0 + tl;
  const int x = 2 * xh + ((y + z + t + p) & 1);
  const size_t lex = x + (size_t)g.L[0] * (y + (size_t)g.L[1] * (z + (size_t)g.L[2] * tl));
  float* h = raw + (lex * ls + s) * 24;
  const size_t dsite = (size_t)p * (stride / 2) + (q + (size_t)t0 * (slice / ls)) * ls + s;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float v[8];
    if (IMPORT) {
      ld256(h + 8 * k, v);
      st256(dev + ((size_t)k * stride + dsite) * 8, v);
    } else {
      ld256_rw(dev + ((size_t)k * stride + dsite) * 8, v);
      st256_cs(h + 8 * k, v);
    }
  }
}

struct HostPipe {
  cudaStream_t h2d = 0, d2h = 0;
  std::vector<cudaEvent_t> ev_in, ev_out;
  float* raw_in = 0;
  float* raw_out = 0;
  size_t raw_bytes = 0;
  cgptb_lattice* f_in = 0;
  cgptb_lattice* f_out = 0;
};
static HostPipe g_pipe;

static cgptb_lattice* pipe_field(cgptb_lattice*& slot, const cgptb_fermion_operator* op) {
  if (slot && (slot->prec != op->prec || slot->Ls != op->Ls || memcmp(slot->dims4, op->dims4, sizeof(op->dims4)) != 0)) {
    cgptb_delete_lattice(slot);
    slot = 0;
  }
  if (!slot) {
    if (cgptb_create_lattice(&slot, op->dims4, op->Ls, op->prec, 12, CGPTB_FULL)) throw Error{g_error};
  }
  return slot;
}

static bool pipeline_usable(const cgptb_fermion_operator* op, int opcode, int nslab) {
  if (getenv("CGPTB_NO_HOSTPIPE")) return false;
  if (opcode != 3001 && opcode != 4001) return false;  // Dhop, DhopDag (register.h:13-14)
  // open boundary conditions in time: op_dhop clears the two boundary slices after the stencil (P D P); the slab pipeline
  // exports slabs as they finish, so these operators take the import -> op_apply -> export path
  if (op->open_bc) return false;
  // a lattice split in t only keeps every hop of the interior slabs on this rank; the first and the last slab go through
  // the regular halo exchange at the end
  if (!dhop_tma_usable(op) || (op->g.comm_mask & ~8)) return false;
  return op->g.L[3] % nslab == 0 && op->g.L[3] / nslab >= 1 && nslab >= 3;
}

static void apply_host_pipelined(cgptb_fermion_operator* op, bool dag, const float* host_src, float* host_dst, int nslab) {
  HostPipe& P = g_pipe;
  const Geom& g = op->g;
  const int ls = op->ls();
  const int T = g.L[3], nt = T / nslab;
  const size_t slab_reals = (size_t)g.L[0] * g.L[1] * g.L[2] * nt * ls * 24;
  const size_t slab_bytes = slab_reals * sizeof(float);
  const size_t total = slab_bytes * nslab;
  if (!P.h2d) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&P.h2d, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&P.d2h, cudaStreamNonBlocking));
  }
  while ((int)P.ev_in.size() < nslab) {
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    P.ev_in.push_back(a);
    P.ev_out.push_back(b);
  }
  if (P.raw_bytes < total) {
    if (P.raw_in) CUDA_CHECK(cudaFree(P.raw_in));
    if (P.raw_out) CUDA_CHECK(cudaFree(P.raw_out));
    CUDA_CHECK(cudaMalloc(&P.raw_in, total));
    CUDA_CHECK(cudaMalloc(&P.raw_out, total));
    P.raw_bytes = total;
  }
  cgptb_lattice* fin = pipe_field(P.f_in, op);
  cgptb_lattice* fout = pipe_field(P.f_out, op);
  const size_t stride = fin->sites;  // 32-byte blocks per component plane
  const size_t half = stride / 2;
  const size_t nthreads = 2 * (size_t)g.hx * g.L[1] * g.L[2] * ls * nt;
  const unsigned blocks = (unsigned)((nthreads + 255) / 256);

  // the staging buffers and fields of the previous call may still be in flight on the copy streams: they are ours again
  // once everything queued so far on the compute stream and the download stream has been ordered before the uploads
  // (calls are synchronous: the previous call ended with a synchronize)
  for (int j = 0; j < nslab; j++) {
    CUDA_CHECK(cudaMemcpyAsync(P.raw_in + (size_t)j * slab_reals, host_src + (size_t)j * slab_reals, slab_bytes, cudaMemcpyHostToDevice, P.h2d));
    CUDA_CHECK(cudaEventRecord(P.ev_in[j], P.h2d));
  }
  auto export_slab = [&](int j) {
    k_slab_layout<false><<<blocks, 256, 0, g_stream>>>(g, ls, j * nt, nt, P.raw_out + (size_t)j * slab_reals, (float*)fout->data, stride);
    LAUNCH_CHECK();
    CUDA_CHECK(cudaEventRecord(P.ev_out[j], g_stream));
    CUDA_CHECK(cudaStreamWaitEvent(P.d2h, P.ev_out[j], 0));
    CUDA_CHECK(cudaMemcpyAsync(host_dst + (size_t)j * slab_reals, P.raw_out + (size_t)j * slab_reals, slab_bytes, cudaMemcpyDeviceToHost, P.d2h));
  };
  auto compute_slab = [&](int j) {
    for (int p = 0; p < 2; p++) {
      const float* pin = (const float*)fin->data + (size_t)(1 - p) * half * 8;
      float* pout = (float*)fout->data + (size_t)p * half * 8;
      dhop_half_f32_tma(op, dag, pin, stride, pout, stride, p, j * nt, nt);
    }
    export_slab(j);
  };
  // the two slabs at the ends of the time direction: periodic neighbours of each other on one GPU; on a lattice split in t
  // their outermost slices also take the faces of the neighbouring ranks (pack -> NCCL -> interior -> exterior, halo.cu)
  auto compute_boundary_slabs = [&]() {
    for (int p = 0; p < 2; p++) {
      const float* pin = (const float*)fin->data + (size_t)(1 - p) * half * 8;
      float* pout = (float*)fout->data + (size_t)p * half * 8;
      if (g.comm_mask) halo_begin(op, dag, p, pin, stride);
      dhop_half_f32_tma(op, dag, pin, stride, pout, stride, p, (nslab - 1) * nt, nt);
      dhop_half_f32_tma(op, dag, pin, stride, pout, stride, p, 0, nt);
      if (g.comm_mask) halo_end(op, dag, p, pout, stride);
    }
    export_slab(nslab - 1);
    export_slab(0);
  };
  for (int j = 0; j < nslab; j++) {
    CUDA_CHECK(cudaStreamWaitEvent(g_stream, P.ev_in[j], 0));
    k_slab_layout<true><<<blocks, 256, 0, g_stream>>>(g, ls, j * nt, nt, P.raw_in + (size_t)j * slab_reals, (float*)fin->data, stride);
    LAUNCH_CHECK();
    if (j >= 2) compute_slab(j - 1);
  }
  compute_boundary_slabs();
  CUDA_CHECK(cudaStreamSynchronize(P.d2h));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
}

}  // namespace cgptb

extern "C" {

// dst_host = op(opcode) src_host; both buffers hold a full 5d (Moebius) or 4d (Wilson) spin-colour field in GPT order
// in the operator's precision (nbytes each).  Synchronous like every cgpt call.
int cgptb_apply_fermion_operator_host(cgptb_fermion_operator* op, int opcode, const void* src_host, void* dst_host, size_t nbytes) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(op && src_host && dst_host && src_host != dst_host);
  cgptb_lattice* fin = pipe_field(g_pipe.f_in, op);
  cgptb_lattice* fout = pipe_field(g_pipe.f_out, op);
  if (nbytes != fin->bytes()) CGPTB_ERR("apply_fermion_operator_host: buffers have %zu bytes, the field needs %zu", nbytes, fin->bytes());
  const char* e = getenv("CGPTB_HOSTPIPE_SLABS");
  int nslab = e ? atoi(e) : 32;  // measured at 32^3x64x12: 8 slabs 63.3 ms, 16: 56.3, 32: 53.6 per call
  while (nslab > 3 && op->g.L[3] % nslab) nslab--;
  if (pipeline_usable(op, opcode, nslab)) {
    apply_host_pipelined(op, opcode == 4001, (const float*)src_host, (float*)dst_host, nslab);
  } else {
    if (cgptb_lattice_import(fin, src_host, nbytes)) throw Error{g_error};
    fin->cb = CGPTB_FULL;
    fout->cb = CGPTB_FULL;
    op_apply(op, opcode, fin, fout);
    if (cgptb_lattice_export(fout, dst_host, nbytes)) throw Error{g_error};
  }
  CGPTB_API_END
}
}

// Multi-GPU plumbing of libcgpt_b200: one process per GPU, a Cartesian processor grid over the 4d lattice
// (GPT's --mpi X.Y.Z.T, lib/gpt/core/grid.py:77-94; the s-direction is never split, grid.py:83-87).
// Replaces what the reference gets from Grid's CartesianCommunicator: halo send/recv (C1 in SURVEY.md 2.3) and
// the global sum of reduction results (C2, lib/cgpt/lib/grid.cc:119-153).
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") so that a single-GPU process has no NCCL
// dependency; inside a torch process this picks up the libnccl torch already mapped.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>
#include <map>
#include <string>
#include <vector>
#include "common.cuh"

namespace cgptb {

Comm g_comm;

struct NcclApi {
  void* handle = 0;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = 0;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = 0;
  ncclResult_t (*CommDestroy)(ncclComm_t) = 0;
  const char* (*GetErrorString)(ncclResult_t) = 0;
  ncclResult_t (*GroupStart)() = 0;
  ncclResult_t (*GroupEnd)() = 0;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = 0;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = 0;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = 0;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = 0;
};
static NcclApi nccl;

#define NCCL_CHECK(x)                                                                               \
  do {                                                                                              \
    ncclResult_t _r = (x);                                                                          \
    if (_r != ncclSuccess) CGPTB_ERR("NCCL error %s at %s:%d", nccl.GetErrorString(_r), __FILE__, __LINE__); \
  } while (0)

static void load_nccl() {
  if (nccl.handle) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (nccl.handle) break;
  }
  if (!nccl.handle) CGPTB_ERR("cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                   \
  *(void**)(&nccl.field) = dlsym(nccl.handle, name);                       \
  if (!nccl.field) CGPTB_ERR("libnccl has no symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(GetErrorString, "ncclGetErrorString")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
#undef SYM
}

int comm_neighbor_rank(int mu, int dir) {
  int c[4];
  for (int i = 0; i < 4; i++) c[i] = g_comm.pcoor[i];
  c[mu] = (c[mu] + dir + g_comm.pgrid[mu]) % g_comm.pgrid[mu];
  return c[0] + g_comm.pgrid[0] * (c[1] + g_comm.pgrid[1] * (c[2] + g_comm.pgrid[2] * c[3]));
}

// exchange with the two neighbours in direction mu: send `to_lo` to rank-mu and `to_hi` to rank+mu,
// receive `from_lo` (sent by rank-mu as its to_hi) and `from_hi` (sent by rank+mu as its to_lo)
void comm_exchange_begin() { NCCL_CHECK(nccl.GroupStart()); }
void comm_exchange_dir(int mu, const void* to_lo, const void* to_hi, void* from_lo, void* from_hi, size_t bytes, cudaStream_t s) {
  int lo = comm_neighbor_rank(mu, -1), hi = comm_neighbor_rank(mu, +1);
  ncclComm_t c = (ncclComm_t)g_comm.nccl;
  NCCL_CHECK(nccl.Send(to_lo, bytes, ncclUint8, lo, c, s));
  NCCL_CHECK(nccl.Send(to_hi, bytes, ncclUint8, hi, c, s));
  // receive order matters when lo == hi (two ranks in this direction): the peer's first message is its
  // to_lo, which is my from_hi
  NCCL_CHECK(nccl.Recv(from_hi, bytes, ncclUint8, hi, c, s));
  NCCL_CHECK(nccl.Recv(from_lo, bytes, ncclUint8, lo, c, s));
}
void comm_exchange_end() { NCCL_CHECK(nccl.GroupEnd()); }

// in-place sum over all ranks of n doubles in device memory, on stream s
void comm_allreduce_device(double* dev, int n, cudaStream_t s) {
  if (!g_comm.active) return;
  NCCL_CHECK(nccl.AllReduce(dev, dev, n, ncclDouble, ncclSum, (ncclComm_t)g_comm.nccl, s));
}

// out = [world][bytes] in device memory, on stream s
void comm_allgather_device(const void* in, void* out, size_t bytes, cudaStream_t s) {
  NCCL_CHECK(nccl.AllGather(in, out, bytes, ncclUint8, (ncclComm_t)g_comm.nccl, s));
}

// every rank contributes `bytes` bytes of host memory; out = [world][bytes] (set-up paths only: synchronises the stream)
void comm_allgather_host(const void* in, void* out, size_t bytes) {
  unsigned char *d_in, *d_out;
  CUDA_CHECK(cudaMalloc(&d_in, bytes));
  CUDA_CHECK(cudaMalloc(&d_out, bytes * g_comm.world));
  CUDA_CHECK(cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, g_stream));
  NCCL_CHECK(nccl.AllGather(d_in, d_out, bytes, ncclUint8, (ncclComm_t)g_comm.nccl, g_stream));
  CUDA_CHECK(cudaMemcpyAsync(out, d_out, bytes * g_comm.world, cudaMemcpyDeviceToHost, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  CUDA_CHECK(cudaFree(d_in));
  CUDA_CHECK(cudaFree(d_out));
}

// ---- peer memory (CUDA IPC) and stream-ordered flags: the halo exchange without NCCL kernels -------------------------------
// A rank exports a device allocation (comm_export), all-gathers the descriptor, and maps the neighbours' allocations
// (comm_import; mappings are cached for the life of the process and never closed: the exporting side keeps its halo arenas
// in a pool instead of freeing them).  cuStreamWriteValue32 / cuStreamWaitValue32 on words of such an allocation order a
// receiver's kernels after a sender's.
static CUresult (*p_cuStreamWriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = 0;
static CUresult (*p_cuStreamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = 0;
static CUresult (*p_cuMemGetAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr) = 0;
static int p2p_state = -1;  // -1 unknown, 0 unavailable, 1 available

static bool load_memops() {
  if (p_cuStreamWriteValue32) return true;
  cudaDriverEntryPointQueryResult q;
  void* f = 0;
  if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) return false;
  *(void**)(&p_cuStreamWriteValue32) = f;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) return false;
  *(void**)(&p_cuStreamWaitValue32) = f;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) return false;
  *(void**)(&p_cuMemGetAddressRange) = f;
  return true;
}

void comm_export(void* ptr, CommExport* e) {
  memset(e, 0, sizeof(*e));
  CUdeviceptr base = 0;
  size_t size = 0;
  if (p_cuMemGetAddressRange(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) CGPTB_ERR("cuMemGetAddressRange failed");
  cudaIpcMemHandle_t h;
  CUDA_CHECK(cudaIpcGetMemHandle(&h, (void*)base));
  static_assert(sizeof(h) == sizeof(e->handle), "ipc handle size");
  memcpy(e->handle, &h, sizeof(h));
  e->offset = (unsigned long long)((CUdeviceptr)ptr - base);
}

void* comm_import(int rank, const CommExport* e) {
  static std::map<std::string, void*> opened;
  std::string key((const char*)e->handle, sizeof(e->handle));
  key += (char)rank;
  auto f = opened.find(key);
  if (f == opened.end()) {
    cudaIpcMemHandle_t h;
    memcpy(&h, e->handle, sizeof(h));
    void* base = 0;
    cudaError_t err = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    f = opened.emplace(key, base).first;
  }
  return (char*)f->second + e->offset;
}

// peer-to-peer halo path usable?  All ranks of a box, memory operations available, not switched off (CGPTB_HALO=nccl);
// decided once, collectively (every rank tries to map an allocation of its successor).
bool comm_p2p_available() {
  if (p2p_state >= 0) return p2p_state == 1;
  p2p_state = 0;
  const char* v = getenv("CGPTB_HALO");
  int ok = g_comm.active && !(v && !strcmp(v, "nccl")) && load_memops() ? 1 : 0;
  void* probe = 0;
  std::vector<CommExport> all(g_comm.world);
  if (ok) {
    CUDA_CHECK(cudaMalloc(&probe, 1 << 21));
    CommExport mine;
    comm_export(probe, &mine);
    comm_allgather_host(&mine, all.data(), sizeof(CommExport));
    int nb = (g_comm.rank + 1) % g_comm.world;
    if (!comm_import(nb, &all[nb])) ok = 0;
  }
  // agree: one rank that cannot map its neighbour switches everybody to NCCL
  double* d = reduce_scratch(8);
  double h = ok ? 0.0 : 1.0;
  CUDA_CHECK(cudaMemcpyAsync(d, &h, sizeof(double), cudaMemcpyHostToDevice, g_stream));
  comm_allreduce_device(d, 1, g_stream);
  CUDA_CHECK(cudaMemcpyAsync(&h, d, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  p2p_state = h == 0.0 ? 1 : 0;
  return p2p_state == 1;  // the probe allocation stays (a neighbour has it mapped)
}

void comm_stream_write32(cudaStream_t s, void* addr, unsigned value) {
  if (p_cuStreamWriteValue32((CUstream)s, (CUdeviceptr)addr, value, 0) != CUDA_SUCCESS) CGPTB_ERR("cuStreamWriteValue32 failed");
}
void comm_stream_wait_geq32(cudaStream_t s, void* addr, unsigned value) {
  if (p_cuStreamWaitValue32((CUstream)s, (CUdeviceptr)addr, value, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) CGPTB_ERR("cuStreamWaitValue32 failed");
}

}  // namespace cgptb

using namespace cgptb;

extern "C" {

int cgptb_comm_unique_id(char* id128) {
  CGPTB_API_BEGIN
  load_nccl();
  ncclUniqueId id;
  NCCL_CHECK(nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
  CGPTB_API_END
}

int cgptb_comm_init(int rank, int world, const int mpi[4], const char* id128) {
  CGPTB_API_BEGIN
  if (g_stream == 0) CGPTB_ERR("cgptb_init() must be called before cgptb_comm_init()");
  if (mpi[0] * mpi[1] * mpi[2] * mpi[3] != world) CGPTB_ERR("processor grid %d.%d.%d.%d does not match %d ranks", mpi[0], mpi[1], mpi[2], mpi[3], world);
  if (mpi[0] != 1) CGPTB_ERR("the x direction (checkerboarded) cannot be split; use --mpi 1.Y.Z.T");
  if (g_comm.active) CGPTB_ERR("communicator already initialised");
  load_nccl();
  ncclUniqueId id;
  memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t c;
  NCCL_CHECK(nccl.CommInitRank(&c, world, id, rank));
  g_comm.nccl = (void*)c;
  g_comm.rank = rank;
  g_comm.world = world;
  int r = rank;
  for (int i = 0; i < 4; i++) {
    g_comm.pgrid[i] = mpi[i];
    g_comm.pcoor[i] = r % mpi[i];
    r /= mpi[i];
  }
  CUDA_CHECK(cudaStreamCreateWithFlags(&g_comm.stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&g_comm.ev_pack, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&g_comm.ev_comm, cudaEventDisableTiming));
  g_comm.active = world > 1;
  CGPTB_API_END
}

int cgptb_comm_finalize(void) {
  CGPTB_API_BEGIN
  if (g_comm.nccl) {
    CUDA_CHECK(cudaStreamSynchronize(g_comm.stream));
    nccl.CommDestroy((ncclComm_t)g_comm.nccl);
    g_comm.nccl = 0;
    g_comm.active = false;
  }
  CGPTB_API_END
}

int cgptb_comm_info(int* rank, int* world, int pgrid[4], int pcoor[4]) {
  *rank = g_comm.rank;
  *world = g_comm.world;
  for (int i = 0; i < 4; i++) {
    pgrid[i] = g_comm.pgrid[i];
    pcoor[i] = g_comm.pcoor[i];
  }
  return 0;
}

// cgpt.grid_globalsum (lib/cgpt/lib/grid.cc:119-160) for an array of doubles in HOST memory
int cgptb_comm_globalsum(double* host, int n) {
  CGPTB_API_BEGIN
  if (g_comm.active) {
    double* d = reduce_scratch((size_t)sm_count() * 8 * 3 + 8 + n) + (size_t)sm_count() * 8 * 3 + 8;
    CUDA_CHECK(cudaMemcpyAsync(d, host, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    comm_allreduce_device(d, n, g_stream);
    CUDA_CHECK(cudaMemcpyAsync(host, d, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CUDA_CHECK(cudaStreamSynchronize(g_stream));
  }
  CGPTB_API_END
}
}

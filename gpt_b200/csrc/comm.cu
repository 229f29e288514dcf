// Multi-GPU plumbing of libcgpt_b200: one process per GPU, a Cartesian processor grid over the 4d lattice
// (GPT's --mpi X.Y.Z.T, lib/gpt/core/grid.py:77-94; the s-direction is never split, grid.py:83-87).
// Replaces what the reference gets from Grid's CartesianCommunicator: halo send/recv (C1 in SURVEY.md 2.3) and
// the global sum of reduction results (C2, lib/cgpt/lib/grid.cc:119-153).
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") so that a single-GPU process has no NCCL
// dependency; inside a torch process this picks up the libnccl torch already mapped.
#include <dlfcn.h>
#include <nccl.h>
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

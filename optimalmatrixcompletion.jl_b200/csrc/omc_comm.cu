// omc_comm.cu -- the engine's only inter-GPU exchange (SURVEY.md sections 2.2 K10, 5, 8b/8e): one process per GPU, the
// frontier of open nodes sharded over the processes, and after every batch
//   * an all-reduce-min of [incumbent upper bound, smallest open lower bound, ...]   (omc_allreduce_min), and
//   * an all-gather of per-rank load figures (open nodes, ADMM iterations of the last batch) from which every rank derives the
//     same re-balanced shard of the global open list (omc_allgather; the node descriptors are pool ids + direction codes that
//     every rank already holds, so no node data has to move).
// The collectives are NCCL's (ncclAllReduce / ncclAllGather over NVLink); payloads are a few doubles, so latency is what
// counts.  NCCL is bound at run time with dlopen (libnccl.so.2): a host that already loaded NCCL (torch.distributed) shares its
// copy, and a single-GPU host never needs the library.  The reference has no counterpart (single process, OMC.jl:700-1073).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/omc_b200.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };                      // ncclDataType_t (nccl.h: ncclDouble = 8)
enum { ncclSumOp = 0, ncclProdOp = 1, ncclMaxOp = 2, ncclMinOp = 3 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;
double* g_buf = nullptr;     // device staging: [0, cap) send, [cap, cap + cap * world) receive
size_t g_cap = 0;
cudaStream_t g_cstream = nullptr;
std::string g_cerr;

int cfail(int code, const std::string& msg) { g_cerr = msg; return code; }

bool load_nccl() {
  if (g_nccl.handle) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) return false;
#define OMC_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name)
  OMC_SYM(GetUniqueId, "ncclGetUniqueId");
  OMC_SYM(CommInitRank, "ncclCommInitRank");
  OMC_SYM(CommDestroy, "ncclCommDestroy");
  OMC_SYM(AllReduce, "ncclAllReduce");
  OMC_SYM(AllGather, "ncclAllGather");
  OMC_SYM(GetErrorString, "ncclGetErrorString");
#undef OMC_SYM
  return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.AllGather;
}

}  // namespace

extern "C" {

const char* omc_comm_last_error(void) { return g_cerr.c_str(); }

int32_t omc_comm_unique_id(uint8_t* id128) {
  if (!id128) return cfail(OMC_ERR_ARG, "null id buffer");
  if (!load_nccl()) return cfail(OMC_ERR_UNSUPPORTED, std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""));
  ncclUniqueId id;
  const ncclResult_t rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return cfail(OMC_ERR_CUDA, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
  memcpy(id128, id.internal, 128);
  return OMC_OK;
}

int32_t omc_comm_init(int32_t rank, int32_t world, const uint8_t* id128) {
  if (world < 1 || rank < 0 || rank >= world) return cfail(OMC_ERR_ARG, "bad rank / world");
  g_rank = rank; g_world = world;
  if (world == 1) return OMC_OK;                       // nothing to exchange
  if (!id128) return cfail(OMC_ERR_ARG, "null id buffer");
  if (!load_nccl()) return cfail(OMC_ERR_UNSUPPORTED, "libnccl.so.2 not found");
  if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  const ncclResult_t rc = g_nccl.CommInitRank(&g_comm, world, id, rank);   // binds to the calling thread's current device (omc_init)
  if (rc != 0) return cfail(OMC_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
  if (!g_cstream && cudaStreamCreateWithFlags(&g_cstream, cudaStreamNonBlocking) != cudaSuccess) return cfail(OMC_ERR_CUDA, "stream");
  return OMC_OK;
}

static int ensure_buf(size_t n) {
  if (n <= g_cap) return OMC_OK;
  if (g_buf) cudaFree(g_buf);
  g_cap = n < 64 ? 64 : n;
  if (cudaMalloc(&g_buf, g_cap * (size_t)(1 + g_world) * sizeof(double)) != cudaSuccess) { g_buf = nullptr; g_cap = 0; return cfail(OMC_ERR_CUDA, "cudaMalloc"); }
  return OMC_OK;
}

/* values[i] <- min over ranks of values[i] (host buffer, n doubles); world = 1: no-op */
int32_t omc_allreduce_min(double* values, int32_t n) {
  if (!values || n <= 0) return cfail(OMC_ERR_ARG, "bad argument");
  if (g_world == 1) return OMC_OK;
  if (!g_comm) return cfail(OMC_ERR_STATE, "omc_comm_init() has not been called");
  if (ensure_buf(n) != OMC_OK) return OMC_ERR_CUDA;
  if (cudaMemcpyAsync(g_buf, values, n * sizeof(double), cudaMemcpyHostToDevice, g_cstream) != cudaSuccess) return cfail(OMC_ERR_CUDA, "H2D");
  const ncclResult_t rc = g_nccl.AllReduce(g_buf, g_buf, n, ncclFloat64, ncclMinOp, g_comm, g_cstream);
  if (rc != 0) return cfail(OMC_ERR_CUDA, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc));
  if (cudaMemcpyAsync(values, g_buf, n * sizeof(double), cudaMemcpyDeviceToHost, g_cstream) != cudaSuccess) return cfail(OMC_ERR_CUDA, "D2H");
  if (cudaStreamSynchronize(g_cstream) != cudaSuccess) return cfail(OMC_ERR_CUDA, "sync");
  return OMC_OK;
}

/* recv[r * n + i] <- send[i] of rank r (host buffers; recv holds world * n doubles); world = 1: copy */
int32_t omc_allgather(const double* send, int32_t n, double* recv) {
  if (!send || !recv || n <= 0) return cfail(OMC_ERR_ARG, "bad argument");
  if (g_world == 1) { memcpy(recv, send, n * sizeof(double)); return OMC_OK; }
  if (!g_comm) return cfail(OMC_ERR_STATE, "omc_comm_init() has not been called");
  if (ensure_buf(n) != OMC_OK) return OMC_ERR_CUDA;
  if (cudaMemcpyAsync(g_buf, send, n * sizeof(double), cudaMemcpyHostToDevice, g_cstream) != cudaSuccess) return cfail(OMC_ERR_CUDA, "H2D");
  const ncclResult_t rc = g_nccl.AllGather(g_buf, g_buf + g_cap, n, ncclFloat64, g_comm, g_cstream);
  if (rc != 0) return cfail(OMC_ERR_CUDA, std::string("ncclAllGather: ") + g_nccl.GetErrorString(rc));
  if (cudaMemcpyAsync(recv, g_buf + g_cap, (size_t)n * g_world * sizeof(double), cudaMemcpyDeviceToHost, g_cstream) != cudaSuccess) return cfail(OMC_ERR_CUDA, "D2H");
  if (cudaStreamSynchronize(g_cstream) != cudaSuccess) return cfail(OMC_ERR_CUDA, "sync");
  return OMC_OK;
}

int32_t omc_comm_info(int32_t* rank, int32_t* world) {
  if (rank) *rank = g_rank;
  if (world) *world = g_world;
  return OMC_OK;
}

int32_t omc_comm_destroy(void) {
  if (g_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_comm);
  g_comm = nullptr;
  if (g_buf) cudaFree(g_buf);
  g_buf = nullptr; g_cap = 0;
  if (g_cstream) cudaStreamDestroy(g_cstream);
  g_cstream = nullptr;
  g_rank = 0; g_world = 1;
  return OMC_OK;
}

}  // extern "C"

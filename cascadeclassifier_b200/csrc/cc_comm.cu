// Feature-sharded training: per tree node every rank contributes its local best split and all ranks
// agree on the global winner.  The exchange is an NCCL all-gather of one 64-byte
// candidate per rank on the training stream, followed by a device-side argmax with the reference's
// ordering rule (highest f32 quality, lowest variable index), i.e. an allreduce-argmax whose payload
// already carries the winner's threshold.  NCCL is loaded lazily (dlopen) so that the single-GPU
// library has no hard dependency on it.
#include <dlfcn.h>

#include <cstring>

#include "cc_internal.h"

namespace cc {

int launch_argmax_candidates(cc_ctx* ctx, const void* cands, int n, void* out);   // cc_boost.cu

namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_t_;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t*);
typedef int (*fn_CommInitRank)(ncclComm_t_*, int, ncclUniqueId_t, int);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, ncclComm_t_, cudaStream_t);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t);
typedef int (*fn_CommDestroy)(ncclComm_t_);
typedef const char* (*fn_GetErrorString)(int);

struct NcclApi {
  void* handle = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_AllGather AllGather = nullptr;
  fn_AllReduce AllReduce = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
} g_nccl;

int load_nccl() {
  if (g_nccl.handle) return CC_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  CC_REQUIRE(g_nccl.handle, CC_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_CommInitRank)dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.AllGather = (fn_AllGather)dlsym(g_nccl.handle, "ncclAllGather");
  g_nccl.AllReduce = (fn_AllReduce)dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.CommDestroy = (fn_CommDestroy)dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.GetErrorString = (fn_GetErrorString)dlsym(g_nccl.handle, "ncclGetErrorString");
  CC_REQUIRE(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllGather && g_nccl.AllReduce && g_nccl.CommDestroy, CC_ERR_NCCL,
             "libnccl is missing required symbols");
  return CC_OK;
}
}  // namespace

constexpr size_t kCandBytes = 64;   // sizeof(BestSplit), checked in cc_boost.cu

// d_best_buf: [0] = local candidate on entry / global winner on exit; [1..world] = gather area
int exchange_candidates(cc_ctx* ctx, void* d_best_buf) {
  if (ctx->comm_world <= 1 || !ctx->nccl_comm) return CC_OK;
  char* base = (char*)d_best_buf;
  int rc = g_nccl.AllGather(base, base + kCandBytes, kCandBytes, /*ncclChar*/ 0, (ncclComm_t_)ctx->nccl_comm, ctx->stream);
  CC_REQUIRE(rc == 0, CC_ERR_NCCL, "ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  return launch_argmax_candidates(ctx, base + kCandBytes, ctx->comm_world, base);
}

// Image-sharded mining: in-place sum over the ranks of a survivor bitmap (uint32 words) or of accepted-window pixels
// (uint8); the contributions are disjoint, so the sum is the union.
int allreduce_sum(cc_ctx* ctx, void* d_buf, size_t count, int elem_bytes) {
  if (ctx->comm_world <= 1 || !ctx->nccl_comm) return CC_OK;
  const int dtype = elem_bytes == 1 ? /*ncclUint8*/ 1 : /*ncclUint32*/ 3;
  int rc = g_nccl.AllReduce(d_buf, d_buf, count, dtype, /*ncclSum*/ 0, (ncclComm_t_)ctx->nccl_comm, ctx->stream);
  CC_REQUIRE(rc == 0, CC_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  return CC_OK;
}

}  // namespace cc

using namespace cc;

extern "C" {

int cc_comm_unique_id(void* out_id128) {
  CC_REQUIRE(out_id128, CC_ERR_BAD_ARG, "cc_comm_unique_id: null output");
  CC_CHECK(load_nccl());
  ncclUniqueId_t id;
  int rc = g_nccl.GetUniqueId(&id);
  CC_REQUIRE(rc == 0, CC_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
  std::memcpy(out_id128, &id, sizeof(id));
  return CC_OK;
}

int cc_comm_init(cc_ctx* ctx, const void* id128, int rank, int world) {
  CC_REQUIRE(ctx && id128, CC_ERR_BAD_ARG, "cc_comm_init: null argument");
  CC_REQUIRE(world >= 1 && world <= CC_MAX_RANKS && rank >= 0 && rank < world, CC_ERR_BAD_ARG,
             "cc_comm_init: rank %d of %d (the candidate buffer of the per-node exchange holds at most %d ranks)", rank, world, CC_MAX_RANKS);
  CC_CHECK(load_nccl());
  CC_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId_t id;
  std::memcpy(&id, id128, sizeof(id));
  ncclComm_t_ comm = nullptr;
  int rc = g_nccl.CommInitRank(&comm, world, id, rank);
  CC_REQUIRE(rc == 0, CC_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  ctx->nccl_comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  // NCCL sets its channels up lazily at the first collective of each kind (hundreds of ms on 8 ranks): do that here, not
  // inside the first stage training / the first mining batch
  {
    unsigned char* warm = nullptr;
    CC_CUDA(cudaMalloc(&warm, (size_t)64 * (world + 1) + 4096));
    CC_CUDA(cudaMemsetAsync(warm, 0, (size_t)64 * (world + 1) + 4096, ctx->stream));
    int r1 = g_nccl.AllGather(warm, warm + 64, 64, /*ncclChar*/ 0, comm, ctx->stream);
    int r2 = g_nccl.AllReduce(warm, warm, 1024, /*ncclUint32*/ 3, /*ncclSum*/ 0, comm, ctx->stream);
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(warm);
    CC_REQUIRE(r1 == 0 && r2 == 0, CC_ERR_NCCL, "NCCL warm-up collectives failed");
  }
  return cc_boost_set_shard(ctx, rank, world);
}

}  // extern "C"

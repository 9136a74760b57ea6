// Context lifetime, error reporting and event-based per-kernel-family timing.
#include <cstring>

#include "cc_internal.h"

namespace cc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Event pairs are recorded around every launch group and resolved lazily in cc_timing_get so that
// timing never inserts a host synchronisation into the training loop.
struct PendingEvent { cudaEvent_t a, b; std::string name; int launches; };
static std::map<cc_ctx*, std::vector<PendingEvent>> g_pending;
static std::map<cc_ctx*, std::vector<cudaEvent_t>> g_free_events;

static cudaEvent_t get_event(cc_ctx* ctx) {
  auto& fe = g_free_events[ctx];
  if (!fe.empty()) { cudaEvent_t e = fe.back(); fe.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

ScopedTimer::ScopedTimer(cc_ctx* c, const char* nm, int launches) : ctx(c), name(nm), n(launches) {
  ctx->launches += launches;
  if (!ctx->timing_enabled) return;
  PendingEvent pe{get_event(ctx), get_event(ctx), nm, launches};
  cudaEventRecord(pe.a, ctx->stream);
  g_pending[ctx].push_back(pe);
  idx = (int)g_pending[ctx].size() - 1;
}
ScopedTimer::~ScopedTimer() {
  if (!ctx->timing_enabled || idx < 0 || idx >= (int)g_pending[ctx].size()) return;
  cudaEventRecord(g_pending[ctx][idx].b, ctx->stream);
}

static void resolve_timers(cc_ctx* ctx) {
  auto& pend = g_pending[ctx];
  if (pend.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  for (auto& pe : pend) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) {
      ctx->timers[pe.name].ms += ms;
      ctx->timers[pe.name].launches += pe.launches;
    }
    g_free_events[ctx].push_back(pe.a);
    g_free_events[ctx].push_back(pe.b);
  }
  pend.clear();
}

}  // namespace cc

using namespace cc;

extern "C" {

const char* cc_last_error(void) { return cc::g_err; }
const char* cc_version(void) { return "cascade_cuda 0.1 (sm_100a)"; }

int cc_create(int device, cc_ctx** out) {
  CC_REQUIRE(out != nullptr, CC_ERR_BAD_ARG, "cc_create: out is null");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("cc_create: no CUDA device available (%s); libcascade_cuda has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return CC_ERR_CUDA;
  }
  CC_REQUIRE(device >= 0 && device < ndev, CC_ERR_BAD_ARG, "cc_create: device %d out of range (%d devices)", device, ndev);
  CC_CUDA(cudaSetDevice(device));
  cc_ctx* ctx = new cc_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  CC_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  CC_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->own_stream = true;
  {
    // temporaries come from the stream-ordered pool (cudaMallocAsync); keep freed blocks in the pool across
    // synchronisations instead of returning them to the driver every time (the default release threshold is 0)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  *out = ctx;
  return CC_OK;
}

void cc_destroy(cc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cc::resolve_timers(ctx);
  for (cudaEvent_t e : cc::g_free_events[ctx]) cudaEventDestroy(e);
  cc::g_free_events.erase(ctx);
  cc::g_pending.erase(ctx);
  cc::mine_free(ctx);
  cc::boost_free(ctx);
  cc::eval_free(ctx);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int cc_set_stream(cc_ctx* ctx, void* cuda_stream) {
  CC_REQUIRE(ctx, CC_ERR_BAD_ARG, "cc_set_stream: null ctx");
  CC_CUDA(cudaSetDevice(ctx->device));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  cc::resolve_timers(ctx);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (cuda_stream) {
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
  } else {
    CC_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return CC_OK;
}

int cc_synchronize(cc_ctx* ctx) {
  CC_REQUIRE(ctx, CC_ERR_BAD_ARG, "cc_synchronize: null ctx");
  CC_CUDA(cudaSetDevice(ctx->device));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  return CC_OK;
}

int cc_timing_reset(cc_ctx* ctx) {
  CC_REQUIRE(ctx, CC_ERR_BAD_ARG, "cc_timing_reset: null ctx");
  cc::resolve_timers(ctx);
  ctx->timers.clear();
  ctx->launches = 0;
  return CC_OK;
}

int cc_timing_get(cc_ctx* ctx, const char* name, double* ms, int64_t* launches) {
  CC_REQUIRE(ctx && name, CC_ERR_BAD_ARG, "cc_timing_get: null argument");
  cc::resolve_timers(ctx);
  auto it = ctx->timers.find(name);
  if (ms) *ms = it == ctx->timers.end() ? 0.0 : it->second.ms;
  if (launches) *launches = it == ctx->timers.end() ? 0 : it->second.launches;
  return CC_OK;
}

int64_t cc_launch_count(cc_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

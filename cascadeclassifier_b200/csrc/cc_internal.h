// Internal declarations shared by the translation units of libcascade_cuda.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/cascade_cuda.h"

namespace cc {

void set_error(const char* fmt, ...);

#define CC_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      cc::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CC_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define CC_CHECK(expr)              \
  do {                              \
    int _s = (expr);                \
    if (_s != CC_OK) return _s;     \
  } while (0)

#define CC_REQUIRE(cond, status, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      cc::set_error(__VA_ARGS__);      \
      return status;                   \
    }                                  \
  } while (0)

constexpr int TILE = 32;   // samples per integral tile (one warp lane per sample)
constexpr int HOG_GRID = 4;  // HOG block step: every cell corner of the catalog is a multiple of it (src/HOGfeatures.cpp:67-106)
constexpr int CC_MAX_RANKS = 72;   // ranks of one communicator (one NVLink domain: 8 on a B200 box, 72 on an NVL72 rack)

// ---- host-side catalogs -------------------------------------------------------------------
struct RectI { int x, y, w, h; };
struct HaarFeatureH { int tilted; RectI r[3]; float wt[3]; int shape, x, y, dx, dy; };   // shape: 0..7 upright (x2,y2,x3,y3,x4,y4,x2y2,cs), 8..13 tilted
struct LbpFeatureH { RectI r; };
struct HogFeatureH { RectI cell0; };

// A Haar feature flattened to unique integral-image taps: value = sum coef[k] * I[off[k]].
// Packed for warp-uniform 128-bit loads: tap k = off (bits 0-15) | coef as int8 (bits 16-23);
// word 0 additionally carries ntaps in bits 24-27 and the tilted flag in bit 28.
constexpr int HAAR_MAX_TAPS = 12;
struct __align__(16) HaarTaps {
  uint32_t w[HAAR_MAX_TAPS];
#if defined(__CUDACC__)
  __host__ __device__
#endif
  int ntaps() const { return (int)((w[0] >> 24) & 15u); }
#if defined(__CUDACC__)
  __host__ __device__
#endif
  int tilted() const { return (int)((w[0] >> 28) & 1u); }
};
static_assert(sizeof(HaarTaps) == 48, "HaarTaps layout");
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int tap_off(uint32_t w) { return (int)(w & 0xFFFFu); }
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int tap_coef(uint32_t w) { return (int)(int8_t)((w >> 16) & 0xFFu); }

// Work item of the separable Haar precalc: all features sharing (x, dx) [kind 0], (y, dy) [kind 1], or, for tilted
// features in rotated coordinates a = x + y, b = y - x, (a0, dx) [kind 2] or (b0 + W, dy) [kind 3]
struct SepPattern { uint8_t kind, p, d, flags; uint8_t lo, hi, pad0, pad1; };   // flags: bit t set = shape t has local features
static_assert(sizeof(SepPattern) == 8, "SepPattern layout");

struct LbpTaps { uint16_t off[16]; };           // 4x4 corner grid
struct HogTaps { uint16_t off[4][4]; };         // 4 cells x 4 corners

std::vector<HaarFeatureH> make_haar_catalog(int W, int H, int mode);
std::vector<LbpFeatureH> make_lbp_catalog(int W, int H);
std::vector<HogFeatureH> make_hog_catalog(int W, int H);
HaarTaps make_haar_taps(const HaarFeatureH& f, int stride);
LbpTaps make_lbp_taps(const LbpFeatureH& f, int stride);
HogTaps make_hog_taps(const HogFeatureH& f, int stride, int grid);      // grid: the planes keep every grid-th row / column (coordinates are divided by it)

// ---- timing -------------------------------------------------------------------------------
struct TimerSlot { double ms = 0; int64_t launches = 0; };

// flat device-side tree used by predict kernels
struct DevNode {
  int left, right, var_idx;
  float c;
  int subset[8];
  double value;
};

struct DevStageRef { int first_tree, num_trees; float threshold; };

struct BoostBuffers;
struct MineState;

// Device storage of one set of evaluator slots (the training set, or the mining scratch batch)
struct EvalStore {
  int max_samples = 0, ntiles = 0, ns_pad = 0;
  uint8_t* d_imgs = nullptr;       // [ns_pad][H][W]
  uint8_t* d_labels = nullptr;     // [ns_pad]
  int32_t* d_sum = nullptr;        // tiles [ntiles][P][32]
  int32_t* d_tilted = nullptr;     // tiles (Haar ALL only)
  float* d_nf = nullptr;           // [ns_pad]
  float* d_hog = nullptr;          // [10][P][ns_pad]
};

}  // namespace cc

struct cc_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::map<std::string, cc::TimerSlot> timers;
  int64_t launches = 0;
  bool timing_enabled = true;

  // ---- evaluator ----
  int ftype = -1, haar_mode = 0, W = 0, H = 0, P = 0, max_samples = 0, ntiles = 0;
  int num_features = 0, feat_size = 1, max_cat = 0, var_count = 0;
  std::vector<cc::HaarFeatureH> haar;
  std::vector<cc::LbpFeatureH> lbp;
  std::vector<cc::HogFeatureH> hog;
  uint8_t* d_imgs = nullptr;       // [max_samples][H][W]
  uint8_t* d_labels = nullptr;     // [max_samples]
  int32_t* d_sum = nullptr;        // tiles [ntiles][P][32]
  int32_t* d_tilted = nullptr;     // tiles (Haar ALL only)
  float* d_nf = nullptr;           // [ntiles*32]
  float* d_hog = nullptr;          // [10][hogP][ns_pad]  sample-fastest; HOG planes keep only the rows / columns features can touch (multiples of HOG_GRID)
  int hogS = 0, hogP = 0;          // compact plane: row stride W / HOG_GRID + 1, entries hogS * (H / HOG_GRID + 1)
  int ns_pad = 0;                  // ntiles*32
  cc::HaarTaps* d_haar_taps = nullptr;   // [num_features]
  cc::LbpTaps* d_lbp_taps = nullptr;
  cc::HogTaps* d_hog_taps = nullptr;
  std::vector<int> upright_rows, tilted_rows;   // global feature indices by plane (Haar)
  cc::HaarTaps* d_taps_upright = nullptr;   // compact per-plane tap lists of the local shard
  cc::HaarTaps* d_taps_tilted = nullptr;
  // separable upright precalc (cc_traindata.cu): feature-index tables and pattern list of the local shard
  int32_t* d_sep_tab_row = nullptr;     // [5][W][W][H][HP]
  int32_t* d_sep_tab_col = nullptr;     // [3][H][H][W][WP]
  int32_t* d_sep_tab_t1 = nullptr;      // tilted, long side along a: [3][W+H][W][W][HP]
  int32_t* d_sep_tab_t2 = nullptr;      // tilted, long side along b: [3][W+H+1][H][W][WP]
  cc::SepPattern* d_sep_patterns = nullptr;   // upright patterns, then tilted patterns
  int n_sep_patterns = 0;               // upright
  int n_sep_tilted = 0;
  bool use_sep = false;

  // ---- train data ----
  int shard_rank = 0, shard_world = 1;
  int row0 = 0, nrows = 0;         // local variable slice
  int N = 0, n_stride = 0;         // samples of the current stage, padded row stride
  bool precalc_valid = false;
  float* d_fval = nullptr;         // [nrows][n_stride] (Haar / HOG)
  uint8_t* d_fcode = nullptr;      // [nrows][n_stride] (LBP)
  uint16_t* d_sidx = nullptr;      // [nrows][n_stride]
  float* d_sval = nullptr;         // [nrows][n_stride] values in sorted order
  uint32_t* d_sidx32 = nullptr;    // wide path (numSamples beyond what one CTA sorts in shared memory): 32-bit sorted-index rows, no d_sidx / d_sval
  bool wide = false;               // the current stage uses the wide path
  bool cap_wide = false;           // what the buffers were allocated for
  int force_wide = 0;              // cc_boost_set_search_mode(ctx, 2): wide kernels regardless of numSamples (tests)
  int* d_hard_rows = nullptr;      // [1 + nrows]: count, then the rows the bucket sort deferred to the radix kernel
  int sort_mode = 0;               // cc_boost_set_sort_mode
  size_t cap_rows = 0, cap_stride = 0;

  // ---- boosting state (device) ----
  cc::BoostBuffers* boost = nullptr;

  // ---- mining ----
  cc::MineState* mine = nullptr;

  // ---- NCCL ----
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_world = 1;
};

#if defined(__CUDACC__)
namespace cc {
// cv::fastAtan32f polynomial (OpenCV 4.x core/mathfuncs_core.simd.hpp), radians; explicit _rn
// intrinsics keep the operation order free of FMA contraction.
__device__ __forceinline__ float cv_atan2_rad(float y, float x) {
  const float deg = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * deg, p3 = -0.3258083974640975f * deg, p5 = 0.1555786518463281f * deg,
              p7 = -0.04432655554792128f * deg;
  float ax = fabsf(x), ay = fabsf(y), a, c, c2;
  const float eps = (float)2.2204460492503131e-16;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return __fmul_rn(a, (float)(3.14159265358979323846 / 180.0));
}

// gradient magnitude and orientation bin of one pixel from its central differences (src/HOGfeatures.cpp:187-222: cartToPolar,
// bin = floor(angle * 9 / pi - 0.5) wrapped into [0, 9)); shared by K1 (k_hog_gradient) and the HOG miner (cc_mine.cu)
__device__ __forceinline__ void hog_mag_bin(float dx, float dy, float& mag, int& b) {
  mag = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  float angle = cv_atan2_rad(dy, dx);
  angle = __fsub_rn(__fmul_rn(angle, (float)(9 / 3.14159265358979323846)), 0.5f);
  b = (int)floorf(angle);
  if (b < 0) b += 9; else if (b >= 9) b -= 9;
}
}  // namespace cc
#endif

namespace cc {
// timing helpers: bracket a kernel family with events on ctx->stream
struct ScopedTimer {      // do not nest (the inner interval would be counted twice)
  cc_ctx* ctx;
  const char* name;
  int n;
  int idx = -1;             // position of this timer's event pair in the pending list
  ScopedTimer(cc_ctx* c, const char* nm, int launches);
  ~ScopedTimer();
};

int eval_free(cc_ctx* ctx);
int boost_free(cc_ctx* ctx);
int mine_free(cc_ctx* ctx);
EvalStore main_store(const cc_ctx* ctx);
int store_alloc(cc_ctx* ctx, EvalStore& st, int max_samples);
void store_free(EvalStore& st);
int launch_integrals(cc_ctx* ctx, const EvalStore& st, int first, int n);
}  // namespace cc

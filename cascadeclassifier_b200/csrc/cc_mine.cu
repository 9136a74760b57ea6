// K5: hard-negative mining.  CvCascadeClassifier::updateTrainingSet / fillPassedSamples / predict
// (src/cascadeclassifier.cpp:297-357) with the window stream of CvCascadeImageReader::NegReader
// (src/imagestorage.cpp:23-126) and PosReader (:138-189).
//
// The reference consumes one window at a time: copy the crop, build its integral images in slot i,
// run the cascade, keep the slot if it passes.  Here the same stream is evaluated speculatively in
// batches: the host enumerates pyramid levels arithmetically (the NegReader state machine without the
// pixels), the GPU (1) builds the INTER_LINEAR_EXACT levels, (2) gathers every window of the batch into
// scratch evaluator slots, (3) builds their integral images with the training kernels (K1) — i.e. with
// exactly the per-window arithmetic of setImage — and (4) runs the cascade with early exit, appending
// surviving stream ordinals.  The host then applies the serial stopping rule to the ordered survivors,
// copies the accepted windows into the training slots and leaves the reader state right behind the last
// consumed window.  Evaluating windows past the stopping point is harmless: they are simply not consumed.
// That literal form is kept as mode 2; by default Haar / LBP windows are evaluated from integral planes of whole pyramid
// levels (k_lvl_build / k_lvl_cascade, with the survivor cache of the window stream) and HOG windows by one CTA per window
// with everything in shared memory (k_hog_window_cascade).
#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>
#include <unordered_map>

#include "cc_internal.h"

namespace cc {

struct EvalView {
  int ftype, P, ns_pad;
  const int32_t* sum;
  const int32_t* tilted;
  const float* nf;
  const float* hog;
  const HaarTaps* haar_taps;
  const LbpTaps* lbp_taps;
  const HogTaps* hog_taps;
};
EvalView make_eval_view_store(const cc_ctx* ctx, const EvalStore& st);
int allreduce_sum(cc_ctx* ctx, void* d_buf, size_t count, int elem_bytes);   // cc_comm.cu: in-place NCCL sum of uint8 / uint32 on ctx->stream

// ---- device-side feature evaluation (same arithmetic as cc_eval.cu / cc_boost.cu) ----------------
__device__ __forceinline__ float m_haar_value(const HaarTaps& t, const EvalView& e, int s) {
  const int32_t* base = (t.tilted() ? e.tilted : e.sum) + (size_t)(s >> 5) * e.P * TILE + (s & 31);
  int acc = 0;
  const int nt = t.ntaps();
  for (int k = 0; k < nt; ++k) acc += tap_coef(t.w[k]) * base[(size_t)tap_off(t.w[k]) * TILE];
  const float n = e.nf[s];
  return n == 0.f ? 0.f : __fdiv_rn((float)acc, n);
}
__device__ __forceinline__ int m_lbp_code(const LbpTaps& t, const EvalView& e, int s) {
  const int32_t* base = e.sum + (size_t)(s >> 5) * e.P * TILE + (s & 31);
  int v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = base[(size_t)t.off[k] * TILE];
  auto cell = [&](int gx, int gy) { const int a = gy * 4 + gx; return v[a] - v[a + 1] - v[a + 4] + v[a + 5]; };
  const int c = cell(1, 1);
  return (cell(0, 0) >= c ? 128 : 0) | (cell(1, 0) >= c ? 64 : 0) | (cell(2, 0) >= c ? 32 : 0) |
         (cell(2, 1) >= c ? 16 : 0) | (cell(2, 2) >= c ? 8 : 0) | (cell(1, 2) >= c ? 4 : 0) |
         (cell(0, 2) >= c ? 2 : 0) | (cell(0, 1) >= c ? 1 : 0);
}
__device__ __forceinline__ float m_hog_value(const HogTaps& t, const EvalView& e, int comp, int s) {
  const int bin = comp % 9, cell = comp / 9;
  const size_t np = (size_t)e.ns_pad;
  const float* ph = e.hog + (size_t)bin * e.P * np + s;
  const float* pn = e.hog + (size_t)9 * e.P * np + s;
  float res = __fadd_rn(__fsub_rn(__fsub_rn(ph[t.off[cell][0] * np], ph[t.off[cell][1] * np]), ph[t.off[cell][2] * np]), ph[t.off[cell][3] * np]);
  float nrm = __fadd_rn(__fsub_rn(__fsub_rn(pn[t.off[0][0] * np], pn[t.off[1][1] * np]), pn[t.off[2][2] * np]), pn[t.off[3][3] * np]);
  return res > 0.001f ? __fdiv_rn(res, __fadd_rn(nrm, 0.001f)) : 0.f;
}
__device__ __forceinline__ float m_eval_var(const EvalView& e, int var, int s) {
  if (e.ftype == CC_FEATURE_HAAR) return m_haar_value(e.haar_taps[var], e, s);
  if (e.ftype == CC_FEATURE_LBP) return (float)m_lbp_code(e.lbp_taps[var], e, s);
  return m_hog_value(e.hog_taps[var / 36], e, var % 36, s);
}

// ---- batch description ------------------------------------------------------------------------------
struct LevelDesc {
  const uint8_t* src;      // background image in HBM
  int sw, sh;              // source size
  int lw, lh;              // level size
  double scale_x, scale_y; // 1.0 / (lw / (double)sw), as cv::resize computes it
  long long px_off;        // offset of this level in the batch level-pixel pool
  int x0, y0, step_x, step_y, nx, ny;
  long long w_skip;        // windows of this level that precede the batch (partial first level)
  long long w_base;        // batch ordinal of the first window taken from this level
  long long w_count;       // windows taken
  // whole-level integral planes (Haar / LBP fast path): element offset of this level's (lh+1) x (lw+1) plane in
  // each plane pool, and cumulative line counts that map a thread to (level, row | column | diagonal)
  long long ip_off;
  int rows_base, cols_base, diag_base;
  int owned;               // image-sharded mining: this rank builds and evaluates the level (1) or skips it (0)
  int row0;                // fast path: this entry is a horizontal strip of its level starting at level row row0 (0: whole level)
  int pitch;               // row pitch of the integral planes (>= lw + 1; the fast path pads it to a multiple of 4)
  int ppitch;              // row pitch of the level pixels in the pool (>= lw)
  int spitch;              // row pitch of the source image in HBM (>= sw, a multiple of 16: rows can be staged with 16-byte copies)
};

// cv::resize(src, dst, INTER_LINEAR_EXACT), 8-bit single channel: 8.8 fixed-point coefficients per axis,
// horizontal pass to 8.8, vertical to 16.16, round by +32768 >> 16.  Un-fused f64.
__device__ __forceinline__ void lin_coef(double scale, int v, int src, int& ofs, int& c0, int& c1, int& mode) {
  double f = __dsub_rn(__dmul_rn(scale, (double)v + 0.5), 0.5);
  const int i = (int)floor(f);
  f = __dsub_rn(f, (double)i);
  if (i >= 0 && src > 1) {
    if (i < src - 1) { ofs = i; c1 = __double2int_rn(__dmul_rn(f, 256.0)); c0 = 256 - c1; mode = 0; }
    else { ofs = src - 1; c0 = 256; c1 = 0; mode = 2; }
  } else { ofs = 0; c0 = 256; c1 = 0; mode = 1; }
}

__global__ void k_resize_levels(const LevelDesc* __restrict__ levels, int nlevels, long long total_px, uint8_t* __restrict__ pool) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total_px; p += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = nlevels - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (levels[mid].px_off <= p) lo = mid; else hi = mid - 1; }
    const LevelDesc& L = levels[lo];
    const int q = (int)(p - L.px_off);
    const int y = q / L.ppitch, x = q - y * L.ppitch;
    if (x >= L.lw) continue;                       // row padding of the pool
    int ox, cx0, cx1, mx, oy, cy0, cy1, my;
    lin_coef(L.scale_x, x, L.sw, ox, cx0, cx1, mx);
    lin_coef(L.scale_y, L.row0 + y, L.sh, oy, cy0, cy1, my);
    auto hrow = [&](int row) -> uint32_t {
      const uint8_t* s = L.src + (size_t)row * L.spitch;
      if (mx == 1) return (uint32_t)s[0] * 256u;
      if (mx == 2) return (uint32_t)s[L.sw - 1] * 256u;
      return (uint32_t)s[ox] * cx0 + (uint32_t)s[ox + 1] * cx1;
    };
    uint32_t val;
    if (my == 1) val = hrow(0) * 256u;
    else if (my == 2) val = hrow(L.sh - 1) * 256u;
    else val = hrow(oy) * cy0 + hrow(oy + 1) * cy1;
    pool[p] = (uint8_t)((val + 32768u) >> 16);
  }
}

// The same levels, tiled: a CTA takes RZ_ROWS rows x RZ_THREADS columns of one level, so the column coefficients (f64) are
// computed once per column instead of once per pixel and the row coefficients once per row; LevelDesc::rows_base holds the
// level's first tile (resize_tiles()).  Same fixed-point arithmetic, bit-identical pixels (test_hog_miner_matches_slot_path
// compares a path that uses this kernel with one that uses k_resize_levels).
constexpr int RZ_ROWS = 8, RZ_THREADS = 256;
__global__ void __launch_bounds__(RZ_THREADS)
k_resize_tiles(const LevelDesc* __restrict__ levels, int nlevels, int ntiles, uint8_t* __restrict__ pool) {
  __shared__ int s_oy[RZ_ROWS], s_cy0[RZ_ROWS], s_cy1[RZ_ROWS], s_step[RZ_ROWS];
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int lo = 0, hi = nlevels - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (levels[mid].rows_base <= tile) lo = mid; else hi = mid - 1; }
    const LevelDesc& L = levels[lo];
    const int t = tile - L.rows_base;
    const int nxc = (L.lw + RZ_THREADS - 1) / RZ_THREADS;
    const int rb = t / nxc, xc = t - rb * nxc;
    const int y0 = rb * RZ_ROWS, ny = min(RZ_ROWS, L.lh - y0);
    __syncthreads();                                   // the previous tile's row coefficients are no longer read
    if ((int)threadIdx.x < ny) {
      int oy, cy0, cy1, my;
      lin_coef(L.scale_y, L.row0 + y0 + (int)threadIdx.x, L.sh, oy, cy0, cy1, my);
      s_oy[threadIdx.x] = oy; s_cy0[threadIdx.x] = cy0; s_cy1[threadIdx.x] = cy1; s_step[threadIdx.x] = my == 0 ? 1 : 0;
    }
    __syncthreads();
    const int x = xc * RZ_THREADS + (int)threadIdx.x;
    if (x >= L.lw) continue;
    int ox, cx0, cx1, mx;
    lin_coef(L.scale_x, x, L.sw, ox, cx0, cx1, mx);
    // outside the interpolating mode the coefficients are (256, 0) and the offset is the clamped tap: the second tap is read
    // from the same place, which gives k_resize_levels' `tap * 256` without a branch
    const int xb = ox + (mx == 0 ? 1 : 0);
    uint8_t* dst = pool + L.px_off + (size_t)y0 * L.ppitch + x;
    for (int r = 0; r < ny; ++r) {
      const uint8_t* a = L.src + (size_t)s_oy[r] * L.spitch;
      const uint8_t* b = a + (size_t)s_step[r] * L.spitch;
      const uint32_t ha = (uint32_t)a[ox] * cx0 + (uint32_t)a[xb] * cx1, hb = (uint32_t)b[ox] * cx0 + (uint32_t)b[xb] * cx1;
      dst[(size_t)r * L.ppitch] = (uint8_t)((ha * (uint32_t)s_cy0[r] + hb * (uint32_t)s_cy1[r] + 32768u) >> 16);
    }
  }
}
// fills LevelDesc::rows_base of levels [l0, l1] with the level's first tile of k_resize_tiles (launched on that range); returns the
// number of tiles
static int resize_tiles(std::vector<LevelDesc>& levels, int l0, int l1) {
  long long n = 0;
  for (int l = l0; l <= l1; ++l) {
    LevelDesc& L = levels[l];
    L.rows_base = (int)n;
    n += (long long)((L.lh + RZ_ROWS - 1) / RZ_ROWS) * ((L.lw + RZ_THREADS - 1) / RZ_THREADS);
  }
  return (int)std::min<long long>(n, INT_MAX);
}

// copy every window of the batch out of its pyramid level into consecutive slots dst[first .. first+n)
__global__ void k_gather_windows(const LevelDesc* __restrict__ levels, int nlevels, const uint8_t* __restrict__ pool, int W, int H,
                                 long long nwin, uint8_t* __restrict__ dst) {
  const int npx = W * H;
  const long long total = nwin * npx;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / npx;
    const int r = (int)(t - b * npx);
    int lo = 0, hi = nlevels - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (levels[mid].w_base <= b) lo = mid; else hi = mid - 1; }
    const LevelDesc& L = levels[lo];
    const long long k = b - L.w_base + L.w_skip;
    const int iy = (int)(k / L.nx), ix = (int)(k - (long long)iy * L.nx);
    const int wy = r / W, wx = r - wy * W;
    dst[t] = pool[L.px_off + (size_t)(L.y0 + iy * L.step_y + wy) * L.ppitch + (L.x0 + ix * L.step_x + wx)];
  }
}

// CvCascadeClassifier::predict(sampleIdx): every stage's CvCascadeBoost::predict with early exit
// (src/cascadeclassifier.cpp:297-306, src/boost.cpp:461-477, src/o_cvcascadeboosttree.cpp:16-39)
__global__ void k_cascade_eval(EvalView e, const DevNode* __restrict__ nodes, const int* __restrict__ tree_first,
                               const DevStageRef* __restrict__ stages, int nstages, int n, unsigned int* __restrict__ surv_count,
                               unsigned int* __restrict__ surv, unsigned int surv_cap) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  bool pass = true;
  for (int st = 0; st < nstages && pass; ++st) {
    const DevStageRef sr = stages[st];
    double sum = 0;
    for (int t = 0; t < sr.num_trees; ++t) {
      const DevNode* tree = nodes + tree_first[sr.first_tree + t];
      int nd = 0;
      while (tree[nd].left >= 0) {
        const float val = m_eval_var(e, tree[nd].var_idx, s);
        bool left;
        if (e.ftype == CC_FEATURE_LBP) { const int c = (int)val; left = (tree[nd].subset[c >> 5] & (1 << (c & 31))) != 0; }
        else left = val <= tree[nd].c;
        nd = left ? tree[nd].left : tree[nd].right;
      }
      sum += tree[nd].value;
    }
    if (sum < (double)(sr.threshold - 0.00001F)) pass = false;
  }
  // ordered compaction inside the warp (ballot), one atomic per warp; the host sorts the short list
  const unsigned m = __ballot_sync(__activemask(), pass);
  if (m) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(surv_count, (unsigned)__popc(m));
    base = __shfl_sync(__activemask(), base, leader);
    if (pass) {
      const unsigned pos = base + __popc(m & ((1u << lane) - 1u));
      if (pos < surv_cap) surv[pos] = (unsigned)s;
    }
  }
}

// Cached survivors (see PyrCache below): slot s holds a window that is known to pass stages [0, first_stage[s]); run the
// remaining stages and leave one flag per slot (the candidates are in stream order, so the flags are too).
__global__ void k_cascade_flags(EvalView e, const DevNode* __restrict__ nodes, const int* __restrict__ tree_first,
                                const DevStageRef* __restrict__ stages, int nstages, int n, const uint8_t* __restrict__ first_stage,
                                uint8_t* __restrict__ flags) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  bool pass = true;
  for (int st = first_stage[s]; st < nstages && pass; ++st) {
    const DevStageRef sr = stages[st];
    double sum = 0;
    for (int t = 0; t < sr.num_trees; ++t) {
      const DevNode* tree = nodes + tree_first[sr.first_tree + t];
      int nd = 0;
      while (tree[nd].left >= 0) {
        const float val = m_eval_var(e, tree[nd].var_idx, s);
        bool left;
        if (e.ftype == CC_FEATURE_LBP) { const int c = (int)val; left = (tree[nd].subset[c >> 5] & (1 << (c & 31))) != 0; }
        else left = val <= tree[nd].c;
        nd = left ? tree[nd].left : tree[nd].right;
      }
      sum += tree[nd].value;
    }
    if (sum < (double)(sr.threshold - 0.00001F)) pass = false;
  }
  flags[s] = pass ? 1 : 0;
}

__global__ void k_copy_slots(const uint8_t* __restrict__ src, const unsigned int* __restrict__ which, int n, int npx, uint8_t* __restrict__ dst,
                             uint8_t* __restrict__ labels, int label) {
  const long long total = (long long)n * npx;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / npx);
    const int r = (int)(t - (long long)i * npx);
    dst[t] = src[(size_t)which[i] * npx + r];
    if (r == 0) labels[i] = (uint8_t)label;
  }
}

// ------------------------------------------------------------------------------------------------
// HOG miner: one CTA per window, nothing per window in HBM.
//
// CvHOGEvaluator::integralHistogram works on the window crop (src/HOGfeatures.cpp:163-256): central differences with
// BORDER_REPLICATE at the crop border, magnitude + orientation bin per pixel, then per bin an f32 integral image whose
// additions run from the crop's first column / first row -- so neither the gradients on the crop's 1-pixel ring nor the integrals
// can be shared between windows.  k_hog_window_cascade copies the window's pixels from its pyramid level into shared memory,
// takes magnitude and bin of every pixel from a table over all 511 x 511 (dx, dy) pairs (differences of 8-bit pixels; the
// table is filled once by k_hog_lut_init with hog_mag_bin, the arithmetic of K1's k_hog_gradient, so the values are
// bit-identical by construction and the division / square root / arctangent polynomial are not re-evaluated per pixel),
// builds the ten compact integral planes in shared memory with the additions of k_hog_integrals in the same order (row
// prefix, then column accumulation) and runs the cascade on them with early exit.  The slot path writes and re-reads ~70 KB
// per 48x96 window (pixels, gradients, planes); this kernel reads the 4.6 KB of pixels.  Survivors == the slot path's
// (test_hog_miner_matches_slot_path), which is the literal setImage + predict arithmetic.
// ------------------------------------------------------------------------------------------------
constexpr int HOG_LUT_W = 512;                                    // row stride of the table; entry [(dy + 255) * 512 + dx + 255]
constexpr int HOG_LUT_N = 511 * HOG_LUT_W;
__global__ void k_hog_lut_init(float2* __restrict__ lut) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HOG_LUT_N) return;
  const int dy = i / HOG_LUT_W - 255, dx = i % HOG_LUT_W - 255;
  float mag = 0.f;
  int b = 0;
  if (dx <= 255) hog_mag_bin((float)dx, (float)dy, mag, b);
  lut[i] = make_float2(mag, __int_as_float(b));
}

// shared-memory geometry of k_hog_window_cascade (strides chosen odd in 32-bit words: a thread per row walks its row
// without bank conflicts)
struct HogWinGeom {
  int W, H;
  int ms;          // row stride of the magnitude tile, floats (odd)
  int bs;          // row stride of the bin tile, bytes (a multiple of 4, odd in words)
  int ss;          // row stride of the row-prefix array, floats (odd)
  int Sc, Pc;      // compact plane: row stride W / HOG_GRID + 1, entries per plane
  int groups, rows;
  int spb;         // bytes per staged pixel row: the aligned 32-bit words that cover W bytes at any byte offset (odd in words)
  int off_bin, off_S, off_px;
  size_t smem;
};
static HogWinGeom hog_win_geometry(int W, int H) {
  HogWinGeom g;
  g.W = W; g.H = H;
  g.ms = W | 1;
  g.bs = (((W + 3) / 4) | 1) * 4;
  g.groups = W / HOG_GRID; g.rows = (H / HOG_GRID) * HOG_GRID;
  g.ss = (g.groups * 10) | 1;
  g.Sc = W / HOG_GRID + 1; g.Pc = g.Sc * (H / HOG_GRID + 1);
  const size_t tile = (size_t)H * g.ms * 4;
  g.off_bin = (int)((tile + 15) & ~(size_t)15);
  g.off_S = (int)(((size_t)g.off_bin + (size_t)H * g.bs + 15) & ~(size_t)15);
  g.spb = (((W + 6) / 4) | 1) * 4;
  // (the compact planes, 10 * Pc floats, later overwrite the row-prefix array: never larger than it, checked here)
  g.off_px = (int)(((size_t)g.off_S + std::max((size_t)g.rows * g.ss * 4, (size_t)10 * g.Pc * 4) + 15) & ~(size_t)15);
  g.smem = (size_t)g.off_px + (size_t)H * g.spb;
  return g;
}

constexpr int HOGW_THREADS = 256;
constexpr int HOGW_GB = 4;            // rows per table-lookup batch of a column thread
constexpr int HOGW_BLK = 64;          // windows per entry of the window -> level table
// run += g if the pixel's bin is `want` -- one predicated FADD (written in PTX: the compiler turns the C++ form of nine
// mutually exclusive tests into a divergent decision tree)
__device__ __forceinline__ void hog_add_if(float& run, uint32_t q, uint32_t want, float gg) {
  asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, %2;\n\t@p add.rn.f32 %0, %0, %3;\n\t}" : "+f"(run) : "r"(q), "r"(want), "f"(gg));
}
// level of every HOGW_BLK-th window of the batch, so that a window finds its level with one load and a short forward walk
__global__ void k_hog_block_levels(const LevelDesc* __restrict__ levels, int nlevels, long long nwin, int* __restrict__ blk) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b * HOGW_BLK >= nwin) return;
  const long long wi = b * HOGW_BLK;
  int lo = 0, hi = nlevels - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (levels[mid].w_base <= wi) lo = mid; else hi = mid - 1; }
  blk[b] = lo;
}
struct HogWinDesc { long long org; int pp, pad; };      // first pixel of the window in the level pool, row pitch of its level
__device__ __forceinline__ HogWinDesc hog_win_locate(const LevelDesc* __restrict__ levels, int nlevels, const int* __restrict__ blk, long long wi) {
  int lo = blk[wi / HOGW_BLK];
  while (lo + 1 < nlevels && levels[lo + 1].w_base <= wi) ++lo;
  const LevelDesc& L = levels[lo];
  const int k = (int)(wi - L.w_base + L.w_skip);
  const int iy = k / L.nx, ix = k - iy * L.nx;
  HogWinDesc d;
  d.pp = L.ppitch; d.pad = 0;
  d.org = L.px_off + (long long)(L.y0 + iy * L.step_y) * L.ppitch + (L.x0 + ix * L.step_x);
  return d;
}
// shared-memory accesses of the gradient pass through 32-bit shared-window addresses (plain LDS / STS with one address add)
__device__ __forceinline__ int hog_lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return (int)v; }
__device__ __forceinline__ void hog_sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void hog_sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
// a column thread of the gradient pass: walks down column x of a segment of rows; the pixel above / below slides through registers
struct HogGradCol {
  uint32_t row;         // shared address of the current pixel row (before its byte shift)
  int sh;               // byte shift of the current row: (address of its first pixel in the pool) & 3
  int prev, cur;        // pixel (y - 1, x) with the border replicated, pixel (y, x)
  uint32_t sm, sb;      // shared addresses of smag[y][x], sbin[y][x]
  int y;
};
template <int GB>
__device__ __forceinline__ void hog_grad_rows(HogGradCol& st, int x, int xm, int xp, int p3, int H, int spb, int ms4, int bs,
                                              const float2* __restrict__ lut) {
  int idx[GB];
#pragma unroll
  for (int r = 0; r < GB; ++r) {
    const int lf = hog_lds_u8(st.row + st.sh + xm), rt = hog_lds_u8(st.row + st.sh + xp);
    const bool last = st.y + 1 >= H;                                  // last row: the pixel below is the pixel itself
    const int shn = (st.sh + p3) & 3;
    const int nxt = hog_lds_u8(last ? st.row + st.sh + x : st.row + spb + shn + x);
    idx[r] = (nxt - st.prev + 255) * HOG_LUT_W + (rt - lf + 255);
    st.prev = st.cur; st.cur = nxt; st.row += spb; st.sh = shn; ++st.y;
  }
  float2 e[GB];
#pragma unroll
  for (int r = 0; r < GB; ++r) e[r] = __ldg(lut + idx[r]);
#pragma unroll
  for (int r = 0; r < GB; ++r) {
    hog_sts_f32(st.sm, e[r].x);
    hog_sts_u8(st.sb, (uint32_t)__float_as_int(e[r].y));
    st.sm += ms4; st.sb += bs;
  }
}

// cascade on the compact planes of one window (a warp): lane t walks weak tree t of the stage, the leaf values are added in tree order
__device__ __forceinline__ bool hog_window_predict(const float* planes, int Pc, const HogTaps* __restrict__ taps, const DevNode* __restrict__ nodes,
                                                   const int* __restrict__ tree_first, const DevStageRef* __restrict__ stages, int nstages, int lane) {
  EvalView e;
  e.ftype = CC_FEATURE_HOG; e.P = Pc; e.ns_pad = 1;
  e.sum = nullptr; e.tilted = nullptr; e.nf = nullptr; e.hog = planes;
  e.haar_taps = nullptr; e.lbp_taps = nullptr; e.hog_taps = taps;
  bool pass = true;
  for (int st = 0; st < nstages && pass; ++st) {
    const DevStageRef sr = stages[st];
    double sum = 0;
    for (int t0 = 0; t0 < sr.num_trees; t0 += 32) {
      double v = 0;
      if (t0 + lane < sr.num_trees) {
        const DevNode* tree = nodes + tree_first[sr.first_tree + t0 + lane];
        int nd = 0;
        while (tree[nd].left >= 0) {
          const float val = m_hog_value(taps[tree[nd].var_idx / 36], e, tree[nd].var_idx % 36, 0);
          nd = val <= tree[nd].c ? tree[nd].left : tree[nd].right;
        }
        v = tree[nd].value;
      }
      const int cnt = min(32, sr.num_trees - t0);
      for (int i = 0; i < cnt; ++i) sum += __shfl_sync(0xffffffffu, v, i);
    }
    if (sum < (double)(sr.threshold - 0.00001F)) pass = false;
  }
  return pass;
}

constexpr int HOGW_MAXQ = 32;         // kept plane rows per window (H / HOG_GRID) held in registers by the column pass
// NQ: H / HOG_GRID at compile time (the column pass is fully unrolled), or 0 for any height up to HOG_GRID * HOGW_MAXQ
template <int NQ>
__global__ void __launch_bounds__(HOGW_THREADS, 3)
k_hog_window_cascade(const LevelDesc* __restrict__ levels, int nlevels, const int* __restrict__ blk, const uint8_t* __restrict__ pool,
                     const float2* __restrict__ lut, HogWinGeom g, const HogTaps* __restrict__ taps, const DevNode* __restrict__ nodes,
                     const int* __restrict__ tree_first, const DevStageRef* __restrict__ stages, int nstages, long long w_begin, long long nwin,
                     unsigned int* __restrict__ surv_count, unsigned int* __restrict__ surv, unsigned int surv_cap) {
  // windows [w_begin, nwin) of the batch (image-sharded mining gives every rank a contiguous slice)
  extern __shared__ __align__(16) unsigned char hw_smem[];
  __shared__ HogWinDesc s_desc[4];                                  // ring: iteration it reads [it & 3], the last thread fills [(it + 2) & 3]
  float* smag = reinterpret_cast<float*>(hw_smem);                  // [H][ms]
  uint8_t* sbin = hw_smem + g.off_bin;                              // [H][bs]
  float* S = reinterpret_cast<float*>(hw_smem + g.off_S);           // [rows][ss]: row prefix of plane b behind pixel group c at [c * 10 + b]
  float* planes = S;                                                // [10][Pc] compact planes of the window, once the column pass has read S
  uint8_t* spx = hw_smem + g.off_px;                                // [H][spb] pixel rows as the aligned words that hold them
  const uint32_t spx_s = (uint32_t)__cvta_generic_to_shared(spx);
  const uint32_t smag_s = (uint32_t)__cvta_generic_to_shared(smag), sbin_s = (uint32_t)__cvta_generic_to_shared(sbin);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = g.W, H = g.H;
  // asynchronous copy of a window's pixel rows into spx (4-byte cp.async: a row starts at any byte, so the words that hold it
  // are copied and the row is read at offset (address & 3)); lands before the next cp.async.wait_group 0
  auto prefetch = [&](const HogWinDesc d) {
    const int nw = g.spb / 4, total = H * nw;
    int row = tid / nw, w = tid - row * nw;
    const int drow = HOGW_THREADS / nw, dw = HOGW_THREADS - drow * nw;
    for (int i = tid; i < total; i += HOGW_THREADS) {
      const long long a = (d.org + (long long)row * d.pp) & ~3ll;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(spx_s + (uint32_t)(row * g.spb + w * 4)), "l"(pool + a + w * 4) : "memory");
      w += dw; row += drow;
      if (w >= nw) { w -= nw; ++row; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto report = [&](long long wi) {      // warp 0: cascade on the planes of window wi
    const bool pass = hog_window_predict(planes, g.Pc, taps, nodes, tree_first, stages, nstages, lane);
    if (pass && lane == 0) {
      const unsigned pos = atomicAdd(surv_count, 1u);
      if (pos < surv_cap) surv[pos] = (unsigned)wi;
    }
  };
  if (tid == 0) {
    s_desc[0] = hog_win_locate(levels, nlevels, blk, w_begin + blockIdx.x);
    if (w_begin + blockIdx.x + gridDim.x < nwin) s_desc[1] = hog_win_locate(levels, nlevels, blk, w_begin + blockIdx.x + gridDim.x);
  }
  __syncthreads();
  prefetch(s_desc[0]);
  int it = 0;
  long long prev_wi = -1;
  for (long long wi = w_begin + blockIdx.x; wi < nwin; wi += gridDim.x, ++it) {
    const HogWinDesc d = s_desc[it & 3];
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // this window's pixels have landed; the previous window's planes are complete
    if (warp == 0) {
      // ---- the cascade of the PREVIOUS window runs beside this window's gradient pass ----
      if (prev_wi >= 0) report(prev_wi);
    } else {
      // ---- magnitude and bin of every pixel: central differences with the window's replicated border, table lookup ----
      const int gt = tid - 32, nseg = (HOGW_THREADS - 32) / W, seg = gt / W, x = gt - seg * W;
      const int rs = (((H + nseg - 1) / nseg + HOGW_GB - 1) / HOGW_GB) * HOGW_GB;
      const int ya = seg * rs, yb = min(H, ya + rs);
      if (seg < nseg && ya < yb) {
        const int xm = max(x - 1, 0), xp = min(x + 1, W - 1);
        const int o3 = (int)(d.org & 3), p3 = d.pp & 3;
        const int yp = max(ya - 1, 0);
        HogGradCol st;
        st.row = spx_s + (uint32_t)(ya * g.spb); st.sh = (o3 + ya * p3) & 3; st.y = ya;
        st.prev = hog_lds_u8(spx_s + (uint32_t)(yp * g.spb + ((o3 + yp * p3) & 3) + x));
        st.cur = hog_lds_u8(st.row + st.sh + x);
        st.sm = smag_s + (uint32_t)((ya * g.ms + x) * 4); st.sb = sbin_s + (uint32_t)(ya * g.bs + x);
        while (st.y + HOGW_GB <= yb) hog_grad_rows<HOGW_GB>(st, x, xm, xp, p3, H, g.spb, g.ms * 4, g.bs, lut);
        while (st.y < yb) hog_grad_rows<1>(st, x, xm, xp, p3, H, g.spb, g.ms * 4, g.bs, lut);
      }
    }
    __syncthreads();
    // the pixel tile and the planes are free: start the copy of the next window; the last thread (idle in the row pass of
    // windows up to 112 rows) looks up the window after that
    {
      const long long wn = wi + gridDim.x;
      if (wn < nwin) prefetch(s_desc[(it + 1) & 3]);
      if (tid == HOGW_THREADS - 1 && wn + gridDim.x < nwin) s_desc[(it + 2) & 3] = hog_win_locate(levels, nlevels, blk, wn + gridDim.x);
    }
    // ---- row prefixes (strSum of src/HOGfeatures.cpp:240-252): two threads walk one row, the first with bins 0..4, the second
    // with bins 5..8 and the norm plane; only the matching bin and the norm plane change at a pixel (the other planes add +0,
    // which leaves an f32 sum unchanged) ----
    for (int t = tid; t < 2 * g.rows; t += HOGW_THREADS) {
      const bool second = t >= g.rows;
      const int y = second ? t - g.rows : t;
      const uint32_t b0 = second ? 5u : 0u;
      float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f;
      const float* mrow = smag + y * g.ms;
      const uint32_t* brow = reinterpret_cast<const uint32_t*>(sbin + y * g.bs);
      float* srow = S + y * g.ss + b0;
      for (int c = 0; c < g.groups; ++c) {
        const uint32_t q4 = brow[c];
#pragma unroll
        for (int kk = 0; kk < HOG_GRID; ++kk) {
          const float gg = mrow[c * HOG_GRID + kk];
          const uint32_t q = (q4 >> (8 * kk)) & 0xFFu;
          hog_add_if(r0, q, b0, gg); hog_add_if(r1, q, b0 + 1u, gg); hog_add_if(r2, q, b0 + 2u, gg); hog_add_if(r3, q, b0 + 3u, gg);
          hog_add_if(r4, q, second ? q : 4u, gg);          // second thread: the norm plane takes every pixel
        }
        srow[c * 10] = r0; srow[c * 10 + 1] = r1; srow[c * 10 + 2] = r2; srow[c * 10 + 3] = r3; srow[c * 10 + 4] = r4;
      }
    }
    __syncthreads();
    // ---- column accumulation: plane[y + 1][4 (c + 1)] = plane[y][4 (c + 1)] + strSum; only rows / columns that are multiples of
    // HOG_GRID are kept (the compact planes of k_hog_integrals).  The planes go where S is, so the kept values wait in registers
    // until every thread has read its column of S ----
    {
      constexpr int QN = NQ ? NQ : HOGW_MAXQ;
      const int nq = NQ ? NQ : g.rows / HOG_GRID;
      const int j = tid;                                 // (use_hog_window_path: groups * 10 <= HOGW_THREADS, one round)
      const bool on = j < g.groups * 10;
      float keep[QN];
      if (on) {
        float col = 0.f;
        const float* sp = S + j;
#pragma unroll
        for (int q = 0; q < QN; ++q) {
          if (NQ || q < nq) {
#pragma unroll
            for (int kk = 0; kk < HOG_GRID; ++kk) { col = __fadd_rn(col, *sp); sp += g.ss; }
            keep[q] = col;
          }
        }
      }
      __syncthreads();
      if (on) {
        const int c = j / 10, b = j - c * 10;
        float* pl = planes + b * g.Pc + (c + 1);
#pragma unroll
        for (int q = 0; q < QN; ++q)
          if (NQ || q < nq) { pl += g.Sc; *pl = keep[q]; }
      }
    }
    {
      const int nr = H / HOG_GRID + 1;                 // first row and first column of every plane are zero
      for (int i = tid; i < 10 * (g.Sc + nr); i += HOGW_THREADS) {
        const int b = i / (g.Sc + nr), e = i - b * (g.Sc + nr);
        planes[b * g.Pc + (e < g.Sc ? e : (e - g.Sc) * g.Sc)] = 0.f;
      }
    }
    prev_wi = wi;
  }
  __syncthreads();
  if (warp == 0 && prev_wi >= 0) report(prev_wi);
}

// ------------------------------------------------------------------------------------------------
// Haar / LBP fast path: integral images of whole pyramid levels instead of one per window.
//
// Sums over a window's rectangles taken from the integral image of the whole level equal those taken
// from the integral image of the 24x24 crop (exact integers; for tilted rectangles the 4-corner
// combination cancels every pixel outside the rectangle point-wise, so it does not matter that the
// level's tilted integral also covers pixels outside the crop) -- SURVEY.md App. A.2, re-checked
// against cv2.integral3 in tests/test_oracle_golden.py.  Planes hold uint32 with wrap-around: every
// window-level difference is < 2^32, so modular arithmetic returns the exact value (sqsum included,
// which the reference keeps in f64).  Per window the arithmetic after the rectangle sums is that of
// setImage + operator(): nf = (float)sqrt(area*sq - s*s) in f64 (src/features.cpp:13-25), value =
// (float)acc / nf in IEEE f32 (include/haarfeatures.h:108-112), LBP code from 9 cell sums.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int level_of_line(const LevelDesc* __restrict__ levels, int nlevels, int line, int which) {
  int lo = 0, hi = nlevels - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    const int b = which == 0 ? levels[mid].rows_base : which == 1 ? levels[mid].cols_base : levels[mid].diag_base;
    if (b <= line) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------
// Fused pyramid builder of the fast path: one CTA per level walks the level row by row and produces, in one
// pass over the source image, the INTER_LINEAR_EXACT pixels (same fixed-point arithmetic as k_resize_levels),
// the sum and sqsum integral rows (block-wide row scan + per-column running sums held in registers) and the
// tilted integral row (D1/D2 diagonal recurrences with the previous row's values in shared memory).  Thread t
// owns pixel columns 2t, 2t+1 and plane columns 2t, 2t+1; one __syncthreads per row (all per-row shared
// buffers are double-buffered).  Levels wider or taller than 2047 take the multi-kernel path below.
// ------------------------------------------------------------------------------------------------
constexpr int LB_MAX_THREADS = 512;
constexpr int LB_PPT = 4;                               // pixels / plane columns per thread
constexpr int LB_MAXDIM = LB_PPT * LB_MAX_THREADS - 1;  // lw + 1 <= 2048 plane columns, lh <= 2047 rows of y coefficients

// THREADS = 128 / 256 / 512 for levels up to 511 / 1023 / 2047 pixels wide: narrow levels get small CTAs, several per SM.
// Thread t owns pixels x = 4t .. 4t+3 and plane columns j = 4t .. 4t+3 (plane column j <-> apex pixel x = j - 1); plane and
// pool rows are padded to multiples of 4, so every row leaves as one 128-bit store per plane and one 32-bit pixel store.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
k_lvl_build(const LevelDesc* __restrict__ levels, const int* __restrict__ order, uint8_t* __restrict__ pool, uint32_t* __restrict__ sum,
            uint32_t* __restrict__ sq, uint32_t* __restrict__ tilted, int sp, int winH) {
  constexpr int NW = THREADS / 32;
  // the two source rows of a level row, staged four level rows ahead with 16-byte cp.async copies: [y & 3][row][sp]
  extern __shared__ __align__(16) uint8_t s_src[];
  __shared__ uint32_t s_bpix[2][THREADS];       // a thread's last pixel of the row (the left neighbour's apex pixel)
  __shared__ uint32_t s_bd1[2][THREADS];        // D1 of the previous row at the thread's last plane column
  __shared__ uint32_t s_bd2[2][THREADS + 1];    // D2 of the previous row at the thread's first plane column
  __shared__ int s_yo[LB_MAXDIM + 1];           // source row | c1 << 16
  __shared__ uint32_t s_wt[2][2][32];
  const LevelDesc L = levels[order[blockIdx.x]];      // tallest strips first: a strip is one serial chain of rows
  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int lw = L.lw, lh = L.lh, sw = L.sw, sh = L.sh, pitch = L.pitch, spitch = L.spitch;
  const uint8_t* __restrict__ src = L.src;
  const int x0 = LB_PPT * t;
  // per-column x coefficients (registers: source column | c1 << 16), per-row y coefficients (shared)
  int xc[LB_PPT];
#pragma unroll
  for (int k = 0; k < LB_PPT; ++k) {
    int o = 0, c0 = 256, c1 = 0, md = 0;
    if (x0 + k < lw) lin_coef(L.scale_x, x0 + k, sw, o, c0, c1, md);
    xc[k] = o | (c1 << 16);
  }
  for (int y = t; y < lh; y += THREADS) {
    int o, c0, c1, md;
    lin_coef(L.scale_y, L.row0 + y, sh, o, c0, c1, md);
    // The sqsum plane is only read by calcNormFactor's four corners (src/features.cpp:13-25): plane rows 1 and winH - 1 below a window's
    // origin row.  Level row y produces plane row y + 1; bit 31 marks the rows some window needs -- the others are not stored
    // (5 of 6 rows for 24-pixel windows at step 12), which takes a third off the builder's write traffic.
    const int a = y - L.y0, b = y + 2 - winH - L.y0, sy = max(1, L.step_y);
    const bool need = (a >= 0 && a % sy == 0 && a / sy < L.ny) || (b >= 0 && b % sy == 0 && b / sy < L.ny);
    s_yo[y] = o | (c1 << 16) | (need ? (int)0x80000000u : 0);
  }
  s_bpix[0][t] = s_bpix[1][t] = 0; s_bd1[0][t] = s_bd1[1][t] = 0; s_bd2[0][t] = s_bd2[1][t] = 0;
  if (t == 0) { s_bd2[0][THREADS] = 0; s_bd2[1][THREADS] = 0; }
  if (t < 32) { s_wt[0][0][t] = s_wt[0][1][t] = s_wt[1][0][t] = s_wt[1][1][t] = 0; }
  uint32_t* S = sum + L.ip_off;
  uint32_t* Q = sq ? sq + L.ip_off : nullptr;
  uint32_t* T = tilted ? tilted + L.ip_off : nullptr;
  for (int j = t; j < pitch; j += THREADS) { S[j] = 0; if (Q) Q[j] = 0; if (T) T[j] = 0; }
  __syncthreads();
  const bool on = x0 <= lw;                          // this thread owns at least one plane column
  const bool warp_on = warp * 32 * LB_PPT <= lw;     // warp-uniform
  uint32_t colS[LB_PPT] = {0, 0, 0, 0}, colQ[LB_PPT] = {0, 0, 0, 0}, Tp[LB_PPT] = {0, 0, 0, 0};
  uint32_t d1[LB_PPT] = {0, 0, 0, 0}, d2[LB_PPT] = {0, 0, 0, 0};
  int xa[LB_PPT], xb[LB_PPT];
  bool in[LB_PPT];
#pragma unroll
  uint32_t cx0[LB_PPT], cx1[LB_PPT];                  // horizontal weights of the thread's columns (row-invariant)
  for (int k = 0; k < LB_PPT; ++k) {
    in[k] = x0 + k < lw; xa[k] = xc[k] & 0xFFFF; xb[k] = min(xa[k] + 1, sw - 1);
    cx1[k] = (uint32_t)xc[k] >> 16; cx0[k] = 256u - cx1[k];
  }
  // source bytes of row y are fetched one row ahead
  uint32_t nb[4 * LB_PPT];
#pragma unroll
  for (int k = 0; k < 4 * LB_PPT; ++k) nb[k] = 0;
  const uint32_t s_src_addr = (uint32_t)__cvta_generic_to_shared(s_src);
  const int chunks = spitch >> 4;                     // 16-byte chunks per source row (the pitch is a multiple of 16)
  auto stage = [&](int y) {                           // all threads: source rows of level row y -> buffer y & 3
    const int yo = s_yo[min(y, lh - 1)];
    const int r0 = yo & 0xFFFF, r1 = min(r0 + 1, sh - 1);
    const uint32_t dst = s_src_addr + (uint32_t)((y & 3) * 2 * sp);
    for (int i = t; i < 2 * chunks; i += THREADS) {
      const int rr = i >= chunks ? 1 : 0, c = i - rr * chunks;
      const uint8_t* g = src + (size_t)(rr ? r1 : r0) * spitch + (size_t)c * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + (uint32_t)(rr * sp + c * 16)), "l"(g) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto fetch = [&](int y) {                           // this thread's 16 source bytes of level row y, from the staged rows
    const uint8_t* a = s_src + (y & 3) * 2 * sp;
    const uint8_t* b = a + sp;
#pragma unroll
    for (int k = 0; k < LB_PPT; ++k)
      if (in[k]) { nb[4 * k] = a[xa[k]]; nb[4 * k + 1] = a[xb[k]]; nb[4 * k + 2] = b[xa[k]]; nb[4 * k + 3] = b[xb[k]]; }
  };
  stage(0);                                           // (row indices beyond the level are clamped: every thread commits
  stage(1);                                           //  the same number of groups)
  stage(2);
  stage(3);
  asm volatile("cp.async.wait_group 2;" ::: "memory");   // rows 0 and 1 have landed
  __syncthreads();
  if (lh > 0 && warp_on) fetch(0);
  __syncthreads();                                    // iteration 0 re-fills buffer 0 (row 4): every thread must have read row 0 first
  uint32_t* Srow = S + x0;
  uint32_t* Qrow = Q ? Q + x0 : nullptr;
  uint32_t* Trow = T ? T + x0 : nullptr;
  uint8_t* drow = pool + L.px_off + x0;
  for (int y = 0; y < lh; ++y) {
    const int buf = y & 1;
    uint32_t v[LB_PPT] = {0, 0, 0, 0};
    uint32_t i1 = 0, i2 = 0, tot1 = 0, tot2 = 0;
    Srow += pitch; if (Q) Qrow += pitch; if (T) Trow += pitch;
    stage(y + 4);                                     // buffer y & 3 was consumed by fetch(y) before the previous barrier
    if (warp_on) {
      const uint32_t cy1 = ((uint32_t)s_yo[y] >> 16) & 0x7FFFu, cy0 = 256u - cy1;
#pragma unroll
      for (int k = 0; k < LB_PPT; ++k) {
        const uint32_t c1 = cx1[k], c0 = cx0[k];
        const uint32_t val = ((nb[4 * k] * c0 + nb[4 * k + 1] * c1) * cy0 + (nb[4 * k + 2] * c0 + nb[4 * k + 3] * c1) * cy1 + 32768u) >> 16;
        v[k] = in[k] ? val : 0u;
      }
      if (y + 1 < lh) fetch(y + 1);
      if (pool && in[0]) *reinterpret_cast<uint32_t*>(drow) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);    // pool rows are padded to 4
      s_bpix[buf][t] = v[3];
      // block-wide exclusive scan of the per-thread row totals (pixels and squares)
      tot1 = (v[0] + v[1]) + (v[2] + v[3]);
      tot2 = (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
      i1 = tot1; i2 = tot2;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, i1, d), b = __shfl_up_sync(0xffffffffu, i2, d);
        if (lane >= d) { i1 += a; i2 += b; }
      }
      if (lane == 31) { s_wt[buf][0][warp] = i1; s_wt[buf][1][warp] = i2; }
    }
    drow += L.ppitch;
    asm volatile("cp.async.wait_group 2;" ::: "memory");   // this thread's copies for row y + 2 have landed (rows y + 3, y + 4 may be in flight)
    __syncthreads();
    if (!warp_on) continue;
    uint32_t w1 = (lane < NW && lane < warp) ? s_wt[buf][0][lane] : 0u, w2 = (lane < NW && lane < warp) ? s_wt[buf][1][lane] : 0u;
    w1 = __reduce_add_sync(0xffffffffu, w1); w2 = __reduce_add_sync(0xffffffffu, w2);       // REDUX: totals of the warps to the left
    if (!on) continue;
    // exclusive row prefixes at the thread's plane columns, accumulated down the columns
    {
      uint32_t e = w1 + i1 - tot1;
#pragma unroll
      for (int k = 0; k < LB_PPT; ++k) { colS[k] += e; e += v[k]; }
      *reinterpret_cast<uint4*>(Srow) = make_uint4(colS[0], colS[1], colS[2], colS[3]);
    }
    if (Q) {
      uint32_t e = w2 + i2 - tot2;
#pragma unroll
      for (int k = 0; k < LB_PPT; ++k) { colQ[k] += e; e += v[k] * v[k]; }
      if (s_yo[y] < 0) *reinterpret_cast<uint4*>(Qrow) = make_uint4(colQ[0], colQ[1], colQ[2], colQ[3]);
    }
    if (T) {
      // apex pixels of the plane columns 4t..4t+3 are x = 4t-1 .. 4t+2; D1(y,x) = I + D1(y-1,x-1), D2(y,x) = I + D2(y-1,x+1)
      const uint32_t ivl = t ? s_bpix[buf][t - 1] : 0u;
      const uint32_t d1l = t ? s_bd1[buf ^ 1][t - 1] : 0u;          // D1 of the previous row at column 4t-1
      const uint32_t d2r = s_bd2[buf ^ 1][t + 1];                   // D2 of the previous row at column 4t+4
      const uint32_t iv[LB_PPT] = {ivl, v[0], v[1], v[2]};
      const uint32_t n1[LB_PPT] = {iv[0] + d1l, iv[1] + d1[0], iv[2] + d1[1], iv[3] + d1[2]};
      const uint32_t n2[LB_PPT] = {iv[0] + d2[1], iv[1] + d2[2], iv[2] + d2[3], iv[3] + d2r};
#pragma unroll
      for (int k = 0; k < LB_PPT; ++k) { d1[k] = n1[k]; d2[k] = n2[k]; Tp[k] += n1[k] + n2[k] - iv[k]; }
      s_bd1[buf][t] = d1[3];
      s_bd2[buf][t] = d2[0];
      *reinterpret_cast<uint4*>(Trow) = make_uint4(Tp[0], Tp[1], Tp[2], Tp[3]);
    }
  }
}

// row prefix sums of I and I^2 into rows 1..lh of the sum / sqsum planes (one warp per level row); row 0 and column 0 = 0
__global__ void __launch_bounds__(256)
k_lvl_rows(const LevelDesc* __restrict__ levels, int nlevels, int total_rows, const uint8_t* __restrict__ pool,
           uint32_t* __restrict__ sum, uint32_t* __restrict__ sq) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int gr = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gr < total_rows; gr += nwarps) {
    const LevelDesc& L = levels[level_of_line(levels, nlevels, gr, 0)];
    const int y = gr - L.rows_base, pitch = L.pitch;
    const uint8_t* src = pool + L.px_off + (size_t)y * L.ppitch;
    uint32_t* S = sum + L.ip_off + (size_t)(y + 1) * pitch;
    uint32_t* Q = sq ? sq + L.ip_off + (size_t)(y + 1) * pitch : nullptr;
    if (y == 0) for (int x = lane; x < pitch; x += 32) { S[x - pitch] = 0; if (Q) Q[x - pitch] = 0; }
    if (lane == 0) { S[0] = 0; if (Q) Q[0] = 0; }
    uint32_t c1 = 0, c2 = 0;
    for (int x0 = 0; x0 < L.lw; x0 += 32) {
      const int x = x0 + lane;
      const uint32_t v = x < L.lw ? src[x] : 0u;
      uint32_t p1 = v, p2 = v * v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t1 = __shfl_up_sync(0xffffffffu, p1, d), t2 = __shfl_up_sync(0xffffffffu, p2, d);
        if (lane >= d) { p1 += t1; p2 += t2; }
      }
      p1 += c1; p2 += c2;
      if (x < L.lw) { S[x + 1] = p1; if (Q) Q[x + 1] = p2; }
      c1 = __shfl_sync(0xffffffffu, p1, 31); c2 = __shfl_sync(0xffffffffu, p2, 31);
    }
  }
}

// tilted plane, step 1: E(y, x) = D1(y,x) + D2(y,x) - I(y,x) stored at plane row y+1, column x+1 (x = -1 .. lw-1), with
// D1(y,x) = I(y,x) + D1(y-1,x-1), D2(y,x) = I(y,x) + D2(y-1,x+1) (the two diagonal running sums of cv::integral's tilted
// definition).  One thread walks one diagonal; pass 0 writes D1, pass 1 adds D2 - I.  The plane is zeroed beforehand.
__global__ void __launch_bounds__(256)
k_lvl_diag(const LevelDesc* __restrict__ levels, int nlevels, int total_diag, const uint8_t* __restrict__ pool,
           uint32_t* __restrict__ tilted, int pass) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total_diag; g += gridDim.x * blockDim.x) {
    const LevelDesc& L = levels[level_of_line(levels, nlevels, g, 2)];
    const int k = g - L.diag_base, lw = L.lw, lh = L.lh, pitch = L.pitch, pp = L.ppitch;
    const uint8_t* img = pool + L.px_off;
    uint32_t* T = tilted + L.ip_off;
    // diagonals start on the top row (k < lw: column k) or on a side column (k >= lw: row k - lw + 1)
    int y = k < lw ? 0 : k - lw + 1;
    uint32_t run = 0;
    if (pass == 0) {
      int x = k < lw ? k : 0;
      for (; y < lh && x < lw; ++y, ++x) { run += img[(size_t)y * pp + x]; T[(size_t)(y + 1) * pitch + x + 1] = run; }
    } else {
      int x = k < lw ? k : lw - 1;
      for (; y < lh && x >= 0; ++y, --x) {
        const uint32_t v = img[(size_t)y * pp + x];
        run += v;
        T[(size_t)(y + 1) * pitch + x + 1] += run - v;
      }
      if (y < lh) T[(size_t)(y + 1) * pitch] += run;        // virtual column x = -1: D2(y,-1) = D2(y-1,0), I = 0
    }
  }
}

// column prefix sums in place (one thread per plane column): turns row prefixes into integrals, and E into
// T(Y,X) = T(Y-1,X) + E(Y-1,X-1)
__global__ void __launch_bounds__(256)
k_lvl_cols(const LevelDesc* __restrict__ levels, int nlevels, int total_cols, uint32_t* __restrict__ p0, uint32_t* __restrict__ p1,
           uint32_t* __restrict__ p2) {
  uint32_t* plane = blockIdx.y == 0 ? p0 : blockIdx.y == 1 ? p1 : p2;
  if (!plane) return;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total_cols; g += gridDim.x * blockDim.x) {
    const LevelDesc& L = levels[level_of_line(levels, nlevels, g, 1)];
    const int j = g - L.cols_base, pitch = L.pitch;
    uint32_t* c = plane + L.ip_off + j;
    uint32_t acc = 0;
    int y = 1;
    for (; y + 3 <= L.lh; y += 4) {
      const uint32_t a0 = c[(size_t)y * pitch], a1 = c[(size_t)(y + 1) * pitch], a2 = c[(size_t)(y + 2) * pitch], a3 = c[(size_t)(y + 3) * pitch];
      acc += a0; c[(size_t)y * pitch] = acc;
      acc += a1; c[(size_t)(y + 1) * pitch] = acc;
      acc += a2; c[(size_t)(y + 2) * pitch] = acc;
      acc += a3; c[(size_t)(y + 3) * pitch] = acc;
    }
    for (; y <= L.lh; ++y) { acc += c[(size_t)y * pitch]; c[(size_t)y * pitch] = acc; }
  }
}

// A tree node's feature flattened for level planes: tap = dx | dy << 8 | (int8 coef) << 16 relative to the window origin
struct LvlFeat { uint32_t hdr; uint32_t tap[16]; };      // hdr = ntaps | tilted << 8

struct LvlPlanes { const uint32_t* sum; const uint32_t* sq; const uint32_t* tilted; };

// CvCascadeClassifier::predict for window k of level (strip) L, stages [s_begin, s_end), early exit.
struct CascadeArgs {
  LvlPlanes pl;
  int ftype, W, H;
  const DevNode* nodes;
  const LvlFeat* feats;
  const int* tree_first;
  const DevStageRef* stages;
  int s_begin, s_end;
};

// window normalisation factor: calcNormFactor (src/features.cpp:13-25) from the level planes
__device__ __forceinline__ float window_nf(const uint32_t* __restrict__ S, const uint32_t* __restrict__ Q, int pitch, const CascadeArgs& A) {
  if (A.ftype != CC_FEATURE_HAAR) return 0.f;
  const int W = A.W, H = A.H;
  const int o0 = pitch + 1, o1 = pitch + (W - 1), o2 = (H - 1) * pitch + 1, o3 = (H - 1) * pitch + (W - 1);
  const uint32_t vs = S[o0] - S[o1] - S[o2] + S[o3];
  const uint32_t vq = Q[o0] - Q[o1] - Q[o2] + Q[o3];
  const double area = (double)((W - 2) * (H - 2));
  const double dvs = (double)(int)vs;
  return (float)sqrt(area * (double)vq - dvs * dvs);
}

// CvCascadeBoostTree::predict (src/o_cvcascadeboosttree.cpp:16-39) of the tree whose nodes start at `first`, on the window at S / Tt
__device__ __forceinline__ double window_tree_value(const uint32_t* __restrict__ S, const uint32_t* __restrict__ Tt, int pitch, const CascadeArgs& A,
                                                    int first, float nf) {
  int nd = 0;
  while (A.nodes[first + nd].left >= 0) {
    const DevNode& node = A.nodes[first + nd];
    const LvlFeat& f = A.feats[first + nd];
    bool left;
    if (A.ftype == CC_FEATURE_HAAR) {
      const uint32_t hdr = f.hdr;
      const uint32_t* P = (hdr >> 8) ? Tt : S;
      const int nt = (int)(hdr & 0xFFu);
      int acc = 0;
      for (int q = 0; q < nt; ++q) {
        const uint32_t tp = f.tap[q];
        acc += (int)(int8_t)(tp >> 16) * (int)P[(int)((tp >> 8) & 0xFFu) * pitch + (int)(tp & 0xFFu)];
      }
      const float val = nf == 0.f ? 0.f : __fdiv_rn((float)acc, nf);
      left = val <= node.c;
    } else {
      int v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) { const uint32_t tp = f.tap[q]; v[q] = (int)S[(int)((tp >> 8) & 0xFFu) * pitch + (int)(tp & 0xFFu)]; }
      auto cell = [&](int gx, int gy) { const int a = gy * 4 + gx; return v[a] - v[a + 1] - v[a + 4] + v[a + 5]; };
      const int c = cell(1, 1);
      const int code = (cell(0, 0) >= c ? 128 : 0) | (cell(1, 0) >= c ? 64 : 0) | (cell(2, 0) >= c ? 32 : 0) |
                       (cell(2, 1) >= c ? 16 : 0) | (cell(2, 2) >= c ? 8 : 0) | (cell(1, 2) >= c ? 4 : 0) |
                       (cell(0, 2) >= c ? 2 : 0) | (cell(0, 1) >= c ? 1 : 0);
      left = (node.subset[code >> 5] & (1 << (code & 31))) != 0;
    }
    nd = left ? node.left : node.right;
  }
  return A.nodes[first + nd].value;
}

// S / Q / Tt point at the window's origin inside planes of row pitch `pitch` (global memory: whole-level planes; shared
// memory: the planes of one tile).  Same arithmetic as setImage + operator() + predict of the reference.
__device__ __forceinline__ bool eval_window_at(const uint32_t* __restrict__ S, const uint32_t* __restrict__ Q, const uint32_t* __restrict__ Tt,
                                               int pitch, const CascadeArgs& A) {
  const float nf = window_nf(S, Q, pitch, A);
  for (int st = A.s_begin; st < A.s_end; ++st) {
    const DevStageRef sr = A.stages[st];
    double sum = 0;
    for (int t = 0; t < sr.num_trees; ++t) sum += window_tree_value(S, Tt, pitch, A, A.tree_first[sr.first_tree + t], nf);
    if (sum < (double)(sr.threshold - 0.00001F)) return false;
  }
  return true;
}

__device__ __forceinline__ bool lvl_eval_window(const LevelDesc& L, long long k, const CascadeArgs& A) {
  const int iy = (int)(k / L.nx), ix = (int)(k - (long long)iy * L.nx);
  const size_t org = (size_t)L.ip_off + (size_t)(L.y0 + iy * L.step_y) * L.pitch + (L.x0 + ix * L.step_x);
  return eval_window_at(A.pl.sum + org, A.pl.sq ? A.pl.sq + org : nullptr, A.pl.tilted ? A.pl.tilted + org : nullptr, L.pitch, A);
}

// ordered-in-warp compaction of the survivors (warp ballot, one atomic per warp); the host sorts the short final list
__device__ __forceinline__ void lvl_append(bool pass, unsigned int b, unsigned int* __restrict__ out_count, unsigned int* __restrict__ out_list) {
  const unsigned m = __ballot_sync(0xffffffffu, pass);
  if (m) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(out_count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pass) out_list[base + __popc(m & ((1u << lane) - 1u))] = b;
  }
}

// first stage group: every window of the batch.  One CTA per level strip (grid-stride), threads over its windows.
__global__ void __launch_bounds__(256)
k_lvl_cascade_all(const LevelDesc* __restrict__ levels, int nlevels, CascadeArgs A, unsigned int* __restrict__ out_count,
                  unsigned int* __restrict__ out_list) {
  for (int lv = blockIdx.x; lv < nlevels; lv += gridDim.x) {
    const LevelDesc& L = levels[lv];
    if (!L.owned) continue;              // another rank evaluates this level (image-sharded mining)
    const int n = (int)L.w_count;
    const int nround = (n + 31) & ~31;
    for (int i = threadIdx.x; i < nround; i += blockDim.x) {
      const bool pass = i < n && lvl_eval_window(L, L.w_skip + i, A);
      lvl_append(pass, (unsigned int)(L.w_base + i), out_count, out_list);
    }
  }
}

// later stage groups: the survivors of the previous group (count read on the device)
__global__ void __launch_bounds__(256)
k_lvl_cascade(const LevelDesc* __restrict__ levels, int nlevels, CascadeArgs A, const unsigned int* __restrict__ in_list,
              const unsigned int* __restrict__ in_count, unsigned int* __restrict__ out_count, unsigned int* __restrict__ out_list) {
  const long long n = (long long)*in_count;
  const long long nround = (n + 31) & ~31ll;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nround; i += (long long)gridDim.x * blockDim.x) {
    bool pass = i < n;
    unsigned int b = 0;
    if (pass) {
      b = in_list[i];
      int lo = 0, hi = nlevels - 1;
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (levels[mid].w_base <= (long long)b) lo = mid; else hi = mid - 1; }
      const LevelDesc& L = levels[lo];
      pass = lvl_eval_window(L, (long long)b - L.w_base + L.w_skip, A);
    }
    lvl_append(pass, b, out_count, out_list);
  }
}

// Deep stage groups: few survivors, many trees per stage.  One thread per window would walk ~150 trees one after the other,
// a chain of dependent loads of ~1 us each that no amount of parallelism across the handful of surviving windows hides.  Here a
// WARP takes a window: the lanes evaluate the trees of a stage in parallel, the stage sum is then accumulated in tree order
// (lane k's leaf value broadcast to all lanes, k = 0, 1, ...), i.e. with the additions of CvCascadeBoost::predict
// (src/boost.cpp:461-477) in the reference's order -- same f64 sum, same decision.
__global__ void __launch_bounds__(256)
k_lvl_cascade_warp(const LevelDesc* __restrict__ levels, int nlevels, CascadeArgs A, const unsigned int* __restrict__ in_list,
                   const unsigned int* __restrict__ in_count, unsigned int* __restrict__ out_count, unsigned int* __restrict__ out_list) {
  const long long n = (long long)*in_count;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
    const unsigned int b = in_list[i];
    int lo = 0, hi = nlevels - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (levels[mid].w_base <= (long long)b) lo = mid; else hi = mid - 1; }
    const LevelDesc& L = levels[lo];
    const long long k = (long long)b - L.w_base + L.w_skip;
    const int iy = (int)(k / L.nx), ix = (int)(k - (long long)iy * L.nx);
    const size_t org = (size_t)L.ip_off + (size_t)(L.y0 + iy * L.step_y) * L.pitch + (L.x0 + ix * L.step_x);
    const uint32_t* S = A.pl.sum + org;
    const uint32_t* Q = A.pl.sq ? A.pl.sq + org : nullptr;
    const uint32_t* Tt = A.pl.tilted ? A.pl.tilted + org : nullptr;
    const float nf = window_nf(S, Q, L.pitch, A);
    bool pass = true;
    for (int st = A.s_begin; st < A.s_end && pass; ++st) {
      const DevStageRef sr = A.stages[st];
      double sum = 0;
      for (int t0 = 0; t0 < sr.num_trees; t0 += 32) {
        const int t = t0 + lane;
        const double val = t < sr.num_trees ? window_tree_value(S, Tt, L.pitch, A, A.tree_first[sr.first_tree + t], nf) : 0.0;
        const int cnt = min(32, sr.num_trees - t0);
        for (int q = 0; q < cnt; ++q) sum += __shfl_sync(0xffffffffu, val, q);
      }
      if (sum < (double)(sr.threshold - 0.00001F)) pass = false;
    }
    if (pass && lane == 0) out_list[atomicAdd(out_count, 1u)] = b;
  }
}

// ------------------------------------------------------------------------------------------------
// Fused tile miner (cc_mine_set_mode 3, Haar / LBP): one CTA takes a tile of TW x TH windows of one pyramid level, builds the tile's
// pixels (INTER_LINEAR_EXACT from the source image) and its sum / sqsum / tilted integral planes in shared memory, runs
// the whole cascade on the tile's windows and appends the survivors.  Window sums are differences, so tile-local planes
// (zero outside the tile) give the same integers as the per-crop cv::integral of setImage; nothing but the survivor
// ordinals leaves the SM.  No phase has a barrier inside a loop: rows are scanned by warps (independent rows), columns
// and diagonals are walked by one thread each.
// ------------------------------------------------------------------------------------------------
struct TileDesc { int level; int tx, ty; };

__global__ void __launch_bounds__(256)
k_tile_cascade(const LevelDesc* __restrict__ levels, const TileDesc* __restrict__ tiles, int TW, int TH, int tpw, int tph, int pp,
               int want_sq, int want_tilted, CascadeArgs A, unsigned int* __restrict__ out_count, unsigned int* __restrict__ out_list) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int plane_words = pp * (tph + 1);
  uint32_t* S = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* Q = S + plane_words;
  uint32_t* T = Q + (want_sq ? plane_words : 0);
  uint8_t* pix = reinterpret_cast<uint8_t*>(T + (want_tilted ? plane_words : 0));      // [tph][tpw]
  int* xcoef = reinterpret_cast<int*>(pix + ((tpw * tph + 3) & ~3));                    // [tpw] source column | c1 << 16
  int* ycoef = xcoef + tpw;                                                             // [tph]
  const TileDesc td = tiles[blockIdx.x];
  const LevelDesc& L = levels[td.level];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int W = A.W, H = A.H, sx = L.step_x, sy = L.step_y;
  const int ix0 = td.tx * TW, iy0 = td.ty * TH;
  const int nwx = min(TW, L.nx - ix0), nwy = min(TH, L.ny - iy0);
  const int aw = (nwx - 1) * sx + W, ah = (nwy - 1) * sy + H;                 // pixels of this tile
  const int lx0 = L.x0 + ix0 * sx, ly0 = L.y0 + iy0 * sy;                     // tile origin in level coordinates
  const int sw = L.sw, sh = L.sh, spitch = L.spitch;
  const uint8_t* __restrict__ src = L.src;
  // ---- phase 0: resize coefficients, zero borders of the planes ----
  for (int i = t; i < aw + ah; i += 256) {
    int o, c0, c1, md;
    if (i < aw) { lin_coef(L.scale_x, lx0 + i, sw, o, c0, c1, md); xcoef[i] = o | (c1 << 16); }
    else { lin_coef(L.scale_y, ly0 + (i - aw), sh, o, c0, c1, md); ycoef[i - aw] = o | (c1 << 16); }
  }
  for (int i = t; i <= aw; i += 256) { S[i] = 0; if (want_sq) Q[i] = 0; if (want_tilted) T[i] = 0; }
  for (int i = t; i <= ah; i += 256) { S[i * pp] = 0; if (want_sq) Q[i * pp] = 0; if (want_tilted) T[i * pp] = 0; }
  __syncthreads();
  // ---- phase 1: pixels + row prefixes (one warp per row) ----
  for (int y = warp; y < ah; y += 8) {
    const int yo = ycoef[y];
    const int r0 = yo & 0xFFFF, r1 = min(r0 + 1, sh - 1);
    const uint32_t cy1 = (uint32_t)yo >> 16, cy0 = 256u - cy1;
    const uint8_t* a = src + (size_t)r0 * spitch;
    const uint8_t* b = src + (size_t)r1 * spitch;
    uint32_t c1s = 0, c2s = 0;
    for (int xb = 0; xb < aw; xb += 32) {
      const int x = xb + lane;
      uint32_t v = 0;
      if (x < aw) {
        const int xo = xcoef[x];
        const int xa = xo & 0xFFFF, xn = min(xa + 1, sw - 1);
        const uint32_t cx1 = (uint32_t)xo >> 16, cx0 = 256u - cx1;
        v = (((uint32_t)a[xa] * cx0 + (uint32_t)a[xn] * cx1) * cy0 + ((uint32_t)b[xa] * cx0 + (uint32_t)b[xn] * cx1) * cy1 + 32768u) >> 16;
        pix[y * tpw + x] = (uint8_t)v;
      }
      uint32_t p1 = v, p2 = v * v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u1 = __shfl_up_sync(0xffffffffu, p1, d), u2 = __shfl_up_sync(0xffffffffu, p2, d);
        if (lane >= d) { p1 += u1; p2 += u2; }
      }
      p1 += c1s; p2 += c2s;
      if (x < aw) { S[(y + 1) * pp + x + 1] = p1; if (want_sq) Q[(y + 1) * pp + x + 1] = p2; }
      c1s = __shfl_sync(0xffffffffu, p1, 31); c2s = __shfl_sync(0xffffffffu, p2, 31);
    }
  }
  __syncthreads();
  // ---- phase 2: column prefixes of S (and Q); tilted step 1: D1 along the down-right diagonals ----
  {
    const int ncol = want_sq ? 2 * aw : aw;
    if (t < ncol) {
      uint32_t* c = (t < aw ? S : Q) + (t < aw ? t : t - aw) + 1;
      uint32_t acc = 0;
      for (int y = 1; y <= ah; ++y) { acc += c[y * pp]; c[y * pp] = acc; }
    }
  }
  if (want_tilted) {
    // cells T[y+1][x+1] (x = -1 .. aw-1): E(y,x) = D1(y,x) + D2(y,x) - I(y,x), then a column prefix gives the tilted integral
    for (int k = t; k < aw + ah - 1; k += 256) {
      int y = k < aw ? 0 : k - aw + 1, x = k < aw ? k : 0;
      uint32_t run = 0;
      for (; y < ah && x < aw; ++y, ++x) { run += pix[y * tpw + x]; T[(y + 1) * pp + x + 1] = run; }
    }
    __syncthreads();
    for (int k = t; k < aw + ah - 1; k += 256) {
      int y = k < aw ? 0 : k - aw + 1, x = k < aw ? k : aw - 1;
      uint32_t run = 0;
      for (; y < ah && x >= 0; ++y, --x) { const uint32_t v = pix[y * tpw + x]; run += v; T[(y + 1) * pp + x + 1] += run - v; }
      if (y < ah) T[(y + 1) * pp] += run;                // virtual column x = -1
    }
    __syncthreads();
    if (t <= aw) {
      uint32_t* c = T + t;
      uint32_t acc = 0;
      for (int y = 1; y <= ah; ++y) { acc += c[y * pp]; c[y * pp] = acc; }
    }
  }
  __syncthreads();
  // ---- phase 3: the cascade on the tile's windows ----
  const int nw = nwx * nwy;
  for (int i0 = 0; i0 < nw; i0 += 256) {
    const int i = i0 + t;
    bool pass = false;
    unsigned int b = 0;
    if (i < nw) {
      const int wy = i / nwx, wx = i - wy * nwx;
      const long long k = (long long)(iy0 + wy) * L.nx + (ix0 + wx);
      if (k >= L.w_skip && k < L.w_skip + L.w_count) {
        const int org = (wy * sy) * pp + wx * sx;
        pass = eval_window_at(S + org, Q + org, T + org, pp, A);
        b = (unsigned int)(L.w_base + (k - L.w_skip));
      }
    }
    lvl_append(pass, b, out_count, out_list);
  }
}

// accepted windows (batch ordinals) recomputed from the source images: INTER_LINEAR_EXACT pixel by pixel (the tile miner
// keeps no level pixels).  Ordinals outside this rank's levels give zeros (image-sharded mining sums the ranks' outputs).
__global__ void k_gather_resize(const LevelDesc* __restrict__ levels, int nlevels, int W, int H, const unsigned int* __restrict__ which, int n,
                                uint8_t* __restrict__ dst, uint8_t* __restrict__ labels, int label) {
  const int npx = W * H;
  const long long total = (long long)n * npx;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / npx);
    const int r = (int)(t - (long long)i * npx);
    if (r == 0) labels[i] = (uint8_t)label;
    if (nlevels == 0) { dst[t] = 0; continue; }
    const long long b = which[i];
    int lo = 0, hi = nlevels - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (levels[mid].w_base <= b) lo = mid; else hi = mid - 1; }
    const LevelDesc& L = levels[lo];
    if (b < L.w_base || b >= L.w_base + L.w_count) { dst[t] = 0; continue; }
    const long long k = b - L.w_base + L.w_skip;
    const int iy = (int)(k / L.nx), ix = (int)(k - (long long)iy * L.nx);
    const int wy = r / W, wx = r - wy * W;
    const int x = L.x0 + ix * L.step_x + wx, y = L.row0 + L.y0 + iy * L.step_y + wy;
    int ox, cx0, cx1, mx, oy, cy0, cy1, my;
    lin_coef(L.scale_x, x, L.sw, ox, cx0, cx1, mx);
    lin_coef(L.scale_y, y, L.sh, oy, cy0, cy1, my);
    const uint8_t* a = L.src + (size_t)oy * L.spitch;
    const uint8_t* bb = L.src + (size_t)min(oy + 1, L.sh - 1) * L.spitch;
    const int xn = min(ox + 1, L.sw - 1);
    const uint32_t val = ((uint32_t)a[ox] * cx0 + (uint32_t)a[xn] * cx1) * (uint32_t)cy0 + ((uint32_t)bb[ox] * cx0 + (uint32_t)bb[xn] * cx1) * (uint32_t)cy1;
    dst[t] = (uint8_t)((val + 32768u) >> 16);
  }
}

// survivor list -> bitmap over the batch ordinals (image-sharded mining: the bitmaps of the ranks are disjoint, so
// their element-wise sum is their union)
__global__ void k_list_to_bitmap(const unsigned int* __restrict__ list, const unsigned int* __restrict__ count, unsigned int* __restrict__ bitmap) {
  const unsigned int n = *count;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned int b = list[i];
    atomicOr(bitmap + (b >> 5), 1u << (b & 31));
  }
}

// ---- host-side state ----------------------------------------------------------------------------------
struct BgImage { int w = 0, h = 0, pitch = 0; uint8_t* d_px = nullptr; };      // pitch: multiple of 16 bytes; w == 0: placeholder of an unreadable file

// Survivor cache of the window stream.  NegReader cycles the background list with a `round` phase
// (src/imagestorage.cpp:57-88): image i is scanned with the offsets (round % W, round / W) clamped to the image, so the
// stream revisits the pyramid (image, offsets) once per W*H list cycles -- in late stages many times per stage when
// the list is short.  Cascade stages are append-only, so the windows of a pyramid that pass stages [0, n) are a superset of those
// that pass any longer cascade: the cache keeps, per pyramid, the ascending stream ordinals (relative to the pyramid's first
// window) of the windows that passed the first n_stages stages.  A revisit evaluates only stages [n_stages, ns) on them
// (k_gather_resize + K1 + k_cascade_flags: the literal per-window setImage + predict path) instead of building the pyramid's
// integral planes again.  Which windows are CONSUMED, the ordered cut, `consumed`, the reader state and the mined pixels are
// those of the uncached evaluation (test_survivor_cache_is_transparent).  Entries are only created from pyramids that
// were evaluated from their first to their last window and have few survivors; the cache is dropped when the background list,
// the window size or an already cached cascade prefix changes.
struct PyrEntry { int n_stages = 0; long long nwin = 0; std::vector<uint32_t> surv; };
struct PyrCache {
  std::unordered_map<uint64_t, PyrEntry> map;
  std::vector<PyrEntry*> table;        // [image][offY][offX] -> entry (nullptr: none); rebuilt lazily, see lookup()
  int table_imgs = 0;
  bool table_valid = false;
  std::vector<uint64_t> stage_hash;    // hash of every stage of the cascade the entries were computed with
  size_t bytes = 0, max_bytes = (size_t)4 << 30;
  bool enabled = true;
  int W = 0, H = 0, ftype = -1, haar_mode = -1;
  void clear() { map.clear(); bytes = 0; table_valid = false; }
  // assemble() looks a pyramid up about a million times per late stage: a flat table instead of the hash map (entries are
  // stable in an unordered_map until erased; every insert / erase invalidates the table)
  const PyrEntry* lookup(int src, int offX, int offY, int nimg) {
    if (!table_valid) {
      table.assign((size_t)nimg * W * H, nullptr);
      table_imgs = nimg;
      for (auto& kv : map) {
        const int s = (int)(kv.first >> 32), oy = (int)((kv.first >> 16) & 0xFFFF), ox = (int)(kv.first & 0xFFFF);
        if (s < nimg && oy < H && ox < W) table[((size_t)s * H + oy) * W + ox] = &kv.second;
      }
      table_valid = true;
    }
    return table[((size_t)src * H + offY) * W + offX];
  }
};

struct MineState {
  std::vector<BgImage> bgs;
  // cascade
  std::vector<DevNode> nodes;
  std::vector<int> tree_first;
  std::vector<DevStageRef> stages;
  DevNode* d_nodes = nullptr;
  int* d_tree_first = nullptr;
  DevStageRef* d_stages = nullptr;
  // scratch
  EvalStore scratch;
  int scratch_cap = 0;
  uint8_t* d_pool = nullptr;
  size_t pool_cap = 0;
  LevelDesc* d_levels = nullptr;
  int levels_cap = 0;
  unsigned int* d_surv = nullptr;      // [0] = counter, [1..cap] = list
  unsigned int surv_cap = 0;
  unsigned int* h_surv = nullptr;      // pinned mirror (count + list) shared by the slot path and the tile miner
  size_t h_surv_cap = 0;               // entries
  float2* d_hog_lut = nullptr;         // HOG miner: (magnitude, bin) of every (dx, dy) pair, see k_hog_lut_init
  int* d_hog_blk = nullptr;            // HOG miner: level of every HOGW_BLK-th window of the batch
  size_t hog_blk_cap = 0;
  // whole-level fast path (Haar / LBP)
  std::vector<LvlFeat> lfeats;         // parallel to `nodes`
  LvlFeat* d_lfeats = nullptr;
  bool cascade_tilted = false;         // some node uses a tilted feature
  bool force_unfused = false;          // cc_mine_set_mode 1: build the planes with the multi-kernel path (k_lvl_rows/diag/cols)
  bool force_slots = false;            // cc_mine_set_mode 2: per-window scratch slots (the literal setImage + predict path)
  uint32_t *d_psum = nullptr, *d_psq = nullptr, *d_ptilt = nullptr;
  size_t plane_cap = 0;
  unsigned int *d_list[2] = {nullptr, nullptr}, *d_counts = nullptr;
  size_t list_cap = 0;
  int* d_order = nullptr;
  int order_cap = 0;
  unsigned int* d_bitmap = nullptr;    // image-sharded mining: survivor bitmap, summed over the ranks
  unsigned int* h_bitmap = nullptr;
  size_t bitmap_cap = 0;
  bool shard_images = true;            // when an NCCL communicator is attached
  std::vector<LevelDesc> dev_levels;   // fast path: the batch's levels (tile miner) or their strips (plane path), as uploaded to d_levels
  std::vector<TileDesc> tiles;
  TileDesc* d_tiles = nullptr;
  int tiles_cap = 0;
  int mode = 0;                        // cc_mine_set_mode
  // plane path, two batches in flight: strips of batch k stay resident for its gather while batch k+1 is evaluated
  LevelDesc* d_strips[2] = {nullptr, nullptr};
  int strips_cap[2] = {0, 0};
  LevelDesc* h_strips[2] = {nullptr, nullptr};       // pinned staging: the upload must not wait for the batch in flight
  std::vector<LevelDesc> strips[2];
  unsigned int* h_bits[2] = {nullptr, nullptr};      // pinned survivor bitmaps
  size_t h_bits_cap[2] = {0, 0};
  cudaEvent_t ev_done[2] = {nullptr, nullptr};
  unsigned int* d_which = nullptr;     // accepted ordinals of a batch
  int which_cap = 0;
  // survivor cache + the re-evaluation of cached survivors (candidates)
  PyrCache cache;
  EvalStore cand_store;                // scratch slots of the candidates of one batch
  int cand_cap = 0;
  uint8_t* d_cand = nullptr;           // device: [LevelDesc table | which u32 | first_stage u8 | flags u8] of one batch
  size_t d_cand_cap = 0;
  uint8_t* h_cand[2] = {nullptr, nullptr};            // pinned staging of the upload, one per batch in flight
  size_t h_cand_cap[2] = {0, 0};
  uint8_t* h_cand_flags[2] = {nullptr, nullptr};      // pinned result flags
  size_t h_cand_flags_cap[2] = {0, 0};
  uint8_t* d_gather = nullptr;         // device: [LevelDesc table | which u32] for the accepted windows of a batch
  size_t d_gather_cap = 0;
  uint8_t* h_gather = nullptr;         // pinned staging of that upload; ev_gather: the last upload has left it
  size_t h_gather_cap = 0;
  cudaEvent_t ev_gather = nullptr;
  long long stat_consumed = 0, stat_built = 0, stat_cand = 0, stat_skipped = 0;
  // the three width classes of the pyramid builder run side by side (their tails overlap): two helper streams forked from / joined to ctx->stream
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
};

static void free_cascade(MineState* m) {
  cudaFree(m->d_nodes); cudaFree(m->d_tree_first); cudaFree(m->d_stages); cudaFree(m->d_lfeats);
  m->d_nodes = nullptr; m->d_tree_first = nullptr; m->d_stages = nullptr; m->d_lfeats = nullptr;
  m->nodes.clear(); m->tree_first.clear(); m->stages.clear(); m->lfeats.clear();
  m->cascade_tilted = false;
}
static void free_scratch(MineState* m) {
  store_free(m->scratch);
  m->scratch_cap = 0;
  cudaFree(m->d_pool); m->d_pool = nullptr; m->pool_cap = 0;
  cudaFree(m->d_levels); m->d_levels = nullptr; m->levels_cap = 0;
  cudaFree(m->d_surv); m->d_surv = nullptr; m->surv_cap = 0;
  cudaFreeHost(m->h_surv); m->h_surv = nullptr; m->h_surv_cap = 0;
  cudaFree(m->d_hog_lut); m->d_hog_lut = nullptr;
  cudaFree(m->d_hog_blk); m->d_hog_blk = nullptr; m->hog_blk_cap = 0;
  cudaFree(m->d_psum); cudaFree(m->d_psq); cudaFree(m->d_ptilt);
  m->d_psum = m->d_psq = m->d_ptilt = nullptr; m->plane_cap = 0;
  cudaFree(m->d_list[0]); cudaFree(m->d_list[1]); cudaFree(m->d_counts); cudaFree(m->d_order);
  m->d_list[0] = m->d_list[1] = m->d_counts = nullptr; m->list_cap = 0; m->d_order = nullptr; m->order_cap = 0;
  cudaFree(m->d_bitmap); cudaFreeHost(m->h_bitmap); m->d_bitmap = nullptr; m->h_bitmap = nullptr; m->bitmap_cap = 0;
  cudaFree(m->d_tiles); m->d_tiles = nullptr; m->tiles_cap = 0;
  for (int s = 0; s < 2; ++s) {
    cudaFree(m->d_strips[s]); m->d_strips[s] = nullptr; m->strips_cap[s] = 0;
    cudaFreeHost(m->h_strips[s]); m->h_strips[s] = nullptr;
    cudaFreeHost(m->h_bits[s]); m->h_bits[s] = nullptr; m->h_bits_cap[s] = 0;
    if (m->ev_done[s]) { cudaEventDestroy(m->ev_done[s]); m->ev_done[s] = nullptr; }
  }
  cudaFree(m->d_which); m->d_which = nullptr; m->which_cap = 0;
  store_free(m->cand_store); m->cand_cap = 0;
  cudaFree(m->d_cand); m->d_cand = nullptr; m->d_cand_cap = 0;
  cudaFree(m->d_gather); m->d_gather = nullptr; m->d_gather_cap = 0;
  cudaFreeHost(m->h_gather); m->h_gather = nullptr; m->h_gather_cap = 0;
  for (int s = 0; s < 2; ++s) {
    if (m->side[s]) { cudaStreamDestroy(m->side[s]); m->side[s] = nullptr; }
    if (m->ev_join[s]) { cudaEventDestroy(m->ev_join[s]); m->ev_join[s] = nullptr; }
  }
  if (m->ev_fork) { cudaEventDestroy(m->ev_fork); m->ev_fork = nullptr; }
  if (m->ev_gather) { cudaEventDestroy(m->ev_gather); m->ev_gather = nullptr; }
  for (int s = 0; s < 2; ++s) {
    cudaFreeHost(m->h_cand[s]); m->h_cand[s] = nullptr; m->h_cand_cap[s] = 0;
    cudaFreeHost(m->h_cand_flags[s]); m->h_cand_flags[s] = nullptr; m->h_cand_flags_cap[s] = 0;
  }
}

int mine_free(cc_ctx* ctx) {
  MineState* m = ctx->mine;
  if (!m) return CC_OK;
  for (auto& b : m->bgs) cudaFree(b.d_px);
  free_cascade(m);
  free_scratch(m);
  delete m;
  ctx->mine = nullptr;
  return CC_OK;
}

static MineState* mine_state(cc_ctx* ctx) {
  if (!ctx->mine) ctx->mine = new MineState();
  return ctx->mine;
}

static int ensure_h_surv(MineState* m, size_t entries) {
  if (entries <= m->h_surv_cap) return CC_OK;
  cudaFreeHost(m->h_surv); m->h_surv = nullptr; m->h_surv_cap = 0;
  CC_CUDA(cudaMallocHost(&m->h_surv, sizeof(unsigned) * entries));
  m->h_surv_cap = entries;
  return CC_OK;
}

// `grow`: how much larger the largest batch of this call can be than this one (HOG window miner: batches grow with the falling
// acceptance rate up to the batch limit) -- the buffers are sized for it at once, a cudaFree + cudaMalloc of gigabytes per growth step
// costs more than a batch
static int ensure_scratch(cc_ctx* ctx, MineState* m, int nwin, size_t pool_px, int nlevels, long long surv_n = -1, double grow = 1.0) {
  if (surv_n < 0) surv_n = nwin;
  grow = std::max(1.0, grow);
  if (nwin > m->scratch_cap) {
    store_free(m->scratch);
    CC_CHECK(store_alloc(ctx, m->scratch, nwin));
    m->scratch_cap = nwin;
  }
  // (capacities grow geometrically: batches grow with the falling acceptance rate, and a cudaFree + cudaMalloc of gigabytes per batch
  // costs more than the batch)
  if (pool_px > m->pool_cap) {
    cudaFree(m->d_pool); m->d_pool = nullptr; m->pool_cap = 0;
    size_t cap = (size_t)std::min((double)pool_px * 1.25 * grow, std::max((double)pool_px * 1.5, 6.0e9));
    if (cudaMalloc(&m->d_pool, cap + 64) != cudaSuccess) {      // slack: the HOG miner copies the aligned words around a row
      (void)cudaGetLastError();
      cap = pool_px + pool_px / 2;
      CC_CUDA(cudaMalloc(&m->d_pool, cap + 64));
    }
    m->pool_cap = cap;
  }
  if (nlevels > m->levels_cap) {
    cudaFree(m->d_levels); m->d_levels = nullptr; m->levels_cap = 0;
    const int cap = (int)std::min((double)nlevels * 1.25 * grow + 16, std::max((double)nlevels * 1.5 + 16, 4.0e6));
    CC_CUDA(cudaMalloc(&m->d_levels, sizeof(LevelDesc) * cap));
    m->levels_cap = cap;
  }
  if ((unsigned long long)surv_n > m->surv_cap) {
    cudaFree(m->d_surv); m->d_surv = nullptr; m->surv_cap = 0;
    const size_t cap = (size_t)std::min<double>(std::min((double)surv_n * 1.25 * grow, std::max((double)surv_n * 1.5, 4.0e7)) + 1024, 4294967280.0);
    CC_CUDA(cudaMalloc(&m->d_surv, sizeof(unsigned) * (cap + 1)));
    m->surv_cap = (unsigned)cap;
  }
  return ensure_h_surv(m, (size_t)m->surv_cap + 1);
}

// ---- NegReader state machine without the pixels (src/imagestorage.cpp:57-126) -------------------------
struct NegStream {
  const std::vector<BgImage>* bgs;
  int W, H;
  long long last = 0, round = 0;
  int src = -1;                 // index of the current source image
  int pointX = 0, pointY = 0, offX = 0, offY = 0;
  int lw = 0, lh = 0;           // current level size (0: no image loaded yet)
  float scale = 1.0f;
  bool fresh = false;           // the last transition loaded a new image (not part of the reader state)
  static constexpr float scaleFactor = 1.4142135623730950488016887242097F;
  static constexpr float stepFactor = 0.5F;

  bool nextImg() {
    const long long count = (long long)bgs->size();
    int ox = 0, oy = 0;
    bool found = false;
    for (long long i = 0; i < count; ++i) {
      src = (int)last;
      last++;
      const BgImage& im = (*bgs)[src];
      if (im.w == 0 || im.h == 0) { last %= count; continue; }      // unreadable file: imread() gave an empty Mat (src/imagestorage.cpp:64-67) -- `round` is not advanced
      round += last / count;
      round = round % ((long long)W * H);
      last %= count;
      ox = std::min((int)round % W, im.w - W);
      oy = std::min((int)round / W, im.h - H);
      if (ox >= 0 && oy >= 0) { found = true; break; }
    }
    if (!found) return false;    // the reference would read out of bounds here; refuse instead
    pointX = offX = ox; pointY = offY = oy;
    const BgImage& im = (*bgs)[src];
    scale = std::max(((float)W + pointX) / ((float)im.w), ((float)H + pointY) / ((float)im.h));
    lw = (int)(scale * im.w + 0.5F);
    lh = (int)(scale * im.h + 0.5F);
    fresh = true;
    return true;
  }
  // the stream stands on the first window of the first level of (image src, offsets): everything the reader does until
  // the next nextImg() is a function of (src, offX, offY) alone
  bool is_pyramid_start() const {
    if (src < 0 || lw == 0 || pointX != offX || pointY != offY) return false;
    const BgImage& im = (*bgs)[src];
    if (im.w == 0) return false;
    const float s0 = std::max(((float)W + offX) / ((float)im.w), ((float)H + offY) / ((float)im.h));
    return scale == s0 && lw == (int)(s0 * im.w + 0.5F) && lh == (int)(s0 * im.h + 0.5F);
  }
  uint64_t pyramid_key() const { return ((uint64_t)(uint32_t)src << 32) | ((uint64_t)(uint32_t)offY << 16) | (uint64_t)(uint32_t)offX; }
  int stepX() const { return (int)(stepFactor * W); }
  int stepY() const { return (int)(stepFactor * H); }
  bool canStepX(int x) const { return (int)(x + (1.0F + stepFactor) * W) < lw; }
  bool canStepY(int y) const { return (int)(y + (1.0F + stepFactor) * H) < lh; }
  // windows per level row / column: 1 + the number of steps NegReader::get takes before canStep fails.  (int)(x + 1.5f * W) ==
  // x + (int)(1.5f * W) for the non-negative integers x that occur (1.5f * W is a multiple of 0.5, everything is exact in f32),
  // so the loop `while (canStepX(x)) x += stepX()` has a closed form; steps of 0 pixels (W == 1) never advance in the reference
  // either and are treated as a single position.
  static int count1(int off, int step, int reach, int len) {
    if (step <= 0) return 1;
    const int room = len - reach - off - 1;          // largest x - off (x = position before a step) that still may step
    return room < 0 ? 1 : 2 + room / step;
  }
  int countX() const { return count1(offX, stepX(), (int)((1.0F + stepFactor) * W), lw); }
  int countY() const { return count1(offY, stepY(), (int)((1.0F + stepFactor) * H), lh); }
  // after the last window of the current level: next level of the same image or the next image
  bool nextLevel() {
    pointX = offX; pointY = offY;
    scale *= scaleFactor;
    if (scale <= 1.0F) {
      const BgImage& im = (*bgs)[src];
      lw = (int)(scale * im.w);
      lh = (int)(scale * im.h);
      fresh = false;
      return true;
    }
    return nextImg();
  }
};

static void stream_from_state(NegStream& ns, const cc_reader_state& st, int nimg) {
  ns.last = st.last; ns.round = st.round;
  ns.pointX = st.point_x; ns.pointY = st.point_y; ns.offX = st.offset_x; ns.offY = st.offset_y;
  ns.lw = st.level_w; ns.lh = st.level_h; ns.scale = st.scale;
  ns.src = nimg > 0 ? (int)((st.last + nimg - 1) % nimg) : -1;
}
static void state_from_stream(const NegStream& ns, cc_reader_state* st) {
  st->last = ns.last; st->round = ns.round;
  st->point_x = ns.pointX; st->point_y = ns.pointY; st->offset_x = ns.offX; st->offset_y = ns.offY;
  st->level_w = ns.lw; st->level_h = ns.lh; st->scale = ns.scale; st->pad = 0;
}

static void fill_level(const NegStream& ns, LevelDesc& L) {
  const BgImage& im = (*ns.bgs)[ns.src];
  L.src = im.d_px; L.sw = im.w; L.sh = im.h; L.lw = ns.lw; L.lh = ns.lh;
  L.scale_x = 1.0 / ((double)ns.lw / (double)im.w);
  L.scale_y = 1.0 / ((double)ns.lh / (double)im.h);
  L.x0 = ns.offX; L.y0 = ns.offY; L.step_x = ns.stepX(); L.step_y = ns.stepY();
  L.nx = ns.countX(); L.ny = ns.countY();
  L.row0 = 0; L.spitch = im.pitch; L.owned = 1; L.pitch = L.lw + 1; L.ppitch = L.lw;
}

static uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001B3ull; }
  return h;
}
static uint64_t stage_hash(const cc_stage& st) {
  uint64_t h = 0xCBF29CE484222325ull;
  h = fnv1a(h, &st.threshold, sizeof(float));
  h = fnv1a(h, &st.num_weak, sizeof(int32_t));
  for (int t = 0; t < st.num_weak; ++t) {
    const cc_tree& tr = st.trees[t];
    h = fnv1a(h, &tr.num_nodes, sizeof(int32_t));
    for (int i = 0; i < tr.num_nodes; ++i) {
      const cc_node& nd = tr.nodes[i];
      h = fnv1a(h, &nd.left, 4); h = fnv1a(h, &nd.right, 4); h = fnv1a(h, &nd.var_idx, 4); h = fnv1a(h, &nd.c, 4);
      h = fnv1a(h, nd.subset, sizeof(nd.subset)); h = fnv1a(h, &nd.value, 8);
    }
  }
  return h;
}
// the cache entries hold survivors of a cascade PREFIX: they stay valid as long as that prefix is unchanged
static void cache_validate(cc_ctx* ctx, MineState* m, const cc_stage* stages, int n) {
  PyrCache& c = m->cache;
  if (c.W != ctx->W || c.H != ctx->H || c.ftype != ctx->ftype || c.haar_mode != ctx->haar_mode) {
    c.clear(); c.stage_hash.clear();
    c.W = ctx->W; c.H = ctx->H; c.ftype = ctx->ftype; c.haar_mode = ctx->haar_mode;
  }
  std::vector<uint64_t> hs(n);
  for (int i = 0; i < n; ++i) hs[i] = stage_hash(stages[i]);
  size_t common = 0;
  while (common < hs.size() && common < c.stage_hash.size() && hs[common] == c.stage_hash[common]) ++common;
  if (common < c.stage_hash.size()) {          // a stage the entries may depend on changed or disappeared
    for (auto it = c.map.begin(); it != c.map.end();) {
      if ((size_t)it->second.n_stages > common) { c.bytes -= it->second.surv.size() * 4 + 64; it = c.map.erase(it); } else ++it;
    }
    c.table_valid = false;
  }
  c.stage_hash.swap(hs);
}

// upload the cascade as flat arrays
static int upload_cascade(cc_ctx* ctx, MineState* m, const cc_stage* stages, int n) {
  free_cascade(m);
  cache_validate(ctx, m, stages, n);
  for (int s = 0; s < n; ++s) {
    DevStageRef sr;
    sr.first_tree = (int)m->tree_first.size();
    sr.num_trees = stages[s].num_weak;
    sr.threshold = stages[s].threshold;
    for (int t = 0; t < stages[s].num_weak; ++t) {
      m->tree_first.push_back((int)m->nodes.size());
      const cc_tree& tr = stages[s].trees[t];
      for (int i = 0; i < tr.num_nodes; ++i) {
        DevNode d;
        d.left = tr.nodes[i].left; d.right = tr.nodes[i].right; d.var_idx = tr.nodes[i].var_idx; d.c = tr.nodes[i].c;
        std::memcpy(d.subset, tr.nodes[i].subset, sizeof(d.subset));
        d.value = tr.nodes[i].value;
        m->nodes.push_back(d);
        LvlFeat lf;
        std::memset(&lf, 0, sizeof(lf));
        if (d.left >= 0 && ctx->ftype != CC_FEATURE_HOG) {
          CC_REQUIRE(d.var_idx >= 0 && d.var_idx < ctx->num_features, CC_ERR_BAD_ARG, "cc_cascade_set_stages: feature index %d out of range", d.var_idx);
          const int S = ctx->W + 1;
          if (ctx->ftype == CC_FEATURE_HAAR) {
            const HaarTaps t = make_haar_taps(ctx->haar[d.var_idx], S);
            const int nt = t.ntaps();
            lf.hdr = (uint32_t)nt | ((uint32_t)t.tilted() << 8);
            for (int q = 0; q < nt; ++q) {
              const int off = tap_off(t.w[q]);
              lf.tap[q] = (uint32_t)(off % S) | ((uint32_t)(off / S) << 8) | ((uint32_t)(uint8_t)(int8_t)tap_coef(t.w[q]) << 16);
            }
            if (t.tilted()) m->cascade_tilted = true;
          } else {
            const LbpTaps t = make_lbp_taps(ctx->lbp[d.var_idx], S);
            lf.hdr = 16;
            for (int q = 0; q < 16; ++q) lf.tap[q] = (uint32_t)(t.off[q] % S) | ((uint32_t)(t.off[q] / S) << 8);
          }
        }
        m->lfeats.push_back(lf);
      }
    }
    m->stages.push_back(sr);
  }
  if (!m->nodes.empty()) {
    CC_CUDA(cudaMalloc(&m->d_nodes, sizeof(DevNode) * m->nodes.size()));
    CC_CUDA(cudaMalloc(&m->d_tree_first, sizeof(int) * m->tree_first.size()));
    CC_CUDA(cudaMalloc(&m->d_stages, sizeof(DevStageRef) * m->stages.size()));
    CC_CUDA(cudaMemcpyAsync(m->d_nodes, m->nodes.data(), sizeof(DevNode) * m->nodes.size(), cudaMemcpyHostToDevice, ctx->stream));
    CC_CUDA(cudaMemcpyAsync(m->d_tree_first, m->tree_first.data(), sizeof(int) * m->tree_first.size(), cudaMemcpyHostToDevice, ctx->stream));
    CC_CUDA(cudaMemcpyAsync(m->d_stages, m->stages.data(), sizeof(DevStageRef) * m->stages.size(), cudaMemcpyHostToDevice, ctx->stream));
    CC_CUDA(cudaMalloc(&m->d_lfeats, sizeof(LvlFeat) * m->lfeats.size()));
    CC_CUDA(cudaMemcpyAsync(m->d_lfeats, m->lfeats.data(), sizeof(LvlFeat) * m->lfeats.size(), cudaMemcpyHostToDevice, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return CC_OK;
}

static int grid_for(cc_ctx* ctx, long long total, int threads) {
  return (int)std::max<long long>(1, std::min<long long>((total + threads - 1) / threads, (long long)ctx->sm_count * 32));
}

// Evaluate `n` windows that sit in scratch slots [0, n): integrals + cascade; returns sorted survivors.
static int eval_scratch(cc_ctx* ctx, MineState* m, int n, std::vector<unsigned>& survivors) {
  survivors.clear();
  if (n == 0) return CC_OK;
  if (m->stages.empty()) {
    survivors.resize(n);
    for (int i = 0; i < n; ++i) survivors[i] = (unsigned)i;
    return CC_OK;
  }
  CC_CHECK(launch_integrals(ctx, m->scratch, 0, n));
  {
    ScopedTimer tm(ctx, "mine", 1);
    CC_CUDA(cudaMemsetAsync(m->d_surv, 0, sizeof(unsigned), ctx->stream));
    k_cascade_eval<<<(n + 127) / 128, 128, 0, ctx->stream>>>(make_eval_view_store(ctx, m->scratch), m->d_nodes, m->d_tree_first, m->d_stages,
                                                             (int)m->stages.size(), n, m->d_surv, m->d_surv + 1, m->surv_cap);
    CC_CUDA(cudaGetLastError());
  }
  CC_CUDA(cudaMemcpyAsync(m->h_surv, m->d_surv, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  const unsigned cnt = std::min(m->h_surv[0], m->surv_cap);
  if (cnt) {
    CC_CUDA(cudaMemcpyAsync(m->h_surv + 1, m->d_surv + 1, sizeof(unsigned) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
    survivors.assign(m->h_surv + 1, m->h_surv + 1 + cnt);
    std::sort(survivors.begin(), survivors.end());
  }
  return CC_OK;
}

// HOG miner (see k_hog_window_cascade): the batch's levels are in m->d_pool / m->d_levels; returns sorted survivors
constexpr long long HOG_WINDOW_MAX_BATCH = 2 << 20;      // ~1600 level pixels per 48x96 window: 3 GB of level pixels per batch
static bool use_hog_window_path(const cc_ctx* ctx, const MineState* m) {
  if (ctx->ftype != CC_FEATURE_HOG || m->force_slots || m->stages.empty() || ctx->W < 3 || ctx->H < 3) return false;
  // one round of the column pass (a thread per kept column and plane), kept rows in registers, a gradient segment per W threads
  if ((ctx->W / HOG_GRID) * 10 > HOGW_THREADS || ctx->H / HOG_GRID > HOGW_MAXQ || ctx->W > HOGW_THREADS - 32) return false;
  return hog_win_geometry(ctx->W, ctx->H).smem <= (size_t)200 * 1024;
}
// Image-sharded mining (an NCCL communicator is attached): every rank evaluates a contiguous slice [w0, w1) of the batch's windows
// and builds only the levels [l0, l0 + nl) that hold them; the survivor sets are merged as bitmaps (collect_survivors).
struct HogShard { long long w0 = 0, w1 = 0; int l0 = 0, nl = 0, tiles = 0; bool sharded = false; };
static HogShard hog_shard(const cc_ctx* ctx, const MineState* m, std::vector<LevelDesc>& levels, long long nwin) {
  HogShard h;
  h.sharded = ctx->comm_world > 1 && ctx->nccl_comm && m->shard_images;
  h.w0 = 0; h.w1 = nwin;
  if (h.sharded) { h.w0 = nwin * ctx->comm_rank / ctx->comm_world; h.w1 = nwin * (ctx->comm_rank + 1) / ctx->comm_world; }
  if (h.w1 <= h.w0 || levels.empty()) return h;
  auto level_of = [&](long long wi) {
    int lo = 0, hi = (int)levels.size() - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (levels[mid].w_base <= wi) lo = mid; else hi = mid - 1; }
    return lo;
  };
  h.l0 = level_of(h.w0);
  const int l1 = level_of(h.w1 - 1);
  h.nl = l1 - h.l0 + 1;
  h.tiles = resize_tiles(levels, h.l0, l1);
  return h;
}
static int collect_survivors(cc_ctx* ctx, MineState* m, bool sharded, long long nwin, const unsigned int* d_count, const unsigned int* d_list,
                             std::vector<unsigned>& survivors);
static int eval_hog_windows(cc_ctx* ctx, MineState* m, int nlevels, long long nwin, const HogShard& hs, std::vector<unsigned>& survivors) {
  survivors.clear();
  if (nwin == 0) return CC_OK;
  if (!m->d_hog_lut) {
    CC_CUDA(cudaMalloc(&m->d_hog_lut, sizeof(float2) * HOG_LUT_N));
    k_hog_lut_init<<<(HOG_LUT_N + 255) / 256, 256, 0, ctx->stream>>>(m->d_hog_lut);
    CC_CUDA(cudaGetLastError());
  }
  const size_t nblk = (size_t)((nwin + HOGW_BLK - 1) / HOGW_BLK);
  if (nblk > m->hog_blk_cap) {
    cudaFree(m->d_hog_blk); m->d_hog_blk = nullptr; m->hog_blk_cap = 0;
    CC_CUDA(cudaMalloc(&m->d_hog_blk, sizeof(int) * (nblk + nblk / 4)));
    m->hog_blk_cap = nblk + nblk / 4;
  }
  const HogWinGeom g = hog_win_geometry(ctx->W, ctx->H);
  {
    ScopedTimer tm(ctx, "mine", 2);
    k_hog_block_levels<<<(unsigned)((nblk + 255) / 256), 256, 0, ctx->stream>>>(m->d_levels, nlevels, nwin, m->d_hog_blk);
    CC_CUDA(cudaGetLastError());
    CC_CUDA(cudaMemsetAsync(m->d_surv, 0, sizeof(unsigned), ctx->stream));
    const int per_sm = std::max(1, (int)((size_t)227 * 1024 / (g.smem + 1024)));
    const int grid = (int)std::min<long long>(hs.w1 - hs.w0, (long long)ctx->sm_count * per_sm * 4);
#define HOGW_LAUNCH(NQ)                                                                                                                      \
  {                                                                                                                                          \
    CC_CUDA(cudaFuncSetAttribute(k_hog_window_cascade<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));                       \
    k_hog_window_cascade<NQ><<<grid, HOGW_THREADS, g.smem, ctx->stream>>>(m->d_levels, nlevels, m->d_hog_blk, m->d_pool, m->d_hog_lut, g,    \
                                                                           ctx->d_hog_taps, m->d_nodes, m->d_tree_first, m->d_stages,        \
                                                                           (int)m->stages.size(), hs.w0, hs.w1, m->d_surv, m->d_surv + 1, m->surv_cap); \
  }
    if (grid > 0) switch (g.rows / HOG_GRID) {
      case 24: HOGW_LAUNCH(24); break;      // 96 rows (BASELINE configs[3])
      case 32: HOGW_LAUNCH(32); break;      // 128
      case 16: HOGW_LAUNCH(16); break;      // 64
      default: HOGW_LAUNCH(0); break;
    }
#undef HOGW_LAUNCH
    CC_CUDA(cudaGetLastError());
  }
  return collect_survivors(ctx, m, hs.sharded, nwin, m->d_surv, m->d_surv + 1, survivors);
}

// copy accepted scratch windows into training slots [first, first+n) and build their integrals
static int accept_windows(cc_ctx* ctx, MineState* m, const unsigned* which, int n, int first, int label) {
  if (n == 0) return CC_OK;
  const int npx = ctx->W * ctx->H;
  // reuse the survivor device buffer as the index list
  CC_CUDA(cudaMemcpyAsync(m->d_surv + 1, which, sizeof(unsigned) * n, cudaMemcpyHostToDevice, ctx->stream));
  {
    ScopedTimer tm(ctx, "mine", 1);
    k_copy_slots<<<grid_for(ctx, (long long)n * npx, 256), 256, 0, ctx->stream>>>(m->scratch.d_imgs, m->d_surv + 1, n, npx,
                                                                                 ctx->d_imgs + (size_t)first * npx, ctx->d_labels + first, label);
    CC_CUDA(cudaGetLastError());
  }
  ctx->precalc_valid = false;
  CC_CHECK(launch_integrals(ctx, main_store(ctx), first, n));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));   // `which` (host) and d_surv are reused by the caller
  return CC_OK;
}

// build one batch: whole levels (or a bounded part of one) starting at the stream position
// A batch is a run of consecutive stream windows: level parts that are evaluated on the GPU (`levels`; w_base = their dense
// "evaluation ordinal") interleaved with whole pyramids answered from the survivor cache (`pyrs`).  `segs` lists both kinds in
// stream order with their batch-relative stream ordinals.
struct CachedPyr {
  NegStream start;                  // stream state on the pyramid's first window (right behind its last one: start.nextImg())
  long long base = 0, nwin = 0;     // stream ordinals [base, base + nwin)
  uint64_t key = 0;
  size_t cand0 = 0, ncand = 0;      // its cached survivors in Batch::cand
};
struct MissPyr { uint64_t key = 0; long long base = 0, nwin = 0; bool complete = false; };      // evaluated pyramid that may enter the cache
// A run of consecutive cached pyramids WITHOUT survivors (most pyramids in late stages): only its extent matters -- no window of it
// can be accepted or re-evaluated -- so the run is one record however long it is; a stop position inside it is found by walking.
struct EmptyRun { NegStream start, after; long long base = 0, nwin = 0; int count = 0; };
struct Seg { long long base, nwin; int kind, idx; };      // kind 0: levels[idx] (a level part), 1: pyrs[idx], 2: runs[idx]
struct Batch {
  std::vector<LevelDesc> levels;
  std::vector<NegStream> before;    // stream state on the first window taken from each level part
  std::vector<NegStream> after;     // stream state right after the last window of each level part (level granularity)
  std::vector<long long> lv_stream; // stream ordinal of the first window taken from each level part
  std::vector<CachedPyr> pyrs;
  std::vector<EmptyRun> runs;
  std::vector<MissPyr> miss;
  std::vector<Seg> segs;
  std::vector<long long> cand;      // stream ordinals of the cached survivors (ascending)
  std::vector<uint8_t> cand_fs;     // stages they are known to pass
  std::vector<uint32_t> cand_eval;  // indices into cand that still have stages to run
  long long nwin = 0;               // stream windows
  long long neval = 0;              // windows in `levels`
  size_t pool_px = 0;
  bool has_stale = false;           // some cached pyramid of the batch has stages to run on its survivors ...
  uint64_t first_stale_key = 0;     // ... the first of them: the batch ends before the stream returns to it (see assemble)
};

// windows the fast path may put in one batch (pyramid pixels ~ 154 per window at 1920x1080 -> ~2 GB of planes)
// per rank (image-sharded mining scales the batch with the number of ranks): 4 M windows = ~8 GB of level pixels + planes.
// Large batches keep every SM busy through the tail of the builder and amortise the per-batch host round trip.
constexpr long long LEVEL_PATH_MAX_BATCH = 4 << 20;
static long long level_path_max_batch(const cc_ctx* ctx, const MineState* m);

static long long level_path_max_batch(const cc_ctx* ctx, const MineState* m) {
  const bool sharded = ctx->comm_world > 1 && ctx->nccl_comm && m->shard_images;
  static const long long override_batch = [] { const char* e = getenv("CC_MINE_BATCH"); return e ? atoll(e) : 0ll; }();   // experiments only
  if (override_batch > 0) return override_batch;
  return LEVEL_PATH_MAX_BATCH * (sharded ? std::min(ctx->comm_world, 16) : 1);
}

// image-sharded mining: a rank builds 1 / world of a batch
static double sharded_alloc_div(const cc_ctx* ctx, const MineState* m) {
  return (ctx->comm_world > 1 && ctx->nccl_comm && m->shard_images) ? (double)std::min(ctx->comm_world, 16) : 1.0;
}

// host wall time of the batch loop by phase (cc_timing_get "mine_host_assemble" / "mine_host_eval" / "mine_host_cut"; launches = batches)
struct HostTimer {
  cc_ctx* ctx; const char* name; std::chrono::steady_clock::time_point t0;
  bool done = false;
  HostTimer(cc_ctx* c, const char* n) : ctx(c), name(n), t0(std::chrono::steady_clock::now()) {}
  ~HostTimer() { stop(); }
  void stop() {
    if (done) return;
    done = true;
    if (!ctx->timing_enabled) return;
    TimerSlot& s = ctx->timers[name];
    s.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    s.launches += 1;
  }
};

// windows per batch of the per-window slot path: HOG keeps 10 compact f32 integral planes per slot + 5 bytes of gradient scratch per
// pixel (36 KB for 48x96 windows), so the batch is sized to ~8 GB of scratch, at most 256 K windows -- large batches amortise the
// launches and the host round trip of a batch
static long long slot_path_max_batch(const cc_ctx* ctx) {
  if (ctx->ftype != CC_FEATURE_HOG) return 262144;
  const double slot_bytes = 10.0 * ctx->hogP * 4 + 6.0 * ctx->W * ctx->H;
  return (long long)std::min(262144.0, std::max(4096.0, 8.0e9 / slot_bytes));
}

static bool use_level_path(const cc_ctx* ctx, const MineState* m) {
  return !m->force_slots && !m->stages.empty() && (ctx->ftype == CC_FEATURE_HAAR || ctx->ftype == CC_FEATURE_LBP) && ctx->W <= 255 && ctx->H <= 255;
}

// survivors of a batch -> ascending ordinals on the host.  Image-sharded: the ranks' (disjoint) survivor sets are merged
// as bitmaps summed over NCCL, so every rank ends up with the same ordered list.
static int collect_survivors(cc_ctx* ctx, MineState* m, bool sharded, long long nwin, const unsigned int* d_count, const unsigned int* d_list,
                             std::vector<unsigned>& survivors) {
  if (sharded) {
    const size_t words = ((size_t)nwin + 31) / 32;
    if (words > m->bitmap_cap) {
      cudaFree(m->d_bitmap); cudaFreeHost(m->h_bitmap);
      const size_t cap = words + words / 4 + 64;
      CC_CUDA(cudaMalloc(&m->d_bitmap, cap * 4)); CC_CUDA(cudaMallocHost(&m->h_bitmap, cap * 4));
      m->bitmap_cap = cap;
    }
    CC_CUDA(cudaMemsetAsync(m->d_bitmap, 0, words * 4, ctx->stream));
    k_list_to_bitmap<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(d_list, d_count, m->d_bitmap);
    CC_CUDA(cudaGetLastError());
    ctx->launches += 1;
    CC_CHECK(allreduce_sum(ctx, m->d_bitmap, words, /*uint32*/ 4));
    CC_CUDA(cudaMemcpyAsync(m->h_bitmap, m->d_bitmap, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t w = 0; w < words; ++w) {
      unsigned int bits = m->h_bitmap[w];
      while (bits) { const int k = __builtin_ctz(bits); bits &= bits - 1; survivors.push_back((unsigned)(w * 32 + k)); }
    }
    return CC_OK;
  }
  CC_CUDA(cudaMemcpyAsync(m->h_surv, d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  const unsigned cnt = m->h_surv[0];
  if (cnt) {
    CC_CUDA(cudaMemcpyAsync(m->h_surv + 1, d_list, sizeof(unsigned) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
    survivors.assign(m->h_surv + 1, m->h_surv + 1 + cnt);
    std::sort(survivors.begin(), survivors.end());
  }
  return CC_OK;
}

static int ensure_lists(cc_ctx* ctx, MineState* m, long long nwin) {
  if ((size_t)nwin > m->list_cap) {
    cudaFree(m->d_list[0]); cudaFree(m->d_list[1]);
    const size_t cap = (size_t)nwin + (size_t)nwin / 2;
    CC_CUDA(cudaMalloc(&m->d_list[0], cap * 4)); CC_CUDA(cudaMalloc(&m->d_list[1], cap * 4));
    m->list_cap = cap;
  }
  CC_CHECK(ensure_h_surv(m, m->list_cap + 1));
  if (!m->d_counts) CC_CUDA(cudaMalloc(&m->d_counts, sizeof(unsigned) * 16));
  return CC_OK;
}

// Tile geometry of the fused miner: TW x TH windows per CTA such that the tile's planes fit ~100 KB (two CTAs per SM)
struct TileGeom { int TW, TH, tpw, tph, pp; size_t smem; bool ok; };
static TileGeom tile_geometry(const cc_ctx* ctx, bool want_sq, bool want_tilted) {
  TileGeom g{};
  const int sx = std::max(1, (int)(0.5F * ctx->W)), sy = std::max(1, (int)(0.5F * ctx->H));
  const int nplanes = 1 + (want_sq ? 1 : 0) + (want_tilted ? 1 : 0);
  for (int t = 8; t >= 1; --t) {
    const int tpw = (t - 1) * sx + ctx->W, tph = (t - 1) * sy + ctx->H;
    const int pp = (tpw + 1) | 1;                  // odd pitch: column, row and diagonal walks are bank-conflict-free
    const size_t smem = (size_t)nplanes * pp * (tph + 1) * 4 + (size_t)((tpw * tph + 3) & ~3) + (size_t)(tpw + tph) * 4;
    if (smem <= 100 * 1024 || (t == 1 && smem <= 200 * 1024)) { g = TileGeom{t, t, tpw, tph, pp, smem, true}; return g; }
  }
  return g;
}

// Fused tile miner: levels of the batch -> tiles -> k_tile_cascade; returns the sorted surviving batch ordinals
static int eval_tiles(cc_ctx* ctx, MineState* m, Batch& bt, const TileGeom& g, std::vector<unsigned>& survivors) {
  survivors.clear();
  if (bt.nwin == 0) return CC_OK;
  const bool sharded = ctx->comm_world > 1 && ctx->nccl_comm && m->shard_images;
  const bool haar = ctx->ftype == CC_FEATURE_HAAR;
  const bool tilted = haar && m->cascade_tilted;
  std::vector<LevelDesc>& dl = m->dev_levels;
  std::vector<TileDesc>& tiles = m->tiles;
  dl.clear(); tiles.clear();
  {
    int img_ord = -1;
    const uint8_t* prev = nullptr;
    for (const LevelDesc& L : bt.levels) {
      if (L.src != prev) { ++img_ord; prev = L.src; }
      if (sharded && (img_ord % ctx->comm_world) != ctx->comm_rank) continue;      // image-sharded mining: another rank's image
      const int li = (int)dl.size();
      dl.push_back(L);
      const long long k_lo = L.w_skip, k_hi = L.w_skip + L.w_count;
      for (int ty = 0; ty * g.TH < L.ny; ++ty) {
        const long long r_lo = (long long)ty * g.TH * L.nx, r_hi = (long long)std::min(L.ny, (ty + 1) * g.TH) * L.nx;
        if (r_hi <= k_lo || r_lo >= k_hi) continue;                               // tile row outside the part of the level taken by the batch
        for (int tx = 0; tx * g.TW < L.nx; ++tx) tiles.push_back(TileDesc{li, tx, ty});
      }
    }
  }
  const int nlv = (int)dl.size(), ntiles = (int)tiles.size();
  if (nlv > m->levels_cap) { cudaFree(m->d_levels); CC_CUDA(cudaMalloc(&m->d_levels, sizeof(LevelDesc) * nlv * 2)); m->levels_cap = nlv * 2; }
  if (ntiles > m->tiles_cap) { cudaFree(m->d_tiles); CC_CUDA(cudaMalloc(&m->d_tiles, sizeof(TileDesc) * ntiles * 2)); m->tiles_cap = ntiles * 2; }
  CC_CHECK(ensure_lists(ctx, m, bt.nwin));
  if (nlv) CC_CUDA(cudaMemcpyAsync(m->d_levels, dl.data(), sizeof(LevelDesc) * nlv, cudaMemcpyHostToDevice, ctx->stream));
  if (ntiles) CC_CUDA(cudaMemcpyAsync(m->d_tiles, tiles.data(), sizeof(TileDesc) * ntiles, cudaMemcpyHostToDevice, ctx->stream));
  CC_CUDA(cudaMemsetAsync(m->d_counts, 0, sizeof(unsigned) * 16, ctx->stream));
  if (ntiles) {
    ScopedTimer tm(ctx, "mine", 1);
    CascadeArgs A;
    A.pl = LvlPlanes{nullptr, nullptr, nullptr};
    A.ftype = ctx->ftype; A.W = ctx->W; A.H = ctx->H;
    A.nodes = m->d_nodes; A.feats = m->d_lfeats; A.tree_first = m->d_tree_first; A.stages = m->d_stages;
    A.s_begin = 0; A.s_end = (int)m->stages.size();
    CC_CUDA(cudaFuncSetAttribute(k_tile_cascade, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    k_tile_cascade<<<ntiles, 256, g.smem, ctx->stream>>>(m->d_levels, m->d_tiles, g.TW, g.TH, g.tpw, g.tph, g.pp, haar ? 1 : 0, tilted ? 1 : 0, A,
                                                          m->d_counts, m->d_list[0]);
    CC_CUDA(cudaGetLastError());
  }
  return collect_survivors(ctx, m, sharded, bt.nwin, m->d_counts, m->d_list[0], survivors);
}

// Fast path, asynchronous half: build the levels and their integral planes, run the cascade in stage groups over the
// shrinking survivor list, turn the survivors into a bitmap (summed over the ranks when the images are sharded) and
// start its copy to the host.  Nothing here waits for the GPU, so the host assembles and launches batch k+1 while
// batch k is evaluated; finish_levels(slot) waits for the bitmap and returns the ascending surviving ordinals.
static int launch_levels(cc_ctx* ctx, MineState* m, Batch& bt, int slot) {
  if (bt.neval == 0) return CC_OK;
  // Tall levels are cut into horizontal strips of STRIP_WIN_ROWS window rows (consecutive strips overlap by H - step
  // pixel rows): window sums are differences, so every strip can carry its own integral planes, and a strip is a
  // much shorter serial chain of rows for k_lvl_build than the whole level.  The window ordinals of a strip are a
  // contiguous range of its level's, so a strip is described like a level (the device code does not distinguish).
  constexpr int STRIP_WIN_ROWS = 8;
  HostTimer hstrips(ctx, "mine_host_strips");
  std::vector<LevelDesc>& dl = m->strips[slot];
  dl.clear();
  size_t pool_px = 0;
  {
    int img_ord = -1;
    const uint8_t* prev = nullptr;
    // image-sharded mining: the background images of the batch go round-robin to the ranks; every rank
    // builds and evaluates only the levels of its images, the survivor bitmaps are summed over NVLink, and every rank
    // then applies the same serial stopping rule to the same ordered survivor list
    const bool sharded_ = ctx->comm_world > 1 && ctx->nccl_comm && m->shard_images;
    for (const LevelDesc& L : bt.levels) {
      if (L.src != prev) { ++img_ord; prev = L.src; }
      const int owned = sharded_ ? ((img_ord % ctx->comm_world) == ctx->comm_rank) : 1;
      if (!owned) continue;              // the owner builds, evaluates and (if accepted) supplies the pixels of these windows
      const long long k_lo = L.w_skip, k_hi = L.w_skip + L.w_count;          // window indices of this level taken by the batch
      for (int a = 0; a < L.ny; a += STRIP_WIN_ROWS) {
        const int b = std::min(L.ny, a + STRIP_WIN_ROWS);
        const long long s_lo = std::max<long long>(k_lo, (long long)a * L.nx), s_hi = std::min<long long>(k_hi, (long long)b * L.nx);
        if (s_lo >= s_hi) continue;
        LevelDesc S = L;
        S.row0 = L.y0 + a * L.step_y;
        S.y0 = 0;
        S.lh = (b - a - 1) * L.step_y + ctx->H;
        S.ny = b - a;
        S.w_skip = s_lo - (long long)a * L.nx;
        S.w_count = s_hi - s_lo;
        S.w_base = L.w_base + (s_lo - k_lo);
        S.px_off = (long long)pool_px;
        S.owned = owned;
        S.pitch = (S.lw + 1 + 3) & ~3;
        S.ppitch = (S.lw + 3) & ~3;
        pool_px += (size_t)S.ppitch * S.lh;
        dl.push_back(S);
      }
    }
  }
  const int nlv = (int)dl.size();
  size_t plane_elems = 0;
  int rows = 0, cols = 0, diag = 0;
  for (LevelDesc& L : dl) {
    L.ip_off = (long long)plane_elems; L.rows_base = rows; L.cols_base = cols; L.diag_base = diag;
    plane_elems += (size_t)L.pitch * (L.lh + 1);
    rows += L.lh; cols += L.lw + 1; diag += L.lw + L.lh - 1;
  }
  const bool sharded = ctx->comm_world > 1 && ctx->nccl_comm && m->shard_images;
  const bool haar = ctx->ftype == CC_FEATURE_HAAR;
  const bool tilted = haar && m->cascade_tilted;
  hstrips.stop();
  // Buffers grow with the batches (which grow from stage to stage as the acceptance rate falls): grow by 4x, but not beyond what
  // a full batch will need, so that a cascade reallocates two or three times instead of once per stage (cudaFree / cudaMalloc /
  // cudaMallocHost of gigabytes synchronise the device and take milliseconds each).
  HostTimer halloc(ctx, "mine_host_alloc");
  const double full_batch = std::max(1.0, (double)level_path_max_batch(ctx, m) / (double)std::max<long long>(1, bt.neval) / (sharded_alloc_div(ctx, m)));
  auto grown = [&](size_t need) -> size_t { return (size_t)((double)need * std::min(4.0, std::max(1.1, 1.1 * full_batch))); };
  bool fused = !m->force_unfused;
  int src_pitch_max = 16;
  for (const LevelDesc& L : dl) { fused = fused && L.lw <= LB_MAXDIM && L.lh <= LB_MAXDIM && L.sh <= 65535 && L.spitch <= 8192; src_pitch_max = std::max(src_pitch_max, L.spitch); }
  if (!fused && pool_px > m->pool_cap) {      // (the fused builder keeps no level pixels)
    cudaFree(m->d_pool); m->d_pool = nullptr;
    const size_t cap = grown(pool_px);
    CC_CUDA(cudaMalloc(&m->d_pool, cap)); m->pool_cap = cap;
  }
  if (nlv > m->strips_cap[slot]) {
    cudaFree(m->d_strips[slot]); cudaFreeHost(m->h_strips[slot]);
    m->d_strips[slot] = nullptr; m->h_strips[slot] = nullptr; m->strips_cap[slot] = 0;
    const int cap = (int)grown((size_t)nlv) + 64;
    CC_CUDA(cudaMalloc(&m->d_strips[slot], sizeof(LevelDesc) * cap));
    CC_CUDA(cudaMallocHost(&m->h_strips[slot], (sizeof(LevelDesc) + sizeof(int)) * cap));      // strips, then the builder's launch order
    m->strips_cap[slot] = cap;
  }
  LevelDesc* d_levels = m->d_strips[slot];
  if (plane_elems > m->plane_cap) {
    cudaFree(m->d_psum); cudaFree(m->d_psq); cudaFree(m->d_ptilt);
    m->d_psum = m->d_psq = m->d_ptilt = nullptr;
    const size_t cap = grown(plane_elems);
    CC_CUDA(cudaMalloc(&m->d_psum, cap * 4));
    if (haar) { CC_CUDA(cudaMalloc(&m->d_psq, cap * 4)); CC_CUDA(cudaMalloc(&m->d_ptilt, cap * 4)); }
    m->plane_cap = cap;
  }
  if ((size_t)bt.neval > m->list_cap) {
    cudaFree(m->d_list[0]); cudaFree(m->d_list[1]);
    const size_t cap = grown((size_t)bt.neval);
    CC_CUDA(cudaMalloc(&m->d_list[0], cap * 4)); CC_CUDA(cudaMalloc(&m->d_list[1], cap * 4));
    m->list_cap = cap;
  }
  if (!m->d_counts) CC_CUDA(cudaMalloc(&m->d_counts, sizeof(unsigned) * 16));
  {
    const size_t words = ((size_t)bt.neval + 31) / 32;
    if (words > m->bitmap_cap) {
      cudaFree(m->d_bitmap);
      const size_t cap = grown(words) + 64;
      CC_CUDA(cudaMalloc(&m->d_bitmap, cap * 4));
      m->bitmap_cap = cap;
    }
    if (words > m->h_bits_cap[slot]) {
      cudaFreeHost(m->h_bits[slot]);
      const size_t cap = grown(words) + 64;
      CC_CUDA(cudaMallocHost(&m->h_bits[slot], cap * 4));
      m->h_bits_cap[slot] = cap;
    }
  }
  halloc.stop();
  if (nlv) {
    memcpy(m->h_strips[slot], dl.data(), sizeof(LevelDesc) * nlv);
    CC_CUDA(cudaMemcpyAsync(d_levels, m->h_strips[slot], sizeof(LevelDesc) * nlv, cudaMemcpyHostToDevice, ctx->stream));
  }
  int ngroups = 0;
  {
    ScopedTimer tm(ctx, "mine", 0);
    int launches = 0;
    if (fused) {
      // three width classes (CTA of 128 / 256 / 512 threads, 4 columns per thread), each tallest-first
      HostTimer horder(ctx, "mine_host_order");
      std::vector<int> order[3];
      for (int i = 0; i < nlv; ++i)
        if (dl[i].owned) order[dl[i].lw <= 511 ? 0 : dl[i].lw <= 1023 ? 1 : 2].push_back(i);
      if (nlv > m->order_cap) { cudaFree(m->d_order); CC_CUDA(cudaMalloc(&m->d_order, sizeof(int) * nlv * 2)); m->order_cap = nlv * 2; }
      std::vector<int> flat;
      int off[4] = {0, 0, 0, 0};
      for (int c = 0; c < 3; ++c) {
        std::stable_sort(order[c].begin(), order[c].end(), [&](int a, int b) { return dl[a].lh > dl[b].lh; });
        flat.insert(flat.end(), order[c].begin(), order[c].end());
        off[c + 1] = (int)flat.size();
      }
      horder.stop();
      if (!flat.empty()) {
        // through the pinned staging buffer: an async copy from pageable memory of this size waits for the stream to drain, i.e.
        // for the batch in flight -- the host would stall for a whole batch and the two-batch pipeline would run in lock step
        int* h_order = reinterpret_cast<int*>(m->h_strips[slot] + m->strips_cap[slot]);
        std::memcpy(h_order, flat.data(), sizeof(int) * flat.size());
        CC_CUDA(cudaMemcpyAsync(m->d_order, h_order, sizeof(int) * flat.size(), cudaMemcpyHostToDevice, ctx->stream));
      }
      uint32_t* psq = haar ? m->d_psq : nullptr;
      uint32_t* ptl = tilted ? m->d_ptilt : nullptr;
      // widest class first: its levels are the longest chains
      // the accepted windows are recomputed from the source images (k_gather_resize), so the fused builder keeps no level pixels
      uint8_t* pool = nullptr;
      const size_t ssm = (size_t)8 * src_pitch_max;
      if (ssm > 24 * 1024) {                          // static + dynamic beyond 48 KB needs the opt-in (images wider than 3072 pixels)
        CC_CUDA(cudaFuncSetAttribute(k_lvl_build<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CC_CUDA(cudaFuncSetAttribute(k_lvl_build<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CC_CUDA(cudaFuncSetAttribute(k_lvl_build<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      }
      // the classes write disjoint strips: run them side by side so that the tail of one overlaps the others
      if (!m->ev_fork) {
        CC_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
          CC_CUDA(cudaStreamCreateWithFlags(&m->side[k], cudaStreamNonBlocking));
          CC_CUDA(cudaEventCreateWithFlags(&m->ev_join[k], cudaEventDisableTiming));
        }
      }
      CC_CUDA(cudaEventRecord(m->ev_fork, ctx->stream));
      CC_CUDA(cudaStreamWaitEvent(m->side[0], m->ev_fork, 0));
      CC_CUDA(cudaStreamWaitEvent(m->side[1], m->ev_fork, 0));
      if (off[3] > off[2]) { k_lvl_build<512><<<off[3] - off[2], 512, ssm, ctx->stream>>>(d_levels, m->d_order + off[2], pool, m->d_psum, psq, ptl, src_pitch_max, ctx->H); ++launches; }
      if (off[2] > off[1]) { k_lvl_build<256><<<off[2] - off[1], 256, ssm, m->side[0]>>>(d_levels, m->d_order + off[1], pool, m->d_psum, psq, ptl, src_pitch_max, ctx->H); ++launches; }
      if (off[1] > off[0]) { k_lvl_build<128><<<off[1] - off[0], 128, ssm, m->side[1]>>>(d_levels, m->d_order + off[0], pool, m->d_psum, psq, ptl, src_pitch_max, ctx->H); ++launches; }
      CC_CUDA(cudaEventRecord(m->ev_join[0], m->side[0]));
      CC_CUDA(cudaEventRecord(m->ev_join[1], m->side[1]));
      CC_CUDA(cudaStreamWaitEvent(ctx->stream, m->ev_join[0], 0));
      CC_CUDA(cudaStreamWaitEvent(ctx->stream, m->ev_join[1], 0));
    } else {
      k_resize_levels<<<grid_for(ctx, (long long)pool_px, 256), 256, 0, ctx->stream>>>(d_levels, nlv, (long long)pool_px, m->d_pool);
      k_lvl_rows<<<grid_for(ctx, (long long)rows * 32, 256), 256, 0, ctx->stream>>>(d_levels, nlv, rows, m->d_pool, m->d_psum, haar ? m->d_psq : nullptr);
      launches += 2;
      if (tilted) {
        CC_CUDA(cudaMemsetAsync(m->d_ptilt, 0, plane_elems * 4, ctx->stream));
        k_lvl_diag<<<grid_for(ctx, diag, 256), 256, 0, ctx->stream>>>(d_levels, nlv, diag, m->d_pool, m->d_ptilt, 0);
        k_lvl_diag<<<grid_for(ctx, diag, 256), 256, 0, ctx->stream>>>(d_levels, nlv, diag, m->d_pool, m->d_ptilt, 1);
        launches += 2;
      }
      dim3 cgrid(grid_for(ctx, cols, 256), 3);
      k_lvl_cols<<<cgrid, 256, 0, ctx->stream>>>(d_levels, nlv, cols, m->d_psum, haar ? m->d_psq : nullptr, tilted ? m->d_ptilt : nullptr);
      launches += 1;
    }
    CC_CUDA(cudaGetLastError());
    CC_CUDA(cudaMemsetAsync(m->d_counts, 0, sizeof(unsigned) * 16, ctx->stream));
    // stage groups [0,2) [2,6) [6,end): each group runs on the survivors of the previous one
    const int ns = (int)m->stages.size();
    CascadeArgs A;
    A.pl = LvlPlanes{m->d_psum, m->d_psq, tilted ? m->d_ptilt : nullptr};
    A.ftype = ctx->ftype; A.W = ctx->W; A.H = ctx->H;
    A.nodes = m->d_nodes; A.feats = m->d_lfeats; A.tree_first = m->d_tree_first; A.stages = m->d_stages;
    int s0 = 0;
    while (s0 < ns) {
      const int s1 = ngroups == 0 ? std::min(ns, 2) : ngroups == 1 ? std::min(ns, 6) : ns;
      A.s_begin = s0; A.s_end = s1;
      if (ngroups == 0) {
        if (nlv > 0) k_lvl_cascade_all<<<std::min(nlv, ctx->sm_count * 32), 256, 0, ctx->stream>>>(d_levels, nlv, A, m->d_counts, m->d_list[0]);
      }
      else if (ngroups == 1)
        k_lvl_cascade<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_levels, nlv, A, m->d_list[(ngroups - 1) & 1], m->d_counts + ngroups - 1,
                                                                   m->d_counts + ngroups, m->d_list[ngroups & 1]);
      else      // deep stages on the few windows left: a warp per window, lanes across the trees of a stage
        k_lvl_cascade_warp<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_levels, nlv, A, m->d_list[(ngroups - 1) & 1], m->d_counts + ngroups - 1,
                                                                        m->d_counts + ngroups, m->d_list[ngroups & 1]);
      ++ngroups; ++launches;
      s0 = s1;
    }
    CC_CUDA(cudaGetLastError());
    ctx->launches += launches;
    if (ctx->timing_enabled) ctx->timers["mine"].launches += launches;
  }
  // survivors -> bitmap (1 bit per batch ordinal) -> pinned host memory, asynchronously
  const size_t words = ((size_t)bt.neval + 31) / 32;
  if (words > m->bitmap_cap) {
    cudaFree(m->d_bitmap);
    const size_t cap = words + words / 2 + 64;
    CC_CUDA(cudaMalloc(&m->d_bitmap, cap * 4));
    m->bitmap_cap = cap;
  }
  if (words > m->h_bits_cap[slot]) {
    cudaFreeHost(m->h_bits[slot]);
    const size_t cap = words + words / 2 + 64;
    CC_CUDA(cudaMallocHost(&m->h_bits[slot], cap * 4));
    m->h_bits_cap[slot] = cap;
  }
  if (!m->ev_done[slot]) CC_CUDA(cudaEventCreateWithFlags(&m->ev_done[slot], cudaEventDisableTiming));
  CC_CUDA(cudaMemsetAsync(m->d_bitmap, 0, words * 4, ctx->stream));
  k_list_to_bitmap<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(m->d_list[(ngroups - 1) & 1], m->d_counts + ngroups - 1, m->d_bitmap);
  CC_CUDA(cudaGetLastError());
  ctx->launches += 1;
  if (sharded) CC_CHECK(allreduce_sum(ctx, m->d_bitmap, words, /*uint32*/ 4));
  CC_CUDA(cudaMemcpyAsync(m->h_bits[slot], m->d_bitmap, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaEventRecord(m->ev_done[slot], ctx->stream));      // (cc_mine_negatives records it again behind the cached survivors' pass)
  return CC_OK;
}

// second half: wait for the batch's bitmap, list the surviving ordinals in ascending order
static int finish_levels(cc_ctx* ctx, MineState* m, int slot, long long nwin, std::vector<unsigned>& survivors) {
  survivors.clear();
  if (m->ev_done[slot]) CC_CUDA(cudaEventSynchronize(m->ev_done[slot]));
  if (nwin == 0) return CC_OK;
  const size_t words = ((size_t)nwin + 31) / 32;
  const unsigned int* hb = m->h_bits[slot];
  for (size_t w = 0; w < words; ++w) {
    unsigned int bits = hb[w];
    while (bits) { const int k = __builtin_ctz(bits); bits &= bits - 1; survivors.push_back((unsigned)(w * 32 + k)); }
  }
  return CC_OK;
}

// ---- gathering windows by stream ordinal ---------------------------------------------------------------
// Pixels of arbitrary stream windows of a batch (cached survivors to re-evaluate, accepted windows) are recomputed from the
// source images by k_gather_resize.  This builds its inputs: one descriptor per pyramid level that holds at least one of
// the requested windows (whole level, private ordinal base) and the private ordinal of every window.  Every rank holds
// every background image, so no exchange is needed for these pixels.
struct GatherPlan { std::vector<LevelDesc> levels; std::vector<uint32_t> which; };

static void level_after(NegStream& w) {        // next level of the SAME pyramid (the caller knows there is one)
  w.pointX = w.offX; w.pointY = w.offY;
  w.scale *= NegStream::scaleFactor;
  const BgImage& im = (*w.bgs)[w.src];
  w.lw = (int)(w.scale * im.w);
  w.lh = (int)(w.scale * im.h);
  w.fresh = false;
}

static int build_gather_plan(const Batch& bt, const long long* ords, size_t n, GatherPlan& gp) {
  gp.levels.clear(); gp.which.resize(n);
  size_t si = 0;
  long long priv = 0;
  int cur_seg = -1;
  NegStream walk;                 // cached pyramid being walked
  long long lvl_base = 0, lvl_n = 0;      // stream ordinals [lvl_base, lvl_base + lvl_n) of the level `walk` stands on
  bool lvl_in_plan = false;
  for (size_t i = 0; i < n; ++i) {
    const long long o = ords[i];
    while (si + 1 < bt.segs.size() && bt.segs[si + 1].base <= o) ++si;
    const Seg& sg = bt.segs[si];
    CC_REQUIRE(o >= sg.base && o < sg.base + sg.nwin, CC_ERR_STATE, "mining: stream ordinal %lld outside the batch", o);
    CC_REQUIRE(sg.kind != 2, CC_ERR_STATE, "mining: a window of a survivor-free cached pyramid was requested");
    if (sg.kind == 0) {
      const LevelDesc& L = bt.levels[sg.idx];
      if (cur_seg != (int)si) {
        LevelDesc T = L;
        T.w_skip = 0; T.w_count = (long long)L.nx * L.ny; T.w_base = priv; T.owned = 1;
        priv += T.w_count;
        gp.levels.push_back(T);
        cur_seg = (int)si;
      }
      gp.which[i] = (uint32_t)(gp.levels.back().w_base + (o - sg.base) + L.w_skip);
    } else {
      const CachedPyr& cp = bt.pyrs[sg.idx];
      if (cur_seg != (int)si) {
        walk = cp.start;
        lvl_base = cp.base; lvl_n = (long long)walk.countX() * walk.countY();
        lvl_in_plan = false;
        cur_seg = (int)si;
      }
      while (o >= lvl_base + lvl_n) { lvl_base += lvl_n; level_after(walk); lvl_n = (long long)walk.countX() * walk.countY(); lvl_in_plan = false; }
      if (!lvl_in_plan) {
        LevelDesc T;
        fill_level(walk, T);
        T.w_skip = 0; T.w_count = lvl_n; T.w_base = priv; T.px_off = 0; T.ip_off = 0; T.rows_base = T.cols_base = T.diag_base = 0;
        priv += lvl_n;
        gp.levels.push_back(T);
        lvl_in_plan = true;
      }
      gp.which[i] = (uint32_t)(gp.levels.back().w_base + (o - lvl_base));
    }
    CC_REQUIRE(priv < 0xFFFFFFFFll, CC_ERR_UNSUPPORTED, "mining: too many distinct levels in one gather");
  }
  return CC_OK;
}

// accepted windows (stream ordinals of the batch, ascending) -> training slots [first, first + n) with label, + their integrals
static int gather_accepted(cc_ctx* ctx, MineState* m, const Batch& bt, const std::vector<long long>& ords, int first, int label) {
  const int n = (int)ords.size();
  if (n == 0) return CC_OK;
  GatherPlan gp;
  CC_CHECK(build_gather_plan(bt, ords.data(), ords.size(), gp));
  const size_t tb = sizeof(LevelDesc) * gp.levels.size(), need = tb + sizeof(uint32_t) * n;
  if (need > m->d_gather_cap) { cudaFree(m->d_gather); m->d_gather = nullptr; CC_CUDA(cudaMalloc(&m->d_gather, need * 2)); m->d_gather_cap = need * 2; }
  if (!m->ev_gather) CC_CUDA(cudaEventCreateWithFlags(&m->ev_gather, cudaEventDisableTiming));
  else CC_CUDA(cudaEventSynchronize(m->ev_gather));           // the previous upload has left the staging buffer (long ago)
  if (need > m->h_gather_cap) { cudaFreeHost(m->h_gather); m->h_gather = nullptr; CC_CUDA(cudaMallocHost(&m->h_gather, need * 2)); m->h_gather_cap = need * 2; }
  std::memcpy(m->h_gather, gp.levels.data(), tb);
  std::memcpy(m->h_gather + tb, gp.which.data(), sizeof(uint32_t) * n);
  CC_CUDA(cudaMemcpyAsync(m->d_gather, m->h_gather, need, cudaMemcpyHostToDevice, ctx->stream));   // pinned: does not wait for the batch in flight
  CC_CUDA(cudaEventRecord(m->ev_gather, ctx->stream));
  const int npx = ctx->W * ctx->H;
  {
    ScopedTimer tm(ctx, "mine", 1);
    k_gather_resize<<<grid_for(ctx, (long long)n * npx, 256), 256, 0, ctx->stream>>>((const LevelDesc*)m->d_gather, (int)gp.levels.size(), ctx->W, ctx->H,
                                                                                    (const unsigned int*)(m->d_gather + tb), n,
                                                                                    ctx->d_imgs + (size_t)first * npx, ctx->d_labels + first, label);
    CC_CUDA(cudaGetLastError());
  }
  ctx->precalc_valid = false;
  CC_CHECK(launch_integrals(ctx, main_store(ctx), first, n));
  return CC_OK;
}

// Cached survivors of the batch that still have stages to run: gather their pixels into scratch slots, build their integrals
// (K1, the arithmetic of setImage) and run stages [first_stage, ns) on every slot.  Asynchronous like launch_levels: the flags
// arrive in m->h_cand_flags[slot] before m->ev_done[slot] fires.
static int launch_candidates(cc_ctx* ctx, MineState* m, const Batch& bt, int slot) {
  const int n = (int)bt.cand_eval.size();
  if (n == 0) return CC_OK;
  std::vector<long long> ords(n);
  for (int i = 0; i < n; ++i) ords[i] = bt.cand[bt.cand_eval[i]];
  GatherPlan gp;
  CC_CHECK(build_gather_plan(bt, ords.data(), ords.size(), gp));
  const size_t tb = sizeof(LevelDesc) * gp.levels.size();
  const size_t off_which = tb, off_fs = off_which + sizeof(uint32_t) * (size_t)n, off_flags = (off_fs + n + 15) & ~(size_t)15, need = off_flags + n;
  if (need > m->d_cand_cap) { cudaFree(m->d_cand); m->d_cand = nullptr; CC_CUDA(cudaMalloc(&m->d_cand, need * 2)); m->d_cand_cap = need * 2; }
  if (need > m->h_cand_cap[slot]) {
    cudaFreeHost(m->h_cand[slot]); m->h_cand[slot] = nullptr;
    CC_CUDA(cudaMallocHost(&m->h_cand[slot], need * 2)); m->h_cand_cap[slot] = need * 2;
  }
  if ((size_t)n > m->h_cand_flags_cap[slot]) {
    cudaFreeHost(m->h_cand_flags[slot]); m->h_cand_flags[slot] = nullptr;
    CC_CUDA(cudaMallocHost(&m->h_cand_flags[slot], (size_t)n * 2)); m->h_cand_flags_cap[slot] = (size_t)n * 2;
  }
  if (n > m->cand_cap) {
    store_free(m->cand_store);
    const int cap = std::max(n + n / 2, 4096);
    CC_CHECK(store_alloc(ctx, m->cand_store, cap));
    m->cand_cap = cap;
  }
  uint8_t* hb = m->h_cand[slot];
  std::memcpy(hb, gp.levels.data(), tb);
  std::memcpy(hb + off_which, gp.which.data(), sizeof(uint32_t) * (size_t)n);
  for (int i = 0; i < n; ++i) hb[off_fs + i] = bt.cand_fs[bt.cand_eval[i]];
  CC_CUDA(cudaMemcpyAsync(m->d_cand, hb, off_fs + n, cudaMemcpyHostToDevice, ctx->stream));
  const int npx = ctx->W * ctx->H;
  {
    ScopedTimer tm(ctx, "mine", 1);
    k_gather_resize<<<grid_for(ctx, (long long)n * npx, 256), 256, 0, ctx->stream>>>((const LevelDesc*)m->d_cand, (int)gp.levels.size(), ctx->W, ctx->H,
                                                                                    (const unsigned int*)(m->d_cand + off_which), n,
                                                                                    m->cand_store.d_imgs, m->cand_store.d_labels, 0);
    CC_CUDA(cudaGetLastError());
  }
  CC_CHECK(launch_integrals(ctx, m->cand_store, 0, n));
  {
    ScopedTimer tm(ctx, "mine", 1);
    k_cascade_flags<<<(n + 127) / 128, 128, 0, ctx->stream>>>(make_eval_view_store(ctx, m->cand_store), m->d_nodes, m->d_tree_first, m->d_stages,
                                                              (int)m->stages.size(), n, m->d_cand + off_fs, m->d_cand + off_flags);
    CC_CUDA(cudaGetLastError());
  }
  CC_CUDA(cudaMemcpyAsync(m->h_cand_flags[slot], m->d_cand + off_flags, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  m->stat_cand += n;
  return CC_OK;
}

// Survivors of the whole batch as ascending stream ordinals: the evaluated level parts (`eval_surv` = ascending evaluation
// ordinals) merged with the cached survivors (known ones, and re-evaluated ones by their flags); then the cache is brought up
// to date with what this batch has learnt about its pyramids.
static void merge_and_update_cache(MineState* m, const Batch& bt, const std::vector<unsigned>& eval_surv, const uint8_t* flags, int ns, bool cache_on,
                                   std::vector<long long>& out) {
  out.clear();
  std::vector<long long> a, c;
  a.reserve(eval_surv.size());
  size_t li = 0;
  for (const unsigned e : eval_surv) {
    while (li + 1 < bt.levels.size() && bt.levels[li + 1].w_base <= (long long)e) ++li;
    a.push_back(bt.lv_stream[li] + ((long long)e - bt.levels[li].w_base));
  }
  size_t fj = 0;
  for (size_t i = 0; i < bt.cand.size(); ++i) {
    const bool pass = bt.cand_fs[i] >= ns ? true : flags[fj++] != 0;
    if (pass) c.push_back(bt.cand[i]);
  }
  out.resize(a.size() + c.size());
  std::merge(a.begin(), a.end(), c.begin(), c.end(), out.begin());
  if (!cache_on) return;
  PyrCache& pc = m->cache;
  auto range = [&](long long base, long long nwin, std::vector<uint32_t>& v) {
    auto lo = std::lower_bound(out.begin(), out.end(), base), hi = std::lower_bound(out.begin(), out.end(), base + nwin);
    v.clear(); v.reserve(hi - lo);
    for (auto it = lo; it != hi; ++it) v.push_back((uint32_t)(*it - base));
  };
  std::vector<uint32_t> v;
  for (const CachedPyr& cp : bt.pyrs) {
    auto it = pc.map.find(cp.key);
    if (it == pc.map.end() || it->second.n_stages >= ns) continue;
    range(cp.base, cp.nwin, v);
    pc.bytes -= it->second.surv.size() * 4;
    it->second.surv = v; it->second.n_stages = ns;
    pc.bytes += v.size() * 4;
  }
  for (const MissPyr& mp : bt.miss) {
    if (!mp.complete) continue;
    auto it = pc.map.find(mp.key);
    if (it != pc.map.end() && it->second.n_stages >= ns) continue;
    range(mp.base, mp.nwin, v);
    if ((long long)v.size() > std::max<long long>(8, mp.nwin / 16)) continue;          // re-evaluating that many windows one by one costs more than the planes
    if (it == pc.map.end()) {
      if (pc.bytes + v.size() * 4 + 64 > pc.max_bytes) continue;
      PyrEntry e; e.n_stages = ns; e.nwin = mp.nwin; e.surv = v;
      pc.map.emplace(mp.key, std::move(e));
      pc.table_valid = false;
      pc.bytes += v.size() * 4 + 64;
    } else {
      pc.bytes -= it->second.surv.size() * 4;
      it->second.surv = v; it->second.n_stages = ns; it->second.nwin = mp.nwin;
      pc.bytes += v.size() * 4;
    }
  }
}

// The serial part of fillPassedSamples (src/cascadeclassifier.cpp:333-355) applied to a batch that has already been
// evaluated: walk the windows in stream order, stop at the count-th acceptance or as soon as the reference's check
// (getcount+1)/consumed <= minimumAcceptanceRatio holds before a window.  `surv` = ascending batch ordinals that pass
// the cascade.  Returns the ordinal of the first unconsumed window; consumed / got are updated, accepted ordinals
// appended.
template <typename Ord>
static long long serial_cut(const Ord* surv, size_t nsurv, long long nwin, int count, double min_ratio, long long& consumed, int& got,
                            std::vector<Ord>& accepted, bool& stopped) {
  accepted.clear();
  stopped = false;
  long long pos = 0;            // next unconsumed batch ordinal
  size_t si = 0;
  // smallest value c >= max(c_from, 1) of `consumed` at which the check holds
  auto first_fire = [&](int g, long long c_from) -> long long {
    if (!(min_ratio > 0)) return INT64_MAX;
    long long c = std::max<long long>(1, c_from);
    const double c0 = std::ceil(((double)g + 1) / min_ratio);
    if (c0 > 9e18) return INT64_MAX;
    const long long guess = (long long)c0 - 2;
    if (guess > c) c = guess;
    while (!(((double)g + 1) / (double)c <= min_ratio)) ++c;
    return c;
  };
  if (got >= count) { stopped = true; return 0; }
  while (pos < nwin) {
    const long long nextPass = si < nsurv ? (long long)surv[si] : nwin;
    const long long gap = nextPass - pos;               // failing windows before the next survivor
    const long long fire = first_fire(got, consumed);
    if (fire <= consumed + gap) {                       // the rule fires before a window of this gap (or before the survivor)
      const long long adv = fire - consumed;
      consumed += adv; pos += adv;
      stopped = true;
      break;
    }
    consumed += gap;
    pos = nextPass;
    if (pos >= nwin) break;
    consumed++; got++;                                  // window `pos` passes the cascade
    accepted.push_back((Ord)pos);
    pos++; si++;
    if (got == count) { stopped = true; break; }
  }
  return pos;
}

}  // namespace cc

using namespace cc;

extern "C" {

int cc_bg_clear(cc_ctx* ctx) {
  CC_REQUIRE(ctx, CC_ERR_BAD_ARG, "cc_bg_clear: null ctx");
  MineState* m = mine_state(ctx);
  CC_CUDA(cudaSetDevice(ctx->device));
  for (auto& b : m->bgs) cudaFree(b.d_px);
  m->bgs.clear();
  m->cache.clear();
  return CC_OK;
}

int cc_bg_register(cc_ctx* ctx, const uint8_t* host_gray, int w, int h) {
  CC_REQUIRE(ctx && ((host_gray && w > 0 && h > 0) || (w == 0 && h == 0)), CC_ERR_BAD_ARG, "cc_bg_register: bad arguments");
  MineState* m = mine_state(ctx);
  CC_CUDA(cudaSetDevice(ctx->device));
  BgImage b;
  m->cache.clear();                                        // image indices / list length are part of the cache keys
  if (w == 0) { m->bgs.push_back(b); return CC_OK; }      // placeholder of an unreadable file: keeps its slot in the list, NegReader::nextImg skips it
  b.w = w; b.h = h; b.pitch = (w + 15) & ~15;
  CC_CUDA(cudaMalloc(&b.d_px, (size_t)b.pitch * h));
  CC_CUDA(cudaMemcpy2DAsync(b.d_px, (size_t)b.pitch, host_gray, (size_t)w, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  m->bgs.push_back(b);
  return CC_OK;
}

int cc_bg_register_device(cc_ctx* ctx, const uint8_t* dev_gray, int w, int h, int pitch) {
  CC_REQUIRE(ctx && dev_gray && w > 0 && h > 0 && pitch >= w, CC_ERR_BAD_ARG, "cc_bg_register_device: bad arguments");
  MineState* m = mine_state(ctx);
  CC_CUDA(cudaSetDevice(ctx->device));
  BgImage b;
  m->cache.clear();
  b.w = w; b.h = h; b.pitch = (w + 15) & ~15;
  CC_CUDA(cudaMalloc(&b.d_px, (size_t)b.pitch * h));
  CC_CUDA(cudaMemcpy2DAsync(b.d_px, (size_t)b.pitch, dev_gray, (size_t)pitch, (size_t)w, (size_t)h, cudaMemcpyDeviceToDevice, ctx->stream));
  m->bgs.push_back(b);
  return CC_OK;
}

int cc_mine_set_mode(cc_ctx* ctx, int mode) {
  CC_REQUIRE(ctx && mode >= 0 && mode <= 3, CC_ERR_BAD_ARG, "cc_mine_set_mode: mode must be 0 (auto: level planes, fused builder), 1 (level planes, separate kernels), 2 (per-window slots) or 3 (fused tile miner)");
  MineState* m = mine_state(ctx);
  m->mode = mode;
  m->force_unfused = mode == 1;
  m->force_slots = mode == 2;
  return CC_OK;
}

int cc_mine_set_cache(cc_ctx* ctx, int enable, int64_t max_bytes) {
  CC_REQUIRE(ctx, CC_ERR_BAD_ARG, "cc_mine_set_cache: null ctx");
  MineState* m = mine_state(ctx);
  m->cache.enabled = enable != 0;
  if (!m->cache.enabled) m->cache.clear();
  if (max_bytes > 0) m->cache.max_bytes = (size_t)max_bytes;
  return CC_OK;
}

int cc_mine_get_stats(cc_ctx* ctx, int64_t* stats6) {
  CC_REQUIRE(ctx && stats6, CC_ERR_BAD_ARG, "cc_mine_get_stats: null argument");
  MineState* m = mine_state(ctx);
  stats6[0] = m->stat_consumed; stats6[1] = m->stat_built; stats6[2] = m->stat_cand; stats6[3] = m->stat_skipped;
  stats6[4] = (int64_t)m->cache.map.size(); stats6[5] = (int64_t)m->cache.bytes;
  return CC_OK;
}

int cc_bg_count(cc_ctx* ctx) { return ctx && ctx->mine ? (int)ctx->mine->bgs.size() : 0; }

int cc_cascade_set_stages(cc_ctx* ctx, const cc_stage* stages, int n) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_cascade_set_stages: evaluator not initialised");
  CC_REQUIRE(n >= 0 && (n == 0 || stages), CC_ERR_BAD_ARG, "cc_cascade_set_stages: bad arguments");
  CC_CUDA(cudaSetDevice(ctx->device));
  return upload_cascade(ctx, mine_state(ctx), stages, n);
}

int cc_mine_negatives(cc_ctx* ctx, int first_slot, int count, double min_ratio, cc_reader_state* state, int64_t* consumed_out, int32_t* got_out) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_mine_negatives: evaluator not initialised");
  CC_REQUIRE(state && consumed_out && got_out, CC_ERR_BAD_ARG, "cc_mine_negatives: null argument");
  CC_REQUIRE(first_slot >= 0 && count >= 0 && first_slot + count <= ctx->max_samples, CC_ERR_BAD_ARG,
             "cc_mine_negatives: slots [%d,%d) exceed maxSampleCount %d", first_slot, first_slot + count, ctx->max_samples);
  MineState* m = mine_state(ctx);
  CC_CUDA(cudaSetDevice(ctx->device));
  *consumed_out = 0; *got_out = 0;
  if (count == 0) return CC_OK;
  CC_REQUIRE(!m->bgs.empty(), CC_ERR_STATE, "cc_mine_negatives: no background images registered");
  const int W = ctx->W, H = ctx->H, npx = W * H;
  NegStream ns;
  ns.bgs = &m->bgs; ns.W = W; ns.H = H;
  stream_from_state(ns, *state, (int)m->bgs.size());

  long long consumed = 0;
  int got = 0;
  double accept_rate = m->stages.empty() ? 1.0 : 0.0;   // running estimate
  const bool level_path = use_level_path(ctx, m);
  const TileGeom tgeom = tile_geometry(ctx, ctx->ftype == CC_FEATURE_HAAR, ctx->ftype == CC_FEATURE_HAAR && m->cascade_tilted);
  const bool tile_path = level_path && m->mode == 3 && tgeom.ok;
  const bool hog_win = !level_path && use_hog_window_path(ctx, m);      // HOG: level gradients + one CTA per window
  const long long max_batch = level_path ? level_path_max_batch(ctx, m) : hog_win ? HOG_WINDOW_MAX_BATCH : slot_path_max_batch(ctx);
  // Plane path: two batches in flight.  While batch k is evaluated the host assembles and launches batch k+1 -- when
  // the running acceptance rate says that batch k will not complete the stage -- and only then waits for batch k's
  // survivor bitmap; a speculative batch that turns out not to be needed is dropped (its windows count as not read).
  const bool pipelined = level_path && !tile_path;
  const bool cache_on = pipelined && m->cache.enabled && m->mode == 0 && m->stages.size() < 255;
  const int nstages = (int)m->stages.size();
  constexpr long long CAND_CAP = 262144;         // cached survivors re-evaluated per batch (scratch slots)
  constexpr size_t SEG_CAP = (size_t)1 << 20;    // level parts + cached pyramids per batch
  bool stream_ok = true;
  if (ns.lw == 0) stream_ok = ns.nextImg();             // first get(): img.empty() -> nextImg()
  std::vector<unsigned> survivors;
  std::vector<long long> surv64, accepted;

  // ---- assemble a batch: level parts (or whole cached pyramids) of the stream starting at `from` ----
  auto assemble = [&](const NegStream& from, Batch& bt, bool& more) {
    HostTimer ht(ctx, "mine_host_assemble");
    // stream windows this batch should cover: enough to finish the stage at the observed acceptance rate.  Only the windows that
    // have to be evaluated (not answered from the cache) are bounded by max_batch.
    double want = 65536;
    if (accept_rate > 0) want = std::max(4096.0, 1.25 * (count - got) / accept_rate + 1024);
    if (!cache_on) want = std::min(want, (double)max_batch);
    want = std::min(want, 4.0e12);
    NegStream cur = from;
    more = true;
    if (cache_on && !m->cache.map.empty()) {
      const size_t guess = (size_t)std::min<double>((double)SEG_CAP, want / 8000.0 + 64);
      bt.pyrs.reserve(guess); bt.segs.reserve(guess);
    }
    int open_pyr = -1;                  // bt.miss entry of the pyramid being assembled from its first window
    while ((double)bt.nwin < want && bt.neval < max_batch && more && (long long)bt.cand_eval.size() < CAND_CAP && bt.segs.size() < SEG_CAP) {
      if (cache_on && ((cur.fresh && cur.pointX == cur.offX && cur.pointY == cur.offY) || cur.is_pyramid_start())) {
        const uint64_t key = cur.pyramid_key();
        const PyrEntry* ep = m->cache.lookup(cur.src, cur.offX, cur.offY, (int)m->bgs.size());
        if (ep && ep->n_stages <= nstages) {
          const PyrEntry& e = *ep;
          if (e.surv.empty()) {
            if (bt.segs.empty() || bt.segs.back().kind != 2) {
              EmptyRun r;
              r.start = cur; r.base = bt.nwin;
              bt.segs.push_back(Seg{bt.nwin, 0, 2, (int)bt.runs.size()});
              bt.runs.push_back(r);
            }
            EmptyRun& r = bt.runs.back();
            r.nwin += e.nwin; r.count += 1;
            bt.segs.back().nwin = r.nwin;
            bt.nwin += e.nwin;
            more = cur.nextImg();
            r.after = cur;
            open_pyr = -1;
            continue;
          }
          if (e.n_stages < nstages) {
            // A stale entry is brought up to date when the batch that re-evaluates it has finished.  The stream returns to the
            // same pyramids in the same order, so a batch that ran on past its first stale pyramid would re-evaluate the same
            // stale lists once per list cycle it spans: end the batch there instead -- the next one finds the entries fresh.
            if (bt.has_stale && key == bt.first_stale_key) break;
            if (!bt.has_stale) { bt.has_stale = true; bt.first_stale_key = key; }
          }
          CachedPyr cp;
          cp.start = cur; cp.base = bt.nwin; cp.nwin = e.nwin; cp.key = key; cp.cand0 = bt.cand.size(); cp.ncand = e.surv.size();
          for (const uint32_t o : e.surv) {
            if (e.n_stages < nstages) bt.cand_eval.push_back((uint32_t)bt.cand.size());
            bt.cand.push_back(cp.base + (long long)o);
            bt.cand_fs.push_back((uint8_t)std::min(e.n_stages, 255));
          }
          bt.segs.push_back(Seg{cp.base, cp.nwin, 1, (int)bt.pyrs.size()});
          bt.pyrs.push_back(cp);
          bt.nwin += cp.nwin;
          more = cur.nextImg();
          open_pyr = -1;
          continue;
        }
        MissPyr mp;
        mp.key = key; mp.base = bt.nwin;
        open_pyr = (int)bt.miss.size();
        bt.miss.push_back(mp);
      }
      LevelDesc L;
      fill_level(cur, L);
      const int ix0 = (cur.pointX - cur.offX) / std::max(1, L.step_x), iy0 = (cur.pointY - cur.offY) / std::max(1, L.step_y);
      L.w_skip = (long long)iy0 * L.nx + ix0;
      const long long avail = (long long)L.nx * L.ny - L.w_skip;
      L.w_count = std::min(avail, max_batch - bt.neval);
      L.w_base = bt.neval;                               // evaluation ordinal of its first window
      L.px_off = (long long)bt.pool_px;
      bt.pool_px += (size_t)L.lw * L.lh;
      bt.segs.push_back(Seg{bt.nwin, L.w_count, 0, (int)bt.levels.size()});
      bt.lv_stream.push_back(bt.nwin);
      bt.neval += L.w_count;
      bt.nwin += L.w_count;
      if (open_pyr >= 0) bt.miss[open_pyr].nwin += L.w_count;
      NegStream nxt = cur;
      if (L.w_count == avail) {
        more = nxt.nextLevel();
        if (!more || nxt.fresh) {                        // that was the pyramid's last level
          if (open_pyr >= 0) bt.miss[open_pyr].complete = true;
          open_pyr = -1;
        }
      } else {
        const long long k = L.w_skip + L.w_count;      // first window not taken
        nxt.pointY = cur.offY + (int)(k / L.nx) * L.step_y;
        nxt.pointX = cur.offX + (int)(k % L.nx) * L.step_x;
      }
      bt.levels.push_back(L);
      bt.before.push_back(cur);
      bt.after.push_back(nxt);
      cur = nxt;
      if (L.w_count < avail) break;
    }
  };
  auto batch_end_state = [&](const Batch& bt) -> NegStream {
    const Seg& sg = bt.segs.back();
    if (sg.kind == 0) return bt.after[sg.idx];
    if (sg.kind == 2) return bt.runs[sg.idx].after;
    NegStream a = bt.pyrs[sg.idx].start;
    a.nextImg();
    return a;
  };
  // asynchronous half of a plane-path batch: pyramid planes + cascade for the level parts, per-window path for the cached survivors
  auto launch_batch = [&](Batch& bt, int slot) -> int {
    HostTimer he(ctx, "mine_host_eval");
    if (!m->ev_done[slot]) CC_CUDA(cudaEventCreateWithFlags(&m->ev_done[slot], cudaEventDisableTiming));
    { HostTimer h1(ctx, "mine_host_eval_levels"); CC_CHECK(launch_levels(ctx, m, bt, slot)); }
    { HostTimer h2(ctx, "mine_host_eval_cand"); CC_CHECK(launch_candidates(ctx, m, bt, slot)); }
    CC_CUDA(cudaEventRecord(m->ev_done[slot], ctx->stream));
    m->stat_built += bt.neval;
    return CC_OK;
  };
  Batch pre;
  bool pre_valid = false, pre_more = false;
  int slot = 0;

  while (got < count && stream_ok) {
    // serial rule, checked before every window: stop when (getcount+1)/consumed <= minRatio
    if (consumed != 0 && ((double)got + 1) / (double)consumed <= min_ratio) break;
    Batch bt;
    bool more = true;
    if (pre_valid) {
      bt = std::move(pre); pre = Batch(); more = pre_more; pre_valid = false; slot ^= 1;
    } else {
      assemble(ns, bt, more);
      if (pipelined) CC_CHECK(launch_batch(bt, slot));
    }
    if (bt.segs.empty()) { stream_ok = more; continue; }
    const long long nwin = bt.nwin;
    // ---- GPU: levels, gather, integrals, cascade ----
    const bool direct = m->stages.empty();       // stage 0: every window passes, gather straight into the training slots
    long long take_direct = 0;
    HogShard hshard;
    if (!level_path) {
      HostTimer halloc(ctx, "mine_host_alloc");
      CC_CHECK(ensure_scratch(ctx, m, (direct || hog_win) ? 1 : (int)nwin, bt.pool_px, (int)bt.levels.size(), direct ? 1 : nwin,
                              hog_win ? (double)max_batch / (double)std::max<long long>(nwin, 1) : 1.0));
      halloc.stop();
      if (hog_win) hshard = hog_shard(ctx, m, bt.levels, nwin);
      CC_CUDA(cudaMemcpyAsync(m->d_levels, bt.levels.data(), sizeof(LevelDesc) * bt.levels.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!level_path) {
      ScopedTimer tm(ctx, "mine", hog_win ? 1 : 2);
      if (hog_win) { if (hshard.tiles > 0) k_resize_tiles<<<std::min(hshard.tiles, ctx->sm_count * 64), RZ_THREADS, 0, ctx->stream>>>(m->d_levels + hshard.l0, hshard.nl, hshard.tiles, m->d_pool); }
      else k_resize_levels<<<grid_for(ctx, (long long)bt.pool_px, 256), 256, 0, ctx->stream>>>(m->d_levels, (int)bt.levels.size(), (long long)bt.pool_px, m->d_pool);
      CC_CUDA(cudaGetLastError());
      if (direct) {
        // how many windows does the serial rule consume?  all pass, so got grows with consumed
        take_direct = std::min<long long>(nwin, count - got);
        k_gather_windows<<<grid_for(ctx, take_direct * npx, 256), 256, 0, ctx->stream>>>(m->d_levels, (int)bt.levels.size(), m->d_pool, W, H, take_direct,
                                                                                       ctx->d_imgs + (size_t)(first_slot + got) * npx);
      } else if (!hog_win) {
        k_gather_windows<<<grid_for(ctx, nwin * npx, 256), 256, 0, ctx->stream>>>(m->d_levels, (int)bt.levels.size(), m->d_pool, W, H, nwin, m->scratch.d_imgs);
      }
      CC_CUDA(cudaGetLastError());
      m->stat_built += direct ? 0 : nwin;
    }
    long long stop_after = -1;     // batch ordinal of the last consumed window (-1: none consumed)
    if (direct) {
      // with no stages every window passes; the ratio rule can only trigger if minRatio >= 1
      long long taken = 0;
      while (taken < take_direct) {
        if (consumed != 0 && ((double)got + 1) / (double)consumed <= min_ratio) break;
        consumed++; got++; taken++;
      }
      CC_CUDA(cudaMemsetAsync(ctx->d_labels + first_slot + got - taken, 0, (size_t)taken, ctx->stream));
      ctx->precalc_valid = false;
      CC_CHECK(launch_integrals(ctx, main_store(ctx), first_slot + got - (int)taken, (int)taken));
      CC_CUDA(cudaStreamSynchronize(ctx->stream));
      stop_after = taken - 1;
      if (taken < take_direct) { /* ratio rule fired */ }
      if (got < count && taken == nwin) { ns = batch_end_state(bt); stream_ok = more; continue; }
    } else {
      if (pipelined) {
        // (no speculation while this batch is bringing stale cache entries up to date: the next batch would re-evaluate the same lists)
        if (more && accept_rate > 0 && accept_rate * (double)bt.nwin < 0.5 * (double)(count - got) && !bt.has_stale) {
          assemble(batch_end_state(bt), pre, pre_more);
          if (!pre.segs.empty()) {
            CC_CHECK(launch_batch(pre, slot ^ 1));
            pre_valid = true;
          }
        }
        {
          HostTimer hw(ctx, "mine_host_wait");
          CC_CHECK(finish_levels(ctx, m, slot, bt.neval, survivors));
        }
        HostTimer hm(ctx, "mine_host_merge");
        merge_and_update_cache(m, bt, survivors, m->h_cand_flags[slot], nstages, cache_on, surv64);
      } else {
        HostTimer he(ctx, "mine_host_eval");
        if (tile_path) CC_CHECK(eval_tiles(ctx, m, bt, tgeom, survivors));
        else if (hog_win) CC_CHECK(eval_hog_windows(ctx, m, (int)bt.levels.size(), nwin, hshard, survivors));
        else CC_CHECK(eval_scratch(ctx, m, (int)nwin, survivors));
        if (tile_path) m->stat_built += nwin;
        surv64.assign(survivors.begin(), survivors.end());
      }
      HostTimer hc(ctx, "mine_host_cut");
      // ---- serial stopping rule over the ordered survivors ----
      bool stopped = false;
      const long long pos = serial_cut<long long>(surv64.data(), surv64.size(), nwin, count, min_ratio, consumed, got, accepted, stopped);
      stop_after = pos - 1;
      const int first_acc = first_slot + got - (int)accepted.size();
      if (level_path || hog_win) CC_CHECK(gather_accepted(ctx, m, bt, accepted, first_acc, 0));      // (the levels of the batch, from the source images)
      else {
        std::vector<unsigned> acc32(accepted.begin(), accepted.end());
        CC_CHECK(accept_windows(ctx, m, acc32.data(), (int)acc32.size(), first_acc, 0));
      }
      if (nwin > 0) accept_rate = std::max((double)surv64.size() / (double)nwin, 1e-12);
      for (const CachedPyr& cp : bt.pyrs) if (cp.base <= stop_after) m->stat_skipped += std::min(cp.nwin, stop_after + 1 - cp.base);
      for (const EmptyRun& r : bt.runs) if (r.base <= stop_after) m->stat_skipped += std::min(r.nwin, stop_after + 1 - r.base);
      if (!stopped) { ns = batch_end_state(bt); stream_ok = more; continue; }
    }
    // ---- stopped inside this batch: reader state right after window `stop_after` ----
    if (stop_after < 0) {
      // nothing consumed from this batch: state unchanged
    } else {
      // find the segment containing stop_after
      size_t si = 0;
      {
        size_t lo = 0, hi = bt.segs.size() - 1;
        while (lo < hi) { const size_t mid = (lo + hi + 1) >> 1; if (bt.segs[mid].base <= stop_after) lo = mid; else hi = mid - 1; }
        si = lo;
      }
      const Seg& sg = bt.segs[si];
      if (sg.kind == 0) {
        const LevelDesc& L = bt.levels[sg.idx];
        const long long k = stop_after - sg.base + L.w_skip + 1;          // index (within the level) of the next window
        if (k >= (long long)L.nx * L.ny) {
          ns = bt.after[sg.idx];                                           // next level / image (already advanced)
          if (si + 1 == bt.segs.size()) stream_ok = more;
        } else {
          // same level: the state of that level, then position the point
          NegStream st = bt.before[sg.idx];
          st.pointY = st.offY + (int)(k / L.nx) * L.step_y;
          st.pointX = st.offX + (int)(k % L.nx) * L.step_x;
          ns = st;
        }
      } else if (sg.kind == 2) {
        // inside a run of survivor-free cached pyramids: walk to the pyramid that holds the stop position, then to its level
        const EmptyRun& r = bt.runs[sg.idx];
        long long k = stop_after - r.base + 1;                             // windows of the run consumed
        if (k >= r.nwin) {
          ns = r.after;
          if (si + 1 == bt.segs.size()) stream_ok = more;
        } else {
          NegStream st = r.start;
          for (;;) {
            const PyrEntry* e = m->cache.lookup(st.src, st.offX, st.offY, (int)m->bgs.size());
            CC_REQUIRE(e, CC_ERR_STATE, "mining: cached pyramid vanished while a batch was in flight");
            if (k < e->nwin) break;
            k -= e->nwin;
            CC_REQUIRE(st.nextImg(), CC_ERR_STATE, "mining: stream ended inside a cached run");
          }
          long long lb = 0, ln = (long long)st.countX() * st.countY();
          while (k >= lb + ln) { lb += ln; level_after(st); ln = (long long)st.countX() * st.countY(); }
          const int nx = st.countX();
          const long long kk = k - lb;
          st.pointY = st.offY + (int)(kk / nx) * st.stepY();
          st.pointX = st.offX + (int)(kk % nx) * st.stepX();
          ns = st;
        }
      } else {
        const CachedPyr& cp = bt.pyrs[sg.idx];
        const long long k = stop_after - cp.base + 1;                      // windows of the pyramid consumed
        if (k >= cp.nwin) {
          ns = cp.start;
          const bool has_next = ns.nextImg();
          if (si + 1 == bt.segs.size()) stream_ok = more; else (void)has_next;
        } else {
          NegStream st = cp.start;
          long long lb = 0, ln = (long long)st.countX() * st.countY();
          while (k >= lb + ln) { lb += ln; level_after(st); ln = (long long)st.countX() * st.countY(); }
          const int nx = st.countX();
          const long long kk = k - lb;
          st.pointY = st.offY + (int)(kk / nx) * st.stepY();
          st.pointX = st.offX + (int)(kk % nx) * st.stepX();
          ns = st;
        }
      }
    }
    break;
  }
  CC_CUDA(cudaStreamSynchronize(ctx->stream));          // gathers / integrals of the last batches; a dropped speculative batch
  state_from_stream(ns, state);
  m->stat_consumed += consumed;
  *consumed_out = consumed;
  *got_out = got;
  return CC_OK;
}

int cc_mine_positives(cc_ctx* ctx, const uint8_t* host_vec, int vec_count, int count, int32_t* vec_pos, int64_t* consumed_out, int32_t* got_out) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_mine_positives: evaluator not initialised");
  CC_REQUIRE(host_vec && vec_pos && consumed_out && got_out, CC_ERR_BAD_ARG, "cc_mine_positives: null argument");
  CC_REQUIRE(count >= 0 && count <= ctx->max_samples, CC_ERR_BAD_ARG, "cc_mine_positives: count %d exceeds maxSampleCount %d", count, ctx->max_samples);
  MineState* m = mine_state(ctx);
  CC_CUDA(cudaSetDevice(ctx->device));
  const int npx = ctx->W * ctx->H;
  long long consumed = 0;
  int got = 0;
  int pos = *vec_pos;
  const int max_batch = (int)std::min<long long>(65536, slot_path_max_batch(ctx));      // positives: the .vec block is small
  std::vector<unsigned> survivors, accepted;
  while (got < count) {
    if (pos >= vec_count) {
      *vec_pos = pos; *consumed_out = consumed; *got_out = got;
      set_error("Can not get new positive sample. The most possible reason is insufficient count of samples in given vec-file.");
      return CC_ERR_BAD_ARG;
    }
    const int nwin = std::min(std::min(max_batch, vec_count - pos), std::max(1024, (count - got) + (count - got) / 8 + 64));
    CC_CHECK(ensure_scratch(ctx, m, nwin, 1, 1));
    CC_CUDA(cudaMemcpyAsync(m->scratch.d_imgs, host_vec + (size_t)pos * npx, (size_t)nwin * npx, cudaMemcpyHostToDevice, ctx->stream));
    CC_CHECK(eval_scratch(ctx, m, nwin, survivors));
    accepted.clear();
    long long last_consumed = -1;
    for (size_t si = 0; si < survivors.size() && got < count; ++si) {
      accepted.push_back(survivors[si]);
      got++;
      last_consumed = survivors[si];
    }
    if (got < count) last_consumed = nwin - 1;          // everything in the batch was consumed
    consumed += last_consumed + 1;
    pos += (int)(last_consumed + 1);
    CC_CHECK(accept_windows(ctx, m, accepted.data(), (int)accepted.size(), got - (int)accepted.size(), 1));
  }
  *vec_pos = pos; *consumed_out = consumed; *got_out = got;
  return CC_OK;
}

// Host-only: advance the NegReader state machine by `nwindows` windows over images of the given sizes
// (no pixels, no GPU).  Used by the CPU tests to pin the stream enumeration against the oracle, and by
// callers that need to skip ahead.  window_xywh (optional) receives x, y, level_w, level_h and the source
// image index of the last consumed window.
int cc_reader_advance(const int32_t* img_w, const int32_t* img_h, int nimg, int win_w, int win_h, cc_reader_state* state,
                      int64_t nwindows, int32_t* window_xywh) {
  CC_REQUIRE(img_w && img_h && nimg > 0 && state && nwindows >= 0, CC_ERR_BAD_ARG, "cc_reader_advance: bad arguments");
  std::vector<BgImage> imgs(nimg);
  for (int i = 0; i < nimg; ++i) { imgs[i].w = img_w[i]; imgs[i].h = img_h[i]; }
  NegStream ns;
  ns.bgs = &imgs; ns.W = win_w; ns.H = win_h;
  stream_from_state(ns, *state, nimg);
  bool ok = true;
  if (ns.lw == 0 && nwindows > 0) ok = ns.nextImg();
  int64_t left = nwindows;
  while (left > 0 && ok) {
    const int nx = ns.countX(), ny = ns.countY();
    const int sx = std::max(1, ns.stepX()), sy = std::max(1, ns.stepY());
    const int64_t k0 = (int64_t)((ns.pointY - ns.offY) / sy) * nx + (ns.pointX - ns.offX) / sx;
    const int64_t avail = (int64_t)nx * ny - k0;
    const int64_t take = std::min(avail, left);
    const int64_t klast = k0 + take - 1;
    if (window_xywh) {
      window_xywh[0] = ns.offX + (int)(klast % nx) * sx; window_xywh[1] = ns.offY + (int)(klast / nx) * sy;
      window_xywh[2] = ns.lw; window_xywh[3] = ns.lh; window_xywh[4] = ns.src;
    }
    left -= take;
    if (take == avail) ok = ns.nextLevel();
    else { const int64_t k = k0 + take; ns.pointY = ns.offY + (int)(k / nx) * sy; ns.pointX = ns.offX + (int)(k % nx) * sx; }
  }
  state_from_stream(ns, state);
  CC_REQUIRE(left == 0, CC_ERR_BAD_ARG, "cc_reader_advance: no usable background image (stream ended with %lld windows left)", (long long)left);
  return CC_OK;
}

int cc_mine_serial_cut(const uint32_t* survivors, int64_t nsurv, int64_t nwin, int32_t count, double min_ratio, int64_t* consumed, int32_t* got,
                       int64_t* next_window, int32_t* stopped, uint32_t* accepted, int32_t* n_accepted) {
  CC_REQUIRE((survivors || nsurv == 0) && nsurv >= 0 && nwin >= 0 && consumed && got && next_window && stopped && n_accepted, CC_ERR_BAD_ARG,
             "cc_mine_serial_cut: bad arguments");
  std::vector<unsigned> acc;
  long long c = *consumed;
  int g = *got;
  bool st = false;
  const long long pos = serial_cut<unsigned>(survivors, (size_t)nsurv, nwin, count, min_ratio, c, g, acc, st);
  *consumed = c; *got = g; *next_window = pos; *stopped = st ? 1 : 0;
  *n_accepted = (int32_t)acc.size();
  if (accepted) std::memcpy(accepted, acc.data(), acc.size() * sizeof(unsigned));
  return CC_OK;
}

int cc_mine_scan(cc_ctx* ctx, int img0, int nimg, int64_t* windows, int64_t* survivors_out) {
  return cc_mine_scan_digest(ctx, img0, nimg, windows, survivors_out, nullptr);
}

int cc_mine_scan_digest(cc_ctx* ctx, int img0, int nimg, int64_t* windows, int64_t* survivors_out, uint64_t* digest_out) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_mine_scan: evaluator not initialised");
  MineState* m = mine_state(ctx);
  CC_REQUIRE(img0 >= 0 && nimg > 0 && img0 + nimg <= (int)m->bgs.size(), CC_ERR_BAD_ARG, "cc_mine_scan: image range out of bounds");
  CC_CUDA(cudaSetDevice(ctx->device));
  const int W = ctx->W, H = ctx->H, npx = W * H;
  const bool level_path = use_level_path(ctx, m);
  const TileGeom tgeom = tile_geometry(ctx, ctx->ftype == CC_FEATURE_HAAR, ctx->ftype == CC_FEATURE_HAAR && m->cascade_tilted);
  const bool tile_path = level_path && m->mode == 3 && tgeom.ok;
  const bool hog_win = !level_path && use_hog_window_path(ctx, m);
  const long long max_batch = level_path ? level_path_max_batch(ctx, m) : hog_win ? HOG_WINDOW_MAX_BATCH : slot_path_max_batch(ctx);
  long long total = 0, surv = 0;
  uint64_t digest = 0;                 // order-independent digest of the surviving windows' scan ordinals
  std::vector<unsigned> survivors;
  auto account = [&](long long base) {
    surv += (long long)survivors.size();
    for (const unsigned o : survivors) {
      uint64_t z = (uint64_t)(base + (long long)o) + 0x9E3779B97F4A7C15ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
      digest += z;
    }
  };
  // every image from its first level (round 0 offsets), independent of the reader state
  std::vector<BgImage> one(1);
  Batch bt;
  int slot = 0;
  long long pend_nwin = 0, pend_base = 0;      // windows / first scan ordinal of the batch in flight in slot ^ 1
  auto flush = [&]() -> int {
    if (bt.nwin == 0) return CC_OK;
    const int nwin = (int)bt.nwin;
    if (level_path && !tile_path) {
      // two batches in flight, as in cc_mine_negatives: launch this one, then collect the previous one
      CC_CHECK(launch_levels(ctx, m, bt, slot));
      if (pend_nwin) { CC_CHECK(finish_levels(ctx, m, slot ^ 1, pend_nwin, survivors)); account(pend_base); }
      pend_nwin = nwin; pend_base = total; slot ^= 1;
      total += nwin;
      bt = Batch();
      return CC_OK;
    }
    if (level_path) {
      CC_CHECK(eval_tiles(ctx, m, bt, tgeom, survivors));
      account(total); total += nwin;
      bt = Batch();
      return CC_OK;
    }
    CC_CHECK(ensure_scratch(ctx, m, hog_win ? 1 : nwin, bt.pool_px, (int)bt.levels.size(), nwin));
    HogShard hshard;
    if (hog_win) hshard = hog_shard(ctx, m, bt.levels, nwin);
    CC_CUDA(cudaMemcpyAsync(m->d_levels, bt.levels.data(), sizeof(LevelDesc) * bt.levels.size(), cudaMemcpyHostToDevice, ctx->stream));
    {
      ScopedTimer tm(ctx, "mine", hog_win ? 1 : 2);
      if (hog_win) { if (hshard.tiles > 0) k_resize_tiles<<<std::min(hshard.tiles, ctx->sm_count * 64), RZ_THREADS, 0, ctx->stream>>>(m->d_levels + hshard.l0, hshard.nl, hshard.tiles, m->d_pool); }
      else k_resize_levels<<<grid_for(ctx, (long long)bt.pool_px, 256), 256, 0, ctx->stream>>>(m->d_levels, (int)bt.levels.size(), (long long)bt.pool_px, m->d_pool);
      if (!hog_win)
        k_gather_windows<<<grid_for(ctx, (long long)nwin * npx, 256), 256, 0, ctx->stream>>>(m->d_levels, (int)bt.levels.size(), m->d_pool, W, H, nwin, m->scratch.d_imgs);
      CC_CUDA(cudaGetLastError());
    }
    if (hog_win) CC_CHECK(eval_hog_windows(ctx, m, (int)bt.levels.size(), nwin, hshard, survivors));
    else CC_CHECK(eval_scratch(ctx, m, nwin, survivors));
    account(total); total += nwin;
    bt = Batch();
    return CC_OK;
  };
  for (int i = img0; i < img0 + nimg; ++i) {
    one[0] = m->bgs[i];
    NegStream s;
    s.bgs = &one; s.W = W; s.H = H;
    if (!s.nextImg()) continue;
    bool more = true;
    while (more) {
      LevelDesc L;
      fill_level(s, L);
      long long done = 0;
      const long long all = (long long)L.nx * L.ny;
      while (done < all) {
        if (bt.nwin >= max_batch) CC_CHECK(flush());
        LevelDesc part = L;
        part.w_skip = done;
        part.w_count = std::min(all - done, max_batch - bt.nwin);
        part.w_base = bt.nwin;
        part.px_off = (long long)bt.pool_px;
        bt.pool_px += (size_t)L.lw * L.lh;
        bt.nwin += part.w_count;
        bt.neval += part.w_count;
        bt.levels.push_back(part);
        done += part.w_count;
      }
      // next level of the same image only
      s.scale *= s.scaleFactor;
      if (s.scale <= 1.0F) { s.lw = (int)(s.scale * one[0].w); s.lh = (int)(s.scale * one[0].h); } else more = false;
    }
  }
  CC_CHECK(flush());
  if (pend_nwin) { CC_CHECK(finish_levels(ctx, m, slot ^ 1, pend_nwin, survivors)); account(pend_base); }
  if (windows) *windows = total;
  if (survivors_out) *survivors_out = surv;
  if (digest_out) *digest_out = digest;
  return CC_OK;
}

}  // extern "C"

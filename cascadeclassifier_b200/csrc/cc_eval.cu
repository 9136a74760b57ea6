// K1: batched integral images (sum / squared / tilted) and HOG integral histograms, plus the
// evaluator half of the C ABI (init / setImage / operator()).
//
// Device layout.  Haar/LBP integrals are stored as tiles of 32 samples, [tile][P][32] int32 with the
// sample index fastest, so that (a) the precalc kernel pulls one tile into shared memory with a single
// bulk copy and reads it bank-conflict-free with lane == sample, and (b) any per-sample gather is a
// coalesced 128-byte line.  HOG planes are [plane][P][ns_pad] float, also sample-fastest.
#include <algorithm>
#include <cstring>

#include "cc_internal.h"

namespace cc {

// ------------------------------------------------------------------------------------------------
// Haar / LBP integral kernel: one CTA per tile of 32 samples, one warp per sample.
//   sum   : row prefix by warp-shuffle inclusive scan (32 columns per step, carry between chunks),
//           column accumulation in a per-lane register.
//   nf    : sums of I and I^2 over the inner (1,1,W-2,H-2) rectangle, warp-reduced, one f64 sqrt
//           (calcNormFactor, src/features.cpp:13-25).  The f64 sqsum integral of the reference holds
//           exact integers, so the direct sum is identical.
//   tilted: T(Y,X) = T(Y-1,X) + D1(Y-1,X-1) + D2(Y-1,X-1) - I(Y-1,X-1) with the two diagonal running
//           sums D1(y,x)=D1(y-1,x-1)+I(y,x), D2(y,x)=D2(y-1,x+1)+I(y,x) (cv::integral tilted definition).
// Results are staged in shared memory as [P][33] and written out as full 128-byte lines.
// ------------------------------------------------------------------------------------------------
template <int NCHUNK>
__global__ void __launch_bounds__(1024, 1)
k_integral_tiles(const uint8_t* __restrict__ imgs, int W, int H, int first, int last /*exclusive*/,
                 int32_t* __restrict__ sum_tiles, int32_t* __restrict__ tilted_tiles, float* __restrict__ nf_out,
                 int want_tilted, int want_nf, int staged) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int S = W + 1, P = S * (H + 1);
  // staged: results go to smem [P][33] first and leave as full 128-byte lines; otherwise (windows whose
  // staging buffer does not fit) each warp writes its sample column of the tile directly.
  int32_t* stage = reinterpret_cast<int32_t*>(smem_raw);
  uint8_t* simg = smem_raw + (staged ? (size_t)P * 33 * sizeof(int32_t) : 0);           // [32][H*W]
  int32_t* diag = reinterpret_cast<int32_t*>(simg + (((size_t)32 * W * H + 15) & ~(size_t)15));   // [32][2][W+2]

  const int tile = first / TILE + blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sample = tile * TILE + warp;
  const bool active = sample >= first && sample < last;
  const int npx = W * H;
  const int lo = max(first - tile * TILE, 0), hi = min(last - tile * TILE, TILE);

  for (int i = threadIdx.x; i < TILE * npx; i += blockDim.x) {
    int s = i / npx;
    int g = tile * TILE + s;
    simg[i] = (g >= first && g < last) ? imgs[(size_t)g * npx + (i - s * npx)] : 0;
  }
  __syncthreads();

  const uint8_t* img = simg + warp * npx;
  const bool compute = staged || active;
  // ---------------- sum plane + norm factor ----------------
  if (compute) {
    int32_t* out = staged ? stage + warp : sum_tiles + (size_t)tile * P * TILE + warp;
    const int ostride = staged ? 33 : TILE;
    int32_t colacc[NCHUNK];
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) colacc[c] = 0;
    for (int x = lane; x <= W; x += 32) out[x * ostride] = 0;              // first row
    long long s1 = 0, s2 = 0;
    for (int y = 0; y < H; ++y) {
      int carry = 0;
      if (lane == 0) out[((y + 1) * S) * ostride] = 0;                     // first column
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const int x = c * 32 + lane;
        int v = (x < W) ? img[y * W + x] : 0;
        if (x >= 1 && x <= W - 2 && y >= 1 && y <= H - 2) { s1 += v; s2 += v * v; }
        int pre = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, pre, d);
          if (lane >= d) pre += t;
        }
        pre += carry;
        carry = __shfl_sync(0xffffffffu, pre, 31);
        colacc[c] += pre;
        if (x < W) out[((y + 1) * S + x + 1) * ostride] = colacc[c];
      }
    }
    if (want_nf) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
      }
      if (lane == 0 && active) {
        double area = (double)((W - 2) * (H - 2));
        double vs = (double)s1;
        nf_out[sample] = (float)sqrt(area * (double)s2 - vs * vs);
      }
    }
  }
  if (staged) {
    __syncthreads();
    int32_t* dst = sum_tiles + (size_t)tile * P * TILE;
    for (int i = threadIdx.x; i < P * TILE; i += blockDim.x) {
      int p = i >> 5, s = i & 31;
      if (s >= lo && s < hi) dst[i] = stage[p * 33 + s];
    }
  }
  if (!want_tilted) return;
  __syncthreads();
  // ---------------- tilted plane ----------------
  if (compute) {
    int32_t* out = staged ? stage + warp : tilted_tiles + (size_t)tile * P * TILE + warp;
    const int ostride = staged ? 33 : TILE;
    int32_t* d1 = diag + warp * 2 * (W + 2);     // index j = x + 1, x in [-1, W-1]
    int32_t* d2 = d1 + (W + 2);
    for (int j = lane; j < W + 2; j += 32) { d1[j] = 0; d2[j] = 0; }
    for (int X = lane; X <= W; X += 32) out[X * ostride] = 0;               // T(0, X) = 0
    __syncwarp();
    for (int y = 0; y < H; ++y) {
      int32_t nd1[NCHUNK + 1], nd2[NCHUNK + 1];
#pragma unroll
      for (int c = 0; c <= NCHUNK; ++c) {
        const int j = c * 32 + lane;           // column X = j, pixel x = j - 1
        nd1[c] = nd2[c] = 0;
        if (j <= W) {
          const int x = j - 1;
          const int iv = (x >= 0) ? img[y * W + x] : 0;
          const int a = (x >= 1) ? d1[j - 1] : 0;            // D1(y-1, x-1)
          const int b = (x + 1 <= W - 1) ? d2[j + 1] : 0;    // D2(y-1, x+1)
          nd1[c] = a + iv;
          nd2[c] = b + iv;
          const int prev = out[(y * S + j) * ostride];
          out[((y + 1) * S + j) * ostride] = prev + nd1[c] + nd2[c] - iv;
        }
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c <= NCHUNK; ++c) {
        const int j = c * 32 + lane;
        if (j <= W) { d1[j] = nd1[c]; d2[j] = nd2[c]; }
      }
      __syncwarp();
    }
  }
  if (staged) {
    __syncthreads();
    int32_t* dst = tilted_tiles + (size_t)tile * P * TILE;
    for (int i = threadIdx.x; i < P * TILE; i += blockDim.x) {
      int p = i >> 5, s = i & 31;
      if (s >= lo && s < hi) dst[i] = stage[p * 33 + s];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// HOG integral histograms (CvHOGEvaluator::integralHistogram, src/HOGfeatures.cpp:163-256)
// ------------------------------------------------------------------------------------------------
// gradient magnitude + orientation bin per pixel; outputs are [pixel][ns_pad] (sample-fastest)
// Gradient magnitude + orientation bin of every pixel of every window (src/HOGfeatures.cpp:163-224: central differences with
// BORDER_REPLICATE at the crop border, cartToPolar, bin = floor(angle * 9 / pi - 0.5) wrapped).  Windows are stored sample-major
// (u8 [sample][H][W]) while the outputs are sample-fastest, so a CTA stages a block of rows (+ one halo row on either side) of 32
// windows in shared memory with coalesced reads -- a warp reads the contiguous bytes of one window at a time -- and computes with
// lane == sample: shared reads are conflict-free (window stride = 1 mod 32 words), global writes are 128-byte lines.  (One thread
// per (pixel, sample) reading the windows directly costs four 32-byte sectors per thread: it was 85 % of the HOG K1 time.)
constexpr int HOG_G_THREADS = 256;
constexpr int HOG_G_ROWS = 8;             // image rows per staged block
__global__ void __launch_bounds__(HOG_G_THREADS)
k_hog_gradient(const uint8_t* __restrict__ imgs, int W, int H, int first, int n, int ns_pad, int wstride /* bytes per staged window, = 4 mod 128 */,
               float* __restrict__ grad, uint8_t* __restrict__ qangle) {
  extern __shared__ __align__(16) uint8_t s_px[];          // [32][wstride]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = HOG_G_THREADS / 32;
  for (int s0 = blockIdx.x * 32; s0 < n; s0 += gridDim.x * 32) {
    const int ns = min(32, n - s0);
    for (int y0 = 0; y0 < H; y0 += HOG_G_ROWS) {
      const int r_lo = max(y0 - 1, 0), r_hi = min(y0 + HOG_G_ROWS + 1, H);      // staged rows [r_lo, r_hi)
      const int nbytes = (r_hi - r_lo) * W;
      __syncthreads();
      for (int k = warp; k < ns; k += nwarps) {
        const uint8_t* src = imgs + (size_t)(first + s0 + k) * W * H + (size_t)r_lo * W;
        uint8_t* dst = s_px + (size_t)k * wstride;
        for (int i = lane; i < nbytes; i += 32) dst[i] = src[i];
      }
      __syncthreads();
      const int rows = min(HOG_G_ROWS, H - y0);
      if (lane < ns) {
        const uint8_t* img = s_px + (size_t)lane * wstride;
        const int s = first + s0 + lane;
        for (int p = warp; p < rows * W; p += nwarps) {
          const int yy = p / W, x = p - yy * W, y = y0 + yy;
          const int xm = max(x - 1, 0), xp = min(x + 1, W - 1), ym = max(y - 1, 0), yp = min(y + 1, H - 1);
          const float dx = (float)((int)img[(y - r_lo) * W + xp] - (int)img[(y - r_lo) * W + xm]);
          const float dy = (float)((int)img[(yp - r_lo) * W + x] - (int)img[(ym - r_lo) * W + x]);
          float mag;
          int b;
          hog_mag_bin(dx, dy, mag, b);
          const size_t o = (size_t)(y * W + x) * ns_pad + s;
          grad[o] = mag;
          qangle[o] = (uint8_t)b;
        }
      }
    }
  }
}

// Integral histograms (src/HOGfeatures.cpp:226-255: per row a sequential prefix strSum along x, plane[y+1][x+1] = plane[y][x+1] +
// strSum -- the same f32 additions in the same order), kept COMPACT: every cell corner of the catalog is a multiple of HOG_GRID = 4
// pixels (block step 4, cell sizes multiples of 8), so only plane entries (4i, 4j) are ever read and only they are stored --
// 325 instead of 4753 floats per plane for a 48x96 window.  A thread owns one plane of one sample: it carries the running
// integral of the kept columns (I(y, 4c) for the current y) in a small local array, adds the row prefix when it passes a kept
// column and writes the array out at every 4th row.  A warp = 32 consecutive samples of one plane (coalesced: the planes are
// sample-fastest), a CTA = the 10 planes of those samples, so the nine re-reads of a gradient line come from L1.
constexpr int HOG_INT_THREADS = 320;
constexpr int HOG_MAX_COLS = 64;          // kept columns: windows up to 252 pixels wide
__global__ void __launch_bounds__(HOG_INT_THREADS)
k_hog_integrals(const float* __restrict__ grad, const uint8_t* __restrict__ qangle, int W, int H, int first, int n, int ns_pad,
                float* __restrict__ planes) {
  const int Sc = W / HOG_GRID + 1, Pc = Sc * (H / HOG_GRID + 1);
  const int lane = threadIdx.x & 31, b = threadIdx.x >> 5;
  for (int s0 = blockIdx.x * 32; s0 < n; s0 += gridDim.x * 32) {
    const int si = s0 + lane;
    if (si >= n) continue;
    const int s = first + si;
    float* out = planes + (size_t)b * Pc * ns_pad + s;
    float col[HOG_MAX_COLS];
    for (int c = 0; c < Sc; ++c) { col[c] = 0.f; out[(size_t)c * ns_pad] = 0.f; }
    const int rows = (H / HOG_GRID) * HOG_GRID, groups = W / HOG_GRID;
    for (int y = 0; y < rows; ++y) {
      float run = 0.f;
      for (int c = 0; c < groups; ++c) {
        float g[HOG_GRID];
        int q[HOG_GRID];
#pragma unroll
        for (int k = 0; k < HOG_GRID; ++k) {
          const size_t px = (size_t)(y * W + c * HOG_GRID + k) * ns_pad + s;
          g[k] = grad[px];
          q[k] = qangle[px];
        }
#pragma unroll
        for (int k = 0; k < HOG_GRID; ++k)
          if (b == 9 || q[k] == b) run = __fadd_rn(run, g[k]);
        col[c + 1] = __fadd_rn(col[c + 1], run);                 // plane[y + 1][4 (c + 1)] = plane[y][4 (c + 1)] + strSum
      }
      if (((y + 1) & (HOG_GRID - 1)) == 0) {
        float* row = out + (size_t)((y + 1) / HOG_GRID) * Sc * ns_pad;
        for (int c = 0; c < Sc; ++c) row[(size_t)c * ns_pad] = col[c];
      }
    }
  }
}

// the full (H + 1) x (W + 1) planes of ONE sample in the reference's layout, for the parity getter (one thread per plane)
__global__ void k_hog_planes_full(const float* __restrict__ grad, const uint8_t* __restrict__ qangle, int W, int H, int s, int ns_pad,
                                  float* __restrict__ out10) {
  const int b = threadIdx.x;
  if (b >= 10) return;
  const int S = W + 1, P = S * (H + 1);
  float* out = out10 + (size_t)b * P;
  for (int X = 0; X <= W; ++X) out[X] = 0.f;
  for (int y = 0; y < H; ++y) {
    float run = 0.f;
    out[(y + 1) * S] = 0.f;
    for (int x = 0; x < W; ++x) {
      const size_t px = (size_t)(y * W + x) * ns_pad + s;
      if (b == 9 || qangle[px] == b) run = __fadd_rn(run, grad[px]);
      out[(y + 1) * S + x + 1] = __fadd_rn(out[y * S + x + 1], run);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// operator()(var, sample) for arbitrary pairs (debug / drop-in queries; not the precalc hot path)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float haar_value(const HaarTaps& t, const int32_t* __restrict__ sum_tiles,
                                            const int32_t* __restrict__ tilted_tiles, const float* __restrict__ nf,
                                            int P, int s) {
  const int32_t* base = (t.tilted() ? tilted_tiles : sum_tiles) + (size_t)(s >> 5) * P * TILE + (s & 31);
  int acc = 0;
  const int nt = t.ntaps();
  for (int k = 0; k < nt; ++k) acc += tap_coef(t.w[k]) * base[(size_t)tap_off(t.w[k]) * TILE];
  const float n = nf[s];
  return n == 0.f ? 0.f : __fdiv_rn((float)acc, n);
}

__device__ __forceinline__ int lbp_code(const LbpTaps& t, const int32_t* __restrict__ sum_tiles, int P, int s) {
  const int32_t* base = sum_tiles + (size_t)(s >> 5) * P * TILE + (s & 31);
  int v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = base[(size_t)t.off[k] * TILE];
  auto cell = [&](int gx, int gy) { const int a = gy * 4 + gx; return v[a] - v[a + 1] - v[a + 4] + v[a + 5]; };
  const int c = cell(1, 1);
  return (cell(0, 0) >= c ? 128 : 0) | (cell(1, 0) >= c ? 64 : 0) | (cell(2, 0) >= c ? 32 : 0) |
         (cell(2, 1) >= c ? 16 : 0) | (cell(2, 2) >= c ? 8 : 0) | (cell(1, 2) >= c ? 4 : 0) |
         (cell(0, 2) >= c ? 2 : 0) | (cell(0, 1) >= c ? 1 : 0);
}

__device__ __forceinline__ float hog_value(const HogTaps& t, const float* __restrict__ planes, int P, int ns_pad,
                                           int comp, int s) {
  const int bin = comp % 9, cell = comp / 9;
  const float* ph = planes + (size_t)bin * P * ns_pad + s;
  const float* pn = planes + (size_t)9 * P * ns_pad + s;
  float res = __fadd_rn(__fsub_rn(__fsub_rn(ph[(size_t)t.off[cell][0] * ns_pad], ph[(size_t)t.off[cell][1] * ns_pad]),
                                  ph[(size_t)t.off[cell][2] * ns_pad]), ph[(size_t)t.off[cell][3] * ns_pad]);
  float nrm = __fadd_rn(__fsub_rn(__fsub_rn(pn[(size_t)t.off[0][0] * ns_pad], pn[(size_t)t.off[1][1] * ns_pad]),
                                  pn[(size_t)t.off[2][2] * ns_pad]), pn[(size_t)t.off[3][3] * ns_pad]);
  return res > 0.001f ? __fdiv_rn(res, __fadd_rn(nrm, 0.001f)) : 0.f;
}

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

__device__ __forceinline__ float eval_var(const EvalView& e, int var, int s) {
  if (e.ftype == CC_FEATURE_HAAR) return haar_value(e.haar_taps[var], e.sum, e.tilted, e.nf, e.P, s);
  if (e.ftype == CC_FEATURE_LBP) return (float)lbp_code(e.lbp_taps[var], e.sum, e.P, s);
  return hog_value(e.hog_taps[var / 36], e.hog, e.P, e.ns_pad, var % 36, s);
}

__global__ void k_eval_values(EvalView e, const int* __restrict__ vars, int nv, int s0, int ns, float* __restrict__ out) {
  const long long total = (long long)nv * ns;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(t / ns), b = (int)(t - (long long)a * ns);
    out[t] = eval_var(e, vars[a], s0 + b);
  }
}

EvalView make_eval_view_store(const cc_ctx* ctx, const EvalStore& st) {
  EvalView e;
  e.ftype = ctx->ftype; e.P = ctx->ftype == CC_FEATURE_HOG ? ctx->hogP : ctx->P; e.ns_pad = st.ns_pad;
  e.sum = st.d_sum; e.tilted = st.d_tilted; e.nf = st.d_nf; e.hog = st.d_hog;
  e.haar_taps = ctx->d_haar_taps; e.lbp_taps = ctx->d_lbp_taps; e.hog_taps = ctx->d_hog_taps;
  return e;
}
EvalView make_eval_view(const cc_ctx* ctx) { return make_eval_view_store(ctx, main_store(ctx)); }

EvalStore main_store(const cc_ctx* ctx) {
  EvalStore s;
  s.max_samples = ctx->max_samples; s.ntiles = ctx->ntiles; s.ns_pad = ctx->ns_pad;
  s.d_imgs = ctx->d_imgs; s.d_labels = ctx->d_labels; s.d_sum = ctx->d_sum; s.d_tilted = ctx->d_tilted;
  s.d_nf = ctx->d_nf; s.d_hog = ctx->d_hog;
  return s;
}

void store_free(EvalStore& st) {
  cudaFree(st.d_imgs); cudaFree(st.d_labels); cudaFree(st.d_sum); cudaFree(st.d_tilted); cudaFree(st.d_nf); cudaFree(st.d_hog);
  st = EvalStore();
}

int store_alloc(cc_ctx* ctx, EvalStore& st, int max_samples) {
  st.max_samples = max_samples;
  st.ntiles = (max_samples + TILE - 1) / TILE;
  st.ns_pad = st.ntiles * TILE;
  const size_t npx = (size_t)ctx->W * ctx->H;
  CC_CUDA(cudaMalloc(&st.d_imgs, (size_t)st.ns_pad * npx));
  CC_CUDA(cudaMemsetAsync(st.d_imgs, 0, (size_t)st.ns_pad * npx, ctx->stream));
  CC_CUDA(cudaMalloc(&st.d_labels, st.ns_pad));
  CC_CUDA(cudaMemsetAsync(st.d_labels, 0, st.ns_pad, ctx->stream));
  const size_t tile_bytes = (size_t)st.ntiles * ctx->P * TILE * sizeof(int32_t);
  if (ctx->ftype == CC_FEATURE_HAAR || ctx->ftype == CC_FEATURE_LBP) {
    CC_CUDA(cudaMalloc(&st.d_sum, tile_bytes));
    CC_CUDA(cudaMemsetAsync(st.d_sum, 0, tile_bytes, ctx->stream));
  }
  if (ctx->ftype == CC_FEATURE_HAAR) {
    if (ctx->haar_mode == CC_HAAR_ALL) {
      CC_CUDA(cudaMalloc(&st.d_tilted, tile_bytes));
      CC_CUDA(cudaMemsetAsync(st.d_tilted, 0, tile_bytes, ctx->stream));
    }
    CC_CUDA(cudaMalloc(&st.d_nf, st.ns_pad * sizeof(float)));
    CC_CUDA(cudaMemsetAsync(st.d_nf, 0, st.ns_pad * sizeof(float), ctx->stream));
  }
  if (ctx->ftype == CC_FEATURE_HOG) {
    const size_t bytes = (size_t)10 * ctx->hogP * st.ns_pad * sizeof(float);
    CC_CUDA(cudaMalloc(&st.d_hog, bytes));
    CC_CUDA(cudaMemsetAsync(st.d_hog, 0, bytes, ctx->stream));
  }
  return CC_OK;
}

int eval_free(cc_ctx* ctx) {
  cudaFree(ctx->d_imgs); cudaFree(ctx->d_labels); cudaFree(ctx->d_sum); cudaFree(ctx->d_tilted);
  cudaFree(ctx->d_nf); cudaFree(ctx->d_hog); cudaFree(ctx->d_haar_taps); cudaFree(ctx->d_lbp_taps);
  cudaFree(ctx->d_hog_taps); cudaFree(ctx->d_taps_upright); cudaFree(ctx->d_taps_tilted); cudaFree(ctx->d_sep_tab_row); cudaFree(ctx->d_sep_tab_col); cudaFree(ctx->d_sep_patterns); cudaFree(ctx->d_sep_tab_t1); cudaFree(ctx->d_sep_tab_t2);
  ctx->d_imgs = ctx->d_labels = nullptr; ctx->d_sum = ctx->d_tilted = nullptr; ctx->d_nf = nullptr; ctx->d_hog = nullptr;
  ctx->d_haar_taps = nullptr; ctx->d_lbp_taps = nullptr; ctx->d_hog_taps = nullptr;
  ctx->d_taps_upright = ctx->d_taps_tilted = nullptr; ctx->d_sep_tab_row = ctx->d_sep_tab_col = ctx->d_sep_tab_t1 = ctx->d_sep_tab_t2 = nullptr; ctx->d_sep_patterns = nullptr; ctx->n_sep_patterns = 0; ctx->n_sep_tilted = 0; ctx->use_sep = false;
  ctx->ftype = -1;
  return CC_OK;
}

// shared by setImage and by mining (windows already in d_imgs)
int launch_integrals(cc_ctx* ctx, const EvalStore& st, int first, int n) {
  if (n <= 0) return CC_OK;
  const int last = first + n;
  if (ctx->ftype == CC_FEATURE_HOG) {
    // bytes per staged window: a multiple of 128 plus 4, so that the 32 windows of a CTA start in 32 different banks
    const int wstride = ((((HOG_G_ROWS + 2) * ctx->W) + 127) & ~127) + 4;
    CC_REQUIRE((size_t)32 * wstride <= 200 * 1024 && ctx->hogS <= HOG_MAX_COLS, CC_ERR_UNSUPPORTED, "HOG windows wider than %d pixels are not supported",
               std::min((200 * 1024 / 32 - 132) / (HOG_G_ROWS + 2), (HOG_MAX_COLS - 1) * HOG_GRID + HOG_GRID - 1));
    ScopedTimer tm(ctx, "integral", 2);
    const size_t npx = (size_t)ctx->W * ctx->H;
    float* grad = nullptr; uint8_t* qa = nullptr;
    CC_CUDA(cudaMallocAsync(&grad, npx * st.ns_pad * sizeof(float), ctx->stream));
    CC_CUDA(cudaMallocAsync(&qa, npx * st.ns_pad, ctx->stream));
    CC_CUDA(cudaFuncSetAttribute(k_hog_gradient, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * wstride));
    k_hog_gradient<<<std::max(1, std::min((n + 31) / 32, ctx->sm_count * 8)), HOG_G_THREADS, (size_t)32 * wstride, ctx->stream>>>(
        st.d_imgs, ctx->W, ctx->H, first, n, st.ns_pad, wstride, grad, qa);
    k_hog_integrals<<<std::max(1, std::min((n + 31) / 32, ctx->sm_count * 16)), HOG_INT_THREADS, 0, ctx->stream>>>(grad, qa, ctx->W, ctx->H, first, n, st.ns_pad, st.d_hog);
    CC_CUDA(cudaGetLastError());
    CC_CUDA(cudaFreeAsync(grad, ctx->stream));
    CC_CUDA(cudaFreeAsync(qa, ctx->stream));
    return CC_OK;
  }
  const int W = ctx->W, H = ctx->H, P = ctx->P;
  const int tile0 = first / TILE, tile1 = (last - 1) / TILE;
  const size_t smem_img = (((size_t)32 * W * H + 15) & ~(size_t)15) + (size_t)32 * 2 * (W + 2) * 4;
  const int staged = (size_t)P * 33 * 4 + smem_img <= 200 * 1024;
  const size_t smem = smem_img + (staged ? (size_t)P * 33 * 4 : 0);
  const int want_tilted = ctx->ftype == CC_FEATURE_HAAR && ctx->haar_mode == CC_HAAR_ALL;
  const int want_nf = ctx->ftype == CC_FEATURE_HAAR;
  const int nchunk = (W + 31) / 32;
  CC_REQUIRE(nchunk <= 4 && smem <= 227 * 1024, CC_ERR_UNSUPPORTED,
             "integral kernel: window %dx%d needs %zu B of shared memory (limit 227 KB, width <= 128)", W, H, smem);
  ScopedTimer tm(ctx, "integral", 1);
#define LAUNCH_INT(NC)                                                                                         \
  {                                                                                                            \
    CC_CUDA(cudaFuncSetAttribute(k_integral_tiles<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_integral_tiles<NC><<<tile1 - tile0 + 1, 1024, smem, ctx->stream>>>(st.d_imgs, W, H, first, last, st.d_sum, \
                                                                         st.d_tilted, st.d_nf, want_tilted, want_nf, staged); \
  }
  switch (nchunk) {
    case 1: LAUNCH_INT(1); break;
    case 2: LAUNCH_INT(2); break;
    case 3: LAUNCH_INT(3); break;
    default: LAUNCH_INT(4); break;
  }
#undef LAUNCH_INT
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

}  // namespace cc

using namespace cc;

extern "C" {

int cc_eval_init(cc_ctx* ctx, int feature_type, int haar_mode, int win_w, int win_h, int max_samples) {
  CC_REQUIRE(ctx, CC_ERR_BAD_ARG, "cc_eval_init: null ctx");
  CC_REQUIRE(max_samples > 0, CC_ERR_BAD_ARG, "cc_eval_init: _maxSampleCount > 0 required (src/features.cpp:75)");
  CC_REQUIRE(feature_type >= CC_FEATURE_HAAR && feature_type <= CC_FEATURE_HOG, CC_ERR_BAD_ARG, "cc_eval_init: unknown feature type %d", feature_type);
  CC_REQUIRE(win_w >= 4 && win_h >= 4 && win_w <= 255 && win_h <= 255, CC_ERR_BAD_ARG, "cc_eval_init: window %dx%d out of range [4,255]", win_w, win_h);
  CC_CUDA(cudaSetDevice(ctx->device));
  cc::mine_free(ctx);
  cc::boost_free(ctx);
  cc::eval_free(ctx);
  ctx->ftype = feature_type; ctx->haar_mode = haar_mode; ctx->W = win_w; ctx->H = win_h;
  ctx->P = (win_w + 1) * (win_h + 1);
  ctx->hogS = win_w / HOG_GRID + 1; ctx->hogP = ctx->hogS * (win_h / HOG_GRID + 1);
  ctx->max_samples = max_samples;
  ctx->ntiles = (max_samples + TILE - 1) / TILE;
  ctx->ns_pad = ctx->ntiles * TILE;
  const int S = win_w + 1;
  if (feature_type == CC_FEATURE_HAAR) {
    CC_REQUIRE(haar_mode >= CC_HAAR_BASIC && haar_mode <= CC_HAAR_ALL, CC_ERR_BAD_ARG, "cc_eval_init: unknown Haar mode %d", haar_mode);
    ctx->haar = make_haar_catalog(win_w, win_h, haar_mode);
    ctx->num_features = (int)ctx->haar.size(); ctx->feat_size = 1; ctx->max_cat = 0;
    std::vector<HaarTaps> taps(ctx->haar.size());
    for (size_t i = 0; i < ctx->haar.size(); ++i) taps[i] = make_haar_taps(ctx->haar[i], S);
    CC_CUDA(cudaMalloc(&ctx->d_haar_taps, taps.size() * sizeof(HaarTaps)));
    CC_CUDA(cudaMemcpyAsync(ctx->d_haar_taps, taps.data(), taps.size() * sizeof(HaarTaps), cudaMemcpyHostToDevice, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
  } else if (feature_type == CC_FEATURE_LBP) {
    ctx->lbp = make_lbp_catalog(win_w, win_h);
    ctx->num_features = (int)ctx->lbp.size(); ctx->feat_size = 1; ctx->max_cat = 256;
    std::vector<LbpTaps> taps(ctx->lbp.size());
    for (size_t i = 0; i < ctx->lbp.size(); ++i) taps[i] = make_lbp_taps(ctx->lbp[i], S);
    CC_CUDA(cudaMalloc(&ctx->d_lbp_taps, taps.size() * sizeof(LbpTaps)));
    CC_CUDA(cudaMemcpyAsync(ctx->d_lbp_taps, taps.data(), taps.size() * sizeof(LbpTaps), cudaMemcpyHostToDevice, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
  } else {
    ctx->hog = make_hog_catalog(win_w, win_h);
    ctx->num_features = (int)ctx->hog.size(); ctx->feat_size = 36; ctx->max_cat = 0;
    for (const HogFeatureH& f : ctx->hog)
      CC_REQUIRE(f.cell0.x % HOG_GRID == 0 && f.cell0.y % HOG_GRID == 0 && f.cell0.w % HOG_GRID == 0 && f.cell0.h % HOG_GRID == 0, CC_ERR_STATE,
                 "cc_eval_init: HOG cell (%d,%d,%d,%d) is not aligned to the plane grid", f.cell0.x, f.cell0.y, f.cell0.w, f.cell0.h);
    std::vector<HogTaps> taps(ctx->hog.size());
    for (size_t i = 0; i < ctx->hog.size(); ++i) taps[i] = make_hog_taps(ctx->hog[i], ctx->hogS, HOG_GRID);
    CC_CUDA(cudaMalloc(&ctx->d_hog_taps, std::max<size_t>(1, taps.size()) * sizeof(HogTaps)));
    CC_CUDA(cudaMemcpyAsync(ctx->d_hog_taps, taps.data(), taps.size() * sizeof(HogTaps), cudaMemcpyHostToDevice, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  {
    EvalStore st;
    CC_CHECK(store_alloc(ctx, st, max_samples));
    ctx->d_imgs = st.d_imgs; ctx->d_labels = st.d_labels; ctx->d_sum = st.d_sum; ctx->d_tilted = st.d_tilted;
    ctx->d_nf = st.d_nf; ctx->d_hog = st.d_hog;
  }
  ctx->var_count = ctx->num_features * ctx->feat_size;
  ctx->precalc_valid = false;
  CC_CHECK(cc_boost_set_shard(ctx, ctx->shard_rank, ctx->shard_world));
  return CC_OK;
}

int cc_eval_num_features(cc_ctx* ctx) { return ctx ? ctx->num_features : 0; }
int cc_eval_feature_size(cc_ctx* ctx) { return ctx ? ctx->feat_size : 0; }
int cc_eval_max_cat_count(cc_ctx* ctx) { return ctx ? ctx->max_cat : 0; }

int cc_eval_get_feature(cc_ctx* ctx, int fi, int32_t rects[12], float weights[3], int32_t* tilted) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_eval_get_feature: evaluator not initialised");
  CC_REQUIRE(fi >= 0 && fi < ctx->num_features, CC_ERR_BAD_ARG, "cc_eval_get_feature: index %d out of range", fi);
  std::memset(rects, 0, sizeof(int32_t) * 12);
  weights[0] = weights[1] = weights[2] = 0.f;
  *tilted = 0;
  if (ctx->ftype == CC_FEATURE_HAAR) {
    const HaarFeatureH& f = ctx->haar[fi];
    for (int j = 0; j < 3; ++j) {
      rects[j * 4] = f.r[j].x; rects[j * 4 + 1] = f.r[j].y; rects[j * 4 + 2] = f.r[j].w; rects[j * 4 + 3] = f.r[j].h;
      weights[j] = f.wt[j];
    }
    *tilted = f.tilted;
  } else {
    const RectI& r = ctx->ftype == CC_FEATURE_LBP ? ctx->lbp[fi].r : ctx->hog[fi].cell0;
    rects[0] = r.x; rects[1] = r.y; rects[2] = r.w; rects[3] = r.h;
  }
  return CC_OK;
}

int cc_eval_set_images_device(cc_ctx* ctx, const uint8_t* dev_imgs, const uint8_t* dev_labels, int first, int n) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_eval_set_images: evaluator not initialised");
  CC_REQUIRE(first >= 0 && n >= 0 && first + n <= ctx->max_samples, CC_ERR_BAD_ARG,
             "cc_eval_set_images: slots [%d,%d) exceed maxSampleCount %d (CV_Assert(idx < cls.rows), src/features.cpp:87)", first, first + n, ctx->max_samples);
  if (n == 0) return CC_OK;
  CC_CUDA(cudaSetDevice(ctx->device));
  const size_t npx = (size_t)ctx->W * ctx->H;
  if (dev_imgs != ctx->d_imgs + (size_t)first * npx)
    CC_CUDA(cudaMemcpyAsync(ctx->d_imgs + (size_t)first * npx, dev_imgs, (size_t)n * npx, cudaMemcpyDeviceToDevice, ctx->stream));
  if (dev_labels && dev_labels != ctx->d_labels + first)
    CC_CUDA(cudaMemcpyAsync(ctx->d_labels + first, dev_labels, n, cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->precalc_valid = false;
  return launch_integrals(ctx, main_store(ctx), first, n);
}

int cc_eval_set_images(cc_ctx* ctx, const uint8_t* host_imgs, const uint8_t* host_labels, int first, int n) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_eval_set_images: evaluator not initialised");
  CC_REQUIRE(host_imgs && host_labels, CC_ERR_BAD_ARG, "cc_eval_set_images: null buffer");
  CC_REQUIRE(first >= 0 && n >= 0 && first + n <= ctx->max_samples, CC_ERR_BAD_ARG,
             "cc_eval_set_images: slots [%d,%d) exceed maxSampleCount %d (CV_Assert(idx < cls.rows), src/features.cpp:87)", first, first + n, ctx->max_samples);
  if (n == 0) return CC_OK;
  CC_CUDA(cudaSetDevice(ctx->device));
  const size_t npx = (size_t)ctx->W * ctx->H;
  CC_CUDA(cudaMemcpyAsync(ctx->d_imgs + (size_t)first * npx, host_imgs, (size_t)n * npx, cudaMemcpyHostToDevice, ctx->stream));
  CC_CUDA(cudaMemcpyAsync(ctx->d_labels + first, host_labels, n, cudaMemcpyHostToDevice, ctx->stream));
  ctx->precalc_valid = false;
  return launch_integrals(ctx, main_store(ctx), first, n);
}

int cc_eval_values(cc_ctx* ctx, const int32_t* var_idx, int nv, int s0, int ns, float* out) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_eval_values: evaluator not initialised");
  CC_REQUIRE(var_idx && out && nv > 0 && ns > 0, CC_ERR_BAD_ARG, "cc_eval_values: bad arguments");
  CC_REQUIRE(s0 >= 0 && s0 + ns <= ctx->max_samples, CC_ERR_BAD_ARG, "cc_eval_values: sample range out of bounds");
  for (int i = 0; i < nv; ++i)
    CC_REQUIRE(var_idx[i] >= 0 && var_idx[i] < ctx->var_count, CC_ERR_BAD_ARG, "cc_eval_values: variable %d out of range", var_idx[i]);
  CC_CUDA(cudaSetDevice(ctx->device));
  int* dv = nullptr; float* dout = nullptr;
  CC_CUDA(cudaMallocAsync(&dv, nv * sizeof(int), ctx->stream));
  CC_CUDA(cudaMallocAsync(&dout, (size_t)nv * ns * sizeof(float), ctx->stream));
  CC_CUDA(cudaMemcpyAsync(dv, var_idx, nv * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  const long long total = (long long)nv * ns;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
  k_eval_values<<<blocks, 256, 0, ctx->stream>>>(make_eval_view(ctx), dv, nv, s0, ns, dout);
  CC_CUDA(cudaGetLastError());
  ctx->launches += 1;
  CC_CUDA(cudaMemcpyAsync(out, dout, (size_t)nv * ns * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  CC_CUDA(cudaFreeAsync(dv, ctx->stream));
  CC_CUDA(cudaFreeAsync(dout, ctx->stream));
  return CC_OK;
}

int cc_eval_value(cc_ctx* ctx, int var_idx, int sample_idx, float* out) {
  int32_t v = var_idx;
  return cc_eval_values(ctx, &v, 1, sample_idx, 1, out);
}

int cc_eval_get_integrals(cc_ctx* ctx, int si, int32_t* sum, int32_t* tilted, float* nf) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_eval_get_integrals: evaluator not initialised");
  CC_REQUIRE(si >= 0 && si < ctx->max_samples, CC_ERR_BAD_ARG, "cc_eval_get_integrals: sample %d out of range", si);
  CC_REQUIRE(ctx->ftype != CC_FEATURE_HOG, CC_ERR_BAD_ARG, "cc_eval_get_integrals: HOG evaluator has no integer integrals");
  CC_CUDA(cudaSetDevice(ctx->device));
  const size_t tile_off = (size_t)(si / TILE) * ctx->P * TILE + (si % TILE);
  if (sum)
    CC_CUDA(cudaMemcpy2DAsync(sum, sizeof(int32_t), ctx->d_sum + tile_off, TILE * sizeof(int32_t), sizeof(int32_t), ctx->P, cudaMemcpyDeviceToHost, ctx->stream));
  if (tilted && ctx->d_tilted)
    CC_CUDA(cudaMemcpy2DAsync(tilted, sizeof(int32_t), ctx->d_tilted + tile_off, TILE * sizeof(int32_t), sizeof(int32_t), ctx->P, cudaMemcpyDeviceToHost, ctx->stream));
  if (nf && ctx->d_nf) CC_CUDA(cudaMemcpyAsync(nf, ctx->d_nf + si, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  return CC_OK;
}

int cc_eval_get_hog_planes(cc_ctx* ctx, int si, float* planes10) {
  CC_REQUIRE(ctx && ctx->ftype == CC_FEATURE_HOG, CC_ERR_STATE, "cc_eval_get_hog_planes: not a HOG evaluator");
  CC_REQUIRE(si >= 0 && si < ctx->max_samples && planes10, CC_ERR_BAD_ARG, "cc_eval_get_hog_planes: sample %d out of range", si);
  CC_CUDA(cudaSetDevice(ctx->device));
  // the resident planes are compact (every HOG_GRID-th row / column); the getter returns the reference's full layout, computed
  // here for this one sample with the same gradient kernel and the same order of additions
  const size_t npx = (size_t)ctx->W * ctx->H;
  const int wstride = ((((HOG_G_ROWS + 2) * ctx->W) + 127) & ~127) + 4;
  float *grad = nullptr, *full = nullptr;
  uint8_t* qa = nullptr;
  CC_CUDA(cudaMallocAsync(&grad, npx * ctx->ns_pad * sizeof(float), ctx->stream));
  CC_CUDA(cudaMallocAsync(&qa, npx * ctx->ns_pad, ctx->stream));
  CC_CUDA(cudaMallocAsync(&full, (size_t)10 * ctx->P * sizeof(float), ctx->stream));
  CC_CUDA(cudaFuncSetAttribute(k_hog_gradient, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * wstride));
  k_hog_gradient<<<1, HOG_G_THREADS, (size_t)32 * wstride, ctx->stream>>>(ctx->d_imgs, ctx->W, ctx->H, si, 1, ctx->ns_pad, wstride, grad, qa);
  k_hog_planes_full<<<1, 32, 0, ctx->stream>>>(grad, qa, ctx->W, ctx->H, si, ctx->ns_pad, full);
  CC_CUDA(cudaGetLastError());
  CC_CUDA(cudaMemcpyAsync(planes10, full, (size_t)10 * ctx->P * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  CC_CUDA(cudaFreeAsync(grad, ctx->stream)); CC_CUDA(cudaFreeAsync(qa, ctx->stream)); CC_CUDA(cudaFreeAsync(full, ctx->stream));
  return CC_OK;
}

/* the resident compact planes of one sample: 10 planes of (H / 4 + 1) x (W / 4 + 1) entries = the reference's planes at rows / columns 0, 4, 8, ... */
int cc_eval_get_hog_planes_compact(cc_ctx* ctx, int si, float* planes10c) {
  CC_REQUIRE(ctx && ctx->ftype == CC_FEATURE_HOG, CC_ERR_STATE, "cc_eval_get_hog_planes_compact: not a HOG evaluator");
  CC_REQUIRE(si >= 0 && si < ctx->max_samples && planes10c, CC_ERR_BAD_ARG, "cc_eval_get_hog_planes_compact: sample %d out of range", si);
  CC_CUDA(cudaSetDevice(ctx->device));
  CC_CUDA(cudaMemcpy2DAsync(planes10c, sizeof(float), ctx->d_hog + si, (size_t)ctx->ns_pad * sizeof(float), sizeof(float),
                            (size_t)10 * ctx->hogP, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  return CC_OK;
}

int cc_eval_get_labels(cc_ctx* ctx, int first, int n, float* cls) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_eval_get_labels: evaluator not initialised");
  CC_REQUIRE(first >= 0 && n >= 0 && first + n <= ctx->max_samples, CC_ERR_BAD_ARG, "cc_eval_get_labels: range out of bounds");
  std::vector<uint8_t> tmp(n);
  CC_CUDA(cudaSetDevice(ctx->device));
  CC_CUDA(cudaMemcpyAsync(tmp.data(), ctx->d_labels + first, n, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) cls[i] = (float)tmp[i];
  return CC_OK;
}

int cc_eval_get_images(cc_ctx* ctx, int first, int n, uint8_t* imgs) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_eval_get_images: evaluator not initialised");
  CC_REQUIRE(first >= 0 && n >= 0 && first + n <= ctx->max_samples, CC_ERR_BAD_ARG, "cc_eval_get_images: range out of bounds");
  const size_t npx = (size_t)ctx->W * ctx->H;
  CC_CUDA(cudaSetDevice(ctx->device));
  CC_CUDA(cudaMemcpyAsync(imgs, ctx->d_imgs + (size_t)first * npx, (size_t)n * npx, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  return CC_OK;
}

}  // extern "C"

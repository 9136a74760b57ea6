// K2 (feature precalc into the fval matrix) and K3 (per-feature segmented radix sort).
// Replaces CvCascadeBoostTrainData::setData / precalculate and the three Feature*Precalc bodies
// (src/o_cvcascadeboosttraindata.cpp:214-368, :490-596); every row is cached in HBM, so the
// reference's -precalcValBufSize / -precalcIdxBufSize become advisory.
#include <algorithm>

#include "cc_internal.h"

namespace cc {

struct EvalView;
EvalView make_eval_view(const cc_ctx* ctx);

// ---- mbarrier / bulk-copy (TMA 1-D) helpers -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// K2 / Haar.  CTA = TILES_PER_CTA tiles of 32 samples resident in shared memory as [tile][P][32]
// (one cp.async.bulk each), lane == sample, so every tap is a conflict-free 128-byte LDS.  A warp
// walks the features of its chunk; the tap words of a feature are warp-uniform 128-bit loads, reused
// for all tiles.  value = (float)(sum coef*I) / nf with IEEE division (haarfeatures.h:108-112).
// Stores are one full 128-byte line per (feature, tile).
// ------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ int haar_acc(const uint32_t (&w)[HAAR_MAX_TAPS], const int32_t* __restrict__ tile_lane) {
  int acc = 0;
#pragma unroll
  for (int k = 0; k < NT; ++k) acc += tap_coef(w[k]) * tile_lane[tap_off(w[k]) * TILE];
  return acc;
}

constexpr int PRECALC_THREADS = 1024;

// taps: compact list of this launch (one plane), w[11] of every entry = global feature index
// IN_GLOBAL: windows whose 32-sample tile (P * 128 bytes) does not fit in shared memory (beyond ~41x41; the reference's
// own end-to-end tests use 75x32) read the taps straight from the tiles in HBM / L2 -- same addresses, same coalescing.
template <int TILES_PER_CTA, bool IN_GLOBAL>
__global__ void __launch_bounds__(PRECALC_THREADS, 1)
k_precalc_haar(const int32_t* __restrict__ plane_tiles, const float* __restrict__ nf, const HaarTaps* __restrict__ taps,
               int nrows_launch, int rows_per_chunk, int row_base, float* __restrict__ fval, int n_stride, int P, int ntiles_used) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tile0 = blockIdx.x * TILES_PER_CTA;
  const int ntl = min(TILES_PER_CTA, ntiles_used - tile0);
  const int32_t* stile = IN_GLOBAL ? plane_tiles + (size_t)tile0 * P * TILE : reinterpret_cast<const int32_t*>(smem_raw);   // [TILES][P][32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)TILES_PER_CTA * P * TILE * sizeof(int32_t));
  if (!IN_GLOBAL) {
    const uint32_t tile_bytes = (uint32_t)P * TILE * sizeof(int32_t);
    if (threadIdx.x == 0) {
      int32_t* dst = reinterpret_cast<int32_t*>(smem_raw);
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(bar, tile_bytes * ntl);
      for (int k = 0; k < ntl; ++k) bulk_g2s(dst + (size_t)k * P * TILE, plane_tiles + (size_t)(tile0 + k) * P * TILE, tile_bytes, bar);
    }
    __syncthreads();
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = PRECALC_THREADS / 32;
  const int r_begin = blockIdx.y * rows_per_chunk;
  const int r_end = min(r_begin + rows_per_chunk, nrows_launch);
  int r = r_begin + warp;
  // software pipeline: the tap words of the next feature are in flight while this one is evaluated
  uint4 a = make_uint4(0, 0, 0, 0), b = a, c = a;
  if (r < r_end) { const uint4* tp = reinterpret_cast<const uint4*>(taps + r); a = __ldg(tp); b = __ldg(tp + 1); c = __ldg(tp + 2); }
  float nfv[TILES_PER_CTA];
#pragma unroll
  for (int k = 0; k < TILES_PER_CTA; ++k) nfv[k] = (k < ntl) ? nf[(tile0 + k) * TILE + lane] : 0.f;
  if (!IN_GLOBAL) mbar_wait(bar, 0);

  for (; r < r_end; r += nwarps) {
    const uint32_t w[HAAR_MAX_TAPS] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    if (r + nwarps < r_end) { const uint4* tp = reinterpret_cast<const uint4*>(taps + r + nwarps); a = __ldg(tp); b = __ldg(tp + 1); c = __ldg(tp + 2); }
    const int nt = (int)((w[0] >> 24) & 15u);
    const int f = (int)w[11];
    float* out = fval + (size_t)(f - row_base) * n_stride + tile0 * TILE + lane;
#pragma unroll
    for (int k = 0; k < TILES_PER_CTA; ++k) {
      if (k < ntl) {
        const int32_t* tl = stile + (size_t)k * P * TILE + lane;
        int acc;
        switch (nt) {
          case 6: acc = haar_acc<6>(w, tl); break;
          case 8: acc = haar_acc<8>(w, tl); break;
          case 9: acc = haar_acc<9>(w, tl); break;
          default: {
            acc = 0;
#pragma unroll
            for (int q = 0; q < 10; ++q)
              if (q < nt) acc += tap_coef(w[q]) * tl[tap_off(w[q]) * TILE];
          }
        }
        const float v = nfv[k] == 0.f ? 0.f : __fdiv_rn((float)acc, nfv[k]);
        __stcs(out + k * TILE, v);      // streaming store: the fval matrix is not re-read by this kernel
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Correctly rounded (int)acc / nf without the generic IEEE division sequence.
// With r = RN(1/nf) precomputed per sample: q0 = RN(a*r); e0 = fma(-q0, nf, a) (exact); q1 = RN(q0 + e0*r);
// e1 = fma(-q1, nf, a); q = RN(q1 + e1*r).  |a| < 2^24 is an integer, so a/nf is never a rounding tie and the
// two correction steps land on RN(a/nf) (Markstein-style iteration; validated against __fdiv_rn on the GPU by
// cc_debug_div_check / tests/test_gpu_features.py::test_fast_division_is_exact).  nf == 0 is handled by r = 0
// (all terms vanish and the result is +0, as the reference returns).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float haar_recip(float nfv) { return nfv == 0.f ? 0.f : __frcp_rn(nfv); }
template <int STEPS>
__device__ __forceinline__ float haar_div(float a, float nfv, float r) {
  float q = __fmul_rn(a, r);
#pragma unroll
  for (int k = 0; k < STEPS; ++k) {
    const float e = __fmaf_rn(-q, nfv, a);
    q = __fmaf_rn(e, r, q);
  }
  return q;
}

// ------------------------------------------------------------------------------------------------
// K2 / upright Haar, separable form.  Every upright shape of the catalog is a difference of a "row
// function" (or, for the y-shapes, a "column function") of the integral image:
//   x2   : acc = gA(y+dy) - gA(y)                 gA(r) = I(r,x) - 2 I(r,x+dx) + I(r,x+2dx)
//   x3   : acc = gB(y+dy) - gB(y)                 gB(r) = I(r,x) - 2 I(r,x+dx) + 2 I(r,x+2dx) - I(r,x+3dx)
//   x4   : acc = gC(y+dy) - gC(y)                 gC(r) = I(r,x) - 2 I(r,x+dx) + 2 I(r,x+3dx) - I(r,x+4dx)
//   x2y2 : acc = gA(y) - 2 gA(y+dy) + gA(y+2dy)
//   3x3  : acc = -(gW(y+3dy) - gW(y)) + 9 (gM(y+2dy) - gM(y+dy)),  gW(r) = I(r,x+3dx) - I(r,x), gM(r) = I(r,x+2dx) - I(r,x+dx)
//   y2/y3/y4 : the transposes (column functions hA/hB/hC of c for fixed (y,dy); acc = h(x+dx) - h(x))
// (expand -1*rect0 + 2*rect1 [+ 2*rect2 | 9*rect1] of src/haarfeatures.cpp:135-195 into corner terms).
// All of it is exact int32 arithmetic, so (float)acc / nf equals the reference's value bit for bit.
// A work item is one (x,dx) [or (y,dy)] pattern: the function is evaluated once per row into a per-warp
// scratch column (lane == sample) and every feature of the pattern then costs ~1 shared load and a subtract,
// instead of 6-9 tap loads and multiply-adds.  The feature -> fval row mapping comes from dense index tables
// (-1 = shape absent from this mode / not in this rank's shard).
// ------------------------------------------------------------------------------------------------
constexpr int SEP_THREADS = 1024;
constexpr int SEP_WARPS = SEP_THREADS / 32;

// one fval cell: row f of this sample's column.  The byte stride of a row fits 32 bits, so the address is one IMAD.WIDE.
template <int STEPS>
__device__ __forceinline__ void sep_store(char* __restrict__ out_bytes, int f, uint32_t stride_bytes, int acc, float nf, float rnf) {
  float* p = reinterpret_cast<float*>(out_bytes + (unsigned long long)(uint32_t)f * stride_bytes);
  __stcs(p, haar_div<STEPS>((float)acc, nf, rnf));
}

// Walk the feature-index table row tp[0 .. nd) (groups of 4, -1 = absent, prefetched one group ahead) and emit
// ACC(d) for every present entry; d = 1-based step (dy or dx).
#define SEP_EMIT(tp, nd, ACC)                                                                              \
  {                                                                                                        \
    int4 cur4 = __ldg(reinterpret_cast<const int4*>(tp));                                                  \
    for (int d0_ = 0; d0_ < (nd); d0_ += 4) {                                                              \
      const int4 f4_ = cur4;                                                                               \
      if (d0_ + 4 < (nd)) cur4 = __ldg(reinterpret_cast<const int4*>((tp) + d0_ + 4));                     \
      const int fs_[4] = {f4_.x, f4_.y, f4_.z, f4_.w};                                                     \
      _Pragma("unroll") for (int k_ = 0; k_ < 4; ++k_) {                                                   \
        const int d = d0_ + k_ + 1;                                                                        \
        if (fs_[k_] >= 0) sep_store<STEPS>(out_bytes, fs_[k_], stride_bytes, (ACC), nfv, rnf);             \
      }                                                                                                    \
    }                                                                                                      \
  }

template <int STEPS>
__device__ __forceinline__ void sep_body(const int32_t* __restrict__ tl, int32_t* __restrict__ sA, int* s_next, int p_begin, int p_end,
                                         const SepPattern* __restrict__ patterns, const int32_t* __restrict__ tab_row,
                                         const int32_t* __restrict__ tab_col, const int32_t* __restrict__ tab_t1,
                                         const int32_t* __restrict__ tab_t2, int W, int H, char* __restrict__ out_bytes,
                                         uint32_t stride_bytes, float nfv, float rnf, int lane) {
  const int S = W + 1;
  const int HP = (H + 3) & ~3, WP = (W + 3) & ~3;
  for (;;) {
    int pi = 0;
    if (lane == 0) pi = p_begin + atomicAdd(s_next, 1);
    pi = __shfl_sync(0xffffffffu, pi, 0);
    if (pi >= p_end) break;
    const SepPattern pat = patterns[pi];
    const int p = pat.p, d_ = pat.d, lo = pat.lo, hi = pat.hi, flags = pat.flags;
    if (pat.kind == 0) {
      // ---------------- row pattern: x = p, dx = d; functions of the row r = 0..H ----------------
      const int x = p, dx = d_;
      const bool has3 = x + 3 * dx <= W, has4 = x + 4 * dx <= W;
      auto tabp = [&](int t, int y) { return tab_row + ((((size_t)t * W + x) * W + (dx - 1)) * H + y) * HP; };
      if (flags & 0x09) {                      // gA: x2 (bit 0), x2y2 (bit 3)
        for (int r = 0; r <= H; ++r) {
          const int32_t* row = tl + (size_t)(r * S + x) * TILE;
          sA[r * TILE] = row[0] - 2 * row[dx * TILE] + row[2 * dx * TILE];
        }
        __syncwarp();
        for (int y = lo; y <= hi; ++y) {
          const int gy = sA[y * TILE];
          const int32_t* sy = sA + y * TILE;
          if (flags & 0x01) SEP_EMIT(tabp(0, y), H - y, sy[d * TILE] - gy)
          if (flags & 0x08) SEP_EMIT(tabp(3, y), (H - y) >> 1, gy - 2 * sy[d * TILE] + sy[2 * d * TILE])
        }
        __syncwarp();
      }
      if (has3 && (flags & 0x02)) {            // gB: x3
        for (int r = 0; r <= H; ++r) {
          const int32_t* row = tl + (size_t)(r * S + x) * TILE;
          sA[r * TILE] = row[0] - 2 * row[dx * TILE] + 2 * row[2 * dx * TILE] - row[3 * dx * TILE];
        }
        __syncwarp();
        for (int y = lo; y <= hi; ++y) {
          const int gy = sA[y * TILE];
          const int32_t* sy = sA + y * TILE;
          SEP_EMIT(tabp(1, y), H - y, sy[d * TILE] - gy)
        }
        __syncwarp();
      }
      if (has3 && (flags & 0x10)) {            // centre-surround: gW from the scratch column, gM straight from the tile
        for (int r = 0; r <= H; ++r) {
          const int32_t* row = tl + (size_t)(r * S + x) * TILE;
          sA[r * TILE] = row[3 * dx * TILE] - row[0];
        }
        __syncwarp();
        const int32_t* c1 = tl + (size_t)(x + dx) * TILE;
        const int32_t* c2 = tl + (size_t)(x + 2 * dx) * TILE;
        for (int y = lo; y <= hi; ++y) {
          const int wy = sA[y * TILE];
          const int32_t* sy = sA + y * TILE;
          const int32_t* m1 = c1 + (size_t)y * S * TILE;
          const int32_t* m2 = c2 + (size_t)y * S * TILE;
          SEP_EMIT(tabp(4, y), (H - y) / 3,
                   wy - sy[3 * d * TILE] + 9 * ((m2[2 * d * S * TILE] - m1[2 * d * S * TILE]) - (m2[d * S * TILE] - m1[d * S * TILE])))
        }
        __syncwarp();
      }
      if (has4 && (flags & 0x04)) {            // gC: x4
        for (int r = 0; r <= H; ++r) {
          const int32_t* row = tl + (size_t)(r * S + x) * TILE;
          sA[r * TILE] = row[0] - 2 * row[dx * TILE] + 2 * row[3 * dx * TILE] - row[4 * dx * TILE];
        }
        __syncwarp();
        for (int y = lo; y <= hi; ++y) {
          const int gy = sA[y * TILE];
          const int32_t* sy = sA + y * TILE;
          SEP_EMIT(tabp(2, y), H - y, sy[d * TILE] - gy)
        }
        __syncwarp();
      }
    } else if (pat.kind >= 2) {
      // ---------------- tilted patterns (the tile is the tilted integral) ----------------
      // In rotated coordinates a = X + Y, b = Y - X a tilted rectangle (x, y, w, h) is the axis-aligned box
      // [a0, a0 + 2w] x [b0, b0 + 2h] with (a0, b0) = (x + y, y - x) (corners of CV_TILTED_OFFSETS), so the tilted
      // shapes separate exactly like the upright ones:
      //   kind 2 (long side along +a: tilted x2/x3/x4):  acc = G(b0) - G(b0 + 2 dy), G a function of b for fixed (a0, dx)
      //   kind 3 (long side along +b: tilted y2/y3/y4):  acc = G(a0) - G(a0 + 2 dx), G a function of a for fixed (b0, dy)
      //   G = T0 - 2 T1 + T2 | -T0 + 3 T1 - 3 T2 + T3 | -T0 + 2 T1 - 2 T3 + T4 at the pattern's 3..5 equidistant points
      // (expand -1*outer + wt*inner of src/haarfeatures.cpp:197-247).  One tile cell step: a += 2 -> off += S + 1,
      // b += 2 -> off += S - 1.
      const bool along_a = pat.kind == 2;
      const int c0 = along_a ? p : p - W;                 // a0, or b0 (stored + W)
      const int LT = ((W + H) >> 1) + 1;
      const int par = along_a ? ((c0 + W) & 1) : (c0 & 1);
      const int step = along_a ? d_ * (S + 1) : d_ * (S - 1);
      for (int t = 0; t < 3; ++t) {
        if (!(flags & (1 << t))) continue;
        const int mmax = t + 2;
        // only the cells the pattern's features read: b in [a0 - 2 hi, a0] (dy <= x), a in [b0 + 2 lo, b0 + 2 W] (x + dx <= W)
        const int jlo = along_a ? max(0, (c0 - 2 * hi + W) >> 1) : max(0, (c0 + 2 * lo) >> 1);
        const int jhi = along_a ? min(LT - 1, (c0 + W) >> 1) : min(LT - 1, (c0 + 2 * W) >> 1);
        for (int j = jlo; j <= jhi; ++j) {
          const int free_c = along_a ? 2 * j - W + par : 2 * j + par;       // b (kind 2) or a (kind 3)
          const int a = along_a ? c0 : free_c, b = along_a ? free_c : c0;
          const int X0 = (a - b) >> 1, Y0 = (a + b) >> 1;
          const int Xe = along_a ? X0 + mmax * d_ : X0 - mmax * d_, Ye = Y0 + mmax * d_;
          int v = 0;
          if (X0 >= 0 && X0 <= W && Y0 >= 0 && Xe >= 0 && Xe <= W && Ye <= H) {
            const int32_t* q = tl + (size_t)(Y0 * S + X0) * TILE;
            const int st = step * TILE;
            if (t == 0) v = q[0] - 2 * q[st] + q[2 * st];
            else if (t == 1) v = -q[0] + 3 * q[st] - 3 * q[2 * st] + q[3 * st];
            else v = -q[0] + 2 * q[st] - 2 * q[3 * st] + q[4 * st];
          }
          sA[j * TILE] = v;
        }
        __syncwarp();
        if (along_a) {
          const int A = W + H;
          for (int x = lo; x <= hi; ++x) {
            const int jb = (c0 - 2 * x + W) >> 1;
            const int gy = sA[jb * TILE];
            const int32_t* sy = sA + jb * TILE;
            const int32_t* tp = tab_t1 + ((((size_t)t * A + c0) * W + (d_ - 1)) * W + x) * HP;
            // catalog bounds (fits iff x+k*dx<=W, y+k*dx+dy<=H, x-dy>=0): dy = 1 .. min(x, H - y - k*dx), y = a0 - x
            const int nd = min(x, H - (c0 - x) - mmax * d_);
            if (nd > 0 && x + mmax * d_ <= W) SEP_EMIT(tp, nd, gy - sy[d * TILE])
          }
        } else {
          const int B = W + H + 1;
          for (int x = lo; x <= hi; ++x) {
            const int ja = (c0 + 2 * x) >> 1;
            const int gy = sA[ja * TILE];
            const int32_t* sy = sA + ja * TILE;
            const int32_t* tp = tab_t2 + ((((size_t)t * B + p) * H + (d_ - 1)) * W + x) * WP;
            // catalog bounds (fits iff x+dx<=W, y+dx+k*dy<=H, x-k*dy>=0): dx = 1 .. min(W - x, H - y - k*dy), y = b0 + x
            const int nd = min(W - x, H - (c0 + x) - mmax * d_);
            if (nd > 0 && x - mmax * d_ >= 0) SEP_EMIT(tp, nd, gy - sy[d * TILE])
          }
        }
        __syncwarp();
      }
    } else {
      // ---------------- column pattern: y = p, dy = d; functions of the column c = 0..W ----------------
      const int y = p, dy = d_;
      auto tabp = [&](int t, int x) { return tab_col + ((((size_t)t * H + y) * H + (dy - 1)) * W + x) * WP; };
      for (int t = 0; t < 3; ++t) {
        if (!(flags & (1 << t))) continue;
        if (y + (t + 2) * dy > H) continue;
        const int32_t* r0 = tl + (size_t)(y * S) * TILE;
        const int32_t* r1 = tl + (size_t)((y + dy) * S) * TILE;
        const int32_t* r2 = tl + (size_t)((y + 2 * dy) * S) * TILE;
        const int32_t* r3 = tl + (size_t)((y + min(3 * dy, H - y)) * S) * TILE;
        const int32_t* r4 = tl + (size_t)((y + min(4 * dy, H - y)) * S) * TILE;
        for (int c = 0; c <= W; ++c) {
          int v;
          if (t == 0) v = r0[c * TILE] - 2 * r1[c * TILE] + r2[c * TILE];
          else if (t == 1) v = r0[c * TILE] - 2 * r1[c * TILE] + 2 * r2[c * TILE] - r3[c * TILE];
          else v = r0[c * TILE] - 2 * r1[c * TILE] + 2 * r3[c * TILE] - r4[c * TILE];
          sA[c * TILE] = v;
        }
        __syncwarp();
        for (int x = lo; x <= hi; ++x) {
          const int hx = sA[x * TILE];
          const int32_t* sx = sA + x * TILE;
          SEP_EMIT(tabp(t, x), W - x, sx[d * TILE] - hx)
        }
        __syncwarp();
      }
    }
  }
}
#undef SEP_EMIT

__global__ void __launch_bounds__(SEP_THREADS, 1)
k_precalc_haar_sep(const int32_t* __restrict__ sum_tiles, const float* __restrict__ nf, const SepPattern* __restrict__ patterns,
                   int npatterns, int patterns_per_chunk, const int32_t* __restrict__ tab_row, const int32_t* __restrict__ tab_col,
                   const int32_t* __restrict__ tab_t1, const int32_t* __restrict__ tab_t2, int W, int H, int row_base,
                   float* __restrict__ fval, int n_stride, int ntiles_used) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = W + 1, P = S * (H + 1);
  const int L = max(W, H) + 1;                                               // scratch column length
  int32_t* stile = reinterpret_cast<int32_t*>(smem_raw);                      // [P][32]
  int32_t* scratch_all = stile + (size_t)P * TILE;                            // [SEP_WARPS][L][32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(scratch_all + (size_t)SEP_WARPS * L * TILE);
  int* s_next = reinterpret_cast<int*>(bar + 1);

  const int tile = blockIdx.x;
  const uint32_t tile_bytes = (uint32_t)P * TILE * sizeof(int32_t);
  if (threadIdx.x == 0) {
    *s_next = 0;
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, tile_bytes);
    bulk_g2s(stile, sum_tiles + (size_t)tile * P * TILE, tile_bytes, bar);
  }
  __syncthreads();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  int32_t* sA = scratch_all + (size_t)warp * L * TILE + lane;                 // element r at sA[r * 32]
  const float nfv = nf[tile * TILE + lane];
  const float rnf = haar_recip(nfv);
  // + f * stride_bytes = this sample's cell of row f
  char* out_bytes = reinterpret_cast<char*>(fval) + ((long long)(tile * TILE + lane) - (long long)row_base * n_stride) * (long long)sizeof(float);
  {   // keep the per-lane base as one opaque 64-bit register so that a cell address is a single IMAD.WIDE
    unsigned long long ob = reinterpret_cast<unsigned long long>(out_bytes);
    asm volatile("" : "+l"(ob));
    out_bytes = reinterpret_cast<char*>(ob);
  }
  const uint32_t stride_bytes = (uint32_t)n_stride * (uint32_t)sizeof(float);
  const int32_t* tl = stile + lane;
  const int p_begin = blockIdx.y * patterns_per_chunk;
  const int p_end = min(p_begin + patterns_per_chunk, npatterns);
  // One correction step of haar_div is exact when 1/nf is correctly rounded, except (Markstein) for divisors whose
  // significand is all ones; a warp holding such a sample takes the two-step variant.
  const bool odd_nf = (__float_as_uint(nfv) & 0x7FFFFFu) == 0x7FFFFFu;
  const bool two_steps = __any_sync(0xffffffffu, odd_nf);
  mbar_wait(bar, 0);
  if (two_steps) sep_body<2>(tl, sA, s_next, p_begin, p_end, patterns, tab_row, tab_col, tab_t1, tab_t2, W, H, out_bytes, stride_bytes, nfv, rnf, lane);
  else sep_body<1>(tl, sA, s_next, p_begin, p_end, patterns, tab_row, tab_col, tab_t1, tab_t2, W, H, out_bytes, stride_bytes, nfv, rnf, lane);
}

// K2 / LBP: same tiling; 16 corner taps -> 9 cell sums -> 8-bit code (lbpfeatures.h:70-83).
constexpr int LBP_THREADS = 512;
template <int TILES_PER_CTA, bool IN_GLOBAL>
__global__ void __launch_bounds__(LBP_THREADS, 1)
k_precalc_lbp(const int32_t* __restrict__ sum_tiles, const LbpTaps* __restrict__ taps, int row_first, int nrows_launch,
              int rows_per_chunk, uint8_t* __restrict__ fcode, int n_stride, int P, int ntiles_used) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tile0 = blockIdx.x * TILES_PER_CTA;
  const int ntl = min(TILES_PER_CTA, ntiles_used - tile0);
  const int32_t* stile = IN_GLOBAL ? sum_tiles + (size_t)tile0 * P * TILE : reinterpret_cast<const int32_t*>(smem_raw);
  if (!IN_GLOBAL) {
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)TILES_PER_CTA * P * TILE * sizeof(int32_t));
    const uint32_t tile_bytes = (uint32_t)P * TILE * sizeof(int32_t);
    if (threadIdx.x == 0) {
      int32_t* dst = reinterpret_cast<int32_t*>(smem_raw);
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(bar, tile_bytes * ntl);
      for (int k = 0; k < ntl; ++k) bulk_g2s(dst + (size_t)k * P * TILE, sum_tiles + (size_t)(tile0 + k) * P * TILE, tile_bytes, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = LBP_THREADS / 32;
  const int r_begin = blockIdx.y * rows_per_chunk;
  const int r_end = min(r_begin + rows_per_chunk, nrows_launch);
  for (int r = r_begin + warp; r < r_end; r += nwarps) {
    const int f = row_first + r;
    const uint4* tp = reinterpret_cast<const uint4*>(taps + f);
    const uint4 a = __ldg(tp), b = __ldg(tp + 1);
    const uint32_t pw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint8_t* out = fcode + (size_t)r * n_stride + tile0 * TILE + lane;
#pragma unroll
    for (int k = 0; k < TILES_PER_CTA; ++k) {
      if (k < ntl) {
        const int32_t* tl = stile + (size_t)k * P * TILE + lane;
        int v[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[2 * q] = tl[(pw[q] & 0xFFFFu) * TILE];
          v[2 * q + 1] = tl[(pw[q] >> 16) * TILE];
        }
        auto cell = [&](int gx, int gy) { const int i = gy * 4 + gx; return v[i] - v[i + 1] - v[i + 4] + v[i + 5]; };
        const int c = cell(1, 1);
        const int code = (cell(0, 0) >= c ? 128 : 0) | (cell(1, 0) >= c ? 64 : 0) | (cell(2, 0) >= c ? 32 : 0) |
                         (cell(2, 1) >= c ? 16 : 0) | (cell(2, 2) >= c ? 8 : 0) | (cell(1, 2) >= c ? 4 : 0) |
                         (cell(0, 2) >= c ? 2 : 0) | (cell(0, 1) >= c ? 1 : 0);
        out[k * TILE] = (uint8_t)code;
      }
    }
  }
}

// K2 / HOG: planes are [plane][P][ns_pad] (sample-fastest), so lane == sample gives coalesced 128-byte
// loads straight from L2/HBM; one warp handles one block feature (36 variables) of one sample tile.
__global__ void __launch_bounds__(256)
k_precalc_hog(const float* __restrict__ planes, const HogTaps* __restrict__ taps, int var_first, int nvars, float* __restrict__ fval,
              int n_stride, int P, int ns_pad, int ntiles_used) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps_total = (gridDim.x * blockDim.x) >> 5;
  const int f_first = var_first / 36, f_last = (var_first + nvars - 1) / 36;
  const int nblocks = f_last - f_first + 1;
  const long long work = (long long)nblocks * ntiles_used;
  for (long long wi = warp_global; wi < work; wi += nwarps_total) {
    const int fb = f_first + (int)(wi / ntiles_used);
    const int s = (int)(wi % ntiles_used) * TILE + lane;
    const HogTaps t = taps[fb];
    const float* pn = planes + (size_t)9 * P * ns_pad + s;
    const float nrm = __fadd_rn(__fsub_rn(__fsub_rn(pn[(size_t)t.off[0][0] * ns_pad], pn[(size_t)t.off[1][1] * ns_pad]),
                                          pn[(size_t)t.off[2][2] * ns_pad]), pn[(size_t)t.off[3][3] * ns_pad]);
    const float den = __fadd_rn(nrm, 0.001f);
    for (int comp = 0; comp < 36; ++comp) {
      const int var = fb * 36 + comp;
      if (var < var_first || var >= var_first + nvars) continue;
      const int bin = comp % 9, cell = comp / 9;
      const float* ph = planes + (size_t)bin * P * ns_pad + s;
      const float res = __fadd_rn(__fsub_rn(__fsub_rn(ph[(size_t)t.off[cell][0] * ns_pad], ph[(size_t)t.off[cell][1] * ns_pad]),
                                            ph[(size_t)t.off[cell][2] * ns_pad]), ph[(size_t)t.off[cell][3] * ns_pad]);
      fval[(size_t)(var - var_first) * n_stride + s] = res > 0.001f ? __fdiv_rn(res, den) : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3: per-feature segmented sort.  One CTA sorts one row of N (value, sample index) pairs in shared memory: stable
// LSD radix sort over the order-preserving float->uint transform, 8 passes of 4 bits.
//
// What moves through the passes is one 32-bit word per element, (16 key bits << 16) | sample index: passes 0-3 sort
// by the low half of the 32-bit key, then every element swaps its key half for the high half (one gather from the
// row's key array, which stays put in shared memory) and passes 4-7 finish the sort -- stable passes keep the order
// of the first four inside equal high halves.  A thread owns 15 consecutive positions and keeps them in registers
// from the load to the scatter of a pass, so the pass is in place; its digit histogram and its running ranks are
// 4-bit fields of two 32-bit registers, so the only shared-memory traffic of a pass is: 15 loads, 16 counter stores,
// the block scan of the [digit][thread] counters, 15 base loads and 15 scattered stores per thread.  Stability: the
// scan order is (digit, thread), a thread walks its positions in order.  Ties therefore come out in ascending
// sample index, which is the canonical order the oracle's stable mode produces (std::sort + LessThanIdx in the
// reference leaves tie order unspecified, o_utils.h:18-23, test_o_utils.cpp:111-128).
// ------------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 1024;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_KPT_MAX = 15;                      // positions per thread: counts fit 4-bit fields, odd stride = conflict-free
constexpr int SORT_CAP = SORT_THREADS * SORT_KPT_MAX; // 15360 elements per row (smaller rows use 3 / 7 / 11 positions per thread)
constexpr int SORT_CNT_WORDS = 16 * SORT_THREADS + 16 * SORT_THREADS / 32;     // [digit][thread] u32, one pad word per 32: both access patterns conflict-free
static size_t sort_smem(int kpt) { return (size_t)SORT_THREADS * kpt * 4 * 2 + (size_t)SORT_CNT_WORDS * 4; }
__device__ __forceinline__ int sort_cnt_slot(int idx) { return idx + (idx >> 5); }

__device__ __forceinline__ uint32_t float_to_key(float v) {
  uint32_t u = __float_as_uint(v + 0.0f);          // -0 -> +0
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

template <int SORT_KPT>
__global__ void __launch_bounds__(SORT_THREADS, 1)
k_sort_rows(const float* __restrict__ fval, int n_stride, int N, int nrows, uint16_t* __restrict__ sidx, float* __restrict__ sval,
            const int* __restrict__ row_list, const unsigned int* __restrict__ row_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CAP = SORT_THREADS * SORT_KPT;
  if (row_list) nrows = (int)*row_count;               // second pass of the bucket sort: only the rows it deferred
  uint32_t* fullkey = reinterpret_cast<uint32_t*>(smem_raw);               // [CAP] by sample index
  uint32_t* pk = fullkey + CAP;                                             // [CAP] packed (key half << 16) | index, by position
  uint32_t* cnt = pk + CAP;                                            // [16][SORT_THREADS], padded (sort_cnt_slot)
  __shared__ uint32_t warp_tot[SORT_WARPS];

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  uint32_t* mine = pk + t * SORT_KPT;

  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const int row = row_list ? row_list[r] : r;
    const float* src = fval + (size_t)row * n_stride;
    __syncthreads();                                   // the previous row's output gathers are done
    for (int i = t; i < CAP; i += SORT_THREADS) {
      const uint32_t k = i < N ? float_to_key(__ldcs(src + i)) : 0xFFFFFFFFu;     // padding sorts behind every real element
      fullkey[i] = k;
      pk[i] = (k << 16) | (uint32_t)i;
    }
    __syncthreads();
#pragma unroll 1
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 16 + 4 * (pass & 3);
      uint32_t e[SORT_KPT];
      uint32_t hlo = 0, hhi = 0;                       // per-digit counts of this thread, 4-bit fields (digits 0-7 | 8-15)
      const uint32_t sh_field = (uint32_t)(shift - 2), top_bit = 1u << (shift + 3);
#pragma unroll
      for (int k = 0; k < SORT_KPT; ++k) {
        uint32_t w = mine[k];
        if (pass == 4) w = (fullkey[w & 0xFFFFu] & 0xFFFF0000u) | (w & 0xFFFFu);     // switch to the high key half
        e[k] = w;
        // field offset = (digit & 7) * 4 straight from the word; digit bit 3 picks the register (predicated adds)
        asm("{\n\t.reg .pred p;\n\t.reg .b32 s, o;\n\t"
            "shr.b32 s, %2, %3;\n\tand.b32 s, s, 28;\n\tand.b32 o, %2, %4;\n\tsetp.ne.b32 p, o, 0;\n\tshl.b32 o, 1, s;\n\t"
            "@p add.u32 %1, %1, o;\n\t@!p add.u32 %0, %0, o;\n\t}"
            : "+r"(hlo), "+r"(hhi) : "r"(w), "r"(sh_field), "r"(top_bit));
      }
      uint32_t* col = cnt + t + (t >> 5);              // entry (d, t) at col[d * (SORT_THREADS + 32)]
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        col[d * (SORT_THREADS + 32)] = (hlo >> (4 * d)) & 15u;
        col[(d + 8) * (SORT_THREADS + 32)] = (hhi >> (4 * d)) & 15u;
      }
      __syncthreads();
      // exclusive scan of the counters in (digit, thread) order: thread t owns flattened entries [16t, 16t+16)
      {
        uint32_t* c16 = cnt + sort_cnt_slot(16 * t);   // 16 consecutive entries never straddle a pad word (16t is a multiple of 16)
        uint32_t w[16];
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) { w[k] = c16[k]; sum += w[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += x; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t wt = warp_tot[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, wt, d); if (lane >= d) wt += x; }
        const uint32_t wbase = __shfl_sync(0xffffffffu, wt, (warp + 31) & 31);
        uint32_t run = (warp ? wbase : 0u) + incl - sum;
#pragma unroll
        for (int k = 0; k < 16; ++k) { const uint32_t c = w[k]; c16[k] = run; run += c; }
      }
      __syncthreads();
      // scatter in place: every thread holds its elements in registers; its column of the counter array now holds the
      // first free slot per digit and is advanced as the thread walks its elements in order (only thread t touches it)
      {
        const uint32_t sh_d = (uint32_t)shift;
#pragma unroll
        for (int k = 0; k < SORT_KPT; ++k) {
          const uint32_t w = e[k];
          uint32_t* slot = col + ((w >> sh_d) & 15u) * (SORT_THREADS + 32);
          const uint32_t pos = *slot;
          *slot = pos + 1;
          pk[pos] = w;
        }
      }
      __syncthreads();
    }
    // output: sorted sample indices and the values in sorted order
    uint16_t* drow = sidx + (size_t)row * n_stride;
    float* vrow = sval + (size_t)row * n_stride;
    for (int i = t; i < N; i += SORT_THREADS) {
      const uint32_t id = pk[i] & 0xFFFFu;
      drow[i] = (uint16_t)id;
      __stcs(vrow + i, key_to_float(fullkey[id]));
    }
    // row padding up to the split-search tile: index N (an always-zero operand), value +max
    for (int i = N + t; i < n_stride; i += SORT_THREADS) { drow[i] = (uint16_t)N; vrow[i] = 3.402823466e+38f; }
  }
}


// ------------------------------------------------------------------------------------------------
// K3, first pass for Haar rows: bucket + rank sort.  The canonical order (value ascending, ties by ascending sample
// index) is the order of the 48-bit composites (key << 16) | index, which are all distinct -- so nothing has to be
// stable, and one distribution pass into many value buckets followed by an exact ranking inside each bucket replaces
// the eight stable radix passes:
//   A  load the row (registers), row statistics -> a monotone linear bucket map  b(v) = clamp((v - lo) * scale)
//      over [mean - 4.5 sd, mean + 4.5 sd] ∩ [min, max]  (any lo / scale gives a correct sort; float subtraction,
//      multiplication by a positive constant and truncation are monotone, so b is non-decreasing in v);
//   B  bucket histogram (shared-memory atomics);   C  exclusive scan of the B counters + the cost sum(k^2);
//   D  distribution: slot = atomicAdd(cursor[b]), composite stored there (order inside a bucket is arbitrary);
//   E  rank: position = bucket start + #(composites of the bucket smaller than mine); the index and the value go
//      straight to their final place in HBM (a warp writes a permutation of 32 consecutive positions).
// With B ~ N / 2 buckets a row costs ~ N * (k_avg + const) compares instead of 8 passes over N.  Rows where the map
// is poor (sum(k^2) above a limit: long runs of equal values, outliers far beyond 4.5 sd) are appended to a list and
// sorted by k_sort_rows afterwards.  Output is bit-identical to k_sort_rows (test_sort_paths_agree).
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long BSORT_ONE = 0x3FF0000000000000ull;     // composites are stored as doubles in [1, 2)
template <int KPT> struct BucketCfg { static constexpr int B = KPT <= 3 ? 4096 : KPT <= 7 ? 8192 : 16384; };     // a multiple of 4 * SORT_THREADS (uint4 scan)
template <int KPT> static size_t bsort_smem() { return (size_t)(SORT_THREADS * KPT + 8) * 8 + (size_t)BucketCfg<KPT>::B * 4; }

template <int KPT>
__global__ void __launch_bounds__(SORT_THREADS, 1)
k_sort_rows_bucket(const float* __restrict__ fval, int n_stride, int N, int nrows, uint16_t* __restrict__ sidx, float* __restrict__ sval,
                   int* __restrict__ hard_rows, unsigned int* __restrict__ hard_count, float hard_limit) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CAP = SORT_THREADS * KPT;
  constexpr int B = BucketCfg<KPT>::B;
  constexpr int CPT = B / SORT_THREADS;                                   // counters per thread in the scan
  unsigned long long* comp = reinterpret_cast<unsigned long long*>(smem_raw);     // [CAP] composites, by bucket after phase D
  uint32_t* cnt = reinterpret_cast<uint32_t*>(comp + CAP + 8);                     // [B] counts -> cursors -> bucket ends
  __shared__ float red[4][SORT_WARPS];
  __shared__ uint32_t warp_tot[SORT_WARPS];
  __shared__ float warp_cost[SORT_WARPS];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  uint32_t comp_s = (uint32_t)__cvta_generic_to_shared(comp), cnt_s = (uint32_t)__cvta_generic_to_shared(cnt);
  asm volatile("" : "+r"(comp_s), "+r"(cnt_s));          // opaque: kept in registers instead of being re-derived from SR_CgaCtaId in the loops

  // every position k < KPT - 1 of a thread is a real element when N covers KPT - 1 full strides (uniform)
  const bool full = N >= (KPT - 1) * SORT_THREADS;
  auto valid = [&](int k) -> bool { return (k < KPT - 1 && full) || t + k * SORT_THREADS < N; };
  float v[KPT];
  auto load_row = [&](int row) {
    const float* src = fval + (size_t)row * n_stride;
#pragma unroll
    for (int k = 0; k < KPT; ++k) v[k] = valid(k) ? __ldcs(src + t + k * SORT_THREADS) : 0.f;
  };
  if ((int)blockIdx.x < nrows) load_row(blockIdx.x);
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    // ---- A: statistics of the row loaded during the previous row's ranking ----
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f, sm = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      v[k] += 0.0f;                                                         // -0 -> +0 (as float_to_key)
      if (valid(k)) { mn = fminf(mn, v[k]); mx = fmaxf(mx, v[k]); sm += v[k]; sq = fmaf(v[k], v[k], sq); }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      sm += __shfl_xor_sync(0xffffffffu, sm, d); sq += __shfl_xor_sync(0xffffffffu, sq, d);
    }
    __syncthreads();                                   // the previous row's ranking has finished reading comp / cnt
    if (lane == 0) { red[0][warp] = mn; red[1][warp] = mx; red[2][warp] = sm; red[3][warp] = sq; }
    {
      uint4* c4 = reinterpret_cast<uint4*>(cnt);
      for (int i = t; i < B / 4; i += SORT_THREADS) c4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    mn = red[0][lane]; mx = red[1][lane]; sm = red[2][lane]; sq = red[3][lane];
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      sm += __shfl_xor_sync(0xffffffffu, sm, d); sq += __shfl_xor_sync(0xffffffffu, sq, d);
    }
    float lo, scale;
    {
      const float inv_n = 1.0f / (float)N;
      const float mean = sm * inv_n;
      const float var = fmaxf(sq * inv_n - mean * mean, 0.f);
      const float sd = sqrtf(var);
      float a = mean - 4.5f * sd, b = mean + 4.5f * sd;
      if (!(a > mn)) a = mn;                           // also taken when the statistics overflowed (inf / nan)
      if (!(b < mx)) b = mx;
      const float w = b - a;
      lo = a;
      scale = (w > 0.f && w < 3.0e38f) ? (float)B / w : 0.f;
      if (!(scale < 3.0e38f)) scale = 0.f;
    }
    auto bucket_of = [&](float x) -> int {
      const int b = __float2int_rz((x - lo) * scale);
      return min(max(b, 0), B - 1);
    };
    // ---- B: histogram ----
#pragma unroll
    for (int k = 0; k < KPT; ++k)
      if (valid(k)) atomicAdd(&cnt[bucket_of(v[k])], 1u);
    __syncthreads();
    // ---- C: exclusive scan of the counters, cost = sum k^2.  Warp w owns the CPT * 32 consecutive counters behind
    // cnt + w * CPT * 32, as Q = CPT / 4 groups of 128; lane l holds counters [4l, 4l + 4) of every group (consecutive
    // lanes read consecutive 16-byte words: conflict-free). ----
    float cost;
    {
      constexpr int Q = CPT / 4;
      uint4* c4 = reinterpret_cast<uint4*>(cnt + warp * (CPT * 32)) + lane;
      uint4 g[Q];
      uint32_t incl[Q], gs[Q], base[Q];
      float c2 = 0.f;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        g[q] = c4[q * 32];
        gs[q] = g[q].x + g[q].y + g[q].z + g[q].w;
        incl[q] = gs[q];
        c2 = fmaf((float)g[q].x, (float)g[q].x, c2); c2 = fmaf((float)g[q].y, (float)g[q].y, c2);
        c2 = fmaf((float)g[q].z, (float)g[q].z, c2); c2 = fmaf((float)g[q].w, (float)g[q].w, c2);
      }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int q = 0; q < Q; ++q) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl[q], d); if (lane >= d) incl[q] += x; }
      }
      uint32_t wsum = 0;
#pragma unroll
      for (int q = 0; q < Q; ++q) { base[q] = wsum; wsum += __shfl_sync(0xffffffffu, incl[q], 31); }
#pragma unroll
      for (int d = 16; d; d >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, d);
      if (lane == 0) { warp_tot[warp] = wsum; warp_cost[warp] = c2; }
      __syncthreads();
      uint32_t wt = warp_tot[lane];
      cost = warp_cost[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, wt, d); if (lane >= d) wt += x; }
#pragma unroll
      for (int d = 16; d; d >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, d);
      const uint32_t wbase = warp ? __shfl_sync(0xffffffffu, wt, (warp + 31) & 31) : 0u;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        uint32_t run = wbase + base[q] + incl[q] - gs[q];
        uint4 x;
        x.x = run; run += g[q].x; x.y = run; run += g[q].y; x.z = run; run += g[q].z; x.w = run;
        c4[q * 32] = x;
      }
    }
    if (cost > hard_limit) {                           // uniform over the CTA: defer the row to the radix kernel
      if (t == 0) hard_rows[atomicAdd(hard_count, 1u)] = row;
      if (row + (int)gridDim.x < nrows) load_row(row + gridDim.x);
      continue;
    }
    __syncthreads();
    // ---- D: distribution ----
    if (t < 8) comp[N + t] = 0x4000000000000000ull;      // sentinel behind the last bucket: 2.0, above every composite
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      if (valid(k)) {
        const uint32_t p = atomicAdd(&cnt[bucket_of(v[k])], 1u);
        comp[p] = BSORT_ONE | ((unsigned long long)float_to_key(v[k]) << 16) | (unsigned long long)(t + k * SORT_THREADS);
      }
    }
    if (row + (int)gridDim.x < nrows) load_row(row + gridDim.x);      // the next row's loads fly during the ranking
    __syncthreads();
    // ---- E: rank, write out.  The buckets are in value order, so every composite left of my bucket is smaller than
    // mine and every one right of it is larger: my final position is w0 + #(smaller composites in [w0, w1)) for ANY
    // window [w0, w1) that covers the bucket -- the window is widened to 16-byte alignment and whole groups of four
    // (comp[N .. N+8) holds a sentinel), so there is neither a remainder loop nor a mask.  Composites carry the
    // exponent of 1.0, i.e. they are doubles in [1, 2) whose order is the order of the 48-bit integers: one DSETP per
    // comparison on the otherwise idle FP64 pipe.  Shared memory is addressed through 32-bit shared-window addresses. ----
    uint16_t* drow = sidx + (size_t)row * n_stride;
    float* vrow = sval + (size_t)row * n_stride;
#pragma unroll 1
    for (int p = t; p < N; p += SORT_THREADS) {
      double cd;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cd) : "r"(comp_s + 8u * (uint32_t)p));
      const unsigned long long c = (unsigned long long)__double_as_longlong(cd);
      const float x = key_to_float((uint32_t)(c >> 16));
      const int b = bucket_of(x);
      uint32_t e, j = 0;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(cnt_s + 4u * (uint32_t)b));
      if (b) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(j) : "r"(cnt_s + 4u * (uint32_t)b - 4u));
      j &= ~1u;
      uint32_t r = j;
#pragma unroll 1
      for (; j < e; j += 4) {
        double q0, q1, q2, q3;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q0), "=d"(q1) : "r"(comp_s + 8u * j));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q2), "=d"(q3) : "r"(comp_s + 8u * j + 16u));
        asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
            "setp.lt.f64 p0, %1, %5;\n\tsetp.lt.f64 p1, %2, %5;\n\tsetp.lt.f64 p2, %3, %5;\n\tsetp.lt.f64 p3, %4, %5;\n\t"
            "@p0 add.u32 %0, %0, 1;\n\t@p1 add.u32 %0, %0, 1;\n\t@p2 add.u32 %0, %0, 1;\n\t@p3 add.u32 %0, %0, 1;\n\t}"
            : "+r"(r) : "d"(q0), "d"(q1), "d"(q2), "d"(q3), "d"(cd));
      }
      drow[r] = (uint16_t)(c & 0xFFFFull);
      __stcs(vrow + r, x);
    }
    // row padding up to the split-search tile: index N (an always-zero operand), value +max
    for (int i = N + t; i < n_stride; i += SORT_THREADS) { drow[i] = (uint16_t)N; vrow[i] = 3.402823466e+38f; }
  }
}

// ------------------------------------------------------------------------------------------------
// K3, wide path: rows longer than one CTA can hold in shared memory (numSamples > 15360, including the reference's
// is_buf_16u == false case numSamples >= 65536).  Stable LSD radix sort, 4 passes of 8 bits, one CTA per row, keys and
// indices ping-pong between two global scratch rows per CTA (L2-resident).  A pass = digit histogram, exclusive scan, then a
// stable scatter chunk by chunk in position order: inside a chunk of 1024 positions a thread's rank among the equal digits
// is (equal digits in earlier warps) + (equal digits in earlier lanes of its warp, by __match_any_sync).  Ties therefore come out
// in ascending sample index, the canonical order of the other sort kernels.  Output: 32-bit sorted-index rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS, 1)
k_sort_rows_wide(const float* __restrict__ fval, int n_stride, int N, int nrows, uint32_t* __restrict__ sidx32, uint32_t* __restrict__ scratch) {
  __shared__ uint32_t s_base[256];                 // running first free slot per digit
  __shared__ uint32_t s_wcnt[SORT_WARPS][256];     // per chunk: equal digits per warp -> exclusive offsets over the warps
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const size_t npad = (size_t)n_stride;
  uint32_t* kbuf[2] = {scratch + (size_t)blockIdx.x * 4 * npad, scratch + (size_t)blockIdx.x * 4 * npad + npad};
  uint32_t* ibuf[2] = {scratch + (size_t)blockIdx.x * 4 * npad + 2 * npad, scratch + (size_t)blockIdx.x * 4 * npad + 3 * npad};
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    const float* src = fval + (size_t)row * n_stride;
    for (int i = t; i < N; i += SORT_THREADS) { kbuf[0][i] = float_to_key(src[i]); ibuf[0][i] = (uint32_t)i; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
      const uint32_t* kin = kbuf[pass & 1];
      const uint32_t* iin = ibuf[pass & 1];
      uint32_t* kout = kbuf[(pass & 1) ^ 1];
      uint32_t* iout = ibuf[(pass & 1) ^ 1];
      const int shift = 8 * pass;
      if (t < 256) s_base[t] = 0;
      __syncthreads();
      for (int i = t; i < N; i += SORT_THREADS) atomicAdd(&s_base[(kin[i] >> shift) & 255u], 1u);
      __syncthreads();
      if (warp == 0) {                               // exclusive scan of the 256 counters
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = s_base[lane * 8 + k]; sum += c[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += x; }
        uint32_t run = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) { s_base[lane * 8 + k] = run; run += c[k]; }
      }
      __syncthreads();
      for (int c0 = 0; c0 < N; c0 += SORT_THREADS) {
        for (int i = t; i < SORT_WARPS * 256; i += SORT_THREADS) (&s_wcnt[0][0])[i] = 0;
        __syncthreads();
        const int i = c0 + t;
        const bool on = i < N;
        const uint32_t key = on ? kin[i] : 0u, id = on ? iin[i] : 0u;
        const uint32_t d = on ? ((key >> shift) & 255u) : 256u;          // 256: not a digit, matches only other idle lanes
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (on && rank_in_warp == 0) s_wcnt[warp][d] = __popc(peers);
        __syncthreads();
        if (t < 256) {                                 // exclusive offsets over the warps, and the digit's base moves on
          uint32_t run = s_base[t];
          for (int w = 0; w < SORT_WARPS; ++w) { const uint32_t c = s_wcnt[w][t]; s_wcnt[w][t] = run; run += c; }
          s_base[t] = run;
        }
        __syncthreads();
        if (on) { const uint32_t p = s_wcnt[warp][d] + rank_in_warp; kout[p] = key; iout[p] = id; }
        __syncthreads();
      }
    }
    // 4 passes: the result is back in buffer 0
    uint32_t* drow = sidx32 + (size_t)row * n_stride;
    for (int i = t; i < n_stride; i += SORT_THREADS) drow[i] = i < N ? ibuf[0][i] : (uint32_t)N;
    __syncthreads();
  }
}

__global__ void k_div_check(int steps, unsigned long long seed, long long n, unsigned long long* mismatches) {
  unsigned long long bad = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // splitmix64
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    const int acc = (int)(z & 0x3FFFFF) - 0x200000;                       // [-2^21, 2^21)
    // nf: positive float with exponent in [2^0, 2^23) and arbitrary mantissa; every 64th sample forces special mantissas
    unsigned mant = (unsigned)(z >> 22) & 0x7FFFFFu;
    const unsigned ex = 127u + (unsigned)((z >> 45) % 23);
    if ((i & 63) == 0) mant = 0x7FFFFFu - (unsigned)((z >> 50) & 7u);
    if ((i & 63) == 1) mant = (unsigned)((z >> 50) & 7u);
    const float nfv = __uint_as_float((ex << 23) | mant);
    const float r = haar_recip(nfv);
    const float ref = __fdiv_rn((float)acc, nfv);
    const float got = steps == 1 ? haar_div<1>((float)acc, nfv, r) : haar_div<2>((float)acc, nfv, r);
    bad += __float_as_uint(ref + 0.0f) != __float_as_uint(got + 0.0f);
  }
  if (bad) atomicAdd(mismatches, bad);
}

// widen LBP codes / copy rows for the parity getters
__global__ void k_codes_to_float(const uint8_t* __restrict__ codes, int n_stride, int N, int nrows, float* __restrict__ out) {
  const long long total = (long long)nrows * N;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / N), s = (int)(t - (long long)r * N);
    out[t] = (float)codes[(size_t)r * n_stride + s];
  }
}

int traindata_free(cc_ctx* ctx) {
  cudaFree(ctx->d_fval); cudaFree(ctx->d_fcode); cudaFree(ctx->d_sidx); cudaFree(ctx->d_sval); cudaFree(ctx->d_hard_rows); cudaFree(ctx->d_sidx32);
  ctx->d_fval = nullptr; ctx->d_fcode = nullptr; ctx->d_sidx = nullptr; ctx->d_sval = nullptr; ctx->d_hard_rows = nullptr; ctx->d_sidx32 = nullptr;
  ctx->cap_rows = ctx->cap_stride = 0;
  ctx->precalc_valid = false;
  return CC_OK;
}

static int ensure_traindata(cc_ctx* ctx, int nrows, int n_stride, bool wide) {
  if ((size_t)nrows <= ctx->cap_rows && (size_t)n_stride <= ctx->cap_stride && wide == ctx->cap_wide) return CC_OK;
  traindata_free(ctx);
  const size_t cells = (size_t)nrows * n_stride;
  ctx->cap_wide = wide;
  if (ctx->ftype == CC_FEATURE_LBP) {
    CC_CUDA(cudaMalloc(&ctx->d_fcode, cells));
  } else if (wide) {
    // CvCascadeBoostTrainData::setData with is_buf_16u == false (src/o_cvcascadeboosttraindata.cpp:250-251,283-309): 32-bit index
    // rows.  Sorted values are not kept: the wide search gathers them from the fval row
    CC_CUDA(cudaMalloc(&ctx->d_fval, cells * sizeof(float)));
    CC_CUDA(cudaMalloc(&ctx->d_sidx32, cells * sizeof(uint32_t)));
  } else {
    CC_CUDA(cudaMalloc(&ctx->d_fval, cells * sizeof(float)));
    CC_CUDA(cudaMalloc(&ctx->d_sidx, cells * sizeof(uint16_t)));
    CC_CUDA(cudaMalloc(&ctx->d_sval, cells * sizeof(float)));
    CC_CUDA(cudaMalloc(&ctx->d_hard_rows, ((size_t)nrows + 1) * sizeof(int)));      // [0] = count, [1..] = rows deferred by the bucket sort
  }
  ctx->cap_rows = nrows; ctx->cap_stride = n_stride;
  return CC_OK;
}

}  // namespace cc

using namespace cc;

extern "C" {

int cc_boost_set_shard(cc_ctx* ctx, int rank, int world) {
  CC_REQUIRE(ctx, CC_ERR_BAD_ARG, "cc_boost_set_shard: null ctx");
  CC_REQUIRE(world >= 1 && rank >= 0 && rank < world, CC_ERR_BAD_ARG, "cc_boost_set_shard: rank %d of %d", rank, world);
  ctx->shard_rank = rank; ctx->shard_world = world;
  if (ctx->ftype >= 0) {
    const long long F = ctx->var_count;
    ctx->row0 = (int)(F * rank / world);
    ctx->nrows = (int)(F * (rank + 1) / world) - ctx->row0;
    ctx->precalc_valid = false;
    // per-plane compact tap lists of this shard (Haar): upright rows read `sum`, tilted rows read `tilted`;
    // word 11 of every entry carries the global feature index
    if (ctx->ftype == CC_FEATURE_HAAR) {
      std::vector<int> up, tl;
      std::vector<HaarTaps> tu, tt;
      const int S = ctx->W + 1;
      for (int f = ctx->row0; f < ctx->row0 + ctx->nrows; ++f) {
        HaarTaps t = make_haar_taps(ctx->haar[f], S);
        t.w[11] = (uint32_t)f;
        if (ctx->haar[f].tilted) { tl.push_back(f); tt.push_back(t); } else { up.push_back(f); tu.push_back(t); }
      }
      ctx->upright_rows = up; ctx->tilted_rows = tl;
      cudaFree(ctx->d_taps_upright); cudaFree(ctx->d_taps_tilted);
      ctx->d_taps_upright = ctx->d_taps_tilted = nullptr;
      CC_CUDA(cudaSetDevice(ctx->device));
      if (!tu.empty()) {
        CC_CUDA(cudaMalloc(&ctx->d_taps_upright, tu.size() * sizeof(HaarTaps)));
        CC_CUDA(cudaMemcpy(ctx->d_taps_upright, tu.data(), tu.size() * sizeof(HaarTaps), cudaMemcpyHostToDevice));
      }
      if (!tt.empty()) {
        CC_CUDA(cudaMalloc(&ctx->d_taps_tilted, tt.size() * sizeof(HaarTaps)));
        CC_CUDA(cudaMemcpy(ctx->d_taps_tilted, tt.data(), tt.size() * sizeof(HaarTaps), cudaMemcpyHostToDevice));
      }
      // ---- separable upright precalc: index tables + pattern list (only if one tile + scratch fit in smem) ----
      cudaFree(ctx->d_sep_tab_row); cudaFree(ctx->d_sep_tab_col); cudaFree(ctx->d_sep_patterns); cudaFree(ctx->d_sep_tab_t1); cudaFree(ctx->d_sep_tab_t2);
      ctx->d_sep_tab_row = ctx->d_sep_tab_col = ctx->d_sep_tab_t1 = ctx->d_sep_tab_t2 = nullptr;
      ctx->d_sep_patterns = nullptr; ctx->n_sep_patterns = 0; ctx->n_sep_tilted = 0; ctx->use_sep = false;
      const int W = ctx->W, H = ctx->H;
      const size_t sep_smem = (size_t)ctx->P * TILE * 4 + (size_t)SEP_WARPS * (std::max(W, H) + 1) * TILE * 4 + 64;
      if (sep_smem <= 227 * 1024 && W + H <= 255 && (!up.empty() || !tl.empty())) {
        const int HP = (H + 3) & ~3, WP = (W + 3) & ~3;
        const int A = W + H, B = W + H + 1;
        std::vector<int32_t> trow((size_t)5 * W * W * H * HP, -1), tcol((size_t)3 * H * H * W * WP, -1);
        std::vector<int32_t> tt1(tl.empty() ? 4 : (size_t)3 * A * W * W * HP, -1), tt2(tl.empty() ? 4 : (size_t)3 * B * H * W * WP, -1);
        // per pattern: which shapes occur and the range of the outer loop variable that has local features
        struct Acc { int flags = 0, lo = 1 << 30, hi = -1; };
        std::vector<Acc> prow((size_t)W * W), pcol((size_t)H * H), pt1((size_t)A * W), pt2((size_t)B * H);
        static const int row_t[8] = {0, -1, 1, -1, 2, -1, 3, 4};   // shape -> row-table type
        static const int col_t[8] = {-1, 0, -1, 1, -1, 2, -1, -1}; // shape -> column-table type
        for (int f : up) {
          const HaarFeatureH& hf = ctx->haar[f];
          if (row_t[hf.shape] >= 0) {
            const int t = row_t[hf.shape];
            trow[((((size_t)t * W + hf.x) * W + (hf.dx - 1)) * H + hf.y) * HP + (hf.dy - 1)] = f;
            Acc& a = prow[(size_t)hf.x * W + (hf.dx - 1)];
            a.flags |= 1 << t; a.lo = std::min(a.lo, hf.y); a.hi = std::max(a.hi, hf.y);
          } else {
            const int t = col_t[hf.shape];
            tcol[((((size_t)t * H + hf.y) * H + (hf.dy - 1)) * W + hf.x) * WP + (hf.dx - 1)] = f;
            Acc& a = pcol[(size_t)hf.y * H + (hf.dy - 1)];
            a.flags |= 1 << t; a.lo = std::min(a.lo, hf.x); a.hi = std::max(a.hi, hf.x);
          }
        }
        for (int f : tl) {       // tilted shapes 8..13: even = long side along a (x2, x3, x4), odd = along b (y2, y3, y4)
          const HaarFeatureH& hf = ctx->haar[f];
          const int t = (hf.shape - 8) >> 1;
          if (((hf.shape - 8) & 1) == 0) {
            const int a0 = hf.x + hf.y;
            tt1[((((size_t)t * A + a0) * W + (hf.dx - 1)) * W + hf.x) * HP + (hf.dy - 1)] = f;
            Acc& a = pt1[(size_t)a0 * W + (hf.dx - 1)];
            a.flags |= 1 << t; a.lo = std::min(a.lo, hf.x); a.hi = std::max(a.hi, hf.x);
          } else {
            const int b0w = hf.y - hf.x + W;
            tt2[((((size_t)t * B + b0w) * H + (hf.dy - 1)) * W + hf.x) * WP + (hf.dx - 1)] = f;
            Acc& a = pt2[(size_t)b0w * H + (hf.dy - 1)];
            a.flags |= 1 << t; a.lo = std::min(a.lo, hf.x); a.hi = std::max(a.hi, hf.x);
          }
        }
        std::vector<SepPattern> pats, patt;
        for (int x = 0; x < W; ++x)
          for (int dx = 1; dx <= W; ++dx) {
            const Acc& a = prow[(size_t)x * W + (dx - 1)];
            if (a.flags) pats.push_back({0, (uint8_t)x, (uint8_t)dx, (uint8_t)a.flags, (uint8_t)a.lo, (uint8_t)a.hi, 0, 0});
          }
        for (int y = 0; y < H; ++y)
          for (int dy = 1; dy <= H; ++dy) {
            const Acc& a = pcol[(size_t)y * H + (dy - 1)];
            if (a.flags) pats.push_back({1, (uint8_t)y, (uint8_t)dy, (uint8_t)a.flags, (uint8_t)a.lo, (uint8_t)a.hi, 0, 0});
          }
        for (int a0 = 0; a0 < A; ++a0)
          for (int dx = 1; dx <= W; ++dx) {
            const Acc& a = pt1[(size_t)a0 * W + (dx - 1)];
            if (a.flags) patt.push_back({2, (uint8_t)a0, (uint8_t)dx, (uint8_t)a.flags, (uint8_t)a.lo, (uint8_t)a.hi, 0, 0});
          }
        for (int b0w = 0; b0w < B; ++b0w)
          for (int dy = 1; dy <= H; ++dy) {
            const Acc& a = pt2[(size_t)b0w * H + (dy - 1)];
            if (a.flags) patt.push_back({3, (uint8_t)b0w, (uint8_t)dy, (uint8_t)a.flags, (uint8_t)a.lo, (uint8_t)a.hi, 0, 0});
          }
        // heavy patterns (small d: most features) first, so that the dynamic per-CTA queue ends with light ones
        auto by_d = [](const SepPattern& a, const SepPattern& b) { return a.d < b.d; };
        std::stable_sort(pats.begin(), pats.end(), by_d);
        std::stable_sort(patt.begin(), patt.end(), by_d);
        ctx->n_sep_patterns = (int)pats.size();
        ctx->n_sep_tilted = (int)patt.size();
        pats.insert(pats.end(), patt.begin(), patt.end());
        auto upload = [&](int32_t** dst, const std::vector<int32_t>& v) -> int {
          CC_CUDA(cudaMalloc(dst, v.size() * sizeof(int32_t)));
          CC_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
          return CC_OK;
        };
        CC_CHECK(upload(&ctx->d_sep_tab_row, trow));
        CC_CHECK(upload(&ctx->d_sep_tab_col, tcol));
        CC_CHECK(upload(&ctx->d_sep_tab_t1, tt1));
        CC_CHECK(upload(&ctx->d_sep_tab_t2, tt2));
        CC_CUDA(cudaMalloc(&ctx->d_sep_patterns, std::max<size_t>(1, pats.size()) * sizeof(SepPattern)));
        CC_CUDA(cudaMemcpy(ctx->d_sep_patterns, pats.data(), pats.size() * sizeof(SepPattern), cudaMemcpyHostToDevice));
        ctx->use_sep = true;
      }
    }
  }
  return CC_OK;
}

int cc_boost_precalc(cc_ctx* ctx, int num_samples) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_boost_precalc: evaluator not initialised");
  CC_REQUIRE(num_samples > 0 && num_samples <= ctx->max_samples, CC_ERR_BAD_ARG, "cc_boost_precalc: numSamples %d not in (0, %d]", num_samples, ctx->max_samples);
  CC_CUDA(cudaSetDevice(ctx->device));
  const int N = num_samples;
  const int n_stride = (N + 255) & ~255;   // rows padded to the split-search tile (256 elements)
  // ordered variables: rows of more than SORT_CAP samples (and cc_boost_set_search_mode(ctx, 2)) take the wide path
  const bool wide = ctx->ftype != CC_FEATURE_LBP && (N > SORT_CAP || ctx->force_wide);
  CC_CHECK(ensure_traindata(ctx, ctx->nrows, n_stride, wide));
  ctx->wide = wide;
  ctx->N = N; ctx->n_stride = n_stride;
  const int ntiles_used = (N + TILE - 1) / TILE;
  const int P = ctx->P;

  if (ctx->ftype == CC_FEATURE_HAAR || ctx->ftype == CC_FEATURE_LBP) {
    const size_t tile_bytes = (size_t)P * TILE * sizeof(int32_t);
    const bool in_global = tile_bytes + 64 > 227 * 1024;         // big windows: taps read from the tiles in HBM / L2
    const int tiles_per_cta = (in_global || 2 * tile_bytes + 64 <= 200 * 1024) ? 2 : 1;
    const size_t smem = in_global ? 0 : tiles_per_cta * tile_bytes + 64;
    const int gx = (ntiles_used + tiles_per_cta - 1) / tiles_per_cta;
    auto chunks_for = [&](int nrows_launch) {
      // enough CTAs for >= ~8 waves, but >= 4096 rows (128 per warp) per CTA so the tile load is amortised
      int want = std::max(1, (ctx->sm_count * 8 + gx - 1) / gx);
      int max_chunks = std::max(1, nrows_launch / 4096);
      int chunks = std::min(want, max_chunks);
      return std::max(1, chunks);
    };
    if (ctx->ftype == CC_FEATURE_HAAR) {
      struct Launch { const int32_t* plane; const HaarTaps* taps; int n; };
      Launch ls[2] = {{ctx->d_sum, ctx->d_taps_upright, (int)ctx->upright_rows.size()},
                      {ctx->d_tilted, ctx->d_taps_tilted, (int)ctx->tilted_rows.size()}};
      if (ctx->use_sep) {
        // separable kernels: upright patterns on the sum tiles, tilted patterns on the tilted tiles
        const size_t sep_smem = (size_t)P * TILE * 4 + (size_t)SEP_WARPS * (std::max(ctx->W, ctx->H) + 1) * TILE * 4 + 64;
        CC_CUDA(cudaFuncSetAttribute(k_precalc_haar_sep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sep_smem));
        struct SepLaunch { const int32_t* plane; const SepPattern* pats; int n; int which; };
        SepLaunch sl[2] = {{ctx->d_sum, ctx->d_sep_patterns, ctx->n_sep_patterns, 0},
                           {ctx->d_tilted, ctx->d_sep_patterns + ctx->n_sep_patterns, ctx->n_sep_tilted, 1}};
        for (const SepLaunch& L : sl) {
          if (L.n == 0 || ls[L.which].n == 0) continue;
          ScopedTimer tm(ctx, "precalc", 1);
          // the list is sorted heavy-first; every CTA drains its chunk through a shared counter, the grid has many CTAs per SM
          const int nchunks = std::max(1, std::min(L.n / (SEP_WARPS * 2), (ctx->sm_count * 6 + ntiles_used - 1) / ntiles_used));
          const int ppc = (L.n + nchunks - 1) / nchunks;
          dim3 grid(ntiles_used, (L.n + ppc - 1) / ppc);
          k_precalc_haar_sep<<<grid, SEP_THREADS, sep_smem, ctx->stream>>>(L.plane, ctx->d_nf, L.pats, L.n, ppc, ctx->d_sep_tab_row, ctx->d_sep_tab_col,
                                                                          ctx->d_sep_tab_t1, ctx->d_sep_tab_t2, ctx->W, ctx->H, ctx->row0,
                                                                          ctx->d_fval, n_stride, ntiles_used);
          CC_CUDA(cudaGetLastError());
          ls[L.which].n = 0;      // these rows are done
        }
      }
      for (const Launch& l : ls) {
        if (l.n == 0) continue;
        ScopedTimer tm(ctx, "precalc", 1);
        const int chunks = chunks_for(l.n);
        const int rpc = (l.n + chunks - 1) / chunks;
        dim3 grid(gx, (l.n + rpc - 1) / rpc);
        if (in_global) {
          k_precalc_haar<2, true><<<grid, PRECALC_THREADS, 0, ctx->stream>>>(l.plane, ctx->d_nf, l.taps, l.n, rpc, ctx->row0,
                                                                            ctx->d_fval, n_stride, P, ntiles_used);
        } else if (tiles_per_cta == 2) {
          CC_CUDA(cudaFuncSetAttribute(k_precalc_haar<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          k_precalc_haar<2, false><<<grid, PRECALC_THREADS, smem, ctx->stream>>>(l.plane, ctx->d_nf, l.taps, l.n, rpc, ctx->row0,
                                                                                 ctx->d_fval, n_stride, P, ntiles_used);
        } else {
          CC_CUDA(cudaFuncSetAttribute(k_precalc_haar<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          k_precalc_haar<1, false><<<grid, PRECALC_THREADS, smem, ctx->stream>>>(l.plane, ctx->d_nf, l.taps, l.n, rpc, ctx->row0,
                                                                                 ctx->d_fval, n_stride, P, ntiles_used);
        }
        CC_CUDA(cudaGetLastError());
      }
    } else {
      ScopedTimer tm(ctx, "precalc", 1);
      const int chunks = chunks_for(ctx->nrows);
      const int rpc = (ctx->nrows + chunks - 1) / chunks;
      dim3 grid(gx, (ctx->nrows + rpc - 1) / rpc);
      if (in_global) {
        k_precalc_lbp<2, true><<<grid, LBP_THREADS, 0, ctx->stream>>>(ctx->d_sum, ctx->d_lbp_taps, ctx->row0, ctx->nrows, rpc, ctx->d_fcode,
                                                                     n_stride, P, ntiles_used);
      } else if (tiles_per_cta == 2) {
        CC_CUDA(cudaFuncSetAttribute(k_precalc_lbp<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_precalc_lbp<2, false><<<grid, LBP_THREADS, smem, ctx->stream>>>(ctx->d_sum, ctx->d_lbp_taps, ctx->row0, ctx->nrows, rpc, ctx->d_fcode,
                                                                          n_stride, P, ntiles_used);
      } else {
        CC_CUDA(cudaFuncSetAttribute(k_precalc_lbp<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_precalc_lbp<1, false><<<grid, LBP_THREADS, smem, ctx->stream>>>(ctx->d_sum, ctx->d_lbp_taps, ctx->row0, ctx->nrows, rpc, ctx->d_fcode,
                                                                          n_stride, P, ntiles_used);
      }
      CC_CUDA(cudaGetLastError());
    }
  } else {
    ScopedTimer tm(ctx, "precalc", 1);
    const int blocks = ctx->sm_count * 8;
    k_precalc_hog<<<blocks, 256, 0, ctx->stream>>>(ctx->d_hog, ctx->d_hog_taps, ctx->row0, ctx->nrows, ctx->d_fval, n_stride, ctx->hogP, ctx->ns_pad,
                                                   ntiles_used);
    CC_CUDA(cudaGetLastError());
  }

  if (wide) {
    ScopedTimer tm(ctx, "sort", 1);
    const int blocks = std::min(ctx->nrows, ctx->sm_count * 2);
    uint32_t* scratch = nullptr;
    CC_CUDA(cudaMallocAsync(&scratch, (size_t)blocks * 4 * n_stride * sizeof(uint32_t), ctx->stream));
    k_sort_rows_wide<<<blocks, SORT_THREADS, 0, ctx->stream>>>(ctx->d_fval, n_stride, N, ctx->nrows, ctx->d_sidx32, scratch);
    CC_CUDA(cudaGetLastError());
    CC_CUDA(cudaFreeAsync(scratch, ctx->stream));
  } else if (ctx->ftype != CC_FEATURE_LBP) {   // categorical variables are not sorted (numPrecalcIdx = 0, :267-273)
    const int blocks = std::min(ctx->nrows, ctx->sm_count);
    // Haar rows: bucket + rank sort, then the radix kernel on the rows it deferred.  HOG rows (long runs of exact zeros)
    // and cc_boost_set_sort_mode(ctx, 1): radix kernel on every row.
    const bool bucket = ctx->sort_mode == 2 || (ctx->sort_mode == 0 && ctx->ftype == CC_FEATURE_HAAR);
    ScopedTimer tm(ctx, "sort", bucket ? 2 : 1);
    unsigned int* hard_count = reinterpret_cast<unsigned int*>(ctx->d_hard_rows);
    int* hard_rows = ctx->d_hard_rows + 1;
    const float hard_limit = 96.0f * (float)N;        // sum k^2 > 96 N: the ranking would cost more than the radix passes
    if (bucket) CC_CUDA(cudaMemsetAsync(hard_count, 0, sizeof(unsigned int), ctx->stream));
#define LAUNCH_SORT(KPT)                                                                                                     \
  {                                                                                                                          \
    if (bucket) {                                                                                                            \
      CC_CUDA(cudaFuncSetAttribute(k_sort_rows_bucket<KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsort_smem<KPT>())); \
      k_sort_rows_bucket<KPT><<<blocks, SORT_THREADS, bsort_smem<KPT>(), ctx->stream>>>(ctx->d_fval, n_stride, N, ctx->nrows, ctx->d_sidx, ctx->d_sval, \
                                                                                      hard_rows, hard_count, hard_limit);   \
    }                                                                                                                        \
    CC_CUDA(cudaFuncSetAttribute(k_sort_rows<KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem(KPT)));       \
    k_sort_rows<KPT><<<blocks, SORT_THREADS, sort_smem(KPT), ctx->stream>>>(ctx->d_fval, n_stride, N, ctx->nrows, ctx->d_sidx, ctx->d_sval, \
                                                                            bucket ? hard_rows : nullptr, bucket ? hard_count : nullptr);   \
  }
    if (N <= SORT_THREADS * 3) LAUNCH_SORT(3)
    else if (N <= SORT_THREADS * 7) LAUNCH_SORT(7)
    else if (N <= SORT_THREADS * 11) LAUNCH_SORT(11)
    else LAUNCH_SORT(15)
#undef LAUNCH_SORT
    CC_CUDA(cudaGetLastError());
  }
  ctx->precalc_valid = true;
  return CC_OK;
}

int cc_boost_set_sort_mode(cc_ctx* ctx, int mode) {
  CC_REQUIRE(ctx && mode >= 0 && mode <= 2, CC_ERR_BAD_ARG, "cc_boost_set_sort_mode: mode must be 0 (auto), 1 (radix passes) or 2 (bucket + rank)");
  ctx->sort_mode = mode;
  return CC_OK;
}

int cc_boost_get_sort_stats(cc_ctx* ctx, int64_t* deferred_rows) {
  CC_REQUIRE(ctx && deferred_rows, CC_ERR_BAD_ARG, "cc_boost_get_sort_stats: null argument");
  CC_REQUIRE(ctx->precalc_valid && ctx->d_hard_rows, CC_ERR_STATE, "cc_boost_get_sort_stats: no sorted rows (precalc not run, or categorical features)");
  CC_CUDA(cudaSetDevice(ctx->device));
  const bool bucket = ctx->sort_mode == 2 || (ctx->sort_mode == 0 && ctx->ftype == CC_FEATURE_HAAR);
  unsigned int h = 0;
  if (bucket) {
    CC_CUDA(cudaMemcpyAsync(&h, ctx->d_hard_rows, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  *deferred_rows = (int64_t)h;
  return CC_OK;
}

int cc_debug_div_check(cc_ctx* ctx, int steps, uint64_t seed, int64_t n, uint64_t* mismatches) {
  CC_REQUIRE(ctx && mismatches && (steps == 1 || steps == 2) && n > 0, CC_ERR_BAD_ARG, "cc_debug_div_check: bad arguments");
  CC_CUDA(cudaSetDevice(ctx->device));
  unsigned long long* d = nullptr;
  CC_CUDA(cudaMallocAsync(&d, sizeof(unsigned long long), ctx->stream));
  CC_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream));
  k_div_check<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(steps, seed, n, d);
  CC_CUDA(cudaGetLastError());
  unsigned long long h = 0;
  CC_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  CC_CUDA(cudaFreeAsync(d, ctx->stream));
  *mismatches = h;
  return CC_OK;
}

int cc_boost_get_fval(cc_ctx* ctx, int row0, int nrows, float* out) {
  CC_REQUIRE(ctx && ctx->precalc_valid, CC_ERR_STATE, "cc_boost_get_fval: precalc has not run");
  CC_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= ctx->nrows, CC_ERR_BAD_ARG, "cc_boost_get_fval: rows [%d,%d) outside local slice of %d rows", row0, row0 + nrows, ctx->nrows);
  CC_CUDA(cudaSetDevice(ctx->device));
  if (ctx->ftype == CC_FEATURE_LBP) {
    float* tmp = nullptr;
    CC_CUDA(cudaMallocAsync(&tmp, (size_t)nrows * ctx->N * sizeof(float), ctx->stream));
    k_codes_to_float<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->d_fcode + (size_t)row0 * ctx->n_stride, ctx->n_stride, ctx->N, nrows, tmp);
    CC_CUDA(cudaGetLastError());
    CC_CUDA(cudaMemcpyAsync(out, tmp, (size_t)nrows * ctx->N * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
    CC_CUDA(cudaFreeAsync(tmp, ctx->stream));
    return CC_OK;
  }
  CC_CUDA(cudaMemcpy2DAsync(out, (size_t)ctx->N * sizeof(float), ctx->d_fval + (size_t)row0 * ctx->n_stride, (size_t)ctx->n_stride * sizeof(float),
                            (size_t)ctx->N * sizeof(float), nrows, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  return CC_OK;
}

int cc_boost_get_sorted_idx(cc_ctx* ctx, int row0, int nrows, uint16_t* out) {
  if (ctx && ctx->precalc_valid && ctx->wide) {      // wide path: narrow the 32-bit rows (numSamples < 65536 only)
    CC_REQUIRE(ctx->N < 65536, CC_ERR_BAD_ARG, "cc_boost_get_sorted_idx: numSamples %d needs cc_boost_get_sorted_idx32", ctx->N);
    CC_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= ctx->nrows, CC_ERR_BAD_ARG, "cc_boost_get_sorted_idx: rows outside the local slice");
    std::vector<uint32_t> tmp((size_t)nrows * ctx->N);
    CC_CHECK(cc_boost_get_sorted_idx32(ctx, row0, nrows, tmp.data()));
    for (size_t i = 0; i < tmp.size(); ++i) out[i] = (uint16_t)tmp[i];
    return CC_OK;
  }
  CC_REQUIRE(ctx && ctx->precalc_valid && ctx->d_sidx, CC_ERR_STATE, "cc_boost_get_sorted_idx: no sorted-index buffer (precalc not run, or categorical features)");
  CC_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= ctx->nrows, CC_ERR_BAD_ARG, "cc_boost_get_sorted_idx: rows out of range");
  CC_CUDA(cudaSetDevice(ctx->device));
  CC_CUDA(cudaMemcpy2DAsync(out, (size_t)ctx->N * sizeof(uint16_t), ctx->d_sidx + (size_t)row0 * ctx->n_stride, (size_t)ctx->n_stride * sizeof(uint16_t),
                            (size_t)ctx->N * sizeof(uint16_t), nrows, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  return CC_OK;
}

int cc_boost_get_sorted_idx32(cc_ctx* ctx, int row0, int nrows, uint32_t* out) {
  CC_REQUIRE(ctx && ctx->precalc_valid && (ctx->d_sidx32 || ctx->d_sidx), CC_ERR_STATE, "cc_boost_get_sorted_idx32: no sorted-index buffer (precalc not run, or categorical features)");
  CC_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= ctx->nrows && out, CC_ERR_BAD_ARG, "cc_boost_get_sorted_idx32: rows outside the local slice");
  CC_CUDA(cudaSetDevice(ctx->device));
  if (ctx->wide) {
    CC_CUDA(cudaMemcpy2DAsync(out, (size_t)ctx->N * sizeof(uint32_t), ctx->d_sidx32 + (size_t)row0 * ctx->n_stride, (size_t)ctx->n_stride * sizeof(uint32_t),
                              (size_t)ctx->N * sizeof(uint32_t), nrows, cudaMemcpyDeviceToHost, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
    return CC_OK;
  }
  std::vector<uint16_t> tmp((size_t)nrows * ctx->N);
  CC_CHECK(cc_boost_get_sorted_idx(ctx, row0, nrows, tmp.data()));
  for (size_t i = 0; i < tmp.size(); ++i) out[i] = tmp[i];
  return CC_OK;
}

int cc_boost_get_sorted_val(cc_ctx* ctx, int row0, int nrows, float* out) {
  if (ctx && ctx->precalc_valid && ctx->wide) {        // wide path keeps no sorted-value rows: gather them
    std::vector<uint32_t> idx((size_t)nrows * ctx->N);
    std::vector<float> fv((size_t)nrows * ctx->N);
    CC_CHECK(cc_boost_get_sorted_idx32(ctx, row0, nrows, idx.data()));
    CC_CHECK(cc_boost_get_fval(ctx, row0, nrows, fv.data()));
    for (int r = 0; r < nrows; ++r)
      for (int i = 0; i < ctx->N; ++i) out[(size_t)r * ctx->N + i] = fv[(size_t)r * ctx->N + idx[(size_t)r * ctx->N + i]];
    return CC_OK;
  }
  CC_REQUIRE(ctx && ctx->precalc_valid && ctx->d_sval, CC_ERR_STATE, "cc_boost_get_sorted_val: no sorted-value buffer");
  CC_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= ctx->nrows, CC_ERR_BAD_ARG, "cc_boost_get_sorted_val: rows out of range");
  CC_CUDA(cudaSetDevice(ctx->device));
  CC_CUDA(cudaMemcpy2DAsync(out, (size_t)ctx->N * sizeof(float), ctx->d_sval + (size_t)row0 * ctx->n_stride, (size_t)ctx->n_stride * sizeof(float),
                            (size_t)ctx->N * sizeof(float), nrows, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  return CC_OK;
}

}  // extern "C"

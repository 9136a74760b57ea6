// Training-sample generation on the GPU: `opencv_createsamples -img <image> -num <n> -vec <out>` without -bg
// (tools/createsamples/utility.cpp:952-1029 cvCreateTrainingSamples; SURVEY.md 8f rank 3 -- the input-preparation step in front
// of the hot path).  Per sample the reference draws a random rotation, warps the source image and its mask with ITS OWN
// scan-line perspective warp (utility.cpp:226-417, not cv::warpPerspective), blurs the mask, crops, resizes to the window with
// INTER_LINEAR_EXACT and alpha-blends onto a bgcolor-filled window (utility.cpp:580-672).
//
// Split of work: everything that is sequential in the reference -- the cv::RNG stream, Rodrigues, the 8x8 linear system of the
// perspective transform and the edge walk that decides which destination pixels a warp writes (incrementally accumulated
// doubles) -- is planned on the host in f64 (a few hundred flops per sample); the per-pixel work (two bilinear warps of
// (2w x 2h) pixels, 3x3 Gaussian, two fixed-point resizes, the blend) runs on the GPU, one CTA row per sample row.
// All f64 pixel arithmetic uses un-fused _rn operations so that it rounds like the reference's scalar C++.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "cc_internal.h"

namespace cc {

// ---- cv::RNG (opencv2/core/operations.hpp): multiply-with-carry; operator double(), uniform(double,double), uniform(int,int) ----
struct CvRng {
  uint64_t state;
  explicit CvRng(uint64_t seed) : state(seed ? seed : 0xffffffffull) {}
  unsigned next() { state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
  double real() { unsigned t = next(); return (double)(((uint64_t)t << 32) | next()) * 5.4210108624275221700372640043497e-20; }
  double uniform(double a, double b) { return real() * (b - a) + a; }
  int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// cv::Rodrigues, rotation vector -> matrix (calib3d; R = cos(t) I + (1 - cos(t)) r r^T + sin(t) [r]x on the unit axis)
static void rodrigues(const double rv[3], double R[9]) {
  const double theta = std::sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
  if (theta < 2.220446049250313e-16) { const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; std::memcpy(R, I, sizeof(I)); return; }
  const double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c, it = theta ? 1. / theta : 0.;
  const double x = rv[0] * it, y = rv[1] * it, z = rv[2] * it;
  const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
  const double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) R[i] = c * I[i] + c1 * rrt[i] + s * rx[i];
}

// cv::solve(A, b, x, DECOMP_LU) for the 8x8 system: Gaussian elimination with partial pivoting (hal::LU), then back substitution
static bool solve8(double A[8][8], double b[8]) {
  const int m = 8;
  for (int i = 0; i < m; ++i) {
    int k = i;
    for (int j = i + 1; j < m; ++j) if (std::fabs(A[j][i]) > std::fabs(A[k][i])) k = j;
    if (std::fabs(A[k][i]) < 2.220446049250313e-16 * 100) return false;
    if (k != i) { for (int j = i; j < m; ++j) std::swap(A[i][j], A[k][j]); std::swap(b[i], b[k]); }
    const double d = -1 / A[i][i];
    for (int j = i + 1; j < m; ++j) {
      const double alpha = A[j][i] * d;
      for (int kk = i + 1; kk < m; ++kk) A[j][kk] += alpha * A[i][kk];
      b[j] += alpha * b[i];
    }
  }
  for (int i = m - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < m; ++k) s -= A[i][k] * b[k];
    b[i] = s / A[i][i];
  }
  return true;
}

// icvRandomQuad, utility.cpp:419-466
static void random_quad(CvRng& rng, int width, int height, double maxx, double maxy, double maxz, double quad[4][2]) {
  const double distfactor = 3.0, distfactor2 = 1.0;
  double rv[3], R[9];
  rv[0] = rng.uniform(-maxx, maxx);
  rv[1] = (maxy - std::fabs(rv[0])) * rng.uniform(-1.0, 1.0);
  rv[2] = rng.uniform(-maxz, maxz);
  const double d = (distfactor + distfactor2 * rng.uniform(-1.0, 1.0)) * width;
  rodrigues(rv, R);
  const double halfw = 0.5 * width, halfh = 0.5 * height;
  const double q0[4][2] = {{-halfw, -halfh}, {halfw, -halfh}, {halfw, halfh}, {-halfw, halfh}};
  for (int i = 0; i < 4; ++i) {
    const double in[3] = {q0[i][0], q0[i][1], 0.0};
    double v[3];
    for (int r = 0; r < 3; ++r) { double s = 0; for (int k = 0; k < 3; ++k) s += R[r * 3 + k] * in[k]; v[r] = s; }      // gemm(rotMat, rotVect)
    quad[i][0] = v[0] * d / (d + v[2]) + halfw;
    quad[i][1] = v[1] * d / (d + v[2]) + halfh;
  }
}

// cvGetPerspectiveTransform, utility.cpp:180-223
static bool perspective_transform(int sw, int sh, const double quad[4][2], double c[9]) {
  double a[8][8], b[8];
  std::memset(a, 0, sizeof(a)); std::memset(b, 0, sizeof(b));
  for (int i = 0; i < 4; ++i) { a[i][0] = quad[i][0]; a[i][1] = quad[i][1]; a[i][2] = 1; }
  for (int i = 4; i < 8; ++i) { a[i][3] = quad[i - 4][0]; a[i][4] = quad[i - 4][1]; a[i][5] = 1; }
  const int u = sw - 1, v = sh - 1;
  a[1][6] = -quad[1][0] * u; a[1][7] = -quad[1][1] * u;
  a[2][6] = -quad[2][0] * u; a[2][7] = -quad[2][1] * u;
  b[1] = b[2] = u;
  a[6][6] = -quad[2][0] * v; a[6][7] = -quad[2][1] * v;
  a[7][6] = -quad[3][0] * v; a[7][7] = -quad[3][1] * v;
  b[6] = b[7] = v;
  if (!solve8(a, b)) return false;
  for (int i = 0; i < 8; ++i) c[i] = b[i];
  c[8] = 1;
  return true;
}

static inline int cv_round(double v) { return (int)std::nearbyint(v); }      // cvRound: round half to even (default FP environment)

struct RowSpan { int y, x0, x1; };

// the destination rows cvWarpPerspective writes and their [ix_min, ix_max] (utility.cpp:226-417, the part around the pixel loop)
static bool warp_rows(const double quad[4][2], int dst_w, int dst_h, std::vector<RowSpan>& rows) {
  rows.clear();
  int direction = 0;
  for (int i = 0; i < 4; ++i) {
    const int ni = (i + 1) & 3, pi = (i + 3) & 3;
    const double d = (quad[i][0] - quad[pi][0]) * (quad[ni][1] - quad[i][1]) - (quad[i][1] - quad[pi][1]) * (quad[ni][0] - quad[i][0]);
    const int cur = d > 0 ? 1 : d < 0 ? -1 : 0;
    if (direction == 0) direction = cur;
    else if (direction * cur < 0) { direction = 0; break; }
  }
  if (direction == 0) return false;            // "Quadrangle is nonconvex or degenerated."
  int left = 0;
  for (int i = 1; i < 4; ++i)
    if (quad[i][1] < quad[left][1] || (quad[i][1] == quad[left][1] && quad[i][0] < quad[left][0])) left = i;
  double q[4][2];
  if (direction > 0) {
    for (int i = left; i < 4; ++i) { q[i - left][0] = quad[i][0]; q[i - left][1] = quad[i][1]; }
    for (int i = 0; i < left; ++i) { q[4 - left + i][0] = quad[i][0]; q[4 - left + i][1] = quad[i][1]; }
  } else {
    for (int i = left; i >= 0; --i) { q[left - i][0] = quad[i][0]; q[left - i][1] = quad[i][1]; }
    for (int i = 3; i > left; --i) { q[4 + left - i][0] = quad[i][0]; q[4 + left - i][1] = quad[i][1]; }
  }
  left = 0;
  int right = 0;
  if (q[left][1] == q[left + 1][1]) right = 1;
  int next_left = 3, next_right = right + 1;
  double y_min = q[left][1] - 1;
  auto edge = [&](int p, int n, double& k, double& bb) {
    k = (q[p][0] - q[n][0]) / (q[p][1] - q[n][1]);
    bb = (q[p][1] * q[n][0] - q[p][0] * q[n][1]) / (q[p][1] - q[n][1]);
  };
  double k_left, b_left, k_right, b_right;
  edge(left, next_left, k_left, b_left);
  edge(right, next_right, k_right, b_right);
  for (;;) {
    const double y_max = std::min(q[next_left][1], q[next_right][1]);
    const int iy_min = std::max(cv_round(y_min), 0) + 1;
    const int iy_max = std::min(cv_round(y_max), dst_h - 1);
    double x_min = k_left * iy_min + b_left;
    double x_max = k_right * iy_min + b_right;
    for (int y = iy_min; y <= iy_max; ++y) {
      rows.push_back(RowSpan{y, std::max(cv_round(x_min), 0), std::min(cv_round(x_max), dst_w - 1)});
      x_min += k_left;
      x_max += k_right;
    }
    if (next_left == next_right || (next_left + 1 == next_right && q[next_left][1] == q[next_right][1])) break;
    if (y_max == q[next_left][1]) { left = next_left; next_left = left - 1; edge(left, next_left, k_left, b_left); }
    if (y_max == q[next_right][1]) { right = next_right; next_right = right + 1; edge(right, next_right, k_right, b_right); }
    y_min = y_max;
  }
  return true;
}

// per-sample plan handed to the device
struct SamplePlan {
  double c[9];                 // perspective coefficients
  int row_first, row_count;    // its RowSpans in the span table
  int roi_x, roi_y, roi_w, roi_h;   // crop of img / maskimg that is resized to the window (already intersected with the image)
  int forecolordev, inverse;
};

// icvStartSampleDistortion, utility.cpp:516-578: mask + border extension (3x3 erode / dilate = min / max over the in-image neighbours)
static void start_distortion(const uint8_t* in, int w, int h, int bgcolor, int bgthreshold, std::vector<uint8_t>& src, std::vector<uint8_t>& mask) {
  src.assign(in, in + (size_t)w * h);
  mask.resize((size_t)w * h);
  std::vector<uint8_t> er((size_t)w * h), di((size_t)w * h);
  for (int r = 0; r < h; ++r)
    for (int c = 0; c < w; ++c) {
      const int p = in[(size_t)r * w + c];
      mask[(size_t)r * w + c] = (bgcolor - bgthreshold <= p && p <= bgcolor + bgthreshold) ? 0 : 255;
      int mn = 255, mx = 0;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int y = r + dy, x = c + dx;
          if (y < 0 || y >= h || x < 0 || x >= w) continue;
          const int v = in[(size_t)y * w + x];
          mn = std::min(mn, v); mx = std::max(mx, v);
        }
      er[(size_t)r * w + c] = (uint8_t)mn; di[(size_t)r * w + c] = (uint8_t)mx;
    }
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    if (mask[i] != 0) continue;
    const uint8_t de = (uint8_t)(bgcolor - er[i]), dd = (uint8_t)(di[i] - bgcolor);
    if (de >= dd && de > bgthreshold) src[i] = er[i];
    if (dd > de && dd > bgthreshold) src[i] = di[i];
  }
}

// the host half of cvCreateTrainingSamples' loop for samples [0, count): RNG stream, quads, coefficients, row spans, crops
static int plan_samples(int w, int h, int count, int invert, int maxintensitydev, double maxx, double maxy, double maxz, int win_w, int win_h,
                        int rngseed, std::vector<SamplePlan>& plans, std::vector<RowSpan>& spans, std::vector<double>* quads_out) {
  CvRng rng((uint64_t)(int64_t)rngseed);        // cv::setRNGSeed(int) -> theRNG() = RNG((uint64)seed)
  const int dx = w / 2, dy = h / 2, iw = w + 2 * dx, ih = h + 2 * dy;
  plans.resize(count);
  spans.clear();
  std::vector<RowSpan> rows;
  int inverse = invert;
  for (int i = 0; i < count; ++i) {
    SamplePlan& p = plans[i];
    if (invert == 0x7FFFFFFF) inverse = rng.uniform(0, 2);         // CV_RANDOM_INVERT
    double quad[4][2];
    random_quad(rng, w, h, maxx, maxy, maxz, quad);
    for (int k = 0; k < 4; ++k) { quad[k][0] += (double)dx; quad[k][1] += (double)dy; }
    if (quads_out) for (int k = 0; k < 4; ++k) { quads_out->push_back(quad[k][0]); quads_out->push_back(quad[k][1]); }
    CC_REQUIRE(perspective_transform(w, h, quad, p.c), CC_ERR_BAD_ARG, "cc_createsamples: singular perspective system for sample %d", i);
    CC_REQUIRE(warp_rows(quad, iw, ih, rows), CC_ERR_BAD_ARG, "Quadrangle is nonconvex or degenerated.");
    p.row_first = (int)spans.size(); p.row_count = (int)rows.size();
    spans.insert(spans.end(), rows.begin(), rows.end());
    // icvPlaceDistortedSample with inscribe = 0, maxshiftf = 0, maxscalef = 0 (utility.cpp:612-650); the RNG is drawn all the same
    int crx = dx, cry = dy, crw = w, crh = h;
    const double maxshiftf = 0.0, maxscalef = 0.0;
    const double xshift = rng.uniform(0., maxshiftf), yshift = rng.uniform(0., maxshiftf);
    crx -= (int)(xshift * crw); cry -= (int)(yshift * crh);
    crw = (int)((1.0 + maxshiftf) * crw); crh = (int)((1.0 + maxshiftf) * crh);
    const double randscale = rng.uniform(0., maxscalef);
    crx -= (int)(0.5 * randscale * crw); cry -= (int)(0.5 * randscale * crh);
    crw = (int)((1.0 + randscale) * crw); crh = (int)((1.0 + randscale) * crh);
    const float scale = std::max(((float)crw) / win_w, ((float)crh) / win_h);
    const int rx = (int)(-0.5F * (scale * win_w - crw) + crx), ry = (int)(-0.5F * (scale * win_h - crh) + cry);
    const int rw = (int)(scale * win_w), rh = (int)(scale * win_h);
    const int x0 = std::max(rx, 0), y0 = std::max(ry, 0), x1 = std::min(rx + rw, iw), y1 = std::min(ry + rh, ih);
    CC_REQUIRE(x1 > x0 && y1 > y0, CC_ERR_BAD_ARG, "cc_createsamples: empty crop for sample %d", i);
    p.roi_x = x0; p.roi_y = y0; p.roi_w = x1 - x0; p.roi_h = y1 - y0;
    p.forecolordev = rng.uniform(-maxintensitydev, maxintensitydev);
    p.inverse = inverse;
  }
  return CC_OK;
}

// ---- device ----------------------------------------------------------------------------------------------------------------
// cvWarpPerspective's pixel loop (utility.cpp:353-393) for one destination row span of one sample, image and mask at once
__global__ void k_cs_warp(const SamplePlan* __restrict__ plans, const RowSpan* __restrict__ spans, int s0, int ns, int total_spans_first,
                          const uint8_t* __restrict__ src, const uint8_t* __restrict__ mask, int sw, int sh, int iw, int ih,
                          uint8_t* __restrict__ img, uint8_t* __restrict__ maskimg) {
  const int s = s0 + blockIdx.y;
  if (blockIdx.y >= ns) return;
  const SamplePlan& p = plans[s];
  uint8_t* di = img + (size_t)blockIdx.y * iw * ih;
  uint8_t* dm = maskimg + (size_t)blockIdx.y * iw * ih;
  for (int r = blockIdx.x; r < p.row_count; r += gridDim.x) {
    const RowSpan sp = spans[p.row_first + r];
    const double y = (double)sp.y;
    for (int xi = sp.x0 + threadIdx.x; xi <= sp.x1; xi += blockDim.x) {
      const double x = (double)xi;
      const double div = __dadd_rn(__dadd_rn(__dmul_rn(p.c[6], x), __dmul_rn(p.c[7], y)), p.c[8]);
      const double src_x = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(p.c[0], x), __dmul_rn(p.c[1], y)), p.c[2]), div);
      const double src_y = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(p.c[3], x), __dmul_rn(p.c[4], y)), p.c[5]), div);
      const int isx = (int)floor(src_x), isy = (int)floor(src_y);
      const double fx = __dsub_rn(src_x, (double)isx), fy = __dsub_rn(src_y, (double)isy);
      const bool x0ok = isx >= 0 && isx < sw, x1ok = isx >= -1 && isx + 1 < sw, y0ok = isy >= 0 && isy < sh, y1ok = isy >= -1 && isy + 1 < sh;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint8_t* S = which ? mask : src;
        const int i00 = (x0ok && y0ok) ? S[(size_t)isy * sw + isx] : 0;
        const int i10 = (x1ok && y0ok) ? S[(size_t)isy * sw + isx + 1] : 0;
        const int i01 = (x0ok && y1ok) ? S[(size_t)(isy + 1) * sw + isx] : 0;
        const int i11 = (x1ok && y1ok) ? S[(size_t)(isy + 1) * sw + isx + 1] : 0;
        const double i0 = __dadd_rn((double)i00, __dmul_rn((double)(i10 - i00), fx));
        const double i1 = __dadd_rn((double)i01, __dmul_rn((double)(i11 - i01), fx));
        const double v = __dadd_rn(i0, __dmul_rn(__dsub_rn(i1, i0), fy));
        (which ? dm : di)[(size_t)sp.y * iw + xi] = (uint8_t)(int)v;       // (uchar)double: truncation, value in [0, 255]
      }
    }
  }
  (void)total_spans_first;
}

// GaussianBlur(maskimg, maskimg, Size(3,3), 0, 0) on 8-bit data: ([1 2 1] x [1 2 1] * p + 8) >> 4, BORDER_REFLECT_101
// (pinned against cv2.GaussianBlur in tests/test_oracle_createsamples.py)
__global__ void k_cs_blur(const uint8_t* __restrict__ in, int iw, int ih, int ns, uint8_t* __restrict__ out) {
  const size_t total = (size_t)ns * iw * ih;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(t / ((size_t)iw * ih));
    const int r = (int)(t - (size_t)s * iw * ih);
    const int y = r / iw, x = r - y * iw;
    const uint8_t* p = in + (size_t)s * iw * ih;
    auto rx = [&](int v) { return v < 0 ? -v : (v >= iw ? 2 * iw - 2 - v : v); };
    auto ry = [&](int v) { return v < 0 ? -v : (v >= ih ? 2 * ih - 2 - v : v); };
    int sum = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) sum += (2 - (dy != 0)) * (2 - (dx != 0)) * (int)p[(size_t)ry(y + dy) * iw + rx(x + dx)];
    out[t] = (uint8_t)((sum + 8) >> 4);
  }
}

// cv::resize(..., INTER_LINEAR_EXACT) coefficient of one axis (same restatement as cc_mine.cu)
__device__ __forceinline__ void cs_lin_coef(double scale, int v, int src, int& ofs, int& c0, int& c1, int& mode) {
  double f = __dsub_rn(__dmul_rn(scale, (double)v + 0.5), 0.5);
  const int i = (int)floor(f);
  f = __dsub_rn(f, (double)i);
  if (i >= 0 && src > 1) {
    if (i < src - 1) { ofs = i; c1 = __double2int_rn(__dmul_rn(f, 256.0)); c0 = 256 - c1; mode = 0; }
    else { ofs = src - 1; c0 = 256; c1 = 0; mode = 2; }
  } else { ofs = 0; c0 = 256; c1 = 0; mode = 1; }
}
__device__ __forceinline__ int cs_resize_px(const uint8_t* __restrict__ base, int pitch, int sw, int sh, double sx, double sy, int x, int y) {
  int ox, cx0, cx1, mx, oy, cy0, cy1, my;
  cs_lin_coef(sx, x, sw, ox, cx0, cx1, mx);
  cs_lin_coef(sy, y, sh, oy, cy0, cy1, my);
  auto hrow = [&](int row) -> uint32_t {
    const uint8_t* s = base + (size_t)row * pitch;
    if (mx == 1) return (uint32_t)s[0] * 256u;
    if (mx == 2) return (uint32_t)s[sw - 1] * 256u;
    return (uint32_t)s[ox] * cx0 + (uint32_t)s[ox + 1] * cx1;
  };
  uint32_t val;
  if (my == 1) val = hrow(0) * 256u;
  else if (my == 2) val = hrow(sh - 1) * 256u;
  else val = hrow(oy) * cy0 + hrow(oy + 1) * cy1;
  return (int)((val + 32768u) >> 16);
}

// crop + resize of image and blurred mask + the blend of utility.cpp:660-671 onto a bgcolor window
__global__ void k_cs_compose(const SamplePlan* __restrict__ plans, int s0, int ns, const uint8_t* __restrict__ img, const uint8_t* __restrict__ blurred,
                             int iw, int ih, int win_w, int win_h, int bgcolor, uint8_t* __restrict__ out) {
  const int npx = win_w * win_h;
  const size_t total = (size_t)ns * npx;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(t / npx), r = (int)(t - (size_t)k * npx);
    const int y = r / win_w, x = r - y * win_w;
    const SamplePlan& p = plans[s0 + k];
    const double sx = 1.0 / ((double)win_w / (double)p.roi_w), sy = 1.0 / ((double)win_h / (double)p.roi_h);
    const size_t off = (size_t)k * iw * ih + (size_t)p.roi_y * iw + p.roi_x;
    const int pi = cs_resize_px(img + off, iw, p.roi_w, p.roi_h, sx, sy, x, y);
    const int alpha = cs_resize_px(blurred + off, iw, p.roi_w, p.roi_h, sx, sy, x, y);
    int chartmp = max(0, min(255, p.forecolordev + pi));
    if (p.inverse) chartmp ^= 0xFF;
    out[(size_t)(s0 + k) * npx + r] = (uint8_t)((chartmp * alpha + (255 - alpha) * bgcolor) / 255);
  }
}

}  // namespace cc

using namespace cc;

extern "C" {

int cc_createsamples_plan(int w, int h, int count, int invert, int maxintensitydev, double maxxangle, double maxyangle, double maxzangle, int win_w,
                          int win_h, int rngseed, double* quads, double* coeffs, int32_t* misc) {
  CC_REQUIRE(w > 1 && h > 1 && count >= 0 && win_w > 0 && win_h > 0, CC_ERR_BAD_ARG, "cc_createsamples_plan: bad arguments");
  std::vector<SamplePlan> plans;
  std::vector<RowSpan> spans;
  std::vector<double> q;
  CC_CHECK(plan_samples(w, h, count, invert, maxintensitydev, maxxangle, maxyangle, maxzangle, win_w, win_h, rngseed, plans, spans, &q));
  if (quads) std::memcpy(quads, q.data(), q.size() * sizeof(double));
  for (int i = 0; i < count; ++i) {
    if (coeffs) std::memcpy(coeffs + (size_t)i * 9, plans[i].c, sizeof(double) * 9);
    if (misc) {
      int32_t* m = misc + (size_t)i * 8;
      m[0] = plans[i].roi_x; m[1] = plans[i].roi_y; m[2] = plans[i].roi_w; m[3] = plans[i].roi_h;
      m[4] = plans[i].forecolordev; m[5] = plans[i].inverse; m[6] = plans[i].row_count;
      long long px = 0;
      for (int r = 0; r < plans[i].row_count; ++r) { const RowSpan& sp = spans[plans[i].row_first + r]; px += std::max(0, sp.x1 - sp.x0 + 1); }
      m[7] = (int32_t)px;
    }
  }
  return CC_OK;
}

int cc_createsamples(cc_ctx* ctx, const uint8_t* host_gray, int w, int h, int bgcolor, int bgthreshold, int count, int invert, int maxintensitydev,
                     double maxxangle, double maxyangle, double maxzangle, int win_w, int win_h, int rngseed, uint8_t* out_samples) {
  CC_REQUIRE(ctx && host_gray && out_samples && w > 1 && h > 1 && count >= 0 && win_w > 0 && win_h > 0, CC_ERR_BAD_ARG, "cc_createsamples: bad arguments");
  if (count == 0) return CC_OK;
  CC_CUDA(cudaSetDevice(ctx->device));
  std::vector<uint8_t> src, mask;
  start_distortion(host_gray, w, h, bgcolor, bgthreshold, src, mask);
  std::vector<SamplePlan> plans;
  std::vector<RowSpan> spans;
  CC_CHECK(plan_samples(w, h, count, invert, maxintensitydev, maxxangle, maxyangle, maxzangle, win_w, win_h, rngseed, plans, spans, nullptr));
  const int dx = w / 2, dy = h / 2, iw = w + 2 * dx, ih = h + 2 * dy;
  const size_t plane = (size_t)iw * ih;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)count, ((size_t)1 << 30) / (3 * plane)));      // <= 1 GiB of scratch
  uint8_t *d_src = nullptr, *d_mask = nullptr, *d_img = nullptr, *d_mimg = nullptr, *d_blur = nullptr, *d_out = nullptr;
  SamplePlan* d_plans = nullptr;
  RowSpan* d_spans = nullptr;
  auto cleanup = [&]() { cudaFree(d_src); cudaFree(d_mask); cudaFree(d_img); cudaFree(d_mimg); cudaFree(d_blur); cudaFree(d_out); cudaFree(d_plans); cudaFree(d_spans); };
#define CS_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); cc::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); return CC_ERR_CUDA; } } while (0)
  CS_CUDA(cudaMalloc(&d_src, (size_t)w * h)); CS_CUDA(cudaMalloc(&d_mask, (size_t)w * h));
  CS_CUDA(cudaMalloc(&d_img, plane * chunk)); CS_CUDA(cudaMalloc(&d_mimg, plane * chunk)); CS_CUDA(cudaMalloc(&d_blur, plane * chunk));
  CS_CUDA(cudaMalloc(&d_out, (size_t)count * win_w * win_h));
  CS_CUDA(cudaMalloc(&d_plans, sizeof(SamplePlan) * count)); CS_CUDA(cudaMalloc(&d_spans, sizeof(RowSpan) * std::max<size_t>(1, spans.size())));
  CS_CUDA(cudaMemcpyAsync(d_src, src.data(), src.size(), cudaMemcpyHostToDevice, ctx->stream));
  CS_CUDA(cudaMemcpyAsync(d_mask, mask.data(), mask.size(), cudaMemcpyHostToDevice, ctx->stream));
  CS_CUDA(cudaMemcpyAsync(d_plans, plans.data(), sizeof(SamplePlan) * count, cudaMemcpyHostToDevice, ctx->stream));
  CS_CUDA(cudaMemcpyAsync(d_spans, spans.data(), sizeof(RowSpan) * spans.size(), cudaMemcpyHostToDevice, ctx->stream));
  {
    ScopedTimer tm(ctx, "createsamples", 0);
    int launches = 0;
    for (int s0 = 0; s0 < count; s0 += chunk) {
      const int ns = std::min(chunk, count - s0);
      CS_CUDA(cudaMemsetAsync(d_img, bgcolor, plane * ns, ctx->stream));       // data->img = bgcolor; data->maskimg = 0
      CS_CUDA(cudaMemsetAsync(d_mimg, 0, plane * ns, ctx->stream));
      for (int y0 = 0; y0 < ns; y0 += 32768) {                                 // gridDim.y <= 65535
        const int nsy = std::min(32768, ns - y0);
        k_cs_warp<<<dim3(std::min(ih, 64), nsy), 128, 0, ctx->stream>>>(d_plans, d_spans, s0 + y0, nsy, 0, d_src, d_mask, w, h, iw, ih,
                                                                         d_img + plane * y0, d_mimg + plane * y0);
        ++launches;
      }
      k_cs_blur<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_mimg, iw, ih, ns, d_blur);
      k_cs_compose<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_plans, s0, ns, d_img, d_blur, iw, ih, win_w, win_h, bgcolor, d_out);
      launches += 2;
      CS_CUDA(cudaGetLastError());
    }
    ctx->launches += launches;
    if (ctx->timing_enabled) ctx->timers["createsamples"].launches += launches;
  }
  CS_CUDA(cudaMemcpyAsync(out_samples, d_out, (size_t)count * win_w * win_h, cudaMemcpyDeviceToHost, ctx->stream));
  CS_CUDA(cudaStreamSynchronize(ctx->stream));
#undef CS_CUDA
  cleanup();
  return CC_OK;
}

}  // extern "C"

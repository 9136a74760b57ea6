// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's feature path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this code.  The product (libcascade_cuda.so) never links or calls it.
//
// Parity status: pinned against (a) the reference's own known answers
// (traincascade/test/test_features.cpp:462-560, flat-image invariants :252-417),
// (b) the published feature counts (traincascade/res/README.md:41,91) and
// (c) python cv2 4.13 for the OpenCV-owned arithmetic (integral3, cartToPolar);
// see tests/test_oracle_*.py and tests/golden/.  Training results are NOT pinned
// by any reference test ("parity unpinned" for chosen splits — SURVEY.md §0).
//
// All path:line citations are relative to /root/reference/traincascade/lib/.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

enum FeatureType { FT_HAAR = 0, FT_LBP = 1, FT_HOG = 2 };   // traincascade_features.h:141
enum HaarMode { HM_BASIC = 0, HM_CORE = 1, HM_ALL = 2 };    // haarfeatures.h:36

static inline int round_half_even(double v) { return (int)std::nearbyint(v); }  // cvRound

// ----------------------------------------------------------------------------------
// Integral images of one W x H u8 window (cv::integral as used by
// src/haarfeatures.cpp:100-114 and src/lbpfeatures.cpp:22-28).
//   sum    : (H+1) x (W+1) int32, first row / column zero
//   sqsum  : same shape, double (exact integers)
//   tilted : same shape, int32; T(Y,X) = sum of I(y,x) with y < Y and |x-(X-1)| <= Y-1-y
// Any output pointer may be null.
// ----------------------------------------------------------------------------------
inline void integral_u8(const uint8_t* img, int W, int H, int32_t* sum, double* sqsum,
                        int32_t* tilted) {
  const int S = W + 1;
  if (sum) {
    for (int x = 0; x <= W; ++x) sum[x] = 0;
    for (int y = 0; y < H; ++y) {
      int32_t run = 0;
      sum[(y + 1) * S] = 0;
      for (int x = 0; x < W; ++x) {
        run += img[y * W + x];
        sum[(y + 1) * S + x + 1] = sum[y * S + x + 1] + run;
      }
    }
  }
  if (sqsum) {
    for (int x = 0; x <= W; ++x) sqsum[x] = 0;
    for (int y = 0; y < H; ++y) {
      double run = 0;
      sqsum[(y + 1) * S] = 0;
      for (int x = 0; x < W; ++x) {
        double v = img[y * W + x];
        run += v * v;
        sqsum[(y + 1) * S + x + 1] = sqsum[y * S + x + 1] + run;
      }
    }
  }
  if (tilted) {
    // direct evaluation of the cone definition through per-row prefix sums
    std::vector<int32_t> pre((size_t)H * (W + 1), 0);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) pre[y * (W + 1) + x + 1] = pre[y * (W + 1) + x] + img[y * W + x];
    for (int Y = 0; Y <= H; ++Y)
      for (int X = 0; X <= W; ++X) {
        int32_t acc = 0;
        for (int y = 0; y < Y; ++y) {
          int half = Y - 1 - y;
          int lo = std::max(0, X - 1 - half), hi = std::min(W - 1, X - 1 + half);
          if (lo <= hi) acc += pre[y * (W + 1) + hi + 1] - pre[y * (W + 1) + lo];
        }
        tilted[Y * S + X] = acc;
      }
  }
}

// calcNormFactor, src/features.cpp:13-25 : inner rect (1,1,W-2,H-2)
inline float norm_factor(const int32_t* sum, const double* sqsum, int W, int H) {
  const int S = W + 1;
  const int rx = 1, ry = 1, rw = W - 2, rh = H - 2;
  const int p0 = rx + S * ry, p1 = rx + rw + S * ry, p2 = rx + S * (ry + rh),
            p3 = rx + rw + S * (ry + rh);
  double area = rw * rh;
  int vs = sum[p0] - sum[p1] - sum[p2] + sum[p3];
  double vq = sqsum[p0] - sqsum[p1] - sqsum[p2] + sqsum[p3];
  return (float)std::sqrt((double)(area * vq - (double)vs * vs));
}

// ----------------------------------------------------------------------------------
// Feature catalogs
// ----------------------------------------------------------------------------------
struct Rect4 { int x, y, w, h; };

struct HaarFeature {            // haarfeatures.h:80-101
  int tilted;
  Rect4 r[3];
  float wt[3];
  int p[3][4];                  // corner offsets, row stride W+1
};

inline void upright_offsets(const Rect4& r, int S, int* p) {   // CV_SUM_OFFSETS, traincascade_features.h:40-51
  p[0] = r.x + S * r.y;
  p[1] = r.x + r.w + S * r.y;
  p[2] = r.x + S * (r.y + r.h);
  p[3] = r.x + r.w + S * (r.y + r.h);
}
inline void tilted_offsets(const Rect4& r, int S, int* p) {    // CV_TILTED_OFFSETS, traincascade_features.h:54-63
  p[0] = r.x + S * r.y;
  p[1] = r.x - r.h + S * (r.y + r.h);
  p[2] = r.x + r.w + S * (r.y + r.w);
  p[3] = r.x + r.w - r.h + S * (r.y + r.w + r.h);
}

inline HaarFeature make_haar(int S, bool tilted, Rect4 a, float wa, Rect4 b, float wb,
                             Rect4 c = {0, 0, 0, 0}, float wc = 0.f) {   // haarfeatures.cpp:266-309
  HaarFeature f;
  std::memset(&f, 0, sizeof(f));
  f.tilted = tilted;
  f.r[0] = a; f.r[1] = b; f.r[2] = c;
  f.wt[0] = wa; f.wt[1] = wb; f.wt[2] = wc;
  for (int j = 0; j < 3; ++j) {
    if (f.wt[j] == 0.f) break;
    if (tilted) tilted_offsets(f.r[j], S, f.p[j]); else upright_offsets(f.r[j], S, f.p[j]);
  }
  return f;
}

// generateFeatures, src/haarfeatures.cpp:127-251.  Enumeration order x -> y -> dx -> dy -> shape.
inline std::vector<HaarFeature> haar_catalog(int W, int H, int mode) {
  std::vector<HaarFeature> out;
  const int S = W + 1;
  for (int x = 0; x < W; ++x)
    for (int y = 0; y < H; ++y)
      for (int dx = 1; dx <= W; ++dx)
        for (int dy = 1; dy <= H; ++dy) {
          auto fits = [&](int mx, int my) { return x + dx * mx <= W && y + dy * my <= H; };
          if (fits(2, 1)) out.push_back(make_haar(S, false, {x, y, dx * 2, dy}, -1.f, {x + dx, y, dx, dy}, 2.f));
          if (fits(1, 2)) out.push_back(make_haar(S, false, {x, y, dx, dy * 2}, -1.f, {x, y + dy, dx, dy}, 2.f));
          if (fits(3, 1)) out.push_back(make_haar(S, false, {x, y, dx * 3, dy}, -1.f, {x + dx, y, dx, dy}, 2.f));
          if (fits(1, 3)) out.push_back(make_haar(S, false, {x, y, dx, dy * 3}, -1.f, {x, y + dy, dx, dy}, 2.f));
          if (mode != HM_BASIC) {
            if (fits(4, 1)) out.push_back(make_haar(S, false, {x, y, dx * 4, dy}, -1.f, {x + dx, y, dx * 2, dy}, 2.f));
            if (fits(1, 4)) out.push_back(make_haar(S, false, {x, y, dx, dy * 4}, -1.f, {x, y + dy, dx, dy * 2}, 2.f));
          }
          if (fits(2, 2))
            out.push_back(make_haar(S, false, {x, y, dx * 2, dy * 2}, -1.f, {x, y, dx, dy}, 2.f,
                                    {x + dx, y + dy, dx, dy}, 2.f));
          if (mode != HM_BASIC && fits(3, 3))
            out.push_back(make_haar(S, false, {x, y, dx * 3, dy * 3}, -1.f, {x + dx, y + dy, dx, dy}, 9.f));
          if (mode == HM_ALL) {
            if (x + 2 * dx <= W && y + 2 * dx + dy <= H && x - dy >= 0)
              out.push_back(make_haar(S, true, {x, y, dx * 2, dy}, -1.f, {x, y, dx, dy}, 2.f));
            if (x + dx <= W && y + dx + 2 * dy <= H && x - 2 * dy >= 0)
              out.push_back(make_haar(S, true, {x, y, dx, 2 * dy}, -1.f, {x, y, dx, dy}, 2.f));
            if (x + 3 * dx <= W && y + 3 * dx + dy <= H && x - dy >= 0)
              out.push_back(make_haar(S, true, {x, y, dx * 3, dy}, -1.f, {x + dx, y + dx, dx, dy}, 3.f));
            if (x + dx <= W && y + dx + 3 * dy <= H && x - 3 * dy >= 0)
              out.push_back(make_haar(S, true, {x, y, dx, 3 * dy}, -1.f, {x - dy, y + dy, dx, dy}, 3.f));
            if (x + 4 * dx <= W && y + 4 * dx + dy <= H && x - dy >= 0)
              out.push_back(make_haar(S, true, {x, y, dx * 4, dy}, -1.f, {x + dx, y + dx, dx * 2, dy}, 2.f));
            if (x + dx <= W && y + dx + 4 * dy <= H && x - 4 * dy >= 0)
              out.push_back(make_haar(S, true, {x, y, dx, 4 * dy}, -1.f, {x - dy, y + dy, dx, 2 * dy}, 2.f));
          }
        }
  return out;
}

// Feature::calc, haarfeatures.h:114-122 (int rect sums, float weights)
inline float haar_calc(const HaarFeature& f, const int32_t* img) {
  float ret = f.wt[0] * (float)(img[f.p[0][0]] - img[f.p[0][1]] - img[f.p[0][2]] + img[f.p[0][3]]) +
              f.wt[1] * (float)(img[f.p[1][0]] - img[f.p[1][1]] - img[f.p[1][2]] + img[f.p[1][3]]);
  if (f.wt[2] != 0.0f)
    ret += f.wt[2] * (float)(img[f.p[2][0]] - img[f.p[2][1]] - img[f.p[2][2]] + img[f.p[2][3]]);
  return ret;
}

struct LbpFeature {             // lbpfeatures.h:55-66
  Rect4 r;                      // one cell; the feature covers 3x3 cells
  int p[16];                    // 4x4 grid of corner offsets, row-major
};

// generateFeatures, src/lbpfeatures.cpp:35-45 ; Feature ctor :52-63
inline std::vector<LbpFeature> lbp_catalog(int W, int H) {
  std::vector<LbpFeature> out;
  const int S = W + 1;
  for (int x = 0; x < W; ++x)
    for (int y = 0; y < H; ++y)
      for (int w = 1; w <= W / 3; ++w)
        for (int h = 1; h <= H / 3; ++h)
          if (x + 3 * w <= W && y + 3 * h <= H) {
            LbpFeature f;
            f.r = {x, y, w, h};
            for (int gy = 0; gy < 4; ++gy)
              for (int gx = 0; gx < 4; ++gx) f.p[gy * 4 + gx] = (x + gx * w) + S * (y + gy * h);
            out.push_back(f);
          }
  return out;
}

// Feature::calc, lbpfeatures.h:70-83 : 8 neighbours >= centre, bit order 0,1,2,5,8,7,6,3
inline uint8_t lbp_calc(const LbpFeature& f, const int32_t* s) {
  auto cell = [&](int gx, int gy) {
    const int a = gy * 4 + gx;
    return s[f.p[a]] - s[f.p[a + 1]] - s[f.p[a + 4]] + s[f.p[a + 5]];
  };
  const int c = cell(1, 1);
  return (uint8_t)((cell(0, 0) >= c ? 128 : 0) | (cell(1, 0) >= c ? 64 : 0) | (cell(2, 0) >= c ? 32 : 0) |
                   (cell(2, 1) >= c ? 16 : 0) | (cell(2, 2) >= c ? 8 : 0) | (cell(1, 2) >= c ? 4 : 0) |
                   (cell(0, 2) >= c ? 2 : 0) | (cell(0, 1) >= c ? 1 : 0));
}

constexpr int HOG_BINS = 9, HOG_CELLS = 4;    // HOGfeatures.h:24-26

struct HogFeature {             // HOGfeatures.h:70-83
  Rect4 cell[4];
  int p[4][4];
};

// generateFeatures, src/HOGfeatures.cpp:67-106
inline std::vector<HogFeature> hog_catalog(int W, int H) {
  std::vector<HogFeature> out;
  const int S = W + 1;
  auto add = [&](int x, int y, int cw, int ch) {
    HogFeature f;
    f.cell[0] = {x, y, cw, ch};
    f.cell[1] = {x + cw, y, cw, ch};
    f.cell[2] = {x, y + ch, cw, ch};
    f.cell[3] = {x + cw, y + ch, cw, ch};
    for (int i = 0; i < 4; ++i) upright_offsets(f.cell[i], S, f.p[i]);
    out.push_back(f);
  };
  for (int t = 8; t <= W / 2; t += 8) {
    for (int x = 0; x <= W - 2 * t; x += 4) for (int y = 0; y <= H - 2 * t; y += 4) add(x, y, t, t);
    for (int x = 0; x <= W - 2 * t; x += 4) for (int y = 0; y <= H - 4 * t; y += 4) add(x, y, t, 2 * t);
    for (int x = 0; x <= W - 4 * t; x += 4) for (int y = 0; y <= H - 2 * t; y += 4) add(x, y, 2 * t, t);
  }
  return out;
}

// OpenCV's polynomial atan2 (cv::fastAtan32f, OpenCV 4.x modules/core/src/mathfuncs_core.simd.hpp);
// third-party dependency of src/HOGfeatures.cpp:207 (cartToPolar, angleInDegrees=false).
// Restated from the published algorithm; compared with cv2.cartToPolar in tests/test_oracle_cv2.py.
inline float cv_fast_atan2_rad(float y, float x) {
  const float deg = (float)(180.0 / M_PI);
  const float p1 = 0.9997878412794807f * deg, p3 = -0.3258083974640975f * deg,
              p5 = 0.1555786518463281f * deg, p7 = -0.04432655554792128f * deg;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a * (float)(M_PI / 180.0);
}

// integralHistogram, src/HOGfeatures.cpp:163-256.  hist[b] and norm are (H+1)x(W+1) float planes.
inline void hog_integral_histogram(const uint8_t* img, int W, int H, float* const* hist, float* norm) {
  const int S = W + 1;
  std::vector<float> grad((size_t)W * H);
  std::vector<uint8_t> qangle((size_t)W * H);
  const float angleScale = (float)(HOG_BINS / M_PI);
  auto clampi = [](int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); };   // BORDER_REPLICATE
  for (int y = 0; y < H; ++y) {
    const uint8_t* cur = img + (size_t)y * W;
    const uint8_t* prev = img + (size_t)clampi(y - 1, H) * W;
    const uint8_t* next = img + (size_t)clampi(y + 1, H) * W;
    for (int x = 0; x < W; ++x) {
      float dx = (float)(cur[clampi(x + 1, W)] - cur[clampi(x - 1, W)]);
      float dy = (float)(next[x] - prev[x]);
      float mag = std::sqrt(dx * dx + dy * dy);
      float angle = cv_fast_atan2_rad(dy, dx);
      angle = angle * angleScale - 0.5f;
      int bidx = (int)std::floor(angle);
      if (bidx < 0) bidx += HOG_BINS; else if (bidx >= HOG_BINS) bidx -= HOG_BINS;
      qangle[(size_t)y * W + x] = (uint8_t)bidx;
      grad[(size_t)y * W + x] = mag;
    }
  }
  // cv::integral(grad, norm, CV_32F): sequential row-sum + previous-row order
  for (int x = 0; x <= W; ++x) norm[x] = 0.f;
  for (int y = 0; y < H; ++y) {
    float run = 0.f;
    norm[(y + 1) * S] = 0.f;
    for (int x = 0; x < W; ++x) {
      run += grad[(size_t)y * W + x];
      norm[(y + 1) * S + x + 1] = norm[y * S + x + 1] + run;
    }
  }
  for (int b = 0; b < HOG_BINS; ++b) {
    float* hb = hist[b];
    for (int x = 0; x <= W; ++x) hb[x] = 0.f;
    for (int y = 0; y < H; ++y) {
      float str = 0.f;
      hb[(y + 1) * S] = 0.f;
      for (int x = 0; x < W; ++x) {
        if (qangle[(size_t)y * W + x] == b) str += grad[(size_t)y * W + x];
        hb[(y + 1) * S + x + 1] = hb[y * S + x + 1] + str;
      }
    }
  }
}

// ----------------------------------------------------------------------------------
// Evaluator : CvFeatureEvaluator + Cv{Haar,LBP,HOG}Evaluator
// (traincascade_features.h:155-188, src/features.cpp:72-89, src/haarfeatures.cpp:89-114,
//  src/lbpfeatures.cpp:15-28, src/HOGfeatures.cpp:17-42)
// ----------------------------------------------------------------------------------
struct Evaluator {
  int type = FT_HAAR, mode = HM_BASIC, W = 0, H = 0, P = 0, maxSamples = 0;
  int numFeatures = 0, featSize = 1, maxCatCount = 0;
  std::vector<float> cls;
  std::vector<HaarFeature> haar;
  std::vector<LbpFeature> lbp;
  std::vector<HogFeature> hog;
  std::vector<int32_t> sum, tilted;      // N x P
  std::vector<float> nf;                 // N
  std::vector<float> hist[HOG_BINS];     // each N x P
  std::vector<float> normSum;            // N x P

  void init(int type_, int mode_, int W_, int H_, int maxSamples_) {
    type = type_; mode = mode_; W = W_; H = H_; maxSamples = maxSamples_;
    P = (W + 1) * (H + 1);
    cls.assign(maxSamples, 0.f);
    if (type == FT_HAAR) {
      sum.assign((size_t)maxSamples * P, 0);
      tilted.assign((size_t)maxSamples * P, 0);
      nf.assign(maxSamples, 0.f);
      haar = haar_catalog(W, H, mode);
      numFeatures = (int)haar.size(); featSize = 1; maxCatCount = 0;
    } else if (type == FT_LBP) {
      sum.assign((size_t)maxSamples * P, 0);
      lbp = lbp_catalog(W, H);
      numFeatures = (int)lbp.size(); featSize = 1; maxCatCount = 256;
    } else {
      for (auto& h : hist) h.assign((size_t)maxSamples * P, 0.f);
      normSum.assign((size_t)maxSamples * P, 0.f);
      hog = hog_catalog(W, H);
      numFeatures = (int)hog.size(); featSize = HOG_BINS * HOG_CELLS; maxCatCount = 0;
    }
  }
  int varCount() const { return numFeatures * featSize; }

  void setImage(const uint8_t* img, uint8_t label, int idx) {
    cls[idx] = label;
    if (type == FT_HAAR) {
      std::vector<double> sq(P);
      integral_u8(img, W, H, &sum[(size_t)idx * P], sq.data(),
                  mode == HM_ALL ? &tilted[(size_t)idx * P] : nullptr);
      nf[idx] = norm_factor(&sum[(size_t)idx * P], sq.data(), W, H);
    } else if (type == FT_LBP) {
      integral_u8(img, W, H, &sum[(size_t)idx * P], nullptr, nullptr);
    } else {
      float* hp[HOG_BINS];
      for (int b = 0; b < HOG_BINS; ++b) hp[b] = &hist[b][(size_t)idx * P];
      hog_integral_histogram(img, W, H, hp, &normSum[(size_t)idx * P]);
    }
  }

  // operator()(featureIdx, sampleIdx): haarfeatures.h:108-112, lbpfeatures.h:44-45, HOGfeatures.h:84-108
  float operator()(int vi, int si) const {
    if (type == FT_HAAR) {
      float n = nf[si];
      const HaarFeature& f = haar[vi];
      const int32_t* img = f.tilted ? &tilted[(size_t)si * P] : &sum[(size_t)si * P];
      return !n ? 0.0f : (haar_calc(f, img) / n);
    } else if (type == FT_LBP) {
      return (float)lbp_calc(lbp[vi], &sum[(size_t)si * P]);
    } else {
      const int fi = vi / (HOG_BINS * HOG_CELLS), comp = vi % (HOG_BINS * HOG_CELLS);
      const int bin = comp % HOG_BINS, cellIdx = comp / HOG_BINS;
      const HogFeature& f = hog[fi];
      const float* ph = &hist[bin][(size_t)si * P];
      float res = ph[f.p[cellIdx][0]] - ph[f.p[cellIdx][1]] - ph[f.p[cellIdx][2]] + ph[f.p[cellIdx][3]];
      const float* pn = &normSum[(size_t)si * P];
      float normFactor = (float)(pn[f.p[0][0]] - pn[f.p[1][1]] - pn[f.p[2][2]] + pn[f.p[3][3]]);
      res = (res > 0.001f) ? (res / (normFactor + 0.001f)) : 0.f;
      return res;
    }
  }
};

}  // namespace orc

// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's sample readers,
// hard-negative mining loop and cascade driver.  See orc_features.hpp for usage rules.
//
// Third-party arithmetic restated here: cv::resize(..., INTER_LINEAR_EXACT) for 8-bit
// single-channel images (OpenCV 4.x, modules/imgproc/src/resize.cpp, fixed-point 8.8 path;
// OpenCV is not vendored in /root/reference, version unpinned — SURVEY.md §8c).  Pinned
// bit-exactly against python cv2 4.13 in tests/test_oracle_cv2.py and tests/golden/.
#pragma once
#include <cstdio>
#include <string>

#include "orc_boost.hpp"

namespace orc {

// ---- INTER_LINEAR_EXACT, SURVEY App. A.6 -------------------------------------------------
struct LinCoef { int ofs; int c0, c1; };   // 8.8 fixed point, c0 + c1 = 256

inline void linear_exact_coeffs(int src, int dst, std::vector<LinCoef>& co, int& minofs, int& maxofs) {
  co.assign(dst, LinCoef{0, 256, 0});
  minofs = 0; maxofs = dst;
  const double inv = (double)dst / src;
  const double scale = 1.0 / inv;
  for (int v = 0; v < dst; ++v) {
    double f = scale * (v + 0.5) - 0.5;
    int i = (int)std::floor(f);
    f -= i;
    if (i >= 0 && src > 1) {
      if (i < src - 1) {
        co[v].ofs = i;
        co[v].c1 = round_half_even(f * 256.0);
        co[v].c0 = 256 - co[v].c1;
      } else {
        co[v].ofs = src - 1;
        maxofs = std::min(maxofs, v);
      }
    } else {
      minofs = std::max(minofs, v + 1);
    }
  }
}

inline void resize_linear_exact(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  std::vector<LinCoef> cx, cy;
  int xmin, xmax, ymin, ymax;
  linear_exact_coeffs(sw, dw, cx, xmin, xmax);
  linear_exact_coeffs(sh, dh, cy, ymin, ymax);
  auto hline = [&](const uint8_t* s, std::vector<uint32_t>& out) {
    for (int x = 0; x < dw; ++x) {
      if (x < xmin) out[x] = (uint32_t)s[0] * 256;
      else if (x >= xmax) out[x] = (uint32_t)s[sw - 1] * 256;
      else out[x] = (uint32_t)s[cx[x].ofs] * cx[x].c0 + (uint32_t)s[cx[x].ofs + 1] * cx[x].c1;
    }
  };
  std::vector<uint32_t> r0(dw), r1(dw);
  for (int y = 0; y < dh; ++y) {
    uint8_t* d = dst + (size_t)y * dw;
    if (y < ymin) {
      hline(src, r0);
      for (int x = 0; x < dw; ++x) d[x] = (uint8_t)((r0[x] * 256 + 32768) >> 16);
    } else if (y >= ymax) {
      hline(src + (size_t)(sh - 1) * sw, r0);
      for (int x = 0; x < dw; ++x) d[x] = (uint8_t)((r0[x] * 256 + 32768) >> 16);
    } else {
      hline(src + (size_t)cy[y].ofs * sw, r0);
      hline(src + (size_t)(cy[y].ofs + 1) * sw, r1);
      for (int x = 0; x < dw; ++x) d[x] = (uint8_t)((r0[x] * cy[y].c0 + r1[x] * cy[y].c1 + 32768) >> 16);
    }
  }
}

// ---- image sources ----------------------------------------------------------------------
struct GrayImage { int w = 0, h = 0; std::vector<uint8_t> px; };

// CvCascadeImageReader::NegReader, src/imagestorage.cpp:23-126 (images already decoded to gray)
struct NegReader {
  const std::vector<GrayImage>* images = nullptr;
  int winW = 0, winH = 0;
  const GrayImage* src = nullptr;
  GrayImage img;
  int pointX = 0, pointY = 0, offX = 0, offY = 0;
  float scale = 1.0f;
  const float scaleFactor = 1.4142135623730950488016887242097F;
  const float stepFactor = 0.5F;
  size_t last = 0, round = 0;

  void create(const std::vector<GrayImage>* imgs, int w, int h) {
    images = imgs; winW = w; winH = h; last = round = 0; src = nullptr; img = GrayImage();
  }
  bool nextImg() {
    int ox = 0, oy = 0;
    size_t count = images->size();
    src = nullptr;
    for (size_t i = 0; i < count; ++i) {
      src = &(*images)[last++];
      if (src->px.empty()) { last %= count; src = nullptr; continue; }
      round += last / count;
      round = round % ((size_t)winW * winH);
      last %= count;
      ox = std::min((int)round % winW, src->w - winW);
      oy = std::min((int)round / winW, src->h - winH);
      if (ox >= 0 && oy >= 0) break;
    }
    if (!src) return false;
    if (ox < 0 || oy < 0) return false;   // the reference reads out of bounds here (UB); the oracle refuses
    pointX = offX = ox; pointY = offY = oy;
    scale = std::max(((float)winW + pointX) / ((float)src->w), ((float)winH + pointY) / ((float)src->h));
    int sw = (int)(scale * src->w + 0.5F), sh = (int)(scale * src->h + 0.5F);
    img.w = sw; img.h = sh; img.px.resize((size_t)sw * sh);
    resize_linear_exact(src->px.data(), src->w, src->h, img.px.data(), sw, sh);
    return true;
  }
  bool get(uint8_t* out) {
    if (img.px.empty())
      if (!nextImg()) return false;
    for (int r = 0; r < winH; ++r)
      std::memcpy(out + (size_t)r * winW, &img.px[(size_t)(pointY + r) * img.w + pointX], winW);
    if ((int)(pointX + (1.0F + stepFactor) * winW) < img.w)
      pointX += (int)(stepFactor * winW);
    else {
      pointX = offX;
      if ((int)(pointY + (1.0F + stepFactor) * winH) < img.h)
        pointY += (int)(stepFactor * winH);
      else {
        pointY = offY;
        scale *= scaleFactor;
        if (scale <= 1.0F) {
          int sw = (int)(scale * src->w), sh = (int)(scale * src->h);
          img.w = sw; img.h = sh; img.px.resize((size_t)sw * sh);
          resize_linear_exact(src->px.data(), src->w, src->h, img.px.data(), sw, sh);
        } else {
          if (!nextImg()) return false;
        }
      }
    }
    return true;
  }
};

// CvCascadeImageReader::PosReader, src/imagestorage.cpp:138-189 (.vec payload already in memory)
struct PosReader {
  const uint8_t* samples = nullptr;   // count x vecSize u8
  int count = 0, vecSize = 0, last = 0;
  void create(const uint8_t* s, int cnt, int vsz) { samples = s; count = cnt; vecSize = vsz; last = 0; }
  void restart() { last = 0; }
  bool get(uint8_t* out) {
    if (last >= count) return false;   // reference raises "vec-file is over"
    std::memcpy(out, samples + (size_t)last * vecSize, vecSize);
    ++last;
    return true;
  }
};

// ---- cascade of trained stages evaluated directly on an Evaluator slot ----------------------
struct CascadeStage {
  float threshold = 0.f;
  std::vector<Tree> weak;
};

inline const TreeNode& tree_predict(const Tree& t, const Evaluator& ev, int si) {   // o_cvcascadeboosttree.cpp:16-39
  int n = 0;
  if (ev.maxCatCount == 0) {
    while (t.nodes[n].left >= 0) {
      float val = ev(t.nodes[n].var_idx, si);
      n = val <= t.nodes[n].c ? t.nodes[n].left : t.nodes[n].right;
    }
  } else {
    while (t.nodes[n].left >= 0) {
      int c = (int)ev(t.nodes[n].var_idx, si);
      n = cat_dir(c, t.nodes[n].subset) < 0 ? t.nodes[n].left : t.nodes[n].right;
    }
  }
  return t.nodes[n];
}
inline bool stage_pass(const CascadeStage& s, const Evaluator& ev, int si) {   // boost.cpp:461-477
  double sum = 0;
  for (const Tree& t : s.weak) sum += tree_predict(t, ev, si).value;
  return !(sum < s.threshold - 0.00001F);
}

struct StageReport {
  int posCount = 0, negCount = 0;
  long long posConsumed = 0, negConsumed = 0;
  double acceptanceRatio = 0;
  int trained = 0;
};

// CvCascadeClassifier (training half), src/cascadeclassifier.cpp:137-357
struct CascadeTrainer {
  Evaluator ev;
  PosReader pos;
  NegReader neg;
  std::vector<CascadeStage> stages;
  std::vector<StageReport> reports;
  std::vector<Stage> stageDetails;
  int numPos = 0, numNeg = 0, numStages = 0, curNumSamples = 0;
  BoostParams bp;
  int sortMode = SORT_STD, threads = 1;
  double acceptanceRatioBreakValue = -1;
  int termination = 0;   // 0 all stages, 1 set not filled, 2 leaf FA reached, 3 acceptance break, 4 stage not trained

  int predict(int si) const {                   // :297-306
    for (const CascadeStage& s : stages)
      if (!stage_pass(s, ev, si)) return 0;
    return 1;
  }
  int fillPassedSamples(int first, int count, bool isPositive, double minRatio, long long& consumed) {   // :329-357
    int getcount = 0;
    std::vector<uint8_t> img((size_t)ev.W * ev.H);
    for (int i = first; i < first + count; ++i) {
      for (;;) {
        if (consumed != 0 && ((double)getcount + 1) / (double)consumed <= minRatio) return getcount;
        bool ok = isPositive ? pos.get(img.data()) : neg.get(img.data());
        if (!ok) return getcount;
        consumed++;
        ev.setImage(img.data(), isPositive ? 1 : 0, i);
        if (predict(i) == 1) { getcount++; break; }
      }
    }
    return getcount;
  }
  bool updateTrainingSet(double minRatio, double& acceptanceRatio, StageReport& rep) {   // :308-327
    long long posConsumed = 0, negConsumed = 0;
    pos.restart();
    int posCount = fillPassedSamples(0, numPos, true, 0, posConsumed);
    rep.posCount = posCount; rep.posConsumed = posConsumed;
    if (!posCount) return false;
    int proNumNeg = round_half_even((((double)numNeg) * ((double)posCount)) / numPos);
    int negCount = fillPassedSamples(posCount, proNumNeg, false, minRatio, negConsumed);
    rep.negCount = negCount; rep.negConsumed = negConsumed;
    if (!negCount)
      if (!(negConsumed > 0 && ((double)negCount + 1) / (double)negConsumed <= minRatio)) return false;
    curNumSamples = posCount + negCount;
    acceptanceRatio = negConsumed == 0 ? 0 : ((double)negCount / (double)negConsumed);
    rep.acceptanceRatio = acceptanceRatio;
    return true;
  }
  void train() {                                 // :137-295
    double requiredLeafFARate = std::pow((double)bp.maxFalseAlarm, (double)numStages) / (double)bp.max_depth;
    double tempLeafFARate = NAN;
    termination = 0;
    for (int i = (int)stages.size(); i < numStages; ++i) {
      StageReport rep;
      if (!updateTrainingSet(requiredLeafFARate, tempLeafFARate, rep)) { reports.push_back(rep); termination = 1; break; }
      if (tempLeafFARate <= requiredLeafFARate) { reports.push_back(rep); termination = 2; break; }
      if (tempLeafFARate <= acceptanceRatioBreakValue && acceptanceRatioBreakValue >= 0) { reports.push_back(rep); termination = 3; break; }
      TrainData td;
      td.precalculate(&ev, curNumSamples, sortMode, threads);
      BoostTrainer bt;
      bool ok = bt.train(&td, bp, sortMode);
      if (!ok) { reports.push_back(rep); termination = 4; break; }
      rep.trained = 1;
      reports.push_back(rep);
      CascadeStage cs; cs.threshold = bt.stage.threshold; cs.weak = bt.stage.weak;
      stages.push_back(cs);
      stageDetails.push_back(bt.stage);
    }
  }
};

}  // namespace orc

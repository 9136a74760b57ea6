// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's stage trainer
// (CvCascadeBoostTrainData / CvCascadeBoost / CvCascadeBoostTree / CvBoostTree / CvDTree).
// See orc_features.hpp for the usage rules and the parity status ("training results are
// parity-unpinned by the reference's own tests"; the restatement follows the files cited).
//
// Simplifications that do not change results:
//  * every feature row is cached (numPrecalcVal = numPrecalcIdx = var_count), i.e. the
//    "-precalcValBufSize/-precalcIdxBufSize large enough" setting of the reference;
//  * node-local sorted index lists are obtained by stably filtering the per-feature global
//    order, which is what the chain subsample_data -> split_node_data produces
//    (src/o_cvcascadeboosttraindata.cpp:80-121, src/o_cvcascadeboosttree.cpp:192-253).
#pragma once
#include <atomic>
#include <functional>
#include <thread>

#include "orc_features.hpp"

namespace orc {

enum BoostType { BT_DISCRETE = 0, BT_REAL = 1, BT_LOGIT = 2, BT_GENTLE = 3 };   // cv::ml::Boost
enum SortMode { SORT_STD = 0, SORT_STABLE = 1 };   // std::sort (faithful, libstdc++) | stable-by-index (canonical)

struct BoostParams {            // boost.cpp:41-56, o_cvboostparams.cpp:7-27, o_cvdtreeparams.cpp:5-14
  int boost_type = BT_GENTLE;
  int weak_count = 100;
  double weight_trim_rate = 0.95;
  int max_depth = 1;
  int min_sample_count = 10;
  float regression_accuracy = 0.01f;
  float minHitRate = 0.995f;
  float maxFalseAlarm = 0.5f;
};

struct TreeNode {
  int left = -1, right = -1, parent = -1, depth = 0;
  int var_idx = -1;
  float c = 0.f;               // ordered threshold
  int subset[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // categorical left-subset bitmask
  float quality = 0.f;
  int split_point = -1;
  double value = 0.;
  double node_risk = 0.;
  int sample_count = 0;
};
struct Tree { std::vector<TreeNode> nodes; };   // nodes[0] = root

struct StageLogRow { int n; float hr, fa; };

struct Stage {
  float threshold = 0.f;
  std::vector<Tree> weak;
  std::vector<StageLogRow> log;
};

static inline double log_ratio(double val) {   // boost.cpp:27-36, o_cvboostree.cpp:11-17
  const double eps = 1e-5;
  val = std::max(val, eps);
  val = std::min(val, 1. - eps);
  return std::log(val / (1. - val));
}
static inline int cat_dir(int idx, const int* subset) {   // CV_DTREE_CAT_DIR, o_utils.cpp:14-16
  return 2 * ((subset[idx >> 5] & (1 << (idx & 31))) == 0) - 1;
}

template <class F>
inline void parallel_rows(int begin, int end, int threads, F body) {   // cv::parallel_for_ shape
  if (threads <= 1 || end - begin < 2) { body(begin, end); return; }
  std::atomic<int> next(begin);
  const int grain = std::max(1, (end - begin) / (threads * 16));
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] {
      for (;;) {
        int b = next.fetch_add(grain);
        if (b >= end) break;
        body(b, std::min(end, b + grain));
      }
    });
  for (auto& th : pool) th.join();
}

// ----------------------------------------------------------------------------------
// TrainData : valCache + sorted index buffer (o_cvcascadeboosttraindata.cpp:214-368, :490-596)
// ----------------------------------------------------------------------------------
struct TrainData {
  const Evaluator* ev = nullptr;
  int N = 0, F = 0;
  bool categorical = false;
  std::vector<float> fval;          // F x N, feature-major ("valCache")
  std::vector<uint16_t> sidx;       // F x N sorted sample indices (ordered variables only), is_buf_16u
  std::vector<uint32_t> sidx32;     // the same buffer when is_buf_16u == false: N >= 65536 (o_cvcascadeboosttraindata.cpp:250-251,283-309)
  bool is_buf_16u = true;
  uint32_t sorted_at(int row, int i) const { return is_buf_16u ? (uint32_t)sidx[(size_t)row * N + i] : sidx32[(size_t)row * N + i]; }
  std::vector<int> feat_subset;     // optional: evaluate only these variables (bench sampling); empty = all

  int var(int row) const { return feat_subset.empty() ? row : feat_subset[row]; }

  void precalculate(const Evaluator* e, int numSamples, int sortMode, int threads,
                    const int* subset = nullptr, int nsubset = 0) {
    ev = e; N = numSamples;
    categorical = e->maxCatCount > 0;
    if (subset) feat_subset.assign(subset, subset + nsubset); else feat_subset.clear();
    F = subset ? nsubset : e->varCount();
    fval.assign((size_t)F * N, 0.f);
    is_buf_16u = N < 65536;
    sidx.clear(); sidx32.clear();
    if (!categorical) { if (is_buf_16u) sidx.assign((size_t)F * N, 0); else sidx32.assign((size_t)F * N, 0); }
    parallel_rows(0, F, threads, [&](int b, int en) {
      for (int fi = b; fi < en; ++fi) {
        float* row = &fval[(size_t)fi * N];
        const int vi = var(fi);
        for (int si = 0; si < N; ++si) row[si] = (*ev)(vi, si);
        if (!categorical && is_buf_16u) {
          uint16_t* idx = &sidx[(size_t)fi * N];
          for (int si = 0; si < N; ++si) idx[si] = (uint16_t)si;
          if (sortMode == SORT_STD)
            std::sort(idx, idx + N, [row](uint16_t a, uint16_t b2) { return row[a] < row[b2]; });   // LessThanIdx, o_utils.h:18-23
          else
            std::stable_sort(idx, idx + N, [row](uint16_t a, uint16_t b2) { return row[a] < row[b2]; });
        } else if (!categorical) {
          uint32_t* idx = &sidx32[(size_t)fi * N];
          for (int si = 0; si < N; ++si) idx[si] = (uint32_t)si;
          if (sortMode == SORT_STD)
            std::sort(idx, idx + N, [row](uint32_t a, uint32_t b2) { return row[a] < row[b2]; });
          else
            std::stable_sort(idx, idx + N, [row](uint32_t a, uint32_t b2) { return row[a] < row[b2]; });
        }
      }
    });
  }
};

// ----------------------------------------------------------------------------------
// BoostTrainer : CvCascadeBoost::train and everything below it
// ----------------------------------------------------------------------------------
struct BoostTrainer {
  const TrainData* data = nullptr;
  BoostParams params;
  int split_criteria = 0;          // 0 GINI, 1 MISCLASS, 2 SQERR  (o_cvboost.cpp:84-92)
  int catSortMode = SORT_STD;
  bool is_classifier = true;
  bool have_subsample = false;
  int N = 0;

  std::vector<double> weights, subtree_weights, weak_eval, sum_response;
  std::vector<int> orig_response;
  std::vector<uint8_t> mask;
  std::vector<float> responses;     // data->responses (cls, or responses_copy for LB/GAB)
  Stage stage;

  // scratch for tree growing
  std::vector<int> pos_in_node;     // sample -> position in current node, or -1
  std::vector<int8_t> dir;

  // hook for tests: called after each weak tree with (tree index)
  std::function<void(int)> on_weak;

  void set_params(const BoostParams& p) {     // o_cvboost.cpp:67-99
    params = p;
    params.weak_count = std::max(params.weak_count, 1);
    params.weight_trim_rate = std::max(params.weight_trim_rate, 0.);
    params.weight_trim_rate = std::min(params.weight_trim_rate, 1.);
    if (params.weight_trim_rate < FLT_EPSILON) params.weight_trim_rate = 1.f;
    if (params.boost_type == BT_DISCRETE) split_criteria = 1;
    else if (params.boost_type == BT_REAL) split_criteria = 0;
    else split_criteria = 2;
  }

  float feature_value(int vi, int si) const { return data->fval[(size_t)vi * N + si]; }   // getVarValue :484-488

  // ---- CvCascadeBoost::update_weights(nullptr), boost.cpp:190-265 ----
  void init_weights() {
    const Evaluator* ev = data->ev;
    orig_response.resize(N); weak_eval.assign(N, 0.); mask.assign(N, 1);
    weights.resize(N); subtree_weights.assign(N + 2, 0.);
    responses.resize(N);
    for (int i = 0; i < N; ++i) responses[i] = ev->cls[i];
    const double w0 = 1. / N;
    for (int i = 0; i < N; ++i) {
      int label = (int)ev->cls[i];
      orig_response[i] = label * 2 - 1;
      weights[i] = w0 * 1.0;
    }
    is_classifier = true;
    if (params.boost_type == BT_LOGIT) {
      sum_response.assign(N, 0.);
      for (int i = 0; i < N; ++i) responses[i] = orig_response[i] > 0 ? 2.f : -2.f;
      is_classifier = false;
    } else if (params.boost_type == BT_GENTLE) {
      for (int i = 0; i < N; ++i) responses[i] = (float)orig_response[i];
      is_classifier = false;
    }
    have_subsample = false;
  }

  // ---- tree prediction, o_cvcascadeboosttree.cpp:16-39 ----
  const TreeNode& predict_node(const Tree& t, int si) const {
    int n = 0;
    if (!data->categorical) {
      while (t.nodes[n].left >= 0) {
        float val = feature_value(t.nodes[n].var_idx, si);
        n = val <= t.nodes[n].c ? t.nodes[n].left : t.nodes[n].right;
      }
    } else {
      while (t.nodes[n].left >= 0) {
        int c = (int)feature_value(t.nodes[n].var_idx, si);
        n = cat_dir(c, t.nodes[n].subset) < 0 ? t.nodes[n].left : t.nodes[n].right;
      }
    }
    return t.nodes[n];
  }

  // ---- CvCascadeBoost::update_weights(tree), boost.cpp:266-406 ----
  void update_weights(Tree& tree) {
    double sumW = 0.;
    if (have_subsample) {
      for (int i = 0; i < N; ++i) mask[i] ^= 1;
      for (int i = 0; i < N; ++i)
        if (mask[i]) weak_eval[i] = predict_node(tree, i).value;
    }
    if (params.boost_type == BT_DISCRETE) {
      double C, err = 0.;
      double scale[2] = {1., 0.};
      for (int i = 0; i < N; ++i) {
        double w = weights[i];
        sumW += w;
        err += w * (weak_eval[i] != orig_response[i]);
      }
      if (sumW != 0) err /= sumW;
      C = err = -log_ratio(err);
      scale[1] = std::exp(err);
      sumW = 0;
      for (int i = 0; i < N; ++i) {
        double w = weights[i] * scale[weak_eval[i] != orig_response[i]];
        sumW += w;
        weights[i] = w;
      }
      for (auto& nd : tree.nodes) nd.value *= C;     // CvBoostTree::scale, o_cvboostree.cpp:50-69
    } else if (params.boost_type == BT_REAL) {
      for (int i = 0; i < N; ++i) weak_eval[i] *= -orig_response[i];
      for (int i = 0; i < N; ++i) weak_eval[i] = std::exp(weak_eval[i]);     // cvExp (<=1ulp vs libm)
      for (int i = 0; i < N; ++i) {
        double w = weights[i] * weak_eval[i];
        sumW += w;
        weights[i] = w;
      }
    } else if (params.boost_type == BT_LOGIT) {
      const double lbWeightThresh = FLT_EPSILON, lbZMax = 10.;
      for (int i = 0; i < N; ++i) {
        double s = sum_response[i] + 0.5 * weak_eval[i];
        sum_response[i] = s;
        weak_eval[i] = -2 * s;
      }
      for (int i = 0; i < N; ++i) weak_eval[i] = std::exp(weak_eval[i]);
      for (int i = 0; i < N; ++i) {
        double p = 1. / (1. + weak_eval[i]);
        double w = p * (1 - p), z;
        w = std::max(w, lbWeightThresh);
        weights[i] = w;
        sumW += w;
        if (orig_response[i] > 0) { z = 1. / p; responses[i] = (float)std::min(z, lbZMax); }
        else { z = 1. / (1 - p); responses[i] = (float)-std::min(z, lbZMax); }
      }
    } else {
      for (int i = 0; i < N; ++i) weak_eval[i] *= -orig_response[i];
      for (int i = 0; i < N; ++i) weak_eval[i] = std::exp(weak_eval[i]);
      for (int i = 0; i < N; ++i) {
        double w = weights[i] * weak_eval[i];
        weights[i] = w;
        sumW += w;
      }
    }
    if (sumW > FLT_EPSILON) {
      sumW = 1. / sumW;
      for (int i = 0; i < N; ++i) weights[i] *= sumW;
    }
  }

  // ---- CvBoost::trim_weights, o_cvboost.cpp:101-139 ----
  void trim_weights() {
    if (params.weight_trim_rate <= 0. || params.weight_trim_rate >= 1.) return;
    weak_eval = weights;                       // weak_eval doubles as the sort buffer
    std::sort(weak_eval.begin(), weak_eval.end());
    double sum = 1. - params.weight_trim_rate;
    int i;
    for (i = 0; i < N; ++i) {
      double w = weak_eval[i];
      if (sum <= 0) break;
      sum -= w;
    }
    double threshold = i < N ? weak_eval[i] : DBL_MAX;
    int nz = 0;
    for (i = 0; i < N; ++i) {
      int f = weights[i] >= threshold;
      mask[i] = (uint8_t)f;
      nz += f;
    }
    have_subsample = nz < N;
  }

  // ---- stage sum / decision, boost.cpp:461-477 ----
  double stage_sum(int si) const {
    double sum = 0;
    for (const Tree& t : stage.weak) sum += predict_node(t, si).value;
    return sum;
  }

  // ---- CvCascadeBoost::isErrDesired, boost.cpp:479-518 ----
  bool is_err_desired() {
    const Evaluator* ev = data->ev;
    int numPos = 0, numNeg = 0, numFalse = 0, numPosTrue = 0;
    std::vector<float> eval(N);
    for (int i = 0; i < N; ++i)
      if (ev->cls[i] == 1.0F) eval[numPos++] = (float)stage_sum(i);
    std::sort(eval.begin(), eval.begin() + numPos);
    int thresholdIdx = (int)((1.0F - params.minHitRate) * numPos);
    stage.threshold = eval[thresholdIdx];
    numPosTrue = numPos - thresholdIdx;
    for (int i = thresholdIdx - 1; i >= 0; --i)
      if (std::fabs(eval[i] - stage.threshold) < FLT_EPSILON) numPosTrue++;
    float hitRate = ((float)numPosTrue) / ((float)numPos);
    for (int i = 0; i < N; ++i)
      if (ev->cls[i] == 0.0F) {
        numNeg++;
        double s = stage_sum(i);
        if (!(s < stage.threshold - 0.00001F)) numFalse++;
      }
    float falseAlarm = ((float)numFalse) / ((float)numNeg);
    stage.log.push_back({(int)stage.weak.size(), hitRate, falseAlarm});
    return falseAlarm <= params.maxFalseAlarm;
  }

  // =============================== tree growing ===================================
  struct NodeWork { std::vector<int> members; };   // members[j] = sample index ( == cv_label, see boost.cpp:226-231 )

  // CvBoostTree::calc_node_value, o_cvboostree.cpp:657-732
  void calc_node_value(TreeNode& node, const std::vector<int>& members, int* cls_count) {
    const int n = (int)members.size();
    double rcw[2] = {0, 0};
    if (is_classifier) {
      cls_count[0] = cls_count[1] = 0;
      for (int i = 0; i < n; ++i) {
        int idx = members[i];
        double w = weights[idx];
        int r = (int)responses[idx];
        rcw[r] += w;
        cls_count[r]++;
        subtree_weights[i] = w;
      }
      int class_idx = rcw[1] > rcw[0];
      if (params.boost_type == BT_DISCRETE) node.value = class_idx * 2 - 1;
      else {
        double p = rcw[1] / (rcw[0] + rcw[1]);
        node.value = 0.5 * log_ratio(p);
      }
    } else {
      double sum = 0, sum2 = 0, iw;
      for (int i = 0; i < n; ++i) {
        int idx = members[i];
        double w = weights[idx];
        double t = responses[idx];
        rcw[0] += w;
        subtree_weights[i] = w;
        sum += t * w;
        sum2 += t * t * w;
      }
      iw = 1. / rcw[0];
      node.value = sum * iw;
      node.node_risk = sum2 - (sum * iw) * sum;
      node.node_risk *= n * iw * n * iw;
    }
    subtree_weights[n] = rcw[0];
    subtree_weights[n + 1] = rcw[1];
  }

  // node-local sorted order of ordered variable vi (values in sorted order, node-local positions)
  void node_sorted(int vi, const std::vector<int>& members, std::vector<int>& sorted, std::vector<float>& values) {
    const float* row = &data->fval[(size_t)vi * N];
    sorted.clear(); values.clear();
    for (int i = 0; i < N; ++i) {
      const uint32_t gi = data->sorted_at(vi, i);
      int p = pos_in_node[gi];
      if (p >= 0) { sorted.push_back(p); values.push_back(row[gi]); }
    }
    (void)members;
  }

  struct SplitCand { bool valid = false; int var_idx = -1; float c = 0; int split_point = -1; int subset[8] = {0,0,0,0,0,0,0,0}; float quality = 0; };

  // CvBoostTree::find_split_ord_reg, o_cvboostree.cpp:361-426
  bool find_split_ord_reg(const TreeNode& node, int vi, float init_quality, const std::vector<int>& members,
                          const std::vector<int>& sorted, const std::vector<float>& values, SplitCand& out) {
    const float epsilon = FLT_EPSILON * 2;
    const int n = (int)members.size();
    int best_i = -1;
    double L = 0, R = subtree_weights[n];
    double best_val = init_quality, lsum = 0, rsum = node.value * R;
    for (int i = 0; i < n - 1; ++i) {
      int idx = sorted[i];
      double w = subtree_weights[idx];
      double t = responses[members[idx]] * w;
      L += w; R -= w;
      lsum += t; rsum -= t;
      if (values[i] + epsilon < values[i + 1]) {
        double val = (lsum * lsum * R + rsum * rsum * L) / (L * R);
        if (best_val < val) { best_val = val; best_i = i; }
      }
    }
    if (best_i < 0) return false;
    out.valid = true; out.var_idx = vi;
    out.c = (values[best_i] + values[best_i + 1]) * 0.5f;
    out.split_point = best_i;
    out.quality = (float)best_val;
    return true;
  }

  // CvBoostTree::find_split_ord_class, o_cvboostree.cpp:151-247
  bool find_split_ord_class(int vi, float init_quality, const std::vector<int>& members,
                            const std::vector<int>& sorted, const std::vector<float>& values, SplitCand& out) {
    const float epsilon = FLT_EPSILON * 2;
    const int n = (int)members.size();
    double lcw[2] = {0, 0}, rcw[2];
    int best_i = -1;
    double best_val = init_quality;
    rcw[0] = subtree_weights[n]; rcw[1] = subtree_weights[n + 1];
    if (split_criteria == 0) {
      double L = 0, R = rcw[0] + rcw[1];
      double lsum2 = 0, rsum2 = rcw[0] * rcw[0] + rcw[1] * rcw[1];
      for (int i = 0; i < n - 1; ++i) {
        int idx = sorted[i];
        double w = subtree_weights[idx], w2 = w * w;
        int k = (int)responses[members[idx]];
        L += w; R -= w;
        double lv = lcw[k], rv = rcw[k];
        lsum2 += 2 * lv * w + w2;
        rsum2 -= 2 * rv * w - w2;
        lcw[k] = lv + w; rcw[k] = rv - w;
        if (values[i] + epsilon < values[i + 1]) {
          double val = (lsum2 * R + rsum2 * L) / (L * R);
          if (best_val < val) { best_val = val; best_i = i; }
        }
      }
    } else {
      for (int i = 0; i < n - 1; ++i) {
        int idx = sorted[i];
        double w = subtree_weights[idx];
        int k = (int)responses[members[idx]];
        lcw[k] += w; rcw[k] -= w;
        if (values[i] + epsilon < values[i + 1]) {
          double val = lcw[0] + rcw[1], val2 = lcw[1] + rcw[0];
          val = std::max(val, val2);
          if (best_val < val) { best_val = val; best_i = i; }
        }
      }
    }
    if (best_i < 0) return false;
    out.valid = true; out.var_idx = vi;
    out.c = (values[best_i] + values[best_i + 1]) * 0.5f;
    out.split_point = best_i;
    out.quality = (float)best_val;
    return true;
  }

  // sort category ids by key; SORT_STD mirrors std::sort(ptrs, LessThanPtr) of the reference
  void sort_cats(std::vector<int>& order, const double* key, int mi) const {
    order.resize(mi);
    for (int i = 0; i < mi; ++i) order[i] = i;
    if (catSortMode == SORT_STD) std::sort(order.begin(), order.end(), [key](int a, int b) { return key[a] < key[b]; });
    else std::stable_sort(order.begin(), order.end(), [key](int a, int b) { return key[a] < key[b]; });
  }

  // CvBoostTree::find_split_cat_reg, o_cvboostree.cpp:428-516
  bool find_split_cat_reg(int vi, float init_quality, const std::vector<int>& members, SplitCand& out) {
    const int n = (int)members.size(), mi = data->ev->maxCatCount;
    std::vector<double> sum(mi, 0.), counts(mi, 0.);
    const float* row = &data->fval[(size_t)vi * N];
    double L = 0, R = 0, best_val = init_quality, lsum = 0, rsum = 0;
    int best_subset = -1;
    for (int i = 0; i < n; ++i) {
      int idx = (int)row[members[i]];
      double w = subtree_weights[i];
      double s = sum[idx] + responses[members[i]] * w;
      double nc = counts[idx] + w;
      sum[idx] = s; counts[idx] = nc;
    }
    for (int i = 0; i < mi; ++i) {
      R += counts[i];
      rsum += sum[i];
      sum[i] = std::fabs(counts[i]) > DBL_EPSILON ? sum[i] / counts[i] : 0;
    }
    std::vector<int> order;
    sort_cats(order, sum.data(), mi);
    for (int i = 0; i < mi; ++i) sum[i] *= counts[i];
    for (int subset_i = 0; subset_i < mi - 1; ++subset_i) {
      int idx = order[subset_i];
      double ni = counts[idx];
      if (ni > FLT_EPSILON) {
        double s = sum[idx];
        lsum += s; L += ni; rsum -= s; R -= ni;
        if (L > FLT_EPSILON && R > FLT_EPSILON) {
          double val = (lsum * lsum * R + rsum * rsum * L) / (L * R);
          if (best_val < val) { best_val = val; best_subset = subset_i; }
        }
      }
    }
    if (best_subset < 0) return false;
    out.valid = true; out.var_idx = vi; out.quality = (float)best_val;
    std::memset(out.subset, 0, sizeof(out.subset));
    for (int i = 0; i <= best_subset; ++i) { int idx = order[i]; out.subset[idx >> 5] |= 1 << (idx & 31); }
    return true;
  }

  // CvBoostTree::find_split_cat_class, o_cvboostree.cpp:249-359
  bool find_split_cat_class(int vi, float init_quality, const std::vector<int>& members, SplitCand& out) {
    const int n = (int)members.size(), mi = data->ev->maxCatCount;
    std::vector<double> cjk((size_t)mi * 2, 0.), key(mi);
    const float* row = &data->fval[(size_t)vi * N];
    double lcw[2] = {0, 0}, rcw[2] = {0, 0};
    double L = 0, R, best_val = init_quality;
    int best_subset = -1;
    for (int i = 0; i < n; ++i) {
      double w = subtree_weights[i];
      int j = (int)row[members[i]];
      int k = (int)responses[members[i]];
      cjk[j * 2 + k] += w;
    }
    for (int j = 0; j < mi; ++j) { rcw[0] += cjk[j * 2]; rcw[1] += cjk[j * 2 + 1]; key[j] = cjk[j * 2 + 1]; }
    R = rcw[0] + rcw[1];
    std::vector<int> order;
    sort_cats(order, key.data(), mi);
    for (int subset_i = 0; subset_i < mi - 1; ++subset_i) {
      int idx = order[subset_i];
      double w0 = cjk[idx * 2], w1 = cjk[idx * 2 + 1];
      double weight = w0 + w1;
      if (weight < FLT_EPSILON) continue;
      lcw[0] += w0; rcw[0] -= w0; lcw[1] += w1; rcw[1] -= w1;
      if (split_criteria == 0) {
        double lsum2 = lcw[0] * lcw[0] + lcw[1] * lcw[1];
        double rsum2 = rcw[0] * rcw[0] + rcw[1] * rcw[1];
        L += weight; R -= weight;
        if (L > FLT_EPSILON && R > FLT_EPSILON) {
          double val = (lsum2 * R + rsum2 * L) / (L * R);
          if (best_val < val) { best_val = val; best_subset = subset_i; }
        }
      } else {
        double val = lcw[0] + rcw[1], val2 = lcw[1] + rcw[0];
        val = std::max(val, val2);
        if (best_val < val) { best_val = val; best_subset = subset_i; }
      }
    }
    if (best_subset < 0) return false;
    out.valid = true; out.var_idx = vi; out.quality = (float)best_val;
    std::memset(out.subset, 0, sizeof(out.subset));
    for (int i = 0; i <= best_subset; ++i) { int idx = order[i]; out.subset[idx >> 5] |= 1 << (idx & 31); }
    return true;
  }

  // CvDTree::find_best_split + DTreeBestSplitFinder, o_cvdtree.cpp:296-357 (serial, running best as init_quality)
  SplitCand find_best_split(const TreeNode& node, const std::vector<int>& members) {
    SplitCand best; best.quality = -1;
    SplitCand cur;
    std::vector<int> sorted; std::vector<float> values;
    const int n = (int)members.size();
    for (int vi = 0; vi < data->F; ++vi) {
      if (n <= 1) continue;
      bool res;
      cur = SplitCand();
      if (data->categorical) {
        res = is_classifier ? find_split_cat_class(vi, best.quality, members, cur)
                            : find_split_cat_reg(vi, best.quality, members, cur);
      } else {
        node_sorted(vi, members, sorted, values);
        res = is_classifier ? find_split_ord_class(vi, best.quality, members, sorted, values, cur)
                            : find_split_ord_reg(node, vi, best.quality, members, sorted, values, cur);
      }
      if (res && best.quality < cur.quality) best = cur;
    }
    if (!(best.quality > 0)) best.valid = false;
    return best;
  }

  // CvBoostTree::try_split_node + CvDTree::try_split_node, o_cvboostree.cpp:71-85, o_cvdtree.cpp:122-187
  void try_split_node(Tree& tree, int ni, std::vector<int> members) {
    const int n = (int)members.size();
    int cls_count[2] = {0, 0};
    bool can_split = true;
    tree.nodes[ni].sample_count = n;
    calc_node_value(tree.nodes[ni], members, cls_count);
    if (n <= params.min_sample_count || tree.nodes[ni].depth >= params.max_depth) can_split = false;
    if (can_split && is_classifier) {
      int nz = (cls_count[0] != 0) + (cls_count[1] != 0);
      if (nz == 1) can_split = false;
    } else if (can_split) {
      if (std::sqrt(tree.nodes[ni].node_risk) / n < params.regression_accuracy) can_split = false;
    }
    SplitCand best;
    if (can_split) {
      std::fill(pos_in_node.begin(), pos_in_node.end(), -1);
      for (int j = 0; j < n; ++j) pos_in_node[members[j]] = j;
      best = find_best_split(tree.nodes[ni], members);
    }
    if (!can_split || !best.valid) {
      const double value = tree.nodes[ni].value;           // leaf: o_cvboostree.cpp:74-84
      for (int i = 0; i < n; ++i) weak_eval[members[i]] = value;
      return;
    }
    TreeNode& nd = tree.nodes[ni];
    nd.var_idx = best.var_idx; nd.c = best.c; nd.split_point = best.split_point; nd.quality = best.quality;
    std::memcpy(nd.subset, best.subset, sizeof(nd.subset));
    // calc_node_dir (o_cvboostree.cpp:87-149) + complete_node_dir (o_cvdtree.cpp:1247-1320)
    dir.assign(n, 0);
    if (data->categorical) {
      const float* row = &data->fval[(size_t)best.var_idx * N];
      for (int i = 0; i < n; ++i) dir[i] = (int8_t)(cat_dir((int)row[members[i]], best.subset) > 0);
    } else {
      std::vector<int> sorted; std::vector<float> values;
      node_sorted(best.var_idx, members, sorted, values);
      for (int i = 0; i < n; ++i) dir[sorted[i]] = (int8_t)(i <= best.split_point ? 0 : 1);
    }
    // split_node_data, o_cvcascadeboosttree.cpp:167-347 : stable partition of the member list
    std::vector<int> lm, rm;
    for (int i = 0; i < n; ++i) (dir[i] ? rm : lm).push_back(members[i]);
    members.clear(); members.shrink_to_fit();
    TreeNode l, r;
    l.parent = r.parent = ni; l.depth = r.depth = tree.nodes[ni].depth + 1;
    int li = (int)tree.nodes.size(); tree.nodes.push_back(l);
    int ri = (int)tree.nodes.size(); tree.nodes.push_back(r);
    tree.nodes[ni].left = li; tree.nodes[ni].right = ri;
    try_split_node(tree, li, std::move(lm));
    try_split_node(tree, ri, std::move(rm));
  }

  // CvBoostTree::train -> CvDTree::do_train, o_cvboostree.cpp:28-35, o_cvdtree.cpp:95-120
  bool train_tree(Tree& tree) {
    std::vector<int> members;     // subsample_data: active samples in ascending order
    for (int i = 0; i < N; ++i) if (mask[i]) members.push_back(i);
    tree.nodes.clear();
    tree.nodes.emplace_back();
    try_split_node(tree, 0, std::move(members));
    return tree.nodes[0].left >= 0;
  }

  // CvCascadeBoost::train, boost.cpp:409-459
  bool train(const TrainData* d, const BoostParams& p, int catSort) {
    data = d; N = d->N; catSortMode = catSort;
    set_params(p);
    stage = Stage();
    pos_in_node.assign(N, -1);
    init_weights();
    do {
      Tree tree;
      if (!train_tree(tree)) break;
      stage.weak.push_back(tree);
      update_weights(stage.weak.back());
      trim_weights();
      if (on_weak) on_weak((int)stage.weak.size() - 1);
      int nz = 0; for (int i = 0; i < N; ++i) nz += mask[i] != 0;
      if (nz == 0) break;
    } while (!is_err_desired() && (int)stage.weak.size() < params.weak_count);
    return !stage.weak.empty();
  }
};

}  // namespace orc

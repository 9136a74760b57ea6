// ORACLE — TEST INFRASTRUCTURE ONLY.  C ABI over the CPU restatement so that tests/ and
// bench.py's cpu_baseline leg can drive it through ctypes.  Never linked into the product.
#include <chrono>

#include "orc_mining.hpp"

using namespace orc;

extern "C" {

struct orc_node {
  int32_t left, right, var_idx, split_point;
  float c, quality;
  int32_t subset[8];
  double value;
  int32_t depth, sample_count;
};

struct orc_boost_params {
  int32_t boost_type, weak_count, max_depth, min_sample_count;
  double weight_trim_rate;
  float regression_accuracy, min_hit_rate, max_false_alarm;
};

struct orc_stage_report {
  int32_t pos_count, neg_count, trained, pad;
  int64_t pos_consumed, neg_consumed;
  double acceptance_ratio;
};

static BoostParams to_params(const orc_boost_params* p) {
  BoostParams b;
  b.boost_type = p->boost_type; b.weak_count = p->weak_count; b.max_depth = p->max_depth;
  b.min_sample_count = p->min_sample_count; b.weight_trim_rate = p->weight_trim_rate;
  b.regression_accuracy = p->regression_accuracy; b.minHitRate = p->min_hit_rate; b.maxFalseAlarm = p->max_false_alarm;
  return b;
}
static void fill_node(const TreeNode& n, orc_node* o) {
  o->left = n.left; o->right = n.right; o->var_idx = n.var_idx; o->split_point = n.split_point;
  o->c = n.c; o->quality = n.quality; std::memcpy(o->subset, n.subset, sizeof(o->subset));
  o->value = n.value; o->depth = n.depth; o->sample_count = n.sample_count;
}

// ---- stateless kernels -------------------------------------------------------------------
void orc_integral(const uint8_t* img, int W, int H, int32_t* sum, double* sq, int32_t* tilted) {
  integral_u8(img, W, H, sum, sq, tilted);
}
float orc_norm_factor(const uint8_t* img, int W, int H) {
  std::vector<int32_t> s((size_t)(W + 1) * (H + 1));
  std::vector<double> q(s.size());
  integral_u8(img, W, H, s.data(), q.data(), nullptr);
  return norm_factor(s.data(), q.data(), W, H);
}
void orc_resize_linear_exact(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  resize_linear_exact(src, sw, sh, dst, dw, dh);
}
void orc_hog_planes(const uint8_t* img, int W, int H, float* planes10) {   // 9 hist planes then norm
  const size_t P = (size_t)(W + 1) * (H + 1);
  float* hp[HOG_BINS];
  for (int b = 0; b < HOG_BINS; ++b) hp[b] = planes10 + b * P;
  hog_integral_histogram(img, W, H, hp, planes10 + HOG_BINS * P);
}
float orc_fast_atan2(float y, float x) { return cv_fast_atan2_rad(y, x); }

int orc_catalog_count(int type, int mode, int W, int H) {
  if (type == FT_HAAR) return (int)haar_catalog(W, H, mode).size();
  if (type == FT_LBP) return (int)lbp_catalog(W, H).size();
  return (int)hog_catalog(W, H).size();
}
// rects: F x 3 x 4 int32 ; weights: F x 3 float ; tilted: F int32
int orc_haar_catalog(int W, int H, int mode, int32_t* rects, float* weights, int32_t* tilted) {
  auto c = haar_catalog(W, H, mode);
  for (size_t i = 0; i < c.size(); ++i) {
    for (int j = 0; j < 3; ++j) {
      rects[(i * 3 + j) * 4 + 0] = c[i].r[j].x; rects[(i * 3 + j) * 4 + 1] = c[i].r[j].y;
      rects[(i * 3 + j) * 4 + 2] = c[i].r[j].w; rects[(i * 3 + j) * 4 + 3] = c[i].r[j].h;
      weights[i * 3 + j] = c[i].wt[j];
    }
    tilted[i] = c[i].tilted;
  }
  return (int)c.size();
}
int orc_lbp_catalog(int W, int H, int32_t* rects) {
  auto c = lbp_catalog(W, H);
  for (size_t i = 0; i < c.size(); ++i) { rects[i * 4] = c[i].r.x; rects[i * 4 + 1] = c[i].r.y; rects[i * 4 + 2] = c[i].r.w; rects[i * 4 + 3] = c[i].r.h; }
  return (int)c.size();
}
int orc_hog_catalog(int W, int H, int32_t* rects) {   // cell 0 of every block
  auto c = hog_catalog(W, H);
  for (size_t i = 0; i < c.size(); ++i) { rects[i * 4] = c[i].cell[0].x; rects[i * 4 + 1] = c[i].cell[0].y; rects[i * 4 + 2] = c[i].cell[0].w; rects[i * 4 + 3] = c[i].cell[0].h; }
  return (int)c.size();
}
// Feature::calc on an arbitrary flattened integral row (reference known-answer tests)
float orc_haar_calc(int stride, int tilted, const int32_t* rects12, const float* w3, const int32_t* integral) {
  Rect4 r[3];
  for (int j = 0; j < 3; ++j) r[j] = {rects12[j * 4], rects12[j * 4 + 1], rects12[j * 4 + 2], rects12[j * 4 + 3]};
  HaarFeature f = make_haar(stride, tilted != 0, r[0], w3[0], r[1], w3[1], r[2], w3[2]);
  return haar_calc(f, integral);
}

// ---- evaluator ---------------------------------------------------------------------------
void* orc_eval_create(int type, int mode, int W, int H, int maxSamples) {
  auto* e = new Evaluator();
  e->init(type, mode, W, H, maxSamples);
  return e;
}
void orc_eval_destroy(void* h) { delete (Evaluator*)h; }
int orc_eval_num_features(void* h) { return ((Evaluator*)h)->numFeatures; }
int orc_eval_var_count(void* h) { return ((Evaluator*)h)->varCount(); }
int orc_eval_feat_size(void* h) { return ((Evaluator*)h)->featSize; }
int orc_eval_max_cat(void* h) { return ((Evaluator*)h)->maxCatCount; }
void orc_eval_set_images(void* h, const uint8_t* imgs, const uint8_t* labels, int first, int n, int threads) {
  auto* e = (Evaluator*)h;
  const size_t px = (size_t)e->W * e->H;
  parallel_rows(0, n, threads, [&](int b, int en) {
    for (int i = b; i < en; ++i) e->setImage(imgs + (size_t)i * px, labels[i], first + i);
  });
}
float orc_eval_value(void* h, int vi, int si) { return (*(Evaluator*)h)(vi, si); }
void orc_eval_values(void* h, const int32_t* vis, int nv, int s0, int ns, float* out) {
  auto* e = (Evaluator*)h;
  for (int a = 0; a < nv; ++a)
    for (int s = 0; s < ns; ++s) out[(size_t)a * ns + s] = (*e)(vis[a], s0 + s);
}
void orc_eval_get_integrals(void* h, int si, int32_t* sum, int32_t* tilted, float* nf) {
  auto* e = (Evaluator*)h;
  if (sum && !e->sum.empty()) std::memcpy(sum, &e->sum[(size_t)si * e->P], sizeof(int32_t) * e->P);
  if (tilted && !e->tilted.empty()) std::memcpy(tilted, &e->tilted[(size_t)si * e->P], sizeof(int32_t) * e->P);
  if (nf && !e->nf.empty()) *nf = e->nf[si];
}
void orc_eval_get_hog(void* h, int si, float* planes10) {
  auto* e = (Evaluator*)h;
  for (int b = 0; b < HOG_BINS; ++b) std::memcpy(planes10 + (size_t)b * e->P, &e->hist[b][(size_t)si * e->P], sizeof(float) * e->P);
  std::memcpy(planes10 + (size_t)HOG_BINS * e->P, &e->normSum[(size_t)si * e->P], sizeof(float) * e->P);
}

// ---- train data (precalc + sort) ---------------------------------------------------------
void* orc_td_create(void* eval, int N, int sortMode, int threads, const int32_t* subset, int nsubset) {
  auto* td = new TrainData();
  td->precalculate((Evaluator*)eval, N, sortMode, threads, subset, nsubset);
  return td;
}
void orc_td_destroy(void* h) { delete (TrainData*)h; }
int orc_td_rows(void* h) { return ((TrainData*)h)->F; }
const float* orc_td_fval(void* h) { return ((TrainData*)h)->fval.data(); }
const uint16_t* orc_td_sidx(void* h) { auto* t = (TrainData*)h; return t->sidx.empty() ? nullptr : t->sidx.data(); }
const uint32_t* orc_td_sidx32(void* h) { auto* t = (TrainData*)h; return t->sidx32.empty() ? nullptr : t->sidx32.data(); }

// ---- stage training ----------------------------------------------------------------------
struct StageResult {
  Stage stage;
  bool ok = false;
  std::vector<std::vector<double>> weights_after;   // per weak tree (optional)
  std::vector<std::vector<uint8_t>> mask_after;
  double seconds = 0;
};
void* orc_boost_train(void* td, const orc_boost_params* p, int catSortMode, int record) {
  auto* r = new StageResult();
  BoostTrainer bt;
  if (record)
    bt.on_weak = [&](int) { r->weights_after.push_back(bt.weights); r->mask_after.push_back(bt.mask); };
  auto t0 = std::chrono::steady_clock::now();
  r->ok = bt.train((TrainData*)td, to_params(p), catSortMode);
  r->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  r->stage = bt.stage;
  return r;
}
void orc_stage_destroy(void* h) { delete (StageResult*)h; }
int orc_stage_ok(void* h) { return ((StageResult*)h)->ok; }
double orc_stage_seconds(void* h) { return ((StageResult*)h)->seconds; }
int orc_stage_num_weak(void* h) { return (int)((StageResult*)h)->stage.weak.size(); }
float orc_stage_threshold(void* h) { return ((StageResult*)h)->stage.threshold; }
int orc_stage_tree_num_nodes(void* h, int t) { return (int)((StageResult*)h)->stage.weak[t].nodes.size(); }
void orc_stage_tree_nodes(void* h, int t, orc_node* out) {
  const Tree& tr = ((StageResult*)h)->stage.weak[t];
  for (size_t i = 0; i < tr.nodes.size(); ++i) fill_node(tr.nodes[i], out + i);
}
int orc_stage_log(void* h, int32_t* n, float* hr, float* fa, int cap) {
  const auto& lg = ((StageResult*)h)->stage.log;
  for (size_t i = 0; i < lg.size() && (int)i < cap; ++i) { n[i] = lg[i].n; hr[i] = lg[i].hr; fa[i] = lg[i].fa; }
  return (int)lg.size();
}
int orc_stage_snapshot(void* h, int t, double* weights, uint8_t* mask) {
  auto* r = (StageResult*)h;
  if (t >= (int)r->weights_after.size()) return 0;
  std::memcpy(weights, r->weights_after[t].data(), sizeof(double) * r->weights_after[t].size());
  std::memcpy(mask, r->mask_after[t].data(), r->mask_after[t].size());
  return 1;
}

// ---- negative reader ---------------------------------------------------------------------
struct NegHandle { std::vector<GrayImage> images; NegReader rd; bool created = false; int W, H; };
void* orc_neg_create(int W, int H) { auto* n = new NegHandle(); n->W = W; n->H = H; return n; }
void orc_neg_destroy(void* h) { delete (NegHandle*)h; }
void orc_neg_add_image(void* h, const uint8_t* px, int w, int hh) {
  GrayImage g; g.w = w; g.h = hh; g.px.assign(px, px + (size_t)w * hh);
  ((NegHandle*)h)->images.push_back(std::move(g));
}
int orc_neg_get(void* h, uint8_t* out) {
  auto* n = (NegHandle*)h;
  if (!n->created) { n->rd.create(&n->images, n->W, n->H); n->created = true; }
  return n->rd.get(out) ? 1 : 0;
}
void orc_neg_state(void* h, int64_t* st /*last, round, pointX, pointY, offX, offY, imgW, imgH*/, float* scale) {
  auto* n = (NegHandle*)h;
  st[0] = (int64_t)n->rd.last; st[1] = (int64_t)n->rd.round; st[2] = n->rd.pointX; st[3] = n->rd.pointY;
  st[4] = n->rd.offX; st[5] = n->rd.offY; st[6] = n->rd.img.w; st[7] = n->rd.img.h;
  *scale = n->rd.scale;
}

// ---- cascade driver ----------------------------------------------------------------------
struct CascadeHandle {
  CascadeTrainer tr;
  std::vector<uint8_t> posSamples;
  std::vector<GrayImage> bgs;
  bool readersReady = false;
  void ready() {
    if (readersReady) return;
    tr.pos.create(posSamples.data(), (int)(posSamples.size() / ((size_t)tr.ev.W * tr.ev.H)), tr.ev.W * tr.ev.H);
    tr.neg.create(&bgs, tr.ev.W, tr.ev.H);
    readersReady = true;
  }
};
void* orc_cascade_create(int type, int mode, int W, int H, int numPos, int numNeg, int numStages,
                         const orc_boost_params* p, int sortMode, int threads, double breakValue) {
  auto* c = new CascadeHandle();
  c->tr.ev.init(type, mode, W, H, numPos + numNeg);
  c->tr.numPos = numPos; c->tr.numNeg = numNeg; c->tr.numStages = numStages;
  c->tr.bp = to_params(p); c->tr.sortMode = sortMode; c->tr.threads = threads;
  c->tr.acceptanceRatioBreakValue = breakValue;
  return c;
}
void orc_cascade_destroy(void* h) { delete (CascadeHandle*)h; }
void orc_cascade_set_pos(void* h, const uint8_t* samples, int count) {
  auto* c = (CascadeHandle*)h;
  c->posSamples.assign(samples, samples + (size_t)count * c->tr.ev.W * c->tr.ev.H);
}
void orc_cascade_add_bg(void* h, const uint8_t* px, int w, int hh) {
  GrayImage g; g.w = w; g.h = hh; g.px.assign(px, px + (size_t)w * hh);
  ((CascadeHandle*)h)->bgs.push_back(std::move(g));
}
void orc_cascade_train(void* h) { auto* c = (CascadeHandle*)h; c->ready(); c->tr.train(); }
int orc_cascade_termination(void* h) { return ((CascadeHandle*)h)->tr.termination; }
int orc_cascade_num_stages(void* h) { return (int)((CascadeHandle*)h)->tr.stages.size(); }
int orc_cascade_num_reports(void* h) { return (int)((CascadeHandle*)h)->tr.reports.size(); }
void orc_cascade_report(void* h, int i, orc_stage_report* out) {
  const StageReport& r = ((CascadeHandle*)h)->tr.reports[i];
  out->pos_count = r.posCount; out->neg_count = r.negCount; out->trained = r.trained; out->pad = 0;
  out->pos_consumed = r.posConsumed; out->neg_consumed = r.negConsumed; out->acceptance_ratio = r.acceptanceRatio;
}
float orc_cascade_stage_threshold(void* h, int s) { return ((CascadeHandle*)h)->tr.stages[s].threshold; }
int orc_cascade_stage_num_weak(void* h, int s) { return (int)((CascadeHandle*)h)->tr.stages[s].weak.size(); }
int orc_cascade_tree_num_nodes(void* h, int s, int t) { return (int)((CascadeHandle*)h)->tr.stages[s].weak[t].nodes.size(); }
void orc_cascade_tree_nodes(void* h, int s, int t, orc_node* out) {
  const Tree& tr = ((CascadeHandle*)h)->tr.stages[s].weak[t];
  for (size_t i = 0; i < tr.nodes.size(); ++i) fill_node(tr.nodes[i], out + i);
}
int orc_cascade_stage_log(void* h, int s, int32_t* n, float* hr, float* fa, int cap) {
  const auto& lg = ((CascadeHandle*)h)->tr.stageDetails[s].log;
  for (size_t i = 0; i < lg.size() && (int)i < cap; ++i) { n[i] = lg[i].n; hr[i] = lg[i].hr; fa[i] = lg[i].fa; }
  return (int)lg.size();
}
// inject an externally trained stage (e.g. from the CUDA path) so that mining can be replayed on the CPU
void orc_cascade_push_stage(void* h, float threshold, int ntrees, const int32_t* nodes_per_tree, const orc_node* nodes) {
  CascadeStage cs; cs.threshold = threshold;
  const orc_node* p = nodes;
  for (int t = 0; t < ntrees; ++t) {
    Tree tr;
    for (int i = 0; i < nodes_per_tree[t]; ++i, ++p) {
      TreeNode n; n.left = p->left; n.right = p->right; n.var_idx = p->var_idx; n.split_point = p->split_point;
      n.c = p->c; n.quality = p->quality; std::memcpy(n.subset, p->subset, sizeof(n.subset)); n.value = p->value;
      n.depth = p->depth; n.sample_count = p->sample_count;
      tr.nodes.push_back(n);
    }
    cs.weak.push_back(tr);
  }
  ((CascadeHandle*)h)->tr.stages.push_back(cs);
}
// one mining pass exactly as updateTrainingSet does it; returns 1 on success
int orc_cascade_update_training_set(void* h, double minRatio, orc_stage_report* out) {
  auto* c = (CascadeHandle*)h; c->ready();
  StageReport rep; double ar = 0;
  bool ok = c->tr.updateTrainingSet(minRatio, ar, rep);
  out->pos_count = rep.posCount; out->neg_count = rep.negCount; out->trained = 0; out->pad = 0;
  out->pos_consumed = rep.posConsumed; out->neg_consumed = rep.negConsumed; out->acceptance_ratio = rep.acceptanceRatio;
  return ok ? 1 : 0;
}
int64_t orc_cascade_fill(void* h, int first, int count, int isPositive, double minRatio, int* got) {
  auto* c = (CascadeHandle*)h; c->ready();
  long long consumed = 0;
  *got = c->tr.fillPassedSamples(first, count, isPositive != 0, minRatio, consumed);
  return consumed;
}
void* orc_cascade_evaluator(void* h) { return &((CascadeHandle*)h)->tr.ev; }
int orc_cascade_cur_num_samples(void* h) { return ((CascadeHandle*)h)->tr.curNumSamples; }
void orc_cascade_neg_state(void* h, int64_t* st, float* scale) {
  auto& rd = ((CascadeHandle*)h)->tr.neg;
  st[0] = (int64_t)rd.last; st[1] = (int64_t)rd.round; st[2] = rd.pointX; st[3] = rd.pointY;
  st[4] = rd.offX; st[5] = rd.offY; st[6] = rd.img.w; st[7] = rd.img.h;
  *scale = rd.scale;
}

}  // extern "C"

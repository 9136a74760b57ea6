// K4: the boosted-tree stage trainer.  Device restatement of CvCascadeBoost::train and everything
// under it (src/boost.cpp:168-518, src/o_cvboost.cpp:67-139, src/o_cvboostree.cpp:71-516,657-732,
// src/o_cvdtree.cpp:95-187,296-357,1247-1320, src/o_cvcascadeboosttree.cpp:16-39,167-347).
//
// State lives in HBM for the whole stage; per weak tree the host only reads a few small structs
// (node statistics, the winning split, the stage summary).  The split search — one pass over the
// sorted-value / sorted-index rows of every local variable per tree node — is the dominant kernel.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

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
EvalView make_eval_view(const cc_ctx* ctx);
int traindata_free(cc_ctx* ctx);

// ---- device copies of the evaluator (kept in sync with cc_eval.cu) --------------------------------
__device__ __forceinline__ float d_haar_value(const HaarTaps& t, const EvalView& e, int s) {
  const int32_t* base = (t.tilted() ? e.tilted : e.sum) + (size_t)(s >> 5) * e.P * TILE + (s & 31);
  int acc = 0;
  const int nt = t.ntaps();
  for (int k = 0; k < nt; ++k) acc += tap_coef(t.w[k]) * base[(size_t)tap_off(t.w[k]) * TILE];
  const float n = e.nf[s];
  return n == 0.f ? 0.f : __fdiv_rn((float)acc, n);
}
__device__ __forceinline__ int d_lbp_code(const LbpTaps& t, const EvalView& e, int s) {
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
__device__ __forceinline__ float d_hog_value(const HogTaps& t, const EvalView& e, int comp, int s) {
  const int bin = comp % 9, cell = comp / 9;
  const size_t np = (size_t)e.ns_pad;
  const float* ph = e.hog + (size_t)bin * e.P * np + s;
  const float* pn = e.hog + (size_t)9 * e.P * np + s;
  float res = __fadd_rn(__fsub_rn(__fsub_rn(ph[t.off[cell][0] * np], ph[t.off[cell][1] * np]), ph[t.off[cell][2] * np]), ph[t.off[cell][3] * np]);
  float nrm = __fadd_rn(__fsub_rn(__fsub_rn(pn[t.off[0][0] * np], pn[t.off[1][1] * np]), pn[t.off[2][2] * np]), pn[t.off[3][3] * np]);
  return res > 0.001f ? __fdiv_rn(res, __fadd_rn(nrm, 0.001f)) : 0.f;
}
__device__ __forceinline__ float d_eval_var(const EvalView& e, int var, int s) {
  if (e.ftype == CC_FEATURE_HAAR) return d_haar_value(e.haar_taps[var], e, s);
  if (e.ftype == CC_FEATURE_LBP) return (float)d_lbp_code(e.lbp_taps[var], e, s);
  return d_hog_value(e.hog_taps[var / 36], e, var % 36, s);
}

// ---- small structs shared with the host ------------------------------------------------------------
enum Crit { CRIT_GINI = 0, CRIT_MISCLASS = 1, CRIT_SQERR = 2 };

struct NodeStats {
  int n;
  int cls_count[2];
  int pad;
  double rcw[2];      // regression: rcw[0] = sum of weights
  double value, risk;
  double rsum0;       // regression: node value * rcw[0] (initial right sum of find_split_ord_reg)
};

struct RowBest { float q; float c; float vleft; int pos; };
struct RowBestCat { float q; int subset[8]; };

struct BestSplit {
  float q;
  int var;            // global variable index
  float c, vleft;
  int pos;
  int subset[8];
  int valid;
  int pad[2];
};

struct TreeSummary {
  double scaleC;      // DAB: tree->scale(C)
  double sumW;
  int nz_count, have_subsample;
  float threshold, hit_rate, false_alarm;
  int num_pos, num_neg;
};

// control block of one pruned search: packed argmax keys (quality bits << 32 | ~row) and the flagged-row counter
struct SearchCtl { unsigned long long key; unsigned int nflag; unsigned int pad; };

struct BoostBuffers {
  int cap_n = 0;
  size_t cap_rows = 0;
  double *weights = nullptr, *weak_eval = nullptr, *stage_sum = nullptr, *sum_response = nullptr, *wnode = nullptr, *treeval = nullptr;
  float *responses = nullptr, *ynode = nullptr;
  int* orig_response = nullptr;
  uint8_t* mask = nullptr;
  int* node_of = nullptr;
  RowBest* rowbest = nullptr;
  RowBestCat* rowbest_cat = nullptr;
  NodeStats* d_stats = nullptr;       // [CC_MAX_TREE_NODES]
  BestSplit* d_best = nullptr;        // [1 + CC_MAX_RANKS]: the local / winning candidate, then the all-gathered candidates of the ranks
  TreeSummary* d_summary = nullptr;
  DevNode* d_tree = nullptr;          // [CC_MAX_TREE_NODES]
  NodeStats* h_stats = nullptr;       // pinned
  BestSplit* h_best = nullptr;
  TreeSummary* h_summary = nullptr;
  int* d_tree_ok = nullptr;           // stump path
  int* h_tree_ok = nullptr;
  // stage state
  int boost_type = CC_BOOST_GENTLE, crit = CRIT_SQERR, is_classifier = 0, have_subsample = 0;
  cc_boost_params params;
  bool row_quality_valid = false;
  // pruned ordered split search (k_ss_bound): control block, seed / flagged row lists, statistics
  struct SearchCtl* d_ctl = nullptr;
  int* d_seed_rows = nullptr;
  int* d_flag_rows = nullptr;
  std::vector<int> seed_rows;         // strided sample of the local rows + the winners of this stage so far
  int seed_cap = 0, n_seed_static = 0;
  int last_nflag = -1;                // rows flagged by the bound pass of the last search read back (-1: none yet in this stage)
  int search_mode = 0;                // 0 = pruned (default), 1 = exhaustive (every row evaluated exactly)
  int tree_path = 0;                  // 0 = stumps grown without a host round trip inside the tree (default), 1 = general node-by-node path
  long long stat_searches = 0, stat_seed_rows = 0, stat_flag_rows = 0;
};

int boost_free(cc_ctx* ctx) {
  BoostBuffers* b = ctx->boost;
  if (b) {
    cudaFree(b->weights); cudaFree(b->weak_eval); cudaFree(b->stage_sum); cudaFree(b->sum_response); cudaFree(b->wnode);
    cudaFree(b->treeval); cudaFree(b->responses); cudaFree(b->ynode); cudaFree(b->orig_response); cudaFree(b->mask);
    cudaFree(b->node_of); cudaFree(b->rowbest); cudaFree(b->rowbest_cat); cudaFree(b->d_stats); cudaFree(b->d_best);
    cudaFree(b->d_summary); cudaFree(b->d_tree); cudaFree(b->d_ctl); cudaFree(b->d_seed_rows); cudaFree(b->d_flag_rows);
    cudaFreeHost(b->h_stats); cudaFreeHost(b->h_best); cudaFreeHost(b->h_summary);
    cudaFree(b->d_tree_ok); cudaFreeHost(b->h_tree_ok);
    delete b;
    ctx->boost = nullptr;
  }
  traindata_free(ctx);
  return CC_OK;
}

constexpr int SEED_TARGET = 2048;      // rows evaluated exactly up front to obtain the pruning threshold

static int ensure_boost(cc_ctx* ctx, int N, size_t nrows) {
  if (!ctx->boost) ctx->boost = new BoostBuffers();
  BoostBuffers* b = ctx->boost;
  if (!b->d_stats) {
    CC_CUDA(cudaMalloc(&b->d_stats, sizeof(NodeStats) * CC_MAX_TREE_NODES));
    CC_CUDA(cudaMalloc(&b->d_best, sizeof(BestSplit) * (1 + CC_MAX_RANKS)));
    CC_CUDA(cudaMalloc(&b->d_summary, sizeof(TreeSummary)));
    CC_CUDA(cudaMalloc(&b->d_tree, sizeof(DevNode) * CC_MAX_TREE_NODES));
    CC_CUDA(cudaMalloc(&b->d_ctl, sizeof(SearchCtl)));
    CC_CUDA(cudaMallocHost(&b->h_stats, sizeof(NodeStats) * CC_MAX_TREE_NODES));
    CC_CUDA(cudaMallocHost(&b->h_best, sizeof(BestSplit) * 16));
    CC_CUDA(cudaMallocHost(&b->h_summary, sizeof(TreeSummary)));
    CC_CUDA(cudaMalloc(&b->d_tree_ok, sizeof(int)));
    CC_CUDA(cudaMallocHost(&b->h_tree_ok, sizeof(int)));
  }
  if (N > b->cap_n) {
    cudaFree(b->weights); cudaFree(b->weak_eval); cudaFree(b->stage_sum); cudaFree(b->sum_response); cudaFree(b->wnode);
    cudaFree(b->treeval); cudaFree(b->responses); cudaFree(b->ynode); cudaFree(b->orig_response); cudaFree(b->mask); cudaFree(b->node_of);
    const size_t n = (size_t)N;
    CC_CUDA(cudaMalloc(&b->weights, n * 8)); CC_CUDA(cudaMalloc(&b->weak_eval, n * 8)); CC_CUDA(cudaMalloc(&b->stage_sum, n * 8));
    CC_CUDA(cudaMalloc(&b->sum_response, n * 8)); CC_CUDA(cudaMalloc(&b->wnode, n * 8)); CC_CUDA(cudaMalloc(&b->treeval, n * 8));
    CC_CUDA(cudaMalloc(&b->responses, n * 4)); CC_CUDA(cudaMalloc(&b->ynode, n * 4)); CC_CUDA(cudaMalloc(&b->orig_response, n * 4));
    CC_CUDA(cudaMalloc(&b->mask, n)); CC_CUDA(cudaMalloc(&b->node_of, n * 4));
    b->cap_n = N;
  }
  if (nrows > b->cap_rows) {
    cudaFree(b->rowbest); cudaFree(b->rowbest_cat); cudaFree(b->d_flag_rows); cudaFree(b->d_seed_rows);
    b->rowbest = nullptr; b->rowbest_cat = nullptr; b->d_flag_rows = nullptr; b->d_seed_rows = nullptr;
    if (ctx->ftype == CC_FEATURE_LBP) CC_CUDA(cudaMalloc(&b->rowbest_cat, nrows * sizeof(RowBestCat)));
    else {
      CC_CUDA(cudaMalloc(&b->rowbest, nrows * sizeof(RowBest)));
      CC_CUDA(cudaMalloc(&b->d_flag_rows, nrows * sizeof(int)));
      b->seed_cap = SEED_TARGET + 256;
      CC_CUDA(cudaMalloc(&b->d_seed_rows, b->seed_cap * sizeof(int)));
    }
    b->cap_rows = nrows;
  }
  return CC_OK;
}

// ---- block-level helpers (single-CTA kernels, blockDim = 1024) ---------------------------------------
constexpr int ONE_CTA = 1024;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* scratch /*[32]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = (lane < (int)(blockDim.x >> 5)) ? scratch[lane] : 0.0;
  r = warp_sum(r);
  return r;     // every thread holds the total
}
__device__ __forceinline__ long long block_sum_ll(long long v, double* scratch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  long long* sc = reinterpret_cast<long long*>(scratch);
  __syncthreads();
  if (lane == 0) sc[warp] = v;
  __syncthreads();
  long long r = (lane < (int)(blockDim.x >> 5)) ? sc[lane] : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) r += __shfl_xor_sync(0xffffffffu, r, d);
  return r;
}

__device__ __forceinline__ double d_log_ratio(double val) {   // boost.cpp:27-36
  const double eps = 1e-5;
  val = fmax(val, eps);
  val = fmin(val, 1. - eps);
  return log(val / (1. - val));
}

// ---- CvCascadeBoost::update_weights(nullptr), src/boost.cpp:190-265 ---------------------------------
__global__ void k_boost_init(int N, const uint8_t* __restrict__ labels, int boost_type, double* weights, double* weak_eval,
                             double* stage_sum, double* sum_response, float* responses, int* orig_response, uint8_t* mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int label = labels[i];
  const int o = label * 2 - 1;
  orig_response[i] = o;
  weights[i] = 1. / N;
  weak_eval[i] = 0.;
  stage_sum[i] = 0.;
  sum_response[i] = 0.;
  mask[i] = 1;
  float r = (float)label;
  if (boost_type == CC_BOOST_LOGIT) r = o > 0 ? 2.f : -2.f;
  else if (boost_type == CC_BOOST_GENTLE) r = (float)o;
  responses[i] = r;
}

__global__ void k_tree_begin(int N, const uint8_t* __restrict__ mask, int* node_of) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) node_of[i] = mask[i] ? 0 : -1;
}

// ---- CvBoostTree::calc_node_value, src/o_cvboostree.cpp:657-732 (single CTA) -------------------------
// Also writes the per-sample operand of the split search for this node:
//   regression, +-1 targets (GAB): wnode = y*w (sign carries the target), 0 for non-members
//   regression, free targets (LB): wnode = w, ynode = y
//   classification (DAB/RAB)     : wnode = +w for class 1, -w for class 0
// and weak_eval[s] = node value for every member (leaf assignment of CvBoostTree::try_split_node,
// :74-84; children overwrite it when the node is split).
__global__ void __launch_bounds__(ONE_CTA, 1)
k_node_prepare(int N, int nid0, int is_classifier, int boost_type, const int* __restrict__ node_of, const double* __restrict__ weights,
               const float* __restrict__ responses, double* wnode, float* ynode, double* weak_eval, NodeStats* stats) {
  __shared__ double scratch[32];
  const int nid = nid0 + blockIdx.x;      // one CTA per node: the two leaves of a stump are prepared by one launch
  const bool write_ops = gridDim.x == 1;  // sibling CTAs would overwrite each other's operands; leaves need none (no further search)
  double a0 = 0, a1 = 0, a2 = 0;
  long long c0 = 0, c1 = 0;
  for (int s = threadIdx.x; s < N; s += blockDim.x) {
    if (node_of[s] != nid) { if (write_ops) wnode[s] = 0.0; continue; }
    const double w = weights[s];
    const float y = responses[s];
    if (is_classifier) {
      const int r = (int)y;
      if (r) { a1 += w; c1++; if (write_ops) wnode[s] = w; } else { a0 += w; c0++; if (write_ops) wnode[s] = -w; }
    } else {
      const double t = (double)y;
      a0 += w; a1 += t * w; a2 += t * t * w; c0++;
      if (write_ops) { if (boost_type == CC_BOOST_LOGIT) { wnode[s] = w; ynode[s] = y; } else wnode[s] = t * w; }
    }
  }
  a0 = block_sum(a0, scratch); a1 = block_sum(a1, scratch); a2 = block_sum(a2, scratch);
  c0 = block_sum_ll(c0, scratch); c1 = block_sum_ll(c1, scratch);
  NodeStats st;
  st.pad = 0;
  if (is_classifier) {
    st.n = (int)(c0 + c1); st.cls_count[0] = (int)c0; st.cls_count[1] = (int)c1;
    st.rcw[0] = a0; st.rcw[1] = a1; st.risk = 0; st.rsum0 = 0;
    const int class_idx = a1 > a0;
    if (boost_type == CC_BOOST_DISCRETE) st.value = class_idx * 2 - 1;
    else { const double p = a1 / (a0 + a1); st.value = 0.5 * d_log_ratio(p); }
  } else {
    st.n = (int)c0; st.cls_count[0] = st.cls_count[1] = 0;
    st.rcw[0] = a0; st.rcw[1] = 0;
    const double iw = 1. / a0;
    st.value = a1 * iw;
    double risk = a2 - (a1 * iw) * a1;
    const double n = (double)st.n;
    risk *= n * iw * n * iw;
    st.risk = risk;
    st.rsum0 = st.value * a0;
  }
  if (threadIdx.x == 0) stats[nid] = st;
  const double value = st.value;
  for (int s = threadIdx.x; s < N; s += blockDim.x)
    if (node_of[s] == nid) weak_eval[s] = value;
}

// ------------------------------------------------------------------------------------------------
// Split search, ordered variables: CvBoostTree::find_split_ord_reg / find_split_ord_class
// (src/o_cvboostree.cpp:361-426, :151-247).
//
// One warp scans one row: 256 sorted positions per step (8 per lane, 128-bit loads of the u16 index
// row and the f32 sorted-value row), per-sample operand gathered from shared memory, two running
// sums carried as warp-level exclusive scans in f64.  Samples that are not in the node (trimmed out,
// or in another node) carry operand 0 and are skipped, which is the stable compaction that
// CvCascadeBoostTrainData::subsample_data / split_node_data perform on the index rows
// (src/o_cvcascadeboosttraindata.cpp:80-121, src/o_cvcascadeboosttree.cpp:192-253).  A split between
// consecutive node samples p < j is admissible iff v[p] + 2*FLT_EPSILON < v[j] (f32).  Per row the
// first position attaining the maximal f64 criterion wins (strict '<' of the reference).
// ------------------------------------------------------------------------------------------------
constexpr int SS_THREADS = 512;
constexpr int SS_WARPS = SS_THREADS / 32;
constexpr int SS_EPT = 8;
constexpr int SS_TILE = 32 * SS_EPT;

template <int CRIT>
__device__ __forceinline__ double criterion(double A, double B, double totA, double totB) {
  if (CRIT == CRIT_SQERR) {         // A = L (sum w), B = lsum (sum w*y); totA = sum w, totB = value*sum w
    const double L = A, R = totA - A, lsum = B, rsum = totB - B;
    if (!(L > 0.0) || !(R > 0.0)) return -DBL_MAX;
    return (lsum * lsum * R + rsum * rsum * L) / (L * R);
  } else if (CRIT == CRIT_GINI) {   // A = lcw[0], B = lcw[1]
    const double r0 = totA - A, r1 = totB - B, L = A + B, R = r0 + r1;
    if (!(L > 0.0) || !(R > 0.0)) return -DBL_MAX;
    const double lsum2 = A * A + B * B, rsum2 = r0 * r0 + r1 * r1;
    return (lsum2 * R + rsum2 * L) / (L * R);
  } else {                          // misclassification
    const double r0 = totA - A, r1 = totB - B;
    return fmax(A + r1, B + r0);
  }
}

// f32 estimate of the same criterion, used only to decide whether a candidate can possibly beat the
// running best; every candidate that survives the filter is re-evaluated with criterion<>() in f64,
// so the selected split is the one full-f64 evaluation would select.
template <int CRIT>
__device__ __forceinline__ float criterion32(float A, float B, float totA, float totB) {
  if (CRIT == CRIT_SQERR) {
    const float L = A, R = totA - A, lsum = B, rsum = totB - B;
    return __fdividef(lsum * lsum * R + rsum * rsum * L, L * R);
  } else if (CRIT == CRIT_GINI) {
    const float r0 = totA - A, r1 = totB - B, L = A + B, R = r0 + r1;
    return __fdividef((A * A + B * B) * R + (r0 * r0 + r1 * r1) * L, L * R);
  } else {
    const float r0 = totA - A, r1 = totB - B;
    return fmaxf(A + r1, B + r0);
  }
}

// exact numerator / denominator of the criterion (f64, the reference's formula); criteria without a division use den = 1
template <int CRIT>
__device__ __forceinline__ bool criterion_frac(double A, double B, double totA, double totB, double& num, double& den) {
  if (CRIT == CRIT_SQERR) {
    const double L = A, R = totA - A, lsum = B, rsum = totB - B;
    if (!(L > 0.0) || !(R > 0.0)) return false;
    num = lsum * lsum * R + rsum * rsum * L; den = L * R;
  } else if (CRIT == CRIT_GINI) {
    const double r0 = totA - A, r1 = totB - B, L = A + B, R = r0 + r1;
    if (!(L > 0.0) || !(R > 0.0)) return false;
    num = (A * A + B * B) * R + (r0 * r0 + r1 * r1) * L; den = L * R;
  } else {
    const double r0 = totA - A, r1 = totB - B;
    num = fmax(A + r1, B + r0); den = 1.0;
  }
  return true;
}

// per-tile state saved by the estimate pass so that the exact pass can resume at any tile
struct TileCarry { double a, b; float v; int cnt_has; };     // cnt_has = (count << 1) | has

// One 256-position tile of one row.  EXACT == false: f32 estimate of every admissible split, returns the lane's
// maximum.  EXACT == true: re-evaluates candidates whose estimate reaches `lim` with the reference's f64 formula,
// in position order, keeping the first maximum (compared as exact fractions).
// TOTALS (with EXACT == false): only the carries are advanced (the tile's totals when they come in as zero).
template <int CRIT, bool HAS_Y, bool EXACT, bool TOTALS = false>
__device__ __forceinline__ float ss_tile(const uint4 ci, const float4 cv0, const float4 cv1, uint32_t sw_addr, uint32_t sy_addr,
                                         int lane, uint32_t lt_mask, double totA, double totB, float ftotA, float ftotB,
                                         double& carryA, double& carryB, float& carryV, int& carryHas, int& carryCnt, float lim,
                                         double& bestNum, double& bestDen, int& bestPos, float& bestVl, float& bestVr) {
  const float eps = FLT_EPSILON * 2;
  const uint32_t iw[4] = {ci.x, ci.y, ci.z, ci.w};
  const float v[SS_EPT] = {cv0.x, cv0.y, cv0.z, cv0.w, cv1.x, cv1.y, cv1.z, cv1.w};
  // ---- gather operands, lane totals, "previous node sample" values, candidate mask ----
  double a[SS_EPT], b[SS_EPT];
  float af[SS_EPT], bf[SS_EPT], pvv[SS_EPT];
  unsigned act = 0, inner = 0;      // inner: candidate whose left neighbour lies in this lane's chunk
  double ta = 0, tb = 0;
  float lastv = 0.f;
#pragma unroll
  for (int j = 0; j < SS_EPT; ++j) {
    const uint32_t idx = (iw[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;   // row padding holds index N (operand 0)
    double w;
    asm("ld.shared.f64 %0, [%1];" : "=d"(w) : "r"(sw_addr + idx * 8u));   // 0 marks "not in this node"
    const float wf = (float)w;
    const bool on = w != 0.0;
    pvv[j] = lastv;
    if (on && act != 0 && __fadd_rn(lastv, eps) < v[j]) inner |= 1u << j;
    act |= on ? (1u << j) : 0u;
    lastv = on ? v[j] : lastv;
    if (CRIT == CRIT_SQERR) {
      if (HAS_Y) {
        float y;
        asm("ld.shared.f32 %0, [%1];" : "=f"(y) : "r"(sy_addr + idx * 4u));
        a[j] = w; b[j] = (double)y * w; af[j] = wf; bf[j] = y * wf;
      } else { a[j] = fabs(w); b[j] = w; af[j] = fabsf(wf); bf[j] = wf; }
    } else {
      a[j] = w < 0.0 ? -w : 0.0; b[j] = w > 0.0 ? w : 0.0;
      af[j] = wf < 0.f ? -wf : 0.f; bf[j] = wf > 0.f ? wf : 0.f;
    }
    ta += a[j]; tb += b[j];
  }
  const int cnt = __popc(act);
  // ---- warp exclusive scans of (ta, tb, cnt) ----
  double ia = ta, ib = tb;
  int ic = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double xa = __shfl_up_sync(0xffffffffu, ia, d), xb = __shfl_up_sync(0xffffffffu, ib, d);
    const int xc = __shfl_up_sync(0xffffffffu, ic, d);
    if (lane >= d) { ia += xa; ib += xb; ic += xc; }
  }
  const double pa0 = carryA + (ia - ta), pb0 = carryB + (ib - tb);    // sums over node samples before this chunk
  const int pc0 = carryCnt + (ic - cnt);
  // value of the last node sample before this lane's chunk (for the first node sample of the chunk)
  const uint32_t has_mask = __ballot_sync(0xffffffffu, act != 0);
  const uint32_t before = has_mask & lt_mask;
  const int src = before ? (31 - __clz(before)) : 0;
  const float pv_other = __shfl_sync(0xffffffffu, lastv, src);
  const float pv_in = before ? pv_other : carryV;
  const bool pHas_in = before ? true : (carryHas != 0);
  unsigned cand = inner;
  const int first = act ? (__ffs(act) - 1) : 0;
  if (act && pHas_in && __fadd_rn(pv_in, eps) < v[first]) cand |= 1u << first;
  // ---- f32 estimate of the criterion at every candidate ----
  float fa = (float)pa0, fb = (float)pb0;
  float tmax = -3.0e38f;
  unsigned pass = 0;
  if (!TOTALS) {
#pragma unroll
    for (int j = 0; j < SS_EPT; ++j) {
      const float est = criterion32<CRIT>(fa, fb, ftotA, ftotB);
      if (EXACT) { if (!(est < lim)) pass |= 1u << j; }
      else tmax = ((cand >> j) & 1u) ? fmaxf(tmax, est == est ? est : 3.0e38f) : tmax;     // NaN (0/0 at the ends) counts as "unknown, maybe large"
      fa += af[j]; fb += bf[j];
    }
  }
  if (EXACT) {
    pass &= cand;
    double pa = pa0, pb = pb0;
    int pc = pc0;
#pragma unroll
    for (int j = 0; j < SS_EPT; ++j) {
      if (pass & (1u << j)) {
        double num, den;
        if (criterion_frac<CRIT>(pa, pb, totA, totB, num, den) && num * bestDen > bestNum * den) {
          bestNum = num; bestDen = den; bestPos = pc - 1;
          bestVl = (j == first) ? pv_in : pvv[j];
          bestVr = v[j];
        }
      }
      pa += a[j]; pb += b[j];
      pc += (act >> j) & 1u;
    }
  }
  // ---- carries for the next tile ----
  carryA += __shfl_sync(0xffffffffu, ia, 31);
  carryB += __shfl_sync(0xffffffffu, ib, 31);
  carryCnt += __shfl_sync(0xffffffffu, ic, 31);
  {
    const int lastLane = has_mask ? 31 - __clz(has_mask) : 0;
    const float lv = __shfl_sync(0xffffffffu, lastv, lastLane);
    if (has_mask) { carryV = lv; carryHas = 1; }
  }
  return tmax;
}

template <int CRIT, bool HAS_Y>
__global__ void __launch_bounds__(SS_THREADS, 1)
k_split_search_ord(const uint16_t* __restrict__ sidx, const float* __restrict__ sval, int n_stride, int N, int nrows,
                   const double* __restrict__ wnode, const float* __restrict__ ynode, const NodeStats* __restrict__ stats,
                   double filter_ymax2, RowBest* __restrict__ out, const int* __restrict__ row_list,
                   const unsigned int* __restrict__ list_count, SearchCtl* __restrict__ ctl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // rows to search: all local rows, or the entries of row_list (seed rows / rows flagged by k_ss_bound)
  const int nwork = row_list ? (list_count ? (int)min(*list_count, (unsigned int)nrows) : nrows) : nrows;
  if (blockIdx.x * SS_WARPS >= nwork) return;
  const int ntile = (N + SS_TILE - 1) / SS_TILE;
  double* sw = reinterpret_cast<double*>(smem_raw);          // [N+1]; entry N = 0 is what the row padding points at
  TileCarry* carries_all = reinterpret_cast<TileCarry*>(sw + N + 1);                 // [SS_WARPS][ntile]
  float* tilemax_all = reinterpret_cast<float*>(carries_all + (size_t)SS_WARPS * ntile);   // [SS_WARPS][ntile]
  float* sy = tilemax_all + (size_t)SS_WARPS * ntile;        // [N+1] (LogitBoost only)
  for (int i = threadIdx.x; i <= N; i += SS_THREADS) { sw[i] = i < N ? wnode[i] : 0.0; if (HAS_Y) sy[i] = i < N ? ynode[i] : 0.f; }
  __syncthreads();
  // 32-bit shared-window addresses: the per-sample gathers become plain LDS with one address add
  const uint32_t sw_addr = (uint32_t)__cvta_generic_to_shared(sw);
  const uint32_t sy_addr = (uint32_t)__cvta_generic_to_shared(sy);
  const NodeStats st = *stats;
  const double totA = st.rcw[0];
  const double totB = (CRIT == CRIT_SQERR) ? st.rsum0 : st.rcw[1];
  const float ftotA = (float)totA, ftotB = (float)totB;
  // absolute slack of the f32 pre-filter: the f32 prefix sums carry ~1e-6 * (node weight) * |y|max^2 of absolute error
  // (see criterion32); LogitBoost targets are clipped to |z| <= 10 (boost.cpp:347).  Taken from the node statistics on the
  // device, so that the host does not have to read them back before it launches the search.
  const float filter_abs = (float)(4e-6 * (st.rcw[0] + st.rcw[1]) * filter_ymax2 + 1e-30);
  // broadcast from lane 0 so that the compiler knows the row loop is warp-uniform (no collective-sync
  // wrappers around the shuffles inside it)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  TileCarry* carries = carries_all + (size_t)warp * ntile;
  float* tilemax = tilemax_all + (size_t)warp * ntile;

  for (int wi = blockIdx.x * SS_WARPS + warp; wi < nwork; wi += gridDim.x * SS_WARPS) {
    const int row = row_list ? row_list[wi] : wi;
    const uint16_t* ri = sidx + (size_t)row * n_stride;
    const float* rv = sval + (size_t)row * n_stride;
    double carryA = 0, carryB = 0;
    float carryV = 0.f;
    int carryHas = 0, carryCnt = 0;
    double bestNum = -1.0, bestDen = 1.0;      // running best as an exact fraction
    int bestPos = -1;
    float bestVl = 0.f, bestVr = 0.f;
    float rmax = -3.0e38f;

    // ---------------- pass 1: f32 estimates, per-tile maxima, per-tile carries ----------------
    uint4 nxt_i = *reinterpret_cast<const uint4*>(ri + lane * SS_EPT);
    float4 nxt_v0 = *reinterpret_cast<const float4*>(rv + lane * SS_EPT);
    float4 nxt_v1 = *reinterpret_cast<const float4*>(rv + lane * SS_EPT + 4);
    for (int t = 0; t < ntile; ++t) {
      const uint4 ci = nxt_i;
      const float4 cv0 = nxt_v0, cv1 = nxt_v1;
      if (t + 1 < ntile) {           // prefetch the next tile while this one is processed
        const int o = (t + 1) * SS_TILE + lane * SS_EPT;
        nxt_i = *reinterpret_cast<const uint4*>(ri + o);
        nxt_v0 = *reinterpret_cast<const float4*>(rv + o);
        nxt_v1 = *reinterpret_cast<const float4*>(rv + o + 4);
      }
      if (lane == 0) { carries[t].a = carryA; carries[t].b = carryB; carries[t].v = carryV; carries[t].cnt_has = (carryCnt << 1) | carryHas; }
      float tm = ss_tile<CRIT, HAS_Y, false>(ci, cv0, cv1, sw_addr, sy_addr, lane, lt_mask, totA, totB, ftotA, ftotB, carryA, carryB, carryV,
                                             carryHas, carryCnt, 0.f, bestNum, bestDen, bestPos, bestVl, bestVr);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, d));
      if (lane == 0) tilemax[t] = tm;
      rmax = fmaxf(rmax, tm);
    }
    __syncwarp();
    // ---------------- pass 2: exact evaluation of the few tiles that can hold the maximum ----------------
    const float lim = rmax - (1e-5f * fabsf(rmax) + filter_abs);
    for (int t = 0; t < ntile; ++t) {
      if (!(tilemax[t] >= lim)) continue;           // warp-uniform (shared memory broadcast)
      const int o = t * SS_TILE + lane * SS_EPT;
      const uint4 ci = *reinterpret_cast<const uint4*>(ri + o);
      const float4 cv0 = *reinterpret_cast<const float4*>(rv + o), cv1 = *reinterpret_cast<const float4*>(rv + o + 4);
      double cA = carries[t].a, cB = carries[t].b;
      float cV = carries[t].v;
      int cH = carries[t].cnt_has & 1, cC = carries[t].cnt_has >> 1;
      ss_tile<CRIT, HAS_Y, true>(ci, cv0, cv1, sw_addr, sy_addr, lane, lt_mask, totA, totB, ftotA, ftotB, cA, cB, cV, cH, cC, lim, bestNum,
                                 bestDen, bestPos, bestVl, bestVr);
    }
    __syncwarp();
    // warp argmax: larger value first, then smaller position (first maximum of the sequential scan)
    double bestVal = bestPos >= 0 ? bestNum / bestDen : -1.0;
    int owner = lane;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bestVal, d);
      const int op = __shfl_xor_sync(0xffffffffu, bestPos, d);
      const int oo = __shfl_xor_sync(0xffffffffu, owner, d);
      const bool take = op >= 0 && (bestPos < 0 || ov > bestVal || (ov == bestVal && op < bestPos));
      if (take) { bestVal = ov; bestPos = op; owner = oo; }
    }
    const float vl = __shfl_sync(0xffffffffu, bestVl, owner), vr = __shfl_sync(0xffffffffu, bestVr, owner);
    if (lane == 0) {
      RowBest rb;
      rb.q = bestPos >= 0 ? (float)bestVal : -1.f;
      rb.c = __fmul_rn(__fadd_rn(vl, vr), 0.5f);
      rb.vleft = vl;
      rb.pos = bestPos;
      out[row] = rb;
      // running argmax over rows: highest f32 quality, lowest row on ties (CvDTree::find_best_split, o_cvdtree.cpp:340)
      if (rb.q > 0.f) {
        const unsigned long long k = ((unsigned long long)__float_as_uint(rb.q) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)row);
        if (k > *reinterpret_cast<volatile unsigned long long*>(&ctl->key)) atomicMax(&ctl->key, k);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same search with one CTA per row, for the short row lists of the pruned search (seed rows of a feature shard, rows flagged
// by k_ss_bound).  k_split_search_ord gives a row to one warp, which walks its 59 tiles one after the other: ~85 us whatever
// the number of rows -- the floor of a tree node on a 1/8 feature shard.  Here the 16 warps of a CTA take contiguous segments
// of the row's tiles: (0) every tile's totals (the carries of ss_tile, started from zero), (1) one thread chains them in tile
// order -- the same additions in the same order as the running carries of the one-warp walk, so every tile sees bit-identical
// carries --, (2) f32 estimates per tile and the row maximum, (3) exact evaluation of the tiles that can hold it, (4) arg max
// over the warps with the rule of the warp-level arg max (larger quality, then smaller position).  Same result bits.
// ------------------------------------------------------------------------------------------------
template <int CRIT, bool HAS_Y>
__global__ void __launch_bounds__(SS_THREADS, 1)
k_split_search_ord_coop(const uint16_t* __restrict__ sidx, const float* __restrict__ sval, int n_stride, int N, int nrows,
                        const double* __restrict__ wnode, const float* __restrict__ ynode, const NodeStats* __restrict__ stats,
                        double filter_ymax2, RowBest* __restrict__ out, const int* __restrict__ row_list,
                        const unsigned int* __restrict__ list_count, SearchCtl* __restrict__ ctl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nwork = list_count ? (int)min(*list_count, (unsigned int)nrows) : nrows;
  if ((int)blockIdx.x >= nwork) return;
  const int ntile = (N + SS_TILE - 1) / SS_TILE;
  double* sw = reinterpret_cast<double*>(smem_raw);                      // [N+1]; entry N = 0 is what the row padding points at
  TileCarry* carries = reinterpret_cast<TileCarry*>(sw + N + 1);         // [ntile]: carries INTO tile t
  TileCarry* tot = carries + ntile;                                      // [ntile]: totals OF tile t
  float* tilemax = reinterpret_cast<float*>(tot + ntile);                // [ntile]
  float* sy = tilemax + ntile;                                           // [N+1] (LogitBoost only)
  __shared__ float s_wmax[SS_WARPS], s_vl[SS_WARPS], s_vr[SS_WARPS];
  __shared__ double s_bval[SS_WARPS];
  __shared__ int s_bpos[SS_WARPS];
  for (int i = threadIdx.x; i <= N; i += SS_THREADS) { sw[i] = i < N ? wnode[i] : 0.0; if (HAS_Y) sy[i] = i < N ? ynode[i] : 0.f; }
  __syncthreads();
  const uint32_t sw_addr = (uint32_t)__cvta_generic_to_shared(sw);
  const uint32_t sy_addr = (uint32_t)__cvta_generic_to_shared(sy);
  const NodeStats st = *stats;
  const double totA = st.rcw[0];
  const double totB = (CRIT == CRIT_SQERR) ? st.rsum0 : st.rcw[1];
  const float ftotA = (float)totA, ftotB = (float)totB;
  const float filter_abs = (float)(4e-6 * (st.rcw[0] + st.rcw[1]) * filter_ymax2 + 1e-30);       // as k_split_search_ord
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int per = (ntile + SS_WARPS - 1) / SS_WARPS;
  const int t0 = min(ntile, warp * per), t1 = min(ntile, t0 + per);

  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const int row = row_list ? row_list[wi] : wi;
    const uint16_t* ri = sidx + (size_t)row * n_stride;
    const float* rv = sval + (size_t)row * n_stride;
    double bestNum = -1.0, bestDen = 1.0;
    int bestPos = -1;
    float bestVl = 0.f, bestVr = 0.f;
    // ---------------- (0) totals of my tiles ----------------
    for (int t = t0; t < t1; ++t) {
      const int o = t * SS_TILE + lane * SS_EPT;
      const uint4 ci = *reinterpret_cast<const uint4*>(ri + o);
      const float4 cv0 = *reinterpret_cast<const float4*>(rv + o), cv1 = *reinterpret_cast<const float4*>(rv + o + 4);
      double cA = 0, cB = 0;
      float cV = 0.f;
      int cH = 0, cC = 0;
      ss_tile<CRIT, HAS_Y, false, true>(ci, cv0, cv1, sw_addr, sy_addr, lane, lt_mask, totA, totB, ftotA, ftotB, cA, cB, cV, cH, cC, 0.f, bestNum,
                                        bestDen, bestPos, bestVl, bestVr);
      if (lane == 0) { tot[t].a = cA; tot[t].b = cB; tot[t].v = cV; tot[t].cnt_has = (cC << 1) | cH; }
    }
    __syncthreads();
    // ---------------- (1) carries into every tile: the totals chained in tile order ----------------
    if (threadIdx.x == 0) {
      double a = 0, b = 0;
      float v = 0.f;
      int has = 0, cnt = 0;
      for (int t = 0; t < ntile; ++t) {
        carries[t].a = a; carries[t].b = b; carries[t].v = v; carries[t].cnt_has = (cnt << 1) | has;
        a += tot[t].a; b += tot[t].b; cnt += tot[t].cnt_has >> 1;
        if (tot[t].cnt_has & 1) { v = tot[t].v; has = 1; }
      }
    }
    __syncthreads();
    // ---------------- (2) f32 estimates, per-tile maxima, row maximum ----------------
    float rmax = -3.0e38f;
    for (int t = t0; t < t1; ++t) {
      const int o = t * SS_TILE + lane * SS_EPT;
      const uint4 ci = *reinterpret_cast<const uint4*>(ri + o);
      const float4 cv0 = *reinterpret_cast<const float4*>(rv + o), cv1 = *reinterpret_cast<const float4*>(rv + o + 4);
      double cA = carries[t].a, cB = carries[t].b;
      float cV = carries[t].v;
      int cH = carries[t].cnt_has & 1, cC = carries[t].cnt_has >> 1;
      float tm = ss_tile<CRIT, HAS_Y, false>(ci, cv0, cv1, sw_addr, sy_addr, lane, lt_mask, totA, totB, ftotA, ftotB, cA, cB, cV, cH, cC, 0.f,
                                             bestNum, bestDen, bestPos, bestVl, bestVr);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, d));
      if (lane == 0) tilemax[t] = tm;
      rmax = fmaxf(rmax, tm);
    }
    if (lane == 0) s_wmax[warp] = rmax;
    __syncthreads();
    rmax = s_wmax[lane & (SS_WARPS - 1)];
#pragma unroll
    for (int d = SS_WARPS / 2; d > 0; d >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, d));
    // ---------------- (3) exact evaluation of my tiles that can hold the maximum ----------------
    const float lim = rmax - (1e-5f * fabsf(rmax) + filter_abs);
    for (int t = t0; t < t1; ++t) {
      if (!(tilemax[t] >= lim)) continue;           // warp-uniform (shared memory broadcast)
      const int o = t * SS_TILE + lane * SS_EPT;
      const uint4 ci = *reinterpret_cast<const uint4*>(ri + o);
      const float4 cv0 = *reinterpret_cast<const float4*>(rv + o), cv1 = *reinterpret_cast<const float4*>(rv + o + 4);
      double cA = carries[t].a, cB = carries[t].b;
      float cV = carries[t].v;
      int cH = carries[t].cnt_has & 1, cC = carries[t].cnt_has >> 1;
      ss_tile<CRIT, HAS_Y, true>(ci, cv0, cv1, sw_addr, sy_addr, lane, lt_mask, totA, totB, ftotA, ftotB, cA, cB, cV, cH, cC, lim, bestNum,
                                 bestDen, bestPos, bestVl, bestVr);
    }
    // ---------------- (4) arg max: lanes, then warps (larger value first, then smaller position) ----------------
    double bestVal = bestPos >= 0 ? bestNum / bestDen : -1.0;
    {
      int owner = lane;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bestVal, d);
        const int op = __shfl_xor_sync(0xffffffffu, bestPos, d);
        const int oo = __shfl_xor_sync(0xffffffffu, owner, d);
        const bool take = op >= 0 && (bestPos < 0 || ov > bestVal || (ov == bestVal && op < bestPos));
        if (take) { bestVal = ov; bestPos = op; owner = oo; }
      }
      const float vl = __shfl_sync(0xffffffffu, bestVl, owner), vr = __shfl_sync(0xffffffffu, bestVr, owner);
      if (lane == 0) { s_bval[warp] = bestVal; s_bpos[warp] = bestPos; s_vl[warp] = vl; s_vr[warp] = vr; }
    }
    __syncthreads();
    if (warp == 0) {
      const int src = lane & (SS_WARPS - 1);
      double bv = s_bval[src];
      int bp = s_bpos[src], owner = src;
#pragma unroll
      for (int d = SS_WARPS / 2; d > 0; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, d);
        const int op = __shfl_xor_sync(0xffffffffu, bp, d);
        const int oo = __shfl_xor_sync(0xffffffffu, owner, d);
        const bool take = op >= 0 && (bp < 0 || ov > bv || (ov == bv && op < bp));
        if (take) { bv = ov; bp = op; owner = oo; }
      }
      if (lane == 0) {
        const float vl = s_vl[owner], vr = s_vr[owner];
        RowBest rb;
        rb.q = bp >= 0 ? (float)bv : -1.f;
        rb.c = __fmul_rn(__fadd_rn(vl, vr), 0.5f);
        rb.vleft = vl;
        rb.pos = bp;
        out[row] = rb;
        if (rb.q > 0.f) {
          const unsigned long long k = ((unsigned long long)__float_as_uint(rb.q) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)row);
          if (k > *reinterpret_cast<volatile unsigned long long*>(&ctl->key)) atomicMax(&ctl->key, k);
        }
      }
    }
    __syncthreads();                                 // tot / carries / tilemax / s_* are reused by the next row
  }
}

// ------------------------------------------------------------------------------------------------
// Ordered split search, wide path (numSamples beyond the shared-memory kernels above; 32-bit sorted-index rows): one warp per
// row, 32 sorted positions per step, per-sample operands and feature values gathered from global memory (operands: L1 / L2
// resident, N * 8 bytes; values: fval[row][index]).  Same arithmetic and same selection as ss_tile<EXACT>: operand 0 = not in
// the node, a split between consecutive node samples p < j is admissible iff v[p] + 2 * FLT_EPSILON < v[j], candidates are
// compared as exact fractions, first maximum wins.  Every row is searched exactly (no pruning pass).
// ------------------------------------------------------------------------------------------------
constexpr int SW_THREADS = 256;
template <int CRIT, bool HAS_Y>
__global__ void __launch_bounds__(SW_THREADS)
k_split_search_ord_wide(const uint32_t* __restrict__ sidx32, const float* __restrict__ fval, int n_stride, int N, int nrows,
                        const double* __restrict__ wnode, const float* __restrict__ ynode, const NodeStats* __restrict__ stats,
                        RowBest* __restrict__ out) {
  const NodeStats st = *stats;
  const double totA = st.rcw[0];
  const double totB = (CRIT == CRIT_SQERR) ? st.rsum0 : st.rcw[1];
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const float eps = FLT_EPSILON * 2;
  const int nwarps = (gridDim.x * SW_THREADS) >> 5;
  for (int row = (blockIdx.x * SW_THREADS + threadIdx.x) >> 5; row < nrows; row += nwarps) {
    const uint32_t* ri = sidx32 + (size_t)row * n_stride;
    const float* fv = fval + (size_t)row * n_stride;
    double carryA = 0, carryB = 0;
    float carryV = 0.f;
    bool carryHas = false;
    int carryCnt = 0;
    double bestNum = -1.0, bestDen = 1.0;
    int bestPos = -1;
    float bestVl = 0.f, bestVr = 0.f;
    for (int base = 0; base < N; base += 32) {
      const int p = base + lane;
      double w = 0.0;
      float v = 0.f, y = 0.f;
      if (p < N) {
        const uint32_t idx = ri[p];
        w = __ldg(wnode + idx);
        if (w != 0.0) { v = __ldg(fv + idx); if (HAS_Y) y = __ldg(ynode + idx); }
      }
      const bool on = w != 0.0;
      double a, b;
      if (CRIT == CRIT_SQERR) {
        if (HAS_Y) { a = w; b = (double)y * w; } else { a = fabs(w); b = w; }
      } else { a = w < 0.0 ? -w : 0.0; b = w > 0.0 ? w : 0.0; }
      double ia = a, ib = b;
      int ic = on ? 1 : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double xa = __shfl_up_sync(0xffffffffu, ia, d), xb = __shfl_up_sync(0xffffffffu, ib, d);
        const int xc = __shfl_up_sync(0xffffffffu, ic, d);
        if (lane >= d) { ia += xa; ib += xb; ic += xc; }
      }
      const double pa = carryA + (ia - a), pb = carryB + (ib - b);          // sums over the node samples before this one
      const int pc = carryCnt + (ic - (on ? 1 : 0));
      const uint32_t onmask = __ballot_sync(0xffffffffu, on);
      const uint32_t before = onmask & lt_mask;
      const int src = before ? (31 - __clz(before)) : 0;
      const float pv_other = __shfl_sync(0xffffffffu, v, src);
      const float pv = before ? pv_other : carryV;
      const bool pHas = before ? true : carryHas;
      if (on && pHas && __fadd_rn(pv, eps) < v) {
        double num, den;
        if (criterion_frac<CRIT>(pa, pb, totA, totB, num, den) && num * bestDen > bestNum * den) {
          bestNum = num; bestDen = den; bestPos = pc - 1; bestVl = pv; bestVr = v;
        }
      }
      carryA += __shfl_sync(0xffffffffu, ia, 31);
      carryB += __shfl_sync(0xffffffffu, ib, 31);
      carryCnt += __shfl_sync(0xffffffffu, ic, 31);
      if (onmask) {
        const int lastl = 31 - __clz(onmask);
        carryV = __shfl_sync(0xffffffffu, v, lastl);
        carryHas = true;
      }
    }
    double bestVal = bestPos >= 0 ? bestNum / bestDen : -1.0;
    int owner = lane;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bestVal, d);
      const int op = __shfl_xor_sync(0xffffffffu, bestPos, d);
      const int oo = __shfl_xor_sync(0xffffffffu, owner, d);
      const bool take = op >= 0 && (bestPos < 0 || ov > bestVal || (ov == bestVal && op < bestPos));
      if (take) { bestVal = ov; bestPos = op; owner = oo; }
    }
    const float vl = __shfl_sync(0xffffffffu, bestVl, owner), vr = __shfl_sync(0xffffffffu, bestVr, owner);
    if (lane == 0) {
      RowBest rb;
      rb.q = bestPos >= 0 ? (float)bestVal : -1.f;
      rb.c = __fmul_rn(__fadd_rn(vl, vr), 0.5f);
      rb.vleft = vl;
      rb.pos = bestPos;
      out[row] = rb;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pruning pass of the ordered split search (all criteria over two class weights: GAB squared error with
// +-1 targets, Gini, misclassification).
//
// Write the state of the reference's running sums after a prefix of the sorted node samples as
// (L, D) = (sum |x|, sum x) with the signed per-sample operand x (y*w for GAB, +-w by class otherwise).
// Every criterion is a convex function of (L, D):
//   squared error  C(L,D) = D^2/L + (S-D)^2/(T-L)            (o_cvboostree.cpp:398-414)
//   Gini           (lsum2*R + rsum2*L)/(L*R) = T/2 + C/2      (:205-225, with lsum2 = (L^2+D^2)/2)
//   misclass       max(T1 - D, T0 + D)                        (:226-243)
// While the scan walks through a chunk of consecutive sorted positions, (L, D) moves by steps (|x|, x), so
// it stays inside the parallelogram spanned by P*(1,1) and M*(1,-1) (P, M = sums of the positive / negative
// operands of the chunk).  A convex function attains its maximum over that parallelogram at a corner, hence
//   max over the chunk's split candidates <= max over 3 corners (the 4th is the previous chunk's end).
// A row none of whose chunks can beat the best quality already known (the exact result over the seed rows,
// minus a margin that covers the f32 arithmetic of this pass) cannot hold the winning split and is dropped;
// the others are appended to the flagged list and evaluated exactly by k_split_search_ord.  The pass reads
// only the u16 index row (2 B per element) and needs no sorted values: the bound holds for admissible and
// inadmissible (tied) positions alike, and samples outside the node carry operand 0.
// ------------------------------------------------------------------------------------------------
constexpr int SB_THREADS = 512;
constexpr int SB_WARPS = SB_THREADS / 32;
constexpr int SB_EPT = 16;
constexpr int SB_TILE = 32 * SB_EPT;

template <bool MISCLASS>
__device__ __forceinline__ bool sb_corner(float L, float D, float T, float S, float Gp, float eps) {
  if (MISCLASS) return fmaxf(S - D, T + D) > Gp;          // T = T0, S = T1
  const float Lc = fmaxf(L, eps), Rc = fmaxf(T - L, eps), Dr = S - D;
  const float num = fmaf(D * D, Rc, (Dr * Dr) * Lc), den = Lc * Rc;
  return fmaf(-Gp, den, num) > 0.f;
}

template <bool MISCLASS>
__global__ void __launch_bounds__(SB_THREADS, 3)
k_ss_bound(const uint16_t* __restrict__ sidx, int n_stride, int N, int nrows, const double* __restrict__ wnode,
           const NodeStats* __restrict__ stats, int crit, SearchCtl* __restrict__ ctl, int* __restrict__ flag_rows) {
  extern __shared__ __align__(16) float sx[];                // [N+1]; entry N = 0 is what the row padding points at
  for (int i = threadIdx.x; i <= N; i += SB_THREADS) sx[i] = i < N ? (float)wnode[i] : 0.f;
  __syncthreads();
  const NodeStats st = *stats;
  const float g0 = __uint_as_float((unsigned)(ctl->key >> 32));   // best exact quality so far (0: none)
  float T, S, Gp, eps;
  {
    const double W = crit == CRIT_SQERR ? st.rcw[0] : st.rcw[0] + st.rcw[1];     // total node weight
    // margin: f32 prefix sums (<= 1.5e-6 W abs. error on L and D, |dC/dL| <= 1, |dC/dD| <= 4), clamping of L, R at
    // eps = 4e-6 W (<= 2 eps), f32 evaluation of the test (<= 1e-6 C) -- see DESIGN.md
    const double m = 1e-5 * fabs((double)g0) + 2.4e-5 * W;
    eps = (float)(4e-6 * W);
    if (MISCLASS) { T = (float)st.rcw[0]; S = (float)st.rcw[1]; Gp = (float)((double)g0 - m); }
    else if (crit == CRIT_GINI) { T = (float)W; S = (float)(st.rcw[1] - st.rcw[0]); Gp = (float)(2.0 * ((double)g0 - m) - W); }
    else { T = (float)W; S = (float)st.rsum0; Gp = (float)((double)g0 - m); }
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int ntile = (n_stride + SB_TILE - 1) / SB_TILE;
  const uint32_t padw = (uint32_t)N | ((uint32_t)N << 16);
  const uint4 pad4 = make_uint4(padw, padw, padw, padw);

  for (int row = blockIdx.x * SB_WARPS + warp; row < nrows; row += gridDim.x * SB_WARPS) {
    const uint16_t* ri = sidx + (size_t)row * n_stride;
    double carryL = 0.0, carryD = 0.0;
    bool fire = false;
    uint4 n0 = pad4, n1 = pad4;
    { const int o = lane * SB_EPT; if (o < n_stride) { n0 = __ldcs(reinterpret_cast<const uint4*>(ri + o)); n1 = __ldcs(reinterpret_cast<const uint4*>(ri + o + 8)); } }
    for (int t = 0; t < ntile; ++t) {
      const uint4 c0 = n0, c1 = n1;
      { const int o = (t + 1) * SB_TILE + lane * SB_EPT;
        if (o < n_stride) { n0 = __ldcs(reinterpret_cast<const uint4*>(ri + o)); n1 = __ldcs(reinterpret_cast<const uint4*>(ri + o + 8)); } else { n0 = pad4; n1 = pad4; } }
      const uint32_t iw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      float ta = 0.f, td = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x0 = sx[iw[j] & 0xFFFFu], x1 = sx[iw[j] >> 16];
        ta += fabsf(x0); td += x0;
        ta += fabsf(x1); td += x1;
      }
      float ia = ta, id = td;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float xa = __shfl_up_sync(0xffffffffu, ia, d), xd = __shfl_up_sync(0xffffffffu, id, d);
        if (lane >= d) { ia += xa; id += xd; }
      }
      const float L0 = (float)carryL + (ia - ta), D0 = (float)carryD + (id - td);
      const float P = 0.5f * (ta + td), M = 0.5f * (ta - td);
      fire = sb_corner<MISCLASS>(L0 + P, D0 + P, T, S, Gp, eps) | sb_corner<MISCLASS>(L0 + M, D0 - M, T, S, Gp, eps) |
             sb_corner<MISCLASS>(L0 + ta, D0 + td, T, S, Gp, eps);
      if (__any_sync(0xffffffffu, fire)) { fire = true; break; }
      carryL += (double)__shfl_sync(0xffffffffu, ia, 31);
      carryD += (double)__shfl_sync(0xffffffffu, id, 31);
    }
    if (fire && lane == 0) flag_rows[atomicAdd(&ctl->nflag, 1u)] = row;
  }
}

// winner of the search from the packed key (pruned path; replaces k_argmax_rows)
__global__ void k_finalize_best(const SearchCtl* __restrict__ ctl, const RowBest* __restrict__ rows, int row_base, BestSplit* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  BestSplit b;
  memset(&b, 0, sizeof(b));
  const unsigned long long k = ctl->key;
  if (k != 0ull) {
    const int r = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
    b.q = rows[r].q; b.valid = b.q > 0.f; b.var = row_base + r; b.c = rows[r].c; b.vleft = rows[r].vleft; b.pos = rows[r].pos;
  } else { b.q = -1.f; b.valid = 0; }
  b.pad[0] = (int)ctl->nflag;
  *out = b;
}

// ------------------------------------------------------------------------------------------------
// Split search, categorical variables (LBP codes): CvBoostTree::find_split_cat_reg / _cat_class
// (src/o_cvboostree.cpp:428-516, :249-359).  One warp per variable.  The 256 per-category sums are
// accumulated in node order (ascending sample index) exactly like the reference: for every sample
// the lane owning bin (code & 31) performs the two f64 additions, so the order of additions inside a
// bin is the sequential one.  Categories are then ranked by (key, category) — the stable order —
// and the prefix scan over the ranked categories is done sequentially by lane 0 with un-fused f64
// operations.
// ------------------------------------------------------------------------------------------------
constexpr int CAT_THREADS_MAX = 1024;      // the launch picks as many warps as shared memory allows (per-warp state: 10.5 KB)

__global__ void __launch_bounds__(CAT_THREADS_MAX, 1)
k_split_search_cat(const uint8_t* __restrict__ codes, int n_stride, int N, int nrows, const double* __restrict__ wnode,
                   const float* __restrict__ ynode, int is_classifier, int crit, int has_y, int ops_global, RowBestCat* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CAT_THREADS = blockDim.x, CAT_WARPS = blockDim.x >> 5;
  // per-sample operands: staged in shared memory, or -- when numSamples is too large for that -- read from global memory
  // (they are read linearly, one coalesced line per 32 samples, and stay in L1 / L2 across the rows of a CTA)
  const int NS = ops_global ? 0 : N;
  double* sw_s = reinterpret_cast<double*>(smem_raw);                 // [NS]
  float* sy_s = reinterpret_cast<float*>(sw_s + NS);                   // [NS] (LogitBoost only)
  double* bins = reinterpret_cast<double*>(sy_s + (has_y ? ((NS + 1) & ~1) : 0));   // [CAT_WARPS][3][256]: s0, s1, key
  const double* sw = ops_global ? wnode : sw_s;
  const float* sy = ops_global ? ynode : sy_s;
  int* order_all = reinterpret_cast<int*>(bins + CAT_WARPS * 3 * 256); // [CAT_WARPS][256]
  double* steps_all = reinterpret_cast<double*>(order_all + CAT_WARPS * 256);   // [CAT_WARPS][6][32]
  double* s_ikey_all = steps_all + CAT_WARPS * 6 * 32;                          // [CAT_WARPS][256] u64
  for (int i = threadIdx.x; i < NS; i += CAT_THREADS) { sw_s[i] = wnode[i]; if (has_y) sy_s[i] = ynode[i]; }
  __syncthreads();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  double* s0 = bins + warp * 3 * 256;     // regression: sum of y*w ; classification: class-0 weight
  double* s1 = s0 + 256;                  // regression: sum of w   ; classification: class-1 weight
  double* key = s1 + 256;
  int* order = order_all + warp * 256;

  for (int row = blockIdx.x * CAT_WARPS + warp; row < nrows; row += gridDim.x * CAT_WARPS) {
    const uint8_t* rc = codes + (size_t)row * n_stride;
    for (int i = lane; i < 256; i += 32) { s0[i] = 0.0; s1[i] = 0.0; }
    __syncwarp();
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      const int mycode = i < N ? rc[i] : 0;
      const double myw = i < N ? sw[i] : 0.0;
      const float myy = (has_y && i < N) ? sy[i] : 0.f;
      // route every node sample of this group of 32 to the lane that owns its bin ((code & 31) == lane): owner masks from
      // 5 ballots; every owner then walks its samples in ascending sample order, so the additions inside a bin happen in
      // the reference's sequential order while different bins proceed in parallel (about 3 rounds instead of 32)
      unsigned mine = __ballot_sync(0xffffffffu, myw != 0.0);
#pragma unroll
      for (int bit = 0; bit < 5; ++bit) {
        const unsigned b = __ballot_sync(0xffffffffu, (mycode >> bit) & 1);
        mine &= ((lane >> bit) & 1) ? b : ~b;
      }
      while (__any_sync(0xffffffffu, mine != 0u)) {
        const int src = mine ? __ffs(mine) - 1 : lane;
        const int code = __shfl_sync(0xffffffffu, mycode, src);
        const double w = __shfl_sync(0xffffffffu, myw, src);
        const float y = has_y ? __shfl_sync(0xffffffffu, myy, src) : 0.f;
        if (mine) {
          mine &= mine - 1u;
          if (is_classifier) {
            if (w > 0.0) s1[code] = __dadd_rn(s1[code], w); else s0[code] = __dadd_rn(s0[code], -w);
          } else if (has_y) {
            s0[code] = __dadd_rn(s0[code], __dmul_rn((double)y, w));
            s1[code] = __dadd_rn(s1[code], w);
          } else {
            s0[code] = __dadd_rn(s0[code], w);            // w already carries the +-1 target
            s1[code] = __dadd_rn(s1[code], fabs(w));
          }
        }
      }
    }
    __syncwarp();
    // totals: sequential over the categories like the reference (o_cvboostree.cpp:459-467, every lane redundantly, same
    // order); sort keys: one division per category, 8 categories per lane
    double R = 0, rsum = 0, rc0 = 0, rc1 = 0;
    if (!is_classifier) {
      for (int i = 0; i < 256; ++i) { R = __dadd_rn(R, s1[i]); rsum = __dadd_rn(rsum, s0[i]); }
      for (int i = lane; i < 256; i += 32) key[i] = fabs(s1[i]) > DBL_EPSILON ? __ddiv_rn(s0[i], s1[i]) : 0.0;
    } else {
      for (int i = 0; i < 256; ++i) { rc0 = __dadd_rn(rc0, s0[i]); rc1 = __dadd_rn(rc1, s1[i]); }
      for (int i = lane; i < 256; i += 32) key[i] = s1[i];
    }
    __syncwarp();
    // stable rank by counting, on order-preserving 64-bit integer images of the keys (the FP64 pipe is narrow; 65536
    // compares per variable belong on the integer pipe).  
    unsigned long long* ikey = reinterpret_cast<unsigned long long*>(s_ikey_all) + warp * 256;
    for (int i = lane; i < 256; i += 32) {
      const unsigned long long u = (unsigned long long)__double_as_longlong(key[i] + 0.0);      // -0 -> +0
      ikey[i] = u ^ ((u >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
    }
    __syncwarp();
    for (int j = lane; j < 256; j += 32) {
      const unsigned long long kj = ikey[j];
      int rank = 0;
#pragma unroll 8
      for (int k = 0; k < 256; ++k) { const unsigned long long kk = ikey[k]; rank += (kk < kj) || (kk == kj && k < j); }
      order[rank] = j;
    }
    __syncwarp();
    // Prefix scan over the ranked categories.  The running sums are a sequential chain (lane 0, the reference's order of
    // un-fused f64 operations, 32 positions at a time); the criterion at a position only reads them, so all lanes
    // evaluate one position each and keep their first maximum -- same values and same winner as the sequential loop.
    double* stp = steps_all + warp * 6 * 32;        // [6][32]: per-position running sums; stp[0][k] < 0 marks "no candidate here"
    double best_val = -1.0;
    int best_subset = -1;
    double L = 0, lsum = 0, lcw0 = 0, lcw1 = 0;
    if (is_classifier) R = __dadd_rn(rc0, rc1);
    for (int blk = 0; blk < 255; blk += 32) {
      if (lane == 0) {
        const int lim = min(32, 255 - blk);
        for (int k = 0; k < lim; ++k) {
          const int idx = order[blk + k];
          bool valid = false;
          if (!is_classifier) {
            const double ni = s1[idx];
            if (ni > FLT_EPSILON) {
              const double sv = __dmul_rn(key[idx], ni);       // "revert back to unnormalized sums"
              lsum = __dadd_rn(lsum, sv); L = __dadd_rn(L, ni);
              rsum = __dsub_rn(rsum, sv); R = __dsub_rn(R, ni);
              valid = L > FLT_EPSILON && R > FLT_EPSILON;
            }
            stp[0 * 32 + k] = valid ? L : -1.0; stp[1 * 32 + k] = R; stp[2 * 32 + k] = lsum; stp[3 * 32 + k] = rsum;
          } else {
            const double w0 = s0[idx], w1 = s1[idx];
            const double weight = __dadd_rn(w0, w1);
            if (!(weight < FLT_EPSILON)) {
              lcw0 = __dadd_rn(lcw0, w0); rc0 = __dsub_rn(rc0, w0); lcw1 = __dadd_rn(lcw1, w1); rc1 = __dsub_rn(rc1, w1);
              if (crit == CRIT_GINI) {
                L = __dadd_rn(L, weight); R = __dsub_rn(R, weight);
                valid = L > FLT_EPSILON && R > FLT_EPSILON;
              } else valid = true;
            }
            stp[0 * 32 + k] = valid ? (crit == CRIT_GINI ? L : 1.0) : -1.0; stp[1 * 32 + k] = R;
            stp[2 * 32 + k] = lcw0; stp[3 * 32 + k] = lcw1; stp[4 * 32 + k] = rc0; stp[5 * 32 + k] = rc1;
          }
        }
        for (int k = lim; k < 32; ++k) stp[k] = -1.0;
      }
      __syncwarp();
      const double a0 = stp[lane];
      if (a0 > 0.0) {
        double val;
        if (!is_classifier) {
          const double Rr = stp[32 + lane], ls = stp[64 + lane], rs = stp[96 + lane];
          const double num = __dadd_rn(__dmul_rn(__dmul_rn(ls, ls), Rr), __dmul_rn(__dmul_rn(rs, rs), a0));
          val = __ddiv_rn(num, __dmul_rn(a0, Rr));
        } else {
          const double l0 = stp[64 + lane], l1 = stp[96 + lane], r0 = stp[128 + lane], r1 = stp[160 + lane];
          if (crit == CRIT_GINI) {
            const double Rr = stp[32 + lane];
            const double lsum2 = __dadd_rn(__dmul_rn(l0, l0), __dmul_rn(l1, l1));
            const double rsum2 = __dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1));
            val = __ddiv_rn(__dadd_rn(__dmul_rn(lsum2, Rr), __dmul_rn(rsum2, a0)), __dmul_rn(a0, Rr));
          } else val = fmax(__dadd_rn(l0, r1), __dadd_rn(l1, r0));
        }
        if (best_val < val) { best_val = val; best_subset = blk + lane; }     // positions ascend per lane: first maximum
      }
      __syncwarp();
    }
    // warp argmax: larger value, then smaller position (the sequential loop keeps the first maximum)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best_val, d);
      const int os = __shfl_xor_sync(0xffffffffu, best_subset, d);
      if (os >= 0 && (best_subset < 0 || ov > best_val || (ov == best_val && os < best_subset))) { best_val = ov; best_subset = os; }
    }
    if (lane == 0) {
      RowBestCat rb;
      rb.q = best_subset >= 0 ? (float)best_val : -1.f;
      for (int k = 0; k < 8; ++k) rb.subset[k] = 0;
      for (int i = 0; i <= best_subset; ++i) { const int idx = order[i]; rb.subset[idx >> 5] |= 1 << (idx & 31); }
      out[row] = rb;
    }
    __syncwarp();
  }
}

// ---- CvDTree::find_best_split over the local rows (src/o_cvdtree.cpp:313-357): highest f32 quality,
//      lowest variable index on ties, accepted only if quality > 0.  Single CTA. ---------------------
__global__ void __launch_bounds__(ONE_CTA, 1)
k_argmax_rows(const RowBest* __restrict__ rows, const RowBestCat* __restrict__ rows_cat, int nrows, int row_base, BestSplit* out) {
  __shared__ float sq[32];
  __shared__ int sr[32];
  float q = -1.f;
  int r = 0x7fffffff;
  for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
    const float qi = rows ? rows[i].q : rows_cat[i].q;
    if (qi > q) { q = qi; r = i; }    // i increases within a thread, so ties keep the lower index
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float oq = __shfl_xor_sync(0xffffffffu, q, d);
    const int orr = __shfl_xor_sync(0xffffffffu, r, d);
    if (oq > q || (oq == q && orr < r)) { q = oq; r = orr; }
  }
  if (lane == 0) { sq[warp] = q; sr[warp] = r; }
  __syncthreads();
  if (warp == 0) {
    q = sq[lane]; r = sr[lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const float oq = __shfl_xor_sync(0xffffffffu, q, d);
      const int orr = __shfl_xor_sync(0xffffffffu, r, d);
      if (oq > q || (oq == q && orr < r)) { q = oq; r = orr; }
    }
    if (lane == 0) {
      BestSplit b;
      memset(&b, 0, sizeof(b));
      b.q = q;
      b.valid = (q > 0.f) && r != 0x7fffffff;
      if (r != 0x7fffffff) {
        b.var = row_base + r;
        if (rows) { b.c = rows[r].c; b.vleft = rows[r].vleft; b.pos = rows[r].pos; }
        else for (int k = 0; k < 8; ++k) b.subset[k] = rows_cat[r].subset[k];
      }
      *out = b;
    }
  }
}

// winner among the per-rank candidates (feature sharding): same ordering rule, variable indices are global
__global__ void k_argmax_candidates(const BestSplit* __restrict__ cands, int n, BestSplit* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  BestSplit best = cands[0];
  for (int i = 1; i < n; ++i) {
    const BestSplit c = cands[i];
    if (!c.valid) continue;
    if (!best.valid || c.q > best.q || (c.q == best.q && c.var < best.var)) best = c;
  }
  *out = best;
}

// ---- calc_node_dir + split_node_data (src/o_cvboostree.cpp:87-149, src/o_cvcascadeboosttree.cpp:167-347):
//      node samples left of the split point are exactly those with value <= v[split_point], because a
//      split never separates equal values; categorical: subset bit set -> left. -------------------------
__global__ void k_apply_split(EvalView e, int N, int nid, int lid, int rid, const BestSplit* __restrict__ best, int* node_of) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N || node_of[s] != nid) return;
  const BestSplit b = *best;
  if (!b.valid) return;                    // speculative launch of the stump path: no split was found
  const float val = d_eval_var(e, b.var, s);
  bool left;
  if (e.ftype == CC_FEATURE_LBP) { const int c = (int)val; left = (b.subset[c >> 5] & (1 << (c & 31))) != 0; }
  else left = val <= b.vleft;
  node_of[s] = left ? lid : rid;
}

// ---- tree prediction for one sample, CvCascadeBoostTree::predict (src/o_cvcascadeboosttree.cpp:16-39)
__device__ __forceinline__ double d_tree_predict(const EvalView& e, const DevNode* __restrict__ tree, int s) {
  int n = 0;
  while (tree[n].left >= 0) {
    const float val = d_eval_var(e, tree[n].var_idx, s);
    bool left;
    if (e.ftype == CC_FEATURE_LBP) { const int c = (int)val; left = (tree[n].subset[c >> 5] & (1 << (c & 31))) != 0; }
    else left = val <= tree[n].c;
    n = left ? tree[n].left : tree[n].right;
  }
  return tree[n].value;
}

__device__ __forceinline__ uint32_t f2key(float v) {
  uint32_t u = __float_as_uint(v);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }

// ------------------------------------------------------------------------------------------------
// Everything that follows a finished weak tree, in one single-CTA kernel:
//   update_weights(tree)  src/boost.cpp:266-406     trim_weights  src/o_cvboost.cpp:101-139
//   isErrDesired          src/boost.cpp:479-518
// The two order statistics (trim threshold = first sorted weight at which the running remainder
// 1-rate-sum drops to <= 0; stage threshold = eval[thresholdIdx] of the sorted positive sums) are
// found by bisection over the value bits instead of by sorting.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ONE_CTA, 1)
k_tree_finish(EvalView e, int N, int boost_type, int have_subsample, double trim_rate, float min_hit_rate,
              const DevNode* __restrict__ tree, const uint8_t* __restrict__ labels, const int* __restrict__ orig_response,
              double* weights, double* weak_eval, double* treeval, double* stage_sum, double* sum_response, float* responses,
              uint8_t* mask, TreeSummary* summary, const int* __restrict__ tree_ok) {
  __shared__ double scratch[32], scratch2[32], scratch3[32];
  __shared__ unsigned long long s_u64;
  if (tree_ok && !*tree_ok) return;        // stump path: the root was not split (tree->train() failed, src/boost.cpp:436-440): nothing is updated
  // (a) responses of the samples that did not take part in the tree, and the threshold-rule output for everyone
  for (int s = threadIdx.x; s < N; s += blockDim.x) {
    const double tv = d_tree_predict(e, tree, s);
    treeval[s] = tv;
    if (have_subsample && !mask[s]) weak_eval[s] = tv;
  }
  __syncthreads();
  // (b) weight update
  double scaleC = 1.0;
  double sumW = 0;
  if (boost_type == CC_BOOST_DISCRETE) {
    double sw = 0, err = 0;
    for (int s = threadIdx.x; s < N; s += blockDim.x) {
      const double w = weights[s];
      sw += w;
      err += w * (weak_eval[s] != (double)orig_response[s]);
    }
    sw = block_sum(sw, scratch); err = block_sum(err, scratch);
    if (sw != 0) err /= sw;
    const double C = -d_log_ratio(err);
    scaleC = C;
    const double sc1 = exp(C);
    double acc = 0;
    for (int s = threadIdx.x; s < N; s += blockDim.x) {
      const double w = weights[s] * ((weak_eval[s] != (double)orig_response[s]) ? sc1 : 1.0);
      acc += w;
      weights[s] = w;
    }
    sumW = block_sum(acc, scratch);
  } else if (boost_type == CC_BOOST_LOGIT) {
    double acc = 0;
    for (int s = threadIdx.x; s < N; s += blockDim.x) {
      const double sr = sum_response[s] + 0.5 * weak_eval[s];
      sum_response[s] = sr;
      const double ex = exp(-2 * sr);
      const double p = 1. / (1. + ex);
      double w = p * (1 - p);
      w = fmax(w, (double)FLT_EPSILON);
      weights[s] = w;
      acc += w;
      if (orig_response[s] > 0) { const double z = 1. / p; responses[s] = (float)fmin(z, 10.); }
      else { const double z = 1. / (1 - p); responses[s] = (float)-fmin(z, 10.); }
    }
    sumW = block_sum(acc, scratch);
  } else {   // REAL and GENTLE share the update w *= exp(-y f)
    double acc = 0;
    for (int s = threadIdx.x; s < N; s += blockDim.x) {
      const double we = weak_eval[s] * (double)(-orig_response[s]);
      const double w = weights[s] * exp(we);
      weights[s] = w;
      acc += w;
    }
    sumW = block_sum(acc, scratch);
  }
  if (sumW > FLT_EPSILON) {
    const double inv = 1. / sumW;
    for (int s = threadIdx.x; s < N; s += blockDim.x) weights[s] *= inv;
  }
  __syncthreads();
  // (c) trim_weights
  int nz = N, hs = 0;
  if (trim_rate > 0. && trim_rate < 1.) {
    const double target = 1. - trim_rate;
    // smallest weight value v with  (sum of weights <= v) - v >= target ; DBL_MAX if none
    unsigned long long lo = 0, hi = 0x7FF0000000000000ull;     // bit patterns of non-negative doubles are ordered
    // predicate(v) true for hi (inf) trivially false if total - ... ; handle "none" below
    bool any = false;
    {
      double tot = 0, mx = 0;
      for (int s = threadIdx.x; s < N; s += blockDim.x) { tot += weights[s]; mx = fmax(mx, weights[s]); }
      tot = block_sum(tot, scratch);
      // block max
      for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      __syncthreads();
      if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = mx;
      __syncthreads();
      mx = scratch[threadIdx.x & 31];
      for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      any = (tot - mx) >= target;
      hi = (unsigned long long)__double_as_longlong(mx);
    }
    double threshold = DBL_MAX;
    if (any) {
      // Q(x) = [ sum{w <= x} - max{w <= x} >= target ] is a monotone step function of x whose first
      // true point is the wanted weight value (the sorted prefix before the last copy of that value).
      // Q only changes at actual weight values, so after testing x = mid both ends snap to weights: true -> hi = largest
      // weight <= mid, false -> lo = smallest weight > mid.  The interval then empties after ~log2(N) + a few steps
      // instead of the 63 halvings of the bit space.
      while (lo < hi) {
        const unsigned long long mid = lo + ((hi - lo) >> 1);
        const double v = __longlong_as_double((long long)mid);
        double acc = 0, top = 0, nxt = DBL_MAX;
        for (int s = threadIdx.x; s < N; s += blockDim.x) {
          const double w = weights[s];
          if (w <= v) { acc += w; top = fmax(top, w); } else nxt = fmin(nxt, w);
        }
        acc = warp_sum(acc);
        for (int d = 16; d > 0; d >>= 1) { top = fmax(top, __shfl_xor_sync(0xffffffffu, top, d)); nxt = fmin(nxt, __shfl_xor_sync(0xffffffffu, nxt, d)); }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { scratch[threadIdx.x >> 5] = acc; scratch2[threadIdx.x >> 5] = top; scratch3[threadIdx.x >> 5] = nxt; }
        __syncthreads();
        acc = scratch[threadIdx.x & 31]; top = scratch2[threadIdx.x & 31]; nxt = scratch3[threadIdx.x & 31];
        acc = warp_sum(acc);
        for (int d = 16; d > 0; d >>= 1) { top = fmax(top, __shfl_xor_sync(0xffffffffu, top, d)); nxt = fmin(nxt, __shfl_xor_sync(0xffffffffu, nxt, d)); }
        if (acc - top >= target) hi = (unsigned long long)__double_as_longlong(top);
        else lo = (unsigned long long)__double_as_longlong(nxt);
      }
      // snap to the smallest actual weight >= bound
      const double bound = __longlong_as_double((long long)lo);
      double mn = DBL_MAX;
      for (int s = threadIdx.x; s < N; s += blockDim.x) { const double w = weights[s]; if (w >= bound) mn = fmin(mn, w); }
      for (int d = 16; d > 0; d >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      __syncthreads();
      if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = mn;
      __syncthreads();
      mn = scratch[threadIdx.x & 31];
      for (int d = 16; d > 0; d >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      threshold = mn;
    }
    long long cnt = 0;
    for (int s = threadIdx.x; s < N; s += blockDim.x) {
      const int f = weights[s] >= threshold;
      mask[s] = (uint8_t)f;
      cnt += f;
    }
    nz = (int)block_sum_ll(cnt, scratch);
    hs = nz < N;
  } else {
    // mask is left as it is (all ones: have_subsample can never become true)
    hs = 0;
  }
  // (d) stage sums and isErrDesired
  long long npos = 0;
  for (int s = threadIdx.x; s < N; s += blockDim.x) {
    stage_sum[s] += treeval[s] * scaleC;
    npos += labels[s] == 1;
  }
  const int numPos = (int)block_sum_ll(npos, scratch);
  const int numNeg = N - numPos;
  const int thresholdIdx = (int)((1.0F - min_hit_rate) * numPos);
  // k-th smallest (k = thresholdIdx) of the float stage sums of the positives
  uint32_t klo = 0, khi = 0xFFFFFFFFu;
  while (klo < khi) {       // same snapping: the count only changes at actual keys
    const uint32_t mid = klo + ((khi - klo) >> 1);
    long long c = 0;
    uint32_t top = 0u, nxt = 0xFFFFFFFFu;
    for (int s = threadIdx.x; s < N; s += blockDim.x)
      if (labels[s] == 1) {
        const uint32_t k = f2key((float)stage_sum[s]);
        if (k <= mid) { c++; top = max(top, k); } else nxt = min(nxt, k);
      }
    for (int d = 16; d > 0; d >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, d);
      top = max(top, __shfl_xor_sync(0xffffffffu, top, d)); nxt = min(nxt, __shfl_xor_sync(0xffffffffu, nxt, d));
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { scratch[threadIdx.x >> 5] = (double)c; scratch2[threadIdx.x >> 5] = (double)top; scratch3[threadIdx.x >> 5] = (double)nxt; }
    __syncthreads();
    double dc = scratch[threadIdx.x & 31], dt = scratch2[threadIdx.x & 31], dn = scratch3[threadIdx.x & 31];      // counts and 32-bit keys are exact in f64
    dc = warp_sum(dc);
    for (int d = 16; d > 0; d >>= 1) { dt = fmax(dt, __shfl_xor_sync(0xffffffffu, dt, d)); dn = fmin(dn, __shfl_xor_sync(0xffffffffu, dn, d)); }
    if ((long long)dc >= thresholdIdx + 1) khi = (uint32_t)dt; else klo = (uint32_t)dn;
  }
  const float thr = key2f(klo);
  long long below = 0, nfalse = 0;
  const float thr_eps = thr - 0.00001F;
  for (int s = threadIdx.x; s < N; s += blockDim.x) {
    const double sum = stage_sum[s];
    if (labels[s] == 1) {
      const float ev = (float)sum;
      if (ev < thr && !(fabsf(ev - thr) < FLT_EPSILON)) below++;
    } else {
      if (!(sum < (double)thr_eps)) nfalse++;
    }
  }
  below = block_sum_ll(below, scratch);
  nfalse = block_sum_ll(nfalse, scratch);
  if (threadIdx.x == 0) {
    TreeSummary ts;
    ts.scaleC = scaleC; ts.sumW = sumW; ts.nz_count = nz; ts.have_subsample = hs;
    ts.threshold = thr;
    ts.num_pos = numPos; ts.num_neg = numNeg;
    ts.hit_rate = ((float)(numPos - (int)below)) / ((float)numPos);
    ts.false_alarm = ((float)nfalse) / ((float)numNeg);
    *summary = ts;
  }
  (void)s_u64;
}

// Stump path (max_depth == 1): the tree is assembled on the device so that k_tree_finish can follow without a host round trip.
// ok = the root may be split (CvDTree::try_split_node, src/o_cvdtree.cpp:122-187 -- same tests as grow_node()) and a split was found.
__global__ void k_build_stump(const NodeStats* __restrict__ stats, const BestSplit* __restrict__ best, int min_sample_count,
                              float regression_accuracy, int is_classifier, DevNode* __restrict__ tree, int* __restrict__ tree_ok) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const NodeStats st = stats[0];
  bool can_split = st.n > min_sample_count;
  if (can_split && is_classifier) { if (((st.cls_count[0] != 0) + (st.cls_count[1] != 0)) == 1) can_split = false; }
  else if (can_split) { if (sqrt(st.risk) / st.n < (double)regression_accuracy) can_split = false; }
  const BestSplit b = *best;
  DevNode root, l, r;
  root.left = 1; root.right = 2; root.var_idx = b.var; root.c = b.c; root.value = st.value;
  for (int i = 0; i < 8; ++i) { root.subset[i] = b.subset[i]; l.subset[i] = 0; r.subset[i] = 0; }
  l.left = l.right = -1; l.var_idx = -1; l.c = 0.f; l.value = stats[1].value;
  r.left = r.right = -1; r.var_idx = -1; r.c = 0.f; r.value = stats[2].value;
  tree[0] = root; tree[1] = l; tree[2] = r;
  *tree_ok = (can_split && b.valid) ? 1 : 0;
}

__global__ void k_row_quality(const RowBest* rows, const RowBestCat* rows_cat, int nrows, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) { const float q = rows ? rows[i].q : rows_cat[i].q; out[i] = q > 0.f ? q : 0.f; }
}

// stage sums for arbitrary stages (cc_stage_predict)
__global__ void k_stage_predict(EvalView e, const DevNode* __restrict__ nodes, const int* __restrict__ tree_first, int ntrees, int s0,
                                int ns, float threshold, int return_sum, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  double sum = 0;
  for (int t = 0; t < ntrees; ++t) sum += d_tree_predict(e, nodes + tree_first[t], s0 + i);
  if (!return_sum) sum = sum < (double)(threshold - 0.00001F) ? 0.0 : 1.0;
  out[i] = (float)sum;
}

// ---- host side ----------------------------------------------------------------------------------------
static int launch_split_search(cc_ctx* ctx, BoostBuffers* b, int nid) {
  const int N = ctx->N;
  ScopedTimer tm(ctx, "split_search", 2);
  if (ctx->ftype == CC_FEATURE_LBP) {
    const bool cat_has_y = b->boost_type == CC_BOOST_LOGIT;
    size_t smem_fixed = (size_t)N * 8 + (cat_has_y ? (size_t)((N + 1) & ~1) * 4 : 0);
    const size_t smem_warp = (size_t)3 * 256 * 8 + 256 * 4 + (size_t)6 * 32 * 8 + (size_t)256 * 8;
    // operands in shared memory as long as at least 8 warps still fit next to them; beyond that (numSamples > ~17000) they are
    // read from global memory
    const int ops_global = (smem_fixed + 8 * smem_warp > 227 * 1024 || b->search_mode == 2) ? 1 : 0;
    if (ops_global) smem_fixed = 0;
    const int CAT_WARPS = (int)std::min<size_t>(CAT_THREADS_MAX / 32, (227 * 1024 - smem_fixed) / smem_warp);
    const int CAT_THREADS = CAT_WARPS * 32;
    const size_t smem = smem_fixed + CAT_WARPS * smem_warp;
    CC_CUDA(cudaFuncSetAttribute(k_split_search_cat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = std::min((ctx->nrows + CAT_WARPS - 1) / CAT_WARPS, ctx->sm_count);
    k_split_search_cat<<<blocks, CAT_THREADS, smem, ctx->stream>>>(ctx->d_fcode, ctx->n_stride, N, ctx->nrows, b->wnode, b->ynode,
                                                                   b->is_classifier, b->crit, b->boost_type == CC_BOOST_LOGIT, ops_global, b->rowbest_cat);
    CC_CUDA(cudaGetLastError());
    k_argmax_rows<<<1, ONE_CTA, 0, ctx->stream>>>(nullptr, b->rowbest_cat, ctx->nrows, ctx->row0, b->d_best);
  } else if (ctx->wide) {
    const bool has_y = b->boost_type == CC_BOOST_LOGIT;
    const int blocks = std::min((ctx->nrows + SW_THREADS / 32 - 1) / (SW_THREADS / 32), ctx->sm_count * 8);
#define LAUNCH_SW(CR, HY) k_split_search_ord_wide<CR, HY><<<blocks, SW_THREADS, 0, ctx->stream>>>(ctx->d_sidx32, ctx->d_fval, ctx->n_stride, N, ctx->nrows, \
                                                                                                  b->wnode, b->ynode, b->d_stats + nid, b->rowbest);
    if (b->crit == CRIT_SQERR) { if (has_y) LAUNCH_SW(CRIT_SQERR, true) else LAUNCH_SW(CRIT_SQERR, false) }
    else if (b->crit == CRIT_GINI) LAUNCH_SW(CRIT_GINI, false)
    else LAUNCH_SW(CRIT_MISCLASS, false)
#undef LAUNCH_SW
    CC_CUDA(cudaGetLastError());
    k_argmax_rows<<<1, ONE_CTA, 0, ctx->stream>>>(b->rowbest, nullptr, ctx->nrows, ctx->row0, b->d_best);
    b->row_quality_valid = true;
  } else {
    const bool has_y = b->boost_type == CC_BOOST_LOGIT;
    const int ntile = (N + SS_TILE - 1) / SS_TILE;
    const size_t smem = (size_t)(N + 1) * 8 + (size_t)SS_WARPS * ntile * (sizeof(TileCarry) + 4) + (has_y ? (size_t)(N + 1) * 4 : 0) + 16;
    CC_REQUIRE(smem <= 227 * 1024, CC_ERR_UNSUPPORTED, "split search: numSamples %d needs %zu B of shared memory (limit 227 KB)", N, smem);
    const int per_sm = std::max(1, std::min(4, (int)((220 * 1024) / (smem + 1024))));
    const double ymax2 = has_y ? 100.0 : 1.0;      // slack of the f32 pre-filter, see k_split_search_ord
    const int blocks = std::min((ctx->nrows + SS_WARPS - 1) / SS_WARPS, ctx->sm_count * per_sm);
    const bool pruned = b->search_mode == 0 && !has_y && ctx->nrows >= 4 * SEED_TARGET;
    CC_CUDA(cudaMemsetAsync(b->d_ctl, 0, sizeof(SearchCtl), ctx->stream));
#define LAUNCH_SS(CR, HY, BLOCKS, NWORK, LIST, COUNT)                                                                       \
  {                                                                                                                         \
    CC_CUDA(cudaFuncSetAttribute(k_split_search_ord<CR, HY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    k_split_search_ord<CR, HY><<<BLOCKS, SS_THREADS, smem, ctx->stream>>>(ctx->d_sidx, ctx->d_sval, ctx->n_stride, N, NWORK, \
                                                                          b->wnode, b->ynode, b->d_stats + nid, ymax2, b->rowbest, \
                                                                          LIST, COUNT, b->d_ctl);                           \
  }
#define LAUNCH_SS_CRIT(BLOCKS, NWORK, LIST, COUNT)                                                             \
  {                                                                                                            \
    if (b->crit == CRIT_SQERR) { if (has_y) LAUNCH_SS(CRIT_SQERR, true, BLOCKS, NWORK, LIST, COUNT) else LAUNCH_SS(CRIT_SQERR, false, BLOCKS, NWORK, LIST, COUNT) } \
    else if (b->crit == CRIT_GINI) LAUNCH_SS(CRIT_GINI, false, BLOCKS, NWORK, LIST, COUNT)                     \
    else LAUNCH_SS(CRIT_MISCLASS, false, BLOCKS, NWORK, LIST, COUNT)                                           \
    CC_CUDA(cudaGetLastError());                                                                               \
  }
    // one CTA per row (k_split_search_ord_coop) for the short row lists of the pruned search
    const size_t csmem = (size_t)(N + 1) * 8 + (size_t)ntile * (2 * sizeof(TileCarry) + 4) + (has_y ? (size_t)(N + 1) * 4 : 0) + 16;
#define LAUNCH_SC(CR, HY, BLOCKS, NWORK, LIST, COUNT)                                                                            \
  {                                                                                                                              \
    CC_CUDA(cudaFuncSetAttribute(k_split_search_ord_coop<CR, HY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));     \
    k_split_search_ord_coop<CR, HY><<<BLOCKS, SS_THREADS, csmem, ctx->stream>>>(ctx->d_sidx, ctx->d_sval, ctx->n_stride, N, NWORK, \
                                                                                b->wnode, b->ynode, b->d_stats + nid, ymax2, b->rowbest, \
                                                                                LIST, COUNT, b->d_ctl);                           \
  }
#define LAUNCH_SC_CRIT(BLOCKS, NWORK, LIST, COUNT)                                                             \
  {                                                                                                            \
    if (b->crit == CRIT_SQERR) { if (has_y) LAUNCH_SC(CRIT_SQERR, true, BLOCKS, NWORK, LIST, COUNT) else LAUNCH_SC(CRIT_SQERR, false, BLOCKS, NWORK, LIST, COUNT) } \
    else if (b->crit == CRIT_GINI) LAUNCH_SC(CRIT_GINI, false, BLOCKS, NWORK, LIST, COUNT)                     \
    else LAUNCH_SC(CRIT_MISCLASS, false, BLOCKS, NWORK, LIST, COUNT)                                           \
    CC_CUDA(cudaGetLastError());                                                                               \
  }
    if (!pruned) {
      LAUNCH_SS_CRIT(blocks, ctx->nrows, (const int*)nullptr, (const unsigned int*)nullptr)
      k_argmax_rows<<<1, ONE_CTA, 0, ctx->stream>>>(b->rowbest, nullptr, ctx->nrows, ctx->row0, b->d_best);
      b->row_quality_valid = true;
    } else {
      // (1) exact search over the seed rows -> pruning threshold, (2) bound pass over every row, (3) exact search
      // over the flagged rows, (4) winner from the packed key
      const int nseed = (int)b->seed_rows.size();
      const int coop_blocks = ctx->sm_count * std::max(1, std::min(2, (int)((220 * 1024) / (csmem + 1024))));
      if (ctx->shard_world > 1 && nseed <= 6 * ctx->sm_count) {      // a feature shard searches SEED_TARGET / world seed rows: a few per SM, one CTA each
        LAUNCH_SC_CRIT(std::min(nseed, coop_blocks), nseed, b->d_seed_rows, (const unsigned int*)nullptr)
      } else {
        const int seed_blocks = std::min((nseed + SS_WARPS - 1) / SS_WARPS, ctx->sm_count * per_sm);
        LAUNCH_SS_CRIT(seed_blocks, nseed, b->d_seed_rows, (const unsigned int*)nullptr)
      }
      const size_t bsmem = (size_t)(N + 1) * sizeof(float);
      const int bper_sm = std::max(1, std::min(3, (int)((220 * 1024) / (bsmem + 1024))));
      const int bblocks = std::min((ctx->nrows + SB_WARPS - 1) / SB_WARPS, ctx->sm_count * bper_sm);
      if (b->crit == CRIT_MISCLASS) {
        CC_CUDA(cudaFuncSetAttribute(k_ss_bound<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
        k_ss_bound<true><<<bblocks, SB_THREADS, bsmem, ctx->stream>>>(ctx->d_sidx, ctx->n_stride, N, ctx->nrows, b->wnode, b->d_stats + nid, b->crit,
                                                                     b->d_ctl, b->d_flag_rows);
      } else {
        CC_CUDA(cudaFuncSetAttribute(k_ss_bound<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
        k_ss_bound<false><<<bblocks, SB_THREADS, bsmem, ctx->stream>>>(ctx->d_sidx, ctx->n_stride, N, ctx->nrows, b->wnode, b->d_stats + nid, b->crit,
                                                                      b->d_ctl, b->d_flag_rows);
      }
      CC_CUDA(cudaGetLastError());
      // The number of flagged rows is only known on the device; the last search read back predicts it.  A few rows per SM:
      // one CTA per row (~25 us per row); more (or nothing known yet): one warp per row (~85 us per wave of 16 rows per SM).
      if (b->last_nflag >= 0 && b->last_nflag <= 3 * ctx->sm_count) {
        LAUNCH_SC_CRIT(std::min(ctx->nrows, coop_blocks), ctx->nrows, b->d_flag_rows, &b->d_ctl->nflag)
      } else {
        LAUNCH_SS_CRIT(blocks, ctx->nrows, b->d_flag_rows, &b->d_ctl->nflag)
      }
      k_finalize_best<<<1, 32, 0, ctx->stream>>>(b->d_ctl, b->rowbest, ctx->row0, b->d_best);
      b->row_quality_valid = false;
      b->stat_searches++; b->stat_seed_rows += nseed;
      ctx->launches += 3;
    }
#undef LAUNCH_SC_CRIT
#undef LAUNCH_SC
#undef LAUNCH_SS_CRIT
#undef LAUNCH_SS
  }
  CC_CUDA(cudaGetLastError());
  if (ctx->ftype == CC_FEATURE_LBP) b->row_quality_valid = true;
  return CC_OK;
}

int exchange_candidates(cc_ctx* ctx, void* d_best_buf);   // cc_comm.cu: all-gather + argmax when sharded
static_assert(sizeof(BestSplit) == 64, "candidate record is exchanged as 64 raw bytes (cc_comm.cu)");
int launch_argmax_candidates(cc_ctx* ctx, const void* cands, int n, void* out) {
  k_argmax_candidates<<<1, 32, 0, ctx->stream>>>((const BestSplit*)cands, n, (BestSplit*)out);
  CC_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return CC_OK;
}

struct HostNode { int left = -1, right = -1, depth = 0, var = -1, pos = -1, n = 0; float c = 0, q = 0; int subset[8] = {0}; double value = 0; };

// CvDTree::try_split_node for one node, depth-first like the reference (the order of nodes does not
// affect results: nodes own disjoint samples)
static int grow_node(cc_ctx* ctx, BoostBuffers* b, std::vector<HostNode>& nodes, int nid) {
  const int N = ctx->N;
  {
    ScopedTimer tm(ctx, "boost_small", 1);
    k_node_prepare<<<1, ONE_CTA, 0, ctx->stream>>>(N, nid, b->is_classifier, b->boost_type, b->node_of, b->weights, b->responses,
                                                   b->wnode, b->ynode, b->weak_eval, b->d_stats);
    CC_CUDA(cudaGetLastError());
  }
  CC_CUDA(cudaMemcpyAsync(b->h_stats + nid, b->d_stats + nid, sizeof(NodeStats), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  const NodeStats st = b->h_stats[nid];
  nodes[nid].value = st.value;
  nodes[nid].n = st.n;
  bool can_split = true;
  if (st.n <= b->params.min_sample_count || nodes[nid].depth >= b->params.max_depth) can_split = false;
  if (can_split && b->is_classifier) {
    const int nz = (st.cls_count[0] != 0) + (st.cls_count[1] != 0);
    if (nz == 1) can_split = false;
  } else if (can_split) {
    if (std::sqrt(st.risk) / st.n < b->params.regression_accuracy) can_split = false;
  }
  if (!can_split) return CC_OK;
  CC_CHECK(launch_split_search(ctx, b, nid));
  CC_CHECK(exchange_candidates(ctx, b->d_best));
  CC_CUDA(cudaMemcpyAsync(b->h_best, b->d_best, sizeof(BestSplit), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  const BestSplit bs = *b->h_best;
  if (b->d_seed_rows && !b->row_quality_valid) { b->stat_flag_rows += bs.pad[0]; b->last_nflag = (int)bs.pad[0]; }
  if (!bs.valid) return CC_OK;
  if (b->d_seed_rows && bs.var >= ctx->row0 && bs.var < ctx->row0 + ctx->nrows && (int)b->seed_rows.size() < b->seed_cap) {
    // the winner of this search stays a strong candidate for the next ones: evaluate it exactly up front
    const int lr = bs.var - ctx->row0;
    if (std::find(b->seed_rows.begin(), b->seed_rows.end(), lr) == b->seed_rows.end()) {
      b->seed_rows.push_back(lr);
      CC_CUDA(cudaMemcpyAsync(b->d_seed_rows + b->seed_rows.size() - 1, &b->seed_rows.back(), sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  CC_REQUIRE((int)nodes.size() + 2 <= CC_MAX_TREE_NODES, CC_ERR_UNSUPPORTED, "tree grew beyond %d nodes (max_depth too large)", CC_MAX_TREE_NODES);
  const int lid = (int)nodes.size(), rid = lid + 1;
  HostNode l, r;
  l.depth = r.depth = nodes[nid].depth + 1;
  nodes.push_back(l); nodes.push_back(r);
  nodes[nid].left = lid; nodes[nid].right = rid;
  nodes[nid].var = bs.var; nodes[nid].c = bs.c; nodes[nid].q = bs.q; nodes[nid].pos = bs.pos;
  std::memcpy(nodes[nid].subset, bs.subset, sizeof(bs.subset));
  {
    ScopedTimer tm(ctx, "boost_small", 1);
    k_apply_split<<<(N + 255) / 256, 256, 0, ctx->stream>>>(make_eval_view(ctx), N, nid, lid, rid, b->d_best, b->node_of);
    CC_CUDA(cudaGetLastError());
  }
  CC_CHECK(grow_node(ctx, b, nodes, lid));
  CC_CHECK(grow_node(ctx, b, nodes, rid));
  return CC_OK;
}

// One weak tree of depth 1 without a host round trip inside: root statistics, split search (+ the exchange between the ranks),
// the split, both leaves, the tree record and everything that follows a finished tree (k_tree_finish) are queued back to back;
// the host reads statistics, winner and summary ONCE per tree.  The launches that depend on "can the root be split / was a split
// found" are speculative: k_apply_split ignores an invalid winner and k_tree_finish returns at once when the tree failed.
// Same results as grow_node() + the tail of cc_boost_train's loop (test_stump_path_equals_general_path).
static int grow_stump_and_finish(cc_ctx* ctx, BoostBuffers* b, const EvalView& ev, float min_hit_rate, std::vector<HostNode>& nodes, bool& ok) {
  const int N = ctx->N;
  {
    ScopedTimer tm(ctx, "boost_small", 1);
    k_node_prepare<<<1, ONE_CTA, 0, ctx->stream>>>(N, 0, b->is_classifier, b->boost_type, b->node_of, b->weights, b->responses,
                                                   b->wnode, b->ynode, b->weak_eval, b->d_stats);
    CC_CUDA(cudaGetLastError());
  }
  CC_CHECK(launch_split_search(ctx, b, 0));
  CC_CHECK(exchange_candidates(ctx, b->d_best));
  {
    ScopedTimer tm(ctx, "boost_small", 4);
    k_apply_split<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ev, N, 0, 1, 2, b->d_best, b->node_of);
    // the leaves: wnode / ynode are no longer needed (no further search), only the node values and weak_eval
    k_node_prepare<<<2, ONE_CTA, 0, ctx->stream>>>(N, 1, b->is_classifier, b->boost_type, b->node_of, b->weights, b->responses,
                                                   b->wnode, b->ynode, b->weak_eval, b->d_stats);
    k_build_stump<<<1, 32, 0, ctx->stream>>>(b->d_stats, b->d_best, b->params.min_sample_count, b->params.regression_accuracy, b->is_classifier,
                                             b->d_tree, b->d_tree_ok);
    k_tree_finish<<<1, ONE_CTA, 0, ctx->stream>>>(ev, N, b->boost_type, b->have_subsample, b->params.weight_trim_rate, min_hit_rate,
                                                  b->d_tree, ctx->d_labels, b->orig_response, b->weights, b->weak_eval, b->treeval,
                                                  b->stage_sum, b->sum_response, b->responses, b->mask, b->d_summary, b->d_tree_ok);
    CC_CUDA(cudaGetLastError());
  }
  CC_CUDA(cudaMemcpyAsync(b->h_stats, b->d_stats, sizeof(NodeStats) * 3, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaMemcpyAsync(b->h_best, b->d_best, sizeof(BestSplit), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaMemcpyAsync(b->h_tree_ok, b->d_tree_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaMemcpyAsync(b->h_summary, b->d_summary, sizeof(TreeSummary), cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  const BestSplit bs = *b->h_best;
  if (b->d_seed_rows && !b->row_quality_valid) { b->stat_flag_rows += bs.pad[0]; b->last_nflag = (int)bs.pad[0]; }
  nodes[0].value = b->h_stats[0].value; nodes[0].n = b->h_stats[0].n;
  ok = *b->h_tree_ok != 0;
  if (!ok) return CC_OK;
  if (b->d_seed_rows && bs.var >= ctx->row0 && bs.var < ctx->row0 + ctx->nrows && (int)b->seed_rows.size() < b->seed_cap) {
    const int lr = bs.var - ctx->row0;
    if (std::find(b->seed_rows.begin(), b->seed_rows.end(), lr) == b->seed_rows.end()) {
      b->seed_rows.push_back(lr);
      CC_CUDA(cudaMemcpyAsync(b->d_seed_rows + b->seed_rows.size() - 1, &b->seed_rows.back(), sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  HostNode l, r;
  l.depth = r.depth = 1;
  l.value = b->h_stats[1].value; l.n = b->h_stats[1].n;
  r.value = b->h_stats[2].value; r.n = b->h_stats[2].n;
  nodes.push_back(l); nodes.push_back(r);
  nodes[0].left = 1; nodes[0].right = 2;
  nodes[0].var = bs.var; nodes[0].c = bs.c; nodes[0].q = bs.q; nodes[0].pos = bs.pos;
  std::memcpy(nodes[0].subset, bs.subset, sizeof(bs.subset));
  return CC_OK;
}

}  // namespace cc

using namespace cc;

extern "C" {

int cc_boost_train(cc_ctx* ctx, int num_samples, const cc_boost_params* p, cc_stage* out) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_boost_train: evaluator not initialised");
  CC_REQUIRE(p && out && out->trees && out->log, CC_ERR_BAD_ARG, "cc_boost_train: null params / output arrays");
  CC_REQUIRE(p->boost_type >= CC_BOOST_DISCRETE && p->boost_type <= CC_BOOST_GENTLE, CC_ERR_BAD_ARG, "Unknown/unsupported boosting type");
  CC_REQUIRE(p->min_hit_rate > 0 && p->min_hit_rate < 1 && p->max_false_alarm > 0 && p->max_false_alarm < 1, CC_ERR_BAD_ARG,
             "cc_boost_train: minHitRate / maxFalseAlarm must be in (0,1) (src/boost.cpp:534-541)");
  CC_REQUIRE(p->max_depth >= 1 && p->max_depth <= 5, CC_ERR_UNSUPPORTED, "cc_boost_train: max_depth %d outside [1,5]", p->max_depth);
  CC_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->precalc_valid || ctx->N != num_samples) CC_CHECK(cc_boost_precalc(ctx, num_samples));
  const int N = ctx->N;
  CC_CHECK(ensure_boost(ctx, N, (size_t)ctx->nrows));
  BoostBuffers* b = ctx->boost;
  // CvBoost::set_params, src/o_cvboost.cpp:67-99
  b->params = *p;
  b->params.weak_count = std::max(p->weak_count, 1);
  double trim = std::min(std::max(p->weight_trim_rate, 0.), 1.);
  if (trim < FLT_EPSILON) trim = 1.f;
  b->params.weight_trim_rate = trim;
  b->boost_type = p->boost_type;
  b->crit = p->boost_type == CC_BOOST_DISCRETE ? CRIT_MISCLASS : p->boost_type == CC_BOOST_REAL ? CRIT_GINI : CRIT_SQERR;
  b->is_classifier = (p->boost_type == CC_BOOST_DISCRETE || p->boost_type == CC_BOOST_REAL) ? 1 : 0;
  b->have_subsample = 0;

  {
    ScopedTimer tm(ctx, "boost_small", 1);
    k_boost_init<<<(N + 255) / 256, 256, 0, ctx->stream>>>(N, ctx->d_labels, b->boost_type, b->weights, b->weak_eval, b->stage_sum,
                                                           b->sum_response, b->responses, b->orig_response, b->mask);
    CC_CUDA(cudaGetLastError());
  }
  if (b->d_seed_rows) {     // seed rows of the pruned split search: a strided sample of the local rows
    b->last_nflag = -1;
    // (a feature shard takes its share of the seed rows: the exact search of the seeds is latency-bound, see k_split_search_ord_coop)
    const int seed_target = std::max(256, SEED_TARGET / std::max(1, ctx->shard_world));
    const int stride = std::max(1, ctx->nrows / seed_target);
    b->seed_rows.clear();
    b->seed_rows.reserve(b->seed_cap);
    for (int r = stride / 2; r < ctx->nrows && (int)b->seed_rows.size() < seed_target; r += stride) b->seed_rows.push_back(r);
    b->n_seed_static = (int)b->seed_rows.size();
    CC_CUDA(cudaMemcpyAsync(b->d_seed_rows, b->seed_rows.data(), sizeof(int) * b->seed_rows.size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  out->num_weak = 0; out->num_log = 0; out->threshold = 0.f; out->trained = 0;
  const EvalView ev = make_eval_view(ctx);
  for (;;) {
    {
      ScopedTimer tm(ctx, "boost_small", 1);
      k_tree_begin<<<(N + 255) / 256, 256, 0, ctx->stream>>>(N, b->mask, b->node_of);
      CC_CUDA(cudaGetLastError());
    }
    std::vector<HostNode> nodes(1);
    if (b->params.max_depth == 1 && b->tree_path == 0) {
      bool ok = false;
      CC_CHECK(grow_stump_and_finish(ctx, b, ev, p->min_hit_rate, nodes, ok));
      if (!ok) break;                  // tree->train() returned false (root not split), boost.cpp:436-440
    } else {
    CC_CHECK(grow_node(ctx, b, nodes, 0));
    if (nodes[0].left < 0) break;      // tree->train() returned false (root not split), boost.cpp:436-440
    // upload the tree for the threshold-rule pass
    DevNode dn[CC_MAX_TREE_NODES];
    for (size_t i = 0; i < nodes.size(); ++i) {
      dn[i].left = nodes[i].left; dn[i].right = nodes[i].right; dn[i].var_idx = nodes[i].var; dn[i].c = nodes[i].c;
      std::memcpy(dn[i].subset, nodes[i].subset, sizeof(dn[i].subset));
      dn[i].value = nodes[i].value;
    }
    CC_CUDA(cudaMemcpyAsync(b->d_tree, dn, sizeof(DevNode) * nodes.size(), cudaMemcpyHostToDevice, ctx->stream));
    {
      ScopedTimer tm(ctx, "boost_small", 1);
      k_tree_finish<<<1, ONE_CTA, 0, ctx->stream>>>(ev, N, b->boost_type, b->have_subsample, b->params.weight_trim_rate, p->min_hit_rate,
                                                    b->d_tree, ctx->d_labels, b->orig_response, b->weights, b->weak_eval, b->treeval,
                                                    b->stage_sum, b->sum_response, b->responses, b->mask, b->d_summary, nullptr);
      CC_CUDA(cudaGetLastError());
    }
    CC_CUDA(cudaMemcpyAsync(b->h_summary, b->d_summary, sizeof(TreeSummary), cudaMemcpyDeviceToHost, ctx->stream));
    CC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    const TreeSummary ts = *b->h_summary;
    b->have_subsample = ts.have_subsample;
    // record the tree (DAB: tree->scale(C), src/boost.cpp:316)
    cc_tree& t = out->trees[out->num_weak];
    t.num_nodes = (int)nodes.size(); t.pad = 0;
    for (size_t i = 0; i < nodes.size(); ++i) {
      cc_node& n = t.nodes[i];
      n.left = nodes[i].left; n.right = nodes[i].right; n.var_idx = nodes[i].var; n.split_point = nodes[i].pos;
      n.c = nodes[i].c; n.quality = nodes[i].q; std::memcpy(n.subset, nodes[i].subset, sizeof(n.subset));
      n.value = nodes[i].value * ts.scaleC; n.depth = nodes[i].depth; n.sample_count = nodes[i].n;
    }
    out->num_weak++;
    if (ts.nz_count == 0) break;       // cvCountNonZero(subsample_mask) == 0, boost.cpp:444
    // isErrDesired (prints one table row in the reference)
    out->threshold = ts.threshold;
    out->log[out->num_log].n = out->num_weak;
    out->log[out->num_log].hit_rate = ts.hit_rate;
    out->log[out->num_log].false_alarm = ts.false_alarm;
    out->num_log++;
    if (ts.false_alarm <= p->max_false_alarm) break;
    if (out->num_weak >= b->params.weak_count) break;
  }
  out->trained = out->num_weak > 0;
  return CC_OK;
}

int cc_boost_get_weights(cc_ctx* ctx, double* weights, uint8_t* mask) {
  CC_REQUIRE(ctx && ctx->boost && ctx->boost->weights, CC_ERR_STATE, "cc_boost_get_weights: no stage has been trained");
  CC_CUDA(cudaSetDevice(ctx->device));
  if (weights) CC_CUDA(cudaMemcpyAsync(weights, ctx->boost->weights, sizeof(double) * ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
  if (mask) CC_CUDA(cudaMemcpyAsync(mask, ctx->boost->mask, ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  return CC_OK;
}

int cc_boost_set_search_mode(cc_ctx* ctx, int mode) {
  CC_REQUIRE(ctx && mode >= 0 && mode <= 2, CC_ERR_BAD_ARG, "cc_boost_set_search_mode: mode must be 0 (pruned), 1 (exhaustive) or 2 (wide kernels)");
  if (!ctx->boost) ctx->boost = new BoostBuffers();
  ctx->boost->search_mode = mode;
  if ((mode == 2) != (ctx->force_wide != 0)) ctx->precalc_valid = false;      // the sorted rows have another layout on the wide path
  ctx->force_wide = mode == 2 ? 1 : 0;
  return CC_OK;
}

int cc_boost_set_tree_path(cc_ctx* ctx, int mode) {
  CC_REQUIRE(ctx && (mode == 0 || mode == 1), CC_ERR_BAD_ARG, "cc_boost_set_tree_path: mode must be 0 (stump fast path) or 1 (general path)");
  if (!ctx->boost) ctx->boost = new BoostBuffers();
  ctx->boost->tree_path = mode;
  return CC_OK;
}

int cc_boost_get_search_stats(cc_ctx* ctx, int64_t* searches, int64_t* seed_rows, int64_t* flagged_rows) {
  CC_REQUIRE(ctx && ctx->boost, CC_ERR_STATE, "cc_boost_get_search_stats: no stage has been trained");
  if (searches) *searches = ctx->boost->stat_searches;
  if (seed_rows) *seed_rows = ctx->boost->stat_seed_rows;
  if (flagged_rows) *flagged_rows = ctx->boost->stat_flag_rows;
  return CC_OK;
}

int cc_boost_get_row_quality(cc_ctx* ctx, int row0, int nrows, float* out) {
  CC_REQUIRE(ctx && ctx->boost && ctx->boost->row_quality_valid, CC_ERR_STATE,
             "cc_boost_get_row_quality: no exhaustive split search has run (cc_boost_set_search_mode(ctx, 1) keeps every row's quality)");
  CC_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= ctx->nrows, CC_ERR_BAD_ARG, "cc_boost_get_row_quality: rows out of range");
  CC_CUDA(cudaSetDevice(ctx->device));
  float* tmp = nullptr;
  CC_CUDA(cudaMallocAsync(&tmp, sizeof(float) * nrows, ctx->stream));
  BoostBuffers* b = ctx->boost;
  k_row_quality<<<(nrows + 255) / 256, 256, 0, ctx->stream>>>(b->rowbest ? b->rowbest + row0 : nullptr, b->rowbest_cat ? b->rowbest_cat + row0 : nullptr, nrows, tmp);
  CC_CUDA(cudaGetLastError());
  CC_CUDA(cudaMemcpyAsync(out, tmp, sizeof(float) * nrows, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  CC_CUDA(cudaFreeAsync(tmp, ctx->stream));
  return CC_OK;
}

int cc_stage_predict(cc_ctx* ctx, const cc_stage* stage, int s0, int ns, int return_sum, float* out) {
  CC_REQUIRE(ctx && ctx->ftype >= 0, CC_ERR_STATE, "cc_stage_predict: evaluator not initialised");
  CC_REQUIRE(stage && out && ns > 0 && s0 >= 0 && s0 + ns <= ctx->max_samples, CC_ERR_BAD_ARG, "cc_stage_predict: bad arguments");
  CC_CUDA(cudaSetDevice(ctx->device));
  std::vector<DevNode> nodes;
  std::vector<int> first;
  for (int t = 0; t < stage->num_weak; ++t) {
    first.push_back((int)nodes.size());
    for (int i = 0; i < stage->trees[t].num_nodes; ++i) {
      const cc_node& n = stage->trees[t].nodes[i];
      DevNode d;
      d.left = n.left; d.right = n.right; d.var_idx = n.var_idx; d.c = n.c; std::memcpy(d.subset, n.subset, sizeof(d.subset)); d.value = n.value;
      nodes.push_back(d);
    }
  }
  DevNode* dn = nullptr; int* df = nullptr; float* dout = nullptr;
  CC_CUDA(cudaMallocAsync(&dn, sizeof(DevNode) * std::max<size_t>(1, nodes.size()), ctx->stream));
  CC_CUDA(cudaMallocAsync(&df, sizeof(int) * std::max<size_t>(1, first.size()), ctx->stream));
  CC_CUDA(cudaMallocAsync(&dout, sizeof(float) * ns, ctx->stream));
  if (!nodes.empty()) CC_CUDA(cudaMemcpyAsync(dn, nodes.data(), sizeof(DevNode) * nodes.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (!first.empty()) CC_CUDA(cudaMemcpyAsync(df, first.data(), sizeof(int) * first.size(), cudaMemcpyHostToDevice, ctx->stream));
  k_stage_predict<<<(ns + 127) / 128, 128, 0, ctx->stream>>>(make_eval_view(ctx), dn, df, stage->num_weak, s0, ns, stage->threshold, return_sum, dout);
  CC_CUDA(cudaGetLastError());
  ctx->launches += 1;
  CC_CUDA(cudaMemcpyAsync(out, dout, sizeof(float) * ns, cudaMemcpyDeviceToHost, ctx->stream));
  CC_CUDA(cudaStreamSynchronize(ctx->stream));
  CC_CUDA(cudaFreeAsync(dn, ctx->stream)); CC_CUDA(cudaFreeAsync(df, ctx->stream)); CC_CUDA(cudaFreeAsync(dout, ctx->stream));
  return CC_OK;
}

}  // extern "C"

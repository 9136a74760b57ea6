/*
 * libcascade_cuda — C ABI of the B200 (sm_100a) stage-training and hard-negative-mining path of
 * opencv_traincascade.  Plain pointers and sizes only; no OpenCV, no torch types.
 *
 * Every entry point names the reference interface it stands in for (paths relative to
 * /root/reference/traincascade/lib/).  INTEGRATION.md shows the C++ adapter a maintainer of the
 * reference would add (CvCudaHaarEvaluator : CvFeatureEvaluator, CvCudaCascadeBoost : CvCascadeBoost).
 *
 * Conventions: every function returns CC_OK (0) or a negative cc_status; cc_last_error() returns a
 * thread-local message for the last failure (the adapter turns it into CV_Error).  All host
 * buffers are caller-owned.  One cc_ctx per GPU / per process rank; a ctx is not thread-safe
 * (the reference calls the boundary from its main thread only; SURVEY.md §8b "Threading").
 * There is no CPU fallback: without a CUDA device cc_create fails.
 */
#ifndef CASCADE_CUDA_H_
#define CASCADE_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cc_ctx cc_ctx;

typedef enum {
  CC_OK = 0,
  CC_ERR_CUDA = -1,          /* CUDA runtime / launch failure */
  CC_ERR_BAD_ARG = -2,       /* CV_Assert / StsBadArg equivalents */
  CC_ERR_STATE = -3,         /* call order violated (e.g. train before init) */
  CC_ERR_UNSUPPORTED = -4,   /* outside the implemented envelope (stated in the message) */
  CC_ERR_NCCL = -5
} cc_status;

/* CvFeatureParams::{HAAR,LBP,HOG}, include/traincascade_features.h:141 */
enum { CC_FEATURE_HAAR = 0, CC_FEATURE_LBP = 1, CC_FEATURE_HOG = 2 };
/* CvHaarFeatureParams::{BASIC,CORE,ALL}, include/haarfeatures.h:36 */
enum { CC_HAAR_BASIC = 0, CC_HAAR_CORE = 1, CC_HAAR_ALL = 2 };
/* CvBoost::{DISCRETE,REAL,LOGIT,GENTLE} (cv::ml::Boost), src/boost.cpp:58-62 */
enum { CC_BOOST_DISCRETE = 0, CC_BOOST_REAL = 1, CC_BOOST_LOGIT = 2, CC_BOOST_GENTLE = 3 };

/* ---- lifetime ------------------------------------------------------------------------------ */
int cc_create(int device, cc_ctx** out);
void cc_destroy(cc_ctx* ctx);
const char* cc_last_error(void);
const char* cc_version(void);
/* Run all work of this ctx on an existing cudaStream_t (e.g. torch's current stream); NULL = own stream. */
int cc_set_stream(cc_ctx* ctx, void* cuda_stream);
int cc_synchronize(cc_ctx* ctx);

/* ---- feature evaluator: CvFeatureEvaluator, include/traincascade_features.h:155-188 ------- */
/* init(): CvFeatureEvaluator::init + Cv{Haar,LBP,HOG}Evaluator::init + generateFeatures
 * (src/features.cpp:72-81, src/haarfeatures.cpp:89-98,127-251, src/lbpfeatures.cpp:15-45,
 *  src/HOGfeatures.cpp:17-27,67-106).  Allocates device storage for max_samples windows. */
int cc_eval_init(cc_ctx* ctx, int feature_type, int haar_mode, int win_w, int win_h, int max_samples);
int cc_eval_num_features(cc_ctx* ctx);      /* getNumFeatures() */
int cc_eval_feature_size(cc_ctx* ctx);      /* getFeatureSize(): 1, or 36 for HOG */
int cc_eval_max_cat_count(cc_ctx* ctx);     /* getMaxCatCount(): 0, or 256 for LBP */
/* Catalog read-back for writeFeatures() (src/haarfeatures.cpp:311-320, src/lbpfeatures.cpp:65-68,
 * src/HOGfeatures.cpp:156-160).  Haar: rects[3][4] = x,y,w,h per rect (w==0 => unused), weights[3].
 * LBP: rects[0] = one cell (x,y,w,h).  HOG: rects[0] = cell 0 (x,y,cellW,cellH). */
int cc_eval_get_feature(cc_ctx* ctx, int feature_idx, int32_t rects[12], float weights[3], int32_t* tilted);

/* setImage() batched: CvFeatureEvaluator::setImage + Cv*Evaluator::setImage
 * (src/features.cpp:83-89, src/haarfeatures.cpp:100-114, src/lbpfeatures.cpp:22-28,
 *  src/HOGfeatures.cpp:29-42).  imgs = n windows of win_h*win_w u8 (row-major), labels = n bytes
 * (1 positive, 0 negative), written to evaluator slots first_idx .. first_idx+n-1. */
int cc_eval_set_images(cc_ctx* ctx, const uint8_t* host_imgs, const uint8_t* host_labels, int first_idx, int n);
/* same with the windows already resident in device memory (used by bench "value" and by mining) */
int cc_eval_set_images_device(cc_ctx* ctx, const uint8_t* dev_imgs, const uint8_t* dev_labels, int first_idx, int n);
/* operator()(featureIdx, sampleIdx), include/haarfeatures.h:108-112, lbpfeatures.h:44-45, HOGfeatures.h:84-108 */
int cc_eval_value(cc_ctx* ctx, int var_idx, int sample_idx, float* out);
/* out[nv][ns] = operator()(var_idx[a], s0 + b) */
int cc_eval_values(cc_ctx* ctx, const int32_t* var_idx, int nv, int s0, int ns, float* out);
/* debug / parity getters in the reference's own layout: one (H+1)*(W+1) row per sample */
int cc_eval_get_integrals(cc_ctx* ctx, int sample_idx, int32_t* sum, int32_t* tilted, float* norm_factor);
int cc_eval_get_hog_planes(cc_ctx* ctx, int sample_idx, float* planes10);   /* 9 hist planes + norm, (H+1) x (W+1) each as in the reference */
/* what is resident: the same planes at rows / columns 0, 4, 8, ... only ((H/4+1) x (W/4+1) each) -- every cell corner of the HOG
 * catalog is a multiple of the block step 4 (src/HOGfeatures.cpp:67-106) */
int cc_eval_get_hog_planes_compact(cc_ctx* ctx, int sample_idx, float* planes10c);
int cc_eval_get_labels(cc_ctx* ctx, int first_idx, int n, float* cls);      /* getCls() */
int cc_eval_get_images(cc_ctx* ctx, int first_idx, int n, uint8_t* imgs);   /* the cached u8 windows */

/* ---- stage training: CvCascadeBoost::train, include/boost.h:68-106, src/boost.cpp:409-459 --- */
typedef struct {
  int32_t boost_type;          /* CC_BOOST_* (CvBoostParams::boost_type) */
  int32_t weak_count;          /* -maxWeakCount */
  int32_t max_depth;           /* -maxDepth */
  int32_t min_sample_count;    /* CvDTreeParams default 10 */
  double weight_trim_rate;     /* -weightTrimRate */
  float regression_accuracy;   /* CvDTreeParams default 0.01f */
  float min_hit_rate;          /* -minHitRate */
  float max_false_alarm;       /* -maxFalseAlarmRate */
} cc_boost_params;

/* One tree node in BFS-free "as grown" order; node 0 is the root (CvDTreeNode + CvDTreeSplit,
 * include/o_cvdtreenode.h, include/o_cvdtreesplit.h). */
typedef struct {
  int32_t left, right;         /* child node indices, -1 for a leaf */
  int32_t var_idx;             /* split->var_idx (global variable index) */
  int32_t split_point;         /* split->ord.split_point */
  float c;                     /* split->ord.c */
  float quality;               /* split->quality */
  int32_t subset[8];           /* split->subset (categorical / LBP) */
  double value;                /* node->value (double, as kept in memory by the reference) */
  int32_t depth, sample_count;
} cc_node;

#define CC_MAX_TREE_NODES 63   /* max_depth <= 5 */

typedef struct {
  int32_t num_nodes;
  int32_t pad;
  cc_node nodes[CC_MAX_TREE_NODES];
} cc_tree;

typedef struct { int32_t n; float hit_rate, false_alarm; } cc_log_row;   /* the "| N | HR | FA |" table, src/boost.cpp:511-515 */

typedef struct {
  float threshold;             /* CvCascadeBoost::getThreshold() */
  int32_t num_weak;
  int32_t num_log;
  int32_t trained;             /* return value of CvCascadeBoost::train */
  cc_tree* trees;              /* caller-provided array, capacity >= weak_count */
  cc_log_row* log;             /* caller-provided array, capacity >= weak_count */
} cc_stage;

/* Restrict this ctx to the variable slice [rank*F/world, (rank+1)*F/world) (feature sharding,
 * SURVEY.md §8e).  Default: rank 0 of 1. */
int cc_boost_set_shard(cc_ctx* ctx, int rank, int world);
/* Optional NCCL communicator for the per-node allreduce-argmax; id = ncclUniqueId bytes (128); 1 <= world <= 72. */
int cc_comm_unique_id(void* out_id128);
int cc_comm_init(cc_ctx* ctx, const void* id128, int rank, int world);

/* CvCascadeBoostTrainData::setData + precalculate (src/o_cvcascadeboosttraindata.cpp:214-368,490-596):
 * K2 feature precalc into the fval matrix + K3 per-feature segmented sort, for samples 0..num_samples-1. */
int cc_boost_precalc(cc_ctx* ctx, int num_samples);
/* parity getters (local rows of this shard): fval rows as float (LBP codes widened), sorted index rows */
int cc_boost_get_fval(cc_ctx* ctx, int row0, int nrows, float* out);          /* out[nrows][num_samples] */
int cc_boost_get_sorted_idx(cc_ctx* ctx, int row0, int nrows, uint16_t* out); /* out[nrows][num_samples] */
int cc_boost_get_sorted_val(cc_ctx* ctx, int row0, int nrows, float* out);
/* same rows as 32-bit indices: the only form that exists when numSamples >= 65536 (is_buf_16u == false,
 * src/o_cvcascadeboosttraindata.cpp:250-251) */
int cc_boost_get_sorted_idx32(cc_ctx* ctx, int row0, int nrows, uint32_t* out);

/* The whole stage: precalc (if not done for this sample set) + boosting loop; fills `out`. */
int cc_boost_train(cc_ctx* ctx, int num_samples, const cc_boost_params* params, cc_stage* out);
/* state after training, for parity tests: AdaBoost weights and subsample mask (src/boost.cpp:168-407,
 * src/o_cvboost.cpp:101-139) */
int cc_boost_get_weights(cc_ctx* ctx, double* weights, uint8_t* mask);
/* Ordered-variable split search strategy.  0 (default) = pruned: a bound pass over the u16 index rows drops every
 * variable that cannot beat the best split already known, the rest is searched exactly -- same winner as the
 * serial loop of DTreeBestSplitFinder (src/o_cvdtree.cpp:313-357), whose init_quality threading prunes in the
 * same spirit.  1 = exhaustive: every variable searched exactly (keeps per-variable qualities for the getter).
 * 2 = wide kernels: the path stages with more than 15360 samples always take (rows sorted by a global-memory radix sort into
 * 32-bit index rows, per-sample operands read from global memory, every variable searched exactly) forced for any numSamples;
 * there is no upper limit on numSamples other than device memory.  All modes grow the same trees. */
int cc_boost_set_search_mode(cc_ctx* ctx, int mode);
/* How a weak tree of depth 1 is grown.  0 (default): root statistics, split search, exchange between the ranks, the split, both
 * leaves and the weight update / trimming / isErrDesired pass are queued back to back and the host reads the result once per
 * tree.  1: the general node-by-node path that deeper trees always take (CvDTree::try_split_node recursion, src/o_cvdtree.cpp:122-187),
 * with a host decision after every node.  Same trees. */
int cc_boost_set_tree_path(cc_ctx* ctx, int mode);
/* How the sorted-index rows (std::sort + LessThanIdx per feature, src/o_cvcascadeboosttraindata.cpp:490-563) are built.
 * 0 (default) = Haar: bucket + rank sort (one distribution pass into value buckets, exact ranking inside a bucket) with
 * the radix kernel on the rows it defers; HOG: radix kernel.  1 = stable LSD radix passes on every row.  2 = bucket +
 * rank for every ordered feature type.  All modes give the same rows: ascending value, ties by ascending sample index. */
int cc_boost_set_sort_mode(cc_ctx* ctx, int mode);
/* rows of the last cc_boost_precalc that the bucket sort deferred to the radix kernel (local slice) */
int cc_boost_get_sort_stats(cc_ctx* ctx, int64_t* deferred_rows);
/* counters since cc_create: pruned searches run, rows searched exactly as seeds, rows flagged by the bound pass */
int cc_boost_get_search_stats(cc_ctx* ctx, int64_t* searches, int64_t* seed_rows, int64_t* flagged_rows);
/* per-variable best split quality of the LAST root-node search (0 where no split), local rows; exhaustive mode only */
int cc_boost_get_row_quality(cc_ctx* ctx, int row0, int nrows, float* out);

/* CvCascadeBoost::predict(sampleIdx), src/boost.cpp:461-477, for stages given as flat arrays */
int cc_stage_predict(cc_ctx* ctx, const cc_stage* stage, int s0, int ns, int return_sum, float* out);

/* ---- mining: CvCascadeClassifier::updateTrainingSet / fillPassedSamples,
 *      src/cascadeclassifier.cpp:297-357, and NegReader, src/imagestorage.cpp:23-126 ------------ */
typedef struct {   /* NegReader state (src/imagestorage.cpp:23-33) so that runs can be resumed exactly */
  int64_t last, round;
  int32_t point_x, point_y, offset_x, offset_y;
  int32_t level_w, level_h;   /* size of the current pyramid level (img), 0 before the first image */
  float scale;
  int32_t pad;
} cc_reader_state;

int cc_bg_clear(cc_ctx* ctx);
/* register one decoded 8-bit gray background (cv::imread(path, 0)); kept resident in HBM.
 * (NULL, 0, 0) registers the placeholder of a file imread() could not decode: it keeps its position in the list and
 * NegReader::nextImg passes over it without advancing `round`, exactly like src/imagestorage.cpp:62-70 */
int cc_bg_register(cc_ctx* ctx, const uint8_t* host_gray, int w, int h);
/* same for an image that is already in device memory (decoded on the GPU / generated there); row pitch in bytes.  The copy is
 * ordered on the ctx's stream: the source may be reused after cc_synchronize */
int cc_bg_register_device(cc_ctx* ctx, const uint8_t* dev_gray, int w, int h, int pitch);
int cc_bg_count(cc_ctx* ctx);
/* the cascade used by predict() during mining: stages 0..n-1 (copied) */
int cc_cascade_set_stages(cc_ctx* ctx, const cc_stage* stages, int n);
/* fillPassedSamples(first, count, isPositive=false, minimumAcceptanceRatio, consumed):
 * runs the cascade over the NegReader window stream starting at *state, stores accepted windows
 * into evaluator slots first..first+got-1 with label 0, and leaves *state right after the last
 * consumed window.  image_rank/image_world shard the background list across GPUs. */
int cc_mine_negatives(cc_ctx* ctx, int first_slot, int count, double min_acceptance_ratio,
                      cc_reader_state* state, int64_t* consumed, int32_t* got);
/* fillPassedSamples(first=0, count, isPositive=true): positives come from the decoded .vec block
 * (PosReader, src/imagestorage.cpp:138-189): host_vec = vec_count windows; *vec_pos = PosReader::last */
int cc_mine_positives(cc_ctx* ctx, const uint8_t* host_vec, int vec_count, int count,
                      int32_t* vec_pos, int64_t* consumed, int32_t* got);
/* How predict() is evaluated over the window stream.  0 (default): Haar / LBP windows are evaluated from integral
 * planes of whole pyramid levels built by one fused kernel (same integers as the per-crop cv::integral of setImage,
 * src/haarfeatures.cpp:100-114), HOG windows by the window miner (one CTA per window: gradients, integral histograms
 * and cascade in shared memory; windows wider than 100 or taller than 128 pixels fall back to slots).  1: same planes,
 * built by separate row / diagonal / column kernels.  2: per-window slots for every feature type (literal setImage + predict per window).  3: fused tile
 * miner -- a CTA builds a tile of a pyramid level (INTER_LINEAR_EXACT pixels, sum / sqsum / tilted integrals) in shared
 * memory and runs the cascade on its windows; nothing but survivor ordinals leaves the SM (measured 2.5x slower than
 * mode 0 on B200: 16 resident warps per SM cannot hide the serial column / diagonal walks). */
int cc_mine_set_mode(cc_ctx* ctx, int mode);
/* Survivor cache of the negative window stream (default: on, 4 GiB of host memory).  NegReader cycles the background list with a
 * `round` phase (src/imagestorage.cpp:57-88), so a short list makes the stream revisit the same (image, offsets) pyramid many
 * times in late stages; stages are append-only, so the windows of a pyramid that passed stages [0, n) are kept and a revisit
 * runs only the newer stages on them (per-window setImage + predict path) instead of rebuilding the pyramid.  Transparent:
 * consumed counts, the ordered cut, the reader state and the mined windows are those of the uncached evaluation. */
int cc_mine_set_cache(cc_ctx* ctx, int enable, int64_t max_bytes);
/* stats6 = { stream windows consumed by cc_mine_negatives, windows evaluated from pyramid planes / slots (speculative ones included),
 *            cached survivors re-evaluated, stream windows answered by the cache, cache entries, cache bytes } since cc_create */
int cc_mine_get_stats(cc_ctx* ctx, int64_t* stats6);
/* NegReader::get() x nwindows without pixels (host only, no GPU needed): advances *state over images of
 * the given sizes; window5 (optional) = x, y, level_w, level_h, source image of the last consumed window. */
int cc_reader_advance(const int32_t* img_w, const int32_t* img_h, int nimg, int win_w, int win_h,
                      cc_reader_state* state, int64_t nwindows, int32_t* window5);
/* The serial part of fillPassedSamples (src/cascadeclassifier.cpp:333-355) for a batch whose windows have already been
 * evaluated (host only, no GPU needed): survivors = ascending batch ordinals that pass the cascade, nwin = windows in
 * the batch.  Walks the stream order, stops at the count-th acceptance or when (getcount+1)/consumed <=
 * minimumAcceptanceRatio holds before a window.  consumed / got are in-out; next_window = first unconsumed ordinal;
 * accepted (capacity >= count) receives the accepted ordinals.  With image-sharded mining every rank runs this on the
 * union of the ranks' survivors, which is what makes the ordered cut identical on all GPUs. */
int cc_mine_serial_cut(const uint32_t* survivors, int64_t nsurv, int64_t nwin, int32_t count, double min_ratio, int64_t* consumed,
                       int32_t* got, int64_t* next_window, int32_t* stopped, uint32_t* accepted, int32_t* n_accepted);
/* throughput probe used by bench.py: evaluate the cascade on every window of backgrounds
 * [img0, img0+nimg) at round 0 without touching evaluator slots; returns windows and survivors */
int cc_mine_scan(cc_ctx* ctx, int img0, int nimg, int64_t* windows, int64_t* survivors);
/* same, plus an order-independent 64-bit digest of the scan ordinals of the surviving windows: two miners (cc_mine_set_mode)
 * or two runs agree on WHICH windows pass iff their digests agree */
int cc_mine_scan_digest(cc_ctx* ctx, int img0, int nimg, int64_t* windows, int64_t* survivors, uint64_t* digest);

/* ---- training-sample generation: opencv_createsamples -img <image> -num <count> -vec <out> (no -bg),
 *      tools/createsamples/utility.cpp:952-1029 cvCreateTrainingSamples + :419-672 (icvRandomQuad, the tool's own scan-line
 *      cvWarpPerspective, icvStartSampleDistortion, icvPlaceDistortedSample), flags of tools/createsamples/createsamples.cpp:60-166 --- */
#define CC_RANDOM_INVERT 0x7FFFFFFF   /* -randinv (CV_RANDOM_INVERT) */
/* host_gray = the source image (imread(path, IMREAD_GRAYSCALE)), w x h.  out_samples receives count windows of win_h * win_w bytes:
 * the payload of the .vec file (header / record layout: utility.cpp:128-153).  The cv::RNG stream (cv::setRNGSeed(rngseed)), the
 * rotations, the perspective systems and the edge walks are computed on the host in f64 exactly in the reference's order; the
 * warps, the mask blur, the INTER_LINEAR_EXACT resizes and the blend run on the GPU. */
int cc_createsamples(cc_ctx* ctx, const uint8_t* host_gray, int w, int h, int bgcolor, int bgthreshold, int count, int invert,
                     int maxintensitydev, double maxxangle, double maxyangle, double maxzangle, int win_w, int win_h, int rngseed,
                     uint8_t* out_samples);
/* the host half alone (no GPU needed; parity tests): per sample quads[8] (the distorted quadrangle in the enlarged image),
 * coeffs[9] (perspective transform), misc[8] = crop x, y, w, h, forecolordev, inverse, warp rows, warp pixels */
int cc_createsamples_plan(int w, int h, int count, int invert, int maxintensitydev, double maxxangle, double maxyangle, double maxzangle,
                          int win_w, int win_h, int rngseed, double* quads, double* coeffs, int32_t* misc);

/* ---- instrumentation ------------------------------------------------------------------------ */
/* CUDA-event time (ms) and launch count accumulated per kernel family since the last reset.
 * names: "integral", "precalc", "sort", "split_search", "boost_small", "mine" */
int cc_timing_reset(cc_ctx* ctx);
int cc_timing_get(cc_ctx* ctx, const char* name, double* ms, int64_t* launches);
int64_t cc_launch_count(cc_ctx* ctx);
/* self-check of the reciprocal-based Haar division used by K2 against IEEE division on n pseudo-random
 * (integer numerator, normalisation factor) pairs; steps = number of correction iterations (1 or 2) */
int cc_debug_div_check(cc_ctx* ctx, int steps, uint64_t seed, int64_t n, uint64_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif /* CASCADE_CUDA_H_ */

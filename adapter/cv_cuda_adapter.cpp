// See cv_cuda_adapter.h.  Construction of the node graph follows CvCascadeBoostTree::read
// (traincascade/lib/src/o_cvcascadeboosttree.cpp:95-165) — the reference's own recipe for building a tree from flat
// data — except that leaf values stay double (in-memory training keeps doubles; only XML rounds to float).
#include "cv_cuda_adapter.h"

#include <vector>

#include "o_cvcascadeboosttraindata.h"
#include "o_cvdtreenode.h"
#include "o_cvdtreesplit.h"

bool CvCudaCascadeBoost::train(const CvFeatureEvaluator* featureEvaluator, int numSamples, int precalcValBufSize,
                               int precalcIdxBufSize, const CvCascadeBoostParams& p) {
  CV_Assert(!data);
  clear();
  if (precalcValBufSize < 0 || precalcIdxBufSize < 0)
    CV_Error(cv::Error::StsOutOfRange, "_numPrecalcVal and _numPrecalcIdx must be positive or 0");
  // the evaluator-only train-data constructor is the one CvCascadeBoost::read uses: node/split heaps, no buffers
  data = new CvCascadeBoostTrainData(featureEvaluator, p);
  set_params(p);
  CvMemStorage* storage = cvCreateMemStorage();
  weak = cvCreateSeq(0, sizeof(CvSeq), sizeof(CvBoostTree*), storage);

  cc_boost_params bp;
  bp.boost_type = p.boost_type; bp.weak_count = p.weak_count; bp.max_depth = p.max_depth;
  bp.min_sample_count = p.min_sample_count; bp.weight_trim_rate = p.weight_trim_rate;
  bp.regression_accuracy = p.regression_accuracy; bp.min_hit_rate = p.minHitRate; bp.max_false_alarm = p.maxFalseAlarm;
  std::vector<cc_tree> trees(std::max(1, p.weak_count));
  std::vector<cc_log_row> log(std::max(1, p.weak_count));
  cc_stage st{0.f, 0, 0, 0, trees.data(), log.data()};
  CC_CALL(cc_boost_train(cu->ctx, numSamples, &bp, &st));

  std::cout << "+----+---------+---------+" << std::endl << "|  N |    HR   |    FA   |" << std::endl
            << "+----+---------+---------+" << std::endl;
  for (int i = 0; i < st.num_log; ++i) {
    std::cout << "|"; std::cout.width(4); std::cout << std::right << log[i].n;
    std::cout << "|"; std::cout.width(9); std::cout << std::right << log[i].hit_rate;
    std::cout << "|"; std::cout.width(9); std::cout << std::right << log[i].false_alarm;
    std::cout << "|" << std::endl << "+----+---------+---------+" << std::endl;
  }
  threshold = st.threshold;
  for (int t = 0; t < st.num_weak; ++t) {
    auto* tree = new CvCudaCascadeBoostTree();
    buildTree(trees[t], tree);
    cvSeqPush(weak, &tree);
  }
  if (weak->total > 0) { data->is_classifier = true; return true; }
  clear();
  return false;
}

void CvCudaCascadeBoost::buildTree(const cc_tree& t, CvCudaCascadeBoostTree* tree) {
  const int maxCat = (dynamic_cast<CvCascadeBoostTrainData*>(data))->featureEvaluator->getMaxCatCount();
  const int subsetN = (maxCat + 31) / 32;
  std::vector<CvDTreeNode*> nodes(t.num_nodes, nullptr);
  for (int i = 0; i < t.num_nodes; ++i) nodes[i] = data->new_node(nullptr, t.nodes[i].sample_count, 0, 0);
  for (int i = 0; i < t.num_nodes; ++i) {
    const cc_node& n = t.nodes[i];
    nodes[i]->value = n.value;
    nodes[i]->depth = n.depth;
    if (n.left >= 0) {
      CvDTreeSplit* s = maxCat ? data->new_split_cat(n.var_idx, n.quality) : data->new_split_ord(n.var_idx, n.c, n.split_point, 0, n.quality);
      if (maxCat) for (int k = 0; k < subsetN; ++k) s->subset[k] = n.subset[k];
      nodes[i]->split = s;
      nodes[i]->left = nodes[n.left];   nodes[n.left]->parent = nodes[i];
      nodes[i]->right = nodes[n.right]; nodes[n.right]->parent = nodes[i];
    }
  }
  tree->adopt(nodes[0], data, this);
}

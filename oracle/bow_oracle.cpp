// bow_oracle.cpp — CPU oracle for DBoW2's descriptor -> (BowVector, FeatureVector) transform (TEST INFRASTRUCTURE,
// NOT PRODUCT).  SURVEY.md §8(f) rank 2: the step between ORBextractor and SearchByBoW on every frame / keyframe
// (Frame::ComputeBoW S/Frame.cpp:268-275, KeyFrame::ComputeBoW S/KeyFrame.cpp:277-286).
//
// Follows D/ = cslam/thirdparty/DBoW2/DBoW2:
//   TemplatedVocabulary::loadFromTextFile       D/TemplatedVocabulary.h:1338-1422  (node ids in file order from 1, word
//                                                                                   ids in order of the leaf flags)
//   TemplatedVocabulary::transform (features)   D/TemplatedVocabulary.h:1127-1192
//   TemplatedVocabulary::transform (one)        D/TemplatedVocabulary.h:1219-1260  (first minimum wins; node at level
//                                                                                   L - levelsup goes to the FeatureVector)
//   FORB::distance                              D/FORB.cpp:77-100
//   BowVector::addWeight/addIfNotExist/normalize D/BowVector.cpp:34-88
//   FeatureVector::addFeature                   D/FeatureVector.cpp:28-43
//   GeneralScoring::mustNormalize               D/ScoringObject.h:53-89  (DOT_PRODUCT: no normalisation, L2_NORM: L2, else L1)
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace {

enum Weighting { TF_IDF = 0, TF = 1, IDF = 2, BINARY = 3 };                                            // D/BowVector.h:36-42
enum Scoring { L1_NORM = 0, L2_NORM = 1, CHI_SQUARE = 2, KL = 3, BHATTACHARYYA = 4, DOT_PRODUCT = 5 };  // D/BowVector.h:45-53

struct Node {
  std::vector<uint32_t> children;
  uint8_t desc[32];
  double weight = 0;
  uint32_t word_id = 0;
};

int forb_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t pa, pb;
    memcpy(&pa, a + 4 * i, 4); memcpy(&pb, b + 4 * i, 4);
    unsigned int v = pa ^ pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

}  // namespace

struct orc_voc {
  int k, L, scoring, weighting;
  std::vector<Node> nodes;
  int n_words = 0;
};

extern "C" {

// rows 1..n_nodes-1 are the lines of the text file: parent id, leaf flag, 32 descriptor bytes, weight; row 0 is the root
orc_voc* orc_voc_create(int32_t k, int32_t L, int32_t scoring, int32_t weighting, int32_t n_nodes, const int32_t* parent,
                        const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
  orc_voc* v = new orc_voc;
  v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
  v->nodes.resize(n_nodes);
  for (int nid = 1; nid < n_nodes; nid++) {
    const int pid = parent[nid];
    if (pid < 0 || pid >= nid) { delete v; return nullptr; }   // the file format lists a parent before its children
    v->nodes[pid].children.push_back((uint32_t)nid);
    memcpy(v->nodes[nid].desc, desc + 32 * (size_t)nid, 32);
    v->nodes[nid].weight = weight[nid];
    if (is_leaf[nid] > 0) v->nodes[nid].word_id = (uint32_t)v->n_words++;
  }
  return v;
}

void orc_voc_destroy(orc_voc* v) { delete v; }

int32_t orc_voc_words(const orc_voc* v) { return v->n_words; }

// per-feature word id / weight / node id, then the two containers flattened in key order.
// Capacities: everything sized n (fv_node_ptr n+1).  Returns 0.
int orc_voc_transform(const orc_voc* v, const uint8_t* feat, int32_t n, int32_t levelsup, uint32_t* word_of_feat, uint32_t* node_of_feat,
                      double* weight_of_feat, uint32_t* bow_id, double* bow_val, int32_t* bow_n, uint32_t* fv_node_id,
                      int32_t* fv_node_ptr, uint32_t* fv_feat, int32_t* fv_n_nodes) {
  std::map<uint32_t, double> bow;
  std::map<uint32_t, std::vector<uint32_t>> fv;
  *bow_n = 0; *fv_n_nodes = 0; fv_node_ptr[0] = 0;
  if (v->n_words == 0) return 0;  // empty()
  const bool must = v->scoring != DOT_PRODUCT;
  const bool l2 = v->scoring == L2_NORM;
  const int nid_level = v->L - levelsup;
  for (int i = 0; i < n; i++) {
    const uint8_t* f = feat + 32 * (size_t)i;
    uint32_t nid = 0, final_id = 0;
    int current_level = 0;
    do {
      ++current_level;
      const std::vector<uint32_t>& nodes = v->nodes[final_id].children;
      final_id = nodes[0];
      double best_d = forb_distance(f, v->nodes[final_id].desc);
      for (size_t c = 1; c < nodes.size(); c++) {
        const uint32_t id = nodes[c];
        const double d = forb_distance(f, v->nodes[id].desc);
        if (d < best_d) { best_d = d; final_id = id; }
      }
      if (current_level == nid_level) nid = final_id;
    } while (!v->nodes[final_id].children.empty());
    const uint32_t id = v->nodes[final_id].word_id;
    const double w = v->nodes[final_id].weight;
    word_of_feat[i] = id; node_of_feat[i] = nid; weight_of_feat[i] = w;
    if (w > 0) {
      if (v->weighting == TF || v->weighting == TF_IDF) bow[id] += w;   // addWeight (a new key starts from w: 0 + w == w)
      else bow.insert({id, w});                                         // addIfNotExist
      fv[nid].push_back((uint32_t)i);
    }
  }
  if ((v->weighting == TF || v->weighting == TF_IDF) && !bow.empty() && !must) {
    const double nd = (double)bow.size();
    for (auto& kv : bow) kv.second /= nd;
  }
  if (must) {
    double norm = 0.0;
    if (!l2) { for (auto& kv : bow) norm += fabs(kv.second); }
    else { for (auto& kv : bow) norm += kv.second * kv.second; norm = sqrt(norm); }
    if (norm > 0.0) for (auto& kv : bow) kv.second /= norm;
  }
  int b = 0;
  for (auto& kv : bow) { bow_id[b] = kv.first; bow_val[b] = kv.second; b++; }
  *bow_n = b;
  int nn = 0, pos = 0;
  for (auto& kv : fv) {
    fv_node_id[nn] = kv.first;
    for (uint32_t fi : kv.second) fv_feat[pos++] = fi;
    fv_node_ptr[++nn] = pos;
  }
  *fv_n_nodes = nn;
  return 0;
}

}  // extern "C"

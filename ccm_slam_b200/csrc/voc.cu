// voc.cu — DBoW2 vocabulary on the device: descriptor -> word / node by tree descent, behind ccm_voc_* (include/ccm_b200.h).
// SURVEY.md §8(f) rank 2; D/ = cslam/thirdparty/DBoW2/DBoW2 under /root/reference.
//
//   TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)   D/TemplatedVocabulary.h:1219-1260
//   TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, l)   D/TemplatedVocabulary.h:1127-1192
//   TemplatedVocabulary::loadFromTextFile (node / word numbering)             D/TemplatedVocabulary.h:1338-1422
//   FORB::distance                                                            D/FORB.cpp:77-100
//   BowVector::addWeight / addIfNotExist / normalize                          D/BowVector.cpp:34-88
//   FeatureVector::addFeature                                                 D/FeatureVector.cpp:28-43
//
// Layout: the tree lives in HBM for the life of the handle (ORBvoc: 1.08 M nodes x 32 B = 35 MB): node descriptors as
// uint4 pairs, children as a CSR (child_ptr, child_idx) in the file's push_back order — for the usual depth-first file the
// k children of a node are consecutive ids, i.e. one 320-byte run.  k_voc_descend: 8 lanes per descriptor, each lane scores
// children c, c+8, ... and the sub-warp keeps the FIRST minimum (ties -> lowest child position, as the reference's strict
// '<' scan); L levels x k distances per descriptor.  The two std::map containers are order dependent f64 sums and stay on
// the host (n <= a few thousand).
#include <cmath>
#include <map>
#include <memory>

#include "common.cuh"

using namespace ccm;

namespace {

enum { W_TF_IDF = 0, W_TF = 1, W_IDF = 2, W_BINARY = 3 };   // DBoW2::WeightingType (D/BowVector.h:36-42)
enum { S_L1 = 0, S_L2 = 1, S_CHI = 2, S_KL = 3, S_BHAT = 4, S_DOT = 5 };  // DBoW2::ScoringType (D/BowVector.h:45-53)

constexpr int SUB = 8;   // lanes per descriptor

__global__ void __launch_bounds__(256) k_voc_descend(const uint4* __restrict__ feat, int n, const uint4* __restrict__ node_desc,
                                                     const int* __restrict__ child_ptr, const int* __restrict__ child_idx,
                                                     int nid_level, int* __restrict__ leaf_of_feat, int* __restrict__ nid_of_feat) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = t / SUB, sub = t % SUB;
  const bool live = f < n;
  const int ff = live ? f : n - 1;                     // idle sub-warps shadow the last descriptor: shuffles stay converged
  const uint4 f0 = feat[(size_t)ff * 2], f1 = feat[(size_t)ff * 2 + 1];
  const unsigned lane = threadIdx.x & 31u;
  // shuffles are confined to the descriptor's own 8 lanes: sub-warps of one warp leave the loop at different depths
  const unsigned mask = 0xffu << (lane & ~(unsigned)(SUB - 1));
  int cur = 0, nid = 0, level = 0;
  while (true) {
    const int beg = child_ptr[cur], end = child_ptr[cur + 1];
    if (beg == end) break;                             // isLeaf(): no children
    level++;
    // key = distance << 16 | position in the child list: the minimum key is the first minimum of the reference's scan
    unsigned best = 0xffffffffu;
    for (int c = beg + sub; c < end; c += SUB) {
      const int id = child_idx[c];
      const uint4 d0 = node_desc[(size_t)id * 2], d1 = node_desc[(size_t)id * 2 + 1];
      const unsigned d = __popc(f0.x ^ d0.x) + __popc(f0.y ^ d0.y) + __popc(f0.z ^ d0.z) + __popc(f0.w ^ d0.w) +
                         __popc(f1.x ^ d1.x) + __popc(f1.y ^ d1.y) + __popc(f1.z ^ d1.z) + __popc(f1.w ^ d1.w);
      const unsigned key = (d << 16) | (unsigned)(c - beg);
      best = min(best, key);
    }
#pragma unroll
    for (int off = SUB / 2; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(mask, best, off));
    cur = child_idx[beg + (int)(best & 0xffffu)];
    if (level == nid_level) nid = cur;
  }
  if (live && sub == 0) {
    leaf_of_feat[f] = cur;
    nid_of_feat[f] = nid;
  }
}

}  // namespace

struct ccm_voc_handle {
  int k = 0, L = 0, scoring = 0, weighting = 0, n_nodes = 0, n_words = 0, device = 0;
  std::vector<uint32_t> word_of_node;   // host: leaf node -> word id
  std::vector<double> weight_of_node;   // host
  DevBuf<uint4> d_desc;
  DevBuf<int> d_child_ptr, d_child_idx;
  DevBuf<uint4> d_feat;
  DevBuf<int> d_leaf, d_nid;
  std::vector<int> h_leaf, h_nid;
  cudaStream_t stream = nullptr;
  ~ccm_voc_handle() {
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

// BowVector + FeatureVector from the per-feature results, in feature order (host; shared by ccm_voc_transform and ccm_bow_assemble)
void assemble(int scoring, int weighting, int n, const uint32_t* word, const double* weight, const uint32_t* node, uint32_t* bow_id,
              double* bow_val, int32_t* bow_n, uint32_t* fv_node_id, int32_t* fv_node_ptr, uint32_t* fv_feat, int32_t* fv_n_nodes) {
  CCM_REQUIRE(scoring >= 0 && scoring <= 5 && weighting >= 0 && weighting <= 3, "ccm_voc: unknown scoring / weighting type");
  CCM_REQUIRE(n >= 0 && bow_n && fv_n_nodes && fv_node_ptr && (n == 0 || (word && weight && node && bow_id && bow_val && fv_node_id && fv_feat)),
              "ccm_voc: null argument");
  std::map<uint32_t, double> bow;
  std::map<uint32_t, std::vector<uint32_t>> fv;
  const bool tf = weighting == W_TF || weighting == W_TF_IDF;
  for (int i = 0; i < n; i++) {
    const double w = weight[i];
    if (!(w > 0)) continue;                              // stopped word
    auto it = bow.lower_bound(word[i]);
    if (it != bow.end() && it->first == word[i]) {
      if (tf) it->second += w;                           // addWeight; addIfNotExist leaves the first value
    } else {
      bow.insert(it, {word[i], w});
    }
    fv[node[i]].push_back((uint32_t)i);
  }
  const bool must = scoring != S_DOT;                    // GeneralScoring::mustNormalize (D/ScoringObject.h:74-89)
  if (tf && !bow.empty() && !must) {
    const double nd = (double)bow.size();
    for (auto& kv : bow) kv.second /= nd;
  }
  if (must) {
    double norm = 0.0;
    if (scoring == S_L2) {
      for (auto& kv : bow) norm += kv.second * kv.second;
      norm = sqrt(norm);
    } else {
      for (auto& kv : bow) norm += fabs(kv.second);
    }
    if (norm > 0.0)
      for (auto& kv : bow) kv.second /= norm;
  }
  int b = 0;
  for (auto& kv : bow) { bow_id[b] = kv.first; bow_val[b] = kv.second; b++; }
  *bow_n = b;
  int nn = 0, pos = 0;
  fv_node_ptr[0] = 0;
  for (auto& kv : fv) {
    fv_node_id[nn] = kv.first;
    for (uint32_t fi : kv.second) fv_feat[pos++] = fi;
    fv_node_ptr[++nn] = pos;
  }
  *fv_n_nodes = nn;
}

}  // namespace

extern "C" {

int ccm_voc_create(int32_t k, int32_t L, int32_t scoring, int32_t weighting, int32_t n_nodes, const int32_t* parent,
                   const uint8_t* is_leaf, const uint8_t* desc, const double* weight, ccm_voc_handle** out) {
  return guarded([&] {
    CCM_REQUIRE(out && parent && is_leaf && desc && weight && n_nodes >= 1, "ccm_voc_create: null argument");
    CCM_REQUIRE(k >= 0 && k <= 20 && L >= 1 && L <= 10 && scoring >= 0 && scoring <= 5 && weighting >= 0 && weighting <= 3,
                "ccm_voc_create: not a vocabulary header (D/TemplatedVocabulary.h:1359)");
    *out = nullptr;
    ensure_device();
    std::unique_ptr<ccm_voc_handle> h(new ccm_voc_handle);
    h->k = k; h->L = L; h->scoring = scoring; h->weighting = weighting; h->n_nodes = n_nodes; h->device = current_device();
    // children CSR in push_back order = ascending node id per parent (counting sort by parent)
    std::vector<int> cptr((size_t)n_nodes + 1, 0), cidx((size_t)std::max(0, n_nodes - 1));
    for (int i = 1; i < n_nodes; i++) {
      CCM_REQUIRE(parent[i] >= 0 && parent[i] < i, "ccm_voc_create: a node must come after its parent");
      cptr[parent[i] + 1]++;
    }
    for (int i = 0; i < n_nodes; i++) {
      CCM_REQUIRE(cptr[i + 1] < 65536, "ccm_voc_create: more than 65535 children under one node");
      cptr[i + 1] += cptr[i];
    }
    {
      std::vector<int> fill(cptr.begin(), cptr.end() - 1);
      for (int i = 1; i < n_nodes; i++) cidx[fill[parent[i]]++] = i;
    }
    h->word_of_node.assign(n_nodes, 0);
    h->weight_of_node.assign(weight, weight + n_nodes);
    h->weight_of_node[0] = 0.0;
    for (int i = 1; i < n_nodes; i++)
      if (is_leaf[i] > 0) h->word_of_node[i] = (uint32_t)h->n_words++;
    CCM_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->d_desc.upload(reinterpret_cast<const uint4*>(desc), (size_t)n_nodes * 2, h->stream);
    h->d_child_ptr.upload(cptr.data(), cptr.size(), h->stream);
    if (!cidx.empty()) h->d_child_idx.upload(cidx.data(), cidx.size(), h->stream);
    else h->d_child_idx.alloc(1);
    CCM_CUDA(cudaStreamSynchronize(h->stream));
    *out = h.release();
  });
}

int ccm_voc_words(const ccm_voc_handle* h) { return h ? h->n_words : 0; }

// desc: host descriptors (uploaded) or, when d_resident != nullptr, descriptors already on the device (kf_store.cu: uploaded once at ingest)
static int voc_transform_any(ccm_voc_handle* h, const uint8_t* desc, const uint4* d_resident, int32_t n, int32_t levelsup, uint32_t* word_of_feat,
                             uint32_t* node_of_feat, double* weight_of_feat, uint32_t* bow_id, double* bow_val, int32_t* bow_n,
                             uint32_t* fv_node_id, int32_t* fv_node_ptr, uint32_t* fv_feat, int32_t* fv_n_nodes) {
  return guarded([&] {
    CCM_REQUIRE(h && n >= 0 && (n == 0 || desc || d_resident) && bow_n && fv_n_nodes && fv_node_ptr, "ccm_voc_transform: null argument");
    *bow_n = 0; *fv_n_nodes = 0; fv_node_ptr[0] = 0;
    if (n == 0 || h->n_words == 0) return;               // transform() of an empty vocabulary clears both containers
    ensure_device();
    CCM_CUDA(cudaSetDevice(h->device));
    const uint4* d_in = d_resident;
    if (!d_in) { h->d_feat.upload(reinterpret_cast<const uint4*>(desc), (size_t)n * 2, h->stream); d_in = h->d_feat.p; }
    if (h->d_leaf.n < (size_t)n) { h->d_leaf.alloc(n + 256); h->d_nid.alloc(n + 256); }
    const int threads = 256;
    k_voc_descend<<<div_up((long long)n * SUB, threads), threads, 0, h->stream>>>(d_in, n, h->d_desc.p, h->d_child_ptr.p,
                                                                                  h->d_child_idx.p, h->L - levelsup, h->d_leaf.p, h->d_nid.p);
    CCM_LAUNCHED();
    h->h_leaf.resize(n); h->h_nid.resize(n);
    h->d_leaf.download(h->h_leaf.data(), n, h->stream);
    h->d_nid.download(h->h_nid.data(), n, h->stream);
    CCM_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<uint32_t> word(n), node(n);
    std::vector<double> w(n);
    for (int i = 0; i < n; i++) {
      const int leaf = h->h_leaf[i];
      CCM_REQUIRE(leaf > 0 && leaf < h->n_nodes, "ccm_voc_transform: descent left the tree");
      word[i] = h->word_of_node[leaf];
      w[i] = h->weight_of_node[leaf];
      node[i] = (uint32_t)h->h_nid[i];
    }
    if (word_of_feat) memcpy(word_of_feat, word.data(), sizeof(uint32_t) * n);
    if (node_of_feat) memcpy(node_of_feat, node.data(), sizeof(uint32_t) * n);
    if (weight_of_feat) memcpy(weight_of_feat, w.data(), sizeof(double) * n);
    assemble(h->scoring, h->weighting, n, word.data(), w.data(), node.data(), bow_id, bow_val, bow_n, fv_node_id, fv_node_ptr, fv_feat,
             fv_n_nodes);
  });
}

int ccm_voc_transform(ccm_voc_handle* h, const uint8_t* desc, int32_t n, int32_t levelsup, uint32_t* word_of_feat, uint32_t* node_of_feat,
                      double* weight_of_feat, uint32_t* bow_id, double* bow_val, int32_t* bow_n, uint32_t* fv_node_id,
                      int32_t* fv_node_ptr, uint32_t* fv_feat, int32_t* fv_n_nodes) {
  return voc_transform_any(h, desc, nullptr, n, levelsup, word_of_feat, node_of_feat, weight_of_feat, bow_id, bow_val, bow_n, fv_node_id,
                           fv_node_ptr, fv_feat, fv_n_nodes);
}

int ccm_bow_assemble(int32_t scoring, int32_t weighting, int32_t n, const uint32_t* word_of_feat, const double* weight_of_feat,
                     const uint32_t* node_of_feat, uint32_t* bow_id, double* bow_val, int32_t* bow_n, uint32_t* fv_node_id,
                     int32_t* fv_node_ptr, uint32_t* fv_feat, int32_t* fv_n_nodes) {
  return guarded([&] {
    assemble(scoring, weighting, n, word_of_feat, weight_of_feat, node_of_feat, bow_id, bow_val, bow_n, fv_node_id, fv_node_ptr, fv_feat,
             fv_n_nodes);
  });
}

void ccm_voc_destroy(ccm_voc_handle* h) { delete h; }

}  // extern "C"

namespace ccm {
int voc_transform_resident(ccm_voc_handle* h, const void* d_desc, int32_t n, int32_t levelsup, uint32_t* word_of_feat, uint32_t* node_of_feat,
                           double* weight_of_feat, uint32_t* bow_id, double* bow_val, int32_t* bow_n, uint32_t* fv_node_id,
                           int32_t* fv_node_ptr, uint32_t* fv_feat, int32_t* fv_n_nodes) {
  return voc_transform_any(h, nullptr, static_cast<const uint4*>(d_desc), n, levelsup, word_of_feat, node_of_feat, weight_of_feat, bow_id, bow_val,
                           bow_n, fv_node_id, fv_node_ptr, fv_feat, fv_n_nodes);
}
}  // namespace ccm

// ccm_voc_double.cpp — link-time stand-in for the DEVICE half of the vocabulary transform (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// ccm_voc_create / ccm_voc_transform for shim/ORBVocabulary_shim.cpp in a container without a GPU: the tree descent of every
// descriptor (k_voc_descend in the product) comes from the CPU oracle, the two containers from the library's OWN host half
// ccm_bow_assemble — the same split the product makes.  Linked only into oracle/_ref/libvoc_shim.so.
#include <cstdint>
#include <vector>

#include "ccm_b200.h"
#include "oracle.h"

struct orc_voc;
extern "C" {
orc_voc* orc_voc_create(int32_t k, int32_t L, int32_t scoring, int32_t weighting, int32_t n_nodes, const int32_t* parent, const uint8_t* is_leaf,
                        const uint8_t* desc, const double* weight);
void orc_voc_destroy(orc_voc* v);
int32_t orc_voc_words(const orc_voc* v);
int orc_voc_transform(const orc_voc* v, const uint8_t* feat, int32_t n, int32_t levelsup, uint32_t* word_of_feat, uint32_t* node_of_feat,
                      double* weight_of_feat, uint32_t* bow_id, double* bow_val, int32_t* bow_n, uint32_t* fv_node_id, int32_t* fv_node_ptr,
                      uint32_t* fv_feat, int32_t* fv_n_nodes);
}

struct ccm_voc_handle { orc_voc* v; int scoring, weighting; };

extern "C" {
int ccm_voc_create(int32_t k, int32_t L, int32_t scoring, int32_t weighting, int32_t n_nodes, const int32_t* parent, const uint8_t* is_leaf,
                   const uint8_t* desc, const double* weight, ccm_voc_handle** out) {
  *out = new ccm_voc_handle{orc_voc_create(k, L, scoring, weighting, n_nodes, parent, is_leaf, desc, weight), scoring, weighting};
  return CCM_OK;
}
int ccm_voc_words(const ccm_voc_handle* h) { return orc_voc_words(h->v); }
void ccm_voc_destroy(ccm_voc_handle* h) { if (h) { orc_voc_destroy(h->v); delete h; } }
int ccm_voc_transform(ccm_voc_handle* h, const uint8_t* desc, int32_t n, int32_t levelsup, uint32_t* word_of_feat, uint32_t* node_of_feat,
                      double* weight_of_feat, uint32_t* bow_id, double* bow_val, int32_t* bow_n, uint32_t* fv_node_id, int32_t* fv_node_ptr,
                      uint32_t* fv_feat, int32_t* fv_n_nodes) {
  std::vector<uint32_t> word(n), node(n), t_id(n), t_nid(n), t_ff(n);
  std::vector<double> w(n), t_val(n);
  std::vector<int32_t> t_ptr(n + 1);
  int32_t tb = 0, tf = 0;
  orc_voc_transform(h->v, desc, n, levelsup, word.data(), node.data(), w.data(), t_id.data(), t_val.data(), &tb, t_nid.data(), t_ptr.data(), t_ff.data(), &tf);
  for (int i = 0; i < n; i++) {
    if (word_of_feat) word_of_feat[i] = word[i];
    if (node_of_feat) node_of_feat[i] = node[i];
    if (weight_of_feat) weight_of_feat[i] = w[i];
  }
  return ccm_bow_assemble(h->scoring, h->weighting, n, word.data(), w.data(), node.data(), bow_id, bow_val, bow_n, fv_node_id, fv_node_ptr, fv_feat,
                          fv_n_nodes);
}
}

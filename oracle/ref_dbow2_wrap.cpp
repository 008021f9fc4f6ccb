// ref_dbow2_wrap.cpp — C entry points over the REFERENCE's own DBoW2 sources (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Built by oracle/Makefile into oracle/_ref/libdbow2_ref.so from cslam/thirdparty/DBoW2/{DBoW2/BowVector.cpp, FeatureVector.cpp,
// ScoringObject.cpp, FORB.cpp, TemplatedVocabulary.h, DUtils/Random.cpp, Timestamp.cpp} where they lie under /root/reference, against
// the stand-in OpenCV header oracle/ref_stub/.  Nothing of the reference is copied into this repository; this file only calls it:
//   loadFromTextFile (TemplatedVocabulary.h:1338-1422), transform x2 (:1127-1192, :1219-1260), FORB::distance (FORB.cpp:77-100).
// tests/test_oracle_vs_reference_dbow2.py holds oracle/bow_oracle.cpp to it.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

namespace {
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Base;
struct Voc : Base {
  using Base::transform;   // the single-descriptor overload is protected in the reference
};
cv::Mat as_mat(const uint8_t* d) {
  cv::Mat m(1, 32, CV_8U);
  memcpy(m.ptr<unsigned char>(), d, 32);
  return m;
}
}  // namespace

extern "C" {

void* ref_voc_load(const char* text_file) {
  Voc* v = new Voc();
  if (!v->loadFromTextFile(text_file)) { delete v; return nullptr; }
  return v;
}
void ref_voc_free(void* v) { delete static_cast<Voc*>(v); }
int ref_voc_words(void* v) { return (int)static_cast<Voc*>(v)->size(); }

int ref_forb_distance(const uint8_t* a, const uint8_t* b) { return DBoW2::FORB::distance(as_mat(a), as_mat(b)); }

// same output layout as orc_voc_transform (oracle/bow_oracle.cpp)
int ref_voc_transform(void* vv, const uint8_t* feat, int32_t n, int32_t levelsup, uint32_t* word_of_feat, uint32_t* node_of_feat,
                      double* weight_of_feat, uint32_t* bow_id, double* bow_val, int32_t* bow_n, uint32_t* fv_node_id,
                      int32_t* fv_node_ptr, uint32_t* fv_feat, int32_t* fv_n_nodes) {
  Voc* v = static_cast<Voc*>(vv);
  std::vector<cv::Mat> f(n);
  for (int i = 0; i < n; i++) f[i] = as_mat(feat + 32 * (size_t)i);
  for (int i = 0; i < n; i++) {
    DBoW2::WordId id = 0; DBoW2::WordValue w = 0; DBoW2::NodeId nid = 0;
    v->transform(f[i], id, w, &nid, levelsup);
    word_of_feat[i] = id; node_of_feat[i] = nid; weight_of_feat[i] = w;
  }
  DBoW2::BowVector bv; DBoW2::FeatureVector fv;
  v->transform(f, bv, fv, levelsup);
  int b = 0;
  for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++b) { bow_id[b] = it->first; bow_val[b] = it->second; }
  *bow_n = b;
  int nn = 0, pos = 0;
  fv_node_ptr[0] = 0;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
    fv_node_id[nn] = it->first;
    for (size_t k = 0; k < it->second.size(); k++) fv_feat[pos++] = it->second[k];
    fv_node_ptr[++nn] = pos;
  }
  *fv_n_nodes = nn;
  return 0;
}

}  // extern "C"

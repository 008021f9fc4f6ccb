// ref_voc_shim_wrap.cpp — runs shim/ORBVocabulary_shim.cpp's Frame::ComputeBoW / KeyFrame::ComputeBoW next to the reference's own
// ORBVocabulary::transform on the same vocabulary file and descriptors (TEST INFRASTRUCTURE, NOT PRODUCT).
#include <cslam/Frame.h>
#include <cslam/KeyFrame.h>

#include <cstdint>
#include <cstring>
#include <string>

namespace cslam { void ccm_b200_load_vocabulary(const ORBVocabulary* voc, const std::string& file); }   // declared by the integrator

using namespace cslam;

namespace {
void flatten(const DBoW2::BowVector& bv, const DBoW2::FeatureVector& fv, uint32_t* bow_id, double* bow_val, int32_t* bow_n, uint32_t* fv_node_id,
             int32_t* fv_node_ptr, uint32_t* fv_feat, int32_t* fv_n) {
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
  *fv_n = nn;
}
}  // namespace

extern "C" {

void* vshim_load(const char* text_file) {
  try {
    vocptr* v = new vocptr(new ORBVocabulary());
    if (!(*v)->loadFromTextFile(text_file)) { delete v; return 0; }   // the reference's loader (ClientSystem.cpp:77)
    ccm_b200_load_vocabulary(v->get(), text_file);                     // the line the integrator adds after it
    return v;
  } catch (...) { return 0; }
}
void vshim_free(void* v) { delete static_cast<vocptr*>(v); }

/* side 0: the shim's Frame::ComputeBoW, 1: the shim's KeyFrame::ComputeBoW, 2: the reference's body of both —
 * mpORBvocabulary->transform(Converter::toDescriptorVector(mDescriptors), mBowVec, mFeatVec, 4) */
int vshim_compute_bow(void* vv, int side, const uint8_t* desc, int32_t n, uint32_t* bow_id, double* bow_val, int32_t* bow_n, uint32_t* fv_node_id,
                      int32_t* fv_node_ptr, uint32_t* fv_feat, int32_t* fv_n) {
  try {
    vocptr voc = *static_cast<vocptr*>(vv);
    cv::Mat D;
    if (n > 0) { D.create(n, 32, CV_8U); std::memcpy(D.ptr<uchar>(0), desc, 32 * (size_t)n); }
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    if (side == 0) { Frame F; F.mpORBvocabulary = voc; F.mDescriptors = D; F.ComputeBoW(); F.ComputeBoW(); bv = F.mBowVec; fv = F.mFeatVec; }
    else if (side == 1) { KeyFrame K; K.mpORBvocabulary = voc; K.mDescriptors = D; K.ComputeBoW(); bv = K.mBowVec; fv = K.mFeatVec; }
    else {
      std::vector<cv::Mat> vCurrentDesc;                               // Converter::toDescriptorVector (S/Converter.cc:30-38)
      vCurrentDesc.reserve(D.rows);
      for (int j = 0; j < D.rows; j++) vCurrentDesc.push_back(D.row(j));
      voc->transform(vCurrentDesc, bv, fv, 4);
    }
    flatten(bv, fv, bow_id, bow_val, bow_n, fv_node_id, fv_node_ptr, fv_feat, fv_n);
    return 0;
  } catch (...) { return -1; }
}

}  // extern "C"

// ORBVocabulary_shim.cpp — DBoW2's descriptor -> (BowVector, FeatureVector) transform on top of libccm_b200.so
// (SURVEY.md §8(f) rank 2).  Replaces the bodies of Frame::ComputeBoW (cslam/src/Frame.cpp:268-275) and
// KeyFrame::ComputeBoW (cslam/src/KeyFrame.cpp:277-286), i.e. the call
//     mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4);
// (thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1192).  Headers stay byte-identical; the DBoW2 containers remain the
// reference's own types.  Compiled against the reference's own ORBVocabulary.h / DBoW2 and run next to ORBVocabulary::transform by
// tests/test_shim_vocabulary.py (oracle/Makefile: _ref/libvoc_shim.so).
//
// The device copy of the vocabulary is made from the same text file the reference loads
// (ClientSystem.cpp:77, ServerSystem.cpp:165: mpVoc->loadFromTextFile(strVocFile)): call ccm_b200_load_vocabulary(mpVoc.get(),
// strVocFile) right after that line.  ORBVocabulary keeps serving score() and the other host-side queries.
#include <cslam/Frame.h>
#include <cslam/KeyFrame.h>

#include <fstream>
#include <map>
#include <mutex>
#include <sstream>

#include "ccm_b200.h"

namespace cslam {

namespace {
std::mutex g_mu;
std::map<const ORBVocabulary*, ccm_voc_handle*> g_voc;   // one device tree per vocabulary object

ccm_voc_handle* handle_of(const ORBVocabulary* v) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_voc.find(v);
  if (it == g_voc.end()) throw estd::infrastructure_ex();   // ccm_b200_load_vocabulary was not called for this vocabulary
  return it->second;
}

void transform_b200(const ORBVocabulary* voc, const cv::Mat& descriptors, DBoW2::BowVector& bow, DBoW2::FeatureVector& fv, int levelsup) {
  bow.clear(); fv.clear();
  const int n = descriptors.rows;
  if (n == 0) return;
  cv::Mat d = descriptors.isContinuous() ? descriptors : descriptors.clone();
  std::vector<uint32_t> bid(n), fid(n), ff(n);
  std::vector<double> bval(n);
  std::vector<int32_t> fptr(n + 1);
  int32_t bn = 0, fn = 0;
  if (ccm_voc_transform(handle_of(voc), d.ptr<uchar>(0), n, levelsup, nullptr, nullptr, nullptr, bid.data(), bval.data(), &bn, fid.data(),
                        fptr.data(), ff.data(), &fn) != CCM_OK)
    throw estd::infrastructure_ex();
  for (int i = 0; i < bn; i++) bow.insert(bow.end(), DBoW2::BowVector::value_type(bid[i], bval[i]));          // ids arrive ascending
  for (int a = 0; a < fn; a++)
    fv.insert(fv.end(), DBoW2::FeatureVector::value_type(fid[a], std::vector<unsigned int>(ff.begin() + fptr[a], ff.begin() + fptr[a + 1])));
}
}  // namespace

// rows of the text file (loadFromTextFile, TemplatedVocabulary.h:1338-1422) -> ccm_voc_create
void ccm_b200_load_vocabulary(const ORBVocabulary* voc, const std::string& file) {
  std::ifstream f(file.c_str());
  std::string line;
  if (!std::getline(f, line)) throw estd::infrastructure_ex();
  int k, L, scoring, weighting;
  { std::stringstream ss(line); ss >> k >> L >> scoring >> weighting; }
  std::vector<int32_t> parent(1, 0);
  std::vector<uint8_t> leaf(1, 0), desc(32, 0);
  std::vector<double> weight(1, 0.0);
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    std::stringstream ss(line);
    int pid, isleaf; ss >> pid >> isleaf;
    parent.push_back(pid); leaf.push_back(isleaf > 0);
    for (int i = 0; i < 32; i++) { int b; ss >> b; desc.push_back((uint8_t)b); }
    double w; ss >> w; weight.push_back(w);
  }
  ccm_voc_handle* h = nullptr;
  if (ccm_voc_create(k, L, scoring, weighting, (int32_t)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), &h) != CCM_OK)
    throw estd::infrastructure_ex();
  std::lock_guard<std::mutex> lk(g_mu);
  g_voc[voc] = h;
}

void Frame::ComputeBoW() {                       // S/Frame.cpp:268-275
  if (mBowVec.empty()) transform_b200(mpORBvocabulary.get(), mDescriptors, mBowVec, mFeatVec, 4);
}

void KeyFrame::ComputeBoW() {                    // S/KeyFrame.cpp:277-286
  if (mBowVec.empty() || mFeatVec.empty()) transform_b200(mpORBvocabulary.get(), mDescriptors, mBowVec, mFeatVec, 4);
}

}  // namespace cslam

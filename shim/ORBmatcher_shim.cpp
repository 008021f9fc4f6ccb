// ORBmatcher_shim.cpp — with ORBmatcher_proj_shim.cpp, replaces cslam/src/ORBmatcher.cpp as a whole: SearchByBoW x2,
// SearchForTriangulation and DescriptorDistance here, the projection-guided searches there, plus the few small members the rest of
// the class needs (constructor, the three thresholds, RadiusByViewingCos).  cslam/include/cslam/ORBmatcher.h stays byte-identical.
// Type-checked against that header and run against the reference's own ORBmatcher.cpp on stand-in Frame / KeyFrame / MapPoint
// classes by tests/test_shim_dropin.py (oracle/Makefile: _ref/libmatch_shim.so).
#include <cslam/ORBmatcher.h>

#include "ccm_b200.h"

namespace cslam {

namespace {
struct FlatFV {                                      // DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned>>
  std::vector<uint32_t> node_id, feat; std::vector<int32_t> node_ptr;
  explicit FlatFV(const DBoW2::FeatureVector& fv) {
    node_ptr.push_back(0);
    for (auto& kv : fv) { node_id.push_back(kv.first); feat.insert(feat.end(), kv.second.begin(), kv.second.end()); node_ptr.push_back((int32_t)feat.size()); }
  }
  ccm_feature_vector c() const { return ccm_feature_vector{(int32_t)node_id.size(), node_id.data(), node_ptr.data(), feat.data()}; }
};
std::vector<float> angles(const std::vector<cv::KeyPoint>& k) { std::vector<float> a(k.size()); for (size_t i = 0; i < k.size(); i++) a[i] = k[i].angle; return a; }
std::vector<uint8_t> good_mps(const std::vector<boost::shared_ptr<MapPoint>>& v) { std::vector<uint8_t> g(v.size()); for (size_t i = 0; i < v.size(); i++) g[i] = v[i] && !v[i]->isBad(); return g; }
}

const int ORBmatcher::TH_HIGH = 100;      // S/ORBmatcher.cpp:63-65
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

float ORBmatcher::RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }   // S/ORBmatcher.cpp:150-156

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
  uint16_t d = 0;
  ccm_hamming_matrix(a.ptr<uchar>(0), 1, b.ptr<uchar>(0), 1, &d);   // batch callers should use the matrix form
  return d;
}

int ORBmatcher::SearchByBoW(kfptr pKF, Frame& F, std::vector<mpptr>& vpMapPointMatches) {   // S/ORBmatcher.cpp:178-306
  const std::vector<mpptr> vpMapPointsKF = pKF->GetMapPointMatches();
  vpMapPointMatches = std::vector<mpptr>(F.N, nullptr);
  FlatFV fk(pKF->mFeatVec), ff(F.mFeatVec);
  ccm_feature_vector cfk = fk.c(), cff = ff.c();
  std::vector<uint8_t> has = good_mps(vpMapPointsKF);
  std::vector<float> ak = angles(pKF->mvKeysUn), af = angles(F.mvKeys);
  std::vector<int32_t> match(F.N);
  int32_t n = 0;
  if (ccm_match_bow_kf_frame(pKF->mDescriptors.ptr<uchar>(0), pKF->mDescriptors.rows, has.data(), ak.data(), &cfk,
                             F.mDescriptors.ptr<uchar>(0), F.N, af.data(), &cff, mfNNratio, mbCheckOrientation, match.data(), &n) != CCM_OK)
    throw estd::infrastructure_ex();
  for (int j = 0; j < F.N; j++) if (match[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[match[j]];
  return n;
}

int ORBmatcher::SearchByBoW(kfptr pKF1, kfptr pKF2, std::vector<mpptr>& vpMatches12) {      // S/ORBmatcher.cpp:565-698
  const std::vector<mpptr> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
  vpMatches12 = std::vector<mpptr>(mp1.size(), nullptr);
  FlatFV f1(pKF1->mFeatVec), f2(pKF2->mFeatVec);
  ccm_feature_vector c1 = f1.c(), c2 = f2.c();
  std::vector<uint8_t> h1 = good_mps(mp1), h2 = good_mps(mp2);
  std::vector<float> a1 = angles(pKF1->mvKeysUn), a2 = angles(pKF2->mvKeysUn);
  std::vector<int32_t> m(mp1.size());
  int32_t n = 0;
  if (ccm_match_bow_kf_kf(pKF1->mDescriptors.ptr<uchar>(0), (int)mp1.size(), h1.data(), a1.data(), &c1, pKF2->mDescriptors.ptr<uchar>(0),
                          (int)mp2.size(), h2.data(), a2.data(), &c2, mfNNratio, mbCheckOrientation, m.data(), &n) != CCM_OK)
    throw estd::infrastructure_ex();
  for (size_t i = 0; i < m.size(); i++) if (m[i] >= 0) vpMatches12[i] = mp2[m[i]];
  return n;
}

int ORBmatcher::SearchForTriangulation(kfptr pKF1, kfptr pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t>>& vMatchedPairs) {  // :700-852
  cv::Mat Cw = pKF1->GetCameraCenter(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
  cv::Mat C2 = R2w * Cw + t2w;                                   // epipole in the second image, :704-712
  const float invz = 1.0f / C2.at<float>(2);
  const float ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx, ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;
  auto view = [](kfptr k, FlatFV& fv, ccm_feature_vector& cfv, std::vector<uint8_t>& has, std::vector<float>& xy, std::vector<int32_t>& oct, std::vector<float>& ang) {
    has.resize(k->N); xy.resize(2 * k->N); oct.resize(k->N); ang.resize(k->N);
    for (int i = 0; i < k->N; i++) { has[i] = k->GetMapPoint(i) ? 1 : 0; xy[2 * i] = k->mvKeysUn[i].pt.x; xy[2 * i + 1] = k->mvKeysUn[i].pt.y; oct[i] = k->mvKeysUn[i].octave; ang[i] = k->mvKeysUn[i].angle; }
    cfv = fv.c();
    return ccm_tri_view{k->mDescriptors.ptr<uchar>(0), k->N, has.data(), xy.data(), oct.data(), ang.data(), &cfv, k->fx, k->fy, k->cx, k->cy};
  };
  FlatFV f1(pKF1->mFeatVec), f2(pKF2->mFeatVec); ccm_feature_vector c1, c2;
  std::vector<uint8_t> h1, h2; std::vector<float> xy1, xy2, a1, a2; std::vector<int32_t> o1, o2;
  ccm_tri_view v1 = view(pKF1, f1, c1, h1, xy1, o1, a1), v2 = view(pKF2, f2, c2, h2, xy2, o2, a2);
  cv::Mat F = F12.isContinuous() ? F12 : F12.clone();
  std::vector<int32_t> pairs(2 * (size_t)std::min(pKF1->N, pKF2->N) + 2);
  int32_t np = 0;
  if (ccm_match_triangulation(&v1, &v2, F.ptr<float>(0), ex, ey, pKF2->mvLevelSigma2.data(), pKF2->mvScaleFactors.data(),
                              (int)pKF2->mvScaleFactors.size(), mbCheckOrientation, pairs.data(), &np) != CCM_OK)
    throw estd::infrastructure_ex();
  vMatchedPairs.clear(); vMatchedPairs.reserve(np);
  for (int i = 0; i < np; i++) vMatchedPairs.push_back(std::make_pair((size_t)pairs[2 * i], (size_t)pairs[2 * i + 1]));
  return np;
}

}  // namespace cslam

// ref_map_update_wrap.cpp — drives shim/MapUpdate_shim.cpp on a stand-in map, next to a restatement of the loop it replaces
// (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// The loop lives inside Map::RunGBA (cslam/src/Map.cpp:1441-1570) and MapMerger::RunGBA (cslam/src/MapMerger.cpp:637-753), two
// member functions that cannot be compiled apart from the ROS-facing rest of their classes; reference_loop() below restates it
// on the stand-in classes of ref_stub_opt with cv::Mat expressions of the same shape (the stand-in Mat rounds small products as
// cv::gemm does — ref_stub/opencv2/core/core.hpp — so values, the set of touched objects and the setter calls are all compared exactly).  mode 1 runs cslam::UpdateMapAfterGBA (the shim; ccm_gba_map_update doubled by the oracle in this library).
#include <cslam/KeyFrame.h>
#include <cslam/Map.h>
#include <cslam/MapPoint.h>

#include <cmath>
#include <cstring>
#include <list>

namespace cslam {
void UpdateMapAfterGBA(boost::shared_ptr<Map> pMap, idpair nLoopKF);
std::mutex MapPoint::mGlobalMutex;
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
}

using namespace cslam;
typedef boost::shared_ptr<KeyFrame> kfptr;
typedef boost::shared_ptr<MapPoint> mpptr;

static void reference_loop(boost::shared_ptr<Map> pMap, idpair nLoopKF) {
  std::list<kfptr> lpKFtoCheck(pMap->mvpKeyFrameOrigins.begin(), pMap->mvpKeyFrameOrigins.end());
  while (!lpKFtoCheck.empty()) {
    kfptr pKF = lpKFtoCheck.front();
    const std::set<kfptr> sChilds = pKF->GetChilds();
    cv::Mat Twc = pKF->GetPoseInverse();
    for (std::set<kfptr>::const_iterator sit = sChilds.begin(); sit != sChilds.end(); sit++) {
      kfptr pChild = *sit;
      if (pChild->mBAGlobalForKF != nLoopKF) {
        cv::Mat Tchildc = pChild->GetPose() * Twc;
        pChild->mTcwGBA = Tchildc * pKF->mTcwGBA;
        pChild->mBAGlobalForKF = nLoopKF;
      }
      lpKFtoCheck.push_back(pChild);
    }
    pKF->mTcwBefGBA = pKF->GetPose();
    pKF->SetPose(pKF->mTcwGBA, true);
    pKF->mbLoopCorrected = true;
    lpKFtoCheck.pop_front();
  }
  const std::vector<mpptr> vpMPs = pMap->GetAllMapPoints();
  for (size_t i = 0; i < vpMPs.size(); i++) {
    mpptr pMP = vpMPs[i];
    if (pMP->isBad()) continue;
    if (pMP->mBAGlobalForKF == nLoopKF) {
      pMP->SetWorldPos(pMP->mPosGBA, true);
      pMP->mbLoopCorrected = true;
    } else {
      kfptr pRefKF = pMP->GetReferenceKeyFrame();
      if (!pRefKF) continue;
      if (pRefKF->mBAGlobalForKF != nLoopKF) continue;
      if (pRefKF->mTcwBefGBA.empty()) continue;          // flagged by the BA but outside the tree: the reference would read an empty Mat here
      cv::Mat Rcw = pRefKF->mTcwBefGBA.rowRange(0, 3).colRange(0, 3);
      cv::Mat tcw = pRefKF->mTcwBefGBA.rowRange(0, 3).col(3);
      cv::Mat Xc = Rcw * pMP->GetWorldPos() + tcw;
      cv::Mat Twc = pRefKF->GetPoseInverse();
      cv::Mat Rwc = Twc.rowRange(0, 3).colRange(0, 3);
      cv::Mat twc = Twc.rowRange(0, 3).col(3);
      pMP->SetWorldPos(Rwc * Xc + twc, true);
      pMP->mbLoopCorrected = true;
    }
  }
}

static cv::Mat mat_from(const float* p, int r, int c) {
  cv::Mat m(r, c, CV_32F);
  for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) m.at<float>(i, j) = p[i * c + j];
  return m;
}
static void mat_to(const cv::Mat& m, float* p, int r, int c) {
  for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) p[i * c + j] = m.empty() ? NAN : m.at<float>(i, j);
}

// scene: ccm_slam_b200.synth.make_map_update layout.  Outputs per keyframe: pose after, mTcwBefGBA, mTcwGBA (NaN where empty),
// info = {mbLoopCorrected, SetPose calls, mBAGlobalForKF == nLoopKF}; per point: position after, info = {mbLoopCorrected, SetWorldPos calls}.
extern "C" int mapw_update(int mode, int32_t K, const int32_t* kf_parent, const uint8_t* kf_optimized, const float* kf_Tcw, const float* kf_TcwGBA,
                           int32_t P, const uint8_t* mp_state, const int32_t* mp_ref, const float* mp_pos, const float* mp_pos_gba,
                           float* kf_pose_after, float* kf_bef, float* kf_gba, int32_t* kf_info, float* mp_after, int32_t* mp_info) {
  const idpair nLoopKF(77, 3);
  boost::shared_ptr<Map> map(new Map);
  for (int k = 0; k < K; k++) {
    kfptr kf(new KeyFrame);
    kf->mId = idpair(k, 0); kf->mUniqueId = k;
    kf->Tcw = mat_from(kf_Tcw + 16 * (size_t)k, 4, 4);
    if (kf_optimized[k]) { kf->mBAGlobalForKF = nLoopKF; kf->mTcwGBA = mat_from(kf_TcwGBA + 16 * (size_t)k, 4, 4); }
    else kf->mBAGlobalForKF = idpair(5, 1);
    map->kfs.push_back(kf);
  }
  for (int k = 0; k < K; k++) {
    if (kf_parent[k] == -1) map->mvpKeyFrameOrigins.push_back(map->kfs[k]);
    else if (kf_parent[k] >= 0) { map->kfs[k]->mpParent = map->kfs[kf_parent[k]]; map->kfs[kf_parent[k]]->mspChildrens.insert(map->kfs[k]); }
  }
  for (int i = 0; i < P; i++) {
    mpptr mp(new MapPoint);
    mp->mId = idpair(i, 0); mp->mUniqueId = i;
    mp->mWorldPos = mat_from(mp_pos + 3 * (size_t)i, 3, 1);
    mp->mbBad = mp_state[i] == 0;
    if (mp_state[i] == 1) { mp->mBAGlobalForKF = nLoopKF; mp->mPosGBA = mat_from(mp_pos_gba + 3 * (size_t)i, 3, 1); }
    else mp->mBAGlobalForKF = idpair(5, 1);
    if (mp_ref[i] >= 0) mp->mpRefKF = map->kfs[mp_ref[i]];
    map->mps.push_back(mp);
  }
  try {
    if (mode == 0) reference_loop(map, nLoopKF); else UpdateMapAfterGBA(map, nLoopKF);
  } catch (const std::exception& e) {
    return 1;
  }
  for (int k = 0; k < K; k++) {
    const kfptr& kf = map->kfs[k];
    mat_to(kf->Tcw, kf_pose_after + 16 * (size_t)k, 4, 4);
    mat_to(kf->mTcwBefGBA, kf_bef + 16 * (size_t)k, 4, 4);
    mat_to(kf->mTcwGBA, kf_gba + 16 * (size_t)k, 4, 4);
    kf_info[3 * k] = kf->mbLoopCorrected; kf_info[3 * k + 1] = kf->n_set_pose; kf_info[3 * k + 2] = kf->mBAGlobalForKF == nLoopKF;
  }
  for (int i = 0; i < P; i++) {
    mat_to(map->mps[i]->mWorldPos, mp_after + 3 * (size_t)i, 3, 1);
    mp_info[2 * i] = map->mps[i]->mbLoopCorrected; mp_info[2 * i + 1] = map->mps[i]->n_set_pos;
  }
  return 0;
}

// MapUpdate_shim.cpp — reference-side translation unit for the map update that follows a global BA.
//
// Map::RunGBA (cslam/src/Map.cpp:1436-1573) and MapMerger::RunGBA (cslam/src/MapMerger.cpp:630-756) both carry the same loop after
// Optimizer::MapFusionGBA returns: walk the spanning tree from the map origins handing mTcwGBA down to keyframes the BA did not
// hold, SetPose every keyframe met, then correct every map point.  A maintainer replaces that loop (Map.cpp:1441-1570,
// MapMerger.cpp:637-753) by one call of UpdateMapAfterGBA(pMap, nLoopKF) — see INTEGRATION.md.  The pointer graph is flattened once,
// libccm_b200.so does the arithmetic (ccm_gba_map_update: tree pass on the host, point pass on the GPU), and the write-back touches
// exactly the objects the reference touches, through the same setters, in the same order.
//
// In this repository it is compiled against stand-in Map / KeyFrame / MapPoint classes (oracle/ref_stub_opt) and run next to a
// literal restatement of the reference loop by tests/test_shim_map_update.py.
#include <cslam/KeyFrame.h>
#include <cslam/Map.h>
#include <cslam/MapPoint.h>

#include <list>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "ccm_b200.h"

namespace cslam {

void UpdateMapAfterGBA(boost::shared_ptr<Map> pMap, idpair nLoopKF) {
  typedef boost::shared_ptr<KeyFrame> kfptr;
  typedef boost::shared_ptr<MapPoint> mpptr;

  // keyframes in the reference's visiting order: breadth-first from the origins through GetChilds()
  std::vector<kfptr> kfs;
  std::vector<int32_t> parent;
  std::unordered_map<KeyFrame*, int> row;
  for (std::vector<kfptr>::const_iterator it = pMap->mvpKeyFrameOrigins.begin(); it != pMap->mvpKeyFrameOrigins.end(); ++it) {
    row[it->get()] = (int)kfs.size(); kfs.push_back(*it); parent.push_back(-1);
  }
  for (size_t q = 0; q < kfs.size(); q++) {
    const std::set<kfptr> sChilds = kfs[q]->GetChilds();
    for (std::set<kfptr>::const_iterator sit = sChilds.begin(); sit != sChilds.end(); ++sit) {
      if (row.count(sit->get())) throw std::runtime_error("UpdateMapAfterGBA: the spanning tree visits a keyframe twice");   // the reference would not terminate on a cycle
      row[sit->get()] = (int)kfs.size(); kfs.push_back(*sit); parent.push_back((int32_t)q);
    }
  }
  const int n_tree = (int)kfs.size();

  // map points; reference keyframes outside the tree are appended as rows the update leaves alone
  const std::vector<mpptr> vpMPs = pMap->GetAllMapPoints();
  const int P = (int)vpMPs.size();
  std::vector<uint8_t> state(P, 0);
  std::vector<int32_t> ref(P, -1);
  std::vector<float> pos((size_t)P * 3, 0.f), pos_gba((size_t)P * 3, 0.f);
  for (int i = 0; i < P; i++) {
    const mpptr& pMP = vpMPs[i];
    if (pMP->isBad()) continue;
    const cv::Mat X = pMP->GetWorldPos();
    for (int j = 0; j < 3; j++) pos[3 * (size_t)i + j] = X.at<float>(j);
    if (pMP->mBAGlobalForKF == nLoopKF) {
      state[i] = 1;
      for (int j = 0; j < 3; j++) pos_gba[3 * (size_t)i + j] = pMP->mPosGBA.at<float>(j);
      continue;
    }
    state[i] = 2;
    kfptr pRefKF = pMP->GetReferenceKeyFrame();
    if (!pRefKF) continue;
    std::unordered_map<KeyFrame*, int>::const_iterator f = row.find(pRefKF.get());
    if (f == row.end()) { row[pRefKF.get()] = (int)kfs.size(); ref[i] = (int)kfs.size(); kfs.push_back(pRefKF); parent.push_back(-2); }
    else ref[i] = f->second;
  }

  const int K = (int)kfs.size();
  std::vector<uint8_t> optimized(K, 0), visited(K, 0), corrected(P, 0);
  std::vector<float> Tcw((size_t)K * 16, 0.f), TcwGBA((size_t)K * 16, 0.f), out((size_t)P * 3, 0.f);
  for (int k = 0; k < K; k++) {
    const cv::Mat T = kfs[k]->GetPose();
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Tcw[16 * (size_t)k + 4 * r + c] = T.at<float>(r, c);
    optimized[k] = kfs[k]->mBAGlobalForKF == nLoopKF;
    if (optimized[k] && k < n_tree)
      for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) TcwGBA[16 * (size_t)k + 4 * r + c] = kfs[k]->mTcwGBA.at<float>(r, c);
  }

  const int rc = ccm_gba_map_update(K, parent.data(), optimized.data(), Tcw.data(), TcwGBA.data(), visited.data(), P, state.data(), ref.data(),
                                    pos.data(), pos_gba.data(), out.data(), corrected.data());
  if (rc != CCM_OK) throw std::runtime_error(std::string("ccm_gba_map_update: ") + ccm_last_error());

  // write-back, keyframes in visiting order (Map.cpp:1455-1490)
  for (int k = 0; k < n_tree; k++) {
    if (!visited[k]) continue;
    const kfptr& pKF = kfs[k];
    if (!optimized[k]) {
      pKF->mTcwGBA.create(4, 4, CV_32F);
      for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) pKF->mTcwGBA.at<float>(r, c) = TcwGBA[16 * (size_t)k + 4 * r + c];
      pKF->mBAGlobalForKF = nLoopKF;
    }
    pKF->mTcwBefGBA = pKF->GetPose();
    pKF->SetPose(pKF->mTcwGBA, true);
    pKF->mbLoopCorrected = true;
  }
  // map points in GetAllMapPoints() order (Map.cpp:1497-1563)
  for (int i = 0; i < P; i++) {
    if (!corrected[i]) continue;
    cv::Mat X(3, 1, CV_32F);
    for (int j = 0; j < 3; j++) X.at<float>(j) = out[3 * (size_t)i + j];
    vpMPs[i]->SetWorldPos(X, true);
    vpMPs[i]->mbLoopCorrected = true;
  }
}

}  // namespace cslam

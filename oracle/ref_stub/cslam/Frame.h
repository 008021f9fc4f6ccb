// Stand-ins for cslam/Frame.h, KeyFrame.h and MapPoint.h (TEST INFRASTRUCTURE, NOT PRODUCT; our own code).
//
// cslam/src/ORBmatcher.cpp reads a few dozen members of Frame, KeyFrame and MapPoint; the real classes drag in ROS messages, the
// communicator, g2o and Eigen and cannot be compiled in this image.  These plain structs expose exactly the members and methods the
// matcher source names (same names, same types), backed by flat arrays that oracle/ref_match_wrap.cpp fills, so that the reference's
// OWN ORBmatcher.cpp compiles where it lies and its eleven search methods run on data the tests control.  What is restated here rather
// than taken from the reference: the feature grid and GetFeaturesInArea (Frame.cpp:103-119, 200-265; KeyFrame.cpp:206-226, 265-275,
// 1162-1201), IsInImage (KeyFrame.cpp:1203), PredictScale and the distance-invariance getters (MapPoint.cpp:823-870).  The map
// surgery methods (AddObservation, AddMapPoint, Replace, RemapMapPointMatch) only record what they were asked to do.
#ifndef CCM_ORACLE_REF_STUB_CSLAM_CLASSES_H
#define CCM_ORACLE_REF_STUB_CSLAM_CLASSES_H
#include <boost/shared_ptr.hpp>
#include <opencv2/opencv.hpp>

#include <cmath>
#include <map>
#include <set>
#include <vector>

#include <cslam/config.h>
#include <cslam/estd.h>
#include <thirdparty/DBoW2/DBoW2/FeatureVector.h>

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 75

namespace cslam {

class Frame;
class KeyFrame;
class MapPoint;

// the lookup grid both image types share
struct FeatureGridStandIn {
  float minX, minY, maxX, maxY, wInv, hInv;
  int cols, rows;
  std::vector<std::vector<size_t> > cell;   // [col * rows + row]
  void build(const std::vector<cv::KeyPoint>& keysUn) {
    cell.assign((size_t)cols * rows, std::vector<size_t>());
    for (size_t i = 0; i < keysUn.size(); i++) {
      const int px = (int)round((keysUn[i].pt.x - minX) * wInv), py = (int)round((keysUn[i].pt.y - minY) * hInv);
      if (px < 0 || px >= cols || py < 0 || py >= rows) continue;
      cell[(size_t)px * rows + py].push_back(i);
    }
  }
  std::vector<size_t> inArea(const std::vector<cv::KeyPoint>& keysUn, float x, float y, float r, int minLevel, int maxLevel) const {
    std::vector<size_t> out;
    const int c0 = std::max(0, (int)floor((x - minX - r) * wInv));
    if (c0 >= cols) return out;
    const int c1 = std::min(cols - 1, (int)ceil((x - minX + r) * wInv));
    if (c1 < 0) return out;
    const int r0 = std::max(0, (int)floor((y - minY - r) * hInv));
    if (r0 >= rows) return out;
    const int r1 = std::min(rows - 1, (int)ceil((y - minY + r) * hInv));
    if (r1 < 0) return out;
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = c0; ix <= c1; ix++)
      for (int iy = r0; iy <= r1; iy++)
        for (size_t k : cell[(size_t)ix * rows + iy]) {
          const cv::KeyPoint& kp = keysUn[k];
          if (check) {
            if (kp.octave < minLevel) continue;
            if (maxLevel >= 0 && kp.octave > maxLevel) continue;
          }
          if (fabs(kp.pt.x - x) < r && fabs(kp.pt.y - y) < r) out.push_back(k);
        }
    return out;
  }
};

class MapPoint {
 public:
  typedef boost::shared_ptr<KeyFrame> kfptr;
  typedef boost::shared_ptr<Frame> frameptr;
  typedef boost::shared_ptr<MapPoint> mpptr;
  int tag = -1;                 // index in the caller's arrays
  bool bad = false;
  cv::Mat pos, normal, desc;    // 3x1 f32, 3x1 f32, 1x32 u8
  float mfMinDistance = 0, mfMaxDistance = 0;
  int nObs = 0;
  std::map<const KeyFrame*, int> indexIn;   // GetIndexInKeyFrame
  bool mbDoNotReplace = false;
  // tracking fields written by Frame::isInFrustum in the reference, here by the caller
  bool mbTrackInView = false;
  float mTrackProjX = 0, mTrackProjY = 0, mTrackViewCos = 0;
  int mnTrackScaleLevel = 0;
  // what the matcher asked for
  std::vector<std::pair<const KeyFrame*, size_t> > added;
  mpptr replacedBy;

  bool isBad() { return bad; }
  cv::Mat GetWorldPos() { return pos.clone(); }
  cv::Mat GetNormal() { return normal.clone(); }
  cv::Mat GetDescriptor() { return desc.clone(); }
  float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
  float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
  int Observations() { return nObs; }
  int GetIndexInKeyFrame(kfptr pKF, bool = false);
  bool IsInKeyFrame(kfptr pKF);
  void AddObservation(kfptr pKF, size_t idx, bool = false);
  void Replace(mpptr pMP, bool = false) { replacedBy = pMP; bad = true; }
  int PredictScale(const float& currentDist, kfptr pKF);
  int PredictScale(const float& currentDist, frameptr pF);
};

class Frame : public boost::enable_shared_from_this<Frame> {
 public:
  typedef boost::shared_ptr<MapPoint> mpptr;
  int N = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
  cv::Mat mDescriptors, mTcw;
  std::vector<mpptr> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  DBoW2::FeatureVector mFeatVec;
  std::vector<float> mvScaleFactors;
  int mnScaleLevels = 8;
  float mfLogScaleFactor = 0;
  static float fx, fy, cx, cy, mnMinX, mnMinY, mnMaxX, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
  FeatureGridStandIn grid;
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const {
    return grid.inArea(mvKeysUn, x, y, r, minLevel, maxLevel);
  }
};

class KeyFrame : public boost::enable_shared_from_this<KeyFrame> {
 public:
  typedef boost::shared_ptr<MapPoint> mpptr;
  int N = 0;
  std::vector<cv::KeyPoint> mvKeysUn;
  cv::Mat mDescriptors, Rcw, tcw, Ow;
  std::vector<mpptr> mvpMapPoints;
  DBoW2::FeatureVector mFeatVec;
  std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  int mnScaleLevels = 8;
  float mfLogScaleFactor = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0;
  int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
  int mnGridCols = 0, mnGridRows = 0;
  float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  FeatureGridStandIn grid;
  // what the matcher asked for
  std::vector<std::pair<mpptr, size_t> > added;
  std::vector<int> remapped;   // triples (map point tag, idx_now, idx_new)

  std::vector<mpptr> GetMapPointMatches() { return mvpMapPoints; }
  std::set<mpptr> GetMapPoints() {
    std::set<mpptr> s;
    for (size_t i = 0; i < mvpMapPoints.size(); i++) if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);
    return s;
  }
  mpptr GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
  cv::Mat GetRotation() { return Rcw.clone(); }
  cv::Mat GetTranslation() { return tcw.clone(); }
  cv::Mat GetCameraCenter() { return Ow.clone(); }
  bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const { return grid.inArea(mvKeysUn, x, y, r, -1, -1); }
  void AddMapPoint(mpptr pMP, const size_t& idx, bool = false) { mvpMapPoints[idx] = pMP; added.push_back(std::make_pair(pMP, idx)); }
  void RemapMapPointMatch(mpptr pMP, const size_t& idx_now, const size_t& idx_new) {
    mvpMapPoints[idx_now] = nullptr; mvpMapPoints[idx_new] = pMP;
    remapped.push_back(pMP->tag); remapped.push_back((int)idx_now); remapped.push_back((int)idx_new);
    pMP->indexIn[this] = (int)idx_new;
  }
};

inline int MapPoint::GetIndexInKeyFrame(kfptr pKF, bool) {
  std::map<const KeyFrame*, int>::const_iterator it = indexIn.find(pKF.get());
  return it == indexIn.end() ? -1 : it->second;
}
inline bool MapPoint::IsInKeyFrame(kfptr pKF) { return indexIn.count(pKF.get()) != 0; }
inline void MapPoint::AddObservation(kfptr pKF, size_t idx, bool) { indexIn[pKF.get()] = (int)idx; nObs++; added.push_back(std::make_pair((const KeyFrame*)pKF.get(), idx)); }
inline int MapPoint::PredictScale(const float& currentDist, kfptr pKF) {      // MapPoint.cpp:837-852
  float ratio = mfMaxDistance / currentDist;
  int nScale = ceil(log(ratio) / pKF->mfLogScaleFactor);
  if (nScale < 0) nScale = 0;
  else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
  return nScale;
}
inline int MapPoint::PredictScale(const float& currentDist, frameptr pF) {    // MapPoint.cpp:854-869
  float ratio = mfMaxDistance / currentDist;
  int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
  if (nScale < 0) nScale = 0;
  else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
  return nScale;
}

}  // namespace cslam
#endif

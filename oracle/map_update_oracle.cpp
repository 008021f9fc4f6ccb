// map_update_oracle.cpp — CPU restatement of the map update that follows a global BA.  TEST INFRASTRUCTURE, NOT PRODUCT: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it.
//
// Follows, statement by statement, the loop that cslam/src/Map.cpp:1441-1570 (Map::RunGBA) and cslam/src/MapMerger.cpp:637-753
// (MapMerger::RunGBA) both contain, on a flat view of the map (indices for pointers, the children of a keyframe = the keyframes that
// name it as parent, in index order — the reference walks a std::set ordered by pointer value, and no result depends on that order).
// cv::Mat arithmetic: f32 throughout; a product entry is the f32 sum, left to right, of f32 products (cv::gemm's path for inner
// dimension <= 4).  PINNED against OpenCV itself: tests/golden/map_update_cv2.npz holds this loop evaluated with cv2.gemm (Python cv2
// 4.13, the OpenCV generation SURVEY.md §8(c') pins) and tests/test_map_update.py requires bit-for-bit agreement, on the fixture and,
// where cv2 can be imported, on fresh scenes.  The loop itself (which objects are touched, in which order) is restated from the two
// member functions, which cannot be compiled apart from their classes; tests/test_shim_map_update.py runs that restatement on objects.
#include <cstring>
#include <list>
#include <vector>

#include "oracle.h"

namespace {

struct M4 { float m[4][4]; };

M4 mul(const M4& a, const M4& b) {                 // cv::Mat * cv::Mat, 4x4
  M4 c;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float s = a.m[i][0] * b.m[0][j];
      for (int k = 1; k < 4; k++) s = s + a.m[i][k] * b.m[k][j];
      c.m[i][j] = s;
    }
  return c;
}

M4 load(const float* p) { M4 a; std::memcpy(a.m, p, sizeof a.m); return a; }

// KeyFrame::SetPose (cslam/src/KeyFrame.cpp:298-306): Rwc = Rcw.t(); Ow = -Rwc*tcw; Twc = [Rwc | Ow]
M4 inverse_of(const M4& Tcw) {
  M4 Twc;
  for (int i = 0; i < 3; i++) {
    float s = Tcw.m[0][i] * Tcw.m[0][3];
    s = s + Tcw.m[1][i] * Tcw.m[1][3];
    s = s + Tcw.m[2][i] * Tcw.m[2][3];
    for (int j = 0; j < 3; j++) Twc.m[i][j] = Tcw.m[j][i];
    Twc.m[i][3] = -s;
  }
  Twc.m[3][0] = Twc.m[3][1] = Twc.m[3][2] = 0.f; Twc.m[3][3] = 1.f;
  return Twc;
}

// R*x + t with R, t the blocks of T: the product first (f32, left to right), then the sum
void affine(const M4& T, const float x[3], float y[3]) {
  for (int i = 0; i < 3; i++) {
    float s = T.m[i][0] * x[0];
    s = s + T.m[i][1] * x[1];
    s = s + T.m[i][2] * x[2];
    y[i] = s + T.m[i][3];
  }
}

}  // namespace

extern "C" int orc_gba_map_update(int32_t n_kf, const int32_t* kf_parent, const uint8_t* kf_optimized, const float* kf_Tcw, float* kf_TcwGBA,
                                  uint8_t* kf_visited, int32_t n_mp, const uint8_t* mp_state, const int32_t* mp_ref, const float* mp_pos,
                                  const float* mp_pos_gba, float* mp_pos_out, uint8_t* mp_corrected) {
  std::vector<uint8_t> flag(kf_optimized, kf_optimized + n_kf);          // mBAGlobalForKF == nLoopKF
  std::vector<M4> pose(n_kf), before(n_kf);                               // GetPose(), mTcwBefGBA
  for (int k = 0; k < n_kf; k++) { pose[k] = load(kf_Tcw + 16 * (size_t)k); kf_visited[k] = 0; }
  std::list<int> toCheck;                                                 // Map.cpp:1442
  for (int k = 0; k < n_kf; k++)
    if (kf_parent[k] == -1) {
      if (!kf_optimized[k]) return 1;                                     // its mTcwGBA would be an empty Mat
      toCheck.push_back(k);
    }
  while (!toCheck.empty()) {                                              // Map.cpp:1455-1490
    const int kf = toCheck.front();
    const M4 Twc = inverse_of(pose[kf]);                                  // GetPoseInverse(): of the pose not yet updated
    const M4 gba = load(kf_TcwGBA + 16 * (size_t)kf);
    for (int c = 0; c < n_kf; c++) {
      if (kf_parent[c] != kf) continue;
      if (!flag[c]) {
        const M4 Tchildc = mul(pose[c], Twc);
        const M4 g = mul(Tchildc, gba);
        std::memcpy(kf_TcwGBA + 16 * (size_t)c, g.m, sizeof g.m);
        flag[c] = 1;
      }
      toCheck.push_back(c);
    }
    before[kf] = pose[kf];
    pose[kf] = gba;                                                       // SetPose(mTcwGBA, true)
    kf_visited[kf] = 1;
    toCheck.pop_front();
  }
  for (int i = 0; i < n_mp; i++) {                                        // Map.cpp:1497-1563
    const float* x = mp_pos + 3 * (size_t)i;
    float* out = mp_pos_out + 3 * (size_t)i;
    out[0] = x[0]; out[1] = x[1]; out[2] = x[2];
    mp_corrected[i] = 0;
    if (mp_state[i] == 0) continue;                                       // isBad()
    if (mp_state[i] == 1) {                                               // optimised by the BA: take mPosGBA
      std::memcpy(out, mp_pos_gba + 3 * (size_t)i, 3 * sizeof(float));
      mp_corrected[i] = 1;
      continue;
    }
    const int r = mp_ref[i];
    if (r < 0) continue;                                                  // no reference keyframe
    if (!flag[r]) continue;                                               // pRefKF->mBAGlobalForKF != nLoopKF
    if (!kf_visited[r]) continue;                                         // flagged by the BA but outside the tree: mTcwBefGBA was never set
    float xc[3];
    affine(before[r], x, xc);                                             // Rcw*Xw + tcw with mTcwBefGBA
    affine(inverse_of(pose[r]), xc, out);                                 // Rwc*Xc + twc with the corrected pose
    mp_corrected[i] = 1;
  }
  return 0;
}

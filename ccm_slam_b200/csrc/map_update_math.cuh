// map_update_math.cuh — arithmetic of the map update that follows a global BA (ccm_gba_map_update, map_update.cu), shared by the
// kernel, the host tree pass and a host build in tests/ (g++).
//
//   Map::RunGBA        S/Map.cpp:1441-1570        MapMerger::RunGBA   S/MapMerger.cpp:637-753   (the same loop twice)
//   KeyFrame::SetPose  S/KeyFrame.cpp:298-306     (Twc = [Rcw^T | -Rcw^T tcw])
//
// The reference does all of this with f32 cv::Mat expressions.  cv::Mat products of these sizes (inner dimension 3 or 4) take
// cv::gemm's small-matrix path: f32 products summed left to right in f32, one result at a time; `A*B + c` and `-A*b` fuse into the same
// gemm call without changing any rounding.  Checked bit for bit against cv2 4.13 (tests/test_map_update.py, tests/golden/
// map_update_cv2.npz).  Every operation below is single-rounded (no FMA contraction on the device).
#pragma once
#include <stdint.h>

#include <vector>

#if defined(__CUDACC__)
#define CCM_MU_HD __host__ __device__ __forceinline__
#else
#define CCM_MU_HD inline
#endif

namespace ccm {
namespace mu {

CCM_MU_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
CCM_MU_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}

// C = A B, 4x4 row-major f32
CCM_MU_HD void mat4_mul(const float* A, const float* B, float* C) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      float s = fmul(A[4 * r], B[c]);
      s = fadd(s, fmul(A[4 * r + 1], B[4 + c]));
      s = fadd(s, fmul(A[4 * r + 2], B[8 + c]));
      s = fadd(s, fmul(A[4 * r + 3], B[12 + c]));
      C[4 * r + c] = s;
    }
}

// y = R x + t with R = T[0:3,0:3] (or its transpose), t a separate vector: the two-step `R*x + t` of the reference
CCM_MU_HD void rot_apply(const float* T, bool transpose, const float x[3], float y[3]) {
  for (int r = 0; r < 3; r++) {
    const float a0 = transpose ? T[r] : T[4 * r], a1 = transpose ? T[4 + r] : T[4 * r + 1], a2 = transpose ? T[8 + r] : T[4 * r + 2];
    float s = fmul(a0, x[0]);
    s = fadd(s, fmul(a1, x[1]));
    s = fadd(s, fmul(a2, x[2]));
    y[r] = s;
  }
}

// KeyFrame::SetPose: Twc from Tcw
CCM_MU_HD void pose_inverse(const float* Tcw, float* Twc) {
  const float t[3] = {Tcw[3], Tcw[7], Tcw[11]};
  float ow[3];
  rot_apply(Tcw, true, t, ow);
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) Twc[4 * r + c] = Tcw[4 * c + r];
    Twc[4 * r + 3] = -ow[r];
  }
  Twc[12] = 0.f; Twc[13] = 0.f; Twc[14] = 0.f; Twc[15] = 1.f;
}

// mTcwGBA of a keyframe the BA did not hold: (Tcw_child * Twc_parent) * TcwGBA_parent, all poses as they were before the update
CCM_MU_HD void propagate_child(const float* Tcw_child, const float* Tcw_parent, const float* TcwGBA_parent, float* TcwGBA_child) {
  float Twc[16], Tchildc[16];
  pose_inverse(Tcw_parent, Twc);
  mat4_mul(Tcw_child, Twc, Tchildc);
  mat4_mul(Tchildc, TcwGBA_parent, TcwGBA_child);
}

// a point the BA did not hold: into the camera of its reference keyframe as that was before, back out through the corrected pose
CCM_MU_HD void correct_point(const float* Tcw_before, const float* Twc_after, const float X[3], float out[3]) {
  float xc[3];
  rot_apply(Tcw_before, false, X, xc);
  xc[0] = fadd(xc[0], Tcw_before[3]); xc[1] = fadd(xc[1], Tcw_before[7]); xc[2] = fadd(xc[2], Tcw_before[11]);
  rot_apply(Twc_after, false, xc, out);
  out[0] = fadd(out[0], Twc_after[3]); out[1] = fadd(out[1], Twc_after[7]); out[2] = fadd(out[2], Twc_after[11]);
}

// The keyframe pass (host): breadth-first from the origins over the children lists implied by the parent array (parent -1 = a map
// origin, below -1 = not part of the tree); fills TcwGBA of keyframes the BA did not hold and marks what was visited.  Siblings may be
// taken in any order: a keyframe depends on its parent alone.  Returns nullptr or what is wrong with the input.
inline const char* update_keyframes(int n_kf, const int32_t* parent, const uint8_t* optimized, const float* Tcw, float* TcwGBA, uint8_t* visited) {
  std::vector<int> head(n_kf, -1), next(n_kf, -1), tail(n_kf, -1), queue;
  queue.reserve(n_kf);
  for (int i = 0; i < n_kf; i++) {
    visited[i] = 0;
    const int p = parent[i];
    if (p >= n_kf || p == i) return "bad parent index";
    if (p == -1) queue.push_back(i);
    else if (p >= 0) {
      if (head[p] < 0) head[p] = i; else next[tail[p]] = i;
      tail[p] = i;
    }
  }
  for (size_t q = 0; q < queue.size(); q++)
    if (!optimized[queue[q]]) return "a map origin the BA did not hold has no mTcwGBA to start from";
  for (size_t q = 0; q < queue.size(); q++) {
    const int k = queue[q];
    visited[k] = 1;
    for (int c = head[k]; c >= 0; c = next[c]) {
      if (!optimized[c]) propagate_child(Tcw + 16 * (size_t)c, Tcw + 16 * (size_t)k, TcwGBA + 16 * (size_t)k, TcwGBA + 16 * (size_t)c);
      queue.push_back(c);
    }
  }
  return nullptr;
}

}  // namespace mu
}  // namespace ccm

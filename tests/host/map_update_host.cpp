// map_update_host.cpp — TEST CODE: the arithmetic of ccm_gba_map_update (ccm_slam_b200/csrc/map_update_math.cuh) run on the host.
// The keyframe pass is the library's own (mu::update_keyframes, host code there too); the point pass is the body of
// k_map_update_points with the grid-stride loop turned into a plain loop.  Compiled by tests/test_map_update.py with g++.
#include <cstdint>
#include <vector>

#include "../../ccm_slam_b200/csrc/map_update_math.cuh"

using namespace ccm;

extern "C" int mu_host_update(int32_t n_kf, const int32_t* kf_parent, const uint8_t* kf_optimized, const float* kf_Tcw, float* kf_TcwGBA,
                              uint8_t* kf_visited, int32_t n_mp, const uint8_t* mp_state, const int32_t* mp_ref, const float* mp_pos,
                              const float* mp_pos_gba, float* mp_pos_out, uint8_t* mp_corrected) {
  if (mu::update_keyframes(n_kf, kf_parent, kf_optimized, kf_Tcw, kf_TcwGBA, kf_visited)) return 1;
  std::vector<float> twc((size_t)n_kf * 16, 0.f);
  for (int k = 0; k < n_kf; k++)
    if (kf_visited[k]) mu::pose_inverse(kf_TcwGBA + 16 * (size_t)k, twc.data() + 16 * (size_t)k);
  for (int i = 0; i < n_mp; i++) {
    float o[3] = {mp_pos[3 * (size_t)i], mp_pos[3 * (size_t)i + 1], mp_pos[3 * (size_t)i + 2]};
    uint8_t done = 0;
    if (mp_state[i] == 1) {
      for (int j = 0; j < 3; j++) o[j] = mp_pos_gba[3 * (size_t)i + j];
      done = 1;
    } else if (mp_state[i] == 2) {
      const int k = mp_ref[i];
      if (k >= 0 && kf_visited[k]) {
        const float x[3] = {o[0], o[1], o[2]};
        mu::correct_point(kf_Tcw + 16 * (size_t)k, twc.data() + 16 * (size_t)k, x, o);
        done = 1;
      }
    }
    for (int j = 0; j < 3; j++) mp_pos_out[3 * (size_t)i + j] = o[j];
    mp_corrected[i] = done;
  }
  return 0;
}

// window_best_host.cpp — TEST CODE: runs the arithmetic of k_window_best (ccm_slam_b200/csrc/window_best.cuh) on the host.
// The kernel is one warp per query: 32 lanes each walk their share of the window (window_lane_scan), a butterfly keeps the
// minimum key, lane 0 stores index and distance.  Here the 32 lanes run one after another and the same butterfly is applied
// to their results, with the host preparation (CellIndex, fill_window_queries) and post-processing (fuse_from_windows,
// by_sim3_from_windows) the library uses.  Compiled by tests/test_window_best_host.py with g++; nothing here ships.
#include <cstring>
#include <vector>

#include "../../ccm_slam_b200/csrc/window_best.cuh"

using namespace ccm;

static int windows(const ccm_feature_grid* g, const ccm_proj_queries* q, const float* w, int nlevels, std::vector<int>& bi, std::vector<int>& bd) {
  bi.assign(q->m, -1); bd.assign(q->m, 0x7fffffff);
  if (q->m == 0 || g->n == 0) return 0;
  if (!window_key_fits(*g)) return 1;
  const CellIndex cells(*g);
  std::vector<WinQuery> hq;
  fill_window_queries(cells, *q, hq);
  std::vector<uint4> qd((size_t)q->m * 2), kd((size_t)g->n * 2);
  std::memcpy(qd.data(), q->desc, (size_t)q->m * 32);
  std::memcpy(kd.data(), g->desc, (size_t)g->n * 32);
  std::vector<float2> xy(g->n);
  std::memcpy(xy.data(), g->kp_xy, (size_t)g->n * 8);
  const std::vector<int> feat = cells.feat.empty() ? std::vector<int>(1, 0) : cells.feat;
  for (int i = 0; i < q->m; i++) {
    unsigned best[32]; int best_j[32];
    for (int lane = 0; lane < 32; lane++) {
      best[lane] = 0xffffffffu; best_j[lane] = -1;
      if (hq[i].c0 <= hq[i].c1 && hq[i].r0 <= hq[i].r1)
        window_lane_scan(hq[i], lane, qd[(size_t)i * 2], qd[(size_t)i * 2 + 1], cells.ptr.data(), feat.data(), g->grid_rows, xy.data(), g->octave,
                         kd.data(), w, nlevels, best[lane], best_j[lane]);
    }
    for (int off = 16; off > 0; off >>= 1) {   // __shfl_xor_sync butterfly: every lane reads its partner's pre-step value
      unsigned nb[32]; int nj[32];
      for (int lane = 0; lane < 32; lane++) {
        const unsigned ob = best[lane ^ off]; const int oj = best_j[lane ^ off];
        nb[lane] = best[lane]; nj[lane] = best_j[lane];
        if (ob < nb[lane]) { nb[lane] = ob; nj[lane] = oj; }
      }
      std::memcpy(best, nb, sizeof best); std::memcpy(best_j, nj, sizeof best_j);
    }
    bi[i] = best_j[0];
    bd[i] = window_key_distance(best[0], best_j[0]);
  }
  return 0;
}

extern "C" int wb_windows(const ccm_feature_grid* g, const ccm_proj_queries* q, const float* w, int nlevels, int* out_idx, int* out_dist) {
  std::vector<int> bi, bd;
  if (windows(g, q, w, nlevels, bi, bd)) return 1;
  std::memcpy(out_idx, bi.data(), sizeof(int) * bi.size());
  std::memcpy(out_dist, bd.data(), sizeof(int) * bd.size());
  return 0;
}

extern "C" int wb_fuse(const ccm_feature_grid* g, const ccm_proj_queries* q, const float* w, int nlevels, int32_t* best_idx, int32_t* nfound) {
  std::vector<int> bi, bd;
  if (windows(g, q, w, nlevels, bi, bd)) return 1;
  fuse_from_windows(q->m, bi.data(), bd.data(), 50, best_idx, nfound);    // TH_LOW
  return 0;
}

extern "C" int wb_by_sim3(const ccm_feature_grid* g1, const ccm_feature_grid* g2, const ccm_proj_queries* q12, const ccm_proj_queries* q21,
                          int32_t* match12, int32_t* nfound) {
  std::vector<int> i12, d12, i21, d21;
  if (windows(g2, q12, nullptr, 0, i12, d12) || windows(g1, q21, nullptr, 0, i21, d21)) return 1;
  by_sim3_from_windows(q12->m, i12.data(), d12.data(), i21.data(), d21.data(), 100, match12, nfound);   // TH_HIGH
  return 0;
}

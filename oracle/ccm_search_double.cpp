// ccm_search_double.cpp — link-time stand-in for the DEVICE half of the projection-guided matchers (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Lets the reference-side shim (shim/ORBmatcher_proj_shim.cpp) run in a container without a GPU: each ccm_search_* / ccm_match_* entry point of
// include/ccm_b200.h is defined here as "Hamming matrix on the CPU, then the library's own host half ccm_select_*" — the product
// library splits the same way, with the matrix coming from k_hamming on the device.  Linked only into oracle/_ref/libmatch_shim.so
// (in front of libccm_b200.so, -Bsymbolic); the product never contains these definitions and still fails loudly without CUDA.
#include <cstddef>
#include <cstdint>
#include <vector>

#include "ccm_b200.h"

namespace {
std::vector<uint16_t> hamming_raw(const uint8_t* A, int nA, const uint8_t* B, int nB) {
  std::vector<uint16_t> D((size_t)(nA > 0 ? nA : 0) * (size_t)(nB > 0 ? nB : 0));
  for (int i = 0; i < nA; i++)
    for (int j = 0; j < nB; j++) {
      int d = 0;
      for (int b = 0; b < 32; b++) d += __builtin_popcount((unsigned)(A[32 * (size_t)i + b] ^ B[32 * (size_t)j + b]));
      D[(size_t)i * nB + j] = (uint16_t)d;
    }
  return D;
}
std::vector<uint16_t> hamming(const ccm_proj_queries* q, const ccm_feature_grid* g) {
  std::vector<uint16_t> D((size_t)(q ? q->m : 0) * (size_t)(g ? g->n : 0));
  if (D.empty()) return D;
  for (int i = 0; i < q->m; i++)
    for (int j = 0; j < g->n; j++) {
      int d = 0;
      for (int b = 0; b < 32; b++) d += __builtin_popcount((unsigned)(q->desc[32 * (size_t)i + b] ^ g->desc[32 * (size_t)j + b]));
      D[(size_t)i * g->n + j] = (uint16_t)d;
    }
  return D;
}
}  // namespace

extern "C" {
int ccm_search_by_projection_track(const ccm_feature_grid* g, const ccm_proj_queries* q, const uint8_t* query_has_obs, const uint8_t* feat_blocked,
                                   float nnratio, int32_t* match_of_feat, int32_t* nmatches) {
  return ccm_select_by_projection_track(g, q, hamming(q, g).data(), query_has_obs, feat_blocked, nnratio, match_of_feat, nmatches);
}
int ccm_search_by_projection_frame(const ccm_feature_grid* g, const ccm_proj_queries* q, const uint8_t* query_has_obs, const uint8_t* feat_blocked,
                                   int32_t reloc, int32_t orb_dist, int32_t check_orientation, int32_t* match_of_feat, int32_t* nmatches) {
  return ccm_select_by_projection_frame(g, q, hamming(q, g).data(), query_has_obs, feat_blocked, reloc, orb_dist, check_orientation, match_of_feat, nmatches);
}
int ccm_search_by_projection_sim3(const ccm_feature_grid* g, const ccm_proj_queries* q, const uint8_t* feat_matched, const int32_t* existing_idx,
                                  int32_t* best_idx, int32_t* match_of_feat, int32_t* nmatches) {
  return ccm_select_by_projection_sim3(g, q, hamming(q, g).data(), feat_matched, existing_idx, best_idx, match_of_feat, nmatches);
}
int ccm_fuse_search(const ccm_feature_grid* g, const ccm_proj_queries* q, const float* inv_level_sigma2, int32_t nlevels, int32_t* best_idx,
                    int32_t* nfound) {
  return ccm_fuse_select(g, q, hamming(q, g).data(), inv_level_sigma2, nlevels, best_idx, nfound);
}
int ccm_search_by_sim3(const ccm_feature_grid* g1, const ccm_feature_grid* g2, const ccm_proj_queries* q12, const ccm_proj_queries* q21,
                       int32_t* match12, int32_t* nfound) {
  return ccm_select_by_sim3(g1, g2, q12, q21, hamming(q12, g2).data(), hamming(q21, g1).data(), match12, nfound);
}
int ccm_search_for_initialization(const ccm_feature_grid* g2, const ccm_proj_queries* q, float nnratio, int32_t check_orientation, int32_t* match12,
                                  int32_t* nmatches) {
  return ccm_select_for_initialization(g2, q, hamming(q, g2).data(), nnratio, check_orientation, match12, nmatches);
}

/* SearchByBoW x2, SearchForTriangulation, DescriptorDistance (shim/ORBmatcher_shim.cpp) */
int ccm_hamming_matrix(const uint8_t* A, int32_t nA, const uint8_t* B, int32_t nB, uint16_t* D) {
  std::vector<uint16_t> h = hamming_raw(A, nA, B, nB);
  for (size_t i = 0; i < h.size(); i++) D[i] = h[i];
  return CCM_OK;
}
int ccm_match_bow_kf_frame(const uint8_t* desc_kf, int32_t n_kf, const uint8_t* kf_has_mp, const float* angle_kf, const ccm_feature_vector* fv_kf,
                           const uint8_t* desc_f, int32_t n_f, const float* angle_f, const ccm_feature_vector* fv_f, float nnratio,
                           int32_t check_orientation, int32_t* match_kf_of_f, int32_t* nmatches) {
  return ccm_select_bow_kf_frame(hamming_raw(desc_kf, n_kf, desc_f, n_f).data(), n_kf, kf_has_mp, angle_kf, fv_kf, n_f, angle_f, fv_f, nnratio,
                                 check_orientation, match_kf_of_f, nmatches);
}
int ccm_match_bow_kf_kf(const uint8_t* desc1, int32_t n1, const uint8_t* has_mp1, const float* angle1, const ccm_feature_vector* fv1,
                        const uint8_t* desc2, int32_t n2, const uint8_t* has_mp2, const float* angle2, const ccm_feature_vector* fv2, float nnratio,
                        int32_t check_orientation, int32_t* match12, int32_t* nmatches) {
  return ccm_select_bow_kf_kf(hamming_raw(desc1, n1, desc2, n2).data(), n1, has_mp1, angle1, fv1, n2, has_mp2, angle2, fv2, nnratio,
                              check_orientation, match12, nmatches);
}
int ccm_match_triangulation(const ccm_tri_view* v1, const ccm_tri_view* v2, const float F12[9], float ex, float ey, const float* level_sigma2,
                            const float* scale_factors, int32_t nlevels, int32_t check_orientation, int32_t* pairs, int32_t* npairs) {
  return ccm_select_triangulation(hamming_raw(v1->desc, v1->n, v2->desc, v2->n).data(), v1, v2, F12, ex, ey, level_sigma2, scale_factors, nlevels,
                                  check_orientation, pairs, npairs);
}
}

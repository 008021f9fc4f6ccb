// ccm_map_update_double.cpp — link-time stand-in for ccm_gba_map_update (TEST INFRASTRUCTURE, NOT PRODUCT): the CPU oracle behind the
// device entry point shim/MapUpdate_shim.cpp calls, so that the shim's flattening and write-back can run where there is no GPU.
// Linked only into oracle/_ref/libmap_update_shim.so.
#include "ccm_b200.h"
#include "oracle.h"

static thread_local const char* g_err = "";
extern "C" const char* ccm_last_error(void) { return g_err; }
extern "C" int ccm_gba_map_update(int32_t n_kf, const int32_t* kf_parent, const uint8_t* kf_optimized, const float* kf_Tcw, float* kf_TcwGBA,
                                  uint8_t* kf_visited, int32_t n_mp, const uint8_t* mp_state, const int32_t* mp_ref, const float* mp_pos,
                                  const float* mp_pos_gba, float* mp_pos_out, uint8_t* mp_corrected) {
  if (orc_gba_map_update(n_kf, kf_parent, kf_optimized, kf_Tcw, kf_TcwGBA, kf_visited, n_mp, mp_state, mp_ref, mp_pos, mp_pos_gba, mp_pos_out,
                         mp_corrected) != 0) {
    g_err = "ccm_gba_map_update: a map origin the BA did not hold has no mTcwGBA to start from";
    return CCM_ERR_INVALID;
  }
  return CCM_OK;
}

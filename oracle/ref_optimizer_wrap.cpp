// ref_optimizer_wrap.cpp — drives shim/Optimizer_shim.cpp through the reference's own cslam::Optimizer interface on a stand-in map
// (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Builds cslam::Map / KeyFrame / MapPoint / Frame stand-ins (oracle/ref_stub_opt/cslam/Frame.h) from flat arrays, calls
// Optimizer::MapFusionGBA, GlobalBundleAdjustemntClient, LocalBundleAdjustmentClient, PoseOptimizationClient, OptimizeSim3,
// OptimizeEssentialGraphMapFusion / LoopClosure exactly as the reference's callers do, and copies back what the shim wrote into the
// objects.  Compiled with the shim, the reference's own Optimizer.h / Converter.h / Converter.cc and the device double
// (oracle/ccm_device_double.cpp) into oracle/_ref/liboptimizer_shim.so.
// Keyframes live in one contiguous array so that shared_ptr ordering (std::map<kfptr, ...>, std::set<kfptr>) is index order.
#include <cslam/Optimizer.h>

#include <cstdint>
#include <cstring>

using namespace cslam;

float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
std::mutex MapPoint::mGlobalMutex;

extern "C" {

typedef struct {
  int32_t K;
  const int64_t* kf_uid; const int64_t* kf_id;      /* mUniqueId; mId = (kf_id[2k], kf_id[2k+1]) */
  const uint8_t* kf_bad;
  const float* kf_Tcw;                               /* K*16 */
  const float* kf_intr;                              /* K*4 fx fy cx cy */
  const int32_t* kp_ptr; const float* kp_uv; const int32_t* kp_octave;   /* keypoints of keyframe k: kp_ptr[k]..kp_ptr[k+1] */
  const float* inv_level_sigma2; int32_t nlevels;
  const int32_t* kf_parent;                          /* index or -1 */
  const int32_t* loop_ptr; const int32_t* loop_kf;   /* GetLoopEdges */
  const int32_t* cov_ptr; const int32_t* cov_kf; const int32_t* cov_w;   /* ordered covisibility with weights */
  int32_t P;
  const int64_t* mp_uid; const int64_t* mp_id;
  const uint8_t* mp_bad;
  const float* mp_pos;                               /* P*3 */
  const int32_t* mp_ref;                             /* reference keyframe index */
  const int32_t* obs_ptr; const int32_t* obs_kf; const int32_t* obs_idx;  /* observations of point j: (keyframe index, keypoint index) */
  int32_t origin;                                    /* mvpKeyFrameOrigins[0] */
  int64_t map_id;
} optw_scene;

typedef struct {
  float* kf_Tcw; float* kf_TcwGBA; int64_t* kf_gba_tag; int32_t* kf_set_pose;   /* K*16, K*16, K*2, K */
  float* mp_pos; float* mp_posGBA; int64_t* mp_gba_tag; int32_t* mp_set_pos; int32_t* mp_update_normal;  /* P*3, P*3, P*2, P, P */
  int32_t* mp_n_obs; int32_t* kf_n_erased;                                       /* P, K: after the call */
} optw_out;

}  // extern "C"

namespace {

struct Null { template <class T> void operator()(T*) const {} };

struct Scene {
  std::vector<KeyFrame> kf_store;
  std::vector<MapPoint> mp_store;
  std::vector<Optimizer::kfptr> kf;
  std::vector<Optimizer::mpptr> mp;
  boost::shared_ptr<Map> map;
  Map map_store;
  explicit Scene(const optw_scene* s) : kf_store(s->K), mp_store(s->P) {
    for (int k = 0; k < s->K; k++) kf.push_back(Optimizer::kfptr(&kf_store[k], Null()));
    for (int j = 0; j < s->P; j++) mp.push_back(Optimizer::mpptr(&mp_store[j], Null()));
    for (int k = 0; k < s->K; k++) {
      KeyFrame& F = kf_store[k];
      F.mUniqueId = (size_t)s->kf_uid[k]; F.mId = idpair((size_t)s->kf_id[2 * k], (size_t)s->kf_id[2 * k + 1]); F.mbBad = s->kf_bad[k] != 0;
      F.Tcw.create(4, 4, CV_32F); std::memcpy(F.Tcw.ptr<float>(0), s->kf_Tcw + 16 * (size_t)k, 16 * sizeof(float));
      F.fx = s->kf_intr[4 * k]; F.fy = s->kf_intr[4 * k + 1]; F.cx = s->kf_intr[4 * k + 2]; F.cy = s->kf_intr[4 * k + 3];
      F.mK = cv::Mat::eye(3, 3, CV_32F); F.mK.at<float>(0, 0) = F.fx; F.mK.at<float>(1, 1) = F.fy; F.mK.at<float>(0, 2) = F.cx; F.mK.at<float>(1, 2) = F.cy;
      for (int q = s->kp_ptr[k]; q < s->kp_ptr[k + 1]; q++) {
        cv::KeyPoint kp; kp.pt.x = s->kp_uv[2 * q]; kp.pt.y = s->kp_uv[2 * q + 1]; kp.octave = s->kp_octave[q];
        F.mvKeysUn.push_back(kp);
      }
      F.mvpMapPoints.resize(F.mvKeysUn.size());
      F.mvInvLevelSigma2.assign(s->inv_level_sigma2, s->inv_level_sigma2 + s->nlevels);
      if (s->kf_parent && s->kf_parent[k] >= 0) { F.mpParent = kf[s->kf_parent[k]]; kf_store[s->kf_parent[k]].mspChildrens.insert(kf[k]); }
      if (s->loop_ptr) for (int q = s->loop_ptr[k]; q < s->loop_ptr[k + 1]; q++) F.mspLoopEdges.insert(kf[s->loop_kf[q]]);
      if (s->cov_ptr) for (int q = s->cov_ptr[k]; q < s->cov_ptr[k + 1]; q++) { F.mvpOrderedConnectedKeyFrames.push_back(kf[s->cov_kf[q]]); F.mvOrderedWeights.push_back(s->cov_w[q]); }
    }
    for (int j = 0; j < s->P; j++) {
      MapPoint& M = mp_store[j];
      M.mUniqueId = (size_t)s->mp_uid[j]; M.mId = idpair((size_t)s->mp_id[2 * j], (size_t)s->mp_id[2 * j + 1]); M.mbBad = s->mp_bad[j] != 0;
      M.mWorldPos.create(3, 1, CV_32F); std::memcpy(M.mWorldPos.ptr<float>(0), s->mp_pos + 3 * (size_t)j, 3 * sizeof(float));
      if (s->mp_ref && s->mp_ref[j] >= 0) M.mpRefKF = kf[s->mp_ref[j]];
      for (int q = s->obs_ptr[j]; q < s->obs_ptr[j + 1]; q++) {
        M.mObservations[kf[s->obs_kf[q]]] = (size_t)s->obs_idx[q];
        kf_store[s->obs_kf[q]].mvpMapPoints[s->obs_idx[q]] = mp[j];
      }
    }
    map_store.kfs = kf; map_store.mps = mp; map_store.mMapId = (size_t)s->map_id;
    if (s->origin >= 0) map_store.mvpKeyFrameOrigins.push_back(kf[s->origin]);
    map = boost::shared_ptr<Map>(&map_store, Null());
  }
  void read(const optw_scene* s, optw_out* o) {
    for (int k = 0; k < s->K; k++) {
      KeyFrame& F = kf_store[k];
      std::memcpy(o->kf_Tcw + 16 * (size_t)k, F.Tcw.ptr<float>(0), 16 * sizeof(float));
      if (!F.mTcwGBA.empty()) std::memcpy(o->kf_TcwGBA + 16 * (size_t)k, F.mTcwGBA.ptr<float>(0), 16 * sizeof(float));
      o->kf_gba_tag[2 * k] = (int64_t)F.mBAGlobalForKF.first; o->kf_gba_tag[2 * k + 1] = (int64_t)F.mBAGlobalForKF.second;
      o->kf_set_pose[k] = F.n_set_pose; o->kf_n_erased[k] = F.n_erased;
    }
    for (int j = 0; j < s->P; j++) {
      MapPoint& M = mp_store[j];
      std::memcpy(o->mp_pos + 3 * (size_t)j, M.mWorldPos.ptr<float>(0), 3 * sizeof(float));
      if (!M.mPosGBA.empty()) std::memcpy(o->mp_posGBA + 3 * (size_t)j, M.mPosGBA.ptr<float>(0), 3 * sizeof(float));
      o->mp_gba_tag[2 * j] = (int64_t)M.mBAGlobalForKF.first; o->mp_gba_tag[2 * j + 1] = (int64_t)M.mBAGlobalForKF.second;
      o->mp_set_pos[j] = M.n_set_pos; o->mp_update_normal[j] = M.n_update_normal; o->mp_n_obs[j] = (int32_t)M.mObservations.size();
    }
  }
};

}  // namespace

extern "C" {

/* which: 0 = MapFusionGBA, 1 = GlobalBundleAdjustemntClient (client id = map id); loop = nLoopKF */
int optw_gba(const optw_scene* s, int which, int iterations, int robust, int64_t loop_first, int64_t loop_second, optw_out* o) {
  try {
    Scene sc(s);
    const idpair nLoopKF((size_t)loop_first, (size_t)loop_second);
    if (which == 0) Optimizer::MapFusionGBA(sc.map, (size_t)s->map_id, iterations, NULL, nLoopKF, robust != 0);
    else Optimizer::GlobalBundleAdjustemntClient(sc.map, (size_t)s->map_id, iterations, NULL, nLoopKF, robust != 0);
    sc.read(s, o);
    return 0;
  } catch (...) { return -1; }
}

/* as optw_gba, but keyframe flip_kf turns bad after the solve and before the write-back (GBA runs in its own thread while culling goes on).
   Needs the device double's hook (liboptimizer_shim.so); returns -2 in the builds without it. */
}
#include <dlfcn.h>
namespace { struct Flip { Scene* sc; int kf; }; void do_flip(void* a) { Flip* f = static_cast<Flip*>(a); f->sc->kf_store[f->kf].mbBad = true; } }
extern "C" {
int optw_gba_flip(const optw_scene* s, int which, int iterations, int robust, int64_t loop_first, int64_t loop_second, int flip_kf, optw_out* o) {
  typedef void (*hook_t)(void*);
  Dl_info self;   /* the library is loaded RTLD_LOCAL by ctypes: look the hook up in this very object */
  if (!dladdr(reinterpret_cast<void*>(&do_flip), &self)) return -2;
  void* me = dlopen(self.dli_fname, RTLD_NOLOAD | RTLD_NOW);
  if (!me) return -2;
  hook_t* hook = reinterpret_cast<hook_t*>(dlsym(me, "ccm_double_after_solve"));
  void** arg = reinterpret_cast<void**>(dlsym(me, "ccm_double_after_solve_arg"));
  dlclose(me);
  if (!hook || !arg) return -2;
  try {
    Scene sc(s);
    Flip f{&sc, flip_kf};
    *hook = do_flip; *arg = &f;
    const idpair nLoopKF((size_t)loop_first, (size_t)loop_second);
    try {
      if (which == 0) Optimizer::MapFusionGBA(sc.map, (size_t)s->map_id, iterations, NULL, nLoopKF, robust != 0);
      else Optimizer::GlobalBundleAdjustemntClient(sc.map, (size_t)s->map_id, iterations, NULL, nLoopKF, robust != 0);
    } catch (...) { *hook = nullptr; *arg = nullptr; throw; }
    *hook = nullptr; *arg = nullptr;
    sc.read(s, o);
    return 0;
  } catch (...) { return -1; }
}

/* MapFusionGBA with a persistent mirror registered for the map (shim builds only): the mirror is fed from the scene the way a server
   would feed it (keyframes, points, observations, in map order), then the optimiser takes its problem from it. */
}
#if defined(CCM_SHIM_BUILD)
#include "ccm_b200.h"
namespace cslam { void ccm_b200_register_mirror(const Map* map, ccm_map_mirror* mirror); }
extern "C" int optw_gba_mirror(const optw_scene* s, int iterations, int robust, int64_t loop_first, int64_t loop_second, optw_out* o) {
  ccm_map_mirror* mir = nullptr;
  try {
    Scene sc(s);
    if (ccm_mirror_create(&mir) != 0) return -3;
    for (int k = 0; k < s->K; k++) {
      KeyFrame& F = sc.kf_store[k];
      const float intr[4] = {F.fx, F.fy, F.cx, F.cy};
      if (ccm_mirror_set_keyframe(mir, (uint64_t)F.mUniqueId, F.Tcw.ptr<float>(0), intr, F.mbBad ? 1 : 0) != 0) throw 1;
    }
    for (int j = 0; j < s->P; j++) {
      MapPoint& M = sc.mp_store[j];
      if (ccm_mirror_set_point(mir, (uint64_t)M.mUniqueId, M.mWorldPos.ptr<float>(0), M.mbBad ? 1 : 0) != 0) throw 1;
      for (auto& ob : M.mObservations) {                       /* std::map<kfptr, size_t>: the order the reference's loop visits them in */
        const KeyFrame& F = *ob.first;
        const cv::KeyPoint& kp = F.mvKeysUn[ob.second];
        if (ccm_mirror_set_observation(mir, (uint64_t)F.mUniqueId, (uint64_t)M.mUniqueId, kp.pt.x, kp.pt.y, F.mvInvLevelSigma2[kp.octave]) != 0) throw 1;
      }
    }
    cslam::ccm_b200_register_mirror(sc.map.get(), mir);
    const idpair nLoopKF((size_t)loop_first, (size_t)loop_second);
    try { Optimizer::MapFusionGBA(sc.map, (size_t)s->map_id, iterations, NULL, nLoopKF, robust != 0); }
    catch (...) { cslam::ccm_b200_register_mirror(sc.map.get(), nullptr); ccm_mirror_destroy(mir); return -1; }
    cslam::ccm_b200_register_mirror(sc.map.get(), nullptr);
    ccm_mirror_destroy(mir);
    sc.read(s, o);
    return 0;
  } catch (...) { if (mir) ccm_mirror_destroy(mir); return -1; }
}
/* two MapFusionGBA calls in a row on one map (direct write-back, so the second starts from the first's result); with a mirror the
   values the first call wrote are reported to it in between (what SetPose / SetWorldPos hooks do on a server): value changes only, so
   the second call must reuse the solver handle.  *creates = handles created in all (device double's counter), -1 if unknown. */
extern "C" int optw_gba_twice(const optw_scene* s, int use_mirror, int iterations, optw_out* o, int32_t* creates) {
  ccm_map_mirror* mir = nullptr;
  Dl_info self;
  int* counter = nullptr;
  if (dladdr(reinterpret_cast<void*>(&do_flip), &self)) {
    void* me = dlopen(self.dli_fname, RTLD_NOLOAD | RTLD_NOW);
    if (me) { counter = reinterpret_cast<int*>(dlsym(me, "ccm_double_ba_creates")); dlclose(me); }
  }
  const int c0 = counter ? *counter : 0;
  try {
    Scene sc(s);
    auto feed_values = [&]() {
      for (int k = 0; k < s->K; k++) { KeyFrame& F = sc.kf_store[k]; if (ccm_mirror_set_keyframe(mir, (uint64_t)F.mUniqueId, F.Tcw.ptr<float>(0), nullptr, F.mbBad ? 1 : 0) != 0) throw 1; }
      for (int j = 0; j < s->P; j++) { MapPoint& M = sc.mp_store[j]; if (ccm_mirror_set_point(mir, (uint64_t)M.mUniqueId, M.mWorldPos.ptr<float>(0), M.mbBad ? 1 : 0) != 0) throw 1; }
    };
    if (use_mirror) {
      if (ccm_mirror_create(&mir) != 0) return -3;
      for (int k = 0; k < s->K; k++) {
        KeyFrame& F = sc.kf_store[k];
        const float intr[4] = {F.fx, F.fy, F.cx, F.cy};
        if (ccm_mirror_set_keyframe(mir, (uint64_t)F.mUniqueId, F.Tcw.ptr<float>(0), intr, F.mbBad ? 1 : 0) != 0) throw 1;
      }
      for (int j = 0; j < s->P; j++) {
        MapPoint& M = sc.mp_store[j];
        if (ccm_mirror_set_point(mir, (uint64_t)M.mUniqueId, M.mWorldPos.ptr<float>(0), M.mbBad ? 1 : 0) != 0) throw 1;
        for (auto& ob : M.mObservations) {
          const KeyFrame& F = *ob.first; const cv::KeyPoint& kp = F.mvKeysUn[ob.second];
          if (ccm_mirror_set_observation(mir, (uint64_t)F.mUniqueId, (uint64_t)M.mUniqueId, kp.pt.x, kp.pt.y, F.mvInvLevelSigma2[kp.octave]) != 0) throw 1;
        }
      }
      cslam::ccm_b200_register_mirror(sc.map.get(), mir);
    }
    const idpair direct((size_t)0, (size_t)s->map_id);
    for (int round = 0; round < 2; round++) {
      if (use_mirror && round == 1) feed_values();
      Optimizer::MapFusionGBA(sc.map, (size_t)s->map_id, iterations, NULL, direct, true);
    }
    if (use_mirror) { cslam::ccm_b200_register_mirror(sc.map.get(), nullptr); ccm_mirror_destroy(mir); mir = nullptr; }
    sc.read(s, o);
    *creates = counter ? *counter - c0 : -1;
    return 0;
  } catch (...) { if (mir) ccm_mirror_destroy(mir); return -1; }
}
#endif
extern "C" {

int optw_local_ba(const optw_scene* s, int kf_index, int server, optw_out* o) {
  try {
    Scene sc(s);
    Optimizer::LocalBundleAdjustmentClient(sc.kf[kf_index], NULL, sc.map, (size_t)s->map_id, server ? eSystemState::SERVER : eSystemState::CLIENT);
    sc.read(s, o);
    return 0;
  } catch (...) { return -1; }
}

/* loop connections: for the keyframe cur, the set conn[0..n_conn); loop closure variant: corrected / non-corrected Sim3 of the listed keyframes */
int optw_essential_graph(const optw_scene* s, int loop_kf, int cur_kf, const int32_t* conn_ptr, const int32_t* conn_kf, int fix_scale,
                         int loop_closure, int n_corr, const int32_t* corr_kf, const double* corrected /*8 each*/, const double* noncorrected,
                         const int32_t* mp_corr_ref /*P: -1 or keyframe uid the point was corrected through*/, optw_out* o) {
  try {
    Scene sc(s);
    map<Optimizer::kfptr, set<Optimizer::kfptr> > LoopConnections;
    for (int k = 0; k < s->K; k++)
      for (int q = conn_ptr[k]; q < conn_ptr[k + 1]; q++) LoopConnections[sc.kf[k]].insert(sc.kf[conn_kf[q]]);
    for (int j = 0; j < s->P; j++)
      if (mp_corr_ref && mp_corr_ref[j] >= 0) {
        if (loop_closure) { sc.mp_store[j].mCorrectedByKF_LC = sc.kf_store[cur_kf].mId; sc.mp_store[j].mCorrectedReference_LC = (size_t)mp_corr_ref[j]; }
        else { sc.mp_store[j].mCorrectedByKF_MM = sc.kf_store[cur_kf].mId; sc.mp_store[j].mCorrectedReference_MM = (size_t)mp_corr_ref[j]; }
      }
    const bool fs = fix_scale != 0;
    if (loop_closure) {
      Optimizer::KeyFrameAndPose Corrected, NonCorrected;
      for (int i = 0; i < n_corr; i++) {
        const double* c = corrected + 8 * (size_t)i; const double* n = noncorrected + 8 * (size_t)i;
        Corrected[sc.kf[corr_kf[i]]] = g2o::Sim3(Eigen::Quaterniond(c[3], c[0], c[1], c[2]), Eigen::Vector3d(c[4], c[5], c[6]), c[7]);
        NonCorrected[sc.kf[corr_kf[i]]] = g2o::Sim3(Eigen::Quaterniond(n[3], n[0], n[1], n[2]), Eigen::Vector3d(n[4], n[5], n[6]), n[7]);
      }
      Optimizer::OptimizeEssentialGraphLoopClosure(sc.map, sc.kf[loop_kf], sc.kf[cur_kf], NonCorrected, Corrected, LoopConnections, fs);
    } else {
      Optimizer::OptimizeEssentialGraphMapFusion(sc.map, sc.kf[loop_kf], sc.kf[cur_kf], LoopConnections, fs);
    }
    sc.read(s, o);
    return 0;
  } catch (...) { return -1; }
}

/* PoseOptimizationClient on a frame: keypoint i carries map point mp_of_kp[i] (index into the scene's points) or -1 */
int optw_pose_optimization(const optw_scene* s, int n_kp, const float* kp_uv, const int32_t* kp_octave, const int32_t* mp_of_kp, const float* Tcw,
                           const float* intr, float* Tcw_out, uint8_t* outlier_out, int32_t* n_set_pose) {
  try {
    Scene sc(s);
    Frame F;
    F.N = n_kp;
    Frame::fx = intr[0]; Frame::fy = intr[1]; Frame::cx = intr[2]; Frame::cy = intr[3];
    for (int i = 0; i < n_kp; i++) {
      cv::KeyPoint kp; kp.pt.x = kp_uv[2 * i]; kp.pt.y = kp_uv[2 * i + 1]; kp.octave = kp_octave[i];
      F.mvKeysUn.push_back(kp);
      F.mvpMapPoints.push_back(mp_of_kp[i] >= 0 ? sc.mp[mp_of_kp[i]] : Optimizer::mpptr());
    }
    F.mvbOutlier.assign(n_kp, true);     // the call must reset the flag of every keypoint that carries a point
    F.mvInvLevelSigma2.assign(s->inv_level_sigma2, s->inv_level_sigma2 + s->nlevels);
    F.mTcw.create(4, 4, CV_32F); std::memcpy(F.mTcw.ptr<float>(0), Tcw, 16 * sizeof(float));
    const int r = Optimizer::PoseOptimizationClient(F);
    std::memcpy(Tcw_out, F.mTcw.ptr<float>(0), 16 * sizeof(float));
    for (int i = 0; i < n_kp; i++) outlier_out[i] = F.mvbOutlier[i] ? 1 : 0;
    *n_set_pose = F.n_set_pose;
    return r;
  } catch (...) { return -1000; }
}

/* OptimizeSim3 between keyframes k1 and k2: match1[i] = point index matched to keypoint i of k1, or -1 */
int optw_optimize_sim3(const optw_scene* s, int k1, int k2, const int32_t* match1, const double* S12_in, float th2, int fix_scale, double* S12_out,
                       int32_t* match1_out) {
  try {
    Scene sc(s);
    const int n = (int)sc.kf_store[k1].mvKeysUn.size();
    std::vector<Optimizer::mpptr> vpMatches1(n);
    for (int i = 0; i < n; i++) if (match1[i] >= 0) vpMatches1[i] = sc.mp[match1[i]];
    g2o::Sim3 S(Eigen::Quaterniond(S12_in[3], S12_in[0], S12_in[1], S12_in[2]), Eigen::Vector3d(S12_in[4], S12_in[5], S12_in[6]), S12_in[7]);
    const int r = Optimizer::OptimizeSim3(sc.kf[k1], sc.kf[k2], vpMatches1, S, th2, fix_scale != 0);
    S12_out[0] = S.rotation().x(); S12_out[1] = S.rotation().y(); S12_out[2] = S.rotation().z(); S12_out[3] = S.rotation().w();
    for (int i = 0; i < 3; i++) S12_out[4 + i] = S.translation()[i];
    S12_out[7] = S.scale();
    for (int i = 0; i < n; i++) match1_out[i] = vpMatches1[i] ? (int32_t)(vpMatches1[i].get() - &sc.mp_store[0]) : -1;
    return r;
  } catch (...) { return -1000; }
}

}  // extern "C"

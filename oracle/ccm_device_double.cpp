// ccm_device_double.cpp — link-time stand-in for the DEVICE entry points shim/Optimizer_shim.cpp calls (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Lets the reference-side optimiser shim run in a container without a GPU: ccm_ba_solve / ccm_ba_create..optimize / ccm_pgo_solve /
// ccm_pose_optimize / ccm_sim3_optimize are defined here on top of the CPU oracle (liboracle.so), so that a test can push a stand-in
// map through Optimizer::MapFusionGBA etc. and check what the shim selects, flattens and writes back.  What the device computes is
// checked elsewhere (tests/test_gpu_*.py against the same oracle).  Linked only into oracle/_ref/liboptimizer_shim.so, in front of
// libccm_b200.so (-Bsymbolic), which still provides the host-side helpers (ccm_pose_from_Tcw_f32, ccm_pose_to_Tcw_f32).
#include <cstring>
#include <vector>

#include "ccm_b200.h"
#include "oracle.h"

namespace {
orc_ba_problem to_orc(const ccm_ba_problem* p) {
  orc_ba_problem o;
  o.K = p->K; o.P = p->P; o.E = p->E; o.poses = p->poses; o.intr = p->intr; o.fixed = p->fixed; o.points = p->points;
  o.obs_kf = p->obs_kf; o.obs_mp = p->obs_mp; o.obs_uv = p->obs_uv; o.obs_w = p->obs_w; o.edge_flags = p->edge_flags;
  return o;
}
int run_ba(const ccm_ba_problem* p, const ccm_ba_options* o, ccm_ba_result* r, std::vector<double>* poses_out, std::vector<double>* points_out) {
  orc_ba_problem op = to_orc(p);
  orc_ba_options oo = {};
  oo.iterations = o->iterations; oo.robust = o->robust; oo.huber_delta = o->huber_delta; oo.lambda_init = o->lambda_init;
  oo.max_trials = o->max_trials > 0 ? o->max_trials : 10; oo.stop = o->stop;
  std::vector<double> poses((size_t)p->K * 7), points((size_t)p->P * 3);
  orc_ba_result rr = {};
  rr.poses = poses.data(); rr.points = points.data(); rr.chi2 = r->chi2; rr.depth_pos = r->depth_pos;
  if (orc_ba_solve(&op, &oo, &rr) != 0) return CCM_ERR_INVALID;
  if (r->poses) std::memcpy(r->poses, poses.data(), sizeof(double) * poses.size());
  if (r->points) std::memcpy(r->points, points.data(), sizeof(double) * points.size());
  r->iters_done = rr.iters_done; r->trials_total = rr.trials_total; r->chi2_initial = rr.chi2_initial; r->chi2_final = rr.chi2_final;
  r->lambda_final = rr.lambda_final; r->trace_len = 0;
  if (poses_out) poses_out->swap(poses);
  if (points_out) points_out->swap(points);
  return CCM_OK;
}
}  // namespace

struct ccm_ba_handle {
  std::vector<double> poses, intr, points;
  std::vector<uint8_t> fixed, flags;
  std::vector<int32_t> obs_kf, obs_mp;
  std::vector<float> uv, w;
  bool has_flags = false;
  ccm_ba_problem view() const {
    ccm_ba_problem p;
    p.K = (int32_t)fixed.size(); p.P = (int32_t)points.size() / 3; p.E = (int32_t)obs_kf.size();
    p.poses = poses.data(); p.intr = intr.data(); p.fixed = fixed.data(); p.points = points.data(); p.obs_kf = obs_kf.data();
    p.obs_mp = obs_mp.data(); p.obs_uv = uv.data(); p.obs_w = w.data(); p.edge_flags = has_flags ? flags.data() : nullptr;
    return p;
  }
};

extern "C" {
const char* ccm_last_error(void) { return "(device double)"; }
int ccm_double_ba_creates = 0;   /* test counter: how many handles ccm_ba_create has made */
/* test hook: runs between the solve and the shim's write-back (a keyframe turning bad while the GBA thread was solving) */
void (*ccm_double_after_solve)(void*) = nullptr;
void* ccm_double_after_solve_arg = nullptr;
int ccm_ba_solve(const ccm_ba_problem* p, const ccm_ba_options* o, ccm_ba_result* r) {
  const int rc = run_ba(p, o, r, nullptr, nullptr);
  if (ccm_double_after_solve) ccm_double_after_solve(ccm_double_after_solve_arg);
  return rc;
}
int ccm_ba_create(const ccm_ba_problem* p, ccm_ba_handle** out) {
  ccm_double_ba_creates++;
  ccm_ba_handle* h = new ccm_ba_handle;
  h->poses.assign(p->poses, p->poses + 7 * (size_t)p->K); h->intr.assign(p->intr, p->intr + 4 * (size_t)p->K);
  h->fixed.assign(p->fixed, p->fixed + p->K); h->points.assign(p->points, p->points + 3 * (size_t)p->P);
  h->obs_kf.assign(p->obs_kf, p->obs_kf + p->E); h->obs_mp.assign(p->obs_mp, p->obs_mp + p->E);
  h->uv.assign(p->obs_uv, p->obs_uv + 2 * (size_t)p->E); h->w.assign(p->obs_w, p->obs_w + p->E);
  if (p->edge_flags) { h->flags.assign(p->edge_flags, p->edge_flags + p->E); h->has_flags = true; }
  *out = h;
  return CCM_OK;
}
int ccm_ba_set_estimate(ccm_ba_handle* h, const double* poses, const double* points) {
  if (poses) h->poses.assign(poses, poses + h->poses.size());
  if (points) h->points.assign(points, points + h->points.size());
  return CCM_OK;
}
int ccm_ba_set_edge_flags(ccm_ba_handle* h, const uint8_t* f) { h->flags.assign(f, f + h->obs_kf.size()); h->has_flags = true; return CCM_OK; }
int ccm_ba_optimize(ccm_ba_handle* h, const ccm_ba_options* o, ccm_ba_result* r) {   // continues from the handle's current estimate
  ccm_ba_problem p = h->view();
  std::vector<double> poses, points;
  const int rc = run_ba(&p, o, r, &poses, &points);
  if (rc == CCM_OK) { h->poses.swap(poses); h->points.swap(points); }
  return rc;
}
void ccm_ba_destroy(ccm_ba_handle* h) { delete h; }

int ccm_pgo_solve(const ccm_pgo_problem* p, const ccm_pgo_options* o, ccm_pgo_result* r) {
  orc_pgo_problem op = {p->K, p->E, p->sim3, p->fixed, p->edge_i, p->edge_j, p->meas, p->fix_scale};
  orc_pgo_result rr = {};
  rr.sim3 = r->sim3;
  if (orc_pgo_solve(&op, o->iterations, o->lambda_init, 0, o->stop, &rr) != 0) return CCM_ERR_INVALID;
  r->iters_done = rr.iters_done; r->chi2_initial = rr.chi2_initial; r->chi2_final = rr.chi2_final; r->lambda_final = rr.lambda_final; r->trace_len = 0;
  return CCM_OK;
}
int ccm_pose_optimize(const ccm_pose_opt_problem* probs, int32_t batch, ccm_pose_opt_result* res) {
  for (int b = 0; b < batch; b++) {
    const ccm_pose_opt_problem& p = probs[b];
    orc_pose_opt_problem op = {p.n, p.Tcw, p.Xw, p.uv, p.inv_sigma2, p.fx, p.fy, p.cx, p.cy};
    res[b].n_inliers = orc_pose_optimize(&op, res[b].Tcw, res[b].outlier);
  }
  return CCM_OK;
}
int ccm_sim3_optimize(const ccm_sim3_opt_problem* probs, int32_t batch, ccm_sim3_opt_result* res) {
  for (int b = 0; b < batch; b++) {
    const ccm_sim3_opt_problem& p = probs[b];
    orc_sim3_opt_problem op = {};
    op.n = p.n; op.S12 = p.S12; op.P1c = p.P1c; op.P2c = p.P2c; op.uv1 = p.uv1; op.uv2 = p.uv2; op.inv_sigma2_1 = p.inv_sigma2_1;
    op.inv_sigma2_2 = p.inv_sigma2_2; std::memcpy(op.K1, p.K1, sizeof op.K1); std::memcpy(op.K2, p.K2, sizeof op.K2);
    op.th2 = p.th2; op.fix_scale = p.fix_scale;
    res[b].n_inliers = orc_sim3_optimize(&op, res[b].S12, res[b].inlier);
  }
  return CCM_OK;
}
}

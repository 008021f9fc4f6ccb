// ref_g2o_wrap.cpp — reference pin for the g2o vertex / edge types of the BA + Sim3 path (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Compiled together with the reference's OWN sources, where they lie (oracle/Makefile, target _ref/libg2o_types_ref.so):
//   G/types/types_sba.cpp, types_six_dof_expmap.cpp, types_seven_dof_expmap.cpp   (vertices, edges, analytic Jacobians)
//   G/types/se3quat.h, sim3.h, se3_ops.hpp                                        (header-only Lie algebra)
//   G/core/base_vertex.h(pp), base_edge.h, base_unary_edge.h(pp), base_binary_edge.h(pp)
//                                          (numeric Jacobians, constructQuadraticForm with the robust weighting, Hessian mapping)
//   G/core/robust_kernel.cpp, robust_kernel_impl.cpp                               (Huber)
// against two stand-ins: oracle/ref_stub/Eigen (the small fixed-size arithmetic, eager and in index order — see its header) and
// oracle/ref_stub/g2o_core_standin.h (the few members of OptimizableGraph::Vertex / Edge those files touch).  The functions below
// only build the reference's objects, call the reference's methods and copy results out in the oracle's layouts (row-major blocks,
// quaternion as x y z w followed by the translation), with the same signatures as the oracle's entry points so that
// tests/test_oracle_vs_reference_g2o.py can call either through one marshalling path.
#include <types/types_seven_dof_expmap.h>
#include <types/types_six_dof_expmap.h>
#include <core/robust_kernel_impl.h>

#include "oracle.h"

using namespace g2o;

namespace {

SE3Quat se3_in(const double* qt) {  // no normalisation on the way in: fromVector stores what it is given
  Vector7d v;
  v[0] = qt[4]; v[1] = qt[5]; v[2] = qt[6]; v[3] = qt[0]; v[4] = qt[1]; v[5] = qt[2]; v[6] = qt[3];
  SE3Quat T;
  T.fromVector(v);
  return T;
}
void se3_out(const SE3Quat& T, double* qt) {
  qt[0] = T.rotation().x(); qt[1] = T.rotation().y(); qt[2] = T.rotation().z(); qt[3] = T.rotation().w();
  for (int i = 0; i < 3; i++) qt[4 + i] = T.translation()[i];
}
Sim3 sim3_in(const double* s) { return Sim3(Quaterniond(s[3], s[0], s[1], s[2]), Vector3d(s[4], s[5], s[6]), s[7]); }
void sim3_out(const Sim3& S, double* s) {
  s[0] = S.rotation().x(); s[1] = S.rotation().y(); s[2] = S.rotation().z(); s[3] = S.rotation().w();
  for (int i = 0; i < 3; i++) s[4 + i] = S.translation()[i];
  s[7] = S.scale();
}
template <class M> void rowmajor(const M& m, double* out) {
  for (int i = 0; i < m.rows(); i++) for (int j = 0; j < m.cols(); j++) out[i * m.cols() + j] = m(i, j);
}

}  // namespace

extern "C" {

// ---- Lie groups -------------------------------------------------------------------------------------------------------
void ref_se3_exp(const double upd[6], double out[7]) { Vector6d u; for (int i = 0; i < 6; i++) u[i] = upd[i]; se3_out(SE3Quat::exp(u), out); }
void ref_se3_mul(const double a[7], const double b[7], double out[7]) { se3_out(se3_in(a) * se3_in(b), out); }
void ref_se3_map(const double qt[7], const double x[3], double out[3]) {
  Vector3d r = se3_in(qt).map(Vector3d(x[0], x[1], x[2]));
  for (int i = 0; i < 3; i++) out[i] = r[i];
}
void ref_se3_from_Rt(const double R[9], const double t[3], double out[7]) {  // SE3Quat(Matrix3d, Vector3d), what Converter::toSE3Quat builds
  Matrix3d m;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m(i, j) = R[i * 3 + j];
  se3_out(SE3Quat(m, Vector3d(t[0], t[1], t[2])), out);
}
void ref_se3_homogeneous(const double qt[7], double M[16]) { rowmajor(se3_in(qt).to_homogeneous_matrix(), M); }
void ref_se3_log(const double qt[7], double out[6]) { Vector6d l = se3_in(qt).log(); for (int i = 0; i < 6; i++) out[i] = l[i]; }
void ref_sim3_exp(const double upd[7], double out[8]) { Vector7d u; for (int i = 0; i < 7; i++) u[i] = upd[i]; sim3_out(Sim3(u), out); }
void ref_sim3_log(const double s[8], double out[7]) { Vector7d l = sim3_in(s).log(); for (int i = 0; i < 7; i++) out[i] = l[i]; }
void ref_sim3_mul(const double a[8], const double b[8], double out[8]) { sim3_out(sim3_in(a) * sim3_in(b), out); }
void ref_sim3_inv(const double a[8], double out[8]) { sim3_out(sim3_in(a).inverse(), out); }
void ref_sim3_from_Rt(const double R[9], const double t[3], const double s[1], double out[8]) {   // Sim3(Matrix3d, Vector3d, double): no normalisation
  Matrix3d m;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m(i, j) = R[i * 3 + j];
  sim3_out(Sim3(m, Vector3d(t[0], t[1], t[2]), s[0]), out);
}
void ref_sim3_map(const double s[8], const double x[3], double out[3]) {
  Vector3d r = sim3_in(s).map(Vector3d(x[0], x[1], x[2]));
  for (int i = 0; i < 3; i++) out[i] = r[i];
}
void ref_huber(double e, double delta, double rho[3]) {
  RobustKernelHuber k;
  k.setDelta(delta);
  Eigen::Vector3d r;
  k.robustify(e, r);
  for (int i = 0; i < 3; i++) rho[i] = r[i];
}
// vertex oplus: kind 0 = VertexSE3Expmap (7 in, 6 update), 1 = VertexSim3Expmap (8, 7; flag = _fix_scale), 2 = VertexSBAPointXYZ (3, 3)
void ref_vertex_oplus(int kind, const double* est, const double* upd, int flag, double* out) {
  if (kind == 0) { VertexSE3Expmap v; v.setEstimate(se3_in(est)); v.oplus(upd); se3_out(v.estimate(), out); }
  else if (kind == 1) {
    VertexSim3Expmap v; v._fix_scale = flag != 0; v.setEstimate(sim3_in(est));
    double u[7]; for (int i = 0; i < 7; i++) u[i] = upd[i];   // oplusImpl writes into the update when the scale is fixed
    v.oplus(u); sim3_out(v.estimate(), out);
  } else { VertexSBAPointXYZ v; v.setEstimate(Vector3d(est[0], est[1], est[2])); v.oplus(upd); for (int i = 0; i < 3; i++) out[i] = v.estimate()[i]; }
}

// ---- bundle adjustment: EdgeSE3ProjectXYZ over VertexSBAPointXYZ / VertexSE3Expmap -------------------------------------------
// The graph of S/Optimizer.cpp:700-787 for a flat problem; edges in index order.  `run` is what BlockSolver::buildSystem does per
// active edge (G/core/block_solver.hpp:518-540): linearizeOplus(workspace), constructQuadraticForm.
struct BAGraph {
  std::vector<VertexSE3Expmap*> kf;
  std::vector<VertexSBAPointXYZ*> mp;
  std::vector<EdgeSE3ProjectXYZ*> edge;
  std::vector<RobustKernelHuber*> kernel;
  BAGraph(const orc_ba_problem* p, int robust, double delta) {
    for (int k = 0; k < p->K; k++) {
      VertexSE3Expmap* v = new VertexSE3Expmap(); v->setEstimate(se3_in(p->poses + 7 * k)); v->setId(k); v->setFixed(p->fixed[k] != 0); kf.push_back(v);
    }
    for (int j = 0; j < p->P; j++) {
      VertexSBAPointXYZ* v = new VertexSBAPointXYZ(); v->setEstimate(Vector3d(p->points[3 * j], p->points[3 * j + 1], p->points[3 * j + 2]));
      v->setId(p->K + j); v->setMarginalized(true); mp.push_back(v);
    }
    for (int e = 0; e < p->E; e++) {
      EdgeSE3ProjectXYZ* ed = new EdgeSE3ProjectXYZ();
      ed->setVertex(0, mp[p->obs_mp[e]]); ed->setVertex(1, kf[p->obs_kf[e]]);
      Vector2d obs(p->obs_uv[2 * e], p->obs_uv[2 * e + 1]);          // float -> double, S/Optimizer.cpp:758-760
      ed->setMeasurement(obs);
      const float& invSigma2 = p->obs_w[e];
      ed->setInformation(Matrix2d::Identity() * invSigma2);
      const uint8_t fl = p->edge_flags ? p->edge_flags[e] : 0;
      ed->setLevel(fl & 1);
      RobustKernelHuber* rk = 0;
      if (robust && !(fl & 2)) { rk = new RobustKernelHuber; ed->setRobustKernel(rk); rk->setDelta(delta); }
      kernel.push_back(rk);
      const double* in = p->intr + 4 * p->obs_kf[e];
      ed->fx = in[0]; ed->fy = in[1]; ed->cx = in[2]; ed->cy = in[3];
      edge.push_back(ed);
    }
  }
  ~BAGraph() {
    for (size_t i = 0; i < edge.size(); i++) { delete edge[i]; delete kernel[i]; }
    for (size_t i = 0; i < kf.size(); i++) delete kf[i];
    for (size_t i = 0; i < mp.size(); i++) delete mp[i];
  }
};

double ref_ba_linearize(const orc_ba_problem* p, int robust, double huber_delta, double* err, double* Jpose, double* Jpoint,
                        double* rho1, double* chi2) {
  BAGraph g(p, robust, huber_delta);
  JacobianWorkspace ws;
  double total = 0;
  for (int e = 0; e < p->E; e++) {
    EdgeSE3ProjectXYZ* ed = g.edge[e];
    ed->computeError();
    // the analytic linearizeOplus of this edge ignores fixedness; both blocks are filled
    ed->BaseBinaryEdge<2, Vector2d, VertexSBAPointXYZ, VertexSE3Expmap>::linearizeOplus(ws);
    if (err) { err[2 * e] = ed->error()[0]; err[2 * e + 1] = ed->error()[1]; }
    if (Jpoint) rowmajor(ed->jacobianOplusXi(), Jpoint + 6 * (size_t)e);
    if (Jpose) rowmajor(ed->jacobianOplusXj(), Jpose + 12 * (size_t)e);
    const double c = ed->chi2();
    Eigen::Vector3d rho(c, 1, 0);
    if (ed->robustKernel()) ed->robustKernel()->robustify(c, rho);
    if (rho1) rho1[e] = rho[1];
    if (chi2) chi2[e] = c;
    if (ed->level() == 0) total += rho[0];   // SparseOptimizer::activeRobustChi2, G/core/sparse_optimizer.cpp:100-114
  }
  return total;
}

void ref_ba_build(const orc_ba_problem* p, int robust, double huber_delta, double* Hpp, double* bp, double* Hll, double* bl, double* W) {
  BAGraph g(p, robust, huber_delta);
  JacobianWorkspace ws;
  // column-major blocks owned here, mapped into the vertices / edges as BlockSolver::buildStructure does
  std::vector<double> hpp((size_t)p->K * 36, 0.), hll((size_t)p->P * 9, 0.), hpl((size_t)p->E * 18, 0.);
  for (int k = 0; k < p->K; k++) { g.kf[k]->mapHessianMemory(&hpp[(size_t)k * 36]); g.kf[k]->clearQuadraticForm(); }
  for (int j = 0; j < p->P; j++) { g.mp[j]->mapHessianMemory(&hll[(size_t)j * 9]); g.mp[j]->clearQuadraticForm(); }
  // the edge's vertex 0 is the marginalised point: BlockSolver::buildStructure hands it the Hpl block (pose, point), 6 x 3, and asks
  // for the transposed write (G/core/block_solver.hpp:240-244)
  for (int e = 0; e < p->E; e++) g.edge[e]->mapHessianMemory(&hpl[(size_t)e * 18], 0, 1, true);
  for (int e = 0; e < p->E; e++) {
    EdgeSE3ProjectXYZ* ed = g.edge[e];
    if (ed->level() != 0) continue;
    ed->computeError();
  }
  for (int e = 0; e < p->E; e++) {
    EdgeSE3ProjectXYZ* ed = g.edge[e];
    if (ed->level() != 0) continue;
    ed->BaseBinaryEdge<2, Vector2d, VertexSBAPointXYZ, VertexSE3Expmap>::linearizeOplus(ws);
    ed->constructQuadraticForm();
  }
  for (int k = 0; k < p->K; k++) {
    for (int i = 0; i < 6; i++) { bp[6 * (size_t)k + i] = g.kf[k]->b(i); for (int j = 0; j < 6; j++) Hpp[36 * (size_t)k + i * 6 + j] = g.kf[k]->hessian(i, j); }
  }
  for (int l = 0; l < p->P; l++)
    for (int i = 0; i < 3; i++) { bl[3 * (size_t)l + i] = g.mp[l]->b(i); for (int j = 0; j < 3; j++) Hll[9 * (size_t)l + i * 3 + j] = g.mp[l]->hessian(i, j); }
  for (int e = 0; e < p->E; e++)   // 6 x 3 column-major -> the oracle's 6 x 3 row-major
    for (int i = 0; i < 6; i++) for (int j = 0; j < 3; j++) W[18 * (size_t)e + i * 3 + j] = hpl[18 * (size_t)e + j * 6 + i];
}

// ---- pose-only optimisation: EdgeSE3ProjectXYZOnlyPose (S/Optimizer.cpp:215-347) -----------------------------------------------
void ref_pose_opt_build(const orc_pose_opt_problem* p, const double* Tcw, int robust, double delta, double* H, double* b, double* err) {
  VertexSE3Expmap v; v.setEstimate(se3_in(Tcw)); v.setId(0); v.setFixed(false);
  double h[36] = {0};
  v.mapHessianMemory(h); v.clearQuadraticForm();
  JacobianWorkspace ws;
  for (int i = 0; i < p->n; i++) {
    EdgeSE3ProjectXYZOnlyPose e;
    e.setVertex(0, &v);
    Vector2d obs(p->uv[2 * i], p->uv[2 * i + 1]);
    e.setMeasurement(obs);
    const float invSigma2 = p->inv_sigma2[i];
    e.setInformation(Matrix2d::Identity() * invSigma2);
    RobustKernelHuber rk;
    if (robust) { e.setRobustKernel(&rk); rk.setDelta(delta); }
    e.fx = p->fx; e.fy = p->fy; e.cx = p->cx; e.cy = p->cy;
    e.Xw[0] = p->Xw[3 * i]; e.Xw[1] = p->Xw[3 * i + 1]; e.Xw[2] = p->Xw[3 * i + 2];
    e.computeError();
    e.BaseUnaryEdge<2, Vector2d, VertexSE3Expmap>::linearizeOplus(ws);
    e.constructQuadraticForm();
    if (err) { err[2 * i] = e.error()[0]; err[2 * i + 1] = e.error()[1]; }
  }
  for (int i = 0; i < 6; i++) { b[i] = v.b(i); for (int j = 0; j < 6; j++) H[i * 6 + j] = v.hessian(i, j); }
}

// ---- Sim3 alignment of two keyframes: EdgeSim3ProjectXYZ + EdgeInverseSim3ProjectXYZ (S/Optimizer.cpp:861-1056) ----------------
// points are fixed vertices; the numeric Jacobian of BaseBinaryEdge::linearizeOplus is taken with respect to the Sim3 vertex only
void ref_sim3_opt_build(const orc_sim3_opt_problem* p, const double* S12, int robust, double delta, double* H, double* b, double* err) {
  VertexSim3Expmap v;
  v._fix_scale = p->fix_scale != 0;
  v.setEstimate(sim3_in(S12)); v.setId(0); v.setFixed(false);
  v._principle_point1[0] = p->K1[2]; v._principle_point1[1] = p->K1[3]; v._focal_length1[0] = p->K1[0]; v._focal_length1[1] = p->K1[1];
  v._principle_point2[0] = p->K2[2]; v._principle_point2[1] = p->K2[3]; v._focal_length2[0] = p->K2[0]; v._focal_length2[1] = p->K2[1];
  double h[49] = {0};
  v.mapHessianMemory(h); v.clearQuadraticForm();
  JacobianWorkspace ws;
  for (int i = 0; i < p->n; i++) {
    for (int side = 0; side < 2; side++) {
      VertexSBAPointXYZ pt;
      const float* X = side == 0 ? p->P2c + 3 * i : p->P1c + 3 * i;   // edge 2i projects P2 through S12, edge 2i+1 P1 through S12^-1
      pt.setEstimate(Vector3d(X[0], X[1], X[2])); pt.setId(1); pt.setFixed(true);
      double hl[9] = {0}, hpl[21] = {0};  // the point is fixed: nothing is written to either
      pt.mapHessianMemory(hl); pt.clearQuadraticForm();
      const float* uv = side == 0 ? p->uv1 + 2 * i : p->uv2 + 2 * i;
      Vector2d obs(uv[0], uv[1]);
      const float invSigma2 = side == 0 ? p->inv_sigma2_1[i] : p->inv_sigma2_2[i];
      RobustKernelHuber rk;
      double* eo = err ? err + 2 * (2 * i + side) : 0;
      if (side == 0) {
        EdgeSim3ProjectXYZ e;
        e.setVertex(0, &pt); e.setVertex(1, &v); e.setMeasurement(obs); e.setInformation(Matrix2d::Identity() * invSigma2);
        if (robust) { e.setRobustKernel(&rk); rk.setDelta(delta); }
        e.mapHessianMemory(hpl, 0, 1, true);
        e.computeError(); e.linearizeOplus(ws); e.constructQuadraticForm();
        if (eo) { eo[0] = e.error()[0]; eo[1] = e.error()[1]; }
      } else {
        EdgeInverseSim3ProjectXYZ e;
        e.setVertex(0, &pt); e.setVertex(1, &v); e.setMeasurement(obs); e.setInformation(Matrix2d::Identity() * invSigma2);
        if (robust) { e.setRobustKernel(&rk); rk.setDelta(delta); }
        e.mapHessianMemory(hpl, 0, 1, true);
        e.computeError(); e.linearizeOplus(ws); e.constructQuadraticForm();
        if (eo) { eo[0] = e.error()[0]; eo[1] = e.error()[1]; }
      }
    }
  }
  for (int i = 0; i < 7; i++) { b[i] = v.b(i); for (int j = 0; j < 7; j++) H[i * 7 + j] = v.hessian(i, j); }
}

// ---- essential graph: EdgeSim3 between two VertexSim3Expmap (S/Optimizer.cpp:1060-1290) -----------------------------------------
void ref_pgo_edge_error(const double meas[8], const double si[8], const double sj[8], double err[7]) {
  VertexSim3Expmap vi, vj; vi.setEstimate(sim3_in(si)); vj.setEstimate(sim3_in(sj));
  EdgeSim3 e; e.setVertex(0, &vi); e.setVertex(1, &vj); e.setMeasurement(sim3_in(meas));
  e.computeError();
  for (int i = 0; i < 7; i++) err[i] = e.error()[i];
}
void ref_pgo_edge_jacobian(const double meas[8], const double si[8], const double sj[8], int fix_scale, double Ji[49], double Jj[49]) {
  VertexSim3Expmap vi, vj;
  vi._fix_scale = vj._fix_scale = fix_scale != 0;
  vi.setEstimate(sim3_in(si)); vj.setEstimate(sim3_in(sj)); vi.setFixed(false); vj.setFixed(false);
  EdgeSim3 e; e.setVertex(0, &vi); e.setVertex(1, &vj); e.setMeasurement(sim3_in(meas));
  JacobianWorkspace ws;
  e.computeError();
  e.linearizeOplus(ws);
  rowmajor(e.jacobianOplusXi(), Ji);
  rowmajor(e.jacobianOplusXj(), Jj);
}

}  // extern "C"

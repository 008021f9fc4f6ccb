// ref_ba_full_wrap.cpp — a whole bundle-adjustment optimize() with the reference's code on both sides of the linear solve
// (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Puts the two reference pins together: g2o's own Levenberg-Marquardt driver (optimization_algorithm{,_with_hessian,_levenberg}.cpp,
// included textually below, as in ref_lm_wrap.cpp) runs over g2o's own graph elements (VertexSE3Expmap, VertexSBAPointXYZ,
// EdgeSE3ProjectXYZ, RobustKernelHuber with their base templates: errors, Jacobians, constructQuadraticForm into mapped blocks, oplus,
// the backup stack behind push / pop — compiled from the reference tree against oracle/ref_stub/Eigen, as in ref_g2o_wrap.cpp).
// What is still the oracle's: the index mapping / active set (init_active), and everything under Solver::solve() — the Schur
// complement and the LDL^T of BlockSolver / LinearSolverEigen (ba_oracle.cpp solve_system).  The glue restates, a few lines each,
// SparseOptimizer::{optimize, computeActiveErrors, activeRobustChi2, update, push, pop} (G/core/sparse_optimizer.cpp:61-114,354-435,
// 600-613) and BlockSolver::{buildStructure's memory mapping, buildSystem} (G/core/block_solver.hpp:175-250,501-560).
// ref_ba_full_solve() has orc_ba_solve()'s signature; tests/test_oracle_vs_reference_lm.py compares the two runs bit for bit.
#include "ba_oracle.cpp"  // the oracle's pieces; its unnamed namespace is visible here

#include <iomanip>
#include <iostream>

#define G2O_SPARSE_BLOCK_MATRIX_
#define G2O_SOLVER_H
#define G2O_GRAPH_OPTIMIZER_CHOL_H_
#include <core/batch_stats.h>
#include <core/hyper_graph.h>
#include <core/robust_kernel_impl.h>
#include <stuff/macros.h>
#include <types/types_six_dof_expmap.h>

namespace g2o {

class MatrixXd;
template <class M> class SparseBlockMatrix;
class OptimizationAlgorithm;

class SparseOptimizer : public OptimizableGraph {
 public:
  SparseOptimizer(BA& s, const orc_ba_problem* p, int robust, double delta, const volatile uint8_t* stop)
      : s(s), stop_(stop), algorithm_(0), last_chi(0), chi_at_push(0), first_chi_(0), have_first_(false), trials(0) {
    for (int k = 0; k < p->K; k++) {   // S/Optimizer.cpp:700-712
      VertexSE3Expmap* v = new VertexSE3Expmap();
      Vector7d est; const double* q = p->poses + 7 * k;
      est[0] = q[4]; est[1] = q[5]; est[2] = q[6]; est[3] = q[0]; est[4] = q[1]; est[5] = q[2]; est[6] = q[3];
      SE3Quat T; T.fromVector(est);
      v->setEstimate(T); v->setId(k); v->setFixed(p->fixed[k] != 0);
      kf.push_back(v);
    }
    for (int j = 0; j < p->P; j++) {   // :740-747
      VertexSBAPointXYZ* v = new VertexSBAPointXYZ();
      v->setEstimate(Vector3d(p->points[3 * j], p->points[3 * j + 1], p->points[3 * j + 2]));
      v->setId(p->K + j); v->setMarginalized(true);
      mp.push_back(v);
    }
    for (int e = 0; e < p->E; e++) {   // :750-784
      EdgeSE3ProjectXYZ* ed = new EdgeSE3ProjectXYZ();
      ed->setVertex(0, mp[p->obs_mp[e]]); ed->setVertex(1, kf[p->obs_kf[e]]);
      ed->setMeasurement(Vector2d(p->obs_uv[2 * e], p->obs_uv[2 * e + 1]));
      const float& invSigma2 = p->obs_w[e];
      ed->setInformation(Matrix2d::Identity() * invSigma2);
      const uint8_t fl = p->edge_flags ? p->edge_flags[e] : 0;
      ed->setLevel(fl & 1);
      RobustKernelHuber* rk = 0;
      if (robust && !(fl & 2)) { rk = new RobustKernelHuber; ed->setRobustKernel(rk); rk->setDelta(delta); }
      kernels.push_back(rk);
      const double* in = p->intr + 4 * p->obs_kf[e];
      ed->fx = in[0]; ed->fy = in[1]; ed->cx = in[2]; ed->cy = in[3];
      edges.push_back(ed);
    }
    // initializeOptimization(0): level-0 edges in id order; index mapping = free poses, then points (the oracle's init_active)
    for (int e : s.active) active.push_back(edges[e]);
    for (int i = 0; i < s.np; i++) iv_.push_back(kf[s.idx_pose[i]]);
    for (int l = 0; l < s.nl; l++) iv_.push_back(mp[s.idx_pt[l]]);
  }
  ~SparseOptimizer() {
    for (size_t i = 0; i < edges.size(); i++) { delete edges[i]; delete kernels[i]; }
    for (size_t i = 0; i < kf.size(); i++) delete kf[i];
    for (size_t i = 0; i < mp.size(); i++) delete mp[i];
  }
  const VertexContainer& indexMapping() const { return iv_; }
  const VertexContainer& activeVertices() const { return iv_; }
  void computeActiveErrors() { for (size_t k = 0; k < active.size(); k++) active[k]->computeError(); }
  double activeRobustChi2() {
    Eigen::Vector3d rho;
    double chi = 0.0;
    for (size_t k = 0; k < active.size(); k++) {
      const EdgeSE3ProjectXYZ* e = active[k];
      if (e->robustKernel()) { e->robustKernel()->robustify(e->chi2(), rho); chi += rho[0]; }
      else chi += e->chi2();
    }
    last_chi = chi;
    if (!have_first_) { first_chi_ = chi; have_first_ = true; }
    return chi;
  }
  void push() { for (size_t i = 0; i < iv_.size(); i++) iv_[i]->push(); chi_at_push = last_chi; trials++; }
  void pop() { for (size_t i = 0; i < iv_.size(); i++) iv_[i]->pop(); last_chi = chi_at_push; }
  void discardTop() { for (size_t i = 0; i < iv_.size(); i++) iv_[i]->discardTop(); }
  void update(const double* update) { for (size_t i = 0; i < iv_.size(); ++i) { iv_[i]->oplus(update); update += iv_[i]->dimension(); } }
  bool terminate() { return stop_ && *stop_; }
  void setAlgorithm(OptimizationAlgorithm* a);
  int optimize(int iterations, orc_ba_result* r);
  BA& s;
  const volatile uint8_t* stop_;
  OptimizationAlgorithm* algorithm_;
  std::vector<VertexSE3Expmap*> kf;
  std::vector<VertexSBAPointXYZ*> mp;
  std::vector<EdgeSE3ProjectXYZ*> edges, active;
  std::vector<RobustKernelHuber*> kernels;
  VertexContainer iv_;
  double last_chi, chi_at_push, first_chi_;
  bool have_first_;
  int trials;
  JacobianWorkspace workspace;
};

class Solver {
 public:
  explicit Solver(BA& s) : s(s), opt_(0), lambda_(0), schur_(false) {}
  virtual ~Solver() {}
  bool init(SparseOptimizer* o, bool) { opt_ = o; return true; }
  SparseOptimizer* optimizer() const { return opt_; }
  bool buildStructure(bool = false) {
    build_structure(s);   // the oracle's Schur pattern and landmark columns, used by solve_system below
    hpp.assign((size_t)s.np * 36, 0.); hll.assign((size_t)s.nl * 9, 0.); hpl.assign(s.active.size() * 18, 0.);
    b_.assign((size_t)s.np * 6 + (size_t)s.nl * 3, 0.);
    for (int i = 0; i < s.np; i++) opt_->iv_[i]->mapHessianMemory(&hpp[(size_t)i * 36]);
    for (int l = 0; l < s.nl; l++) opt_->iv_[s.np + l]->mapHessianMemory(&hll[(size_t)l * 9]);
    for (size_t a = 0; a < opt_->active.size(); a++)   // edges to a fixed pose get no block (ind == -1: continue)
      if (!opt_->active[a]->vertex(1)->fixed()) opt_->active[a]->mapHessianMemory(&hpl[a * 18], 0, 1, true);
    return true;
  }
  bool updateStructure(const std::vector<HyperGraph::Vertex*>&, const HyperGraph::EdgeSet&) { return false; }
  bool buildSystem() {
    for (size_t i = 0; i < opt_->iv_.size(); ++i) opt_->iv_[i]->clearQuadraticForm();
    std::fill(hpp.begin(), hpp.end(), 0.); std::fill(hll.begin(), hll.end(), 0.); std::fill(hpl.begin(), hpl.end(), 0.);
    for (size_t k = 0; k < opt_->active.size(); ++k) {
      EdgeSE3ProjectXYZ* e = opt_->active[k];
      e->BaseBinaryEdge<2, Vector2d, VertexSBAPointXYZ, VertexSE3Expmap>::linearizeOplus(opt_->workspace);
      e->constructQuadraticForm();
    }
    double* b = b_.data();
    for (size_t i = 0; i < opt_->iv_.size(); ++i) b += opt_->iv_[i]->copyB(b);
    return true;
  }
  bool setLambda(double lambda, bool = false) { lambda_ = lambda; return true; }
  void restoreDiagonal() {}
  bool solve() {
    // hand the system to the oracle's Schur complement + LDL^T in its own layout (row-major blocks)
    for (int i = 0; i < s.np; i++)
      for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) s.Hpp[(size_t)i * 36 + r * 6 + c] = hpp[(size_t)i * 36 + c * 6 + r];
    for (int l = 0; l < s.nl; l++)
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) s.Hll[(size_t)l * 9 + r * 3 + c] = hll[(size_t)l * 9 + c * 3 + r];
    for (size_t a = 0; a < s.active.size(); a++)
      for (int r = 0; r < 6; r++) for (int c = 0; c < 3; c++) s.Hpl[a * 18 + r * 3 + c] = hpl[a * 18 + c * 6 + r];
    std::copy(b_.begin(), b_.begin() + (size_t)s.np * 6, s.bp.begin());
    std::copy(b_.begin() + (size_t)s.np * 6, b_.end(), s.bl.begin());
    return solve_system(s, lambda_);
  }
  bool computeMarginals(SparseBlockMatrix<MatrixXd>&, const std::vector<std::pair<int, int> >&) { return false; }
  double* x() { return s.x.data(); }
  double* b() { return b_.data(); }
  size_t vectorSize() const { return b_.size(); }
  bool schur() { return schur_; }
  bool supportsSchur() { return true; }
  void setSchur(bool v) { schur_ = v; }
  void setWriteDebug(bool) {}
  BA& s;
  SparseOptimizer* opt_;
  double lambda_;
  bool schur_;
  std::vector<double> hpp, hll, hpl, b_;
};

}  // namespace g2o

// the reference's Levenberg-Marquardt sources, compiled in place
#include <core/optimization_algorithm.cpp>
#include <core/optimization_algorithm_with_hessian.cpp>
#include <core/optimization_algorithm_levenberg.cpp>

namespace g2o {

void SparseOptimizer::setAlgorithm(OptimizationAlgorithm* a) { algorithm_ = a; a->setOptimizer(this); }

int SparseOptimizer::optimize(int iterations, orc_ba_result* r) {  // G/core/sparse_optimizer.cpp:354-419
  if (iv_.size() == 0 || active.empty()) return -1;
  OptimizationAlgorithmLevenberg* lm = static_cast<OptimizationAlgorithmLevenberg*>(algorithm_);
  int cjIterations = 0;
  bool ok = algorithm_->init(false);
  if (!ok) return -1;
  OptimizationAlgorithm::SolverResult result = OptimizationAlgorithm::OK;
  for (int i = 0; i < iterations && !terminate() && ok; i++) {
    const int trials_before = trials;
    result = algorithm_->solve(i, false);
    ok = (result == OptimizationAlgorithm::OK);
    if (i == 0) r->chi2_initial = first_chi_;
    if (r->trace && r->trace_len < r->trace_cap) {
      double* tr = r->trace + (size_t)r->trace_len * ORC_TRACE_COLS;
      tr[0] = i; tr[1] = static_cast<Solver*>(lm->solver())->lambda_; tr[2] = last_chi;
      tr[3] = std::numeric_limits<double>::quiet_NaN();
      tr[4] = lm->levenbergIteration(); tr[5] = lm->currentLambda();
      r->trace_len++;
    }
    r->trials_total += trials - trials_before;
    r->chi2_final = last_chi; r->lambda_final = lm->currentLambda();
    ++cjIterations;
  }
  if (result == OptimizationAlgorithm::Fail) return 0;
  return cjIterations;
}

}  // namespace g2o

extern "C" int ref_ba_full_solve(const orc_ba_problem* p, const orc_ba_options* o, orc_ba_result* r) {
  BA s;
  load(s, p, o->robust, o->huber_delta);
  init_active(s);
  r->trace_len = 0; r->iters_done = 0; r->trials_total = 0;
  r->chi2_initial = r->chi2_final = 0; r->lambda_final = 0;
  g2o::SparseOptimizer optimizer(s, p, o->robust, o->huber_delta, o->stop);
  {
    g2o::OptimizationAlgorithmLevenberg* lm = new g2o::OptimizationAlgorithmLevenberg(new g2o::Solver(s));
    if (o->lambda_init > 0) lm->setUserLambdaInit(o->lambda_init);
    if (o->max_trials > 0) lm->setMaxTrialsAfterFailure(o->max_trials);
    optimizer.setAlgorithm(lm);
    r->iters_done = optimizer.optimize(o->iterations, r);
    delete lm;
  }
  // read the result out of the reference's vertices and edges (S/Optimizer.cpp:803-857 reads estimate(); :540-566 chi2(), isDepthPositive())
  for (int k = 0; k < p->K; k++) {
    const g2o::SE3Quat& T = optimizer.kf[k]->estimate();
    double* q = r->poses + 7 * (size_t)k;
    q[0] = T.rotation().x(); q[1] = T.rotation().y(); q[2] = T.rotation().z(); q[3] = T.rotation().w();
    for (int i = 0; i < 3; i++) q[4 + i] = T.translation()[i];
  }
  for (int j = 0; j < p->P; j++) for (int i = 0; i < 3; i++) r->points[3 * (size_t)j + i] = optimizer.mp[j]->estimate()[i];
  if (r->chi2) for (int e : s.active) r->chi2[e] = optimizer.edges[e]->chi2();
  if (r->depth_pos) for (int e = 0; e < p->E; e++) r->depth_pos[e] = optimizer.edges[e]->isDepthPositive();
  return 0;
}

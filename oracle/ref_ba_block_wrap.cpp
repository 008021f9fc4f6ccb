// ref_ba_block_wrap.cpp — a whole bundle-adjustment optimize() in which only the sparse factorisation is not the reference's
// (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// One step further than ref_ba_full_wrap.cpp: next to g2o's own Levenberg-Marquardt driver and graph elements, g2o's own
// BlockSolver_6_3 — buildStructure (block allocation, Hschur pattern, the transposed pose-landmark blocks), buildSystem, setLambda /
// restoreDiagonal, and solve() with the Schur complement and the landmark back-substitution (G/core/block_solver.h(pp), solver.cpp,
// sparse_block_matrix*.h(pp), matrix_operations.h, compiled where they lie against oracle/ref_stub/Eigen) — runs as the reference's
// code.  The oracle supplies one thing: LinearSolver::solve on the reduced camera system (the sparse LDL^T of sparse_ldlt.hpp, standing
// in for LinearSolverEigen, G/solvers/linear_solver_eigen.h:106-133), fed from the reference's SparseBlockMatrix.  The glue restates
// what ref_ba_full_wrap.cpp restates of SparseOptimizer (optimize, computeActiveErrors, activeRobustChi2, update, push, pop, the
// index mapping with its hessian indices).  ref_ba_block_solve() has orc_ba_solve()'s signature.
#include "ba_oracle.cpp"  // BA::load / init_active for the active set and index mapping, BlockSym + SparseLDLT; unnamed namespace visible here

#include <iomanip>
#include <iostream>

#define G2O_GRAPH_OPTIMIZER_CHOL_H_
#include <core/batch_stats.h>
#include <core/hyper_graph.h>
#include <core/robust_kernel_impl.h>
#include <stuff/macros.h>
#include <types/types_six_dof_expmap.h>

namespace g2o {

class OptimizationAlgorithm;

class SparseOptimizer : public OptimizableGraph {
 public:
  typedef OptimizableGraph::EdgeContainer EdgeContainer;
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
      mp[p->obs_mp[e]]->edges().insert(ed); kf[p->obs_kf[e]]->edges().insert(ed);   // HyperGraph::addEdge (G/core/hyper_graph.cpp:80-95)
      edges.push_back(ed);
    }
    // initializeOptimization(0) + buildIndexMapping (G/core/sparse_optimizer.cpp:166-267): level-0 edges in id order; free poses then
    // points, each given its position as hessian index; fixed vertices keep -1
    for (int e : s.active) active.push_back(edges[e]);
    for (int i = 0; i < s.np; i++) iv_.push_back(kf[s.idx_pose[i]]);
    for (int l = 0; l < s.nl; l++) iv_.push_back(mp[s.idx_pt[l]]);
    for (size_t i = 0; i < iv_.size(); i++) iv_[i]->setHessianIndex((int)i);
  }
  ~SparseOptimizer() {
    for (size_t i = 0; i < edges.size(); i++) { delete edges[i]; delete kernels[i]; }
    for (size_t i = 0; i < kf.size(); i++) delete kf[i];
    for (size_t i = 0; i < mp.size(); i++) delete mp[i];
  }
  const VertexContainer& indexMapping() const { return iv_; }
  const VertexContainer& activeVertices() const { return iv_; }
  const EdgeContainer& activeEdges() const { return active; }
  JacobianWorkspace& jacobianWorkspace() { return workspace; }
  void computeActiveErrors() { for (size_t k = 0; k < active.size(); k++) active[k]->computeError(); }
  double activeRobustChi2() {
    Eigen::Vector3d rho;
    double chi = 0.0;
    for (size_t k = 0; k < active.size(); k++) {
      const OptimizableGraph::Edge* e = active[k];
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
  bool verbose() const { return false; }
  void setAlgorithm(OptimizationAlgorithm* a);
  int optimize(int iterations, orc_ba_result* r);
  BA& s;
  const volatile uint8_t* stop_;
  OptimizationAlgorithm* algorithm_;
  std::vector<VertexSE3Expmap*> kf;
  std::vector<VertexSBAPointXYZ*> mp;
  std::vector<EdgeSE3ProjectXYZ*> edges;
  EdgeContainer active;
  std::vector<RobustKernelHuber*> kernels;
  VertexContainer iv_;
  double last_chi, chi_at_push, first_chi_;
  bool have_first_;
  int trials;
  JacobianWorkspace workspace;
};

}  // namespace g2o

#include <core/block_solver.h>   // the reference's own: Solver, BlockSolver<Traits>, SparseBlockMatrix*, LinearSolver

namespace g2o {

// LinearSolverEigen's place: the reduced camera system arrives as the reference's SparseBlockMatrix (upper triangle, column maps of
// column-major 6x6 blocks) and is factorised by the oracle's sparse LDL^T
class OracleLDLT : public LinearSolver<BlockSolver_6_3::PoseMatrixType> {
 public:
  typedef BlockSolver_6_3::PoseMatrixType M;
  OracleLDLT() : analyzed_(false), last_lambda_(0) {}
  virtual bool init() { analyzed_ = false; return true; }
  virtual bool solve(const SparseBlockMatrix<M>& A, double* x, double* b) {
    const int n = (int)A.blockCols().size();
    std::vector<std::vector<std::pair<int, const M*> > > rows(n);
    for (int j = 0; j < n; j++)
      for (SparseBlockMatrix<M>::IntBlockMap::const_iterator it = A.blockCols()[j].begin(); it != A.blockCols()[j].end(); ++it)
        if (it->first <= j) rows[it->first].push_back(std::make_pair(j, (const M*)it->second));
    S_.nb = n; S_.bs = 6; S_.rowptr.assign(n + 1, 0); S_.col.clear();
    for (int i = 0; i < n; i++) { for (size_t q = 0; q < rows[i].size(); q++) S_.col.push_back(rows[i][q].first); S_.rowptr[i + 1] = (int)S_.col.size(); }
    S_.val.assign(S_.col.size() * 36, 0.);
    size_t q = 0;
    for (int i = 0; i < n; i++)
      for (size_t k = 0; k < rows[i].size(); k++, q++)
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) S_.val[q * 36 + r * 6 + c] = (*rows[i][k].second)(r, c);
    if (!analyzed_) { ldlt_.analyze(S_); analyzed_ = true; }
    if (!ldlt_.factorize(S_)) return false;
    ldlt_.solve(b, x);
    return true;
  }
 private:
  BlockSym S_;
  SparseLDLT ldlt_;
  bool analyzed_;
  double last_lambda_;
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
    const double lambda_before = lm->currentLambda();
    result = algorithm_->solve(i, false);
    ok = (result == OptimizationAlgorithm::OK);
    if (i == 0) r->chi2_initial = first_chi_;
    if (r->trace && r->trace_len < r->trace_cap) {
      double* tr = r->trace + (size_t)r->trace_len * ORC_TRACE_COLS;
      tr[0] = i; tr[1] = std::numeric_limits<double>::quiet_NaN();   // the lambda of the last trial lives inside the real BlockSolver
      tr[2] = last_chi; tr[3] = std::numeric_limits<double>::quiet_NaN();
      tr[4] = lm->levenbergIteration(); tr[5] = lm->currentLambda();
      r->trace_len++;
      (void)lambda_before;
    }
    r->trials_total += trials - trials_before;
    r->chi2_final = last_chi; r->lambda_final = lm->currentLambda();
    ++cjIterations;
  }
  if (result == OptimizationAlgorithm::Fail) return 0;
  return cjIterations;
}

}  // namespace g2o

extern "C" int ref_ba_block_solve(const orc_ba_problem* p, const orc_ba_options* o, orc_ba_result* r) {
  BA s;
  load(s, p, o->robust, o->huber_delta);
  init_active(s);
  r->trace_len = 0; r->iters_done = 0; r->trials_total = 0;
  r->chi2_initial = r->chi2_final = 0; r->lambda_final = 0;
  g2o::SparseOptimizer optimizer(s, p, o->robust, o->huber_delta, o->stop);
  {
    g2o::BlockSolver_6_3* solver_ptr = new g2o::BlockSolver_6_3(new g2o::OracleLDLT());   // S/Optimizer.cpp:681-687 with the solver swapped
    g2o::OptimizationAlgorithmLevenberg* lm = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
    if (o->lambda_init > 0) lm->setUserLambdaInit(o->lambda_init);
    if (o->max_trials > 0) lm->setMaxTrialsAfterFailure(o->max_trials);
    optimizer.setAlgorithm(lm);
    r->iters_done = optimizer.optimize(o->iterations, r);
    delete lm;
  }
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

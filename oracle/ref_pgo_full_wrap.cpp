// ref_pgo_full_wrap.cpp — the essential-graph optimisation with the reference's code around the sparse solve
// (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Same construction as ref_ba_full_wrap.cpp for Optimizer::OptimizeEssentialGraph's graph (S/Optimizer.cpp:1060-1290): g2o's own
// Levenberg-Marquardt driver over g2o's own VertexSim3Expmap / EdgeSim3 (errors through Sim3::log, numeric Jacobians of
// BaseBinaryEdge::linearizeOplus, constructQuadraticForm into mapped 7x7 blocks, oplus with _fix_scale, push / pop), compiled from
// the reference tree against oracle/ref_stub/Eigen.  The oracle supplies the block pattern and the sparse LDL^T under
// Solver::solve() (pgo_oracle.cpp / sparse_ldlt.hpp, standing in for BlockSolver_7_3 + LinearSolverEigen).  The glue restates the
// upper-triangle block allocation with the transposed write (G/core/block_solver.hpp:215-232) and, as in the other wrappers,
// SparseOptimizer::{optimize, computeActiveErrors, activeRobustChi2, update, push, pop} and BlockSolver::buildSystem.
// ref_pgo_solve() has orc_pgo_solve()'s signature.
#include "pgo_oracle.cpp"

#include <iomanip>
#include <iostream>

#define G2O_SPARSE_BLOCK_MATRIX_
#define G2O_SOLVER_H
#define G2O_GRAPH_OPTIMIZER_CHOL_H_
#include <core/batch_stats.h>
#include <core/hyper_graph.h>
#include <stuff/macros.h>
#include <types/types_seven_dof_expmap.h>

namespace g2o {

class MatrixXd;
template <class M> class SparseBlockMatrix;
class OptimizationAlgorithm;

class SparseOptimizer : public OptimizableGraph {
 public:
  explicit SparseOptimizer(const volatile uint8_t* stop) : stop_(stop), algorithm_(0), last_chi(0), chi_at_push(0), first_chi_(0), have_first_(false) {}
  const VertexContainer& indexMapping() const { return iv_; }
  const VertexContainer& activeVertices() const { return iv_; }
  void computeActiveErrors() { for (size_t k = 0; k < active.size(); k++) active[k]->computeError(); }
  double activeRobustChi2() {
    double chi = 0.0;
    for (size_t k = 0; k < active.size(); k++) chi += active[k]->chi2();   // no kernels on this graph
    last_chi = chi;
    if (!have_first_) { first_chi_ = chi; have_first_ = true; }
    return chi;
  }
  void push() { for (size_t i = 0; i < iv_.size(); i++) iv_[i]->push(); chi_at_push = last_chi; }
  void pop() { for (size_t i = 0; i < iv_.size(); i++) iv_[i]->pop(); last_chi = chi_at_push; }
  void discardTop() { for (size_t i = 0; i < iv_.size(); i++) iv_[i]->discardTop(); }
  void update(const double* update) { for (size_t i = 0; i < iv_.size(); ++i) { iv_[i]->oplus(update); update += iv_[i]->dimension(); } }
  bool terminate() { return stop_ && *stop_; }
  void setAlgorithm(OptimizationAlgorithm* a);
  int optimize(int iterations, orc_pgo_result* r);
  const volatile uint8_t* stop_;
  OptimizationAlgorithm* algorithm_;
  VertexContainer iv_;
  std::vector<EdgeSim3*> active;
  double last_chi, chi_at_push, first_chi_;
  bool have_first_;
  JacobianWorkspace workspace;
};

class Solver {
 public:
  Solver(PGO& s, const std::vector<VertexSim3Expmap*>& v) : s(s), v_(v), opt_(0), lambda_(0) {}
  virtual ~Solver() {}
  bool init(SparseOptimizer* o, bool) { opt_ = o; return true; }
  SparseOptimizer* optimizer() const { return opt_; }
  bool buildStructure(bool = false) {
    // pattern: diagonal + one upper-triangle block per connected pair of free vertices
    std::vector<std::vector<int>> rows(s.n);
    for (int i = 0; i < s.n; i++) rows[i].push_back(i);
    for (int e : s.active) {
      int a = s.vidx[s.edges[e].i], b = s.vidx[s.edges[e].j];
      if (a >= 0 && b >= 0 && a != b) rows[std::min(a, b)].push_back(std::max(a, b));
    }
    s.H.nb = s.n; s.H.bs = 7; s.H.rowptr.assign(s.n + 1, 0); s.H.col.clear();
    for (int i = 0; i < s.n; i++) {
      std::sort(rows[i].begin(), rows[i].end());
      rows[i].erase(std::unique(rows[i].begin(), rows[i].end()), rows[i].end());
      s.H.col.insert(s.H.col.end(), rows[i].begin(), rows[i].end());
      s.H.rowptr[i + 1] = (int)s.H.col.size();
    }
    s.H.val.assign(s.H.col.size() * 49, 0);
    s.ldlt.analyze(s.H);
    s.b.assign((size_t)s.n * 7, 0); s.x.assign((size_t)s.n * 7, 0);
    blocks.assign(s.H.col.size() * 49, 0.);   // column-major 7x7 blocks in the pattern's order, handed to vertices and edges
    for (int i = 0; i < s.n; i++) opt_->iv_[i]->mapHessianMemory(&blocks[(size_t)s.H.find(i, i) * 49]);
    for (size_t k = 0; k < opt_->active.size(); k++) {
      const PEdge& ed = s.edges[s.active[k]];
      int ind1 = s.vidx[ed.i], ind2 = s.vidx[ed.j];
      if (ind1 < 0 || ind2 < 0) continue;
      bool transposedBlock = ind1 > ind2;
      if (transposedBlock) std::swap(ind1, ind2);
      opt_->active[k]->mapHessianMemory(&blocks[(size_t)s.H.find(ind1, ind2) * 49], 0, 1, transposedBlock);
    }
    return true;
  }
  bool updateStructure(const std::vector<HyperGraph::Vertex*>&, const HyperGraph::EdgeSet&) { return false; }
  bool buildSystem() {
    for (size_t i = 0; i < opt_->iv_.size(); ++i) opt_->iv_[i]->clearQuadraticForm();
    std::fill(blocks.begin(), blocks.end(), 0.);
    for (size_t k = 0; k < opt_->active.size(); ++k) {
      opt_->active[k]->linearizeOplus(opt_->workspace);
      opt_->active[k]->constructQuadraticForm();
    }
    double* b = s.b.data();
    for (size_t i = 0; i < opt_->iv_.size(); ++i) b += opt_->iv_[i]->copyB(b);
    return true;
  }
  bool setLambda(double lambda, bool = false) { lambda_ = lambda; return true; }
  void restoreDiagonal() {}
  bool solve() {
    for (size_t q = 0; q < s.H.col.size(); q++)
      for (int r = 0; r < 7; r++) for (int c = 0; c < 7; c++) s.H.val[q * 49 + r * 7 + c] = blocks[q * 49 + c * 7 + r];
    for (int i = 0; i < s.n; i++) {
      double* d = &s.H.val[(size_t)s.H.find(i, i) * 49];
      for (int k = 0; k < 7; k++) d[k * 8] += lambda_;
    }
    bool ok = s.ldlt.factorize(s.H);
    if (ok) s.ldlt.solve(s.b.data(), s.x.data());
    return ok;
  }
  bool computeMarginals(SparseBlockMatrix<MatrixXd>&, const std::vector<std::pair<int, int> >&) { return false; }
  double* x() { return s.x.data(); }
  double* b() { return s.b.data(); }
  size_t vectorSize() const { return s.b.size(); }
  bool schur() { return false; }
  bool supportsSchur() { return true; }
  void setSchur(bool) {}
  void setWriteDebug(bool) {}
  PGO& s;
  const std::vector<VertexSim3Expmap*>& v_;
  SparseOptimizer* opt_;
  double lambda_;
  std::vector<double> blocks;
};

}  // namespace g2o

#include <core/optimization_algorithm.cpp>
#include <core/optimization_algorithm_with_hessian.cpp>
#include <core/optimization_algorithm_levenberg.cpp>

namespace g2o {

void SparseOptimizer::setAlgorithm(OptimizationAlgorithm* a) { algorithm_ = a; a->setOptimizer(this); }

int SparseOptimizer::optimize(int iterations, orc_pgo_result* r) {  // G/core/sparse_optimizer.cpp:354-419
  if (iv_.size() == 0 || active.empty()) return -1;
  OptimizationAlgorithmLevenberg* lm = static_cast<OptimizationAlgorithmLevenberg*>(algorithm_);
  int cjIterations = 0;
  bool ok = algorithm_->init(false);
  if (!ok) return -1;
  OptimizationAlgorithm::SolverResult result = OptimizationAlgorithm::OK;
  for (int i = 0; i < iterations && !terminate() && ok; i++) {
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
    r->chi2_final = last_chi; r->lambda_final = lm->currentLambda();
    ++cjIterations;
  }
  if (result == OptimizationAlgorithm::Fail) return 0;
  return cjIterations;
}

}  // namespace g2o

extern "C" int ref_pgo_solve(const orc_pgo_problem* p, int32_t iterations, double lambda_init, int32_t analytic_jac,
                             const volatile uint8_t* stop, orc_pgo_result* r) {
  (void)analytic_jac;
  using namespace g2o;
  PGO s;
  s.K = p->K; s.E = p->E; s.fix_scale = p->fix_scale != 0;
  s.fixed.assign(p->fixed, p->fixed + s.K);
  s.edges.resize(s.E);
  for (int e = 0; e < s.E; e++) { s.edges[e].i = p->edge_i[e]; s.edges[e].j = p->edge_j[e]; }
  std::vector<VertexSim3Expmap*> v;
  for (int k = 0; k < s.K; k++) {   // S/Optimizer.cpp:1095-1120
    VertexSim3Expmap* VSim3 = new VertexSim3Expmap();
    const double* q = p->sim3 + 8 * (size_t)k;
    VSim3->setEstimate(g2o::Sim3(Quaterniond(q[3], q[0], q[1], q[2]), Vector3d(q[4], q[5], q[6]), q[7]));
    VSim3->setFixed(p->fixed[k] != 0); VSim3->setId(k); VSim3->setMarginalized(false);
    VSim3->_fix_scale = s.fix_scale;
    v.push_back(VSim3);
  }
  std::vector<EdgeSim3*> edges;
  for (int e = 0; e < s.E; e++) {   // :1135-1150: identity information, no kernel
    EdgeSim3* ed = new EdgeSim3();
    ed->setVertex(1, v[p->edge_j[e]]); ed->setVertex(0, v[p->edge_i[e]]);
    const double* q = p->meas + 8 * (size_t)e;
    ed->setMeasurement(g2o::Sim3(Quaterniond(q[3], q[0], q[1], q[2]), Vector3d(q[4], q[5], q[6]), q[7]));
    ed->information() = Matrix<double, 7, 7>::Identity();
    edges.push_back(ed);
  }
  // initializeOptimization: edges with a free vertex, in id order; index mapping = free vertices with an active edge, in id order
  std::vector<char> has(s.K, 0);
  for (int e = 0; e < s.E; e++) {
    if (s.fixed[s.edges[e].i] && s.fixed[s.edges[e].j]) continue;
    s.active.push_back(e);
    has[s.edges[e].i] = has[s.edges[e].j] = 1;
  }
  s.vidx.assign(s.K, -1);
  for (int k = 0; k < s.K; k++)
    if (has[k] && !s.fixed[k]) { s.vidx[k] = (int)s.idxv.size(); s.idxv.push_back(k); }
  s.n = (int)s.idxv.size();
  r->trace_len = 0; r->iters_done = 0; r->chi2_initial = r->chi2_final = 0; r->lambda_final = 0;
  {
    SparseOptimizer optimizer(stop);
    for (int i = 0; i < s.n; i++) optimizer.iv_.push_back(v[s.idxv[i]]);
    for (int e : s.active) optimizer.active.push_back(edges[e]);
    OptimizationAlgorithmLevenberg* lm = new OptimizationAlgorithmLevenberg(new Solver(s, v));
    if (lambda_init > 0) lm->setUserLambdaInit(lambda_init);
    optimizer.setAlgorithm(lm);
    r->iters_done = optimizer.optimize(iterations, r);
    delete lm;
  }
  for (int k = 0; k < s.K; k++) {
    const g2o::Sim3& S = v[k]->estimate();
    double* q = r->sim3 + 8 * (size_t)k;
    q[0] = S.rotation().x(); q[1] = S.rotation().y(); q[2] = S.rotation().z(); q[3] = S.rotation().w();
    for (int i = 0; i < 3; i++) q[4 + i] = S.translation()[i];
    q[7] = S.scale();
  }
  for (size_t e = 0; e < edges.size(); e++) delete edges[e];
  for (size_t k = 0; k < v.size(); k++) delete v[k];
  return 0;
}

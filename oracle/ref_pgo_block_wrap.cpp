// ref_pgo_block_wrap.cpp — the essential-graph optimisation in which only the sparse factorisation is not the reference's
// (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// As ref_pgo_full_wrap.cpp, plus g2o's own BlockSolver_7_3 (block allocation with the transposed write for vertex pairs in descending
// order, buildSystem, setLambda / restoreDiagonal, the non-Schur solve() — G/core/block_solver.h(pp), solver.cpp,
// sparse_block_matrix*.h(pp) compiled where they lie over oracle/ref_stub/Eigen).  The oracle supplies LinearSolver::solve only (its
// sparse LDL^T in place of LinearSolverEigen).  ref_pgo_block_solve() has orc_pgo_solve()'s signature.
#include "pgo_oracle.cpp"

#include <iomanip>
#include <iostream>

#define G2O_GRAPH_OPTIMIZER_CHOL_H_
#include <core/batch_stats.h>
#include <core/hyper_graph.h>
#include <stuff/macros.h>
#include <types/types_seven_dof_expmap.h>

namespace g2o {

class OptimizationAlgorithm;

class SparseOptimizer : public OptimizableGraph {
 public:
  typedef OptimizableGraph::EdgeContainer EdgeContainer;
  const EdgeContainer& activeEdges() const { return active_base; }
  JacobianWorkspace& jacobianWorkspace() { return workspace; }
  bool verbose() const { return false; }
  EdgeContainer active_base;
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

}  // namespace g2o

#include <core/block_solver.h>   // the reference's own Solver, BlockSolver<Traits>, SparseBlockMatrix*, LinearSolver

namespace g2o {
class OracleLDLT7 : public LinearSolver<BlockSolver_7_3::PoseMatrixType> {
 public:
  typedef BlockSolver_7_3::PoseMatrixType M;
  OracleLDLT7() : analyzed_(false) {}
  virtual bool init() { analyzed_ = false; return true; }
  virtual bool solve(const SparseBlockMatrix<M>& A, double* x, double* b) {
    const int n = (int)A.blockCols().size();
    std::vector<std::vector<std::pair<int, const M*> > > rows(n);
    for (int j = 0; j < n; j++)
      for (SparseBlockMatrix<M>::IntBlockMap::const_iterator it = A.blockCols()[j].begin(); it != A.blockCols()[j].end(); ++it)
        if (it->first <= j) rows[it->first].push_back(std::make_pair(j, (const M*)it->second));
    S_.nb = n; S_.bs = 7; S_.rowptr.assign(n + 1, 0); S_.col.clear();
    for (int i = 0; i < n; i++) { for (size_t q = 0; q < rows[i].size(); q++) S_.col.push_back(rows[i][q].first); S_.rowptr[i + 1] = (int)S_.col.size(); }
    S_.val.assign(S_.col.size() * 49, 0.);
    size_t q = 0;
    for (int i = 0; i < n; i++)
      for (size_t k = 0; k < rows[i].size(); k++, q++)
        for (int r = 0; r < 7; r++) for (int c = 0; c < 7; c++) S_.val[q * 49 + r * 7 + c] = (*rows[i][k].second)(r, c);
    if (!analyzed_) { ldlt_.analyze(S_); analyzed_ = true; }
    if (!ldlt_.factorize(S_)) return false;
    ldlt_.solve(b, x);
    return true;
  }
 private:
  BlockSym S_;
  SparseLDLT ldlt_;
  bool analyzed_;
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
      tr[0] = i; tr[1] = std::numeric_limits<double>::quiet_NaN(); tr[2] = last_chi;
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

extern "C" int ref_pgo_block_solve(const orc_pgo_problem* p, int32_t iterations, double lambda_init, int32_t analytic_jac,
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
    for (int i = 0; i < s.n; i++) { optimizer.iv_.push_back(v[s.idxv[i]]); v[s.idxv[i]]->setHessianIndex(i); }   // buildIndexMapping
    for (int e : s.active) { optimizer.active.push_back(edges[e]); optimizer.active_base.push_back(edges[e]); }
    for (int e = 0; e < s.E; e++) { v[p->edge_i[e]]->edges().insert(edges[e]); v[p->edge_j[e]]->edges().insert(edges[e]); }      // addEdge
    OptimizationAlgorithmLevenberg* lm = new OptimizationAlgorithmLevenberg(new BlockSolver_7_3(new OracleLDLT7()));
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

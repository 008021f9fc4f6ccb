// ref_lm_wrap.cpp — reference pin for the Levenberg-Marquardt schedule (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// g2o as a whole cannot be built here (no Eigen).  Its damping schedule, trial loop and stop rules, however, live in three
// translation units that touch Eigen only through the headers they include for the types of their collaborators:
//   G/core/optimization_algorithm.cpp, G/core/optimization_algorithm_with_hessian.cpp, G/core/optimization_algorithm_levenberg.cpp.
// This file compiles those three sources WHERE THEY LIE (textual #include below, nothing copied) against stand-ins for the
// collaborators — g2o::SparseOptimizer, g2o::Solver, g2o::OptimizableGraph::Vertex — whose bodies are the oracle's own pieces
// (ba_oracle.cpp: errors, quadratic form, damped Schur solve, oplus, push/pop).  The result drives the oracle's linear algebra with
// the reference's OptimizationAlgorithmLevenberg::solve(), so comparing ref_lm_solve() with orc_ba_solve() on the same problem
// checks the oracle's restatement of SURVEY.md §8 row a8 (lambda initialisation, rho, scale factor, the nu doubling, qmax,
// the three-strike stop of the vendored copy) against the reference's compiled code, trace line by trace line.
// What stays restated: the ten-line outer loop of SparseOptimizer::optimize (G/core/sparse_optimizer.cpp:354-419), reproduced in
// SparseOptimizer::optimize below, and everything under the Solver interface (pinned by the scipy witnesses instead).
// g2o's property, string, time and statistics helpers are STL-only and are compiled in place as well (see the Makefile).
#include "ba_oracle.cpp"  // the oracle's pieces; its unnamed namespace is visible in this translation unit

#include <iomanip>
#include <iostream>
#include <map>
#include <set>
#include <utility>

// The headers of the collaborators pull in Eigen: mark them as already seen and declare what the three sources use of them.
#define G2O_SPARSE_BLOCK_MATRIX_
#define G2O_SOLVER_H
#define G2O_AIS_OPTIMIZABLE_GRAPH_HH_
#define G2O_GRAPH_OPTIMIZER_CHOL_H_
#include <core/batch_stats.h>
#include <core/hyper_graph.h>
#include <stuff/macros.h>

namespace g2o {

class MatrixXd;
template <class M> class SparseBlockMatrix;
class OptimizationAlgorithm;

class OptimizableGraph {
 public:
  class Vertex {
   public:
    Vertex(const BA* s, bool point, int index) : s_(s), point_(point), i_(index) {}
    bool marginalized() const { return point_; }  // S/Optimizer.cpp:746 vPoint->setMarginalized(true)
    int dimension() const { return point_ ? 3 : 6; }
    const double& hessian(int i, int j) const {
      return point_ ? s_->Hll[(size_t)i_ * 9 + i * 3 + j] : s_->Hpp[(size_t)i_ * 36 + i * 6 + j];
    }
   private:
    const BA* s_; bool point_; int i_;
  };
  typedef std::vector<Vertex*> VertexContainer;
};

class SparseOptimizer : public OptimizableGraph {
 public:
  SparseOptimizer(BA& s, const volatile uint8_t* stop) : s(s), stop_(stop), algorithm_(0), last_chi(0), chi_at_push(0), first_chi_(0), have_first_(false), trials(0) {
    for (int i = 0; i < s.np; i++) iv_.push_back(new Vertex(&s, false, i));  // buildIndexMapping: free poses, then points
    for (int l = 0; l < s.nl; l++) iv_.push_back(new Vertex(&s, true, l));
  }
  ~SparseOptimizer() { for (size_t i = 0; i < iv_.size(); i++) delete iv_[i]; }
  const VertexContainer& indexMapping() const { return iv_; }
  const VertexContainer& activeVertices() const { return iv_; }
  void computeActiveErrors() { compute_active_errors(s); }
  double activeRobustChi2() {
    last_chi = active_robust_chi2(s);
    if (!have_first_) { first_chi_ = last_chi; have_first_ = true; }
    return last_chi;
  }
  void push() { ::push(s); chi_at_push = last_chi; trials++; }
  void pop() { ::pop(s); last_chi = chi_at_push; }
  void discardTop() {}
  void update(const double* x) { (void)x; ::update(s); }  // x is the solver's vector, which is s.x
  bool terminate() { return stop_ && *stop_; }            // G/core/sparse_optimizer.h:188
  bool verbose() const { return false; }
  void setAlgorithm(OptimizationAlgorithm* a);
  int optimize(int iterations, orc_ba_result* r);
  BA& s;
  const volatile uint8_t* stop_;
  OptimizationAlgorithm* algorithm_;
  VertexContainer iv_;
  double last_chi, chi_at_push, first_chi_;
  bool have_first_;
  int trials;
};

class Solver {
 public:
  explicit Solver(BA& s) : s(s), opt_(0), lambda_(0), schur_(false) {}
  virtual ~Solver() {}
  bool init(SparseOptimizer* o, bool) { opt_ = o; return true; }
  SparseOptimizer* optimizer() const { return opt_; }
  bool buildStructure(bool = false) { build_structure(s); b_.assign((size_t)s.np * 6 + (size_t)s.nl * 3, 0.); return true; }
  bool updateStructure(const std::vector<HyperGraph::Vertex*>&, const HyperGraph::EdgeSet&) { return false; }
  bool buildSystem() {
    build_system(s);
    std::copy(s.bp.begin(), s.bp.end(), b_.begin());
    std::copy(s.bl.begin(), s.bl.end(), b_.begin() + (size_t)s.np * 6);
    return true;
  }
  bool setLambda(double lambda, bool = false) { lambda_ = lambda; return true; }  // damping is applied to a copy inside solve_system
  void restoreDiagonal() {}
  bool solve() { return solve_system(s, lambda_); }
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
  std::vector<double> b_;
};

}  // namespace g2o

// the reference's sources, compiled in place
#include <core/optimization_algorithm.cpp>
#include <core/optimization_algorithm_with_hessian.cpp>
#include <core/optimization_algorithm_levenberg.cpp>

namespace g2o {

void SparseOptimizer::setAlgorithm(OptimizationAlgorithm* a) { algorithm_ = a; a->setOptimizer(this); }

// The outer loop, G/core/sparse_optimizer.cpp:354-419 without statistics and verbose output.
int SparseOptimizer::optimize(int iterations, orc_ba_result* r) {
  if (iv_.size() == 0 || s.active.empty()) return -1;
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
      tr[3] = std::numeric_limits<double>::quiet_NaN();  // rho is a local of solve()
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

extern "C" int ref_lm_solve(const orc_ba_problem* p, const orc_ba_options* o, orc_ba_result* r) {
  BA s;
  load(s, p, o->robust, o->huber_delta);
  init_active(s);
  r->trace_len = 0; r->iters_done = 0; r->trials_total = 0;
  r->chi2_initial = r->chi2_final = 0; r->lambda_final = 0;
  {
    g2o::SparseOptimizer optimizer(s, o->stop);
    g2o::OptimizationAlgorithmLevenberg* lm = new g2o::OptimizationAlgorithmLevenberg(new g2o::Solver(s));  // owns the solver
    if (o->lambda_init > 0) lm->setUserLambdaInit(o->lambda_init);
    if (o->max_trials > 0) lm->setMaxTrialsAfterFailure(o->max_trials);
    optimizer.setAlgorithm(lm);
    r->iters_done = optimizer.optimize(o->iterations, r);
    delete lm;
  }
  for (int k = 0; k < s.K; k++) se3_store(s.pose[k], r->poses + 7 * (size_t)k);
  std::memcpy(r->points, s.pt.data(), sizeof(double) * 3 * (size_t)s.P);
  if (r->chi2)
    for (int e : s.active) r->chi2[e] = edge_chi2(s.edges[e]);
  if (r->depth_pos)
    for (int e = 0; e < s.E; e++) {
      double xc[3];
      se3_map(s.pose[s.edges[e].kf], &s.pt[3 * (size_t)s.edges[e].mp], xc);
      r->depth_pos[e] = xc[2] > 0.0;
    }
  return 0;
}

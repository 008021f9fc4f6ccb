// ref_single_full_wrap.cpp — the two single-vertex optimisations with the reference's code around the 6x6 / 7x7 solve
// (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Same construction as ref_ba_full_wrap.cpp: g2o's own Levenberg-Marquardt driver (optimization_algorithm*.cpp, included textually)
// over g2o's own VertexSE3Expmap / EdgeSE3ProjectXYZOnlyPose (Optimizer::PoseOptimizationClient) and VertexSim3Expmap /
// EdgeSim3ProjectXYZ / EdgeInverseSim3ProjectXYZ with fixed VertexSBAPointXYZ (Optimizer::OptimizeSim3), Huber kernels included, all
// compiled from the reference tree against oracle/ref_stub/Eigen.  The oracle supplies the dense Cholesky under Solver::solve()
// (single_oracle.cpp chol_solve, standing in for LinearSolverDense, G/solvers/linear_solver_dense.h:64-113).  The glue restates
// SparseOptimizer::{initializeOptimization(level), optimize, computeActiveErrors, activeRobustChi2, update, push, pop} and
// BlockSolver::buildSystem as in ref_ba_full_wrap.cpp, plus the two short protocols around optimize() — S/Optimizer.cpp:290-347
// (four rounds, chi2 classification as float, levels, kernels dropped in round 3) and :995-1055 (two passes, edge removal).
// ref_pose_optimize / ref_sim3_optimize have the signatures of orc_pose_optimize / orc_sim3_optimize.
#include "single_oracle.cpp"  // chol_solve<N>; the unnamed namespace is visible here

#include <iomanip>
#include <iostream>

#define G2O_SPARSE_BLOCK_MATRIX_
#define G2O_SOLVER_H
#define G2O_GRAPH_OPTIMIZER_CHOL_H_
#include <core/batch_stats.h>
#include <core/hyper_graph.h>
#include <core/robust_kernel_impl.h>
#include <stuff/macros.h>
#include <types/types_seven_dof_expmap.h>

namespace g2o {

class MatrixXd;
template <class M> class SparseBlockMatrix;
class OptimizationAlgorithm;

class SparseOptimizer : public OptimizableGraph {
 public:
  SparseOptimizer() : algorithm_(0) {}
  const VertexContainer& indexMapping() const { return iv_; }
  const VertexContainer& activeVertices() const { return iv_; }
  void initializeOptimization(int level = 0) {   // the edges of that level; the one free vertex is the index mapping
    active.clear();
    for (size_t k = 0; k < edges.size(); k++) if (edges[k] && edges[k]->level() == level) active.push_back(edges[k]);
  }
  void computeActiveErrors() { for (size_t k = 0; k < active.size(); k++) active[k]->computeError(); }
  double activeRobustChi2() {
    Eigen::Vector3d rho;
    double chi = 0.0;
    for (size_t k = 0; k < active.size(); k++) {
      const Edge* e = active[k];
      if (e->robustKernel()) { e->robustKernel()->robustify(e->chi2(), rho); chi += rho[0]; }
      else chi += e->chi2();
    }
    return chi;
  }
  void push() { for (size_t i = 0; i < iv_.size(); i++) iv_[i]->push(); }
  void pop() { for (size_t i = 0; i < iv_.size(); i++) iv_[i]->pop(); }
  void discardTop() { for (size_t i = 0; i < iv_.size(); i++) iv_[i]->discardTop(); }
  void update(const double* update) { for (size_t i = 0; i < iv_.size(); ++i) { iv_[i]->oplus(update); update += iv_[i]->dimension(); } }
  bool terminate() { return false; }
  void setAlgorithm(OptimizationAlgorithm* a);
  int optimize(int iterations);
  OptimizationAlgorithm* algorithm_;
  VertexContainer iv_;
  std::vector<Edge*> edges, active;   // removed edges are null
  JacobianWorkspace workspace;
};

class Solver {
 public:
  Solver() : opt_(0), lambda_(0) {}
  virtual ~Solver() {}
  bool init(SparseOptimizer* o, bool) { opt_ = o; return true; }
  SparseOptimizer* optimizer() const { return opt_; }
  virtual bool buildStructure(bool = false) = 0;
  bool updateStructure(const std::vector<HyperGraph::Vertex*>&, const HyperGraph::EdgeSet&) { return false; }
  virtual bool buildSystem() = 0;
  bool setLambda(double lambda, bool = false) { lambda_ = lambda; return true; }
  void restoreDiagonal() {}
  virtual bool solve() = 0;
  bool computeMarginals(SparseBlockMatrix<MatrixXd>&, const std::vector<std::pair<int, int> >&) { return false; }
  virtual double* x() = 0;
  virtual double* b() = 0;
  virtual size_t vectorSize() const = 0;
  bool schur() { return false; }
  bool supportsSchur() { return false; }
  void setSchur(bool) {}
  void setWriteDebug(bool) {}
  SparseOptimizer* opt_;
  double lambda_;
};

template <int D> class DenseSolver : public Solver {   // BlockSolverX + LinearSolverDense on one D-dimensional vertex
 public:
  bool buildStructure(bool = false) { opt_->iv_[0]->mapHessianMemory(h_); return true; }
  bool buildSystem() {
    opt_->iv_[0]->clearQuadraticForm();
    std::fill(h_, h_ + D * D, 0.);
    for (size_t k = 0; k < opt_->active.size(); ++k) {
      opt_->active[k]->linearizeOplus(opt_->workspace);
      opt_->active[k]->constructQuadraticForm();
    }
    opt_->iv_[0]->copyB(b_);
    return true;
  }
  bool solve() {
    double Hd[D * D];
    for (int r = 0; r < D; r++) for (int c = 0; c < D; c++) Hd[r * D + c] = h_[c * D + r];
    for (int i = 0; i < D; i++) Hd[i * D + i] += lambda_;
    const bool ok = chol_solve<D>(Hd, b_, x_);
    if (!ok) std::fill(x_, x_ + D, 0.0);
    return ok;
  }
  double* x() { return x_; }
  double* b() { return b_; }
  size_t vectorSize() const { return D; }
  double h_[D * D], b_[D], x_[D];
};

}  // namespace g2o

#include <core/optimization_algorithm.cpp>
#include <core/optimization_algorithm_with_hessian.cpp>
#include <core/optimization_algorithm_levenberg.cpp>

namespace g2o {

void SparseOptimizer::setAlgorithm(OptimizationAlgorithm* a) { algorithm_ = a; a->setOptimizer(this); }

int SparseOptimizer::optimize(int iterations) {  // G/core/sparse_optimizer.cpp:354-419
  if (iv_.size() == 0 || active.empty()) return -1;
  int cjIterations = 0;
  bool ok = algorithm_->init(false);
  if (!ok) return -1;
  OptimizationAlgorithm::SolverResult result = OptimizationAlgorithm::OK;
  for (int i = 0; i < iterations && !terminate() && ok; i++) {
    result = algorithm_->solve(i, false);
    ok = (result == OptimizationAlgorithm::OK);
    ++cjIterations;
  }
  if (result == OptimizationAlgorithm::Fail) return 0;
  return cjIterations;
}

}  // namespace g2o

using namespace g2o;

extern "C" int ref_pose_optimize(const orc_pose_opt_problem* p, double* Tcw_out, uint8_t* outlier) {
  const int N = p->n;
  if (N < 3) { std::memcpy(Tcw_out, p->Tcw, 7 * sizeof(double)); return 0; }
  SparseOptimizer optimizer;
  OptimizationAlgorithmLevenberg* solver = new OptimizationAlgorithmLevenberg(new DenseSolver<6>());
  optimizer.setAlgorithm(solver);
  Vector7d v7; v7[0] = p->Tcw[4]; v7[1] = p->Tcw[5]; v7[2] = p->Tcw[6]; v7[3] = p->Tcw[0]; v7[4] = p->Tcw[1]; v7[5] = p->Tcw[2]; v7[6] = p->Tcw[3];
  SE3Quat T0; T0.fromVector(v7);
  VertexSE3Expmap* vSE3 = new VertexSE3Expmap();
  vSE3->setEstimate(T0); vSE3->setId(0); vSE3->setFixed(false);
  optimizer.iv_.push_back(vSE3);
  const float deltaMono = sqrt(5.991);
  std::vector<EdgeSE3ProjectXYZOnlyPose*> vpEdgesMono;
  std::vector<RobustKernelHuber*> kernels;
  for (int i = 0; i < N; i++) {
    outlier[i] = 0;
    EdgeSE3ProjectXYZOnlyPose* e = new EdgeSE3ProjectXYZOnlyPose();
    e->setVertex(0, vSE3);
    e->setMeasurement(Vector2d(p->uv[2 * i], p->uv[2 * i + 1]));
    const float invSigma2 = p->inv_sigma2[i];
    e->setInformation(Matrix2d::Identity() * invSigma2);
    RobustKernelHuber* rk = new RobustKernelHuber;
    e->setRobustKernel(rk); rk->setDelta(deltaMono);
    e->fx = p->fx; e->fy = p->fy; e->cx = p->cx; e->cy = p->cy;
    e->Xw[0] = p->Xw[3 * i]; e->Xw[1] = p->Xw[3 * i + 1]; e->Xw[2] = p->Xw[3 * i + 2];
    optimizer.edges.push_back(e); vpEdgesMono.push_back(e); kernels.push_back(rk);
  }
  const float chi2Mono[4] = {5.991, 5.991, 5.991, 5.991};
  const int its[4] = {10, 10, 10, 10};
  int nBad = 0;
  for (size_t it = 0; it < 4; it++) {
    vSE3->setEstimate(T0);
    optimizer.initializeOptimization(0);
    optimizer.optimize(its[it]);
    nBad = 0;
    for (size_t i = 0; i < vpEdgesMono.size(); i++) {
      EdgeSE3ProjectXYZOnlyPose* e = vpEdgesMono[i];
      if (outlier[i]) e->computeError();
      const float chi2 = e->chi2();
      if (chi2 > chi2Mono[it]) { outlier[i] = 1; e->setLevel(1); nBad++; }
      else { outlier[i] = 0; e->setLevel(0); }
      if (it == 2) e->setRobustKernel(0);
    }
    if (optimizer.edges.size() < 10) break;
  }
  const SE3Quat T = vSE3->estimate();
  Tcw_out[0] = T.rotation().x(); Tcw_out[1] = T.rotation().y(); Tcw_out[2] = T.rotation().z(); Tcw_out[3] = T.rotation().w();
  for (int i = 0; i < 3; i++) Tcw_out[4 + i] = T.translation()[i];
  delete solver;
  for (int i = 0; i < N; i++) { delete vpEdgesMono[i]; delete kernels[i]; }
  delete vSE3;
  return N - nBad;
}

extern "C" int ref_sim3_optimize(const orc_sim3_opt_problem* p, double* S12_out, uint8_t* inlier) {
  const int N = p->n;
  SparseOptimizer optimizer;
  OptimizationAlgorithmLevenberg* solver = new OptimizationAlgorithmLevenberg(new DenseSolver<7>());
  optimizer.setAlgorithm(solver);
  VertexSim3Expmap* vSim3 = new VertexSim3Expmap();
  vSim3->_fix_scale = p->fix_scale != 0;
  vSim3->setEstimate(g2o::Sim3(Quaterniond(p->S12[3], p->S12[0], p->S12[1], p->S12[2]), Vector3d(p->S12[4], p->S12[5], p->S12[6]), p->S12[7]));
  vSim3->setId(0); vSim3->setFixed(false);
  vSim3->_principle_point1[0] = p->K1[2]; vSim3->_principle_point1[1] = p->K1[3];
  vSim3->_focal_length1[0] = p->K1[0]; vSim3->_focal_length1[1] = p->K1[1];
  vSim3->_principle_point2[0] = p->K2[2]; vSim3->_principle_point2[1] = p->K2[3];
  vSim3->_focal_length2[0] = p->K2[0]; vSim3->_focal_length2[1] = p->K2[1];
  optimizer.iv_.push_back(vSim3);
  std::memcpy(S12_out, p->S12, 8 * sizeof(double));
  const float th2 = p->th2;
  const float deltaHuber = sqrt(th2);
  std::vector<EdgeSim3ProjectXYZ*> vpEdges12;
  std::vector<EdgeInverseSim3ProjectXYZ*> vpEdges21;
  std::vector<VertexSBAPointXYZ*> points;
  std::vector<RobustKernelHuber*> kernels;
  for (int i = 0; i < N; i++) {
    inlier[i] = 1;
    VertexSBAPointXYZ* vPoint1 = new VertexSBAPointXYZ();
    vPoint1->setEstimate(Vector3d(p->P1c[3 * i], p->P1c[3 * i + 1], p->P1c[3 * i + 2])); vPoint1->setId(2 * i + 1); vPoint1->setFixed(true);
    VertexSBAPointXYZ* vPoint2 = new VertexSBAPointXYZ();
    vPoint2->setEstimate(Vector3d(p->P2c[3 * i], p->P2c[3 * i + 1], p->P2c[3 * i + 2])); vPoint2->setId(2 * i + 2); vPoint2->setFixed(true);
    points.push_back(vPoint1); points.push_back(vPoint2);
    EdgeSim3ProjectXYZ* e12 = new EdgeSim3ProjectXYZ();        // x1 = S12 * X2
    e12->setVertex(0, vPoint2); e12->setVertex(1, vSim3);
    e12->setMeasurement(Vector2d(p->uv1[2 * i], p->uv1[2 * i + 1]));
    const float invSigmaSquare1 = p->inv_sigma2_1[i];
    e12->setInformation(Matrix2d::Identity() * invSigmaSquare1);
    RobustKernelHuber* rk1 = new RobustKernelHuber; e12->setRobustKernel(rk1); rk1->setDelta(deltaHuber);
    EdgeInverseSim3ProjectXYZ* e21 = new EdgeInverseSim3ProjectXYZ();  // x2 = S21 * X1
    e21->setVertex(0, vPoint1); e21->setVertex(1, vSim3);
    e21->setMeasurement(Vector2d(p->uv2[2 * i], p->uv2[2 * i + 1]));
    const float invSigmaSquare2 = p->inv_sigma2_2[i];
    e21->setInformation(Matrix2d::Identity() * invSigmaSquare2);
    RobustKernelHuber* rk2 = new RobustKernelHuber; e21->setRobustKernel(rk2); rk2->setDelta(deltaHuber);
    optimizer.edges.push_back(e12); optimizer.edges.push_back(e21);
    vpEdges12.push_back(e12); vpEdges21.push_back(e21); kernels.push_back(rk1); kernels.push_back(rk2);
  }
  int result = 0;
  bool done = false;
  optimizer.initializeOptimization();
  optimizer.optimize(5);
  int nBad = 0;
  for (size_t i = 0; i < vpEdges12.size(); i++) {
    if (vpEdges12[i]->chi2() > th2 || vpEdges21[i]->chi2() > th2) {
      inlier[i] = 0;
      optimizer.edges[2 * i] = 0; optimizer.edges[2 * i + 1] = 0;   // removeEdge
      nBad++;
    }
  }
  const int nMoreIterations = nBad > 0 ? 10 : 5;
  if (N - nBad < 10) { result = 0; done = true; }
  if (!done) {
    optimizer.initializeOptimization();
    optimizer.optimize(nMoreIterations);
    int nIn = 0;
    for (size_t i = 0; i < vpEdges12.size(); i++) {
      if (!inlier[i]) continue;
      if (vpEdges12[i]->chi2() > th2 || vpEdges21[i]->chi2() > th2) inlier[i] = 0; else nIn++;
    }
    const g2o::Sim3 S = vSim3->estimate();
    S12_out[0] = S.rotation().x(); S12_out[1] = S.rotation().y(); S12_out[2] = S.rotation().z(); S12_out[3] = S.rotation().w();
    for (int i = 0; i < 3; i++) S12_out[4 + i] = S.translation()[i];
    S12_out[7] = S.scale();
    result = nIn;
  }
  delete solver;
  for (size_t i = 0; i < vpEdges12.size(); i++) { delete vpEdges12[i]; delete vpEdges21[i]; }
  for (size_t i = 0; i < kernels.size(); i++) delete kernels[i];
  for (size_t i = 0; i < points.size(); i++) delete points[i];
  delete vSim3;
  return result;
}

// ba_oracle.cpp — CPU oracle for the bundle-adjustment hot path (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Single-threaded restatement of what cslam::Optimizer::{MapFusionGBA, LocalBundleAdjustmentClient,
// BundleAdjustmentClient} make vendored g2o do between initializeOptimization() and the end of
// optimize(n) (S/Optimizer.cpp:788-797, 535-567, 164-166), on a flat problem.
//   LM loop ........ G/core/optimization_algorithm_levenberg.cpp:61-189
//   outer loop ..... G/core/sparse_optimizer.cpp:354-419 (terminate(): G/core/sparse_optimizer.h:188)
//   active set ..... G/core/sparse_optimizer.cpp:166-190,206-267
//   buildSystem .... G/core/block_solver.hpp:502-560, G/core/base_binary_edge.hpp:55-120
//   Schur + solve .. G/core/block_solver.hpp:354-486  (+ direct LDL^T, G/solvers/linear_solver_eigen.h:106-133)
//   edge ........... G/types/types_six_dof_expmap.{h:80-109,cpp:103-147}
//   Huber .......... G/core/robust_kernel_impl.cpp:77-91, G/core/base_edge.h:96-102
// Parity pinning: no reference tests exist for this path and g2o as a whole cannot be built here (no Eigen);
// pinned by scipy / finite-difference witnesses in tests/test_oracle_ba.py, and the LM control flow of orc_ba_solve by the
// reference's own optimization_algorithm_levenberg.cpp compiled over these pieces (ref_lm_wrap.cpp, tests/test_oracle_vs_reference_lm.py),
// edge errors / Jacobians / Huber / constructQuadraticForm by the reference's own types compiled over a stand-in Eigen
// (ref_g2o_wrap.cpp, tests/test_oracle_vs_reference_g2o.py), structure / Schur complement / back-substitution by the reference's own
// BlockSolver_6_3 (ref_ba_block_wrap.cpp) — bit for bit.  Only the sparse LDL^T itself has no reference counterpart here.
#include "oracle.h"

#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "lie.hpp"
#include "sparse_ldlt.hpp"

namespace {
using namespace orc;

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Edge {
  int kf, mp;
  double u, v, w;  // measurement (f32 widened, S/Optimizer.cpp:758-765), information = w * I2
  bool active, robust;
  double err[2];   // _error, only refreshed by computeActiveErrors
};

struct Lin {  // linearizeOplus output, G/types/types_six_dof_expmap.cpp:103-139
  double Jl[6];   // 2x3 d e / d point
  double Jp[12];  // 2x6 d e / d pose (omega, upsilon)
};

struct BA {
  int K, P, E;
  std::vector<SE3> pose, pose_bak;
  std::vector<double> pt, pt_bak;  // 3P
  std::vector<double> intr;        // 4K
  std::vector<uint8_t> fixed;
  std::vector<Edge> edges;
  std::vector<int> active;         // active edge ids, sorted by id (sortVectorContainers)
  // index mapping (buildIndexMapping): free active poses, then active points
  std::vector<int> pose_idx, pt_idx;  // -1 if not in mapping
  std::vector<int> idx_pose, idx_pt;  // inverse
  int np = 0, nl = 0;
  double huber_delta = 0;
  // system
  std::vector<double> Hpp, bp;     // np*36, np*6
  std::vector<double> Hll, bl;     // nl*9, nl*3
  std::vector<double> Hpl;         // per active edge 18 (6x3), zero if pose fixed
  std::vector<int> lm_ptr, lm_edges;  // per landmark index: active edges with a free pose, sorted by pose idx
  BlockSym S;
  SparseLDLT ldlt;
  std::vector<double> x, bschur, coeff, Dinv;
  double t_build = 0, t_schur = 0, t_solve = 0, t_resid = 0, t_structure = 0;
};

inline void edge_error(const BA& s, Edge& e) {  // computeError, G/types/types_six_dof_expmap.h:90-95
  double xc[3];
  se3_map(s.pose[e.kf], &s.pt[3 * e.mp], xc);
  const double* in = &s.intr[4 * e.kf];
  double px = xc[0] / xc[2], py = xc[1] / xc[2];  // project2d
  e.err[0] = e.u - (px * in[0] + in[2]);
  e.err[1] = e.v - (py * in[1] + in[3]);
}
inline double edge_chi2(const Edge& e) {  // BaseEdge::chi2, G/core/base_edge.h:58-61
  return e.err[0] * (e.w * e.err[0]) + e.err[1] * (e.w * e.err[1]);
}
inline void huber(double e, double delta, double rho[3]) {  // G/core/robust_kernel_impl.cpp:64-91
  // the vendored RobustKernelHuber keeps delta^2 in a FLOAT member (G/core/robust_kernel_impl.h:84, set in setDelta): both the
  // inlier test and rho(e) use the rounded value.  Found by compiling the reference's kernel (oracle/ref_g2o_wrap.cpp).
  double dsqr = (double)(float)(delta * delta);
  if (e <= dsqr) {
    rho[0] = e; rho[1] = 1.; rho[2] = 0.;
  } else {
    double sqrte = std::sqrt(e);
    rho[0] = 2 * sqrte * delta - dsqr;
    rho[1] = delta / sqrte;
    rho[2] = -0.5 * rho[1] / e;
  }
}

inline void linearize(const BA& s, const Edge& e, Lin& L) {
  const SE3& T = s.pose[e.kf];
  double xc[3];
  se3_map(T, &s.pt[3 * e.mp], xc);
  const double fx = s.intr[4 * e.kf], fy = s.intr[4 * e.kf + 1];
  double x = xc[0], y = xc[1], z = xc[2], z_2 = z * z;
  double tmp[6] = {fx, 0, -x / z * fx, 0, fy, -y / z * fy};
  double R[9];
  q2R(T.r, R);
  // `-1./z * tmp * R` groups as ((-1/z) * tmp) * R  (types_six_dof_expmap.cpp:127; pinned by oracle/ref_g2o_wrap.cpp)
  double st[6];
  for (int i = 0; i < 6; i++) st[i] = -1. / z * tmp[i];
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++) L.Jl[i * 3 + j] = st[i * 3] * R[j] + st[i * 3 + 1] * R[3 + j] + st[i * 3 + 2] * R[6 + j];
  L.Jp[0] = x * y / z_2 * fx;
  L.Jp[1] = -(1 + (x * x / z_2)) * fx;
  L.Jp[2] = y / z * fx;
  L.Jp[3] = -1. / z * fx;
  L.Jp[4] = 0;
  L.Jp[5] = x / z_2 * fx;
  L.Jp[6] = (1 + y * y / z_2) * fy;
  L.Jp[7] = -x * y / z_2 * fy;
  L.Jp[8] = -x / z * fy;
  L.Jp[9] = 0;
  L.Jp[10] = -1. / z * fy;
  L.Jp[11] = y / z_2 * fy;
}

void load(BA& s, const orc_ba_problem* p, int robust, double huber_delta) {
  s.K = p->K; s.P = p->P; s.E = p->E;
  s.pose.resize(s.K);
  for (int k = 0; k < s.K; k++) s.pose[k] = se3_load(p->poses + 7 * k);
  s.pt.assign(p->points, p->points + 3 * (size_t)s.P);
  s.intr.assign(p->intr, p->intr + 4 * (size_t)s.K);
  s.fixed.assign(p->fixed, p->fixed + s.K);
  s.edges.resize(s.E);
  for (int e = 0; e < s.E; e++) {
    Edge& ed = s.edges[e];
    ed.kf = p->obs_kf[e]; ed.mp = p->obs_mp[e];
    ed.u = p->obs_uv[2 * e]; ed.v = p->obs_uv[2 * e + 1]; ed.w = p->obs_w[e];
    uint8_t fl = p->edge_flags ? p->edge_flags[e] : 0;
    ed.active = !(fl & 1);
    ed.robust = robust && !(fl & 2);
    ed.err[0] = ed.err[1] = 0;
  }
  s.huber_delta = huber_delta;
}

// initializeOptimization(level 0) + buildIndexMapping
void init_active(BA& s) {
  s.active.clear();
  std::vector<char> pose_has(s.K, 0), pt_has(s.P, 0);
  for (int e = 0; e < s.E; e++)
    if (s.edges[e].active) {  // points are never fixed -> allVerticesFixed() is false
      s.active.push_back(e);
      pose_has[s.edges[e].kf] = 1;
      pt_has[s.edges[e].mp] = 1;
    }
  s.pose_idx.assign(s.K, -1); s.pt_idx.assign(s.P, -1);
  s.idx_pose.clear(); s.idx_pt.clear();
  for (int k = 0; k < s.K; k++)
    if (pose_has[k] && !s.fixed[k]) { s.pose_idx[k] = (int)s.idx_pose.size(); s.idx_pose.push_back(k); }
  for (int j = 0; j < s.P; j++)
    if (pt_has[j]) { s.pt_idx[j] = (int)s.idx_pt.size(); s.idx_pt.push_back(j); }
  s.np = (int)s.idx_pose.size(); s.nl = (int)s.idx_pt.size();
}

// BlockSolver::buildStructure (G/core/block_solver.hpp:143-295): Hpl column lists + Schur pattern
void build_structure(BA& s) {
  double t0 = now_s();
  s.Hpp.assign((size_t)s.np * 36, 0); s.bp.assign((size_t)s.np * 6, 0);
  s.Hll.assign((size_t)s.nl * 9, 0); s.bl.assign((size_t)s.nl * 3, 0);
  s.Hpl.assign(s.active.size() * 18, 0);
  // landmark columns of Hpl, sorted by pose row
  std::vector<int> cnt(s.nl + 1, 0);
  for (size_t a = 0; a < s.active.size(); a++) {
    const Edge& e = s.edges[s.active[a]];
    if (s.pose_idx[e.kf] >= 0) cnt[s.pt_idx[e.mp] + 1]++;
  }
  s.lm_ptr.assign(s.nl + 1, 0);
  for (int l = 0; l < s.nl; l++) s.lm_ptr[l + 1] = s.lm_ptr[l] + cnt[l + 1];
  s.lm_edges.resize(s.lm_ptr[s.nl]);
  std::vector<int> fill(s.lm_ptr.begin(), s.lm_ptr.end() - 1);
  for (size_t a = 0; a < s.active.size(); a++) {
    const Edge& e = s.edges[s.active[a]];
    if (s.pose_idx[e.kf] >= 0) s.lm_edges[fill[s.pt_idx[e.mp]]++] = (int)a;
  }
  for (int l = 0; l < s.nl; l++)
    std::sort(s.lm_edges.begin() + s.lm_ptr[l], s.lm_edges.begin() + s.lm_ptr[l + 1], [&](int a, int b) {
      return s.pose_idx[s.edges[s.active[a]].kf] < s.pose_idx[s.edges[s.active[b]].kf];
    });
  // Schur pattern: diagonal + every (i1<=i2) pair co-observing a landmark
  std::vector<std::vector<int>> rows(s.np);
  for (int i = 0; i < s.np; i++) rows[i].push_back(i);
  for (int l = 0; l < s.nl; l++)
    for (int a = s.lm_ptr[l]; a < s.lm_ptr[l + 1]; a++) {
      int i1 = s.pose_idx[s.edges[s.active[s.lm_edges[a]]].kf];
      for (int b = a + 1; b < s.lm_ptr[l + 1]; b++) rows[i1].push_back(s.pose_idx[s.edges[s.active[s.lm_edges[b]]].kf]);
      if (rows[i1].size() > 4096) { std::sort(rows[i1].begin(), rows[i1].end()); rows[i1].erase(std::unique(rows[i1].begin(), rows[i1].end()), rows[i1].end()); }
    }
  s.S.nb = s.np; s.S.bs = 6;
  s.S.rowptr.assign(s.np + 1, 0); s.S.col.clear();
  for (int i = 0; i < s.np; i++) {
    std::sort(rows[i].begin(), rows[i].end());
    rows[i].erase(std::unique(rows[i].begin(), rows[i].end()), rows[i].end());
    s.S.col.insert(s.S.col.end(), rows[i].begin(), rows[i].end());
    s.S.rowptr[i + 1] = (int)s.S.col.size();
  }
  s.S.val.assign(s.S.col.size() * 36, 0);
  s.ldlt.analyze(s.S);  // LinearSolverEigen::computeSymbolicDecomposition, once per optimize()
  s.x.assign((size_t)s.np * 6 + (size_t)s.nl * 3, 0);
  s.bschur.assign((size_t)s.np * 6, 0); s.coeff.assign((size_t)s.np * 6, 0);
  s.Dinv.assign((size_t)s.nl * 9, 0);
  s.t_structure += now_s() - t0;
}

void compute_active_errors(BA& s) {  // G/core/sparse_optimizer.cpp:61-88
  double t0 = now_s();
  for (int e : s.active) edge_error(s, s.edges[e]);
  s.t_resid += now_s() - t0;
}
double active_robust_chi2(const BA& s) {  // G/core/sparse_optimizer.cpp:100-114
  double chi = 0, rho[3];
  for (int id : s.active) {
    const Edge& e = s.edges[id];
    if (e.robust) { huber(edge_chi2(e), s.huber_delta, rho); chi += rho[0]; }
    else chi += edge_chi2(e);
  }
  return chi;
}

// buildSystem: linearizeOplus + constructQuadraticForm per active edge
void build_system(BA& s) {
  double t0 = now_s();
  std::fill(s.Hpp.begin(), s.Hpp.end(), 0.); std::fill(s.bp.begin(), s.bp.end(), 0.);
  std::fill(s.Hll.begin(), s.Hll.end(), 0.); std::fill(s.bl.begin(), s.bl.end(), 0.);
  std::fill(s.Hpl.begin(), s.Hpl.end(), 0.);
  Lin L;
  for (size_t a = 0; a < s.active.size(); a++) {
    const Edge& e = s.edges[s.active[a]];
    linearize(s, e, L);
    double w = e.w, rho1 = 1.0;
    if (e.robust) { double rho[3]; huber(edge_chi2(e), s.huber_delta, rho); rho1 = rho[1]; }
    // omega_r = -omega*e (*rho1); weightedOmega = rho1*omega  (G/core/base_binary_edge.hpp:75-113)
    double or0 = -(w * e.err[0]) * rho1, or1 = -(w * e.err[1]) * rho1;
    double wo = rho1 * w;
    const int li = s.pt_idx[e.mp], pi = s.pose_idx[e.kf];
    double* Hl = &s.Hll[(size_t)li * 9];
    double* b_l = &s.bl[(size_t)li * 3];
    for (int i = 0; i < 3; i++) {
      b_l[i] += L.Jl[i] * or0 + L.Jl[3 + i] * or1;
      for (int j = 0; j < 3; j++) Hl[i * 3 + j] += L.Jl[i] * wo * L.Jl[j] + L.Jl[3 + i] * wo * L.Jl[3 + j];
    }
    if (pi >= 0) {
      double* Hp = &s.Hpp[(size_t)pi * 36];
      double* b_p = &s.bp[(size_t)pi * 6];
      double* W = &s.Hpl[a * 18];
      for (int i = 0; i < 6; i++) {
        b_p[i] += L.Jp[i] * or0 + L.Jp[6 + i] * or1;
        for (int j = 0; j < 6; j++) Hp[i * 6 + j] += L.Jp[i] * wo * L.Jp[j] + L.Jp[6 + i] * wo * L.Jp[6 + j];
        // the transposed pose-landmark block (block_solver.hpp:244): with a kernel  (B^T wOmega) A, without one  B^T (A^T Omega)^T —
        // the weight multiplies the other factor first (base_binary_edge.hpp:78-81 vs :98-101)
        if (e.robust) for (int j = 0; j < 3; j++) W[i * 3 + j] += L.Jp[i] * wo * L.Jl[j] + L.Jp[6 + i] * wo * L.Jl[3 + j];
        else for (int j = 0; j < 3; j++) W[i * 3 + j] += L.Jp[i] * (L.Jl[j] * wo) + L.Jp[6 + i] * (L.Jl[3 + j] * wo);
      }
    }
  }
  s.t_build += now_s() - t0;
}

// BlockSolver::solve with Schur (G/core/block_solver.hpp:354-486); lambda already NOT in H -> added here and
// removed again implicitly (setLambda/restoreDiagonal operate on copies).
bool solve_system(BA& s, double lambda) {
  double t0 = now_s();
  std::fill(s.S.val.begin(), s.S.val.end(), 0.);
  for (int i = 0; i < s.np; i++) {
    double* d = &s.S.val[(size_t)s.S.find(i, i) * 36];
    std::memcpy(d, &s.Hpp[(size_t)i * 36], 36 * sizeof(double));
    for (int k = 0; k < 6; k++) d[k * 7] += lambda;
  }
  std::fill(s.coeff.begin(), s.coeff.end(), 0.);
  for (int l = 0; l < s.nl; l++) {
    double D[9];
    std::memcpy(D, &s.Hll[(size_t)l * 9], sizeof(D));
    D[0] += lambda; D[4] += lambda; D[8] += lambda;
    double* Dinv = &s.Dinv[(size_t)l * 9];
    inv3(D, Dinv);
    double db[3];
    mat3vec(Dinv, &s.bl[(size_t)l * 3], db);
    for (int a = s.lm_ptr[l]; a < s.lm_ptr[l + 1]; a++) {
      const int ea = s.lm_edges[a];
      const int i1 = s.pose_idx[s.edges[s.active[ea]].kf];
      const double* Bi = &s.Hpl[(size_t)ea * 18];
      double BD[18];
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 3; c++) BD[r * 3 + c] = Bi[r * 3] * Dinv[c] + Bi[r * 3 + 1] * Dinv[3 + c] + Bi[r * 3 + 2] * Dinv[6 + c];
      for (int r = 0; r < 6; r++) s.coeff[(size_t)i1 * 6 + r] += Bi[r * 3] * db[0] + Bi[r * 3 + 1] * db[1] + Bi[r * 3 + 2] * db[2];
      int q = s.S.rowptr[i1];
      for (int b = a; b < s.lm_ptr[l + 1]; b++) {
        const int eb = s.lm_edges[b];
        const int i2 = s.pose_idx[s.edges[s.active[eb]].kf];
        while (s.S.col[q] < i2) ++q;
        double* H = &s.S.val[(size_t)q * 36];
        const double* Bj = &s.Hpl[(size_t)eb * 18];
        for (int r = 0; r < 6; r++)
          for (int c = 0; c < 6; c++) H[r * 6 + c] -= BD[r * 3] * Bj[c * 3] + BD[r * 3 + 1] * Bj[c * 3 + 1] + BD[r * 3 + 2] * Bj[c * 3 + 2];
      }
    }
  }
  for (int i = 0; i < s.np * 6; i++) s.bschur[i] = s.bp[i] - s.coeff[i];
  s.t_schur += now_s() - t0;
  t0 = now_s();
  bool ok = s.ldlt.factorize(s.S);
  if (ok) s.ldlt.solve(s.bschur.data(), s.x.data());
  s.t_solve += now_s() - t0;
  if (!ok) return false;
  // xl = Dinv * (bl - Hpl^T xp)
  t0 = now_s();
  double* xl = s.x.data() + (size_t)s.np * 6;
  for (int l = 0; l < s.nl; l++) {
    double c[3] = {s.bl[(size_t)l * 3], s.bl[(size_t)l * 3 + 1], s.bl[(size_t)l * 3 + 2]};
    for (int a = s.lm_ptr[l]; a < s.lm_ptr[l + 1]; a++) {
      const int ea = s.lm_edges[a];
      const int i1 = s.pose_idx[s.edges[s.active[ea]].kf];
      const double* B = &s.Hpl[(size_t)ea * 18];
      const double* xp = &s.x[(size_t)i1 * 6];
      // _HplCCS->rightMultiply(cl, -xp): per block  y.segment<3>() += B^T * (-x).segment<6>()  (sparse_block_matrix_ccs.h:108-128,
      // matrix_operations.h:54-57) — the six-term product is formed first, then added
      for (int cc = 0; cc < 3; cc++) {
        double d = B[cc] * -xp[0];
        for (int r = 1; r < 6; r++) d += B[r * 3 + cc] * -xp[r];
        c[cc] += d;
      }
    }
    mat3vec(&s.Dinv[(size_t)l * 9], c, xl + (size_t)l * 3);
  }
  s.t_schur += now_s() - t0;
  return true;
}

void push(BA& s) { s.pose_bak = s.pose; s.pt_bak = s.pt; }
void pop(BA& s) { s.pose = s.pose_bak; s.pt = s.pt_bak; }

void update(BA& s) {  // SparseOptimizer::update -> oplus (G/types/types_six_dof_expmap.h:73-76, G/types/types_sba.h:52-56)
  for (int i = 0; i < s.np; i++) {
    int k = s.idx_pose[i];
    s.pose[k] = se3_mul(se3_exp(&s.x[(size_t)i * 6]), s.pose[k]);
  }
  const double* xl = s.x.data() + (size_t)s.np * 6;
  for (int l = 0; l < s.nl; l++) {
    int j = s.idx_pt[l];
    for (int c = 0; c < 3; c++) s.pt[3 * (size_t)j + c] += xl[(size_t)l * 3 + c];
  }
}

double lambda_init(const BA& s) {  // computeLambdaInit, G/core/optimization_algorithm_levenberg.cpp:166-180
  double m = 0;
  for (int i = 0; i < s.np; i++)
    for (int j = 0; j < 6; j++) m = std::max(std::fabs(s.Hpp[(size_t)i * 36 + j * 7]), m);
  for (int l = 0; l < s.nl; l++)
    for (int j = 0; j < 3; j++) m = std::max(std::fabs(s.Hll[(size_t)l * 9 + j * 4]), m);
  return 1e-5 * m;
}

double compute_scale(const BA& s, double lambda) {  // G/core/optimization_algorithm_levenberg.cpp:182-189
  double scale = 0;
  for (int j = 0; j < s.np * 6; j++) scale += s.x[j] * (lambda * s.x[j] + s.bp[j]);
  const double* xl = s.x.data() + (size_t)s.np * 6;
  for (int j = 0; j < s.nl * 3; j++) scale += xl[j] * (lambda * xl[j] + s.bl[j]);
  return scale;
}

}  // namespace

extern "C" int orc_ba_solve(const orc_ba_problem* p, const orc_ba_options* o, orc_ba_result* r) {
  double T0 = now_s();
  BA s;
  load(s, p, o->robust, o->huber_delta);
  init_active(s);
  r->trace_len = 0; r->iters_done = 0; r->trials_total = 0;
  r->chi2_initial = r->chi2_final = 0; r->lambda_final = 0;
  auto terminate = [&]() { return o->stop && *o->stop; };
  int ret_iters = 0;
  if (s.np + s.nl == 0 || s.active.empty()) {
    ret_iters = -1;  // "0 vertices to optimize"
  } else {
    double lambda = -1, ni = 2;
    int nBad = 0;
    bool ok = true;
    const int max_trials = o->max_trials > 0 ? o->max_trials : 10;
    for (int it = 0; it < o->iterations && !terminate() && ok; it++) {
      if (it == 0) build_structure(s);
      compute_active_errors(s);
      double currentChi = active_robust_chi2(s);
      double tempChi = currentChi;
      const double iniChi = currentChi;
      if (it == 0) r->chi2_initial = currentChi;
      build_system(s);
      if (it == 0) {
        lambda = o->lambda_init > 0 ? o->lambda_init : lambda_init(s);
        ni = 2; nBad = 0;
      }
      double rho = 0, lambda_used = lambda;
      int qmax = 0;
      do {
        push(s);
        lambda_used = lambda;
        bool ok2 = solve_system(s, lambda);
        update(s);
        compute_active_errors(s);
        tempChi = active_robust_chi2(s);
        if (!ok2) tempChi = std::numeric_limits<double>::max();
        rho = (currentChi - tempChi);
        double scale = compute_scale(s, lambda);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(tempChi)) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, 2. / 3.);
          double scaleFactor = std::max(1. / 3., alpha);
          lambda *= scaleFactor;
          ni = 2;
          currentChi = tempChi;
        } else {
          lambda *= ni;
          ni *= 2;
          pop(s);
        }
        qmax++;
        r->trials_total++;
      } while (rho < 0 && qmax < max_trials && !terminate());
      ret_iters++;
      if (r->trace && r->trace_len < r->trace_cap) {
        double* tr = r->trace + (size_t)r->trace_len * ORC_TRACE_COLS;
        tr[0] = it; tr[1] = lambda_used; tr[2] = currentChi; tr[3] = rho; tr[4] = qmax; tr[5] = lambda;
        r->trace_len++;
      }
      r->chi2_final = currentChi; r->lambda_final = lambda;
      if (qmax == max_trials || rho == 0) { ok = false; continue; }  // Terminate
      if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
      if (nBad >= 3) { ok = false; continue; }
    }
  }
  r->iters_done = ret_iters;
  for (int k = 0; k < s.K; k++) se3_store(s.pose[k], r->poses + 7 * (size_t)k);
  std::memcpy(r->points, s.pt.data(), sizeof(double) * 3 * (size_t)s.P);
  if (r->chi2)
    for (int e : s.active) r->chi2[e] = edge_chi2(s.edges[e]);  // stale-by-design for a rejected last trial
  if (r->depth_pos)
    for (int e = 0; e < s.E; e++) {  // isDepthPositive on the final estimate, G/types/types_six_dof_expmap.h:97-101
      double xc[3];
      se3_map(s.pose[s.edges[e].kf], &s.pt[3 * (size_t)s.edges[e].mp], xc);
      r->depth_pos[e] = xc[2] > 0.0;
    }
  r->t_build_s = s.t_build; r->t_schur_s = s.t_schur; r->t_solve_s = s.t_solve; r->t_resid_s = s.t_resid;
  r->t_structure_s = s.t_structure;
  r->t_total_s = now_s() - T0;
  return 0;
}

extern "C" double orc_ba_linearize(const orc_ba_problem* p, int robust, double huber_delta, double* err,
                                   double* Jpose, double* Jpoint, double* rho1, double* chi2) {
  BA s;
  load(s, p, robust, huber_delta);
  init_active(s);
  compute_active_errors(s);
  Lin L;
  for (int e = 0; e < s.E; e++) {
    Edge& ed = s.edges[e];
    if (!ed.active) edge_error(s, ed);
    linearize(s, ed, L);
    if (err) { err[2 * e] = ed.err[0]; err[2 * e + 1] = ed.err[1]; }
    if (Jpose) std::memcpy(Jpose + 12 * (size_t)e, L.Jp, sizeof(L.Jp));
    if (Jpoint) std::memcpy(Jpoint + 6 * (size_t)e, L.Jl, sizeof(L.Jl));
    double c = edge_chi2(ed), rho[3] = {c, 1, 0};
    if (ed.robust) huber(c, huber_delta, rho);
    if (rho1) rho1[e] = rho[1];
    if (chi2) chi2[e] = c;
  }
  return active_robust_chi2(s);
}

extern "C" void orc_ba_build(const orc_ba_problem* p, int robust, double huber_delta, double* Hpp, double* bp,
                             double* Hll, double* bl, double* W) {
  BA s;
  load(s, p, robust, huber_delta);
  init_active(s);
  build_structure(s);
  compute_active_errors(s);
  build_system(s);
  std::memset(Hpp, 0, sizeof(double) * 36 * (size_t)s.K);
  std::memset(bp, 0, sizeof(double) * 6 * (size_t)s.K);
  std::memset(Hll, 0, sizeof(double) * 9 * (size_t)s.P);
  std::memset(bl, 0, sizeof(double) * 3 * (size_t)s.P);
  std::memset(W, 0, sizeof(double) * 18 * (size_t)s.E);
  for (int i = 0; i < s.np; i++) {
    std::memcpy(Hpp + 36 * (size_t)s.idx_pose[i], &s.Hpp[(size_t)i * 36], 36 * sizeof(double));
    std::memcpy(bp + 6 * (size_t)s.idx_pose[i], &s.bp[(size_t)i * 6], 6 * sizeof(double));
  }
  for (int l = 0; l < s.nl; l++) {
    std::memcpy(Hll + 9 * (size_t)s.idx_pt[l], &s.Hll[(size_t)l * 9], 9 * sizeof(double));
    std::memcpy(bl + 3 * (size_t)s.idx_pt[l], &s.bl[(size_t)l * 3], 3 * sizeof(double));
  }
  for (size_t a = 0; a < s.active.size(); a++) std::memcpy(W + 18 * (size_t)s.active[a], &s.Hpl[a * 18], 18 * sizeof(double));
}

extern "C" int orc_ba_schur_solve(const orc_ba_problem* p, int robust, double huber_delta, double lambda,
                                  double* dx_pose, double* dx_point, double* S_dense, double* bschur) {
  BA s;
  load(s, p, robust, huber_delta);
  init_active(s);
  build_structure(s);
  compute_active_errors(s);
  build_system(s);
  bool ok = solve_system(s, lambda);
  std::memset(dx_pose, 0, sizeof(double) * 6 * (size_t)s.K);
  std::memset(dx_point, 0, sizeof(double) * 3 * (size_t)s.P);
  for (int i = 0; i < s.np; i++) std::memcpy(dx_pose + 6 * (size_t)s.idx_pose[i], &s.x[(size_t)i * 6], 6 * sizeof(double));
  const double* xl = s.x.data() + (size_t)s.np * 6;
  for (int l = 0; l < s.nl; l++) std::memcpy(dx_point + 3 * (size_t)s.idx_pt[l], xl + (size_t)l * 3, 3 * sizeof(double));
  if (S_dense) {
    const size_t n = 6 * (size_t)s.K;
    std::memset(S_dense, 0, sizeof(double) * n * n);
    for (int i = 0; i < s.np; i++)
      for (int q = s.S.rowptr[i]; q < s.S.rowptr[i + 1]; q++) {
        int j = s.S.col[q];
        size_t gi = (size_t)s.idx_pose[i] * 6, gj = (size_t)s.idx_pose[j] * 6;
        for (int r = 0; r < 6; r++)
          for (int c = 0; c < 6; c++) {
            double v = s.S.val[(size_t)q * 36 + r * 6 + c];
            S_dense[(gi + r) * n + gj + c] = v;
            S_dense[(gj + c) * n + gi + r] = v;
          }
      }
  }
  if (bschur) {
    std::memset(bschur, 0, sizeof(double) * 6 * (size_t)s.K);
    for (int i = 0; i < s.np; i++) std::memcpy(bschur + 6 * (size_t)s.idx_pose[i], &s.bschur[(size_t)i * 6], 6 * sizeof(double));
  }
  return ok ? 0 : 1;
}

extern "C" void orc_se3_exp(const double upd[6], double out[7]) { se3_store(se3_exp(upd), out); }
extern "C" void orc_se3_mul(const double a[7], const double b[7], double out[7]) {
  se3_store(se3_mul(se3_load(a), se3_load(b)), out);
}
extern "C" void orc_se3_map(const double qt[7], const double x[3], double out[3]) { se3_map(se3_load(qt), x, out); }
extern "C" void orc_pose_from_Tcw_f32(const float T[16], double out[7]) {
  double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  double t[3] = {T[3], T[7], T[11]};
  se3_store(se3_from_Rt(R, t), out);
}
extern "C" void orc_pose_to_Tcw_f32(const double qt[7], float T[16]) {
  SE3 s = se3_load(qt);
  double R[9];
  q2R(s.r, R);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[i * 4 + j] = (float)R[i * 3 + j];
    T[i * 4 + 3] = (float)s.t[i];
  }
  T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}
extern "C" void orc_huber(double e, double delta, double rho[3]) { huber(e, delta, rho); }

// sparse LDL^T as a service, for the LinearSolverEigen stand-in (oracle/ref_stub_g2o): the factorisation orc_ba_solve / orc_pgo_solve use
namespace { struct LdltHandle { BlockSym S; SparseLDLT ldlt; bool analyzed = false; }; }
extern "C" void* orc_ldlt_new(void) { return new LdltHandle; }
extern "C" void orc_ldlt_free(void* h) { delete static_cast<LdltHandle*>(h); }
extern "C" void orc_ldlt_reset(void* h) { static_cast<LdltHandle*>(h)->analyzed = false; }
extern "C" int orc_ldlt_solve(void* hv, int nb, int bs, const int* rowptr, const int* col, const double* val, const double* b, double* x) {
  LdltHandle* h = static_cast<LdltHandle*>(hv);
  h->S.nb = nb; h->S.bs = bs;
  h->S.rowptr.assign(rowptr, rowptr + nb + 1);
  h->S.col.assign(col, col + rowptr[nb]);
  h->S.val.assign(val, val + (size_t)rowptr[nb] * bs * bs);
  if (!h->analyzed) { h->ldlt.analyze(h->S); h->analyzed = true; }
  if (!h->ldlt.factorize(h->S)) return 1;
  h->ldlt.solve(b, x);
  return 0;
}


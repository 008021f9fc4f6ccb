// pgo_oracle.cpp — CPU oracle for the Sim3 essential-graph optimisation (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Restates what cslam::Optimizer::OptimizeEssentialGraph{LoopClosure,MapFusion} make g2o do in
// optimizer.optimize(20) (S/Optimizer.cpp:1277, 1513): BlockSolver_7_3 without Schur
// (G/core/optimization_algorithm_with_hessian.cpp:50-74), LinearSolverEigen (direct LDL^T),
// Levenberg with user lambda 1e-16 (S/Optimizer.cpp:1071,1344), EdgeSim3 with identity information
// (G/types/types_seven_dof_expmap.h:99-126) and NUMERIC Jacobians, central differences delta = 1e-9
// (G/core/base_binary_edge.hpp:131-205).  VertexSim3Expmap::oplusImpl zeroes update[6] when
// _fix_scale (G/types/types_seven_dof_expmap.h:60-69).
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "lie.hpp"
#include "oracle.h"
#include "sparse_ldlt.hpp"

namespace {
using namespace orc;

struct PEdge { int i, j; Sim3 C; double err[7]; };

struct PGO {
  int K, E;
  bool fix_scale;
  std::vector<Sim3> v, v_bak;
  std::vector<uint8_t> fixed;
  std::vector<PEdge> edges;
  std::vector<int> active;
  std::vector<int> vidx, idxv;
  int n = 0;
  BlockSym H;
  std::vector<int> eblk;  // per active edge: index of off-diagonal block in H (or -1), and whether transposed
  std::vector<char> etr;
  std::vector<double> b, x;
  SparseLDLT ldlt;
};

inline Sim3 oplus(const Sim3& s, double* upd, bool fix_scale) {
  if (fix_scale) upd[6] = 0;
  return sim3_mul(sim3_exp(upd), s);
}
inline void edge_err(const PEdge& e, const Sim3& vi, const Sim3& vj, double err[7]) {  // computeError, :105-114
  Sim3 er = sim3_mul(sim3_mul(e.C, vi), sim3_inv(vj));
  sim3_log(er, err);
}

void numeric_jac(const PGO& s, const PEdge& e, bool iFree, bool jFree, double Ji[49], double Jj[49], bool analytic) {
  (void)analytic;
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  double ep[7], em[7], add[7];
  std::memset(Ji, 0, 49 * sizeof(double));
  std::memset(Jj, 0, 49 * sizeof(double));
  if (iFree)
    for (int d = 0; d < 7; d++) {
      std::memset(add, 0, sizeof(add));
      add[d] = delta;
      edge_err(e, oplus(s.v[e.i], add, s.fix_scale), s.v[e.j], ep);
      add[d] = -delta;
      edge_err(e, oplus(s.v[e.i], add, s.fix_scale), s.v[e.j], em);
      for (int r = 0; r < 7; r++) Ji[r * 7 + d] = scalar * (ep[r] - em[r]);
    }
  if (jFree)
    for (int d = 0; d < 7; d++) {
      std::memset(add, 0, sizeof(add));
      add[d] = delta;
      edge_err(e, s.v[e.i], oplus(s.v[e.j], add, s.fix_scale), ep);
      add[d] = -delta;
      edge_err(e, s.v[e.i], oplus(s.v[e.j], add, s.fix_scale), em);
      for (int r = 0; r < 7; r++) Jj[r * 7 + d] = scalar * (ep[r] - em[r]);
    }
}

}  // namespace

extern "C" int orc_pgo_solve(const orc_pgo_problem* p, int32_t iterations, double lambda_init, int32_t analytic_jac,
                             const volatile uint8_t* stop, orc_pgo_result* r) {
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double T0 = now();
  PGO s;
  s.K = p->K; s.E = p->E; s.fix_scale = p->fix_scale != 0;
  s.v.resize(s.K);
  for (int k = 0; k < s.K; k++) s.v[k] = sim3_load(p->sim3 + 8 * (size_t)k);
  s.fixed.assign(p->fixed, p->fixed + s.K);
  s.edges.resize(s.E);
  for (int e = 0; e < s.E; e++) {
    s.edges[e].i = p->edge_i[e]; s.edges[e].j = p->edge_j[e];
    s.edges[e].C = sim3_load(p->meas + 8 * (size_t)e);
  }
  std::vector<char> has(s.K, 0);
  for (int e = 0; e < s.E; e++) {
    const PEdge& ed = s.edges[e];
    if (s.fixed[ed.i] && s.fixed[ed.j]) continue;
    s.active.push_back(e);
    has[ed.i] = has[ed.j] = 1;
  }
  s.vidx.assign(s.K, -1);
  for (int k = 0; k < s.K; k++)
    if (has[k] && !s.fixed[k]) { s.vidx[k] = (int)s.idxv.size(); s.idxv.push_back(k); }
  s.n = (int)s.idxv.size();
  r->trace_len = 0; r->iters_done = 0; r->chi2_initial = r->chi2_final = 0; r->lambda_final = 0;
  auto terminate = [&] { return stop && *stop; };
  auto errors = [&] { for (int e : s.active) { PEdge& ed = s.edges[e]; edge_err(ed, s.v[ed.i], s.v[ed.j], ed.err); } };
  // activeRobustChi2: chi += e->chi2(), each edge's e'e summed on its own first (G/core/sparse_optimizer.cpp:100-114, base_edge.h:58-61)
  auto chi2 = [&] {
    double c = 0;
    for (int e : s.active) {
      const double* er = s.edges[e].err;
      double ce = er[0] * er[0];
      for (int k = 1; k < 7; k++) ce += er[k] * er[k];
      c += ce;
    }
    return c;
  };
  int ret = 0;
  if (s.n == 0 || s.active.empty()) {
    ret = -1;
  } else {
    // structure
    std::vector<std::vector<int>> rows(s.n);
    for (int i = 0; i < s.n; i++) rows[i].push_back(i);
    for (int e : s.active) {
      int a = s.vidx[s.edges[e].i], b = s.vidx[s.edges[e].j];
      if (a >= 0 && b >= 0 && a != b) rows[std::min(a, b)].push_back(std::max(a, b));
    }
    s.H.nb = s.n; s.H.bs = 7; s.H.rowptr.assign(s.n + 1, 0);
    for (int i = 0; i < s.n; i++) {
      std::sort(rows[i].begin(), rows[i].end());
      rows[i].erase(std::unique(rows[i].begin(), rows[i].end()), rows[i].end());
      s.H.col.insert(s.H.col.end(), rows[i].begin(), rows[i].end());
      s.H.rowptr[i + 1] = (int)s.H.col.size();
    }
    s.H.val.assign(s.H.col.size() * 49, 0);
    s.ldlt.analyze(s.H);
    s.b.assign((size_t)s.n * 7, 0); s.x.assign((size_t)s.n * 7, 0);
    std::vector<double> Hbak;
    double lambda = -1, ni = 2;
    int nBad = 0;
    bool ok = true;
    for (int it = 0; it < iterations && !terminate() && ok; it++) {
      errors();
      double currentChi = chi2(), tempChi = currentChi;
      const double iniChi = currentChi;
      if (it == 0) r->chi2_initial = currentChi;
      // buildSystem
      std::fill(s.H.val.begin(), s.H.val.end(), 0.);
      std::fill(s.b.begin(), s.b.end(), 0.);
      double Ji[49], Jj[49];
      for (int e : s.active) {
        const PEdge& ed = s.edges[e];
        int a = s.vidx[ed.i], b = s.vidx[ed.j];
        numeric_jac(s, ed, a >= 0, b >= 0, Ji, Jj, analytic_jac != 0);
        if (a >= 0) {
          double* Haa = &s.H.val[(size_t)s.H.find(a, a) * 49];
          for (int r_ = 0; r_ < 7; r_++) {
            double acc = 0;
            for (int k = 0; k < 7; k++) acc += Ji[k * 7 + r_] * (-ed.err[k]);
            s.b[(size_t)a * 7 + r_] += acc;
            for (int c = 0; c < 7; c++) {
              double h = 0;
              for (int k = 0; k < 7; k++) h += Ji[k * 7 + r_] * Ji[k * 7 + c];
              Haa[r_ * 7 + c] += h;
            }
          }
        }
        if (b >= 0) {
          double* Hbb = &s.H.val[(size_t)s.H.find(b, b) * 49];
          for (int r_ = 0; r_ < 7; r_++) {
            double acc = 0;
            for (int k = 0; k < 7; k++) acc += Jj[k * 7 + r_] * (-ed.err[k]);
            s.b[(size_t)b * 7 + r_] += acc;
            for (int c = 0; c < 7; c++) {
              double h = 0;
              for (int k = 0; k < 7; k++) h += Jj[k * 7 + r_] * Jj[k * 7 + c];
              Hbb[r_ * 7 + c] += h;
            }
          }
        }
        if (a >= 0 && b >= 0 && a != b) {
          bool tr = a > b;
          double* Hab = &s.H.val[(size_t)s.H.find(std::min(a, b), std::max(a, b)) * 49];
          for (int r_ = 0; r_ < 7; r_++)
            for (int c = 0; c < 7; c++) {
              double h = 0;
              for (int k = 0; k < 7; k++) h += Ji[k * 7 + r_] * Jj[k * 7 + c];  // A^T B  (block (a,b))
              if (!tr) Hab[r_ * 7 + c] += h; else Hab[c * 7 + r_] += h;
            }
        }
      }
      if (it == 0) {
        if (lambda_init > 0) lambda = lambda_init;
        else {
          double m = 0;
          for (int i = 0; i < s.n; i++)
            for (int k = 0; k < 7; k++) m = std::max(m, std::fabs(s.H.val[(size_t)s.H.find(i, i) * 49 + k * 8]));
          lambda = 1e-5 * m;
        }
        ni = 2; nBad = 0;
      }
      double rho = 0;
      int qmax = 0;
      double lambda_used = lambda;
      do {
        s.v_bak = s.v;
        lambda_used = lambda;
        Hbak = s.H.val;
        for (int i = 0; i < s.n; i++) {
          double* d = &s.H.val[(size_t)s.H.find(i, i) * 49];
          for (int k = 0; k < 7; k++) d[k * 8] += lambda;
        }
        bool ok2 = s.ldlt.factorize(s.H);
        if (ok2) s.ldlt.solve(s.b.data(), s.x.data());
        for (int i = 0; i < s.n; i++) s.v[s.idxv[i]] = oplus(s.v[s.idxv[i]], &s.x[(size_t)i * 7], s.fix_scale);
        s.H.val = Hbak;  // restoreDiagonal
        errors();
        tempChi = chi2();
        if (!ok2) tempChi = std::numeric_limits<double>::max();
        rho = currentChi - tempChi;
        double scale = 0;
        for (int j = 0; j < s.n * 7; j++) scale += s.x[j] * (lambda * s.x[j] + s.b[j]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(tempChi)) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, 2. / 3.);
          lambda *= std::max(1. / 3., alpha);
          ni = 2;
          currentChi = tempChi;
        } else {
          lambda *= ni;
          ni *= 2;
          s.v = s.v_bak;
        }
        qmax++;
      } while (rho < 0 && qmax < 10 && !terminate());
      ret++;
      if (r->trace && r->trace_len < r->trace_cap) {
        double* tr = r->trace + (size_t)r->trace_len * ORC_TRACE_COLS;
        tr[0] = it; tr[1] = lambda_used; tr[2] = currentChi; tr[3] = rho; tr[4] = qmax; tr[5] = lambda;
        r->trace_len++;
      }
      r->chi2_final = currentChi; r->lambda_final = lambda;
      if (qmax == 10 || rho == 0) { ok = false; continue; }
      if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
      if (nBad >= 3) { ok = false; continue; }
    }
  }
  r->iters_done = ret;
  for (int k = 0; k < s.K; k++) sim3_store(s.v[k], r->sim3 + 8 * (size_t)k);
  r->t_total_s = now() - T0;
  return 0;
}

extern "C" void orc_sim3_exp(const double upd[7], double out[8]) { sim3_store(sim3_exp(upd), out); }
extern "C" void orc_sim3_log(const double s[8], double out[7]) { sim3_log(sim3_load(s), out); }
extern "C" void orc_sim3_mul(const double a[8], const double b[8], double out[8]) {
  sim3_store(sim3_mul(sim3_load(a), sim3_load(b)), out);
}
extern "C" void orc_sim3_inv(const double a[8], double out[8]) { sim3_store(sim3_inv(sim3_load(a)), out); }
extern "C" void orc_pgo_edge_error(const double meas[8], const double si[8], const double sj[8], double err[7]) {
  PEdge e;
  e.C = sim3_load(meas);
  edge_err(e, sim3_load(si), sim3_load(sj), err);
}
extern "C" void orc_pgo_edge_jacobian(const double meas[8], const double si[8], const double sj[8], int fix_scale, double Ji[49], double Jj[49]) {
  PGO s;
  s.fix_scale = fix_scale != 0;
  s.v = {sim3_load(si), sim3_load(sj)};
  PEdge e;
  e.i = 0; e.j = 1; e.C = sim3_load(meas);
  numeric_jac(s, e, true, true, Ji, Jj, false);
}
extern "C" void orc_sim3_map(const double s[8], const double x[3], double out[3]) { sim3_map(sim3_load(s), x, out); }


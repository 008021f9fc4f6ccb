// sparse_ldlt.hpp — direct sparse LDL^T of the CPU oracle (TEST INFRASTRUCTURE).
//
// Stands in for Eigen::SimplicialLDLT<SparseMatrix, Upper> + AMD as used by g2o's
// LinearSolverEigen (G/solvers/linear_solver_eigen.h:72-88,106-133,159-213): one fill-reducing
// ordering per optimize(), then a scalar up-looking LDL^T (elimination tree + sparse triangular
// solves — the algorithm of T. Davis' LDL package, which is what Eigen's SimplicialCholesky
// implements) per LM trial.  Eigen is absent from this container, so the ordering is a
// block-level greedy minimum-degree (equivalent quality: all scalars of a block share one
// adjacency).  Failure convention as Eigen's LDLT: a pivot exactly 0 -> false.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace orc {

// Block-symmetric matrix, upper block-triangle stored row-wise: row i holds sorted cols j >= i.
struct BlockSym {
  int nb = 0, bs = 0;                 // number of block rows, block size
  std::vector<int> rowptr, col;       // CSR over upper blocks
  std::vector<double> val;            // nnzb * bs*bs, each block row-major
  int find(int i, int j) const {      // j >= i
    const int* b = col.data() + rowptr[i];
    const int* e = col.data() + rowptr[i + 1];
    const int* it = std::lower_bound(b, e, j);
    return (it != e && *it == j) ? int(it - col.data()) : -1;
  }
};

// greedy minimum degree on the block graph; returns perm (new -> old)
inline std::vector<int> min_degree_order(const BlockSym& A) {
  const int n = A.nb;
  std::vector<std::vector<int>> adj(n);
  for (int i = 0; i < n; i++)
    for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; p++) {
      int j = A.col[p];
      if (j != i) { adj[i].push_back(j); adj[j].push_back(i); }
    }
  for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
  std::vector<char> done(n, 0);
  std::vector<int> perm;
  perm.reserve(n);
  std::vector<int> mark(n, -1), tmp;
  for (int step = 0; step < n; step++) {
    int best = -1;
    size_t bd = SIZE_MAX;
    for (int v = 0; v < n; v++)
      if (!done[v] && adj[v].size() < bd) { bd = adj[v].size(); best = v; }
    const int v = best;
    done[v] = 1;
    perm.push_back(v);
    // eliminate v: neighbours form a clique
    const std::vector<int> nb = adj[v];
    for (int u : nb) {
      // adj[u] = (adj[u] U nb) \ {u, v}
      tmp.clear();
      std::set_union(adj[u].begin(), adj[u].end(), nb.begin(), nb.end(), std::back_inserter(tmp));
      tmp.erase(std::remove_if(tmp.begin(), tmp.end(), [&](int w) { return w == u || w == v; }), tmp.end());
      adj[u].swap(tmp);
    }
    adj[v].clear();
    adj[v].shrink_to_fit();
  }
  return perm;
}

struct SparseLDLT {
  int n = 0;
  std::vector<int> Ap, Ai;       // permuted upper-triangular CSC pattern
  std::vector<int> srcblk, srcoff, srctr;  // where each CSC entry comes from in BlockSym.val
  std::vector<double> Ax;
  std::vector<int> Lp, Parent, Lnz, Li, Flag, Pattern;
  std::vector<double> Lx, D, Y;
  std::vector<int> pinv_scalar;  // old scalar -> new scalar

  // symbolic analysis once per optimize(): ordering + elimination tree + column counts
  void analyze(const BlockSym& A) {
    const int bs = A.bs;
    n = A.nb * bs;
    std::vector<int> perm = min_degree_order(A);
    std::vector<int> pinvb(A.nb);
    for (int k = 0; k < A.nb; k++) pinvb[perm[k]] = k;
    pinv_scalar.resize(n);
    for (int b = 0; b < A.nb; b++)
      for (int r = 0; r < bs; r++) pinv_scalar[b * bs + r] = pinvb[b] * bs + r;
    // count entries per permuted column
    std::vector<int> cnt(n + 1, 0);
    auto visit = [&](auto&& f) {
      for (int i = 0; i < A.nb; i++)
        for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; p++) {
          int j = A.col[p];
          for (int r = 0; r < bs; r++)
            for (int c = 0; c < bs; c++) {
              if (i == j && c < r) continue;  // scalar upper triangle of a diagonal block only
              int pr = pinv_scalar[i * bs + r], pc = pinv_scalar[j * bs + c];
              int rr = std::min(pr, pc), cc = std::max(pr, pc);
              f(rr, cc, p, r * bs + c);
            }
        }
    };
    visit([&](int, int cc, int, int) { cnt[cc + 1]++; });
    Ap.assign(n + 1, 0);
    for (int k = 0; k < n; k++) Ap[k + 1] = Ap[k] + cnt[k + 1];
    const int nnz = Ap[n];
    Ai.resize(nnz); srcblk.resize(nnz); srcoff.resize(nnz); Ax.resize(nnz);
    std::vector<int> fill(Ap.begin(), Ap.end() - 1);
    visit([&](int rr, int cc, int p, int off) {
      int q = fill[cc]++;
      Ai[q] = rr; srcblk[q] = p; srcoff[q] = off;
    });
    // ldl_symbolic
    Lp.assign(n + 1, 0); Parent.assign(n, -1); Lnz.assign(n, 0); Flag.assign(n, 0);
    for (int k = 0; k < n; k++) {
      Parent[k] = -1; Flag[k] = k; Lnz[k] = 0;
      for (int p = Ap[k]; p < Ap[k + 1]; p++) {
        int i = Ai[p];
        if (i < k)
          for (; Flag[i] != k; i = Parent[i]) {
            if (Parent[i] == -1) Parent[i] = k;
            Lnz[i]++;
            Flag[i] = k;
          }
      }
    }
    for (int k = 0; k < n; k++) Lp[k + 1] = Lp[k] + Lnz[k];
    Li.resize(Lp[n]); Lx.resize(Lp[n]); D.resize(n); Y.resize(n); Pattern.resize(n);
  }

  // numeric factorisation; false on an exactly-zero pivot (Eigen LDLT convention)
  bool factorize(const BlockSym& A) {
    const int bb = A.bs * A.bs;
    for (size_t q = 0; q < Ax.size(); q++) Ax[q] = A.val[(size_t)srcblk[q] * bb + srcoff[q]];
    for (int k = 0; k < n; k++) {
      Y[k] = 0.0;
      int top = n;
      Flag[k] = k;
      Lnz[k] = 0;
      for (int p = Ap[k]; p < Ap[k + 1]; p++) {
        int i = Ai[p];
        if (i <= k) {
          Y[i] += Ax[p];
          int len;
          for (len = 0; Flag[i] != k; i = Parent[i]) { Pattern[len++] = i; Flag[i] = k; }
          while (len > 0) Pattern[--top] = Pattern[--len];
        }
      }
      D[k] = Y[k];
      Y[k] = 0.0;
      for (; top < n; top++) {
        int i = Pattern[top];
        double yi = Y[i];
        Y[i] = 0.0;
        int p2 = Lp[i] + Lnz[i], p;
        for (p = Lp[i]; p < p2; p++) Y[Li[p]] -= Lx[p] * yi;
        double l_ki = yi / D[i];
        D[k] -= l_ki * yi;
        Li[p] = k;
        Lx[p] = l_ki;
        Lnz[i]++;
      }
      if (D[k] == 0.0) return false;
    }
    return true;
  }

  void solve(const double* b, double* x) const {
    std::vector<double> w(n);
    for (int i = 0; i < n; i++) w[pinv_scalar[i]] = b[i];
    for (int j = 0; j < n; j++) {
      double xj = w[j];
      for (int p = Lp[j]; p < Lp[j + 1]; p++) w[Li[p]] -= Lx[p] * xj;
    }
    for (int j = 0; j < n; j++) w[j] /= D[j];
    for (int j = n - 1; j >= 0; j--) {
      double s = w[j];
      for (int p = Lp[j]; p < Lp[j + 1]; p++) s -= Lx[p] * w[Li[p]];
      w[j] = s;
    }
    for (int i = 0; i < n; i++) x[i] = w[pinv_scalar[i]];
  }
};

}  // namespace orc

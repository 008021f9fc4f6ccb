// Stand-in for g2o/solvers/linear_solver_dense.h (TEST INFRASTRUCTURE): the reference's LinearSolverDense copies the block matrix into a
// dense one and factorises it with Eigen::LDLT (G/solvers/linear_solver_dense.h:64-113).  Same class name and interface; the
// factorisation is the CPU oracle's dense Cholesky (liboracle.so: orc_chol_solve), as in the oracle's single-vertex optimisations.
#ifndef G2O_LINEAR_SOLVER_DENSE_H
#define G2O_LINEAR_SOLVER_DENSE_H
#include <vector>

#include <core/linear_solver.h>

extern "C" int orc_chol_solve(int n, const double* A /*n x n row-major*/, const double* b, double* x);   /* 0 = solved, 1 = not positive definite */

namespace g2o {

template <typename MatrixType>
class LinearSolverDense : public LinearSolver<MatrixType> {
 public:
  LinearSolverDense() : LinearSolver<MatrixType>() {}
  virtual ~LinearSolverDense() {}
  virtual bool init() { return true; }
  bool solve(const SparseBlockMatrix<MatrixType>& A, double* x, double* b) {
    // densify: A keeps block (i, j), i <= j, in the map of block column j; mirror it below the diagonal
    const int n = A.cols();
    std::vector<double> dense((size_t)n * n, 0.);
    const int ncol = (int)A.blockCols().size();
    for (int j = 0; j < ncol; j++) {
      const int c0 = A.colBaseOfBlock(j);
      for (const auto& entry : A.blockCols()[j]) {
        const int i = entry.first;
        if (i > j) continue;
        const int r0 = A.rowBaseOfBlock(i);
        const MatrixType& blk = *entry.second;
        for (int c = 0; c < blk.cols(); c++)
          for (int r = 0; r < blk.rows(); r++) {
            dense[(size_t)(r0 + r) * n + (c0 + c)] = blk(r, c);
            if (i != j) dense[(size_t)(c0 + c) * n + (r0 + r)] = blk(r, c);
          }
      }
    }
    return orc_chol_solve(n, dense.data(), b, x) == 0;
  }
};

}  // namespace g2o
#endif

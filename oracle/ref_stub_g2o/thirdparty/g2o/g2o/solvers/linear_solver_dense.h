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
    const int n = A.cols();
    std::vector<double> H((size_t)n * n, 0.);
    for (size_t i = 0; i < A.blockCols().size(); ++i) {
      const int c_idx = A.colBaseOfBlock((int)i);
      const typename SparseBlockMatrix<MatrixType>::IntBlockMap& col = A.blockCols()[i];
      for (typename SparseBlockMatrix<MatrixType>::IntBlockMap::const_iterator it = col.begin(); it != col.end(); ++it) {
        if (it->first > (int)i) continue;                       // only the upper triangular block is processed
        const int r_idx = A.rowBaseOfBlock(it->first);
        const MatrixType& m = *(it->second);
        for (int r = 0; r < m.rows(); r++)
          for (int c = 0; c < m.cols(); c++) {
            H[(size_t)(r_idx + r) * n + c_idx + c] = m(r, c);
            if (r_idx != c_idx) H[(size_t)(c_idx + c) * n + r_idx + r] = m(r, c);
          }
      }
    }
    return orc_chol_solve(n, H.data(), b, x) == 0;
  }
};

}  // namespace g2o
#endif

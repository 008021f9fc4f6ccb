// Stand-in for g2o/solvers/linear_solver_eigen.h (TEST INFRASTRUCTURE): the reference's LinearSolverEigen wraps Eigen's sparse simplicial
// LDL^T, which the Eigen stand-in of oracle/ref_stub does not carry.  Same class name and interface (G/solvers/linear_solver_eigen.h:48-133);
// solve() hands the upper-triangle block matrix to the CPU oracle's sparse LDL^T (liboracle.so: orc_ldlt_*).
#ifndef G2O_LINEAR_SOLVER_EIGEN_H
#define G2O_LINEAR_SOLVER_EIGEN_H
#include <utility>
#include <vector>

#include <core/linear_solver.h>

extern "C" {
void* orc_ldlt_new(void);
void orc_ldlt_free(void* h);
void orc_ldlt_reset(void* h);
/* block CSR of the upper triangle (rows ascending, columns ascending within a row), bs x bs row-major blocks; analyses the pattern on the
 * first call after new / reset; returns 0 on success, 1 when the factorisation fails */
int orc_ldlt_solve(void* h, int nb, int bs, const int* rowptr, const int* col, const double* val, const double* b, double* x);
}

namespace g2o {

template <typename MatrixType>
class LinearSolverEigen : public LinearSolver<MatrixType> {
 public:
  LinearSolverEigen() : LinearSolver<MatrixType>(), h_(orc_ldlt_new()), _blockOrdering(false), _writeDebug(true) {}
  virtual ~LinearSolverEigen() { orc_ldlt_free(h_); }
  virtual bool init() { orc_ldlt_reset(h_); return true; }
  bool solve(const SparseBlockMatrix<MatrixType>& A, double* x, double* b) {
    const int n = (int)A.blockCols().size();
    if (n == 0) return true;
    const int bs = A.rowsOfBlock(0);
    std::vector<std::vector<std::pair<int, const MatrixType*> > > rows(n);
    for (int j = 0; j < n; j++)
      for (typename SparseBlockMatrix<MatrixType>::IntBlockMap::const_iterator it = A.blockCols()[j].begin(); it != A.blockCols()[j].end(); ++it)
        if (it->first <= j) rows[it->first].push_back(std::make_pair(j, (const MatrixType*)it->second));
    std::vector<int> rowptr(n + 1, 0), col;
    for (int i = 0; i < n; i++) { for (size_t q = 0; q < rows[i].size(); q++) col.push_back(rows[i][q].first); rowptr[i + 1] = (int)col.size(); }
    std::vector<double> val(col.size() * (size_t)bs * bs, 0.);
    size_t q = 0;
    for (int i = 0; i < n; i++)
      for (size_t k = 0; k < rows[i].size(); k++, q++)
        for (int r = 0; r < bs; r++) for (int c = 0; c < bs; c++) val[q * bs * bs + r * bs + c] = (*rows[i][k].second)(r, c);
    return orc_ldlt_solve(h_, n, bs, rowptr.data(), col.data(), val.data(), b, x) == 0;
  }
  bool blockOrdering() const { return _blockOrdering; }
  void setBlockOrdering(bool blockOrdering) { _blockOrdering = blockOrdering; }
  virtual bool writeDebug() const { return _writeDebug; }
  virtual void setWriteDebug(bool b) { _writeDebug = b; }
 protected:
  void* h_;
  bool _blockOrdering, _writeDebug;
};

}  // namespace g2o
#endif

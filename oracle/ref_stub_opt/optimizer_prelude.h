// Force-included in front of shim/Optimizer_shim.cpp when it is compiled against the reference's own cslam/Optimizer.h
// (TEST INFRASTRUCTURE): the stand-in graph classes for g2o's types/ headers, and "already seen" marks for the g2o solver headers
// Optimizer.h names but the shim does not use (they need Eigen's sparse module).
#include "../ref_stub/g2o_core_standin.h"
#define G2O_BLOCK_SOLVER_H
#define G2O_SOLVER_LEVENBERG_H
#define G2O_LINEAR_SOLVER_EIGEN_H
#define G2O_LINEAR_SOLVER_DENSE_H

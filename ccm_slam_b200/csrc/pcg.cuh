// pcg.cuh — persistent cooperative preconditioned conjugate gradients on a block-CSR SPD system (block size BS).
//
// Replaces the direct sparse LDL^T of g2o's LinearSolverEigen (G/solvers/linear_solver_eigen.h:106-133) on the reduced
// camera system (BS = 6, bundle adjustment) and on the Sim3 pose-graph Hessian (BS = 7, essential graph).  One warp per
// block row; the whole iteration runs inside one kernel launched with cudaLaunchCooperativeKernel, grid-wide barriers go
// through a global counter, dot products are reduced in a fixed order so that every CTA sees identical scalars.
// Preconditioner: block-Jacobi (inverse diagonal blocks, computed by the caller).
#pragma once
#include <cuda_runtime.h>

namespace ccm {

constexpr int TPB = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0.  smem must hold blockDim/32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double r = 0;
  if (wid == 0) {
    r = lane < (blockDim.x >> 5) ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

struct PcgArgs {
  int n;  // block rows
  const int* rowptr; const int* col; const double* val; const double* Minv; const double* b;
  double *x, *r, *z, *p, *q;
  double* partials;  // 3 * gridDim.x
  unsigned* bar;     // zeroed before launch
  double tol; int max_iter;
  double* status;    // [iters, relres, flag(0 converged, 1 max_iter, 2 breakdown: p'Sp <= 0)]
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    while (*(volatile unsigned*)bar < target) {}
    __threadfence();
  }
  __syncthreads();
}

// sum of partials[0..g) in a fixed order, same value in every thread of the calling warp
__device__ __forceinline__ double sum_partials_dev(const double* partials, int g) {
  double v = 0.0;
  for (int i = threadIdx.x & 31; i < g; i += 32) v += __ldcg(partials + i);
  return warp_sum(v);
}

template <int BS>
__global__ void __launch_bounds__(TPB) k_pcg(PcgArgs A) {
  constexpr int BB = BS * BS;
  __shared__ double red[TPB / 32];
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * TPB + threadIdx.x) >> 5;
  const int nw = (gridDim.x * TPB) >> 5;
  const int G = gridDim.x;
  unsigned target = 0;
  double* part0 = A.partials;
  double* part1 = A.partials + G;
  double* part2 = A.partials + 2 * G;

  // x = 0, r = b, z = Minv r, p = z
  double acc_rz = 0.0, acc_bb = 0.0;
  for (int a = gw; a < A.n; a += nw) {
    if (lane < BS) {
      const double rv = A.b[(size_t)a * BS + lane];
      const double* M = A.Minv + (size_t)a * BB + lane * BS;
      const double* ra = A.b + (size_t)a * BS;
      double zv = 0.0;
#pragma unroll
      for (int k = 0; k < BS; k++) zv += M[k] * ra[k];
      A.x[(size_t)a * BS + lane] = 0.0;
      A.r[(size_t)a * BS + lane] = rv;
      A.z[(size_t)a * BS + lane] = zv;
      A.p[(size_t)a * BS + lane] = zv;
      acc_rz += rv * zv;
      acc_bb += rv * rv;
    }
  }
  {
    const double t0 = block_sum(acc_rz, red);
    const double t1 = block_sum(acc_bb, red);
    if (threadIdx.x == 0) { part0[blockIdx.x] = t0; part1[blockIdx.x] = t1; }
  }
  grid_barrier(A.bar, target);
  double rz = sum_partials_dev(part0, G);
  const double bb = sum_partials_dev(part1, G);
  const double stop2 = A.tol * A.tol * bb;
  int it = 0, flag = 1;
  double rr = bb;
  if (!(bb > 0.0)) {
    flag = 0;
  } else {
    for (it = 0; it < A.max_iter; it++) {
      // q = S p ; pq = p.q
      double acc_pq = 0.0;
      for (int a = gw; a < A.n; a += nw) {
        double y[BS];
#pragma unroll
        for (int k = 0; k < BS; k++) y[k] = 0.0;
        const int beg = A.rowptr[a], end = A.rowptr[a + 1];
        for (int j = beg + lane; j < end; j += 32) {
          const double* v = A.val + (size_t)j * BB;
          const double* pj = A.p + (size_t)A.col[j] * BS;
          double pv[BS];
#pragma unroll
          for (int k = 0; k < BS; k++) pv[k] = __ldcg(pj + k);
          if constexpr (BS % 2 == 0) {
            const double2* v2 = reinterpret_cast<const double2*>(v);
#pragma unroll
            for (int rI = 0; rI < BS; rI++)
#pragma unroll
              for (int c = 0; c < BS / 2; c++) {
                const double2 t = __ldg(v2 + rI * (BS / 2) + c);
                y[rI] += t.x * pv[2 * c] + t.y * pv[2 * c + 1];
              }
          } else {
#pragma unroll
            for (int rI = 0; rI < BS; rI++)
#pragma unroll
              for (int c = 0; c < BS; c++) y[rI] += __ldg(v + rI * BS + c) * pv[c];
          }
        }
#pragma unroll
        for (int k = 0; k < BS; k++) y[k] = warp_sum(y[k]);
        if (lane < BS) {
          double yl = y[0];
#pragma unroll
          for (int k = 1; k < BS; k++) yl = lane == k ? y[k] : yl;
          A.q[(size_t)a * BS + lane] = yl;
          acc_pq += yl * A.p[(size_t)a * BS + lane];
        }
      }
      {
        const double t0 = block_sum(acc_pq, red);
        if (threadIdx.x == 0) part0[blockIdx.x] = t0;
      }
      grid_barrier(A.bar, target);
      const double pq = sum_partials_dev(part0, G);
      if (!(pq > 0.0) || !isfinite(pq)) { flag = 2; break; }
      const double alpha = rz / pq;
      // x += alpha p ; r -= alpha q ; z = Minv r   (rows owned by this warp)
      double acc_rz2 = 0.0, acc_rr = 0.0;
      for (int a = gw; a < A.n; a += nw) {
        double rv = 0.0;
        if (lane < BS) {
          rv = A.r[(size_t)a * BS + lane] - alpha * A.q[(size_t)a * BS + lane];
          A.x[(size_t)a * BS + lane] += alpha * A.p[(size_t)a * BS + lane];
          A.r[(size_t)a * BS + lane] = rv;
        }
        double r6[BS];
#pragma unroll
        for (int k = 0; k < BS; k++) r6[k] = __shfl_sync(0xffffffffu, rv, k);
        if (lane < BS) {
          const double* M = A.Minv + (size_t)a * BB + lane * BS;
          double zv = 0.0;
#pragma unroll
          for (int k = 0; k < BS; k++) zv += M[k] * r6[k];
          A.z[(size_t)a * BS + lane] = zv;
          acc_rz2 += rv * zv;
          acc_rr += rv * rv;
        }
      }
      {
        const double t0 = block_sum(acc_rz2, red);
        const double t1 = block_sum(acc_rr, red);
        if (threadIdx.x == 0) { part1[blockIdx.x] = t0; part2[blockIdx.x] = t1; }
      }
      grid_barrier(A.bar, target);
      const double rz_new = sum_partials_dev(part1, G);
      rr = sum_partials_dev(part2, G);
      if (rr <= stop2) { flag = 0; it++; break; }
      const double beta = rz_new / rz;
      rz = rz_new;
      for (int a = gw; a < A.n; a += nw)
        if (lane < BS) A.p[(size_t)a * BS + lane] = A.z[(size_t)a * BS + lane] + beta * A.p[(size_t)a * BS + lane];
      grid_barrier(A.bar, target);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    A.status[0] = (double)it;
    A.status[1] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    A.status[2] = (double)flag;
  }
}

}  // namespace ccm

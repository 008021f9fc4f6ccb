// pgo.cu — Sim3 essential-graph optimisation on the GPU behind ccm_pgo_solve (include/ccm_b200.h).
//
// Replaces optimizer.optimize(20) of Optimizer::OptimizeEssentialGraph{LoopClosure,MapFusion} (S/Optimizer.cpp:1277, :1513):
//   vertex  VertexSim3Expmap, oplus = Sim3(update) * estimate, update[6] := 0 when _fix_scale  (G/types/types_seven_dof_expmap.h:60-69)
//   edge    EdgeSim3, e = log(C * S_i * S_j^-1), information = I7                               (:105-114, S/Optimizer.cpp:1125)
//   solver  BlockSolver_7_3 without Schur, Levenberg with user lambda 1e-16                     (S/Optimizer.cpp:1066-1072)
// The reference differentiates numerically (central differences, delta = 1e-9, G/core/base_binary_edge.hpp:131-205);
// k_pgo_linearize does the same per edge (one thread per edge, 28 error evaluations) so that the Jacobians carry the same
// discretisation, then scatters J^T J / J^T e into the 7x7 block-CSR Hessian with red.add.f64.  The linear solve is the
// shared persistent PCG (pcg.cuh, BS = 7) instead of the reference's sparse LDL^T.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>

#include "ba_math.cuh"
#include "common.cuh"
#include "pcg.cuh"
#include "sim3_math.cuh"

using namespace ccm;

namespace {

struct PgoEdge { int i, j, ai, aj, idx_ii, idx_jj, idx_ij; int ij_transposed; };  // ai/aj: free index or -1

__global__ void __launch_bounds__(128) k_pgo_error(const double* __restrict__ v, const PgoEdge* __restrict__ edges,
                                                   const double* __restrict__ meas, int E, double* __restrict__ partials) {
  __shared__ double red[4];
  double acc = 0.0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    double err[7];
    edge_error(s3_load(meas + 8 * (size_t)e), s3_load(v + 8 * (size_t)edges[e].i), s3_load(v + 8 * (size_t)edges[e].j), err);
#pragma unroll
    for (int k = 0; k < 7; k++) acc += err[k] * err[k];
  }
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// one thread per edge: error, numeric Jacobians (central differences through oplus), scatter into H and b
__global__ void __launch_bounds__(64) k_pgo_linearize(const double* __restrict__ v, const PgoEdge* __restrict__ edges,
                                                      const double* __restrict__ meas, int E, int fix_scale,
                                                      double* __restrict__ H, double* __restrict__ b) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const PgoEdge ed = edges[e];
  const S3 C = s3_load(meas + 8 * (size_t)e), vi = s3_load(v + 8 * (size_t)ed.i), vj = s3_load(v + 8 * (size_t)ed.j);
  double err[7];
  edge_error(C, vi, vj, err);
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  double Ji[49], Jj[49];
  for (int side = 0; side < 2; side++) {
    double* J = side ? Jj : Ji;
    const bool free_v = (side ? ed.aj : ed.ai) >= 0;
    for (int d = 0; d < 7; d++) {
      double ep[7], em[7], add[7] = {0, 0, 0, 0, 0, 0, 0};
      if (free_v) {
        add[d] = delta;
        if (side) edge_error(C, vi, s3_oplus(vj, add, fix_scale), ep); else edge_error(C, s3_oplus(vi, add, fix_scale), vj, ep);
        add[d] = -delta;
        if (side) edge_error(C, vi, s3_oplus(vj, add, fix_scale), em); else edge_error(C, s3_oplus(vi, add, fix_scale), vj, em);
      }
      for (int r = 0; r < 7; r++) J[r * 7 + d] = free_v ? scalar * (ep[r] - em[r]) : 0.0;
    }
  }
  if (ed.ai >= 0) {
    for (int r = 0; r < 7; r++) {
      double g = 0;
      for (int k = 0; k < 7; k++) g -= Ji[k * 7 + r] * err[k];
      atomicAdd(b + (size_t)ed.ai * 7 + r, g);
      for (int c = 0; c < 7; c++) {
        double hh = 0;
        for (int k = 0; k < 7; k++) hh += Ji[k * 7 + r] * Ji[k * 7 + c];
        atomicAdd(H + (size_t)ed.idx_ii * 49 + r * 7 + c, hh);
      }
    }
  }
  if (ed.aj >= 0) {
    for (int r = 0; r < 7; r++) {
      double g = 0;
      for (int k = 0; k < 7; k++) g -= Jj[k * 7 + r] * err[k];
      atomicAdd(b + (size_t)ed.aj * 7 + r, g);
      for (int c = 0; c < 7; c++) {
        double hh = 0;
        for (int k = 0; k < 7; k++) hh += Jj[k * 7 + r] * Jj[k * 7 + c];
        atomicAdd(H + (size_t)ed.idx_jj * 49 + r * 7 + c, hh);
      }
    }
  }
  if (ed.ai >= 0 && ed.aj >= 0 && ed.ai != ed.aj) {
    // block (ai, aj) += Ji^T Jj and its mirror (aj, ai) += Jj^T Ji: the PCG works on the full symmetric block-CSR
    for (int r = 0; r < 7; r++)
      for (int c = 0; c < 7; c++) {
        double hh = 0;
        for (int k = 0; k < 7; k++) hh += Ji[k * 7 + r] * Jj[k * 7 + c];
        atomicAdd(H + (size_t)ed.idx_ij * 49 + r * 7 + c, hh);
        atomicAdd(H + (size_t)ed.ij_transposed * 49 + c * 7 + r, hh);
      }
  }
}

// per free vertex: Hd = H + lambda I on the diagonal block (written to Hs), Minv = inverse of that block (Gauss-Jordan)
__global__ void __launch_bounds__(64) k_pgo_damp(const double* __restrict__ H, double* __restrict__ Hs, long long nnz49,
                                                 const int* __restrict__ diag, int n, double lambda, double* __restrict__ Minv,
                                                 double* __restrict__ maxdiag_bits, int* __restrict__ fail) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long k = i; k < nnz49; k += (long long)gridDim.x * blockDim.x) Hs[k] = H[k];
  (void)maxdiag_bits;
  if (i >= n) return;
  double A[49], I[49];
  const double* src = H + (size_t)diag[i] * 49;
  for (int k = 0; k < 49; k++) { A[k] = src[k]; I[k] = 0.0; }
  for (int k = 0; k < 7; k++) { A[k * 8] += lambda; I[k * 8] = 1.0; }
  bool ok = true;
  for (int c = 0; c < 7; c++) {
    int piv = c;
    double best = fabs(A[c * 7 + c]);
    for (int r = c + 1; r < 7; r++)
      if (fabs(A[r * 7 + c]) > best) { best = fabs(A[r * 7 + c]); piv = r; }
    if (!(best > 0.0)) { ok = false; break; }
    if (piv != c)
      for (int k = 0; k < 7; k++) {
        double t = A[c * 7 + k]; A[c * 7 + k] = A[piv * 7 + k]; A[piv * 7 + k] = t;
        t = I[c * 7 + k]; I[c * 7 + k] = I[piv * 7 + k]; I[piv * 7 + k] = t;
      }
    const double ip = 1.0 / A[c * 7 + c];
    for (int k = 0; k < 7; k++) { A[c * 7 + k] *= ip; I[c * 7 + k] *= ip; }
    for (int r = 0; r < 7; r++) {
      if (r == c) continue;
      const double f = A[r * 7 + c];
      for (int k = 0; k < 7; k++) { A[r * 7 + k] -= f * A[c * 7 + k]; I[r * 7 + k] -= f * I[c * 7 + k]; }
    }
  }
  for (int k = 0; k < 49; k++) Minv[(size_t)i * 49 + k] = ok ? I[k] : (k % 8 == 0 ? 1.0 : 0.0);
  if (!ok) atomicExch(fail, 1);
}

__global__ void k_pgo_add_lambda(double* __restrict__ Hs, const int* __restrict__ diag, int n, double lambda) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 7) return;
  Hs[(size_t)diag[i / 7] * 49 + (i % 7) * 8] += lambda;
}

__global__ void __launch_bounds__(128) k_pgo_update(const double* __restrict__ v, const int* __restrict__ vidx,
                                                    const double* __restrict__ x, const double* __restrict__ b, int K,
                                                    int fix_scale, double lambda, double* __restrict__ vt,
                                                    double* __restrict__ partials) {
  __shared__ double red[4];
  double acc = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
    S3 s = s3_load(v + 8 * (size_t)k);
    const int a = vidx[k];
    if (a >= 0) {
      double u[7];
      for (int d = 0; d < 7; d++) u[d] = x[(size_t)a * 7 + d];
      if (fix_scale) u[6] = 0;  // oplusImpl zeroes the solver's x[6] in place before computeScale reads it
      for (int d = 0; d < 7; d++) acc += u[d] * (lambda * u[d] + b[(size_t)a * 7 + d]);
      s = s3_oplus(s, u, fix_scale);
    }
    s3_store(s, vt + 8 * (size_t)k);
  }
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(1024) k_sum(const double* __restrict__ p, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += p[i];
  const double t = block_sum(v, red);
  if (threadIdx.x == 0) out[0] = t;
}

__global__ void k_pgo_maxdiag(const double* __restrict__ H, const int* __restrict__ diag, int n, unsigned long long* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 7) return;
  const double v = fabs(H[(size_t)diag[i / 7] * 49 + (i % 7) * 8]);
  atomicMax(out, (unsigned long long)__double_as_longlong(v));
}

void pgo_solve(const ccm_pgo_problem* p, const ccm_pgo_options* o, ccm_pgo_result* r) {
  const auto T0 = std::chrono::steady_clock::now();
  ensure_device();
  CCM_REQUIRE(p && o && r && p->K > 0 && p->E >= 0 && p->sim3 && p->fixed && r->sim3, "ccm_pgo_solve: bad argument");
  cudaStream_t s;
  CCM_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{s};
  const int K = p->K, E = p->E;
  // active set (edges with at least one free vertex) and index mapping
  std::vector<char> has(K, 0);
  std::vector<int> act;
  for (int e = 0; e < E; e++) {
    const int i = p->edge_i[e], j = p->edge_j[e];
    CCM_REQUIRE(i >= 0 && i < K && j >= 0 && j < K, "ccm_pgo_solve: edge index out of range");
    if (p->fixed[i] && p->fixed[j]) continue;
    act.push_back(e);
    has[i] = has[j] = 1;
  }
  std::vector<int> vidx(K, -1), idxv;
  for (int k = 0; k < K; k++)
    if (has[k] && !p->fixed[k]) { vidx[k] = (int)idxv.size(); idxv.push_back(k); }
  const int n = (int)idxv.size(), Ea = (int)act.size();
  r->trace_len = 0; r->iters_done = 0; r->chi2_initial = r->chi2_final = 0; r->lambda_final = 0;
  memcpy(r->sim3, p->sim3, sizeof(double) * 8 * K);
  if (n == 0 || Ea == 0) { r->iters_done = -1; return; }
  // full symmetric block-CSR pattern
  std::vector<std::vector<int>> rows(n);
  for (int a = 0; a < n; a++) rows[a].push_back(a);
  for (int e : act) {
    const int a = vidx[p->edge_i[e]], b = vidx[p->edge_j[e]];
    if (a >= 0 && b >= 0 && a != b) { rows[a].push_back(b); rows[b].push_back(a); }
  }
  std::vector<int> rowptr(n + 1, 0), col, diag(n);
  for (int a = 0; a < n; a++) {
    std::sort(rows[a].begin(), rows[a].end());
    rows[a].erase(std::unique(rows[a].begin(), rows[a].end()), rows[a].end());
    col.insert(col.end(), rows[a].begin(), rows[a].end());
    rowptr[a + 1] = (int)col.size();
  }
  auto find = [&](int a, int b) {
    const int* bb = col.data() + rowptr[a];
    const int* ee = col.data() + rowptr[a + 1];
    return (int)(std::lower_bound(bb, ee, b) - col.data());
  };
  for (int a = 0; a < n; a++) diag[a] = find(a, a);
  std::vector<PgoEdge> edges(Ea);
  std::vector<double> meas((size_t)Ea * 8);
  for (int k = 0; k < Ea; k++) {
    const int e = act[k];
    PgoEdge& ed = edges[k];
    ed.i = p->edge_i[e]; ed.j = p->edge_j[e]; ed.ai = vidx[ed.i]; ed.aj = vidx[ed.j];
    ed.idx_ii = ed.ai >= 0 ? diag[ed.ai] : 0; ed.idx_jj = ed.aj >= 0 ? diag[ed.aj] : 0;
    ed.idx_ij = (ed.ai >= 0 && ed.aj >= 0 && ed.ai != ed.aj) ? find(ed.ai, ed.aj) : 0;
    ed.ij_transposed = (ed.ai >= 0 && ed.aj >= 0 && ed.ai != ed.aj) ? find(ed.aj, ed.ai) : 0;
    memcpy(&meas[(size_t)k * 8], p->meas + 8 * (size_t)e, 8 * sizeof(double));
  }
  const long long nnzb = (long long)col.size();
  DevBuf<double> d_v, d_vt, d_meas, H, Hs, b, Minv, x, pr, pz, pp, pq, partials, scal, pcg_partials, pcg_status;
  DevBuf<int> d_vidx, d_rowptr, d_col, d_diag, fail;
  DevBuf<PgoEdge> d_edges;
  DevBuf<unsigned> bar;
  d_v.upload(p->sim3, (size_t)K * 8, s); d_vt.alloc((size_t)K * 8);
  d_meas.upload(meas.data(), meas.size(), s); d_edges.upload(edges.data(), Ea, s);
  d_vidx.upload(vidx.data(), K, s); d_rowptr.upload(rowptr.data(), n + 1, s); d_col.upload(col.data(), nnzb, s);
  d_diag.upload(diag.data(), n, s);
  H.alloc(nnzb * 49); Hs.alloc(nnzb * 49); b.alloc((size_t)n * 7); Minv.alloc((size_t)n * 49);
  x.alloc_zero((size_t)n * 7, s); pr.alloc((size_t)n * 7); pz.alloc((size_t)n * 7); pp.alloc((size_t)2 * n * 7); pq.alloc((size_t)n * 7);
  const int gsm = sm_count();
  partials.alloc((size_t)gsm * 8 + 8); scal.alloc_zero(8, s); pcg_status.alloc_zero(4, s); bar.alloc_zero(2, s); fail.alloc_zero(1, s);
  const int pcg_block = ((long long)n * 32 >= (long long)gsm * PCG_TPB) ? 512 : 256;
  void* pcg_fn = pcg_block == 512 ? (void*)k_pcg<7, 512, 1> : (void*)k_pcg<7, 256, 2>;
  int per_sm = 0;
  CCM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)pcg_fn, pcg_block, 0));
  CCM_REQUIRE(per_sm >= 1, "k_pcg does not fit on an SM");
  per_sm = std::min(per_sm, pcg_block == 256 ? 2 : 1);
  const int pcg_grid = std::max(1, std::min(gsm * per_sm, div_up((long long)n * 32, pcg_block)));
  pcg_partials.alloc((size_t)3 * pcg_grid);
  int c_agg = 0, c_nc = 0;
  {
    const char* e = getenv("CCM_PCG_NC");
    pcg_coarse_shape(n, e ? atoi(e) : 64, &c_agg, &c_nc);
  }
  DevBuf<double> cAc, crc, cyc;
  {
    const size_t nC = (size_t)7 * c_nc;
    cAc.alloc(std::max(2 * nC * nC, (size_t)1)); crc.alloc(std::max(2 * nC, (size_t)1)); cyc.alloc(std::max(nC, (size_t)1));
  }
  double* h_scal = nullptr;
  CCM_CUDA(cudaMallocHost((void**)&h_scal, 8 * sizeof(double)));
  struct HostGuard { double* p; ~HostGuard() { cudaFreeHost(p); } } hg{h_scal};
  double *cur = d_v.p, *trial = d_vt.p;
  const int pcg_max = o->pcg_max_iter > 0 ? o->pcg_max_iter : 5000;
  const double pcg_tol = o->pcg_tol > 0 ? o->pcg_tol : 1e-10;
  auto terminate = [&] { return o->stop && *o->stop; };
  auto chi2_of = [&](const double* state, double* dev_out) {
    const int g = std::max(1, std::min(div_up(Ea, 128), gsm * 8));
    k_pgo_error<<<g, 128, 0, s>>>(state, d_edges.p, d_meas.p, Ea, partials.p);
    CCM_LAUNCHED();
    k_sum<<<1, 1024, 0, s>>>(partials.p, g, dev_out);
    CCM_LAUNCHED();
  };
  double lambda = -1, ni = 2;
  int nBad = 0, ret = 0;
  bool ok = true;
  for (int it = 0; it < o->iterations && !terminate() && ok; it++) {
    chi2_of(cur, scal.p);
    CCM_CUDA(cudaMemsetAsync(H.p, 0, H.bytes(), s));
    CCM_CUDA(cudaMemsetAsync(b.p, 0, b.bytes(), s));
    k_pgo_linearize<<<div_up(Ea, 64), 64, 0, s>>>(cur, d_edges.p, d_meas.p, Ea, p->fix_scale, H.p, b.p);
    CCM_LAUNCHED();
    if (it == 0 && !(o->lambda_init > 0)) {
      CCM_CUDA(cudaMemsetAsync(scal.p + 4, 0, sizeof(double), s));
      k_pgo_maxdiag<<<div_up(n * 7, 128), 128, 0, s>>>(H.p, d_diag.p, n, reinterpret_cast<unsigned long long*>(scal.p + 4));
      CCM_LAUNCHED();
    }
    CCM_CUDA(cudaMemcpyAsync(h_scal, scal.p, 5 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CCM_CUDA(cudaStreamSynchronize(s));
    double currentChi = h_scal[0];
    const double iniChi = currentChi;
    if (it == 0) {
      r->chi2_initial = currentChi;
      lambda = o->lambda_init > 0 ? o->lambda_init : 1e-5 * h_scal[4];
      ni = 2; nBad = 0;
    }
    double rho = 0, tempChi = currentChi, lambda_used = lambda, relres = 0;
    int qmax = 0, pcg_it = 0;
    do {
      lambda_used = lambda;
      CCM_CUDA(cudaMemsetAsync(fail.p, 0, sizeof(int), s));
      k_pgo_damp<<<std::max(div_up(n, 64), std::min(div_up(nnzb * 49, 64), gsm * 16)), 64, 0, s>>>(
          H.p, Hs.p, nnzb * 49, d_diag.p, n, lambda, Minv.p, nullptr, fail.p);
      CCM_LAUNCHED();
      k_pgo_add_lambda<<<div_up(n * 7, 128), 128, 0, s>>>(Hs.p, d_diag.p, n, lambda);
      CCM_LAUNCHED();
      CCM_CUDA(cudaMemsetAsync(bar.p, 0, 2 * sizeof(unsigned), s));
      PcgArgs a;
      a.n = n; a.rowptr = d_rowptr.p; a.col = d_col.p; a.val = Hs.p; a.Minv = Minv.p; a.b = b.p;
      a.x = x.p; a.r = pr.p; a.z = pz.p; a.p = pp.p; a.q = pq.p; a.partials = pcg_partials.p; a.bar = bar.p;
      a.tol = pcg_tol; a.max_iter = pcg_max; a.status = pcg_status.p;
      a.agg = c_agg; a.nc = c_nc; a.Ac = cAc.p; a.rc = crc.p; a.yc = cyc.p; a.coarse_mode = 1; a.prof = nullptr;
      void* args[] = {&a};
      CCM_CUDA(cudaLaunchCooperativeKernel(pcg_fn, dim3(pcg_grid), dim3(pcg_block), args, 0, s));
      CCM_LAUNCHED();
      const int g = std::max(1, std::min(div_up(K, 128), gsm * 8));
      k_pgo_update<<<g, 128, 0, s>>>(cur, d_vidx.p, x.p, b.p, K, p->fix_scale, lambda, trial, partials.p);
      CCM_LAUNCHED();
      k_sum<<<1, 1024, 0, s>>>(partials.p, g, scal.p + 1);
      CCM_LAUNCHED();
      chi2_of(trial, scal.p + 2);
      CCM_CUDA(cudaMemcpyAsync(h_scal, scal.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
      CCM_CUDA(cudaMemcpyAsync(h_scal + 3, pcg_status.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
      CCM_CUDA(cudaMemcpyAsync(h_scal + 6, fail.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      CCM_CUDA(cudaStreamSynchronize(s));
      int jfail;
      memcpy(&jfail, h_scal + 6, sizeof(int));
      tempChi = h_scal[2];
      pcg_it = (int)h_scal[3]; relres = h_scal[4];
      const bool ok2 = !((int)h_scal[5] == 2 || jfail);
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = currentChi - tempChi;
      double scale = h_scal[1];
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        lambda *= std::max(1. / 3., alpha);
        ni = 2;
        currentChi = tempChi;
        std::swap(cur, trial);
      } else {
        lambda *= ni;
        ni *= 2;
      }
      qmax++;
    } while (rho < 0 && qmax < 10 && !terminate());
    ret++;
    if (r->trace && r->trace_len < r->trace_cap) {
      double* tr = r->trace + (size_t)r->trace_len * CCM_TRACE_COLS;
      tr[0] = it; tr[1] = lambda_used; tr[2] = currentChi; tr[3] = rho; tr[4] = qmax; tr[5] = lambda; tr[6] = pcg_it; tr[7] = relres;
      r->trace_len++;
    }
    r->chi2_final = currentChi; r->lambda_final = lambda;
    if (qmax == 10 || rho == 0) { ok = false; continue; }
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
    if (nBad >= 3) { ok = false; continue; }
  }
  r->iters_done = ret;
  CCM_CUDA(cudaMemcpyAsync(r->sim3, cur, sizeof(double) * 8 * K, cudaMemcpyDeviceToHost, s));
  CCM_CUDA(cudaStreamSynchronize(s));
  r->t_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
}

}  // namespace

extern "C" int ccm_pgo_solve(const ccm_pgo_problem* p, const ccm_pgo_options* o, ccm_pgo_result* r) {
  return guarded([&] { pgo_solve(p, o, r); });
}

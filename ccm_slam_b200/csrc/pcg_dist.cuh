// pcg_dist.cuh — the PCG of pcg.cuh with its block rows distributed over the ranks of one NVLink/NVSwitch node.
//
// STATUS: written after the round-1 GPU budget was spent; it has never run on a device.  It is reachable only with
// CCM_PCG_DIST=1 (with one rank it talks to its own window: a single-GPU dry run of everything but the NVLink hop); the default
// multi-rank path is the replicated k_pcg.
//
// Why: with landmarks sharded everything but the reduced-camera solve scales with the number of GPUs; the
// replicated solve is 62 % of the single-GPU step and caps the 1 -> 8 speed-up at 1.6x.  Distributing the solve through NCCL
// would put three collectives per iteration behind host launches (~3 x 10 us against a 9 us SpMV share at N = 8), so the
// exchange is done by the persistent kernel itself: every rank maps every peer's exchange window (cudaIpc) and
//   * stores its slice of z, its partial dot products and its partial restricted residual straight into every peer's window
//     (plain st.global over NVLink, then __threadfence_system),
//   * meets the peers at a flag barrier in peer memory (one 64-bit epoch word per source rank),
//   * and sums the partials of all ranks in rank order, so that every rank holds bit-identical scalars, z and p and takes
//     the same branches (same iteration count, same exit) without any broadcast.
// Row ownership: rank k owns block rows [r0, r1) (balanced by block count).  S (all rows), Minv and b are present on every rank
// (the all-reduce of the Schur blocks leaves them there); x, r, q are touched on own rows only; z lives in the window
// (double buffered by iteration parity), p is recomputed for all rows by every rank from z (60 k fmas).
// The coarse level: the Galerkin matrix and its inverse are built by the replicated k_pcg (launched with max_iter = 0) and
// reused here (coarse_mode 2); each rank restricts its own rows, the partials are exchanged and summed, the dense coarse solve
// is replicated (10.6 MB with a linear P over 192 nodes).
// Hazards: a window region written in iteration i is next written in iteration i + 2; a rank can only get there through two
// barriers that need this rank's arrival after it finished reading iteration i.  Spins carry a clock64 time-out that raises a
// node-visible abort word instead of hanging the GPU.
#pragma once
#include "pcg.cuh"

namespace ccm {

struct PcgDistArgs {
  PcgArgs a;                 // as for k_pcg; a.z is unused (z lives in the window), a.coarse_mode must be 2 or agg <= 0
  int rank, nranks;
  int r0, r1;                // own block rows
  char* const* win;          // [nranks] base address of every rank's exchange window as mapped in THIS process
  size_t off_z, off_scal, off_rc, off_x, off_flags, off_ctl;   // byte offsets inside a window
  unsigned long long epoch0; // first barrier epoch of this launch (flags only ever grow)
  unsigned long long* rel;   // local release word of the combined barrier
  long long timeout_cycles;
};

// local arrive -> (CTA 0: signal peers, wait for peers) -> local release.  Returns false when the node aborted.
__device__ __forceinline__ bool xbarrier(const PcgDistArgs& D, unsigned& target, unsigned long long& epoch) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // this CTA's stores (local and remote) are ordered before its arrival
    target += gridDim.x;
    epoch += 1;
    volatile unsigned* ctl = reinterpret_cast<volatile unsigned*>(D.win[D.rank] + D.off_ctl);
    const long long t0 = clock64();
    int ok = 1;
    atomicAdd(D.a.bar, 1u);
    if (blockIdx.x == 0) {
      while (*(volatile unsigned*)D.a.bar < target) {
        if (*ctl || clock64() - t0 > D.timeout_cycles) { ok = 0; break; }
      }
      __threadfence_system();
      if (ok) {
        for (int k = 0; k < D.nranks; k++)
          reinterpret_cast<volatile unsigned long long*>(D.win[k] + D.off_flags)[D.rank] = epoch;
        volatile unsigned long long* mine = reinterpret_cast<volatile unsigned long long*>(D.win[D.rank] + D.off_flags);
        for (int k = 0; k < D.nranks && ok; k++)
          while (mine[k] < epoch) {
            if (*ctl || clock64() - t0 > D.timeout_cycles) { ok = 0; break; }
          }
      }
      if (!ok)  // tell every rank (and every local CTA) to leave
        for (int k = 0; k < D.nranks; k++) *reinterpret_cast<volatile unsigned*>(D.win[k] + D.off_ctl) = 1u;
      __threadfence_system();
      *(volatile unsigned long long*)D.rel = epoch;
    } else {
      while (*(volatile unsigned long long*)D.rel < epoch) {
        if (*ctl || clock64() - t0 > D.timeout_cycles) { ok = 0; break; }
      }
    }
    if (*ctl) ok = 0;
    __threadfence_system();
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

template <int BS, int MAXT = PCG_TPB, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB) k_pcg_dist(PcgDistArgs D) {
  constexpr int BB = BS * BS;
  const PcgArgs& A = D.a;
  __shared__ double red[PCG_TPB / 32];
  const int bdim = blockDim.x;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * bdim + threadIdx.x) >> 5;
  const int nw = (gridDim.x * bdim) >> 5;
  const int G = gridDim.x;
  const int N = D.nranks, me = D.rank;
  unsigned target = 0;
  unsigned long long epoch = D.epoch0;
  double* part0 = A.partials;
  double* part1 = A.partials + G;
  double* part2 = A.partials + 2 * G;
  const int nC = A.agg > 0 ? BS * A.nc : 0;
  // status[3] is what the set-up launch (k_pcg, max_iter 0) or the previous solve left: > 0 iff Ac holds a usable inverse here.
  // The ranks build their inverses independently, so they agree on using the coarse level only after the exchange below.
  const double coarse_here = (nC > 0 && A.status[3] > 0.0) ? 1.0 : 0.0;
  bool coarse = false;
  const double* Ainv = nC > 0 ? ((((nC + GJB - 1) / GJB) & 1) ? A.Ac + (size_t)nC * nC : A.Ac) : nullptr;
  const size_t nv = (size_t)A.n * BS;
  auto zbuf = [&](int k, int par) { return reinterpret_cast<double*>(D.win[k] + D.off_z) + (size_t)par * nv; };
  auto scal = [&](int k, int par, int src) { return reinterpret_cast<double*>(D.win[k] + D.off_scal) + ((size_t)par * N + src) * 4; };
  auto rcpart = [&](int k, int par, int src) { return reinterpret_cast<double*>(D.win[k] + D.off_rc) + ((size_t)par * N + src) * nC; };
  int status_flag = 1;
  bool alive = true;

  // sum over ranks, rank order, of slot `which` of this rank's copy of the scalars (identical on every rank)
  auto sum_ranks = [&](int par, int which) {
    double v = 0.0;
    for (int k = 0; k < N; k++) v += __ldcg(scal(me, par, k) + which);
    return v;
  };
  // Scalar slots of one parity: [0] p.q  [1] r.z  [2] r.r  [3] set-up values (b.b / the coarse flag).  Threads 0..N-1 of CTA 0 each
  // sum the per-CTA partials in index order (the same value in each of them) and store it into one rank's window.
  auto publish = [&](int par, int slot, const double* partials_or_null, double direct) {
    if (blockIdx.x == 0 && threadIdx.x < N) {
      double v = direct;
      if (partials_or_null) {
        v = 0.0;
        for (int i = 0; i < G; i++) v += __ldcg(partials_or_null + i);
      }
      scal(threadIdx.x, par, me)[slot] = v;
    }
  };
  // restriction of own rows into the local buffer rc[buf] (red.add), the same arithmetic as k_pcg's
  auto restrict_own = [&](int buf) {
    for (int a = D.r0 + gw; a < D.r1; a += nw)
      if (lane < BS) {
        const CoarseParents pa = coarse_parents(a, A.agg, A.nc, A.prolong);
        const double rv = A.r[(size_t)a * BS + lane];
        atomicAdd(A.rc + (size_t)buf * nC + (size_t)pa.lo * BS + lane, pa.w0 * rv);
        if (pa.w1 != 0.0) atomicAdd(A.rc + (size_t)buf * nC + (size_t)pa.hi * BS + lane, pa.w1 * rv);
      }
  };
  // coarse correction for the current residual: local restrict -> exchange partials -> rank-ordered sum -> replicated dense solve
  auto coarse_correct = [&](int par, int buf) -> bool {
    restrict_own(buf);
    grid_barrier(A.bar, target);
    for (long long i = (long long)blockIdx.x * bdim + threadIdx.x; i < (long long)nC * N; i += (long long)G * bdim) {
      const int k = (int)(i / nC), j = (int)(i % nC);
      rcpart(k, par, me)[j] = __ldcg(A.rc + (size_t)buf * nC + j);
    }
    if (!xbarrier(D, target, epoch)) return false;
    for (long long j = (long long)blockIdx.x * bdim + threadIdx.x; j < nC; j += (long long)G * bdim) {
      double s = 0.0;
      for (int k = 0; k < N; k++) s += __ldcg(rcpart(me, par, k) + j);
      A.rc[(size_t)buf * nC + j] = s;             // own partial replaced by the node-wide sum
      A.rc[(size_t)(buf ^ 1) * nC + j] = 0.0;     // the other buffer is idle in this phase: clear it for the next restriction
    }
    grid_barrier(A.bar, target);
    const double* rcv = A.rc + (size_t)buf * nC;
    for (int i = gw; i < nC; i += nw) {
      double s = 0.0;
      for (int j = lane; j < nC; j += 32) s += __ldcg(Ainv + (size_t)i * nC + j) * __ldcg(rcv + j);
      s = warp_sum(s);
      if (lane == 0) A.yc[i] = s;
    }
    grid_barrier(A.bar, target);
    return true;
  };
  // z = Minv r (+ P yc) on own rows, stored into every rank's z buffer `par`; accumulates r.z and r.r
  auto precond_own = [&](int par, double& acc_rz, double& acc_rr) {
    for (int a = D.r0 + gw; a < D.r1; a += nw) {
      double rv = 0.0;
      if (lane < BS) rv = A.r[(size_t)a * BS + lane];
      double r6[BS];
#pragma unroll
      for (int k = 0; k < BS; k++) r6[k] = __shfl_sync(0xffffffffu, rv, k);
      if (lane < BS) {
        const double* M = A.Minv + (size_t)a * BB + lane * BS;
        double zv = 0.0;
#pragma unroll
        for (int k = 0; k < BS; k++) zv += M[k] * r6[k];
        if (coarse) {
          const CoarseParents pa = coarse_parents(a, A.agg, A.nc, A.prolong);
          zv += pa.w0 * __ldcg(A.yc + (size_t)pa.lo * BS + lane);
          if (pa.w1 != 0.0) zv += pa.w1 * __ldcg(A.yc + (size_t)pa.hi * BS + lane);
        }
        for (int k = 0; k < N; k++) zbuf(k, par)[(size_t)a * BS + lane] = zv;
        acc_rz += rv * zv;
        acc_rr += rv * rv;
      }
    }
  };

  // ---- x = 0, r = b on own rows; p_old = 0 everywhere -----------------------------------------------------------------
  for (int a = gw; a < A.n; a += nw)
    if (lane < BS) {
      A.p[(size_t)a * BS + lane] = 0.0;
      if (a >= D.r0 && a < D.r1) {
        A.x[(size_t)a * BS + lane] = 0.0;
        A.r[(size_t)a * BS + lane] = A.b[(size_t)a * BS + lane];
      }
    }
  for (long long i = (long long)blockIdx.x * bdim + threadIdx.x; i < 2ll * nC; i += (long long)G * bdim) A.rc[i] = 0.0;
  grid_barrier(A.bar, target);
  int par = 0, buf = 0;
  // the set-up scalars travel in the slots of parity 1, which iteration 0 (parity 0) does not write
  publish(1, 3, nullptr, coarse_here);
  alive = xbarrier(D, target, epoch);
  if (alive) {
    double all = 1.0;
    for (int k = 0; k < N; k++) all = fmin(all, __ldcg(scal(me, 1, k) + 3));
    coarse = all > 0.0;
  }
  if (alive && coarse) {
    alive = coarse_correct(par, buf);
    buf ^= 1;
  }
  double rz = 0.0, bb = 0.0, rr = 0.0;
  int it = 0;
  if (alive) {
    double acc_rz = 0.0, acc_bb = 0.0;
    precond_own(par, acc_rz, acc_bb);
    const double t0 = block_sum(acc_rz, red);
    const double t1 = block_sum(acc_bb, red);
    if (threadIdx.x == 0) { part0[blockIdx.x] = t0; part1[blockIdx.x] = t1; }
    grid_barrier(A.bar, target);
    publish(1, 1, part0, 0.0);
    publish(1, 2, part1, 0.0);
    alive = xbarrier(D, target, epoch);
  }
  if (alive) {
    rz = sum_ranks(1, 1);
    bb = sum_ranks(1, 2);
    rr = bb;
  }
  const double stop2 = A.tol * A.tol * bb;
  if (alive && !(bb > 0.0)) {
    status_flag = 0;
  } else if (alive) {
    double beta = 0.0;
    int pc = 0;
    for (it = 0; it < A.max_iter; it++, pc ^= 1, par ^= 1) {
      const double* pold = A.p + (size_t)pc * nv;
      double* pnew = A.p + (size_t)(pc ^ 1) * nv;
      const double* z = zbuf(me, par);
      // p = z + beta p_old for every row (each rank keeps the whole direction; readers below form the same values on the fly)
      for (int a = gw; a < A.n; a += nw)
        if (lane < BS) pnew[(size_t)a * BS + lane] = fma(beta, __ldcg(pold + (size_t)a * BS + lane), __ldcg(z + (size_t)a * BS + lane));
      // q = S p on own rows ; partial p.q
      double acc_pq = 0.0;
      for (int a = D.r0 + gw; a < D.r1; a += nw) {
        double y[BS];
#pragma unroll
        for (int k = 0; k < BS; k++) y[k] = 0.0;
        const int beg = A.rowptr[a], end = A.rowptr[a + 1];
        for (int j = beg + lane; j < end; j += 32) {
          const double* v = A.val + (size_t)j * BB;
          const size_t cj = (size_t)A.col[j] * BS;
          double pv[BS];
#pragma unroll
          for (int k = 0; k < BS; k++) pv[k] = fma(beta, __ldcg(pold + cj + k), __ldcg(z + cj + k));
#pragma unroll
          for (int rI = 0; rI < BS; rI++)
#pragma unroll
            for (int c = 0; c < BS; c++) y[rI] += __ldg(v + rI * BS + c) * pv[c];
        }
#pragma unroll
        for (int k = 0; k < BS; k++) y[k] = warp_sum(y[k]);
        if (lane < BS) {
          double yl = y[0];
#pragma unroll
          for (int k = 1; k < BS; k++) yl = lane == k ? y[k] : yl;
          const double pn = fma(beta, __ldcg(pold + (size_t)a * BS + lane), __ldcg(z + (size_t)a * BS + lane));
          A.q[(size_t)a * BS + lane] = yl;
          acc_pq += yl * pn;
        }
      }
      {
        const double t0 = block_sum(acc_pq, red);
        if (threadIdx.x == 0) part0[blockIdx.x] = t0;
      }
      grid_barrier(A.bar, target);
      publish(par, 0, part0, 0.0);
      if (!xbarrier(D, target, epoch)) { alive = false; break; }
      const double pq = sum_ranks(par, 0);
      if (!(pq > 0.0) || !isfinite(pq)) { status_flag = 2; break; }
      const double alpha = rz / pq;
      for (int a = D.r0 + gw; a < D.r1; a += nw)
        if (lane < BS) {
          A.x[(size_t)a * BS + lane] += alpha * pnew[(size_t)a * BS + lane];
          A.r[(size_t)a * BS + lane] -= alpha * A.q[(size_t)a * BS + lane];
        }
      if (coarse) {
        if (!coarse_correct(par, buf)) { alive = false; break; }
        buf ^= 1;
      }
      double acc_rz2 = 0.0, acc_rr = 0.0;
      precond_own(par ^ 1, acc_rz2, acc_rr);   // the next iteration reads z from the other parity
      {
        const double t0 = block_sum(acc_rz2, red);
        const double t1 = block_sum(acc_rr, red);
        if (threadIdx.x == 0) { part1[blockIdx.x] = t0; part2[blockIdx.x] = t1; }
      }
      grid_barrier(A.bar, target);
      publish(par, 1, part1, 0.0);
      publish(par, 2, part2, 0.0);
      if (!xbarrier(D, target, epoch)) { alive = false; break; }
      const double rz_new = sum_ranks(par, 1);
      rr = sum_ranks(par, 2);
      if (rr <= stop2) { status_flag = 0; it++; par ^= 1; break; }
      beta = rz_new / rz;
      rz = rz_new;
    }
  }
  // ---- gather x: own slice into every window, then the whole vector into the caller's x ---------------------------------
  if (alive) {
    for (int a = D.r0 + gw; a < D.r1; a += nw)
      if (lane < BS) {
        const double xv = A.x[(size_t)a * BS + lane];
        for (int k = 0; k < N; k++) reinterpret_cast<double*>(D.win[k] + D.off_x)[(size_t)a * BS + lane] = xv;
      }
    alive = xbarrier(D, target, epoch);
  }
  if (alive) {
    const double* xs = reinterpret_cast<const double*>(D.win[me] + D.off_x);
    for (long long i = (long long)blockIdx.x * bdim + threadIdx.x; i < (long long)nv; i += (long long)G * bdim) A.x[i] = __ldcg(xs + i);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    A.status[0] = (double)it;
    A.status[1] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    A.status[2] = alive ? (double)status_flag : 3.0;   // 3: a peer did not show up at a barrier (time-out / abort)
    A.status[3] = coarse ? (double)nC : 0.0;
  }
}

// byte layout of one exchange window
struct PcgDistLayout {
  size_t off_z, off_scal, off_rc, off_x, off_flags, off_ctl, bytes;
};
inline PcgDistLayout pcg_dist_layout(int n, int bs, int nranks, int nC) {
  PcgDistLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  const size_t nv = (size_t)n * bs;
  L.off_z = take(2 * nv * sizeof(double));
  L.off_scal = take((size_t)2 * nranks * 4 * sizeof(double));
  L.off_rc = take((size_t)2 * nranks * (nC > 0 ? nC : 1) * sizeof(double));
  L.off_x = take(nv * sizeof(double));
  L.off_flags = take((size_t)nranks * sizeof(unsigned long long));
  L.off_ctl = take(sizeof(unsigned));
  L.bytes = o;
  return L;
}

}  // namespace ccm

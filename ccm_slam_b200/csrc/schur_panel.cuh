// schur_panel.cuh — the Schur complement products of block_solver.hpp:370-439, landmark-synchronous: k_schur_panel.
//
// The list kernel (k_schur_mma, ba_kernels.cuh) walks one product list per upper block of S and gathers two 144-byte rows of Z per
// product from L2: 60 GB of L2 -> L1 traffic for 2.9 GB of Z on cfg5, every row fetched about 20 times.  It is bound by that gather
// (L1 wavefronts 65 %, L2 -> L1 6.8 TB/s), not by the 36 % busy DMMA pipe, and neither deeper prefetch, sorted lists nor L1 tiling
// moves it (profiles/r2/schur_tiled_cfg5.log).  This kernel turns the loop inside out:
//
//   * a CTA owns a PANEL of SP_R = 8 consecutive block rows, one warp per row a;
//   * it streams the landmarks of the panel's landmark range [min first landmark, max last landmark] of its eight poses through
//     shared memory: the observations of consecutive landmarks are consecutive rows of Z, so one stage (<= SP_SOBS rows, whole
//     landmarks) is ONE cp.async.bulk copy completing on an mbarrier, SP_NST stages deep;
//   * for a landmark its pose observes, warp a loads its own row once (A fragment) and meets every co-observing pose b in the band
//     a <= b < a + SP_DMAX with one f64 mma.sync.m8n8k4 whose B fragment comes from shared memory; the 6x6 block (a, b) lives in
//     the accumulator fragment number b - a: 2 x SP_DMAX doubles per lane, chosen by a warp-uniform switch, so every row of Z is
//     read from L2 once per panel that touches it (about 5 times) instead of once per product, and never gathered.
//
// Blocks outside the band (loop closures, merged maps) and panels whose landmark range is much wider than what their poses observe
// stay with the list kernel: `covered[u]` says which kernel owns upper block u, and product lists are only built for the others
// (and for the diagonal blocks, which also feed the pose pass).
#pragma once
#include <stdint.h>

#include "ba_kernels.cuh"
#include "pcg2.cuh"   // mbarrier / bulk-copy wrappers

namespace ccm {
namespace ba {

constexpr int SP_R = 8;         // block rows per panel
constexpr int SP_HALVES = 2;    // warps per block row: each keeps half of the band
constexpr int SP_DH = 22;       // block offsets per warp
constexpr int SP_DMAX = SP_HALVES * SP_DH;   // in-band block offsets kept in accumulator fragments (44)
constexpr int SP_SOBS = 160;    // observations (rows of Z) per stage: 23 040 bytes
constexpr int SP_NST = 3;       // stages
constexpr int SP_MAXL = 32;     // landmarks per stage (the boundary search is one warp wide)
constexpr int SP_THREADS = 32 * SP_R * SP_HALVES;
constexpr size_t SP_SMEM = (size_t)SP_NST * SP_SOBS * 18 * sizeof(double) + 256;

struct SchurPanelArgs {
  const double* Z;             // [El][18]
  const int* o_slot;           // [El] free-pose slot of every observation (-1: fixed pose)
  const int* lm_ptr;           // [Pl + 1]
  const double* gvec;          // [Pl][3]
  const int* pose_lmin; const int* pose_lmax;   // [Kf] first / last local landmark every free pose observes (lmin > lmax: none)
  const unsigned char* pan_on; // [panels]
  int Kf;
  const unsigned* bitmap; const int* word_prefix; const int* s_rowptr; const int* csr_u; int words;
  double* U_val; double* bneg;
};

#define SP_CASE(D) case D: dmma_884(c[D][0], c[D][1], av, bv); break;
#define SP_CASES4(D) SP_CASE(D) SP_CASE(D + 1) SP_CASE(D + 2) SP_CASE(D + 3)

__global__ void __launch_bounds__(SP_THREADS, 1) k_schur_panel(SchurPanelArgs A) {
  extern __shared__ __align__(128) unsigned char sp_smem[];
  double* Zs = reinterpret_cast<double*>(sp_smem);                        // [NST][SOBS * 18]
  __shared__ uint64_t mbar[SP_NST];
  __shared__ int d_la[SP_NST], d_nl[SP_NST], d_e0[SP_NST], d_direct[SP_NST];
  __shared__ int d_lp[SP_NST][SP_MAXL + 1];
  __shared__ int slot_s[SP_NST][SP_SOBS];           // free-pose slot of every staged observation
  __shared__ double g_s[SP_NST][SP_MAXL * 3];        // g_l of every staged landmark
  __shared__ int s_L0, s_L1;
  const int p = blockIdx.x;
  if (!A.pan_on[p]) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int half = wid & 1;
  const int a = p * SP_R + (wid >> 1);
  const int dlo = half * SP_DH;          // this warp keeps the blocks (a, a + dlo .. a + dlo + SP_DH - 1)
  const bool row_ok = a < A.Kf;
  const int m = lane >> 2, k = lane & 3;
  const bool ld = m < 6 && k < 3;
  const int off = ld ? m * 3 + k : 0;
  const bool gl = m == 6 && k < 3;       // lanes that carry g_l in column 6 of B for the diagonal block
  if (tid == 0) {
    int L0 = 0x7fffffff, L1 = -1;
    for (int r = 0; r < SP_R && p * SP_R + r < A.Kf; r++) {
      const int lo = A.pose_lmin[p * SP_R + r], hi = A.pose_lmax[p * SP_R + r];
      if (lo <= hi) { L0 = lo < L0 ? lo : L0; L1 = hi + 1 > L1 ? hi + 1 : L1; }
    }
    s_L0 = L0; s_L1 = L1;
    for (int s = 0; s < SP_NST; s++) p2_mbar_init(mbar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int L1 = s_L1;
  double c[SP_DH][2];
#pragma unroll
  for (int d = 0; d < SP_DH; d++) { c[d][0] = 0.0; c[d][1] = 0.0; }

  // warp 0 cuts the landmark range into stages of whole landmarks and issues their copies; `next_la` is the first landmark not yet staged
  int next_la = s_L0;
  auto stage_issue = [&](int kst) {   // warp 0 only; kst = running stage number
    const int slot = kst % SP_NST;
    int nl = 0, e0 = 0, ne = 0, direct = 0;
    const int la = next_la;
    if (la < L1) {
      e0 = __ldg(A.lm_ptr + la);
      const int li = la + 1 + lane;
      const int pe = li <= L1 ? __ldg(A.lm_ptr + li) : 0x7fffffff;
      const unsigned fits = __ballot_sync(0xffffffffu, li <= L1 && pe - e0 <= SP_SOBS);
      nl = fits == 0xffffffffu ? 32 : __ffs(~fits) - 1;   // fits is monotone: its run of low ones is the number of landmarks that fit
      if (nl == 0) { nl = 1; direct = 1; }   // one landmark with more observations than a stage holds: read from global memory
      const int prev = __shfl_up_sync(0xffffffffu, pe, 1);
      if (lane <= nl) d_lp[slot][lane] = lane == 0 ? e0 : prev;
      if (lane == 31 && nl == 32) d_lp[slot][32] = pe;
      ne = __shfl_sync(0xffffffffu, pe, nl - 1) - e0;
    }
    if (lane == 0) {
      d_la[slot] = la; d_nl[slot] = nl; d_e0[slot] = e0; d_direct[slot] = direct;
      if (nl > 0 && !direct) {
        const uint32_t bytes = (uint32_t)ne * 18u * (uint32_t)sizeof(double);
        p2_mbar_expect_tx(mbar + slot, bytes);
        p2_bulk_g2s(Zs + (size_t)slot * SP_SOBS * 18, A.Z + (size_t)e0 * 18, bytes, mbar + slot);
      } else if (nl > 0) {   // nothing to copy, but the slot's phase must still advance: the waiters count uses of the slot
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(p2_smem_u32(mbar + slot)) : "memory");
      }
    }
    next_la = la + nl;
  };
  // The pose slots and g_l of a stage travel one stage ahead too: every thread requests one element of stage k + 1 while stage k is
  // processed and parks it in shared memory at the end of the iteration (the barrier at the top of the next one publishes it).
  auto side_load = [&](int kst, int& sv, double& gv) {   // request
    const int slot = kst % SP_NST;
    sv = -2; gv = 0.0;
    if (d_nl[slot] > 0 && !d_direct[slot]) {
      const int e0 = d_e0[slot], ne = d_lp[slot][d_nl[slot]] - e0;
      if (tid < ne) sv = __ldg(A.o_slot + e0 + tid);
      const int t = tid - SP_SOBS;
      if (t >= 0 && t < 3 * d_nl[slot]) gv = __ldg(A.gvec + 3 * (size_t)d_la[slot] + t);
    }
  };
  auto side_store = [&](int kst, int sv, double gv) {
    const int slot = kst % SP_NST;
    if (tid < SP_SOBS) slot_s[slot][tid] = sv;
    const int t = tid - SP_SOBS;
    if (t >= 0 && t < 3 * SP_MAXL) g_s[slot][t] = gv;
  };
  static_assert(SP_SOBS + 3 * SP_MAXL <= SP_THREADS, "one thread per side element");
  if (wid == 0)
    for (int s = 0; s < SP_NST - 1; s++) stage_issue(s);
  __syncthreads();
  { int sv; double gv; side_load(0, sv, gv); side_store(0, sv, gv); }

  for (int kst = 0;; kst++) {
    __syncthreads();   // descriptors and side data of stage kst are visible; every warp is done with stage kst - 1, whose slot is refilled
    if (wid == 0) stage_issue(kst + SP_NST - 1);
    const int slot = kst % SP_NST;
    const int nl = d_nl[slot];
    if (nl == 0) break;   // the landmark range is exhausted (uniform: every warp reads the same descriptor)
    int sv_next; double gv_next;
    side_load(kst + 1, sv_next, gv_next);   // (the descriptor of stage kst + 1 was published at least one barrier ago)
    const int la = d_la[slot], e0 = d_e0[slot];
    const bool direct = d_direct[slot] != 0;
    {
      const uint32_t par = (uint32_t)(kst / SP_NST) & 1u;
      if (!p2_mbar_try_wait(mbar + slot, par)) {
        const long long t0 = clock64();
        while (!p2_mbar_try_wait(mbar + slot, par))
          if (clock64() - t0 > 4000000000ll) __trap();
      }
    }
    if (row_ok) {
      const double* rows = direct ? A.Z + (size_t)e0 * 18 : Zs + (size_t)slot * SP_SOBS * 18;
      const int* slots = direct ? A.o_slot + e0 : slot_s[slot];
      for (int li = 0; li < nl; li++) {
        const int s0 = d_lp[slot][li] - e0, n = d_lp[slot][li + 1] - d_lp[slot][li];
        for (int c0 = 0; c0 < n; c0 += 32) {           // own observations of this landmark (one, unless the input repeats a pose)
          const int sl = c0 + lane < n ? slots[s0 + c0 + lane] : -2;
          unsigned mine = __ballot_sync(0xffffffffu, sl == a);
          while (mine) {
            const int i = c0 + __ffs(mine) - 1;
            mine &= mine - 1;
            const double av = ld ? rows[(size_t)(s0 + i) * 18 + off] : 0.0;
            if (half == 0) {  // the diagonal product of this observation with itself; column 6 of B carries g_l: C[.][6] accumulates bneg
              double bv = ld ? av : 0.0;
              if (gl) bv = direct ? __ldg(A.gvec + 3 * (size_t)(la + li) + k) : g_s[slot][3 * li + k];
              dmma_884(c[0][0], c[0][1], av, bv);
            }
            for (int c1 = 0; c1 < n; c1 += 32) {       // co-observing poses in this warp's half of the band above a
              const int sb = c1 + lane < n ? slots[s0 + c1 + lane] : -2;
              const int dd = sb - a - dlo;              // offset inside the half: 0 .. SP_DH - 1 (0 is the diagonal for half 0: excluded)
              unsigned part = __ballot_sync(0xffffffffu, sb > a && dd >= 0 && dd < SP_DH);
              while (part) {
                const int j = __ffs(part) - 1;
                part &= part - 1;
                const int d = __shfl_sync(0xffffffffu, dd, j);
                const double bv = ld ? rows[(size_t)(s0 + c1 + j) * 18 + off] : 0.0;
                switch (d) {
                  SP_CASES4(0) SP_CASES4(4) SP_CASES4(8) SP_CASES4(12) SP_CASES4(16) SP_CASE(20) SP_CASE(21)
                  default: break;
                }
              }
            }
          }
        }
      }
    }
    side_store(kst + 1, sv_next, gv_next);
  }
  if (!row_ok) return;
  // ---- the row's in-band blocks of this half: those the pattern holds are written, the others never received a product
#pragma unroll
  for (int d = 0; d < SP_DH; d++) {
    const int b = a + dlo + d;
    if (b >= A.Kf) continue;
    if (!((A.bitmap[(size_t)a * A.words + (b >> 5)] >> (b & 31)) & 1u)) continue;
    const int u = A.csr_u[csr_pos(A.bitmap, A.word_prefix, A.s_rowptr, A.words, a, b)];
    if (ld) {
      A.U_val[(size_t)u * 36 + m * 6 + 2 * k] = -c[d][0];
      A.U_val[(size_t)u * 36 + m * 6 + 2 * k + 1] = -c[d][1];
    }
    if (half == 0 && d == 0 && m < 6 && k == 3) A.bneg[(size_t)a * 6 + m] = -c[0][0];   // C[m][6]
  }
}
#undef SP_CASE
#undef SP_CASES4

// first / last local landmark of every free pose (atomicMin / atomicMax over the observations), and the slot of every observation
__global__ void __launch_bounds__(TPB) k_pose_lm_range(const int* __restrict__ o_kf, const int* __restrict__ o_lm,
                                                       const int* __restrict__ pose_slot, int El, int* __restrict__ o_slot,
                                                       int* __restrict__ lmin, int* __restrict__ lmax, int* __restrict__ cnt) {
  const int e = blockIdx.x * TPB + threadIdx.x;
  if (e >= El) return;
  const int a = pose_slot[o_kf[e]];
  o_slot[e] = a;
  if (a < 0) return;
  atomicMin(lmin + a, o_lm[e]);
  atomicMax(lmax + a, o_lm[e]);
  atomicAdd(cnt + a, 1);
}
__global__ void __launch_bounds__(TPB) k_fill_int(int* __restrict__ p, int n, int v) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i < n) p[i] = v;
}
// a panel is taken when the observations inside its landmark range are at most `factor` times the observations of its own poses
__global__ void __launch_bounds__(TPB) k_panel_on(const int* __restrict__ lmin, const int* __restrict__ lmax, const int* __restrict__ lm_ptr,
                                                  const int* __restrict__ cnt, int Kf, int npan, int factor,
                                                  unsigned char* __restrict__ pan_on) {
  const int p = blockIdx.x * TPB + threadIdx.x;
  if (p >= npan) return;
  int L0 = 0x7fffffff, L1 = -1;
  long long own = 0;
  for (int r = 0; r < SP_R && p * SP_R + r < Kf; r++) {
    const int a = p * SP_R + r;
    if (lmin[a] <= lmax[a]) { L0 = min(L0, lmin[a]); L1 = max(L1, lmax[a] + 1); }
    own += (long long)cnt[a];
  }
  pan_on[p] = (L1 > L0 && (long long)(lm_ptr[L1] - lm_ptr[L0]) <= (long long)factor * own) ? 1 : 0;
}
// covered[u] = 1: the panel kernel owns upper block u
__global__ void __launch_bounds__(TPB) k_covered(const int* __restrict__ u_row, const int* __restrict__ u_col, int nub,
                                                 const unsigned char* __restrict__ pan_on, unsigned char* __restrict__ covered) {
  const int u = blockIdx.x * TPB + threadIdx.x;
  if (u >= nub) return;
  const int a = u_row[u], b = u_col[u];
  covered[u] = (pan_on[a / SP_R] && b - a < SP_DMAX) ? 1 : 0;
}

}  // namespace ba
}  // namespace ccm

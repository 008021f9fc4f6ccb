// ba_kernels.cuh — sm_100a kernels of the Levenberg-Marquardt bundle-adjustment hot path.
//
// Data layout in HBM (one landmark shard per GPU; poses replicated):
//   pose      [K][7]  f64  AoS  (gathered per observation, K is small -> L1/L2 resident)
//   pt        [Pl][3] f64  AoS  (observations are sorted by landmark -> consecutive threads share lines)
//   o_kf,o_lm [El]    i32       observation -> keyframe index / local landmark index (landmark order)
//   o_uv      [El]    float2    undistorted pixel
//   o_w       [El]    f32       invSigma2; sign bit = "no robust kernel"; 0 = inactive edge (level 1)
//   W         [18][Ep] f64 SoA  Hpl block per observation (6x3 row-major entry c at W[c*Ep+e]) -> coalesced stores
//   Z         [El][18] f64 AoS  W * U^-1 with (Hll + lambda I) = U^T U; gathered 144 B rows by the Schur products
//   Hll       [6][Pl]  f64 SoA  upper triangle (00 01 02 11 12 22);  bl [3][Pl]
//   Hpp       [Kf][36], bp [Kf][6] ; U_val [nub][36] upper Schur blocks ; s_val [nnzb][36] full block-CSR for PCG
//
// Kernels (reference loop each one replaces: SURVEY.md §2.2 K1..K7):
//   k_linearize   K1+K2  residual + Huber + Jacobians + W store + Hll/bl warp-segmented reduction + chi2
//   k_pose_pass   K2     Hpp/bp per free pose (CTA per pose, register accumulation, no atomics)
//   k_residual    K1     robust chi2 of a trial state
//   k_scale       K3/K4  Z = W U^-1, g = U^-T bl        (lambda folded in here, no setLambda/restoreDiagonal passes)
//   k_schur       K4     S_ab = sum_l Z_al Z_bl^T over precomputed product lists (register accumulation, no atomics)
//   k_finalize_S  K4     S = [a==b](Hpp + lambda I) - products, mirrored to full block-CSR
//   k_block_jacobi K5    6x6 inverses of the diagonal blocks + bschur
//   k_pcg         K5     persistent cooperative PCG on the reduced camera system
//   k_update_poses / k_backsub_points  K6+K7  back-substitution, oplus into the trial state, gain-ratio denominator
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ba_math.cuh"
#include "pcg.cuh"

namespace ccm {
namespace ba {

using ccm::TPB;
using ccm::warp_sum;
using ccm::block_sum;

__device__ __forceinline__ Pose load_pose(const double* __restrict__ pose, int k) {
  const double* p = pose + 7 * (size_t)k;
  Pose T;
  T.qx = __ldg(p); T.qy = __ldg(p + 1); T.qz = __ldg(p + 2); T.qw = __ldg(p + 3);
  T.tx = __ldg(p + 4); T.ty = __ldg(p + 5); T.tz = __ldg(p + 6);
  return T;
}

// ------------------------------------------------------------------------------------------------------------
// K1+K2: one thread per observation (landmark order).
//   reads  20 B/obs (kf, lm, uv, w) + gathers (pose 56 B, intr 32 B, point 24 B: cache resident)
//   writes 144 B/obs (W, SoA, fully coalesced) + 72 B/landmark (Hll, bl) via warp-segmented reduction
template <int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_linearize(
    const int* __restrict__ o_kf, const int* __restrict__ o_lm, const float2* __restrict__ o_uv,
    const float* __restrict__ o_w, const double* __restrict__ pose, const double* __restrict__ intr,
    const int* __restrict__ pose_slot, const double* __restrict__ pt, int E, size_t Ep, int Pl, int robust,
    double delta, double* __restrict__ W, double* __restrict__ Hll, double* __restrict__ bl,
    double* __restrict__ chi2_partials) {
  __shared__ double red[TPB / 32];
  double chi_acc = 0.0;
  const int lane = threadIdx.x & 31;
  for (long long base = (long long)blockIdx.x * TPB; base < E; base += (long long)gridDim.x * TPB) {
    const long long e = base + threadIdx.x;
    const bool valid = e < E;
    int lm = -1;
    double h[9];  // Hll upper (6) + bl (3)
#pragma unroll
    for (int i = 0; i < 9; i++) h[i] = 0.0;
    if (valid) {
      const int kf = o_kf[e];
      lm = o_lm[e];
      const float2 uv = o_uv[e];
      const float wf = o_w[e];
      const double w = fabs((double)wf);
      const bool rob = robust && !signbit(wf);
      const Pose T = load_pose(pose, kf);
      const double in4[4] = {__ldg(intr + 4 * (size_t)kf), __ldg(intr + 4 * (size_t)kf + 1),
                             __ldg(intr + 4 * (size_t)kf + 2), __ldg(intr + 4 * (size_t)kf + 3)};
      const double* X = pt + 3 * (size_t)lm;
      ObsLin L;
      linearize_obs(T, in4, X[0], X[1], X[2], (double)uv.x, (double)uv.y, w, L);
      double rho0 = L.chi2, rho1 = 1.0;
      if (rob) huber(L.chi2, delta, rho0, rho1);
      chi_acc += rho0;
      const double wo = rho1 * w;               // weightedOmega = rho'(chi2) * omega
      const double r0 = -wo * L.ex, r1 = -wo * L.ey;  // omega_r = -omega e * rho'
      // point block
      h[0] = wo * (L.Jl[0] * L.Jl[0] + L.Jl[3] * L.Jl[3]);
      h[1] = wo * (L.Jl[0] * L.Jl[1] + L.Jl[3] * L.Jl[4]);
      h[2] = wo * (L.Jl[0] * L.Jl[2] + L.Jl[3] * L.Jl[5]);
      h[3] = wo * (L.Jl[1] * L.Jl[1] + L.Jl[4] * L.Jl[4]);
      h[4] = wo * (L.Jl[1] * L.Jl[2] + L.Jl[4] * L.Jl[5]);
      h[5] = wo * (L.Jl[2] * L.Jl[2] + L.Jl[5] * L.Jl[5]);
      h[6] = L.Jl[0] * r0 + L.Jl[3] * r1;
      h[7] = L.Jl[1] * r0 + L.Jl[4] * r1;
      h[8] = L.Jl[2] * r0 + L.Jl[5] * r1;
      // pose-landmark block W = Jp^T (wo I) Jl  (zero when the pose vertex is fixed)
      const double wz = __ldg(pose_slot + kf) >= 0 ? wo : 0.0;
#pragma unroll
      for (int r = 0; r < 6; r++) {
        const double a = wz * L.Jp[r], b = wz * L.Jp[6 + r];
#pragma unroll
        for (int c = 0; c < 3; c++) W[(size_t)(r * 3 + c) * Ep + e] = a * L.Jl[c] + b * L.Jl[3 + c];
      }
    }
    // segmented (by landmark) warp reduction of the 9 point-side sums; observations of one landmark are contiguous
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int lm_o = __shfl_down_sync(0xffffffffu, lm, off);
      const bool take = (lane + off < 32) && (lm_o == lm);
#pragma unroll
      for (int i = 0; i < 9; i++) {
        const double v = __shfl_down_sync(0xffffffffu, h[i], off);
        if (take) h[i] += v;
      }
    }
    const int lm_prev = __shfl_up_sync(0xffffffffu, lm, 1);
    if (lm >= 0 && (lane == 0 || lm_prev != lm)) {
      // a landmark may straddle warps: red.add into the zeroed accumulators (<= a few partial sums per landmark)
#pragma unroll
      for (int i = 0; i < 6; i++) atomicAdd(Hll + (size_t)i * Pl + lm, h[i]);
#pragma unroll
      for (int i = 0; i < 3; i++) atomicAdd(bl + (size_t)i * Pl + lm, h[6 + i]);
    }
  }
  const double tot = block_sum(chi_acc, red);
  if (threadIdx.x == 0) chi2_partials[blockIdx.x] = tot;
}

// K1 on a (trial) state: robust chi2 only.  20 B/obs read.
__global__ void __launch_bounds__(TPB) k_residual(
    const int* __restrict__ o_kf, const int* __restrict__ o_lm, const float2* __restrict__ o_uv,
    const float* __restrict__ o_w, const double* __restrict__ pose, const double* __restrict__ intr,
    const double* __restrict__ pt, int E, int robust, double delta, double* __restrict__ chi2_partials) {
  __shared__ double red[TPB / 32];
  double chi_acc = 0.0;
  for (long long e = (long long)blockIdx.x * TPB + threadIdx.x; e < E; e += (long long)gridDim.x * TPB) {
    const int kf = o_kf[e];
    const int lm = o_lm[e];
    const float2 uv = o_uv[e];
    const float wf = o_w[e];
    const double w = fabs((double)wf);
    const Pose T = load_pose(pose, kf);
    const double in4[4] = {__ldg(intr + 4 * (size_t)kf), __ldg(intr + 4 * (size_t)kf + 1),
                           __ldg(intr + 4 * (size_t)kf + 2), __ldg(intr + 4 * (size_t)kf + 3)};
    const double* X = pt + 3 * (size_t)lm;
    double ex, ey, chi2, Xc[3];
    project_residual(T, in4, X[0], X[1], X[2], (double)uv.x, (double)uv.y, w, ex, ey, chi2, Xc);
    double rho0 = chi2, rho1;
    if (robust && !signbit(wf)) huber(chi2, delta, rho0, rho1);
    chi_acc += rho0;
  }
  const double tot = block_sum(chi_acc, red);
  if (threadIdx.x == 0) chi2_partials[blockIdx.x] = tot;
}

// per-edge report for the caller: plain chi2 at the last evaluated state, depth sign at the final estimate
__global__ void __launch_bounds__(TPB) k_edge_report(
    const int* __restrict__ o_kf, const int* __restrict__ o_lm, const float2* __restrict__ o_uv,
    const float* __restrict__ o_w, const double* __restrict__ pose_eval, const double* __restrict__ pt_eval,
    const double* __restrict__ pose_fin, const double* __restrict__ pt_fin, const double* __restrict__ intr, int E,
    double* __restrict__ chi2_out, uint8_t* __restrict__ depth_out) {
  const long long e = (long long)blockIdx.x * TPB + threadIdx.x;
  if (e >= E) return;
  const int kf = o_kf[e], lm = o_lm[e];
  const float2 uv = o_uv[e];
  const double w = fabs((double)o_w[e]);
  const double in4[4] = {intr[4 * (size_t)kf], intr[4 * (size_t)kf + 1], intr[4 * (size_t)kf + 2], intr[4 * (size_t)kf + 3]};
  double ex, ey, chi2, Xc[3];
  {
    const Pose T = load_pose(pose_eval, kf);
    const double* X = pt_eval + 3 * (size_t)lm;
    project_residual(T, in4, X[0], X[1], X[2], (double)uv.x, (double)uv.y, w, ex, ey, chi2, Xc);
    chi2_out[e] = chi2;
  }
  {
    const Pose T = load_pose(pose_fin, kf);
    const double* X = pt_fin + 3 * (size_t)lm;
    project_residual(T, in4, X[0], X[1], X[2], (double)uv.x, (double)uv.y, w, ex, ey, chi2, Xc);
    depth_out[e] = Xc[2] > 0.0 ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------------------
// per free pose: its observations packed as (u, v, signed weight, landmark) in diagonal-product-list order, so that the pose
// pass streams 16 coalesced bytes per observation instead of chasing four index arrays
__global__ void __launch_bounds__(128) k_pack_pose_obs(const uint2* __restrict__ prod, const unsigned* __restrict__ u_prod_ptr,
                                                       const int* __restrict__ u_diag, const unsigned* __restrict__ kobs_ptr,
                                                       const int* __restrict__ o_lm, const float2* __restrict__ o_uv,
                                                       const float* __restrict__ o_w, float4* __restrict__ kobs) {
  const int a = blockIdx.x;
  const unsigned beg = u_prod_ptr[u_diag[a]], n = u_prod_ptr[u_diag[a] + 1] - beg, dst = kobs_ptr[a];
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned e = prod[beg + i].x;
    const float2 uv = o_uv[e];
    kobs[dst + i] = make_float4(uv.x, uv.y, o_w[e], __int_as_float(o_lm[e]));
  }
}

// K2 pose side: one CTA per free pose; its observations are the product list of the diagonal Schur block (a,a).
__global__ void __launch_bounds__(128) k_pose_pass(
    const float4* __restrict__ kobs, const unsigned* __restrict__ kobs_ptr,
    const int* __restrict__ slot_pose, const double* __restrict__ pose, const double* __restrict__ intr,
    const double* __restrict__ pt, int robust, double delta, double* __restrict__ Hpp, double* __restrict__ bp) {
  __shared__ double red[27][4];
  const int a = blockIdx.x;
  const int kf = slot_pose[a];
  const unsigned beg = kobs_ptr[a], end = kobs_ptr[a + 1];
  const Pose T = load_pose(pose, kf);
  const double in4[4] = {intr[4 * (size_t)kf], intr[4 * (size_t)kf + 1], intr[4 * (size_t)kf + 2], intr[4 * (size_t)kf + 3]};
  double acc[27];
#pragma unroll
  for (int i = 0; i < 27; i++) acc[i] = 0.0;
  for (unsigned p = beg + threadIdx.x; p < end; p += blockDim.x) {
    const float4 ob = kobs[p];
    const int lm = __float_as_int(ob.w);
    const float wf = ob.z;
    const double w = fabs((double)wf);
    const double* X = pt + 3 * (size_t)lm;
    ObsLin L;
    linearize_obs(T, in4, X[0], X[1], X[2], (double)ob.x, (double)ob.y, w, L);
    double rho0, rho1 = 1.0;
    if (robust && !signbit(wf)) huber(L.chi2, delta, rho0, rho1);
    const double wo = rho1 * w;
    const double r0 = -wo * L.ex, r1 = -wo * L.ey;
    int q = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = i; j < 6; j++) acc[q++] += wo * (L.Jp[i] * L.Jp[j] + L.Jp[6 + i] * L.Jp[6 + j]);
#pragma unroll
    for (int i = 0; i < 6; i++) acc[21 + i] += L.Jp[i] * r0 + L.Jp[6 + i] * r1;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 27; i++) {
    const double v = warp_sum(acc[i]);
    if (lane == 0) red[i][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 27) {
    const double v = red[threadIdx.x][0] + red[threadIdx.x][1] + red[threadIdx.x][2] + red[threadIdx.x][3];
    if (threadIdx.x < 21) {
      // unpack upper-triangular index -> (i,j)
      int i = 0, q = threadIdx.x;
      while (q >= 6 - i) { q -= 6 - i; i++; }
      const int j = i + q;
      Hpp[(size_t)a * 36 + i * 6 + j] = v;
      Hpp[(size_t)a * 36 + j * 6 + i] = v;
    } else {
      bp[(size_t)a * 6 + (threadIdx.x - 21)] = v;
    }
  }
}

// max |diag| over Hpp and Hll -> bits of a non-negative double, atomicMax as unsigned long long
__global__ void __launch_bounds__(TPB) k_max_diag(const double* __restrict__ Hpp, int Kf, const double* __restrict__ Hll,
                                                  int Pl, unsigned long long* __restrict__ out) {
  double m = 0.0;
  const long long n1 = (long long)Kf * 6, n2 = (long long)Pl * 3;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n1 + n2; i += (long long)gridDim.x * TPB) {
    double v;
    if (i < n1) {
      v = Hpp[(i / 6) * 36 + (i % 6) * 7];
    } else {
      const long long j = i - n1;
      const int d = (int)(j / Pl);  // 0,1,2 -> rows 0,3,5 of the packed upper triangle
      const int row = d == 0 ? 0 : (d == 1 ? 3 : 5);
      v = Hll[(size_t)row * Pl + (j % Pl)];
    }
    m = fmax(m, fabs(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// single-block deterministic sum of partials: out[0] = sum
__global__ void __launch_bounds__(1024) k_sum_partials(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
  const double t = block_sum(v, red);
  if (threadIdx.x == 0) out[0] = t;
}

// ------------------------------------------------------------------------------------------------------------
// K3/K4 first half: Z = W U^-1 (AoS, staged through shared memory for coalesced rows), g = U^-T bl.
// reads 144 B/obs (W) + cached Hll ; writes 144 B/obs (Z)
__global__ void __launch_bounds__(TPB) k_scale(const int* __restrict__ o_lm, const double* __restrict__ W, size_t Ep,
                                               const double* __restrict__ Hll, const double* __restrict__ bl, int Pl,
                                               int E, double lambda, double* __restrict__ Z, double* __restrict__ gvec) {
  __shared__ double tile[TPB * 19];
  const long long e0 = (long long)blockIdx.x * TPB;
  const long long e = e0 + threadIdx.x;
  if (e < E) {
    const int lm = o_lm[e];
    double d[6], u[6];
#pragma unroll
    for (int i = 0; i < 6; i++) d[i] = __ldg(Hll + (size_t)i * Pl + lm);
    d[0] += lambda; d[3] += lambda; d[5] += lambda;
    chol3_upper(d, u);
#pragma unroll
    for (int r = 0; r < 6; r++) {
      const double w0 = W[(size_t)(r * 3) * Ep + e], w1 = W[(size_t)(r * 3 + 1) * Ep + e], w2 = W[(size_t)(r * 3 + 2) * Ep + e];
      double z0, z1, z2;
      row_times_Uinv(u, w0, w1, w2, z0, z1, z2);
      tile[threadIdx.x * 19 + r * 3] = z0;
      tile[threadIdx.x * 19 + r * 3 + 1] = z1;
      tile[threadIdx.x * 19 + r * 3 + 2] = z2;
    }
    const bool head = (e == 0) || (o_lm[e - 1] != lm);
    if (head) {
      const double b3[3] = {__ldg(bl + lm), __ldg(bl + (size_t)Pl + lm), __ldg(bl + 2 * (size_t)Pl + lm)};
      double g[3];
      UTinv_times(u, b3, g);
      gvec[3 * (size_t)lm] = g[0]; gvec[3 * (size_t)lm + 1] = g[1]; gvec[3 * (size_t)lm + 2] = g[2];
    }
  }
  __syncthreads();
  const long long nvalid = (E - e0 < TPB ? E - e0 : TPB) * 18;
  double* out = Z + e0 * 18;
  for (int i = threadIdx.x; i < nvalid; i += TPB) out[i] = tile[(i / 18) * 19 + (i % 18)];
}

// ------------------------------------------------------------------------------------------------------------
// K4: Schur products.  One warp per upper block u = (a,b): lane = (row r = lane % 6, stream s = lane / 6), 5 streams.
//   acc[r][c] += sum_k Z_oa[r][k] * Z_ob[c][k]  over the block's product list; diagonal blocks also build
//   bneg_a = sum_o Z_o g_l(o).  Outputs are written NEGATED (S = Hpp + lambda I - sum).
__global__ void __launch_bounds__(TPB) k_schur(const uint2* __restrict__ prod, const unsigned* __restrict__ u_prod_ptr,
                                               const int* __restrict__ u_row, const int* __restrict__ u_col, int nub,
                                               const double* __restrict__ Z, const int* __restrict__ o_lm,
                                               const double* __restrict__ gvec, double* __restrict__ U_val,
                                               double* __restrict__ bneg) {
  const int warp = (int)(((long long)blockIdx.x * TPB + threadIdx.x) >> 5);
  if (warp >= nub) return;
  const int lane = threadIdx.x & 31;
  const int r = lane % 6, s = lane / 6;  // lanes 30,31: s == 5 -> idle stream
  const unsigned beg = u_prod_ptr[warp], end = u_prod_ptr[warp + 1];
  const bool diag = u_row[warp] == u_col[warp];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  double bacc = 0.0;
  if (s < 5) {
    for (unsigned p = beg + s; p < end; p += 5) {
      const uint2 pr = prod[p];
      const double* za = Z + (size_t)pr.x * 18 + r * 3;
      const double a0 = za[0], a1 = za[1], a2 = za[2];
      const double2* zb = reinterpret_cast<const double2*>(Z + (size_t)pr.y * 18);
      double b[18];
#pragma unroll
      for (int i = 0; i < 9; i++) {
        const double2 t = zb[i];
        b[2 * i] = t.x; b[2 * i + 1] = t.y;
      }
#pragma unroll
      for (int c = 0; c < 6; c++) acc[c] += a0 * b[c * 3] + a1 * b[c * 3 + 1] + a2 * b[c * 3 + 2];
      if (diag) {
        const double* g = gvec + 3 * (size_t)o_lm[pr.x];
        bacc += a0 * g[0] + a1 * g[1] + a2 * g[2];
      }
    }
  }
  // combine the 5 streams: lane r gathers lanes r+6, r+12, r+18, r+24
#pragma unroll
  for (int c = 0; c < 6; c++) {
    double v = acc[c];
    v += __shfl_sync(0xffffffffu, acc[c], (r + 6) & 31) + __shfl_sync(0xffffffffu, acc[c], (r + 12) & 31) +
         __shfl_sync(0xffffffffu, acc[c], (r + 18) & 31) + __shfl_sync(0xffffffffu, acc[c], (r + 24) & 31);
    acc[c] = v;
  }
  {
    double v = bacc;
    v += __shfl_sync(0xffffffffu, bacc, (r + 6) & 31) + __shfl_sync(0xffffffffu, bacc, (r + 12) & 31) +
         __shfl_sync(0xffffffffu, bacc, (r + 18) & 31) + __shfl_sync(0xffffffffu, bacc, (r + 24) & 31);
    bacc = v;
  }
  if (lane < 6) {
#pragma unroll
    for (int c = 0; c < 6; c++) U_val[(size_t)warp * 36 + r * 6 + c] = -acc[c];
    if (diag) bneg[(size_t)u_row[warp] * 6 + r] = -bacc;
  }
}

// K4, tensor-core form (CCM_SCHUR=mma).  The gather form above spends its time in L1TEX wavefronts: every lane of a
// stream pulls 21 doubles per product.  Here one product Z_oa (6x3) . Z_ob^T (3x6) is ONE mma.sync.m8n8k4.f64 whose
// operand fragments are exactly one coalesced 144-byte row each:
//   A (8x4 row-major): lane t holds A[t/4][t%4] = Z_oa[t/4][t%4]        rows 6,7 and column 3 are zero padding
//   B (4x8 col-major): lane t holds B[t%4][t/4] = Z_ob[t/4][t%4]        -> the same offset (t/4)*3 + t%4 into the row
//   C (8x8)          : lane t holds C[t/4][2*(t%4)], C[t/4][2*(t%4)+1]  accumulated in place over the product list
// i.e. 18 lanes x 8 B per operand (2 cache lines), no shuffles, no cross-lane reduction.  Diagonal blocks put g_l into
// B's column 6, so C[r][6] = sum_o Z_o[r][:] . g_l(o) = bneg comes out of the same instruction.  Two accumulator sets
// (products alternate) keep two MMA chains in flight; they are added in a fixed order: deterministic per list order.
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// TILED: the CTA takes the upper blocks of one T x T tile of S (tile_ptr / tile_u, built by build_schur_tiles): its warps share
// the Z rows of T block rows and T block columns, and -- the product lists being sorted by landmark -- meet them at about the same
// time, so most of the 2 x 144 bytes per product come from L1 instead of L2.
// VEC: the UNROLL list entries of a batch come in with ONE coalesced load (lane j takes entry j) and reach the other lanes by shuffle,
// instead of UNROLL broadcast loads: the kernel sits at 65 % of the L1 wavefront rate, and the entry loads are a fifth of its wavefronts.
// PRED (with VEC): the 14 padding lanes of a fragment do not load at all (predicated off) instead of re-reading element 0 of the row.
// WIDE (with VEC): a 144-byte row comes in as nine 16-byte loads (lanes 0..8) and reaches its fragment lanes by two 64-bit shuffles,
// instead of eighteen 8-byte lanes: fewer L1 wavefronts per row, more shuffles.
// SMB (with VEC, CCM_SCHUR=15): the entries of a batch go from the loading lanes through a per-warp shared-memory slot (double-buffered:
// one __syncwarp per batch) and come back as broadcast 16-byte loads, two entries each: UNROLL / 2 + 1 shared-memory wavefronts per
// batch instead of 2 UNROLL shuffles (a shuffle is a wavefront of the same L1 data pipe).  Measured: 7.62 ms against 6.90 ms -- the
// store / barrier / load chain in front of every batch costs more latency than the wavefronts are worth.
template <int UNROLL, int CTA, bool PIPE = false, bool TILED = false, bool VEC = false, bool PRED = false, bool WIDE = false,
          bool SMB = false>
__global__ void __launch_bounds__(CTA) k_schur_mma(const uint2* __restrict__ prod, const unsigned* __restrict__ u_prod_ptr,
                                                   const int* __restrict__ u_row, const int* __restrict__ u_col, int nub,
                                                   const double* __restrict__ Z, const int* __restrict__ o_lm,
                                                   const double* __restrict__ gvec, double* __restrict__ U_val,
                                                   double* __restrict__ bneg, const int* __restrict__ tile_ptr = nullptr,
                                                   const int* __restrict__ tile_u = nullptr,
                                                   const unsigned char* __restrict__ covered = nullptr, int only_diag = 0) {
  static_assert(UNROLL % 2 == 0, "products alternate between two accumulator sets");
  int warp;
  if (TILED) {
    const int t0 = tile_ptr[blockIdx.x], t1 = tile_ptr[blockIdx.x + 1];
    const int w = threadIdx.x >> 5;
    if (w >= t1 - t0) return;  // ragged tile (diagonal, band edge): warp-uniform
    warp = tile_u[t0 + w];
  } else {
    warp = (int)(((long long)blockIdx.x * CTA + threadIdx.x) >> 5);
    if (warp >= nub) return;  // warp-uniform
  }
  if (covered != nullptr && covered[warp]) return;  // this block belongs to the panel kernel (schur_panel.cuh)
  const int lane = threadIdx.x & 31;
  const int m = lane >> 2, k = lane & 3;
  const bool ld = m < 6 && k < 3;
  constexpr int zs = 18;   // doubles per row of Z.  Padded rows (a 128-byte line per half-warp, 160 / 192 / 256-byte strides) were measured
                           // and are no faster: 7.16 .. 7.36 ms against 7.21 ms (profiles/r2/zlayout_cfg5.log)
  // padding lanes read an element their own half-warp reads anyway (no extra line: 7.21 -> 7.13 ms) and discard it
  const int off = ld ? m * 3 + k : (lane < 16 ? 0 : 12);
  const unsigned beg = u_prod_ptr[warp], end = u_prod_ptr[warp + 1];
  const int row = u_row[warp];
  const bool diag = row == u_col[warp];
  if (only_diag && !diag) return;  // the off-diagonal blocks belong to k_schur_rowsync
  double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
  unsigned p = beg;
  __shared__ uint4 s_ent[SMB ? (CTA / 32) * 2 * (UNROLL / 2) : 1];   // [warp][buffer][UNROLL entries of 8 bytes]
  uint4* const my_ent = s_ent + (SMB ? (threadIdx.x >> 5) * 2 * (UNROLL / 2) : 0);
  unsigned batch = 0;
  if (!diag && PIPE && VEC) {
    uint2 nxv = make_uint2(0u, 0u);
    if (p + UNROLL <= end && lane < UNROLL) nxv = prod[p + lane];
    for (; p + UNROLL <= end; p += UNROLL) {
      uint2 pr[UNROLL];
      if (SMB) {
        uint4* slot = my_ent + (batch & 1u) * (UNROLL / 2);
        batch++;
        if (lane < UNROLL) reinterpret_cast<uint2*>(slot)[lane] = nxv;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < UNROLL; j += 2) {
          const uint4 t = slot[j >> 1];
          pr[j] = make_uint2(t.x, t.y); pr[j + 1] = make_uint2(t.z, t.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < UNROLL; j++) { pr[j].x = __shfl_sync(0xffffffffu, nxv.x, j); pr[j].y = __shfl_sync(0xffffffffu, nxv.y, j); }
      }
      if (p + 2 * UNROLL <= end && lane < UNROLL) nxv = prod[p + UNROLL + lane];
      double a[UNROLL], b[UNROLL];
#pragma unroll
      for (int j = 0; j < UNROLL; j++) {
        if (WIDE) {
          double2 ra = make_double2(0.0, 0.0), rb = make_double2(0.0, 0.0);
          if (lane < 9) {
            ra = reinterpret_cast<const double2*>(Z + (size_t)pr[j].x * zs)[lane];
            rb = reinterpret_cast<const double2*>(Z + (size_t)pr[j].y * zs)[lane];
          }
          const int cl = ld ? m * 3 + k : 0, src = cl >> 1;
          const double ax = __shfl_sync(0xffffffffu, ra.x, src), ay = __shfl_sync(0xffffffffu, ra.y, src);
          const double bx = __shfl_sync(0xffffffffu, rb.x, src), by = __shfl_sync(0xffffffffu, rb.y, src);
          a[j] = (cl & 1) ? ay : ax;
          b[j] = (cl & 1) ? by : bx;
        } else if (PRED) {
          a[j] = 0.0; b[j] = 0.0;
          if (ld) { a[j] = Z[(size_t)pr[j].x * zs + off]; b[j] = Z[(size_t)pr[j].y * zs + off]; }
        } else {
          a[j] = Z[(size_t)pr[j].x * zs + off];
          b[j] = Z[(size_t)pr[j].y * zs + off];
        }
      }
#pragma unroll
      for (int j = 0; j < UNROLL; j += 2) {
        dmma_884(c00, c01, ld ? a[j] : 0.0, ld ? b[j] : 0.0);
        dmma_884(c10, c11, ld ? a[j + 1] : 0.0, ld ? b[j + 1] : 0.0);
      }
    }
    for (; p < end; p++) {
      const uint2 pr = prod[p];
      const double a = Z[(size_t)pr.x * zs + off], b = Z[(size_t)pr.y * zs + off];
      if ((p - beg) & 1u) dmma_884(c10, c11, ld ? a : 0.0, ld ? b : 0.0);
      else dmma_884(c00, c01, ld ? a : 0.0, ld ? b : 0.0);
    }
  } else if (!diag && PIPE) {
    // software-pipelined form (8.43 ms against 8.73 ms on cfg5: the default): the product entries of batch k+1 are requested before the rows of batch k,
    // so the entry -> row dependence costs one memory latency per batch instead of two
    uint2 nx[UNROLL];
    if (p + UNROLL <= end) {
#pragma unroll
      for (int j = 0; j < UNROLL; j++) nx[j] = prod[p + j];
    }
    for (; p + UNROLL <= end; p += UNROLL) {
      uint2 pr[UNROLL];
#pragma unroll
      for (int j = 0; j < UNROLL; j++) pr[j] = nx[j];
      if (p + 2 * UNROLL <= end) {
#pragma unroll
        for (int j = 0; j < UNROLL; j++) nx[j] = prod[p + UNROLL + j];
      }
      double a[UNROLL], b[UNROLL];
#pragma unroll
      for (int j = 0; j < UNROLL; j++) {
        a[j] = Z[(size_t)pr[j].x * zs + off];
        b[j] = Z[(size_t)pr[j].y * zs + off];
      }
#pragma unroll
      for (int j = 0; j < UNROLL; j += 2) {
        dmma_884(c00, c01, ld ? a[j] : 0.0, ld ? b[j] : 0.0);
        dmma_884(c10, c11, ld ? a[j + 1] : 0.0, ld ? b[j + 1] : 0.0);
      }
    }
    for (; p < end; p++) {
      const uint2 pr = prod[p];
      const double a = Z[(size_t)pr.x * zs + off], b = Z[(size_t)pr.y * zs + off];
      if ((p - beg) & 1u) dmma_884(c10, c11, ld ? a : 0.0, ld ? b : 0.0);
      else dmma_884(c00, c01, ld ? a : 0.0, ld ? b : 0.0);
    }
  } else if (!diag) {
    for (; p + UNROLL <= end; p += UNROLL) {
      uint2 pr[UNROLL];
#pragma unroll
      for (int j = 0; j < UNROLL; j++) pr[j] = prod[p + j];  // one address for the whole warp
      double a[UNROLL], b[UNROLL];
#pragma unroll
      for (int j = 0; j < UNROLL; j++) {
        a[j] = Z[(size_t)pr[j].x * zs + off];
        b[j] = Z[(size_t)pr[j].y * zs + off];
      }
#pragma unroll
      for (int j = 0; j < UNROLL; j += 2) {
        dmma_884(c00, c01, ld ? a[j] : 0.0, ld ? b[j] : 0.0);
        dmma_884(c10, c11, ld ? a[j + 1] : 0.0, ld ? b[j + 1] : 0.0);
      }
    }
    for (; p < end; p++) {
      const uint2 pr = prod[p];
      const double a = Z[(size_t)pr.x * zs + off], b = Z[(size_t)pr.y * zs + off];
      if ((p - beg) & 1u) dmma_884(c10, c11, ld ? a : 0.0, ld ? b : 0.0);  // warp-uniform branch
      else dmma_884(c00, c01, ld ? a : 0.0, ld ? b : 0.0);
    }
  } else {
    // diagonal block (Kf of them): lanes 24..26 carry g_l in column 6 of B, so C[r][6] accumulates bneg
    const bool gl = m == 6 && k < 3;
    if (VEC) {   // batches of UNROLL products: entries by one coalesced load, then all rows, landmark ids and g_l of the batch in flight together
      for (; p + UNROLL <= end; p += UNROLL) {
        unsigned ex = 0u;
        if (lane < UNROLL) ex = prod[p + lane].x;
        unsigned e[UNROLL];
        if (SMB && UNROLL % 4 == 0) {
          uint4* slot = my_ent + (batch & 1u) * (UNROLL / 2);
          batch++;
          if (lane < UNROLL) reinterpret_cast<unsigned*>(slot)[lane] = ex;
          __syncwarp();
#pragma unroll
          for (int j = 0; j < UNROLL; j += 4) {
            const uint4 t = slot[j >> 2];
            e[j] = t.x; e[j + 1] = t.y; e[j + 2] = t.z; e[j + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < UNROLL; j++) e[j] = __shfl_sync(0xffffffffu, ex, j);
        }
        double a[UNROLL], g[UNROLL];
        int lm[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
          a[j] = Z[(size_t)e[j] * zs + off];
          lm[j] = gl ? o_lm[e[j]] : 0;
        }
#pragma unroll
        for (int j = 0; j < UNROLL; j++) g[j] = gl ? gvec[3 * (size_t)lm[j] + k] : 0.0;
#pragma unroll
        for (int j = 0; j < UNROLL; j += 2) {
          dmma_884(c00, c01, ld ? a[j] : 0.0, gl ? g[j] : (ld ? a[j] : 0.0));
          dmma_884(c10, c11, ld ? a[j + 1] : 0.0, gl ? g[j + 1] : (ld ? a[j + 1] : 0.0));
        }
      }
    }
    for (; p < end; p++) {
      const uint2 pr = prod[p];
      const double a = Z[(size_t)pr.x * zs + off];
      double b = a;   // a diagonal list holds (e, e) pairs only: one row serves both operands
      if (!ld) b = 0.0;
      if (gl) b = gvec[3 * (size_t)o_lm[pr.x] + k];
      if ((p - beg) & 1u) dmma_884(c10, c11, ld ? a : 0.0, b);
      else dmma_884(c00, c01, ld ? a : 0.0, b);
    }
  }
  c00 += c10; c01 += c11;
  if (ld) {
    U_val[(size_t)warp * 36 + m * 6 + 2 * k] = -c00;
    U_val[(size_t)warp * 36 + m * 6 + 2 * k + 1] = -c01;
  }
  if (diag && m < 6 && k == 3) bneg[(size_t)row * 6 + m] = -c00;  // C[m][6]
}

// Row-synchronous form (CCM_SCHUR=10): a CTA takes up to RS_W consecutive OFF-DIAGONAL upper blocks of ONE block row a (the schedule
// rs_first / rs_count is cut at row boundaries) and its warps walk their product lists -- sorted by the observation of a, i.e. by
// landmark -- chunk by chunk of 2^RS_SHIFT observation indices with a CTA barrier after every chunk.  All warps then need the same
// rows Z_(l, a) at the same time: they are fetched from L2 once per CTA and served from L1 to the other warps, which the free-running
// list kernel does not achieve (its warps drift apart).  The diagonal blocks (lists four times as long, and the g_l column) stay with
// k_schur_mma, launched with only_diag = 1.
constexpr int RS_W = 8;
constexpr int RS_SHIFT = 13;
template <int UNROLL>
__global__ void __launch_bounds__(32 * RS_W) k_schur_rowsync(const uint2* __restrict__ prod, const unsigned* __restrict__ u_prod_ptr,
                                                             const int* __restrict__ rs_first, const int* __restrict__ rs_count,
                                                             const double* __restrict__ Z, double* __restrict__ U_val) {
  __shared__ unsigned s_cmin, s_cmax;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int count = rs_count[blockIdx.x];
  const bool active = w < count;
  const int u = active ? rs_first[blockIdx.x] + w : 0;
  const int m = lane >> 2, k = lane & 3;
  const bool ld = m < 6 && k < 3;
  const int off = ld ? m * 3 + k : 0;
  unsigned p = 0, end = 0;
  if (active) { p = u_prod_ptr[u]; end = u_prod_ptr[u + 1]; }
  if (threadIdx.x == 0) { s_cmin = 0xffffffffu; s_cmax = 0u; }
  __syncthreads();
  if (active && lane == 0 && p < end) {
    atomicMin(&s_cmin, prod[p].x >> RS_SHIFT);
    atomicMax(&s_cmax, prod[end - 1].x >> RS_SHIFT);
  }
  __syncthreads();
  const unsigned cmin = s_cmin, cmax = s_cmax;
  double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
  // (the last pass, c = cmax + 1 with no limit, takes what a list too long for the set-up sort left out of order)
  if (cmin != 0xffffffffu)
    for (unsigned c = cmin; c <= cmax + 1u; c++) {
      const unsigned lim = c > cmax ? 0xfffffffeu : c;
      while (p < end) {
        uint2 pr[UNROLL];
        int n_in = 0;
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
          pr[j] = p + j < end ? prod[p + j] : make_uint2(0xffffffffu, 0u);
          if (pr[j].x != 0xffffffffu && (pr[j].x >> RS_SHIFT) <= lim) n_in = j + 1;   // sorted: the entries of this chunk are a prefix
        }
        if (n_in == 0) break;
        double a[UNROLL], b[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
          const bool in = j < n_in;
          a[j] = (in && ld) ? Z[(size_t)pr[j].x * 18 + off] : 0.0;
          b[j] = (in && ld) ? Z[(size_t)pr[j].y * 18 + off] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < UNROLL; j += 2) {
          dmma_884(c00, c01, a[j], b[j]);
          dmma_884(c10, c11, a[j + 1], b[j + 1]);
        }
        p += n_in;
      }
      __syncthreads();
    }
  if (!active) return;
  c00 += c10; c01 += c11;
  if (ld) {
    U_val[(size_t)u * 36 + m * 6 + 2 * k] = -c00;
    U_val[(size_t)u * 36 + m * 6 + 2 * k + 1] = -c01;
  }
}

// Grouped form (CCM_SCHUR=16 / 17; measured, NOT the default).  The list kernel is bound by the L1 data pipe: every product costs two
// row loads of 144 bytes (two lines each: ~3 wavefronts), and a row Z_(l,a) is loaded once for every block (a, b) whose list holds l.
// Here the off-diagonal upper blocks of a block row are cut into groups of QG consecutive blocks (a, b_1 .. b_QG), one warp per group,
// and the lists are kept per group: an entry is (observation of a, observation of b_1 | none, .., observation of b_QG | none) for one
// landmark, so the row of a is loaded ONCE per entry and multiplied into up to QG accumulator sets; an entry with one member costs
// what a list product costs, so the form never loads more rows than the list kernel.  Diagonal blocks (their lists also feed the pose
// pass, and they carry the g_l column) stay with k_schur_mma, launched with only_diag = 1.
// Outcome on cfg5 (profiles/r2/quad_cfg5.log): 95.2 M entries for 190 M products (2.0 products per entry: 25 % fewer row loads), parity
// green (tests/test_gpu_ba.py under CCM_SCHUR=16 and 17), and 13.6 ms per launch against 6.9 ms for the list kernel (cfg4: 0.20 against
// 0.11 ms): 78-128 registers and sixteen conditional MMAs per batch cost more than the row loads saved.  Kept as a switch.
constexpr int QG = 4;
constexpr unsigned Q_NONE = 0xffffffffu;
__device__ __forceinline__ int csr_pos(const unsigned* __restrict__ bitmap, const int* __restrict__ word_prefix,
                                       const int* __restrict__ s_rowptr, int words, int a, int b);   // defined with the pattern kernels below

// count (fill == 0) or fill (fill == 1) the grouped lists; one thread per local observation e (pose a, landmark l) walks the other
// observations of l in list order and keeps ONE open entry: an observation of a later pose joins it if it falls into the same group
// and its member slot is free, otherwise the entry is closed (a slot of the group's list taken by atomicAdd) and a new one opened.
// Keyframe-ordered observation lists (the usual case) therefore give the densest entries; any other order only costs sharing.
__global__ void __launch_bounds__(TPB) k_quad_entries(const int* __restrict__ o_kf, const int* __restrict__ o_lm,
                                                      const int* __restrict__ lm_ptr, const int* __restrict__ pose_slot,
                                                      const unsigned* __restrict__ bitmap, const int* __restrict__ word_prefix,
                                                      const int* __restrict__ s_rowptr, const int* __restrict__ csr_u, int words,
                                                      int E, int fill, const int* __restrict__ u_diag,
                                                      const int* __restrict__ g_rowstart, unsigned* __restrict__ counters,
                                                      const unsigned* __restrict__ g_ptr, unsigned* __restrict__ gent) {
  const long long e = (long long)blockIdx.x * TPB + threadIdx.x;
  if (e >= E) return;
  const int a = pose_slot[o_kf[e]];
  if (a < 0) return;
  const int lm = o_lm[e];
  const int beg = lm_ptr[lm], end = lm_ptr[lm + 1];
  const int ud = u_diag[a], g0 = g_rowstart[a];
  int open = -1;
  unsigned m0 = Q_NONE, m1 = Q_NONE, m2 = Q_NONE, m3 = Q_NONE;
  auto flush = [&]() {
    if (open < 0) return;
    const unsigned slot = atomicAdd(counters + open, 1u);
    if (fill) {
      unsigned* d = gent + ((size_t)g_ptr[open] + slot) * (QG + 1);
      d[0] = (unsigned)e; d[1] = m0; d[2] = m1; d[3] = m2; d[4] = m3;
    }
  };
  for (int o = beg; o < end; o++) {
    const int b = pose_slot[o_kf[o]];
    if (b <= a) continue;   // strictly upper blocks only (b < 0: fixed pose)
    const int j = csr_u[csr_pos(bitmap, word_prefix, s_rowptr, words, a, b)] - ud - 1;
    const int grp = g0 + (j >> 2), mi = j & 3;
    const unsigned cur = mi == 0 ? m0 : mi == 1 ? m1 : mi == 2 ? m2 : m3;
    if (grp != open || cur != Q_NONE) {
      flush();
      open = grp; m0 = m1 = m2 = m3 = Q_NONE;
    }
    if (mi == 0) m0 = (unsigned)o; else if (mi == 1) m1 = (unsigned)o; else if (mi == 2) m2 = (unsigned)o; else m3 = (unsigned)o;
  }
  flush();
}

// a row element if the member is present, 0 otherwise.  Volatile asm: the compiler would otherwise sink every conditional load into the
// conditional MMA that consumes it (load, wait, multiply, sixteen times in a row) instead of keeping the batch's loads in flight together.
__device__ __forceinline__ double ld_row_if(const double* p, unsigned present) {
  double v;
  asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %2, 0;\n mov.f64 %0, 0d0000000000000000;\n @q ld.global.nc.f64 %0, [%1];\n}"
               : "=d"(v)
               : "l"(p), "r"(present));
  return v;
}

// one warp per group; EB entries per batch: their 5 EB words come in with one coalesced load and reach the lanes by shuffle (SMB:
// through a double-buffered shared-memory slot and broadcast 16-byte loads), then the EB rows of a and the up to QG EB rows of
// the b_i are in flight together; a member that is absent (warp-uniform) neither loads nor multiplies.
template <int EB, int CTA, bool SMB>
__global__ void __launch_bounds__(CTA) k_schur_quad(const unsigned* __restrict__ gent, const unsigned* __restrict__ g_ptr,
                                                    const int* __restrict__ g_first, const int* __restrict__ g_count, int ng,
                                                    const double* __restrict__ Z, double* __restrict__ U_val) {
  static_assert(EB == 4, "a batch is 20 words: five 16-byte pieces");
  constexpr int NW = EB * (QG + 1);
  const int warp = (int)(((long long)blockIdx.x * CTA + threadIdx.x) >> 5);
  if (warp >= ng) return;  // warp-uniform
  const int lane = threadIdx.x & 31;
  const int m = lane >> 2, k = lane & 3;
  const bool ld = m < 6 && k < 3;
  const int off = ld ? m * 3 + k : (lane < 16 ? 0 : 12);
  const unsigned beg = g_ptr[warp], end = g_ptr[warp + 1];
  __shared__ uint4 s_w[SMB ? (CTA / 32) * 2 * (NW / 4) : 1];
  uint4* const my_w = s_w + (SMB ? (threadIdx.x >> 5) * 2 * (NW / 4) : 0);
  double c[QG][2][2];
#pragma unroll
  for (int i = 0; i < QG; i++) { c[i][0][0] = c[i][0][1] = c[i][1][0] = c[i][1][1] = 0.0; }
  unsigned p = beg, batch = 0;
  unsigned nx = Q_NONE;
  if (p + EB <= end && lane < NW) nx = gent[(size_t)p * (QG + 1) + lane];
  for (; p + EB <= end; p += EB) {
    unsigned w[NW];
    if (SMB) {
      uint4* slot = my_w + (batch & 1u) * (NW / 4);
      batch++;
      if (lane < NW) reinterpret_cast<unsigned*>(slot)[lane] = nx;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NW; j += 4) {
        const uint4 t = slot[j >> 2];
        w[j] = t.x; w[j + 1] = t.y; w[j + 2] = t.z; w[j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NW; j++) w[j] = __shfl_sync(0xffffffffu, nx, j);
    }
    if (p + 2 * EB <= end && lane < NW) nx = gent[(size_t)(p + EB) * (QG + 1) + lane];
    double a[EB], b[EB][QG];
#pragma unroll
    for (int e = 0; e < EB; e++) {
      a[e] = Z[(size_t)w[e * (QG + 1)] * 18 + off];
#pragma unroll
      for (int i = 0; i < QG; i++) {
        const unsigned ob = w[e * (QG + 1) + 1 + i];
        b[e][i] = ld_row_if(Z + (size_t)ob * 18 + off, ob != Q_NONE ? 1u : 0u);
      }
    }
#pragma unroll
    for (int e = 0; e < EB; e++) {
#pragma unroll
      for (int i = 0; i < QG; i++)
        if (w[e * (QG + 1) + 1 + i] != Q_NONE) dmma_884(c[i][e & 1][0], c[i][e & 1][1], ld ? a[e] : 0.0, ld ? b[e][i] : 0.0);
    }
  }
  for (; p < end; p++) {
    unsigned t = Q_NONE;
    if (lane < QG + 1) t = gent[(size_t)p * (QG + 1) + lane];
    const unsigned oa = __shfl_sync(0xffffffffu, t, 0);
    const double a = Z[(size_t)oa * 18 + off];
#pragma unroll
    for (int i = 0; i < QG; i++) {
      const unsigned ob = __shfl_sync(0xffffffffu, t, 1 + i);
      if (ob != Q_NONE) {
        const double b = Z[(size_t)ob * 18 + off];
        dmma_884(c[i][0][0], c[i][0][1], ld ? a : 0.0, ld ? b : 0.0);
      }
    }
  }
  const int u0 = g_first[warp], cnt = g_count[warp];
  if (ld) {
#pragma unroll
    for (int i = 0; i < QG; i++)
      if (i < cnt) {
        U_val[(size_t)(u0 + i) * 36 + m * 6 + 2 * k] = -(c[i][0][0] + c[i][1][0]);
        U_val[(size_t)(u0 + i) * 36 + m * 6 + 2 * k + 1] = -(c[i][0][1] + c[i][1][1]);
      }
  }
}

// S (full block-CSR) from the upper blocks: diagonal gets Hpp + lambda I, lower blocks are transposed copies.
__global__ void __launch_bounds__(TPB) k_finalize_S(const int* __restrict__ s_row, const int* __restrict__ s_col,
                                                    const int* __restrict__ csr_u, long long nnzb,
                                                    const double* __restrict__ U_val, const double* __restrict__ Hpp,
                                                    double lambda, double* __restrict__ s_val) {
  const long long i = (long long)blockIdx.x * TPB + threadIdx.x;
  if (i >= nnzb * 36) return;
  const long long pos = i / 36;
  const int ent = (int)(i % 36);
  const int a = s_row[pos], b = s_col[pos];
  const int u = csr_u[pos];
  double v;
  if (a <= b) {
    v = U_val[(size_t)u * 36 + ent];
    if (a == b) {
      v += Hpp[(size_t)a * 36 + ent];
      if (ent % 7 == 0) v += lambda;
    }
  } else {
    v = U_val[(size_t)u * 36 + (ent % 6) * 6 + ent / 6];
  }
  s_val[i] = v;
}

// block-Jacobi preconditioner: Minv_a = inverse(S_aa) by Cholesky; bschur = bp + bneg
__global__ void __launch_bounds__(128) k_block_jacobi(const int* __restrict__ s_diag, const double* __restrict__ s_val,
                                                      const double* __restrict__ bp, const double* __restrict__ bneg,
                                                      int Kf, double* __restrict__ Minv, double* __restrict__ bschur,
                                                      int* __restrict__ fail) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= Kf) return;
  double A[36], Li[36];
  const double* src = s_val + (size_t)s_diag[a] * 36;
#pragma unroll
  for (int i = 0; i < 36; i++) A[i] = src[i];
  // Cholesky A = L L^T (lower), in place
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double d = A[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= A[j * 6 + k] * A[j * 6 + k];
    if (!(d > 0.0)) { ok = false; d = 1.0; }
    const double l = sqrt(d);
    A[j * 6 + j] = l;
    const double il = 1.0 / l;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double v = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; k++) v -= A[i * 6 + k] * A[j * 6 + k];
      A[i * 6 + j] = v * il;
    }
  }
  // Li = L^-1 (lower)
#pragma unroll
  for (int i = 0; i < 36; i++) Li[i] = 0.0;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    Li[c * 6 + c] = 1.0 / A[c * 6 + c];
#pragma unroll
    for (int i = c + 1; i < 6; i++) {
      double v = 0.0;
#pragma unroll
      for (int k = c; k < i; k++) v -= A[i * 6 + k] * Li[k * 6 + c];
      Li[i * 6 + c] = v / A[i * 6 + i];
    }
  }
  // Minv = Li^T Li
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 6; j++) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) v += Li[k * 6 + i] * Li[k * 6 + j];
      Minv[(size_t)a * 36 + i * 6 + j] = v;
    }
#pragma unroll
  for (int i = 0; i < 6; i++) bschur[(size_t)a * 6 + i] = bp[(size_t)a * 6 + i] + bneg[(size_t)a * 6 + i];
  if (!ok) atomicExch(fail, 1);
}

// ------------------------------------------------------------------------------------------------------------
// K6+K7 (poses): trial = exp(x) * T ; partial of sum x (lambda x + b) over pose unknowns
__global__ void __launch_bounds__(TPB) k_update_poses(const double* __restrict__ pose, const int* __restrict__ pose_slot,
                                                      const double* __restrict__ x, const double* __restrict__ bp, int K,
                                                      double lambda, double* __restrict__ pose_trial,
                                                      double* __restrict__ scale_partials) {
  __shared__ double red[TPB / 32];
  double acc = 0.0;
  for (int k = blockIdx.x * TPB + threadIdx.x; k < K; k += gridDim.x * TPB) {
    Pose T;
    const double* p = pose + 7 * (size_t)k;
    T.qx = p[0]; T.qy = p[1]; T.qz = p[2]; T.qw = p[3]; T.tx = p[4]; T.ty = p[5]; T.tz = p[6];
    const int s = pose_slot[k];
    if (s >= 0) {
      double u[6];
#pragma unroll
      for (int i = 0; i < 6; i++) {
        u[i] = x[(size_t)s * 6 + i];
        acc += u[i] * (lambda * u[i] + bp[(size_t)s * 6 + i]);
      }
      T = se3_exp_times(u, T);
    }
    double* o = pose_trial + 7 * (size_t)k;
    o[0] = T.qx; o[1] = T.qy; o[2] = T.qz; o[3] = T.qw; o[4] = T.tx; o[5] = T.ty; o[6] = T.tz;
  }
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) scale_partials[blockIdx.x] = t;
}

// K6+K7 (landmarks): xl = U^-1 (g - sum_o Z_o^T xp), trial point = point + xl, partial of sum xl (lambda xl + bl)
__global__ void __launch_bounds__(TPB) k_backsub_points(const int* __restrict__ lm_ptr, const int* __restrict__ o_kf,
                                                        const int* __restrict__ pose_slot, const double* __restrict__ Z,
                                                        const double* __restrict__ Hll, const double* __restrict__ bl,
                                                        const double* __restrict__ x, const double* __restrict__ pt,
                                                        int Pl, double lambda, double* __restrict__ pt_trial,
                                                        double* __restrict__ dx_out, double* __restrict__ scale_partials) {
  __shared__ double red[TPB / 32];
  double acc = 0.0;
  for (int l = blockIdx.x * TPB + threadIdx.x; l < Pl; l += gridDim.x * TPB) {
    double d[6], u[6], b3[3], g[3], t[3] = {0, 0, 0}, xl[3];
#pragma unroll
    for (int i = 0; i < 6; i++) d[i] = Hll[(size_t)i * Pl + l];
    d[0] += lambda; d[3] += lambda; d[5] += lambda;
    chol3_upper(d, u);
    b3[0] = bl[l]; b3[1] = bl[(size_t)Pl + l]; b3[2] = bl[2 * (size_t)Pl + l];
    UTinv_times(u, b3, g);
    const int beg = lm_ptr[l], end = lm_ptr[l + 1];
    for (int o = beg; o < end; o++) {
      const int s = __ldg(pose_slot + o_kf[o]);
      if (s < 0) continue;
      const double2* z2 = reinterpret_cast<const double2*>(Z + (size_t)o * 18);
      double z[18];
#pragma unroll
      for (int i = 0; i < 9; i++) { const double2 v = z2[i]; z[2 * i] = v.x; z[2 * i + 1] = v.y; }
#pragma unroll
      for (int r = 0; r < 6; r++) {
        const double xr = __ldg(x + (size_t)s * 6 + r);
        t[0] += z[r * 3] * xr; t[1] += z[r * 3 + 1] * xr; t[2] += z[r * 3 + 2] * xr;
      }
    }
    const double gm[3] = {g[0] - t[0], g[1] - t[1], g[2] - t[2]};
    Uinv_times(u, gm, xl);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      pt_trial[3 * (size_t)l + c] = pt[3 * (size_t)l + c] + xl[c];
      acc += xl[c] * (lambda * xl[c] + b3[c]);
      if (dx_out) dx_out[3 * (size_t)l + c] = xl[c];
    }
  }
  const double tt = block_sum(acc, red);
  if (threadIdx.x == 0) scale_partials[blockIdx.x] = tt;
}

// ------------------------------------------------------------------------------------------------------------
// structure building (once per create): covisibility bitmap -> block-CSR pattern -> Schur product lists
__global__ void __launch_bounds__(TPB) k_pattern_bitmap(const int* __restrict__ g_kf, const int* __restrict__ g_lm_of,
                                                        const int* __restrict__ g_lm_ptr, const int* __restrict__ pose_slot,
                                                        long long E, int words, unsigned* __restrict__ bitmap) {
  const long long e = (long long)blockIdx.x * TPB + threadIdx.x;
  if (e >= E) return;
  const int a = pose_slot[g_kf[e]];
  if (a < 0) return;
  const int lm = g_lm_of[e];
  const int beg = g_lm_ptr[lm], end = g_lm_ptr[lm + 1];
  for (int o = beg; o < end; o++) {
    const int b = pose_slot[g_kf[o]];
    if (b < 0) continue;
    unsigned* wp = bitmap + (size_t)a * words + (b >> 5);
    const unsigned bit = 1u << (b & 31);
    if (!(*(volatile unsigned*)wp & bit)) atomicOr(wp, bit);
  }
}

__global__ void __launch_bounds__(TPB) k_set_diag_bits(int Kf, int words, unsigned* __restrict__ bitmap) {
  const int a = blockIdx.x * TPB + threadIdx.x;
  if (a < Kf) atomicOr(bitmap + (size_t)a * words + (a >> 5), 1u << (a & 31));
}

// per row: exclusive popcount prefix per word + row total
__global__ void __launch_bounds__(TPB) k_row_prefix(const unsigned* __restrict__ bitmap, int Kf, int words,
                                                    int* __restrict__ word_prefix, int* __restrict__ row_count) {
  const int a = blockIdx.x * TPB + threadIdx.x;
  if (a >= Kf) return;
  int run = 0;
  for (int w = 0; w < words; w++) {
    word_prefix[(size_t)a * words + w] = run;
    run += __popc(bitmap[(size_t)a * words + w]);
  }
  row_count[a] = run;
}

__global__ void __launch_bounds__(TPB) k_fill_cols(const unsigned* __restrict__ bitmap, const int* __restrict__ word_prefix,
                                                   const int* __restrict__ s_rowptr, int Kf, int words,
                                                   int* __restrict__ s_col, int* __restrict__ s_row) {
  const long long i = (long long)blockIdx.x * TPB + threadIdx.x;
  if (i >= (long long)Kf * words) return;
  const int a = (int)(i / words), w = (int)(i % words);
  unsigned bits = bitmap[i];
  int pos = s_rowptr[a] + word_prefix[i];
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    s_col[pos] = w * 32 + b;
    s_row[pos] = a;
    pos++;
  }
}

__device__ __forceinline__ int csr_pos(const unsigned* __restrict__ bitmap, const int* __restrict__ word_prefix,
                                       const int* __restrict__ s_rowptr, int words, int a, int b) {
  const size_t wi = (size_t)a * words + (b >> 5);
  return s_rowptr[a] + word_prefix[wi] + __popc(bitmap[wi] & ((1u << (b & 31)) - 1u));
}

// CSR position of the diagonal block of every row and the number of upper blocks (diagonal included) of the row
__global__ void __launch_bounds__(TPB) k_diag_pos(const unsigned* __restrict__ bitmap, const int* __restrict__ word_prefix,
                                                 const int* __restrict__ s_rowptr, int Kf, int words, int* __restrict__ s_diag,
                                                 int* __restrict__ upper_count) {
  const int a = blockIdx.x * TPB + threadIdx.x;
  if (a >= Kf) return;
  const int q = csr_pos(bitmap, word_prefix, s_rowptr, words, a, a);
  s_diag[a] = q;
  upper_count[a] = s_rowptr[a + 1] - q;
}

// Upper-block numbering (row-major over the entries with column >= row: the columns of a row are sorted, so its upper part is
// [s_diag[a], s_rowptr[a+1]) and numbers consecutively from u_rowstart[a]), the (row, column) of every upper block, and for every
// position of the full pattern the upper block that holds it (the transposed one for the lower triangle).
__global__ void __launch_bounds__(TPB) k_upper_index(const int* __restrict__ s_row, const int* __restrict__ s_col, long long nnzb,
                                                    const unsigned* __restrict__ bitmap, const int* __restrict__ word_prefix,
                                                    const int* __restrict__ s_rowptr, int words, const int* __restrict__ s_diag,
                                                    const int* __restrict__ u_rowstart, int* __restrict__ csr_u,
                                                    int* __restrict__ u_row, int* __restrict__ u_col, int* __restrict__ err) {
  const long long q = (long long)blockIdx.x * TPB + threadIdx.x;
  if (q >= nnzb) return;
  const int a = s_row[q], b = s_col[q];
  if (b >= a) {
    const int u = u_rowstart[a] + (int)(q - s_diag[a]);
    csr_u[q] = u; u_row[u] = a; u_col[u] = b;
  } else {
    if (!((bitmap[(size_t)b * words + (a >> 5)] >> (a & 31)) & 1u)) { atomicOr(err, 1); return; }  // pattern must be symmetric
    const int qb = csr_pos(bitmap, word_prefix, s_rowptr, words, b, a);
    csr_u[q] = u_rowstart[b] + (qb - s_diag[b]);
  }
}

// out[i] = in[i] - off  (local landmark ids / local observation offsets of a shard)
__global__ void __launch_bounds__(TPB) k_shift(const int* __restrict__ in, long long n, int off, int* __restrict__ out) {
  const long long i = (long long)blockIdx.x * TPB + threadIdx.x;
  if (i < n) out[i] = in[i] - off;
}

// Product lists come out of k_products in the order the atomic slots were handed out.  Sorting every list by (first observation,
// second observation) = by landmark makes the Schur sums deterministic and lets the warps of a tile walk Z together (k_schur_mma,
// TILED).  One CTA per list, bitonic network in shared memory; lists beyond the shared buffer (a keyframe with more than 4096
// observations in one block) keep their order, which only costs locality.
constexpr int SORT_CAP = 4096;
__global__ void __launch_bounds__(256) k_sort_products(const unsigned* __restrict__ u_prod_ptr, int nub, uint2* __restrict__ prod) {
  __shared__ unsigned long long s[SORT_CAP];
  const int u = blockIdx.x;
  if (u >= nub) return;
  const unsigned beg = u_prod_ptr[u], end = u_prod_ptr[u + 1];
  const int n = (int)(end - beg);
  if (n <= 1 || n > SORT_CAP) return;
  int m = 2;
  while (m < n) m <<= 1;
  for (int i = threadIdx.x; i < m; i += 256) {
    unsigned long long v = ~0ull;
    if (i < n) { const uint2 p = prod[beg + i]; v = ((unsigned long long)p.x << 32) | p.y; }
    s[i] = v;
  }
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m; i += 256) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s[i], b = s[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s[i] = b; s[ixj] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += 256) prod[beg + i] = make_uint2((unsigned)(s[i] >> 32), (unsigned)(s[i] & 0xffffffffu));
}

// tile schedule of the upper blocks: key = (row group, column group | position inside the T x T tile), value = block number
__global__ void __launch_bounds__(TPB) k_tile_keys(const int* __restrict__ u_row, const int* __restrict__ u_col, int nub, int T,
                                                   unsigned long long ngroups, unsigned long long* __restrict__ keys,
                                                   int* __restrict__ vals) {
  const int u = blockIdx.x * TPB + threadIdx.x;
  if (u >= nub) return;
  const int a = u_row[u], b = u_col[u];
  keys[u] = (((unsigned long long)(a / T) * ngroups + (unsigned long long)(b / T)) << 8) | (unsigned)((a % T) * T + (b % T));
  vals[u] = u;
}
// head[i] = 1 where a new tile starts in the sorted key list
__global__ void __launch_bounds__(TPB) k_tile_heads(const unsigned long long* __restrict__ keys, int nub, int* __restrict__ head) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= nub) return;
  head[i] = (i == 0 || (keys[i] >> 8) != (keys[i - 1] >> 8)) ? 1 : 0;
}
// tile_ptr[t] = first sorted position of tile t (rank = inclusive scan of the heads), tile_ptr[ntiles] = nub
__global__ void __launch_bounds__(TPB) k_tile_ptr(const int* __restrict__ head, const int* __restrict__ rank, int nub,
                                                  int* __restrict__ tile_ptr) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= nub) return;
  if (head[i]) tile_ptr[rank[i] - 1] = i;
  if (i == nub - 1) tile_ptr[rank[i]] = nub;
}

// count (fill == 0) or fill (fill == 1) the product lists of the upper blocks; one thread per local observation
__global__ void __launch_bounds__(TPB) k_products(const int* __restrict__ o_kf, const int* __restrict__ o_lm,
                                                  const int* __restrict__ lm_ptr, const int* __restrict__ pose_slot,
                                                  const unsigned* __restrict__ bitmap, const int* __restrict__ word_prefix,
                                                  const int* __restrict__ s_rowptr, const int* __restrict__ csr_u, int words,
                                                  int E, int fill, unsigned* __restrict__ counters,
                                                  const unsigned* __restrict__ u_prod_ptr, uint2* __restrict__ prod,
                                                  const unsigned char* __restrict__ covered = nullptr) {
  const long long e = (long long)blockIdx.x * TPB + threadIdx.x;
  if (e >= E) return;
  const int a = pose_slot[o_kf[e]];
  if (a < 0) return;
  const int lm = o_lm[e];
  const int beg = lm_ptr[lm], end = lm_ptr[lm + 1];
  for (int o = beg; o < end; o++) {
    const int b = pose_slot[o_kf[o]];
    if (b < a || (b == a && o != (int)e)) continue;
    const int u = csr_u[csr_pos(bitmap, word_prefix, s_rowptr, words, a, b)];
    if (covered != nullptr && b != a && covered[u]) continue;  // the panel kernel forms this block; diagonal lists also feed the pose pass
    const unsigned slot = atomicAdd(counters + u, 1u);
    if (fill) prod[u_prod_ptr[u] + slot] = make_uint2((unsigned)e, (unsigned)o);
  }
}

// validation of the caller's observation arrays: flags[0] |= out-of-range / negative weight, flags[1] |= not grouped by landmark
__global__ void __launch_bounds__(TPB) k_check_obs(const int* __restrict__ kf, const int* __restrict__ mp,
                                                  const float* __restrict__ w, int E, int K, int P, int* __restrict__ flags) {
  int bad = 0, unsorted = 0;
  for (long long e = (long long)blockIdx.x * TPB + threadIdx.x; e < E; e += (long long)gridDim.x * TPB) {
    const int k = kf[e], m = mp[e];
    if (k < 0 || k >= K || m < 0 || m >= P || (w != nullptr && !(w[e] >= 0.0f))) bad = 1;
    if (e > 0 && mp[e - 1] > m) unsorted = 1;
  }
  if (bad) atomicOr(flags, 1);
  if (unsorted) atomicOr(flags + 1, 1);
}

// lm_ptr[l] = first observation of landmark l in the landmark-sorted observation list (lower bound), lm_ptr[P] = E
__global__ void __launch_bounds__(TPB) k_lm_ptr(const int* __restrict__ mp, int E, int P, int* __restrict__ lm_ptr) {
  const int l = blockIdx.x * TPB + threadIdx.x;
  if (l > P) return;
  int lo = 0, hi = E;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (mp[mid] < l) lo = mid + 1; else hi = mid;
  }
  lm_ptr[l] = lo;
}

__global__ void __launch_bounds__(TPB) k_apply_flags(const float* __restrict__ w_raw, const uint8_t* __restrict__ flags,
                                                     int E, float* __restrict__ o_w) {
  const long long e = (long long)blockIdx.x * TPB + threadIdx.x;
  if (e >= E) return;
  const uint8_t f = flags ? flags[e] : 0;
  float w = (f & 1) ? 0.0f : w_raw[e];
  if (f & 2) w = -w;  // -0.0f keeps the sign bit: inactive + no kernel
  o_w[e] = w;
}

}  // namespace ba
}  // namespace ccm

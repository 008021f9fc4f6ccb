// ba_math.cuh — SE3 / projection arithmetic shared by the BA kernels (device) and their host-side checks.
//
// What is computed follows the reference's edge and vertex types:
//   EdgeSE3ProjectXYZ::computeError / linearizeOplus   G/types/types_six_dof_expmap.{h:90-95,cpp:103-139}
//   VertexSE3Expmap::oplusImpl = exp(d) * T            G/types/types_six_dof_expmap.h:73-76, G/types/se3quat.h:223-257
//   RobustKernelHuber::robustify                       G/core/robust_kernel_impl.cpp:77-91
// (G/ = cslam/thirdparty/g2o/g2o/ of the reference).  Written from those definitions for a register-resident
// per-observation thread: no matrices in memory, rotation applied straight from the quaternion.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define CCM_HD __host__ __device__ __forceinline__
#else
#define CCM_HD inline
#endif

namespace ccm {

struct Pose {  // Tcw : Xc = R(q) X + t
  double qx, qy, qz, qw, tx, ty, tz;
};

CCM_HD void quat_rotate(double qx, double qy, double qz, double qw, double vx, double vy, double vz,
                        double& ox, double& oy, double& oz) {
  // v + 2 w (q x v) + 2 q x (q x v)
  double ux = qy * vz - qz * vy, uy = qz * vx - qx * vz, uz = qx * vy - qy * vx;
  ux += ux; uy += uy; uz += uz;
  ox = vx + qw * ux + (qy * uz - qz * uy);
  oy = vy + qw * uy + (qz * ux - qx * uz);
  oz = vz + qw * uz + (qx * uy - qy * ux);
}

CCM_HD void quat_to_R(double x, double y, double z, double w, double R[9]) {
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// rotation matrix -> quaternion with the branch structure of Eigen's Quaterniond(Matrix3d)
CCM_HD void R_to_quat(const double R[9], double& x, double& y, double& z, double& w) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    w = 0.5 * t;
    t = 0.5 / t;
    x = (R[7] - R[5]) * t; y = (R[2] - R[6]) * t; z = (R[3] - R[1]) * t;
  } else if (R[0] >= R[4] && R[0] >= R[8]) {  // i = 0
    t = sqrt(R[0] - R[4] - R[8] + 1.0);
    x = 0.5 * t; t = 0.5 / t;
    w = (R[7] - R[5]) * t; y = (R[3] + R[1]) * t; z = (R[6] + R[2]) * t;
  } else if (R[4] > R[0] && R[4] >= R[8]) {  // i = 1
    t = sqrt(R[4] - R[8] - R[0] + 1.0);
    y = 0.5 * t; t = 0.5 / t;
    w = (R[2] - R[6]) * t; z = (R[7] + R[5]) * t; x = (R[1] + R[3]) * t;
  } else {  // i = 2
    t = sqrt(R[8] - R[0] - R[4] + 1.0);
    z = 0.5 * t; t = 0.5 / t;
    w = (R[3] - R[1]) * t; x = (R[2] + R[6]) * t; y = (R[5] + R[7]) * t;
  }
}

CCM_HD void quat_normalize_pos_w(double& x, double& y, double& z, double& w) {
  if (w < 0) { x = -x; y = -y; z = -z; w = -w; }
  const double n = sqrt(x * x + y * y + z * z + w * w);  // Eigen's normalize(): coefficient-wise division by the norm
  x /= n; y /= n; z /= n; w /= n;
}

// out = exp(upd) * T with upd = (omega, upsilon).  Includes g2o's theta < 1e-5 branch (R = I + W + W^2, V = R).
CCM_HD Pose se3_exp_times(const double upd[6], const Pose& T) {
  const double ox = upd[0], oy = upd[1], oz = upd[2];
  const double theta2 = ox * ox + oy * oy + oz * oz;
  const double theta = sqrt(theta2);
  double a, b, va, vb;  // R = I + a W + b W^2 ; V = I + va W + vb W^2
  if (theta < 0.00001) {
    a = 1.0; b = 1.0; va = 1.0; vb = 1.0;
  } else {
    const double s = sin(theta), c = cos(theta);
    a = s / theta;
    b = (1 - c) / theta2;
    va = b;
    vb = (theta - s) / (theta2 * theta);
  }
  // W^2 = omega omega^T - theta^2 I
  double W2[9] = {ox * ox - theta2, ox * oy, ox * oz, ox * oy, oy * oy - theta2, oy * oz, ox * oz, oy * oz, oz * oz - theta2};
  double R[9], V[9];
  const double Wm[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    R[i] = id + a * Wm[i] + b * W2[i];
    V[i] = id + va * Wm[i] + vb * W2[i];
  }
  double ex, ey, ez, ew;
  R_to_quat(R, ex, ey, ez, ew);
  quat_normalize_pos_w(ex, ey, ez, ew);
  const double etx = V[0] * upd[3] + V[1] * upd[4] + V[2] * upd[5];
  const double ety = V[3] * upd[3] + V[4] * upd[4] + V[5] * upd[5];
  const double etz = V[6] * upd[3] + V[7] * upd[4] + V[8] * upd[5];
  Pose o;
  double rx, ry, rz;
  quat_rotate(ex, ey, ez, ew, T.tx, T.ty, T.tz, rx, ry, rz);
  o.tx = etx + rx; o.ty = ety + ry; o.tz = etz + rz;
  o.qw = ew * T.qw - ex * T.qx - ey * T.qy - ez * T.qz;
  o.qx = ew * T.qx + ex * T.qw + ey * T.qz - ez * T.qy;
  o.qy = ew * T.qy + ey * T.qw + ez * T.qx - ex * T.qz;
  o.qz = ew * T.qz + ez * T.qw + ex * T.qy - ey * T.qx;
  quat_normalize_pos_w(o.qx, o.qy, o.qz, o.qw);
  return o;
}

// Huber: returns rho(e) and the weight rho'(e)
CCM_HD void huber(double e, double delta, double& rho0, double& rho1) {
  const double dsqr = (double)(float)(delta * delta);  // the vendored kernel keeps delta^2 in a float (G/core/robust_kernel_impl.h:84)
  if (e <= dsqr) {
    rho0 = e; rho1 = 1.0;
  } else {
    const double sq = sqrt(e);
    rho0 = 2 * sq * delta - dsqr;
    rho1 = delta / sq;
  }
}

struct ObsLin {
  double ex, ey;    // residual  z - proj
  double chi2;      // w * |e|^2
  double Xc[3];
  double Jl[6];     // 2x3  d e / d point
  double Jp[12];    // 2x6  d e / d (omega, upsilon)
};

CCM_HD void project_residual(const Pose& T, const double intr[4], double X, double Y, double Z, double u, double v,
                             double w, double& ex, double& ey, double& chi2, double Xc[3]) {
  quat_rotate(T.qx, T.qy, T.qz, T.qw, X, Y, Z, Xc[0], Xc[1], Xc[2]);
  Xc[0] += T.tx; Xc[1] += T.ty; Xc[2] += T.tz;
  const double iz = 1.0 / Xc[2];
  ex = u - (Xc[0] * iz * intr[0] + intr[2]);
  ey = v - (Xc[1] * iz * intr[1] + intr[3]);
  chi2 = w * (ex * ex + ey * ey);
}

CCM_HD void linearize_obs(const Pose& T, const double intr[4], double X, double Y, double Z, double u, double v,
                          double w, ObsLin& L) {
  project_residual(T, intr, X, Y, Z, u, v, w, L.ex, L.ey, L.chi2, L.Xc);
  const double fx = intr[0], fy = intr[1];
  const double x = L.Xc[0], y = L.Xc[1], z = L.Xc[2];
  const double iz = 1.0 / z, iz2 = iz * iz;
  double R[9];
  quat_to_R(T.qx, T.qy, T.qz, T.qw, R);
  // Jl = -1/z * [[fx, 0, -x/z fx], [0, fy, -y/z fy]] * R
  const double a0 = -fx * iz, a2 = x * iz2 * fx;
  const double b1 = -fy * iz, b2 = y * iz2 * fy;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    L.Jl[j] = a0 * R[j] + a2 * R[6 + j];
    L.Jl[3 + j] = b1 * R[3 + j] + b2 * R[6 + j];
  }
  L.Jp[0] = x * y * iz2 * fx;
  L.Jp[1] = -(1 + (x * x * iz2)) * fx;
  L.Jp[2] = y * iz * fx;
  L.Jp[3] = -iz * fx;
  L.Jp[4] = 0;
  L.Jp[5] = x * iz2 * fx;
  L.Jp[6] = (1 + y * y * iz2) * fy;
  L.Jp[7] = -x * y * iz2 * fy;
  L.Jp[8] = -x * iz * fy;
  L.Jp[9] = 0;
  L.Jp[10] = -iz * fy;
  L.Jp[11] = y * iz2 * fy;
}

// upper Cholesky factor of a symmetric 3x3 D (d00 d01 d02 d11 d12 d22): D = U^T U
CCM_HD void chol3_upper(const double d[6], double u[6]) {
  u[0] = sqrt(d[0]);
  const double i0 = 1.0 / u[0];
  u[1] = d[1] * i0;
  u[2] = d[2] * i0;
  u[3] = sqrt(d[3] - u[1] * u[1]);
  const double i1 = 1.0 / u[3];
  u[4] = (d[4] - u[1] * u[2]) * i1;
  u[5] = sqrt(d[5] - u[2] * u[2] - u[4] * u[4]);
}
// solve z U = w for a row vector (3): z = w U^-1
CCM_HD void row_times_Uinv(const double u[6], double w0, double w1, double w2, double& z0, double& z1, double& z2) {
  z0 = w0 / u[0];
  z1 = (w1 - z0 * u[1]) / u[3];
  z2 = (w2 - z0 * u[2] - z1 * u[4]) / u[5];
}
// g = U^-T b  (solve U^T g = b)
CCM_HD void UTinv_times(const double u[6], const double b[3], double g[3]) {
  g[0] = b[0] / u[0];
  g[1] = (b[1] - u[1] * g[0]) / u[3];
  g[2] = (b[2] - u[2] * g[0] - u[4] * g[1]) / u[5];
}
// x = U^-1 g  (solve U x = g)
CCM_HD void Uinv_times(const double u[6], const double g[3], double x[3]) {
  x[2] = g[2] / u[5];
  x[1] = (g[1] - u[4] * x[2]) / u[3];
  x[0] = (g[0] - u[1] * x[1] - u[2] * x[2]) / u[0];
}

}  // namespace ccm

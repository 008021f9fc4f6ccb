// sim3_math.cuh — g2o::Sim3 arithmetic shared by the Sim3 pose-graph kernels (pgo.cu) and the single-vertex Sim3 alignment
// (single.cu).  Follows G/types/sim3.h of the reference (G/ = cslam/thirdparty/g2o/g2o/): exp ctor :70-142, log :148-230,
// inverse :233-236, operator* :266-272, map :144-146.
#pragma once
#include "ba_math.cuh"

namespace ccm {

struct S3 { double qx, qy, qz, qw, tx, ty, tz, s; };

CCM_HD S3 s3_load(const double* p) { return S3{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]}; }
CCM_HD void s3_store(const S3& a, double* p) { p[0] = a.qx; p[1] = a.qy; p[2] = a.qz; p[3] = a.qw; p[4] = a.tx; p[5] = a.ty; p[6] = a.tz; p[7] = a.s; }

CCM_HD S3 s3_mul(const S3& a, const S3& b) {  // G/types/sim3.h:266-272 (no quaternion normalisation)
  S3 r;
  r.qw = a.qw * b.qw - a.qx * b.qx - a.qy * b.qy - a.qz * b.qz;
  r.qx = a.qw * b.qx + a.qx * b.qw + a.qy * b.qz - a.qz * b.qy;
  r.qy = a.qw * b.qy + a.qy * b.qw + a.qz * b.qx - a.qx * b.qz;
  r.qz = a.qw * b.qz + a.qz * b.qw + a.qx * b.qy - a.qy * b.qx;
  double rx, ry, rz;
  quat_rotate(a.qx, a.qy, a.qz, a.qw, b.tx, b.ty, b.tz, rx, ry, rz);
  r.tx = a.s * rx + a.tx; r.ty = a.s * ry + a.ty; r.tz = a.s * rz + a.tz;
  r.s = a.s * b.s;
  return r;
}

CCM_HD S3 s3_inv(const S3& a) {  // G/types/sim3.h:233-236
  S3 r;
  r.qx = -a.qx; r.qy = -a.qy; r.qz = -a.qz; r.qw = a.qw;
  const double k = -1. / a.s;
  quat_rotate(r.qx, r.qy, r.qz, r.qw, k * a.tx, k * a.ty, k * a.tz, r.tx, r.ty, r.tz);
  r.s = 1. / a.s;
  return r;
}

// coefficients A, B, C of W = A*Omega + B*Omega^2 + C*I shared by exp and log (G/types/sim3.h:88-135, 166-208)
CCM_HD void s3_abc(double sigma, double s, double theta, bool small_theta, double& A, double& B, double& C) {
  const double eps = 0.00001;
  if (fabs(sigma) < eps) {
    C = 1;
    if (small_theta) { A = 1. / 2.; B = 1. / 6.; }
    else {
      const double theta2 = theta * theta;
      A = (1 - cos(theta)) / theta2;
      B = (theta - sin(theta)) / (theta2 * theta);
    }
  } else {
    C = (s - 1) / sigma;
    if (small_theta) {
      const double sigma2 = sigma * sigma;
      A = ((sigma - 1) * s + 1) / sigma2;
      B = ((0.5 * sigma2 - sigma + 1) * s) / (sigma2 * sigma);
    } else {
      const double a = s * sin(theta), b = s * cos(theta);
      const double theta2 = theta * theta, sigma2 = sigma * sigma;
      const double c = theta2 + sigma2;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / (c)) * 1. / (theta2);
    }
  }
}

CCM_HD S3 s3_exp(const double u[7]) {  // Sim3(Vector7d), G/types/sim3.h:70-142
  const double ox = u[0], oy = u[1], oz = u[2], sigma = u[6];
  const double theta2 = ox * ox + oy * oy + oz * oz, theta = sqrt(theta2);
  const double eps = 0.00001;
  S3 r;
  r.s = exp(sigma);
  double A, B, C;
  s3_abc(sigma, r.s, theta, theta < eps, A, B, C);
  double ra, rb;  // R = I + ra*Omega + rb*Omega^2
  if (theta < eps) { ra = 1.0; rb = 1.0; }
  else { ra = sin(theta) / theta; rb = (1 - cos(theta)) / (theta * theta); }
  const double Om[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
  const double O2[9] = {ox * ox - theta2, ox * oy, ox * oz, ox * oy, oy * oy - theta2, oy * oz, ox * oz, oy * oz, oz * oz - theta2};
  double R[9], W[9];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    R[i] = id + ra * Om[i] + rb * O2[i];
    W[i] = A * Om[i] + B * O2[i] + C * id;
  }
  R_to_quat(R, r.qx, r.qy, r.qz, r.qw);
  r.tx = W[0] * u[3] + W[1] * u[4] + W[2] * u[5];
  r.ty = W[3] * u[3] + W[4] * u[4] + W[5] * u[5];
  r.tz = W[6] * u[3] + W[7] * u[4] + W[8] * u[5];
  return r;
}

CCM_HD void s3_log(const S3& S, double res[7]) {  // Sim3::log, G/types/sim3.h:148-230
  const double sigma = log(S.s);
  double R[9];
  quat_to_R(S.qx, S.qy, S.qz, S.qw, R);
  const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double eps = 0.00001;
  const bool small_theta = d > 1 - eps;
  const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  double om[3], theta = 0.0;
  if (small_theta) {
    om[0] = 0.5 * dR[0]; om[1] = 0.5 * dR[1]; om[2] = 0.5 * dR[2];
  } else {
    theta = acos(d);
    const double f = theta / (2 * sqrt(1 - d * d));
    om[0] = f * dR[0]; om[1] = f * dR[1]; om[2] = f * dR[2];
  }
  double A, B, C;
  s3_abc(sigma, S.s, theta, small_theta, A, B, C);
  const double ox = om[0], oy = om[1], oz = om[2];
  const double n2 = ox * ox + oy * oy + oz * oz;
  const double Om[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
  const double O2[9] = {ox * ox - n2, ox * oy, ox * oz, ox * oy, oy * oy - n2, oy * oz, ox * oz, oy * oz, oz * oz - n2};
  double W[9];
#pragma unroll
  for (int i = 0; i < 9; i++) W[i] = A * Om[i] + B * O2[i] + C * ((i == 0 || i == 4 || i == 8) ? 1.0 : 0.0);
  // upsilon = W.lu().solve(t): Gaussian elimination with partial pivoting
  double y[3] = {S.tx, S.ty, S.tz};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int piv = k;
    double best = fabs(W[k * 3 + k]);
    for (int i = k + 1; i < 3; i++)
      if (fabs(W[i * 3 + k]) > best) { best = fabs(W[i * 3 + k]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < 3; j++) { const double t = W[k * 3 + j]; W[k * 3 + j] = W[piv * 3 + j]; W[piv * 3 + j] = t; }
      const double t = y[k]; y[k] = y[piv]; y[piv] = t;
    }
    for (int i = k + 1; i < 3; i++) {
      const double f = W[i * 3 + k] / W[k * 3 + k];
      for (int j = k + 1; j < 3; j++) W[i * 3 + j] -= f * W[k * 3 + j];
      y[i] -= f * y[k];
    }
  }
  double ups[3];
  ups[2] = y[2] / W[8];
  ups[1] = (y[1] - W[5] * ups[2]) / W[4];
  ups[0] = (y[0] - W[1] * ups[1] - W[2] * ups[2]) / W[0];
  res[0] = om[0]; res[1] = om[1]; res[2] = om[2];
  res[3] = ups[0]; res[4] = ups[1]; res[5] = ups[2];
  res[6] = sigma;
}

CCM_HD void edge_error(const S3& C, const S3& vi, const S3& vj, double e[7]) {
  s3_log(s3_mul(s3_mul(C, vi), s3_inv(vj)), e);
}

CCM_HD S3 s3_oplus(const S3& v, double u[7], int fix_scale) {
  if (fix_scale) u[6] = 0;
  return s3_mul(s3_exp(u), v);
}

CCM_HD void s3_map(const S3& S, double x, double y, double z, double& ox, double& oy, double& oz) {  // G/types/sim3.h:144-146
  double rx, ry, rz;
  quat_rotate(S.qx, S.qy, S.qz, S.qw, x, y, z, rx, ry, rz);
  ox = S.s * rx + S.tx; oy = S.s * ry + S.ty; oz = S.s * rz + S.tz;
}

}  // namespace ccm

// lie.hpp — SE3 / Sim3 / quaternion arithmetic of the CPU oracle (TEST INFRASTRUCTURE).
// Restates g2o's value types and the Eigen routines they call, in plain double arithmetic.
// G/ = /root/reference/cslam/thirdparty/g2o/g2o/
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

struct Quat { double x, y, z, w; };
struct SE3 { Quat r; double t[3]; };
struct Sim3 { Quat r; double t[3]; double s; };

// Eigen QuaternionBase::operator*(Quaternion): Hamilton product.
inline Quat qmul(const Quat& a, const Quat& b) {
  Quat q;
  q.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  q.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  q.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  q.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return q;
}
inline Quat qconj(const Quat& a) { return Quat{-a.x, -a.y, -a.z, a.w}; }

// Eigen QuaternionBase::_transformVector: v + w*(2 q x v) + q x (2 q x v)
inline void qrot(const Quat& q, const double v[3], double out[3]) {
  double ux = q.y * v[2] - q.z * v[1];
  double uy = q.z * v[0] - q.x * v[2];
  double uz = q.x * v[1] - q.y * v[0];
  ux += ux; uy += uy; uz += uz;
  out[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
  out[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
  out[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}

// Eigen QuaternionBase::toRotationMatrix (row-major R[9]).
inline void q2R(const Quat& q, double R[9]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen quaternionbase_assign_impl<Matrix3>: trace branch, else largest-diagonal branch.
inline Quat R2q(const double R[9]) {
  Quat q;
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R[7] - R[5]) * t;
    q.y = (R[2] - R[6]) * t;
    q.z = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    double c[3];
    c[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (R[k * 3 + j] - R[j * 3 + k]) * t;
    c[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    c[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    q.x = c[0]; q.y = c[1]; q.z = c[2];
  }
  return q;
}

// SE3Quat::normalizeRotation (G/types/se3quat.h:280-285)
inline void normalize_rotation(Quat& q) {
  if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

inline void skew(const double v[3], double m[9]) {  // G/types/se3_ops.hpp:27-38
  m[0] = 0;     m[1] = -v[2]; m[2] = v[1];
  m[3] = v[2];  m[4] = 0;     m[5] = -v[0];
  m[6] = -v[1]; m[7] = v[0];  m[8] = 0;
}
inline void mat3mul(const double a[9], const double b[9], double c[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
inline void mat3vec(const double a[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; i++) o[i] = a[i * 3] * v[0] + a[i * 3 + 1] * v[1] + a[i * 3 + 2] * v[2];
}

// SE3Quat::map (G/types/se3quat.h:217-220)
inline void se3_map(const SE3& T, const double x[3], double out[3]) {
  qrot(T.r, x, out);
  out[0] += T.t[0]; out[1] += T.t[1]; out[2] += T.t[2];
}

// SE3Quat::operator* (G/types/se3quat.h:105-111)
inline SE3 se3_mul(const SE3& a, const SE3& b) {
  SE3 r = a;
  double rt[3];
  qrot(a.r, b.t, rt);
  r.t[0] += rt[0]; r.t[1] += rt[1]; r.t[2] += rt[2];
  r.r = qmul(a.r, b.r);
  normalize_rotation(r.r);
  return r;
}

// SE3Quat::exp (G/types/se3quat.h:223-257) incl. the small-angle quirk R = I + W + W^2, V = R.
inline SE3 se3_exp(const double upd[6]) {
  const double* omega = upd;
  const double* ups = upd + 3;
  double theta = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  double Om[9], Om2[9], R[9], V[9];
  skew(omega, Om);
  mat3mul(Om, Om, Om2);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) { R[i] = I[i] + Om[i] + Om2[i]; V[i] = R[i]; }
  } else {
    double a = std::sin(theta) / theta;
    double b = (1 - std::cos(theta)) / (theta * theta);
    double c = (theta - std::sin(theta)) / (std::pow(theta, 3));
    for (int i = 0; i < 9; i++) { R[i] = I[i] + a * Om[i] + b * Om2[i]; V[i] = I[i] + b * Om[i] + c * Om2[i]; }
  }
  SE3 T;
  T.r = R2q(R);
  normalize_rotation(T.r);  // SE3Quat(Quaterniond, Vector3d) ctor, G/types/se3quat.h:62-64
  mat3vec(V, ups, T.t);
  return T;
}

// SE3Quat(Matrix3d R, Vector3d t) (G/types/se3quat.h:58-60)
inline SE3 se3_from_Rt(const double R[9], const double t[3]) {
  SE3 T;
  T.r = R2q(R);
  normalize_rotation(T.r);
  T.t[0] = t[0]; T.t[1] = t[1]; T.t[2] = t[2];
  return T;
}

inline SE3 se3_load(const double* p) { return SE3{Quat{p[0], p[1], p[2], p[3]}, {p[4], p[5], p[6]}}; }
inline void se3_store(const SE3& T, double* p) {
  p[0] = T.r.x; p[1] = T.r.y; p[2] = T.r.z; p[3] = T.r.w; p[4] = T.t[0]; p[5] = T.t[1]; p[6] = T.t[2];
}

// ---- Sim3 (G/types/sim3.h) ----
inline Sim3 sim3_load(const double* p) { return Sim3{Quat{p[0], p[1], p[2], p[3]}, {p[4], p[5], p[6]}, p[7]}; }
inline void sim3_store(const Sim3& S, double* p) {
  p[0] = S.r.x; p[1] = S.r.y; p[2] = S.r.z; p[3] = S.r.w; p[4] = S.t[0]; p[5] = S.t[1]; p[6] = S.t[2]; p[7] = S.s;
}

// Sim3(Vector7d) G/types/sim3.h:70-142
inline Sim3 sim3_exp(const double upd[7]) {
  const double* omega = upd;
  const double* ups = upd + 3;
  double sigma = upd[6];
  double theta = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  double Om[9], Om2[9], R[9];
  skew(omega, Om);
  Sim3 S;
  S.s = std::exp(sigma);
  mat3mul(Om, Om, Om2);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const double eps = 0.00001;
  double A, B, C;
  if (std::fabs(sigma) < eps) {
    C = 1;
    if (theta < eps) {
      A = 1. / 2.; B = 1. / 6.;
      for (int i = 0; i < 9; i++) R[i] = I[i] + Om[i] + Om2[i];
    } else {
      double theta2 = theta * theta;
      A = (1 - std::cos(theta)) / theta2;
      B = (theta - std::sin(theta)) / (theta2 * theta);
      double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
      for (int i = 0; i < 9; i++) R[i] = I[i] + a * Om[i] + b * Om2[i];
    }
  } else {
    C = (S.s - 1) / sigma;
    if (theta < eps) {
      double sigma2 = sigma * sigma;
      A = ((sigma - 1) * S.s + 1) / sigma2;
      B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
      for (int i = 0; i < 9; i++) R[i] = I[i] + Om[i] + Om2[i];
    } else {
      double ra = std::sin(theta) / theta, rb = (1 - std::cos(theta)) / (theta * theta);
      for (int i = 0; i < 9; i++) R[i] = I[i] + ra * Om[i] + rb * Om2[i];
      double a = S.s * std::sin(theta);
      double b = S.s * std::cos(theta);
      double theta2 = theta * theta;
      double sigma2 = sigma * sigma;
      double c = theta2 + sigma2;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / (c)) * 1. / (theta2);
    }
  }
  S.r = R2q(R);  // no normalisation in the reference
  double W[9];
  for (int i = 0; i < 9; i++) W[i] = A * Om[i] + B * Om2[i] + C * I[i];
  mat3vec(W, ups, S.t);
  return S;
}

// Eigen PartialPivLU<Matrix3d>::solve restated: Gaussian elimination with partial (row) pivoting.
inline void lu3_solve(const double Ain[9], const double b[3], double x[3]) {
  double A[9];
  std::memcpy(A, Ain, sizeof(A));
  double y[3] = {b[0], b[1], b[2]};
  for (int k = 0; k < 3; k++) {
    int piv = k;
    double best = std::fabs(A[k * 3 + k]);
    for (int i = k + 1; i < 3; i++)
      if (std::fabs(A[i * 3 + k]) > best) { best = std::fabs(A[i * 3 + k]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < 3; j++) { double tmp = A[k * 3 + j]; A[k * 3 + j] = A[piv * 3 + j]; A[piv * 3 + j] = tmp; }
      double tmp = y[k]; y[k] = y[piv]; y[piv] = tmp;
    }
    for (int i = k + 1; i < 3; i++) {
      double f = A[i * 3 + k] / A[k * 3 + k];
      A[i * 3 + k] = f;
      for (int j = k + 1; j < 3; j++) A[i * 3 + j] -= f * A[k * 3 + j];
    }
  }
  // forward (unit lower)
  for (int i = 1; i < 3; i++)
    for (int j = 0; j < i; j++) y[i] -= A[i * 3 + j] * y[j];
  // backward
  for (int i = 2; i >= 0; i--) {
    for (int j = i + 1; j < 3; j++) y[i] -= A[i * 3 + j] * x[j];
    x[i] = y[i] / A[i * 3 + i];
  }
}

inline void deltaR(const double R[9], double v[3]) {  // G/types/se3_ops.hpp:40-47
  v[0] = R[7] - R[5]; v[1] = R[2] - R[6]; v[2] = R[3] - R[1];
}

// Sim3::log G/types/sim3.h:148-230
inline void sim3_log(const Sim3& S, double res[7]) {
  double sigma = std::log(S.s);
  double omega[3], R[9], Om[9], Om2[9], dR[3];
  q2R(S.r, R);
  double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double eps = 0.00001;
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double A, B, C;
  deltaR(R, dR);
  if (std::fabs(sigma) < eps) {
    C = 1;
    if (d > 1 - eps) {
      for (int i = 0; i < 3; i++) omega[i] = 0.5 * dR[i];
      A = 1. / 2.; B = 1. / 6.;
    } else {
      double theta = std::acos(d);
      double theta2 = theta * theta;
      double f = theta / (2 * std::sqrt(1 - d * d));
      for (int i = 0; i < 3; i++) omega[i] = f * dR[i];
      A = (1 - std::cos(theta)) / theta2;
      B = (theta - std::sin(theta)) / (theta2 * theta);
    }
  } else {
    C = (S.s - 1) / sigma;
    if (d > 1 - eps) {
      double sigma2 = sigma * sigma;
      for (int i = 0; i < 3; i++) omega[i] = 0.5 * dR[i];
      A = ((sigma - 1) * S.s + 1) / (sigma2);
      B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
    } else {
      double theta = std::acos(d);
      double f = theta / (2 * std::sqrt(1 - d * d));
      for (int i = 0; i < 3; i++) omega[i] = f * dR[i];
      double theta2 = theta * theta;
      double a = S.s * std::sin(theta);
      double b = S.s * std::cos(theta);
      double c = theta2 + sigma * sigma;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / (c)) * 1. / (theta2);
    }
  }
  skew(omega, Om);
  // sim3.h:221 writes  A*Omega + B*Omega*Omega + C*I : C++ groups it (B*Omega)*Omega, unlike the exponential's B*Omega2
  // (found by compiling the reference's header, oracle/ref_g2o_wrap.cpp; a last-bit matter that the numeric Jacobians amplify)
  double BOm[9];
  for (int i = 0; i < 9; i++) BOm[i] = B * Om[i];
  mat3mul(BOm, Om, Om2);
  double W[9], ups[3];
  for (int i = 0; i < 9; i++) W[i] = A * Om[i] + Om2[i] + C * I[i];
  lu3_solve(W, S.t, ups);
  for (int i = 0; i < 3; i++) { res[i] = omega[i]; res[i + 3] = ups[i]; }
  res[6] = sigma;
}

inline Sim3 sim3_inv(const Sim3& S) {  // G/types/sim3.h:233-236
  Sim3 r;
  r.r = qconj(S.r);
  double tt[3] = {(-1. / S.s) * S.t[0], (-1. / S.s) * S.t[1], (-1. / S.s) * S.t[2]};
  qrot(r.r, tt, r.t);
  r.s = 1. / S.s;
  return r;
}

inline Sim3 sim3_mul(const Sim3& a, const Sim3& b) {  // G/types/sim3.h:266-272
  Sim3 r;
  r.r = qmul(a.r, b.r);
  double rt[3];
  qrot(a.r, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.s * rt[i] + a.t[i];
  r.s = a.s * b.s;
  return r;
}

inline void sim3_map(const Sim3& S, const double x[3], double out[3]) {  // G/types/sim3.h:144-146
  double r[3];
  qrot(S.r, x, r);
  for (int i = 0; i < 3; i++) out[i] = S.s * r[i] + S.t[i];
}

// Eigen Matrix3d::inverse, row-major in / out.  Eigen/src/LU/Inverse.h (compute_inverse<MatrixType, ResultType, 3>): the cofactors of
// column 0 give the determinant as (cofactors_col0 . col(0)).sum() — a three-term unrolled reduction, a0 + (a1 + a2) — and every
// entry is its cofactor times 1/det.  (Which association a given Eigen build uses for that sum is a property of the build; the
// reference's g2o sources compiled over oracle/ref_stub/Eigen — ref_ba_block_wrap.cpp — use this one.)
inline void inv3(const double m[9], double o[9]) {
  const double c0 = m[4] * m[8] - m[5] * m[7];      // cofactor(0,0)
  const double c1 = m[7] * m[2] - m[8] * m[1];      // cofactor(1,0)
  const double c2 = m[1] * m[5] - m[2] * m[4];      // cofactor(2,0)
  const double det = c0 * m[0] + (c1 * m[3] + c2 * m[6]);
  const double id = 1.0 / det;
  o[0] = c0 * id; o[1] = c1 * id; o[2] = c2 * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[8] * m[0] - m[6] * m[2]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[6] * m[1] - m[7] * m[0]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

}  // namespace orc

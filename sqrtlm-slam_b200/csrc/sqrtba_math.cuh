// Per-observation / per-vertex arithmetic of the BA hot path, written once as __host__ __device__
// inline functions so that the CUDA kernels and the CPU-side unit check (tests/cpu_math_check.cpp)
// execute the very same source.  Reference formulas (paths relative to /root/reference):
//   residuals   Thirdparty/g2o/g2o/types/types_six_dof_expmap.h:90-95,122-127, .cpp:141-157
//   Jacobians   Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:103-139 (mono), 188-234 (stereo)
//   Huber       Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91, base_edge.h:96-102
//   SE3 update  Thirdparty/g2o/g2o/types/se3quat.h:104-110,223-257,280-285, types_six_dof_expmap.h:73-76
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SQ_HD __host__ __device__ __forceinline__
#else
#define SQ_HD inline
#endif

namespace sqrtba {

// rotation matrix of a unit quaternion (x,y,z,w), row-major
SQ_HD void quat_to_R(const double q[4], double R[9]) {
  const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

// Xc = R X + t
SQ_HD void transform(const double R[9], const double t[3], const double X[3], double Xc[3]) {
  Xc[0] = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  Xc[1] = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  Xc[2] = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
}

// e = obs - cam_project(Xc).  Stereo keeps the reference's float32 inverse depth and float bf*invz product
// (types_six_dof_expmap.cpp:150-157); `ur < 0` selects the monocular edge (g2oOptimizer.cc:208, 877).
SQ_HD void reproj_error(const double Xc[3], float u, float v, float ur, const double cam[5], bool stereo,
                        double e[3]) {
  if (stereo) {
    const float invz = (float)(1.0 / Xc[2]);
    const float bfi = (float)cam[4] * invz;  // float * float, rounded to float
    const double pu = Xc[0] * (double)invz * cam[0] + cam[2];
    const double pv = Xc[1] * (double)invz * cam[1] + cam[3];
    e[0] = (double)u - pu;
    e[1] = (double)v - pv;
    e[2] = (double)ur - (pu - (double)bfi);
  } else {
    e[0] = (double)u - (Xc[0] / Xc[2] * cam[0] + cam[2]);
    e[1] = (double)v - (Xc[1] / Xc[2] * cam[1] + cam[3]);
    e[2] = 0.0;
  }
}

// Huber: rho0 (robustified chi2) and rho1 (weight); delta/dsqr as RobustKernelHuber::setDelta stores them -- the
// reference keeps dsqr in a FLOAT member (core/robust_kernel_impl.h:84), so callers pass huber_dsqr(delta)
SQ_HD double huber_dsqr(double delta) { return (double)(float)(delta * delta); }
SQ_HD void huber(double c, double delta, double dsqr, double* rho0, double* rho1) {
  if (c <= dsqr) {
    *rho0 = c;
    *rho1 = 1.0;
  } else {
    const double s = sqrt(c);
    *rho0 = 2.0 * s * delta - dsqr;
    *rho1 = delta / s;
  }
}

// Jacobians of e wrt the pose tangent (omega first, then upsilon; T <- exp(xi) T) and wrt the world point.
// Jp is 3x6 row-major, Jl is 3x3 row-major; the third rows are zero for a monocular edge.
SQ_HD void reproj_jacobians(const double R[9], const double Xc[3], const double cam[5], bool stereo, double Jp[18],
                            double Jl[9]) {
  const double x = Xc[0], y = Xc[1], z = Xc[2];
  const double iz = 1.0 / z, iz2 = iz * iz;
  const double fx = cam[0], fy = cam[1], bf = cam[4];
  Jp[0] = x * y * iz2 * fx;
  Jp[1] = -(1.0 + x * x * iz2) * fx;
  Jp[2] = y * iz * fx;
  Jp[3] = -iz * fx;
  Jp[4] = 0.0;
  Jp[5] = x * iz2 * fx;
  Jp[6] = (1.0 + y * y * iz2) * fy;
  Jp[7] = -x * y * iz2 * fy;
  Jp[8] = -x * iz * fy;
  Jp[9] = 0.0;
  Jp[10] = -iz * fy;
  Jp[11] = y * iz2 * fy;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    Jl[c] = -fx * R[c] * iz + fx * x * R[6 + c] * iz2;
    Jl[3 + c] = -fy * R[3 + c] * iz + fy * y * R[6 + c] * iz2;
  }
  if (stereo) {
    Jp[12] = Jp[0] - bf * y * iz2;
    Jp[13] = Jp[1] + bf * x * iz2;
    Jp[14] = Jp[2];
    Jp[15] = Jp[3];
    Jp[16] = 0.0;
    Jp[17] = Jp[5] - bf * iz2;
#pragma unroll
    for (int c = 0; c < 3; c++) Jl[6 + c] = Jl[c] - bf * R[6 + c] * iz2;
  } else {
#pragma unroll
    for (int c = 0; c < 6; c++) Jp[12 + c] = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) Jl[6 + c] = 0.0;
  }
}

// ---- the pose Jacobian as a function of four numbers per observation ------------------------------------------------
// The weighted 3x6 pose block of an edge depends on the camera-frame point only through x/z, y/z and 1/z (and on the
// keyframe's fx, fy, bf): 13 distinct non-zero entries out of 18.  The PCG operand therefore stores
//   g = { x/z, y/z, 1/z, w }      (w = sqrt(rho1 * invSigma2); an excluded edge stores four zeros)
// per observation -- 32 bytes instead of 144 -- and every kernel that needs the block rebuilds it in registers with the
// SAME function, so the gradient, the block-Jacobi blocks, the reduced right-hand side and the matvec all see identical
// values.  (types_six_dof_expmap.cpp:126-138, 188-234 written in these variables.)
struct JpC {  // the 13 distinct entries, named by their position in the 3x6 row-major block
  double j0, j1, j2, j3, j5, j6, j7, j8, j10, j11, j12, j13, j17;
};
SQ_HD JpC jp_compact(const double g[4], double fx, double fy, double bf, bool stereo) {
  const double a = g[0], b = g[1], iz = g[2], w = g[3];
  const double fxw = fx * w, fyw = fy * w;
  JpC J;
  J.j0 = a * b * fxw;
  J.j1 = -(1.0 + a * a) * fxw;
  J.j2 = b * fxw;
  J.j3 = -iz * fxw;
  J.j5 = a * iz * fxw;
  J.j6 = (1.0 + b * b) * fyw;
  J.j7 = -a * b * fyw;
  J.j8 = -a * fyw;
  J.j10 = -iz * fyw;
  J.j11 = b * iz * fyw;
  const double bw = stereo ? bf * w * iz : 0.0;
  J.j12 = stereo ? J.j0 - bw * b : 0.0;
  J.j13 = stereo ? J.j1 + bw * a : 0.0;
  J.j17 = stereo ? J.j5 - bw * iz : 0.0;
  return J;
}
// v = Jp * p (6-vector); the third row is zero for a monocular edge
SQ_HD void jp_mul(const JpC& J, bool stereo, const double p[6], double v[3]) {
  v[0] = J.j0 * p[0] + J.j1 * p[1] + J.j2 * p[2] + J.j3 * p[3] + J.j5 * p[5];
  v[1] = J.j6 * p[0] + J.j7 * p[1] + J.j8 * p[2] + J.j10 * p[4] + J.j11 * p[5];
  v[2] = stereo ? (J.j12 * p[0] + J.j13 * p[1] + J.j2 * p[2] + J.j3 * p[3] + J.j17 * p[5]) : 0.0;
}
// out = Jp^T * u
SQ_HD void jp_mulT(const JpC& J, bool stereo, const double u[3], double out[6]) {
  const double u02 = stereo ? u[0] + u[2] : u[0];
  const double u2 = stereo ? u[2] : 0.0;
  out[0] = J.j0 * u[0] + J.j6 * u[1] + J.j12 * u2;
  out[1] = J.j1 * u[0] + J.j7 * u[1] + J.j13 * u2;
  out[2] = J.j2 * u02 + J.j8 * u[1];
  out[3] = J.j3 * u02;
  out[4] = J.j10 * u[1];
  out[5] = J.j5 * u[0] + J.j11 * u[1] + J.j17 * u2;
}
// the full 3x6 row-major block (the landmark QR's block-Jacobi products, debug read-back)
SQ_HD void jp_full(const JpC& J, bool stereo, double Jp[18]) {
  Jp[0] = J.j0; Jp[1] = J.j1; Jp[2] = J.j2; Jp[3] = J.j3; Jp[4] = 0.0; Jp[5] = J.j5;
  Jp[6] = J.j6; Jp[7] = J.j7; Jp[8] = J.j8; Jp[9] = 0.0; Jp[10] = J.j10; Jp[11] = J.j11;
  Jp[12] = J.j12; Jp[13] = J.j13; Jp[14] = stereo ? J.j2 : 0.0; Jp[15] = stereo ? J.j3 : 0.0; Jp[16] = 0.0; Jp[17] = J.j17;
}
// Jacobian wrt the world point only (3x3 row-major; third row zero for a monocular edge) -- the Jl half of
// reproj_jacobians
SQ_HD void reproj_jacobian_point(const double R[9], const double Xc[3], const double cam[5], bool stereo, double Jl[9]) {
  const double x = Xc[0], y = Xc[1], z = Xc[2];
  const double iz = 1.0 / z, iz2 = iz * iz;
  const double fx = cam[0], fy = cam[1], bf = cam[4];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    Jl[c] = -fx * R[c] * iz + fx * x * R[6 + c] * iz2;
    Jl[3 + c] = -fy * R[3 + c] * iz + fy * y * R[6 + c] * iz2;
    Jl[6 + c] = stereo ? Jl[c] - bf * R[6 + c] * iz2 : 0.0;
  }
}

// ---- SE3 left update  T <- exp(xi) * T, state stored as (tx,ty,tz,qx,qy,qz,qw) ---------------------

SQ_HD void quat_mul(const double a[4], const double b[4], double c[4]) {  // (x,y,z,w)
  c[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  c[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  c[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  c[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}

SQ_HD void quat_normalize_pos_w(double q[4]) {  // SE3Quat::normalizeRotation
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

SQ_HD void R_to_quat(const double m[9], double q[4]) {  // Eigen's Quaterniond(Matrix3d) branches
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
    double qq[4];
    qq[i] = 0.5 * t;
    t = 0.5 / t;
    qq[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
    qq[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    qq[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
  }
}

// SE3Quat::exp including the reference's small-angle branch R = I + Om + Om^2, V = R (se3quat.h:237-243)
SQ_HD void se3_exp(const double xi[6], double t_out[3], double q_out[4]) {
  const double wx = xi[0], wy = xi[1], wz = xi[2];
  const double theta = sqrt(wx * wx + wy * wy + wz * wz);
  const double Om[9] = {0.0, -wz, wy, wz, 0.0, -wx, -wy, wx, 0.0};
  double Om2[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Om2[i * 3 + j] = Om[i * 3] * Om[j] + Om[i * 3 + 1] * Om[3 + j] + Om[i * 3 + 2] * Om[6 + j];
  double a, b, c;
  if (theta < 0.00001) {
    a = 1.0; b = 1.0; c = 1.0;  // R = I + Om + Om2 ; V = R
  } else {
    a = sin(theta) / theta;
    b = (1.0 - cos(theta)) / (theta * theta);
    c = (theta - sin(theta)) / (theta * theta * theta);
  }
  double R[9], V[9];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    R[i] = id + a * Om[i] + b * Om2[i];
    V[i] = (theta < 0.00001) ? R[i] : id + b * Om[i] + c * Om2[i];
  }
  R_to_quat(R, q_out);
  quat_normalize_pos_w(q_out);
#pragma unroll
  for (int i = 0; i < 3; i++) t_out[i] = V[i * 3] * xi[3] + V[i * 3 + 1] * xi[4] + V[i * 3 + 2] * xi[5];
}

// VertexSE3Expmap::oplusImpl: pose7 <- exp(xi) * pose7
SQ_HD void pose_oplus(double pose[7], const double xi[6]) {
  double te[3], qe[4];
  se3_exp(xi, te, qe);
  double Re[9];
  quat_to_R(qe, Re);
  const double t0 = pose[0], t1 = pose[1], t2 = pose[2];
  const double qo[4] = {pose[3], pose[4], pose[5], pose[6]};
  pose[0] = te[0] + Re[0] * t0 + Re[1] * t1 + Re[2] * t2;
  pose[1] = te[1] + Re[3] * t0 + Re[4] * t1 + Re[5] * t2;
  pose[2] = te[2] + Re[6] * t0 + Re[7] * t1 + Re[8] * t2;
  double qn[4];
  quat_mul(qe, qo, qn);
  quat_normalize_pos_w(qn);
  pose[3] = qn[0]; pose[4] = qn[1]; pose[5] = qn[2]; pose[6] = qn[3];
}

// Cholesky inverse of a symmetric positive definite 6x6 (row-major, full storage). Returns false if not SPD.
SQ_HD bool spd6_inverse(const double A[36], double Ainv[36]) {
  double L[36];
#pragma unroll
  for (int i = 0; i < 36; i++) L[i] = 0.0;
  for (int j = 0; j < 6; j++) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k];
    if (!(d > 0.0)) return false;
    const double ljj = sqrt(d);
    L[j * 6 + j] = ljj;
    for (int i = j + 1; i < 6; i++) {
      double s = A[i * 6 + j];
      for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k];
      L[i * 6 + j] = s / ljj;
    }
  }
  // invert L (lower) in place into Li, then Ainv = Li^T Li
  double Li[36];
#pragma unroll
  for (int i = 0; i < 36; i++) Li[i] = 0.0;
  for (int j = 0; j < 6; j++) {
    Li[j * 6 + j] = 1.0 / L[j * 6 + j];
    for (int i = j + 1; i < 6; i++) {
      double s = 0.0;
      for (int k = j; k < i; k++) s -= L[i * 6 + k] * Li[k * 6 + j];
      Li[i * 6 + j] = s / L[i * 6 + i];
    }
  }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j <= i; j++) {
      double s = 0.0;
      for (int k = i; k < 6; k++) s += Li[k * 6 + i] * Li[k * 6 + j];
      Ainv[i * 6 + j] = s;
      Ainv[j * 6 + i] = s;
    }
  return true;
}

}  // namespace sqrtba

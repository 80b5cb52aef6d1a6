// Sim3 arithmetic of the essential-graph optimisation (g2oOptimizer.cc:1212-1558), written once
// as __host__ __device__ code like sqrtba_math.cuh: g2o::Sim3 (Thirdparty/g2o/g2o/types/sim3.h:40-285),
// VertexSim3Expmap::oplusImpl and EdgeSim3::computeError (types/types_seven_dof_expmap.h:48-110).
//
// The arithmetic only: it is compiled for the host by tests/cpu_math_check.cpp and checked against vectors recorded from
// the reference's own binary (tests/golden/libg2o_vectors.npz: sim3_*).  Used by csrc/sqrtba_posegraph.cuh (essential
// graph), csrc/sqrtba_sim3opt.cuh (OptimizeSim3) and the host adapter.
//
// A Sim3 is stored as 8 doubles in g2o's operator[] order: qx qy qz qw | tx ty tz | s.
#pragma once
#include "sqrtba_math.cuh"

namespace sqrtba {

SQ_HD void sim3_rotate(const double q[4], const double v[3], double out[3]) {  // Eigen: v + w*uv + q.vec x uv, uv = 2 q.vec x v
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  out[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
  out[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
  out[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}

// the coefficients A, B, C of W = A Om + B Om^2 + C I shared by exp and log (sim3.h:86-131, 172-215)
SQ_HD void sim3_abc(double theta, double sigma, double s, bool small_theta, double* A, double* B, double* C) {
  const double eps = 0.00001;
  if (fabs(sigma) < eps) {
    *C = 1;
    if (small_theta) { *A = 1. / 2.; *B = 1. / 6.; }
    else {
      const double theta2 = theta * theta;
      *A = (1 - cos(theta)) / theta2;
      *B = (theta - sin(theta)) / (theta2 * theta);
    }
  } else {
    *C = (s - 1) / sigma;
    const double sigma2 = sigma * sigma;
    if (small_theta) {
      *A = ((sigma - 1) * s + 1) / sigma2;
      *B = ((0.5 * sigma2 - sigma + 1) * s) / (sigma2 * sigma);
    } else {
      const double a = s * sin(theta), b = s * cos(theta), theta2 = theta * theta, c = theta2 + sigma2;
      *A = (a * sigma + (1 - b) * theta) / (theta * c);
      *B = (*C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
    }
  }
}

SQ_HD void sim3_skew2(const double om[3], double Om[9], double Om2[9]) {
  const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  for (int i = 0; i < 9; i++) Om[i] = O[i];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Om2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
}

// Sim3(const Vector7d&): exponential map of (omega, upsilon, sigma)   (sim3.h:68-144)
SQ_HD void sim3_exp(const double u[7], double out[8]) {
  const double theta = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  double Om[9], Om2[9], R[9], A, B, C;
  sim3_skew2(u, Om, Om2);
  const double s = exp(u[6]);
  const bool small = theta < 0.00001;
  sim3_abc(theta, u[6], s, small, &A, &B, &C);
  for (int i = 0; i < 9; i++) {
    const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    R[i] = small ? id + Om[i] + Om2[i] : id + sin(theta) / theta * Om[i] + (1 - cos(theta)) / (theta * theta) * Om2[i];
  }
  R_to_quat(R, out);
  for (int i = 0; i < 3; i++) {
    double acc = 0;
    for (int j = 0; j < 3; j++) {
      const double id = (i == j) ? 1.0 : 0.0;
      // W = A Om + B Om2 + C I, accumulated in Eigen's left-to-right order of the row product
      const double w = A * Om[i * 3 + j] + B * Om2[i * 3 + j] + C * id;
      acc = (j == 0) ? w * u[3 + j] : acc + w * u[3 + j];
    }
    out[4 + i] = acc;
  }
  out[7] = s;
}

// Sim3::log (sim3.h:150-235) incl. Matrix3d::lu().solve: 3x3 LU with partial pivoting
SQ_HD void sim3_log(const double S[8], double res[7]) {
  const double sigma = log(S[7]);
  double R[9];
  quat_to_R(S, R);
  const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  const bool small = d > 1 - 0.00001;
  double om[3], theta = 0, A, B, C;
  if (small) {
    for (int i = 0; i < 3; i++) om[i] = 0.5 * dR[i];
  } else {
    theta = acos(d);
    for (int i = 0; i < 3; i++) om[i] = theta / (2 * sqrt(1 - d * d)) * dR[i];
  }
  sim3_abc(theta, sigma, S[7], small, &A, &B, &C);
  double Om[9], Om2[9], W[9];
  sim3_skew2(om, Om, Om2);
  for (int i = 0; i < 9; i++) W[i] = A * Om[i] + B * Om2[i] + C * ((i == 0 || i == 4 || i == 8) ? 1.0 : 0.0);
  int perm[3] = {0, 1, 2};
  for (int k = 0; k < 3; k++) {
    int piv = k;
    for (int i = k + 1; i < 3; i++)
      if (fabs(W[i * 3 + k]) > fabs(W[piv * 3 + k])) piv = i;
    if (piv != k) {
      for (int j = 0; j < 3; j++) { const double t = W[k * 3 + j]; W[k * 3 + j] = W[piv * 3 + j]; W[piv * 3 + j] = t; }
      const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    for (int i = k + 1; i < 3; i++) {
      W[i * 3 + k] /= W[k * 3 + k];
      for (int j = k + 1; j < 3; j++) W[i * 3 + j] -= W[i * 3 + k] * W[k * 3 + j];
    }
  }
  double y[3], x[3];
  for (int i = 0; i < 3; i++) {
    y[i] = S[4 + perm[i]];
    for (int j = 0; j < i; j++) y[i] -= W[i * 3 + j] * y[j];
  }
  for (int i = 2; i >= 0; i--) {
    x[i] = y[i];
    for (int j = i + 1; j < 3; j++) x[i] -= W[i * 3 + j] * x[j];
    x[i] /= W[i * 3 + i];
  }
  for (int i = 0; i < 3; i++) { res[i] = om[i]; res[3 + i] = x[i]; }
  res[6] = sigma;
}

// Sim3::operator* and inverse (sim3.h:238-278); no quaternion re-normalisation, as in the reference
SQ_HD void sim3_mul(const double a[8], const double b[8], double out[8]) {
  double q[4], rt[3];
  quat_mul(a, b, q);
  sim3_rotate(a, b + 4, rt);
  for (int i = 0; i < 4; i++) out[i] = q[i];
  for (int i = 0; i < 3; i++) out[4 + i] = a[7] * rt[i] + a[4 + i];
  out[7] = a[7] * b[7];
}
SQ_HD void sim3_inv(const double a[8], double out[8]) {
  const double qi[4] = {-a[0], -a[1], -a[2], a[3]};
  const double v[3] = {(-1. / a[7]) * a[4], (-1. / a[7]) * a[5], (-1. / a[7]) * a[6]};
  double t[3];
  sim3_rotate(qi, v, t);
  for (int i = 0; i < 4; i++) out[i] = qi[i];
  for (int i = 0; i < 3; i++) out[4 + i] = t[i];
  out[7] = 1. / a[7];
}

// VertexSim3Expmap::oplusImpl (types_seven_dof_expmap.h:63-72): est <- Sim3(update) * est, scale frozen if fix_scale
SQ_HD void sim3_oplus(double est[8], const double upd[7], bool fix_scale) {
  double u[7], e[8], r[8];
  for (int i = 0; i < 7; i++) u[i] = upd[i];
  if (fix_scale) u[6] = 0;
  sim3_exp(u, e);
  sim3_mul(e, est, r);
  for (int i = 0; i < 8; i++) est[i] = r[i];
}

// EdgeSim3::computeError (types_seven_dof_expmap.h:95-103): log(C * v1 * v2^-1)
SQ_HD void sim3_edge_error(const double C8[8], const double v1[8], const double v2[8], double err[7]) {
  double a[8], b[8], c[8];
  sim3_mul(C8, v1, a);
  sim3_inv(v2, b);
  sim3_mul(a, b, c);
  sim3_log(c, err);
}

// EdgeSim3's Jacobians are the numeric ones it inherits (BaseBinaryEdge::linearizeOplus, base_binary_edge.hpp:122-195):
// per non-fixed vertex and per tangent direction d, oplus(+delta e_d) and oplus(-delta e_d) on that vertex (through
// oplusImpl, so a fixed scale zeroes column 6), column d = (e(+) - e(-)) / (2 delta), delta = 1e-9.  Ji, Jj: 7x7
// row-major (row = error component); the Jacobian of a fixed vertex is left untouched.
SQ_HD void sim3_edge_linearize(const double C8[8], const double v1[8], const double v2[8], bool fixed1, bool fixed2,
                               bool fix_scale, double Ji[49], double Jj[49]) {
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  for (int side = 0; side < 2; side++) {
    if (side == 0 ? fixed1 : fixed2) continue;
    double* J = side == 0 ? Ji : Jj;
    for (int d = 0; d < 7; d++) {
      double add[7] = {0, 0, 0, 0, 0, 0, 0}, p[8], e1[7], e2[7];
      add[d] = delta;
      for (int i = 0; i < 8; i++) p[i] = side == 0 ? v1[i] : v2[i];
      sim3_oplus(p, add, fix_scale);
      sim3_edge_error(C8, side == 0 ? p : v1, side == 0 ? v2 : p, e1);
      add[d] = -delta;
      for (int i = 0; i < 8; i++) p[i] = side == 0 ? v1[i] : v2[i];
      sim3_oplus(p, add, fix_scale);
      sim3_edge_error(C8, side == 0 ? p : v1, side == 0 ? v2 : p, e2);
      for (int r = 0; r < 7; r++) J[r * 7 + d] = scalar * (e1[r] - e2[r]);
    }
  }
}

}  // namespace sqrtba

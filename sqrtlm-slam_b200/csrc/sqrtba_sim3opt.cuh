// Sim3 alignment of loop-candidate keyframe pairs: the CUDA counterpart of g2oOptimizer::OptimizeSim3
// (src/backend/g2oOptimizer.cc:1560-1796), which LoopClosing::ComputeSim3 calls for every candidate that survives the
// RANSAC Sim3Solver (src/backend/LoopClosing.cc:513, th2 = 10).
//
// One CTA per keyframe pair, the whole schedule in ONE launch: optimize(5), chi2 test on the stored errors, optimize(10)
// if a match was dropped (else 5), final count.  The graph has ONE free vertex (the Sim3 S12 with both cameras'
// intrinsics, types_seven_dof_expmap.h:48-97) and two binary edges per match whose other end is a fixed point:
//   EdgeSim3ProjectXYZ        e12 = obs1 - cam_map1(project(S12.map(P2c)))            (:130-149)
//   EdgeInverseSim3ProjectXYZ e21 = obs2 - cam_map2(project(S12.inverse().map(P1c)))  (:152-171)
// Their Jacobians are the NUMERIC ones of BaseBinaryEdge::linearizeOplus (base_binary_edge.hpp:122-195): central
// differences with delta = 1e-9 through VertexSim3Expmap::oplusImpl.  The 14 perturbed similarities (and their inverses)
// are the same for every match, so 14 threads build them once per iteration and every thread then evaluates its
// matches against them.  The 7x7 system (LinearSolverDense: LDL^T, positive pivots or fail) and g2o's Levenberg policy
// (optimization_algorithm_levenberg.cpp:61-189) run on thread 0, as in csrc/sqrtba_poseopt.cuh.
#pragma once
#include "sqrtba_poseopt.cuh"
#include "sqrtba_sim3.cuh"

namespace sqrtba {

constexpr int S3_TRACE_COLS = 8;   // pass, iter, trial, lambda, chi_before, chi_trial, rho, accepted
constexpr int S3_MAX_TRACE = 150;  // (5 + 10) iterations x <= 10 trials
constexpr int S3_NV = 37;          // H (28 upper) | b (7) | chi2 | number of active matches

struct Sim3OptArgs {
  int n_pairs;
  const long long* pair_ptr;  // n_pairs + 1 match offsets
  double* s12;                // n_pairs x 8, in/out (qx qy qz qw | t | s)
  const double* cam8;         // n_pairs x 8: fx1 fy1 cx1 cy1 fx2 fy2 cx2 cy2
  const double* p1c;          // n_match x 3: matched map point of keyframe 1 in the frame of keyframe 1
  const double* p2c;          // n_match x 3: matched map point of keyframe 2 in the frame of keyframe 2
  const float* meas6;         // n_match x 6: u1 v1 invSigma2_1 u2 v2 invSigma2_2
  double* err;                // n_match x 4: g2o's stored _error of e12 | e21
  uint8_t* on;                // n_match: both edges still in the graph
  uint8_t* keep;              // n_match: vpMatches1 keeps the match
  int* n_in;                  // n_pairs: the function's return value
  double* trace;              // n_pairs x S3_MAX_TRACE x S3_TRACE_COLS
  int* trace_len;             // n_pairs
  float th2;
  int fix_scale;
};

// obs - cam_map(project(S.map(X)))
__device__ __forceinline__ void s3_error(const double* S, const double* X, double f0, double f1, double c0, double c1,
                                         double o0, double o1, double e[2]) {
  double r[3];
  sim3_rotate(S, X, r);
  const double x = S[7] * r[0] + S[4], y = S[7] * r[1] + S[5], z = S[7] * r[2] + S[6];
  e[0] = o0 - (x / z * f0 + c0);
  e[1] = o1 - (y / z * f1 + c1);
}

// (H + lambda I) x = b by LDL^T without pivoting; false when a pivot is not positive (LDLT::isPositive())
template <int N>
__device__ __forceinline__ bool ldlt_solve(const double* Hu /*upper, row-major*/, double lambda, const double* b, double* x) {
  double A[N * N], L[N * N], D[N], y[N];
  int idx = 0;
  for (int i = 0; i < N; i++)
    for (int j = i; j < N; j++) { A[i * N + j] = A[j * N + i] = Hu[idx++] + (i == j ? lambda : 0.0); }
  for (int i = 0; i < N * N; i++) L[i] = 0.0;
  for (int j = 0; j < N; j++) {
    double d = A[j * N + j];
    for (int k = 0; k < j; k++) d -= L[j * N + k] * L[j * N + k] * D[k];
    if (!(d > 0.0)) return false;
    D[j] = d;
    for (int i = j + 1; i < N; i++) {
      double v = A[i * N + j];
      for (int k = 0; k < j; k++) v -= L[i * N + k] * L[j * N + k] * D[k];
      L[i * N + j] = v / d;
    }
  }
  for (int i = 0; i < N; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[i * N + k] * y[k]; y[i] = v; }
  for (int i = 0; i < N; i++) y[i] /= D[i];
  for (int i = N - 1; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < N; k++) v -= L[k * N + i] * x[k]; x[i] = v; }
  return true;
}

__global__ void __launch_bounds__(PO_CTA) k_sim3_opt(Sim3OptArgs A) {
  __shared__ double sh_part[PO_WARPS * S3_NV];
  __shared__ double sh_red[S3_NV];
  __shared__ double sh_S[8], sh_Sinv[8], sh_bak[8], sh_x[7], sh_H[28], sh_b[7];
  __shared__ double sh_pert[14 * 8], sh_pinv[14 * 8];  // S (+) (+-delta e_d) and their inverses, index 2 d + (minus ? 1 : 0)
  __shared__ double sh_lambda, sh_ni, sh_cur, sh_ini;
  __shared__ int sh_flag[4];  // [0] again (retry with a larger lambda), [1] ok (next iteration), [2] ok2, [3] nBad of the LM
  const int p = blockIdx.x, tid = threadIdx.x;
  const long long o0 = A.pair_ptr[p], o1 = A.pair_ptr[p + 1];
  const int n = (int)(o1 - o0);
  const bool fix_scale = A.fix_scale != 0;
  double cam[8];
#pragma unroll
  for (int i = 0; i < 8; i++) cam[i] = A.cam8[p * 8 + i];
  if (tid < 8) sh_S[tid] = A.s12[p * 8 + tid];
  for (long long o = o0 + tid; o < o1; o += PO_CTA) {
    A.on[o] = 1;
    A.keep[o] = 1;
#pragma unroll
    for (int c = 0; c < 4; c++) A.err[o * 4 + c] = 0.0;
  }
  __syncthreads();
  const double th2 = (double)A.th2;
  const double hub = (double)sqrtf(A.th2), hub2 = huber_dsqr(hub);  // const float deltaHuber = sqrt(th2), :1619
  int trace_len = 0;                                                 // thread 0 only
  int iters = 5;
  for (int pass = 0; pass < 2; pass++) {
    if (tid == 0) sh_flag[1] = 1;
    __syncthreads();
    for (int iter = 0; iter < iters; iter++) {
      if (!sh_flag[1]) break;  // CTA-uniform (written before the last barrier of the previous iteration)
      // ---- the 14 perturbed estimates of the numeric Jacobian and the inverse of the estimate itself
      if (tid < 14) {
        double add[7] = {0, 0, 0, 0, 0, 0, 0}, q[8], qi[8];
        add[tid >> 1] = (tid & 1) ? -1e-9 : 1e-9;
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = sh_S[i];
        sim3_oplus(q, add, fix_scale);
        sim3_inv(q, qi);
#pragma unroll
        for (int i = 0; i < 8; i++) { sh_pert[tid * 8 + i] = q[i]; sh_pinv[tid * 8 + i] = qi[i]; }
      } else if (tid == 32) {
        double qi[8];
        sim3_inv(sh_S, qi);
#pragma unroll
        for (int i = 0; i < 8; i++) sh_Sinv[i] = qi[i];
      }
      __syncthreads();
      // ---- computeActiveErrors + activeRobustChi2 + buildSystem at the current estimate
      double acc[S3_NV];
#pragma unroll
      for (int k = 0; k < S3_NV; k++) acc[k] = 0.0;
      for (long long o = o0 + tid; o < o1; o += PO_CTA) {
        if (!A.on[o]) continue;
        const double* P1 = A.p1c + o * 3;
        const double* P2 = A.p2c + o * 3;
        const float* m = A.meas6 + o * 6;
        const double X1[3] = {P1[0], P1[1], P1[2]}, X2[3] = {P2[0], P2[1], P2[2]};
        const double obs[4] = {(double)m[0], (double)m[1], (double)m[3], (double)m[4]};
        const double info[2] = {(double)m[2], (double)m[5]};
#pragma unroll
        for (int side = 0; side < 2; side++) {  // 0: e12 (S, P2c, camera 1)   1: e21 (S^-1, P1c, camera 2)
          const double* X = side ? X1 : X2;
          const double f0 = cam[side * 4], f1 = cam[side * 4 + 1], c0 = cam[side * 4 + 2], c1 = cam[side * 4 + 3];
          double e[2], J[14];
          s3_error(side ? sh_Sinv : sh_S, X, f0, f1, c0, c1, obs[side * 2], obs[side * 2 + 1], e);
          A.err[o * 4 + side * 2] = e[0];
          A.err[o * 4 + side * 2 + 1] = e[1];
          for (int d = 0; d < 7; d++) {
            double ea[2], eb[2];
            s3_error((side ? sh_pinv : sh_pert) + (2 * d) * 8, X, f0, f1, c0, c1, obs[side * 2], obs[side * 2 + 1], ea);
            s3_error((side ? sh_pinv : sh_pert) + (2 * d + 1) * 8, X, f0, f1, c0, c1, obs[side * 2], obs[side * 2 + 1], eb);
            J[d] = (1.0 / (2 * 1e-9)) * (ea[0] - eb[0]);
            J[7 + d] = (1.0 / (2 * 1e-9)) * (ea[1] - eb[1]);
          }
          const double w0 = info[side];
          const double c2 = e[0] * (w0 * e[0]) + e[1] * (w0 * e[1]);
          double rho0, rho1;
          huber(c2, hub, hub2, &rho0, &rho1);
          acc[35] += rho0;
          const double w = rho1 * w0;
          int idx = 0;
#pragma unroll
          for (int i = 0; i < 7; i++) {
            acc[28 + i] -= rho1 * (J[i] * (w0 * e[0]) + J[7 + i] * (w0 * e[1]));
#pragma unroll
            for (int j = i; j < 7; j++) {
              acc[idx] += J[i] * (w * J[j]) + J[7 + i] * (w * J[7 + j]);
              idx++;
            }
          }
        }
        acc[36] += 1.0;
      }
      po_reduce<S3_NV>(acc, sh_part, sh_red);
      if (sh_red[36] == 0.0) break;  // no active edge: optimize() returns without touching the vertex (CTA-uniform)
      if (tid < 28) sh_H[tid] = sh_red[tid];
      if (tid < 7) sh_b[tid] = sh_red[28 + tid];
      if (tid == 0) {
        sh_cur = sh_red[35];
        sh_ini = sh_red[35];
        if (iter == 0) {  // computeLambdaInit, optimization_algorithm_levenberg.cpp:166-180
          double md = 0.0;
          int d = 0;
          for (int i = 0; i < 7; i++) { md = fmax(md, fabs(sh_red[d])); d += 7 - i; }
          sh_lambda = 1e-5 * md;
          sh_ni = 2.0;
          sh_flag[3] = 0;
        }
      }
      __syncthreads();
      // ---- trial loop (<= 10 trials)
      double rho = 0.0;  // thread 0
      for (int qmax = 0; qmax < 10; qmax++) {
        if (tid == 0) {
          for (int i = 0; i < 8; i++) sh_bak[i] = sh_S[i];
          double x[7] = {0, 0, 0, 0, 0, 0, 0};
          const bool ok2 = ldlt_solve<7>(sh_H, sh_lambda, sh_b, x);
          if (!ok2) for (int i = 0; i < 7; i++) x[i] = 0.0;
          for (int i = 0; i < 7; i++) sh_x[i] = x[i];
          double q[8], qi[8];
          for (int i = 0; i < 8; i++) q[i] = sh_S[i];
          sim3_oplus(q, x, fix_scale);
          sim3_inv(q, qi);
          for (int i = 0; i < 8; i++) { sh_S[i] = q[i]; sh_Sinv[i] = qi[i]; }
          sh_flag[2] = ok2 ? 1 : 0;
        }
        __syncthreads();
        double chi[1] = {0.0};
        for (long long o = o0 + tid; o < o1; o += PO_CTA) {
          if (!A.on[o]) continue;
          const float* m = A.meas6 + o * 6;
          double e[2], rho0, rho1;
          s3_error(sh_S, A.p2c + o * 3, cam[0], cam[1], cam[2], cam[3], (double)m[0], (double)m[1], e);
          A.err[o * 4] = e[0];  // a rejected trial leaves these behind (stale _error)
          A.err[o * 4 + 1] = e[1];
          huber(e[0] * ((double)m[2] * e[0]) + e[1] * ((double)m[2] * e[1]), hub, hub2, &rho0, &rho1);
          chi[0] += rho0;
          s3_error(sh_Sinv, A.p1c + o * 3, cam[4], cam[5], cam[6], cam[7], (double)m[3], (double)m[4], e);
          A.err[o * 4 + 2] = e[0];
          A.err[o * 4 + 3] = e[1];
          huber(e[0] * ((double)m[5] * e[0]) + e[1] * ((double)m[5] * e[1]), hub, hub2, &rho0, &rho1);
          chi[0] += rho0;
        }
        po_reduce<1>(chi, sh_part, sh_red);
        if (tid == 0) {
          double tempChi = sh_red[0];
          if (!sh_flag[2]) tempChi = 1.7976931348623157e308;
          rho = sh_cur - tempChi;
          double scale = 0.0;
          for (int j = 0; j < 7; j++) scale += sh_x[j] * (sh_lambda * sh_x[j] + sh_b[j]);
          scale += 1e-3;
          rho /= scale;
          const bool good = (rho > 0.0) && isfinite(tempChi);
          if (trace_len < S3_MAX_TRACE) {
            double* tr = A.trace + ((size_t)p * S3_MAX_TRACE + trace_len) * S3_TRACE_COLS;
            tr[0] = pass; tr[1] = iter; tr[2] = qmax; tr[3] = sh_lambda; tr[4] = sh_cur; tr[5] = tempChi; tr[6] = rho;
            tr[7] = good ? 1.0 : 0.0;
            trace_len++;
          }
          if (good) {
            double alpha = 1.0 - pow((2.0 * rho - 1.0), 3.0);
            alpha = fmin(alpha, 2.0 / 3.0);
            sh_lambda *= fmax(1.0 / 3.0, alpha);
            sh_ni = 2.0;
            sh_cur = tempChi;
          } else {
            sh_lambda *= sh_ni;
            sh_ni *= 2.0;
            for (int i = 0; i < 8; i++) sh_S[i] = sh_bak[i];
          }
          const bool again = (rho < 0.0) && (qmax + 1 < 10);
          sh_flag[0] = again ? 1 : 0;
          if (!again) {  // exit logic of OptimizationAlgorithmLevenberg::solve, :156-163
            bool ok = true;
            if (qmax + 1 == 10 || rho == 0.0) ok = false;
            else {
              if ((sh_ini - sh_cur) * 1e3 < sh_ini) sh_flag[3]++; else sh_flag[3] = 0;
              if (sh_flag[3] >= 3) ok = false;
            }
            sh_flag[1] = ok ? 1 : 0;
          }
        }
        __syncthreads();
        if (!sh_flag[0]) break;
      }
      __syncthreads();
    }
    __syncthreads();
    // ---- chi2 test on the stored errors (:1724-1745 after pass 0, :1765-1779 after pass 1)
    double cnt[2] = {0.0, 0.0};  // dropped now | still on
    for (long long o = o0 + tid; o < o1; o += PO_CTA) {
      if (!A.on[o]) continue;
      const float* m = A.meas6 + o * 6;
      const double a0 = A.err[o * 4], a1 = A.err[o * 4 + 1], b0 = A.err[o * 4 + 2], b1 = A.err[o * 4 + 3];
      const double c12 = a0 * ((double)m[2] * a0) + a1 * ((double)m[2] * a1);
      const double c21 = b0 * ((double)m[5] * b0) + b1 * ((double)m[5] * b1);
      if (c12 > th2 || c21 > th2) {
        A.keep[o] = 0;
        if (pass == 0) A.on[o] = 0;  // removeEdge; after pass 1 only the match is nulled
        cnt[0] += 1.0;
      } else {
        cnt[1] += 1.0;
      }
    }
    po_reduce<2>(cnt, sh_part, sh_red);
    const int n_drop = (int)(sh_red[0] + 0.5), n_left = (int)(sh_red[1] + 0.5);
    __syncthreads();  // sh_red is reused by the next reduction
    if (pass == 0) {
      if (n - n_drop < 10) {  // nCorrespondences - nBad < 10: return 0, g2oS12 untouched (:1755-1756)
        if (tid == 0) { A.n_in[p] = 0; A.trace_len[p] = trace_len; }
        return;
      }
      iters = n_drop > 0 ? 10 : 5;
    } else {
      if (tid == 0) { A.n_in[p] = n_left; A.trace_len[p] = trace_len; }
      if (tid < 8) A.s12[p * 8 + tid] = sh_S[tid];
    }
  }
}

}  // namespace sqrtba

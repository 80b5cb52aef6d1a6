// Pose-only optimisation: the CUDA counterpart of g2oOptimizer::PoseOptimization
// (src/backend/g2oOptimizer.cc:385-559, 655-690; the lidar block :560-640 is csrc/sqrtba_poseopt_lidar.cuh), the per-frame sibling of
// local BA that the tracking thread calls on every frame (Tracking.cc:1347,1547,1617,2466-2517).
//
// One CTA per frame, the whole 4 x optimize(10) schedule in ONE launch: every LM trial is a sweep of the CTA over the
// frame's observations (residual, 2x6 / 3x6 Jacobian, Huber weight) reduced to the 6x6 system with warp shuffles, one
// thread solves (H + lambda I) x = b (LDL^T, the analogue of LinearSolverDense, linear_solver_dense.h:65-113) and runs
// g2o's Levenberg policy (optimization_algorithm_levenberg.cpp:61-189), then the chi2 re-classification of
// g2oOptimizer.cc:518-547.  A batch of frames (relocalisation candidates, Tracking.cc:2466-2517) is one launch.
// Edge arithmetic: EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose, types_six_dof_expmap.h:143-202,
// .cpp:266-306, 335-364 (note: float invz but DOUBLE bf in the stereo projection, unlike the binary stereo edge).
#pragma once
#include "sqrtba_kernels.cuh"

namespace sqrtba {

constexpr int PO_CTA = 256;
constexpr int PO_WARPS = PO_CTA / 32;
constexpr int PO_TRACE_COLS = 8;   // round, iter, trial, lambda, chi_before, chi_trial, rho, accepted
constexpr int PO_MAX_TRACE = 500;  // (4 rounds + the lidar round) x 10 iterations x <= 10 trials

struct PoseOptArgs {
  int n_frames;
  const long long* frame_ptr;  // n_frames + 1 observation offsets
  double* pose;                // n_frames x 7, in/out
  const double* cam;           // n_frames x 5
  const double* xyz;           // n_obs x 3
  const float4* meas;          // n_obs: u, v, ur (<0 mono), invSigma2
  double* err;                 // n_obs x 3: g2o's stored _error
  uint8_t* level;              // n_obs: 0 active, 1 excluded
  uint8_t* outlier;            // n_obs: Frame::mvbOutlier
  int* inliers;                // n_frames
  double* trace;               // n_frames x PO_MAX_TRACE x PO_TRACE_COLS
  int* trace_len;              // n_frames
  int skip_final;              // 1: stop after the rounds (pose, err, level, outlier stay for k_pose_opt_lidar)
};

struct PoEdge {
  double e[3], J[18], info;
  bool stereo;
};

// residual (and Jacobian) of one observation at pose (R, t)
__device__ __forceinline__ void po_eval(const double R[9], const double t[3], const double cam[5], const double* X,
                                        const float4 m, bool want_jac, PoEdge& E) {
  double Xc[3];
  transform(R, t, X, Xc);
  E.stereo = !(m.z < 0.0f);
  E.info = (double)m.w;
  if (!E.stereo) {
    const double px = Xc[0] / Xc[2], py = Xc[1] / Xc[2];
    E.e[0] = (double)m.x - (px * cam[0] + cam[2]);
    E.e[1] = (double)m.y - (py * cam[1] + cam[3]);
    E.e[2] = 0.0;
  } else {
    const float invz = (float)(1.0f / Xc[2]);
    const double r0 = Xc[0] * invz * cam[0] + cam[2], r1 = Xc[1] * invz * cam[1] + cam[3];
    E.e[0] = (double)m.x - r0;
    E.e[1] = (double)m.y - r1;
    E.e[2] = (double)m.z - (r0 - cam[4] * invz);
  }
  if (want_jac) {
    const double x = Xc[0], y = Xc[1], invz = 1.0 / Xc[2], invz_2 = invz * invz, fx = cam[0], fy = cam[1];
    double* J = E.J;
    J[0] = x * y * invz_2 * fx;
    J[1] = -(1 + (x * x * invz_2)) * fx;
    J[2] = y * invz * fx;
    J[3] = -invz * fx;
    J[4] = 0;
    J[5] = x * invz_2 * fx;
    J[6] = (1 + y * y * invz_2) * fy;
    J[7] = -x * y * invz_2 * fy;
    J[8] = -x * invz * fy;
    J[9] = 0;
    J[10] = -invz * fy;
    J[11] = y * invz_2 * fy;
    if (E.stereo) {
      J[12] = J[0] - cam[4] * y * invz_2;
      J[13] = J[1] + cam[4] * x * invz_2;
      J[14] = J[2];
      J[15] = J[3];
      J[16] = 0;
      J[17] = J[5] - cam[4] * invz_2;
    } else {
#pragma unroll
      for (int c = 0; c < 6; c++) J[12 + c] = 0;
    }
  }
}

__device__ __forceinline__ double po_chi2(const double e[3], double info, bool stereo) {
  double c = e[0] * (info * e[0]) + e[1] * (info * e[1]);
  if (stereo) c += e[2] * (info * e[2]);
  return c;
}

// sum NV values over the CTA; result in sh_out[0..NV) (valid after the trailing barrier)
template <int NV>
__device__ __forceinline__ void po_reduce(double* v, double* sh_part, double* sh_out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double x = v[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(FULL, x, off);
    if (lane == 0) sh_part[wid * NV + k] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < PO_WARPS; w++) s += sh_part[w * NV + threadIdx.x];
    sh_out[threadIdx.x] = s;
  }
  __syncthreads();
}

// (H + lambda I) x = b by LDL^T without pivoting; false when a pivot is not positive (LDLT::isPositive())
__device__ __forceinline__ bool po_solve6(const double* Hu /*21 upper, row-major*/, double lambda, const double* b, double* x) {
  double A[36], L[36], D[6];
  int idx = 0;
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++) { A[i * 6 + j] = A[j * 6 + i] = Hu[idx++] + (i == j ? lambda : 0.0); }
  for (int i = 0; i < 36; i++) L[i] = 0.0;
  for (int j = 0; j < 6; j++) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
    if (!(d > 0.0)) return false;
    D[j] = d;
    for (int i = j + 1; i < 6; i++) {
      double v = A[i * 6 + j];
      for (int k = 0; k < j; k++) v -= L[i * 6 + k] * L[j * 6 + k] * D[k];
      L[i * 6 + j] = v / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[i * 6 + k] * y[k]; y[i] = v; }
  for (int i = 0; i < 6; i++) y[i] /= D[i];
  for (int i = 5; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < 6; k++) v -= L[k * 6 + i] * x[k]; x[i] = v; }
  return true;
}

__global__ void __launch_bounds__(PO_CTA) k_pose_opt(PoseOptArgs A) {
  __shared__ double sh_part[PO_WARPS * 29];
  __shared__ double sh_red[29];
  __shared__ double sh_pose[7], sh_pose0[7], sh_bak[7], sh_x[6], sh_H[21], sh_b[6];
  __shared__ double sh_lambda, sh_ni, sh_cur, sh_ini;
  __shared__ int sh_flag[4];  // [0] again (retry with a larger lambda), [1] ok (next iteration), [2] ok2 (solve succeeded), [3] nBad
  const int f = blockIdx.x, tid = threadIdx.x;
  const long long o0 = A.frame_ptr[f], o1 = A.frame_ptr[f + 1];
  const int n = (int)(o1 - o0);
  double cam[5];
#pragma unroll
  for (int i = 0; i < 5; i++) cam[i] = A.cam[f * 5 + i];
  if (tid == 0) {
    double q[4] = {A.pose[f * 7 + 3], A.pose[f * 7 + 4], A.pose[f * 7 + 5], A.pose[f * 7 + 6]};
    quat_normalize_pos_w(q);
    for (int i = 0; i < 3; i++) sh_pose0[i] = A.pose[f * 7 + i];
    for (int i = 0; i < 4; i++) sh_pose0[3 + i] = q[i];
    A.trace_len[f] = 0;
  }
  for (long long o = o0 + tid; o < o1; o += PO_CTA) { A.level[o] = 0; A.outlier[o] = 0; }
  __syncthreads();
  if (n < 3) {  // g2oOptimizer.cc:491-492
    if (tid == 0) A.inliers[f] = 0;
    return;
  }
  const double dMono = (double)(float)sqrt(5.991), dStereo = (double)(float)sqrt(7.815);  // const float delta*, :426-428
  int trace_len = 0;  // thread 0 only
  for (int round = 0; round < 4; round++) {
    const bool robust = round < 3;  // kernels are removed while classifying round 2, :544-545
    if (tid < 7) sh_pose[tid] = sh_pose0[tid];  // vSE3->setEstimate(mTcw) before every round, :510
    if (tid == 0) sh_flag[1] = 1;
    __syncthreads();
    for (int iter = 0; iter < 10; iter++) {
      if (!sh_flag[1]) break;  // CTA-uniform (written before the last barrier of the previous iteration)
      // ---- computeActiveErrors + activeRobustChi2 + buildSystem at the current estimate
      double R[9];
      quat_to_R(sh_pose + 3, R);
      double acc[29];  // H (21 upper) | b (6) | chi2 | number of active edges
#pragma unroll
      for (int k = 0; k < 29; k++) acc[k] = 0.0;
      for (long long o = o0 + tid; o < o1; o += PO_CTA) {
        if (A.level[o]) continue;
        PoEdge E;
        po_eval(R, sh_pose, cam, A.xyz + o * 3, A.meas[o], true, E);
#pragma unroll
        for (int c = 0; c < 3; c++) A.err[o * 3 + c] = E.e[c];
        const double c2 = po_chi2(E.e, E.info, E.stereo);
        double rho0 = c2, rho1 = 1.0;
        if (robust) { const double d = E.stereo ? dStereo : dMono; huber(c2, d, huber_dsqr(d), &rho0, &rho1); }
        acc[27] += rho0;
        acc[28] += 1.0;
        const double w = rho1 * E.info;
        int idx = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
          acc[21 + i] -= rho1 * (E.J[i] * (E.info * E.e[0]) + E.J[6 + i] * (E.info * E.e[1]) + E.J[12 + i] * (E.info * E.e[2]));
#pragma unroll
          for (int j = i; j < 6; j++) {
            acc[idx] += E.J[i] * (w * E.J[j]) + E.J[6 + i] * (w * E.J[6 + j]) + E.J[12 + i] * (w * E.J[12 + j]);
            idx++;
          }
        }
      }
      po_reduce<29>(acc, sh_part, sh_red);
      if (sh_red[28] == 0.0) break;  // no active edge left: optimize() returns without touching the vertex (CTA-uniform)
      if (tid < 21) sh_H[tid] = sh_red[tid];
      if (tid < 6) sh_b[tid] = sh_red[21 + tid];
      if (tid == 0) {
        sh_cur = sh_red[27];
        sh_ini = sh_red[27];
        if (iter == 0) {  // computeLambdaInit, optimization_algorithm_levenberg.cpp:166-180
          double md = 0.0;
          int d = 0;
          for (int i = 0; i < 6; i++) { md = fmax(md, fabs(sh_red[d])); d += 6 - i; }
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
          for (int i = 0; i < 7; i++) sh_bak[i] = sh_pose[i];
          double x[6] = {0, 0, 0, 0, 0, 0};
          const bool ok2 = po_solve6(sh_H, sh_lambda, sh_b, x);
          if (!ok2) for (int i = 0; i < 6; i++) x[i] = 0.0;
          for (int i = 0; i < 6; i++) sh_x[i] = x[i];
          double p[7];
          for (int i = 0; i < 7; i++) p[i] = sh_pose[i];
          pose_oplus(p, x);
          for (int i = 0; i < 7; i++) sh_pose[i] = p[i];
          sh_flag[2] = ok2 ? 1 : 0;
        }
        __syncthreads();
        quat_to_R(sh_pose + 3, R);
        double chi[1] = {0.0};
        for (long long o = o0 + tid; o < o1; o += PO_CTA) {
          if (A.level[o]) continue;
          PoEdge E;
          po_eval(R, sh_pose, cam, A.xyz + o * 3, A.meas[o], false, E);
#pragma unroll
          for (int c = 0; c < 3; c++) A.err[o * 3 + c] = E.e[c];  // a rejected trial leaves these behind (stale _error)
          const double c2 = po_chi2(E.e, E.info, E.stereo);
          double rho0 = c2, rho1 = 1.0;
          if (robust) { const double d = E.stereo ? dStereo : dMono; huber(c2, d, huber_dsqr(d), &rho0, &rho1); }
          chi[0] += rho0;
        }
        po_reduce<1>(chi, sh_part, sh_red);
        if (tid == 0) {
          double tempChi = sh_red[0];
          if (!sh_flag[2]) tempChi = 1.7976931348623157e308;
          rho = sh_cur - tempChi;
          double scale = 0.0;
          for (int j = 0; j < 6; j++) scale += sh_x[j] * (sh_lambda * sh_x[j] + sh_b[j]);
          scale += 1e-3;
          rho /= scale;
          const bool good = (rho > 0.0) && isfinite(tempChi);
          if (trace_len < PO_MAX_TRACE) {
            double* tr = A.trace + ((size_t)f * PO_MAX_TRACE + trace_len) * PO_TRACE_COLS;
            tr[0] = round; tr[1] = iter; tr[2] = qmax; tr[3] = sh_lambda; tr[4] = sh_cur; tr[5] = tempChi; tr[6] = rho;
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
            for (int i = 0; i < 7; i++) sh_pose[i] = sh_bak[i];
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
    // ---- chi2 re-classification (g2oOptimizer.cc:518-547): float chi2 against float thresholds; an observation
    //      flagged as outlier gets its error recomputed at the new estimate, an inlier keeps the stored one
    {
      double R[9];
      quat_to_R(sh_pose + 3, R);
      for (long long o = o0 + tid; o < o1; o += PO_CTA) {
        const float4 m = A.meas[o];
        const bool stereo = !(m.z < 0.0f);
        double e[3] = {A.err[o * 3], A.err[o * 3 + 1], A.err[o * 3 + 2]};
        if (A.outlier[o]) {
          PoEdge E;
          po_eval(R, sh_pose, cam, A.xyz + o * 3, m, false, E);
#pragma unroll
          for (int c = 0; c < 3; c++) { e[c] = E.e[c]; A.err[o * 3 + c] = E.e[c]; }
        }
        const float chi2 = (float)po_chi2(e, (double)m.w, stereo);
        const bool bad = chi2 > (stereo ? 7.815f : 5.991f);
        A.outlier[o] = bad ? 1 : 0;
        A.level[o] = bad ? 1 : 0;
      }
    }
    __syncthreads();
    if (n < 10) break;  // optimizer.edges().size() < 10, :549-550
  }
  if (A.skip_final) {  // the lidar block follows (csrc/sqrtba_poseopt_lidar.cuh): it classifies after its own round
    if (tid == 0) A.trace_len[f] = trace_len;
    if (tid < 7) A.pose[f * 7 + tid] = sh_pose[tid];
    return;
  }
  // ---- final classification (:656-680): double thresholds, same stale/recomputed error rule
  double nb[1] = {0.0};
  {
    double R[9];
    quat_to_R(sh_pose + 3, R);
    for (long long o = o0 + tid; o < o1; o += PO_CTA) {
      const float4 m = A.meas[o];
      const bool stereo = !(m.z < 0.0f);
      double e[3] = {A.err[o * 3], A.err[o * 3 + 1], A.err[o * 3 + 2]};
      if (A.outlier[o]) {
        PoEdge E;
        po_eval(R, sh_pose, cam, A.xyz + o * 3, m, false, E);
#pragma unroll
        for (int c = 0; c < 3; c++) e[c] = E.e[c];
      }
      const float chi2 = (float)po_chi2(e, (double)m.w, stereo);
      const bool bad = (double)chi2 > (stereo ? 7.815 : 5.991);
      A.outlier[o] = bad ? 1 : 0;
      nb[0] += bad ? 1.0 : 0.0;
    }
  }
  po_reduce<1>(nb, sh_part, sh_red);
  if (tid == 0) {
    A.inliers[f] = n - (int)(sh_red[0] + 0.5);
    A.trace_len[f] = trace_len;
  }
  if (tid < 7) A.pose[f * 7 + tid] = sh_pose[tid];
}

}  // namespace sqrtba

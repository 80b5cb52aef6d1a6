// The lidar block this fork adds to g2oOptimizer::PoseOptimization (src/backend/g2oOptimizer.cc:560-640; SURVEY.md §8(f)
// N1 cites :385-679): when the tracker's local lidar map has more than 100 points, the frame's flat / sharp feature
// points are moved to the world frame with the pose of the four visual rounds, matched to their nearest map point
// (pcl::KdTreeFLANN, k = 1, squared distance below lidarConfig::distance_sq_threshold), every match becomes a unary
// EdgeLidarFlatPoint / EdgeLidarCornerPoint on the pose (types_six_dof_expmap.h:206-262: numeric Jacobians,
// information = weight, no robust kernel), the estimate is reset to the float-rounded pose (pFrame->SetPose followed by
// setEstimate(toSE3Quat(mTcw)), :557, :632) and a fifth optimize(10) runs over the level-0 visual edges and the lidar
// edges; the final classification (:656-680) follows.
//
// Three launches after k_pose_opt(skip_final): the association kernels of the local-BA lidar pass (csrc/sqrtba_lidar.cuh:
// k_lidar_to_world for the frame's own points, k_lidar_nn per point kind against the SAME map cloud, k_lidar_edges) and
// k_pose_opt_lidar below -- one CTA: the pose-only LM of csrc/sqrtba_poseopt.cuh with the unary edges added to the
// 6x6 system and to the cost of every trial.
#pragma once
#include "sqrtba_lidar.cuh"
#include "sqrtba_poseopt.cuh"

namespace sqrtba {

__global__ void __launch_bounds__(PO_CTA) k_pose_opt_lidar(PoseOptArgs A, LidarDev L, int* n_match2) {
  __shared__ double sh_part[PO_WARPS * 29];
  __shared__ double sh_red[29];
  __shared__ double sh_pose[7], sh_bak[7], sh_x[6], sh_H[21], sh_b[6];
  __shared__ double sh_lambda, sh_ni, sh_cur, sh_ini;
  __shared__ int sh_flag[4];
  const int tid = threadIdx.x;
  const long long o0 = A.frame_ptr[0], o1 = A.frame_ptr[1];
  const int n = (int)(o1 - o0);
  double cam[5];
#pragma unroll
  for (int i = 0; i < 5; i++) cam[i] = A.cam[i];
  if (tid == 0) {
    // Converter::toCvMat(estimate) -> CV_32F -> Converter::toSE3Quat: rotation matrix and translation rounded to float,
    // Eigen's Quaterniond(Matrix3d), normalised with w >= 0 by the SE3Quat constructor
    double R[9], q[4];
    quat_to_R(A.pose + 3, R);
    for (int i = 0; i < 9; i++) R[i] = (double)(float)R[i];
    R_to_quat(R, q);
    quat_normalize_pos_w(q);
    for (int i = 0; i < 3; i++) sh_pose[i] = (double)(float)A.pose[i];
    for (int i = 0; i < 4; i++) sh_pose[3 + i] = q[i];
    sh_flag[1] = 1;
  }
  // matched flat / corner points (the two counts the reference prints, :626-627)
  {
    double cnt[2] = {0.0, 0.0};
    for (int e = tid; e < L.n_edge; e += PO_CTA)
      if (L.w[e] > 0.0) cnt[e >= L.n_flat ? 1 : 0] += 1.0;
    po_reduce<2>(cnt, sh_part, sh_red);
    if (tid < 2) n_match2[tid] = (int)(sh_red[tid] + 0.5);
    __syncthreads();
  }
  const double dMono = (double)(float)sqrt(5.991), dStereo = (double)(float)sqrt(7.815);
  const bool robust = n < 10;  // the kernels are dropped while classifying the third round; with fewer than 10 edges the
                               // schedule stops after the first (:549-550) and they are still attached
  int trace_len = A.trace_len[0];  // thread 0 only
  for (int iter = 0; iter < 10; iter++) {
    if (!sh_flag[1]) break;
    double R[9];
    quat_to_R(sh_pose + 3, R);
    double acc[29];
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
    {
      double pose[7];
#pragma unroll
      for (int i = 0; i < 7; i++) pose[i] = sh_pose[i];
      for (int e = tid; e < L.n_edge; e += PO_CTA) {
        const double w = L.w[e];
        if (!(w > 0.0)) continue;
        const bool corner = e >= L.n_flat;
        const double* pc = L.pc + (size_t)e * 3;
        const double* qw = L.qw + (size_t)e * 3;
        const double* nv = L.nv + (size_t)e * 3;
        const double err = lidar_error(pose, pc, qw, nv, corner);
        double J[6];
        lidar_jacobian(L, pose, pc, qw, nv, corner, J);
        int idx = 0;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
          for (int j = i; j < 6; j++) acc[idx++] += J[i] * w * J[j];
#pragma unroll
        for (int i = 0; i < 6; i++) acc[21 + i] -= J[i] * w * err;
        acc[27] += err * (w * err);
        acc[28] += 1.0;
      }
    }
    po_reduce<29>(acc, sh_part, sh_red);
    if (sh_red[28] == 0.0) break;  // no active edge at all
    if (tid < 21) sh_H[tid] = sh_red[tid];
    if (tid < 6) sh_b[tid] = sh_red[21 + tid];
    if (tid == 0) {
      sh_cur = sh_red[27];
      sh_ini = sh_red[27];
      if (iter == 0) {
        double md = 0.0;
        int d = 0;
        for (int i = 0; i < 6; i++) { md = fmax(md, fabs(sh_red[d])); d += 6 - i; }
        sh_lambda = 1e-5 * md;
        sh_ni = 2.0;
        sh_flag[3] = 0;
      }
    }
    __syncthreads();
    double rho = 0.0;
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
        for (int c = 0; c < 3; c++) A.err[o * 3 + c] = E.e[c];
        const double c2 = po_chi2(E.e, E.info, E.stereo);
        double rho0 = c2, rho1 = 1.0;
        if (robust) { const double d = E.stereo ? dStereo : dMono; huber(c2, d, huber_dsqr(d), &rho0, &rho1); }
        chi[0] += rho0;
      }
      {
        double pose[7];
#pragma unroll
        for (int i = 0; i < 7; i++) pose[i] = sh_pose[i];
        for (int e = tid; e < L.n_edge; e += PO_CTA) {
          const double w = L.w[e];
          if (!(w > 0.0)) continue;
          const double err = lidar_error(pose, L.pc + (size_t)e * 3, L.qw + (size_t)e * 3, L.nv + (size_t)e * 3, e >= L.n_flat);
          chi[0] += err * (w * err);
        }
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
          double* tr = A.trace + (size_t)trace_len * PO_TRACE_COLS;
          tr[0] = 4; tr[1] = iter; tr[2] = qmax; tr[3] = sh_lambda; tr[4] = sh_cur; tr[5] = tempChi; tr[6] = rho;
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
        if (!again) {
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
  // ---- final classification (:656-680)
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
    A.inliers[0] = n - (int)(sh_red[0] + 0.5);
    A.trace_len[0] = trace_len;
  }
  if (tid < 7) A.pose[tid] = sh_pose[tid];
}

}  // namespace sqrtba

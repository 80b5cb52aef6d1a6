// Lidar tight-coupling pass: the CUDA counterpart of the block this fork adds to
// g2oOptimizer::LocalBundleAdjustment between the second visual pass and the final outlier test
// (src/backend/g2oOptimizer.cc:979-1117).
//
//   1. local lidar map: the flat / corner feature clouds of every OTHER local keyframe are moved to the world frame with
//      the pose estimates of pass 2 (:985-1013)                                             -> k_lidar_to_world
//   2. association: each feature point of the current keyframe, moved to the world frame, is matched to its nearest
//      map point (pcl::KdTreeFLANN, k = 1) and kept when the squared distance is below distance_sq_threshold
//      (:1043-1049, :1081-1087)                                                              -> k_lidar_nn, k_lidar_edges
//   3. every match becomes a unary 1-D edge on the current keyframe's pose: point-to-plane (EdgeLidarFlatPoint) or
//      point-to-point distance (EdgeLidarCornerPoint), types_six_dof_expmap.h:206-262, information = weight, no robust
//      kernel; then optimize(20) over visual + lidar edges (:1113-1114)
//
// A unary edge touches no landmark, so in the square-root formulation it never enters the landmark QR: its 6x6
// information block H_u = sum w J^T J and gradient b_u = -sum w J^T e are ADDED to the reduced camera system of the
// current keyframe's slot -- to the gradient / Hessian diagonal after the linearisation (b, lambda_0), to the reduced
// right-hand side and the block-Jacobi block after the landmark QR, to q = A p after every matvec, and its chi2 to the
// cost of every trial.  Each of these is one tiny kernel launched between the existing ones, which stay untouched.
#pragma once
#include "sqrtba_kernels.cuh"

namespace sqrtba {

constexpr int LD_CTA = 256;
constexpr int LD_WARPS = LD_CTA / 32;
constexpr int LD_ACC = 32;  // [0,21) H_u upper triangle row-major, [21,27) b_u, 27 chi2 at the linearisation point, 28 chi2 of the trial

struct LidarDev {
  int n_edge, n_flat;  // edges [0, n_flat) are flat (point-to-plane), [n_flat, n_edge) corner (point-to-point)
  int pose;            // pose index of the current keyframe
  int numeric;         // 1: central-difference Jacobians, delta 1e-9, as BaseUnaryEdge::linearizeOplus (the reference)
  const double* pc;    // n_edge x 3  curpoint_cameraframe_
  const double* qw;    // n_edge x 3  lastpoint_worldframe_
  const double* nv;    // n_edge x 3  curr_point_norm (flat edges)
  const double* w;     // n_edge      information; 0 = no correspondence, the edge does not exist
  double* acc;         // LD_ACC
};

// computeError of both edge types (types_six_dof_expmap.h:216-226, 245-253): with M = Tcw^-1, Rwc = M.block(0,0,3,3),
// Ow = M.col(3).head(3):  d = Rwc^-1 (lastpoint_world - Ow) - curpoint_camera;  flat: d . n, corner: |d|.
// The inverses of the rigid transform are taken in closed form (Rwc = R^T, Ow = -R^T t, Rwc^-1 = R).
__device__ __forceinline__ double lidar_error(const double pose[7], const double* pc, const double* qw, const double* n,
                                              bool corner) {
  double R[9];
  quat_to_R(pose + 3, R);
  double Ow[3], v[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; i++) Ow[i] = -(R[0 * 3 + i] * pose[0] + R[1 * 3 + i] * pose[1] + R[2 * 3 + i] * pose[2]);
#pragma unroll
  for (int i = 0; i < 3; i++) v[i] = qw[i] - Ow[i];
#pragma unroll
  for (int i = 0; i < 3; i++) d[i] = (R[i * 3 + 0] * v[0] + R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2]) - pc[i];
  if (corner) return sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  return d[0] * n[0] + d[1] * n[1] + d[2] * n[2];
}

// 1x6 Jacobian w.r.t. the left-multiplicative update (omega, upsilon) of VertexSE3Expmap
__device__ __forceinline__ void lidar_jacobian(const LidarDev& L, const double pose[7], const double* pc, const double* qw,
                                               const double* n, bool corner, double J[6]) {
  if (L.numeric) {  // base_unary_edge.hpp:82-123: per direction push / oplus(+-delta) / computeError / pop
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    for (int d = 0; d < 6; d++) {
      double xi[6] = {0, 0, 0, 0, 0, 0}, pp[7];
      xi[d] = delta;
      for (int i = 0; i < 7; i++) pp[i] = pose[i];
      pose_oplus(pp, xi);
      const double e1 = lidar_error(pp, pc, qw, n, corner);
      xi[d] = -delta;
      for (int i = 0; i < 7; i++) pp[i] = pose[i];
      pose_oplus(pp, xi);
      const double e2 = lidar_error(pp, pc, qw, n, corner);
      J[d] = scalar * (e1 - e2);
    }
    return;
  }
  // closed form: Xc = R q + t, err = g . (Xc - pc) with g = n (flat) or (Xc - pc)/|Xc - pc| (corner);
  // dXc = omega x Xc + upsilon  ->  J = [Xc x g, g]
  double R[9], Xc[3], g[3];
  quat_to_R(pose + 3, R);
  transform(R, pose, qw, Xc);
  if (corner) {
    const double d[3] = {Xc[0] - pc[0], Xc[1] - pc[1], Xc[2] - pc[2]};
    const double nn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
    for (int i = 0; i < 3; i++) g[i] = nn > 0.0 ? d[i] / nn : 0.0;
  } else {
#pragma unroll
    for (int i = 0; i < 3; i++) g[i] = n[i];
  }
  J[0] = Xc[1] * g[2] - Xc[2] * g[1];
  J[1] = Xc[2] * g[0] - Xc[0] * g[2];
  J[2] = Xc[0] * g[1] - Xc[1] * g[0];
  J[3] = g[0]; J[4] = g[1]; J[5] = g[2];
}

// CTA sum of NV per-thread values; the totals end up in out[0..NV) of thread 0 only
template <int NV>
__device__ __forceinline__ void lidar_reduce(double* v, double (*sh)[LD_WARPS]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    const double s = warp_sum(v[k]);
    if (lane == 0) sh[k][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double s = 0.0;
      for (int i = 0; i < LD_WARPS; i++) s += sh[k][i];
      v[k] = s;
    }
  }
}

// after the visual linearisation of an LM iteration: errors + Jacobians of the lidar edges at the current estimate,
// H_u / b_u / chi2 into acc, and their share of the gradient, the Hessian diagonal (lambda_0 = tau * max diag,
// optimization_algorithm_levenberg.cpp:166-180) and the cost of the linearisation point.  One CTA, single window.
__global__ void __launch_bounds__(LD_CTA) k_lidar_lin(Dev P, LidarDev L) {
  __shared__ double sh[28][LD_WARPS];
  if (P.ctl[0].phase != PH_LIN) return;
  const int slot = P.pose_slot[L.pose];
  if (slot < 0) return;  // fixed vertex: the edges are not active (sparse_optimizer.cpp:218-235)
  double pose[7];
  for (int i = 0; i < 7; i++) pose[i] = P.pose[L.pose * 7 + i];
  double a[28];
#pragma unroll
  for (int k = 0; k < 28; k++) a[k] = 0.0;
  for (int e = threadIdx.x; e < L.n_edge; e += LD_CTA) {
    const double w = L.w[e];
    if (!(w > 0.0)) continue;
    const bool corner = e >= L.n_flat;
    const double* pc = L.pc + (size_t)e * 3;
    const double* qw = L.qw + (size_t)e * 3;
    const double* n = L.nv + (size_t)e * 3;
    const double err = lidar_error(pose, pc, qw, n, corner);
    double J[6];
    lidar_jacobian(L, pose, pc, qw, n, corner, J);
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = i; j < 6; j++) a[idx++] += J[i] * w * J[j];  // base_unary_edge.hpp:60-61
#pragma unroll
    for (int i = 0; i < 6; i++) a[21 + i] -= J[i] * w * err;
    a[27] += err * (w * err);
  }
  lidar_reduce<28>(a, sh);
  if (threadIdx.x == 0) {
    for (int k = 0; k < 28; k++) L.acc[k] = a[k];
    int idx = 0;
    for (int i = 0; i < 6; i++) {
      P.bp[slot * 6 + i] += a[21 + i];
      P.hd[slot * 6 + i] += a[idx];
      idx += 6 - i;
    }
    P.chi_part[P.win_item_ptr[0]] += a[27];
  }
}

// after the landmark QR of a trial (and before the block-Jacobi inverse): reduced rhs += b_u, block += H_u
__global__ void k_lidar_trial_add(Dev P, LidarDev L) {
  if (P.ctl[0].phase != PH_TRIAL) return;
  const int slot = P.pose_slot[L.pose];
  if (slot < 0) return;
  const int k = threadIdx.x;
  if (k < 21) P.D[(size_t)slot * 21 + k] += L.acc[k];
  else if (k < 27) P.bs[(size_t)slot * 6 + (k - 21)] += L.acc[k];
}

// after every matvec of the PCG (before the dot products): q[slot] += H_u p[slot]
__global__ void k_lidar_matvec(Dev P, LidarDev L, const double* __restrict__ pvec, double* __restrict__ qvec) {
  if (!P.ctl[0].cg_active) return;
  const int slot = P.pose_slot[L.pose];
  if (slot < 0) return;
  const int r = threadIdx.x;
  if (r >= 6) return;
  double s = 0.0;
  for (int c = 0; c < 6; c++) {
    const int i = r < c ? r : c, j = r < c ? c : r;
    const int idx = i * 6 - (i * (i - 1)) / 2 + (j - i);  // upper triangle, row-major
    s += L.acc[idx] * pvec[(size_t)slot * 6 + c];
  }
  qvec[(size_t)slot * 6 + r] += s;
}

// after the visual cost evaluation of a trial: chi2 of the lidar edges at the trial estimate
__global__ void __launch_bounds__(LD_CTA) k_lidar_cost(Dev P, LidarDev L) {
  __shared__ double sh[1][LD_WARPS];
  if (P.ctl[0].phase != PH_TRIAL) return;
  if (P.pose_slot[L.pose] < 0) return;
  double pose[7];
  for (int i = 0; i < 7; i++) pose[i] = P.pose[L.pose * 7 + i];
  double a[1] = {0.0};
  for (int e = threadIdx.x; e < L.n_edge; e += LD_CTA) {
    const double w = L.w[e];
    if (!(w > 0.0)) continue;
    const double err = lidar_error(pose, L.pc + (size_t)e * 3, L.qw + (size_t)e * 3, L.nv + (size_t)e * 3, e >= L.n_flat);
    a[0] += err * (w * err);
  }
  lidar_reduce<1>(a, sh);
  if (threadIdx.x == 0) {
    L.acc[28] = a[0];
    P.chi_part[P.win_item_ptr[0]] += a[0];
  }
}

// ---------------------------------------------------------------------------------------------------- association
struct LidarAssoc {
  int pose;                      // current keyframe
  int n_flat, n_corner;          // its feature points (own frame)
  const float* flat;  const float* flat_n;  const float* corner;
  long long n_map_flat, n_map_corner;  // feature points of the other local keyframes, each in its keyframe's frame
  const float* map_flat;   const int* map_flat_pose;
  const float* map_corner; const int* map_corner_pose;
  float* map_flat_w;  float* map_corner_w;  // the same points in the world frame (scratch)
  float* cur_w;                  // (n_flat + n_corner) x 3: current points in the world frame (scratch)
  unsigned long long* best;      // n_flat + n_corner: (squared distance bits << 32) | map index, reduced by atomicMin
  int* match;                    // n_flat + n_corner: matched map index or -1
  double thr, w_flat, w_corner;
  int use_flat, use_corner;
};

// Twc the way the reference obtains it: Converter::toCvMat(SE3Quat) (double -> CV_32F 4x4) followed by cv::Mat::inv().
// The inverse of the float-rounded rigid transform is taken in closed form, evaluated in double, rounded to float.
__device__ __forceinline__ void twc_float(const double* pose, float Rwc[9], float Ow[3]) {
  double R[9];
  quat_to_R(pose + 3, R);
  float Rf[9], tf[3];
#pragma unroll
  for (int i = 0; i < 9; i++) Rf[i] = (float)R[i];
#pragma unroll
  for (int i = 0; i < 3; i++) tf[i] = (float)pose[i];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Rwc[i * 3 + j] = Rf[j * 3 + i];
#pragma unroll
  for (int i = 0; i < 3; i++)
    Ow[i] = (float)(-((double)Rf[0 * 3 + i] * tf[0] + (double)Rf[1 * 3 + i] * tf[1] + (double)Rf[2 * 3 + i] * tf[2]));
}
// pcl::transformPointCloud(in, out, Eigen::Affine3d): evaluated in double, cast to the float fields
__device__ __forceinline__ void to_world_float(const float Rwc[9], const float Ow[3], const float* p, float* out) {
#pragma unroll
  for (int i = 0; i < 3; i++)
    out[i] = (float)((double)Rwc[i * 3 + 0] * p[0] + (double)Rwc[i * 3 + 1] * p[1] + (double)Rwc[i * 3 + 2] * p[2] + (double)Ow[i]);
}

// which = 0: flat map, 1: corner map, 2: the current keyframe's own points (flat then corner); also resets `best`
__global__ void k_lidar_to_world(Dev P, LidarAssoc A, int which) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = which == 0 ? A.n_map_flat : (which == 1 ? A.n_map_corner : (long long)A.n_flat + A.n_corner);
  if (i >= n) return;
  const float* src;
  float* dst;
  int pose;
  if (which == 0) { src = A.map_flat + i * 3; dst = A.map_flat_w + i * 3; pose = A.map_flat_pose[i]; }
  else if (which == 1) { src = A.map_corner + i * 3; dst = A.map_corner_w + i * 3; pose = A.map_corner_pose[i]; }
  else {
    src = i < A.n_flat ? A.flat + i * 3 : A.corner + (i - A.n_flat) * 3;
    dst = A.cur_w + i * 3;
    pose = A.pose;
    A.best[i] = ~0ull;
  }
  float Rwc[9], Ow[3];
  twc_float(P.pose + (size_t)pose * 7, Rwc, Ow);
  to_world_float(Rwc, Ow, src, dst);
}

// exact nearest neighbour by brute force: grid = (query blocks, map chunks); every CTA stages its map chunk in shared
// memory, every thread owns one query and keeps its best candidate, one atomicMin on the packed (distance, index) word
// per (query, chunk).  Distances as flann::L2_Simple<float>: float accumulation of squared differences, no FMA.
// A non-negative float orders like its bit pattern, so the packed minimum is the nearest point, ties to the smallest index.
constexpr int LD_NN_CTA = 128;
constexpr int LD_NN_CHUNK = 2048;
__global__ void __launch_bounds__(LD_NN_CTA) k_lidar_nn(LidarAssoc A, int corner) {
  __shared__ float ms[LD_NN_CHUNK * 3];
  const int nq = corner ? A.n_corner : A.n_flat;
  const int qoff = corner ? A.n_flat : 0;
  const long long nm = corner ? A.n_map_corner : A.n_map_flat;
  const float* mapw = corner ? A.map_corner_w : A.map_flat_w;
  const long long m0 = (long long)blockIdx.y * LD_NN_CHUNK;
  const int mc = (int)min((long long)LD_NN_CHUNK, nm - m0);
  for (int i = threadIdx.x; i < mc * 3; i += LD_NN_CTA) ms[i] = mapw[m0 * 3 + i];
  __syncthreads();
  const int q = blockIdx.x * LD_NN_CTA + threadIdx.x;
  if (q >= nq) return;
  const float qx = A.cur_w[(size_t)(qoff + q) * 3], qy = A.cur_w[(size_t)(qoff + q) * 3 + 1], qz = A.cur_w[(size_t)(qoff + q) * 3 + 2];
  float bd = __int_as_float(0x7f800000);
  int bi = -1;
  for (int j = 0; j < mc; j++) {
    const float dx = __fsub_rn(qx, ms[j * 3]), dy = __fsub_rn(qy, ms[j * 3 + 1]), dz = __fsub_rn(qz, ms[j * 3 + 2]);
    const float r = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    if (r < bd) { bd = r; bi = j; }
  }
  if (bi >= 0) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(bd) << 32) | (unsigned long long)(unsigned)(m0 + bi);
    atomicMin(A.best + qoff + q, key);
  }
}

// matches -> edges (g2oOptimizer.cc:1049-1071, 1087-1105); an unmatched point keeps weight 0
__global__ void k_lidar_edges(LidarAssoc A, LidarDev L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n_flat + A.n_corner) return;
  const bool corner = i >= A.n_flat;
  const bool used = corner ? (A.use_corner && A.n_map_corner > 0) : (A.use_flat && A.n_map_flat > 0);
  double w = 0.0;
  int match = -1;
  double* pc = const_cast<double*>(L.pc) + (size_t)i * 3;
  double* qw = const_cast<double*>(L.qw) + (size_t)i * 3;
  double* nv = const_cast<double*>(L.nv) + (size_t)i * 3;
  for (int k = 0; k < 3; k++) { pc[k] = 0.0; qw[k] = 0.0; nv[k] = 0.0; }
  if (used) {
    const unsigned long long key = A.best[i];
    const float d2 = __uint_as_float((unsigned)(key >> 32));
    if (key != ~0ull && (double)d2 < A.thr) {
      match = (int)(unsigned)(key & 0xffffffffull);
      const float* src = corner ? A.corner + (size_t)(i - A.n_flat) * 3 : A.flat + (size_t)i * 3;
      const float* mp = (corner ? A.map_corner_w : A.map_flat_w) + (size_t)match * 3;
      for (int k = 0; k < 3; k++) {
        pc[k] = (double)src[k];
        qw[k] = (double)mp[k];
        nv[k] = corner ? 0.0 : (double)A.flat_n[(size_t)i * 3 + k];
      }
      w = corner ? A.w_corner : A.w_flat;
    }
  }
  const_cast<double*>(L.w)[i] = w;
  A.match[i] = match;
}

}  // namespace sqrtba

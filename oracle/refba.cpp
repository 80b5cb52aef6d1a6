// refba -- CPU ORACLE for the bundle-adjustment hot path of lutao98/SqrtLM-SLAM.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (sqrtlm-slam_b200/, include/) may link,
// import or execute this file; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
// `--impl reference` legs use it, and only as the checker / the timed CPU baseline.
//
// It is a dependency-free FP64 restatement of the reference's default BA back-end (vendored g2o
// driven by src/backend/g2oOptimizer.cc).  The reference itself cannot be compiled here (needs
// Eigen, OpenCV, PCL, Ceres, ROS -- none installed, no network), so this is a "port" oracle.
// PARITY PARTLY PINNED: the reference has no tests, golden vectors or fixtures.  Unpinned upstream: the
// Schur complement + sparse LDLT numerics (BlockSolver and LinearSolverEigen are header-only and absent from the
// shipped binary; exact-in-exact-arithmetic operations), the adapter control flow of g2oOptimizer.cc (two passes,
// outlier policy: needs OpenCV/PCL types) and the lidar pass.  Everything else IS pinned against
// the reference's own compiled code: its tree ships a prebuilt Thirdparty/g2o/lib/libg2o.so that exports
// SE3Quat::exp, project2d, the mono / stereo cam_project and RobustKernelHuber::robustify; oracle/pin_libg2o.py calls
// them through ctypes and tests/test_pin_libg2o.py checks this file against the recorded outputs
// (tests/golden/libg2o_vectors.npz); oracle/pin_libg2o_edges.py does the same with REAL vertex and edge objects of that
// binary (oplusImpl, computeError, linearizeOplus): residuals and Jacobians agree to 1e-15.  The check found one quirk
// the sources hide in a header: `float dsqr`.  oracle/pin_libg2o_graph.py pins the GRAPH semantics against a real
// g2o::SparseOptimizer of that binary: index mapping of initializeOptimization (free poses, then landmarks, ascending
// id; fixed and edge-less vertices excluded), computeActiveErrors leaving level-1 edges' _error stale, activeChi2 /
// activeRobustChi2, update() in index order, push/pop, and the normal equations that the binary's linearizeOplus +
// constructQuadraticForm accumulate into mapped Hessian blocks (refba_debug_phase / refba_debug_system below
// reproduce all of it: indices exactly, numbers to 1e-12).  The same script runs the binary's own
// SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve on a g2o::Solver it supplies (vtable of callbacks;
// the binary's constructQuadraticForm + a dense solve): lmSolve / optimize below reproduce its lambda sequence trial by
// trial, including rejected trials, both clamps of the lambda factor and the _nBad stop rule; and the two-pass
// local-BA schedule driven over the binary's objects gives the lambda sequences, level-1 set, outlier flags and
// estimates that refba_solve_local gives; likewise refba_pose_opt against the four-round pose-only schedule over the
// binary's EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose objects, and refba_sim3_* / refba_pose_graph
// (essential-graph optimisation) against the binary's Sim3, VertexSim3Expmap and EdgeSim3.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Eigen is not vendored in the reference; where g2o calls into Eigen (quaternion*vector,
// Quaterniond(Matrix3d), toRotationMatrix, Matrix3d::inverse, SimplicialLDLT) the published Eigen 3
// algorithm is restated -- exact-in-exact-arithmetic operations, so only rounding can differ.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// SE3Quat  (Thirdparty/g2o/g2o/types/se3quat.h)
// ---------------------------------------------------------------------------------------------
struct Quat { double x, y, z, w; };
struct SE3 { Quat r; double t[3]; };

// Eigen::Quaternion::operator*(Quaternion)  (used by se3quat.h:104-110, `result._r*=tr2._r`)
inline Quat qmul(const Quat& a, const Quat& b) {
  Quat c;
  c.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  c.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  c.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  c.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return c;
}

// Eigen::Quaternion::_transformVector: v + w*uv + vec x uv with uv = 2 (vec x v)   (se3quat.h:217-220 `_r*xyz`)
inline void qrot(const Quat& q, const double v[3], double out[3]) {
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
  out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
  out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}

// Eigen::Quaternion::toRotationMatrix
inline void qtoR(const Quat& q, double R[9]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen::Quaterniond(Matrix3d)  (se3quat.h:58, 256)
inline Quat RtoQ(const double m[9]) {
  double q[4];  // x y z w
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 3 + i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
    q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
  }
  return Quat{q[0], q[1], q[2], q[3]};
}

// SE3Quat::normalizeRotation  (se3quat.h:280-285)
inline void normalizeRotation(SE3& T) {
  if (T.r.w < 0) { T.r.x *= -1; T.r.y *= -1; T.r.z *= -1; T.r.w *= -1; }
  const double n = std::sqrt(T.r.x * T.r.x + T.r.y * T.r.y + T.r.z * T.r.z + T.r.w * T.r.w);
  T.r.x /= n; T.r.y /= n; T.r.z /= n; T.r.w /= n;
}

// SE3Quat::operator*  (se3quat.h:104-110)
inline SE3 se3mul(const SE3& a, const SE3& b) {
  SE3 r = a;
  double rt[3];
  qrot(a.r, b.t, rt);
  r.t[0] += rt[0]; r.t[1] += rt[1]; r.t[2] += rt[2];
  r.r = qmul(a.r, b.r);
  normalizeRotation(r);
  return r;
}

// SE3Quat::map  (se3quat.h:217-220)
inline void se3map(const SE3& T, const double p[3], double out[3]) {
  qrot(T.r, p, out);
  out[0] += T.t[0]; out[1] += T.t[1]; out[2] += T.t[2];
}

inline void mat3mul(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}

// SE3Quat::exp  (se3quat.h:223-257); update = (omega, upsilon), rotation first.
inline SE3 se3exp(const double upd[6]) {
  const double om[3] = {upd[0], upd[1], upd[2]};
  const double up[3] = {upd[3], upd[4], upd[5]};
  const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};  // skew(), se3_ops.hpp:27-47
  double Om2[9];
  mat3mul(Om, Om, Om2);
  double R[9], V[9];
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) { R[i] = I[i] + Om[i] + Om2[i]; V[i] = R[i]; }  // sic, se3quat.h:237-243
  } else {
    const double a = std::sin(theta) / theta;
    const double b = (1 - std::cos(theta)) / (theta * theta);
    const double c = (theta - std::sin(theta)) / (std::pow(theta, 3));
    for (int i = 0; i < 9; i++) { R[i] = I[i] + a * Om[i] + b * Om2[i]; V[i] = I[i] + b * Om[i] + c * Om2[i]; }
  }
  SE3 T;
  T.r = RtoQ(R);
  for (int i = 0; i < 3; i++) T.t[i] = V[i * 3 + 0] * up[0] + V[i * 3 + 1] * up[1] + V[i * 3 + 2] * up[2];
  normalizeRotation(T);  // SE3Quat(const Quaterniond&, const Vector3d&) ctor, se3quat.h:62-64
  return T;
}

// ---------------------------------------------------------------------------------------------
// graph
// ---------------------------------------------------------------------------------------------
struct Edge {
  int pose, point;
  bool stereo;
  double obs[3];
  double info;    // invSigma2: information = invSigma2 * I_d  (g2oOptimizer.cc:226-227, 262-264)
  double fx, fy, cx, cy, bf;
  int level = 0;
  bool robust = false;
  double delta = 0, dsqr = 0;  // RobustKernelHuber::setDelta, robust_kernel_impl.cpp:65-69.  NB the reference stores
                               // dsqr in a FLOAT member (robust_kernel_impl.h:84): delta^2 is rounded to float32 -- found by
                               // pinning against the prebuilt libg2o.so (oracle/pin_libg2o.py)
  double err[3] = {0, 0, 0};   // g2o's Edge::_error: only rewritten by computeError() on ACTIVE edges
  double Jl[9];                // _jacobianOplusXi (d x 3)
  double Jp[18];               // _jacobianOplusXj (d x 6)
  int hpl = -1;                // index of the Hpl block shared by all edges with the same (pose,landmark)
};

struct TraceRow { double pass, iter, trial, lambda, chi_before, chi_trial, rho, accepted; };

struct Graph {
  int n_pose = 0, n_point = 0, n_obs = 0;
  std::vector<SE3> pose;
  std::vector<uint8_t> fixed;
  std::vector<double> point;  // 3 per landmark
  std::vector<Edge> edges;    // insertion order == internalId order == _activeEdges order (sparse_optimizer.cpp:482-487)
  int threads = 1;

  // ---- active set / index mapping (sparse_optimizer.cpp:166-190, 199-267)
  std::vector<int> active;        // active edge ids, ascending
  std::vector<int> pose_slot;     // hessianIndex of pose (or -1: fixed / inactive)
  std::vector<int> point_slot;    // hessianIndex of landmark minus numPoses (or -1)
  std::vector<int> slot_pose, slot_point;
  int Np = 0, Nl = 0;

  // ---- BlockSolver<6,3> storage (block_solver.hpp:143-295)
  std::vector<double> Hpp, Hll, Hpl, b, x, coeff, bschur, Dinv;
  std::vector<int> hpl_slot;                 // pose slot of each Hpl block
  std::vector<int> lm_hpl_ptr, lm_hpl;       // CCS column per landmark: its Hpl blocks sorted by pose slot
  // reduced system, scalar skyline (envelope) storage of the upper triangle, row-wise on the transposed lower
  std::vector<int> sky_first;
  std::vector<int64_t> sky_ptr;
  std::vector<double> S, Sfac;
  std::vector<double> diagBackupPose, diagBackupLm;

  // ---- LM state (optimization_algorithm_levenberg.cpp:43-55)
  double lambda = -1, ni = 2;
  int nBad = 0;
  const volatile bool* stop = nullptr;

  // ---- state stack (push/pop/discardTop, sparse_optimizer.cpp:600-613; one level deep is all LM uses)
  std::vector<SE3> pose_bak;
  std::vector<double> point_bak;

  std::vector<TraceRow> trace;
  int cur_pass = 0;
  double t_solve_s = 0;

  // ---- lidar tight-coupling pass (g2oOptimizer.cc:979-1117): unary edges on the current keyframe, added after the
  // second visual pass; `lidar` holds the clouds for the kd-tree association, `uedges` the resulting edges
  struct UEdge {  // EdgeLidarFlatPoint / EdgeLidarCornerPoint, types_six_dof_expmap.h:206-262 (D = 1, no robust kernel)
    int pose;
    bool corner;
    double pc[3], qw[3], n[3];  // curpoint_cameraframe_, lastpoint_worldframe_, curr_point_norm
    double info;                // flat_optimized_weight / corner_optimized_weight
    double err = 0;
    double J[6];
  };
  std::vector<UEdge> uedges;
  bool uactive = false;      // the edges exist in the graph only from the third pass on
  bool unary_numeric = true; // BaseUnaryEdge::linearizeOplus central differences (the reference) vs closed form
  struct Lidar {
    bool set = false;
    int cur_pose = 0;
    std::vector<float> flat, flat_n, corner;                 // current keyframe, its own frame
    std::vector<float> map_flat, map_corner;                 // the other local keyframes, each point in its keyframe's frame
    std::vector<int32_t> map_flat_pose, map_corner_pose;
    double thr = 0, w_flat = 0, w_corner = 0;
    bool use_flat = false, use_corner = false;
    std::vector<int32_t> match;                              // per current point: matched map index or -1
  } lidar;

  bool terminate() const { return stop ? *stop : false; }  // sparse_optimizer.h:188
};

// EdgeSE3ProjectXYZ::cam_project / EdgeStereoSE3ProjectXYZ::cam_project
// (types_six_dof_expmap.cpp:141-157).  The stereo version takes `const float& bf` and computes
// `const float invz = 1.0f/trans_xyz[2]`: inverse depth is rounded to float32 and `bf*invz` is a
// float*float product.
inline void computeError(const Graph& g, Edge& e) {
  double Xc[3];
  se3map(g.pose[e.pose], &g.point[3 * e.point], Xc);
  if (!e.stereo) {
    // types_six_dof_expmap.h:90-95 + project2d (.cpp:37-42)
    const double px = Xc[0] / Xc[2], py = Xc[1] / Xc[2];
    e.err[0] = e.obs[0] - (px * e.fx + e.cx);
    e.err[1] = e.obs[1] - (py * e.fy + e.cy);
    e.err[2] = 0;
  } else {
    // types_six_dof_expmap.h:122-127 + .cpp:150-157
    const float bf_f = (float)e.bf;
    const float invz = (float)(1.0f / Xc[2]);
    double res[3];
    res[0] = Xc[0] * invz * e.fx + e.cx;
    res[1] = Xc[1] * invz * e.fy + e.cy;
    res[2] = res[0] - (double)(bf_f * invz);
    e.err[0] = e.obs[0] - res[0];
    e.err[1] = e.obs[1] - res[1];
    e.err[2] = e.obs[2] - res[2];
  }
}

// BaseEdge::chi2  (base_edge.h:58-61): _error.dot(information()*_error), information = info*I
inline double chi2(const Edge& e) {
  if (!e.stereo) return e.err[0] * (e.info * e.err[0]) + e.err[1] * (e.info * e.err[1]);
  return e.err[0] * (e.info * e.err[0]) + e.err[1] * (e.info * e.err[1]) + e.err[2] * (e.info * e.err[2]);
}

// isDepthPositive  (types_six_dof_expmap.h:97-101, 129-133)
inline bool depthPositive(const Graph& g, const Edge& e) {
  double Xc[3];
  se3map(g.pose[e.pose], &g.point[3 * e.point], Xc);
  return Xc[2] > 0.0;
}

// RobustKernelHuber::robustify  (robust_kernel_impl.cpp:78-91)
inline void robustify(const Edge& e, double c, double rho[3]) {
  if (c <= e.dsqr) {
    rho[0] = c; rho[1] = 1.; rho[2] = 0.;
  } else {
    const double sqrte = std::sqrt(c);
    rho[0] = 2 * sqrte * e.delta - e.dsqr;
    rho[1] = e.delta / sqrte;
    rho[2] = -0.5 * rho[1] / c;
  }
}

// linearizeOplus  (types_six_dof_expmap.cpp:103-139 mono, 188-234 stereo)
inline void linearize(const Graph& g, Edge& e) {
  const SE3& T = g.pose[e.pose];
  double Xc[3], R[9];
  se3map(T, &g.point[3 * e.point], Xc);
  qtoR(T.r, R);
  const double x = Xc[0], y = Xc[1], z = Xc[2], z_2 = z * z;
  const double fx = e.fx, fy = e.fy;
  double* Ji = e.Jl;
  double* Jj = e.Jp;
  if (!e.stereo) {
    // _jacobianOplusXi = -1./z * tmp * R   (.cpp:114-124)
    const double tmp[6] = {fx, 0, -x / z * fx, 0, fy, -y / z * fy};
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 3; c++) {
        const double tr = tmp[r * 3 + 0] * R[0 * 3 + c] + tmp[r * 3 + 1] * R[1 * 3 + c] + tmp[r * 3 + 2] * R[2 * 3 + c];
        Ji[r * 3 + c] = (-1. / z) * tr;
      }
    for (int c = 0; c < 3; c++) Ji[6 + c] = 0;
  } else {
    for (int c = 0; c < 3; c++) {
      Ji[0 * 3 + c] = -fx * R[0 * 3 + c] / z + fx * x * R[2 * 3 + c] / z_2;
      Ji[1 * 3 + c] = -fy * R[1 * 3 + c] / z + fy * y * R[2 * 3 + c] / z_2;
      Ji[2 * 3 + c] = Ji[0 * 3 + c] - e.bf * R[2 * 3 + c] / z_2;
    }
  }
  Jj[0] = x * y / z_2 * fx;
  Jj[1] = -(1 + (x * x / z_2)) * fx;
  Jj[2] = y / z * fx;
  Jj[3] = -1. / z * fx;
  Jj[4] = 0;
  Jj[5] = x / z_2 * fx;
  Jj[6] = (1 + y * y / z_2) * fy;
  Jj[7] = -x * y / z_2 * fy;
  Jj[8] = -x / z * fy;
  Jj[9] = 0;
  Jj[10] = -1. / z * fy;
  Jj[11] = y / z_2 * fy;
  if (e.stereo) {
    Jj[12] = Jj[0] - e.bf * y / z_2;
    Jj[13] = Jj[1] + e.bf * x / z_2;
    Jj[14] = Jj[2];
    Jj[15] = Jj[3];
    Jj[16] = 0;
    Jj[17] = Jj[5] - e.bf / z_2;
  } else {
    for (int c = 0; c < 6; c++) Jj[12 + c] = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// lidar unary edges (types_six_dof_expmap.h:206-262)
// ---------------------------------------------------------------------------------------------
// computeError of both edge types: with M = estimate.to_homogeneous_matrix().inverse(), Rwc = M.block(0,0,3,3),
// Ow = M.col(3).head(3):  d = Rwc.inverse() * (lastpoint_worldframe_ - Ow) - curpoint_cameraframe_;
// flat: d.dot(curr_point_norm) (:223-224), corner: d.norm() (:252).  The two matrix inverses of a rigid transform are
// restated in closed form (Rwc = R^T, Ow = -R^T t, Rwc^-1 = R); Eigen's general inverse differs from that by rounding.
inline double uError(const SE3& T, const Graph::UEdge& e) {
  double R[9];
  qtoR(T.r, R);
  double Ow[3], v[3], d[3];
  for (int i = 0; i < 3; i++) Ow[i] = -(R[0 * 3 + i] * T.t[0] + R[1 * 3 + i] * T.t[1] + R[2 * 3 + i] * T.t[2]);
  for (int i = 0; i < 3; i++) v[i] = e.qw[i] - Ow[i];
  for (int i = 0; i < 3; i++) d[i] = (R[i * 3 + 0] * v[0] + R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2]) - e.pc[i];
  if (e.corner) return std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  return d[0] * e.n[0] + d[1] * e.n[1] + d[2] * e.n[2];
}

// BaseUnaryEdge::linearizeOplus (base_unary_edge.hpp:82-123): central differences, delta = 1e-9, through the vertex's
// own oplus (push / oplus / computeError / pop per direction).  unary_numeric = false: the closed form
// d err / d(omega, upsilon) = [Xc x g, g] with Xc = R q + t and g = n (flat) or d/|d| (corner).
inline void uLinearize(const Graph& g, Graph::UEdge& e) {
  const SE3& T = g.pose[e.pose];
  if (g.unary_numeric) {
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    for (int d = 0; d < 6; d++) {
      double add[6] = {0, 0, 0, 0, 0, 0};
      add[d] = delta;
      const double e1 = uError(se3mul(se3exp(add), T), e);
      add[d] = -delta;
      const double e2 = uError(se3mul(se3exp(add), T), e);
      e.J[d] = scalar * (e1 - e2);
    }
    return;
  }
  double Xc[3], gr[3];
  se3map(T, e.qw, Xc);
  if (e.corner) {
    const double d[3] = {Xc[0] - e.pc[0], Xc[1] - e.pc[1], Xc[2] - e.pc[2]};
    const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int i = 0; i < 3; i++) gr[i] = n > 0 ? d[i] / n : 0.0;
  } else {
    for (int i = 0; i < 3; i++) gr[i] = e.n[i];
  }
  e.J[0] = Xc[1] * gr[2] - Xc[2] * gr[1];
  e.J[1] = Xc[2] * gr[0] - Xc[0] * gr[2];
  e.J[2] = Xc[0] * gr[1] - Xc[1] * gr[0];
  e.J[3] = gr[0]; e.J[4] = gr[1]; e.J[5] = gr[2];
}

// BaseUnaryEdge::constructQuadraticForm without robust kernel (base_unary_edge.hpp:58-61)
inline void uQuadraticForm(Graph& g, const Graph::UEdge& e) {
  const int ps = g.pose_slot[e.pose];
  if (ps < 0) return;
  double* bp = &g.b[(size_t)ps * 6];
  double* Hp = &g.Hpp[(size_t)ps * 36];
  for (int i = 0; i < 6; i++) {
    bp[i] -= e.J[i] * e.info * e.err;
    for (int j = 0; j < 6; j++) Hp[i * 6 + j] += e.J[i] * e.info * e.J[j];
  }
}

// a unary edge is active unless its only vertex is fixed (sparse_optimizer.cpp:218-235 allVerticesFixed)
inline bool uActive(const Graph& g, const Graph::UEdge& e) { return g.uactive && !g.fixed[e.pose]; }

// ---------------------------------------------------------------------------------------------
// SparseOptimizer::initializeOptimization(level) + buildIndexMapping + BlockSolver::buildStructure
// (sparse_optimizer.cpp:199-267, 166-190; block_solver.hpp:143-295)
// ---------------------------------------------------------------------------------------------
void initializeOptimization(Graph& g, int level) {
  g.active.clear();
  std::vector<uint8_t> pose_active(g.n_pose, 0), point_active(g.n_point, 0);
  for (int k = 0; k < (int)g.edges.size(); k++) {
    const Edge& e = g.edges[k];
    if (level >= 0 && e.level != level) continue;
    // allVerticesFixed(): landmark vertices are never fixed in BA, so every edge qualifies
    g.active.push_back(k);
    pose_active[e.pose] = 1;
    point_active[e.point] = 1;
  }
  if (level <= 0)  // the lidar edges are created with the default level 0
    for (const Graph::UEdge& e : g.uedges)
      if (uActive(g, e)) pose_active[e.pose] = 1;
  // index mapping: non-fixed, non-marginalised vertices first (poses), then marginalised (landmarks),
  // each in ascending vertex-id order (sortVectorContainers, sparse_optimizer.cpp:482-487)
  g.pose_slot.assign(g.n_pose, -1);
  g.point_slot.assign(g.n_point, -1);
  g.slot_pose.clear();
  g.slot_point.clear();
  for (int i = 0; i < g.n_pose; i++)
    if (pose_active[i] && !g.fixed[i]) { g.pose_slot[i] = (int)g.slot_pose.size(); g.slot_pose.push_back(i); }
  for (int i = 0; i < g.n_point; i++)
    if (point_active[i]) { g.point_slot[i] = (int)g.slot_point.size(); g.slot_point.push_back(i); }
  g.Np = (int)g.slot_pose.size();
  g.Nl = (int)g.slot_point.size();
}

void buildStructure(Graph& g) {
  const int Np = g.Np, Nl = g.Nl;
  g.Hpp.assign((size_t)Np * 36, 0.0);
  g.Hll.assign((size_t)Nl * 9, 0.0);
  g.Dinv.assign((size_t)Nl * 9, 0.0);
  g.b.assign((size_t)Np * 6 + (size_t)Nl * 3, 0.0);
  g.x.assign(g.b.size(), 0.0);
  g.coeff.assign(g.b.size(), 0.0);
  g.bschur.assign((size_t)Np * 6, 0.0);
  // Hpl blocks: one per distinct (pose slot, landmark slot) pair among active edges
  std::vector<std::vector<std::pair<int, int>>> per_lm(Nl);  // (pose slot, block id)
  g.hpl_slot.clear();
  for (int k : g.active) {
    Edge& e = g.edges[k];
    e.hpl = -1;
    const int ps = g.pose_slot[e.pose], ls = g.point_slot[e.point];
    if (ps < 0) continue;
    auto& lst = per_lm[ls];
    int found = -1;
    for (auto& pr : lst)
      if (pr.first == ps) { found = pr.second; break; }
    if (found < 0) {
      found = (int)g.hpl_slot.size();
      g.hpl_slot.push_back(ps);
      lst.emplace_back(ps, found);
    }
    e.hpl = found;
  }
  g.Hpl.assign(g.hpl_slot.size() * 18, 0.0);
  g.lm_hpl_ptr.assign(Nl + 1, 0);
  g.lm_hpl.clear();
  for (int l = 0; l < Nl; l++) {
    std::sort(per_lm[l].begin(), per_lm[l].end());  // SparseBlockMatrixCCS column: rows ascending
    for (auto& pr : per_lm[l]) g.lm_hpl.push_back(pr.second);
    g.lm_hpl_ptr[l + 1] = (int)g.lm_hpl.size();
  }
  // Schur pattern -> scalar skyline of the reduced system (upper triangle stored by column == lower by row)
  std::vector<int> first_blk(Np);
  for (int i = 0; i < Np; i++) first_blk[i] = i;
  for (int l = 0; l < Nl; l++) {
    const int a = g.lm_hpl_ptr[l], bnd = g.lm_hpl_ptr[l + 1];
    if (bnd <= a) continue;
    const int mn = g.hpl_slot[g.lm_hpl[a]];
    for (int k = a; k < bnd; k++) {
      const int s = g.hpl_slot[g.lm_hpl[k]];
      first_blk[s] = std::min(first_blk[s], mn);
    }
  }
  g.sky_first.assign((size_t)Np * 6, 0);
  g.sky_ptr.assign((size_t)Np * 6 + 1, 0);
  for (int i = 0; i < Np; i++)
    for (int r = 0; r < 6; r++) {
      const int row = i * 6 + r;
      g.sky_first[row] = first_blk[i] * 6;
      g.sky_ptr[row + 1] = g.sky_ptr[row] + (row - g.sky_first[row] + 1);
    }
  g.S.assign((size_t)g.sky_ptr[(size_t)Np * 6], 0.0);
  g.Sfac = g.S;
}

inline double& Sat(Graph& g, std::vector<double>& S, int row, int col) {  // row >= col
  return S[g.sky_ptr[row] + (col - g.sky_first[row])];
}

// SparseOptimizer::computeActiveErrors (sparse_optimizer.cpp:61-88)
void computeActiveErrors(Graph& g) {
  const int n = (int)g.active.size();
#pragma omp parallel for schedule(static) num_threads(g.threads) if (g.threads > 1)
  for (int k = 0; k < n; k++) computeError(g, g.edges[g.active[k]]);
  for (Graph::UEdge& e : g.uedges)
    if (uActive(g, e)) e.err = uError(g.pose[e.pose], e);
}

// SparseOptimizer::activeRobustChi2 (sparse_optimizer.cpp:100-114): sequential sum in edge order
double activeRobustChi2(const Graph& g) {
  double chi = 0.0, rho[3];
  for (int k : g.active) {
    const Edge& e = g.edges[k];
    if (e.robust) {
      robustify(e, chi2(e), rho);
      chi += rho[0];
    } else
      chi += chi2(e);
  }
  for (const Graph::UEdge& e : g.uedges)  // inserted after every visual edge -> last in _activeEdges
    if (uActive(g, e)) chi += e.err * (e.info * e.err);
  return chi;
}

// BaseBinaryEdge::constructQuadraticForm (base_binary_edge.hpp:55-120); A = Jl (vertex 0 = landmark),
// B = Jp (vertex 1 = pose).  Products are accumulated in Eigen's fixed-size evaluation order:
// (A^T * W) * A etc. with W = w * I.
inline void constructQuadraticForm(Graph& g, Edge& e, double* Hpp, double* bp_base) {
  const int d = e.stereo ? 3 : 2;
  const int ls = g.point_slot[e.point];
  const int ps = g.pose_slot[e.pose];
  double w = e.info;
  double omega_r[3];
  for (int i = 0; i < d; i++) omega_r[i] = -(e.info * e.err[i]);
  if (e.robust) {
    double rho[3];
    robustify(e, chi2(e), rho);
    w = rho[1] * e.info;  // robustInformation, base_edge.h:96-102 (second-order term commented out)
    for (int i = 0; i < d; i++) omega_r[i] *= rho[1];
  }
  const double* A = e.Jl;
  const double* B = e.Jp;
  // landmark: from->b += A^T omega_r ; from->A += A^T W A
  double* bl = &g.b[(size_t)g.Np * 6 + (size_t)ls * 3];
  double* Hl = &g.Hll[(size_t)ls * 9];
  for (int i = 0; i < 3; i++) {
    double s = 0;
    for (int r = 0; r < d; r++) s += A[r * 3 + i] * omega_r[r];
    bl[i] += s;
    for (int j = 0; j < 3; j++) {
      double h = 0;
      for (int r = 0; r < d; r++) h += (A[r * 3 + i] * w) * A[r * 3 + j];
      Hl[i * 3 + j] += h;
    }
  }
  if (ps >= 0) {
    double* Wpl = &g.Hpl[(size_t)e.hpl * 18];  // 6x3: B^T W A
    double* bp = bp_base + (size_t)ps * 6;
    double* Hp = Hpp + (size_t)ps * 36;
    for (int i = 0; i < 6; i++) {
      double s = 0;
      for (int r = 0; r < d; r++) s += B[r * 6 + i] * omega_r[r];
      bp[i] += s;
      for (int j = 0; j < 6; j++) {
        double h = 0;
        for (int r = 0; r < d; r++) h += (B[r * 6 + i] * w) * B[r * 6 + j];
        Hp[i * 6 + j] += h;
      }
      for (int j = 0; j < 3; j++) {
        double h = 0;
        for (int r = 0; r < d; r++) h += (B[r * 6 + i] * w) * A[r * 3 + j];
        Wpl[i * 3 + j] += h;
      }
    }
  }
}

// BlockSolver::buildSystem (block_solver.hpp:502-560)
void buildSystemVisual(Graph& g);
void buildSystem(Graph& g) {
  buildSystemVisual(g);
  for (Graph::UEdge& e : g.uedges)
    if (uActive(g, e)) {
      uLinearize(g, e);
      uQuadraticForm(g, e);
    }
}
void buildSystemVisual(Graph& g) {
  std::fill(g.b.begin(), g.b.end(), 0.0);
  std::fill(g.Hpp.begin(), g.Hpp.end(), 0.0);
  std::fill(g.Hll.begin(), g.Hll.end(), 0.0);
  std::fill(g.Hpl.begin(), g.Hpl.end(), 0.0);
  const int n = (int)g.active.size();
  if (g.threads <= 1) {
    for (int k = 0; k < n; k++) {
      Edge& e = g.edges[g.active[k]];
      linearize(g, e);
      constructQuadraticForm(g, e, g.Hpp.data(), g.b.data());
    }
    return;
  }
  // best-effort multi-core variant (g2o's dormant G2O_OPENMP site, block_solver.hpp:526-528): Jacobians in
  // parallel, landmark-side accumulation is race-free because edges arrive grouped by landmark; pose-side
  // accumulators are per-thread and reduced afterwards (changes summation order => rounding only).
#pragma omp parallel for schedule(static) num_threads(g.threads)
  for (int k = 0; k < n; k++) linearize(g, g.edges[g.active[k]]);
  // split active edges into landmark-aligned chunks
  std::vector<int> cut(g.threads + 1, n);
  cut[0] = 0;
  for (int t = 1; t < g.threads; t++) {
    int c = (int)((int64_t)n * t / g.threads);
    while (c < n && c > 0 && g.edges[g.active[c]].point == g.edges[g.active[c - 1]].point) c++;
    cut[t] = std::max(c, cut[t - 1]);
  }
  std::vector<std::vector<double>> Hloc(g.threads), bloc(g.threads);
#pragma omp parallel num_threads(g.threads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    Hloc[t].assign(g.Hpp.size(), 0.0);
    bloc[t].assign((size_t)g.Np * 6, 0.0);
    for (int k = cut[t]; k < cut[t + 1]; k++) constructQuadraticForm(g, g.edges[g.active[k]], Hloc[t].data(), bloc[t].data());
  }
  for (int t = 0; t < g.threads; t++) {
    for (size_t i = 0; i < g.Hpp.size(); i++) g.Hpp[i] += Hloc[t][i];
    for (size_t i = 0; i < (size_t)g.Np * 6; i++) g.b[i] += bloc[t][i];
  }
}

// Eigen::Matrix3d::inverse() (cofactor form, compute_inverse_size3_helper)  (block_solver.hpp:389)
inline void inv3(const double* m, double* inv) {
  const double c00 = m[4] * m[8] - m[5] * m[7];
  const double c10 = m[5] * m[6] - m[3] * m[8];
  const double c20 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c10 + m[2] * c20;
  const double invdet = 1.0 / det;
  inv[0] = c00 * invdet;
  inv[3] = c10 * invdet;
  inv[6] = c20 * invdet;
  inv[1] = (m[2] * m[7] - m[1] * m[8]) * invdet;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) * invdet;
  inv[7] = (m[1] * m[6] - m[0] * m[7]) * invdet;
  inv[2] = (m[1] * m[5] - m[2] * m[4]) * invdet;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) * invdet;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) * invdet;
}

// Eigen::SimplicialLDLT (no pivoting; fails on a zero pivot) on the skyline-stored reduced system
// (linear_solver_eigen.h:94-124).  AMD ordering is irrelevant to the exact-arithmetic result.
bool ldltSolve(Graph& g, const double* rhs, double* x) {
  const int n = g.Np * 6;
  g.Sfac = g.S;
  std::vector<double>& L = g.Sfac;  // unit-lower L strictly below the diagonal, D on the diagonal
  std::vector<double> dinv(n);
  for (int i = 0; i < n; i++) {
    const int fi = g.sky_first[i];
    double* Li = &L[g.sky_ptr[i]] - fi;  // Li[c] = element (i,c)
    // pass 1: Y_ij = L_ij*d_j = A_ij - sum_k Y_ik L_jk   (finished rows j already hold L_jk)
    for (int j = fi; j < i; j++) {
      const int fj = g.sky_first[j];
      const double* Lj = &L[g.sky_ptr[j]] - fj;
      double s = Li[j];
      for (int k = std::max(fi, fj); k < j; k++) s -= Li[k] * Lj[k];
      Li[j] = s;
    }
    // pass 2: d_i = A_ii - sum_j Y_ij L_ij ; convert the row from Y to L
    double d = Li[i];
    for (int j = fi; j < i; j++) {
      const double lij = Li[j] * dinv[j];
      d -= Li[j] * lij;
      Li[j] = lij;
    }
    if (d == 0.0 || !std::isfinite(d)) return false;  // Eigen: info() != Success on a zero pivot
    dinv[i] = 1.0 / d;
    Li[i] = d;
  }
  for (int i = 0; i < n; i++) x[i] = rhs[i];
  for (int i = 0; i < n; i++) {
    const int fi = g.sky_first[i];
    const double* Li = &L[g.sky_ptr[i]] - fi;
    double s = x[i];
    for (int k = fi; k < i; k++) s -= Li[k] * x[k];
    x[i] = s;
  }
  for (int i = 0; i < n; i++) x[i] *= dinv[i];
  for (int i = n - 1; i >= 0; i--) {
    const int fi = g.sky_first[i];
    const double* Li = &L[g.sky_ptr[i]] - fi;
    const double xi = x[i];
    for (int k = fi; k < i; k++) x[k] -= Li[k] * xi;
  }
  return true;
}

// BlockSolver::setLambda / restoreDiagonal (block_solver.hpp:564-604): lambda on EVERY diagonal of Hpp and Hll
void setLambda(Graph& g, double lambda) {
  g.diagBackupPose.resize((size_t)g.Np * 6);
  g.diagBackupLm.resize((size_t)g.Nl * 3);
  for (int i = 0; i < g.Np; i++)
    for (int r = 0; r < 6; r++) {
      double& dd = g.Hpp[(size_t)i * 36 + r * 6 + r];
      g.diagBackupPose[(size_t)i * 6 + r] = dd;
      dd += lambda;
    }
  for (int i = 0; i < g.Nl; i++)
    for (int r = 0; r < 3; r++) {
      double& dd = g.Hll[(size_t)i * 9 + r * 3 + r];
      g.diagBackupLm[(size_t)i * 3 + r] = dd;
      dd += lambda;
    }
}
void restoreDiagonal(Graph& g) {
  for (int i = 0; i < g.Np; i++)
    for (int r = 0; r < 6; r++) g.Hpp[(size_t)i * 36 + r * 6 + r] = g.diagBackupPose[(size_t)i * 6 + r];
  for (int i = 0; i < g.Nl; i++)
    for (int r = 0; r < 3; r++) g.Hll[(size_t)i * 9 + r * 3 + r] = g.diagBackupLm[(size_t)i * 3 + r];
}

// Schur complement of one landmark into (S, coeff)  (block_solver.hpp:381-432)
inline void schurLandmark(Graph& g, int l, std::vector<double>& S, double* coeff) {
  const int Np = g.Np;
  double* Dinv = &g.Dinv[(size_t)l * 9];
  inv3(&g.Hll[(size_t)l * 9], Dinv);
  const double* bl = &g.b[(size_t)Np * 6 + (size_t)l * 3];
  double db[3];
  for (int i = 0; i < 3; i++) db[i] = Dinv[i * 3 + 0] * bl[0] + Dinv[i * 3 + 1] * bl[1] + Dinv[i * 3 + 2] * bl[2];
  const int a = g.lm_hpl_ptr[l], bnd = g.lm_hpl_ptr[l + 1];
  for (int ko = a; ko < bnd; ko++) {
    const int blk1 = g.lm_hpl[ko];
    const int i1 = g.hpl_slot[blk1];
    const double* Bi = &g.Hpl[(size_t)blk1 * 18];
    double BDinv[18];
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 3; c++)
        BDinv[r * 3 + c] = Bi[r * 3 + 0] * Dinv[0 * 3 + c] + Bi[r * 3 + 1] * Dinv[1 * 3 + c] + Bi[r * 3 + 2] * Dinv[2 * 3 + c];
    for (int r = 0; r < 6; r++) coeff[(size_t)i1 * 6 + r] += Bi[r * 3 + 0] * db[0] + Bi[r * 3 + 1] * db[1] + Bi[r * 3 + 2] * db[2];
    for (int ki = ko; ki < bnd; ki++) {  // upper triangle only: i2 >= i1
      const int blk2 = g.lm_hpl[ki];
      const int i2 = g.hpl_slot[blk2];
      const double* Bj = &g.Hpl[(size_t)blk2 * 18];
      // Hschur(i1,i2) -= BDinv * Bj^T ; stored as the lower-triangle element (row of i2, col of i1) transposed
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
          const double v = BDinv[r * 3 + 0] * Bj[c * 3 + 0] + BDinv[r * 3 + 1] * Bj[c * 3 + 1] + BDinv[r * 3 + 2] * Bj[c * 3 + 2];
          const int gr = i1 * 6 + r, gc = i2 * 6 + c;  // gr <= gc within the upper triangle unless i1==i2
          if (gc >= gr) S[g.sky_ptr[gc] + (gr - g.sky_first[gc])] -= v;
        }
    }
  }
}

// BlockSolver::solve (block_solver.hpp:354-486)
bool blockSolve(Graph& g) {
  const int Np = g.Np, Nl = g.Nl;
  // _Hschur = _Hpp (upper triangle)
  std::fill(g.S.begin(), g.S.end(), 0.0);
  for (int i = 0; i < Np; i++)
    for (int r = 0; r < 6; r++)
      for (int c = r; c < 6; c++) Sat(g, g.S, i * 6 + c, i * 6 + r) = g.Hpp[(size_t)i * 36 + r * 6 + c];
  std::fill(g.coeff.begin(), g.coeff.begin() + (size_t)Np * 6, 0.0);
  if (g.threads <= 1) {
    for (int l = 0; l < Nl; l++) schurLandmark(g, l, g.S, g.coeff.data());
  } else {
    // g2o's dormant `#pragma omp parallel for` over landmarks (block_solver.hpp:378-380), with per-thread
    // accumulators instead of per-block mutexes
    std::vector<std::vector<double>> Sl(g.threads), cl(g.threads);
#pragma omp parallel num_threads(g.threads)
    {
#ifdef _OPENMP
      const int t = omp_get_thread_num();
#else
      const int t = 0;
#endif
      Sl[t].assign(g.S.size(), 0.0);
      cl[t].assign((size_t)Np * 6, 0.0);
#pragma omp for schedule(static)
      for (int l = 0; l < Nl; l++) schurLandmark(g, l, Sl[t], cl[t].data());
    }
    for (int t = 0; t < g.threads; t++) {
      for (size_t i = 0; i < g.S.size(); i++) g.S[i] += Sl[t][i];
      for (size_t i = 0; i < (size_t)Np * 6; i++) g.coeff[i] += cl[t][i];
    }
  }
  for (int i = 0; i < Np * 6; i++) g.bschur[i] = g.b[i] - g.coeff[i];
  const bool ok = ldltSolve(g, g.bschur.data(), g.x.data());
  if (!ok) return false;
  // landmark back-substitution (block_solver.hpp:461-483): xl = Dinv * (bl - Hpl^T xp)
  double* xl = g.x.data() + (size_t)Np * 6;
  const double* bl = g.b.data() + (size_t)Np * 6;
#pragma omp parallel for schedule(static) num_threads(g.threads) if (g.threads > 1)
  for (int l = 0; l < Nl; l++) {
    double cl[3] = {bl[l * 3 + 0], bl[l * 3 + 1], bl[l * 3 + 2]};
    for (int k = g.lm_hpl_ptr[l]; k < g.lm_hpl_ptr[l + 1]; k++) {
      const int blk = g.lm_hpl[k];
      const double* Bi = &g.Hpl[(size_t)blk * 18];
      const double* xp = &g.x[(size_t)g.hpl_slot[blk] * 6];
      for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int r = 0; r < 6; r++) s += Bi[r * 3 + c] * (-xp[r]);  // cp = -xp; rightMultiply adds B^T * cp
        cl[c] += s;
      }
    }
    const double* Dinv = &g.Dinv[(size_t)l * 9];
    for (int i = 0; i < 3; i++) xl[l * 3 + i] = Dinv[i * 3 + 0] * cl[0] + Dinv[i * 3 + 1] * cl[1] + Dinv[i * 3 + 2] * cl[2];
  }
  return true;
}

// SparseOptimizer::update (sparse_optimizer.cpp:422-435) -> oplusImpl
void updateState(Graph& g, const double* upd) {
  for (int s = 0; s < g.Np; s++) {
    SE3& T = g.pose[g.slot_pose[s]];
    T = se3mul(se3exp(upd + (size_t)s * 6), T);  // VertexSE3Expmap::oplusImpl, types_six_dof_expmap.h:73-76
  }
  const double* ul = upd + (size_t)g.Np * 6;
  for (int s = 0; s < g.Nl; s++) {
    double* X = &g.point[(size_t)g.slot_point[s] * 3];  // VertexSBAPointXYZ::oplusImpl, types_sba.h:52-56
    X[0] += ul[s * 3 + 0]; X[1] += ul[s * 3 + 1]; X[2] += ul[s * 3 + 2];
  }
}

void pushState(Graph& g) { g.pose_bak = g.pose; g.point_bak = g.point; }  // only indexed vertices change, so a full copy is equivalent
void popState(Graph& g) { g.pose = g.pose_bak; g.point = g.point_bak; }

// OptimizationAlgorithmLevenberg::computeLambdaInit (optimization_algorithm_levenberg.cpp:166-180)
double computeLambdaInit(const Graph& g) {
  double maxDiagonal = 0.;
  for (int i = 0; i < g.Np; i++)
    for (int j = 0; j < 6; j++) maxDiagonal = std::max(std::fabs(g.Hpp[(size_t)i * 36 + j * 6 + j]), maxDiagonal);
  for (int i = 0; i < g.Nl; i++)
    for (int j = 0; j < 3; j++) maxDiagonal = std::max(std::fabs(g.Hll[(size_t)i * 9 + j * 3 + j]), maxDiagonal);
  return 1e-5 * maxDiagonal;  // _tau
}

enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };

// OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-164)
SolverResult lmSolve(Graph& g, int iteration) {
  if (iteration == 0) buildStructure(g);
  computeActiveErrors(g);
  double currentChi = activeRobustChi2(g);
  double tempChi = currentChi;
  const double iniChi = currentChi;
  buildSystem(g);
  if (iteration == 0) {
    g.lambda = computeLambdaInit(g);
    g.ni = 2;
    g.nBad = 0;
  }
  double rho = 0;
  int qmax = 0;
  do {
    pushState(g);
    setLambda(g, g.lambda);
    const bool ok2 = blockSolve(g);
    updateState(g, g.x.data());
    restoreDiagonal(g);
    computeActiveErrors(g);
    tempChi = activeRobustChi2(g);
    if (!ok2) tempChi = std::numeric_limits<double>::max();
    rho = (currentChi - tempChi);
    double scale = 0.;  // computeScale, :182-189
    for (size_t j = 0; j < g.x.size(); j++) scale += g.x[j] * (g.lambda * g.x[j] + g.b[j]);
    scale += 1e-3;
    rho /= scale;
    TraceRow tr{(double)g.cur_pass, (double)iteration, (double)qmax, g.lambda, currentChi, tempChi, rho, 0.0};
    if (rho > 0 && std::isfinite(tempChi)) {
      double alpha = 1. - std::pow((2 * rho - 1), 3);
      alpha = std::min(alpha, 2. / 3.);
      const double scaleFactor = std::max(1. / 3., alpha);
      g.lambda *= scaleFactor;
      g.ni = 2;
      currentChi = tempChi;
      tr.accepted = 1.0;  // discardTop
    } else {
      g.lambda *= g.ni;
      g.ni *= 2;
      popState(g);  // NB: edge errors keep the rejected trial's values (stale _error)
    }
    g.trace.push_back(tr);
    qmax++;
  } while (rho < 0 && qmax < 10 && !g.terminate());
  if (qmax == 10 || rho == 0) return Terminate;
  if ((iniChi - currentChi) * 1e3 < iniChi)
    g.nBad++;
  else
    g.nBad = 0;
  if (g.nBad >= 3) return Terminate;
  return OK;
}

// SparseOptimizer::optimize (sparse_optimizer.cpp:354-419)
int optimize(Graph& g, int iterations) {
  if (g.Np + g.Nl == 0) return -1;
  int cj = 0;
  bool ok = true;
  for (int i = 0; i < iterations && !g.terminate() && ok; i++) {
    const SolverResult r = lmSolve(g, i);
    ok = (r == OK);
    ++cj;
  }
  return cj;
}

}  // namespace

// =============================================================================================
// C API (ctypes / bench)
// =============================================================================================
extern "C" {

struct refba { Graph g; std::vector<uint8_t> outlier; };

// Graph construction mirrors g2oOptimizer::BundleAdjustment / LocalBundleAdjustment edge wiring
// (src/backend/g2oOptimizer.cc:210-283, 875-916): obs with ur<0 -> EdgeSE3ProjectXYZ, else EdgeStereoSE3ProjectXYZ.
refba* refba_create(int n_pose, int n_point, int n_obs, const double* pose_qt, const uint8_t* pose_fixed,
                    const double* point_xyz, const int32_t* obs_pose, const int32_t* obs_point,
                    const float* obs_meas, const double* cam) {
  refba* h = new refba();
  Graph& g = h->g;
  g.n_pose = n_pose; g.n_point = n_point; g.n_obs = n_obs;
  g.pose.resize(n_pose);
  g.fixed.assign(pose_fixed, pose_fixed + n_pose);
  for (int i = 0; i < n_pose; i++) {
    const double* v = pose_qt + (size_t)i * 7;
    SE3 T;
    T.t[0] = v[0]; T.t[1] = v[1]; T.t[2] = v[2];
    T.r = Quat{v[3], v[4], v[5], v[6]};
    normalizeRotation(T);  // SE3Quat(R,t) ctor
    g.pose[i] = T;
  }
  g.point.assign(point_xyz, point_xyz + (size_t)n_point * 3);
  g.edges.resize(n_obs);
  for (int k = 0; k < n_obs; k++) {
    Edge& e = g.edges[k];
    e.pose = obs_pose[k];
    e.point = obs_point[k];
    const float* m = obs_meas + (size_t)k * 4;
    e.stereo = !(m[2] < 0);
    e.obs[0] = m[0]; e.obs[1] = m[1]; e.obs[2] = e.stereo ? m[2] : 0.0;
    e.info = m[3];
    const double* c = cam + (size_t)e.pose * 5;
    e.fx = c[0]; e.fy = c[1]; e.cx = c[2]; e.cy = c[3]; e.bf = c[4];
  }
  h->outlier.assign(n_obs, 0);
  return h;
}

void refba_destroy(refba* h) { delete h; }
void refba_set_threads(refba* h, int t) { h->g.threads = t < 1 ? 1 : t; }

static void setHuber(Edge& e, float th2d, float th3d, bool robust) {
  e.robust = robust;
  const double d = e.stereo ? (double)th3d : (double)th2d;  // `const float thHuber.. = sqrt(..)` then setDelta(double)
  e.delta = d;
  e.dsqr = (double)(float)(d * d);  // `float dsqr`, robust_kernel_impl.h:84
}

// g2oOptimizer::LocalBundleAdjustment control flow (g2oOptimizer.cc:704-976, 1119-1142) with the stereo edge
// wired as in ::BundleAdjustment (:247-283) and upstream ORB-SLAM2's stereo thresholds (thHuberStereo :853, 7.815).
// third_pass_iters = 0 (north_star default) or 20 (the fork's unconditional lidar-coupling pass, :1113-1114).
// ---- lidar pass: association (g2oOptimizer.cc:981-1107) ------------------------------------------------------------
// Twc of a keyframe the way the reference gets it: Converter::toCvMat(SE3Quat) (double -> CV_32F 4x4, Converter.cc)
// followed by cv::Mat::inv().  OpenCV inverts the float matrix by LU; here the inverse of the float-rounded rigid
// transform is written in closed form, evaluated in double and rounded to float (agrees to float rounding).
static void twcFloat(const SE3& T, float Rwc[9], float Ow[3]) {
  double R[9];
  qtoR(T.r, R);
  float Rf[9], tf[3];
  for (int i = 0; i < 9; i++) Rf[i] = (float)R[i];
  for (int i = 0; i < 3; i++) tf[i] = (float)T.t[i];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Rwc[i * 3 + j] = Rf[j * 3 + i];
  for (int i = 0; i < 3; i++)
    Ow[i] = (float)(-((double)Rf[0 * 3 + i] * tf[0] + (double)Rf[1 * 3 + i] * tf[1] + (double)Rf[2 * 3 + i] * tf[2]));
}
// pcl::transformPointCloud(in, out, Eigen::Affine3d): the product is evaluated in double and cast to the float fields
static void toWorldFloat(const float Rwc[9], const float Ow[3], const float* p, float* out) {
  for (int i = 0; i < 3; i++)
    out[i] = (float)((double)Rwc[i * 3 + 0] * p[0] + (double)Rwc[i * 3 + 1] * p[1] + (double)Rwc[i * 3 + 2] * p[2] + (double)Ow[i]);
}
// pcl::KdTreeFLANN<PointI>::nearestKSearch(k=1): exact nearest neighbour under flann::L2_Simple<float> (float
// accumulation of squared differences, no fused multiply-add in a generic x86-64 build).  Brute force here; ties go
// to the smallest index (a kd-tree may return either).
__attribute__((optimize("fp-contract=off"))) static int nearestFloat(const std::vector<float>& map, const float* q, float* d2out) {
  int best = -1;
  float bd = std::numeric_limits<float>::infinity();
  const size_t n = map.size() / 3;
  for (size_t i = 0; i < n; i++) {
    float r = 0.f;
    for (int k = 0; k < 3; k++) {
      const float diff = q[k] - map[i * 3 + k];
      r += diff * diff;
    }
    if (r < bd) { bd = r; best = (int)i; }
  }
  *d2out = bd;
  return best;
}

static void lidarAssociate(Graph& g) {
  Graph::Lidar& L = g.lidar;
  g.uedges.clear();
  const size_t nf = L.flat.size() / 3, nc = L.corner.size() / 3;
  L.match.assign(nf + nc, -1);
  auto worldMap = [&](const std::vector<float>& pts, const std::vector<int32_t>& pose) {
    std::vector<float> out(pts.size());
    for (size_t i = 0; i < pose.size(); i++) {
      float Rwc[9], Ow[3];
      twcFloat(g.pose[pose[i]], Rwc, Ow);
      toWorldFloat(Rwc, Ow, &pts[i * 3], &out[i * 3]);
    }
    return out;
  };
  float Rc[9], Oc[3];
  twcFloat(g.pose[L.cur_pose], Rc, Oc);
  auto pass = [&](const std::vector<float>& cur, const std::vector<float>& mapw, bool corner, size_t off) {
    if (mapw.empty()) return;
    for (size_t i = 0; i < cur.size() / 3; i++) {
      float qw[3], d2;
      toWorldFloat(Rc, Oc, &cur[i * 3], qw);
      const int j = nearestFloat(mapw, qw, &d2);
      if (!(d2 < L.thr)) continue;  // float compared against the double threshold, :1049/:1087
      L.match[off + i] = j;
      Graph::UEdge e;
      e.pose = L.cur_pose;
      e.corner = corner;
      for (int k = 0; k < 3; k++) {
        e.pc[k] = cur[i * 3 + k];
        e.qw[k] = mapw[(size_t)j * 3 + k];
        e.n[k] = corner ? 0.0 : L.flat_n[i * 3 + k];
      }
      e.info = corner ? L.w_corner : L.w_flat;
      g.uedges.push_back(e);
    }
  };
  if (L.use_flat) pass(L.flat, worldMap(L.map_flat, L.map_flat_pose), false, 0);
  if (L.use_corner) pass(L.corner, worldMap(L.map_corner, L.map_corner_pose), true, nf);
}

// explicit correspondences (what the association would produce); w[i] == 0 -> no edge.  Flat edges first.
void refba_set_lidar_edges(refba* h, int cur_pose, int n_flat, int n_corner, const double* pc, const double* qw,
                           const double* normal, const double* w, int numeric_jacobian) {
  Graph& g = h->g;
  g.lidar.set = false;
  g.uedges.clear();
  g.unary_numeric = numeric_jacobian != 0;
  for (int i = 0; i < n_flat + n_corner; i++) {
    if (!(w[i] > 0)) continue;
    Graph::UEdge e;
    e.pose = cur_pose;
    e.corner = i >= n_flat;
    for (int k = 0; k < 3; k++) {
      e.pc[k] = pc[i * 3 + k];
      e.qw[k] = qw[i * 3 + k];
      e.n[k] = e.corner ? 0.0 : normal[i * 3 + k];
    }
    e.info = w[i];
    g.uedges.push_back(e);
  }
}

// the clouds of the lidar pass: current keyframe features (own frame) and the features of the other local keyframes
// (each in its keyframe's frame, with the pose index of that keyframe); thresholds/weights from lidarConfig
void refba_set_lidar(refba* h, int cur_pose, int n_flat, const float* flat_xyz, const float* flat_normal, int n_corner,
                     const float* corner_xyz, int64_t n_map_flat, const float* map_flat_xyz, const int32_t* map_flat_pose,
                     int64_t n_map_corner, const float* map_corner_xyz, const int32_t* map_corner_pose,
                     double distance_sq_threshold, double flat_weight, double corner_weight, int use_flat, int use_corner,
                     int numeric_jacobian) {
  Graph::Lidar& L = h->g.lidar;
  L.set = true;
  L.cur_pose = cur_pose;
  L.flat.assign(flat_xyz, flat_xyz + (size_t)n_flat * 3);
  L.flat_n.assign(flat_normal, flat_normal + (size_t)n_flat * 3);
  L.corner.assign(corner_xyz, corner_xyz + (size_t)n_corner * 3);
  L.map_flat.assign(map_flat_xyz, map_flat_xyz + (size_t)n_map_flat * 3);
  L.map_flat_pose.assign(map_flat_pose, map_flat_pose + n_map_flat);
  L.map_corner.assign(map_corner_xyz, map_corner_xyz + (size_t)n_map_corner * 3);
  L.map_corner_pose.assign(map_corner_pose, map_corner_pose + n_map_corner);
  L.thr = distance_sq_threshold; L.w_flat = flat_weight; L.w_corner = corner_weight;
  L.use_flat = use_flat != 0; L.use_corner = use_corner != 0;
  h->g.unary_numeric = numeric_jacobian != 0;
  h->g.uedges.clear();
}
int refba_lidar_num_matches(refba* h) { return (int)h->g.lidar.match.size(); }
void refba_get_lidar_matches(refba* h, int32_t* out) {
  std::memcpy(out, h->g.lidar.match.data(), h->g.lidar.match.size() * sizeof(int32_t));
}
int refba_num_lidar_edges(refba* h) { return (int)h->g.uedges.size(); }

int refba_solve_local(refba* h, const volatile bool* stop, int third_pass_iters) {
  Graph& g = h->g;
  g.stop = stop;
  g.uactive = false;
  g.trace.clear();
  const float thHuberMono = std::sqrt(5.991);     // :851 `const float`
  const float thHuberStereo = std::sqrt(7.815);   // :853
  for (Edge& e : g.edges) { e.level = 0; setHuber(e, thHuberMono, thHuberStereo, true); }
  if (stop && *stop) return 0;  // :923-928 early-out, map untouched
  auto t0 = std::chrono::steady_clock::now();
  g.cur_pass = 0;
  initializeOptimization(g, 0);
  optimize(g, 5);
  bool bDoMore = true;
  if (stop && *stop) bDoMore = false;
  if (bDoMore) {
    for (Edge& e : g.edges) {  // :947-970
      const double thr = e.stereo ? 7.815 : 5.991;
      if (chi2(e) > thr || !depthPositive(g, e)) e.level = 1;
      e.robust = false;
    }
    g.cur_pass = 1;
    initializeOptimization(g, 0);
    optimize(g, 10);
  }
  if (third_pass_iters > 0) {
    g.cur_pass = 2;
    if (g.lidar.set) lidarAssociate(g);  // local lidar map + kd-tree matches at the pass-2 estimates, :981-1107
    g.uactive = !g.uedges.empty();
    initializeOptimization(g, 0);
    optimize(g, third_pass_iters);
  }
  for (size_t k = 0; k < g.edges.size(); k++) {  // :1125-1142
    const Edge& e = g.edges[k];
    const double thr = e.stereo ? 7.815 : 5.991;
    h->outlier[k] = (chi2(e) > thr || !depthPositive(g, e)) ? 1 : 0;
  }
  g.t_solve_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 1;
}

// g2oOptimizer::BundleAdjustment (g2oOptimizer.cc:110-362): single pass, optional Huber with
// thHuber2D = sqrt(5.99), thHuber3D = sqrt(7.815) (:163-164), no outlier step.
int refba_solve_global(refba* h, int iters, int robust, const volatile bool* stop) {
  Graph& g = h->g;
  g.stop = stop;
  g.trace.clear();
  const float thHuber2D = std::sqrt(5.99);
  const float thHuber3D = std::sqrt(7.815);
  for (Edge& e : g.edges) { e.level = 0; setHuber(e, thHuber2D, thHuber3D, robust != 0); }
  auto t0 = std::chrono::steady_clock::now();
  g.cur_pass = 0;
  initializeOptimization(g, 0);
  optimize(g, iters);
  for (size_t k = 0; k < g.edges.size(); k++) h->outlier[k] = 0;
  g.t_solve_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 1;
}

double refba_solve_seconds(refba* h) { return h->g.t_solve_s; }

void refba_get_poses(refba* h, double* out) {
  for (int i = 0; i < h->g.n_pose; i++) {  // SE3Quat::toVector order
    const SE3& T = h->g.pose[i];
    double* v = out + (size_t)i * 7;
    v[0] = T.t[0]; v[1] = T.t[1]; v[2] = T.t[2]; v[3] = T.r.x; v[4] = T.r.y; v[5] = T.r.z; v[6] = T.r.w;
  }
}
void refba_get_points(refba* h, double* out) { std::memcpy(out, h->g.point.data(), h->g.point.size() * sizeof(double)); }
void refba_get_outliers(refba* h, uint8_t* out) { std::memcpy(out, h->outlier.data(), h->outlier.size()); }
int refba_trace_len(refba* h) { return (int)h->g.trace.size(); }
void refba_get_trace(refba* h, double* out) { std::memcpy(out, h->g.trace.data(), h->g.trace.size() * sizeof(TraceRow)); }

// ---- stage-level introspection used by the kernel parity tests --------------------------------

// One "phase" of graph-level semantics, used to pin the oracle against a REAL g2o::SparseOptimizer of the reference's
// prebuilt binary (oracle/pin_libg2o_graph.py): set the edge levels and robust flags, initializeOptimization(0),
// computeActiveErrors(), report the index mapping (hessianIndex of every vertex; landmarks offset by the number of
// free poses as in g2o), every edge's stored _error (inactive edges keep what they had), activeChi2 and
// activeRobustChi2, then -- if `update` is given -- SparseOptimizer::update(update) and the resulting estimates.
void refba_debug_phase(refba* h, const int32_t* levels, int robust, int32_t* pose_index, int32_t* point_index,
                       double* err, double* chi2_out, const double* update, double* poses, double* points) {
  Graph& g = h->g;
  const float thHuberMono = std::sqrt(5.991), thHuberStereo = std::sqrt(7.815);
  for (size_t k = 0; k < g.edges.size(); k++) {
    g.edges[k].level = levels[k];
    setHuber(g.edges[k], thHuberMono, thHuberStereo, robust != 0);
  }
  initializeOptimization(g, 0);
  computeActiveErrors(g);
  for (int i = 0; i < g.n_pose; i++) pose_index[i] = g.pose_slot[i];
  for (int i = 0; i < g.n_point; i++) point_index[i] = g.point_slot[i] < 0 ? -1 : g.Np + g.point_slot[i];
  for (size_t k = 0; k < g.edges.size(); k++)
    for (int c = 0; c < 3; c++) err[k * 3 + c] = g.edges[k].err[c];
  double plain = 0.0;
  for (int k : g.active) plain += chi2(g.edges[k]);  // SparseOptimizer::activeChi2, sparse_optimizer.cpp:90-98
  chi2_out[0] = plain;
  chi2_out[1] = activeRobustChi2(g);
  if (update) updateState(g, update);
  refba_get_poses(h, poses);
  refba_get_points(h, points);
}

// The normal equations of the active set chosen by the last refba_debug_phase, at the current state:
// buildStructure + computeActiveErrors + buildSystem (linearizeOplus + constructQuadraticForm of every active edge,
// base_binary_edge.hpp:55-120, block_solver.hpp:502-560).  Hpp: 36 per pose (row-major 6x6, zero for poses without an
// index), Hll: 9 per landmark, Hpl: 18 per EDGE (6x3 row-major: B^T W A of the (pose, landmark) block the edge adds to;
// zero when the pose is fixed or the edge inactive), b: 6 per pose then 3 per landmark (vertex order, not index order).
void refba_debug_system(refba* h, double* Hpp, double* Hll, double* Hpl, double* b) {
  Graph& g = h->g;
  buildStructure(g);
  computeActiveErrors(g);
  buildSystem(g);
  std::fill(Hpp, Hpp + (size_t)g.n_pose * 36, 0.0);
  std::fill(Hll, Hll + (size_t)g.n_point * 9, 0.0);
  std::fill(Hpl, Hpl + g.edges.size() * 18, 0.0);
  std::fill(b, b + (size_t)g.n_pose * 6 + (size_t)g.n_point * 3, 0.0);
  for (int i = 0; i < g.n_pose; i++) {
    const int s = g.pose_slot[i];
    if (s < 0) continue;
    std::memcpy(Hpp + (size_t)i * 36, &g.Hpp[(size_t)s * 36], 36 * sizeof(double));
    std::memcpy(b + (size_t)i * 6, &g.b[(size_t)s * 6], 6 * sizeof(double));
  }
  for (int j = 0; j < g.n_point; j++) {
    const int s = g.point_slot[j];
    if (s < 0) continue;
    std::memcpy(Hll + (size_t)j * 9, &g.Hll[(size_t)s * 9], 9 * sizeof(double));
    std::memcpy(b + (size_t)g.n_pose * 6 + (size_t)j * 3, &g.b[(size_t)g.Np * 6 + (size_t)s * 3], 3 * sizeof(double));
  }
  for (int k : g.active) {
    const Edge& e = g.edges[k];
    if (e.hpl >= 0 && g.pose_slot[e.pose] >= 0) std::memcpy(Hpl + (size_t)k * 18, &g.Hpl[(size_t)e.hpl * 18], 18 * sizeof(double));
  }
}

// Residuals / Jacobians / robust weights of every edge at the CURRENT state.
// err (n_obs,3), Jp (n_obs,18), Jl (n_obs,9), w (n_obs,) = rho1*invSigma2, rho0 (n_obs,) robustified chi2.
// huber: 0 = none, 1 = LBA deltas (sqrt(5.991)/sqrt(7.815)), 2 = GBA deltas (sqrt(5.99)/sqrt(7.815)).
void refba_linearize_all(refba* h, int huber, double* err, double* Jp, double* Jl, double* w, double* rho0) {
  Graph& g = h->g;
  const float t2 = huber == 2 ? std::sqrt(5.99) : std::sqrt(5.991);
  const float t3 = std::sqrt(7.815);
  for (size_t k = 0; k < g.edges.size(); k++) {
    Edge& e = g.edges[k];
    setHuber(e, t2, t3, huber != 0);
    computeError(g, e);
    linearize(g, e);
    std::memcpy(err + k * 3, e.err, 3 * sizeof(double));
    std::memcpy(Jp + k * 18, e.Jp, 18 * sizeof(double));
    std::memcpy(Jl + k * 9, e.Jl, 9 * sizeof(double));
    const double c = chi2(e);
    double rho[3] = {c, 1, 0};
    if (e.robust) robustify(e, c, rho);
    w[k] = rho[1] * e.info;
    rho0[k] = rho[0];
  }
}

// One damped Schur solve at the current state with all level-0 edges: returns the dense reduced matrix
// S (6Np x 6Np, symmetric), bschur (6Np), b (6Np+3Nl), x (6Np+3Nl), slot->pose, slot->point maps.
// Returns Np; the caller sizes buffers from n_pose/n_point upper bounds.
int refba_schur_solve(refba* h, int huber, double lambda, double* S_dense, double* bschur, double* b, double* x,
                      int32_t* slot_pose, int32_t* slot_point, double* max_diag) {
  Graph& g = h->g;
  const float t2 = huber == 2 ? std::sqrt(5.99) : std::sqrt(5.991);
  const float t3 = std::sqrt(7.815);
  for (Edge& e : g.edges) { e.level = 0; setHuber(e, t2, t3, huber != 0); }
  initializeOptimization(g, 0);
  buildStructure(g);
  computeActiveErrors(g);
  buildSystem(g);
  *max_diag = computeLambdaInit(g) / 1e-5;
  setLambda(g, lambda);
  const bool ok = blockSolve(g);
  restoreDiagonal(g);
  const int n = g.Np * 6;
  for (int r = 0; r < n; r++)
    for (int c = g.sky_first[r]; c <= r; c++) {
      const double v = g.S[g.sky_ptr[r] + (c - g.sky_first[r])];
      S_dense[(size_t)r * n + c] = v;
      S_dense[(size_t)c * n + r] = v;
    }
  std::memcpy(bschur, g.bschur.data(), n * sizeof(double));
  std::memcpy(b, g.b.data(), g.b.size() * sizeof(double));
  std::memcpy(x, g.x.data(), g.x.size() * sizeof(double));
  for (int i = 0; i < g.Np; i++) slot_pose[i] = g.slot_pose[i];
  for (int i = 0; i < g.Nl; i++) slot_point[i] = g.slot_point[i];
  return ok ? g.Np : -1;
}

// Apply oplus to one pose: out7 = exp(upd6) * in7   (VertexSE3Expmap::oplusImpl)
void refba_pose_oplus(const double* in7, const double* upd6, double* out7) {
  SE3 T;
  T.t[0] = in7[0]; T.t[1] = in7[1]; T.t[2] = in7[2];
  T.r = Quat{in7[3], in7[4], in7[5], in7[6]};
  SE3 r = se3mul(se3exp(upd6), T);
  out7[0] = r.t[0]; out7[1] = r.t[1]; out7[2] = r.t[2];
  out7[3] = r.r.x; out7[4] = r.r.y; out7[5] = r.r.z; out7[6] = r.r.w;
}

// SE3Quat::exp alone: out7 = (t, q)
void refba_se3_exp(const double* upd6, double* out7) {
  SE3 r = se3exp(upd6);
  out7[0] = r.t[0]; out7[1] = r.t[1]; out7[2] = r.t[2];
  out7[3] = r.r.x; out7[4] = r.r.y; out7[5] = r.r.z; out7[6] = r.r.w;
}

// cam_project restatements: Xc(3) -> mono (2) / stereo (3)
void refba_cam_project_mono(const double* Xc, double fx, double fy, double cx, double cy, double* out2) {
  out2[0] = Xc[0] / Xc[2] * fx + cx;
  out2[1] = Xc[1] / Xc[2] * fy + cy;
}
void refba_cam_project_stereo(const double* Xc, double fx, double fy, double cx, double cy, float bf, double* out3) {
  const float invz = (float)(1.0f / Xc[2]);
  out3[0] = Xc[0] * invz * fx + cx;
  out3[1] = Xc[1] * invz * fy + cy;
  out3[2] = out3[0] - (double)(bf * invz);
}
void refba_huber(double delta, double e, double* rho3) {
  Edge ed;
  ed.delta = delta; ed.dsqr = (double)(float)(delta * delta);
  robustify(ed, e, rho3);
}

// =============================================================================================
// Pose-only optimisation: g2oOptimizer::PoseOptimization, src/backend/g2oOptimizer.cc:385-559,
// 655-690 (the lidar block :560-640 is out of scope).  One VertexSE3Expmap, one unary EdgeSE3ProjectXYZOnlyPose per
// observation with ur<0 (the fork leaves the stereo branch empty, :481-483; an observation with ur>=0 is wired as the
// vendored EdgeStereoSE3ProjectXYZOnlyPose, types_six_dof_expmap.h:175-202, the way upstream ORB-SLAM2 does).
// BlockSolver_6_3 + LinearSolverDense (linear_solver_dense.h:65-113: Eigen LDLT of the 6x6 block, isPositive() or fail)
// + OptimizationAlgorithmLevenberg; 4 rounds x optimize(10) from the SAME initial pose, chi2 re-classification after
// each round (float chi2 against float thresholds), Huber kernels dropped from round 2 on.
// =============================================================================================
namespace {
struct PEdge {
  bool stereo;
  double obs[3], info, Xw[3];
  int level = 0;
  bool robust = true;
  double delta, dsqr;
  double err[3] = {0, 0, 0};
  double J[18];
};
struct PoseProblem {
  SE3 T, Tbak;
  double fx, fy, cx, cy, bf;
  std::vector<PEdge> e;
  std::vector<int> active;
  double H[36], b[6], x[6];
  double lambda = -1, ni = 2;
  int nBad = 0;
  std::vector<TraceRow> trace;
  int round = 0;
  std::vector<Graph::UEdge> u;  // lidar edges of the fifth optimisation (g2oOptimizer.cc:560-640), no robust kernel
};
// EdgeSE3ProjectXYZOnlyPose::computeError (types_six_dof_expmap.h:152-156, .cpp:290-296) /
// EdgeStereoSE3ProjectXYZOnlyPose::computeError (.h:184-188, .cpp:299-306: float invz, DOUBLE bf)
inline void pComputeError(const PoseProblem& P, PEdge& e) {
  double Xc[3];
  se3map(P.T, e.Xw, Xc);
  if (!e.stereo) {
    const double px = Xc[0] / Xc[2], py = Xc[1] / Xc[2];
    e.err[0] = e.obs[0] - (px * P.fx + P.cx);
    e.err[1] = e.obs[1] - (py * P.fy + P.cy);
    e.err[2] = 0;
  } else {
    const float invz = (float)(1.0f / Xc[2]);
    const double r0 = Xc[0] * invz * P.fx + P.cx, r1 = Xc[1] * invz * P.fy + P.cy;
    e.err[0] = e.obs[0] - r0;
    e.err[1] = e.obs[1] - r1;
    e.err[2] = e.obs[2] - (r0 - P.bf * invz);
  }
}
inline double pChi2(const PEdge& e) {
  if (!e.stereo) return e.err[0] * (e.info * e.err[0]) + e.err[1] * (e.info * e.err[1]);
  return e.err[0] * (e.info * e.err[0]) + e.err[1] * (e.info * e.err[1]) + e.err[2] * (e.info * e.err[2]);
}
// linearizeOplus (.cpp:266-288 mono, 335-364 stereo): invz = 1/z, invz_2 = invz*invz
inline void pLinearize(const PoseProblem& P, PEdge& e) {
  double Xc[3];
  se3map(P.T, e.Xw, Xc);
  const double x = Xc[0], y = Xc[1], invz = 1.0 / Xc[2], invz_2 = invz * invz, fx = P.fx, fy = P.fy;
  double* J = e.J;
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
  if (e.stereo) {
    J[12] = J[0] - P.bf * y * invz_2;
    J[13] = J[1] + P.bf * x * invz_2;
    J[14] = J[2];
    J[15] = J[3];
    J[16] = 0;
    J[17] = J[5] - P.bf * invz_2;
  } else {
    for (int c = 0; c < 6; c++) J[12 + c] = 0;
  }
}
inline void pRobustify(const PEdge& e, double c, double rho[3]) {
  if (c <= e.dsqr) { rho[0] = c; rho[1] = 1.; rho[2] = 0.; }
  else { const double sq = std::sqrt(c); rho[0] = 2 * sq * e.delta - e.dsqr; rho[1] = e.delta / sq; rho[2] = -0.5 * rho[1] / c; }
}
void pComputeActiveErrors(PoseProblem& P) {
  for (int k : P.active) pComputeError(P, P.e[k]);
  for (Graph::UEdge& e : P.u) e.err = uError(P.T, e);
}
double pActiveRobustChi2(const PoseProblem& P) {  // sparse_optimizer.cpp:100-114
  double chi = 0;
  for (int k : P.active) {
    const PEdge& e = P.e[k];
    const double c = pChi2(e);
    if (e.robust) { double rho[3]; pRobustify(e, c, rho); chi += rho[0]; } else chi += c;
  }
  for (const Graph::UEdge& e : P.u) chi += e.err * e.info * e.err;  // added after the visual edges -> last in _activeEdges
  return chi;
}
// BaseUnaryEdge::constructQuadraticForm (base_unary_edge.hpp:42-72) over the active edges
void pBuildSystem(PoseProblem& P) {
  for (double& v : P.H) v = 0;
  for (double& v : P.b) v = 0;
  for (int k : P.active) {
    PEdge& e = P.e[k];
    pLinearize(P, e);
    double w = e.info, rho1 = 1.0;
    if (e.robust) { double rho[3]; pRobustify(e, pChi2(e), rho); rho1 = rho[1]; w = rho[1] * e.info; }
    const int d = e.stereo ? 3 : 2;
    for (int i = 0; i < 6; i++) {
      double g = 0;
      for (int r = 0; r < d; r++) g += e.J[r * 6 + i] * (e.info * e.err[r]);
      P.b[i] -= rho1 * g;
      for (int j = 0; j < 6; j++) {
        double h = 0;
        for (int r = 0; r < d; r++) h += e.J[r * 6 + i] * (w * e.J[r * 6 + j]);
        P.H[i * 6 + j] += h;
      }
    }
  }
  for (Graph::UEdge& e : P.u) {  // BaseUnaryEdge numeric linearizeOplus + constructQuadraticForm (base_unary_edge.hpp:58-123)
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    for (int d = 0; d < 6; d++) {
      double add[6] = {0, 0, 0, 0, 0, 0};
      add[d] = delta;
      const double e1 = uError(se3mul(se3exp(add), P.T), e);
      add[d] = -delta;
      const double e2 = uError(se3mul(se3exp(add), P.T), e);
      e.J[d] = scalar * (e1 - e2);
    }
    for (int i = 0; i < 6; i++) {
      P.b[i] -= e.J[i] * e.info * e.err;
      for (int j = 0; j < 6; j++) P.H[i * 6 + j] += e.J[i] * e.info * e.J[j];
    }
  }
}
// Eigen::LDLT + isPositive() (linear_solver_dense.h:105-111), restated without pivoting
bool pSolve6(const double* H, const double* b, double* x) {
  double L[36] = {0}, D[6];
  for (int j = 0; j < 6; j++) {
    double d = H[j * 6 + j];
    for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
    if (!(d > 0.0)) return false;
    D[j] = d;
    for (int i = j + 1; i < 6; i++) {
      double v = H[i * 6 + j];
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
// OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-164) for the one-vertex graph
SolverResult pLmSolve(PoseProblem& P, int iteration) {
  pComputeActiveErrors(P);
  double currentChi = pActiveRobustChi2(P);
  double tempChi = currentChi;
  const double iniChi = currentChi;
  pBuildSystem(P);
  if (iteration == 0) {
    double md = 0;
    for (int j = 0; j < 6; j++) md = std::max(std::fabs(P.H[j * 6 + j]), md);
    P.lambda = 1e-5 * md;
    P.ni = 2;
    P.nBad = 0;
  }
  double rho = 0;
  int qmax = 0;
  do {
    P.Tbak = P.T;
    double Hd[36];
    std::memcpy(Hd, P.H, sizeof Hd);
    for (int j = 0; j < 6; j++) Hd[j * 6 + j] += P.lambda;
    const bool ok2 = pSolve6(Hd, P.b, P.x);
    if (!ok2) for (double& v : P.x) v = 0;  // BlockSolver::solve leaves x untouched on failure; the step is rejected below
    P.T = se3mul(se3exp(P.x), P.T);
    pComputeActiveErrors(P);
    tempChi = pActiveRobustChi2(P);
    if (!ok2) tempChi = std::numeric_limits<double>::max();
    rho = (currentChi - tempChi);
    double scale = 0.;
    for (int j = 0; j < 6; j++) scale += P.x[j] * (P.lambda * P.x[j] + P.b[j]);
    scale += 1e-3;
    rho /= scale;
    TraceRow tr{(double)P.round, (double)iteration, (double)qmax, P.lambda, currentChi, tempChi, rho, 0.0};
    if (rho > 0 && std::isfinite(tempChi)) {
      double alpha = 1. - std::pow((2 * rho - 1), 3);
      alpha = std::min(alpha, 2. / 3.);
      P.lambda *= std::max(1. / 3., alpha);
      P.ni = 2;
      currentChi = tempChi;
      tr.accepted = 1.0;
    } else {
      P.lambda *= P.ni;
      P.ni *= 2;
      P.T = P.Tbak;  // pop: edge errors keep the rejected trial's values
    }
    P.trace.push_back(tr);
    qmax++;
  } while (rho < 0 && qmax < 10);
  if (qmax == 10 || rho == 0) return Terminate;
  if ((iniChi - currentChi) * 1e3 < iniChi) P.nBad++; else P.nBad = 0;
  if (P.nBad >= 3) return Terminate;
  return OK;
}
}  // namespace

// pose7 in/out (SE3Quat::toVector order), cam = fx,fy,cx,cy,bf, xyz n x 3 (the reference copies float map points into
// double Xw, g2oOptimizer.cc:472-475), meas n x 4 float (u, v, ur<0 => mono, invSigma2).  outlier: n flags
// (Frame::mvbOutlier).  trace: up to max_trace rows of 8 doubles (round, iter, trial, lambda, chi_before, chi_trial,
// rho, accepted); *n_trace = rows written.  Returns nInitialCorrespondences - nBad (0 if fewer than 3 observations).
struct refba_frame_lidar {  // the clouds and lidarConfig fields the lidar block of PoseOptimization reads (:560-640)
  int32_t n_flat; const float* flat_xyz; const float* flat_normal;  // Frame::surface_points_flat_ / _normal_
  int32_t n_corner; const float* corner_xyz;                        // Frame::corner_points_sharp_
  int64_t n_map; const float* map_xyz;                              // local_lidarmap_cloud_ptr (world frame)
  double distance_sq_threshold, flat_weight, corner_weight;
  int32_t use_flat, use_corner;
};
static int poseOptRun(double* pose7, const double* cam, int n, const double* xyz, const float* meas, uint8_t* outlier,
                      double* trace, int max_trace, int* n_trace, const refba_frame_lidar* L, int32_t* n_match2) {
  PoseProblem P;
  P.T.t[0] = pose7[0]; P.T.t[1] = pose7[1]; P.T.t[2] = pose7[2];
  P.T.r = Quat{pose7[3], pose7[4], pose7[5], pose7[6]};
  normalizeRotation(P.T);
  const SE3 T0 = P.T;
  P.fx = cam[0]; P.fy = cam[1]; P.cx = cam[2]; P.cy = cam[3]; P.bf = cam[4];
  const float deltaMono = std::sqrt(5.991), deltaStereo = std::sqrt(7.815);  // :426-428
  P.e.resize(n);
  for (int i = 0; i < n; i++) {
    PEdge& e = P.e[i];
    e.stereo = !(meas[i * 4 + 2] < 0.0f);
    e.obs[0] = meas[i * 4 + 0]; e.obs[1] = meas[i * 4 + 1]; e.obs[2] = e.stereo ? meas[i * 4 + 2] : 0.0;
    e.info = meas[i * 4 + 3];
    for (int c = 0; c < 3; c++) e.Xw[c] = xyz[i * 3 + c];
    e.delta = e.stereo ? deltaStereo : deltaMono;
    e.dsqr = (double)(float)(e.delta * e.delta);  // `float dsqr`, robust_kernel_impl.h:84
    outlier[i] = 0;
  }
  if (n_trace) *n_trace = 0;
  if (n < 3) return 0;  // :491-492
  const float chi2Mono[4] = {5.991f, 5.991f, 5.991f, 5.991f}, chi2Stereo[4] = {7.815f, 7.815f, 7.815f, 7.815f};
  int nBad = 0;
  for (int it = 0; it < 4; it++) {
    P.T = T0;  // vSE3->setEstimate(Converter::toSE3Quat(pFrame->mTcw)), :510
    P.round = it;
    P.active.clear();
    for (int i = 0; i < n; i++) if (P.e[i].level == 0) P.active.push_back(i);
    if (!P.active.empty()) {  // optimize(10), sparse_optimizer.cpp:354-419
      bool ok = true;
      for (int i = 0; i < 10 && ok; i++) ok = (pLmSolve(P, i) == OK);
    }
    nBad = 0;
    for (int i = 0; i < n; i++) {  // :518-547
      PEdge& e = P.e[i];
      if (outlier[i]) pComputeError(P, e);
      const float chi2 = (float)pChi2(e);
      if (chi2 > (e.stereo ? chi2Stereo[it] : chi2Mono[it])) { outlier[i] = 1; e.level = 1; nBad++; }
      else { outlier[i] = 0; e.level = 0; }
      if (it == 2) e.robust = false;
    }
    if (n < 10) break;  // optimizer.edges().size()<10, :549-550
  }
  if (n_match2) n_match2[0] = n_match2[1] = 0;
  if (L && L->n_map > 100) {  // optimizer for lidar and visual, :560-640
    // pose_temp = Converter::toCvMat(vSE3->estimate()).inv(): the float pose, inverted (closed form, see twcFloat)
    float Rc[9], Oc[3];
    twcFloat(P.T, Rc, Oc);
    const std::vector<float> mapw(L->map_xyz, L->map_xyz + (size_t)L->n_map * 3);
    auto pass = [&](const float* cur, const float* nrm, int cnt, bool corner) {
      for (int i = 0; i < cnt; i++) {
        float qw[3], d2;
        toWorldFloat(Rc, Oc, cur + (size_t)i * 3, qw);
        const int j = nearestFloat(mapw, qw, &d2);
        if (!(d2 < L->distance_sq_threshold)) continue;  // float against the double threshold, :579 / :611
        Graph::UEdge e;
        e.pose = 0;
        e.corner = corner;
        for (int k = 0; k < 3; k++) {
          e.pc[k] = cur[(size_t)i * 3 + k];
          e.qw[k] = mapw[(size_t)j * 3 + k];
          e.n[k] = corner ? 0.0 : nrm[(size_t)i * 3 + k];
        }
        e.info = corner ? L->corner_weight : L->flat_weight;
        P.u.push_back(e);
        if (n_match2) n_match2[corner ? 1 : 0]++;
      }
    };
    if (L->use_flat) pass(L->flat_xyz, L->flat_normal, L->n_flat, false);
    if (L->use_corner) pass(L->corner_xyz, nullptr, L->n_corner, true);
    // vSE3->setEstimate(Converter::toSE3Quat(pFrame->mTcw)) after pFrame->SetPose(pose): the float round trip (:632)
    {
      double R[9];
      qtoR(P.T.r, R);
      for (double& v : R) v = (double)(float)v;
      for (double& v : P.T.t) v = (double)(float)v;
      P.T.r = RtoQ(R);
      normalizeRotation(P.T);
    }
    P.round = 4;
    P.active.clear();
    for (int i = 0; i < n; i++) if (P.e[i].level == 0) P.active.push_back(i);
    if (!P.active.empty() || !P.u.empty()) {
      bool ok = true;
      for (int i = 0; i < 10 && ok; i++) ok = (pLmSolve(P, i) == OK);
    }
  }
  nBad = 0;
  for (int i = 0; i < n; i++) {  // final classification, :656-680 (double 5.991 / upstream stereo 7.815)
    PEdge& e = P.e[i];
    if (outlier[i]) pComputeError(P, e);
    const float chi2 = (float)pChi2(e);
    if ((double)chi2 > (e.stereo ? 7.815 : 5.991)) { outlier[i] = 1; nBad++; } else outlier[i] = 0;
  }
  pose7[0] = P.T.t[0]; pose7[1] = P.T.t[1]; pose7[2] = P.T.t[2];
  pose7[3] = P.T.r.x; pose7[4] = P.T.r.y; pose7[5] = P.T.r.z; pose7[6] = P.T.r.w;
  const int nt = std::min<int>((int)P.trace.size(), max_trace);
  if (trace && nt > 0) std::memcpy(trace, P.trace.data(), (size_t)nt * sizeof(TraceRow));
  if (n_trace) *n_trace = nt;
  return n - nBad;
}
int refba_pose_opt(double* pose7, const double* cam, int n, const double* xyz, const float* meas, uint8_t* outlier,
                   double* trace, int max_trace, int* n_trace) {
  return poseOptRun(pose7, cam, n, xyz, meas, outlier, trace, max_trace, n_trace, nullptr, nullptr);
}
// the same with this fork's lidar block (g2oOptimizer.cc:560-640): when the local lidar map has more than 100 points,
// the frame's flat / sharp feature points -- moved to the world frame with the pose of the four visual rounds -- are
// matched to their nearest map point (k = 1, squared distance below the threshold), every match becomes an
// EdgeLidarFlatPoint / EdgeLidarCornerPoint on the pose (numeric Jacobians, information = weight, no kernel), and a
// fifth optimize(10) runs from the float-rounded pose over the level-0 visual edges and the lidar edges.
// n_match2: matched flat / corner points (the two counts the reference prints, :626-627).
int refba_pose_opt_lidar(double* pose7, const double* cam, int n, const double* xyz, const float* meas, uint8_t* outlier,
                         double* trace, int max_trace, int* n_trace, const refba_frame_lidar* lidar, int32_t* n_match2) {
  return poseOptRun(pose7, cam, n, xyz, meas, outlier, trace, max_trace, n_trace, lidar, n_match2);
}

// =====================================================================================================
// Essential-graph (Sim3 pose-graph) optimisation -- SURVEY.md 8(f) row N3, oracle only (no CUDA counterpart yet).
// Restates g2o's Sim3 (types/sim3.h:40-285), VertexSim3Expmap::oplusImpl and EdgeSim3::computeError
// (types/types_seven_dof_expmap.h:48-110), the NUMERIC Jacobians EdgeSim3 inherits (base_binary_edge.hpp:122-195:
// central differences, delta 1e-9, through the vertex's own oplus), constructQuadraticForm with identity information,
// and the optimisation set up by g2oOptimizer::OptimizeEssentialGraph (g2oOptimizer.cc:1212-1232, 1472-1478):
// Levenberg with setUserLambdaInit(1e-16), no marginalised vertices (BlockSolver_7_3 solves the whole system).
// Pinned against the reference binary by oracle/pin_libg2o_graph.py (make_sim3, make_posegraph).
// =====================================================================================================
namespace {
struct Sim3 { Quat r; double t[3]; double s; };

// Sim3(const Vector7d& update): exp map, update = (omega, upsilon, sigma)   (sim3.h:68-144)
inline Sim3 sim3exp(const double u[7]) {
  const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]}, sigma = u[6];
  const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double Om2[9];
  mat3mul(Om, Om, Om2);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  Sim3 S;
  S.s = std::exp(sigma);
  double R[9], A, B, C;
  const double eps = 0.00001;
  if (std::fabs(sigma) < eps) {
    C = 1;
    if (theta < eps) {
      A = 1. / 2.; B = 1. / 6.;
      for (int i = 0; i < 9; i++) R[i] = I[i] + Om[i] + Om2[i];
    } else {
      const double theta2 = theta * theta;
      A = (1 - std::cos(theta)) / theta2;
      B = (theta - std::sin(theta)) / (theta2 * theta);
      for (int i = 0; i < 9; i++) R[i] = I[i] + std::sin(theta) / theta * Om[i] + (1 - std::cos(theta)) / (theta * theta) * Om2[i];
    }
  } else {
    C = (S.s - 1) / sigma;
    if (theta < eps) {
      const double sigma2 = sigma * sigma;
      A = ((sigma - 1) * S.s + 1) / sigma2;
      B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
      for (int i = 0; i < 9; i++) R[i] = I[i] + Om[i] + Om2[i];
    } else {
      for (int i = 0; i < 9; i++) R[i] = I[i] + std::sin(theta) / theta * Om[i] + (1 - std::cos(theta)) / (theta * theta) * Om2[i];
      const double a = S.s * std::sin(theta), b = S.s * std::cos(theta);
      const double theta2 = theta * theta, sigma2 = sigma * sigma, c = theta2 + sigma2;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
    }
  }
  S.r = RtoQ(R);
  double W[9];
  for (int i = 0; i < 9; i++) W[i] = A * Om[i] + B * Om2[i] + C * I[i];
  for (int i = 0; i < 3; i++) S.t[i] = W[i * 3] * up[0] + W[i * 3 + 1] * up[1] + W[i * 3 + 2] * up[2];
  return S;
}

// Eigen's PartialPivLU solve of a 3x3 system (Matrix3d::lu().solve, sim3.h:224)
inline void lu3solve(const double Win[9], const double b[3], double x[3]) {
  double A[9];
  std::memcpy(A, Win, sizeof A);
  int perm[3] = {0, 1, 2};
  for (int k = 0; k < 3; k++) {
    int piv = k;
    for (int i = k + 1; i < 3; i++)
      if (std::fabs(A[i * 3 + k]) > std::fabs(A[piv * 3 + k])) piv = i;
    if (piv != k) {
      for (int j = 0; j < 3; j++) std::swap(A[k * 3 + j], A[piv * 3 + j]);
      std::swap(perm[k], perm[piv]);
    }
    for (int i = k + 1; i < 3; i++) {
      A[i * 3 + k] /= A[k * 3 + k];
      for (int j = k + 1; j < 3; j++) A[i * 3 + j] -= A[i * 3 + k] * A[k * 3 + j];
    }
  }
  double y[3];
  for (int i = 0; i < 3; i++) {
    y[i] = b[perm[i]];
    for (int j = 0; j < i; j++) y[i] -= A[i * 3 + j] * y[j];
  }
  for (int i = 2; i >= 0; i--) {
    x[i] = y[i];
    for (int j = i + 1; j < 3; j++) x[i] -= A[i * 3 + j] * x[j];
    x[i] /= A[i * 3 + i];
  }
}

// Sim3::log   (sim3.h:150-235)
inline void sim3log(const Sim3& S, double res[7]) {
  const double sigma = std::log(S.s);
  double R[9];
  qtoR(S.r, R);
  const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};  // deltaR, se3_ops.hpp:40-47
  const double eps = 0.00001;
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double om[3], A, B, C;
  if (std::fabs(sigma) < eps) {
    C = 1;
    if (d > 1 - eps) {
      for (int i = 0; i < 3; i++) om[i] = 0.5 * dR[i];
      A = 1. / 2.; B = 1. / 6.;
    } else {
      const double theta = std::acos(d), theta2 = theta * theta;
      for (int i = 0; i < 3; i++) om[i] = theta / (2 * std::sqrt(1 - d * d)) * dR[i];
      A = (1 - std::cos(theta)) / theta2;
      B = (theta - std::sin(theta)) / (theta2 * theta);
    }
  } else {
    C = (S.s - 1) / sigma;
    if (d > 1 - eps) {
      const double sigma2 = sigma * sigma;
      for (int i = 0; i < 3; i++) om[i] = 0.5 * dR[i];
      A = ((sigma - 1) * S.s + 1) / sigma2;
      B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
    } else {
      const double theta = std::acos(d);
      for (int i = 0; i < 3; i++) om[i] = theta / (2 * std::sqrt(1 - d * d)) * dR[i];
      const double theta2 = theta * theta;
      const double a = S.s * std::sin(theta), b = S.s * std::cos(theta), c = theta2 + sigma * sigma;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
    }
  }
  const double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double Om2[9], W[9], ups[3];
  mat3mul(Om, Om, Om2);
  for (int i = 0; i < 9; i++) W[i] = A * Om[i] + B * Om2[i] + C * I[i];
  lu3solve(W, S.t, ups);
  for (int i = 0; i < 3; i++) { res[i] = om[i]; res[i + 3] = ups[i]; }
  res[6] = sigma;
}

// Sim3::operator* and inverse (sim3.h:238-278).  No re-normalisation of the quaternion, unlike SE3Quat.
inline Sim3 sim3mul(const Sim3& a, const Sim3& b) {
  Sim3 r;
  r.r = qmul(a.r, b.r);
  double rt[3];
  qrot(a.r, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.s * rt[i] + a.t[i];
  r.s = a.s * b.s;
  return r;
}
inline Sim3 sim3inv(const Sim3& a) {
  Sim3 r;
  r.r = Quat{-a.r.x, -a.r.y, -a.r.z, a.r.w};
  const double v[3] = {(-1. / a.s) * a.t[0], (-1. / a.s) * a.t[1], (-1. / a.s) * a.t[2]};
  qrot(r.r, v, r.t);
  r.s = 1. / a.s;
  return r;
}
inline Sim3 sim3from8(const double* v) {  // (qx, qy, qz, qw, tx, ty, tz, s): Sim3::operator[] order
  Sim3 S;
  S.r = Quat{v[0], v[1], v[2], v[3]};
  S.t[0] = v[4]; S.t[1] = v[5]; S.t[2] = v[6];
  S.s = v[7];
  return S;
}
inline void sim3to8(const Sim3& S, double* v) {
  v[0] = S.r.x; v[1] = S.r.y; v[2] = S.r.z; v[3] = S.r.w;
  v[4] = S.t[0]; v[5] = S.t[1]; v[6] = S.t[2]; v[7] = S.s;
}
// VertexSim3Expmap::oplusImpl (types_seven_dof_expmap.h:63-72)
inline Sim3 sim3oplus(const Sim3& est, const double upd[7], bool fix_scale) {
  double u[7];
  std::memcpy(u, upd, sizeof u);
  if (fix_scale) u[6] = 0;
  return sim3mul(sim3exp(u), est);
}
// EdgeSim3::computeError (types_seven_dof_expmap.h:95-103): log(C * v1 * v2^-1)
inline void sim3edgeError(const Sim3& C, const Sim3& v1, const Sim3& v2, double e[7]) {
  sim3log(sim3mul(sim3mul(C, v1), sim3inv(v2)), e);
}

struct PoseGraph {
  int n = 0;
  std::vector<Sim3> v, bak;
  std::vector<uint8_t> fixed;
  bool fix_scale = true;
  struct E { int i, j; Sim3 meas; double err[7]; double Ji[49], Jj[49]; };
  std::vector<E> e;
  std::vector<int> slot;  // hessianIndex
  int N = 0;              // free vertices
  std::vector<double> H, b, x;
  std::vector<TraceRow> trace;
};

inline void pgErrors(PoseGraph& G) {
  for (auto& e : G.e) sim3edgeError(e.meas, G.v[e.i], G.v[e.j], e.err);
}
inline double pgChi2(const PoseGraph& G) {  // information = identity
  double c = 0;
  for (const auto& e : G.e) {
    double s = 0;
    for (int k = 0; k < 7; k++) s += e.err[k] * e.err[k];
    c += s;
  }
  return c;
}
// BaseBinaryEdge::linearizeOplus, numeric version (base_binary_edge.hpp:122-195)
inline void pgLinearize(PoseGraph& G, PoseGraph::E& e) {
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  for (int side = 0; side < 2; side++) {
    const int vi = side == 0 ? e.i : e.j;
    double* J = side == 0 ? e.Ji : e.Jj;
    if (G.fixed[vi]) continue;
    const Sim3 keep = G.v[vi];
    for (int d = 0; d < 7; d++) {
      double add[7] = {0, 0, 0, 0, 0, 0, 0}, e1[7], e2[7];
      add[d] = delta;
      G.v[vi] = sim3oplus(keep, add, G.fix_scale);
      sim3edgeError(e.meas, G.v[e.i], G.v[e.j], e1);
      add[d] = -delta;
      G.v[vi] = sim3oplus(keep, add, G.fix_scale);
      sim3edgeError(e.meas, G.v[e.i], G.v[e.j], e2);
      G.v[vi] = keep;
      for (int r = 0; r < 7; r++) J[r * 7 + d] = scalar * (e1[r] - e2[r]);
    }
  }
}
inline void pgBuildSystem(PoseGraph& G) {
  const int n = G.N * 7;
  std::fill(G.H.begin(), G.H.end(), 0.0);
  std::fill(G.b.begin(), G.b.end(), 0.0);
  for (auto& e : G.e) {
    pgLinearize(G, e);
    const int si = G.slot[e.i], sj = G.slot[e.j];
    // constructQuadraticForm without robust kernel, omega = I (base_binary_edge.hpp:97-117)
    if (si >= 0)
      for (int a = 0; a < 7; a++) {
        double s = 0;
        for (int r = 0; r < 7; r++) s += e.Ji[r * 7 + a] * e.err[r];
        G.b[si * 7 + a] -= s;
        for (int c = 0; c < 7; c++) {
          double h = 0;
          for (int r = 0; r < 7; r++) h += e.Ji[r * 7 + a] * e.Ji[r * 7 + c];
          G.H[(size_t)(si * 7 + a) * n + si * 7 + c] += h;
        }
      }
    if (sj >= 0)
      for (int a = 0; a < 7; a++) {
        double s = 0;
        for (int r = 0; r < 7; r++) s += e.Jj[r * 7 + a] * e.err[r];
        G.b[sj * 7 + a] -= s;
        for (int c = 0; c < 7; c++) {
          double h = 0;
          for (int r = 0; r < 7; r++) h += e.Jj[r * 7 + a] * e.Jj[r * 7 + c];
          G.H[(size_t)(sj * 7 + a) * n + sj * 7 + c] += h;
        }
      }
    if (si >= 0 && sj >= 0)
      for (int a = 0; a < 7; a++)
        for (int c = 0; c < 7; c++) {
          double h = 0;
          for (int r = 0; r < 7; r++) h += e.Ji[r * 7 + a] * e.Jj[r * 7 + c];
          G.H[(size_t)(si * 7 + a) * n + sj * 7 + c] += h;
          G.H[(size_t)(sj * 7 + c) * n + si * 7 + a] += h;
        }
  }
}
// dense LDL^T solve of (H + lambda I) x = b (stands for LinearSolverEigen's sparse Cholesky: an exact solve)
inline bool pgSolve(PoseGraph& G, double lambda) {
  const int n = G.N * 7;
  std::vector<double> A(G.H);
  for (int i = 0; i < n; i++) A[(size_t)i * n + i] += lambda;
  std::vector<double> D(n);
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k] * D[k];
    if (!(d > 0)) return false;
    D[j] = d;
    for (int i = j + 1; i < n; i++) {
      double l = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) l -= A[(size_t)i * n + k] * A[(size_t)j * n + k] * D[k];
      A[(size_t)i * n + j] = l / d;
    }
  }
  std::vector<double>& x = G.x;
  for (int i = 0; i < n; i++) {
    double v = G.b[i];
    for (int k = 0; k < i; k++) v -= A[(size_t)i * n + k] * x[k];
    x[i] = v;
  }
  for (int i = 0; i < n; i++) x[i] /= D[i];
  for (int i = n - 1; i >= 0; i--)
    for (int k = i + 1; k < n; k++) x[i] -= A[(size_t)k * n + i] * x[k];
  return true;
}
}  // namespace

// arithmetic probes (8-vectors are qx,qy,qz,qw,tx,ty,tz,s)
void refba_sim3_exp(const double* upd7, double* out8) { sim3to8(sim3exp(upd7), out8); }
void refba_sim3_log(const double* in8, double* out7) { sim3log(sim3from8(in8), out7); }
void refba_sim3_oplus(const double* in8, const double* upd7, int fix_scale, double* out8) {
  sim3to8(sim3oplus(sim3from8(in8), upd7, fix_scale != 0), out8);
}
void refba_sim3_edge_error(const double* meas8, const double* v1_8, const double* v2_8, double* err7) {
  sim3edgeError(sim3from8(meas8), sim3from8(v1_8), sim3from8(v2_8), err7);
}

// Sim3 pose-graph optimisation as OptimizeEssentialGraph sets it up: LM, lambda_0 = lambda_init (1e-16 in the reference),
// `iters` iterations, identity information, vertices in ascending id.  vertices n x 8 in/out; edges (i, j) with
// measurement S_ji (vertex 0 = i, vertex 1 = j).  trace rows as in refba_get_trace.  Returns the iterations performed.
int refba_pose_graph(int n, double* vert8, const uint8_t* fixed, int fix_scale, int n_e, const int32_t* edge_ij,
                     const double* meas8, int iters, double lambda_init, double* trace, int max_trace, int* n_trace) {
  PoseGraph G;
  G.n = n;
  G.fix_scale = fix_scale != 0;
  G.v.resize(n);
  G.fixed.assign(fixed, fixed + n);
  for (int i = 0; i < n; i++) G.v[i] = sim3from8(vert8 + (size_t)i * 8);
  G.e.resize(n_e);
  for (int k = 0; k < n_e; k++) {
    G.e[k].i = edge_ij[2 * k];
    G.e[k].j = edge_ij[2 * k + 1];
    G.e[k].meas = sim3from8(meas8 + (size_t)k * 8);
  }
  // index mapping: non-fixed vertices that have an edge, ascending id
  std::vector<uint8_t> act(n, 0);
  for (auto& e : G.e) { act[e.i] = 1; act[e.j] = 1; }
  G.slot.assign(n, -1);
  for (int i = 0; i < n; i++)
    if (act[i] && !G.fixed[i]) G.slot[i] = G.N++;
  const int dim = G.N * 7;
  G.H.assign((size_t)dim * dim, 0.0);
  G.b.assign(dim, 0.0);
  G.x.assign(dim, 0.0);
  double lambda = 0, ni = 2;
  int nBad = 0, done = 0;
  bool ok = true;
  for (int it = 0; it < iters && ok && dim > 0; it++) {  // SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve
    pgErrors(G);
    double currentChi = pgChi2(G), tempChi = currentChi;
    const double iniChi = currentChi;
    pgBuildSystem(G);
    if (it == 0) {
      if (lambda_init > 0) lambda = lambda_init;  // computeLambdaInit: _userLambdaInit (optimization_algorithm_levenberg.cpp:168-169)
      else {
        double md = 0;
        for (int i = 0; i < dim; i++) md = std::max(md, std::fabs(G.H[(size_t)i * dim + i]));
        lambda = 1e-5 * md;
      }
      ni = 2;
      nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      G.bak = G.v;
      const bool ok2 = pgSolve(G, lambda);
      for (int i = 0; i < n; i++)
        if (G.slot[i] >= 0) G.v[i] = sim3oplus(G.v[i], &G.x[(size_t)G.slot[i] * 7], G.fix_scale);
      pgErrors(G);
      tempChi = pgChi2(G);
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = currentChi - tempChi;
      double scale = 0;
      for (int j = 0; j < dim; j++) scale += G.x[j] * (lambda * G.x[j] + G.b[j]);
      scale += 1e-3;
      rho /= scale;
      TraceRow tr{0.0, (double)it, (double)qmax, lambda, currentChi, tempChi, rho, 0.0};
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        lambda *= std::max(1. / 3., alpha);
        ni = 2;
        currentChi = tempChi;
        tr.accepted = 1.0;
      } else {
        lambda *= ni;
        ni *= 2;
        G.v = G.bak;
      }
      G.trace.push_back(tr);
      qmax++;
    } while (rho < 0 && qmax < 10);
    done++;
    if (qmax == 10 || rho == 0) { ok = false; break; }
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
    if (nBad >= 3) ok = false;
  }
  for (int i = 0; i < n; i++) sim3to8(G.v[i], vert8 + (size_t)i * 8);
  const int nt = std::min<int>((int)G.trace.size(), max_trace);
  if (trace && nt > 0) std::memcpy(trace, G.trace.data(), (size_t)nt * sizeof(TraceRow));
  if (n_trace) *n_trace = nt;
  return done;
}

// =====================================================================================================
// Sim3 alignment of a loop-candidate keyframe pair: g2oOptimizer::OptimizeSim3,
// src/backend/g2oOptimizer.cc:1560-1796.  ONE free VertexSim3Expmap (S12, carrying both cameras' focal lengths and
// principal points), the matched map points as FIXED vertices in their own camera frames, and per match two binary
// edges with a Huber kernel of delta = (float)sqrt(th2):
//   EdgeSim3ProjectXYZ        e12 = obs1 - cam_map1(project(S12.map(P2c)))            (types_seven_dof_expmap.h:130-149)
//   EdgeInverseSim3ProjectXYZ e21 = obs2 - cam_map2(project(S12.inverse().map(P1c)))  (:152-171)
// Both edges leave linearizeOplus to BaseBinaryEdge (the analytic one is commented out, types_seven_dof_expmap.cpp), so
// the 2x7 Jacobian is the NUMERIC one (base_binary_edge.hpp:122-195, delta 1e-9, central differences through oplusImpl)
// and only the Sim3 side is linearised (the point vertex is fixed).  BlockSolverX + LinearSolverDense + Levenberg:
// optimize(5), drop the pairs with chi2 > th2 on either edge (stored errors, i.e. possibly those of a rejected last
// trial), give up (return 0, S12 untouched) with fewer than 10 pairs left, optimize(10 if something was dropped else 5),
// count the pairs that still pass.
namespace {
struct S3Pair { double P1c[3], P2c[3], obs1[2], obs2[2], info1, info2; double e12[2], e21[2]; bool on = true; };
struct Sim3Problem {
  Sim3 S, Sbak;
  bool fix_scale;
  double f1[2], c1[2], f2[2], c2[2];
  double delta, dsqr;
  std::vector<S3Pair> m;
  double H[49], b[7], x[7];
  double lambda = -1, ni = 2;
  int nBad = 0, round = 0;
  std::vector<TraceRow> trace;
};
inline void sim3map(const Sim3& S, const double* p, double* out) {  // Sim3::map, sim3.h:144-146
  double r[3];
  qrot(S.r, p, r);
  for (int i = 0; i < 3; i++) out[i] = S.s * r[i] + S.t[i];
}
inline void s3Error12(const Sim3Problem& P, const Sim3& S, const S3Pair& m, double e[2]) {
  double X[3];
  sim3map(S, m.P2c, X);
  e[0] = m.obs1[0] - (X[0] / X[2] * P.f1[0] + P.c1[0]);
  e[1] = m.obs1[1] - (X[1] / X[2] * P.f1[1] + P.c1[1]);
}
inline void s3Error21(const Sim3Problem& P, const Sim3& S, const S3Pair& m, double e[2]) {
  double X[3];
  sim3map(sim3inv(S), m.P1c, X);
  e[0] = m.obs2[0] - (X[0] / X[2] * P.f2[0] + P.c2[0]);
  e[1] = m.obs2[1] - (X[1] / X[2] * P.f2[1] + P.c2[1]);
}
inline void s3ComputeActiveErrors(Sim3Problem& P) {
  for (auto& m : P.m) if (m.on) { s3Error12(P, P.S, m, m.e12); s3Error21(P, P.S, m, m.e21); }
}
inline void s3Rho(const Sim3Problem& P, double c, double rho[2]) {
  if (c <= P.dsqr) { rho[0] = c; rho[1] = 1.; }
  else { const double sq = std::sqrt(c); rho[0] = 2 * sq * P.delta - P.dsqr; rho[1] = P.delta / sq; }
}
inline double s3Chi2(const double e[2], double info) { return e[0] * (info * e[0]) + e[1] * (info * e[1]); }
inline double s3ActiveRobustChi2(const Sim3Problem& P) {
  double chi = 0, rho[2];
  for (const auto& m : P.m) if (m.on) {
    s3Rho(P, s3Chi2(m.e12, m.info1), rho); chi += rho[0];
    s3Rho(P, s3Chi2(m.e21, m.info2), rho); chi += rho[0];
  }
  return chi;
}
inline void s3AddEdge(Sim3Problem& P, const double e[2], double info, const double J[14]) {
  double rho[2];
  s3Rho(P, s3Chi2(e, info), rho);
  const double w = rho[1] * info;
  for (int i = 0; i < 7; i++) {
    P.b[i] -= rho[1] * (J[i] * (info * e[0]) + J[7 + i] * (info * e[1]));
    for (int j = 0; j < 7; j++) P.H[i * 7 + j] += J[i] * (w * J[j]) + J[7 + i] * (w * J[7 + j]);
  }
}
inline void s3BuildSystem(Sim3Problem& P) {
  for (double& v : P.H) v = 0;
  for (double& v : P.b) v = 0;
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  for (auto& m : P.m) {
    if (!m.on) continue;
    double J12[14], J21[14];
    for (int d = 0; d < 7; d++) {
      double add[7] = {0, 0, 0, 0, 0, 0, 0}, a[2], c[2];
      add[d] = delta;
      const Sim3 Sp = sim3oplus(P.S, add, P.fix_scale);
      add[d] = -delta;
      const Sim3 Sm = sim3oplus(P.S, add, P.fix_scale);
      s3Error12(P, Sp, m, a); s3Error12(P, Sm, m, c);
      J12[d] = scalar * (a[0] - c[0]); J12[7 + d] = scalar * (a[1] - c[1]);
      s3Error21(P, Sp, m, a); s3Error21(P, Sm, m, c);
      J21[d] = scalar * (a[0] - c[0]); J21[7 + d] = scalar * (a[1] - c[1]);
    }
    s3AddEdge(P, m.e12, m.info1, J12);
    s3AddEdge(P, m.e21, m.info2, J21);
  }
}
// Eigen::LDLT + isPositive() (linear_solver_dense.h), restated without pivoting, 7x7
inline bool s3Solve7(const double* H, const double* b, double* x) {
  double L[49] = {0}, D[7], y[7];
  for (int j = 0; j < 7; j++) {
    double d = H[j * 7 + j];
    for (int k = 0; k < j; k++) d -= L[j * 7 + k] * L[j * 7 + k] * D[k];
    if (!(d > 0.0)) return false;
    D[j] = d;
    for (int i = j + 1; i < 7; i++) {
      double v = H[i * 7 + j];
      for (int k = 0; k < j; k++) v -= L[i * 7 + k] * L[j * 7 + k] * D[k];
      L[i * 7 + j] = v / d;
    }
  }
  for (int i = 0; i < 7; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[i * 7 + k] * y[k]; y[i] = v; }
  for (int i = 0; i < 7; i++) y[i] /= D[i];
  for (int i = 6; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < 7; k++) v -= L[k * 7 + i] * x[k]; x[i] = v; }
  return true;
}
SolverResult s3LmSolve(Sim3Problem& P, int iteration) {  // optimization_algorithm_levenberg.cpp:61-164
  s3ComputeActiveErrors(P);
  double currentChi = s3ActiveRobustChi2(P), tempChi = currentChi;
  const double iniChi = currentChi;
  s3BuildSystem(P);
  if (iteration == 0) {
    double md = 0;
    for (int j = 0; j < 7; j++) md = std::max(std::fabs(P.H[j * 7 + j]), md);
    P.lambda = 1e-5 * md;
    P.ni = 2;
    P.nBad = 0;
  }
  double rho = 0;
  int qmax = 0;
  do {
    P.Sbak = P.S;
    double Hd[49];
    std::memcpy(Hd, P.H, sizeof Hd);
    for (int j = 0; j < 7; j++) Hd[j * 7 + j] += P.lambda;
    const bool ok2 = s3Solve7(Hd, P.b, P.x);
    if (!ok2) for (double& v : P.x) v = 0;
    P.S = sim3oplus(P.S, P.x, P.fix_scale);
    s3ComputeActiveErrors(P);
    tempChi = s3ActiveRobustChi2(P);
    if (!ok2) tempChi = std::numeric_limits<double>::max();
    rho = (currentChi - tempChi);
    double scale = 0.;
    for (int j = 0; j < 7; j++) scale += P.x[j] * (P.lambda * P.x[j] + P.b[j]);
    scale += 1e-3;
    rho /= scale;
    TraceRow tr{(double)P.round, (double)iteration, (double)qmax, P.lambda, currentChi, tempChi, rho, 0.0};
    if (rho > 0 && std::isfinite(tempChi)) {
      double alpha = 1. - std::pow((2 * rho - 1), 3);
      alpha = std::min(alpha, 2. / 3.);
      P.lambda *= std::max(1. / 3., alpha);
      P.ni = 2;
      currentChi = tempChi;
      tr.accepted = 1.0;
    } else {
      P.lambda *= P.ni;
      P.ni *= 2;
      P.S = P.Sbak;
    }
    P.trace.push_back(tr);
    qmax++;
  } while (rho < 0 && qmax < 10);
  if (qmax == 10 || rho == 0) return Terminate;
  if ((iniChi - currentChi) * 1e3 < iniChi) P.nBad++; else P.nBad = 0;
  if (P.nBad >= 3) return Terminate;
  return OK;
}
}  // namespace

// s12 in/out (qx qy qz qw | t | s); cam8 = fx1 fy1 cx1 cy1 fx2 fy2 cx2 cy2; p1c / p2c: n x 3 (the matched map points
// in the frames of keyframe 1 / 2); meas6: n x 6 float (u1 v1 invSigma2_1 u2 v2 invSigma2_2).  keep[i] = 1 while
// vpMatches1 keeps the match.  Returns nIn (0 -- and s12 untouched -- when fewer than 10 pairs survive the first test).
int refba_optimize_sim3(double* s12, const double* cam8, int n, const double* p1c, const double* p2c, const float* meas6,
                        float th2, int fix_scale, uint8_t* keep, double* trace, int max_trace, int* n_trace) {
  Sim3Problem P;
  P.S = sim3from8(s12);
  P.fix_scale = fix_scale != 0;
  P.f1[0] = cam8[0]; P.f1[1] = cam8[1]; P.c1[0] = cam8[2]; P.c1[1] = cam8[3];
  P.f2[0] = cam8[4]; P.f2[1] = cam8[5]; P.c2[0] = cam8[6]; P.c2[1] = cam8[7];
  const float deltaHuber = std::sqrt(th2);  // :1619
  P.delta = deltaHuber;
  P.dsqr = (double)(float)(P.delta * P.delta);
  P.m.resize(n);
  for (int i = 0; i < n; i++) {
    S3Pair& m = P.m[i];
    for (int c = 0; c < 3; c++) { m.P1c[c] = p1c[i * 3 + c]; m.P2c[c] = p2c[i * 3 + c]; }
    m.obs1[0] = meas6[i * 6]; m.obs1[1] = meas6[i * 6 + 1]; m.info1 = meas6[i * 6 + 2];
    m.obs2[0] = meas6[i * 6 + 3]; m.obs2[1] = meas6[i * 6 + 4]; m.info2 = meas6[i * 6 + 5];
    m.e12[0] = m.e12[1] = m.e21[0] = m.e21[1] = 0;
    keep[i] = 1;
  }
  auto optimize = [&](int iters) {  // sparse_optimizer.cpp:354-419; no active edge -> nothing happens
    bool any = false;
    for (auto& m : P.m) any |= m.on;
    bool ok = any;
    for (int i = 0; i < iters && ok; i++) ok = (s3LmSolve(P, i) == OK);
  };
  auto finish = [&](int ret) {
    const int nt = std::min<int>((int)P.trace.size(), max_trace);
    if (trace && nt > 0) std::memcpy(trace, P.trace.data(), (size_t)nt * sizeof(TraceRow));
    if (n_trace) *n_trace = nt;
    return ret;
  };
  P.round = 0;
  optimize(5);
  int nBad = 0;
  for (int i = 0; i < n; i++) {
    S3Pair& m = P.m[i];
    if (s3Chi2(m.e12, m.info1) > th2 || s3Chi2(m.e21, m.info2) > th2) { keep[i] = 0; m.on = false; nBad++; }
  }
  const int more = nBad > 0 ? 10 : 5;
  if (n - nBad < 10) return finish(0);
  P.round = 1;
  optimize(more);
  int nIn = 0;
  for (int i = 0; i < n; i++) {
    S3Pair& m = P.m[i];
    if (!m.on) continue;
    if (s3Chi2(m.e12, m.info1) > th2 || s3Chi2(m.e21, m.info2) > th2) keep[i] = 0; else nIn++;
  }
  sim3to8(P.S, s12);
  return finish(nIn);
}

// errors and numeric Jacobians (2x7 row-major each) of one match's two edges at S12 -- for the pin against the binary
void refba_sim3_match_linearize(const double* s12, const double* cam8, const double* p1c, const double* p2c,
                                const float* meas6, int fix_scale, double* e12, double* e21, double* J12, double* J21) {
  Sim3Problem P;
  P.S = sim3from8(s12);
  P.fix_scale = fix_scale != 0;
  P.f1[0] = cam8[0]; P.f1[1] = cam8[1]; P.c1[0] = cam8[2]; P.c1[1] = cam8[3];
  P.f2[0] = cam8[4]; P.f2[1] = cam8[5]; P.c2[0] = cam8[6]; P.c2[1] = cam8[7];
  S3Pair m;
  for (int c = 0; c < 3; c++) { m.P1c[c] = p1c[c]; m.P2c[c] = p2c[c]; }
  m.obs1[0] = meas6[0]; m.obs1[1] = meas6[1]; m.info1 = meas6[2];
  m.obs2[0] = meas6[3]; m.obs2[1] = meas6[4]; m.info2 = meas6[5];
  s3Error12(P, P.S, m, e12);
  s3Error21(P, P.S, m, e21);
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  for (int d = 0; d < 7; d++) {
    double add[7] = {0, 0, 0, 0, 0, 0, 0}, a[2], c[2];
    add[d] = delta;
    const Sim3 Sp = sim3oplus(P.S, add, P.fix_scale);
    add[d] = -delta;
    const Sim3 Sm = sim3oplus(P.S, add, P.fix_scale);
    s3Error12(P, Sp, m, a); s3Error12(P, Sm, m, c);
    J12[d] = scalar * (a[0] - c[0]); J12[7 + d] = scalar * (a[1] - c[1]);
    s3Error21(P, Sp, m, a); s3Error21(P, Sm, m, c);
    J21[d] = scalar * (a[0] - c[0]); J21[7 + d] = scalar * (a[1] - c[1]);
  }
}

int refba_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"

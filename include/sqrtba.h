/* sqrtba.h -- C ABI of the B200-native square-root Levenberg-Marquardt bundle-adjustment back-end.
 *
 * Drop-in boundary for the bundle-adjustment hot path of lutao98/SqrtLM-SLAM.  Each entry point names the
 * reference interface it replaces (paths relative to the reference checkout):
 *
 *   sqrtba_set_problem      <- graph construction in g2oOptimizer::LocalBundleAdjustment / ::BundleAdjustment
 *                              (src/backend/g2oOptimizer.cc:805-919, 142-295): VertexSE3Expmap per keyframe,
 *                              VertexSBAPointXYZ per map point, Edge(Stereo)SE3ProjectXYZ per observation
 *   sqrtba_solve_local      <- the optimise part of Optimizer::LocalBundleAdjustment
 *                              (include/backend/Optimizer.h:55, src/backend/g2oOptimizer.cc:923-976[,1113-1114])
 *   sqrtba_solve_global     <- the optimise part of Optimizer::BundleAdjustment / GlobalBundleAdjustemnt
 *                              (include/backend/Optimizer.h:50-53, src/backend/g2oOptimizer.cc:300-301)
 *   sqrtba_get_poses/points <- vSE3->estimate() / vPoint->estimate() read-back (g2oOptimizer.cc:1167-1189, 308-361)
 *   sqrtba_get_outliers     <- the chi2 / depth test that fills vToErase (g2oOptimizer.cc:1119-1142)
 *   stop_flag               <- bool* pbStopFlag / optimizer.setForceStopFlag (g2oOptimizer.cc:797-798, 923-947)
 *
 * Conventions: plain C, host pointers unless a name says `_device`, caller keeps ownership of every buffer
 * (inputs are copied), all functions return 0 on success and a negative sqrtba_status on failure and never
 * throw; sqrtba_last_error() gives the message.  A handle is owned by one calling thread at a time; different
 * handles may be used concurrently (the reference runs local and global BA on different threads).
 * There is no CPU fallback: every solve runs on the CUDA device or fails with SQRTBA_ERR_CUDA.
 */
#ifndef SQRTBA_H_
#define SQRTBA_H_

#include <stdint.h>
#ifndef __cplusplus
#include <stdbool.h> /* the stop flags are the reference's `bool* pbStopFlag` */
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sqrtba_handle sqrtba_handle;

typedef enum {
  SQRTBA_OK = 0,
  SQRTBA_ERR_INVALID = -1, /* bad argument / problem not set */
  SQRTBA_ERR_CUDA = -2,    /* CUDA runtime error or no device */
  SQRTBA_ERR_ALLOC = -3,
  SQRTBA_ERR_COMM = -4     /* NCCL / multi-GPU set-up error */
} sqrtba_status;

typedef struct {
  int32_t device;            /* CUDA device ordinal */
  double pcg_rtol;           /* PCG stops when sqrt(r'M^-1 r / r0'M^-1 r0) <= pcg_rtol.  0 (default) = automatic: 1e-7 for
                                two-pass local BA on windows of up to 128 free keyframes (single or batched), 1e-8 for
                                global BA with its two-level preconditioner (1e-9 with the 6x6 blocks), 1e-9 for handles
                                with a third pass (third_pass_iters > 0: the fork's schedule, lidar edges with
                                central-difference Jacobians); the value in effect is stats.reserved[4].  The result tolerances (identical trial sequence and
                                outlier flags, cost 1e-6, pose RMS 1e-5 m) hold unchanged up to 1e-6 on local windows */
  int32_t pcg_max_iters;     /* hard cap per linear solve (default 2000; a safety net -- global BA needs a few hundred) */
  int32_t third_pass_iters;  /* 0 = ORB-SLAM2 two-pass 5+10 (default); 20 = this fork's extra pass (g2oOptimizer.cc:1113) */
  int32_t pcg_mode;          /* 0 = auto: single-window problems run the whole PCG solve in ONE persistent cooperative
                                kernel (grid barriers between the matvec and the vector updates, in-kernel NVLink
                                exchange when landmark-sharded), batches use one launch per CG phase;
                                1 = always one launch per phase; 2 = persistent whenever possible;
                                3 = like 0 but without the CUDA graph of the LM macro step (A/B: single-window problems
                                otherwise run every LM trial as ONE graph launch and read its verdict from mapped
                                pinned memory instead of synchronising the stream);
                                4 = reproducible: every pose-side sum leaves the CTAs as per-(CTA, window) partial vectors
                                that are added in a fixed order -- no floating-point atomics anywhere, bit-identical
                                results from run to run; one launch per CG phase (a single window pays for it: the
                                persistent PCG kernel is what makes C0 fast).  Needs <= 128 free poses per window, no
                                landmark with more than 32 observations, one GPU; otherwise the default path runs and
                                stats.reserved[5] says 0;
                                5 = like 0, but a big single window (global BA; any single window of 64 or more free
                                keyframes takes that path by default, SQRTBA_BIG_MIN_SLOTS) keeps the 6x6
                                block-Jacobi preconditioner instead of the 20-pose chunk blocks (A/B; stats.reserved[6]);
                                6 = like 0, but the big-window preconditioner keeps its chunk level only, without the
                                coarse correction over the chunks (A/B; stats.reserved[7]) */
  int32_t pcg_check_every;   /* multi-launch mode: host polls the convergence counter every N iterations */
  int32_t reserved[8];       /* tuning / A-B knobs, all 0 by default:
                                [0] record per-stage CUDA-event times (stats.ms_linearize ...)
                                [1] 1 = general (non-TMA) matvec kernel instead of the pipelined one
                                [2] force the depth of the matvec's shared-memory ring (2 or 3 stages)
                                [3] host threads of sqrtba_set_problem's preprocessing (0 = all cores)
                                [4] 1 = keep the caller's landmark order in big windows (no internal re-ordering)
                                [5] big-window matvec: pose slots of a shared-memory accumulator window (0 = global atomics)
                                [6] 1 = landmark-sharded global BA without peer-mapped buffers (NCCL all-reduce per iteration)
                                [7] linearise / landmark-QR kernel variant: 0 = pipelined v2 (default), 1 = one tile per
                                    CTA, 2 = first pipelined version, 5 = v2 QR compiled for 5 CTAs/SM, 6 = linearisation with
                                    L1 prefetch of the next tile's pose / landmark lines (experimental), 7 = fused
                                    linearise + landmark QR kernel for the iterations whose lambda is known and the retries
                                    (parity-tested; measured slower than the two separate kernels) */
} sqrtba_config;

/* one row per LM trial (g2o "levenbergIterations"), per window */
typedef struct {
  double pass;       /* 0,1,(2) */
  double iter;       /* outer iteration index inside the pass */
  double trial;      /* qmax before increment */
  double lambda;     /* damping used for this trial */
  double chi_before; /* currentChi */
  double chi_trial;  /* tempChi */
  double rho;        /* gain ratio */
  double accepted;   /* 1/0 */
  double cg_iters;   /* PCG iterations spent on this trial */
  double cg_relres;  /* final sqrt(rz/rz0) */
} sqrtba_trace_row;

typedef struct {
  int32_t n_windows;
  int32_t lm_trials;        /* total trials (max over windows = lock-step macro steps) */
  int32_t cg_iters_total;   /* matvec launches */
  int32_t kernel_launches;  /* launches of this library's kernels during the last solve */
  double ms_total;          /* device time of the last solve (CUDA events on the handle's stream) */
  double ms_linearize, ms_qr, ms_pcg, ms_backsub, ms_cost; /* per-stage device time (events) */
  double ms_matvec;         /* sum of matvec kernel time (multi-launch mode; 0 in persistent mode) */
  double reserved[8];       /* [0] 1 = the persistent PCG kernel ran, [1] 1 = in-kernel NVLink exchange was active,
                               [2] grid of the persistent kernel, [4] PCG tolerance in effect,
                               [5] 1 = reproducible mode (pcg_mode 4) was active,
                               [6] 1 = global BA ran PCG with the chunk preconditioner (one 120x120 block of the reduced
                               system per 20 consecutive keyframes; the persistent kernel's default for big windows),
                               [7] 1 = ... plus the additive coarse correction Z (Z^T S Z)^-1 Z^T over the chunks */
} sqrtba_stats;

int sqrtba_default_config(sqrtba_config* cfg);
int sqrtba_create(const sqrtba_config* cfg, sqrtba_handle** out);
int sqrtba_destroy(sqrtba_handle* h);
const char* sqrtba_last_error(const sqrtba_handle* h);
const char* sqrtba_version(void);

/* One problem = one BA graph.
 *  pose_qt    n_pose x 7  tx,ty,tz,qx,qy,qz,qw  (SE3Quat::toVector, Thirdparty/g2o/g2o/types/se3quat.h:138-148), Tcw
 *  pose_fixed n_pose      1 = vSE3->setFixed(true) (g2oOptimizer.cc:154, 813, 829)
 *  cam        n_pose x 5  fx,fy,cx,cy,bf of the observing keyframe (KeyFrame.h:394)
 *  point_xyz  n_point x 3 world position
 *  obs_pose / obs_point   indices; observations MUST be grouped by landmark (non-decreasing obs_point), which is
 *                         the order the reference adapter creates edges in (g2oOptimizer.cc:856-919)
 *  obs_meas   n_obs x 4   float u, v, ur (ur < 0 => monocular EdgeSE3ProjectXYZ), invSigma2 */
int sqrtba_set_problem(sqrtba_handle* h, int32_t n_pose, int32_t n_point, int32_t n_obs, const double* pose_qt,
                       const uint8_t* pose_fixed, const double* cam, const double* point_xyz,
                       const int32_t* obs_pose, const int32_t* obs_point, const float* obs_meas);

/* Batch of independent windows (block-diagonal problem): same arrays concatenated, indices global, plus CSR
 * offsets per window (n_win+1 entries each).  Every window runs its own LM (own lambda, accept/reject, stop). */
int sqrtba_set_problem_batch(sqrtba_handle* h, int32_t n_win, const int64_t* win_pose_ptr,
                             const int64_t* win_point_ptr, const int64_t* win_obs_ptr, const double* pose_qt,
                             const uint8_t* pose_fixed, const double* cam, const double* point_xyz,
                             const int32_t* obs_pose, const int32_t* obs_point, const float* obs_meas);

/* Restore the initial estimate given to set_problem (device-to-device), so a solve can be repeated. */
int sqrtba_reset_state(sqrtba_handle* h);

/* Local BA: robust pass (5 its) -> chi2/depth outlier exclusion -> non-robust pass (10 its) [-> third pass].
 * stop_flag may be NULL.  The host polls it before every LM iteration (g2o: `for (i < iterations && !terminate())`)
 * and mirrors it into mapped pinned memory, where the LM decision kernel reads it at the end of every trial (g2o's
 * do-while condition); the caller's bool itself is never read by the device.  With a communicator
 * (sqrtba_comm_init) the solve calls are collective and every poll max-reduces the ranks' flags over NCCL, so all
 * ranks stop in the same LM step even if their callers raise the flag at different moments. */
int sqrtba_solve_local(sqrtba_handle* h, const volatile bool* stop_flag, sqrtba_stats* stats);
/* Global BA: one pass of `iters` LM iterations, Huber iff robust (deltas sqrt(5.99)/sqrt(7.815)), no outlier step. */
int sqrtba_solve_global(sqrtba_handle* h, int32_t iters, int32_t robust, const volatile bool* stop_flag,
                        sqrtba_stats* stats);

int sqrtba_get_poses(sqrtba_handle* h, double* pose_qt_out);   /* n_pose x 7 */
int sqrtba_get_points(sqrtba_handle* h, double* point_xyz_out); /* n_point x 3 */
int sqrtba_get_outliers(sqrtba_handle* h, uint8_t* flags_out);  /* n_obs, 1 = erase this observation */
int sqrtba_get_trace_len(sqrtba_handle* h, int32_t window);
int sqrtba_get_trace(sqrtba_handle* h, int32_t window, sqrtba_trace_row* rows_out, int32_t max_rows);

/* ---- multi-GPU: landmark-sharded bundle adjustment (global BA, BASELINE config 3) ---------------------------------
 * One process (one handle) per GPU.  Every rank passes ALL poses (replicated, identical) and ITS OWN shard of the map
 * points with all their observations to sqrtba_set_problem; the solve calls are collective.  Per CG iteration the
 * ranks all-reduce the pose-sized matvec result (NCCL over NVLink); per LM trial a few pose-sized vectors and three
 * scalars per window.  All ranks take identical LM decisions and end with identical poses; each holds its own points.
 * sqrtba_comm_unique_id: rank 0 creates the NCCL id (128 bytes) and distributes it by any means (MPI, torch.distributed,
 * a file); sqrtba_comm_init joins the communicator on the handle's device.  Without a communicator every handle is an
 * independent single-GPU solver (batched windows shard by window with no collective at all). */
int sqrtba_comm_unique_id(uint8_t* id128_out);
int sqrtba_comm_init(sqrtba_handle* h, int32_t nranks, int32_t rank, const uint8_t* id128);
int sqrtba_comm_destroy(sqrtba_handle* h);

/* ---- pose-only optimisation (tracking thread) --------------------------------------------------------------------
 * Replaces Optimizer::PoseOptimization (include/backend/Optimizer.h:58-60 -> src/backend/g2oOptimizer.cc:385-559,
 * 655-690; the lidar block :560-640 is sqrtba_pose_opt_lidar below): one SE3 vertex per frame, one unary reprojection edge per matched
 * map point (EdgeSE3ProjectXYZOnlyPose when ur < 0, EdgeStereoSE3ProjectXYZOnlyPose otherwise -- the reference fork
 * only creates the monocular ones, :441-483), Huber deltas sqrt(5.991)/sqrt(7.815), four rounds of optimize(10) from
 * the same initial pose with chi2 re-classification in between, kernels dropped from the third classification on.
 * A batch of frames (relocalisation candidates) is one launch: frame f owns observations
 * [frame_obs_ptr[f], frame_obs_ptr[f+1]).  Independent of sqrtba_set_problem (the handle only lends stream/buffers).
 *   pose_qt     n_frames x 7 in/out (Tcw, SE3Quat::toVector order)     cam  n_frames x 5 (fx,fy,cx,cy,bf)
 *   obs_xyz     n_obs x 3 world positions of the matched map points    obs_meas n_obs x 4 float u,v,ur,invSigma2
 *   outlier_out n_obs (Frame::mvbOutlier)    inliers_out n_frames (the function's return value; 0 if < 3 observations)
 * sqrtba_pose_opt_trace: LM trials of one frame of the LAST call, rows of 8 doubles
 *   (round, iteration, trial, lambda, chi2 before, chi2 trial, rho, accepted); returns the number of rows. */
int sqrtba_pose_opt(sqrtba_handle* h, int32_t n_frames, const int64_t* frame_obs_ptr, double* pose_qt, const double* cam,
                    const double* obs_xyz, const float* obs_meas, uint8_t* outlier_out, int32_t* inliers_out,
                    sqrtba_stats* stats);
int sqrtba_pose_opt_trace(sqrtba_handle* h, int32_t frame, double* rows_out, int32_t max_rows);

/* The same with the lidar block this fork adds to PoseOptimization (src/backend/g2oOptimizer.cc:560-640, facade
 * include/backend/Optimizer.h:58-59: PoseOptimization(pFrame, local_lidarmap_cloud_ptr, kdtree_local_map, lidarconfig)):
 * when the local lidar map holds more than 100 points, the frame's flat / sharp feature points (Frame.h:273-277) are
 * moved to the world frame with the pose of the four visual rounds, matched to their nearest map point (k = 1; exact
 * search on the device instead of the caller's kd-tree) and kept below distance_sq_threshold; every match becomes an
 * EdgeLidarFlatPoint / EdgeLidarCornerPoint on the pose (types_six_dof_expmap.h:206-262, numeric Jacobians, information
 * = weight, no kernel); the estimate restarts from the float-rounded pose (:557, :632) for a fifth optimize(10) over the
 * level-0 visual edges and the lidar edges, then the final classification.  One frame per call.  With n_map <= 100 or
 * fewer than 3 observations the call is sqrtba_pose_opt.  n_match2_out (may be NULL): matched flat / corner points, the
 * two counts the reference prints (:626-627).  sqrtba_pose_opt_trace(h, 0, ...) then holds the trials of all five
 * rounds (round index 4 = the lidar round). */
typedef struct {
  int32_t n_flat;  const float* flat_xyz;  const float* flat_normal; /* Frame::surface_points_flat_ / _flat_normal_, n x 3 */
  int32_t n_corner; const float* corner_xyz;                          /* Frame::corner_points_sharp_ */
  int64_t n_map;   const float* map_xyz;                              /* local_lidarmap_cloud_ptr, world frame, n x 3 */
  double distance_sq_threshold, flat_weight, corner_weight;           /* lidarConfig (cfg/lidar_slam.yaml:52-61) */
  int32_t use_flat, use_corner;                                       /* lidarConfig::using_flat_point / using_sharp_point */
} sqrtba_frame_lidar;
int sqrtba_pose_opt_lidar(sqrtba_handle* h, double* pose_qt, const double* cam, int32_t n_obs, const double* obs_xyz,
                          const float* obs_meas, uint8_t* outlier_out, int32_t* inliers_out, const sqrtba_frame_lidar* lidar,
                          int32_t* n_match2_out, sqrtba_stats* stats);

/* ---- lidar tight-coupling pass of this fork's LocalBundleAdjustment (src/backend/g2oOptimizer.cc:979-1117) ----
 * After the two visual passes the reference (1) builds a local lidar map from the flat / corner feature clouds of every
 * OTHER local keyframe, moved to the world frame with the freshly optimised poses (:985-1013), (2) matches every
 * feature point of the current keyframe to its nearest map point (pcl::KdTreeFLANN, k = 1) and keeps matches with a
 * squared distance below lidarConfig::distance_sq_threshold (:1043-1049, :1081-1087), (3) adds one unary edge per match
 * to the current keyframe's pose -- EdgeLidarFlatPoint (point-to-plane) / EdgeLidarCornerPoint (point-to-point),
 * types_six_dof_expmap.h:206-262, information = flat/corner_optimized_weight, no robust kernel -- and (4) runs
 * optimize(20) over visual + lidar edges (:1113-1114).  Here all four steps run on the device inside
 * sqrtba_solve_local, between pass 2 and the final outlier test, when cfg.third_pass_iters > 0 (20 in the reference).
 *
 * sqrtba_set_lidar        hands over the clouds (copied; valid until the next sqrtba_set_problem).  Single-window,
 *                         single-GPU handles only.  Points are float32 as in pcl::PointXYZI:
 *     flat_xyz/flat_normal/corner_xyz  the current keyframe's surface_points_less_flat_(+_normal_) /
 *                         corner_points_less_sharp_ in ITS OWN frame (KeyFrame.h:438-442)
 *     map_*_xyz/map_*_pose the same clouds of the other local keyframes, each point in its keyframe's frame, with the
 *                         pose index of that keyframe
 *     numeric_jacobian    1: central differences with delta = 1e-9 exactly as BaseUnaryEdge::linearizeOplus
 *                         (base_unary_edge.hpp:82-123), the reference's behaviour; 0: closed-form Jacobians
 * sqrtba_set_lidar_edges  explicit correspondences instead (callers with their own association, tests): n_flat flat edges
 *                         followed by n_corner corner edges; point_cam/point_world/normal are n x 3, weight n (0 = no edge)
 * sqrtba_get_lidar_matches after a solve: per current feature point (flat then corner) the matched map index or -1;
 *                         returns the number of entries.   sqrtba_num_lidar_edges: edges that took part. */
typedef struct sqrtba_lidar {
  int32_t cur_pose;
  int32_t n_flat;
  const float* flat_xyz;
  const float* flat_normal;
  int32_t n_corner;
  int32_t numeric_jacobian;
  const float* corner_xyz;
  int64_t n_map_flat;
  const float* map_flat_xyz;
  const int32_t* map_flat_pose;
  int64_t n_map_corner;
  const float* map_corner_xyz;
  const int32_t* map_corner_pose;
  double distance_sq_threshold; /* lidarConfig::distance_sq_threshold (cfg/lidar_slam.yaml:52) */
  double flat_weight;           /* lidarConfig::flat_optimized_weight */
  double corner_weight;         /* lidarConfig::corner_optimized_weight */
  int32_t use_flat;             /* lidarConfig::using_flat_point */
  int32_t use_corner;           /* lidarConfig::using_sharp_point */
} sqrtba_lidar;
int sqrtba_set_lidar(sqrtba_handle* h, const sqrtba_lidar* clouds);
int sqrtba_set_lidar_edges(sqrtba_handle* h, int32_t cur_pose, int32_t n_flat, int32_t n_corner, const double* point_cam,
                           const double* point_world, const double* normal, const double* weight, int32_t numeric_jacobian);
int sqrtba_get_lidar_matches(sqrtba_handle* h, int32_t* match_out);
int sqrtba_num_lidar_edges(sqrtba_handle* h);

/* ---- essential-graph (Sim3 pose-graph) optimisation: the loop-closing step before global BA ------
 * Replaces the optimiser of g2oOptimizer::OptimizeEssentialGraph (include/backend/Optimizer.h:62-67,
 * src/backend/g2oOptimizer.cc:1212-1460): VertexSim3Expmap vertices, EdgeSim3 edges (identity information, the numeric
 * Jacobians that edge type inherits), Levenberg with setUserLambdaInit(lambda_init) and `iters` iterations (reference:
 * 1e-16 and 20), exact sparse Cholesky per trial.  The adapter builds the graph from the map (spanning tree, loop
 * edges, covisibility >= 100, new loop connections) and applies the result (g2oOptimizer.cc:1462-1520).
 *   vert8    n_vert x 8, in/out: qx qy qz qw | tx ty tz | s  (g2o::Sim3 operator[] order), S_iw of every keyframe
 *   fixed    n_vert: 1 = setFixed (the loop keyframe)     fix_scale: VertexSim3Expmap::_fix_scale (stereo / RGB-D)
 *   edge_ij  n_edge x 2: vertex 0 = i, vertex 1 = j;  meas8 n_edge x 8: S_ji
 * Vertices without an edge and fixed vertices are not moved.  Independent of sqrtba_set_problem.
 * sqrtba_pose_graph_trace: LM trials of the LAST call, rows of 8 doubles (pass, iteration, trial, lambda, chi2 before,
 * chi2 of the trial, rho, accepted); rows_out == NULL returns the number of rows. */
int sqrtba_pose_graph(sqrtba_handle* h, int32_t n_vert, double* vert8, const uint8_t* fixed, int32_t fix_scale, int32_t n_edge,
                      const int32_t* edge_ij, const double* meas8, int32_t iters, double lambda_init, sqrtba_stats* stats);
int sqrtba_pose_graph_trace(sqrtba_handle* h, double* rows_out, int32_t max_rows);

/* ---- Sim3 alignment of a loop candidate: the optimisation inside LoopClosing::ComputeSim3 -------
 * Replaces Optimizer::OptimizeSim3 (include/backend/Optimizer.h:68-69 -> src/backend/g2oOptimizer.cc:1560-1796; called from
 * src/backend/LoopClosing.cc:513 with th2 = 10): ONE free
 * VertexSim3Expmap S12 carrying both cameras' intrinsics; the matched map points enter as fixed vertices in their own
 * camera frames; per match EdgeSim3ProjectXYZ (point of keyframe 2 into image 1) and EdgeInverseSim3ProjectXYZ (point of
 * keyframe 1 into image 2), information invSigma2 * I, Huber delta = (float)sqrt(th2), numeric Jacobians (1e-9 central
 * differences) as inherited by both edge types; BlockSolverX + LinearSolverDense + Levenberg: optimize(5), drop every
 * match with chi2 > th2 on either edge, give up with fewer than 10 left, optimize(10 if something was dropped else 5),
 * count the matches that still pass.  A batch of candidate pairs is one launch: pair k owns matches
 * [pair_match_ptr[k], pair_match_ptr[k+1]).  Independent of sqrtba_set_problem.
 *   sim3_12    n_pairs x 8 in/out (qx qy qz qw | t | s); untouched for a pair that gives up
 *   cam8       n_pairs x 8: fx1 fy1 cx1 cy1 fx2 fy2 cx2 cy2
 *   p1c, p2c   n_match x 3: the matched map points, R1w * X1 + t1w and R2w * X2 + t2w (:1650-1662)
 *   match_meas n_match x 6 float: u1 v1 invSigma2_1 u2 v2 invSigma2_2 (undistorted keypoints, :1680-1712)
 *   keep_out   n_match: 1 while vpMatches1 keeps the match, 0 = set to NULL by the reference
 *   inliers_out n_pairs: the function's return value nIn (0 when it gives up)
 * sqrtba_optimize_sim3_trace: LM trials of one pair of the LAST call, rows of 8 doubles (pass, iteration, trial, lambda,
 *   chi2 before, chi2 of the trial, rho, accepted); returns the number of rows. */
int sqrtba_optimize_sim3(sqrtba_handle* h, int32_t n_pairs, const int64_t* pair_match_ptr, double* sim3_12, const double* cam8,
                         const double* p1c, const double* p2c, const float* match_meas, float th2, int32_t fix_scale,
                         uint8_t* keep_out, int32_t* inliers_out, sqrtba_stats* stats);
int sqrtba_optimize_sim3_trace(sqrtba_handle* h, int32_t pair, double* rows_out, int32_t max_rows);

/* ---- stage-level entry points (kernel parity tests, profiling) --------------------------------------------
 * sqrtba_debug_linearize : run the fused residual+Jacobian+Huber kernel at the current state.
 *    huber: 0 none, 1 local-BA deltas, 2 global-BA deltas.  Outputs (any may be NULL):
 *    err n_obs x 3 (unweighted), Jp n_obs x 18, Jl n_obs x 9 (both scaled by sqrt(rho1*invSigma2)),
 *    r n_obs x 3 (weighted residual), chi2 per window (robustified sum).
 * sqrtba_debug_step      : one damped square-root step with the given lambda at the current linearisation:
 *    landmark QR, block-Jacobi, PCG, back-substitution.  Outputs: dp 6 per free pose slot (pose order),
 *    dl 3 per landmark, bs reduced rhs (6 per slot), cg_iters. State is NOT updated.
 * sqrtba_debug_matvec    : y = (S_reduced) p for the current QR factors (p, y: 6 per free pose slot). */
int sqrtba_debug_linearize(sqrtba_handle* h, int32_t huber, double* err, double* Jp, double* Jl, double* r,
                           double* chi2);
int sqrtba_debug_step(sqrtba_handle* h, double lambda, double* dp, double* dl, double* bs, int32_t* cg_iters);
int sqrtba_debug_matvec(sqrtba_handle* h, const double* p, double* y);
int sqrtba_num_free_poses(sqrtba_handle* h);

/* The host-side plan of sqrtba_set_problem[_batch] WITHOUT a device (pure host code, used by the CPU tests of the host
 * logic): landmarks packed into items (<= 32 observations, one warp), items into tiles (<= 4 items of one window, one
 * CTA), per tile the pose-sorted ranks + run table, and for windows with more than 128 free poses the internal landmark
 * order (sorted by first free pose).  Outputs (any may be NULL):
 *   summary8       n_item, n_tile, n_run_ints, slots_fit_16_bits, pq_shared, reordered, jq_doubles lo31, jq_doubles hi
 *   tiles20        per tile (20 ints): item0, nitem, o0, o1, win, nfree, nt, is_long, nrun, cnt[4], fcnt[4],
 *                  blk_doubles, jq_off lo31, jq_off hi  (first max_tiles tiles)
 *   obs_lp         n_obs meta words (low 16 bits: window-relative free slot or 0xffff; high 16: rank in the tile)
 *   tile_run_ptr / tile_runs   CSR of the run tables (nrun+1 rank offsets, then nrun slots)
 *   landmark_order n_point: internal index -> caller's landmark index
 * Observation indices in tiles20 / obs_lp refer to the INTERNAL order when the problem was re-ordered. */
int sqrtba_debug_plan(int32_t n_win, const int64_t* win_pose_ptr, const int64_t* win_point_ptr, const int64_t* win_obs_ptr,
                      int32_t n_pose, int32_t n_point, int32_t n_obs, const uint8_t* pose_fixed, const int32_t* obs_pose,
                      const int32_t* obs_point, int32_t host_threads, int32_t* summary8, int32_t* tiles20, int32_t max_tiles,
                      uint32_t* obs_lp, int32_t* tile_run_ptr, int32_t* tile_runs, int64_t max_runs, int32_t* landmark_order);

/* Time `reps` launches of one stage kernel on the handle's stream with CUDA events (after `warmup` untimed
 * launches); returns the average ms per launch.  stage: 0 matvec, 1 linearize, 2 landmark QR, 3 cost, 4 backsub,
 * 5 fused linearise + landmark QR. */
int sqrtba_time_stage(sqrtba_handle* h, int32_t stage, int32_t warmup, int32_t reps, double* ms_avg);

#ifdef __cplusplus
}
#endif
#endif /* SQRTBA_H_ */

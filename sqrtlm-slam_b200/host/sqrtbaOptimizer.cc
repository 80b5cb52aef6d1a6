// sqrtbaOptimizer -- the adapter between the reference's map types and the sqrtba C ABI (include/sqrtba.h).
//
// It plays the role of src/backend/g2oOptimizer.cc for the two bundle-adjustment entry points:
//   LocalBundleAdjustment   window selection as g2oOptimizer.cc:709-780, gather instead of graph construction
//                           (:805-919), solve (:923-976), outlier erase + write-back (:1119-1189)
//   BundleAdjustment        g2oOptimizer.cc:110-362, GlobalBundleAdjustemnt :80-89
// Differences that are deliberate:
//   * stereo observations (mvuRight >= 0) get the stereo edge in local BA too, as upstream ORB-SLAM2 does; the
//     reference's local BA silently drops them (empty branch, g2oOptimizer.cc:914-916) although it declares their
//     thresholds (:853).
//   * vertices are handed over in ascending mnId order -- the order g2o itself imposes (sparse_optimizer.cpp:482-487).
// Error convention of the reference is kept: void, silent; the last sqrtba error string is available for logging.
#include "Optimizer.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <list>
#include <string>

#include "../../include/sqrtba.h"

namespace ORB_SLAM2 {

#ifndef SQRTBA_WITH_REFERENCE_HEADERS
std::mutex MapPoint::mGlobalMutex;
#endif

namespace {

struct Handle {
  sqrtba_handle* h = nullptr;
  std::string err;
  ~Handle() {
    if (h) sqrtba_destroy(h);
  }
  int third_pass_iters = 0;
  explicit Handle(int third = 0) : third_pass_iters(third) {}
  sqrtba_handle* get() {
    if (!h) {
      sqrtba_config cfg;
      sqrtba_default_config(&cfg);
      cfg.third_pass_iters = third_pass_iters;
      if (sqrtba_create(&cfg, &h) != SQRTBA_OK) {
        err = sqrtba_last_error(nullptr);
        h = nullptr;
      }
    }
    return h;
  }
};
thread_local Handle tl_handle;  // LocalMapping and the GBA thread each get their own (Optimizer.h is re-entrant across threads)
thread_local Handle tl_handle_lidar(20);  // local BA with the fork's lidar pass: optimize(20) after the two visual passes (:1113-1114)

// Converter::toSE3Quat (src/utils/Converter.cc:55-68) without Eigen: float 4x4 Tcw -> (t, unit quaternion with w >= 0)
void toSE3Quat(const cv::Mat& T, double out7[7]) {
  double m[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) m[i * 3 + j] = T.at<float>(i, j);
  double q[4];
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
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
    q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
  }
  if (q[3] < 0)
    for (double& v : q) v = -v;
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  out7[0] = T.at<float>(0, 3);
  out7[1] = T.at<float>(1, 3);
  out7[2] = T.at<float>(2, 3);
  for (int i = 0; i < 4; i++) out7[3 + i] = q[i] / n;
}

// Converter::toCvMat(SE3Quat) (Converter.cc:73-79, 98-109): homogeneous matrix rounded to float
cv::Mat toCvMat(const double p[7]) {
  const double x = p[3], y = p[4], z = p[5], w = p[6];
  const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w),     2 * (x * z + y * w),
                       2 * (x * y + z * w),     1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                       2 * (x * z - y * w),     2 * (y * z + x * w),     1 - 2 * (x * x + y * y)};
  cv::Mat T(4, 4, CV_32F);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T.at<float>(i, j) = (float)R[i * 3 + j];
    T.at<float>(i, 3) = (float)p[i];
  }
  T.at<float>(3, 3) = 1.f;
  return T;
}

cv::Mat toCvMat3(const double* X) {
  cv::Mat P(3, 1, CV_32F);
  for (int i = 0; i < 3; i++) P.at<float>(i) = (float)X[i];
  return P;
}

// Flat problem in the layout of sqrtba_set_problem, plus the back-references needed for write-back
struct Gathered {
  std::vector<KeyFrame*> kfs;   // ascending mnId
  std::vector<MapPoint*> mps;   // ascending mnId, only points with >= 1 usable observation
  std::vector<double> pose_qt, cam, point_xyz;
  std::vector<uint8_t> pose_fixed;
  std::vector<int32_t> obs_pose, obs_point;
  std::vector<float> obs_meas;
  std::vector<std::pair<KeyFrame*, MapPoint*>> obs_ref;
};

// kf_fixed(kf) decides setFixed; usable(kf) mirrors the per-observation filters of the reference adapters
template <class FixedFn, class UsableFn>
void gather(std::vector<KeyFrame*> kfs, std::vector<MapPoint*> mps, FixedFn kf_fixed, UsableFn usable, Gathered& g) {
  std::sort(kfs.begin(), kfs.end(), [](KeyFrame* a, KeyFrame* b) { return a->mnId < b->mnId; });
  std::sort(mps.begin(), mps.end(), [](MapPoint* a, MapPoint* b) { return a->mnId < b->mnId; });
  std::map<KeyFrame*, int> kf_index;
  g.kfs = kfs;
  for (size_t i = 0; i < kfs.size(); i++) {
    KeyFrame* kf = kfs[i];
    kf_index[kf] = (int)i;
    double p[7];
    toSE3Quat(kf->GetPose(), p);
    g.pose_qt.insert(g.pose_qt.end(), p, p + 7);
    g.pose_fixed.push_back(kf_fixed(kf) ? 1 : 0);
    const double c[5] = {kf->fx, kf->fy, kf->cx, kf->cy, kf->mbf};
    g.cam.insert(g.cam.end(), c, c + 5);
  }
  for (MapPoint* mp : mps) {
    const std::map<KeyFrame*, size_t> observations = mp->GetObservations();
    std::vector<std::pair<int, size_t>> obs;  // (pose index, keypoint index)
    for (auto& kv : observations) {
      auto it = kf_index.find(kv.first);
      if (it == kf_index.end() || !usable(kv.first)) continue;
      obs.emplace_back(it->second, kv.second);
    }
    if (obs.empty()) continue;  // vbNotIncludedMP (g2oOptimizer.cc:287-295)
    std::sort(obs.begin(), obs.end());
    const int ip = (int)g.mps.size();
    g.mps.push_back(mp);
    const cv::Mat X = mp->GetWorldPos();
    for (int i = 0; i < 3; i++) g.point_xyz.push_back(X.at<float>(i));
    for (auto& o : obs) {
      KeyFrame* kf = kfs[o.first];
      const cv::KeyPoint& kp = kf->mvKeysUn[o.second];
      const float ur = kf->mvuRight[o.second];
      g.obs_pose.push_back(o.first);
      g.obs_point.push_back(ip);
      g.obs_meas.push_back(kp.pt.x);
      g.obs_meas.push_back(kp.pt.y);
      g.obs_meas.push_back(ur < 0 ? -1.f : ur);  // mvuRight < 0 => monocular edge (g2oOptimizer.cc:208, 877)
      g.obs_meas.push_back(kf->mvInvLevelSigma2[kp.octave]);
      g.obs_ref.emplace_back(kf, mp);
    }
  }
}

bool upload(Handle& H, const Gathered& g) {
  sqrtba_handle* h = H.get();
  if (!h) return false;
  if (g.obs_pose.empty() || g.mps.empty() || g.kfs.empty()) return false;
  const int rc = sqrtba_set_problem(h, (int)g.kfs.size(), (int)g.mps.size(), (int)g.obs_pose.size(), g.pose_qt.data(),
                                    g.pose_fixed.data(), g.cam.data(), g.point_xyz.data(), g.obs_pose.data(),
                                    g.obs_point.data(), g.obs_meas.data());
  if (rc != SQRTBA_OK) {
    H.err = sqrtba_last_error(h);
    return false;
  }
  return true;
}

}  // namespace

const char* sqrtbaOptimizer::LastError() { return tl_handle.err.empty() ? tl_handle_lidar.err.c_str() : tl_handle.err.c_str(); }

void sqrtbaOptimizer::GlobalBundleAdjustemnt(Map* pMap, int nIterations, bool* pbStopFlag, const unsigned long nLoopKF,
                                             const bool bRobust) {
  std::vector<KeyFrame*> vpKFs = pMap->GetAllKeyFrames();
  std::vector<MapPoint*> vpMP = pMap->GetAllMapPoints();
  BundleAdjustment(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust);
}

void sqrtbaOptimizer::BundleAdjustment(const std::vector<KeyFrame*>& vpKFs, const std::vector<MapPoint*>& vpMP,
                                       int nIterations, bool* pbStopFlag, const unsigned long nLoopKF, const bool bRobust) {
  std::vector<KeyFrame*> kfs;
  std::vector<MapPoint*> mps;
  for (KeyFrame* kf : vpKFs)
    if (!kf->isBad()) kfs.push_back(kf);  // :146-148
  for (MapPoint* mp : vpMP)
    if (!mp->isBad()) mps.push_back(mp);  // :176-178
  Gathered g;
  gather(kfs, mps, [](KeyFrame* kf) { return kf->mnId == 0; },  // only the first keyframe is fixed (:154)
         [](KeyFrame* kf) { return !kf->isBad(); }, g);
  Handle& H = tl_handle;
  if (!upload(H, g)) return;
  sqrtba_handle* h = H.get();
  if (sqrtba_solve_global(h, nIterations, bRobust ? 1 : 0, pbStopFlag, nullptr) != SQRTBA_OK) {
    H.err = sqrtba_last_error(h);
    return;
  }
  std::vector<double> P(g.kfs.size() * 7), X(g.mps.size() * 3);
  if (sqrtba_get_poses(h, P.data()) != SQRTBA_OK || sqrtba_get_points(h, X.data()) != SQRTBA_OK) {
    H.err = sqrtba_last_error(h);
    return;
  }
  for (size_t i = 0; i < g.kfs.size(); i++) {  // :308-330
    KeyFrame* kf = g.kfs[i];
    if (nLoopKF == 0) {
      kf->SetPose(toCvMat(&P[i * 7]));
    } else {
      kf->mTcwGBA.create(4, 4, CV_32F);
      toCvMat(&P[i * 7]).copyTo(kf->mTcwGBA);
      kf->mnBAGlobalForKF = nLoopKF;
    }
  }
  for (size_t i = 0; i < g.mps.size(); i++) {  // :334-361
    MapPoint* mp = g.mps[i];
    if (nLoopKF == 0) {
      mp->SetWorldPos(toCvMat3(&X[i * 3]));
      mp->UpdateNormalAndDepth();
    } else {
      mp->mPosGBA.create(3, 1, CV_32F);
      toCvMat3(&X[i * 3]).copyTo(mp->mPosGBA);
      mp->mnBAGlobalForKF = nLoopKF;
    }
  }
}

namespace {
// The clouds of the lidar pass (g2oOptimizer.cc:985-1013, 1033-1107) in the layout of sqrtba_set_lidar: the current
// keyframe's features in its own frame, and the features of every OTHER local keyframe tagged with that keyframe's pose
// index.  The reference moves the map clouds to the world frame and searches a kd-tree on the host; both happen on the
// device here, at the estimates of the second pass.
struct LidarClouds {
  std::vector<float> flat, flat_n, corner, map_flat, map_corner;
  std::vector<int32_t> map_flat_pose, map_corner_pose;
};
void append(const PointIRTCloud& c, std::vector<float>& xyz, std::vector<int32_t>* pose, int idx) {
  for (const PointIRT& p : c.points) {
    xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z);
    if (pose) pose->push_back(idx);
  }
}
bool set_lidar(sqrtba_handle* h, const Gathered& g, KeyFrame* pKF, const lidarConfig* cfg) {
  LidarClouds L;
  int cur = -1;
  for (size_t i = 0; i < g.kfs.size(); i++) {
    KeyFrame* kf = g.kfs[i];
    if (kf == pKF) { cur = (int)i; continue; }
    if (kf->mnBALocalForKF != pKF->mnId) continue;  // lLocalKeyFrames only (:986-992); fixed keyframes add nothing
    append(kf->corner_points_less_sharp_, L.map_corner, &L.map_corner_pose, (int)i);
    append(kf->surface_points_less_flat_, L.map_flat, &L.map_flat_pose, (int)i);
  }
  if (cur < 0) return false;
  append(pKF->surface_points_less_flat_, L.flat, nullptr, 0);
  append(pKF->surface_points_less_flat_normal_, L.flat_n, nullptr, 0);
  append(pKF->corner_points_less_sharp_, L.corner, nullptr, 0);
  if (L.flat_n.size() != L.flat.size()) return false;  // the normal cloud is index-aligned with the flat cloud (:1060-1062)
  sqrtba_lidar c{};
  c.cur_pose = cur;
  c.n_flat = (int32_t)(L.flat.size() / 3); c.flat_xyz = L.flat.data(); c.flat_normal = L.flat_n.data();
  c.n_corner = (int32_t)(L.corner.size() / 3); c.corner_xyz = L.corner.data();
  c.numeric_jacobian = 1;  // BaseUnaryEdge::linearizeOplus, as the reference
  c.n_map_flat = (int64_t)L.map_flat_pose.size(); c.map_flat_xyz = L.map_flat.data(); c.map_flat_pose = L.map_flat_pose.data();
  c.n_map_corner = (int64_t)L.map_corner_pose.size(); c.map_corner_xyz = L.map_corner.data(); c.map_corner_pose = L.map_corner_pose.data();
  c.distance_sq_threshold = cfg->distance_sq_threshold;
  c.flat_weight = cfg->flat_optimized_weight;
  c.corner_weight = cfg->corner_optimized_weight;
  c.use_flat = cfg->using_flat_point;
  c.use_corner = cfg->using_sharp_point;
  if (c.n_flat + c.n_corner == 0) return true;  // nothing to match: the plain third pass
  return sqrtba_set_lidar(h, &c) == SQRTBA_OK;
}
}  // namespace

void sqrtbaOptimizer::LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, const lidarConfig* lidarconfig) {
  // ---- local keyframes: pKF + its covisible keyframes (g2oOptimizer.cc:709-727)
  std::list<KeyFrame*> lLocalKeyFrames;
  lLocalKeyFrames.push_back(pKF);
  pKF->mnBALocalForKF = pKF->mnId;
  const std::vector<KeyFrame*> vNeighKFs = pKF->GetVectorCovisibleKeyFrames();
  for (KeyFrame* pKFi : vNeighKFs) {
    pKFi->mnBALocalForKF = pKF->mnId;
    if (!pKFi->isBad()) lLocalKeyFrames.push_back(pKFi);
  }
  // ---- local map points: everything the local keyframes see (:731-755)
  std::list<MapPoint*> lLocalMapPoints;
  for (KeyFrame* kf : lLocalKeyFrames) {
    std::vector<MapPoint*> vpMPs = kf->GetMapPointMatches();
    for (MapPoint* pMP : vpMPs)
      if (pMP && !pMP->isBad() && pMP->mnBALocalForKF != pKF->mnId) {
        lLocalMapPoints.push_back(pMP);
        pMP->mnBALocalForKF = pKF->mnId;
      }
  }
  // ---- fixed keyframes: other observers of the local points (:759-780)
  std::list<KeyFrame*> lFixedCameras;
  for (MapPoint* mp : lLocalMapPoints) {
    std::map<KeyFrame*, size_t> observations = mp->GetObservations();
    for (auto& kv : observations) {
      KeyFrame* pKFi = kv.first;
      if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) {
        pKFi->mnBAFixedForKF = pKF->mnId;
        if (!pKFi->isBad()) lFixedCameras.push_back(pKFi);
      }
    }
  }
  std::vector<KeyFrame*> kfs(lLocalKeyFrames.begin(), lLocalKeyFrames.end());
  kfs.insert(kfs.end(), lFixedCameras.begin(), lFixedCameras.end());
  std::vector<MapPoint*> mps(lLocalMapPoints.begin(), lLocalMapPoints.end());
  const unsigned long cur = pKF->mnId;
  Gathered g;
  gather(kfs, mps,
         [cur](KeyFrame* kf) { return kf->mnBALocalForKF != cur || kf->mnId == 0; },  // :813, :829
         [](KeyFrame* kf) { return !kf->isBad(); }, g);                                // :872
  if (pbStopFlag && *pbStopFlag) return;  // :923-928
  // the fork's third pass (lidar edges on pKF + optimize(20), :979-1117) runs when the configuration asks for lidar
  // features; without them the two-pass schedule of ORB-SLAM2 is used
  const bool with_lidar = lidarconfig && (lidarconfig->using_flat_point || lidarconfig->using_sharp_point);
  Handle& H = with_lidar ? tl_handle_lidar : tl_handle;
  if (!upload(H, g)) { tl_handle.err = H.err; return; }
  sqrtba_handle* h = H.get();
  if (with_lidar && !set_lidar(h, g, pKF, lidarconfig)) {
    tl_handle.err = sqrtba_last_error(h);
    return;
  }
  if (sqrtba_solve_local(h, pbStopFlag, nullptr) != SQRTBA_OK) {
    H.err = sqrtba_last_error(h);
    return;
  }
  std::vector<double> P(g.kfs.size() * 7), X(g.mps.size() * 3);
  std::vector<uint8_t> flags(g.obs_ref.size());
  if (sqrtba_get_poses(h, P.data()) != SQRTBA_OK || sqrtba_get_points(h, X.data()) != SQRTBA_OK ||
      sqrtba_get_outliers(h, flags.data()) != SQRTBA_OK) {
    H.err = sqrtba_last_error(h);
    return;
  }
  // ---- erase outlier observations and write the estimates back under the map mutex (:1145-1189)
  std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
  for (size_t k = 0; k < flags.size(); k++)
    if (flags[k]) {
      KeyFrame* pKFi = g.obs_ref[k].first;
      MapPoint* pMPi = g.obs_ref[k].second;
      pKFi->EraseMapPointMatch(pMPi);
      pMPi->EraseObservation(pKFi);
    }
  for (size_t i = 0; i < g.kfs.size(); i++)
    if (g.kfs[i]->mnBALocalForKF == cur) g.kfs[i]->SetPose(toCvMat(&P[i * 7]));
  for (size_t i = 0; i < g.mps.size(); i++) {
    g.mps[i]->SetWorldPos(toCvMat3(&X[i * 3]));
    g.mps[i]->UpdateNormalAndDepth();
  }
}

// ---- the facade (src/backend/Optimizer.cc:26-79) with the new selector value
Optimizer::eSolver solver = Optimizer::SQRTBA;

// g2oOptimizer::PoseOptimization (g2oOptimizer.cc:385-559, 655-690) for one or several frames.  Edge wiring as in the
// reference: a matched map point with mvuRight < 0 becomes a monocular edge; the fork leaves the stereo branch empty
// (:481-483), so a point with a right coordinate is neither optimised nor counted -- define SQRTBA_POSEOPT_STEREO to get
// upstream ORB-SLAM2's stereo edges instead.
void sqrtbaOptimizer::PoseOptimizationBatch(const std::vector<Frame*>& frames, std::vector<int>& inliers) {
  inliers.assign(frames.size(), 0);
  Handle& H = tl_handle;
  if (frames.empty() || !H.get()) return;
  const int nf = (int)frames.size();
  std::vector<int64_t> ptr(nf + 1, 0);
  std::vector<double> pose((size_t)nf * 7), cam((size_t)nf * 5), xyz;
  std::vector<float> meas;
  std::vector<std::pair<int, int>> ref;  // (frame, keypoint index) of every edge
  {
    std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);  // :433
    for (int f = 0; f < nf; f++) {
      Frame* F = frames[f];
      toSE3Quat(F->mTcw, &pose[(size_t)f * 7]);
      cam[f * 5 + 0] = F->fx; cam[f * 5 + 1] = F->fy; cam[f * 5 + 2] = F->cx; cam[f * 5 + 3] = F->cy; cam[f * 5 + 4] = F->mbf;
      for (int i = 0; i < F->N; i++) {
        MapPoint* pMP = F->mvpMapPoints[i];
        if (!pMP) continue;
        const bool mono = F->mvuRight[i] < 0;
#ifndef SQRTBA_POSEOPT_STEREO
        if (!mono) continue;
#endif
        F->mvbOutlier[i] = false;
        const cv::KeyPoint& kp = F->mvKeysUn[i];
        const cv::Mat Xw = pMP->GetWorldPos();
        for (int c = 0; c < 3; c++) xyz.push_back(Xw.at<float>(c));
        meas.push_back(kp.pt.x);
        meas.push_back(kp.pt.y);
        meas.push_back(mono ? -1.f : F->mvuRight[i]);
        meas.push_back(F->mvInvLevelSigma2[kp.octave]);
        ref.emplace_back(f, i);
      }
      ptr[f + 1] = (int64_t)ref.size();
    }
  }
  std::vector<uint8_t> out(std::max<size_t>(ref.size(), 1), 0);
  std::vector<int32_t> inl(nf, 0);
  if (sqrtba_pose_opt(H.h, nf, ptr.data(), pose.data(), cam.data(), xyz.data(), meas.data(), out.data(), inl.data(), nullptr) !=
      SQRTBA_OK) {
    H.err = sqrtba_last_error(H.h);
    return;
  }
  for (size_t k = 0; k < ref.size(); k++) frames[ref[k].first]->mvbOutlier[ref[k].second] = out[k] != 0;
  for (int f = 0; f < nf; f++) {
    if (ptr[f + 1] - ptr[f] >= 3) frames[f]->SetPose(toCvMat(&pose[(size_t)f * 7]));  // :491-492: nothing happens below 3
    inliers[f] = inl[f];
  }
}

int sqrtbaOptimizer::PoseOptimization(Frame* pFrame) {
  std::vector<int> inl;
  PoseOptimizationBatch(std::vector<Frame*>{pFrame}, inl);
  return inl.empty() ? 0 : inl[0];
}

void Optimizer::GlobalBundleAdjustemnt(Map* pMap, int nIterations, bool* pbStopFlag, const unsigned long nLoopKF,
                                       const bool bRobust) {
  sqrtbaOptimizer::GlobalBundleAdjustemnt(pMap, nIterations, pbStopFlag, nLoopKF, bRobust);
}
void Optimizer::BundleAdjustment(const std::vector<KeyFrame*>& vpKFs, const std::vector<MapPoint*>& vpMP, int nIterations,
                                 bool* pbStopFlag, const unsigned long nLoopKF, const bool bRobust) {
  sqrtbaOptimizer::BundleAdjustment(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust);
}
void Optimizer::LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, const lidarConfig* lidarconfig) {
  sqrtbaOptimizer::LocalBundleAdjustment(pKF, pbStopFlag, pMap, lidarconfig);
}

}  // namespace ORB_SLAM2

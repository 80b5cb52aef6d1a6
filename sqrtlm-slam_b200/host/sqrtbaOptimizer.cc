// sqrtbaOptimizer -- the adapter between the reference's map types and the sqrtba C ABI (include/sqrtba.h).
//
// It plays the role of src/backend/g2oOptimizer.cc for the two bundle-adjustment entry points:
//   LocalBundleAdjustment   window selection as g2oOptimizer.cc:709-780, gather instead of graph construction
//                           (:805-919), solve (:923-976), outlier erase + write-back (:1119-1189)
//   BundleAdjustment        g2oOptimizer.cc:110-362, GlobalBundleAdjustemnt :80-89
// Defaults follow the reference fork (sqrtbaOptimizer::options(), Optimizer.h): local BA creates monocular edges only
// -- observations with mvuRight >= 0 fall into the fork's empty stereo branch (g2oOptimizer.cc:914-916) -- and always
// ends with the third optimize(20) (:1113-1114); upstream ORB-SLAM2's stereo edges / two-pass schedule are opt-in.
// Differences that are deliberate:
//   * vertices are handed over in ascending mnId order -- the order g2o itself imposes (sparse_optimizer.cpp:482-487).
// Error convention of the reference is kept: void, silent; the last sqrtba error string is available for logging.
#include "Optimizer.h"

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <list>
#include <thread>
#include <unordered_map>
#include <string>

#include "../../include/sqrtba.h"
#include "../csrc/sqrtba_sim3.cuh"  // g2o::Sim3's product / inverse restated (host build of the device arithmetic)
#include "host_pool.h"
#include "map_mirror.h"

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
  // local map points without a usable observation: vertices without edges in the reference's graph (:856-866) -- their
  // estimate cannot move, but the write-back still visits them (SetWorldPos + UpdateNormalAndDepth, :1180-1188)
  std::vector<MapPoint*> edgeless;
};

// One observation of a map point as MapPoint::GetObservations() reports it
struct ObsRec { KeyFrame* kf; size_t idx; };

// The observation lists of many map points, copied ONCE from the map (MapPoint::GetObservations() returns a copy of a
// std::map under the point's mutex, MapPoint.cc:212-215 -- ~12 node allocations per point; the reference walks it twice
// per local-BA call, :759-780 and :870-919) into one flat array.  Filled on all host cores.
struct ObsCache {
  std::vector<size_t> ptr;   // n_mp + 1
  std::vector<ObsRec> rec;
};

int host_threads(size_t work_items) {
  static const int hw = [] {
    const char* e = std::getenv("SQRTBA_ADAPTER_THREADS");   // override for measurements
    return e ? std::max(1, std::atoi(e)) : (int)std::max(1u, std::thread::hardware_concurrency());
  }();
  return (int)std::max<size_t>(1, std::min<size_t>({(size_t)hw, (size_t)16, work_items / 512 + 1}));
}

// helper threads of the calling thread (LocalMapping and the global-BA thread each get their own), created on first
// use and parked between loops
thread_local sqrtba::HostPool tl_pool;

template <class F>
void parallel_ranges(size_t n, int n_thr, F f) {   // f(part, begin, end) over n_thr contiguous ranges
  tl_pool.ranges(n, n_thr, f);
}

// With the incremental mirror switched on (host/map_mirror.h: the map's mutation sites keep flat per-point lists up to
// date) the lists are read from it -- same order, same content, no std::map copies; a point the mirror does not know
// sends the whole call back to the map copies.
bool fill_obs_cache_from_mirror(const std::vector<MapPoint*>& mps, ObsCache& c) {
  sqrtba::MapMirror& mm = sqrtba::MapMirror::Global();
  if (!mm.enabled() || mps.empty()) return false;
  // a snapshot is ~0.25 us per point (one shard lock, one hash lookup, one small copy): a few threads saturate it
  const int T = std::min(4, host_threads(mps.size() / 4));
  std::vector<std::vector<size_t>> ptr(T);
  std::vector<std::vector<sqrtba::MapMirror::Obs>> rec(T);
  std::vector<uint8_t> ok(T, 1);
  parallel_ranges(mps.size(), T, [&](int t, size_t a, size_t b) {
    ok[t] = mm.Snapshot(mps.data() + a, b - a, ptr[t], rec[t]) ? 1 : 0;
  });
  for (int t = 0; t < T; t++)
    if (!ok[t]) return false;
  size_t total = 0;
  for (auto& r : rec) total += r.size();
  c.ptr.assign(1, 0);
  c.ptr.reserve(mps.size() + 1);
  c.rec.clear();
  c.rec.reserve(total);
  static_assert(sizeof(ObsRec) == sizeof(sqrtba::MapMirror::Obs) && offsetof(ObsRec, idx) == offsetof(sqrtba::MapMirror::Obs, idx),
                "the mirror's records are copied as they are");
  c.rec.resize(total);
  size_t base = 0;
  for (int t = 0; t < T; t++) {
    for (size_t i = 1; i < ptr[t].size(); i++) c.ptr.push_back(base + ptr[t][i]);
    if (!rec[t].empty()) std::memcpy((void*)(c.rec.data() + base), rec[t].data(), rec[t].size() * sizeof(ObsRec));
    base += rec[t].size();
  }
  return c.ptr.size() == mps.size() + 1;
}

void fill_obs_cache(const std::vector<MapPoint*>& mps, ObsCache& c) {
  if (fill_obs_cache_from_mirror(mps, c)) return;
  const int T = host_threads(mps.size());
  std::vector<std::vector<ObsRec>> part(T);
  std::vector<std::vector<size_t>> cnt(T);
  parallel_ranges(mps.size(), T, [&](int t, size_t a, size_t b) {
    part[t].reserve((b - a) * 12);
    cnt[t].reserve(b - a);
    for (size_t i = a; i < b; i++) {
      const std::map<KeyFrame*, size_t> observations = mps[i]->GetObservations();
      for (auto& kv : observations) part[t].push_back(ObsRec{kv.first, kv.second});
      cnt[t].push_back(observations.size());
    }
  });
  c.ptr.assign(1, 0);
  c.ptr.reserve(mps.size() + 1);
  size_t total = 0;
  for (auto& p : part) total += p.size();
  c.rec.clear();
  c.rec.reserve(total);
  for (int t = 0; t < T; t++) {
    for (size_t n : cnt[t]) c.ptr.push_back(c.ptr.back() + n);
    c.rec.insert(c.rec.end(), part[t].begin(), part[t].end());
  }
}

// kf_fixed(kf) decides setFixed; usable(kf) mirrors the per-observation filters of the reference adapters.
// `cache` (optional) holds the observation lists of `mps` in the order given (copied earlier by the caller).
// Two parallel sweeps over the map points in mnId order -- count the usable observations, then (after a prefix sum) fill
// the flat arrays in place -- so nothing is merged or re-allocated and the result does not depend on the thread count.
// drop_stereo: observations with a right coordinate are left out (the fork's local BA, g2oOptimizer.cc:883-916).
template <class FixedFn, class UsableFn>
void gather(std::vector<KeyFrame*> kfs, std::vector<MapPoint*> mps, FixedFn kf_fixed, UsableFn usable, Gathered& g,
            const ObsCache* cache = nullptr, bool drop_stereo = false) {
  ObsCache own;
  if (!cache) {
    fill_obs_cache(mps, own);
    cache = &own;
  }
  std::sort(kfs.begin(), kfs.end(), [](KeyFrame* a, KeyFrame* b) { return a->mnId < b->mnId; });
  std::vector<size_t> order(mps.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return mps[a]->mnId < mps[b]->mnId; });
  std::unordered_map<KeyFrame*, int> kf_index;
  kf_index.reserve(kfs.size() * 2);
  std::vector<uint8_t> kf_usable(kfs.size());
  g.kfs = kfs;
  g.pose_qt.reserve(kfs.size() * 7);
  g.cam.reserve(kfs.size() * 5);
  for (size_t i = 0; i < kfs.size(); i++) {
    KeyFrame* kf = kfs[i];
    kf_index[kf] = (int)i;
    kf_usable[i] = usable(kf) ? 1 : 0;
    double p[7];
    toSE3Quat(kf->GetPose(), p);
    g.pose_qt.insert(g.pose_qt.end(), p, p + 7);
    g.pose_fixed.push_back(kf_fixed(kf) ? 1 : 0);
    const double c[5] = {kf->fx, kf->fy, kf->cx, kf->cy, kf->mbf};
    g.cam.insert(g.cam.end(), c, c + 5);
  }
  const size_t n = order.size();
  const int T = host_threads(n);
  // sweep 1: pose index of every cached observation (-1: keyframe not in the problem or not usable), count per point
  std::vector<int> rec_pose(cache->rec.size());
  std::vector<size_t> cnt(n + 1, 0);  // cnt[r + 1] = usable observations of the r-th point in mnId order
  parallel_ranges(n, T, [&](int, size_t a, size_t b) {
    for (size_t r = a; r < b; r++) {
      size_t c = 0;
      for (size_t k = cache->ptr[order[r]]; k < cache->ptr[order[r] + 1]; k++) {
        auto it = kf_index.find(cache->rec[k].kf);
        int ip = (it != kf_index.end() && kf_usable[it->second]) ? it->second : -1;
        if (ip >= 0 && drop_stereo && !(cache->rec[k].kf->mvuRight[cache->rec[k].idx] < 0)) ip = -1;
        rec_pose[k] = ip;
        c += ip >= 0;
      }
      cnt[r + 1] = c;
    }
  });
  // prefix sums: observation offset and compact point index (points without a usable observation drop out,
  // vbNotIncludedMP, g2oOptimizer.cc:287-295)
  std::vector<size_t> pt_index(n + 1, 0);
  for (size_t r = 0; r < n; r++) {
    pt_index[r + 1] = pt_index[r] + (cnt[r + 1] > 0);
    if (cnt[r + 1] == 0) g.edgeless.push_back(mps[order[r]]);
    cnt[r + 1] += cnt[r];
  }
  const size_t n_pt = pt_index[n], n_ob = cnt[n];
  g.mps.resize(n_pt);
  g.point_xyz.resize(n_pt * 3);
  g.obs_pose.resize(n_ob);
  g.obs_point.resize(n_ob);
  g.obs_meas.resize(n_ob * 4);
  g.obs_ref.resize(n_ob);
  // sweep 2: fill in place, observations of a point sorted by pose index
  parallel_ranges(n, T, [&](int, size_t a, size_t b) {
    std::vector<std::pair<int, size_t>> obs;  // (pose index, keypoint index), reused
    for (size_t r = a; r < b; r++) {
      if (cnt[r + 1] == cnt[r]) continue;
      MapPoint* mp = mps[order[r]];
      obs.clear();
      for (size_t k = cache->ptr[order[r]]; k < cache->ptr[order[r] + 1]; k++)
        if (rec_pose[k] >= 0) obs.emplace_back(rec_pose[k], cache->rec[k].idx);
      std::sort(obs.begin(), obs.end());
      const size_t ip = pt_index[r];
      g.mps[ip] = mp;
      const cv::Mat X = mp->GetWorldPos();
      for (int i = 0; i < 3; i++) g.point_xyz[ip * 3 + i] = X.at<float>(i);
      size_t o = cnt[r];
      for (auto& ob : obs) {
        KeyFrame* kf = kfs[ob.first];
        const cv::KeyPoint& kp = kf->mvKeysUn[ob.second];
        const float ur = kf->mvuRight[ob.second];
        g.obs_pose[o] = ob.first;
        g.obs_point[o] = (int32_t)ip;
        g.obs_meas[o * 4 + 0] = kp.pt.x;
        g.obs_meas[o * 4 + 1] = kp.pt.y;
        g.obs_meas[o * 4 + 2] = ur < 0 ? -1.f : ur;  // mvuRight < 0 => monocular edge (g2oOptimizer.cc:208, 877)
        g.obs_meas[o * 4 + 3] = kf->mvInvLevelSigma2[kp.octave];
        g.obs_ref[o] = std::make_pair(kf, mp);
        o++;
      }
    }
  });
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
  parallel_ranges(g.mps.size(), host_threads(g.mps.size()), [&](int, size_t a, size_t b) {  // :334-361, per-point state only
    for (size_t i = a; i < b; i++) {
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
  });
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

namespace {
// Window selection + gather of LocalBundleAdjustment (g2oOptimizer.cc:709-780, 805-919): everything the host does
// before the solve.  SQRTBA_HOST_TIMING=1 prints the wall time of each step on stderr.
void gather_local_window(KeyFrame* pKF, Gathered& g) {
  static const bool host_timing = std::getenv("SQRTBA_HOST_TIMING") != nullptr;
  auto tick = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!host_timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[sqrtba adapter] %-24s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
    tick = now;
  };
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
  lap("local keyframes + points");
  // ---- fixed keyframes: other observers of the local points (:759-780).  The observation lists are copied from the
  //      map once, on all host cores, and serve both this loop and the gather below (the reference copies every
  //      std::map twice, :761 and :870)
  std::vector<MapPoint*> mps(lLocalMapPoints.begin(), lLocalMapPoints.end());
  ObsCache cache;
  fill_obs_cache(mps, cache);
  lap("observation copies");
  std::list<KeyFrame*> lFixedCameras;
  for (const ObsRec& r : cache.rec) {
    KeyFrame* pKFi = r.kf;
    if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) {
      pKFi->mnBAFixedForKF = pKF->mnId;
      if (!pKFi->isBad()) lFixedCameras.push_back(pKFi);
    }
  }
  lap("window selection");
  std::vector<KeyFrame*> kfs(lLocalKeyFrames.begin(), lLocalKeyFrames.end());
  kfs.insert(kfs.end(), lFixedCameras.begin(), lFixedCameras.end());
  const unsigned long cur = pKF->mnId;
  gather(kfs, mps,
         [cur](KeyFrame* kf) { return kf->mnBALocalForKF != cur || kf->mnId == 0; },  // :813, :829
         [](KeyFrame* kf) { return !kf->isBad(); }, g, &cache,                         // :872
         !sqrtbaOptimizer::options().local_ba_stereo_edges);                           // :883-916
  lap("gather");
}
}  // namespace

namespace {
// Erase the outlier observations and write the estimates back under the map mutex (g2oOptimizer.cc:1145-1189).  The
// erasures and SetPose stay serial (they touch shared keyframes); SetWorldPos + UpdateNormalAndDepth only touch their
// own map point (per-point mutexes, MapPoint.cc:118-124, 531-575) and read keyframe poses that are final by then, so
// they run on all host cores -- UpdateNormalAndDepth copies the point's observation map again and is the bulk of the
// time mMutexMapUpdate is held, which is time the tracking thread waits.
void write_back_local(const Gathered& g, unsigned long cur, const double* P, const double* X, const uint8_t* flags, Map* pMap) {
  std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
  for (size_t k = 0; k < g.obs_ref.size(); k++)
    if (flags[k]) {
      KeyFrame* pKFi = g.obs_ref[k].first;
      MapPoint* pMPi = g.obs_ref[k].second;
      pKFi->EraseMapPointMatch(pMPi);
      pMPi->EraseObservation(pKFi);
    }
  for (size_t i = 0; i < g.kfs.size(); i++)
    if (g.kfs[i]->mnBALocalForKF == cur) g.kfs[i]->SetPose(toCvMat(&P[i * 7]));
  parallel_ranges(g.mps.size(), host_threads(g.mps.size()), [&](int, size_t a, size_t b) {
    for (size_t i = a; i < b; i++) {
      g.mps[i]->SetWorldPos(toCvMat3(&X[i * 3]));
      g.mps[i]->UpdateNormalAndDepth();
    }
  });
  // points that had no edge keep their position (float -> double -> float is exact) but see the new keyframe poses
  for (MapPoint* mp : g.edgeless) mp->UpdateNormalAndDepth();
}
}  // namespace

sqrtbaOptimizer::Options& sqrtbaOptimizer::options() {
  static Options o;
  return o;
}

void sqrtbaOptimizer::LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, const lidarConfig* lidarconfig) {
  const unsigned long cur = pKF->mnId;
  Gathered g;
  gather_local_window(pKF, g);
  if (pbStopFlag && *pbStopFlag) return;  // :923-928
  // The fork runs its third optimize(20) unconditionally after the lidar block (:1113-1114) -- with both feature kinds
  // switched off, or without a single match, it is simply 20 more iterations over the visual edges.  The lidar clouds
  // are handed over only when the configuration asks for a feature kind.  options().local_ba_two_pass selects upstream
  // ORB-SLAM2's 5 + 10 schedule instead (then there is no pass the lidar edges could join).
  const bool two_pass = options().local_ba_two_pass;
  const bool with_lidar = !two_pass && lidarconfig && (lidarconfig->using_flat_point || lidarconfig->using_sharp_point);
  Handle& H = two_pass ? tl_handle : tl_handle_lidar;
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
  write_back_local(g, cur, P.data(), X.data(), flags.data(), pMap);
}


// ---- essential graph (g2oOptimizer::OptimizeEssentialGraph, src/backend/g2oOptimizer.cc:1212-1520) -------------------
namespace {
// g2o::Sim3 <-> the 8 doubles of sqrtba_pose_graph (qx qy qz qw | tx ty tz | s)
void sim3_to8(const g2o::Sim3& S, double* o) {
  o[0] = S.rotation().x(); o[1] = S.rotation().y(); o[2] = S.rotation().z(); o[3] = S.rotation().w();
  o[4] = S.translation()[0]; o[5] = S.translation()[1]; o[6] = S.translation()[2];
  o[7] = S.scale();
}
// g2o::Sim3(Rcw, tcw, 1.0) from the keyframe's float pose (:1276-1280): Eigen's Quaterniond(Matrix3d), not normalised
void pose_to_sim3(KeyFrame* pKF, double* o) {
  const cv::Mat R = pKF->GetRotation(), t = pKF->GetTranslation();
  double m[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) m[i * 3 + j] = R.at<float>(i, j);
  sqrtba::R_to_quat(m, o);
  for (int i = 0; i < 3; i++) o[4 + i] = t.at<float>(i);
  o[7] = 1.0;
}

struct EssentialGraph {
  std::vector<KeyFrame*> vpKFs;
  std::vector<double> vScw;          // (nMaxKFid + 1) x 8: the vertices' initial estimates, indexed by mnId (:1244)
  std::vector<uint8_t> fixed, present;
  std::vector<int32_t> edge_ij;
  std::vector<double> meas8;
};

void build_essential_graph(Map* pMap, KeyFrame* pLoopKF, KeyFrame* pCurKF, const LoopClosing::KeyFrameAndPose& NonCorrectedSim3,
                           const LoopClosing::KeyFrameAndPose& CorrectedSim3,
                           const std::map<KeyFrame*, std::set<KeyFrame*>>& LoopConnections, EssentialGraph& G) {
  G.vpKFs = pMap->GetAllKeyFrames();
  const size_t nMax = pMap->GetMaxKFid();
  G.vScw.assign((nMax + 1) * 8, 0.0);
  G.fixed.assign(nMax + 1, 0);
  G.present.assign(nMax + 1, 0);
  const int minFeat = 100;  // :1251
  // Step 2 (:1258-1298): one vertex per good keyframe, the Sim3-corrected pose where loop closing has one
  for (KeyFrame* pKF : G.vpKFs) {
    if (pKF->isBad()) continue;
    const size_t id = pKF->mnId;
    auto it = CorrectedSim3.find(pKF);
    if (it != CorrectedSim3.end()) sim3_to8(it->second, &G.vScw[id * 8]);
    else pose_to_sim3(pKF, &G.vScw[id * 8]);
    G.present[id] = 1;
    if (pKF == pLoopKF) G.fixed[id] = 1;
  }
  auto S = [&](size_t id) { return &G.vScw[id * 8]; };
  auto add_edge = [&](size_t i, size_t j, const double* Sjw, const double* Swi) {  // measurement S_ji = S_jw * S_wi
    if (i > nMax || j > nMax || !G.present[i] || !G.present[j]) return;  // the reference would dereference a null vertex
    double Sji[8];
    sqrtba::sim3_mul(Sjw, Swi, Sji);
    G.edge_ij.push_back((int32_t)i);
    G.edge_ij.push_back((int32_t)j);
    G.meas8.insert(G.meas8.end(), Sji, Sji + 8);
  };
  std::set<std::pair<unsigned long, unsigned long>> sInsertedEdges;
  // Step 3 (:1306-1336): the new connections the loop fusion created
  for (auto& kv : LoopConnections) {
    KeyFrame* pKF = kv.first;
    const unsigned long nIDi = pKF->mnId;
    if (nIDi > nMax || !G.present[nIDi]) continue;
    double Swi[8];
    sqrtba::sim3_inv(S(nIDi), Swi);
    for (KeyFrame* pKFn : kv.second) {
      const unsigned long nIDj = pKFn->mnId;
      if ((nIDi != pCurKF->mnId || nIDj != pLoopKF->mnId) && pKF->GetWeight(pKFn) < minFeat) continue;
      if (nIDj > nMax || !G.present[nIDj]) continue;
      add_edge(nIDi, nIDj, S(nIDj), Swi);
      sInsertedEdges.insert(std::make_pair(std::min(nIDi, nIDj), std::max(nIDi, nIDj)));
    }
  }
  // Step 4 (:1339-1448): spanning tree, earlier loop edges, strong covisibility -- relative poses from BEFORE the correction
  auto prior = [&](KeyFrame* pKF, double* out) {  // S_kw: the non-corrected pose if loop closing recorded one
    auto it = NonCorrectedSim3.find(pKF);
    if (it != NonCorrectedSim3.end()) sim3_to8(it->second, out);
    else std::memcpy(out, S(pKF->mnId), 8 * sizeof(double));
  };
  for (KeyFrame* pKF : G.vpKFs) {
    const unsigned long nIDi = pKF->mnId;
    if (pKF->isBad()) continue;  // (the reference would index a vertex that was never added)
    double Siw[8], Swi[8];
    prior(pKF, Siw);
    sqrtba::sim3_inv(Siw, Swi);
    KeyFrame* pParentKF = pKF->GetParent();
    if (pParentKF && !pParentKF->isBad()) {
      double Sjw[8];
      prior(pParentKF, Sjw);
      add_edge(nIDi, pParentKF->mnId, Sjw, Swi);
    }
    const std::set<KeyFrame*> sLoopEdges = pKF->GetLoopEdges();
    for (KeyFrame* pLKF : sLoopEdges)
      if (pLKF->mnId < pKF->mnId && !pLKF->isBad()) {
        double Slw[8];
        prior(pLKF, Slw);
        add_edge(nIDi, pLKF->mnId, Slw, Swi);
      }
    const std::vector<KeyFrame*> vpConnectedKFs = pKF->GetCovisiblesByWeight(minFeat);
    for (KeyFrame* pKFn : vpConnectedKFs)
      if (pKFn && pKFn != pParentKF && !pKF->hasChild(pKFn) && !sLoopEdges.count(pKFn))
        if (!pKFn->isBad() && pKFn->mnId < pKF->mnId) {
          if (sInsertedEdges.count(std::make_pair(std::min(pKF->mnId, pKFn->mnId), std::max(pKF->mnId, pKFn->mnId)))) continue;
          double Snw[8];
          prior(pKFn, Snw);
          add_edge(nIDi, pKFn->mnId, Snw, Swi);
        }
  }
}
}  // namespace

void sqrtbaOptimizer::GatherEssentialGraph(Map* pMap, KeyFrame* pLoopKF, KeyFrame* pCurKF,
                                           const LoopClosing::KeyFrameAndPose& NonCorrectedSim3,
                                           const LoopClosing::KeyFrameAndPose& CorrectedSim3,
                                           const std::map<KeyFrame*, std::set<KeyFrame*>>& LoopConnections, PoseGraphProblem& out) {
  EssentialGraph G;
  build_essential_graph(pMap, pLoopKF, pCurKF, NonCorrectedSim3, CorrectedSim3, LoopConnections, G);
  out.vert8 = G.vScw; out.meas8 = G.meas8; out.fixed = G.fixed; out.present = G.present;
  out.edge_ij.assign(G.edge_ij.begin(), G.edge_ij.end());
}

void sqrtbaOptimizer::OptimizeEssentialGraph(Map* pMap, KeyFrame* pLoopKF, KeyFrame* pCurKF,
                                             const LoopClosing::KeyFrameAndPose& NonCorrectedSim3,
                                             const LoopClosing::KeyFrameAndPose& CorrectedSim3,
                                             const std::map<KeyFrame*, std::set<KeyFrame*>>& LoopConnections, const bool& bFixScale) {
  EssentialGraph G;
  build_essential_graph(pMap, pLoopKF, pCurKF, NonCorrectedSim3, CorrectedSim3, LoopConnections, G);
  Handle& H = tl_handle;
  sqrtba_handle* h = H.get();
  if (!h) return;
  std::vector<double> est = G.vScw;  // absent ids are vertices without edges for the solver: never touched
  const int n_vert = (int)G.fixed.size();
  // solver->setUserLambdaInit(1e-16); optimizer.optimize(20)  (:1230, :1452-1453)
  if (sqrtba_pose_graph(h, n_vert, est.data(), G.fixed.data(), bFixScale ? 1 : 0, (int)(G.edge_ij.size() / 2), G.edge_ij.data(),
                        G.meas8.data(), 20, 1e-16, nullptr) != SQRTBA_OK) {
    H.err = sqrtba_last_error(h);
    return;
  }
  std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);  // :1456
  // SE3 pose recovery (:1459-1476): Sim3 [sR t; 0 1] -> SE3 [R t/s; 0 1]
  std::vector<double> vCorrectedSwc(est.size(), 0.0);
  for (KeyFrame* pKFi : G.vpKFs) {
    const size_t id = pKFi->mnId;
    if (id >= G.present.size() || !G.present[id]) continue;  // (a bad keyframe has no vertex)
    const double* Siw = &est[id * 8];
    sqrtba::sim3_inv(Siw, &vCorrectedSwc[id * 8]);
    double R[9];
    sqrtba::quat_to_R(Siw, R);  // Quaterniond::toRotationMatrix
    const double inv_s = 1. / Siw[7];
    cv::Mat Tiw(4, 4, CV_32F);  // Converter::toCvSE3
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) Tiw.at<float>(i, j) = (float)R[i * 3 + j];
      Tiw.at<float>(i, 3) = (float)(Siw[4 + i] * inv_s);
    }
    Tiw.at<float>(3, 3) = 1.f;
    pKFi->SetPose(Tiw);
  }
  // map points (:1479-1518): through their reference keyframe, from its pose before to its pose after the optimisation
  const std::vector<MapPoint*> vpMPs = pMap->GetAllMapPoints();
  parallel_ranges(vpMPs.size(), host_threads(vpMPs.size()), [&](int, size_t a, size_t b) {
    for (size_t i = a; i < b; i++) {
      MapPoint* pMP = vpMPs[i];
      if (pMP->isBad()) continue;
      size_t nIDr;
      if (pMP->mnCorrectedByKF == pCurKF->mnId) {
        nIDr = pMP->mnCorrectedReference;
      } else {
        KeyFrame* pRefKF = pMP->GetReferenceKeyFrame();
        nIDr = pRefKF->mnId;
      }
      if (nIDr >= G.present.size() || !G.present[nIDr]) continue;
      const double* Srw = &G.vScw[nIDr * 8];
      const double* Swr = &vCorrectedSwc[nIDr * 8];
      const cv::Mat P3Dw = pMP->GetWorldPos();
      const double X[3] = {P3Dw.at<float>(0), P3Dw.at<float>(1), P3Dw.at<float>(2)};
      double rc[3], c[3], rw[3], w[3];
      sqrtba::sim3_rotate(Srw, X, rc);  // Sim3::map: s * (r * xyz) + t
      for (int k = 0; k < 3; k++) c[k] = Srw[7] * rc[k] + Srw[4 + k];
      sqrtba::sim3_rotate(Swr, c, rw);
      for (int k = 0; k < 3; k++) w[k] = Swr[7] * rw[k] + Swr[4 + k];
      pMP->SetWorldPos(toCvMat3(w));
      pMP->UpdateNormalAndDepth();
    }
  });
}

// ---- Sim3 of a loop candidate (g2oOptimizer::OptimizeSim3, src/backend/g2oOptimizer.cc:1560-1796) --------------------
namespace {
#ifdef SQRTBA_WITH_REFERENCE_HEADERS
g2o::Sim3 sim3_from8(const double* v) {
  return g2o::Sim3(Eigen::Quaterniond(v[3], v[0], v[1], v[2]), Eigen::Vector3d(v[4], v[5], v[6]), v[7]);
}
#else
g2o::Sim3 sim3_from8(const double* v) {
  g2o::Sim3 S;
  S.r.x_ = v[0]; S.r.y_ = v[1]; S.r.z_ = v[2]; S.r.w_ = v[3];
  S.t[0] = v[4]; S.t[1] = v[5]; S.t[2] = v[6];
  S.s = v[7];
  return S;
}
#endif
// R * X + t on CV_32F matrices the way cv::Mat evaluates it (:1650-1651, :1660-1661): one GEMM with the translation as
// its additive term -- double accumulator, ONE rounding to float -- then Converter::toVector3d
void to_camera(const cv::Mat& R, const cv::Mat& t, const cv::Mat& Xw, double* out) {
  for (int i = 0; i < 3; i++) {
    double a = 0.0;
    for (int j = 0; j < 3; j++) a += (double)R.at<float>(i, j) * (double)Xw.at<float>(j);
    out[i] = (double)(float)(a + (double)t.at<float>(i));
  }
}
}  // namespace

void sqrtbaOptimizer::GatherSim3(KeyFrame* pKF1, KeyFrame* pKF2, const std::vector<MapPoint*>& vpMatches1, Sim3Problem& out) {
  const cv::Mat& K1 = pKF1->mK;
  const cv::Mat& K2 = pKF2->mK;
  out.cam8[0] = K1.at<float>(0, 0); out.cam8[1] = K1.at<float>(1, 1); out.cam8[2] = K1.at<float>(0, 2); out.cam8[3] = K1.at<float>(1, 2);
  out.cam8[4] = K2.at<float>(0, 0); out.cam8[5] = K2.at<float>(1, 1); out.cam8[6] = K2.at<float>(0, 2); out.cam8[7] = K2.at<float>(1, 2);
  const cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation();
  const cv::Mat R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
  const int N = (int)vpMatches1.size();
  const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
  out.p1c.clear(); out.p2c.clear(); out.meas6.clear(); out.index.clear();
  for (int i = 0; i < N; i++) {
    if (!vpMatches1[i]) continue;
    MapPoint* pMP1 = vpMapPoints1[i];
    MapPoint* pMP2 = vpMatches1[i];
    const int i2 = pMP2->GetIndexInKeyFrame(pKF2);
    if (!pMP1 || !pMP2) continue;                                  // :1640-1668
    if (pMP1->isBad() || pMP2->isBad() || i2 < 0) continue;
    double c1[3], c2[3];
    to_camera(R1w, t1w, pMP1->GetWorldPos(), c1);
    to_camera(R2w, t2w, pMP2->GetWorldPos(), c2);
    out.p1c.insert(out.p1c.end(), c1, c1 + 3);
    out.p2c.insert(out.p2c.end(), c2, c2 + 3);
    const cv::KeyPoint& kpUn1 = pKF1->mvKeysUn[i];
    const cv::KeyPoint& kpUn2 = pKF2->mvKeysUn[i2];
    const float m[6] = {kpUn1.pt.x, kpUn1.pt.y, pKF1->mvInvLevelSigma2[kpUn1.octave],
                        kpUn2.pt.x, kpUn2.pt.y, pKF2->mvInvLevelSigma2[kpUn2.octave]};
    out.meas6.insert(out.meas6.end(), m, m + 6);
    out.index.push_back((size_t)i);
  }
}

int sqrtbaOptimizer::OptimizeSim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches1, g2o::Sim3& g2oS12,
                                  const float th2, const bool bFixScale) {
  Sim3Problem P;
  GatherSim3(pKF1, pKF2, vpMatches1, P);
  Handle& H = tl_handle;
  sqrtba_handle* h = H.get();
  if (!h) return 0;
  const int64_t ptr[2] = {0, (int64_t)P.index.size()};
  double s12[8];
  sim3_to8(g2oS12, s12);
  std::vector<uint8_t> keep(std::max<size_t>(P.index.size(), 1), 1);
  int32_t nIn = 0;
  if (sqrtba_optimize_sim3(h, 1, ptr, s12, P.cam8, P.p1c.data(), P.p2c.data(), P.meas6.data(), th2, bFixScale ? 1 : 0,
                           keep.data(), &nIn, nullptr) != SQRTBA_OK) {
    H.err = sqrtba_last_error(h);
    return 0;
  }
  for (size_t k = 0; k < P.index.size(); k++)
    if (!keep[k]) vpMatches1[P.index[k]] = static_cast<MapPoint*>(NULL);  // :1733, :1773
  g2oS12 = sim3_from8(s12);  // comes back untouched when the reference returns early (:1755-1756)
  return nIn;
}

// ---- the flat problems the adapters hand to the C ABI, without solving (host-side tests, no GPU needed)
static void to_flat(const Gathered& g, sqrtbaOptimizer::FlatProblem& out) {
  out.pose_qt = g.pose_qt; out.cam = g.cam; out.point_xyz = g.point_xyz; out.pose_fixed = g.pose_fixed;
  out.obs_pose = g.obs_pose; out.obs_point = g.obs_point; out.obs_meas = g.obs_meas;
  out.kf_ids.clear(); out.mp_ids.clear();
  for (KeyFrame* kf : g.kfs) out.kf_ids.push_back(kf->mnId);
  for (MapPoint* mp : g.mps) out.mp_ids.push_back(mp->mnId);
}
// Switch the incremental observation mirror on for a map that already exists: one pass over its points (the last time
// their std::maps are copied), after which the hook lines in MapPoint keep it current (host/map_mirror.h).
void sqrtbaOptimizer::AttachMirror(Map* pMap) {
  sqrtba::MapMirror& mm = sqrtba::MapMirror::Global();
  mm.Enable(false);
  mm.Clear();
  const std::vector<MapPoint*> mps = pMap->GetAllMapPoints();
  parallel_ranges(mps.size(), host_threads(mps.size()), [&](int, size_t a, size_t b) {
    std::vector<sqrtba::MapMirror::Obs> obs;
    for (size_t i = a; i < b; i++) {
      if (!mps[i]) continue;
      const std::map<KeyFrame*, size_t> m = mps[i]->GetObservations();
      obs.clear();
      for (auto& kv : m) obs.push_back(sqrtba::MapMirror::Obs{kv.first, kv.second});
      mm.SetPoint(mps[i], obs);
    }
  });
  mm.Enable(true);
}
void sqrtbaOptimizer::DetachMirror() {
  sqrtba::MapMirror::Global().Enable(false);
  sqrtba::MapMirror::Global().Clear();
}

void sqrtbaOptimizer::GatherLocalWindow(KeyFrame* pKF, FlatProblem& out) {
  Gathered g;
  gather_local_window(pKF, g);
  to_flat(g, out);
}
void sqrtbaOptimizer::ApplyLocalResult(KeyFrame* pKF, Map* pMap, const std::vector<double>& pose_qt,
                                       const std::vector<double>& point_xyz, const std::vector<unsigned char>& outlier) {
  Gathered g;
  gather_local_window(pKF, g);
  if (pose_qt.size() != g.kfs.size() * 7 || point_xyz.size() != g.mps.size() * 3 || outlier.size() != g.obs_ref.size()) return;
  write_back_local(g, pKF->mnId, pose_qt.data(), point_xyz.data(), outlier.data(), pMap);
}
void sqrtbaOptimizer::GatherGlobal(const std::vector<KeyFrame*>& vpKFs, const std::vector<MapPoint*>& vpMP, FlatProblem& out) {
  std::vector<KeyFrame*> kfs;
  std::vector<MapPoint*> mps;
  for (KeyFrame* kf : vpKFs)
    if (!kf->isBad()) kfs.push_back(kf);
  for (MapPoint* mp : vpMP)
    if (!mp->isBad()) mps.push_back(mp);
  Gathered g;
  gather(kfs, mps, [](KeyFrame* kf) { return kf->mnId == 0; }, [](KeyFrame* kf) { return !kf->isBad(); }, g);
  to_flat(g, out);
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

// with the fork's lidar block (g2oOptimizer.cc:560-640): same gather for the visual edges, the frame's flat / sharp
// feature clouds and the tracker's local lidar map go along (pcl point fields x y z, packed to n x 3 floats)
int sqrtbaOptimizer::PoseOptimization(Frame* pFrame, PointICloudPtr local_lidarmap_cloud_ptr,
                                      pcl::KdTreeFLANN<PointI>::Ptr /*kdtree_local_map*/, const lidarConfig* lidarconfig) {
  if (!local_lidarmap_cloud_ptr || !lidarconfig || !(local_lidarmap_cloud_ptr->size() > 100)) return PoseOptimization(pFrame);
  Handle& H = tl_handle;
  if (!H.get()) return 0;
  Frame* F = pFrame;
  double pose[7], cam[5] = {F->fx, F->fy, F->cx, F->cy, F->mbf};
  std::vector<double> xyz;
  std::vector<float> meas;
  std::vector<int> ref;
  {
    std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);  // :433
    toSE3Quat(F->mTcw, pose);
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
      ref.push_back(i);
    }
  }
  auto pack = [](const auto& cloud, std::vector<float>& out) {
    out.resize(cloud.points.size() * 3);
    for (size_t i = 0; i < cloud.points.size(); i++) {
      out[i * 3] = cloud.points[i].x; out[i * 3 + 1] = cloud.points[i].y; out[i * 3 + 2] = cloud.points[i].z;
    }
  };
  std::vector<float> flat, nrm, corner, map;
  pack(F->surface_points_flat_, flat);
  pack(F->surface_points_flat_normal_, nrm);
  pack(F->corner_points_sharp_, corner);
  pack(*local_lidarmap_cloud_ptr, map);
  sqrtba_frame_lidar L;
  std::memset(&L, 0, sizeof L);
  L.n_flat = (int32_t)std::min(flat.size(), nrm.size()) / 3; L.flat_xyz = flat.data(); L.flat_normal = nrm.data();
  L.n_corner = (int32_t)(corner.size() / 3); L.corner_xyz = corner.data();
  L.n_map = (int64_t)(map.size() / 3); L.map_xyz = map.data();
  L.distance_sq_threshold = lidarconfig->distance_sq_threshold;
  L.flat_weight = lidarconfig->flat_optimized_weight; L.corner_weight = lidarconfig->corner_optimized_weight;
  L.use_flat = lidarconfig->using_flat_point ? 1 : 0; L.use_corner = lidarconfig->using_sharp_point ? 1 : 0;
  std::vector<uint8_t> out(std::max<size_t>(ref.size(), 1), 0);
  int32_t inl = 0;
  if (sqrtba_pose_opt_lidar(H.h, pose, cam, (int32_t)ref.size(), xyz.data(), meas.data(), out.data(), &inl, &L, nullptr, nullptr) !=
      SQRTBA_OK) {
    H.err = sqrtba_last_error(H.h);
    return 0;
  }
  for (size_t k = 0; k < ref.size(); k++) F->mvbOutlier[ref[k]] = out[k] != 0;
  if (ref.size() >= 3) F->SetPose(toCvMat(pose));  // :491-492: nothing happens below 3
  return inl;
}

void Optimizer::GlobalBundleAdjustemnt(Map* pMap, int nIterations, bool* pbStopFlag, const unsigned long nLoopKF,
                                       const bool bRobust) {
  sqrtbaOptimizer::GlobalBundleAdjustemnt(pMap, nIterations, pbStopFlag, nLoopKF, bRobust);
}
void Optimizer::BundleAdjustment(const std::vector<KeyFrame*>& vpKFs, const std::vector<MapPoint*>& vpMP, int nIterations,
                                 bool* pbStopFlag, const unsigned long nLoopKF, const bool bRobust) {
  sqrtbaOptimizer::BundleAdjustment(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust);
}
void Optimizer::OptimizeEssentialGraph(Map* pMap, KeyFrame* pLoopKF, KeyFrame* pCurKF,
                                       const LoopClosing::KeyFrameAndPose& NonCorrectedSim3,
                                       const LoopClosing::KeyFrameAndPose& CorrectedSim3,
                                       const std::map<KeyFrame*, std::set<KeyFrame*>>& LoopConnections, const bool& bFixScale) {
  sqrtbaOptimizer::OptimizeEssentialGraph(pMap, pLoopKF, pCurKF, NonCorrectedSim3, CorrectedSim3, LoopConnections, bFixScale);
}
void Optimizer::LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, const lidarConfig* lidarconfig) {
  sqrtbaOptimizer::LocalBundleAdjustment(pKF, pbStopFlag, pMap, lidarconfig);
}
int Optimizer::PoseOptimization(Frame* pFrame, PointICloudPtr local_lidarmap_cloud_ptr,
                                pcl::KdTreeFLANN<PointI>::Ptr kdtree_local_map, const lidarConfig* lidarconfig) {
  return sqrtbaOptimizer::PoseOptimization(pFrame, local_lidarmap_cloud_ptr, kdtree_local_map, lidarconfig);
}
int Optimizer::OptimizeSim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches1, g2o::Sim3& g2oS12,
                            const float th2, const bool bFixScale) {
  return sqrtbaOptimizer::OptimizeSim3(pKF1, pKF2, vpMatches1, g2oS12, th2, bFixScale);
}

}  // namespace ORB_SLAM2

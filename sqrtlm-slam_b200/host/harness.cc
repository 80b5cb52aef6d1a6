// C harness around the adapter for the Python tests: builds a map of the header-compatible types from flat arrays,
// calls ORB_SLAM2::Optimizer::{LocalBundleAdjustment,GlobalBundleAdjustemnt} exactly as LocalMapping / LoopClosing do
// (src/backend/LocalMapping.cc:131, src/backend/LoopClosing.cc:987), and exposes the resulting map state.
#include <algorithm>
#include <cstdint>
#include <list>
#include <memory>

#include "Optimizer.h"
#include <atomic>
#include <chrono>
#include <thread>

using namespace ORB_SLAM2;

struct hh_map {
  Map map;
  std::vector<std::unique_ptr<KeyFrame>> kfs;
  std::vector<std::unique_ptr<MapPoint>> mps;
  lidarConfig lidar;
};

extern "C" {

hh_map* hh_build(int n_kf, const float* Tcw, const float* cam5, const float* inv_sigma2, int n_levels, int n_mp,
                 const float* Xw, int n_obs, const int32_t* obs_kf, const int32_t* obs_mp, const float* obs_uvr,
                 const int32_t* obs_octave) {
  hh_map* m = new hh_map();
  for (int i = 0; i < n_kf; i++) {
    auto kf = std::make_unique<KeyFrame>();
    kf->mnId = (unsigned long)i;
    kf->Tcw.create(4, 4, CV_32F);
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) kf->Tcw.at<float>(r, c) = Tcw[i * 16 + r * 4 + c];
    kf->fx = cam5[0]; kf->fy = cam5[1]; kf->cx = cam5[2]; kf->cy = cam5[3]; kf->mbf = cam5[4];
    kf->mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + n_levels);
    kf->mK.create(3, 3, CV_32F);
    kf->mK.at<float>(0, 0) = cam5[0]; kf->mK.at<float>(1, 1) = cam5[1]; kf->mK.at<float>(0, 2) = cam5[2];
    kf->mK.at<float>(1, 2) = cam5[3]; kf->mK.at<float>(2, 2) = 1.f;
    m->map.mspKeyFrames.push_back(kf.get());
    m->kfs.push_back(std::move(kf));
  }
  for (int i = 0; i < n_mp; i++) {
    auto mp = std::make_unique<MapPoint>();
    mp->mnId = (unsigned long)i;
    mp->mWorldPos.create(3, 1, CV_32F);
    for (int r = 0; r < 3; r++) mp->mWorldPos.at<float>(r) = Xw[i * 3 + r];
    m->map.mspMapPoints.push_back(mp.get());
    m->mps.push_back(std::move(mp));
  }
  for (int k = 0; k < n_obs; k++) {
    KeyFrame* kf = m->kfs[obs_kf[k]].get();
    MapPoint* mp = m->mps[obs_mp[k]].get();
    cv::KeyPoint kp;
    kp.pt.x = obs_uvr[k * 3 + 0];
    kp.pt.y = obs_uvr[k * 3 + 1];
    kp.octave = obs_octave[k];
    const size_t idx = kf->mvKeysUn.size();
    kf->mvKeysUn.push_back(kp);
    kf->mvuRight.push_back(obs_uvr[k * 3 + 2]);
    kf->mvpMapPoints.push_back(mp);
    mp->mObservations[kf] = idx;
  }
  return m;
}

void hh_destroy(hh_map* m) { delete m; }

void hh_set_covisible(hh_map* m, int kf, const int32_t* list, int n) {
  auto& v = m->kfs[kf]->mvpOrderedConnectedKeyFrames;
  v.clear();
  for (int i = 0; i < n; i++) v.push_back(m->kfs[list[i]].get());
}

void hh_set_bad(hh_map* m, int kf, int mp) {
  if (kf >= 0) m->kfs[kf]->mbBad = true;
  if (mp >= 0) m->mps[mp]->mbBad = true;
}

// lidar features of one keyframe (own frame) and the lidarConfig fields of the lidar pass
void hh_set_lidar_cloud(hh_map* m, int kf, int n_flat, const float* flat_xyz, const float* flat_normal, int n_corner,
                        const float* corner_xyz) {
  KeyFrame* k = m->kfs[kf].get();
  auto fill = [](PointIRTCloud& c, int n, const float* xyz) {
    c.points.resize((size_t)n);
    for (int i = 0; i < n; i++) { c.points[i].x = xyz[i * 3]; c.points[i].y = xyz[i * 3 + 1]; c.points[i].z = xyz[i * 3 + 2]; }
  };
  fill(k->surface_points_less_flat_, n_flat, flat_xyz);
  fill(k->surface_points_less_flat_normal_, flat_normal ? n_flat : 0, flat_normal);
  fill(k->corner_points_less_sharp_, n_corner, corner_xyz);
}
void hh_set_lidar_config(hh_map* m, int use_flat, int use_corner, double thr, double w_flat, double w_corner) {
  m->lidar.using_flat_point = use_flat != 0;
  m->lidar.using_sharp_point = use_corner != 0;
  m->lidar.distance_sq_threshold = thr;
  m->lidar.flat_optimized_weight = w_flat;
  m->lidar.corner_optimized_weight = w_corner;
}

// the flat problem of the adapter's gather (no solve, no GPU): sizes first, then the arrays
static sqrtbaOptimizer::FlatProblem g_flat;
void hh_gather(hh_map* m, int kf /* >= 0: local window of that keyframe, -1: whole map */, int32_t* sizes3) {
  if (kf >= 0) sqrtbaOptimizer::GatherLocalWindow(m->kfs[kf].get(), g_flat);
  else sqrtbaOptimizer::GatherGlobal(m->map.GetAllKeyFrames(), m->map.GetAllMapPoints(), g_flat);
  sizes3[0] = (int32_t)g_flat.kf_ids.size(); sizes3[1] = (int32_t)g_flat.mp_ids.size(); sizes3[2] = (int32_t)g_flat.obs_pose.size();
}
void hh_gather_get(double* pose_qt, uint8_t* pose_fixed, double* cam, double* point_xyz, int32_t* obs_pose, int32_t* obs_point,
                   float* obs_meas, int64_t* kf_ids, int64_t* mp_ids) {
  const auto& f = g_flat;
  std::copy(f.pose_qt.begin(), f.pose_qt.end(), pose_qt);
  std::copy(f.pose_fixed.begin(), f.pose_fixed.end(), pose_fixed);
  std::copy(f.cam.begin(), f.cam.end(), cam);
  std::copy(f.point_xyz.begin(), f.point_xyz.end(), point_xyz);
  std::copy(f.obs_pose.begin(), f.obs_pose.end(), obs_pose);
  std::copy(f.obs_point.begin(), f.obs_point.end(), obs_point);
  std::copy(f.obs_meas.begin(), f.obs_meas.end(), obs_meas);
  for (size_t i = 0; i < f.kf_ids.size(); i++) kf_ids[i] = (int64_t)f.kf_ids[i];
  for (size_t i = 0; i < f.mp_ids.size(); i++) mp_ids[i] = (int64_t)f.mp_ids[i];
}

void hh_apply_local(hh_map* m, int kf, int n_kf, const double* pose_qt, int n_mp, const double* point_xyz, int n_obs,
                    const uint8_t* outlier) {
  sqrtbaOptimizer::ApplyLocalResult(m->kfs[kf].get(), &m->map, std::vector<double>(pose_qt, pose_qt + (size_t)n_kf * 7),
                                    std::vector<double>(point_xyz, point_xyz + (size_t)n_mp * 3),
                                    std::vector<unsigned char>(outlier, outlier + n_obs));
}

// ---- essential graph: spanning tree, loop edges, covisibility weights, reference keyframes; then the reference call
void hh_set_parent(hh_map* m, int kf, int parent) {
  m->kfs[kf]->mpParent = m->kfs[parent].get();
  m->kfs[parent]->mspChildrens.insert(m->kfs[kf].get());
}
void hh_add_loop_edge(hh_map* m, int a, int b) {
  m->kfs[a]->mspLoopEdges.insert(m->kfs[b].get());
  m->kfs[b]->mspLoopEdges.insert(m->kfs[a].get());
}
void hh_set_weight(hh_map* m, int a, int b, int w) {
  m->kfs[a]->mConnectedKeyFrameWeights[m->kfs[b].get()] = w;
  m->kfs[b]->mConnectedKeyFrameWeights[m->kfs[a].get()] = w;
}
void hh_set_ref_kf(hh_map* m, int mp, int kf, long corrected_by, long corrected_ref) {
  m->mps[mp]->mpRefKF = m->kfs[kf].get();
  if (corrected_by >= 0) { m->mps[mp]->mnCorrectedByKF = (unsigned long)corrected_by; m->mps[mp]->mnCorrectedReference = (unsigned long)corrected_ref; }
}
static g2o::Sim3 sim3_from8(const double* v) {
  g2o::Sim3 S;
  S.r.x_ = v[0]; S.r.y_ = v[1]; S.r.z_ = v[2]; S.r.w_ = v[3];
  S.t[0] = v[4]; S.t[1] = v[5]; S.t[2] = v[6];
  S.s = v[7];
  return S;
}
// mode 0: gather only (sizes3 <- n_vert, n_edge, -), mode 1: Optimizer::OptimizeEssentialGraph
static sqrtbaOptimizer::PoseGraphProblem g_pg;
void hh_essential_graph(hh_map* m, int mode, int loop_kf, int cur_kf, int n_corr, const int32_t* corr_kf, const double* corr8,
                        const double* noncorr8, int n_conn, const int32_t* conn_ab, int fix_scale, int32_t* sizes3) {
  LoopClosing::KeyFrameAndPose Corrected, NonCorrected;
  for (int k = 0; k < n_corr; k++) {
    Corrected[m->kfs[corr_kf[k]].get()] = sim3_from8(corr8 + (size_t)k * 8);
    NonCorrected[m->kfs[corr_kf[k]].get()] = sim3_from8(noncorr8 + (size_t)k * 8);
  }
  std::map<KeyFrame*, std::set<KeyFrame*>> Conn;
  for (int k = 0; k < n_conn; k++) Conn[m->kfs[conn_ab[2 * k]].get()].insert(m->kfs[conn_ab[2 * k + 1]].get());
  if (mode == 0) {
    sqrtbaOptimizer::GatherEssentialGraph(&m->map, m->kfs[loop_kf].get(), m->kfs[cur_kf].get(), NonCorrected, Corrected, Conn, g_pg);
    sizes3[0] = (int)g_pg.fixed.size(); sizes3[1] = (int)(g_pg.edge_ij.size() / 2); sizes3[2] = 0;
  } else {
    const bool fs = fix_scale != 0;
    Optimizer::OptimizeEssentialGraph(&m->map, m->kfs[loop_kf].get(), m->kfs[cur_kf].get(), NonCorrected, Corrected, Conn, fs);
  }
}
void hh_essential_graph_get(double* vert8, uint8_t* fixed, uint8_t* present, int32_t* edge_ij, double* meas8) {
  std::copy(g_pg.vert8.begin(), g_pg.vert8.end(), vert8);
  std::copy(g_pg.fixed.begin(), g_pg.fixed.end(), fixed);
  std::copy(g_pg.present.begin(), g_pg.present.end(), present);
  std::copy(g_pg.edge_ij.begin(), g_pg.edge_ij.end(), edge_ij);
  std::copy(g_pg.meas8.begin(), g_pg.meas8.end(), meas8);
}
// ---- what LoopClosing::RunGlobalBundleAdjustment does with the staged global-BA result (src/backend/LoopClosing.cc:
// 1014-1103), restated as the CALLER's code for the integration test: keyframes and points that did not take part in
// the optimisation (created while it ran) follow their parent / reference keyframe.  Float 4x4 arithmetic like cv::Mat.
static cv::Mat mul4(const cv::Mat& A, const cv::Mat& B) {
  cv::Mat Cm(4, 4, CV_32F);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float acc = 0.f;
      for (int k = 0; k < 4; k++) acc += A.at<float>(i, k) * B.at<float>(k, j);
      Cm.at<float>(i, j) = acc;
    }
  return Cm;
}
void hh_forget_gba(hh_map* m, int kf, int mp) {  // as if created after the global BA started
  if (kf >= 0) m->kfs[kf]->mnBAGlobalForKF = 0;
  if (mp >= 0) m->mps[mp]->mnBAGlobalForKF = 0;
}
void hh_apply_gba(hh_map* m, unsigned long nLoopKF) {
  std::unique_lock<std::mutex> lock(m->map.mMutexMapUpdate);
  std::list<KeyFrame*> lpKFtoCheck(m->map.mvpKeyFrameOrigins.begin(), m->map.mvpKeyFrameOrigins.end());
  while (!lpKFtoCheck.empty()) {
    KeyFrame* pKF = lpKFtoCheck.front();
    const std::set<KeyFrame*> sChilds = pKF->GetChilds();
    cv::Mat Twc = pKF->GetPoseInverse();
    for (KeyFrame* pChild : sChilds) {
      if (pChild->mnBAGlobalForKF != nLoopKF) {
        cv::Mat Tchildc = mul4(pChild->GetPose(), Twc);
        pChild->mTcwGBA = mul4(Tchildc, pKF->mTcwGBA);
        pChild->mnBAGlobalForKF = nLoopKF;
      }
      lpKFtoCheck.push_back(pChild);
    }
    pKF->mTcwBefGBA = pKF->GetPose();
    pKF->SetPose(pKF->mTcwGBA);
    lpKFtoCheck.pop_front();
  }
  for (MapPoint* pMP : m->map.GetAllMapPoints()) {
    if (pMP->isBad()) continue;
    if (pMP->mnBAGlobalForKF == nLoopKF) {
      pMP->SetWorldPos(pMP->mPosGBA);
    } else {
      KeyFrame* pRefKF = pMP->GetReferenceKeyFrame();
      if (!pRefKF || pRefKF->mnBAGlobalForKF != nLoopKF) continue;
      const cv::Mat& B = pRefKF->mTcwBefGBA;
      const cv::Mat X = pMP->GetWorldPos();
      float Xc[3];
      for (int i = 0; i < 3; i++)
        Xc[i] = B.at<float>(i, 0) * X.at<float>(0) + B.at<float>(i, 1) * X.at<float>(1) + B.at<float>(i, 2) * X.at<float>(2) + B.at<float>(i, 3);
      const cv::Mat Twc = pRefKF->GetPoseInverse();
      cv::Mat W(3, 1, CV_32F);
      for (int i = 0; i < 3; i++)
        W.at<float>(i) = Twc.at<float>(i, 0) * Xc[0] + Twc.at<float>(i, 1) * Xc[1] + Twc.at<float>(i, 2) * Xc[2] + Twc.at<float>(i, 3);
      pMP->SetWorldPos(W);
    }
  }
}
// ---- OptimizeSim3: match_mp[i] = index of the map point matched to keypoint i of kf1 (-1: no match), in/out
static std::vector<MapPoint*> hh_matches(hh_map* m, int n, const int32_t* match_mp) {
  std::vector<MapPoint*> v((size_t)n, nullptr);
  for (int i = 0; i < n; i++)
    if (match_mp[i] >= 0) v[i] = m->mps[match_mp[i]].get();
  return v;
}
void hh_set_K(hh_map* m, int kf, float fx, float fy, float cx, float cy) {
  cv::Mat& K = m->kfs[kf]->mK;
  K.at<float>(0, 0) = fx; K.at<float>(1, 1) = fy; K.at<float>(0, 2) = cx; K.at<float>(1, 2) = cy;
}
int hh_gather_sim3(hh_map* m, int kf1, int kf2, int n, const int32_t* match_mp, double* cam8, double* p1c, double* p2c,
                   float* meas6, int32_t* index) {
  sqrtbaOptimizer::Sim3Problem P;
  sqrtbaOptimizer::GatherSim3(m->kfs[kf1].get(), m->kfs[kf2].get(), hh_matches(m, n, match_mp), P);
  std::copy(P.cam8, P.cam8 + 8, cam8);
  std::copy(P.p1c.begin(), P.p1c.end(), p1c);
  std::copy(P.p2c.begin(), P.p2c.end(), p2c);
  std::copy(P.meas6.begin(), P.meas6.end(), meas6);
  for (size_t k = 0; k < P.index.size(); k++) index[k] = (int32_t)P.index[k];
  return (int)P.index.size();
}
int hh_optimize_sim3(hh_map* m, int kf1, int kf2, int n, int32_t* match_mp, double* s12, float th2, int fix_scale) {
  std::vector<MapPoint*> v = hh_matches(m, n, match_mp);
  g2o::Sim3 S;
  S.r.x_ = s12[0]; S.r.y_ = s12[1]; S.r.z_ = s12[2]; S.r.w_ = s12[3];
  S.t[0] = s12[4]; S.t[1] = s12[5]; S.t[2] = s12[6];
  S.s = s12[7];
  const int nIn = Optimizer::OptimizeSim3(m->kfs[kf1].get(), m->kfs[kf2].get(), v, S, th2, fix_scale != 0);
  for (int i = 0; i < n; i++)
    if (!v[i]) match_mp[i] = -1;
  s12[0] = S.rotation().x(); s12[1] = S.rotation().y(); s12[2] = S.rotation().z(); s12[3] = S.rotation().w();
  for (int i = 0; i < 3; i++) s12[4 + i] = S.translation()[i];
  s12[7] = S.scale();
  return nIn;
}

void hh_set_origin(hh_map* m, int kf) { m->map.mvpKeyFrameOrigins.push_back(m->kfs[kf].get()); }
void hh_set_options(int local_ba_stereo_edges, int local_ba_two_pass) {
  sqrtbaOptimizer::options().local_ba_stereo_edges = local_ba_stereo_edges != 0;
  sqrtbaOptimizer::options().local_ba_two_pass = local_ba_two_pass != 0;
}
void hh_local_ba(hh_map* m, int kf, bool* stop) {
  Optimizer::LocalBundleAdjustment(m->kfs[kf].get(), stop, &m->map, &m->lidar);
}

void hh_global_ba(hh_map* m, int iters, int robust, unsigned long nLoopKF, bool* stop) {
  Optimizer::GlobalBundleAdjustemnt(&m->map, iters, stop, nLoopKF, robust != 0);
}

void hh_get_pose(hh_map* m, int kf, int gba, float* out16) {
  const cv::Mat T = gba ? m->kfs[kf]->mTcwGBA : m->kfs[kf]->GetPose();
  for (int i = 0; i < 16; i++) out16[i] = T.empty() ? 0.f : T.at<float>(i / 4, i % 4);
}
void hh_get_point(hh_map* m, int mp, int gba, float* out3) {
  const cv::Mat P = gba ? m->mps[mp]->mPosGBA : m->mps[mp]->GetWorldPos();
  for (int i = 0; i < 3; i++) out3[i] = P.empty() ? 0.f : P.at<float>(i);
}
int hh_has_observation(hh_map* m, int kf, int mp) {
  auto obs = m->mps[mp]->GetObservations();
  return obs.count(m->kfs[kf].get()) ? 1 : 0;
}
int hh_keyframe_sees(hh_map* m, int kf, int mp) {
  for (MapPoint* p : m->kfs[kf]->GetMapPointMatches())
    if (p == m->mps[mp].get()) return 1;
  return 0;
}
int hh_point_updates(hh_map* m, int mp) { return m->mps[mp]->nUpdateNormalAndDepth; }
unsigned long hh_gba_marker(hh_map* m, int kf) { return m->kfs[kf]->mnBAGlobalForKF; }
const char* hh_last_error() { return sqrtbaOptimizer::LastError(); }

// ---- pose-only optimisation: a Frame with n keypoints, every one matched to its own map point
struct hh_frame {
  Frame frame;
  std::vector<std::unique_ptr<MapPoint>> mps;
};
hh_frame* hh_frame_build(const float* Tcw16, const float* cam5, const float* inv_sigma2, int n_levels, int n,
                         const float* Xw, const float* uvr, const int32_t* octave) {
  hh_frame* f = new hh_frame();
  Frame& F = f->frame;
  F.N = n;
  F.mTcw.create(4, 4, CV_32F);
  for (int i = 0; i < 16; i++) F.mTcw.at<float>(i / 4, i % 4) = Tcw16[i];
  F.fx = cam5[0]; F.fy = cam5[1]; F.cx = cam5[2]; F.cy = cam5[3]; F.mbf = cam5[4];
  F.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + n_levels);
  F.mvKeysUn.resize(n); F.mvuRight.resize(n); F.mvpMapPoints.resize(n); F.mvbOutlier.assign(n, false);
  for (int i = 0; i < n; i++) {
    auto mp = std::make_unique<MapPoint>();
    mp->mWorldPos.create(3, 1, CV_32F);
    for (int c = 0; c < 3; c++) mp->mWorldPos.at<float>(c) = Xw[i * 3 + c];
    F.mvKeysUn[i].pt.x = uvr[i * 3]; F.mvKeysUn[i].pt.y = uvr[i * 3 + 1]; F.mvKeysUn[i].octave = octave[i];
    F.mvuRight[i] = uvr[i * 3 + 2];
    F.mvpMapPoints[i] = mp.get();
    f->mps.push_back(std::move(mp));
  }
  return f;
}
void hh_frame_destroy(hh_frame* f) { delete f; }
void hh_frame_unmatch(hh_frame* f, int i) { f->frame.mvpMapPoints[i] = nullptr; }
int hh_pose_opt(hh_frame* f) { return sqrtbaOptimizer::PoseOptimization(&f->frame); }
// Optimizer::PoseOptimization(pFrame, local_lidarmap_cloud_ptr, kdtree_local_map, lidarconfig) as Tracking calls it
// (Tracking.cc:1347, 1547, 1617)
int hh_pose_opt_lidar(hh_frame* f, int n_flat, const float* flat_xyz, const float* flat_normal, int n_corner,
                      const float* corner_xyz, int n_map, const float* map_xyz, int use_flat, int use_corner, double thr,
                      double w_flat, double w_corner) {
  auto fill = [](PointIRTCloud& c, int n, const float* xyz) {
    c.points.resize((size_t)n);
    for (int i = 0; i < n; i++) { c.points[i].x = xyz[i * 3]; c.points[i].y = xyz[i * 3 + 1]; c.points[i].z = xyz[i * 3 + 2]; }
  };
  fill(f->frame.surface_points_flat_, n_flat, flat_xyz);
  fill(f->frame.surface_points_flat_normal_, n_flat, flat_normal);
  fill(f->frame.corner_points_sharp_, n_corner, corner_xyz);
  PointICloudPtr map = std::make_shared<PointICloud>();
  map->points.resize((size_t)n_map);
  for (int i = 0; i < n_map; i++) { map->points[i].x = map_xyz[i * 3]; map->points[i].y = map_xyz[i * 3 + 1]; map->points[i].z = map_xyz[i * 3 + 2]; }
  lidarConfig cfg;
  cfg.using_flat_point = use_flat != 0; cfg.using_sharp_point = use_corner != 0;
  cfg.distance_sq_threshold = thr; cfg.flat_optimized_weight = w_flat; cfg.corner_optimized_weight = w_corner;
  return Optimizer::PoseOptimization(&f->frame, map, pcl::KdTreeFLANN<PointI>::Ptr(), &cfg);
}
void hh_pose_opt_batch(hh_frame** fs, int n, int32_t* inliers) {
  std::vector<Frame*> v;
  for (int i = 0; i < n; i++) v.push_back(&fs[i]->frame);
  std::vector<int> inl;
  sqrtbaOptimizer::PoseOptimizationBatch(v, inl);
  for (int i = 0; i < n; i++) inliers[i] = inl[i];
}
void hh_frame_get(hh_frame* f, float* Tcw16, uint8_t* outlier) {
  for (int i = 0; i < 16; i++) Tcw16[i] = f->frame.mTcw.at<float>(i / 4, i % 4);
  for (int i = 0; i < f->frame.N; i++) outlier[i] = f->frame.mvbOutlier[i] ? 1 : 0;
}


// ---- incremental observation mirror (host/map_mirror.h): attach / detach, mutations through the map's own methods (which
// carry the hook lines), consistency of the mirror with the map, a concurrent stress run and the cost of the two paths
void hh_mirror_attach(hh_map* m) { sqrtbaOptimizer::AttachMirror(&m->map); }
void hh_mirror_detach() { sqrtbaOptimizer::DetachMirror(); }
int hh_mirror_points() { return (int)sqrtba::MapMirror::Global().NumPoints(); }
// a new keypoint of keyframe kf observing map point mp (MapPoint::AddObservation + KeyFrame::AddMapPoint)
void hh_add_observation(hh_map* m, int kf, int mp, float u, float v, float ur, int octave) {
  KeyFrame* k = m->kfs[kf].get();
  MapPoint* p = m->mps[mp].get();
  cv::KeyPoint kp;
  kp.pt.x = u; kp.pt.y = v; kp.octave = octave;
  const size_t idx = k->mvKeysUn.size();
  k->mvKeysUn.push_back(kp);
  k->mvuRight.push_back(ur);
  k->mvpMapPoints.push_back(p);
  p->AddObservation(k, idx);
}
void hh_erase_observation(hh_map* m, int kf, int mp) {
  m->kfs[kf]->EraseMapPointMatch(m->mps[mp].get());
  m->mps[mp]->EraseObservation(m->kfs[kf].get());
}
void hh_set_point_bad(hh_map* m, int mp) { m->mps[mp]->SetBadFlag(); }
// number of map points whose mirrored list differs from MapPoint::GetObservations() (order included); -1: unknown point
int hh_mirror_mismatches(hh_map* m) {
  std::vector<MapPoint*> mps = m->map.GetAllMapPoints();
  std::vector<size_t> ptr;
  std::vector<sqrtba::MapMirror::Obs> rec;
  if (!sqrtba::MapMirror::Global().Snapshot(mps.data(), mps.size(), ptr, rec)) return -1;
  int bad = 0;
  for (size_t i = 0; i < mps.size(); i++) {
    const std::map<KeyFrame*, size_t> obs = mps[i]->GetObservations();
    bool same = obs.size() == ptr[i + 1] - ptr[i];
    size_t k = ptr[i];
    if (same)
      for (auto& kv : obs) {
        if (rec[k].kf != (const void*)kv.first || rec[k].idx != kv.second) { same = false; break; }
        k++;
      }
    bad += same ? 0 : 1;
  }
  return bad;
}
// writers add / erase observations of disjoint slices of the map points through the hooked methods while one reader
// keeps taking snapshots; returns the number of snapshots taken (the caller checks hh_mirror_mismatches afterwards)
int hh_mirror_stress(hh_map* m, int n_writers, int rounds) {
  std::atomic<bool> done(false);
  std::atomic<int> snaps(0);
  std::vector<MapPoint*> mps = m->map.GetAllMapPoints();
  std::thread reader([&] {
    std::vector<size_t> ptr;
    std::vector<sqrtba::MapMirror::Obs> rec;
    while (!done.load()) {
      if (sqrtba::MapMirror::Global().Snapshot(mps.data(), mps.size(), ptr, rec)) snaps++;
    }
  });
  while (snaps.load() == 0) std::this_thread::yield();  // the reader is running before the writers start
  std::vector<std::thread> ws;
  const int nk = (int)m->kfs.size();
  for (int w = 0; w < n_writers; w++)
    ws.emplace_back([&, w] {
      unsigned rng = 12345u + 977u * (unsigned)w;
      for (int r = 0; r < rounds; r++)
        for (size_t i = (size_t)w; i < mps.size(); i += (size_t)n_writers) {
          rng = rng * 1664525u + 1013904223u;
          KeyFrame* k = m->kfs[(rng >> 8) % nk].get();
          if ((rng >> 4) & 1) mps[i]->AddObservation(k, (size_t)(rng >> 20) % std::max<size_t>(1, k->mvKeysUn.size()));
          else mps[i]->EraseObservation(k);
        }
    });
  for (auto& t : ws) t.join();
  done.store(true);
  reader.join();
  return snaps.load();
}
// the window-selection markers are per-call stamps (mnBALocalForKF == current keyframe id means "already taken",
// g2oOptimizer.cc:713-775): selecting the same window twice needs them cleared in between
void hh_reset_markers(hh_map* m) {
  for (auto& k : m->kfs) { k->mnBALocalForKF = ~0ul; k->mnBAFixedForKF = ~0ul; }
  for (auto& p : m->mps) p->mnBALocalForKF = ~0ul;
}
// microseconds per gather of the local window of keyframe kf (best of reps), with whatever path is active
double hh_time_gather(hh_map* m, int kf, int reps) {
  double best = 1e30;
  sqrtbaOptimizer::FlatProblem f;
  for (int r = 0; r < reps; r++) {
    hh_reset_markers(m);
    const auto t0 = std::chrono::steady_clock::now();
    sqrtbaOptimizer::GatherLocalWindow(m->kfs[kf].get(), f);
    best = std::min(best, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
  }
  return best;
}

}  // extern "C"

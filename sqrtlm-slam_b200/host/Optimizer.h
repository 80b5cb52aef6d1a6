// Mirror of the bundle-adjustment part of the reference facade (include/backend/Optimizer.h:42-56): same class, same
// static signatures (the reference's spelling `GlobalBundleAdjustemnt` included), plus the new back-end selector value.
#pragma once
#include <map>
#include <set>
#include <vector>

#ifdef SQRTBA_WITH_REFERENCE_HEADERS
#include "data_structure/Frame.h"
#include "data_structure/KeyFrame.h"
#include "data_structure/Map.h"
#include "data_structure/MapPoint.h"
#include "utils/lidarconfig.h"
#else
#include "map_types.h"
#endif

namespace ORB_SLAM2 {

class Optimizer {
 public:
  enum eSolver { CERES = 0, G2O = 1, MYOPT = 2, SQRTBA = 3 };
  void static BundleAdjustment(const std::vector<KeyFrame*>& vpKF, const std::vector<MapPoint*>& vpMP,
                               int nIterations = 5, bool* pbStopFlag = NULL, const unsigned long nLoopKF = 0,
                               const bool bRobust = true);
  void static GlobalBundleAdjustemnt(Map* pMap, int nIterations = 5, bool* pbStopFlag = NULL,
                                     const unsigned long nLoopKF = 0, const bool bRobust = true);
  void static LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, const lidarConfig* lidarconfig);
  // include/backend/Optimizer.h:58-59
  int static PoseOptimization(Frame* pFrame, PointICloudPtr local_lidarmap_cloud_ptr,
                              pcl::KdTreeFLANN<PointI>::Ptr kdtree_local_map, const lidarConfig* lidarconfig);
  // include/backend/Optimizer.h:62-67
  void static OptimizeEssentialGraph(Map* pMap, KeyFrame* pLoopKF, KeyFrame* pCurKF,
                                     const LoopClosing::KeyFrameAndPose& NonCorrectedSim3,
                                     const LoopClosing::KeyFrameAndPose& CorrectedSim3,
                                     const std::map<KeyFrame*, std::set<KeyFrame*>>& LoopConnections, const bool& bFixScale);
  // include/backend/Optimizer.h:68-69
  static int OptimizeSim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches1, g2o::Sim3& g2oS12,
                          const float th2, const bool bFixScale);
};

// the adapter that sits beside g2oOptimizer / CeresOptimizer / MyOptimizer (src/backend/)
class sqrtbaOptimizer {
 public:
  void static BundleAdjustment(const std::vector<KeyFrame*>& vpKF, const std::vector<MapPoint*>& vpMP, int nIterations,
                               bool* pbStopFlag, const unsigned long nLoopKF, const bool bRobust);
  void static GlobalBundleAdjustemnt(Map* pMap, int nIterations, bool* pbStopFlag, const unsigned long nLoopKF,
                                     const bool bRobust);
  void static LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, const lidarConfig* lidarconfig);
  // the visual part of Optimizer::PoseOptimization (include/backend/Optimizer.h:58-59), same shape as the reference's
  // CeresOptimizer::PoseOptimization(Frame*) / MyOptimizer::PoseOptimization(Frame*) adapters (Optimizer.cc:55-60)
  int static PoseOptimization(Frame* pFrame);
  // g2oOptimizer::PoseOptimization with the fork's lidar block (src/backend/g2oOptimizer.cc:385-690): the frame's flat /
  // sharp points against the tracker's local lidar map; the kd-tree argument is unused (exact search on the device)
  int static PoseOptimization(Frame* pFrame, PointICloudPtr local_lidarmap_cloud_ptr,
                              pcl::KdTreeFLANN<PointI>::Ptr kdtree_local_map, const lidarConfig* lidarconfig);
  // relocalisation: several candidate frames in one launch (Tracking.cc:2466-2517 calls PoseOptimization per candidate)
  void static PoseOptimizationBatch(const std::vector<Frame*>& frames, std::vector<int>& inliers);
  // g2oOptimizer::OptimizeEssentialGraph (src/backend/g2oOptimizer.cc:1212-1520): the graph is built from the map by the
  // reference's rules (keyframe vertices, new loop connections, spanning tree, loop edges, covisibility >= 100), the
  // optimisation runs on the device (sqrtba_pose_graph), poses and map points are corrected under mMutexMapUpdate
  void static OptimizeEssentialGraph(Map* pMap, KeyFrame* pLoopKF, KeyFrame* pCurKF,
                                     const LoopClosing::KeyFrameAndPose& NonCorrectedSim3,
                                     const LoopClosing::KeyFrameAndPose& CorrectedSim3,
                                     const std::map<KeyFrame*, std::set<KeyFrame*>>& LoopConnections, const bool& bFixScale);
  // the graph that OptimizeEssentialGraph hands to sqrtba_pose_graph, without optimising (host-side tests, no GPU):
  // vertices indexed by keyframe mnId (8 doubles each, absent / bad keyframes are zero rows with present = 0)
  struct PoseGraphProblem {
    std::vector<double> vert8, meas8;
    std::vector<unsigned char> fixed, present;
    std::vector<int> edge_ij;
  };
  void static GatherEssentialGraph(Map* pMap, KeyFrame* pLoopKF, KeyFrame* pCurKF,
                                   const LoopClosing::KeyFrameAndPose& NonCorrectedSim3,
                                   const LoopClosing::KeyFrameAndPose& CorrectedSim3,
                                   const std::map<KeyFrame*, std::set<KeyFrame*>>& LoopConnections, PoseGraphProblem& out);
  // g2oOptimizer::OptimizeSim3 (src/backend/g2oOptimizer.cc:1560-1796): the matches are gathered by the reference's
  // rules (both map points good, the second one observed in pKF2; points moved into their keyframe's frame in float),
  // the optimisation runs on the device (sqrtba_optimize_sim3), dropped matches are set to NULL in vpMatches1
  static int OptimizeSim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches1, g2o::Sim3& g2oS12,
                          const float th2, const bool bFixScale);
  // the arrays OptimizeSim3 hands to sqrtba_optimize_sim3, without optimising (host-side tests, no GPU)
  struct Sim3Problem {
    double cam8[8];
    std::vector<double> p1c, p2c;
    std::vector<float> meas6;
    std::vector<size_t> index;  // vnIndexEdge: position in vpMatches1 of every match
  };
  void static GatherSim3(KeyFrame* pKF1, KeyFrame* pKF2, const std::vector<MapPoint*>& vpMatches1, Sim3Problem& out);
  // Behaviour switches of the local-BA adapter.  The defaults are what THIS reference does:
  //  * local_ba_stereo_edges = false: the fork's local BA only creates monocular edges -- an observation with a right
  //    coordinate falls into an empty branch (g2oOptimizer.cc:914-916) and takes no part in the optimisation or in the
  //    outlier erasure.  true = upstream ORB-SLAM2's stereo edges (the solver supports them; thresholds :853).
  //  * local_ba_two_pass = false: the fork ALWAYS runs initializeOptimization(0); optimize(20) after the lidar block,
  //    with or without lidar features or matches (:1113-1114), i.e. the schedule is 5 + 10 + 20.  true = upstream
  //    ORB-SLAM2's two-pass 5 + 10 schedule (what BASELINE.json's configs time).
  struct Options {
    bool local_ba_stereo_edges = false;
    bool local_ba_two_pass = false;
  };
  static Options& options();
  // last error of the calling thread's handle ("" if none); the reference API itself is void / silent
  static const char* LastError();
  // the flat problem (layout of sqrtba_set_problem) that LocalBundleAdjustment / BundleAdjustment build from the map,
  // without solving it: lets the window selection and the gather be tested and timed on a machine without a GPU
  struct FlatProblem {
    std::vector<double> pose_qt, cam, point_xyz;
    std::vector<unsigned char> pose_fixed;
    std::vector<int> obs_pose, obs_point;
    std::vector<float> obs_meas;
    std::vector<unsigned long> kf_ids, mp_ids;
  };
  void static GatherLocalWindow(KeyFrame* pKF, FlatProblem& out);
  // the write-back half of LocalBundleAdjustment (outlier erasure, SetPose, SetWorldPos + UpdateNormalAndDepth under
  // mMutexMapUpdate) applied to a result given in the layout of GatherLocalWindow -- again for tests without a GPU
  void static ApplyLocalResult(KeyFrame* pKF, Map* pMap, const std::vector<double>& pose_qt,
                               const std::vector<double>& point_xyz, const std::vector<unsigned char>& outlier);
  void static GatherGlobal(const std::vector<KeyFrame*>& vpKF, const std::vector<MapPoint*>& vpMP, FlatProblem& out);
  // incremental observation mirror (host/map_mirror.h): attach to an existing map / switch off again
  void static AttachMirror(Map* pMap);
  void static DetachMirror();
};

}  // namespace ORB_SLAM2

// Header-compatible stand-ins for the reference's map types, restricted to the member surface the BA adapters touch
// (SURVEY.md §8(b)): include/data_structure/KeyFrame.h:83-84,134,226-250,306,345-403,427; MapPoint.h:82-88,108,129,
// 161,230,242-287; Map.h:100,107,144; and the sliver of cv::Mat / cv::KeyPoint they use.  They let
// sqrtbaOptimizer.cc compile and be tested in an image without OpenCV / PCL / ROS.  In the reference tree the adapter
// includes the real headers instead -- same names, same signatures.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <vector>

#include "map_mirror.h"   // the hook lines below are the ones INTEGRATION.md adds to the reference's MapPoint.cc

#define CV_32F 5
namespace cv {
struct Point2f { float x = 0, y = 0; };
struct KeyPoint { Point2f pt; int octave = 0; };
class Mat {  // row-major float matrix, value semantics like a cloned cv::Mat
 public:
  Mat() {}
  Mat(int r, int c, int /*type*/) { create(r, c, CV_32F); }
  void create(int r, int c, int /*type*/) { rows = r; cols = c; d_.assign((size_t)r * c, 0.f); }
  template <class T> T& at(int i, int j) { return d_[(size_t)i * cols + j]; }
  template <class T> const T& at(int i, int j) const { return d_[(size_t)i * cols + j]; }
  template <class T> T& at(int i) { return d_[(size_t)i]; }
  template <class T> const T& at(int i) const { return d_[(size_t)i]; }
  Mat clone() const { return *this; }
  void copyTo(Mat& o) const { o = *this; }
  bool empty() const { return d_.empty(); }
  int rows = 0, cols = 0;
 private:
  std::vector<float> d_;
};
}  // namespace cv

struct lidarConfig {  // include/utils/lidarconfig.h:51-56 -- the fields the lidar pass of LocalBundleAdjustment reads
  bool using_flat_point = false, using_sharp_point = false;
  double distance_sq_threshold = 0, flat_optimized_weight = 0, corner_optimized_weight = 0;
};

// include/data_structure/point_types.h:16-23,52-53: PointXYZIRT and pcl::PointCloud<PointIRT>, cut down to what
// g2oOptimizer.cc:994-1096 touches (points[i].x/y/z, size())
struct PointIRT { float x = 0, y = 0, z = 0; };
struct PointIRTCloud {
  std::vector<PointIRT> points;
  size_t size() const { return points.size(); }
};

// include/data_structure/point_types.h:42-45: PointI = pcl::PointXYZI, PointICloud(Ptr) -- the tracker's local lidar map
// handed to PoseOptimization (Tracking.cc:1347) -- and the kd-tree over it (pcl::KdTreeFLANN<PointI>::Ptr; the sqrtba
// adapter searches on the device and only takes the pointer to keep the reference's signature)
struct PointI { float x = 0, y = 0, z = 0, intensity = 0; };
struct PointICloud {
  std::vector<PointI> points;
  size_t size() const { return points.size(); }
};
typedef std::shared_ptr<PointICloud> PointICloudPtr;
namespace pcl {
template <class T> struct KdTreeFLANN { typedef std::shared_ptr<KdTreeFLANN<T>> Ptr; };
}  // namespace pcl

// Thirdparty/g2o/g2o/types/sim3.h: the accessor surface of g2o::Sim3 that the essential-graph adapter uses
// (rotation().x() .. w(), translation()[i], scale()); the real class stores an Eigen quaternion and vector
namespace g2o {
struct Sim3Quat { double x_ = 0, y_ = 0, z_ = 0, w_ = 1; double x() const { return x_; } double y() const { return y_; }
                  double z() const { return z_; } double w() const { return w_; } };
struct Sim3 {
  Sim3Quat r;
  double t[3] = {0, 0, 0};
  double s = 1.0;
  const Sim3Quat& rotation() const { return r; }
  const double* translation() const { return t; }
  double scale() const { return s; }
};
}  // namespace g2o

namespace ORB_SLAM2 {
class MapPoint;
class Map;

class KeyFrame {
 public:
  long unsigned int mnId = 0;
  long unsigned int mnBALocalForKF = ~0ul, mnBAFixedForKF = ~0ul, mnBAGlobalForKF = 0;
  cv::Mat mTcwGBA;
  float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
  cv::Mat mK;  // KeyFrame.h:434 (const in the reference): 3x3 float calibration matrix, read by OptimizeSim3
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvuRight;
  std::vector<float> mvInvLevelSigma2;

  cv::Mat GetPose() { std::unique_lock<std::mutex> l(mMutexPose); return Tcw.clone(); }
  void SetPose(const cv::Mat& T) { std::unique_lock<std::mutex> l(mMutexPose); T.copyTo(Tcw); }
  bool isBad() { return mbBad; }
  std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return mvpOrderedConnectedKeyFrames; }
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  void EraseMapPointMatch(MapPoint* pMP) {
    for (auto& p : mvpMapPoints)
      if (p == pMP) p = nullptr;
  }
  // spanning tree / loop edges / covisibility weights: what OptimizeEssentialGraph reads (KeyFrame.h:149-210)
  KeyFrame* GetParent() { return mpParent; }
  std::set<KeyFrame*> GetLoopEdges() { return mspLoopEdges; }
  bool hasChild(KeyFrame* pKF) { return mspChildrens.count(pKF) != 0; }
  int GetWeight(KeyFrame* pKF) { auto it = mConnectedKeyFrameWeights.find(pKF); return it == mConnectedKeyFrameWeights.end() ? 0 : it->second; }
  std::vector<KeyFrame*> GetCovisiblesByWeight(const int& w) {  // ordered by decreasing weight (KeyFrame.cc:268-287)
    std::vector<std::pair<int, KeyFrame*>> v;
    for (auto& kv : mConnectedKeyFrameWeights)
      if (kv.second >= w) v.emplace_back(-kv.second, kv.first);
    std::stable_sort(v.begin(), v.end(), [](const std::pair<int, KeyFrame*>& a, const std::pair<int, KeyFrame*>& b) { return a.first < b.first; });
    std::vector<KeyFrame*> out;
    for (auto& e : v) out.push_back(e.second);
    return out;
  }
  cv::Mat GetRotation() { std::unique_lock<std::mutex> l(mMutexPose); cv::Mat R(3, 3, CV_32F);
                          for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R.at<float>(i, j) = Tcw.at<float>(i, j); return R; }
  cv::Mat GetTranslation() { std::unique_lock<std::mutex> l(mMutexPose); cv::Mat t(3, 1, CV_32F);
                             for (int i = 0; i < 3; i++) t.at<float>(i) = Tcw.at<float>(i, 3); return t; }
  std::set<KeyFrame*> GetChilds() { return mspChildrens; }
  cv::Mat GetPoseInverse() {  // Twc = [Rcw^T | -Rcw^T tcw], float as in KeyFrame::SetPose (KeyFrame.cc:102-124)
    std::unique_lock<std::mutex> l(mMutexPose);
    cv::Mat Twc(4, 4, CV_32F);
    for (int i = 0; i < 3; i++) {
      float o = 0.f;
      for (int j = 0; j < 3; j++) { Twc.at<float>(i, j) = Tcw.at<float>(j, i); o += Tcw.at<float>(j, i) * Tcw.at<float>(j, 3); }
      Twc.at<float>(i, 3) = -o;
    }
    Twc.at<float>(3, 3) = 1.f;
    return Twc;
  }
  cv::Mat mTcwBefGBA;  // KeyFrame.h:389
  KeyFrame* mpParent = nullptr;
  std::set<KeyFrame*> mspLoopEdges, mspChildrens;
  std::map<KeyFrame*, int> mConnectedKeyFrameWeights;
  // lidar features of the keyframe in its own frame (KeyFrame.h:437-442)
  PointIRTCloud corner_points_less_sharp_, surface_points_less_flat_, surface_points_less_flat_normal_;
  // test-side construction
  cv::Mat Tcw;
  bool mbBad = false;
  std::vector<KeyFrame*> mvpOrderedConnectedKeyFrames;
  std::vector<MapPoint*> mvpMapPoints;  // indexed by keypoint
  std::mutex mMutexPose;
};

class MapPoint {
 public:
  long unsigned int mnId = 0;
  long unsigned int mnBALocalForKF = ~0ul, mnBAGlobalForKF = 0;
  cv::Mat mPosGBA;
  long unsigned int mnCorrectedByKF = 0, mnCorrectedReference = 0;  // MapPoint.h:282-283 (set by LoopClosing::CorrectLoop)
  KeyFrame* GetReferenceKeyFrame() { return mpRefKF; }
  KeyFrame* mpRefKF = nullptr;
  int nUpdateNormalAndDepth = 0;

  cv::Mat GetWorldPos() { std::unique_lock<std::mutex> l(mMutexPos); return mWorldPos.clone(); }
  void SetWorldPos(const cv::Mat& P) { std::unique_lock<std::mutex> l(mMutexPos); P.copyTo(mWorldPos); }
  std::map<KeyFrame*, size_t> GetObservations() { std::unique_lock<std::mutex> l(mMutexFeatures); return mObservations; }
  void AddObservation(KeyFrame* pKF, size_t idx) {  // MapPoint.cc:178-189
    std::unique_lock<std::mutex> l(mMutexFeatures);
    if (mObservations.count(pKF)) return;
    mObservations[pKF] = idx;
    sqrtba::MapMirror::Global().OnAddObservation(this, pKF, idx);   // <- hook
  }
  void EraseObservation(KeyFrame* pKF) {            // MapPoint.cc:191-215
    std::unique_lock<std::mutex> l(mMutexFeatures);
    mObservations.erase(pKF);
    sqrtba::MapMirror::Global().OnEraseObservation(this, pKF);      // <- hook
  }
  void SetBadFlag() {                               // MapPoint.cc:228-246
    std::unique_lock<std::mutex> l(mMutexFeatures);
    mbBad = true;
    mObservations.clear();
    sqrtba::MapMirror::Global().OnClearObservations(this);          // <- hook
  }
  ~MapPoint() { sqrtba::MapMirror::Global().OnDeletePoint(this); }
  bool isBad() { return mbBad; }
  int GetIndexInKeyFrame(KeyFrame* pKF) {  // MapPoint.h:137, MapPoint.cc:493-500
    std::unique_lock<std::mutex> l(mMutexFeatures);
    auto it = mObservations.find(pKF);
    return it == mObservations.end() ? -1 : (int)it->second;
  }
  void UpdateNormalAndDepth() { nUpdateNormalAndDepth++; }  // MapPoint.cc:531-575 is outside the BA path
  cv::Mat mWorldPos;
  bool mbBad = false;
  std::map<KeyFrame*, size_t> mObservations;
  std::mutex mMutexPos, mMutexFeatures;
  static std::mutex mGlobalMutex;  // MapPoint.h:252, held while PoseOptimization copies map points (g2oOptimizer.cc:433)
};

// include/data_structure/Frame.h -- the members g2oOptimizer::PoseOptimization touches (g2oOptimizer.cc:405-559,655-690):
// mTcw (:407,510), N (:412), mvpMapPoints (:436), mvuRight (:442), mvbOutlier (:445,525-541,663-677), mvKeysUn (:449),
// mvInvLevelSigma2 (:457), fx fy cx cy (:465-468) [mbf for the stereo edge of upstream ORB-SLAM2], SetPose (:557)
class Frame {
 public:
  int N = 0;
  cv::Mat mTcw;
  float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvuRight;
  std::vector<float> mvInvLevelSigma2;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  // lidar features of the frame in its own frame (Frame.h:273-277), read by the lidar block of PoseOptimization
  PointIRTCloud corner_points_sharp_, surface_points_flat_, surface_points_flat_normal_;
  void SetPose(const cv::Mat& T) { T.copyTo(mTcw); }
};

class Map {
 public:
  std::vector<KeyFrame*> GetAllKeyFrames() { return mspKeyFrames; }
  std::vector<MapPoint*> GetAllMapPoints() { return mspMapPoints; }
  long unsigned int GetMaxKFid() { long unsigned int m = 0; for (KeyFrame* k : mspKeyFrames) m = std::max(m, k->mnId); return m; }
  std::mutex mMutexMapUpdate;
  std::vector<KeyFrame*> mspKeyFrames;
  std::vector<MapPoint*> mspMapPoints;
  std::vector<KeyFrame*> mvpKeyFrameOrigins;  // Map.h:141
};
// include/backend/LoopClosing.h:60-65
class LoopClosing {
 public:
  typedef std::map<KeyFrame*, g2o::Sim3, std::less<KeyFrame*>> KeyFrameAndPose;
};
}  // namespace ORB_SLAM2

// Incremental mirror of the map's observation graph.
//
// Every bundle-adjustment call of the reference starts by copying, for each map point of the window,
// `MapPoint::GetObservations()` -- a std::map<KeyFrame*, size_t> copied node by node under the point's mutex
// (MapPoint.cc:212-215; local BA walks it twice, g2oOptimizer.cc:759-780 and :870-919).  Once the solve takes a few
// milliseconds that copy is a visible share of the call.  The mirror keeps the same information in flat per-point
// vectors that are updated where the map itself is updated -- MapPoint::AddObservation (MapPoint.cc:178-189),
// EraseObservation (:191-215), SetBadFlag (:228-246) and Replace (:253-291) get one hook line each --
// so the adapter's gather reads a contiguous snapshot instead of re-walking 6 000 red-black trees.
//
// Contract:
//  * a point's list is kept sorted by KeyFrame* exactly like the std::map it mirrors, so a snapshot is bit-identical
//    to what the map copies give (tests/test_map_mirror.py) and the gathered problem does not depend on the path taken;
//  * hooks may be called from any thread while the caller holds the point's own mutex (the reference's mutation sites
//    do); the mirror has its own locks, sharded by point, writers exclusive / snapshots shared;
//  * a point the mirror has never seen makes Snapshot() return false and the adapter falls back to the map copies --
//    a map that was built before the mirror was switched on is attached once with Rebuild().
// Header-only, no dependency beyond the C++17 standard library; KeyFrame / MapPoint are opaque pointers here.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <shared_mutex>
#include <unordered_map>
#include <utility>
#include <vector>

namespace sqrtba {

class MapMirror {
 public:
  struct Obs {
    const void* kf;   // KeyFrame*
    size_t idx;       // index of the keypoint in that keyframe
  };

  static MapMirror& Global() {
    static MapMirror m;
    return m;
  }

  void Enable(bool on) { enabled_.store(on, std::memory_order_release); }
  bool enabled() const { return enabled_.load(std::memory_order_acquire); }

  void Clear() {
    for (Shard& s : shards_) {
      std::unique_lock<std::shared_mutex> l(s.mu);
      s.points.clear();
    }
  }

  // ---- hooks (no-ops while the mirror is off)
  // MapPoint::AddObservation: `if(mObservations.count(pKF)) return; mObservations[pKF]=idx;`
  void OnAddObservation(const void* mp, const void* kf, size_t idx) {
    if (!enabled()) return;
    Shard& s = shard(mp);
    std::unique_lock<std::shared_mutex> l(s.mu);
    std::vector<Obs>& v = s.points[mp];
    auto it = std::lower_bound(v.begin(), v.end(), kf, [](const Obs& o, const void* k) { return std::less<const void*>()(o.kf, k); });
    if (it != v.end() && it->kf == kf) return;  // the reference keeps the first index
    v.insert(it, Obs{kf, idx});
  }
  // MapPoint::EraseObservation: `mObservations.erase(pKF)`
  void OnEraseObservation(const void* mp, const void* kf) {
    if (!enabled()) return;
    Shard& s = shard(mp);
    std::unique_lock<std::shared_mutex> l(s.mu);
    auto pit = s.points.find(mp);
    if (pit == s.points.end()) return;
    std::vector<Obs>& v = pit->second;
    auto it = std::lower_bound(v.begin(), v.end(), kf, [](const Obs& o, const void* k) { return std::less<const void*>()(o.kf, k); });
    if (it != v.end() && it->kf == kf) v.erase(it);
  }
  // MapPoint::SetBadFlag / Replace: `mObservations.clear()`
  void OnClearObservations(const void* mp) {
    if (!enabled()) return;
    Shard& s = shard(mp);
    std::unique_lock<std::shared_mutex> l(s.mu);
    s.points[mp].clear();
  }
  // a freshly constructed point (no observation yet) -- so that Snapshot() knows it
  void OnNewPoint(const void* mp) {
    if (!enabled()) return;
    Shard& s = shard(mp);
    std::unique_lock<std::shared_mutex> l(s.mu);
    s.points[mp];
  }
  // ~MapPoint (the reference never deletes map points while the system runs; tests do)
  void OnDeletePoint(const void* mp) {
    Shard& s = shard(mp);
    std::unique_lock<std::shared_mutex> l(s.mu);
    s.points.erase(mp);
  }

  // Attach one point of an existing map: `obs` = its current observations in any order.
  void SetPoint(const void* mp, std::vector<Obs> obs) {
    std::sort(obs.begin(), obs.end(), [](const Obs& a, const Obs& b) { return std::less<const void*>()(a.kf, b.kf); });
    Shard& s = shard(mp);
    std::unique_lock<std::shared_mutex> l(s.mu);
    s.points[mp] = std::move(obs);
  }

  // Observations of `n` points, in the order given, into one flat array: ptr[i] .. ptr[i + 1] index rec.  `out_ptr` has
  // n + 1 entries.  Returns false (outputs unspecified) if the mirror is off or does not know one of the points.
  // Safe to call from several threads at once (disjoint or overlapping point sets).
  template <class PointPtr>
  bool Snapshot(const PointPtr* mps, size_t n, std::vector<size_t>& out_ptr, std::vector<Obs>& out_rec) const {
    if (!enabled()) return false;
    out_ptr.assign(1, 0);
    out_ptr.reserve(n + 1);
    out_rec.clear();
    out_rec.reserve(n * 12);
    for (size_t i = 0; i < n; i++) {
      const void* mp = (const void*)mps[i];
      const Shard& s = shard(mp);
      std::shared_lock<std::shared_mutex> l(s.mu);
      auto it = s.points.find(mp);
      if (it == s.points.end()) return false;
      out_rec.insert(out_rec.end(), it->second.begin(), it->second.end());
      out_ptr.push_back(out_rec.size());
    }
    return true;
  }

  size_t NumPoints() const {
    size_t n = 0;
    for (const Shard& s : shards_) {
      std::shared_lock<std::shared_mutex> l(s.mu);
      n += s.points.size();
    }
    return n;
  }

 private:
  static constexpr int kShards = 64;
  struct Shard {
    mutable std::shared_mutex mu;
    std::unordered_map<const void*, std::vector<Obs>> points;
  };
  static size_t hash(const void* p) {
    uint64_t x = (uint64_t)(uintptr_t)p;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33;
    return (size_t)x;
  }
  Shard& shard(const void* p) { return shards_[hash(p) % kShards]; }
  const Shard& shard(const void* p) const { return shards_[hash(p) % kShards]; }
  Shard shards_[kShards];
  std::atomic<bool> enabled_{false};
};

}  // namespace sqrtba

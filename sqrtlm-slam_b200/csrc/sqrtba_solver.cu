// Host orchestration of the sqrtba CUDA path + the extern "C" ABI declared in include/sqrtba.h.
//
// One handle = one CUDA stream + device buffers for one (possibly batched) BA problem.  The LM loop of the
// reference (Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:354-419 driving
// optimization_algorithm_levenberg.cpp:61-164) is run in lock-step over all windows of the batch: every
// "macro step" is  [linearise windows that accepted]  ->  landmark QR with each window's lambda  ->  PCG  ->
// back-substitution  ->  trial update  ->  trial cost  ->  per-window accept/reject on the device.
// The host only enqueues kernels, reads back two counters per macro step and polls the caller's stop flag
// (bool* pbStopFlag, src/backend/g2oOptimizer.cc:797-798) -- it never touches problem data after upload.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/sqrtba.h"
#include "sqrtba_kernels.cuh"
#include "sqrtba_chunkprec.cuh"
#include "sqrtba_poseopt.cuh"
#include "sqrtba_lidar.cuh"
#include "sqrtba_posegraph.cuh"
#include "sqrtba_sim3opt.cuh"
#include "sqrtba_poseopt_lidar.cuh"
#include "../host/host_pool.h"

namespace sqrtba {

#define CU_CHECK(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      char buf__[512];                                                                         \
      snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      err_ = buf__;                                                                            \
      return SQRTBA_ERR_CUDA;                                                                  \
    }                                                                                          \
  } while (0)

template <class T>
struct DBuf {  // device buffer that only grows
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <class T>
struct HBuf {  // pinned host staging buffer that only grows (fast H2D, reused across set_problem calls)
  T* p = nullptr;
  size_t cap = 0;
  bool pinned = true;
  int ensure(size_t n, bool want_pinned = true) {
    if (n <= cap) return 0;
    release();
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    if (want_pinned && cudaMallocHost((void**)&p, bytes) == cudaSuccess) {
      pinned = true;
    } else {  // plan-only mode (no device needed): ordinary memory
      if (want_pinned) cudaGetLastError();
      p = (T*)std::malloc(bytes);
      pinned = false;
      if (!p) return 1;
    }
    cap = n;
    return 0;
  }
  void release() {
    if (p) { if (pinned) cudaFreeHost(p); else std::free(p); }
    p = nullptr;
    cap = 0;
  }
};

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// NCCL is loaded at run time and only when a multi-GPU communicator is requested, so the single-GPU path has no
// dependency on it.  (In a Python process that already imported torch, dlopen resolves to torch's bundled libnccl.)
struct NcclApi {
  struct UniqueId { char internal[128]; };
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void* lib = nullptr;
  bool load() {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
    GetUniqueId = (int (*)(UniqueId*))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (int (*)(void**, int, UniqueId, int))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !AllGather) { lib = nullptr; return false; }
    return true;
  }
};
static NcclApi g_nccl;

class Solver {
 public:
  explicit Solver(const sqrtba_config& cfg) : cfg_(cfg) {}
  ~Solver() { destroy(); }

  int init() {
    int ndev = 0;
    CU_CHECK(cudaGetDeviceCount(&ndev));
    if (ndev <= 0 || cfg_.device >= ndev) {
      err_ = "no CUDA device (sqrtba has no CPU fallback)";
      return SQRTBA_ERR_CUDA;
    }
    CU_CHECK(cudaSetDevice(cfg_.device));
    CU_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    CU_CHECK(cudaMallocHost((void**)&h_counters_, 8 * sizeof(int)));
    CU_CHECK(cudaHostAlloc((void**)&h_ctl_, sizeof(HostCtl), cudaHostAllocMapped));
    std::memset((void*)h_ctl_, 0, sizeof(HostCtl));
    CU_CHECK(cudaHostGetDevicePointer((void**)&d_ctl_host_, (void*)h_ctl_, 0));
    CU_CHECK(cudaEventCreate(&ev0_));
    CU_CHECK(cudaEventCreate(&ev1_));
    for (auto& e : stage_ev_) CU_CHECK(cudaEventCreate(&e));
    cudaDeviceProp prop;
    CU_CHECK(cudaGetDeviceProperties(&prop, cfg_.device));
    n_sm_ = prop.multiProcessorCount;
    CU_CHECK(cudaFuncSetAttribute(k_matvec_pipe<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pipe_smem_bytes(2, MAXSLOT, false)));
    CU_CHECK(cudaFuncSetAttribute(k_matvec_pipe<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pipe_smem_bytes(3, MAXSLOT, false)));
    CU_CHECK(cudaFuncSetAttribute(k_matvec_pipe<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pipe_smem_bytes(2, MAXSLOT, false)));
    CU_CHECK(cudaFuncSetAttribute(k_matvec_pipe<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pipe_smem_bytes(3, MAXSLOT, false)));
    CU_CHECK(cudaFuncSetAttribute(k_pcg_persist<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)persist_smem_bytes(2, MAXSLOT, false)));
    CU_CHECK(cudaFuncSetAttribute(k_pcg_persist<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)persist_smem_bytes(3, MAXSLOT, false)));
    CU_CHECK(cudaFuncSetAttribute(k_pcg_persist<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)persist_smem_bytes(2, MAXSLOT, false)));
    CU_CHECK(cudaFuncSetAttribute(k_pcg_persist<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)persist_smem_bytes(3, MAXSLOT, false)));
    CU_CHECK(cudaFuncSetAttribute(k_pcg_persist<2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)persist_smem_bytes(2, MAXSLOT, false)));
    CU_CHECK(cudaFuncSetAttribute(k_pcg_persist<3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)persist_smem_bytes(3, MAXSLOT, false)));
    CU_CHECK(cudaFuncSetAttribute(k_pcg_persist<2, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)persist_smem_bytes(2, 48, true, true)));
    CU_CHECK(cudaFuncSetAttribute(k_pcg_persist<2, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)persist_smem_bytes(2, 48, true, true)));
    CU_CHECK(cudaFuncSetAttribute(k_chunk_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_FACTOR_SMEM));
    CU_CHECK(cudaFuncSetAttribute(k_pg_factor_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU_CHECK(cudaFuncSetAttribute(k_qr_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QR_PIPE_SMEM));
    CU_CHECK(cudaFuncSetAttribute(k_qr_pipe2<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QR_PIPE2_SMEM));
    CU_CHECK(cudaFuncSetAttribute(k_qr_pipe2<0, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QR_PIPE2_SMEM));
    int coop = 0;
    CU_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg_.device));
    coop_ok_ = coop != 0;
    return SQRTBA_OK;
  }

  void destroy() {
    if (stream_) {
      cudaSetDevice(cfg_.device);
      cudaStreamSynchronize(stream_);
    }
    comm_destroy();
    release_all();
    if (h_counters_) cudaFreeHost(h_counters_);
    h_counters_ = nullptr;
    if (step_exec_) cudaGraphExecDestroy(step_exec_);
    step_exec_ = nullptr;
    if (h_ctl_) cudaFreeHost((void*)h_ctl_);
    h_ctl_ = nullptr;
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    for (auto& e : stage_ev_)
      if (e) cudaEventDestroy(e);
    ev0_ = ev1_ = nullptr;
    for (auto& e : mv_events_) cudaEventDestroy(e);
    mv_events_.clear();
    if (stream_) cudaStreamDestroy(stream_);
    stream_ = nullptr;
  }

  const char* last_error() const { return err_.c_str(); }

  // ------------------------------------------------------------------------------------------ problem upload
  int set_problem(int n_win, const int64_t* wpose, const int64_t* wpoint, const int64_t* wobs, int n_pose, int n_point,
                  int n_obs, const double* pose_qt, const uint8_t* pose_fixed, const double* cam,
                  const double* point_xyz_in, const int32_t* obs_pose_in, const int32_t* obs_point_in,
                  const float* obs_meas_in) {
    const double* point_xyz = point_xyz_in;
    const int32_t* obs_pose = obs_pose_in;
    const int32_t* obs_point = obs_point_in;
    const float* obs_meas = obs_meas_in;
    if (n_pose <= 0 || n_point <= 0 || n_obs <= 0 || n_win <= 0 || !pose_qt || !pose_fixed || !cam || !point_xyz ||
        !obs_pose || !obs_point || !obs_meas) {
      err_ = "set_problem: empty problem or null pointer";
      return SQRTBA_ERR_INVALID;
    }
    // SQRTBA_HOST_TIMING=1: wall time of the host-side steps on stderr (set_problem is the part of the end-to-end call
    // that is neither a copy nor a kernel)
    static const bool host_timing = std::getenv("SQRTBA_HOST_TIMING") != nullptr;
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!host_timing) return;
      const auto now = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[sqrtba host] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
      tick = now;
    };
    if (!plan_only_) CU_CHECK(cudaSetDevice(cfg_.device));
    have_problem_ = false;
    lidar_edges_set_ = lidar_assoc_set_ = lidar_active_ = false;  // pose indices of the old problem
    std::vector<int> pose_win(n_pose, 0), point_win(n_point, 0);
    if (n_win > 1) {
      if (!wpose || !wpoint || !wobs || wpose[n_win] != n_pose || wpoint[n_win] != n_point || wobs[n_win] != n_obs ||
          wpose[0] != 0 || wpoint[0] != 0 || wobs[0] != 0) {
        err_ = "set_problem_batch: window offsets inconsistent with array sizes (every offset array starts at 0 and ends at the array size)";
        return SQRTBA_ERR_INVALID;
      }
      for (int w = 0; w < n_win; w++) {
        if (wpose[w] > wpose[w + 1] || wpoint[w] > wpoint[w + 1] || wobs[w] > wobs[w + 1]) {
          err_ = "set_problem_batch: window offsets must be non-decreasing";
          return SQRTBA_ERR_INVALID;
        }
        for (int64_t i = wpose[w]; i < wpose[w + 1]; i++) pose_win[i] = w;
        for (int64_t i = wpoint[w]; i < wpoint[w + 1]; i++) point_win[i] = w;
      }
    }
    // free-pose slots in pose order (g2o: hessianIndex, poses first; sparse_optimizer.cpp:166-190)
    std::vector<int> pose_slot(n_pose, -1), slot_pose, slot_win, win_slot_ptr(n_win + 1, 0);
    for (int i = 0; i < n_pose; i++)
      if (!pose_fixed[i]) {
        pose_slot[i] = (int)slot_pose.size();
        slot_pose.push_back(i);
        slot_win.push_back(pose_win[i]);
        win_slot_ptr[pose_win[i] + 1]++;
      }
    for (int w = 0; w < n_win; w++) win_slot_ptr[w + 1] += win_slot_ptr[w];
    const int n_slot = (int)slot_pose.size();
    lap("windows + slots");
    // ---- observation pre-processing (multi-threaded; all outputs land in pinned staging buffers)
    const int ld = ((n_obs + 31) / 32) * 32;
    int max_win_slots = 0;
    for (int w = 0; w < n_win; w++) max_win_slots = std::max(max_win_slots, win_slot_ptr[w + 1] - win_slot_ptr[w]);
    const int smallwin = (max_win_slots < 65535) ? 1 : 0;       // window-relative slots fit the 16-bit meta field
    // p/q of a window fit the matvec CTA's shared memory -- unless it is ONE window with enough keyframes for the chunked
    // vector phase of global BA to pay (owner CTAs, two-level preconditioner): big_window_min_slots_
    const int pq_shared = (max_win_slots <= MAXSLOT && !(n_win == 1 && max_win_slots >= big_window_min_slots())) ? 1 : 0;
    if (h_obs_slot_.ensure(n_obs, !plan_only_) || h_obs_lp_.ensure(n_obs, !plan_only_)) { err_ = "pinned host allocation failed"; return SQRTBA_ERR_ALLOC; }
    int* obs_slot = h_obs_slot_.p;
    unsigned* obs_lp = h_obs_lp_.p;
    const int n_thr = std::max(1, std::min<int>(cfg_.reserved[3] > 0 ? cfg_.reserved[3] : (int)std::thread::hardware_concurrency(), 64));
    // A. validate, free slot per observation, first observation of every landmark
    std::vector<int> lm_first(n_point, -1), lm_cnt(n_point, 0);
    {
      std::atomic<int> bad(0);
      pool_.chunks(n_obs, 1 << 16, n_thr, [&](long long k0, long long k1) {
        for (long long k = k0; k < k1; k++) {
          const int ip = obs_pose[k], il = obs_point[k];
          if (ip < 0 || ip >= n_pose || il < 0 || il >= n_point) { bad.store(1); return; }
          if (k && il < obs_point[k - 1]) { bad.store(2); return; }
          if (pose_win[ip] != point_win[il]) { bad.store(3); return; }
          if (n_win > 1 && (k < wobs[point_win[il]] || k >= wobs[point_win[il] + 1])) { bad.store(4); return; }
          obs_slot[k] = pose_slot[ip];
          // meta word: "takes no part in the in-CTA reduction" until step C says otherwise + the stereo bit
          obs_lp[k] = 0xffffu | (!(obs_meas[(size_t)k * 4 + 2] < 0.0f) ? LP_STEREO : 0u);
          if (k == 0 || il != obs_point[k - 1]) lm_first[il] = (int)k;
        }
      });
      if (bad.load()) {
        err_ = bad.load() == 1 ? "set_problem: observation index out of range"
             : bad.load() == 2 ? "set_problem: observations must be grouped by landmark (non-decreasing obs_point)"
             : bad.load() == 3 ? "set_problem_batch: observation links a pose and a point of different windows"
                               : "set_problem_batch: observation lies outside the observation range of its window";
        return SQRTBA_ERR_INVALID;
      }
      int next = n_obs;
      for (int l = n_point - 1; l >= 0; l--)
        if (lm_first[l] >= 0) { lm_cnt[l] = next - lm_first[l]; next = lm_first[l]; }
    }
    lap("A validate + slots per obs");
    // A2. big windows (global BA): order the landmarks of each window by their first free pose, so that consecutive
    //     tiles touch a narrow, slowly drifting band of pose slots (shared accumulator window of the matvec, few
    //     distinct slots per tile for the run-table reductions, cache-resident pose gathers).  The order the caller
    //     uses (std::set<MapPoint*> pointer order in the reference, g2oOptimizer.cc:117,176) is restored on read-back.
    perm_.clear(); old_first_.clear(); new_first_.clear();
    if (!pq_shared && smallwin && cfg_.reserved[4] == 0) {
      std::vector<int> key(n_point);
      pool_.chunks(n_point, 1 << 14, n_thr, [&](long long l0, long long l1) {
        for (long long l = l0; l < l1; l++) {
          int k = 0x7fffffff;
          if (lm_first[l] >= 0)
            for (int o = lm_first[l]; o < lm_first[l] + lm_cnt[l]; o++)
              if (obs_slot[o] >= 0) k = std::min(k, obs_slot[o]);
          key[l] = k;
        }
      });
      perm_.resize(n_point);
      for (int l = 0; l < n_point; l++) perm_[l] = l;
      for (int w = 0; w < n_win; w++) {
        const int p0 = (n_win > 1) ? (int)wpoint[w] : 0, p1 = (n_win > 1) ? (int)wpoint[w + 1] : n_point;
        std::stable_sort(perm_.begin() + p0, perm_.begin() + p1, [&](int a, int b) { return key[a] < key[b]; });
      }
      old_first_ = lm_first;
      new_first_.assign(n_point, -1);
      std::vector<int> cnt_new(n_point, 0);
      int pos = 0;
      for (int j = 0; j < n_point; j++) {
        const int l = perm_[j];
        cnt_new[j] = lm_cnt[l];
        if (lm_cnt[l] > 0) { new_first_[j] = pos; pos += lm_cnt[l]; }
      }
      if (h_perm_pose_.ensure(n_obs, !plan_only_) || h_perm_point_.ensure(n_obs, !plan_only_) || h_perm_meas_.ensure((size_t)n_obs * 4, !plan_only_) ||
          h_perm_xyz_.ensure((size_t)n_point * 3, !plan_only_) || h_perm_slot_.ensure(n_obs, !plan_only_)) {
        err_ = "pinned host allocation failed";
        return SQRTBA_ERR_ALLOC;
      }
      pool_.chunks(n_point, 1 << 12, n_thr, [&](long long j0, long long j1) {
        for (long long j = j0; j < j1; j++) {
          const int l = perm_[j];
          for (int c = 0; c < 3; c++) h_perm_xyz_.p[j * 3 + c] = point_xyz_in[(size_t)l * 3 + c];
          const int src = old_first_[l], dst = new_first_[j], k = cnt_new[j];
          for (int i = 0; i < k; i++) {
            h_perm_pose_.p[dst + i] = obs_pose_in[src + i];
            h_perm_point_.p[dst + i] = (int)j;
            h_perm_slot_.p[dst + i] = obs_slot[src + i];
            std::memcpy(h_perm_meas_.p + (size_t)(dst + i) * 4, obs_meas_in + (size_t)(src + i) * 4, 4 * sizeof(float));
          }
        }
      });
      std::memcpy(obs_slot, h_perm_slot_.p, (size_t)n_obs * sizeof(int));
      lm_first = new_first_;
      lm_cnt.swap(cnt_new);
      lm_count_new_ = lm_cnt;
      point_xyz = h_perm_xyz_.p; obs_pose = h_perm_pose_.p; obs_point = h_perm_point_.p; obs_meas = h_perm_meas_.p;
    }
    if (!perm_.empty())  // re-ordered observations: the stereo bits follow the new order
      pool_.chunks(n_obs, 1 << 16, n_thr, [&](long long k0, long long k1) {
        for (long long k = k0; k < k1; k++) obs_lp[k] = 0xffffu | (!(obs_meas[(size_t)k * 4 + 2] < 0.0f) ? LP_STEREO : 0u);
      });
    lap("A2 landmark order");
    // The caller's big arrays (or their re-ordered copies) are final now: start their host-to-device copies so that they
    // overlap the remaining host-side preprocessing (truly asynchronous when the caller's buffers are pinned).
    auto up = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_); };
    if (!plan_only_) {
    CU_CHECK(d_cam_.ensure((size_t)n_pose * 5));
    CU_CHECK(d_meas_.ensure((size_t)n_obs));
    CU_CHECK(d_obs_pose_.ensure((size_t)n_obs));
    CU_CHECK(d_obs_point_.ensure((size_t)n_obs));
    CU_CHECK(d_point0_.ensure((size_t)n_point * 3));
    CU_CHECK(up(d_cam_.p, cam, (size_t)n_pose * 5 * sizeof(double)));
    CU_CHECK(up(d_meas_.p, obs_meas, (size_t)n_obs * sizeof(float4)));
    CU_CHECK(up(d_obs_pose_.p, obs_pose, (size_t)n_obs * sizeof(int)));
    CU_CHECK(up(d_obs_point_.p, obs_point, (size_t)n_obs * sizeof(int)));
    CU_CHECK(up(d_point0_.p, point_xyz, (size_t)n_point * 3 * sizeof(double)));
    }
    lap("early uploads issued");
    // B. work chunks: landmark ranges inside one window of roughly equal observation count.  Items and tiles never
    //    span chunks (a chunk boundary merely ends an item early), so chunks are processed independently.
    struct Chunk { int win, l0, l1; std::vector<int> it_start, it_cnt, run_ptr, runs; std::vector<TileInfo> tiles; long long jq = 0; };
    std::vector<Chunk> chunks;
    {
      const long long target = std::max<long long>(1 << 13, (long long)n_obs / (8LL * n_thr));
      for (int w = 0; w < n_win; w++) {
        const int p0 = (n_win > 1) ? (int)wpoint[w] : 0, p1 = (n_win > 1) ? (int)wpoint[w + 1] : n_point;
        int l0 = p0;
        long long acc = 0;
        for (int l = p0; l < p1; l++) {
          acc += lm_cnt[l];
          if (acc >= target || l == p1 - 1) {
            Chunk c; c.win = w; c.l0 = l0; c.l1 = l + 1;
            chunks.push_back(std::move(c));
            l0 = l + 1; acc = 0;
          }
        }
      }
    }
    lap("B chunks");
    // C. per chunk: pack whole landmarks into items of <= 32 observations, items into tiles of <= WARPS items, and per
    //    tile the pose-sorted ranks + run table (see tile_scatter / k_matvec_pipe)
    {
      std::atomic<int> next_chunk(0);
      auto worker = [&]() {
        std::vector<int> stamp(std::max(max_win_slots, 1), -1), local_of(std::max(max_win_slots, 1), 0), distinct, lptr, cursor;
        for (;;) {
          const int ci = next_chunk.fetch_add(1);
          if (ci >= (int)chunks.size()) break;
          Chunk& c = chunks[ci];
          const int ws0 = win_slot_ptr[c.win];
          std::fill(stamp.begin(), stamp.end(), -1);
          int cur_start = 0, cur_cnt = 0;
          auto flush = [&]() {
            if (cur_cnt > 0) { c.it_start.push_back(cur_start); c.it_cnt.push_back(cur_cnt); }
            cur_cnt = 0;
          };
          for (int l = c.l0; l < c.l1; l++) {
            const int k = lm_cnt[l];
            if (k == 0) continue;
            if (k > 32) {
              flush();
              c.it_start.push_back(lm_first[l]);
              c.it_cnt.push_back(k);
            } else {
              if (cur_cnt > 0 && cur_cnt + k > 32) flush();
              if (cur_cnt == 0) cur_start = lm_first[l];
              cur_cnt += k;
            }
          }
          flush();
          const int ni = (int)c.it_start.size();
          c.run_ptr.push_back(0);
          int it = 0;
          while (it < ni) {
            TileInfo ti{};
            ti.item0 = it;  // chunk-local, rebased in the merge
            ti.win = c.win;
            ti.o0 = c.it_start[it];
            ti.is_long = c.it_cnt[it] > 32 ? 1 : 0;
            if (ti.is_long) {
              ti.nitem = 1;
            } else {
              int n = 0;
              while (it + n < ni && n < WARPS && c.it_cnt[it + n] <= 32) n++;
              ti.nitem = n;
            }
            const int last = it + ti.nitem - 1;
            ti.o1 = c.it_start[last] + c.it_cnt[last];
            for (int i = 0; i < 4; i++) ti.cnt[i] = (i < ti.nitem) ? c.it_cnt[it + i] : 0;
            const int t = (int)c.tiles.size();
            if (!ti.is_long) {
              distinct.clear();
              for (int o = ti.o0; o < ti.o1; o++) {
                const int sl = obs_slot[o] - ws0;
                if (obs_slot[o] >= 0 && stamp[sl] != t) { stamp[sl] = t; distinct.push_back(sl); }
              }
              std::sort(distinct.begin(), distinct.end());
              const int nl = (int)distinct.size();
              for (int i = 0; i < nl; i++) local_of[distinct[i]] = i;
              lptr.assign(nl + 1, 0);
              for (int o = ti.o0; o < ti.o1; o++)
                if (obs_slot[o] >= 0) lptr[local_of[obs_slot[o] - ws0] + 1]++;
              for (int i = 0; i < nl; i++) lptr[i + 1] += lptr[i];
              cursor.assign(lptr.begin(), lptr.end());
              for (int o = ti.o0; o < ti.o1; o++) {
                if (obs_slot[o] < 0) continue;
                const int rank = cursor[local_of[obs_slot[o] - ws0]]++;
                const unsigned low = smallwin ? (unsigned)(obs_slot[o] - ws0) : 0u;
                obs_lp[o] = low | ((unsigned)rank << 16) | (obs_lp[o] & LP_STEREO);
              }
              ti.nfree = lptr[nl];
              ti.nrun = nl;
              // run table: nl+1 rank offsets, then the nl slots (window-relative when every window is small)
              c.runs.insert(c.runs.end(), lptr.begin(), lptr.end());
              for (int i = 0; i < nl; i++) c.runs.push_back(smallwin ? distinct[i] : distinct[i] + ws0);
            }
            c.run_ptr.push_back((int)c.runs.size());
            for (int i = 0; i < 4; i++) {
              ti.fcnt[i] = 0;
              if (i < ti.nitem && !ti.is_long)
                for (int o = c.it_start[it + i]; o < c.it_start[it + i] + c.it_cnt[it + i]; o++) ti.fcnt[i] += obs_slot[o] >= 0;
            }
            // columns of the JQ block: free-pose observations only (a long tile keeps one per observation); even
            ti.nt = ((ti.is_long ? (ti.o1 - ti.o0) : ti.nfree) + 1) & ~1;
            const int run_ints = ti.is_long ? 0 : 2 * ti.nrun + 1;
            ti.blk_doubles = JQ_HDR + JQ_ROWS * ti.nt + 2 * ((run_ints + 3) / 4);
            ti.jq_off = c.jq + JQ_HDR;  // chunk-local, rebased in the merge
            c.jq += ti.blk_doubles;
            c.tiles.push_back(ti);
            it += ti.nitem;
          }
        }
      };
      const int n_work = std::min<int>(n_thr, (int)chunks.size());  // never more workers than chunks
      pool_.run(n_work, n_work, [&](int) { worker(); });
    }
    lap("C items/tiles/run tables");
    // D. merge the chunks (offsets are prefix sums over chunks in order)
    std::vector<long long> item_off(chunks.size() + 1, 0), tile_off(chunks.size() + 1, 0), run_off(chunks.size() + 1, 0),
        jq_off(chunks.size() + 1, 0);
    std::vector<int> win_item_ptr(n_win + 1, 0);
    for (size_t c = 0; c < chunks.size(); c++) {
      item_off[c + 1] = item_off[c] + (long long)chunks[c].it_start.size();
      tile_off[c + 1] = tile_off[c] + (long long)chunks[c].tiles.size();
      run_off[c + 1] = run_off[c] + (long long)chunks[c].runs.size();
      jq_off[c + 1] = jq_off[c] + chunks[c].jq;
      win_item_ptr[chunks[c].win + 1] += (int)chunks[c].it_start.size();
    }
    for (int w = 0; w < n_win; w++) win_item_ptr[w + 1] += win_item_ptr[w];
    const int n_item = (int)item_off.back(), n_tile = (int)tile_off.back();
    const long long jq_total = jq_off.back();
    const size_t n_runs = (size_t)run_off.back();
    if (h_item_start_.ensure(n_item, !plan_only_) || h_item_cnt_.ensure(n_item, !plan_only_) || h_item_win_.ensure(n_item, !plan_only_) || h_tiles_pin_.ensure(n_tile, !plan_only_) ||
        h_tile_run_ptr_.ensure(n_tile + 1, !plan_only_) || h_tile_runs_.ensure(n_runs, !plan_only_)) {
      err_ = "pinned host allocation failed";
      cudaStreamSynchronize(stream_);  // the early host-to-device copies still read the caller's buffers
      return SQRTBA_ERR_ALLOC;
    }
    {
      std::atomic<int> next_chunk(0);
      auto worker = [&]() {
        for (;;) {
          const int ci = next_chunk.fetch_add(1);
          if (ci >= (int)chunks.size()) break;
          const Chunk& c = chunks[ci];
          const long long io = item_off[ci], to = tile_off[ci], ro = run_off[ci];
          for (size_t i = 0; i < c.it_start.size(); i++) {
            h_item_start_.p[io + i] = c.it_start[i];
            h_item_cnt_.p[io + i] = c.it_cnt[i];
            h_item_win_.p[io + i] = c.win;
          }
          for (size_t t = 0; t < c.tiles.size(); t++) {
            TileInfo ti = c.tiles[t];
            ti.item0 += (int)io;
            ti.jq_off += jq_off[ci];
            h_tiles_pin_.p[to + t] = ti;
            h_tile_run_ptr_.p[to + t] = (int)ro + c.run_ptr[t];
          }
          if (!c.runs.empty()) std::memcpy(h_tile_runs_.p + ro, c.runs.data(), c.runs.size() * sizeof(int));
        }
      };
      const int n_work = std::min<int>(n_thr, (int)chunks.size());
      pool_.run(n_work, n_work, [&](int) { worker(); });
      h_tile_run_ptr_.p[n_tile] = (int)n_runs;
    }
    lap("D merge");
    plan_n_item_ = n_item; plan_n_tile_ = n_tile; plan_n_runs_ = (long long)n_runs; plan_jq_total_ = jq_total;
    plan_smallwin_ = smallwin; plan_pq_shared_ = pq_shared;
    if (plan_only_) return SQRTBA_OK;  // sqrtba_debug_plan: the host-side tiling only, no device needed
    // ---- device allocation
    P_ = Dev{};
    P_.n_pose = n_pose; P_.n_point = n_point; P_.n_obs = n_obs; P_.n_win = n_win; P_.n_slot = n_slot; P_.n_item = n_item;
    P_.n_tile = n_tile; P_.ld = ld; P_.smallwin = smallwin; P_.pq_shared = pq_shared;
    // The fused linearise + QR kernel (k_linqr_pipe) is an OPT-IN A/B variant (reserved[7] = 7): it removes the J_l / r
    // round trip through HBM (264 instead of ~600 B per observation) but measured SLOWER than the two separate kernels
    // -- 3.87 ms against 1.30 + 1.73 ms on 256 C0 windows -- because neither stage is bandwidth-bound any more: both
    // are bound by dependency chains at 16 warps/SM, and the fused kernel needs 168 registers (12 warps/SM).
    P_.fused = (cfg_.reserved[7] == 7) ? 1 : 0;
    const size_t No = n_obs, Nl = n_point, Ns = std::max(n_slot, 1), Ld = ld;
    CU_CHECK(d_cam_.ensure((size_t)n_pose * 5));
    CU_CHECK(d_pose_slot_.ensure(n_pose));
    CU_CHECK(d_slot_pose_.ensure(Ns));
    CU_CHECK(d_slot_win_.ensure(Ns));
    CU_CHECK(d_slot_cam_.ensure(Ns * 3));
    CU_CHECK(d_pose_win_.ensure(n_pose));
    CU_CHECK(d_point_win_.ensure(n_point));
    CU_CHECK(d_meas_.ensure(No));
    CU_CHECK(d_obs_pose_.ensure(No));
    CU_CHECK(d_obs_point_.ensure(No));
    CU_CHECK(d_obs_slot_.ensure(No));
    CU_CHECK(d_item_start_.ensure(n_item));
    CU_CHECK(d_item_cnt_.ensure(n_item));
    CU_CHECK(d_item_win_.ensure(n_item));
    CU_CHECK(d_win_item_ptr_.ensure(n_win + 1));
    CU_CHECK(d_win_slot_ptr_.ensure(n_win + 1));
    CU_CHECK(d_tiles_.ensure(n_tile));
    CU_CHECK(d_obs_lp_.ensure(No));
    CU_CHECK(d_tile_run_ptr_.ensure(n_tile + 1));
    CU_CHECK(d_tile_runs_.ensure(n_runs));
    CU_CHECK(d_pose_.ensure((size_t)n_pose * 7));
    CU_CHECK(d_pose0_.ensure((size_t)n_pose * 7));
    CU_CHECK(d_pose_bak_.ensure((size_t)n_pose * 7));
    CU_CHECK(d_point_.ensure(Nl * 3));
    CU_CHECK(d_point0_.ensure(Nl * 3));
    CU_CHECK(d_point_bak_.ensure(Nl * 3));
    CU_CHECK(d_level_.ensure(No));
    CU_CHECK(d_outlier_.ensure(No));
    CU_CHECK(d_err_.ensure(Ld * 3));
    CU_CHECK(d_JQ_.ensure((size_t)jq_total));
    CU_CHECK(d_Jl_.ensure(Ld * 9));
    CU_CHECK(d_r_.ensure(Ld * 3));
    CU_CHECK(d_R_.ensure(Nl * 6));
    CU_CHECK(d_tl_.ensure(Nl * 3));
    CU_CHECK(d_bl_.ensure(Nl * 3));
    CU_CHECK(d_dl_.ensure(Nl * 3));
    CU_CHECK(d_slotvec_.ensure(Ns * (6 * 8 + 21 + 36)));
    CU_CHECK(d_chi_part_.ensure(n_item));
    CU_CHECK(d_scale_part_.ensure(n_item));
    CU_CHECK(d_ctl_.ensure(n_win));
    CU_CHECK(d_wred_.ensure((size_t)3 * n_win));
    max_trace_ = 200;
    CU_CHECK(d_trace_.ensure((size_t)n_win * max_trace_ * TRACE_COLS));
    CU_CHECK(d_counters_.ensure(8));

    CU_CHECK(up(d_pose_slot_.p, pose_slot.data(), n_pose * sizeof(int)));
    if (n_slot) {
      CU_CHECK(up(d_slot_pose_.p, slot_pose.data(), n_slot * sizeof(int)));
      CU_CHECK(up(d_slot_win_.p, slot_win.data(), n_slot * sizeof(int)));
    }
    h_slot_cam_.assign((size_t)Ns * 3, 0.0);
    for (int sl = 0; sl < n_slot; sl++) {  // fx, fy, bf of the keyframe in every free slot (jp_compact's intrinsics)
      const double* c5 = cam + (size_t)slot_pose[sl] * 5;
      h_slot_cam_[(size_t)sl * 3] = c5[0]; h_slot_cam_[(size_t)sl * 3 + 1] = c5[1]; h_slot_cam_[(size_t)sl * 3 + 2] = c5[4];
    }
    CU_CHECK(up(d_slot_cam_.p, h_slot_cam_.data(), h_slot_cam_.size() * sizeof(double)));
    CU_CHECK(up(d_pose_win_.p, pose_win.data(), n_pose * sizeof(int)));
    CU_CHECK(up(d_point_win_.p, point_win.data(), n_point * sizeof(int)));
    CU_CHECK(up(d_obs_slot_.p, h_obs_slot_.p, No * sizeof(int)));
    CU_CHECK(up(d_item_start_.p, h_item_start_.p, n_item * sizeof(int)));
    CU_CHECK(up(d_item_cnt_.p, h_item_cnt_.p, n_item * sizeof(int)));
    CU_CHECK(up(d_item_win_.p, h_item_win_.p, n_item * sizeof(int)));
    CU_CHECK(up(d_win_item_ptr_.p, win_item_ptr.data(), (n_win + 1) * sizeof(int)));
    CU_CHECK(up(d_win_slot_ptr_.p, win_slot_ptr.data(), (n_win + 1) * sizeof(int)));
    CU_CHECK(up(d_tiles_.p, h_tiles_pin_.p, (size_t)n_tile * sizeof(TileInfo)));
    CU_CHECK(up(d_tile_run_ptr_.p, h_tile_run_ptr_.p, (size_t)(n_tile + 1) * sizeof(int)));
    if (n_runs) CU_CHECK(up(d_tile_runs_.p, h_tile_runs_.p, n_runs * sizeof(int)));
    CU_CHECK(up(d_obs_lp_.p, h_obs_lp_.p, No * sizeof(unsigned)));
    // poses: normalise the quaternion the way SE3Quat's constructor does (se3quat.h:58-64)
    std::vector<double> pq(pose_qt, pose_qt + (size_t)n_pose * 7);
    for (int i = 0; i < n_pose; i++) quat_normalize_pos_w(&pq[(size_t)i * 7 + 3]);
    CU_CHECK(up(d_pose0_.p, pq.data(), (size_t)n_pose * 7 * sizeof(double)));
    CU_CHECK(cudaStreamSynchronize(stream_));  // host staging vectors go out of scope

    P_.cam = d_cam_.p; P_.pose_slot = d_pose_slot_.p; P_.slot_pose = d_slot_pose_.p; P_.slot_win = d_slot_win_.p;
    P_.pose_win = d_pose_win_.p; P_.point_win = d_point_win_.p; P_.obs_meas = d_meas_.p; P_.obs_pose = d_obs_pose_.p;
    P_.obs_point = d_obs_point_.p; P_.obs_slot = d_obs_slot_.p; P_.item_start = d_item_start_.p;
    P_.item_cnt = d_item_cnt_.p; P_.item_win = d_item_win_.p; P_.win_item_ptr = d_win_item_ptr_.p;
    P_.win_slot_ptr = d_win_slot_ptr_.p;
    P_.slot_cam = d_slot_cam_.p;
    P_.tiles = d_tiles_.p;
    P_.obs_lp = d_obs_lp_.p;
    P_.tile_run_ptr = d_tile_run_ptr_.p; P_.tile_runs = d_tile_runs_.p;
    P_.pose = d_pose_.p; P_.point = d_point_.p; P_.pose_bak = d_pose_bak_.p; P_.point_bak = d_point_bak_.p;
    P_.obs_level = d_level_.p; P_.obs_outlier = d_outlier_.p;
    P_.err = d_err_.p; P_.JQ = d_JQ_.p; P_.Jl = d_Jl_.p; P_.r = d_r_.p;
    P_.R = d_R_.p; P_.tl = d_tl_.p; P_.bl = d_bl_.p; P_.dl = d_dl_.p;
    double* sv = d_slotvec_.p;  // [bp hd | bs D | x res z p q | Dinv]: the two groups that cross ranks are contiguous
    P_.bp = sv; sv += Ns * 6;
    P_.hd = sv; sv += Ns * 6;
    P_.bs = sv; sv += Ns * 6;
    P_.D = sv; sv += Ns * 21;
    P_.x = sv; sv += Ns * 6;
    P_.res = sv; sv += Ns * 6;
    P_.z = sv; sv += Ns * 6;
    P_.p = sv; sv += Ns * 6;
    P_.q = sv; sv += Ns * 6;
    P_.Dinv = sv;
    P_.chi_part = d_chi_part_.p; P_.scale_part = d_scale_part_.p; P_.wred = d_wred_.p;
    P_.ctl = d_ctl_.p; P_.trace = d_trace_.p; P_.max_trace = max_trace_; P_.counters = d_counters_.p;
    P_.prof = nullptr;
#ifdef SQRTBA_PIPE_PROF
    CU_CHECK(d_prof_.ensure((size_t)4096 * WARPS * 8));
    CU_CHECK(cudaMemsetAsync(d_prof_.p, 0, d_prof_.cap * sizeof(long long), stream_));
    P_.prof = d_prof_.p;
#endif
    h_tiles_.assign(h_tiles_pin_.p, h_tiles_pin_.p + n_tile);
    max_win_slots_ = std::max(max_win_slots, 1);
    // pad columns of the JQ blocks are streamed by the TMA copies: keep them defined; then headers + per-column meta
    CU_CHECK(cudaMemsetAsync(d_JQ_.p, 0, (size_t)jq_total * sizeof(double), stream_));
    k_init_jq<<<n_tile, 128, 0, stream_>>>(P_);
    CU_CHECK(cudaGetLastError());
    if (smallwin) {  // pipeline depth / residency for this problem's window size
      // big windows: run sums go straight to global atomics (measured faster than a shared accumulator window:
      // 5.70 vs 4.70 TB/s on C3); reserved[5] > 0 brings the window back for A/B profiling
      pipe_slots_ = pq_shared ? max_win_slots_ : (cfg_.reserved[5] > 0 ? std::min(cfg_.reserved[5], MAXSLOT) : 0);
      int best_ctas = 0;
      for (int S : {2, 3}) {
        int per_sm = 0;
        const size_t bytes = pipe_smem_bytes(S, pipe_slots_, !pq_shared);
        cudaError_t e;
        if (pq_shared)
          e = (S == 2) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_matvec_pipe<2, false>, PIPE_THREADS, bytes)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_matvec_pipe<3, false>, PIPE_THREADS, bytes);
        else
          e = (S == 2) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_matvec_pipe<2, true>, PIPE_THREADS, bytes)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_matvec_pipe<3, true>, PIPE_THREADS, bytes);
        if (e != cudaSuccess) per_sm = 0;
        if (cfg_.reserved[2] > 0 && S != cfg_.reserved[2]) continue;   // forced depth (profiling)
        // With the 104-byte columns a tile is 14 KB and the kernel is bound by the per-tile dependency chain of a CTA
        // (meta -> operands -> segmented sums -> run sums), not by bytes in flight: resident CTAs first, depth second
        if (per_sm > best_ctas || (per_sm == best_ctas && S > pipe_stages_) || best_ctas == 0) { best_ctas = per_sm; pipe_stages_ = S; }
      }
      pipe_ctas_ = std::max(1, best_ctas) * n_sm_;
      // the persistent PCG kernel shares the ring configuration; its grid must be co-resident (cooperative launch)
      int best = 0;
      persist_stages_ = 2;
      persist_slots_ = pq_shared ? max_win_slots_ : 48;  // big windows: slots in the shared window of the search direction
      for (int S : {2, 3}) {
        int per_sm = 0;
        const size_t pbytes = persist_smem_bytes(S, persist_slots_, !pq_shared);
        cudaError_t e;
        if (pq_shared)
          e = (S == 2) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_persist<2, false>, PIPE_THREADS, pbytes)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_persist<3, false>, PIPE_THREADS, pbytes);
        else
          e = (S == 2) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_persist<2, true>, PIPE_THREADS, pbytes)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_persist<3, true>, PIPE_THREADS, pbytes);
        if (e != cudaSuccess) per_sm = 0;
        if (cfg_.reserved[2] > 0 && S != cfg_.reserved[2]) continue;
        if (per_sm * S > best * persist_stages_ || best == 0) { best = per_sm; persist_stages_ = S; }
      }
      // (4 CTAs/SM at 96 registers was measured SLOWER for the persistent kernel -- C0 5.4 vs 4.7 ms: every extra CTA adds
      // arrivals to the grid barrier, flush atomics on the q copies and a redundant vector update)
      // Big single window (global BA): chunk preconditioner (sqrtba_chunkprec.cuh).  The owner CTA of a chunk keeps its
      // 28.6 KB inverse block in shared memory, which leaves room for a 2-stage ring at the same residency; taken only
      // if the residency holds and every chunk gets a CTA of its own.  pcg_mode = 5 keeps the 6x6 blocks (A/B).
      chunk_active_ = false;
      if (!pq_shared && n_win == 1 && n_slot > 0 && cfg_.pcg_mode != 5 && cfg_.pcg_mode != 1 && cfg_.pcg_mode != 4 &&
          (cfg_.reserved[2] == 0 || cfg_.reserved[2] == 2)) {
        const int nchunk = (n_slot + VSLOT - 1) / VSLOT;
        for (int slots : {48, 40, 32}) {
          int per_sm = 0;
          const size_t pbytes = persist_smem_bytes(2, slots, true, true);
          if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_persist<2, true, false, true>, PIPE_THREADS, pbytes) != cudaSuccess) {
            cudaGetLastError();
            per_sm = 0;
          }
          if (std::getenv("SQRTBA_HOST_TIMING"))
            std::fprintf(stderr, "[sqrtba host] chunk preconditioner: %zu B of shared memory at %d window slots -> %d CTAs/SM (6x6 path: %d)\n",
                         pbytes, slots, per_sm, best);
          if (per_sm >= best && best > 0 && std::min(n_tile, per_sm * n_sm_) >= nchunk) {
            chunk_active_ = true;
            persist_stages_ = 2;
            persist_slots_ = slots;
            break;
          }
        }
      }
      if (const char* e = std::getenv("SQRTBA_PERSIST_CTAS_PER_SM")) best = std::max(1, std::min(best, std::atoi(e)));  // A/B
      persist_ctas_ = best * n_sm_;
    }
    persist_grid_ = std::min(n_tile, persist_ctas_);
    if (smallwin && n_win == 1 && persist_grid_ > 0) {
      // Static tile ranges of the persistent PCG kernel, balanced by a cost model in SM cycles fitted to the per-CTA cycle
      // counters of the instrumented build (tools/runs/r2_v.sh; 8-way shard of C3, C2): a tile is a latency chain, not a
      // bandwidth item -- 2700 cycles whatever it holds, + ~80 per item (warp) and ~3 per observation; an observation
      // whose pose lies outside the CTA's shared window of the search direction (loop-closure tracks, very wide
      // covisibility) gathers it from global memory; a long landmark is swept by one warp.
      std::vector<double> cum(n_tile + 1, 0.0);
      for (int t = 0; t < n_tile; t++) {
        const TileInfo& ti = h_tiles_pin_.p[t];
        double cost = 2700.0 + 80.0 * ti.nitem + 3.0 * (ti.o1 - ti.o0);
        if (!pq_shared && !ti.is_long && ti.nrun > 0) {
          const int* runs = h_tile_runs_.p + h_tile_run_ptr_.p[t];
          const int lo = runs[ti.nrun + 1];
          for (int i = 0; i < ti.nrun; i++)
            if (runs[ti.nrun + 1 + i] >= lo + persist_slots_ - 6) cost += 40.0 * (runs[i + 1] - runs[i]);
        }
        if (ti.is_long) cost += 60.0 * (ti.o1 - ti.o0);
        cum[t + 1] = cum[t] + cost;
      }
      std::vector<int> ptr(persist_grid_ + 1, 0);
      int t = 0;
      for (int b = 1; b < persist_grid_; b++) {
        const double target = cum[n_tile] * b / persist_grid_;
        while (t < n_tile && cum[t + 1] <= target) t++;
        // every CTA keeps at least one tile
        t = std::max(t, ptr[b - 1] + 1);
        t = std::min(t, n_tile - (persist_grid_ - b));
        ptr[b] = t;
      }
      ptr[persist_grid_] = n_tile;
      CU_CHECK(d_ptile_.ensure(persist_grid_ + 1));
      CU_CHECK(cudaMemcpyAsync(d_ptile_.p, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
      CU_CHECK(cudaStreamSynchronize(stream_));
    }
    {
      const int nchunk = (std::max(n_slot, 1) + VSLOT - 1) / VSLOT;
      CU_CHECK(d_part_.ensure((size_t)4 * nchunk));
      CU_CHECK(d_gbar_.ensure(2));
      CU_CHECK(d_q3_.ensure((size_t)3 * KQ * 6 * std::max(n_slot, 1)));
      CU_CHECK(d_dq_.ensure((size_t)12 * std::max(n_slot, 1)));  // Dq and qf
      if (chunk_active_ && std::min(n_tile, persist_ctas_) < nchunk) chunk_active_ = false;
      // second level (coarse correction over the chunks): pcg_mode = 6 keeps the chunk level only (A/B)
      coarse_active_ = chunk_active_ && nchunk >= 2 && nchunk <= CO_MAXCH && cfg_.pcg_mode != 6 && nchunk <= n_sm_;
      if (chunk_active_) {
        CU_CHECK(d_chunk_M_.ensure((size_t)nchunk * CH_MSIZE + (coarse_active_ ? (size_t)nchunk * nchunk * CH_MBLK : 0)));
        if (coarse_active_) {
          CU_CHECK(d_co_dblk_.ensure((size_t)nchunk * 36));
          CU_CHECK(d_co_R_.ensure((size_t)nchunk * (6 * CO_LD + 36)));
          CU_CHECK(d_co_aci_.ensure((size_t)36 * nchunk * nchunk));
          CU_CHECK(d_co_w_.ensure((size_t)6 * nchunk));
          CU_CHECK(d_co_ctl_.ensure(2 + CO_MAXCH));
        }
        CU_CHECK(d_chunk_pack_.ensure((size_t)nchunk * CH_PACK));
        CU_CHECK(d_chunk_diag_.ensure((size_t)nchunk * CHB));
        CU_CHECK(d_chunk_rz_.ensure((size_t)nchunk));
      }
      if (comm_) {
        if (int rc = peer_setup((size_t)std::max(n_slot, 1) * 6)) return rc;
      }
    }
    {  // tile range of every window (tiles are ordered by window) + the reproducible mode's partial vectors
      std::vector<int> wtp(n_win + 1, 0);
      bool any_long = false;
      for (int t = 0; t < n_tile; t++) { wtp[h_tiles_pin_.p[t].win + 1]++; any_long |= h_tiles_pin_.p[t].is_long != 0; }
      for (int w = 0; w < n_win; w++) wtp[w + 1] += wtp[w];
      CU_CHECK(d_win_tile_ptr_.ensure(n_win + 1));
      CU_CHECK(cudaMemcpyAsync(d_win_tile_ptr_.p, wtp.data(), wtp.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
      CU_CHECK(cudaStreamSynchronize(stream_));
      P_.win_tile_ptr = d_win_tile_ptr_.p;
      P_.maxslot = max_win_slots_;
      // pcg_mode = 4: fixed-order reductions everywhere.  Needs every window's free poses in the CTAs' shared accumulators
      // (<= 128), no landmark with more than 32 observations (those go through direct atomics), one GPU
      det_active_ = cfg_.pcg_mode == 4 && pq_shared && smallwin && !any_long && !comm_ && n_slot > 0 && cfg_.reserved[7] == 0 &&
                    cfg_.reserved[1] == 0;
      P_.det = det_active_ ? 1 : 0;
      if (det_active_) {
        const size_t ms = (size_t)max_win_slots_;
        const size_t v_lin = (size_t)cdiv(n_tile, LIN_TPB) + n_win, v_qr = (size_t)cdiv(n_tile, QR_TPB) + n_win;
        const size_t v_q = (size_t)std::min(n_tile, pipe_ctas_) + n_win;
        CU_CHECK(d_part_lin_.ensure(v_lin * 12 * ms));
        CU_CHECK(d_part_qr_.ensure(v_qr * 27 * ms));
        CU_CHECK(d_part_q_.ensure(v_q * 6 * ms));
        P_.part_lin = d_part_lin_.p; P_.part_qr = d_part_qr_.p; P_.part_q = d_part_q_.p;
        CU_CHECK(cudaFuncSetAttribute(k_qr_pipe2<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(QR_PIPE2_SMEM + 27 * MAXSLOT * sizeof(double))));
      }
    }
    have_problem_ = true;
    step_graph_valid_ = false;  // the captured macro step carries this problem's sizes and pointers
    return reset_state();
  }

  int reset_state() {
    if (!have_problem_) { err_ = "no problem set"; return SQRTBA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(cfg_.device));
    CU_CHECK(cudaMemcpyAsync(d_pose_.p, d_pose0_.p, (size_t)P_.n_pose * 7 * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
    CU_CHECK(cudaMemcpyAsync(d_point_.p, d_point0_.p, (size_t)P_.n_point * 3 * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
    CU_CHECK(cudaMemsetAsync(d_level_.p, 0, P_.n_obs, stream_));
    CU_CHECK(cudaMemsetAsync(d_outlier_.p, 0, P_.n_obs, stream_));
    CU_CHECK(cudaMemsetAsync(d_err_.p, 0, (size_t)P_.ld * 3 * sizeof(double), stream_));
    CU_CHECK(cudaMemsetAsync(d_ctl_.p, 0, (size_t)P_.n_win * sizeof(WinCtl), stream_));
    CU_CHECK(cudaMemsetAsync(d_slotvec_.p, 0, d_slotvec_.cap * sizeof(double), stream_));
    CU_CHECK(cudaMemsetAsync(d_dl_.p, 0, (size_t)P_.n_point * 3 * sizeof(double), stream_));
    CU_CHECK(cudaStreamSynchronize(stream_));
    return SQRTBA_OK;
  }

  // ------------------------------------------------------------------------------------------ solve
  int solve_local(const volatile bool* stop, sqrtba_stats* st) {
    if (!have_problem_) { err_ = "no problem set"; return SQRTBA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(cfg_.device));
    begin_stats();
    // `const float thHuberMono = sqrt(5.991)` / thHuberStereo = sqrt(7.815) then setDelta(double) (g2oOptimizer.cc:851-853)
    const double d2 = (double)(float)std::sqrt(5.991), d3 = (double)(float)std::sqrt(7.815);
    CU_CHECK(cudaMemsetAsync(d_level_.p, 0, P_.n_obs, stream_));
    CU_CHECK(clear_traces());
    int term = 0, rc = 0;
    if ((rc = poll_stop(stop, &term))) return rc;
    if (term) return finish_stats(st);  // g2oOptimizer.cc:923-928: early out, estimates untouched
    rc = run_pass(5, 0, 1, d2, d3, stop);
    if (rc) return rc;
    if ((rc = poll_stop(stop, &term))) return rc;
    const bool do_more = !term;  // :936-947
    if (do_more) {
      launch_classify(0, 5.991, 7.815);
      rc = run_pass(10, 1, 0, d2, d3, stop);
      if (rc) return rc;
    }
    if (cfg_.third_pass_iters > 0) {
      if (lidar_assoc_set_)
        if ((rc = lidar_associate())) return rc;
      // the Huber kernels are only dropped inside `if (bDoMore)` (:947-970): a skipped second pass leaves them on
      rc = run_pass(cfg_.third_pass_iters, 2, do_more ? 0 : 1, d2, d3, stop);
      lidar_active_ = false;
      if (rc) return rc;
    }
    launch_classify(1, 5.991, 7.815);
    dump_persist_prof();
    return finish_stats(st);
  }

  int solve_global(int iters, int robust, const volatile bool* stop, sqrtba_stats* st) {
    if (!have_problem_) { err_ = "no problem set"; return SQRTBA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(cfg_.device));
    begin_stats();
    const double d2 = (double)(float)std::sqrt(5.99), d3 = (double)(float)std::sqrt(7.815);  // g2oOptimizer.cc:163-164
    CU_CHECK(cudaMemsetAsync(d_level_.p, 0, P_.n_obs, stream_));
    CU_CHECK(cudaMemsetAsync(d_outlier_.p, 0, P_.n_obs, stream_));
    CU_CHECK(clear_traces());
    int rc = run_pass(iters, 0, robust ? 1 : 0, d2, d3, stop);
    if (rc) return rc;
    dump_persist_prof();
    return finish_stats(st);
  }

  // ------------------------------------------------------------------------------------------ read-back
  int get_poses(double* out) { return download(out, d_pose_.p, (size_t)P_.n_pose * 7 * sizeof(double)); }
  int get_points(double* out) {
    if (perm_.empty()) return download(out, d_point_.p, (size_t)P_.n_point * 3 * sizeof(double));
    std::vector<double> tmp((size_t)P_.n_point * 3);
    if (int rc = download(tmp.data(), d_point_.p, tmp.size() * sizeof(double))) return rc;
    unperm_points(tmp.data(), out, 3);
    return SQRTBA_OK;
  }
  int get_outliers(uint8_t* out) {
    if (perm_.empty()) return download(out, d_outlier_.p, (size_t)P_.n_obs);
    std::vector<uint8_t> tmp((size_t)P_.n_obs);
    if (int rc = download(tmp.data(), d_outlier_.p, tmp.size())) return rc;
    unperm_obs(tmp.data(), out, 1);
    return SQRTBA_OK;
  }
  // big windows are solved in an internal landmark order (set_problem step A2): back to the caller's order
  template <class T>
  void unperm_points(const T* in, T* out, int width) const {
    for (size_t j = 0; j < perm_.size(); j++)
      for (int c = 0; c < width; c++) out[(size_t)perm_[j] * width + c] = in[j * width + c];
  }
  template <class T>
  void unperm_obs(const T* in, T* out, int width) const {
    const int n_point = (int)perm_.size();
    for (int j = 0; j < n_point; j++) {
      if (new_first_[j] < 0) continue;
      const int l = perm_[j];
      const int next_new = (j + 1 < n_point && new_first_[j + 1] >= 0) ? new_first_[j + 1] : -1;
      (void)next_new;
      const size_t src = (size_t)new_first_[j], dst = (size_t)old_first_[l];
      const size_t k = lm_count_new_[j];
      std::memcpy(out + dst * width, in + src * width, k * width * sizeof(T));
    }
  }
  int trace_len(int win) {
    if (!have_problem_ || win < 0 || win >= P_.n_win) return SQRTBA_ERR_INVALID;
    WinCtl c;
    if (download(&c, d_ctl_.p + win, sizeof(WinCtl))) return SQRTBA_ERR_CUDA;
    return std::min(c.trace_len, max_trace_);
  }
  int get_trace(int win, sqrtba_trace_row* rows, int max_rows) {
    const int n = trace_len(win);
    if (n < 0) return n;
    const int m = std::min(n, max_rows);
    static_assert(sizeof(sqrtba_trace_row) == TRACE_COLS * sizeof(double), "trace row layout");
    if (m > 0 && download(rows, d_trace_.p + (size_t)win * max_trace_ * TRACE_COLS, (size_t)m * sizeof(sqrtba_trace_row)))
      return SQRTBA_ERR_CUDA;
    return m;
  }
  int num_free() const { return have_problem_ ? P_.n_slot : SQRTBA_ERR_INVALID; }

  // host-side plan of the last set_problem (plan-only mode): tile table, meta words, run tables, landmark order
  void plan_export(int32_t* summary, int32_t* tiles20, int32_t max_tiles, uint32_t* obs_lp, int32_t* run_ptr, int32_t* runs,
                   int64_t max_runs, int32_t* perm, int32_t n_obs, int32_t n_point) const {
    if (summary) {
      summary[0] = plan_n_item_; summary[1] = plan_n_tile_; summary[2] = (int32_t)plan_n_runs_;
      summary[3] = plan_smallwin_; summary[4] = plan_pq_shared_; summary[5] = perm_.empty() ? 0 : 1;
      summary[6] = (int32_t)(plan_jq_total_ & 0x7fffffff); summary[7] = (int32_t)(plan_jq_total_ >> 31);
    }
    for (int t = 0; tiles20 && t < std::min(plan_n_tile_, max_tiles); t++) {
      const TileInfo& ti = h_tiles_pin_.p[t];
      int32_t* o = tiles20 + (size_t)t * 20;
      o[0] = ti.item0; o[1] = ti.nitem; o[2] = ti.o0; o[3] = ti.o1; o[4] = ti.win; o[5] = ti.nfree; o[6] = ti.nt;
      o[7] = ti.is_long; o[8] = ti.nrun;
      for (int i = 0; i < 4; i++) { o[9 + i] = ti.cnt[i]; o[13 + i] = ti.fcnt[i]; }
      o[17] = ti.blk_doubles; o[18] = (int32_t)(ti.jq_off & 0x7fffffff); o[19] = (int32_t)(ti.jq_off >> 31);
    }
    if (obs_lp) std::memcpy(obs_lp, h_obs_lp_.p, (size_t)n_obs * sizeof(uint32_t));
    if (run_ptr) std::memcpy(run_ptr, h_tile_run_ptr_.p, (size_t)(std::min(plan_n_tile_, max_tiles) + 1) * sizeof(int32_t));
    if (runs) std::memcpy(runs, h_tile_runs_.p, (size_t)std::min<long long>(plan_n_runs_, max_runs) * sizeof(int32_t));
    if (perm) for (int l = 0; l < n_point; l++) perm[l] = perm_.empty() ? l : perm_[l];
  }

  // ------------------------------------------------------------------------------------------ pose-only optimisation
  // g2oOptimizer::PoseOptimization for a batch of frames (one CTA per frame, one launch); independent of set_problem.
  // ------------------------------------------------------------------------------------------ lidar pass (row N4)
  int lidar_check(int cur_pose, const char* who) {
    if (!have_problem_) { err_ = std::string(who) + ": no problem set"; return SQRTBA_ERR_INVALID; }
    if (P_.n_win != 1 || comm_) { err_ = std::string(who) + ": the lidar pass belongs to one local-BA window on one GPU"; return SQRTBA_ERR_INVALID; }
    if (cur_pose < 0 || cur_pose >= P_.n_pose) { err_ = std::string(who) + ": pose index out of range"; return SQRTBA_ERR_INVALID; }
    return SQRTBA_OK;
  }
  int lidar_alloc_edges(int n_edge) {
    const size_t n = (size_t)std::max(n_edge, 1);
    CU_CHECK(d_l_pc_.ensure(n * 3));
    CU_CHECK(d_l_qw_.ensure(n * 3));
    CU_CHECK(d_l_nv_.ensure(n * 3));
    CU_CHECK(d_l_w_.ensure(n));
    CU_CHECK(d_l_acc_.ensure(LD_ACC));
    CU_CHECK(d_l_match_.ensure(n));
    lidar_.pc = d_l_pc_.p; lidar_.qw = d_l_qw_.p; lidar_.nv = d_l_nv_.p; lidar_.w = d_l_w_.p; lidar_.acc = d_l_acc_.p;
    return SQRTBA_OK;
  }
  // explicit correspondences: what the association would produce (flat edges first, weight 0 = no edge)
  int set_lidar_edges(int cur_pose, int n_flat, int n_corner, const double* pc, const double* qw, const double* normal,
                      const double* w, int numeric) {
    if (int rc = lidar_check(cur_pose, "set_lidar_edges")) return rc;
    const int n = n_flat + n_corner;
    if (n_flat < 0 || n_corner < 0 || (n > 0 && (!pc || !qw || !normal || !w))) { err_ = "set_lidar_edges: bad arrays"; return SQRTBA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(cfg_.device));
    if (int rc = lidar_alloc_edges(n)) return rc;
    if (n > 0) {
      CU_CHECK(cudaMemcpyAsync(d_l_pc_.p, pc, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, stream_));
      CU_CHECK(cudaMemcpyAsync(d_l_qw_.p, qw, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, stream_));
      CU_CHECK(cudaMemcpyAsync(d_l_nv_.p, normal, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, stream_));
      CU_CHECK(cudaMemcpyAsync(d_l_w_.p, w, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream_));
      CU_CHECK(cudaStreamSynchronize(stream_));  // the caller's arrays may go away
    }
    lidar_.n_edge = n; lidar_.n_flat = n_flat; lidar_.pose = cur_pose; lidar_.numeric = numeric ? 1 : 0;
    lidar_edges_set_ = n > 0;
    lidar_assoc_set_ = false;
    return SQRTBA_OK;
  }
  // the clouds of the pass; the association itself runs on the device inside solve_local, at the pass-2 estimates
  int set_lidar(const sqrtba_lidar* c) {
    if (!c) { err_ = "set_lidar: null"; return SQRTBA_ERR_INVALID; }
    if (int rc = lidar_check(c->cur_pose, "set_lidar")) return rc;
    if (c->n_flat < 0 || c->n_corner < 0 || c->n_map_flat < 0 || c->n_map_corner < 0 ||
        (c->n_flat > 0 && (!c->flat_xyz || !c->flat_normal)) || (c->n_corner > 0 && !c->corner_xyz) ||
        (c->n_map_flat > 0 && (!c->map_flat_xyz || !c->map_flat_pose)) ||
        (c->n_map_corner > 0 && (!c->map_corner_xyz || !c->map_corner_pose)) ||
        c->n_map_flat > 0x7fffffffLL || c->n_map_corner > 0x7fffffffLL) {
      err_ = "set_lidar: bad arrays";
      return SQRTBA_ERR_INVALID;
    }
    for (long long i = 0; i < c->n_map_flat; i++)
      if (c->map_flat_pose[i] < 0 || c->map_flat_pose[i] >= P_.n_pose) { err_ = "set_lidar: map pose index out of range"; return SQRTBA_ERR_INVALID; }
    for (long long i = 0; i < c->n_map_corner; i++)
      if (c->map_corner_pose[i] < 0 || c->map_corner_pose[i] >= P_.n_pose) { err_ = "set_lidar: map pose index out of range"; return SQRTBA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(cfg_.device));
    const int n = c->n_flat + c->n_corner;
    if (int rc = lidar_alloc_edges(n)) return rc;
    const size_t nf = (size_t)std::max(c->n_flat, 1), nc = (size_t)std::max(c->n_corner, 1);
    const size_t mf = (size_t)std::max<long long>(c->n_map_flat, 1), mc = (size_t)std::max<long long>(c->n_map_corner, 1);
    CU_CHECK(d_lc_flat_.ensure(nf * 3)); CU_CHECK(d_lc_normal_.ensure(nf * 3)); CU_CHECK(d_lc_corner_.ensure(nc * 3));
    CU_CHECK(d_lm_flat_.ensure(mf * 3)); CU_CHECK(d_lm_flat_w_.ensure(mf * 3)); CU_CHECK(d_lm_flat_pose_.ensure(mf));
    CU_CHECK(d_lm_corner_.ensure(mc * 3)); CU_CHECK(d_lm_corner_w_.ensure(mc * 3)); CU_CHECK(d_lm_corner_pose_.ensure(mc));
    CU_CHECK(d_lc_world_.ensure((size_t)std::max(n, 1) * 3)); CU_CHECK(d_l_best_.ensure((size_t)std::max(n, 1)));
    auto up = [&](void* dst, const void* src, size_t bytes) {
      return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_) : cudaSuccess;
    };
    CU_CHECK(up(d_lc_flat_.p, c->flat_xyz, (size_t)c->n_flat * 3 * sizeof(float)));
    CU_CHECK(up(d_lc_normal_.p, c->flat_normal, (size_t)c->n_flat * 3 * sizeof(float)));
    CU_CHECK(up(d_lc_corner_.p, c->corner_xyz, (size_t)c->n_corner * 3 * sizeof(float)));
    CU_CHECK(up(d_lm_flat_.p, c->map_flat_xyz, (size_t)c->n_map_flat * 3 * sizeof(float)));
    CU_CHECK(up(d_lm_flat_pose_.p, c->map_flat_pose, (size_t)c->n_map_flat * sizeof(int)));
    CU_CHECK(up(d_lm_corner_.p, c->map_corner_xyz, (size_t)c->n_map_corner * 3 * sizeof(float)));
    CU_CHECK(up(d_lm_corner_pose_.p, c->map_corner_pose, (size_t)c->n_map_corner * sizeof(int)));
    CU_CHECK(cudaStreamSynchronize(stream_));
    LidarAssoc& A = assoc_;
    A.pose = c->cur_pose; A.n_flat = c->n_flat; A.n_corner = c->n_corner;
    A.flat = d_lc_flat_.p; A.flat_n = d_lc_normal_.p; A.corner = d_lc_corner_.p;
    A.n_map_flat = c->n_map_flat; A.n_map_corner = c->n_map_corner;
    A.map_flat = d_lm_flat_.p; A.map_flat_pose = d_lm_flat_pose_.p; A.map_corner = d_lm_corner_.p; A.map_corner_pose = d_lm_corner_pose_.p;
    A.map_flat_w = d_lm_flat_w_.p; A.map_corner_w = d_lm_corner_w_.p; A.cur_w = d_lc_world_.p;
    A.best = d_l_best_.p; A.match = d_l_match_.p;
    A.thr = c->distance_sq_threshold; A.w_flat = c->flat_weight; A.w_corner = c->corner_weight;
    A.use_flat = c->use_flat ? 1 : 0; A.use_corner = c->use_corner ? 1 : 0;
    lidar_.n_edge = n; lidar_.n_flat = c->n_flat; lidar_.pose = c->cur_pose; lidar_.numeric = c->numeric_jacobian ? 1 : 0;
    lidar_assoc_set_ = n > 0;
    lidar_edges_set_ = false;
    return SQRTBA_OK;
  }
  // local lidar map + nearest-neighbour matches + edges, all at the current device estimates (g2oOptimizer.cc:981-1107)
  int lidar_associate() {
    const LidarAssoc& A = assoc_;
    const int n = A.n_flat + A.n_corner;
    if (A.n_map_flat > 0) k_lidar_to_world<<<cdiv((int)A.n_map_flat, 256), 256, 0, stream_>>>(P_, A, 0);
    if (A.n_map_corner > 0) k_lidar_to_world<<<cdiv((int)A.n_map_corner, 256), 256, 0, stream_>>>(P_, A, 1);
    k_lidar_to_world<<<cdiv(n, 256), 256, 0, stream_>>>(P_, A, 2);
    if (A.use_flat && A.n_flat > 0 && A.n_map_flat > 0)
      k_lidar_nn<<<dim3(cdiv(A.n_flat, LD_NN_CTA), cdiv((int)A.n_map_flat, LD_NN_CHUNK)), LD_NN_CTA, 0, stream_>>>(A, 0);
    if (A.use_corner && A.n_corner > 0 && A.n_map_corner > 0)
      k_lidar_nn<<<dim3(cdiv(A.n_corner, LD_NN_CTA), cdiv((int)A.n_map_corner, LD_NN_CHUNK)), LD_NN_CTA, 0, stream_>>>(A, 1);
    k_lidar_edges<<<cdiv(n, 256), 256, 0, stream_>>>(A, lidar_);
    launches_ += 6;
    CU_CHECK(cudaGetLastError());
    return SQRTBA_OK;
  }
  int get_lidar_matches(int32_t* out) {
    if (!lidar_assoc_set_ || !out) { err_ = "get_lidar_matches: no lidar clouds set"; return SQRTBA_ERR_INVALID; }
    if (download(out, d_l_match_.p, (size_t)lidar_.n_edge * sizeof(int))) return SQRTBA_ERR_CUDA;
    return lidar_.n_edge;
  }
  int num_lidar_edges() {
    if (!lidar_assoc_set_ && !lidar_edges_set_) return 0;
    std::vector<double> w((size_t)lidar_.n_edge);
    if (download(w.data(), d_l_w_.p, w.size() * sizeof(double))) return SQRTBA_ERR_CUDA;
    int n = 0;
    for (double v : w) n += v > 0.0;
    return n;
  }

  int pose_opt(int n_frames, const int64_t* frame_ptr, double* pose_qt, const double* cam, const double* obs_xyz,
               const float* obs_meas, uint8_t* outlier_out, int32_t* inliers_out, sqrtba_stats* st) {
    if (n_frames <= 0 || !frame_ptr || !pose_qt || !cam || !inliers_out) { err_ = "pose_opt: bad arguments"; return SQRTBA_ERR_INVALID; }
    const long long n_obs = frame_ptr[n_frames];
    if (frame_ptr[0] != 0 || n_obs < 0 || (n_obs > 0 && (!obs_xyz || !obs_meas || !outlier_out))) {
      err_ = "pose_opt: bad observation arrays";
      return SQRTBA_ERR_INVALID;
    }
    for (int f = 0; f < n_frames; f++)
      if (frame_ptr[f] > frame_ptr[f + 1]) { err_ = "pose_opt: frame offsets must be non-decreasing"; return SQRTBA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(cfg_.device));
    const size_t No = (size_t)std::max<long long>(n_obs, 1);
    CU_CHECK(d_po_ptr_.ensure(n_frames + 1));
    CU_CHECK(d_po_pose_.ensure((size_t)n_frames * 7));
    CU_CHECK(d_po_cam_.ensure((size_t)n_frames * 5));
    CU_CHECK(d_po_xyz_.ensure(No * 3));
    CU_CHECK(d_po_meas_.ensure(No));
    CU_CHECK(d_po_err_.ensure(No * 3));
    CU_CHECK(d_po_level_.ensure(No));
    CU_CHECK(d_po_outlier_.ensure(No));
    CU_CHECK(d_po_inl_.ensure(2 * (size_t)n_frames));
    CU_CHECK(d_po_trace_.ensure((size_t)n_frames * PO_MAX_TRACE * PO_TRACE_COLS));
    auto up = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_); };
    static_assert(sizeof(long long) == sizeof(int64_t), "frame offsets");
    CU_CHECK(cudaEventRecord(ev0_, stream_));
    CU_CHECK(up(d_po_ptr_.p, frame_ptr, (size_t)(n_frames + 1) * sizeof(int64_t)));
    CU_CHECK(up(d_po_pose_.p, pose_qt, (size_t)n_frames * 7 * sizeof(double)));
    CU_CHECK(up(d_po_cam_.p, cam, (size_t)n_frames * 5 * sizeof(double)));
    if (n_obs > 0) {
      CU_CHECK(up(d_po_xyz_.p, obs_xyz, (size_t)n_obs * 3 * sizeof(double)));
      CU_CHECK(up(d_po_meas_.p, obs_meas, (size_t)n_obs * sizeof(float4)));
    }
    PoseOptArgs A{};
    A.n_frames = n_frames; A.frame_ptr = d_po_ptr_.p; A.pose = d_po_pose_.p; A.cam = d_po_cam_.p; A.xyz = d_po_xyz_.p;
    A.meas = d_po_meas_.p; A.err = d_po_err_.p; A.level = d_po_level_.p; A.outlier = d_po_outlier_.p;
    A.inliers = d_po_inl_.p; A.trace = d_po_trace_.p; A.trace_len = d_po_inl_.p + n_frames;
    k_pose_opt<<<n_frames, PO_CTA, 0, stream_>>>(A);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaMemcpyAsync(pose_qt, d_po_pose_.p, (size_t)n_frames * 7 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaMemcpyAsync(inliers_out, d_po_inl_.p, (size_t)n_frames * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    if (n_obs > 0) CU_CHECK(cudaMemcpyAsync(outlier_out, d_po_outlier_.p, (size_t)n_obs, cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaEventRecord(ev1_, stream_));
    CU_CHECK(cudaEventSynchronize(ev1_));
    po_frames_ = n_frames;
    s3_pairs_ = 0;  // shared buffers
    if (st) {
      std::memset(st, 0, sizeof *st);
      float ms = 0;
      CU_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
      st->n_windows = n_frames;
      st->kernel_launches = 1;
      st->ms_total = ms;
    }
    return SQRTBA_OK;
  }
  // ------------------------------------------------------------------------ pose-only optimisation with the lidar block
  // g2oOptimizer::PoseOptimization as the tracking thread of this fork calls it (g2oOptimizer.cc:385-690, lidar block
  // :560-640): the four visual rounds (k_pose_opt), then -- when the local lidar map has more than 100 points -- the
  // association at the new pose, the float round trip of the estimate and a fifth optimize(10) over visual + lidar edges
  // (csrc/sqrtba_poseopt_lidar.cuh).  One frame per call; independent of set_problem.
  int pose_opt_lidar(double* pose_qt, const double* cam, int n_obs, const double* obs_xyz, const float* obs_meas,
                     uint8_t* outlier_out, int32_t* inliers_out, const sqrtba_frame_lidar* c, int32_t* n_match2_out,
                     sqrtba_stats* st) {
    if (!c) { err_ = "pose_opt_lidar: null lidar block"; return SQRTBA_ERR_INVALID; }
    if (c->n_flat < 0 || c->n_corner < 0 || c->n_map < 0 || c->n_map > 0x7fffffffLL ||
        (c->n_flat > 0 && (!c->flat_xyz || !c->flat_normal)) || (c->n_corner > 0 && !c->corner_xyz) || (c->n_map > 0 && !c->map_xyz)) {
      err_ = "pose_opt_lidar: bad lidar arrays";
      return SQRTBA_ERR_INVALID;
    }
    if (n_match2_out) n_match2_out[0] = n_match2_out[1] = 0;
    const int64_t ptr[2] = {0, n_obs};
    const bool lidar_round = c->n_map > 100 && n_obs >= 3;  // :491-492 returns before the lidar block; :560 needs > 100 map points
    if (!lidar_round) return pose_opt(1, ptr, pose_qt, cam, obs_xyz, obs_meas, outlier_out, inliers_out, st);
    if (!pose_qt || !cam || !inliers_out || !obs_xyz || !obs_meas || !outlier_out) { err_ = "pose_opt_lidar: bad arguments"; return SQRTBA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(cfg_.device));
    const size_t No = (size_t)n_obs;
    CU_CHECK(d_po_ptr_.ensure(2));
    CU_CHECK(d_po_pose_.ensure(8));
    CU_CHECK(d_po_cam_.ensure(8));
    CU_CHECK(d_po_xyz_.ensure(No * 3));
    CU_CHECK(d_po_meas_.ensure(No));
    CU_CHECK(d_po_err_.ensure(No * 3));
    CU_CHECK(d_po_level_.ensure(No));
    CU_CHECK(d_po_outlier_.ensure(No));
    CU_CHECK(d_po_inl_.ensure(4));
    CU_CHECK(d_po_trace_.ensure((size_t)PO_MAX_TRACE * PO_TRACE_COLS));
    const int n = c->n_flat + c->n_corner;
    const size_t ne = (size_t)std::max(n, 1);
    CU_CHECK(d_fl_pc_.ensure(ne * 3)); CU_CHECK(d_fl_qw_.ensure(ne * 3)); CU_CHECK(d_fl_nv_.ensure(ne * 3)); CU_CHECK(d_fl_w_.ensure(ne));
    CU_CHECK(d_fl_match_.ensure(ne)); CU_CHECK(d_fl_best_.ensure(ne)); CU_CHECK(d_fl_world_.ensure(ne * 3));
    CU_CHECK(d_fl_flat_.ensure((size_t)std::max(c->n_flat, 1) * 3)); CU_CHECK(d_fl_normal_.ensure((size_t)std::max(c->n_flat, 1) * 3));
    CU_CHECK(d_fl_corner_.ensure((size_t)std::max(c->n_corner, 1) * 3)); CU_CHECK(d_fl_map_.ensure((size_t)c->n_map * 3));
    auto up = [&](void* dst, const void* src, size_t bytes) {
      return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_) : cudaSuccess;
    };
    CU_CHECK(cudaEventRecord(ev0_, stream_));
    CU_CHECK(up(d_po_ptr_.p, ptr, sizeof ptr));
    CU_CHECK(up(d_po_pose_.p, pose_qt, 7 * sizeof(double)));
    CU_CHECK(up(d_po_cam_.p, cam, 5 * sizeof(double)));
    CU_CHECK(up(d_po_xyz_.p, obs_xyz, No * 3 * sizeof(double)));
    CU_CHECK(up(d_po_meas_.p, obs_meas, No * sizeof(float4)));
    CU_CHECK(up(d_fl_flat_.p, c->flat_xyz, (size_t)c->n_flat * 3 * sizeof(float)));
    CU_CHECK(up(d_fl_normal_.p, c->flat_normal, (size_t)c->n_flat * 3 * sizeof(float)));
    CU_CHECK(up(d_fl_corner_.p, c->corner_xyz, (size_t)c->n_corner * 3 * sizeof(float)));
    CU_CHECK(up(d_fl_map_.p, c->map_xyz, (size_t)c->n_map * 3 * sizeof(float)));
    PoseOptArgs A{};
    A.n_frames = 1; A.frame_ptr = d_po_ptr_.p; A.pose = d_po_pose_.p; A.cam = d_po_cam_.p; A.xyz = d_po_xyz_.p;
    A.meas = d_po_meas_.p; A.err = d_po_err_.p; A.level = d_po_level_.p; A.outlier = d_po_outlier_.p;
    A.inliers = d_po_inl_.p; A.trace = d_po_trace_.p; A.trace_len = d_po_inl_.p + 1; A.skip_final = 1;
    k_pose_opt<<<1, PO_CTA, 0, stream_>>>(A);
    // association at the pose of the four rounds: both point kinds against the one local map cloud (world frame already)
    LidarAssoc S{};
    S.pose = 0; S.n_flat = c->n_flat; S.n_corner = c->n_corner;
    S.flat = d_fl_flat_.p; S.flat_n = d_fl_normal_.p; S.corner = d_fl_corner_.p;
    S.n_map_flat = S.n_map_corner = c->n_map;
    S.map_flat_w = S.map_corner_w = d_fl_map_.p;
    S.cur_w = d_fl_world_.p; S.best = d_fl_best_.p; S.match = d_fl_match_.p;
    S.thr = c->distance_sq_threshold; S.w_flat = c->flat_weight; S.w_corner = c->corner_weight;
    S.use_flat = c->use_flat ? 1 : 0; S.use_corner = c->use_corner ? 1 : 0;
    LidarDev L{};
    L.n_edge = n; L.n_flat = c->n_flat; L.pose = 0; L.numeric = 1;
    L.pc = d_fl_pc_.p; L.qw = d_fl_qw_.p; L.nv = d_fl_nv_.p; L.w = d_fl_w_.p; L.acc = nullptr;
    int launched = 1;
    if (n > 0) {
      Dev Pf{};
      Pf.pose = d_po_pose_.p;
      k_lidar_to_world<<<cdiv(n, 256), 256, 0, stream_>>>(Pf, S, 2);
      if (S.use_flat && S.n_flat > 0) {
        k_lidar_nn<<<dim3(cdiv(S.n_flat, LD_NN_CTA), cdiv((int)c->n_map, LD_NN_CHUNK)), LD_NN_CTA, 0, stream_>>>(S, 0);
        launched++;
      }
      if (S.use_corner && S.n_corner > 0) {
        k_lidar_nn<<<dim3(cdiv(S.n_corner, LD_NN_CTA), cdiv((int)c->n_map, LD_NN_CHUNK)), LD_NN_CTA, 0, stream_>>>(S, 1);
        launched++;
      }
      k_lidar_edges<<<cdiv(n, 256), 256, 0, stream_>>>(S, L);
      launched += 2;
    }
    k_pose_opt_lidar<<<1, PO_CTA, 0, stream_>>>(A, L, d_po_inl_.p + 2);
    launched++;
    CU_CHECK(cudaGetLastError());
    int32_t nm[2] = {0, 0};
    CU_CHECK(cudaMemcpyAsync(pose_qt, d_po_pose_.p, 7 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaMemcpyAsync(inliers_out, d_po_inl_.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaMemcpyAsync(nm, d_po_inl_.p + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaMemcpyAsync(outlier_out, d_po_outlier_.p, No, cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaEventRecord(ev1_, stream_));
    CU_CHECK(cudaEventSynchronize(ev1_));
    if (n_match2_out) { n_match2_out[0] = nm[0]; n_match2_out[1] = nm[1]; }
    po_frames_ = 1;
    s3_pairs_ = 0;
    if (st) {
      std::memset(st, 0, sizeof *st);
      float ms = 0;
      CU_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
      st->n_windows = 1;
      st->kernel_launches = launched;
      st->ms_total = ms;
    }
    return SQRTBA_OK;
  }
  int pose_opt_trace(int frame, double* rows, int max_rows) {
    if (frame < 0 || frame >= po_frames_ || !rows) { err_ = "pose_opt_trace: no such frame"; return SQRTBA_ERR_INVALID; }
    int len = 0;
    if (download(&len, d_po_inl_.p + po_frames_ + frame, sizeof(int))) return SQRTBA_ERR_CUDA;
    const int m = std::min(std::min(len, max_rows), PO_MAX_TRACE);
    if (m > 0 && download(rows, d_po_trace_.p + (size_t)frame * PO_MAX_TRACE * PO_TRACE_COLS, (size_t)m * PO_TRACE_COLS * sizeof(double)))
      return SQRTBA_ERR_CUDA;
    return m;
  }


  // -------------------------------------------------------------------------------------- Sim3 of a loop candidate (row N3)
  // g2oOptimizer::OptimizeSim3 (g2oOptimizer.cc:1560-1796) for a batch of keyframe pairs, one CTA each
  // (csrc/sqrtba_sim3opt.cuh).  Independent of set_problem.
  int optimize_sim3(int n_pairs, const int64_t* pair_ptr, double* s12, const double* cam8, const double* p1c,
                    const double* p2c, const float* meas6, float th2, int fix_scale, uint8_t* keep_out, int32_t* n_in_out,
                    sqrtba_stats* st) {
    if (n_pairs <= 0 || !pair_ptr || !s12 || !cam8 || !n_in_out || !(th2 > 0.0f)) { err_ = "optimize_sim3: bad arguments"; return SQRTBA_ERR_INVALID; }
    const long long n_m = pair_ptr[n_pairs];
    if (pair_ptr[0] != 0 || n_m < 0 || (n_m > 0 && (!p1c || !p2c || !meas6 || !keep_out))) {
      err_ = "optimize_sim3: bad match arrays";
      return SQRTBA_ERR_INVALID;
    }
    for (int f = 0; f < n_pairs; f++)
      if (pair_ptr[f] > pair_ptr[f + 1]) { err_ = "optimize_sim3: pair offsets must be non-decreasing"; return SQRTBA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(cfg_.device));
    const size_t Nm = (size_t)std::max<long long>(n_m, 1);
    CU_CHECK(d_po_ptr_.ensure(n_pairs + 1));
    CU_CHECK(d_po_pose_.ensure((size_t)n_pairs * 8));
    CU_CHECK(d_po_cam_.ensure((size_t)n_pairs * 8));
    CU_CHECK(d_po_xyz_.ensure(Nm * 6));
    CU_CHECK(d_s3_meas_.ensure(Nm * 6));
    CU_CHECK(d_po_err_.ensure(Nm * 4));
    CU_CHECK(d_po_level_.ensure(Nm));
    CU_CHECK(d_po_outlier_.ensure(Nm));
    CU_CHECK(d_po_inl_.ensure(2 * (size_t)n_pairs));
    CU_CHECK(d_po_trace_.ensure((size_t)n_pairs * S3_MAX_TRACE * S3_TRACE_COLS));
    auto up = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_); };
    CU_CHECK(cudaEventRecord(ev0_, stream_));
    CU_CHECK(up(d_po_ptr_.p, pair_ptr, (size_t)(n_pairs + 1) * sizeof(int64_t)));
    CU_CHECK(up(d_po_pose_.p, s12, (size_t)n_pairs * 8 * sizeof(double)));
    CU_CHECK(up(d_po_cam_.p, cam8, (size_t)n_pairs * 8 * sizeof(double)));
    if (n_m > 0) {
      CU_CHECK(up(d_po_xyz_.p, p1c, (size_t)n_m * 3 * sizeof(double)));
      CU_CHECK(up(d_po_xyz_.p + Nm * 3, p2c, (size_t)n_m * 3 * sizeof(double)));
      CU_CHECK(up(d_s3_meas_.p, meas6, (size_t)n_m * 6 * sizeof(float)));
    }
    Sim3OptArgs A{};
    A.n_pairs = n_pairs; A.pair_ptr = d_po_ptr_.p; A.s12 = d_po_pose_.p; A.cam8 = d_po_cam_.p; A.p1c = d_po_xyz_.p;
    A.p2c = d_po_xyz_.p + Nm * 3; A.meas6 = d_s3_meas_.p; A.err = d_po_err_.p; A.on = d_po_level_.p; A.keep = d_po_outlier_.p;
    A.n_in = d_po_inl_.p; A.trace = d_po_trace_.p; A.trace_len = d_po_inl_.p + n_pairs; A.th2 = th2; A.fix_scale = fix_scale;
    k_sim3_opt<<<n_pairs, PO_CTA, 0, stream_>>>(A);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaMemcpyAsync(s12, d_po_pose_.p, (size_t)n_pairs * 8 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaMemcpyAsync(n_in_out, d_po_inl_.p, (size_t)n_pairs * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    if (n_m > 0) CU_CHECK(cudaMemcpyAsync(keep_out, d_po_outlier_.p, (size_t)n_m, cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaEventRecord(ev1_, stream_));
    CU_CHECK(cudaEventSynchronize(ev1_));
    s3_pairs_ = n_pairs;
    po_frames_ = 0;  // the pose-only trace buffers were reused
    if (st) {
      std::memset(st, 0, sizeof *st);
      float ms = 0;
      CU_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
      st->n_windows = n_pairs;
      st->kernel_launches = 1;
      st->ms_total = ms;
    }
    return SQRTBA_OK;
  }
  int optimize_sim3_trace(int pair, double* rows, int max_rows) {
    if (pair < 0 || pair >= s3_pairs_ || !rows) { err_ = "optimize_sim3_trace: no such pair"; return SQRTBA_ERR_INVALID; }
    int len = 0;
    if (download(&len, d_po_inl_.p + s3_pairs_ + pair, sizeof(int))) return SQRTBA_ERR_CUDA;
    const int m = std::min(std::min(len, max_rows), S3_MAX_TRACE);
    if (m > 0 && download(rows, d_po_trace_.p + (size_t)pair * S3_MAX_TRACE * S3_TRACE_COLS, (size_t)m * S3_TRACE_COLS * sizeof(double)))
      return SQRTBA_ERR_CUDA;
    return m;
  }

  // ------------------------------------------------------------------------------------------ essential graph (row N3)
  // The optimisation of g2oOptimizer::OptimizeEssentialGraph (g2oOptimizer.cc:1212-1460) for a pose graph the adapter
  // has built: Levenberg with lambda_0 = lambda_init (setUserLambdaInit, 1e-16 in the reference; <= 0: tau * max diag),
  // `iters` iterations, exact block-skyline Cholesky per trial (csrc/sqrtba_posegraph.cuh).  Independent of set_problem.
  int pose_graph(int n_vert, double* vert8, const uint8_t* fixed, int fix_scale, int n_edge, const int32_t* edge_ij,
                 const double* meas8, int iters, double lambda_init, sqrtba_stats* st) {
    if (n_vert <= 0 || n_edge < 0 || !vert8 || !fixed || (n_edge > 0 && (!edge_ij || !meas8)) || iters < 0) {
      err_ = "pose_graph: bad arguments";
      return SQRTBA_ERR_INVALID;
    }
    for (int k = 0; k < 2 * n_edge; k++)
      if (edge_ij[k] < 0 || edge_ij[k] >= n_vert) { err_ = "pose_graph: edge vertex out of range"; return SQRTBA_ERR_INVALID; }
    for (int k = 0; k < n_edge; k++)
      if (edge_ij[2 * k] == edge_ij[2 * k + 1]) { err_ = "pose_graph: an edge joins a vertex with itself"; return SQRTBA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(cfg_.device));
    pg_trace_.clear();
    // index mapping: non-fixed vertices that have an edge, ascending id (sparse_optimizer.cpp:166-190, 482-487)
    std::vector<int> slot(n_vert, -1), slot_vert;
    {
      std::vector<uint8_t> act(n_vert, 0);
      for (int k = 0; k < 2 * n_edge; k++) act[edge_ij[k]] = 1;
      for (int i = 0; i < n_vert; i++)
        if (act[i] && !fixed[i]) { slot[i] = (int)slot_vert.size(); slot_vert.push_back(i); }
    }
    const int n_slot = (int)slot_vert.size();
    if (st) std::memset(st, 0, sizeof *st);
    if (n_slot == 0 || iters == 0) return SQRTBA_OK;  // nothing to optimise: estimates untouched
    // block skyline: row r stores the block columns [first[r], r]
    std::vector<int> first(n_slot);
    for (int r = 0; r < n_slot; r++) first[r] = r;
    for (int k = 0; k < n_edge; k++) {
      const int a = slot[edge_ij[2 * k]], b = slot[edge_ij[2 * k + 1]];
      if (a < 0 || b < 0) continue;
      const int hi = std::max(a, b), lo = std::min(a, b);
      first[hi] = std::min(first[hi], lo);
    }
    std::vector<long long> rowptr(n_slot + 1, 0);
    for (int r = 0; r < n_slot; r++) rowptr[r + 1] = rowptr[r] + (r - first[r] + 1);
    const long long n_blocks = rowptr[n_slot];
    if ((size_t)n_slot * 7 * sizeof(double) > 160 * 1024) {  // the solve keeps its work vector in shared memory
      err_ = "pose_graph: more than 2900 free keyframes (shared-memory work vector of the triangular solves)";
      return SQRTBA_ERR_INVALID;
    }
    if (n_blocks > (1ll << 23)) {  // 8M blocks = 3.3 GB per copy: the natural (keyframe id) order does not fit this graph
      err_ = "pose_graph: the block skyline of this graph in keyframe-id order is too large";
      return SQRTBA_ERR_INVALID;
    }
    std::vector<int> col_ptr(n_slot + 1, 0), col_rows((size_t)(n_blocks - n_slot));
    for (int r = 0; r < n_slot; r++)
      for (int c = first[r]; c < r; c++) col_ptr[c + 1]++;
    for (int c = 0; c < n_slot; c++) col_ptr[c + 1] += col_ptr[c];
    {
      std::vector<int> cur(col_ptr.begin(), col_ptr.end() - 1);
      for (int r = 0; r < n_slot; r++)   // ascending r: every column's row list comes out sorted
        for (int c = first[r]; c < r; c++) col_rows[(size_t)cur[c]++] = r;
    }
    // the edges of every free vertex in ascending edge index (deterministic assembly, no atomics)
    std::vector<int> inc_ptr(n_slot + 1, 0), inc_edge;
    for (int k = 0; k < 2 * n_edge; k++)
      if (slot[edge_ij[k]] >= 0) inc_ptr[slot[edge_ij[k]] + 1]++;
    for (int r = 0; r < n_slot; r++) inc_ptr[r + 1] += inc_ptr[r];
    inc_edge.resize((size_t)inc_ptr[n_slot]);
    {
      std::vector<int> cur(inc_ptr.begin(), inc_ptr.end() - 1);
      for (int k = 0; k < n_edge; k++)
        for (int side = 0; side < 2; side++) {
          const int sl = slot[edge_ij[2 * k + side]];
          if (sl >= 0) inc_edge[(size_t)cur[sl]++] = 2 * k + side;
        }
    }
    const int n_err_cta = std::max(cdiv(n_edge, 256), 1);
    const size_t Nv = n_vert, Ne = (size_t)std::max(n_edge, 1), Ns = n_slot, Nb = (size_t)n_blocks;
    CU_CHECK(d_pg_inc_ptr_.ensure(Ns + 1)); CU_CHECK(d_pg_inc_edge_.ensure(std::max<size_t>(inc_edge.size(), 1)));
    CU_CHECK(d_pg_chi_part_.ensure(n_err_cta));
    CU_CHECK(d_pg_vert_.ensure(Nv * 8)); CU_CHECK(d_pg_bak_.ensure(Nv * 8)); CU_CHECK(d_pg_fixed_.ensure(Nv));
    CU_CHECK(d_pg_slot_.ensure(Nv)); CU_CHECK(d_pg_slot_vert_.ensure(Ns)); CU_CHECK(d_pg_edge_.ensure(Ne * 2));
    CU_CHECK(d_pg_meas_.ensure(Ne * 8)); CU_CHECK(d_pg_err_.ensure(Ne * 7)); CU_CHECK(d_pg_Ji_.ensure(Ne * 49));
    CU_CHECK(d_pg_Jj_.ensure(Ne * 49)); CU_CHECK(d_pg_first_.ensure(Ns)); CU_CHECK(d_pg_rowptr_.ensure(Ns + 1));
    CU_CHECK(d_pg_H_.ensure(Nb * 49)); CU_CHECK(d_pg_L_.ensure(Nb * 49)); CU_CHECK(d_pg_Linv_.ensure(Ns * 49));
    CU_CHECK(d_pg_col_ptr_.ensure(Ns + 1)); CU_CHECK(d_pg_col_rows_.ensure(std::max<size_t>(col_rows.size(), 1)));
    CU_CHECK(d_pg_b_.ensure(Ns * 7)); CU_CHECK(d_pg_y_.ensure(Ns * 7)); CU_CHECK(d_pg_x_.ensure(Ns * 7)); CU_CHECK(d_pg_scal_.ensure(4));
    auto up = [&](void* dst, const void* src, size_t bytes) {
      return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_) : cudaSuccess;
    };
    CU_CHECK(cudaEventRecord(ev0_, stream_));
    CU_CHECK(up(d_pg_vert_.p, vert8, Nv * 8 * sizeof(double)));
    CU_CHECK(up(d_pg_fixed_.p, fixed, Nv));
    CU_CHECK(up(d_pg_slot_.p, slot.data(), Nv * sizeof(int)));
    CU_CHECK(up(d_pg_slot_vert_.p, slot_vert.data(), Ns * sizeof(int)));
    CU_CHECK(up(d_pg_edge_.p, edge_ij, (size_t)n_edge * 2 * sizeof(int)));
    CU_CHECK(up(d_pg_meas_.p, meas8, (size_t)n_edge * 8 * sizeof(double)));
    CU_CHECK(up(d_pg_first_.p, first.data(), Ns * sizeof(int)));
    CU_CHECK(up(d_pg_rowptr_.p, rowptr.data(), (Ns + 1) * sizeof(long long)));
    CU_CHECK(up(d_pg_col_ptr_.p, col_ptr.data(), (Ns + 1) * sizeof(int)));
    CU_CHECK(up(d_pg_col_rows_.p, col_rows.data(), col_rows.size() * sizeof(int)));
    CU_CHECK(up(d_pg_inc_ptr_.p, inc_ptr.data(), (Ns + 1) * sizeof(int)));
    CU_CHECK(up(d_pg_inc_edge_.p, inc_edge.data(), inc_edge.size() * sizeof(int)));
    PgDev G{};
    G.n_vert = n_vert; G.n_edge = n_edge; G.n_slot = n_slot; G.fix_scale = fix_scale ? 1 : 0;
    G.vert = d_pg_vert_.p; G.vert_bak = d_pg_bak_.p; G.fixed = d_pg_fixed_.p; G.slot = d_pg_slot_.p; G.slot_vert = d_pg_slot_vert_.p;
    G.edge_ij = d_pg_edge_.p; G.meas = d_pg_meas_.p; G.err = d_pg_err_.p; G.Ji = d_pg_Ji_.p; G.Jj = d_pg_Jj_.p;
    G.first = d_pg_first_.p; G.rowptr = d_pg_rowptr_.p; G.H = d_pg_H_.p; G.L = d_pg_L_.p; G.Linv = d_pg_Linv_.p;
    G.col_ptr = d_pg_col_ptr_.p; G.col_rows = d_pg_col_rows_.p; G.b = d_pg_b_.p; G.y = d_pg_y_.p; G.x = d_pg_x_.p; G.scal = d_pg_scal_.p;
    G.chi_part = d_pg_chi_part_.p; G.inc_ptr = d_pg_inc_ptr_.p; G.inc_edge = d_pg_inc_edge_.p;
    int launches = 0;
    double h_scal[4];
    std::vector<double> h_part(n_err_cta, 0.0);
    // computeActiveErrors + activeChi2: per-CTA partial sums, added up in order here (reproducible); with_scal also
    // reads the solve's scalars in the same synchronisation
    auto chi2_now = [&](double* out, bool with_scal) -> int {
      if (n_edge > 0) { k_pg_errors<<<n_err_cta, 256, 0, stream_>>>(G); launches++; }
      else CU_CHECK(cudaMemsetAsync(G.chi_part, 0, sizeof(double), stream_));
      CU_CHECK(cudaMemcpyAsync(h_part.data(), G.chi_part, (size_t)n_err_cta * sizeof(double), cudaMemcpyDeviceToHost, stream_));
      if (with_scal) CU_CHECK(cudaMemcpyAsync(h_scal, G.scal, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
      CU_CHECK(cudaStreamSynchronize(stream_));
      CU_CHECK(cudaGetLastError());
      double c = 0.0;
      for (double v : h_part) c += v;
      *out = c;
      return SQRTBA_OK;
    };
    double lambda = 0.0, ni = 2.0;
    int nbad = 0;
    bool ok = true;
    for (int it = 0; it < iters && ok; it++) {  // SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve
      double currentChi = 0.0;
      if (int rc = chi2_now(&currentChi, false)) return rc;
      double tempChi = currentChi;
      const double iniChi = currentChi;
      CU_CHECK(cudaMemsetAsync(G.H, 0, Nb * 49 * sizeof(double), stream_));
      if (n_edge > 0) { k_pg_linearize<<<cdiv((long long)n_edge * 14, 128), 128, 0, stream_>>>(G); launches++; }
      k_pg_assemble<<<cdiv((long long)n_slot * 49, 256), 256, 0, stream_>>>(G);
      launches++;
      if (it == 0) {
        if (lambda_init > 0) {
          lambda = lambda_init;  // computeLambdaInit returns _userLambdaInit (optimization_algorithm_levenberg.cpp:168-169)
        } else {                // tau * max diag(H)
          std::vector<double> Hh(Nb * 49);
          if (download(Hh.data(), G.H, Hh.size() * sizeof(double))) return SQRTBA_ERR_CUDA;
          double md = 0.0;
          for (int r = 0; r < n_slot; r++)
            for (int a = 0; a < 7; a++) md = std::max(md, std::fabs(Hh[(size_t)(rowptr[r] + (r - first[r])) * 49 + a * 8]));
          lambda = 1e-5 * md;
        }
        ni = 2.0;
        nbad = 0;
      }
      double rho = 0.0;
      int qmax = 0;
      do {
        k_pg_prepare<<<cdiv(std::max<long long>(n_blocks * 49, (long long)n_slot * 7), 256), 256, 0, stream_>>>(G, n_blocks);
        k_pg_damp<<<cdiv(n_slot * 7, 256), 256, 0, stream_>>>(G, lambda);
        static const bool pg_timing = std::getenv("SQRTBA_HOST_TIMING") != nullptr;
        if (pg_timing) cudaEventRecord(stage_ev_[0], stream_);
        k_pg_factor_solve<<<1, PG_THREADS, (size_t)n_slot * 7 * sizeof(double), stream_>>>(G, lambda);
        if (pg_timing) {
          cudaEventRecord(stage_ev_[1], stream_);
          cudaEventSynchronize(stage_ev_[1]);
          float fms = 0;
          cudaEventElapsedTime(&fms, stage_ev_[0], stage_ev_[1]);
          std::fprintf(stderr, "[sqrtba host] pose graph: factor + solve %8.3f ms (%d block rows, %lld blocks)\n", fms, n_slot, n_blocks);
        }
        k_pg_update<<<cdiv(n_slot, 128), 128, 0, stream_>>>(G);   // push + update
        launches += 4;
        if (int rc = chi2_now(&tempChi, true)) return rc;
        if (h_scal[2] != 0.0) tempChi = std::numeric_limits<double>::max();  // Cholesky failure => the step is rejected
        rho = currentChi - tempChi;
        const double scale = h_scal[1] + 1e-3;
        rho /= scale;
        double row[8] = {0.0, (double)it, (double)qmax, lambda, currentChi, tempChi, rho, 0.0};
        if (rho > 0 && std::isfinite(tempChi)) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, 2. / 3.);
          lambda *= std::max(1. / 3., alpha);
          ni = 2;
          currentChi = tempChi;
          row[7] = 1.0;
        } else {
          lambda *= ni;
          ni *= 2;
          k_pg_restore<<<cdiv(n_slot, 128), 128, 0, stream_>>>(G);  // pop
          launches++;
        }
        pg_trace_.insert(pg_trace_.end(), row, row + 8);
        qmax++;
      } while (rho < 0 && qmax < 10);
      if (qmax == 10 || rho == 0) { ok = false; break; }
      if ((iniChi - currentChi) * 1e3 < iniChi) nbad++; else nbad = 0;
      if (nbad >= 3) ok = false;
    }
    CU_CHECK(cudaMemcpyAsync(vert8, G.vert, Nv * 8 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaEventRecord(ev1_, stream_));
    CU_CHECK(cudaEventSynchronize(ev1_));
    CU_CHECK(cudaGetLastError());
    if (st) {
      float ms = 0;
      CU_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
      st->n_windows = 1;
      st->lm_trials = (int)(pg_trace_.size() / 8);
      st->kernel_launches = launches;
      st->ms_total = ms;
      st->reserved[3] = (double)n_blocks;  // 7x7 blocks of the skyline
      st->reserved[4] = (double)n_slot;
    }
    return SQRTBA_OK;
  }
  int pose_graph_trace(double* rows, int max_rows) {
    const int n = (int)(pg_trace_.size() / 8), m = std::min(n, std::max(max_rows, 0));
    if (m > 0 && rows) std::memcpy(rows, pg_trace_.data(), (size_t)m * 8 * sizeof(double));
    return rows ? m : n;
  }

  // ------------------------------------------------------------------------------------------ stage-level entry points
  // the stage-level entry points below use the kernels' atomic flushes (they launch single stages out of sequence)
  struct DetOff {
    Dev& P; int saved;
    explicit DetOff(Dev& p) : P(p), saved(p.det) { P.det = 0; }
    ~DetOff() { P.det = saved; }
  };
  int debug_linearize(int huber, double* err, double* Jp, double* Jl, double* r, double* chi2) {
    if (!have_problem_) { err_ = "no problem set"; return SQRTBA_ERR_INVALID; }
    DetOff det_off(P_);
    CU_CHECK(cudaSetDevice(cfg_.device));
    const double d2 = (double)(float)std::sqrt(huber == 2 ? 5.99 : 5.991), d3 = (double)(float)std::sqrt(7.815);
    k_pass_init<<<cdiv(P_.n_win, 128), 128, 0, stream_>>>(P_, 1, 0, huber != 0 ? 1 : 0, eff_rtol() * eff_rtol());
    if (P_.n_slot) k_zero_lin<<<cdiv(P_.n_slot, 128), 128, 0, stream_>>>(P_);
    launch_linearize(huber != 0, d2, d3, 1);
    if (int rc = begin_after_linearize()) return rc;
    CU_CHECK(cudaGetLastError());
    const size_t No = P_.n_obs, Ld = P_.ld;
    std::vector<double> tmp;
    auto planes = [&](double* dst, const double* dsrc, int np) -> int {
      if (!dst) return 0;
      tmp.resize(Ld * np);
      if (download(tmp.data(), dsrc, Ld * np * sizeof(double))) return SQRTBA_ERR_CUDA;
      std::vector<double> rows;
      double* d = dst;
      if (!perm_.empty()) { rows.resize(No * np); d = rows.data(); }
      for (size_t o = 0; o < No; o++)
        for (int c = 0; c < np; c++) d[o * np + c] = tmp[(size_t)c * Ld + o];
      if (!perm_.empty()) unperm_obs(rows.data(), dst, np);
      return 0;
    };
    if (planes(err, d_err_.p, 3) || planes(Jl, d_Jl_.p, 9) || planes(r, d_r_.p, 3)) return SQRTBA_ERR_CUDA;
    if (Jp) {  // un-block the tile-blocked matvec operand
      tmp.resize(d_JQ_.cap);
      if (download(tmp.data(), d_JQ_.p, d_JQ_.cap * sizeof(double))) return SQRTBA_ERR_CUDA;
      std::vector<double> rows;
      double* d = Jp;
      if (!perm_.empty()) { rows.resize(No * 18); d = rows.data(); }
      for (const TileInfo& ti : h_tiles_) {  // fixed-keyframe observations have no pose columns: zeros
        int fcol = 0;
        for (int o = ti.o0; o < ti.o1; o++) {
          const bool free_pose = h_obs_slot_.p[o] >= 0;
          const int col = ti.is_long ? (o - ti.o0) : fcol;
          double Jp[18];
          for (int c = 0; c < 18; c++) Jp[c] = 0.0;
          if (free_pose) {  // the block stores {x/z, y/z, 1/z, w}; the 3x6 Jacobian is rebuilt the way the kernels do
            double g[JG];
            for (int c = 0; c < JG; c++) g[c] = tmp[(size_t)ti.jq_off + (size_t)c * ti.nt + col];
            const double* cm = &h_slot_cam_[(size_t)h_obs_slot_.p[o] * 3];
            const bool stereo = (h_obs_lp_.p[o] & LP_STEREO) != 0;
            jp_full(jp_compact(g, cm[0], cm[1], cm[2], stereo), stereo, Jp);
          }
          for (int c = 0; c < 18; c++) d[(size_t)o * 18 + c] = Jp[c];
          fcol += free_pose;
        }
      }
      if (!perm_.empty()) unperm_obs(rows.data(), Jp, 18);
    }
    if (chi2) {
      std::vector<WinCtl> c(P_.n_win);
      if (download(c.data(), d_ctl_.p, c.size() * sizeof(WinCtl))) return SQRTBA_ERR_CUDA;
      for (int w = 0; w < P_.n_win; w++) chi2[w] = c[w].cur_chi;
    }
    return SQRTBA_OK;
  }

  // one damped square-root step at the linearisation left by debug_linearize (or the last solve)
  int debug_step(double lambda, double* dp, double* dl, double* bs, int* cg_iters) {
    if (!have_problem_) { err_ = "no problem set"; return SQRTBA_ERR_INVALID; }
    DetOff det_off(P_);
    CU_CHECK(cudaSetDevice(cfg_.device));
    // put every window into PH_TRIAL with the requested lambda
    std::vector<WinCtl> c(P_.n_win);
    if (download(c.data(), d_ctl_.p, c.size() * sizeof(WinCtl))) return SQRTBA_ERR_CUDA;
    for (auto& w : c) { w.phase = PH_TRIAL; w.lambda = lambda; w.iter = 0; w.qmax = 0; w.tol2 = eff_rtol() * eff_rtol(); }  // first trial of a pass: the separate QR kernel
    CU_CHECK(cudaMemcpyAsync(d_ctl_.p, c.data(), c.size() * sizeof(WinCtl), cudaMemcpyHostToDevice, stream_));
    CU_CHECK(cudaStreamSynchronize(stream_));
    int rc = factor_and_solve();
    if (rc) return rc;
    k_backsub<<<P_.n_tile, CTA, 0, stream_>>>(P_, 0, 0.0);
    CU_CHECK(cudaGetLastError());
    if (dp && P_.n_slot && download(dp, P_.x, (size_t)P_.n_slot * 6 * sizeof(double))) return SQRTBA_ERR_CUDA;
    if (bs && P_.n_slot && download(bs, P_.bs, (size_t)P_.n_slot * 6 * sizeof(double))) return SQRTBA_ERR_CUDA;
    if (dl) {
      std::vector<double> tmp((size_t)P_.n_point * 3);
      if (download(tmp.data(), d_dl_.p, tmp.size() * sizeof(double))) return SQRTBA_ERR_CUDA;
      for (int l = 0; l < P_.n_point; l++) {
        const size_t lo = perm_.empty() ? (size_t)l : (size_t)perm_[l];
        for (int k = 0; k < 3; k++) dl[lo * 3 + k] = tmp[(size_t)k * P_.n_point + l];
      }
    }
    if (cg_iters) {
      if (download(c.data(), d_ctl_.p, c.size() * sizeof(WinCtl))) return SQRTBA_ERR_CUDA;
      int m = 0;
      for (auto& w : c) m = std::max(m, w.cg_iters);
      *cg_iters = m;
    }
    return SQRTBA_OK;
  }

  int debug_matvec(const double* p, double* y) {
    if (!have_problem_ || !P_.n_slot) { err_ = "no problem / no free pose"; return SQRTBA_ERR_INVALID; }
    DetOff det_off(P_);
    CU_CHECK(cudaSetDevice(cfg_.device));
    const size_t bytes = (size_t)P_.n_slot * 6 * sizeof(double);
    CU_CHECK(cudaMemcpyAsync(P_.p, p, bytes, cudaMemcpyHostToDevice, stream_));
    CU_CHECK(cudaMemsetAsync(P_.q, 0, bytes, stream_));
    launch_matvec(P_.p, P_.q, 1);
    CU_CHECK(cudaGetLastError());
    return download(y, P_.q, bytes);
  }

  int time_stage(int stage, int warmup, int reps, double* ms_avg) {
    if (!have_problem_) { err_ = "no problem set"; return SQRTBA_ERR_INVALID; }
    DetOff det_off(P_);
    CU_CHECK(cudaSetDevice(cfg_.device));
    const double d2 = (double)(float)std::sqrt(5.991), d3 = (double)(float)std::sqrt(7.815);
    const int gi = cdiv(P_.n_item, WARPS);
    auto one = [&]() {
      switch (stage) {
        case 0: launch_matvec(P_.p, P_.q, 1); break;
        case 1: launch_linearize(1, d2, d3, 1); break;
        case 2: launch_qr(1, 1.0); break;
        case 3: k_cost<<<gi, CTA, 0, stream_>>>(P_, 1, d2, d3); break;
        case 4: k_backsub<<<P_.n_tile, CTA, 0, stream_>>>(P_, 1, 1.0); break;
        case 5: k_linqr_pipe<<<cdiv(P_.n_tile, LIN_TPB), CTA, 0, stream_>>>(P_, 1, d2, d3, 1, 1.0); break;
        default: break;
      }
    };
    if (stage < 0 || stage > 5) { err_ = "time_stage: unknown stage"; return SQRTBA_ERR_INVALID; }
    if (stage == 3) {  // the cost kernel only runs for windows in a trial
      std::vector<WinCtl> c(P_.n_win);
      if (download(c.data(), d_ctl_.p, c.size() * sizeof(WinCtl))) return SQRTBA_ERR_CUDA;
      for (auto& w : c) w.phase = PH_TRIAL;
      CU_CHECK(cudaMemcpyAsync(d_ctl_.p, c.data(), c.size() * sizeof(WinCtl), cudaMemcpyHostToDevice, stream_));
      CU_CHECK(cudaStreamSynchronize(stream_));
    }
    for (int i = 0; i < warmup; i++) one();
    CU_CHECK(cudaStreamSynchronize(stream_));
    CU_CHECK(cudaEventRecord(ev0_, stream_));
    for (int i = 0; i < reps; i++) one();
    CU_CHECK(cudaEventRecord(ev1_, stream_));
    CU_CHECK(cudaEventSynchronize(ev1_));
    CU_CHECK(cudaGetLastError());
    float ms = 0;
    CU_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
    *ms_avg = (double)ms / std::max(reps, 1);
#ifdef SQRTBA_PIPE_PROF
    if (stage == 0 && P_.smallwin) {  // cycle breakdown of the last pipelined matvec launch, averaged over consumer warps
      const int nw = std::min(P_.n_tile, pipe_ctas_) * WARPS;
      std::vector<long long> hp((size_t)nw * 8);
      if (download(hp.data(), d_prof_.p, hp.size() * sizeof(long long))) return SQRTBA_ERR_CUDA;
      double acc[7] = {0, 0, 0, 0, 0, 0, 0};
      for (int w = 0; w < nw; w++)
        for (int i = 0; i < 7; i++) acc[i] += (double)hp[(size_t)w * 8 + i];
      const double tiles = acc[6] / WARPS;  // each consumer warp counts every tile of its CTA
      fprintf(stderr, "[pipe prof] tiles/CTA %.1f | cycles per tile: wait_full %.0f phase1 %.0f barA %.0f phase2 %.0f barB+arrive %.0f loop %.0f\n",
              acc[6] / nw, acc[0] / acc[6], acc[1] / acc[6], acc[2] / acc[6], acc[3] / acc[6], acc[4] / acc[6], acc[5] / acc[6]);
      (void)tiles;
    }
#endif
    return SQRTBA_OK;
  }

 private:
  int download(void* dst, const void* src, size_t bytes) {
    CU_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaStreamSynchronize(stream_));
    return SQRTBA_OK;
  }
  cudaError_t clear_traces() {
    // trace_len lives in WinCtl; zero the whole control block (lambda etc. are re-initialised at iteration 0)
    return cudaMemsetAsync(d_ctl_.p, 0, (size_t)P_.n_win * sizeof(WinCtl), stream_);
  }
  void launch_classify(int mode, double t2, double t3) {
    k_classify<<<cdiv(P_.n_obs, 256), 256, 0, stream_>>>(P_, mode, t2, t3);
    launches_++;
  }

  // second half of "linearise": cross-rank sums (landmark-sharded mode), then lambda init / currentChi per window
  int begin_after_linearize() {
    if (int rc = allreduce(P_.bp, (size_t)P_.n_slot * 12, false)) return rc;  // b_p and diag(Jp^T Jp)
    k_lm_reduce_lin<<<P_.n_win, RCTA, 0, stream_>>>(P_);
    if (int rc = allreduce(P_.wred, (size_t)P_.n_win, false)) return rc;                     // chi2: sum
    if (int rc = allreduce(P_.wred + 2 * (size_t)P_.n_win, (size_t)P_.n_win, true)) return rc;  // landmark max diag: max
    k_lm_begin<<<P_.n_win, RCTA, 0, stream_>>>(P_);
    CU_CHECK(cudaGetLastError());
    return SQRTBA_OK;
  }

  // The caller's stop flag (bool* pbStopFlag).  Landmark-sharded over several ranks every LM step issues collectives,
  // so all ranks must act on the SAME value: the flags are max-reduced (a rank that sees the flag a step earlier than
  // its peers would otherwise leave the loop while they block in the next all-reduce).  One extra small all-reduce +
  // read-back per LM trial; single-GPU solves just read the flag.
  int poll_stop(const volatile bool* stop, int* term) {
    const int local = (stop && *stop) ? 1 : 0;
    *term = local;
    if (!comm_) return SQRTBA_OK;
    h_counters_[3] = local;
    CU_CHECK(cudaMemcpyAsync(d_counters_.p + 3, h_counters_ + 3, sizeof(int), cudaMemcpyHostToDevice, stream_));
    const int rc = g_nccl.AllReduce(d_counters_.p + 3, d_counters_.p + 3, 1, /*ncclInt32*/ 2, /*ncclMax*/ 2, comm_, stream_);
    if (rc != 0) {
      err_ = std::string("ncclAllReduce (stop flag) failed: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
      return SQRTBA_ERR_COMM;
    }
    CU_CHECK(cudaMemcpyAsync(h_counters_ + 3, d_counters_.p + 3, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaStreamSynchronize(stream_));
    *term = h_counters_[3];
    return SQRTBA_OK;
  }

  // in-place all-reduce over the ranks that share this problem (no-op for a single GPU)
  int allreduce(double* buf, size_t count, bool is_max) {
    if (!comm_ || count == 0) return SQRTBA_OK;
    const int rc = g_nccl.AllReduce(buf, buf, count, /*ncclFloat64*/ 8, is_max ? /*ncclMax*/ 2 : /*ncclSum*/ 0, comm_, stream_);
    if (rc != 0) {
      err_ = std::string("ncclAllReduce failed: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
      return SQRTBA_ERR_COMM;
    }
    return SQRTBA_OK;
  }

 public:
  int comm_init(int nranks, int rank, const uint8_t* id128) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) { err_ = "comm_init: bad arguments"; return SQRTBA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(cfg_.device));
    if (!g_nccl.load()) { err_ = "comm_init: cannot load libnccl.so.2"; return SQRTBA_ERR_COMM; }
    comm_destroy();
    NcclApi::UniqueId id;
    std::memcpy(id.internal, id128, 128);
    const int rc = g_nccl.CommInitRank(&comm_, nranks, id, rank);
    if (rc != 0) {
      comm_ = nullptr;
      err_ = std::string("ncclCommInitRank failed: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
      return SQRTBA_ERR_COMM;
    }
    n_ranks_ = nranks;
    rank_ = rank;
    return SQRTBA_OK;
  }
  void comm_destroy() {
    if (stream_) cudaStreamSynchronize(stream_);
    peer_release();
    if (d_seq_) cudaFree(d_seq_);
    d_seq_ = nullptr;
    if (d_peer_tbl_) cudaFree(d_peer_tbl_);
    d_peer_tbl_ = nullptr;
    if (comm_ && g_nccl.CommDestroy) g_nccl.CommDestroy(comm_);
    comm_ = nullptr;
    n_ranks_ = 1;
    rank_ = 0;
  }
  int uses_peer_exchange() const { return (comm_ && peer_ok_) ? 1 : 0; }
  void dump_persist_prof() {
#ifdef SQRTBA_PIPE_PROF
    const int grid = persist_grid_;
    std::vector<long long> hp((size_t)grid * 16);
    if (download(hp.data(), d_prof_.p, hp.size() * sizeof(long long))) return;
    double acc[9] = {};
    double mx3 = 0;
    for (int b = 0; b < grid; b++) {
      for (int i = 0; i < 9; i++) acc[i] += (double)hp[(size_t)b * 16 + i];
      mx3 = std::max(mx3, (double)hp[(size_t)b * 16 + 3]);
    }
    const double its = std::max(acc[8], 1.0);
    fprintf(stderr, "[persist prof] grid %d, iterations/CTA %.0f | cycles per iteration (CTA average): tiles %.0f (max-CTA %.0f) flush %.0f "
            "B1 %.0f ph1 %.0f B2|ph2 %.0f ph3 %.0f B3 %.0f other %.0f\n", grid, its / grid, acc[3] / its, mx3 / (its / grid), acc[4] / its,
            acc[0] / its, acc[1] / its, acc[2] / its, acc[5] / its, acc[6] / its, acc[7] / its);
    {
      const int nchunk = (std::max(P_.n_slot, 1) + VSLOT - 1) / VSLOT;
      const int nown = std::min(nchunk, grid);
      double o9 = 0, o10 = 0, o1 = 0, o2 = 0, o11 = 0, o12 = 0, o13 = 0;
      for (int b = 0; b < nown; b++) { o9 += hp[(size_t)b * 16 + 9]; o10 += hp[(size_t)b * 16 + 10]; o1 += hp[(size_t)b * 16 + 1]; o2 += hp[(size_t)b * 16 + 2];
        o11 += hp[(size_t)b * 16 + 11]; o12 += hp[(size_t)b * 16 + 12]; o13 += hp[(size_t)b * 16 + 13]; }
      fprintf(stderr, "[persist prof] owner detail: Dq %.0f, dot shuffles %.0f, wait for the CTA's other warps %.0f\n", o11 / (its / grid * nown), o12 / (its / grid * nown), o13 / (its / grid * nown));
      const double oi = its / grid * nown;
      fprintf(stderr, "[persist prof] owner CTAs (%d): cycles per iteration: B1->exchange start %.0f, exchange %.0f, rest of phase 1 %.0f, B2 wait %.0f\n",
              nown, o9 / oi, o10 / oi, o1 / oi, o2 / oi);
    }
    {
      std::vector<std::pair<long long, int>> v;
      for (int b = 0; b < grid; b++) v.push_back({hp[(size_t)b * 16 + 3], b});
      std::sort(v.begin(), v.end());
      std::vector<int> ptr(grid + 1);
      download(ptr.data(), d_ptile_.p, ptr.size() * sizeof(int));
      fprintf(stderr, "[persist prof] slowest CTAs (cta:tiles:cycles/iter):");
      for (int i = 0; i < 12; i++) { auto& x = v[grid - 1 - i]; fprintf(stderr, " %d:%d:%.0f", x.second, ptr[x.second + 1] - ptr[x.second], x.first / (its / grid)); }
      fprintf(stderr, "\n[persist prof] fastest:");
      for (int i = 0; i < 6; i++) { auto& x = v[i]; fprintf(stderr, " %d:%d:%.0f", x.second, ptr[x.second + 1] - ptr[x.second], x.first / (its / grid)); }
      fprintf(stderr, "\n");
    }
    if (const char* path = std::getenv("SQRTBA_PROF_CSV")) {  // per-CTA cycles + the features a cost model can use
      std::vector<int> ptr(grid + 1);
      download(ptr.data(), d_ptile_.p, ptr.size() * sizeof(int));
      if (FILE* f = std::fopen(path, "a")) {
        std::fprintf(f, "cta,cycles,tiles,obs,nrun,outside,slides,long_obs,items\n");
        const bool big = max_win_slots_ > MAXSLOT;
        for (int b = 0; b < grid; b++) {
          long long obs = 0, nrun = 0, outside = 0, slides = 0, lobs = 0, items = 0;
          int abase = 0;
          for (int t = ptr[b]; t < ptr[b + 1]; t++) {
            const TileInfo& ti = h_tiles_pin_.p[t];
            obs += ti.o1 - ti.o0; nrun += ti.nrun; items += ti.nitem;
            if (ti.is_long) lobs += ti.o1 - ti.o0;
            if (big && !ti.is_long && ti.nrun > 0) {
              const int* runs = h_tile_runs_.p + h_tile_run_ptr_.p[t];
              const int lo = runs[ti.nrun + 1];
              const int want = std::max(0, std::min(lo, P_.n_slot - persist_slots_));
              if (t == ptr[b] || want < abase || want > abase + 6) { abase = want; slides++; }
              for (int i = 0; i < ti.nrun; i++)
                if (runs[ti.nrun + 1 + i] >= abase + persist_slots_ || runs[ti.nrun + 1 + i] < abase) outside += runs[i + 1] - runs[i];
            }
          }
          std::fprintf(f, "%d,%.0f,%d,%lld,%lld,%lld,%lld,%lld,%lld\n", b, hp[(size_t)b * 16 + 3] / (its / grid), ptr[b + 1] - ptr[b], obs, nrun,
                       outside, slides, lobs, items);
        }
        std::fclose(f);
      }
    }
    cudaMemsetAsync(d_prof_.p, 0, d_prof_.cap * sizeof(long long), stream_);
#endif
  }

 private:
  // matvec dispatch: persistent TMA-pipelined kernel when every window is small, general tile kernel otherwise
  static size_t pipe_smem_bytes(int S, int maxslot, bool big) {
    return ((size_t)S * JQ_STAGE_D + 12 * (CTA + 1) + 2 + (big ? 6 : 15) * (size_t)maxslot) * sizeof(double) +
           2 * (2 * CTA + 4) * sizeof(int) + 2 * S * sizeof(uint64_t);
  }
  static size_t persist_smem_bytes(int S, int maxslot, bool big, bool chunk = false) {
    return ((size_t)S * JQ_STAGE_D + 12 * (CTA + 1) + 2 + (big ? 6 : 18) * (size_t)maxslot + 16) * sizeof(double) +
           2 * S * sizeof(uint64_t) + 4 * sizeof(int) + 2 * (size_t)pipe_run_cap(maxslot, big) * sizeof(int) +
           (chunk ? (size_t)CH_PACK * sizeof(float) : 0);
  }
  // the chunk preconditioner is live for this solve: decided per problem, and only the persistent kernel applies it
  // PCG tolerance.  0 = automatic: 1e-7 for windows whose poses fit the shared-memory path (local BA, batches), 1e-9 for
  // big windows (global BA).  Measured on C0-shaped windows, 12 seeds (tools/rtol_probe_local.py):
  // trial sequence, outlier flags, cost (6e-8) and pose RMS (5e-8 m) are the same from 1e-9 up to 1e-6 and break at 1e-5
  // -- the cost error of an inexact step is second order in the linear residual -- while 1e-7 needs 10 % fewer iterations.
  // A handle with a third pass (the fork's schedule: 20 more iterations at the minimum, lidar edges with central-difference
  // Jacobians of step 1e-9, g2oOptimizer.cc:979-1117) keeps 1e-9 throughout: finite differences of step 1e-9 turn a 1e-8
  // difference of the state that enters the pass into per-trial cost differences above 1e-6 (tests/test_gpu_lidar.py).
  // Global BA with the two-level preconditioner: 1e-8.  Full C3 against the reference algorithm's exact solve (tools/rtol_probe_c3.py): pose RMS
  // 1.6e-7 / 2.0e-7 / 2.4e-7 / 5.2e-7 m at 1e-9 / 1e-8 / 1e-7 / 1e-6, identical trial sequences, 553 / 490 / 431 / 366
  // iterations -- the preconditioned residual of a strong preconditioner is close to the energy norm of the error.  The
  // 6x6 blocks (pcg_mode 5, sharded runs without peer access) keep 1e-9.
  double eff_rtol() const {
    if (cfg_.pcg_rtol > 0) return cfg_.pcg_rtol;
    if (cfg_.third_pass_iters > 0) return 1e-9;
    if (P_.pq_shared) return 1e-7;
    return (chunk_on() && coarse_active_) ? 1e-8 : 1e-9;
  }
  bool chunk_on() const { return chunk_active_ && use_persist(); }
  // From how many free keyframes on ONE window takes the big-window PCG path (chunked vector phase, two-level
  // preconditioner) instead of the shared-memory one.  Measured on C2-shaped windows (tools/big_window_crossover.py):
  // 40 keyframes 10.0 against 12.5 ms, 60: 16.6 / 15.7, 80: 24.2 / 19.3, 99 (C2): 31.8 / 24.0, 120: 42.5 / 27.5 -- the
  // iteration count halves, an iteration costs more (three grid barriers, q through global atomics).  64 = four chunks:
  // the preconditioner then forms at most a quarter of the blocks of the reduced system.  The modes that need the
  // shared-memory path (multi-launch, reproducible) or ask for the 6x6 blocks keep the old limit.
  int big_window_min_slots() const {
    static const int env = [] {
      const char* e = std::getenv("SQRTBA_BIG_MIN_SLOTS");
      return e ? std::max(2, std::atoi(e)) : 0;
    }();
    if (env > 0) return env;
    return (cfg_.pcg_mode == 1 || cfg_.pcg_mode == 4 || cfg_.pcg_mode == 5) ? MAXSLOT + 1 : 64;
  }
  // off-diagonal blocks of every chunk -> [all-reduce] -> inverse + CG start vectors + r0.z0
  int enqueue_chunk_prec() {
    int nchunk = (P_.n_slot + VSLOT - 1) / VSLOT;
    const size_t m_floats = (size_t)nchunk * CH_MSIZE, c_floats = coarse_active_ ? (size_t)nchunk * nchunk * CH_MBLK : 0;
    float* cacc = coarse_active_ ? d_chunk_M_.p + m_floats : nullptr;
    CU_CHECK(cudaMemsetAsync(d_chunk_M_.p, 0, (m_floats + c_floats) * sizeof(float), stream_));
    // how old a level of the preconditioner may be, in outer LM iterations (1 = rebuilt for every trial).  Measured on C3,
    // one GPU: both levels every 2nd iteration cost 8 CG iterations of 490 and save five block builds (1.9 ms), chunk
    // inversions and coarse inversions (0.54 ms): 96.6 -> 84.7 ms; every 3rd: 507 iterations, 83.4 ms; a coarse level that
    // is never refreshed within the pass: 1148 iterations (lambda falls by 3^9 over the pass and sits on its diagonal)
    static const int refresh_chunk_env = [] { const char* e = std::getenv("SQRTBA_CHUNK_REFRESH"); return e ? std::max(1, std::atoi(e)) : 2; }();
    static const int refresh_coarse_env = [] { const char* e = std::getenv("SQRTBA_COARSE_REFRESH"); return e ? std::max(1, std::atoi(e)) : 2; }();
    const int refresh_chunk = refresh_chunk_env;
    int refresh = std::max(refresh_coarse_env, refresh_chunk);  // the coarse level is built from the chunk level's sums
    if (refresh % refresh_chunk != 0) refresh = refresh_chunk * ((refresh + refresh_chunk - 1) / refresh_chunk);
    k_chunk_blocks<<<P_.n_tile, CTA, 0, stream_>>>(P_, d_chunk_M_.p, cacc, nchunk, refresh_chunk, refresh);
    if (comm_) {  // every rank adds its landmarks' pairs; the sum is the same bits everywhere
      const int rc = g_nccl.AllReduce(d_chunk_M_.p, d_chunk_M_.p, m_floats + c_floats, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm_, stream_);
      if (rc != 0) {
        err_ = std::string("ncclAllReduce (chunk blocks) failed: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
        return SQRTBA_ERR_COMM;
      }
    }
    k_chunk_factor<<<nchunk, CH_FACTOR_THREADS, CH_FACTOR_SMEM, stream_>>>(P_, d_chunk_M_.p, d_chunk_pack_.p, d_chunk_diag_.p, d_chunk_rz_.p,
                                                                         coarse_active_ ? d_co_dblk_.p : nullptr,
                                                                         coarse_active_ ? d_co_ctl_.p : nullptr, refresh_chunk);
    if (coarse_active_) {
      const float* cacc_c = cacc;
      const double* dblk = d_co_dblk_.p;
      double* rbuf = d_co_R_.p;
      double* aci = d_co_aci_.p;
      unsigned* ctl = d_co_ctl_.p;
      void* args[] = {(void*)&P_, (void*)&cacc_c, (void*)&dblk, (void*)&rbuf, (void*)&aci, (void*)&ctl, (void*)&nchunk, (void*)&refresh};
      CU_CHECK(cudaLaunchCooperativeKernel((const void*)k_coarse_invert, dim3(nchunk), dim3(CO_THREADS), args, 0, stream_));
    }
    k_chunk_z0<<<nchunk, CO_THREADS, 0, stream_>>>(P_, d_chunk_pack_.p, d_chunk_diag_.p, coarse_active_ ? d_co_aci_.p : nullptr,
                                                   d_chunk_rz_.p, nchunk);
    k_chunk_rz<<<1, 32, 0, stream_>>>(P_, d_chunk_rz_.p, nchunk);
    return SQRTBA_OK;
  }
  int chunk_prec_launches() const { return coarse_active_ ? 5 : 4; }
  void launch_linearize(int robust, double d2, double d3, int force_all) {
    if (cfg_.reserved[7] == 1) k_linearize<<<P_.n_tile, CTA, 0, stream_>>>(P_, robust, d2, d3, force_all);
    else if (cfg_.reserved[7] == 2) k_linearize_pipe<false><<<cdiv(P_.n_tile, LIN_TPB), CTA, 0, stream_>>>(P_, robust, d2, d3, force_all);
    else if (cfg_.reserved[7] == 6) k_linearize_pipe<true, true><<<cdiv(P_.n_tile, LIN_TPB), CTA, 0, stream_>>>(P_, robust, d2, d3, force_all);
    else k_linearize_pipe<true><<<cdiv(P_.n_tile, LIN_TPB), CTA, P_.det ? 12 * (size_t)P_.maxslot * sizeof(double) : 0, stream_>>>(
        P_, robust, d2, d3, force_all);  // 0 and 7
  }
  // fused linearisation + landmark QR for the windows whose lambda is already known (iterations > 0, retries)
  void launch_linqr(int robust, double d2, double d3) {
    k_linqr_pipe<<<cdiv(P_.n_tile, LIN_TPB), CTA, 0, stream_>>>(P_, robust, d2, d3, 0, 0.0);
  }
  // landmark QR: cp.async-pipelined kernel (QR_TPB tiles per CTA); reserved[7] != 0 selects the plain one-tile-per-CTA kernel
  void launch_qr(int force_all, double lam_override) {
    // reserved[7]: 0 = pipelined v2 (default), 1 = plain one-tile-per-CTA kernels, 2 = first pipelined version,
    // 5 = v2 compiled for 5 CTAs/SM (96 registers) (A/B)
    if (cfg_.reserved[7] == 1) k_qr<<<P_.n_tile, CTA, 0, stream_>>>(P_, force_all, lam_override);
    else if (cfg_.reserved[7] == 2) k_qr_pipe<<<cdiv(P_.n_tile, QR_TPB), CTA, QR_PIPE_SMEM, stream_>>>(P_, force_all, lam_override);
    else if (cfg_.reserved[7] == 5) k_qr_pipe2<0, 5><<<cdiv(P_.n_tile, QR_TPB), CTA, QR_PIPE2_SMEM, stream_>>>(P_, force_all, lam_override);
    else k_qr_pipe2<0, 4><<<cdiv(P_.n_tile, QR_TPB), CTA, QR_PIPE2_SMEM + (P_.det ? 27 * (size_t)P_.maxslot * sizeof(double) : 0), stream_>>>(
        P_, force_all, lam_override);
    if (P_.det) { k_reduce_qr<<<P_.n_win, RCTA, 0, stream_>>>(P_); launches_++; }  // the CTAs' partial vectors, in order
  }
  void launch_matvec(const double* pvec, double* qvec, int force_all) {
    if (P_.smallwin && cfg_.reserved[1] == 0) {
      const int grid = std::min(P_.n_tile, pipe_ctas_);
      const bool big = !P_.pq_shared;
      const size_t bytes = pipe_smem_bytes(pipe_stages_, pipe_slots_, big);
      if (big) {
        if (pipe_stages_ == 2)
          k_matvec_pipe<2, true><<<grid, PIPE_THREADS, bytes, stream_>>>(P_, pvec, qvec, force_all, pipe_slots_);
        else
          k_matvec_pipe<3, true><<<grid, PIPE_THREADS, bytes, stream_>>>(P_, pvec, qvec, force_all, pipe_slots_);
      } else if (pipe_stages_ == 2) {
        k_matvec_pipe<2, false><<<grid, PIPE_THREADS, bytes, stream_>>>(P_, pvec, qvec, force_all, pipe_slots_);
      } else {
        k_matvec_pipe<3, false><<<grid, PIPE_THREADS, bytes, stream_>>>(P_, pvec, qvec, force_all, pipe_slots_);
      }
    } else {
      k_matvec<<<P_.n_tile, CTA, 0, stream_>>>(P_, pvec, qvec, force_all);
    }
  }

  // pcg_mode: 0 = auto (persistent kernel for single-window problems, one launch per CG phase for batches),
  //           1 = always one launch per phase, 2 = persistent whenever possible
  bool grow_mv_events() {
    if (mv_events_.size() >= 8192) return false;  // 4096 timed launches per solve are plenty
    for (int i = 0; i < 256; i++) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) return false;
      mv_events_.push_back(e);
    }
    return true;
  }
  bool use_persist() const {
    if (cfg_.pcg_mode == 1 || !coop_ok_ || P_.n_win != 1 || !P_.smallwin || persist_ctas_ <= 0 || cfg_.reserved[1] != 0) return false;
    if (lidar_active_) return false;  // the unary term H_u p is added between the matvec and the CG update (multi-launch loop)
    if (cfg_.pcg_mode == 4) return false;  // reproducible mode: one launch per CG phase, fixed-order sums in k_cg_step
    if (comm_ && (!peer_ok_ || P_.pq_shared)) return false;  // sharded without peer-mapped buffers, or a small window: NCCL all-reduce per iteration
    return true;
  }
  int launch_pcg_persist(double tol2, int use_override, double lam_override) {
    PcgArgs A{};
    A.tol2 = tol2;
    A.max_iters = cfg_.pcg_max_iters;
    A.nranks = comm_ ? n_ranks_ : 1;
    A.rank = comm_ ? rank_ : 0;
    A.gbar = d_gbar_.p;
    A.part = d_part_.p;
    A.q3 = d_q3_.p;
    A.dq = d_dq_.p;
    A.qf = d_dq_.p + (size_t)6 * std::max(P_.n_slot, 1);
    if (A.nranks > 1) {
      A.recv = (uint4*)peer_recv_[rank_];
      A.peer_tbl = (uint4* const*)d_peer_tbl_;
      A.seq_state = d_seq_;
      A.nelem_cap = (int)peer_nelem_cap_;
    }
    const bool big = !P_.pq_shared;
    const int grid = persist_grid_;
    A.tile_ptr = d_ptile_.p;
    const bool chunk = big && chunk_on();
    A.cpack = chunk ? d_chunk_pack_.p : nullptr;
    A.cdiag = chunk ? d_chunk_diag_.p : nullptr;
    A.aci = (chunk && coarse_active_) ? d_co_aci_.p : nullptr;
    A.wvec = (chunk && coarse_active_) ? d_co_w_.p : nullptr;
    const size_t bytes = persist_smem_bytes(persist_stages_, persist_slots_, big, chunk);
    int maxslot = persist_slots_;
    void* args[] = {(void*)&P_, (void*)&A, (void*)&maxslot, (void*)&lam_override, (void*)&use_override};
    const bool multi = A.nranks > 1;  // use_persist() admits several ranks only for big windows
    const void* fn = chunk ? (multi ? (const void*)k_pcg_persist<2, true, true, true> : (const void*)k_pcg_persist<2, true, false, true>)
                     : big ? (multi ? (persist_stages_ == 2 ? (const void*)k_pcg_persist<2, true, true> : (const void*)k_pcg_persist<3, true, true>)
                                  : (persist_stages_ == 2 ? (const void*)k_pcg_persist<2, true> : (const void*)k_pcg_persist<3, true>))
                         : (persist_stages_ == 2 ? (const void*)k_pcg_persist<2, false> : (const void*)k_pcg_persist<3, false>);
    CU_CHECK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(PIPE_THREADS), args, bytes, stream_));
    return SQRTBA_OK;
  }

  // Peer-mapped exchange buffers of the landmark-sharded mode (collective: every rank calls with the same size).
  // Each rank allocates its receive buffer with cudaMalloc, the cudaIpc handles travel through an NCCL all-gather on
  // the existing communicator, and every rank maps every peer's buffer.  If any rank cannot map a peer (no P2P), all
  // ranks agree to fall back to ncclAllReduce per CG iteration.
  int peer_setup(size_t nelem) {
    if (!comm_ || n_ranks_ <= 1) return SQRTBA_OK;
    if (peer_ok_ && nelem <= peer_nelem_cap_) return SQRTBA_OK;
    if (n_ranks_ > 8 || cfg_.reserved[6] != 0) { peer_ok_ = false; return SQRTBA_OK; }
    peer_release();
    peer_nelem_cap_ = nelem;
    const size_t recv_bytes = 2 * (size_t)n_ranks_ * nelem * 16;  // 16-byte {value, sequence} records, two parities
    double* recv = nullptr;
    CU_CHECK(cudaMalloc((void**)&recv, recv_bytes));
    if (!d_seq_) CU_CHECK(cudaMalloc((void**)&d_seq_, sizeof(unsigned long long)));
    CU_CHECK(cudaMemsetAsync(recv, 0, recv_bytes, stream_));
    CU_CHECK(cudaMemsetAsync(d_seq_, 0, sizeof(unsigned long long), stream_));
    peer_recv_[rank_] = recv;
    std::vector<cudaIpcMemHandle_t> hs(n_ranks_);
    int ok = 1;
    if (cudaIpcGetMemHandle(&hs[rank_], recv) != cudaSuccess) {
      ok = 0;
      cudaGetLastError();
      std::memset(&hs[rank_], 0, sizeof(cudaIpcMemHandle_t));
    }
    cudaIpcMemHandle_t* d_h = nullptr;
    CU_CHECK(cudaMalloc((void**)&d_h, sizeof(cudaIpcMemHandle_t) * n_ranks_));
    CU_CHECK(cudaMemcpyAsync(d_h + rank_, &hs[rank_], sizeof(cudaIpcMemHandle_t), cudaMemcpyHostToDevice, stream_));
    if (g_nccl.AllGather(d_h + rank_, d_h, sizeof(cudaIpcMemHandle_t), /*ncclInt8*/ 0, comm_, stream_) != 0) {
      cudaFree(d_h);
      err_ = "ncclAllGather of the peer handles failed";
      return SQRTBA_ERR_COMM;
    }
    CU_CHECK(cudaMemcpyAsync(hs.data(), d_h, sizeof(cudaIpcMemHandle_t) * n_ranks_, cudaMemcpyDeviceToHost, stream_));
    CU_CHECK(cudaStreamSynchronize(stream_));
    cudaFree(d_h);
    for (int r = 0; r < n_ranks_ && ok; r++) {
      if (r == rank_) continue;
      void* pr = nullptr;
      if (cudaIpcOpenMemHandle(&pr, hs[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ok = 0;
        cudaGetLastError();
        break;
      }
      peer_recv_[r] = (double*)pr;
    }
    // every rank must take the same path: agree on success (min over ranks)
    double* d_ok = nullptr;
    CU_CHECK(cudaMalloc((void**)&d_ok, sizeof(double)));
    const double okd = ok;
    CU_CHECK(cudaMemcpyAsync(d_ok, &okd, sizeof(double), cudaMemcpyHostToDevice, stream_));
    double all_ok = 0.0;
    const int rc = g_nccl.AllReduce(d_ok, d_ok, 1, /*ncclFloat64*/ 8, /*ncclMin*/ 3, comm_, stream_);
    if (rc == 0) {
      CU_CHECK(cudaMemcpyAsync(&all_ok, d_ok, sizeof(double), cudaMemcpyDeviceToHost, stream_));
      CU_CHECK(cudaStreamSynchronize(stream_));
    }
    cudaFree(d_ok);
    peer_ok_ = (rc == 0) && all_ok > 0.5;
    if (peer_ok_) {
      if (!d_peer_tbl_) CU_CHECK(cudaMalloc((void**)&d_peer_tbl_, 8 * sizeof(void*)));
      CU_CHECK(cudaMemcpyAsync(d_peer_tbl_, peer_recv_, 8 * sizeof(void*), cudaMemcpyHostToDevice, stream_));
      CU_CHECK(cudaStreamSynchronize(stream_));
    }
    if (!peer_ok_) peer_release();
    return SQRTBA_OK;
  }
  void peer_release() {
    for (int r = 0; r < 8; r++) {
      if (peer_recv_[r]) {
        if (r == rank_) cudaFree(peer_recv_[r]);
        else cudaIpcCloseMemHandle(peer_recv_[r]);
      }
      peer_recv_[r] = nullptr;
    }
    peer_ok_ = false;
    peer_nelem_cap_ = 0;
  }

  // QR + block-Jacobi + PCG for every window in PH_TRIAL
  int factor_and_solve() {
    const int gi = cdiv(P_.n_item, WARPS);
    if (P_.n_slot) k_zero_trial<<<cdiv(P_.n_slot, 128), 128, 0, stream_>>>(P_);
    stage_begin(1);
    launch_qr(0, 0.0);
    stage_end(1);
    launches_ += 2;
    if (!P_.n_slot) return SQRTBA_OK;
    if (int rc = allreduce(P_.bs, (size_t)P_.n_slot * 27, false)) return rc;  // reduced rhs + block-Jacobi blocks
    if (lidar_active_) { k_lidar_trial_add<<<1, 32, 0, stream_>>>(P_, lidar_); launches_++; }
    if (!chunk_on()) k_dinv<<<cdiv(P_.n_slot, 64), 64, 0, stream_>>>(P_, 0, 0.0);
    stage_begin(2);
    CU_CHECK(cudaMemsetAsync(P_.counters + 1, 0, sizeof(int), stream_));
    k_cg_init<<<P_.n_win, RCTA, 0, stream_>>>(P_, 0, chunk_on() ? 1 : 0);
    launches_ += 2;
    if (chunk_on()) {
      if (int rc = enqueue_chunk_prec()) return rc;
      launches_ += chunk_prec_launches();
    }
    const double tol2 = cfg_.pcg_rtol > 0 ? cfg_.pcg_rtol * cfg_.pcg_rtol : -1.0;  // <= 0: the pass's own (WinCtl::tol2)
    const int check = std::max(1, cfg_.pcg_check_every);
    const size_t qbytes = (size_t)P_.n_slot * 6 * sizeof(double);
    if (use_persist()) {  // whole PCG solve in one cooperative launch (single-window problems)
      if (P_.pq_shared) CU_CHECK(cudaMemsetAsync(d_q3_.p, 0, 3 * KQ * qbytes, stream_));
      else {
        CU_CHECK(cudaMemsetAsync(P_.q, 0, qbytes, stream_));
        CU_CHECK(cudaMemsetAsync(d_dq_.p, 0, 2 * qbytes, stream_));
      }
      CU_CHECK(cudaMemsetAsync(d_gbar_.p, 0, 2 * sizeof(unsigned), stream_));
      if (int rc = launch_pcg_persist(tol2, 0, 0.0)) return rc;
      launches_ += 1;
      stage_end(2);
      CU_CHECK(cudaGetLastError());
      return SQRTBA_OK;
    }
    for (int it = 0; it < cfg_.pcg_max_iters; it++) {
      CU_CHECK(cudaMemsetAsync(P_.q, 0, qbytes, stream_));
      // live duration of the dominant kernel: an event pair per launch, read back after the solve (no extra sync)
      cudaEvent_t ea = nullptr, eb = nullptr;
      if (mv_used_ + 2 <= (int)mv_events_.size() || grow_mv_events()) { ea = mv_events_[mv_used_]; eb = mv_events_[mv_used_ + 1]; mv_used_ += 2; }
      if (ea) cudaEventRecord(ea, stream_);
      launch_matvec(P_.p, P_.q, 0);
      if (eb) cudaEventRecord(eb, stream_);
      if (int rc = allreduce(P_.q, (size_t)P_.n_slot * 6, false)) return rc;  // the one exchange step of a CG iteration
      if (lidar_active_) { k_lidar_matvec<<<1, 32, 0, stream_>>>(P_, lidar_, P_.p, P_.q); launches_++; }
      k_cg_step<<<P_.n_win, RCTA, 0, stream_>>>(P_, tol2, cfg_.pcg_max_iters, 0, 0.0, P_.det ? std::min(P_.n_tile, pipe_ctas_) : 0);
      launches_ += 2;
      cg_iters_total_++;
      if ((it + 1) % check == 0) {
        CU_CHECK(cudaMemcpyAsync(h_counters_, P_.counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream_));
        CU_CHECK(cudaStreamSynchronize(stream_));
        if (h_counters_[1] <= 0) break;
      }
    }
    stage_end(2);
    CU_CHECK(cudaGetLastError());
    return SQRTBA_OK;
  }

  // ---- single-window problems: one LM macro step = ONE CUDA graph launch, verdict read from mapped host memory
  // A C0-shaped window spends ~150 us per LM trial in kernels; enqueueing ~20 small launches + memsets per trial and a
  // stream synchronisation + copy to read three counters cost about as much again.  The graph has 10 kernel nodes (the
  // small per-window kernels fused: k_trial_begin, k_cg_prep, k_push_update, k_decide_publish), the cooperative PCG
  // kernel included; the last decision kernel writes the step's counters and a sequence number into mapped pinned
  // memory (HostCtl) and the host spins on it.  The stop flag is mirrored into the same block, so the device reads it
  // where g2o's do-while evaluates terminate() -- at the end of the trial.
  bool use_graph() const {
    return graph_ok_ && cfg_.pcg_mode != 3 && !stage_timing_ && !comm_ && P_.n_win == 1 && P_.n_slot > 0 && use_persist();
  }
  int enqueue_step(double d2, double d3) {
    const int gi = cdiv(P_.n_item, WARPS);
    const int n6 = P_.n_slot * 6;
    launch_linearize(-1, d2, d3, 0);
    if (P_.fused) launch_linqr(-1, d2, d3);
    k_trial_begin<<<1, RCTA, 0, stream_>>>(P_);
    launch_qr(0, 0.0);
    if (P_.pq_shared) k_cg_prep<<<1, RCTA, 0, stream_>>>(P_, d_q3_.p, 3 * KQ * n6, nullptr, 0, d_gbar_.p, 0);
    else k_cg_prep<<<1, RCTA, 0, stream_>>>(P_, P_.q, n6, d_dq_.p, 2 * n6, d_gbar_.p, chunk_on() ? 1 : 0);
    if (chunk_on()) {
      if (int rc = enqueue_chunk_prec()) return rc;
    }
    const double tol2 = cfg_.pcg_rtol > 0 ? cfg_.pcg_rtol * cfg_.pcg_rtol : -1.0;  // <= 0: the pass's own (WinCtl::tol2)
    if (int rc = launch_pcg_persist(tol2, 0, 0.0)) return rc;
    k_backsub<<<P_.n_tile, CTA, 0, stream_>>>(P_, 0, 0.0);
    k_push_update<<<cdiv(std::max(P_.n_slot, P_.n_point), 256), 256, 0, stream_>>>(P_);
    k_cost<<<gi, CTA, 0, stream_>>>(P_, -1, d2, d3);
    k_decide_publish<<<1, RCTA, 0, stream_>>>(P_, d_ctl_host_);
    k_restore<<<cdiv(std::max(P_.n_pose, P_.n_point), 256), 256, 0, stream_>>>(P_);
    return SQRTBA_OK;
  }
  static constexpr int STEP_NODES = 11;
  // (re)capture the macro step for the current problem; an existing executable graph is updated in place
  int ensure_step_graph(double d2, double d3) {
    if (step_graph_valid_ && step_d2_ == d2 && step_d3_ == d3) return SQRTBA_OK;
    static const bool host_timing = std::getenv("SQRTBA_HOST_TIMING") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); graph_ok_ = false; return SQRTBA_OK; }
    const int rc = enqueue_step(d2, d3);
    const cudaError_t le = cudaGetLastError();
    const cudaError_t ce = cudaStreamEndCapture(stream_, &g);
    if (rc || le != cudaSuccess || ce != cudaSuccess || !g) {  // e.g. a driver that cannot capture cooperative launches
      cudaGetLastError();
      if (g) cudaGraphDestroy(g);
      graph_ok_ = false;
      return SQRTBA_OK;  // the caller falls back to one launch per kernel
    }
    if (step_exec_) {
      cudaGraphExecUpdateResultInfo info;
      if (cudaGraphExecUpdate(step_exec_, g, &info) != cudaSuccess) {
        cudaGetLastError();
        cudaGraphExecDestroy(step_exec_);
        step_exec_ = nullptr;
      }
    }
    if (!step_exec_ && cudaGraphInstantiate(&step_exec_, g, 0) != cudaSuccess) {
      cudaGetLastError();
      step_exec_ = nullptr;
      graph_ok_ = false;
    }
    cudaGraphDestroy(g);
    step_graph_valid_ = step_exec_ != nullptr;
    step_d2_ = d2; step_d3_ = d3;
    if (host_timing)
      std::fprintf(stderr, "[sqrtba host] %-28s %8.3f ms\n", "macro-step graph (re)capture",
                   std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    return SQRTBA_OK;
  }
  int run_pass_graph(int iters, const volatile bool* stop) {
    const int max_macro = iters * 10 + 1;
    for (int step = 0; step < max_macro; step++) {
      if (stop && *stop) {  // `for (i < iterations && !terminate())`: the next iteration does not start
        if (step > 0) { k_terminate<<<cdiv(P_.n_win, 128), 128, 0, stream_>>>(P_); launches_++; }
        break;
      }
      h_ctl_->term = 0;
      const int want = ++step_seq_;
      CU_CHECK(cudaGraphLaunch(step_exec_, stream_));
      unsigned spins = 0;
      while (h_ctl_->seq < want) {
        if (stop && *stop) h_ctl_->term = 1;  // seen by k_decide_publish when the trial ends
        if ((++spins & 0xfffu) == 0) {
          const cudaError_t q = cudaStreamQuery(stream_);
          if (q == cudaSuccess) {
            if (h_ctl_->seq >= want) break;
            err_ = "LM macro step finished without publishing its counters";
            return SQRTBA_ERR_CUDA;
          }
          if (q != cudaErrorNotReady) {
            err_ = std::string("LM macro step failed: ") + cudaGetErrorString(q);
            return SQRTBA_ERR_CUDA;
          }
        }
        __builtin_ia32_pause();
      }
      launches_ += STEP_NODES + (chunk_on() ? chunk_prec_launches() : 0);
      lm_trials_++;
      cg_iters_total_ = h_ctl_->counters[2];
      if (h_ctl_->counters[0] >= P_.n_win) break;
    }
    return SQRTBA_OK;
  }

  // one optimizer.optimize(iters) call over all windows (sparse_optimizer.cpp:354-419)
  int run_pass(int iters, int pass, int robust, double d2, double d3, const volatile bool* stop) {
    const int gi = cdiv(P_.n_item, WARPS);
    lidar_active_ = pass == 2 && (lidar_edges_set_ || lidar_assoc_set_);  // the edges join the graph for the third pass only
    k_pass_init<<<cdiv(P_.n_win, 128), 128, 0, stream_>>>(P_, iters, pass, robust, eff_rtol() * eff_rtol());
    CU_CHECK(cudaMemsetAsync(P_.counters, 0, 2 * sizeof(int), stream_));
    launches_++;
    if (iters <= 0) return SQRTBA_OK;
    if (use_graph()) {
      if (int rc = ensure_step_graph(d2, d3)) return rc;
      if (step_graph_valid_) {
        if (P_.n_slot) { k_zero_lin<<<cdiv(P_.n_slot, 128), 128, 0, stream_>>>(P_); launches_++; }  // later steps: k_decide_publish
        return run_pass_graph(iters, stop);
      }
    }
    const int max_macro = iters * 10 + 1;
    for (int step = 0; step < max_macro; step++) {
      int term = 0;
      if (int rc = poll_stop(stop, &term)) return rc;
      if (term) {  // `for (i < iterations && !terminate())`: the next iteration does not start (sparse_optimizer.cpp:376)
        if (step > 0) { k_terminate<<<cdiv(P_.n_win, 128), 128, 0, stream_>>>(P_); launches_++; }
        break;
      }
      if (P_.n_slot) k_zero_lin<<<cdiv(P_.n_slot, 128), 128, 0, stream_>>>(P_);
      stage_begin(0);
      launch_linearize(robust, d2, d3, 0);
      if (P_.fused) { launch_linqr(robust, d2, d3); launches_++; }
      stage_end(0);
      if (lidar_active_) { k_lidar_lin<<<1, LD_CTA, 0, stream_>>>(P_, lidar_); launches_++; }
      if (int rc = begin_after_linearize()) return rc;
      launches_ += 4;
      int rc = factor_and_solve();
      if (rc) return rc;
      stage_begin(3);
      k_backsub<<<P_.n_tile, CTA, 0, stream_>>>(P_, 0, 0.0);
      stage_end(3);
      CU_CHECK(cudaMemcpyAsync(d_pose_bak_.p, d_pose_.p, (size_t)P_.n_pose * 7 * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
      CU_CHECK(cudaMemcpyAsync(d_point_bak_.p, d_point_.p, (size_t)P_.n_point * 3 * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
      if (P_.n_slot) k_update_pose<<<cdiv(P_.n_slot, 128), 128, 0, stream_>>>(P_);
      k_update_point<<<cdiv(P_.n_point, 256), 256, 0, stream_>>>(P_);
      stage_begin(4);
      k_cost<<<gi, CTA, 0, stream_>>>(P_, robust, d2, d3);
      stage_end(4);
      if (lidar_active_) { k_lidar_cost<<<1, LD_CTA, 0, stream_>>>(P_, lidar_); launches_++; }
      k_lm_reduce_trial<<<P_.n_win, RCTA, 0, stream_>>>(P_);
      if (int rc = allreduce(P_.wred, (size_t)2 * P_.n_win, false)) return rc;  // trial chi2 + landmark part of the scale
      // the flag is looked at again where g2o's do-while evaluates terminate(): after the trial's solve was enqueued
      // (single-rank only: sharded ranks must act on the value they agreed on at the top of the step)
      if (!comm_ && stop && *stop) term = 1;
      k_lm_decide<<<P_.n_win, RCTA, 0, stream_>>>(P_, term);
      k_restore<<<cdiv(std::max(P_.n_pose, P_.n_point), 256), 256, 0, stream_>>>(P_);
      launches_ += 7;
      lm_trials_++;
      CU_CHECK(cudaMemcpyAsync(h_counters_, P_.counters, 3 * sizeof(int), cudaMemcpyDeviceToHost, stream_));
      CU_CHECK(cudaStreamSynchronize(stream_));
      CU_CHECK(cudaGetLastError());
      if (use_persist()) cg_iters_total_ = h_counters_[2];
      if (h_counters_[0] >= P_.n_win) break;
    }
    return SQRTBA_OK;
  }

  // ---- statistics: per-stage CUDA-event timing is accumulated lazily (events are read after the solve)
  void begin_stats() {
    launches_ = 0; lm_trials_ = 0; cg_iters_total_ = 0;
    for (auto& v : stage_ms_) v = 0.0;
    mv_used_ = 0;
    cudaMemsetAsync(d_counters_.p + 2, 0, 6 * sizeof(int), stream_);  // CG iterations, stop-flag scratch, published steps
    step_seq_ = 0;
    if (h_ctl_) { h_ctl_->seq = 0; h_ctl_->term = 0; }
    cudaEventRecord(ev0_, stream_);
  }
  void stage_begin(int s) {
    if (!stage_timing_) return;
    cudaEventRecord(stage_ev_[2 * s], stream_);
  }
  void stage_end(int s) {
    if (!stage_timing_) return;
    cudaEventRecord(stage_ev_[2 * s + 1], stream_);
    cudaEventSynchronize(stage_ev_[2 * s + 1]);
    float ms = 0;
    cudaEventElapsedTime(&ms, stage_ev_[2 * s], stage_ev_[2 * s + 1]);
    stage_ms_[s] += ms;
  }
  int finish_stats(sqrtba_stats* st) {
    CU_CHECK(cudaEventRecord(ev1_, stream_));
    CU_CHECK(cudaEventSynchronize(ev1_));
    CU_CHECK(cudaGetLastError());
    float ms = 0;
    CU_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
    if (st) {
      std::memset(st, 0, sizeof *st);
      st->n_windows = P_.n_win;
      st->lm_trials = lm_trials_;
      st->cg_iters_total = cg_iters_total_;
      st->kernel_launches = launches_;
      st->ms_total = ms;
      st->ms_linearize = stage_ms_[0];
      st->ms_qr = stage_ms_[1];
      st->ms_pcg = stage_ms_[2];
      st->ms_backsub = stage_ms_[3];
      st->ms_cost = stage_ms_[4];
      double mv = 0.0;
      for (int i = 0; i + 1 < mv_used_; i += 2) {
        float t = 0;
        if (cudaEventElapsedTime(&t, mv_events_[i], mv_events_[i + 1]) == cudaSuccess) mv += t;
      }
      st->ms_matvec = mv;  // sum over the matvec launches of this solve (0 when the persistent PCG kernel ran)
      st->reserved[0] = use_persist() ? 1.0 : 0.0;           // the PCG solves ran in the persistent cooperative kernel
      st->reserved[1] = (comm_ && peer_ok_) ? 1.0 : 0.0;     // landmark-sharded: in-kernel NVLink exchange available
      st->reserved[2] = persist_grid_;
      st->reserved[4] = eff_rtol();                          // PCG tolerance in effect
      st->reserved[7] = (chunk_on() && coarse_active_) ? 1.0 : 0.0;  // ... with the coarse (second) level
      st->reserved[6] = chunk_on() ? 1.0 : 0.0;              // big window: the 20-pose chunk preconditioner was active
      st->reserved[5] = det_active_ ? 1.0 : 0.0;             // reproducible mode active (pcg_mode = 4 and the problem qualifies)
    }
    return SQRTBA_OK;
  }

  void release_all() {
    d_cam_.release(); d_slot_cam_.release(); d_pose_slot_.release(); d_slot_pose_.release(); d_slot_win_.release(); d_pose_win_.release();
    d_point_win_.release(); d_meas_.release(); d_obs_pose_.release(); d_obs_point_.release(); d_obs_slot_.release();
    d_item_start_.release(); d_item_cnt_.release(); d_item_win_.release(); d_win_item_ptr_.release();
    d_win_slot_ptr_.release(); d_pose_.release(); d_pose0_.release(); d_pose_bak_.release(); d_point_.release();
    d_point0_.release(); d_point_bak_.release(); d_level_.release(); d_outlier_.release(); d_err_.release();
    d_JQ_.release(); d_Jl_.release(); d_r_.release(); d_R_.release(); d_tl_.release();
    d_bl_.release(); d_dl_.release(); d_slotvec_.release(); d_chi_part_.release(); d_scale_part_.release();
    d_ctl_.release(); d_trace_.release(); d_counters_.release(); d_wred_.release();
    d_tiles_.release(); d_obs_lp_.release(); d_tile_run_ptr_.release(); d_tile_runs_.release();
    d_win_tile_ptr_.release(); d_part_lin_.release(); d_part_qr_.release(); d_part_q_.release();
    d_gbar_.release(); d_part_.release(); d_q3_.release(); d_dq_.release(); d_ptile_.release();
    d_chunk_M_.release(); d_chunk_rz_.release(); d_chunk_pack_.release(); d_chunk_diag_.release();
    d_co_dblk_.release(); d_co_R_.release(); d_co_aci_.release(); d_co_w_.release(); d_co_ctl_.release();
    d_l_pc_.release(); d_l_qw_.release(); d_l_nv_.release(); d_l_w_.release(); d_l_acc_.release(); d_l_match_.release();
    d_lm_flat_pose_.release(); d_lm_corner_pose_.release(); d_lc_flat_.release(); d_lc_normal_.release(); d_lc_corner_.release();
    d_lc_world_.release(); d_lm_flat_.release(); d_lm_flat_w_.release(); d_lm_corner_.release(); d_lm_corner_w_.release(); d_l_best_.release();
    d_po_ptr_.release(); d_po_pose_.release(); d_po_cam_.release(); d_po_xyz_.release(); d_po_err_.release(); d_po_trace_.release();
    d_po_meas_.release(); d_s3_meas_.release(); d_fl_pc_.release(); d_fl_qw_.release(); d_fl_nv_.release(); d_fl_w_.release(); d_fl_match_.release();
    d_fl_best_.release(); d_fl_world_.release(); d_fl_flat_.release(); d_fl_normal_.release(); d_fl_corner_.release(); d_fl_map_.release(); d_po_level_.release(); d_po_outlier_.release(); d_po_inl_.release();
    d_pg_vert_.release(); d_pg_bak_.release(); d_pg_meas_.release(); d_pg_err_.release(); d_pg_Ji_.release(); d_pg_Jj_.release();
    d_pg_H_.release(); d_pg_L_.release(); d_pg_Linv_.release(); d_pg_b_.release(); d_pg_y_.release(); d_pg_x_.release();
    d_pg_scal_.release(); d_pg_fixed_.release(); d_pg_slot_.release(); d_pg_slot_vert_.release(); d_pg_edge_.release();
    d_pg_first_.release(); d_pg_col_ptr_.release(); d_pg_col_rows_.release(); d_pg_rowptr_.release();
    d_pg_inc_ptr_.release(); d_pg_inc_edge_.release(); d_pg_chi_part_.release();
    h_obs_slot_.release(); h_item_start_.release(); h_item_cnt_.release(); h_item_win_.release();
    h_tile_run_ptr_.release(); h_tile_runs_.release(); h_obs_lp_.release(); h_tiles_pin_.release();
    h_perm_pose_.release(); h_perm_point_.release(); h_perm_slot_.release(); h_perm_meas_.release(); h_perm_xyz_.release();
  }

 public:
  bool stage_timing_ = false;
  bool plan_only_ = false;
  HostPool pool_;  // helper threads of set_problem, created on first use and kept for the life of the handle
  int plan_n_item_ = 0, plan_n_tile_ = 0, plan_smallwin_ = 0, plan_pq_shared_ = 0;
  long long plan_n_runs_ = 0, plan_jq_total_ = 0;

 private:
  sqrtba_config cfg_;
  std::string err_;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  cudaEvent_t stage_ev_[10] = {};
  double stage_ms_[5] = {};
  int* h_counters_ = nullptr;
  volatile HostCtl* h_ctl_ = nullptr;   // mapped pinned memory: the device publishes every macro step here
  HostCtl* d_ctl_host_ = nullptr;       // its device-side address
  cudaGraphExec_t step_exec_ = nullptr; // the captured LM macro step of single-window problems
  bool step_graph_valid_ = false, graph_ok_ = true;
  double step_d2_ = 0.0, step_d3_ = 0.0;
  int step_seq_ = 0;
  bool have_problem_ = false;
  int max_trace_ = 200;
  int launches_ = 0, lm_trials_ = 0, cg_iters_total_ = 0;
  int n_sm_ = 148, pipe_ctas_ = 296, pipe_stages_ = 3, max_win_slots_ = 1, pipe_slots_ = 1;
  std::vector<int> perm_, old_first_, new_first_, lm_count_new_;  // landmark re-ordering of big windows: new -> old, first observation in either order
  void* comm_ = nullptr;  // ncclComm_t
  int n_ranks_ = 1, rank_ = 0;
  bool coop_ok_ = false, peer_ok_ = false;
  int persist_ctas_ = 0, persist_stages_ = 2, persist_slots_ = 1, persist_grid_ = 0;
  std::vector<cudaEvent_t> mv_events_;
  int mv_used_ = 0;
  DBuf<int> d_ptile_, d_win_tile_ptr_;
  DBuf<double> d_part_lin_, d_part_qr_, d_part_q_;
  bool det_active_ = false;
  double* peer_recv_[8] = {};
  unsigned long long* d_seq_ = nullptr;
  void** d_peer_tbl_ = nullptr;
  size_t peer_nelem_cap_ = 0;
  DBuf<unsigned> d_gbar_;
  DBuf<double> d_part_, d_q3_, d_dq_;
  bool chunk_active_ = false;  // big single window: 20-pose blocks as the PCG preconditioner (sqrtba_chunkprec.cuh)
  bool coarse_active_ = false;  // ... plus the coarse correction over the chunks (second level)
  DBuf<double> d_chunk_rz_, d_co_dblk_, d_co_R_, d_co_aci_, d_co_w_;
  DBuf<unsigned> d_co_ctl_;
  DBuf<float> d_chunk_M_, d_chunk_pack_, d_chunk_diag_;
  // lidar pass
  LidarDev lidar_{};
  LidarAssoc assoc_{};
  bool lidar_edges_set_ = false, lidar_assoc_set_ = false, lidar_active_ = false;
  DBuf<double> d_l_pc_, d_l_qw_, d_l_nv_, d_l_w_, d_l_acc_;
  DBuf<int> d_l_match_, d_lm_flat_pose_, d_lm_corner_pose_;
  DBuf<float> d_lc_flat_, d_lc_normal_, d_lc_corner_, d_lc_world_, d_lm_flat_, d_lm_flat_w_, d_lm_corner_, d_lm_corner_w_;
  DBuf<unsigned long long> d_l_best_;
  DBuf<long long> d_po_ptr_;
  DBuf<double> d_po_pose_, d_po_cam_, d_po_xyz_, d_po_err_, d_po_trace_;
  DBuf<float4> d_po_meas_;
  DBuf<uint8_t> d_po_level_, d_po_outlier_;
  DBuf<int> d_po_inl_;
  int po_frames_ = 0;
  DBuf<float> d_s3_meas_;
  DBuf<double> d_fl_pc_, d_fl_qw_, d_fl_nv_, d_fl_w_;
  DBuf<int> d_fl_match_;
  DBuf<unsigned long long> d_fl_best_;
  DBuf<float> d_fl_world_, d_fl_flat_, d_fl_normal_, d_fl_corner_, d_fl_map_;
  int s3_pairs_ = 0;
  // essential-graph optimisation (independent of set_problem)
  DBuf<double> d_pg_vert_, d_pg_bak_, d_pg_meas_, d_pg_err_, d_pg_Ji_, d_pg_Jj_, d_pg_H_, d_pg_L_, d_pg_Linv_, d_pg_b_, d_pg_y_,
      d_pg_x_, d_pg_scal_;
  DBuf<uint8_t> d_pg_fixed_;
  DBuf<int> d_pg_slot_, d_pg_slot_vert_, d_pg_edge_, d_pg_first_, d_pg_col_ptr_, d_pg_col_rows_, d_pg_inc_ptr_, d_pg_inc_edge_;
  DBuf<double> d_pg_chi_part_;
  DBuf<long long> d_pg_rowptr_;
  std::vector<double> pg_trace_;
  Dev P_{};
  DBuf<double> d_cam_, d_pose_, d_pose0_, d_pose_bak_, d_point_, d_point0_, d_point_bak_, d_err_, d_JQ_, d_Jl_,
      d_r_, d_R_, d_tl_, d_bl_, d_dl_, d_slotvec_, d_chi_part_, d_scale_part_, d_trace_, d_wred_;
  DBuf<double> d_slot_cam_;
  std::vector<double> h_slot_cam_;
  DBuf<int> d_pose_slot_, d_slot_pose_, d_slot_win_, d_pose_win_, d_point_win_, d_obs_pose_, d_obs_point_, d_obs_slot_,
      d_item_start_, d_item_cnt_, d_item_win_, d_win_item_ptr_, d_win_slot_ptr_, d_counters_, d_tile_run_ptr_,
      d_tile_runs_;
  DBuf<unsigned> d_obs_lp_;
  DBuf<TileInfo> d_tiles_;
  DBuf<long long> d_prof_;
  HBuf<int> h_obs_slot_, h_item_start_, h_item_cnt_, h_item_win_, h_tile_run_ptr_, h_tile_runs_;
  HBuf<unsigned> h_obs_lp_;
  HBuf<int> h_perm_pose_, h_perm_point_, h_perm_slot_;
  HBuf<float> h_perm_meas_;
  HBuf<double> h_perm_xyz_;
  HBuf<TileInfo> h_tiles_pin_;
  std::vector<TileInfo> h_tiles_;
  DBuf<float4> d_meas_;
  DBuf<uint8_t> d_level_, d_outlier_;
  DBuf<WinCtl> d_ctl_;
};

}  // namespace sqrtba

// =================================================================================================== C ABI
using sqrtba::Solver;

struct sqrtba_handle {
  Solver* s;
  std::string err;
};

static thread_local std::string g_create_err;

extern "C" {

const char* sqrtba_version(void) { return "sqrtba 0.1 (sm_100a)"; }

int sqrtba_default_config(sqrtba_config* cfg) {
  if (!cfg) return SQRTBA_ERR_INVALID;
  std::memset(cfg, 0, sizeof *cfg);
  cfg->device = 0;
  cfg->pcg_rtol = 0.0;  // automatic, see sqrtba.h
  cfg->pcg_max_iters = 2000;
  cfg->third_pass_iters = 0;
  cfg->pcg_mode = 0;
  cfg->pcg_check_every = 4;
  return SQRTBA_OK;
}

int sqrtba_create(const sqrtba_config* cfg, sqrtba_handle** out) {
  if (!out) return SQRTBA_ERR_INVALID;
  *out = nullptr;
  sqrtba_config c;
  if (cfg) c = *cfg; else sqrtba_default_config(&c);
  if (!(c.pcg_rtol > 0)) c.pcg_rtol = 0.0;  // automatic
  if (c.pcg_max_iters <= 0) c.pcg_max_iters = 2000;
  if (c.pcg_check_every <= 0) c.pcg_check_every = 4;
  sqrtba_handle* h = new (std::nothrow) sqrtba_handle();
  if (!h) return SQRTBA_ERR_ALLOC;
  h->s = new (std::nothrow) Solver(c);
  if (!h->s) { delete h; return SQRTBA_ERR_ALLOC; }
  h->s->stage_timing_ = c.reserved[0] != 0;
  const int rc = h->s->init();
  if (rc) {
    g_create_err = h->s->last_error();
    delete h->s;
    delete h;
    return rc;
  }
  *out = h;
  return SQRTBA_OK;
}

int sqrtba_destroy(sqrtba_handle* h) {
  if (!h) return SQRTBA_ERR_INVALID;
  delete h->s;
  delete h;
  return SQRTBA_OK;
}

const char* sqrtba_last_error(const sqrtba_handle* h) { return h ? h->s->last_error() : g_create_err.c_str(); }

int sqrtba_set_problem(sqrtba_handle* h, int32_t n_pose, int32_t n_point, int32_t n_obs, const double* pose_qt,
                       const uint8_t* pose_fixed, const double* cam, const double* point_xyz,
                       const int32_t* obs_pose, const int32_t* obs_point, const float* obs_meas) {
  if (!h) return SQRTBA_ERR_INVALID;
  return h->s->set_problem(1, nullptr, nullptr, nullptr, n_pose, n_point, n_obs, pose_qt, pose_fixed, cam, point_xyz,
                           obs_pose, obs_point, obs_meas);
}

int sqrtba_set_problem_batch(sqrtba_handle* h, int32_t n_win, const int64_t* win_pose_ptr,
                             const int64_t* win_point_ptr, const int64_t* win_obs_ptr, const double* pose_qt,
                             const uint8_t* pose_fixed, const double* cam, const double* point_xyz,
                             const int32_t* obs_pose, const int32_t* obs_point, const float* obs_meas) {
  if (!h || n_win <= 0 || !win_pose_ptr || !win_point_ptr || !win_obs_ptr) return SQRTBA_ERR_INVALID;
  return h->s->set_problem(n_win, win_pose_ptr, win_point_ptr, win_obs_ptr, (int)win_pose_ptr[n_win],
                           (int)win_point_ptr[n_win], (int)win_obs_ptr[n_win], pose_qt, pose_fixed, cam, point_xyz,
                           obs_pose, obs_point, obs_meas);
}

int sqrtba_reset_state(sqrtba_handle* h) { return h ? h->s->reset_state() : SQRTBA_ERR_INVALID; }
int sqrtba_solve_local(sqrtba_handle* h, const volatile bool* stop_flag, sqrtba_stats* stats) {
  return h ? h->s->solve_local(stop_flag, stats) : SQRTBA_ERR_INVALID;
}
int sqrtba_solve_global(sqrtba_handle* h, int32_t iters, int32_t robust, const volatile bool* stop_flag,
                        sqrtba_stats* stats) {
  return h ? h->s->solve_global(iters, robust, stop_flag, stats) : SQRTBA_ERR_INVALID;
}
int sqrtba_get_poses(sqrtba_handle* h, double* out) { return (h && out) ? h->s->get_poses(out) : SQRTBA_ERR_INVALID; }
int sqrtba_get_points(sqrtba_handle* h, double* out) { return (h && out) ? h->s->get_points(out) : SQRTBA_ERR_INVALID; }
int sqrtba_get_outliers(sqrtba_handle* h, uint8_t* out) { return (h && out) ? h->s->get_outliers(out) : SQRTBA_ERR_INVALID; }
int sqrtba_get_trace_len(sqrtba_handle* h, int32_t window) { return h ? h->s->trace_len(window) : SQRTBA_ERR_INVALID; }
int sqrtba_get_trace(sqrtba_handle* h, int32_t window, sqrtba_trace_row* rows_out, int32_t max_rows) {
  return (h && rows_out) ? h->s->get_trace(window, rows_out, max_rows) : SQRTBA_ERR_INVALID;
}
int sqrtba_debug_linearize(sqrtba_handle* h, int32_t huber, double* err, double* Jp, double* Jl, double* r,
                           double* chi2) {
  return h ? h->s->debug_linearize(huber, err, Jp, Jl, r, chi2) : SQRTBA_ERR_INVALID;
}
int sqrtba_debug_step(sqrtba_handle* h, double lambda, double* dp, double* dl, double* bs, int32_t* cg_iters) {
  return h ? h->s->debug_step(lambda, dp, dl, bs, cg_iters) : SQRTBA_ERR_INVALID;
}
int sqrtba_debug_matvec(sqrtba_handle* h, const double* p, double* y) {
  return (h && p && y) ? h->s->debug_matvec(p, y) : SQRTBA_ERR_INVALID;
}
int sqrtba_num_free_poses(sqrtba_handle* h) { return h ? h->s->num_free() : SQRTBA_ERR_INVALID; }
int sqrtba_debug_plan(int32_t n_win, const int64_t* win_pose_ptr, const int64_t* win_point_ptr, const int64_t* win_obs_ptr,
                      int32_t n_pose, int32_t n_point, int32_t n_obs, const uint8_t* pose_fixed, const int32_t* obs_pose,
                      const int32_t* obs_point, int32_t host_threads, int32_t* summary8, int32_t* tiles20, int32_t max_tiles,
                      uint32_t* obs_lp, int32_t* tile_run_ptr, int32_t* tile_runs, int64_t max_runs, int32_t* landmark_order) {
  if (!pose_fixed || !obs_pose || !obs_point || n_pose <= 0 || n_point <= 0 || n_obs <= 0 || n_win <= 0) return SQRTBA_ERR_INVALID;
  sqrtba_config c;
  sqrtba_default_config(&c);
  c.reserved[3] = host_threads;
  Solver s(c);
  s.plan_only_ = true;
  // the plan never reads values, only indices: give the value arrays harmless stand-ins
  std::vector<double> pose((size_t)n_pose * 7, 0.0), cam((size_t)n_pose * 5, 0.0), xyz((size_t)n_point * 3, 0.0);
  std::vector<float> meas((size_t)n_obs * 4, 0.f);
  const int rc = s.set_problem(n_win, win_pose_ptr, win_point_ptr, win_obs_ptr, n_pose, n_point, n_obs, pose.data(), pose_fixed,
                               cam.data(), xyz.data(), obs_pose, obs_point, meas.data());
  if (rc != SQRTBA_OK) return rc;
  s.plan_export(summary8, tiles20, max_tiles, obs_lp, tile_run_ptr, tile_runs, max_runs, landmark_order, n_obs, n_point);
  return SQRTBA_OK;
}
int sqrtba_pose_opt(sqrtba_handle* h, int32_t n_frames, const int64_t* frame_obs_ptr, double* pose_qt, const double* cam,
                    const double* obs_xyz, const float* obs_meas, uint8_t* outlier_out, int32_t* inliers_out,
                    sqrtba_stats* stats) {
  return h ? h->s->pose_opt(n_frames, frame_obs_ptr, pose_qt, cam, obs_xyz, obs_meas, outlier_out, inliers_out, stats)
           : SQRTBA_ERR_INVALID;
}
int sqrtba_pose_opt_lidar(sqrtba_handle* h, double* pose_qt, const double* cam, int32_t n_obs, const double* obs_xyz,
                          const float* obs_meas, uint8_t* outlier_out, int32_t* inliers_out, const sqrtba_frame_lidar* lidar,
                          int32_t* n_match2_out, sqrtba_stats* stats) {
  return h ? h->s->pose_opt_lidar(pose_qt, cam, n_obs, obs_xyz, obs_meas, outlier_out, inliers_out, lidar, n_match2_out, stats)
           : SQRTBA_ERR_INVALID;
}
int sqrtba_pose_opt_trace(sqrtba_handle* h, int32_t frame, double* rows_out, int32_t max_rows) {
  return h ? h->s->pose_opt_trace(frame, rows_out, max_rows) : SQRTBA_ERR_INVALID;
}
int sqrtba_pose_graph(sqrtba_handle* h, int32_t n_vert, double* vert8, const uint8_t* fixed, int32_t fix_scale, int32_t n_edge,
                      const int32_t* edge_ij, const double* meas8, int32_t iters, double lambda_init, sqrtba_stats* stats) {
  return h ? h->s->pose_graph(n_vert, vert8, fixed, fix_scale, n_edge, edge_ij, meas8, iters, lambda_init, stats)
           : SQRTBA_ERR_INVALID;
}
int sqrtba_pose_graph_trace(sqrtba_handle* h, double* rows_out, int32_t max_rows) {
  return h ? h->s->pose_graph_trace(rows_out, max_rows) : SQRTBA_ERR_INVALID;
}
int sqrtba_optimize_sim3(sqrtba_handle* h, int32_t n_pairs, const int64_t* pair_match_ptr, double* sim3_12, const double* cam8,
                         const double* p1c, const double* p2c, const float* match_meas, float th2, int32_t fix_scale,
                         uint8_t* keep_out, int32_t* inliers_out, sqrtba_stats* stats) {
  return h ? h->s->optimize_sim3(n_pairs, pair_match_ptr, sim3_12, cam8, p1c, p2c, match_meas, th2, fix_scale, keep_out,
                                 inliers_out, stats)
           : SQRTBA_ERR_INVALID;
}
int sqrtba_optimize_sim3_trace(sqrtba_handle* h, int32_t pair, double* rows_out, int32_t max_rows) {
  return h ? h->s->optimize_sim3_trace(pair, rows_out, max_rows) : SQRTBA_ERR_INVALID;
}
int sqrtba_set_lidar(sqrtba_handle* h, const sqrtba_lidar* clouds) { return h ? h->s->set_lidar(clouds) : SQRTBA_ERR_INVALID; }
int sqrtba_set_lidar_edges(sqrtba_handle* h, int32_t cur_pose, int32_t n_flat, int32_t n_corner, const double* point_cam,
                           const double* point_world, const double* normal, const double* weight, int32_t numeric_jacobian) {
  return h ? h->s->set_lidar_edges(cur_pose, n_flat, n_corner, point_cam, point_world, normal, weight, numeric_jacobian)
           : SQRTBA_ERR_INVALID;
}
int sqrtba_get_lidar_matches(sqrtba_handle* h, int32_t* match_out) { return h ? h->s->get_lidar_matches(match_out) : SQRTBA_ERR_INVALID; }
int sqrtba_num_lidar_edges(sqrtba_handle* h) { return h ? h->s->num_lidar_edges() : SQRTBA_ERR_INVALID; }
int sqrtba_comm_unique_id(uint8_t* id128) {
  if (!id128) return SQRTBA_ERR_INVALID;
  if (!sqrtba::g_nccl.load()) return SQRTBA_ERR_COMM;
  sqrtba::NcclApi::UniqueId id;
  if (sqrtba::g_nccl.GetUniqueId(&id) != 0) return SQRTBA_ERR_COMM;
  std::memcpy(id128, id.internal, 128);
  return SQRTBA_OK;
}
int sqrtba_comm_init(sqrtba_handle* h, int32_t nranks, int32_t rank, const uint8_t* id128) {
  return h ? h->s->comm_init(nranks, rank, id128) : SQRTBA_ERR_INVALID;
}
int sqrtba_comm_destroy(sqrtba_handle* h) {
  if (!h) return SQRTBA_ERR_INVALID;
  h->s->comm_destroy();
  return SQRTBA_OK;
}
int sqrtba_time_stage(sqrtba_handle* h, int32_t stage, int32_t warmup, int32_t reps, double* ms_avg) {
  return (h && ms_avg) ? h->s->time_stage(stage, warmup, reps, ms_avg) : SQRTBA_ERR_INVALID;
}

}  // extern "C"

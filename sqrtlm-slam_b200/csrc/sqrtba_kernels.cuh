// CUDA kernels of the square-root LM bundle-adjustment path (sm_100a).
//
// Work decomposition.  Observations arrive grouped by landmark.  The host packs whole landmarks into "items"
// of at most 32 observations; one warp owns one item, one lane owns one observation, so every per-landmark
// reduction (Householder norms, Q1^T v, J_l^T r ...) is a segmented warp-shuffle reduction and every
// per-observation plane access is a fully coalesced stream.  A landmark with more than 32 observations gets
// an item of its own and the warp sweeps it in chunks of 32.  Items never straddle windows (batched BA).
//
// Storage is SoA by component ("planes" of ld doubles): Jl 9, r 3, err 3; the matvec operand (Jp 18 + Q1 9) is
// tile-blocked SoA (see Dev::JQ).
// All arithmetic is FP64 (tolerances 1e-6 on cost / 1e-5 m on poses); the roofline is HBM bandwidth.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sqrtba_math.cuh"

namespace sqrtba {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CTA = 128;            // item kernels: one CTA = one tile of up to WARPS items (warps)
constexpr int WARPS = CTA / 32;
constexpr int RCTA = 256;           // per-window reduction / vector kernels
constexpr int RWARPS = RCTA / 32;
constexpr int MAXSLOT = 128;        // "small window": all free poses of a window fit the CTA's shared accumulators
constexpr int JG = 4;               // geometry rows of the matvec operand: {x/z, y/z, 1/z, w} per observation -- the weighted
                                    // 3x6 pose Jacobian is rebuilt from them in registers (jp_compact, sqrtba_math.cuh)
constexpr int NPLANE = JG + 9;      // matvec streams the geometry rows (4) + Q1 (9): 104 bytes per observation
constexpr unsigned LP_STEREO = 0x80000000u;  // obs_lp / meta bit: the edge is a stereo edge (third residual row exists)
constexpr int JQ_HDR = 8;           // doubles (64 bytes) of per-tile header in front of every JQ data block
constexpr int JQ_ROWS = NPLANE + 1; // data planes + one 8-byte meta entry per column

enum Phase { PH_LIN = 0, PH_TRIAL = 1, PH_DONE = 2 };

constexpr int TRACE_COLS = 10;

struct WinCtl {
  double lambda, ni, cur_chi, ini_chi, tmp_chi, rho;
  unsigned long long maxdiag_bits;  // landmark-side max |diag(Hll)| as ordered bits (non-negative doubles)
  double rz, rz0;
  double tol2;  // square of the PCG tolerance of this pass (set by k_pass_init; kernels launched with tol2 <= 0 read it here)
  int iter, qmax, nbad, phase;
  int max_iter, pass, cg_active, cg_iters;
  int trace_len, need_restore, lin_count;
  int robust;  // Huber kernels on in this pass (set by k_pass_init; kernels launched with robust < 0 read it here)
};

struct TileInfo {
  int item0, nitem;  // items [item0, item0+nitem)
  int o0, o1;        // observation range [o0, o1)
  int win;           // window of the tile
  int nfree;         // observations that take part in the in-CTA reduction (ranks 0..nfree-1)
  int nt;            // columns of this tile's JQ block, rounded up to even (16-byte rows for TMA): one per FREE-pose
                     // observation (short tiles) or one per observation (long tiles)
  int is_long;       // the tile is one landmark with more than 32 observations
  long long jq_off;  // offset (doubles) of the tile's [27][nt] data block inside Dev::JQ (its header sits right before)
  int nrun;          // distinct free pose slots among the participating observations (= runs of equal slot by rank)
  int blk_doubles;   // size of the whole JQ block (header + data + meta + run table), in doubles, even
  int cnt[4];        // observations of each item of the tile (0 for absent items): saves the item_start / item_cnt lookups
  int fcnt[4];       // free-pose observations of each item: the JQ block of a short tile only has columns for those
};

// item geometry from the tile descriptor alone (items of a tile are consecutive observation ranges; selects, not
// indexing, so the descriptor stays in registers)
__device__ __forceinline__ int tile_item_cnt(const TileInfo& ti, int wid) {
  return wid == 0 ? ti.cnt[0] : (wid == 1 ? ti.cnt[1] : (wid == 2 ? ti.cnt[2] : (wid == 3 ? ti.cnt[3] : 0)));
}
__device__ __forceinline__ int tile_item_start(const TileInfo& ti, int wid) {
  return ti.o0 + (wid > 0 ? ti.cnt[0] : 0) + (wid > 1 ? ti.cnt[1] : 0) + (wid > 2 ? ti.cnt[2] : 0);
}

// Column of a free-pose observation inside its tile's JQ block.  Short tiles store columns for FREE-pose observations
// only (in landmark order): observations of fixed keyframes have no pose columns (g2o hessianIndex -1), so neither the
// matvec nor the back-substitution ever reads them -- leaving them out cuts the dominant kernel's DRAM traffic by the
// share of fixed-keyframe observations (14 % on the KITTI-shaped windows).  Must be called by all lanes of the warp.
__device__ __forceinline__ int tile_fcol(const TileInfo& ti, int wid, bool has, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, has);
  return (wid > 0 ? ti.fcnt[0] : 0) + (wid > 1 ? ti.fcnt[1] : 0) + (wid > 2 ? ti.fcnt[2] : 0) + __popc(m & ((1u << lane) - 1u));
}

struct Dev {
  int n_pose, n_point, n_obs, n_win, n_slot, n_item;
  // ---- problem (static)
  const double* cam;       // n_pose*5
  const int* pose_slot;    // n_pose: free slot or -1
  const int* slot_pose;    // n_slot
  const int* slot_win;     // n_slot
  const int* pose_win;     // n_pose
  const int* point_win;    // n_point
  const float4* obs_meas;  // n_obs
  const int* obs_pose;     // n_obs
  const int* obs_point;    // n_obs
  const int* obs_slot;     // n_obs: pose_slot[obs_pose]
  const int* item_start;   // n_item
  const int* item_cnt;     // n_item
  const int* item_win;     // n_item
  const int* win_item_ptr;  // n_win+1
  const int* win_slot_ptr;  // n_win+1
  const double* slot_cam;   // n_slot*3: fx, fy, bf of the keyframe in that free slot (for jp_compact)
  // tile = up to WARPS consecutive short items of ONE window (or one long item); one CTA per tile.
  // obs_lp: low 16 bits = window-relative free slot when the problem is "smallwin" (0 otherwise), 0xffff = the
  // observation takes no part in the in-CTA reduction (fixed pose or long item); bits 16-30 = its rank in the
  // tile's pose-sorted order; bit 31 (LP_STEREO) = stereo edge.
  int n_tile;
  int ld;                     // leading dimension (stride) of every per-observation plane, multiple of 32
  int smallwin;               // window-relative slots fit 16 bits (every window has < 65535 free poses): obs_lp and the
                              // run tables carry window-relative slots and the TMA-pipelined matvec is used
  int pq_shared;              // every window has <= MAXSLOT free poses: p and q of a window live in shared memory
  int det;                    // reproducible mode (pcg_mode = 4): pose-side sums leave the CTAs as per-(CTA, window)
                              // partial vectors that are added up in a fixed order, never through atomics
  int maxslot;                // det: free poses of the largest window (stride of the partial vectors)
  double* part_lin;           // det: [(linearise CTA + window)][12 * maxslot]  b_p | diag(Jp^T Jp) per slot
  double* part_qr;            // det: [(QR CTA + window)][27 * maxslot]         reduced rhs | block-Jacobi block per slot
  double* part_q;             // det: [(matvec CTA + window)][6 * maxslot]      partial q per slot
  const int* win_tile_ptr;    // n_win + 1: tile range of every window
  int fused;                  // k_linqr_pipe serves the iterations with a known lambda and the retries; the separate
                              // linearise / QR kernels only the first trial of a pass (see lin_phase / qr_phase)
  const struct TileInfo* tiles;
  const unsigned* obs_lp;     // n_obs
  const int* tile_run_ptr;    // n_tile+1 -> tile_runs (host-built run tables, copied into the JQ blocks by k_init_jq)
  const int* tile_runs;
  // ---- state
  double* pose;       // n_pose*7 (t,q)
  double* point;      // n_point*3
  double* pose_bak;
  double* point_bak;
  uint8_t* obs_level;  // n_obs (0 active, 1 excluded)
  uint8_t* obs_outlier;
  // ---- linearisation (planes)
  double* err;  // 3 planes: g2o's stored _error (unweighted)
  double* Jl;   // 9
  double* r;    // 3
  // matvec operand, TILE-BLOCKED: tile t owns one contiguous block of JQ:
  //   [JQ_HDR doubles header][13][nt] data][nt x 8-byte per-column meta][run table: nrun+1 offsets, nrun slots (int)]
  // run table: ranks [run_ptr[r], run_ptr[r+1]) all belong to pose slot run_slot[r] (window-relative when smallwin)
  // data rows 0-3 = {x/z, y/z, 1/z, w} (the weighted 3x6 Jp is a function of these and of the keyframe's fx, fy, bf:
  // jp_compact), rows 4-12 = observation rows of Q1 (3x3 row-major), column = free-pose observation;
  // meta = {obs_lp, landmark id}; header = 16 ints (see k_init_jq).  jq_off points at the data part.
  // One tile = one contiguous chunk, so the matvec fetches everything it needs with a single TMA bulk copy.
  double* JQ;
  // ---- per landmark (planes of n_point)
  double* R;   // 6: r00 r01 r02 r11 r12 r22
  double* tl;  // 3: Q1^T r
  double* bl;  // 3: -Jl^T r
  double* dl;  // 3: landmark step
  // ---- per slot
  double* bp;   // 6*n_slot  -Jp^T r
  double* hd;   // 6*n_slot  diag(Jp^T Jp)
  double* bs;   // 6*n_slot  reduced rhs
  double* D;    // 21*n_slot upper triangle of the block-Jacobi block (without lambda)
  double* Dinv; // 36*n_slot
  double* x;    // 6*n_slot  pose step
  double* res;
  double* z;
  double* p;
  double* q;
  // ---- per item
  double* chi_part;    // n_item
  double* scale_part;  // n_item
  // ---- per window partial sums that cross ranks in the landmark-sharded (multi-GPU) mode:
  // wred[0*n_win + w] chi2, [1*n_win + w] landmark part of computeScale, [2*n_win + w] landmark-side max diagonal
  double* wred;
  // ---- control
  WinCtl* ctl;      // n_win
  double* trace;    // n_win * max_trace * TRACE_COLS
  int max_trace;
  int* counters;    // [0] windows done, [1] cg-active windows
  long long* prof;  // optional (SQRTBA_PIPE_PROF builds): per-CTA cycle counters of the pipelined matvec
};

// Division of labour between the separate kernels and the fused one (Dev::fused): a window's phase as the SEPARATE
// linearisation / landmark-QR kernels see it -- PH_DONE means "not mine this step".
__device__ __forceinline__ int lin_phase(const Dev& P, const WinCtl& c) {
  return (P.fused && c.phase == PH_LIN && c.iter > 0) ? (int)PH_DONE : c.phase;
}
__device__ __forceinline__ int qr_phase(const Dev& P, const WinCtl& c) {
  return (P.fused && c.phase == PH_TRIAL && (c.iter > 0 || c.qmax > 0)) ? (int)PH_DONE : c.phase;
}
// ... and which windows the fused kernel serves: 0 not mine, 1 FUSED (re-linearise + factorise), 2 RETRY (new lambda)
__device__ __forceinline__ int linqr_mode(const Dev& P, const WinCtl& c) {
  if (!P.fused) return 0;
  if (c.phase == PH_LIN) return c.iter > 0 ? 1 : 0;
  if (c.phase == PH_TRIAL) return c.qmax > 0 ? 2 : 0;
  return 0;
}

// ------------------------------------------------------------------------------------------------ warp helpers

struct Seg {
  int start, end;  // lane range [start,end) of this lane's landmark inside the warp
};

__device__ __forceinline__ Seg seg_of(int key, int lane) {
  const int prev = __shfl_up_sync(FULL, key, 1);
  const bool head = (lane == 0) || (key != prev);
  const unsigned heads = __ballot_sync(FULL, head);
  Seg s;
  s.start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
  const unsigned above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
  s.end = above ? (__ffs(above) - 1) : 32;
  return s;
}

// sum over the lane's segment, result broadcast to every lane of the segment (fixed order => deterministic)
__device__ __forceinline__ double seg_sum(double v, const Seg& s, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double t = __shfl_down_sync(FULL, v, off);
    if (lane + off < s.end) v += t;
  }
  return __shfl_sync(FULL, v, s.start);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, off));
  return v;
}

// reduction of one value over the whole landmark, for both item shapes:
//   short item (LONG=false): segmented over the lane's landmark;  long item: whole warp
template <bool LONG>
__device__ __forceinline__ double lm_sum(double v, const Seg& s, int lane) {
  if (LONG) return warp_sum(v);
  return seg_sum(v, s, lane);
}

__device__ __forceinline__ double block_sum(double v, double* sh) {  // RCTA threads; sh: RWARPS doubles; result to all
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < RWARPS; i++) t += sh[i];
  return t;
}

__device__ __forceinline__ void atomic_max_pos(unsigned long long* addr, double v) {
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

// Sum of one run c[a..b) of a shared-memory row.  Four independent partial sums: the run loop is a chain of dependent
// LDS + DADD (~40 cycles per element) and was the hottest line of the matvec (14 %) and of the landmark QR (22 %).
__device__ __forceinline__ double run_sum(const double* c, int a, int b) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int j = a;
  for (; j + 3 < b; j += 4) {
    s0 += c[j];
    s1 += c[j + 1];
    s2 += c[j + 2];
    s3 += c[j + 3];
  }
  if (j < b) s0 += c[j];
  if (j + 1 < b) s1 += c[j + 1];
  if (j + 2 < b) s2 += c[j + 2];
  return (s0 + s1) + (s2 + s3);
}

// CTA-level reduction of per-observation pose-side contributions.  Every participating observation of the tile owns
// one column `rank` of c_sh[NV][CTA]; ranks are ordered by pose slot, so the sum for one (slot, value) is a
// contiguous run that exactly one thread (the run head) adds up in a fixed order -> no shared-memory atomics, and the
// CTA issues one global atomicAdd per (tile, slot, value) instead of one per observation.  The runs come from the
// tile's run table (host-built, stored at the tail of its JQ block).  Must be called by all threads of the CTA.
template <int NV>
__device__ __forceinline__ void tile_scatter(const int* __restrict__ run_ptr, const int* __restrict__ run_slot, int nrun,
                                             int slot_base, double* c_sh, const double* vals, bool has, int rank,
                                             double* __restrict__ target, int stride, int offset) {
  if (has) {
#pragma unroll
    for (int k = 0; k < NV; k++) c_sh[k * CTA + rank] = vals[k];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nrun * 8; idx += CTA) {  // 8 threads per run, value k = idx & 7
    const int r = idx >> 3, k = idx & 7;
    if (k < NV) {
      const int a = run_ptr[r], b = run_ptr[r + 1];
      const double sum = run_sum(c_sh + k * CTA, a, b);
      atomicAdd(&target[(size_t)(slot_base + run_slot[r]) * stride + offset + k], sum);
    }
  }
  __syncthreads();
}

// Single-pass variant for kernels that reduce several per-observation vectors at once: every participating observation
// has already written its NV values into column `rank` of c_sh[NV][SCST] (SCST = CTA + 1: the value-lanes of one run
// fall into different banks); one thread per (run, value) adds the run up and issues ONE global atomic.  `addr(slot, k)`
// maps value k of pose slot `slot` to its accumulator.  Must be called by all threads of the CTA after a __syncthreads().
constexpr int SCST = CTA + 1;
template <int NV, class Addr>
__device__ __forceinline__ void tile_scatter_all(const int* __restrict__ run_ptr, const int* __restrict__ run_slot, int nrun,
                                                 int slot_base, const double* c_sh, Addr addr) {
  for (int idx = threadIdx.x; idx < nrun * NV; idx += CTA) {
    const int r = idx / NV, k = idx - r * NV;
    const int a = run_ptr[r], b = run_ptr[r + 1];
    const double sum = run_sum(c_sh + k * SCST, a, b);
    atomicAdd(addr(slot_base + run_slot[r], k), sum);
  }
}

// Reproducible mode: the run sums of a tile are added to the CTA's shared accumulators acc[slot * NV + k] (window-relative
// slot; a slot occurs once per tile, tiles follow each other behind a CTA barrier: fixed order), and a CTA's accumulators
// leave it as ONE partial vector per (CTA, window) visit -- visit = CTA + window is unique because both grow along a
// CTA's consecutive tiles -- that a per-window reduction adds up in visit order.
template <int NV>
__device__ __forceinline__ void tile_scatter_acc(const int* __restrict__ run_ptr, const int* __restrict__ run_slot, int nrun,
                                                 const double* c_sh, double* acc) {
  for (int idx = threadIdx.x; idx < nrun * NV; idx += CTA) {
    const int r = idx / NV, k = idx - r * NV;
    acc[run_slot[r] * NV + k] += run_sum(c_sh + k * SCST, run_ptr[r], run_ptr[r + 1]);
  }
}
template <int NV>
__device__ __forceinline__ void det_flush(double* __restrict__ part, int visit, int maxslot, int wn, double* acc) {
  __syncthreads();
  double* dst = part + (size_t)visit * NV * maxslot;
  for (int i = threadIdx.x; i < NV * wn; i += CTA) { dst[i] = acc[i]; acc[i] = 0.0; }
  __syncthreads();
}
// sum of element i over the visits [c_lo + win, c_hi + win] of a window, in order
__device__ __forceinline__ double det_sum(const double* __restrict__ part, int c_lo, int c_hi, int win, size_t stride, int i) {
  double v = 0.0;
  for (int c = c_lo; c <= c_hi; c++) v += part[(size_t)(c + win) * stride + i];
  return v;
}

// ---- TMA (bulk async copy) + mbarrier primitives, raw PTX (sm_90+/sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA prefetch of a contiguous global range into L2 (no shared-memory destination, no registers): used to pull the
// operands of a LATER phase of a kernel towards the SMs while the current phase computes.  size: multiple of 16 bytes.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}

// ---- cp.async (LDGSTS) staging primitives
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ K0: zeroing

// per-slot accumulators that the linearise kernel fills with atomics (only for windows that re-linearise)
__global__ void k_zero_lin(Dev P) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.n_slot) return;
  const WinCtl& c = P.ctl[P.slot_win[s]];
  if (c.phase == PH_LIN) {
#pragma unroll
    for (int k = 0; k < 6; k++) { P.bp[s * 6 + k] = 0.0; P.hd[s * 6 + k] = 0.0; }
  }
  if (linqr_mode(P, c) != 0) {  // the fused kernel also fills this trial's reduced rhs and block-Jacobi blocks
#pragma unroll
    for (int k = 0; k < 6; k++) P.bs[s * 6 + k] = 0.0;
#pragma unroll
    for (int k = 0; k < 21; k++) P.D[s * 21 + k] = 0.0;
  }
}

// per-slot accumulators of one trial (reduced rhs, block-Jacobi block) for the windows the separate QR kernel serves
__global__ void k_zero_trial(Dev P) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.n_slot) return;
  if (qr_phase(P, P.ctl[P.slot_win[s]]) != PH_TRIAL) return;
#pragma unroll
  for (int c = 0; c < 6; c++) P.bs[s * 6 + c] = 0.0;
#pragma unroll
  for (int c = 0; c < 21; c++) P.D[s * 21 + c] = 0.0;
}

// ------------------------------------------------------------------------------------------------ K1: linearise
// Fused residual + SE3/point Jacobians + Huber weighting (reference: computeActiveErrors + linearizeOplus +
// the weighting half of constructQuadraticForm; sparse_optimizer.cpp:61-88, types_six_dof_expmap.cpp:103-234,
// base_binary_edge.hpp:55-120).  Also produces what LM needs without ever forming H: chi2 (deterministic
// per-item partials), b_p = -Jp^T r, b_l = -Jl^T r, diag(Jp^T Jp), max diag(Jl^T Jl).
struct ObsLin {
  double e[3], Jl[9], w, rho0;
  double g[4];   // {x/z, y/z, 1/z, w}: what the matvec operand stores instead of the 3x6 pose block
  double cam3[3];  // fx, fy, bf of the observing keyframe
  bool depth_pos, stereo;
};

// the observation's own words (measurement, pose index, landmark index) already in registers
__device__ __forceinline__ void obs_eval_ops(const Dev& P, const float4 m, int ip, int il, bool want_jac, bool robust,
                                             double d2, double d3, ObsLin& L);
__device__ __forceinline__ void obs_eval(const Dev& P, int o, bool want_jac, bool robust, double d2, double d3,
                                         ObsLin& L) {
  obs_eval_ops(P, __ldg(&P.obs_meas[o]), __ldg(&P.obs_pose[o]), __ldg(&P.obs_point[o]), want_jac, robust, d2, d3, L);
}
__device__ __forceinline__ void obs_eval_ops(const Dev& P, const float4 m, int ip, int il, bool want_jac, bool robust,
                                             double d2, double d3, ObsLin& L) {
  double pose[7], X[3], cam[5], R[9], Xc[3];
#pragma unroll
  for (int i = 0; i < 7; i++) pose[i] = P.pose[ip * 7 + i];
#pragma unroll
  for (int i = 0; i < 3; i++) X[i] = P.point[il * 3 + i];
#pragma unroll
  for (int i = 0; i < 5; i++) cam[i] = __ldg(&P.cam[ip * 5 + i]);
  quat_to_R(pose + 3, R);
  transform(R, pose, X, Xc);
  const bool stereo = !(m.z < 0.0f);
  reproj_error(Xc, m.x, m.y, m.z, cam, stereo, L.e);
  L.depth_pos = Xc[2] > 0.0;
  const double info = (double)m.w;
  const double c = L.e[0] * (info * L.e[0]) + L.e[1] * (info * L.e[1]) + L.e[2] * (info * L.e[2]);
  double rho1 = 1.0;
  L.rho0 = c;
  if (robust) {
    const double delta = stereo ? d3 : d2;
    huber(c, delta, huber_dsqr(delta), &L.rho0, &rho1);
  }
  L.w = sqrt(rho1 * info);
  L.stereo = stereo;
  if (want_jac) {
    reproj_jacobian_point(R, Xc, cam, stereo, L.Jl);
    const double iz = 1.0 / Xc[2];
    L.g[0] = Xc[0] * iz; L.g[1] = Xc[1] * iz; L.g[2] = iz; L.g[3] = L.w;
    L.cam3[0] = cam[0]; L.cam3[1] = cam[1]; L.cam3[2] = cam[4];
  }
}

// per-observation words of one lane, loadable one tile ahead of their use (8 registers)
struct LinOps {
  float4 m;
  int ip, lm;
  unsigned lp;
  int live;
};
__device__ __forceinline__ LinOps lin_load_ops(const Dev& P, int o) {
  LinOps q;
  q.m = __ldg(&P.obs_meas[o]);
  q.ip = __ldg(&P.obs_pose[o]);
  q.lm = __ldg(&P.obs_point[o]);
  q.lp = __ldg(&P.obs_lp[o]);
  q.live = P.obs_level[o];  // raw level byte; compared where it is used, so this pre-load never waits for the data
  return q;
}

// long landmark of the linearisation (one warp, more than 32 observations): chunks of 32 observations, whole-warp
// reductions, direct atomics on the pose side; adds to chi_acc, returns the landmark's max diagonal in maxd
__device__ __forceinline__ void linearize_long_item(const Dev& P, const TileInfo& ti, int lane, int start, int cnt, int robust,
                                                    double d2, double d3, double& chi_acc, double& maxd) {
  const size_t No = (size_t)P.ld; const int Nl = P.n_point;
  double* __restrict__ jq = P.JQ + ti.jq_off;
    // long landmark: chunks of 32 observations, whole-warp reductions, direct atomics on the pose side
    double bl_acc[3] = {0, 0, 0}, hl_acc[3] = {0, 0, 0};
    for (int base = 0; base < cnt; base += 32) {
      const int i = base + lane;
      const bool act = i < cnt;
      const int o = start + (act ? i : 0);
      ObsLin L;
#pragma unroll
      for (int c = 0; c < 9; c++) L.Jl[c] = 0.0;
      L.e[0] = L.e[1] = L.e[2] = 0.0;
      if (act) {
        const int slot = P.obs_slot[o];
        const bool live = P.obs_level[o] == 0;
        obs_eval(P, o, true, robust != 0, d2, d3, L);
        if (live) {
#pragma unroll
          for (int c = 0; c < 3; c++) P.err[(size_t)c * No + o] = L.e[c];
          chi_acc += L.rho0;
        }
#pragma unroll
        for (int c = 0; c < JG; c++) { L.g[c] = live ? L.g[c] : 0.0; jq[(size_t)c * ti.nt + (o - ti.o0)] = L.g[c]; }
#pragma unroll
        for (int c = 0; c < 9; c++) { L.Jl[c] = live ? L.Jl[c] * L.w : 0.0; P.Jl[(size_t)c * No + o] = L.Jl[c]; }
#pragma unroll
        for (int c = 0; c < 3; c++) { L.e[c] = live ? L.e[c] * L.w : 0.0; P.r[(size_t)c * No + o] = L.e[c]; }
        if (slot >= 0 && live) {
          double Jp[18];
          jp_full(jp_compact(L.g, L.cam3[0], L.cam3[1], L.cam3[2], L.stereo), L.stereo, Jp);
#pragma unroll
          for (int c = 0; c < 6; c++) {
            atomicAdd(&P.bp[slot * 6 + c], -(Jp[c] * L.e[0] + Jp[6 + c] * L.e[1] + Jp[12 + c] * L.e[2]));
            atomicAdd(&P.hd[slot * 6 + c], Jp[c] * Jp[c] + Jp[6 + c] * Jp[6 + c] + Jp[12 + c] * Jp[12 + c]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        bl_acc[c] += warp_sum(L.Jl[c] * L.e[0] + L.Jl[3 + c] * L.e[1] + L.Jl[6 + c] * L.e[2]);
        hl_acc[c] += warp_sum(L.Jl[c] * L.Jl[c] + L.Jl[3 + c] * L.Jl[3 + c] + L.Jl[6 + c] * L.Jl[6 + c]);
      }
    }
    const int lm = P.obs_point[start];
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 3; c++) P.bl[(size_t)c * Nl + lm] = -bl_acc[c];
    }
    maxd = fmax(hl_acc[0], fmax(hl_acc[1], hl_acc[2]));
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// One tile of the linearisation.  PRE: the lanes of short items already hold their LinOps (pipelined kernel).
template <bool PRE>
__device__ __forceinline__ void linearize_tile(const Dev& P, const TileInfo& ti, const LinOps& pre, int robust, double d2,
                                               double d3, double* c_sh, const int* runs_staged = nullptr,
                                               int pf_pose = -1, int pf_point = -1, double* acc_det = nullptr) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int w = ti.item0 + wid;
  const bool valid = wid < ti.nitem;
  const int win = ti.win;
  const bool on = valid;
  const int start = tile_item_start(ti, wid), cnt = tile_item_cnt(ti, wid);
  const size_t No = (size_t)P.ld; const int Nl = P.n_point;
  const bool is_long = cnt > 32;
  double* __restrict__ jq = P.JQ + ti.jq_off;
  const int sbase = P.smallwin ? P.win_slot_ptr[win] : 0;  // for the pose-side reduction at the end: requested now
  double chi_acc = 0.0, maxd = 0.0;
  double vb[6] = {0, 0, 0, 0, 0, 0}, vh[6] = {0, 0, 0, 0, 0, 0};  // this observation's -Jp^T r and diag(Jp^T Jp)
  bool has = false;
  int rank = 0;
  if (valid && !is_long) {
    const bool act = lane < cnt;
    const int o = start + (act ? lane : 0);
    int lm = -1 - lane;
    ObsLin L;
#pragma unroll
    for (int c = 0; c < 9; c++) L.Jl[c] = 0.0;
    L.e[0] = L.e[1] = L.e[2] = 0.0;
    LinOps q = pre;
    if (act) {
      if (!PRE) q = lin_load_ops(P, o);
      has = (q.lp & 0xffffu) != 0xffffu;
      rank = (int)((q.lp >> 16) & 0x7fffu);
      lm = q.lm;
    }
    const int fcol = tile_fcol(ti, wid, has, lane);
    if (act && on) {
      const bool live = q.live == 0;
      obs_eval_ops(P, q.m, q.ip, q.lm, true, robust != 0, d2, d3, L);
      if (live) {
#pragma unroll
        for (int c = 0; c < 3; c++) P.err[(size_t)c * No + o] = L.e[c];
        chi_acc += L.rho0;
      }
      // an excluded (level-1) edge contributes zero rows; select, do not multiply (its Jacobian may be inf/NaN)
#pragma unroll
      for (int c = 0; c < JG; c++) { L.g[c] = live ? L.g[c] : 0.0; if (has) jq[(size_t)c * ti.nt + fcol] = L.g[c]; }
#pragma unroll
      for (int c = 0; c < 9; c++) { L.Jl[c] = live ? L.Jl[c] * L.w : 0.0; P.Jl[(size_t)c * No + o] = L.Jl[c]; }
#pragma unroll
      for (int c = 0; c < 3; c++) { L.e[c] = live ? L.e[c] * L.w : 0.0; P.r[(size_t)c * No + o] = L.e[c]; }
      if (has) {
        double Jp[18];
        jp_full(jp_compact(L.g, L.cam3[0], L.cam3[1], L.cam3[2], L.stereo), L.stereo, Jp);
#pragma unroll
        for (int c = 0; c < 6; c++) {
          vb[c] = -(Jp[c] * L.e[0] + Jp[6 + c] * L.e[1] + Jp[12 + c] * L.e[2]);
          vh[c] = Jp[c] * Jp[c] + Jp[6 + c] * Jp[6 + c] + Jp[12 + c] * Jp[12 + c];
        }
      }
    }
    if (on) {  // warp-uniform: landmark-side gradient and Hessian diagonal by segmented shuffles
      const Seg sg = seg_of(lm, lane);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double g = L.Jl[c] * L.e[0] + L.Jl[3 + c] * L.e[1] + L.Jl[6 + c] * L.e[2];
        double h = L.Jl[c] * L.Jl[c] + L.Jl[3 + c] * L.Jl[3 + c] + L.Jl[6 + c] * L.Jl[6 + c];
        g = seg_sum(g, sg, lane);
        h = seg_sum(h, sg, lane);
        if (act && lane == sg.start) P.bl[(size_t)c * Nl + lm] = -g;
        maxd = fmax(maxd, h);
      }
    }
  } else if (on) {
    linearize_long_item(P, ti, lane, start, cnt, robust, d2, d3, chi_acc, maxd);
  }
  if (on) {
    chi_acc = warp_sum(chi_acc);
    maxd = warp_max(maxd);
    if (lane == 0) {
      P.chi_part[w] = chi_acc;
      atomic_max_pos(&P.ctl[win].maxdiag_bits, maxd);
    }
  }
  if (has) {
#pragma unroll
    for (int c = 0; c < 6; c++) { c_sh[c * SCST + rank] = vb[c]; c_sh[(6 + c) * SCST + rank] = vh[c]; }
  }
  if (pf_pose >= 0) {  // the arithmetic of this tile is done and the next tile's indices arrived long ago: pull its
                       // pose (56 B: two lines at most), camera and landmark lines into L1 while the reduction runs
    prefetch_l1(P.pose + (size_t)pf_pose * 7);
    prefetch_l1(P.pose + (size_t)pf_pose * 7 + 6);
    prefetch_l1(P.cam + (size_t)pf_pose * 5);
    prefetch_l1(P.point + (size_t)pf_point * 3);
  }
  __syncthreads();
  const int* runs = runs_staged ? runs_staged : reinterpret_cast<const int*>(jq + (size_t)JQ_ROWS * ti.nt);
  double* bp = P.bp;
  double* hd = P.hd;
  if (acc_det) { tile_scatter_acc<12>(runs, runs + ti.nrun + 1, ti.nrun, c_sh, acc_det); return; }
  tile_scatter_all<12>(runs, runs + ti.nrun + 1, ti.nrun, sbase, c_sh,
                       [bp, hd](int slot, int k) { return (k < 6) ? bp + (size_t)slot * 6 + k : hd + (size_t)slot * 6 + (k - 6); });
}

__global__ void __launch_bounds__(CTA, 4) k_linearize(Dev P, int robust, double d2, double d3, int force_all) {
  __shared__ double c_sh[12 * SCST];
  const TileInfo ti = P.tiles[blockIdx.x];
  if (!force_all && lin_phase(P, P.ctl[ti.win]) != PH_LIN) return;  // a tile lies inside one window: CTA-uniform
  if (robust < 0) robust = P.ctl[ti.win].robust;
  LinOps none{};
  linearize_tile<false>(P, ti, none, robust, d2, d3, c_sh);
}

// Pipelined variant (same idea as k_qr_pipe): every CTA walks LIN_TPB consecutive tiles; the tile descriptor two tiles
// ahead is staged in shared memory with cp.async and every lane loads its own observation words (measurement, pose and
// landmark index, meta, level: 8 registers) for the NEXT tile while it linearises the current one, so only the
// L2-resident pose / point gathers stay on the critical path.
constexpr int LIN_TPB = 8;
constexpr int LIN_RUN_INTS = 2 * CTA + 2;
// STAGE: the run table of the next tile is copied to shared memory with cp.async together with the descriptor (it sits
// at the cold tail of the tile's JQ block, which nothing else touches before the pose-side reduction needs it), and the
// landmark coordinates of the following tiles -- consecutive in memory, landmarks are tiled in order -- are pulled into
// L2 by a TMA prefetch, so the only exposed global latency left is the L1/L2-resident pose gather.
// PF (experimental, off by default; A/B knob reserved[7] = 6): when the arithmetic of a tile is done, the indices of
// the next tile's observations have long arrived in registers -- linearize_tile prefetches their pose, camera and
// landmark lines into L1 before the pose-side reduction, so the gathers at the top of the next tile hit L1, not L2.
template <bool STAGE, bool PF = false>
__global__ void __launch_bounds__(CTA, 4) k_linearize_pipe(Dev P, int robust, double d2, double d3, int force_all) {
  __shared__ double c_sh[12 * SCST];
  __shared__ __align__(16) TileInfo ti_sh[3];
  __shared__ int run_sh[STAGE ? 2 : 1][STAGE ? LIN_RUN_INTS : 1];
  extern __shared__ double lin_acc[];  // reproducible mode: 12 * maxslot accumulators of the window being walked
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t0 = blockIdx.x * LIN_TPB, t1 = min(t0 + LIN_TPB, P.n_tile);
  int det_win = -1;
  if (P.det) {
    for (int i = tid; i < 12 * P.maxslot; i += CTA) lin_acc[i] = 0.0;
    __syncthreads();
  }
  auto stage_runs = [&](const TileInfo& ti, int rb) {
    if (!STAGE || ti.is_long) return;
    const int* runs = reinterpret_cast<const int*>(P.JQ + ti.jq_off + (size_t)JQ_ROWS * ti.nt);
    for (int i = tid; i < 2 * ti.nrun + 1; i += CTA) cp_async4(&run_sh[rb][i], runs + i);
  };
  auto ops_of = [&](const TileInfo& ti) {
    LinOps q{};
    if (!ti.is_long) {
      const int cnt = tile_item_cnt(ti, wid), start = tile_item_start(ti, wid);
      if (wid < ti.nitem && lane < cnt) q = lin_load_ops(P, start + lane);
    }
    return q;
  };
  if (tid < 5) {
    cp_async16(reinterpret_cast<char*>(&ti_sh[0]) + tid * 16, reinterpret_cast<const char*>(P.tiles + t0) + tid * 16);
    if (t0 + 1 < t1)
      cp_async16(reinterpret_cast<char*>(&ti_sh[1]) + tid * 16, reinterpret_cast<const char*>(P.tiles + t0 + 1) + tid * 16);
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  LinOps nxt = ops_of(ti_sh[0]);
  stage_runs(ti_sh[0], 0);
  cp_async_commit();
  int nphase = force_all ? PH_LIN : lin_phase(P, P.ctl[ti_sh[0].win]);
  int nrob = robust < 0 ? P.ctl[ti_sh[0].win].robust : robust;  // robust < 0: the pass's setting, from the control block
  for (int k = 0; k < t1 - t0; k++) {
    cp_async_wait_all();  // descriptor of tile k+1 (and the run table of tile k)
    __syncthreads();      // ... visible to everybody; c_sh of the previous tile is free
    const TileInfo ti = ti_sh[k % 3];
    const LinOps cur = nxt;
    const int phase = nphase;
    const int rob = nrob;
    if (k + 1 < t1 - t0) {
      const TileInfo& tn = ti_sh[(k + 1) % 3];
      if (k + 2 < t1 - t0 && tid < 5)
        cp_async16(reinterpret_cast<char*>(&ti_sh[(k + 2) % 3]) + tid * 16,
                   reinterpret_cast<const char*>(P.tiles + t0 + k + 2) + tid * 16);
      nxt = ops_of(tn);
      stage_runs(tn, (k + 1) & 1);
      if (!force_all) nphase = lin_phase(P, P.ctl[tn.win]);
      if (robust < 0) nrob = P.ctl[tn.win].robust;
    }
    cp_async_commit();
    if (phase != PH_LIN) continue;  // CTA-uniform
    if (P.det && ti.win != det_win) {  // a new window starts: the finished one leaves the CTA as one partial vector
      if (det_win >= 0)
        det_flush<12>(P.part_lin, blockIdx.x + det_win, P.maxslot, P.win_slot_ptr[det_win + 1] - P.win_slot_ptr[det_win], lin_acc);
      det_win = ti.win;
    }
    if (STAGE && !ti.is_long) {
      // the last lane of the tile's last item knows the tile's last landmark: pull the coordinates of the next ~96
      // landmarks (the next two or three tiles of this CTA) towards L2
      const int li = ti.nitem - 1;
      if (wid == li && lane == tile_item_cnt(ti, li) - 1) {
        long long first = ((long long)cur.lm + 1) & ~1ll;  // 16-byte aligned start
        long long bytes = ((long long)P.n_point - first) * 24;
        bytes = (bytes < 2304 ? bytes : 2304) & ~15ll;
        if (bytes > 0) bulk_prefetch_l2(P.point + first * 3, (uint32_t)bytes);
      }
    }
    bool pf = false;  // lanes that own an observation of the next tile prefetch its operands (PF only)
    if (PF && k + 1 < t1 - t0) {
      const TileInfo& tn = ti_sh[(k + 1) % 3];
      pf = !tn.is_long && wid < tn.nitem && lane < tile_item_cnt(tn, wid);
    }
    linearize_tile<true>(P, ti, cur, rob, d2, d3, c_sh, (STAGE && !ti.is_long) ? run_sh[k & 1] : nullptr,
                         pf ? nxt.ip : -1, pf ? nxt.lm : -1, P.det ? lin_acc : nullptr);
  }
  cp_async_wait_all();
  if (P.det && det_win >= 0)
    det_flush<12>(P.part_lin, blockIdx.x + det_win, P.maxslot, P.win_slot_ptr[det_win + 1] - P.win_slot_ptr[det_win], lin_acc);
}

// ------------------------------------------------------------------------------------------------ K8a: LM begin
// OptimizationAlgorithmLevenberg::solve up to the trial loop (optimization_algorithm_levenberg.cpp:75-100):
// currentChi, iniChi, lambda init = tau * max diag(H) at iteration 0 (computeLambdaInit :166-180).
// Split in two so that, with landmarks sharded over several GPUs, the per-window partial sums (wred) and the pose-side
// vectors can be all-reduced between the two kernels; with one GPU they simply run back to back.
__device__ __forceinline__ void lm_reduce_lin_body(const Dev& P, int win, double* sh) {
  WinCtl& c = P.ctl[win];
  if (c.phase != PH_LIN) {  // nothing new from this window: contribute the neutral element
    if (threadIdx.x == 0) { P.wred[win] = 0.0; P.wred[2 * P.n_win + win] = 0.0; }
    return;
  }
  if (P.det && P.win_tile_ptr[win + 1] > P.win_tile_ptr[win]) {  // b_p and diag(Jp^T Jp): the CTAs' partial vectors in order
    const int c_lo = P.win_tile_ptr[win] / LIN_TPB, c_hi = (P.win_tile_ptr[win + 1] - 1) / LIN_TPB;
    const int s0 = P.win_slot_ptr[win], wn = P.win_slot_ptr[win + 1] - s0;
    for (int i = threadIdx.x; i < 12 * wn; i += RCTA) {
      const double v = det_sum(P.part_lin, c_lo, c_hi, win, (size_t)12 * P.maxslot, i);
      const int sl = i / 12, k = i - sl * 12;
      if (k < 6) P.bp[(size_t)(s0 + sl) * 6 + k] = v; else P.hd[(size_t)(s0 + sl) * 6 + (k - 6)] = v;
    }
  }
  double chi = 0.0;
  for (int i = P.win_item_ptr[win] + threadIdx.x; i < P.win_item_ptr[win + 1]; i += RCTA) chi += P.chi_part[i];
  chi = block_sum(chi, sh);
  if (threadIdx.x == 0) {
    P.wred[win] = chi;
    P.wred[2 * P.n_win + win] = __longlong_as_double((long long)c.maxdiag_bits);
    c.maxdiag_bits = 0ull;
  }
}
__global__ void __launch_bounds__(RCTA) k_lm_reduce_lin(Dev P) {
  __shared__ double sh[RWARPS];
  lm_reduce_lin_body(P, blockIdx.x, sh);
}

__device__ __forceinline__ void lm_begin_body(const Dev& P, int win, double* sh) {
  WinCtl& c = P.ctl[win];
  if (c.phase != PH_LIN) return;
  double md = 0.0;
  for (int i = P.win_slot_ptr[win] * 6 + threadIdx.x; i < P.win_slot_ptr[win + 1] * 6; i += RCTA)
    md = fmax(md, fabs(P.hd[i]));
  md = warp_max(md);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = md;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < RWARPS; i++) md = fmax(md, sh[i]);
    md = fmax(md, P.wred[2 * P.n_win + win]);
    const double chi = P.wred[win];
    c.cur_chi = chi;
    c.ini_chi = chi;
    c.tmp_chi = chi;
    if (c.iter == 0) {
      c.lambda = 1e-5 * md;  // _tau
      c.ni = 2.0;
      c.nbad = 0;
    }
    c.qmax = 0;
    c.rho = 0.0;
    c.phase = PH_TRIAL;
    c.lin_count++;
  }
}
__global__ void __launch_bounds__(RCTA) k_lm_begin(Dev P) {
  __shared__ double sh[RWARPS];
  lm_begin_body(P, blockIdx.x, sh);
}

// Single-rank fusion of the three small kernels between the linearisation and the landmark QR (one launch instead of
// three; used by the CUDA-graph macro step of single-window problems): per-window chi2 / max-diagonal reduction, the LM
// "begin" step, and the zeroing of this trial's pose-side accumulators (reduced rhs, block-Jacobi blocks).
__global__ void __launch_bounds__(RCTA) k_trial_begin(Dev P) {
  __shared__ double sh[RWARPS];
  const int win = blockIdx.x;
  lm_reduce_lin_body(P, win, sh);
  __syncthreads();
  lm_begin_body(P, win, sh);
  __syncthreads();  // thread 0's phase write is visible to the CTA
  if (qr_phase(P, P.ctl[win]) != PH_TRIAL) return;  // (the fused kernel's windows were zeroed before it ran)
  const int s0 = P.win_slot_ptr[win], s1 = P.win_slot_ptr[win + 1];
  for (int i = s0 * 6 + threadIdx.x; i < s1 * 6; i += RCTA) P.bs[i] = 0.0;
  for (int i = s0 * 21 + threadIdx.x; i < s1 * 21; i += RCTA) P.D[i] = 0.0;
}

// ------------------------------------------------------------------------------------------------ K2: landmark QR
// In-place Householder QR of A_l = [sqrt(lambda) I3 ; J_l] (damping rows first, so reflector j touches damping
// row j and every observation row), one lane per observation, norms/dots by (segmented) warp shuffles.
// Replaces the per-landmark (Hll+lambda I)^-1 and Schur products of block_solver.hpp:381-432 without forming
// them: keeps R (3x3), the observation rows of the thin factor Q1 (compact-WY: Q1_obs = -V T diag(v0)) and
// t_l = Q1^T r.  Also emits the reduced right-hand side b_s = -sum Jp^T (r - Q1 t_l) and the 6x6 block-Jacobi
// blocks sum Jp^T Jp - (Jp^T Q1)(Jp^T Q1)^T of the (never formed) reduced camera matrix.
struct LmFactor {
  double beta[3], v0[3], w01, w02, w12;
  double Rm[6];
  double M[6];  // upper-tri of T*diag(v0): m00 m01 m02 m11 m12 m22
};

__device__ __forceinline__ void wy_from_gram(LmFactor& F, double g01, double g02, double g12) {
  const double T00 = F.beta[0], T11 = F.beta[1], T22 = F.beta[2];
  const double T01 = -F.beta[1] * (T00 * g01);
  const double T02 = -F.beta[2] * (T00 * g02 + T01 * g12);
  const double T12 = -F.beta[2] * (T11 * g12);
  F.M[0] = T00 * F.v0[0]; F.M[1] = T01 * F.v0[1]; F.M[2] = T02 * F.v0[2];
  F.M[3] = T11 * F.v0[1]; F.M[4] = T12 * F.v0[2];
  F.M[5] = T22 * F.v0[2];
}

// rows of V for one observation from its rows of J_l (a, 3x3 row-major): V0=a0, V1=a1-w01 a0, V2=a2-w02 a0-w12 V1
__device__ __forceinline__ void v_rows(const double a[9], const LmFactor& F, double V[9]) {
#pragma unroll
  for (int r = 0; r < 3; r++) {
    V[r * 3 + 0] = a[r * 3 + 0];
    V[r * 3 + 1] = a[r * 3 + 1] - F.w01 * V[r * 3 + 0];
    V[r * 3 + 2] = a[r * 3 + 2] - F.w02 * V[r * 3 + 0] - F.w12 * V[r * 3 + 1];
  }
}
__device__ __forceinline__ void q1_rows(const double V[9], const LmFactor& F, double Q[9]) {
#pragma unroll
  for (int r = 0; r < 3; r++) {
    Q[r * 3 + 0] = -(V[r * 3 + 0] * F.M[0]);
    Q[r * 3 + 1] = -(V[r * 3 + 0] * F.M[1] + V[r * 3 + 1] * F.M[3]);
    Q[r * 3 + 2] = -(V[r * 3 + 0] * F.M[2] + V[r * 3 + 1] * F.M[4] + V[r * 3 + 2] * F.M[5]);
  }
}

// one Householder column step given the reduced sums; updates the factor
__device__ __forceinline__ void hh_col(LmFactor& F, int j, double lam, double sl, double sig) {
  const double norm = sqrt(lam + sig);
  F.v0[j] = sl + norm;
  F.beta[j] = 1.0 / (norm * F.v0[j]);
  F.Rm[j == 0 ? 0 : (j == 1 ? 3 : 5)] = -norm;
}

__device__ __forceinline__ void load9(const double* planes, size_t stride, size_t o, double a[9]) {
#pragma unroll
  for (int c = 0; c < 9; c++) a[c] = planes[(size_t)c * stride + o];
}

// this observation's weighted 3x6 pose block, rebuilt from the geometry rows of its JQ column (`slot` = global free slot)
__device__ __forceinline__ void load_Jp(const Dev& P, const double* __restrict__ jq, int nt, int col, int slot, bool stereo,
                                        double J[18]) {
  double g[JG];
#pragma unroll
  for (int c = 0; c < JG; c++) g[c] = __ldg(jq + (size_t)c * nt + col);
  const double* cam = P.slot_cam + (size_t)slot * 3;
  jp_full(jp_compact(g, __ldg(cam), __ldg(cam + 1), __ldg(cam + 2), stereo), stereo, J);
}
// global free slot of an observation from its meta word (window-relative slot in the low 16 bits when smallwin)
__device__ __forceinline__ int slot_of(const Dev& P, unsigned lp, int sbase, int o) {
  return P.smallwin ? sbase + (int)(lp & 0xffffu) : P.obs_slot[o];
}

// one observation's pose-side contributions for the trial: reduced rhs (6) and block-Jacobi block (21, upper tri)
__device__ __forceinline__ void trial_contrib(const double J[18], const double Q[9],
                                              const double rr[3], const double tl[3], double out[27]) {
  double u[3];
#pragma unroll
  for (int r = 0; r < 3; r++) u[r] = rr[r] - (Q[r * 3] * tl[0] + Q[r * 3 + 1] * tl[1] + Q[r * 3 + 2] * tl[2]);
#pragma unroll
  for (int c = 0; c < 6; c++) out[c] = -(J[c] * u[0] + J[6 + c] * u[1] + J[12 + c] * u[2]);
  double G[18];  // Jp^T Q1 (6x3)
#pragma unroll
  for (int c = 0; c < 6; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) G[c * 3 + k] = J[c] * Q[k] + J[6 + c] * Q[3 + k] + J[12 + c] * Q[6 + k];
  int idx = 6;
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = a; b < 6; b++) {
      out[idx] = J[a] * J[b] + J[6 + a] * J[6 + b] + J[12 + a] * J[12 + b] -
                 (G[a * 3] * G[b * 3] + G[a * 3 + 1] * G[b * 3 + 1] + G[a * 3 + 2] * G[b * 3 + 2]);
      idx++;
    }
}

// short-item variant of trial_contrib: the 27 values go straight into column `rank` of c_sh[27][SCST]
__device__ __forceinline__ void trial_contrib_regs(const double J[18], const double Q[9], const double rr[3],
                                                   const double tl[3], double* c_sh, int rank) {
  {
    double u[3];
#pragma unroll
    for (int r = 0; r < 3; r++) u[r] = rr[r] - (Q[r * 3] * tl[0] + Q[r * 3 + 1] * tl[1] + Q[r * 3 + 2] * tl[2]);
#pragma unroll
    for (int c = 0; c < 6; c++) c_sh[c * SCST + rank] = -(J[c] * u[0] + J[6 + c] * u[1] + J[12 + c] * u[2]);
  }
  double G[18];  // Jp^T Q1 (6x3)
#pragma unroll
  for (int c = 0; c < 6; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) G[c * 3 + k] = J[c] * Q[k] + J[6 + c] * Q[3 + k] + J[12 + c] * Q[6 + k];
  int idx = 6;
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = a; b < 6; b++) {
      c_sh[idx * SCST + rank] = J[a] * J[b] + J[6 + a] * J[6 + b] + J[12 + a] * J[12 + b] -
                                (G[a * 3] * G[b * 3] + G[a * 3 + 1] * G[b * 3 + 1] + G[a * 3 + 2] * G[b * 3 + 2]);
      idx++;
    }
}

// ---- the two item shapes of the landmark QR, shared by the plain and the pipelined kernel
// short item: operands of this lane's observation are already in registers (a = its rows of J_l, rr = its residual)
__device__ __forceinline__ void qr_short_item(const Dev& P, const TileInfo& ti, int wid, int lane, bool act, int o, int lm,
                                              bool has, int rank, int slot, bool stereo, const double a[9],
                                              const double rr[3], double lam, double* c_sh) {
  const double sl = sqrt(lam);
  const int Nl = P.n_point, nt = ti.nt;
  const bool on = true;
  double* __restrict__ jq = P.JQ + ti.jq_off;
  LmFactor F;
    const int fcol = tile_fcol(ti, wid, has, lane);
    if (on) {
      const Seg sg = seg_of(lm, lane);
      // column 0
      double s0 = seg_sum(a[0] * a[0] + a[3] * a[3] + a[6] * a[6], sg, lane);
      double d01 = seg_sum(a[0] * a[1] + a[3] * a[4] + a[6] * a[7], sg, lane);
      double d02 = seg_sum(a[0] * a[2] + a[3] * a[5] + a[6] * a[8], sg, lane);
      hh_col(F, 0, lam, sl, s0);
      F.w01 = F.beta[0] * d01;
      F.w02 = F.beta[0] * d02;
      F.Rm[1] = -F.w01 * F.v0[0];
      F.Rm[2] = -F.w02 * F.v0[0];
      double V[9];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        V[r * 3] = a[r * 3];
        V[r * 3 + 1] = a[r * 3 + 1] - F.w01 * a[r * 3];
        V[r * 3 + 2] = a[r * 3 + 2] - F.w02 * a[r * 3];
      }
      // column 1
      double s1 = seg_sum(V[1] * V[1] + V[4] * V[4] + V[7] * V[7], sg, lane);
      double d12 = seg_sum(V[1] * V[2] + V[4] * V[5] + V[7] * V[8], sg, lane);
      hh_col(F, 1, lam, sl, s1);
      F.w12 = F.beta[1] * d12;
      F.Rm[4] = -F.w12 * F.v0[1];
#pragma unroll
      for (int r = 0; r < 3; r++) V[r * 3 + 2] -= F.w12 * V[r * 3 + 1];
      // column 2
      double s2 = seg_sum(V[2] * V[2] + V[5] * V[5] + V[8] * V[8], sg, lane);
      hh_col(F, 2, lam, sl, s2);
      // Gram of the reflector observation parts -> compact WY
      double g01 = seg_sum(V[0] * V[1] + V[3] * V[4] + V[6] * V[7], sg, lane);
      double g02 = seg_sum(V[0] * V[2] + V[3] * V[5] + V[6] * V[8], sg, lane);
      double g12 = seg_sum(V[1] * V[2] + V[4] * V[5] + V[7] * V[8], sg, lane);
      wy_from_gram(F, g01, g02, g12);
      double Q[9];
      q1_rows(V, F, Q);
      double tl[3];
#pragma unroll
      for (int k = 0; k < 3; k++) tl[k] = seg_sum(Q[k] * rr[0] + Q[3 + k] * rr[1] + Q[6 + k] * rr[2], sg, lane);
      if (act) {
        if (has) {
#pragma unroll
          for (int c = 0; c < 9; c++) jq[(size_t)(JG + c) * nt + fcol] = Q[c];
        }
        if (lane == sg.start) {
#pragma unroll
          for (int c = 0; c < 6; c++) P.R[(size_t)c * Nl + lm] = F.Rm[c];
#pragma unroll
          for (int c = 0; c < 3; c++) P.tl[(size_t)c * Nl + lm] = tl[c];
        }
        if (has) {
          double J[18];
          load_Jp(P, jq, nt, fcol, slot, stereo, J);
          trial_contrib_regs(J, Q, rr, tl, c_sh, rank);
        }
      }
    }
}

// Second version of the short item, used by the pipelined kernel.  Same factorisation, shorter critical path:
//  * nine segmented reductions in three dependent groups instead of thirteen in five.  The Gram entries of the
//    reflectors follow from sums that are already there -- with norm_j = sqrt(lambda + sigma_j) the column-j reflector
//    leaves  sum V_j . V_k = d_jk * sqrt(lambda) / norm_j  (k > j, before later updates) -- and t_l = Q1^T r is
//    -M^T (V^T r), where V^T r comes from c_j = sum a_j . r (reduced together with the first group) by the same
//    column updates as V itself;
//  * the 18 Jp rows of the observation are requested right after the first group (read-only path: the kernel only
//    writes rows 18-26 of the block), so their L2 latency is covered by the remaining two reduction groups.
template <int JPOS>  // where the Jp rows are requested: 0 = right before they are used, 1 = after reduction group 1, 2 = after group 2
__device__ __forceinline__ void qr_short_item_v2(const Dev& P, const TileInfo& ti, int wid, int lane, bool act, int lm,
                                                 bool has, int rank, int slot, bool stereo, const double a[9],
                                                 const double rr[3], double lam, double* c_sh) {
  const double sl = sqrt(lam);
  const int Nl = P.n_point, nt = ti.nt;
  double* __restrict__ jq = P.JQ + ti.jq_off;
  LmFactor F;
  const int fcol = tile_fcol(ti, wid, has, lane);
  const Seg sg = seg_of(lm, lane);
  // group 1: column 0 and the three products with the residual
  const double s0 = seg_sum(a[0] * a[0] + a[3] * a[3] + a[6] * a[6], sg, lane);
  const double d01 = seg_sum(a[0] * a[1] + a[3] * a[4] + a[6] * a[7], sg, lane);
  const double d02 = seg_sum(a[0] * a[2] + a[3] * a[5] + a[6] * a[8], sg, lane);
  const double c0 = seg_sum(a[0] * rr[0] + a[3] * rr[1] + a[6] * rr[2], sg, lane);
  const double c1 = seg_sum(a[1] * rr[0] + a[4] * rr[1] + a[7] * rr[2], sg, lane);
  const double c2 = seg_sum(a[2] * rr[0] + a[5] * rr[1] + a[8] * rr[2], sg, lane);
  double J[18];
  auto load_J = [&]() {
    if (act && has) {
      load_Jp(P, jq, nt, fcol, slot, stereo, J);
    } else {
#pragma unroll
      for (int c = 0; c < 18; c++) J[c] = 0.0;
    }
  };
  if (JPOS == 1) load_J();
  hh_col(F, 0, lam, sl, s0);
  const double norm0 = -F.Rm[0];
  F.w01 = F.beta[0] * d01;
  F.w02 = F.beta[0] * d02;
  F.Rm[1] = -F.w01 * F.v0[0];
  F.Rm[2] = -F.w02 * F.v0[0];
  double V[9];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    V[r * 3] = a[r * 3];
    V[r * 3 + 1] = a[r * 3 + 1] - F.w01 * a[r * 3];
    V[r * 3 + 2] = a[r * 3 + 2] - F.w02 * a[r * 3];
  }
  // group 2: column 1
  const double s1 = seg_sum(V[1] * V[1] + V[4] * V[4] + V[7] * V[7], sg, lane);
  const double d12 = seg_sum(V[1] * V[2] + V[4] * V[5] + V[7] * V[8], sg, lane);
  if (JPOS == 2) load_J();
  hh_col(F, 1, lam, sl, s1);
  const double norm1 = -F.Rm[3];
  F.w12 = F.beta[1] * d12;
  F.Rm[4] = -F.w12 * F.v0[1];
#pragma unroll
  for (int r = 0; r < 3; r++) V[r * 3 + 2] -= F.w12 * V[r * 3 + 1];
  // group 3: column 2
  const double s2 = seg_sum(V[2] * V[2] + V[5] * V[5] + V[8] * V[8], sg, lane);
  hh_col(F, 2, lam, sl, s2);
  const double g01 = d01 * (sl / norm0);
  const double g02 = d02 * (sl / norm0) - F.w12 * g01;
  const double g12 = d12 * (sl / norm1);
  wy_from_gram(F, g01, g02, g12);
  double Q[9];
  q1_rows(V, F, Q);
  const double u0 = c0, u1 = c1 - F.w01 * c0, u2 = c2 - F.w02 * c0 - F.w12 * u1;
  double tl[3];
  tl[0] = -(F.M[0] * u0);
  tl[1] = -(F.M[1] * u0 + F.M[3] * u1);
  tl[2] = -(F.M[2] * u0 + F.M[4] * u1 + F.M[5] * u2);
  if (JPOS == 0) load_J();
  if (act) {
    if (has) {
#pragma unroll
      for (int c = 0; c < 9; c++) jq[(size_t)(JG + c) * nt + fcol] = Q[c];
    }
    if (lane == sg.start) {
#pragma unroll
      for (int c = 0; c < 6; c++) P.R[(size_t)c * Nl + lm] = F.Rm[c];
#pragma unroll
      for (int c = 0; c < 3; c++) P.tl[(size_t)c * Nl + lm] = tl[c];
    }
    if (has) trial_contrib_regs(J, Q, rr, tl, c_sh, rank);
  }
}

// The factorisation itself (nine segmented reductions in three dependent groups, see qr_short_item_v2), operands and
// results in registers: R (F.Rm), this observation's rows of Q1 and t_l = Q1^T r.  Must be called by the whole warp.
__device__ __forceinline__ void qr_factor_v2(const double a[9], const double rr[3], double lam, const Seg& sg, int lane,
                                             LmFactor& F, double Q[9], double tl[3]) {
  const double sl = sqrt(lam);
  const double s0 = seg_sum(a[0] * a[0] + a[3] * a[3] + a[6] * a[6], sg, lane);
  const double d01 = seg_sum(a[0] * a[1] + a[3] * a[4] + a[6] * a[7], sg, lane);
  const double d02 = seg_sum(a[0] * a[2] + a[3] * a[5] + a[6] * a[8], sg, lane);
  const double c0 = seg_sum(a[0] * rr[0] + a[3] * rr[1] + a[6] * rr[2], sg, lane);
  const double c1 = seg_sum(a[1] * rr[0] + a[4] * rr[1] + a[7] * rr[2], sg, lane);
  const double c2 = seg_sum(a[2] * rr[0] + a[5] * rr[1] + a[8] * rr[2], sg, lane);
  hh_col(F, 0, lam, sl, s0);
  const double norm0 = -F.Rm[0];
  F.w01 = F.beta[0] * d01;
  F.w02 = F.beta[0] * d02;
  F.Rm[1] = -F.w01 * F.v0[0];
  F.Rm[2] = -F.w02 * F.v0[0];
  double V[9];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    V[r * 3] = a[r * 3];
    V[r * 3 + 1] = a[r * 3 + 1] - F.w01 * a[r * 3];
    V[r * 3 + 2] = a[r * 3 + 2] - F.w02 * a[r * 3];
  }
  const double s1 = seg_sum(V[1] * V[1] + V[4] * V[4] + V[7] * V[7], sg, lane);
  const double d12 = seg_sum(V[1] * V[2] + V[4] * V[5] + V[7] * V[8], sg, lane);
  hh_col(F, 1, lam, sl, s1);
  const double norm1 = -F.Rm[3];
  F.w12 = F.beta[1] * d12;
  F.Rm[4] = -F.w12 * F.v0[1];
#pragma unroll
  for (int r = 0; r < 3; r++) V[r * 3 + 2] -= F.w12 * V[r * 3 + 1];
  const double s2 = seg_sum(V[2] * V[2] + V[5] * V[5] + V[8] * V[8], sg, lane);
  hh_col(F, 2, lam, sl, s2);
  const double g01 = d01 * (sl / norm0);
  const double g02 = d02 * (sl / norm0) - F.w12 * g01;
  const double g12 = d12 * (sl / norm1);
  wy_from_gram(F, g01, g02, g12);
  q1_rows(V, F, Q);
  const double u0 = c0, u1 = c1 - F.w01 * c0, u2 = c2 - F.w02 * c0 - F.w12 * u1;
  tl[0] = -(F.M[0] * u0);
  tl[1] = -(F.M[1] * u0 + F.M[3] * u1);
  tl[2] = -(F.M[2] * u0 + F.M[4] * u1 + F.M[5] * u2);
}

// long landmark (one warp, more than 32 observations): operands straight from global memory, direct atomics
__device__ __forceinline__ void qr_long_item(const Dev& P, const TileInfo& ti, int lane, int start, int cnt, double lam) {
  const double sl = sqrt(lam);
  const size_t No = (size_t)P.ld;
  const int Nl = P.n_point, nt = ti.nt;
  double* __restrict__ jq = P.JQ + ti.jq_off;
  LmFactor F;
    // long landmark: five sweeps over its rows (L1/L2 resident), whole-warp reductions
    const int lm = P.obs_point[start];
    double acc[3] = {0, 0, 0};
    double a[9];
    for (int i = lane; i < cnt; i += 32) {
      load9(P.Jl, No, start + i, a);
      acc[0] += a[0] * a[0] + a[3] * a[3] + a[6] * a[6];
      acc[1] += a[0] * a[1] + a[3] * a[4] + a[6] * a[7];
      acc[2] += a[0] * a[2] + a[3] * a[5] + a[6] * a[8];
    }
    hh_col(F, 0, lam, sl, warp_sum(acc[0]));
    F.w01 = F.beta[0] * warp_sum(acc[1]);
    F.w02 = F.beta[0] * warp_sum(acc[2]);
    F.Rm[1] = -F.w01 * F.v0[0];
    F.Rm[2] = -F.w02 * F.v0[0];
    F.w12 = 0.0;
    acc[0] = acc[1] = 0.0;
    double V[9];
    for (int i = lane; i < cnt; i += 32) {
      load9(P.Jl, No, start + i, a);
      v_rows(a, F, V);  // w12 = 0: V2 is the column-0-reduced third column
      acc[0] += V[1] * V[1] + V[4] * V[4] + V[7] * V[7];
      acc[1] += V[1] * V[2] + V[4] * V[5] + V[7] * V[8];
    }
    hh_col(F, 1, lam, sl, warp_sum(acc[0]));
    F.w12 = F.beta[1] * warp_sum(acc[1]);
    F.Rm[4] = -F.w12 * F.v0[1];
    double g[4] = {0, 0, 0, 0};
    for (int i = lane; i < cnt; i += 32) {
      load9(P.Jl, No, start + i, a);
      v_rows(a, F, V);
      g[0] += V[2] * V[2] + V[5] * V[5] + V[8] * V[8];
      g[1] += V[0] * V[1] + V[3] * V[4] + V[6] * V[7];
      g[2] += V[0] * V[2] + V[3] * V[5] + V[6] * V[8];
      g[3] += V[1] * V[2] + V[4] * V[5] + V[7] * V[8];
    }
    hh_col(F, 2, lam, sl, warp_sum(g[0]));
    const double g01 = warp_sum(g[1]), g02 = warp_sum(g[2]), g12 = warp_sum(g[3]);
    wy_from_gram(F, g01, g02, g12);
    double tl[3] = {0, 0, 0}, Q[9], rr[3];
    for (int i = lane; i < cnt; i += 32) {
      const int o = start + i;
      load9(P.Jl, No, o, a);
      v_rows(a, F, V);
      q1_rows(V, F, Q);
#pragma unroll
      for (int c = 0; c < 9; c++) jq[(size_t)(JG + c) * nt + (o - ti.o0)] = Q[c];
#pragma unroll
      for (int c = 0; c < 3; c++) rr[c] = P.r[(size_t)c * No + o];
#pragma unroll
      for (int k = 0; k < 3; k++) tl[k] += Q[k] * rr[0] + Q[3 + k] * rr[1] + Q[6 + k] * rr[2];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) tl[k] = warp_sum(tl[k]);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 6; c++) P.R[(size_t)c * Nl + lm] = F.Rm[c];
#pragma unroll
      for (int c = 0; c < 3; c++) P.tl[(size_t)c * Nl + lm] = tl[c];
    }
    double lc[27];
    for (int i = lane; i < cnt; i += 32) {
      const int o = start + i;
      const int slot = P.obs_slot[o];
      if (slot < 0) continue;
      load9(jq + (size_t)JG * nt, nt, o - ti.o0, Q);  // written by this same lane above
#pragma unroll
      for (int c = 0; c < 3; c++) rr[c] = P.r[(size_t)c * No + o];
      double J[18];
      load_Jp(P, jq, nt, o - ti.o0, slot, (P.obs_lp[o] & LP_STEREO) != 0, J);
      trial_contrib(J, Q, rr, tl, lc);
#pragma unroll
      for (int c = 0; c < 6; c++) atomicAdd(&P.bs[slot * 6 + c], lc[c]);
#pragma unroll
      for (int c = 0; c < 21; c++) atomicAdd(&P.D[slot * 21 + c], lc[6 + c]);
    }
}

__global__ void __launch_bounds__(CTA, 4) k_qr(Dev P, int force_all, double lam_override) {
  __shared__ double c_sh[27 * SCST];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const TileInfo ti = P.tiles[blockIdx.x];
  const bool valid = wid < ti.nitem;
  const int win = ti.win;
  const WinCtl& wc = P.ctl[win];
  if (!force_all && qr_phase(P, wc) != PH_TRIAL) return;  // CTA-uniform
  const double lam = force_all ? lam_override : wc.lambda;
  const int cnt = tile_item_cnt(ti, wid);
  const int start = tile_item_start(ti, wid);
  const size_t No = (size_t)P.ld;
  double* __restrict__ jq = P.JQ + ti.jq_off;
  // the 18 Jp rows are only needed after the factorisation: start pulling them into L2 now (one TMA prefetch per CTA)
  if (threadIdx.x == 0 && !ti.is_long) bulk_prefetch_l2(jq, (uint32_t)(JG * ti.nt * sizeof(double)));
  if (valid && cnt <= 32) {
    const bool act = lane < cnt;
    const int o = start + (act ? lane : 0);
    int lm = -1 - lane, rank = 0, slot = 0;
    bool has = false, stereo = false;
    double a[9], rr[3];
    if (act) {  // every load of this phase is issued before the first use
      const unsigned lp = P.obs_lp[o];
      lm = P.obs_point[o];
      load9(P.Jl, No, o, a);
#pragma unroll
      for (int c = 0; c < 3; c++) rr[c] = P.r[(size_t)c * No + o];
      has = (lp & 0xffffu) != 0xffffu;
      rank = (int)((lp >> 16) & 0x7fffu);
      stereo = (lp & LP_STEREO) != 0;
      if (has) slot = slot_of(P, lp, P.smallwin ? P.win_slot_ptr[win] : 0, o);
    } else {
#pragma unroll
      for (int c = 0; c < 9; c++) a[c] = 0.0;
      rr[0] = rr[1] = rr[2] = 0.0;
    }
    qr_short_item(P, ti, wid, lane, act, o, lm, has, rank, slot, stereo, a, rr, lam, c_sh);
  } else if (valid) {
    qr_long_item(P, ti, lane, start, cnt, lam);
  }
  __syncthreads();
  const int* runs = reinterpret_cast<const int*>(jq + (size_t)JQ_ROWS * ti.nt);
  const int sbase = P.smallwin ? P.win_slot_ptr[win] : 0;
  double* bs = P.bs;
  double* D = P.D;
  tile_scatter_all<27>(runs, runs + ti.nrun + 1, ti.nrun, sbase, c_sh,
                       [bs, D](int slot, int k) { return (k < 6) ? bs + (size_t)slot * 6 + k : D + (size_t)slot * 21 + (k - 6); });
}

// Pipelined variant: every CTA walks QR_TPB consecutive tiles and stages the NEXT tile's operands (this lane's rows of
// J_l, its residual, its meta words -- 14 cp.async per thread -- plus the tile descriptor two tiles ahead) in shared
// memory while it factorises the current one, so the three dependent global-load phases of the plain kernel
// (descriptor -> operands -> Jp) no longer sit on the critical path; the Jp rows and the run table of the next tile are
// pulled into L2 by a TMA prefetch.  Per-lane data is private (a lane reads back exactly what it copied), so the only
// barrier the staging needs is the one that publishes the tile descriptors.
constexpr int QR_TPB = 8;

constexpr size_t QR_PIPE_SMEM = (27 * SCST + 2 * 12 * CTA) * sizeof(double) + 2 * CTA * (sizeof(unsigned) + sizeof(int)) + 3 * 80;
__global__ void __launch_bounds__(CTA, 4) k_qr_pipe(Dev P, int force_all, double lam_override) {
  extern __shared__ __align__(16) unsigned char qr_smem[];
  static_assert(sizeof(TileInfo) == 80, "tile descriptors are staged with five 16-byte copies");
  TileInfo* ti_sh = reinterpret_cast<TileInfo*>(qr_smem);                       // [3]
  double* c_sh = reinterpret_cast<double*>(qr_smem + 3 * 80);                   // [27][SCST]
  double(*op_sh)[12][CTA] = reinterpret_cast<double(*)[12][CTA]>(c_sh + 27 * SCST);  // [2]: 9 rows of J_l, 3 of r, column = thread
  unsigned(*lp_sh)[CTA] = reinterpret_cast<unsigned(*)[CTA]>(op_sh + 2);        // [2]
  int(*lm_sh)[CTA] = reinterpret_cast<int(*)[CTA]>(lp_sh + 2);                  // [2]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t0 = blockIdx.x * QR_TPB, t1 = min(t0 + QR_TPB, P.n_tile);
  const size_t No = (size_t)P.ld;
  // stage the operands of tile t (descriptor `ti`) into buffer b; returns nothing -- completion via wait_group
  auto stage = [&](const TileInfo& ti, int b) {
    if (ti.is_long) return;
    const int cnt = tile_item_cnt(ti, wid), start = tile_item_start(ti, wid);
    if (wid < ti.nitem && lane < cnt) {
      const int o = start + lane;
#pragma unroll
      for (int c = 0; c < 9; c++) cp_async8(&op_sh[b][c][tid], P.Jl + (size_t)c * No + o);
#pragma unroll
      for (int c = 0; c < 3; c++) cp_async8(&op_sh[b][9 + c][tid], P.r + (size_t)c * No + o);
      cp_async4(&lp_sh[b][tid], P.obs_lp + o);
      cp_async4(&lm_sh[b][tid], P.obs_point + o);
    }
    if (tid == 0) bulk_prefetch_l2(P.JQ + ti.jq_off - JQ_HDR, (uint32_t)(ti.blk_doubles * sizeof(double)));
  };
  if (tid < 5) {
    cp_async16(reinterpret_cast<char*>(&ti_sh[0]) + tid * 16, reinterpret_cast<const char*>(P.tiles + t0) + tid * 16);
    if (t0 + 1 < t1)
      cp_async16(reinterpret_cast<char*>(&ti_sh[1]) + tid * 16, reinterpret_cast<const char*>(P.tiles + t0 + 1) + tid * 16);
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  stage(ti_sh[0], 0);
  cp_async_commit();
  // window state of the tile being processed next (phase, lambda): loaded one tile ahead as well
  int nphase = force_all ? PH_TRIAL : qr_phase(P, P.ctl[ti_sh[0].win]);
  double nlam = force_all ? lam_override : P.ctl[ti_sh[0].win].lambda;
  for (int k = 0; k < t1 - t0; k++) {
    const int buf = k & 1;
    cp_async_wait_all();  // operands of tile k (own lanes) and the descriptor of tile k+1 have landed
    __syncthreads();      // ... and the descriptor is visible to everybody; c_sh of the previous tile is free
    const TileInfo ti = ti_sh[k % 3];
    const int phase = nphase;
    const double lam = nlam;
    if (k + 1 < t1 - t0) {
      const TileInfo& tn = ti_sh[(k + 1) % 3];
      if (k + 2 < t1 - t0 && tid < 5)
        cp_async16(reinterpret_cast<char*>(&ti_sh[(k + 2) % 3]) + tid * 16,
                   reinterpret_cast<const char*>(P.tiles + t0 + k + 2) + tid * 16);
      stage(tn, buf ^ 1);
      if (!force_all) { nphase = qr_phase(P, P.ctl[tn.win]); nlam = P.ctl[tn.win].lambda; }
    }
    cp_async_commit();
    if (phase != PH_TRIAL) continue;  // CTA-uniform: the tile's window is not in a trial
    const bool valid = wid < ti.nitem;
    const int cnt = tile_item_cnt(ti, wid);
    const int start = tile_item_start(ti, wid);
    if (valid && cnt <= 32) {
      const bool act = lane < cnt;
      const int o = start + (act ? lane : 0);
      int lm = -1 - lane, rank = 0, slot = 0;
      bool has = false, stereo = false;
      double a[9], rr[3];
      if (act) {
        const unsigned lp = lp_sh[buf][tid];
        lm = lm_sh[buf][tid];
#pragma unroll
        for (int c = 0; c < 9; c++) a[c] = op_sh[buf][c][tid];
#pragma unroll
        for (int c = 0; c < 3; c++) rr[c] = op_sh[buf][9 + c][tid];
        has = (lp & 0xffffu) != 0xffffu;
        rank = (int)((lp >> 16) & 0x7fffu);
        stereo = (lp & LP_STEREO) != 0;
        if (has) slot = slot_of(P, lp, P.smallwin ? P.win_slot_ptr[ti.win] : 0, o);
      } else {
#pragma unroll
        for (int c = 0; c < 9; c++) a[c] = 0.0;
        rr[0] = rr[1] = rr[2] = 0.0;
      }
      qr_short_item(P, ti, wid, lane, act, o, lm, has, rank, slot, stereo, a, rr, lam, c_sh);
    } else if (valid) {
      qr_long_item(P, ti, lane, start, cnt, lam);
    }
    __syncthreads();
    const double* jq = P.JQ + ti.jq_off;
    const int* runs = reinterpret_cast<const int*>(jq + (size_t)JQ_ROWS * ti.nt);
    const int sbase = P.smallwin ? P.win_slot_ptr[ti.win] : 0;
    double* bs = P.bs;
    double* D = P.D;
    tile_scatter_all<27>(runs, runs + ti.nrun + 1, ti.nrun, sbase, c_sh,
                         [bs, D](int slot, int k2) { return (k2 < 6) ? bs + (size_t)slot * 6 + k2 : D + (size_t)slot * 21 + (k2 - 6); });
  }
  cp_async_wait_all();
}

// Second version of the pipelined kernel (the default): qr_short_item_v2, the run table of the next tile staged in
// shared memory together with its operands (the pose-side reduction no longer starts with a dependent L2 round trip),
// and ONE operand buffer -- a lane copies the next tile's operands into the slots it has just read its own from, so no
// second buffer and no extra barrier are needed.
constexpr int QR_RUN_INTS = 2 * CTA + 2;  // a tile has at most CTA runs: nrun + 1 offsets, nrun slots
constexpr size_t QR_PIPE2_SMEM = (27 * SCST + 12 * CTA) * sizeof(double) + CTA * (sizeof(unsigned) + sizeof(int)) +
                                 2 * QR_RUN_INTS * sizeof(int) + 3 * 80;
template <int JPOS, int MINB>
__global__ void __launch_bounds__(CTA, MINB) k_qr_pipe2(Dev P, int force_all, double lam_override) {
  extern __shared__ __align__(16) unsigned char qr_smem[];
  TileInfo* ti_sh = reinterpret_cast<TileInfo*>(qr_smem);                             // [3]
  double* c_sh = reinterpret_cast<double*>(qr_smem + 3 * 80);                         // [27][SCST]
  double(*op_sh)[CTA] = reinterpret_cast<double(*)[CTA]>(c_sh + 27 * SCST);           // [12]: 9 rows of J_l, 3 of r
  unsigned* lp_sh = reinterpret_cast<unsigned*>(op_sh + 12);                          // [CTA]
  int* lm_sh = reinterpret_cast<int*>(lp_sh + CTA);                                   // [CTA]
  int(*run_sh)[QR_RUN_INTS] = reinterpret_cast<int(*)[QR_RUN_INTS]>(lm_sh + CTA);     // [2]
  double* qr_acc = reinterpret_cast<double*>(qr_smem + QR_PIPE2_SMEM);                // reproducible mode: [27 * maxslot]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t0 = blockIdx.x * QR_TPB, t1 = min(t0 + QR_TPB, P.n_tile);
  const size_t No = (size_t)P.ld;
  int det_win = -1;
  if (P.det) {
    for (int i = tid; i < 27 * P.maxslot; i += CTA) qr_acc[i] = 0.0;
    __syncthreads();
  }
  auto stage = [&](const TileInfo& ti, int rb) {
    if (ti.is_long) return;
    const int cnt = tile_item_cnt(ti, wid), start = tile_item_start(ti, wid);
    if (wid < ti.nitem && lane < cnt) {
      const int o = start + lane;
#pragma unroll
      for (int c = 0; c < 9; c++) cp_async8(&op_sh[c][tid], P.Jl + (size_t)c * No + o);
#pragma unroll
      for (int c = 0; c < 3; c++) cp_async8(&op_sh[9 + c][tid], P.r + (size_t)c * No + o);
      cp_async4(&lp_sh[tid], P.obs_lp + o);
      cp_async4(&lm_sh[tid], P.obs_point + o);
    }
    const int* runs = reinterpret_cast<const int*>(P.JQ + ti.jq_off + (size_t)JQ_ROWS * ti.nt);
    for (int i = tid; i < 2 * ti.nrun + 1; i += CTA) cp_async4(&run_sh[rb][i], runs + i);
    if (tid == 0) bulk_prefetch_l2(P.JQ + ti.jq_off, (uint32_t)(JG * ti.nt * sizeof(double)));
  };
  if (tid < 5) {
    cp_async16(reinterpret_cast<char*>(&ti_sh[0]) + tid * 16, reinterpret_cast<const char*>(P.tiles + t0) + tid * 16);
    if (t0 + 1 < t1)
      cp_async16(reinterpret_cast<char*>(&ti_sh[1]) + tid * 16, reinterpret_cast<const char*>(P.tiles + t0 + 1) + tid * 16);
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  stage(ti_sh[0], 0);
  cp_async_commit();
  int nphase = force_all ? PH_TRIAL : qr_phase(P, P.ctl[ti_sh[0].win]);
  double nlam = force_all ? lam_override : P.ctl[ti_sh[0].win].lambda;
  for (int k = 0; k < t1 - t0; k++) {
    cp_async_wait_all();  // operands + run table of tile k, descriptor of tile k+1
    __syncthreads();      // ... visible to everybody; c_sh and the other run buffer are free
    const TileInfo ti = ti_sh[k % 3];
    const int phase = nphase;
    const double lam = nlam;
    const bool valid = wid < ti.nitem;
    const int cnt = tile_item_cnt(ti, wid);
    const int start = tile_item_start(ti, wid);
    const bool is_short = valid && cnt <= 32;
    const bool act = is_short && lane < cnt;
    int lm = -1 - lane, rank = 0;
    unsigned lpw = 0xffffu;
    bool has = false;
    double a[9], rr[3];
    if (act) {  // this lane's operands, copied by itself one tile ago
      const unsigned lp = lp_sh[tid];
      lpw = lp;
      lm = lm_sh[tid];
#pragma unroll
      for (int c = 0; c < 9; c++) a[c] = op_sh[c][tid];
#pragma unroll
      for (int c = 0; c < 3; c++) rr[c] = op_sh[9 + c][tid];
      has = (lp & 0xffffu) != 0xffffu;
      rank = (int)((lp >> 16) & 0x7fffu);
    } else {
#pragma unroll
      for (int c = 0; c < 9; c++) a[c] = 0.0;
      rr[0] = rr[1] = rr[2] = 0.0;
    }
    if (k + 1 < t1 - t0) {  // the slots just read are free again: bring in tile k+1
      const TileInfo& tn = ti_sh[(k + 1) % 3];
      if (k + 2 < t1 - t0 && tid < 5)
        cp_async16(reinterpret_cast<char*>(&ti_sh[(k + 2) % 3]) + tid * 16,
                   reinterpret_cast<const char*>(P.tiles + t0 + k + 2) + tid * 16);
      stage(tn, (k + 1) & 1);
      if (!force_all) { nphase = qr_phase(P, P.ctl[tn.win]); nlam = P.ctl[tn.win].lambda; }
    }
    cp_async_commit();
    if (phase != PH_TRIAL) continue;  // CTA-uniform: the tile's window is not in a trial
    const int sbase = P.smallwin ? P.win_slot_ptr[ti.win] : 0;  // for the reduction below: requested before the factorisation
    if (P.det && ti.win != det_win) {  // a new window starts: the finished one leaves the CTA as one partial vector
      if (det_win >= 0)
        det_flush<27>(P.part_qr, blockIdx.x + det_win, P.maxslot, P.win_slot_ptr[det_win + 1] - P.win_slot_ptr[det_win], qr_acc);
      det_win = ti.win;
    }
    if (is_short) {
      const int slot = has ? slot_of(P, lpw, sbase, start + lane) : 0;
      qr_short_item_v2<JPOS>(P, ti, wid, lane, act, lm, has, rank, slot, (lpw & LP_STEREO) != 0, a, rr, lam, c_sh);
    } else if (valid) {
      qr_long_item(P, ti, lane, start, cnt, lam);
    }
    __syncthreads();
    double* bs = P.bs;
    double* D = P.D;
    const int* runs = run_sh[k & 1];
    if (P.det) tile_scatter_acc<27>(runs, runs + ti.nrun + 1, ti.nrun, c_sh, qr_acc);
    else
      tile_scatter_all<27>(runs, runs + ti.nrun + 1, ti.nrun, sbase, c_sh,
                           [bs, D](int slot, int k2) { return (k2 < 6) ? bs + (size_t)slot * 6 + k2 : D + (size_t)slot * 21 + (k2 - 6); });
  }
  cp_async_wait_all();
  if (P.det && det_win >= 0)
    det_flush<27>(P.part_qr, blockIdx.x + det_win, P.maxslot, P.win_slot_ptr[det_win + 1] - P.win_slot_ptr[det_win], qr_acc);
}

// Reproducible mode: reduced right-hand side and block-Jacobi blocks of a window = its QR CTAs' partial vectors in order
__global__ void __launch_bounds__(RCTA) k_reduce_qr(Dev P) {
  const int win = blockIdx.x;
  if (qr_phase(P, P.ctl[win]) != PH_TRIAL || P.win_tile_ptr[win + 1] <= P.win_tile_ptr[win]) return;
  const int c_lo = P.win_tile_ptr[win] / QR_TPB, c_hi = (P.win_tile_ptr[win + 1] - 1) / QR_TPB;
  const int s0 = P.win_slot_ptr[win], wn = P.win_slot_ptr[win + 1] - s0;
  for (int i = threadIdx.x; i < 27 * wn; i += RCTA) {
    const double v = det_sum(P.part_qr, c_lo, c_hi, win, (size_t)27 * P.maxslot, i);
    const int sl = i / 27, k = i - sl * 27;
    if (k < 6) P.bs[(size_t)(s0 + sl) * 6 + k] = v; else P.D[(size_t)(s0 + sl) * 21 + (k - 6)] = v;
  }
}

// ------------------------------------------------------------------------------------------------ K1+K2 fused
// Linearisation and landmark QR in ONE pass over the observations, for the LM iterations whose lambda is already known
// when the window re-linearises (every iteration but the first of a pass: lambda_0 = tau * max diag needs a complete
// linearisation first) and for the retries after a rejected trial.  The rows of J_l and the weighted residual never
// leave the registers: against the two separate kernels the round trip of 96 B per observation through the J_l / r
// planes (written by the linearisation, read back by the QR) and the re-read of the geometry rows disappear -- what is
// left is SURVEY 8(d)'s figure for K1+K2: the observation record in, e and the matvec operand out, R / t_l / b_l per
// landmark.
//   mode FUSED (window in PH_LIN, iteration > 0): everything the linearisation produces (stored error, chi2, b_p,
//        diag(Jp^T Jp), b_l, max diag) + everything the QR produces (Q1 rows, R, t_l, reduced rhs, block-Jacobi blocks);
//   mode RETRY (window in PH_TRIAL after a rejected trial, new lambda): the estimates were restored to the linearisation
//        point, so the SAME Jacobians are recomputed (bit-identical) instead of being read back; only the QR's outputs
//        are written -- the stored errors keep the rejected trial's values (g2o's stale _error).
// Long landmarks (> 32 observations) run the two long-item routines back to back (they go through the planes).
// Structure: the pipelined linearisation's (LIN_TPB tiles per CTA, descriptors / run tables / observation words
// staged one tile ahead).
constexpr int FZ_NV = 39;  // pose-side values per observation: reduced rhs 6 + block-Jacobi 21 + gradient 6 + diagonal 6
__device__ __forceinline__ void linqr_tile(const Dev& P, const TileInfo& ti, const LinOps& pre, int robust, double d2, double d3,
                                           bool full, double lam, double* c_sh, const int* runs_staged) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int w = ti.item0 + wid;
  const bool valid = wid < ti.nitem;
  const int win = ti.win;
  const int start = tile_item_start(ti, wid), cnt = tile_item_cnt(ti, wid);
  const size_t No = (size_t)P.ld; const int Nl = P.n_point, nt = ti.nt;
  const bool is_long = cnt > 32;
  double* __restrict__ jq = P.JQ + ti.jq_off;
  const int sbase = P.smallwin ? P.win_slot_ptr[win] : 0;
  double chi_acc = 0.0, maxd = 0.0;
  if (valid && !is_long) {
    const bool act = lane < cnt;
    const int o = start + (act ? lane : 0);
    int lm = -1 - lane, rank = 0;
    bool has = false, live = false;
    ObsLin L;
    double a[9], rr[3];
#pragma unroll
    for (int c = 0; c < 9; c++) a[c] = 0.0;
    rr[0] = rr[1] = rr[2] = 0.0;
#pragma unroll
    for (int c = 0; c < JG; c++) L.g[c] = 0.0;
    L.cam3[0] = L.cam3[1] = L.cam3[2] = 0.0;
    L.stereo = false;
    const LinOps q = pre;
    if (act) {
      has = (q.lp & 0xffffu) != 0xffffu;
      rank = (int)((q.lp >> 16) & 0x7fffu);
      lm = q.lm;
      live = q.live == 0;
      obs_eval_ops(P, q.m, q.ip, q.lm, true, robust != 0, d2, d3, L);
      if (full && live) {
#pragma unroll
        for (int c = 0; c < 3; c++) P.err[(size_t)c * No + o] = L.e[c];
        chi_acc += L.rho0;
      }
      // an excluded (level-1) edge contributes zero rows; select, do not multiply (its Jacobian may be inf/NaN)
#pragma unroll
      for (int c = 0; c < JG; c++) L.g[c] = live ? L.g[c] : 0.0;
#pragma unroll
      for (int c = 0; c < 9; c++) a[c] = live ? L.Jl[c] * L.w : 0.0;
#pragma unroll
      for (int c = 0; c < 3; c++) rr[c] = live ? L.e[c] * L.w : 0.0;
    }
    const int fcol = tile_fcol(ti, wid, has, lane);
    if (full && act && has) {
#pragma unroll
      for (int c = 0; c < JG; c++) jq[(size_t)c * nt + fcol] = L.g[c];
    }
    const Seg sg = seg_of(lm, lane);
    if (full) {  // landmark-side gradient and Hessian diagonal
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double g = a[c] * rr[0] + a[3 + c] * rr[1] + a[6 + c] * rr[2];
        double h = a[c] * a[c] + a[3 + c] * a[3 + c] + a[6 + c] * a[6 + c];
        g = seg_sum(g, sg, lane);
        h = seg_sum(h, sg, lane);
        if (act && lane == sg.start) P.bl[(size_t)c * Nl + lm] = -g;
        maxd = fmax(maxd, h);
      }
    }
    LmFactor F;
    double Q[9], tl[3];
    qr_factor_v2(a, rr, lam, sg, lane, F, Q, tl);
    if (act) {
      if (lane == sg.start) {
#pragma unroll
        for (int c = 0; c < 6; c++) P.R[(size_t)c * Nl + lm] = F.Rm[c];
#pragma unroll
        for (int c = 0; c < 3; c++) P.tl[(size_t)c * Nl + lm] = tl[c];
      }
      if (has) {
#pragma unroll
        for (int c = 0; c < 9; c++) jq[(size_t)(JG + c) * nt + fcol] = Q[c];
        double J[18];
        jp_full(jp_compact(L.g, L.cam3[0], L.cam3[1], L.cam3[2], L.stereo), L.stereo, J);
        trial_contrib_regs(J, Q, rr, tl, c_sh, rank);  // rows 0-26 of this observation's column
        if (full) {
#pragma unroll
          for (int c = 0; c < 6; c++) {
            c_sh[(27 + c) * SCST + rank] = -(J[c] * rr[0] + J[6 + c] * rr[1] + J[12 + c] * rr[2]);
            c_sh[(33 + c) * SCST + rank] = J[c] * J[c] + J[6 + c] * J[6 + c] + J[12 + c] * J[12 + c];
          }
        }
      }
    }
  } else if (valid) {
    if (full) linearize_long_item(P, ti, lane, start, cnt, robust, d2, d3, chi_acc, maxd);
    qr_long_item(P, ti, lane, start, cnt, lam);  // its planes are those of the window's last linearisation
  }
  if (valid && full) {
    chi_acc = warp_sum(chi_acc);
    maxd = warp_max(maxd);
    if (lane == 0) {
      P.chi_part[w] = chi_acc;
      atomic_max_pos(&P.ctl[win].maxdiag_bits, maxd);
    }
  }
  __syncthreads();
  double* bs = P.bs; double* D = P.D; double* bp = P.bp; double* hd = P.hd;
  auto addr = [bs, D, bp, hd](int slot, int k) {
    return (k < 6) ? bs + (size_t)slot * 6 + k
                   : (k < 27) ? D + (size_t)slot * 21 + (k - 6)
                              : (k < 33) ? bp + (size_t)slot * 6 + (k - 27) : hd + (size_t)slot * 6 + (k - 33);
  };
  if (full) tile_scatter_all<FZ_NV>(runs_staged, runs_staged + ti.nrun + 1, ti.nrun, sbase, c_sh, addr);
  else tile_scatter_all<27>(runs_staged, runs_staged + ti.nrun + 1, ti.nrun, sbase, c_sh, addr);
}

__global__ void __launch_bounds__(CTA, 3) k_linqr_pipe(Dev P, int robust, double d2, double d3, int force_mode, double lam_override) {
  __shared__ double c_sh[FZ_NV * SCST];
  __shared__ __align__(16) TileInfo ti_sh[3];
  __shared__ int run_sh[2][LIN_RUN_INTS];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t0 = blockIdx.x * LIN_TPB, t1 = min(t0 + LIN_TPB, P.n_tile);
  auto stage_runs = [&](const TileInfo& ti, int rb) {
    if (ti.is_long) return;
    const int* runs = reinterpret_cast<const int*>(P.JQ + ti.jq_off + (size_t)JQ_ROWS * ti.nt);
    for (int i = tid; i < 2 * ti.nrun + 1; i += CTA) cp_async4(&run_sh[rb][i], runs + i);
  };
  auto ops_of = [&](const TileInfo& ti) {
    LinOps q{};
    if (!ti.is_long) {
      const int cnt = tile_item_cnt(ti, wid), start = tile_item_start(ti, wid);
      if (wid < ti.nitem && lane < cnt) q = lin_load_ops(P, start + lane);
    }
    return q;
  };
  if (tid < 5) {
    cp_async16(reinterpret_cast<char*>(&ti_sh[0]) + tid * 16, reinterpret_cast<const char*>(P.tiles + t0) + tid * 16);
    if (t0 + 1 < t1)
      cp_async16(reinterpret_cast<char*>(&ti_sh[1]) + tid * 16, reinterpret_cast<const char*>(P.tiles + t0 + 1) + tid * 16);
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  LinOps nxt = ops_of(ti_sh[0]);
  stage_runs(ti_sh[0], 0);
  cp_async_commit();
  // window state of the tile processed next (mode, lambda, robust), loaded one tile ahead
  int nmode = force_mode ? force_mode : linqr_mode(P, P.ctl[ti_sh[0].win]);
  double nlam = force_mode ? lam_override : P.ctl[ti_sh[0].win].lambda;
  int nrob = robust < 0 ? P.ctl[ti_sh[0].win].robust : robust;
  for (int k = 0; k < t1 - t0; k++) {
    cp_async_wait_all();
    __syncthreads();
    const TileInfo ti = ti_sh[k % 3];
    const LinOps cur = nxt;
    const int mode = nmode, rob = nrob;
    const double lam = nlam;
    if (k + 1 < t1 - t0) {
      const TileInfo& tn = ti_sh[(k + 1) % 3];
      if (k + 2 < t1 - t0 && tid < 5)
        cp_async16(reinterpret_cast<char*>(&ti_sh[(k + 2) % 3]) + tid * 16,
                   reinterpret_cast<const char*>(P.tiles + t0 + k + 2) + tid * 16);
      nxt = ops_of(tn);
      stage_runs(tn, (k + 1) & 1);
      if (!force_mode) { nmode = linqr_mode(P, P.ctl[tn.win]); nlam = P.ctl[tn.win].lambda; }
      if (robust < 0) nrob = P.ctl[tn.win].robust;
    }
    cp_async_commit();
    if (mode == 0) continue;  // CTA-uniform
    if (!ti.is_long) {
      const int li = ti.nitem - 1;
      if (wid == li && lane == tile_item_cnt(ti, li) - 1) {  // pull the next tiles' landmark coordinates towards L2
        long long first = ((long long)cur.lm + 1) & ~1ll;
        long long bytes = ((long long)P.n_point - first) * 24;
        bytes = (bytes < 2304 ? bytes : 2304) & ~15ll;
        if (bytes > 0) bulk_prefetch_l2(P.point + first * 3, (uint32_t)bytes);
      }
    }
    linqr_tile(P, ti, cur, rob, d2, d3, mode == 1, lam, c_sh, run_sh[k & 1]);
  }
  cp_async_wait_all();
}

// 6x6 block-Jacobi inverse per pose slot: (D + lambda I)^-1
__device__ __forceinline__ void dinv_slot(const Dev& P, int s, double lam) {
  double A[36], Ai[36];
  int idx = 0;
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = a; b < 6; b++) {
      const double v = P.D[s * 21 + idx] + (a == b ? lam : 0.0);
      A[a * 6 + b] = v;
      A[b * 6 + a] = v;
      idx++;
    }
  if (!spd6_inverse(A, Ai)) {
    for (int i = 0; i < 36; i++) Ai[i] = 0.0;
    for (int i = 0; i < 6; i++) Ai[i * 6 + i] = 1.0 / fmax(fabs(A[i * 6 + i]), 1e-300);
  }
  for (int i = 0; i < 36; i++) P.Dinv[s * 36 + i] = Ai[i];
}
__global__ void k_dinv(Dev P, int force_all, double lam_override) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.n_slot) return;
  const WinCtl& c = P.ctl[P.slot_win[s]];
  if (!force_all && c.phase != PH_TRIAL) return;
  dinv_slot(P, s, force_all ? lam_override : c.lambda);
}

// ------------------------------------------------------------------------------------------------ K3: matvec
// q += sum_l Jp^T (I - Q1 Q1^T) Jp p  over the landmarks of this warp's item (lambda*p is added by the CG
// update).  Per observation: v = Jp p[slot] (d-vector), s_l = sum Q1^T v (3-vector, landmark reduction),
// u = v - Q1 s_l, scatter-add Jp^T u.  Streams Jp (18) + Q1 (9) planes: 216 B/observation.
// Observations of fixed poses have no pose columns (g2o hessianIndex -1): they are skipped entirely.
// The search direction of the persistent big-window PCG is never stored between iterations: owners apply the CG update
// one phase late, and everybody else evaluates  p = (z - alpha Dq) + beta p_old  where it is needed.
struct PEff {
  const double *z, *dq, *p;
  double alpha, beta;
};
// Plain (L1-cached) loads: the three vectors only change between grid barriers, whose acquire makes the owners' writes
// visible to ordinary loads (the cooperative-groups grid.sync contract); repeated far-band poses then hit L1.
__device__ __forceinline__ double peff_at(const PEff& pe, size_t e) {
  return (pe.z[e] - pe.alpha * pe.dq[e]) + pe.beta * pe.p[e];
}

// PSRC: where p lives -- 0 global (slot-major), 1 global read through L2 (rewritten by other CTAs during the kernel),
// 2 shared memory, component-major with stride `pstride`, 3 evaluated on the fly from a PEff
template <int PSRC = 0>
__device__ __forceinline__ void matvec_obs_v(const Dev& P, const double* __restrict__ jq, int nt, int col,
                                             const double* pvec, int slot, bool stereo, double J[18], double Q[9],
                                             double v[3], int pstride = 0, const PEff* pe = nullptr) {
  load_Jp(P, jq, nt, col, slot, stereo, J);
#pragma unroll
  for (int c = 0; c < 9; c++) Q[c] = jq[(size_t)(JG + c) * nt + col];
  double pp[6];
#pragma unroll
  for (int c = 0; c < 6; c++)
    pp[c] = (PSRC == 3) ? peff_at(*pe, (size_t)slot * 6 + c)
                        : (PSRC == 2) ? pvec[c * pstride + slot] : ((PSRC == 1) ? __ldcg(pvec + slot * 6 + c) : pvec[slot * 6 + c]);
#pragma unroll
  for (int r = 0; r < 3; r++)
    v[r] = J[r * 6] * pp[0] + J[r * 6 + 1] * pp[1] + J[r * 6 + 2] * pp[2] + J[r * 6 + 3] * pp[3] +
           J[r * 6 + 4] * pp[4] + J[r * 6 + 5] * pp[5];
}

// long landmark (one warp, more than 32 observations): two sweeps, direct atomics
template <int PSRC = 0>
__device__ __forceinline__ void matvec_long_item(const Dev& P, const double* __restrict__ jq, int nt,
                                                 const double* pvec, double* qvec,
                                                 int start, int cnt, int lane, int pstride = 0, const PEff* pe = nullptr) {
  double J[18], Q[9], v[3], sv[3] = {0, 0, 0};
  for (int i = lane; i < cnt; i += 32) {
    const int o = start + i;
    const int slot = P.obs_slot[o];
    if (slot < 0) continue;
    matvec_obs_v<PSRC>(P, jq, nt, i, pvec, slot, (P.obs_lp[o] & LP_STEREO) != 0, J, Q, v, pstride, pe);  // a long item is a tile of its own: column = i
#pragma unroll
    for (int k = 0; k < 3; k++) sv[k] += Q[k] * v[0] + Q[3 + k] * v[1] + Q[6 + k] * v[2];
  }
#pragma unroll
  for (int k = 0; k < 3; k++) sv[k] = warp_sum(sv[k]);
  for (int i = lane; i < cnt; i += 32) {
    const int o = start + i;
    const int slot = P.obs_slot[o];
    if (slot < 0) continue;
    matvec_obs_v<PSRC>(P, jq, nt, i, pvec, slot, (P.obs_lp[o] & LP_STEREO) != 0, J, Q, v, pstride, pe);
#pragma unroll
    for (int r = 0; r < 3; r++) v[r] -= Q[r * 3] * sv[0] + Q[r * 3 + 1] * sv[1] + Q[r * 3 + 2] * sv[2];
#pragma unroll
    for (int c = 0; c < 6; c++) atomicAdd(&qvec[slot * 6 + c], J[c] * v[0] + J[6 + c] * v[1] + J[12 + c] * v[2]);
  }
}

// per-observation matvec math once J (3x6), Q (3x3) and p (6) are in registers; the landmark sum is a segmented shuffle
__device__ __forceinline__ void matvec_lane(bool has, const double J[18], const double Q[9], const double pp[6],
                                            const Seg& sg, int lane, double out[6]) {
  double v[3] = {0, 0, 0}, sv[3];
  if (has) {
#pragma unroll
    for (int r = 0; r < 3; r++)
      v[r] = J[r * 6] * pp[0] + J[r * 6 + 1] * pp[1] + J[r * 6 + 2] * pp[2] + J[r * 6 + 3] * pp[3] +
             J[r * 6 + 4] * pp[4] + J[r * 6 + 5] * pp[5];
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double tt = has ? (Q[k] * v[0] + Q[3 + k] * v[1] + Q[6 + k] * v[2]) : 0.0;
    sv[k] = seg_sum(tt, sg, lane);
  }
  if (has) {
#pragma unroll
    for (int r = 0; r < 3; r++) v[r] -= Q[r * 3] * sv[0] + Q[r * 3 + 1] * sv[1] + Q[r * 3 + 2] * sv[2];
#pragma unroll
    for (int c = 0; c < 6; c++) out[c] = J[c] * v[0] + J[6 + c] * v[1] + J[12 + c] * v[2];
  }
}

// General matvec (any window size): one CTA per tile, planes read straight from global memory, p gathered from L2,
// pose-side sums reduced in the CTA and flushed with one atomic per (tile, slot, component).
__global__ void __launch_bounds__(CTA) k_matvec(Dev P, const double* __restrict__ pvec, double* __restrict__ qvec,
                                                int force_all) {
  __shared__ double c_sh[6 * CTA];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const TileInfo ti = P.tiles[blockIdx.x];
  if (!force_all && !P.ctl[ti.win].cg_active) return;  // CTA-uniform
  const int w = ti.item0 + wid;
  const bool valid = wid < ti.nitem;
  const int start = valid ? P.item_start[w] : 0, cnt = valid ? P.item_cnt[w] : 0;
  const double* __restrict__ jq = P.JQ + ti.jq_off;
  const int nt = ti.nt;
  double out[6] = {0, 0, 0, 0, 0, 0};
  bool has = false;
  int rank = 0, key = 0;
  if (valid && cnt <= 32) {
    const bool act = lane < cnt;
    const int o = start + (act ? lane : 0);
    int lm = -1 - lane;
    bool stereo = false;
    if (act) {
      const unsigned lp = P.obs_lp[o];
      has = (lp & 0xffffu) != 0xffffu;
      rank = (int)((lp >> 16) & 0x7fffu);
      stereo = (lp & LP_STEREO) != 0;
      key = P.obs_slot[o];
      lm = P.obs_point[o];
    }
    const Seg sg = seg_of(lm, lane);
    const int fcol = tile_fcol(ti, wid, has, lane);
    double J[18], Q[9], pp[6];
    if (has) {
      const int col = fcol;
      load_Jp(P, jq, nt, col, key, stereo, J);
#pragma unroll
      for (int c = 0; c < 9; c++) Q[c] = jq[(size_t)(JG + c) * nt + col];
#pragma unroll
      for (int c = 0; c < 6; c++) pp[c] = pvec[(size_t)key * 6 + c];
    }
    matvec_lane(has, J, Q, pp, sg, lane, out);
  } else if (valid) {
    matvec_long_item(P, jq, nt, pvec, qvec, start, cnt, lane);
  }
  const int* runs = reinterpret_cast<const int*>(jq + (size_t)JQ_ROWS * nt);
  const int sbase = P.smallwin ? P.win_slot_ptr[ti.win] : 0;
  tile_scatter<6>(runs, runs + ti.nrun + 1, ti.nrun, sbase, c_sh, out, has, rank, qvec, 6, 0);
}

// Writes the static part of every JQ block: the 64-byte header and the per-column meta entries.
// header ints: [0] nitem, [1] win, [2] first free slot of the window, [3] free slots of the window, [4] nfree,
//              [5] nt, [6] o0, [7]/[8] jq_off lo/hi, [9..12] observations per item, [13] is_long, [14] nrun,
//              [15] lowest run slot of the tile (-1: none)
__global__ void k_init_jq(Dev P) {
  const TileInfo ti = P.tiles[blockIdx.x];
  double* blk = P.JQ + ti.jq_off;
  if (threadIdx.x == 0) {
    int* h = reinterpret_cast<int*>(blk - JQ_HDR);
    h[0] = ti.nitem; h[1] = ti.win; h[2] = P.win_slot_ptr[ti.win];
    h[3] = P.win_slot_ptr[ti.win + 1] - P.win_slot_ptr[ti.win];
    h[4] = ti.nfree; h[5] = ti.nt; h[6] = ti.o0;
    h[7] = (int)(ti.jq_off & 0xffffffffll); h[8] = (int)(ti.jq_off >> 32);
    // matvec lanes = columns: free-pose observations per item (a long tile keeps one column per observation)
    for (int i = 0; i < 4; i++) h[9 + i] = ti.is_long ? (i == 0 ? ti.cnt[0] : 0) : ti.fcnt[i];
    h[13] = ti.is_long; h[14] = ti.nrun;
    // lowest slot of the tile (run slots are sorted): anchors the shared accumulator window of the big-window matvec
    h[15] = (ti.nrun > 0) ? P.tile_runs[P.tile_run_ptr[blockIdx.x] + ti.nrun + 1] : -1;
  }
  int* runs = reinterpret_cast<int*>(blk + (size_t)JQ_ROWS * ti.nt);
  const int r0 = P.tile_run_ptr[blockIdx.x], nr = P.tile_run_ptr[blockIdx.x + 1] - r0;
  for (int i = threadIdx.x; i < nr; i += blockDim.x) runs[i] = P.tile_runs[r0 + i];
  uint2* meta = reinterpret_cast<uint2*>(blk + (size_t)NPLANE * ti.nt);
  if (ti.is_long) {
    for (int col = threadIdx.x; col < ti.nt; col += blockDim.x) {
      const int o = ti.o0 + col;
      uint2 m;
      m.x = (o < ti.o1) ? P.obs_lp[o] : 0xffffu;
      m.y = (o < ti.o1) ? (unsigned)P.obs_point[o] : 0xffffffffu;
      meta[col] = m;
    }
    return;
  }
  // short tile (blockDim.x = 128 = one warp per item): one meta entry per FREE-pose observation, in landmark order
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int cnt = tile_item_cnt(ti, wid), start = tile_item_start(ti, wid);
  const bool act = wid < ti.nitem && lane < cnt;
  const unsigned lp = act ? P.obs_lp[start + lane] : 0xffffu;
  const bool has = (lp & 0xffffu) != 0xffffu;
  const int fcol = tile_fcol(ti, wid, has, lane);
  if (has) meta[fcol] = make_uint2(lp, (unsigned)P.obs_point[start + lane]);
  if (threadIdx.x == 0 && (ti.nfree & 1)) meta[ti.nfree] = make_uint2(0xffffu, 0xffffffffu);  // pad column
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Small-window matvec, warp-specialised and persistent.  One producer warp streams whole JQ blocks (header + 27 planes
// + meta) of the CTA's contiguous tile range into an S-stage shared-memory ring with ONE TMA bulk copy per tile
// (mbarrier complete_tx); four consumer warps wait on the stage, read everything from shared memory (no dependent
// global loads on the critical path), and hand the stage back through an "empty" mbarrier.  p and the q accumulators
// of the current window stay in shared memory; q is flushed with one atomic per (CTA, window, component).
//
// BIG = false: every window has <= MAXSLOT free poses; p and the q accumulators of the current window live in shared
// memory (maxslot = free poses of the largest window).
// BIG = true (global BA): p is gathered from global memory (L1/L2 resident: 48 bytes per pose) and the q accumulators
// are a sliding WINDOW of `maxslot` consecutive slots anchored at the lowest slot of the tile being processed.  The
// host orders the landmarks of a big window by their first free pose, so the consecutive tiles of one CTA touch a
// narrow, slowly drifting band of poses; run sums that fall outside the window go straight to global atomics.
constexpr int PIPE_THREADS = CTA + 32;
constexpr int JQ_STAGE_D = JQ_HDR + JQ_ROWS * CTA + (2 * CTA + 4) / 2;
template <int S, bool BIG>
__global__ void __maxnreg__(96) k_matvec_pipe(Dev P, const double* __restrict__ pvec,
                                                               double* __restrict__ qvec, int force_all, int maxslot) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int STAGE_D = JQ_STAGE_D;  // doubles per stage: header + 28 rows + run table of a full tile
  double* stage = reinterpret_cast<double*>(smem_raw);
  constexpr int CST = CTA + 1;  // padded row stride of c_sh: the 6 value-lanes of a run hit 6 different bank pairs
  double* c_sh = stage + (size_t)S * STAGE_D;  // two buffers of [6][CST]
  double* p_sh = c_sh + 2 * (6 * CST) + 2;      // component-major: p_sh[c * maxslot + slot]  (unused when BIG)
  double* acc_sh = p_sh + (BIG ? 0 : 6 * maxslot);
  double* cam_sh = acc_sh + 6 * maxslot;                         // !BIG: fx, fy, bf of the window's slots [c * maxslot + slot]
  int* runs_sh = reinterpret_cast<int*>(cam_sh + (BIG ? 0 : 3 * maxslot));  // two buffers of 2*CTA+4 ints
  uint64_t* full = reinterpret_cast<uint64_t*>(runs_sh + 2 * (2 * CTA + 4));
  uint64_t* empty = full + S;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t0 = (int)((long long)P.n_tile * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)P.n_tile * (blockIdx.x + 1) / gridDim.x);
  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_fence_init();
  }
  __syncthreads();
  if (wid == WARPS) {
    // ------------------------------------------------------------------ producer warp (one elected lane)
    if (lane == 0) {
      int n = 0, seen_win = -1;
      bool win_on = true;
      TileInfo nxt = P.tiles[t0 < t1 ? t0 : 0];
      for (int t = t0; t < t1; t++) {
        const TileInfo ti = nxt;
        if (t + 1 < t1) nxt = P.tiles[t + 1];  // prefetch the next descriptor before blocking on the ring
        if (ti.win != seen_win) {
          seen_win = ti.win;
          win_on = force_all || P.ctl[ti.win].cg_active;
        }
        if (!win_on) continue;
        const int s = n % S;
        if (n >= S) mbar_wait(&empty[s], (uint32_t)(((n / S) - 1) & 1));
        const uint32_t bytes = ti.is_long ? (uint32_t)(JQ_HDR * sizeof(double))
                                          : (uint32_t)(ti.blk_doubles * sizeof(double));
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(stage + (size_t)s * STAGE_D, P.JQ + ti.jq_off - JQ_HDR, bytes, &full[s]);
        n++;
      }
      const int s = n % S;  // end-of-stream marker
      if (n >= S) mbar_wait(&empty[s], (uint32_t)(((n / S) - 1) & 1));
      reinterpret_cast<volatile int*>(stage + (size_t)s * STAGE_D)[0] = -1;
      mbar_arrive(&full[s]);
    }
    return;
  }
  // -------------------------------------------------------------------- consumer warps
  int n = 0, cur_win = -1, ws0 = 0, wn = 0;
  int abase = 0;  // BIG: first (window-relative) slot of the shared accumulator window
#ifdef SQRTBA_PIPE_PROF
  long long tp[6] = {0, 0, 0, 0, 0, 0}, tc = clock64(), tn;
#define PROF_MARK(i) { tn = clock64(); tp[i] += tn - tc; tc = tn; }
#else
#define PROF_MARK(i)
#endif
  while (true) {
    const int s = n % S;
    PROF_MARK(5)
    mbar_wait(&full[s], (uint32_t)((n / S) & 1));
    PROF_MARK(0)
    const double* st = stage + (size_t)s * STAGE_D;
    const int* hdr = reinterpret_cast<const int*>(st);
    const int nitem = reinterpret_cast<const volatile int*>(hdr)[0];
    if (nitem < 0) break;
    if (BIG) {
      const int lo = hdr[15];
      const bool new_win = hdr[1] != cur_win;
      if (new_win || (maxslot > 0 && lo >= 0 && (lo < abase || lo >= abase + (maxslot >> 1)))) {
        // flush the accumulator window and re-anchor it at this tile's lowest slot
        named_bar_sync(1, CTA);  // previous tile's reduction complete
        const int aw = min(maxslot, wn - abase);
        for (int i = tid; i < aw * 6; i += CTA) {
          const double v = acc_sh[i];
          if (v != 0.0) atomicAdd(&qvec[(size_t)(ws0 + abase) * 6 + i], v);
        }
        cur_win = hdr[1];
        ws0 = hdr[2];
        wn = hdr[3];
        abase = max(0, min(lo, wn - maxslot));
        named_bar_sync(1, CTA);
        for (int i = tid; i < maxslot * 6; i += CTA) acc_sh[i] = 0.0;
        named_bar_sync(1, CTA);
      }
    } else if (hdr[1] != cur_win) {
      // flush the finished window's accumulators, stage p of the new one
      named_bar_sync(1, CTA);  // previous tile's reduction complete
      if (P.det) {  // reproducible mode: one partial vector per (CTA, window) visit, summed in order by k_cg_step
        if (cur_win >= 0)
          for (int i = tid; i < wn * 6; i += CTA) P.part_q[(size_t)(blockIdx.x + cur_win) * 6 * P.maxslot + i] = acc_sh[i];
      } else {
        for (int i = tid; i < wn * 6; i += CTA) atomicAdd(&qvec[(size_t)ws0 * 6 + i], acc_sh[i]);
      }
      cur_win = hdr[1];
      ws0 = hdr[2];
      wn = hdr[3];
      named_bar_sync(1, CTA);
      for (int i = tid; i < wn * 6; i += CTA) {
        acc_sh[i] = 0.0;
        const int sl = i / 6;
        p_sh[(i - sl * 6) * maxslot + sl] = pvec[(size_t)ws0 * 6 + i];
      }
      for (int i = tid; i < wn * 3; i += CTA) {
        const int sl = i / 3;
        cam_sh[(i - sl * 3) * maxslot + sl] = __ldg(P.slot_cam + (size_t)ws0 * 3 + i);
      }
      named_bar_sync(1, CTA);
    }
    const int nt = hdr[5];
    if (hdr[13]) {  // a long landmark is a tile of its own: operands straight from global memory, direct atomics
      const long long jq_off = ((long long)hdr[8] << 32) | (unsigned)hdr[7];
      if (wid == 0) matvec_long_item(P, P.JQ + jq_off, nt, pvec, qvec, hdr[6], hdr[9], lane);
      named_bar_sync(1, CTA);
      if (tid == 0) mbar_arrive(&empty[s]);
      n++;
      continue;
    }
    const double* data = st + JQ_HDR;
    const uint2* meta = reinterpret_cast<const uint2*>(data + (size_t)NPLANE * nt);
    double* cb = c_sh + (size_t)(n & 1) * (6 * CST);       // double-buffered: no barrier needed after the reduction
    int* rb = runs_sh + (size_t)(n & 1) * (2 * CTA + 4);
    const int nrun = hdr[14];
    {  // copy the tile's run table out of the stage so the stage can be released right after the first barrier
      const int* rsrc = reinterpret_cast<const int*>(data + (size_t)JQ_ROWS * nt);
      for (int i = tid; i < 2 * nrun + 1; i += CTA) rb[i] = rsrc[i];
    }
    if (wid < nitem) {
      const int col0 = (wid > 0 ? hdr[9] : 0) + (wid > 1 ? hdr[10] : 0) + (wid > 2 ? hdr[11] : 0);
      const int cnt = hdr[9 + wid];
      const bool act = lane < cnt;
      const int col = col0 + (act ? lane : 0);
      int lm = -1 - lane, rank = 0, ls = 0;
      bool has = false, stereo = false;
      if (act) {
        const uint2 m = meta[col];
        ls = (int)(m.x & 0xffffu);
        has = ls != 0xffff;
        rank = (int)((m.x >> 16) & 0x7fffu);
        stereo = (m.x & LP_STEREO) != 0;
        lm = (int)m.y;
      }
      const Seg sg = seg_of(lm, lane);
      const double* dcol = data + col;
      double v[3] = {0, 0, 0}, t[3] = {0, 0, 0};
      JpC Jc;
      if (has) {
        double pp[6], g[JG], cm[3];
        if (BIG) {
          const double* pg = pvec + (size_t)(ws0 + ls) * 6;
#pragma unroll
          for (int c = 0; c < 6; c++) pp[c] = __ldg(pg + c);
#pragma unroll
          for (int c = 0; c < 3; c++) cm[c] = __ldg(P.slot_cam + (size_t)(ws0 + ls) * 3 + c);
        } else {
#pragma unroll
          for (int c = 0; c < 6; c++) pp[c] = p_sh[c * maxslot + ls];
#pragma unroll
          for (int c = 0; c < 3; c++) cm[c] = cam_sh[c * maxslot + ls];
        }
#pragma unroll
        for (int c = 0; c < JG; c++) g[c] = dcol[c * nt];
        Jc = jp_compact(g, cm[0], cm[1], cm[2], stereo);  // the weighted 3x6 pose block, 13 distinct entries
        jp_mul(Jc, stereo, pp, v);
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
          for (int k = 0; k < 3; k++) t[k] += dcol[(JG + r * 3 + k) * nt] * v[r];
        }
      }
      double sv[3];
#pragma unroll
      for (int k = 0; k < 3; k++) sv[k] = seg_sum(t[k], sg, lane);
      if (has) {
        // Q1 rows are re-read from the stage instead of being kept live across the shuffles (register pressure)
#pragma unroll
        for (int r = 0; r < 3; r++)
          v[r] -= dcol[(JG + r * 3) * nt] * sv[0] + dcol[(JG + 1 + r * 3) * nt] * sv[1] + dcol[(JG + 2 + r * 3) * nt] * sv[2];
        double out[6];
        jp_mulT(Jc, stereo, v, out);
#pragma unroll
        for (int c = 0; c < 6; c++) cb[c * CST + rank] = out[c];
      }
    }
    PROF_MARK(1)
    named_bar_sync(1, CTA);  // every consumer has drained stage s and published its column / the run table
    if (tid == 0) mbar_arrive(&empty[s]);
    PROF_MARK(2)
    // pose-side reduction: one thread per (run of equal slot, component); window accumulators stay in shared memory.
    // The next tile's barrier orders these read-modify-writes against the following tile's, so no barrier here.
    for (int idx = tid; idx < nrun * 6; idx += CTA) {
      const int r = (idx * 10923) >> 16, k = idx - r * 6;
      const int a = rb[r], b = rb[r + 1];
      const double sum = run_sum(cb + k * CST, a, b);
      if (BIG) {
        const int sl = rb[nrun + 1 + r], rel = sl - abase;
        if ((unsigned)rel < (unsigned)maxslot) acc_sh[rel * 6 + k] += sum;
        else atomicAdd(&qvec[(size_t)(ws0 + sl) * 6 + k], sum);
      } else {
        acc_sh[rb[nrun + 1 + r] * 6 + k] += sum;
      }
    }
    PROF_MARK(3)
    n++;
  }
  named_bar_sync(1, CTA);  // the last tile's reduction is complete before the accumulators are flushed
  if (BIG) {
    const int aw = min(maxslot, wn - abase);
    for (int i = tid; i < aw * 6; i += CTA) {
      const double v = acc_sh[i];
      if (v != 0.0) atomicAdd(&qvec[(size_t)(ws0 + abase) * 6 + i], v);
    }
  } else if (P.det) {
    if (cur_win >= 0)
      for (int i = tid; i < wn * 6; i += CTA) P.part_q[(size_t)(blockIdx.x + cur_win) * 6 * P.maxslot + i] = acc_sh[i];
  } else {
    for (int i = tid; i < wn * 6; i += CTA) atomicAdd(&qvec[(size_t)ws0 * 6 + i], acc_sh[i]);
  }
#ifdef SQRTBA_PIPE_PROF
  if (P.prof && (tid & 31) == 0) {
    long long* o = P.prof + ((size_t)blockIdx.x * WARPS + wid) * 8;
    for (int i = 0; i < 6; i++) o[i] = tp[i];
    o[6] = n;
  }
#endif
}

// ------------------------------------------------------------------------------------------------ K3-K5 fused: persistent PCG
// Whole preconditioned-CG solve of ONE window (local BA window, large window, global BA) in a single cooperative launch:
// the TMA-pipelined matvec above plus the CG vector updates, separated by grid barriers, so a CG iteration costs no
// kernel launch and no host round trip.  A grid barrier is ~1.7 us on 444 CTAs (tools/microbench/gridbar.cu) and every
// dependent L2 access ~0.5 us, so the design minimises barriers per iteration:
//
// BIG = false (<= MAXSLOT free poses: local BA): every CTA keeps the WHOLE CG state (p, res) replicated in shared
//   memory and performs the vector updates redundantly with identical arithmetic -> ONE grid barrier per iteration
//   (q complete).  q rotates through three buffers so that clearing never races with accumulation.
// BIG = true (global BA): vector phases are chunked by pose (chunk = VSLOT poses, one CTA per chunk, six lanes per
//   pose) -> THREE barriers per iteration: q complete | dot products | p complete.  The second dot product of textbook
//   CG is taken from the recurrence  r'.z' = r.z - 2 alpha q.z + alpha^2 q.Dinv q, evaluated together with p.q and
//   with the TRUE r.z of the current iterate (so the recurrence error never accumulates).
//   With landmarks sharded over several GPUs the pose-sized partial products are exchanged INSIDE
//   the kernel over NVLink peer memory: every rank stores its chunk of q into the receive buffer of every peer, raises
//   a per-chunk flag (st.release.sys), waits for the peers' flags and adds the partial vectors in rank order -- all
//   ranks hold bit-identical q, alpha, beta and iterates.  Dot products are per-chunk partials summed in chunk order
//   by every CTA, so they do not depend on the grid size (ranks with different shards launch different grids).
constexpr int VSLOT = 20;  // 4 warps x 5 poses x 6 lanes
// CHUNK (global BA, persistent kernel): the preconditioner is block-Jacobi with ONE block per chunk of VSLOT consecutive
// pose slots (120 x 120) instead of one per pose (6 x 6): the landmarks of a big window are ordered by their first pose,
// so most of the coupling of a pose is with its neighbours in the same chunk -- 2.2-2.4x fewer CG iterations on the
// KITTI-shaped loop (sqrtba_chunkprec.cuh builds and inverts the blocks).  The owner CTA of a chunk keeps the inverse in
// shared memory for the whole solve: strictly-lower triangle, packed, FP32 (28.6 KB; entrywise rounding of an SPD
// inverse whose Jacobi-scaled condition number is a few hundred keeps it SPD), diagonal in a register.
constexpr int CHB = VSLOT * 6;
constexpr int CH_PACK = CHB * (CHB - 1) / 2;
#ifndef PERSIST_REREAD
#define PERSIST_REREAD false
#endif
// !BIG: the per-CTA partial products are added into one of KQ copies of q (CTA b -> copy b % KQ).  FP64 atomics on
// one L2 line serialise (~30 cycles each: 444 CTAs flushing the same 120 doubles cost ~8 us), so the copies cut the
// depth of that queue; every CTA then adds the KQ copies in a fixed order.
constexpr int KQ = 4;

struct PcgArgs {
  double tol2;
  int max_iters;
  int nranks, rank;
  unsigned* gbar;                 // grid barrier arrival counter (monotonic; zeroed by the host before the launch)
  const int* tile_ptr;            // [gridDim.x + 1] cost-balanced tile ranges of the CTAs
  double* part;                   // BIG: [nchunk][4] partial dot products
  double* q3;                     // !BIG: three rotating buffers of KQ copies of q, [3][KQ][6*n_slot] (zeroed by the host)
  double* dq;                     // BIG: Dinv * qf   (zeroed by the host)
  double* qf;                     // BIG: q + lambda p (zeroed by the host)
  uint4* recv;                    // this rank's receive buffer  [2 (parity)][nranks][nelem_cap] of 16-byte records
                                  // {value lo, seq, value hi, seq}
  uint4* const* peer_tbl;         // device table of the receive buffers of all ranks, peer-mapped (cudaIpc)
  unsigned long long* seq_state;  // running exchange sequence number (device resident, advanced by the kernel)
  int nelem_cap;
  const float* cpack;             // CHUNK: [nchunk][CH_PACK] strictly-lower triangle of every chunk's inverse block (FP32)
  const float* cdiag;             // CHUNK: [nchunk][CHB] its diagonal
  const double* aci;              // CHUNK, optional: inverse of the coarse matrix Z^T S Z, [6 nchunk][6 nchunk] (Z = one
                                  // column per chunk and tangent component: the second level of the preconditioner)
  double* wvec;                   // CHUNK + coarse: Z^T qf of the current iteration, [6 nchunk]
};

__device__ __forceinline__ void st_volatile_v4(uint4* p, const uint4& v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Sum of one element of q over the ranks that share the problem, over NVLink peer memory, flag-in-data ("LL") style:
// every 8-byte half-record {4 bytes of the value, 4-byte sequence number} is written with one store (8-byte stores are
// not torn), so the receiver needs neither a fence nor a separate flag: it polls the record until both sequence
// numbers match.  The partial values are added in rank order: identical bits on every rank.
// (__noinline__: its registers -- eight records in flight -- stay out of the matvec loop's allocation.)
__device__ __forceinline__ double peer_sum(uint4* const* peer_tbl, const uint4* recv, int nranks, int rank, int nelem_cap,
                                        double qv, int e, unsigned long long seq) {
  const int par = (int)(seq & 1ull);
  const unsigned sq = (unsigned)seq;
  {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(qv);
    const uint4 rec = make_uint4((unsigned)bits, sq, (unsigned)(bits >> 32), sq);
#pragma unroll
    for (int r = 0; r < 8; r++)
      if (r < nranks && r != rank) st_volatile_v4(peer_tbl[r] + ((size_t)par * nranks + rank) * nelem_cap + e, rec);
  }
  const uint4* mine = recv + (size_t)par * nranks * nelem_cap + e;
  // four peers' records in flight at a time (two L2 round trips at eight ranks when the data is already there instead
  // of four; eight at a time cost the matvec loop registers), re-polling only the ones that have not arrived; the sum
  // itself is taken in rank order
  double tot = 0.0;
#pragma unroll 1
  for (int r0 = 0; r0 < nranks; r0 += 4) {
    uint4 v[4];
    unsigned pending = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      v[j] = make_uint4(0, sq, 0, sq);
      if (r0 + j < nranks && r0 + j != rank) pending |= 1u << j;
    }
    while (pending) {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (pending & (1u << j)) v[j] = ld_volatile_v4(mine + (size_t)(r0 + j) * nelem_cap);
#pragma unroll
      for (int j = 0; j < 4; j++)
        if ((pending & (1u << j)) && v[j].y == sq && v[j].w == sq) pending &= ~(1u << j);
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (r0 + j < nranks)
        tot += (r0 + j == rank) ? qv : __longlong_as_double((long long)(((unsigned long long)v[j].z << 32) | v[j].x));
  }
  return tot;
}

// Barrier over the consumer threads of every CTA of a cooperative launch: CTA barrier, one thread does a release-add
// on a monotonically increasing counter and acquire-spins until all CTAs of this generation arrived, CTA barrier.
// `gen` = how many barriers this launch has completed including this one (the counter is monotonic).
__device__ __forceinline__ void grid_bar(unsigned* bar, unsigned nblk, unsigned gen, int tid) {
  named_bar_sync(1, CTA);
  if (tid == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const unsigned target = gen * nblk;
    while (ld_acquire_gpu(bar) < target) {}
  }
  named_bar_sync(1, CTA);
}

// deterministic sum over the 128 consumer threads, result to all of them (slot: 0/1 selects the scratch cell)
__device__ __forceinline__ double cta_sum(double v, double* red_sh, int which, int lane, int wid) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(FULL, v, off);
  if (lane == 0) red_sh[which * 4 + wid] = v;
  named_bar_sync(1, CTA);
  return (red_sh[which * 4] + red_sh[which * 4 + 1]) + (red_sh[which * 4 + 2] + red_sh[which * 4 + 3]);
}

// per-lane matvec products of one staged tile: writes this observation's Jp^T u (6 values) into column `rank` of cb
// REREAD: Jp is re-read from the stage for the final product instead of being kept live across the shuffles -- slower
// (+25 % per tile) but 36 registers lighter; used by the multi-GPU instantiation, whose exchange code otherwise pushes
// ptxas into spilling those rows inside the loop (+80 % per tile).
template <bool BIG, bool REREAD = false>
__device__ __forceinline__ void tile_products(const Dev& P, const double* data, const int* hdr, int nt, int wid, int lane,
                                              const double* p_sh, int maxslot, const PEff& pe, int abase, double* cb) {
  constexpr int CST = CTA + 1;
  const uint2* meta = reinterpret_cast<const uint2*>(data + (size_t)NPLANE * nt);
  const int col0 = (wid > 0 ? hdr[9] : 0) + (wid > 1 ? hdr[10] : 0) + (wid > 2 ? hdr[11] : 0);
  const int cnt = hdr[9 + wid];
  const bool act = lane < cnt;
  const int col = col0 + (act ? lane : 0);
  int lm = -1 - lane, rank = 0, ls = 0;
  bool has = false, stereo = false;
  if (act) {
    const uint2 m = meta[col];
    ls = (int)(m.x & 0xffffu);
    has = ls != 0xffff;
    rank = (int)((m.x >> 16) & 0x7fffu);
    stereo = (m.x & LP_STEREO) != 0;
    lm = (int)m.y;
  }
  const Seg sg = seg_of(lm, lane);
  const double* dcol = data + col;
  double v[3] = {0, 0, 0}, t[3] = {0, 0, 0};
  JpC Jc;
  if (has) {
    double pp[6], g[JG], cm[3];
    const int rel = BIG ? ls - abase : ls;
    if (!BIG || (unsigned)rel < (unsigned)maxslot) {  // BIG: inside the CTA's window of the search direction
#pragma unroll
      for (int cc = 0; cc < 6; cc++) pp[cc] = p_sh[cc * maxslot + rel];
    } else {                                           // outside (loop-closure observations, wide covisibility)
#pragma unroll
      for (int cc = 0; cc < 6; cc++) pp[cc] = peff_at(pe, (size_t)ls * 6 + cc);
    }
    // one window per problem here: the window-relative slot is the global one (intrinsics: three L1-resident doubles)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) cm[cc] = __ldg(P.slot_cam + (size_t)ls * 3 + cc);
#pragma unroll
    for (int cc = 0; cc < JG; cc++) g[cc] = dcol[cc * nt];
    Jc = jp_compact(g, cm[0], cm[1], cm[2], stereo);
    jp_mul(Jc, stereo, pp, v);
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
      for (int k = 0; k < 3; k++) t[k] += dcol[(JG + r * 3 + k) * nt] * v[r];
    }
  }
  double sv[3];
#pragma unroll
  for (int k = 0; k < 3; k++) sv[k] = seg_sum(t[k], sg, lane);
  if (has) {
#pragma unroll
    for (int r = 0; r < 3; r++)
      v[r] -= dcol[(JG + r * 3) * nt] * sv[0] + dcol[(JG + 1 + r * 3) * nt] * sv[1] + dcol[(JG + 2 + r * 3) * nt] * sv[2];
    double out[6];
    jp_mulT(Jc, stereo, v, out);
#pragma unroll
    for (int cc = 0; cc < 6; cc++) cb[cc * CST + rank] = out[cc];
  }
}

__host__ __device__ inline int pipe_run_cap(int maxslot, bool big) {  // ints per run-table buffer
  return 2 * ((big || maxslot > CTA) ? CTA : maxslot) + 4;
}

// MULTI: landmark-sharded over several GPUs (compiled separately so that the exchange code cannot disturb the register
// allocation of the single-GPU matvec loop)
template <int S, bool BIG, bool MULTI = false, bool CHUNK = false>
__global__ void __maxnreg__(120) k_pcg_persist(Dev P, PcgArgs A, int maxslot, double lam_override, int use_override) {
  static_assert(!CHUNK || BIG, "the chunk preconditioner belongs to the big-window path");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int STAGE_D = JQ_STAGE_D;
  constexpr int CST = CTA + 1;
  double* stage = reinterpret_cast<double*>(smem_raw);
  double* c_sh = stage + (size_t)S * STAGE_D;            // two buffers of [6][CST]
  double* p_sh = c_sh + 2 * (6 * CST) + 2;                // p, component-major [c * maxslot + slot]; BIG: a WINDOW of
                                                          // maxslot consecutive slots starting at abase
  double* acc_sh = p_sh + 6 * maxslot;                    // !BIG: q accumulators, then scratch of the vector phase
  double* res_sh = acc_sh + (BIG ? 0 : 6 * maxslot);      // !BIG: residual, slot-major [slot * 6 + c]
  double* red_sh = res_sh + (BIG ? 0 : 6 * maxslot);      // 8 doubles of reduction scratch
  // CG scalars that live across iterations are parked here during the matvec phase instead of in registers: the hot
  // loop needs every register it can get (ptxas otherwise spills the Jacobian rows around the shuffles)
  volatile double* st_sh = red_sh + 8;                    // [0] r.z  [1] r0.z0  [2] lambda
  uint64_t* full = reinterpret_cast<uint64_t*>(red_sh + 16);
  uint64_t* empty = full + S;
  volatile int* sig = reinterpret_cast<volatile int*>(empty + S);  // [0] approved iteration, [1] stop, [2] fills consumed
  int* runs_sh = const_cast<int*>(sig) + 4;
  const int rcap = pipe_run_cap(maxslot, BIG);
  float* cpk_sh = reinterpret_cast<float*>(runs_sh + 2 * rcap);  // CHUNK: this CTA's chunk inverse, strictly lower, packed
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t0 = A.tile_ptr[blockIdx.x], t1 = A.tile_ptr[blockIdx.x + 1];
  const int ntile = t1 - t0;
  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_fence_init();
    sig[0] = 0;
    sig[1] = 0;
    sig[2] = 0;
  }
  __syncthreads();
  WinCtl& c = P.ctl[0];
  if (!c.cg_active) return;  // grid-uniform: set by k_cg_init before this launch, cleared only after the last barrier
  if (wid == WARPS) {
    // ------------------------------------------------------------------ producer warp (one elected lane)
    // JQ does not change during the solve, so the first tiles of the next iteration are prefetched speculatively while
    // the consumers run the vector phases; past the ring depth the iteration must have been approved.
    if (lane == 0 && ntile > 0) {
      int n = 0;
      TileInfo nxt = P.tiles[t0];
      for (int it = 0;; it++) {
        bool stop = false;
        for (int k = 0; k < ntile; k++) {
          if (it > 0 && k >= S) {
            while (sig[0] < it && !sig[1]) {}
            if (sig[1]) { stop = true; break; }
          }
          const TileInfo ti = nxt;
          nxt = P.tiles[(k + 1 < ntile) ? t0 + k + 1 : t0];
          const int s = n % S;
          if (n >= S) mbar_wait(&empty[s], (uint32_t)(((n / S) - 1) & 1));
          const uint32_t bytes = ti.is_long ? (uint32_t)(JQ_HDR * sizeof(double))
                                            : (uint32_t)(ti.blk_doubles * sizeof(double));
          mbar_expect_tx(&full[s], bytes);
          bulk_g2s(stage + (size_t)s * STAGE_D, P.JQ + ti.jq_off - JQ_HDR, bytes, &full[s]);
          n++;
        }
        if (!stop && it > 0 && ntile <= S) {  // every tile of this iteration was speculative: still needs the verdict
          while (sig[0] < it && !sig[1]) {}
          if (sig[1]) stop = true;
        }
        if (stop) break;
      }
      // drain: speculative copies of an iteration that never runs must land before the CTA's shared memory is released
      const int used = sig[2];
      for (int f = used; f < n; f++) mbar_wait(&full[f % S], (uint32_t)((f / S) & 1));
    }
    return;
  }
  // -------------------------------------------------------------------- consumer warps
  if (tid == 0) {
    st_sh[0] = c.rz;
    st_sh[1] = c.rz0;
    st_sh[2] = use_override ? lam_override : c.lambda;
    st_sh[3] = A.tol2 > 0.0 ? A.tol2 : (c.tol2 > 0.0 ? c.tol2 : 1e-18);  // tolerance of this pass (WinCtl::tol2)
  }
  const int n6 = P.n_slot * 6;
  int n = 0, abase = 0;
  double rz_out = 0.0;
  int iters_out = 0, exch_out = 0;
  PEff pe;
  pe.z = P.z; pe.dq = A.dq; pe.p = P.p;
  pe.alpha = 0.0; pe.beta = 0.0;  // iteration 0: p = z (k_cg_init), Dq zeroed by the host
#ifdef SQRTBA_PIPE_PROF
  long long tpp[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tcc = clock64(), tnn;
#define PROFP(i) { tnn = clock64(); tpp[i] += tnn - tcc; tcc = tnn; }
#else
#define PROFP(i)
#endif
  if (CHUNK) {  // owner of a chunk: its inverse block stays in shared memory for the whole solve (the host guarantees
                // gridDim.x >= number of chunks, so a CTA owns at most one)
    if ((int)blockIdx.x * VSLOT < P.n_slot) {
      const float* src = A.cpack + (size_t)blockIdx.x * CH_PACK;
      for (int i = tid; i < CH_PACK; i += CTA) cpk_sh[i] = __ldg(src + i);
    }
  }
  if (!BIG) {  // replicate the CG state
    for (int e = tid; e < n6; e += CTA) {
      const int sl = e / 6;
      p_sh[(e - sl * 6) * maxslot + sl] = P.p[e];
      res_sh[e] = P.res[e];
      acc_sh[e] = 0.0;
    }
  }
  named_bar_sync(1, CTA);
  for (int it = 0;; it++) {
    double* qcur = BIG ? P.q : A.q3 + ((size_t)(it % 3) * KQ + (blockIdx.x % KQ)) * n6;
    PROFP(7)
    // ================================================================== matvec over this CTA's tiles
    for (int kt = 0; kt < ntile; kt++) {
      const int s = n % S;
      mbar_wait(&full[s], (uint32_t)((n / S) & 1));
      const double* st = stage + (size_t)s * STAGE_D;
      const int* hdr = reinterpret_cast<const int*>(st);
      const int nitem = hdr[0];
      const int nt = hdr[5];
      if (BIG) {
        // slide the window of the search direction with the (sorted) landmarks: re-evaluate it when this tile starts
        // below it or more than a few poses above its anchor
        const int lo = hdr[15];
        const int want = max(0, min(lo, P.n_slot - maxslot));
        if (kt == 0 || (lo >= 0 && (want < abase || want > abase + 6))) {
          named_bar_sync(1, CTA);  // every warp is done reading the old window
          abase = want;
          for (int i = tid; i < maxslot * 6; i += CTA) {
            const int rel = i / 6, cc = i - rel * 6;
            if (abase + rel < P.n_slot) p_sh[cc * maxslot + rel] = peff_at(pe, (size_t)(abase + rel) * 6 + cc);
          }
          named_bar_sync(1, CTA);
        }
      }
      if (hdr[13]) {  // long landmark: operands straight from global memory, direct atomics
        const long long jq_off = ((long long)hdr[8] << 32) | (unsigned)hdr[7];
        if (wid == 0) {
          if (BIG) matvec_long_item<3>(P, P.JQ + jq_off, nt, nullptr, qcur, hdr[6], hdr[9], lane, 0, &pe);
          else matvec_long_item<2>(P, P.JQ + jq_off, nt, p_sh, qcur, hdr[6], hdr[9], lane, maxslot);
        }
        named_bar_sync(1, CTA);
        if (tid == 0) mbar_arrive(&empty[s]);
        n++;
        continue;
      }
      const double* data = st + JQ_HDR;
      double* cb = c_sh + (size_t)(n & 1) * (6 * CST);
      int* rb = runs_sh + (size_t)(n & 1) * rcap;
      const int nrun = hdr[14];
      {
        const int* rsrc = reinterpret_cast<const int*>(data + (size_t)JQ_ROWS * nt);
        for (int i = tid; i < 2 * nrun + 1; i += CTA) rb[i] = rsrc[i];
      }
      if (wid < nitem) tile_products<BIG, MULTI && PERSIST_REREAD>(P, data, hdr, nt, wid, lane, p_sh, maxslot, pe, abase, cb);
      named_bar_sync(1, CTA);
      if (tid == 0) mbar_arrive(&empty[s]);
      for (int idx = tid; idx < nrun * 6; idx += CTA) {
        const int r = (idx * 10923) >> 16, k = idx - r * 6;
        const int a = rb[r], b = rb[r + 1];
        const double sum = run_sum(cb + k * CST, a, b);
        if (BIG) atomicAdd(&qcur[(size_t)rb[nrun + 1 + r] * 6 + k], sum);
        else acc_sh[rb[nrun + 1 + r] * 6 + k] += sum;
      }
      n++;
    }
    PROFP(3)
    if (!BIG) {
      named_bar_sync(1, CTA);
      for (int e = tid; e < n6; e += CTA) atomicAdd(&qcur[e], acc_sh[e]);
    }
    PROFP(4)
    // BIG: an owner CTA requests everything of its (first) chunk that does not depend on this iteration's matvec --
    // the vectors it wrote itself one iteration ago and its rows of the block-Jacobi inverse -- BEFORE the barrier, so
    // that after it only q itself is a dependent L2 round trip
    double pf_po = 0.0, pf_dq = 0.0, pf_qf = 0.0, pf_z = 0.0, pf_r = 0.0, pf_dv[6] = {0, 0, 0, 0, 0, 0};
    if (BIG) {
      const int slot0 = (int)blockIdx.x * VSLOT + wid * 5 + lane / 6;
      if ((int)blockIdx.x * VSLOT < P.n_slot && lane < 30 && slot0 < P.n_slot) {
        const int e0 = slot0 * 6 + (lane - (lane / 6) * 6);
        pf_po = P.p[e0]; pf_dq = A.dq[e0]; pf_qf = A.qf[e0]; pf_z = P.z[e0]; pf_r = P.res[e0];
        if (CHUNK) pf_dv[0] = (double)__ldg(A.cdiag + (size_t)blockIdx.x * CHB + wid * 30 + lane);
        else if (!MULTI) {  // sharded: the inverse's rows are requested right before the exchange and hide behind it
#pragma unroll
          for (int k = 0; k < 6; k++) pf_dv[k] = __ldg(&P.Dinv[(size_t)e0 * 6 + k]);
        }
      }
    }
    const bool coarse = CHUNK && A.aci != nullptr;  // grid-uniform
    const int nbar = BIG ? (coarse ? 3 : 2) : 1;     // grid barriers per iteration
    grid_bar(A.gbar, gridDim.x, (unsigned)(nbar * it + 1), tid);  // B1: q complete
    PROFP(0)
    const double lam = st_sh[2], rz0 = st_sh[1];
    double rz = st_sh[0];
    if (!BIG) {
      // ================================================================ replicated vector update (one barrier / iteration)
      if (blockIdx.x == 0) {  // clear the buffer of iteration it+2: its readers (iteration it-1) all passed B1 above
        double* qz = A.q3 + (size_t)((it + 2) % 3) * KQ * n6;
        for (int e = tid; e < KQ * n6; e += CTA) qz[e] = 0.0;
      }
      constexpr int EPT = MAXSLOT * 6 / CTA;  // elements per thread
      double d = 0.0;
      {
        const double* qall = A.q3 + (size_t)(it % 3) * KQ * n6;
        double qg[EPT][KQ];
#pragma unroll
        for (int j = 0; j < EPT; j++) {
          const int e = tid + j * CTA;
#pragma unroll
          for (int k = 0; k < KQ; k++) qg[j][k] = (e < n6) ? __ldcg(&qall[(size_t)k * n6 + e]) : 0.0;  // all loads in flight together
        }
#pragma unroll
        for (int j = 0; j < EPT; j++) {
          const int e = tid + j * CTA;
          if (e < n6) {
            const int sl = e / 6, cc = e - sl * 6;
            const double pv = p_sh[cc * maxslot + sl];
            double qs = qg[j][0];
#pragma unroll
            for (int k = 1; k < KQ; k++) qs += qg[j][k];
            const double qv = qs + lam * pv;
            acc_sh[e] = qv;
            d += pv * qv;
          }
        }
      }
      PROFP(1)
      const double pq = cta_sum(d, red_sh, 0, lane, wid);
      const double alpha = rz / pq;
      if (!(pq > 0.0) || !isfinite(alpha)) {  // breakdown: keep the iterate, LM judges the step by its gain ratio
        rz_out = rz;
        iters_out = it;
        break;
      }
      for (int e = tid; e < n6; e += CTA) {
        res_sh[e] -= alpha * acc_sh[e];
        if (blockIdx.x == 0) {  // the step itself is only needed once: fire-and-forget adds (x was zeroed by k_cg_init)
          const int sl = e / 6, cc = e - sl * 6;
          atomicAdd(&P.x[e], alpha * p_sh[cc * maxslot + sl]);
        }
      }
      named_bar_sync(1, CTA);
      d = 0.0;
      {
        double dg[EPT][6];
#pragma unroll
        for (int j = 0; j < EPT; j++) {
          const int e = tid + j * CTA;
          const double* Di = P.Dinv + (size_t)e * 6;  // row cc of block sl = 6 doubles at (sl*36 + cc*6) = e*6
#pragma unroll
          for (int k = 0; k < 6; k++) dg[j][k] = (e < n6) ? __ldg(Di + k) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < EPT; j++) {
          const int e = tid + j * CTA;
          if (e < n6) {
            const int sl = e / 6;
            const double* rs = res_sh + sl * 6;
            double z = 0.0;
#pragma unroll
            for (int k = 0; k < 6; k++) z += dg[j][k] * rs[k];
            acc_sh[e] = z;
            d += res_sh[e] * z;
          }
        }
      }
      PROFP(2)
      const double rzn = cta_sum(d, red_sh, 1, lane, wid);
      const bool last = !(rzn > st_sh[3] * rz0) || it + 1 >= A.max_iters;
      if (tid == 0) {
        if (last) { sig[2] = n; __threadfence_block(); sig[1] = 1; }
        else sig[0] = it + 1;
        st_sh[0] = rzn;  // read again after the next grid barrier
      }
      const double beta = rzn / rz;
      if (last) {
        rz_out = rzn;
        iters_out = it + 1;
        break;
      }
      for (int e = tid; e < n6; e += CTA) {
        const int sl = e / 6, cc = e - sl * 6;
        p_sh[cc * maxslot + sl] = acc_sh[e] + beta * p_sh[cc * maxslot + sl];
        acc_sh[e] = 0.0;
      }
      named_bar_sync(1, CTA);
      PROFP(5)
      continue;
    }
    // ================================================================== chunked vector update (BIG)
    // ---- owners: apply the PREVIOUS iteration's update (deferred so that nobody had to wait for it), then
    //      [cross-rank sum of q]  qf = q + lambda p,  Dq = Dinv qf,  partial r.z, p.qf, qf.z, qf.Dq;  q = 0
    const unsigned long long seq = (MULTI ? *A.seq_state : 0ull) + (unsigned long long)it + 1ull;
    const int nchunk = (P.n_slot + VSLOT - 1) / VSLOT;
    const int vsl = wid * 5 + lane / 6, vcc = lane - (lane / 6) * 6;
    const int vbase = min((lane / 6) * 6, 24);
    exch_out = it + 1;
    // partial dot products of one chunk: r.z, p.qf, qf.z, qf.Dq -> A.part[chunk]
    auto chunk_dots = [&](int ch, double rv, double zv, double pv, double qv, double dq) {
      double d0 = rv * zv, d1 = pv * qv, d2 = qv * zv, d3 = qv * dq;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        d0 += __shfl_down_sync(FULL, d0, off);
        d1 += __shfl_down_sync(FULL, d1, off);
        d2 += __shfl_down_sync(FULL, d2, off);
        d3 += __shfl_down_sync(FULL, d3, off);
      }
      PROFP(6)
      named_bar_sync(1, CTA);  // scratch free (previous chunk's partials consumed)
      PROFP(8)
      if (lane == 0) { red_sh[wid] = d0; red_sh[4 + wid] = d1; c_sh[wid] = d2; c_sh[4 + wid] = d3; }
      named_bar_sync(1, CTA);
      if (tid < 4) {
        const double* src = (tid < 2) ? red_sh + tid * 4 : c_sh + (tid - 2) * 4;
        A.part[(size_t)ch * 4 + tid] = (src[0] + src[1]) + (src[2] + src[3]);
      }
    };
    // two-level preconditioner: the owner's state between the chunk-local and the coarse part (one chunk per CTA)
    bool co_ok = false;
    int co_e = 0;
    double co_qv = 0.0, co_pv = 0.0, co_zv = 0.0, co_rv = 0.0, co_dq = 0.0;
    for (int ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
      const int slot = ch * VSLOT + vsl;
      const bool ok = lane < 30 && slot < P.n_slot;
      const int e = slot * 6 + vcc;
      double qv = 0.0, pv = 0.0, zv = 0.0, rv = 0.0;
      const bool first = ch == (int)blockIdx.x;  // this chunk's state was requested before the barrier
      double dv[6];
      if (ok) {
        qv = __ldcg(&P.q[e]);
        const double po = first ? pf_po : P.p[e], dqo = first ? pf_dq : A.dq[e], qfo = first ? pf_qf : A.qf[e];
        zv = (first ? pf_z : P.z[e]) - pe.alpha * dqo;
        pv = zv + pe.beta * po;
        rv = (first ? pf_r : P.res[e]) - pe.alpha * qfo;
        P.x[e] += pe.alpha * po;
        P.z[e] = zv;
        P.p[e] = pv;
        P.res[e] = rv;
        P.q[e] = 0.0;
      }
      // this thread's row of the block-Jacobi inverse (later chunks of a CTA: requested before the exchange so that its
      // latency hides behind the NVLink round trip)
      if (!CHUNK) {
#pragma unroll
        for (int k = 0; k < 6; k++) dv[k] = (first && !MULTI) ? pf_dv[k] : (ok ? __ldg(&P.Dinv[(size_t)slot * 36 + vcc * 6 + k]) : 0.0);
      }
      PROFP(9)
      if (MULTI && ok) qv = peer_sum(A.peer_tbl, A.recv, A.nranks, A.rank, A.nelem_cap, qv, e, seq);
      PROFP(10)
      if (ok) {
        qv += lam * pv;
        A.qf[e] = qv;
      }
      double dq = 0.0;
      if (CHUNK) {
        // dq = (chunk inverse) qf: the chunk's 120 values through shared memory (the tile buffers are idle in this
        // phase), every thread one row of the symmetric block: left of the diagonal its own packed row, below it a
        // column walk (consecutive threads read consecutive words)
        double* y_sh = c_sh + 16;
        if (lane < 30) y_sh[wid * 30 + lane] = ok ? qv : 0.0;
        named_bar_sync(1, CTA);
        if (ok) {
          const int r = wid * 30 + lane;
          const float* row = cpk_sh + (r * (r - 1)) / 2;
          double a0 = pf_dv[0] * qv, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          int k = 0;
          for (; k + 3 < r; k += 4) {
            a0 += (double)row[k] * y_sh[k];
            a1 += (double)row[k + 1] * y_sh[k + 1];
            a2 += (double)row[k + 2] * y_sh[k + 2];
            a3 += (double)row[k + 3] * y_sh[k + 3];
          }
          for (; k < r; k++) a0 += (double)row[k] * y_sh[k];
          int idx = ((r + 1) * r) / 2 + r;  // entry (r + 1, r)
          for (k = r + 1; k + 1 < CHB; k += 2) {
            a1 += (double)cpk_sh[idx] * y_sh[k];
            a2 += (double)cpk_sh[idx + k] * y_sh[k + 1];
            idx += 2 * k + 1;
          }
          if (k < CHB) a3 += (double)cpk_sh[idx] * y_sh[k];
          dq = (a0 + a1) + (a2 + a3);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 6; k++) {
          const double qk = __shfl_sync(FULL, qv, vbase + k);
          dq += dv[k] * qk;
        }
      }
      if (coarse) {
        // second level: publish Z^T qf of this chunk (the chunk's qf is in y_sh), finish after the extra grid barrier
        if (tid < 6) {
          const double* y_sh = c_sh + 16;
          double w = 0.0;
#pragma unroll
          for (int sl = 0; sl < VSLOT; sl++) w += y_sh[sl * 6 + tid];
          A.wvec[(size_t)ch * 6 + tid] = w;
        }
        co_ok = ok; co_e = e; co_qv = qv; co_pv = pv; co_zv = zv; co_rv = rv; co_dq = dq;
        break;  // CHUNK: at most one chunk per CTA
      }
      if (ok) A.dq[e] = dq;
      PROFP(11)
      chunk_dots(ch, rv, zv, pv, qv, dq);
    }
    if (coarse) {
      grid_bar(A.gbar, gridDim.x, (unsigned)(nbar * it + 2), tid);  // Bx: Z^T qf complete
      const int ch = blockIdx.x;
      if (ch < nchunk) {
        // y = rows [6 ch, 6 ch + 6) of (Z^T S Z)^-1 times Z^T qf; every element of the chunk gets its component of y
        const int nc6 = nchunk * 6;
        double acc[6] = {0, 0, 0, 0, 0, 0};
        const double* arow = A.aci + (size_t)ch * 6 * nc6;
        for (int k = tid; k < nc6; k += CTA) {
          const double wk = __ldcg(&A.wvec[k]);
#pragma unroll
          for (int a = 0; a < 6; a++) acc[a] += __ldg(arow + (size_t)a * nc6 + k) * wk;
        }
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) acc[a] += __shfl_down_sync(FULL, acc[a], off);
        }
        double* yc_sh = c_sh + 16 + CHB;  // [4 warps][6]
        if (lane == 0) {
#pragma unroll
          for (int a = 0; a < 6; a++) yc_sh[wid * 6 + a] = acc[a];
        }
        named_bar_sync(1, CTA);
        const double yc = (yc_sh[vcc] + yc_sh[6 + vcc]) + (yc_sh[12 + vcc] + yc_sh[18 + vcc]);
        const double dq = co_dq + yc;
        if (co_ok) A.dq[co_e] = dq;
        chunk_dots(ch, co_rv, co_zv, co_pv, co_qv, co_ok ? dq : 0.0);
      }
    }
    PROFP(1)
    grid_bar(A.gbar, gridDim.x, (unsigned)(nbar * it + nbar), tid);  // B2: state of this iterate and the partial dot products complete
    PROFP(2)
    if (wid == 0) {  // one warp per CTA reads the partials (every CTA reads the same few lines)
      // three records per lane in flight at a time (96 chunks = 1920 poses per L2 round trip instead of one round trip
      // per record); the summation order is unchanged, so every rank still gets the same bits
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (int i0 = lane; i0 < nchunk; i0 += 96) {
        double2 ab[3], cd[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const int i = i0 + 32 * j;
          if (i < nchunk) {
            ab[j] = __ldcg(reinterpret_cast<const double2*>(A.part + (size_t)i * 4));
            cd[j] = __ldcg(reinterpret_cast<const double2*>(A.part + (size_t)i * 4 + 2));
          } else {
            ab[j] = make_double2(0.0, 0.0);
            cd[j] = make_double2(0.0, 0.0);
          }
        }
#pragma unroll
        for (int j = 0; j < 3; j++)
          if (i0 + 32 * j < nchunk) { s0 += ab[j].x; s1 += ab[j].y; s2 += cd[j].x; s3 += cd[j].y; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        s0 += __shfl_down_sync(FULL, s0, off);
        s1 += __shfl_down_sync(FULL, s1, off);
        s2 += __shfl_down_sync(FULL, s2, off);
        s3 += __shfl_down_sync(FULL, s3, off);
      }
      if (lane == 0) { red_sh[0] = s0; red_sh[1] = s1; red_sh[2] = s2; red_sh[3] = s3; }
    }
    named_bar_sync(1, CTA);
    const double s0 = red_sh[0], s1 = red_sh[1], s2 = red_sh[2], s3 = red_sh[3];
    rz = s0;  // true r.z of the current iterate
    const double alpha = rz / s1;
    pe.alpha = 0.0;
    pe.beta = 0.0;
    if (!(s1 > 0.0) || !isfinite(alpha)) {  // breakdown: keep the iterate, LM judges the step by its gain ratio
      rz_out = rz;
      iters_out = it;
      break;
    }
    double rzn = rz - 2.0 * alpha * s2 + alpha * alpha * s3;  // r'.z' after the step (exact in exact arithmetic)
    if (!(rzn > 0.0)) rzn = 0.0;
    const bool last = !(rzn > st_sh[3] * rz0) || it + 1 >= A.max_iters;
    if (tid == 0) {
      if (last) { sig[2] = n; __threadfence_block(); sig[1] = 1; }
      else sig[0] = it + 1;
      st_sh[0] = rzn;
    }
    pe.alpha = alpha;
    pe.beta = rzn / rz;
    PROFP(5)
    if (last) {
      rz_out = rzn;
      iters_out = it + 1;
      break;
    }
  }
  if (BIG && pe.alpha != 0.0) {  // the last step is still pending
    const int nchunk = (P.n_slot + VSLOT - 1) / VSLOT;
    const int vsl = wid * 5 + lane / 6, vcc = lane - (lane / 6) * 6;
    for (int ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
      const int slot = ch * VSLOT + vsl;
      if (lane < 30 && slot < P.n_slot) {
        const int e = slot * 6 + vcc;
        P.x[e] += pe.alpha * P.p[e];
      }
    }
  }
#ifdef SQRTBA_PIPE_PROF
  if (P.prof && tid == 0) {
    long long* o = P.prof + (size_t)blockIdx.x * 16;
    for (int i = 0; i < 8; i++) o[i] += tpp[i];
    o[8] += iters_out;
    o[9] += tpp[9]; o[10] += tpp[10]; o[11] += tpp[11]; o[12] += tpp[6]; o[13] += tpp[8];
  }
#endif
  if (tid == 0 && !sig[1]) { sig[2] = n; __threadfence_block(); sig[1] = 1; }  // breakdown exit: release the producer
  if (blockIdx.x == 0 && tid == 0) {
    c.rz = rz_out;
    c.cg_iters = iters_out;
    c.cg_active = 0;
    atomicAdd(&P.counters[2], iters_out);
    if (BIG && MULTI) *A.seq_state += (unsigned long long)exch_out;
  }
}

// ------------------------------------------------------------------------------------------------ K4/K5: PCG vector ops
// One CTA owns one window's pose-sized vectors, so dot products are block reductions with no global sync.
// no_dinv: the chunk preconditioner (sqrtba_chunkprec.cuh) writes z, p and r.z itself afterwards -- the 6x6 inverses were
// not computed, so start from z = b
__device__ __forceinline__ void cg_init_body(const Dev& P, int win, int force_all, double* sh, bool no_dinv = false) {
  WinCtl& c = P.ctl[win];
  if (!force_all && c.phase != PH_TRIAL) return;
  const int s0 = P.win_slot_ptr[win], s1 = P.win_slot_ptr[win + 1];
  double rz = 0.0;
  for (int e = s0 * 6 + threadIdx.x; e < s1 * 6; e += RCTA) {
    const int s = e / 6, rr = e - s * 6;
    double z = 0.0;
    if (no_dinv) z = P.bs[e];
    else {
#pragma unroll
      for (int k = 0; k < 6; k++) z += P.Dinv[s * 36 + rr * 6 + k] * P.bs[s * 6 + k];
    }
    const double b = P.bs[e];
    P.x[e] = 0.0;
    P.res[e] = b;
    P.z[e] = z;
    P.p[e] = z;
    rz += b * z;
  }
  rz = block_sum(rz, sh);
  if (threadIdx.x == 0) {
    c.rz = rz;
    c.rz0 = rz;
    c.cg_iters = 0;
    c.cg_active = (rz > 0.0) ? 1 : 0;
    if (c.cg_active) atomicAdd(&P.counters[1], 1);
  }
}
__global__ void __launch_bounds__(RCTA) k_cg_init(Dev P, int force_all, int no_dinv) {
  __shared__ double sh[RWARPS];
  cg_init_body(P, blockIdx.x, force_all, sh, no_dinv != 0);
}

// Single-window fusion of everything between the landmark QR and the persistent PCG kernel (CUDA-graph macro step):
// block-Jacobi inverses, CG start vectors, and the zeroing the host used to enqueue as memset nodes (q buffers of the
// persistent kernel, its grid-barrier counter, the "PCG active" counter).  One CTA (the problem has one window).
__global__ void __launch_bounds__(RCTA) k_cg_prep(Dev P, double* zero_a, int n_a, double* zero_b, int n_b, unsigned* gbar, int no_dinv) {
  __shared__ double sh[RWARPS];
  const WinCtl& c = P.ctl[0];
  if (threadIdx.x == 0) { P.counters[1] = 0; gbar[0] = 0u; gbar[1] = 0u; }
  for (int i = threadIdx.x; i < n_a; i += RCTA) zero_a[i] = 0.0;
  for (int i = threadIdx.x; i < n_b; i += RCTA) zero_b[i] = 0.0;
  if (c.phase != PH_TRIAL) return;
  const double lam = c.lambda;
  if (!no_dinv) {
    for (int s = threadIdx.x; s < P.n_slot; s += RCTA) dinv_slot(P, s, lam);
  }
  __syncthreads();
  cg_init_body(P, 0, 0, sh, no_dinv != 0);
}

__global__ void __launch_bounds__(RCTA) k_cg_step(Dev P, double tol2, int max_iters, int force_all, double lam_override,
                                                  int det_grid) {
  __shared__ double sh[RWARPS];
  const int win = blockIdx.x;
  WinCtl& c = P.ctl[win];
  if (!c.cg_active) return;
  const double lam = force_all ? lam_override : c.lambda;
  if (!(tol2 > 0.0)) tol2 = c.tol2 > 0.0 ? c.tol2 : 1e-18;  // tolerance of this pass
  const int e0 = P.win_slot_ptr[win] * 6, e1 = P.win_slot_ptr[win + 1] * 6;
  // reproducible mode (det_grid = grid of the matvec launch): q of the window = the partial vectors of the matvec CTAs
  // that own its tiles, in CTA order.  CTA c owns tiles [n_tile c / G, n_tile (c+1) / G): tile t belongs to
  // ceil((t+1) G / n_tile) - 1
  int c_lo = 0, c_hi = -1;
  if (det_grid > 0 && P.win_tile_ptr[win + 1] > P.win_tile_ptr[win]) {
    const long long G = det_grid, nt = P.n_tile;
    c_lo = (int)((((long long)P.win_tile_ptr[win] + 1) * G + nt - 1) / nt - 1);
    c_hi = (int)(((long long)P.win_tile_ptr[win + 1] * G + nt - 1) / nt - 1);
  }
  double pq = 0.0;
  for (int e = e0 + threadIdx.x; e < e1; e += RCTA) {
    const double qraw = det_grid > 0 ? det_sum(P.part_q, c_lo, c_hi, win, (size_t)6 * P.maxslot, e - e0) : P.q[e];
    const double qq = qraw + lam * P.p[e];
    P.q[e] = qq;
    pq += P.p[e] * qq;
  }
  pq = block_sum(pq, sh);
  const double rz = c.rz;
  const double alpha = rz / pq;
  const bool broke = !(pq > 0.0) || !isfinite(alpha);
  if (broke) {  // breakdown: keep the current iterate; LM will judge the step by its gain ratio
    if (threadIdx.x == 0) { c.cg_active = 0; atomicSub(&P.counters[1], 1); }
    return;
  }
  for (int e = e0 + threadIdx.x; e < e1; e += RCTA) {
    P.x[e] += alpha * P.p[e];
    P.res[e] -= alpha * P.q[e];
  }
  __syncthreads();
  double rzn = 0.0;
  for (int e = e0 + threadIdx.x; e < e1; e += RCTA) {
    const int s = e / 6, rr = e - s * 6;
    double z = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) z += P.Dinv[s * 36 + rr * 6 + k] * P.res[s * 6 + k];
    P.z[e] = z;
    rzn += P.res[e] * z;
  }
  rzn = block_sum(rzn, sh);
  const double beta = rzn / rz;
  for (int e = e0 + threadIdx.x; e < e1; e += RCTA) P.p[e] = P.z[e] + beta * P.p[e];
  if (threadIdx.x == 0) {
    c.rz = rzn;
    c.cg_iters++;
    if (!(rzn > tol2 * c.rz0) || c.cg_iters >= max_iters) {
      c.cg_active = 0;
      atomicSub(&P.counters[1], 1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K6: back-substitution
// dl = -R^-1 (t_l + sum_o Q1_o^T Jp_o dp[slot])  == g2o's Dinv (b_l - Hpl^T dp) (block_solver.hpp:461-483);
// also the landmark part of computeScale: sum dl (lambda dl + b_l)  (optimization_algorithm_levenberg.cpp:182-189)
__global__ void __launch_bounds__(CTA) k_backsub(Dev P, int force_all, double lam_override) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const TileInfo ti = P.tiles[blockIdx.x];
  if (wid >= ti.nitem) return;
  const int w = ti.item0 + wid;
  const int win = ti.win;
  if (!force_all && P.ctl[win].phase != PH_TRIAL) return;
  const double lam = force_all ? lam_override : P.ctl[win].lambda;
  const int start = P.item_start[w], cnt = P.item_cnt[w];
  const int Nl = P.n_point;
  const bool is_long = cnt > 32;
  const double* __restrict__ jq = P.JQ + ti.jq_off;
  double g[3] = {0, 0, 0};
  double scale = 0.0;
  int lm = -1 - lane;
  Seg sg;
  sg.start = 0; sg.end = 32;
  bool act = false;
  for (int base = 0; base < cnt; base += 32) {
    const int i = base + lane;
    act = i < cnt;
    const int o = start + (act ? i : 0);
    const int slot = act ? P.obs_slot[o] : -1;
    if (act) lm = P.obs_point[o];
    double J[18], Q[9], v[3];
    double t[3] = {0, 0, 0};
    const int fcol = is_long ? i : tile_fcol(ti, wid, slot >= 0, lane);  // long tiles keep a column per observation
    if (slot >= 0) {
      matvec_obs_v(P, jq, ti.nt, fcol, P.x, slot, (P.obs_lp[o] & LP_STEREO) != 0, J, Q, v);
#pragma unroll
      for (int k = 0; k < 3; k++) t[k] = Q[k] * v[0] + Q[3 + k] * v[1] + Q[6 + k] * v[2];
    }
    if (is_long) {
#pragma unroll
      for (int k = 0; k < 3; k++) g[k] += t[k];
    } else {
      sg = seg_of(lm, lane);
#pragma unroll
      for (int k = 0; k < 3; k++) g[k] = seg_sum(t[k], sg, lane);
    }
  }
  bool head;
  if (is_long) {
#pragma unroll
    for (int k = 0; k < 3; k++) g[k] = warp_sum(g[k]);
    lm = P.obs_point[start];
    head = lane == 0;
  } else {
    head = act && lane == sg.start;
  }
  if (head) {
    double Rm[6], tl[3], bl[3], d[3];
#pragma unroll
    for (int c = 0; c < 6; c++) Rm[c] = P.R[(size_t)c * Nl + lm];
#pragma unroll
    for (int c = 0; c < 3; c++) { tl[c] = P.tl[(size_t)c * Nl + lm]; bl[c] = P.bl[(size_t)c * Nl + lm]; }
    const double y0 = tl[0] + g[0], y1 = tl[1] + g[1], y2 = tl[2] + g[2];
    d[2] = y2 / Rm[5];
    d[1] = (y1 - Rm[4] * d[2]) / Rm[3];
    d[0] = (y0 - Rm[1] * d[1] - Rm[2] * d[2]) / Rm[0];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      d[c] = -d[c];
      P.dl[(size_t)c * Nl + lm] = d[c];
      scale += d[c] * (lam * d[c] + bl[c]);
    }
  }
  scale = warp_sum(scale);
  if (lane == 0) P.scale_part[w] = scale;
}

// ------------------------------------------------------------------------------------------------ K7: state update
// SparseOptimizer::update -> oplusImpl on every free vertex of the windows in a trial (sparse_optimizer.cpp:422-435);
// push() is the device-to-device backup the host enqueues right before (sparse_optimizer.cpp:600-603).
__global__ void k_update_pose(Dev P) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.n_slot) return;
  if (P.ctl[P.slot_win[s]].phase != PH_TRIAL) return;
  const int ip = P.slot_pose[s];
  double pose[7], xi[6];
  for (int i = 0; i < 7; i++) pose[i] = P.pose[ip * 7 + i];
  for (int i = 0; i < 6; i++) xi[i] = P.x[s * 6 + i];
  pose_oplus(pose, xi);
  for (int i = 0; i < 7; i++) P.pose[ip * 7 + i] = pose[i];
}
__global__ void k_update_point(Dev P) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= P.n_point) return;
  if (P.ctl[P.point_win[l]].phase != PH_TRIAL) return;
#pragma unroll
  for (int c = 0; c < 3; c++) P.point[l * 3 + c] += P.dl[(size_t)c * P.n_point + l];
}
// push() + update() in one launch (CUDA-graph macro step): every free pose / point of a window in a trial saves its own
// estimate right before it moves, so no device-to-device copy of the whole state is needed.  (Fixed poses never move
// and k_restore never touches them.)
__global__ void k_push_update(Dev P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n_slot && P.ctl[P.slot_win[i]].phase == PH_TRIAL) {
    const int ip = P.slot_pose[i];
    double pose[7], xi[6];
    for (int k = 0; k < 7; k++) { pose[k] = P.pose[ip * 7 + k]; P.pose_bak[ip * 7 + k] = pose[k]; }
    for (int k = 0; k < 6; k++) xi[k] = P.x[i * 6 + k];
    pose_oplus(pose, xi);
    for (int k = 0; k < 7; k++) P.pose[ip * 7 + k] = pose[k];
  }
  if (i < P.n_point && P.ctl[P.point_win[i]].phase == PH_TRIAL) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double v = P.point[i * 3 + c];
      P.point_bak[i * 3 + c] = v;
      P.point[i * 3 + c] = v + P.dl[(size_t)c * P.n_point + i];
    }
  }
}
// pop(): restore the pre-trial estimate of the windows whose trial was rejected (sparse_optimizer.cpp:605-608)
__global__ void k_restore(Dev P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n_pose) {
    if (P.pose_slot[i] >= 0 && P.ctl[P.pose_win[i]].need_restore)  // fixed poses never moved
      for (int c = 0; c < 7; c++) P.pose[i * 7 + c] = P.pose_bak[i * 7 + c];
  }
  if (i < P.n_point) {
    if (P.ctl[P.point_win[i]].need_restore)
      for (int c = 0; c < 3; c++) P.point[i * 3 + c] = P.point_bak[i * 3 + c];
  }
}

// ------------------------------------------------------------------------------------------------ K1b: trial cost
// computeActiveErrors + activeRobustChi2 at the trial state (optimization_algorithm_levenberg.cpp:123-124):
// rewrites the stored error of ACTIVE edges only -- level-1 edges keep their pass-1 value, and a rejected last
// trial leaves the trial-state errors behind, exactly like g2o's _error.
__global__ void __launch_bounds__(CTA) k_cost(Dev P, int robust, double d2, double d3) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (w >= P.n_item) return;
  const WinCtl& wc = P.ctl[P.item_win[w]];
  if (wc.phase != PH_TRIAL) return;
  if (robust < 0) robust = wc.robust;  // captured macro step: the pass's setting lives in the window's control block
  const int start = P.item_start[w], cnt = P.item_cnt[w];
  const size_t No = (size_t)P.ld;
  double chi = 0.0;
  for (int i = lane; i < cnt; i += 32) {
    const int o = start + i;
    if (P.obs_level[o] != 0) continue;
    ObsLin L;
    obs_eval(P, o, false, robust != 0, d2, d3, L);
#pragma unroll
    for (int c = 0; c < 3; c++) P.err[(size_t)c * No + o] = L.e[c];
    chi += L.rho0;
  }
  chi = warp_sum(chi);
  if (lane == 0) P.chi_part[w] = chi;
}

// ------------------------------------------------------------------------------------------------ K8b: LM decision
// The body of the do-while of OptimizationAlgorithmLevenberg::solve and its exit logic
// (optimization_algorithm_levenberg.cpp:126-163), one CTA per window.
__device__ __forceinline__ void lm_reduce_trial_body(const Dev& P, int win, double* sh) {
  if (P.ctl[win].phase != PH_TRIAL) {
    if (threadIdx.x == 0) { P.wred[win] = 0.0; P.wred[P.n_win + win] = 0.0; }
    return;
  }
  double chi = 0.0, scale = 0.0;
  for (int i = P.win_item_ptr[win] + threadIdx.x; i < P.win_item_ptr[win + 1]; i += RCTA) {
    chi += P.chi_part[i];
    scale += P.scale_part[i];
  }
  chi = block_sum(chi, sh);
  scale = block_sum(scale, sh);
  if (threadIdx.x == 0) { P.wred[win] = chi; P.wred[P.n_win + win] = scale; }
}
__global__ void __launch_bounds__(RCTA) k_lm_reduce_trial(Dev P) {
  __shared__ double sh[RWARPS];
  lm_reduce_trial_body(P, blockIdx.x, sh);
}

// term_host: optional pointer into mapped pinned host memory that mirrors the caller's stop flag (bool* pbStopFlag):
// read HERE, at the end of the trial -- the point where g2o's do-while evaluates terminate()
// (optimization_algorithm_levenberg.cpp:149) -- instead of at the top of the macro step on the host.
__device__ __forceinline__ void lm_decide_body(const Dev& P, int win, int terminate, const volatile int* term_host, double* sh) {
  WinCtl& c = P.ctl[win];
  if (c.phase != PH_TRIAL) {
    if (threadIdx.x == 0) c.need_restore = 0;
    return;
  }
  double chi = P.wred[win], scale = 0.0;
  const double lam = c.lambda;
  for (int e = P.win_slot_ptr[win] * 6 + threadIdx.x; e < P.win_slot_ptr[win + 1] * 6; e += RCTA)
    scale += P.x[e] * (lam * P.x[e] + P.bp[e]);
  scale = block_sum(scale, sh) + P.wred[P.n_win + win];
  if (threadIdx.x != 0) return;
  if (term_host && *term_host) terminate = 1;
  const double tempChi = chi;
  double rho = (c.cur_chi - tempChi);
  scale += 1e-3;
  rho /= scale;
  double* tr = nullptr;
  if (c.trace_len < P.max_trace) tr = P.trace + ((size_t)win * P.max_trace + c.trace_len) * TRACE_COLS;
  if (tr) {
    tr[0] = c.pass; tr[1] = c.iter; tr[2] = c.qmax; tr[3] = lam; tr[4] = c.cur_chi; tr[5] = tempChi; tr[6] = rho;
    tr[8] = c.cg_iters; tr[9] = (c.rz0 > 0.0) ? sqrt(fabs(c.rz) / c.rz0) : 0.0;
  }
  c.tmp_chi = tempChi;
  c.rho = rho;
  const bool good = (rho > 0.0) && isfinite(tempChi);
  if (good) {
    double alpha = 1.0 - pow((2.0 * rho - 1.0), 3.0);
    alpha = fmin(alpha, 2.0 / 3.0);
    const double scaleFactor = fmax(1.0 / 3.0, alpha);
    c.lambda *= scaleFactor;
    c.ni = 2.0;
    c.cur_chi = tempChi;
    c.need_restore = 0;
  } else {
    c.lambda *= c.ni;
    c.ni *= 2.0;
    c.need_restore = 1;
  }
  if (tr) tr[7] = good ? 1.0 : 0.0;
  c.trace_len++;
  c.qmax++;
  const bool again = (rho < 0.0) && (c.qmax < 10) && !terminate;
  if (again) return;  // stay in PH_TRIAL: same linearisation, new lambda
  bool done = false;
  if (c.qmax == 10 || rho == 0.0) {
    done = true;  // Terminate
  } else {
    if ((c.ini_chi - c.cur_chi) * 1e3 < c.ini_chi) c.nbad++; else c.nbad = 0;
    if (c.nbad >= 3) done = true;
  }
  c.iter++;
  if (c.iter >= c.max_iter || terminate) done = true;
  c.phase = done ? PH_DONE : PH_LIN;
  if (done) atomicAdd(&P.counters[0], 1);
}
__global__ void __launch_bounds__(RCTA) k_lm_decide(Dev P, int terminate) {
  __shared__ double sh[RWARPS];
  lm_decide_body(P, blockIdx.x, terminate, nullptr, sh);
}

// What the host reads after every macro step, in mapped pinned memory: no copy, no stream synchronisation -- the host
// spins on `seq` (written last, after a system-scope fence).
struct HostCtl {
  volatile int term;         // host -> device: the caller's stop flag was seen raised
  volatile int seq;          // device -> host: macro steps published so far
  volatile int counters[3];  // windows done, PCG-active windows, CG iterations of the solve so far
  int pad[3];
};

// Single-window fusion of the trial's tail (CUDA-graph macro step): cost / scale reduction, the LM decision with the
// stop flag read from mapped host memory, zeroing of the gradient accumulators when the window re-linearises next, and
// the publication of the step's counters to the host.
__global__ void __launch_bounds__(RCTA) k_decide_publish(Dev P, HostCtl* hc) {
  __shared__ double sh[RWARPS];
  lm_reduce_trial_body(P, 0, sh);
  __syncthreads();
  lm_decide_body(P, 0, 0, &hc->term, sh);
  __syncthreads();
  if (P.ctl[0].phase == PH_LIN)
    for (int i = threadIdx.x; i < P.n_slot * 6; i += RCTA) { P.bp[i] = 0.0; P.hd[i] = 0.0; }
  if (linqr_mode(P, P.ctl[0]) != 0) {  // the next step's fused kernel accumulates the trial's rhs / blocks as well
    for (int i = threadIdx.x; i < P.n_slot * 6; i += RCTA) P.bs[i] = 0.0;
    for (int i = threadIdx.x; i < P.n_slot * 21; i += RCTA) P.D[i] = 0.0;
  }
  if (threadIdx.x == 0) {
    hc->counters[0] = P.counters[0];
    hc->counters[1] = P.counters[1];
    hc->counters[2] = P.counters[2];
    const int seq = ++P.counters[5];
    __threadfence_system();
    hc->seq = seq;
  }
}

// the stop flag was seen raised at the top of an iteration: `for (i < iterations && !terminate())` does not start it
// (sparse_optimizer.cpp:376) -- windows that would have re-linearised or re-tried are finished as they are
__global__ void k_terminate(Dev P) {
  const int win = blockIdx.x * blockDim.x + threadIdx.x;
  if (win >= P.n_win) return;
  WinCtl& c = P.ctl[win];
  if (c.phase != PH_DONE) { c.phase = PH_DONE; c.need_restore = 0; c.cg_active = 0; }
}

// start of one optimize(n) call: SparseOptimizer::optimize + the iteration==0 re-initialisation of lambda
__global__ void k_pass_init(Dev P, int max_iter, int pass, int robust, double tol2) {
  const int win = blockIdx.x * blockDim.x + threadIdx.x;
  if (win >= P.n_win) return;
  WinCtl& c = P.ctl[win];
  c.iter = 0;
  c.qmax = 0;
  c.nbad = 0;
  c.phase = (max_iter > 0) ? PH_LIN : PH_DONE;
  c.max_iter = max_iter;
  c.pass = pass;
  c.robust = robust;
  c.tol2 = tol2;
  c.need_restore = 0;
  c.cg_active = 0;
  c.maxdiag_bits = 0ull;
}

// ------------------------------------------------------------------------------------------------ K9: outlier flags
// mode 0: after pass 1 -> e->setLevel(1) (g2oOptimizer.cc:952-970); mode 1: final -> erase flag (:1125-1142).
// chi2 comes from the STORED error (stale for level-1 edges by design), depth from the current estimates.
__global__ void k_classify(Dev P, int mode, double thr2d, double thr3d) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= P.n_obs) return;
  const float4 m = P.obs_meas[o];
  const bool stereo = !(m.z < 0.0f);
  const double info = (double)m.w;
  const double e0 = P.err[o], e1 = P.err[(size_t)P.ld + o], e2 = P.err[(size_t)2 * P.ld + o];
  const double c = e0 * (info * e0) + e1 * (info * e1) + e2 * (info * e2);
  const int ip = P.obs_pose[o], il = P.obs_point[o];
  double R[9], Xc[3];
  quat_to_R(P.pose + ip * 7 + 3, R);
  transform(R, P.pose + ip * 7, P.point + il * 3, Xc);
  const bool bad = (c > (stereo ? thr3d : thr2d)) || !(Xc[2] > 0.0);
  if (mode == 0) {
    if (bad) P.obs_level[o] = 1;
  } else {
    P.obs_outlier[o] = bad ? 1 : 0;
  }
}

}  // namespace sqrtba

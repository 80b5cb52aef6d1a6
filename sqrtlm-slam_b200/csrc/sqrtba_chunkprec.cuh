// Chunk preconditioner of the big-window (global BA) PCG -- the kernels that build it once per LM trial.
//
// The reduced camera system S = sum_l Jp^T (I - Q1 Q1^T) Jp + lambda I is never formed.  Its 6x6 diagonal blocks are the
// block-Jacobi preconditioner of the local-BA path (north_star).  On the 1500-keyframe loop of global BA that
// preconditioner needs ~320 CG iterations per solve, every one of them a grid-wide (and, sharded, an NVLink-wide)
// synchronisation.  Here the preconditioner keeps one block per CHUNK of VSLOT = 20 consecutive pose slots instead:
// the 120 x 120 principal submatrix of S that belongs to the chunk.  Big windows are ordered by first pose, so this is
// the coupling of a keyframe with its nearest neighbours; the coupling between chunks stays with CG.  Measured on the
// KITTI-shaped loop: 2.2-2.4x fewer iterations at the same tolerance.
//
//   k_chunk_blocks  per landmark, for every PAIR of its free-pose observations whose poses share a chunk:
//                   M[chunk][pose_i][pose_j] -= G_i G_j^T,  G = Jp^T Q1 (6x3) -- the off-diagonal 6x6 blocks (FP32);
//                   the diagonal ones are the block-Jacobi blocks the landmark QR already reduces (Dev::D)
//   [sharded: the blocks are all-reduced over the ranks]
//   k_chunk_factor  one CTA per chunk: assemble (M + M^T, D + lambda I), invert in shared memory (Gauss-Jordan on an SPD
//                   matrix, no pivoting), round to FP32, publish the packed strictly-lower triangle + diagonal for the
//                   owner CTA of the persistent PCG kernel, and start CG: z0 = p0 = M^-1 b_s, per-chunk part of r.z
//   k_chunk_rz      r0.z0 summed in chunk order (identical bits on every rank)
//
// Landmarks with more than 32 observations (a tile of their own) contribute only their diagonal blocks: the result is
// still SPD (every dropped term is the off-diagonal part of a positive semi-definite contribution whose diagonal blocks
// Jp^T (I - Q1_i Q1_i^T) Jp stay), merely a slightly weaker preconditioner.
#pragma once
#include "sqrtba_kernels.cuh"

namespace sqrtba {

constexpr int CH_LDA = CHB + 1;     // padded row of the shared-memory block (column walks without bank conflicts)
constexpr int CH_FACTOR_THREADS = 512;
constexpr size_t CH_FACTOR_SMEM = ((size_t)CHB * CH_LDA + 3 * CHB + 16) * sizeof(double);
constexpr int CH_MBLK = 36;                              // one 6x6 block, row-major, contiguous: nine 16-byte vectors
constexpr size_t CH_MSIZE = (size_t)VSLOT * VSLOT * CH_MBLK;  // floats per chunk: [pose i][pose j][6x6]

// One CTA per tile, one warp per item, one lane per observation (the layout of k_backsub).  The blocks are accumulated
// in FP32 with vector reductions (REDG.ADD.F32x4: nine per pair instead of 36 scalar FP64 ones -- the kernel is bound by
// the L2 atomic units); FP32 is what the PCG kernel keeps of the inverse anyway, and k_chunk_factor falls back to the 6x6
// blocks should a chunk ever lose definiteness.
__global__ void __launch_bounds__(CTA) k_chunk_blocks(Dev P, float* __restrict__ M) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const TileInfo ti = P.tiles[blockIdx.x];
  if (ti.is_long || wid >= ti.nitem) return;
  if (P.ctl[ti.win].phase != PH_TRIAL) return;
  const int cnt = tile_item_cnt(ti, wid), start = tile_item_start(ti, wid);
  const bool act = lane < cnt;
  const int o = start + (act ? lane : 0);
  const int slot = act ? P.obs_slot[o] : -1;
  const int lm = act ? P.obs_point[o] : -1 - lane;
  const int fcol = tile_fcol(ti, wid, slot >= 0, lane);
  float G[18];
#pragma unroll
  for (int i = 0; i < 18; i++) G[i] = 0.f;
  bool live = false;
  if (slot >= 0) {
    const double* __restrict__ jq = P.JQ + ti.jq_off;
    live = jq[(size_t)3 * ti.nt + fcol] != 0.0;  // weight 0: the edge is excluded from this pass (level 1)
    if (live) {
      double J[18], Q[9];
      load_Jp(P, jq, ti.nt, fcol, slot, (P.obs_lp[o] & LP_STEREO) != 0, J);
#pragma unroll
      for (int c = 0; c < 9; c++) Q[c] = jq[(size_t)(JG + c) * ti.nt + fcol];
#pragma unroll
      for (int c = 0; c < 6; c++)
#pragma unroll
        for (int k = 0; k < 3; k++) G[c * 3 + k] = (float)(J[c] * Q[k] + J[6 + c] * Q[3 + k] + J[12 + c] * Q[6 + k]);
    }
  }
  const Seg sg = seg_of(lm, lane);
  const int ch = live ? slot / VSLOT : -1;
  int maxlen = act ? sg.end - sg.start : 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) maxlen = max(maxlen, __shfl_xor_sync(FULL, maxlen, off));
  for (int d = 1; d < maxlen; d++) {
    const int sj = __shfl_down_sync(FULL, live ? slot : -1, d);
    float Gj[18];
#pragma unroll
    for (int i = 0; i < 18; i++) Gj[i] = __shfl_down_sync(FULL, G[i], d);
    const bool valid = live && lane + d < sg.end && sj >= 0 && sj != slot && sj / VSLOT == ch;
    if (valid) {
      float4* blk = reinterpret_cast<float4*>(M + (size_t)ch * CH_MSIZE +
                                              ((size_t)(slot - ch * VSLOT) * VSLOT + (sj - ch * VSLOT)) * CH_MBLK);
      float v[36];
#pragma unroll
      for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = 0; b < 6; b++)
          v[a * 6 + b] = -(G[a * 3] * Gj[b * 3] + G[a * 3 + 1] * Gj[b * 3 + 1] + G[a * 3 + 2] * Gj[b * 3 + 2]);
#pragma unroll
      for (int q = 0; q < 9; q++) atomicAdd(blk + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    }
  }
}

// One CTA per chunk.  rzpart[chunk] = this chunk's part of r0.z0.
__global__ void __launch_bounds__(CH_FACTOR_THREADS) k_chunk_factor(Dev P, const float* __restrict__ M, float* __restrict__ cpack,
                                                                    float* __restrict__ cdiag, double* __restrict__ rzpart) {
  extern __shared__ __align__(16) double ch_sm[];
  double* A = ch_sm;                   // [CHB][CH_LDA]
  double* colk = A + CHB * CH_LDA;     // pivot column of the current step
  double* rowk = colk + CHB;           // scaled pivot row
  double* vec = rowk + CHB;            // b_s of the chunk
  __shared__ int fail;
  const int tid = threadIdx.x;
  const int ch = blockIdx.x;
  const WinCtl& c = P.ctl[0];
  if (c.phase != PH_TRIAL) return;
  const double lam = c.lambda;
  const int s0 = ch * VSLOT;
  const int ns = min(VSLOT, P.n_slot - s0), n = ns * 6;
  const float* Mc = M + (size_t)ch * CH_MSIZE;
  // thread -> (column j, rows i0, i0 + RG, ...): no index arithmetic inside the elimination loop
  constexpr int RG = CH_FACTOR_THREADS / 128;
  static_assert((RG & (RG - 1)) == 0, "row groups: power of two");
  const int j = tid & 127, i0 = tid >> 7;
  if (tid == 0) fail = 0;
  if (j < n) {
    const int bc = j / 6, b = j - bc * 6;
    for (int r = i0; r < n; r += RG) {
      const int br = r / 6, a = r - br * 6;
      double v;
      if (br == bc) {
        const int lo = min(a, b), hi = max(a, b);
        v = P.D[(size_t)(s0 + br) * 21 + lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo)] + (r == j ? lam : 0.0);
      } else {  // every pair was added once, at one of the two places
        v = (double)Mc[((size_t)br * VSLOT + bc) * CH_MBLK + a * 6 + b] + (double)Mc[((size_t)bc * VSLOT + br) * CH_MBLK + b * 6 + a];
      }
      A[r * CH_LDA + j] = v;
    }
  }
  for (int i = tid; i < n; i += CH_FACTOR_THREADS) vec[i] = P.bs[(size_t)s0 * 6 + i];
  __syncthreads();
  // in-place Gauss-Jordan inversion; the pivots of an SPD matrix are positive
  for (int k = 0; k < n; k++) {
    const double piv = A[k * CH_LDA + k];
    if (!(piv > 0.0) || !isfinite(piv)) {
      if (tid == 0) fail = 1;
      break;  // uniform: every thread reads the same pivot
    }
    const double pinv = 1.0 / piv;
    if (tid < n) {
      colk[tid] = A[tid * CH_LDA + k];
      rowk[tid] = (tid == k) ? pinv : A[k * CH_LDA + tid] * pinv;
    }
    __syncthreads();
    if (j < n) {  // thread = one column, every RG-th row; the pivot row / column cases stay outside the inner loops
      const double rj = rowk[j];
      if (j == k) {
        for (int i = i0; i < n; i += RG)
          if (i != k) A[i * CH_LDA + k] = -colk[i] * pinv;
      } else {
        double* a = A + j;
#pragma unroll 5
        for (int i = i0; i < n; i += RG) {
          const double v = a[i * CH_LDA] - colk[i] * rj;
          if (i != k) a[i * CH_LDA] = v;
        }
      }
      if (((k - i0) & (RG - 1)) == 0) A[k * CH_LDA + j] = (j == k) ? pinv : rj;
    }
    __syncthreads();
  }
  __syncthreads();
  if (fail) {  // not positive definite in floating point: this chunk falls back to the 6x6 block-Jacobi inverses
    if (j < n) {
      const int bc = j / 6;
      for (int r = i0; r < n; r += RG) {
        const int br = r / 6;
        A[r * CH_LDA + j] = (br == bc) ? P.Dinv[(size_t)(s0 + br) * 36 + (r - br * 6) * 6 + (j - bc * 6)] : 0.0;
      }
    }
    __syncthreads();
  }
  // symmetrise, round to the precision the PCG kernel keeps, publish
  float* pk = cpack + (size_t)ch * CH_PACK;
  if (j < CHB) {
    for (int r = i0; r < CHB; r += RG) {
      if (j > r) continue;
      float f = 0.f;
      if (r < n) f = (float)(j == r ? A[r * CH_LDA + r] : 0.5 * (A[r * CH_LDA + j] + A[j * CH_LDA + r]));
      if (j == r) cdiag[(size_t)ch * CHB + r] = f;
      else pk[(r * (r - 1)) / 2 + j] = f;
    }
  }
  // CG start of the chunk with the same (rounded) operator the PCG kernel applies: z0 = p0 = M^-1 b_s
  double part = 0.0;
  if (tid < n) {
    const int r = tid;
    double z = (double)(float)A[r * CH_LDA + r] * vec[r];
    for (int k = 0; k < n; k++)
      if (k != r) z += (double)(float)(0.5 * (A[r * CH_LDA + k] + A[k * CH_LDA + r])) * vec[k];
    const size_t e = (size_t)s0 * 6 + r;
    P.z[e] = z;
    P.p[e] = z;
    part = vec[r] * z;
  }
  __syncthreads();
  part = warp_sum(part);
  if ((tid & 31) == 0) colk[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < CH_FACTOR_THREADS / 32; i++) t += colk[i];  // fixed order
    rzpart[ch] = t;
  }
}

// r0.z0 of the chunk-preconditioned start, summed in chunk order; replaces what k_cg_init / k_cg_prep left in the
// window's control block (those ran with the 6x6 inverses)
__global__ void __launch_bounds__(32) k_chunk_rz(Dev P, const double* __restrict__ rzpart, int nchunk) {
  WinCtl& c = P.ctl[0];
  if (c.phase != PH_TRIAL) return;
  const int lane = threadIdx.x;
  double s = 0.0;
  for (int i = lane; i < nchunk; i += 32) s += rzpart[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(FULL, s, off);
  if (lane == 0) {
    const int was = c.cg_active;
    const int now = (s > 0.0) ? 1 : 0;
    c.rz = s;
    c.rz0 = s;
    c.cg_active = now;
    if (now != was) atomicAdd(&P.counters[1], now - was);
  }
}

}  // namespace sqrtba

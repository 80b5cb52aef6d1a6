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

constexpr int CO_MAXCH = 128;              // second level: chunks (2560 free poses); beyond that only the chunk level is used
constexpr int CO_LD = CO_MAXCH * 6;
constexpr int CO_THREADS = 128;
constexpr int CH_LDA = CHB + 1;     // padded row of the shared-memory block (column walks without bank conflicts)
constexpr int CH_FACTOR_THREADS = 512;
constexpr size_t CH_FACTOR_SMEM = ((size_t)CHB * CH_LDA + 6 * CHB + CH_FACTOR_THREADS + 16) * sizeof(double);
constexpr int CH_MBLK = 36;                              // one 6x6 block, row-major, contiguous: nine 16-byte vectors
constexpr size_t CH_MSIZE = (size_t)VSLOT * VSLOT * CH_MBLK;  // floats per chunk: [pose i][pose j][6x6]

// Any SPD preconditioner gives the same PCG solution to tolerance, so a block that is one LM iteration old only costs
// iterations, never correctness.  A level is rebuilt on the first trial of a pass and then every `refresh`-th outer
// iteration; retries after a rejected trial (only lambda changed) reuse it as well.  Grid-uniform.
__device__ __forceinline__ bool prec_is_stale_ok(const WinCtl& c, int refresh) {
  return refresh > 1 && (c.qmax > 0 || (c.iter % refresh) != 0);
}

// One CTA per tile, one warp per item, one lane per observation (the layout of k_backsub).  The blocks are accumulated
// in FP32 with vector reductions (REDG.ADD.F32x4: nine per pair instead of 36 scalar FP64 ones -- the kernel is bound by
// the L2 atomic units); FP32 is what the PCG kernel keeps of the inverse anyway, and k_chunk_factor falls back to the 6x6
// blocks should a chunk ever lose definiteness.
__global__ void __launch_bounds__(CTA) k_chunk_blocks(Dev P, float* __restrict__ M, float* __restrict__ Cacc, int nchunk,
                                                      int refresh_chunk, int refresh_coarse) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const TileInfo ti = P.tiles[blockIdx.x];
  if (ti.is_long || wid >= ti.nitem) return;
  if (P.ctl[ti.win].phase != PH_TRIAL) return;
  const bool skip_chunk = prec_is_stale_ok(P.ctl[ti.win], refresh_chunk);
  if (Cacc != nullptr && prec_is_stale_ok(P.ctl[ti.win], refresh_coarse)) Cacc = nullptr;
  if (skip_chunk && Cacc == nullptr) return;
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
  // cross-chunk pairs towards the NEXT chunk (nearly all of them: the observations of a landmark follow the keyframe
  // order) are first summed per lane, then per warp and chunk: the few blocks of the coarse matrix are hot addresses,
  // and the L2 serialises reductions on one address (~30 cycles each)
  float cn[36];
#pragma unroll
  for (int i = 0; i < 36; i++) cn[i] = 0.f;
  bool has_cn = false;
  for (int d = 1; d < maxlen; d++) {
    const int sj = __shfl_down_sync(FULL, live ? slot : -1, d);
    float Gj[18];
#pragma unroll
    for (int i = 0; i < 18; i++) Gj[i] = __shfl_down_sync(FULL, G[i], d);
    const bool pair = live && lane + d < sg.end && sj >= 0 && sj != slot;
    const int chj = pair ? sj / VSLOT : -1;
    const bool valid = pair && (chj == ch ? !skip_chunk : Cacc != nullptr);
    if (valid) {
      float v[36];
#pragma unroll
      for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = 0; b < 6; b++)
          v[a * 6 + b] = -(G[a * 3] * Gj[b * 3] + G[a * 3 + 1] * Gj[b * 3 + 1] + G[a * 3 + 2] * Gj[b * 3 + 2]);
      if (chj == ch + 1) {
#pragma unroll
        for (int i = 0; i < 36; i++) cn[i] += v[i];
        has_cn = true;
      } else {
        // same chunk: the pair's 6x6 block of the chunk matrix; another chunk (loop closure, reversed order): its share
        // of block (chunk_i, chunk_j) of the coarse matrix Z^T S Z
        float4* blk = (chj == ch) ? reinterpret_cast<float4*>(M + (size_t)ch * CH_MSIZE +
                                                              ((size_t)(slot - ch * VSLOT) * VSLOT + (sj - ch * VSLOT)) * CH_MBLK)
                                  : reinterpret_cast<float4*>(Cacc + ((size_t)ch * nchunk + chj) * CH_MBLK);
#pragma unroll
        for (int q = 0; q < 9; q++) atomicAdd(blk + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
      }
    }
  }
  if (Cacc != nullptr) {
    unsigned todo = __ballot_sync(FULL, has_cn);
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int key = __shfl_sync(FULL, ch, leader);
      const bool mine = has_cn && ch == key;
      todo &= ~__ballot_sync(FULL, mine);
#pragma unroll
      for (int i = 0; i < 36; i++) {
        float t = mine ? cn[i] : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(FULL, t, off);
        if (lane == leader) cn[i] = t;
      }
      if (lane == leader) {
        float4* blk = reinterpret_cast<float4*>(Cacc + ((size_t)key * nchunk + key + 1) * CH_MBLK);
#pragma unroll
        for (int q = 0; q < 9; q++) atomicAdd(blk + q, make_float4(cn[4 * q], cn[4 * q + 1], cn[4 * q + 2], cn[4 * q + 3]));
      }
    }
  }
}

// One CTA per chunk.  rzpart[chunk] = this chunk's part of r0.z0.
// The Gauss-Jordan sweep keeps the matrix in REGISTERS (a 5 x 6 tile per thread, 480 of the 512 threads): with the
// matrix in shared memory every step moved all 14 400 entries through the SM's 128 B/clk shared-memory port (1.2 us per
// step); now a step writes the pivot row and column (240 doubles), one barrier, and every thread reads 11 of them.
constexpr int CH_TR = 5, CH_TC = 6;
constexpr int CH_NTC = CHB / CH_TC;                 // 20 tile columns
constexpr int CH_TILES = (CHB / CH_TR) * CH_NTC;    // 480
static_assert(CHB % CH_TR == 0 && CHB % CH_TC == 0 && CH_TILES <= CH_FACTOR_THREADS, "register tiling of the chunk block");

__global__ void __launch_bounds__(CH_FACTOR_THREADS) k_chunk_factor(Dev P, const float* __restrict__ M, float* __restrict__ cpack,
                                                                    float* __restrict__ cdiag, double* __restrict__ rzpart,
                                                                    double* __restrict__ dblk, unsigned* __restrict__ co_ctl,
                                                                    int refresh_chunk) {
  extern __shared__ __align__(16) double ch_sm[];
  double* A = ch_sm;                   // [CHB][CH_LDA]
  double* colk = A + CHB * CH_LDA;     // [2][CHB] pivot column of the current step (two parities)
  double* rowk = colk + 2 * CHB;       // [2][CHB] pivot row
  double* vec = rowk + 2 * CHB;        // b_s of the chunk
  double* dsc = vec + CHB;             // Jacobi scaling 1 / sqrt(diagonal): the sweep runs on the unit-diagonal matrix
  double* scr = dsc + CHB;             // [CH_FACTOR_THREADS] reduction scratch
  const int tid = threadIdx.x;
  const int ch = blockIdx.x;
  const WinCtl& c = P.ctl[0];
  if (c.phase != PH_TRIAL) return;
  if (co_ctl && ch == 0 && tid < 2 + CO_MAXCH) co_ctl[tid] = 0u;  // failure flag and per-step flags of k_coarse_invert
  if (prec_is_stale_ok(c, refresh_chunk)) return;  // (the coarse level is never fresher than the chunk level: dblk stays too)
  const double lam = c.lambda;
  const int s0 = ch * VSLOT;
  const int ns = min(VSLOT, P.n_slot - s0), n = ns * 6;
  const float* Mc = M + (size_t)ch * CH_MSIZE;
  constexpr int RG = CH_FACTOR_THREADS / 128;
  const int j = tid & 127, i0 = tid >> 7;
  // ---- assemble (a partly filled last chunk is padded with the identity: the sweep always runs over 120 pivots)
  if (j < CHB) {
    const int bc = j / 6, b = j - bc * 6;
#pragma unroll 2
    for (int r = i0; r < CHB; r += RG) {
      const int br = r / 6, a = r - br * 6;
      double v;
      if (r >= n || j >= n) v = (r == j) ? 1.0 : 0.0;
      else if (br == bc) {
        const int lo = min(a, b), hi = max(a, b);
        v = P.D[(size_t)(s0 + br) * 21 + lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo)] + (r == j ? lam : 0.0);
      } else {  // every pair was added once, at one of the two places
        v = (double)Mc[((size_t)br * VSLOT + bc) * CH_MBLK + a * 6 + b] + (double)Mc[((size_t)bc * VSLOT + br) * CH_MBLK + b * 6 + a];
      }
      A[r * CH_LDA + j] = v;
    }
  }
  for (int i = tid; i < CHB; i += CH_FACTOR_THREADS) vec[i] = (i < n) ? P.bs[(size_t)s0 * 6 + i] : 0.0;
  __syncthreads();
  if (dblk) {  // Z^T (chunk matrix) Z: the 6x6 sum of all its 6x6 blocks -- 14 partial sums per entry, added in order
    constexpr int NP = CH_FACTOR_THREADS / 36;
    if (tid < NP * 36) {
      const int en = tid % 36, part = tid / 36, a = en / 6, b = en - a * 6;
      double sum = 0.0;
      for (int q = part; q < ns * ns; q += NP) {
        const int bi = q / ns, bj = q - bi * ns;
        sum += A[(bi * 6 + a) * CH_LDA + bj * 6 + b];
      }
      scr[part * 36 + en] = sum;
    }
    __syncthreads();
    if (tid < 36) {
      double sum = 0.0;
      for (int part = 0; part < NP; part++) sum += scr[part * 36 + tid];
      dblk[(size_t)ch * 36 + tid] = sum;
    }
    __syncthreads();
  }
  // ---- Gauss-Jordan sweep over all pivots; the pivots of an SPD matrix are positive.
  // The matrix is first scaled to unit diagonal (pivots in (0, 1]).  That makes it safe to fold the special cases of a
  // sweep step -- pivot row, pivot column, pivot -- into the SAME rank-one update as everything else:
  //   a_ij <- a_ij - c_i r_j   with  r_j = a_kj / p (j != k),  r_k = 1 + 1/p,  c_i = a_ik (i != k),  c_k = p - 1
  // gives a_kj / p on the pivot row, -a_ik / p on the pivot column and 1/p at the pivot, without cancellation for p <= 1.
  // A step is then 30 DFMA per thread and no branches on the position of the pivot (every warp holds a piece of the
  // pivot column, so any special-casing is paid by all warps).
  for (int i = tid; i < CHB; i += CH_FACTOR_THREADS) {
    const double d = A[i * CH_LDA + i];
    dsc[i] = (d > 0.0 && isfinite(d)) ? rsqrt(d) : 0.0;   // 0: the first pivot check fails -> fallback
  }
  __syncthreads();
  if (j < CHB) {
    const double dj = dsc[j];
    for (int r = i0; r < CHB; r += RG) A[r * CH_LDA + j] *= dsc[r] * dj;
  }
  __syncthreads();
  const bool worker = tid < CH_TILES;
  const int tr = tid / CH_NTC, tc = tid - tr * CH_NTC;
  double t[CH_TR][CH_TC];
  if (worker) {
#pragma unroll
    for (int a = 0; a < CH_TR; a++)
#pragma unroll
      for (int b = 0; b < CH_TC; b++) t[a][b] = A[(tr * CH_TR + a) * CH_LDA + tc * CH_TC + b];
  }
  bool failed = false;
  double* pvs = scr;  // [2][2]: pivot and its reciprocal, by parity
  for (int k = 0; k < CHB; k++) {
    double* ck = colk + (k & 1) * CHB;
    double* rk = rowk + (k & 1) * CHB;
    const int ik = k - tr * CH_TR, jk = k - tc * CH_TC;  // position of the pivot row / column inside this thread's tile
    const bool has_row = worker && (unsigned)ik < (unsigned)CH_TR, has_col = worker && (unsigned)jk < (unsigned)CH_TC;
    if (has_row) {
#pragma unroll
      for (int a = 0; a < CH_TR; a++)
        if (a == ik) {
#pragma unroll
          for (int b = 0; b < CH_TC; b++) rk[tc * CH_TC + b] = t[a][b];
        }
    }
    if (has_col) {
#pragma unroll
      for (int b = 0; b < CH_TC; b++)
        if (b == jk) {
#pragma unroll
          for (int a = 0; a < CH_TR; a++) ck[tr * CH_TR + a] = t[a][b];
        }
    }
    if (has_row && has_col) {  // the thread that holds the pivot patches the two special entries behind its own writes
      double pv = 0.0;
#pragma unroll
      for (int a = 0; a < CH_TR; a++)
#pragma unroll
        for (int b = 0; b < CH_TC; b++)
          if (a == ik && b == jk) pv = t[a][b];
      rk[k] = pv + 1.0;  // times 1/p: r_k = 1 + 1/p
      ck[k] = pv - 1.0;  // c_k = p - 1
      pvs[(k & 1) * 2] = pv;
      pvs[(k & 1) * 2 + 1] = 1.0 / pv;
    }
    __syncthreads();
    const double piv = pvs[(k & 1) * 2];
    if (!(piv > 0.0) || !isfinite(piv)) { failed = true; break; }  // uniform: every thread reads the same pivot
    const double pinv = pvs[(k & 1) * 2 + 1];
    if (worker) {
      double cc[CH_TR], rr[CH_TC];
#pragma unroll
      for (int a = 0; a < CH_TR; a++) cc[a] = ck[tr * CH_TR + a];
#pragma unroll
      for (int b = 0; b < CH_TC; b++) rr[b] = rk[tc * CH_TC + b] * pinv;
#pragma unroll
      for (int a = 0; a < CH_TR; a++)
#pragma unroll
        for (int b = 0; b < CH_TC; b++) t[a][b] -= cc[a] * rr[b];
    }
    // no second barrier: the next step writes the other parity, and nobody reaches the step after it before everyone
    // has passed the next barrier
  }
  __syncthreads();
  if (!failed) {
    if (worker) {
#pragma unroll
      for (int a = 0; a < CH_TR; a++)
#pragma unroll
        for (int b = 0; b < CH_TC; b++) A[(tr * CH_TR + a) * CH_LDA + tc * CH_TC + b] = t[a][b];
    }
    __syncthreads();
    if (j < CHB) {  // undo the scaling: inverse of the original block
      const double dj = dsc[j];
      for (int r = i0; r < CHB; r += RG) A[r * CH_LDA + j] *= dsc[r] * dj;
    }
  } else {  // not positive definite in floating point: this chunk falls back to 6x6 block-Jacobi inverses
    for (int idx = tid; idx < CHB * CHB; idx += CH_FACTOR_THREADS) A[(idx / CHB) * CH_LDA + (idx % CHB)] = 0.0;
    __syncthreads();
    if (tid < ns) {
      double B6[36], Bi[36];
      int idx = 0;
      for (int a = 0; a < 6; a++)
        for (int b = a; b < 6; b++) {
          const double v = P.D[(size_t)(s0 + tid) * 21 + idx] + (a == b ? lam : 0.0);
          B6[a * 6 + b] = v;
          B6[b * 6 + a] = v;
          idx++;
        }
      if (!spd6_inverse(B6, Bi)) {
        for (int i = 0; i < 36; i++) Bi[i] = 0.0;
        for (int i = 0; i < 6; i++) Bi[i * 6 + i] = 1.0 / fmax(fabs(B6[i * 6 + i]), 1e-300);
      }
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) A[(tid * 6 + a) * CH_LDA + tid * 6 + b] = Bi[a * 6 + b];
    }
  }
  __syncthreads();
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
  (void)rzpart;
  (void)vec;
}

// ---------------------------------------------------------------------------------------------- second level (coarse)
// Block-Jacobi over chunks leaves the coupling BETWEEN chunks -- the smooth, loop-wide error modes -- to CG, and those
// are exactly what a 1500-keyframe loop converges slowly on.  The second level is the classical additive coarse
// correction  M^-1 = blockdiag(chunk inverses) + Z (Z^T S Z)^-1 Z^T  with Z = one column per (chunk, tangent component):
// 6 nchunk unknowns (450 on C3).  Z^T S Z is assembled from sums that are already there: its diagonal blocks are the
// 6x6 block sums of the chunk matrices (k_chunk_factor), block (c, c') the sum of the cross-chunk pair blocks
// (k_chunk_blocks).  Measured on the KITTI-shaped loop at small lambda: another 3.7x fewer iterations on top of the chunks.

// Cooperative launch (co-residency), one CTA per chunk = one 6-row block row of the coarse matrix in shared memory.  Block
// Gauss-Jordan with 6x6 pivots: at step k the owner of row k inverts its pivot block, publishes R = P^-1 row_k in the
// step's own buffer and raises the step's flag; every other row waits for that flag (no grid barrier: the critical path
// is the chain owner k -> owner k+1) and eliminates its block column k.
__global__ void __launch_bounds__(CO_THREADS) k_coarse_invert(Dev P, const float* __restrict__ Cacc, const double* __restrict__ dblk,
                                                              double* __restrict__ Rbuf, double* __restrict__ aci,
                                                              unsigned* __restrict__ co_ctl, int nchunk, int refresh) {
  __shared__ double row[6][CO_LD];
  __shared__ double Pm[36], Pinv[36], Cc[36];
  __shared__ int bad;
  const int tid = threadIdx.x, c = blockIdx.x;
  if (P.ctl[0].phase != PH_TRIAL) return;  // grid-uniform
  // Any symmetric positive semi-definite coarse operator keeps the preconditioner SPD, so a slightly stale one only costs
  // iterations, never correctness: refresh on the first trial of a pass and then every `refresh`-th outer iteration
  // (retries after a rejected trial only change lambda and reuse it as well).
  if (prec_is_stale_ok(P.ctl[0], refresh)) return;  // grid-uniform
  const int nc6 = nchunk * 6;
  for (int k = tid; k < nc6; k += CO_THREADS) {
    const int c2 = k / 6, b = k - c2 * 6;
#pragma unroll
    for (int a = 0; a < 6; a++)
      row[a][k] = (c2 == c) ? dblk[(size_t)c * 36 + a * 6 + b]
                            : (double)Cacc[((size_t)c * nchunk + c2) * CH_MBLK + a * 6 + b] +
                                  (double)Cacc[((size_t)c2 * nchunk + c) * CH_MBLK + b * 6 + a];
  }
  __syncthreads();
  for (int kb = 0; kb < nchunk; kb++) {
    double* R = Rbuf + (size_t)kb * (6 * CO_LD + 36);
    if (c == kb) {
      if (tid < 36) Pm[tid] = row[tid / 6][kb * 6 + (tid % 6)];
      if (tid == 0) bad = 0;
      __syncthreads();
      if (tid < 32) {  // 6x6 Gauss-Jordan by one warp: lane = entry (and entry + 32 for the first four lanes)
        for (int k = 0; k < 6; k++) {
          const double piv = Pm[k * 6 + k];
          if (!(piv > 0.0) || !isfinite(piv)) { if (tid == 0) bad = 1; break; }
          const double pinv = 1.0 / piv;
          double nv[2];
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int en = tid + 32 * h;
            nv[h] = 0.0;
            if (en < 36) {
              const int i = en / 6, j = en - i * 6;
              if (i == k) nv[h] = (j == k) ? pinv : Pm[k * 6 + j] * pinv;
              else if (j == k) nv[h] = -Pm[i * 6 + k] * pinv;
              else nv[h] = Pm[en] - Pm[i * 6 + k] * Pm[k * 6 + j] * pinv;
            }
          }
          __syncwarp();
          Pm[tid] = nv[0];
          if (tid < 4) Pm[tid + 32] = nv[1];
          __syncwarp();
        }
      }
      __syncthreads();
      if (tid < 36) Pinv[tid] = bad ? 0.0 : Pm[tid];
      if (bad && tid == 0) atomicExch(&co_ctl[1], 1u);
      __syncthreads();
      for (int k = tid; k < nc6; k += CO_THREADS) {
        double rk[6];
        const bool pc = k >= kb * 6 && k < kb * 6 + 6;
#pragma unroll
        for (int a = 0; a < 6; a++) {
          double v = 0.0;
          if (pc) v = Pinv[a * 6 + (k - kb * 6)];
          else {
#pragma unroll
            for (int q = 0; q < 6; q++) v += Pinv[a * 6 + q] * row[q][k];
          }
          rk[a] = v;
        }
#pragma unroll
        for (int a = 0; a < 6; a++) { row[a][k] = rk[a]; R[(size_t)a * CO_LD + k] = rk[a]; }
      }
      if (tid < 36) R[6 * CO_LD + tid] = Pinv[tid];
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(co_ctl + 2 + kb), "r"(1u) : "memory");
      }
    } else {
      if (tid == 0) {
        while (ld_acquire_gpu(co_ctl + 2 + kb) == 0u) {}
      }
      __syncthreads();
      if (tid < 36) { Cc[tid] = row[tid / 6][kb * 6 + (tid % 6)]; Pinv[tid] = __ldcg(R + 6 * CO_LD + tid); }
      __syncthreads();
      // all of R this thread needs in flight at once (one L2 round trip), then the arithmetic
      constexpr int KU = CO_LD / CO_THREADS;
      double rk[KU][6];
#pragma unroll
      for (int u = 0; u < KU; u++) {
        const int k = tid + u * CO_THREADS;
        const bool pc = k >= kb * 6 && k < kb * 6 + 6;
#pragma unroll
        for (int q = 0; q < 6; q++) rk[u][q] = (k < nc6 && !pc) ? __ldcg(R + (size_t)q * CO_LD + k) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < KU; u++) {
        const int k = tid + u * CO_THREADS;
        if (k < nc6) {
          const bool pc = k >= kb * 6 && k < kb * 6 + 6;
#pragma unroll
          for (int a = 0; a < 6; a++) {
            double v = pc ? 0.0 : row[a][k];
#pragma unroll
            for (int q = 0; q < 6; q++) v -= Cc[a * 6 + q] * (pc ? Pinv[q * 6 + (k - kb * 6)] : rk[u][q]);
            row[a][k] = v;
          }
        }
      }
      __syncthreads();
    }
  }
  const bool failed = __ldcg(&co_ctl[1]) != 0u;  // written before the flag of a step every CTA has waited for (or by this CTA)
  for (int k = tid; k < nc6; k += CO_THREADS) {
#pragma unroll
    for (int a = 0; a < 6; a++) aci[(size_t)(c * 6 + a) * nc6 + k] = failed ? 0.0 : row[a][k];
  }
}

// CG start with the operator the PCG kernel applies: z0 = p0 = (packed FP32 chunk inverse) b_s + Z (Z^T S Z)^-1 Z^T b_s;
// r0.z0 per chunk.  Reads the PUBLISHED inverses, so it is the same whether they were rebuilt for this trial or not.
__global__ void __launch_bounds__(CO_THREADS) k_chunk_z0(Dev P, const float* __restrict__ cpack, const float* __restrict__ cdiag,
                                                         const double* __restrict__ aci, double* __restrict__ rzpart, int nchunk) {
  __shared__ double w[CO_LD];
  __shared__ double ysh[4][6];
  __shared__ double bsh[CHB];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, ch = blockIdx.x;
  if (P.ctl[0].phase != PH_TRIAL) return;
  const int nc6 = nchunk * 6;
  const int n = min(VSLOT, P.n_slot - ch * VSLOT) * 6;
  if (tid < CHB) bsh[tid] = (tid < n) ? P.bs[(size_t)ch * CHB + tid] : 0.0;
  if (aci != nullptr) {
    for (int k = tid; k < nc6; k += CO_THREADS) {
      const int c2 = k / 6, a = k - c2 * 6;
      const int ns = min(VSLOT, P.n_slot - c2 * VSLOT);
      double sum = 0.0;
      for (int sl = 0; sl < ns; sl++) sum += P.bs[(size_t)(c2 * VSLOT + sl) * 6 + a];
      w[k] = sum;
    }
  }
  __syncthreads();
  double acc[6] = {0, 0, 0, 0, 0, 0};
  if (aci != nullptr) {
    for (int k = tid; k < nc6; k += CO_THREADS) {
#pragma unroll
      for (int a = 0; a < 6; a++) acc[a] += aci[(size_t)(ch * 6 + a) * nc6 + k] * w[k];
    }
  }
#pragma unroll
  for (int a = 0; a < 6; a++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc[a] += __shfl_down_sync(FULL, acc[a], off);
    if (lane == 0) ysh[wid][a] = acc[a];
  }
  __syncthreads();
  double part = 0.0;
  if (tid < n) {
    const int r = tid, a = tid % 6;
    const float* pk = cpack + (size_t)ch * CH_PACK;
    double z0 = (double)cdiag[(size_t)ch * CHB + r] * bsh[r], z1 = 0.0;
    const float* row = pk + (r * (r - 1)) / 2;
    int k = 0;
    for (; k + 1 < r; k += 2) {
      z0 += (double)row[k] * bsh[k];
      z1 += (double)row[k + 1] * bsh[k + 1];
    }
    if (k < r) z0 += (double)row[k] * bsh[k];
    int idx = ((r + 1) * r) / 2 + r;
    for (k = r + 1; k < n; k++) {
      z1 += (double)pk[idx] * bsh[k];
      idx += k;
    }
    const double z = (z0 + z1) + ((ysh[0][a] + ysh[1][a]) + (ysh[2][a] + ysh[3][a]));
    const size_t e = (size_t)ch * CHB + tid;
    P.z[e] = z;
    P.p[e] = z;
    part = bsh[r] * z;
  }
  part = warp_sum(part);
  __syncthreads();
  if (lane == 0) w[wid] = part;
  __syncthreads();
  if (tid == 0) rzpart[ch] = (w[0] + w[1]) + (w[2] + w[3]);
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

// Essential-graph optimisation on the device: the Sim3 pose-graph Levenberg that
// g2oOptimizer::OptimizeEssentialGraph sets up (src/backend/g2oOptimizer.cc:1212-1460) -- VertexSim3Expmap vertices,
// EdgeSim3 edges with identity information and the NUMERIC Jacobians that edge type inherits (central differences,
// delta = 1e-9, through oplusImpl; Thirdparty/g2o/g2o/core/base_binary_edge.hpp:122-195), BlockSolver_7_3 without
// marginalised vertices, LinearSolverEigen (an exact sparse Cholesky), setUserLambdaInit(1e-16), optimize(20).
//
// Device design.  The system has one 7x7 block row per free keyframe (10 500 unknowns at KITTI-00 length) and is
// banded along the keyframe ids (spanning tree + covisibility edges) plus a few long rows (loop edges).  It is stored
// as a BLOCK SKYLINE -- block row r keeps the blocks [first[r], r] -- and factorised exactly by a right-looking block
// Cholesky inside ONE thread block: per block column the 7x7 pivot is factorised, the column's blocks are scaled by its
// inverse and every pair of rows in the column gets its rank-7 update; all the rows of a column (band rows and long
// rows alike) are updated in parallel, so a loop edge costs work, not sequential steps.  Fill never leaves the skyline.
// An exact solve reproduces the reference's LM trial sequence; the linearisation (28 Sim3 error evaluations per edge)
// and the assembly run over all SMs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sqrtba_sim3.cuh"

namespace sqrtba {

struct PgDev {
  int n_vert, n_edge, n_slot, fix_scale;
  double* vert;            // n_vert x 8 (qx qy qz qw | tx ty tz | s), current estimate
  double* vert_bak;        // push() copy
  const uint8_t* fixed;    // n_vert
  const int* slot;         // n_vert: block row of the vertex or -1 (fixed / no edge)
  const int* slot_vert;    // n_slot
  const int* edge_ij;      // n_edge x 2 (vertex 0 = i, vertex 1 = j)
  const double* meas;      // n_edge x 8: S_ji
  double* err;             // n_edge x 7
  double* Ji;              // n_edge x 49 (row = error component)
  double* Jj;
  const int* first;        // n_slot: first block column stored for the row
  const long long* rowptr; // n_slot + 1, in blocks
  double* H;               // blocks x 49: J^T J (lower triangle by block rows, blocks row-major 7x7)
  double* L;               // blocks x 49: H + lambda I, factorised in place (L of L L^T, lower triangle)
  double* Linv;            // n_slot x 49: inverse of every diagonal factor block
  const int* col_ptr;      // n_slot + 1 -> col_rows: the rows r > k whose skyline contains column k, ascending
  const int* col_rows;
  double* b;               // 7 x n_slot: -J^T e
  double* y;               // work vector of the solve
  double* x;               // the step
  double* scal;            // [1] sum x (lambda x + b) [2] factorisation failed (non-positive pivot)
  double* chi_part;        // one partial chi2 per CTA of k_pg_errors (summed in order by the host: reproducible)
  const int* inc_ptr;      // n_slot + 1 -> inc_edge: the edges of every free vertex, ascending edge index;
  const int* inc_edge;     //   entry = 2 * edge + side (side 0: the vertex is the edge's vertex i, 1: vertex j)
};

__global__ void k_pg_errors(PgDev G) {
  __shared__ double sh[8];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double c = 0.0;
  if (e < G.n_edge) {
    const int i = G.edge_ij[2 * e], j = G.edge_ij[2 * e + 1];
    double er[7];
    sim3_edge_error(G.meas + (size_t)e * 8, G.vert + (size_t)i * 8, G.vert + (size_t)j * 8, er);
#pragma unroll
    for (int k = 0; k < 7; k++) { G.err[(size_t)e * 7 + k] = er[k]; c += er[k] * er[k]; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
    G.chi_part[blockIdx.x] = t;
  }
}

// one thread per (edge, vertex side, tangent direction): the two perturbed error evaluations of that Jacobian column
__global__ void k_pg_linearize(PgDev G) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)G.n_edge * 14) return;
  const int e = (int)(t / 14), sd = (int)(t - (long long)e * 14), side = sd / 7, d = sd - side * 7;
  const int i = G.edge_ij[2 * e], j = G.edge_ij[2 * e + 1];
  const int vi = side == 0 ? i : j;
  double* J = (side == 0 ? G.Ji : G.Jj) + (size_t)e * 49;
  if (G.fixed[vi]) {  // the Jacobian of a fixed vertex is never used
#pragma unroll
    for (int r = 0; r < 7; r++) J[r * 7 + d] = 0.0;
    return;
  }
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  const double* C8 = G.meas + (size_t)e * 8;
  const double* v1 = G.vert + (size_t)i * 8;
  const double* v2 = G.vert + (size_t)j * 8;
  double add[7] = {0, 0, 0, 0, 0, 0, 0}, p[8], e1[7], e2[7];
  add[d] = delta;
#pragma unroll
  for (int k = 0; k < 8; k++) p[k] = side == 0 ? v1[k] : v2[k];
  sim3_oplus(p, add, G.fix_scale != 0);
  sim3_edge_error(C8, side == 0 ? p : v1, side == 0 ? v2 : p, e1);
  add[d] = -delta;
#pragma unroll
  for (int k = 0; k < 8; k++) p[k] = side == 0 ? v1[k] : v2[k];
  sim3_oplus(p, add, G.fix_scale != 0);
  sim3_edge_error(C8, side == 0 ? p : v1, side == 0 ? v2 : p, e2);
#pragma unroll
  for (int r = 0; r < 7; r++) J[r * 7 + d] = scalar * (e1[r] - e2[r]);
}

// constructQuadraticForm with omega = I, no robust kernel (base_binary_edge.hpp:97-117).  One thread per (block row,
// a, c) walks the row's edges in ascending edge index and adds its diagonal block, its off-diagonal blocks (stored in
// the row of the larger slot) and its part of b: no atomics, so the system -- and with it every LM decision -- is
// reproducible from run to run (the numeric Jacobians amplify rounding noise by 1/delta = 1e9).
__global__ void k_pg_assemble(PgDev G) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)G.n_slot * 49) return;
  const int R = (int)(t / 49), ac = (int)(t - (long long)R * 49), a = ac / 7, c = ac - a * 7;
  const long long rowbase = G.rowptr[R] - G.first[R];
  double diag = 0.0, bb = 0.0;
  for (int k = G.inc_ptr[R]; k < G.inc_ptr[R + 1]; k++) {
    const int e = G.inc_edge[k] >> 1, side = G.inc_edge[k] & 1;
    const double* Jr = (side == 0 ? G.Ji : G.Jj) + (size_t)e * 49;  // Jacobian of this row's vertex
    const double* Jo = (side == 0 ? G.Jj : G.Ji) + (size_t)e * 49;  // ... of the other vertex of the edge
    const int other = G.slot[G.edge_ij[2 * e + (side == 0 ? 1 : 0)]];
    double h = 0.0;
#pragma unroll
    for (int r = 0; r < 7; r++) h += Jr[r * 7 + a] * Jr[r * 7 + c];
    diag += h;
    if (c == 0) {
      const double* er = G.err + (size_t)e * 7;
      double sacc = 0.0;
#pragma unroll
      for (int r = 0; r < 7; r++) sacc += Jr[r * 7 + a] * er[r];
      bb -= sacc;
    }
    if (other >= 0 && other < R) {
      double x = 0.0;
#pragma unroll
      for (int r = 0; r < 7; r++) x += Jr[r * 7 + a] * Jo[r * 7 + c];
      G.H[(size_t)(rowbase + other) * 49 + ac] += x;  // only this thread touches this element
    }
  }
  G.H[(size_t)(rowbase + R) * 49 + ac] = diag;
  if (c == 0) G.b[(size_t)R * 7 + a] = bb;
}

// L = H, y = b; then lambda on every diagonal entry (setLambda, block_solver.hpp:564-589)
__global__ void k_pg_prepare(PgDev G, long long n_blocks) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_blocks * 49) G.L[t] = G.H[t];
  if (t < (long long)G.n_slot * 7) G.y[t] = G.b[t];
}
__global__ void k_pg_damp(PgDev G, double lambda) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G.n_slot * 7) return;
  const int r = t / 7, a = t - r * 7;
  G.L[(size_t)(G.rowptr[r] + (r - G.first[r])) * 49 + a * 8] += lambda;
}

constexpr int PG_THREADS = 1024;
constexpr int PG_CHUNK = (PG_THREADS / 49) * 49;  // whole 7x7 blocks per sweep of the column scaling
constexpr int PG_SM_ROWS = 96;  // column blocks staged in shared memory when the column has at most this many rows

// Right-looking block Cholesky of the skyline + both triangular solves + the pose part of computeScale, ONE thread block.
__global__ void __launch_bounds__(PG_THREADS) k_pg_factor_solve(PgDev G, double lambda) {
  __shared__ double Akk[49], Lkk[49], Lki[49], yk[7], red[32];
  __shared__ double colb[PG_SM_ROWS * 49];
  __shared__ int fail;
  const int tid = threadIdx.x, n = G.n_slot;
  if (tid == 0) fail = 0;
  __syncthreads();
  for (int k = 0; k < n; k++) {
    const long long dk = G.rowptr[k] + (k - G.first[k]);
    if (tid < 49) Akk[tid] = G.L[(size_t)dk * 49 + tid];
    __syncthreads();
    if (tid < 32) {  // 7x7 Cholesky of the pivot by one warp: lane r < 7 owns row r in registers, the pivot row travels
                     // by shuffles (a single-thread version over shared memory is one chain of dependent 30-cycle
                     // accesses: 8 us per pivot)
      const int r = tid;
      double a[7], l[7];
#pragma unroll
      for (int c = 0; c < 7; c++) { a[c] = (r < 7) ? Akk[r * 7 + c] : 0.0; l[c] = 0.0; }
      bool ok = true;
#pragma unroll
      for (int c = 0; c < 7; c++) {
        double v = a[c];  // row r, column c minus the part already eliminated:  sum_{s<c} L[r][s] * L[c][s]
#pragma unroll
        for (int s2 = 0; s2 < 7; s2++)
          if (s2 < c) v -= l[s2] * __shfl_sync(0xffffffffu, l[s2], c);
        double d = __shfl_sync(0xffffffffu, v, c);  // the pivot's own diagonal entry
        if (!(d > 0.0) || !isfinite(d)) { ok = false; d = 1.0; }
        const double lcc = sqrt(d);
        l[c] = (r == c) ? lcc : ((r > c) ? v / lcc : 0.0);
      }
      if (r < 7) {
#pragma unroll
        for (int c = 0; c < 7; c++) Lkk[r * 7 + c] = l[c];
      }
      __syncwarp();
      if (r < 7) {  // lane r: column r of the inverse of L by forward substitution (L read from shared memory: broadcasts)
        double x[7];
#pragma unroll
        for (int i = 0; i < 7; i++) {
          double v = (i == r) ? 1.0 : 0.0;
#pragma unroll
          for (int s2 = 0; s2 < 7; s2++)
            if (s2 < i) v -= Lkk[i * 7 + s2] * x[s2];
          x[i] = (i < r) ? 0.0 : v / Lkk[i * 7 + i];
        }
#pragma unroll
        for (int i = 0; i < 7; i++) Lki[i * 7 + r] = x[i];
      }
      if (r == 0 && !ok) fail = 1;
    }
    __syncthreads();
    if (tid < 49) {
      G.L[(size_t)dk * 49 + tid] = Lkk[tid];
      G.Linv[(size_t)k * 49 + tid] = Lki[tid];
    }
    const int c0 = G.col_ptr[k], m = G.col_ptr[k + 1] - c0;
    const bool staged = m <= PG_SM_ROWS;
    // column blocks: L_ik = A_ik * Lkk^-T   (element (p,q) = sum_s A_ik[p][s] * Lki[q][s])
    for (int base = 0; base < m * 49; base += PG_CHUNK) {  // whole blocks per sweep: an element reads its block's row
      const int t = base + tid;
      const bool mine = tid < PG_CHUNK && t < m * 49;
      double row[7], out = 0.0;
      size_t addr = 0;
      int q = 0;
      if (mine) {
        const int a = t / 49, pq = t - a * 49, p = pq / 7;
        q = pq - p * 7;
        const int i = G.col_rows[c0 + a];
        addr = (size_t)(G.rowptr[i] + (k - G.first[i])) * 49;
#pragma unroll
        for (int s = 0; s < 7; s++) row[s] = G.L[addr + p * 7 + s];
#pragma unroll
        for (int s = 0; s < 7; s++) out += row[s] * Lki[q * 7 + s];
        addr += pq;
      }
      __syncthreads();  // every element of the sweep has read its row before any element is overwritten
      if (mine) {
        G.L[addr] = out;
        if (staged) colb[t] = out;
      }
    }
    __syncthreads();
    // trailing update over every pair of rows of the column: A[ia][ib] -= L_ia,k * L_ib,k^T   (ia >= ib)
    const long long npair = (long long)m * (m + 1) / 2;
    for (long long t = tid; t < npair * 49; t += PG_THREADS) {
      const long long pr = t / 49;
      const int pq = (int)(t - pr * 49), p = pq / 7, q = pq - p * 7;
      int a = (int)((sqrt(8.0 * (double)pr + 1.0) - 1.0) * 0.5);
      while ((long long)a * (a + 1) / 2 > pr) a--;
      while ((long long)(a + 1) * (a + 2) / 2 <= pr) a++;
      const int bb = (int)(pr - (long long)a * (a + 1) / 2);
      const int ia = G.col_rows[c0 + a], ib = G.col_rows[c0 + bb];
      double acc = 0.0;
      if (staged) {
        const double* La = colb + a * 49 + p * 7;
        const double* Lb = colb + bb * 49 + q * 7;
#pragma unroll
        for (int s = 0; s < 7; s++) acc += La[s] * Lb[s];
      } else {
        const double* La = G.L + (size_t)(G.rowptr[ia] + (k - G.first[ia])) * 49 + p * 7;
        const double* Lb = G.L + (size_t)(G.rowptr[ib] + (k - G.first[ib])) * 49 + q * 7;
#pragma unroll
        for (int s = 0; s < 7; s++) acc += La[s] * Lb[s];
      }
      G.L[(size_t)(G.rowptr[ia] + (ib - G.first[ia])) * 49 + pq] -= acc;
    }
    __syncthreads();
  }
  // forward substitution  L y = b  and backward substitution  L^T x = y, in place on one shared-memory vector
  extern __shared__ double ysh[];  // 7 * n_slot doubles (dynamic), preceded by nothing
  for (int t = tid; t < n * 7; t += PG_THREADS) ysh[t] = G.y[t];
  __syncthreads();
  for (int k = 0; k < n; k++) {
    if (tid < 7) {
      double v = 0.0;
#pragma unroll
      for (int s2 = 0; s2 < 7; s2++) v += G.Linv[(size_t)k * 49 + tid * 7 + s2] * ysh[k * 7 + s2];
      yk[tid] = v;
    }
    __syncthreads();
    if (tid < 7) ysh[k * 7 + tid] = yk[tid];
    const int c0 = G.col_ptr[k], m = G.col_ptr[k + 1] - c0;
    for (int t = tid; t < m * 7; t += PG_THREADS) {
      const int a = t / 7, p = t - a * 7;
      const int i = G.col_rows[c0 + a];
      const double* Lik = G.L + (size_t)(G.rowptr[i] + (k - G.first[i])) * 49 + p * 7;
      double v = 0.0;
#pragma unroll
      for (int s2 = 0; s2 < 7; s2++) v += Lik[s2] * yk[s2];
      ysh[i * 7 + p] -= v;
    }
    __syncthreads();
  }
  for (int k = n - 1; k >= 0; k--) {
    const int c0 = G.col_ptr[k], m = G.col_ptr[k + 1] - c0;
    const int q = tid >> 5, lane = tid & 31;  // warp q < 7 gathers component q of  sum_i L_ik^T x_i
    if (q < 7) {
      double v = 0.0;
      for (int a = lane; a < m; a += 32) {
        const int i = G.col_rows[c0 + a];
        const double* Lik = G.L + (size_t)(G.rowptr[i] + (k - G.first[i])) * 49;
#pragma unroll
        for (int p = 0; p < 7; p++) v += Lik[p * 7 + q] * ysh[i * 7 + p];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) yk[q] = ysh[k * 7 + q] - v;
    }
    __syncthreads();
    if (tid < 7) {  // x_k = Lkk^-T * yk
      double v = 0.0;
#pragma unroll
      for (int s2 = 0; s2 < 7; s2++) v += G.Linv[(size_t)k * 49 + s2 * 7 + tid] * yk[s2];
      ysh[k * 7 + tid] = v;
    }
    __syncthreads();
  }
  for (int t = tid; t < n * 7; t += PG_THREADS) G.x[t] = ysh[t];
  __syncthreads();
  // computeScale: sum_j x_j (lambda x_j + b_j)   (optimization_algorithm_levenberg.cpp:182-189)
  double sc = 0.0;
  for (int t = tid; t < n * 7; t += PG_THREADS) sc += G.x[t] * (lambda * G.x[t] + G.b[t]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, off);
  if ((tid & 31) == 0) red[tid >> 5] = sc;
  __syncthreads();
  if (tid == 0) {
    double tsum = 0.0;
    for (int w = 0; w < PG_THREADS / 32; w++) tsum += red[w];
    G.scal[1] = tsum;
    G.scal[2] = fail ? 1.0 : 0.0;
  }
}

// push + update (VertexSim3Expmap::oplusImpl on every free vertex)
__global__ void k_pg_update(PgDev G) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= G.n_slot) return;
  const int v = G.slot_vert[s];
  double est[8], upd[7];
#pragma unroll
  for (int k = 0; k < 8; k++) { est[k] = G.vert[(size_t)v * 8 + k]; G.vert_bak[(size_t)v * 8 + k] = est[k]; }
#pragma unroll
  for (int k = 0; k < 7; k++) upd[k] = G.x[(size_t)s * 7 + k];
  sim3_oplus(est, upd, G.fix_scale != 0);
#pragma unroll
  for (int k = 0; k < 8; k++) G.vert[(size_t)v * 8 + k] = est[k];
}
__global__ void k_pg_restore(PgDev G) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= G.n_slot) return;
  const int v = G.slot_vert[s];
#pragma unroll
  for (int k = 0; k < 8; k++) G.vert[(size_t)v * 8 + k] = G.vert_bak[(size_t)v * 8 + k];
}

}  // namespace sqrtba

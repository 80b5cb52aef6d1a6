// Microbenchmark: how fast can cp.async.bulk (1-D TMA) stream a DRAM-resident buffer into shared memory?
// Sweeps chunk size, ring depth and CTAs/SM.  Consumers only touch one word per chunk (no compute), so this is the
// ceiling for the TMA-fed matvec.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)), "l"(s), "r"(n), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t a = smem_u32(b), ok;
  do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(a), "r"(ph) : "memory"); } while (!ok);
}

// chunk = bytes per stage; split = number of bulk copies per stage
__global__ void k_stream(const char* src, size_t total, int chunk, int S, int split, double* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = (uint64_t*)(sm + (size_t)S * chunk);
  uint64_t* empty = full + S;
  const int tid = threadIdx.x;
  const size_t nchunk = total / chunk;
  const size_t c0 = nchunk * blockIdx.x / gridDim.x, c1 = nchunk * (blockIdx.x + 1) / gridDim.x;
  if (tid == 0) { for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (tid >= 128) {
    if (tid == 128) {
      int n = 0;
      for (size_t c = c0; c < c1; c++, n++) {
        int s = n % S;
        if (n >= S) mbar_wait(&empty[s], ((n / S) - 1) & 1);
        mbar_expect_tx(&full[s], chunk);
        int piece = chunk / split;
        for (int i = 0; i < split; i++) bulk(sm + (size_t)s * chunk + i * piece, src + c * chunk + i * piece, piece, &full[s]);
      }
    }
    return;
  }
  double acc = 0;
  int n = 0;
  for (size_t c = c0; c < c1; c++, n++) {
    int s = n % S;
    mbar_wait(&full[s], (n / S) & 1);
    acc += ((const double*)(sm + (size_t)s * chunk))[tid];
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid == 0) mbar_arrive(&empty[s]);
  }
  if (acc == 12345.678) sink[0] = acc;
}

int main() {
  const size_t total = (size_t)2 << 30;
  char* d; double* sink;
  cudaMalloc(&d, total); cudaMalloc(&sink, 8);
  cudaMemset(d, 1, total);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int nsm = 148;
  printf("chunkKB S split ctas/SM  GB/s\n");
  for (int chunk : {8192, 16384, 28672}) for (int S : {2, 3, 4, 6}) for (int split : {1, 4}) for (int per : {1, 2, 3, 4, 6}) {
    size_t smem = (size_t)S * chunk + 2 * S * 8;
    if (smem * per > 220 * 1024) continue;
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stream, 160, smem);
    if (occ < per) continue;
    k_stream<<<nsm * per, 160, smem>>>(d, total, chunk, S, split, sink);
    cudaEventRecord(e0);
    k_stream<<<nsm * per, 160, smem>>>(d, total, chunk, S, split, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    printf("%5d %2d %3d %3d   %8.1f %s\n", chunk / 1024, S, split, per, total / (ms * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
  }
  return 0;
}

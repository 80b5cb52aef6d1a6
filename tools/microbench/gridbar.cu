// Grid-barrier latency on B200: how long does one device-wide barrier take as a function of the number of CTAs and
// of the polling style?  (Design input for the persistent PCG kernel.)   nvcc -arch=sm_100a -O3 gridbar.cu -o gridbar
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <int MODE>
__global__ void k(unsigned* bar, int iters, double* sink) {
  unsigned gen = 0;
  double acc = 0;
  for (int i = 0; i < iters; i++) {
    __syncthreads();
    if (threadIdx.x == 0) {
      gen++;
      if (MODE == 0) {  // fence + atomic + acquire spin + fence (cooperative-groups style)
        __threadfence();
        unsigned prev = atomicAdd(&bar[0], 1u);
        if (prev == gridDim.x - 1) { atomicExch(&bar[0], 0u); __threadfence(); atomicExch(&bar[1], gen); }
        else while (ld_acquire_gpu(&bar[1]) != gen) {}
        __threadfence();
      } else if (MODE == 1) {  // release-add on a monotonically increasing counter, relaxed spin, one acquire fence
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&bar[0]) : "memory");
        const unsigned target = gen * gridDim.x;
        while (ld_relaxed_gpu(&bar[0]) < target) {}
        __threadfence();
      } else if (MODE == 2) {  // as 1 with nanosleep backoff
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&bar[0]) : "memory");
        const unsigned target = gen * gridDim.x;
        while (ld_relaxed_gpu(&bar[0]) < target) { __nanosleep(32); }
        __threadfence();
      } else if (MODE == 4) {  // arrivals and the release flag on DIFFERENT lines: the last arriver (it sees it in the
                               // value its atomic returns) publishes the generation, everybody else spins on a
                               // read-only line that no atomic is queued on
        unsigned prev;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&bar[0]) : "memory");
        if (prev == gen * gridDim.x - 1) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&bar[32]), "r"(gen) : "memory");
        else while (ld_acquire_gpu(&bar[32]) < gen) {}
      } else if (MODE == 3) {  // as 1, acquire load instead of relaxed + fence
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&bar[0]) : "memory");
        const unsigned target = gen * gridDim.x;
        while (ld_acquire_gpu(&bar[0]) < target) {}
      }
    }
    __syncthreads();
    acc += i;
  }
  if (sink) sink[blockIdx.x] = acc;
}
template <int MODE>
void run(int grid, int iters, unsigned* bar) {
  cudaMemset(bar, 0, 1024);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  int it = iters;
  double* sink = nullptr;
  void* args[] = {&bar, &it, &sink};
  cudaLaunchCooperativeKernel((void*)k<MODE>, dim3(grid), dim3(160), args, 70 * 1024, 0);
  cudaDeviceSynchronize();
  cudaMemset(bar, 0, 1024);
  cudaEventRecord(a);
  cudaLaunchCooperativeKernel((void*)k<MODE>, dim3(grid), dim3(160), args, 70 * 1024, 0);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  printf("mode %d grid %4d: %.3f us per barrier (%s)\n", MODE, grid, 1e3 * ms / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  unsigned* bar;
  cudaMalloc(&bar, 1024);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  for (int grid : {148, 296, 444}) {
    run<0>(grid, 2000, bar);
    run<1>(grid, 2000, bar);
    run<2>(grid, 2000, bar);
    run<3>(grid, 2000, bar);
    run<4>(grid, 2000, bar);
  }
  return 0;
}

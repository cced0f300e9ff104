// Probe: cost of a hand-rolled grid barrier (atomic arrive + spin) on B200 for the grid sizes the single-launch BatchNorm
// kernels use, with three polling flavours, and of an empty kernel for reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/bin/grid_barrier_probe scripts/grid_barrier_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_volatile(const unsigned* p) {
  unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// counter is monotonic: barrier k completes when it reaches (k + 1) * gridDim.x
template <int MODE>
__global__ void __launch_bounds__(256) barrier_kernel(unsigned* counter, int reps, unsigned long long* cycles) {
  const long long t0 = clock64();
  for (int k = 0; k < reps; ++k) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned target = (unsigned)(k + 1) * gridDim.x;
      __threadfence();
      if (MODE == 3) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory"); }
      else atomicAdd(counter, 1u);
      if (MODE == 0) { while (ld_acquire(counter) < target) __nanosleep(64); }
      else if (MODE == 1) { while (ld_relaxed(counter) < target) { } }
      else if (MODE == 2) { while (ld_volatile(counter) < target) { } }
      else { while (ld_relaxed(counter) < target) __nanosleep(20); }
      __threadfence();
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = (unsigned long long)(clock64() - t0);
}

__global__ void empty_kernel() {}

int main() {
  unsigned* counter; unsigned long long* cyc;
  cudaMalloc(&counter, 4); cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 20; ++i) empty_kernel<<<148, 256>>>();
  cudaEventRecord(e0);
  for (int i = 0; i < 200; ++i) empty_kernel<<<148, 256>>>();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("empty kernel, back to back: %.2f us per launch\n", ms * 1e3 / 200);
  const int reps = 200;
  for (int grid : {64, 128, 148, 296, 592, 1184}) {
    for (int mode = 0; mode < 4; ++mode) {
      cudaMemset(counter, 0, 4);
      unsigned long long h = 0;
      cudaEventRecord(e0);
      if (mode == 0) barrier_kernel<0><<<grid, 256>>>(counter, reps, cyc);
      else if (mode == 1) barrier_kernel<1><<<grid, 256>>>(counter, reps, cyc);
      else if (mode == 2) barrier_kernel<2><<<grid, 256>>>(counter, reps, cyc);
      else barrier_kernel<3><<<grid, 256>>>(counter, reps, cyc);
      cudaEventRecord(e1);
      cudaError_t err = cudaEventSynchronize(e1);
      if (err != cudaSuccess) { printf("grid %d mode %d: %s\n", grid, mode, cudaGetErrorString(err)); return 1; }
      cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const char* names[4] = {"acquire+nanosleep(64)", "relaxed spin", "volatile spin", "red.release + relaxed + nanosleep(20)"};
      printf("grid %4d  %-38s %.2f us per barrier (%.0f cycles)\n", grid, names[mode], ms * 1e3 / reps, (double)h / reps);
    }
  }
  return 0;
}

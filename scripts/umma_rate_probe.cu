// Probe: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands) as a function of N, swizzle mode and
// whether consecutive MMAs hit the same accumulator.  One CTA per SM (148 CTAs) or a single CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/bin/umma_rate_probe scripts/umma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../deep-calcium_b200/csrc/tc_common.cuh"

namespace dcb {
unsigned long long g_launches = 0;
char* last_error_buf() { static char b[512]; return b; }
int fail(int code, const char*, ...) { return code; }
int sm_count() { return 148; }
}
using namespace dcb::tc;

// smem: A tile 256 rows x 128 B (32 KB), B tile 256 rows x 128 B (32 KB); contents irrelevant (zeros)
__global__ void rate_kernel(int N, int BK, int iters, int alt_acc, int row_shift, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 80 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t pitch = BK * 2, sbo = 8 * pitch, swz = BK == 64 ? SWZ_128B : SWZ_64B;
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t sa = smem_u32(smem) + row_shift * pitch, sb = smem_u32(smem) + 40 * 1024;
    const int ksteps = BK / 16;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + ((alt_acc && (it & 1)) ? 256 : 0);
      for (int k = 0; k < ksteps; ++k)
        umma_bf16(d, make_smem_desc(sa + k * 32, 16, sbo, swz), make_smem_desc(sb + k * 32, 16, sbo, swz), idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) *cycles_out = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000;
  for (int grid : {1, 148}) {
    for (int BK : {64, 32}) {
      for (int N : {32, 64, 128, 256}) {
        for (int alt : {0, 1}) {
          for (int shift : {0, 1}) {
            if (alt == 1 && N > 256) continue;
            rate_kernel<<<grid, 128, 90 * 1024>>>(N, BK, iters, alt, shift, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            const double per = (double)c / (iters * (BK / 16));
            printf("grid=%3d BK=%2d N=%3d alt_acc=%d row_shift=%d : %.1f cycles per UMMA (128xNx16) -> %.0f MAC/cycle/SM\n", grid, BK, N,
                   alt, shift, per, 128.0 * N * 16 / per);
          }
        }
      }
    }
  }
  return 0;
}

// Probe: why is the dy-folded issue pattern (N = 96, accumulators shifted by 32 TMEM columns per input row) slow?
// Times several issue patterns of tcgen05.mma (kind::f16, M=128, K=16, SW64 operands, BK=32) from one thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/bin/umma_fold_probe scripts/umma_fold_probe.cu
#include <cstdio>
#include <cstdlib>
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

// variant: 0 standard (18 x N=32 per tile, 3 A rows dy-innermost, 4 rotating accumulators)
//          1 fold (6 x N=96 per row, same A row, D advances 32 columns per row)
//          2 fold, D fixed
//          3 fold, D advances 96 columns per row (no overlap between consecutive rows)
//          4 fold, D advances 32 columns, consecutive MMAs alternate between two A rows (two input rows interleaved)
//          5 standard order but N=96 into a fixed D (18 MMAs)
//          6 fold, D advances 32, N = 32 only (one accumulator per MMA)
//          7 fold, D advances 32 columns per row, but a different A row region per k/dx (no shared A rows)
__global__ void probe_kernel(int variant, int iters, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 150 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t pitch = 64, sbo = 512, rowb = 9 * 1024;        // halo row slot: 130 px x 64 B -> 9 KB
    const uint64_t dbase = make_smem_desc(0, 16, sbo, SWZ_64B);
    const uint32_t a16 = smem_u32(smem) >> 4, w16 = (smem_u32(smem) + 120 * 1024) >> 4, rowb16 = rowb >> 4, pitch16 = pitch >> 4;
    const uint32_t wblk16 = (32 * 64) >> 4;
    const uint32_t id32 = make_idesc_bf16(128, 32, 0, 0), id96 = make_idesc_bf16(128, 96, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (variant == 0 || variant == 5) {
        const uint32_t d = variant == 0 ? tmem_base + (it & 3) * 32 : tmem_base;
        const uint32_t r0 = a16 + (it % 10) * rowb16;
        uint32_t accf = 0;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              umma_bf16(d, dbase + (r0 + dy * rowb16 + dx * pitch16 + 2 * k), dbase + (w16 + (dy * 3 + dx) * wblk16 + 2 * k),
                        variant == 0 ? id32 : id96, accf);
              accf = 1;
            }
      } else {
        uint32_t dcol = 0;
        if (variant == 1 || variant == 4 || variant == 6 || variant == 7) dcol = (it % 13) * 32;
        if (variant == 3) dcol = (it % 4) * 96;
        const uint32_t d = tmem_base + dcol;
        const uint32_t r0 = a16 + (it % 10) * rowb16;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            uint32_t ar = r0 + dx * pitch16 + 2 * k;
            uint32_t dd = d;
            if (variant == 4 && (k & 1)) { ar += rowb16; dd += 32; }
            if (variant == 7) ar = r0 + (dx * 2 + k) * (rowb16 / 8) * 1 + 2 * k;   // 1152 B apart: disjoint 128-row windows? no - just shifted further
            umma_bf16(dd, dbase + ar, dbase + (w16 + dx * 3 * wblk16 + 2 * k), variant == 6 ? id32 : id96, 1u);
          }
      }
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
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000;
  const char* names[8] = {"standard 18 x N=32, 3 rows dy-innermost, 4 accumulators", "fold 6 x N=96, D += 32 cols per row",
                          "fold 6 x N=96, D fixed", "fold 6 x N=96, D += 96 cols per row",
                          "fold, two input rows interleaved (D alternates +0/+32)", "standard order, N=96, D fixed (18 MMAs)",
                          "fold pattern with N=32, D += 32 per row", "fold, D += 32, A windows spread"};
  for (int v = 0; v < 8; ++v) {
    probe_kernel<<<148, 128, 190 * 1024>>>(v, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const int per = (v == 0 || v == 5) ? 18 : 6;
    printf("variant %d (%s): %.1f cycles per row/tile, %.1f per MMA\n", v, names[v], (double)c / iters, (double)c / iters / per);
  }
  return 0;
}

// Probe for round 2: does a CTA PAIR (cluster of 2, tcgen05 cta_group::2) compute D[256 x N] = A[256 x K] * B[N x K]^T
// correctly with each CTA holding its 128 rows of A and HALF of the N rows of B in its own shared memory?
// Minimal on purpose: operands are written to shared memory by ordinary threads in the canonical K-major SWIZZLE_128B
// layout (no TMA), one K block of 64 (four MMAs), leader CTA issues, commit multicast to both CTAs, each CTA drains its
// own 128 TMEM lanes.  Checked against a CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/bin/umma_cta2_probe scripts/umma_cta2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cooperative_groups.h>
#include "../deep-calcium_b200/csrc/tc_common.cuh"

namespace dcb {
unsigned long long g_launches = 0;
char* last_error_buf() { static char b[512]; return b; }
int fail(int code, const char*, ...) { return code; }
int sm_count() { return 148; }
}
using namespace dcb::tc;
namespace cg = cooperative_groups;

constexpr int K = 64, N = 128;

// element (row, k) of a K-major tile with 128-byte rows, SWIZZLE_128B: 16-byte chunk index XOR (row % 8)
__device__ __forceinline__ uint32_t swz128_off(int row, int k) {
  const int chunk = (k >> 3) ^ (row & 7);
  return (uint32_t)row * 128u + (uint32_t)chunk * 16u + (uint32_t)(k & 7) * 2u;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
cta2_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                  // [128 rows][64 k] bf16 = 16 KB
  uint8_t* sB = smem + 16384;          // [64 rows (this CTA's half of N)][64 k] = 8 KB
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + swz128_off(r, k)) = A[(size_t)(rank * 128 + r) * K + k];
  }
  for (int i = threadIdx.x; i < (N / 2) * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + swz128_off(r, k)) = B[(size_t)(rank * (N / 2) + r) * K + k];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_fence_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
    const uint64_t da = make_smem_desc(smem_u32(sA), 16, 1024, SWZ_128B);
    const uint64_t db = make_smem_desc(smem_u32(sB), 16, 1024, SWZ_128B);
    for (int k = 0; k < K / 16; ++k) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem_base), "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(k > 0 ? 1u : 0u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar_done)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  // each CTA drains its own 128 lanes: row = rank*128 + lane index
  const int row = rank * 128 + warp * 32 + lane;
  for (int c = 0; c < N; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
  }
}

// issue rate of the pair MMA: the leader issues `iters` x 4 MMAs (M = 256, K = 16) on zeroed operands
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
cta2_rate_kernel(int Nn, int iters, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_fence_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(256, Nn, 0, 0);
    const uint64_t da = make_smem_desc(smem_u32(smem), 16, 1024, SWZ_128B);
    const uint64_t db = make_smem_desc(smem_u32(smem) + 16384, 16, 1024, SWZ_128B);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + ((it & 1) ? 256u : 0u);
      for (int k = 0; k < 4; ++k)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d), "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar_done)), "h"((uint16_t)3) : "memory");
    mbar_wait(&bar_done, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cycles_out = t1 - t0;
  } else {
    mbar_wait(&bar_done, 0);
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

int main() {
  std::vector<__nv_bfloat16> hA(256 * K), hB(N * K);
  std::vector<float> fA(256 * K), fB(N * K), ref(256 * N), out(256 * N);
  srand(7);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.f); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0; for (int k = 0; k < K; ++k) s += fA[m * K + k] * fB[n * K + k];
      ref[m * N + n] = s;
    }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, out.size() * 4);
  cudaFuncSetAttribute(cta2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  cta2_kernel<<<2, 128, 26 * 1024>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (size_t i = 0; i < out.size(); ++i) {
    const double d = fabs((double)out[i] - ref[i]);
    if (!(d <= 1e-3)) ++bad;
    if (d > maxerr || d != d) maxerr = d;
  }
  printf("cta_group::2 M=256 N=%d K=%d: max |err| = %g, mismatches = %d of %zu  (rows 0..127 = CTA 0, 128..255 = CTA 1)\n", N, K,
         maxerr, bad, out.size());
  if (bad) {
    for (int m : {0, 127, 128, 255}) printf("  row %3d: got %g %g %g ... ref %g %g %g\n", m, out[m * N], out[m * N + 1], out[m * N + 64],
                                            ref[m * N], ref[m * N + 1], ref[m * N + 64]);
  }
  if (bad) return 2;
  cudaFuncSetAttribute(cta2_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  long long* dc; cudaMalloc(&dc, 8);
  for (int grid : {2, 148}) {
    for (int Nn : {64, 128, 256}) {
      const int iters = 2000;
      cta2_rate_kernel<<<grid, 128, 50 * 1024>>>(Nn, iters, dc);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error (rate): %s\n", cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
      const double per = (double)c / (iters * 4);
      printf("grid=%3d pair MMA M=256 N=%3d K=16: %.1f cycles per MMA -> %.0f MAC/cycle/SM (each of the two SMs)\n", grid, Nn, per,
             256.0 * Nn * 16 / per / 2);
    }
  }
  return 0;
}

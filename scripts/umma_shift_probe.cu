// Probe: can a tcgen05.mma A operand start at a row offset inside a TMA-written swizzled tile?
// (needed to reuse one halo tile for the three horizontal taps of a 3x3 convolution)
//   A_g [R rows][BK] bf16 -> TMA box {BK, R} with 128B / 64B swizzle -> smem (1024-aligned)
//   B = identity [BK][BK]; D[m][n] = A[m + row_off][n] is expected.
// For each (swizzle, row_off, base-offset rule) print whether D matches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/bin/umma_shift_probe scripts/umma_shift_probe.cu
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

constexpr int R = 144;   // rows in the smem tile (>= 128 + max shift, multiple of 8)

__global__ void probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int BK,
                             int row_off, int base_rule, float* D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + 32768;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t pitch = BK * 2;
  if (threadIdx.x == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, R * pitch + BK * pitch);
    tma_load_2d(&mapA, &bar_load, sa, 0, 0);
    tma_load_2d(&mapB, &bar_load, sb, 0, 0);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t swz = BK == 64 ? SWZ_128B : SWZ_64B;
    const uint32_t sbo = 8 * pitch;
    const uint32_t idesc = make_idesc_bf16(128, BK, 0, 0);
    for (int k = 0; k < BK / 16; ++k) {
      const uint32_t a_addr = smem_u32(sa) + row_off * pitch + k * 32;
      uint64_t da = make_smem_desc(a_addr, 16, sbo, swz);
      uint32_t bo = 0;
      if (base_rule == 1) bo = (a_addr >> 7) & 7;
      else if (base_rule == 2) bo = row_off & 7;
      else if (base_rule == 3) bo = (a_addr >> 6) & 7;
      da |= (uint64_t)bo << 49;
      const uint64_t db = make_smem_desc(smem_u32(sb) + k * 32, 16, sbo, swz);
      umma_bf16(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar_mma);
  }
  __syncthreads();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int c = 0; c < BK; c += 32) {
    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * BK + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int BK : {64, 32}) {
    std::vector<__nv_bfloat16> A(R * BK), B(BK * BK);
    for (int r = 0; r < R; ++r) for (int c = 0; c < BK; ++c) A[r * BK + c] = __float2bfloat16((float)(((r * 7 + c * 3) % 17) - 8));
    for (int n = 0; n < BK; ++n) for (int k = 0; k < BK; ++k) B[n * BK + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * BK * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap mA, mB;
    cuuint64_t dimsA[2] = {(cuuint64_t)BK, (cuuint64_t)R}, strA[1] = {(cuuint64_t)BK * 2};
    cuuint32_t boxA[2] = {(cuuint32_t)BK, (cuuint32_t)R}, es[2] = {1, 1};
    cuuint64_t dimsB[2] = {(cuuint64_t)BK, (cuuint64_t)BK};
    cuuint32_t boxB[2] = {(cuuint32_t)BK, (cuuint32_t)BK};
    CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r1 = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, strA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, strA, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 || r2) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
    std::vector<float> D(128 * BK);
    for (int row_off = 0; row_off <= 10; ++row_off) {
      for (int rule = 0; rule < 4; ++rule) {
        cudaMemset(dD, 0, 128 * BK * 4);
        probe_kernel<<<1, 128, 48 * 1024>>>(mA, mB, BK, row_off, rule, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("BK=%d off=%d rule=%d CUDA error %s\n", BK, row_off, rule, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < BK; ++n)
          if (D[m * BK + n] != __bfloat162float(A[(m + row_off) * BK + n])) ++bad;
        printf("BK=%d (swizzle %dB) row_off=%2d base_rule=%d -> %s (%d mismatches)\n", BK, BK * 2, row_off, rule,
               bad ? "WRONG" : "ok", bad);
      }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}

// Probe: fixed per-row overheads of the folded strip loop in the single MMA-issuing thread:
// tcgen05.commit, mbarrier try_wait on an already-complete phase, tcgen05.fence::after_thread_sync.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/bin/umma_overhead_probe scripts/umma_overhead_probe.cu
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

// both phase checks in flight at once
__device__ __forceinline__ void dual_wait_a(uint32_t bar0, uint32_t par0, uint32_t bar1, uint32_t par1) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred P, Q;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 Q, [%3], %4;\n\tand.pred P, P, Q;\n\tselp.b32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok) : "r"(bar0), "r"(par0), "r"(bar1), "r"(par1) : "memory");
  } while (!ok);
}
// plain shared-memory flag polling
__device__ __forceinline__ void flag_wait_a(uint32_t addr, uint32_t want) {
  uint32_t v;
  do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); } while (v != want);
}
__device__ __forceinline__ void test_wait_a(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
  } while (!ok);
}

// flags: 1 = two commits per row, 2 = two ready try_waits per row, 4 = fence after the waits, 8 = no MMAs,
//        16 = one commit per row instead of two, 32 = commits every 4th row only, 64 = precomputed barrier addresses
__global__ void probe_kernel(int flags, int iters, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar, bars[16], ready[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t flags_s[2];
  if (threadIdx.x == 0) { flags_s[0] = 0; flags_s[1] = 0; }
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 150 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
    mbar_init(&ready[0], 1); mbar_init(&ready[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t pitch = 64, sbo = 512, rowb = 9 * 1024;
    const uint64_t dbase = make_smem_desc(0, 16, sbo, SWZ_64B);
    const uint32_t a16 = smem_u32(smem) >> 4, w16 = (smem_u32(smem) + 120 * 1024) >> 4, rowb16 = rowb >> 4, pitch16 = pitch >> 4;
    const uint32_t wblk16 = (32 * 64) >> 4;
    const uint32_t id96 = make_idesc_bf16(128, 96, 0, 0);
    const uint32_t a_flag = smem_u32_pinned(&flags_s[0]);
    const uint32_t a_ready0 = smem_u32_pinned(&ready[0]), a_ready1 = smem_u32_pinned(&ready[1]), a_bars = smem_u32_pinned(&bars[0]);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if ((flags & 2) && (flags & 256)) {
        dual_wait_a(a_ready0, 1, a_ready1, 1);
      } else if ((flags & 2) && (flags & 512)) {
        flag_wait_a(a_flag, 0); flag_wait_a(a_flag + 4, 0);
      } else if ((flags & 2) && (flags & 128)) {
        test_wait_a(a_ready0, 1);
        test_wait_a(a_ready1, 1);
      } else if ((flags & 2) && (flags & 64)) {
        mbar_wait_a(a_ready0, 1);
        mbar_wait_a(a_ready1, 1);
      } else if (flags & 2) {
        mbar_wait(&ready[0], 1);      // fresh barrier: the phase with parity 1 counts as complete
        mbar_wait(&ready[1], 1);
      }
      if (flags & 4) tc_fence_after();
      if (!(flags & 8)) {
        const uint32_t d = tmem_base + (it % 13) * 32;
        const uint32_t r0 = a16 + (it % 10) * rowb16;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16(d, dbase + (r0 + dx * pitch16 + 2 * k), dbase + (w16 + dx * 3 * wblk16 + 2 * k), id96, 1u);
        umma_bf16(d, dbase + r0, dbase + w16, id96, 1u);
      }
      if ((flags & 1) && (!(flags & 32) || (it & 3) == 3)) {
        if (flags & 64) {
          umma_commit_a(a_bars + 8 * (it & 15));
          if (!(flags & 16)) umma_commit_a(a_bars + 8 * ((it + 8) & 15));
        } else {
          umma_commit(&bars[it & 15]);
          if (!(flags & 16)) umma_commit(&bars[(it + 8) & 15]);
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
  const int variants[] = {0, 1, 1 | 16, 1 | 32, 2, 2 | 4, 1 | 2 | 4, 8 | 1, 8 | 2, 8 | 2 | 4, 8 | 1 | 2 | 4,
                          64 | 1, 64 | 2, 64 | 1 | 2 | 4, 64 | 8 | 1, 64 | 8 | 2, 64 | 8 | 1 | 2 | 4,
                          128 | 64 | 2, 128 | 64 | 8 | 2, 128 | 64 | 1 | 2 | 4,
                          256 | 64 | 8 | 2, 256 | 64 | 2, 256 | 64 | 1 | 2 | 4, 512 | 64 | 8 | 2, 512 | 64 | 2, 512 | 64 | 1 | 2 | 4};
  for (int v : variants) {
    probe_kernel<<<148, 128, 190 * 1024>>>(v, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("flags %3d [%s%s%s%s%s%s%s]: %.1f cycles per row\n", v, (v & 512) ? "LDS flag polling, " : (v & 256) ? "dual try_wait, " : (v & 128) ? "test_wait, " : (v & 64) ? "hoisted addresses, " : "", (v & 8) ? "no MMA " : "7 MMAs ", (v & 1) ? "+commits " : "",
           (v & 16) ? "(one) " : "", (v & 32) ? "(every 4th row) " : "", (v & 2) ? "+2 ready waits " : "", (v & 4) ? "+fence" : "",
           (double)c / iters);
  }
  return 0;
}

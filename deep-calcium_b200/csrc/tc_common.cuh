// sm_100a primitives used by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences), UMMA shared-memory and instruction descriptors.
// Written as inline PTX; the bit layouts follow the PTX ISA "tcgen05" tables.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace dcb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// 32-bit shared address of `p` that the compiler cannot rematerialise (an opaque mov): the plain smem_u32() is a pure
// expression that ptxas happily recomputes - S2R SR_CgaCtaId included - at every use inside a loop.
__device__ __forceinline__ uint32_t smem_u32_pinned(const void* p) {
  uint32_t a;
  asm volatile("mov.u32 %0, %1;" : "=r"(a) : "r"(smem_u32(p)));
  return a;
}
// Variants taking a precomputed 32-bit shared-memory address.  smem_u32() of a __shared__ variable re-reads
// SR_CgaCtaId (S2UR, ~100+ cycles of latency) every time it is evaluated next to a volatile asm; a single-thread
// producer / MMA loop that waits and commits once per tile pays for that on its critical path
// (profiles/r1_umma_overhead_probe.log), so those loops hoist the addresses out.
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_a(bar_addr, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (surfacing as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}

// ---------------------------------------------------------------- TMA loads (tile mode)
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------- TMA stores (tile mode, bulk async group)
// The epilogues stage a [pixels x 32 channels] block in shared memory (64-byte rows, SWIZZLE_64B so that the 16-byte
// chunk writes of consecutive lanes fall into different banks) and hand it to the TMA unit: the store then leaves the
// SM as full lines through the async proxy instead of 32 scattered 32-byte sectors per warp instruction through the
// LSU pipe (which the short-K layers saturate: profiles/r2_convT_ncu.txt).  Out-of-bounds parts of a box are clipped.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all earlier bulk groups of this thread have finished READING their shared-memory source (it may be overwritten)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// one pixel's 32 packed channels (64 bytes) into row `m` of a SWIZZLE_64B staging tile at shared address `tile`
// (512-byte aligned): 16-byte chunk j lands at chunk position j ^ ((m >> 1) & 3)
__device__ __forceinline__ void stage_pk16_swz64(uint32_t tile, int m, const uint32_t (&pk)[16]) {
  const uint32_t row = tile + (uint32_t)m * 64u, x = ((uint32_t)m >> 1) & 3u;
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j) st_shared_v4(row + ((j ^ x) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by one thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit_a(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets lane (base_lane + i), registers = consecutive columns
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (PTX ISA, tcgen05 "matrix descriptor"):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//   [46,48) version = 1 (sm_100)   [49,52) base offset   [52] LBO mode   [61,64) swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
enum { SWZ_NONE = 0, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6 };
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t swz) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(swz & 7) << 61;
  return d;
}
// Instruction descriptor for kind::f16 (PTX ISA "instruction descriptor"):
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format (1 = bf16)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
// f16 != 0: fp16 operands (format code 0) instead of bf16 (format code 1); same MMA rate, 10 instead of 7 mantissa bits
__host__ __device__ constexpr uint32_t make_idesc_16(int M, int N, int a_mn_major, int b_mn_major, int f16) {
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return make_idesc_16(M, N, a_mn_major, b_mn_major, 0);
}

// ---------------------------------------------------------------- 16-bit element helpers (bf16 / fp16 by a runtime flag)
__device__ __forceinline__ uint16_t cvt16(float v, int f16) {
  if (f16) { const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t*>(&h); }
  const __nv_bfloat16 b = __float2bfloat16_rn(v);
  return *reinterpret_cast<const uint16_t*>(&b);
}
__device__ __forceinline__ float2 unpack16x2(uint32_t u, int f16) {
  if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b, int f16) {
  if (f16) {
    const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&m);
  }
  const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&m);
}

// ---------------------------------------------------------------- epilogue stores
// 32-byte store (sm_100, PTX 8.8): one FULL 32-byte sector per lane.  The epilogues write 64 B (bf16) or 128 B (fp32)
// per pixel and lane; as 16-byte stores every warp instruction touched 32 sectors half-filled - twice the
// store requests for the same bytes.  p must be 32-byte aligned.
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                             uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
               "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}
// 16 packed bf16 pairs (32 channels, 64 bytes) of one pixel
__device__ __forceinline__ void store_pk16(void* p, const uint32_t (&pk)[16]) {
  st_global_v8(p, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
  st_global_v8(reinterpret_cast<uint8_t*>(p) + 32, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
}

// ---------------------------------------------------------------- epilogue math
// packed fp32x2 FMA (sm_100): d = a * b + c on two lanes at once
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ua, ub, uc, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(uc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
  return d;
}

// y = [relu](acc * scale + shift) for 32 consecutive accumulator columns, packed to 16 bf16 pairs.  scale/shift are
// read from (16-byte aligned) shared memory as float4; ReLU is applied to the packed bf16 pairs (identical result:
// rounding is monotonic and keeps the sign).
__device__ __forceinline__ void bn_relu_pack32(const uint32_t (&r)[32], const float* s_sc, const float* s_sh, int relu,
                                               uint32_t (&pk)[16], int f16 = 0) {
  const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
  if (f16) {
    const __half2 hzero2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 sc = *reinterpret_cast<const float4*>(s_sc + 4 * g);
      const float4 sh = *reinterpret_cast<const float4*>(s_sh + 4 * g);
      const float2 v0 = ffma2(make_float2(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1])), make_float2(sc.x, sc.y),
                              make_float2(sh.x, sh.y));
      const float2 v1 = ffma2(make_float2(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])), make_float2(sc.z, sc.w),
                              make_float2(sh.z, sh.w));
      __half2 b0 = __float22half2_rn(v0), b1 = __float22half2_rn(v1);
      if (relu) { b0 = __hmax2(b0, hzero2); b1 = __hmax2(b1, hzero2); }
      pk[2 * g] = *reinterpret_cast<uint32_t*>(&b0);
      pk[2 * g + 1] = *reinterpret_cast<uint32_t*>(&b1);
    }
    return;
  }
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 sc = *reinterpret_cast<const float4*>(s_sc + 4 * g);
    const float4 sh = *reinterpret_cast<const float4*>(s_sh + 4 * g);
    const float2 v0 = ffma2(make_float2(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1])), make_float2(sc.x, sc.y),
                            make_float2(sh.x, sh.y));
    const float2 v1 = ffma2(make_float2(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])), make_float2(sc.z, sc.w),
                            make_float2(sh.z, sh.w));
    __nv_bfloat162 b0 = __float22bfloat162_rn(v0), b1 = __float22bfloat162_rn(v1);
    if (relu) { b0 = __hmax2(b0, zero2); b1 = __hmax2(b1, zero2); }
    pk[2 * g] = *reinterpret_cast<uint32_t*>(&b0);
    pk[2 * g + 1] = *reinterpret_cast<uint32_t*>(&b1);
  }
}

// the same for two accumulator rows at once (the two rows of a folded strip pair): one scale / shift fetch serves both,
// and the two rows are independent dependency chains
__device__ __forceinline__ void bn_relu_pack32_x2(const uint32_t (&ra)[32], const uint32_t (&rb)[32], const float* s_sc,
                                                  const float* s_sh, int relu, uint32_t (&pka)[16], uint32_t (&pkb)[16],
                                                  int f16) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 sc = *reinterpret_cast<const float4*>(s_sc + 4 * g);
    const float4 sh = *reinterpret_cast<const float4*>(s_sh + 4 * g);
    const float2 sc0 = make_float2(sc.x, sc.y), sc1 = make_float2(sc.z, sc.w);
    const float2 sh0 = make_float2(sh.x, sh.y), sh1 = make_float2(sh.z, sh.w);
    const float2 a0 = ffma2(make_float2(__uint_as_float(ra[4 * g]), __uint_as_float(ra[4 * g + 1])), sc0, sh0);
    const float2 a1 = ffma2(make_float2(__uint_as_float(ra[4 * g + 2]), __uint_as_float(ra[4 * g + 3])), sc1, sh1);
    const float2 b0 = ffma2(make_float2(__uint_as_float(rb[4 * g]), __uint_as_float(rb[4 * g + 1])), sc0, sh0);
    const float2 b1 = ffma2(make_float2(__uint_as_float(rb[4 * g + 2]), __uint_as_float(rb[4 * g + 3])), sc1, sh1);
    if (f16) {
      const __half2 z = __floats2half2_rn(0.f, 0.f);
      __half2 ha0 = __float22half2_rn(a0), ha1 = __float22half2_rn(a1), hb0 = __float22half2_rn(b0), hb1 = __float22half2_rn(b1);
      if (relu) { ha0 = __hmax2(ha0, z); ha1 = __hmax2(ha1, z); hb0 = __hmax2(hb0, z); hb1 = __hmax2(hb1, z); }
      pka[2 * g] = *reinterpret_cast<uint32_t*>(&ha0); pka[2 * g + 1] = *reinterpret_cast<uint32_t*>(&ha1);
      pkb[2 * g] = *reinterpret_cast<uint32_t*>(&hb0); pkb[2 * g + 1] = *reinterpret_cast<uint32_t*>(&hb1);
    } else {
      const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
      __nv_bfloat162 ha0 = __float22bfloat162_rn(a0), ha1 = __float22bfloat162_rn(a1), hb0 = __float22bfloat162_rn(b0),
                     hb1 = __float22bfloat162_rn(b1);
      if (relu) { ha0 = __hmax2(ha0, z); ha1 = __hmax2(ha1, z); hb0 = __hmax2(hb0, z); hb1 = __hmax2(hb1, z); }
      pka[2 * g] = *reinterpret_cast<uint32_t*>(&ha0); pka[2 * g + 1] = *reinterpret_cast<uint32_t*>(&ha1);
      pkb[2 * g] = *reinterpret_cast<uint32_t*>(&hb0); pkb[2 * g + 1] = *reinterpret_cast<uint32_t*>(&hb1);
    }
  }
}

}  // namespace tc
}  // namespace dcb

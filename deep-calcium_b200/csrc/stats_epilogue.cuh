// Batch statistics of a training-mode conv output taken INSIDE the conv epilogue (Keras BatchNormalization in training
// mode, unet_2d_summary.py:157,165: per-channel mean / biased variance over the batch): the epilogue already holds every
// output value in registers, so the per-channel sums (x, x^2) cost a shared-memory transpose instead of another pass over
// the tensor and a grid barrier (the single-launch BatchNorm kernel of bn_fused.cu).
//
//   per 32 px x 32 ch block : stats_block()       - the packed 16-bit block a warp is about to store goes through a
//                                                   2 KB per-warp scratch tile; lane l sums channel pair (l & 15) over
//                                                   16 of the 32 pixels, the two half-warps meet with one shuffle
//   per CTA, once           : stats_cta_finish()  - per-warp register accumulators -> this CTA's fp32 sums (fixed order) ->
//                                                   64-bit fixed point -> integer atomics into sums_q[0..C) = sum x,
//                                                   sums_q[C..2C) = sum x^2
// Every floating-point sum has a fixed order (lane, block, warp) and the cross-CTA total is an INTEGER sum, so the totals
// are bit-reproducible run to run; there are no floating-point atomics.  The statistics are taken over the ROUNDED (stored)
// values - exactly what a separate statistics pass over the stored tensor would see.
#pragma once
#include "tc_common.cuh"

namespace dcb {
namespace tc {

constexpr int STATS_SCRATCH_WORDS = 512;      // per warp: 32 px x 16 packed pairs

// pk = 32 channels (16 packed pairs) of this lane's pixel; invalid pixels contribute zeros.  s / q accumulate the sums of
// channel pair (lane & 15): .x = even channel, .y = odd channel (both half-warps end up with the same totals).
__device__ __forceinline__ void stats_block(const uint32_t (&pk)[16], bool valid, uint32_t* scratch, int lane, int f16,
                                            float2& s, float2& q) {
  __syncwarp();
  {
    // row = pixel (64 B = four 16-byte chunks); chunk k is stored at k ^ ((lane >> 1) & 3): the eight lanes of a
    // quarter-warp write eight different 16-byte bank groups
    uint4* row = reinterpret_cast<uint4*>(scratch) + lane * 4;
    const int f = (lane >> 1) & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      row[k ^ f] = valid ? make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]) : make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  const int cp = lane & 15, hi = lane >> 4;
  float2 ss = make_float2(0.f, 0.f), qq = make_float2(0.f, 0.f);
  const float2 one = make_float2(1.f, 1.f);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    // the upper half-warp visits its pixels with the parity flipped, so the two halves read different banks
    const int px = hi * 16 + (i ^ hi);
    const uint32_t w = scratch[px * 16 + ((((cp >> 2) ^ (i >> 1)) & 3) << 2) + (cp & 3)];
    const float2 v = unpack16x2(w, f16);
    ss = ffma2(v, one, ss);
    qq = ffma2(v, v, qq);
  }
  ss.x += __shfl_xor_sync(0xffffffffu, ss.x, 16); ss.y += __shfl_xor_sync(0xffffffffu, ss.y, 16);
  qq.x += __shfl_xor_sync(0xffffffffu, qq.x, 16); qq.y += __shfl_xor_sync(0xffffffffu, qq.y, 16);
  s.x += ss.x; s.y += ss.y; q.x += qq.x; q.y += qq.y;
}

// Cross-CTA total WITHOUT a second phase: every CTA converts its fp32 sums to 64-bit fixed point (2^-20 units) and adds them
// to the totals with integer atomics.  Integer addition is associative, so the result does not depend on the arrival order
// (bit-reproducible) - and nothing has to wait for a last CTA, re-read partial rows or fence.  Resolution 1e-6 per CTA
// contribution (the fp32 partials themselves are only good to ~1e-7 relative); range |total| < 8.8e12, i.e. an rms
// activation of ~4000 over 2^19 pixels.  The caller zeroes sums_q[2 * C] before the launch.
constexpr float STATS_Q = 1048576.f;               // 2^20
constexpr double STATS_Q_INV = 1.0 / 1048576.0;
__device__ __forceinline__ void stats_add_q(long long* sums_q, int idx, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(sums_q) + idx, (unsigned long long)__float2ll_rn(v * STATS_Q));
}

// dump: shared memory, [nwarps][nblk][16][4] floats = {s.x, s.y, q.x, q.y} of every epilogue warp's accumulators (written by
// the caller, followed by __syncthreads()).  chan_base(bi) = first channel of accumulator block bi of this CTA (several
// blocks may cover the same channels - the sub-positions of a transposed conv; a block outside [0, C) is ignored).
// Called by all threads of the CTA.
template <typename ChanFn>
__device__ __forceinline__ void stats_cta_finish(const float* dump, int nwarps, int nblk, ChanFn chan_base, int C, long long* sums_q) {
  for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
    float s = 0.f, q = 0.f;
    const int cp = (cc & 31) >> 1, odd = cc & 1;
    bool any = false;
    for (int bi = 0; bi < nblk; ++bi) {
      if (chan_base(bi) != (cc & ~31)) continue;
      any = true;
      for (int w = 0; w < nwarps; ++w) {
        const float* d = dump + ((size_t)(w * nblk + bi) * 16 + cp) * 4;
        s += d[odd]; q += d[2 + odd];
      }
    }
    if (any) { stats_add_q(sums_q, cc, s); stats_add_q(sums_q, C + cc, q); }
  }
}

}  // namespace tc
}  // namespace dcb

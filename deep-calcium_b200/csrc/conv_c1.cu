// First layer of the U-Net: Conv2D(3x3,'same') on the 1-channel summary image
// (unet_2d_summary.py:169-172).  K = 9 is far too small for the tensor pipe; the layer is
// bound by writing its [N,H,W,Cout] output, so it runs on CUDA cores: one thread per pixel,
// weights broadcast from shared memory, 16-byte coalesced stores.
#include "elementwise.cuh"

namespace dcb {
extern unsigned long long g_launches;
void launch_reduce_splits(const float* part, int splits, size_t n, float* out, cudaStream_t st);

// One thread per pixel, weights broadcast from shared memory.  A warp's 32 pixels x COUT channels form ONE
// contiguous run of the NHWC output, so the results are staged through a per-warp shared-memory tile and written
// with fully coalesced 16-byte stores (a direct per-thread store would issue 32 partial-sector requests per
// instruction and is ~3x slower for this write-bound layer).
template <typename T, int COUT>
__global__ void __launch_bounds__(256)
conv3x3_c1_fwd_kernel(const float* __restrict__ x, int N, int H, int W, const float* __restrict__ w,
                      const float* __restrict__ scale, const float* __restrict__ shift, int relu, T* __restrict__ out) {
  __shared__ __align__(16) float ws[9 * COUT];
  __shared__ __align__(16) float sc[COUT], sh[COUT];
  constexpr int ROWB = COUT * (int)sizeof(T);                 // bytes per pixel
  constexpr int PITCH = ROWB + 16;                            // padded row: conflict-free 16-byte accesses
  constexpr int U = ROWB <= 64 ? 2 : 1;                       // pixels per thread (stage tile must fit 48 KB of static smem)
  constexpr int PXW = 32 * U;                                 // pixels per warp and iteration
  __shared__ __align__(16) uint8_t stage[8][PXW * PITCH];
  for (int i = threadIdx.x; i < 9 * COUT; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) { sc[i] = scale ? scale[i] : 1.f; sh[i] = shift ? shift[i] : 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* st = stage[warp];
  const long long M = (long long)N * H * W;
  // The kernel is bound by the shared-memory pipe (ncu: l1tex 90 %): the weight broadcasts dominate, so every thread
  // computes U = 2 pixels (lane and lane + 32 of the warp's 64) per weight load.  Block-uniform trip count.
  for (long long base = ((long long)blockIdx.x * 8 + warp) * PXW; base < M; base += (long long)gridDim.x * 8 * PXW) {
    float v[U][9];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = base + lane + 32 * u;
      const bool mv = m < M;
      const long long mm = mv ? m : 0;
      const int wq = (int)(mm % W), hq = (int)((mm / W) % H);
      const float* img = x + (mm - (long long)hq * W - wq);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int ih = hq + t / 3 - 1, iw = wq + t % 3 - 1;
        v[u][t] = (mv && ih >= 0 && ih < H && iw >= 0 && iw < W) ? img[(long long)ih * W + iw] : 0.f;
      }
    }
    // packed fp32x2 FMAs: two output channels per instruction
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 4) {
      float2 a[U][2];
#pragma unroll
      for (int u = 0; u < U; ++u) { a[u][0] = make_float2(0.f, 0.f); a[u][1] = make_float2(0.f, 0.f); }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 wv = *reinterpret_cast<const float4*>(&ws[t * COUT + c0]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float2 vv = make_float2(v[u][t], v[u][t]);
          a[u][0] = ffma2(vv, make_float2(wv.x, wv.y), a[u][0]);
          a[u][1] = ffma2(vv, make_float2(wv.z, wv.w), a[u][1]);
        }
      }
      const float4 s4 = *reinterpret_cast<const float4*>(&sc[c0]), h4 = *reinterpret_cast<const float4*>(&sh[c0]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float2 r0 = ffma2(a[u][0], make_float2(s4.x, s4.y), make_float2(h4.x, h4.y));
        float2 r1 = ffma2(a[u][1], make_float2(s4.z, s4.w), make_float2(h4.z, h4.w));
        if (relu) { r0.x = fmaxf(r0.x, 0.f); r0.y = fmaxf(r0.y, 0.f); r1.x = fmaxf(r1.x, 0.f); r1.y = fmaxf(r1.y, 0.f); }
        store4<T>(reinterpret_cast<T*>(st + (lane + 32 * u) * PITCH) + c0, make_float4(r0.x, r0.y, r1.x, r1.y));
      }
    }
    __syncwarp();
    // 64 pixels x ROWB bytes = one contiguous block of the output
    const long long npix = (M - base) < PXW ? (M - base) : PXW;
    uint8_t* dst = reinterpret_cast<uint8_t*>(out + base * COUT);
#pragma unroll
    for (int q = 0; q < PXW * ROWB / 16 / 32; ++q) {
      const int idx = q * 32 + lane;                           // 16-byte chunk index within the block
      const int px = idx / (ROWB / 16), ch = idx % (ROWB / 16);
      if (px < npix)
        *reinterpret_cast<uint4*>(dst + (size_t)idx * 16) = *reinterpret_cast<const uint4*>(st + px * PITCH + ch * 16);
    }
    __syncwarp();
  }
}

// dW[t][co] = sum_pixels x[pixel + tap t] * dy[pixel][co].  A warp takes 4 consecutive pixels per step:
// lane l reads 4 channels ((l & 7) * 4 ..) of pixel (l >> 3) - one contiguous 256-byte (bf16) row of dy per
// warp load - and keeps 9 taps x 4 channels of partial sums.  Per-CTA partials, then a fixed-order reduce.
template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_c1_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, int N, int H, int W, int Cout,
                        float* __restrict__ part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int psub = lane >> 3, c4 = (lane & 7) * 4;
  const long long M = (long long)N * H * W;
  const long long per_cta = ((M + gridDim.x - 1) / gridDim.x + 3) & ~3LL;
  const long long m0 = (long long)blockIdx.x * per_cta;
  long long m1 = m0 + per_cta; if (m1 > M) m1 = M;
  __shared__ float red[8][9][32];
  for (int cbase = 0; cbase < Cout; cbase += 32) {
    float acc[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
    const bool cvalid = cbase + c4 < Cout;
    for (long long mb = m0 + warp * 4; mb < m1; mb += 32) {
      const long long m = mb + psub;
      if (m < m1 && cvalid) {
        const int wq = (int)(m % W), hq = (int)((m / W) % H);
        const float* img = x + (m - (long long)hq * W - wq);
        const float4 g = load4<T>(dy + m * Cout + cbase + c4);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ih = hq + t / 3 - 1, iw = wq + t % 3 - 1;
          const float v = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(img + (long long)ih * W + iw) : 0.f;
          acc[t][0] = fmaf(v, g.x, acc[t][0]); acc[t][1] = fmaf(v, g.y, acc[t][1]);
          acc[t][2] = fmaf(v, g.z, acc[t][2]); acc[t][3] = fmaf(v, g.w, acc[t][3]);
        }
      }
    }
    // combine the 4 pixel sub-groups of the warp, then the 8 warps
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[t][j];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (psub == 0) red[warp][t][c4 + j] = v;
      }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * 32; i += blockDim.x) {
      const int t = i / 32, l = i % 32;
      float sacc = 0.f;
      for (int wv = 0; wv < 8; ++wv) sacc += red[wv][t][l];
      if (cbase + l < Cout) part[((size_t)blockIdx.x * 9 + t) * Cout + cbase + l] = sacc;
    }
    __syncthreads();
  }
}

}  // namespace dcb

using namespace dcb;

template <typename T>
static int launch_c1_fwd(const float* x, int N, int H, int W, const float* w, int Cout, const float* scale,
                         const float* shift, int relu, T* out, cudaStream_t st) {
  const long long M = (long long)N * H * W;
  long long grid = (M + 255) / 256;
  if (grid > (long long)sm_count() * 16) grid = (long long)sm_count() * 16;
  switch (Cout) {
    case 8: conv3x3_c1_fwd_kernel<T, 8><<<(int)grid, 256, 0, st>>>(x, N, H, W, w, scale, shift, relu, out); break;
    case 16: conv3x3_c1_fwd_kernel<T, 16><<<(int)grid, 256, 0, st>>>(x, N, H, W, w, scale, shift, relu, out); break;
    case 32: conv3x3_c1_fwd_kernel<T, 32><<<(int)grid, 256, 0, st>>>(x, N, H, W, w, scale, shift, relu, out); break;
    case 64:
      if constexpr (sizeof(T) == 2) { conv3x3_c1_fwd_kernel<T, 64><<<(int)grid, 256, 0, st>>>(x, N, H, W, w, scale, shift, relu, out); break; }
      return fail(DCB_ERR_UNSUPPORTED, "dcb_conv3x3_c1_fwd: Cout=64 is built for bf16 output only");
    default: return fail(DCB_ERR_UNSUPPORTED, "dcb_conv3x3_c1_fwd: Cout=%d unsupported (8,16,32,64)", Cout);
  }
  g_launches += 1;
  DCB_LAUNCH_OK("conv3x3_c1_fwd_kernel");
  return DCB_OK;
}

extern "C" int dcb_conv3x3_c1_fwd(int dtype, const float* x, int N, int H, int W, const float* w, int Cout,
                                  const float* scale, const float* shift, int relu, void* out, dcb_stream_t stream) {
  DCB_CHECK_ARG(x && w && out && N > 0 && H > 0 && W > 0, "dcb_conv3x3_c1_fwd: bad arguments");
  if (dtype == DCB_F32) return launch_c1_fwd<float>(x, N, H, W, w, Cout, scale, shift, relu, (float*)out, (cudaStream_t)stream);
  if (dtype == DCB_BF16)
    return launch_c1_fwd<__nv_bfloat16>(x, N, H, W, w, Cout, scale, shift, relu, (__nv_bfloat16*)out, (cudaStream_t)stream);
  if (dtype == DCB_F16)
    return launch_c1_fwd<__half>(x, N, H, W, w, Cout, scale, shift, relu, (__half*)out, (cudaStream_t)stream);
  return fail(DCB_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
}

static int c1_wgrad_ctas() { return sm_count() * 8; }

extern "C" int dcb_conv3x3_c1_wgrad_workspace_bytes(int Cout, size_t* bytes) {
  DCB_CHECK_ARG(bytes && Cout > 0, "dcb_conv3x3_c1_wgrad_workspace_bytes: bad arguments");
  *bytes = (size_t)c1_wgrad_ctas() * 9 * Cout * sizeof(float);
  return DCB_OK;
}

extern "C" int dcb_conv3x3_c1_wgrad(int dtype, const float* x, const void* dy, int N, int H, int W, int Cout, float* dW,
                                    void* ws, size_t ws_bytes, dcb_stream_t stream) {
  DCB_CHECK_ARG(x && dy && dW && N > 0 && H > 0 && W > 0 && Cout > 0, "dcb_conv3x3_c1_wgrad: bad arguments");
  const int grid = c1_wgrad_ctas();
  const size_t need = (size_t)grid * 9 * Cout * sizeof(float);
  if (!ws || ws_bytes < need) return fail(DCB_ERR_WORKSPACE, "dcb_conv3x3_c1_wgrad: workspace %zu B < %zu B", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DCB_F32) conv3x3_c1_wgrad_kernel<float><<<grid, 256, 0, st>>>(x, (const float*)dy, N, H, W, Cout, (float*)ws);
  else if (dtype == DCB_BF16)
    conv3x3_c1_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, (const __nv_bfloat16*)dy, N, H, W, Cout, (float*)ws);
  else return fail(DCB_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
  DCB_LAUNCH_OK("conv3x3_c1_wgrad_kernel");
  const size_t n = (size_t)9 * Cout;
  launch_reduce_splits((const float*)ws, grid, n, dW, st);
  g_launches += 2;
  DCB_LAUNCH_OK("reduce_splits_kernel");
  return DCB_OK;
}

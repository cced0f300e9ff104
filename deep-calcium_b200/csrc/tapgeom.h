// Geometry of a "tap GEMM": out[m][n] = sum_tap sum_k A(m,tap,k) * B[tap][k][n]
// where m = (image, gh, gw) walks an iteration grid and A gathers the input pixel
// (gh*sy + dy[tap], gw*sx + dx[tap]) of up to two channel-concatenated NHWC sources
// (zero outside the image).  One description covers every dense contraction of the
// UNet2DS graph (deepcalcium/models/neurons/unet_2d_summary.py:154-167):
//   conv3x3 'same' fwd : 9 taps dy,dx in {-1,0,1}, stride 1
//   conv3x3 dgrad      : same geometry with flipped / in-out swapped weights
//   convT2x2 s2 fwd    : 1 tap, output scattered to (2gh+a, 2gw+b), one launch slice per (a,b)
//   convT2x2 s2 dgrad  : 4 taps (a,b), input stride 2
#pragma once

struct TapGeom {
  int N;            // images
  int IH, IW;       // input tensor height/width
  int GH, GW;       // iteration grid per image
  int sy, sx;       // input stride
  int ntaps;
  int dy[9], dx[9];
  int OH, OW;       // output tensor height/width
  int osy, osx;     // output position = g*os + od
  int ody, odx;
  int zsub;         // >1: blockIdx.z selects sub-position (ody,odx)=(z/2,z%2) and weight slice z
  int f16;          // 16-bit element format of the tensor-core path: 0 = bf16, 1 = fp16 (DCB_F16, inference only)
};

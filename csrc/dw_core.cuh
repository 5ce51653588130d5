// Depthwise 3x3 convolution (pad 1, stride 1 | 2), NHWC bf16: the per-thread work of the three kernels in
// depthwise.cu as __host__ __device__ functions, so that csrc/tests/dw_host_test.cu can run exactly this index
// logic on the CPU against the convolution definition (no GPU, no launch) — tests/test_cpu_units.py builds and runs it.
//
// Reference site: torchvision.models.mobilenet_v2's `Conv2d(C, C, 3, stride, 1, groups=C)` behind the legacy
// container entrypoint (train.py:60-68, MODEL_TYPE=mobilenet).  K = 9 per output element: there is no GEMM in a
// depthwise convolution, so this is a SIMT kernel (8 channels = one 16-byte vector per thread, the thread's 72
// filter taps live in registers); the 1x1 expand / project convolutions around it are the tcgen05 kernels.
#pragma once
#include <cuda_bf16.h>
#include <stddef.h>

#define HZ_HD __host__ __device__ __forceinline__

namespace hz {
namespace dw {

struct Geo {
  int N, H, W, C;       // input  [N, H, W, C]
  int Ho, Wo, stride;   // output [N, Ho, Wo, C],  Ho = (H - 1) / stride + 1
};

HZ_HD Geo make_geo(int N, int H, int W, int C, int stride) {
  Geo g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.stride = stride;
  g.Ho = (H + 2 - 3) / stride + 1;
  g.Wo = (W + 2 - 3) / stride + 1;
  return g;
}

struct alignas(16) Vec8 {
  __nv_bfloat162 v[4];
};

HZ_HD void load8(const __nv_bfloat16* p, float* f) {
  const Vec8 t = *reinterpret_cast<const Vec8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = __bfloat1622float2(t.v[i]);
    f[2 * i] = a.x;
    f[2 * i + 1] = a.y;
  }
}

// round 8 floats to bf16, store them, and hand the rounded values back (the BN statistics are taken from what the
// next kernel will actually read)
HZ_HD void store8(__nv_bfloat16* p, float* f) {
  Vec8 t;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    t.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    const float2 a = __bfloat1622float2(t.v[i]);
    f[2 * i] = a.x;
    f[2 * i + 1] = a.y;
  }
  *reinterpret_cast<Vec8*>(p) = t;
}

// filter taps of channels [cv*8, cv*8+8): w is [C][3][3] (torch [C,1,3,3], contiguous) -> wr[tap][channel]
HZ_HD void load_taps(const __nv_bfloat16* w, int cv, float (*wr)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t][i] = __bfloat162float(w[(size_t)(cv * 8 + i) * 9 + t]);
}

// y[n, ho, wo, c] = sum_{r,s} x[n, ho*stride - 1 + r, wo*stride - 1 + s, c] * w[c, r, s]     (p = (n*Ho + ho)*Wo + wo)
HZ_HD void fwd_pixel(const Geo& g, const __nv_bfloat16* x, int p, int cv, const float (*wr)[8], float* acc) {
  const int wo = p % g.Wo;
  const int t = p / g.Wo;
  const int ho = t % g.Ho;
  const int n = t / g.Ho;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int h = ho * g.stride - 1 + r;
    if ((unsigned)h >= (unsigned)g.H) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int w = wo * g.stride - 1 + s;
      if ((unsigned)w >= (unsigned)g.W) continue;
      float f[8];
      load8(x + (((size_t)n * g.H + h) * g.W + w) * g.C + cv * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i] * wr[r * 3 + s][i];
    }
  }
}

// dx[n, h, w, c] = sum over (r, s, ho, wo) with ho*stride - 1 + r == h and wo*stride - 1 + s == w of
//                  dy[n, ho, wo, c] * w[c, r, s]                                            (q = (n*H + h)*W + w)
HZ_HD void dgrad_pixel(const Geo& g, const __nv_bfloat16* dy, int q, int cv, const float (*wr)[8], float* acc) {
  const int w = q % g.W;
  const int t = q / g.W;
  const int h = t % g.H;
  const int n = t / g.H;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hh = h + 1 - r;                     // = ho * stride
    if (hh < 0 || (hh % g.stride) != 0) continue;
    const int ho = hh / g.stride;
    if (ho >= g.Ho) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ww = w + 1 - s;
      if (ww < 0 || (ww % g.stride) != 0) continue;
      const int wo = ww / g.stride;
      if (wo >= g.Wo) continue;
      float f[8];
      load8(dy + (((size_t)n * g.Ho + ho) * g.Wo + wo) * g.C + cv * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i] * wr[r * 3 + s][i];
    }
  }
}

// dw[c, r, s] += dy[n, ho, wo, c] * x[n, ho*stride - 1 + r, wo*stride - 1 + s, c]   for output pixel p
HZ_HD void wgrad_pixel(const Geo& g, const __nv_bfloat16* dy, const __nv_bfloat16* x, int p, int cv, float (*acc)[8]) {
  const int wo = p % g.Wo;
  const int t = p / g.Wo;
  const int ho = t % g.Ho;
  const int n = t / g.Ho;
  float gy[8];
  load8(dy + (size_t)p * g.C + cv * 8, gy);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int h = ho * g.stride - 1 + r;
    if ((unsigned)h >= (unsigned)g.H) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int w = wo * g.stride - 1 + s;
      if ((unsigned)w >= (unsigned)g.W) continue;
      float f[8];
      load8(x + (((size_t)n * g.H + h) * g.W + w) * g.C + cv * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[r * 3 + s][i] += gy[i] * f[i];
    }
  }
}

// Thread layout of all three kernels: 256 threads = (C/8 channel vectors) x (256 / (C/8) row lanes); the
// 256 % (C/8) left-over threads idle.  A thread keeps one channel vector for its whole life (taps and partial sums
// stay in registers) and walks rows  blockIdx*rlanes + rl,  += gridDim*rlanes.
struct Lane {
  int cv, rl, rlanes, nvec;
  bool active;
};
HZ_HD Lane make_lane(int tid, int C) {
  Lane l;
  l.nvec = C >> 3;
  l.rlanes = 256 / l.nvec;
  l.cv = tid % l.nvec;
  l.rl = tid / l.nvec;
  l.active = l.rl < l.rlanes;
  return l;
}

}  // namespace dw
}  // namespace hz

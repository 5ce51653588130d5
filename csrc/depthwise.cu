// Depthwise 3x3 convolution kernels (forward with the BatchNorm statistics in its epilogue, input gradient, weight
// gradient straight into the flat fp32 gradient bucket) for MobileNetV2's inverted-residual blocks
// (reference: torchvision mobilenet_v2 behind train.py:60-68, MODEL_TYPE=mobilenet).
//
// A depthwise convolution has K = 9 multiply-adds per output element and no reuse across channels — it is a
// memory-bound stencil, not a GEMM, so it stays off the tensor cores: one thread owns 8 channels (one 16-byte NHWC
// vector), keeps its 72 filter taps in registers, and walks output rows; the 3x3 neighbourhood re-reads hit L1/L2
// (a whole MobileNetV2 activation at CIFAR shapes is a few MB).  The per-thread index logic lives in dw_core.cuh and
// is unit-tested on the host.  All kernels are PDL-aware like the rest of the step.
#include "common.cuh"
#include "dw_core.cuh"
#include "launchers.h"

namespace hz {

using dw::Geo;
using dw::Lane;

// y = dwconv(x, w);  kStats: sums[0:C] += sum y, sums[C:2C] += sum y^2 (of the bf16-rounded outputs; pre-zeroed)
template <bool kStats>
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ w,
                                                         __nv_bfloat16* __restrict__ y, float* __restrict__ sums,
                                                         const Geo g) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float red[];                 // kStats: [rlanes][nvec][16]
  const Lane l = dw::make_lane(threadIdx.x, g.C);
  const int M = g.N * g.Ho * g.Wo;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (l.active) {
    float wr[9][8];
    dw::load_taps(w, l.cv, wr);
    for (int p = blockIdx.x * l.rlanes + l.rl; p < M; p += gridDim.x * l.rlanes) {
      float acc[8];
      dw::fwd_pixel(g, x, p, l.cv, wr, acc);
      dw::store8(y + (size_t)p * g.C + l.cv * 8, acc);
      if (kStats) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += acc[i]; q[i] += acc[i] * acc[i]; }
      }
    }
  }
  if (kStats) {
    if (l.active) {
      float* mine = red + ((size_t)l.rl * l.nvec + l.cv) * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) { mine[i] = s[i]; mine[8 + i] = q[i]; }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < l.nvec * 16; o += 256) {
      const int v = o >> 4, j = o & 15;
      float t = 0.f;
      for (int k = 0; k < l.rlanes; ++k) t += red[((size_t)k * l.nvec + v) * 16 + j];
      atomicAdd(&sums[(j >> 3) * g.C + v * 8 + (j & 7)], t);
    }
  }
}

// the producing layer's BatchNorm (the expand conv's, whose only consumer is this depthwise conv): its backward sums
// (sum g, sum g*xhat), g = dx * mask(out), are taken from the registers that store dx (see conv_gemm.cu kBnBwd)
struct DwBnBwd {
  const __nv_bfloat16* out;      // BN output (activation mask); null = no activation
  const __nv_bfloat16* yraw;
  const float* mean; const float* invstd;
  float* sums;                   // [2C], pre-zeroed
  int cap6;                      // ReLU6 mask (0 < out < 6) instead of ReLU
};

// dx = dwconv_transposed(dy, w)
template <bool kBn>
__global__ void __launch_bounds__(256) dwconv_dgrad_kernel(const __nv_bfloat16* __restrict__ dy,
                                                           const __nv_bfloat16* __restrict__ w,
                                                           __nv_bfloat16* __restrict__ dx, const Geo g, const DwBnBwd b) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float red[];                 // kBn: [rlanes][nvec][16]
  const Lane l = dw::make_lane(threadIdx.x, g.C);
  const int Q = g.N * g.H * g.W;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (l.active) {
    float wr[9][8];
    dw::load_taps(w, l.cv, wr);
    float mu[8], is[8];
    if (kBn) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { mu[i] = b.mean[l.cv * 8 + i]; is[i] = b.invstd[l.cv * 8 + i]; }
    }
    for (int p = blockIdx.x * l.rlanes + l.rl; p < Q; p += gridDim.x * l.rlanes) {
      float acc[8];
      dw::dgrad_pixel(g, dy, p, l.cv, wr, acc);
      const size_t off = (size_t)p * g.C + l.cv * 8;
      dw::store8(dx + off, acc);                   // (acc now holds the bf16-rounded values)
      if (kBn) {
        float y[8];
        dw::load8(b.yraw + off, y);
        if (b.out != nullptr) {
          float o[8];
          dw::load8(b.out + off, o);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = (o[i] > 0.f && (!b.cap6 || o[i] < 6.f)) ? acc[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += acc[i]; q[i] += acc[i] * (y[i] - mu[i]) * is[i]; }
      }
    }
  }
  if (kBn) {
    if (l.active) {
      float* mine = red + ((size_t)l.rl * l.nvec + l.cv) * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) { mine[i] = s[i]; mine[8 + i] = q[i]; }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < l.nvec * 16; o += 256) {
      const int v = o >> 4, j = o & 15;
      float t = 0.f;
      for (int k = 0; k < l.rlanes; ++k) t += red[((size_t)k * l.nvec + v) * 16 + j];
      atomicAdd(&b.sums[(j >> 3) * g.C + v * 8 + (j & 7)], t);
    }
  }
}

// dw[c][tap] += sum over output pixels (fp32 atomics into the flat gradient bucket: [C][3][3] contiguous)
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const __nv_bfloat16* __restrict__ dy,
                                                           const __nv_bfloat16* __restrict__ x,
                                                           float* __restrict__ dwt, const Geo g) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float red[];                 // [rlanes][nvec][8], reused for each of the 9 taps
  const Lane l = dw::make_lane(threadIdx.x, g.C);
  const int M = g.N * g.Ho * g.Wo;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  if (l.active)
    for (int p = blockIdx.x * l.rlanes + l.rl; p < M; p += gridDim.x * l.rlanes) dw::wgrad_pixel(g, dy, x, p, l.cv, acc);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    if (l.active) {
      float* mine = red + ((size_t)l.rl * l.nvec + l.cv) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) mine[i] = acc[t][i];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < l.nvec * 8; o += 256) {
      const int v = o >> 3, i = o & 7;
      float sum = 0.f;
      for (int k = 0; k < l.rlanes; ++k) sum += red[((size_t)k * l.nvec + v) * 8 + i];
      atomicAdd(&dwt[(size_t)(v * 8 + i) * 9 + t], sum);
    }
    __syncthreads();
  }
}

}  // namespace hz

namespace {
inline int rows_grid(int rows, int C, int rows_per_lane, int cap) {
  const int rlanes = 256 / (C / 8);
  int g = (rows + rlanes - 1) / rlanes;
  g = (g + rows_per_lane - 1) / rows_per_lane;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return g;
}
}  // namespace

extern "C" {

int hz_dwconv_ok(int N, int H, int W, int C, int stride) {
  if (C < 8 || C > 2048 || (C & 7)) return 0;
  if (stride != 1 && stride != 2) return 0;
  if (N < 1 || H < 1 || W < 1) return 0;
  return (long long)N * H * W < (1ll << 30);           // 32-bit row indices
}

// stats: optional [2C] fp32 (stats_is_zero: already cleared, e.g. a slice of the per-step statistics arena)
int hz_dwconv_fwd(const void* x, const void* w, void* y, float* stats, int stats_is_zero, int N, int H, int W, int C,
                  int stride, cudaStream_t st) {
  if (!hz_dwconv_ok(N, H, W, C, stride)) return -1;
  const hz::dw::Geo g = hz::dw::make_geo(N, H, W, C, stride);
  const int grid = rows_grid(g.N * g.Ho * g.Wo, C, 2, 148 * 4);
  cudaError_t e;
  if (stats != nullptr) {
    if (!stats_is_zero) hz::zero_f32(stats, (size_t)2 * C, st);
    e = hz::launch(hz::dwconv_fwd_kernel<true>, dim3(grid), dim3(256), sizeof(float) * 256 * 16, st,
                   (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (__nv_bfloat16*)y, stats, g);
  } else {
    e = hz::launch(hz::dwconv_fwd_kernel<false>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)x,
                   (const __nv_bfloat16*)w, (__nv_bfloat16*)y, (float*)nullptr, g);
  }
  return e == cudaSuccess ? 0 : -2;
}

int hz_dwconv_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int C, int stride, cudaStream_t st) {
  if (!hz_dwconv_ok(N, H, W, C, stride)) return -1;
  const hz::dw::Geo g = hz::dw::make_geo(N, H, W, C, stride);
  const int grid = rows_grid(g.N * g.H * g.W, C, 2, 148 * 4);
  hz::DwBnBwd b{};
  return hz::launch(hz::dwconv_dgrad_kernel<false>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)dy,
                    (const __nv_bfloat16*)w, (__nv_bfloat16*)dx, g, b) == cudaSuccess ? 0 : -2;
}

// depthwise dgrad that also takes the BatchNorm-backward sums of the layer producing the conv's input (struct HzBnBwd)
int hz_dwconv_dgrad_bnbwd(const void* dy, const void* w, void* dx, int N, int H, int W, int C, int stride,
                          const struct HzBnBwd* bnb, cudaStream_t st) {
  if (!hz_dwconv_ok(N, H, W, C, stride) || bnb == nullptr || bnb->sums == nullptr || bnb->yraw == nullptr) return -1;
  const hz::dw::Geo g = hz::dw::make_geo(N, H, W, C, stride);
  const int grid = rows_grid(g.N * g.H * g.W, C, 2, 148 * 4);
  if (!bnb->sums_is_zero) hz::zero_f32(bnb->sums, (size_t)2 * C, st);
  hz::DwBnBwd b;
  b.out = (const __nv_bfloat16*)bnb->out; b.yraw = (const __nv_bfloat16*)bnb->yraw;
  b.mean = bnb->mean; b.invstd = bnb->invstd; b.sums = bnb->sums; b.cap6 = bnb->cap6;
  return hz::launch(hz::dwconv_dgrad_kernel<true>, dim3(grid), dim3(256), sizeof(float) * 256 * 16, st,
                    (const __nv_bfloat16*)dy, (const __nv_bfloat16*)w, (__nv_bfloat16*)dx, g, b) == cudaSuccess ? 0 : -2;
}

// dw: fp32 [C][3][3]; accumulate = add to what is there; otherwise the buffer is cleared first unless `prezeroed`
int hz_dwconv_wgrad(const void* dy, const void* x, float* dwt, int N, int H, int W, int C, int stride, int accumulate,
                    int prezeroed, cudaStream_t st) {
  if (!hz_dwconv_ok(N, H, W, C, stride)) return -1;
  const hz::dw::Geo g = hz::dw::make_geo(N, H, W, C, stride);
  if (!accumulate && !prezeroed) hz::zero_f32(dwt, (size_t)C * 9, st);
  const int grid = rows_grid(g.N * g.Ho * g.Wo, C, 8, 148 * 2);   // >= 8 rows per lane before the atomics
  return hz::launch(hz::dwconv_wgrad_kernel, dim3(grid), dim3(256), sizeof(float) * 256 * 8, st,
                    (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, dwt, g) == cudaSuccess ? 0 : -2;
}

}  // extern "C"

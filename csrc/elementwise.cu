// Memory-bound sm_100a kernels of the ResNet-18 training step:
//   BatchNorm(train)+residual+ReLU forward / backward, channel statistics, 3x3/2 max-pool,
//   stem input normalisation + im2col, fused avgpool+FC+softmax-CE head (fwd+bwd), flat fused Adam,
//   gradient-divergence reduction, on-device step statistics.
// Activations are NHWC bf16 ([M = N*H*W, C] row-major), 16-byte vectorised (8 channels / thread).
// Reference sites: torchvision BatchNorm2d/ReLU/MaxPool2d inside resnet18 (data_parallel_train.py:198),
// CrossEntropyLoss + argmax accuracy (:114-115,:128-130), optim.Adam.step (:122), grad divergence (:132-145).
#include "common.cuh"
#include "launchers.h"

namespace hz {

// ------------------------------------------------------------------------------------------------
// per-channel reductions over rows: block = (C/8 channel-vectors) x (256/(C/8) row lanes)
// ------------------------------------------------------------------------------------------------
// kGen: any channel count that is a multiple of 8 (the row-lane split leaves 256 % (C/8) threads idle) and the
// ReLU6 mask (relu == 2) — MobileNetV2's 24/96/144/160/320/...-channel layers; the default instantiation is the
// ResNet one (C/8 a power of two, plain ReLU).
template <bool kBwd, bool kGen = false>
__global__ void __launch_bounds__(256) channel_reduce_kernel(
    const __nv_bfloat16* __restrict__ a,      // fwd: y_raw           bwd: dout
    const __nv_bfloat16* __restrict__ outp,   // bwd: bn output (ReLU mask), may be null
    const __nv_bfloat16* __restrict__ yraw,   // bwd: y_raw
    const float* __restrict__ mean, const float* __restrict__ invstd,
    float* __restrict__ sums,                 // [2C], pre-zeroed: fwd (Σy, Σy²)  bwd (Σg, Σg·x̂)
    int M, int C, int relu) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float red[];              // [rlanes][nvec][16]
  const int nvec = C >> 3;
  const int rlanes = 256 / nvec;
  const int vec = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  float mu[8], is[8];
  if (kBwd) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { mu[i] = mean[vec * 8 + i]; is[i] = invstd[vec * 8 + i]; }
  }
  const bool active = !kGen || rl < rlanes;
  if (active) {
  for (int r = blockIdx.x * rlanes + rl; r < M; r += gridDim.x * rlanes) {
    const size_t off = (size_t)r * C + vec * 8;
    float f[8];
    unpack8(ld8(a + off), f);
    if (!kBwd) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * f[i]; }
    } else {
      float y[8];
      unpack8(ld8(yraw + off), y);
      if (relu) {
        float o[8];
        unpack8(ld8(outp + off), o);
        if (kGen && relu == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = (o[i] > 0.f && o[i] < 6.f) ? f[i] : 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = o[i] > 0.f ? f[i] : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * (y[i] - mu[i]) * is[i]; }
    }
  }
  float* mine = red + ((size_t)rl * nvec + vec) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { mine[i] = s[i]; mine[8 + i] = q[i]; }
  }
  __syncthreads();
  // nvec*16 outputs, each summed over rlanes
  for (int o = threadIdx.x; o < nvec * 16; o += 256) {
    const int v = o >> 4, j = o & 15;
    float t = 0.f;
    for (int k = 0; k < rlanes; ++k) t += red[((size_t)k * nvec + v) * 16 + j];
    const int c = v * 8 + (j & 7);
    atomicAdd(&sums[(j >> 3) * C + c], t);
  }
}

// ------------------------------------------------------------------------------------------------
// BN apply (+ residual) (+ ReLU); every CTA derives scale/shift from the sums in smem,
// CTA 0 publishes mean / invstd / running statistics.
// ------------------------------------------------------------------------------------------------
template <bool kGen = false>     // kGen: relu == 2 clamps at 6 (ReLU6)
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(
    const __nv_bfloat16* __restrict__ y, const float* __restrict__ sums,
    const float* __restrict__ gamma, const float* __restrict__ beta,
    const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
    float* __restrict__ mean_out, float* __restrict__ invstd_out,
    float* __restrict__ rmean, float* __restrict__ rvar,
    int M, int C, float eps, float momentum, int relu, int training) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float sm[];   // scale[C], shift[C]
  float* scale = sm;
  float* shift = sm + C;
  const float inv_cnt = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mu, var;
    if (training) {
      mu = sums[c] * inv_cnt;
      var = fmaxf(sums[C + c] * inv_cnt - mu * mu, 0.f);
    } else {
      mu = rmean[c];
      var = rvar[c];
    }
    const float is = rsqrtf(var + eps);
    const float g = gamma[c];
    scale[c] = g * is;
    shift[c] = beta[c] - mu * g * is;
    if (blockIdx.x == 0) {
      mean_out[c] = mu;
      invstd_out[c] = is;
      if (training && rmean != nullptr) {
        const float unb = var * ((float)M / (float)max(M - 1, 1));
        rmean[c] = (1.f - momentum) * rmean[c] + momentum * mu;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * unb;
      }
    }
  }
  __syncthreads();
  const int nvec = C >> 3;
  const size_t total = (size_t)M * nvec;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(v % nvec) * 8;
    float f[8];
    unpack8(ld8(y + v * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = f[i] * scale[c0 + i] + shift[c0 + i];
    if (residual != nullptr) {
      float r[8];
      unpack8(ld8(residual + v * 8), r);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += r[i];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
      if (kGen && relu == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fminf(f[i], 6.f);
      }
    }
    st8(out + v * 8, pack8(f));
  }
}

// dy_raw = γ·invstd·(g − Σg/M − x̂·Σ(g·x̂)/M), g = dout·[out>0];  dres = g;  CTA0: dγ, dβ
template <bool kGen = false>     // kGen: relu == 2 is the ReLU6 mask (0 < out < 6)
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ outp,
    const __nv_bfloat16* __restrict__ yraw, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ sums, __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dres,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int acc_gamma, int acc_beta,
    int M, int C, int relu) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float sm[];   // k[C], a[C], b[C], mu[C], is[C]
  float *k = sm, *a = sm + C, *b = sm + 2 * C, *mu = sm + 3 * C, *is = sm + 4 * C;
  const float inv_cnt = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float i_s = invstd[c];
    k[c] = gamma[c] * i_s;
    a[c] = sums[c] * inv_cnt;
    b[c] = sums[C + c] * inv_cnt;
    mu[c] = mean[c];
    is[c] = i_s;
    if (blockIdx.x == 0) {
      if (dgamma != nullptr) dgamma[c] = (acc_gamma ? dgamma[c] : 0.f) + sums[C + c];
      if (dbeta != nullptr) dbeta[c] = (acc_beta ? dbeta[c] : 0.f) + sums[c];
    }
  }
  __syncthreads();
  const int nvec = C >> 3;
  const size_t total = (size_t)M * nvec;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(v % nvec) * 8;
    float g[8], y[8];
    unpack8(ld8(dout + v * 8), g);
    unpack8(ld8(yraw + v * 8), y);
    if (relu) {
      float o[8];
      unpack8(ld8(outp + v * 8), o);
      if (kGen && relu == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = (o[i] > 0.f && o[i] < 6.f) ? g[i] : 0.f;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = o[i] > 0.f ? g[i] : 0.f;
      }
    }
    if (dres != nullptr) st8(dres + v * 8, pack8(g));
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      const float xh = (y[i] - mu[c]) * is[c];
      d[i] = k[c] * (g[i] - a[c] - xh * b[c]);
    }
    st8(dy + v * 8, pack8(d));
  }
}

// BN-backward apply of a layer whose output had a residual added (a BasicBlock's bn2 with a downsample branch) that also
// takes the backward sums of the BatchNorm producing that residual (the downsample BN: no activation, and this kernel's
// `dres` is its only upstream gradient): res_sums[0:C] += sum dres, res_sums[C:2C] += sum dres * xhat_res, from the
// bf16 values being stored — the downsample BN's own reduction pass disappears (ops.BNBackLink, HZ_BN_BWD_IN_DGRAD).
// Row-lane thread layout of the reductions (C/8 a power of two); arithmetic of bn_act_bwd_apply_kernel.
__global__ void __launch_bounds__(256) bn_act_bwd_apply_res_kernel(
    const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ outp,
    const __nv_bfloat16* __restrict__ yraw, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ sums, __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dres,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int acc_gamma, int acc_beta,
    int M, int C, int relu, const __nv_bfloat16* __restrict__ res_yraw, const float* __restrict__ res_mean,
    const float* __restrict__ res_invstd, float* __restrict__ res_sums) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float sm[];   // k[C], a[C], b[C], mu[C], is[C], red[256 * 16]
  float *k = sm, *a = sm + C, *b = sm + 2 * C, *mu = sm + 3 * C, *is = sm + 4 * C, *red = sm + 5 * C;
  const float inv_cnt = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float i_s = invstd[c];
    k[c] = gamma[c] * i_s;
    a[c] = sums[c] * inv_cnt;
    b[c] = sums[C + c] * inv_cnt;
    mu[c] = mean[c];
    is[c] = i_s;
    if (blockIdx.x == 0) {
      if (dgamma != nullptr) dgamma[c] = (acc_gamma ? dgamma[c] : 0.f) + sums[C + c];
      if (dbeta != nullptr) dbeta[c] = (acc_beta ? dbeta[c] : 0.f) + sums[c];
    }
  }
  __syncthreads();
  const int nvec = C >> 3;
  const int rlanes = 256 / nvec;
  const int cv = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  const int c0 = cv * 8;
  float rmu[8], ris[8], s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { rmu[i] = res_mean[c0 + i]; ris[i] = res_invstd[c0 + i]; s[i] = q[i] = 0.f; }
  for (int p = blockIdx.x * rlanes + rl; p < M; p += gridDim.x * rlanes) {
    const size_t off = (size_t)p * C + c0;
    float g[8], y[8];
    unpack8(ld8(dout + off), g);
    unpack8(ld8(yraw + off), y);
    if (relu) {
      float o[8];
      unpack8(ld8(outp + off), o);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = o[i] > 0.f ? g[i] : 0.f;
    }
    const bf16x8 gv = pack8(g);
    st8(dres + off, gv);
    float gr[8], yr[8];
    unpack8(gv, gr);                                  // the rounded values the downsample BN's reduction would read
    unpack8(ld8(res_yraw + off), yr);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += gr[i]; q[i] += gr[i] * (yr[i] - rmu[i]) * ris[i]; }
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      const float xh = (y[i] - mu[c]) * is[c];
      d[i] = k[c] * (g[i] - a[c] - xh * b[c]);
    }
    st8(dy + off, pack8(d));
  }
  float* mine = red + ((size_t)rl * nvec + cv) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { mine[i] = s[i]; mine[8 + i] = q[i]; }
  __syncthreads();
  for (int o = threadIdx.x; o < nvec * 16; o += 256) {
    const int v = o >> 4, j = o & 15;
    float t = 0.f;
    for (int kk = 0; kk < rlanes; ++kk) t += red[((size_t)kk * nvec + v) * 16 + j];
    atomicAdd(&res_sums[(j >> 3) * C + v * 8 + (j & 7)], t);
  }
}

// Single-kernel BN backward: per-channel reduction -> device-wide barrier -> apply.
// grid <= #SMs so all CTAs are co-resident (the barrier spins); sums[2C] and the barrier counter come
// pre-zeroed from the statistics arena.
HZ_DEVINL unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) bn_act_bwd_fused_kernel(
    const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ outp,
    const __nv_bfloat16* __restrict__ yraw, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, float* __restrict__ sums,
    unsigned* __restrict__ counter, __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dres,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int acc_gamma, int acc_beta, int M, int C,
    int relu) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float sm[];      // phase 1: red[256*16]   phase 2: k,a,b,mu,is [5C]
  {
    float* red = sm;
    const int nvec = C >> 3;
    const int rlanes = 256 / nvec;
    const int vec = threadIdx.x % nvec, rl = threadIdx.x / nvec;
    float s[8], q[8], mu8[8], is8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = q[i] = 0.f; mu8[i] = mean[vec * 8 + i]; is8[i] = invstd[vec * 8 + i]; }
    for (int r = blockIdx.x * rlanes + rl; r < M; r += gridDim.x * rlanes) {
      const size_t off = (size_t)r * C + vec * 8;
      float f[8], y[8];
      unpack8(ld8(dout + off), f);
      unpack8(ld8(yraw + off), y);
      if (relu) {
        float o[8];
        unpack8(ld8(outp + off), o);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = o[i] > 0.f ? f[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * (y[i] - mu8[i]) * is8[i]; }
    }
    float* mine = red + ((size_t)rl * nvec + vec) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) { mine[i] = s[i]; mine[8 + i] = q[i]; }
    __syncthreads();
    for (int o = threadIdx.x; o < nvec * 16; o += 256) {
      const int v = o >> 4, j = o & 15;
      float t = 0.f;
      for (int k = 0; k < rlanes; ++k) t += red[((size_t)k * nvec + v) * 16 + j];
      atomicAdd(&sums[(j >> 3) * C + v * 8 + (j & 7)], t);
    }
  }
  // ---- device-wide barrier
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while (ld_acquire_gpu_u32(counter) < gridDim.x) {
      if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
  __syncthreads();
  // ---- apply
  float *k = sm, *a = sm + C, *b = sm + 2 * C, *mu = sm + 3 * C, *is = sm + 4 * C;
  const float inv_cnt = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float i_s = invstd[c];
    const float sg = __ldcg(&sums[c]), sq = __ldcg(&sums[C + c]);
    k[c] = gamma[c] * i_s;
    a[c] = sg * inv_cnt;
    b[c] = sq * inv_cnt;
    mu[c] = mean[c];
    is[c] = i_s;
    if (blockIdx.x == 0) {
      if (dgamma != nullptr) dgamma[c] = (acc_gamma ? dgamma[c] : 0.f) + sq;
      if (dbeta != nullptr) dbeta[c] = (acc_beta ? dbeta[c] : 0.f) + sg;
    }
  }
  __syncthreads();
  const int nvec = C >> 3;
  const size_t total = (size_t)M * nvec;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(v % nvec) * 8;
    float g[8], y[8];
    unpack8(ld8(dout + v * 8), g);
    unpack8(ld8(yraw + v * 8), y);
    if (relu) {
      float o[8];
      unpack8(ld8(outp + v * 8), o);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = o[i] > 0.f ? g[i] : 0.f;
    }
    if (dres != nullptr) st8(dres + v * 8, pack8(g));
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      const float xh = (y[i] - mu[c]) * is[c];
      d[i] = k[c] * (g[i] - a[c] - xh * b[c]);
    }
    st8(dy + v * 8, pack8(d));
  }
}

// ------------------------------------------------------------------------------------------------
// max-pool 3x3 / stride 2 / pad 1, NHWC; backward recomputes the (first) arg-max: no index tensor
// ------------------------------------------------------------------------------------------------
// forward also records the window position (0..8, row-major) of the first maximum: backward is a gather
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                          __nv_bfloat16* __restrict__ y,
                                                          uint8_t* __restrict__ idx, int N, int H, int W, int C,
                                                          int Ho, int Wo) {
  pdl_launch();
  pdl_wait();
  const int nvec = C >> 3;
  const size_t total = (size_t)N * Ho * Wo * nvec;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(v % nvec);
    size_t p = v / nvec;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float m[8];
    int am[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { m[i] = -INFINITY; am[i] = 0; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = ho * 2 - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = wo * 2 - 1 + s;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(ld8(x + (((size_t)n * H + h) * W + w) * C + cv * 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (f[i] > m[i]) { m[i] = f[i]; am[i] = r * 3 + s; }
      }
    }
    st8(y + v * 8, pack8(m));
    if (idx != nullptr) {
      uint2 pk;
      pk.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
      pk.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
      *reinterpret_cast<uint2*>(idx + v * 8) = pk;
    }
  }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                          const uint8_t* __restrict__ idx,
                                                          __nv_bfloat16* __restrict__ dx, int N, int H, int W,
                                                          int C, int Ho, int Wo) {
  pdl_launch();
  pdl_wait();
  const int nvec = C >> 3;
  const size_t total = (size_t)N * H * W * nvec;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(v % nvec);
    size_t p = v / nvec;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int n = (int)(p / H);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int ho = h / 2; ho <= (h + 1) / 2; ++ho) {
      if (ho >= Ho) continue;
      const int r = h - (ho * 2 - 1);
      for (int wo = w / 2; wo <= (w + 1) / 2; ++wo) {
        if (wo >= Wo) continue;
        const int me = r * 3 + (w - (wo * 2 - 1));
        const size_t oo = (((size_t)n * Ho + ho) * Wo + wo) * C + cv * 8;
        const uint2 pk = *reinterpret_cast<const uint2*>(idx + oo);
        float gv[8];
        unpack8(ld8(dy + oo), gv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int a = ((i < 4 ? pk.x : pk.y) >> (8 * (i & 3))) & 0xFF;
          acc[i] += (a == me) ? gv[i] : 0.f;
        }
      }
    }
    st8(dx + v * 8, pack8(acc));
  }
}

// max-pool backward that also takes the BatchNorm-backward sums of the layer feeding the pool (the stem's bn1: the pool
// is the only consumer of its output, so dx IS that BatchNorm's upstream gradient): sums[0:C] += sum g,
// sums[C:2C] += sum g*xhat with g = dx * [bn_out > 0] from the bf16 values being stored — the separate reduction pass
// over dx / out / y_raw (the largest one of the step: 64 x 16 x 16 x 64) disappears.  Thread layout of the reductions
// (C/8 channel vectors x 256/(C/8) row lanes, C/8 a power of two), pixel math of maxpool_bwd_kernel.
__global__ void __launch_bounds__(256) maxpool_bwd_bn_kernel(
    const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx, __nv_bfloat16* __restrict__ dx,
    const __nv_bfloat16* __restrict__ bn_out, const __nv_bfloat16* __restrict__ bn_yraw,
    const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ sums, int relu,
    int N, int H, int W, int C, int Ho, int Wo) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float red[];              // [rlanes][nvec][16]
  const int nvec = C >> 3;
  const int rlanes = 256 / nvec;
  const int cv = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  float s[8], q[8], mu[8], is[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = q[i] = 0.f; mu[i] = mean[cv * 8 + i]; is[i] = invstd[cv * 8 + i]; }
  const int P = N * H * W;
  for (int p = blockIdx.x * rlanes + rl; p < P; p += gridDim.x * rlanes) {
    const int w = p % W;
    const int t = p / W;
    const int h = t % H;
    const int n = t / H;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int ho = h / 2; ho <= (h + 1) / 2; ++ho) {
      if (ho >= Ho) continue;
      const int r = h - (ho * 2 - 1);
      for (int wo = w / 2; wo <= (w + 1) / 2; ++wo) {
        if (wo >= Wo) continue;
        const int me = r * 3 + (w - (wo * 2 - 1));
        const size_t oo = (((size_t)n * Ho + ho) * Wo + wo) * C + cv * 8;
        const uint2 pk = *reinterpret_cast<const uint2*>(idx + oo);
        float gv[8];
        unpack8(ld8(dy + oo), gv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int a = ((i < 4 ? pk.x : pk.y) >> (8 * (i & 3))) & 0xFF;
          acc[i] += (a == me) ? gv[i] : 0.f;
        }
      }
    }
    const size_t off = (size_t)p * C + cv * 8;
    const bf16x8 v = pack8(acc);
    st8(dx + off, v);
    float g[8], y[8];
    unpack8(v, g);                                   // the rounded values a separate pass would read back
    unpack8(ld8(bn_yraw + off), y);
    if (relu) {
      float o[8];
      unpack8(ld8(bn_out + off), o);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = o[i] > 0.f ? g[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += g[i]; q[i] += g[i] * (y[i] - mu[i]) * is[i]; }
  }
  float* mine = red + ((size_t)rl * nvec + cv) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { mine[i] = s[i]; mine[8 + i] = q[i]; }
  __syncthreads();
  for (int o = threadIdx.x; o < nvec * 16; o += 256) {
    const int v = o >> 4, j = o & 15;
    float tt = 0.f;
    for (int k = 0; k < rlanes; ++k) tt += red[((size_t)k * nvec + v) * 16 + j];
    atomicAdd(&sums[(j >> 3) * C + v * 8 + (j & 7)], tt);
  }
}

// ------------------------------------------------------------------------------------------------
// stem: uint8 NHWC -> normalised bf16 NHWC;  small-Cin im2col -> A[M, Kp] (K = (r,s,c), zero padded)
// ------------------------------------------------------------------------------------------------
__global__ void u8_normalize_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                    size_t n, float mean, float inv_std) {
  pdl_launch();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(((float)in[i] * (1.f / 255.f) - mean) * inv_std);
}

template <int CIN, int RR, int SS>
__global__ void __launch_bounds__(256) im2col_small_cin_kernel(
    const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ A, int N, int H, int W, int Cin_,
    int R_, int S_, int stride, int pad, int Ho, int Wo, int K, int Kp) {
  pdl_launch();
  pdl_wait();
  const int Cin = CIN > 0 ? CIN : Cin_;
  const int S = SS > 0 ? SS : S_;
  (void)R_;
  const int kvec = Kp >> 3;
  const size_t total = (size_t)N * Ho * Wo * kvec;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (size_t)gridDim.x * blockDim.x) {
    const int kv = (int)(v % kvec);
    size_t m = v / kvec;
    const int wo = (int)(m % Wo);
    size_t t = m / Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kv * 8 + i;
      float val = 0.f;
      if (k < K) {
        const int c = k % Cin;
        const int rs = k / Cin;
        const int s = rs % S, r = rs / S;
        const int h = ho * stride - pad + r, w = wo * stride - pad + s;
        if (h >= 0 && h < H && w >= 0 && w < W)
          val = __bfloat162float(x[(((size_t)n * H + h) * W + w) * Cin + c]);
      }
      f[i] = val;
    }
    st8(A + v * 8, pack8(f));
  }
}

// The 7x7/2 stem in one launch: CTAs [0, a_blocks) build the im2col matrix A[N*Ho*Wo][KP] of the 3-channel
// input (k = (r*7+s)*3+c, zero padded to KP), the remaining CTAs pad the [Cout,147] weight rows to KP —
// 32-bit index math with compile-time divisors (the generic kernel above spends its time in 64-bit divides).
template <int KP>
__global__ void __launch_bounds__(256) stem_pack_kernel(const __nv_bfloat16* __restrict__ x,
                                                        __nv_bfloat16* __restrict__ A,
                                                        const __nv_bfloat16* __restrict__ w_in,
                                                        __nv_bfloat16* __restrict__ w_out, int N, int H, int W,
                                                        int stride, int pad, int Ho, int Wo, int rows_w,
                                                        int a_blocks) {
  constexpr int CIN = 3, SS = 7, K = 147, KVEC = KP / 8, RUN = SS * CIN;
  pdl_launch();
  pdl_wait();
  if ((int)blockIdx.x >= a_blocks) {
    const int total = rows_w * KP;
    for (int i = ((int)blockIdx.x - a_blocks) * 256 + threadIdx.x; i < total; i += ((int)gridDim.x - a_blocks) * 256) {
      const int r = i / KP, k = i % KP;
      w_out[i] = k < K ? w_in[r * K + k] : __float2bfloat16_rn(0.f);
    }
    return;
  }
  const int total = N * Ho * Wo * KVEC;
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
  for (int v = (int)blockIdx.x * 256 + threadIdx.x; v < total; v += a_blocks * 256) {
    const int kv = v % KVEC;
    const int m = v / KVEC;
    const int wo = m % Wo;
    const int t = m / Wo;
    const int ho = t % Ho;
    const int n = t / Ho;
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    const int img = n * H;
    unsigned short e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kv * 8 + i;
      unsigned short val = 0;
      if (k < K) {
        const int r = k / RUN, j = k % RUN;           // j = s*3 + c: 21 contiguous input elements per patch row
        const int h = h0 + r, w = w0 + j / CIN;
        if ((unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W)
          val = __ldg(xs + ((img + h) * W + w0) * CIN + j);
      }
      e[i] = val;
    }
    uint4 pk;
    pk.x = e[0] | ((uint32_t)e[1] << 16);
    pk.y = e[2] | ((uint32_t)e[3] << 16);
    pk.z = e[4] | ((uint32_t)e[5] << 16);
    pk.w = e[6] | ((uint32_t)e[7] << 16);
    *reinterpret_cast<uint4*>(A + (size_t)v * 8) = pk;
  }
}

// rows of length K (bf16) -> rows of length Kp (zero padded): packs conv1's [64,147] weights for TMA
__global__ void pad_rows_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                int rows, int K, int Kp) {
  pdl_launch();
  pdl_wait();
  const int total = rows * Kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / Kp, k = i % Kp;
    out[i] = k < K ? in[(size_t)r * K + k] : __float2bfloat16_rn(0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// head: global-avg-pool -> FC -> softmax-CE (+accuracy) forward AND backward.
//   kernel A: one CTA per sample (pooled, logits, loss, correct, dlogits, dfeat)
//   kernel B: dW[k,c] = Σ_n dlogits[n,k]·pooled[n,c], db[k] = Σ_n dlogits[n,k]
// ------------------------------------------------------------------------------------------------
constexpr int kHeadMaxK = 64;

__global__ void __launch_bounds__(128) head_sample_kernel(
    const __nv_bfloat16* __restrict__ feat, const float* Wt, const float* __restrict__ bias,
    const int64_t* __restrict__ labels, float* __restrict__ pooled, float* __restrict__ dlogits,
    float* __restrict__ logits_out, __nv_bfloat16* __restrict__ dfeat,
    float* __restrict__ loss_out, float* __restrict__ correct_out, int N, int C, int HW, int K,
    int n_valid, float loss_scale, int w_in_smem) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float sm[];            // pooled[C], logit[kHeadMaxK], dl[kHeadMaxK], (W[n_valid][C])
  float* pl = sm;
  float* lg = sm + C;
  float* dl = lg + kHeadMaxK;
  const int n = blockIdx.x;
  const float inv_hw = 1.f / (float)HW;
  // stage the classifier weights once (coalesced float4): logits and dfeat then read shared memory only,
  // so the kernel has one global-load latency instead of three dependent ones
  if (w_in_smem) {
    float* wsm = dl + kHeadMaxK;
    const int nv4 = (n_valid * C) >> 2;
    for (int i = threadIdx.x; i < nv4; i += blockDim.x)
      reinterpret_cast<float4*>(wsm)[i] = reinterpret_cast<const float4*>(Wt)[i];
    Wt = wsm;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += __bfloat162float(feat[((size_t)n * HW + p) * C + c]);
    s *= inv_hw;
    pl[c] = s;
    pooled[(size_t)n * C + c] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int k = warp; k < K; k += nwarp) {
    float s = 0.f;
    if (k < n_valid) {                       // padded classes are masked: never touch their (unstaged) rows
      for (int c = lane; c < C; c += 32) s += pl[c] * Wt[(size_t)k * C + c];
      s = warp_sum(s);
    }
    if (lane == 0) lg[k] = (k < n_valid) ? s + (bias ? bias[k] : 0.f) : -INFINITY;
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    int amax = 0;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, lg[k]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < K; k += 32) se += __expf(lg[k] - mx);
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    const int lab = (int)labels[n];
    // first arg-max (torch.argmax tie-break is irrelevant for real data; pick the lowest index)
    int best = K;
    for (int k = lane; k < K; k += 32) if (lg[k] == mx) best = min(best, k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    amax = best;
    for (int k = lane; k < K; k += 32) {
      const float p = __expf(lg[k] - lse);
      const float d = (p - (k == lab ? 1.f : 0.f)) * (loss_scale / (float)N);
      dl[k] = d;
      dlogits[(size_t)n * K + k] = d;
      if (logits_out) logits_out[(size_t)n * K + k] = lg[k];
    }
    if (lane == 0) {
      atomicAdd(loss_out, (lse - lg[lab]) * (loss_scale / (float)N));
      if (amax == lab) atomicAdd(correct_out, 1.f);
    }
  }
  __syncthreads();
  if (dfeat != nullptr) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
      for (int k = 0; k < n_valid; ++k) s += dl[k] * Wt[(size_t)k * C + c];
      const __nv_bfloat16 g = __float2bfloat16_rn(s * inv_hw);
      for (int p = 0; p < HW; ++p) dfeat[((size_t)n * HW + p) * C + c] = g;
    }
  }
}

// CTA = 32 input channels x 4 sample groups; 16 classes in registers; dlogits tile staged in smem
__global__ void __launch_bounds__(128) head_wgrad_kernel(const float* __restrict__ pooled,
                                                         const float* __restrict__ dlogits,
                                                         float* __restrict__ dW, float* __restrict__ db,
                                                         int N, int C, int K, int accumulate) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float hsm[];                // dl[N][16] | part[4][32][16]
  float* dl = hsm;
  float* part = hsm + N * 16;
  const int k0 = blockIdx.y * 16;
  for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
    const int n = i >> 4, j = i & 15;
    dl[i] = (k0 + j < K) ? dlogits[(size_t)n * K + k0 + j] : 0.f;
  }
  __syncthreads();
  const int cl = threadIdx.x & 31, ng = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  if (c < C) {
#pragma unroll 4
    for (int n = ng; n < N; n += 4) {
      const float pv = pooled[(size_t)n * C + c];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += dl[n * 16 + j] * pv;
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) part[(ng * 32 + cl) * 16 + j] = acc[j];
  __syncthreads();
  // 32 channels x 16 classes = 512 outputs, 128 threads x 4
  for (int o = threadIdx.x; o < 512; o += 128) {
    const int cc = o >> 4, j = o & 15;
    const float t = part[(0 * 32 + cc) * 16 + j] + part[(1 * 32 + cc) * 16 + j] + part[(2 * 32 + cc) * 16 + j] +
                    part[(3 * 32 + cc) * 16 + j];
    const int ch = blockIdx.x * 32 + cc;
    if (ch < C && k0 + j < K) {
      const size_t oo = (size_t)(k0 + j) * C + ch;
      dW[oo] = (accumulate ? dW[oo] : 0.f) + t;
    }
  }
  if (db != nullptr && blockIdx.x == 0 && threadIdx.x < 16 && k0 + threadIdx.x < K) {
    float t = 0.f;
    for (int n = 0; n < N; ++n) t += dl[n * 16 + threadIdx.x];
    db[k0 + threadIdx.x] = (accumulate ? db[k0 + threadIdx.x] : 0.f) + t;
  }
}

// ------------------------------------------------------------------------------------------------
// flat fused Adam (fp32 master + moments, bf16 shadow refresh)  — torch.optim.Adam semantics
// ------------------------------------------------------------------------------------------------
__global__ void bump_step_kernel(float* step, float* diff) {
  pdl_launch();
  pdl_wait();
  step[0] += 1.f;
  if (diff != nullptr) diff[0] = 0.f;
}

// One pass over the flat buffers: Adam update, bf16 shadow refresh, optional gradient-divergence
// accumulation Σ(g − prev)² with prev ← g (reference metric, data_parallel_train.py:132-145) and
// optional g ← 0 (so the next step's wgrad kernels can accumulate without any memset).
template <bool kDiff, bool kZero>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   __nv_bfloat16* __restrict__ shadow,
                                                   const float* __restrict__ step, float* __restrict__ prev,
                                                   float* __restrict__ diff_out, size_t n, float lr,
                                                   float b1, float b2, float eps, float gscale,
                                                   const int* __restrict__ live) {
  pdl_launch();
  pdl_wait();
  __shared__ float wsum[8];
  const float t = step[0];
  const float bc1 = 1.f - __powf(b1, t), bc2 = 1.f - __powf(b2, t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  // n counts the elements actually visited: all of them, or 64 per live block (dead conv taps never get a
  // gradient, so their m = v = 0 and their weights never move: skipping them is exact, not an approximation)
  const size_t nv = n >> 2;
  float dacc = 0.f;
  for (size_t iv = (size_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nv;
       iv += (size_t)gridDim.x * blockDim.x) {
    const size_t i = live ? (size_t)live[iv >> 4] * 16 + (iv & 15) : iv;
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; float* G = &gg.x; float* Mo = &mm.x; float* V = &vv.x;
    if (kDiff) {
      const float4 pr = reinterpret_cast<float4*>(prev)[i];
      const float dx = gg.x - pr.x, dy = gg.y - pr.y, dz = gg.z - pr.z, dw = gg.w - pr.w;
      dacc += dx * dx + dy * dy + dz * dz + dw * dw;
      reinterpret_cast<float4*>(prev)[i] = gg;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = G[j] * gscale;
      Mo[j] = b1 * Mo[j] + (1.f - b1) * gr;
      V[j] = b2 * V[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(V[j]) * inv_sqrt_bc2 + eps;
      P[j] -= step_size * Mo[j] / denom;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (kZero) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(shadow)[i] = pk;
    }
  }
  if (kDiff) {
    dacc = warp_sum(dacc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = dacc;
    __syncthreads();
    if (threadIdx.x < 8) {
      float tt = wsum[threadIdx.x];
      tt += __shfl_xor_sync(0xffu, tt, 4);
      tt += __shfl_xor_sync(0xffu, tt, 2);
      tt += __shfl_xor_sync(0xffu, tt, 1);
      if (threadIdx.x == 0) atomicAdd(diff_out, tt);
    }
  }
}

// Σ (g − prev)², prev ← g      (out pre-zeroed)
__global__ void __launch_bounds__(256) grad_diff_kernel(const float* __restrict__ g,
                                                        float* __restrict__ prev,
                                                        float* __restrict__ out, size_t n) {
  pdl_launch();
  pdl_wait();
  __shared__ float wsum[8];
  float acc = 0.f;
  const size_t nv = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(g)[i];
    const float4 b = reinterpret_cast<float4*>(prev)[i];
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
    acc += dx * dx + dy * dy + dz * dz + dw * dw;
    reinterpret_cast<float4*>(prev)[i] = a;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = wsum[threadIdx.x];
    t += __shfl_xor_sync(0xffu, t, 4);
    t += __shfl_xor_sync(0xffu, t, 2);
    t += __shfl_xor_sync(0xffu, t, 1);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

// stats[0..5] += (loss, correct, batch, sqrt(diff)*has_prev, has_prev, 1); has_prev = 1
__global__ void stats_update_kernel(float* stats, float* has_prev, const float* loss,
                                    const float* correct, float batch, float* diff_sq) {
  pdl_launch();
  pdl_wait();
  stats[0] += loss[0];
  stats[1] += correct[0];
  stats[2] += batch;
  stats[5] += 1.f;
  if (diff_sq != nullptr) {
    const float hp = has_prev[0];
    stats[3] += sqrtf(diff_sq[0]) * hp;
    stats[4] += hp;
    has_prev[0] = 1.f;
    diff_sq[0] = 0.f;      // consumed: the bucket-wise optimizer passes of the next step accumulate into it from zero
  }
}

}  // namespace hz

// ================================================================================================
// launchers (plain C++ interface, no torch headers)
// ================================================================================================
namespace {
inline int grid_for(size_t work_items, int block, int cap = 148 * 8) {
  size_t g = (work_items + block - 1) / block;
  if (g < 1) g = 1;
  if (g > (size_t)cap) g = cap;
  return (int)g;
}
// head_wgrad_kernel stages dlogits[N][16] in shared memory: beyond N = 640 that exceeds the 48 KB a kernel gets without
// opting in (a --batch_size 1024 step used to fail at this launch)
inline bool head_wgrad_smem_ok(size_t bytes) {
  if (bytes <= 48 * 1024) return true;
  static size_t granted_on[16] = {0};                 // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& granted = granted_on[dev & 15];
  if (bytes <= granted) return true;
  if (bytes > 200 * 1024) return false;
  if (cudaFuncSetAttribute(hz::head_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  granted = bytes;
  return true;
}
inline int reduce_grid(int M, int C) {
  const int rlanes = 256 / (C / 8);
  int g = (M + rlanes - 1) / rlanes;
  g = (g + 3) / 4;                       // >= 4 rows per lane
  if (g < 1) g = 1;
  if (g > 148 * 2) g = 148 * 2;
  return g;
}
}  // namespace

extern "C" {

// 1: the default (ResNet) instantiations serve this channel count; 2: the generic ones do (any multiple of 8)
int hz_channel_ok(int C) {
  if (C < 8 || C > 2048 || (C & 7)) return 0;
  const int nvec = C >> 3;
  return (nvec & (nvec - 1)) == 0 ? 1 : 2;
}
namespace {
inline bool bn_generic(int C, int relu) { return relu == 2 || hz_channel_ok(C) == 2; }
}  // namespace

void hz_channel_sums(const void* y, float* sums, int M, int C, cudaStream_t st) {
  hz::zero_f32(sums, (size_t)2 * C, st);
  const size_t smem = sizeof(float) * 256 * 16;
  if (bn_generic(C, 0))
    hz::launch(hz::channel_reduce_kernel<false, true>, dim3(reduce_grid(M, C)), dim3(256), smem, st,
        (const __nv_bfloat16*)y, nullptr, nullptr, nullptr, nullptr, sums, M, C, 0);
  else
  hz::launch(hz::channel_reduce_kernel<false>, dim3(reduce_grid(M, C)), dim3(256), smem, st, 
      (const __nv_bfloat16*)y, nullptr, nullptr, nullptr, nullptr, sums, M, C, 0);
}

void hz_bn_act_fwd(const void* y, const float* sums, const float* gamma, const float* beta,
                   const void* residual, void* out, float* mean, float* invstd, float* rmean,
                   float* rvar, int M, int C, float eps, float momentum, int relu, int training,
                   cudaStream_t st) {
  const size_t smem = sizeof(float) * 2 * C;
  if (relu == 2)
    hz::launch(hz::bn_act_fwd_kernel<true>, dim3(grid_for((size_t)M * (C / 8), 256, 148 * 4)), dim3(256), smem, st,
        (const __nv_bfloat16*)y, sums, gamma, beta, (const __nv_bfloat16*)residual, (__nv_bfloat16*)out,
        mean, invstd, rmean, rvar, M, C, eps, momentum, relu, training);
  else
  hz::launch(hz::bn_act_fwd_kernel<false>, dim3(grid_for((size_t)M * (C / 8), 256, 148 * 4)), dim3(256), smem, st, 
      (const __nv_bfloat16*)y, sums, gamma, beta, (const __nv_bfloat16*)residual, (__nv_bfloat16*)out,
      mean, invstd, rmean, rvar, M, C, eps, momentum, relu, training);
}

void hz_bn_act_bwd(const void* dout, const void* outp, const void* yraw, const float* mean,
                   const float* invstd, const float* gamma, float* sums_scratch, void* dy, void* dres,
                   float* dgamma, float* dbeta, int acc_gamma, int acc_beta, int M, int C, int relu,
                   int scratch_is_zero, cudaStream_t st) {
  const bool gen = bn_generic(C, relu);
  if (scratch_is_zero == 2 && !gen) {
    // arena slice [2C sums | 32-float pad holding the barrier counter], all zero: one fused kernel
    int grid = grid_for((size_t)M * (C / 8), 256, 148);
    size_t smem = sizeof(float) * 256 * 16;
    if (smem < sizeof(float) * 5 * C) smem = sizeof(float) * 5 * C;
    hz::launch(hz::bn_act_bwd_fused_kernel, dim3(grid), dim3(256), smem, st, 
        (const __nv_bfloat16*)dout, (const __nv_bfloat16*)outp, (const __nv_bfloat16*)yraw, mean, invstd, gamma,
        sums_scratch, (unsigned*)(sums_scratch + 2 * C), (__nv_bfloat16*)dy, (__nv_bfloat16*)dres, dgamma, dbeta,
        acc_gamma, acc_beta, M, C, relu);
    return;
  }
  const size_t smem_r = sizeof(float) * 256 * 16;
  const size_t smem = sizeof(float) * 5 * C;
  if (scratch_is_zero == 3) {
    // the sums are already final: the dgrad kernel that produced `dout` took them in its epilogue
    // (hz_conv_dgrad_bnbwd) — only the apply pass is left
    if (gen)
      hz::launch(hz::bn_act_bwd_apply_kernel<true>, dim3(grid_for((size_t)M * (C / 8), 256, 148 * 4)), dim3(256), smem, st,
          (const __nv_bfloat16*)dout, (const __nv_bfloat16*)outp, (const __nv_bfloat16*)yraw, mean, invstd,
          gamma, sums_scratch, (__nv_bfloat16*)dy, (__nv_bfloat16*)dres, dgamma, dbeta, acc_gamma, acc_beta, M, C, relu);
    else
      hz::launch(hz::bn_act_bwd_apply_kernel<false>, dim3(grid_for((size_t)M * (C / 8), 256, 148 * 4)), dim3(256), smem, st,
          (const __nv_bfloat16*)dout, (const __nv_bfloat16*)outp, (const __nv_bfloat16*)yraw, mean, invstd,
          gamma, sums_scratch, (__nv_bfloat16*)dy, (__nv_bfloat16*)dres, dgamma, dbeta, acc_gamma, acc_beta, M, C, relu);
    return;
  }
  if (!scratch_is_zero) hz::zero_f32(sums_scratch, (size_t)2 * C, st);
  if (gen) {
    hz::launch(hz::channel_reduce_kernel<true, true>, dim3(reduce_grid(M, C)), dim3(256), smem_r, st,
        (const __nv_bfloat16*)dout, (const __nv_bfloat16*)outp, (const __nv_bfloat16*)yraw, mean, invstd,
        sums_scratch, M, C, relu);
    hz::launch(hz::bn_act_bwd_apply_kernel<true>, dim3(grid_for((size_t)M * (C / 8), 256, 148 * 4)), dim3(256), smem, st,
        (const __nv_bfloat16*)dout, (const __nv_bfloat16*)outp, (const __nv_bfloat16*)yraw, mean, invstd,
        gamma, sums_scratch, (__nv_bfloat16*)dy, (__nv_bfloat16*)dres, dgamma, dbeta, acc_gamma, acc_beta,
        M, C, relu);
    return;
  }
  hz::launch(hz::channel_reduce_kernel<true>, dim3(reduce_grid(M, C)), dim3(256), smem_r, st, 
      (const __nv_bfloat16*)dout, (const __nv_bfloat16*)outp, (const __nv_bfloat16*)yraw, mean, invstd,
      sums_scratch, M, C, relu);
  hz::launch(hz::bn_act_bwd_apply_kernel<false>, dim3(grid_for((size_t)M * (C / 8), 256, 148 * 4)), dim3(256), smem, st, 
      (const __nv_bfloat16*)dout, (const __nv_bfloat16*)outp, (const __nv_bfloat16*)yraw, mean, invstd,
      gamma, sums_scratch, (__nv_bfloat16*)dy, (__nv_bfloat16*)dres, dgamma, dbeta, acc_gamma, acc_beta,
      M, C, relu);
}

// hz_bn_act_bwd for a layer with a residual whose producer is a BatchNorm without activation (the downsample branch):
// also leaves that BatchNorm's backward sums in res_sums [2C] (see bn_act_bwd_apply_res_kernel).  scratch_is_zero as in
// hz_bn_act_bwd (3: the own sums are already final).  Returns -1 when the shape / activation is not covered.
int hz_bn_act_bwd_res(const void* dout, const void* outp, const void* yraw, const float* mean, const float* invstd,
                      const float* gamma, float* sums_scratch, void* dy, void* dres, float* dgamma, float* dbeta,
                      int acc_gamma, int acc_beta, int M, int C, int relu, int scratch_is_zero, const void* res_yraw,
                      const float* res_mean, const float* res_invstd, float* res_sums, int res_sums_is_zero,
                      cudaStream_t st) {
  if (hz_channel_ok(C) != 1 || relu == 2 || dres == nullptr || res_yraw == nullptr || res_sums == nullptr) return -1;
  if (scratch_is_zero != 3) {
    if (!scratch_is_zero) hz::zero_f32(sums_scratch, (size_t)2 * C, st);
    hz::launch(hz::channel_reduce_kernel<true>, dim3(reduce_grid(M, C)), dim3(256), sizeof(float) * 256 * 16, st,
        (const __nv_bfloat16*)dout, (const __nv_bfloat16*)outp, (const __nv_bfloat16*)yraw, mean, invstd,
        sums_scratch, M, C, relu);
  }
  if (!res_sums_is_zero) hz::zero_f32(res_sums, (size_t)2 * C, st);
  const int rlanes = 256 / (C / 8);
  int grid = (M + rlanes - 1) / rlanes;
  grid = (grid + 1) / 2;
  if (grid < 1) grid = 1;
  if (grid > 148 * 4) grid = 148 * 4;
  const size_t smem = sizeof(float) * ((size_t)5 * C + 256 * 16);
  return hz::launch(hz::bn_act_bwd_apply_res_kernel, dim3(grid), dim3(256), smem, st, (const __nv_bfloat16*)dout,
                    (const __nv_bfloat16*)outp, (const __nv_bfloat16*)yraw, mean, invstd, gamma, sums_scratch,
                    (__nv_bfloat16*)dy, (__nv_bfloat16*)dres, dgamma, dbeta, acc_gamma, acc_beta, M, C, relu,
                    (const __nv_bfloat16*)res_yraw, res_mean, res_invstd, res_sums) == cudaSuccess ? 0 : -1;
}

void hz_maxpool_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, cudaStream_t st) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  hz::launch(hz::maxpool_fwd_kernel, dim3(grid_for((size_t)N * Ho * Wo * (C / 8), 256)), dim3(256), 0, st, 
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (uint8_t*)idx, N, H, W, C, Ho, Wo);
}

void hz_maxpool_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, cudaStream_t st) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  hz::launch(hz::maxpool_bwd_kernel, dim3(grid_for((size_t)N * H * W * (C / 8), 256)), dim3(256), 0, st, 
      (const __nv_bfloat16*)dy, (const uint8_t*)idx, (__nv_bfloat16*)dx, N, H, W, C, Ho, Wo);
}

// max-pool backward + BatchNorm-backward sums of the layer that feeds the pool (see maxpool_bwd_bn_kernel); C/8 must be a
// power of two; sums [2C] fp32 (sums_is_zero: already cleared).  Returns 0, or -1 when the shape is not covered.
int hz_maxpool_bwd_bn(const void* dy, const void* idx, void* dx, const void* bn_out, const void* bn_yraw, const float* mean,
                      const float* invstd, float* sums, int sums_is_zero, int relu, int N, int H, int W, int C,
                      cudaStream_t st) {
  if (hz_channel_ok(C) != 1 || (long long)N * H * W >= (1ll << 31) || (relu && bn_out == nullptr)) return -1;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (!sums_is_zero) hz::zero_f32(sums, (size_t)2 * C, st);
  const int rlanes = 256 / (C / 8);
  int grid = (N * H * W + rlanes - 1) / rlanes;
  grid = (grid + 1) / 2;                                  // >= 2 pixels per lane
  if (grid < 1) grid = 1;
  if (grid > 148 * 4) grid = 148 * 4;
  return hz::launch(hz::maxpool_bwd_bn_kernel, dim3(grid), dim3(256), sizeof(float) * 256 * 16, st,
                    (const __nv_bfloat16*)dy, (const uint8_t*)idx, (__nv_bfloat16*)dx, (const __nv_bfloat16*)bn_out,
                    (const __nv_bfloat16*)bn_yraw, mean, invstd, sums, relu, N, H, W, C, Ho, Wo) == cudaSuccess ? 0 : -1;
}

void hz_u8_normalize(const void* in, void* out, size_t n, float mean, float std, cudaStream_t st) {
  hz::launch(hz::u8_normalize_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, (const uint8_t*)in, (__nv_bfloat16*)out, n,
                                                             mean, 1.f / std);
}

void hz_im2col_small(const void* x, void* A, int N, int H, int W, int Cin, int R, int S, int stride,
                     int pad, int Ho, int Wo, int Kp, cudaStream_t st) {
  const int grid = grid_for((size_t)N * Ho * Wo * (Kp / 8), 256, 148 * 16);
  if (Cin == 3 && R == 7 && S == 7)
    hz::launch(hz::im2col_small_cin_kernel<3, 7, 7>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)A, N, H, W,
                                                                Cin, R, S, stride, pad, Ho, Wo, R * S * Cin, Kp);
  else
    hz::launch(hz::im2col_small_cin_kernel<0, 0, 0>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)A, N, H, W,
                                                                Cin, R, S, stride, pad, Ho, Wo, R * S * Cin, Kp);
}

// stem (Cin=3, 7x7): im2col of x and zero-padding of the weight rows in one launch; returns 0 if handled
int hz_stem_pack(const void* x, void* A, const void* w, void* wp, int N, int H, int W, int Cin, int R, int stride,
                 int pad, int Ho, int Wo, int Kp, int rows_w, cudaStream_t st) {
  if (Cin != 3 || R != 7 || Kp != 192) return -1;
  const long long vecs = (long long)N * Ho * Wo * (Kp / 8);
  if (vecs > (1ll << 30)) return -1;
  const int a_blocks = (int)((vecs + 255) / 256) < 148 * 16 ? (int)((vecs + 255) / 256) : 148 * 16;
  const int w_blocks = (rows_w * Kp + 255) / 256;
  hz::launch(hz::stem_pack_kernel<192>, dim3(a_blocks + w_blocks), dim3(256), 0, st, (const __nv_bfloat16*)x,
             (__nv_bfloat16*)A, (const __nv_bfloat16*)w, (__nv_bfloat16*)wp, N, H, W, stride, pad, Ho, Wo, rows_w,
             a_blocks);
  return 0;
}

void hz_pad_rows(const void* in, void* out, int rows, int K, int Kp, cudaStream_t st) {
  hz::launch(hz::pad_rows_kernel, dim3(grid_for((size_t)rows * Kp, 256)), dim3(256), 0, st, 
      (const __nv_bfloat16*)in, (__nv_bfloat16*)out, rows, K, Kp);
}

void hz_head_fwd_bwd(const void* feat, const float* W, const float* bias, const int64_t* labels,
                     float* pooled, float* dlogits, float* logits, void* dfeat, float* loss,
                     float* correct, float* dW, float* db, int N, int C, int HW, int K, int n_valid,
                     float loss_scale, int accumulate, int out_is_zero, cudaStream_t st) {
  if (!out_is_zero) {
    hz::zero_f32(loss, 1, st);
    hz::zero_f32(correct, 1, st);
  }
  size_t smem = sizeof(float) * (C + 2 * hz::kHeadMaxK);
  const int w_in_smem = ((size_t)n_valid * C * sizeof(float) <= 40 * 1024 && (C & 3) == 0) ? 1 : 0;
  if (w_in_smem) smem += sizeof(float) * (size_t)n_valid * C;
  hz::launch(hz::head_sample_kernel, dim3(N), dim3(128), smem, st, (const __nv_bfloat16*)feat, W, bias, labels, pooled,
                                               dlogits, logits, (__nv_bfloat16*)dfeat, loss, correct, N,
                                               C, HW, K, n_valid, loss_scale, w_in_smem);
  hz_head_wgrad(pooled, dlogits, dW, db, N, C, K, accumulate, st);
}

// dW[k,c] (+)= sum_n dlogits[n,k] * pooled[n,c], db[k] (+)= sum_n dlogits[n,k]   (tensor-parallel head: local shard)
void hz_head_wgrad(const float* pooled, const float* dlogits, float* dW, float* db, int N, int C, int K,
                   int accumulate, cudaStream_t st) {
  dim3 grid((C + 31) / 32, (K + 15) / 16);
  constexpr int kChunk = 2048;                       // samples per launch: dlogits[chunk][16] must fit in shared memory
  for (int n0 = 0; n0 < N; n0 += kChunk) {
    const int nn = N - n0 < kChunk ? N - n0 : kChunk;
    const size_t smem = sizeof(float) * ((size_t)nn * 16 + 4 * 32 * 16);
    head_wgrad_smem_ok(smem);
    hz::launch(hz::head_wgrad_kernel, dim3(grid), dim3(128), smem, st, pooled + (size_t)n0 * C, dlogits + (size_t)n0 * K,
               dW, db, nn, C, K, (accumulate || n0 > 0) ? 1 : 0);
  }
}

void hz_adam(float* p, float* g, float* m, float* v, void* shadow, float* step, float* prev, float* diff_out,
             int zero_grad, size_t n, float lr, float b1, float b2, float eps, float gscale, const int* live,
             size_t n_live_blocks, int bump, int max_ctas, cudaStream_t st) {
  if (live != nullptr) n = n_live_blocks * 64;
  const bool diff = prev != nullptr && diff_out != nullptr;
  // bump = 0: a later bucket of the same optimizer step (step counter / divergence accumulator already set up)
  if (bump) hz::launch(hz::bump_step_kernel, dim3(1), dim3(1), 0, st, step, diff ? diff_out : nullptr);
  if (n == 0) return;
  // max_ctas > 0: a bucket pass that runs beside backward kernels on another stream — leave thread slots free
  const int grid = grid_for(n / 4, 256, max_ctas > 0 ? max_ctas : 148 * 8);
#define HZ_ADAM(D, Z)                                                                                   \
  hz::launch(hz::adam_kernel<D, Z>, dim3(grid), dim3(256), 0, st, p, g, m, v, (__nv_bfloat16*)shadow, step, prev, diff_out, n, lr, \
                                              b1, b2, eps, gscale, live)
  if (diff && zero_grad) HZ_ADAM(true, true);
  else if (diff) HZ_ADAM(true, false);
  else if (zero_grad) HZ_ADAM(false, true);
  else HZ_ADAM(false, false);
#undef HZ_ADAM
}

void hz_grad_diff(const float* g, float* prev, float* out, size_t n, cudaStream_t st) {
  hz::zero_f32(out, 1, st);
  hz::launch(hz::grad_diff_kernel, dim3(grid_for(n / 4, 256, 148 * 4)), dim3(256), 0, st, g, prev, out, n);
}

void hz_stats_update(float* stats, float* has_prev, const float* loss, const float* correct,
                     float batch, float* diff_sq, cudaStream_t st) {
  hz::launch(hz::stats_update_kernel, dim3(1), dim3(1), 0, st, stats, has_prev, loss, correct, batch, diff_sq);
}

}  // extern "C"

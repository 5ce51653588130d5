// Peer-memory gradient all-reduce for NVLink 5 / NVSwitch (§5.8).
//
// Replaces the reference's DDP bucket all-reduce on gloo (data_parallel_train.py:118,202; K2) and the
// 62 blocking per-parameter all-reduces of the tensor-parallel script (tensor_parallel_train.py:215-218; K9),
// and its per-step host barrier (K4/K10) with device-side flag barriers.
//
// One kernel per bucket does:   pack   : grad(fp32) * (1/W) -> wire dtype (bf16 | fp32) into the local,
//                                         peer-visible staging buffer                       [fused cast+scale]
//                               barrier: per-block flag exchange with the same block on every peer
//                               reduce : one-shot  - read all W staged copies, sum in fp32, write grad
//                                        two-shot  - reduce own 1/W slice, push it to every peer's `out`
//                                                    buffer (NVLink stores), barrier, unpack out -> grad
//                                        nvls      - like two-shot but the reduction is done by the switch:
//                                                    multimem.ld_reduce on the multicast address, result
//                                                    broadcast with multimem.st
// Summation order is rank 0..W-1 on every rank, so replicas stay bit-identical.
// Staging/out buffers alternate between calls (parity), which together with the in-call barrier makes
// back-to-back calls race-free without a trailing barrier.  All counters live in device memory so the
// kernels are CUDA-graph replayable.  Spins are bounded (error flag + trap instead of a hang or silent garbage).
// With the optimizer epilogue (AdamFuse) the reduced values go straight into the Adam update of the bucket's
// parameters: no gradient round trip through HBM, no separate optimizer kernel waiting for "all buckets".
#include "common.cuh"
#include "launchers.h"
#include "tc05.cuh"

#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <cstring>
#include <vector>

namespace hz {

constexpr int kMaxRanks = 8;
constexpr int kCommThreads = 512;
constexpr size_t kFlagBytes = 1 << 16;          // flags + counters region at the start of each rank's block
constexpr long long kDefaultSpinTimeoutNs = 30000000000LL;   // 30 s (HZ_COMM_TIMEOUT_S overrides)

struct CommDev {
  char* base[kMaxRanks];      // every rank's region (flags first)
  char* mc_base;              // multicast mapping of the NVLS data region (or null)
  char* mc_local;             // this rank's local mapping of the same region
  size_t buf_bytes;           // size of ONE staging/out buffer
  long long timeout_ns;       // bound on every peer spin: error flag + trap instead of a hang / silent garbage
  int rank, world;
};

// region layout: [flags: maxBlocks*world u32][counters: maxBlocks u32 (+calls)][err u32] ... [data @kFlagBytes]
// data: stage[0], stage[1], out[0], out[1]  (each buf_bytes)
HZ_DEVINL uint32_t* flags_of(char* base) { return reinterpret_cast<uint32_t*>(base); }
HZ_DEVINL uint32_t* counters_of(char* base) { return reinterpret_cast<uint32_t*>(base + 32768); }
HZ_DEVINL uint32_t* calls_of(char* base) { return reinterpret_cast<uint32_t*>(base + 49152); }
HZ_DEVINL uint32_t* err_of(char* base) { return reinterpret_cast<uint32_t*>(base + 61440); }

HZ_DEVINL void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
HZ_DEVINL uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
HZ_DEVINL long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Barrier between block `blockIdx.x` of every rank.  All prior global writes of this block are made
// visible system-wide before the flag is published; peers' writes are visible after it returns.
HZ_DEVINL void peer_block_barrier(const CommDev& c, uint32_t* s_epoch) {
  __syncthreads();
  char* my = c.base[c.rank];
  uint32_t* cnt = counters_of(my) + blockIdx.x;
  if (threadIdx.x == 0) *s_epoch = *cnt + 1;
  __syncthreads();
  const uint32_t e = *s_epoch;
  if (threadIdx.x < c.world) {
    const int r = threadIdx.x;
    __threadfence_system();
    st_release_sys(flags_of(c.base[r]) + blockIdx.x * c.world + c.rank, e);
    const uint32_t* mine = flags_of(my) + blockIdx.x * c.world + r;
    const long long t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      if (globaltimer_ns() - t0 > c.timeout_ns) {
        // a peer never arrived: reducing whatever sits in its staging buffer would silently corrupt the
        // gradients, so record the error (host-visible through hz_comm_error) and kill the context
        atomicExch(err_of(my), 1u);
        __threadfence_system();
        __trap();
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *cnt = e;
}

template <bool kBf16>
struct Wire;
template <>
struct Wire<true> {          // 8 elements / 16 B
  static constexpr int kVec = 8;
  HZ_DEVINL static uint4 pack(const float* g, float s) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = g[i] * s;
    bf16x8 p = pack8(f);
    return *reinterpret_cast<uint4*>(&p);
  }
  HZ_DEVINL static void accum(float* a, const uint4& w) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(&w), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] += f[i];
  }
  HZ_DEVINL static uint4 from_acc(const float* a) {
    bf16x8 p = pack8(a);
    return *reinterpret_cast<uint4*>(&p);
  }
  HZ_DEVINL static uint4 mc_ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
template <>
struct Wire<false> {         // 4 elements / 16 B
  static constexpr int kVec = 4;
  HZ_DEVINL static uint4 pack(const float* g, float s) {
    float4 f = make_float4(g[0] * s, g[1] * s, g[2] * s, g[3] * s);
    return *reinterpret_cast<uint4*>(&f);
  }
  HZ_DEVINL static void accum(float* a, const uint4& w) {
    const float4 f = *reinterpret_cast<const float4*>(&w);
    a[0] += f.x; a[1] += f.y; a[2] += f.z; a[3] += f.w;
  }
  HZ_DEVINL static uint4 from_acc(const float* a) {
    float4 f = make_float4(a[0], a[1], a[2], a[3]);
    return *reinterpret_cast<uint4*>(&f);
  }
  HZ_DEVINL static uint4 mc_ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};

HZ_DEVINL void mc_st(void* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// wire vector index -> element offset in the gradient buffer.  With `live` (indices of the 64-element blocks
// that can ever be non-zero) the wire buffer is the *compacted* gradient: provably-dead parameters (conv taps
// that only ever see padding) are neither packed, sent, reduced nor unpacked.
template <int V>
HZ_DEVINL size_t goff(const int* __restrict__ live, size_t vec) {
  constexpr int kPer = 64 / V;
  return live ? (size_t)live[vec / kPer] * 64 + (vec % kPer) * V : vec * V;
}
template <int V>
HZ_DEVINL void load_grad(const float* g, size_t off, float* f) {
#pragma unroll
  for (int i = 0; i < V / 4; ++i) {
    const float4 t = reinterpret_cast<const float4*>(g + off)[i];
    f[4 * i] = t.x; f[4 * i + 1] = t.y; f[4 * i + 2] = t.z; f[4 * i + 3] = t.w;
  }
}
template <int V>
HZ_DEVINL void store_grad(float* g, size_t off, const float* f) {
#pragma unroll
  for (int i = 0; i < V / 4; ++i)
    reinterpret_cast<float4*>(g + off)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}

enum Algo { kOneShot = 0, kTwoShot = 1, kNvls = 2, kLL = 3, kBulk = 4 };
constexpr size_t kLLMaxElems = 512 * 1024;            // largest bucket the latency protocol takes (bf16 wire)
constexpr size_t kLLSlotBytes = kLLMaxElems * 4;      // one rank's slot: 4 wire bytes per element ({2 bf16, flag} words)

// "LL" words: 16 bytes = {data, flag, data, flag}; 8-byte halves are single-copy atomic, so a receiver that sees the
// flag sees the data next to it: no fence, no separate flag, no barrier (NCCL's small-message protocol).
HZ_DEVINL void ll_store(void* dst, const uint4& v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
HZ_DEVINL uint4 ll_load(const void* src) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
  return v;
}

// sub-range b of slice r of nv vectors split over W ranks and B blocks
HZ_DEVINL void sub_range(size_t nv, int W, int B, int r, int b, size_t& lo, size_t& hi) {
  const size_t q = (nv + W - 1) / W;
  const size_t slo = min((size_t)r * q, nv), shi = min(slo + q, nv);
  const size_t qq = (shi - slo + B - 1) / B;
  lo = min(slo + (size_t)b * qq, shi);
  hi = min(lo + qq, shi);
}

constexpr int kUnroll = 8;     // independent 16-byte requests in flight per thread and phase

// pack [lo,hi): grad(fp32)*scale -> wire vectors in the local staging buffer
template <bool kBf16>
HZ_DEVINL void pack_range(const float* __restrict__ grad, uint4* __restrict__ stage, size_t lo, size_t hi,
                          float scale, const int* __restrict__ live) {
  using Wt = Wire<kBf16>;
  constexpr int V = Wt::kVec;
  for (size_t v0 = lo + threadIdx.x; v0 < hi; v0 += (size_t)blockDim.x * kUnroll) {
    float f[kUnroll][V];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t v = v0 + (size_t)u * blockDim.x;
      if (v < hi) load_grad<V>(grad, goff<V>(live, v), f[u]);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t v = v0 + (size_t)u * blockDim.x;
      if (v < hi) stage[v] = Wt::pack(f[u], scale);
    }
  }
}

// unpack [lo,hi): wire vectors -> grad(fp32)
template <bool kBf16>
HZ_DEVINL void unpack_range(float* __restrict__ grad, const uint4* __restrict__ src, size_t lo, size_t hi,
                            const int* __restrict__ live) {
  using Wt = Wire<kBf16>;
  constexpr int V = Wt::kVec;
  for (size_t v0 = lo + threadIdx.x; v0 < hi; v0 += (size_t)blockDim.x * kUnroll) {
    uint4 w[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t v = v0 + (size_t)u * blockDim.x;
      if (v < hi) w[u] = src[v];
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t v = v0 + (size_t)u * blockDim.x;
      if (v < hi) {
        float a[V];
#pragma unroll
        for (int i = 0; i < V; ++i) a[i] = 0.f;
        Wt::accum(a, w[u]);
        store_grad<V>(grad, goff<V>(live, v), a);
      }
    }
  }
}

// Optimizer epilogue of the all-reduce (kAdam): the reduced gradient never goes back to HBM as a gradient — the
// thread that holds it applies torch.optim.Adam to the parameter slice of the bucket right away (fp32 master +
// moments, bf16 shadow refresh, the reference's gradient-divergence term sum (g - g_prev)^2 with g_prev <- g, and the
// gradient clear the producers rely on).  All pointers address the bucket's slice (same element offsets as `grad`).
struct AdamFuse {
  float* p; float* m; float* v;
  __nv_bfloat16* shadow;         // may be null (fp32 compute)
  float* prev; float* diff_out;  // may be null (metric off)
  float* step;                   // [1] optimizer step counter: this call uses step+1; `bump` makes the call's last
  float lr, b1, b2, eps;         //     block store step+1 (the final bucket of the optimizer step)
  int bump;
};

struct AdamCoef { float step_size, inv_sqrt_bc2, b1, b2, eps; };

template <int V>
HZ_DEVINL void adam_apply(const AdamFuse& a, const AdamCoef& k, float* __restrict__ grad, size_t off, const float* g,
                          float& dacc) {
#pragma unroll
  for (int q = 0; q < V / 4; ++q) {
    const size_t o = off + 4 * q;
    float4 pp = *reinterpret_cast<float4*>(a.p + o);
    float4 mm = *reinterpret_cast<float4*>(a.m + o);
    float4 vv = *reinterpret_cast<float4*>(a.v + o);
    float* P = &pp.x; float* Mo = &mm.x; float* Vv = &vv.x;
    const float* G = g + 4 * q;
    if (a.prev != nullptr) {
      const float4 pr = *reinterpret_cast<const float4*>(a.prev + o);
      const float dx = G[0] - pr.x, dy = G[1] - pr.y, dz = G[2] - pr.z, dw = G[3] - pr.w;
      dacc += dx * dx + dy * dy + dz * dz + dw * dw;
      *reinterpret_cast<float4*>(a.prev + o) = make_float4(G[0], G[1], G[2], G[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = G[j];
      Mo[j] = k.b1 * Mo[j] + (1.f - k.b1) * gr;
      Vv[j] = k.b2 * Vv[j] + (1.f - k.b2) * gr * gr;
      const float denom = sqrtf(Vv[j]) * k.inv_sqrt_bc2 + k.eps;
      P[j] -= k.step_size * Mo[j] / denom;
    }
    *reinterpret_cast<float4*>(a.p + o) = pp;
    *reinterpret_cast<float4*>(a.m + o) = mm;
    *reinterpret_cast<float4*>(a.v + o) = vv;
    *reinterpret_cast<float4*>(grad + o) = make_float4(0.f, 0.f, 0.f, 0.f);     // producers only ever accumulate
    if (a.shadow != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(a.shadow + o) = pk;
    }
  }
}

// [lo,hi) of the wire buffer `src` -> Adam on the bucket (two-shot / NVLS final phase with kAdam)
template <bool kBf16>
HZ_DEVINL void unpack_adam_range(float* __restrict__ grad, const uint4* __restrict__ src, size_t lo, size_t hi,
                                 const int* __restrict__ live, const AdamFuse& ad, const AdamCoef& k, float& dacc) {
  using Wt = Wire<kBf16>;
  constexpr int V = Wt::kVec;
  for (size_t v0 = lo + threadIdx.x; v0 < hi; v0 += (size_t)blockDim.x * 2) {
    uint4 w[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const size_t v = v0 + (size_t)u * blockDim.x;
      if (v < hi) w[u] = src[v];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const size_t v = v0 + (size_t)u * blockDim.x;
      if (v < hi) {
        float a[V];
#pragma unroll
        for (int i = 0; i < V; ++i) a[i] = 0.f;
        Wt::accum(a, w[u]);
        adam_apply<V>(ad, k, grad, goff<V>(live, v), a, dacc);
      }
    }
  }
}

template <bool kBf16, int kAlgo, bool kAdam>
__global__ void __launch_bounds__(kCommThreads) allreduce_kernel(CommDev c, float* __restrict__ grad,
                                                                 size_t n, float scale,
                                                                 const int* __restrict__ live, const AdamFuse ad) {
  using Wt = Wire<kBf16>;
  constexpr int V = Wt::kVec;
  __shared__ uint32_t s_epoch;
  const int W = c.world, B = gridDim.x, b = blockIdx.x;
  const size_t nv = n / V;
  char* my = c.base[c.rank];
  // Staging/out buffers alternate by call.  The parity comes from ONE per-communicator call counter that every
  // block reads at entry and the last-finishing block bumps (ticket): a per-block counter would let a block that
  // did not exist in the previous (smaller-grid) call reuse that call's parity and overwrite a region a slow peer
  // is still reading (kernels of one communicator are stream-ordered, so the counter is stable while they run).
  const uint32_t parity = calls_of(my)[0] & 1u;
  const size_t stage_off = kFlagBytes + (size_t)parity * c.buf_bytes;
  const size_t out_off = kFlagBytes + (size_t)(2 + parity) * c.buf_bytes;
  // NVLS uses the symmetric (multicast-mapped) region instead of the IPC region for data
  char* my_data = (kAlgo == kNvls) ? c.mc_local - kFlagBytes : my;
  uint4* my_stage = reinterpret_cast<uint4*>(my_data + stage_off);
  AdamCoef ak{};
  float dacc = 0.f;
  if (kAdam) {
    const float t = ad.step[0] + 1.f;
    const float bc1 = 1.f - __powf(ad.b1, t), bc2 = 1.f - __powf(ad.b2, t);
    ak.step_size = ad.lr / bc1; ak.inv_sqrt_bc2 = rsqrtf(bc2); ak.b1 = ad.b1; ak.b2 = ad.b2; ak.eps = ad.eps;
  }

  if (kAlgo == kLL) {
    // ---- latency protocol (small buckets, e.g. the one that is only complete after the very last gradient kernel):
    //      scaled bf16 gradients are pushed as {data, epoch} words straight into slot [rank] of every peer (ONE
    //      multimem.st per 16 bytes through the NVSwitch when the region is multicast-mapped) and the W slots in local
    //      memory are polled and summed in rank order.  No staging pass, no flag barrier: ~one NVLink flight time.
    const uint32_t e = calls_of(my)[0] + 1u;
    const bool mc = c.mc_base != nullptr;
    const size_t ll_off = kFlagBytes + 4 * c.buf_bytes + (size_t)parity * (size_t)W * kLLSlotBytes;
    char* rd_base = (mc ? c.mc_local - kFlagBytes : my) + ll_off;
    const size_t per = (nv + B - 1) / B;
    const size_t lo = min((size_t)b * per, nv), hi = min(lo + per, nv);
    for (size_t v = lo + threadIdx.x; v < hi; v += blockDim.x) {
      float f[V];
      load_grad<V>(grad, goff<V>(live, v), f);
      const uint4 w = Wt::pack(f, scale);
      const uint4 w0 = make_uint4(w.x, e, w.y, e), w1 = make_uint4(w.z, e, w.w, e);
      const size_t off = ll_off + (size_t)c.rank * kLLSlotBytes + v * 32;
      if (mc) {
        mc_st(c.mc_base - kFlagBytes + off, w0);
        mc_st(c.mc_base - kFlagBytes + off + 16, w1);
      } else {
        for (int d = 0; d < W; ++d) {
          char* dst = c.base[(c.rank + d) % W] + off;
          ll_store(dst, w0);
          ll_store(dst + 16, w1);
        }
      }
    }
    const long long t0 = globaltimer_ns();
    for (size_t v = lo + threadIdx.x; v < hi; v += blockDim.x) {
      float a[V];
#pragma unroll
      for (int i = 0; i < V; ++i) a[i] = 0.f;
      for (int r = 0; r < W; ++r) {
        const char* q = rd_base + (size_t)r * kLLSlotBytes + v * 32;
        uint4 w0 = ll_load(q), w1 = ll_load(q + 16);
        while (w0.y != e || w0.w != e || w1.y != e || w1.w != e) {
          if (globaltimer_ns() - t0 > c.timeout_ns) { atomicExch(err_of(my), 1u); __threadfence_system(); __trap(); }
          w0 = ll_load(q); w1 = ll_load(q + 16);
        }
        Wt::accum(a, make_uint4(w0.x, w0.z, w1.x, w1.z));
      }
      if (kAdam) adam_apply<V>(ad, ak, grad, goff<V>(live, v), a, dacc);
      else store_grad<V>(grad, goff<V>(live, v), a);
    }
  } else {
  // ---- pack: fused 1/W scale + cast into the peer-visible staging buffer -------------------------
  if (kAlgo == kOneShot) {
    const size_t per = (nv + B - 1) / B;
    const size_t lo = min((size_t)b * per, nv), hi = min(lo + per, nv);
    pack_range<kBf16>(grad, my_stage, lo, hi, scale, live);
  } else {
    for (int r = 0; r < W; ++r) {
      size_t lo, hi;
      sub_range(nv, W, B, r, b, lo, hi);
      pack_range<kBf16>(grad, my_stage, lo, hi, scale, live);
    }
  }
  peer_block_barrier(c, &s_epoch);

  if (kAlgo == kOneShot) {
    const size_t per = (nv + B - 1) / B;
    const size_t lo = min((size_t)b * per, nv), hi = min(lo + per, nv);
    // all W copies of U vectors are requested before the first add: one NVLink round trip per U vectors per thread
    // (with one vector per trip a CTA moved ~25 GB/s and the kernel needed ~100 CTAs next to the backward kernels)
    auto body = [&](auto ucount) {
      constexpr int U = decltype(ucount)::value;
      for (size_t v0 = lo + threadIdx.x; v0 < hi; v0 += (size_t)blockDim.x * U) {
        uint4 w[U][kMaxRanks];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t v = v0 + (size_t)u * blockDim.x;
#pragma unroll
          for (int r = 0; r < kMaxRanks; ++r)
            if (r < W && v < hi) w[u][r] = reinterpret_cast<const uint4*>(c.base[r] + stage_off)[v];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t v = v0 + (size_t)u * blockDim.x;
          if (v >= hi) continue;
          float a[V];
#pragma unroll
          for (int i = 0; i < V; ++i) a[i] = 0.f;
#pragma unroll
          for (int r = 0; r < kMaxRanks; ++r)
            if (r < W) Wt::accum(a, w[u][r]);
          if (kAdam) adam_apply<V>(ad, ak, grad, goff<V>(live, v), a, dacc);
          else store_grad<V>(grad, goff<V>(live, v), a);
        }
      }
    };
    if (W <= 2) body(std::integral_constant<int, 4>{});
    else if (W <= 4) body(std::integral_constant<int, 2>{});
    else body(std::integral_constant<int, 1>{});
  } else {
    size_t lo, hi;
    sub_range(nv, W, B, c.rank, b, lo, hi);
    if (kAlgo == kTwoShot) {
      auto body = [&](auto ucount) {
        constexpr int U = decltype(ucount)::value;
        for (size_t v0 = lo + threadIdx.x; v0 < hi; v0 += (size_t)blockDim.x * U) {
          uint4 w[U][kMaxRanks];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t v = v0 + (size_t)u * blockDim.x;
#pragma unroll
            for (int r = 0; r < kMaxRanks; ++r)
              if (r < W && v < hi) w[u][r] = reinterpret_cast<const uint4*>(c.base[r] + stage_off)[v];
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t v = v0 + (size_t)u * blockDim.x;
            if (v >= hi) continue;
            float a[V];
#pragma unroll
            for (int i = 0; i < V; ++i) a[i] = 0.f;
#pragma unroll
            for (int r = 0; r < kMaxRanks; ++r)
              if (r < W) Wt::accum(a, w[u][r]);
            const uint4 o = Wt::from_acc(a);
#pragma unroll
            for (int r = 0; r < kMaxRanks; ++r)
              if (r < W) reinterpret_cast<uint4*>(c.base[r] + out_off)[v] = o;     // NVLink push
          }
        }
      };
      if (W <= 2) body(std::integral_constant<int, 4>{});
      else if (W <= 4) body(std::integral_constant<int, 2>{});
      else body(std::integral_constant<int, 1>{});
    } else {
      const char* mc_stage = c.mc_base - kFlagBytes + stage_off;
      char* mc_out = c.mc_base - kFlagBytes + out_off;
      for (size_t v0 = lo + threadIdx.x; v0 < hi; v0 += (size_t)blockDim.x * kUnroll) {
        uint4 o[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const size_t v = v0 + (size_t)u * blockDim.x;
          if (v < hi) o[u] = Wt::mc_ld_reduce(mc_stage + v * 16);   // reduced inside the NVSwitch
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const size_t v = v0 + (size_t)u * blockDim.x;
          if (v < hi) mc_st(mc_out + v * 16, o[u]);                 // broadcast by the NVSwitch
        }
      }
    }
    peer_block_barrier(c, &s_epoch);
    const uint4* my_out = reinterpret_cast<const uint4*>(my_data + out_off);
    for (int r = 0; r < W; ++r) {
      size_t l2, h2;
      sub_range(nv, W, B, r, b, l2, h2);
      if (kAdam) unpack_adam_range<kBf16>(grad, my_out, l2, h2, live, ad, ak, dacc);
      else unpack_range<kBf16>(grad, my_out, l2, h2, live);
    }
  }
  }   // staged algorithms
  if (kAdam && ad.prev != nullptr && ad.diff_out != nullptr) {
    __shared__ float wsum[kCommThreads / 32];
    dacc = warp_sum(dacc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = dacc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float tt = threadIdx.x < kCommThreads / 32 ? wsum[threadIdx.x] : 0.f;
      tt = warp_sum(tt);
      if (threadIdx.x == 0) atomicAdd(ad.diff_out, tt);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    uint32_t* ticket = calls_of(my) + 1;
    if (atomicAdd(ticket, 1u) == (uint32_t)(B - 1)) {      // last block of this call
      *ticket = 0u;
      calls_of(my)[0] += 1u;
      if (kAdam && ad.bump) ad.step[0] += 1.f;             // every block has read the counter long ago (barriers above)
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// ZeRO-1 (optimizer-state sharding) step of one gradient bucket as ONE kernel: the two-shot all-reduce with the
// optimizer in the middle.  Rank r owns slice r of the bucket's wire vectors — fp32 master values and both Adam
// moments exist only there (m, v: 2·P/W floats per rank instead of 2·P):
//     pack    : every slice of grad(fp32)·(1/W) -> bf16 into the local staging buffer, gradient cleared in the same pass
//     barrier
//     own slice: sum the W staged copies in rank order (fp32), Adam on the owned master slice and moment shards,
//                new parameters -> bf16 -> pushed into the `out` buffer of EVERY rank (NVLink stores)
//     barrier
//     unpack  : out -> the local bf16 parameter shadow the compute kernels read (all slices)
// Wire traffic is that of a bf16 all-reduce (gradients in, bf16 parameters out); the optimizer costs 1/W of the
// replicated pass and there is no separate reduce-scatter / all-gather / shadow-refresh kernel (the NCCL formulation in
// parallel/zero.py is the baseline).  The reference replicates its optimizer state (SURVEY §2.3 "ZeRO: NO").
// The gradient-divergence term sum (g - g_prev)^2 is accumulated per slice and exchanged through per-(rank, block)
// slots next to the flags, so every rank logs the whole-gradient value without an extra collective.
struct ZeroFuse {
  float* p;                        // fp32 master, the bucket's slice of the flat buffer (authoritative on the owner only)
  float* m; float* v;              // this rank's moment shards of the bucket: [ceil(nv / W) * 8]
  __nv_bfloat16* shadow;           // bf16 parameters, the bucket's slice of the flat shadow
  float* prev; float* diff_out;    // optional: previous reduced gradient of the own slice (shard-sized) / accumulator [1]
  float* step;                     // as AdamFuse
  float lr, b1, b2, eps;
  int bump;
};
constexpr int kZeroMaxBlocks = 64;
HZ_DEVINL float* zdiff_of(char* base) { return reinterpret_cast<float*>(base + 53248); }   // [2][kMaxRanks][kZeroMaxBlocks]

__global__ void __launch_bounds__(kCommThreads) zero1_kernel(CommDev c, float* __restrict__ grad, size_t n, float scale,
                                                             const int* __restrict__ live, const ZeroFuse z) {
  using Wt = Wire<true>;
  constexpr int V = Wt::kVec;
  __shared__ uint32_t s_epoch;
  __shared__ float wsum[kCommThreads / 32];
  const int W = c.world, B = gridDim.x, b = blockIdx.x;
  const size_t nv = n / V;
  char* my = c.base[c.rank];
  const uint32_t parity = calls_of(my)[0] & 1u;        // per-communicator call counter, see allreduce_kernel
  const size_t stage_off = kFlagBytes + (size_t)parity * c.buf_bytes;
  const size_t out_off = kFlagBytes + (size_t)(2 + parity) * c.buf_bytes;
  uint4* my_stage = reinterpret_cast<uint4*>(my + stage_off);
  AdamCoef ak;
  {
    const float t = z.step[0] + 1.f;
    const float bc1 = 1.f - __powf(z.b1, t), bc2 = 1.f - __powf(z.b2, t);
    ak.step_size = z.lr / bc1; ak.inv_sqrt_bc2 = rsqrtf(bc2); ak.b1 = z.b1; ak.b2 = z.b2; ak.eps = z.eps;
  }

  // ---- pack all slices, clear the gradient (producers only ever accumulate into it)
  for (int r = 0; r < W; ++r) {
    size_t lo, hi;
    sub_range(nv, W, B, r, b, lo, hi);
    for (size_t v0 = lo + threadIdx.x; v0 < hi; v0 += (size_t)blockDim.x * 4) {
      float f[4][V];
      size_t off[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t v = v0 + (size_t)u * blockDim.x;
        if (v < hi) { off[u] = goff<V>(live, v); load_grad<V>(grad, off[u], f[u]); }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t v = v0 + (size_t)u * blockDim.x;
        if (v < hi) {
          my_stage[v] = Wt::pack(f[u], scale);
          reinterpret_cast<float4*>(grad + off[u])[0] = make_float4(0.f, 0.f, 0.f, 0.f);
          reinterpret_cast<float4*>(grad + off[u])[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
  peer_block_barrier(c, &s_epoch);

  // ---- own slice: reduce, Adam on the shard, push the new bf16 parameters to every rank
  float dacc = 0.f;
  {
    size_t lo, hi;
    sub_range(nv, W, B, c.rank, b, lo, hi);
    const size_t q = (nv + W - 1) / W;
    const size_t slice_lo = min((size_t)c.rank * q, nv);
    for (size_t v = lo + threadIdx.x; v < hi; v += blockDim.x) {
      uint4 w[kMaxRanks];
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)
        if (r < W) w[r] = reinterpret_cast<const uint4*>(c.base[r] + stage_off)[v];
      float a[V];
#pragma unroll
      for (int i = 0; i < V; ++i) a[i] = 0.f;
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)
        if (r < W) Wt::accum(a, w[r]);                                   // rank order: identical on every rank
      const size_t off = goff<V>(live, v);                                // element offset in the bucket
      const size_t so = (v - slice_lo) * V;                               // element offset in the shard
      float pn[V];
#pragma unroll
      for (int h = 0; h < V / 4; ++h) {
        float4 pp = *reinterpret_cast<float4*>(z.p + off + 4 * h);
        float4 mm = *reinterpret_cast<float4*>(z.m + so + 4 * h);
        float4 vv = *reinterpret_cast<float4*>(z.v + so + 4 * h);
        float* P = &pp.x; float* Mo = &mm.x; float* Vv = &vv.x;
        const float* G = a + 4 * h;
        if (z.prev != nullptr) {
          const float4 pr = *reinterpret_cast<const float4*>(z.prev + so + 4 * h);
          const float dx = G[0] - pr.x, dy = G[1] - pr.y, dz = G[2] - pr.z, dw = G[3] - pr.w;
          dacc += dx * dx + dy * dy + dz * dz + dw * dw;
          *reinterpret_cast<float4*>(z.prev + so + 4 * h) = make_float4(G[0], G[1], G[2], G[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float gr = G[j];
          Mo[j] = ak.b1 * Mo[j] + (1.f - ak.b1) * gr;
          Vv[j] = ak.b2 * Vv[j] + (1.f - ak.b2) * gr * gr;
          const float denom = sqrtf(Vv[j]) * ak.inv_sqrt_bc2 + ak.eps;
          P[j] -= ak.step_size * Mo[j] / denom;
          pn[4 * h + j] = P[j];
        }
        *reinterpret_cast<float4*>(z.p + off + 4 * h) = pp;
        *reinterpret_cast<float4*>(z.m + so + 4 * h) = mm;
        *reinterpret_cast<float4*>(z.v + so + 4 * h) = vv;
      }
      const uint4 o = Wt::from_acc(pn);
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)
        if (r < W) reinterpret_cast<uint4*>(c.base[r] + out_off)[v] = o;   // NVLink push (own copy included)
    }
  }
  const bool want_diff = z.prev != nullptr && z.diff_out != nullptr;
  if (want_diff) {
    dacc = warp_sum(dacc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = dacc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float tt = threadIdx.x < kCommThreads / 32 ? wsum[threadIdx.x] : 0.f;
      tt = warp_sum(tt);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
          if (r < W) zdiff_of(c.base[r])[((size_t)parity * kMaxRanks + c.rank) * kZeroMaxBlocks + b] = tt;
      }
    }
  }
  peer_block_barrier(c, &s_epoch);

  // ---- unpack: every rank's new parameters -> the local bf16 shadow
  {
    const uint4* my_out = reinterpret_cast<const uint4*>(my + out_off);
    for (int r = 0; r < W; ++r) {
      size_t lo, hi;
      sub_range(nv, W, B, r, b, lo, hi);
      for (size_t v0 = lo + threadIdx.x; v0 < hi; v0 += (size_t)blockDim.x * 4) {
        uint4 w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t v = v0 + (size_t)u * blockDim.x;
          if (v < hi) w[u] = my_out[v];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t v = v0 + (size_t)u * blockDim.x;
          if (v < hi) *reinterpret_cast<uint4*>(z.shadow + goff<V>(live, v)) = w[u];
        }
      }
    }
  }
  if (want_diff && threadIdx.x == 0) {
    float tot = 0.f;
    for (int r = 0; r < W; ++r) tot += zdiff_of(my)[((size_t)parity * kMaxRanks + r) * kZeroMaxBlocks + b];
    atomicAdd(z.diff_out, tot);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    uint32_t* ticket = calls_of(my) + 1;
    if (atomicAdd(ticket, 1u) == (uint32_t)(B - 1)) {      // last block of this call
      *ticket = 0u;
      calls_of(my)[0] += 1u;
      if (z.bump) z.step[0] += 1.f;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One-shot all-reduce whose reduce phase PULLS with bulk async copies (algo "bulk"): instead of every thread loading one
// 16-byte vector from each peer into registers (~32 KB in flight per CTA: the staged kernels need ~100 CTAs to fill
// NVLink, and lose bandwidth under the 32-CTA background cap), one elected thread streams 8 KB chunks of all W peers'
// staging buffers into a 3-deep shared-memory ring with cp.async.bulk (completion on an mbarrier, W x 8 KB per stage:
// up to 192 KB in flight per CTA), and the CTA sums from shared memory in rank order.  Same pack / flag barrier /
// parity / ticket protocol and the same arithmetic (bit-identical results) as allreduce_kernel<.., kOneShot, false>.
// The copy engine path is the one the fused tensor-parallel kernels' pull uses (csrc/tp_fused.cu, measured there).
// Selected explicitly (`--allreduce peer-bulk` / PeerComm.allreduce(.., "bulk", ..)); not part of the auto rule yet:
// written after the round's GPU budget was spent, untimed.
// ------------------------------------------------------------------------------------------------
constexpr int kBulkChunkVec = 512;          // 16-byte vectors per peer and ring stage (8 KB) = one per thread
constexpr int kBulkStages = 3;
HZ_DEVINL void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool kBf16>
__global__ void __launch_bounds__(kCommThreads) allreduce_bulk_kernel(CommDev c, float* __restrict__ grad, size_t n,
                                                                      float scale, const int* __restrict__ live) {
  using Wt = Wire<kBf16>;
  constexpr int V = Wt::kVec;
  static_assert(kBulkChunkVec == kCommThreads, "one vector per thread and chunk");
  extern __shared__ __align__(128) uint8_t bulk_ring[];          // [kBulkStages][W][kBulkChunkVec] uint4
  __shared__ uint32_t s_epoch;
  __shared__ __align__(8) uint64_t full[kBulkStages], empty[kBulkStages];
  const int W = c.world, B = gridDim.x, b = blockIdx.x;
  const size_t nv = n / V;
  char* my = c.base[c.rank];
  const uint32_t parity = calls_of(my)[0] & 1u;
  const size_t stage_off = kFlagBytes + (size_t)parity * c.buf_bytes;
  uint4* my_stage = reinterpret_cast<uint4*>(my + stage_off);
  uint4* ring = reinterpret_cast<uint4*>(bulk_ring);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kBulkStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], kCommThreads / 32); }
    tc::fence_barrier_init();
  }
  const size_t per = (nv + B - 1) / B;
  const size_t lo = min((size_t)b * per, nv), hi = min(lo + per, nv);
  pack_range<kBf16>(grad, my_stage, lo, hi, scale, live);
  peer_block_barrier(c, &s_epoch);                       // (its __syncthreads also publish the mbarrier inits)

  const int nchunks = (int)((hi - lo + kBulkChunkVec - 1) / kBulkChunkVec);
  auto issue = [&](int ck) {                             // thread 0: request chunk ck of every peer into ring slot ck % S
    const int s = ck % kBulkStages;
    const uint32_t ph = (ck / kBulkStages) & 1;
    const size_t v0 = lo + (size_t)ck * kBulkChunkVec;
    const uint32_t bytes = (uint32_t)(min((size_t)kBulkChunkVec, hi - v0) * 16);
    tc::mbar_wait(&empty[s], ph ^ 1);                    // every warp is done with the previous content of the slot
    tc::mbar_arrive_expect_tx(&full[s], (uint32_t)W * bytes);
    for (int d = 0; d < W; ++d) {
      const int r = (c.rank + d) % W;                    // own copy first, peers staggered
      bulk_g2s(ring + ((size_t)s * W + r) * kBulkChunkVec, c.base[r] + stage_off + v0 * 16, bytes, &full[s]);
    }
  };
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");     // the peers' staging stores (seen through the flag barrier) -> async proxy
    for (int ck = 0; ck < kBulkStages - 1 && ck < nchunks; ++ck) issue(ck);
  }
  for (int ck = 0; ck < nchunks; ++ck) {
    if (threadIdx.x == 0 && ck + kBulkStages - 1 < nchunks) issue(ck + kBulkStages - 1);
    __syncwarp();
    const int s = ck % kBulkStages;
    const uint32_t ph = (ck / kBulkStages) & 1;
    tc::mbar_wait(&full[s], ph);
    const size_t v = lo + (size_t)ck * kBulkChunkVec + threadIdx.x;
    if (v < hi) {
      float a[V];
#pragma unroll
      for (int i = 0; i < V; ++i) a[i] = 0.f;
      for (int r = 0; r < W; ++r) Wt::accum(a, ring[((size_t)s * W + r) * kBulkChunkVec + threadIdx.x]);   // rank order
      store_grad<V>(grad, goff<V>(live, v), a);
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) tc::mbar_arrive(&empty[s]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    uint32_t* ticket = calls_of(my) + 1;
    if (atomicAdd(ticket, 1u) == (uint32_t)(B - 1)) {      // last block of this call
      *ticket = 0u;
      calls_of(my)[0] += 1u;
      __threadfence();
    }
  }
}

__global__ void barrier_kernel(CommDev c, long long* stamp_ns) {
  __shared__ uint32_t s_epoch;
  const long long t0 = globaltimer_ns();
  peer_block_barrier(c, &s_epoch);
  if (threadIdx.x == 0 && stamp_ns != nullptr) {
    stamp_ns[0] = t0;                       // arrival
    stamp_ns[1] = globaltimer_ns();         // release: (release - arrival) = device-side idle time
  }
}

}  // namespace hz

// ================================================================================================
// host side
// ================================================================================================
struct HzComm {
  hz::CommDev dev;
  int device;
  int max_blocks;
  int block_cap;               // 0 = max_blocks; else a (smaller) cap for the following calls
  size_t region_bytes;
  char* local;                 // cudaMalloc'd region
  size_t heap_off, heap_bytes; // symmetric heap (optional)
  size_t ll_bytes;             // latency-protocol slot area behind the staging buffers
  bool imported[hz::kMaxRanks];
  bool local_group;
};

#define HZ_CUDA(x)                                                                              \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "[hz comm] %s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, \
              __LINE__);                                                                        \
      return -1;                                                                                \
    }                                                                                           \
  } while (0)

extern "C" {

HzComm* hz_comm_create2(int rank, int world, int device, size_t max_wire_bytes, int max_blocks, size_t heap_bytes);
HzComm* hz_comm_create(int rank, int world, int device, size_t max_wire_bytes, int max_blocks) {
  return hz_comm_create2(rank, world, device, max_wire_bytes, max_blocks, 0);
}

// heap_bytes > 0 appends a zero-initialised symmetric heap (same offsets on every rank) used by the fused
// tensor-parallel kernels for peer-writable activations, partial-tile slots and flags.
HzComm* hz_comm_create2(int rank, int world, int device, size_t max_wire_bytes, int max_blocks, size_t heap_bytes) {
  if (world > hz::kMaxRanks || max_blocks * world * 4 > 32768 || max_blocks * 4 > 12288) return nullptr;
  HzComm* c = new HzComm();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->max_blocks = max_blocks;
  cudaSetDevice(device);
  const size_t buf = (max_wire_bytes + 255) / 256 * 256;
  c->dev.buf_bytes = buf;
  c->dev.rank = rank;
  c->dev.world = world;
  {
    const char* e = getenv("HZ_COMM_TIMEOUT_S");
    const double sec = e ? atof(e) : 0.0;
    c->dev.timeout_ns = sec > 0.0 ? (long long)(sec * 1e9) : hz::kDefaultSpinTimeoutNs;
  }
  c->ll_bytes = 2 * (size_t)world * hz::kLLSlotBytes;              // latency-protocol slots [2 parities][world]
  c->heap_off = hz::kFlagBytes + 4 * buf + c->ll_bytes;
  c->heap_bytes = (heap_bytes + 1023) / 1024 * 1024;
  c->region_bytes = c->heap_off + c->heap_bytes;
  if (cudaMalloc(&c->local, c->region_bytes) != cudaSuccess) { delete c; return nullptr; }
  cudaMemset(c->local, 0, hz::kFlagBytes);
  cudaMemset(c->local + hz::kFlagBytes + 4 * buf, 0, c->ll_bytes);      // LL flags start at 0 (epochs start at 1)
  if (c->heap_bytes) cudaMemset(c->local + c->heap_off, 0, c->heap_bytes);
  c->dev.base[rank] = c->local;
  cudaDeviceSynchronize();
  return c;
}

int hz_comm_export(HzComm* c, void* handle64) {
  cudaIpcMemHandle_t h;
  HZ_CUDA(cudaIpcGetMemHandle(&h, c->local));
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(handle64, &h, 64);
  return 0;
}

int hz_comm_import(HzComm* c, const void* handles) {
  cudaSetDevice(c->device);
  for (int r = 0; r < c->dev.world; ++r) {
    if (r == c->dev.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * r, 64);
    void* p = nullptr;
    HZ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->dev.base[r] = (char*)p;
    c->imported[r] = true;
  }
  return 0;
}

// same-process "virtual ranks" on one device (single-GPU tests of the multi-rank protocol)
int hz_comm_link_local(HzComm** comms, int world) {
  for (int i = 0; i < world; ++i)
    for (int r = 0; r < world; ++r) comms[i]->dev.base[r] = comms[r]->local;
  for (int i = 0; i < world; ++i) comms[i]->local_group = true;
  return 0;
}

size_t hz_comm_symm_bytes(HzComm* c) { return 4 * c->dev.buf_bytes + c->ll_bytes; }

// mc_ptr / local_ptr: multicast and local mapping of a zero-initialised symmetric buffer of hz_comm_symm_bytes()
void hz_comm_set_multicast(HzComm* c, void* mc_ptr, void* local_ptr, size_t bytes) {
  if (bytes < 4 * c->dev.buf_bytes + c->ll_bytes) return;
  c->dev.mc_base = (char*)mc_ptr;
  c->dev.mc_local = (char*)local_ptr;
}

int hz_comm_blocks_for(HzComm* c, size_t n, int algo, int wire_bf16) {
  const int V = wire_bf16 ? 8 : 4;
  const size_t nv = n / V;
  // ~4 vectors (64 B) per thread and phase: enough CTAs to pull HBM + NVLink bandwidth on big buckets,
  // a single CTA for latency-bound small ones
  size_t want = (nv + (size_t)hz::kCommThreads * 4 - 1) / ((size_t)hz::kCommThreads * 4);
  if (algo == hz::kLL) want = (nv + (size_t)hz::kCommThreads * 2 - 1) / ((size_t)hz::kCommThreads * 2);   // latency first
  if (want < 1) want = 1;
  if (want > (size_t)c->max_blocks) want = c->max_blocks;
  if (c->block_cap > 0 && want > (size_t)c->block_cap) want = c->block_cap;
  (void)algo;
  return (int)want;
}

// n = number of gradient elements on the wire (= all of them, or 64 * #live blocks when `live` is given)
static int comm_allreduce_impl(HzComm* c, float* grad, size_t n, int algo, int wire_bf16, float scale, const int* live,
                               const hz::AdamFuse* adam, cudaStream_t st) {
  const int V = wire_bf16 ? 8 : 4;
  if (n % V != 0) return -2;
  if (n * (wire_bf16 ? 2 : 4) > c->dev.buf_bytes) return -3;
  if (algo == hz::kNvls && c->dev.mc_base == nullptr) return -4;
  if (algo == hz::kLL && (!wire_bf16 || n > hz::kLLMaxElems)) return -5;
  const int blocks = hz_comm_blocks_for(c, n, algo, wire_bf16);
  if (algo == hz::kBulk) {
    if (adam) return -7;                                   // plain reduction only
    const size_t smem = (size_t)hz::kBulkStages * c->dev.world * hz::kBulkChunkVec * 16;
    static bool attr = [] {
      return cudaFuncSetAttribute(hz::allreduce_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  hz::kBulkStages * hz::kMaxRanks * hz::kBulkChunkVec * 16) == cudaSuccess &&
             cudaFuncSetAttribute(hz::allreduce_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  hz::kBulkStages * hz::kMaxRanks * hz::kBulkChunkVec * 16) == cudaSuccess;
    }();
    if (!attr) return -8;
    if (wire_bf16) hz::allreduce_bulk_kernel<true><<<blocks, hz::kCommThreads, smem, st>>>(c->dev, grad, n, scale, live);
    else hz::allreduce_bulk_kernel<false><<<blocks, hz::kCommThreads, smem, st>>>(c->dev, grad, n, scale, live);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  hz::AdamFuse ad;
  memset(&ad, 0, sizeof(ad));
  if (adam) ad = *adam;
#define HZ_LAUNCH(BF, AL)                                                                                     \
  do {                                                                                                        \
    if (adam) hz::allreduce_kernel<BF, AL, true><<<blocks, hz::kCommThreads, 0, st>>>(c->dev, grad, n, scale, live, ad);  \
    else hz::allreduce_kernel<BF, AL, false><<<blocks, hz::kCommThreads, 0, st>>>(c->dev, grad, n, scale, live, ad);      \
  } while (0)
  if (wire_bf16) {
    if (algo == hz::kOneShot) HZ_LAUNCH(true, hz::kOneShot);
    else if (algo == hz::kTwoShot) HZ_LAUNCH(true, hz::kTwoShot);
    else if (algo == hz::kLL) HZ_LAUNCH(true, hz::kLL);
    else HZ_LAUNCH(true, hz::kNvls);
  } else {
    if (algo == hz::kOneShot) HZ_LAUNCH(false, hz::kOneShot);
    else if (algo == hz::kTwoShot) HZ_LAUNCH(false, hz::kTwoShot);
    else HZ_LAUNCH(false, hz::kNvls);
  }
#undef HZ_LAUNCH
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int hz_comm_allreduce(HzComm* c, float* grad, size_t n, int algo, int wire_bf16, float scale, const int* live,
                      cudaStream_t st) {
  return comm_allreduce_impl(c, grad, n, algo, wire_bf16, scale, live, nullptr, st);
}

// All-reduce (average) of a gradient bucket with the Adam update of the bucket's parameters fused into the
// reduction's final phase (see AdamFuse).  `bump`: this is the last bucket of the optimizer step.
int hz_comm_allreduce_adam(HzComm* c, float* grad, size_t n, int algo, int wire_bf16, float scale, const int* live,
                           float* p, float* m, float* v, void* shadow, float* prev, float* diff_out, float* step,
                           float lr, float b1, float b2, float eps, int bump, cudaStream_t st) {
  hz::AdamFuse ad;
  ad.p = p; ad.m = m; ad.v = v; ad.shadow = (__nv_bfloat16*)shadow; ad.prev = prev; ad.diff_out = diff_out;
  ad.step = step; ad.lr = lr; ad.b1 = b1; ad.b2 = b2; ad.eps = eps; ad.bump = bump;
  return comm_allreduce_impl(c, grad, n, algo, wire_bf16, scale, live, &ad, st);
}

// ZeRO-1 step of one bucket (zero1_kernel): n wire elements (all of the bucket, or 64 per live block), bf16 wire and
// bf16 parameter shadow only.  m / v / prev are THIS rank's shards: hz_comm_zero1_shard(n, world) elements each.
size_t hz_comm_zero1_shard(size_t n, int world) { return ((n / 8 + (size_t)world - 1) / (size_t)world) * 8; }

int hz_comm_zero1_step(HzComm* c, float* grad, size_t n, float scale, const int* live, float* p, float* m, float* v,
                       void* shadow, float* prev, float* diff_out, float* step, float lr, float b1, float b2, float eps,
                       int bump, cudaStream_t st) {
  if (n % 8 != 0) return -2;
  if (n * 2 > c->dev.buf_bytes) return -3;
  if (shadow == nullptr || p == nullptr || m == nullptr || v == nullptr || step == nullptr) return -6;
  int blocks = hz_comm_blocks_for(c, n, hz::kTwoShot, 1);
  if (blocks > hz::kZeroMaxBlocks) blocks = hz::kZeroMaxBlocks;
  hz::ZeroFuse z;
  z.p = p; z.m = m; z.v = v; z.shadow = (__nv_bfloat16*)shadow; z.prev = prev; z.diff_out = diff_out; z.step = step;
  z.lr = lr; z.b1 = b1; z.b2 = b2; z.eps = eps; z.bump = bump;
  hz::zero1_kernel<<<blocks, hz::kCommThreads, 0, st>>>(c->dev, grad, n, scale, live, z);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Grid cap for the following collectives (same value on every rank: blocks pair up with their peers).  A bucket
// whose result is not needed for a long time is reduced by a few CTAs so the backward kernels keep the SMs.
void hz_comm_set_block_cap(HzComm* c, int cap) { c->block_cap = cap > 0 ? cap : 0; }

int hz_comm_barrier(HzComm* c, long long* stamps, cudaStream_t st) {
  hz::barrier_kernel<<<1, 32, 0, st>>>(c->dev, stamps);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

char* hz_comm_heap_base(HzComm* c, int r) { return c->dev.base[r] ? c->dev.base[r] + c->heap_off : nullptr; }
size_t hz_comm_heap_bytes(HzComm* c) { return c->heap_bytes; }

int hz_comm_error(HzComm* c) {
  uint32_t e = 0;
  cudaMemcpy(&e, c->local + 61440, 4, cudaMemcpyDeviceToHost);
  return (int)e;
}

void hz_comm_destroy(HzComm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (int r = 0; r < c->dev.world; ++r)
    if (c->imported[r] && c->dev.base[r]) cudaIpcCloseMemHandle(c->dev.base[r]);
  if (c->local) cudaFree(c->local);
  delete c;
}

}  // extern "C"

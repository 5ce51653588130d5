// tcgen05 / TMEM / TMA implicit-GEMM convolutions for NHWC bf16.
//
// Replaces every torchvision conv of ResNet-18 (data_parallel_train.py:198 — ATen CPU convs in the
// reference) in forward, dgrad and wgrad:
//
//   igemm_kernel  (fwd, dgrad, dense GEMM):   D[128 x BLOCK_N] = sum over (tap, 64-channel block) A * B
//       A : activation tile, K-major.  One 4-D TMA box (64 ch, BW, BH, BN) per (tap, channel block) is
//           exactly a 128-row x 128-byte SWIZZLE_128B operand tile; padding comes for free from TMA
//           out-of-bounds zero fill (negative / overflowing coordinates).  Stride-2 convs read four
//           "parity views" of the input (strided tensor maps), so no im2col buffer ever exists.
//       B : weights [Cout, R*S*Cin]: K-major for fwd, MN-major for dgrad (no transposed weight copy).
//       D : fp32 in TMEM; epilogue tcgen05.ld -> bf16 -> smem -> coalesced stores, plus the BatchNorm
//           partial sums (sum y, sum y^2 per channel) so BN statistics cost no extra pass.
//       Stride-2 dgrad is 4 output-parity classes (blockIdx.z), each a stride-1 gather over its taps.
//       Taps whose receptive field is entirely padding are dropped on the host (layer4's 1x1 maps use
//       only the centre tap: 9x less work).
//
//   wgrad_kernel:  dW[Cout, tap, Cin] = sum over pixels dY^T * X_tap.  Both operands MN-major straight
//       from their NHWC tensors (64-pixel TMA boxes), split-K over pixel blocks (blockIdx.z), fp32 result
//       written / accumulated directly into the flat gradient bucket (no flatten copy).
//
// Warp roles (128 threads, one output tile per CTA): warp0/lane0 TMA producer, warp1/lane0 MMA issuer,
// all four warps epilogue (each owns its 32 TMEM lanes).  6-stage smem ring, mbarrier full/empty pairs,
// tcgen05.commit releases stages and signals the epilogue.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "igemm_common.cuh"
#include "launchers.h"
#include "tc05.cuh"

namespace hz {


struct IgemmParams {
  TapList cls[4];
  long long cls_out_off[4];
  long long out_n_stride, out_h_stride, out_w_stride;   // elements
  int num_classes, cblocks;
  int BN, BH, BW, tiles_per_img;
  int n_images;
  int ncols;                 // valid output columns (Cout / Cin total)
  int splits;                // split-K factor (blockIdx.z = class * splits + split)
  int cluster;               // 1: the `splits` CTAs of a tile form a thread-block cluster and reduce through DSMEM
  int prefetch_b;            // 1: weight tiles of the first stages are requested BEFORE griddepcontrol.wait
  float* ws;                 // fp32 split-K workspace [tiles][128][BLOCK_N], all-zero between launches
  unsigned* sem;             // per-tile arrival counters, all-zero between launches
  __nv_bfloat16* out;
  const __nv_bfloat16* addend;   // optional, same layout as out: out = tile + addend (residual-gradient fusion)
  float* stats;              // [2*ncols] or null
  // ---- fused BatchNorm(batch stats) + residual + ReLU (forward): after the tile's Σy, Σy² are in `stats`, a
  //      device-wide barrier makes the totals final and the CTA normalises the tile it still holds in shared memory
  __nv_bfloat16* bn_out;         // null = no fusion
  const __nv_bfloat16* bn_residual;
  const float* bn_gamma; const float* bn_beta;
  float* bn_mean; float* bn_invstd;      // saved for backward
  float* bn_rmean; float* bn_rvar;       // running statistics (may be null)
  unsigned* bn_counter;          // zero before the launch
  float bn_inv_count, bn_unbias, bn_eps, bn_momentum;
  int bn_relu;
  long long* dbg;                // optional: per-CTA clock64 stamps of the kernel's phases (tools/conv_timeline.py)
  // ---- dgrad only (kBnBwd instantiation): this dgrad's output IS the upstream gradient of the producing layer's
  //      BatchNorm, so its backward sums  (sum g, sum g*xhat),  g = dx * [bn_out > 0],  are taken from the registers
  //      that store dx (into `stats`, pre-zeroed) — the separate channel_reduce pass over dx / out / y_raw disappears
  const __nv_bfloat16* bnb_out;  // the producing layer's BN output (ReLU mask), same layout as `out`; null = no ReLU
  const __nv_bfloat16* bnb_yraw; // the producing layer's raw conv output
  const float* bnb_mean; const float* bnb_invstd;     // [ncols]
  int bnb_cap6;                  // the producing layer's activation is ReLU6: the mask is 0 < out < 6
};

HZ_DEVINL unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Cluster split-K reduction of this CTA's row slice: the fp32 partial tiles of all SX CTAs are read through
// distributed shared memory.  Measured with the in-kernel phase stamps (tools/conv_timeline.py): issuing the SX
// remote loads of an item one after the other, each feeding an add, serialises ~16 DSMEM round trips (2.3 us of a
// 6 us kernel) — here all of a thread's remote loads are in flight before the first add.
template <int SX, int BLOCK_N, int RED_LD, int STAGING_LD>
HZ_DEVINL void cluster_reduce_rows(uint32_t red_base, int row_lo, __nv_bfloat16* staging) {
  constexpr int kChunks = BLOCK_N / 4;
  constexpr int kItems = (kTileM / SX) * kChunks / 128;
  static_assert(kItems >= 1 && kItems * SX <= 32, "remote loads in flight per thread");
  float4 v[kItems][SX];
#pragma unroll
  for (int it = 0; it < kItems; ++it) {
    const int idx = threadIdx.x + it * 128;
    const int rr = row_lo + idx / kChunks, ch = idx % kChunks;
    const uint32_t off = (uint32_t)((rr * RED_LD + ch * 4) * 4);
#pragma unroll
    for (int s2 = 0; s2 < SX; ++s2) v[it][s2] = ld_dsmem_f4_nb(map_to_cta(red_base + off, (uint32_t)s2));
  }
#pragma unroll
  for (int it = 0; it < kItems; ++it) {
    const int idx = threadIdx.x + it * 128;
    const int rr = row_lo + idx / kChunks, ch = idx % kChunks;
    float4 acc = v[it][0];
#pragma unroll
    for (int s2 = 1; s2 < SX; ++s2) {                                     // fixed order: deterministic
      acc.x += v[it][s2].x; acc.y += v[it][s2].y; acc.z += v[it][s2].z; acc.w += v[it][s2].w;
    }
    __nv_bfloat162 lo2 = __floats2bfloat162_rn(acc.x, acc.y), hi2 = __floats2bfloat162_rn(acc.z, acc.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo2);
    pk.y = *reinterpret_cast<uint32_t*>(&hi2);
    *reinterpret_cast<uint2*>(staging + rr * STAGING_LD + ch * 4) = pk;
  }
}

template <int BLOCK_N, bool B_MN, bool kBnBwd = false>
__global__ void __launch_bounds__(128) igemm_kernel(const __grid_constant__ AMaps amaps,
                                                    const __grid_constant__ CUtensorMap bmap,
                                                    const __grid_constant__ IgemmParams p) {
  using S = IgemmSmem<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  pdl_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, nt = blockIdx.y;
  long long* dbg = p.dbg ? p.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 : nullptr;
#define HZ_STAMP(i) do { if (dbg != nullptr) dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 0) HZ_STAMP(0);                     // kernel entry
  const int cls = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const TapList& taps = p.cls[cls];
  const int n0 = (p.BN == 1) ? mt / p.tiles_per_img : mt * p.BN;
  const int h0 = (p.BN == 1) ? (mt % p.tiles_per_img) * p.BH : 0;
  // split-K: one SM pulls only ~80 GB/s out of L2, so the K loop of a tile is spread over `splits` CTAs
  const int k_total = taps.n * p.cblocks;
  const int k_per = (k_total + p.splits - 1) / p.splits;
  const int k_lo = min(split * k_per, k_total);
  const int k_iters = min(k_lo + k_per, k_total) - k_lo;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) tc::prefetch_tmap(&amaps.m[i]);
    tc::prefetch_tmap(&bmap);
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, BLOCK_N);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;

  auto load_b = [&](int it) {
    const int t = (k_lo + it) / p.cblocks, cb = (k_lo + it) % p.cblocks;
    const int s = it % kStages;
    uint8_t* sb = smem + s * S::kStageBytes + kABytes;
    if (!B_MN) {
      tc::tma_load_2d(sb, &bmap, &full[s], taps.bk[t] + cb * kKBlock, nt * BLOCK_N);
    } else {
#pragma unroll
      for (int j = 0; j < BLOCK_N / 64; ++j)
        tc::tma_load_2d(sb + j * 8192, &bmap, &full[s], taps.bk[t] + nt * BLOCK_N + j * 64, cb * kKBlock);
    }
  };
  // The B operand is the weight tensor: it was last written by the optimizer at least two kernels upstream, so
  // (unlike A, the previous kernel's output) it may be requested before griddepcontrol.wait — the first ring of
  // weight tiles (cold in L2: DRAM latency) streams in under the previous kernel's tail.
  int n_pre = 0;
  if (warp == 0 && lane == 0 && p.prefetch_b) {
    n_pre = min(k_iters, kStages);
    for (int it = 0; it < n_pre; ++it) {
      tc::mbar_arrive_expect_tx(&full[it], S::kStageBytes);
      load_b(it);
    }
  }
  if (threadIdx.x == 0) HZ_STAMP(1);                     // prologue done (barriers, TMEM, first weight requests)
  pdl_wait();          // everything above overlapped the tail of the previous kernel
  if (threadIdx.x == 0) HZ_STAMP(2);                     // upstream kernel complete

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    for (int it = 0; it < k_iters; ++it) {
      const int t = (k_lo + it) / p.cblocks, cb = (k_lo + it) % p.cblocks;
      const CUtensorMap* am = &amaps.m[taps.map[t]];
      const int cw = taps.dw[t], ch = h0 + taps.dh[t];
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      uint8_t* sa = smem + s * S::kStageBytes;
      if (it >= n_pre) {
        tc::mbar_wait(&empty[s], ph ^ 1);
        tc::mbar_arrive_expect_tx(&full[s], S::kStageBytes);
      }
      tc::tma_load_4d(sa, am, &full[s], cb * kKBlock, cw, ch, n0);
      if (it >= n_pre) load_b(it);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = tc::make_idesc(kTileM, BLOCK_N, false, B_MN);
    for (int it = 0; it < k_iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      tc::mbar_wait(&full[s], ph);
      if (it == 0) HZ_STAMP(3);                          // first operand stage landed
      tc::fence_after_sync();
      const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
      const uint32_t sb = sa + kABytes;
#pragma unroll
      for (int k = 0; k < kKBlock / 16; ++k) {
        const uint64_t da = tc::make_sdesc(sa + k * 32, 16, 1024);
        const uint64_t db = B_MN ? tc::make_sdesc(sb + k * 2048, 8192, 1024)
                                 : tc::make_sdesc(sb + k * 32, 16, 1024);
        tc::umma_f16(tmem_d, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
      }
      tc::umma_commit(&empty[s]);
    }
    if (k_iters > 0) tc::umma_commit(tmem_full);
    HZ_STAMP(4);                                         // last MMA issued
  }
  __syncwarp();

  // ===================== epilogue (all 4 warps) =====================
  __nv_bfloat16* staging = reinterpret_cast<__nv_bfloat16*>(smem);
  const int row = warp * 32 + lane;
  int row_lo = 0, row_hi = kTileM;          // rows of the tile this CTA writes out
  if (p.splits > 1 && p.cluster) {
    row_lo = split * (kTileM / p.splits);
    row_hi = row_lo + kTileM / p.splits;
  }
  constexpr int kVecPerRow = BLOCK_N / 8;
  constexpr int kRowsPerPass = 128 / kVecPerRow;
  constexpr int kPasses = kTileM / kRowsPerPass;
  // global element offset of (tile row r0, this thread's 8-channel vector); -1 for rows past the last image
  auto out_offset = [&](int r0) -> long long {
    const int wi = r0 % p.BW;
    const int hi = (r0 / p.BW) % p.BH;
    const int n = n0 + r0 / (p.BW * p.BH);
    // rows past the last image, and columns past the last channel (channel counts that are not a multiple of 64:
    // the TMA boxes were zero-filled / over-read there, the results are simply not stored)
    if (n >= p.n_images || nt * BLOCK_N + (int)(threadIdx.x % kVecPerRow) * 8 >= p.ncols) return -1;
    return (long long)n * p.out_n_stride + (long long)(h0 + hi) * p.out_h_stride + (long long)wi * p.out_w_stride +
           p.cls_out_off[cls] + nt * BLOCK_N + (threadIdx.x % kVecPerRow) * 8;
  };
  // residual-gradient fusion: the addend rows this thread will write are requested now, so their L2 latency
  // hides under the MMA tail instead of serialising with the stores
  // all of this thread's output addresses, computed while the MMAs are still running (the runtime divisions in
  // out_offset() used to sit between the accumulator and the stores)
  long long offs[kPasses];
#pragma unroll
  for (int i = 0; i < kPasses; ++i) {
    const int r0 = row_lo + threadIdx.x / kVecPerRow + i * kRowsPerPass;
    offs[i] = r0 < row_hi ? out_offset(r0) : -1;
  }
  bf16x8 addv[kPasses];
  const __nv_bfloat16* pre = p.addend != nullptr ? p.addend : p.bn_residual;      // (never both: dgrad vs forward)
  if (pre != nullptr) {
#pragma unroll
    for (int i = 0; i < kPasses; ++i)
      if (offs[i] >= 0) addv[i] = ld8(pre + offs[i]);
  }
  if (k_iters > 0) {
    tc::mbar_wait(tmem_full, 0);
    tc::fence_after_sync();
  }
  if (threadIdx.x == 64) HZ_STAMP(5);                    // accumulator complete
  if (p.splits > 1 && p.cluster) {
    // ---- cluster split-K: every CTA of the cluster parks its fp32 partial tile in its own shared memory,
    //      then CTA r reduces rows [r*128/S, (r+1)*128/S) of all S partials through distributed shared memory
    constexpr int kRedLd = BLOCK_N + 4;                                  // fp32 words per row (conflict-free)
    float* red = reinterpret_cast<float*>(smem);                          // [128][kRedLd]  (pipeline buffers are idle)
    staging = reinterpret_cast<__nv_bfloat16*>(smem + 40960);            // keep the bf16 staging clear of `red`
#pragma unroll
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      if (k_iters > 0) {
        tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(red + row * kRedLd + c0 + j) =
            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                        __uint_as_float(r[j + 3]));
    }
    cluster_sync();
    const int Sx = p.splits;
    const int rows_per = kTileM / Sx;
    constexpr int kChunks = BLOCK_N / 4;
    const uint32_t red_base = smem_u32(red);
    if (Sx == 4) {
      cluster_reduce_rows<4, BLOCK_N, kRedLd, S::kStagingLd>(red_base, row_lo, staging);
    } else if (Sx == 8) {
      cluster_reduce_rows<8, BLOCK_N, kRedLd, S::kStagingLd>(red_base, row_lo, staging);
    } else if (Sx == 2) {
      cluster_reduce_rows<2, BLOCK_N, kRedLd, S::kStagingLd>(red_base, row_lo, staging);
    } else
    for (int idx = threadIdx.x; idx < rows_per * kChunks; idx += 128) {
      const int rr = row_lo + idx / kChunks, ch = idx % kChunks;
      const uint32_t off = (uint32_t)((rr * kRedLd + ch * 4) * 4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s2 = 0; s2 < Sx; ++s2) {                                   // fixed order: deterministic
        const float4 v = ld_dsmem_f4(map_to_cta(red_base + off, (uint32_t)s2));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      __nv_bfloat162 lo2 = __floats2bfloat162_rn(acc.x, acc.y), hi2 = __floats2bfloat162_rn(acc.z, acc.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo2);
      pk.y = *reinterpret_cast<uint32_t*>(&hi2);
      *reinterpret_cast<uint2*>(staging + rr * S::kStagingLd + ch * 4) = pk;
    }
    // Peers may still be reading my `red` (it is not written again), so the only hazard left is this CTA
    // exiting early: arrive now, wait right before the kernel ends — the stores below overlap the barrier.
    cluster_arrive();
  } else if (p.splits == 1) {
    if (k_iters > 0) {
#pragma unroll
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
        tc::tmem_ld_wait();
        __nv_bfloat16* dst = staging + row * S::kStagingLd + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[j + i]);
          st8(dst + j, pack8(f));
        }
      }
    } else {
      // a class with no taps (1x1 stride-2 dgrad, odd parities): the gradient is exactly zero
      float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c0 = 0; c0 < BLOCK_N; c0 += 8) st8(staging + row * S::kStagingLd + c0, pack8(z));
    }
  } else {
    // ---- split-K: accumulate the partial tile into the fp32 workspace (vector red), the last-arriving
    //      CTA of the tile turns the sum into the output tile and leaves workspace + counter zeroed
    const int tile = (cls * gridDim.y + nt) * gridDim.x + mt;
    float* wrow = p.ws + ((size_t)tile * kTileM + row) * BLOCK_N;
    if (k_iters > 0) {
#pragma unroll
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wrow + c0 + j),
                       "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])),
                       "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                       : "memory");
      }
    }
    __threadfence();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
      const unsigned old = atomicAdd(&p.sem[tile], 1u);
      const int last = (old == (unsigned)(p.splits - 1));
      if (last) { p.sem[tile] = 0u; __threadfence(); }
      s_last = last;
    }
    __syncthreads();
    if (!s_last) {                       // uniform per CTA
      tc::fence_before_sync();
      __syncthreads();
      if (warp == 2) tc::tmem_dealloc(tmem_d, BLOCK_N);
      return;
    }
#pragma unroll
    for (int c0 = 0; c0 < BLOCK_N; c0 += 8) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(wrow + c0));
      const float4 b = __ldcg(reinterpret_cast<const float4*>(wrow + c0 + 4));
      __stcg(reinterpret_cast<float4*>(wrow + c0), make_float4(0.f, 0.f, 0.f, 0.f));
      __stcg(reinterpret_cast<float4*>(wrow + c0 + 4), make_float4(0.f, 0.f, 0.f, 0.f));
      const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      st8(staging + row * S::kStagingLd + c0, pack8(f));
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 64) HZ_STAMP(6);                    // tile (reduced over the cluster) staged in shared memory

  // coalesced stores: BLOCK_N/8 16-byte vectors per row.  The BatchNorm sums (forward) come from the same
  // registers: a separate per-column pass over the staged tile was a chain of dependent shared-memory loads
  // (~1.1 us for a 128-row tile, measured with the phase stamps).
  static_assert(BLOCK_N == 64, "stats reduction below assumes 8 vectors per row");
  float ssum[8], ssq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ssum[j] = ssq[j] = 0.f;
  {
    const int vec = threadIdx.x % kVecPerRow;
    float bmu[8], bis[8];
    if (kBnBwd) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = nt * BLOCK_N + vec * 8 + j;
        bmu[j] = c < p.ncols ? p.bnb_mean[c] : 0.f;
        bis[j] = c < p.ncols ? p.bnb_invstd[c] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < kPasses; ++i) {
      const int r0 = row_lo + threadIdx.x / kVecPerRow + i * kRowsPerPass;
      const long long off = offs[i];
      if (off < 0) continue;              // rows past the last image are zero-filled: they add nothing to the sums
      bf16x8 v = ld8(staging + r0 * S::kStagingLd + vec * 8);
      if (!kBnBwd && p.stats != nullptr) {
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { ssum[j] += f[j]; ssq[j] += f[j] * f[j]; }
      }
      if (p.addend != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v.v[j] = __hadd2(v.v[j], addv[i].v[j]);  // bf16 + bf16 -> bf16, as the separate add did
      }
      if (kBnBwd) {
        // BatchNorm-backward sums of the layer that produced this conv's input, from the bf16 values being stored
        // (exactly what channel_reduce_kernel<true> would read back): g = dx * [out > 0], xhat = (y - mean) * invstd
        float g[8], y[8];
        unpack8(v, g);
        unpack8(ld8(p.bnb_yraw + off), y);
        if (p.bnb_out != nullptr) {
          float o[8];
          unpack8(ld8(p.bnb_out + off), o);
          if (p.bnb_cap6) {
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = (o[j] > 0.f && o[j] < 6.f) ? g[j] : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { ssum[j] += g[j]; ssq[j] += g[j] * (y[j] - bmu[j]) * bis[j]; }
      }
      st8(p.out + off, v);
    }
  }
  if (p.stats != nullptr) {
    // lanes l, l^8, l^16, l^24 hold the same 8 columns (different rows): fold them, then the 4 warps through smem
    __shared__ float stat_sm[4][2][BLOCK_N];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 8);
      ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], 8);
      ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 16);
      ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], 16);
    }
    if (lane < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        stat_sm[warp][0][lane * 8 + j] = ssum[j];
        stat_sm[warp][1][lane * 8 + j] = ssq[j];
      }
    }
    __syncthreads();
    const int col = threadIdx.x % BLOCK_N, which = threadIdx.x / BLOCK_N;       // 128 threads = 64 columns x {Σ, Σ²}
    const float tot = stat_sm[0][which][col] + stat_sm[1][which][col] + stat_sm[2][which][col] + stat_sm[3][which][col];
    if (nt * BLOCK_N + col < p.ncols) atomicAdd(&p.stats[which * p.ncols + nt * BLOCK_N + col], tot);
  }
  if (threadIdx.x == 64) HZ_STAMP(7);                    // output rows + BN sums written
  if (p.bn_out != nullptr) {
    // ---- fused BN + residual + ReLU.  grid <= #SMs (host-checked) and 1 CTA/SM: all CTAs are co-resident, the
    //      barrier cannot deadlock; PDL dependents are only scheduled once every CTA of this grid has started.
    __shared__ float bn_scale[BLOCK_N], bn_shift[BLOCK_N];
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(p.bn_counter, 1u);
      const unsigned total = gridDim.x * gridDim.y * gridDim.z;
      const long long t0 = clock64();
      while (ld_acquire_gpu_u32(p.bn_counter) < total) {
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
    __syncthreads();
    if (threadIdx.x < BLOCK_N) {
      const int c = nt * BLOCK_N + threadIdx.x;
      const float mu = __ldcg(&p.stats[c]) * p.bn_inv_count;
      const float var = fmaxf(__ldcg(&p.stats[p.ncols + c]) * p.bn_inv_count - mu * mu, 0.f);
      const float is = rsqrtf(var + p.bn_eps);
      const float g = p.bn_gamma[c];
      bn_scale[threadIdx.x] = g * is;
      bn_shift[threadIdx.x] = p.bn_beta[c] - mu * g * is;
      if (mt == 0 && blockIdx.z == 0) {            // one CTA per channel block publishes the statistics
        p.bn_mean[c] = mu;
        p.bn_invstd[c] = is;
        if (p.bn_rmean != nullptr) {
          p.bn_rmean[c] = (1.f - p.bn_momentum) * p.bn_rmean[c] + p.bn_momentum * mu;
          p.bn_rvar[c] = (1.f - p.bn_momentum) * p.bn_rvar[c] + p.bn_momentum * var * p.bn_unbias;
        }
      }
    }
    __syncthreads();
    const int vec = threadIdx.x % kVecPerRow;
#pragma unroll
    for (int i = 0; i < kPasses; ++i) {
      const int r0 = row_lo + threadIdx.x / kVecPerRow + i * kRowsPerPass;
      const long long off = offs[i];
      if (off < 0) continue;
      float f[8];
      unpack8(ld8(staging + r0 * S::kStagingLd + vec * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = f[j] * bn_scale[vec * 8 + j] + bn_shift[vec * 8 + j];
      if (p.bn_residual != nullptr) {
        float r[8];
        unpack8(addv[i], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += r[j];
      }
      if (p.bn_relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      st8(p.bn_out + off, pack8(f));
    }
  }
  if (p.splits > 1 && p.cluster) cluster_wait();         // no peer reads this CTA's shared memory any more
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_d, BLOCK_N);
  if (threadIdx.x == 64) HZ_STAMP(8);                    // done
#undef HZ_STAMP
}

// ------------------------------------------------------------------------------------------------
// Throughput variant of igemm_kernel for grids of many waves (large --batch_size): PERSISTENT CTAs, one per SM, that
// walk a static tile schedule with the three pipelines of the canonical sm_100 GEMM —
//     warp 0 (one lane)  TMA producer: runs ahead across tile boundaries (the operand ring never drains)
//     warp 1 (one lane)  tcgen05.mma issuer into TWO TMEM accumulators (2 x BLOCK_N columns), alternating per tile
//     warps 2..5         epilogue: tcgen05.ld -> bf16 -> dedicated staging tile -> coalesced stores (+ BN sums,
//                        + residual-gradient addend); the accumulator is handed back (tmem_empty) as soon as it is in
//                        registers, so the MMAs of tile i+1 run under the stores of tile i
// One-tile-per-CTA pays barrier init + TMEM alloc + descriptor prefetch + pipeline fill + a serial epilogue per tile
// (measured: MMA ~0.6 us of a ~4.5 us CTA at batch 4096, 14-16 % of the bf16 peak); here they are paid once per SM or
// overlapped.  Same operands, taps, parity classes, masking and epilogue arithmetic as igemm_kernel (no split-K: there
// are more tiles than SMs).  Dispatch: hz_conv_fwd / hz_conv_dgrad via use_persistent() — opt-in (HZ_CONV_PERSIST=1 |
// auto) until its first hardware run has passed.
// ------------------------------------------------------------------------------------------------
constexpr int kPersistThreads = 192;
template <int BLOCK_N>
struct PersistSmem {
  static constexpr int kNumStages = BLOCK_N == 64 ? 6 : 5;             // 24 KB / 32 KB operand stages
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kPipeBytes = kNumStages * kStageBytes;
  static constexpr int kStagingLd = BLOCK_N + 8;
  static constexpr int kStagingOff = kPipeBytes;                       // dedicated: the ring is busy with the next tile
  static constexpr int kStagingBytes = kTileM * kStagingLd * 2;
  static constexpr int kBarOff = kStagingOff + ((kStagingBytes + 1023) / 1024) * 1024;
  static constexpr int kTotal = kBarOff + 256 + 1024;                  // + barriers + alignment slack
};

HZ_DEVINL void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps only

// BLOCK_N = 64 | 128 output columns per tile (128: the activation tile is fetched once per 128 output channels instead
// of once per 64 — layers with >= 128 output channels); the epilogue works in 64-column halves either way.
template <int BLOCK_N, bool B_MN>
__global__ void __launch_bounds__(kPersistThreads, 1) igemm_persist_kernel(const __grid_constant__ AMaps amaps,
                                                                           const __grid_constant__ CUtensorMap bmap,
                                                                           const __grid_constant__ IgemmParams p,
                                                                           const int m_tiles, const int n_tiles) {
  using S = PersistSmem<BLOCK_N>;
  static_assert(BLOCK_N == 64 || BLOCK_N == 128, "epilogue below works in 64-column halves");
  constexpr int NS = S::kNumStages;
  constexpr int kHalves = BLOCK_N / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty = full + NS;
  uint64_t* tmem_full = empty + NS;               // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  __shared__ float stat_sm[4][2][64];

  pdl_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.num_classes * n_tiles * m_tiles;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) tc::prefetch_tmap(&amaps.m[i]);
    tc::prefetch_tmap(&bmap);
    for (int s = 0; s < NS; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 2 * BLOCK_N);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;
  pdl_wait();

  // tile t -> (class, n tile, m tile), m fastest: neighbouring CTAs share the weight tile and adjacent input rows
  auto decode = [&](int t, int& cls, int& nt, int& mt) {
    mt = t % m_tiles;
    const int r = t / m_tiles;
    nt = r % n_tiles;
    cls = r / n_tiles;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int it = 0;                                                       // ring position, continues across tiles
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int cls, nt, mt;
        decode(t, cls, nt, mt);
        const TapList& taps = p.cls[cls];
        const int n0 = (p.BN == 1) ? mt / p.tiles_per_img : mt * p.BN;
        const int h0 = (p.BN == 1) ? (mt % p.tiles_per_img) * p.BH : 0;
        const int k_total = taps.n * p.cblocks;
        for (int k = 0; k < k_total; ++k, ++it) {
          const int tp = k / p.cblocks, cb = k % p.cblocks;
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          uint8_t* sa = smem + s * S::kStageBytes;
          uint8_t* sb = sa + kABytes;
          tc::mbar_wait(&empty[s], ph ^ 1);
          tc::mbar_arrive_expect_tx(&full[s], S::kStageBytes);
          tc::tma_load_4d(sa, &amaps.m[taps.map[tp]], &full[s], cb * kKBlock, taps.dw[tp], h0 + taps.dh[tp], n0);
          if (!B_MN) {
            tc::tma_load_2d(sb, &bmap, &full[s], taps.bk[tp] + cb * kKBlock, nt * BLOCK_N);      // box: 64 k x BLOCK_N rows
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)                                               // 64-column MN atoms
              tc::tma_load_2d(sb + j * 8192, &bmap, &full[s], taps.bk[tp] + nt * BLOCK_N + j * 64, cb * kKBlock);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = tc::make_idesc(kTileM, BLOCK_N, false, B_MN);
      int it = 0, acc_it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int cls, nt, mt;
        decode(t, cls, nt, mt);
        const int k_total = p.cls[cls].n * p.cblocks;
        if (k_total == 0) continue;                                     // a class without taps: no accumulator used
        const int acc = acc_it & 1;
        const uint32_t acc_ph = (acc_it >> 1) & 1;
        ++acc_it;
        tc::mbar_wait(&tmem_empty[acc], acc_ph ^ 1);                    // the epilogue has drained this accumulator
        tc::fence_after_sync();
        const uint32_t td = tmem_d + (uint32_t)(acc * BLOCK_N);
        for (int k = 0; k < k_total; ++k, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          tc::mbar_wait(&full[s], ph);
          tc::fence_after_sync();
          const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int kk = 0; kk < kKBlock / 16; ++kk) {
            const uint64_t da = tc::make_sdesc(sa + kk * 32, 16, 1024);
            const uint64_t db = B_MN ? tc::make_sdesc(sb + kk * 2048, 8192, 1024)
                                     : tc::make_sdesc(sb + kk * 32, 16, 1024);
            tc::umma_f16(td, da, db, idesc, (k > 0 || kk > 0) ? 1u : 0u);
          }
          tc::umma_commit(&empty[s]);
        }
        tc::umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ===================== epilogue: warps 2..5 =====================
    __nv_bfloat16* staging = reinterpret_cast<__nv_bfloat16*>(smem + S::kStagingOff);
    const int et = threadIdx.x - 64;                       // 0..127
    const int ew = warp - 2;                               // statistics slot of this warp
    const int lq = warp & 3;                               // TMEM lane quarter this warp may read (warp id mod 4)
    const int row = lq * 32 + lane;                        // accumulator row held by this thread
    constexpr int kVecPerRow = 8;                          // per 64-column half
    constexpr int kRowsPerPass = 128 / kVecPerRow;
    constexpr int kPasses = kTileM / kRowsPerPass;
    const int vec = et % kVecPerRow;
    int acc_it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int cls, nt, mt;
      decode(t, cls, nt, mt);
      const int n0 = (p.BN == 1) ? mt / p.tiles_per_img : mt * p.BN;
      const int h0 = (p.BN == 1) ? (mt % p.tiles_per_img) * p.BH : 0;
      const int k_total = p.cls[cls].n * p.cblocks;
      // element offset of (row, column 0) for this thread's 8 rows; -1: row past the last image
      long long rowoff[kPasses];
#pragma unroll
      for (int i = 0; i < kPasses; ++i) {
        const int r0 = et / kVecPerRow + i * kRowsPerPass;
        const int wi = r0 % p.BW;
        const int hi = (r0 / p.BW) % p.BH;
        const int n = n0 + r0 / (p.BW * p.BH);
        rowoff[i] = n >= p.n_images ? -1
                                    : (long long)n * p.out_n_stride + (long long)(h0 + hi) * p.out_h_stride +
                                          (long long)wi * p.out_w_stride + p.cls_out_off[cls];
      }
      if (k_total > 0) {
        const int acc = acc_it & 1;
        const uint32_t acc_ph = (acc_it >> 1) & 1;
        ++acc_it;
        tc::mbar_wait(&tmem_full[acc], acc_ph);
        tc::fence_after_sync();
#pragma unroll
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          uint32_t r[32];
          tc::tmem_ld32(tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), r);
          tc::tmem_ld_wait();
          __nv_bfloat16* dst = staging + row * S::kStagingLd + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[j + i]);
            st8(dst + j, pack8(f));
          }
        }
        // the accumulator is in shared memory: give it back before the (long) global stores
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
      } else {
        float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c0 = 0; c0 < BLOCK_N; c0 += 8) st8(staging + row * S::kStagingLd + c0, pack8(z));
      }
      epi_bar_sync();                                      // the whole tile is staged
#pragma unroll 1
      for (int half = 0; half < kHalves; ++half) {
        const int cbase = nt * BLOCK_N + half * 64;        // first output column of this half
        const bool col_ok = cbase + vec * 8 < p.ncols;     // columns past the last channel are not stored
        bf16x8 addv[kPasses];
        if (p.addend != nullptr) {
#pragma unroll
          for (int i = 0; i < kPasses; ++i)
            if (rowoff[i] >= 0 && col_ok) addv[i] = ld8(p.addend + rowoff[i] + cbase + vec * 8);
        }
        float ssum[8], ssq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ssum[j] = ssq[j] = 0.f;
#pragma unroll
        for (int i = 0; i < kPasses; ++i) {
          const int r0 = et / kVecPerRow + i * kRowsPerPass;
          if (rowoff[i] < 0 || !col_ok) continue;          // zero-filled rows add nothing to the sums
          bf16x8 v = ld8(staging + r0 * S::kStagingLd + half * 64 + vec * 8);
          if (p.stats != nullptr) {
            float f[8];
            unpack8(v, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { ssum[j] += f[j]; ssq[j] += f[j] * f[j]; }
          }
          if (p.addend != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v.v[j] = __hadd2(v.v[j], addv[i].v[j]);
          }
          st8(p.out + rowoff[i] + cbase + vec * 8, v);
        }
        if (p.stats != nullptr) {
          // lanes l, l^8, l^16, l^24 hold the same 8 columns (different rows): fold them, then the 4 warps through smem
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 8);
            ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], 8);
            ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 16);
            ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], 16);
          }
          if (lane < 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              stat_sm[ew][0][lane * 8 + j] = ssum[j];
              stat_sm[ew][1][lane * 8 + j] = ssq[j];
            }
          }
          epi_bar_sync();
          const int col = et % 64, which = et / 64;        // 128 threads = 64 columns x {sum, sum of squares}
          const float tot = stat_sm[0][which][col] + stat_sm[1][which][col] + stat_sm[2][which][col] + stat_sm[3][which][col];
          if (cbase + col < p.ncols) atomicAdd(&p.stats[which * p.ncols + cbase + col], tot);
          epi_bar_sync();                                  // stat_sm is rewritten by the next half / tile
        }
      }
      epi_bar_sync();                                      // staging may be overwritten by the next tile
    }
  }
  __syncwarp();                                            // re-converge the single-lane role warps
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_d, 2 * BLOCK_N);
}

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
struct WgradParams {
  TapList taps;                  // bk[] = output tap index
  int KBN, KBH, kb_per_img;      // 64-pixel k-block box over dY
  int kblocks, splits;
  int n_tiles;                   // Cin / BLOCK_N
  int Cout;
  long long ld_out, tap_stride;  // dW row stride / per-tap column offset (elements)
  int n_valid;                   // valid columns per tap (Cin, or 147 for the padded stem)
  int mode;                      // 0 = store, 1 = read-add-store, 2 = atomic add (split-K)
  float* out;
};

template <int BLOCK_N>
struct WgradSmem {
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kTotal = kBarOff + 128 + 1024;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(128) wgrad_kernel(const __grid_constant__ CUtensorMap dymap,
                                                    const __grid_constant__ AMaps xmaps,
                                                    const __grid_constant__ WgradParams p) {
  using S = WgradSmem<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  pdl_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x / p.n_tiles, nt = blockIdx.x % p.n_tiles;
  const int t = blockIdx.y;
  const int per = (p.kblocks + p.splits - 1) / p.splits;
  const int kb_lo = blockIdx.z * per;
  const int kb_hi = min(kb_lo + per, p.kblocks);
  const int k_iters = max(kb_hi - kb_lo, 0);

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&dymap);
    for (int i = 0; i < 4; ++i) tc::prefetch_tmap(&xmaps.m[i]);
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, BLOCK_N);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;
  pdl_wait();

  if (warp == 0 && lane == 0) {
    const CUtensorMap* xm = &xmaps.m[p.taps.map[t]];
    for (int it = 0; it < k_iters; ++it) {
      const int kb = kb_lo + it;
      const int n0 = (p.KBN == 1) ? kb / p.kb_per_img : kb * p.KBN;
      const int h0 = (p.KBN == 1) ? (kb % p.kb_per_img) * p.KBH : 0;
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      tc::mbar_wait(&empty[s], ph ^ 1);
      uint8_t* sa = smem + s * S::kStageBytes;
      uint8_t* sb = sa + kABytes;
      tc::mbar_arrive_expect_tx(&full[s], S::kStageBytes);
      // A: dY^T, two 64-channel atoms (the second is zero-filled by TMA when Cout == 64)
      tc::tma_load_4d(sa, &dymap, &full[s], mt * 128, 0, h0, n0);
      tc::tma_load_4d(sa + 8192, &dymap, &full[s], mt * 128 + 64, 0, h0, n0);
#pragma unroll
      for (int j = 0; j < BLOCK_N / 64; ++j)
        tc::tma_load_4d(sb + j * 8192, xm, &full[s], nt * BLOCK_N + j * 64, p.taps.dw[t], h0 + p.taps.dh[t], n0);
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = tc::make_idesc(kTileM, BLOCK_N, true, true);
    for (int it = 0; it < k_iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      tc::mbar_wait(&full[s], ph);
      tc::fence_after_sync();
      const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
      const uint32_t sb = sa + kABytes;
#pragma unroll
      for (int k = 0; k < kKBlock / 16; ++k) {
        const uint64_t da = tc::make_sdesc(sa + k * 2048, 8192, 1024);
        const uint64_t db = tc::make_sdesc(sb + k * 2048, 8192, 1024);
        tc::umma_f16(tmem_d, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
      }
      tc::umma_commit(&empty[s]);
    }
    tc::umma_commit(tmem_full);
  }
  __syncwarp();

  // epilogue: TMEM -> fp32 tile in smem (pipeline buffers are idle now) -> coalesced row-wise float4 writes
  constexpr int kLd = BLOCK_N + 4;                       // fp32 words per staged row (conflict-free 16B rows)
  float* tile = reinterpret_cast<float*>(smem);
  if (k_iters > 0) {
    tc::mbar_wait(tmem_full, 0);
    tc::fence_after_sync();
    const int row = warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(tile + row * kLd + c0 + j) =
            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                        __uint_as_float(r[j + 3]));
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (k_iters > 0) {
    constexpr int kChunks = BLOCK_N / 4;                 // float4 chunks per row
    const bool vec_ok = ((p.ld_out | p.tap_stride) & 3) == 0 && (p.n_valid & 3) == 0;
    for (int idx = threadIdx.x; idx < kTileM * kChunks; idx += 128) {
      const int r = idx / kChunks, ch = idx % kChunks;
      const int co = mt * 128 + r;
      if (co >= p.Cout) continue;
      const int c = nt * BLOCK_N + ch * 4;
      float* dst = p.out + (long long)co * p.ld_out + (long long)p.taps.bk[t] * p.tap_stride + c;
      const float4 v = *reinterpret_cast<const float4*>(tile + r * kLd + ch * 4);
      if (vec_ok && c + 3 < p.n_valid) {
        if (p.mode == 2) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z),
                       "f"(v.w) : "memory");
        } else if (p.mode == 1) {
          float4 o = *reinterpret_cast<float4*>(dst);
          o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
          *reinterpret_cast<float4*>(dst) = o;
        } else {
          *reinterpret_cast<float4*>(dst) = v;
        }
      } else {
        const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (c + j < p.n_valid) {
            if (p.mode == 2) atomicAdd(dst + j, f[j]);
            else dst[j] = (p.mode == 1 ? dst[j] : 0.f) + f[j];
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_d, BLOCK_N);
}

}  // namespace hz

// ================================================================================================
// host: tensor maps + launch
// ================================================================================================
namespace {
using namespace hz::host;

// persistent split-K workspace (per device): self-cleaning, so it is zeroed exactly once
constexpr size_t kWsTiles = 512;
struct SplitWs { float* ws = nullptr; unsigned* sem = nullptr; };
SplitWs get_split_ws() {
  static SplitWs per_dev[16];
  int dev = 0;
  cudaGetDevice(&dev);
  SplitWs& w = per_dev[dev & 15];
  if (w.ws == nullptr) {
    const size_t bytes = kWsTiles * 128 * 64 * sizeof(float);
    if (cudaMalloc(&w.ws, bytes + kWsTiles * sizeof(unsigned)) != cudaSuccess) { w.ws = nullptr; return w; }
    cudaMemset(w.ws, 0, bytes + kWsTiles * sizeof(unsigned));
    w.sem = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(w.ws) + bytes);
  }
  return w;
}
// cluster (DSMEM) split-K factor: power of two <= 8; only when the K loop is long enough to pay for the
// two cluster barriers (~1.5 us) and the grid still fits in one wave
long long* g_conv_dbg = nullptr;

int hz_num_sms() {
  static const int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

// How many clusters of `cluster` CTAs of this kernel can be resident at once (1 CTA/SM, clusters live inside a GPC:
// on B200 a cluster of 8 does not tile the 148 SMs evenly).  A grid with more clusters than that runs in two
// waves — and a kernel with a device-wide barrier would deadlock — so the split factor is chosen to fit.
template <typename K>
int max_active_clusters(K kernel, size_t smem, int cluster) {
  static int cache[17] = {0};
  if (cluster < 1 || cluster > 16) return 0;
  if (cache[cluster] == 0) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1, 1, (unsigned)cluster);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)cluster;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cluster == 1) {
      n = hz_num_sms();
    } else if (cudaOccupancyMaxActiveClusters(&n, (const void*)kernel, &cfg) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      n = hz_num_sms() / cluster - 2;          // conservative guess
    }
    cache[cluster] = n > 0 ? n : 1;
  }
  return cache[cluster];
}

int pick_cluster_splits(int tiles, int k_total, int (*max_clusters)(int)) {
  // validated on B200 (all conv numerics tests; 0.670 -> 0.612 ms/step); HZ_CLUSTER_SPLITK=0 disables
  static const bool off = [] { const char* e = getenv("HZ_CLUSTER_SPLITK"); return e && e[0] == '0'; }();
  static const int min_k = [] { const char* e = getenv("HZ_CLUSTER_MIN_K"); return e ? atoi(e) : 8; }();
  static const int min_per = [] { const char* e = getenv("HZ_CLUSTER_MIN_PER"); return e ? atoi(e) : 2; }();
  if (off || tiles <= 0 || k_total < min_k) return 1;
  static const bool one_wave = [] { const char* e = getenv("HZ_CLUSTER_ONE_WAVE"); return !(e && e[0] == '0'); }();
  int s = 8;
  while (s > 1 && (tiles * s > hz_num_sms() || k_total / s < min_per || (one_wave && tiles > max_clusters(s)))) s >>= 1;
  return s;
}

// Persistent (throughput) kernel selection: 0 never (DEFAULT: the kernel was written after the round's GPU budget was
// spent and has not run on hardware yet — tests/test_gpu_persist.py is its first execution), 1 always, -1 auto (grids
// of >= HZ_CONV_PERSIST_WAVES x #SMs tiles).  HZ_CONV_PERSIST = 0 | 1 | auto sets the initial mode,
// hz_conv_set_persist() changes it at run time (tests, tools/conv_roofline.py).
int g_persist_mode = [] { const char* e = getenv("HZ_CONV_PERSIST"); return e ? (e[0] == '1' ? 1 : (e[0] == 'a' ? -1 : 0)) : 0; }();
bool use_persistent(int total_tiles) {
  static const int waves = [] { const char* e = getenv("HZ_CONV_PERSIST_WAVES"); const int v = e ? atoi(e) : 4; return v > 0 ? v : 4; }();
  if (g_persist_mode == 0 || total_tiles <= 0) return false;
  if (g_persist_mode >= 1) return true;
  return total_tiles >= waves * hz_num_sms();
}
template <int BLOCK_N, bool B_MN>
int launch_persistent_n(const hz::AMaps& am, const CUtensorMap& bm, hz::IgemmParams& p, int m_tiles, cudaStream_t st) {
  using SM = hz::PersistSmem<BLOCK_N>;
  static bool attr = set_smem(hz::igemm_persist_kernel<BLOCK_N, B_MN>, SM::kTotal);
  (void)attr;
  p.splits = 1; p.cluster = 0; p.prefetch_b = 0; p.dbg = nullptr; p.ws = nullptr; p.sem = nullptr;
  const int n_tiles = (p.ncols + BLOCK_N - 1) / BLOCK_N;
  const int total = p.num_classes * n_tiles * m_tiles;
  const int grid = total < hz_num_sms() ? total : hz_num_sms();
  return hz::launch(hz::igemm_persist_kernel<BLOCK_N, B_MN>, dim3(grid), dim3(hz::kPersistThreads), SM::kTotal, st, am, bm,
                    p, m_tiles, n_tiles) == cudaSuccess ? 0 : -1;
}
// 128-column tiles when the output has >= 128 channels in whole 128-column tiles (mode 2 / HZ_CONV_PERSIST_N=64 pin 64)
bool persist_wide(int ncols) {
  static const int pin = [] { const char* e = getenv("HZ_CONV_PERSIST_N"); return e ? atoi(e) : 0; }();
  if (pin == 64 || g_persist_mode == 2) return false;
  return ncols >= 128 && ncols % 128 == 0;
}

int prefetch_weights_enabled() {
  static const int on = [] { const char* e = getenv("HZ_PREFETCH_B"); return (e && e[0] == '0') ? 0 : 1; }();
  return on;
}

int pick_splits(int tiles, int k_total) {
  // Measured on B200: accumulating 32 KB fp32 tiles with red.global.add.v4.f32 costs more than the
  // K-loop time it saves (layer1 conv 7.7 -> 26 us), so workspace split-K stays opt-in (HZ_SPLITK=1).
  static const bool enabled = [] { const char* e = getenv("HZ_SPLITK"); return e && e[0] == '1'; }();
  if (!enabled) return 1;
  if (tiles <= 0 || k_total <= 1 || (size_t)tiles > kWsTiles) return 1;
  int s = 132 / tiles;
  if (s > k_total) s = k_total;
  if (s < 1) s = 1;
  // never leave a split with nothing to do
  const int per = (k_total + s - 1) / s;
  s = (k_total + per - 1) / per;
  return s;
}

}  // namespace

extern "C" {

// the shape rules of hz_conv_supported alone (host-only: also answers on a machine without a CUDA driver)
int hz_conv_shape_ok(int N, int H, int W, int Cin, int Cout, int R, int stride, int pad) {
  // channel counts only need 16-byte rows (TMA): a 64-wide box over a narrower tensor is zero-filled, surplus output
  // columns are masked in the epilogue — tensor-parallel shards (256/8 = 32 channels) stay on these kernels
  if ((Cin & 7) || (Cout & 7) || Cin < 8 || Cout < 8) return 0;
  if (!((R == 3 && pad == 1) || (R == 1 && pad == 0))) return 0;
  if (stride != 1 && stride != 2) return 0;
  if (stride == 2 && ((H | W) & 1)) return 0;
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - R) / stride + 1;
  Tile t;
  if (!pick_tile(128, N, Ho, Wo, &t)) return 0;
  if (!pick_tile(64, N, Ho, Wo, &t)) return 0;
  if (!pick_tile(128, N, H, W, &t) && stride == 1) return 0;
  return 1;
}

int hz_conv_supported(int N, int H, int W, int Cin, int Cout, int R, int stride, int pad) {
  return hz_conv_shape_ok(N, H, W, Cin, Cout, R, stride, pad) && get_encode() != nullptr;
}

// per-CTA phase stamps (16 x int64 per CTA) for the next forward / dgrad launches; nullptr switches them off
void hz_conv_set_debug(long long* buf) { g_conv_dbg = buf; }

// persistent-kernel selection for the following forward / dgrad launches: -1 auto, 0 never, 1 always (128-column tiles
// where the channel count allows), 2 always with 64-column tiles only; returns the old mode
int hz_conv_set_persist(int mode) {
  const int old = g_persist_mode;
  g_persist_mode = mode < 0 ? -1 : (mode > 2 ? 1 : mode);
  return old;
}

// resident clusters of the forward conv kernel for cluster sizes 1,2,4,8 (diagnostics)
void hz_cluster_capacity(int out[4]) {
  static bool attr = set_smem(hz::igemm_kernel<64, false>, hz::IgemmSmem<64>::kTotal);
  (void)attr;
  const int cs[4] = {1, 2, 4, 8};
  for (int i = 0; i < 4; ++i) out[i] = max_active_clusters(hz::igemm_kernel<64, false>, hz::IgemmSmem<64>::kTotal, cs[i]);
}

// y[N,Ho,Wo,Cout] = conv(x[N,H,W,Cin], w[Cout,R,S,Cin]); stats (2*Cout fp32, zeroed here) optional
int hz_conv_fwd(const void* x, const void* w, void* y, float* stats, int stats_is_zero, int N, int H, int W,
                int Cin, int Cout, int R, int stride, int pad, int weights_stable, const HzBnFuse* bn,
                cudaStream_t st) {
  const int S_ = R;
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S_) / stride + 1;
  static bool attr = set_smem(hz::igemm_kernel<64, false>, hz::IgemmSmem<64>::kTotal);   // before any occupancy query
  (void)attr;
  Tile t;
  if (!pick_tile(128, N, Ho, Wo, &t)) return -10;
  hz::AMaps am;
  if (!make_x_maps(&am, x, N, H, W, Cin, stride, t)) return -11;
  constexpr int BLOCK_N = 64;
  CUtensorMap bm;
  if (!make_map2(&bm, w, (long long)R * S_ * Cin, Cout, (long long)R * S_ * Cin, 64, BLOCK_N)) return -12;
  hz::IgemmParams p;
  memset(&p, 0, sizeof(p));
  input_taps(&p.cls[0], R, S_, stride, pad, Ho, Wo, H, W, Cin, false);
  p.num_classes = 1;
  p.cblocks = (Cin + 63) / 64;
  p.BN = t.BN; p.BH = t.BH; p.BW = t.BW; p.tiles_per_img = t.per_img;
  p.n_images = N;
  p.out_n_stride = (long long)Ho * Wo * Cout; p.out_h_stride = (long long)Wo * Cout; p.out_w_stride = Cout;
  p.ncols = Cout;
  p.out = (__nv_bfloat16*)y;
  p.addend = nullptr;
  p.stats = stats;
  if (stats && !stats_is_zero) hz::zero_f32(stats, (size_t)2 * Cout, st);
  if (bn == nullptr && use_persistent(t.tiles * ((Cout + BLOCK_N - 1) / BLOCK_N))) {
    if (persist_wide(Cout)) {
      CUtensorMap bm2;      // weight box of 128 rows (output channels)
      if (!make_map2(&bm2, w, (long long)R * S_ * Cin, Cout, (long long)R * S_ * Cin, 64, 128)) return -12;
      return launch_persistent_n<128, false>(am, bm2, p, t.tiles, st);
    }
    return launch_persistent_n<64, false>(am, bm, p, t.tiles, st);
  }
  {
    const SplitWs w = get_split_ws();
    const int tiles = t.tiles * ((Cout + BLOCK_N - 1) / BLOCK_N);
    p.splits = w.ws ? pick_splits(tiles, p.cls[0].n * p.cblocks) : 1;
    p.ws = w.ws; p.sem = w.sem;
    if (bn != nullptr) p.splits = 1;      // workspace split-K retires CTAs early: incompatible with a grid barrier
    if (p.splits == 1) {
      p.splits = pick_cluster_splits(tiles, p.cls[0].n * p.cblocks, [](int c) {
        return max_active_clusters(hz::igemm_kernel<64, false>, hz::IgemmSmem<64>::kTotal, c);
      });
      p.cluster = p.splits > 1;
    }
    if (bn != nullptr) {
      if (stats == nullptr || !stats_is_zero) return -21;
      if (Cout % BLOCK_N) return -20;            // the fused BN epilogue assumes full 64-column tiles
      // every CTA must be resident for the barrier: clusters are already limited to one wave by
      // pick_cluster_splits; plain grids keep a few SMs spare for kernels of other streams (NCCL p2p)
      if (p.cluster ? tiles > max_active_clusters(hz::igemm_kernel<64, false>, hz::IgemmSmem<64>::kTotal, p.splits)
                    : tiles > hz_num_sms() - 16)
        return -20;
      p.bn_out = (__nv_bfloat16*)bn->out;
      p.bn_residual = (const __nv_bfloat16*)bn->residual;
      p.bn_gamma = bn->gamma; p.bn_beta = bn->beta;
      p.bn_mean = bn->mean; p.bn_invstd = bn->invstd;
      p.bn_rmean = bn->rmean; p.bn_rvar = bn->rvar;
      p.bn_counter = bn->counter;
      const long long M = (long long)N * Ho * Wo;
      p.bn_inv_count = 1.f / (float)M;
      p.bn_unbias = (float)M / (float)(M > 1 ? M - 1 : 1);
      p.bn_eps = bn->eps; p.bn_momentum = bn->momentum; p.bn_relu = bn->relu;
    }
  }
  p.prefetch_b = weights_stable ? prefetch_weights_enabled() : 0;
  p.dbg = g_conv_dbg;
  using SM = hz::IgemmSmem<BLOCK_N>;
  dim3 grid(t.tiles, (Cout + BLOCK_N - 1) / BLOCK_N, p.splits);
  return hz::launch_cluster(hz::igemm_kernel<BLOCK_N, false>, grid, dim3(128), SM::kTotal, st,
                            p.cluster ? (unsigned)p.splits : 1u, am, bm, p) == cudaSuccess ? 0 : -1;
}

}  // extern "C"

namespace {
// dx[N,H,W,Cin] = conv_transpose(dy[N,Ho,Wo,Cout], w);  bnb: also the BatchNorm-backward sums of the layer that produced x
int conv_dgrad_impl(const void* dy, const void* w, void* dx, const void* addend, int N, int H, int W, int Cin, int Cout, int R,
                    int stride, int pad, int weights_stable, const HzBnBwd* bnb, cudaStream_t st) {
  const int S_ = R;
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S_) / stride + 1;
  // output lattice per class: stride 1 -> (H,W); stride 2 -> (H/2,W/2) == (Ho,Wo)
  const int Lh = stride == 1 ? H : H / 2, Lw = stride == 1 ? W : W / 2;
  if (stride == 2 && (Lh != Ho || Lw != Wo)) return -13;
  static bool attr = set_smem(hz::igemm_kernel<64, true>, hz::IgemmSmem<64>::kTotal);    // before any occupancy query
  (void)attr;
  Tile t;
  if (!pick_tile(128, N, Lh, Lw, &t)) return -10;
  hz::AMaps am;
  if (!make_map4(&am.m[0], dy, Cout, Wo, Ho, N, Cout, (long long)Wo * Cout, (long long)Ho * Wo * Cout, 64, t.BW,
                 t.BH, t.BN))
    return -11;
  for (int i = 1; i < 4; ++i) am.m[i] = am.m[0];
  constexpr int BLOCK_N = 64;
  CUtensorMap bm;   // MN-major B: rows = Cout (K), inner = (r,s,ci)
  if (!make_map2(&bm, w, (long long)R * S_ * Cin, Cout, (long long)R * S_ * Cin, 64, 64)) return -12;
  hz::IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.num_classes = stride == 1 ? 1 : 4;
  for (int c = 0; c < p.num_classes; ++c) {
    const int ph = c >> 1, pw = c & 1;
    hz::TapList& tl = p.cls[c];
    tl.n = 0;
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S_; ++s) {
        int dh, dw;
        if (stride == 1) { dh = pad - r; dw = pad - s; }
        else {
          if (((ph + pad - r) & 1) || ((pw + pad - s) & 1)) continue;
          dh = (ph + pad - r) / 2; dw = (pw + pad - s) / 2;
        }
        if (!tap_hits(dh, Lh, Ho) || !tap_hits(dw, Lw, Wo)) continue;
        const int i = tl.n++;
        tl.dh[i] = (int8_t)dh; tl.dw[i] = (int8_t)dw; tl.map[i] = 0;
        tl.bk[i] = (r * S_ + s) * Cin;
      }
    p.cls_out_off[c] = stride == 1 ? 0 : ((long long)ph * W + pw) * Cin;
  }
  p.cblocks = (Cout + 63) / 64;
  p.BN = t.BN; p.BH = t.BH; p.BW = t.BW; p.tiles_per_img = t.per_img;
  p.n_images = N;
  p.out_n_stride = (long long)H * W * Cin;
  p.out_h_stride = (long long)stride * W * Cin;
  p.out_w_stride = (long long)stride * Cin;
  p.ncols = Cin;
  p.out = (__nv_bfloat16*)dx;
  p.addend = (const __nv_bfloat16*)addend;
  p.stats = nullptr;
  if (bnb != nullptr) {
    if (bnb->sums == nullptr || bnb->yraw == nullptr || bnb->mean == nullptr || bnb->invstd == nullptr) return -15;
    if (!bnb->sums_is_zero) hz::zero_f32(bnb->sums, (size_t)2 * Cin, st);
    p.stats = bnb->sums;
    p.bnb_out = (const __nv_bfloat16*)bnb->out;
    p.bnb_yraw = (const __nv_bfloat16*)bnb->yraw;
    p.bnb_mean = bnb->mean; p.bnb_invstd = bnb->invstd;
    p.bnb_cap6 = bnb->cap6;
  }
  if (bnb == nullptr && use_persistent(t.tiles * ((Cin + BLOCK_N - 1) / BLOCK_N) * p.num_classes))
    return persist_wide(Cin) ? launch_persistent_n<128, true>(am, bm, p, t.tiles, st)      // same 64 x 64 weight boxes,
                             : launch_persistent_n<64, true>(am, bm, p, t.tiles, st);      // two MN atoms per stage
  {
    const SplitWs w = get_split_ws();
    const int tiles = t.tiles * ((Cin + BLOCK_N - 1) / BLOCK_N) * p.num_classes;
    int kmax = 0;
    for (int c = 0; c < p.num_classes; ++c) kmax = kmax > p.cls[c].n ? kmax : p.cls[c].n;
    p.splits = w.ws ? pick_splits(tiles, kmax * p.cblocks) : 1;
    p.ws = w.ws; p.sem = w.sem;
    if (p.splits == 1) {
      p.splits = pick_cluster_splits(tiles, kmax * p.cblocks, [](int c) {
        return max_active_clusters(hz::igemm_kernel<64, true>, hz::IgemmSmem<64>::kTotal, c);
      });
      p.cluster = p.splits > 1;
    }
  }
  p.prefetch_b = weights_stable ? prefetch_weights_enabled() : 0;
  p.dbg = g_conv_dbg;
  using SM = hz::IgemmSmem<BLOCK_N>;
  dim3 grid(t.tiles, (Cin + BLOCK_N - 1) / BLOCK_N, p.num_classes * p.splits);
  if (bnb != nullptr) {
    static bool attr2 = set_smem(hz::igemm_kernel<64, true, true>, hz::IgemmSmem<64>::kTotal);
    (void)attr2;
    return hz::launch_cluster(hz::igemm_kernel<BLOCK_N, true, true>, grid, dim3(128), SM::kTotal, st,
                              p.cluster ? (unsigned)p.splits : 1u, am, bm, p) == cudaSuccess ? 0 : -1;
  }
  return hz::launch_cluster(hz::igemm_kernel<BLOCK_N, true>, grid, dim3(128), SM::kTotal, st,
                            p.cluster ? (unsigned)p.splits : 1u, am, bm, p) == cudaSuccess ? 0 : -1;
}
}  // namespace

extern "C" {

int hz_conv_dgrad(const void* dy, const void* w, void* dx, const void* addend, int N, int H, int W, int Cin, int Cout, int R,
                  int stride, int pad, int weights_stable, cudaStream_t st) {
  return conv_dgrad_impl(dy, w, dx, addend, N, H, W, Cin, Cout, R, stride, pad, weights_stable, nullptr, st);
}

// dgrad whose epilogue also leaves the BatchNorm-backward sums of the layer that produced x (see IgemmParams::bnb_*)
int hz_conv_dgrad_bnbwd(const void* dy, const void* w, void* dx, const void* addend, int N, int H, int W, int Cin, int Cout,
                        int R, int stride, int pad, int weights_stable, const HzBnBwd* bnb, cudaStream_t st) {
  if (bnb == nullptr) return -15;
  return conv_dgrad_impl(dy, w, dx, addend, N, H, W, Cin, Cout, R, stride, pad, weights_stable, bnb, st);
}

// dw[Cout, R*S*Cin (ld_out)] (+)= dy^T * x_taps.   ld_out / n_valid allow the padded stem (Cin=192 -> 147)
int hz_conv_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Cout, int R,
                  int stride, int pad, int accumulate, int prezeroed, long long ld_out, int n_valid,
                  cudaStream_t st) {
  const int S_ = R;
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S_) / stride + 1;
  Tile t;
  if (!pick_tile(64, N, Ho, Wo, &t)) return -10;
  CUtensorMap dym;
  if (!make_map4(&dym, dy, Cout, Wo, Ho, N, Cout, (long long)Wo * Cout, (long long)Ho * Wo * Cout, 64, t.BW, t.BH,
                 t.BN))
    return -11;
  hz::AMaps xm;
  if (!make_x_maps(&xm, x, N, H, W, Cin, stride, t)) return -11;
  hz::WgradParams p;
  memset(&p, 0, sizeof(p));
  input_taps(&p.taps, R, S_, stride, pad, Ho, Wo, H, W, Cin, true);
  if (p.taps.n == 0) return 0;
  constexpr int BLOCK_N = 64;
  p.KBN = t.BN; p.KBH = t.BH; p.kb_per_img = t.per_img;
  p.kblocks = t.tiles;
  p.n_tiles = (Cin + BLOCK_N - 1) / BLOCK_N;
  const int m_tiles = (Cout + 127) / 128;
  const int ctas = m_tiles * p.n_tiles * p.taps.n;
  // small problems: leave SMs for the concurrently running dgrad chain; in throughput mode (see use_persistent) a long
  // pixel reduction is spread over two waves of CTAs instead
  const int target = (g_persist_mode != 0 && p.kblocks >= 512) ? 2 * hz_num_sms() : 72;
  int splits = (target + ctas - 1) / ctas;
  if (splits > p.kblocks) splits = p.kblocks;
  if (splits > 32) splits = 32;
  if (splits < 1) splits = 1;
  p.splits = splits;
  p.Cout = Cout;
  p.ld_out = ld_out > 0 ? ld_out : (long long)R * S_ * Cin;
  p.tap_stride = Cin;
  p.n_valid = n_valid > 0 ? n_valid : Cin;
  p.mode = splits > 1 ? 2 : (accumulate ? 1 : 0);
  p.out = dw;
  if (splits > 1 && !accumulate && !prezeroed) {
    // split-K accumulates with atomics: clear exactly the region this conv owns (rows are ld_out apart)
    if (p.ld_out == (long long)R * S_ * Cin || R == 1)
      hz::zero_f32(dw, (size_t)Cout * p.ld_out, st);
    else
      return -14;
  }
  using SM = hz::WgradSmem<BLOCK_N>;
  static bool attr = set_smem(hz::wgrad_kernel<BLOCK_N>, SM::kTotal);
  (void)attr;
  dim3 grid(m_tiles * p.n_tiles, p.taps.n, splits);
  return hz::launch(hz::wgrad_kernel<BLOCK_N>, grid, dim3(128), SM::kTotal, st, dym, xm, p) == cudaSuccess ? 0 : -1;
}

}  // extern "C"

// Tensor-parallel fused compute+collective kernels: one tcgen05 implicit-GEMM kernel
// that ALSO moves its data over NVLink from inside the kernel, tile by tile.
//
//   GEMM -> all-reduce / reduce-scatter  (row-parallel conv/linear forward; column-parallel dgrad):
//       every rank computes a partial 128xBLOCK_N tile in TMEM.  Tiles are owned round-robin
//       (tile % W).  A non-owner *pushes* its fp32 partial straight into the owner's workspace slot
//       with NVLink stores and raises a per-tile flag; the owner waits for the W-1 flags, adds the
//       partials in rank order (deterministic), converts to bf16 and - for all-reduce - broadcasts the
//       finished tile into every rank's output buffer with peer stores, then raises the tile's result
//       flag on every rank.  No NCCL call, no separate reduction pass: the transfer of tile i overlaps
//       the MMA of tile j running on another SM.   (reference site: the missing reduction of
//       tensor_parallel_train.py:215-218 / SURVEY Q4; row-split is required by BASELINE.json.)
//
//   all-gather -> GEMM  (A operand row/image-sharded across ranks):
//       the TMA producer loads A tiles *directly from the owning peer's memory* (tensor maps built on
//       the peer-mapped addresses) after a per-kernel ready-flag handshake - the all-gather never
//       materialises.   (reference site: the ws-broadcast "all-gather" of tensor_parallel_train.py:49-62.)
//
// All buffers live in a symmetric heap (same offsets on every rank, CUDA-IPC mapped).  Flags are epoch
// numbered (device-side counter) so the kernels are re-launchable / CUDA-graph replayable; spins are
// bounded (trap instead of hang).  Grids are <= #SMs so all CTAs are co-resident while they spin.
#include <cuda.h>

#include "igemm_common.cuh"
#include "launchers.h"

namespace hz {

constexpr int kTpMaxRanks = 8;

struct PeerAMaps {
  CUtensorMap m[kTpMaxRanks];     // A tensor map on rank r's buffer (AG mode); m[rank] is the local one
};

struct TpParams {
  TapList taps;
  long long out_n_stride, out_h_stride, out_w_stride;   // elements
  int cblocks;
  int BN, BH, BW, tiles_per_img;
  int n_images;                  // total images (all ranks)
  int ncols;
  // ---- peer part
  int world, rank;
  int reduce;                    // 1: push-to-owner reduce (owner = tile % world); 2: one-shot (push to all, reduce everywhere)
  int bcast;                     // 1: all-reduce (owner broadcasts), 0: reduce-scatter (owner keeps)
  int ag;                        // 1: A is image-sharded, ag_imgs images per rank
  int ag_imgs;
  char* heap[kTpMaxRanks];       // symmetric heap base of every rank
  long long out_off;             // bf16 output [n_images, ...] (bytes from heap base)
  long long ws_off;              // fp32 partial slots [world][tiles][128][BLOCK_N]
  long long ws_stride;           // one-shot: second copy of the slots, selected by call parity (no back-pressure needed)
  long long arrive_off;          // u32 [tiles][world]
  long long result_off;          // u32 [tiles]
  long long ready_off;           // u32 [world]      (AG: "my A shard is ready")
  unsigned* epoch;               // local: number of completed calls
  unsigned* done;                // local: CTAs finished in this call
};

HZ_DEVINL void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
HZ_DEVINL unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
HZ_DEVINL void spin_until_ge(const unsigned* p, unsigned e) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys_u32(p) - e) < 0) {
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}

template <int BLOCK_N, bool B_MN>
__global__ void __launch_bounds__(128) igemm_tp_kernel(const __grid_constant__ PeerAMaps amaps,
                                                       const __grid_constant__ CUtensorMap bmap,
                                                       const __grid_constant__ TpParams p) {
  using S = IgemmSmem<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, nt = blockIdx.y;
  const int tiles = gridDim.x * gridDim.y;
  const int tile = nt * gridDim.x + mt;
  const int n0 = (p.BN == 1) ? mt / p.tiles_per_img : mt * p.BN;
  const int h0 = (p.BN == 1) ? (mt % p.tiles_per_img) * p.BH : 0;
  const int k_iters = p.taps.n * p.cblocks;
  const int W = p.world, me = p.rank;
  const unsigned e = *p.epoch + 1u;             // epoch of this call (same on every rank)
  char* my_heap = p.heap[me];
  const long long ws_off = p.ws_off + (long long)(e & 1u) * p.ws_stride;

  // which rank holds this tile's A rows (all-gather mode) and the image index inside that shard
  const int src = p.ag ? min(n0 / p.ag_imgs, W - 1) : me;
  const int n0_src = p.ag ? n0 - src * p.ag_imgs : n0;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&amaps.m[src]);
    tc::prefetch_tmap(&bmap);
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, BLOCK_N);
    tc::tmem_relinquish();
  }
  if (p.ag && warp == 3) {
    // ready handshake: our shard was produced by earlier kernels of this stream, so any CTA may vouch
    // for it; then wait until the shard we are about to read is published by its owner
    if (blockIdx.x == 0 && blockIdx.y == 0 && lane < W)
      st_release_sys_u32(reinterpret_cast<unsigned*>(p.heap[lane] + p.ready_off) + me, e);
    if (lane == 0 && src != me) spin_until_ge(reinterpret_cast<unsigned*>(my_heap + p.ready_off) + src, e);
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (A possibly straight out of a peer's HBM over NVLink) ==========
    const CUtensorMap* am = &amaps.m[src];
    for (int it = 0; it < k_iters; ++it) {
      const int t = it / p.cblocks, cb = it % p.cblocks;
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      tc::mbar_wait(&empty[s], ph ^ 1);
      uint8_t* sa = smem + s * S::kStageBytes;
      uint8_t* sb = sa + kABytes;
      tc::mbar_arrive_expect_tx(&full[s], S::kStageBytes);
      tc::tma_load_4d(sa, am, &full[s], cb * kKBlock, p.taps.dw[t], h0 + p.taps.dh[t], n0_src);
      if (!B_MN) {
        tc::tma_load_2d(sb, &bmap, &full[s], p.taps.bk[t] + cb * kKBlock, nt * BLOCK_N);
      } else {
#pragma unroll
        for (int j = 0; j < BLOCK_N / 64; ++j)
          tc::tma_load_2d(sb + j * 8192, &bmap, &full[s], p.taps.bk[t] + nt * BLOCK_N + j * 64, cb * kKBlock);
      }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = tc::make_idesc(kTileM, BLOCK_N, false, B_MN);
    for (int it = 0; it < k_iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      tc::mbar_wait(&full[s], ph);
      tc::fence_after_sync();
      const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
      const uint32_t sb = sa + kABytes;
#pragma unroll
      for (int k = 0; k < kKBlock / 16; ++k) {
        const uint64_t da = tc::make_sdesc(sa + k * 32, 16, 1024);
        const uint64_t db = B_MN ? tc::make_sdesc(sb + k * 2048, 8192, 1024) : tc::make_sdesc(sb + k * 32, 16, 1024);
        tc::umma_f16(tmem_d, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
      }
      tc::umma_commit(&empty[s]);
    }
    tc::umma_commit(tmem_full);
  }
  __syncwarp();

  // ===================== epilogue =====================
  __nv_bfloat16* staging = reinterpret_cast<__nv_bfloat16*>(smem);
  const int row = warp * 32 + lane;
  tc::mbar_wait(tmem_full, 0);
  tc::fence_after_sync();

  const int owner = p.reduce == 1 ? tile % W : me;
  bool have_result = true;            // staging holds the final bf16 tile
  if (p.reduce == 2 && W > 1) {
    // ---- one-shot: push the fp32 partial into slot [me][tile] of EVERY peer (one NVLink hop), then each
    //      rank reduces all W partials itself in rank order -> bit-identical results, no second hop
#pragma unroll
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
      tc::tmem_ld_wait();
      for (int d = 1; d < W; ++d) {
        const int rr = (me + d) % W;
        float* dst = reinterpret_cast<float*>(p.heap[rr] + ws_off) + ((size_t)(me * tiles + tile) * kTileM + row) * BLOCK_N + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                            __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < W && threadIdx.x != me)
      st_release_sys_u32(reinterpret_cast<unsigned*>(p.heap[threadIdx.x] + p.arrive_off) + tile * W + me, e);
  }
  if (p.reduce == 1 && owner != me) {
    // ---- push the fp32 partial into the owner's slot [me][tile] over NVLink
    float* dst = reinterpret_cast<float*>(p.heap[owner] + ws_off) + ((size_t)(me * tiles + tile) * kTileM + row) * BLOCK_N;
#pragma unroll
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                               __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      st_release_sys_u32(reinterpret_cast<unsigned*>(p.heap[owner] + p.arrive_off) + tile * W + me, e);
    have_result = false;
  } else {
    if (p.reduce && W > 1) {
      if (threadIdx.x < W && threadIdx.x != me)
        spin_until_ge(reinterpret_cast<unsigned*>(my_heap + p.arrive_off) + tile * W + threadIdx.x, e);
      __syncthreads();
    }
    const float* slots = reinterpret_cast<const float*>(my_heap + ws_off);
#pragma unroll
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
      tc::tmem_ld_wait();
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.f;
      if (p.reduce) {
        for (int rr = 0; rr < W; ++rr) {          // fixed rank order: deterministic (and rank-identical) sum
          if (rr == me) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
          } else {
            const float* sp = slots + ((size_t)(rr * tiles + tile) * kTileM + row) * BLOCK_N + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(sp + j));
              acc[j] += v.x; acc[j + 1] += v.y; acc[j + 2] += v.z; acc[j + 3] += v.w;
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
      }
      __nv_bfloat16* dstg = staging + row * S::kStagingLd + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) st8(dstg + j, pack8(acc + j));
    }
  }
  tc::fence_before_sync();
  __syncthreads();

  constexpr int kVecPerRow = BLOCK_N / 8;
  constexpr int kRowsPerPass = 128 / kVecPerRow;
  const int vec = threadIdx.x % kVecPerRow;
  if (have_result) {
    // store the finished tile into the local output and (all-reduce) into every peer's output
    const int n_dst = (p.reduce == 1 && p.bcast) ? W : 1;
    for (int d = 0; d < n_dst; ++d) {
      const int rr = (p.reduce == 1 && p.bcast) ? (me + d) % W : me;   // start with self, stagger peers
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.heap[rr] + p.out_off);
      for (int r0 = threadIdx.x / kVecPerRow; r0 < kTileM; r0 += kRowsPerPass) {
        const int wi = r0 % p.BW;
        const int hi = (r0 / p.BW) % p.BH;
        const int ni = r0 / (p.BW * p.BH);
        const int n = n0 + ni;
        if (n >= p.n_images) continue;
        const long long off = (long long)n * p.out_n_stride + (long long)(h0 + hi) * p.out_h_stride +
                              (long long)wi * p.out_w_stride + nt * BLOCK_N + vec * 8;
        st8(out + off, ld8(staging + r0 * S::kStagingLd + vec * 8));
      }
    }
    if (p.reduce == 1 && p.bcast && W > 1) {
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x < W && threadIdx.x != me)
        st_release_sys_u32(reinterpret_cast<unsigned*>(p.heap[threadIdx.x] + p.result_off) + tile, e);
    }
  } else if (p.bcast) {
    // non-owner of an all-reduce: the owner delivers the finished tile into our output buffer
    if (threadIdx.x == 0) spin_until_ge(reinterpret_cast<unsigned*>(my_heap + p.result_off) + tile, e);
  }
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_d, BLOCK_N);
  // ---- the last CTA of the grid publishes the new epoch (every CTA has already read the old one)
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned fin = atomicAdd(p.done, 1u);
    if (fin == (unsigned)(tiles - 1)) {
      *p.done = 0u;
      *p.epoch = e;
      __threadfence();
    }
  }
}

}  // namespace hz

// ================================================================================================
// host
// ================================================================================================
using namespace hz::host;

extern "C" {

// layout helper: bytes needed in the symmetric heap for the flag/workspace part of one fused op
size_t hz_tp_ws_bytes(int world, int tiles) {
  return (size_t)world * tiles * 128 * 64 * sizeof(float);
}

// y (in the symmetric heap at out_off on every rank) = [all-reduce|reduce-scatter] over ranks of conv(x_r, w_r)
// or, with ag=1, conv(all-gather(x shards), w_local).  kind: 0 = forward (w K-major), 1 = dgrad (w MN-major).
// heaps[r]: heap base of rank r.  x_ptrs[r]: rank r's A buffer (only [rank] is used unless ag).
// Stride-1 convs and dense GEMMs only (H=W=1, R=1).
int hz_tp_conv(int kind, const void* const* x_ptrs, const void* w, char* const* heaps, long long out_off,
               long long ws_off, long long ws_stride, long long arrive_off, long long result_off,
               long long ready_off, unsigned* epoch, unsigned* done, int world, int rank, int reduce, int bcast,
               int ag, int N, int H, int W_, int Cin, int Cout, int R, int pad, cudaStream_t st) {
  // logical conv: x [N,H,W,Cin] -> y [N,H,W,Cout] (fwd)   |   dy [N,H,W,Cout] -> dx [N,H,W,Cin] (dgrad)
  const int S_ = R;
  const int Ka = kind == 0 ? Cin : Cout;       // channels of the A operand
  const int Nn = kind == 0 ? Cout : Cin;       // output channels
  if (Ka % 64 || Nn % 64) return -20;
  Tile t;
  if (!pick_tile(128, N, H, W_, &t)) return -10;
  constexpr int BLOCK_N = 64;
  const int tiles = t.tiles * (Nn / BLOCK_N);
  if (tiles > 148) return -21;                 // all CTAs must be co-resident (they spin on peers)
  const int imgs_per_rank = ag ? N / world : N;
  if (ag && (N % world || (t.BN > 1 && imgs_per_rank % t.BN))) return -22;
  hz::PeerAMaps am;
  memset(&am, 0, sizeof(am));
  for (int r = 0; r < world; ++r) {
    const void* base = ag ? x_ptrs[r] : x_ptrs[rank];
    if (!ag && r != rank) { am.m[r] = am.m[rank]; continue; }
    if (!make_map4(&am.m[r], base, Ka, W_, H, imgs_per_rank, Ka, (long long)W_ * Ka, (long long)H * W_ * Ka, 64, t.BW,
                   t.BH, t.BN))
      return -11;
    if (!ag) break;
  }
  if (!ag) for (int r = 0; r < world; ++r) if (r != rank) am.m[r] = am.m[rank];
  CUtensorMap bm;
  if (kind == 0) {
    if (!make_map2(&bm, w, (long long)R * S_ * Cin, Cout, (long long)R * S_ * Cin, 64, BLOCK_N)) return -12;
  } else {
    if (!make_map2(&bm, w, (long long)R * S_ * Cin, Cout, (long long)R * S_ * Cin, 64, 64)) return -12;
  }
  hz::TpParams p;
  memset(&p, 0, sizeof(p));
  p.taps.n = 0;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S_; ++s) {
      const int dh = kind == 0 ? r - pad : pad - r, dw = kind == 0 ? s - pad : pad - s;
      if (!tap_hits(dh, H, H) || !tap_hits(dw, W_, W_)) continue;
      const int i = p.taps.n++;
      p.taps.dh[i] = (int8_t)dh; p.taps.dw[i] = (int8_t)dw; p.taps.map[i] = 0;
      p.taps.bk[i] = (r * S_ + s) * Cin;
    }
  p.cblocks = Ka / 64;
  p.BN = t.BN; p.BH = t.BH; p.BW = t.BW; p.tiles_per_img = t.per_img;
  p.n_images = N;
  p.out_n_stride = (long long)H * W_ * Nn; p.out_h_stride = (long long)W_ * Nn; p.out_w_stride = Nn;
  p.ncols = Nn;
  p.world = world; p.rank = rank; p.reduce = reduce; p.bcast = bcast; p.ag = ag; p.ag_imgs = imgs_per_rank;
  for (int r = 0; r < world; ++r) p.heap[r] = heaps[r];
  p.out_off = out_off; p.ws_off = ws_off; p.ws_stride = ws_stride; p.arrive_off = arrive_off; p.result_off = result_off;
  p.ready_off = ready_off;
  p.epoch = epoch; p.done = done;
  using SM = hz::IgemmSmem<BLOCK_N>;
  dim3 grid(t.tiles, Nn / BLOCK_N, 1);
  if (kind == 0) {
    static bool attr = set_smem(hz::igemm_tp_kernel<BLOCK_N, false>, SM::kTotal);
    (void)attr;
    hz::igemm_tp_kernel<BLOCK_N, false><<<grid, 128, SM::kTotal, st>>>(am, bm, p);
  } else {
    static bool attr = set_smem(hz::igemm_tp_kernel<BLOCK_N, true>, SM::kTotal);
    (void)attr;
    hz::igemm_tp_kernel<BLOCK_N, true><<<grid, 128, SM::kTotal, st>>>(am, bm, p);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"

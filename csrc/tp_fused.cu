// Tensor-parallel fused compute+collective kernels: ONE tcgen05 implicit-GEMM kernel that
// also performs its collective over NVLink / NVSwitch from inside the kernel, tile by tile.
//
//   GEMM -> all-reduce   (row-parallel conv/linear forward; column-parallel dgrad)      mode 1
//   GEMM -> reduce-scatter (tile t is reduced and kept by rank t % W only)              mode 2
//       every rank computes its partial 128x64 tile in TMEM, drops it as bf16 into ITS OWN slot of the symmetric
//       heap (a local store, no NVLink traffic yet) and bumps the tile's arrival counter on the ranks that need
//       the tile - with ONE `multimem.red` on the NVSwitch multicast address when the heap is multicast-mapped,
//       else one `red.release.sys` per peer.  A rank that needs the tile waits for the W arrivals and then
//       *pulls*: either `multimem.ld_reduce` (the switch adds the W copies in flight - one NVLink round trip,
//       1/W of the ingress bytes) or W `cp.async.bulk` copies of the peers' slots straight into the (now idle)
//       operand ring in shared memory, summed in rank order (bit-identical on every rank).  The reduced tile goes
//       through the normal conv epilogue: BatchNorm partial sums, optional residual-gradient addend, bf16 store
//       into a plain local tensor.  Round 1 pushed fp32 partials to every peer and used 7 flags per tile: 2x the
//       bytes, W-1 posted-store streams per thread and a flag fan-out - it lost to conv + NCCL at 8 GPUs.
//       (reference site: the missing reduction of tensor_parallel_train.py:215-218 / SURVEY Q4; the row split
//       is required by BASELINE.json.)
//
//   all-gather -> GEMM   (A operand image-sharded across ranks)                          ag = 1
//       the TMA producer loads A tiles *directly from the owning peer's memory* (tensor maps built on the
//       peer-mapped addresses) after a ready-flag handshake - the gathered tensor never materialises.
//       (reference site: the ws-broadcast "all-gather" of tensor_parallel_train.py:49-62.)
//
// Channel counts need not be multiples of 64: a 64-wide TMA box over a narrower tensor is zero-filled by the
// TMA unit (and columns past `ncols` are masked in the epilogue), so the 32-channel shards of layer3 at W=8 run
// here too.  Stride-2 dgrad runs as 4 output-parity classes (blockIdx.z) like the dense kernel.
// All counters are epoch based and live in device memory: re-launchable and CUDA-graph replayable; partial
// slots alternate by call parity (every call contains an all-to-all dependency, so a slot is never rewritten
// while a peer can still read it).  Spins are bounded (trap instead of a hang); grids are <= #SMs so all CTAs
// are co-resident while they spin.  Programmatic dependent launch like every other compute kernel.
#include <cuda.h>

#include "igemm_common.cuh"
#include "launchers.h"

namespace hz {

constexpr int kTpMaxRanks = 8;
constexpr int kTpPartBytes = kTileM * 64 * 2;      // one bf16 partial tile

struct PeerAMaps {
  CUtensorMap m[kTpMaxRanks];     // A tensor map on rank r's buffer (AG mode); m[rank] is the local one
};

struct TpParams {
  TapList cls[4];
  long long cls_out_off[4];
  long long out_n_stride, out_h_stride, out_w_stride;   // elements
  int num_classes, cblocks;
  int BN, BH, BW, tiles_per_img;
  int n_images;                  // total images (all ranks)
  int ncols;                     // valid output columns
  // ---- peer part
  int world, rank;
  int mode;                      // 0: no reduction, 1: all-reduce, 2: reduce-scatter (tile % world keeps)
  int nvls;                      // 1: the heap is multicast-mapped: multimem.st / multimem.red / multimem.ld_reduce
  int ll;                        // 1: latency protocol (flag-in-data push), 0: bandwidth protocol (counter + pull)
  int ag;                        // 1: A is image-sharded, ag_imgs images per rank
  int ag_imgs;
  char* heap[kTpMaxRanks];       // symmetric heap base of every rank, as mapped in this process
  char* mc_heap;                 // multicast mapping of the same heap (nvls)
  long long part_off;            // partial tiles (bytes from heap base): bandwidth protocol bf16 [2 parities][tiles][128][64];
  long long part_stride;         //   latency protocol LL words [2][world][tiles][128][64] (2x the bytes); stride of a parity
  long long cnt_off;             // u32 [tiles] arrival counters (monotonic: += W per call)
  long long ready_off;           // u32 [world]      (AG: "my A shard is ready")
  unsigned* epoch;               // local u32 [tiles]: calls completed by each tile's CTA (every CTA keeps its own
                                 // counter: no cross-CTA atomic / fence at the tail of the kernel)
  __nv_bfloat16* out;            // local output tensor
  const __nv_bfloat16* addend;   // optional, layout of out: out = reduced tile + addend (residual gradient)
  float* stats;                  // optional [2*ncols]: sum y, sum y^2 of the REDUCED output (pre-zeroed)
  long long timeout_clk;
  long long* dbg;                // optional: per-CTA clock64 stamps of the kernel's phases (tools/tp_timeline.py)
};

HZ_DEVINL void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
HZ_DEVINL unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
HZ_DEVINL void red_release_sys_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// one instruction, delivered by the NVSwitch to the same offset on every rank of the multicast group
HZ_DEVINL void multimem_red_add_u32(unsigned* mc_p, unsigned v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_p), "r"(v) : "memory");
}
HZ_DEVINL uint4 multimem_ld_reduce_bf16x8(const void* mc_p) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc_p) : "memory");
  return v;
}
HZ_DEVINL void spin_until_ge(const unsigned* p, unsigned e, long long timeout_clk) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys_u32(p) - e) < 0) {
    if (clock64() - t0 > timeout_clk) __trap();
  }
}
// "LL" words: 16 bytes = {data, flag, data, flag}.  8-byte halves are single-copy atomic, so a receiver that sees the
// flag sees the data next to it - no fence between payload and flag (the protocol NCCL uses for small messages).
HZ_DEVINL void ll_store(void* dst, const uint4& v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
HZ_DEVINL void ll_mc_store(void* mc_dst, const uint4& v) {      // replicated to every rank by the NVSwitch
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_dst), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w) : "memory");
}
HZ_DEVINL uint4 ll_load(const void* src) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
  return v;
}
// 1-D bulk copy global (possibly a peer's HBM over NVLink) -> shared, completion on an mbarrier
HZ_DEVINL void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool B_MN>
__global__ void __launch_bounds__(128) igemm_tp_kernel(const __grid_constant__ PeerAMaps amaps,
                                                       const __grid_constant__ CUtensorMap bmap,
                                                       const __grid_constant__ TpParams p) {
  constexpr int BLOCK_N = 64;
  using S = IgemmSmem<BLOCK_N>;
  static_assert(kTpMaxRanks * kTpPartBytes <= S::kPipeBytes, "peer partial tiles are pulled into the idle operand ring");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint64_t* pull_bar = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pull_bar + 1);

  pdl_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, nt = blockIdx.y, cls = blockIdx.z;
  const int tiles = gridDim.x * gridDim.y * gridDim.z;
  const int tile = (cls * gridDim.y + nt) * gridDim.x + mt;
  long long* dbg = p.dbg ? p.dbg + (size_t)tile * 16 : nullptr;
#define HZ_STAMP(i) do { if (dbg != nullptr && threadIdx.x == 0) dbg[i] = clock64(); } while (0)
  HZ_STAMP(0);                                   // kernel entry
  const TapList& taps = p.cls[cls];
  const int n0 = (p.BN == 1) ? mt / p.tiles_per_img : mt * p.BN;
  const int h0 = (p.BN == 1) ? (mt % p.tiles_per_img) * p.BH : 0;
  const int k_iters = taps.n * p.cblocks;
  const int W = p.world, me = p.rank;
  char* my_heap = p.heap[me];

  // which rank holds this tile's A rows (all-gather mode) and the image index inside that shard
  const int src = p.ag ? min(n0 / p.ag_imgs, W - 1) : me;
  const int n0_src = p.ag ? n0 - src * p.ag_imgs : n0;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&amaps.m[src]);
    tc::prefetch_tmap(&bmap);
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::mbar_init(pull_bar, 1);
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
  HZ_STAMP(1);                                  // prologue done
  pdl_wait();                                   // the A operand / addend / epoch counter come from upstream kernels
  HZ_STAMP(2);                                  // upstream kernel complete
  const unsigned e = p.epoch[tile] + 1u;        // epoch of this call (same for every tile and on every rank)

  if (p.ag && warp == 3) {
    // ready handshake: our shard was produced by earlier kernels of this stream (complete: pdl_wait above), so
    // any CTA may vouch for it; then wait until the shard we are about to read is published by its owner
    if (tile == 0 && lane < W) st_release_sys_u32(reinterpret_cast<unsigned*>(p.heap[lane] + p.ready_off) + me, e);
    if (lane == 0 && src != me)
      spin_until_ge(reinterpret_cast<unsigned*>(my_heap + p.ready_off) + src, e, p.timeout_clk);
    __syncwarp();
  }
  if (p.ag) __syncthreads();

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
      tc::tma_load_4d(sa, am, &full[s], cb * kKBlock, taps.dw[t], h0 + taps.dh[t], n0_src);
      if (!B_MN) {
        tc::tma_load_2d(sb, &bmap, &full[s], taps.bk[t] + cb * kKBlock, nt * BLOCK_N);
      } else {
        tc::tma_load_2d(sb, &bmap, &full[s], taps.bk[t] + nt * BLOCK_N, cb * kKBlock);
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
    if (k_iters > 0) tc::umma_commit(tmem_full);
  }
  __syncwarp();

  // ===================== epilogue =====================
  const int row = warp * 32 + lane;
  constexpr int kVecPerRow = BLOCK_N / 8;              // 16-byte vectors per tile row
  constexpr int kRowsPerPass = 128 / kVecPerRow;
  constexpr int kPasses = kTileM / kRowsPerPass;
  const int vec = threadIdx.x % kVecPerRow;
  const bool col_ok = nt * BLOCK_N + vec * 8 < p.ncols;
  // global element offsets of the rows this thread stores, computed while the MMAs are still running
  long long offs[kPasses];
#pragma unroll
  for (int i = 0; i < kPasses; ++i) {
    const int r0 = threadIdx.x / kVecPerRow + i * kRowsPerPass;
    const int wi = r0 % p.BW;
    const int hi = (r0 / p.BW) % p.BH;
    const int n = n0 + r0 / (p.BW * p.BH);
    offs[i] = (n < p.n_images && col_ok)
                  ? (long long)n * p.out_n_stride + (long long)(h0 + hi) * p.out_h_stride +
                        (long long)wi * p.out_w_stride + p.cls_out_off[cls] + nt * BLOCK_N + vec * 8
                  : -1;
  }
  const bool reducing = p.mode != 0 && W > 1;
  const bool keeper = !reducing || p.mode == 1 || (tile % W) == me;     // this rank finishes / stores the tile
  bf16x8 addv[kPasses];
  if (p.addend != nullptr && keeper) {
#pragma unroll
    for (int i = 0; i < kPasses; ++i)
      if (offs[i] >= 0) addv[i] = ld8(p.addend + offs[i]);
  }
  if (k_iters > 0) {
    tc::mbar_wait(tmem_full, 0);
    tc::fence_after_sync();
  }
  HZ_STAMP(3);                                  // accumulator complete

  // ---- accumulator -> bf16 tile in shared memory (the operand ring is idle now)
  __nv_bfloat16* staging = reinterpret_cast<__nv_bfloat16*>(smem + S::kPipeBytes - S::kStagingBytes);   // [128][kStagingLd]
  {
#pragma unroll
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      if (k_iters > 0) {
        tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;            // a parity class without taps: exactly zero
      }
      __nv_bfloat16* dst = staging + row * S::kStagingLd + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[j + i]);
        st8(dst + j, pack8(f));
      }
    }
    tc::fence_before_sync();
    __syncthreads();
  }
  uint4 red[kPasses];                                                        // reduced bf16x8 vectors (modes 1, 2)
  if (reducing && p.ll) {
    // ================= latency protocol ("LL"): data and flag travel in the same 8 bytes =================
    // Every 16-byte store carries {2 bf16, epoch, 2 bf16, epoch}: the receiver polls the payload itself, so there is
    // no fence, no separate flag and no counter on the critical path (phase stamps of the fence + counter protocol on
    // 2 GPUs: 7.3 us for the system fences, 3.4 us for the releasing arrival, 2 us waiting, 3.9 us pulling).
    // Push: my partial tile goes into slot [parity][me][tile] of every rank that keeps the tile - ONE multimem.st
    // per 16 bytes when the heap is multicast-mapped (the NVSwitch replicates it), else one store per peer.
    const long long slot0 = p.part_off + (long long)(e & 1u) * p.part_stride;
    const long long tile_bytes = 2LL * kTpPartBytes;                        // 32 KB on the wire per tile
    const long long my_slot = slot0 + ((long long)me * tiles + tile) * tile_bytes;
    const bool use_mc = p.nvls && p.mode == 1;
#pragma unroll
    for (int i = 0; i < kPasses; ++i) {
      const int r0 = threadIdx.x / kVecPerRow + i * kRowsPerPass;
      const bf16x8 v = ld8(staging + r0 * S::kStagingLd + vec * 8);
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
      const long long off = my_slot + (long long)(r0 * BLOCK_N + vec * 8) * 4;           // 4 wire bytes per element
      const uint4 lo = make_uint4(w[0], e, w[1], e), hi = make_uint4(w[2], e, w[3], e);
      if (use_mc) {
        ll_mc_store(p.mc_heap + off, lo);
        ll_mc_store(p.mc_heap + off + 16, hi);
      } else if (p.mode == 1) {
        for (int d = 0; d < W; ++d) {
          char* dst = p.heap[(me + d) % W] + off;
          ll_store(dst, lo);
          ll_store(dst + 16, hi);
        }
      } else {
        char* dst = p.heap[tile % W] + off;
        ll_store(dst, lo);
        ll_store(dst + 16, hi);
      }
    }
    HZ_STAMP(4);                                // partial tile pushed
    HZ_STAMP(5);
    if (keeper) {
      // Receive: every rank's copy of the tile sits in MY memory; poll the words until they carry this epoch and sum
      // in rank order (deterministic, bit-identical on every rank).
      float acc[kPasses][8];
#pragma unroll
      for (int i = 0; i < kPasses; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      const long long t0 = clock64();
      for (int rr = 0; rr < W; ++rr) {
        const char* src = my_heap + slot0 + ((long long)rr * tiles + tile) * tile_bytes;
        uint4 lo[kPasses], hi[kPasses];
#pragma unroll
        for (int i = 0; i < kPasses; ++i) {
          const int r0 = threadIdx.x / kVecPerRow + i * kRowsPerPass;
          const char* q = src + (long long)(r0 * BLOCK_N + vec * 8) * 4;
          lo[i] = ll_load(q);
          hi[i] = ll_load(q + 16);
        }
#pragma unroll
        for (int i = 0; i < kPasses; ++i) {
          const int r0 = threadIdx.x / kVecPerRow + i * kRowsPerPass;
          const char* q = src + (long long)(r0 * BLOCK_N + vec * 8) * 4;
          while (lo[i].y != e || lo[i].w != e || hi[i].y != e || hi[i].w != e) {
            if (clock64() - t0 > p.timeout_clk) __trap();
            lo[i] = ll_load(q);
            hi[i] = ll_load(q + 16);
          }
          const uint32_t w[4] = {lo[i].x, lo[i].z, hi[i].x, hi[i].z};
          float f[8];
          unpack8(*reinterpret_cast<const bf16x8*>(w), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] += f[j];
        }
        if (rr == 0) HZ_STAMP(6);               // first rank's tile has arrived
      }
#pragma unroll
      for (int i = 0; i < kPasses; ++i) {
        const bf16x8 pk = pack8(acc[i]);
        red[i] = *reinterpret_cast<const uint4*>(&pk);
      }
    }
  } else if (reducing) {
    // ================= bandwidth protocol: partial in my own slot, arrival counter, pull =================
    // ---- 1. my partial tile -> my own slot (row-major [128][64] bf16, coalesced 16-byte stores)
    const long long slot = p.part_off + (long long)(e & 1u) * p.part_stride + (long long)tile * kTpPartBytes;
#pragma unroll
    for (int i = 0; i < kPasses; ++i) {
      const int r0 = threadIdx.x / kVecPerRow + i * kRowsPerPass;
      st8(reinterpret_cast<__nv_bfloat16*>(my_heap + slot) + r0 * BLOCK_N + vec * 8, ld8(staging + r0 * S::kStagingLd + vec * 8));
    }
    // one releasing arrival per CTA orders the whole tile: the CTA barrier makes every thread's stores happen-before
    // thread 0's release (cumulativity) - no per-thread system fence (measured 7 us for 128 of them)
    __syncthreads();
    HZ_STAMP(4);                                // partial tile written
    // ---- 2. arrival: bump the tile's counter wherever the tile is needed
    unsigned* cnt_local = reinterpret_cast<unsigned*>(my_heap + p.cnt_off) + tile;
    if (threadIdx.x == 0) {
      if (p.mode == 1) {
        if (p.nvls) {
          multimem_red_add_u32(reinterpret_cast<unsigned*>(p.mc_heap + p.cnt_off) + tile, 1u);
        } else {
          for (int d = 0; d < W; ++d) red_release_sys_add_u32(reinterpret_cast<unsigned*>(p.heap[(me + d) % W] + p.cnt_off) + tile, 1u);
        }
      } else {
        red_release_sys_add_u32(reinterpret_cast<unsigned*>(p.heap[tile % W] + p.cnt_off) + tile, 1u);
      }
    }
    HZ_STAMP(5);                                // arrival posted
    if (keeper) {
      // ---- 3. wait for all W partials of this tile, 4. pull them
      if (threadIdx.x == 0) spin_until_ge(cnt_local, (unsigned)W * e, p.timeout_clk);
      __syncthreads();
      HZ_STAMP(6);                              // all W partials have arrived
      if (p.nvls) {
        const char* mc_tile = p.mc_heap + slot;
#pragma unroll
        for (int i = 0; i < kPasses; ++i) {
          const int r0 = threadIdx.x / kVecPerRow + i * kRowsPerPass;
          red[i] = multimem_ld_reduce_bf16x8(mc_tile + (r0 * BLOCK_N + vec * 8) * 2);   // summed inside the NVSwitch
        }
      } else {
        if (threadIdx.x == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");
          tc::mbar_arrive_expect_tx(pull_bar, (uint32_t)(W * kTpPartBytes));
          for (int d = 0; d < W; ++d) {
            const int rr = (me + d) % W;                    // own (local) slot first, peers staggered
            bulk_g2s(smem + rr * kTpPartBytes, p.heap[rr] + slot, kTpPartBytes, pull_bar);
          }
        }
        tc::mbar_wait(pull_bar, 0);
#pragma unroll
        for (int i = 0; i < kPasses; ++i) {
          const int r0 = threadIdx.x / kVecPerRow + i * kRowsPerPass;
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
          for (int rr = 0; rr < W; ++rr) {                  // fixed rank order: deterministic and rank-identical
            float f[8];
            unpack8(ld8(reinterpret_cast<const __nv_bfloat16*>(smem + rr * kTpPartBytes) + r0 * BLOCK_N + vec * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
          }
          const bf16x8 pk = pack8(acc);
          red[i] = *reinterpret_cast<const uint4*>(&pk);
        }
      }
    }
  }
  HZ_STAMP(7);                                  // reduced tile in registers (pull complete)
  if (keeper) {
    // ---- coalesced stores of the finished tile (+ BatchNorm sums, + residual-gradient addend)
    float ssum[8], ssq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ssum[j] = ssq[j] = 0.f;
#pragma unroll
    for (int i = 0; i < kPasses; ++i) {
      const long long off = offs[i];
      if (off < 0) continue;
      const int r0 = threadIdx.x / kVecPerRow + i * kRowsPerPass;
      bf16x8 v = reducing ? *reinterpret_cast<const bf16x8*>(&red[i]) : ld8(staging + r0 * S::kStagingLd + vec * 8);
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
      st8(p.out + off, v);
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
      const int col = threadIdx.x % BLOCK_N, which = threadIdx.x / BLOCK_N;
      const float tot = stat_sm[0][which][col] + stat_sm[1][which][col] + stat_sm[2][which][col] + stat_sm[3][which][col];
      if (nt * BLOCK_N + col < p.ncols) atomicAdd(&p.stats[which * p.ncols + nt * BLOCK_N + col], tot);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  HZ_STAMP(8);                                  // output rows (+ BN sums) written
  if (warp == 2) tc::tmem_dealloc(tmem_d, BLOCK_N);
  if (threadIdx.x == 0) p.epoch[tile] = e;
  HZ_STAMP(9);
#undef HZ_STAMP
}

// ------------------------------------------------------------------------------------------------
// Tensor-parallel classifier head in ONE kernel (reference: TensorParallelLinear + CrossEntropy,
// tensor_parallel_train.py:27-64,203-204): global-avg-pool -> column-parallel FC (this rank's k classes)
// -> all-gather of the logit columns by peer stores -> softmax-CE + accuracy (replicated) -> dlogits ->
// dW_r / db_r inputs (pooled, dlogits) -> dX = sum_r dY_r . W_r by a peer pull-reduce (multimem.ld_reduce when
// the heap is multicast-mapped) -> bf16 dfeat.  One CTA per sample; two cross-rank rendezvous per sample.
// ------------------------------------------------------------------------------------------------
struct TpHeadParams {
  int world, rank, nvls;
  int N, C, HW, K, k_local, n_valid;       // K = k_local * world padded classes
  float loss_scale;
  char* heap[kTpMaxRanks];
  char* mc_heap;
  long long logits_off;          // LL words {f32, epoch}: [2 parities][N][K]
  long long dfeat_off;           // LL words {f32, epoch}: [2 parities][world][N][C]   partial dX of every rank
  long long par_stride_logits, par_stride_dfeat;
  unsigned* epoch;               // local u32 [N]: one call counter per sample CTA
  long long timeout_clk;
};

HZ_DEVINL void ll_store2(void* dst, uint32_t data, uint32_t flag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(dst), "r"(data), "r"(flag) : "memory");
}
HZ_DEVINL uint2 ll_load2(const void* src) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(src) : "memory");
  return v;
}

__global__ void __launch_bounds__(128) tp_head_kernel(const __nv_bfloat16* __restrict__ feat,
                                                      const float* __restrict__ Wl, const float* __restrict__ bl,
                                                      const int64_t* __restrict__ labels, float* __restrict__ pooled,
                                                      float* __restrict__ dl_local, float* __restrict__ logits_out,
                                                      __nv_bfloat16* __restrict__ dfeat, float* __restrict__ loss_out,
                                                      float* __restrict__ correct_out, const TpHeadParams p) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float hsm[];            // pooled[C] | logit[K] | dl[K]
  float* pl = hsm;
  float* lg = hsm + p.C;
  float* dl = lg + p.K;
  const int n = blockIdx.x, W = p.world, me = p.rank, C = p.C, K = p.K, kl = p.k_local;
  const unsigned e = p.epoch[n] + 1u;
  const unsigned par = e & 1u;
  char* my_heap = p.heap[me];
  const float inv_hw = 1.f / (float)p.HW;
  const long long t0 = clock64();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < p.HW; ++q) s += __bfloat162float(feat[((size_t)n * p.HW + q) * C + c]);
    s *= inv_hw;
    pl[c] = s;
    pooled[(size_t)n * C + c] = s;
  }
  __syncthreads();
  // ---- my logit columns, pushed into the logits row of EVERY rank (the reference's K8 all-gather: k floats per peer)
  //      as {value, epoch} words: the receiver polls the word itself, no fence and no separate flag
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const long long lrow = p.logits_off + (long long)par * p.par_stride_logits + ((long long)n * K) * 8;
  for (int k = warp; k < kl; k += nwarp) {
    const int kg = me * kl + k;
    float s = 0.f;
    if (kg < p.n_valid) {
      for (int c = lane; c < C; c += 32) s += pl[c] * Wl[(size_t)k * C + c];
      s = warp_sum(s);
      s += bl ? bl[k] : 0.f;
    } else {
      s = -INFINITY;                          // class padding (10 classes over 8 ranks -> 16): masked
    }
    if (lane < W) ll_store2(p.heap[lane] + lrow + (long long)kg * 8, __float_as_uint(s), e);
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const char* q = my_heap + lrow + (long long)k * 8;
    uint2 v = ll_load2(q);
    while (v.y != e) {
      if (clock64() - t0 > p.timeout_clk) __trap();
      v = ll_load2(q);
    }
    lg[k] = __uint_as_float(v.x);
  }
  __syncthreads();
  // ---- softmax cross-entropy over the gathered row (identical on every rank)
  if (warp == 0) {
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, lg[k]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < K; k += 32) se += __expf(lg[k] - mx);
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    const int lab = (int)labels[n];
    int best = K;
    for (int k = lane; k < K; k += 32) if (lg[k] == mx) best = min(best, k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    for (int k = lane; k < K; k += 32) {
      const float pr = __expf(lg[k] - lse);
      const float d = (pr - (k == lab ? 1.f : 0.f)) * (p.loss_scale / (float)p.N);
      dl[k] = d;
      if (logits_out) logits_out[(size_t)n * K + k] = lg[k];
      if (k >= me * kl && k < (me + 1) * kl) dl_local[(size_t)n * kl + (k - me * kl)] = d;
    }
    if (lane == 0) {
      atomicAdd(loss_out, (lse - lg[lab]) * (p.loss_scale / (float)p.N));
      if (best == lab) atomicAdd(correct_out, 1.f);
    }
  }
  __syncthreads();
  if (dfeat != nullptr) {
    // ---- dX = sum_r dY_r . W_r : my partial (fp32) pushed as {value, epoch} words into slot [me] of every rank, then
    //      the W slots of MY memory are polled and summed in rank order (bit-identical everywhere)
    const long long dbase = p.dfeat_off + (long long)par * p.par_stride_dfeat;
    for (int c4 = threadIdx.x * 4; c4 < C; c4 += blockDim.x * 4) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < kl; ++k)
        if (me * kl + k < p.n_valid) {
          const float d = dl[me * kl + k];
          const float4 w = *reinterpret_cast<const float4*>(Wl + (size_t)k * C + c4);
          s[0] += d * w.x; s[1] += d * w.y; s[2] += d * w.z; s[3] += d * w.w;
        }
      const long long off = dbase + (((long long)me * p.N + n) * C + c4) * 8;
      const uint4 lo = make_uint4(__float_as_uint(s[0]), e, __float_as_uint(s[1]), e);
      const uint4 hi = make_uint4(__float_as_uint(s[2]), e, __float_as_uint(s[3]), e);
      if (W > 1 && p.nvls) {
        ll_mc_store(p.mc_heap + off, lo);
        ll_mc_store(p.mc_heap + off + 16, hi);
      } else {
        for (int d = 0; d < W; ++d) {
          char* dst = p.heap[(me + d) % W] + off;
          ll_store(dst, lo);
          ll_store(dst + 16, hi);
        }
      }
    }
    for (int c4 = threadIdx.x * 4; c4 < C; c4 += blockDim.x * 4) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      for (int rr = 0; rr < W; ++rr) {
        const char* q = my_heap + dbase + (((long long)rr * p.N + n) * C + c4) * 8;
        uint4 lo = ll_load(q), hi = ll_load(q + 16);
        while (lo.y != e || lo.w != e || hi.y != e || hi.w != e) {
          if (clock64() - t0 > p.timeout_clk) __trap();
          lo = ll_load(q); hi = ll_load(q + 16);
        }
        a[0] += __uint_as_float(lo.x); a[1] += __uint_as_float(lo.z);
        a[2] += __uint_as_float(hi.x); a[3] += __uint_as_float(hi.z);
      }
      const float g[4] = {a[0] * inv_hw, a[1] * inv_hw, a[2] * inv_hw, a[3] * inv_hw};
      for (int q = 0; q < p.HW; ++q) {
        __nv_bfloat16* d = dfeat + ((size_t)n * p.HW + q) * C + c4;
        *reinterpret_cast<__nv_bfloat162*>(d) = __floats2bfloat162_rn(g[0], g[1]);
        *reinterpret_cast<__nv_bfloat162*>(d + 2) = __floats2bfloat162_rn(g[2], g[3]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) p.epoch[n] = e;
}

// ------------------------------------------------------------------------------------------------
// Stand-alone bf16 all-reduce (sum) of a small activation tensor over the symmetric heap - the reduction point
// of a tensor-parallel layer whose GEMM cannot take the fused kernel (grid larger than the SM count).  Same
// protocol as the fused epilogue: copy in -> arrival counter -> multimem.ld_reduce / rank-ordered peer pull.
// ------------------------------------------------------------------------------------------------
struct TpArParams {
  int world, rank, nvls, ll;
  char* heap[kTpMaxRanks];
  char* mc_heap;
  long long buf_off, buf_stride;     // bw: bf16 [2 parities][n];  ll: LL words [2 parities][world][n/8] x 32 B
  long long cnt_off;                 // u32 [gridDim.x]
  unsigned* epoch;                   // local u32 [gridDim.x]: one call counter per block
  long long timeout_clk;
};

__global__ void __launch_bounds__(256) tp_allreduce_bf16_kernel(const __nv_bfloat16* __restrict__ in,
                                                                __nv_bfloat16* __restrict__ out, size_t nvec,
                                                                const TpArParams p) {
  pdl_launch();
  pdl_wait();
  const int W = p.world, me = p.rank;
  const unsigned e = p.epoch[blockIdx.x] + 1u;
  const long long boff = p.buf_off + (long long)(e & 1u) * p.buf_stride;
  const size_t per = (nvec + gridDim.x - 1) / gridDim.x;
  const size_t lo = min((size_t)blockIdx.x * per, nvec), hi = min(lo + per, nvec);
  if (p.ll) {
    // latency protocol: {2 bf16, epoch} words pushed into slot [me] of every rank, W local slots polled and summed
    const long long lbase = p.buf_off + (long long)(e & 1u) * p.buf_stride;          // [world][nvec] x 32 B
    const long long t0 = clock64();
    for (size_t v = lo + threadIdx.x; v < hi; v += blockDim.x) {
      const uint4 d = reinterpret_cast<const uint4*>(in)[v];
      const uint4 w0 = make_uint4(d.x, e, d.y, e), w1 = make_uint4(d.z, e, d.w, e);
      const long long off = lbase + ((long long)me * nvec + v) * 32;
      if (p.nvls) {
        ll_mc_store(p.mc_heap + off, w0);
        ll_mc_store(p.mc_heap + off + 16, w1);
      } else {
        for (int dd = 0; dd < W; ++dd) {
          char* dst = p.heap[(me + dd) % W] + off;
          ll_store(dst, w0);
          ll_store(dst + 16, w1);
        }
      }
    }
    for (size_t v = lo + threadIdx.x; v < hi; v += blockDim.x) {
      float a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = 0.f;
      for (int r = 0; r < W; ++r) {
        const char* q = p.heap[me] + lbase + ((long long)r * nvec + v) * 32;
        uint4 w0 = ll_load(q), w1 = ll_load(q + 16);
        while (w0.y != e || w0.w != e || w1.y != e || w1.w != e) {
          if (clock64() - t0 > p.timeout_clk) __trap();
          w0 = ll_load(q); w1 = ll_load(q + 16);
        }
        const uint32_t w[4] = {w0.x, w0.z, w1.x, w1.z};
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(w), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += f[i];
      }
      const bf16x8 pk = pack8(a);
      reinterpret_cast<uint4*>(out)[v] = *reinterpret_cast<const uint4*>(&pk);
    }
    __syncthreads();
    if (threadIdx.x == 0) p.epoch[blockIdx.x] = e;
    return;
  }
  uint4* mine = reinterpret_cast<uint4*>(p.heap[me] + boff);
  for (size_t v = lo + threadIdx.x; v < hi; v += blockDim.x) mine[v] = reinterpret_cast<const uint4*>(in)[v];
  __syncthreads();              // one releasing arrival per CTA orders the block's stores (no per-thread system fence)
  if (threadIdx.x == 0) {
    if (p.nvls) {
      multimem_red_add_u32(reinterpret_cast<unsigned*>(p.mc_heap + p.cnt_off) + blockIdx.x, 1u);
    } else {
      for (int dd = 0; dd < W; ++dd)
        red_release_sys_add_u32(reinterpret_cast<unsigned*>(p.heap[(me + dd) % W] + p.cnt_off) + blockIdx.x, 1u);
    }
    spin_until_ge(reinterpret_cast<unsigned*>(p.heap[me] + p.cnt_off) + blockIdx.x, (unsigned)W * e, p.timeout_clk);
  }
  __syncthreads();
  for (size_t v = lo + threadIdx.x; v < hi; v += blockDim.x) {
    uint4 o;
    if (p.nvls) {
      o = multimem_ld_reduce_bf16x8(p.mc_heap + boff + v * 16);
    } else {
      uint4 w[kTpMaxRanks];
#pragma unroll
      for (int r = 0; r < kTpMaxRanks; ++r)
        if (r < W) w[r] = __ldcg(reinterpret_cast<const uint4*>(p.heap[r] + boff) + v);
      float a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = 0.f;
#pragma unroll
      for (int r = 0; r < kTpMaxRanks; ++r)
        if (r < W) {
          float f[8];
          unpack8(*reinterpret_cast<const bf16x8*>(&w[r]), f);
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] += f[i];
        }
      const bf16x8 pk = pack8(a);
      o = *reinterpret_cast<const uint4*>(&pk);
    }
    reinterpret_cast<uint4*>(out)[v] = o;
  }
  __syncthreads();
  if (threadIdx.x == 0) p.epoch[blockIdx.x] = e;
}

}  // namespace hz

// ================================================================================================
// host
// ================================================================================================
using namespace hz::host;

namespace {
long long* g_tp_dbg = nullptr;
long long tp_timeout_clk() {
  static const long long v = [] {
    const char* e = getenv("HZ_COMM_TIMEOUT_S");
    const double sec = e ? atof(e) : 0.0;
    return (long long)((sec > 0.0 ? sec : 20.0) * 1.9e9);       // clock64 ticks at ~1.9 GHz
  }();
  return v;
}
}  // namespace

extern "C" {

// per-CTA phase stamps (16 x int64 per tile) for the next fused launches; nullptr switches them off
void hz_tp_set_debug(long long* buf) { g_tp_dbg = buf; }

// tiles of the fused op over output [N, H, W, n_out] (per parity class for stride-2 dgrad), or -1
int hz_tp_tiles(int kind, int N, int H, int W_, int Cin, int Cout, int stride) {
  const int Nn = kind == 0 ? Cout : Cin;
  const int Lh = (kind == 1 && stride == 2) ? H / 2 : H, Lw = (kind == 1 && stride == 2) ? W_ / 2 : W_;
  Tile t;
  if (!pick_tile(128, N, Lh, Lw, &t)) return -1;
  return t.tiles * ((Nn + 63) / 64) * ((kind == 1 && stride == 2) ? 4 : 1);
}

// y = [all-reduce | reduce-scatter | nothing] over ranks of conv(x_r, w_r)   (kind 0: forward, stride 1)
// or of dgrad(dy_r, w_r) (kind 1: w MN-major, stride 1 or 2; H, W_ are the INPUT (dx) extents).
// With ag=1 (kind 0) the A operand is image-sharded: x_ptrs[r] is rank r's shard inside its heap.
// heaps[r]: heap base of rank r as mapped here; mc_heap: multicast mapping (nvls) or null.
int hz_tp_conv(int kind, const void* const* x_ptrs, const void* w, void* out, const void* addend, float* stats,
               char* const* heaps, char* mc_heap, long long part_off, long long part_stride, long long cnt_off,
               long long ready_off, unsigned* epoch, int world, int rank, int mode, int nvls, int ll,
               int ag, int N, int H, int W_, int Cin, int Cout, int R, int stride, int pad, cudaStream_t st) {
  const int S_ = R;
  if (kind == 0 && stride != 1) return -23;
  if ((Cin | Cout) & 7) return -20;
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W_ + 2 * pad - S_) / stride + 1;
  const int Ka = kind == 0 ? Cin : Cout;       // channels of the A operand
  const int Nn = kind == 0 ? Cout : Cin;       // output channels
  // output lattice (per class): fwd (Ho,Wo) == (H,W_); dgrad stride 1 (H,W_), stride 2 (H/2,W_/2)
  const int Lh = (kind == 1 && stride == 2) ? H / 2 : H, Lw = (kind == 1 && stride == 2) ? W_ / 2 : W_;
  if (kind == 1 && stride == 2 && (Lh != Ho || Lw != Wo)) return -13;
  Tile t;
  if (!pick_tile(128, N, Lh, Lw, &t)) return -10;
  const int classes = (kind == 1 && stride == 2) ? 4 : 1;
  const int n_tiles = (Nn + 63) / 64;
  const int tiles = t.tiles * n_tiles * classes;
  if (tiles > num_sms()) return -21;           // all CTAs must be co-resident (they spin on peers)
  if (mode == 2 && tiles < world) return -24;  // reduce-scatter: every rank must own a tile (keeps ranks in step)
  const int imgs_per_rank = ag ? N / world : N;
  if (ag && (kind != 0 || N % world || (t.BN > 1 && imgs_per_rank % t.BN))) return -22;
  hz::PeerAMaps am;
  memset(&am, 0, sizeof(am));
  const int Ha = kind == 0 ? H : Ho, Wa = kind == 0 ? W_ : Wo;      // extents of the A tensor
  for (int r = 0; r < world; ++r) {
    if (!ag && r != rank) continue;
    if (!make_map4(&am.m[r], x_ptrs[r], Ka, Wa, Ha, imgs_per_rank, Ka, (long long)Wa * Ka, (long long)Ha * Wa * Ka, 64,
                   t.BW, t.BH, t.BN))
      return -11;
  }
  if (!ag) for (int r = 0; r < world; ++r) if (r != rank) am.m[r] = am.m[rank];
  CUtensorMap bm;
  if (!make_map2(&bm, w, (long long)R * S_ * Cin, Cout, (long long)R * S_ * Cin, 64, 64)) return -12;
  hz::TpParams p;
  memset(&p, 0, sizeof(p));
  p.num_classes = classes;
  for (int c = 0; c < classes; ++c) {
    hz::TapList& tl = p.cls[c];
    tl.n = 0;
    const int ph = c >> 1, pw = c & 1;
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S_; ++s) {
        int dh, dw;
        if (kind == 0) { dh = r - pad; dw = s - pad; }
        else if (stride == 1) { dh = pad - r; dw = pad - s; }
        else {
          if (((ph + pad - r) & 1) || ((pw + pad - s) & 1)) continue;
          dh = (ph + pad - r) / 2; dw = (pw + pad - s) / 2;
        }
        if (!tap_hits(dh, Lh, Ha) || !tap_hits(dw, Lw, Wa)) continue;
        const int i = tl.n++;
        tl.dh[i] = (int8_t)dh; tl.dw[i] = (int8_t)dw; tl.map[i] = 0;
        tl.bk[i] = (r * S_ + s) * Cin;
      }
    p.cls_out_off[c] = (kind == 1 && stride == 2) ? ((long long)ph * W_ + pw) * Nn : 0;
  }
  p.cblocks = (Ka + 63) / 64;
  p.BN = t.BN; p.BH = t.BH; p.BW = t.BW; p.tiles_per_img = t.per_img;
  p.n_images = N;
  const int so = (kind == 1) ? stride : 1;
  const int Hout = kind == 0 ? Ho : H, Wout = kind == 0 ? Wo : W_;
  p.out_n_stride = (long long)Hout * Wout * Nn; p.out_h_stride = (long long)so * Wout * Nn; p.out_w_stride = (long long)so * Nn;
  p.ncols = Nn;
  p.world = world; p.rank = rank; p.mode = world > 1 ? mode : 0; p.nvls = (nvls && mc_heap != nullptr && world > 1) ? 1 : 0;
  p.ll = ll ? 1 : 0;
  p.ag = ag; p.ag_imgs = imgs_per_rank;
  for (int r = 0; r < world; ++r) p.heap[r] = heaps[r];
  p.mc_heap = mc_heap;
  p.part_off = part_off; p.part_stride = part_stride; p.cnt_off = cnt_off; p.ready_off = ready_off;
  p.epoch = epoch;
  p.out = (__nv_bfloat16*)out; p.addend = (const __nv_bfloat16*)addend; p.stats = stats;
  p.timeout_clk = tp_timeout_clk();
  p.dbg = g_tp_dbg;
  using SM = hz::IgemmSmem<64>;
  dim3 grid(t.tiles, n_tiles, classes);
  if (kind == 0) {
    static bool attr = set_smem(hz::igemm_tp_kernel<false>, SM::kTotal);
    (void)attr;
    return hz::launch(hz::igemm_tp_kernel<false>, grid, dim3(128), SM::kTotal, st, am, bm, p) == cudaSuccess ? 0 : -1;
  }
  static bool attr = set_smem(hz::igemm_tp_kernel<true>, SM::kTotal);
  (void)attr;
  return hz::launch(hz::igemm_tp_kernel<true>, grid, dim3(128), SM::kTotal, st, am, bm, p) == cudaSuccess ? 0 : -1;
}

// Tensor-parallel classifier head (see tp_head_kernel).  Wl [k_local, C] fp32 (this rank's class rows), bl [k_local].
// Outputs: pooled [N,C] f32, dl_local [N,k_local] f32 (inputs of the local dW/db kernel), logits [N,K] (optional),
// dfeat [N,HW,C] bf16 (optional), loss / correct accumulators (pre-zeroed).
int hz_tp_head(const void* feat, const float* Wl, const float* bl, const int64_t* labels, float* pooled,
               float* dl_local, float* logits, void* dfeat, float* loss, float* correct, char* const* heaps,
               char* mc_heap, long long logits_off, long long dfeat_off, unsigned* epoch,
               int world, int rank, int nvls, int N, int C, int HW, int k_local, int n_valid,
               float loss_scale, cudaStream_t st) {
  if (C & 3) return -20;
  hz::TpHeadParams p;
  memset(&p, 0, sizeof(p));
  p.world = world; p.rank = rank; p.nvls = (nvls && mc_heap != nullptr && world > 1) ? 1 : 0;
  p.N = N; p.C = C; p.HW = HW; p.K = k_local * world; p.k_local = k_local; p.n_valid = n_valid;
  p.loss_scale = loss_scale;
  for (int r = 0; r < world; ++r) p.heap[r] = heaps[r];
  p.mc_heap = mc_heap;
  p.logits_off = logits_off; p.dfeat_off = dfeat_off;
  p.par_stride_logits = (long long)N * p.K * 8;
  p.par_stride_dfeat = (long long)world * N * C * 8;
  p.epoch = epoch;
  p.timeout_clk = tp_timeout_clk();
  if (N > num_sms()) return -21;
  const size_t smem = sizeof(float) * (C + 2 * p.K);
  return hz::launch(hz::tp_head_kernel, dim3(N), dim3(128), smem, st, (const __nv_bfloat16*)feat, Wl, bl, labels,
                    pooled, dl_local, logits, (__nv_bfloat16*)dfeat, loss, correct, p) == cudaSuccess ? 0 : -1;
}

size_t hz_tp_head_bytes(int N, int C, int K, int world) { return 2 * ((size_t)N * K * 8 + (size_t)world * N * C * 8) + 64; }

// out = sum over ranks of `in` (bf16, n elements, n % 8 == 0); buf: bf16 [2][n] in the heap, cnt: u32 [blocks]
int hz_tp_allreduce_bf16(const void* in, void* out, size_t n, char* const* heaps, char* mc_heap, long long buf_off,
                         long long cnt_off, unsigned* epoch, int world, int rank, int nvls, int ll,
                         int blocks, cudaStream_t st) {
  if (n & 7) return -20;
  hz::TpArParams p;
  memset(&p, 0, sizeof(p));
  p.world = world; p.rank = rank; p.nvls = (nvls && mc_heap != nullptr && world > 1) ? 1 : 0;
  for (int r = 0; r < world; ++r) p.heap[r] = heaps[r];
  p.mc_heap = mc_heap;
  p.ll = ll ? 1 : 0;
  p.buf_off = buf_off; p.buf_stride = ll ? (long long)world * n * 4 : (long long)n * 2; p.cnt_off = cnt_off;
  p.epoch = epoch;
  p.timeout_clk = tp_timeout_clk();
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;
  return hz::launch(hz::tp_allreduce_bf16_kernel, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)in,
                    (__nv_bfloat16*)out, n / 8, p) == cudaSuccess ? 0 : -1;
}

}  // extern "C"

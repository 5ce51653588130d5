// Shared definitions of the implicit-GEMM kernels (conv_gemm.cu, tp_fused.cu): tile constants, tap lists,
// shared-memory plan, TMA tensor-map construction and tile/tap selection on the host.
#pragma once
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "tc05.cuh"

namespace hz {

constexpr int kStages = 6;
constexpr int kTileM = 128;
constexpr int kKBlock = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int kABytes = kTileM * 128;             // 16 KB
constexpr int kMaxTaps = 9;

struct TapList {
  int n;
  int bk[kMaxTaps];          // element offset of this tap along the weight's (r,s,c) axis
  int8_t dh[kMaxTaps], dw[kMaxTaps], map[kMaxTaps];
};

struct AMaps {
  CUtensorMap m[4];
};

template <int BLOCK_N>
struct IgemmSmem {
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kPipeBytes = kStages * kStageBytes;
  static constexpr int kStagingLd = BLOCK_N + 8;                 // bf16 elements, conflict-free 16B rows
  static constexpr int kStagingBytes = kTileM * kStagingLd * 2;
  static constexpr int kBarOff = (kPipeBytes > kStagingBytes ? kPipeBytes : kStagingBytes);
  static constexpr int kTotal = kBarOff + 128 + 1024;            // + barriers + alignment slack
};

}  // namespace hz

namespace hz {
namespace host {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 4-D bf16 tensor map over an NHWC tensor view: dims (C, W, H, N) with element strides (1, sw, sh, sn)
inline bool make_map4(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long sw, long long sh,
               long long sn, int boxC, int boxW, int boxH, int boxN) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)boxW, (cuuint32_t)boxH, (cuuint32_t)boxN};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "[hz conv] cuTensorMapEncodeTiled(4d) failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

inline bool make_map2(CUtensorMap* m, const void* base, long long inner, long long rows, long long row_stride,
               int box_inner, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "[hz conv] cuTensorMapEncodeTiled(2d) failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

struct Tile {
  int BN, BH, BW, per_img, tiles;
};
// rows-per-tile box over an [N, H, W] pixel lattice
inline bool pick_tile(int rows, int N, int H, int W, Tile* t) {
  if (H * W >= rows) {
    if (W > rows || rows % W) return false;
    t->BN = 1; t->BW = W; t->BH = rows / W;
    if (H % t->BH) return false;
    t->per_img = H / t->BH;
    t->tiles = N * t->per_img;
  } else {
    if (rows % (H * W)) return false;
    t->BN = rows / (H * W); t->BH = H; t->BW = W; t->per_img = 1;
    t->tiles = (N + t->BN - 1) / t->BN;
  }
  return t->BW <= 256 && t->BH <= 256 && t->BN <= 256;
}

// does offset d hit any valid index: exists i in [0,n_out) with 0 <= i + d < n_in
inline bool tap_hits(int d, int n_out, int n_in) { return d < n_in && d + n_out - 1 >= 0 && d > -n_out - n_in; }

// x maps for a conv input: plain view (stride 1) or 4 parity views (stride 2)
inline bool make_x_maps(hz::AMaps* am, const void* x, int N, int H, int W, int C, int stride, const Tile& t) {
  if (stride == 1) {
    if (!make_map4(&am->m[0], x, C, W, H, N, C, (long long)W * C, (long long)H * W * C, 64, t.BW, t.BH, t.BN))
      return false;
    for (int i = 1; i < 4; ++i) am->m[i] = am->m[0];
    return true;
  }
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const char* base = (const char*)x + ((long long)ph * W + pw) * C * 2;
      if (!make_map4(&am->m[ph * 2 + pw], base, C, W / 2, H / 2, N, 2LL * C, 2LL * W * C, (long long)H * W * C,
                     64, t.BW, t.BH, t.BN))
        return false;
    }
  return true;
}

// taps of a conv reading its input: coordinate offsets in (parity-)view space
inline void input_taps(hz::TapList* tl, int R, int S, int stride, int pad, int Ho, int Wo, int H, int W, int Cin,
                bool bk_is_tap_index) {
  tl->n = 0;
  const int Hv = stride == 1 ? H : H / 2, Wv = stride == 1 ? W : W / 2;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      int dh, dw, map = 0;
      if (stride == 1) { dh = r - pad; dw = s - pad; }
      else {
        const int th = r - pad, tw = s - pad;
        const int ph = th & 1, pw = tw & 1;
        dh = (th - ph) / 2; dw = (tw - pw) / 2;
        map = ph * 2 + pw;
      }
      if (!tap_hits(dh, Ho, Hv) || !tap_hits(dw, Wo, Wv)) continue;
      const int i = tl->n++;
      tl->dh[i] = (int8_t)dh; tl->dw[i] = (int8_t)dw; tl->map[i] = (int8_t)map;
      tl->bk[i] = bk_is_tap_index ? (r * S + s) : (r * S + s) * Cin;
    }
}


// SM count of the current device (co-residency bound of kernels whose CTAs wait for each other / for peers)
inline int num_sms() {
  static const int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

template <typename K>
inline bool set_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
}


}  // namespace host
}  // namespace hz

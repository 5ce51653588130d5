// Host-only checks of the implicit-GEMM addressing logic (no kernel is launched, no GPU needed):
//   tap_hits / input_taps  — which filter taps survive, and that (view, dh, dw) addresses the pixel the
//                            convolution definition says, for stride 1 and the stride-2 parity views;
//   pick_tile              — 128-row tiles tile the [N, H, W] output lattice exactly.
// Built and run by tests/test_cpu_units.py::test_host_tiling_logic (nvcc, host code only).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../igemm_common.cuh"

using namespace hz::host;

static int fails = 0;
#define CHECK(cond, ...)                                                   \
  do {                                                                     \
    if (!(cond)) { ++fails; fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } \
  } while (0)

static bool brute_hits(int d, int n_out, int n_in) {
  for (int i = 0; i < n_out; ++i)
    if (i + d >= 0 && i + d < n_in) return true;
  return false;
}

int main() {
  // ---- tap_hits == brute force
  for (int n_out = 1; n_out <= 9; ++n_out)
    for (int n_in = 1; n_in <= 9; ++n_in)
      for (int d = -12; d <= 12; ++d)
        CHECK(tap_hits(d, n_out, n_in) == brute_hits(d, n_out, n_in), "tap_hits d=%d n_out=%d n_in=%d", d, n_out, n_in);

  // ---- input_taps: every distinct conv geometry of ResNet-18 at 32x32 (+ a few others)
  struct G { int H, R, stride, pad; };
  const G geos[] = {{8, 3, 1, 1}, {8, 3, 2, 1}, {8, 1, 2, 0}, {4, 3, 1, 1}, {4, 3, 2, 1}, {4, 1, 2, 0},
                    {2, 3, 1, 1}, {2, 3, 2, 1}, {2, 1, 2, 0}, {1, 3, 1, 1}, {16, 3, 1, 1}, {16, 3, 2, 1}, {32, 1, 1, 0}};
  for (const G& g : geos) {
    const int H = g.H, W = g.H, R = g.R, S = g.R, st = g.stride, pad = g.pad, Cin = 64;
    const int Ho = (H + 2 * pad - R) / st + 1, Wo = (W + 2 * pad - S) / st + 1;
    hz::TapList tl;
    input_taps(&tl, R, S, st, pad, Ho, Wo, H, W, Cin, false);
    std::vector<int> kept(R * S, 0);
    for (int i = 0; i < tl.n; ++i) {
      CHECK(tl.bk[i] % Cin == 0, "bk not a tap offset");
      const int tap = tl.bk[i] / Cin, r = tap / S, s = tap % S;
      kept[tap] = 1;
      // the (view, dh, dw) triple must address input pixel (ho*st + r - pad, wo*st + s - pad) for every output pixel
      for (int ho = 0; ho < Ho; ++ho)
        for (int wo = 0; wo < Wo; ++wo) {
          const int h_true = ho * st + r - pad, w_true = wo * st + s - pad;
          int h_map, w_map;
          if (st == 1) { h_map = ho + tl.dh[i]; w_map = wo + tl.dw[i]; CHECK(tl.map[i] == 0, "stride-1 map"); }
          else {
            const int ph = tl.map[i] >> 1, pw = tl.map[i] & 1;
            h_map = 2 * (ho + tl.dh[i]) + ph; w_map = 2 * (wo + tl.dw[i]) + pw;
          }
          CHECK(h_map == h_true && w_map == w_true, "H=%d R=%d st=%d tap(%d,%d) out(%d,%d): mapped (%d,%d) != (%d,%d)", H, R,
                st, r, s, ho, wo, h_map, w_map, h_true, w_true);
        }
    }
    // a tap is dropped iff no output pixel reads a valid input pixel through it (TMA zero-fills the rest)
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S; ++s) {
        bool any = false;
        for (int ho = 0; ho < Ho && !any; ++ho)
          for (int wo = 0; wo < Wo && !any; ++wo) {
            const int h = ho * st + r - pad, w = wo * st + s - pad;
            any = h >= 0 && h < H && w >= 0 && w < W;
          }
        // separable test in the kernel: rows and columns are checked independently (never drops a live tap)
        bool any_h = false, any_w = false;
        for (int ho = 0; ho < Ho; ++ho) any_h |= (ho * st + r - pad >= 0 && ho * st + r - pad < H);
        for (int wo = 0; wo < Wo; ++wo) any_w |= (wo * st + s - pad >= 0 && wo * st + s - pad < W);
        CHECK(kept[r * S + s] == (any_h && any_w), "H=%d R=%d st=%d tap(%d,%d) kept=%d expected=%d", H, R, st, r, s,
              kept[r * S + s], (int)(any_h && any_w));
        CHECK(!any || kept[r * S + s], "live tap dropped");
      }
  }
  // layer4 at 32x32 input: 1x1 maps keep only the centre tap
  {
    hz::TapList tl;
    input_taps(&tl, 3, 3, 1, 1, 1, 1, 1, 1, 512, true);
    CHECK(tl.n == 1 && tl.bk[0] == 4, "centre tap only, got n=%d", tl.n);
  }

  // ---- pick_tile: exact tiling of the output lattice with 128-row boxes
  const int lattices[][3] = {{64, 16, 16}, {64, 8, 8}, {64, 4, 4}, {64, 2, 2}, {64, 1, 1}, {16, 8, 8}, {8, 16, 16},
                             {4096, 8, 8}, {100, 1, 1}, {3, 2, 2}};
  for (auto& l : lattices) {
    Tile t;
    const int N = l[0], H = l[1], W = l[2];
    if (!pick_tile(128, N, H, W, &t)) { CHECK(false, "pick_tile(128,%d,%d,%d) refused", N, H, W); continue; }
    CHECK(t.BN * t.BH * t.BW == 128, "box %dx%dx%d", t.BN, t.BH, t.BW);
    if (t.BN == 1) CHECK(t.per_img * t.BH == H && t.BW == W && t.tiles == N * t.per_img, "intra-image tiling");
    else CHECK(t.BH == H && t.BW == W && t.tiles == (N + t.BN - 1) / t.BN, "multi-image tiling");
  }
  {
    Tile t;
    CHECK(!pick_tile(128, 64, 3, 3, &t), "9-pixel maps cannot form 128-row boxes");
    CHECK(!pick_tile(128, 64, 24, 24, &t), "W=24 does not divide 128");
  }
  if (fails) { fprintf(stderr, "%d check(s) failed\n", fails); return 1; }
  printf("host tiling logic ok\n");
  return 0;
}

// Host-only check of the depthwise-convolution kernels' per-thread logic (csrc/dw_core.cuh): the thread layout
// (channel vector x row lane, idle left-over threads), the grid-stride row walk and the forward / input-gradient /
// weight-gradient index math are run on the CPU — one loop iteration per (block, thread) exactly as depthwise.cu
// schedules them — and compared with the convolution definition.  No kernel is launched, no GPU needed.
// Built and run by tests/test_cpu_units.py::test_host_depthwise_logic (nvcc, host code only).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../dw_core.cuh"

using namespace hz::dw;

static int fails = 0;
#define CHECK(cond, ...)                                                   \
  do {                                                                     \
    if (!(cond)) { if (++fails < 20) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } \
  } while (0)

static float frand(unsigned& s) {
  s = s * 1664525u + 1013904223u;
  return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static std::vector<__nv_bfloat16> rnd_bf16(size_t n, unsigned seed, float scale = 1.f) {
  std::vector<__nv_bfloat16> v(n);
  unsigned s = seed;
  for (auto& e : v) e = __float2bfloat16_rn(frand(s) * scale);
  return v;
}
static inline double f(const __nv_bfloat16& b) { return (double)__bfloat162float(b); }

struct Case { int N, H, W, C, stride, grid; };

static void run(const Case& cs) {
  const Geo g = make_geo(cs.N, cs.H, cs.W, cs.C, cs.stride);
  const int C = g.C, M = g.N * g.Ho * g.Wo, Q = g.N * g.H * g.W;
  // aligned (16-byte) buffers: std::vector<bf16> of a multiple of 8 elements from operator new is 16-byte aligned
  auto x = rnd_bf16((size_t)Q * C, 1), w = rnd_bf16((size_t)C * 9, 2, 0.5f), dy = rnd_bf16((size_t)M * C, 3);
  std::vector<__nv_bfloat16> y((size_t)M * C), dx((size_t)Q * C);
  std::vector<int> hits_y((size_t)M * (C / 8), 0), hits_dx((size_t)Q * (C / 8), 0);
  std::vector<float> dwt((size_t)C * 9, 0.f), s1(C, 0.f), s2(C, 0.f);

  // ---- the kernels, thread by thread
  for (int b = 0; b < cs.grid; ++b)
    for (int tid = 0; tid < 256; ++tid) {
      const Lane l = make_lane(tid, C);
      CHECK(l.rlanes >= 1 && l.cv < l.nvec, "lane layout C=%d tid=%d", C, tid);
      if (!l.active) continue;
      float wr[9][8];
      load_taps(w.data(), l.cv, wr);
      for (int p = b * l.rlanes + l.rl; p < M; p += cs.grid * l.rlanes) {      // forward
        float acc[8];
        fwd_pixel(g, x.data(), p, l.cv, wr, acc);
        store8(y.data() + (size_t)p * C + l.cv * 8, acc);
        for (int i = 0; i < 8; ++i) { s1[l.cv * 8 + i] += acc[i]; s2[l.cv * 8 + i] += acc[i] * acc[i]; }
        ++hits_y[(size_t)p * l.nvec + l.cv];
      }
      for (int q = b * l.rlanes + l.rl; q < Q; q += cs.grid * l.rlanes) {      // input gradient
        float acc[8];
        dgrad_pixel(g, dy.data(), q, l.cv, wr, acc);
        store8(dx.data() + (size_t)q * C + l.cv * 8, acc);
        ++hits_dx[(size_t)q * l.nvec + l.cv];
      }
      float acc9[9][8];
      for (int t = 0; t < 9; ++t) for (int i = 0; i < 8; ++i) acc9[t][i] = 0.f;
      for (int p = b * l.rlanes + l.rl; p < M; p += cs.grid * l.rlanes) wgrad_pixel(g, dy.data(), x.data(), p, l.cv, acc9);
      for (int t = 0; t < 9; ++t) for (int i = 0; i < 8; ++i) dwt[(size_t)(l.cv * 8 + i) * 9 + t] += acc9[t][i];
    }
  for (int h : hits_y) CHECK(h == 1, "output vector written %d times (C=%d grid=%d)", h, C, cs.grid);
  for (int h : hits_dx) CHECK(h == 1, "dx vector written %d times (C=%d grid=%d)", h, C, cs.grid);

  // ---- the definition
  std::vector<double> dx_ref((size_t)Q * C, 0.0), dw_ref((size_t)C * 9, 0.0), s1_ref(C, 0.0), s2_ref(C, 0.0);
  for (int n = 0; n < g.N; ++n)
    for (int ho = 0; ho < g.Ho; ++ho)
      for (int wo = 0; wo < g.Wo; ++wo)
        for (int c = 0; c < C; ++c) {
          const size_t po = (((size_t)n * g.Ho + ho) * g.Wo + wo) * C + c;
          double acc = 0.0;
          for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
              const int h = ho * g.stride - 1 + r, ww = wo * g.stride - 1 + s;
              if (h < 0 || h >= g.H || ww < 0 || ww >= g.W) continue;
              const size_t pi = (((size_t)n * g.H + h) * g.W + ww) * C + c;
              acc += f(x[pi]) * f(w[(size_t)c * 9 + r * 3 + s]);
              dx_ref[pi] += f(dy[po]) * f(w[(size_t)c * 9 + r * 3 + s]);
              dw_ref[(size_t)c * 9 + r * 3 + s] += f(dy[po]) * f(x[pi]);
            }
          const double got = f(y[po]);
          CHECK(std::fabs(got - acc) <= 0.02 * std::fabs(acc) + 0.02, "fwd n=%d ho=%d wo=%d c=%d: %g vs %g", n, ho, wo, c, got, acc);
          s1_ref[c] += got; s2_ref[c] += got * got;
        }
  for (size_t i = 0; i < dx_ref.size(); ++i)
    CHECK(std::fabs(f(dx[i]) - dx_ref[i]) <= 0.02 * std::fabs(dx_ref[i]) + 0.02, "dgrad elem %zu: %g vs %g", i, f(dx[i]), dx_ref[i]);
  for (size_t i = 0; i < dw_ref.size(); ++i)
    CHECK(std::fabs(dwt[i] - dw_ref[i]) <= 1e-3 * std::fabs(dw_ref[i]) + 1e-2, "wgrad c=%zu tap=%zu: %g vs %g", i / 9, i % 9, dwt[i], dw_ref[i]);
  for (int c = 0; c < C; ++c) {
    CHECK(std::fabs(s1[c] - s1_ref[c]) <= 1e-3 * std::fabs(s1_ref[c]) + 1e-2, "sum y c=%d", c);
    CHECK(std::fabs(s2[c] - s2_ref[c]) <= 1e-3 * std::fabs(s2_ref[c]) + 1e-2, "sum y^2 c=%d", c);
  }
}

int main() {
  const Case cases[] = {
      {2, 8, 8, 16, 1, 1},   {2, 8, 8, 24, 2, 3},  {3, 5, 7, 144, 1, 2}, {2, 7, 5, 40, 2, 5},  {4, 1, 1, 960, 1, 2},
      {2, 2, 2, 64, 2, 1},   {1, 16, 16, 8, 1, 7}, {2, 4, 4, 2048, 1, 3}, {3, 3, 3, 96, 2, 4}, {2, 16, 16, 32, 2, 2},
      {5, 2, 2, 576, 2, 1},  {2, 6, 6, 192, 1, 64},
  };
  for (const Case& c : cases) run(c);
  // random geometries: any batch, odd / even extents, every channel-vector count class (power of two or not), both strides
  unsigned s = 12345u;
  auto rnd = [&](int lo, int hi) { s = s * 1664525u + 1013904223u; return lo + (int)((s >> 8) % (unsigned)(hi - lo + 1)); };
  int n_random = 0;
  for (int i = 0; i < 60; ++i) {
    Case c;
    c.N = rnd(1, 5); c.H = rnd(1, 9); c.W = rnd(1, 9); c.C = 8 * rnd(1, 40); c.stride = rnd(1, 2); c.grid = rnd(1, 9);
    run(c);
    ++n_random;
  }
  if (fails) { fprintf(stderr, "%d check(s) failed\n", fails); return 1; }
  printf("host depthwise logic ok (%zu cases + %d random)\n", sizeof(cases) / sizeof(cases[0]), n_random);
  return 0;
}

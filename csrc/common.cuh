// Shared device helpers for the horizonml_b200 sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>

#define HZ_DEVINL __device__ __forceinline__

namespace hz {

constexpr int kWarp = 32;

HZ_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

HZ_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 8 bf16 values in one 16-byte vector
struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

HZ_DEVINL void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

HZ_DEVINL bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

HZ_DEVINL bf16x8 ld8(const __nv_bfloat16* p) { return *reinterpret_cast<const bf16x8*>(p); }
HZ_DEVINL void st8(__nv_bfloat16* p, const bf16x8& v) { *reinterpret_cast<bf16x8*>(p) = v; }

// Programmatic dependent launch (PDL): `pdl_launch()` lets the next kernel in the stream become resident and
// run its prologue while this one drains; `pdl_wait()` blocks until every prerequisite grid has completed and
// flushed its memory.  Both are no-ops when the kernel was not launched with the PDL attribute.
HZ_DEVINL void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
HZ_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- thread-block cluster / distributed shared memory
HZ_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
HZ_DEVINL void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// split form: arrive as soon as this CTA is done with its peers' memory, wait only where the hazard really is
HZ_DEVINL void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
HZ_DEVINL void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// shared::cluster address of `local_smem_addr` inside CTA `rank` of this cluster
HZ_DEVINL uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
HZ_DEVINL float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// same load without the compiler memory barrier: lets a batch of independent remote loads be issued back to back
// (ordering against the surrounding barrier.cluster is kept by `volatile`)
HZ_DEVINL float4 ld_dsmem_f4_nb(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

HZ_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

}  // namespace hz

// ---- host: launch with the PDL attribute (HZ_PDL=0 disables) ------------------------------------
#ifdef __CUDACC__
#include <cstdlib>
namespace hz {
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("HZ_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  unsigned cluster_z, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_z > 1) {          // thread-block cluster along z: the split-K CTAs of one output tile
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 1;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = cluster_z;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                          Args&&... args) {
  return launch_cluster(kernel, grid, block, smem, st, 1u, std::forward<Args>(args)...);
}
// Zero-fill as a *kernel* (not cudaMemsetAsync): a memset node sitting between two PDL-launched kernels is not
// a grid, so `griddepcontrol.wait` in the consumer does not order against it (observed: BN sums cleared after
// the conv epilogue had already accumulated into them inside captured graphs).
static __global__ void zero_f32_kernel(float* __restrict__ p, size_t n) {
  pdl_launch();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}
inline void zero_f32(float* p, size_t n, cudaStream_t st) {
  size_t g = (n + 255) / 256;
  if (g > 592) g = 592;
  if (g < 1) g = 1;
  launch(zero_f32_kernel, dim3((unsigned)g), dim3(256), 0, st, p, n);
}
}  // namespace hz
#endif

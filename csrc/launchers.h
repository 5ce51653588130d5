// C interface between the CUDA translation units (no torch headers) and bindings.cpp.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// ---- elementwise.cu
int hz_channel_ok(int C);
void hz_channel_sums(const void* y, float* sums, int M, int C, cudaStream_t st);
void hz_bn_act_fwd(const void* y, const float* sums, const float* gamma, const float* beta,
                   const void* residual, void* out, float* mean, float* invstd, float* rmean, float* rvar,
                   int M, int C, float eps, float momentum, int relu, int training, cudaStream_t st);
void hz_bn_act_bwd(const void* dout, const void* outp, const void* yraw, const float* mean,
                   const float* invstd, const float* gamma, float* sums_scratch, void* dy, void* dres,
                   float* dgamma, float* dbeta, int acc_gamma, int acc_beta, int M, int C, int relu,
                   int scratch_is_zero, cudaStream_t st);
int hz_bn_act_bwd_res(const void* dout, const void* outp, const void* yraw, const float* mean, const float* invstd,
                      const float* gamma, float* sums_scratch, void* dy, void* dres, float* dgamma, float* dbeta,
                      int acc_gamma, int acc_beta, int M, int C, int relu, int scratch_is_zero, const void* res_yraw,
                      const float* res_mean, const float* res_invstd, float* res_sums, int res_sums_is_zero,
                      cudaStream_t st);
void hz_maxpool_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, cudaStream_t st);
void hz_maxpool_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, cudaStream_t st);
int hz_maxpool_bwd_bn(const void* dy, const void* idx, void* dx, const void* bn_out, const void* bn_yraw, const float* mean,
                      const float* invstd, float* sums, int sums_is_zero, int relu, int N, int H, int W, int C,
                      cudaStream_t st);
void hz_u8_normalize(const void* in, void* out, size_t n, float mean, float std, cudaStream_t st);
void hz_im2col_small(const void* x, void* A, int N, int H, int W, int Cin, int R, int S, int stride, int pad,
                     int Ho, int Wo, int Kp, cudaStream_t st);
void hz_pad_rows(const void* in, void* out, int rows, int K, int Kp, cudaStream_t st);
int hz_stem_pack(const void* x, void* A, const void* w, void* wp, int N, int H, int W, int Cin, int R, int stride,
                 int pad, int Ho, int Wo, int Kp, int rows_w, cudaStream_t st);
void hz_head_fwd_bwd(const void* feat, const float* W, const float* bias, const int64_t* labels, float* pooled,
                     float* dlogits, float* logits, void* dfeat, float* loss, float* correct, float* dW,
                     float* db, int N, int C, int HW, int K, int n_valid, float loss_scale, int accumulate,
                     int out_is_zero, cudaStream_t st);
void hz_adam(float* p, float* g, float* m, float* v, void* shadow, float* step, float* prev, float* diff_out,
             int zero_grad, size_t n, float lr, float b1, float b2, float eps, float gscale, const int* live,
             size_t n_live_blocks, int bump, int max_ctas, cudaStream_t st);
void hz_grad_diff(const float* g, float* prev, float* out, size_t n, cudaStream_t st);
void hz_stats_update(float* stats, float* has_prev, const float* loss, const float* correct, float batch,
                     float* diff_sq, cudaStream_t st);

// ---- depthwise.cu (3x3 depthwise convolution, pad 1, stride 1|2; x/y NHWC bf16, w bf16 [C][3][3], dw fp32 [C][3][3])
int hz_dwconv_ok(int N, int H, int W, int C, int stride);
int hz_dwconv_fwd(const void* x, const void* w, void* y, float* stats, int stats_is_zero, int N, int H, int W, int C,
                  int stride, cudaStream_t st);
int hz_dwconv_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int C, int stride, cudaStream_t st);
int hz_dwconv_wgrad(const void* dy, const void* x, float* dwt, int N, int H, int W, int C, int stride, int accumulate,
                    int prezeroed, cudaStream_t st);

// ---- conv_gemm.cu (tcgen05 implicit GEMM)
int hz_conv_supported(int N, int H, int W, int Cin, int Cout, int R, int stride, int pad);
int hz_conv_shape_ok(int N, int H, int W, int Cin, int Cout, int R, int stride, int pad);
void hz_cluster_capacity(int out[4]);
void hz_conv_set_debug(long long* buf);
int hz_conv_set_persist(int mode);      // persistent (throughput) conv kernel: -1 auto by grid size, 0 never, 1 always
// optional fused BatchNorm(batch statistics) + residual + ReLU epilogue of the forward convolution
struct HzBnFuse {
  const float* gamma; const float* beta;
  float* mean; float* invstd;          // [C] saved for backward
  float* rmean; float* rvar;           // running statistics, updated in place (may be null)
  const void* residual;                // bf16, layout of the output, or null
  void* out;                           // bf16 activation
  unsigned* counter;                   // device-wide barrier counter, zero before the launch
  float eps, momentum;
  int relu;
};
// rc -20: grid larger than the SM count (barrier needs co-residency) — caller runs the unfused pair instead
int hz_conv_fwd(const void* x, const void* w, void* y, float* stats, int stats_is_zero, int N, int H, int W,
                int Cin, int Cout, int R, int stride, int pad, int weights_stable, const HzBnFuse* bn,
                cudaStream_t st);
int hz_conv_dgrad(const void* dy, const void* w, void* dx, const void* addend, int N, int H, int W, int Cin, int Cout, int R,
                  int stride, int pad, int weights_stable, cudaStream_t st);
// BatchNorm-backward sums (sum g, sum g*xhat; g = dx * [out > 0]) of the layer that produced the conv's input, taken in
// the dgrad epilogue: `out` (null = no ReLU) / `yraw` are that layer's BN output / raw conv output (layout of dx),
// mean / invstd its saved statistics [Cin], sums [2*Cin] fp32 (sums_is_zero: already cleared)
struct HzBnBwd {
  const void* out; const void* yraw;
  const float* mean; const float* invstd;
  float* sums;
  int sums_is_zero;
  int cap6;                      // the producing layer's activation is ReLU6 (mask 0 < out < 6) instead of ReLU
};
int hz_conv_dgrad_bnbwd(const void* dy, const void* w, void* dx, const void* addend, int N, int H, int W, int Cin, int Cout,
                        int R, int stride, int pad, int weights_stable, const struct HzBnBwd* bnb, cudaStream_t st);
int hz_dwconv_dgrad_bnbwd(const void* dy, const void* w, void* dx, int N, int H, int W, int C, int stride,
                          const struct HzBnBwd* bnb, cudaStream_t st);      // depthwise.cu
int hz_conv_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Cout, int R,
                  int stride, int pad, int accumulate, int prezeroed, long long ld_out, int n_valid,
                  cudaStream_t st);

// ---- comm.cu (peer-memory all-reduce)
struct HzComm;
struct HzComm* hz_comm_create(int rank, int world, int device, size_t max_wire_bytes, int max_blocks);
struct HzComm* hz_comm_create2(int rank, int world, int device, size_t max_wire_bytes, int max_blocks,
                               size_t heap_bytes);
char* hz_comm_heap_base(struct HzComm* c, int r);
size_t hz_comm_heap_bytes(struct HzComm* c);
int hz_comm_export(struct HzComm* c, void* handle64);
int hz_comm_import(struct HzComm* c, const void* handles);
int hz_comm_link_local(struct HzComm** comms, int world);
size_t hz_comm_symm_bytes(struct HzComm* c);
void hz_comm_set_multicast(struct HzComm* c, void* mc_ptr, void* local_ptr, size_t bytes);
int hz_comm_blocks_for(struct HzComm* c, size_t n, int algo, int wire_bf16);
void hz_comm_set_block_cap(struct HzComm* c, int cap);
int hz_comm_allreduce(struct HzComm* c, float* grad, size_t n, int algo, int wire_bf16, float scale,
                      const int* live, cudaStream_t st);
int hz_comm_allreduce_adam(struct HzComm* c, float* grad, size_t n, int algo, int wire_bf16, float scale,
                           const int* live, float* p, float* m, float* v, void* shadow, float* prev, float* diff_out,
                           float* step, float lr, float b1, float b2, float eps, int bump, cudaStream_t st);
size_t hz_comm_zero1_shard(size_t n, int world);
int hz_comm_zero1_step(struct HzComm* c, float* grad, size_t n, float scale, const int* live, float* p, float* m, float* v,
                       void* shadow, float* prev, float* diff_out, float* step, float lr, float b1, float b2, float eps,
                       int bump, cudaStream_t st);
int hz_comm_barrier(struct HzComm* c, long long* stamps, cudaStream_t st);
int hz_comm_error(struct HzComm* c);
void hz_comm_destroy(struct HzComm* c);

// ---- tp_fused.cu (GEMM fused with its collective over peer memory; TP head; small bf16 all-reduce)
void hz_tp_set_debug(long long* buf);
int hz_tp_tiles(int kind, int N, int H, int W_, int Cin, int Cout, int stride);
int hz_tp_conv(int kind, const void* const* x_ptrs, const void* w, void* out, const void* addend, float* stats,
               char* const* heaps, char* mc_heap, long long part_off, long long part_stride, long long cnt_off,
               long long ready_off, unsigned* epoch, int world, int rank, int mode, int nvls, int ll,
               int ag, int N, int H, int W_, int Cin, int Cout, int R, int stride, int pad, cudaStream_t st);
int hz_tp_head(const void* feat, const float* Wl, const float* bl, const int64_t* labels, float* pooled,
               float* dl_local, float* logits, void* dfeat, float* loss, float* correct, char* const* heaps,
               char* mc_heap, long long logits_off, long long dfeat_off, unsigned* epoch,
               int world, int rank, int nvls, int N, int C, int HW, int k_local, int n_valid,
               float loss_scale, cudaStream_t st);
size_t hz_tp_head_bytes(int N, int C, int K, int world);
int hz_tp_allreduce_bf16(const void* in, void* out, size_t n, char* const* heaps, char* mc_heap, long long buf_off,
                         long long cnt_off, unsigned* epoch, int world, int rank, int nvls, int ll,
                         int blocks, cudaStream_t st);
void hz_head_wgrad(const float* pooled, const float* dlogits, float* dW, float* db, int N, int C, int K,
                   int accumulate, cudaStream_t st);

#ifdef __cplusplus
}
#endif

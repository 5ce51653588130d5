// Python bindings (torch tensors in, raw pointers + current CUDA stream out to the launchers).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <pybind11/stl.h>
#include <torch/extension.h>

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "launchers.h"

namespace py = pybind11;
using at::Tensor;

namespace {

cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

inline const void* cptr(const Tensor& t) { return t.data_ptr(); }
inline void* mptr(Tensor& t) { return t.data_ptr(); }
inline float* fptr(const c10::optional<Tensor>& t) { return t.has_value() && t->defined() ? t->data_ptr<float>() : nullptr; }

// physical NHWC dims of a logical NCHW channels_last tensor
struct Dims { int N, C, H, W; };
Dims dims_of(const Tensor& t) {
  TORCH_CHECK(t.dim() == 4, "expected 4-D tensor");
  return {(int)t.size(0), (int)t.size(1), (int)t.size(2), (int)t.size(3)};
}
void check_cl(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kBFloat16, name, " must be bf16");
  TORCH_CHECK(t.is_contiguous(at::MemoryFormat::ChannelsLast) || t.is_contiguous(), name,
              " must be channels_last contiguous");
  if (!(t.size(2) == 1 && t.size(3) == 1) && t.size(1) != 1)
    TORCH_CHECK(t.is_contiguous(at::MemoryFormat::ChannelsLast), name, " must be channels_last");
}
Tensor empty_cl(const Tensor& like, int N, int C, int H, int W) {
  return at::empty({N, C, H, W}, like.options().memory_format(at::MemoryFormat::ChannelsLast));
}

int channel_ok(int64_t C) { return hz_channel_ok((int)C); }

Tensor channel_sums(const Tensor& y) {
  check_cl(y, "y");
  c10::cuda::CUDAGuard g(y.device());
  auto d = dims_of(y);
  Tensor sums = at::empty({2, d.C}, y.options().dtype(at::kFloat));
  hz_channel_sums(cptr(y), sums.data_ptr<float>(), d.N * d.H * d.W, d.C, cur_stream());
  return sums;
}

std::vector<Tensor> bn_act_fwd(const Tensor& y, const Tensor& sums, const Tensor& gamma, const Tensor& beta,
                               const c10::optional<Tensor>& rmean, const c10::optional<Tensor>& rvar,
                               double momentum, double eps, const c10::optional<Tensor>& residual, int64_t relu,
                               bool training) {   // relu: 0 none, 1 ReLU, 2 ReLU6
  check_cl(y, "y");
  c10::cuda::CUDAGuard g(y.device());
  auto d = dims_of(y);
  Tensor out = at::empty_like(y);
  Tensor mean = at::empty({d.C}, y.options().dtype(at::kFloat));
  Tensor invstd = at::empty({d.C}, y.options().dtype(at::kFloat));
  const void* res = nullptr;
  if (residual.has_value() && residual->defined()) { check_cl(*residual, "residual"); res = residual->data_ptr(); }
  hz_bn_act_fwd(cptr(y), sums.data_ptr<float>(), gamma.data_ptr<float>(), beta.data_ptr<float>(), res,
                out.data_ptr(), mean.data_ptr<float>(), invstd.data_ptr<float>(), fptr(rmean), fptr(rvar),
                d.N * d.H * d.W, d.C, (float)eps, (float)momentum, (int)relu, training ? 1 : 0, cur_stream());
  return {out, mean, invstd};
}

// writes dgamma/dbeta straight into their gradient slots
std::vector<Tensor> bn_act_bwd(const Tensor& dout, const Tensor& out, const Tensor& yraw, const Tensor& mean,
                               const Tensor& invstd, const Tensor& gamma, int64_t relu, bool has_res,
                               Tensor dgamma, Tensor dbeta, bool acc_gamma, bool acc_beta,
                               c10::optional<Tensor> zeroed_scratch, bool sums_ready) {
  // sums_ready: `zeroed_scratch` already holds the final (sum g, sum g*xhat) — left there by conv_dgrad_bnbwd
  check_cl(dout, "dout"); check_cl(yraw, "yraw");
  c10::cuda::CUDAGuard g(dout.device());
  auto d = dims_of(yraw);
  Tensor dy = at::empty_like(yraw);
  Tensor dres;
  if (has_res) dres = at::empty_like(yraw);
  const bool pre = zeroed_scratch.has_value() && zeroed_scratch->defined();
  Tensor scratch = pre ? *zeroed_scratch : at::empty({2, d.C}, yraw.options().dtype(at::kFloat));
  hz_bn_act_bwd(cptr(dout), cptr(out), cptr(yraw), mean.data_ptr<float>(), invstd.data_ptr<float>(),
                gamma.data_ptr<float>(), scratch.data_ptr<float>(), dy.data_ptr(),
                has_res ? dres.data_ptr() : nullptr, dgamma.data_ptr<float>(), dbeta.data_ptr<float>(),
                acc_gamma ? 1 : 0, acc_beta ? 1 : 0, d.N * d.H * d.W, d.C, (int)relu,
                (sums_ready && pre) ? 3 : (pre ? (scratch.numel() >= 2 * d.C + 32 ? 2 : 1) : 0), cur_stream());
  return {dy, dres};
}

// bn_act_bwd of a layer with a residual branch that ends in a BatchNorm without activation: {dy, dres, res_sums[2,C]}
std::vector<Tensor> bn_act_bwd_res(const Tensor& dout, const Tensor& out, const Tensor& yraw, const Tensor& mean,
                                   const Tensor& invstd, const Tensor& gamma, int64_t relu, Tensor dgamma, Tensor dbeta,
                                   bool acc_gamma, bool acc_beta, c10::optional<Tensor> zeroed_scratch, bool sums_ready,
                                   const Tensor& res_yraw, const Tensor& res_mean, const Tensor& res_invstd,
                                   c10::optional<Tensor> res_sums_pre) {
  check_cl(dout, "dout"); check_cl(yraw, "yraw"); check_cl(res_yraw, "res_yraw");
  c10::cuda::CUDAGuard g(dout.device());
  auto d = dims_of(yraw);
  TORCH_CHECK(res_yraw.sizes() == yraw.sizes() && res_mean.numel() == d.C && res_invstd.numel() == d.C &&
              res_mean.scalar_type() == at::kFloat && res_invstd.scalar_type() == at::kFloat, "bn_act_bwd_res: operand mismatch");
  Tensor dy = at::empty_like(yraw), dres = at::empty_like(yraw);
  const bool pre = zeroed_scratch.has_value() && zeroed_scratch->defined();
  Tensor scratch = pre ? *zeroed_scratch : at::empty({2, d.C}, yraw.options().dtype(at::kFloat));
  const bool rpre = res_sums_pre.has_value() && res_sums_pre->defined();
  Tensor rs = rpre ? *res_sums_pre : at::empty({2, d.C}, yraw.options().dtype(at::kFloat));
  TORCH_CHECK(rs.numel() >= 2 * d.C && rs.scalar_type() == at::kFloat && rs.is_contiguous());
  int rc = hz_bn_act_bwd_res(cptr(dout), cptr(out), cptr(yraw), mean.data_ptr<float>(), invstd.data_ptr<float>(),
                             gamma.data_ptr<float>(), scratch.data_ptr<float>(), dy.data_ptr(), dres.data_ptr(),
                             dgamma.data_ptr<float>(), dbeta.data_ptr<float>(), acc_gamma ? 1 : 0, acc_beta ? 1 : 0,
                             d.N * d.H * d.W, d.C, (int)relu, (sums_ready && pre) ? 3 : (pre ? 1 : 0), res_yraw.data_ptr(),
                             res_mean.data_ptr<float>(), res_invstd.data_ptr<float>(), rs.data_ptr<float>(), rpre ? 1 : 0,
                             cur_stream());
  TORCH_CHECK(rc == 0, "hz_bn_act_bwd_res: shape / activation not covered");
  return {dy, dres, rs};
}

std::vector<Tensor> maxpool_fwd(const Tensor& x, bool want_idx) {
  check_cl(x, "x");
  c10::cuda::CUDAGuard g(x.device());
  auto d = dims_of(x);
  const int Ho = (d.H + 2 - 3) / 2 + 1, Wo = (d.W + 2 - 3) / 2 + 1;
  Tensor y = empty_cl(x, d.N, d.C, Ho, Wo);
  Tensor idx;
  if (want_idx) idx = at::empty({d.N, Ho, Wo, d.C}, x.options().dtype(at::kByte));
  hz_maxpool_fwd(cptr(x), y.data_ptr(), want_idx ? idx.data_ptr() : nullptr, d.N, d.H, d.W, d.C, cur_stream());
  return {y, idx};
}

Tensor maxpool_bwd(const Tensor& dy, const Tensor& idx, std::vector<int64_t> x_shape) {
  check_cl(dy, "dy");
  c10::cuda::CUDAGuard g(dy.device());
  const int N = (int)x_shape[0], Cc = (int)x_shape[1], H = (int)x_shape[2], W = (int)x_shape[3];
  Tensor dx = empty_cl(dy, N, Cc, H, W);
  hz_maxpool_bwd(cptr(dy), idx.data_ptr(), dx.data_ptr(), N, H, W, Cc, cur_stream());
  return dx;
}

// max-pool backward + the BatchNorm-backward sums of the layer feeding the pool: {dx, sums[2,C]}
std::vector<Tensor> maxpool_bwd_bn(const Tensor& dy, const Tensor& idx, std::vector<int64_t> x_shape,
                                   c10::optional<Tensor> bn_out, const Tensor& bn_yraw, const Tensor& bn_mean,
                                   const Tensor& bn_invstd, c10::optional<Tensor> sums_pre) {
  check_cl(dy, "dy"); check_cl(bn_yraw, "bn_yraw");
  c10::cuda::CUDAGuard g(dy.device());
  const int N = (int)x_shape[0], Cc = (int)x_shape[1], H = (int)x_shape[2], W = (int)x_shape[3];
  Tensor dx = empty_cl(dy, N, Cc, H, W);
  TORCH_CHECK(bn_yraw.sizes() == dx.sizes() && bn_mean.numel() == Cc && bn_invstd.numel() == Cc &&
              bn_mean.scalar_type() == at::kFloat && bn_invstd.scalar_type() == at::kFloat, "maxpool_bwd_bn: BN operand mismatch");
  const void* bo = nullptr;
  if (bn_out.has_value() && bn_out->defined()) {
    check_cl(*bn_out, "bn_out");
    TORCH_CHECK(bn_out->sizes() == dx.sizes(), "maxpool_bwd_bn: bn_out shape");
    bo = bn_out->data_ptr();
  }
  const bool pre = sums_pre.has_value() && sums_pre->defined();
  Tensor sums = pre ? *sums_pre : at::empty({2, Cc}, dy.options().dtype(at::kFloat));
  TORCH_CHECK(sums.numel() >= 2 * Cc && sums.scalar_type() == at::kFloat && sums.is_contiguous());
  int rc = hz_maxpool_bwd_bn(cptr(dy), idx.data_ptr(), dx.data_ptr(), bo, bn_yraw.data_ptr(), bn_mean.data_ptr<float>(),
                             bn_invstd.data_ptr<float>(), sums.data_ptr<float>(), pre ? 1 : 0, bo != nullptr ? 1 : 0, N, H, W, Cc,
                             cur_stream());
  TORCH_CHECK(rc == 0, "hz_maxpool_bwd_bn: shape not covered");
  return {dx, sums};
}

// uint8 NHWC-physical (logical NCHW channels_last view) -> normalised bf16, same layout
Tensor u8_normalize(const Tensor& img, double mean, double std) {
  TORCH_CHECK(img.is_cuda() && img.scalar_type() == at::kByte);
  c10::cuda::CUDAGuard g(img.device());
  Tensor out = at::empty_strided(img.sizes(), img.strides(), img.options().dtype(at::kBFloat16));
  hz_u8_normalize(img.data_ptr(), out.data_ptr(), (size_t)img.numel(), (float)mean, (float)std, cur_stream());
  return out;
}

Tensor im2col_small(const Tensor& x, int64_t R, int64_t stride, int64_t pad, int64_t Kp) {
  check_cl(x, "x");
  c10::cuda::CUDAGuard g(x.device());
  auto d = dims_of(x);
  const int Ho = (d.H + 2 * pad - R) / stride + 1, Wo = (d.W + 2 * pad - R) / stride + 1;
  Tensor A = at::empty({(int64_t)d.N * Ho * Wo, Kp}, x.options());
  hz_im2col_small(cptr(x), A.data_ptr(), d.N, d.H, d.W, d.C, (int)R, (int)R, (int)stride, (int)pad, Ho, Wo,
                  (int)Kp, cur_stream());
  return A;
}

// stem: (im2col matrix [N*Ho*Wo, Kp], zero-padded weights [Cout, Kp]) from one launch
std::vector<Tensor> stem_pack(const Tensor& x, const Tensor& w2d, int64_t R, int64_t stride, int64_t pad, int64_t Kp) {
  check_cl(x, "x");
  TORCH_CHECK(w2d.is_cuda() && w2d.scalar_type() == at::kBFloat16 && w2d.dim() == 2 && w2d.is_contiguous());
  c10::cuda::CUDAGuard g(x.device());
  auto d = dims_of(x);
  const int Ho = (d.H + 2 * pad - R) / stride + 1, Wo = (d.W + 2 * pad - R) / stride + 1;
  Tensor A = at::empty({(int64_t)d.N * Ho * Wo, Kp}, x.options());
  Tensor wp = at::empty({w2d.size(0), Kp}, w2d.options());
  TORCH_CHECK(w2d.size(1) == d.C * R * R, "stem_pack: weight rows must be R*R*Cin long");
  int rc = hz_stem_pack(cptr(x), A.data_ptr(), cptr(w2d), wp.data_ptr(), d.N, d.H, d.W, d.C, (int)R, (int)stride,
                        (int)pad, Ho, Wo, (int)Kp, (int)w2d.size(0), cur_stream());
  if (rc != 0) {   // shapes the fused kernel does not cover: the two generic kernels
    hz_im2col_small(cptr(x), A.data_ptr(), d.N, d.H, d.W, d.C, (int)R, (int)R, (int)stride, (int)pad, Ho, Wo, (int)Kp,
                    cur_stream());
    hz_pad_rows(cptr(w2d), wp.data_ptr(), (int)w2d.size(0), (int)w2d.size(1), (int)Kp, cur_stream());
  }
  return {A, wp};
}

// ------------------------------------------------------------------ depthwise 3x3 convolution (csrc/depthwise.cu)
// w: bf16 [C,1,3,3] (C*9 contiguous elements); stats_pre: optional pre-zeroed [2,C] fp32 slice of the statistics arena
void check_dw_weight(const Tensor& w, int C) {
  TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kBFloat16 && w.numel() == (int64_t)C * 9 && w.dim() == 4 &&
              w.size(0) == C && w.stride(0) == 9 && w.stride(2) == 3 && w.stride(3) == 1,
              "depthwise weight must be bf16 [C,1,3,3] with C*9 contiguous elements");
}
std::vector<Tensor> dwconv_fwd(const Tensor& x, const Tensor& w, int64_t stride, bool want_stats,
                               c10::optional<Tensor> stats_pre) {
  check_cl(x, "x");
  c10::cuda::CUDAGuard g(x.device());
  auto d = dims_of(x);
  check_dw_weight(w, d.C);
  TORCH_CHECK(hz_dwconv_ok(d.N, d.H, d.W, d.C, (int)stride), "dwconv_fwd: unsupported shape");
  const int Ho = (d.H - 1) / (int)stride + 1, Wo = (d.W - 1) / (int)stride + 1;
  Tensor y = empty_cl(x, d.N, d.C, Ho, Wo);
  Tensor stats;
  const bool pre = want_stats && stats_pre.has_value() && stats_pre->defined();
  if (want_stats) stats = pre ? *stats_pre : at::empty({2, d.C}, x.options().dtype(at::kFloat));
  if (pre) TORCH_CHECK(stats.numel() == 2 * d.C && stats.scalar_type() == at::kFloat && stats.is_contiguous());
  int rc = hz_dwconv_fwd(cptr(x), cptr(w), y.data_ptr(), want_stats ? stats.data_ptr<float>() : nullptr, pre ? 1 : 0,
                         d.N, d.H, d.W, d.C, (int)stride, cur_stream());
  TORCH_CHECK(rc == 0, "hz_dwconv_fwd failed rc=", rc);
  return {y, stats};
}

Tensor dwconv_dgrad(const Tensor& dy, const Tensor& w, std::vector<int64_t> x_shape, int64_t stride) {
  check_cl(dy, "dy");
  c10::cuda::CUDAGuard g(dy.device());
  const int N = (int)x_shape[0], Cc = (int)x_shape[1], H = (int)x_shape[2], W = (int)x_shape[3];
  check_dw_weight(w, Cc);
  TORCH_CHECK(hz_dwconv_ok(N, H, W, Cc, (int)stride), "dwconv_dgrad: unsupported shape");
  TORCH_CHECK(dy.size(0) == N && dy.size(1) == Cc && dy.size(2) == (H - 1) / stride + 1 && dy.size(3) == (W - 1) / stride + 1,
              "dwconv_dgrad: dy does not match x_shape / stride");
  Tensor dx = empty_cl(dy, N, Cc, H, W);
  int rc = hz_dwconv_dgrad(cptr(dy), cptr(w), dx.data_ptr(), N, H, W, Cc, (int)stride, cur_stream());
  TORCH_CHECK(rc == 0, "hz_dwconv_dgrad failed rc=", rc);
  return dx;
}

// depthwise dgrad + the BatchNorm-backward sums of the layer that produced the conv's input: {dx, sums[2,C]}
std::vector<Tensor> dwconv_dgrad_bnbwd(const Tensor& dy, const Tensor& w, std::vector<int64_t> x_shape, int64_t stride,
                                       c10::optional<Tensor> bn_out, const Tensor& bn_yraw, const Tensor& bn_mean,
                                       const Tensor& bn_invstd, c10::optional<Tensor> sums_pre, bool cap6) {
  check_cl(dy, "dy"); check_cl(bn_yraw, "bn_yraw");
  c10::cuda::CUDAGuard g(dy.device());
  const int N = (int)x_shape[0], Cc = (int)x_shape[1], H = (int)x_shape[2], W = (int)x_shape[3];
  check_dw_weight(w, Cc);
  TORCH_CHECK(hz_dwconv_ok(N, H, W, Cc, (int)stride), "dwconv_dgrad_bnbwd: unsupported shape");
  TORCH_CHECK(dy.size(0) == N && dy.size(1) == Cc && dy.size(2) == (H - 1) / stride + 1 && dy.size(3) == (W - 1) / stride + 1,
              "dwconv_dgrad_bnbwd: dy does not match x_shape / stride");
  Tensor dx = empty_cl(dy, N, Cc, H, W);
  TORCH_CHECK(bn_yraw.sizes() == dx.sizes() && bn_mean.numel() == Cc && bn_invstd.numel() == Cc &&
              bn_mean.scalar_type() == at::kFloat && bn_invstd.scalar_type() == at::kFloat, "dwconv_dgrad_bnbwd: BN operand mismatch");
  const void* bo = nullptr;
  if (bn_out.has_value() && bn_out->defined()) {
    check_cl(*bn_out, "bn_out");
    TORCH_CHECK(bn_out->sizes() == dx.sizes(), "dwconv_dgrad_bnbwd: bn_out shape");
    bo = bn_out->data_ptr();
  }
  const bool pre = sums_pre.has_value() && sums_pre->defined();
  Tensor sums = pre ? *sums_pre : at::empty({2, Cc}, dy.options().dtype(at::kFloat));
  TORCH_CHECK(sums.numel() >= 2 * Cc && sums.scalar_type() == at::kFloat && sums.is_contiguous());
  HzBnBwd b;
  b.out = bo; b.yraw = bn_yraw.data_ptr(); b.mean = bn_mean.data_ptr<float>(); b.invstd = bn_invstd.data_ptr<float>();
  b.sums = sums.data_ptr<float>(); b.sums_is_zero = pre ? 1 : 0; b.cap6 = cap6 ? 1 : 0;
  int rc = hz_dwconv_dgrad_bnbwd(cptr(dy), cptr(w), dx.data_ptr(), N, H, W, Cc, (int)stride, &b, cur_stream());
  TORCH_CHECK(rc == 0, "hz_dwconv_dgrad_bnbwd failed rc=", rc);
  return {dx, sums};
}

// dW (fp32 [C,1,3,3], C*9 contiguous: a view of the flat gradient bucket) written or accumulated in place
void dwconv_wgrad(const Tensor& dy, const Tensor& x, Tensor dw, int64_t stride, bool accumulate, bool prezeroed) {
  check_cl(dy, "dy"); check_cl(x, "x");
  c10::cuda::CUDAGuard g(x.device());
  auto d = dims_of(x);
  TORCH_CHECK(dw.is_cuda() && dw.scalar_type() == at::kFloat && dw.numel() == (int64_t)d.C * 9 && dw.dim() == 4 &&
              dw.stride(0) == 9 && dw.stride(2) == 3 && dw.stride(3) == 1, "dwconv_wgrad: dw must be fp32 [C,1,3,3], C*9 contiguous");
  TORCH_CHECK(hz_dwconv_ok(d.N, d.H, d.W, d.C, (int)stride), "dwconv_wgrad: unsupported shape");
  TORCH_CHECK(dy.size(0) == d.N && dy.size(1) == d.C && dy.size(2) == (d.H - 1) / stride + 1 &&
              dy.size(3) == (d.W - 1) / stride + 1, "dwconv_wgrad: dy does not match x / stride");
  int rc = hz_dwconv_wgrad(cptr(dy), cptr(x), dw.data_ptr<float>(), d.N, d.H, d.W, d.C, (int)stride, accumulate ? 1 : 0,
                           prezeroed ? 1 : 0, cur_stream());
  TORCH_CHECK(rc == 0, "hz_dwconv_wgrad failed rc=", rc);
}

Tensor pad_rows(const Tensor& w2d, int64_t Kp) {
  TORCH_CHECK(w2d.is_cuda() && w2d.scalar_type() == at::kBFloat16 && w2d.dim() == 2 && w2d.is_contiguous());
  c10::cuda::CUDAGuard g(w2d.device());
  Tensor out = at::empty({w2d.size(0), Kp}, w2d.options());
  hz_pad_rows(cptr(w2d), out.data_ptr(), (int)w2d.size(0), (int)w2d.size(1), (int)Kp, cur_stream());
  return out;
}

std::vector<Tensor> head_fwd_bwd(const Tensor& feat, const Tensor& W, const c10::optional<Tensor>& bias,
                                 const Tensor& labels, double loss_scale, int64_t n_valid, Tensor dW,
                                 c10::optional<Tensor> db, bool accumulate, bool need_dfeat,
                                 c10::optional<Tensor> zeroed2) {
  check_cl(feat, "feat");
  TORCH_CHECK(W.scalar_type() == at::kFloat && W.is_contiguous());
  TORCH_CHECK(labels.scalar_type() == at::kLong);
  c10::cuda::CUDAGuard g(feat.device());
  auto d = dims_of(feat);
  const int K = (int)W.size(0);
  TORCH_CHECK(K <= 64, "head supports at most 64 (padded) classes");
  auto fo = feat.options().dtype(at::kFloat);
  Tensor pooled = at::empty({d.N, d.C}, fo), dlogits = at::empty({d.N, K}, fo), logits = at::empty({d.N, K}, fo);
  const bool pre = zeroed2.has_value() && zeroed2->defined() && zeroed2->numel() >= 2;
  Tensor loss = pre ? zeroed2->view({-1})[0] : at::empty({}, fo);
  Tensor correct = pre ? zeroed2->view({-1})[1] : at::empty({}, fo);
  Tensor dfeat;
  if (need_dfeat) dfeat = at::empty_like(feat);
  hz_head_fwd_bwd(cptr(feat), W.data_ptr<float>(), fptr(bias), labels.data_ptr<int64_t>(),
                  pooled.data_ptr<float>(), dlogits.data_ptr<float>(), logits.data_ptr<float>(),
                  need_dfeat ? dfeat.data_ptr() : nullptr, loss.data_ptr<float>(), correct.data_ptr<float>(),
                  dW.data_ptr<float>(), fptr(db), d.N, d.C, d.H * d.W, K, (int)n_valid, (float)loss_scale,
                  accumulate ? 1 : 0, pre ? 1 : 0, cur_stream());
  return {loss, correct, dfeat, logits};
}

void adam_step(Tensor master, Tensor grad, Tensor m, Tensor v, c10::optional<Tensor> shadow, Tensor step,
               double lr, double b1, double b2, double eps, double gscale, c10::optional<Tensor> prev,
               c10::optional<Tensor> diff_out, bool zero_grad, c10::optional<Tensor> live_blocks, bool bump,
               int64_t max_ctas) {
  TORCH_CHECK(master.is_cuda() && master.scalar_type() == at::kFloat && master.numel() % 4 == 0);
  c10::cuda::CUDAGuard g(master.device());
  void* sh = nullptr;
  if (shadow.has_value() && shadow->defined()) { TORCH_CHECK(shadow->scalar_type() == at::kBFloat16); sh = shadow->data_ptr(); }
  hz_adam(master.data_ptr<float>(), grad.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), sh,
          step.data_ptr<float>(), fptr(prev), fptr(diff_out), zero_grad ? 1 : 0, (size_t)master.numel(), (float)lr,
          (float)b1, (float)b2, (float)eps, (float)gscale,
          live_blocks.has_value() && live_blocks->defined() ? live_blocks->data_ptr<int>() : nullptr,
          live_blocks.has_value() && live_blocks->defined() ? (size_t)live_blocks->numel() : 0, bump ? 1 : 0, (int)max_ctas, cur_stream());
}

Tensor grad_diff_sq(const Tensor& grad, Tensor prev) {
  TORCH_CHECK(grad.is_cuda() && grad.numel() % 4 == 0);
  c10::cuda::CUDAGuard g(grad.device());
  Tensor out = at::empty({}, grad.options());
  hz_grad_diff(grad.data_ptr<float>(), prev.data_ptr<float>(), out.data_ptr<float>(), (size_t)grad.numel(),
               cur_stream());
  return out;
}

void stats_update(Tensor stats, Tensor has_prev, const Tensor& loss, const Tensor& correct, double batch,
                  const c10::optional<Tensor>& diff_sq) {
  c10::cuda::CUDAGuard g(stats.device());
  hz_stats_update(stats.data_ptr<float>(), has_prev.data_ptr<float>(), loss.data_ptr<float>(),
                  correct.data_ptr<float>(), (float)batch, fptr(diff_sq), cur_stream());
}

// ------------------------------------------------------------------ tcgen05 convs
bool conv_supported(int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t R, int64_t stride,
                    int64_t pad) {
  return hz_conv_supported((int)N, (int)H, (int)W, (int)Cin, (int)Cout, (int)R, (int)stride, (int)pad) != 0;
}

std::vector<Tensor> conv_fwd(const Tensor& x, const Tensor& w, int64_t stride, int64_t pad, bool want_stats,
                             c10::optional<Tensor> zeroed_stats, bool weights_stable) {
  check_cl(x, "x"); check_cl(w, "w");
  c10::cuda::CUDAGuard g(x.device());
  auto d = dims_of(x);
  const int Cout = (int)w.size(0), R = (int)w.size(2);
  const int Ho = (d.H + 2 * pad - R) / stride + 1, Wo = (d.W + 2 * pad - R) / stride + 1;
  Tensor y = empty_cl(x, d.N, Cout, Ho, Wo);
  Tensor stats;
  const bool pre = want_stats && zeroed_stats.has_value() && zeroed_stats->defined();
  if (pre) stats = *zeroed_stats;
  else if (want_stats) stats = at::empty({2, Cout}, x.options().dtype(at::kFloat));
  int rc = hz_conv_fwd(cptr(x), cptr(w), y.data_ptr(), want_stats ? stats.data_ptr<float>() : nullptr, pre ? 1 : 0,
                       d.N, d.H, d.W, d.C, Cout, R, (int)stride, (int)pad, weights_stable ? 1 : 0, nullptr, cur_stream());
  TORCH_CHECK(rc == 0, "hz_conv_fwd failed rc=", rc);
  return {y, stats};
}

// conv -> BatchNorm(batch statistics) -> (+residual) -> (ReLU) in ONE kernel (device-wide barrier inside the conv
// epilogue).  `scratch`: pre-zeroed fp32 slice of >= 2*Cout + 32 elements ([Σy | Σy² | barrier counter]).
// Returns {y_raw, out, mean, invstd}; when the grid would exceed the SM count (no co-residency) the unfused pair of
// kernels runs instead — same results either way.
std::vector<Tensor> conv_bn_act_fwd(const Tensor& x, const Tensor& w, int64_t stride, int64_t pad, Tensor scratch,
                                    const Tensor& gamma, const Tensor& beta, const c10::optional<Tensor>& rmean,
                                    const c10::optional<Tensor>& rvar, double momentum, double eps,
                                    const c10::optional<Tensor>& residual, bool relu, bool weights_stable) {
  check_cl(x, "x"); check_cl(w, "w");
  c10::cuda::CUDAGuard g(x.device());
  auto d = dims_of(x);
  const int Cout = (int)w.size(0), R = (int)w.size(2);
  const int Ho = (d.H + 2 * pad - R) / stride + 1, Wo = (d.W + 2 * pad - R) / stride + 1;
  TORCH_CHECK(scratch.scalar_type() == at::kFloat && scratch.numel() >= 2 * Cout + 32 && scratch.is_contiguous());
  Tensor y = empty_cl(x, d.N, Cout, Ho, Wo);
  Tensor out = at::empty_like(y);
  Tensor mean = at::empty({Cout}, x.options().dtype(at::kFloat));
  Tensor invstd = at::empty({Cout}, x.options().dtype(at::kFloat));
  const void* res = nullptr;
  if (residual.has_value() && residual->defined()) {
    check_cl(*residual, "residual");
    TORCH_CHECK(residual->sizes() == y.sizes(), "residual shape mismatch");
    res = residual->data_ptr();
  }
  HzBnFuse bn;
  bn.gamma = gamma.data_ptr<float>(); bn.beta = beta.data_ptr<float>();
  bn.mean = mean.data_ptr<float>(); bn.invstd = invstd.data_ptr<float>();
  bn.rmean = fptr(rmean); bn.rvar = fptr(rvar);
  bn.residual = res; bn.out = out.data_ptr();
  bn.counter = reinterpret_cast<unsigned*>(scratch.data_ptr<float>() + 2 * Cout);
  bn.eps = (float)eps; bn.momentum = (float)momentum; bn.relu = relu ? 1 : 0;
  int rc = hz_conv_fwd(cptr(x), cptr(w), y.data_ptr(), scratch.data_ptr<float>(), 1, d.N, d.H, d.W, d.C, Cout, R,
                       (int)stride, (int)pad, weights_stable ? 1 : 0, &bn, cur_stream());
  if (rc == -20) {
    rc = hz_conv_fwd(cptr(x), cptr(w), y.data_ptr(), scratch.data_ptr<float>(), 1, d.N, d.H, d.W, d.C, Cout, R,
                     (int)stride, (int)pad, weights_stable ? 1 : 0, nullptr, cur_stream());
    TORCH_CHECK(rc == 0, "hz_conv_fwd failed rc=", rc);
    hz_bn_act_fwd(y.data_ptr(), scratch.data_ptr<float>(), bn.gamma, bn.beta, res, out.data_ptr(), bn.mean, bn.invstd,
                  bn.rmean, bn.rvar, d.N * Ho * Wo, Cout, (float)eps, (float)momentum, relu ? 1 : 0, 1, cur_stream());
    return {y, out, mean, invstd};
  }
  TORCH_CHECK(rc == 0, "hz_conv_fwd (fused BN) failed rc=", rc);
  return {y, out, mean, invstd};
}

// addend (optional, same shape/layout as dx): dx = dgrad(dy, w) + addend, fused into the epilogue
// dgrad + the BatchNorm-backward sums of the layer that produced the conv's input (see HzBnBwd): returns {dx, sums[2,Cin]}
std::vector<Tensor> conv_dgrad_bnbwd(const Tensor& dy, const Tensor& w, std::vector<int64_t> x_shape, int64_t stride,
                                     int64_t pad, c10::optional<Tensor> addend, bool weights_stable,
                                     c10::optional<Tensor> bn_out, const Tensor& bn_yraw, const Tensor& bn_mean,
                                     const Tensor& bn_invstd, c10::optional<Tensor> sums_pre, bool cap6) {
  check_cl(dy, "dy"); check_cl(w, "w"); check_cl(bn_yraw, "bn_yraw");
  c10::cuda::CUDAGuard g(dy.device());
  const int N = (int)x_shape[0], Cin = (int)x_shape[1], H = (int)x_shape[2], W = (int)x_shape[3];
  Tensor dx = empty_cl(dy, N, Cin, H, W);
  TORCH_CHECK(bn_yraw.sizes() == dx.sizes() && bn_mean.numel() == Cin && bn_invstd.numel() == Cin &&
              bn_mean.scalar_type() == at::kFloat && bn_invstd.scalar_type() == at::kFloat, "conv_dgrad_bnbwd: BN operand mismatch");
  const void* add = nullptr;
  if (addend.has_value() && addend->defined()) {
    check_cl(*addend, "addend");
    TORCH_CHECK(addend->sizes() == dx.sizes() && addend->scalar_type() == dx.scalar_type(), "dgrad addend mismatch");
    add = addend->data_ptr();
  }
  const void* bo = nullptr;
  if (bn_out.has_value() && bn_out->defined()) {
    check_cl(*bn_out, "bn_out");
    TORCH_CHECK(bn_out->sizes() == dx.sizes(), "conv_dgrad_bnbwd: bn_out shape");
    bo = bn_out->data_ptr();
  }
  const bool pre = sums_pre.has_value() && sums_pre->defined();
  Tensor sums = pre ? *sums_pre : at::empty({2, Cin}, dy.options().dtype(at::kFloat));
  TORCH_CHECK(sums.numel() >= 2 * Cin && sums.scalar_type() == at::kFloat && sums.is_contiguous());
  HzBnBwd b;
  b.out = bo; b.yraw = bn_yraw.data_ptr(); b.mean = bn_mean.data_ptr<float>(); b.invstd = bn_invstd.data_ptr<float>();
  b.sums = sums.data_ptr<float>(); b.sums_is_zero = pre ? 1 : 0;
  b.cap6 = cap6 ? 1 : 0;
  int rc = hz_conv_dgrad_bnbwd(cptr(dy), cptr(w), dx.data_ptr(), add, N, H, W, Cin, (int)w.size(0), (int)w.size(2),
                               (int)stride, (int)pad, weights_stable ? 1 : 0, &b, cur_stream());
  TORCH_CHECK(rc == 0, "hz_conv_dgrad_bnbwd failed rc=", rc);
  return {dx, sums};
}

Tensor conv_dgrad(const Tensor& dy, const Tensor& w, std::vector<int64_t> x_shape, int64_t stride, int64_t pad,
                  c10::optional<Tensor> addend, bool weights_stable) {
  check_cl(dy, "dy"); check_cl(w, "w");
  c10::cuda::CUDAGuard g(dy.device());
  const int N = (int)x_shape[0], Cin = (int)x_shape[1], H = (int)x_shape[2], W = (int)x_shape[3];
  Tensor dx = empty_cl(dy, N, Cin, H, W);
  const void* add = nullptr;
  if (addend.has_value() && addend->defined()) {
    check_cl(*addend, "addend");
    TORCH_CHECK(addend->sizes() == dx.sizes() && addend->scalar_type() == dx.scalar_type(), "dgrad addend mismatch");
    add = addend->data_ptr();
  }
  int rc = hz_conv_dgrad(cptr(dy), cptr(w), dx.data_ptr(), add, N, H, W, Cin, (int)w.size(0), (int)w.size(2),
                         (int)stride, (int)pad, weights_stable ? 1 : 0, cur_stream());
  TORCH_CHECK(rc == 0, "hz_conv_dgrad failed rc=", rc);
  return dx;
}

// dw_out: fp32 tensor whose storage is [Cout, R, S, Cin]-physical (a view of the flat gradient bucket)
void conv_wgrad(const Tensor& dy, const Tensor& x, Tensor dw_out, int64_t R, int64_t stride, int64_t pad,
                bool accumulate, bool prezeroed, int64_t ld_out, int64_t n_valid) {
  check_cl(dy, "dy"); check_cl(x, "x");
  TORCH_CHECK(dw_out.scalar_type() == at::kFloat);
  c10::cuda::CUDAGuard g(dy.device());
  auto d = dims_of(x);
  int rc = hz_conv_wgrad(cptr(dy), cptr(x), dw_out.data_ptr<float>(), d.N, d.H, d.W, d.C, (int)dy.size(1), (int)R,
                         (int)stride, (int)pad, accumulate ? 1 : 0, prezeroed ? 1 : 0, (long long)ld_out, (int)n_valid,
                         cur_stream());
  TORCH_CHECK(rc == 0, "hz_conv_wgrad failed rc=", rc);
}

// ------------------------------------------------------------------ peer all-reduce
class PeerComm {
 public:
  PeerComm(int rank, int world, int device, int64_t max_wire_bytes, int max_blocks, int64_t heap_bytes)
      : rank_(rank), world_(world), device_(device) {
    c_ = hz_comm_create2(rank, world, device, (size_t)max_wire_bytes, max_blocks, (size_t)heap_bytes);
    TORCH_CHECK(c_ != nullptr, "hz_comm_create failed");
  }
  int64_t heap_bytes() { return (int64_t)hz_comm_heap_bytes(c_); }
  // torch view of [offset, offset+numel*itemsize) of the LOCAL symmetric heap (storage owned by the comm)
  Tensor heap_tensor(int64_t offset, std::vector<int64_t> sizes, std::vector<int64_t> strides, const std::string& dtype) {
    auto dt = dtype == "bf16" ? at::kBFloat16 : dtype == "f32" ? at::kFloat : dtype == "i32" ? at::kInt : at::kByte;
    char* base = hz_comm_heap_base(c_, rank_);
    TORCH_CHECK(base != nullptr && offset >= 0, "no symmetric heap");
    auto opts = at::TensorOptions().dtype(dt).device(at::kCUDA, device_);
    return at::from_blob(base + offset, sizes, strides, opts);
  }
  // base address of rank r's symmetric heap as mapped in this process (0 when not mapped)
  int64_t heap_ptr(int64_t r) { return (int64_t)(uintptr_t)hz_comm_heap_base(c_, (int)r); }
  ~PeerComm() { hz_comm_destroy(c_); }
  py::bytes export_handles() {
    char h[64];
    TORCH_CHECK(hz_comm_export(c_, h) == 0, "cudaIpcGetMemHandle failed");
    return py::bytes(h, 64);
  }
  void import_handles(const std::vector<std::string>& hs) {
    TORCH_CHECK((int)hs.size() == world_);
    std::string all;
    for (auto& s : hs) { TORCH_CHECK(s.size() == 64); all += s; }
    TORCH_CHECK(hz_comm_import(c_, all.data()) == 0, "cudaIpcOpenMemHandle failed");
  }
  static void link_local(std::vector<PeerComm*> comms) {
    std::vector<HzComm*> raw;
    for (auto* c : comms) raw.push_back(c->c_);
    hz_comm_link_local(raw.data(), (int)raw.size());
  }
  int64_t symm_bytes() { return (int64_t)hz_comm_symm_bytes(c_); }
  void set_multicast(int64_t mc_ptr, int64_t local_ptr, int64_t bytes) {
    hz_comm_set_multicast(c_, (void*)mc_ptr, (void*)local_ptr, (size_t)bytes);
  }
  void allreduce(Tensor grad, const std::string& algo, bool wire_bf16, double scale,
                 c10::optional<Tensor> live_blocks) {
    TORCH_CHECK(grad.is_cuda() && grad.scalar_type() == at::kFloat && grad.is_contiguous());
    const bool has_live = live_blocks.has_value() && live_blocks->defined();
    if (has_live) TORCH_CHECK(live_blocks->scalar_type() == at::kInt && live_blocks->is_cuda());
    int a = algo == "oneshot" ? 0 : algo == "twoshot" ? 1 : algo == "nvls" ? 2 : algo == "ll" ? 3 : algo == "bulk" ? 4 : -1;
    TORCH_CHECK(a >= 0, "unknown all-reduce algorithm ", algo);
    c10::cuda::CUDAGuard g(grad.device());
    const size_t n = has_live ? (size_t)live_blocks->numel() * 64 : (size_t)grad.numel();
    int rc = hz_comm_allreduce(c_, grad.data_ptr<float>(), n, a, wire_bf16 ? 1 : 0, (float)scale,
                               has_live ? live_blocks->data_ptr<int>() : nullptr, cur_stream());
    TORCH_CHECK(rc == 0, "hz_comm_allreduce failed rc=", rc);
  }
  // all-reduce (average) + Adam on the bucket in one kernel; every tensor is the bucket's slice of its flat buffer
  void allreduce_adam(Tensor grad, const std::string& algo, bool wire_bf16, double scale,
                      c10::optional<Tensor> live_blocks, Tensor master, Tensor m, Tensor v, c10::optional<Tensor> shadow,
                      c10::optional<Tensor> prev, c10::optional<Tensor> diff_out, Tensor step, double lr, double b1,
                      double b2, double eps, bool bump) {
    TORCH_CHECK(grad.is_cuda() && grad.scalar_type() == at::kFloat && grad.is_contiguous());
    TORCH_CHECK(master.numel() == grad.numel() && m.numel() == grad.numel() && v.numel() == grad.numel());
    const bool has_live = live_blocks.has_value() && live_blocks->defined();
    if (has_live) TORCH_CHECK(live_blocks->scalar_type() == at::kInt && live_blocks->is_cuda());
    int a = algo == "oneshot" ? 0 : algo == "twoshot" ? 1 : algo == "nvls" ? 2 : algo == "ll" ? 3 : algo == "bulk" ? 4 : -1;
    TORCH_CHECK(a >= 0, "unknown all-reduce algorithm ", algo);
    c10::cuda::CUDAGuard g(grad.device());
    const size_t n = has_live ? (size_t)live_blocks->numel() * 64 : (size_t)grad.numel();
    void* sh = nullptr;
    if (shadow.has_value() && shadow->defined()) { TORCH_CHECK(shadow->scalar_type() == at::kBFloat16); sh = shadow->data_ptr(); }
    int rc = hz_comm_allreduce_adam(c_, grad.data_ptr<float>(), n, a, wire_bf16 ? 1 : 0, (float)scale,
                                    has_live ? live_blocks->data_ptr<int>() : nullptr, master.data_ptr<float>(),
                                    m.data_ptr<float>(), v.data_ptr<float>(), sh, fptr(prev), fptr(diff_out),
                                    step.data_ptr<float>(), (float)lr, (float)b1, (float)b2, (float)eps, bump ? 1 : 0,
                                    cur_stream());
    TORCH_CHECK(rc == 0, "hz_comm_allreduce_adam failed rc=", rc);
  }
  // ZeRO-1 step of one bucket in one kernel (csrc/comm.cu zero1_kernel): grad / master / shadow are the bucket's
  // slices of the flat buffers, m / v / prev this rank's shards (zero1_shard(n_wire) elements)
  void zero1_step(Tensor grad, double scale, c10::optional<Tensor> live_blocks, Tensor master, Tensor m, Tensor v,
                  Tensor shadow, c10::optional<Tensor> prev, c10::optional<Tensor> diff_out, Tensor step, double lr,
                  double b1, double b2, double eps, bool bump) {
    TORCH_CHECK(grad.is_cuda() && grad.scalar_type() == at::kFloat && grad.is_contiguous());
    TORCH_CHECK(master.scalar_type() == at::kFloat && master.numel() == grad.numel() && master.is_contiguous());
    TORCH_CHECK(shadow.scalar_type() == at::kBFloat16 && shadow.numel() == grad.numel() && shadow.is_contiguous());
    const bool has_live = live_blocks.has_value() && live_blocks->defined();
    if (has_live) TORCH_CHECK(live_blocks->scalar_type() == at::kInt && live_blocks->is_cuda());
    const size_t n = has_live ? (size_t)live_blocks->numel() * 64 : (size_t)grad.numel();
    const int64_t shard = (int64_t)hz_comm_zero1_shard(n, world_);
    TORCH_CHECK(m.scalar_type() == at::kFloat && v.scalar_type() == at::kFloat && m.numel() == shard && v.numel() == shard &&
                m.is_contiguous() && v.is_contiguous(), "zero1_step: moment shards must have ", shard, " elements");
    if (prev.has_value() && prev->defined()) TORCH_CHECK(prev->scalar_type() == at::kFloat && prev->numel() == shard);
    c10::cuda::CUDAGuard g(grad.device());
    int rc = hz_comm_zero1_step(c_, grad.data_ptr<float>(), n, (float)scale, has_live ? live_blocks->data_ptr<int>() : nullptr,
                                master.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), shadow.data_ptr(),
                                fptr(prev), fptr(diff_out), step.data_ptr<float>(), (float)lr, (float)b1, (float)b2,
                                (float)eps, bump ? 1 : 0, cur_stream());
    TORCH_CHECK(rc == 0, "hz_comm_zero1_step failed rc=", rc);
  }
  int64_t zero1_shard(int64_t n_wire) { return (int64_t)hz_comm_zero1_shard((size_t)n_wire, world_); }
  int64_t blocks_for(int64_t n, const std::string& algo, bool wire_bf16) {
    int a = algo == "oneshot" ? 0 : algo == "twoshot" ? 1 : algo == "ll" ? 3 : algo == "bulk" ? 4 : 2;
    return hz_comm_blocks_for(c_, (size_t)n, a, wire_bf16 ? 1 : 0);
  }
  void set_block_cap(int64_t cap) { hz_comm_set_block_cap(c_, (int)cap); }
  void barrier(c10::optional<Tensor> stamps) {
    long long* s = nullptr;
    if (stamps.has_value() && stamps->defined()) s = (long long*)stamps->data_ptr<int64_t>();
    TORCH_CHECK(hz_comm_barrier(c_, s, cur_stream()) == 0);
  }
  int error() { return hz_comm_error(c_); }

 private:
  HzComm* c_;
  int rank_, world_, device_;
};


// ------------------------------------------------------------------ fused tensor-parallel ops (csrc/tp_fused.cu)
// `heaps`: base address of every rank's symmetric heap as mapped in this process; `mc`: multicast mapping or 0.
// All *_off arguments are byte offsets into the heap.  ctrl_off: u32 call counters of this op (one per CTA).
struct HeapPtrs { char* h[8]; };
HeapPtrs heap_ptrs(const std::vector<int64_t>& heaps) {
  TORCH_CHECK(heaps.size() >= 1 && heaps.size() <= 8, "1..8 tensor-parallel ranks");
  HeapPtrs o{};
  for (size_t i = 0; i < heaps.size(); ++i) { TORCH_CHECK(heaps[i] != 0, "peer heap not mapped"); o.h[i] = (char*)(uintptr_t)heaps[i]; }
  return o;
}

int64_t tp_tiles(int64_t kind, std::vector<int64_t> x_shape, int64_t Cout, int64_t stride) {
  return hz_tp_tiles((int)kind, (int)x_shape[0], (int)x_shape[2], (int)x_shape[3], (int)x_shape[1], (int)Cout, (int)stride);
}

// kind 0: out[N,Cout,H,W] = reduce_r conv(a_r[N,Cin,H,W], w_r[Cout,Cin,R,R]) (stride 1)
// kind 1: out[N,Cin,H,W]  = reduce_r dgrad(a_r = dy_r[N,Cout,Ho,Wo], w_r[Cout,Cin,R,R])  (x_shape = [N,Cin,H,W])
// mode 0 none | 1 all-reduce | 2 reduce-scatter (rank tile%W keeps).  ag: a is None, A shards live at a_off in each heap.
Tensor tp_conv(int64_t kind, c10::optional<Tensor> a, int64_t a_off, const Tensor& w, std::vector<int64_t> x_shape,
               int64_t stride, int64_t pad, c10::optional<Tensor> addend, c10::optional<Tensor> stats,
               std::vector<int64_t> heaps, int64_t mc, int64_t part_off, int64_t part_stride, int64_t cnt_off,
               int64_t ready_off, int64_t ctrl_off, int64_t rank, int64_t mode, bool nvls, bool ag, bool ll) {
  check_cl(w, "w");
  c10::cuda::CUDAGuard g(w.device());
  const int world = (int)heaps.size();
  HeapPtrs hp = heap_ptrs(heaps);
  const int N = (int)x_shape[0], Cin = (int)x_shape[1], H = (int)x_shape[2], Wd = (int)x_shape[3];
  const int Cout = (int)w.size(0), R = (int)w.size(2);
  TORCH_CHECK((int)w.size(1) == Cin, "tp_conv: weight Cin mismatch");
  const void* xp[8] = {nullptr};
  if (ag) {
    for (int r = 0; r < world; ++r) xp[r] = hp.h[r] + a_off;
  } else {
    TORCH_CHECK(a.has_value() && a->defined(), "tp_conv: operand missing");
    check_cl(*a, "a");
    xp[rank] = a->data_ptr();
  }
  const int Ho = (H + 2 * (int)pad - R) / (int)stride + 1, Wo = (Wd + 2 * (int)pad - R) / (int)stride + 1;
  Tensor out = kind == 0 ? at::empty({N, Cout, Ho, Wo}, w.options().memory_format(at::MemoryFormat::ChannelsLast))
                         : at::empty({N, Cin, H, Wd}, w.options().memory_format(at::MemoryFormat::ChannelsLast));
  const void* add = nullptr;
  if (addend.has_value() && addend->defined()) {
    check_cl(*addend, "addend");
    TORCH_CHECK(addend->sizes() == out.sizes(), "tp_conv: addend shape mismatch");
    add = addend->data_ptr();
  }
  float* st = nullptr;
  if (stats.has_value() && stats->defined()) {
    TORCH_CHECK(stats->scalar_type() == at::kFloat && stats->numel() >= 2 * out.size(1) && stats->is_contiguous());
    st = stats->data_ptr<float>();
  }
  unsigned* ctrl = reinterpret_cast<unsigned*>(hp.h[rank] + ctrl_off);
  int rc = hz_tp_conv((int)kind, xp, w.data_ptr(), out.data_ptr(), add, st, hp.h, (char*)(uintptr_t)mc, part_off,
                      part_stride, cnt_off, ready_off, ctrl, world, (int)rank, (int)mode, nvls ? 1 : 0,
                      ll ? 1 : 0, ag ? 1 : 0, N, H, Wd, Cin, Cout, R, (int)stride, (int)pad, cur_stream());
  TORCH_CHECK(rc == 0, "hz_tp_conv failed rc=", rc);
  return out;
}

// Tensor-parallel head: returns {loss, correct, dfeat, logits[N,K]}; dW_r/db_r are written into dW/db.
std::vector<Tensor> tp_head(const Tensor& feat, const Tensor& Wl, const c10::optional<Tensor>& bl, const Tensor& labels,
                            double loss_scale, int64_t n_valid, Tensor dW, c10::optional<Tensor> db, bool accumulate,
                            bool need_dfeat, c10::optional<Tensor> zeroed2, std::vector<int64_t> heaps, int64_t mc,
                            int64_t logits_off, int64_t dfeat_off, int64_t ctrl_off, int64_t rank, bool nvls) {
  check_cl(feat, "feat");
  TORCH_CHECK(Wl.scalar_type() == at::kFloat && Wl.is_contiguous() && labels.scalar_type() == at::kLong);
  c10::cuda::CUDAGuard g(feat.device());
  auto d = dims_of(feat);
  const int world = (int)heaps.size();
  HeapPtrs hp = heap_ptrs(heaps);
  const int kl = (int)Wl.size(0), K = kl * world;
  TORCH_CHECK(K <= 64 && kl <= 16, "tp_head: at most 64 padded classes");
  auto fo = feat.options().dtype(at::kFloat);
  Tensor pooled = at::empty({d.N, d.C}, fo), dl = at::empty({d.N, kl}, fo), logits = at::empty({d.N, K}, fo);
  const bool pre = zeroed2.has_value() && zeroed2->defined() && zeroed2->numel() >= 2;
  Tensor acc2 = pre ? zeroed2->view({-1}) : at::zeros({2}, fo);
  Tensor loss = acc2[0], correct = acc2[1];
  Tensor dfeat;
  if (need_dfeat) dfeat = at::empty_like(feat);
  unsigned* ctrl = reinterpret_cast<unsigned*>(hp.h[rank] + ctrl_off);
  int rc = hz_tp_head(cptr(feat), Wl.data_ptr<float>(), fptr(bl), labels.data_ptr<int64_t>(), pooled.data_ptr<float>(),
                      dl.data_ptr<float>(), logits.data_ptr<float>(), need_dfeat ? dfeat.data_ptr() : nullptr,
                      loss.data_ptr<float>(), correct.data_ptr<float>(), hp.h, (char*)(uintptr_t)mc, logits_off, dfeat_off,
                      ctrl, world, (int)rank, nvls ? 1 : 0, d.N, d.C, d.H * d.W, kl, (int)n_valid,
                      (float)loss_scale, cur_stream());
  TORCH_CHECK(rc == 0, "hz_tp_head failed rc=", rc);
  hz_head_wgrad(pooled.data_ptr<float>(), dl.data_ptr<float>(), dW.data_ptr<float>(), fptr(db), d.N, d.C, kl,
                accumulate ? 1 : 0, cur_stream());
  return {loss, correct, dfeat, logits};
}

int64_t tp_head_bytes(int64_t N, int64_t C, int64_t K, int64_t world) {
  return (int64_t)hz_tp_head_bytes((int)N, (int)C, (int)K, (int)world);
}

// sum over the tensor-parallel ranks of a small bf16 tensor (dense layout preserved)
Tensor tp_allreduce_bf16(const Tensor& in, std::vector<int64_t> heaps, int64_t mc, int64_t buf_off, int64_t cnt_off,
                         int64_t ctrl_off, int64_t rank, bool nvls, int64_t blocks, bool ll) {
  TORCH_CHECK(in.is_cuda() && in.scalar_type() == at::kBFloat16 && in.is_non_overlapping_and_dense() && in.numel() % 8 == 0);
  c10::cuda::CUDAGuard g(in.device());
  HeapPtrs hp = heap_ptrs(heaps);
  Tensor out = at::empty_like(in);
  unsigned* ctrl = reinterpret_cast<unsigned*>(hp.h[rank] + ctrl_off);
  int rc = hz_tp_allreduce_bf16(in.data_ptr(), out.data_ptr(), (size_t)in.numel(), hp.h, (char*)(uintptr_t)mc, buf_off,
                                cnt_off, ctrl, (int)heaps.size(), (int)rank, nvls ? 1 : 0, ll ? 1 : 0,
                                (int)blocks, cur_stream());
  TORCH_CHECK(rc == 0, "hz_tp_allreduce_bf16 failed rc=", rc);
  return out;
}

}  // namespace

// Failure diagnostics: a native crash inside a rank otherwise dies silently under the launcher; with this
// handler the rank prints its C/C++ frames (resolve with `addr2line -e horizonml_b200/_C.so`) before re-raising.
static void hz_segv_handler(int sig) {
  void* frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "\n[hz] fatal signal in native code, backtrace:\n";
  (void)!write(2, msg, sizeof(msg) - 1);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "horizonml_b200 sm_100a kernels";
  m.def("install_crash_backtrace", [] { signal(SIGSEGV, hz_segv_handler); signal(SIGBUS, hz_segv_handler); signal(SIGABRT, hz_segv_handler); });
  m.def("channel_ok", &channel_ok);
  m.def("channel_sums", &channel_sums);
  m.def("bn_act_fwd", &bn_act_fwd);
  m.def("bn_act_bwd", &bn_act_bwd);
  m.def("bn_act_bwd_res", &bn_act_bwd_res);
  m.def("maxpool_fwd", &maxpool_fwd);
  m.def("maxpool_bwd", &maxpool_bwd);
  m.def("maxpool_bwd_bn", &maxpool_bwd_bn);
  m.def("u8_normalize", &u8_normalize);
  m.def("im2col_small", &im2col_small);
  m.def("pad_rows", &pad_rows);
  m.def("dwconv_ok", [](int64_t N, int64_t H, int64_t W, int64_t C, int64_t stride) {
    return hz_dwconv_ok((int)N, (int)H, (int)W, (int)C, (int)stride) != 0; });
  m.def("dwconv_fwd", &dwconv_fwd);
  m.def("dwconv_dgrad", &dwconv_dgrad);
  m.def("dwconv_dgrad_bnbwd", &dwconv_dgrad_bnbwd);
  m.def("dwconv_wgrad", &dwconv_wgrad);
  m.def("conv_set_debug", [](c10::optional<Tensor> buf) {
    if (buf.has_value() && buf->defined()) {
      TORCH_CHECK(buf->is_cuda() && buf->scalar_type() == at::kLong && buf->is_contiguous());
      hz_conv_set_debug(reinterpret_cast<long long*>(buf->data_ptr<int64_t>()));
    } else {
      hz_conv_set_debug(nullptr);
    }
  });
  m.def("conv_set_persist", [](int64_t mode) { return (int64_t)hz_conv_set_persist((int)mode); });
  m.def("cluster_capacity", [] { int v[4]; hz_cluster_capacity(v); return std::vector<int64_t>{v[0], v[1], v[2], v[3]}; });
  m.def("conv_bn_act_fwd", &conv_bn_act_fwd);
  m.def("stem_pack", &stem_pack);
  m.def("head_fwd_bwd", &head_fwd_bwd);
  m.def("adam_step", &adam_step);
  m.def("grad_diff_sq", &grad_diff_sq);
  m.def("stats_update", &stats_update);
  m.def("conv_supported", &conv_supported);
  m.def("conv_shape_ok", [](int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t R, int64_t stride,
                            int64_t pad) {
    return hz_conv_shape_ok((int)N, (int)H, (int)W, (int)Cin, (int)Cout, (int)R, (int)stride, (int)pad) != 0;
  });
  m.def("conv_fwd", &conv_fwd);
  m.def("conv_dgrad", &conv_dgrad);
  m.def("conv_dgrad_bnbwd", &conv_dgrad_bnbwd);
  m.def("conv_wgrad", &conv_wgrad);
  m.def("tp_set_debug", [](c10::optional<Tensor> buf) {
    if (buf.has_value() && buf->defined()) {
      TORCH_CHECK(buf->is_cuda() && buf->scalar_type() == at::kLong && buf->is_contiguous());
      hz_tp_set_debug(reinterpret_cast<long long*>(buf->data_ptr<int64_t>()));
    } else {
      hz_tp_set_debug(nullptr);
    }
  });
  m.def("tp_tiles", &tp_tiles);
  m.def("tp_conv", &tp_conv);
  m.def("tp_head", &tp_head);
  m.def("tp_head_bytes", &tp_head_bytes);
  m.def("tp_allreduce_bf16", &tp_allreduce_bf16);
  py::class_<PeerComm>(m, "PeerComm")
      .def(py::init<int, int, int, int64_t, int, int64_t>(), py::arg("rank"), py::arg("world"), py::arg("device"),
           py::arg("max_wire_bytes"), py::arg("max_blocks"), py::arg("heap_bytes") = 0)
      .def("heap_bytes", &PeerComm::heap_bytes)
      .def("heap_tensor", &PeerComm::heap_tensor)
      .def("heap_ptr", &PeerComm::heap_ptr)
      .def("export_handles", &PeerComm::export_handles)
      .def("import_handles", &PeerComm::import_handles)
      .def_static("link_local", &PeerComm::link_local)
      .def("set_multicast", &PeerComm::set_multicast)
      .def("symm_bytes", &PeerComm::symm_bytes)
      .def("allreduce", &PeerComm::allreduce, py::arg("grad"), py::arg("algo"), py::arg("wire_bf16"),
           py::arg("scale"), py::arg("live_blocks") = py::none())
      .def("allreduce_adam", &PeerComm::allreduce_adam)
      .def("zero1_step", &PeerComm::zero1_step)
      .def("zero1_shard", &PeerComm::zero1_shard)
      .def("blocks_for", &PeerComm::blocks_for)
      .def("set_block_cap", &PeerComm::set_block_cap)
      .def("barrier", &PeerComm::barrier)
      .def("error", &PeerComm::error);
}

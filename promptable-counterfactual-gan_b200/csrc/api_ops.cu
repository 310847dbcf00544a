// extern "C" wrappers of the primitive operators used by the Python-composed step plans
// (pcg_b200/ops.py).  fp32, row-major / NHWC; see include/pcg.h for the contracts.
#include "../../include/pcg.h"

#include <vector>

#include "common.cuh"
#include "conv_auto.cuh"
#include "conv_c1k4.cuh"
#include "linear_small.cuh"
#include "conv_generic.cuh"
#include "elementwise.cuh"
#include "ops.cuh"
#include "instnorm.cuh"
#include "film_layer.cuh"
#include "frozen_mlp.cuh"

using namespace pcg;

#define PCG_API_BEGIN try {
#define PCG_API_END                                   \
  return 0;                                           \
  }                                                   \
  catch (const pcg::Error& e) {                       \
    pcg::set_last_error(e.what());                    \
    return e.code;                                    \
  }                                                   \
  catch (const std::exception& e) {                   \
    pcg::set_last_error(e.what());                    \
    return 99;                                        \
  }
#define ST ((cudaStream_t)stream)

// A/B switch: PCG_SKINNY=0 keeps the one-channel / one-output layers on the generic kernels
static bool skinny_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PCG_SKINNY"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v == 1;
}
static ConvGeom geom(int N, int H, int W, int Cin, int Cout, int k, int stride, int pad) {
  return ConvGeom{N, H, W, Cin, Cout, k, stride, pad};
}

extern "C" {

int pcg_conv_fprop(const float* in, int N, int H, int W, int Cin, const float* wf, int Cout, int k, int stride, int pad,
                   const float* bias, int act, float slope, const float* add_src, float* out, void* stream) {
  PCG_API_BEGIN
  GenEpilogue<float> e;
  e.bias = bias; e.act = act; e.slope = slope; e.add_src = add_src;
  if (bias == nullptr && add_src == nullptr && skinny_on() && c1k4_supported(geom(N, H, W, Cin, Cout, k, stride, pad))) {
    c1k4_fprop(in, geom(N, H, W, Cin, Cout, k, stride, pad), wf, act, slope, out, ST,
               conv_auto_cache_producer(out, (long long)N * (H / 2) * (W / 2) * Cout));
    return 0;
  }
  if (skinny_on() && !conv_auto_tensor_cores() && linear_small_supported(geom(N, H, W, Cin, Cout, k, stride, pad), N)) {
    linear_small(in, N, Cin, Cout, wf, e, out, ST);                         // wf of a 1x1 layer is [Cout][Cin]
    return 0;
  }
  if (act == ACT_NONE && add_src == nullptr && skinny_on() && full1_supported(geom(N, H, W, Cin, Cout, k, stride, pad))) {
    full1_fprop(in, geom(N, H, W, Cin, Cout, k, stride, pad), wf, bias, out, ST);
    return 0;
  }
  if (conv_fprop_auto(in, geom(N, H, W, Cin, Cout, k, stride, pad), wf, e, out, ST)) return 0;
  conv_fprop_generic<float, float>(in, geom(N, H, W, Cin, Cout, k, stride, pad), wf, e, out, ST);
  PCG_API_END
}
int pcg_conv_dgrad(const float* dout, int N, int H, int W, int Cin, const float* wd, int Cout, int k, int stride, int pad,
                   const float* add_src, const float* act_ref, int ref_act, float ref_slope, float* din, void* stream) {
  PCG_API_BEGIN
  GenEpilogue<float> e;
  e.add_src = add_src; e.act_ref = act_ref; e.ref_act = ref_act; e.ref_slope = ref_slope;
  if (add_src == nullptr && (act_ref == nullptr || ref_act == ACT_NONE) &&
      skinny_on() && c1k4_supported(geom(N, H, W, Cin, Cout, k, stride, pad))) {
    c1k4_dgrad(dout, geom(N, H, W, Cin, Cout, k, stride, pad), wd, din, ST);
    return 0;
  }
  if (skinny_on() && !conv_auto_tensor_cores() && linear_small_supported(geom(N, H, W, Cin, Cout, k, stride, pad), N)) {
    linear_small(dout, N, Cout, Cin, wd, e, din, ST);                       // wd of a 1x1 layer is [Cin][Cout]: "[out][in]" here
    return 0;
  }
  if (skinny_on() &&
      full_window_dgrad_supported(geom(N, H, W, Cin, Cout, k, stride, pad))) {
    // ConvTranspose2d on a 1x1 input: [N x Cout] x [Cout x (tap, Cin)], weight rows permuted from wd's [Cin][taps][Cout]
    linear_small(dout, N, Cout, k * k * Cin, wd, e, din, ST, Cin, k * k);
    return 0;
  }
  if (add_src == nullptr && (act_ref == nullptr || ref_act == ACT_NONE) && Cin % 4 == 0 &&
      skinny_on() && full1_supported(geom(N, H, W, Cin, Cout, k, stride, pad))) {
    full1_dgrad(dout, geom(N, H, W, Cin, Cout, k, stride, pad), wd, din, ST);
    return 0;
  }
  if (conv_dgrad_auto(dout, geom(N, H, W, Cin, Cout, k, stride, pad), wd, e, din, ST)) return 0;
  conv_dgrad_generic<float, float>(dout, geom(N, H, W, Cin, Cout, k, stride, pad), wd, e, din, ST);
  PCG_API_END
}
long long pcg_conv_wgrad_scratch(int N, int H, int W, int Cin, int Cout, int k, int stride, int pad) {
  const size_t generic = conv_wgrad_generic_scratch(geom(N, H, W, Cin, Cout, k, stride, pad));
  const size_t c1 = skinny_on() && c1k4_supported(geom(N, H, W, Cin, Cout, k, stride, pad)) ? c1k4_wgrad_scratch() : 0;
  return (long long)(generic > c1 ? generic : c1);
}
int pcg_conv_wgrad(const float* in, const float* dout, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                   float* scratch, float* dw, void* stream) {
  PCG_API_BEGIN
  if (skinny_on() && c1k4_supported(geom(N, H, W, Cin, Cout, k, stride, pad))) {
    c1k4_wgrad(in, dout, geom(N, H, W, Cin, Cout, k, stride, pad), scratch, dw, ST);
    return 0;
  }
  // (the weight-gradient kernel parallelises over columns only: it needs K >= 4096 to fill the GPU; a critic's last
  // Linear(128, 1) over 4096 samples stays on the generic split-K kernel)
  if (Cin % 4 == 0 && skinny_on() && full1_supported(geom(N, H, W, Cin, Cout, k, stride, pad)) &&
      geom(N, H, W, Cin, Cout, k, stride, pad).K() >= 4096) {
    full1_wgrad(in, dout, geom(N, H, W, Cin, Cout, k, stride, pad), dw, ST);
    return 0;
  }
  if (conv_wgrad_auto(in, dout, geom(N, H, W, Cin, Cout, k, stride, pad), dw, ST)) return 0;
  conv_wgrad_generic<float, float>(in, dout, geom(N, H, W, Cin, Cout, k, stride, pad), scratch, dw, ST);
  PCG_API_END
}
int pcg_pack_conv_weights(const float* w, int Cout, int Cin, int k, int perm_hw, float* wf, float* wd, void* stream) {
  PCG_API_BEGIN
  pack_conv_weights_generic(w, Cout, Cin, k, perm_hw, wf, wd, ST);
  PCG_API_END
}
int pcg_colsum(const float* a, long long M, int C, float* scratch, float* out, void* stream) {
  PCG_API_BEGIN
  colsum_partial<float>(a, M, C, scratch, ST);
  colsum_finalize(scratch, STAT_PARTS, C, C, out, ST);
  PCG_API_END
}
int pcg_set_conv_tensor_cores(int on) {
  conv_auto_set_tensor_cores(on != 0);
  return 0;
}
int pcg_set_operand_cache(int on) {
  const int prev = conv_auto_operand_cache() ? 1 : 0;
  conv_auto_set_operand_cache(on != 0);
  return prev;
}
void pcg_operand_cache_clear(void) { conv_auto_cache_clear(); }
void pcg_operand_cache_invalidate(const void* p, long long bytes) { conv_auto_cache_invalidate(p, (size_t)bytes); }
int pcg_get_conv_tensor_cores(void) { return conv_auto_tensor_cores() ? 1 : 0; }
int pcg_set_conv_tensor_core_terms(int terms) {
  const int prev = conv_auto_terms();
  conv_auto_set_terms(terms);
  return prev;
}
long long pcg_stat_scratch_floats(int C) { return (long long)STAT_PARTS * 2 * C; }

int pcg_bn_train_fwd(const float* y, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var, long long* nbt, float* mean, float* rstd, float* scale,
                     float* shift, int act, float slope, float* z, float* scratch, void* stream) {
  PCG_API_BEGIN
  if (skinny_on() && bn_cluster_supported(M, C)) {
    bn_cluster_fwd(y, M, C, gamma, beta, eps, momentum, running_mean, running_var, nbt, mean, rstd, scale, shift, act, slope, z,
                   ST);
    return 0;
  }
  bn_stats_partial<float>(y, M, C, scratch, ST);
  bn_finalize(scratch, STAT_PARTS, M, C, gamma, beta, eps, momentum, running_mean, running_var, nbt, mean, rstd, scale,
              shift, ST);
  // if a tensor-core consumer keeps a bf16 conversion of z (operand cache), write it here instead of in a pass of its own
  bn_apply_act<float>(y, scale, shift, M, C, act, slope, z, ST, conv_auto_cache_producer(z, M * C));
  PCG_API_END
}
int pcg_bn_train_bwd(const float* dz, const float* y, long long M, int C, const float* gamma, const float* mean,
                     const float* rstd, const float* scale, const float* shift, float gscale, int act, float slope,
                     float* dy, float* dgamma, float* dbeta, float* dbias_prev, float* c12, float* scratch,
                     float* scratch2, void* stream) {
  PCG_API_BEGIN
  if (skinny_on() && bn_cluster_supported(M, C)) {
    bn_cluster_bwd(dz, y, M, C, gamma, mean, rstd, scale, shift, gscale, act, slope, dy, dgamma, dbeta, dbias_prev, ST);
    return 0;
  }
  bn_bwd_partial<float>(dz, y, mean, rstd, scale, shift, gscale, act, slope, M, C, scratch, ST);
  bn_bwd_finalize(scratch, STAT_PARTS, M, C, dgamma, dbeta, c12, ST);
  bn_bwd_apply<float>(dz, y, mean, rstd, scale, shift, gamma, c12, gscale, act, slope, M, C, dy, scratch2, ST,
                      conv_auto_cache_producer(dy, M * C));
  if (dbias_prev) colsum_finalize(scratch2, STAT_PARTS, C, C, dbias_prev, ST);
  PCG_API_END
}
int pcg_bn_eval(const float* x, long long rows, int C, const float* gamma, const float* beta, const float* rm,
                const float* rv, float eps, float* y, float* scale_out, void* stream) {
  PCG_API_BEGIN
  bn_eval(x, rows, C, gamma, beta, rm, rv, eps, y, scale_out, ST);
  PCG_API_END
}
int pcg_scale_cols(const float* dy, long long rows, int C, const float* scale, float* dx, void* stream) {
  PCG_API_BEGIN
  scale_cols(dy, rows, C, scale, dx, ST);
  PCG_API_END
}
int pcg_unary(const float* x, long long n, int op, float a, float* y, void* stream) {
  PCG_API_BEGIN
  unary(x, n, op, a, y, ST);
  PCG_API_END
}
int pcg_unary_bwd(const float* dy, const float* y, long long n, int op, float a, float* dx, void* stream) {
  PCG_API_BEGIN
  unary_bwd(dy, y, n, op, a, dx, ST);
  PCG_API_END
}
int pcg_binary(const float* a, const float* b, long long n, int op, float alpha, float beta, float* out, void* stream) {
  PCG_API_BEGIN
  binary(a, b, n, op, alpha, beta, out, ST);
  PCG_API_END
}
int pcg_film_fwd(const float* gamma, const float* n_, const float* beta, const float* res, long long n, int relu,
                 float* out, void* stream) {
  PCG_API_BEGIN
  film_fwd(gamma, n_, beta, res, n, relu, out, ST);
  PCG_API_END
}
int pcg_film_bwd(const float* df, const float* gamma, const float* n_, long long n, int accumulate, float* dn,
                 float* dgamma, float* dbeta, void* stream) {
  PCG_API_BEGIN
  film_bwd(df, gamma, n_, n, accumulate, dn, dgamma, dbeta, ST);
  PCG_API_END
}
int pcg_transpose_multi(int n, const float* const* src, float* const* dst, const int* rows, const int* cols, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(n >= 1 && n <= TRANSPOSE_MAX, "1..64 matrices per call");
  TransposeTable t;
  t.n = n;
  int off = 0;
  for (int i = 0; i < n; ++i) {
    t.src[i] = src[i]; t.dst[i] = dst[i]; t.rows[i] = rows[i]; t.cols[i] = cols[i]; t.begin[i] = off;
    off += rows[i] * cols[i];
  }
  t.begin[n] = off;
  transpose_multi(t, ST);
  PCG_API_END
}
int pcg_copy_cols(const float* src, int src_ld, int c0_src, float* dst, int dst_ld, int c0_dst, long long rows, int ncols,
                  float alpha, int accumulate, void* stream) {
  PCG_API_BEGIN
  copy_cols(src, src_ld, c0_src, dst, dst_ld, c0_dst, rows, ncols, alpha, accumulate, ST);
  PCG_API_END
}
int pcg_onehot(const long long* lab, long long rows, int nc, float* dst, int dst_ld, int c0, void* stream) {
  PCG_API_BEGIN
  onehot(lab, rows, nc, dst, dst_ld, c0, ST);
  PCG_API_END
}
int pcg_reduce_scalar(const float* x, long long n, int absval, float scale, float* out, float gscale, float* dx,
                      void* stream) {
  PCG_API_BEGIN
  reduce_scalar(x, n, absval, scale, out, gscale, dx, ST);
  PCG_API_END
}
int pcg_rownorm_mean(const float* x, long long rows, int cols, int p, float* out, float gscale, float* dx, void* stream) {
  PCG_API_BEGIN
  rownorm_mean(x, rows, cols, p, out, gscale, dx, ST);
  PCG_API_END
}
int pcg_gan_loss(const float* z, int n, int kind, float t, float wgt, float* out_loss, float* out_aux, float* dz,
                 void* stream) {
  PCG_API_BEGIN
  gan_loss(z, n, kind, t, wgt, out_loss, out_aux, dz, ST);
  PCG_API_END
}
int pcg_combine_scalars(int n, const float* coeffs, const float* const* ptrs, float* out, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(n >= 1 && n <= 6, "1..6 terms");
  ScalarTerms t;
  t.n = n;
  for (int i = 0; i < n; ++i) { t.c[i] = coeffs[i]; t.p[i] = ptrs[i]; }
  combine_scalars(t, out, ST);
  PCG_API_END
}
int pcg_mlp_gan_step(int B, int z_dim, int label_dim, int hidden, const float* real, const float* real_oh,
                     const float* z1, const float* oh1, const float* z2, const float* oh2, float* g_param, float* g_grad,
                     float* g_m, float* g_v, int* g_step, float* d_param, float* d_grad, float* d_m, float* d_v,
                     int* d_step, float lr, float* scal, void* stream) {
  PCG_API_BEGIN
  mlp_gan_step(B, z_dim, label_dim, hidden, real, real_oh, z1, oh1, z2, oh2, g_param, g_grad, g_m, g_v, g_step, d_param,
               d_grad, d_m, d_v, d_step, lr, scal, ST);
  PCG_API_END
}
int pcg_spectral_norm_fwd(const float* W, int N, int K, float* u, float* v, float eps, int do_iter, float* Wn,
                          float* sigma, void* stream) {
  PCG_API_BEGIN
  spectral_norm_fwd(W, N, K, u, v, eps, do_iter, Wn, sigma, ST);
  PCG_API_END
}
int pcg_spectral_norm_fwd2(const float* W, int N, int K, float* u, float* v, float eps, int do_iter, float* Wn,
                           float* WnT, float* us, float* vs, float* sigma, void* stream) {
  PCG_API_BEGIN
  spectral_norm_fwd(W, N, K, u, v, eps, do_iter, Wn, sigma, ST, WnT, us, vs);
  PCG_API_END
}
int pcg_spectral_norm_bwd(const float* dWn, const float* Wn, int N, int K, const float* u, const float* v,
                          const float* sigma, float* dW, void* stream) {
  PCG_API_BEGIN
  spectral_norm_bwd(dWn, Wn, N, K, u, v, sigma, dW, ST);
  PCG_API_END
}
int pcg_gumbel_softmax_fwd(const float* logits, const float* g, long long rows, int n, float tau, float* y, void* stream) {
  PCG_API_BEGIN
  gumbel_softmax_fwd(logits, g, rows, n, tau, y, ST);
  PCG_API_END
}
int pcg_onehot_argmax(const float* x, long long rows, int n, float* y, void* stream) {
  PCG_API_BEGIN
  onehot_argmax(x, rows, n, y, ST);
  PCG_API_END
}
int pcg_softmax_bwd(const float* dy, const float* y, long long rows, int n, float tau, float* dl, void* stream) {
  PCG_API_BEGIN
  softmax_bwd(dy, y, rows, n, tau, dl, ST);
  PCG_API_END
}
int pcg_ce_loss(const float* logits, const long long* target, int B, int NC, float wgt, float* loss, float* dlogits,
                void* stream) {
  PCG_API_BEGIN
  ce_loss(logits, target, B, NC, wgt, loss, dlogits, ST);
  PCG_API_END
}
int pcg_ce_loss_weighted(const float* logits, const long long* target, const float* class_weights, int B, int NC,
                         float* loss, float* dlogits, float* correct, void* stream) {
  PCG_API_BEGIN
  ce_loss_weighted(logits, target, class_weights, B, NC, loss, dlogits, correct, ST);
  PCG_API_END
}
int pcg_adamw_flat(float* p, const float* g, float* m, float* v, long long n, int* step, const float* lr_dev, float beta1,
                   float beta2, float eps, float weight_decay, void* stream) {
  PCG_API_BEGIN
  adamw_flat(p, g, m, v, n, step, lr_dev, beta1, beta2, eps, weight_decay, ST);
  PCG_API_END
}
int pcg_u8_batch(const unsigned char* images, const long long* labels, const long long* index, int B, int HW, float mean,
                 float stdv, float* x, long long* y, void* stream) {
  PCG_API_BEGIN
  u8_batch(images, labels, index, B, HW, mean, stdv, x, y, ST);
  PCG_API_END
}
int pcg_instnorm_fwd(const float* x, int N, int P, int C, const float* gamma, const float* beta, float eps, int act, float slope,
                     float* y, float* mean, float* rstd, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(x && gamma && beta && y && mean && rstd, "instnorm_fwd: null pointer");
  instnorm_fwd(x, N, P, C, gamma, beta, eps, act, slope, y, mean, rstd, ST);
  PCG_API_END
}
int pcg_instnorm_bwd(const float* gy, const float* act_ref, int act, float slope, const float* x, const float* mean,
                     const float* rstd, const float* gamma, int N, int P, int C, const float* add_src, float* dx,
                     float* dgamma_part, float* dbeta_part, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(gy && x && mean && rstd && gamma && dx, "instnorm_bwd: null pointer");
  instnorm_bwd(gy, act_ref, act, slope, x, mean, rstd, gamma, N, P, C, add_src, dx, dgamma_part, dbeta_part, ST);
  PCG_API_END
}
int pcg_instnorm_bwd_bwd(const float* q, const float* gy, const float* act_ref, int act, float slope, const float* x,
                         const float* mean, const float* rstd, const float* gamma, int N, int P, int C, float* gy_bar,
                         float* x_bar, float* dgamma_part, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(q && gy && x && mean && rstd && gamma && gy_bar && x_bar, "instnorm_bwd_bwd: null pointer");
  instnorm_bwd_bwd(q, gy, act_ref, act, slope, x, mean, rstd, gamma, N, P, C, gy_bar, x_bar, dgamma_part, ST);
  PCG_API_END
}
int pcg_flatten_nchw(const float* src, int B, int R, int C, float* dst, int ld, int c0, int inverse, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(src && dst, "flatten_nchw: null pointer");
  flatten_nchw(src, B, R, C, dst, ld, c0, inverse != 0, ST);
  PCG_API_END
}
long long pcg_linear_wgrad_small_scratch(long long M, int K, int N) {
  return linear_wgrad_small_supported(M, K, N) ? linear_wgrad_small_scratch(M, K, N) : -1;
}
int pcg_linear_wgrad_small(const float* x, const float* dy, long long M, int K, int N, float* scratch, float* dw, float* db,
                           void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(x && dy && scratch && dw, "linear_wgrad_small: null pointer");
  linear_wgrad_small(x, dy, M, K, N, scratch, dw, db, ST);
  PCG_API_END
}
int pcg_frozen_mlp_parts(int L, const int* dims, int B) { return frozen_mlp_supported(L, dims) ? frozen_mlp_parts(B) : -1; }
int pcg_frozen_mlp_ce_grad(int L, const int* dims, const float* const* W, const float* const* WT, const float* const* b,
                           float slope, const float* x, const long long* target, int loss_kind, int B, float wgt,
                           float* logits, float* loss_part, float* dx, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(dims && W && WT && b && x && loss_part && dx && B >= 1, "frozen_mlp_ce_grad: null pointer / empty batch");
  frozen_mlp_ce_grad(L, dims, W, WT, b, slope, x, target, loss_kind, B, wgt, logits, loss_part, dx, ST);
  PCG_API_END
}
int pcg_mlp_fwd_bwd(int L, const int* dims, const float* const* W, const float* const* WT, const float* const* b, float slope,
                    const float* x, const long long* target, int loss_kind, int B, float wgt, float* logits, float* loss_part,
                    float* dx, float* const* act_out, float* const* grad_out, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(dims && W && WT && b && x && loss_part && dx && act_out && grad_out && B >= 1, "mlp_fwd_bwd: null pointer / empty batch");
  frozen_mlp_ce_grad(L, dims, W, WT, b, slope, x, target, loss_kind, B, wgt, logits, loss_part, dx, ST, act_out, grad_out);
  PCG_API_END
}
int pcg_film_layer_supported(long long M, int H) { return film_layer_supported(M, H) ? 1 : 0; }
int pcg_film_layer_fwd(const float* x, long long M, int H, const float* W, const float* bias, const float* gamma,
                       const float* beta, float eps, float momentum, float* running_mean, float* running_var, long long* nbt,
                       float* mean, float* rstd, float* scale, float* shift, const float* fg, const float* fb,
                       const float* res, int relu, float* u, float* n, float* out, float* scratch, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(x && W && bias && gamma && beta && mean && rstd && scale && shift && fg && fb && u && n && out && scratch,
              "film_layer_fwd: null pointer");
  film_layer_fwd(x, M, H, W, bias, gamma, beta, eps, momentum, running_mean, running_var, nbt, mean, rstd, scale, shift, fg,
                 fb, res, relu != 0, u, n, out, scratch, ST);
  PCG_API_END
}
int pcg_film_chain_fwd(const float* x, long long M, int H, int n, const pcg_film_half* halves, float eps, float momentum,
                       void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(x && halves && n >= 1 && n <= 64, "film_chain_fwd: null pointer / 1..64 half blocks");
  std::vector<FilmHalfFwd> hs(n);
  for (int k = 0; k < n; ++k) {
    const pcg_film_half& p = halves[k];
    PCG_REQUIRE(p.W && p.bias && p.gamma && p.beta && p.mean && p.rstd && p.scale && p.shift && p.fg && p.fb && p.u && p.n &&
                    p.out && p.scratch, "film_chain_fwd: null pointer in a half block");
    hs[k] = FilmHalfFwd{p.W, p.bias, p.gamma, p.beta, p.running_mean, p.running_var, p.nbt, p.mean, p.rstd, p.scale, p.shift,
                        p.fg, p.fb, p.res, p.res == nullptr ? 1 : 0, p.u, p.n, p.out, p.scratch};
  }
  film_chain_fwd(x, M, H, n, hs.data(), eps, momentum, ST);
  PCG_API_END
}
int pcg_film_layer_bwd(const float* d_f, long long M, int H, const float* fg, const float* n, const float* u,
                       const float* mean, const float* rstd, const float* gamma, const float* W, const float* add_src,
                       const float* act_ref, int accumulate, float* dfg, float* dfb, float* du, float* dx, float* dgamma,
                       float* dbeta, float* scratch, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(d_f && fg && n && u && mean && rstd && gamma && W && dfg && dfb && du && dx && dgamma && dbeta && scratch,
              "film_layer_bwd: null pointer");
  film_layer_bwd(d_f, M, H, fg, n, u, mean, rstd, gamma, W, add_src, act_ref, accumulate != 0, dfg, dfb, du, dx, dgamma,
                 dbeta, scratch, ST);
  PCG_API_END
}
int pcg_bias_act(const float* x, long long rows, int C, const float* bias, int tanh_out, float* y, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(x && bias && y && rows >= 0 && C >= 1, "bias_act: null pointer / shape");
  bias_act(x, rows, C, bias, tanh_out, y, ST);
  PCG_API_END
}
int pcg_dilate(const float* src, int N, int Ho, int Wo, int C, int stride, int off, int Hp, int Wp, float* dst, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(src && dst, "dilate: null pointer");
  dilate(src, N, Ho, Wo, C, stride, off, Hp, Wp, dst, ST);
  PCG_API_END
}
int pcg_pack_dgrad_classes(const float* w, int Cout, int Cin, int k, float* wc, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(w && wc, "pack_dgrad_classes: null pointer");
  pack_dgrad_classes(w, Cout, Cin, k, wc, ST);
  PCG_API_END
}
int pcg_parity_interleave(const float* src, int N, int Hc, int Wc, int C, int pad, int H, int W, int stacked, float* dx,
                          void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(src && dx, "parity_interleave: null pointer");
  parity_interleave(src, N, Hc, Wc, C, pad, H, W, stacked != 0, dx, ST);
  PCG_API_END
}
int pcg_gp_penalty(const float* g, int B, int D, float lambda, float* out, float* gbar, float* norms, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(g && out && gbar, "gp_penalty: null pointer");
  gp_penalty(g, B, D, lambda, out, gbar, norms, ST);
  PCG_API_END
}
int pcg_cf_scratch_floats(void) { return cf_parts(); }
int pcg_cf_apply(const float* x, const float* residual, long long n, float lo, float hi, float* x_cf, float* scratch,
                 void* stream) {
  PCG_API_BEGIN
  cf_apply(x, residual, n, lo, hi, x_cf, scratch, ST);
  PCG_API_END
}
int pcg_cf_metrics(const float* logits, const long long* y_true, const long long* y_target, int B, int NC,
                   const float* scratch, long long n_elems, float* out3, void* stream) {
  PCG_API_BEGIN
  cf_metrics(logits, y_true, y_target, B, NC, scratch, n_elems, out3, ST);
  PCG_API_END
}
int pcg_dropout_mask(long long rows, int inner, int C, float p, int channelwise, unsigned long long seed,
                     unsigned long long* rng_state, float* mask, void* stream) {
  PCG_API_BEGIN
  dropout_mask(rows, inner, C, p, channelwise, seed, rng_state, mask, ST);
  PCG_API_END
}
int pcg_build_mask(int B, int C, int H, int W, int patch, int num_modifiable_patches, int num_classes,
                   unsigned long long seed, unsigned long long* rng_state, float* mask, long long* target, void* stream) {
  PCG_API_BEGIN
  build_mask(B, C, H, W, patch, num_modifiable_patches, num_classes, seed, rng_state, mask, target, ST);
  PCG_API_END
}
int pcg_adam_flat(float* p, const float* g, float* m, float* v, long long n, int* step, float lr, float beta1,
                  float beta2, float eps, float grad_scale, void* stream) {
  PCG_API_BEGIN
  adam_flat(p, g, m, v, n, step, lr, beta1, beta2, eps, grad_scale, ST);
  PCG_API_END
}

}  // extern "C"

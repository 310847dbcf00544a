// HBM-bound kernels of the GAN step: BatchNorm (train mode) forward/backward, input assembly,
// residual head, discriminator head, losses, embedding gradient, fused Adam — host interface.
#pragma once
#include "common.cuh"

namespace pcg {

constexpr int STAT_PARTS = 592;   // 4 x 148 row-slices (4 resident 256-thread blocks per SM) for two-stage column reductions

// ---- BatchNorm2d, train mode (generator.py:12,15; torch native_batch_norm) ------------------
// Column statistics of y[M][C]: part[STAT_PARTS][2*C] (sum, sum of squares) per row slice.
template <typename T>
void bn_stats_partial(const T* y, long long M, int C, float* part, cudaStream_t s);
// Reduces `nparts` partial rows in fp64, writes mean/rstd (saved for backward), scale/shift
// (a = gamma*rstd, b = beta - mean*a) and updates running_mean / running_var (momentum 0.1,
// unbiased variance) and num_batches_tracked (int64) in place.
void bn_finalize(const float* part, int nparts, long long M, int C, const float* gamma, const float* beta,
                 float eps, float momentum, float* running_mean, float* running_var, long long* nbt,
                 float* mean, float* rstd, float* scale, float* shift, cudaStream_t s);
// z = act(scale*y + shift)
template <typename T>
void bn_apply_act(const T* y, const float* scale, const float* shift, long long M, int C, int act, float slope,
                  T* z, cudaStream_t s, bf16* side = nullptr);     // side: optional bf16 copy of z (operand cache)
// out = h + res_scale * (scale*y + shift)          (generator.py:22)
template <typename T>
void bn_apply_residual(const T* y, const T* h, const float* scale, const float* shift, float res_scale,
                       long long M, int C, T* out, cudaStream_t s);
// Backward, pass 1: g = gscale * dsrc * act'(scale*y+shift);  part[.][2C] = (sum g, sum g*xhat)
template <typename T>
void bn_bwd_partial(const T* dsrc, const T* y, const float* mean, const float* rstd, const float* scale,
                    const float* shift, float gscale, int act, float slope, long long M, int C, float* part,
                    cudaStream_t s);
// Backward, pass 2 prologue: totals -> dgamma, dbeta, and the two per-channel means (c1, c2).
void bn_bwd_finalize(const float* part, int nparts, long long M, int C, float* dgamma, float* dbeta, float* c12,
                     cudaStream_t s);
// Backward, pass 2: dy = gamma*rstd * (g - c1 - xhat*c2); also part_db[.][C] = column sums of dy
// (gradient of the conv bias in front of the BatchNorm).
template <typename T>
void bn_bwd_apply(const T* dsrc, const T* y, const float* mean, const float* rstd, const float* scale,
                  const float* shift, const float* gamma, const float* c12, float gscale, int act, float slope,
                  long long M, int C, T* dy, float* part_db, cudaStream_t s, bf16* side = nullptr);
// out[C] = sum over `nparts` rows of part[.][stride] (first C columns), fixed order, fp64 accumulate.
void colsum_finalize(const float* part, int nparts, int stride, int C, float* out, cudaStream_t s);
struct ColsumOuts { static constexpr int MAX = 16; float* p[MAX]; };
void colsum_finalize_multi(const float* part, int nparts, int C, const ColsumOuts& outs, int nslots, cudaStream_t s);
// part[STAT_PARTS][C] = column sums of a[M][C]
template <typename T>
void colsum_partial(const T* a, long long M, int C, float* part, cudaStream_t s);

// ---- input assembly (generator.py:73-74, discriminator.py:35-36) ---------------------------------
// out[n][p][0..2] = x[n][p], embed[label[n]][p], mask[n][p]
template <typename T>
void g_input(const float* x, const float* embed, const long long* label, const float* mask, int B, int HW, T* out,
             cudaStream_t s);
// out[n][p][0..1] = x[n][p], embed[label[n]][p]
template <typename T>
void d_input(const float* x, const float* embed, const long long* label, int B, int HW, T* out, cudaStream_t s);
// dE[cls][p] = sum_{n : label[n] == cls} src[n][p][ch]   (src has `nch` channels); deterministic.
template <typename T>
void embed_grad(const T* src, int nch, int ch, const long long* label, int B, int HW, int num_classes, float* dE,
                cudaStream_t s);

// ---- residual head (generator.py:80-82, trainer.py:97,99,119) ------------------------------------
// raw = rs*c ; masked = raw*mask ; x_cf = clamp(x+masked, -1, 1);
// part[STAT_PARTS][2] = (sum |masked|, sum |raw*(1-mask)|)
void residual_head_fwd(const float* c, const float* x, const float* mask, float rs, long long n, float* raw,
                       float* masked, float* x_cf, float* part, cudaStream_t s);
// g_c = rs * ( mask * ( pass*(wd*dxd[.][0 of dxd_ch] + dxc) + lreg*sign(masked)/n ) + lmask*sign(raw*(1-mask))*(1-mask)/n )
//   pass = 1 where -1 <= x+masked <= 1 (clamp backward); dxd is the discriminator-path gradient
//   wrt its 2-channel input (channel 0 = image), dxc the classifier-path gradient.
template <typename T>
void residual_head_bwd(const float* dxd, int dxd_ch, const float* dxc, const float* raw, const float* x,
                       const float* mask, float rs, float lreg, float lmask, long long n, T* g_c, cudaStream_t s);

// ---- discriminator head + BCE-with-logits (discriminator.py:26-38, trainer.py:106-107,117) -----------
// logits[n] = b + sum_c w[c] * mean_hw z[n][hw][c]
template <typename T>
void d_head_fwd(const T* z, int B, int HW, int C, const float* w, const float* b, float* logits, cudaStream_t s);
// Segment i (i < nseg) covers samples [i*seg, (i+1)*seg) with target t[i] and weight wgt[i]:
//   out_loss[i] = mean BCEWithLogits, out_p[i] = mean sigmoid, dlogit[n] = wgt[i]*(sigmoid(z)-t)/seg
void bce_logits(const float* logits, int seg, int nseg, float t0, float t1, float w0, float w1, float* out_loss,
                float* out_p, float* dlogit, cudaStream_t s);
// g[n][hw][c] = dlogit[n]*w[c]/HW * lrelu'(z);  dw[c] = sum_n dlogit[n]*mean_hw z ; db = sum dlogit
// (dw/db only if dw != nullptr; `scratch` then holds 64*C floats of per-slice partial sums)
template <typename T>
void d_head_bwd(const T* z, const float* dlogit, int B, int HW, int C, const float* w, float slope, T* g, float* dw,
                float* db, float* scratch, cudaStream_t s);

// ---- classifier loss (trainer.py:118) ----------------------------------------------------------------
// loss = mean_n (logsumexp(l[n]) - l[n][t[n]]);  dl = wgt * (softmax - onehot) / B
void ce_loss(const float* logits, const long long* target, int B, int NC, float wgt, float* loss, float* dlogits,
             cudaStream_t s);

// ---- scalars ------------------------------------------------------------------------------------------
// out[0] = sum(part[.][0]) * inv_n ; out[1] = sum(part[.][1]) * inv_n
void l1_finalize(const float* part, int nparts, float inv_n, float* out2, cudaStream_t s);
// g_loss = la*g_adv + lc*g_cls + lr*reg + lm*mask_pen
void g_loss_combine(const float* g_adv, const float* g_cls, const float* reg, const float* mpen, float la, float lc,
                    float lr, float lm, float* out, cudaStream_t s);

// ---- Adam over a flat arena (torch/optim/adam.py:347-547; trainer.py:77-78,112,123) -------------------
// step counter lives on the device (`step`, int32, incremented by the kernel) so the launch can be
// captured in a CUDA graph.  grad is multiplied by grad_scale first (1/world_size for data parallel).
void adam_flat(float* p, const float* g, float* m, float* v, long long n, int* step, float lr, float beta1,
               float beta2, float eps, float grad_scale, cudaStream_t s);

template <typename T>
void fill_zero(T* p, long long n, cudaStream_t s);
// fp32 -> T copy (used to stage NCHW==NHWC single-channel images into the activation type)
template <typename T>
void convert_from_f32(const float* src, long long n, T* dst, cudaStream_t s);

void ce_loss_weighted(const float* logits, const long long* target, const float* w, int B, int NC, float* loss,
                      float* dlogits, float* correct, cudaStream_t s);
void adamw_flat(float* p, const float* g, float* m, float* v, long long n, int* step, const float* lr_dev, float beta1,
                float beta2, float eps, float weight_decay, cudaStream_t s);
// Train-mode BatchNorm of a small [M][C] problem as ONE cluster launch (bn_cluster.cu); same contract as the three-launch
// pipelines above.  bn_cluster_supported: C a power of two <= 256 and M small enough for the rows to stay in registers.
bool bn_cluster_supported(long long M, int C);
void bn_cluster_fwd(const float* y, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                    float* running_mean, float* running_var, long long* nbt, float* mean, float* rstd, float* scale,
                    float* shift, int act, float slope, float* z, cudaStream_t s);
void bn_cluster_bwd(const float* dz, const float* y, long long M, int C, const float* gamma, const float* mean,
                    const float* rstd, const float* scale, const float* shift, float gscale, int act, float slope, float* dy,
                    float* dgamma, float* dbeta, float* dbias_prev, cudaStream_t s);

}  // namespace pcg

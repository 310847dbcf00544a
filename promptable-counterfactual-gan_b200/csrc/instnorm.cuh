// InstanceNorm2d (affine) on NHWC activations [N][P][C], P = H*W: forward, backward and the backward OF the backward, which
// the gradient penalty of the conditional WGAN-GP critic needs (conditional_gan/mnist/mnist_wgan_conditional.py:87-95 the
// layers, :146-150 autograd.grad(create_graph=True)).  Statistics per (sample, channel) over P, biased variance.
#pragma once
#include "common.cuh"

namespace pcg {

// y = act(gamma * xhat + beta); mean / rstd [N][C] are saved for the backward passes
void instnorm_fwd(const float* x, int N, int P, int C, const float* gamma, const float* beta, float eps, int act, float slope,
                  float* y, float* mean, float* rstd, cudaStream_t s);
// p = gy * act'(act_ref) (act_ref = the layer's output y, NULL: p = gy);  dx = gamma * rstd * (p - mean(p) - xhat * mean(p * xhat))
// (+ add_src);  dgamma_part / dbeta_part [N][C] = per-sample sums (NULL: skipped), reduce over N with pcg_colsum
void instnorm_bwd(const float* gy, const float* act_ref, int act, float slope, const float* x, const float* mean,
                  const float* rstd, const float* gamma, int N, int P, int C, const float* add_src, float* dx,
                  float* dgamma_part, float* dbeta_part, cudaStream_t s);
// Given q = the cotangent of dx above: gy_bar (cotangent of gy), x_bar (cotangent of x through xhat and rstd) and the
// per-sample parts of gamma's cotangent.  beta and the activation argument get none (act'' = 0 almost everywhere).
void instnorm_bwd_bwd(const float* q, const float* gy, const float* act_ref, int act, float slope, const float* x,
                      const float* mean, const float* rstd, const float* gamma, int N, int P, int C, float* gy_bar,
                      float* x_bar, float* dgamma_part, cudaStream_t s);
// dst[b][c0 + c * R + r] = src[b][r][c] (rows of ld_dst floats): NHWC features -> torch's NCHW nn.Flatten order, written
// into a column window of a wider matrix;  inverse = true: src[b][r][c] layout is the DESTINATION (dst[b][r][c] = src[b][c0 + c*R + r])
void flatten_nchw(const float* src, int B, int R, int C, float* dst, int ld, int c0, bool inverse, cudaStream_t s);

// y[r][c] = f(x[r][c] + bias[c]); f: 0 identity, 1 tanh (the bias of a ConvTranspose2d computed as a data gradient)
void bias_act(const float* x, long long rows, int C, const float* bias, int tanh_out, float* y, cudaStream_t s);
// Zero-dilated, padded copy of an NHWC tensor: dst[n][off + s*y][off + s*x][:] = src[n][y][x][:], zero elsewhere
// ([N][Hp][Wp][C]).  The data gradient of a strided convolution is the stride-1 convolution of this tensor (off = k-1-pad,
// Hp = H + k - 1) with the tap-reversed weight (pcg_pack_conv_weights perm_hw = -1) - a forward convolution, which runs on
// the tcgen05 kernel for any kernel size / stride.
void dilate(const float* src, int N, int Ho, int Wo, int C, int stride, int off, int Hp, int Wp, float* dst, cudaStream_t s);
// Stride-2 data gradient as four stride-1 2x2 forward convolutions of the UNdilated gradient, one per parity class of the
// input position (no multiplications by the zeros a dilated gradient carries): pack_dgrad_classes builds their weights
// wc[4][Cin][2*2][Cout] from the torch OIHW weight (k = 3 or 4), parity_interleave merges the four class maps
// [4][N][Hc][Wc][C] (Hc = Ho + 1) - or, stacked, the result [N][Hc][Wc][4][C] of ONE convolution with all four weight sets
// (wc as a whole is the forward weight of a Cout -> 4 * Cin layer) - into dx [N][H][W][C].
void pack_dgrad_classes(const float* w, int Cout, int Cin, int k, float* wc, cudaStream_t s);
void parity_interleave(const float* src, int N, int Hc, int Wc, int C, int pad, int H, int W, bool stacked, float* dx,
                       cudaStream_t s);
// WGAN-GP penalty of mnist_wgan_conditional.py:147: n_b = ||g[b][:]||_2, out[0] = lambda * mean_b (n_b - 1)^2,
// gbar[b][:] = lambda * 2 (n_b - 1) / (B * n_b) * g[b][:] (the cotangent of g), norms[b] = n_b (optional)
void gp_penalty(const float* g, int B, int D, float lambda, float* out, float* gbar, float* norms, cudaStream_t s);

}  // namespace pcg

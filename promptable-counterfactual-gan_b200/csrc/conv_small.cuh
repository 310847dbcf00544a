// Convolutions whose channel count on one side is tiny (1-3): the generator's conv_in / conv_out, the first layer
// of the discriminator and of the classifier (conditional_counteRGAN/mnist/models/generator.py:39,50,
// discriminator.py:15, classifier.py:8).  They carry < 2 % of the step's FLOPs but read or write a full
// 64-channel activation map, so each must run at HBM speed: one TMA halo tile per block of output rows, the
// contraction on the (legacy, warp-level) tensor-core path straight out of the swizzled tile.
#pragma once
#include "common.cuh"

namespace pcg {

// out[N*H*W] (fp32) = bias + sum_{tap, c} in[n][h + r - 1][w + s - 1][c] * w9[tap = r*3+s][c]
//   in : NHWC bf16, Cin = 32 or 64;  w9 : bf16 [9][Cin]
// Serves conv_out's forward (w9 = its packed weights) and the single-input-channel data gradients of conv_in and of
// the classifier's first convolution (w9 = that channel's row of the rotated dgrad packing).
bool conv_to1_supported(int H, int W, int Cin);
void conv_to1(const bf16* in, int N, int H, int W, int Cin, const bf16* w9, const float* bias, float* out,
              cudaStream_t stream);

// out[M][Cout] (bf16 NHWC) = epi( sum_{tap, c} in[n][ho*stride - 1 + r][wo*stride - 1 + s][c] * wnk[co][tap*Cs + c] )
//   in  : NHWC with Cs = 1, 2 or 3 channels, bf16 or fp32;  wnk : bf16 [Cout][9*Cs];  Cout = 32 or 64; stride 1 or 2; pad 1
//   epi : v = acc + bias ; v = act(v) ; v *= act'(act_ref) (LeakyReLU/ReLU derivative from the sign of act_ref[M][Cout])
// Serves conv_in's and the discriminator's / classifier's first forward convolutions (wnk = the fprop packing) and
// conv_out's data gradient (Cs = 1, wnk = the rotated dgrad packing [ci][8 - tap]).
struct FewEpilogue {
  const float* bias = nullptr;
  int act = ACT_NONE;
  float slope = 0.2f;
  const bf16* act_ref = nullptr;
  int ref_act = ACT_NONE;
  float ref_slope = 0.2f;
};
bool conv_few_supported(int Cs, int Cout, int ksize, int stride, int pad);
template <typename TIn>
void conv_few(const TIn* in, int N, int H, int W, int Cs, const bf16* wnk, int Cout, int stride, const FewEpilogue& epi,
              bf16* out, cudaStream_t stream);

// Weight gradient of a convolution with Cs = 1..3 input channels and 64 output channels (3x3, pad 1, stride 1|2):
//   dw[co][c][r][s] (torch OIHW fp32) = sum_m dy[m][co] * in[n][ho*stride - 1 + r][wo*stride - 1 + s][c]
//   db[co] = sum_m dy[m][co]           (only if db != nullptr; costs nothing: a constant-1 im2col column)
// `part` must hold wgrad_few_parts() * 2048 floats.
int wgrad_few_parts();
bool wgrad_few_supported(int Cs, int Cout, int ksize, int stride, int pad);
template <typename TIn>
void wgrad_few(const TIn* in, const bf16* dy, int N, int H, int W, int Cs, int stride, float* part, float* dw, float* db,
               cudaStream_t stream);

// Weight gradient of a 64 -> 1 convolution (3x3, stride 1, pad 1):
//   dw[0][ci][r][s] = sum_p g[p] * x[n][h + r - 1][w + s - 1][ci];   db[0] = sum_p g[p]
//   x : NHWC bf16 [N*H*W][64];  g : bf16 [N*H*W].  `part` must hold wgrad_to1_parts() * 1088 floats.
int wgrad_to1_parts();
void wgrad_to1(const bf16* x, const bf16* g, int N, int H, int W, float* part, float* dw, float* db, cudaStream_t stream);

// One input channel of the data gradient of a 64-output-channel 3x3 / stride-2 / pad-1 convolution:
//   dx[n][hi][wi] (fp32) = sum_{r,s,co : hi+1-r, wi+1-s even} dy[n][(hi+1-r)/2][(wi+1-s)/2][co] * wrot[8 - (r*3+s)][co]
//   dy : NHWC bf16 [N][Ho][Wo][64] with Ho = H/2, Wo = W/2 (H, W even, Ho*Wo <= 208);  wrot : bf16 [9][64], that input
//   channel's row of the rotated dgrad packing ([ci][8 - tap][co]).
// Per image: P[pos][tap] = dy[pos][:] . w[:][tap] on the tensor path, then a col2im gather of <= 4 taps per pixel.
bool dgrad_s2_to1_supported(int H, int W, int Cout);
void dgrad_s2_to1(const bf16* dy, int N, int H, int W, const bf16* wrot, float* dx, cudaStream_t stream);

}  // namespace pcg

// CUDA-core (fp32 FFMA) convolution kernels on NHWC tensors — host interface.
// Used for every layer the tensor-core kernels do not cover (tiny channel counts, strided data
// gradients, the fully connected layers as 1x1 convolutions) and as the exact-fp32 mode of the step.
#pragma once
#include "common.cuh"

namespace pcg {

struct ConvGeom {
  int N, H, W, Cin;      // input  NHWC
  int Cout, ksize, stride, pad;
  int Ho() const { return (H + 2 * pad - ksize) / stride + 1; }
  int Wo() const { return (W + 2 * pad - ksize) / stride + 1; }
  long long Mout() const { return (long long)N * Ho() * Wo(); }
  long long Min() const { return (long long)N * H * W; }
  int K() const { return ksize * ksize * Cin; }
};

// Epilogue of fprop / dgrad:  v = acc (+bias) ; v = act(v) ; v += add_src ; v *= act'(act_ref)
template <typename TOut>
struct GenEpilogue {
  const float* bias = nullptr;
  int act = ACT_NONE;           // forward activation applied to acc + bias
  float slope = 0.2f;
  const TOut* add_src = nullptr;   // same shape as the output
  const TOut* act_ref = nullptr;   // same shape as the output: multiply by d act / d pre at act_ref
  int ref_act = ACT_NONE;          // which activation produced act_ref
  float ref_slope = 0.2f;
};

// out[Mout][Cout] = epi( im2col(in) * wf^T ),  wf fp32 [Cout][taps][Cin]
template <typename TIn, typename TOut>
void conv_fprop_generic(const TIn* in, const ConvGeom& g, const float* wf, const GenEpilogue<TOut>& epi,
                        TOut* out, cudaStream_t stream);

// din[Min][Cin] = epi( gather(dout) * wd^T ),  wd fp32 [Cin][taps][Cout] (taps NOT rotated)
template <typename TIn, typename TOut>
void conv_dgrad_generic(const TIn* dout, const ConvGeom& g, const float* wd, const GenEpilogue<TOut>& epi,
                        TOut* din, cudaStream_t stream, int ch_select = -1);
// ch_select >= 0 (Cin <= 4 only): compute just that input channel, compact [Min][1] output.

// dw (torch OIHW fp32 [Cout][Cin][k][k]) = sum_pixels dout[p][co] * in[p@tap][ci]
// `scratch` must hold conv_wgrad_generic_scratch(g) floats.  If db != nullptr also db[co] = sum dout.
size_t conv_wgrad_generic_scratch(const ConvGeom& g);
template <typename TIn, typename TDy>
void conv_wgrad_generic(const TIn* in, const TDy* dout, const ConvGeom& g, float* scratch, float* dw,
                        cudaStream_t stream);

// dw[co][ci][tap] (torch OIHW) = sum_z part[z][co][(tap, ci)], fixed order.
void wgrad_reduce_generic(const float* part, int nz, int Cout, int Cin, int taps, float* dw, cudaStream_t stream);

// torch OIHW fp32 -> wf [Cout][taps][Cin] and wd [Cin][taps][Cout] (either may be nullptr).
// `perm_hw` > 0: the Cin axis of a Linear that consumes a flattened NCHW [C][perm_hw] map is
// re-ordered to NHWC flatten order (classifier fc.1, classifier.py:17-18).
// dst[c][r] = src[r][c]; taps / rev: see conv_generic.cu (tap-reversed weight packing)
void transpose_tiled(const float* src, int R, int C, float* dst, cudaStream_t stream, int taps = 1, bool rev = false,
                     int batch = 1);      // batch: that many consecutive [R][C] matrices
void pack_conv_weights_generic(const float* w, int Cout, int Cin, int ksize, int perm_hw, float* wf,
                               float* wd, cudaStream_t stream);

}  // namespace pcg

// Tensor-core (tcgen05 / TMEM / TMA-im2col) convolution kernels for sm_100a — host interface.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace pcg {

// Finalisation of the per-CTA `stats` partial rows, run by the LAST CTA of conv_tc64_fprop to finish (an atomic ticket
// after its row is written and fenced) instead of by a separate one-block-per-channel launch: the rows are summed per
// column in CTA order in fp64 (deterministic), then
//   mode 1 (train-mode BatchNorm forward: rows = sum y, sum y^2): mean, rstd, scale = gamma*rstd, shift = beta - mean*scale,
//           running_mean / running_var (momentum, unbiased variance) and num_batches_tracked, as bn_finalize;
//   mode 2 (BatchNorm backward: rows = sum g, sum g*xhat): dbeta, dgamma, c12 = (sum g / M, sum g*xhat / M), as
//           bn_bwd_finalize.
struct StatsFinalize {
  int mode = 0;
  unsigned int* counter = nullptr;   // device, zero before the first use; the last CTA resets it
  long long M = 0;                   // elements per channel
  const float *gamma = nullptr, *beta = nullptr;
  float eps = 1e-5f, momentum = 0.1f;
  float *running_mean = nullptr, *running_var = nullptr;
  long long* nbt = nullptr;
  float *mean = nullptr, *rstd = nullptr, *scale = nullptr, *shift = nullptr;
  float *dgamma = nullptr, *dbeta = nullptr, *c12 = nullptr;
};

// Epilogue description shared by the fprop/dgrad implicit-GEMM kernel.
struct ConvEpilogue {
  const float* bias = nullptr;   // [Cout] fp32, added before activation
  int act = ACT_NONE;            // ACT_NONE / ACT_LRELU / ACT_RELU applied to (acc + bias)
  float slope = 0.2f;
  const bf16* add_src = nullptr; // [M][Cout] bf16 added after activation (residual / skip gradient)
  const bf16* act_ref = nullptr; // [M][Cout] bf16: result is multiplied by d act/d pre evaluated at
  int ref_act = ACT_NONE;        // act_ref (the activation OUTPUT; LReLU/ReLU derivative from its sign)
  float ref_slope = 0.2f;
  float* out_f32 = nullptr;      // conv_tc_fprop / conv_tc_dgrad_s2 only: store the result as fp32 here (same shape)
  float* stats = nullptr;        // [grid][2*Cout] fp32 per-CTA partial (sum, sum of squares) of the
                                 // pre-rounding output values, for train-mode BatchNorm
  // conv_tc64_fprop only - reduction pass of the BatchNorm backward that consumes this data gradient, fused into the
  // epilogue: with v = the result, y = bn_y[pixel][c], g = v * act'(scale*y + shift):
  //   stats[cta][c] = sum g,  stats[cta][64 + c] = sum g * (y - mean) * rstd        (bn_y == nullptr: off)
  const bf16* bn_y = nullptr;
  const float *bn_mean = nullptr, *bn_rstd = nullptr, *bn_scale = nullptr, *bn_shift = nullptr;
  int bn_act = ACT_NONE;
  float bn_slope = 0.2f;
  StatsFinalize fin;             // conv_tc64_fprop only: finish the statistics in the same launch (mode 0: off)
  float bn_gscale = 1.f;         // the reduction is taken of bn_gscale * v (BN2 of a residual block: 0.1, generator.py:22)
  // add_src may be combined with bn_y on the row-class kernel (H % 4 == 0): v = accumulator + add_src, reduced as above
};

// Number of CTAs conv_tc_fprop launches for a problem (rows of the `stats` partial buffer).
int conv_tc_grid(long long M, int Cout);

// out[M][Cout] (NHWC bf16) = epilogue( im2col(in)[M][taps*Cin] * wpk[Cout][taps*Cin]^T )
//   in  : NHWC bf16 [N][H][W][Cin], Cin % 64 == 0
//   wpk : bf16 [Cout][ksize*ksize][Cin] (tap-major, channel fastest), Cout % 64 == 0
// The data-gradient of a stride-1 convolution is the same call on dY with the rotated/transposed
// packing produced by pack_conv_weights_tc().
void conv_tc_fprop(const bf16* in, int N, int H, int W, int Cin, const bf16* wpk, int Cout, int ksize,
                   int stride, int pad, const ConvEpilogue& epi, bf16* out, cudaStream_t stream);

// Data gradient of a 3x3 / stride-2 / pad-1 convolution on the tensor cores.
//   dy bf16 NHWC [N][Ho][Wo][Cout] -> dx bf16 NHWC [N][H][W][Cin]; `packed` = 9*Cout*Cin bf16 written by
//   pack_dgrad_s2_tc (four parity-class matrices [Cin][taps_class][Cout], 1 + 2 + 2 + 4 taps).
// Four launches of the fprop kernel, one per (hi%2, wi%2) class: each is a stride-1 implicit GEMM over dY whose
// output rows scatter to every other pixel of dx.  Epilogue: act / add_src / act_ref as in fprop (no bias/stats).
size_t conv_tc_dgrad_s2_pack_elems(int Cout, int Cin);
void pack_dgrad_s2_tc(const float* w, int Cout, int Cin, bf16* packed, cudaStream_t stream);
// `class_streams` (optional, 4 entries): the four parity classes write disjoint pixels, so they may run on different
// streams; the caller orders those streams around the call.  conv_tc_dgrad_s2_class_ctas = CTAs of the largest class.
void conv_tc_dgrad_s2(const bf16* dy, int N, int H, int W, int Cin, int Cout, const bf16* packed,
                      const ConvEpilogue& epi, bf16* dx, cudaStream_t stream,
                      const cudaStream_t* class_streams = nullptr, int ksize = 3);
// 4x4 / stride 2 / pad 1 variant of the packing: wd = fp32 [Cin][16][Cout] (the CUDA-core dgrad layout), packed =
// 16*Cout*Cin bf16 (four [Cin][2x2][Cout] class matrices); use with conv_tc_dgrad_s2(..., ksize = 4).
// dup = 3: packed = 48*Cout*Cin bf16, every Cout run as [W_hi | W_hi | W_lo], for a dY split into hi | lo | hi (3*Cout channels).
void pack_dgrad_s2_k4_tc(const float* wd, int Cout, int Cin, bf16* packed, cudaStream_t stream, int dup = 1);
int conv_tc_dgrad_s2_class_ctas(int N, int H, int W, int Cin);

// dW partials of a 3x3 / stride-1 / pad-1 convolution with Cin = Cout = 64:
//   part[cta][tap][ci][co] (fp32) = sum over the CTA's pixels of x[p + tap][ci] * dy[p][co]
// `part` must hold conv_tc_wgrad_grid(M) * 9*64*64 floats; reduce with wgrad_reduce_tc().
int conv_tc_wgrad_grid(long long M);
void conv_tc_wgrad64(const bf16* x, const bf16* dy, int N, int H, int W, float* part,
                     cudaStream_t stream);
// dw[co][ci][3][3] (fp32, torch OIHW) = sum_cta part[cta][tap][ci][co];  db[co] = sum_p dy[p][co]
// is produced by the caller's BN/bias kernels, not here.
// swizzled: the partials come from conv_tc64_wgrad's staged write-out (conv_tc64_wgrad_swizzled())
void wgrad_reduce_tc(const float* part, int nparts, float* dw, cudaStream_t stream, bool swizzled = false);
bool conv_tc64_wgrad_swizzled();

// General tensor-core wgrad (3x3, pad 1, stride 1|2, Cin % 64 == 0, Cout % 64 == 0):
//   part[z][co][(tap, ci)] fp32, z < conv_tc_wgrad_general_splits(...); reduce with wgrad_reduce_generic().
int conv_tc_wgrad_general_splits(int N, int H, int W, int Cin, int Cout, int stride, int ksize = 3, int pad = 1);
void conv_tc_wgrad_general(const bf16* x, const bf16* dy, int N, int H, int W, int Cin, int Cout, int stride,
                           float* part, cudaStream_t stream, int ksize = 3, int pad = 1);

// fp32 OIHW [Cout][Cin][k][k] -> bf16 [Cout][k*k][Cin] (fprop) and, if dgrad != nullptr,
// bf16 [Cin][k*k][Cout] with the taps rotated by 180 degrees (dgrad of a stride-1 conv).
void pack_conv_weights_tc(const float* w, int Cout, int Cin, int ksize, bf16* fprop, bf16* dgrad,
                          cudaStream_t stream);

// ---- tensor-map builders shared by the tensor-core kernels (driver entry points resolved at run time)
CUtensorMap make_tmap_2d(const bf16* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
CUtensorMap make_tmap_im2col_box(const bf16* base, int N, int H, int W, int C, int lower_w, int lower_h,
                                 int upper_w, int upper_h, int stride);
CUtensorMap make_tmap_nhwc_box(const bf16* base, int N, int H, int W, int C, int box_w, int box_h);
CUtensorMap make_tmap_nhwc_rowclass(const bf16* base, int N, int H, int W, int C, int box_w, int box_rows);
CUtensorMap make_tmap_nhwc_box_c(const bf16* base, int N, int H, int W, int C, int box_c, int box_w, int box_h);

// ---- 64 -> 64, 3x3, stride 1, pad 1 convolutions with a shared-memory halo tile (conv_tc64.cu) ----------
// One TMA box per tile brings (R+2) x (W+2) zero-padded pixels; the nine filter taps are shifted views of that
// tile (UMMA descriptors offset by (r*(W+2)+s) rows), the 72 KB of weights stay resident in shared memory.
// L2->SM traffic per output tile drops from 216 KB (im2col per tap + weights) to ~23 KB.
int conv_tc64_grid(int N, int H, int W);             // CTAs conv_tc64_wgrad launches = slices of its `part` buffer
bool conv_tc64_supported(int H, int W);
int conv_tc64_fprop_grid(int N, int H, int W);       // CTAs conv_tc64_fprop launches = rows of its `stats` partial buffer
void conv_tc64_fprop(const bf16* in, int N, int H, int W, const bf16* wpk, const ConvEpilogue& epi, bf16* out,
                     cudaStream_t stream);
// part[conv_tc64_grid][9][64 ci][64 co] fp32; reduce with wgrad_reduce_tc().
void conv_tc64_wgrad(const bf16* x, const bf16* dy, int N, int H, int W, float* part, cudaStream_t stream);
void conv_tc64_set_variant(int v);                   // bring-up: descriptor base-offset policy

// Debug: what one im2col TMA box of 128 pixels x 64 channels delivers (de-swizzled), for tests.
void debug_im2col_tile(const bf16* in, int N, int H, int W, int Cin, int ksize, int stride, int pad,
                       int first_pixel, int tap_r, int tap_s, int cblock, bf16* out128x64,
                       cudaStream_t stream);

}  // namespace pcg

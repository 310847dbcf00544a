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

}  // namespace pcg

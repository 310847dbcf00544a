// Small fully connected layers (linear_small.cu): out[M][N] = epilogue(in[M][K] * W[N][K]^T), K, N <= 128, fp32.
#pragma once
#include "common.cuh"
#include "conv_generic.cuh"

namespace pcg {

bool linear_small_supported(const ConvGeom& g, long long M);
// w = [N][K] row major ("[out][in]"); for a data gradient pass dy as `in` and the transposed weight [K_layer][N_layer]
void linear_small(const float* in, long long M, int K, int N, const float* w, const GenEpilogue<float>& e, float* out,
                  cudaStream_t s);

}  // namespace pcg

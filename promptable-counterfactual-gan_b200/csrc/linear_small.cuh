// Small fully connected layers (linear_small.cu): out[M][N] = epilogue(in[M][K] * W[N][K]^T), K, N <= 128, fp32.
#pragma once
#include "common.cuh"
#include "conv_generic.cuh"

namespace pcg {

bool linear_small_supported(const ConvGeom& g, long long M);
// w = [N][K] row major ("[out][in]"); for a data gradient pass dy as `in` and the transposed weight [K_layer][N_layer]
// N may exceed 128 (column tiles); perm_c / perm_taps > 0: output column tap * perm_c + c reads weight row c * perm_taps + tap
void linear_small(const float* in, long long M, int K, int N, const float* w, const GenEpilogue<float>& e, float* out,
                  cudaStream_t s, int perm_c = 0, int perm_taps = 0);
bool full_window_dgrad_supported(const ConvGeom& g);

}  // namespace pcg

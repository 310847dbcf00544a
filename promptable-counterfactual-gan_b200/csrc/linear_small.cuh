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
// dw[N][K] = dy^T x and db[N] = column sums of dy in one call, two launches (K, N <= 128); scratch:
// linear_wgrad_small_scratch() floats (per-CTA partials; no initialisation needed), not shared by concurrent launches
bool linear_wgrad_small_supported(long long M, int K, int N);
long long linear_wgrad_small_scratch(long long M, int K, int N);
void linear_wgrad_small(const float* x, const float* dy, long long M, int K, int N, float* scratch, float* dw, float* db,
                        cudaStream_t s);

}  // namespace pcg

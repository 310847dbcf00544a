// Frozen MLP classifier: forward + cross-entropy + input gradient in one launch (see frozen_mlp.cu).
#pragma once
#include "common.cuh"

namespace pcg {

constexpr int FM_MAX_LAYERS = 6;
constexpr int FM_MAX_CLASSES = 8;

bool frozen_mlp_supported(int L, const int* dims);
int frozen_mlp_parts(int B);         // floats of loss_part
void frozen_mlp_ce_grad(int L, const int* dims, const float* const* W, const float* const* WT, const float* const* b,
                        float slope, const float* x,
                        const long long* target, int loss_kind, int B, float wgt, float* logits, float* loss_part, float* dx,
                        cudaStream_t s, float* const* act_out = nullptr, float* const* grad_out = nullptr);
// loss_kind 0: cross-entropy against target; 1: mean of the outputs.  act_out / grad_out: optional row-major copies of the
// hidden activations / pre-activation gradients (L - 1 pointers each) for a caller that computes weight gradients too

}  // namespace pcg

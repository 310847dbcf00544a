// Fused halves of a FiLM residual block (see film_layer.cu).
#pragma once
#include "common.cuh"

namespace pcg {

bool film_layer_supported(long long M, int H);
// one half block of a forward chain: the input of half k + 1 is the output of half k
struct FilmHalfFwd {
  const float* W; const float* bias; const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* nbt;
  float* mean; float* rstd; float* scale; float* shift;
  const float* fg; const float* fb; const float* res; int relu;
  float* u; float* n; float* out; float* part;
};
// n chained half blocks in n + 1 launches (the apply of one and the Linear of the next share a launch)
void film_chain_fwd(const float* x, long long M, int H, int n, const FilmHalfFwd* halves, float eps, float momentum,
                    cudaStream_t s);
void film_layer_fwd(const float* x, long long M, int H, const float* W, const float* bias, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var, long long* nbt,
                    float* mean, float* rstd, float* scale, float* shift, const float* fg, const float* fb, const float* res,
                    bool relu, float* u, float* n, float* out, float* part, cudaStream_t s);
void film_layer_bwd(const float* d_f, long long M, int H, const float* fg, const float* n, const float* u, const float* mean,
                    const float* rstd, const float* gamma, const float* W, const float* add_src, const float* act_ref,
                    bool accumulate, float* dfg, float* dfb, float* du, float* dx, float* dgamma, float* dbeta, float* part,
                    cudaStream_t s);   // part: pcg_stat_scratch_floats(H) floats, private to the call

}  // namespace pcg

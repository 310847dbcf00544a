// fp32-storage front end of the tensor-core convolution kernels (see conv_auto.cu).  Each function returns false
// when the layer is not eligible (or the mode is off); the caller then runs the CUDA-core kernel.
#pragma once
#include "common.cuh"
#include "conv_generic.cuh"

namespace pcg {

void conv_auto_set_tensor_cores(bool on);
bool conv_auto_tensor_cores();
// operand terms of the tensor-core mode: 3 = bf16x3 (fp32-equivalent products), 1 = plain bf16 operands
void conv_auto_set_terms(int t);
int conv_auto_terms();
// operand cache (see conv_auto.cu): reuse of the bf16 conversions of unchanged operands across calls
void conv_auto_set_operand_cache(bool on);
bool conv_auto_operand_cache();
void conv_auto_cache_clear();
void conv_auto_cache_invalidate(const void* p, size_t bytes);
bf16* conv_auto_cache_producer(const float* dst, long long n);
bool conv_fprop_auto(const float* in, const ConvGeom& g, const float* wf, const GenEpilogue<float>& e, float* out,
                     cudaStream_t s);
bool conv_dgrad_auto(const float* dout, const ConvGeom& g, const float* wd, const GenEpilogue<float>& e, float* din,
                     cudaStream_t s);
bool conv_wgrad_auto(const float* in, const float* dout, const ConvGeom& g, float* dw, cudaStream_t s);

}  // namespace pcg

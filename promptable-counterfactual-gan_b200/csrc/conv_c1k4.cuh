// One-channel 4x4 / stride-2 / pad-1 convolution kernels (conv_c1k4.cu): DCGAN's first discriminator and last
// generator layer (dconv_gan/mnist/mnist_dcgan.py:89,100).  fp32 NHWC, geometry [N,H,W,1] <-> [N,H/2,W/2,64].
#pragma once
#include "common.cuh"
#include "conv_generic.cuh"

namespace pcg {

bool c1k4_supported(const ConvGeom& g);
// out[N][H/2][W/2][64] = act(conv(x[N][H][W][1])), wf = [64][16]
void c1k4_fprop(const float* x, const ConvGeom& g, const float* wf, int act, float slope, float* out, cudaStream_t s,
                bf16* side = nullptr);      // side: optional bf16 copy of out (tensor-core operand cache)
// dx[N][H][W][1] = data gradient of that convolution, wd = [16][64]
void c1k4_dgrad(const float* dy, const ConvGeom& g, const float* wd, float* dx, cudaStream_t s);
// dw[64][1][4][4]; scratch holds c1k4_wgrad_scratch() floats
size_t c1k4_wgrad_scratch();
void c1k4_wgrad(const float* x, const float* dy, const ConvGeom& g, float* scratch, float* dw, cudaStream_t s);

// Full-window convolution to one output (Conv2d(C, 1, k, 1, 0) on a k x k map; mnist_dcgan.py:112): dot product,
// outer product, weighted column sum.
bool full1_supported(const ConvGeom& g);
void full1_fprop(const float* x, const ConvGeom& g, const float* wf, const float* bias, float* out, cudaStream_t s);
void full1_dgrad(const float* dz, const ConvGeom& g, const float* wd, float* dx, cudaStream_t s);
void full1_wgrad(const float* x, const float* dz, const ConvGeom& g, float* dw, cudaStream_t s);

}  // namespace pcg

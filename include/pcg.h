/* libpcg — C ABI of the B200-native GAN training step.
 *
 * Drop-in boundary for the hot path of flash4242/Promptable-Counterfactual-GAN.
 * The reference has no FFI of its own: its boundary is the Python surface
 *   conditional_counteRGAN/mnist/trainer.py:76   train_countergan(G, D, C, loader, cfg, device)
 *   conditional_counteRGAN/mnist/models/{generator,discriminator,classifier}.py      nn.Module ctor/forward/state_dict
 * and every operation below replaces the torch call cited next to it.  The Python mirror of that
 * surface (promptable-counterfactual-gan_b200/mnist/) binds these symbols with ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; pcg_last_error() gives the message
 *     (thread local).  The Python side raises RuntimeError.
 *   - all pointers are DEVICE pointers unless named host_*; the library borrows them, it never
 *     frees or reallocates caller memory.  Workspaces owned by a plan are freed by *_destroy.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless stated.
 *   - activations inside the library are NHWC ("pixels x channels"); images with one channel are
 *     identical in NCHW and NHWC, so x / mask / residuals cross the boundary unchanged.
 *   - parameters and gradients cross the boundary in torch layout (OIHW conv weights, [out][in]
 *     linear weights) as ONE flat fp32 arena per network, tensors in `module.parameters()` order,
 *     each tensor starting at a multiple of 4 floats (see pcg_mnist_*_layout).
 */
#ifndef PCG_H
#define PCG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCG_VERSION 1

const char* pcg_last_error(void);
int pcg_version(void);
/* Kernels launched by this library since load (bench.py "gpu_launches"). */
unsigned long long pcg_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Per-kernel entry points (used by the parity tests; the step plan calls the same code).
 * ------------------------------------------------------------------------------------------ */

/* precision codes */
#define PCG_F32 0
#define PCG_BF16 1
/* activation codes */
#define PCG_ACT_NONE 0
#define PCG_ACT_LRELU 1
#define PCG_ACT_RELU 2

/* Tensor-core implicit-GEMM convolution (tcgen05 + TMA im2col), replaces nn.Conv2d forward
 * (generator.py:11,14,49) and, with dgrad-packed weights, its input gradient.
 *   in  bf16 NHWC [N][H][W][Cin]; wpk bf16 [Cout][k*k][Cin]; out bf16 NHWC [N][Ho][Wo][Cout]
 *   bias fp32[Cout] or NULL; add_src bf16 like out or NULL; stats fp32 [pcg_conv_tc_grid][2*Cout]
 *   or NULL (per-CTA partial sum / sum of squares for train-mode BatchNorm). */
int pcg_conv_tc_grid(long long M, int Cout);
int pcg_conv_tc_fprop(const void* in, int N, int H, int W, int Cin, const void* wpk, int Cout,
                      int ksize, int stride, int pad, const float* bias, int act, float slope,
                      const void* add_src, void* out, float* stats, void* stream);
/* Weight gradient of the 64->64 3x3 s1 p1 convolution, replaces ConvolutionBackward0's wgrad.
 *   x, dy bf16 NHWC [N][H][W][64]; part fp32 scratch [pcg_conv_tc_wgrad_grid(M)][9*64*64];
 *   dw fp32 OIHW [64][64][3][3]. */
int pcg_conv_tc_wgrad_grid(long long M);
int pcg_conv_tc_wgrad64(const void* x, const void* dy, int N, int H, int W, float* part, float* dw,
                        void* stream);
/* fp32 OIHW -> bf16 [Cout][taps][Cin] (fprop) and rotated [Cin][taps][Cout] (dgrad; may be NULL). */
int pcg_pack_conv_weights_tc(const float* w, int Cout, int Cin, int ksize, void* fprop, void* dgrad,
                             void* stream);
/* Debug/test: one 128-pixel x 64-channel im2col TMA box, de-swizzled, bf16 [128][64]. */
int pcg_debug_im2col_tile(const void* in, int N, int H, int W, int Cin, int ksize, int stride, int pad,
                          int first_pixel, int tap_r, int tap_s, int cblock, void* out128x64,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCG_H */

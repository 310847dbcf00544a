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
/* Programmatic dependent launch: every kernel of the library lets its successor in the stream be launched early and
 * waits for its predecessor's completion before touching global memory (griddepcontrol), hiding launch latency and
 * prologues in the 100-1000-kernel steps.  Off by default (as a blanket policy it measured ~3 % slower on the MNIST
 * step and 17 % slower on the 1000-node KC graph, profiles/exp_pdl_r1.md); env PCG_PDL=1 enables; returns the previous
 * setting.
 * A step captured into a CUDA graph keeps the mode it was captured with. */
int pcg_set_pdl(int on);

/* Per-launcher device timing: begin() enables CUDA-event brackets around every kernel launcher
 * (not usable during graph capture); end() synchronises, disables, and writes a JSON object
 * {"launcher": {"ms": total, "launches": n}, ...} into out (host buffer of cap bytes). */
int pcg_profile_begin(void);
int pcg_profile_end(char* out, size_t cap);

/* Device-to-device copy on `stream` (test helper for reading plan-owned tensors). */
int pcg_memcpy_d2d(void* dst, const void* src, size_t nbytes, void* stream);

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
/* Halo-tile variants for the 64->64 / 3x3 / stride-1 / pad-1 case (conv_tc64.cu): one TMA box per tile,
 * weights resident in shared memory, the nine taps are shifted views of the tile.  Same operand layouts as
 * pcg_conv_tc_fprop / pcg_conv_tc_wgrad64.  When H % 4 == 0 the forward / data-gradient kernel is the row-class stacked
 * one (four output row classes per accumulator, tap matrices stacked along N: half the MMA instructions, 27 % less
 * shared-memory traffic; profiles/exp_tc64_stacked_r1.md); variant bit 256 of pcg_conv_tc64_set_variant selects the
 * one-class-per-tile kernel (same results).
 * stats rows = pcg_conv_tc64_fprop_grid(N, H, W); wgrad part slices = pcg_conv_tc64_grid(N, H, W). */
int pcg_conv_tc64_grid(int N, int H, int W);
int pcg_conv_tc64_fprop_grid(int N, int H, int W);
int pcg_conv_tc64_fprop(const void* in, int N, int H, int W, const void* wpk, const float* bias, int act,
                        float slope, const void* add_src, const void* act_ref, int ref_act, void* out,
                        float* stats, void* stream);
/* The same kernel with the reduction pass of a train-mode BatchNorm backward fused into its epilogue
 * (torch native_batch_norm_backward's two column sums; generator.py:12-15,22): v = conv(in) [+ add_src],
 * g = bn_gscale * v * act'(bn_scale*y + bn_shift) with y = bn_y at the same pixel,
 *   stats[cta][c] = sum g,  stats[cta][64 + c] = sum g * (y - bn_mean) * bn_rstd      (cta < pcg_conv_tc64_fprop_grid);
 * out = v as bf16.  add_src together with bn_y needs H % 4 == 0 (row-class kernel: the skip-gradient tile arrives by TMA
 * in the staging slot the result leaves from). */
int pcg_conv_tc64_dgrad_bnred(const void* in, int N, int H, int W, const void* wpk, const void* add_src, const void* bn_y,
                              const float* bn_mean, const float* bn_rstd, const float* bn_scale, const float* bn_shift,
                              int bn_act, float bn_slope, float bn_gscale, void* out, float* stats, void* stream);
int pcg_conv_tc64_wgrad(const void* x, const void* dy, int N, int H, int W, float* part, float* dw,
                        void* stream);
int pcg_conv_tc64_set_variant(int v);

/* Skinny layers of the MNIST step on the warp-level tensor path (conv_small.cu); 3x3, pad 1, NHWC bf16 activations.
 * Replace nn.Conv2d forward / ConvolutionBackward0 of conv_in (3->64), conv_out (64->1) (generator.py:39,50),
 * Discriminator main.0 (2->64, stride 2; discriminator.py:15) and CNNClassifier conv.0 (1->32; classifier.py:8).
 *   pcg_conv_to1     out[N*H*W] f32 = bias + sum_{tap,c} in[..][c] * w9[tap][c]                (Cin = 32 | 64)
 *   pcg_conv_few     out bf16 [M][Cout] = epi(sum in[..][c] * wnk[co][tap*Cs + c]), Cs = 1..3, Cout = 32 | 64, stride 1|2;
 *                    epi: + bias, act, * act'(act_ref)
 *   pcg_dgrad_s2_to1 dx[N][H][W] f32 = one input channel of the data gradient of a 64-output-channel stride-2 conv;
 *                    wrot = that channel's row of the rotated packing, bf16 [9][64]
 *   pcg_wgrad_few    dw [64][Cs][3][3] f32 (+ db[64] if not NULL) from in [N][H][W][Cs] and dy [M][64]
 *   pcg_wgrad_to1    dw [1][64][3][3], db[1] from x [N][H][W][64] and g [N*H*W]
 *   part: pcg_wgrad_small_parts() * 2048 floats of scratch. */
int pcg_conv_to1(const void* in, int N, int H, int W, int Cin, const void* w9, const float* bias, float* out, void* stream);
int pcg_conv_few(const void* in, int in_is_f32, int N, int H, int W, int Cs, const void* wnk, int Cout, int stride,
                 const float* bias, int act, float slope, const void* act_ref, int ref_act, float ref_slope, void* out,
                 void* stream);
int pcg_dgrad_s2_to1(const void* dy, int N, int H, int W, const void* wrot, float* dx, void* stream);
int pcg_wgrad_small_parts(void);
int pcg_wgrad_few(const void* in, const void* dy, int N, int H, int W, int Cs, int stride, float* part, float* dw, float* db,
                  void* stream);
int pcg_wgrad_to1(const void* x, const void* g, int N, int H, int W, float* part, float* dw, float* db, void* stream);

/* fp32 OIHW -> bf16 [Cout][taps][Cin] (fprop) and rotated [Cin][taps][Cout] (dgrad; may be NULL). */
int pcg_pack_conv_weights_tc(const float* w, int Cout, int Cin, int ksize, void* fprop, void* dgrad,
                             void* stream);
/* Debug/test: one 128-pixel x 64-channel im2col TMA box, de-swizzled, bf16 [128][64]. */
int pcg_debug_im2col_tile(const void* in, int N, int H, int W, int Cin, int ksize, int stride, int pad,
                          int first_pixel, int tap_r, int tap_s, int cblock, void* out128x64,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * MNIST conv CounteRGAN step plan
 *   replaces the loop body of conditional_counteRGAN/mnist/trainer.py:89-132 (train_countergan)
 *   and the forwards of models/generator.py:71-86, models/discriminator.py:33-38,
 *   models/classifier.py:25-28.
 * ------------------------------------------------------------------------------------------ */
typedef struct pcg_mnist_plan pcg_mnist_plan;

typedef struct pcg_mnist_config {
  int batch;              /* samples per step on this GPU                                        */
  int base_ch;            /* ResidualGenerator(base_ch=...)   generator.py:32  (64)              */
  int n_resblocks;        /* ResidualGenerator(n_resblocks=...)                (6)               */
  int precision;          /* PCG_F32: CUDA-core fp32 everywhere; PCG_BF16: bf16 activations,
                             tcgen05 tensor-core convolutions with fp32 accumulation             */
  float g_lr, d_lr;       /* config.py:10-11                                                     */
  float beta1, beta2, adam_eps;   /* torch.optim.Adam defaults .9 / .999 / 1e-8 (trainer.py:77)  */
  float lambda_adv, lambda_cls, lambda_reg, lambda_mask;   /* config.py:12-15                    */
  float residual_scaling; /* generator.py:33 (0.1)                                               */
  float grad_scale;       /* gradients are multiplied by this before Adam (1/world_size)         */
  int use_tensor_cores;   /* PCG_BF16 only: 0 keeps bf16 storage but runs every convolution on the CUDA-core
                             kernels (A/B check of the tcgen05 path); 1 = tcgen05 where eligible        */
  int pollute_d_grads;    /* !=0: the G step also accumulates its weight gradients into d_grads,
                             as the reference's g_loss.backward() does (diagnostic only)         */
} pcg_mnist_config;

/* Caller-owned device memory the plan borrows.  All fp32 unless stated.  Arena layouts are given
 * by pcg_mnist_layout().  Gradients are WRITTEN (not accumulated) by every step. */
typedef struct pcg_mnist_buffers {
  float* g_params; float* g_grads; float* g_adam_m; float* g_adam_v; int* g_step;
  float* g_bn_running;        /* [2*n_resblocks][2][base_ch]: (running_mean, running_var) per BN,
                                 order bn1, bn2 of block 0, bn1, bn2 of block 1, ...             */
  long long* g_bn_nbt;        /* [2*n_resblocks] num_batches_tracked (int64)                      */
  float* d_params; float* d_grads; float* d_adam_m; float* d_adam_v; int* d_step;
  const float* c_params;      /* frozen classifier                                               */
} pcg_mnist_buffers;

typedef struct pcg_mnist_inputs {   /* device pointers                                            */
  const float* x;             /* [B][1][28][28] in [-1,1]                                         */
  const long long* y;         /* [B] int64 labels              trainer.py:90                      */
  const long long* target;    /* [B] int64 target classes      trainer.py:94 (injected draw)      */
  const float* mask;          /* [B][1][28][28] {0,1}          trainer.py:95 (injected draw)      */
} pcg_mnist_inputs;

/* indices into the scalar block written by the step (device float[PCG_MNIST_NSCALARS]) */
enum {
  PCG_S_D_LOSS = 0, PCG_S_G_LOSS = 1, PCG_S_G_ADV = 2, PCG_S_G_CLS = 3, PCG_S_REG_L1 = 4,
  PCG_S_MASK_PEN = 5, PCG_S_D_REAL_P = 6, PCG_S_D_FAKE_P = 7, PCG_S_D_LOSS_REAL = 8,
  PCG_S_D_LOSS_FAKE = 9, PCG_MNIST_NSCALARS = 16
};

/* net: 0 = generator, 1 = discriminator, 2 = classifier.  Returns the number of parameter
 * tensors; if idx >= 0 also the arena offset / element count (floats) of tensor idx, tensors in
 * module.parameters() order.  *total = arena size in floats. */
int pcg_mnist_layout(int net, int base_ch, int n_resblocks, int idx, long long* offset,
                     long long* numel, long long* total);

int pcg_mnist_plan_create(const pcg_mnist_config* cfg, const pcg_mnist_buffers* buf,
                          pcg_mnist_plan** out);
int pcg_mnist_plan_destroy(pcg_mnist_plan* plan);
/* Re-derive the packed (kernel-layout) weights from the fp32 arenas; call after the caller
 * modified parameters behind the plan's back (load_state_dict, optimizer of its own...). */
int pcg_mnist_refresh_weights(pcg_mnist_plan* plan, void* stream);

/* One full iteration = the four phases below in order.  scalars: device float[PCG_MNIST_NSCALARS]. */
int pcg_mnist_step(pcg_mnist_plan* plan, const pcg_mnist_inputs* in, float* scalars, void* stream);
/* phase 1: G forward, x_cf, D forward/backward on (x,y) and (x_cf.detach(), target) -> d_grads; beside it (side
 *          stream, joined before the phase returns) the frozen classifier's forward / input gradient on x_cf, which
 *          needs nothing else: scalars[PCG_S_G_CLS] and the plan-owned d CE / d x_cf are consumed by phase 3 */
int pcg_mnist_step_d_grads(pcg_mnist_plan* plan, const pcg_mnist_inputs* in, float* scalars, void* stream);
/* phase 2: Adam on D (gradient all-reduce, if any, happens between phase 1 and 2) */
int pcg_mnist_step_d_update(pcg_mnist_plan* plan, void* stream);
/* phase 3: D forward with the updated D, losses, G backward -> g_grads (same `in` and `scalars` as phase 1) */
int pcg_mnist_step_g_grads(pcg_mnist_plan* plan, const pcg_mnist_inputs* in, float* scalars, void* stream);
/* Data-parallel refinements of phases 1 and 3 ("each [reduction] can be bucketed and overlapped with its own
 * backward pass (D's also with the classifier fwd/dgrad ...)"):
 *   pcg_mnist_set_defer_c_bwd(plan, 1): phase 1 stops the frozen classifier's branch after its forward + loss; the caller
 *     runs pcg_mnist_step_c_bwd (the input-gradient chain) on another stream beside the all-reduce of d_grads and joins
 *     it before phase 3;
 *   pcg_mnist_step_g_grads_part(part = 1, split): phase 3 down to and including residual block `split` (weight-gradient
 *     stream joined): the gradients of [resblocks.split .. conv_out], the tail of the flat arena, are final and can be
 *     all-reduced while part = 2 (blocks split-1 .. 0, conv_in, embedding) runs.  1 <= split < n_resblocks. */
int pcg_mnist_set_defer_c_bwd(pcg_mnist_plan* plan, int on);
int pcg_mnist_step_c_bwd(pcg_mnist_plan* plan, void* stream);
int pcg_mnist_step_g_grads_part(pcg_mnist_plan* plan, const pcg_mnist_inputs* in, float* scalars, int part, int split_block,
                                void* stream);
/* phase 4: Adam on G */
int pcg_mnist_step_g_update(pcg_mnist_plan* plan, void* stream);

/* Module forwards (nn.Module.forward mirrors).  training != 0: BatchNorm uses batch statistics and
 * updates the running buffers (generator.py train mode); == 0: running statistics. */
int pcg_mnist_g_forward(pcg_mnist_plan* plan, const float* x, const long long* target,
                        const float* mask, int training, float* raw, float* masked, void* stream);
int pcg_mnist_d_forward(pcg_mnist_plan* plan, const float* x, const long long* cond, float* logits,
                        void* stream);
int pcg_mnist_c_forward(pcg_mnist_plan* plan, const float* x, float* logits, void* stream);

/* Test hook: device pointer / element count / dtype (PCG_F32 or PCG_BF16) of an internal NHWC
 * tensor by name ("h0", "y1.3", "z1.0", "y2.5", "h.6", "hm", "x_cf", "raw", "dxd", "dxc", ...). */
int pcg_mnist_debug_tensor(pcg_mnist_plan* plan, const char* name, void** ptr, long long* numel,
                           int* dtype);

/* ------------------------------------------------------------------------------------------
 * Primitive operators (fp32, NHWC / row-major) composed by the Python step plans of the other
 * hot-path configurations: conditional_gan/moons/make_moons_cgan.py:83-135,
 * simple_gan/moons/make_moons_gan.py:49-93, conditional_counteRGAN/{moons,house_sales_kc_usa}/trainer.py,
 * dconv_gan/mnist/mnist_dcgan.py:143-175.  nn.Linear is the 1x1 case of the convolution entry points
 * (weights [out][in] == [Cout][1][Cin]); nn.ConvTranspose2d forward is pcg_conv_dgrad of the mirrored
 * convolution, its input gradient pcg_conv_fprop.
 * ------------------------------------------------------------------------------------------ */
int pcg_conv_fprop(const float* in, int N, int H, int W, int Cin, const float* wf /*[Cout][k*k][Cin]*/, int Cout,
                   int k, int stride, int pad, const float* bias, int act, float slope, const float* add_src,
                   float* out, void* stream);
int pcg_conv_dgrad(const float* dout, int N, int H, int W, int Cin, const float* wd /*[Cin][k*k][Cout]*/, int Cout,
                   int k, int stride, int pad, const float* add_src, const float* act_ref, int ref_act,
                   float ref_slope, float* din, void* stream);
long long pcg_conv_wgrad_scratch(int N, int H, int W, int Cin, int Cout, int k, int stride, int pad);
int pcg_conv_wgrad(const float* in, const float* dout, int N, int H, int W, int Cin, int Cout, int k, int stride,
                   int pad, float* scratch, float* dw /*torch OIHW*/, void* stream);
/* perm_hw > 0: Cin is a flattened (c, hw) index in torch's NCHW order, re-ordered to NHWC (hw, c); perm_hw == -1: wd is
 * written with its taps reversed (the forward weight of the dilated-gradient convolution, see pcg_dilate). */
int pcg_pack_conv_weights(const float* w /*torch OIHW*/, int Cout, int Cin, int k, int perm_hw, float* wf, float* wd,
                          void* stream);
/* Tensor-core mode of the three convolution entry points above (default off = exact fp32 on the CUDA cores): layers
 * with 64-multiple channel counts (fprop: any k/stride/pad; dgrad: 4x4 stride 2 pad 1 on even sizes, i.e. the
 * ConvTranspose2d forward and Conv2d input gradient of mnist_dcgan.py:72-116; wgrad: 3x3 / 4x4, stride 1|2) round their
 * operands to bf16 into library-owned scratch and run the tcgen05 implicit-GEMM kernels with fp32 accumulation.
 * Scratch is sized by the first (eager) call, so run one un-captured pass before CUDA-graph capture. */
int pcg_set_conv_tensor_cores(int on);
int pcg_get_conv_tensor_cores(void);
/* Operand cache of the tensor-core mode.  Every tensor-core call converts its fp32 operands to bf16; with the cache on
 * (returns the previous setting) a conversion is kept in its own buffer, keyed by (pointer, size, layout), and reused by
 * later calls while the source tensor is unchanged - the CALLER says when it changed: pcg_operand_cache_invalidate(ptr,
 * bytes) after any write into [ptr, ptr + bytes), pcg_operand_cache_clear() to drop everything (start of a step, after
 * inputs were written outside the library).  Producers that are told nothing stay correct with the cache off (default).
 * New buffers are cudaMalloc'ed: populate by one eager pass before stream capture. */
int pcg_set_operand_cache(int on);
void pcg_operand_cache_clear(void);
void pcg_operand_cache_invalidate(const void* p, long long bytes);
/* Operand precision of that mode: 3 (default) = bf16x3, every fp32 product emulated as hi*hi + lo*hi + hi*lo (fp32-level
 * results, 3x the tensor work); 1 = plain bf16 operands with fp32 accumulation (what torch autocast runs).  Returns the
 * previous setting. */
int pcg_set_conv_tensor_core_terms(int terms);
long long pcg_stat_scratch_floats(int C);
int pcg_colsum(const float* a, long long M, int C, float* scratch, float* out, void* stream);
/* nn.BatchNorm{1,2}d in train mode over M rows x C channels (+ fused activation), and its backward
 * (dz = gradient wrt the activation output; dbias_prev = column sums of dy, the bias gradient of the layer in front) */
int pcg_bn_train_fwd(const float* y, long long M, int C, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, long long* nbt, float* mean,
                     float* rstd, float* scale, float* shift, int act, float slope, float* z, float* scratch,
                     void* stream);
int pcg_bn_train_bwd(const float* dz, const float* y, long long M, int C, const float* gamma, const float* mean,
                     const float* rstd, const float* scale, const float* shift, float gscale, int act, float slope,
                     float* dy, float* dgamma, float* dbeta, float* dbias_prev, float* c12, float* scratch,
                     float* scratch2, void* stream);
int pcg_bn_eval(const float* x, long long rows, int C, const float* gamma, const float* beta, const float* rm,
                const float* rv, float eps, float* y, float* scale_out, void* stream);
int pcg_scale_cols(const float* dy, long long rows, int C, const float* scale, float* dx, void* stream);
/* elementwise: op codes 1 relu, 2 lrelu(a), 3 sigmoid, 4 tanh, 5 scale(a), 6 copy; binary 0 add (alpha*a+beta*b), 1 mul */
int pcg_unary(const float* x, long long n, int op, float a, float* y, void* stream);
int pcg_unary_bwd(const float* dy, const float* y, long long n, int op, float a, float* dx, void* stream);
int pcg_binary(const float* a, const float* b, long long n, int op, float alpha, float beta, float* out, void* stream);
/* FiLM modulation (house_sales_kc_usa/models/generator.py:13-16,28-35) in one launch each way:
 *   fwd: out = [relu](gamma * n + beta) [+ res]          (res may be NULL)
 *   bwd: dn = df * gamma ; dgamma (+)= df * n ; dbeta (+)= df   (accumulate != 0 adds: the same FiLM is used twice per block) */
int pcg_film_fwd(const float* gamma, const float* n_, const float* beta, const float* res, long long n, int relu,
                 float* out, void* stream);
int pcg_film_bwd(const float* df, const float* gamma, const float* n_, long long n, int accumulate, float* dn,
                 float* dgamma, float* dbeta, void* stream);
/* dst[i] = src[i]^T (src[i] is [rows[i]][cols[i]]) for n <= 64 small matrices in one launch; the pointer / size arrays
 * are HOST arrays (copied into the kernel arguments). */
int pcg_transpose_multi(int n, const float* const* src, float* const* dst, const int* rows, const int* cols, void* stream);
int pcg_copy_cols(const float* src, int src_ld, int c0_src, float* dst, int dst_ld, int c0_dst, long long rows,
                  int ncols, float alpha, int accumulate, void* stream);
int pcg_onehot(const long long* lab, long long rows, int nc, float* dst, int dst_ld, int c0, void* stream);
int pcg_reduce_scalar(const float* x, long long n, int absval, float scale, float* out, float gscale, float* dx,
                      void* stream);
int pcg_rownorm_mean(const float* x, long long rows, int cols, int p, float* out, float gscale, float* dx,
                     void* stream);
/* kind 0: -mean log(sigmoid) terms on logits; 1: BCELoss (log clamped at -100); 2: Wasserstein mean */
int pcg_gan_loss(const float* z, int n, int kind, float t, float wgt, float* out_loss, float* out_aux, float* dz,
                 void* stream);
int pcg_combine_scalars(int n, const float* host_coeffs, const float* const* host_ptrs, float* out, void* stream);
/* One whole iteration of the two-layer MLP GANs on 2-D points in ONE launch (mlp_gan.cu): replaces the loop bodies of
 * conditional_gan/moons/make_moons_cgan.py:90-129 and simple_gan/moons/make_moons_gan.py:61-88 - generator forward on
 * (z1, oh1), discriminator on real + fake, loss_D = -mean(log D(real) + log(1 - D(fake))), Adam on D, generator forward on
 * (z2, oh2), loss_G = -mean(log D(fake2)) through the UPDATED discriminator, Adam on G (lr, betas 0.9/0.999, eps 1e-8).
 * A cluster of up to 8 CTAs, all weights in shared memory, gradients summed over the cluster in rank order.
 *   hidden = 128, label_dim <= 2 (0: unconditional, the *_oh pointers may be NULL), z_dim + label_dim <= 36, B <= 1024.
 *   *_param / *_grad / *_m / *_v: flat fp32 buffers [0.weight | 0.bias | 2.weight | 2.bias], every slice padded to a
 *   multiple of 4 floats; *_step: int32 Adam step counters (incremented); scal[8]: 0 loss_D, 1 loss_G, 2 real term,
 *   3 fake term, 4 mean D(real), 5 mean D(fake). */
int pcg_mlp_gan_step(int B, int z_dim, int label_dim, int hidden, const float* real, const float* real_oh,
                     const float* z1, const float* oh1, const float* z2, const float* oh2, float* g_param, float* g_grad,
                     float* g_m, float* g_v, int* g_step, float* d_param, float* d_grad, float* d_m, float* d_v,
                     int* d_step, float lr, float* scal, void* stream);
int pcg_spectral_norm_fwd(const float* W, int N, int K, float* u, float* v, float eps, int do_iter, float* Wn,
                          float* sigma, void* stream);
/* Same with the extra outputs a training pass needs in the same launch: WnT = Wn^T [K][N] (operand of the data
 * gradient), us / vs = copies of u / v after the power iteration (any of the three may be NULL). */
int pcg_spectral_norm_fwd2(const float* W, int N, int K, float* u, float* v, float eps, int do_iter, float* Wn,
                           float* WnT, float* us, float* vs, float* sigma, void* stream);
int pcg_spectral_norm_bwd(const float* dWn, const float* Wn, int N, int K, const float* u, const float* v,
                          const float* sigma, float* dW, void* stream);
int pcg_gumbel_softmax_fwd(const float* logits, const float* g, long long rows, int n, float tau, float* y,
                           void* stream);
/* y[r] = one_hot(argmax_j x[r][j]) (first maximum): the forward value of F.gumbel_softmax(hard=True) from its soft
 * sample, house_sales_kc_usa/eval_utils.py:75.  x and y may alias. */
int pcg_onehot_argmax(const float* x, long long rows, int n, float* y, void* stream);
int pcg_softmax_bwd(const float* dy, const float* y, long long rows, int n, float tau, float* dl, void* stream);
int pcg_ce_loss(const float* logits, const long long* target, int B, int NC, float wgt, float* loss, float* dlogits,
                void* stream);
int pcg_adam_flat(float* p, const float* g, float* m, float* v, long long n, int* step, float lr, float beta1,
                  float beta2, float eps, float grad_scale, void* stream);

/* nn.CrossEntropyLoss(weight=class_weights) with mean reduction (house_sales_kc_usa/trainer.py:58): loss =
 * sum_n w[t_n] nll_n / sum_n w[t_n]; dlogits (may be NULL) its gradient; correct (may be NULL) = number of rows whose
 * arg-max equals the target (trainer.py:93-95).  class_weights == NULL: unweighted mean. */
int pcg_ce_loss_weighted(const float* logits, const long long* target, const float* class_weights, int B, int NC,
                         float* loss, float* dlogits, float* correct, void* stream);
/* torch.optim.AdamW on a flat buffer (trainer.py:60): p *= 1 - lr*weight_decay, then the Adam update; lr is read from
 * DEVICE memory (lr_dev[0]) so a scheduler (ReduceLROnPlateau, trainer.py:61) can change it under a captured graph. */
int pcg_adamw_flat(float* p, const float* g, float* m, float* v, long long n, int* step, const float* lr_dev, float beta1,
                   float beta2, float eps, float weight_decay, void* stream);

/* On-device input pipeline, replaces DataLoader + transforms.ToTensor() + transforms.Normalize((mean,), (std,)) of
 * conditional_counteRGAN/mnist/data_utils.py:9-12,26 for a dataset kept resident in HBM as uint8 [N][HW] (HW % 16 == 0):
 *   x[b][.] = (float(images[index[b]][.]) / 255 - mean) / std  (bit-identical to torchvision),  y[b] = labels[index[b]];
 * index == NULL gathers rows 0..B-1, y / labels may be NULL. */
int pcg_u8_batch(const unsigned char* images, const long long* labels, const long long* index, int B, int HW, float mean,
                 float stdv, float* x, long long* y, void* stream);

/* Random patch mask and target draw of one iteration, replaces build_mask (conditional_counteRGAN/mnist/trainer.py:45-72:
 * a Python loop of B randperm calls + F.interpolate + repeat) and `target_y = torch.randint(0, num_classes, (bs,))`
 * (trainer.py:94) by ONE launch:
 *   mask[B][C][H][W] (fp32, 0/1): per sample a uniformly random subset of `num_modifiable_patches` of the
 *   (H / patch) x (W / patch) patches (at most 64) is set to 1 - independent fair coins per patch when
 *   num_modifiable_patches < 0 or >= the patch count (trainer.py:59-61) - then nearest-upsampled to H x W
 *   (source index = floor(dst * patches / size), as F.interpolate(mode="nearest")) and repeated over the C channels;
 *   target[b] uniform in [0, num_classes) (target == NULL: not drawn).
 * Randomness: Philox4x32-10, counter = (sample, draw, stream offset).  rng_state points at three device uint64
 * {stream offset, 0, key}: the kernel uses key' = seed + key and the offset, and advances the offset, so the launch may
 * be replayed inside a CUDA graph and re-seeded without re-capturing; NULL = key seed, offset 0 (deterministic, for
 * tests).  Same distribution as the reference, not the same stream. */
int pcg_build_mask(int B, int C, int H, int W, int patch, int num_modifiable_patches, int num_classes,
                   unsigned long long seed, unsigned long long* rng_state, float* mask, long long* target, void* stream);
/* Counterfactual evaluation metrics of conditional_counteRGAN/mnist/eval_utils.py:46-75 (evaluate_counterfactuals) and
 * house_sales_kc_usa/eval_utils.py:232-262:  pcg_cf_apply: x_cf = clamp(x + residual, lo, hi) over n elements and the
 * partial sums of |x_cf - x| into scratch (pcg_cf_scratch_floats() floats);  pcg_cf_metrics (after the classifier ran on
 * x_cf): out3 = {class-flip rate = mean(argmax logits == y_target), prediction gain = mean(softmax[y_target] -
 * softmax[y_true]), actionability = sum |x_cf - x| / n_elems}. */
int pcg_cf_scratch_floats(void);
/* nn.InstanceNorm2d(C, affine=True) of the conditional WGAN-GP critic (conditional_gan/mnist/mnist_wgan_conditional.py:
 * 87-95) on NHWC activations [N][P = H*W][C], statistics per (sample, channel), biased variance:
 *   pcg_instnorm_fwd      y = act(gamma * xhat + beta), saves mean / rstd [N][C];
 *   pcg_instnorm_bwd      p = gy * act'(act_ref) (act_ref = the layer's output, NULL: p = gy);
 *                         dx = gamma * rstd * (p - mean(p) - xhat * mean(p * xhat)) (+ add_src);
 *                         dgamma_part / dbeta_part [N][C] per-sample sums (NULL: skipped; reduce with pcg_colsum);
 *   pcg_instnorm_bwd_bwd  the backward of pcg_instnorm_bwd, which the gradient penalty (:146-150, autograd.grad with
 *                         create_graph=True) back-propagates through: given q = the cotangent of dx, writes gy_bar (cotangent
 *                         of gy), x_bar (cotangent of x through xhat and rstd) and per-sample parts of gamma's cotangent.
 *   pcg_flatten_nchw      dst[b][c0 + c * R + r] = src[b][r][c] (dst rows of ld floats): NHWC features into torch's
 *                         nn.Flatten (NCHW) column order inside a wider matrix (:106 torch.cat); inverse != 0: the same map
 *                         read backwards (dst[b][r][c] = src[b][c0 + c * R + r]). */
int pcg_instnorm_fwd(const float* x, int N, int P, int C, const float* gamma, const float* beta, float eps, int act, float slope,
                     float* y, float* mean, float* rstd, void* stream);
int pcg_instnorm_bwd(const float* gy, const float* act_ref, int act, float slope, const float* x, const float* mean,
                     const float* rstd, const float* gamma, int N, int P, int C, const float* add_src, float* dx,
                     float* dgamma_part, float* dbeta_part, void* stream);
int pcg_instnorm_bwd_bwd(const float* q, const float* gy, const float* act_ref, int act, float slope, const float* x,
                         const float* mean, const float* rstd, const float* gamma, int N, int P, int C, float* gy_bar,
                         float* x_bar, float* dgamma_part, void* stream);
int pcg_flatten_nchw(const float* src, int B, int R, int C, float* dst, int ld, int c0, int inverse, void* stream);
/* The frozen classifier of the tabular CounteRGAN generator step (house_sales_kc_usa/trainer.py:300-303, moons/trainer.py:
 * 85-87: F.cross_entropy(clf_model(x_cf), target) with clf_model in eval mode and frozen) in ONE launch: an MLP of L Linear
 * layers (torch [out][in] weights W[j], their transposes WT[j] [in][out], biases b[j]: host arrays of device pointers;
 * dims[0..L] = the widths; LeakyReLU(slope) between the layers - slope 0 is ReLU; eval-mode BatchNorms folded into the
 * following Linear by the caller), its
 * cross-entropy against `target` (loss_kind 0) or the mean of its outputs (loss_kind 1: a critic score inside the generator
 * step, trainer.py:298; target may be NULL), and the backward pass down to dx = wgt * d loss / d x.  logits (NULL: not stored)
 * [B][dims[L]]; loss_part: pcg_frozen_mlp_parts() floats (-1: shape not supported), sum(loss_part) / B = the mean loss.
 * Hidden widths in {16, 32, 64, 128, 256}, dims[0] <= 64, dims[L] <= 8. */
int pcg_frozen_mlp_parts(int L, const int* dims, int B);
int pcg_frozen_mlp_ce_grad(int L, const int* dims, const float* const* W, const float* const* WT, const float* const* b,
                           float slope, const float* x, const long long* target, int loss_kind, int B, float wgt,
                           float* logits, float* loss_part, float* dx, void* stream);
/* pcg_frozen_mlp_ce_grad that also stores what the weight gradients of the SAME network need (a critic's own update,
 * trainer.py:290-295): act_out[j] [B][dims[j + 1]] = activation of hidden layer j, grad_out[j] = gradient of the loss with
 * respect to that layer's pre-activation (j = 0 .. L-2; host arrays of device pointers).  The gradient of the output layer's
 * pre-activation is the caller's (wgt / B for the mean score).  dW_j = grad_out[j]^T act_out[j - 1] (pcg_linear_wgrad_small). */
int pcg_mlp_fwd_bwd(int L, const int* dims, const float* const* W, const float* const* WT, const float* const* b, float slope,
                    const float* x, const long long* target, int loss_kind, int B, float wgt, float* logits, float* loss_part,
                    float* dx, float* const* act_out, float* const* grad_out, void* stream);
/* Weight and bias gradient of a small nn.Linear (K inputs, N outputs, both <= 128) in ONE call of two launches (product over
 * up to 128 row slices, then a wide deterministic sum): dw[N][K] = dy^T x (torch layout), db[N] = column sums of dy (NULL:
 * skipped) - instead of pcg_conv_wgrad + pcg_colsum (four launches).  scratch: pcg_linear_wgrad_small_scratch(M, K, N) floats
 * (-1: shape not supported), 16-byte aligned, uninitialised, not shared by launches that may run concurrently. */
long long pcg_linear_wgrad_small_scratch(long long M, int K, int N);
int pcg_linear_wgrad_small(const float* x, const float* dy, long long M, int K, int N, float* scratch, float* dw, float* db,
                           void* stream);
/* One call (two launches: the batch statistics are one grid-wide dependency) per half of a FiLM residual block of the
 * tabular generator (house_sales_kc_usa/models/generator.py:19-35: nn.Linear(H, H) -> nn.BatchNorm1d(H) in training mode ->
 * FiLM -> ReLU or residual add), H in {32, 64} (pcg_film_layer_supported), instead of five / six primitive operators.
 * scratch: pcg_stat_scratch_floats(H) floats private to the call (per-CTA partial sums).
 *   fwd:  u = x W^T + bias;  n = BN(u) (saves mean / rstd / scale / shift, updates the running buffers like
 *         pcg_bn_train_fwd);  f = fg * n + fb ([M][H] FiLM tensors);  out = relu ? max(f, 0) : res + f
 *   bwd:  given d_f = dL/df:  dfg (+)= d_f * n;  dfb (+)= d_f (accumulate != 0: +=);  du = BatchNorm backward of d_f * fg
 *         (written: the weight gradient needs it), dgamma / dbeta;  dx = du W (+ add_src), zeroed where act_ref <= 0.
 * W is the torch [out][in] weight in both calls. */
int pcg_film_layer_supported(long long M, int H);
int pcg_film_layer_fwd(const float* x, long long M, int H, const float* W, const float* bias, const float* gamma,
                       const float* beta, float eps, float momentum, float* running_mean, float* running_var, long long* nbt,
                       float* mean, float* rstd, float* scale, float* shift, const float* fg, const float* fb,
                       const float* res, int relu, float* u, float* n, float* out, float* scratch, void* stream);
/* A chain of n half blocks (the output of one is the input of the next: the generator's five residual blocks are ten) in
 * n + 1 launches instead of 2 n: the BatchNorm + FiLM + activation of half k and the Linear of half k + 1 share a launch.
 * Same results as n calls of pcg_film_layer_fwd.  res == NULL selects out = relu(f), otherwise out = res + f. */
typedef struct pcg_film_half {
  const float* W; const float* bias; const float* gamma; const float* beta;          /* Linear and BatchNorm parameters */
  float* running_mean; float* running_var; long long* nbt;                            /* BatchNorm buffers (may be NULL) */
  float* mean; float* rstd; float* scale; float* shift;                               /* saved statistics [H] */
  const float* fg; const float* fb; const float* res;                                 /* FiLM tensors [M][H], residual */
  float* u; float* n; float* out; float* scratch;                                     /* [M][H] each; scratch as above */
} pcg_film_half;
int pcg_film_chain_fwd(const float* x, long long M, int H, int n, const pcg_film_half* halves, float eps, float momentum,
                       void* stream);
int pcg_film_layer_bwd(const float* d_f, long long M, int H, const float* fg, const float* n, const float* u,
                       const float* mean, const float* rstd, const float* gamma, const float* W, const float* add_src,
                       const float* act_ref, int accumulate, float* dfg, float* dfb, float* du, float* dx, float* dgamma,
                       float* dbeta, float* scratch, void* stream);
/* y[r][c] = x[r][c] + bias[c], tanh applied when tanh_out != 0: the bias (and final Tanh, mnist_wgan_conditional.py:62-72)
 * of a ConvTranspose2d whose product was computed as a data gradient.  In place (y == x) allowed. */
int pcg_bias_act(const float* x, long long rows, int C, const float* bias, int tanh_out, float* y, void* stream);
/* dst [N][Hp][Wp][C] = zero-dilated, padded copy of src [N][Ho][Wo][C]: dst[n][off + stride*y][off + stride*x] = src[n][y][x],
 * zero elsewhere.  With off = k - 1 - pad and Hp = H + k - 1 the data gradient of Conv2d(k, stride, pad) on H x W inputs
 * (= the forward of the mirrored ConvTranspose2d) equals pcg_conv_fprop(dst, N, Hp, Wp, Cout -> Cin, k, stride 1, pad 0)
 * with the wd packed by pcg_pack_conv_weights(perm_hw = -1) as its forward weight: any geometry on the tensor-core
 * forward kernel, at stride^2 times the arithmetic.  C % 4 == 0. */
int pcg_dilate(const float* src, int N, int Ho, int Wo, int C, int stride, int off, int Hp, int Wp, float* dst, void* stream);
/* Data gradient of a stride-2 convolution (k = 3 or 4, any padding; = forward of the mirrored ConvTranspose2d) without the
 * zeros of pcg_dilate: for each parity class (cy, cx) of the input position a stride-1 2x2 forward convolution of the
 * gradient,  pcg_conv_fprop(dy, N, Ho, Wo, Cout -> Cin, k 2, stride 1, pad 1, weight wc + cls * Cin * 4 * Cout)
 * -> class map [N][Ho + 1][Wo + 1][Cin],  then pcg_parity_interleave of the four maps (stored back to back) into dx
 * [N][H][W][Cin]:  dx[n][y][x] = class (y + pad) & 1, (x + pad) & 1 at ((y + pad) >> 1, (x + pad) >> 1).
 * pcg_pack_dgrad_classes builds wc [4][Cin][4][Cout] from the torch OIHW weight.  stacked != 0: src is the result
 * [N][Hc][Wc][4][Cin] of ONE forward convolution Cout -> 4 * Cin with the whole of wc as its weight (class-major output
 * channels) instead of four class maps. */
int pcg_pack_dgrad_classes(const float* w, int Cout, int Cin, int k, float* wc, void* stream);
int pcg_parity_interleave(const float* src, int N, int Hc, int Wc, int C, int pad, int H, int W, int stacked, float* dx,
                          void* stream);
/* Gradient penalty of mnist_wgan_conditional.py:147 on the critic's input gradient g [B][D]: n_b = ||g[b]||_2,
 * out[0] = lambda * mean_b (n_b - 1)^2, gbar = its cotangent lambda * 2 (n_b - 1) / (B n_b) * g[b], norms[b] = n_b
 * (NULL: not stored).  Not re-entrant across streams (one internal arrival counter). */
int pcg_gp_penalty(const float* g, int B, int D, float lambda, float* out, float* gbar, float* norms, void* stream);
int pcg_cf_apply(const float* x, const float* residual, long long n, float lo, float hi, float* x_cf, float* scratch,
                 void* stream);
int pcg_cf_metrics(const float* logits, const long long* y_true, const long long* y_target, int B, int NC,
                   const float* scratch, long long n_elems, float* out3, void* stream);
/* Keep-mask of nn.Dropout(p) / nn.Dropout2d(p) in training mode (conditional_counteRGAN/mnist/models/classifier.py:14,19,
 * house_sales_kc_usa/models/nn_classifier.py): mask[rows][inner][C] (NHWC: inner = H*W) = Bernoulli(1 - p) / (1 - p);
 * channelwise != 0 draws once per (row, c) and repeats it over `inner` (Dropout2d zeroes whole feature maps).  The
 * activation is multiplied by the mask in the forward pass and its gradient in the backward pass.  rng_state as in
 * pcg_build_mask. */
int pcg_dropout_mask(long long rows, int inner, int C, float p, int channelwise, unsigned long long seed,
                     unsigned long long* rng_state, float* mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCG_H */

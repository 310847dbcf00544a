// Small fp32 primitives for the MLP / tabular / DCGAN step plans — host interface (see ops.cu).
#pragma once
#include "common.cuh"

namespace pcg {

enum UnaryOp { OP_RELU = 1, OP_LRELU = 2, OP_SIGMOID = 3, OP_TANH = 4, OP_SCALE = 5, OP_COPY = 6 };
enum BinaryOp { OP_ADD = 0, OP_MUL = 1 };

void unary(const float* x, long long n, int op, float a, float* y, cudaStream_t s);
void unary_bwd(const float* dy, const float* y, long long n, int op, float a, float* dx, cudaStream_t s);
void binary(const float* a, const float* b, long long n, int op, float alpha, float beta, float* out, cudaStream_t s);
void film_fwd(const float* gamma, const float* n_, const float* beta, const float* res, long long n, int relu, float* out,
              cudaStream_t s);
void film_bwd(const float* df, const float* gamma, const float* n_, long long n, int accumulate, float* dn, float* dgamma,
              float* dbeta, cudaStream_t s);
constexpr int TRANSPOSE_MAX = 64;
struct TransposeTable {
  int n;
  const float* src[TRANSPOSE_MAX];     // [rows][cols]
  float* dst[TRANSPOSE_MAX];           // [cols][rows]
  int rows[TRANSPOSE_MAX], cols[TRANSPOSE_MAX], begin[TRANSPOSE_MAX + 1];
};
void transpose_multi(const TransposeTable& t, cudaStream_t s);
void copy_cols(const float* src, int src_ld, int c0_src, float* dst, int dst_ld, int c0_dst, long long rows, int ncols,
               float alpha, int accumulate, cudaStream_t s);
void onehot(const long long* lab, long long rows, int nc, float* dst, int dst_ld, int c0, cudaStream_t s);
void reduce_scalar(const float* x, long long n, int absval, float scale, float* out, float gscale, float* dx,
                   cudaStream_t s);
void rownorm_mean(const float* x, long long rows, int cols, int p, float* out, float gscale, float* dx, cudaStream_t s);
void gan_loss(const float* z, int n, int kind, float t, float wgt, float* out_loss, float* out_aux, float* dz,
              cudaStream_t s);
struct ScalarTerms {
  int n;
  float c[6];
  const float* p[6];
};
void combine_scalars(const ScalarTerms& t, float* out, cudaStream_t s);
void spectral_norm_fwd(const float* W, int N, int K, float* u, float* v, float eps, int do_iter, float* Wn,
                       float* sigma, cudaStream_t s, float* WnT = nullptr, float* us = nullptr, float* vs = nullptr);
void spectral_norm_bwd(const float* dWn, const float* Wn, int N, int K, const float* u, const float* v,
                       const float* sigma, float* dW, cudaStream_t s);
// Whole iteration of the two-layer MLP GANs (make_moons_cgan.py:90-129, make_moons_gan.py:61-88) in one cluster launch
// (mlp_gan.cu); parameter / gradient / moment buffers use the flat layout of ops.py FlatParams (slices padded to 4).
void mlp_gan_step(int B, int z_dim, int label_dim, int hidden, const float* real, const float* real_oh, const float* z1,
                  const float* oh1, const float* z2, const float* oh2, float* g_param, float* g_grad, float* g_m,
                  float* g_v, int* g_step, float* d_param, float* d_grad, float* d_m, float* d_v, int* d_step, float lr,
                  float* scal, cudaStream_t stream);
void gumbel_softmax_fwd(const float* logits, const float* g, long long rows, int n, float tau, float* y, cudaStream_t s);
void onehot_argmax(const float* x, long long rows, int n, float* y, cudaStream_t s);
void softmax_bwd(const float* dy, const float* y, long long rows, int n, float tau, float* dl, cudaStream_t s);
void bn_eval(const float* x, long long rows, int C, const float* gamma, const float* beta, const float* rm,
             const float* rv, float eps, float* y, float* scale_out, cudaStream_t s);
void scale_cols(const float* dy, long long rows, int C, const float* scale, float* dx, cudaStream_t s);
// x[b][HW] = (float(images[index[b]][.]) / 255 - mean) / std ; y[b] = labels[index[b]]  (index == nullptr: identity)
void u8_batch(const uint8_t* images, const long long* labels, const long long* index, int B, int HW, float mean, float stdv,
              float* x, long long* y, cudaStream_t s);
// mask[B][C][H][W] in {0,1}: k_sel of the (H/patch)*(W/patch) patches per sample chosen uniformly at random (k_sel < 0 or
// >= total: independent fair coins), nearest-upsampled; target[b] ~ U{0..num_classes-1} (nullptr: no draw).
// rng = device {stream offset, ticket} (advanced by the kernel) or nullptr (offset 0).
// counterfactual evaluation: x_cf = clamp(x + r), partials of sum |x_cf - x| (cf_parts() floats); then flip rate,
// prediction gain, actionability from the classifier's logits on x_cf
int cf_parts();
void cf_apply(const float* x, const float* r, long long n, float lo, float hi, float* x_cf, float* part, cudaStream_t s);
void cf_metrics(const float* logits, const long long* y_true, const long long* y_target, int B, int NC, const float* part,
                long long n_elems, float* out, cudaStream_t s);
// mask[rows][inner][C] = Bernoulli(1 - p) / (1 - p) (nn.Dropout); channelwise: one draw per (row, c) (nn.Dropout2d)
void dropout_mask(long long rows, int inner, int C, float p, int channelwise, unsigned long long seed, unsigned long long* rng,
                  float* mask, cudaStream_t s);
void build_mask(int B, int C, int H, int W, int patch, int k_sel, int num_classes, unsigned long long seed,
                unsigned long long* rng, float* mask, long long* target, cudaStream_t s);

}  // namespace pcg

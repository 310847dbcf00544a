// Small fp32 primitives used by the MLP / tabular / DCGAN step plans (conditional_gan/moons,
// simple_gan/moons, conditional_counteRGAN/{moons,house_sales_kc_usa}, dconv_gan/mnist).
// Row-major [rows][cols] matrices.  Every kernel here is latency/HBM-bound and tiny at the reference's
// problem sizes; the step plans capture them in a CUDA graph.
#include "ops.cuh"

#include "cluster_reduce.cuh"

namespace pcg {

static int blocks_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}
#define GRID_STRIDE(i, n) \
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

// ------------------------------------------------------------------ elementwise
__global__ void unary_kernel(const float* __restrict__ x, long long n, int op, float a, float* __restrict__ y) {
  pdl_enter();
  GRID_STRIDE(i, n) {
    const float v = x[i];
    float r;
    switch (op) {
      case OP_RELU: r = fmaxf(v, 0.f); break;
      case OP_LRELU: r = v > 0.f ? v : v * a; break;
      case OP_SIGMOID: r = 1.f / (1.f + expf(-v)); break;
      case OP_TANH: r = tanhf(v); break;
      case OP_SCALE: r = v * a; break;
      case OP_COPY: r = v; break;
      default: r = v;
    }
    y[i] = r;
  }
}
void unary(const float* x, long long n, int op, float a, float* y, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(unary_kernel, dim3(blocks_for(n)), dim3(256), 0, s, x, n, op, a, y);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// dx = dy * f'(.) evaluated from the activation OUTPUT y
__global__ void unary_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n, int op,
                                 float a, float* __restrict__ dx) {
  pdl_enter();
  GRID_STRIDE(i, n) {
    const float o = y[i], g = dy[i];
    float r;
    switch (op) {
      case OP_RELU: r = o > 0.f ? g : 0.f; break;
      case OP_LRELU: r = o > 0.f ? g : g * a; break;
      case OP_SIGMOID: r = g * o * (1.f - o); break;
      case OP_TANH: r = g * (1.f - o * o); break;
      case OP_SCALE: r = g * a; break;
      default: r = g;
    }
    dx[i] = r;
  }
}
void unary_bwd(const float* dy, const float* y, long long n, int op, float a, float* dx, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(unary_bwd_kernel, dim3(blocks_for(n)), dim3(256), 0, s, dy, y, n, op, a, dx);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// out = alpha * a (op) b   with op in {add, mul}; b may broadcast over rows when b_cols > 0 (b is [cols])
__global__ void binary_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int op,
                              float alpha, float beta, float* __restrict__ out) {
  pdl_enter();
  GRID_STRIDE(i, n) {
    const float x = a[i], y = b[i];
    out[i] = op == OP_MUL ? alpha * x * y : alpha * x + beta * y;
  }
}
void binary(const float* a, const float* b, long long n, int op, float alpha, float beta, float* out, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(binary_kernel, dim3(blocks_for(n)), dim3(256), 0, s, a, b, n, op, alpha, beta, out);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// FiLM (house_sales_kc_usa/models/generator.py:13-16,28-35): one launch for  out = [relu](gamma * n + beta) [+ res]
__global__ void film_fwd_kernel(const float* __restrict__ gamma, const float* __restrict__ n_, const float* __restrict__ beta,
                                const float* __restrict__ res, long long n, int relu, float* __restrict__ out) {
  pdl_enter();
  GRID_STRIDE(i, n) {
    float v = fmaf(gamma[i], n_[i], beta[i]);
    if (relu) v = fmaxf(v, 0.f);
    out[i] = res != nullptr ? res[i] + v : v;
  }
}
void film_fwd(const float* gamma, const float* n_, const float* beta, const float* res, long long n, int relu, float* out,
              cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(film_fwd_kernel, dim3(blocks_for(n)), dim3(256), 0, s, gamma, n_, beta, res, n, relu, out);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
// ... and its backward for an upstream gradient df:  dn = df * gamma ;  dgamma (+)= df * n ;  dbeta (+)= df
__global__ void film_bwd_kernel(const float* __restrict__ df, const float* __restrict__ gamma, const float* __restrict__ n_,
                                long long n, int accumulate, float* __restrict__ dn, float* __restrict__ dgamma,
                                float* __restrict__ dbeta) {
  pdl_enter();
  GRID_STRIDE(i, n) {
    const float d = df[i];
    dn[i] = d * gamma[i];
    const float dg = d * n_[i];
    dgamma[i] = accumulate ? dgamma[i] + dg : dg;
    dbeta[i] = accumulate ? dbeta[i] + d : d;
  }
}
void film_bwd(const float* df, const float* gamma, const float* n_, long long n, int accumulate, float* dn, float* dgamma,
              float* dbeta, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(film_bwd_kernel, dim3(blocks_for(n)), dim3(256), 0, s, df, gamma, n_, n, accumulate, dn, dgamma, dbeta);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// dst_i = src_i^T for up to TRANSPOSE_MAX small matrices in one launch (the dgrad operands of every Linear layer of a
// tabular net after its Adam update)
__global__ void transpose_multi_kernel(const TransposeTable t) {
  pdl_enter();
  GRID_STRIDE(g, (long long)t.begin[t.n]) {
    int l = 0;
    while (l + 1 < t.n && g >= t.begin[l + 1]) ++l;
    const int i = (int)g - t.begin[l];
    const int r = i / t.cols[l], c = i - r * t.cols[l];
    t.dst[l][(size_t)c * t.rows[l] + r] = t.src[l][i];
  }
}
void transpose_multi(const TransposeTable& t, cudaStream_t s) {
  PCG_PROFILE("pack_weights", s);
  launch_k(transpose_multi_kernel, dim3(blocks_for(t.begin[t.n])), dim3(256), 0, s, t);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// out[r][c0_out + j] = src[r][c0_src + j], j < ncols  (column block copy between matrices of different widths);
// accumulate != 0 adds instead of overwriting (gradient of a tensor used twice).
__global__ void copy_cols_kernel(const float* __restrict__ src, int src_ld, int c0_src, float* __restrict__ dst,
                                 int dst_ld, int c0_dst, long long rows, int ncols, float alpha, int accumulate) {
  pdl_enter();
  const long long n = rows * ncols;
  GRID_STRIDE(i, n) {
    const long long r = i / ncols;
    const int j = (int)(i - r * ncols);
    const float v = alpha * src[r * src_ld + c0_src + j];
    float* d = dst + r * dst_ld + c0_dst + j;
    *d = accumulate ? *d + v : v;
  }
}
void copy_cols(const float* src, int src_ld, int c0_src, float* dst, int dst_ld, int c0_dst, long long rows, int ncols,
               float alpha, int accumulate, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(copy_cols_kernel, dim3(blocks_for(rows * ncols)), dim3(256), 0, s, src, src_ld, c0_src, dst, dst_ld, c0_dst, rows, ncols, alpha,
                                                           accumulate);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

__global__ void onehot_kernel(const long long* __restrict__ lab, long long rows, int nc, float* __restrict__ dst,
                              int dst_ld, int c0) {
  pdl_enter();
  const long long n = rows * nc;
  GRID_STRIDE(i, n) {
    const long long r = i / nc;
    const int j = (int)(i - r * nc);
    dst[r * dst_ld + c0 + j] = lab[r] == j ? 1.f : 0.f;
  }
}
void onehot(const long long* lab, long long rows, int nc, float* dst, int dst_ld, int c0, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(onehot_kernel, dim3(blocks_for(rows * nc)), dim3(256), 0, s, lab, rows, nc, dst, dst_ld, c0);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------ reductions to a scalar (single block, fixed order)
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;
}

// out = scale * sum_i f(x_i), f in {id, abs}; optional dx_i = gscale * f'(x_i)   (sign(0) = 0 as torch.abs backward)
__global__ void reduce_scalar_kernel(const float* __restrict__ x, long long n, int absval, float scale, float* out,
                                     float gscale, float* __restrict__ dx) {
  pdl_enter();
  __shared__ float red[32];
  __shared__ float slot;
  long long i0, i1;
  cluster_slice(n, i0, i1);
  float s = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const float v = x[i];
    s += absval ? fabsf(v) : v;
    if (dx) dx[i] = absval ? gscale * ((v > 0.f) - (v < 0.f)) : gscale;
  }
  s = cluster_total(block_sum(s, red), &slot);
  if (cluster_leader()) out[0] = s * scale;
}
void reduce_scalar(const float* x, long long n, int absval, float scale, float* out, float gscale, float* dx,
                   cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k_cluster(reduce_scalar_kernel, n / 8, dim3(1024), 0, s, x, n, absval, scale, out, gscale, dx);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// out = mean_r ||x_r||_p (p = 1 or 2); dx_r = gscale/rows * d||x_r||_p / dx   (zero sub-gradient at x_r = 0 for p=2)
__global__ void rownorm_mean_kernel(const float* __restrict__ x, long long rows, int cols, int p, float* out,
                                    float gscale, float* __restrict__ dx) {
  pdl_enter();
  __shared__ float red[32];
  __shared__ float slot;
  long long rb, re;
  cluster_slice(rows, rb, re);
  float s = 0.f;
  for (long long r = rb + threadIdx.x; r < re; r += blockDim.x) {
    float a = 0.f;
    for (int c = 0; c < cols; ++c) {
      const float v = x[r * cols + c];
      a += p == 1 ? fabsf(v) : v * v;
    }
    const float nrm = p == 1 ? a : sqrtf(a);
    s += nrm;
    if (dx) {
      for (int c = 0; c < cols; ++c) {
        const float v = x[r * cols + c];
        float g;
        if (p == 1) g = (float)((v > 0.f) - (v < 0.f));
        else g = nrm > 0.f ? v / nrm : 0.f;
        dx[r * cols + c] = gscale * g / (float)rows;
      }
    }
  }
  s = cluster_total(block_sum(s, red), &slot);
  if (cluster_leader()) out[0] = s / (float)rows;
}
void rownorm_mean(const float* x, long long rows, int cols, int p, float* out, float gscale, float* dx, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k_cluster(rownorm_mean_kernel, rows, dim3(rows >= CR_MIN_ITEMS ? 512 : 1024), 0, s, x, rows, cols, p, out, gscale, dx);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------ GAN losses on discriminator outputs
// kind 0 (log / saturating GAN, make_moons_gan.py:70,81): z are pre-sigmoid logits, p = sigmoid(z)
//        term t=1: loss = -mean log p       dz = wgt * -(1 - p) / n
//        term t=0: loss = -mean log(1 - p)  dz = wgt *  p / n
// kind 1 (BCELoss on probabilities, mnist_dcgan.py:125): same formulas with torch's log clamp at -100
// kind 2 (Wasserstein, moons/trainer.py:77,83): loss = sign * mean z ; dz = wgt * sign / n   (t=1 -> sign=-1)
__global__ void gan_loss_kernel(const float* __restrict__ z, int n, int kind, float t, float wgt, float* out_loss,
                                float* out_aux, float* __restrict__ dz) {
  pdl_enter();
  __shared__ float red[32];
  __shared__ float slot;
  long long i0, i1;
  cluster_slice(n, i0, i1);
  float sl = 0.f, sp = 0.f;
  for (int i = (int)i0 + threadIdx.x; i < (int)i1; i += blockDim.x) {
    const float v = z[i];
    if (kind == 2) {
      const float sign = t > 0.5f ? -1.f : 1.f;
      sl += sign * v;
      sp += 1.f / (1.f + expf(-v));
      dz[i] = wgt * sign / (float)n;
    } else {
      const float p = 1.f / (1.f + expf(-v));
      float lp = logf(p), l1p = logf(1.f - p);
      if (kind == 1) { lp = fmaxf(lp, -100.f); l1p = fmaxf(l1p, -100.f); }
      sl += t > 0.5f ? -lp : -l1p;
      sp += p;
      dz[i] = wgt * (t > 0.5f ? -(1.f - p) : p) / (float)n;
    }
  }
  sl = cluster_total(block_sum(sl, red), &slot);
  sp = cluster_total(block_sum(sp, red), &slot);
  if (cluster_leader()) {
    out_loss[0] = sl / (float)n;
    if (out_aux) out_aux[0] = sp / (float)n;
  }
}
void gan_loss(const float* z, int n, int kind, float t, float wgt, float* out_loss, float* out_aux, float* dz,
              cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k_cluster(gan_loss_kernel, n, dim3(n >= CR_MIN_ITEMS ? 512 : 256), 0, s, z, n, kind, t, wgt, out_loss, out_aux, dz);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// out[0] = sum_i c_i * in_i[0]  (up to 6 scalar terms)
__global__ void combine_kernel(ScalarTerms t, float* out) {
  pdl_enter();
  float s = 0.f;
  for (int i = 0; i < t.n; ++i) s += t.c[i] * t.p[i][0];
  out[0] = s;
}
void combine_scalars(const ScalarTerms& t, float* out, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(combine_kernel, dim3(1), dim3(1), 0, s, t, out);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------ spectral norm (torch/nn/utils/spectral_norm.py:62-113)
// One power iteration in place on (u, v) (train mode), sigma = u^T W v, Wn = W / sigma.  Single block: the
// matrices are at most 128 x 64.
// Optional extra outputs save three launches per layer and pass: WnT = Wn^T ([K][N], the dgrad operand) and the
// snapshots us / vs of u / v that this pass's backward needs (torch's graph holds clones taken at call time).
// W is staged in shared memory once ([N][K + 1]); W^T u runs a thread per column, W v a warp per row (the product is
// reused for the u update and for sigma); the first version read W from global memory three times with one thread per
// output and a sequential inner loop: 25 us for the 128 x 64 layer, on the critical path behind the critic's Adam.
__global__ void __launch_bounds__(256)
spectral_norm_fwd_kernel(const float* __restrict__ W, int N, int K, float* u, float* v, float eps, int do_iter,
                         float* __restrict__ Wn, float* sigma_out, float* __restrict__ WnT, float* __restrict__ us,
                         float* __restrict__ vs) {
  pdl_enter();
  extern __shared__ float sm[];     // u[N], v[K], a[N] = W v, red[40], W[N][K + 1]
  float* su = sm;
  float* sv = sm + N;
  float* sa = sv + K;
  float* red = sa + N;
  float* sW = red + 40;
  const int ld = K + 1;
  __shared__ float bc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i - n * K;
    sW[n * ld + k] = W[i];
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) su[i] = u[i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) sv[i] = v[i];
  __syncthreads();
  if (do_iter) {
    // v = normalize(W^T u)
    float part = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float a = 0.f;
      for (int n = 0; n < N; ++n) a = fmaf(sW[n * ld + k], su[n], a);
      sv[k] = a;
      part += a * a;
    }
    const float nrm = sqrtf(block_sum(part, red));
    if (threadIdx.x == 0) bc = fmaxf(nrm, eps);
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) sv[k] /= bc;
    __syncthreads();
  }
  // a = W v (a warp per row)
  for (int n = warp; n < N; n += (blockDim.x >> 5)) {
    float a = 0.f;
    for (int k = lane; k < K; k += 32) a = fmaf(sW[n * ld + k], sv[k], a);
    a = warp_sum(a);
    if (lane == 0) sa[n] = a;
  }
  __syncthreads();
  if (do_iter) {
    // u = normalize(W v)
    float part = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) part += sa[n] * sa[n];
    const float nrm = sqrtf(block_sum(part, red));
    if (threadIdx.x == 0) bc = fmaxf(nrm, eps);
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) su[n] = sa[n] / bc;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) u[i] = su[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) v[i] = sv[i];
  }
  // sigma = u^T (W v)
  float part = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) part = fmaf(su[n], sa[n], part);
  const float sg = block_sum(part, red);
  __shared__ float s_sigma;
  if (threadIdx.x == 0) { s_sigma = sg; sigma_out[0] = sg; }
  __syncthreads();
  const float inv = 1.f / s_sigma;
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i - n * K;
    Wn[i] = sW[n * ld + k] * inv;
  }
  if (WnT != nullptr)
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {      // coalesced writes of the transpose, reads from shared memory
      const int k = i / N, n = i - k * N;
      WnT[i] = sW[n * ld + k] * inv;
    }
  if (us != nullptr)
    for (int i = threadIdx.x; i < N; i += blockDim.x) us[i] = su[i];
  if (vs != nullptr)
    for (int i = threadIdx.x; i < K; i += blockDim.x) vs[i] = sv[i];
}
void spectral_norm_fwd(const float* W, int N, int K, float* u, float* v, float eps, int do_iter, float* Wn,
                       float* sigma, cudaStream_t s, float* WnT, float* us, float* vs) {
  PCG_PROFILE("ops_small", s);
  const size_t sm = (size_t)(2 * N + K + 40 + (size_t)N * (K + 1)) * sizeof(float);
  PCG_REQUIRE(sm <= 200 * 1024, "spectral_norm_fwd: the weight matrix must fit in shared memory (N * K <= ~50 K)");
  if (sm > 48 * 1024)
    PCG_CHECK_CUDA(cudaFuncSetAttribute(spectral_norm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  launch_k(spectral_norm_fwd_kernel, dim3(1), dim3(256), sm, s, W, N, K, u, v, eps, do_iter, Wn, sigma, WnT, us, vs);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
// dW = (dWn - (sum dWn .* Wn) * u v^T) / sigma      (u, v treated as constants, as torch does)
__global__ void spectral_norm_bwd_kernel(const float* __restrict__ dWn, const float* __restrict__ Wn, int N, int K,
                                         const float* __restrict__ u, const float* __restrict__ v,
                                         const float* __restrict__ sigma, float* __restrict__ dW) {
  pdl_enter();
  __shared__ float red[32];
  __shared__ float s_dot;
  float part = 0.f;
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) part = fmaf(dWn[i], Wn[i], part);
  const float dot = block_sum(part, red);
  if (threadIdx.x == 0) s_dot = dot;
  __syncthreads();
  const float inv = 1.f / sigma[0];
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i - n * K;
    dW[i] = (dWn[i] - s_dot * u[n] * v[k]) * inv;
  }
}
void spectral_norm_bwd(const float* dWn, const float* Wn, int N, int K, const float* u, const float* v,
                       const float* sigma, float* dW, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(spectral_norm_bwd_kernel, dim3(1), dim3(256), 0, s, dWn, Wn, N, K, u, v, sigma, dW);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------ Gumbel-softmax (soft), noise injected
// y = softmax((logits + g) / tau) ; backward dlogits = y .* (dy - sum_j dy_j y_j) / tau
__global__ void gumbel_softmax_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ g, long long rows,
                                          int n, float tau, float* __restrict__ y) {
  pdl_enter();
  GRID_STRIDE(r, rows) {
    const float* l = logits + r * n;
    const float* gg = g + r * n;
    float mx = -INFINITY;
    for (int j = 0; j < n; ++j) mx = fmaxf(mx, (l[j] + gg[j]) / tau);
    float se = 0.f;
    for (int j = 0; j < n; ++j) se += expf((l[j] + gg[j]) / tau - mx);
    for (int j = 0; j < n; ++j) y[r * n + j] = expf((l[j] + gg[j]) / tau - mx) / se;
  }
}
void gumbel_softmax_fwd(const float* logits, const float* g, long long rows, int n, float tau, float* y, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(gumbel_softmax_fwd_kernel, dim3(blocks_for(rows)), dim3(256), 0, s, logits, g, rows, n, tau, y);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
// y = one_hot(argmax_j x[r][j]): the forward value of F.gumbel_softmax(..., hard=True) given its soft sample
// (house_sales_kc_usa/eval_utils.py:75: the evaluation path asks the generator for hard categorical samples)
__global__ void onehot_argmax_kernel(const float* __restrict__ x, long long rows, int n, float* __restrict__ y) {
  pdl_enter();
  GRID_STRIDE(r, rows) {
    int am = 0;
    float mx = x[r * n];
    for (int j = 1; j < n; ++j)
      if (x[r * n + j] > mx) { mx = x[r * n + j]; am = j; }
    for (int j = 0; j < n; ++j) y[r * n + j] = j == am ? 1.f : 0.f;
  }
}
void onehot_argmax(const float* x, long long rows, int n, float* y, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(onehot_argmax_kernel, dim3(blocks_for(rows)), dim3(256), 0, s, x, rows, n, y);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
__global__ void softmax_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long rows, int n,
                                   float tau, float* __restrict__ dl) {
  pdl_enter();
  GRID_STRIDE(r, rows) {
    float dot = 0.f;
    for (int j = 0; j < n; ++j) dot = fmaf(dy[r * n + j], y[r * n + j], dot);
    for (int j = 0; j < n; ++j) dl[r * n + j] = y[r * n + j] * (dy[r * n + j] - dot) / tau;
  }
}
void softmax_bwd(const float* dy, const float* y, long long rows, int n, float tau, float* dl, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(softmax_bwd_kernel, dim3(blocks_for(rows)), dim3(256), 0, s, dy, y, rows, n, tau, dl);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------ BatchNorm in eval mode (running stats) as affine
__global__ void bn_eval_kernel(const float* __restrict__ x, long long rows, int C, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ rm,
                               const float* __restrict__ rv, float eps, float* __restrict__ y, float* __restrict__ scale_out) {
  pdl_enter();
  const long long n = rows * C;
  GRID_STRIDE(i, n) {
    const int c = (int)(i % C);
    const float sc = gamma[c] * rsqrtf(rv[c] + eps);
    y[i] = (x[i] - rm[c]) * sc + beta[c];
    if (scale_out && i < C) scale_out[c] = sc;
  }
}
void bn_eval(const float* x, long long rows, int C, const float* gamma, const float* beta, const float* rm,
             const float* rv, float eps, float* y, float* scale_out, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(bn_eval_kernel, dim3(blocks_for(rows * C)), dim3(256), 0, s, x, rows, C, gamma, beta, rm, rv, eps, y, scale_out);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
// dx = dy * scale[c]
__global__ void scale_cols_kernel(const float* __restrict__ dy, long long rows, int C, const float* __restrict__ scale,
                                  float* __restrict__ dx) {
  pdl_enter();
  const long long n = rows * C;
  GRID_STRIDE(i, n) dx[i] = dy[i] * scale[i % C];
}
void scale_cols(const float* dy, long long rows, int C, const float* scale, float* dx, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  launch_k(scale_cols_kernel, dim3(blocks_for(rows * C)), dim3(256), 0, s, dy, rows, C, scale, dx);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ---- on-device input pipeline (conditional_counteRGAN/mnist/data_utils.py:9-12,26): the uint8 dataset stays resident
// in HBM; a batch is gathered through an index vector and normalised exactly as torchvision does it:
// ToTensor = float(u8) / 255 (true division), Normalize = (x - mean) / std.  16 pixels (one 16-byte load) per thread.
__global__ void u8_batch_kernel(const uint8_t* __restrict__ images, const long long* __restrict__ labels,
                                const long long* __restrict__ index, int B, int HW, float mean, float stdv,
                                float* __restrict__ x, long long* __restrict__ y) {
  pdl_enter();
  const int v16 = HW / 16;                                   // HW % 16 == 0 (784 = 49 * 16)
  const long long total = (long long)B * v16;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / v16), c = (int)(i - (long long)b * v16);
    const long long src = index != nullptr ? index[b] : b;
    const uint4 u = *reinterpret_cast<const uint4*>(images + src * HW + c * 16);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float4* dst = reinterpret_cast<float4*>(x + (long long)b * HW + c * 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        f[j] = __fdiv_rn(__fdiv_rn((float)((w[k] >> (8 * j)) & 0xffu), 255.f) - mean, stdv);
      dst[k] = make_float4(f[0], f[1], f[2], f[3]);
    }
    if (c == 0 && y != nullptr) y[b] = labels[src];
  }
}
void u8_batch(const uint8_t* images, const long long* labels, const long long* index, int B, int HW, float mean, float stdv,
              float* x, long long* y, cudaStream_t s) {
  PCG_PROFILE("input_pipeline", s);
  PCG_REQUIRE(HW % 16 == 0 && (reinterpret_cast<uintptr_t>(images) & 15) == 0, "image size must be a multiple of 16 bytes");
  launch_k(u8_batch_kernel, dim3(blocks_for((long long)B * (HW / 16))), dim3(256), 0, s, images, labels, index, B, HW, mean, stdv, x, y);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ---- random patch mask + target draw (conditional_counteRGAN/mnist/trainer.py:45-72 build_mask, :94 target_y) -------
// Philox4x32-10 counter-based generator (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11): counter =
// (sample index, draw index, stream offset lo, hi), key = seed.  The stream offset lives in device memory (rng[0]) and
// is advanced by the last block to finish, so the launch can sit inside a replayed CUDA graph and still draw fresh
// numbers every replay.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// uniform integer in [0, n) from a 32-bit draw (multiply-shift; bias < n / 2^32)
__device__ __forceinline__ uint32_t bounded(uint32_t r, uint32_t n) { return __umulhi(r, n); }

constexpr int MASK_SPB = 8;          // samples per block
__global__ void build_mask_kernel(int B, int C, int H, int W, int nph, int npw, int k_sel, int num_classes,
                                  unsigned long long seed, unsigned long long* __restrict__ rng,
                                  float* __restrict__ mask, long long* __restrict__ target) {
  pdl_enter();
  __shared__ unsigned long long words[MASK_SPB];
  const unsigned long long offset = rng != nullptr ? rng[0] : 0ull;
  if (rng != nullptr) seed += rng[2];                        // the key may live in device memory too (graph replays)
  const int b0 = blockIdx.x * MASK_SPB;
  const int total = nph * npw;
  if (threadIdx.x < MASK_SPB && b0 + (int)threadIdx.x < B) {
    const int b = b0 + threadIdx.x;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t olo = (uint32_t)offset, ohi = (uint32_t)(offset >> 32);
    unsigned long long word = 0ull;
    if (k_sel < 0 || k_sel >= total) {
      // trainer.py:59-61: every patch an independent fair coin
      for (int i = 0; i < total; i += 32) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)b, 1u + (uint32_t)(i >> 5), olo, ohi), key);
        word |= (unsigned long long)r.x << i;
      }
      if (total < 64) word &= (1ull << total) - 1ull;
    } else {
      // trainer.py:63-65: randperm(total)[:k] = a uniformly random k-subset; partial Fisher-Yates over a patch list
      // packed in registers is replaced by selection sampling on the bit set: the j-th pick is the (r mod remaining)-th
      // still-free patch, which is the same distribution
      unsigned long long free_bits = total < 64 ? (1ull << total) - 1ull : ~0ull;
      uint4 r = make_uint4(0, 0, 0, 0);
      for (int j = 0; j < k_sel; ++j) {
        if ((j & 3) == 0) r = philox4x32_10(make_uint4((uint32_t)b, 1u + (uint32_t)(j >> 2), olo, ohi), key);
        const uint32_t draw = (j & 3) == 0 ? r.x : (j & 3) == 1 ? r.y : (j & 3) == 2 ? r.z : r.w;
        int nth = (int)bounded(draw, (uint32_t)(total - j));
        unsigned long long f = free_bits;
        for (int t = 0; t < nth; ++t) f &= f - 1ull;           // drop the nth lowest free bits
        const unsigned long long pick = f & (~f + 1ull);       // lowest remaining free bit
        word |= pick;
        free_bits &= ~pick;
      }
    }
    words[threadIdx.x] = word;
    if (target != nullptr) {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)b, 0u, olo, ohi), key);
      target[b] = (long long)bounded(r.x, (uint32_t)num_classes);           // trainer.py:94 randint(0, num_classes)
    }
  }
  __syncthreads();
  // nearest up-sampling (F.interpolate(..., size=(h, w), mode="nearest"), trainer.py:68-70: source = floor(dst * in / out))
  // and the repeat over channels; one float4 (four pixels of a row) per thread iteration when W % 4 == 0
  const int nb = min(MASK_SPB, B - b0);
  const long long per = (long long)C * H * W;
  if ((W & 3) == 0) {
    const long long n4 = nb * per / 4;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      const long long e = i * 4;
      const int sb = (int)(e / per);
      const int rem = (int)(e - sb * per);
      const int hw = rem % (H * W), h = hw / W, w0 = hw - h * W;
      const unsigned long long word = words[sb];
      const int ph = min((h * nph) / H, nph - 1);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pw = min(((w0 + j) * npw) / W, npw - 1);
        v[j] = (float)((word >> (ph * npw + pw)) & 1ull);
      }
      *reinterpret_cast<float4*>(mask + (long long)b0 * per + e) = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    for (long long e = threadIdx.x; e < nb * per; e += blockDim.x) {
      const int sb = (int)(e / per);
      const int rem = (int)(e - sb * per);
      const int hw = rem % (H * W), h = hw / W, w = hw - h * W;
      const int ph = min((h * nph) / H, nph - 1), pw = min((w * npw) / W, npw - 1);
      mask[(long long)b0 * per + e] = (float)((words[sb] >> (ph * npw + pw)) & 1ull);
    }
  }
  if (rng != nullptr) {
    // advance the stream offset once per launch: the last block to arrive does it (every block read rng[0] above)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long ticket = atomicAdd(rng + 1, 1ull);
      if (ticket == (unsigned long long)gridDim.x - 1ull) {
        rng[1] = 0ull;
        rng[0] = offset + 1ull;
        __threadfence();
      }
    }
  }
}
// ---- counterfactual evaluation (conditional_counteRGAN/mnist/eval_utils.py:46-75) -----------------------------------
// x_cf = clamp(x + residual, lo, hi) and the per-block partial of sum |x_cf - x| (actionability numerator)
__global__ void cf_apply_kernel(const float* __restrict__ x, const float* __restrict__ r, long long n, float lo, float hi,
                                float* __restrict__ x_cf, float* __restrict__ part) {
  pdl_enter();
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = x[i];
    const float c = fminf(fmaxf(xv + r[i], lo), hi);
    x_cf[i] = c;
    s += fabsf(c - xv);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    part[blockIdx.x] = t;
  }
}
// out[0] = class-flip rate (arg-max of logits == target), out[1] = mean(softmax[target] - softmax[true]),
// out[2] = mean |x_cf - x| from the partials of cf_apply_kernel.  One block, fixed order.
__global__ void cf_metrics_kernel(const float* __restrict__ logits, const long long* __restrict__ y_true,
                                  const long long* __restrict__ y_target, int B, int NC, const float* __restrict__ part,
                                  int nparts, long long n_elems, float* __restrict__ out) {
  pdl_enter();
  __shared__ double red[3][8];
  double flips = 0.0, gain = 0.0, act = 0.0;
  for (int n = threadIdx.x; n < B; n += blockDim.x) {
    const float* l = logits + (size_t)n * NC;
    float mx = l[0];
    int am = 0;
    for (int j = 1; j < NC; ++j)
      if (l[j] > mx) { mx = l[j]; am = j; }
    float se = 0.f;
    for (int j = 0; j < NC; ++j) se += expf(l[j] - mx);
    const int t = (int)y_target[n], y = (int)y_true[n];
    flips += am == t ? 1.0 : 0.0;
    gain += (double)((expf(l[t] - mx) - expf(l[y] - mx)) / se);
  }
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) act += (double)part[i];
  flips = warp_sum(flips); gain = warp_sum(gain); act = warp_sum(act);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = flips; red[1][threadIdx.x >> 5] = gain; red[2][threadIdx.x >> 5] = act; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
    out[0] = (float)(a / (double)B);
    out[1] = (float)(b / (double)B);
    out[2] = (float)(c / (double)n_elems);
  }
}
constexpr int CF_PARTS = 296;
void cf_apply(const float* x, const float* r, long long n, float lo, float hi, float* x_cf, float* part, cudaStream_t s) {
  PCG_PROFILE("cf_eval", s);
  launch_k(cf_apply_kernel, dim3(CF_PARTS), dim3(256), 0, s, x, r, n, lo, hi, x_cf, part);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
void cf_metrics(const float* logits, const long long* y_true, const long long* y_target, int B, int NC, const float* part,
                long long n_elems, float* out, cudaStream_t s) {
  PCG_PROFILE("cf_eval", s);
  launch_k(cf_metrics_kernel, dim3(1), dim3(256), 0, s, logits, y_true, y_target, B, NC, part, CF_PARTS, n_elems, out);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
int cf_parts() { return CF_PARTS; }

// ---- dropout keep-mask (nn.Dropout / nn.Dropout2d in training mode: mnist/models/classifier.py:14,19) ------------------
// mask[row][inner][c] = Bernoulli(1 - p) / (1 - p); channelwise (Dropout2d): one draw per (row, c), repeated over `inner`.
// Same Philox stream convention as build_mask (key = seed + rng[2], offset rng[0], advanced by the last block).
__global__ void dropout_mask_kernel(long long rows, int inner, int C, float p, int channelwise, unsigned long long seed,
                                    unsigned long long* __restrict__ rng, float* __restrict__ mask) {
  pdl_enter();
  const unsigned long long offset = rng != nullptr ? rng[0] : 0ull;
  if (rng != nullptr) seed += rng[2];
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const float keep = 1.f / (1.f - p);
  const long long draws = channelwise ? rows * C : rows * inner * C;       // one 32-bit draw each, four per Philox call
  const long long calls = (draws + 3) / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < calls; i += (long long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32) | 0x80000000u, (uint32_t)offset, (uint32_t)(offset >> 32)), key);
    const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long d = i * 4 + j;
      if (d >= draws) break;
      // uniform in [0, 1): keep when u >= p (torch's bernoulli_(1 - p) keeps with probability 1 - p)
      const float u = (float)(rv[j] >> 8) * (1.f / 16777216.f);
      const float v = u >= p ? keep : 0.f;
      if (channelwise) {
        const long long row = d / C;
        const int c = (int)(d - row * C);
        float* dst = mask + row * inner * C + c;
        for (int k = 0; k < inner; ++k) dst[(long long)k * C] = v;
      } else {
        mask[d] = v;
      }
    }
  }
  if (rng != nullptr) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long ticket = atomicAdd(rng + 1, 1ull);
      if (ticket == (unsigned long long)gridDim.x - 1ull) {
        rng[1] = 0ull;
        rng[0] = offset + 1ull;
        __threadfence();
      }
    }
  }
}
void dropout_mask(long long rows, int inner, int C, float p, int channelwise, unsigned long long seed, unsigned long long* rng,
                  float* mask, cudaStream_t s) {
  PCG_PROFILE("dropout_mask", s);
  PCG_REQUIRE(rows >= 1 && inner >= 1 && C >= 1 && p >= 0.f && p < 1.f, "dropout geometry / probability");
  const long long draws = channelwise ? rows * C : rows * inner * C;
  launch_k(dropout_mask_kernel, dim3(blocks_for((draws + 3) / 4)), dim3(256), 0, s, rows, inner, C, p, channelwise, seed, rng, mask);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

void build_mask(int B, int C, int H, int W, int patch, int k_sel, int num_classes, unsigned long long seed,
                unsigned long long* rng, float* mask, long long* target, cudaStream_t s) {
  PCG_PROFILE("build_mask", s);
  PCG_REQUIRE(B >= 1 && C >= 1 && patch >= 1 && H >= patch && W >= patch, "mask geometry");
  const int nph = H / patch, npw = W / patch;
  PCG_REQUIRE(nph * npw <= 64, "at most 64 patches per image (one 64-bit word per sample)");
  PCG_REQUIRE(target == nullptr || num_classes >= 1, "num_classes");
  PCG_REQUIRE((reinterpret_cast<uintptr_t>(mask) & 15) == 0, "mask must be 16-byte aligned");
  launch_k(build_mask_kernel, dim3((B + MASK_SPB - 1) / MASK_SPB), dim3(256), 0, s, B, C, H, W, nph, npw, k_sel, num_classes,
           seed, rng, mask, target);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

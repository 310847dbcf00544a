// One C call (two launches) per half of a FiLM residual block of the tabular generator
// (conditional_counteRGAN/house_sales_kc_usa/models/generator.py:19-35: Linear -> BatchNorm1d(train) -> FiLM -> ReLU /
// residual add), forward and backward.
//
// As primitive operators a half block is five dependent launches forward (linear, BN statistics, finalize, apply, FiLM) and
// six backward; at [4096 x 32] each is a few microseconds of pure latency, and the five blocks of the generator are the
// critical path of the whole iteration.  The batch statistics force one grid-wide dependency per half block,
// so two launches is the minimum without a grid barrier:
//   forward A   u = x W^T + bias (written);  per-CTA column sums of u and u^2
//   forward B   every CTA adds the partials (fixed order, double) -> mean / rstd / scale / shift (CTA 0 stores them and
//               updates the running buffers);  n = BN(u);  f = fg * n + fb;  out = relu ? max(f, 0) : res + f
//   backward A  dfg (+)= d_f * n;  dfb (+)= d_f;  per-CTA column sums of dn = d_f * fg and dn * xhat
//   backward B  totals -> dbeta, dgamma;  du = gamma * rstd * (dn - dbeta / M - xhat * dgamma / M) (written: the weight
//               gradient consumes it);  dx = du W (+ add_src) (* [act_ref > 0])
// A first version ran each half block as ONE launch of an 8-CTA cluster with the reduction in distributed shared memory: 57 us
// per launch (64 warps for the whole batch, every load latency exposed) against ~25 us for the operators it replaced.
// Same arithmetic contract as those operators (pcg_conv_fprop 1x1, pcg_bn_train_fwd / _bwd, pcg_film_fwd / _bwd,
// pcg_conv_dgrad); tests/test_film_layer_gpu.py checks one against the other and against float64 torch.
#include "film_layer.cuh"

#include "elementwise.cuh"

namespace pcg {

constexpr int FL_THREADS = 256;
constexpr int FL_TILE = 32;          // rows per tile (a CTA walks its tiles)

bool film_layer_supported(long long M, int H) { return (H == 32 || H == 64) && M >= 1; }

static int fl_grid(long long M) {
  const long long tiles = (M + FL_TILE - 1) / FL_TILE;
  return (int)(tiles < STAT_PARTS ? tiles : STAT_PARTS);
}

// (s, q) of this thread's column c over the CTA's row lanes -> part[blockIdx.x][2][H]
template <int H>
__device__ __forceinline__ void cta_colsum2(float s, float q, float* flat, float* __restrict__ part, int c, int lane_row) {
  constexpr int LANES = FL_THREADS / H;
  flat[(lane_row * 2 + 0) * H + c] = s;
  flat[(lane_row * 2 + 1) * H + c] = q;
  __syncthreads();
  if (lane_row == 0) {
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int l = 0; l < LANES; ++l) { sa += flat[(l * 2 + 0) * H + c]; sb += flat[(l * 2 + 1) * H + c]; }
    part[((size_t)blockIdx.x * 2 + 0) * H + c] = sa;
    part[((size_t)blockIdx.x * 2 + 1) * H + c] = sb;
  }
}

// totals of part[nparts][2][H] for column c, identical in every thread of that column (fixed order)
template <int H>
__device__ __forceinline__ void total_colsum2(const float* __restrict__ part, int nparts, float* flat, int c, int lane_row,
                                              double& ts, double& tq) {
  constexpr int LANES = FL_THREADS / H;
  float sa = 0.f, sb = 0.f;
  for (int p = lane_row; p < nparts; p += LANES) {
    sa += part[((size_t)p * 2 + 0) * H + c];
    sb += part[((size_t)p * 2 + 1) * H + c];
  }
  flat[(lane_row * 2 + 0) * H + c] = sa;
  flat[(lane_row * 2 + 1) * H + c] = sb;
  __syncthreads();
  ts = 0.0; tq = 0.0;
#pragma unroll
  for (int l = 0; l < LANES; ++l) { ts += (double)flat[(l * 2 + 0) * H + c]; tq += (double)flat[(l * 2 + 1) * H + c]; }
}

template <int H>
__global__ void __launch_bounds__(FL_THREADS)
film_fwd_a_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, long long M,
                  float* __restrict__ u_o, float* __restrict__ part) {
  pdl_enter();
  constexpr int LANES = FL_THREADS / H, RPT = FL_TILE / LANES;      // rows per thread and tile: 4 (H = 32) or 8 (H = 64)
  __shared__ __align__(16) float sx[FL_TILE * H];
  __shared__ float flat[2 * FL_THREADS];
  const int c = threadIdx.x % H, lane_row = threadIdx.x / H;
  float w[H];
#pragma unroll
  for (int k = 0; k < H; k += 4) {
    const float4 t = *reinterpret_cast<const float4*>(W + (size_t)c * H + k);
    w[k] = t.x; w[k + 1] = t.y; w[k + 2] = t.z; w[k + 3] = t.w;
  }
  const float bc = bias[c];
  float s = 0.f, q = 0.f;
  for (long long t0 = (long long)blockIdx.x * FL_TILE; t0 < M; t0 += (long long)gridDim.x * FL_TILE) {
    const int nrows = M - t0 < FL_TILE ? (int)(M - t0) : FL_TILE;
    __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(x + t0 * H);
      float4* dst = reinterpret_cast<float4*>(sx);
      for (int i = threadIdx.x; i < nrows * (H / 4); i += FL_THREADS) dst[i] = src[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int rl = lane_row + k * LANES;
      if (rl < nrows) {
        float acc = bc;
        const float4* xr = reinterpret_cast<const float4*>(sx + rl * H);       // a warp reads one row: broadcast
#pragma unroll
        for (int j = 0; j < H / 4; ++j) {
          const float4 xv = xr[j];
          acc = fmaf(xv.x, w[4 * j], acc); acc = fmaf(xv.y, w[4 * j + 1], acc);
          acc = fmaf(xv.z, w[4 * j + 2], acc); acc = fmaf(xv.w, w[4 * j + 3], acc);
        }
        u_o[(size_t)(t0 + rl) * H + c] = acc;
        s += acc;
        q = fmaf(acc, acc, q);
      }
    }
  }
  cta_colsum2<H>(s, q, flat, part, c, lane_row);
}

template <int H>
__global__ void __launch_bounds__(FL_THREADS)
film_fwd_b_kernel(const float* __restrict__ u_i, const float* __restrict__ part, int nparts, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, float momentum, float* running_mean, float* running_var,
                  long long* nbt, float* mean_o, float* rstd_o, float* scale_o, float* shift_o, const float* __restrict__ fg,
                  const float* __restrict__ fb, const float* __restrict__ res, int relu, long long M, float* __restrict__ n_o,
                  float* __restrict__ out) {
  pdl_enter();
  constexpr int LANES = FL_THREADS / H, RPT = FL_TILE / LANES;
  __shared__ float flat[2 * FL_THREADS];
  const int c = threadIdx.x % H, lane_row = threadIdx.x / H;
  double ts, tq;
  total_colsum2<H>(part, nparts, flat, c, lane_row, ts, tq);
  const double mean = ts / (double)M;
  double var = tq / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float a = gamma[c] * rstd;
  const float b = beta[c] - (float)mean * a;
  if (blockIdx.x == 0 && lane_row == 0) {
    mean_o[c] = (float)mean;
    rstd_o[c] = rstd;
    scale_o[c] = a;
    shift_o[c] = b;
    if (running_mean != nullptr) {
      const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
    if (c == 0 && nbt != nullptr) *nbt += 1;
  }
  for (long long t0 = (long long)blockIdx.x * FL_TILE; t0 < M; t0 += (long long)gridDim.x * FL_TILE) {
    float uv[RPT], gv[RPT], bv[RPT], rv[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {                       // all loads of the tile first
      const long long r = t0 + lane_row + k * LANES;
      const bool ok = r < M;
      const size_t i = (size_t)r * H + c;
      uv[k] = ok ? u_i[i] : 0.f;
      gv[k] = ok ? fg[i] : 0.f;
      bv[k] = ok ? fb[i] : 0.f;
      rv[k] = (ok && !relu) ? res[i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const long long r = t0 + lane_row + k * LANES;
      if (r < M) {
        const size_t i = (size_t)r * H + c;
        const float n = fmaf(uv[k], a, b);
        const float f = fmaf(gv[k], n, bv[k]);
        n_o[i] = n;
        out[i] = relu ? fmaxf(f, 0.f) : rv[k] + f;
      }
    }
  }
}

// B of half block k and A of half block k + 1 in one launch: the BatchNorm + FiLM + activation of a row tile feeds the
// next Linear directly (both are row-local); only the batch statistics separate launches.  A chain of n half blocks is
// n + 1 launches (A, BA x (n - 1), B) instead of 2 n.
template <int H>
__global__ void __launch_bounds__(FL_THREADS)
film_fwd_ba_kernel(const float* __restrict__ u_i, const float* __restrict__ part, int nparts, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, float momentum, float* running_mean, float* running_var,
                   long long* nbt, float* mean_o, float* rstd_o, float* scale_o, float* shift_o, const float* __restrict__ fg,
                   const float* __restrict__ fb, const float* __restrict__ res, int relu, long long M, float* __restrict__ n_o,
                   float* __restrict__ out, const float* __restrict__ Wn, const float* __restrict__ biasn,
                   float* __restrict__ un_o, float* __restrict__ partn) {
  pdl_enter();
  constexpr int LANES = FL_THREADS / H, RPT = FL_TILE / LANES;
  __shared__ __align__(16) float sx[FL_TILE * H];
  __shared__ float flat[2 * FL_THREADS];
  const int c = threadIdx.x % H, lane_row = threadIdx.x / H;
  double ts, tq;
  total_colsum2<H>(part, nparts, flat, c, lane_row, ts, tq);
  const double mean = ts / (double)M;
  double var = tq / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float a = gamma[c] * rstd;
  const float b = beta[c] - (float)mean * a;
  if (blockIdx.x == 0 && lane_row == 0) {
    mean_o[c] = (float)mean;
    rstd_o[c] = rstd;
    scale_o[c] = a;
    shift_o[c] = b;
    if (running_mean != nullptr) {
      const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
    if (c == 0 && nbt != nullptr) *nbt += 1;
  }
  float w[H];                          // row c of the NEXT layer's weight
#pragma unroll
  for (int k = 0; k < H; k += 4) {
    const float4 t = *reinterpret_cast<const float4*>(Wn + (size_t)c * H + k);
    w[k] = t.x; w[k + 1] = t.y; w[k + 2] = t.z; w[k + 3] = t.w;
  }
  const float bcn = biasn[c];
  float s = 0.f, q = 0.f;
  for (long long t0 = (long long)blockIdx.x * FL_TILE; t0 < M; t0 += (long long)gridDim.x * FL_TILE) {
    float uv[RPT], gv[RPT], bv[RPT], rv[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const long long r = t0 + lane_row + k * LANES;
      const bool ok = r < M;
      const size_t i = (size_t)r * H + c;
      uv[k] = ok ? u_i[i] : 0.f;
      gv[k] = ok ? fg[i] : 0.f;
      bv[k] = ok ? fb[i] : 0.f;
      rv[k] = (ok && !relu) ? res[i] : 0.f;
    }
    __syncthreads();                   // the previous tile's rows are no longer read
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int rl = lane_row + k * LANES;
      const long long r = t0 + rl;
      const float n = fmaf(uv[k], a, b);
      const float f = fmaf(gv[k], n, bv[k]);
      const float o = relu ? fmaxf(f, 0.f) : rv[k] + f;
      sx[rl * H + c] = o;
      if (r < M) {
        const size_t i = (size_t)r * H + c;
        n_o[i] = n;
        out[i] = o;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int rl = lane_row + k * LANES;
      if (t0 + rl < M) {
        float acc = bcn;
        const float4* xr = reinterpret_cast<const float4*>(sx + rl * H);
#pragma unroll
        for (int j = 0; j < H / 4; ++j) {
          const float4 xv = xr[j];
          acc = fmaf(xv.x, w[4 * j], acc); acc = fmaf(xv.y, w[4 * j + 1], acc);
          acc = fmaf(xv.z, w[4 * j + 2], acc); acc = fmaf(xv.w, w[4 * j + 3], acc);
        }
        un_o[(size_t)(t0 + rl) * H + c] = acc;
        s += acc;
        q = fmaf(acc, acc, q);
      }
    }
  }
  cta_colsum2<H>(s, q, flat, partn, c, lane_row);
}

template <int H>
__global__ void __launch_bounds__(FL_THREADS)
film_bwd_a_kernel(const float* __restrict__ d_f, const float* __restrict__ fg, const float* __restrict__ n_i,
                  const float* __restrict__ u_i, const float* __restrict__ mean, const float* __restrict__ rstd,
                  int accumulate, long long M, float* __restrict__ dfg, float* __restrict__ dfb, float* __restrict__ part) {
  pdl_enter();
  constexpr int LANES = FL_THREADS / H, RPT = FL_TILE / LANES;
  __shared__ float flat[2 * FL_THREADS];
  const int c = threadIdx.x % H, lane_row = threadIdx.x / H;
  const float mu = mean[c], rs = rstd[c];
  float s = 0.f, q = 0.f;
  for (long long t0 = (long long)blockIdx.x * FL_TILE; t0 < M; t0 += (long long)gridDim.x * FL_TILE) {
    float dv[RPT], gv[RPT], nv[RPT], uv[RPT], ag[RPT], ab[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const long long r = t0 + lane_row + k * LANES;
      const bool ok = r < M;
      const size_t i = (size_t)r * H + c;
      dv[k] = ok ? d_f[i] : 0.f;
      gv[k] = ok ? fg[i] : 0.f;
      nv[k] = ok ? n_i[i] : 0.f;
      uv[k] = ok ? u_i[i] : mu;
      ag[k] = (ok && accumulate) ? dfg[i] : 0.f;
      ab[k] = (ok && accumulate) ? dfb[i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const long long r = t0 + lane_row + k * LANES;
      if (r < M) {
        const size_t i = (size_t)r * H + c;
        dfg[i] = fmaf(dv[k], nv[k], ag[k]);
        dfb[i] = ab[k] + dv[k];
        const float dn = dv[k] * gv[k];
        s += dn;
        q = fmaf(dn, (uv[k] - mu) * rs, q);
      }
    }
  }
  cta_colsum2<H>(s, q, flat, part, c, lane_row);
}

template <int H>
__global__ void __launch_bounds__(FL_THREADS)
film_bwd_b_kernel(const float* __restrict__ d_f, const float* __restrict__ fg, const float* __restrict__ u_i,
                  const float* __restrict__ part, int nparts, const float* __restrict__ mean, const float* __restrict__ rstd,
                  const float* __restrict__ gamma, const float* __restrict__ W, const float* __restrict__ add_src,
                  const float* __restrict__ act_ref, long long M, float* __restrict__ du_o, float* __restrict__ dx,
                  float* dgamma, float* dbeta) {
  pdl_enter();
  constexpr int LANES = FL_THREADS / H, RPT = FL_TILE / LANES;
  __shared__ __align__(16) float sdu[FL_TILE * H];
  __shared__ float flat[2 * FL_THREADS];
  const int c = threadIdx.x % H, lane_row = threadIdx.x / H;
  double ts, tq;
  total_colsum2<H>(part, nparts, flat, c, lane_row, ts, tq);
  const float inv = 1.f / (float)M;
  const float mu = mean[c], rs = rstd[c];
  const float k1 = (float)ts * inv, k2 = (float)tq * inv, gr = gamma[c] * rs;
  if (blockIdx.x == 0 && lane_row == 0) {
    dbeta[c] = (float)ts;
    dgamma[c] = (float)tq;
  }
  float wc[H];                         // column c of W: dx[r][c] = sum_j du[r][j] * W[j][c]
#pragma unroll
  for (int j = 0; j < H; ++j) wc[j] = W[(size_t)j * H + c];
  for (long long t0 = (long long)blockIdx.x * FL_TILE; t0 < M; t0 += (long long)gridDim.x * FL_TILE) {
    float dv[RPT], gv[RPT], uv[RPT], av[RPT], rv[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const long long r = t0 + lane_row + k * LANES;
      const bool ok = r < M;
      const size_t i = (size_t)r * H + c;
      dv[k] = ok ? d_f[i] : 0.f;
      gv[k] = ok ? fg[i] : 0.f;
      uv[k] = ok ? u_i[i] : mu;
      av[k] = (ok && add_src) ? add_src[i] : 0.f;
      rv[k] = (ok && act_ref) ? act_ref[i] : 1.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int rl = lane_row + k * LANES;
      const float d = gr * (dv[k] * gv[k] - k1 - (uv[k] - mu) * rs * k2);
      sdu[rl * H + c] = d;
      if (t0 + rl < M) du_o[(size_t)(t0 + rl) * H + c] = d;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int rl = lane_row + k * LANES;
      if (t0 + rl < M) {
        const float4* dr = reinterpret_cast<const float4*>(sdu + rl * H);
        float acc = av[k];
#pragma unroll
        for (int j = 0; j < H / 4; ++j) {
          const float4 t = dr[j];
          acc = fmaf(t.x, wc[4 * j], acc); acc = fmaf(t.y, wc[4 * j + 1], acc);
          acc = fmaf(t.z, wc[4 * j + 2], acc); acc = fmaf(t.w, wc[4 * j + 3], acc);
        }
        dx[(size_t)(t0 + rl) * H + c] = rv[k] > 0.f ? acc : 0.f;
      }
    }
  }
}

void film_layer_fwd(const float* x, long long M, int H, const float* W, const float* bias, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var, long long* nbt,
                    float* mean, float* rstd, float* scale, float* shift, const float* fg, const float* fb, const float* res,
                    bool relu, float* u, float* n, float* out, float* part, cudaStream_t s) {
  PCG_PROFILE("film_layer", s);
  PCG_REQUIRE(film_layer_supported(M, H), "film_layer: H in {32, 64}");
  PCG_REQUIRE(relu || res != nullptr, "film_layer_fwd: the residual form needs res");
  PCG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "film_layer: 16-byte alignment");
  const int grid = fl_grid(M);
#define PCG_L(HH)                                                                                                         \
  launch_k(film_fwd_a_kernel<HH>, dim3(grid), dim3(FL_THREADS), 0, s, x, W, bias, M, u, part);                              \
  launch_k(film_fwd_b_kernel<HH>, dim3(grid), dim3(FL_THREADS), 0, s, u, part, grid, gamma, beta, eps, momentum,            \
           running_mean, running_var, nbt, mean, rstd, scale, shift, fg, fb, res, relu ? 1 : 0, M, n, out)
  if (H == 32) { PCG_L(32); } else { PCG_L(64); }
#undef PCG_L
  PCG_COUNT_LAUNCH();
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

void film_chain_fwd(const float* x, long long M, int H, int n, const FilmHalfFwd* h, float eps, float momentum, cudaStream_t s) {
  PCG_PROFILE("film_layer", s);
  PCG_REQUIRE(film_layer_supported(M, H) && n >= 1, "film_chain: H in {32, 64}, at least one half block");
  for (int k = 0; k < n; ++k) {
    PCG_REQUIRE(h[k].relu || h[k].res != nullptr, "film_chain: the residual form needs res");
    PCG_REQUIRE((reinterpret_cast<uintptr_t>(h[k].W) & 15) == 0, "film_chain: 16-byte aligned weights");
  }
  PCG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "film_chain: 16-byte aligned input");
  const int grid = fl_grid(M);
#define PCG_L(HH)                                                                                                         \
  launch_k(film_fwd_a_kernel<HH>, dim3(grid), dim3(FL_THREADS), 0, s, x, h[0].W, h[0].bias, M, h[0].u, h[0].part);         \
  for (int k = 0; k + 1 < n; ++k)                                                                                         \
    launch_k(film_fwd_ba_kernel<HH>, dim3(grid), dim3(FL_THREADS), 0, s, h[k].u, h[k].part, grid, h[k].gamma, h[k].beta,   \
             eps, momentum, h[k].running_mean, h[k].running_var, h[k].nbt, h[k].mean, h[k].rstd, h[k].scale, h[k].shift,   \
             h[k].fg, h[k].fb, h[k].res, h[k].relu, M, h[k].n, h[k].out, h[k + 1].W, h[k + 1].bias, h[k + 1].u,           \
             h[k + 1].part);                                                                                              \
  launch_k(film_fwd_b_kernel<HH>, dim3(grid), dim3(FL_THREADS), 0, s, h[n - 1].u, h[n - 1].part, grid, h[n - 1].gamma,     \
           h[n - 1].beta, eps, momentum, h[n - 1].running_mean, h[n - 1].running_var, h[n - 1].nbt, h[n - 1].mean,        \
           h[n - 1].rstd, h[n - 1].scale, h[n - 1].shift, h[n - 1].fg, h[n - 1].fb, h[n - 1].res, h[n - 1].relu, M,        \
           h[n - 1].n, h[n - 1].out)
  if (H == 32) { PCG_L(32); } else { PCG_L(64); }
#undef PCG_L
  for (int k = 0; k <= n; ++k) PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

void film_layer_bwd(const float* d_f, long long M, int H, const float* fg, const float* n, const float* u, const float* mean,
                    const float* rstd, const float* gamma, const float* W, const float* add_src, const float* act_ref,
                    bool accumulate, float* dfg, float* dfb, float* du, float* dx, float* dgamma, float* dbeta, float* part,
                    cudaStream_t s) {
  PCG_PROFILE("film_layer", s);
  PCG_REQUIRE(film_layer_supported(M, H), "film_layer: H in {32, 64}");
  const int grid = fl_grid(M);
#define PCG_L(HH)                                                                                                         \
  launch_k(film_bwd_a_kernel<HH>, dim3(grid), dim3(FL_THREADS), 0, s, d_f, fg, n, u, mean, rstd, accumulate ? 1 : 0, M, dfg, \
           dfb, part);                                                                                                    \
  launch_k(film_bwd_b_kernel<HH>, dim3(grid), dim3(FL_THREADS), 0, s, d_f, fg, u, part, grid, mean, rstd, gamma, W, add_src, \
           act_ref, M, du, dx, dgamma, dbeta)
  if (H == 32) { PCG_L(32); } else { PCG_L(64); }
#undef PCG_L
  PCG_COUNT_LAUNCH();
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

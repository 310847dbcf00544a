// Train-mode BatchNorm over [M][C] rows as ONE launch for small problems (the tabular generators: M = batch <= a few
// thousand rows, C <= 256 channels): a thread-block cluster of 8 CTAs splits the rows, reduces the per-channel sums
// through distributed shared memory in rank order (deterministic), and applies the normalisation to its rows, which it
// kept in registers.  Replaces the three-launch pipeline (partial sums over 592 blocks, one-block-per-channel finalize,
// apply) of elementwise.cu where that is pure latency (a 64 x 32 problem: three dependent launches of ~4 us against one).  Same arithmetic contract as bn_stats_partial / bn_finalize / bn_apply_act and
// bn_bwd_partial / bn_bwd_finalize / bn_bwd_apply (torch.nn.BatchNorm1d in training mode, nn/functional.py batch_norm:
// biased variance for the normalisation, unbiased for running_var, momentum update, num_batches_tracked += 1).
#include <cooperative_groups.h>

#include "elementwise.cuh"

namespace cg = cooperative_groups;

namespace pcg {

constexpr int BNC_CTAS = 8;          // portable cluster size
constexpr int BNC_THREADS = 512;
constexpr int BNC_MAX_C = 256;
constexpr int BNC_REGS = 32;         // elements a thread can keep in registers between the two passes
constexpr int BNC_SMALL = 4;         // ... and how many it is allowed to have for the cluster path to be chosen

bool bn_cluster_supported(long long M, int C) {
  if (C < 1 || C > BNC_MAX_C || (BNC_THREADS % C) != 0) return false;      // C in {1, 2, 4, ..., 256}
  const long long rows_per_cta = (M + BNC_CTAS - 1) / BNC_CTAS;
  const int lanes = BNC_THREADS / C;                                        // row lanes per CTA
  // Taken only while a thread holds <= 4 elements (M*C <= 16 K: the moons generator, 64 x 32): at 4096 x 32 the eight
  // CTAs of the cluster are slower than the three launches that spread over all SMs (KC step 1.06 -> 1.14 ms, measured).
  return M >= 1 && (rows_per_cta + lanes - 1) / lanes <= BNC_SMALL;
}

__global__ void __cluster_dims__(BNC_CTAS, 1, 1) __launch_bounds__(BNC_THREADS)
bn_cluster_fwd_kernel(const float* __restrict__ y, long long M, int C, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, float momentum, float* running_mean, float* running_var,
                      long long* nbt, float* mean_o, float* rstd_o, float* scale_o, float* shift_o, int act, float slope,
                      float* __restrict__ z) {
  pdl_enter();
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float flat[2 * BNC_THREADS];                      // [row lane][sum | sum of squares][C], lanes * C == 512
  __shared__ float cta_sum[2 * BNC_MAX_C];
  const int rank = (int)cluster.block_rank();
  const int lanes = BNC_THREADS / C, c = threadIdx.x % C, lane_row = threadIdx.x / C;
  const long long per = (M + BNC_CTAS - 1) / BNC_CTAS;
  const long long r0 = rank * per, r1 = r0 + per < M ? r0 + per : M;
  float v[BNC_REGS];
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int k = 0; k < BNC_REGS; ++k) {
    const long long r = r0 + lane_row + (long long)k * lanes;
    v[k] = r < r1 ? y[r * C + c] : 0.f;
    s += v[k];
    q = fmaf(v[k], v[k], q);
  }
  flat[(lane_row * 2 + 0) * C + c] = s;
  flat[(lane_row * 2 + 1) * C + c] = q;
  __syncthreads();
  if (lane_row == 0) {
    float sa = 0.f, sb = 0.f;
    for (int l = 0; l < lanes; ++l) { sa += flat[(l * 2 + 0) * C + c]; sb += flat[(l * 2 + 1) * C + c]; }
    cta_sum[c] = sa;
    cta_sum[BNC_MAX_C + c] = sb;
  }
  cluster.sync();
  double ts = 0.0, tq = 0.0;
  for (int r = 0; r < BNC_CTAS; ++r) {
    const float* remote = cluster.map_shared_rank(cta_sum, r);
    ts += (double)remote[c];
    tq += (double)remote[BNC_MAX_C + c];
  }
  const double mean = ts / (double)M;
  double var = tq / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float a = gamma[c] * rstd;
  const float b = beta[c] - (float)mean * a;
  if (rank == 0 && lane_row == 0) {
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
#pragma unroll
  for (int k = 0; k < BNC_REGS; ++k) {
    const long long r = r0 + lane_row + (long long)k * lanes;
    if (r < r1) {
      float o = fmaf(v[k], a, b);
      if (act == ACT_LRELU) o = lrelu(o, slope);
      else if (act == ACT_RELU) o = fmaxf(o, 0.f);
      z[r * C + c] = o;
    }
  }
  cluster.sync();                                              // keep cta_sum alive until every remote read is done
}

// g = gscale * dz * act'(scale*y + shift); dbeta = sum g, dgamma = sum g*xhat; dy = gamma*rstd*(g - dbeta/M - xhat*dgamma/M)
__global__ void __cluster_dims__(BNC_CTAS, 1, 1) __launch_bounds__(BNC_THREADS)
bn_cluster_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ y, long long M, int C,
                      const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ scale, const float* __restrict__ shift, float gscale, int act, float slope,
                      float* __restrict__ dy, float* dgamma, float* dbeta, float* dbias_prev) {
  pdl_enter();
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float flat[2 * BNC_THREADS];
  __shared__ float cta_sum[3 * BNC_MAX_C];
  const int rank = (int)cluster.block_rank();
  const int lanes = BNC_THREADS / C, c = threadIdx.x % C, lane_row = threadIdx.x / C;
  const long long per = (M + BNC_CTAS - 1) / BNC_CTAS;
  const long long r0 = rank * per, r1 = r0 + per < M ? r0 + per : M;
  const float mu = mean[c], rs = rstd[c], a = scale[c], b = shift[c];
  float g[BNC_REGS], xh[BNC_REGS];
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int k = 0; k < BNC_REGS; ++k) {
    const long long r = r0 + lane_row + (long long)k * lanes;
    g[k] = 0.f; xh[k] = 0.f;
    if (r < r1) {
      const float yv = y[r * C + c];
      float d = dz[r * C + c];
      if (act == ACT_LRELU) d = fmaf(yv, a, b) > 0.f ? d : d * slope;
      else if (act == ACT_RELU) d = fmaf(yv, a, b) > 0.f ? d : 0.f;
      g[k] = d * gscale;
      xh[k] = (yv - mu) * rs;
    }
    s += g[k];
    q = fmaf(g[k], xh[k], q);
  }
  flat[(lane_row * 2 + 0) * C + c] = s;
  flat[(lane_row * 2 + 1) * C + c] = q;
  __syncthreads();
  if (lane_row == 0) {
    float sa = 0.f, sb = 0.f;
    for (int l = 0; l < lanes; ++l) { sa += flat[(l * 2 + 0) * C + c]; sb += flat[(l * 2 + 1) * C + c]; }
    cta_sum[c] = sa;
    cta_sum[BNC_MAX_C + c] = sb;
  }
  cluster.sync();
  double ts = 0.0, tq = 0.0;
  for (int r = 0; r < BNC_CTAS; ++r) {
    const float* remote = cluster.map_shared_rank(cta_sum, r);
    ts += (double)remote[c];
    tq += (double)remote[BNC_MAX_C + c];
  }
  if (rank == 0 && lane_row == 0) {
    dbeta[c] = (float)ts;
    dgamma[c] = (float)tq;
  }
  const float c1 = (float)(ts / (double)M), c2 = (float)(tq / (double)M);
  const float gr = gamma[c] * rs;
  float sdy = 0.f;
#pragma unroll
  for (int k = 0; k < BNC_REGS; ++k) {
    const long long r = r0 + lane_row + (long long)k * lanes;
    if (r < r1) {
      const float o = gr * (g[k] - c1 - xh[k] * c2);
      dy[r * C + c] = o;
      sdy += o;
    }
  }
  if (dbias_prev != nullptr) {                                 // column sums of dy: the bias gradient of the layer in front
    __syncthreads();
    flat[lane_row * C + c] = sdy;
    __syncthreads();
    if (lane_row == 0) {
      float sa = 0.f;
      for (int l = 0; l < lanes; ++l) sa += flat[l * C + c];
      cta_sum[2 * BNC_MAX_C + c] = sa;
    }
    cluster.sync();
    if (rank == 0 && lane_row == 0) {
      double t = 0.0;
      for (int r = 0; r < BNC_CTAS; ++r) t += (double)cluster.map_shared_rank(cta_sum, r)[2 * BNC_MAX_C + c];
      dbias_prev[c] = (float)t;
    }
  }
  cluster.sync();
}

void bn_cluster_fwd(const float* y, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                    float* running_mean, float* running_var, long long* nbt, float* mean, float* rstd, float* scale,
                    float* shift, int act, float slope, float* z, cudaStream_t s) {
  PCG_PROFILE("bn_cluster", s);
  PCG_REQUIRE(bn_cluster_supported(M, C), "cluster BatchNorm: problem too large");
  launch_k(bn_cluster_fwd_kernel, dim3(BNC_CTAS), dim3(BNC_THREADS), 0, s, y, M, C, gamma, beta, eps, momentum, running_mean,
           running_var, nbt, mean, rstd, scale, shift, act, slope, z);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

void bn_cluster_bwd(const float* dz, const float* y, long long M, int C, const float* gamma, const float* mean,
                    const float* rstd, const float* scale, const float* shift, float gscale, int act, float slope, float* dy,
                    float* dgamma, float* dbeta, float* dbias_prev, cudaStream_t s) {
  PCG_PROFILE("bn_cluster", s);
  PCG_REQUIRE(bn_cluster_supported(M, C), "cluster BatchNorm: problem too large");
  launch_k(bn_cluster_bwd_kernel, dim3(BNC_CTAS), dim3(BNC_THREADS), 0, s, dz, y, M, C, gamma, mean, rstd, scale, shift, gscale, act,
           slope, dy, dgamma, dbeta, dbias_prev);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

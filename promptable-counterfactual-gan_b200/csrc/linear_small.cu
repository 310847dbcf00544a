// Small fully connected layers: out[M][N] = epilogue( in[M][K] * W[N][K]^T ), K, N <= 128 (fp32, CUDA cores).
//
// The tabular generators, critics and their backward passes are chains of Linear layers 2..128 wide over a few thousand
// rows (house_sales_kc_usa/models/generator.py:13-92, discriminator.py:9-20; moons/models/*.py).  On the generic
// implicit-GEMM kernel (64 x 64 x 16 tiles built for convolutions with K in the hundreds) a [4096 x 32] x [32 x 32] layer
// takes 8-12 us, and ~46 of them sit on the critical path of a KC iteration.  Here a block owns 64 rows and ALL outputs:
// the whole weight matrix (<= 64 KB) and the block's input rows live in shared memory (row stride K + 1: conflict free),
// a warp owns 8 rows, a lane owns the output columns lane, lane + 32, ..., so weights are read conflict-free, inputs as
// warp broadcasts, and a row of outputs leaves as one coalesced store.  Same epilogue contract as the generic kernel
// (GenEpilogue: bias, activation, add_src, derivative of an activation reference).  The data gradient of a Linear layer
// is the same call with the transposed weight the plans already keep ([K][N] = "[out][in]" of the backward product).
#include "linear_small.cuh"

namespace pcg {

constexpr int LS_ROWS = 64, LS_THREADS = 256, LS_MAXD = 128;

bool linear_small_supported(const ConvGeom& g, long long M) {
  return g.H == 1 && g.W == 1 && g.ksize == 1 && g.stride == 1 && g.pad == 0 && g.Cin >= 1 && g.Cin <= LS_MAXD &&
         g.Cout >= 1 && g.Cout <= LS_MAXD && M >= 1 && M <= 2048;     // measured: no gain over the generic kernel at 4096 rows
                                                                       // (both are 64 latency-bound blocks; KC 1.06 vs 1.12 ms)
}

// Ntot > N: the block computes the column tile [blockIdx.y * 128, +N) of a wider output (N = columns of this tile).
// perm_c > 0: output column j = tap * perm_c + c reads weight row c * perm_taps + tap (the data gradient of a full-window
// convolution, i.e. ConvTranspose2d on a 1x1 input: wd is [Cin][taps][Cout] while the NHWC output runs (tap, Cin)).
// Data gradient of a full-window convolution with a small output count (ConvTranspose2d(100, 512, 4, 1, 0) on a 1x1 input,
// mnist_dcgan.py:77): din[n][(tap, ci)] = sum_co dout[n][co] * wd[ci][tap][co] - a [N x Cout] x [Cout x taps*Cin] product.
bool full_window_dgrad_supported(const ConvGeom& g) {
  return g.stride == 1 && g.pad == 0 && g.ksize == g.H && g.ksize == g.W && g.ksize > 1 && g.Cout <= LS_MAXD && g.Cout >= 1;
}

template <int J>       // J = ceil(N / 32) output columns per lane
__global__ void __launch_bounds__(LS_THREADS)
linear_small_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out, long long M, int K,
                    int Ntot, int perm_c, int perm_taps, const float* __restrict__ bias, int act, float slope,
                    const float* __restrict__ add_src, const float* __restrict__ act_ref, int ref_act, float ref_slope) {
  pdl_enter();
  extern __shared__ float sm[];
  const int ld = K + 1;
  const int n0 = blockIdx.y * LS_MAXD;
  const int N = Ntot - n0 < LS_MAXD ? Ntot - n0 : LS_MAXD;
  float* sw = sm;                       // [N][K + 1]
  float* sx = sm + N * ld;              // [LS_ROWS][K + 1]
  const long long m0 = (long long)blockIdx.x * LS_ROWS;
  for (int i = threadIdx.x; i < N * K; i += LS_THREADS) {
    const int n = i / K, k = i - n * K;
    const int j = n0 + n;
    const int wrow = perm_c > 0 ? (j % perm_c) * perm_taps + j / perm_c : j;
    sw[n * ld + k] = w[(size_t)wrow * K + k];
  }
  const int rows = (int)(M - m0 < LS_ROWS ? M - m0 : LS_ROWS);
  for (int i = threadIdx.x; i < rows * K; i += LS_THREADS) {
    const int r = i / K, k = i - r * K;
    sx[r * ld + k] = in[m0 * K + i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[8][J];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < J; ++j) acc[i][j] = 0.f;
  const float* xr = sx + warp * 8 * ld;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float wv[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int n = lane + 32 * j;
      wv[j] = n < N ? sw[n * ld + k] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xv = xr[i * ld + k];                  // same address for the whole warp: broadcast
#pragma unroll
      for (int j = 0; j < J; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp * 8 + i;
    if (r >= rows) break;
    const long long row = m0 + r;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int n = lane + 32 * j;
      if (n >= N) continue;
      float v = acc[i][j];
      const size_t o = (size_t)row * Ntot + n0 + n;
      if (bias) v += __ldg(bias + n0 + n);
      if (act == ACT_LRELU) v = v > 0.f ? v : v * slope;
      else if (act == ACT_RELU) v = fmaxf(v, 0.f);
      if (add_src) v += add_src[o];
      if (act_ref) {
        const float a = act_ref[o];
        if (ref_act == ACT_LRELU) v *= (a > 0.f ? 1.f : ref_slope);
        else if (ref_act == ACT_RELU) v *= (a > 0.f ? 1.f : 0.f);
      }
      out[o] = v;
    }
  }
}

void linear_small(const float* in, long long M, int K, int N, const float* w, const GenEpilogue<float>& e, float* out,
                  cudaStream_t s, int perm_c, int perm_taps) {
  PCG_PROFILE("linear_small", s);
  PCG_REQUIRE(K >= 1 && K <= LS_MAXD && N >= 1, "linear_small: K <= 128");
  const int ntile = N < LS_MAXD ? N : LS_MAXD;
  const size_t smem = (size_t)(ntile + LS_ROWS) * (K + 1) * sizeof(float);
  const dim3 grid((unsigned)((M + LS_ROWS - 1) / LS_ROWS), (unsigned)((N + LS_MAXD - 1) / LS_MAXD));
  const int J = (ntile + 31) / 32;
  static bool configured = false;
  if (!configured) {
    const int max_smem = (LS_MAXD + LS_ROWS) * (LS_MAXD + 1) * (int)sizeof(float);
    PCG_CHECK_CUDA(cudaFuncSetAttribute(linear_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(linear_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(linear_small_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(linear_small_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
#define PCG_LS(JJ)                                                                                                       \
  launch_k(linear_small_kernel<JJ>, grid, dim3(LS_THREADS), smem, s, in, w, out, M, K, N, perm_c, perm_taps, e.bias, e.act, \
           e.slope, e.add_src, e.act_ref, e.ref_act, e.ref_slope)
  if (J == 1) PCG_LS(1);
  else if (J == 2) PCG_LS(2);
  else if (J == 3) PCG_LS(3);
  else PCG_LS(4);
#undef PCG_LS
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

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

// ---------------------------------------------------------------------------------------------------------------
// Weight AND bias gradient of a small Linear layer in ONE call (two launches):  dw[N][K] = dy^T x,  db[N] = column sums of
// dy (K, N <= 128; the tabular networks: 21..128 wide, 4096 rows).  As primitive operators this is four launches (split-K
// product, its reduction, column sums, their reduction) per layer, 37 layers per KC iteration.
//   product  up to 128 CTAs take 32-row slices (one shared-memory tile each at 4096 rows); a thread owns up to 16 float4
//            outputs (n, k4 .. k4+3); partials [P][N * K4 * 4 + 128] go to scratch
//   sum      eight lanes per float4 output: lane l adds partials l, l + 8, ... (all loads in flight), then a fixed butterfly
// Deterministic; ~7 us as a dependency chain against 12-14 us for the four launches (tools/bench_wgrad_small.py).  A first
// version added the partials in the last CTA to arrive (one launch, a device-scope ticket): 20-50 us, the product -> fence ->
// ticket -> 64 sequential partial reads chain is pure latency.
constexpr int LW_TR = 32;            // rows per shared-memory tile
constexpr int LW_MAXP = 128;         // row slices (CTAs)

static int lw_ctas(long long M) {
  const long long p = (M + LW_TR - 1) / LW_TR;
  return (int)(p < 1 ? 1 : (p < LW_MAXP ? p : LW_MAXP));
}
bool linear_wgrad_small_supported(long long M, int K, int N) { return M >= 1 && K >= 1 && K <= LS_MAXD && N >= 1 && N <= LS_MAXD; }
long long linear_wgrad_small_scratch(long long M, int K, int N) {
  const int K4 = (K + 3) / 4;
  return 4 + (long long)lw_ctas(M) * ((long long)N * K4 * 4 + LS_MAXD);
}

template <int J>
__global__ void __launch_bounds__(LS_THREADS)
linear_wgrad_small_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long M, int K, int N,
                          float* __restrict__ scratch) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  const int K4 = (K + 3) >> 2, KP = K4 * 4;
  float* sx = sm;                       // [LW_TR][KP], zero padded columns
  float* sdy = sm + LW_TR * KP;         // [LW_TR][N]
  const int nout4 = N * K4;
  const long long per = ((M + gridDim.x - 1) / gridDim.x + LW_TR - 1) / LW_TR * LW_TR;
  const long long r0 = (long long)blockIdx.x * per, r1 = r0 + per < M ? r0 + per : M;
  float4 acc[J];
  int on[J], ok4[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int o = threadIdx.x + LS_THREADS * j;
    on[j] = o < nout4 ? o / K4 : -1;
    ok4[j] = o < nout4 ? o - (o / K4) * K4 : 0;
  }
  float bsum = 0.f;
  for (long long t0 = r0; t0 < r1; t0 += LW_TR) {
    const int nrows = r1 - t0 < LW_TR ? (int)(r1 - t0) : LW_TR;
    __syncthreads();
    for (int i = threadIdx.x; i < LW_TR * KP; i += LS_THREADS) {
      const int r = i / KP, k = i - r * KP;
      sx[i] = (r < nrows && k < K) ? x[(t0 + r) * K + k] : 0.f;
    }
    for (int i = threadIdx.x; i < LW_TR * N; i += LS_THREADS) {
      const int r = i / N;
      sdy[i] = r < nrows ? dy[(t0 + r) * N + (i - r * N)] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < LW_TR; ++r) {
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (on[j] >= 0) {
          const float d = sdy[r * N + on[j]];
          const float4 xv = *reinterpret_cast<const float4*>(sx + r * KP + ok4[j] * 4);
          acc[j].x = fmaf(d, xv.x, acc[j].x); acc[j].y = fmaf(d, xv.y, acc[j].y);
          acc[j].z = fmaf(d, xv.z, acc[j].z); acc[j].w = fmaf(d, xv.w, acc[j].w);
        }
      }
      if (threadIdx.x < N) bsum += sdy[r * N + threadIdx.x];
    }
  }
  const size_t stride = (size_t)nout4 * 4 + LS_MAXD;
  float* mine = scratch + 4 + (size_t)blockIdx.x * stride;
#pragma unroll
  for (int j = 0; j < J; ++j)
    if (on[j] >= 0) *reinterpret_cast<float4*>(mine + (size_t)(threadIdx.x + LS_THREADS * j) * 4) = acc[j];
  if (threadIdx.x < LS_MAXD) mine[(size_t)nout4 * 4 + threadIdx.x] = threadIdx.x < N ? bsum : 0.f;
}

// eight lanes per float4 of dw (then per float4 of db): lane l adds partials l, l + 8, ... in that order, a fixed butterfly
// over the eight lanes finishes
__global__ void __launch_bounds__(LS_THREADS)
linear_wgrad_small_sum_kernel(const float* __restrict__ scratch, int P, int K, int N, float* __restrict__ dw,
                              float* __restrict__ db) {
  pdl_enter();
  const int K4 = (K + 3) >> 2, nout4 = N * K4, nb4 = (N + 3) >> 2;
  const size_t stride = (size_t)nout4 * 4 + LS_MAXD;
  const int o = (blockIdx.x * LS_THREADS + threadIdx.x) >> 3, lane = threadIdx.x & 7;
  const bool live = o < nout4 + nb4;
  float4 v[LW_MAXP / 8];
#pragma unroll
  for (int u = 0; u < LW_MAXP / 8; ++u) {
    const int p = lane + 8 * u;
    v[u] = (live && p < P) ? *reinterpret_cast<const float4*>(scratch + 4 + (size_t)o * 4 + (size_t)p * stride)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 t = v[0];
#pragma unroll
  for (int u = 1; u < LW_MAXP / 8; ++u) { t.x += v[u].x; t.y += v[u].y; t.z += v[u].z; t.w += v[u].w; }
#pragma unroll
  for (int off = 4; off >= 1; off >>= 1) {
    t.x += __shfl_xor_sync(0xffffffffu, t.x, off); t.y += __shfl_xor_sync(0xffffffffu, t.y, off);
    t.z += __shfl_xor_sync(0xffffffffu, t.z, off); t.w += __shfl_xor_sync(0xffffffffu, t.w, off);
  }
  if (!live || lane != 0) return;
  const float tv[4] = {t.x, t.y, t.z, t.w};
  if (o < nout4) {
    const int n = o / K4, k4 = o - n * K4;
    for (int e = 0; e < 4 && k4 * 4 + e < K; ++e) dw[(size_t)n * K + k4 * 4 + e] = tv[e];
  } else if (db != nullptr) {
    const int n0 = (o - nout4) * 4;
    for (int e = 0; e < 4 && n0 + e < N; ++e) db[n0 + e] = tv[e];
  }
}

void linear_wgrad_small(const float* x, const float* dy, long long M, int K, int N, float* scratch, float* dw, float* db,
                        cudaStream_t s) {
  PCG_PROFILE("wgrad_small", s);
  PCG_REQUIRE(linear_wgrad_small_supported(M, K, N), "linear_wgrad_small: K, N <= 128");
  PCG_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "linear_wgrad_small: 16-byte aligned scratch");
  const int K4 = (K + 3) / 4, nout4 = N * K4;
  const int J = (nout4 + LS_THREADS - 1) / LS_THREADS;
  const size_t smem = (size_t)LW_TR * (K4 * 4 + N) * sizeof(float);
  const dim3 grid(lw_ctas(M));
#define PCG_LW(JJ) launch_k(linear_wgrad_small_kernel<JJ>, grid, dim3(LS_THREADS), smem, s, x, dy, M, K, N, scratch)
  if (J <= 1) PCG_LW(1);
  else if (J <= 2) PCG_LW(2);
  else if (J <= 4) PCG_LW(4);
  else if (J <= 8) PCG_LW(8);
  else PCG_LW(16);
#undef PCG_LW
  PCG_COUNT_LAUNCH();
  const int outs = nout4 + (N + 3) / 4;
  launch_k(linear_wgrad_small_sum_kernel, dim3((outs * 8 + LS_THREADS - 1) / LS_THREADS), dim3(LS_THREADS), 0, s, scratch,
           (int)grid.x, K, N, dw, db);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

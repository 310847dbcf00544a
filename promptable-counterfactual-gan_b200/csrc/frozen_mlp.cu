// The frozen classifier of the tabular CounteRGAN step as ONE launch: forward of an MLP (Linear + LeakyReLU, the eval-mode
// BatchNorms folded into the Linear layers by the caller), cross-entropy against the target classes, and the backward
// pass down to the gradient with respect to the INPUT (house_sales_kc_usa/trainer.py:300-303: `F.cross_entropy(clf_model(
// x_cf), target)`; the classifier's parameters are frozen, only d loss / d x_cf is used).
//
// As primitive operators this is 5 + 1 + 5 dependent launches on [4096 x 256] activations - short GEMMs (K <= 256) that
// the generic kernel runs at a quarter of the fp32 rate because nothing hides their load latency - and the longest single
// stretch of the KC iteration's critical path (0.15 of 0.41 ms, tools/kc_critical_path.py).  Here a CTA owns 32 rows and
// keeps every activation of them in shared memory (transposed, [width][32]: four rows of a column are one 16-byte read),
// weights stream through a double-buffered 16 x C shared-memory panel (cp.async), a thread accumulates 4 rows x C/64
// columns.  Nothing but x, the weights and dx touches global memory.
#include "frozen_mlp.cuh"

#include <cuda_pipeline.h>

namespace pcg {

constexpr int FM_ROWS = 32;          // rows per CTA
constexpr int FM_THREADS = 512;      // 16 warps: 8 row groups of 4 rows x 2 column halves (4 warps per scheduler hide the
                                     // shared-memory latency that 2 could not: 160 -> see DESIGN.md)
constexpr int FM_QC = 16;            // reduction chunk staged per panel
constexpr int FM_MAXW = 256;         // widest hidden layer
constexpr int FM_PLD = FM_MAXW + 4;  // panel row length (floats)

struct FmArgs {
  int L;                     // number of Linear layers
  int dims[FM_MAX_LAYERS + 1];
  const float* W[FM_MAX_LAYERS];    // [out][in]
  const float* WT[FM_MAX_LAYERS];   // [in][out]
  const float* b[FM_MAX_LAYERS];
  int act_off[FM_MAX_LAYERS];  // shared-memory offsets (floats) of the transposed activations a_0 .. a_{L-2}
  int x_off, dy_off[2], panel_off, w0_off, wl_off;
  float slope;
  // optional copies for a caller that also wants the weight gradients (a critic's own update): row-major [B][dims[j + 1]]
  // activations a_j and pre-activation gradients g_j of the hidden layers j = 0 .. L-2 (NULL: not stored)
  float* act_out[FM_MAX_LAYERS];
  float* grad_out[FM_MAX_LAYERS];
};

// outT[c][r] = sum_q inT[q][r] * Wq[q][c] for the CTA's 32 rows; Wq is a row-major [Q][C] matrix (forward: the transposed
// weight, q = input feature; data gradient: the weight itself, q = output feature).  C in {32, 64, 128, 256}.
// Epilogue: FWD: + bias[c], LeakyReLU;  !FWD: * LeakyReLU'(maskT[c][r]).  Panels of 16 rows of Wq are copied to shared memory
// with cp.async one chunk ahead of the FMAs.
template <bool FWD>
__device__ void fm_gemm(const float* __restrict__ Wq, int C, int Q, const float* __restrict__ inT,
                        const float* __restrict__ bias, const float* __restrict__ maskT, float slope,
                        float* __restrict__ outT, float* __restrict__ panel, float* __restrict__ copy, long long r0,
                        int nrows) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 7, half = threadIdx.x >> 8;
  // columns of this lane: half * C/2 + lane * CW + e, e < CW = C/64 (C = 32 / 16: one column, the second half - and for
  // 16 the upper lanes - idle)
  const int CW = C >= 64 ? C >> 6 : 1;
  const bool live = C >= 64 || (half == 0 && lane < C);
  const int cbase = (C >= 64 ? half * (C >> 1) : 0) + lane * CW;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  const int nchunk = (Q + FM_QC - 1) / FM_QC;
  const int c4n = C >> 2;                      // float4 per panel row
  auto stage = [&](int ch, int buf) {
    float* p = panel + buf * FM_QC * FM_PLD;
    const int q0 = ch * FM_QC;
    for (int i = threadIdx.x; i < FM_QC * c4n; i += FM_THREADS) {
      const int qq = i / c4n, c4 = i - qq * c4n;
      float* dst = p + qq * FM_PLD + c4 * 4;
      if (q0 + qq < Q) __pipeline_memcpy_async(dst, Wq + (size_t)(q0 + qq) * C + c4 * 4, 16);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __pipeline_commit();
  };
  stage(0, 0);
  for (int ch = 0; ch < nchunk; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunk) {
      stage(ch + 1, buf ^ 1);
      __pipeline_wait_prior(1);                // chunk ch has landed, chunk ch + 1 stays in flight
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
    const float* p = panel + buf * FM_QC * FM_PLD;
    const int q0 = ch * FM_QC;
    const int nq = Q - q0 < FM_QC ? Q - q0 : FM_QC;
#pragma unroll 8
    for (int qq = 0; qq < nq; ++qq) {
      if (!live) break;
      const float4 a = *reinterpret_cast<const float4*>(inT + (size_t)(q0 + qq) * FM_ROWS + warp * 4);   // broadcast
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float* pr = p + qq * FM_PLD + cbase;
      if (CW == 4) {
        const float4 w = *reinterpret_cast<const float4*>(pr);
        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[e][i] = fmaf(av[i], wv[e], acc[e][i]);
      } else if (CW == 2) {
        const float2 w = *reinterpret_cast<const float2*>(pr);
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[0][i] = fmaf(av[i], w.x, acc[0][i]); acc[1][i] = fmaf(av[i], w.y, acc[1][i]); }
      } else {
        const float w = pr[0];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[0][i] = fmaf(av[i], w, acc[0][i]);
      }
    }
    __syncthreads();                           // everyone is done with this buffer before it is refilled
  }
  // epilogue: column of accumulator j
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j >= CW || !live) break;
    const int c = cbase + j;
    float4 o;
    float* ov = reinterpret_cast<float*>(&o);
    if (FWD) {
      const float bb = bias[c];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v = acc[j][i] + bb;
        ov[i] = v > 0.f ? v : v * slope;
      }
    } else {
      const float4 m = *reinterpret_cast<const float4*>(maskT + (size_t)c * FM_ROWS + warp * 4);
      const float mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) ov[i] = mv[i] > 0.f ? acc[j][i] : acc[j][i] * slope;
    }
    *reinterpret_cast<float4*>(outT + (size_t)c * FM_ROWS + warp * 4) = o;
    if (copy != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (warp * 4 + i < nrows) copy[(size_t)(r0 + warp * 4 + i) * C + c] = ov[i];
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FM_THREADS)
frozen_mlp_ce_grad_kernel(FmArgs m, const float* __restrict__ x, const long long* __restrict__ target, int loss_kind, int B,
                          float wgt, float* __restrict__ logits_out, float* __restrict__ loss_part, float* __restrict__ dx) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  __shared__ float s_logit[FM_ROWS][FM_MAX_CLASSES];
  __shared__ float s_dlogit[FM_ROWS][FM_MAX_CLASSES];
  __shared__ float s_red[FM_ROWS];
  const int L = m.L, D0 = m.dims[0], NC = m.dims[L], H = m.dims[L - 1], H1 = m.dims[1];
  const long long r0 = (long long)blockIdx.x * FM_ROWS;
  const int nrows = B - r0 < FM_ROWS ? (int)(B - r0) : FM_ROWS;
  float* xT = sm + m.x_off;
  float* panel = sm + m.panel_off;
  float* w0s = sm + m.w0_off;          // W_0 [H1][D0] for the input gradient
  float* wls = sm + m.wl_off;          // W_{L-1} [NC][H] for the logits and their gradient
  // x tile, transposed [D0][32] (rows past the end are zero), and the two small weight matrices the hand-written ends use
  for (int i = threadIdx.x; i < FM_ROWS * D0; i += FM_THREADS) {
    const int r = i / D0, k = i - r * D0;
    xT[k * FM_ROWS + r] = r < nrows ? x[(r0 + r) * D0 + k] : 0.f;
  }
  for (int i = threadIdx.x; i < H1 * D0; i += FM_THREADS) w0s[i] = m.W[0][i];
  for (int i = threadIdx.x; i < NC * H; i += FM_THREADS) wls[i] = m.W[L - 1][i];
  __syncthreads();
  // ---- forward through the hidden layers
  const float* inT = xT;
  for (int j = 0; j + 1 < L; ++j) {
    float* outT = sm + m.act_off[j];
    fm_gemm<true>(m.WT[j], m.dims[j + 1], m.dims[j], inT, m.b[j], nullptr, m.slope, outT, panel, m.act_out[j], r0, nrows);
    inT = outT;
  }
  // ---- last layer (NC <= 8 classes), softmax cross-entropy, gradient of the logits
  for (int i = threadIdx.x; i < FM_ROWS * NC; i += FM_THREADS) {
    const int n = i / FM_ROWS, r = i - n * FM_ROWS;            // a warp: one class, 32 rows (conflict-free reads of inT)
    const float* w = wls + n * H;
    float a = m.b[L - 1][n];
    for (int k = 0; k < H; ++k) a = fmaf(inT[k * FM_ROWS + r], w[k], a);
    s_logit[r][n] = a;
    if (logits_out != nullptr && r < nrows) logits_out[(r0 + r) * NC + n] = a;
  }
  __syncthreads();
  if (threadIdx.x < FM_ROWS) {
    const int r = threadIdx.x;
    float term = 0.f;
    if (r < nrows && loss_kind == 1) {                 // mean of the outputs (a critic score): d / d out = wgt / B
      for (int n = 0; n < NC; ++n) { term += s_logit[r][n]; s_dlogit[r][n] = wgt / (float)B; }
    } else if (r < nrows) {
      float mx = s_logit[r][0];
      for (int n = 1; n < NC; ++n) mx = fmaxf(mx, s_logit[r][n]);
      float se = 0.f;
      for (int n = 0; n < NC; ++n) se += expf(s_logit[r][n] - mx);
      const float lse = mx + logf(se);
      const int t = (int)target[r0 + r];
      term = lse - s_logit[r][t];
      for (int n = 0; n < NC; ++n) s_dlogit[r][n] = wgt * (expf(s_logit[r][n] - lse) - (n == t ? 1.f : 0.f)) / (float)B;
    } else {
      for (int n = 0; n < NC; ++n) s_dlogit[r][n] = 0.f;
    }
    s_red[r] = term;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int r = 0; r < FM_ROWS; ++r) t += s_red[r];
    loss_part[blockIdx.x] = t;                       // the caller adds the parts and divides by B
  }
  // ---- backward: last layer by hand (reduction over NC), then the hidden layers, then the input layer
  float* dyT = sm + m.dy_off[0];
  for (int i = threadIdx.x; i < H * FM_ROWS; i += FM_THREADS) {
    const int k = i / FM_ROWS, r = i - k * FM_ROWS;
    float a = 0.f;
    for (int n = 0; n < NC; ++n) a = fmaf(s_dlogit[r][n], wls[n * H + k], a);
    const float act = inT[k * FM_ROWS + r];
    const float gv = act > 0.f ? a : a * m.slope;
    dyT[i] = gv;
    if (m.grad_out[L - 2] != nullptr && r < nrows) m.grad_out[L - 2][(size_t)(r0 + r) * H + k] = gv;
  }
  __syncthreads();
  int cur = 0;
  for (int j = L - 2; j >= 1; --j) {
    // g_{j-1} = (g_j W_j) * LeakyReLU'(a_{j-1}): reduction over layer j's outputs, W_j read row by row
    float* nxt = sm + m.dy_off[cur ^ 1];
    fm_gemm<false>(m.W[j], m.dims[j], m.dims[j + 1], sm + m.dy_off[cur], nullptr, sm + m.act_off[j - 1], m.slope, nxt, panel,
                   m.grad_out[j - 1], r0, nrows);
    cur ^= 1;
  }
  // input layer: dx[r][k] = sum_n g_0[n][r] * W_0[n][k]  (no activation in front of x); a warp: one k, 32 rows
  const float* d0 = sm + m.dy_off[cur];
  for (int i = threadIdx.x; i < FM_ROWS * D0; i += FM_THREADS) {
    const int k = i / FM_ROWS, r = i - k * FM_ROWS;
    float a = 0.f;
    for (int n = 0; n < H1; ++n) a = fmaf(d0[n * FM_ROWS + r], w0s[n * D0 + k], a);
    if (r < nrows) dx[(r0 + r) * D0 + k] = a;
  }
}

bool frozen_mlp_supported(int L, const int* dims) {
  if (L < 2 || L > FM_MAX_LAYERS) return false;
  if (dims[0] < 1 || dims[0] > 64 || dims[L] < 1 || dims[L] > FM_MAX_CLASSES) return false;
  for (int j = 1; j < L; ++j)
    if (dims[j] != 16 && dims[j] != 32 && dims[j] != 64 && dims[j] != 128 && dims[j] != 256) return false;
  return true;
}
int frozen_mlp_parts(int B) { return (B + FM_ROWS - 1) / FM_ROWS; }

void frozen_mlp_ce_grad(int L, const int* dims, const float* const* W, const float* const* WT, const float* const* b,
                        float slope, const float* x,
                        const long long* target, int loss_kind, int B, float wgt, float* logits, float* loss_part, float* dx,
                        cudaStream_t s, float* const* act_out, float* const* grad_out) {
  PCG_PROFILE("frozen_mlp", s);
  PCG_REQUIRE(frozen_mlp_supported(L, dims), "frozen_mlp: 2..6 layers, input <= 64, hidden widths in {16, 32, 64, 128, 256}, <= 8 classes");
  FmArgs m;
  m.L = L;
  m.slope = slope;
  int off = 0, widest = 32;
  for (int j = 0; j <= L; ++j) m.dims[j] = dims[j];
  for (int j = 0; j < FM_MAX_LAYERS; ++j) {
    m.act_out[j] = (act_out != nullptr && j + 1 < L) ? act_out[j] : nullptr;
    m.grad_out[j] = (grad_out != nullptr && j + 1 < L) ? grad_out[j] : nullptr;
  }
  for (int j = 0; j < L; ++j) {
    m.W[j] = W[j]; m.WT[j] = WT[j]; m.b[j] = b[j];
    PCG_REQUIRE((reinterpret_cast<uintptr_t>(W[j]) & 15) == 0 && (reinterpret_cast<uintptr_t>(WT[j]) & 15) == 0,
                "frozen_mlp: 16-byte aligned weights");
  }
  m.x_off = off; off += 64 * FM_ROWS;
  for (int j = 0; j + 1 < L; ++j) {
    m.act_off[j] = off;
    off += dims[j + 1] * FM_ROWS;
    widest = dims[j + 1] > widest ? dims[j + 1] : widest;
  }
  m.dy_off[0] = off; off += widest * FM_ROWS;
  m.dy_off[1] = off; off += widest * FM_ROWS;
  m.panel_off = off; off += 2 * FM_QC * FM_PLD;
  m.w0_off = off; off += (dims[1] * dims[0] + 3) / 4 * 4;
  m.wl_off = off; off += (dims[L] * dims[L - 1] + 3) / 4 * 4;
  const size_t smem = (size_t)off * sizeof(float);
  PCG_REQUIRE(smem <= 220 * 1024, "frozen_mlp: activations of 32 rows must fit in shared memory");
  PCG_CHECK_CUDA(cudaFuncSetAttribute(frozen_mlp_ce_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PCG_REQUIRE(loss_kind == 1 || target != nullptr, "frozen_mlp: the cross-entropy needs targets");
  launch_k(frozen_mlp_ce_grad_kernel, dim3(frozen_mlp_parts(B)), dim3(FM_THREADS), smem, s, m, x, target, loss_kind, B, wgt,
           logits, loss_part, dx);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

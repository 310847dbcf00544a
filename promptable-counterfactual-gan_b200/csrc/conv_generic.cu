// CUDA-core convolution kernels (fp32 math, NHWC storage in fp32 or bf16).
//
// Replaces, for the layers the tcgen05 kernels do not take: nn.Conv2d / nn.Linear forward and
// ConvolutionBackward0 / AddmmBackward0 of
//   conditional_counteRGAN/mnist/models/generator.py:39,50        (conv_in 3->64, conv_out 64->1)
//   conditional_counteRGAN/mnist/models/discriminator.py:14-24     (4 stride-2 convs, no bias)
//   conditional_counteRGAN/mnist/models/classifier.py:8-21         (3 convs + 2 linears, eval)
//
// Three kernel families:
//   gemm_conv_kernel<MODE>   64x64x16 smem-tiled implicit GEMM; MODE = fprop gather or dgrad gather
//   skinny_conv_kernel<MODE> <= 4 output channels (conv_out fprop, data-gradients wrt 1-3 channel
//                            images): one warp per pixel, HBM/L2-bound
//   wgrad kernels            split-K over pixels with per-slice partials reduced in fixed order
#include "conv_generic.cuh"

namespace pcg {

enum { MODE_FPROP = 0, MODE_DGRAD = 1 };

// Row (output pixel) -> base coordinates; shared by fprop/dgrad gathers.
struct RowCoord {
  int n, h, w;   // pixel of the OUTPUT tensor of this GEMM (fprop: (n,ho,wo); dgrad: (n,hi,wi))
};

template <int MODE>
struct Gather {
  // Source tensor spatial dims / channels and GEMM-K decomposition.
  int srcH, srcW, srcC;     // tensor being gathered from
  int ksize, stride, pad;
  // Returns element offset into the source tensor (pixel base, channel 0) or -1 if the tap is void.
  __device__ __forceinline__ long long pixel_offset(const RowCoord& rc, int tap) const {
    const int r = tap / ksize, s = tap - r * ksize;
    int sh, sw;
    if (MODE == MODE_FPROP) {
      sh = rc.h * stride - pad + r;
      sw = rc.w * stride - pad + s;
    } else {
      const int th = rc.h + pad - r, tw = rc.w + pad - s;
      if (th < 0 || tw < 0) return -1;
      if (stride > 1 && ((th % stride) != 0 || (tw % stride) != 0)) return -1;
      sh = th / stride;
      sw = tw / stride;
    }
    if (sh < 0 || sh >= srcH || sw < 0 || sw >= srcW) return -1;
    return (((long long)rc.n * srcH + sh) * srcW + sw) * srcC;
  }
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 f = *reinterpret_cast<const float4*>(p);
  v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}
template <>
__device__ __forceinline__ void load4<bf16>(const bf16* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

constexpr int GT = 64;     // tile M = tile N
constexpr int GK = 16;     // tile K
constexpr int GPAD = 4;

template <typename TOut>
struct EpiDev {
  const float* bias; int act; float slope;
  const TOut* add_src; const TOut* act_ref; int ref_act; float ref_slope;
};

template <typename TOut>
__device__ __forceinline__ float apply_epilogue(float v, const EpiDev<TOut>& e, long long row, int col, int ld) {
  if (e.bias) v += __ldg(e.bias + col);
  if (e.act == ACT_LRELU) v = v > 0.f ? v : v * e.slope;
  else if (e.act == ACT_RELU) v = fmaxf(v, 0.f);
  if (e.add_src) v += to_f(e.add_src[row * ld + col]);
  if (e.act_ref) {
    const float a = to_f(e.act_ref[row * ld + col]);
    if (e.ref_act == ACT_LRELU) v *= (a > 0.f ? 1.f : e.ref_slope);
    else if (e.ref_act == ACT_RELU) v *= (a > 0.f ? 1.f : 0.f);
  }
  return v;
}

// C[M][Nc] = A_gather[M][K] * B[Nc][K]^T ; K = taps * srcC with (tap, channel) ordering.
template <int MODE, typename TIn, typename TOut, bool VEC>
__global__ void __launch_bounds__(256)
gemm_conv_kernel(const TIn* __restrict__ src, const float* __restrict__ wgt, TOut* __restrict__ dst,
                 long long M, int Nc, int K, int outH, int outW, Gather<MODE> ga, EpiDev<TOut> epi) {
  pdl_enter();
  __shared__ float As[2][GK][GT + GPAD];
  __shared__ float Bs[2][GK][GT + GPAD];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * GT;
  const int n0 = blockIdx.y * GT;

  // loader mapping: 64 rows x 16 k, thread -> (row = tid/4, 4 consecutive k at (tid%4)*4)
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  RowCoord rc;
  const long long am = m0 + lrow;
  const bool arow_ok = am < M;
  {
    long long t = arow_ok ? am : 0;
    rc.w = (int)(t % outW); t /= outW;
    rc.h = (int)(t % outH); rc.n = (int)(t / outH);
  }
  const int bn = n0 + lrow;
  const bool brow_ok = bn < Nc;

  float ra[4], rb[4];
  auto load_tile = [&](int k0) {
    const int k = k0 + lk;
#pragma unroll
    for (int i = 0; i < 4; ++i) ra[i] = rb[i] = 0.f;
    if (VEC) {
      // srcC % 4 == 0 and K % 4 == 0: the 4 k's share a tap and are contiguous channels
      if (k < K) {
        if (arow_ok) {
          const int tap = k / ga.srcC, c = k - tap * ga.srcC;
          const long long off = ga.pixel_offset(rc, tap);
          if (off >= 0) load4<TIn>(src + off + c, ra);
        }
        if (brow_ok) load4<float>(wgt + (size_t)bn * K + k, rb);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = k + i;
        if (kk < K) {
          if (arow_ok) {
            const int tap = kk / ga.srcC, c = kk - tap * ga.srcC;
            const long long off = ga.pixel_offset(rc, tap);
            if (off >= 0) ra[i] = to_f(src[off + c]);
          }
          if (brow_ok) rb[i] = wgt[(size_t)bn * K + kk];
        }
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[buf][lk + i][lrow] = ra[i];
      Bs[buf][lk + i][lrow] = rb[i];
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + GK - 1) / GK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col < Nc) dst[row * Nc + col] = from_f<TOut>(apply_epilogue(acc[i][j], epi, row, col, Nc));
    }
  }
}

// <= 4 output channels: one warp per output pixel, lanes over the source channels.
template <int MODE, typename TIn, typename TOut, int NC>
__global__ void __launch_bounds__(256)
skinny_conv_kernel(const TIn* __restrict__ src, const float* __restrict__ wgt, TOut* __restrict__ dst,
                   long long M, int K, int outH, int outW, Gather<MODE> ga, EpiDev<TOut> epi) {
  pdl_enter();
  extern __shared__ float wsm[];   // [NC][K]
  for (int i = threadIdx.x; i < NC * K; i += blockDim.x) wsm[i] = wgt[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int taps = ga.ksize * ga.ksize;
  for (long long m = warp_global; m < M; m += nwarps) {
    RowCoord rc;
    long long t = m;
    rc.w = (int)(t % outW); t /= outW;
    rc.h = (int)(t % outH); rc.n = (int)(t / outH);
    float acc[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = 0.f;
    for (int tap = 0; tap < taps; ++tap) {
      const long long off = ga.pixel_offset(rc, tap);
      if (off < 0) continue;
      for (int c = lane; c < ga.srcC; c += 32) {
        const float a = to_f(src[off + c]);
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[j] = fmaf(a, wsm[j * K + tap * ga.srcC + c], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < NC; ++j) dst[m * NC + j] = from_f<TOut>(apply_epilogue(acc[j], epi, m, j, NC));
    }
  }
}


// ---- vectorised skinny kernel: LPP lanes span one pixel's source channels (8 channels = 16/32 B per lane),
// 32/LPP pixels per warp instruction, fully coalesced; shuffle-reduce over the LPP lanes.
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __bfloat1622float2(h[e]);
    v[2 * e] = f.x; v[2 * e + 1] = f.y;
  }
}

template <int MODE, typename TIn, typename TOut, int NC, int LPP>
__global__ void __launch_bounds__(256)
skinny_conv_v2_kernel(const TIn* __restrict__ src, const float* __restrict__ wgt, TOut* __restrict__ dst,
                      long long M, int K, int outH, int outW, Gather<MODE> ga, EpiDev<TOut> epi) {
  pdl_enter();
  extern __shared__ float wsm[];   // [NC][K]
  for (int i = threadIdx.x; i < NC * K; i += blockDim.x) wsm[i] = wgt[i];
  __syncthreads();
  constexpr int PPW = 32 / LPP;                 // pixels per warp pass
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPP, cl = lane % LPP;   // pixel slot in the warp, channel-group of this lane
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int taps = ga.ksize * ga.ksize;
  for (long long m0 = warp_global * PPW; m0 < M; m0 += nwarps * PPW) {
    const long long m = m0 + sub;
    float acc[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = 0.f;
    if (m < M) {
      RowCoord rc;
      long long t = m;
      rc.w = (int)(t % outW); t /= outW;
      rc.h = (int)(t % outH); rc.n = (int)(t / outH);
      // data gradient: only taps with (h + pad - r) % stride == 0 contribute -> visit exactly those
      const int r0 = (MODE == MODE_DGRAD) ? (rc.h + ga.pad) % ga.stride : 0;
      const int s0 = (MODE == MODE_DGRAD) ? (rc.w + ga.pad) % ga.stride : 0;
      const int step = (MODE == MODE_DGRAD) ? ga.stride : 1;
      for (int r = r0; r < ga.ksize; r += step)
      for (int sx = s0; sx < ga.ksize; sx += step) {
        const int tap = r * ga.ksize + sx;
        const long long off = ga.pixel_offset(rc, tap);
        if (off < 0) continue;
        float a[8];
        load8<TIn>(src + off + cl * 8, a);
        const float* wrow = wsm + tap * ga.srcC + cl * 8;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const float4 w0 = *reinterpret_cast<const float4*>(wrow + j * K);
          const float4 w1 = *reinterpret_cast<const float4*>(wrow + j * K + 4);
          acc[j] = fmaf(a[0], w0.x, acc[j]); acc[j] = fmaf(a[1], w0.y, acc[j]);
          acc[j] = fmaf(a[2], w0.z, acc[j]); acc[j] = fmaf(a[3], w0.w, acc[j]);
          acc[j] = fmaf(a[4], w1.x, acc[j]); acc[j] = fmaf(a[5], w1.y, acc[j]);
          acc[j] = fmaf(a[6], w1.z, acc[j]); acc[j] = fmaf(a[7], w1.w, acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (cl == 0 && m < M) {
#pragma unroll
      for (int j = 0; j < NC; ++j) dst[m * NC + j] = from_f<TOut>(apply_epilogue(acc[j], epi, m, j, NC));
    }
  }
}

template <int MODE, typename TIn, typename TOut, int NC>
static bool launch_skinny_v2(const TIn* src, const float* wgt, TOut* dst, long long M, int K, int outH, int outW,
                             const Gather<MODE>& ga, const EpiDev<TOut>& e, cudaStream_t stream) {
  const int lpp = ga.srcC / 8;
  if (ga.srcC % 8 != 0 || lpp > 32 || (lpp & (lpp - 1)) != 0 || (((uintptr_t)src) & 31) != 0 || (K % 4) != 0)
    return false;
  const size_t sm = (size_t)NC * K * sizeof(float);
  if (sm > 48 * 1024) return false;
  const long long warps_needed = (M * lpp + 31) / 32;
  long long blocks = (warps_needed + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
#define PCG_SK(L) launch_k(skinny_conv_v2_kernel<MODE, TIn, TOut, NC, L>, dim3((int)blocks), dim3(256), sm, stream, src, wgt, dst, M, K, outH, outW, ga, e)
  switch (lpp) {
    case 1: PCG_SK(1); break;
    case 2: PCG_SK(2); break;
    case 4: PCG_SK(4); break;
    case 8: PCG_SK(8); break;
    case 16: PCG_SK(16); break;
    default: PCG_SK(32); break;
  }
#undef PCG_SK
  return true;
}

template <typename TOut>
static EpiDev<TOut> to_dev(const GenEpilogue<TOut>& e) {
  EpiDev<TOut> d;
  d.bias = e.bias; d.act = e.act; d.slope = e.slope; d.add_src = e.add_src;
  d.act_ref = e.act_ref; d.ref_act = e.ref_act; d.ref_slope = e.ref_slope;
  return d;
}

template <int MODE, typename TIn, typename TOut>
static void launch_conv(const TIn* src, const float* wgt, TOut* dst, long long M, int Nc, int K, int outH,
                        int outW, const Gather<MODE>& ga, const GenEpilogue<TOut>& epi, cudaStream_t stream) {
  PCG_PROFILE("conv_generic", stream);
  const EpiDev<TOut> e = to_dev(epi);
  bool done = false;
  if (Nc <= 4) {
    switch (Nc) {
      case 1: done = launch_skinny_v2<MODE, TIn, TOut, 1>(src, wgt, dst, M, K, outH, outW, ga, e, stream); break;
      case 2: done = launch_skinny_v2<MODE, TIn, TOut, 2>(src, wgt, dst, M, K, outH, outW, ga, e, stream); break;
      case 3: done = launch_skinny_v2<MODE, TIn, TOut, 3>(src, wgt, dst, M, K, outH, outW, ga, e, stream); break;
      default: done = launch_skinny_v2<MODE, TIn, TOut, 4>(src, wgt, dst, M, K, outH, outW, ga, e, stream); break;
    }
  }
  if (done) {
  } else if (Nc <= 4) {
    const int blocks = (int)((M * 32 + 255) / 256 < (long long)sm_count() * 16 ? (M * 32 + 255) / 256
                                                                               : (long long)sm_count() * 16);
    const size_t sm = (size_t)Nc * K * sizeof(float);
    PCG_REQUIRE(sm <= 48 * 1024, "skinny conv weights must fit 48 KB of shared memory");
    switch (Nc) {
      case 1: launch_k(skinny_conv_kernel<MODE, TIn, TOut, 1>, dim3(blocks), dim3(256), sm, stream, src, wgt, dst, M, K, outH, outW, ga, e); break;
      case 2: launch_k(skinny_conv_kernel<MODE, TIn, TOut, 2>, dim3(blocks), dim3(256), sm, stream, src, wgt, dst, M, K, outH, outW, ga, e); break;
      case 3: launch_k(skinny_conv_kernel<MODE, TIn, TOut, 3>, dim3(blocks), dim3(256), sm, stream, src, wgt, dst, M, K, outH, outW, ga, e); break;
      default: launch_k(skinny_conv_kernel<MODE, TIn, TOut, 4>, dim3(blocks), dim3(256), sm, stream, src, wgt, dst, M, K, outH, outW, ga, e); break;
    }
  } else {
    dim3 grid((unsigned)((M + GT - 1) / GT), (unsigned)((Nc + GT - 1) / GT));
    const bool vec = (ga.srcC % 4 == 0) && (K % 4 == 0) && (((uintptr_t)src & 15) == 0) &&
                     (((uintptr_t)wgt & 15) == 0);
    if (vec) launch_k(gemm_conv_kernel<MODE, TIn, TOut, true>, dim3(grid), dim3(256), 0, stream, src, wgt, dst, M, Nc, K, outH, outW, ga, e);
    else launch_k(gemm_conv_kernel<MODE, TIn, TOut, false>, dim3(grid), dim3(256), 0, stream, src, wgt, dst, M, Nc, K, outH, outW, ga, e);
  }
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

template <typename TIn, typename TOut>
void conv_fprop_generic(const TIn* in, const ConvGeom& g, const float* wf, const GenEpilogue<TOut>& epi,
                        TOut* out, cudaStream_t stream) {
  Gather<MODE_FPROP> ga;
  ga.srcH = g.H; ga.srcW = g.W; ga.srcC = g.Cin; ga.ksize = g.ksize; ga.stride = g.stride; ga.pad = g.pad;
  launch_conv<MODE_FPROP>(in, wf, out, g.Mout(), g.Cout, g.K(), g.Ho(), g.Wo(), ga, epi, stream);
}

template <typename TIn, typename TOut>
void conv_dgrad_generic(const TIn* dout, const ConvGeom& g, const float* wd, const GenEpilogue<TOut>& epi,
                        TOut* din, cudaStream_t stream, int ch_select) {
  Gather<MODE_DGRAD> ga;
  ga.srcH = g.Ho(); ga.srcW = g.Wo(); ga.srcC = g.Cout; ga.ksize = g.ksize; ga.stride = g.stride; ga.pad = g.pad;
  const int K = g.ksize * g.ksize * g.Cout;
  if (ch_select >= 0) {
    // only input channel `ch_select` is wanted: a 1-column problem on that row of wd, compact [Min][1] output
    PCG_REQUIRE(ch_select < g.Cin && epi.add_src == nullptr && epi.act_ref == nullptr && epi.bias == nullptr,
                "channel-select dgrad takes no epilogue tensors");
    launch_conv<MODE_DGRAD>(dout, wd + (size_t)ch_select * K, din, g.Min(), 1, K, g.H, g.W, ga, epi, stream);
    return;
  }
  launch_conv<MODE_DGRAD>(dout, wd, din, g.Min(), g.Cin, K, g.H, g.W, ga, epi, stream);
}

// ------------------------------------------------------------------------------------------
// wgrad: part[z][co][k] = sum over pixel slice z of dy[p][co] * x[p@tap][ci],  k = (tap, ci)
// ------------------------------------------------------------------------------------------
template <typename TIn, typename TDy, bool VEC>
__global__ void __launch_bounds__(256)
wgrad_gemm_kernel(const TIn* __restrict__ x, const TDy* __restrict__ dy, float* __restrict__ part, long long P,
                  int Cout, int K, int outH, int outW, long long slice, Gather<MODE_FPROP> ga) {
  pdl_enter();
  __shared__ float As[2][GK][GT + GPAD];   // [pixel][co]
  __shared__ float Bs[2][GK][GT + GPAD];   // [pixel][k]
  const int tid = threadIdx.x;
  const int co0 = blockIdx.x * GT, k0 = blockIdx.y * GT;
  const long long p_begin = (long long)blockIdx.z * slice;
  const long long p_end = p_begin + slice < P ? p_begin + slice : P;

  const int lp = tid >> 4, lc = (tid & 15) * 4;   // 16 pixels x 64 columns, 4 consecutive columns each
  float ra[4], rb[4];
  auto load_tile = [&](long long pbase) {
    const long long p = pbase + lp;
#pragma unroll
    for (int i = 0; i < 4; ++i) ra[i] = rb[i] = 0.f;
    if (p < p_end) {
      if (VEC) {
        if (co0 + lc < Cout) load4<TDy>(dy + p * Cout + co0 + lc, ra);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (co0 + lc + i < Cout) ra[i] = to_f(dy[p * Cout + co0 + lc + i]);
      }
      RowCoord rc;
      long long t = p;
      rc.w = (int)(t % outW); t /= outW;
      rc.h = (int)(t % outH); rc.n = (int)(t / outH);
      const int k = k0 + lc;
      if (VEC) {
        if (k < K) {
          const int tap = k / ga.srcC, c = k - tap * ga.srcC;
          const long long off = ga.pixel_offset(rc, tap);
          if (off >= 0) load4<TIn>(x + off + c, rb);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kk = k + i;
          if (kk < K) {
            const int tap = kk / ga.srcC, c = kk - tap * ga.srcC;
            const long long off = ga.pixel_offset(rc, tap);
            if (off >= 0) rb[i] = to_f(x[off + c]);
          }
        }
      }
    }
  };
  auto store_tile = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lp][lc]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    *reinterpret_cast<float4*>(&Bs[buf][lp][lc]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const long long npix = p_end > p_begin ? p_end - p_begin : 0;
  const int nt = (int)((npix + GK - 1) / GK);
  if (nt > 0) {
    load_tile(p_begin);
    store_tile(0);
    __syncthreads();
  }
  for (int t = 0; t < nt; ++t) {
    const int buf = t & 1;
    if (t + 1 < nt) load_tile(p_begin + (long long)(t + 1) * GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < nt) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }
  float* my = part + (size_t)blockIdx.z * Cout * K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) my[(size_t)co * K + k] = acc[i][j];
    }
  }
}


// ---- skinny wgrad (Cout <= 4, Cin % 8 == 0): lanes over input channels, accumulators in registers,
// one partial row per block: part[block][co][k].
template <typename TIn, typename TDy, int NC, int LPP>
__global__ void __launch_bounds__(256)
wgrad_skinny_kernel(const TIn* __restrict__ x, const TDy* __restrict__ dy, float* __restrict__ part, long long P,
                    int K, int outH, int outW, Gather<MODE_FPROP> ga) {
  pdl_enter();
  constexpr int PPW = 32 / LPP;
  constexpr int MAXTAPS = 9;
  extern __shared__ float red[];    // [8 warps][NC][K]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPP, cl = lane % LPP;
  const int taps = ga.ksize * ga.ksize;
  float acc[NC][MAXTAPS][8];
#pragma unroll
  for (int j = 0; j < NC; ++j)
#pragma unroll
    for (int t = 0; t < MAXTAPS; ++t)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[j][t][e] = 0.f;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p0 = warp_global * PPW; p0 < P; p0 += nwarps * PPW) {
    const long long p = p0 + sub;
    if (p >= P) continue;
    float g[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) g[j] = to_f(dy[p * NC + j]);
    RowCoord rc;
    long long t = p;
    rc.w = (int)(t % outW); t /= outW;
    rc.h = (int)(t % outH); rc.n = (int)(t / outH);
#pragma unroll
    for (int tap = 0; tap < MAXTAPS; ++tap) {
      if (tap >= taps) break;
      const long long off = ga.pixel_offset(rc, tap);
      if (off < 0) continue;
      float a[8];
      load8<TIn>(x + off + cl * 8, a);
#pragma unroll
      for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[j][tap][e] = fmaf(g[j], a[e], acc[j][tap][e]);
    }
  }
  // reduce over the PPW pixel slots of the warp, then over the 8 warps of the block (fixed order)
#pragma unroll
  for (int j = 0; j < NC; ++j)
#pragma unroll
    for (int tap = 0; tap < MAXTAPS; ++tap)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = acc[j][tap][e];
#pragma unroll
        for (int o = LPP; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (sub == 0 && tap < taps) red[((size_t)warp * NC + j) * K + tap * ga.srcC + cl * 8 + e] = v;
      }
  __syncthreads();
  for (int i = threadIdx.x; i < NC * K; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[(size_t)w * NC * K + i];
    part[(size_t)blockIdx.x * NC * K + i] = s;
  }
}

// part[z][co][(tap,ci)] -> dw[co][ci][tap] (torch OIHW), fixed order over z.
// 32 outputs x 8 slice groups per block: a thread sums the slices z = g, g + 8, ... as two independent chains (the loads
// of one thread are all in flight: a Linear(32, 32) over 4096 rows has 128 slices of 1024 elements, which one thread per
// element summed as 128 dependent loads = 14 us), then the eight groups are added in fixed order through shared memory.
constexpr int WRG_THREADS = 256;
__global__ void __launch_bounds__(WRG_THREADS)
wgrad_reduce_generic_kernel(const float* __restrict__ part, int nz, int Cout, int Cin, int taps, int groups,
                            float* __restrict__ dw) {
  pdl_enter();
  __shared__ float sm[WRG_THREADS];
  const int K = taps * Cin, total = Cout * K;
  const int vec = WRG_THREADS / groups;                      // outputs per block
  const int v = threadIdx.x % vec, g = threadIdx.x / vec;
  const int idx = blockIdx.x * vec + v;
  float s0 = 0.f, s1 = 0.f;
  if (idx < total) {
    const float* p = part + idx;
    int z = g;
#pragma unroll 4
    for (; z + groups < nz; z += 2 * groups) {
      s0 += p[(size_t)z * total];
      s1 += p[(size_t)(z + groups) * total];
    }
    if (z < nz) s0 += p[(size_t)z * total];
  }
  sm[threadIdx.x] = s0 + s1;
  __syncthreads();
  if (g == 0 && idx < total) {
    float t = sm[v];
    for (int k = 1; k < groups; ++k) t += sm[k * vec + v];
    const int co = idx / K, kk = idx - co * K;
    const int tap = kk / Cin, ci = kk - tap * Cin;
    dw[((size_t)co * Cin + ci) * taps + tap] = t;
  }
}

// slice groups per block: enough to keep ~16 loads per thread in flight, never more groups than slices
static int wrg_groups(int nz) { return nz >= 64 ? 8 : nz >= 16 ? 4 : nz >= 4 ? 2 : 1; }
static void launch_wgrad_reduce(const float* part, int nz, int Cout, int Cin, int taps, float* dw, cudaStream_t stream) {
  const int groups = wrg_groups(nz);
  launch_k(wgrad_reduce_generic_kernel, dim3(cdiv((long long)Cout * Cin * taps, WRG_THREADS / groups)), dim3(WRG_THREADS), 0, stream,
           part, nz, Cout, Cin, taps, groups, dw);
}

void wgrad_reduce_generic(const float* part, int nz, int Cout, int Cin, int taps, float* dw, cudaStream_t stream) {
  PCG_PROFILE("wgrad_reduce", stream);
  launch_wgrad_reduce(part, nz, Cout, Cin, taps, dw, stream);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

static bool wgrad_is_skinny(const ConvGeom& g) {
  const int lpp = g.Cin / 8;
  return g.Cout == 1 && g.Cin % 8 == 0 && lpp <= 32 && (lpp & (lpp - 1)) == 0 && g.ksize <= 3 &&
         (size_t)8 * g.Cout * g.K() * sizeof(float) <= 48 * 1024;
}
constexpr int WG_SKINNY_BLOCKS = 592;

static int wgrad_slices(const ConvGeom& g) {
  if (wgrad_is_skinny(g)) return WG_SKINNY_BLOCKS;
  const long long P = g.Mout();
  const int tiles = cdiv(g.Cout, GT) * cdiv(g.K(), GT);
  long long want = (4LL * 148 + tiles - 1) / tiles;
  long long maxz = (P + 31) / 32;         // at least 32 pixels per slice (a 4096-row Linear layer has ONE output tile:
                                          // 16 slices of 256 rows left 132 SMs idle and cost 34 us per weight gradient)
  if (want > maxz) want = maxz;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}

size_t conv_wgrad_generic_scratch(const ConvGeom& g) {
  return (size_t)wgrad_slices(g) * g.Cout * g.K();
}

template <typename TIn, typename TDy>
void conv_wgrad_generic(const TIn* in, const TDy* dout, const ConvGeom& g, float* scratch, float* dw,
                        cudaStream_t stream) {
  PCG_PROFILE("wgrad_generic", stream);
  Gather<MODE_FPROP> ga;
  ga.srcH = g.H; ga.srcW = g.W; ga.srcC = g.Cin; ga.ksize = g.ksize; ga.stride = g.stride; ga.pad = g.pad;
  const long long P = g.Mout();
  const int nz = wgrad_slices(g);
  if (wgrad_is_skinny(g) && (((uintptr_t)in) & 31) == 0) {
    const size_t sm = (size_t)8 * g.Cout * g.K() * sizeof(float);
#define PCG_WS(L) launch_k(wgrad_skinny_kernel<TIn, TDy, 1, L>, dim3(WG_SKINNY_BLOCKS), dim3(256), sm, stream, in, dout, scratch, P, g.K(), g.Ho(), g.Wo(), ga)
    switch (g.Cin / 8) {
      case 1: PCG_WS(1); break;
      case 2: PCG_WS(2); break;
      case 4: PCG_WS(4); break;
      case 8: PCG_WS(8); break;
      case 16: PCG_WS(16); break;
      default: PCG_WS(32); break;
    }
#undef PCG_WS
    PCG_COUNT_LAUNCH();
    PCG_LAUNCH_CHECK();
    launch_wgrad_reduce(scratch, nz, g.Cout, g.Cin, g.ksize * g.ksize, dw, stream);
    PCG_COUNT_LAUNCH();
    PCG_LAUNCH_CHECK();
    return;
  }
  const long long slice = ((P + nz - 1) / nz + GK - 1) / GK * GK;
  dim3 grid(cdiv(g.Cout, GT), cdiv(g.K(), GT), nz);
  const bool vec = (g.Cin % 4 == 0) && (g.Cout % 4 == 0) && (((uintptr_t)in & 15) == 0) &&
                   (((uintptr_t)dout & 15) == 0);
  if (vec) launch_k(wgrad_gemm_kernel<TIn, TDy, true>, dim3(grid), dim3(256), 0, stream, in, dout, scratch, P, g.Cout, g.K(), g.Ho(), g.Wo(), slice, ga);
  else launch_k(wgrad_gemm_kernel<TIn, TDy, false>, dim3(grid), dim3(256), 0, stream, in, dout, scratch, P, g.Cout, g.K(), g.Ho(), g.Wo(), slice, ga);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
  launch_wgrad_reduce(scratch, nz, g.Cout, g.Cin, g.ksize * g.ksize, dw, stream);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
__global__ void pack_generic_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int perm_hw,
                                    float* __restrict__ wf, float* __restrict__ wd) {
  pdl_enter();
  const int total = Cout * Cin * taps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % taps;
    int ci = (i / taps) % Cin;
    const int co = i / (taps * Cin);
    if (perm_hw > 0) {           // torch flatten index c*HW + hw  ->  NHWC flatten index hw*C + c
      const int C = Cin / perm_hw;
      const int c = ci / perm_hw, hw = ci - c * perm_hw;
      ci = hw * C + c;
    }
    const float v = w[i];
    if (wf) wf[((size_t)co * taps + tap) * Cin + ci] = v;
    // perm_hw == -1: taps reversed - wd is then the FORWARD weight [Cin][taps][Cout] of the stride-1 convolution over the
    // zero-dilated output gradient that equals this layer's data gradient (pcg_dilate)
    if (wd) wd[((size_t)ci * taps + (perm_hw == -1 ? taps - 1 - tap : tap)) * Cout + co] = v;
  }
}

// Tiled transpose dst[c][r] = src[r][c] (32 x 32 tiles through shared memory, both sides coalesced).  taps > 1 with
// rev != 0: the column index is (ci, tap) and lands on row (ci, taps - 1 - tap) - the tap-reversed wd of perm_hw == -1.
__global__ void __launch_bounds__(256) transpose_tiled_kernel(const float* __restrict__ src, int R, int C, int taps, int rev,
                                                             float* __restrict__ dst) {
  pdl_enter();
  __shared__ float tile[32][33];
  src += (size_t)blockIdx.z * R * C;                 // batch of matrices (wf: one [Cin][taps] -> [taps][Cin] per Cout)
  dst += (size_t)blockIdx.z * R * C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    if (r0 + j < R && c0 + tx < C) tile[j][tx] = src[(size_t)(r0 + j) * C + c0 + tx];
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j;
    if (c < C && r0 + tx < R) {
      if (rev) c = c / taps * taps + (taps - 1 - c % taps);
      dst[(size_t)c * R + r0 + tx] = tile[tx][j];
    }
  }
}
void transpose_tiled(const float* src, int R, int C, float* dst, cudaStream_t stream, int taps, bool rev, int batch) {
  PCG_REQUIRE((R + 31) / 32 <= 65535 && batch >= 1 && batch <= 65535, "transpose: at most 2M rows, 65535 matrices");
  launch_k(transpose_tiled_kernel, dim3((C + 31) / 32, (R + 31) / 32, batch), dim3(256), 0, stream, src, R, C, taps,
           rev ? 1 : 0, dst);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

void pack_conv_weights_generic(const float* w, int Cout, int Cin, int ksize, int perm_hw, float* wf, float* wd,
                               cudaStream_t stream) {
  PCG_PROFILE("pack_weights", stream);
  if (wd != nullptr && perm_hw <= 0 && (long long)Cout * Cin * ksize * ksize >= (1 << 16)) {
    // wd [Cin][taps][Cout] is the transpose of w viewed as [Cout][Cin * taps]: the scatter below would write it 4 bytes
    // at a time across rows (8.4 M elements for the WGAN critic's Linear(8192, 1024))
    transpose_tiled(w, Cout, Cin * ksize * ksize, wd, stream, ksize * ksize, perm_hw == -1);
    wd = nullptr;
    if (wf != nullptr && perm_hw == 0 && ksize > 1 && Cout <= 65535) {
      // wf [Cout][taps][Cin]: per output channel the [Cin][taps] block of w transposed
      transpose_tiled(w, Cin, ksize * ksize, wf, stream, 1, false, Cout);
      wf = nullptr;
    }
    if (wf == nullptr) return;
  }
  const int total = Cout * Cin * ksize * ksize;
  int blocks = cdiv(total, 256);
  if (blocks > 1184) blocks = 1184;
  launch_k(pack_generic_kernel, dim3(blocks), dim3(256), 0, stream, w, Cout, Cin, ksize * ksize, perm_hw, wf, wd);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// explicit instantiations
#define INST(TI, TO)                                                                                      \
  template void conv_fprop_generic<TI, TO>(const TI*, const ConvGeom&, const float*, const GenEpilogue<TO>&, \
                                           TO*, cudaStream_t);                                            \
  template void conv_dgrad_generic<TI, TO>(const TI*, const ConvGeom&, const float*, const GenEpilogue<TO>&, \
                                           TO*, cudaStream_t, int);                                          \
  template void conv_wgrad_generic<TI, TO>(const TI*, const TO*, const ConvGeom&, float*, float*, cudaStream_t);
INST(float, float)
INST(bf16, bf16)
INST(float, bf16)
INST(bf16, float)
#undef INST

}  // namespace pcg

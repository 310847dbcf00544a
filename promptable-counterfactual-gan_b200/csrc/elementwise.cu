// HBM-bound kernels of the GAN step (see elementwise.cuh for the reference call sites).
#include "elementwise.cuh"

#include "cluster_reduce.cuh"

#include <initializer_list>

namespace pcg {

// ---------------------------------------------------------------------------------------------
// vector helpers: V (4 or 8) consecutive channels per thread.  V = 8 makes every bf16 access 16 bytes
// (fp32: 2 x 16), which together with two rows in flight per thread is what it takes to keep
// enough bytes outstanding for HBM3e (~5 MB at 6.5 TB/s x 800 ns).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 f = *reinterpret_cast<const float4*>(p);
  v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void ldv(const float* p, float (&v)[4]) { ld4(p, v); }
__device__ __forceinline__ void ldv(const bf16* p, float (&v)[4]) { ld4(p, v); }
__device__ __forceinline__ void stv(float* p, const float (&v)[4]) { st4(p, v); }
__device__ __forceinline__ void stv(bf16* p, const float (&v)[4]) { st4(p, v); }
__device__ __forceinline__ void ldv(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ldv(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
// evict-first ("streaming") variant for inputs that are read for the last time; only the 8 x bf16 vector has one
template <typename T, int V>
__device__ __forceinline__ void ldv_cs(const T* p, float (&v)[V]) { ldv(p, v); }
template <>
__device__ __forceinline__ void ldv_cs<bf16, 8>(const bf16* p, float (&v)[8]) {
  const uint4 u = __ldcs(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void stv(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void stv(bf16* p, const float (&v)[8]) {
  uint4 u;
  uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&a);
  }
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  if (act == ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  return v;
}
__device__ __forceinline__ float act_grad(float out, int act, float slope) {
  if (act == ACT_LRELU) return out > 0.f ? 1.f : slope;
  if (act == ACT_RELU) return out > 0.f ? 1.f : 0.f;
  return 1.f;
}

// ---------------------------------------------------------------------------------------------
// Column reductions over a row slice of an [M][C] matrix.
// Block = 256 threads; C/V threads span a row (V channels each); 256/(C/V) rows per pass, two passes in
// flight per thread.  Each block writes part[blockIdx.x][NV*C].
// ---------------------------------------------------------------------------------------------
struct RowSlice {
  long long begin, end;
};
__device__ __forceinline__ RowSlice row_slice(long long M) {
  const long long rps = (M + gridDim.x - 1) / gridDim.x;
  RowSlice r;
  r.begin = (long long)blockIdx.x * rps;
  r.end = r.begin + rps < M ? r.begin + rps : M;
  return r;
}

template <int NV, int V>
__device__ __forceinline__ void block_col_reduce(float (&acc)[NV][V], int C, float* part_row) {
  // threads with the same (threadIdx.x % (C/V)) own the same V channels
  __shared__ float red[256 * V];
  const int lpr = C / V;
  const int cg = threadIdx.x % lpr;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < V; ++j) red[threadIdx.x * V + j] = acc[v][j];
    __syncthreads();
    if (threadIdx.x < lpr) {
      float s[V];
#pragma unroll
      for (int j = 0; j < V; ++j) s[j] = 0.f;
      for (int r = threadIdx.x; r < 256; r += lpr) {
#pragma unroll
        for (int j = 0; j < V; ++j) s[j] += red[r * V + j];
      }
#pragma unroll
      for (int j = 0; j < V; ++j) part_row[v * C + cg * V + j] = s[j];
    }
  }
}

static void check_colshape(int C) {
  PCG_REQUIRE(C % 4 == 0 && C <= 1024 && (256 % (C / 4)) == 0, "channel count must be 4*2^k <= 1024");
}
// 8-wide vectors need C % 8 == 0, a power-of-two lane group and (V * sizeof(T))-byte aligned rows
template <typename T>
static bool wide_ok(int C, std::initializer_list<const void*> ptrs) {
  if (C % 8 != 0 || (256 % (C / 8)) != 0) return false;
  for (const void* p : ptrs)
    if (p != nullptr && (reinterpret_cast<uintptr_t>(p) % (8 * sizeof(T))) != 0) return false;
  return true;
}

// Walks the block's row slice two rows at a time: body(r) is called for each row with its loads issued
// back to back (the compiler hoists both rows' loads above the arithmetic).
#define PCG_ROWS2(sl, r0, rpp, BODY)                                   \
  {                                                                    \
    long long r = (sl).begin + (r0);                                   \
    for (; r + (rpp) < (sl).end; r += 2 * (rpp)) { BODY(r, r + (rpp)) } \
    if (r < (sl).end) { BODY(r, -1) }                                  \
  }

template <typename T, int V>
__global__ void __launch_bounds__(256, 4) bn_stats_kernel(const T* __restrict__ y, long long M, int C,
                                                      float* __restrict__ part) {
  pdl_enter();
  const int lpr = C / V, rpp = 256 / lpr;
  const int cg = threadIdx.x % lpr, r0 = threadIdx.x / lpr;
  const RowSlice sl = row_slice(M);
  float acc[2][V] = {};
#define BODY(ra, rb)                                                   \
  {                                                                    \
    float va[V], vb[V];                                                \
    ldv(y + (ra) * C + cg * V, va);                                    \
    if ((rb) >= 0) ldv(y + (rb) * C + cg * V, vb);                     \
    _Pragma("unroll") for (int j = 0; j < V; ++j) {                    \
      acc[0][j] += va[j];                                              \
      acc[1][j] = fmaf(va[j], va[j], acc[1][j]);                       \
    }                                                                  \
    if ((rb) >= 0) {                                                   \
      _Pragma("unroll") for (int j = 0; j < V; ++j) {                  \
        acc[0][j] += vb[j];                                            \
        acc[1][j] = fmaf(vb[j], vb[j], acc[1][j]);                     \
      }                                                                \
    }                                                                  \
  }
  PCG_ROWS2(sl, r0, rpp, BODY)
#undef BODY
  block_col_reduce<2, V>(acc, C, part + (size_t)blockIdx.x * 2 * C);
}

template <typename T>
void bn_stats_partial(const T* y, long long M, int C, float* part, cudaStream_t s) {
  PCG_PROFILE("bn_stats", s);
  check_colshape(C);
  if (wide_ok<T>(C, {y})) launch_k(bn_stats_kernel<T, 8>, dim3(STAT_PARTS), dim3(256), 0, s, y, M, C, part);
  else launch_k(bn_stats_kernel<T, 4>, dim3(STAT_PARTS), dim3(256), 0, s, y, M, C, part);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// Fixed-order parallel sum of column c over `nparts` partial rows by one warp (lane l takes rows
// l, l+32, ... then a shuffle tree): deterministic and ~30x shorter dependency chain than a serial loop.
__device__ __forceinline__ double warp_colsum(const float* __restrict__ part, int nparts, size_t stride, int c) {
  // four loads in flight per lane: these kernels are pure L2-latency chains (a lane reads up to 19 rows)
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = threadIdx.x & 31;
  for (; i + 96 < nparts; i += 128) {
    const float a = part[(size_t)i * stride + c], b = part[(size_t)(i + 32) * stride + c];
    const float d = part[(size_t)(i + 64) * stride + c], e = part[(size_t)(i + 96) * stride + c];
    s0 += (double)a; s1 += (double)b; s2 += (double)d; s3 += (double)e;
  }
  for (; i < nparts; i += 32) s0 += (double)part[(size_t)i * stride + c];
  return warp_sum((s0 + s1) + (s2 + s3));
}

// One 128-thread block per channel: every thread reads at most ceil(nparts/128) rows per column (one L2 round trip
// instead of a 19-deep chain per lane), then a fixed-order tree (warp shuffles, 4 warps through shared memory):
// deterministic.  These kernels sit on the critical path 36 times per step.
constexpr int FIN_THREADS = 128;
template <int NS>
__device__ __forceinline__ void block_colsum(const float* __restrict__ part, int nparts, size_t stride,
                                             const int (&cols)[NS], double (&out)[NS]) {
  __shared__ double red[FIN_THREADS / 32][NS];
  double s[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
  for (int i = threadIdx.x; i < nparts; i += FIN_THREADS) {
#pragma unroll
    for (int k = 0; k < NS; ++k) s[k] += (double)part[(size_t)i * stride + cols[k]];
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = warp_sum(s[k]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) red[threadIdx.x >> 5][k] = s[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NS; ++k) out[k] = (red[0][k] + red[1][k]) + (red[2][k] + red[3][k]);
}

__global__ void bn_finalize_kernel(const float* __restrict__ part, int nparts, long long M, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, long long* nbt,
                                   float* mean_o, float* rstd_o, float* scale, float* shift) {
  pdl_enter();
  const int c = blockIdx.x;                                        // one block per channel
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
  const int cols[2] = {c, C + c};
  double sums[2];
  block_colsum<2>(part, nparts, 2 * (size_t)C, cols, sums);
  const double s = sums[0], q = sums[1];
  if (threadIdx.x != 0) return;
  const double mean = s / (double)M;
  double var = q / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float a = gamma[c] * rstd;
  mean_o[c] = (float)mean;
  rstd_o[c] = rstd;
  scale[c] = a;
  shift[c] = beta[c] - (float)mean * a;
  if (running_mean != nullptr) {
    const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

void bn_finalize(const float* part, int nparts, long long M, int C, const float* gamma, const float* beta,
                 float eps, float momentum, float* running_mean, float* running_var, long long* nbt, float* mean,
                 float* rstd, float* scale, float* shift, cudaStream_t s) {
  PCG_PROFILE("bn_finalize", s);
  launch_k_small(1, bn_finalize_kernel, dim3(C), dim3(FIN_THREADS), 0, s, part, nparts, M, C, gamma, beta, eps, momentum, running_mean,
                                               running_var, nbt, mean, rstd, scale, shift);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

static int ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// Element-wise BatchNorm-apply kernels: thread i owns vectors i, i + S, ... of V channels where the stride S (total
// threads) is a multiple of C/V, so a thread always sees the same channels and keeps scale/shift in registers;
// two vectors in flight; the activation is a template parameter.
template <typename T, int V, int ACT>
__global__ void __launch_bounds__(256) bn_apply_act_kernel(const T* __restrict__ y, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, long long nv, int C,
                                                          float slope, T* __restrict__ z, bf16* __restrict__ side,
                                                          int stream_loads) {
  pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)((i0 * V) % C);
  float a[V], b[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { a[j] = scale[c0 + j]; b[j] = shift[c0 + j]; }
  for (long long i = i0; i < nv; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < nv;
    float va[V], vb[V];
    if (stream_loads) {                            // PCG_L2_HINTS bit 8: y is not read again before the backward
      ldv_cs<T, V>(y + i * V, va);
      if (two) ldv_cs<T, V>(y + i2 * V, vb);
    } else {
      ldv(y + i * V, va);
      if (two) ldv(y + i2 * V, vb);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) va[j] = act_fwd(fmaf(va[j], a[j], b[j]), ACT, slope);
    stv(z + i * V, va);
    if (side) stv(side + i * V, va);           // bf16 copy for the tensor-core operand cache (fp32-storage plans)
    if (two) {
#pragma unroll
      for (int j = 0; j < V; ++j) vb[j] = act_fwd(fmaf(vb[j], a[j], b[j]), ACT, slope);
      stv(z + i2 * V, vb);
      if (side) stv(side + i2 * V, vb);
    }
  }
}

// grid whose total thread count is a multiple of `period` vectors (so a thread's channel group never changes)
static int ew_blocks_periodic(long long nvec, int period) {
  int blocks = ew_blocks((nvec + 1) / 2);
  while (((long long)blocks * 256) % period != 0) ++blocks;
  return blocks;
}

template <typename T>
void bn_apply_act(const T* y, const float* scale, const float* shift, long long M, int C, int act, float slope, T* z,
                  cudaStream_t s, bf16* side) {
  PCG_PROFILE("bn_apply", s);
  PCG_REQUIRE(C % 4 == 0, "C % 4");
#define PCG_L(V, A) launch_k_small(2, bn_apply_act_kernel<T, V, A>, dim3(ew_blocks_periodic(nv, C / V)), dim3(256), 0, s, y, scale, shift, nv, C, slope, z, side, (g_l2_hints & 8) ? 1 : 0)
#define PCG_LA(V) { if (act == ACT_LRELU) PCG_L(V, ACT_LRELU); else if (act == ACT_RELU) PCG_L(V, ACT_RELU); else PCG_L(V, ACT_NONE); }
  if (C % 8 == 0 && wide_ok<T>(8, {y, z})) {
    const long long nv = M * C / 8;
    PCG_LA(8)
  } else {
    const long long nv = M * C / 4;
    PCG_LA(4)
  }
#undef PCG_LA
#undef PCG_L
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
bn_apply_residual_kernel(const T* __restrict__ y, const T* __restrict__ h, const float* __restrict__ scale,
                         const float* __restrict__ shift, float res_scale, long long nv, int C, T* __restrict__ out,
                         int stream_loads) {
  pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)((i0 * V) % C);
  float a[V], b[V];          // out = h + (res_scale*scale)*y + res_scale*shift
#pragma unroll
  for (int j = 0; j < V; ++j) { a[j] = res_scale * scale[c0 + j]; b[j] = res_scale * shift[c0 + j]; }
  for (long long i = i0; i < nv; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < nv;
    float va[V], ha[V], vb[V], hb[V];
    if (stream_loads) {                            // PCG_L2_HINTS bit 8
      ldv_cs<T, V>(y + i * V, va);
      ldv_cs<T, V>(h + i * V, ha);
      if (two) { ldv_cs<T, V>(y + i2 * V, vb); ldv_cs<T, V>(h + i2 * V, hb); }
    } else {
      ldv(y + i * V, va);
      ldv(h + i * V, ha);
      if (two) { ldv(y + i2 * V, vb); ldv(h + i2 * V, hb); }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) va[j] = ha[j] + fmaf(va[j], a[j], b[j]);
    stv(out + i * V, va);
    if (two) {
#pragma unroll
      for (int j = 0; j < V; ++j) vb[j] = hb[j] + fmaf(vb[j], a[j], b[j]);
      stv(out + i2 * V, vb);
    }
  }
}

template <typename T>
void bn_apply_residual(const T* y, const T* h, const float* scale, const float* shift, float res_scale, long long M,
                       int C, T* out, cudaStream_t s) {
  PCG_PROFILE("bn_apply", s);
  PCG_REQUIRE(C % 4 == 0, "C % 4");
  if (C % 8 == 0 && wide_ok<T>(8, {y, h, out})) {
    const long long nv = M * C / 8;
    launch_k_small(2, bn_apply_residual_kernel<T, 8>, dim3(ew_blocks_periodic(nv, C / 8)), dim3(256), 0, s, y, h, scale, shift, res_scale, nv, C, out, (g_l2_hints & 8) ? 1 : 0);
  } else {
    const long long nv = M * C / 4;
    launch_k_small(2, bn_apply_residual_kernel<T, 4>, dim3(ew_blocks_periodic(nv, C / 4)), dim3(256), 0, s, y, h, scale, shift, res_scale, nv, C, out, (g_l2_hints & 8) ? 1 : 0);
  }
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ---- BatchNorm backward.  These kernels are instruction-issue bound before they are HBM bound (bf16 unpacking plus
// ~10 operations per element), so: the activation is a template parameter, the per-channel algebra is folded into
// as few constants as possible, each thread owns 4 channels (constants stay in registers under 64 registers =
// four resident blocks per SM) and keeps four rows in flight as raw (still packed) loads.
template <typename T> struct Raw4;
template <> struct Raw4<float> { typedef float4 type; };
template <> struct Raw4<bf16> { typedef uint2 type; };
__device__ __forceinline__ void unpack4(const float4& r, float (&v)[4]) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
__device__ __forceinline__ void unpack4(const uint2& r, float (&v)[4]) {
  v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
  v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ typename Raw4<T>::type ldraw(const T* p) {
  return *reinterpret_cast<const typename Raw4<T>::type*>(p);
}
template <typename T>
__device__ __forceinline__ typename Raw4<T>::type ldraw_cs(const T* p) {
  return __ldcs(reinterpret_cast<const typename Raw4<T>::type*>(p));
}
template <typename T>
__device__ __forceinline__ typename Raw4<T>::type zero_raw() {
  typename Raw4<T>::type z;
  memset(&z, 0, sizeof(z));
  return z;
}
constexpr int BNB_UN = 4;    // rows in flight per thread

// g = gscale * dsrc * act'(scale*y + shift);  part[.][2C] = (sum g, sum g*xhat)
template <typename T, int ACT>
__global__ void __launch_bounds__(256, 4)
bn_bwd_partial_kernel(const T* __restrict__ dsrc, const T* __restrict__ y, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                      float gscale, float slope, long long M, int C, float* __restrict__ part) {
  pdl_enter();
  const int lpr = C >> 2, rpp = 256 / lpr;
  const int cg = threadIdx.x % lpr, r0 = threadIdx.x / lpr;
  const RowSlice sl = row_slice(M);
  float rs[4], nmr[4], a[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cg * 4 + j;
    rs[j] = rstd[c]; nmr[j] = -mean[c] * rstd[c];        // xhat = v*rs + nmr
    a[j] = scale[c]; b[j] = shift[c];
  }
  float acc[2][4] = {};
  for (long long r = sl.begin + r0; r < sl.end; r += BNB_UN * rpp) {
    typename Raw4<T>::type rd[BNB_UN], rv[BNB_UN];
#pragma unroll
    for (int u = 0; u < BNB_UN; ++u) {
      const long long rr = r + u * rpp;
      const bool ok = rr < sl.end;
      rd[u] = ok ? ldraw<T>(dsrc + rr * C + cg * 4) : zero_raw<T>();
      rv[u] = ok ? ldraw<T>(y + rr * C + cg * 4) : zero_raw<T>();
    }
#pragma unroll
    for (int u = 0; u < BNB_UN; ++u) {
      float d[4], v[4];
      unpack4(rd[u], d);
      unpack4(rv[u], v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float g = d[j];
        if (ACT == ACT_LRELU) g = fmaf(v[j], a[j], b[j]) > 0.f ? g : g * slope;
        if (ACT == ACT_RELU) g = fmaf(v[j], a[j], b[j]) > 0.f ? g : 0.f;
        acc[0][j] += g;
        acc[1][j] = fmaf(g, fmaf(v[j], rs[j], nmr[j]), acc[1][j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { acc[0][j] *= gscale; acc[1][j] *= gscale; }
  block_col_reduce<2, 4>(acc, C, part + (size_t)blockIdx.x * 2 * C);
}

template <typename T>
void bn_bwd_partial(const T* dsrc, const T* y, const float* mean, const float* rstd, const float* scale,
                    const float* shift, float gscale, int act, float slope, long long M, int C, float* part,
                    cudaStream_t s) {
  PCG_PROFILE("bn_bwd_reduce", s);
  check_colshape(C);
#define PCG_L(A) launch_k(bn_bwd_partial_kernel<T, A>, dim3(STAT_PARTS), dim3(256), 0, s, dsrc, y, mean, rstd, scale, shift, gscale, slope, M, C, part)
  if (act == ACT_LRELU) PCG_L(ACT_LRELU);
  else if (act == ACT_RELU) PCG_L(ACT_RELU);
  else PCG_L(ACT_NONE);
#undef PCG_L
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int nparts, long long M, int C, float* dgamma,
                                       float* dbeta, float* c12) {
  pdl_enter();
  const int c = blockIdx.x;
  const int cols[2] = {c, C + c};
  double sums[2];
  block_colsum<2>(part, nparts, 2 * (size_t)C, cols, sums);
  const double s1 = sums[0], s2 = sums[1];
  if (threadIdx.x != 0) return;
  dbeta[c] = (float)s1;
  dgamma[c] = (float)s2;
  c12[c] = (float)(s1 / (double)M);
  c12[C + c] = (float)(s2 / (double)M);
}

void bn_bwd_finalize(const float* part, int nparts, long long M, int C, float* dgamma, float* dbeta, float* c12,
                     cudaStream_t s) {
  PCG_PROFILE("bn_finalize", s);
  launch_k_small(1, bn_bwd_finalize_kernel, dim3(C), dim3(FIN_THREADS), 0, s, part, nparts, M, C, dgamma, dbeta, c12);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// dy = gamma*rstd * (g - c1 - xhat*c2) with g as above, folded per channel into
//   dy = ag * (dsrc * act') + nk2 * y + k0,   ag = a*gscale, nk2 = -a*c2*rstd, k0 = a*(c2*rstd*mean - c1), a = gamma*rstd
template <typename T, int ACT>
__global__ void __launch_bounds__(256, 4)
bn_bwd_apply_kernel(const T* __restrict__ dsrc, const T* __restrict__ y, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ c12, float gscale, float slope, long long M, int C,
                    T* __restrict__ dy, float* __restrict__ part_db, bf16* __restrict__ side, int stream_loads) {
  pdl_enter();
  const int lpr = C >> 2, rpp = 256 / lpr;
  const int cg = threadIdx.x % lpr, r0 = threadIdx.x / lpr;
  const RowSlice sl = row_slice(M);
  float a[4], b[4], ag[4], nk2[4], k0[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cg * 4 + j;
    a[j] = scale[c]; b[j] = shift[c];
    const float c2rs = c12[C + c] * rstd[c];
    ag[j] = a[j] * gscale;
    nk2[j] = -a[j] * c2rs;
    k0[j] = a[j] * (c2rs * mean[c] - c12[c]);
  }
  float acc[1][4] = {};
  for (long long r = sl.begin + r0; r < sl.end; r += BNB_UN * rpp) {
    typename Raw4<T>::type rd[BNB_UN], rv[BNB_UN];
#pragma unroll
    for (int u = 0; u < BNB_UN; ++u) {
      const long long rr = r + u * rpp;
      const bool ok = rr < sl.end;
      if (stream_loads) {                          // last use of both inputs: evict-first loads (PCG_L2_HINTS bit 2)
        rd[u] = ok ? ldraw_cs<T>(dsrc + rr * C + cg * 4) : zero_raw<T>();
        rv[u] = ok ? ldraw_cs<T>(y + rr * C + cg * 4) : zero_raw<T>();
      } else {
        rd[u] = ok ? ldraw<T>(dsrc + rr * C + cg * 4) : zero_raw<T>();
        rv[u] = ok ? ldraw<T>(y + rr * C + cg * 4) : zero_raw<T>();
      }
    }
#pragma unroll
    for (int u = 0; u < BNB_UN; ++u) {
      const long long rr = r + u * rpp;
      if (rr < sl.end) {
        float d[4], v[4], o[4];
        unpack4(rd[u], d);
        unpack4(rv[u], v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float g = d[j];
          if (ACT == ACT_LRELU) g = fmaf(v[j], a[j], b[j]) > 0.f ? g : g * slope;
          if (ACT == ACT_RELU) g = fmaf(v[j], a[j], b[j]) > 0.f ? g : 0.f;
          o[j] = fmaf(ag[j], g, fmaf(nk2[j], v[j], k0[j]));
          acc[0][j] += o[j];
        }
        st4(dy + rr * C + cg * 4, o);
        if (side) st4(side + rr * C + cg * 4, o);
      }
    }
  }
  block_col_reduce<1, 4>(acc, C, part_db + (size_t)blockIdx.x * C);
}

template <typename T>
void bn_bwd_apply(const T* dsrc, const T* y, const float* mean, const float* rstd, const float* scale,
                  const float* shift, const float* gamma, const float* c12, float gscale, int act, float slope,
                  long long M, int C, T* dy, float* part_db, cudaStream_t s, bf16* side) {
  PCG_PROFILE("bn_bwd_apply", s);
  (void)gamma;
  check_colshape(C);
#define PCG_L(A) launch_k_small(2, bn_bwd_apply_kernel<T, A>, dim3(STAT_PARTS), dim3(256), 0, s, dsrc, y, mean, rstd, scale, shift, c12, gscale, slope, M, C, dy, part_db, side, (g_l2_hints & 2) ? 1 : 0)
  if (act == ACT_LRELU) PCG_L(ACT_LRELU);
  else if (act == ACT_RELU) PCG_L(ACT_RELU);
  else PCG_L(ACT_NONE);
#undef PCG_L
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

__global__ void colsum_finalize_kernel(const float* __restrict__ part, int nparts, int stride, int C, float* out) {
  pdl_enter();
  const int c = blockIdx.x;
  const int cols[1] = {c};
  double sums[1];
  block_colsum<1>(part, nparts, (size_t)stride, cols, sums);
  if (threadIdx.x == 0) out[c] = (float)sums[0];
}
void colsum_finalize(const float* part, int nparts, int stride, int C, float* out, cudaStream_t s) {
  PCG_PROFILE("small", s);
  launch_k_small(1, colsum_finalize_kernel, dim3(C), dim3(FIN_THREADS), 0, s, part, nparts, stride, C, out);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// Several layers' column sums in one launch: slot k's partials are part + k * nparts * C (row stride C), its result goes
// to outs.p[k].  Same fixed-order sum as colsum_finalize, so the results are bit-identical to one launch per layer.
__global__ void colsum_finalize_multi_kernel(const float* __restrict__ part, int nparts, int C, ColsumOuts outs) {
  pdl_enter();
  const int c = blockIdx.x, k = blockIdx.y;
  const int cols[1] = {c};
  double sums[1];
  block_colsum<1>(part + (size_t)k * nparts * C, nparts, (size_t)C, cols, sums);
  if (threadIdx.x == 0) outs.p[k][c] = (float)sums[0];
}
void colsum_finalize_multi(const float* part, int nparts, int C, const ColsumOuts& outs, int nslots, cudaStream_t s) {
  PCG_PROFILE("small", s);
  PCG_REQUIRE(nslots >= 1 && nslots <= ColsumOuts::MAX, "colsum_finalize_multi: 1..16 slots per launch");
  launch_k(colsum_finalize_multi_kernel, dim3(C, nslots), dim3(FIN_THREADS), 0, s, part, nparts, C, outs);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

template <typename T, int V>
__global__ void __launch_bounds__(256, 4) colsum_kernel(const T* __restrict__ a, long long M, int C,
                                                    float* __restrict__ part) {
  pdl_enter();
  const RowSlice sl = row_slice(M);
  if (V > 1) {
    const int lpr = C / V, rpp = 256 / lpr;
    const int cg = threadIdx.x % lpr, r0 = threadIdx.x / lpr;
    float acc[1][V] = {};
#define BODY(ra, rb)                                                   \
  {                                                                    \
    float va[V], vb[V];                                                \
    ldv(a + (ra) * C + cg * V, va);                                    \
    if ((rb) >= 0) ldv(a + (rb) * C + cg * V, vb);                     \
    _Pragma("unroll") for (int j = 0; j < V; ++j) acc[0][j] += va[j];  \
    if ((rb) >= 0) {                                                   \
      _Pragma("unroll") for (int j = 0; j < V; ++j) acc[0][j] += vb[j]; \
    }                                                                  \
  }
    PCG_ROWS2(sl, r0, rpp, BODY)
#undef BODY
    block_col_reduce<1, V>(acc, C, part + (size_t)blockIdx.x * C);
  }
}
// tiny / odd channel counts (e.g. C = 1): one column at a time, block tree reduction
template <typename T>
__global__ void __launch_bounds__(256) colsum_scalar_kernel(const T* __restrict__ a, long long M, int C,
                                                           float* __restrict__ part) {
  pdl_enter();
  const RowSlice sl = row_slice(M);
  __shared__ float red[256];
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
    for (long long r = sl.begin + threadIdx.x; r < sl.end; r += 256) s += to_f(a[r * C + c]);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * C + c] = red[0];
    __syncthreads();
  }
}
// wide matrices whose column count fits neither vector kernel (e.g. C = 4096): a thread per column, rows of the slice in turn
template <typename T>
__global__ void __launch_bounds__(256) colsum_wide_kernel(const T* __restrict__ a, long long M, int C,
                                                         float* __restrict__ part) {
  pdl_enter();
  const RowSlice sl = row_slice(M);
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.f;
    for (long long r = sl.begin; r < sl.end; ++r) s += to_f(a[r * C + c]);
    part[(size_t)blockIdx.x * C + c] = s;
  }
}
template <typename T>
void colsum_partial(const T* a, long long M, int C, float* part, cudaStream_t s) {
  PCG_PROFILE("colsum", s);
  if (wide_ok<T>(C, {a})) launch_k(colsum_kernel<T, 8>, dim3(STAT_PARTS), dim3(256), 0, s, a, M, C, part);
  else if ((C & 3) == 0 && (256 % (C >> 2)) == 0 && C <= 1024 && (reinterpret_cast<uintptr_t>(a) % (4 * sizeof(T))) == 0)
    launch_k(colsum_kernel<T, 4>, dim3(STAT_PARTS), dim3(256), 0, s, a, M, C, part);
  else if (C >= 64) launch_k(colsum_wide_kernel<T>, dim3(STAT_PARTS), dim3(256), 0, s, a, M, C, part);
  else launch_k(colsum_scalar_kernel<T>, dim3(STAT_PARTS), dim3(256), 0, s, a, M, C, part);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// input assembly / embedding gradient
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void g_input_kernel(const float* __restrict__ x, const float* __restrict__ embed,
                               const long long* __restrict__ label, const float* __restrict__ mask, long long total,
                               int HW, T* __restrict__ out) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW), p = (int)(i - (long long)n * HW);
    out[i * 3 + 0] = from_f<T>(x[i]);
    out[i * 3 + 1] = from_f<T>(embed[(size_t)label[n] * HW + p]);
    out[i * 3 + 2] = from_f<T>(mask[i]);
  }
}
template <typename T>
void g_input(const float* x, const float* embed, const long long* label, const float* mask, int B, int HW, T* out,
             cudaStream_t s) {
  PCG_PROFILE("small", s);
  const long long total = (long long)B * HW;
  launch_k(g_input_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, s, x, embed, label, mask, total, HW, out);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

template <typename T>
__global__ void d_input_kernel(const float* __restrict__ x, const float* __restrict__ embed,
                               const long long* __restrict__ label, long long total, int HW, T* __restrict__ out) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW), p = (int)(i - (long long)n * HW);
    out[i * 2 + 0] = from_f<T>(x[i]);
    out[i * 2 + 1] = from_f<T>(embed[(size_t)label[n] * HW + p]);
  }
}
template <typename T>
void d_input(const float* x, const float* embed, const long long* label, int B, int HW, T* out, cudaStream_t s) {
  PCG_PROFILE("small", s);
  const long long total = (long long)B * HW;
  launch_k(d_input_kernel<T>, dim3(ew_blocks(total)), dim3(256), 0, s, x, embed, label, total, HW, out);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// Generic fallback: one thread per (class, pixel), serial over the batch.
template <typename T>
__global__ void embed_grad_serial_kernel(const T* __restrict__ src, int nch, int ch, const long long* __restrict__ label,
                                         int B, int HW, float* __restrict__ dE) {
  pdl_enter();
  const int cls = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  float s = 0.f;
  for (int n = 0; n < B; ++n) {
    if (label[n] == cls) s += to_f(src[((size_t)n * HW + p) * nch + ch]);
  }
  dE[(size_t)cls * HW + p] = s;
}
// <= 16 classes: a block owns 32 pixels (one per lane); warp w walks samples w, w+16, ... with one register
// accumulator per class (predicated adds, 8 loads in flight), then the 16 warps are summed in fixed order.
constexpr int EG_WARPS = 16, EG_CLS = 16;
template <typename T>
__global__ void __launch_bounds__(EG_WARPS * 32)
embed_grad_kernel(const T* __restrict__ src, int nch, int ch, const long long* __restrict__ label, int B, int HW,
                  int num_classes, float* __restrict__ dE) {
  pdl_enter();
  __shared__ float red[EG_WARPS][EG_CLS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + lane;
  const bool ok = p < HW;
  float acc[EG_CLS];
#pragma unroll
  for (int c = 0; c < EG_CLS; ++c) acc[c] = 0.f;
  for (int n0 = warp; n0 < B; n0 += EG_WARPS * 8) {
    float v[8];
    int lab[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int n = n0 + u * EG_WARPS;
      const bool in = n < B;
      lab[u] = in ? (int)label[n] : -1;
      v[u] = (in && ok) ? to_f(src[((size_t)n * HW + p) * nch + ch]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < EG_CLS; ++c) acc[c] += (lab[u] == c) ? v[u] : 0.f;
  }
#pragma unroll
  for (int c = 0; c < EG_CLS; ++c) red[warp][c][lane] = acc[c];
  __syncthreads();
  for (int i = threadIdx.x; i < num_classes * 32; i += blockDim.x) {
    const int c = i >> 5, l = i & 31;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < EG_WARPS; ++w) s += red[w][c][l];
    const int pp = blockIdx.x * 32 + l;
    if (pp < HW) dE[(size_t)c * HW + pp] = s;
  }
}
template <typename T>
void embed_grad(const T* src, int nch, int ch, const long long* label, int B, int HW, int num_classes, float* dE,
                cudaStream_t s) {
  PCG_PROFILE("embed_grad", s);
  if (num_classes <= EG_CLS) {
    launch_k(embed_grad_kernel<T>, dim3(cdiv(HW, 32)), dim3(EG_WARPS * 32), 0, s, src, nch, ch, label, B, HW, num_classes, dE);
  } else {
    dim3 grid(cdiv(HW, 128), num_classes);
    launch_k(embed_grad_serial_kernel<T>, dim3(grid), dim3(128), 0, s, src, nch, ch, label, B, HW, dE);
  }
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// residual head
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
residual_head_fwd_kernel(const float* __restrict__ c, const float* __restrict__ x, const float* __restrict__ mask,
                         float rs, long long n, float* __restrict__ raw, float* __restrict__ masked,
                         float* __restrict__ x_cf, float* __restrict__ part) {
  pdl_enter();
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long b = (long long)blockIdx.x * per, e = b + per < n ? b + per : n;
  float s0 = 0.f, s1 = 0.f;
  for (long long i = b + threadIdx.x; i < e; i += 256) {
    const float r = c[i] * rs, m = mask[i];
    const float mr = r * m;
    raw[i] = r;
    masked[i] = mr;
    x_cf[i] = fminf(fmaxf(x[i] + mr, -1.f), 1.f);
    s0 += fabsf(mr);
    s1 += fabsf(r * (1.f - m));
  }
  __shared__ float red[2][8];
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, d = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; d += red[1][w]; }
    part[blockIdx.x * 2 + 0] = a;
    part[blockIdx.x * 2 + 1] = d;
  }
}
void residual_head_fwd(const float* c, const float* x, const float* mask, float rs, long long n, float* raw,
                       float* masked, float* x_cf, float* part, cudaStream_t s) {
  PCG_PROFILE("residual_head", s);
  launch_k(residual_head_fwd_kernel, dim3(STAT_PARTS), dim3(256), 0, s, c, x, mask, rs, n, raw, masked, x_cf, part);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

template <typename T>
__global__ void residual_head_bwd_kernel(const float* __restrict__ dxd, int dxd_ch, const float* __restrict__ dxc,
                                         const float* __restrict__ raw, const float* __restrict__ x,
                                         const float* __restrict__ mask, float rs, float lreg, float lmask,
                                         long long n, T* __restrict__ g_c) {
  pdl_enter();
  const float inv_n = 1.f / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float r = raw[i], m = mask[i];
    const float mr = r * m;
    const float pre = x[i] + mr;
    const float pass = (pre >= -1.f && pre <= 1.f) ? 1.f : 0.f;
    const float dxcf = dxd[i * dxd_ch] + dxc[i];
    const float d_masked = pass * dxcf + lreg * sgn(mr) * inv_n;
    const float d_raw = d_masked * m + lmask * sgn(r * (1.f - m)) * (1.f - m) * inv_n;
    g_c[i] = from_f<T>(rs * d_raw);
  }
}
template <typename T>
void residual_head_bwd(const float* dxd, int dxd_ch, const float* dxc, const float* raw, const float* x,
                       const float* mask, float rs, float lreg, float lmask, long long n, T* g_c, cudaStream_t s) {
  PCG_PROFILE("residual_head", s);
  launch_k(residual_head_bwd_kernel<T>, dim3(ew_blocks(n)), dim3(256), 0, s, dxd, dxd_ch, dxc, raw, x, mask, rs, lreg, lmask, n, g_c);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// discriminator head, BCE, CE
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void d_head_fwd_kernel(const T* __restrict__ z, int B, int HW, int C, const float* __restrict__ w,
                                  const float* __restrict__ b, float* __restrict__ logits) {
  pdl_enter();
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= B) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    float m = 0.f;
    for (int p = 0; p < HW; ++p) m += to_f(z[((size_t)n * HW + p) * C + c]);
    s = fmaf(m / (float)HW, w[c], s);
  }
  s = warp_sum(s);
  if (lane == 0) logits[n] = s + b[0];
}
template <typename T>
void d_head_fwd(const T* z, int B, int HW, int C, const float* w, const float* b, float* logits, cudaStream_t s) {
  PCG_PROFILE("small", s);
  launch_k(d_head_fwd_kernel<T>, dim3(cdiv(B, 8)), dim3(256), 0, s, z, B, HW, C, w, b, logits);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

__device__ __forceinline__ float block_sum_1024(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;   // valid in warp 0
}

// one block per segment
__global__ void bce_logits_kernel(const float* __restrict__ logits, int seg, float t0, float t1, float w0, float w1,
                                  float* out_loss, float* out_p, float* __restrict__ dlogit) {
  pdl_enter();
  __shared__ float red[32];
  const int sidx = blockIdx.x;
  const float t = sidx == 0 ? t0 : t1, wgt = sidx == 0 ? w0 : w1;
  float sl = 0.f, sp = 0.f;
  for (int i = threadIdx.x; i < seg; i += blockDim.x) {
    const float z = logits[sidx * seg + i];
    const float sig = 1.f / (1.f + expf(-z));
    sl += fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));
    sp += sig;
    dlogit[sidx * seg + i] = wgt * (sig - t) / (float)seg;
  }
  sl = block_sum_1024(sl, red);
  sp = block_sum_1024(sp, red);
  if (threadIdx.x == 0) {
    out_loss[sidx] = sl / (float)seg;
    out_p[sidx] = sp / (float)seg;
  }
}
void bce_logits(const float* logits, int seg, int nseg, float t0, float t1, float w0, float w1, float* out_loss,
                float* out_p, float* dlogit, cudaStream_t s) {
  PCG_PROFILE("small", s);
  PCG_REQUIRE(nseg == 1 || nseg == 2, "1 or 2 segments");
  launch_k(bce_logits_kernel, dim3(nseg), dim3(256), 0, s, logits, seg, t0, t1, w0, w1, out_loss, out_p, dlogit);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// grid.x = B (one block per sample) for g; dw/db by a second tiny kernel.
template <typename T>
__global__ void d_head_bwd_kernel(const T* __restrict__ z, const float* __restrict__ dlogit, int HW, int C,
                                  const float* __restrict__ w, float slope, T* __restrict__ g) {
  pdl_enter();
  const int n = blockIdx.x;
  const float dl = dlogit[n] / (float)HW;
  for (int i = threadIdx.x; i < HW * C; i += blockDim.x) {
    const int c = i % C;
    const float zz = to_f(z[(size_t)n * HW * C + i]);
    g[(size_t)n * HW * C + i] = from_f<T>(dl * w[c] * (zz > 0.f ? 1.f : slope));
  }
}
// part[blockIdx.x][c] = sum over the block's sample slice of dlogit[n] * mean_hw z[n][.][c]; one thread per channel
// (coalesced over c), fixed order within the slice; the slices are summed by colsum_finalize.
constexpr int DHW_SLICES = 64;
template <typename T>
__global__ void __launch_bounds__(256)
d_head_wgrad_kernel(const T* __restrict__ z, const float* __restrict__ dlogit, int B, int HW, int C,
                    float* __restrict__ part, float* __restrict__ db) {
  pdl_enter();
  const int per = (B + gridDim.x - 1) / gridDim.x;
  const int n0 = blockIdx.x * per, n1 = n0 + per < B ? n0 + per : B;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int n = n0; n < n1; ++n) {
      float m = 0.f;
      for (int p = 0; p < HW; ++p) m += to_f(z[((size_t)n * HW + p) * C + c]);
      s = fmaf(dlogit[n], m / (float)HW, s);
    }
    part[(size_t)blockIdx.x * C + c] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    float t = 0.f;
    for (int n = threadIdx.x; n < B; n += 32) t += dlogit[n];
    t = warp_sum(t);
    if (threadIdx.x == 0) db[0] = t;
  }
}
template <typename T>
void d_head_bwd(const T* z, const float* dlogit, int B, int HW, int C, const float* w, float slope, T* g, float* dw,
                float* db, float* scratch, cudaStream_t s) {
  {
    PCG_PROFILE("small", s);
    launch_k(d_head_bwd_kernel<T>, dim3(B), dim3(256), 0, s, z, dlogit, HW, C, w, slope, g);
    PCG_COUNT_LAUNCH();
    PCG_LAUNCH_CHECK();
    if (dw != nullptr) {
      launch_k(d_head_wgrad_kernel<T>, dim3(DHW_SLICES), dim3(256), 0, s, z, dlogit, B, HW, C, scratch, db);
      PCG_COUNT_LAUNCH();
      PCG_LAUNCH_CHECK();
    }
  }
  if (dw != nullptr) colsum_finalize(scratch, DHW_SLICES, C, C, dw, s);
}

__global__ void ce_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int B, int NC,
                               float wgt, float* loss, float* __restrict__ dlogits) {
  pdl_enter();
  __shared__ float red[32];
  __shared__ float slot;
  long long rb, re;
  cluster_slice(B, rb, re);
  float sl = 0.f;
  for (int n = (int)rb + threadIdx.x; n < (int)re; n += blockDim.x) {
    const float* l = logits + (size_t)n * NC;
    float mx = l[0];
    for (int j = 1; j < NC; ++j) mx = fmaxf(mx, l[j]);
    float se = 0.f;
    for (int j = 0; j < NC; ++j) se += expf(l[j] - mx);
    const float lse = mx + logf(se);
    const int t = (int)target[n];
    sl += lse - l[t];
    for (int j = 0; j < NC; ++j) {
      const float p = expf(l[j] - lse);
      dlogits[(size_t)n * NC + j] = wgt * (p - (j == t ? 1.f : 0.f)) / (float)B;
    }
  }
  sl = cluster_total(block_sum_1024(sl, red), &slot);
  if (cluster_leader()) loss[0] = sl / (float)B;
}
void ce_loss(const float* logits, const long long* target, int B, int NC, float wgt, float* loss, float* dlogits,
             cudaStream_t s) {
  PCG_PROFILE("small", s);
  launch_k_cluster(ce_loss_kernel, B, dim3(B >= CR_MIN_ITEMS ? 512 : 256), 0, s, logits, target, B, NC, wgt, loss, dlogits);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// CrossEntropyLoss(weight = w) with mean reduction (torch: sum_n w[t_n] * nll_n / sum_n w[t_n]; w == nullptr: plain mean)
// + optionally the number of rows whose arg-max is the target (the accuracy counters of
// house_sales_kc_usa/trainer.py:92-95,110-112).  One block; deterministic.
__global__ void ce_loss_weighted_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                        const float* __restrict__ w, int B, int NC, float* loss,
                                        float* __restrict__ dlogits, float* correct) {
  pdl_enter();
  __shared__ float red[32];
  __shared__ float s_wsum;
  float sw = 0.f;
  for (int n = threadIdx.x; n < B; n += blockDim.x) sw += w ? w[(int)target[n]] : 1.f;
  sw = block_sum_1024(sw, red);
  if (threadIdx.x == 0) s_wsum = sw;
  __syncthreads();
  const float wsum = s_wsum;
  float sl = 0.f, sc = 0.f;
  for (int n = threadIdx.x; n < B; n += blockDim.x) {
    const float* l = logits + (size_t)n * NC;
    float mx = l[0];
    int am = 0;
    for (int j = 1; j < NC; ++j)
      if (l[j] > mx) { mx = l[j]; am = j; }
    float se = 0.f;
    for (int j = 0; j < NC; ++j) se += expf(l[j] - mx);
    const float lse = mx + logf(se);
    const int t = (int)target[n];
    const float wn = w ? w[t] : 1.f;
    sl += wn * (lse - l[t]);
    sc += am == t ? 1.f : 0.f;
    if (dlogits != nullptr)
      for (int j = 0; j < NC; ++j) dlogits[(size_t)n * NC + j] = wn * (expf(l[j] - lse) - (j == t ? 1.f : 0.f)) / wsum;
  }
  __syncthreads();
  sl = block_sum_1024(sl, red);
  __syncthreads();
  sc = block_sum_1024(sc, red);
  if (threadIdx.x == 0) {
    loss[0] = sl / wsum;
    if (correct != nullptr) correct[0] = sc;
  }
}
void ce_loss_weighted(const float* logits, const long long* target, const float* w, int B, int NC, float* loss,
                      float* dlogits, float* correct, cudaStream_t s) {
  PCG_PROFILE("small", s);
  launch_k(ce_loss_weighted_kernel, dim3(1), dim3(256), 0, s, logits, target, w, B, NC, loss, dlogits, correct);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

__global__ void l1_finalize_kernel(const float* __restrict__ part, int nparts, float inv_n, float* out2) {
  pdl_enter();
  const int c = threadIdx.x >> 5;
  if (c < 2) {
    const double s = warp_colsum(part, nparts, 2, c);
    if ((threadIdx.x & 31) == 0) out2[c] = (float)(s * (double)inv_n);
  }
}
void l1_finalize(const float* part, int nparts, float inv_n, float* out2, cudaStream_t s) {
  PCG_PROFILE("small", s);
  launch_k(l1_finalize_kernel, dim3(1), dim3(64), 0, s, part, nparts, inv_n, out2);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

__global__ void g_loss_combine_kernel(const float* g_adv, const float* g_cls, const float* reg, const float* mpen,
                                      float la, float lc, float lr, float lm, float* out) {
  pdl_enter();
  out[0] = la * g_adv[0] + lc * g_cls[0] + lr * reg[0] + lm * mpen[0];
}
void g_loss_combine(const float* g_adv, const float* g_cls, const float* reg, const float* mpen, float la, float lc,
                    float lr, float lm, float* out, cudaStream_t s) {
  PCG_PROFILE("small", s);
  launch_k(g_loss_combine_kernel, dim3(1), dim3(1), 0, s, g_adv, g_cls, reg, mpen, la, lc, lr, lm, out);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// Adam
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, const int* __restrict__ step, float lr, float beta1, float beta2, float eps,
                 float grad_scale) {
  pdl_enter();
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = (double)(*step + 1);
    const double bc1 = 1.0 - pow((double)beta1, t);
    const double bc2 = 1.0 - pow((double)beta2, t);
    s_step_size = (float)((double)lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  // four elements per thread and load when the arenas allow it (they do: FlatParams pads every tensor to 4 floats): the
  // 14-24 M parameter networks of the conditional WGAN-GP are HBM-bound here
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                      reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long nv = vec ? n >> 2 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    const float gs[4] = {g4.x, g4.y, g4.z, g4.w};
    float* ms = reinterpret_cast<float*>(&m4);
    float* vs = reinterpret_cast<float*>(&v4);
    float* ps = reinterpret_cast<float*>(&p4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gi = gs[e] * grad_scale;
      ms[e] = ms[e] + (gi - ms[e]) * (1.f - beta1);
      vs[e] = vs[e] * beta2 + (1.f - beta2) * gi * gi;
      const float denom = sqrtf(vs[e]) / bc2_sqrt + eps;
      ps[e] = ps[e] - step_size * (ms[e] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
  }
  for (long long i = nv * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float mi = m[i], vi = v[i];
    mi = mi + (gi - mi) * (1.f - beta1);                 // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * beta2 + (1.f - beta2) * gi * gi;           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);              // param.addcdiv_(exp_avg, denom, -step_size)
    m[i] = mi;
    v[i] = vi;
  }
}
__global__ void adam_step_inc_kernel(int* step) {
  pdl_enter(); *step += 1; }

void adam_flat(float* p, const float* g, float* m, float* v, long long n, int* step, float lr, float beta1,
               float beta2, float eps, float grad_scale, cudaStream_t s) {
  PCG_PROFILE("adam", s);
  launch_k(adam_flat_kernel, dim3(ew_blocks(n)), dim3(256), 0, s, p, g, m, v, n, step, lr, beta1, beta2, eps, grad_scale);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
  launch_k(adam_step_inc_kernel, dim3(1), dim3(1), 0, s, step);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// torch.optim.AdamW (decoupled weight decay, optim/adamw.py: param.mul_(1 - lr * weight_decay), then the Adam update).
// The learning rate is read from device memory so that a scheduler (ReduceLROnPlateau, house_sales_kc_usa/trainer.py:61)
// can change it between replays of a captured graph.
__global__ void __launch_bounds__(256)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  long long n, const int* __restrict__ step, const float* __restrict__ lr_dev, float beta1, float beta2,
                  float eps, float weight_decay) {
  pdl_enter();
  __shared__ float s_step_size, s_bc2_sqrt, s_decay;
  if (threadIdx.x == 0) {
    const double t = (double)(*step + 1), lr = (double)lr_dev[0];
    s_step_size = (float)(lr / (1.0 - pow((double)beta1, t)));
    s_bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
    s_decay = 1.f - lr_dev[0] * weight_decay;
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt, decay = s_decay;
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                      reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long nv = vec ? n >> 2 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    const float gs[4] = {g4.x, g4.y, g4.z, g4.w};
    float* ms = reinterpret_cast<float*>(&m4);
    float* vs = reinterpret_cast<float*>(&v4);
    float* ps = reinterpret_cast<float*>(&p4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ms[e] = ms[e] + (gs[e] - ms[e]) * (1.f - beta1);
      vs[e] = vs[e] * beta2 + (1.f - beta2) * gs[e] * gs[e];
      const float denom = sqrtf(vs[e]) / bc2_sqrt + eps;
      ps[e] = ps[e] * decay - step_size * (ms[e] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
  }
  for (long long i = nv * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float mi = m[i], vi = v[i];
    mi = mi + (gi - mi) * (1.f - beta1);
    vi = vi * beta2 + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] * decay - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}
void adamw_flat(float* p, const float* g, float* m, float* v, long long n, int* step, const float* lr_dev, float beta1,
                float beta2, float eps, float weight_decay, cudaStream_t s) {
  PCG_PROFILE("adam", s);
  launch_k(adamw_flat_kernel, dim3(ew_blocks(n)), dim3(256), 0, s, p, g, m, v, n, step, lr_dev, beta1, beta2, eps, weight_decay);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
  launch_k(adam_step_inc_kernel, dim3(1), dim3(1), 0, s, step);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

template <typename T>
__global__ void fill_zero_kernel(T* p, long long n) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = from_f<T>(0.f);
}
template <typename T>
void fill_zero(T* p, long long n, cudaStream_t s) {
  PCG_PROFILE("small", s);
  launch_k(fill_zero_kernel<T>, dim3(ew_blocks(n)), dim3(256), 0, s, p, n);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
template <typename T>
__global__ void convert_kernel(const float* __restrict__ src, long long n, T* __restrict__ dst) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f<T>(src[i]);
}
template <typename T>
void convert_from_f32(const float* src, long long n, T* dst, cudaStream_t s) {
  PCG_PROFILE("pack_weights", s);
  launch_k(convert_kernel<T>, dim3(ew_blocks(n)), dim3(256), 0, s, src, n, dst);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
#define INST(T)                                                                                                    \
  template void bn_stats_partial<T>(const T*, long long, int, float*, cudaStream_t);                               \
  template void bn_apply_act<T>(const T*, const float*, const float*, long long, int, int, float, T*, cudaStream_t, bf16*); \
  template void bn_apply_residual<T>(const T*, const T*, const float*, const float*, float, long long, int, T*,    \
                                     cudaStream_t);                                                                \
  template void bn_bwd_partial<T>(const T*, const T*, const float*, const float*, const float*, const float*,      \
                                  float, int, float, long long, int, float*, cudaStream_t);                        \
  template void bn_bwd_apply<T>(const T*, const T*, const float*, const float*, const float*, const float*,        \
                                const float*, const float*, float, int, float, long long, int, T*, float*,          \
                                cudaStream_t, bf16*);                                                              \
  template void colsum_partial<T>(const T*, long long, int, float*, cudaStream_t);                                 \
  template void g_input<T>(const float*, const float*, const long long*, const float*, int, int, T*, cudaStream_t); \
  template void d_input<T>(const float*, const float*, const long long*, int, int, T*, cudaStream_t);              \
  template void embed_grad<T>(const T*, int, int, const long long*, int, int, int, float*, cudaStream_t);          \
  template void residual_head_bwd<T>(const float*, int, const float*, const float*, const float*, const float*,    \
                                     float, float, float, long long, T*, cudaStream_t);                            \
  template void d_head_fwd<T>(const T*, int, int, int, const float*, const float*, float*, cudaStream_t);          \
  template void d_head_bwd<T>(const T*, const float*, int, int, int, const float*, float, T*, float*, float*,      \
                              float*, cudaStream_t);                                                                       \
  template void fill_zero<T>(T*, long long, cudaStream_t);                                                         \
  template void convert_from_f32<T>(const float*, long long, T*, cudaStream_t);
INST(float)
INST(bf16)
#undef INST

}  // namespace pcg

// One-channel 4x4 / stride-2 / pad-1 convolution kernels (fp32 NHWC, CUDA cores).
//
// DCGAN's first discriminator layer Conv2d(1, 64, 4, 2, 1) and last generator layer ConvTranspose2d(64, 1, 4, 2, 1)
// (dconv_gan/mnist/mnist_dcgan.py:89,100) carry 1 % of the step's FLOPs but each pass streams a full 64-channel map
// (67 MB at batch 256): they are HBM-bound (~11 us), and took ~95 us each on the generic implicit-GEMM kernels, whose
// tiles are built for K = taps x Cin >> 16.  Both layers are the SAME three operators on the geometry
// [N, H, W, 1] <-> [N, H/2, W/2, C] (a ConvTranspose2d forward is the data gradient of the mirrored convolution):
//   c1k4_fprop   out[n, oy, ox, c] = act( sum_{ky,kx} x[n, 2oy-1+ky, 2ox-1+kx] * wf[c][ky*4+kx] )
//                  D conv0 forward (+ LeakyReLU); input gradient of G's last ConvTranspose2d
//   c1k4_dgrad   dx[n, iy, ix]     = sum_c sum_{taps} dy[n, oy, ox, c] * wd[ky*4+kx][c]   (iy = 2oy-1+ky, ix = 2ox-1+kx)
//                  G's last ConvTranspose2d forward; D conv0 input gradient (the generator's adversarial gradient)
//   c1k4_wgrad   dw[c][ky*4+kx]    = sum_{n,oy,ox} x[n, 2oy-1+ky, 2ox-1+kx] * dy[n, oy, ox, c]
// Requirements: C == 64, W in {16, 32, 48, 64}, H % 8 == 0.  Anything else stays on the generic kernels.
#include "conv_c1k4.cuh"

#include "elementwise.cuh"

namespace pcg {

constexpr int C1_C = 64;            // output channels
constexpr int C1_ROWS = 4;          // output rows (of the C-channel map) per block

bool c1k4_supported(const ConvGeom& g) {
  return g.Cin == 1 && g.Cout == C1_C && g.ksize == 4 && g.stride == 2 && g.pad == 1 && (g.H & 1) == 0 && (g.W & 1) == 0 &&
         g.W <= 64 && g.W % 16 == 0 && g.H >= 8 && (g.H / 2) % C1_ROWS == 0;    // W % 16: whole warps in c1k4_dgrad step 1
}

// ------------------------------------------------------------------------------------------------ forward
// Block: one image, C1_ROWS output rows.  The 2*C1_ROWS + 2 input rows live in shared memory (zero padded, one pad
// column on the left so that input column 2ox-1 sits at even index 2ox); a thread owns four channels (its 16 x 4 weights
// stay in registers) and walks the block's output positions; consecutive threads write consecutive 16-byte groups.
__global__ void __launch_bounds__(256)
c1k4_fprop_kernel(const float* __restrict__ x, const float* __restrict__ wf, int H, int W, int act, float slope,
                  float* __restrict__ out, bf16* __restrict__ side) {
  pdl_enter();
  __shared__ __align__(16) float sx[2 * C1_ROWS + 2][68];
  __shared__ __align__(16) float sw[16][C1_C + 4];            // [tap][c]: a thread's four channels are one 16-byte read
  const int Ho = H >> 1, Wo = W >> 1;
  const int n = blockIdx.x / (Ho / C1_ROWS), oy0 = (blockIdx.x % (Ho / C1_ROWS)) * C1_ROWS;
  const float* xi = x + (size_t)n * H * W;
  for (int i = threadIdx.x; i < (2 * C1_ROWS + 2) * 68; i += 256) {
    const int r = i / 68, col = i - r * 68;
    const int iy = 2 * oy0 - 1 + r, ix = col - 1;
    sx[r][col] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? xi[(size_t)iy * W + ix] : 0.f;
  }
  for (int i = threadIdx.x; i < C1_C * 16; i += 256) sw[i & 15][i >> 4] = wf[i];      // coalesced read of wf[c][tap]
  __syncthreads();
  const int c4 = threadIdx.x & 15, p0 = threadIdx.x >> 4;       // 16 position lanes x 16 channel quads
  float4 w[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) w[t] = *reinterpret_cast<const float4*>(&sw[t][c4 * 4]);
  float* o = out + ((size_t)n * Ho + oy0) * Wo * C1_C;
  bf16* o16 = side ? side + ((size_t)n * Ho + oy0) * Wo * C1_C : nullptr;     // bf16 copy for the tensor-core operand cache
  for (int p = p0; p < C1_ROWS * Wo; p += 16) {
    const int oyl = p / Wo, ox = p - oyl * Wo;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const float2 x01 = *reinterpret_cast<const float2*>(&sx[2 * oyl + ky][2 * ox]);
      const float2 x23 = *reinterpret_cast<const float2*>(&sx[2 * oyl + ky][2 * ox + 2]);
      const float xv[4] = {x01.x, x01.y, x23.x, x23.y};
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const float4 ww = w[ky * 4 + kx];
        a.x = fmaf(xv[kx], ww.x, a.x); a.y = fmaf(xv[kx], ww.y, a.y);
        a.z = fmaf(xv[kx], ww.z, a.z); a.w = fmaf(xv[kx], ww.w, a.w);
      }
    }
    if (act == ACT_LRELU) {
      a.x = lrelu(a.x, slope); a.y = lrelu(a.y, slope); a.z = lrelu(a.z, slope); a.w = lrelu(a.w, slope);
    } else if (act == ACT_RELU) {
      a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    }
    *reinterpret_cast<float4*>(o + (size_t)p * C1_C + c4 * 4) = a;
    if (o16) {
      __nv_bfloat162 h[2] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w)};
      *reinterpret_cast<uint2*>(o16 + (size_t)p * C1_C + c4 * 4) = *reinterpret_cast<const uint2*>(h);
    }
  }
}

void c1k4_fprop(const float* x, const ConvGeom& g, const float* wf, int act, float slope, float* out, cudaStream_t s,
                bf16* side) {
  PCG_PROFILE("conv_c1k4", s);
  PCG_REQUIRE(c1k4_supported(g), "c1k4 geometry");
  launch_k(c1k4_fprop_kernel, dim3(g.N * (g.H / 2 / C1_ROWS)), dim3(256), 0, s, x, wf, g.H, g.W, act, slope, out, side);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------ data gradient
// dy row oy feeds input rows 2oy-1 .. 2oy+2, so the 2*C1_ROWS input rows [2*oy0, 2*oy0 + 2*C1_ROWS) of a block need the
// dy rows oy0-1 .. oy0+C1_ROWS (C1_ROWS + 2 rows).  Step 1: P[row][ox][tap] = sum_c dy[row][ox][c] * wd[tap][c] (a
// [positions x 64] x [64 x 16] product).  Step 2: every input pixel gathers its four (row, ox, tap) entries.
__global__ void __launch_bounds__(256)
c1k4_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ wd, int H, int W, float* __restrict__ dx) {
  pdl_enter();
  __shared__ float sp[C1_ROWS + 2][32][17];                    // [row][ox][tap], padded
  const int Ho = H >> 1, Wo = W >> 1;
  const int n = blockIdx.x / (Ho / C1_ROWS), oy0 = (blockIdx.x % (Ho / C1_ROWS)) * C1_ROWS;
  // Step 1.  A thread owns four channels (c4) with all 16 taps of them in registers; the 16 lanes of a half warp cover
  // the 64 channels of ONE position (one coalesced 256-byte read of dy), each forms its 16 partial tap sums, and a
  // transpose-reduce over the half warp (15 shuffles) leaves tap (lane & 15) of that position in every lane.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = lane & 15, half = lane >> 4;
  float4 w[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) w[t] = __ldg(reinterpret_cast<const float4*>(wd + t * C1_C + c4 * 4));     // wd is [tap][c]
  const int npos = (C1_ROWS + 2) * Wo;
  for (int pos = warp * 2 + half; pos < npos; pos += 16) {
    const int r = pos / Wo, ox = pos - r * Wo;
    const int oy = oy0 - 1 + r;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (oy >= 0 && oy < Ho) d = __ldg(reinterpret_cast<const float4*>(dy + (((size_t)n * Ho + oy) * Wo + ox) * C1_C) + c4);
    float v[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = fmaf(d.w, w[t].w, fmaf(d.z, w[t].z, fmaf(d.y, w[t].y, d.x * w[t].x)));
#pragma unroll
    for (int off = 8, cnt = 16; off >= 1; off >>= 1, cnt >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int j = 0; j < cnt / 2; ++j) {
        const float send = upper ? v[j] : v[j + cnt / 2];
        const float keep = upper ? v[j + cnt / 2] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    sp[r][ox][c4] = v[0];                                       // tap index = lane & 15
  }
  __syncthreads();
  float* o = dx + ((size_t)n * H + 2 * oy0) * W;
  for (int p = threadIdx.x; p < 2 * C1_ROWS * W; p += 256) {
    const int iyl = p / W, ix = p - iyl * W;
    const int iy = 2 * oy0 + iyl;
    // iy = 2oy - 1 + ky: oy = (iy+1)/2 with ky = (iy+1)&1, and oy - 1 with ky + 2
    const int my = (iy + 1) >> 1, ry = (iy + 1) & 1, mx = (ix + 1) >> 1, rx = (ix + 1) & 1;
    float a = 0.f;
#pragma unroll
    for (int dyy = 0; dyy < 2; ++dyy) {
      const int oy = my - dyy, ky = ry + 2 * dyy;
      if (oy < 0 || oy >= Ho) continue;
      const int r = oy - (oy0 - 1);
#pragma unroll
      for (int dxx = 0; dxx < 2; ++dxx) {
        const int ox = mx - dxx, kx = rx + 2 * dxx;
        if (ox < 0 || ox >= Wo) continue;
        a += sp[r][ox][ky * 4 + kx];
      }
    }
    o[p] = a;
  }
}

void c1k4_dgrad(const float* dy, const ConvGeom& g, const float* wd, float* dx, cudaStream_t s) {
  PCG_PROFILE("conv_c1k4", s);
  PCG_REQUIRE(c1k4_supported(g), "c1k4 geometry");
  launch_k(c1k4_dgrad_kernel, dim3(g.N * (g.H / 2 / C1_ROWS)), dim3(256), 0, s, dy, wd, g.H, g.W, dx);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------ weight gradient
// A block walks a contiguous range of (image, row group) work items and keeps its 64 x 16 sums in registers: thread
// (c, g) owns channel c and the 16 taps, g = 0..3 picks one of the C1_ROWS output rows of the item.  Per position: one
// coalesced dy load and the 4 x 4 input window as eight 8-byte shared-memory reads (broadcast within the warp).
// The four row groups are added in fixed order, one partial row [64*16] per block leaves; colsum_finalize adds the rows.
__global__ void __launch_bounds__(256)
c1k4_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int H, int W, int items,
                  float* __restrict__ part) {
  pdl_enter();
  __shared__ __align__(16) float sx[2 * C1_ROWS + 2][68];
  __shared__ float sred[3][C1_C][17];
  const int Ho = H >> 1, Wo = W >> 1, groups = Ho / C1_ROWS;
  const int c = threadIdx.x & 63, g = threadIdx.x >> 6;
  float acc[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) acc[t] = 0.f;
  const int per = (items + gridDim.x - 1) / gridDim.x;
  const int i_begin = blockIdx.x * per, i_end = min(items, i_begin + per);
  for (int item = i_begin; item < i_end; ++item) {
    const int n = item / groups, oy0 = (item % groups) * C1_ROWS;
    const float* xi = x + (size_t)n * H * W;
    __syncthreads();
    for (int i = threadIdx.x; i < (2 * C1_ROWS + 2) * 68; i += 256) {
      const int r = i / 68, col = i - r * 68;
      const int iy = 2 * oy0 - 1 + r, ix = col - 1;
      sx[r][col] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? xi[(size_t)iy * W + ix] : 0.f;
    }
    __syncthreads();
    const float* d = dy + (((size_t)n * Ho + oy0 + g) * Wo) * C1_C + c;
#pragma unroll 2
    for (int ox = 0; ox < Wo; ++ox) {
      const float dv = __ldg(d + (size_t)ox * C1_C);
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const float2 x01 = *reinterpret_cast<const float2*>(&sx[2 * g + ky][2 * ox]);
        const float2 x23 = *reinterpret_cast<const float2*>(&sx[2 * g + ky][2 * ox + 2]);
        acc[ky * 4 + 0] = fmaf(dv, x01.x, acc[ky * 4 + 0]);
        acc[ky * 4 + 1] = fmaf(dv, x01.y, acc[ky * 4 + 1]);
        acc[ky * 4 + 2] = fmaf(dv, x23.x, acc[ky * 4 + 2]);
        acc[ky * 4 + 3] = fmaf(dv, x23.y, acc[ky * 4 + 3]);
      }
    }
  }
  __syncthreads();
  if (g > 0) {
#pragma unroll
    for (int t = 0; t < 16; ++t) sred[g - 1][c][t] = acc[t];
  }
  __syncthreads();
  if (g == 0) {
    float* dst = part + (size_t)blockIdx.x * (C1_C * 16) + c * 16;
#pragma unroll
    for (int t = 0; t < 16; ++t) dst[t] = ((acc[t] + sred[0][c][t]) + sred[1][c][t]) + sred[2][c][t];
  }
}

int c1k4_wgrad_parts() { return sm_count() * 2; }
size_t c1k4_wgrad_scratch() { return (size_t)c1k4_wgrad_parts() * C1_C * 16; }

void c1k4_wgrad(const float* x, const float* dy, const ConvGeom& g, float* scratch, float* dw, cudaStream_t s) {
  PCG_PROFILE("wgrad_c1k4", s);
  PCG_REQUIRE(c1k4_supported(g), "c1k4 geometry");
  static_assert(C1_ROWS == 4, "thread layout: four row groups of 64 channels");
  const int items = g.N * (g.H / 2 / C1_ROWS);
  const int parts = c1k4_wgrad_parts();
  launch_k(c1k4_wgrad_kernel, dim3(parts), dim3(256), 0, s, x, dy, g.H, g.W, items, scratch);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
  colsum_finalize(scratch, parts, C1_C * 16, C1_C * 16, dw, s);        // dw is torch OIHW [64][1][4][4] = [c][tap]
}

}  // namespace pcg

// ================================================================================================
// Full-window convolution to ONE output: Conv2d(C, 1, k, 1, 0) on a k x k map (DCGAN's last discriminator layer,
// Conv2d(512, 1, 4, 1, 0) on 4 x 4, mnist_dcgan.py:112): per sample a dot product of length K = k*k*C, its data gradient
// an outer product, its weight gradient a weighted column sum.  8 MB at batch 256: ~3 us of HBM time each, against
// ~80 us on the generic implicit-GEMM kernels (whose tiles have 1 useful column of 64).
// ================================================================================================
namespace pcg {

bool full1_supported(const ConvGeom& g) {
  return g.Cout == 1 && g.stride == 1 && g.pad == 0 && g.ksize == g.H && g.ksize == g.W && (g.K() % 4) == 0 && g.K() <= 65536;
}

// out[n] = bias + sum_k x[n][k] * wf[k]   (x NHWC flattened = (tap, ci) order = wf's order)
__global__ void __launch_bounds__(256) full1_fprop_kernel(const float* __restrict__ x, const float* __restrict__ wf, int K,
                                                          const float* __restrict__ bias, float* __restrict__ out) {
  pdl_enter();
  __shared__ float red[8];
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)blockIdx.x * K);
  const float4* w4 = reinterpret_cast<const float4*>(wf);
  float a = 0.f;
  for (int i = threadIdx.x; i < K / 4; i += 256) {
    const float4 xv = __ldg(xr + i), wv = __ldg(w4 + i);
    a = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, a))));
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = bias ? bias[0] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i];
    out[blockIdx.x] = s;
  }
}
void full1_fprop(const float* x, const ConvGeom& g, const float* wf, const float* bias, float* out, cudaStream_t s) {
  PCG_PROFILE("conv_full1", s);
  launch_k(full1_fprop_kernel, dim3(g.N), dim3(256), 0, s, x, wf, g.K(), bias, out);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// dx[n][tap*C + ci] = dz[n] * wd[ci*taps + tap]
__global__ void __launch_bounds__(256) full1_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ wd, int N,
                                                          int C, int taps, float* __restrict__ dx) {
  pdl_enter();
  const int K = C * taps;
  const long long total4 = (long long)N * K / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    const int n = (int)(e / K), k = (int)(e - (long long)n * K);
    const int tap = k / C, ci = k - tap * C;                 // C % 4 == 0: the four elements share the tap
    const float z = __ldg(dz + n);
    float4 o;
    o.x = z * __ldg(wd + (ci + 0) * taps + tap);
    o.y = z * __ldg(wd + (ci + 1) * taps + tap);
    o.z = z * __ldg(wd + (ci + 2) * taps + tap);
    o.w = z * __ldg(wd + (ci + 3) * taps + tap);
    *reinterpret_cast<float4*>(dx + e) = o;
  }
}
void full1_dgrad(const float* dz, const ConvGeom& g, const float* wd, float* dx, cudaStream_t s) {
  PCG_PROFILE("conv_full1", s);
  PCG_REQUIRE(g.Cin % 4 == 0, "full-window data gradient: Cin % 4");
  const long long total4 = (long long)g.N * g.K() / 4;
  long long b = (total4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  launch_k(full1_dgrad_kernel, dim3((int)(b < cap ? (b > 0 ? b : 1) : cap)), dim3(256), 0, s, dz, wd, g.N, g.Cin,
           g.ksize * g.ksize, dx);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// dw[ci*taps + tap] = sum_n dz[n] * x[n][tap*C + ci]; 8 sample groups per 4 columns, added in fixed order
__global__ void __launch_bounds__(256) full1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, int N,
                                                          int C, int taps, float* __restrict__ dw) {
  pdl_enter();
  __shared__ float4 sm[8][32];
  const int K = C * taps;
  const int v = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int k4 = blockIdx.x * 32 + v;                         // columns 4*k4 .. 4*k4+3
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k4 * 4 < K) {
    const float4* col = reinterpret_cast<const float4*>(x) + k4;
#pragma unroll 4
    for (int n = grp; n < N; n += 8) {
      const float z = __ldg(dz + n);
      const float4 xv = __ldg(col + (size_t)n * (K / 4));
      a.x = fmaf(z, xv.x, a.x); a.y = fmaf(z, xv.y, a.y); a.z = fmaf(z, xv.z, a.z); a.w = fmaf(z, xv.w, a.w);
    }
  }
  sm[grp][v] = a;
  __syncthreads();
  if (grp == 0 && k4 * 4 < K) {
    float4 t = sm[0][v];
#pragma unroll
    for (int j = 1; j < 8; ++j) { const float4 u = sm[j][v]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    const int k = k4 * 4, tap = k / C, ci = k - tap * C;
    dw[(ci + 0) * taps + tap] = t.x;
    dw[(ci + 1) * taps + tap] = t.y;
    dw[(ci + 2) * taps + tap] = t.z;
    dw[(ci + 3) * taps + tap] = t.w;
  }
}
void full1_wgrad(const float* x, const float* dz, const ConvGeom& g, float* dw, cudaStream_t s) {
  PCG_PROFILE("wgrad_full1", s);
  PCG_REQUIRE(g.Cin % 4 == 0, "full-window weight gradient: Cin % 4");
  launch_k(full1_wgrad_kernel, dim3((g.K() / 4 + 31) / 32), dim3(256), 0, s, x, dz, g.N, g.Cin, g.ksize * g.ksize, dw);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

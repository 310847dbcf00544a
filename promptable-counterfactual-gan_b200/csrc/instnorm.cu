// InstanceNorm2d forward / backward / double backward on NHWC activations (see instnorm.cuh).
// One block per (sample, 32 channels): threadIdx.x = channel (coalesced rows), threadIdx.y = 8 slices of the P positions;
// the per-channel sums meet in shared memory.  The tensors are small (<= 128 x 169 x 256 floats) and L2 resident across the
// 2-3 sweeps a kernel makes over its sample.
#include "instnorm.cuh"

#include "conv_generic.cuh"

namespace pcg {

constexpr int IN_CX = 32, IN_PY = 8;

template <int K>
__device__ __forceinline__ void slice_sum(float (&v)[K], float (*sh)[IN_PY][IN_CX]) {
#pragma unroll
  for (int k = 0; k < K; ++k) sh[k][threadIdx.y][threadIdx.x] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < IN_PY; ++y) t += sh[k][y][threadIdx.x];
    v[k] = t;
  }
  __syncthreads();
}

__device__ __forceinline__ float act_apply(float v, int act, float slope) {
  return act == ACT_LRELU ? (v > 0.f ? v : v * slope) : (act == ACT_RELU ? fmaxf(v, 0.f) : v);
}
__device__ __forceinline__ float act_deriv(float ref, int act, float slope) {
  return act == ACT_LRELU ? (ref > 0.f ? 1.f : slope) : (act == ACT_RELU ? (ref > 0.f ? 1.f : 0.f) : 1.f);
}

__global__ void __launch_bounds__(IN_CX * IN_PY)
instnorm_fwd_kernel(const float* __restrict__ x, int P, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float eps, int act, float slope, float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_enter();
  __shared__ float sh[1][IN_PY][IN_CX];
  const int c = blockIdx.x * IN_CX + threadIdx.x, n = blockIdx.y;
  const bool on = c < C;
  const size_t base = (size_t)n * P * C + c;
  float v[1] = {0.f};
  if (on) for (int p = threadIdx.y; p < P; p += IN_PY) v[0] += x[base + (size_t)p * C];
  slice_sum(v, sh);
  const float mu = v[0] / P;
  v[0] = 0.f;
  if (on) for (int p = threadIdx.y; p < P; p += IN_PY) { const float d = x[base + (size_t)p * C] - mu; v[0] += d * d; }
  slice_sum(v, sh);
  const float rs = rsqrtf(v[0] / P + eps);
  if (!on) return;
  const float g = gamma[c], b = beta[c];
  for (int p = threadIdx.y; p < P; p += IN_PY)
    y[base + (size_t)p * C] = act_apply((x[base + (size_t)p * C] - mu) * rs * g + b, act, slope);
  if (threadIdx.y == 0) { mean[(size_t)n * C + c] = mu; rstd[(size_t)n * C + c] = rs; }
}

__global__ void __launch_bounds__(IN_CX * IN_PY)
instnorm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ act_ref, int act, float slope,
                    const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, int P, int C, const float* __restrict__ add_src, float* __restrict__ dx,
                    float* __restrict__ dgamma_part, float* __restrict__ dbeta_part) {
  pdl_enter();
  __shared__ float sh[2][IN_PY][IN_CX];
  const int c = blockIdx.x * IN_CX + threadIdx.x, n = blockIdx.y;
  const bool on = c < C;
  const size_t base = (size_t)n * P * C + c;
  const float mu = on ? mean[(size_t)n * C + c] : 0.f, rs = on ? rstd[(size_t)n * C + c] : 0.f;
  float v[2] = {0.f, 0.f};
  if (on) for (int p = threadIdx.y; p < P; p += IN_PY) {
    const size_t i = base + (size_t)p * C;
    const float pp = gy[i] * (act_ref ? act_deriv(act_ref[i], act, slope) : 1.f);
    v[0] += pp;
    v[1] += pp * (x[i] - mu) * rs;
  }
  slice_sum(v, sh);
  if (!on) return;
  const float c1 = v[0] / P, c2 = v[1] / P, g = gamma[c] * rs;
  for (int p = threadIdx.y; p < P; p += IN_PY) {
    const size_t i = base + (size_t)p * C;
    const float pp = gy[i] * (act_ref ? act_deriv(act_ref[i], act, slope) : 1.f);
    float o = g * (pp - c1 - (x[i] - mu) * rs * c2);
    if (add_src) o += add_src[i];
    dx[i] = o;
  }
  if (threadIdx.y == 0) {
    if (dgamma_part) dgamma_part[(size_t)n * C + c] = v[1];
    if (dbeta_part) dbeta_part[(size_t)n * C + c] = v[0];
  }
}

__global__ void __launch_bounds__(IN_CX * IN_PY)
instnorm_bwd_bwd_kernel(const float* __restrict__ q, const float* __restrict__ gy, const float* __restrict__ act_ref, int act,
                        float slope, const float* __restrict__ x, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ gamma, int P, int C,
                        float* __restrict__ gy_bar, float* __restrict__ x_bar, float* __restrict__ dgamma_part) {
  pdl_enter();
  __shared__ float sh[5][IN_PY][IN_CX];
  const int c = blockIdx.x * IN_CX + threadIdx.x, n = blockIdx.y;
  const bool on = c < C;
  const size_t base = (size_t)n * P * C + c;
  const float mu = on ? mean[(size_t)n * C + c] : 0.f, rs = on ? rstd[(size_t)n * C + c] : 0.f;
  float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};          // sum p, p*xhat, q, q*xhat, q*p
  if (on) for (int p = threadIdx.y; p < P; p += IN_PY) {
    const size_t i = base + (size_t)p * C;
    const float pp = gy[i] * (act_ref ? act_deriv(act_ref[i], act, slope) : 1.f);
    const float xh = (x[i] - mu) * rs, qq = q[i];
    v[0] += pp; v[1] += pp * xh; v[2] += qq; v[3] += qq * xh; v[4] += qq * pp;
  }
  slice_sum(v, sh);
  if (!on) return;
  const float c1 = v[0] / P, c2 = v[1] / P, d1 = v[2] / P, d2 = v[3] / P;
  const float sqr = v[4] - c1 * v[2] - c2 * v[3];          // sum q * (p - c1 - xhat * c2)
  const float g = gamma[c] * rs;
  const float mw = -g * (c2 * d1 + d2 * c1), mwx = -g * (2.f * c2 * d2);
  const float tail = g * rs * sqr / P;
  for (int p = threadIdx.y; p < P; p += IN_PY) {
    const size_t i = base + (size_t)p * C;
    const float dv = act_ref ? act_deriv(act_ref[i], act, slope) : 1.f;
    const float pp = gy[i] * dv, xh = (x[i] - mu) * rs, qq = q[i];
    gy_bar[i] = g * (qq - d1 - xh * d2) * dv;
    const float w = -g * (c2 * qq + d2 * pp);
    x_bar[i] = rs * (w - mw - xh * mwx) - tail * xh;
  }
  if (threadIdx.y == 0 && dgamma_part) dgamma_part[(size_t)n * C + c] = rs * sqr;
}

__global__ void flatten_nchw_kernel(const float* __restrict__ src, int R, int C, float* __restrict__ dst, int ld, int c0,
                                    int inverse) {
  pdl_enter();
  const int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * C; i += gridDim.x * blockDim.x) {
    const int r = i / C, c = i - r * C;
    const size_t wide = (size_t)b * ld + c0 + (size_t)c * R + r, dense = (size_t)b * R * C + i;
    if (inverse) dst[dense] = src[wide];
    else dst[wide] = src[dense];
  }
}

__global__ void bias_act_kernel(const float* __restrict__ x, long long n, int C, const float* __restrict__ bias, int tanh_out,
                                float* __restrict__ y) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i] + bias[i % C];
    y[i] = tanh_out ? tanhf(v) : v;
  }
}

// one block per sample; the mean over the batch is accumulated by the last block to finish (deterministic order: the
// per-sample terms are stored and summed by that block in index order)
__global__ void __launch_bounds__(256)
gp_penalty_kernel(const float* __restrict__ g, int B, int D, float lambda, float* __restrict__ out, float* __restrict__ gbar,
                  float* __restrict__ norms, float* __restrict__ terms, unsigned int* __restrict__ counter) {
  pdl_enter();
  __shared__ float sh[8];
  __shared__ float bc;
  __shared__ bool last;
  const int b = blockIdx.x;
  const float* row = g + (size_t)b * D;
  float acc = 0.f;
  for (int i = threadIdx.x; i < D; i += 256) acc += row[i] * row[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    const float nrm = sqrtf(t);
    if (norms) norms[b] = nrm;
    terms[b] = (nrm - 1.f) * (nrm - 1.f);
    bc = nrm > 0.f ? lambda * 2.f * (nrm - 1.f) / ((float)B * nrm) : 0.f;
  }
  __syncthreads();
  const float k = bc;
  for (int i = threadIdx.x; i < D; i += 256) gbar[(size_t)b * D + i] = k * row[i];
  __threadfence();
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == (unsigned)(B - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.f;
    for (int i = 0; i < B; ++i) t += ((volatile float*)terms)[i];
    out[0] = lambda * t / (float)B;
    *counter = 0u;
  }
}

// dst[n][y][x][c] = src[n][(y - off) / s][(x - off) / s][c] where that is a whole in-range position, else 0
__global__ void dilate_kernel(const float4* __restrict__ src, int Ho, int Wo, int C4, int s, int off, int Hp, int Wp,
                              float4* __restrict__ dst, long long total) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long r = i / C4;
    const int x = (int)(r % Wp); r /= Wp;
    const int y = (int)(r % Hp);
    const long long n = r / Hp;
    const int yy = y - off, xx = x - off;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && xx >= 0 && yy % s == 0 && xx % s == 0 && yy / s < Ho && xx / s < Wo)
      v = src[((n * Ho + yy / s) * Wo + xx / s) * C4 + c];
    dst[i] = v;
  }
}

// Stride-2 data gradient by parity classes.  With y = 2 i + c - pad (c = (y + pad) & 1) the taps that reach input row y are
// ky = c, c + 2, and dx_c[i] = sum_{m = 0, 1} dy[i - m] * W[c + 2 m]: per class (cy, cx) a stride-1 2x2 convolution of the
// UNdilated gradient (pad 1, output (Ho + 1) x (Wo + 1)) with the weights wc[cls][ci][jy * 2 + jx][co] =
// W[co][ci][cy + 2 (1 - jy)][cx + 2 (1 - jx)] (zero where that tap does not exist, k = 3), then an interleave.
__global__ void pack_dgrad_classes_kernel(const float* __restrict__ w, int Cout, int Cin, int k, float* __restrict__ wc) {
  pdl_enter();
  const long long total = 4LL * Cin * 4 * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    long long r = i / Cout;
    const int j = (int)(r % 4); r /= 4;
    const int ci = (int)(r % Cin);
    const int cls = (int)(r / Cin);
    const int ky = (cls >> 1) + 2 * (1 - (j >> 1)), kx = (cls & 1) + 2 * (1 - (j & 1));
    wc[i] = (ky < k && kx < k) ? w[(((size_t)co * Cin + ci) * k + ky) * k + kx] : 0.f;
  }
}

// dx[n][2 i + cy - pad][2 j + cx - pad][:] = src[cls][n][i][j][:] for the positions inside the H x W map; src holds the four
// class results [4][N][Hc][Wc][C] back to back
__global__ void parity_interleave_kernel(const float4* __restrict__ src, int N, int Hc, int Wc, int C4, int pad, int H, int W,
                                         int stacked, float4* __restrict__ dx, long long total) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long r = i / C4;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const long long n = r / H;
    const int cy = (y + pad) & 1, cx = (x + pad) & 1;
    const int iy = (y + pad - cy) >> 1, ix = (x + pad - cx) >> 1;
    const int cls = cy * 2 + cx;
    dx[i] = stacked ? src[(((n * Hc + iy) * Wc + ix) * 4 + cls) * C4 + c]                 // [N][Hc][Wc][4][C]: ONE convolution
                    : src[((((long long)cls * N + n) * Hc + iy) * Wc + ix) * C4 + c];     // [4][N][Hc][Wc][C]
  }
}

static dim3 in_grid(int N, int C) { return dim3((C + IN_CX - 1) / IN_CX, N); }

void instnorm_fwd(const float* x, int N, int P, int C, const float* gamma, const float* beta, float eps, int act, float slope,
                  float* y, float* mean, float* rstd, cudaStream_t s) {
  PCG_PROFILE("instnorm", s);
  PCG_REQUIRE(N >= 1 && N <= 65535 && P >= 1 && C >= 1, "instnorm: N in [1, 65535], P, C >= 1");
  launch_k(instnorm_fwd_kernel, in_grid(N, C), dim3(IN_CX, IN_PY), 0, s, x, P, C, gamma, beta, eps, act, slope, y, mean, rstd);
  PCG_COUNT_LAUNCH();
}

void instnorm_bwd(const float* gy, const float* act_ref, int act, float slope, const float* x, const float* mean,
                  const float* rstd, const float* gamma, int N, int P, int C, const float* add_src, float* dx,
                  float* dgamma_part, float* dbeta_part, cudaStream_t s) {
  PCG_PROFILE("instnorm", s);
  PCG_REQUIRE(N >= 1 && N <= 65535 && P >= 1 && C >= 1, "instnorm: N in [1, 65535], P, C >= 1");
  launch_k(instnorm_bwd_kernel, in_grid(N, C), dim3(IN_CX, IN_PY), 0, s, gy, act_ref, act, slope, x, mean, rstd, gamma, P, C,
           add_src, dx, dgamma_part, dbeta_part);
  PCG_COUNT_LAUNCH();
}

void instnorm_bwd_bwd(const float* q, const float* gy, const float* act_ref, int act, float slope, const float* x,
                      const float* mean, const float* rstd, const float* gamma, int N, int P, int C, float* gy_bar,
                      float* x_bar, float* dgamma_part, cudaStream_t s) {
  PCG_PROFILE("instnorm", s);
  PCG_REQUIRE(N >= 1 && N <= 65535 && P >= 1 && C >= 1, "instnorm: N in [1, 65535], P, C >= 1");
  launch_k(instnorm_bwd_bwd_kernel, in_grid(N, C), dim3(IN_CX, IN_PY), 0, s, q, gy, act_ref, act, slope, x, mean, rstd, gamma,
           P, C, gy_bar, x_bar, dgamma_part);
  PCG_COUNT_LAUNCH();
}

void flatten_nchw(const float* src, int B, int R, int C, float* dst, int ld, int c0, bool inverse, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  PCG_REQUIRE(B >= 1 && B <= 65535 && R >= 1 && C >= 1 && c0 >= 0 && c0 + R * C <= ld, "flatten_nchw: window inside the row");
  if (B == 1 && (long long)R * C >= (1 << 16)) {          // one big matrix: a plain transpose, tiled through shared memory
    if (inverse) transpose_tiled(src + c0, C, R, dst, s);
    else transpose_tiled(src, R, C, dst + c0, s);
    return;
  }
  const int blocks = (R * C + 255) / 256;
  const int cap = 64 > 2368 / B ? 64 : 2368 / B;
  launch_k(flatten_nchw_kernel, dim3(blocks < cap ? blocks : cap, B), dim3(256), 0, s, src, R, C, dst, ld, c0, inverse ? 1 : 0);
  PCG_COUNT_LAUNCH();
}

void bias_act(const float* x, long long rows, int C, const float* bias, int tanh_out, float* y, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  const long long n = rows * C;
  const long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  launch_k(bias_act_kernel, dim3((unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap)), dim3(256), 0, s, x, n, C, bias,
           tanh_out, y);
  PCG_COUNT_LAUNCH();
}

void dilate(const float* src, int N, int Ho, int Wo, int C, int stride, int off, int Hp, int Wp, float* dst, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  PCG_REQUIRE(C % 4 == 0 && stride >= 1 && off >= 0 && off + (Ho - 1) * stride < Hp && off + (Wo - 1) * stride < Wp,
              "dilate: C % 4 == 0 and the dilated grid inside the destination");
  PCG_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "dilate: 16-byte alignment");
  const long long total = (long long)N * Hp * Wp * (C / 4);
  const long long blocks = (total + 255) / 256, cap = (long long)sm_count() * 16;
  launch_k(dilate_kernel, dim3((unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap)), dim3(256), 0, s,
           reinterpret_cast<const float4*>(src), Ho, Wo, C / 4, stride, off, Hp, Wp, reinterpret_cast<float4*>(dst), total);
  PCG_COUNT_LAUNCH();
}

void pack_dgrad_classes(const float* w, int Cout, int Cin, int k, float* wc, cudaStream_t s) {
  PCG_PROFILE("pack_weights", s);
  PCG_REQUIRE(k == 3 || k == 4, "pack_dgrad_classes: k in {3, 4}");
  const long long total = 16LL * Cin * Cout;
  const long long blocks = (total + 255) / 256, cap = (long long)sm_count() * 16;
  launch_k(pack_dgrad_classes_kernel, dim3((unsigned)(blocks < cap ? blocks : cap)), dim3(256), 0, s, w, Cout, Cin, k, wc);
  PCG_COUNT_LAUNCH();
}

void parity_interleave(const float* src, int N, int Hc, int Wc, int C, int pad, int H, int W, bool stacked, float* dx,
                       cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  PCG_REQUIRE(C % 4 == 0 && (H - 1 + pad) / 2 < Hc && (W - 1 + pad) / 2 < Wc && pad >= 0,
              "parity_interleave: C % 4 == 0 and class maps that cover the H x W map");
  const long long total = (long long)N * H * W * (C / 4);
  const long long blocks = (total + 255) / 256, cap = (long long)sm_count() * 16;
  launch_k(parity_interleave_kernel, dim3((unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap)), dim3(256), 0, s,
           reinterpret_cast<const float4*>(src), N, Hc, Wc, C / 4, pad, H, W, stacked ? 1 : 0, reinterpret_cast<float4*>(dx),
           total);
  PCG_COUNT_LAUNCH();
}

void gp_penalty(const float* g, int B, int D, float lambda, float* out, float* gbar, float* norms, cudaStream_t s) {
  PCG_PROFILE("ops_small", s);
  PCG_REQUIRE(B >= 1 && B <= 65536 && D >= 1, "gp_penalty: B in [1, 65536]");
  static float* terms = nullptr;                   // [65536] per-sample terms + the arrival counter (zero between launches)
  static unsigned int* counter = nullptr;
  if (terms == nullptr) {
    PCG_CHECK_CUDA(cudaMalloc(&terms, 65536 * sizeof(float) + sizeof(unsigned int)));
    counter = reinterpret_cast<unsigned int*>(terms + 65536);
    PCG_CHECK_CUDA(cudaMemset(counter, 0, sizeof(unsigned int)));
  }
  launch_k(gp_penalty_kernel, dim3(B), dim3(256), 0, s, g, B, D, lambda, out, gbar, norms, terms, counter);
  PCG_COUNT_LAUNCH();
}

}  // namespace pcg

// Skinny convolutions of the MNIST CounteRGAN step (see conv_small.cuh).
#include "conv_small.cuh"

#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace pcg {
using namespace tc;

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------
// Cin -> 1, 3x3, stride 1, pad 1.
// A tile = R output rows of one image: one TMA box brings the (R+2) x (W+2) zero-padded input pixels (Cin bf16 per
// row, hardware swizzle).  Legacy mma.sync is ~16x slower than tcgen05 on this part (measured ~32 cycles per
// m16n8k16 and sub-core), so the contraction is arranged to need as few of them as possible: instead of one dot
// product of 9*Cin per output pixel (36 k-steps, 1 of 8 B columns used), every INPUT pixel q of the halo tile gets
//   P[q][tap] = sum_c x[q][c] * w[tap][c]      (K = Cin: 4 k-steps, 9 of 16 B columns used)
// and an output is the 9-term gather out[i] = bias + sum_{r,s} P[i + r*(W+2) + s][r*3 + s] from shared memory.
// ------------------------------------------------------------------------------------------
constexpr int C1_STAGES = 3;
constexpr int C1_ROWS = 192;                 // halo rows padded to 12 m16 blocks; >= (R+2)*(W+2) for W <= 28
constexpr int C1_CONSUMERS = 12;             // one m16 block of the halo tile per warp
constexpr int C1_THREADS = 32 * (C1_CONSUMERS + 1);   // + warp 12: TMA producer

template <int CIN>
__device__ __forceinline__ uint32_t c1_swizzle(uint32_t off) {
  return CIN == 64 ? (off ^ (((off >> 7) & 7u) << 4)) : (off ^ (((off >> 7) & 3u) << 4));
}

template <int CIN>
__global__ void __launch_bounds__(C1_THREADS, 2)
conv_to1_kernel(const __grid_constant__ CUtensorMap tmX, const bf16* __restrict__ w9, const float* __restrict__ bias,
                float* __restrict__ out, int H, int W, int WP, int R, int tiles_per_img, int total_tiles) {
  pdl_enter();
  constexpr int ROWB = CIN * 2, KQ = CIN / 16, STAGE_BYTES = C1_ROWS * ROWB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C1_STAGES * STAGE_BYTES);
  uint64_t* empty = full + C1_STAGES;
  float* P = reinterpret_cast<float*>(empty + C1_STAGES + 2);       // [2][C1_ROWS][9]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < C1_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], C1_CONSUMERS); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < C1_STAGES * STAGE_BYTES / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  __syncthreads();
  const int box_bytes = (R + 2) * WP * ROWB;

  if (warp == C1_CONSUMERS) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img, h0 = (tile % tiles_per_img) * R;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], box_bytes);
        tma_load_4d(&tmX, &full[stage], smem + stage * STAGE_BYTES, 0, -1, h0 - 1, n);
        if (++stage == C1_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    return;
  }

  const int q = lane & 3, rr = lane >> 2;
  // B[k = channel][n = tap]: n-tile 0 holds taps 0..7 (column rr), n-tile 1 tap 8 in column 0
  uint32_t breg[KQ][2][2];
#pragma unroll
  for (int ks = 0; ks < KQ; ++ks)
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      const int c = ks * 16 + 2 * q + hb * 8;
      breg[ks][0][hb] = *reinterpret_cast<const uint32_t*>(w9 + rr * CIN + c);
      breg[ks][1][hb] = rr == 0 ? *reinterpret_cast<const uint32_t*>(w9 + 8 * CIN + c) : 0u;
    }
  const float bv = bias != nullptr ? __ldg(bias) : 0.f;
  int stage = 0, it = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int n = tile / tiles_per_img, h0 = (tile % tiles_per_img) * R;
    float* Pt = P + (it & 1) * C1_ROWS * 9;
    mbar_wait(&full[stage], phase);
    const uint32_t base = smem_u32(smem + stage * STAGE_BYTES);
    float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const uint32_t rowoff = (uint32_t)(warp * 16 + (lane & 15)) * ROWB;
#pragma unroll
    for (int ks = 0; ks < KQ; ++ks) {
      uint32_t a[4];
      ldmatrix_x4(base + c1_swizzle<CIN>(rowoff + (uint32_t)(ks * 2 + (lane >> 4)) * 16u), a);
      mma_bf16_16816(c[0], a, breg[ks][0][0], breg[ks][0][1]);
      mma_bf16_16816(c[1], a, breg[ks][1][0], breg[ks][1][1]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      float* row = Pt + (warp * 16 + rr + t * 8) * 9;
      row[2 * q] = c[0][2 * t];
      row[2 * q + 1] = c[0][2 * t + 1];
      if (q == 0) row[8] = c[1][2 * t];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(C1_CONSUMERS * 32) : "memory");
    // gather: P is double-buffered, so the next tile's writes (other buffer) need no second barrier; the buffer
    // written two tiles later is protected by the barrier of the tile in between
    if (threadIdx.x < 128) {
      const int i = threadIdx.x;
      const int hh = i / WP, ww = i - hh * WP;
      if (hh < R && ww < W && h0 + hh < H) {
        float s = bv;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int sx = 0; sx < 3; ++sx) s += Pt[(i + r * WP + sx) * 9 + r * 3 + sx];
        out[((long long)n * H + h0 + hh) * W + ww] = s;
      }
    }
    if (++stage == C1_STAGES) { stage = 0; phase ^= 1; }
  }
}

bool conv_to1_supported(int H, int W, int Cin) {
  const int WP = W + 2;
  if (WP > 128 || (Cin != 32 && Cin != 64)) return false;
  const int R = 128 / WP;
  return R >= 1 && 127 + 2 * WP + 3 <= C1_ROWS && (R + 2) * WP <= C1_ROWS && R + 2 <= 256;
}

void conv_to1(const bf16* in, int N, int H, int W, int Cin, const bf16* w9, const float* bias, float* out,
              cudaStream_t stream) {
  PCG_PROFILE("conv_small", stream);
  PCG_REQUIRE(conv_to1_supported(H, W, Cin), "conv_to1: unsupported geometry");
  const int WP = W + 2, R = 128 / WP;
  const int tiles_per_img = (H + R - 1) / R, total = N * tiles_per_img;
  CUtensorMap tmX = make_tmap_nhwc_box_c(in, N, H, W, Cin, Cin, WP, R + 2);
  const int smem = 1024 + C1_STAGES * C1_ROWS * Cin * 2 + 64 + 2 * C1_ROWS * 9 * 4;
  int grid = 2 * sm_count();
  if (grid > total) grid = total;
  if (Cin == 64) {
    static bool configured = false;
    if (!configured) {
      PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_to1_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = true;
    }
    launch_k(conv_to1_kernel<64>, dim3(grid), dim3(C1_THREADS), smem, stream, tmX, w9, bias, out, H, W, WP, R, tiles_per_img, total);
  } else {
    static bool configured = false;
    if (!configured) {
      PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_to1_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = true;
    }
    launch_k(conv_to1_kernel<32>, dim3(grid), dim3(C1_THREADS), smem, stream, tmX, w9, bias, out, H, W, WP, R, tiles_per_img, total);
  }
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// 1-3 input channels -> 32/64 output channels, 3x3, pad 1, stride 1|2.
// K = 9*Cs <= 27 is padded to 16 or 32; the A fragments (im2col rows) are assembled in registers from scalar loads
// of the tiny input (L1/L2 resident), the B fragments (all weights) live in registers, a warp owns 16 consecutive
// output pixels.  The kernel is bound by writing the Cout-channel output: the accumulators are staged through
// shared memory so that every lane stores 16 contiguous bytes of a 128-byte pixel row.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf16_bits(float v) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ uint32_t load_bits(const bf16* p) { return (uint32_t)*reinterpret_cast<const uint16_t*>(p); }
__device__ __forceinline__ uint32_t load_bits(const float* p) { return bf16_bits(*p); }

constexpr int FEW_STAGE_LD = 72;   // floats per staged row: 64 + 8 keeps the float2 writes and float4 reads conflict-free

template <typename TIn, int CS, int COUT, int STRIDE>
__global__ void __launch_bounds__(256, 2)
conv_few_kernel(const TIn* __restrict__ in, const bf16* __restrict__ wnk, const float* __restrict__ bias, int act,
                float slope, const bf16* __restrict__ act_ref, float ref_neg, bf16* __restrict__ out, int H, int W,
                int Ho, int Wo, long long M) {
  pdl_enter();
  constexpr int K = 9 * CS, KS = (K + 15) / 16, NT = COUT / 8, CH = COUT / 8;
  __shared__ __align__(16) float stage[8][16][FEW_STAGE_LD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane & 3, rr = lane >> 2;

  // B fragments: b0 = B[k = ks*16 + 2q + {0,1}][n = j*8 + rr], b1 = same at k + 8;  B[k][n] = wnk[n*K + k]
  uint32_t breg[KS][NT][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        const int k = ks * 16 + 2 * q + hb * 8, n = j * 8 + rr;
        const uint32_t lo = k < K ? load_bits(wnk + n * K + k) : 0u;
        const uint32_t hi = k + 1 < K ? load_bits(wnk + n * K + k + 1) : 0u;
        breg[ks][j][hb] = lo | (hi << 16);
      }
  // this thread's im2col columns: k = ks*16 + 2q + {0, 1, 8, 9} -> (tap, c) -> element offset relative to the window corner
  int kdelta[KS * 4], kdr[KS * 4], kds[KS * 4];
#pragma unroll
  for (int e = 0; e < KS * 4; ++e) {
    const int k = (e >> 2) * 16 + 2 * q + (e & 1) + ((e >> 1) & 1) * 8;
    const int tap = k / CS, c = k - tap * CS;
    kdr[e] = k < K ? tap / 3 : -100000;          // pushes invalid columns out of range
    kds[e] = tap % 3;
    kdelta[e] = ((tap / 3) * W + (tap % 3)) * CS + c;
  }
  float bcol[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    bcol[j][0] = bias != nullptr ? __ldg(bias + j * 8 + 2 * q) : 0.f;
    bcol[j][1] = bias != nullptr ? __ldg(bias + j * 8 + 2 * q + 1) : 0.f;
  }

  const long long nblk = (M + 15) / 16;
  // im2col fragment of one 16-pixel block (two rows per thread: rr and rr + 8)
  auto load_a = [&](long long blk_, uint32_t (&a)[KS][4]) {
    const long long m0_ = blk_ * 16;
    int hb_[2], wb_[2];
    long long base[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      long long m = m0_ + rr + t * 8;
      if (m >= M) m = M - 1;
      const int wo = (int)(m % Wo);
      const long long t2 = m / Wo;
      const int ho = (int)(t2 % Ho);
      const long long n = t2 / Ho;
      hb_[t] = ho * STRIDE - 1;
      wb_[t] = wo * STRIDE - 1;
      base[t] = ((n * H + hb_[t]) * W + wb_[t]) * CS;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) {
        const int t = r4 & 1;                    // a0/a2: row rr, a1/a3: row rr + 8
        const int e0 = ks * 4 + (r4 >> 1) * 2;   // a0/a1: k pair 0, a2/a3: k pair +8
        if (CS == 2 && sizeof(TIn) == 2) {
          // two bf16 channels of one tap: the k pair (2q, 2q + 1) is ONE aligned 32-bit word of the NHWC input
          const int hi = hb_[t] + kdr[e0], wi = wb_[t] + kds[e0];
          const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
          a[ks][r4] = ok ? *reinterpret_cast<const uint32_t*>(in + base[t] + kdelta[e0]) : 0u;
          continue;
        }
        uint32_t bits[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = e0 + u;
          const int hi = hb_[t] + kdr[e], wi = wb_[t] + kds[e];
          const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
          bits[u] = ok ? load_bits(in + base[t] + kdelta[e]) : 0u;
        }
        a[ks][r4] = bits[0] | (bits[1] << 16);
      }
  };
  const long long bstride = (long long)gridDim.x * 8;
  long long blk = (long long)blockIdx.x * 8 + warp;
  uint32_t a[KS][4], a_next[KS][4];
  if (blk < nblk) load_a(blk, a);
  for (; blk < nblk; blk += bstride) {
    const long long m0 = blk * 16;
    // the next block's (L2-latency) loads are in flight while this block is computed and stored
    if (blk + bstride < nblk) load_a(blk + bstride, a_next);
    // the activation-reference rows of THIS block (only their signs are used) are fetched now, under the MMAs
    uint4 aref[16 * CH / 32];
    if (act_ref != nullptr) {
#pragma unroll
      for (int it = 0; it < 16 * CH / 32; ++it) {
        const int id = it * 32 + lane;
        const int row = id / CH, cc = id - row * CH;
        const long long m = m0 + row;
        aref[it] = m < M ? *reinterpret_cast<const uint4*>(act_ref + m * COUT + cc * 8) : make_uint4(0, 0, 0, 0);
      }
    }
    float c[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) mma_bf16_16816(c[j], a[ks], breg[ks][j][0], breg[ks][j][1]);
    }
    // phase 1: bias + activation, stage as fp32
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float v0 = c[j][t * 2] + bcol[j][0], v1 = c[j][t * 2 + 1] + bcol[j][1];
        if (act == ACT_LRELU) { v0 = v0 > 0.f ? v0 : v0 * slope; v1 = v1 > 0.f ? v1 : v1 * slope; }
        else if (act == ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        *reinterpret_cast<float2*>(&stage[warp][rr + t * 8][j * 8 + 2 * q]) = make_float2(v0, v1);
      }
    __syncwarp();
    // phase 2: 16 bytes (8 channels) per lane
#pragma unroll
    for (int it = 0; it < 16 * CH / 32; ++it) {
      const int id = it * 32 + lane;
      const int row = id / CH, cc = id - row * CH;
      const long long m = m0 + row;
      const float4 f0 = *reinterpret_cast<const float4*>(&stage[warp][row][cc * 8]);
      const float4 f1 = *reinterpret_cast<const float4*>(&stage[warp][row][cc * 8 + 4]);
      float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
      if (m < M) {
        if (act_ref != nullptr) {
          const uint4 u = aref[it];
          const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = __uint_as_float(w4[e] << 16), hi = __uint_as_float(w4[e] & 0xffff0000u);
            v[2 * e] *= lo > 0.f ? 1.f : ref_neg;
            v[2 * e + 1] *= hi > 0.f ? 1.f : ref_neg;
          }
        }
        uint4 o;
        o.x = bf16_bits(v[0]) | (bf16_bits(v[1]) << 16);
        o.y = bf16_bits(v[2]) | (bf16_bits(v[3]) << 16);
        o.z = bf16_bits(v[4]) | (bf16_bits(v[5]) << 16);
        o.w = bf16_bits(v[6]) | (bf16_bits(v[7]) << 16);
        *reinterpret_cast<uint4*>(out + m * COUT + cc * 8) = o;
      }
    }
    __syncwarp();
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) a[ks][r4] = a_next[ks][r4];
  }
}

bool conv_few_supported(int Cs, int Cout, int ksize, int stride, int pad) {
  return Cs >= 1 && Cs <= 3 && (Cout == 32 || Cout == 64) && ksize == 3 && pad == 1 && (stride == 1 || stride == 2);
}

template <typename TIn, int CS, int COUT, int STRIDE>
static void launch_few(const TIn* in, int N, int H, int W, const bf16* wnk, const FewEpilogue& e, bf16* out,
                       cudaStream_t stream) {
  const int Ho = (H + 2 - 3) / STRIDE + 1, Wo = (W + 2 - 3) / STRIDE + 1;
  const long long M = (long long)N * Ho * Wo;
  long long blocks = ((M + 15) / 16 + 7) / 8;
  const long long cap = (long long)sm_count() * 6;
  if (blocks > cap) blocks = cap;
  const float ref_neg = e.ref_act == ACT_LRELU ? e.ref_slope : 0.f;
  launch_k(conv_few_kernel<TIn, CS, COUT, STRIDE>, dim3((int)blocks), dim3(256), 0, stream, 
      in, wnk, e.bias, e.act, e.slope, e.ref_act != ACT_NONE ? e.act_ref : nullptr, ref_neg, out, H, W, Ho, Wo, M);
}

template <typename TIn>
void conv_few(const TIn* in, int N, int H, int W, int Cs, const bf16* wnk, int Cout, int stride, const FewEpilogue& epi,
              bf16* out, cudaStream_t stream) {
  PCG_PROFILE("conv_small", stream);
  PCG_REQUIRE(conv_few_supported(Cs, Cout, 3, stride, 1), "conv_few: unsupported geometry");
#define PCG_F(CS_, CO_, ST_) launch_few<TIn, CS_, CO_, ST_>(in, N, H, W, wnk, epi, out, stream)
  const int key = Cs * 1000 + Cout * 10 + stride;
  switch (key) {
    case 1000 + 320 + 1: PCG_F(1, 32, 1); break;
    case 1000 + 640 + 1: PCG_F(1, 64, 1); break;
    case 2000 + 640 + 2: PCG_F(2, 64, 2); break;
    case 3000 + 640 + 1: PCG_F(3, 64, 1); break;
    default: throw Error(1, "conv_few: this (Cs, Cout, stride) combination is not instantiated");
  }
#undef PCG_F
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
template void conv_few<bf16>(const bf16*, int, int, int, int, const bf16*, int, int, const FewEpilogue&, bf16*, cudaStream_t);
template void conv_few<float>(const float*, int, int, int, int, const bf16*, int, int, const FewEpilogue&, bf16*, cudaStream_t);

// ------------------------------------------------------------------------------------------
// Weight gradients of the skinny layers.  Both kernels stream the 64-channel tensor ([M][64] bf16, M = pixels) through
// a TMA ring of 192-row tiles; the 12 consumer warps each contract 16 pixels per tile on the warp-level tensor path
// with fp32 accumulators that live in registers for the whole kernel, are summed in a fixed order inside the block
// and written as one partial row per block; a small second kernel adds the rows (fixed order) and writes torch's
// OIHW layout.
// ------------------------------------------------------------------------------------------
constexpr int WG_TILE = 192, WG_STAGES = 4, WG_CONSUMERS = 12, WG_THREADS = 32 * (WG_CONSUMERS + 1);
constexpr int WG_SMEM = 1024 + WG_STAGES * WG_TILE * 128 + 64 + 32 * 64 * 4;

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

struct WgRing {
  uint8_t* tiles;
  uint64_t *full, *empty;
  float* red;
};
__device__ __forceinline__ WgRing wg_setup(uint8_t* smem_raw, const CUtensorMap* tm) {
  const uint32_t raw_addr = smem_u32(smem_raw);
  WgRing r;
  r.tiles = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  r.full = reinterpret_cast<uint64_t*>(r.tiles + WG_STAGES * WG_TILE * 128);
  r.empty = r.full + WG_STAGES;
  r.red = reinterpret_cast<float*>(r.empty + WG_STAGES);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(tm);
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&r.full[s], 1); mbar_init(&r.empty[s], WG_CONSUMERS); }
    fence_barrier_init();
  }
  __syncthreads();
  return r;
}
// producer warp: one 192-row box per tile (rows past M arrive as zeros)
__device__ __forceinline__ void wg_produce(const WgRing& r, const CUtensorMap* tm, int ntiles) {
  if ((threadIdx.x & 31) != 0) return;
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    mbar_wait(&r.empty[stage], phase ^ 1);
    mbar_expect_tx(&r.full[stage], WG_TILE * 128);
    tma_load_2d(tm, &r.full[stage], r.tiles + stage * WG_TILE * 128, 0, tile * WG_TILE);
    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
  }
}

__device__ __forceinline__ void decode_pixel(long long m, int Ho, int Wo, int& n, int& ho, int& wo) {
  wo = (int)(m % Wo);
  const long long t = m / Wo;
  ho = (int)(t % Ho);
  n = (int)(t / Ho);
}

template <typename TIn, int CS, int STRIDE>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_few_kernel(const __grid_constant__ CUtensorMap tmDY, const TIn* __restrict__ in, int H, int W, int Ho, int Wo,
                 long long M, int ntiles, float* __restrict__ part) {
  pdl_enter();
  constexpr int K = 9 * CS;                       // rows 0..K-1 of D: (tap, c); row K: the bias gradient; K + 1 <= 32
  extern __shared__ uint8_t smem_raw[];
  const WgRing ring = wg_setup(smem_raw, &tmDY);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == WG_CONSUMERS) { wg_produce(ring, &tmDY, ntiles); return; }
  const int q = lane & 3, rr = lane >> 2;

  // this thread's four D rows: k = rr + 8*i  -> window offsets
  int kdelta[4], kdr[4], kds[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = rr + 8 * i;
    const int tap = k / CS, c = k - tap * CS;
    kdr[i] = k < K ? tap / 3 : (k == K ? 100000 : -100000);   // +100000 marks the constant-1 column
    kds[i] = tap % 3;
    kdelta[i] = ((tap / 3) * W + (tap % 3)) * CS + c;
  }
  float acc[2][8][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[a][j][0] = acc[a][j][1] = acc[a][j][2] = acc[a][j][3] = 0.f;

  int stage = 0;
  uint32_t phase = 0;
  const uint32_t one = 0x3f80u;                   // bf16 1.0
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long m0 = (long long)tile * WG_TILE + warp * 16;
    // A^T fragments: element (k, m) for m = m0 + 2q + {0, 1, 8, 9}
    uint32_t bits[4][4];                          // [k index i][pixel index u]
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long m = m0 + 2 * q + (u & 1) + (u >> 1) * 8;
      const bool mok = m < M;
      int n, ho, wo;
      decode_pixel(mok ? m : 0, Ho, Wo, n, ho, wo);
      const int hb = ho * STRIDE - 1, wb = wo * STRIDE - 1;
      const long long base = (((long long)n * H + hb) * W + wb) * CS;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int hi = hb + kdr[i], wi = wb + kds[i];
        const bool ok = mok && hi >= 0 && hi < H && wi >= 0 && wi < W;
        uint32_t v = ok ? load_bits(in + base + kdelta[i]) : 0u;
        if (kdr[i] == 100000 && mok) v = one;
        bits[i][u] = v;
      }
    }
    uint32_t a[2][4];
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      a[kb][0] = bits[2 * kb][0] | (bits[2 * kb][1] << 16);          // row k = kb*16 + rr,     m = 2q, 2q+1
      a[kb][1] = bits[2 * kb + 1][0] | (bits[2 * kb + 1][1] << 16);  // row k = kb*16 + rr + 8
      a[kb][2] = bits[2 * kb][2] | (bits[2 * kb][3] << 16);          // m = 2q+8, 2q+9
      a[kb][3] = bits[2 * kb + 1][2] | (bits[2 * kb + 1][3] << 16);
    }
    mbar_wait(&ring.full[stage], phase);
    const uint32_t tbase = smem_u32(ring.tiles + stage * WG_TILE * 128);
    const int j8 = lane >> 3;                     // matrix index supplied by this lane
    const uint32_t rowoff = (uint32_t)(warp * 16 + (j8 & 1) * 8 + (lane & 7)) * 128u;
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldmatrix_x4_trans(tbase + swz128(rowoff + (uint32_t)(np * 2 + (j8 >> 1)) * 16u), b);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        mma_bf16_16816(acc[kb][2 * np], a[kb], b[0], b[1]);
        mma_bf16_16816(acc[kb][2 * np + 1], a[kb], b[2], b[3]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ring.empty[stage]);
    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
  }
  // fixed-order sum of the 12 warps' fragments into red[32][64]
  for (int w = 0; w < WG_CONSUMERS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            float2* p = reinterpret_cast<float2*>(&ring.red[(kb * 16 + rr + t * 8) * 64 + j * 8 + 2 * q]);
            float2 v = make_float2(acc[kb][j][2 * t], acc[kb][j][2 * t + 1]);
            if (w > 0) { const float2 o = *p; v.x += o.x; v.y += o.y; }
            *p = v;
          }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(WG_CONSUMERS * 32) : "memory");
  }
  for (int i = threadIdx.x; i < 32 * 64; i += WG_CONSUMERS * 32) part[(size_t)blockIdx.x * 2048 + i] = ring.red[i];
}

// dw[co][c][tap] = sum_b part[b][(tap*Cs + c)*64 + co];  db[co] = sum_b part[b][K*64 + co]
__global__ void wgrad_few_reduce_kernel(const float* __restrict__ part, int nparts, int Cs, float* __restrict__ dw,
                                        float* __restrict__ db) {
  pdl_enter();
  const int K = 9 * Cs;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (K + 1) * 64) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int b = 0;
  for (; b + 3 < nparts; b += 4) {
    s0 += part[(size_t)b * 2048 + i];
    s1 += part[(size_t)(b + 1) * 2048 + i];
    s2 += part[(size_t)(b + 2) * 2048 + i];
    s3 += part[(size_t)(b + 3) * 2048 + i];
  }
  for (; b < nparts; ++b) s0 += part[(size_t)b * 2048 + i];
  const float s = (s0 + s1) + (s2 + s3);
  const int k = i >> 6, co = i & 63;
  if (k == K) {
    if (db != nullptr) db[co] = s;
  } else {
    const int tap = k / Cs, c = k - tap * Cs;
    dw[((size_t)co * Cs + c) * 9 + tap] = s;
  }
}

int wgrad_few_parts() { return sm_count(); }
bool wgrad_few_supported(int Cs, int Cout, int ksize, int stride, int pad) {
  return Cs >= 1 && Cs <= 3 && Cout == 64 && ksize == 3 && pad == 1 && (stride == 1 || stride == 2);
}

template <typename TIn, int CS, int STRIDE>
static void launch_wgrad_few(const TIn* in, const bf16* dy, int N, int H, int W, float* part, int grid, cudaStream_t stream) {
  const int Ho = (H + 2 - 3) / STRIDE + 1, Wo = (W + 2 - 3) / STRIDE + 1;
  const long long M = (long long)N * Ho * Wo;
  const int ntiles = (int)((M + WG_TILE - 1) / WG_TILE);
  CUtensorMap tm = make_tmap_2d(dy, (uint64_t)M, 64, WG_TILE);
  static bool configured = false;
  if (!configured) {
    PCG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_few_kernel<TIn, CS, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    configured = true;
  }
  launch_k(wgrad_few_kernel<TIn, CS, STRIDE>, dim3(grid), dim3(WG_THREADS), WG_SMEM, stream, tm, in, H, W, Ho, Wo, M, ntiles, part);
}

template <typename TIn>
void wgrad_few(const TIn* in, const bf16* dy, int N, int H, int W, int Cs, int stride, float* part, float* dw, float* db,
               cudaStream_t stream) {
  PCG_PROFILE("wgrad_small", stream);
  PCG_REQUIRE(wgrad_few_supported(Cs, 64, 3, stride, 1), "wgrad_few: unsupported geometry");
  const int grid = wgrad_few_parts();
  const int key = Cs * 10 + stride;
  switch (key) {
    case 31: launch_wgrad_few<TIn, 3, 1>(in, dy, N, H, W, part, grid, stream); break;
    case 22: launch_wgrad_few<TIn, 2, 2>(in, dy, N, H, W, part, grid, stream); break;
    default: throw Error(1, "wgrad_few: this (Cs, stride) combination is not instantiated");
  }
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
  launch_k(wgrad_few_reduce_kernel, dim3(cdiv((9 * Cs + 1) * 64, 128)), dim3(128), 0, stream, part, grid, Cs, dw, db);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
template void wgrad_few<bf16>(const bf16*, const bf16*, int, int, int, int, int, float*, float*, float*, cudaStream_t);

// ---- 64 -> 1: D[ci][tap] = sum_q x[q][ci] * G[q][tap],  G[q][tap = (r, s)] = g[n][h - (r-1)][w - (s-1)] (0 outside)
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_to1_kernel(const __grid_constant__ CUtensorMap tmX, const bf16* __restrict__ g, int H, int W, long long M,
                 int ntiles, float* __restrict__ part) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const WgRing ring = wg_setup(smem_raw, &tmX);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == WG_CONSUMERS) { wg_produce(ring, &tmX, ntiles); return; }
  const int q = lane & 3, rr = lane >> 2;
  // this thread's G columns: tap = rr (n-tile 0) and tap = 8 when rr == 0 (n-tile 1)
  const int dr0 = rr / 3 - 1, ds0 = rr % 3 - 1;
  float acc[4][2][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[a][j][0] = acc[a][j][1] = acc[a][j][2] = acc[a][j][3] = 0.f;
  float gsum = 0.f;                               // lanes with rr == 4 (centre tap) see every g[q] exactly once

  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long m0 = (long long)tile * WG_TILE + warp * 16;
    uint32_t gb[2][4];                            // [n-tile][pixel u]
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long m = m0 + 2 * q + (u & 1) + (u >> 1) * 8;
      const bool mok = m < M;
      int n, h, w;
      decode_pixel(mok ? m : 0, H, W, n, h, w);
      {
        const int hh = h - dr0, ww = w - ds0;
        const bool ok = mok && hh >= 0 && hh < H && ww >= 0 && ww < W;
        gb[0][u] = ok ? load_bits(g + ((long long)n * H + hh) * W + ww) : 0u;
      }
      {
        const int hh = h - 1, ww = w - 1;         // tap 8 = (r, s) = (2, 2)
        const bool ok = mok && rr == 0 && hh >= 0 && ww >= 0;
        gb[1][u] = ok ? load_bits(g + ((long long)n * H + hh) * W + ww) : 0u;
      }
      if (rr == 4) gsum += __uint_as_float(gb[0][u] << 16);
    }
    uint32_t b[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      b[j][0] = gb[j][0] | (gb[j][1] << 16);
      b[j][1] = gb[j][2] | (gb[j][3] << 16);
    }
    mbar_wait(&ring.full[stage], phase);
    const uint32_t tbase = smem_u32(ring.tiles + stage * WG_TILE * 128);
    const int j8 = lane >> 3;
    // A = x^T: matrix j8 -> pixel rows (j8 >> 1)*8 + lane%8, channel chunk mt*2 + (j8 & 1)
    const uint32_t rowoff = (uint32_t)(warp * 16 + (j8 >> 1) * 8 + (lane & 7)) * 128u;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      uint32_t a[4];
      ldmatrix_x4_trans(tbase + swz128(rowoff + (uint32_t)(mt * 2 + (j8 & 1)) * 16u), a);
      mma_bf16_16816(acc[mt][0], a, b[0][0], b[0][1]);
      mma_bf16_16816(acc[mt][1], a, b[1][0], b[1][1]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ring.empty[stage]);
    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
  }
  // red[64 ci][16 taps] + red[1024 + warp*32 + lane] for the bias-gradient partials
  for (int w = 0; w < WG_CONSUMERS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            float2* p = reinterpret_cast<float2*>(&ring.red[(mt * 16 + rr + t * 8) * 16 + j * 8 + 2 * q]);
            float2 v = make_float2(acc[mt][j][2 * t], acc[mt][j][2 * t + 1]);
            if (w > 0) { const float2 o = *p; v.x += o.x; v.y += o.y; }
            *p = v;
          }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(WG_CONSUMERS * 32) : "memory");
  }
  ring.red[1024 + warp * 32 + lane] = gsum;
  asm volatile("bar.sync 1, %0;" ::"n"(WG_CONSUMERS * 32) : "memory");
  for (int i = threadIdx.x; i < 1024; i += WG_CONSUMERS * 32) part[(size_t)blockIdx.x * 1088 + i] = ring.red[i];
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < WG_CONSUMERS * 32; ++i) s += ring.red[1024 + i];
    part[(size_t)blockIdx.x * 1088 + 1024] = s;
  }
}
__global__ void wgrad_to1_reduce_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dw,
                                        float* __restrict__ db) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > 1024) return;
  float s0 = 0.f, s1 = 0.f;
  int b = 0;
  for (; b + 1 < nparts; b += 2) {
    s0 += part[(size_t)b * 1088 + i];
    s1 += part[(size_t)(b + 1) * 1088 + i];
  }
  if (b < nparts) s0 += part[(size_t)b * 1088 + i];
  const float s = s0 + s1;
  if (i == 1024) {
    if (db != nullptr) db[0] = s;
  } else {
    const int ci = i >> 4, tap = i & 15;
    if (tap < 9) dw[ci * 9 + tap] = s;
  }
}
int wgrad_to1_parts() { return sm_count(); }
void wgrad_to1(const bf16* x, const bf16* g, int N, int H, int W, float* part, float* dw, float* db, cudaStream_t stream) {
  PCG_PROFILE("wgrad_small", stream);
  const long long M = (long long)N * H * W;
  const int ntiles = (int)((M + WG_TILE - 1) / WG_TILE);
  CUtensorMap tm = make_tmap_2d(x, (uint64_t)M, 64, WG_TILE);
  static bool configured = false;
  if (!configured) {
    PCG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_to1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    configured = true;
  }
  const int grid = wgrad_to1_parts();
  launch_k(wgrad_to1_kernel, dim3(grid), dim3(WG_THREADS), WG_SMEM, stream, tm, g, H, W, M, ntiles, part);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
  launch_k(wgrad_to1_reduce_kernel, dim3(cdiv(1025, 128)), dim3(128), 0, stream, part, grid, dw, db);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// Stride-2 data gradient towards one input channel (discriminator conv0 -> the image / the label map).
// ------------------------------------------------------------------------------------------
constexpr int S2_ROWS = 208, S2_STAGES = 3, S2_THREADS = 288;
constexpr int S2_SMEM = 1024 + S2_STAGES * S2_ROWS * 128 + 64 + S2_ROWS * 9 * 4;

__global__ void __launch_bounds__(S2_THREADS, 2)
dgrad_s2_to1_kernel(const __grid_constant__ CUtensorMap tmDY, const bf16* __restrict__ wrot, float* __restrict__ dx,
                    int N, int H, int W, int Ho, int Wo) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S2_STAGES * S2_ROWS * 128);
  uint64_t* empty = full + S2_STAGES;
  float* P = reinterpret_cast<float*>(empty + S2_STAGES + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npos = Ho * Wo;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < S2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < S2_STAGES * S2_ROWS * 128 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  __syncthreads();

  if (warp == 8) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int img = blockIdx.x; img < N; img += gridDim.x) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], npos * 128);
        tma_load_2d(&tmDY, &full[stage], smem + stage * S2_ROWS * 128, 0, img * npos);
        if (++stage == S2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    return;
  }
  const int q = lane & 3, rr = lane >> 2;
  // B[k = co][n = tap] = wrot[8 - tap][co]
  uint32_t breg[4][2][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      const int co = ks * 16 + 2 * q + hb * 8;
      breg[ks][0][hb] = *reinterpret_cast<const uint32_t*>(wrot + (8 - rr) * 64 + co);
      breg[ks][1][hb] = rr == 0 ? *reinterpret_cast<const uint32_t*>(wrot + co) : 0u;     // tap 8 -> row 0
    }
  const int nblk = (npos + 15) / 16;
  int stage = 0;
  uint32_t phase = 0;
  for (int img = blockIdx.x; img < N; img += gridDim.x) {
    mbar_wait(&full[stage], phase);
    const uint32_t base = smem_u32(smem + stage * S2_ROWS * 128);
    for (int blk = warp; blk < nblk; blk += 8) {
      float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      const uint32_t rowoff = (uint32_t)(blk * 16 + (lane & 15)) * 128u;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
        ldmatrix_x4(base + swz128(rowoff + (uint32_t)(ks * 2 + (lane >> 4)) * 16u), a);
        mma_bf16_16816(c[0], a, breg[ks][0][0], breg[ks][0][1]);
        mma_bf16_16816(c[1], a, breg[ks][1][0], breg[ks][1][1]);
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float* row = P + (blk * 16 + rr + t * 8) * 9;
        row[2 * q] = c[0][2 * t];
        row[2 * q + 1] = c[0][2 * t + 1];
        if (q == 0) row[8] = c[1][2 * t];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float* o = dx + (size_t)img * H * W;
    for (int i = threadIdx.x; i < H * W; i += 256) {
      const int hi = i / W, wi = i - hi * W;
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int th = hi + 1 - r;
        if (th < 0 || (th & 1)) continue;
        const int ho = th >> 1;
        if (ho >= Ho) continue;
#pragma unroll
        for (int sx = 0; sx < 3; ++sx) {
          const int tw = wi + 1 - sx;
          if (tw < 0 || (tw & 1)) continue;
          const int wo = tw >> 1;
          if (wo >= Wo) continue;
          s += P[(ho * Wo + wo) * 9 + r * 3 + sx];
        }
      }
      o[i] = s;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (++stage == S2_STAGES) { stage = 0; phase ^= 1; }
  }
}

bool dgrad_s2_to1_supported(int H, int W, int Cout) {
  return Cout == 64 && H % 2 == 0 && W % 2 == 0 && (H / 2) * (W / 2) <= 196 && (H / 2) * (W / 2) >= 1;
}
void dgrad_s2_to1(const bf16* dy, int N, int H, int W, const bf16* wrot, float* dx, cudaStream_t stream) {
  PCG_PROFILE("conv_small", stream);
  PCG_REQUIRE(dgrad_s2_to1_supported(H, W, 64), "dgrad_s2_to1: unsupported geometry");
  const int Ho = H / 2, Wo = W / 2;
  CUtensorMap tm = make_tmap_2d(dy, (uint64_t)N * Ho * Wo, 64, Ho * Wo);
  static bool configured = false;
  if (!configured) {
    PCG_CHECK_CUDA(cudaFuncSetAttribute(dgrad_s2_to1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM));
    configured = true;
  }
  int grid = 2 * sm_count();
  if (grid > N) grid = N;
  launch_k(dgrad_s2_to1_kernel, dim3(grid), dim3(S2_THREADS), S2_SMEM, stream, tm, wrot, dx, N, H, W, Ho, Wo);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

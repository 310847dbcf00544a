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
// row, hardware swizzle); output position i = hh*(W+2) + ww reads tile row i + r*(W+2) + s for tap (r, s), so every
// tap's A fragment is an ldmatrix at a shifted row of the same tile.  The single output channel occupies column 0
// of the m16n8k16 B fragment (lanes 0-3 hold the weights, the rest zeros); the 128 positions of a tile are the
// 8 consumer warps' m16 blocks.
// ------------------------------------------------------------------------------------------
constexpr int C1_STAGES = 3;
constexpr int C1_ROWS = 192;                 // >= 127 + 2*(W+2) + 2 + 1 for W <= 28
constexpr int C1_THREADS = 288;              // warp 0: TMA producer, warps 1-8: consumers

template <int CIN>
__device__ __forceinline__ uint32_t c1_swizzle(uint32_t off) {
  return CIN == 64 ? (off ^ (((off >> 7) & 7u) << 4)) : (off ^ (((off >> 7) & 3u) << 4));
}

template <int CIN>
__global__ void __launch_bounds__(C1_THREADS, 2)
conv_to1_kernel(const __grid_constant__ CUtensorMap tmX, const bf16* __restrict__ w9, const float* __restrict__ bias,
                float* __restrict__ out, int H, int W, int WP, int R, int tiles_per_img, int total_tiles) {
  constexpr int ROWB = CIN * 2, KQ = CIN / 16, STAGE_BYTES = C1_ROWS * ROWB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C1_STAGES * STAGE_BYTES);
  uint64_t* empty = full + C1_STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < C1_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < C1_STAGES * STAGE_BYTES / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  __syncthreads();
  const int box_bytes = (R + 2) * WP * ROWB;

  if (warp == 0) {
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

  // B fragments of all 9*KQ k-steps: b0 = W[k = 2*(lane%4) + {0,1}][n = lane/4], b1 = same at k + 8; only n = 0 is real
  uint32_t breg[9 * KQ][2];
  {
    const int kb = (lane & 3) * 2;
    const bool real = (lane >> 2) == 0;
#pragma unroll
    for (int ks = 0; ks < 9 * KQ; ++ks) {
      const bf16* p = w9 + ks * 16 + kb;            // (tap, kq) -> tap*CIN + kq*16 == ks*16
      breg[ks][0] = real ? *reinterpret_cast<const uint32_t*>(p) : 0u;
      breg[ks][1] = real ? *reinterpret_cast<const uint32_t*>(p + 8) : 0u;
    }
  }
  const float bv = bias != nullptr ? __ldg(bias) : 0.f;
  const int i0 = (warp - 1) * 16;
  const int arow = i0 + (lane & 15), ahalf = lane >> 4;
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / tiles_per_img, h0 = (tile % tiles_per_img) * R;
    mbar_wait(&full[stage], phase);
    const uint32_t base = smem_u32(smem + stage * STAGE_BYTES);
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const uint32_t rowoff = (uint32_t)(arow + (tap / 3) * WP + (tap % 3)) * ROWB;
#pragma unroll
      for (int kq = 0; kq < KQ; ++kq) {
        uint32_t a[4];
        ldmatrix_x4(base + c1_swizzle<CIN>(rowoff + (uint32_t)(kq * 2 + ahalf) * 16u), a);
        mma_bf16_16816(c, a, breg[tap * KQ + kq][0], breg[tap * KQ + kq][1]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if ((lane & 3) == 0) {
#pragma unroll
      for (int hsel = 0; hsel < 2; ++hsel) {
        const int row = i0 + (lane >> 2) + hsel * 8;
        const int hh = row / WP, ww = row - hh * WP;
        if (hh < R && ww < W && h0 + hh < H)
          out[((long long)n * H + h0 + hh) * W + ww] = c[hsel * 2] + bv;
      }
    }
    if (++stage == C1_STAGES) { stage = 0; phase ^= 1; }
  }
}

bool conv_to1_supported(int H, int W, int Cin) {
  const int WP = W + 2;
  if (WP > 128 || (Cin != 32 && Cin != 64)) return false;
  const int R = 128 / WP;
  return R >= 1 && 127 + 2 * WP + 3 <= C1_ROWS && R + 2 <= 256;
}

void conv_to1(const bf16* in, int N, int H, int W, int Cin, const bf16* w9, const float* bias, float* out,
              cudaStream_t stream) {
  PCG_PROFILE("conv_small", stream);
  PCG_REQUIRE(conv_to1_supported(H, W, Cin), "conv_to1: unsupported geometry");
  const int WP = W + 2, R = 128 / WP;
  const int tiles_per_img = (H + R - 1) / R, total = N * tiles_per_img;
  CUtensorMap tmX = make_tmap_nhwc_box_c(in, N, H, W, Cin, Cin, WP, R + 2);
  const int smem = 1024 + C1_STAGES * C1_ROWS * Cin * 2 + 64;
  int grid = 2 * sm_count();
  if (grid > total) grid = total;
  if (Cin == 64) {
    static bool configured = false;
    if (!configured) {
      PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_to1_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = true;
    }
    conv_to1_kernel<64><<<grid, C1_THREADS, smem, stream>>>(tmX, w9, bias, out, H, W, WP, R, tiles_per_img, total);
  } else {
    static bool configured = false;
    if (!configured) {
      PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_to1_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = true;
    }
    conv_to1_kernel<32><<<grid, C1_THREADS, smem, stream>>>(tmX, w9, bias, out, H, W, WP, R, tiles_per_img, total);
  }
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

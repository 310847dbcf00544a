#include <cstdlib>
// Tensor-core convolution kernels for sm_100a.
//
//  * conv_tc_fprop_kernel<BN>: persistent, warp-specialised implicit GEMM
//        Y[M = N*Ho*Wo][Cout] = im2col(X)[M][taps*Cin] * W[Cout][taps*Cin]^T
//    - A operand: NHWC bf16 activations fetched by TMA in *im2col mode* (128 pixels x 64 channels
//      per filter tap, zero fill at the image border, 128B swizzle) -> no im2col buffer in HBM.
//    - B operand: packed bf16 weights fetched by tiled TMA (BN rows x 64 k, 128B swizzle).
//    - tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) issued by one thread, fp32
//      accumulators double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of i+1.
//    - epilogue warps: tcgen05.ld -> +bias -> activation -> +residual -> bf16 store, plus the
//      per-channel sum / sum-of-squares partials train-mode BatchNorm needs (reference:
//      conditional_counteRGAN/mnist/models/generator.py:11-15,17-22).
//  * conv_tc_wgrad64_kernel: dW = X^T * dY with K = pixels; both operands MN-major straight from
//    the NHWC tensors (TMA im2col for the shifted X taps), two taps per M=128 MMA, five fp32
//    accumulators resident in TMEM for the whole kernel, per-CTA partials reduced deterministically.
//
// Reference semantics being replaced: torch.nn.Conv2d forward / ConvolutionBackward0 as called by
// conditional_counteRGAN/mnist/models/generator.py:11,14,49.
#include "conv_tc.cuh"
#include "tc_common.cuh"

#include <cudaTypedefs.h>

#include <mutex>

namespace pcg {
using namespace tc;

// ------------------------------------------------------------------------------------------
// Tensor-map creation (driver entry points resolved at run time; libcuda is not linked)
// ------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
static PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;

static void load_driver_entry_points() {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn);
  });
  if (!g_encode_tiled || !g_encode_im2col)
    throw Error(3, "cuTensorMapEncodeTiled/Im2col driver entry points unavailable");
}

// [rows][cols] bf16 row-major matrix, box = box_rows x 64 columns, 128B swizzle.
CUtensorMap make_tmap_2d(const bf16* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  load_driver_entry_points();
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(bf16)};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(3, "cuTensorMapEncodeTiled failed: " + std::to_string(int(r)));
  return m;
}

// NHWC bf16 activation, im2col mode: 128 pixels x 64 channels per box.
// The base pixel walks the box [lower, dim + upper) in steps of `stride`; the element fetched for a
// filter tap is base + tap offset, zero-filled outside the tensor.
CUtensorMap make_tmap_im2col_box(const bf16* base, int N, int H, int W, int C, int lower_w, int lower_h,
                                        int upper_w, int upper_h, int stride) {
  load_driver_entry_points();
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims,
                               strides, lower, upper, /*channelsPerPixel=*/64,
                               /*pixelsPerColumn=*/128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(3, "cuTensorMapEncodeIm2col failed: " + std::to_string(int(r)));
  return m;
}
// NHWC bf16 activation, tiled mode: box = 64 channels x box_w x box_h x 1 image, 128B swizzle, zero OOB fill.
CUtensorMap make_tmap_nhwc_box(const bf16* base, int N, int H, int W, int C, int box_w, int box_h) {
  load_driver_entry_points();
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(3, "cuTensorMapEncodeTiled(4d) failed: " + std::to_string(int(r)));
  return m;
}
// Row-class view of an NHWC tensor with H % 4 == 0: dims (C, W, class = h % 4, idx = h / 4, N); box = 64 channels x
// box_w x one class x box_rows consecutive rows of that class x 1 image (conv_tc64s_fprop_kernel).
CUtensorMap make_tmap_nhwc_rowclass(const bf16* base, int N, int H, int W, int C, int box_w, int box_rows) {
  load_driver_entry_points();
  PCG_REQUIRE(H % 4 == 0, "row-class tensor map: H must be a multiple of 4");
  CUtensorMap m;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, 4, (cuuint64_t)(H / 4), (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)4 * W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)box_w, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<bf16*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(3, "cuTensorMapEncodeTiled(5d row classes) failed: " + std::to_string(int(r)));
  return m;
}
// Same with `box_c` (16, 32 or 64) channels per pixel: 32/64/128-byte rows, matching swizzle.
CUtensorMap make_tmap_nhwc_box_c(const bf16* base, int N, int H, int W, int C, int box_c, int box_w, int box_h) {
  load_driver_entry_points();
  PCG_REQUIRE(box_c == 16 || box_c == 32 || box_c == 64, "box_c must be 16, 32 or 64");
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = g_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(3, "cuTensorMapEncodeTiled(4d, box_c) failed: " + std::to_string(int(r)));
  return m;
}
// Convolution window: lower = -pad, upper = pad - (k-1)  (dilation 1).
static CUtensorMap make_tmap_im2col(const bf16* base, int N, int H, int W, int C, int ksize, int stride, int pad) {
  return make_tmap_im2col_box(base, N, H, W, C, -pad, -pad, pad - (ksize - 1), pad - (ksize - 1), stride);
}

// ------------------------------------------------------------------------------------------
// fprop / dgrad implicit GEMM
// ------------------------------------------------------------------------------------------
constexpr int TILE_M = 128;
constexpr int A_STAGE_BYTES = TILE_M * 128;       // 128 pixels x 64 bf16
constexpr int PIPE_BYTES = 196608;                // operand ring budget
constexpr int FPROP_THREADS = 192;                // warp0 TMA, warp1 MMA, warps 2-5 epilogue

struct FpropParams {
  int M, Ho, Wo;                 // GEMM rows = N*Ho*Wo base pixels
  int tstride, lower_h, lower_w; // base coordinate of output (ho, wo) = ho*tstride + lower_h, ...
  int taps_h, taps_w, cin_blocks, Cout;
  // output row m = (n, ho, wo) is stored at pixel (n*OH + ho*os + oh0)*OW + wo*os + ow0 of the output tensor
  int OH, OW, os, oh0, ow0;
  int num_m_tiles, num_n_tiles;
  const float* bias;
  int act;
  float slope;
  const bf16* add_src;
  const bf16* act_ref;
  int ref_act;
  float ref_slope;
  bf16* out;
  float* out_f32;                // if set, results are stored here as fp32 instead of bf16 into `out`
  float* stats;
};

// One launch can carry NC independent implicit GEMMs that share the tile shape (the four parity classes of a stride-2
// data gradient): tiles [tile_end[c-1], tile_end[c]) belong to problem c.
template <int NC>
struct FpropArgs {
  CUtensorMap a[NC], b[NC];
  FpropParams c[NC];
  int tile_end[NC];
};

template <int BN>
struct FpropCfg {
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = PIPE_BYTES / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int STATS_BYTES = 4 * 2 * BN * 4;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STATS_BYTES + 256;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Sum over the 32 lanes of a warp of 32 per-lane values; lane l ends up with column l's total.
__device__ __forceinline__ float butterfly_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, cnt = 32; off >= 1; off >>= 1, cnt >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < cnt / 2; ++j) {
      float send = upper ? v[j] : v[j + cnt / 2];
      float keep = upper ? v[j + cnt / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int BN, int NC>
__global__ void __launch_bounds__(FPROP_THREADS, 1)
conv_tc_fprop_kernel(const __grid_constant__ FpropArgs<NC> P) {
  pdl_launch_dependents();                        // prologue first (no global memory), pdl_wait() below
  using Cfg = FpropCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* ring = smem;
  float* stats_smem = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::STATS_BYTES);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]
  uint64_t* tfull = bars + 2 * Cfg::STAGES;    // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = P.tile_end[NC - 1];
  // problem of a tile and the tile's index inside it
  auto locate = [&](int tile, int& lt) {
    int c = 0;
#pragma unroll
    for (int k = 0; k + 1 < NC; ++k) c += tile >= P.tile_end[k] ? 1 : 0;
    lt = tile - (c > 0 ? P.tile_end[c > 0 ? c - 1 : 0] : 0);
    return c;
  };

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      tma_prefetch_desc(&P.a[k]);
      tma_prefetch_desc(&P.b[k]);
    }
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                     // the predecessor kernel has completed and flushed

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int lt;
        const int cls = locate(tile, lt);
        const FpropParams& p = P.c[cls];
        const CUtensorMap& tmA = P.a[cls];
        const CUtensorMap& tmB = P.b[cls];
        const int hw = p.Ho * p.Wo, taps = p.taps_h * p.taps_w;
        const int m_tile = lt / p.num_n_tiles, n_tile = lt % p.num_n_tiles;
        const int p0 = m_tile * TILE_M;
        const int n_img = p0 / hw, rem = p0 % hw;
        const int cw = (rem % p.Wo) * p.tstride + p.lower_w;
        const int ch = (rem / p.Wo) * p.tstride + p.lower_h;
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / p.taps_w, s = tap % p.taps_w;
          for (int cb = 0; cb < p.cin_blocks; ++cb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
            uint8_t* sa = ring + stage * Cfg::STAGE_BYTES;
            tma_load_im2col_4d(&tmA, &full[stage], sa, cb * 64, cw, ch, n_img, (uint16_t)s, (uint16_t)r);
            tma_load_2d(&tmB, &full[stage], sa + A_STAGE_BYTES, (tap * p.cin_blocks + cb) * 64,
                        n_tile * BN);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int lt;
        const FpropParams& p = P.c[locate(tile, lt)];
        const int num_kb = p.taps_h * p.taps_w * p.cin_blocks;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(ring + stage * Cfg::STAGE_BYTES);
          const uint64_t da = umma_smem_desc(a_base, 16, 1024);
          const uint64_t db = umma_smem_desc(a_base + A_STAGE_BYTES, 16, 1024);
          // 4 x (K = 16 bf16 = 32 B) inside the 128 B swizzle row: advance the start-address field by 2
          if (kb == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, 1u);
          }
          umma_commit(&empty[stage]);       // frees the smem slot once these MMAs have read it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);           // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    float acc_s[BN / 32], acc_q[BN / 32];
#pragma unroll
    for (int i = 0; i < BN / 32; ++i) acc_s[i] = acc_q[i] = 0.f;

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int lt;
      const FpropParams& p = P.c[locate(tile, lt)];
      const int m_tile = lt / p.num_n_tiles, n_tile = lt % p.num_n_tiles;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const long long mrow = (long long)m_tile * TILE_M + row;
      const bool valid = mrow < p.M;
      long long pix = mrow;
      if (p.os != 1 || p.OH != p.Ho || p.OW != p.Wo) {
        const int hw = p.Ho * p.Wo;
        const int n_img = (int)(mrow / hw), rem = (int)(mrow % hw);
        pix = ((long long)n_img * p.OH + (rem / p.Wo) * p.os + p.oh0) * p.OW + (rem % p.Wo) * p.os + p.ow0;
      }
#pragma unroll
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + acc * BN + chunk * 32, r);
        tmem_ld_wait();
        const int col0 = n_tile * BN + chunk * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + col0 + j);
        }
        if (p.act == ACT_LRELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
        } else if (p.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.add_src != nullptr && valid) {
          const uint4* src = reinterpret_cast<const uint4*>(p.add_src + pix * p.Cout + col0);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            uint4 u = __ldg(src + j4);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 f = __bfloat1622float2(h[e]);
              v[j4 * 8 + e * 2] += f.x;
              v[j4 * 8 + e * 2 + 1] += f.y;
            }
          }
        }
        if (p.act_ref != nullptr && valid) {
          const uint4* src = reinterpret_cast<const uint4*>(p.act_ref + pix * p.Cout + col0);
          const float neg = p.ref_act == ACT_LRELU ? p.ref_slope : 0.f;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            uint4 u = __ldg(src + j4);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 f = __bfloat1622float2(h[e]);
              v[j4 * 8 + e * 2] *= (f.x > 0.f ? 1.f : neg);
              v[j4 * 8 + e * 2 + 1] *= (f.y > 0.f ? 1.f : neg);
            }
          }
        }
        if (valid && p.out_f32 != nullptr) {          // fp32 result (fp32-storage callers, conv_auto.cu)
          float4* dst = reinterpret_cast<float4*>(p.out_f32 + pix * p.Cout + col0);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) dst[j4] = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
        } else if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + pix * p.Cout + col0);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            uint4 u;
            u.x = pack_bf16x2(v[j4 * 8 + 0], v[j4 * 8 + 1]);
            u.y = pack_bf16x2(v[j4 * 8 + 2], v[j4 * 8 + 3]);
            u.z = pack_bf16x2(v[j4 * 8 + 4], v[j4 * 8 + 5]);
            u.w = pack_bf16x2(v[j4 * 8 + 6], v[j4 * 8 + 7]);
            dst[j4] = u;
          }
        }
        if (p.stats != nullptr) {
          float sq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = valid ? v[j] : 0.f;
            sq[j] = v[j] * v[j];
          }
          acc_s[chunk] += butterfly_colsum32(v, lane);
          acc_q[chunk] += butterfly_colsum32(sq, lane);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }

    const FpropParams& p = P.c[0];            // statistics: single-problem launches only
    if (p.stats != nullptr) {
      // stats_smem[q][2*BN]: (sum[BN], sumsq[BN]) per epilogue warp, then a fixed-order 4-way add.
#pragma unroll
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        stats_smem[q * 2 * BN + chunk * 32 + lane] = acc_s[chunk];
        stats_smem[q * 2 * BN + BN + chunk * 32 + lane] = acc_q[chunk];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int e = threadIdx.x - 64;
      for (int i = e; i < 2 * BN; i += 128) {
        float t = stats_smem[i] + stats_smem[2 * BN + i] + stats_smem[4 * BN + i] + stats_smem[6 * BN + i];
        p.stats[(size_t)blockIdx.x * 2 * BN + i] = t;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static int pick_bn(int Cout) { return (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : (Cout % 64 == 0) ? 64 : 32; }
// Small-M problems (the last discriminator layers: 16-64 M tiles) leave most SMs idle while each busy CTA streams the
// whole K extent of a 256-wide weight panel through one SM's TMA path (1.7 MB per CTA for 256->256: ~25 us for 1 us of
// math).  Narrower N tiles put more SMs to work on the same bytes: halve the tile while the launch still fits `budget`
// CTAs (one wave).
static int pick_bn_parallel(long long m_tiles, int Cout, int budget) {
  static const bool narrow = [] { const char* e = getenv("PCG_TC_NARROW"); return e == nullptr || atoi(e) != 0; }();
  int bn = pick_bn(Cout);
  while (narrow && bn > 64 && m_tiles * (Cout / (bn / 2)) <= budget) bn /= 2;
  return bn;
}

int conv_tc_grid(long long M, int Cout) {
  int bn = pick_bn(Cout);
  long long tiles = ((M + TILE_M - 1) / TILE_M) * (Cout / bn);
  int sms = sm_count();
  return (int)(tiles < sms ? tiles : sms);
}

template <int BN, int NC>
static void launch_fprop_args(const FpropArgs<NC>& args, cudaStream_t stream) {
  PCG_PROFILE("conv_tc_fprop", stream);
  using Cfg = FpropCfg<BN>;
  static bool configured = false;
  if (!configured) {
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_fprop_kernel<BN, NC>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = args.tile_end[NC - 1];
  const int grid = tiles < sm_count() ? tiles : sm_count();
  launch_k_pdl(conv_tc_fprop_kernel<BN, NC>, dim3(grid), dim3(FPROP_THREADS), Cfg::SMEM_BYTES, stream, args);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}
template <int BN>
static void launch_fprop(const CUtensorMap& tmA, const CUtensorMap& tmB, const FpropParams& p, int grid,
                         cudaStream_t stream) {
  (void)grid;                                   // == min(tiles, SMs)
  FpropArgs<1> args;
  args.a[0] = tmA; args.b[0] = tmB; args.c[0] = p;
  args.tile_end[0] = p.num_m_tiles * p.num_n_tiles;
  launch_fprop_args<BN, 1>(args, stream);
}

void conv_tc_fprop(const bf16* in, int N, int H, int W, int Cin, const bf16* wpk, int Cout, int ksize,
                   int stride, int pad, const ConvEpilogue& epi, bf16* out, cudaStream_t stream) {
  PCG_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tensor-core conv needs Cin, Cout multiples of 64");
  PCG_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)wpk & 15) == 0 && ((uintptr_t)out & 15) == 0,
              "16-byte alignment");
  const int Ho = (H + 2 * pad - ksize) / stride + 1;
  const int Wo = (W + 2 * pad - ksize) / stride + 1;
  const long long M = (long long)N * Ho * Wo;
  const int bn = epi.stats != nullptr ? pick_bn(Cout) : pick_bn_parallel((M + TILE_M - 1) / TILE_M, Cout, sm_count());
  PCG_REQUIRE(epi.stats == nullptr || Cout == bn, "BN statistics need a single N tile");
  FpropParams p;
  p.M = (int)M; p.Ho = Ho; p.Wo = Wo; p.tstride = stride; p.lower_h = -pad; p.lower_w = -pad;
  p.taps_h = ksize; p.taps_w = ksize;
  p.OH = Ho; p.OW = Wo; p.os = 1; p.oh0 = 0; p.ow0 = 0;
  p.cin_blocks = Cin / 64; p.Cout = Cout;
  p.num_m_tiles = (int)((M + TILE_M - 1) / TILE_M);
  p.num_n_tiles = Cout / bn;
  p.bias = epi.bias; p.act = epi.act; p.slope = epi.slope; p.add_src = epi.add_src;
  p.act_ref = epi.ref_act != ACT_NONE ? epi.act_ref : nullptr; p.ref_act = epi.ref_act; p.ref_slope = epi.ref_slope;
  p.out = out; p.out_f32 = epi.out_f32; p.stats = epi.stats;
  CUtensorMap tmA = make_tmap_im2col(in, N, H, W, Cin, ksize, stride, pad);
  CUtensorMap tmB = make_tmap_2d(wpk, Cout, (uint64_t)ksize * ksize * Cin, bn);
  const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());     // == conv_tc_grid(M, Cout) when statistics are taken
  if (bn == 256) launch_fprop<256>(tmA, tmB, p, grid, stream);
  else if (bn == 128) launch_fprop<128>(tmA, tmB, p, grid, stream);
  else launch_fprop<64>(tmA, tmB, p, grid, stream);
}


// ------------------------------------------------------------------------------------------
// data gradient of a 3x3 / stride-2 / pad-1 convolution: four parity classes (hi%2, wi%2), each a
// stride-1 implicit GEMM over dY with 1, 2, 2 or 4 taps whose rows scatter to every other pixel of dX
// ------------------------------------------------------------------------------------------
__global__ void pack_dgrad_s2_kernel(const float* __restrict__ w, int Cout, int Cin, bf16* __restrict__ c00,
                                     bf16* __restrict__ c01, bf16* __restrict__ c10, bf16* __restrict__ c11) {
  pdl_enter();
  // class (ph,pw): taps_h = ph ? 2 : 1; window offset oh -> filter row r = ph ? (oh == 0 ? 2 : 0) : 1
  const int total = Cout * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int s = i % 3, r = (i / 3) % 3, ci = (i / 9) % Cin, co = i / (9 * Cin);
    const int ph = (r == 1) ? 0 : 1, pw = (s == 1) ? 0 : 1;
    const int oh = (r == 0) ? 1 : 0, ow = (s == 0) ? 1 : 0;
    const int taps_w = pw ? 2 : 1, taps = (ph ? 2 : 1) * taps_w;
    bf16* dst = ph ? (pw ? c11 : c10) : (pw ? c01 : c00);
    dst[((size_t)ci * taps + oh * taps_w + ow) * Cout + co] = __float2bfloat16_rn(w[i]);
  }
}

size_t conv_tc_dgrad_s2_pack_elems(int Cout, int Cin) { return (size_t)Cout * Cin * 9; }

// 4x4 / stride 2 / pad 1: every parity class has 2 x 2 taps; window offset oh -> filter row r = (ph ? 2 : 3) - 2*oh.
// Source layout wd fp32 [Cin][16][Cout] (taps not rotated); packed = four [Cin][4][Cout] bf16 matrices.
// `dup` = 3: every [Cout] run becomes [W_hi | W_hi | W_lo] (3*Cout per tap) for a dY operand split into hi | lo | hi
// (bf16x3 emulation of an fp32 product: hi*W_hi + lo*W_hi + hi*W_lo).
__global__ void pack_dgrad_s2_k4_kernel(const float* __restrict__ wd, int Cout, int Cin, int dup,
                                        bf16* __restrict__ packed) {
  pdl_enter();
  const int total = Cout * Cin * 16;
  const int ld = dup * Cout;
  const size_t u4 = (size_t)ld * Cin * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % Cout, tap = (i / Cout) % 16, ci = i / (Cout * 16);
    const int r = tap >> 2, s = tap & 3;
    const int ph = (r & 1) ? 0 : 1, pw = (s & 1) ? 0 : 1;     // r in {1,3} serves even rows, {0,2} odd rows
    const int oh = ((ph ? 2 : 3) - r) >> 1, ow = ((pw ? 2 : 3) - s) >> 1;
    const bf16 v = __float2bfloat16_rn(wd[i]);
    bf16* dst = packed + (size_t)(ph * 2 + pw) * u4 + ((size_t)ci * 4 + oh * 2 + ow) * ld + co;
    dst[0] = v;
    if (dup == 3) {
      dst[Cout] = v;
      dst[2 * Cout] = __float2bfloat16_rn(wd[i] - __bfloat162float(v));
    }
  }
}
void pack_dgrad_s2_k4_tc(const float* wd, int Cout, int Cin, bf16* packed, cudaStream_t stream, int dup) {
  PCG_PROFILE("pack_weights", stream);
  const long long total = (long long)Cout * Cin * 16;
  launch_k(pack_dgrad_s2_k4_kernel, dim3(cdiv(total, 256) > 1184 ? 1184 : cdiv(total, 256)), dim3(256), 0, stream, wd, Cout, Cin, dup, packed);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

void pack_dgrad_s2_tc(const float* w, int Cout, int Cin, bf16* packed, cudaStream_t stream) {
  PCG_PROFILE("pack_weights", stream);
  const size_t u = (size_t)Cout * Cin;
  launch_k(pack_dgrad_s2_kernel, dim3(cdiv((long long)u * 9, 256) > 1184 ? 1184 : cdiv((long long)u * 9, 256)), dim3(256), 0, stream, 
      w, Cout, Cin, packed, packed + u, packed + 3 * u, packed + 5 * u);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

void conv_tc_dgrad_s2(const bf16* dy, int N, int H, int W, int Cin, int Cout, const bf16* packed,
                      const ConvEpilogue& epi, bf16* dx, cudaStream_t stream, const cudaStream_t* class_streams,
                      int ksize) {
  PCG_REQUIRE(Cout % 64 == 0 && Cin % 32 == 0, "strided tensor-core dgrad needs Cout % 64 == 0, Cin % 32 == 0");
  PCG_REQUIRE(epi.stats == nullptr && epi.bias == nullptr, "no bias / statistics in the dgrad epilogue");
  PCG_REQUIRE(ksize == 3 || (ksize == 4 && H % 2 == 0 && W % 2 == 0), "stride-2 dgrad: 3x3, or 4x4 on even sizes");
  const int Ho = (H + 2 - ksize) / 2 + 1, Wo = (W + 2 - ksize) / 2 + 1;
  const size_t u = (size_t)Cout * Cin;
  const size_t cls_off3[4] = {0, u, 3 * u, 5 * u};
  // The four parity classes go out as ONE launch (tiles of the class with the most taps first): one graph node instead
  // of four and one tile queue, so the tail of a class overlaps the head of the next.
  (void)class_streams;
  long long all_m_tiles = 0;
  for (int cls = 0; cls < 4; ++cls)
    all_m_tiles += ((long long)N * ((H + 1 - (cls >> 1)) / 2) * ((W + 1 - (cls & 1)) / 2) + TILE_M - 1) / TILE_M;
  const int bn = pick_bn_parallel(all_m_tiles, Cin, sm_count());
  FpropArgs<4> args;
  int slot = 0, tile_end = 0;
  for (int cls = 3; cls >= 0; --cls) {
    const int ph = cls >> 1, pw = cls & 1;
    const int Ah = (H + 1 - ph) / 2, Aw = (W + 1 - pw) / 2;
    PCG_REQUIRE(Ah > 0 && Aw > 0, "stride-2 dgrad: every parity class needs at least one pixel");
    // 3x3: 1 or 2 taps per dimension, windows start at the class pixel; 4x4: always 2 taps, even rows look one back
    const int th = ksize == 4 ? 2 : (ph ? 2 : 1), tw = ksize == 4 ? 2 : (pw ? 2 : 1);
    const int lo_h = (ksize == 4 && ph == 0) ? -1 : 0, lo_w = (ksize == 4 && pw == 0) ? -1 : 0;
    const size_t cls_off = ksize == 4 ? (size_t)cls * 4 * u : cls_off3[cls];
    FpropParams p;
    p.M = N * Ah * Aw; p.Ho = Ah; p.Wo = Aw; p.tstride = 1; p.lower_h = lo_h; p.lower_w = lo_w;
    p.taps_h = th; p.taps_w = tw;
    p.OH = H; p.OW = W; p.os = 2; p.oh0 = ph; p.ow0 = pw;
    p.cin_blocks = Cout / 64; p.Cout = Cin;
    p.num_m_tiles = (p.M + TILE_M - 1) / TILE_M;
    p.num_n_tiles = Cin / bn;
    p.bias = nullptr; p.act = epi.act; p.slope = epi.slope; p.add_src = epi.add_src;
    p.act_ref = epi.ref_act != ACT_NONE ? epi.act_ref : nullptr; p.ref_act = epi.ref_act; p.ref_slope = epi.ref_slope;
    p.out = dx; p.out_f32 = epi.out_f32; p.stats = nullptr;
    args.a[slot] = make_tmap_im2col_box(dy, N, Ho, Wo, Cout, lo_w, lo_h, Aw - Wo + lo_w, Ah - Ho + lo_h, 1);
    args.b[slot] = make_tmap_2d(packed + cls_off, Cin, (uint64_t)th * tw * Cout, bn);
    args.c[slot] = p;
    tile_end += p.num_m_tiles * p.num_n_tiles;
    args.tile_end[slot] = tile_end;
    ++slot;
  }
  if (bn == 256) launch_fprop_args<256, 4>(args, stream);
  else if (bn == 128) launch_fprop_args<128, 4>(args, stream);
  else if (bn == 64) launch_fprop_args<64, 4>(args, stream);
  else launch_fprop_args<32, 4>(args, stream);
}
int conv_tc_dgrad_s2_class_ctas(int N, int H, int W, int Cin) {
  (void)N; (void)H; (void)W; (void)Cin;
  return sm_count();                            // the four classes are one launch: nothing to run side by side
}

// ------------------------------------------------------------------------------------------
// wgrad (Cin = Cout = 64, 3x3, stride 1, pad 1)
// ------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 192;
constexpr int WG_X_STAGE = 2 * A_STAGE_BYTES;      // two taps per stage
constexpr int WG_STAGES = 4;
constexpr int WG_DY_BYTES = A_STAGE_BYTES;
constexpr int WG_TMEM_COLS = 512;                  // 5 accumulators x 64 columns, power of two
constexpr int WG_SMEM_BYTES = 1024 + WG_STAGES * WG_X_STAGE + 2 * WG_DY_BYTES + 256;

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_tc_wgrad64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                       int M, int H, int W, int num_pblocks, float* __restrict__ part) {
  pdl_launch_dependents();                        // prologue first (no global memory), pdl_wait() below
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sx = smem;                                    // [WG_STAGES][2][128 px][64 ci]
  uint8_t* sdy = smem + WG_STAGES * WG_X_STAGE;          // [2][128 px][64 co]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdy + 2 * WG_DY_BYTES);
  uint64_t* full = bars;                 // [WG_STAGES]
  uint64_t* empty = bars + WG_STAGES;    // [WG_STAGES]
  uint64_t* dyfull = bars + 2 * WG_STAGES;   // [2]
  uint64_t* dyempty = dyfull + 2;            // [2]
  uint64_t* done = dyempty + 2;              // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&dyfull[s], 1); mbar_init(&dyempty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, WG_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                     // the predecessor kernel has completed and flushed

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, db = 0;
      uint32_t phase = 0, dphase = 0;
      const int hw = H * W;
      for (int pb = blockIdx.x; pb < num_pblocks; pb += gridDim.x) {
        const int p0 = pb * TILE_M;
        const int n_img = p0 / hw, rem = p0 % hw;
        const int cw = rem % W - 1, ch = rem / W - 1;
        mbar_wait(&dyempty[db], dphase ^ 1);
        mbar_expect_tx(&dyfull[db], WG_DY_BYTES);
        tma_load_2d(&tmDY, &dyfull[db], sdy + db * WG_DY_BYTES, 0, p0);
        for (int j = 0; j < 5; ++j) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], j < 4 ? WG_X_STAGE : A_STAGE_BYTES);
          uint8_t* dst = sx + stage * WG_X_STAGE;
          const int t0 = 2 * j;
          tma_load_im2col_4d(&tmX, &full[stage], dst, 0, cw, ch, n_img, (uint16_t)(t0 % 3),
                             (uint16_t)(t0 / 3));
          if (j < 4) {
            const int t1 = t0 + 1;
            tma_load_im2col_4d(&tmX, &full[stage], dst + A_STAGE_BYTES, 0, cw, ch, n_img,
                               (uint16_t)(t1 % 3), (uint16_t)(t1 / 3));
          }
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        db ^= 1;
        if (db == 0) dphase ^= 1;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // A = [X@tap(2j) | X@tap(2j+1)] : MN-major, M = 128 (two 64-channel blocks LBO apart)
      // B = dY tile                     : MN-major, N = 64
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      int stage = 0, db = 0;
      uint32_t phase = 0, dphase = 0;
      bool first = true;
      for (int pb = blockIdx.x; pb < num_pblocks; pb += gridDim.x) {
        mbar_wait(&dyfull[db], dphase);
        tc_fence_after();
        const uint32_t b_base = smem_u32(sdy + db * WG_DY_BYTES);
        for (int j = 0; j < 5; ++j) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_smem_desc(smem_u32(sx + stage * WG_X_STAGE), A_STAGE_BYTES, 1024);
          const uint64_t dbd = umma_smem_desc(b_base, A_STAGE_BYTES, 1024);
          // 8 x (K = 16 pixels = 2 swizzle atoms of 8 rows = 2048 B): start-address field += 128
          if (first) {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_f16(tmem_base + j * 64, da + 128 * k, dbd + 128 * k, idesc, k != 0 ? 1u : 0u);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_f16(tmem_base + j * 64, da + 128 * k, dbd + 128 * k, idesc, 1u);
          }
          umma_commit(&empty[stage]);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&dyempty[db]);
        first = false;
        db ^= 1;
        if (db == 0) dphase ^= 1;
      }
      umma_commit(done);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;            // accumulator row = (tap parity, ci)
    mbar_wait(done, 0);
    tc_fence_after();
    float* my = part + (size_t)blockIdx.x * 9 * 64 * 64;
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
      const int tap = 2 * j + (row >> 6);
      const int ci = row & 63;
#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + j * 64 + chunk * 32, r);
        tmem_ld_wait();
        if (tap < 9) {
          float4* dst = reinterpret_cast<float4*>(my + ((size_t)tap * 64 + ci) * 64 + chunk * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            dst[j4] = make_float4(__uint_as_float(r[j4 * 4]), __uint_as_float(r[j4 * 4 + 1]),
                                  __uint_as_float(r[j4 * 4 + 2]), __uint_as_float(r[j4 * 4 + 3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, WG_TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------
// general wgrad on the tensor cores (3x3, pad 1, stride 1 or 2, Cin % 64 == 0, Cout % 64 == 0)
//   part[z][co][(tap, ci)] = sum over the z-th pixel slice of dY[p][co] * X[p@tap][ci]
// CTA = (pair of (tap, 64-channel block) units, K slice).  A = two im2col tiles of X (MN-major, M = 128),
// B = the dY tile (MN-major, N = Cout tile up to 256), K = 128 pixels per pipeline stage; one fp32
// accumulator (128 lanes x Ntile columns) in TMEM per CTA; partials reduced by wgrad_reduce_generic.
// ------------------------------------------------------------------------------------------
constexpr int WGG_THREADS = 192;
constexpr int WGG_STAGES = 2;

template <int NT>   // Cout tile: 64, 128 or 256
struct WggCfg {
  static constexpr int DY_BYTES = (NT / 64) * A_STAGE_BYTES;
  static constexpr int STAGE_BYTES = 2 * A_STAGE_BYTES + DY_BYTES;
  static constexpr int SMEM_BYTES = 1024 + WGG_STAGES * STAGE_BYTES + 256;
};

struct WggParams {
  int P, Ho, Wo, stride, cin_blocks, Cout, K;       // K = ksize^2 * Cin
  int ksize, pad;
  int units, pairs, n_tiles, ksplit, kb_per_split, num_kb;
  float* part;
};

template <int NT>
__global__ void __launch_bounds__(WGG_THREADS, 1)
conv_tc_wgrad_general_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                             const WggParams p) {
  pdl_launch_dependents();                        // prologue first (no global memory), pdl_wait() below
  using Cfg = WggCfg<NT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WGG_STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + WGG_STAGES;
  uint64_t* done = bars + 2 * WGG_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item
  int w = blockIdx.x;
  const int z = w % p.ksplit; w /= p.ksplit;
  const int nt = w % p.n_tiles; w /= p.n_tiles;
  const int pair = w;
  const int u0 = 2 * pair, u1 = (u0 + 1 < p.units) ? u0 + 1 : u0;
  const int kb0 = z * p.kb_per_split;
  const int kb1 = (kb0 + p.kb_per_split < p.num_kb) ? kb0 + p.kb_per_split : p.num_kb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < WGG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, NT < 32 ? 32 : NT);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                     // the predecessor kernel has completed and flushed

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int hw = p.Ho * p.Wo;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int p0 = kb * TILE_M;
        const int n_img = p0 / hw, rem = p0 % hw;
        const int cw = (rem % p.Wo) * p.stride - p.pad, ch = (rem / p.Wo) * p.stride - p.pad;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
        uint8_t* dst = smem + stage * Cfg::STAGE_BYTES;
        for (int h = 0; h < 2; ++h) {
          const int u = h ? u1 : u0;
          const int tap = u / p.cin_blocks, cb = u % p.cin_blocks;
          tma_load_im2col_4d(&tmX, &full[stage], dst + h * A_STAGE_BYTES, cb * 64, cw, ch, n_img,
                             (uint16_t)(tap % p.ksize), (uint16_t)(tap / p.ksize));
        }
        for (int nb = 0; nb < NT / 64; ++nb)
          tma_load_2d(&tmDY, &full[stage], dst + (2 + nb) * A_STAGE_BYTES, nt * NT + nb * 64, p0);
        if (++stage == WGG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, NT, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
      const uint64_t da = umma_smem_desc(base, A_STAGE_BYTES, 1024);                        // M blocks 16 KB apart
      const uint64_t db = umma_smem_desc(base + 2 * A_STAGE_BYTES, A_STAGE_BYTES, 1024);    // N blocks 16 KB apart
      if (elect_one()) {
        if (kb == kb0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_f16(tmem_base, da + 128 * k, db + 128 * k, idesc, k != 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_f16(tmem_base, da + 128 * k, db + 128 * k, idesc, 1u);
        }
        umma_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == WGG_STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int u = (row >> 6) ? u1 : u0;
    const bool row_ok = (row < 64) || (u1 != u0);
    const int kcol = u * 64 + (row & 63);           // (tap, ci) column of the partial matrix
    mbar_wait(done, 0);
    tc_fence_after();
    float* my = p.part + (size_t)z * p.Cout * p.K;
    const bool any = kb1 > kb0;
#pragma unroll 1
    for (int chunk = 0; chunk < NT / 32; ++chunk) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + chunk * 32, r);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int co = nt * NT + chunk * 32 + j;
          my[(size_t)co * p.K + kcol] = any ? __uint_as_float(r[j]) : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NT < 32 ? 32 : NT);
  }
}

static void wgg_plan(int N, int H, int W, int Cin, int Cout, int stride, WggParams& p, int& nt_size, int ksize = 3,
                     int pad = 1) {
  const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
  p.P = N * Ho * Wo; p.Ho = Ho; p.Wo = Wo; p.stride = stride; p.ksize = ksize; p.pad = pad;
  p.cin_blocks = Cin / 64; p.Cout = Cout; p.K = ksize * ksize * Cin;
  p.units = ksize * ksize * p.cin_blocks; p.pairs = (p.units + 1) / 2;
  nt_size = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : 64;
  p.n_tiles = Cout / nt_size;
  p.num_kb = (p.P + TILE_M - 1) / TILE_M;
  int want = (2 * sm_count()) / (p.pairs * p.n_tiles);
  if (want < 1) want = 1;
  if (want > p.num_kb) want = p.num_kb;
  p.kb_per_split = (p.num_kb + want - 1) / want;
  p.ksplit = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
}

int conv_tc_wgrad_general_splits(int N, int H, int W, int Cin, int Cout, int stride, int ksize, int pad) {
  WggParams p; int nt;
  wgg_plan(N, H, W, Cin, Cout, stride, p, nt, ksize, pad);
  return p.ksplit;
}

void conv_tc_wgrad_general(const bf16* x, const bf16* dy, int N, int H, int W, int Cin, int Cout, int stride,
                           float* part, cudaStream_t stream, int ksize, int pad) {
  PCG_PROFILE("conv_tc_wgrad_general", stream);
  PCG_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0 && (stride == 1 || stride == 2), "unsupported wgrad shape");
  WggParams p; int nt;
  wgg_plan(N, H, W, Cin, Cout, stride, p, nt, ksize, pad);
  p.part = part;
  CUtensorMap tmX = make_tmap_im2col(x, N, H, W, Cin, ksize, stride, pad);
  CUtensorMap tmDY = make_tmap_2d(dy, (uint64_t)p.P, (uint64_t)Cout, 128);
  const int grid = p.pairs * p.n_tiles * p.ksplit;
#define PCG_WGG(NTV)                                                                                               \
  {                                                                                                                \
    static bool configured = false;                                                                                \
    if (!configured) {                                                                                             \
      PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_general_kernel<NTV>,                                       \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, WggCfg<NTV>::SMEM_BYTES)); \
      configured = true;                                                                                           \
    }                                                                                                              \
    launch_k_pdl(conv_tc_wgrad_general_kernel<NTV>, dim3(grid), dim3(WGG_THREADS), WggCfg<NTV>::SMEM_BYTES, stream, tmX, tmDY, p);       \
  }
  if (nt == 256) PCG_WGG(256) else if (nt == 128) PCG_WGG(128) else PCG_WGG(64)
#undef PCG_WGG
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

int conv_tc_wgrad_grid(long long M) {
  long long nb = (M + TILE_M - 1) / TILE_M;
  int sms = sm_count();
  return (int)(nb < sms ? nb : sms);
}

void conv_tc_wgrad64(const bf16* x, const bf16* dy, int N, int H, int W, float* part,
                     cudaStream_t stream) {
  PCG_PROFILE("conv_tc_wgrad64", stream);
  const long long M = (long long)N * H * W;
  CUtensorMap tmX = make_tmap_im2col(x, N, H, W, 64, 3, 1, 1);
  CUtensorMap tmDY = make_tmap_2d(dy, (uint64_t)M, 64, 128);
  static bool configured = false;
  if (!configured) {
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_wgrad64_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
    configured = true;
  }
  const int nb = (int)((M + TILE_M - 1) / TILE_M);
  const int grid = conv_tc_wgrad_grid(M);
  launch_k_pdl(conv_tc_wgrad64_kernel, dim3(grid), dim3(WG_THREADS), WG_SMEM_BYTES, stream, tmX, tmDY, (int)M, H, W, nb, part);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// part[cta][tap][ci][co] -> dw[co][ci][tap]; fixed summation order (deterministic).
// One thread per (four consecutive outputs, partial group g of 8): it sums the partials c = g, g + 8, ... as two
// independent float4 chains, so ~18 sixteen-byte loads per thread are in flight instead of 148 dependent scalar ones
// (the partials are L2-resident: 148 x 147 KB); the eight groups are then added in fixed order through shared memory.
constexpr int WR_GROUPS = 8, WR_VEC = 32;
__global__ void __launch_bounds__(WR_GROUPS * WR_VEC)
wgrad_reduce_tc_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dw, int swizzled) {
  pdl_enter();
  __shared__ float4 sm[WR_GROUPS][WR_VEC];
  const int v = threadIdx.x % WR_VEC, g = threadIdx.x / WR_VEC;
  const int idx4 = blockIdx.x * WR_VEC + v;                  // outputs idx4*4 .. idx4*4+3 of (tap, ci, co), co fastest
  const float4* p4 = reinterpret_cast<const float4*>(part) + idx4;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  int c = g;
#pragma unroll 4
  for (; c + WR_GROUPS < nparts; c += 2 * WR_GROUPS) {
    const float4 a = p4[(size_t)c * 9216], b = p4[(size_t)(c + WR_GROUPS) * 9216];
    s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
    s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
  }
  if (c < nparts) {
    const float4 a = p4[(size_t)c * 9216];
    s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
  }
  sm[g][v] = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
  __syncthreads();
  if (g == 0) {
    float4 t = sm[0][v];
#pragma unroll
    for (int k = 1; k < WR_GROUPS; ++k) {
      const float4 u = sm[k][v];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    // swizzled partials (conv_tc64_wgrad): 16-byte chunk c of row (tap, ci) is stored at chunk c ^ (ci & 15)
    const int ci = (idx4 >> 4) & 63, tap = idx4 >> 10;
    const int co = ((idx4 & 15) ^ (swizzled ? (ci & 15) : 0)) * 4;
    dw[((co + 0) * 64 + ci) * 9 + tap] = t.x;
    dw[((co + 1) * 64 + ci) * 9 + tap] = t.y;
    dw[((co + 2) * 64 + ci) * 9 + tap] = t.z;
    dw[((co + 3) * 64 + ci) * 9 + tap] = t.w;
  }
}

void wgrad_reduce_tc(const float* part, int nparts, float* dw, cudaStream_t stream, bool swizzled) {
  PCG_PROFILE("wgrad_reduce_tc", stream);
  launch_k(wgrad_reduce_tc_kernel, dim3(9216 / WR_VEC), dim3(WR_GROUPS * WR_VEC), 0, stream, part, nparts, dw, swizzled ? 1 : 0);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------
__global__ void pack_conv_weights_tc_kernel(const float* __restrict__ w, int Cout, int Cin, int ksize,
                                            bf16* __restrict__ fprop, bf16* __restrict__ dgrad) {
  pdl_enter();
  const int taps = ksize * ksize;
  const int total = Cout * Cin * taps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % taps, ci = (i / taps) % Cin, co = i / (taps * Cin);
    const bf16 v = __float2bfloat16_rn(w[i]);
    if (fprop) fprop[((size_t)co * taps + tap) * Cin + ci] = v;
    if (dgrad) dgrad[((size_t)ci * taps + (taps - 1 - tap)) * Cout + co] = v;
  }
}

void pack_conv_weights_tc(const float* w, int Cout, int Cin, int ksize, bf16* fprop, bf16* dgrad,
                          cudaStream_t stream) {
  PCG_PROFILE("pack_weights", stream);
  const int total = Cout * Cin * ksize * ksize;
  launch_k(pack_conv_weights_tc_kernel, dim3(cdiv(total, 256)), dim3(256), 0, stream, w, Cout, Cin, ksize, fprop, dgrad);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// debug: one im2col box
// ------------------------------------------------------------------------------------------
__global__ void debug_im2col_kernel(const __grid_constant__ CUtensorMap tmA, int cblock, int cw, int ch,
                                    int n_img, int s, int r, bf16* __restrict__ out) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_STAGE_BYTES);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, A_STAGE_BYTES);
    tma_load_im2col_4d(&tmA, bar, smem, cblock * 64, cw, ch, n_img, (uint16_t)s, (uint16_t)r);
  }
  mbar_wait(bar, 0);
  // de-swizzle: row = pixel (128 B), 16-byte chunk c stored at chunk (c ^ (row & 7))
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int row = i >> 3, c = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(smem + row * 128 + ((c ^ (row & 7)) << 4));
    reinterpret_cast<uint4*>(out)[row * 8 + c] = v;
  }
}

void debug_im2col_tile(const bf16* in, int N, int H, int W, int Cin, int ksize, int stride, int pad,
                       int first_pixel, int tap_r, int tap_s, int cblock, bf16* out128x64,
                       cudaStream_t stream) {
  const int Ho = (H + 2 * pad - ksize) / stride + 1;
  const int Wo = (W + 2 * pad - ksize) / stride + 1;
  CUtensorMap tmA = make_tmap_im2col(in, N, H, W, Cin, ksize, stride, pad);
  const int hw = Ho * Wo;
  const int n_img = first_pixel / hw, rem = first_pixel % hw;
  const int cw = (rem % Wo) * stride - pad, ch = (rem / Wo) * stride - pad;
  launch_k(debug_im2col_kernel, dim3(1), dim3(128), A_STAGE_BYTES + 1024 + 64, stream, tmA, cblock, cw, ch, n_img, tap_s,
                                                                      tap_r, out128x64);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

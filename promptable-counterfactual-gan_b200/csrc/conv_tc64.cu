// 64 -> 64 channel, 3x3, stride-1, pad-1 convolution kernels with a shared-memory halo tile.
//
// The im2col kernels of conv_tc.cu fetch every filter tap from L2 separately (9 x 16 KB per 128 output
// pixels) and re-fetch the weights for every tile: ~216 KB of L2->SM traffic per tile, which makes them
// L2-bandwidth bound (8 TB/s) at 6x the MMA time.  Here one 4-D TMA box per tile brings the (R+2) x (W+2)
// zero-padded input pixels of R output rows (23 KB for 28x28 images), and the nine taps are *shifted views*
// of that tile: output position i = hh*(W+2) + ww reads tile row i + r*(W+2) + s for tap (r, s), so the A
// operand of tap (r, s) is the same shared-memory matrix started (r*(W+2)+s) rows later.  The packed weights
// (72 KB) are loaded once per CTA.  Two of every W+2 accumulator rows are padding columns and are discarded
// (R*W of 128 MMA rows are useful: 112/128 for 28x28).
//
// Replaces: nn.Conv2d(64, 64, 3, padding=1) forward and ConvolutionBackward0 (dgrad with rotated weights,
// wgrad) of conditional_counteRGAN/mnist/models/generator.py:11,14,49.
//
// Kernels in this file (profiles/exp_tc64_stacked_r1.md):
//   conv_tc64s_fprop_kernel  forward / data gradient, default when H % 4 == 0: row classes h % 4, tap matrices stacked
//                            along N (72 instead of 144 MMAs per 512 positions), store warp + mbarrier-handed staging
//   conv_tc64_fprop_kernel   forward / data gradient, one output class per tile: every other height, and the A/B
//                            partner of the stacked kernel (variant bit 256); same epilogue contract
//   conv_tc64_wgrad_kernel   weight gradient: shifted views on both operands (16 instead of 40 MMAs per tile;
//                            variant bit 1024 = the two-taps-per-accumulator scheme)
// Variant bits (pcg_conv_tc64_set_variant / PCG_TC64_VARIANT): 1 descriptor base-offset policy (bring-up), 2 / 4 / 8 /
// 16 / 64 / 128 timing experiments (results invalid), 256 one-class forward kernel, 1024 original weight-gradient scheme.
// 2048: weight-gradient partials written by direct stores (linear layout) instead of staged bulk copies.
#include <cstdlib>

#include "conv_tc.cuh"
#include "elementwise.cuh"
#include "tc_common.cuh"

namespace pcg {
using namespace tc;

static int env_variant() {
  const char* e = getenv("PCG_TC64_VARIANT");      // A/B switch for bench runs (same bits as pcg_conv_tc64_set_variant)
  return e ? atoi(e) : 0;
}
static int g_variant = env_variant();
void conv_tc64_set_variant(int v) { g_variant = v; }

constexpr int C64 = 64;
constexpr int W_BYTES = 9 * 64 * 128;            // resident weights: 9 taps x [64 rows][64 k] bf16
constexpr int IN_STAGE_BYTES = 24576;            // wgrad: >= (127 + 2*(W+2) + 2 + 1) * 128 for W <= 28
constexpr int F_EPI_WARPS = 16;                  // 4 TMEM lane quarters x 4 column quarters (16 columns per thread)
constexpr int F_EPI_THREADS = F_EPI_WARPS * 32;
constexpr int F_EPI_WARP0 = 3;                  // warp0 TMA (input), warps 1-2 MMA (even / odd tiles), warps 3-18 epilogue,
constexpr int F_THREADS = 32 * (F_EPI_WARP0 + F_EPI_WARPS + 1);   // warp19 TMA (epilogue operand)
constexpr int F_TAIL_BYTES = (16 * 2 * 16 + 64 + 4 * 64) * 4 + 512;   // statistics scratch, bias, BN constants; barriers   // statistics scratch + bias, mbarriers + TMEM pointer
constexpr int SMEM_LIMIT = 232448;               // 227 KB opt-in maximum per CTA

struct F64Params {
  int N, H, W, WP, R, tiles_per_img, total_tiles;
  int in_stages, in_stage_bytes, out_tile_bytes; // shared-memory ring geometry (host-computed)
  int n_extra;                                   // 1: one epilogue operand tile per output tile arrives by TMA
  int extra_is_add;                              // that operand is add_src (else act_ref)
  int bn_bwd;                                    // the operand is the BatchNorm input y: fused backward reduction
  int dual;                                      // stacked kernel: bn_bwd AND a residual add; the residual tile arrives by TMA
                                                 // in the staging slot its result leaves from (three in-place slots)
  float bn_gscale;                               // the reduction is taken of bn_gscale * result
  const float *bn_mean, *bn_rstd, *bn_scale, *bn_shift;
  int bn_act;
  float bn_slope;
  const float* bias;
  int act;
  float slope;
  const bf16* add_src;                           // read from global only when it is not the TMA operand
  const bf16* act_ref;
  int ref_act;
  float ref_slope;
  float* stats;
  StatsFinalize fin;
  int variant;
  int add_evict_first;                           // dual: the skip operand is read for the last time (PCG_L2_HINTS bit 4)
  int in_evict_first;                            // forward + statistics: the input is not read again soon (bit 16)
};

// Runs in the epilogue warps (512 threads, named barrier 1) of the last CTA to finish: see StatsFinalize.
__device__ __forceinline__ void finalize_stats_last_cta(const StatsFinalize& f, const float* __restrict__ stats, int nrows,
                                                        double* red /* [4][128] shared */, int t) {
  {
    // 512 threads: column (t & 127) of [nrows][128] (first sums in columns 0-63, second sums in 64-127), rows
    // part, part + 4, ... with part = t >> 7; all loads of a thread are independent (L2 round trips overlap)
    const float* col = stats + (t & 127);
    const int part = t >> 7;
    float v[10];
    double a = 0.0;
    for (int i0 = part; i0 < nrows; i0 += 40) {
#pragma unroll
      for (int k = 0; k < 10; ++k) v[k] = (i0 + 4 * k) < nrows ? __ldcg(col + (size_t)(i0 + 4 * k) * 128) : 0.f;
#pragma unroll
      for (int k = 0; k < 10; ++k) a += (double)v[k];
    }
    red[part * 128 + (t & 127)] = a;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
  if (t < 64) {
    const int c = t;
    const double s1 = (red[c] + red[128 + c]) + (red[256 + c] + red[384 + c]);
    const double s2 = (red[64 + c] + red[192 + c]) + (red[320 + c] + red[448 + c]);
    const double M = (double)f.M;
    if (f.mode == 1) {
      const double mean = s1 / M;
      double var = s2 / M - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)f.eps));
      const float a = f.gamma[c] * rstd;
      f.mean[c] = (float)mean;
      f.rstd[c] = rstd;
      f.scale[c] = a;
      f.shift[c] = f.beta[c] - (float)mean * a;
      if (f.running_mean != nullptr) {
        const double unbiased = f.M > 1 ? var * (M / (M - 1.0)) : var;
        f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * (float)mean;
        f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unbiased;
      }
      if (c == 0 && f.nbt != nullptr) *f.nbt += 1;
    } else {
      f.dbeta[c] = (float)s1;
      f.dgamma[c] = (float)s2;
      f.c12[c] = (float)(s1 / M);
      f.c12[64 + c] = (float)(s2 / M);
    }
  }
}

bool conv_tc64_supported(int H, int W) {
  const int WP = W + 2;
  if (WP > 128) return false;
  const int R = 128 / WP;
  return R >= 1 && (127 + 2 * WP + 3) * 128 <= IN_STAGE_BYTES && (R + 2) <= 256 && WP <= 256;
}

int conv_tc64_grid(int N, int H, int W) {
  const int R = 128 / (W + 2);
  const long long tiles = (long long)N * ((H + R - 1) / R);
  const int sms = sm_count();
  return (int)(tiles < sms ? tiles : sms);
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, cnt = 32; off >= 1; off >>= 1, cnt >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < cnt / 2; ++j) {
      const float send = upper ? v[j] : v[j + cnt / 2];
      const float keep = upper ? v[j + cnt / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// 16-column variant: lanes l and l ^ 16 end up holding column (l & 15) summed over the warp's 32 rows
__device__ __forceinline__ float colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
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
  return v[0];
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__device__ __forceinline__ uint64_t a_desc(uint32_t addr, uint32_t lbo, int variant) {
  // variant 0: swizzle phase taken from the absolute shared-memory address (base_offset = 0)
  // variant 1: base_offset = (address >> 7) & 7, for starts that are not 1024-byte aligned
  return umma_smem_desc(addr, lbo, 1024, variant == 1 ? ((addr >> 7) & 7u) : 0u);
}

// Epilogue I/O goes through shared memory: a thread owns one accumulator row (= one pixel), so direct global
// accesses would touch 32 different 128-byte lines per warp instruction (32 LSU transactions each, ~1000 cycles per
// tile and per tensor - more than the tile's MMAs).  Instead the output tile is written to a 128B-swizzled staging
// buffer (conflict-free) and leaves by one TMA store; the residual / activation-reference tile of the same pixels
// arrives by TMA into a two-slot ring.
__global__ void __launch_bounds__(F_THREADS, 1)
conv_tc64_fprop_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmExtra,
                       const F64Params p) {
  pdl_launch_dependents();                        // the prologue below touches no global memory: see pdl_wait() further down
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sw = smem;                                   // weights
  uint8_t* sin = smem + W_BYTES;                        // input halo tiles
  uint8_t* sout = sin + p.in_stages * p.in_stage_bytes; // output staging, 2 slots
  uint8_t* sx = sout + 2 * p.out_tile_bytes;            // epilogue operand ring, 2 slots (if n_extra)
  float* stats_smem = reinterpret_cast<float*>(sx + (p.n_extra ? 2 : 0) * p.out_tile_bytes);   // [16][2][16] + bias[64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats_smem) + (16 * 2 * 16 + 64 + 4 * 64) * 4);
  uint64_t* full = bars;                 // [4]
  uint64_t* empty = bars + 4;            // [4]
  uint64_t* wfull = bars + 8;            // [1]
  uint64_t* tfull = wfull + 1;           // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint64_t* xfull = tempty + 2;          // [2]
  uint64_t* xempty = xfull + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmOut);
    if (p.n_extra) tma_prefetch_desc(&tmExtra);
    for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(wfull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1); mbar_init(&tempty[s], F_EPI_THREADS);
      mbar_init(&xfull[s], 1); mbar_init(&xempty[s], F_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 128);
    tmem_relinquish();
  }
  // rows of a halo tile beyond the TMA box are read by the (discarded) padding rows of the MMA: keep them finite
  {
    const int zbytes = (int)(reinterpret_cast<uint8_t*>(stats_smem) - sin);
    for (int i = threadIdx.x; i < zbytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sin)[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                     // from here on the predecessor kernel has completed and flushed
  const int box_bytes = (p.R + 2) * p.WP * 128;
  const int tile_bytes = p.R * p.W * 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, W_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(&tmW, wfull, sw + tap * 8192, tap * 64, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n = tile / p.tiles_per_img, h0 = (tile % p.tiles_per_img) * p.R;
        mbar_wait(&empty[stage], phase ^ 1);
        if (p.variant & 16) {                      // experiment: no input traffic
          mbar_arrive(&full[stage]);
        } else {
          mbar_expect_tx(&full[stage], box_bytes);
          tma_load_4d(&tmX, &full[stage], sin + stage * p.in_stage_bytes, 0, -1, h0 - 1, n);
        }
        if (++stage == p.in_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == F_EPI_WARP0 + F_EPI_WARPS) {
    if (lane == 0 && p.n_extra) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int n = tile / p.tiles_per_img, h0 = (tile % p.tiles_per_img) * p.R;
        const int slot = it & 1;
        mbar_wait(&xempty[slot], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&xfull[slot], tile_bytes);
        tma_load_4d(&tmExtra, &xfull[slot], sx + slot * p.out_tile_bytes, 0, 0, h0, n);
      }
    }
  } else if (warp == 1 || warp == 2) {
    {
      // Two MMA warps: issuing one M128xN64xK16 tcgen05.mma costs the issuing warp ~70 cycles (descriptor moves into
      // uniform registers) while the tensor core needs ~48, so warp 1 issues the even tiles of this CTA into
      // accumulator 0 and warp 2 the odd tiles into accumulator 1; the two instruction streams are independent.
      // The whole warp walks the loop (warp-uniform control flow), one elected lane issues and commits.
      const int mw = warp - 1;
      const int nsel = (p.variant >> 6) & 3;       // experiment: MMA N = 64 / 16 / 32 / 128
      const uint32_t idesc = nsel == 1 ? umma_idesc_bf16(128, 16, 0, 0) : nsel == 2 ? umma_idesc_bf16(128, 32, 0, 0)
                             : nsel == 3 ? umma_idesc_bf16(128, 128, 0, 0) : umma_idesc_bf16(128, 64, 0, 0);
      mbar_wait(wfull, 0);
      const uint64_t b0 = umma_smem_desc(smem_u32(sw), 16, 1024);
      const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
      const uint32_t d_tmem = tmem_base + mw * 64;
      const int ntap = (p.variant & 4) ? 1 : 9;    // experiment: one tap only
      int it = mw;                                 // CTA-local tile counter
      for (int tile = blockIdx.x + mw * gridDim.x; tile < p.total_tiles; tile += 2 * gridDim.x, it += 2) {
        const int stage = it % p.in_stages;
        const uint32_t phase = (uint32_t)(it / p.in_stages) & 1u, acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&tempty[mw], acc_phase ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t a0 = a_desc(smem_u32(sin + stage * p.in_stage_bytes), 16, p.variant & 1);
        const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (tap >= ntap) break;
            const uint32_t at = a_lo + (uint32_t)(((tap / 3) * p.WP + (tap % 3)) * 8);     // (r*WP+s)*128 B >> 4
            const uint32_t bt = b_lo + (uint32_t)(tap * 512);                               // tap*8192 B >> 4
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_lohi(d_tmem, at + 2 * k, a_hi, bt + 2 * k, b_hi, idesc, (tap | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          umma_commit(&tfull[mw]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue: 16 warps; warp e -> TMEM lane quarter (warp & 3), column quarter (e >> 2); four warps per
    // scheduler hide the TMEM-load / shared-memory latencies, so an epilogue pass costs less than the tile's MMAs
    const int e = warp - F_EPI_WARP0;
    const int q = warp & 3, cq = e >> 2;
    const int row = q * 32 + lane;                  // accumulator row = padded position hh*WP + ww
    const int hh = row / p.WP, ww = row - hh * p.WP;
    const bool row_ok = hh < p.R && ww < p.W;
    const int col0 = cq * 16;
    const int drow = hh * p.W + ww;                 // row of the dense [R*W][64] staging / operand tiles
    uint32_t soff[2];                               // 128B-swizzled byte offsets of this thread's two 16-byte chunks
#pragma unroll
    for (int j2 = 0; j2 < 2; ++j2) soff[j2] = (uint32_t)drow * 128u + ((uint32_t)((cq * 2 + j2) ^ (drow & 7)) << 4);
    const bool issuer = threadIdx.x == F_EPI_WARP0 * 32;
    // bias lives in shared memory (read as broadcast float4) to keep registers for the statistics
    float* bias_s = stats_smem + 16 * 2 * 16;       // [64], behind the [16][2][16] statistics block
    float* bnc = bias_s + 64;                       // [4][64]: scale, shift, mean, rstd
    if (e == 0) {
      bias_s[lane] = p.bias ? __ldg(p.bias + lane) : 0.f;
      bias_s[lane + 32] = p.bias ? __ldg(p.bias + lane + 32) : 0.f;
    }
    if (e == 1 && p.bn_bwd) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        bnc[c] = __ldg(p.bn_scale + c);
        bnc[64 + c] = __ldg(p.bn_shift + c);
        bnc[128 + c] = __ldg(p.bn_mean + c);
        bnc[192 + c] = __ldg(p.bn_rstd + c);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
    // BatchNorm statistics: per-thread running sums of the raw accumulators over all tiles of this CTA; the bias is
    // folded in analytically and the cross-row reduction runs once, after the last tile.
    float acc_s[16], acc_q[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc_s[j] = acc_q[j] = 0.f;
    int nvalid = 0;
    const bool want_stats = p.stats != nullptr;
    const bf16* g_add = (p.n_extra && p.extra_is_add) ? nullptr : p.add_src;    // operands still read from global
    const bf16* g_ref = (p.n_extra && !p.extra_is_add && !p.bn_bwd) ? nullptr : p.act_ref;
    const float neg = p.ref_act == ACT_LRELU ? p.ref_slope : 0.f;
    int acc = 0, it = 0;
    uint32_t acc_phase = 0;
    int n = blockIdx.x / p.tiles_per_img, tin = blockIdx.x % p.tiles_per_img;   // tile -> (image, row block), advanced incrementally
    const int dn = gridDim.x / p.tiles_per_img, dt = gridDim.x % p.tiles_per_img;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int h0 = tin * p.R;
      const int slot = it & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + (uint32_t(q * 32) << 16) + acc * 64 + col0, r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tempty[acc]);                    // accumulator is in registers: release TMEM early
      if (p.variant & 2) {                          // experiment: no epilogue work
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      const bool valid = row_ok && (h0 + hh) < p.H;
      float v[16];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 b = *reinterpret_cast<const float4*>(bias_s + col0 + j4 * 4);
        v[j4 * 4 + 0] = __uint_as_float(r[j4 * 4 + 0]) + b.x;
        v[j4 * 4 + 1] = __uint_as_float(r[j4 * 4 + 1]) + b.y;
        v[j4 * 4 + 2] = __uint_as_float(r[j4 * 4 + 2]) + b.z;
        v[j4 * 4 + 3] = __uint_as_float(r[j4 * 4 + 3]) + b.w;
      }
      if (p.act == ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
      } else if (p.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.n_extra) {
        mbar_wait(&xfull[slot], (it >> 1) & 1);
        if (valid) {
          const uint8_t* xt = sx + slot * p.out_tile_bytes;
#pragma unroll
          for (int j2 = 0; j2 < 2; ++j2) {
            const uint4 u = *reinterpret_cast<const uint4*>(xt + soff[j2]);
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
            if (p.bn_bwd) {
              // the operand is y of the BatchNorm this gradient enters next: g = v * act'(scale*y + shift);
              // accumulate sum g and sum g*(y - mean) (scaled by rstd once, at the end).  Constants come as
              // warp-broadcast 16-byte shared-memory loads (6 per 8 columns), not per element.
              float ca[8], cb[8], cm[8];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int c = col0 + j2 * 8 + h * 4;
                const float4 fa = *reinterpret_cast<const float4*>(bnc + c);
                const float4 fb = *reinterpret_cast<const float4*>(bnc + 64 + c);
                const float4 fm = *reinterpret_cast<const float4*>(bnc + 128 + c);
                ca[h * 4] = fa.x; ca[h * 4 + 1] = fa.y; ca[h * 4 + 2] = fa.z; ca[h * 4 + 3] = fa.w;
                cb[h * 4] = fb.x; cb[h * 4 + 1] = fb.y; cb[h * 4 + 2] = fb.z; cb[h * 4 + 3] = fb.w;
                cm[h * 4] = fm.x; cm[h * 4 + 1] = fm.y; cm[h * 4 + 2] = fm.z; cm[h * 4 + 3] = fm.w;
              }
              const float neg_bn = p.bn_act == ACT_LRELU ? p.bn_slope : (p.bn_act == ACT_RELU ? 0.f : 1.f);
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const float yv = (t & 1) ? __uint_as_float(w4[t >> 1] & 0xffff0000u) : __uint_as_float(w4[t >> 1] << 16);
                const int j = j2 * 8 + t;
                const float g = fmaf(yv, ca[t], cb[t]) > 0.f ? v[j] : v[j] * neg_bn;
                acc_s[j] += g;
                acc_q[j] = fmaf(g, yv - cm[t], acc_q[j]);
              }
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float lo = __uint_as_float(w4[t] << 16), hi = __uint_as_float(w4[t] & 0xffff0000u);
                if (p.extra_is_add) {
                  v[j2 * 8 + t * 2] += lo;
                  v[j2 * 8 + t * 2 + 1] += hi;
                } else {
                  v[j2 * 8 + t * 2] *= (lo > 0.f ? 1.f : neg);
                  v[j2 * 8 + t * 2 + 1] *= (hi > 0.f ? 1.f : neg);
                }
              }
            }
          }
        }
        mbar_arrive(&xempty[slot]);
      }
      if (valid && (g_add != nullptr || g_ref != nullptr)) {
        const long long pix = ((long long)n * p.H + h0 + hh) * p.W + ww;
        if (g_add != nullptr) {
          const uint4* src = reinterpret_cast<const uint4*>(g_add + pix * C64 + col0);
#pragma unroll
          for (int j2 = 0; j2 < 2; ++j2) {
            const uint4 u = __ldg(src + j2);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = __bfloat1622float2(h2[t]);
              v[j2 * 8 + t * 2] += f.x;
              v[j2 * 8 + t * 2 + 1] += f.y;
            }
          }
        }
        if (g_ref != nullptr) {
          const uint4* src = reinterpret_cast<const uint4*>(g_ref + pix * C64 + col0);
#pragma unroll
          for (int j2 = 0; j2 < 2; ++j2) {
            const uint4 u = __ldg(src + j2);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = __bfloat1622float2(h2[t]);
              v[j2 * 8 + t * 2] *= (f.x > 0.f ? 1.f : neg);
              v[j2 * 8 + t * 2 + 1] *= (f.y > 0.f ? 1.f : neg);
            }
          }
        }
      }
      // staging slot `slot` was last read by the TMA store issued two tiles ago
      if (issuer) tma_store_wait_read<1>();
      asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
      if (valid) {
        uint8_t* st = sout + slot * p.out_tile_bytes;
#pragma unroll
        for (int j2 = 0; j2 < 2; ++j2) {
          uint4 u;
          u.x = pack2(v[j2 * 8 + 0], v[j2 * 8 + 1]);
          u.y = pack2(v[j2 * 8 + 2], v[j2 * 8 + 3]);
          u.z = pack2(v[j2 * 8 + 4], v[j2 * 8 + 5]);
          u.w = pack2(v[j2 * 8 + 6], v[j2 * 8 + 7]);
          *reinterpret_cast<uint4*>(st + soff[j2]) = u;
        }
        if (want_stats && !p.bn_bwd) {   // forward statistics are only requested with act == NONE and no add/act_ref
          ++nvalid;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = __uint_as_float(r[j]);
            acc_s[j] += a;
            acc_q[j] = fmaf(a, a, acc_q[j]);
          }
        }
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
      if (issuer) {
        tma_store_4d(&tmOut, sout + slot * p.out_tile_bytes, 0, 0, h0, n);
        tma_store_commit();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      n += dn; tin += dt;
      if (tin >= p.tiles_per_img) { tin -= p.tiles_per_img; ++n; }
    }
    if (issuer) tma_store_wait<0>();
    if (want_stats) {
      // sum(a+b) = sum a + n b ; sum (a+b)^2 = sum a^2 + 2 b sum a + n b^2
      const float nv = (float)nvalid;
      if (p.bn_bwd) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          acc_q[j] *= bnc[192 + col0 + j] * p.bn_gscale;      // sum g*(y - mean) -> sum g*xhat
          acc_s[j] *= p.bn_gscale;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float b = bias_s[col0 + j], sa = acc_s[j];
          acc_q[j] = acc_q[j] + 2.f * b * sa + nv * b * b;
          acc_s[j] = sa + nv * b;
        }
      }
      const float ts = colsum16(acc_s, lane);      // lanes l, l^16 hold column col0 + (l & 15) over this warp's 32 rows
      const float tq = colsum16(acc_q, lane);
      // stats_smem[e][2][16]: per epilogue warp (sum, sumsq) of its 16 columns; fixed-order add over the 4 lane quarters
      if (lane < 16) {
        stats_smem[(e * 2 + 0) * 16 + lane] = ts;
        stats_smem[(e * 2 + 1) * 16 + lane] = tq;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
      const int t = threadIdx.x - F_EPI_WARP0 * 32;   // 0..511
      if (t < 128) {
        const int which = t >> 6, col = t & 63;    // which: 0 sum, 1 sumsq
        const int c4 = col >> 4, l = col & 15;
        float s = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) s += stats_smem[(((c4 * 4 + qq) * 2) + which) * 16 + l];
        p.stats[(size_t)blockIdx.x * 128 + which * 64 + col] = s;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------
// Row-class stacked variant (H % 4 == 0).  The M=128 x N=64 MMAs above are bound by the A-operand read (a 4 KB pixel
// block streamed from shared memory per instruction, ~64 cycles against 32 cycles of math), so the lever is to use
// every A read for more output columns.  Image rows are split into four classes h % 4 (a 5-D tensor-map view of the same
// NHWC tensor: no data movement); a super-tile is R consecutive rows of EACH class of one image and its accumulator is
// M=128 positions x N=256 = four 64-channel blocks, one per output class.  An input block of class c (same row index)
// is the tap-row  r = c - co + 1  operand of the output classes co = c-1, c, c+1, so ONE MMA with the three tap
// matrices stacked along N (resident weights ordered [s][r = 2, 1, 0]) feeds three output blocks:
//     class 1 -> co 0,1,2 (N=192)    class 2 -> co 1,2,3 (N=192)    class 0 -> co 0,1 (N=128)    class 3 -> co 2,3 (N=128)
// plus the two wrap-arounds, class 0 one row further down -> co 3 (r = 2) and class 3 one row further up -> co 0 (r = 0), N=64.
// 72 MMAs per 512 positions instead of 144.  The epilogue is the one above, run once per output class.
// ------------------------------------------------------------------------------------------
// Epilogue specialisations (SPEC >= 0): the epilogue of the fused launches is bound by instruction issue (sixteen warps
// x ~250 instructions per sub-tile against the ~4000 cycles of a super-tile's MMAs), and a good part of those
// instructions only test launch-uniform switches.  The launches the MNIST step makes fix the switches at compile time;
// SPEC < 0 reads every switch from the parameters (any other caller, every experiment variant).  A fixed SPEC implies:
// act == NONE, no operand read from global memory, no experiment variant, last-CTA finalisation only with SP_FIN.
enum : int { SP_BIAS = 1, SP_BNBWD = 2, SP_DUAL = 4, SP_BN_NOACT = 8, SP_EXTRA = 16, SP_STATS = 32,
              SP_F2 = 64, SP_FIN = 128,     // SP_FIN: the last CTA finishes the statistics (StatsFinalize)
              SP_LRELU = 256,               // act == LRELU (conv_mid, folded-BatchNorm inference)
              SP_EXTRA_ADD = 512 };         // the TMA operand is the residual (folded-BatchNorm inference)    // SP_F2: the sums and the residual add as packed two-lane fp32 instructions (FADD2 / FFMA2:
                               // the same round-to-nearest results, half the issue slots)
constexpr int SPEC_FWD_STATS = SP_BIAS | SP_STATS;                                    // conv + bias, BatchNorm statistics
constexpr int SPEC_BN1 = SP_BNBWD | SP_EXTRA | SP_STATS;                              // data gradient + BN1 backward sums
constexpr int SPEC_BN2 = SP_BNBWD | SP_BN_NOACT | SP_EXTRA | SP_STATS;                // ... + BN2 backward sums
constexpr int SPEC_BN2_DUAL = SPEC_BN2 | SP_DUAL;                                     // ... and the skip gradient added
constexpr int SPEC_FWD_LRELU = SP_BIAS | SP_LRELU;                                    // conv + bias + LeakyReLU
constexpr int SPEC_FWD_ADD = SP_BIAS | SP_EXTRA | SP_EXTRA_ADD;                       // conv + bias + residual

__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  const float2 r = __fadd2_rn(make_float2(a0, a1), make_float2(b0, b1));
  a0 = r.x; a1 = r.y;
}
// (c0, c1) = (a0 * b0 + c0, a1 * b1 + c1)
__device__ __forceinline__ void fma2(float& c0, float& c1, float a0, float a1, float b0, float b1) {
  const float2 r = __ffma2_rn(make_float2(a0, a1), make_float2(b0, b1), make_float2(c0, c1));
  c0 = r.x; c1 = r.y;
}

template <int SPEC>
__global__ void __launch_bounds__(F_THREADS, 1)
conv_tc64s_fprop_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                        const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmExtra,
                        const __grid_constant__ CUtensorMap tmAdd, const F64Params p) {
  pdl_launch_dependents();                        // the prologue below touches no global memory: see pdl_wait() further down
  constexpr bool kS = SPEC >= 0;
  const bool k_bias = kS ? (SPEC & SP_BIAS) != 0 : true;       // run-time: a missing bias is added as zeros
  const bool k_bn_bwd = kS ? (SPEC & SP_BNBWD) != 0 : p.bn_bwd != 0;
  const bool k_dual = kS ? (SPEC & SP_DUAL) != 0 : p.dual != 0;
  const bool k_bn_noact = kS ? (SPEC & SP_BN_NOACT) != 0 : p.bn_act == ACT_NONE;
  const bool k_extra = kS ? (SPEC & SP_EXTRA) != 0 : p.n_extra != 0;
  const bool k_extra_add = kS ? (SPEC & SP_EXTRA_ADD) != 0 : p.extra_is_add != 0;
  const bool k_stats = kS ? (SPEC & SP_STATS) != 0 : p.stats != nullptr;
  const int k_act = kS ? ((SPEC & SP_LRELU) ? (int)ACT_LRELU : (int)ACT_NONE) : p.act;
  const int k_variant = kS ? 0 : p.variant;
  const bool k_fin = kS ? (SPEC & SP_FIN) != 0 : p.fin.mode != 0;
  constexpr bool kF2 = kS && (SPEC & SP_F2) != 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sw = smem;                                   // weights, slot s*3 + (2 - r)
  uint8_t* sin = smem + W_BYTES;                        // ring of class regions ((R+1) x (W+2) pixels each)
  // Dual mode (fused BatchNorm-backward reduction AND a residual add: two epilogue operands): the residual tile is
  // loaded by TMA straight into the staging slot its result will leave from - every thread reads its 32 bytes, adds,
  // and writes the result back in place - so the slot ring (three deep) replaces a second operand ring that would not
  // fit: load (afull) -> epilogue in place (sfull) -> TMA store -> read out (sempty) -> next load.
  const int nst = k_dual ? 3 : 2;
  uint8_t* sout = sin + p.in_stages * p.in_stage_bytes; // output staging, 2 slots (dual: 3 in-place slots)
  uint8_t* sx = sout + nst * p.out_tile_bytes;          // epilogue operand ring, 2 slots (if n_extra)
  float* stats_smem = reinterpret_cast<float*>(sx + (k_extra ? 2 : 0) * p.out_tile_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats_smem) + (16 * 2 * 16 + 64 + 4 * 64) * 4);
  uint64_t* full = bars;                 // [<= 8] ring of class regions
  uint64_t* empty = bars + 8;            // [<= 8]
  uint64_t* wfull = bars + 16;           // [1]
  uint64_t* tfull = wfull + 1;           // [2][4] accumulator, output class: committed as soon as the class is complete
  uint64_t* tempty = tfull + 8;          // [2]
  uint64_t* xfull = tempty + 2;          // [2]
  uint64_t* xempty = xfull + 2;          // [2]
  uint64_t* sfull = xempty + 2;          // [3] output staging slot written by all epilogue threads
  uint64_t* sempty = sfull + 3;          // [3] ... and read out by its TMA store
  uint64_t* afull = sempty + 3;          // [3] dual: residual tile landed in the staging slot
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(afull + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmOut);
    if (k_extra) tma_prefetch_desc(&tmExtra);
    if (k_dual) tma_prefetch_desc(&tmAdd);
    for (int s = 0; s < 8; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&tfull[s], 1); }
    mbar_init(wfull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tempty[s], F_EPI_THREADS);
      mbar_init(&xfull[s], 1); mbar_init(&xempty[s], F_EPI_THREADS);
    }
    for (int s = 0; s < 3; ++s) { mbar_init(&sfull[s], F_EPI_THREADS); mbar_init(&sempty[s], 1); mbar_init(&afull[s], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  // rows of a class region beyond the TMA box are read by the (discarded) padding rows of the MMA: keep them finite
  {
    const int zbytes = (int)(reinterpret_cast<uint8_t*>(stats_smem) - sin);
    for (int i = threadIdx.x; i < zbytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sin)[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                     // from here on the predecessor kernel has completed and flushed
  const int box_bytes = (p.R + 1) * p.WP * 128;
  const int tile_bytes = p.R * p.W * 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, W_BYTES);
      for (int tap = 0; tap < 9; ++tap)
        tma_load_2d(&tmW, wfull, sw + ((tap % 3) * 3 + (2 - tap / 3)) * 8192, tap * 64, 0);
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t in_pol = l2_evict_first_policy();
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n = tile / p.tiles_per_img, i0 = (tile % p.tiles_per_img) * p.R;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cls = j == 0 ? 1 : (j == 1 ? 0 : j);          // class order 1, 0, 2, 3 (first touches of the accumulator)
          mbar_wait(&empty[stage], phase ^ 1);
          if (k_variant & 16) {                      // experiment: no input traffic
            mbar_arrive(&full[stage]);
          } else {
            mbar_expect_tx(&full[stage], box_bytes);
            if (p.in_evict_first)
              tma_load_5d_hint(&tmX, &full[stage], sin + stage * p.in_stage_bytes, 0, -1, cls, cls == 3 ? i0 - 1 : i0, n, in_pol);
            else
              tma_load_5d(&tmX, &full[stage], sin + stage * p.in_stage_bytes, 0, -1, cls, cls == 3 ? i0 - 1 : i0, n);
          }
          if (++stage == p.in_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == F_EPI_WARP0 + F_EPI_WARPS) {
    if (lane == 0 && k_extra) {
      int sub = 0, s3 = 0;
      uint32_t s3use = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n = tile / p.tiles_per_img, i0 = (tile % p.tiles_per_img) * p.R;
        for (int c = 0; c < 4; ++c, ++sub) {
          const int co = c < 2 ? 1 - c : c;                        // the epilogue's class order 1, 0, 2, 3
          const int slot = sub & 1;
          if (k_dual) {                                            // residual tile into the next free in-place slot
            mbar_wait(&sempty[s3], (s3use & 1u) ^ 1u);
            mbar_expect_tx(&afull[s3], tile_bytes);
            if (p.add_evict_first) tma_load_5d_hint(&tmAdd, &afull[s3], sout + s3 * p.out_tile_bytes, 0, 0, co, i0, n, l2_evict_first_policy());
            else tma_load_5d(&tmAdd, &afull[s3], sout + s3 * p.out_tile_bytes, 0, 0, co, i0, n);
            if (++s3 == 3) { s3 = 0; ++s3use; }
          }
          mbar_wait(&xempty[slot], ((sub >> 1) & 1) ^ 1);
          mbar_expect_tx(&xfull[slot], tile_bytes);
          tma_load_5d(&tmExtra, &xfull[slot], sx + slot * p.out_tile_bytes, 0, 0, co, i0, n);
        }
      }
    }
  } else if (warp == 2) {
    // Store warp: the epilogue warps never meet at a CTA-wide barrier; each thread writes its 32 bytes of the staging
    // slot, fences and arrives on sfull, this thread sends the slot off and hands it back through sempty once the
    // TMA engine has read it, so the sixteen epilogue warps drift apart and hide each other's latencies.
    if (lane == 0 && !(k_variant & 2)) {
      int sub = 0, slot = 0, prev = 0;                // slot = sub % nst, use = sub / nst
      uint32_t use = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n = tile / p.tiles_per_img, i0 = (tile % p.tiles_per_img) * p.R;
        for (int c = 0; c < 4; ++c, ++sub) {
          const int co = c < 2 ? 1 - c : c;
          mbar_wait(&sfull[slot], use & 1u);
          tma_store_5d(&tmOut, sout + slot * p.out_tile_bytes, 0, 0, co, i0, n);
          tma_store_commit();
          if (sub >= 1) {
            tma_store_wait_read<1>();                 // the previous store has read its slot
            mbar_arrive(&sempty[prev]);
          }
          prev = slot;
          if (++slot == nst) { slot = 0; ++use; }
        }
      }
      tma_store_wait<0>();
    }
  } else if (warp == 1) {
    // one MMA-issuing warp (the stacked MMAs last ~100 cycles, more than their issue cost): even super-tiles of this
    // CTA into accumulator 0 (TMEM columns 0-255), odd ones into accumulator 1
    constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64, 0, 0), idesc128 = umma_idesc_bf16(128, 128, 0, 0),
                       idesc192 = umma_idesc_bf16(128, 192, 0, 0);
    mbar_wait(wfull, 0);
    const uint64_t b0 = umma_smem_desc(smem_u32(sw), 16, 1024);
    const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
    const uint32_t row_adv = (uint32_t)(p.WP * 8);          // one region row further down: WP * 128 B >> 4
    int it = 0, stage = 0;                                  // CTA-local super-tile counter, ring position
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int mw = it & 1;
      const uint32_t d_tmem = tmem_base + mw * 256;
      mbar_wait(&tempty[mw], ((uint32_t)(it >> 1) & 1u) ^ 1u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t a0 = a_desc(smem_u32(sin + stage * p.in_stage_bytes), 16, 0);
        const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
        if (elect_one()) {
          if (j == 0) {                                     // class 1 -> co 0,1,2
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_lohi(d_tmem, a_lo + s * 8 + 2 * k, a_hi, b_lo + s * 1536 + 2 * k, b_hi, idesc192, (s | k) != 0 ? 1u : 0u);
          } else if (k_variant & 8) {                       // experiment: 12 of the 72 MMAs
          } else if (j == 1) {                              // class 0: one row down -> co 3 (r = 2); same row -> co 0,1 (r = 1, 0)
            if (!(k_variant & 4))                           // experiment bit 4: no wrap-around MMAs
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_lohi(d_tmem + 192, a_lo + row_adv + s * 8 + 2 * k, a_hi, b_lo + s * 1536 + 2 * k, b_hi, idesc64,
                              (s | k) != 0 ? 1u : 0u);
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_lohi(d_tmem, a_lo + s * 8 + 2 * k, a_hi, b_lo + s * 1536 + 512 + 2 * k, b_hi, idesc128, 1u);
          } else if (j == 2) {                              // class 2 -> co 1,2,3
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_lohi(d_tmem + 64, a_lo + s * 8 + 2 * k, a_hi, b_lo + s * 1536 + 2 * k, b_hi, idesc192, 1u);
          } else {                                          // class 3 (region starts one row up): one row up -> co 0 (r = 0); same row -> co 2,3
            if (!(k_variant & 4))
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_lohi(d_tmem, a_lo + s * 8 + 2 * k, a_hi, b_lo + s * 1536 + 1024 + 2 * k, b_hi, idesc64, 1u);
            umma_commit(&tfull[mw * 4 + 0]);                // class 0 is complete
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_lohi(d_tmem + 128, a_lo + row_adv + s * 8 + 2 * k, a_hi, b_lo + s * 1536 + 2 * k, b_hi, idesc128, 1u);
          }
          umma_commit(&empty[stage]);
          if (j == 2) umma_commit(&tfull[mw * 4 + 1]);      // class 1 is complete after the class-2 input block
          if (j == 3) { umma_commit(&tfull[mw * 4 + 2]); umma_commit(&tfull[mw * 4 + 3]); }
        }
        __syncwarp();
        if (++stage == p.in_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ---- epilogue: as in conv_tc64_fprop_kernel, one pass per output class (TMEM columns acc*256 + co*64 + ...)
    const int e = warp - F_EPI_WARP0;
    const int q = warp & 3, cq = e >> 2;
    const int row = q * 32 + lane;                  // accumulator row = padded position hh*WP + ww
    const int hh = row / p.WP, ww = row - hh * p.WP;
    const bool row_ok = hh < p.R && ww < p.W;
    const int col0 = cq * 16;
    const int drow = hh * p.W + ww;
    uint32_t soff[2];
#pragma unroll
    for (int j2 = 0; j2 < 2; ++j2) soff[j2] = (uint32_t)drow * 128u + ((uint32_t)((cq * 2 + j2) ^ (drow & 7)) << 4);
    float* bias_s = stats_smem + 16 * 2 * 16;
    float* bnc = bias_s + 64;
    if (e == 0 && k_bias) {
      bias_s[lane] = p.bias ? __ldg(p.bias + lane) : 0.f;
      bias_s[lane + 32] = p.bias ? __ldg(p.bias + lane + 32) : 0.f;
    }
    if (e == 1 && k_bn_bwd) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        bnc[c] = __ldg(p.bn_scale + c);
        bnc[64 + c] = __ldg(p.bn_shift + c);
        bnc[128 + c] = __ldg(p.bn_mean + c);
        bnc[192 + c] = __ldg(p.bn_rstd + c);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
    float acc_s[16], acc_q[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc_s[j] = acc_q[j] = 0.f;
    int nvalid = 0;
    const bool want_stats = k_stats;
    const bf16* g_add = (kS || (k_extra && k_extra_add)) ? nullptr : p.add_src;
    const bf16* g_ref = (kS || (k_extra && !k_extra_add && !k_bn_bwd)) ? nullptr : p.act_ref;
    const float neg = p.ref_act == ACT_LRELU ? p.ref_slope : 0.f;
    const int nidx = p.H >> 2;                      // rows per class
    int acc = 0, sub = 0, sslot = 0;                // sslot = sub % nst: staging slot
    uint32_t acc_phase = 0, suse = 0;               // suse = sub / nst
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n = tile / p.tiles_per_img, i0 = (tile % p.tiles_per_img) * p.R;
      const bool valid = row_ok && (i0 + hh) < nidx;
#pragma unroll 1
      for (int c = 0; c < 4; ++c, ++sub) {
        const int co = c < 2 ? 1 - c : c;           // classes in the order the MMA warp completes them: 1, 0, 2, 3
        const int slot = sub & 1;
        mbar_wait(&tfull[acc * 4 + co], acc_phase);
        tc_fence_after();
        uint32_t r[16];
        tmem_ld_32x16(tmem_base + (uint32_t(q * 32) << 16) + acc * 256 + co * 64 + col0, r);
        tmem_ld_wait();
        if (co == 3) {
          tc_fence_before();
          mbar_arrive(&tempty[acc]);                // the last block is in registers: release the accumulator
        }
        if (k_variant & 2) {                        // experiment: no epilogue work
          if (++sslot == nst) { sslot = 0; ++suse; }
          continue;
        }
        float v[16];
        if (k_bias) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = *reinterpret_cast<const float4*>(bias_s + col0 + j4 * 4);
            if constexpr (kF2) {
              v[j4 * 4 + 0] = __uint_as_float(r[j4 * 4 + 0]); v[j4 * 4 + 1] = __uint_as_float(r[j4 * 4 + 1]);
              v[j4 * 4 + 2] = __uint_as_float(r[j4 * 4 + 2]); v[j4 * 4 + 3] = __uint_as_float(r[j4 * 4 + 3]);
              add2(v[j4 * 4 + 0], v[j4 * 4 + 1], b.x, b.y);
              add2(v[j4 * 4 + 2], v[j4 * 4 + 3], b.z, b.w);
            } else {
              v[j4 * 4 + 0] = __uint_as_float(r[j4 * 4 + 0]) + b.x;
              v[j4 * 4 + 1] = __uint_as_float(r[j4 * 4 + 1]) + b.y;
              v[j4 * 4 + 2] = __uint_as_float(r[j4 * 4 + 2]) + b.z;
              v[j4 * 4 + 3] = __uint_as_float(r[j4 * 4 + 3]) + b.w;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        }
        if (k_act == ACT_LRELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
        } else if (k_act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (k_dual) {
          // the residual tile sits in the staging slot this pass will write its result to
          mbar_wait(&afull[sslot], suse & 1u);
          if (valid) {
            const uint8_t* at = sout + sslot * p.out_tile_bytes;
#pragma unroll
            for (int j2 = 0; j2 < 2; ++j2) {
              const uint4 u = *reinterpret_cast<const uint4*>(at + soff[j2]);
              const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int j = j2 * 8 + t * 2;
                if constexpr (kF2) {
                  add2(v[j], v[j + 1], __uint_as_float(w4[t] << 16), __uint_as_float(w4[t] & 0xffff0000u));
                } else {
                  v[j] += __uint_as_float(w4[t] << 16);
                  v[j + 1] += __uint_as_float(w4[t] & 0xffff0000u);
                }
              }
            }
          }
        }
        if (k_extra) {
          mbar_wait(&xfull[slot], (sub >> 1) & 1);
          if (valid) {
            const uint8_t* xt = sx + slot * p.out_tile_bytes;
#pragma unroll
            for (int j2 = 0; j2 < 2; ++j2) {
              const uint4 u = *reinterpret_cast<const uint4*>(xt + soff[j2]);
              const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
              if (k_bn_bwd && k_bn_noact) {
                // no activation between the BatchNorm and this gradient (BN2 of a residual block): g = v.  The second
                // sum is taken as sum g*y and turned into sum g*(y - mean) after the last tile (acc_q -= mean * acc_s)
                if constexpr (kF2) {
#pragma unroll
                  for (int t = 0; t < 4; ++t) {
                    const int j = j2 * 8 + t * 2;
                    add2(acc_s[j], acc_s[j + 1], v[j], v[j + 1]);
                    fma2(acc_q[j], acc_q[j + 1], v[j], v[j + 1], __uint_as_float(w4[t] << 16),
                         __uint_as_float(w4[t] & 0xffff0000u));
                  }
                } else {
#pragma unroll
                  for (int t = 0; t < 8; ++t) {
                    const float yv = (t & 1) ? __uint_as_float(w4[t >> 1] & 0xffff0000u) : __uint_as_float(w4[t >> 1] << 16);
                    const int j = j2 * 8 + t;
                    acc_s[j] += v[j];
                    acc_q[j] = fmaf(v[j], yv, acc_q[j]);
                  }
                }
              } else if (k_bn_bwd) {
                float ca[8], cb[8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int c = col0 + j2 * 8 + h * 4;
                  const float4 fa = *reinterpret_cast<const float4*>(bnc + c);
                  const float4 fb = *reinterpret_cast<const float4*>(bnc + 64 + c);
                  ca[h * 4] = fa.x; ca[h * 4 + 1] = fa.y; ca[h * 4 + 2] = fa.z; ca[h * 4 + 3] = fa.w;
                  cb[h * 4] = fb.x; cb[h * 4 + 1] = fb.y; cb[h * 4 + 2] = fb.z; cb[h * 4 + 3] = fb.w;
                }
                const float neg_bn = p.bn_act == ACT_LRELU ? p.bn_slope : 0.f;
                if constexpr (kF2) {
#pragma unroll
                  for (int t = 0; t < 4; ++t) {
                    const int j = j2 * 8 + t * 2;
                    const float y0 = __uint_as_float(w4[t] << 16), y1 = __uint_as_float(w4[t] & 0xffff0000u);
                    float z0 = cb[t * 2], z1 = cb[t * 2 + 1];
                    fma2(z0, z1, y0, y1, ca[t * 2], ca[t * 2 + 1]);
                    const float g0 = z0 > 0.f ? v[j] : v[j] * neg_bn, g1 = z1 > 0.f ? v[j + 1] : v[j + 1] * neg_bn;
                    add2(acc_s[j], acc_s[j + 1], g0, g1);
                    fma2(acc_q[j], acc_q[j + 1], g0, g1, y0, y1);
                  }
                } else {
#pragma unroll
                  for (int t = 0; t < 8; ++t) {
                    const float yv = (t & 1) ? __uint_as_float(w4[t >> 1] & 0xffff0000u) : __uint_as_float(w4[t >> 1] << 16);
                    const int j = j2 * 8 + t;
                    const float g = fmaf(yv, ca[t], cb[t]) > 0.f ? v[j] : v[j] * neg_bn;
                    acc_s[j] += g;
                    acc_q[j] = fmaf(g, yv, acc_q[j]);
                  }
                }
              } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float lo = __uint_as_float(w4[t] << 16), hi = __uint_as_float(w4[t] & 0xffff0000u);
                  if (k_extra_add) {
                    v[j2 * 8 + t * 2] += lo;
                    v[j2 * 8 + t * 2 + 1] += hi;
                  } else {
                    v[j2 * 8 + t * 2] *= (lo > 0.f ? 1.f : neg);
                    v[j2 * 8 + t * 2 + 1] *= (hi > 0.f ? 1.f : neg);
                  }
                }
              }
            }
          }
          mbar_arrive(&xempty[slot]);
        }
        if (valid && (g_add != nullptr || g_ref != nullptr)) {
          const long long pix = ((long long)n * p.H + 4 * (i0 + hh) + co) * p.W + ww;
          if (g_add != nullptr) {
            const uint4* src = reinterpret_cast<const uint4*>(g_add + pix * C64 + col0);
#pragma unroll
            for (int j2 = 0; j2 < 2; ++j2) {
              const uint4 u = __ldg(src + j2);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = __bfloat1622float2(h2[t]);
                v[j2 * 8 + t * 2] += f.x;
                v[j2 * 8 + t * 2 + 1] += f.y;
              }
            }
          }
          if (g_ref != nullptr) {
            const uint4* src = reinterpret_cast<const uint4*>(g_ref + pix * C64 + col0);
#pragma unroll
            for (int j2 = 0; j2 < 2; ++j2) {
              const uint4 u = __ldg(src + j2);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = __bfloat1622float2(h2[t]);
                v[j2 * 8 + t * 2] *= (f.x > 0.f ? 1.f : neg);
                v[j2 * 8 + t * 2 + 1] *= (f.y > 0.f ? 1.f : neg);
              }
            }
          }
        }
        // staging slot: last read by the TMA store issued two passes ago (dual: already owned, the residual came in it)
        if (!k_dual) mbar_wait(&sempty[sslot], (suse & 1u) ^ 1u);
        if (valid) {
          uint8_t* st = sout + sslot * p.out_tile_bytes;
#pragma unroll
          for (int j2 = 0; j2 < 2; ++j2) {
            uint4 u;
            u.x = pack2(v[j2 * 8 + 0], v[j2 * 8 + 1]);
            u.y = pack2(v[j2 * 8 + 2], v[j2 * 8 + 3]);
            u.z = pack2(v[j2 * 8 + 4], v[j2 * 8 + 5]);
            u.w = pack2(v[j2 * 8 + 6], v[j2 * 8 + 7]);
            *reinterpret_cast<uint4*>(st + soff[j2]) = u;
          }
          if (want_stats && !k_bn_bwd) {
            ++nvalid;
            if constexpr (kF2) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float a0 = __uint_as_float(r[j]), a1 = __uint_as_float(r[j + 1]);
                add2(acc_s[j], acc_s[j + 1], a0, a1);
                fma2(acc_q[j], acc_q[j + 1], a0, a1, a0, a1);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float a = __uint_as_float(r[j]);
                acc_s[j] += a;
                acc_q[j] = fmaf(a, a, acc_q[j]);
              }
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&sfull[sslot]);
        if (++sslot == nst) { sslot = 0; ++suse; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (want_stats) {
      const float nv = (float)nvalid;
      if (k_bn_bwd) {
        // sum g*y -> sum g*(y - mean) * rstd = sum g*xhat; both sums scaled by bn_gscale
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          acc_q[j] = (acc_q[j] - bnc[128 + col0 + j] * acc_s[j]) * bnc[192 + col0 + j] * p.bn_gscale;
          acc_s[j] *= p.bn_gscale;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float b = bias_s[col0 + j], sa = acc_s[j];
          acc_q[j] = acc_q[j] + 2.f * b * sa + nv * b * b;
          acc_s[j] = sa + nv * b;
        }
      }
      const float ts = colsum16(acc_s, lane);
      const float tq = colsum16(acc_q, lane);
      if (lane < 16) {
        stats_smem[(e * 2 + 0) * 16 + lane] = ts;
        stats_smem[(e * 2 + 1) * 16 + lane] = tq;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
      const int t = threadIdx.x - F_EPI_WARP0 * 32;
      if (t < 128) {
        const int which = t >> 6, col = t & 63;
        const int c4 = col >> 4, l = col & 15;
        float s = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) s += stats_smem[(((c4 * 4 + qq) * 2) + which) * 16 + l];
        p.stats[(size_t)blockIdx.x * 128 + which * 64 + col] = s;
      }
      if (k_fin) {
        // the last CTA to get here finishes the statistics (fixed CTA order: deterministic)
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
        uint32_t* flag = reinterpret_cast<uint32_t*>(stats_smem);       // the partial sums in stats_smem are consumed
        if (t == 0) *flag = atomicAdd(p.fin.counter, 1u) == gridDim.x - 1 ? 1u : 0u;
        asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
        const bool last = *flag != 0u;
        asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
        if (last) {
          __threadfence();
          // scratch: the input ring (every MMA of this CTA has completed, nothing reads it any more)
          finalize_stats_last_cta(p.fin, p.stats, (int)gridDim.x, reinterpret_cast<double*>(sin), t);
          if (t == 0) *p.fin.counter = 0u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// The compile-time epilogue of a launch, or -1 when its switches are not one of the instantiated combinations.
// PCG_TC64_SPEC=0 keeps every launch on the run-time epilogue (A/B switch).
static int epilogue_spec(const F64Params& p) {
  static const int mode = [] { const char* e = getenv("PCG_TC64_SPEC"); return e == nullptr ? 3 : atoi(e); }();
  if (mode == 0 || (p.variant & ~(1024 | 2048)) != 0 || p.act == ACT_RELU) return -1;   // 1024, 2048: weight-gradient bits
  if (p.act_ref != nullptr) return -1;
  if (p.add_src != nullptr && !(p.n_extra && p.extra_is_add)) return -1;   // (dual: the host has moved add_src to the TMA map)
  int s = 0;
  if (p.act == ACT_LRELU) s |= SP_LRELU;
  if (p.n_extra && p.extra_is_add) s |= SP_EXTRA_ADD;
  if (p.bias != nullptr) s |= SP_BIAS;
  if (p.stats != nullptr) s |= SP_STATS;
  if (p.n_extra) s |= SP_EXTRA;
  if (p.bn_bwd) s |= SP_BNBWD | (p.bn_act == ACT_NONE ? SP_BN_NOACT : 0);
  if (p.dual) s |= SP_DUAL;
  if (p.fin.mode != 0) return (s == SPEC_FWD_STATS && mode >= 3) ? (s | SP_F2 | SP_FIN) : -1;
  if (s == SPEC_FWD_STATS) return mode >= 3 ? (s | SP_F2) : s;
  if (s == SPEC_FWD_LRELU || s == SPEC_FWD_ADD) return s;
  if (s == SPEC_BN1 || s == SPEC_BN2 || s == SPEC_BN2_DUAL) return mode >= 2 ? (s | SP_F2) : s;
  return -1;
}

// The stacked kernel is the default wherever it applies (H % 4 == 0); variant bit 256 selects the one-class-per-tile
// kernel.  Measured at B=512, 28x28 (profiles/exp_tc64_stacked_r1.md): forward + statistics 34.1-35.4 us against
// 35.9-36.0 us, data gradient + skip add 37.3 against 37.4 us, whole MNIST step 3.207 against 3.249 ms.  Both kernels
// sit on the shared-memory bandwidth of the SM (operand reads of the MMAs + staging + TMA fills: ~790 KB against
// ~1090 KB per 512 positions), not on the MMA count.
static bool use_stacked(int H, int W) { return (g_variant & 256) == 0 && H % 4 == 0 && conv_tc64_supported(H, W); }

int conv_tc64_fprop_grid(int N, int H, int W) {
  if (!use_stacked(H, W)) return conv_tc64_grid(N, H, W);
  const int R = 128 / (W + 2);
  const long long tiles = (long long)N * ((H / 4 + R - 1) / R);
  const int sms = sm_count();
  return (int)(tiles < sms ? tiles : sms);
}

void conv_tc64_fprop(const bf16* in, int N, int H, int W, const bf16* wpk, const ConvEpilogue& epi, bf16* out,
                     cudaStream_t stream) {
  PCG_PROFILE("conv_tc64_fprop", stream);
  PCG_REQUIRE(conv_tc64_supported(H, W), "image too wide for the halo-tile kernel");
  F64Params p;
  p.N = N; p.H = H; p.W = W; p.WP = W + 2; p.R = 128 / p.WP;
  p.tiles_per_img = (H + p.R - 1) / p.R;
  p.total_tiles = N * p.tiles_per_img;
  p.bias = epi.bias; p.act = epi.act; p.slope = epi.slope; p.add_src = epi.add_src;
  p.act_ref = epi.ref_act != ACT_NONE ? epi.act_ref : nullptr; p.ref_act = epi.ref_act; p.ref_slope = epi.ref_slope;
  p.stats = epi.stats; p.variant = g_variant; p.add_evict_first = (g_l2_hints & 4) ? 1 : 0;
  p.in_evict_first = ((g_l2_hints & 16) && epi.stats != nullptr && epi.bn_y == nullptr) ? 1 : 0;
  p.fin = epi.fin;
  PCG_REQUIRE(p.fin.mode == 0 || (epi.stats != nullptr && p.fin.counter != nullptr && p.fin.M > 0),
              "statistics finalisation needs the partial buffer, a ticket counter and the element count");
  PCG_REQUIRE(epi.stats == nullptr || epi.bn_y != nullptr ||
                  (epi.act == ACT_NONE && epi.add_src == nullptr && p.act_ref == nullptr),
              "BatchNorm statistics are taken of (accumulator + bias) only");
  p.bn_bwd = epi.bn_y != nullptr ? 1 : 0;
  p.bn_gscale = epi.bn_gscale;
  const bool stacked = use_stacked(H, W);
  p.dual = (p.bn_bwd && p.add_src != nullptr) ? 1 : 0;
  PCG_REQUIRE(!p.dual || stacked, "reduction + residual in one epilogue needs the row-class kernel (H % 4 == 0)");
  p.bn_mean = epi.bn_mean; p.bn_rstd = epi.bn_rstd; p.bn_scale = epi.bn_scale; p.bn_shift = epi.bn_shift;
  p.bn_act = epi.bn_act; p.bn_slope = epi.bn_slope;
  PCG_REQUIRE(!p.bn_bwd || (epi.stats != nullptr && epi.bias == nullptr && epi.act == ACT_NONE &&
                            p.act_ref == nullptr && epi.bn_mean && epi.bn_rstd && epi.bn_scale && epi.bn_shift),
              "fused BatchNorm-backward reduction: partial buffer and the four per-channel vectors are required");
  // one epilogue operand travels by TMA (the BatchNorm input, else the residual if there is one, else the activation
  // reference)
  p.n_extra = (p.bn_bwd || p.add_src != nullptr || p.act_ref != nullptr) ? 1 : 0;
  p.extra_is_add = (!p.bn_bwd && p.add_src != nullptr) ? 1 : 0;
  const bf16* extra = p.bn_bwd ? epi.bn_y : (p.add_src != nullptr ? p.add_src : p.act_ref);
  auto round1k = [](int b) { return (b + 1023) / 1024 * 1024; };
  const bf16* add_tma = p.dual ? p.add_src : nullptr;
  if (p.dual) p.add_src = nullptr;                 // the kernel takes it from the staging slot, never from global
  if (stacked) {                                   // a tile is a super-tile: R rows of each of the four row classes
    p.tiles_per_img = (H / 4 + p.R - 1) / p.R;
    p.total_tiles = N * p.tiles_per_img;
  }
  // stacked: a class region holds R+1 rows, and the last view starts WP + 2 rows in and spans 128 rows
  p.in_stage_bytes = stacked ? round1k((128 + p.WP + 2) * 128) : round1k((p.R + 2) * p.WP * 128);
  p.out_tile_bytes = round1k(p.R * p.W * 128);
  p.in_stages = stacked ? 5 : 4;
  auto total = [&]() {
    return 1024 + W_BYTES + p.in_stages * p.in_stage_bytes + ((p.dual ? 3 : 2) + 2 * p.n_extra) * p.out_tile_bytes + F_TAIL_BYTES;
  };
  while (total() > SMEM_LIMIT && p.in_stages > 2) --p.in_stages;
  PCG_REQUIRE(total() <= SMEM_LIMIT, "halo-tile kernel: shared-memory budget exceeded");
  CUtensorMap tmW = make_tmap_2d(wpk, 64, 576, 64);
  static bool configured = false;
  if (!configured) {
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_FWD_STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_BN1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_BN2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_BN2_DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_FWD_STATS | SP_F2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_FWD_STATS | SP_F2 | SP_FIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_FWD_LRELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_FWD_ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_BN1 | SP_F2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_BN2 | SP_F2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64s_fprop_kernel<SPEC_BN2_DUAL | SP_F2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    configured = true;
  }
  if (stacked) {
    PCG_REQUIRE(p.in_stages >= 4, "stacked halo-tile kernel: four class regions must fit the ring");
    CUtensorMap tmX = make_tmap_nhwc_rowclass(in, N, H, W, 64, p.WP, p.R + 1);
    CUtensorMap tmOut = make_tmap_nhwc_rowclass(out, N, H, W, 64, W, p.R);
    CUtensorMap tmExtra = make_tmap_nhwc_rowclass(extra != nullptr ? extra : out, N, H, W, 64, W, p.R);
    CUtensorMap tmAdd = make_tmap_nhwc_rowclass(add_tma != nullptr ? add_tma : out, N, H, W, 64, W, p.R);
    auto kernel = conv_tc64s_fprop_kernel<-1>;
    switch (epilogue_spec(p)) {
      case SPEC_FWD_STATS: kernel = conv_tc64s_fprop_kernel<SPEC_FWD_STATS>; break;
      case SPEC_BN1: kernel = conv_tc64s_fprop_kernel<SPEC_BN1>; break;
      case SPEC_BN2: kernel = conv_tc64s_fprop_kernel<SPEC_BN2>; break;
      case SPEC_BN2_DUAL: kernel = conv_tc64s_fprop_kernel<SPEC_BN2_DUAL>; break;
      case SPEC_FWD_STATS | SP_F2: kernel = conv_tc64s_fprop_kernel<SPEC_FWD_STATS | SP_F2>; break;
      case SPEC_FWD_STATS | SP_F2 | SP_FIN: kernel = conv_tc64s_fprop_kernel<SPEC_FWD_STATS | SP_F2 | SP_FIN>; break;
      case SPEC_FWD_LRELU: kernel = conv_tc64s_fprop_kernel<SPEC_FWD_LRELU>; break;
      case SPEC_FWD_ADD: kernel = conv_tc64s_fprop_kernel<SPEC_FWD_ADD>; break;
      case SPEC_BN1 | SP_F2: kernel = conv_tc64s_fprop_kernel<SPEC_BN1 | SP_F2>; break;
      case SPEC_BN2 | SP_F2: kernel = conv_tc64s_fprop_kernel<SPEC_BN2 | SP_F2>; break;
      case SPEC_BN2_DUAL | SP_F2: kernel = conv_tc64s_fprop_kernel<SPEC_BN2_DUAL | SP_F2>; break;
      default: break;
    }
    launch_k_pdl(kernel, dim3(conv_tc64_fprop_grid(N, H, W)), dim3(F_THREADS), total(), stream, tmX, tmW, tmOut, tmExtra,
             tmAdd, p);
    PCG_COUNT_LAUNCH();
    PCG_LAUNCH_CHECK();
    return;
  }
  const StatsFinalize fin = p.fin;                 // the one-class kernel leaves the finalisation to a second launch
  p.fin.mode = 0;
  CUtensorMap tmX = make_tmap_nhwc_box(in, N, H, W, 64, p.WP, p.R + 2);
  CUtensorMap tmOut = make_tmap_nhwc_box(out, N, H, W, 64, W, p.R);
  CUtensorMap tmExtra = make_tmap_nhwc_box(extra != nullptr ? extra : out, N, H, W, 64, W, p.R);
  launch_k_pdl(conv_tc64_fprop_kernel, dim3(conv_tc64_grid(N, H, W)), dim3(F_THREADS), total(), stream, tmX, tmW, tmOut, tmExtra, p);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
  if (fin.mode == 1)
    bn_finalize(p.stats, conv_tc64_grid(N, H, W), fin.M, 64, fin.gamma, fin.beta, fin.eps, fin.momentum, fin.running_mean,
                fin.running_var, fin.nbt, fin.mean, fin.rstd, fin.scale, fin.shift, stream);
  else if (fin.mode == 2)
    bn_bwd_finalize(p.stats, conv_tc64_grid(N, H, W), fin.M, 64, fin.dgamma, fin.dbeta, fin.c12, stream);
}

// ------------------------------------------------------------------------------------------
// wgrad: dW[tap][ci][co] = sum_p x[p@tap][ci] * dy[p][co]; K = the 128 padded positions of a tile
// ------------------------------------------------------------------------------------------
constexpr int G_STAGES = 4;
constexpr int DY_PAD_BYTES = 1024;               // eight zero rows in front of the dY tile (read by its shifted views)
constexpr int DY_STAGE_BYTES = DY_PAD_BYTES + 16384;   // + 128 rows x 128 B (rows >= R*(W+2) stay zero)
constexpr int G_THREADS = 192;
constexpr int G_SMEM_BYTES = 1024 + G_STAGES * (IN_STAGE_BYTES + DY_STAGE_BYTES) + 256;

__global__ void __launch_bounds__(G_THREADS, 1)
conv_tc64_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, int N, int H,
                       int W, int WP, int R, int tiles_per_img, int total_tiles, int variant,
                       float* __restrict__ part) {
  pdl_launch_dependents();                        // the prologue below touches no global memory: see pdl_wait() further down
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sx = smem;
  uint8_t* sdy = smem + G_STAGES * IN_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdy + G_STAGES * DY_STAGE_BYTES);
  uint64_t* full = bars;               // [G_STAGES]
  uint64_t* empty = bars + G_STAGES;   // [G_STAGES]
  uint64_t* done = bars + 2 * G_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  // zero everything once: x rows past the box are read by padding positions whose dy rows are zero (must be
  // finite), dy rows [R*WP, 128) must be exactly zero
  for (int i = threadIdx.x; i < G_STAGES * (IN_STAGE_BYTES + DY_STAGE_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                     // from here on the predecessor kernel has completed and flushed
  const int xbytes = (R + 2) * WP * 128, dybytes = R * WP * 128;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol = l2_evict_first_policy();
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img, h0 = (tile % tiles_per_img) * R;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], xbytes + dybytes);
        if (variant & 4096) {                      // both operands are read for the last time (PCG_L2_HINTS bit 1)
          tma_load_4d_hint(&tmX, &full[stage], sx + stage * IN_STAGE_BYTES, 0, -1, h0 - 1, n, pol);
          tma_load_4d_hint(&tmDY, &full[stage], sdy + stage * DY_STAGE_BYTES + DY_PAD_BYTES, 0, 0, h0, n, pol);
        } else {
          tma_load_4d(&tmX, &full[stage], sx + stage * IN_STAGE_BYTES, 0, -1, h0 - 1, n);
          tma_load_4d(&tmDY, &full[stage], sdy + stage * DY_STAGE_BYTES + DY_PAD_BYTES, 0, 0, h0, n);
        }
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);     // A, B both MN-major
      constexpr uint32_t idesc192 = umma_idesc_bf16(128, 192, 1, 1), idesc128 = umma_idesc_bf16(128, 128, 1, 1);
      // Default: BOTH operands carry shifted views.  dW[r][s] = sum_pos X[pos + r*WP + s] dY[pos] = sum_p X[p + r*WP] dY[p - s]
      // (p = pos + s; dY is zero outside its tile), so the column shift moves to the dY side: B = the dY tile started
      // 2, 1, 0 rows early, stacked along N (three 64-channel blocks 128 B apart, N = 192), A = the X views of tap rows
      // r = 0, 1 stacked along M.  One M128 x N192 MMA per k-step covers six taps; for r = 2 the column shift is split
      // between the operands, s = a + b: A = the r = 2 view shifted by a = 0, 1 pixels (M = 128), B = dY started b = 2, 0
      // rows early (N = 128), which yields s = 2, 3 (unused), 0, 1.  16 MMAs per tile instead of 40.  Accumulator 0 (TMEM
      // columns 0-191): lanes (r, ci), columns (2 - s, co); accumulator 1 (columns 192-319): lanes (a, ci), columns (b, co).
      // variant bit 1024: the original scheme (five M128 x N64 accumulators, two taps each).
      const bool stacked = (variant & 1024) == 0;
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t x_base = smem_u32(sx + stage * IN_STAGE_BYTES);
        const uint32_t dy_base = smem_u32(sdy + stage * DY_STAGE_BYTES + DY_PAD_BYTES);
        if (elect_one()) {
          const uint32_t acc0 = first ? 0u : 1u;
          if (stacked) {
            const uint64_t a1 = a_desc(x_base, (uint32_t)WP * 128u, 0);
            const uint64_t a2 = a_desc(x_base + 2u * (uint32_t)WP * 128u, 128u, 0);
            const uint64_t b3 = umma_smem_desc(dy_base - 256u, 128, 1024);
            const uint64_t b2 = umma_smem_desc(dy_base - 256u, 256, 1024);      // blocks b = 2, 0
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_f16(tmem_base, a1 + 128 * k, b3 + 128 * k, idesc192, k != 0 ? 1u : acc0);
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_f16(tmem_base + 192, a2 + 128 * k, b2 + 128 * k, idesc128, k != 0 ? 1u : acc0);
          } else {
            const uint64_t b0 = umma_smem_desc(dy_base, 16, 1024);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              const int t0 = 2 * j, t1 = (j < 4) ? t0 + 1 : t0;
              const uint32_t o0 = (uint32_t)((t0 / 3) * WP + (t0 % 3)) * 128u;
              const uint32_t o1 = (uint32_t)((t1 / 3) * WP + (t1 % 3)) * 128u;
              const uint32_t lbo = (j < 4) ? (o1 - o0) : 128u;            // distance between the two 64-channel M blocks
              const uint64_t aj = a_desc(x_base + o0, lbo, variant & 1);
#pragma unroll
              for (int k = 0; k < 8; ++k) umma_f16(tmem_base + j * 64, aj + 128 * k, b0 + 128 * k, idesc, k != 0 ? 1u : acc0);
            }
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        first = false;
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
    float* my = part + (size_t)blockIdx.x * 9 * 64 * 64;
    if ((variant & (1024 | 2048)) == 0) {
      // Default write-out.  A thread owns one accumulator row (tap, ci) = 256 contiguous bytes of the partial, so direct
      // stores touch 32 different lines per warp instruction (~9,200 LSU transactions per CTA, ~4.7 us at the end of
      // every launch).  Instead each warp stages its 32 rows of a tap (8 KB, contiguous in the partial) in the idle
      // operand ring - 16-byte chunk c of row ci at chunk c ^ (ci & 15), which keeps the staging stores free of bank
      // conflicts - and one bulk copy per warp and tap sends them off.  The partial therefore carries that chunk
      // swizzle inside every 256-byte row; wgrad_reduce_tc(..., swizzled = true) undoes it when it writes dW.
      uint8_t* stage = smem + (size_t)(warp - 2) * 5 * 8192;           // every MMA has completed: the ring is idle
      const int ci = row & 63, hi = row >> 6;
      int it = 0;
#pragma unroll 1
      for (int a = 0; a < 2; ++a) {
#pragma unroll 1
        for (int jb = 0; jb < (a == 0 ? 3 : 2); ++jb, ++it) {
          const int sc = a == 0 ? 2 - jb : hi + (jb == 0 ? 2 : 0);       // as below
          const bool live = sc < 3;                                      // warp-uniform (hi is)
          const int tap = (a == 0 ? hi : 2) * 3 + sc;
          uint8_t* st = stage + it * 8192;
#pragma unroll
          for (int chunk = 0; chunk < 2; ++chunk) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + a * 192 + jb * 64 + chunk * 32, v);
            tmem_ld_wait();
            if (live) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<uint4*>(st + lane * 256 + (((chunk * 8 + j4) ^ (lane & 15)) << 4)) =
                    make_uint4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            }
          }
          if (live) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) bulk_store(my + ((size_t)tap * 64 + (ci & 32)) * 64, st, 8192);
          }
        }
      }
      if (lane == 0) {
        tma_store_commit();
        tma_store_wait<0>();
      }
      __syncwarp();
    } else if ((variant & 1024) == 0) {
#pragma unroll 1
      for (int a = 0; a < 2; ++a) {
        const int ci = row & 63, hi = row >> 6;
#pragma unroll 1
        for (int jb = 0; jb < (a == 0 ? 3 : 2); ++jb) {
          // accumulator 0: tap (r = hi, s = 2 - jb); accumulator 1: tap (2, s = hi + (jb == 0 ? 2 : 0)), s = 3 unused
          const int sc = a == 0 ? 2 - jb : hi + (jb == 0 ? 2 : 0);
          const bool live = sc < 3;
          const int tap = (a == 0 ? hi : 2) * 3 + sc;
#pragma unroll
          for (int chunk = 0; chunk < 2; ++chunk) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + a * 192 + jb * 64 + chunk * 32, v);
            tmem_ld_wait();
            if (live) {
              float4* dst = reinterpret_cast<float4*>(my + ((size_t)tap * 64 + ci) * 64 + chunk * 32);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                dst[j4] = make_float4(__uint_as_float(v[j4 * 4]), __uint_as_float(v[j4 * 4 + 1]),
                                      __uint_as_float(v[j4 * 4 + 2]), __uint_as_float(v[j4 * 4 + 3]));
            }
          }
        }
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < 5; ++j) {
        const int tap = 2 * j + (row >> 6), ci = row & 63;
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// layout of the partials the kernel writes under the current variant (see the write-out above)
bool conv_tc64_wgrad_swizzled() { return (g_variant & (1024 | 2048)) == 0; }

void conv_tc64_wgrad(const bf16* x, const bf16* dy, int N, int H, int W, float* part, cudaStream_t stream) {
  PCG_PROFILE("conv_tc64_wgrad", stream);
  PCG_REQUIRE(conv_tc64_supported(H, W), "image too wide for the halo-tile kernel");
  const int WP = W + 2, R = 128 / WP;
  const int tiles_per_img = (H + R - 1) / R;
  CUtensorMap tmX = make_tmap_nhwc_box(x, N, H, W, 64, WP, R + 2);
  CUtensorMap tmDY = make_tmap_nhwc_box(dy, N, H, W, 64, WP, R);
  static bool configured = false;
  if (!configured) {
    PCG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc64_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES));
    configured = true;
  }
  launch_k_pdl(conv_tc64_wgrad_kernel, dim3(conv_tc64_grid(N, H, W)), dim3(G_THREADS), G_SMEM_BYTES, stream, 
      tmX, tmDY, N, H, W, WP, R, tiles_per_img, N * tiles_per_img, g_variant | ((g_l2_hints & 1) ? 4096 : 0), part);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

// fp32-storage front end of the tensor-core convolution kernels.
//
// The Python-composed step plans (DCGAN, pcg_b200/dcgan/plan.py) keep NHWC fp32 activations and call the primitive
// operators pcg_conv_fprop / pcg_conv_dgrad / pcg_conv_wgrad.  With the tensor-core mode switched on
// (pcg_set_conv_tensor_cores(1)) eligible layers - Conv2d / ConvTranspose2d with 64-multiple channel counts,
// dconv_gan/mnist/mnist_dcgan.py:72-116 - are routed here: operands are rounded to bf16 into library-owned scratch,
// the tcgen05 implicit-GEMM kernels of conv_tc.cu run with fp32 accumulation, results return as fp32.
// Scratch grows on demand with cudaMalloc, which is illegal during stream capture: plans run one eager pass first.
#include "conv_auto.cuh"

#include <stdlib.h>

#include <vector>

#include "conv_tc.cuh"
#include "elementwise.cuh"

namespace pcg {

static bool g_tc_mode = false;
static int g_tc_ops = -1;        // bring-up: PCG_TC_OPS bit mask (1 fprop, 2 dgrad, 4 wgrad) narrows the mode
static bool op_on(int bit) {
  if (g_tc_ops < 0) {
    const char* e = getenv("PCG_TC_OPS");
    g_tc_ops = e ? atoi(e) : 7;
  }
  return g_tc_mode && (g_tc_ops & bit);
}
void conv_auto_set_tensor_cores(bool on) { g_tc_mode = on; }
bool conv_auto_tensor_cores() { return g_tc_mode; }

enum { SL_X = 0, SL_Y = 1, SL_W = 2, SL_E = 3, SL_E2 = 4, SL_PART = 5, SL_COUNT = 6 };
struct Slot {
  void* p = nullptr;
  size_t cap = 0;
};
static Slot g_slots[SL_COUNT];
// Blocks replaced by a larger one are retired, not freed: a CUDA graph captured by an earlier plan of this process
// (another batch size: the tail batch of an epoch, a second model) still holds their addresses.  Growth is geometric,
// so the retired blocks of a slot add up to less than its final size.
static std::vector<void*> g_retired;

static void* scratch(int slot, size_t bytes, cudaStream_t stream) {
  Slot& s = g_slots[slot];
  if (bytes <= s.cap) return s.p;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &st);
  if (st != cudaStreamCaptureStatusNone)
    throw Error(5, "tensor-core scratch must be sized by an eager pass before stream capture");
  if (s.p) g_retired.push_back(s.p);
  const size_t cap = bytes > 2 * s.cap ? bytes + bytes / 4 + 4096 : 2 * s.cap;
  PCG_CHECK_CUDA(cudaMalloc(&s.p, cap));
  s.cap = cap;
  return s.p;
}

// ---- operand cache ----------------------------------------------------------------------------------------------
// Every tensor-core call converts its fp32 operands to bf16 first; in a training step each activation is an operand twice
// (the next layer's forward and weight gradient), each gradient twice (data and weight gradient), each weight once per
// pass.  With the cache on (pcg_set_operand_cache(1), the step plans do), a conversion gets its own buffer keyed by
// (source pointer, size, layout) and is reused while the source is unchanged.  "Unchanged" is the CALLER's knowledge:
// pcg_operand_cache_invalidate(ptr, bytes) after every write to a tensor (pcg_b200/ops.py does it for every operator it
// wraps, from the operator's declared write set), pcg_operand_cache_clear() at the start of a step body (inputs were
// copied in by torch, outside the library).  The decisions are made when a launch is issued, so a captured graph
// replays exactly the conversions the capturing pass made.
struct CacheEntry {
  const float* src;
  long long n;
  int layout;          // T | kind << 4 | lo_last << 8 | C << 12 (C only where the layout depends on it)
  bf16* buf;
  size_t cap;
  bool valid;
};
static std::vector<CacheEntry> g_cache;
static bool g_cache_on = false;
void conv_auto_set_operand_cache(bool on) { g_cache_on = on; }
bool conv_auto_operand_cache() { return g_cache_on; }
void conv_auto_cache_clear() {
  for (auto& e : g_cache) e.valid = false;
}
void conv_auto_cache_invalidate(const void* p, size_t bytes) {
  const char* lo = static_cast<const char*>(p);
  const char* hi = lo + bytes;
  for (auto& e : g_cache) {
    const char* a = reinterpret_cast<const char*>(e.src);
    if (e.valid && a < hi && lo < a + e.n * sizeof(float)) e.valid = false;
  }
}
enum { KIND_SPLIT = 0, KIND_DGRAD_PACK = 2 };
// Returns the entry for this key; *fresh = true when the caller must (re)fill entry->buf.
static CacheEntry* cache_get(const float* src, long long n, int layout, size_t bytes, cudaStream_t stream, bool* fresh) {
  for (auto& e : g_cache)
    if (e.src == src && e.n == n && e.layout == layout && e.cap >= bytes) {
      *fresh = !e.valid;
      e.valid = true;
      return &e;
    }
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &st);
  if (st != cudaStreamCaptureStatusNone)
    throw Error(5, "tensor-core operand cache must be populated by an eager pass before stream capture");
  CacheEntry e{src, n, layout, nullptr, bytes, true};
  PCG_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&e.buf), bytes));
  g_cache.push_back(e);
  *fresh = true;
  return &g_cache.back();
}
// bf16 side output for a producer kernel: the plain-bf16 (T = 1) conversion buffer some consumer has already asked for
// this tensor, marked valid - or NULL (cache off, bf16x3 mode, or nobody converts this tensor)
bf16* conv_auto_cache_producer(const float* dst, long long n) {
  if (!g_cache_on) return nullptr;
  for (auto& e : g_cache)
    if (e.src == dst && e.n == n && e.layout == 1) {
      e.valid = true;
      return e.buf;
    }
  return nullptr;
}

static bf16* to_bf16(int slot, const float* src, long long n, cudaStream_t s) {
  bf16* dst = reinterpret_cast<bf16*>(scratch(slot, (size_t)n * sizeof(bf16), s));
  convert_from_f32<bf16>(src, n, dst, s);
  return dst;
}

// bf16x3: an fp32 product a*w is emulated as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (hi = bf16(v), lo = bf16(v - hi); the
// dropped a_lo*w_lo term is ~2^-16 relative), accumulated in fp32 by the tensor cores.  No kernel change is needed:
// the contraction dimension is tripled.  For fprop / dgrad the data operand becomes [M][hi | lo | hi] (3C channels)
// and the weights [.. | W_hi | W_hi | W_lo]; for wgrad (contraction over pixels) the batch is tripled: images
// (hi, lo, hi) of X against (hi, hi, lo) of dY.
// Why not plain bf16 operands: on the DCGAN step the rounding noise is amplified by the mean subtractions of three
// stacked train-mode BatchNorm backwards (measured 6-15 % relative L2 on the gradients against the fp32 oracle, whose
// own fp32 CUDA-core realisation only agrees to ~1e-2 there); bf16x3 brings the tensor-core path to the fp32 level.
// PCG_TC_TERMS=1 selects plain bf16 operands (1/3 of the tensor work).
static int g_terms = -1;
void conv_auto_set_terms(int t) { g_terms = (t == 1) ? 1 : 3; }
int conv_auto_terms();
static int terms() {
  if (g_terms < 0) {
    const char* e = getenv("PCG_TC_TERMS");
    g_terms = (e && atoi(e) == 1) ? 1 : 3;
  }
  return g_terms;
}

// src fp32 [M][C] -> dst bf16 [M][T*C]; T = 3: (hi | lo | hi) or, with lo_last, (hi | hi | lo); T = 1: (hi)
__global__ void split_bf16_kernel(const float* __restrict__ src, long long M, int C, int T, int lo_last,
                                  bf16* __restrict__ dst) {
  pdl_enter();
  const long long n4 = M * C / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4, row = e / C;
    const int c = (int)(e - row * C);
    const float4 v = *reinterpret_cast<const float4*>(src + e);
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hi[j] = __float2bfloat16_rn(f[j]);
      lo[j] = __float2bfloat16_rn(f[j] - __bfloat162float(hi[j]));
    }
    bf16* d = dst + row * T * C + c;
    *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(hi);
    if (T == 3) {
      *reinterpret_cast<uint2*>(d + C) = *reinterpret_cast<const uint2*>(lo_last ? hi : lo);
      *reinterpret_cast<uint2*>(d + 2 * C) = *reinterpret_cast<const uint2*>(lo_last ? lo : hi);
    }
  }
}
int conv_auto_terms() { return terms(); }

static bf16* split(int slot, const float* src, long long M, int C, int T, bool lo_last, cudaStream_t s) {
  PCG_REQUIRE(C % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0, "split operand: C % 4 and 16-byte alignment");
  bf16* dst;
  if (g_cache_on) {
    bool fresh;
    const int layout = T == 1 ? 1 : (T | (lo_last ? 1 << 8 : 0) | C << 12);      // T = 1: a plain copy, whatever the shape
    dst = cache_get(src, M * C, layout, (size_t)M * T * C * sizeof(bf16), s, &fresh)->buf;
    if (!fresh) return dst;
  } else {
    dst = reinterpret_cast<bf16*>(scratch(slot, (size_t)M * T * C * sizeof(bf16), s));
  }
  PCG_PROFILE("convert", s);
  long long b = (M * C / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  launch_k(split_bf16_kernel, dim3((int)(b < cap ? (b > 0 ? b : 1) : cap)), dim3(256), 0, s, src, M, C, T, lo_last ? 1 : 0, dst);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
  return dst;
}
// batch-tripled variant for wgrad: src fp32 [n] -> dst bf16 [T][n]; blocks (hi, lo, hi) or (hi, hi, lo)
__global__ void split_batch_kernel(const float* __restrict__ src, long long n, int T, int lo_last, bf16* __restrict__ dst) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float f = src[i];
    const bf16 hi = __float2bfloat16_rn(f);
    dst[i] = hi;
    if (T == 3) {
      const bf16 lo = __float2bfloat16_rn(f - __bfloat162float(hi));
      dst[n + i] = lo_last ? hi : lo;
      dst[2 * n + i] = lo_last ? lo : hi;
    }
  }
}
static bf16* split_batch(int slot, const float* src, long long n, int T, bool lo_last, cudaStream_t s) {
  bf16* dst;
  if (g_cache_on) {
    bool fresh;
    const int layout = T == 1 ? 1 : (T | 1 << 4 | (lo_last ? 1 << 8 : 0));
    dst = cache_get(src, n, layout, (size_t)n * T * sizeof(bf16), s, &fresh)->buf;
    if (!fresh) return dst;
  } else {
    dst = reinterpret_cast<bf16*>(scratch(slot, (size_t)n * T * sizeof(bf16), s));
  }
  PCG_PROFILE("convert", s);
  long long b = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  launch_k(split_batch_kernel, dim3((int)(b < cap ? (b > 0 ? b : 1) : cap)), dim3(256), 0, s, src, n, T, lo_last ? 1 : 0, dst);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
  return dst;
}

static bool ch64(int a, int b) { return a % 64 == 0 && b % 64 == 0; }

bool conv_fprop_auto(const float* in, const ConvGeom& g, const float* wf, const GenEpilogue<float>& e, float* out,
                     cudaStream_t s) {
  if (!op_on(1) || !ch64(g.Cin, g.Cout) || e.act_ref != nullptr) return false;
  const int T = terms();
  const long long nout = g.Mout() * g.Cout;
  const bf16* xb = split(SL_X, in, g.Min(), g.Cin, T, false, s);                                       // [Min][hi | lo | hi]
  const bf16* wb = split(SL_W, wf, (long long)g.Cout * g.ksize * g.ksize, g.Cin, T, true, s);           // [Cout][tap][hi | hi | lo]
  ConvEpilogue c;
  c.bias = e.bias; c.act = e.act; c.slope = e.slope;
  if (e.add_src != nullptr) c.add_src = to_bf16(SL_E, e.add_src, nout, s);
  c.out_f32 = out;                                   // fp32 straight from the accumulators
  conv_tc_fprop(xb, g.N, g.H, g.W, T * g.Cin, wb, g.Cout, g.ksize, g.stride, g.pad, c, nullptr, s);
  return true;
}

bool conv_dgrad_auto(const float* dout, const ConvGeom& g, const float* wd, const GenEpilogue<float>& e, float* din,
                     cudaStream_t s) {
  if (!op_on(2) || !ch64(g.Cin, g.Cout) || g.stride != 2 || g.pad != 1 || g.ksize != 4 || (g.H & 1) || (g.W & 1) ||
      e.bias != nullptr)
    return false;
  const int T = terms();
  const long long nin = g.Min() * g.Cin;
  const bf16* dyb = split(SL_X, dout, g.Mout(), g.Cout, T, false, s);                                  // [Mout][hi | lo | hi]
  bf16* packed;
  bool fresh = true;
  if (g_cache_on)
    packed = cache_get(wd, (long long)16 * g.Cout * g.Cin, T | KIND_DGRAD_PACK << 4, (size_t)16 * T * g.Cout * g.Cin * sizeof(bf16),
                       s, &fresh)->buf;
  else
    packed = reinterpret_cast<bf16*>(scratch(SL_W, (size_t)16 * T * g.Cout * g.Cin * sizeof(bf16), s));
  if (fresh) pack_dgrad_s2_k4_tc(wd, g.Cout, g.Cin, packed, s, T);
  ConvEpilogue c;
  c.act = e.act; c.slope = e.slope;
  if (e.add_src != nullptr) c.add_src = to_bf16(SL_E, e.add_src, nin, s);
  if (e.act_ref != nullptr && e.ref_act != ACT_NONE) {
    c.act_ref = to_bf16(SL_E2, e.act_ref, nin, s);      // only its sign is used
    c.ref_act = e.ref_act; c.ref_slope = e.ref_slope;
  }
  c.out_f32 = din;
  conv_tc_dgrad_s2(dyb, g.N, g.H, g.W, g.Cin, T * g.Cout, packed, c, nullptr, s, nullptr, 4);
  return true;
}

bool conv_wgrad_auto(const float* in, const float* dout, const ConvGeom& g, float* dw, cudaStream_t s) {
  if (!op_on(4) || !ch64(g.Cin, g.Cout) || (g.stride != 1 && g.stride != 2) || (g.ksize != 3 && g.ksize != 4))
    return false;
  const int T = terms();
  const bf16* xb = split_batch(SL_X, in, g.Min() * g.Cin, T, false, s);          // images (hi, lo, hi)
  const bf16* dyb = split_batch(SL_Y, dout, g.Mout() * g.Cout, T, true, s);      // images (hi, hi, lo)
  const int N3 = T * g.N;
  const int splits = conv_tc_wgrad_general_splits(N3, g.H, g.W, g.Cin, g.Cout, g.stride, g.ksize, g.pad);
  float* part = reinterpret_cast<float*>(scratch(SL_PART, (size_t)splits * g.Cout * g.K() * sizeof(float), s));
  conv_tc_wgrad_general(xb, dyb, N3, g.H, g.W, g.Cin, g.Cout, g.stride, part, s, g.ksize, g.pad);
  wgrad_reduce_generic(part, splits, g.Cout, g.Cin, g.ksize * g.ksize, dw, s);
  return true;
}

}  // namespace pcg

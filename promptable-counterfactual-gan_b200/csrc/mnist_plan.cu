// MNIST conv CounteRGAN step plan: owns the workspaces and orchestrates the kernels of one
// iteration of conditional_counteRGAN/mnist/trainer.py:89-132 (see include/pcg.h).
//
// Data layout in HBM
//   activations   NHWC, element type T (fp32 or bf16 per pcg_mnist_config.precision)
//   images        [B][784] fp32 (x, mask, raw, masked, x_cf, per-pixel gradients)
//   parameters    caller-owned flat fp32 arenas in torch layout (pcg_mnist_layout); the plan keeps
//                 packed copies ([Cout][tap][Cin] fp32 / bf16, rotated for dgrad) refreshed after
//                 every Adam update
//   saved for backward: h0, per block (y1, z1, y2, h_next), hm; BN mean/rstd/scale/shift
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/pcg.h"
#include "common.cuh"
#include "conv_generic.cuh"
#include "conv_small.cuh"
#include "conv_tc.cuh"
#include "elementwise.cuh"

namespace pcg {

struct TensorSpec {
  long long offset, numel;
};

static std::vector<TensorSpec> net_layout(int net, int ch, int nres, long long* total) {
  std::vector<long long> sizes;
  if (net == 0) {
    sizes = {10LL * 784, (long long)ch * 3 * 9, ch};
    for (int i = 0; i < nres; ++i)
      for (int j = 0; j < 2; ++j) {
        sizes.push_back((long long)ch * ch * 9);
        sizes.push_back(ch);
        sizes.push_back(ch);
        sizes.push_back(ch);
      }
    sizes.push_back((long long)ch * ch * 9);
    sizes.push_back(ch);
    sizes.push_back((long long)ch * 9);
    sizes.push_back(1);
  } else if (net == 1) {
    sizes = {10LL * 784, 64LL * 2 * 9, 128LL * 64 * 9, 256LL * 128 * 9, 256LL * 256 * 9, 256, 1};
  } else if (net == 2) {
    sizes = {32LL * 9, 32, 64LL * 32 * 9, 64, 128LL * 64 * 9, 128, 256LL * 6272, 256, 2560, 10};
  } else {
    throw Error(1, "net must be 0 (G), 1 (D) or 2 (C)");
  }
  std::vector<TensorSpec> out;
  long long off = 0;
  for (long long n : sizes) {
    out.push_back({off, n});
    off += (n + 3) / 4 * 4;
  }
  if (total) *total = off;
  return out;
}

// ---------------------------------------------------------------------------------------------
// Weight re-packing after an Adam update: ONE launch per network walks a device-resident table of layers and
// writes every layout the step's kernels read (fp32 [Cout][tap][Cin] / [Cin][tap][Cout] for the CUDA-core kernels,
// bf16 [Cout][tap][Cin], 180-degree-rotated [Cin][tap][Cout] and the stride-2 parity-class matrices for tcgen05).
struct PackDesc {
  const float* w;
  float *wf, *wd;
  bf16 *tcf, *tcd, *tcs2;
  int Cout, Cin, taps, perm_hw;   // dimensions of the torch tensor
  int ld_cout, ld_cin;            // channel counts of the packed layouts (>= Cout / Cin: zero-padded layers)
  int begin;                 // first flat element of this layer in the table-wide index space
};
constexpr int PACK_MAX_LAYERS = 40;
__global__ void __launch_bounds__(256) pack_multi_kernel(const PackDesc* __restrict__ table, int nlayers, int total) {
  pdl_enter();
  __shared__ PackDesc d[PACK_MAX_LAYERS];
  for (int i = threadIdx.x; i < nlayers * (int)(sizeof(PackDesc) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(d)[i] = reinterpret_cast<const uint32_t*>(table)[i];
  __syncthreads();
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
    int l = 0;
    while (l + 1 < nlayers && g >= d[l + 1].begin) ++l;
    const PackDesc& L = d[l];
    const int i = g - L.begin;
    const int taps = L.taps, Cin = L.Cin, Cout = L.Cout;
    const int tap = i % taps, ci0 = (i / taps) % Cin, co = i / (taps * Cin);
    int ci = ci0;
    if (L.perm_hw > 0) {         // torch flatten index c*HW + hw  ->  NHWC flatten index hw*C + c
      const int C = Cin / L.perm_hw;
      const int c = ci / L.perm_hw, hw = ci - c * L.perm_hw;
      ci = hw * C + c;
    }
    const float v = L.w[i];
    const bf16 vb = __float2bfloat16_rn(v);
    const int ldi = L.ld_cin, ldo = L.ld_cout;
    const size_t f = ((size_t)co * taps + tap) * ldi + ci;
    if (L.wf) L.wf[f] = v;
    if (L.wd) L.wd[((size_t)ci * taps + tap) * ldo + co] = v;
    if (L.tcf) L.tcf[f] = vb;
    if (L.tcd) L.tcd[((size_t)ci * taps + (taps - 1 - tap)) * ldo + co] = vb;
    if (L.tcs2) {                // parity classes of the stride-2 data gradient (conv_tc.cu: pack_dgrad_s2_kernel)
      const int s2 = tap % 3, r = tap / 3;
      const int ph = (r == 1) ? 0 : 1, pw = (s2 == 1) ? 0 : 1;
      const int oh = (r == 0) ? 1 : 0, ow = (s2 == 0) ? 1 : 0;
      const int taps_w = pw ? 2 : 1, ntap = (ph ? 2 : 1) * taps_w;
      const size_t u = (size_t)ldo * ldi;
      bf16* dst = L.tcs2 + (ph ? (pw ? 5 * u : 3 * u) : (pw ? u : 0));
      dst[((size_t)ci * ntap + oh * taps_w + ow) * ldo + co] = vb;
    }
  }
}
struct PackTable {
  PackDesc* dev = nullptr;
  int nlayers = 0, total = 0;
  void launch(cudaStream_t s) const {
    if (nlayers == 0) return;
    PCG_PROFILE("pack_weights", s);
    int blocks = cdiv(total, 256 * 4);
    if (blocks > 1184) blocks = 1184;
    launch_k(pack_multi_kernel, dim3(blocks), dim3(256), 0, s, dev, nlayers, total);
    PCG_COUNT_LAUNCH();
    PCG_LAUNCH_CHECK();
  }
};

// Eval-mode BatchNorm folded into the convolution in front of it (inference forward):
//   BN(conv(x) + b) * extra = conv'(x) + b'  with  w'[co] = w[co] * s[co] * extra,  b'[co] = extra * ((b[co] - rm[co]) * s[co] + beta[co]),
//   s = gamma / sqrt(running_var + eps).  One launch folds every (conv, BN) pair of the generator's residual blocks
//   into bf16 [Cout][tap][Cin] weights + fp32 biases, so the eval forward is conv -> conv with fused epilogues only.
struct FoldDesc {
  const float *w, *b, *gamma, *beta, *rm, *rv;
  bf16* wq;
  float* bq;
  float extra;
  int Cout, Cin, taps;
};
__global__ void __launch_bounds__(256) fold_bn_kernel(const FoldDesc* __restrict__ table, float eps) {
  pdl_enter();
  const FoldDesc L = table[blockIdx.y];
  const int total = L.Cout * L.Cin * L.taps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % L.taps, ci = (i / L.taps) % L.Cin, co = i / (L.taps * L.Cin);
    const float sc = L.gamma[co] / sqrtf(L.rv[co] + eps) * L.extra;
    L.wq[((size_t)co * L.taps + tap) * L.Cin + ci] = __float2bfloat16_rn(L.w[i] * sc);
    if (tap == 0 && ci == 0)
      L.bq[co] = L.extra * ((L.b[co] - L.rm[co]) * (L.gamma[co] / sqrtf(L.rv[co] + eps)) + L.beta[co]);
  }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
struct ConvLayer {
  ConvGeom g{};
  const float* w = nullptr;   // torch OIHW in the parameter arena
  const float* b = nullptr;
  float* dw = nullptr;        // gradient arena slots (nullptr for the frozen classifier)
  float* db = nullptr;
  float* wf = nullptr;        // fp32 [Cout][taps][Cin]
  float* wd = nullptr;        // fp32 [Cin][taps][Cout]
  bf16* tcf = nullptr;        // bf16 [Cout][taps][Cin]
  bf16* tcd = nullptr;        // bf16 [Cin][taps][Cout], taps rotated
  bf16* tcs2 = nullptr;       // bf16 parity-class matrices of the stride-2 data gradient
  bool tc_fprop = false, tc_dgrad = false, tc_wgrad = false, tc_dgrad_s2 = false;
  bool tc64 = false;          // 64->64 3x3 s1 p1: halo-tile kernels (conv_tc64.cu)
  bool tc_wgrad_gen = false;  // general tensor-core wgrad (conv_tc.cu)
  bf16* tcf_eval = nullptr;   // eval-mode BatchNorm folded in (fold_bn_kernel)
  float* b_eval = nullptr;
  bool to1_fprop = false;     // Cout == 1: conv_to1 on tcf (conv_small.cu)
  bool to1_dgrad = false;     // Cin <= 4: one input channel's data gradient = conv_to1 on a row of tcd
  bool few_fprop = false;     // Cin <= 3: conv_few on tcf
  bool few_dgrad = false;     // Cout <= 3: data gradient = conv_few on tcd
  bool s2_to1_dgrad = false;  // Cin <= 4, stride 2: one input channel's data gradient = dgrad_s2_to1 on a row of tcd
  bool few_wgrad = false;     // Cin <= 3, Cout == 64: wgrad_few (also produces the bias gradient)
  bool to1_wgrad = false;     // 64 -> 1: wgrad_to1 (also produces the bias gradient)
  int perm_hw = 0;
  int src_cin = 0, src_cout = 0;   // channel counts of the torch weight (g.Cin / g.Cout may be zero-padded)
  std::string name, tag_f, tag_d, tag_w;

  // Only the layouts a kernel of this plan actually reads are written: the fprop tensor-core packing is the bf16
  // image of wf (same [Cout][taps][Cin] order, incl. the NHWC-flatten permutation of fc.1); the dgrad packing
  // additionally rotates the taps.  `wd_generic_too`: a channel-select data gradient runs on the CUDA cores.
  PackDesc desc(bool wd_generic_too) const {
    PCG_REQUIRE(!tcs2 || perm_hw == 0, "permuted stride-2 dgrad packing unsupported");
    PackDesc d;
    d.w = w; d.Cout = src_cout; d.Cin = src_cin; d.taps = g.ksize * g.ksize; d.perm_hw = perm_hw; d.begin = 0;
    d.ld_cout = g.Cout; d.ld_cin = g.Cin;
    d.wf = (tc_fprop || to1_fprop || few_fprop) ? nullptr : wf;
    d.wd = (((tc_dgrad || tc_dgrad_s2) && !wd_generic_too) || to1_dgrad || few_dgrad || s2_to1_dgrad) ? nullptr : wd;
    d.tcf = tcf; d.tcd = tcd; d.tcs2 = tcs2;
    return d;
  }
};

static ConvEpilogue to_tc(const GenEpilogue<bf16>& e, float* stats) {
  ConvEpilogue c;
  c.bias = e.bias; c.act = e.act; c.slope = e.slope; c.add_src = e.add_src;
  c.act_ref = e.act_ref; c.ref_act = e.ref_act; c.ref_slope = e.ref_slope; c.stats = stats;
  return c;
}

// ---------------------------------------------------------------------------------------------
struct PlanBase {
  virtual ~PlanBase() {}
  virtual void refresh_weights(cudaStream_t s) = 0;
  virtual void step_d_grads(const pcg_mnist_inputs& in, float* scal, cudaStream_t s) = 0;
  virtual void step_d_update(cudaStream_t s) = 0;
  virtual void step_g_grads(const pcg_mnist_inputs& in, float* scal, cudaStream_t s) = 0;
  virtual void step_g_update(cudaStream_t s) = 0;
  virtual void step_c_bwd(cudaStream_t s) = 0;
  virtual void set_defer_c_bwd(bool on) = 0;
  virtual void step_g_grads_part(const pcg_mnist_inputs& in, float* scal, cudaStream_t s, int part, int split) = 0;
  virtual void g_forward(const float* x, const long long* target, const float* mask, int training, float* raw,
                         float* masked, cudaStream_t s) = 0;
  virtual void d_forward(const float* x, const long long* cond, float* logits, cudaStream_t s) = 0;
  virtual void c_forward(const float* x, float* logits, cudaStream_t s) = 0;
  virtual void debug_tensor(const std::string& name, void** ptr, long long* numel, int* dtype) = 0;
};

template <typename T>
struct MnistPlan : PlanBase {
  static constexpr bool kBf16 = std::is_same<T, bf16>::value;
  pcg_mnist_config cfg;
  pcg_mnist_buffers buf;
  int B, ch, nres;
  long long MG;                 // B * 784
  std::vector<void*> owned;
  std::map<std::string, std::pair<void*, std::pair<long long, int>>> dbg;

  // ---- layers
  ConvLayer<T> g_in, g_mid, g_out;
  std::vector<ConvLayer<T>> g_c1, g_c2;
  ConvLayer<T> d_conv[4];
  ConvLayer<T> c_conv[3], c_fc1, c_fc2;
  struct BN {
    const float *gamma, *beta;
    float *dgamma, *dbeta, *running_mean, *running_var;
    long long* nbt;
    float *mean, *rstd, *scale, *shift;
  };
  std::vector<BN> bn1, bn2;
  const float *g_embed, *d_embed, *d_head_w, *d_head_b;
  float *g_dembed, *d_dembed, *d_dhead_w, *d_dhead_b;

  // ---- workspaces
  T* inp3;                       // [B][784][3]
  std::vector<T*> h;             // nres + 1
  std::vector<T*> y1, z1, y2;
  T* hm;
  float *cimg, *raw, *masked, *x_cf;
  T *dhA, *dhB, *dz1;            // backward ping-pong
  T* g_hm;                       // gradient wrt conv_mid's output
  std::vector<T*> dy1, dy2;      // per-block gradients wrt the conv outputs (own buffers: the weight-gradient
                                 // stream reads them while the main stream runs ahead)
  // side streams: weight gradients (off the data-gradient critical path) and the classifier branch of the G step
  cudaStream_t side_w = nullptr, side_c = nullptr, aux[3] = {nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_next = 0;
  T* g_c;                        // [B][784] gradient wrt conv_out output
  float* dinp;                   // [B][784][3]
  // D (sized for 2B)
  T* a0;
  T* dz[4];
  T* dg[4];
  float *dlogits_d, *ddlogit, *dxd;
  long long* labels2;
  // C
  T* cz[3];
  T* cf1;
  float *clogits, *cdlogits, *dxc;
  T *cdf1, *cd3, *cd2, *cd1;
  // scratch
  float *stat_part3 = nullptr;                // partial rows of the bias gradient that runs on the weight-gradient stream
  float *stat_db = nullptr;                   // [2 * nres][STAT_PARTS][ch]: conv-bias gradient partials of the residual blocks
  bool batch_db = false;                      // PCG_BATCH_DB bit 0: one finalize launch for all of them at the end of the backward
  bool side_mid_db = false;                   // PCG_BATCH_DB bit 1: conv_mid's bias gradient on the weight-gradient stream
  float *stat_part, *stat_part2, *stat_bn2 = nullptr, *c12, *wg_scratch, *tc_part, *l1_part, *scal_tmp, *small_part;
  bool fuse_bn2_reduce = true, fuse_finalize = false, fuse_finalize_fwd = false, defer_c_bwd = false;
  unsigned int* fin_counter = nullptr;
  size_t wg_scratch_elems = 0;

  template <typename U>
  U* alloc(size_t n, const char* name = nullptr, int dtype = -1) {
    void* p = nullptr;
    PCG_CHECK_CUDA(cudaMalloc(&p, (n ? n : 1) * sizeof(U)));
    PCG_CHECK_CUDA(cudaMemset(p, 0, (n ? n : 1) * sizeof(U)));
    owned.push_back(p);
    if (name) dbg[name] = {p, {(long long)n, dtype >= 0 ? dtype : (std::is_same<U, bf16>::value ? PCG_BF16 : PCG_F32)}};
    return reinterpret_cast<U*>(p);
  }

  void setup_conv(ConvLayer<T>& L, const std::string& name, int N, int H, int W, int Cin, int Cout, int k, int stride,
                  int pad, const float* w, const float* b, float* dw, float* db, bool need_wd, int perm_hw = 0,
                  int cin_pad = 0, int cout_pad = 0) {
    L.name = name; L.tag_f = name + ".fprop"; L.tag_d = name + ".dgrad"; L.tag_w = name + ".wgrad";
    // cin_pad / cout_pad: the kernels see zero-padded channel counts (the packed weights and the activation maps
    // carry zero channels) so that a 32-channel layer can use the 64-channel tensor-core kernels
    L.src_cin = Cin; L.src_cout = Cout;
    if (cin_pad > Cin) Cin = cin_pad;
    if (cout_pad > Cout) {
      PCG_REQUIRE(dw == nullptr, "output-channel padding is for frozen layers");
      if (b != nullptr) {
        float* bp = alloc<float>(cout_pad);
        padded_bias.push_back({bp, b, Cout});
        b = bp;
      }
      Cout = cout_pad;
    }
    L.g = ConvGeom{N, H, W, Cin, Cout, k, stride, pad};
    L.w = w; L.b = b; L.dw = dw; L.db = db; L.perm_hw = perm_hw;
    const size_t n = (size_t)Cout * Cin * k * k;
    L.wf = alloc<float>(n);
    if (need_wd) L.wd = alloc<float>(n);
    if (dw) {
      const size_t sc = conv_wgrad_generic_scratch(L.g);
      if (sc > wg_scratch_elems) wg_scratch_elems = sc;
    }
    if (kBf16 && cfg.use_tensor_cores && Cin % 64 == 0 && Cout % 64 == 0) {
      L.tcf = alloc<bf16>(n);
      L.tc_fprop = true;
      // stride-1 data gradient = the same implicit GEMM on dY with transposed (and, for 3x3, rotated) weights;
      // a Linear (k = 1) may carry the NHWC-flatten permutation, which the packing applies to the Cin index
      if (stride == 1 && need_wd && ((k == 3 && pad == 1 && perm_hw == 0) || (k == 1 && pad == 0))) {
        L.tcd = alloc<bf16>(n);
        L.tc_dgrad = true;
      }
      if (Cin == 64 && Cout == 64 && k == 3 && stride == 1 && pad == 1 && dw) L.tc_wgrad = true;
      L.tc64 = Cin == 64 && Cout == 64 && k == 3 && stride == 1 && pad == 1 && conv_tc64_supported(H, W);
      if (dw && k == 3 && pad == 1 && !L.tc_wgrad) {
        L.tc_wgrad_gen = true;
        const size_t sc = (size_t)conv_tc_wgrad_general_splits(N, H, W, Cin, Cout, stride) * Cout * Cin * 9;
        if (sc > wg_scratch_elems) wg_scratch_elems = sc;
      }
    }
    if (kBf16 && cfg.use_tensor_cores && k == 3 && stride == 1 && pad == 1 && perm_hw == 0) {
      if (Cout == 1 && conv_to1_supported(H, W, Cin)) {
        L.tcf = alloc<bf16>(n);
        L.to1_fprop = true;
      }
      if (Cin <= 4 && need_wd && conv_to1_supported(H, W, Cout)) {
        L.tcd = alloc<bf16>(n);
        L.to1_dgrad = true;
      }
      if (Cout <= 3 && need_wd && conv_few_supported(Cout, Cin, k, 1, pad)) {
        if (!L.tcd) L.tcd = alloc<bf16>(n);
        L.few_dgrad = true;
      }
    }
    if (kBf16 && cfg.use_tensor_cores && k == 3 && stride == 2 && pad == 1 && Cin <= 4 && need_wd && perm_hw == 0 &&
        dgrad_s2_to1_supported(H, W, Cout)) {
      if (!L.tcd) L.tcd = alloc<bf16>(n);
      L.s2_to1_dgrad = true;
    }
    if (kBf16 && cfg.use_tensor_cores && dw != nullptr) {
      L.few_wgrad = wgrad_few_supported(Cin, Cout, k, stride, pad) && ((Cin == 3 && stride == 1) || (Cin == 2 && stride == 2));
      L.to1_wgrad = Cin == 64 && Cout == 1 && k == 3 && stride == 1 && pad == 1;
    }
    if (kBf16 && cfg.use_tensor_cores && perm_hw == 0 && conv_few_supported(Cin, Cout, k, stride, pad)) {
      if (!L.tcf) L.tcf = alloc<bf16>(n);
      L.few_fprop = true;
    }
    if (kBf16 && cfg.use_tensor_cores && stride == 2 && k == 3 && pad == 1 && Cout % 64 == 0 && Cin % 32 == 0 &&
        need_wd && perm_hw == 0) {
      L.tcs2 = alloc<bf16>(conv_tc_dgrad_s2_pack_elems(Cout, Cin));
      L.tc_dgrad_s2 = true;
    }
  }

  MnistPlan(const pcg_mnist_config& c, const pcg_mnist_buffers& bf) : cfg(c), buf(bf) {
    B = c.batch; ch = c.base_ch; nres = c.n_resblocks;
    PCG_REQUIRE(B > 0 && ch >= 4 && ch % 4 == 0 && (256 % (ch / 4)) == 0 && nres >= 0, "bad config");
    MG = (long long)B * 784;
    long long gt, dt, ct;
    auto gl = net_layout(0, ch, nres, &gt);
    auto dl = net_layout(1, ch, nres, &dt);
    auto cl = net_layout(2, ch, nres, &ct);
    auto GP = [&](int i) { return buf.g_params + gl[i].offset; };
    auto GG = [&](int i) { return buf.g_grads + gl[i].offset; };
    auto DP = [&](int i) { return buf.d_params + dl[i].offset; };
    auto DG = [&](int i) { return buf.d_grads + dl[i].offset; };
    auto CP = [&](int i) { return buf.c_params + cl[i].offset; };

    // ---- generator
    g_embed = GP(0); g_dembed = GG(0);
    setup_conv(g_in, "g.conv_in", B, 28, 28, 3, ch, 3, 1, 1, GP(1), GP(2), GG(1), GG(2), true);
    g_c1.resize(nres); g_c2.resize(nres); bn1.resize(nres); bn2.resize(nres);
    for (int i = 0; i < nres; ++i) {
      const int t = 3 + i * 8;
      setup_conv(g_c1[i], "g.res", B, 28, 28, ch, ch, 3, 1, 1, GP(t), GP(t + 1), GG(t), GG(t + 1), true);
      setup_conv(g_c2[i], "g.res", B, 28, 28, ch, ch, 3, 1, 1, GP(t + 4), GP(t + 5), GG(t + 4), GG(t + 5), true);
      BN* bns[2] = {&bn1[i], &bn2[i]};
      for (int j = 0; j < 2; ++j) {
        BN& q = *bns[j];
        q.gamma = GP(t + 2 + 4 * j); q.beta = GP(t + 3 + 4 * j);
        q.dgamma = GG(t + 2 + 4 * j); q.dbeta = GG(t + 3 + 4 * j);
        q.running_mean = buf.g_bn_running + (size_t)(2 * i + j) * 2 * ch;
        q.running_var = q.running_mean + ch;
        q.nbt = buf.g_bn_nbt + (2 * i + j);
        q.mean = alloc<float>(ch); q.rstd = alloc<float>(ch); q.scale = alloc<float>(ch); q.shift = alloc<float>(ch);
      }
    }
    {
      const int t = 3 + nres * 8;
      setup_conv(g_mid, "g.res", B, 28, 28, ch, ch, 3, 1, 1, GP(t), GP(t + 1), GG(t), GG(t + 1), true);
      setup_conv(g_out, "g.conv_out", B, 28, 28, ch, 1, 3, 1, 1, GP(t + 2), GP(t + 3), GG(t + 2), GG(t + 3), true);
    }
    // ---- discriminator (batched real+fake in the D step -> 2B)
    d_embed = DP(0); d_dembed = DG(0);
    {
      const int cin[4] = {2, 64, 128, 256}, cout[4] = {64, 128, 256, 256}, hw[4] = {28, 14, 7, 4};
      for (int l = 0; l < 4; ++l)
        setup_conv(d_conv[l], "d.conv" + std::to_string(l), 2 * B, hw[l], hw[l], cin[l], cout[l], 3, 2, 1, DP(1 + l), nullptr, DG(1 + l), nullptr, true);
    }
    d_head_w = DP(5); d_head_b = DP(6); d_dhead_w = DG(5); d_dhead_b = DG(6);
    // ---- classifier (frozen; data gradients only)
    // the classifier's 32-channel map is carried as 64 channels (upper half zero) in tensor-core mode: conv1 then is a
    // 64 -> 64 stride-2 convolution the tcgen05 im2col kernel takes
    const int c1 = (kBf16 && cfg.use_tensor_cores) ? 64 : 32;
    setup_conv(c_conv[0], "c.conv0", B, 28, 28, 1, 32, 3, 1, 1, CP(0), CP(1), nullptr, nullptr, true, 0, 0, c1);
    setup_conv(c_conv[1], "c.conv1", B, 28, 28, 32, 64, 3, 2, 1, CP(2), CP(3), nullptr, nullptr, true, 0, c1, 0);
    setup_conv(c_conv[2], "c.conv2", B, 14, 14, 64, 128, 3, 2, 1, CP(4), CP(5), nullptr, nullptr, true);
    setup_conv(c_fc1, "c.fc1", B, 1, 1, 6272, 256, 1, 1, 0, CP(6), CP(7), nullptr, nullptr, true, /*perm_hw=*/49);
    setup_conv(c_fc2, "c.fc2", B, 1, 1, 256, 10, 1, 1, 0, CP(8), CP(9), nullptr, nullptr, true);

    // ---- workspaces
    const size_t act = (size_t)MG * ch;
    inp3 = alloc<T>((size_t)MG * 3, "inp3");
    h.resize(nres + 1); y1.resize(nres); z1.resize(nres); y2.resize(nres);
    h[0] = alloc<T>(act, "h.0");
    for (int i = 0; i < nres; ++i) {
      y1[i] = alloc<T>(act, ("y1." + std::to_string(i)).c_str());
      z1[i] = alloc<T>(act, ("z1." + std::to_string(i)).c_str());
      y2[i] = alloc<T>(act, ("y2." + std::to_string(i)).c_str());
      h[i + 1] = alloc<T>(act, ("h." + std::to_string(i + 1)).c_str());
    }
    hm = alloc<T>(act, "hm");
    cimg = alloc<float>(MG, "cimg"); raw = alloc<float>(MG, "raw"); masked = alloc<float>(MG, "masked");
    x_cf = alloc<float>(MG, "x_cf");
    dhA = alloc<T>(act, "dhA"); dhB = alloc<T>(act, "dhB"); dz1 = alloc<T>(act, "dz1");
    g_hm = alloc<T>(act, "g_hm");
    dy1.resize(nres); dy2.resize(nres);
    for (int i = 0; i < nres; ++i) { dy1[i] = alloc<T>(act); dy2[i] = alloc<T>(act); }
    dbg["dy"] = {nres ? dy1[0] : nullptr, {(long long)act, kBf16 ? PCG_BF16 : PCG_F32}};
    PCG_CHECK_CUDA(cudaStreamCreateWithFlags(&side_w, cudaStreamNonBlocking));
    PCG_CHECK_CUDA(cudaStreamCreateWithFlags(&side_c, cudaStreamNonBlocking));
    for (auto& a : aux) PCG_CHECK_CUDA(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    ev_pool.resize(256);
    for (auto& e : ev_pool) PCG_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_c = alloc<T>(MG, "g_c");
    dinp = alloc<float>((size_t)MG * 3, "dinp");
    a0 = alloc<T>((size_t)2 * MG * 2, "a0");
    {
      const int cout[4] = {64, 128, 256, 256}, hw[4] = {14, 7, 4, 2};
      for (int l = 0; l < 4; ++l) {
        const size_t n = (size_t)2 * B * hw[l] * hw[l] * cout[l];
        dz[l] = alloc<T>(n, ("dz." + std::to_string(l)).c_str());
        dg[l] = alloc<T>(n, ("dg." + std::to_string(l)).c_str());
      }
    }
    dlogits_d = alloc<float>(2 * B, "d_logits"); ddlogit = alloc<float>(2 * B, "d_dlogit");
    dxd = alloc<float>((size_t)2 * MG * 2, "dxd");
    labels2 = alloc<long long>(2 * B);
    cz[0] = alloc<T>((size_t)MG * c_conv[0].g.Cout, "c.0"); cz[1] = alloc<T>((size_t)B * 196 * 64, "c.1");
    cz[2] = alloc<T>((size_t)B * 49 * 128, "c.2"); cf1 = alloc<T>((size_t)B * 256, "c.f1");
    clogits = alloc<float>((size_t)B * 10, "c_logits"); cdlogits = alloc<float>((size_t)B * 10, "c_dlogits");
    dxc = alloc<float>(MG, "dxc");
    cdf1 = alloc<T>((size_t)B * 256); cd3 = alloc<T>((size_t)B * 49 * 128); cd2 = alloc<T>((size_t)B * 196 * 64);
    cd1 = alloc<T>((size_t)MG * c_conv[0].g.Cout);
    const int maxC = ch > 256 ? ch : 256;
    stat_part = alloc<float>((size_t)STAT_PARTS * 2 * maxC);
    stat_part2 = alloc<float>((size_t)STAT_PARTS * 2 * maxC);
    {
      const char* e = getenv("PCG_BATCH_DB");
      const int mode = e ? atoi(e) : 0;            // default off: measured slower; bit 0 / bit 1
      batch_db = (mode & 1) != 0 && 2 * nres <= ColsumOuts::MAX;
      side_mid_db = (mode & 2) != 0;
      if (batch_db) stat_db = alloc<float>((size_t)2 * nres * STAT_PARTS * ch);
      if (side_mid_db) stat_part3 = alloc<float>((size_t)STAT_PARTS * maxC);
    }
    c12 = alloc<float>(2 * maxC);
    wg_scratch = alloc<float>(wg_scratch_elems);
    PCG_REQUIRE(sm_count() <= STAT_PARTS, "more SMs than partial-row slots (STAT_PARTS)");
    tc_part = kBf16 ? alloc<float>((size_t)sm_count() * 9 * 64 * 64) : nullptr;   // one weight-gradient partial per CTA (grid <= SMs)
    l1_part = alloc<float>(STAT_PARTS * 2);
    stat_bn2 = alloc<float>((size_t)STAT_PARTS * 2 * maxC);
    fin_counter = alloc<unsigned int>(4);
    {
      const char* e = getenv("PCG_FUSE_FINALIZE");  // A/B switch: 1 = the last CTA of the convolution finishes the statistics
      fuse_finalize = e && atoi(e) == 1;       // default OFF: measured slower (negative results)
      fuse_finalize_fwd = e && (atoi(e) == 1 || atoi(e) == 2);   // 2: the forward statistics only (no weight-gradient stream there)
    }
    {
      const char* e = getenv("PCG_FUSE_BN2");      // A/B switch: 0 = separate bn_bwd_partial pass (round-1 schedule)
      fuse_bn2_reduce = !(e && atoi(e) == 0);
    }
    small_part = alloc<float>((size_t)sm_count() * 2048);
    scal_tmp = alloc<float>(16);
    dbg["dinp"] = {dinp, {MG * 3, PCG_F32}};
    build_pack_tables();
    build_fold_table();
  }

  ~MnistPlan() override {
    for (void* p : owned) cudaFree(p);
    for (auto e : ev_pool) cudaEventDestroy(e);
    if (side_w) cudaStreamDestroy(side_w);
    if (side_c) cudaStreamDestroy(side_c);
    for (auto a : aux) if (a) cudaStreamDestroy(a);
  }
  // Per-launcher profiling (bench.py roofline pass) times kernels one at a time: everything stays on the caller's stream.
  cudaStream_t wstream(cudaStream_t s) const { return g_profile_on ? s : side_w; }
  cudaStream_t cstream(cudaStream_t s) const { return g_profile_on ? s : side_c; }
  // `to` continues after everything enqueued on `from` so far (a graph edge under stream capture)
  void after(cudaStream_t from, cudaStream_t to) {
    if (from == to) return;
    cudaEvent_t e = ev_pool[ev_next++ % ev_pool.size()];
    PCG_CHECK_CUDA(cudaEventRecord(e, from));
    PCG_CHECK_CUDA(cudaStreamWaitEvent(to, e, 0));
  }

  // ---------------------------------------------------------------- conv dispatch
  template <typename TIn, typename TOut>
  void fprop(const ConvLayer<T>& L, const TIn* in, GenEpilogue<TOut> e, TOut* out, cudaStream_t s, float* stats = nullptr,
             int* nparts = nullptr, int n_override = 0) {
    ProfTag _tag(L.tag_f.c_str());
    ConvGeom g = L.g;
    if (n_override) g.N = n_override;
    if constexpr (kBf16 && std::is_same<TIn, bf16>::value && std::is_same<TOut, bf16>::value) {
      if (L.tc_fprop && L.tc64) {
        conv_tc64_fprop(in, g.N, g.H, g.W, L.tcf, to_tc(e, stats), out, s);
        if (nparts) *nparts = conv_tc64_fprop_grid(g.N, g.H, g.W);
        return;
      }
      if (L.tc_fprop) {
        conv_tc_fprop(in, g.N, g.H, g.W, g.Cin, L.tcf, g.Cout, g.ksize, g.stride, g.pad, to_tc(e, stats), out, s);
        if (nparts) *nparts = conv_tc_grid(g.Mout(), g.Cout);
        return;
      }
    }
    if constexpr (kBf16 && std::is_same<TOut, bf16>::value) {
      if (L.few_fprop && e.add_src == nullptr && stats == nullptr) {
        FewEpilogue fe;
        fe.bias = e.bias; fe.act = e.act; fe.slope = e.slope;
        fe.act_ref = e.act_ref; fe.ref_act = e.ref_act; fe.ref_slope = e.ref_slope;
        conv_few<TIn>(in, g.N, g.H, g.W, g.Cin, L.tcf, g.Cout, g.stride, fe, out, s);
        return;
      }
    }
    if constexpr (kBf16 && std::is_same<TIn, bf16>::value && std::is_same<TOut, float>::value) {
      if (L.to1_fprop && e.act == ACT_NONE && e.add_src == nullptr && e.act_ref == nullptr && stats == nullptr) {
        conv_to1(in, g.N, g.H, g.W, g.Cin, L.tcf, e.bias, out, s);
        return;
      }
    }
    conv_fprop_generic<TIn, TOut>(in, g, L.wf, e, out, s);
    if (stats) {
      if constexpr (std::is_same<TOut, T>::value) bn_stats_partial<T>(out, g.Mout(), g.Cout, stats, s);
      *nparts = STAT_PARTS;
    }
  }
  template <typename TIn, typename TOut>
  void dgrad(const ConvLayer<T>& L, const TIn* dout, GenEpilogue<TOut> e, TOut* din, cudaStream_t s, int n_override = 0,
             int ch_select = -1) {
    ProfTag _tag(L.tag_d.c_str());
    ConvGeom g = L.g;
    if (n_override) g.N = n_override;
    if constexpr (kBf16 && std::is_same<TIn, bf16>::value && std::is_same<TOut, float>::value) {
      if (L.s2_to1_dgrad && (ch_select >= 0 || g.Cin == 1) && e.bias == nullptr && e.act == ACT_NONE &&
          e.add_src == nullptr && e.act_ref == nullptr) {
        const int ci = ch_select >= 0 ? ch_select : 0;
        dgrad_s2_to1(dout, g.N, g.H, g.W, L.tcd + (size_t)ci * 9 * g.Cout, din, s);
        return;
      }
      if (L.to1_dgrad && (ch_select >= 0 || g.Cin == 1) && e.bias == nullptr && e.act == ACT_NONE &&
          e.add_src == nullptr && e.act_ref == nullptr) {
        const int ci = ch_select >= 0 ? ch_select : 0;
        conv_to1(dout, g.N, g.H, g.W, g.Cout, L.tcd + (size_t)ci * 9 * g.Cout, nullptr, din, s);
        return;
      }
    }
    if constexpr (kBf16 && std::is_same<TIn, bf16>::value && std::is_same<TOut, bf16>::value) {
      if (L.few_dgrad && ch_select < 0 && e.bias == nullptr && e.act == ACT_NONE && e.add_src == nullptr) {
        FewEpilogue fe;
        fe.act_ref = e.act_ref; fe.ref_act = e.ref_act; fe.ref_slope = e.ref_slope;
        conv_few<bf16>(dout, g.N, g.H, g.W, g.Cout, L.tcd, g.Cin, 1, fe, din, s);
        return;
      }
    }
    if (ch_select >= 0) {
      conv_dgrad_generic<TIn, TOut>(dout, g, L.wd, e, din, s, ch_select);
      return;
    }
    if constexpr (kBf16 && std::is_same<TIn, bf16>::value && std::is_same<TOut, bf16>::value) {
      if (L.tc_dgrad && L.tc64) {
        conv_tc64_fprop(dout, g.N, g.H, g.W, L.tcd, to_tc(e, nullptr), din, s);
        return;
      }
      if (L.tc_dgrad) {
        conv_tc_fprop(dout, g.N, g.H, g.W, g.Cout, L.tcd, g.Cin, g.ksize, 1, g.pad, to_tc(e, nullptr), din, s);
        return;
      }
      if (L.tc_dgrad_s2) {
        // small problems (a class needs at most a quarter of the SMs): the four parity classes side by side
        if (!g_profile_on && 4 * conv_tc_dgrad_s2_class_ctas(g.N, g.H, g.W, g.Cin) <= sm_count() + 20) {
          cudaStream_t cs[4] = {s, aux[0], aux[1], aux[2]};
          for (int i = 0; i < 3; ++i) after(s, aux[i]);
          conv_tc_dgrad_s2(dout, g.N, g.H, g.W, g.Cin, g.Cout, L.tcs2, to_tc(e, nullptr), din, s, cs);
          for (int i = 0; i < 3; ++i) after(aux[i], s);
        } else {
          conv_tc_dgrad_s2(dout, g.N, g.H, g.W, g.Cin, g.Cout, L.tcs2, to_tc(e, nullptr), din, s);
        }
        return;
      }
    }
    conv_dgrad_generic<TIn, TOut>(dout, g, L.wd, e, din, s);
  }
  template <typename TIn, typename TDy>
  void wgrad(const ConvLayer<T>& L, const TIn* in, const TDy* dout, cudaStream_t s, int n_override = 0) {
    ProfTag _tag(L.tag_w.c_str());
    ConvGeom g = L.g;
    if (n_override) g.N = n_override;
    if constexpr (kBf16 && std::is_same<TIn, bf16>::value && std::is_same<TDy, bf16>::value) {
      if (L.few_wgrad) {
        wgrad_few<bf16>(in, dout, g.N, g.H, g.W, g.Cin, g.stride, small_part, L.dw, L.db, s);
        return;
      }
      if (L.to1_wgrad) {
        wgrad_to1(in, dout, g.N, g.H, g.W, small_part, L.dw, L.db, s);
        return;
      }
      if (L.tc_wgrad && L.tc64) {
        conv_tc64_wgrad(in, dout, g.N, g.H, g.W, tc_part, s);
        wgrad_reduce_tc(tc_part, conv_tc64_grid(g.N, g.H, g.W), L.dw, s, conv_tc64_wgrad_swizzled());
        return;
      }
      if (L.tc_wgrad_gen) {
        conv_tc_wgrad_general(in, dout, g.N, g.H, g.W, g.Cin, g.Cout, g.stride, wg_scratch, s);
        wgrad_reduce_generic(wg_scratch, conv_tc_wgrad_general_splits(g.N, g.H, g.W, g.Cin, g.Cout, g.stride), g.Cout,
                             g.Cin, 9, L.dw, s);
        return;
      }
      if (L.tc_wgrad) {
        conv_tc_wgrad64(in, dout, g.N, g.H, g.W, tc_part, s);
        wgrad_reduce_tc(tc_part, conv_tc_wgrad_grid(g.Mout()), L.dw, s);
        return;
      }
    }
    conv_wgrad_generic<TIn, TDy>(in, dout, g, wg_scratch, L.dw, s);
  }
  void bias_grad(const T* dy, long long M, int C, float* db, cudaStream_t s, float* part = nullptr) {
    if (part == nullptr) part = stat_part2;
    colsum_partial<T>(dy, M, C, part, s);
    colsum_finalize(part, STAT_PARTS, C, C, db, s);
  }

  PackTable pack_g, pack_d, pack_c;
  struct PaddedBias { float* dst; const float* src; int n; };
  std::vector<PaddedBias> padded_bias;
  PackTable make_pack_table(const std::vector<PackDesc>& layers) {
    PCG_REQUIRE((int)layers.size() <= PACK_MAX_LAYERS, "too many layers for one pack table");
    std::vector<PackDesc> v = layers;
    long long off = 0;
    for (auto& d : v) {
      d.begin = (int)off;
      off += (long long)d.Cout * d.Cin * d.taps;
    }
    PackTable t;
    t.nlayers = (int)v.size(); t.total = (int)off;
    t.dev = alloc<PackDesc>(v.size());
    PCG_CHECK_CUDA(cudaMemcpy(t.dev, v.data(), v.size() * sizeof(PackDesc), cudaMemcpyHostToDevice));
    return t;
  }
  void build_pack_tables() {
    std::vector<PackDesc> g{g_in.desc(true)};
    for (int i = 0; i < nres; ++i) { g.push_back(g_c1[i].desc(false)); g.push_back(g_c2[i].desc(false)); }
    g.push_back(g_mid.desc(false)); g.push_back(g_out.desc(false));
    pack_g = make_pack_table(g);
    std::vector<PackDesc> d;
    for (int l = 0; l < 4; ++l) d.push_back(d_conv[l].desc(l == 0));
    pack_d = make_pack_table(d);
    pack_c = make_pack_table({c_conv[0].desc(false), c_conv[1].desc(false), c_conv[2].desc(false), c_fc1.desc(false),
                              c_fc2.desc(false)});
  }
  // ---- eval-mode forward with folded BatchNorm (tensor-core plans with the halo-tile kernels only)
  FoldDesc* fold_dev = nullptr;
  int fold_n = 0;
  void build_fold_table() {
    if (!kBf16 || nres == 0 || !g_c1[0].tc64) return;
    std::vector<FoldDesc> v;
    for (int i = 0; i < nres; ++i)
      for (int j = 0; j < 2; ++j) {
        ConvLayer<T>& L = j ? g_c2[i] : g_c1[i];
        const BN& q = j ? bn2[i] : bn1[i];
        L.tcf_eval = alloc<bf16>((size_t)ch * ch * 9);
        L.b_eval = alloc<float>(ch);
        FoldDesc d;
        d.w = L.w; d.b = L.b; d.gamma = q.gamma; d.beta = q.beta; d.rm = q.running_mean; d.rv = q.running_var;
        d.wq = L.tcf_eval; d.bq = L.b_eval; d.extra = j ? 0.1f : 1.f;      // generator.py:22: x + 0.1 * out
        d.Cout = ch; d.Cin = ch; d.taps = 9;
        v.push_back(d);
      }
    fold_n = (int)v.size();
    fold_dev = alloc<FoldDesc>(v.size());
    PCG_CHECK_CUDA(cudaMemcpy(fold_dev, v.data(), v.size() * sizeof(FoldDesc), cudaMemcpyHostToDevice));
  }
  void g_fwd_eval(const float* x, const long long* target, const float* mask, cudaStream_t s) {
    {
      PCG_PROFILE("pack_weights", s);
      launch_k(fold_bn_kernel, dim3(dim3(36, fold_n)), dim3(256), 0, s, fold_dev, 1e-5f);
      PCG_COUNT_LAUNCH();
      PCG_LAUNCH_CHECK();
    }
    g_input<T>(x, g_embed, target, mask, B, 784, inp3, s);
    GenEpilogue<T> e;
    e.bias = g_in.b; e.act = ACT_LRELU; e.slope = 0.2f;
    fprop<T, T>(g_in, inp3, e, h[0], s);
    if constexpr (kBf16) {
      for (int i = 0; i < nres; ++i) {
        ConvEpilogue e1;                               // z1 = LReLU(BN1(conv1(h)))
        e1.bias = g_c1[i].b_eval; e1.act = ACT_LRELU; e1.slope = 0.2f;
        conv_tc64_fprop(h[i], B, 28, 28, g_c1[i].tcf_eval, e1, z1[i], s);
        ConvEpilogue e2;                               // h' = h + 0.1 * BN2(conv2(z1))
        e2.bias = g_c2[i].b_eval; e2.add_src = h[i];
        conv_tc64_fprop(z1[i], B, 28, 28, g_c2[i].tcf_eval, e2, h[i + 1], s);
      }
    }
    GenEpilogue<T> em; em.bias = g_mid.b; em.act = ACT_LRELU; em.slope = 0.2f;
    fprop<T, T>(g_mid, h[nres], em, hm, s);
    GenEpilogue<float> eo; eo.bias = g_out.b;
    fprop<T, float>(g_out, hm, eo, cimg, s);
    residual_head_fwd(cimg, x, mask, cfg.residual_scaling, MG, raw, masked, x_cf, l1_part, s);
  }
  void refresh_g(cudaStream_t s) { pack_g.launch(s); }
  void refresh_d(cudaStream_t s) { pack_d.launch(s); }
  void refresh_weights(cudaStream_t s) override {
    refresh_g(s); refresh_d(s);
    pack_c.launch(s);
    for (const auto& pb : padded_bias)
      PCG_CHECK_CUDA(cudaMemcpyAsync(pb.dst, pb.src, pb.n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }

  // ---------------------------------------------------------------- generator forward
  void bn_coeffs(const BN& q, const float* part, int nparts, bool training, cudaStream_t s) {
    if (training) {
      bn_finalize(part, nparts, MG, ch, q.gamma, q.beta, 1e-5f, 0.1f, q.running_mean, q.running_var, q.nbt, q.mean,
                  q.rstd, q.scale, q.shift, s);
    } else {
      // eval: scale = gamma / sqrt(running_var + eps), shift = beta - running_mean*scale; expressed as
      // a one-row "partial" so the same finalize kernel serves: sum = mean*M, sumsq = (var+mean^2)*M
      bn_eval_coeffs(q, s);
    }
  }
  void bn_eval_coeffs(const BN& q, cudaStream_t s);

  void g_fwd(const float* x, const long long* target, const float* mask, bool training, cudaStream_t s) {
    g_input<T>(x, g_embed, target, mask, B, 784, inp3, s);
    GenEpilogue<T> e;
    e.bias = g_in.b; e.act = ACT_LRELU; e.slope = 0.2f;
    fprop<T, T>(g_in, inp3, e, h[0], s);
    // convolution + the statistics of the BatchNorm behind it; on the tensor-core path the last CTA of the convolution
    // also finishes them (mean, rstd, scale, shift, running buffers), which removes a launch from the serial chain
    auto conv_bn = [&](const ConvLayer<T>& L, const T* in, T* out, const BN& q) {
      if constexpr (kBf16) {
        if (training && L.tc_fprop && L.tc64 && fuse_finalize_fwd) {
          ProfTag _tag(L.tag_f.c_str());
          ConvEpilogue c;
          c.bias = L.b; c.stats = stat_part;
          c.fin.mode = 1; c.fin.counter = fin_counter; c.fin.M = MG;
          c.fin.gamma = q.gamma; c.fin.beta = q.beta; c.fin.eps = 1e-5f; c.fin.momentum = 0.1f;
          c.fin.running_mean = q.running_mean; c.fin.running_var = q.running_var; c.fin.nbt = q.nbt;
          c.fin.mean = q.mean; c.fin.rstd = q.rstd; c.fin.scale = q.scale; c.fin.shift = q.shift;
          conv_tc64_fprop(in, B, L.g.H, L.g.W, L.tcf, c, out, s);
          return;
        }
      }
      int np = 0;
      GenEpilogue<T> e; e.bias = L.b;
      if (training) fprop<T, T>(L, in, e, out, s, stat_part, &np);
      else fprop<T, T>(L, in, e, out, s);
      bn_coeffs(q, stat_part, np, training, s);
    };
    for (int i = 0; i < nres; ++i) {
      conv_bn(g_c1[i], h[i], y1[i], bn1[i]);
      bn_apply_act<T>(y1[i], bn1[i].scale, bn1[i].shift, MG, ch, ACT_LRELU, 0.2f, z1[i], s);
      conv_bn(g_c2[i], z1[i], y2[i], bn2[i]);
      bn_apply_residual<T>(y2[i], h[i], bn2[i].scale, bn2[i].shift, 0.1f, MG, ch, h[i + 1], s);
    }
    GenEpilogue<T> em; em.bias = g_mid.b; em.act = ACT_LRELU; em.slope = 0.2f;
    fprop<T, T>(g_mid, h[nres], em, hm, s);
    GenEpilogue<float> eo; eo.bias = g_out.b;
    fprop<T, float>(g_out, hm, eo, cimg, s);
    residual_head_fwd(cimg, x, mask, cfg.residual_scaling, MG, raw, masked, x_cf, l1_part, s);
  }

  // ---------------------------------------------------------------- discriminator
  void d_fwd(int n, cudaStream_t s) {          // a0[0:n] already assembled
    const T* in = a0;
    for (int l = 0; l < 4; ++l) {
      GenEpilogue<T> e; e.act = ACT_LRELU; e.slope = 0.2f;
      fprop<T, T>(d_conv[l], in, e, dz[l], s, nullptr, nullptr, n);
      in = dz[l];
    }
    d_head_fwd<T>(dz[3], n, 4, 256, d_head_w, d_head_b, dlogits_d, s);
  }
  // backward from ddlogit; weight grads written when `wg`; dxd[n][784] = gradient wrt input channel `in_ch`
  // (0 = image, needed by the G step; 1 = label-embedding map, needed by the D step)
  // Weight gradients run on side_w: each only needs dg[l] (written once per pass) and a forward activation, so the
  // data-gradient chain on `s` never waits for them; the caller joins side_w before the optimizer step.
  void d_bwd(int n, bool wg, int in_ch, cudaStream_t s) {
    cudaStream_t side_w = wstream(s);
    d_head_bwd<T>(dz[3], ddlogit, n, 4, 256, d_head_w, 0.2f, dg[3], wg ? d_dhead_w : nullptr, wg ? d_dhead_b : nullptr,
                  stat_part2, s);
    for (int l = 3; l >= 1; --l) {
      if (wg) {
        after(s, side_w);
        wgrad<T, T>(d_conv[l], dz[l - 1], dg[l], side_w, n);
      }
      GenEpilogue<T> e; e.act_ref = dz[l - 1]; e.ref_act = ACT_LRELU; e.ref_slope = 0.2f;
      dgrad<T, T>(d_conv[l], dg[l], e, dg[l - 1], s, n);
    }
    if (wg) {
      after(s, side_w);
      wgrad<T, T>(d_conv[0], a0, dg[0], side_w, n);
    }
    GenEpilogue<float> e0;
    dgrad<T, float>(d_conv[0], dg[0], e0, dxd, s, n, in_ch);
  }

  // ---------------------------------------------------------------- phases
  // Frozen classifier forward + input gradient (trainer.py:118, 121): lambda_cls * CE(C(x_cf), target) -> G_CLS, dxc.
  // It needs nothing but x_cf (the classifier is in eval mode and is never updated), so it runs on side_c beside the
  // whole discriminator step instead of in front of the generator backward: both are chains of small kernels that do
  // not fill the GPU alone.
  // Data parallel: the input-gradient half can be deferred (defer_c_bwd) and run as its own phase (step_c_bwd) on a side
  // stream BESIDE the all-reduce of the discriminator's gradients, which nothing else in the step can overlap (the
  // generator phase needs the updated discriminator).
  void c_branch(const pcg_mnist_inputs& in, float* scal, cudaStream_t c) {
    c_fwd(x_cf, c);
    ce_loss(clogits, in.target, B, 10, cfg.lambda_cls, scal + PCG_S_G_CLS, cdlogits, c);
    if (!defer_c_bwd) c_branch_bwd(c);
  }
  void step_c_bwd(cudaStream_t s) override { c_branch_bwd(s); }
  void set_defer_c_bwd(bool on) override { defer_c_bwd = on; }
  void c_branch_bwd(cudaStream_t c) {
    GenEpilogue<T> e; e.ref_act = ACT_RELU;
    e.act_ref = cf1; dgrad<float, T>(c_fc2, cdlogits, e, cdf1, c);
    e.act_ref = cz[2]; dgrad<T, T>(c_fc1, cdf1, e, cd3, c);
    e.act_ref = cz[1]; dgrad<T, T>(c_conv[2], cd3, e, cd2, c);
    e.act_ref = cz[0]; dgrad<T, T>(c_conv[1], cd2, e, cd1, c);
    GenEpilogue<float> e0;
    dgrad<T, float>(c_conv[0], cd1, e0, dxc, c);
  }

  void step_d_grads(const pcg_mnist_inputs& in, float* scal, cudaStream_t s) override {
    cudaStream_t side_w = wstream(s), side_c = cstream(s);
    // The real half of the discriminator's input and the label vector need nothing from the generator: they are
    // assembled on side_c beside the generator forward instead of between it and the discriminator forward.
    auto d_real_inputs = [&](cudaStream_t c) {
      d_input<T>(in.x, d_embed, in.y, B, 784, a0, c);
      PCG_CHECK_CUDA(cudaMemcpyAsync(labels2, in.y, B * sizeof(long long), cudaMemcpyDeviceToDevice, c));
      PCG_CHECK_CUDA(cudaMemcpyAsync(labels2 + B, in.target, B * sizeof(long long), cudaMemcpyDeviceToDevice, c));
    };
    cudaEvent_t pre = nullptr;
    if (side_c != s) {
      after(s, side_c);
      d_real_inputs(side_c);
      pre = ev_pool[ev_next++ % ev_pool.size()];
      PCG_CHECK_CUDA(cudaEventRecord(pre, side_c));
    }
    g_fwd(in.x, in.target, in.mask, true, s);
    after(s, side_c);
    l1_finalize(l1_part, STAT_PARTS, 1.f / (float)MG, scal + PCG_S_REG_L1, side_c);   // writes REG_L1, MASK_PEN
    c_branch(in, scal, side_c);
    if (pre != nullptr) PCG_CHECK_CUDA(cudaStreamWaitEvent(s, pre, 0));
    else d_real_inputs(s);
    d_input<T>(x_cf, d_embed, in.target, B, 784, a0 + (size_t)MG * 2, s);
    d_fwd(2 * B, s);
    bce_logits(dlogits_d, B, 2, 1.f, 0.f, 1.f, 1.f, scal + PCG_S_D_LOSS_REAL, scal + PCG_S_D_REAL_P, ddlogit, s);
    after(s, side_w);                              // the loss scalar is off the gradient chain
    g_loss_combine(scal + PCG_S_D_LOSS_REAL, scal + PCG_S_D_LOSS_FAKE, scal + PCG_S_D_LOSS_REAL,
                   scal + PCG_S_D_LOSS_REAL, 1.f, 1.f, 0.f, 0.f, scal + PCG_S_D_LOSS, side_w);
    d_bwd(2 * B, true, 1, s);
    embed_grad<float>(dxd, 1, 0, labels2, 2 * B, 784, 10, d_dembed, s);
    after(side_w, s);
    after(side_c, s);
  }

  void step_d_update(cudaStream_t s) override {
    long long dt;
    net_layout(1, ch, nres, &dt);
    adam_flat(buf.d_params, buf.d_grads, buf.d_adam_m, buf.d_adam_v, dt, buf.d_step, cfg.d_lr, cfg.beta1, cfg.beta2,
              cfg.adam_eps, cfg.grad_scale, s);
    refresh_d(s);
  }

  void c_fwd(const float* x, cudaStream_t s) {
    GenEpilogue<T> e; e.act = ACT_RELU;
    e.bias = c_conv[0].b; fprop<float, T>(c_conv[0], x, e, cz[0], s);
    e.bias = c_conv[1].b; fprop<T, T>(c_conv[1], cz[0], e, cz[1], s);
    e.bias = c_conv[2].b; fprop<T, T>(c_conv[2], cz[1], e, cz[2], s);
    e.bias = c_fc1.b; fprop<T, T>(c_fc1, cz[2], e, cf1, s);
    GenEpilogue<float> eo; eo.bias = c_fc2.b;
    fprop<T, float>(c_fc2, cf1, eo, clogits, s);
  }

  // part 0: the whole phase.  Data parallel: part 1 = everything down to (and including) residual block `split`, with
  // the weight-gradient stream joined, so the gradients of [block split .. conv_out] - the tail of the flat arena - can
  // be all-reduced while part 2 (blocks split-1 .. 0, conv_in, the embedding) still runs.
  void step_g_grads(const pcg_mnist_inputs& in, float* scal, cudaStream_t s) override { g_grads(in, scal, s, 0, 0); }
  void step_g_grads_part(const pcg_mnist_inputs& in, float* scal, cudaStream_t s, int part, int split) override {
    PCG_REQUIRE((part == 1 || part == 2) && split >= 1 && split < nres, "g_grads part 1|2, 1 <= split < n_resblocks");
    g_grads(in, scal, s, part, split);
  }
  T* gg_dh = nullptr;
  T* gg_dh_other = nullptr;
  int gg_bn2_parts = 0;
  void g_grads(const pcg_mnist_inputs& in, float* scal, cudaStream_t s, int part, int split) {
    cudaStream_t side_w = wstream(s);
    T*& dh = gg_dh;
    T*& dh_other = gg_dh_other;
    int& bn2_parts = gg_bn2_parts;
    auto dgrad_bn2 = [&](const ConvLayer<T>& L, const T* dout, const T* add, int blk, T* din) -> bool {
      if constexpr (kBf16) {
        if (L.tc_dgrad && L.tc64 && fuse_bn2_reduce) {
          ProfTag _tag("g.res.dgrad_bnred");
          const BN& qb = bn2[blk];
          ConvEpilogue c;
          c.add_src = add;
          c.stats = stat_bn2;
          c.bn_y = y2[blk]; c.bn_mean = qb.mean; c.bn_rstd = qb.rstd; c.bn_scale = qb.scale; c.bn_shift = qb.shift;
          c.bn_act = ACT_NONE; c.bn_gscale = 0.1f;
          if (fuse_finalize) {
            c.fin.mode = 2; c.fin.counter = fin_counter; c.fin.M = MG;
            c.fin.dgamma = qb.dgamma; c.fin.dbeta = qb.dbeta; c.fin.c12 = c12;
          }
          conv_tc64_fprop(dout, B, 28, 28, L.tcd, c, din, s);
          bn2_parts = conv_tc64_fprop_grid(B, 28, 28);
          return true;
        }
      }
      return false;
    };
    const int i_hi = part == 2 ? split - 1 : nres - 1, i_lo = part == 1 ? split : 0;
    if (part != 2) {
    // --- adversarial path through the UPDATED discriminator (trainer.py:116-117)
    // --- classifier path (trainer.py:118): independent of the discriminator path until the two input gradients meet
    //     in residual_head_bwd
    //     ... has already run beside the discriminator step (c_branch in step_d_grads): dxc and G_CLS are ready
    d_input<T>(x_cf, d_embed, in.target, B, 784, a0, s);
    d_fwd(B, s);
    bce_logits(dlogits_d, B, 1, 1.f, 1.f, cfg.lambda_adv, cfg.lambda_adv, scal + PCG_S_G_ADV, scal_tmp, ddlogit, s);
    d_bwd(B, cfg.pollute_d_grads != 0, 0, s);
    // --- through clamp / mask / scaling (trainer.py:97,99,119; generator.py:80-82)
    residual_head_bwd<T>(dxd, 1, dxc, raw, in.x, in.mask, cfg.residual_scaling, cfg.lambda_reg, cfg.lambda_mask, MG,
                         g_c, s);
    // --- generator backward
    // weight gradients go to side_w (see d_bwd); every dY they read has its own buffer
    after(s, side_w);
    g_loss_combine(scal + PCG_S_G_ADV, scal + PCG_S_G_CLS, scal + PCG_S_REG_L1, scal + PCG_S_MASK_PEN, cfg.lambda_adv,
                   cfg.lambda_cls, cfg.lambda_reg, cfg.lambda_mask, scal + PCG_S_G_LOSS, side_w);   // off the gradient chain
    wgrad<T, T>(g_out, hm, g_c, side_w);
    if (!g_out.to1_wgrad) bias_grad(g_c, MG, 1, g_out.db, s);
    {
      GenEpilogue<T> e; e.act_ref = hm; e.ref_act = ACT_LRELU; e.ref_slope = 0.2f;
      dgrad<T, T>(g_out, g_c, e, g_hm, s);
    }
    // conv_mid's bias gradient (a full read of g_hm) is needed by the optimizer only; side_mid_db runs it on the
    // weight-gradient stream, with its own partial rows (experiment switch)
    if (!side_mid_db) bias_grad(g_hm, MG, ch, g_mid.db, s);
    dh = dhA;
    dh_other = dhB;
    // The reduction pass of BN2's backward (sum g, sum g*xhat over g = 0.1*dh, generator.py:22) rides in the epilogue of
    // the tensor-core kernel that PRODUCES dh - conv_mid's data gradient for the last block, conv1's data gradient of
    // block i+1 (which also adds the skip gradient) for block i - instead of re-reading dh and y2 in a separate pass.
    bn2_parts = 0;
    if (!dgrad_bn2(g_mid, g_hm, nullptr, nres - 1, dh)) {
      GenEpilogue<T> e;
      dgrad<T, T>(g_mid, g_hm, e, dh, s);
    }
    after(s, side_w);
    if (side_mid_db) bias_grad(g_hm, MG, ch, g_mid.db, side_w, stat_part3);
    wgrad<T, T>(g_mid, h[nres], g_hm, side_w);
    }   // part != 2
    for (int i = i_hi; i >= i_lo; --i) {
      // BN2 backward: upstream = 0.1 * dh (generator.py:22)
      const BN& q2 = bn2[i];
      if (bn2_parts > 0) {
        // c12 still holds this block's BN2 sums: the fused launch finished them, and nothing ran in between
        if (!fuse_finalize) bn_bwd_finalize(stat_bn2, bn2_parts, MG, ch, q2.dgamma, q2.dbeta, c12, s);
        bn2_parts = 0;
      } else {
        bn_bwd_partial<T>(dh, y2[i], q2.mean, q2.rstd, q2.scale, q2.shift, 0.1f, ACT_NONE, 0.f, MG, ch, stat_part, s);
        bn_bwd_finalize(stat_part, STAT_PARTS, MG, ch, q2.dgamma, q2.dbeta, c12, s);
      }
      // conv-bias gradients: nothing reads them before the optimizer; with batch_db every layer keeps its own partial
      // rows and ONE launch sums them all after the loop (experiment switch)
      bn_bwd_apply<T>(dh, y2[i], q2.mean, q2.rstd, q2.scale, q2.shift, q2.gamma, c12, 0.1f, ACT_NONE, 0.f, MG, ch, dy2[i],
                      batch_db ? stat_db + (size_t)(2 * i) * STAT_PARTS * ch : stat_part2, s);
      if (!batch_db) colsum_finalize(stat_part2, STAT_PARTS, ch, ch, g_c2[i].db, s);
      // data gradient of conv2; on the tensor-core path its epilogue also does the reduction pass of BN1's backward
      // (sum g, sum g*xhat with g = dz1 * LReLU'(BN1(y1))), which saves one full read of dz1 and y1
      int bn1_parts = 0;
      bool bn1_fin_done = false;
      if constexpr (kBf16) {
        if (g_c2[i].tc_dgrad && g_c2[i].tc64) {
          ProfTag _tag("g.res.dgrad_bnred");
          const BN& qb = bn1[i];
          ConvEpilogue c;
          c.stats = stat_part;
          c.bn_y = y1[i]; c.bn_mean = qb.mean; c.bn_rstd = qb.rstd; c.bn_scale = qb.scale; c.bn_shift = qb.shift;
          c.bn_act = ACT_LRELU; c.bn_slope = 0.2f;
          if (fuse_finalize) {
            c.fin.mode = 2; c.fin.counter = fin_counter; c.fin.M = MG;
            c.fin.dgamma = qb.dgamma; c.fin.dbeta = qb.dbeta; c.fin.c12 = c12;
            bn1_fin_done = true;
          }
          conv_tc64_fprop(dy2[i], B, 28, 28, g_c2[i].tcd, c, dz1, s);
          bn1_parts = conv_tc64_fprop_grid(B, 28, 28);
        }
      }
      if (bn1_parts == 0) {
        GenEpilogue<T> e;
        dgrad<T, T>(g_c2[i], dy2[i], e, dz1, s);
      }
      // forked AFTER the data gradient: the two tcgen05 kernels cannot share an SM (shared memory), and this order
      // makes the weight gradient overlap the HBM-bound BatchNorm kernels that follow on the main stream
      after(s, side_w);
      wgrad<T, T>(g_c2[i], z1[i], dy2[i], side_w);
      // LeakyReLU + BN1 backward
      const BN& q1 = bn1[i];
      if (bn1_parts == 0) {
        bn_bwd_partial<T>(dz1, y1[i], q1.mean, q1.rstd, q1.scale, q1.shift, 1.f, ACT_LRELU, 0.2f, MG, ch, stat_part, s);
        bn1_parts = STAT_PARTS;
      }
      if (!bn1_fin_done) bn_bwd_finalize(stat_part, bn1_parts, MG, ch, q1.dgamma, q1.dbeta, c12, s);
      bn_bwd_apply<T>(dz1, y1[i], q1.mean, q1.rstd, q1.scale, q1.shift, q1.gamma, c12, 1.f, ACT_LRELU, 0.2f, MG, ch,
                      dy1[i], batch_db ? stat_db + (size_t)(2 * i + 1) * STAT_PARTS * ch : stat_part2, s);
      if (!batch_db) colsum_finalize(stat_part2, STAT_PARTS, ch, ch, g_c1[i].db, s);
      if (i == 0 || !dgrad_bn2(g_c1[i], dy1[i], dh, i - 1, dh_other)) {
        GenEpilogue<T> e; e.add_src = dh;
        if (i == 0) { e.act_ref = h[0]; e.ref_act = ACT_LRELU; e.ref_slope = 0.2f; }
        dgrad<T, T>(g_c1[i], dy1[i], e, dh_other, s);
      }
      after(s, side_w);
      wgrad<T, T>(g_c1[i], h[i], dy1[i], side_w);
      T* t = dh; dh = dh_other; dh_other = t;
    }
    if (nres == 0) {
      // no residual block consumed h0's activation derivative: apply it here via a copy-free trick
      GenEpilogue<T> e; (void)e;
      throw Error(1, "n_resblocks == 0 is not supported by the fused backward");
    }
    if (batch_db && i_hi >= i_lo) {
      ColsumOuts outs;
      for (int i = i_lo; i <= i_hi; ++i) {
        outs.p[2 * (i - i_lo)] = g_c2[i].db;
        outs.p[2 * (i - i_lo) + 1] = g_c1[i].db;
      }
      colsum_finalize_multi(stat_db + (size_t)(2 * i_lo) * STAT_PARTS * ch, STAT_PARTS, ch, outs, 2 * (i_hi - i_lo + 1), s);
    }
    if (part == 1) {                               // the caller reduces the finished half of the gradients now
      after(side_w, s);
      return;
    }
    // dh now holds d loss / d (pre-activation of conv_in)
    after(s, side_w);
    wgrad<T, T>(g_in, inp3, dh, side_w);
    if (!g_in.few_wgrad) bias_grad(dh, MG, ch, g_in.db, s);
    {
      GenEpilogue<float> e;
      dgrad<T, float>(g_in, dh, e, dinp, s, 0, /*ch_select=*/1);    // only the label-embedding channel
    }
    after(side_w, s);
    embed_grad<float>(dinp, 1, 0, in.target, B, 784, 10, g_dembed, s);
  }

  void step_g_update(cudaStream_t s) override {
    long long gt;
    net_layout(0, ch, nres, &gt);
    adam_flat(buf.g_params, buf.g_grads, buf.g_adam_m, buf.g_adam_v, gt, buf.g_step, cfg.g_lr, cfg.beta1, cfg.beta2,
              cfg.adam_eps, cfg.grad_scale, s);
    refresh_g(s);
  }

  // ---------------------------------------------------------------- module forwards
  void g_forward(const float* x, const long long* target, const float* mask, int training, float* raw_o, float* masked_o,
                 cudaStream_t s) override {
    if (training == 0 && fold_n > 0) g_fwd_eval(x, target, mask, s);
    else g_fwd(x, target, mask, training != 0, s);
    PCG_CHECK_CUDA(cudaMemcpyAsync(raw_o, raw, MG * sizeof(float), cudaMemcpyDeviceToDevice, s));
    PCG_CHECK_CUDA(cudaMemcpyAsync(masked_o, masked, MG * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  void d_forward(const float* x, const long long* cond, float* logits, cudaStream_t s) override {
    d_input<T>(x, d_embed, cond, B, 784, a0, s);
    d_fwd(B, s);
    PCG_CHECK_CUDA(cudaMemcpyAsync(logits, dlogits_d, B * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  void c_forward(const float* x, float* logits, cudaStream_t s) override {
    c_fwd(x, s);
    PCG_CHECK_CUDA(cudaMemcpyAsync(logits, clogits, (size_t)B * 10 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  void debug_tensor(const std::string& name, void** ptr, long long* numel, int* dtype) override {
    auto it = dbg.find(name);
    if (it == dbg.end()) throw Error(1, "unknown debug tensor " + name);
    *ptr = it->second.first; *numel = it->second.second.first; *dtype = it->second.second.second;
  }
};

__global__ void bn_eval_coeffs_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                      int C, float* mean, float* rstd, float* scale, float* shift) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float r = 1.0f / sqrtf(rv[c] + eps);
  mean[c] = rm[c]; rstd[c] = r;
  scale[c] = gamma[c] * r;
  shift[c] = beta[c] - rm[c] * gamma[c] * r;
}
template <typename T>
void MnistPlan<T>::bn_eval_coeffs(const BN& q, cudaStream_t s) {
  launch_k(bn_eval_coeffs_kernel, dim3(cdiv(ch, 64)), dim3(64), 0, s, q.gamma, q.beta, q.running_mean, q.running_var, 1e-5f, ch, q.mean,
                                                   q.rstd, q.scale, q.shift);
  PCG_COUNT_LAUNCH();
  PCG_LAUNCH_CHECK();
}

}  // namespace pcg

using namespace pcg;

struct pcg_mnist_plan {
  PlanBase* impl;
};

#define PCG_API_BEGIN try {
#define PCG_API_END                                   \
  return 0;                                           \
  }                                                   \
  catch (const pcg::Error& e) {                       \
    pcg::set_last_error(e.what());                    \
    return e.code;                                    \
  }                                                   \
  catch (const std::exception& e) {                   \
    pcg::set_last_error(e.what());                    \
    return 99;                                        \
  }

extern "C" {

int pcg_mnist_layout(int net, int base_ch, int n_resblocks, int idx, long long* offset, long long* numel,
                     long long* total) {
  try {
    auto l = net_layout(net, base_ch, n_resblocks, total);
    if (idx >= 0) {
      if (idx >= (int)l.size()) throw Error(1, "tensor index out of range");
      if (offset) *offset = l[idx].offset;
      if (numel) *numel = l[idx].numel;
    }
    return (int)l.size();
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return -1;
  }
}

int pcg_mnist_plan_create(const pcg_mnist_config* cfg, const pcg_mnist_buffers* buf, pcg_mnist_plan** out) {
  PCG_API_BEGIN
  PCG_REQUIRE(cfg && buf && out, "null argument");
  PCG_REQUIRE(buf->g_params && buf->g_grads && buf->g_adam_m && buf->g_adam_v && buf->g_step && buf->g_bn_running &&
                  buf->g_bn_nbt && buf->d_params && buf->d_grads && buf->d_adam_m && buf->d_adam_v && buf->d_step &&
                  buf->c_params,
              "all arenas must be bound");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw Error(4, "no CUDA device: libpcg has no CPU fallback");
  pcg_mnist_plan* p = new pcg_mnist_plan;
  p->impl = nullptr;
  try {
    if (cfg->precision == PCG_BF16) p->impl = new MnistPlan<bf16>(*cfg, *buf);
    else if (cfg->precision == PCG_F32) p->impl = new MnistPlan<float>(*cfg, *buf);
    else throw Error(1, "precision must be PCG_F32 or PCG_BF16");
    p->impl->refresh_weights(0);
    PCG_CHECK_CUDA(cudaStreamSynchronize(0));
  } catch (...) {
    delete p->impl;
    delete p;
    throw;
  }
  *out = p;
  PCG_API_END
}

int pcg_mnist_plan_destroy(pcg_mnist_plan* plan) {
  PCG_API_BEGIN
  if (plan) {
    delete plan->impl;
    delete plan;
  }
  PCG_API_END
}

int pcg_mnist_refresh_weights(pcg_mnist_plan* plan, void* stream) {
  PCG_API_BEGIN
  plan->impl->refresh_weights((cudaStream_t)stream);
  PCG_API_END
}

int pcg_mnist_step_d_grads(pcg_mnist_plan* plan, const pcg_mnist_inputs* in, float* scalars, void* stream) {
  PCG_API_BEGIN
  plan->impl->step_d_grads(*in, scalars, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_mnist_step_d_update(pcg_mnist_plan* plan, void* stream) {
  PCG_API_BEGIN
  plan->impl->step_d_update((cudaStream_t)stream);
  PCG_API_END
}
int pcg_mnist_step_g_grads(pcg_mnist_plan* plan, const pcg_mnist_inputs* in, float* scalars, void* stream) {
  PCG_API_BEGIN
  plan->impl->step_g_grads(*in, scalars, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_mnist_set_defer_c_bwd(pcg_mnist_plan* plan, int on) {
  PCG_API_BEGIN
  PCG_REQUIRE(plan, "null plan");
  plan->impl->set_defer_c_bwd(on != 0);
  PCG_API_END
}
int pcg_mnist_step_c_bwd(pcg_mnist_plan* plan, void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(plan, "null plan");
  plan->impl->step_c_bwd((cudaStream_t)stream);
  PCG_API_END
}
int pcg_mnist_step_g_grads_part(pcg_mnist_plan* plan, const pcg_mnist_inputs* in, float* scalars, int part, int split_block,
                                void* stream) {
  PCG_API_BEGIN
  PCG_REQUIRE(plan && in && scalars, "null argument");
  plan->impl->step_g_grads_part(*in, scalars, (cudaStream_t)stream, part, split_block);
  PCG_API_END
}
int pcg_mnist_step_g_update(pcg_mnist_plan* plan, void* stream) {
  PCG_API_BEGIN
  plan->impl->step_g_update((cudaStream_t)stream);
  PCG_API_END
}
int pcg_mnist_step(pcg_mnist_plan* plan, const pcg_mnist_inputs* in, float* scalars, void* stream) {
  PCG_API_BEGIN
  cudaStream_t s = (cudaStream_t)stream;
  plan->impl->step_d_grads(*in, scalars, s);
  plan->impl->step_d_update(s);
  plan->impl->step_g_grads(*in, scalars, s);
  plan->impl->step_g_update(s);
  PCG_API_END
}

int pcg_mnist_g_forward(pcg_mnist_plan* plan, const float* x, const long long* target, const float* mask, int training,
                        float* raw, float* masked, void* stream) {
  PCG_API_BEGIN
  plan->impl->g_forward(x, target, mask, training, raw, masked, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_mnist_d_forward(pcg_mnist_plan* plan, const float* x, const long long* cond, float* logits, void* stream) {
  PCG_API_BEGIN
  plan->impl->d_forward(x, cond, logits, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_mnist_c_forward(pcg_mnist_plan* plan, const float* x, float* logits, void* stream) {
  PCG_API_BEGIN
  plan->impl->c_forward(x, logits, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_mnist_debug_tensor(pcg_mnist_plan* plan, const char* name, void** ptr, long long* numel, int* dtype) {
  PCG_API_BEGIN
  plan->impl->debug_tensor(name, ptr, numel, dtype);
  PCG_API_END
}

}  // extern "C"

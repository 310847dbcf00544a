// extern "C" surface of libpcg (declared in include/pcg.h).  Converts C++ exceptions to codes.
#include "../../include/pcg.h"

#include "common.cuh"
#include "conv_small.cuh"
#include "conv_tc.cuh"

#include <string.h>

#include <map>
#include <vector>

namespace pcg {
static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }
unsigned long long g_launch_count = 0;

bool g_profile_on = false;
static int env_pdl() {
  const char* e = getenv("PCG_PDL");
  return e ? (atoi(e) != 0) : 0;      // off by default: measured slower as a blanket policy (profiles/exp_pdl_r1.md)
}
int g_pdl = env_pdl();
static int env_pdl_edges() {
  const char* e = getenv("PCG_PDL_EDGES");
  return e ? (atoi(e) != 0) : 1;      // on by default: only the tcgen05 64->64 kernels, whose prologue runs before their wait
}
int g_pdl_edges = env_pdl_edges();
static int env_pdl_small() {
  const char* e = getenv("PCG_PDL_SMALL");
  return e ? atoi(e) : 0;
}
int g_pdl_small = env_pdl_small();
static int env_l2_hints() {
  const char* e = getenv("PCG_L2_HINTS");
  // all five on: MNIST step 2.84-2.87 -> 2.81 ms with bits 1 | 2 | 4, 2.73 -> 2.70 ms with 8 | 16 added, interleaved on
  // one box each (profiles/exp_l2_hints_r2i.txt)
  return e ? atoi(e) : 31;
}
int g_l2_hints = env_l2_hints();
const char* g_prof_tag = nullptr;
struct ProfRec { std::string name; cudaEvent_t e0, e1; };
static std::vector<ProfRec> g_prof;
ProfileScope::ProfileScope(const char* n, cudaStream_t s) : name(n), stream(s) {
  if (!g_profile_on) return;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
}
ProfileScope::~ProfileScope() {
  if (e0 == nullptr) return;
  cudaEventRecord(e1, stream);
  g_prof.push_back({g_prof_tag ? std::string(name) + ":" + g_prof_tag : std::string(name), e0, e1});
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    PCG_CHECK_CUDA(cudaGetDevice(&dev));
    PCG_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}
}  // namespace pcg

using namespace pcg;

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

const char* pcg_last_error(void) { return t_last_error.c_str(); }
int pcg_version(void) { return PCG_VERSION; }
unsigned long long pcg_launch_count(void) { return g_launch_count; }
int pcg_set_pdl(int on) { const int prev = g_pdl; g_pdl = on ? 1 : 0; return prev; }

int pcg_profile_begin(void) {
  PCG_API_BEGIN
  PCG_CHECK_CUDA(cudaDeviceSynchronize());
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_profile_on = true;
  PCG_API_END
}

int pcg_profile_end(char* out, size_t cap) {
  PCG_API_BEGIN
  g_profile_on = false;
  PCG_CHECK_CUDA(cudaDeviceSynchronize());
  std::map<std::string, std::pair<double, int>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      auto& a = agg[r.name];
      a.first += ms;
      a.second += 1;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  std::string js = "{";
  bool first = true;
  for (auto& kv : agg) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s\"%s\": {\"ms\": %.6f, \"launches\": %d}", first ? "" : ", ", kv.first.c_str(),
             kv.second.first, kv.second.second);
    js += buf;
    first = false;
  }
  js += "}";
  PCG_REQUIRE(out != nullptr && js.size() + 1 <= cap, "profile buffer too small");
  memcpy(out, js.c_str(), js.size() + 1);
  PCG_API_END
}

int pcg_memcpy_d2d(void* dst, const void* src, size_t nbytes, void* stream) {
  PCG_API_BEGIN
  PCG_CHECK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  PCG_API_END
}

int pcg_conv_tc_grid(long long M, int Cout) {
  try { return conv_tc_grid(M, Cout); } catch (const std::exception& e) { set_last_error(e.what()); return -1; }
}
int pcg_conv_tc_wgrad_grid(long long M) {
  try { return conv_tc_wgrad_grid(M); } catch (const std::exception& e) { set_last_error(e.what()); return -1; }
}

int pcg_conv_tc_fprop(const void* in, int N, int H, int W, int Cin, const void* wpk, int Cout, int ksize,
                      int stride, int pad, const float* bias, int act, float slope, const void* add_src,
                      void* out, float* stats, void* stream) {
  PCG_API_BEGIN
  ConvEpilogue e;
  e.bias = bias; e.act = act; e.slope = slope; e.add_src = (const bf16*)add_src; e.stats = stats;
  conv_tc_fprop((const bf16*)in, N, H, W, Cin, (const bf16*)wpk, Cout, ksize, stride, pad, e, (bf16*)out,
                (cudaStream_t)stream);
  PCG_API_END
}

int pcg_conv_tc_wgrad64(const void* x, const void* dy, int N, int H, int W, float* part, float* dw,
                        void* stream) {
  PCG_API_BEGIN
  conv_tc_wgrad64((const bf16*)x, (const bf16*)dy, N, H, W, part, (cudaStream_t)stream);
  wgrad_reduce_tc(part, conv_tc_wgrad_grid((long long)N * H * W), dw, (cudaStream_t)stream);
  PCG_API_END
}

int pcg_conv_tc64_grid(int N, int H, int W) {
  try { return conv_tc64_grid(N, H, W); } catch (const std::exception& e) { set_last_error(e.what()); return -1; }
}
int pcg_conv_tc64_fprop_grid(int N, int H, int W) {
  try { return conv_tc64_fprop_grid(N, H, W); } catch (const std::exception& e) { set_last_error(e.what()); return -1; }
}
int pcg_conv_tc64_set_variant(int v) { conv_tc64_set_variant(v); return 0; }

// ---- skinny-layer kernels (conv_small.cu): test entry points
int pcg_conv_to1(const void* in, int N, int H, int W, int Cin, const void* w9, const float* bias, float* out, void* stream) {
  PCG_API_BEGIN
  conv_to1((const bf16*)in, N, H, W, Cin, (const bf16*)w9, bias, out, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_conv_few(const void* in, int in_is_f32, int N, int H, int W, int Cs, const void* wnk, int Cout, int stride,
                 const float* bias, int act, float slope, const void* act_ref, int ref_act, float ref_slope, void* out,
                 void* stream) {
  PCG_API_BEGIN
  FewEpilogue e;
  e.bias = bias; e.act = act; e.slope = slope; e.act_ref = (const bf16*)act_ref; e.ref_act = ref_act; e.ref_slope = ref_slope;
  if (in_is_f32) conv_few<float>((const float*)in, N, H, W, Cs, (const bf16*)wnk, Cout, stride, e, (bf16*)out, (cudaStream_t)stream);
  else conv_few<bf16>((const bf16*)in, N, H, W, Cs, (const bf16*)wnk, Cout, stride, e, (bf16*)out, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_dgrad_s2_to1(const void* dy, int N, int H, int W, const void* wrot, float* dx, void* stream) {
  PCG_API_BEGIN
  dgrad_s2_to1((const bf16*)dy, N, H, W, (const bf16*)wrot, dx, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_wgrad_small_parts(void) {
  try { return wgrad_few_parts(); } catch (const std::exception& e) { set_last_error(e.what()); return -1; }
}
int pcg_wgrad_few(const void* in, const void* dy, int N, int H, int W, int Cs, int stride, float* part, float* dw, float* db,
                  void* stream) {
  PCG_API_BEGIN
  wgrad_few<bf16>((const bf16*)in, (const bf16*)dy, N, H, W, Cs, stride, part, dw, db, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_wgrad_to1(const void* x, const void* g, int N, int H, int W, float* part, float* dw, float* db, void* stream) {
  PCG_API_BEGIN
  wgrad_to1((const bf16*)x, (const bf16*)g, N, H, W, part, dw, db, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_conv_tc64_fprop(const void* in, int N, int H, int W, const void* wpk, const float* bias, int act, float slope,
                        const void* add_src, const void* act_ref, int ref_act, void* out, float* stats, void* stream) {
  PCG_API_BEGIN
  ConvEpilogue e;
  e.bias = bias; e.act = act; e.slope = slope; e.add_src = (const bf16*)add_src; e.stats = stats;
  e.act_ref = (const bf16*)act_ref; e.ref_act = ref_act; e.ref_slope = slope;
  conv_tc64_fprop((const bf16*)in, N, H, W, (const bf16*)wpk, e, (bf16*)out, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_conv_tc64_dgrad_bnred(const void* in, int N, int H, int W, const void* wpk, const void* add_src, const void* bn_y,
                              const float* bn_mean, const float* bn_rstd, const float* bn_scale, const float* bn_shift,
                              int bn_act, float bn_slope, float bn_gscale, void* out, float* stats, void* stream) {
  PCG_API_BEGIN
  ConvEpilogue e;
  e.add_src = (const bf16*)add_src; e.stats = stats;
  e.bn_y = (const bf16*)bn_y; e.bn_mean = bn_mean; e.bn_rstd = bn_rstd; e.bn_scale = bn_scale; e.bn_shift = bn_shift;
  e.bn_act = bn_act; e.bn_slope = bn_slope; e.bn_gscale = bn_gscale;
  conv_tc64_fprop((const bf16*)in, N, H, W, (const bf16*)wpk, e, (bf16*)out, (cudaStream_t)stream);
  PCG_API_END
}
int pcg_conv_tc64_wgrad(const void* x, const void* dy, int N, int H, int W, float* part, float* dw, void* stream) {
  PCG_API_BEGIN
  conv_tc64_wgrad((const bf16*)x, (const bf16*)dy, N, H, W, part, (cudaStream_t)stream);
  wgrad_reduce_tc(part, conv_tc64_grid(N, H, W), dw, (cudaStream_t)stream, conv_tc64_wgrad_swizzled());
  PCG_API_END
}

int pcg_pack_conv_weights_tc(const float* w, int Cout, int Cin, int ksize, void* fprop, void* dgrad,
                             void* stream) {
  PCG_API_BEGIN
  pack_conv_weights_tc(w, Cout, Cin, ksize, (bf16*)fprop, (bf16*)dgrad, (cudaStream_t)stream);
  PCG_API_END
}

int pcg_debug_im2col_tile(const void* in, int N, int H, int W, int Cin, int ksize, int stride, int pad,
                          int first_pixel, int tap_r, int tap_s, int cblock, void* out128x64, void* stream) {
  PCG_API_BEGIN
  debug_im2col_tile((const bf16*)in, N, H, W, Cin, ksize, stride, pad, first_pixel, tap_r, tap_s, cblock,
                    (bf16*)out128x64, (cudaStream_t)stream);
  PCG_API_END
}

}  // extern "C"

// extern "C" surface of libpcg (declared in include/pcg.h).  Converts C++ exceptions to codes.
#include "../../include/pcg.h"

#include "common.cuh"
#include "conv_tc.cuh"

namespace pcg {
static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }
unsigned long long g_launch_count = 0;

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

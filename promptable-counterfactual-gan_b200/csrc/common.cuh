// Shared host/device helpers for libpcg.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>

namespace pcg {

// ---- error plumbing: internal code throws, the extern "C" layer converts to codes --------------
void set_last_error(const std::string& msg);

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PCG_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      throw ::pcg::Error(2, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +       \
                                __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
    }                                                                                         \
  } while (0)

#define PCG_REQUIRE(cond, msg)                                                                \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      throw ::pcg::Error(1, std::string("invalid argument: ") + (msg) + " [" #cond "] (" +    \
                                __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
    }                                                                                         \
  } while (0)

#define PCG_LAUNCH_CHECK() PCG_CHECK_CUDA(cudaGetLastError())

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// Steps are chains of 100-1000 short dependent kernels; between two of them the GPU idles for the launch latency, the
// CTA ramp and the successor's prologue.  Every kernel of this library therefore starts with pdl_enter(): it lets the
// NEXT kernel of the stream be launched right away (griddepcontrol.launch_dependents) and then waits until the PREVIOUS
// kernel has completed and flushed (griddepcontrol.wait) before touching global memory.  launch_k() launches with the
// programmatic-stream-serialization attribute when PDL is on (env PCG_PDL=1 or pcg_set_pdl(1)); otherwise both
// instructions are no-ops.  Works in eager streams and in captured graphs (programmatic kernel-node edges; all 70 GPU
// parity tests pass in that mode).  OFF by default: with the wait at the top of every kernel there is no prologue to
// overlap and a programmatic graph edge costs ~0.4 us more than a plain one (profiles/exp_pdl_r1.md: MNIST step 3.22
// -> 3.32 ms, KC 2.18 -> 2.62 ms).  It pays only for kernels that do real work before pdl_wait().
// Rule for kernel authors: no global-memory access before pdl_enter() / pdl_wait().
extern int g_pdl;
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_launch_dependents();
  pdl_wait();
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  PCG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// Per-edge variant for kernels with a real prologue (the tcgen05 convolutions: barrier init, TMEM allocation, shared-memory
// zero fill - ~2 us that touch no global memory): they call pdl_launch_dependents() first, run the prologue, and only then
// pdl_wait(); launched with launch_k_pdl the prologue (and the launch latency) overlaps the predecessor's tail.  Switch:
// g_pdl_edges (env PCG_PDL_EDGES, default on).
extern int g_pdl_edges;
template <typename... KArgs, typename... Args>
inline void launch_k_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_pdl || g_pdl_edges) ? 1 : 0;
  PCG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// Experiment switch PCG_PDL_SMALL (default 0): 1 = the statistics-finalize kernels, 2 = also the BatchNorm apply kernels
// are launched with the programmatic attribute, so that their launch latency overlaps the producer they wait for.
extern int g_pdl_small;
template <typename... KArgs, typename... Args>
inline void launch_k_small(int level, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                           Args&&... args) {
  if (g_pdl_small >= level) launch_k_pdl(kernel, grid, block, smem, stream, std::forward<Args>(args)...);
  else launch_k(kernel, grid, block, smem, stream, std::forward<Args>(args)...);
}

// L2 eviction hints (PCG_L2_HINTS bit mask, experiment): 1 = the 64->64 weight gradient streams both operands with
// evict_first (their last use), 2 = bn_bwd_apply reads with streaming loads (last use of both inputs), 4 = the skip
// operand of a fused data gradient with evict_first, 8 = the forward BatchNorm apply kernels read with streaming loads,
// 16 = the 64->64 forward convolution (with statistics) loads its input with evict_first.
extern int g_l2_hints;

// Number of SMs of the current device (cached).
int sm_count();

// Counts kernels launched by this library (bench.py's "gpu_launches").
extern unsigned long long g_launch_count;
#define PCG_COUNT_LAUNCH() (++::pcg::g_launch_count)

// ---- optional per-launcher device timing (bench.py roofline / kernel-share breakdown) ----------
// When enabled (pcg_profile_begin), every launcher brackets its kernels with CUDA events on the
// launching stream; pcg_profile_end() synchronises and reports total ms + launch count per name.
// Must be off during CUDA-graph capture.
extern bool g_profile_on;
extern const char* g_prof_tag;     // optional finer label set by the plan ("d.conv2.dgrad", ...)
struct ProfTag {
  const char* prev;
  explicit ProfTag(const char* t) : prev(g_prof_tag) { g_prof_tag = t; }
  ~ProfTag() { g_prof_tag = prev; }
};
struct ProfileScope {
  const char* name;
  cudaStream_t stream;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfileScope(const char* n, cudaStream_t s);
  ~ProfileScope();
};
#define PCG_PROFILE(name, stream) ::pcg::ProfileScope _pcg_prof_scope(name, stream)

// ---- storage-type conversion ---------------------------------------------------------------
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// Activation codes shared by every kernel.
enum Act { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2 };

}  // namespace pcg

// Scalar reductions over more rows than one block should walk: the loss / regulariser kernels (cross-entropy, GAN losses,
// row norms, |x| means) write a per-element gradient that the rest of the iteration waits for and ONE scalar.  As single
// blocks they take 11-45 us at 4096 rows (16 rows per thread, every load latency in sequence).  Launched as a thread-block
// cluster of 8 CTAs each CTA takes a row slice and rank 0 adds the eight block totals through distributed shared memory
// in rank order: deterministic, no scratch, no second launch.  With one CTA (small inputs) the same code is the old kernel.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace pcg {

constexpr int CR_CTAS = 8;                 // portable cluster size
constexpr long long CR_MIN_ITEMS = 1024;   // below this one CTA is faster than a cluster launch

// [begin, end) of this CTA's slice of n items (whole slices of the cluster: the last ones may be empty)
__device__ __forceinline__ void cluster_slice(long long n, long long& begin, long long& end) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const long long per = (n + cluster.num_blocks() - 1) / cluster.num_blocks();
  begin = (long long)cluster.block_rank() * per;
  end = begin + per < n ? begin + per : n;
  if (begin > n) begin = n;
}

// block_total: this CTA's sum, valid in thread 0.  Returns the cluster's sum (rank order), valid in thread 0 of rank 0.
// slot: one __shared__ float of the caller.  Every thread of every CTA must call it.
__device__ __forceinline__ float cluster_total(float block_total, float* slot) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  if (threadIdx.x == 0) *slot = block_total;
  cluster.sync();
  float t = 0.f;
  if (cluster.block_rank() == 0 && threadIdx.x == 0)
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) t += *cluster.map_shared_rank(slot, r);
  cluster.sync();                          // remote shared memory stays alive until rank 0 has read it
  return t;
}

__device__ __forceinline__ bool cluster_leader() {
  return cooperative_groups::this_cluster().block_rank() == 0 && threadIdx.x == 0;
}

// launch as one CTA, or as a cluster of CR_CTAS when there are enough items
template <typename... KArgs, typename... Args>
inline void launch_k_cluster(void (*kernel)(KArgs...), long long items, dim3 block, size_t smem, cudaStream_t stream,
                             Args&&... args) {
  const unsigned ctas = items >= CR_MIN_ITEMS ? CR_CTAS : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 2 : 1;
  PCG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

}  // namespace pcg

// One whole iteration of the two-layer MLP GANs on 2-D points in ONE kernel launch.
//
// Replaces the loop bodies of conditional_gan/moons/make_moons_cgan.py:90-129 and simple_gan/moons/make_moons_gan.py:61-88
// (generator forward, discriminator forward on real + fake, log losses, both backward passes, both Adam updates;
// Generator = Linear(z+label -> 128)-ReLU-Linear(128 -> 2), Discriminator = Linear(2+label -> 128)-ReLU-Linear(128 -> 1)-Sigmoid).
//
// The nets are tiny (5.5 k parameters, 47 kFLOP per sample): as ~100 dependent graph nodes the iteration is pure launch
// latency (245 us at batch 1024).  Here one thread-block cluster of up to 16 CTAs runs the whole iteration:
//   * every CTA keeps ALL weights of both nets in shared memory and owns B / cluster_size samples;
//   * "phase A" (TPS = 4..32 threads per sample, 128 / TPS hidden units each, warp-shuffle reductions over them; small
//     batches get more threads per sample) runs the forwards, the
//     losses and the gradients with respect to the layer outputs; "phase B" (one thread per hidden unit and sample
//     group) accumulates the weight gradients over the CTA's samples in registers;
//   * the per-CTA gradient partials are summed over the cluster through distributed shared memory in rank order
//     (deterministic), each CTA owning a slice of the parameters: it applies Adam to the slice, writes parameters /
//     gradients / moments to global memory and, for the discriminator, broadcasts the new weights into every CTA's
//     shared memory, because the generator step sees the updated discriminator (make_moons_cgan.py:111 -> :121).
// Four cluster barriers per iteration; no global-memory round trip between the layers.
#include <cooperative_groups.h>

#include "common.cuh"
#include "ops.cuh"

namespace cg = cooperative_groups;

namespace pcg {

constexpr int MG_H = 128;          // hidden width
constexpr int MG_GS = 36;          // padded generator input width (z + label <= 36), row stride of the layer-1 weights
constexpr int MG_DS = 4;           // padded discriminator input width (2 + label <= 4)
constexpr int MG_NB = 128;         // samples per CTA (4 threads each in phase A)
constexpr int MG_T = 512;
constexpr int MG_HS = MG_H + 4;    // row stride of the stored generator hidden activations (bank spread)
constexpr int MG_PG_MAX = MG_H * MG_GS + MG_H + 2 * MG_H + 4;
constexpr int MG_PD_MAX = MG_H * MG_DS + MG_H + MG_H + 4;

struct MlpGanArgs {
  int B, zd, ld, nb;
  const float *real, *real_oh, *z1, *oh1, *z2, *oh2;
  float *gp, *gg, *gm, *gv;
  int* gstep;
  float *dp, *dg, *dm, *dv;
  int* dstep;
  float lr, beta1, beta2, eps;
  float* scal;
};

struct MlpGanSmem {
  float gw1[MG_H * MG_GS], gb1[MG_H], gw2[2 * MG_H], gb2[4];
  float dw1[MG_H * MG_DS], db1[MG_H], dw2[MG_H], db2[4];
  float x[MG_NB * MG_GS];            // generator inputs of the generator step (phase B operand)
  float h[MG_NB * MG_HS];            // generator hidden activations of the generator step
  float din[2 * MG_NB * MG_DS];      // discriminator inputs of the discriminator step: real rows, fake rows
  float dz[2 * MG_NB];               // d loss / d logit of those rows
  float df[MG_NB * 2];               // d loss_G / d fake
  float part_g[MG_PG_MAX], part_d[MG_PD_MAX];   // this CTA's gradient partials, flat parameter layout
  float loss[8];                     // partial sums: -log D(real), -log(1 - D(fake)), D(real), D(fake), -log D(fake2)
  float adam[4];                     // step_size / sqrt(bias_correction2) of D, of G
  float red[32];
};

template <int TPS>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = 1; o < TPS; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the CTA in a fixed order (warp shuffles, then the 16 warp partials in order); result valid in thread 0
__device__ __forceinline__ float cta_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < MG_T / 32; ++w) s += red[w];
  return s;
}

// Generator forward for one sample, this thread's 128 / TPS hidden units (j = jj*TPS + jq); the two outputs are summed
// over the TPS threads of the sample.
template <int TPS, bool STORE>
__device__ __forceinline__ void g_forward(const float (&x)[MG_GS], const MlpGanSmem& S, int jq, float* hrow, float& o0,
                                          float& o1) {
  float p0 = 0.f, p1 = 0.f;
#pragma unroll 2
  for (int jj = 0; jj < MG_H / TPS; ++jj) {
    const int j = jj * TPS + jq;
    const float4* w = reinterpret_cast<const float4*>(S.gw1 + j * MG_GS);
    float a0 = S.gb1[j], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k4 = 0; k4 < MG_GS / 4; ++k4) {
      const float4 ww = w[k4];
      a0 = fmaf(x[4 * k4], ww.x, a0);
      a1 = fmaf(x[4 * k4 + 1], ww.y, a1);
      a2 = fmaf(x[4 * k4 + 2], ww.z, a2);
      a3 = fmaf(x[4 * k4 + 3], ww.w, a3);
    }
    const float hv = fmaxf((a0 + a1) + (a2 + a3), 0.f);
    if (STORE) hrow[j] = hv;
    p0 = fmaf(hv, S.gw2[j], p0);
    p1 = fmaf(hv, S.gw2[MG_H + j], p1);
  }
  o0 = group_sum<TPS>(p0) + S.gb2[0];
  o1 = group_sum<TPS>(p1) + S.gb2[1];
}

// Discriminator forward for one row: logit (summed over the four threads) and the ReLU mask of this thread's units.
template <int TPS>
__device__ __forceinline__ float d_forward(const float (&x)[MG_DS], const MlpGanSmem& S, int jq, uint32_t& mask) {
  float p = 0.f;
  mask = 0u;
#pragma unroll 4
  for (int jj = 0; jj < MG_H / TPS; ++jj) {
    const int j = jj * TPS + jq;
    const float4 w = *reinterpret_cast<const float4*>(S.dw1 + j * MG_DS);
    const float pre = fmaf(x[3], w.w, fmaf(x[2], w.z, fmaf(x[1], w.y, fmaf(x[0], w.x, S.db1[j]))));
    if (pre > 0.f) {
      mask |= 1u << jj;
      p = fmaf(pre, S.dw2[j], p);
    }
  }
  return group_sum<TPS>(p) + S.db2[0];
}

__device__ __forceinline__ void load_gen_input(float (&x)[MG_GS], const float* z, const float* oh, int zd, int ld,
                                               long long s, bool active) {
#pragma unroll
  for (int k = 0; k < MG_GS; ++k) x[k] = 0.f;
  if (!active) return;
#pragma unroll
  for (int k = 0; k < MG_GS; ++k) {
    if (k < zd) x[k] = __ldg(z + s * zd + k);
    else if (k < zd + ld) x[k] = __ldg(oh + s * ld + (k - zd));
  }
}

// Adam on this CTA's slice of a flat parameter buffer; the gradient is the rank-ordered sum of the CTAs' partials.
// `bcast` (discriminator): new values are also written into every CTA's shared-memory copy of the weights.
template <bool IS_D>
__device__ __forceinline__ void reduce_and_adam(cg::cluster_group& cluster, MlpGanSmem& S, float* part, int P, int in_dim,
                                                float* gp, float* gg, float* gm, float* gv, float step_size,
                                                float bc2_sqrt, const MlpGanArgs& a) {
  const int cs = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int slice = ((P + cs - 1) / cs + 3) / 4 * 4;
  const int lo = rank * slice, hi = min(P, lo + slice);
  const int ob1 = (MG_H * in_dim + 3) / 4 * 4, ow2 = ob1 + MG_H, ob2 = ow2 + (IS_D ? MG_H : 2 * MG_H);
  for (int i = lo + (int)threadIdx.x; i < hi; i += MG_T) {
    float g = 0.f;
    for (int r = 0; r < cs; ++r) g += cluster.map_shared_rank(part, r)[i];
    gg[i] = g;
    float mi = gm[i], vi = gv[i];
    mi = mi + (g - mi) * (1.f - a.beta1);
    vi = vi * a.beta2 + (1.f - a.beta2) * g * g;
    const float denom = sqrtf(vi) / bc2_sqrt + a.eps;
    const float pn = gp[i] - step_size * (mi / denom);
    gp[i] = pn;
    gm[i] = mi;
    gv[i] = vi;
    if (IS_D) {
      float* dst = nullptr;
      if (i < MG_H * in_dim) dst = S.dw1 + (i / in_dim) * MG_DS + (i % in_dim);
      else if (i >= ob1 && i < ob1 + MG_H) dst = S.db1 + (i - ob1);
      else if (i >= ow2 && i < ow2 + MG_H) dst = S.dw2 + (i - ow2);
      else if (i == ob2) dst = S.db2;
      if (dst != nullptr)
        for (int r = 0; r < cs; ++r) *cluster.map_shared_rank(dst, r) = pn;
    }
  }
}

template <int TPS>
__global__ void __launch_bounds__(MG_T, 1) mlp_gan_step_kernel(const MlpGanArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  MlpGanSmem& S = *reinterpret_cast<MlpGanSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), cs = (int)cluster.num_blocks();
  const int tid = threadIdx.x;
  const int gi = a.zd + a.ld, di = 2 + a.ld;
  const int g_ob1 = (MG_H * gi + 3) / 4 * 4, g_ow2 = g_ob1 + MG_H, g_ob2 = g_ow2 + 2 * MG_H, PG = g_ob2 + 4;
  const int d_ob1 = (MG_H * di + 3) / 4 * 4, d_ow2 = d_ob1 + MG_H, d_ob2 = d_ow2 + MG_H, PD = d_ob2 + 4;
  const int first = rank * a.nb;                                // first sample of this CTA
  const int nv = max(0, min(a.nb, a.B - first));                // its number of samples
  const float inv_b = 1.f / (float)a.B;

  // ---- weights -> shared memory (padded rows), partial buffers zeroed, Adam constants
  {
    // all global loads of a thread are issued before its first shared-memory store (one memory latency, not nine)
    constexpr int NG = MG_H * MG_GS / MG_T;                      // 9
    float vg[NG], vd, vs[5];
#pragma unroll
    for (int u = 0; u < NG; ++u) {
      const int i = tid + u * MG_T, j = i / MG_GS, k = i - j * MG_GS;
      vg[u] = k < gi ? __ldg(a.gp + j * gi + k) : 0.f;
    }
    {
      const int j = tid / MG_DS, k = tid - j * MG_DS;              // MG_H * MG_DS == MG_T
      vd = k < di ? __ldg(a.dp + j * di + k) : 0.f;
    }
    if (tid < MG_H) {
      vs[0] = __ldg(a.gp + g_ob1 + tid);
      vs[1] = __ldg(a.gp + g_ow2 + tid);
      vs[2] = __ldg(a.gp + g_ow2 + MG_H + tid);
      vs[3] = __ldg(a.dp + d_ob1 + tid);
      vs[4] = __ldg(a.dp + d_ow2 + tid);
    }
#pragma unroll
    for (int u = 0; u < NG; ++u) S.gw1[tid + u * MG_T] = vg[u];
    S.dw1[tid] = vd;
    if (tid < MG_H) {
      S.gb1[tid] = vs[0];
      S.gw2[tid] = vs[1];
      S.gw2[MG_H + tid] = vs[2];
      S.db1[tid] = vs[3];
      S.dw2[tid] = vs[4];
    }
  }
  if (tid < 2) S.gb2[tid] = __ldg(a.gp + g_ob2 + tid);
  if (tid == 2) S.db2[0] = __ldg(a.dp + d_ob2);
  for (int i = tid; i < MG_PG_MAX; i += MG_T) S.part_g[i] = 0.f;
  for (int i = tid; i < MG_PD_MAX; i += MG_T) S.part_d[i] = 0.f;
  if (tid >= MG_T - 2) {                                       // torch/optim/adam.py:_single_tensor_adam bias corrections
    const int w = tid - (MG_T - 2);                            // 0: discriminator, 1: generator
    unsigned t = (unsigned)(*(w == 0 ? a.dstep : a.gstep) + 1);
    double p1 = 1.0, p2 = 1.0, q1 = (double)a.beta1, q2 = (double)a.beta2;   // beta^t by squaring
    for (; t != 0u; t >>= 1) {
      if (t & 1u) { p1 *= q1; p2 *= q2; }
      q1 *= q1; q2 *= q2;
    }
    S.adam[w * 2] = (float)((double)a.lr / (1.0 - p1));
    S.adam[w * 2 + 1] = (float)sqrt(1.0 - p2);
  }
  __syncthreads();

  const int sl = tid / TPS, jq = tid % TPS;                      // phase A: sample, hidden-unit residue
  const bool active = sl < nv;
  const bool warp_active = (tid & ~31) / TPS < nv;              // some sample of this warp exists (warp-uniform)
  const long long sgl = (long long)first + sl;
  const int hj = tid & (MG_H - 1), sgp = tid >> 7;              // phase B: hidden unit, sample group
  const int per_group = (nv + 3) / 4;
  const int s_lo = min(nv, sgp * per_group), s_hi = min(nv, s_lo + per_group);
  float l_dr = 0.f, l_df = 0.f, p_r = 0.f, p_f = 0.f, l_g = 0.f;

  // ================= discriminator step: phase A =================
  if (warp_active) {
    float x[MG_GS];
    load_gen_input(x, a.z1, a.oh1, a.zd, a.ld, sgl, active);
    float f0, f1;
    g_forward<TPS, false>(x, S, jq, nullptr, f0, f1);
    float xr[MG_DS] = {0.f, 0.f, 0.f, 0.f}, xf[MG_DS] = {f0, f1, 0.f, 0.f};
    if (active) {
      xr[0] = __ldg(a.real + sgl * 2);
      xr[1] = __ldg(a.real + sgl * 2 + 1);
      for (int k = 0; k < a.ld; ++k) {
        xr[2 + k] = __ldg(a.real_oh + sgl * a.ld + k);
        xf[2 + k] = __ldg(a.oh1 + sgl * a.ld + k);
      }
    }
    uint32_t m_unused;
    const float zr = d_forward<TPS>(xr, S, jq, m_unused), zf = d_forward<TPS>(xf, S, jq, m_unused);
    if (jq == 0) {
      float dzr = 0.f, dzf = 0.f;
      if (active) {
        const float pr = 1.f / (1.f + expf(-zr)), pf = 1.f / (1.f + expf(-zf));
        l_dr = -logf(pr);
        l_df = -logf(1.f - pf);
        p_r = pr;
        p_f = pf;
        dzr = -(1.f - pr) * inv_b;                             // d/dz of -mean log sigmoid(z)
        dzf = pf * inv_b;                                      // d/dz of -mean log(1 - sigmoid(z))
      }
      S.dz[sl] = dzr;
      S.dz[MG_NB + sl] = dzf;
#pragma unroll
      for (int k = 0; k < MG_DS; ++k) {
        S.din[sl * MG_DS + k] = xr[k];
        S.din[(MG_NB + sl) * MG_DS + k] = xf[k];
      }
    }
  }
  __syncthreads();
  // ================= discriminator step: phase B (weight gradients of this CTA's rows) =================
  {
    const float4 w1 = *reinterpret_cast<const float4*>(S.dw1 + hj * MG_DS);
    const float b1 = S.db1[hj], w2 = S.dw2[hj];
    float aw1[MG_DS] = {0.f, 0.f, 0.f, 0.f}, ab1 = 0.f, aw2 = 0.f, ab2 = 0.f;
    for (int half = 0; half < 2; ++half)
      for (int s = s_lo; s < s_hi; ++s) {
        const int row = half * MG_NB + s;
        const float4 xv = *reinterpret_cast<const float4*>(S.din + row * MG_DS);
        const float dz = S.dz[row];
        const float pre = fmaf(xv.w, w1.w, fmaf(xv.z, w1.z, fmaf(xv.y, w1.y, fmaf(xv.x, w1.x, b1))));
        ab2 += dz;
        if (pre > 0.f) {
          const float ddh = dz * w2;
          aw2 = fmaf(dz, pre, aw2);
          ab1 += ddh;
          aw1[0] = fmaf(ddh, xv.x, aw1[0]);
          aw1[1] = fmaf(ddh, xv.y, aw1[1]);
          aw1[2] = fmaf(ddh, xv.z, aw1[2]);
          aw1[3] = fmaf(ddh, xv.w, aw1[3]);
        }
      }
    for (int r = 0; r < 4; ++r) {                               // fixed-order sum over the four sample groups
      if (sgp == r) {
        for (int k = 0; k < di; ++k) S.part_d[hj * di + k] += aw1[k];
        S.part_d[d_ob1 + hj] += ab1;
        S.part_d[d_ow2 + hj] += aw2;
        if (hj == 0) S.part_d[d_ob2] += ab2;
      }
      __syncthreads();
    }
  }
  cluster.sync();
  reduce_and_adam<true>(cluster, S, S.part_d, PD, di, a.dp, a.dg, a.dm, a.dv, S.adam[0], S.adam[1], a);
  cluster.sync();

  // ================= generator step: phase A =================
  if (warp_active) {
    float x[MG_GS];
    load_gen_input(x, a.z2, a.oh2, a.zd, a.ld, sgl, active);
    if (jq == 0) {
#pragma unroll
      for (int k4 = 0; k4 < MG_GS / 4; ++k4)
        *reinterpret_cast<float4*>(S.x + sl * MG_GS + 4 * k4) = make_float4(x[4 * k4], x[4 * k4 + 1], x[4 * k4 + 2], x[4 * k4 + 3]);
    }
    float f0, f1;
    g_forward<TPS, true>(x, S, jq, S.h + sl * MG_HS, f0, f1);
    float xf[MG_DS] = {f0, f1, 0.f, 0.f};
    if (active)
      for (int k = 0; k < a.ld; ++k) xf[2 + k] = __ldg(a.oh2 + sgl * a.ld + k);
    uint32_t mask;
    const float zf = d_forward<TPS>(xf, S, jq, mask);
    float dz = 0.f;
    if (active) {
      const float pf = 1.f / (1.f + expf(-zf));
      if (jq == 0) l_g = -logf(pf);
      dz = -(1.f - pf) * inv_b;
    }
    // d fake = DW1[:, 0:2]^T (dz * DW2 * relu'), over this thread's units, then over the four threads
    float d0 = 0.f, d1 = 0.f;
#pragma unroll 4
    for (int jj = 0; jj < MG_H / TPS; ++jj)
      if (mask & (1u << jj)) {
        const int j = jj * TPS + jq;
        const float ddh = dz * S.dw2[j];
        d0 = fmaf(ddh, S.dw1[j * MG_DS], d0);
        d1 = fmaf(ddh, S.dw1[j * MG_DS + 1], d1);
      }
    d0 = group_sum<TPS>(d0);
    d1 = group_sum<TPS>(d1);
    if (jq == 0) {
      S.df[sl * 2] = d0;
      S.df[sl * 2 + 1] = d1;
    }
  }
  __syncthreads();
  // ================= generator step: phase B =================
  {
    const float w20 = S.gw2[hj], w21 = S.gw2[MG_H + hj];
    float aw1[MG_GS], ab1 = 0.f, aw20 = 0.f, aw21 = 0.f, ab20 = 0.f, ab21 = 0.f;
#pragma unroll
    for (int k = 0; k < MG_GS; ++k) aw1[k] = 0.f;
    for (int s = s_lo; s < s_hi; ++s) {
      const float hv = S.h[s * MG_HS + hj];
      const float d0 = S.df[s * 2], d1 = S.df[s * 2 + 1];
      ab20 += d0;
      ab21 += d1;
      aw20 = fmaf(d0, hv, aw20);
      aw21 = fmaf(d1, hv, aw21);
      if (hv > 0.f) {
        const float dgh = fmaf(d1, w21, d0 * w20);
        ab1 += dgh;
        const float4* xs = reinterpret_cast<const float4*>(S.x + s * MG_GS);
#pragma unroll
        for (int k4 = 0; k4 < MG_GS / 4; ++k4) {
          const float4 xv = xs[k4];
          aw1[4 * k4] = fmaf(dgh, xv.x, aw1[4 * k4]);
          aw1[4 * k4 + 1] = fmaf(dgh, xv.y, aw1[4 * k4 + 1]);
          aw1[4 * k4 + 2] = fmaf(dgh, xv.z, aw1[4 * k4 + 2]);
          aw1[4 * k4 + 3] = fmaf(dgh, xv.w, aw1[4 * k4 + 3]);
        }
      }
    }
    for (int r = 0; r < 4; ++r) {
      if (sgp == r) {
#pragma unroll
        for (int k = 0; k < MG_GS; ++k)
          if (k < gi) S.part_g[hj * gi + k] += aw1[k];
        S.part_g[g_ob1 + hj] += ab1;
        S.part_g[g_ow2 + hj] += aw20;
        S.part_g[g_ow2 + MG_H + hj] += aw21;
        if (hj == 0) {
          S.part_g[g_ob2] += ab20;
          S.part_g[g_ob2 + 1] += ab21;
        }
      }
      __syncthreads();
    }
  }
  {
    const float t0 = cta_sum(l_dr, S.red), t1 = cta_sum(l_df, S.red), t2 = cta_sum(p_r, S.red), t3 = cta_sum(p_f, S.red),
                t4 = cta_sum(l_g, S.red);
    if (tid == 0) {
      S.loss[0] = t0; S.loss[1] = t1; S.loss[2] = t2; S.loss[3] = t3; S.loss[4] = t4;
    }
  }
  cluster.sync();
  reduce_and_adam<false>(cluster, S, S.part_g, PG, gi, a.gp, a.gg, a.gm, a.gv, S.adam[2], S.adam[3], a);
  if (rank == 0 && tid == 0) {
    float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < cs; ++r) {
      const float* l = cluster.map_shared_rank(S.loss, r);
      for (int q = 0; q < 5; ++q) t[q] += l[q];
    }
    a.scal[2] = t[0] * inv_b;
    a.scal[3] = t[1] * inv_b;
    a.scal[0] = t[0] * inv_b + t[1] * inv_b;
    a.scal[1] = t[4] * inv_b;
    a.scal[4] = t[2] * inv_b;
    a.scal[5] = t[3] * inv_b;
    *a.dstep += 1;
    *a.gstep += 1;
  }
  cluster.sync();                                               // keep shared memory alive until every remote read is done
}

void mlp_gan_step(int B, int z_dim, int label_dim, int hidden, const float* real, const float* real_oh, const float* z1,
                  const float* oh1, const float* z2, const float* oh2, float* g_param, float* g_grad, float* g_m,
                  float* g_v, int* g_step, float* d_param, float* d_grad, float* d_m, float* d_v, int* d_step, float lr,
                  float* scal, cudaStream_t stream) {
  PCG_PROFILE("mlp_gan_step", stream);
  PCG_REQUIRE(hidden == MG_H, "fused MLP GAN step: hidden width must be 128");
  PCG_REQUIRE(label_dim >= 0 && label_dim <= MG_DS - 2 && z_dim >= 1 && z_dim + label_dim <= MG_GS,
              "fused MLP GAN step: label_dim <= 2 and z_dim + label_dim <= 36");
  PCG_REQUIRE(label_dim == 0 || (real_oh != nullptr && oh1 != nullptr && oh2 != nullptr), "label one-hots are required");
  // cluster size: >= 16 samples per CTA; 16 CTAs (non-portable size) for the largest batches when the device can host
  // such a cluster.  Threads per sample: as many as the CTA's samples leave room for.
  static int max16 = -1;
  static bool configured = false;
  auto launch_cfg = [&](int cs, cudaLaunchAttribute* attr) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(MG_T);
    cfg.dynamicSmemBytes = sizeof(MlpGanSmem);
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
  };
  if (!configured) {
    void (*kernels[4])(MlpGanArgs) = {mlp_gan_step_kernel<4>, mlp_gan_step_kernel<8>, mlp_gan_step_kernel<16>,
                                      mlp_gan_step_kernel<32>};
    for (auto k : kernels) {
      PCG_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MlpGanSmem)));
      PCG_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = launch_cfg(16, attr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, mlp_gan_step_kernel<8>, &cfg) != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
    max16 = n;
    configured = true;
  }
  int cs = (B > 8 * 32 && max16 >= 1) ? 16 : 8;
  while (cs > 1 && B < cs * 16) cs /= 2;
  PCG_REQUIRE(B >= 1 && (B + cs - 1) / cs <= MG_NB, "fused MLP GAN step: batch must be 1..1024");
  MlpGanArgs a;
  a.B = B; a.zd = z_dim; a.ld = label_dim; a.nb = (B + cs - 1) / cs;
  a.real = real; a.real_oh = real_oh; a.z1 = z1; a.oh1 = oh1; a.z2 = z2; a.oh2 = oh2;
  a.gp = g_param; a.gg = g_grad; a.gm = g_m; a.gv = g_v; a.gstep = g_step;
  a.dp = d_param; a.dg = d_grad; a.dm = d_m; a.dv = d_v; a.dstep = d_step;
  a.lr = lr; a.beta1 = 0.9f; a.beta2 = 0.999f; a.eps = 1e-8f;
  a.scal = scal;
  cudaLaunchAttribute attr[1];
  cudaLaunchConfig_t cfg = launch_cfg(cs, attr);
  if (a.nb <= MG_T / 32) PCG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mlp_gan_step_kernel<32>, a));
  else if (a.nb <= MG_T / 16) PCG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mlp_gan_step_kernel<16>, a));
  else if (a.nb <= MG_T / 8) PCG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mlp_gan_step_kernel<8>, a));
  else PCG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mlp_gan_step_kernel<4>, a));
  PCG_COUNT_LAUNCH();
}

}  // namespace pcg

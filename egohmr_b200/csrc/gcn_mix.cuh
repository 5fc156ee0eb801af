// Joint-mix warps of the hidden-layer kernels (gcn_umma.cu: row-major product, gcn_umma_t.cu: transposed product).
//
// ModulatedGraphConv mixes the 24 joints of a (body, pass) slot with the layer's symmetric adjacency
// (modulated_gcn_conv.py:38-50):  y[j] = adj[j][j] (M o h0)[j] + sum_{i != j} adj[j][i] (M o h1)[i], then BatchNorm1d(eval) + ReLU
// (+ the _ResGraphConv residual, modulated_gcn.py:38-42).  The tcgen05.ld warps leave  D = diag (M o h0)  and  G = M o h1  of one
// hand-off chunk (5 slots x 32 channels) in channel-major shared-memory tiles; a mix warp owns (slot w, 12 output joints from
// j0), lane = channel: it reads its slot's 24 G values and 12 D values, releases the chunk, does the 12 x 24 FMAs with the
// adjacency straight from the kernel-parameter constant bank, and writes the fp32 block-boundary activations and / or the
// next layer's fp16 hi|lo operand.
#pragma once
#include "kernels.cuh"
#include "ptx.cuh"

namespace ehb {

constexpr int MIX_NJH = NJ / 2;   // output joints per mix warp

// `p` must be the kernel's __grid_constant__ parameter (the adjacency is indexed with compile-time constants only).
// row0 = first output row of this warp (tile row of slot w, joint j0); c = this lane's channel.
struct MixNoHook {
  __device__ __forceinline__ void operator()() const {}
};

// after_wait: called right after the chunk's hand-off wait returns (the fused kernel publishes the PREVIOUS unit's completion
// there: its stores have drained during the wait, so the fence that must precede the count costs nothing).
template <int GT_LD, bool LATE_RES = false, typename AfterWait = MixNoHook>
__device__ __forceinline__ void mix_chunk(const HiddenLayerParams& p, const float* G_T, const float* D_T, uint64_t* cfull,
                                          uint64_t* cempty, uint32_t chunk_it, bool valid, int c, size_t row0, int w, int j0,
                                          int lane, float& amax, AfterWait after_wait = AfterWait()) {
  constexpr int NJH = MIX_NJH;
  float g[NJ], y[NJH], rsd[NJH];
  // residual rows are fetched before the hand-off wait: independent loads in flight, latency hidden.  (LATE_RES: after it —
  // in the fused kernel a unit's rows may still be in flight from another CTA until its first chunk is staged; the double-
  // buffered accumulator leaves the mix warps a whole unit of slack for the exposed latency.)
  const float* rp = p.res + row0 * p.C + c;
#define EHB_LOAD_RES()                                                                     \
  do {                                                                                     \
    _Pragma("unroll") for (int jj = 0; jj < NJH; ++jj)                                     \
      rsd[jj] = (p.add_res && valid) ? __ldcg(rp + static_cast<size_t>(jj) * p.C) : 0.f;   \
  } while (0)
  if constexpr (!LATE_RES) EHB_LOAD_RES();
  ptx::mbar_wait(cfull, chunk_it & 1);
  after_wait();
  if constexpr (LATE_RES) EHB_LOAD_RES();
#undef EHB_LOAD_RES
  {
    const float4* gp = reinterpret_cast<const float4*>(G_T + lane * GT_LD + NJ * w);
    const float4* dp = reinterpret_cast<const float4*>(D_T + lane * GT_LD + NJ * w + j0);
#pragma unroll
    for (int v = 0; v < NJ / 4; ++v) {
      const float4 a = gp[v];
      g[4 * v + 0] = a.x; g[4 * v + 1] = a.y; g[4 * v + 2] = a.z; g[4 * v + 3] = a.w;
    }
#pragma unroll
    for (int v = 0; v < NJH / 4; ++v) {
      const float4 b = dp[v];
      y[4 * v + 0] = b.x; y[4 * v + 1] = b.y; y[4 * v + 2] = b.z; y[4 * v + 3] = b.w;
    }
  }
  ptx::mbar_arrive(cempty);
  if (!valid) return;
  if (j0 == 0) {
#pragma unroll
    for (int jj = 0; jj < NJH; ++jj) {
      float acc = y[jj];
#pragma unroll
      for (int i = 0; i < NJ; ++i) acc = fmaf(p.adj.off[jj][i], g[i], acc);
      y[jj] = acc;
    }
  } else {
#pragma unroll
    for (int jj = 0; jj < NJH; ++jj) {
      float acc = y[jj];
#pragma unroll
      for (int i = 0; i < NJ; ++i) acc = fmaf(p.adj.off[NJH + jj][i], g[i], acc);
      y[jj] = acc;
    }
  }
  const float sc = __ldg(p.bn_scale + c);
  const float sh = __ldg(p.bn_shift + c);
  float* fp = p.res + row0 * p.C + c;
  __half* hp = p.out_hl + row0 * (2 * static_cast<size_t>(p.C)) + c;
#pragma unroll
  for (int jj = 0; jj < NJH; ++jj) {
    const float v = fmaxf(fmaf(y[jj], sc, sh), 0.f) + rsd[jj];
    if (p.write_f32) fp[static_cast<size_t>(jj) * p.C] = v;
    if (p.write_hl) {
      const float sv = v * p.act_scale;
      const __half hi = __float2half_rn(sv);
      const __half lo = __float2half_rn(sv - __half2float(hi));
      hp[static_cast<size_t>(jj) * 2 * p.C] = hi;
      hp[static_cast<size_t>(jj) * 2 * p.C + p.C] = lo;
      amax = fmaxf(amax, fabsf(sv));
    }
  }
}

}  // namespace ehb

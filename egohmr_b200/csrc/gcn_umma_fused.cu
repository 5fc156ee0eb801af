// K1, fused over the hidden layers of one reverse step — the transposed-product kernel of gcn_umma_t.cu as ONE persistent
// launch for all (up to 8) ModulatedGraphConv layers.
//
// Reference semantics: as gcn_umma.cu / gcn_umma_t.cu (modulated_gcn_conv.py:38-50, modulated_gcn.py:21-28,38-42,99-113: the
// residual blocks are applied one after the other, each layer reads the previous layer's activations).
//
// Why.  A per-layer launch costs its ramp and its tail: the first unit waits 12.5 k cycles for operands while all 148 CTAs
// fill their rings at once, and the last unit's epilogue runs 18 k cycles with the tensor pipe idle — 8 % of a launch whose
// steady state is at 96-99 % (profiles/r02b_k1_timeline.txt).  Here the units of all layers form one sequence in layer-major
// order, CTA pair q takes units q, q + G, ..., and the only thing a unit (layer l, row group r, channel group c) needs from
// layer l - 1 is the 8 channel-group units of ITS row group r.  Every mix warp bumps a counter per (layer, row group) when
// its part of a unit is stored — through the CTA's publisher warp, see below; the TMA producer of a unit acquires the
// counter of (l - 1, r) before it loads that unit's activation rows.  For all but the stragglers that counter has been complete
// for ~1000 units, so the head of layer l + 1 overlaps the tail of layer l and the pipeline never drains between layers.
// The same counter covers the write-after-read hazard of the ping-pong operand buffers (layer l + 1 overwrites the rows
// layer l - 1 ... l read only after all of layer l's units on those rows are done).  All CTAs are co-resident (one per SM,
// grid <= SM count), units only depend on units earlier in the sequence, every CTA works through its units in order: no
// deadlock.  Kernel parameters (8 layers' adjacencies and pointers, 10 tensor maps) are ~21 KB of __grid_constant__ data.
#include "gcn_mix.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

#ifndef EHB_FUSED_LATE_RES
#define EHB_FUSED_LATE_RES true
#endif

namespace ehb {
namespace {

constexpr int WROWS = 128;        // weight rows (TMEM lanes) per CTA: 64 channels x {h0, h1}
constexpr int XROWS = SLOTS_PER_TILE * NJ;   // 120 activation rows per CTA: the real rows of one 128-row tile
constexpr int UMMA_N = 2 * XROWS; // 240
constexpr int ACC_STRIDE = 256;   // TMEM columns between the two accumulator buffers
constexpr int BK = EHB_UMMA_BK;
static_assert(BK == 64, "the transposed kernel is written for 128-byte swizzled k-blocks");
constexpr int SWZ = BK * 2;
constexpr int UMMA_K = 16;
constexpr int W_BYTES = WROWS * BK * 2;      // 16 KiB: rows 0-63 h0, 64-127 h1
constexpr int X_BYTES = XROWS * BK * 2;      // 15 KiB (a multiple of the 1 KiB swizzle atom)
constexpr int STAGE_BYTES = 2 * W_BYTES + 2 * X_BYTES;   // hi + lo of both operands
constexpr int STAGES = 3;
constexpr int CHUNK = 32;         // channels per epilogue hand-off
constexpr int GT_LD = 132;        // padded row length (floats) of the channel-major staging tiles
constexpr int EPI_BYTES = 2 * CHUNK * GT_LD * 4;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB of dynamic shared memory");
// warp roles as in gcn_umma.cu: 0-3 tcgen05.ld (warp index == TMEM lane quadrant: quadrants 0,1 hold h0 of channels
// 0-31 / 32-63 of this CTA's 64, quadrants 2,3 the h1 lanes), 4-13 joint mix / store, 14 TMA producer, 15 MMA issuer.
constexpr int NUM_WARPS = 16;
constexpr int NUM_MIX_WARPS = 10;
constexpr int NUM_THREADS = NUM_WARPS * 32;
constexpr int LD_WARP0 = 0, MIX_WARP0 = 4, TMA_WARP = 14, MMA_WARP = 15;
constexpr int TMEM_COLS = 512;
constexpr int CHUNKS_PER_UNIT = 4;   // (tile of the pair) x (32-channel half)

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint64_t cfull;
  uint64_t cempty;
  uint64_t udone;     // the mix warps have stored a unit
  uint32_t tmem_base;
};
static_assert(sizeof(Barriers) <= BAR_BYTES, "barrier block too small");

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gcn_hidden_fused_kernel(const __grid_constant__ FusedHiddenMaps maps, const __grid_constant__ FusedHiddenParams fp) {
  const HiddenLayerParams& p = fp.layer[0];   // the fields every layer shares (C, tile counts, slots, res, flags)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  float* D_T = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  float* G_T = D_T + CHUNK * GT_LD;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);

  __shared__ float s_diag[MAX_FUSED_LAYERS * NJ];   // the layers' adjacency diagonals (a runtime layer index cannot be an
                                                    // immediate constant-bank operand)
  for (int i = threadIdx.x; i < fp.n_layers * NJ; i += NUM_THREADS) s_diag[i] = fp.layer[i / NJ].adj.diag[i % NJ];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = p.C / BK;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  // work units: (pair of 128-row activation tiles) x (128-channel group); the channel group runs fastest so that the
  // units in flight share their activation rows in L2
  const int n_cgroups = p.n_ntiles;
  const int n_rgroups = p.n_mtiles / 2;
  const int upl = n_rgroups * n_cgroups;        // units per layer
  const int total_units = fp.n_layers * upl;
  const int done_target = 2 * n_cgroups;   // both CTAs of every channel-group unit of the row group
  const int unit0 = blockIdx.x / 2;
  const int unit_step = gridDim.x / 2;

  if (warp == TMA_WARP && lane == 0) {
    ptx::prefetch_tensormap(&maps.x[0]);
    ptx::prefetch_tensormap(&maps.x[1]);
    for (int l = 0; l < fp.n_layers; ++l) ptx::prefetch_tensormap(&maps.w[l]);
  }
  if (warp == MMA_WARP && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&bars->full[s], 1);      // the leader's producer arms it; both CTAs' bytes are tracked by expect_tx
      ptx::mbar_init(&bars->empty[s], 1);     // one multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bars->tfull[s], 1);
      ptx::mbar_init(&bars->tempty[s], 4 * 2);   // one elected lane of each tcgen05.ld warp of each CTA
    }
    ptx::mbar_init(&bars->cfull, 2 * 32);     // the two tcgen05.ld warps that own a chunk's channels
    ptx::mbar_init(&bars->cempty, NUM_MIX_WARPS * 32);
    ptx::mbar_init(&bars->udone, NUM_MIX_WARPS);
    ptx::fence_mbar_init();
  }
  if (warp == TMA_WARP) {
    ptx::tmem_alloc_2sm(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();   // the peer's barriers must be initialised before anything remote touches them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  // programmatic dependent launch: the set-up above overlaps the previous layer's tail (see gcn_umma.cu)
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();

  if (warp == TMA_WARP) {
    // ------------------------------------------------------------------ TMA producer (both CTAs of the pair)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int g = unit0; g < total_units; g += unit_step) {
        const int layer = g / upl, u = g - layer * upl, rg = u / n_cgroups;
        const CUtensorMap* tX = &maps.x[layer & 1];   // layer l reads operand buffer l & 1 and writes the other one
        const CUtensorMap* tW = &maps.w[layer];
        const int x_row = (rg * 2 + static_cast<int>(rank)) * TILE_ROWS;   // this CTA's tile, rows 0..119
        // weight rows of channel group cg are stored as [128 x h0 | 128 x h1]; this CTA takes 64 of each
        const int w_row0 = (u % n_cgroups) * 256 + static_cast<int>(rank) * 64;
        const int w_row1 = w_row0 + 128;
#ifndef EHB_FUSED_NOFLAG
        if (layer > 0) {
          // the previous layer's 8 channel-group units of this row group: all their mix warps have stored and fenced
          const int* flag = fp.done + (layer - 1) * n_rgroups + rg;
          while (ld_acquire_gpu(flag) < done_target) __nanosleep(100);
#ifndef EHB_FUSED_NOPROXY
          ptx::fence_proxy_async_all();   // their generic-proxy stores -> this thread's async-proxy (TMA) reads
#endif
        }
#endif
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* s = stage_base + stage * STAGE_BYTES;
          const uint32_t lfull = ptx::mapa(ptx::smem_u32(&bars->full[stage]), 0);  // the leader's barrier
          if (leader) ptx::mbar_arrive_expect_tx(&bars->full[stage], 2 * STAGE_BYTES);
          ptx::tma_load_2d_2sm(s, tW, lfull, kb * BK, w_row0);
          ptx::tma_load_2d_2sm(s + W_BYTES / 2, tW, lfull, kb * BK, w_row1);
          ptx::tma_load_2d_2sm(s + W_BYTES, tW, lfull, p.C + kb * BK, w_row0);
          ptx::tma_load_2d_2sm(s + W_BYTES + W_BYTES / 2, tW, lfull, p.C + kb * BK, w_row1);
          ptx::tma_load_2d_2sm(s + 2 * W_BYTES, tX, lfull, kb * BK, x_row);
          ptx::tma_load_2d_2sm(s + 2 * W_BYTES + X_BYTES, tX, lfull, p.C + kb * BK, x_row);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(2 * WROWS, UMMA_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int g = unit0; g < total_units; g += unit_step) {
        ptx::mbar_wait_cluster(&bars->tempty[as], aphase ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t tacc = tmem_base + as * ACC_STRIDE;
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(&bars->full[stage], phase);
          ptx::tc_fence_after_sync();
          {
            const uint32_t sa = ptx::smem_u32(stage_base + stage * STAGE_BYTES);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint32_t koff = ks * UMMA_K * 2;
              const uint64_t w_hi = ptx::make_kmajor_desc<SWZ>(sa + koff);
              const uint64_t w_lo = ptx::make_kmajor_desc<SWZ>(sa + W_BYTES + koff);
              const uint64_t x_hi = ptx::make_kmajor_desc<SWZ>(sa + 2 * W_BYTES + koff);
              const uint64_t x_lo = ptx::make_kmajor_desc<SWZ>(sa + 2 * W_BYTES + X_BYTES + koff);
              const uint32_t first = (kb | ks) != 0 ? 1u : 0u;
              // same term order as gcn_umma.cu: x_hi.w_hi, x_hi.w_lo, x_lo.w_hi
              ptx::umma_f16_2sm_elect(tacc, w_hi, x_hi, idesc, first);
              ptx::umma_f16_2sm_elect(tacc, w_lo, x_hi, idesc, 1u);
              ptx::umma_f16_2sm_elect(tacc, w_hi, x_lo, idesc, 1u);
            }
            ptx::umma_commit_2sm_mc_elect(&bars->empty[stage], 0b11);
            if (kb == KB - 1) ptx::umma_commit_2sm_mc_elect(&bars->tfull[as], 0b11);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp < LD_WARP0 + 4) {
    // ------------------------------------------------------------------ TMEM -> modulate -> shared (channel-major)
    const int q = warp - LD_WARP0;  // == warp % 4: the TMEM lane quadrant this warp may read
    const bool is_h1 = q >= 2;      // quadrants 0,1: h0 lanes; 2,3: h1 lanes
    const int ck = q & 1;           // which 32-channel half of the CTA's 64 channels
    float* dst = (is_h1 ? G_T : D_T) + lane * GT_LD;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t chunk_it = 0;
    uint32_t uphase = 0;
    for (int g = unit0; g < total_units; g += unit_step, chunk_it += CHUNKS_PER_UNIT) {
      const int layer = g / upl, u = g - layer * upl;
      const HiddenLayerParams& lp = fp.layer[layer];
      const int c = (u % n_cgroups) * 128 + static_cast<int>(rank) * 64 + ck * CHUNK + lane;
      // this lane's channel of the modulation matrix; the h0 lanes also apply the adjacency diagonal, as
      // diag * (M * h0) to keep gcn_umma.cu's rounding
      float m[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) m[j] = __ldg(lp.mod + static_cast<size_t>(j) * p.C + c);
      const float* dg = s_diag + layer * NJ;   // the h0 lanes fold the adjacency diagonal in below: diag * (M * h0)
      ptx::mbar_wait(&bars->tfull[as], aphase);
      ptx::tc_fence_after_sync();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * ACC_STRIDE;
#pragma unroll 1
      for (int k = 0; k < CHUNKS_PER_UNIT; ++k) {
        // chunk k = (tile k / 2 of the pair = accumulator columns 0..119 / 120..239) x (32-channel half k % 2); every
        // tcgen05.ld warp observes every chunk's release in order (a warp that skipped a phase of `cempty` could not tell
        // it from the one two chunks later), and fills the chunks of its own channel half
        const uint32_t it = chunk_it + k;
        const bool mine = (k & 1) == ck;
        const uint32_t tcol = trow + (k >> 1) * XROWS;
        float va[NJ], vb[NJ];
        if (mine) {
          ptx::tmem_ld8(tcol, va);
          ptx::tmem_ld16(tcol + 8, va + 8);
          ptx::tmem_ld_wait();
        }
        ptx::mbar_wait(&bars->cempty, (it & 1) ^ 1);
        if (!mine) continue;
#pragma unroll
        for (int s = 0; s < SLOTS_PER_TILE; ++s) {
          float* cur = (s & 1) ? vb : va;
          float* nxt = (s & 1) ? va : vb;
          if (s + 1 < SLOTS_PER_TILE) {
            ptx::tmem_ld8(tcol + (s + 1) * NJ, nxt);
            ptx::tmem_ld16(tcol + (s + 1) * NJ + 8, nxt + 8);
          }
          float o[NJ];
          if (is_h1) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) o[j] = m[j] * cur[j];
          } else {
#pragma unroll
            for (int j = 0; j < NJ; ++j) o[j] = dg[j] * (m[j] * cur[j]);
          }
#pragma unroll
          for (int t = 0; t < NJ / 4; ++t)
            *reinterpret_cast<float4*>(dst + s * NJ + 4 * t) = make_float4(o[4 * t], o[4 * t + 1], o[4 * t + 2], o[4 * t + 3]);
          if (s + 1 < SLOTS_PER_TILE) ptx::tmem_ld_wait();
        }
        if (k >= 2) {
          // this warp's lanes of the accumulator are drained: hand the TMEM stage back to the leader's MMA warp
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            if (leader) ptx::mbar_arrive(&bars->tempty[as]);
            else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->tempty[as]), 0));
          }
        }
        ptx::mbar_arrive(&bars->cfull);
      }
      if (q == 0) {
        // publisher: all mix warps of this CTA have stored the unit (acquire at CTA scope) -> release at device scope ->
        // count this CTA in for (layer, row group).  This warp is idle until the next accumulator completes.
        ptx::mbar_wait(&bars->udone, uphase);
        uphase ^= 1;
        if (lane == 0) {
#ifndef EHB_FUSED_NOFENCE
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
          atomicAdd(fp.done + layer * n_rgroups + u / n_cgroups, 1);
        }
        __syncwarp();
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else if (warp >= MIX_WARP0 && warp < MIX_WARP0 + NUM_MIX_WARPS) {
    // ------------------------------------------------------------------ joint mix + BN + ReLU (+res) + store
    const int w = (warp - MIX_WARP0) % SLOTS_PER_TILE;   // slot within the tile
    const int j0 = ((warp - MIX_WARP0) / SLOTS_PER_TILE) * MIX_NJH;  // first output joint of this warp
    uint32_t chunk_it = 0;
    float amax = 0.f;
    for (int g = unit0; g < total_units; g += unit_step) {
      const int layer = g / upl, u = g - layer * upl;
#pragma unroll 1
      for (int ch = 0; ch < CHUNKS_PER_UNIT; ++ch, ++chunk_it) {
        // chunk = (tile ch / 2 of the pair) x (32-channel half ch % 2 of this CTA's 64 channels)
        const int m_tile = (u / n_cgroups) * 2 + (ch >> 1);
        const bool valid = (m_tile * SLOTS_PER_TILE + w) < p.n_slots;
        const int c = (u % n_cgroups) * 128 + static_cast<int>(rank) * 64 + (ch & 1) * CHUNK + lane;
        const size_t row0 = static_cast<size_t>(m_tile) * TILE_ROWS + NJ * w + j0;
        // the residual rows are read only once the chunk is staged: by then this CTA knows that the unit's dependency on the
        // previous layers is met (LATE_RES).  One case per layer keeps the adjacency an immediate constant-bank operand.
#define EHB_MIX_CASE(L)                                                                                                     \
  case L:                                                                                                                  \
    mix_chunk<GT_LD, EHB_FUSED_LATE_RES>(fp.layer[L], G_T, D_T, &bars->cfull, &bars->cempty, chunk_it, valid, c, row0, w, j0, \
                                         lane, amax);                                                                      \
    break;
        switch (layer) {
          EHB_MIX_CASE(0) EHB_MIX_CASE(1) EHB_MIX_CASE(2) EHB_MIX_CASE(3)
          EHB_MIX_CASE(4) EHB_MIX_CASE(5) EHB_MIX_CASE(6) EHB_MIX_CASE(7)
        }
#undef EHB_MIX_CASE
      }
      // this warp's share of the unit is stored: tell the CTA's publisher (tcgen05.ld warp 0).  A device-scope fence here, in
      // each of the 10 mix warps, stalls them for the store round trip and costs 8-10 % of the kernel (tools/ab_libs.sh);
      // the arrive is a CTA-scope release, and the publisher's device-scope release is cumulative over it.
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->udone);
    }
    if (!(amax <= 65504.f)) atomicExch(p.overflow_flag, 1);  // also catches NaN
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();   // the peer may still be reading this CTA's operands / signalling its barriers
  if (warp == TMA_WARP) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace

cudaError_t launch_gcn_hidden_fused(const FusedHiddenMaps& maps, const FusedHiddenParams& fp, int num_sms, bool pdl,
                                    cudaStream_t stream) {
  const HiddenLayerParams& p = fp.layer[0];
  if (p.C % 128 != 0 || p.n_ntiles != p.C / 128 || p.n_mtiles % 2 != 0 || fp.n_layers < 1 || fp.n_layers > MAX_FUSED_LAYERS)
    return cudaErrorInvalidValue;
  static bool attr_set = false;
  auto kern = gcn_hidden_fused_kernel;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int n_rgroups = p.n_mtiles / 2;
  const int units = fp.n_layers * n_rgroups * p.n_ntiles;
  if (units == 0) return cudaSuccess;
  // the dependency counters: [n_layers][n_rgroups], zero before every launch
  cudaError_t e = cudaMemsetAsync(fp.done, 0, sizeof(int) * fp.n_layers * n_rgroups, stream);
  if (e != cudaSuccess) return e;
  // every CTA must be resident (units wait for units of other CTAs): one CTA per SM, never more CTAs than SMs — and never
  // more CTA pairs than the device reports as co-resident for this kernel (fewer than SMs / 2 on a partitioned device)
  int grid = units * 2 < num_sms ? units * 2 : (num_sms / 2) * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  static int max_pairs = -1;
  if (max_pairs < 0) {
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) return e;
    max_pairs = n;
  }
  if (max_pairs < 1) return cudaErrorLaunchOutOfResources;
  if (grid > 2 * max_pairs) grid = 2 * max_pairs;
  cfg.gridDim = dim3(grid);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, maps, fp);
}

}  // namespace ehb

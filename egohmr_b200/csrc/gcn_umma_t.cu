// K1, transposed — one hidden ModulatedGraphConv (+ BatchNorm1d(eval) + ReLU [+ residual]) of the EgoHMR denoiser with the
// WEIGHTS as the M operand and the activation rows as the N operand of the tcgen05 product (CTA pairs, cta_group::2).
//
// Reference semantics: as gcn_umma.cu (models/egohmr/modulated_gcn/modulated_gcn_conv.py:38-50, modulated_gcn.py:21-28,38-42).
//
// Why transposed.  A (body, pass) slot is 24 rows, and the M side of a UMMA is 128 lanes per CTA: gcn_umma.cu packs 5 slots
// into a 128-row tile and pays for 8 pad rows (6.25 % of the MMA time).  The N side only has to be a multiple of 16, and a
// cta_group::2 MMA costs time proportional to N (tools/probes/umma_n_probe.cu: 69.6 ns at N = 256, 65.3 ns at N = 240), so
// here a unit is  D[256 weight rows][240 activation rows]:
//   M: per CTA 64 channels of h0 | the same 64 channels of h1   (two 64-row TMA boxes out of the existing weight layout)
//   N: 10 slots = rows 0..119 of the pair's two 128-row activation tiles, 120 per CTA (the pair shares the N operand)
//   one 240-column fp32 accumulator, double-buffered at TMEM columns 0 and 256.
// Shared-memory operand reads per CTA and MMA: 4 KB of weights + 3.75 KB of activations per 60 cycles = the same
// 128 B/clk as the row-major kernel.  A TMEM lane is now a channel and a slot's 24 joints are 24 accumulator columns: the
// tcgen05.ld warps (lane = channel) apply the modulation and write the (channel-major) staging tiles with 128-bit stores,
// and the mix warps are the ones of gcn_umma.cu unchanged (slot x joint-half each, lane = channel).  A hand-off chunk is
// (one of the pair's two tiles) x (32 channels): filled by the two tcgen05.ld warps that own those channels' h0 / h1 lanes
// while the other two already fetch the next chunk from TMEM.
#include "gcn_mix.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

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
  uint32_t tmem_base;
};
static_assert(sizeof(Barriers) <= BAR_BYTES, "barrier block too small");

// ---- optional timeline (nvcc -DEHB_K1_TRACE; tools/k1_trace.sh): per unit of CTA 0 of the first launch after arming, the cycles
// the MMA warp waited for a free accumulator and for operands, the TMA warp for free stages, one tcgen05.ld warp for the
// accumulator and one mix warp for its chunks.  Compiled out of the product build.
#ifdef EHB_K1_TRACE
constexpr int TRACE_UNITS = 16, TRACE_SLOTS = 16;
__device__ long long g_k1_trace[TRACE_UNITS][TRACE_SLOTS];
__device__ int g_k1_armed = 0;
__device__ unsigned long long g_k1_pair_t[128][2];   // per CTA pair: globaltimer (ns) when its MMA warp started / issued its last unit
__device__ __forceinline__ unsigned long long k1_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define K1_TRACE(unit_i, slot, val)                                                        \
  do {                                                                                     \
    if (trace_on && (unit_i) < TRACE_UNITS) g_k1_trace[(unit_i)][(slot)] = (val);          \
  } while (0)
#define K1_CLOCK() clock64()
#else
#define K1_TRACE(unit_i, slot, val) \
  do {                              \
  } while (0)
#define K1_CLOCK() 0LL
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gcn_hidden_umma_t_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                         const __grid_constant__ HiddenLayerParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  float* D_T = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  float* G_T = D_T + CHUNK * GT_LD;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = p.C / BK;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  // work units: (pair of 128-row activation tiles) x (128-channel group); the channel group runs fastest so that the
  // units in flight share their activation rows in L2
  const int n_cgroups = p.n_ntiles;
  const int total_units = (p.n_mtiles / 2) * n_cgroups;
  const int unit0 = blockIdx.x / 2;
  const int unit_step = gridDim.x / 2;
#ifdef EHB_K1_TRACE
  const bool trace_on = blockIdx.x == 0 && g_k1_armed == 1 && lane == 0;
#endif
  [[maybe_unused]] int ui = 0;   // unit ordinal of this CTA (timeline only)

  if (warp == TMA_WARP && lane == 0) {
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW);
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
      for (int u = unit0; u < total_units; u += unit_step) {
        const int x_row = ((u / n_cgroups) * 2 + static_cast<int>(rank)) * TILE_ROWS;   // this CTA's tile, rows 0..119
        // weight rows of channel group cg are stored as [128 x h0 | 128 x h1]; this CTA takes 64 of each
        const int w_row0 = (u % n_cgroups) * 256 + static_cast<int>(rank) * 64;
        const int w_row1 = w_row0 + 128;
        [[maybe_unused]] long long w_empty = 0;
        for (int kb = 0; kb < KB; ++kb) {
          const long long c0 = K1_CLOCK();
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          w_empty += K1_CLOCK() - c0;
          uint8_t* s = stage_base + stage * STAGE_BYTES;
          const uint32_t lfull = ptx::mapa(ptx::smem_u32(&bars->full[stage]), 0);  // the leader's barrier
          if (leader) ptx::mbar_arrive_expect_tx(&bars->full[stage], 2 * STAGE_BYTES);
          ptx::tma_load_2d_2sm(s, &tmW, lfull, kb * BK, w_row0);
          ptx::tma_load_2d_2sm(s + W_BYTES / 2, &tmW, lfull, kb * BK, w_row1);
          ptx::tma_load_2d_2sm(s + W_BYTES, &tmW, lfull, p.C + kb * BK, w_row0);
          ptx::tma_load_2d_2sm(s + W_BYTES + W_BYTES / 2, &tmW, lfull, p.C + kb * BK, w_row1);
          ptx::tma_load_2d_2sm(s + 2 * W_BYTES, &tmX, lfull, kb * BK, x_row);
          ptx::tma_load_2d_2sm(s + 2 * W_BYTES + X_BYTES, &tmX, lfull, p.C + kb * BK, x_row);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        K1_TRACE(ui, 8, w_empty);
        K1_TRACE(ui, 9, K1_CLOCK());
        ++ui;
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
#ifdef EHB_K1_TRACE
      if (g_k1_armed == 1 && lane == 0 && unit0 < 128) g_k1_pair_t[unit0][0] = k1_globaltimer();
#endif
      for (int u = unit0; u < total_units; u += unit_step) {
        [[maybe_unused]] const long long t_unit = K1_CLOCK();
        ptx::mbar_wait_cluster(&bars->tempty[as], aphase ^ 1);
        [[maybe_unused]] const long long t_acc = K1_CLOCK();
        [[maybe_unused]] long long w_full = 0;
        ptx::tc_fence_after_sync();
        const uint32_t tacc = tmem_base + as * ACC_STRIDE;
        for (int kb = 0; kb < KB; ++kb) {
          const long long c0 = K1_CLOCK();
          ptx::mbar_wait(&bars->full[stage], phase);
          w_full += K1_CLOCK() - c0;
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
        K1_TRACE(ui, 0, t_unit);
        K1_TRACE(ui, 1, t_acc - t_unit);     // waited for a free accumulator
        K1_TRACE(ui, 2, w_full);             // waited for operands (sum over the unit's k-blocks)
        K1_TRACE(ui, 3, K1_CLOCK());         // all MMAs of the unit issued
#ifdef EHB_K1_TRACE
        if (g_k1_armed == 1 && lane == 0 && unit0 < 128) g_k1_pair_t[unit0][1] = k1_globaltimer();
#endif
        ++ui;
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
    for (int u = unit0; u < total_units; u += unit_step, chunk_it += CHUNKS_PER_UNIT) {
      const int c = (u % n_cgroups) * 128 + static_cast<int>(rank) * 64 + ck * CHUNK + lane;
      // this lane's channel of the modulation matrix; the h0 lanes also apply the adjacency diagonal, as
      // diag * (M * h0) to keep gcn_umma.cu's rounding
      float m[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) m[j] = __ldg(p.mod + static_cast<size_t>(j) * p.C + c);
      [[maybe_unused]] const long long t_ld0 = K1_CLOCK();
      ptx::mbar_wait(&bars->tfull[as], aphase);
      if (q == 0) {
        K1_TRACE(ui, 4, K1_CLOCK() - t_ld0);   // tcgen05.ld warp 0 waited for the accumulator
        K1_TRACE(ui, 5, K1_CLOCK());           // accumulator complete (MMAs retired)
      }
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
            for (int j = 0; j < NJ; ++j) o[j] = p.adj.diag[j] * (m[j] * cur[j]);
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
      if (q == 3) K1_TRACE(ui, 6, K1_CLOCK());   // last chunk staged
      ++ui;
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
    for (int u = unit0; u < total_units; u += unit_step) {
#pragma unroll 1
      for (int ch = 0; ch < CHUNKS_PER_UNIT; ++ch, ++chunk_it) {
        // chunk = (tile ch / 2 of the pair) x (32-channel half ch % 2 of this CTA's 64 channels)
        const int m_tile = (u / n_cgroups) * 2 + (ch >> 1);
        const bool valid = (m_tile * SLOTS_PER_TILE + w) < p.n_slots;
        const int c = (u % n_cgroups) * 128 + static_cast<int>(rank) * 64 + (ch & 1) * CHUNK + lane;
        const size_t row0 = static_cast<size_t>(m_tile) * TILE_ROWS + NJ * w + j0;
        mix_chunk<GT_LD>(p, G_T, D_T, &bars->cfull, &bars->cempty, chunk_it, valid, c, row0, w, j0, lane, amax);
      }
      if (warp == MIX_WARP0) K1_TRACE(ui, 7, K1_CLOCK());   // unit's outputs stored (mix warp 0)
      ++ui;
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

#ifdef EHB_K1_TRACE
extern "C" int ehb_k1_trace_arm(int on) {
  static long long zeros[TRACE_UNITS][TRACE_SLOTS] = {};
  if (on && cudaMemcpyToSymbol(g_k1_trace, zeros, sizeof(zeros)) != cudaSuccess) return 1;
  return cudaMemcpyToSymbol(g_k1_armed, &on, sizeof(int)) != cudaSuccess;
}
extern "C" int ehb_k1_trace_pairs(unsigned long long* out) {   // [128][2] globaltimer ns: start, last unit issued
  return cudaMemcpyFromSymbol(out, g_k1_pair_t, sizeof(unsigned long long) * 128 * 2) != cudaSuccess;
}
extern "C" int ehb_k1_trace_read(long long* out) {   // [TRACE_UNITS][TRACE_SLOTS]
  return cudaMemcpyFromSymbol(out, g_k1_trace, sizeof(long long) * TRACE_UNITS * TRACE_SLOTS) != cudaSuccess;
}
#endif

cudaError_t launch_gcn_hidden_umma_t(const CUtensorMap& tmX, const CUtensorMap& tmW, const HiddenLayerParams& p, int num_sms,
                                     bool pdl, cudaStream_t stream) {
  if (p.C % 128 != 0 || p.n_ntiles != p.C / 128 || p.n_mtiles % 2 != 0) return cudaErrorInvalidValue;
  static bool attr_set = false;
  auto kern = gcn_hidden_umma_t_kernel;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int units = (p.n_mtiles / 2) * p.n_ntiles;
  if (units == 0) return cudaSuccess;
  const int grid = units * 2 < num_sms ? units * 2 : (num_sms / 2) * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, tmX, tmW, p);
}

}  // namespace ehb

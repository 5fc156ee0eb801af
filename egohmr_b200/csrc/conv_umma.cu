// K9 — convolutions of the ResNet-50 image encoder (models/resnet.py:60-150) as GEMMs on tcgen05, fp32-class accuracy.
//
//   Y[M, Cout] = A[M, K] W^T + bias (+ identity) (-> ReLU),  M = images x out-pixels (NHWC rows), K = kh*kw*Cin
//
// A is always a plain row-major fp16 [hi(K) | lo(K)] matrix: a 1x1/stride-1 convolution reads the previous layer's
// activation matrix as it is, everything else (3x3, strided 1x1, the 7x7 stem) reads an im2col matrix written by
// resnet_ops.cu.  BatchNorm (eval) is folded into W and bias on the host.  Same error-compensated scheme as the GCN
// layer kernel (hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator), same CTA-pair / TMA / mbarrier pipeline as
// linear_umma.cu, generalised over the N tile (64 / 128 / 256 output channels per tile, so the 64- and 128-channel
// layers do not pay for 256) and over n-tiles inside the persistent work loop.  The epilogue writes the NEXT layer's
// operand directly: bias, residual add (identity read back from its own hi/lo operand), ReLU, fp16 hi/lo split.
#include <cstdio>
#include <cstdlib>

#include "epilogue.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace ehb {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // 128-byte swizzled rows
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;   // 16 KiB
constexpr int BAR_BYTES = 256;
constexpr int ADDV_BYTES = 2 * 256 * 4;
constexpr int NUM_EPI_WARPS = 8;  // two per TMEM lane quadrant: warp w and w+4 split the tile's columns
constexpr int TMA_WARP = 8, MMA_WARP = 9;
constexpr int NUM_THREADS = 10 * 32;
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 4;

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2 / 2;                 // per CTA of the pair: half of the tile's N rows
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 64 / 48 / 40 KiB
  static constexpr int STAGES = BN == 256 ? 3 : 4;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + BAR_BYTES + ADDV_BYTES + NUM_EPI_WARPS * EPI_SCRATCH_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB of dynamic shared memory");
};

constexpr int MAX_ACC = 8;        // TMEM accumulator ring: 512 columns / tile_n stages (2 / 4 / 8)
struct Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tfull[MAX_ACC];
  uint64_t tempty[MAX_ACC];
  uint32_t tmem_base;
};
static_assert(sizeof(Barriers) <= BAR_BYTES, "barrier block too small");

// ---- optional per-warp timeline (nvcc -DEHB_CONV_TRACE; tools/conv_trace.sh): clock64 stamps of the first units of CTA 0 in
// one chosen launch (ordinal modulo the 53 convolution GEMMs of a ResNet-50 forward).  Compiled out of the product build.
#ifdef EHB_CONV_TRACE
constexpr int TRACE_UNITS = 12, TRACE_SLOTS = 40;
__device__ long long g_trace[TRACE_UNITS][TRACE_SLOTS];
__device__ int g_trace_counter = 0, g_trace_target = -1, g_trace_period = 53;
#define EHB_STAMP(cond, unit_i, slot)                                                                   \
  do {                                                                                                  \
    if (trace_on && (cond) && (unit_i) < TRACE_UNITS) g_trace[(unit_i)][(slot)] = clock64();            \
  } while (0)
#define EHB_NEXT_UNIT() ++ui
#else
#define EHB_NEXT_UNIT() \
  do {                  \
  } while (0)
#define EHB_STAMP(cond, unit_i, slot) \
  do {                                \
  } while (0)
#endif

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ ConvGemmParams p) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, B_BYTES = C::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES);
  uint8_t* scratch_s = smem + STAGES * STAGE_BYTES + BAR_BYTES + ADDV_BYTES;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int n_units = (p.n_mtiles / 2) * p.n_ntiles;
  const int unit0 = blockIdx.x / 2, unit_step = gridDim.x / 2;
  // two operand pairs may accumulate into one TMEM accumulator: conv3(y2) + downsample(x) of a stage's first bottleneck
  // (models/resnet.py:90-94) is ONE launch, the projected identity never exists in memory
  const int KB1 = p.K / BK, KB = KB1 + p.K2 / BK;
  constexpr bool CHUNKED = BN <= 128;   // running sums live in registers: 64 / 32 per epilogue thread
  constexpr int NACC = TMEM_COLS / BN;  // accumulator stages: the MMA warp may run NACC - 1 chunks ahead of the epilogue
  const int kc = (CHUNKED && p.kc > 0 && p.kc < KB) ? p.kc : KB;   // k-blocks per accumulation chunk
#ifdef EHB_CONV_TRACE
  const bool trace_on = blockIdx.x == 0 && (g_trace_counter % g_trace_period) == g_trace_target;
  int ui = 0;   // unit ordinal of this CTA
#endif

  if (warp == TMA_WARP && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    if (p.K2) {
      ptx::prefetch_tensormap(&tmA2);
      ptx::prefetch_tensormap(&tmB2);
    }
  }
  if (warp == MMA_WARP && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < NACC; ++s) {
      ptx::mbar_init(&bars->tfull[s], 1);
      ptx::mbar_init(&bars->tempty[s], 2 * NUM_EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == TMA_WARP) {
    ptx::tmem_alloc_2sm(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  // programmatic dependent launch (see gcn_umma.cu): the set-up above overlaps the tail of the previous kernel in the
  // stream; everything below reads its output or overwrites buffers it may still read
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();

  if (warp == TMA_WARP) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = unit0; u < n_units; u += unit_step) {
        const int tile = (u / p.n_ntiles) * 2 + static_cast<int>(rank);
        const int m_row = tile * BM;
        const int b_row = (u % p.n_ntiles) * BN + static_cast<int>(rank) * (BN / 2);
        // implicit mode: this CTA's tile = images [n0, n0 + nb) x output rows [ho0, ho0 + th)
        const int n0 = p.implicit ? (tile / p.tiles_per_img) * p.nb : 0;
        const int hi0 = p.implicit ? (tile % p.tiles_per_img) * p.th * p.stride - p.pad : 0;
        const int cpk = p.implicit ? p.Cin / BK : 1;
        const uint32_t a_box_bytes = p.implicit ? static_cast<uint32_t>(p.nb * p.th * p.Wo) * BK * 2 : A_BYTES;
        for (int kb = 0; kb < KB; ++kb) {
          EHB_STAMP(kb == 0, ui, 30);
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          EHB_STAMP(kb == 0, ui, 31);
          uint8_t* s = smem + stage * STAGE_BYTES;
          const uint32_t lfull = ptx::mapa(ptx::smem_u32(&bars->full[stage]), 0);
          if (leader) ptx::mbar_arrive_expect_tx(&bars->full[stage], 2 * (2 * a_box_bytes + 2 * B_BYTES));
          if (p.implicit) {
            {
              const int tap = kb / cpk, c = (kb % cpk) * BK;
              const int wi = tap % p.kw - p.pad_w, hi = hi0 + tap / p.kw;
              ptx::tma_load_4d_2sm(s, &tmA, lfull, c, wi, hi, n0);
              // the lo half: the upper channels of the same pixel, or (lo_plane) a second plane of images behind the first
              if (p.lo_plane) ptx::tma_load_4d_2sm(s + A_BYTES, &tmA, lfull, c, wi, hi, n0 + p.lo_plane);
              else ptx::tma_load_4d_2sm(s + A_BYTES, &tmA, lfull, p.Cin + c, wi, hi, n0);
            }
            ptx::tma_load_2d_2sm(s + 2 * A_BYTES, &tmB, lfull, kb * BK, b_row);
            ptx::tma_load_2d_2sm(s + 2 * A_BYTES + B_BYTES, &tmB, lfull, p.K + kb * BK, b_row);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          const bool second = kb >= KB1;
          const CUtensorMap* ta = second ? &tmA2 : &tmA;
          const CUtensorMap* tb = second ? &tmB2 : &tmB;
          const int K = second ? p.K2 : p.K;
          const int kk = (second ? kb - KB1 : kb) * BK;
          ptx::tma_load_2d_2sm(s, ta, lfull, kk, m_row);
          ptx::tma_load_2d_2sm(s + A_BYTES, ta, lfull, K + kk, m_row);
          ptx::tma_load_2d_2sm(s + 2 * A_BYTES, tb, lfull, kk, b_row);
          ptx::tma_load_2d_2sm(s + 2 * A_BYTES + B_BYTES, tb, lfull, K + kk, b_row);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        EHB_STAMP(true, ui, 32);   // this unit's loads are issued
        EHB_NEXT_UNIT();
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(2 * BM, BN);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int u = unit0; u < n_units; u += unit_step) {
        // chunked accumulation: every `kc` k-blocks the accumulator is handed to the epilogue warps, which sum the chunks
        // in fp32 registers (round-to-nearest) — the tensor core's own fp32 accumulation truncates, a bias that grows
        // with the number of MMAs chained into one accumulator (DESIGN.md, K9 numerics)
        for (int kb0 = 0; kb0 < KB; kb0 += kc) {
          EHB_STAMP(lane == 0 && kb0 == 0, ui, 24);
          ptx::mbar_wait_cluster(&bars->tempty[as], aphase ^ 1);
          EHB_STAMP(lane == 0 && kb0 == 0, ui, 25);   // accumulator stage free
          ptx::tc_fence_after_sync();
          const uint32_t tacc = tmem_base + as * BN;
          const int kend = min(KB, kb0 + kc);
          for (int kb = kb0; kb < kend; ++kb) {
            ptx::mbar_wait(&bars->full[stage], phase);
            EHB_STAMP(lane == 0 && kb == 0, ui, 26);    // first operands landed
            ptx::tc_fence_after_sync();
            const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
            // the small cross terms (hi*lo, lo*hi: 2^-11 of the main term) of the whole k-block go first: the tensor
            // core truncates after every MMA by up to one ulp of the CURRENT accumulator, so terms added while the
            // accumulator is still small cost (almost) nothing — only the 4 main MMAs see its full magnitude
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint32_t koff = ks * UMMA_K * 2;
              const uint64_t a_hi = ptx::make_kmajor_desc<128>(sa + koff);
              const uint64_t a_lo = ptx::make_kmajor_desc<128>(sa + A_BYTES + koff);
              const uint64_t b_hi = ptx::make_kmajor_desc<128>(sa + 2 * A_BYTES + koff);
              const uint64_t b_lo = ptx::make_kmajor_desc<128>(sa + 2 * A_BYTES + B_BYTES + koff);
              ptx::umma_f16_2sm_elect(tacc, a_hi, b_lo, idesc, (kb > kb0 || ks != 0) ? 1u : 0u);
              ptx::umma_f16_2sm_elect(tacc, a_lo, b_hi, idesc, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint32_t koff = ks * UMMA_K * 2;
              const uint64_t a_hi = ptx::make_kmajor_desc<128>(sa + koff);
              const uint64_t b_hi = ptx::make_kmajor_desc<128>(sa + 2 * A_BYTES + koff);
              ptx::umma_f16_2sm_elect(tacc, a_hi, b_hi, idesc, 1u);
            }
            ptx::umma_commit_2sm_mc_elect(&bars->empty[stage], 0b11);
            if (kb == kend - 1) ptx::umma_commit_2sm_mc_elect(&bars->tfull[as], 0b11);
            __syncwarp();
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (++as == NACC) {
            as = 0;
            aphase ^= 1;
          }
        }
        EHB_STAMP(lane == 0, ui, 27);   // all MMAs of the unit issued and committed
        EHB_NEXT_UNIT();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread = row
    constexpr int HALF = BN / 2;          // columns per warp
    constexpr int CHUNKS = HALF / 32;
    const int q = warp & 3;               // TMEM lane quadrant
    const int col_half = warp >> 2;
    const float inv_act = 1.f / p.act_scale;
    uint8_t* scratch = scratch_s + warp * EPI_SCRATCH_BYTES;
    int as = 0;
    uint32_t aphase = 0;
    float amax = 0.f;
    const int n_chunks = (KB + kc - 1) / kc;
    float acc[CHUNKED ? HALF : 1];        // fp32 running sum of the first n_chunks - 1 accumulation chunks
    for (int u = unit0; u < n_units; u += unit_step) {
      EHB_STAMP(threadIdx.x == 0, ui, 0);
      const int n_tile = u % p.n_ntiles;
      const int tile = (u / p.n_ntiles) * 2 + static_cast<int>(rank);
      long long tile_first = static_cast<long long>(tile) * BM, tile_rows = BM;
      if (p.implicit) {   // the tile's output pixels are contiguous rows: nb whole images or th full-width rows of one
        const int n0 = (tile / p.tiles_per_img) * p.nb, ho0 = (tile % p.tiles_per_img) * p.th;
        tile_first = (static_cast<long long>(n0) * p.Ho + ho0) * p.Wo;
        const int n_here = min(p.nb, p.n_img - n0), h_here = min(p.th, p.Ho - ho0);
        tile_rows = n_here <= 0 ? 0 : (p.nb > 1 ? static_cast<long long>(n_here) * p.Ho * p.Wo : static_cast<long long>(h_here) * p.Wo);
      }
      tile_rows = max(0LL, min(tile_rows, p.M - tile_first));
      const long long row = tile_first + q * 32 + lane;
      const bool valid = q * 32 + lane < tile_rows;
      const long long row0w = row - lane;                  // first row of this warp's 32
      // bias of this warp's columns: broadcast loads (every lane reads the same 16 bytes), no block-wide staging barrier
      const float4* bias4 = reinterpret_cast<const float4*>(p.bias + n_tile * BN + col_half * HALF);
      // the identity block of the first chunk is requested before the accumulator is awaited, every further chunk's
      // while the previous one is processed: its global latency never sits on the epilogue's critical path
      const long long rows_valid = tile_rows - q * 32;     // rows of this warp's 32 that exist
      uint4 raw_h[4], raw_l[4];
      if (p.res_hl) {
        const int cg0 = n_tile * BN + col_half * HALF;
        warp_issue_rows_64B(raw_h, p.res_hl + row0w * p.out_ld + cg0, p.out_ld, rows_valid, lane);
        warp_issue_rows_64B(raw_l, p.res_hl + row0w * p.out_ld + p.Cout + cg0, p.out_ld, rows_valid, lane);
        // The identity tile of this CTA's NEXT unit is pulled into L2 now, one unit (~10 us) ahead: the epilogue is bound by
        // the latency of its global loads (~32 KB in flight per SM), and an L2 hit costs less than half of an HBM miss.
        // Lane = row; HALF halves of hi and of lo = HALF / 64 128-byte lines each (non-implicit tiles: rows are tile * 128 + r).
        const int un = u + unit_step;
        if (un < n_units && !p.implicit) {
          const long long rn = (static_cast<long long>(un / p.n_ntiles) * 2 + static_cast<int>(rank)) * BM + q * 32 + lane;
          if (rn < p.M) {
            const __half* pn = p.res_hl + rn * p.out_ld + (un % p.n_ntiles) * BN + col_half * HALF;
#pragma unroll
            for (int l = 0; l < (HALF + 63) / 64; ++l) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + l * 64));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + p.Cout + l * 64));
            }
          }
        }
      }
      if constexpr (CHUNKED) {
        if (n_chunks > 1) {
#pragma unroll
          for (int c = 0; c < HALF; ++c) acc[c] = 0.f;
          for (int c0 = 0; c0 + 1 < n_chunks; ++c0) {
            ptx::mbar_wait(&bars->tfull[as], aphase);
            ptx::tc_fence_after_sync();
            const uint32_t tr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + col_half * HALF;
#pragma unroll
            for (int ch = 0; ch < CHUNKS; ++ch) {
              float v[32];
              ptx::tmem_ld_32x32b_x32(tr + ch * 32, v);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) acc[ch * 32 + c] += v[c];
            }
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
              if (leader) ptx::mbar_arrive(&bars->tempty[as]);
              else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->tempty[as]), 0));
            }
            if (++as == NACC) {
              as = 0;
              aphase ^= 1;
            }
          }
        }
      }
      EHB_STAMP(threadIdx.x == 0, ui, 1);
      ptx::mbar_wait(&bars->tfull[as], aphase);
      EHB_STAMP(threadIdx.x == 0, ui, 2);   // final accumulator ready
      ptx::tc_fence_after_sync();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + col_half * HALF;
#pragma unroll(CHUNKED ? CHUNKS : 1)
      for (int ch = 0; ch < CHUNKS; ++ch) {
        float v[32];
        ptx::tmem_ld_32x32b_x32(trow + ch * 32, v);
        ptx::tmem_ld_wait();
        EHB_STAMP(threadIdx.x == 0, ui, 4 + 4 * ch);   // chunk's accumulator in registers
        if (ch == CHUNKS - 1) {
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            if (leader) ptx::mbar_arrive(&bars->tempty[as]);
            else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->tempty[as]), 0));
          }
        }
        if constexpr (CHUNKED) {
          if (n_chunks > 1) {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] += acc[ch * 32 + c];
          }
        }
        const int cg = n_tile * BN + col_half * HALF + ch * 32;   // output channel
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 b = __ldg(bias4 + ch * 8 + c4);
          v[4 * c4 + 0] = v[4 * c4 + 0] * p.acc_scale_inv + b.x;
          v[4 * c4 + 1] = v[4 * c4 + 1] * p.acc_scale_inv + b.y;
          v[4 * c4 + 2] = v[4 * c4 + 2] * p.acc_scale_inv + b.z;
          v[4 * c4 + 3] = v[4 * c4 + 3] * p.acc_scale_inv + b.w;
        }
        // global accesses go through the warp transposes of epilogue.cuh: 8 rows x 64 B per instruction
        if (p.res_hl) {   // identity branch of the bottleneck (models/resnet.py:93-94), stored as its own hi/lo operand
          uint4 rh[4], rl[4];
          warp_finish_rows_64B(scratch, raw_h, rh, lane);
          warp_finish_rows_64B(scratch, raw_l, rl, lane);
          EHB_STAMP(threadIdx.x == 0, ui, 5 + 4 * ch);   // identity block arrived and transposed
          if (ch + 1 < CHUNKS) {
            warp_issue_rows_64B(raw_h, p.res_hl + row0w * p.out_ld + cg + 32, p.out_ld, rows_valid, lane);
            warp_issue_rows_64B(raw_l, p.res_hl + row0w * p.out_ld + p.Cout + cg + 32, p.out_ld, rows_valid, lane);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const __half2* ah = reinterpret_cast<const __half2*>(&rh[g]);
            const __half2* bh = reinterpret_cast<const __half2*>(&rl[g]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fa = __half22float2(ah[e]), fb = __half22float2(bh[e]);
              v[g * 8 + 2 * e] += (fa.x + fb.x) * inv_act;
              v[g * 8 + 2 * e + 1] += (fa.y + fb.y) * inv_act;
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
        }
        if (p.out_f32 && valid) {
          float4* o = reinterpret_cast<float4*>(p.out_f32 + row * p.Cout + cg);
#pragma unroll
          for (int c = 0; c < 8; ++c) o[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
        EHB_STAMP(threadIdx.x == 0, ui, 6 + 4 * ch);     // arithmetic done
        if (p.out_hl) {
          __align__(16) __half2 hi[16], lo[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float s0 = v[2 * c] * p.act_scale, s1 = v[2 * c + 1] * p.act_scale;
            hi[c] = __floats2half2_rn(s0, s1);
            lo[c] = __floats2half2_rn(s0 - __low2float(hi[c]), s1 - __high2float(hi[c]));
            if (valid) amax = fmaxf(amax, fmaxf(fabsf(s0), fabsf(s1)));
          }
          warp_store_rows_64B(scratch, *reinterpret_cast<const uint4(*)[4]>(hi), p.out_hl + row0w * p.out_ld + cg, p.out_ld,
                              rows_valid, lane);
          warp_store_rows_64B(scratch, *reinterpret_cast<const uint4(*)[4]>(lo), p.out_hl + row0w * p.out_ld + p.Cout + cg,
                              p.out_ld, rows_valid, lane);
        }
        EHB_STAMP(threadIdx.x == 0, ui, 7 + 4 * ch);     // stores issued
      }
      EHB_NEXT_UNIT();
      if (++as == NACC) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (!(amax <= 65504.f)) atomicExch(p.overflow_flag, 1);
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == TMA_WARP) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
#ifdef EHB_CONV_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) g_trace_counter = g_trace_counter + 1;   // launches are stream-ordered
#endif
}

#ifdef EHB_CONV_TRACE
}  // namespace
// trace control (tools/conv_trace.sh builds a library with -DEHB_CONV_TRACE and calls these through ctypes)
extern "C" int ehb_conv_trace_arm(int target, int period) {
  int zero = 0;
  static long long zeros[TRACE_UNITS][TRACE_SLOTS] = {};
  if (cudaMemcpyToSymbol(g_trace_counter, &zero, sizeof(int)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(g_trace_target, &target, sizeof(int)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(g_trace_period, &period, sizeof(int)) != cudaSuccess) return 1;
  return cudaMemcpyToSymbol(g_trace, zeros, sizeof(zeros)) != cudaSuccess;
}
extern "C" int ehb_conv_trace_read(long long* out) {   // [TRACE_UNITS][TRACE_SLOTS]
  return cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * TRACE_UNITS * TRACE_SLOTS) != cudaSuccess;
}
namespace {
#endif

template <int BN>
cudaError_t launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2, const CUtensorMap& tmB2,
                      const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int units = (p.n_mtiles / 2) * p.n_ntiles;
  if (units == 0) return cudaSuccess;
  const int grid = units * 2 < num_sms ? units * 2 : (num_sms / 2) * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg<BN>::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BN>, tmA, tmB, tmA2, tmB2, p);
}

}  // namespace

// N tile (256 / 128 / 64 output channels) of one layer: the one with the lowest estimated time = rounds of work units
// over the CTA pairs x relative cost of a unit (narrower tiles re-read A and run the MMA less efficiently).  Small-M
// layers (the 14 x 14 and 7 x 7 stages) trade MMA width for parallelism.
int conv_gemm_tile_n(int cout, long long rows, int num_sms, int max_bn) {
  static float cost[3] = {1.0f, 0.6f, 0.4f};   // swept on B200: 2.83 ms for ResNet-50 x 64 images (always-largest tile: 2.98)
  static bool init = false;
  if (!init) {   // bring-up override: EHB_CONV_COST="c256,c128,c64"
    if (const char* e = getenv("EHB_CONV_COST")) sscanf(e, "%f,%f,%f", &cost[0], &cost[1], &cost[2]);
    init = true;
  }
  const long long m_units = (rows + 255) / 256;
  const int pairs = num_sms / 2;
  const int bns[3] = {256, 128, 64};
  int best = 0;
  float best_t = 1e30f;
  for (int i = 0; i < 3; ++i) {
    if (cout % bns[i] || bns[i] > max_bn) continue;
    const long long units = m_units * (cout / bns[i]);
    const float t = static_cast<float>((units + pairs - 1) / pairs) * cost[i];
    if (t < best_t) {
      best_t = t;
      best = bns[i];
    }
  }
  return best;
}

cudaError_t launch_conv_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2, const CUtensorMap& tmB2,
                             const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  if (p.K <= 0 || p.K % BK || p.K2 < 0 || p.K2 % BK || p.n_mtiles % 2 || p.Cout % 64 || p.n_ntiles <= 0 || p.Cout % p.n_ntiles)
    return cudaErrorInvalidValue;
  const int bn = p.Cout / p.n_ntiles;
  if (bn != 256 && bn != 128 && bn != 64) return cudaErrorInvalidValue;
  if (bn == 256) return launch_bn<256>(tmA, tmB, tmA2, tmB2, p, num_sms, stream);
  if (bn == 128) return launch_bn<128>(tmA, tmB, tmA2, tmB2, p, num_sms, stream);
  return launch_bn<64>(tmA, tmB, tmA2, tmB2, p, num_sms, stream);
}

}  // namespace ehb

// K1 — one hidden ModulatedGraphConv (+ BatchNorm1d(eval) + ReLU [+ residual]) of the EgoHMR denoiser as a
// persistent, warp-specialised tcgen05 kernel for sm_100a.
//
// Reference semantics (models/egohmr/modulated_gcn/modulated_gcn_conv.py:38-50, modulated_gcn.py:21-28,38-42):
//     h0 = X W[0];  h1 = X W[1];  A = sym(adj + adj2)
//     Y  = (A o I)(M o h0) + (A o (1-I))(M o h1) + bias ;  out = relu(BN(Y)) (+ residual)
//
// GEMM part.  X is [rows, K] with rows = (body,pass) slots x 24 joints, K = C = hid.  W[0]|W[1] are concatenated so
// that one 256-wide N tile holds the h0 and the h1 columns of the same 128 channels.  fp32-class accuracy on the
// fp16 tensor pipe comes from an error-compensated split: every operand v (pre-scaled by a power of two) is stored as
// hi = fp16(v), lo = fp16(v - hi) and the kernel accumulates  hi*hi + hi*lo + lo*hi  into ONE fp32 TMEM accumulator
// (3 tcgen05.mma kind::f16 per 16-wide k slice).  The dropped lo*lo term and the fp16 rounding of lo are ~2^-22
// relative, i.e. fp32 SGEMM territory.
//
// Pipeline.  warp 4: TMA producer (A_hi, A_lo, B_hi, B_lo tiles, 64B swizzle, 4-stage mbarrier ring);
// warp 5: single-thread tcgen05.mma issuer, fp32 accumulator double-buffered in TMEM (2 x 256 columns);
// warps 0-3: tcgen05.ld the accumulator (one TMEM lane quadrant each), apply the modulation M and stage
// (M o h1), diag(A)(M o h0) transposed in shared memory; warps 6-10: one (body,pass) slot each, lane = channel:
// 24x24 joint mix with the adjacency taken straight from the kernel-parameter constant bank, BN scale/shift, ReLU,
// residual, then write the next layer's operand (hi|lo fp16) and/or the fp32 block-boundary activations.
#include "kernels.cuh"
#include "ptx.cuh"

namespace ehb {
namespace {

constexpr int BM = 128;           // rows per tile (5 slots x 24 joints + 8 pad)
constexpr int BN = 256;           // 128 channels of h0 | the same 128 channels of h1
constexpr int BK = 32;            // fp16 elements per k block = 64 bytes = one SWIZZLE_64B span
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // hi+lo of both operands = 48 KiB
constexpr int CHUNK = 32;         // channels per epilogue hand-off
constexpr int GT_LD = 132;        // padded row length (floats) of the transposed staging tiles
constexpr int EPI_BYTES = 2 * CHUNK * GT_LD * 4;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
// warp roles: 0-3 tcgen05.ld (warp index == TMEM lane quadrant), 4 TMA producer (+ TMEM alloc/dealloc),
// 5 MMA issuer, 6-10 joint-mix/store.  11 warps keep the per-thread register budget at 168.
constexpr int NUM_WARPS = 11;
constexpr int NUM_THREADS = NUM_WARPS * 32;
constexpr int LD_WARP0 = 0, TMA_WARP = 4, MMA_WARP = 5, MIX_WARP0 = 6;
constexpr int TMEM_COLS = 512;

static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB of dynamic shared memory");

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

__global__ void __launch_bounds__(NUM_THREADS, 1)
gcn_hidden_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
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
  const int total_tiles = p.n_mtiles * p.n_ntiles;

  if (warp == TMA_WARP && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == MMA_WARP && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bars->tfull[s], 1);
      ptx::mbar_init(&bars->tempty[s], 4);   // one elected lane of each tcgen05.ld warp
    }
    ptx::mbar_init(&bars->cfull, 4 * 32);    // every thread of the tcgen05.ld warps
    ptx::mbar_init(&bars->cempty, 5 * 32);   // every thread of the mix warps
    ptx::fence_mbar_init();
  }
  if (warp == TMA_WARP) {
    ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == TMA_WARP) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_ntiles;
        const int n_tile = tile % p.n_ntiles;
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* s = stage_base + stage * STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&bars->full[stage], STAGE_BYTES);
          ptx::tma_load_2d(s, &tmA, &bars->full[stage], kb * BK, m_tile * BM);
          ptx::tma_load_2d(s + A_BYTES, &tmA, &bars->full[stage], p.C + kb * BK, m_tile * BM);
          ptx::tma_load_2d(s + 2 * A_BYTES, &tmB, &bars->full[stage], kb * BK, n_tile * BN);
          ptx::tma_load_2d(s + 2 * A_BYTES + B_BYTES, &tmB, &bars->full[stage], p.C + kb * BK, n_tile * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::make_idesc_f16_f32(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&bars->tempty[as], aphase ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t tacc = tmem_base + as * BN;
      for (int kb = 0; kb < KB; ++kb) {
        ptx::mbar_wait(&bars->full[stage], phase);
        ptx::tc_fence_after_sync();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(stage_base + stage * STAGE_BYTES);
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            const uint32_t koff = ks * UMMA_K * 2;
            const uint64_t a_hi = ptx::make_kmajor_desc<64>(sa + koff);
            const uint64_t a_lo = ptx::make_kmajor_desc<64>(sa + A_BYTES + koff);
            const uint64_t b_hi = ptx::make_kmajor_desc<64>(sa + 2 * A_BYTES + koff);
            const uint64_t b_lo = ptx::make_kmajor_desc<64>(sa + 2 * A_BYTES + B_BYTES + koff);
            ptx::umma_f16(tacc, a_hi, b_hi, idesc, (kb | ks) != 0 ? 1u : 0u);
            ptx::umma_f16(tacc, a_hi, b_lo, idesc, 1u);
            ptx::umma_f16(tacc, a_lo, b_hi, idesc, 1u);
          }
          ptx::umma_commit(&bars->empty[stage]);
          if (kb == KB - 1) ptx::umma_commit(&bars->tfull[as]);
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
  } else if (warp < LD_WARP0 + 4) {
    // ------------------------------------------------------------------ TMEM -> modulate -> shared (transposed)
    const int q = warp - LD_WARP0;  // == warp % 4: the TMEM lane quadrant this warp may read
    const int r = q * 32 + lane;    // tile row
    const bool valid_row = r < SLOTS_PER_TILE * NJ;
    const int j = r % NJ;
    const float adiag = p.adj.diag[j];
    int as = 0;
    uint32_t aphase = 0;
    uint32_t chunk_it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_ntiles;
      ptx::mbar_wait(&bars->tfull[as], aphase);
      ptx::tc_fence_after_sync();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int ch = 0; ch < BN / 2 / CHUNK; ++ch, ++chunk_it) {
        float h0[32], h1[32], m[32];
        ptx::tmem_ld_32x32b_x32(trow + ch * CHUNK, h0);
        ptx::tmem_ld_32x32b_x32(trow + BN / 2 + ch * CHUNK, h1);
        const float4* mp = reinterpret_cast<const float4*>(p.mod + static_cast<size_t>(j) * p.C + n_tile * (BN / 2) +
                                                            ch * CHUNK);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const float4 t = __ldg(mp + v);
          m[4 * v + 0] = t.x;
          m[4 * v + 1] = t.y;
          m[4 * v + 2] = t.z;
          m[4 * v + 3] = t.w;
        }
        ptx::tmem_ld_wait();
        if (ch == BN / 2 / CHUNK - 1) {
          // accumulator fully drained: hand the TMEM stage back to the MMA warp
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars->tempty[as]);
        }
        ptx::mbar_wait(&bars->cempty, (chunk_it & 1) ^ 1);
        if (valid_row) {
#pragma unroll
          for (int c = 0; c < CHUNK; ++c) {
            G_T[c * GT_LD + r] = m[c] * h1[c];
            D_T[c * GT_LD + r] = adiag * (m[c] * h0[c]);
          }
        }
        ptx::mbar_arrive(&bars->cfull);
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else if (warp >= MIX_WARP0) {
    // ------------------------------------------------------------------ joint mix + BN + ReLU (+res) + store
    const int w = warp - MIX_WARP0;  // slot within the tile
    uint32_t chunk_it = 0;
    float amax = 0.f;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_ntiles;
      const int n_tile = tile % p.n_ntiles;
      const bool valid = (m_tile * SLOTS_PER_TILE + w) < p.n_slots;
#pragma unroll 1
      for (int ch = 0; ch < BN / 2 / CHUNK; ++ch, ++chunk_it) {
        float g[NJ], y[NJ], rsd[NJ];
        const int c = n_tile * (BN / 2) + ch * CHUNK + lane;
        const size_t row0 = static_cast<size_t>(m_tile) * BM + NJ * w;
        // residual rows are fetched before the hand-off wait: independent loads in flight, latency hidden
        if (p.add_res && valid) {
#pragma unroll
          for (int jj = 0; jj < NJ; ++jj) rsd[jj] = __ldcg(p.res + (row0 + jj) * p.C + c);
        } else {
#pragma unroll
          for (int jj = 0; jj < NJ; ++jj) rsd[jj] = 0.f;
        }
        ptx::mbar_wait(&bars->cfull, chunk_it & 1);
        {
          const float4* gp = reinterpret_cast<const float4*>(G_T + lane * GT_LD + NJ * w);
          const float4* dp = reinterpret_cast<const float4*>(D_T + lane * GT_LD + NJ * w);
#pragma unroll
          for (int v = 0; v < NJ / 4; ++v) {
            const float4 a = gp[v];
            const float4 b = dp[v];
            g[4 * v + 0] = a.x; g[4 * v + 1] = a.y; g[4 * v + 2] = a.z; g[4 * v + 3] = a.w;
            y[4 * v + 0] = b.x; y[4 * v + 1] = b.y; y[4 * v + 2] = b.z; y[4 * v + 3] = b.w;
          }
        }
        ptx::mbar_arrive(&bars->cempty);
        if (!valid) continue;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          float acc = y[jj];
#pragma unroll
          for (int i = 0; i < NJ; ++i) acc = fmaf(p.adj.off[jj][i], g[i], acc);
          y[jj] = acc;
        }
        const float sc = __ldg(p.bn_scale + c);
        const float sh = __ldg(p.bn_shift + c);
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          const size_t row = row0 + jj;
          const float v = fmaxf(fmaf(y[jj], sc, sh), 0.f) + rsd[jj];
          if (p.write_f32) p.res[row * p.C + c] = v;
          if (p.write_hl) {
            const float sv = v * p.act_scale;
            const __half hi = __float2half_rn(sv);
            const __half lo = __float2half_rn(sv - __half2float(hi));
            p.out_hl[row * (2 * p.C) + c] = hi;
            p.out_hl[row * (2 * p.C) + p.C + c] = lo;
            amax = fmaxf(amax, fabsf(sv));
          }
        }
      }
    }
    if (!(amax <= 65504.f)) atomicExch(p.overflow_flag, 1);  // also catches NaN
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == TMA_WARP) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

size_t gcn_hidden_umma_smem_bytes() { return SMEM_BYTES; }

cudaError_t launch_gcn_hidden_umma(const CUtensorMap& tmA, const CUtensorMap& tmB, const HiddenLayerParams& p,
                                   int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(gcn_hidden_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (p.C % 128 != 0 || p.n_ntiles != p.C / 128) return cudaErrorInvalidValue;
  const int total = p.n_mtiles * p.n_ntiles;
  if (total == 0) return cudaSuccess;
  const int grid = total < num_sms ? total : num_sms;
  gcn_hidden_umma_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
  return cudaGetLastError();
}

}  // namespace ehb

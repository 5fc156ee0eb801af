// K7 — point-wise linear layers of the ResPointNet scene encoder (models/respointnet.py:33-59, 88-97) on tcgen05:
//     Y[M,256] = A1[M,K1] W1^T (+ A2[M,K2] W2^T) (+ bias) (+ rowvec[cloud(row)])
// with M = clouds x points, and a fused epilogue that writes what the NEXT layer consumes instead of fp32 activations:
// the fp16 hi/lo operand of relu(Y) and/or of Y, and the per-cloud column max (the PointNet pooling, :38,42,46,55).
// Same fp16x3 error-compensated scheme, CTA pairs (cta_group::2) and TMA/mbarrier pipeline as gcn_umma.cu; two operand
// pairs may accumulate into one TMEM accumulator, which is how a ResnetBlockFC's `shortcut(x) + fc_1(relu(h))` becomes
// ONE launch with no fp32 round trip.  Step-invariant: runs once per image batch.
#include "epilogue.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace ehb {
namespace {

constexpr int BM = 128;
constexpr int BN = 256;           // ResPointNet hidden width
constexpr int BK = 64;            // 128-byte swizzled rows
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;            // 16 KiB
constexpr int B_BYTES = BN * BK * 2 / 2;        // per CTA of the pair: half of the N rows
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 64 KiB
constexpr int STAGES = 3;
constexpr int BAR_BYTES = 256;
constexpr int ADDV_BYTES = (2 + 3) * 256 * 4;   // per-tile additive row (bias + per-cloud row), double-buffered; + ptw[3][256]
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + BAR_BYTES + ADDV_BYTES + 8 * EPI_SCRATCH_BYTES;
constexpr int NUM_EPI_WARPS = 8;  // two per TMEM lane quadrant: warp w and w+4 split the 256 columns
constexpr int TMA_WARP = 8, MMA_WARP = 9;
constexpr int NUM_THREADS = 10 * 32;
constexpr int TMEM_COLS = 512;

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Barriers) <= BAR_BYTES, "barrier block too small");

__device__ __forceinline__ int encode_max(float v) {
  const int i = __float_as_int(v);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;   // order-preserving float -> int
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
linear_umma_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                   const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                   const __grid_constant__ LinearParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES);
  float* addv_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + BAR_BYTES);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int n_units = p.n_mtiles / 2;
  const int unit0 = blockIdx.x / 2, unit_step = gridDim.x / 2;
  const int KB1 = p.K1 / BK, KB2 = p.K2 / BK;

  if (warp == TMA_WARP && lane == 0) {
    ptx::prefetch_tensormap(&tmA1);
    ptx::prefetch_tensormap(&tmB1);
    if (KB2) {
      ptx::prefetch_tensormap(&tmA2);
      ptx::prefetch_tensormap(&tmB2);
    }
  }
  if (warp == MMA_WARP && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
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
        const int m_row = (u * 2 + static_cast<int>(rank)) * BM;
        const int b_row = static_cast<int>(rank) * (BN / 2);
        for (int kb = 0; kb < KB1 + KB2; ++kb) {
          const bool second = kb >= KB1;
          const CUtensorMap* ta = second ? &tmA2 : &tmA1;
          const CUtensorMap* tb = second ? &tmB2 : &tmB1;
          const int K = second ? p.K2 : p.K1;
          const int kk = (second ? kb - KB1 : kb) * BK;
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* s = smem + stage * STAGE_BYTES;
          const uint32_t lfull = ptx::mapa(ptx::smem_u32(&bars->full[stage]), 0);
          if (leader) ptx::mbar_arrive_expect_tx(&bars->full[stage], 2 * STAGE_BYTES);
          ptx::tma_load_2d_2sm(s, ta, lfull, kk, m_row);
          ptx::tma_load_2d_2sm(s + A_BYTES, ta, lfull, K + kk, m_row);
          ptx::tma_load_2d_2sm(s + 2 * A_BYTES, tb, lfull, kk, b_row);
          ptx::tma_load_2d_2sm(s + 2 * A_BYTES + B_BYTES, tb, lfull, K + kk, b_row);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(2 * BM, BN);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int u = unit0; u < n_units; u += unit_step) {
        ptx::mbar_wait_cluster(&bars->tempty[as], aphase ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < KB1 + KB2; ++kb) {
          ptx::mbar_wait(&bars->full[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            const uint32_t koff = ks * UMMA_K * 2;
            const uint64_t a_hi = ptx::make_kmajor_desc<128>(sa + koff);
            const uint64_t a_lo = ptx::make_kmajor_desc<128>(sa + A_BYTES + koff);
            const uint64_t b_hi = ptx::make_kmajor_desc<128>(sa + 2 * A_BYTES + koff);
            const uint64_t b_lo = ptx::make_kmajor_desc<128>(sa + 2 * A_BYTES + B_BYTES + koff);
            ptx::umma_f16_2sm_elect(tacc, a_hi, b_hi, idesc, (kb | ks) != 0 ? 1u : 0u);
            ptx::umma_f16_2sm_elect(tacc, a_hi, b_lo, idesc, 1u);
            ptx::umma_f16_2sm_elect(tacc, a_lo, b_hi, idesc, 1u);
          }
          ptx::umma_commit_2sm_mc_elect(&bars->empty[stage], 0b11);
          if (kb == KB1 + KB2 - 1) ptx::umma_commit_2sm_mc_elect(&bars->tfull[as], 0b11);
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
  } else {
    // ------------------------------------------------------------------ epilogue: thread = row, 4 x 32 columns each
    const int q = warp & 3;             // TMEM lane quadrant
    const int col_half = warp >> 2;     // columns [128*col_half, +128)
    const int et = threadIdx.x;         // 0..255 within the epilogue warps
    uint8_t* scratch = smem + STAGES * STAGE_BYTES + BAR_BYTES + ADDV_BYTES + warp * EPI_SCRATCH_BYTES;
    int as = 0;
    uint32_t aphase = 0;
    float amax = 0.f;
    float* ptw_s = addv_s + 2 * 256;
    if (p.pts) {   // staged once per launch; the first per-tile barrier below orders it before any use
      for (int e = et; e < 3 * 256; e += 256) ptw_s[e] = __ldg(p.ptw + e);
    }
    for (int u = unit0; u < n_units; u += unit_step) {
      const long long tile_first = static_cast<long long>(u * 2 + static_cast<int>(rank)) * BM;
      const long long row = tile_first + q * 32 + lane;
      const bool valid = row < p.M;
      const int cloud = valid ? static_cast<int>(row / p.pts_per_cloud) : 0;
      // pooling fast path: all 32 rows of this warp are valid and in one cloud
      const long long row_first = row - lane;
      const bool warp_one_cloud = (row_first + 31 < p.M) && (row_first / p.pts_per_cloud == (row_first + 31) / p.pts_per_cloud);
      // The additive row (bias + the cloud's row vector) is the same for every row of a tile that lies inside one
      // cloud: stage it in shared memory once per tile instead of 2 x 128 global loads per thread.
      const long long tile_last = min(tile_first + BM, p.M) - 1;
      const bool tile_one_cloud = tile_last >= tile_first && (tile_first / p.pts_per_cloud == tile_last / p.pts_per_cloud);
      float* addv = addv_s + as * 256;
      {
        float a = p.bias ? __ldg(p.bias + et) : 0.f;
        if (p.rowvec && tile_one_cloud) a += __ldg(p.rowvec + static_cast<size_t>(tile_first / p.pts_per_cloud) * BN + et);
        addv[et] = a;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const bool row_from_global = p.rowvec && !tile_one_cloud;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (p.pts && valid) {
        px = __ldg(p.pts + row * 3);
        py = __ldg(p.pts + row * 3 + 1);
        pz = __ldg(p.pts + row * 3 + 2);
      }
      ptx::mbar_wait(&bars->tfull[as], aphase);
      ptx::tc_fence_after_sync();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + col_half * 128;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        float v[32];
        ptx::tmem_ld_32x32b_x32(trow + ch * 32, v);
        ptx::tmem_ld_wait();
        if (ch == 3) {
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            if (leader) ptx::mbar_arrive(&bars->tempty[as]);
            else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->tempty[as]), 0));
          }
        }
        const int c0 = col_half * 128 + ch * 32;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float y = v[c] * p.acc_scale_inv + addv[c0 + c];
          if (row_from_global) y += __ldg(p.rowvec + static_cast<size_t>(cloud) * BN + c0 + c);
          if (p.pts) y += fmaf(pz, ptw_s[512 + c0 + c], fmaf(py, ptw_s[256 + c0 + c], px * ptw_s[c0 + c]));
          v[c] = y;
        }
        if (p.out_f32 && valid) {
          float4* o = reinterpret_cast<float4*>(p.out_f32 + row * BN + c0);
#pragma unroll
          for (int c = 0; c < 8; ++c) o[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
        if (p.out_hl || p.out_hl_relu) {
          // 32 consecutive channels of every row: 64 B of hi and 64 B of lo per destination, written through the warp
          // transposes of epilogue.cuh (8 rows x 64 B per store instruction instead of 32 rows x 16 B)
          const long long rows_valid = p.M - row_first;
          __align__(16) __half2 hi[16], lo[16];
          if (p.out_hl) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float s0 = v[2 * c] * p.act_scale, s1 = v[2 * c + 1] * p.act_scale;
              hi[c] = __floats2half2_rn(s0, s1);
              lo[c] = __floats2half2_rn(s0 - __low2float(hi[c]), s1 - __high2float(hi[c]));
              if (valid) amax = fmaxf(amax, fmaxf(fabsf(s0), fabsf(s1)));
            }
            warp_store_rows_64B(scratch, *reinterpret_cast<const uint4(*)[4]>(hi), p.out_hl + row_first * (2 * BN) + c0,
                                2 * BN, rows_valid, lane);
            warp_store_rows_64B(scratch, *reinterpret_cast<const uint4(*)[4]>(lo), p.out_hl + row_first * (2 * BN) + BN + c0,
                                2 * BN, rows_valid, lane);
            if (p.out_hl_relu) {
              // relu(y) splits into the same (hi, lo) when y > 0 and into (0, 0) otherwise: mask the packed halves
              const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const __half2 pos = __hgt2(hi[c], zero);          // 1.0 where hi > 0
                hi[c] = __hmul2(hi[c], pos);
                lo[c] = __hmul2(lo[c], pos);
              }
              warp_store_rows_64B(scratch, *reinterpret_cast<const uint4(*)[4]>(hi), p.out_hl_relu + row_first * (2 * BN) + c0,
                                  2 * BN, rows_valid, lane);
              warp_store_rows_64B(scratch, *reinterpret_cast<const uint4(*)[4]>(lo),
                                  p.out_hl_relu + row_first * (2 * BN) + BN + c0, 2 * BN, rows_valid, lane);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float s0 = fmaxf(v[2 * c], 0.f) * p.act_scale, s1 = fmaxf(v[2 * c + 1], 0.f) * p.act_scale;
              hi[c] = __floats2half2_rn(s0, s1);
              lo[c] = __floats2half2_rn(s0 - __low2float(hi[c]), s1 - __high2float(hi[c]));
              if (valid) amax = fmaxf(amax, fmaxf(s0, s1));
            }
            warp_store_rows_64B(scratch, *reinterpret_cast<const uint4(*)[4]>(hi), p.out_hl_relu + row_first * (2 * BN) + c0,
                                2 * BN, rows_valid, lane);
            warp_store_rows_64B(scratch, *reinterpret_cast<const uint4(*)[4]>(lo),
                                p.out_hl_relu + row_first * (2 * BN) + BN + c0, 2 * BN, rows_valid, lane);
          }
        }
        if (p.pool) {
          if (warp_one_cloud) {
            // column max over the warp's 32 rows as a butterfly reduce-scatter: 31 shuffles for 32 columns instead of
            // 160; afterwards lane l holds the maximum of column l (destroys v, hence last)
#pragma unroll
            for (int s = 16, n = 32; s >= 1; s >>= 1, n >>= 1) {
              const bool up = (lane & s) != 0;
#pragma unroll
              for (int i = 0; i < n / 2; ++i) {
                const float send = up ? v[i] : v[i + n / 2];
                const float keep = up ? v[i + n / 2] : v[i];
                v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, s));
              }
            }
            atomicMax(p.pool + static_cast<size_t>(cloud) * BN + c0 + lane, encode_max(v[0]));
          } else if (valid) {
#pragma unroll
            for (int c = 0; c < 32; ++c) atomicMax(p.pool + static_cast<size_t>(cloud) * BN + c0 + c, encode_max(v[c]));
          }
        }
      }
      if (++as == 2) {
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
}

// fc_pos_0 (3 -> 2*hidden, models/respointnet.py:21,35): K = 3 is not a tensor-core shape; this writes the fp16 hi/lo
// operand block_0's fc_0 consumes, relu(net).  (net itself, the shortcut's input, is optional: block_0's shortcut is
// folded into a K = 3 epilogue term of the fc_1 GEMM, see ehb_pointnet_load.)
__global__ void __launch_bounds__(256) pointnet_pos_kernel(const float* __restrict__ pts, const float* __restrict__ w /*[3][C]*/,
                                                           const float* __restrict__ b, __half* __restrict__ out_hl,
                                                           __half* __restrict__ out_hl_relu, long long M, int C,
                                                           float act_scale, int* overflow_flag) {
  // thread = 8 channels, kept in registers with their weights while the thread walks POS_ROWS rows; the C/8 threads of
  // a row write C * 2 contiguous bytes per operand half
  constexpr int POS_ROWS = 16;
  const int groups = C / 8;                         // channel groups per row
  const int rows_per_block = 256 / groups;          // row lanes of this block
  const int cgp = threadIdx.x % groups, rl = threadIdx.x / groups;
  if (rl >= rows_per_block) return;
  const int c0 = cgp * 8;
  float wx[8], wy[8], wz[8], bb[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    wx[c] = w[c0 + c];
    wy[c] = w[C + c0 + c];
    wz[c] = w[2 * C + c0 + c];
    bb[c] = b[c0 + c];
  }
  float amax = 0.f;
  const long long row_begin = static_cast<long long>(blockIdx.x) * rows_per_block * POS_ROWS;
  for (int i = 0; i < POS_ROWS; ++i) {
    const long long row = row_begin + static_cast<long long>(i) * rows_per_block + rl;
    if (row >= M) break;
    const float x = pts[row * 3 + 0], y = pts[row * 3 + 1], z = pts[row * 3 + 2];
    __align__(16) __half hi[8], lo[8], rhi[8], rlo[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float v = fmaf(z, wz[c], fmaf(y, wy[c], fmaf(x, wx[c], bb[c])));
      const float sv = v * act_scale, rv = fmaxf(v, 0.f) * act_scale;
      hi[c] = __float2half_rn(sv);
      lo[c] = __float2half_rn(sv - __half2float(hi[c]));
      rhi[c] = __float2half_rn(rv);
      rlo[c] = __float2half_rn(rv - __half2float(rhi[c]));
      amax = fmaxf(amax, fabsf(sv));
    }
    if (out_hl) {
      *reinterpret_cast<uint4*>(out_hl + row * (2 * C) + c0) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(out_hl + row * (2 * C) + C + c0) = *reinterpret_cast<const uint4*>(lo);
    }
    *reinterpret_cast<uint4*>(out_hl_relu + row * (2 * C) + c0) = *reinterpret_cast<const uint4*>(rhi);
    *reinterpret_cast<uint4*>(out_hl_relu + row * (2 * C) + C + c0) = *reinterpret_cast<const uint4*>(rlo);
  }
  if (!(amax <= 65504.f)) atomicExch(overflow_flag, 1);
}

__global__ void pool_init_kernel(int* pool, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pool[i] = encode_max(-INFINITY);
}
__global__ void pool_decode_kernel(const int* pool, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int e = pool[i];
    out[i] = __int_as_float(e >= 0 ? e : e ^ 0x7FFFFFFF);
  }
}

}  // namespace

cudaError_t launch_linear_umma(const CUtensorMap& tmA1, const CUtensorMap& tmB1, const CUtensorMap& tmA2,
                               const CUtensorMap& tmB2, const LinearParams& p, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(linear_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (p.K1 <= 0 || p.K1 % BK || p.K2 % BK || p.n_mtiles % 2) return cudaErrorInvalidValue;
  const int units = p.n_mtiles / 2;
  if (units == 0) return cudaSuccess;
  const int grid = units * 2 < num_sms ? units * 2 : (num_sms / 2) * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
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
  return cudaLaunchKernelEx(&cfg, linear_umma_kernel, tmA1, tmB1, tmA2, tmB2, p);
}

cudaError_t launch_pointnet_pos(const float* pts, const float* w, const float* b, __half* out_hl, __half* out_hl_relu,
                                long long M, int C, float act_scale, int* overflow_flag, cudaStream_t stream) {
  if (M <= 0) return cudaSuccess;
  if (C % 8 != 0 || C / 8 > 256) return cudaErrorInvalidValue;
  const long long rows_per_block = static_cast<long long>(256 / (C / 8)) * 16;
  pointnet_pos_kernel<<<static_cast<unsigned>((M + rows_per_block - 1) / rows_per_block), 256, 0, stream>>>(
      pts, w, b, out_hl, out_hl_relu, M, C, act_scale, overflow_flag);
  return cudaGetLastError();
}

cudaError_t launch_pool_init(int* pool, int n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  pool_init_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pool, n);
  return cudaGetLastError();
}
cudaError_t launch_pool_decode(const int* pool, float* out, int n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  pool_decode_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pool, out, n);
  return cudaGetLastError();
}

}  // namespace ehb

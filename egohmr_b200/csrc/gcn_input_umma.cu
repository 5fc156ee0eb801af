// K2 on the tensor pipe: the input ModulatedGraphConv layer with its 24x24 joint mix as a tcgen05 GEMM.
//
// Reference: EgoHMR.forward builds feat = [img*vis | scene | transl | cam | Linear6->512(x_t) | temb] (egohmr.py:190-236)
// and feeds it to gconv_input (modulated_gcn.py:99-101); ModulatedGraphConv.forward (modulated_graph_conv.py) is
//   y[j] = adj[j][j] * (M[j] . h0[j]) + sum_{i != j} adj[j][i] * (M[i] . h1[i])        per channel.
// The per-joint products g[i] = M[i] . h1[i] are cheap (the feature splits into per-image, per-step and per-joint terms,
// SURVEY 7.2), but the off-diagonal mix is 24 x 24 FMAs per (slot, channel): 53 % of the FFMA kernel's instructions
// (profiles/r02_ncu_full_k2_k3.txt: issue-active 80 %, i.e. FFMA-bound, at 1.9x its HBM time).  Here the mix is
//   D[channel][j] = sum_i G[channel][i] * off[j][i]              (M = 128 channels, N = 32, K = 32; joints 24..31 are zero)
// on the fp16 tensor pipe with fp32-class operands: G and off are split hi + lo in fp16 and the three significant
// products are accumulated in TMEM (cross terms first, so the large hi.hi products are added last and the accumulator's
// truncation, 0.5 ulp per MMA, applies to two instructions only).  The kernel is then bound by its 2 x 98.3 KB of output
// per slot.  One CTA owns a 128-channel chunk and walks the slots; thread = channel = accumulator lane.
#include "kernels.cuh"
#include "ptx.cuh"

namespace ehb {
namespace {

constexpr int K2T_THREADS = 128;
constexpr int K2T_CTAS_PER_SM = 4;
constexpr uint32_t K2T_TMEM_COLS = 32;
constexpr float K2T_OFF_SCALE = 1024.f;   // adjacency entries are O(1): hi/lo fp16 of off * 2^10 keep 22 bits
constexpr uint32_t K2T_A_BYTES = 128 * 64;          // [128 channels][32 joints] fp16, 64-byte swizzled rows
constexpr uint32_t K2T_B_BYTES = 32 * 64;           // [32 output joints][32 input joints] fp16
constexpr uint32_t K2T_OFF_MOD = 2 * K2T_A_BYTES + 2 * K2T_B_BYTES;   // 20480
constexpr uint32_t K2T_OFF_XS = K2T_OFF_MOD + NJ * 128 * 4;            // 32768
constexpr uint32_t K2T_OFF_BAR = K2T_OFF_XS + (XDIM + NJ + 8) * 4;     // 8-byte aligned
constexpr uint32_t K2T_SMEM = K2T_OFF_BAR + 16 + 1024;                 // + alignment slack

__device__ __forceinline__ size_t slot_row0(int slot) {
  return static_cast<size_t>(slot / SLOTS_PER_TILE) * TILE_ROWS + static_cast<size_t>(slot % SLOTS_PER_TILE) * NJ;
}
__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}

__global__ void __launch_bounds__(K2T_THREADS, K2T_CTAS_PER_SM) gcn_input_umma_kernel(const __grid_constant__ InputLayerParams p) {
  ptx::pdl_launch_dependents();   // the first hidden layer (programmatic serialization) sets up meanwhile
  extern __shared__ uint8_t k2t_smem_raw[];
  // align inside the extern array by an offset (not through an integer cast) so the compiler keeps the shared address space
  uint8_t* sm = k2t_smem_raw + ((1024u - (ptx::smem_u32(k2t_smem_raw) & 1023u)) & 1023u);
  uint8_t* a_hi = sm;
  uint8_t* a_lo = sm + K2T_A_BYTES;
  uint8_t* b_hi = sm + 2 * K2T_A_BYTES;
  uint8_t* b_lo = b_hi + K2T_B_BYTES;
  float* mods = reinterpret_cast<float*>(sm + K2T_OFF_MOD);   // [24][128]
  float* xs = reinterpret_cast<float*>(sm + K2T_OFF_XS);      // [144] x_t of the slot's body, then [24] visibility
  float* visf = xs + XDIM;
  float* tile = reinterpret_cast<float*>(a_hi);   // [24][128] output staging, aliases the A operand after its MMAs
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + K2T_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + K2T_OFF_BAR + 8);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int C = p.C;
  const int n_chunks = C / 128;                     // C % 128 == 0
  const int chunk = blockIdx.x % n_chunks;          // gridDim.x % n_chunks == 0 (launcher)
  const int c = chunk * 128 + tid;

  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, K2T_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
  }
  {
    // B operand: off[j][i] * 2^10 as hi + lo fp16, K-major rows (output joint j) of 64 bytes with the 64-byte swizzle
    // (16-byte chunk index ^= (row >> 1) & 3); rows / columns 24..31 are zero
    const int n = tid >> 2, q = tid & 3;
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __half h[2], l[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = q * 8 + e * 2 + u;
        const float v = (n < NJ && i < NJ) ? p.adj.off[n][i] * K2T_OFF_SCALE : 0.f;
        h[u] = __float2half_rn(v);
        l[u] = __float2half_rn(v - __half2float(h[u]));
      }
      hw[e] = pack_half2(h[0], h[1]);
      lw[e] = pack_half2(l[0], l[1]);
    }
    const uint32_t off = n * 64 + ((q ^ ((n >> 1) & 3)) << 4);
    *reinterpret_cast<uint4*>(b_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(b_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) mods[j * 128 + tid] = p.mod[j * C + c] * p.act_scale;   // power of two: exact
  float w0[6], w1[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    w0[d] = p.wx01[(0 * 6 + d) * C + c];
    w1[d] = p.wx01[(1 * 6 + d) * C + c];
  }
  const float ct0 = p.ct01[(static_cast<size_t>(p.step) * 2 + 0) * C + c];
  const float ct1 = p.ct01[(static_cast<size_t>(p.step) * 2 + 1) * C + c];
  // the mix runs on act_scale * (M . h) (the fp16 operand scaling); BatchNorm's scale undoes it.  Powers of two: exact.
  const float sc = p.bn_scale[c] / p.act_scale, sh = p.bn_shift[c];
  const float inv = 1.f / K2T_OFF_SCALE;
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = ptx::make_idesc_f16_f32(128, 32);
  const uint64_t da_hi = ptx::make_kmajor_desc<64>(ptx::smem_u32(a_hi)), da_lo = ptx::make_kmajor_desc<64>(ptx::smem_u32(a_lo));
  const uint64_t db_hi = ptx::make_kmajor_desc<64>(ptx::smem_u32(b_hi)), db_lo = ptx::make_kmajor_desc<64>(ptx::smem_u32(b_lo));
  const uint32_t sw = (tid >> 1) & 3;
  const size_t C2 = 2 * static_cast<size_t>(C);
  uint32_t phase = 0;
  float amax = 0.f;

  // Software pipeline over this CTA's slots: the index -> image -> operand loads of a slot are a three-deep dependent
  // chain (about 1.5 us of L2 latency), so each level is fetched one slot further ahead and lands during the compute of
  // the slots before it.  Out-of-range look-ahead slots clamp to the last slot (loads stay in bounds, values unused).
  const int stride = gridDim.x / n_chunks;
  const int slot0 = blockIdx.x / n_chunks;
  const int last = p.n_slots - 1;
  struct Operands {
    float xa, xb, a0, a1, base0, base1;
  };
  auto load_operands = [&](int body, int img, bool cond) {
    Operands o;
    o.xa = p.x_t[static_cast<size_t>(body) * XDIM + tid];
    o.xb = 0.f;
    if (tid < XDIM - 128) o.xb = p.x_t[static_cast<size_t>(body) * XDIM + 128 + tid];
    else if (tid < XDIM - 128 + NJ) o.xb = (cond && p.vis[img * NJ + (tid - (XDIM - 128))]) ? 1.f : 0.f;
    o.a0 = p.a01[(static_cast<size_t>(img) * 2 + 0) * C + c];
    o.a1 = p.a01[(static_cast<size_t>(img) * 2 + 1) * C + c];
    const bool drop_all = !cond && p.mask_all;
    o.base0 = (drop_all ? p.cx01[c] : p.be01[(static_cast<size_t>(img) * 2 + 0) * C + c]) + ct0;
    o.base1 = (drop_all ? p.cx01[C + c] : p.be01[(static_cast<size_t>(img) * 2 + 1) * C + c]) + ct1;
    return o;
  };
  int body1 = p.slot_body[min(slot0 + stride, last)];
  int body2 = p.slot_body[min(slot0 + 2 * stride, last)];
  int img1 = p.img_of_body[body1];
  bool cond1 = p.slot_cond[min(slot0 + stride, last)] != 0;
  Operands cur;
  {
    const int body0 = p.slot_body[min(slot0, last)];
    cur = load_operands(body0, p.img_of_body[body0], p.slot_cond[min(slot0, last)] != 0);
  }

  for (int slot = slot0; slot < p.n_slots; slot += stride) {
    xs[tid] = cur.xa;
    if (tid < XDIM - 128 + NJ) xs[128 + tid] = cur.xb;
    const float a0 = cur.a0, a1 = cur.a1, base0 = cur.base0, base1 = cur.base1;
    // look-ahead: operands of the next slot, image of the one after, body of the third
    cur = load_operands(body1, img1, cond1);
    img1 = p.img_of_body[body2];
    cond1 = p.slot_cond[min(slot + 2 * stride, last)] != 0;
    body1 = body2;
    body2 = p.slot_body[min(slot + 3 * stride, last)];
    __syncthreads();   // xs / visf staged (their previous readers passed the barrier wait below)

    float y0[NJ];
    uint32_t hw[12], lw[12];
#pragma unroll
    for (int j2 = 0; j2 < NJ / 2; ++j2) {
      __half h[2], l[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = 2 * j2 + u;
        float x0 = 0.f, x1 = 0.f;
#pragma unroll
        for (int d = 0; d < 6; ++d) {
          x0 = fmaf(xs[j * 6 + d], w0[d], x0);
          x1 = fmaf(xs[j * 6 + d], w1[d], x1);
        }
        const float h0 = fmaf(visf[j], a0, base0) + x0;
        const float h1 = fmaf(visf[j], a1, base1) + x1;
        const float m = mods[j * 128 + tid];
        y0[j] = p.adj.diag[j] * (m * h0);
        const float g = m * h1;
        h[u] = __float2half_rn(g);
        l[u] = __float2half_rn(g - __half2float(h[u]));
        amax = fmaxf(amax, fabsf(g));
      }
      hw[j2] = pack_half2(h[0], h[1]);
      lw[j2] = pack_half2(l[0], l[1]);
    }
    {
      uint8_t* rh = a_hi + tid * 64;
      uint8_t* rl = a_lo + tid * 64;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        *reinterpret_cast<uint4*>(rh + ((q ^ sw) << 4)) = make_uint4(hw[4 * q], hw[4 * q + 1], hw[4 * q + 2], hw[4 * q + 3]);
        *reinterpret_cast<uint4*>(rl + ((q ^ sw) << 4)) = make_uint4(lw[4 * q], lw[4 * q + 1], lw[4 * q + 2], lw[4 * q + 3]);
      }
      *reinterpret_cast<uint4*>(rh + ((3 ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(rl + ((3 ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
    }
    ptx::fence_proxy_async_smem();   // generic-proxy writes of the A operand -> visible to the tensor pipe's async proxy
    ptx::tc_fence_before_sync();     // this thread's tcgen05.ld of the previous slot precedes the next MMA
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after_sync();
      // k-slices of 16 joints are 32 bytes apart inside the swizzled row (descriptor start address is in 16-byte units)
      ptx::umma_f16(tmem_base, da_hi, db_lo, idesc, 0);
      ptx::umma_f16(tmem_base, da_hi + 2, db_lo + 2, idesc, 1);
      ptx::umma_f16(tmem_base, da_lo, db_hi, idesc, 1);
      ptx::umma_f16(tmem_base, da_lo + 2, db_hi + 2, idesc, 1);
      ptx::umma_f16(tmem_base, da_hi, db_hi, idesc, 1);
      ptx::umma_f16(tmem_base, da_hi + 2, db_hi + 2, idesc, 1);
      ptx::umma_commit(bar);
    }
    ptx::mbar_wait(bar, phase);
    phase ^= 1;
    ptx::tc_fence_after_sync();
    float acc[32];
    ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), acc);
    ptx::tmem_ld_wait();

    // BatchNorm(eval) + ReLU, then the 24 x 128 output tile goes through shared memory (the A operand's space: the MMAs
    // that read it have completed) so that it leaves as 128-bit fp32 / 64-bit fp16 hi|lo row segments
#pragma unroll
    for (int j = 0; j < NJ; ++j) tile[j * 128 + tid] = fmaxf(fmaf(fmaf(acc[j], inv, y0[j]), sc, sh), 0.f);
    __syncthreads();
    float* res_row = p.res + slot_row0(slot) * C + chunk * 128;
    __half* hl_row = p.out_hl + slot_row0(slot) * C2 + chunk * 128;
#pragma unroll
    for (int it = 0; it < NJ * 32 / K2T_THREADS; ++it) {
      const int e = it * K2T_THREADS + tid;
      const int j = e >> 5, q = (e & 31) * 4;
      const float4 v = *reinterpret_cast<const float4*>(&tile[j * 128 + q]);
      *reinterpret_cast<float4*>(res_row + static_cast<size_t>(j) * C + q) = v;
      const float s0 = v.x * p.act_scale, s1 = v.y * p.act_scale, s2 = v.z * p.act_scale, s3 = v.w * p.act_scale;
      const __half2 h01 = __floats2half2_rn(s0, s1), h23 = __floats2half2_rn(s2, s3);
      const __half2 l01 = __floats2half2_rn(s0 - __low2float(h01), s1 - __high2float(h01));
      const __half2 l23 = __floats2half2_rn(s2 - __low2float(h23), s3 - __high2float(h23));
      uint2 hi, lo;
      hi.x = *reinterpret_cast<const uint32_t*>(&h01);
      hi.y = *reinterpret_cast<const uint32_t*>(&h23);
      lo.x = *reinterpret_cast<const uint32_t*>(&l01);
      lo.y = *reinterpret_cast<const uint32_t*>(&l23);
      *reinterpret_cast<uint2*>(hl_row + static_cast<size_t>(j) * C2 + q) = hi;
      *reinterpret_cast<uint2*>(hl_row + static_cast<size_t>(j) * C2 + C + q) = lo;
      amax = fmaxf(amax, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));   // post-ReLU: non-negative
    }
  }
  if (!(amax <= 65504.f)) atomicExch(p.overflow_flag, 1);
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, K2T_TMEM_COLS);
}

}  // namespace

cudaError_t launch_gcn_input_umma(const InputLayerParams& p, int num_sms, cudaStream_t stream) {
  if (p.n_slots <= 0) return cudaSuccess;
  const int n_chunks = p.C / 128;
  int per_chunk = (num_sms * K2T_CTAS_PER_SM) / n_chunks;
  if (per_chunk > p.n_slots) per_chunk = p.n_slots;
  if (per_chunk < 1) per_chunk = 1;
  gcn_input_umma_kernel<<<per_chunk * n_chunks, K2T_THREADS, K2T_SMEM, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ehb

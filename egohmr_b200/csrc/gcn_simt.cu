// SIMT (fp32 FFMA) kernels of the denoise step that are HBM/latency bound rather than contraction bound:
//   K2  gcn_input_kernel   — the algebraically folded 3718-wide input ModulatedGraphConv + BN + ReLU
//   K3  gcn_output_kernel  — output ModulatedGraphConv (hid -> 6), cond/uncond fuse-select and the DDIM/DDPM update
//   sgemm_nn               — plain fp32 GEMM for the once-per-batch conditioning folds and for the fp32 check path
//   gcn_hidden_epilogue    — check-path twin of the tcgen05 kernel's epilogue (reads H = X.Wcat from global)
#include "kernels.cuh"
#include "ptx.cuh"

namespace ehb {
namespace {

// ------------------------------------------------------------------------------------------------ sgemm
// GT x GT output tile per 256-thread block, (GT/16)^2 outputs per thread, K consumed GK at a time.  GT = 32 is used for
// skinny problems (the per-batch conditioning folds have M = n_img rows).  blockIdx.z selects a K chunk (split-K): chunk
// z accumulates A[:, z*kc : (z+1)*kc] . B[z*kc : (z+1)*kc, :] into its own partial plane C + z*plane (summed in a fixed
// order by splitk_reduce_kernel, so results do not depend on scheduling).
template <int GT, int GK>
__global__ void __launch_bounds__(256) sgemm_nn_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ C, int M, int N, int K, int lda, int ldb,
                                                       int ldc, int accumulate, int kc, size_t plane) {
  __shared__ float As[GK][GT + 4];
  __shared__ float Bs[GK][GT + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int kbeg = blockIdx.z * kc, kend = min(K, kbeg + kc);
  C += static_cast<size_t>(blockIdx.z) * plane;
  constexpr int R = GT / 16;
  float acc[R][R] = {};
  for (int k0 = kbeg; k0 < kend; k0 += GK) {
    for (int e = threadIdx.x; e < GT * GK; e += 256) {
      const int mm = e / GK, kk = e % GK;
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < kend) ? A[static_cast<size_t>(gm) * lda + gk] : 0.f;
    }
    for (int e = threadIdx.x; e < GT * GK; e += 256) {
      const int kk = e / GT, nn = e % GT;
      const int gk = k0 + kk, gn = n0 + nn;
      Bs[kk][nn] = (gk < kend && gn < N) ? B[static_cast<size_t>(gk) * ldb + gn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[R], b[R];
#pragma unroll
      for (int i = 0; i < R; ++i) a[i] = As[kk][ty * R + i];
#pragma unroll
      for (int i = 0; i < R; ++i) b[i] = Bs[kk][tx * R + i];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int jn = 0; jn < R; ++jn) acc[i][jn] = fmaf(a[i], b[jn], acc[i][jn]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int gm = m0 + ty * R + i;
    if (gm >= M) continue;
#pragma unroll
    for (int jn = 0; jn < R; ++jn) {
      const int gn = n0 + tx * R + jn;
      if (gn >= N) continue;
      float* dst = C + static_cast<size_t>(gm) * ldc + gn;
      *dst = accumulate ? (*dst + acc[i][jn]) : acc[i][jn];
    }
  }
}

// C[m][n] = (accumulate ? C[m][n] : 0) + sum_z partial[z][m][n], z ascending
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ C, int M, int N, int ldc,
                                     int splits, int accumulate) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(M) * N) return;
  const int m = static_cast<int>(i / N), n = static_cast<int>(i % N);
  float* dst = C + static_cast<size_t>(m) * ldc + n;
  float acc = accumulate ? *dst : 0.f;
  for (int z = 0; z < splits; ++z) acc += partial[static_cast<size_t>(z) * M * N + i];
  *dst = acc;
}

// fp16 hi/lo operand split shared by every producer of a GEMM A operand
__device__ __forceinline__ float store_hl(__half* out_hl, size_t row, int C, int c, float v, float act_scale) {
  const float sv = v * act_scale;
  const __half hi = __float2half_rn(sv);
  const __half lo = __float2half_rn(sv - __half2float(hi));
  out_hl[row * (2 * static_cast<size_t>(C)) + c] = hi;
  out_hl[row * (2 * static_cast<size_t>(C)) + C + c] = lo;
  return fabsf(sv);
}

__device__ __forceinline__ size_t slot_row0(int slot) {
  return static_cast<size_t>(slot / SLOTS_PER_TILE) * TILE_ROWS + static_cast<size_t>(slot % SLOTS_PER_TILE) * NJ;
}

// ------------------------------------------------------------------------------------------------ K2
// Reference: EgoHMR.forward builds feat = [img*vis | scene | transl | cam | Linear6->512(x_t) | temb] (egohmr.py:190-236)
// and feeds it to gconv_input (modulated_gcn.py:99-101).  Because the feature is a concatenation and the layer is
// linear before the joint mix, feat.W_k splits into per-image, per-step and per-joint terms:
//   h_k[j] = vis[j]*a_k[img] (conditioned pass only) + be_k[img] + ct_k[step] + x_t[j,:].wx_k
// The image-masked pass drops the a_k term (mask_cond(force_mask=True, only_mask_img_cond=True), egohmr.py:150-156);
// with only_mask_img_cond=False it drops every condition (:157-158), i.e. be_k falls back to its constant part cx_k.
// HBM-bound: 2 x 98.3 KB written per slot.  Compute is thread = channel (coalesced parameter reads); the 24 x 128
// output tile is staged in shared memory and written out as 128-bit fp32 / 64-bit fp16 row segments.
__global__ void __launch_bounds__(128) gcn_input_kernel(const __grid_constant__ InputLayerParams p) {
  ptx::pdl_launch_dependents();   // the first hidden layer (launched with programmatic serialization) sets up meanwhile
  __shared__ float xs[NJ][6];
  __shared__ float visf[NJ];
  __shared__ __align__(16) float tile[NJ][128];
  const int slot = blockIdx.x;
  const int body = p.slot_body[slot];
  const int img = p.img_of_body[body];
  const bool cond = p.slot_cond[slot] != 0;
  for (int e = threadIdx.x; e < XDIM; e += blockDim.x) xs[e / 6][e % 6] = p.x_t[static_cast<size_t>(body) * XDIM + e];
  if (threadIdx.x < NJ) visf[threadIdx.x] = (cond && p.vis[img * NJ + threadIdx.x]) ? 1.f : 0.f;
  __syncthreads();
  const int C = p.C;
  const size_t row0 = slot_row0(slot);
  const int c0 = blockIdx.y * 128;  // C % 128 == 0
  {
    const int c = c0 + threadIdx.x;
    const float a0 = p.a01[(static_cast<size_t>(img) * 2 + 0) * C + c];
    const float a1 = p.a01[(static_cast<size_t>(img) * 2 + 1) * C + c];
    const bool drop_all = !cond && p.mask_all;
    const float be0 = drop_all ? p.cx01[c] : p.be01[(static_cast<size_t>(img) * 2 + 0) * C + c];
    const float be1 = drop_all ? p.cx01[C + c] : p.be01[(static_cast<size_t>(img) * 2 + 1) * C + c];
    const float base0 = be0 + p.ct01[(static_cast<size_t>(p.step) * 2 + 0) * C + c];
    const float base1 = be1 + p.ct01[(static_cast<size_t>(p.step) * 2 + 1) * C + c];
    float w0[6], w1[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      w0[d] = p.wx01[(0 * 6 + d) * C + c];
      w1[d] = p.wx01[(1 * 6 + d) * C + c];
    }
    float g[NJ], y[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float x0 = 0.f, x1 = 0.f;
#pragma unroll
      for (int d = 0; d < 6; ++d) {
        x0 = fmaf(xs[j][d], w0[d], x0);
        x1 = fmaf(xs[j][d], w1[d], x1);
      }
      const float h0 = fmaf(visf[j], a0, base0) + x0;
      const float h1 = fmaf(visf[j], a1, base1) + x1;
      const float m = p.mod[j * C + c];
      g[j] = m * h1;
      y[j] = p.adj.diag[j] * (m * h0);
    }
    const float sc = p.bn_scale[c], sh = p.bn_shift[c];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float acc = y[j];
#pragma unroll
      for (int i = 0; i < NJ; ++i) acc = fmaf(p.adj.off[j][i], g[i], acc);
      tile[j][threadIdx.x] = fmaxf(fmaf(acc, sc, sh), 0.f);
    }
  }
  __syncthreads();
  float amax = 0.f;
  const size_t C2 = 2 * static_cast<size_t>(C);
  for (int e = threadIdx.x; e < NJ * 32; e += 128) {
    const int j = e >> 5, q = (e & 31) * 4;
    const float4 v = *reinterpret_cast<const float4*>(&tile[j][q]);
    *reinterpret_cast<float4*>(p.res + (row0 + j) * C + c0 + q) = v;
    const float s0 = v.x * p.act_scale, s1 = v.y * p.act_scale, s2 = v.z * p.act_scale, s3 = v.w * p.act_scale;
    const __half2 h01 = __floats2half2_rn(s0, s1), h23 = __floats2half2_rn(s2, s3);
    const __half2 l01 = __floats2half2_rn(s0 - __low2float(h01), s1 - __high2float(h01));
    const __half2 l23 = __floats2half2_rn(s2 - __low2float(h23), s3 - __high2float(h23));
    uint2 hi, lo;
    hi.x = *reinterpret_cast<const uint32_t*>(&h01);
    hi.y = *reinterpret_cast<const uint32_t*>(&h23);
    lo.x = *reinterpret_cast<const uint32_t*>(&l01);
    lo.y = *reinterpret_cast<const uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(p.out_hl + (row0 + j) * C2 + c0 + q) = hi;
    *reinterpret_cast<uint2*>(p.out_hl + (row0 + j) * C2 + C + c0 + q) = lo;
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(s0), fabsf(s1)), fmaxf(fabsf(s2), fabsf(s3))));
  }
  if (!(amax <= 65504.f)) atomicExch(p.overflow_flag, 1);
}

// One element of the reverse-diffusion update, in the reference's fp32 op order with contraction disabled.  `x0` is
// in/out: the guided DDIM step (ddim_sample_with_grad, gaussian_diffusion.py:579-592) replaces pred_xstart.
__device__ __forceinline__ float sampler_update_one(const StepCoef& coef, int kind, float x, float& x0, const float* noise,
                                                    const float* grad, size_t idx) {
  if (kind == SAMPLER_DDIM) {
    const float cx = __fmul_rn(coef.c[0], x);
    float eps = __fdiv_rn(__fsub_rn(cx, x0), coef.c[1]);                       // _predict_eps_from_xstart (:286-290)
    if (grad && coef.c[5] != 0.f) {
      eps = __fsub_rn(eps, __fmul_rn(__fmul_rn(coef.c[5], grad[idx]), 1.0f));  // eps - sqrt(1 - alpha_bar) * grad * scale
      x0 = __fsub_rn(cx, __fmul_rn(coef.c[1], eps));                           // _predict_xstart_from_eps (:277-283)
      eps = __fdiv_rn(__fsub_rn(cx, x0), coef.c[1]);
    }
    const float mean = __fadd_rn(__fmul_rn(x0, coef.c[2]), __fmul_rn(coef.c[3], eps));
    // nonzero_mask * sigma * noise (:552-555); sigma = 0 for eta = 0, where the draw is not even read
    return (noise && coef.c[4] != 0.f) ? __fadd_rn(mean, __fmul_rn(coef.c[4], noise[idx])) : mean;
  }
  float mean = __fadd_rn(__fmul_rn(coef.c[0], x0), __fmul_rn(coef.c[1], x));
  if (grad) mean = __fadd_rn(mean, __fmul_rn(coef.c[3], grad[idx]));
  const float nz = noise ? __fmul_rn(coef.c[2], noise[idx]) : 0.f;
  return __fadd_rn(mean, nz);
}

__global__ void sampler_update_kernel(const StepCoef coef, int kind, const float* __restrict__ x_t,
                                      const float* __restrict__ x0, const float* __restrict__ noise,
                                      const float* __restrict__ grad, float* __restrict__ x_prev,
                                      float* __restrict__ x0_out, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    float v0 = x0[i];
    x_prev[i] = sampler_update_one(coef, kind, x_t[i], v0, noise, grad, i);
    if (x0_out) x0_out[i] = v0;
  }
}

// ------------------------------------------------------------------------------------------------ K3
// Output ModulatedGraphConv (no BN/ReLU, modulated_gcn.py:111), the diffuse_fuse select of egohmr.py:239-254
// (guidance_param == 0: invisible joints take the image-masked pass, visible joints the image-conditioned pass)
// and one sampler update (gaussian_diffusion.py:298-337 p_sample, :340-388 with gradient, :511-556 ddim_sample).
// The update replays the reference's fp32 op order with contraction disabled so equal inputs give equal bits.
__global__ void __launch_bounds__(256, 2) gcn_output_fallback_kernel(const __grid_constant__ OutputLayerParams p) {
  __shared__ float hs[2][NJ][12];
  const int body = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C;
  const int slots[2] = {p.body_slot[body * 2 + 0], p.body_slot[body * 2 + 1]};
  // HBM-bound: 2 x 98.3 KB read per body.  Each warp reduces six (pass, joint) rows at once so that one pass over the
  // [C][12] output weights (L1-resident, 48 KB) serves six rows: the kernel moves 2 B of L1 traffic per HBM byte
  // instead of 12.  Lane l owns k = 128*it + 4*l .. +3 of every row (128-bit coalesced loads).
  constexpr int RW = 6;   // rows per warp = 2*NJ / 8 warps
  const float* arow[RW];
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int rr = warp + 8 * r, pass = rr / NJ, j = rr % NJ;
    arow[r] = slots[pass] < 0 ? nullptr : p.act + (slot_row0(slots[pass]) + j) * C;
  }
  float acc[RW][12] = {};
  for (int k = lane * 4; k < C; k += 128) {
    float4 a[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r)
      a[r] = arow[r] ? *reinterpret_cast<const float4*>(arow[r] + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* wp = reinterpret_cast<const float4*>(p.wout + static_cast<size_t>(k) * 12);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w0 = __ldg(wp + 3 * kk), w1 = __ldg(wp + 3 * kk + 1), w2 = __ldg(wp + 3 * kk + 2);
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const float av = kk == 0 ? a[r].x : (kk == 1 ? a[r].y : (kk == 2 ? a[r].z : a[r].w));
        acc[r][0] = fmaf(av, w0.x, acc[r][0]); acc[r][1] = fmaf(av, w0.y, acc[r][1]);
        acc[r][2] = fmaf(av, w0.z, acc[r][2]); acc[r][3] = fmaf(av, w0.w, acc[r][3]);
        acc[r][4] = fmaf(av, w1.x, acc[r][4]); acc[r][5] = fmaf(av, w1.y, acc[r][5]);
        acc[r][6] = fmaf(av, w1.z, acc[r][6]); acc[r][7] = fmaf(av, w1.w, acc[r][7]);
        acc[r][8] = fmaf(av, w2.x, acc[r][8]); acc[r][9] = fmaf(av, w2.y, acc[r][9]);
        acc[r][10] = fmaf(av, w2.z, acc[r][10]); acc[r][11] = fmaf(av, w2.w, acc[r][11]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int rr = warp + 8 * r, pass = rr / NJ, j = rr % NJ;
#pragma unroll
    for (int o = 0; o < 12; ++o) {
      float v = acc[r][o];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (lane == 0) hs[pass][j][o] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x >= XDIM) return;
  const int e = threadIdx.x, j = e / 6, d = e % 6;
  float out[2] = {0.f, 0.f};
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (slots[pass] < 0) continue;
    float acc = p.adj.diag[j] * (p.mod[j * 6 + d] * hs[pass][j][d]);
    for (int i = 0; i < NJ; ++i) acc = fmaf(p.adj.off[j][i], p.mod[i * 6 + d] * hs[pass][i][6 + d], acc);
    out[pass] = acc + p.bias[d];
  }
  const int img = p.img_of_body[body];
  float x0;
  if (p.diffuse_fuse) {
    x0 = p.vis[img * NJ + j] ? out[0] : out[1];
  } else {
    x0 = out[0];
  }
  const size_t idx = static_cast<size_t>(body) * XDIM + e;
  if (p.out_cond) p.out_cond[idx] = out[0];
  if (p.out_uncond) p.out_uncond[idx] = out[1];
  if (p.x0_model) p.x0_model[idx] = x0;
  p.x_prev[idx] = sampler_update_one(p.coef, p.kind, p.x_t[idx], x0, p.noise, p.grad, idx);
  p.x0[idx] = x0;
}

// ------------------------------------------------------------------------------------------------ K3 (bulk-staged)
// Same arithmetic as gcn_output_fallback_kernel, restructured around the memory system: the activations of one
// (body, pass) slot are ONE contiguous 24 x C fp32 block, so a producer warp streams them with 1-D bulk copies
// (cp.async.bulk + mbarrier transaction counts) through a 6-stage shared-memory ring of 8-row chunks while 8 consumer
// warps reduce them.  Persistent blocks (one per SM) keep ~190 KB of loads in flight per SM without spending registers
// on them; every consumer warp owns a fixed C/8-channel slice whose [k][12] output weights stay in registers for the
// whole launch, so the weights cost no memory traffic at all.
constexpr int K3_STAGES = 6, K3_ROWS = 8, K3_CONSUMERS = 8, K3_KPL = 4;   // K3_KPL: channels per lane, C <= 1024
constexpr int K3_NACC = K3_ROWS * 12;

struct K3Barriers {
  uint64_t full[K3_STAGES];
  uint64_t empty[K3_STAGES];
};

__global__ void __launch_bounds__((K3_CONSUMERS + 1) * 32, 1) gcn_output_kernel(const __grid_constant__ OutputLayerParams p) {
  extern __shared__ __align__(128) uint8_t k3_smem[];
  const int C = p.C;
  const uint32_t chunk_bytes = K3_ROWS * C * sizeof(float);
  float* ring = reinterpret_cast<float*>(k3_smem);
  float* part = reinterpret_cast<float*>(k3_smem + K3_STAGES * chunk_bytes);     // [2][K3_CONSUMERS][K3_NACC]
  float* hs = part + 2 * K3_CONSUMERS * K3_NACC;                                   // [2][NJ][12]
  K3Barriers* bars = reinterpret_cast<K3Barriers*>(hs + 2 * NJ * 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int passes = p.diffuse_fuse ? 2 : 1;
  const int n_chunks = passes * (NJ / K3_ROWS);
  if (threadIdx.x == 0) {
    for (int s = 0; s < K3_STAGES; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], K3_CONSUMERS);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp == K3_CONSUMERS) {
    // ---------------------------------------------------------------- producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int body = blockIdx.x; body < p.n_bodies; body += gridDim.x) {
        for (int ch = 0; ch < n_chunks; ++ch) {
          const int slot = p.body_slot[body * 2 + ch / (NJ / K3_ROWS)];
          const float* src = p.act + (slot_row0(slot) + static_cast<size_t>(ch % (NJ / K3_ROWS)) * K3_ROWS) * C;
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->full[stage], chunk_bytes);
          ptx::bulk_load_1d(ring + static_cast<size_t>(stage) * K3_ROWS * C, src, chunk_bytes, &bars->full[stage]);
          if (++stage == K3_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumers
  const int slice = C / K3_CONSUMERS;            // channels of this warp
  const int k0 = warp * slice;
  float w[K3_KPL][12];
#pragma unroll
  for (int i = 0; i < K3_KPL; ++i) {
    const int kk = lane + 32 * i;
#pragma unroll
    for (int o = 0; o < 12; ++o) w[i][o] = kk < slice ? __ldg(p.wout + static_cast<size_t>(k0 + kk) * 12 + o) : 0.f;
  }
  // after the butterfly below lane l holds the three sums with flat index (row * 12 + out) = base .. base + 2
  const int base = ((lane & 16) ? 48 : 0) + ((lane & 8) ? 24 : 0) + ((lane & 4) ? 12 : 0) + ((lane & 2) ? 6 : 0) + ((lane & 1) ? 3 : 0);
  int stage = 0;
  uint32_t phase = 0;
  uint32_t it = 0;   // chunk counter of this block (selects the partial-sum buffer)
  for (int body = blockIdx.x; body < p.n_bodies; body += gridDim.x) {
    float* hsb = hs;   // rewritten per body; the tail barrier below orders reuse
    for (int ch = 0; ch < n_chunks; ++ch, ++it) {
      ptx::mbar_wait(&bars->full[stage], phase);
      const float* a = ring + static_cast<size_t>(stage) * K3_ROWS * C + k0;
      float acc[K3_NACC];
#pragma unroll
      for (int e = 0; e < K3_NACC; ++e) acc[e] = 0.f;
#pragma unroll
      for (int i = 0; i < K3_KPL; ++i) {
        const int kk = lane + 32 * i;
        if (kk < slice) {
#pragma unroll
          for (int r = 0; r < K3_ROWS; ++r) {
            const float av = a[static_cast<size_t>(r) * C + kk];
#pragma unroll
            for (int o = 0; o < 12; ++o) acc[r * 12 + o] = fmaf(av, w[i][o], acc[r * 12 + o]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->empty[stage]);   // this warp is done reading the stage
      if (++stage == K3_STAGES) {
        stage = 0;
        phase ^= 1;
      }
      // sum over the warp's 32 lanes: butterfly reduce-scatter, 96 -> 3 values per lane in 93 shuffles
#pragma unroll
      for (int s = 16, n = K3_NACC; s >= 1; s >>= 1, n >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int e = 0; e < n / 2; ++e) {
          const float send = up ? acc[e] : acc[e + n / 2];
          const float keep = up ? acc[e + n / 2] : acc[e];
          acc[e] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      float* pb = part + ((it & 1) * K3_CONSUMERS + warp) * K3_NACC;
      pb[base] = acc[0];
      pb[base + 1] = acc[1];
      pb[base + 2] = acc[2];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x < K3_NACC) {   // sum over the 8 channel slices, fixed order
        const float* q = part + (it & 1) * K3_CONSUMERS * K3_NACC + threadIdx.x;
        float sum = 0.f;
#pragma unroll
        for (int ww = 0; ww < K3_CONSUMERS; ++ww) sum += q[ww * K3_NACC];
        const int pass = ch / (NJ / K3_ROWS), j = (ch % (NJ / K3_ROWS)) * K3_ROWS + threadIdx.x / 12;
        hsb[(pass * NJ + j) * 12 + threadIdx.x % 12] = sum;
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x < XDIM) {
      const int e = threadIdx.x, j = e / 6, d = e % 6;
      float out[2] = {0.f, 0.f};
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        if (pass >= passes) continue;
        const float* h = hsb + pass * NJ * 12;
        float acc = p.adj.diag[j] * (p.mod[j * 6 + d] * h[j * 12 + d]);
        for (int i = 0; i < NJ; ++i) acc = fmaf(p.adj.off[j][i], p.mod[i * 6 + d] * h[i * 12 + 6 + d], acc);
        out[pass] = acc + p.bias[d];
      }
      const int img = p.img_of_body[body];
      float x0 = p.diffuse_fuse ? (p.vis[img * NJ + j] ? out[0] : out[1]) : out[0];
      const size_t idx = static_cast<size_t>(body) * XDIM + e;
      if (p.out_cond) p.out_cond[idx] = out[0];
      if (p.out_uncond) p.out_uncond[idx] = out[1];
      if (p.x0_model) p.x0_model[idx] = x0;
      p.x_prev[idx] = sampler_update_one(p.coef, p.kind, p.x_t[idx], x0, p.noise, p.grad, idx);
      p.x0[idx] = x0;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // hs is rewritten by the next body's first chunk
  }
}

// ------------------------------------------------------------------------------------------------ non-local block
// NONLocalBlock2D(hid, sub_sample=False) between the last residual block and the output layer when
// ModulatedGCN(nonlocal_layer=True) (modulated_gcn.py:93-109, nets/non_local_embedded_gaussian.py:61-85):
//   theta, phi, g = 1x1 convs hid -> hid/2;  f = softmax_j(theta_i . phi_j);  y_i = sum_j f_ij g_j;  z = x + BN(W y)
// The three projections and W run on the tcgen05 linear kernel (linear_umma.cu) as 256-column planes; this kernel is
// the 24 x 24 attention of one (body, pass) slot in between: it reads the theta|phi|g planes and writes y as the fp16
// hi/lo A operand of the W projection.
__global__ void __launch_bounds__(256) nonlocal_attention_kernel(const __grid_constant__ NonLocalParams p) {
  __shared__ float f[NJ][NJ + 1];
  const int slot = blockIdx.x;
  const size_t row0 = slot_row0(slot);
  const int I = p.inter;
  auto at = [&](int which, size_t row, int c) -> const float* {   // column n = which*inter + c of the projection output
    const int n = which * I + c;
    return p.tpg + static_cast<size_t>(n >> 8) * p.plane_stride + row * 256 + (n & 255);
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int pair = warp; pair < NJ * NJ; pair += 8) {
    const int i = pair / NJ, j = pair % NJ;
    float acc = 0.f;
    for (int c = lane; c < I; c += 32) acc = fmaf(*at(0, row0 + i, c), *at(1, row0 + j, c), acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) f[i][j] = acc;
  }
  __syncthreads();
  if (threadIdx.x < NJ) {   // F.softmax(f, dim=-1)
    const int i = threadIdx.x;
    float mx = f[i][0];
    for (int j = 1; j < NJ; ++j) mx = fmaxf(mx, f[i][j]);
    float sum = 0.f;
    for (int j = 0; j < NJ; ++j) {
      const float e = expf(f[i][j] - mx);
      f[i][j] = e;
      sum += e;
    }
    for (int j = 0; j < NJ; ++j) f[i][j] = f[i][j] / sum;
  }
  __syncthreads();
  float amax = 0.f;
  for (int c = threadIdx.x; c < I; c += blockDim.x) {
    float y[NJ];
#pragma unroll
    for (int i = 0; i < NJ; ++i) y[i] = 0.f;
    for (int j = 0; j < NJ; ++j) {
      const float gv = *at(2, row0 + j, c);
#pragma unroll
      for (int i = 0; i < NJ; ++i) y[i] = fmaf(f[i][j], gv, y[i]);
    }
#pragma unroll
    for (int i = 0; i < NJ; ++i) amax = fmaxf(amax, store_hl(p.y_hl, row0 + i, I, c, y[i], p.act_scale));
  }
  if (!(amax <= 65504.f)) atomicExch(p.overflow_flag, 1);
}

// z = x + BN(W y + b): res[row][c] += wy[row][c] * scale[c] + shift[c] on the valid slot rows
__global__ void __launch_bounds__(256) nonlocal_residual_kernel(const __grid_constant__ NonLocalParams p) {
  const int slot = blockIdx.x;
  const size_t row0 = slot_row0(slot);
  for (int e = threadIdx.x; e < NJ * p.C; e += blockDim.x) {
    const int j = e / p.C, c = e % p.C;
    const size_t row = row0 + j;
    const float wy = p.wy[static_cast<size_t>(c >> 8) * p.plane_stride + row * 256 + (c & 255)];
    p.res[row * p.C + c] += fmaf(wy, p.bn_scale[c], p.bn_shift[c]);
  }
}

// ------------------------------------------------------------------------------------------------ check path
struct HiddenSimtParams {
  AdjMix adj;
  const float* H;       // [rows_pad][2C]  h0 | h1
  const float* mod;     // [24][C] unscaled M
  const float* bn_scale;
  const float* bn_shift;
  const float* res_in;  // may alias out_f32
  float* out_f32;
  __half* out_hl;
  int* overflow_flag;
  float act_scale;
  int C, n_slots, add_res, write_hl;
};

__global__ void __launch_bounds__(128) gcn_hidden_epilogue_kernel(const __grid_constant__ HiddenSimtParams p) {
  const int slot = blockIdx.x;
  const int C = p.C;
  const size_t row0 = slot_row0(slot);
  float amax = 0.f;
  {
    const int c = blockIdx.y * 128 + threadIdx.x;  // C % 128 == 0
    float g[NJ], y[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float m = p.mod[j * C + c];
      const float h0 = p.H[(row0 + j) * (2 * static_cast<size_t>(C)) + c];
      const float h1 = p.H[(row0 + j) * (2 * static_cast<size_t>(C)) + C + c];
      g[j] = m * h1;
      y[j] = p.adj.diag[j] * (m * h0);
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float acc = y[j];
#pragma unroll
      for (int i = 0; i < NJ; ++i) acc = fmaf(p.adj.off[j][i], g[i], acc);
      y[j] = acc;
    }
    const float sc = p.bn_scale[c], sh = p.bn_shift[c];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float v = fmaxf(fmaf(y[j], sc, sh), 0.f);
      if (p.add_res) v += p.res_in[(row0 + j) * C + c];
      p.out_f32[(row0 + j) * C + c] = v;
      if (p.write_hl) amax = fmaxf(amax, store_hl(p.out_hl, row0 + j, C, c, v, p.act_scale));
    }
  }
  if (!(amax <= 65504.f)) atomicExch(p.overflow_flag, 1);
}

}  // namespace

cudaError_t launch_sgemm_nn(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                            int accumulate, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (static_cast<long long>((N + 63) / 64) * ((M + 63) / 64) < 256) {   // skinny: smaller tiles, more blocks
    dim3 grid((N + 31) / 32, (M + 31) / 32);
    sgemm_nn_kernel<32, 16><<<grid, 256, 0, stream>>>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, K, 0);
  } else {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    sgemm_nn_kernel<64, 16><<<grid, 256, 0, stream>>>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, K, 0);
  }
  return cudaGetLastError();
}

int sgemm_splitk_plan(int M, int N, int K, int num_sms) {
  // skinny problems only (few output tiles, long K): enough K chunks for ~8 blocks per SM (the loads of a block are latency-bound), each at least 32 deep
  const long long tiles = static_cast<long long>((N + 31) / 32) * ((M + 31) / 32);
  if (tiles >= 2LL * num_sms || K < 128) return 1;
  long long want = (8LL * num_sms + tiles - 1) / tiles;
  long long cap = K / 32;
  long long s = want < cap ? want : cap;
  return static_cast<int>(s < 1 ? 1 : (s > 64 ? 64 : s));
}

cudaError_t launch_sgemm_nn_splitk(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb,
                                   int ldc, int accumulate, float* scratch, int splits, int* n_launches,
                                   cudaStream_t stream) {
  if (n_launches) *n_launches = 0;
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (splits <= 1 || !scratch) {
    if (n_launches) *n_launches = 1;
    return launch_sgemm_nn(A, B, C, M, N, K, lda, ldb, ldc, accumulate, stream);
  }
  int kc = (K + splits - 1) / splits;
  kc = (kc + 31) / 32 * 32;
  splits = (K + kc - 1) / kc;
  dim3 grid((N + 31) / 32, (M + 31) / 32, splits);
  sgemm_nn_kernel<32, 32><<<grid, 256, 0, stream>>>(A, B, scratch, M, N, K, lda, ldb, N, 0, kc,
                                                    static_cast<size_t>(M) * N);
  const size_t n = static_cast<size_t>(M) * N;
  splitk_reduce_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(scratch, C, M, N, ldc, splits,
                                                                                   accumulate);
  if (n_launches) *n_launches = 2;
  return cudaGetLastError();
}

cudaError_t launch_gcn_input(const InputLayerParams& p, cudaStream_t stream) {
  if (p.n_slots <= 0) return cudaSuccess;
  gcn_input_kernel<<<dim3(p.n_slots, p.C / 128), 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_nonlocal_attention(const NonLocalParams& p, cudaStream_t stream) {
  if (p.n_slots <= 0) return cudaSuccess;
  nonlocal_attention_kernel<<<p.n_slots, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_nonlocal_residual(const NonLocalParams& p, cudaStream_t stream) {
  if (p.n_slots <= 0) return cudaSuccess;
  nonlocal_residual_kernel<<<p.n_slots, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_sampler_update(const StepCoef& coef, int kind, const float* x_t, const float* x0,
                                  const float* noise, const float* grad, float* x_prev, float* x0_out, size_t n,
                                  cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  sampler_update_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(coef, kind, x_t, x0, noise, grad,
                                                                                    x_prev, x0_out, n);
  return cudaGetLastError();
}

cudaError_t launch_gcn_output(const OutputLayerParams& p, int num_sms, cudaStream_t stream) {
  if (p.n_bodies <= 0) return cudaSuccess;
  if (p.C > 256 * K3_KPL || p.C % (8 * 4) != 0) {   // register-resident weight slices cover C <= 1024
    gcn_output_fallback_kernel<<<p.n_bodies, 256, 0, stream>>>(p);
    return cudaGetLastError();
  }
  const size_t smem = static_cast<size_t>(K3_STAGES) * K3_ROWS * p.C * sizeof(float) +
                      (2 * K3_CONSUMERS * K3_NACC + 2 * NJ * 12) * sizeof(float) + sizeof(K3Barriers);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(gcn_output_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    attr_smem = smem;
  }
  const int grid = p.n_bodies < num_sms ? p.n_bodies : num_sms;
  gcn_output_kernel<<<grid, (K3_CONSUMERS + 1) * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_gcn_hidden_simt(const float* x_f32, const float* wcat, float* h_tmp, const HiddenLayerParams& p,
                                   const float* mod_unscaled, float* out_f32, cudaStream_t stream) {
  // Check path: x_f32 is the fp32 layer input, p.res the residual source (add_res), out_f32 the fp32 destination
  // (may alias p.res: every element is read and written by the same thread).
  const int rows = p.n_mtiles * TILE_ROWS;
  cudaError_t e = launch_sgemm_nn(x_f32, wcat, h_tmp, rows, 2 * p.C, p.C, p.C, 2 * p.C, 2 * p.C, 0, stream);
  if (e != cudaSuccess) return e;
  HiddenSimtParams q;
  q.adj = p.adj;
  q.H = h_tmp;
  q.mod = mod_unscaled;
  q.bn_scale = p.bn_scale;
  q.bn_shift = p.bn_shift;
  q.res_in = p.res;
  q.out_f32 = out_f32;
  q.out_hl = p.out_hl;
  q.overflow_flag = p.overflow_flag;
  q.act_scale = p.act_scale;
  q.C = p.C;
  q.n_slots = p.n_slots;
  q.add_res = p.add_res;
  q.write_hl = p.write_hl;
  if (p.n_slots <= 0) return cudaSuccess;
  gcn_hidden_epilogue_kernel<<<dim3(p.n_slots, p.C / 128), 128, 0, stream>>>(q);
  return cudaGetLastError();
}

}  // namespace ehb

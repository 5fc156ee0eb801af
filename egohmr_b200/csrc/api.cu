// C ABI (include/egohmr_b200.h): context, weight ingestion/repacking, step orchestration.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cudaTypedefs.h>

#include "../../include/egohmr_b200.h"
#include "kernels.cuh"

namespace {

thread_local std::string g_last_error;

int fail(const std::string& msg) {
  g_last_error = msg;
  return 1;
}
int fail_cuda(const char* what, cudaError_t e) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return 1;
}
#define EHB_CUDA(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return fail_cuda(#expr, _e);  \
  } while (0)

// bumped whenever a workspace moves: a captured CUDA graph holds raw workspace pointers, so its owner compares this
// against the value at capture time before every replay (ehb_alloc_epoch)
uint64_t g_alloc_epoch = 0;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t n, bool zero = false) {
    if (n <= bytes && p) return cudaSuccess;
    ++g_alloc_epoch;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, n ? n : 1);
    if (e != cudaSuccess) return e;
    bytes = n;
    if (zero) e = cudaMemset(p, 0, n ? n : 1);
    return e;
  }
  template <typename T>
  cudaError_t upload(const std::vector<T>& h) {
    cudaError_t e = ensure(h.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  return fn;
}

// 2-D fp16 row-major [rows][cols] tensor, box = [box_rows][BK halfs]; the swizzle span equals the box row (64 B or 128 B)
// and matches the UMMA shared-memory descriptors of gcn_umma.cu
int make_tmap_f16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  auto fn = get_encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * sizeof(__half)};
  const int bk = ehb::gcn_hidden_umma_bk();
  cuuint32_t box[2] = {static_cast<cuuint32_t>(bk), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
  return 0;
}

// 4-D fp16 NHWC activation [N][H][W][C2] (C2 = hi(C) | lo(C) halves per pixel): box = nb images x th*stride rows x
// tw*stride columns x 64 halves, traversed with element stride `stride` in H and W (so th x tw pixels are loaded);
// 128-byte swizzle like the 2-D maps.
// `window` > 0: an overlapping-window view for the space-to-depth stem — a "pixel" of the map is the 64 halves starting
// at physical pixel x (physical pixels hold `window` = 16 halves, so consecutive map pixels overlap by 48 halves); W is
// then the number of valid window positions.
int make_tmap_f16_nhwc(CUtensorMap* tm, const void* base, uint64_t N, uint64_t H, uint64_t W, uint64_t C2, uint32_t nb,
                       uint32_t th, uint32_t tw, uint32_t stride, uint32_t window = 0, uint64_t W_phys = 0) {
  auto fn = get_encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[4] = {C2, W, H, N};
  cuuint64_t gstride[3] = {C2 * sizeof(__half), W * C2 * sizeof(__half), H * W * C2 * sizeof(__half)};
  if (window) {
    gdim[0] = 64;
    gstride[0] = window * sizeof(__half);
    gstride[1] = W_phys * window * sizeof(__half);
    gstride[2] = H * W_phys * window * sizeof(__half);
  }
  cuuint32_t box[4] = {64, tw * stride, th * stride, nb};
  cuuint32_t estr[4] = {1, stride, stride, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (4-D) failed with CUresult " + std::to_string(int(r)));
  return 0;
}

ehb::AdjMix make_adjmix(const float* adj, const float* adj2) {
  // modulated_gcn_conv.py:42-43: adj = self.adj + self.adj2; adj = (adj.T + adj) / 2   (fp32)
  float a[ehb::NJ][ehb::NJ];
  for (int i = 0; i < ehb::NJ; ++i)
    for (int j = 0; j < ehb::NJ; ++j) a[i][j] = adj[i * ehb::NJ + j] + adj2[i * ehb::NJ + j];
  ehb::AdjMix m;
  for (int j = 0; j < ehb::NJ; ++j)
    for (int i = 0; i < ehb::NJ; ++i) {
      const float s = (a[i][j] + a[j][i]) / 2.f;
      if (i == j) {
        m.diag[j] = s;
        m.off[j][i] = 0.f;
      } else {
        m.off[j][i] = s;
      }
    }
  return m;
}

void fold_bn(const ehb_gconv& g, std::vector<float>& scale, std::vector<float>& shift) {
  const int C = g.out_dim;
  scale.resize(C);
  shift.resize(C);
  for (int c = 0; c < C; ++c) {
    const double sc = double(g.bn_weight[c]) / std::sqrt(double(g.bn_var[c]) + double(g.bn_eps));
    scale[c] = float(sc);
    shift[c] = float(double(g.bn_bias[c]) + (double(g.bias[c]) - double(g.bn_mean[c])) * sc);
  }
}

__global__ void fill_rows_kernel(float* dst, const float* row, int n_rows, int width) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < static_cast<size_t>(n_rows) * width) dst[i] = row[i % width];
}

__global__ void denorm_kernel(const float* x, const float* mean, const float* std_, float* out, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fadd_rn(__fmul_rn(x[i], std_[i % ehb::XDIM]), mean[i % ehb::XDIM]);
}

}  // namespace

static void split_hl(const float* W, int rows, int cols, int ld, int col0, float scale, std::vector<__half>& out) {
  // rows x cols block of W (row stride ld, first column col0) -> [rows][hi(cols) | lo(cols)] fp16 of scale*W
  out.assign(static_cast<size_t>(rows) * 2 * cols, __half());
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const float sv = W[static_cast<size_t>(r) * ld + col0 + c] * scale;
      const __half hi = __float2half_rn(sv);
      out[static_cast<size_t>(r) * 2 * cols + c] = hi;
      out[static_cast<size_t>(r) * 2 * cols + cols + c] = __float2half_rn(sv - __half2float(hi));
    }
}
static float pow2_scale(float maxabs) {
  if (!(maxabs > 0.f)) return 1.f;
  float s = std::exp2(std::floor(std::log2(16384.f / maxabs)));
  if (maxabs * s >= 16384.f) s *= 0.5f;
  return s;
}

struct ehb_ctx {
  int device = 0;
  int num_sms = 0;
  int64_t launches = 0;
  int gemm_mode = 0;   // 0 = tcgen05 CTA pairs, transposed product (product path), 1 = fp32 FFMA check path,
                       // 2 = tcgen05 one-CTA row-major kernel, 3 = tcgen05 CTA pairs, row-major product
  int k1_fused = 0;    // 1 = the hidden layers of a reverse step as ONE persistent launch (gcn_umma_fused.cu): same bits, measured
                       // 7-8 % slower than one launch per layer, opt-in through ehb_debug_set_k1_fused
  DevBuf k1_done;      // its per-(layer, row group) completion counters
  int input_umma = 0;  // 1 = K2's joint mix on tcgen05 (gcn_input_umma.cu): measured equal to the FFMA kernel, opt-in
  int pdl = 1;         // hidden-layer launches use programmatic dependent launch (set-up overlaps the previous layer's tail)
  float act_scale = 8.f;

  // ---- denoiser
  bool gcn_loaded = false;
  int hid = 0, n_blocks = 0, img_dim = 0, cond_dim = 0, xfeat_dim = 0, temb_dim = 0, diffuse_fuse = 0, mask_all = 0;
  struct Hidden {
    ehb::AdjMix adj;
    DevBuf mod_scaled, mod, bn_scale, bn_shift, w_hl, wcat;
    float w_scale = 1.f;
    CUtensorMap tmB;    // box 256 rows (one-CTA kernel)
    CUtensorMap tmB2;   // box 128 rows (CTA-pair kernel: each CTA stages half of the B tile)
    CUtensorMap tmW;    // box 64 rows (transposed kernel: 64 channels of h0 / h1 per CTA)
  };
  std::vector<Hidden*> hidden;
  ehb::AdjMix adj_in, adj_out;
  DevBuf w_img, w_rest, w_temb, wx01, cx01, mod_in, bn_scale_in, bn_shift_in;
  DevBuf wout, mod_out, bias_out;

  // ---- optional non-local block (gcn_nonlocal_layer=True)
  bool nl_loaded = false;
  int nl_inter = 0, nl_tpg_planes = 0, nl_w_planes = 0;
  float nl_wscale_tpg = 1.f, nl_wscale_w = 1.f;
  DevBuf nl_wtpg_hl, nl_btpg, nl_ww_hl, nl_scale, nl_shift, nl_tpg, nl_wy, nl_y_hl;

  // ---- conditioning
  int n_img = 0, n_steps_cond = 0;
  DevBuf a01, be01, ct01, vis;

  // ---- chains
  int n_bodies = 0, n_slots = 0, n_mtiles = 0, max_img_of_body = 0;
  DevBuf img_of_body, slot_body, slot_cond, body_slot;
  DevBuf act_hl[2], res, mid, h_tmp;
  CUtensorMap tmA[2];
  CUtensorMap tmX[2];  // the same buffers with 120-row boxes (transposed kernel: the real rows of a 128-row tile)

  // ---- sampler
  int kind = 0;
  std::vector<ehb::StepCoef> coef;
  DevBuf mean, std_;
  bool norm_set = false;

  // ---- SMPL
  bool smpl_loaded = false;
  ehb::SmplDevice smpl{};
  DevBuf s_vt, s_sd, s_pd, s_w, s_jt, s_jsd, s_ex;
  DevBuf sc_R, sc_A, sc_j24, sc_pf, sc_p6, sc_dvp, sc_dA, sc_dpf;
  DevBuf s_pdT_hl, s_zero_bias, sc_pf_hl, sc_Y, s_w4, s_j4, s_sdT;   // tensor-core pose blend: posedirs^T operand, pose-feature operand, Y
  float s_pd_scale = 1.f;

  // ---- ResPointNet
  bool pn_loaded = false;
  int pn_hidden = 0, pn_out = 0;
  float pn_wscale[4][2] = {};   // [block][0 = fc_0, 1 = fc_1 + shortcut (shared)]
  DevBuf pn_pos_w, pn_pos_b, pn_fc0[4], pn_fc1[4], pn_sc[4], pn_b0[4], pn_b1[4], pn_w0b_t[4], pn_wsb_t[4], pn_fcc_t, pn_fcc_b;
  DevBuf pn_x[2], pn_y[2], pn_h, pn_pool, pn_pooled, pn_pooled_relu, pn_row0, pn_rows, pn_fold_w, pn_fold_b;

  // ---- ResNet-50 image encoder (conv_umma.cu)
  struct ConvPlan {
    int cout = 0, cin = 0, kh = 0, kw = 0, stride = 1, pad = 0, Kp = 0;
    bool stem = false;   // the space-to-depth stem: windowed tensor map, hi / lo planes
    float w_scale = 1.f;
    DevBuf w_hl, bias;
    std::vector<float> wf, bias_h;   // folded fp32 weights / bias (kept on the host until the scales are final)
  };
  std::vector<ConvPlan*> rn_convs;
  int rn_blocks[4] = {0, 0, 0, 0};
  bool rn_loaded = false;
  // Operand scale of the ResNet activations.  fp16's narrow exponent makes the lo half of a small value denormal
  // (|scale * v| < 0.125 starts losing relative precision) and ResNet activations are mostly << 1, so they are scaled
  // by 64 (exact, a power of two) instead of the GCN's 8: values up to 1023 stay in range, the overflow flag reports
  // beyond.  (Measured: the scale does not change the 1e-5 feature error, which grows linearly with depth — the
  // signature of the tensor core's truncating fp32 accumulation, not of operand rounding.)
  float rn_act_scale = 64.f;
  int rn_implicit = 1;   // 3x3 convolutions as implicit GEMMs through 4-D TMA boxes (0: explicit im2col matrix)
  // k-blocks (of 64) chained into one TMEM accumulation; longer contractions are summed chunk by chunk in fp32 registers
  // by the epilogue warps (0 = never split).  See DESIGN.md "K9 numerics".
  int rn_kc = 2;
  int rn_kc_1x1 = 2;
  DevBuf rn_col, rn_x[2], rn_y1, rn_y2, rn_s2d;

  DevBuf overflow, splitk;

  ~ehb_ctx() {
    for (auto* h : hidden) delete h;
    for (auto* c : rn_convs) delete c;
  }
};

// skinny fp32 GEMM of the once-per-batch folds / per-cloud rows, split over K so that it fills the GPU
static int ctx_sgemm(ehb_ctx* ctx, const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb,
                     int ldc, int accumulate, cudaStream_t stream) {
  const int splits = ehb::sgemm_splitk_plan(M, N, K, ctx->num_sms);
  if (splits > 1) EHB_CUDA(ctx->splitk.ensure(static_cast<size_t>(splits) * M * N * sizeof(float)));
  int nl = 0;
  EHB_CUDA(ehb::launch_sgemm_nn_splitk(A, B, C, M, N, K, lda, ldb, ldc, accumulate, ctx->splitk.as<float>(), splits, &nl,
                                       stream));
  ctx->launches += nl;
  return 0;
}

extern "C" {

const char* ehb_last_error(void) { return g_last_error.c_str(); }

int64_t ehb_launch_count(const ehb_ctx* ctx) { return ctx ? ctx->launches : 0; }

uint64_t ehb_alloc_epoch(void) { return g_alloc_epoch; }

int ehb_ctx_create(int device, ehb_ctx** out) {
  if (!out) return fail("ehb_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(std::string("ehb_ctx_create: no CUDA device (") + cudaGetErrorString(e) +
                "); this library has no CPU fallback");
  if (device < 0 || device >= n) return fail("ehb_ctx_create: bad device index");
  EHB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  EHB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail("ehb_ctx_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                ", this library is built for sm_100a only");
  ehb_ctx* c = new ehb_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  if (c->overflow.ensure(sizeof(int), true) != cudaSuccess) {
    delete c;
    return fail("ehb_ctx_create: cudaMalloc failed");
  }
  *out = c;
  return 0;
}

void ehb_ctx_destroy(ehb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  delete ctx;
}

int ehb_debug_set_gemm_mode(ehb_ctx* ctx, int gemm_mode) {
  if (!ctx) return fail("null ctx");
  if (gemm_mode < 0 || gemm_mode > 3)
    return fail("gemm_mode must be 0 (tcgen05 CTA pairs, transposed product), 1 (fp32 check), 2 (tcgen05 single CTA) or 3 (tcgen05 CTA pairs, row-major product)");
  ctx->gemm_mode = gemm_mode;
  return 0;
}

int ehb_gcn_load(ehb_ctx* ctx, const ehb_gcn_weights* w) {
  if (!ctx || !w) return fail("ehb_gcn_load: null argument");
  EHB_CUDA(cudaSetDevice(ctx->device));
  const int C = w->hid;
  if (C <= 0 || C % 128 != 0) return fail("ehb_gcn_load: hid must be a positive multiple of 128");
  if (w->n_layers != 2 * w->n_blocks + 2) return fail("ehb_gcn_load: n_layers must be 2*n_blocks + 2");
  const int in_dim = w->cond_dim + w->xfeat_dim + w->temb_dim;
  const ehb_gconv& gi = w->layers[0];
  const ehb_gconv& go = w->layers[w->n_layers - 1];
  if (gi.in_dim != in_dim || gi.out_dim != C) return fail("ehb_gcn_load: gconv_input shape mismatch");
  if (go.in_dim != C || go.out_dim != 6) return fail("ehb_gcn_load: gconv_output shape mismatch");
  if (!gi.bn_weight) return fail("ehb_gcn_load: gconv_input needs BatchNorm parameters");
  ctx->hid = C;
  ctx->n_blocks = w->n_blocks;
  ctx->img_dim = w->img_dim;
  ctx->cond_dim = w->cond_dim;
  ctx->xfeat_dim = w->xfeat_dim;
  ctx->temb_dim = w->temb_dim;
  ctx->diffuse_fuse = w->diffuse_fuse;
  ctx->mask_all = w->mask_all_cond ? 1 : 0;
  const size_t C2 = 2 * static_cast<size_t>(C);

  // ---- input layer: split W[2][in_dim][C] by feature block, columns concatenated as k*C + c
  {
    auto Wat = [&](int k, int r, int c) { return gi.W[(static_cast<size_t>(k) * in_dim + r) * C + c]; };
    const int rest = w->cond_dim - w->img_dim;
    std::vector<float> wi(static_cast<size_t>(w->img_dim) * C2), wr(static_cast<size_t>(rest) * C2),
        wt(static_cast<size_t>(w->temb_dim) * C2);
    for (int k = 0; k < 2; ++k)
      for (int c = 0; c < C; ++c) {
        for (int r = 0; r < w->img_dim; ++r) wi[r * C2 + k * C + c] = Wat(k, r, c);
        for (int r = 0; r < rest; ++r) wr[r * C2 + k * C + c] = Wat(k, w->img_dim + r, c);
        for (int r = 0; r < w->temb_dim; ++r) wt[r * C2 + k * C + c] = Wat(k, w->cond_dim + w->xfeat_dim + r, c);
      }
    // InputProcess folded through the x_feat rows: wx[k][d][c] = sum_f inproc_w[f][d] W_k[x0+f][c]; cx = inproc_b . W_k
    std::vector<float> wx(2 * 6 * static_cast<size_t>(C)), cx(C2);
    for (int k = 0; k < 2; ++k)
      for (int c = 0; c < C; ++c) {
        double acc[6] = {0, 0, 0, 0, 0, 0}, accb = 0;
        for (int f = 0; f < w->xfeat_dim; ++f) {
          const double wv = Wat(k, w->cond_dim + f, c);
          for (int d = 0; d < 6; ++d) acc[d] += double(w->inproc_w[f * 6 + d]) * wv;
          accb += double(w->inproc_b[f]) * wv;
        }
        for (int d = 0; d < 6; ++d) wx[(static_cast<size_t>(k) * 6 + d) * C + c] = float(acc[d]);
        cx[k * C + c] = float(accb);
      }
    std::vector<float> mod(gi.M, gi.M + static_cast<size_t>(ehb::NJ) * C), sc, sh;
    fold_bn(gi, sc, sh);
    EHB_CUDA(ctx->w_img.upload(wi));
    EHB_CUDA(ctx->w_rest.upload(wr));
    EHB_CUDA(ctx->w_temb.upload(wt));
    EHB_CUDA(ctx->wx01.upload(wx));
    EHB_CUDA(ctx->cx01.upload(cx));
    EHB_CUDA(ctx->mod_in.upload(mod));
    EHB_CUDA(ctx->bn_scale_in.upload(sc));
    EHB_CUDA(ctx->bn_shift_in.upload(sh));
    ctx->adj_in = make_adjmix(w->adj, gi.adj2);
  }

  // ---- hidden layers
  for (auto* h : ctx->hidden) delete h;
  ctx->hidden.clear();
  for (int l = 1; l <= 2 * w->n_blocks; ++l) {
    const ehb_gconv& g = w->layers[l];
    if (g.in_dim != C || g.out_dim != C || !g.bn_weight) return fail("ehb_gcn_load: hidden layer shape mismatch");
    auto* h = new ehb_ctx::Hidden();
    ctx->hidden.push_back(h);
    h->adj = make_adjmix(w->adj, g.adj2);
    float maxabs = 0.f;
    const size_t nW = 2 * static_cast<size_t>(C) * C;
    for (size_t i = 0; i < nW; ++i) maxabs = std::max(maxabs, std::fabs(g.W[i]));
    if (!std::isfinite(maxabs)) return fail("ehb_gcn_load: non-finite weight");
    // power-of-two scale putting max|W| in [8192, 16384): fp16 lo parts stay normal, no overflow
    h->w_scale = maxabs > 0.f ? std::exp2(std::floor(std::log2(16384.f / maxabs)) ) : 1.f;
    if (maxabs * h->w_scale >= 16384.f) h->w_scale *= 0.5f;
    // B operand: row n' = nt*256 + k*128 + cl  (channel nt*128+cl of W[k]), columns [hi(K) | lo(K)]
    std::vector<__half> whl(C2 * C2);
    std::vector<float> wcat(static_cast<size_t>(C) * C2);
    for (int k = 0; k < 2; ++k)
      for (int kk = 0; kk < C; ++kk)
        for (int c = 0; c < C; ++c) {
          const float v = g.W[(static_cast<size_t>(k) * C + kk) * C + c];
          wcat[kk * C2 + k * C + c] = v;
          const float sv = v * h->w_scale;
          const __half hi = __float2half_rn(sv);
          const __half lo = __float2half_rn(sv - __half2float(hi));
          const size_t np = static_cast<size_t>(c / 128) * 256 + k * 128 + (c % 128);
          whl[np * C2 + kk] = hi;
          whl[np * C2 + C + kk] = lo;
        }
    const float inv = 1.f / (ctx->act_scale * h->w_scale);
    std::vector<float> mod(g.M, g.M + static_cast<size_t>(ehb::NJ) * C), mods(mod.size()), sc, sh;
    for (size_t i = 0; i < mod.size(); ++i) mods[i] = mod[i] * inv;
    fold_bn(g, sc, sh);
    EHB_CUDA(h->w_hl.upload(whl));
    EHB_CUDA(h->wcat.upload(wcat));
    EHB_CUDA(h->mod.upload(mod));
    EHB_CUDA(h->mod_scaled.upload(mods));
    EHB_CUDA(h->bn_scale.upload(sc));
    EHB_CUDA(h->bn_shift.upload(sh));
    if (make_tmap_f16(&h->tmB, h->w_hl.p, C2, C2, 256)) return 1;
    if (make_tmap_f16(&h->tmB2, h->w_hl.p, C2, C2, 128)) return 1;
    if (make_tmap_f16(&h->tmW, h->w_hl.p, C2, C2, 64)) return 1;
  }

  // ---- output layer
  {
    std::vector<float> wo(static_cast<size_t>(C) * 12);
    for (int k = 0; k < 2; ++k)
      for (int kk = 0; kk < C; ++kk)
        for (int d = 0; d < 6; ++d) wo[static_cast<size_t>(kk) * 12 + k * 6 + d] = go.W[(static_cast<size_t>(k) * C + kk) * 6 + d];
    std::vector<float> mo(go.M, go.M + ehb::NJ * 6), bo(go.bias, go.bias + 6);
    EHB_CUDA(ctx->wout.upload(wo));
    EHB_CUDA(ctx->mod_out.upload(mo));
    EHB_CUDA(ctx->bias_out.upload(bo));
    ctx->adj_out = make_adjmix(w->adj, go.adj2);
  }
  ctx->gcn_loaded = true;
  ctx->nl_loaded = false;   // the optional non-local block is (re)loaded separately
  ctx->n_bodies = 0;
  ctx->n_slots = 0;
  ctx->n_mtiles = 0;        // activation buffers / tensor maps are rebuilt for the new `hid` by the next ehb_set_bodies
  return 0;
}

int ehb_set_norm(ehb_ctx* ctx, const float* mean, const float* std) {
  if (!ctx || !mean || !std) return fail("ehb_set_norm: null argument");
  EHB_CUDA(cudaSetDevice(ctx->device));
  std::vector<float> m(mean, mean + ehb::XDIM), s(std, std + ehb::XDIM);
  EHB_CUDA(ctx->mean.upload(m));
  EHB_CUDA(ctx->std_.upload(s));
  ctx->norm_set = true;
  return 0;
}

int ehb_set_schedule(ehb_ctx* ctx, int kind, int n_steps, const float* coef) {
  if (!ctx || !coef) return fail("ehb_set_schedule: null argument");
  if (kind != ehb::SAMPLER_DDIM && kind != ehb::SAMPLER_DDPM) return fail("ehb_set_schedule: kind must be 0 or 1");
  if (n_steps <= 0) return fail("ehb_set_schedule: n_steps must be positive");
  ctx->kind = kind;
  ctx->coef.resize(n_steps);
  std::memcpy(ctx->coef.data(), coef, sizeof(ehb::StepCoef) * n_steps);
  return 0;
}

int ehb_set_cond(ehb_ctx* ctx, int n_img, const float* img_feat, const float* rest_feat, const uint8_t* vis,
                 void* stream_) {
  if (!ctx || !img_feat || !rest_feat || !vis) return fail("ehb_set_cond: null argument");
  if (!ctx->gcn_loaded) return fail("ehb_set_cond: call ehb_gcn_load first");
  if (n_img <= 0) return fail("ehb_set_cond: n_img must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int C = ctx->hid, C2 = 2 * C, rest = ctx->cond_dim - ctx->img_dim;
  EHB_CUDA(ctx->a01.ensure(sizeof(float) * n_img * C2));
  EHB_CUDA(ctx->be01.ensure(sizeof(float) * n_img * C2));
  EHB_CUDA(ctx->vis.ensure(static_cast<size_t>(n_img) * ehb::NJ));
  EHB_CUDA(cudaMemcpyAsync(ctx->vis.p, vis, static_cast<size_t>(n_img) * ehb::NJ, cudaMemcpyDeviceToDevice, stream));
  if (ctx_sgemm(ctx, img_feat, ctx->w_img.as<float>(), ctx->a01.as<float>(), n_img, C2, ctx->img_dim, ctx->img_dim, C2, C2,
                0, stream))
    return 1;
  {
    const size_t n = static_cast<size_t>(n_img) * C2;
    fill_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(ctx->be01.as<float>(),
                                                                                 ctx->cx01.as<float>(), n_img, C2);
    EHB_CUDA(cudaGetLastError());
  }
  if (ctx_sgemm(ctx, rest_feat, ctx->w_rest.as<float>(), ctx->be01.as<float>(), n_img, C2, rest, rest, C2, C2, 1, stream))
    return 1;
  ctx->launches += 1;
  ctx->n_img = n_img;
  return 0;
}

int ehb_set_temb(ehb_ctx* ctx, int n_steps, const float* temb, void* stream_) {
  if (!ctx || !temb) return fail("ehb_set_temb: null argument");
  if (!ctx->gcn_loaded) return fail("ehb_set_temb: call ehb_gcn_load first");
  if (n_steps <= 0) return fail("ehb_set_temb: n_steps must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int C2 = 2 * ctx->hid;
  EHB_CUDA(ctx->ct01.ensure(sizeof(float) * n_steps * C2));
  if (ctx_sgemm(ctx, temb, ctx->w_temb.as<float>(), ctx->ct01.as<float>(), n_steps, C2, ctx->temb_dim, ctx->temb_dim, C2,
                C2, 0, stream))
    return 1;
  ctx->n_steps_cond = n_steps;
  return 0;
}

int ehb_set_bodies(ehb_ctx* ctx, int n_bodies, const int32_t* img_of_body) {
  if (!ctx || !img_of_body) return fail("ehb_set_bodies: null argument");
  if (!ctx->gcn_loaded) return fail("ehb_set_bodies: call ehb_gcn_load first");
  if (n_bodies <= 0) return fail("ehb_set_bodies: n_bodies must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  const int passes = ctx->diffuse_fuse ? 2 : 1;
  const int n_slots = n_bodies * passes;
  std::vector<int32_t> iob(img_of_body, img_of_body + n_bodies), sb(n_slots), bs(static_cast<size_t>(n_bodies) * 2, -1);
  std::vector<uint8_t> scnd(n_slots);
  for (int p = 0; p < passes; ++p)
    for (int b = 0; b < n_bodies; ++b) {
      const int s = p * n_bodies + b;
      sb[s] = b;
      scnd[s] = p == 0 ? 1 : 0;  // pass 0 = image-conditioned, pass 1 = image-masked
      bs[b * 2 + p] = s;
    }
  int max_img = 0;
  for (int b = 0; b < n_bodies; ++b) {
    if (iob[b] < 0) return fail("ehb_set_bodies: negative image index");
    max_img = std::max(max_img, iob[b]);
  }
  ctx->max_img_of_body = max_img;
  EHB_CUDA(ctx->img_of_body.upload(iob));
  EHB_CUDA(ctx->slot_body.upload(sb));
  EHB_CUDA(ctx->slot_cond.upload(scnd));
  EHB_CUDA(ctx->body_slot.upload(bs));
  // whole CTA-pair tiles: an even number of 128-row tiles (pad tiles hold zeros and are never stored)
  const int n_mtiles = ((n_slots + ehb::SLOTS_PER_TILE - 1) / ehb::SLOTS_PER_TILE + 1) / 2 * 2;
  const size_t rows = static_cast<size_t>(n_mtiles) * ehb::TILE_ROWS;
  const size_t C = ctx->hid;
  if (n_mtiles != ctx->n_mtiles || !ctx->res.p) {
    for (int i = 0; i < 2; ++i) {
      EHB_CUDA(ctx->act_hl[i].ensure(rows * 2 * C * sizeof(__half), true));
      EHB_CUDA(cudaMemset(ctx->act_hl[i].p, 0, rows * 2 * C * sizeof(__half)));
      if (make_tmap_f16(&ctx->tmA[i], ctx->act_hl[i].p, rows, 2 * C, 128)) return 1;
      if (make_tmap_f16(&ctx->tmX[i], ctx->act_hl[i].p, rows, 2 * C, ehb::SLOTS_PER_TILE * ehb::NJ)) return 1;
    }
    EHB_CUDA(ctx->res.ensure(rows * C * sizeof(float), true));
    EHB_CUDA(cudaMemset(ctx->res.p, 0, rows * C * sizeof(float)));
    EHB_CUDA(ctx->k1_done.ensure(sizeof(int) * ehb::MAX_FUSED_LAYERS * (n_mtiles / 2), true));
  }
  ctx->n_bodies = n_bodies;
  ctx->n_slots = n_slots;
  ctx->n_mtiles = n_mtiles;
  EHB_CUDA(cudaDeviceSynchronize());
  return 0;
}

static void fill_hidden_params(ehb_ctx* ctx, int l, ehb::HiddenLayerParams& p);

// All hidden layers of the step as one persistent launch (product path).
static int run_hidden_fused(ehb_ctx* ctx, cudaStream_t stream) {
  const int L = static_cast<int>(ctx->hidden.size());
  ehb::FusedHiddenMaps maps;
  ehb::FusedHiddenParams fp;
  maps.x[0] = ctx->tmX[0];
  maps.x[1] = ctx->tmX[1];
  for (int l = 0; l < L; ++l) {
    maps.w[l] = ctx->hidden[l]->tmW;
    fill_hidden_params(ctx, l, fp.layer[l]);
  }
  for (int l = L; l < ehb::MAX_FUSED_LAYERS; ++l) {   // unused slots: defined contents
    maps.w[l] = maps.w[0];
    fp.layer[l] = fp.layer[0];
  }
  fp.n_layers = L;
  fp.done = ctx->k1_done.as<int>();
  EHB_CUDA(ehb::launch_gcn_hidden_fused(maps, fp, ctx->num_sms, ctx->pdl != 0, stream));
  ctx->launches += 1;
  return 0;
}

static void fill_hidden_params(ehb_ctx* ctx, int l, ehb::HiddenLayerParams& p) {
  const int L = static_cast<int>(ctx->hidden.size());
  ehb_ctx::Hidden& h = *ctx->hidden[l];
  p.adj = h.adj;
  p.mod = h.mod_scaled.as<float>();
  p.bn_scale = h.bn_scale.as<float>();
  p.bn_shift = h.bn_shift.as<float>();
  p.res = ctx->res.as<float>();
  p.out_hl = ctx->act_hl[(l + 1) & 1].as<__half>();
  p.overflow_flag = ctx->overflow.as<int>();
  p.act_scale = ctx->act_scale;
  p.C = ctx->hid;
  p.n_mtiles = ctx->n_mtiles;
  p.n_ntiles = ctx->hid / 128;
  p.n_slots = ctx->n_slots;
  const bool second = (l & 1) == 1;  // gconv2 of a _ResGraphConv: residual add, block-boundary output
  p.add_res = second ? 1 : 0;
  p.write_f32 = second ? 1 : 0;
  p.write_hl = (l != L - 1 || ctx->nl_loaded) ? 1 : 0;   // the non-local block consumes the last layer's operand
}

static int run_hidden(ehb_ctx* ctx, int l, cudaStream_t stream) {
  ehb_ctx::Hidden& h = *ctx->hidden[l];
  ehb::HiddenLayerParams p;
  fill_hidden_params(ctx, l, p);
  const bool second = (l & 1) == 1;
  if (ctx->gemm_mode == 0) {
    // product path: transposed product, N = 240 activation rows = 10 slots per CTA pair, no pad rows (gcn_umma_t.cu)
    EHB_CUDA(ehb::launch_gcn_hidden_umma_t(ctx->tmX[l & 1], h.tmW, p, ctx->num_sms, ctx->pdl != 0, stream));
    ctx->launches += 1;
  } else if (ctx->gemm_mode == 2) {
    EHB_CUDA(ehb::launch_gcn_hidden_umma(ctx->tmA[l & 1], h.tmB, p, ctx->num_sms, 1, ctx->pdl != 0, stream));
    ctx->launches += 1;
  } else if (ctx->gemm_mode == 3) {
    // the round-1 kernel: activations on the M side, 5 slots + 8 pad rows per 128-row tile (same bits, 6 % more MMA time)
    EHB_CUDA(ehb::launch_gcn_hidden_umma(ctx->tmA[l & 1], h.tmB2, p, ctx->num_sms, 2, ctx->pdl != 0, stream));
    ctx->launches += 1;
  } else {
    const size_t rows = static_cast<size_t>(ctx->n_mtiles) * ehb::TILE_ROWS;
    EHB_CUDA(ctx->mid.ensure(rows * ctx->hid * sizeof(float), true));
    EHB_CUDA(ctx->h_tmp.ensure(rows * 2 * ctx->hid * sizeof(float)));
    const float* xin = second ? ctx->mid.as<float>() : ctx->res.as<float>();
    float* xout = second ? ctx->res.as<float>() : ctx->mid.as<float>();
    EHB_CUDA(ehb::launch_gcn_hidden_simt(xin, h.wcat.as<float>(), ctx->h_tmp.as<float>(), p, h.mod.as<float>(), xout,
                                         stream));
    ctx->launches += 2;
  }
  return 0;
}

static int run_nonlocal(ehb_ctx* ctx, cudaStream_t stream);

static int run_input(ehb_ctx* ctx, int step, const float* x_t, cudaStream_t stream) {
  ehb::InputLayerParams p;
  p.adj = ctx->adj_in;
  p.a01 = ctx->a01.as<float>();
  p.be01 = ctx->be01.as<float>();
  p.cx01 = ctx->cx01.as<float>();
  p.ct01 = ctx->ct01.as<float>();
  p.wx01 = ctx->wx01.as<float>();
  p.mod = ctx->mod_in.as<float>();
  p.bn_scale = ctx->bn_scale_in.as<float>();
  p.bn_shift = ctx->bn_shift_in.as<float>();
  p.vis = ctx->vis.as<uint8_t>();
  p.slot_body = ctx->slot_body.as<int32_t>();
  p.slot_cond = ctx->slot_cond.as<uint8_t>();
  p.img_of_body = ctx->img_of_body.as<int32_t>();
  p.x_t = x_t;
  p.res = ctx->res.as<float>();
  p.out_hl = ctx->act_hl[0].as<__half>();
  p.overflow_flag = ctx->overflow.as<int>();
  p.act_scale = ctx->act_scale;
  p.C = ctx->hid;
  p.n_slots = ctx->n_slots;
  p.step = step;
  p.mask_all = ctx->mask_all;
  // the tensor-pipe joint mix is opt-in (ehb_debug_set_input_mode): both kernels take 0.083-0.087 ms at 640 x 2 slots
  if (ctx->gemm_mode == 1 || !ctx->input_umma) EHB_CUDA(ehb::launch_gcn_input(p, stream));
  else EHB_CUDA(ehb::launch_gcn_input_umma(p, ctx->num_sms, stream));
  ctx->launches += 1;
  return 0;
}

static int run_output(ehb_ctx* ctx, int step, const float* x_t, const float* noise, const float* grad, float* x_prev,
                      float* x0, float* out_cond, float* out_uncond, cudaStream_t stream, float* x0_model = nullptr) {
  ehb::OutputLayerParams p;
  p.adj = ctx->adj_out;
  p.act = ctx->res.as<float>();
  p.wout = ctx->wout.as<float>();
  p.mod = ctx->mod_out.as<float>();
  p.bias = ctx->bias_out.as<float>();
  p.vis = ctx->vis.as<uint8_t>();
  p.img_of_body = ctx->img_of_body.as<int32_t>();
  p.body_slot = ctx->body_slot.as<int32_t>();
  p.x_t = x_t;
  p.noise = noise;
  p.grad = grad;
  p.x_prev = x_prev;
  p.x0 = x0;
  p.out_cond = out_cond;
  p.out_uncond = out_uncond;
  p.x0_model = x0_model;
  p.coef = ctx->coef[step];
  p.kind = ctx->kind;
  p.C = ctx->hid;
  p.n_bodies = ctx->n_bodies;
  p.diffuse_fuse = ctx->diffuse_fuse;
  EHB_CUDA(ehb::launch_gcn_output(p, ctx->num_sms, stream));
  ctx->launches += 1;
  return 0;
}

static int denoise_step_impl(ehb_ctx* ctx, int step, const float* x_t, const float* noise, const float* grad,
                             float* x_prev, float* x0, float* out_cond, float* out_uncond, float* x0_model, void* stream_) {
  if (!ctx || !x_t || !x_prev || !x0) return fail("ehb_denoise_step: null argument");
  if (!ctx->gcn_loaded || ctx->n_bodies <= 0 || ctx->n_img <= 0) return fail("ehb_denoise_step: context not set up");
  if (ctx->max_img_of_body >= ctx->n_img)
    return fail("ehb_denoise_step: a body refers to image " + std::to_string(ctx->max_img_of_body) + " but ehb_set_cond holds " +
                std::to_string(ctx->n_img) + " images");
  if (step < 0 || step >= static_cast<int>(ctx->coef.size()) || step >= ctx->n_steps_cond)
    return fail("ehb_denoise_step: step out of range of the schedule / conditioning tables");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (run_input(ctx, step, x_t, stream)) return 1;
  const int L = static_cast<int>(ctx->hidden.size());
  if (ctx->gemm_mode == 0 && ctx->k1_fused && L >= 2 && L <= ehb::MAX_FUSED_LAYERS) {
    if (run_hidden_fused(ctx, stream)) return 1;
  } else {
    for (int l = 0; l < L; ++l)
      if (run_hidden(ctx, l, stream)) return 1;
  }
  if (ctx->nl_loaded && run_nonlocal(ctx, stream)) return 1;
  return run_output(ctx, step, x_t, noise, grad, x_prev, x0, out_cond, out_uncond, stream, x0_model);
}

int ehb_denoise_step_debug(ehb_ctx* ctx, int step, const float* x_t, const float* noise, const float* grad,
                           float* x_prev, float* x0, float* out_cond, float* out_uncond, void* stream) {
  return denoise_step_impl(ctx, step, x_t, noise, grad, x_prev, x0, out_cond, out_uncond, nullptr, stream);
}

int ehb_denoise_step_ex(ehb_ctx* ctx, int step, const float* x_t, const float* noise, const float* grad, float* x_prev,
                        float* x0, float* x0_model, void* stream) {
  return denoise_step_impl(ctx, step, x_t, noise, grad, x_prev, x0, nullptr, nullptr, x0_model, stream);
}

namespace {
struct EventPair {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t create() {
    cudaError_t e = cudaEventCreate(&e0);
    return e != cudaSuccess ? e : cudaEventCreate(&e1);
  }
  ~EventPair() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};
}  // namespace

/* stage 0 = folded input layer (K2), 1 .. 2*n_blocks = hidden layers (K1), 2*n_blocks + 1 = output layer + sampler
 * update (K3); runs it `iters` times on the context's current activations and returns the mean device time. */
int ehb_time_stage(ehb_ctx* ctx, int stage, int step, const float* x_t, float* x_prev, float* x0, int iters, float* ms,
                   void* stream_) {
  if (!ctx || !ms || !x_t || !x_prev || !x0) return fail("ehb_time_stage: null argument");
  if (!ctx->gcn_loaded || ctx->n_bodies <= 0 || ctx->n_img <= 0) return fail("ehb_time_stage: context not set up");
  if (ctx->max_img_of_body >= ctx->n_img) return fail("ehb_time_stage: a body refers to an image ehb_set_cond does not hold");
  const int L = static_cast<int>(ctx->hidden.size());
  if (stage < 0 || stage > L + 1) return fail("ehb_time_stage: bad stage");
  if (step < 0 || step >= static_cast<int>(ctx->coef.size()) || step >= ctx->n_steps_cond)
    return fail("ehb_time_stage: step out of range");
  if (iters <= 0) return fail("ehb_time_stage: iters must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EventPair ev;   // destroyed on every return path
  EHB_CUDA(ev.create());
  EHB_CUDA(cudaEventRecord(ev.e0, stream));
  for (int i = 0; i < iters; ++i) {
    int rc;
    if (stage == 0) rc = run_input(ctx, step, x_t, stream);
    else if (stage <= L) rc = run_hidden(ctx, stage - 1, stream);
    else rc = run_output(ctx, step, x_t, nullptr, nullptr, x_prev, x0, nullptr, nullptr, stream);
    if (rc) return 1;
  }
  EHB_CUDA(cudaEventRecord(ev.e1, stream));
  EHB_CUDA(cudaEventSynchronize(ev.e1));
  float t = 0.f;
  EHB_CUDA(cudaEventElapsedTime(&t, ev.e0, ev.e1));
  *ms = t / iters;
  return 0;
}

int ehb_denoise_step(ehb_ctx* ctx, int step, const float* x_t, const float* noise, const float* grad, float* x_prev,
                     float* x0, void* stream) {
  return ehb_denoise_step_debug(ctx, step, x_t, noise, grad, x_prev, x0, nullptr, nullptr, stream);
}

int ehb_check_overflow(ehb_ctx* ctx, void* stream_) {
  if (!ctx) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int flag = 0;
  cudaMemcpyAsync(&flag, ctx->overflow.p, sizeof(int), cudaMemcpyDeviceToHost, stream);
  cudaStreamSynchronize(stream);
  if (flag) cudaMemsetAsync(ctx->overflow.p, 0, sizeof(int), stream);
  return flag;
}

int ehb_overflow_flag_async(ehb_ctx* ctx, int32_t* host_flag, void* stream_) {
  if (!ctx || !host_flag) return fail("ehb_overflow_flag_async: null argument");
  EHB_CUDA(cudaMemcpyAsync(host_flag, ctx->overflow.p, sizeof(int), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream_)));
  return 0;
}

int ehb_time_hidden_layer(ehb_ctx* ctx, int layer, int iters, float* ms, void* stream_) {
  if (!ctx || !ms) return fail("ehb_time_hidden_layer: null argument");
  if (!ctx->gcn_loaded || ctx->n_bodies <= 0) return fail("ehb_time_hidden_layer: context not set up");
  if (layer < 1 || layer > static_cast<int>(ctx->hidden.size())) return fail("ehb_time_hidden_layer: bad layer");
  if (iters <= 0) return fail("ehb_time_hidden_layer: iters must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EventPair ev;
  EHB_CUDA(ev.create());
  EHB_CUDA(cudaEventRecord(ev.e0, stream));
  for (int i = 0; i < iters; ++i)
    if (run_hidden(ctx, layer - 1, stream)) return 1;
  EHB_CUDA(cudaEventRecord(ev.e1, stream));
  EHB_CUDA(cudaEventSynchronize(ev.e1));
  float t = 0.f;
  EHB_CUDA(cudaEventElapsedTime(&t, ev.e0, ev.e1));
  *ms = t / iters;
  return 0;
}

int ehb_sampler_update_ex(ehb_ctx* ctx, int step, int n, const float* x_t, const float* x0, const float* noise,
                          const float* grad, float* x_prev, float* x0_out, void* stream_) {
  if (!ctx || !x_t || !x0 || !x_prev) return fail("ehb_sampler_update: null argument");
  if (step < 0 || step >= static_cast<int>(ctx->coef.size())) return fail("ehb_sampler_update: step out of range");
  if (n <= 0) return fail("ehb_sampler_update: n must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  EHB_CUDA(ehb::launch_sampler_update(ctx->coef[step], ctx->kind, x_t, x0, noise, grad, x_prev, x0_out,
                                      static_cast<size_t>(n) * ehb::XDIM, static_cast<cudaStream_t>(stream_)));
  ctx->launches += 1;
  return 0;
}

int ehb_sampler_update(ehb_ctx* ctx, int step, int n, const float* x_t, const float* x0, const float* noise,
                       const float* grad, float* x_prev, void* stream_) {
  return ehb_sampler_update_ex(ctx, step, n, x_t, x0, noise, grad, x_prev, nullptr, stream_);
}

int ehb_debug_get_buffer(ehb_ctx* ctx, int which, void** ptr, uint64_t* bytes) {
  if (!ctx || !ptr || !bytes) return fail("ehb_debug_get_buffer: null argument");
  const DevBuf* b = which == 0 ? &ctx->res : (which == 1 ? &ctx->act_hl[0] : (which == 2 ? &ctx->act_hl[1] : nullptr));
  if (!b) return fail("ehb_debug_get_buffer: which must be 0, 1 or 2");
  *ptr = b->p;
  *bytes = b->bytes;
  return 0;
}

int ehb_rot6d_to_rotmat(ehb_ctx* ctx, const float* x6, int n, float* R, void* stream_) {
  if (!ctx || !x6 || !R) return fail("ehb_rot6d_to_rotmat: null argument");
  if (n < 0) return fail("ehb_rot6d_to_rotmat: negative n");
  EHB_CUDA(cudaSetDevice(ctx->device));
  EHB_CUDA(ehb::launch_rot6d_flat(x6, R, n, static_cast<cudaStream_t>(stream_)));
  ctx->launches += n > 0;
  return 0;
}

int ehb_smpl_load(ehb_ctx* ctx, const ehb_smpl_model* m) {
  if (!ctx || !m) return fail("ehb_smpl_load: null argument");
  if (m->n_verts <= 0 || m->n_betas <= 0 || m->n_betas > 16 || m->n_extra < 0)
    return fail("ehb_smpl_load: need n_verts > 0, 0 < n_betas <= 16, n_extra >= 0");
  if (m->parents[0] >= 0) return fail("ehb_smpl_load: parents[0] must be -1");
  for (int j = 1; j < ehb::NJ; ++j)
    if (m->parents[j] < 0 || m->parents[j] >= j) return fail("ehb_smpl_load: parents must be topologically ordered");
  for (int i = 0; i < m->n_extra; ++i)
    if (m->extra_vertex_ids[i] < 0 || m->extra_vertex_ids[i] >= m->n_verts)
      return fail("ehb_smpl_load: extra vertex id out of range");
  EHB_CUDA(cudaSetDevice(ctx->device));
  const int V = m->n_verts, NB = m->n_betas;
  // J = J_regressor . v_shaped is linear in beta: J_template + J_shapedirs . beta  (smplx lbs.py vertices2joints)
  std::vector<float> jt(ehb::NJ * 3), jsd(static_cast<size_t>(ehb::NJ) * 3 * NB);
  for (int j = 0; j < ehb::NJ; ++j)
    for (int k = 0; k < 3; ++k) {
      double acc = 0;
      for (int v = 0; v < V; ++v) acc += double(m->J_regressor[static_cast<size_t>(j) * V + v]) * m->v_template[v * 3 + k];
      jt[j * 3 + k] = float(acc);
      for (int l = 0; l < NB; ++l) {
        double a2 = 0;
        for (int v = 0; v < V; ++v)
          a2 += double(m->J_regressor[static_cast<size_t>(j) * V + v]) * m->shapedirs[(static_cast<size_t>(v) * 3 + k) * NB + l];
        jsd[(static_cast<size_t>(j) * 3 + k) * NB + l] = float(a2);
      }
    }
  auto up = [&](DevBuf& b, const float* src, size_t n) {
    std::vector<float> h(src, src + n);
    return b.upload(h);
  };
  EHB_CUDA(up(ctx->s_vt, m->v_template, static_cast<size_t>(V) * 3));
  EHB_CUDA(up(ctx->s_sd, m->shapedirs, static_cast<size_t>(V) * 3 * NB));
  EHB_CUDA(up(ctx->s_pd, m->posedirs, static_cast<size_t>(207) * V * 3));
  EHB_CUDA(up(ctx->s_w, m->lbs_weights, static_cast<size_t>(V) * ehb::NJ));
  EHB_CUDA(ctx->s_jt.upload(jt));
  EHB_CUDA(ctx->s_jsd.upload(jsd));
  std::vector<int32_t> ex(m->extra_vertex_ids, m->extra_vertex_ids + m->n_extra);
  if (ex.empty()) ex.push_back(0);
  EHB_CUDA(ctx->s_ex.upload(ex));
  {
    // A operand of the pose-blend GEMM: row r = 3*v + k holds posedirs[:, r] (K = 207 padded to 256) as fp16 hi | lo
    const size_t rows = static_cast<size_t>(V) * 3, rows_pad = (rows + 255) / 256 * 256;
    float mx = 0.f;
    for (size_t i = 0; i < static_cast<size_t>(207) * rows; ++i) mx = std::max(mx, std::fabs(m->posedirs[i]));
    ctx->s_pd_scale = pow2_scale(mx);
    std::vector<float> pt(rows_pad * 256, 0.f);
    for (int k = 0; k < 207; ++k)
      for (size_t r = 0; r < rows; ++r) pt[r * 256 + k] = m->posedirs[static_cast<size_t>(k) * rows + r];
    std::vector<__half> hl;
    split_hl(pt.data(), static_cast<int>(rows_pad), 256, 256, 0, ctx->s_pd_scale, hl);
    EHB_CUDA(ctx->s_pdT_hl.upload(hl));
  }
  ehb::SmplDevice& d = ctx->smpl;
  {
    // compact skinning weights: SMPL vertices follow at most four joints
    std::vector<float> w4(static_cast<size_t>(V) * 4, 0.f);
    std::vector<uint8_t> j4(static_cast<size_t>(V) * 4, 0);
    bool sparse = true;
    for (int v = 0; v < V && sparse; ++v) {
      int q = 0;
      for (int j = 0; j < ehb::NJ; ++j) {
        const float wv = m->lbs_weights[static_cast<size_t>(v) * ehb::NJ + j];
        if (wv != 0.f) {
          if (q == 4) {
            sparse = false;
            break;
          }
          w4[v * 4 + q] = wv;
          j4[v * 4 + q] = static_cast<uint8_t>(j);
          ++q;
        }
      }
    }
    std::vector<float> sdt(static_cast<size_t>(3) * NB * V);
    for (int v = 0; v < V; ++v)
      for (int k = 0; k < 3; ++k)
        for (int l = 0; l < NB; ++l) sdt[(static_cast<size_t>(k) * NB + l) * V + v] = m->shapedirs[(static_cast<size_t>(v) * 3 + k) * NB + l];
    EHB_CUDA(ctx->s_sdT.upload(sdt));
    d.shapedirs_t = ctx->s_sdT.as<float>();
    d.skin_w4 = nullptr;
    d.skin_j4 = nullptr;
    if (sparse) {
      EHB_CUDA(ctx->s_w4.upload(w4));
      EHB_CUDA(ctx->s_j4.upload(j4));
      d.skin_w4 = ctx->s_w4.as<float>();
      d.skin_j4 = ctx->s_j4.as<uint8_t>();
    }
  }
  d.v_template = ctx->s_vt.as<float>();
  d.shapedirs = ctx->s_sd.as<float>();
  d.posedirs = ctx->s_pd.as<float>();
  d.lbs_weights = ctx->s_w.as<float>();
  d.j_template = ctx->s_jt.as<float>();
  d.j_shapedirs = ctx->s_jsd.as<float>();
  d.extra_vids = ctx->s_ex.as<int32_t>();
  for (int j = 0; j < ehb::NJ; ++j) d.parents[j] = m->parents[j];
  d.V = V;
  d.NB = NB;
  d.n_extra = m->n_extra;
  ctx->smpl_loaded = true;
  return 0;
}

static int smpl_run(ehb_ctx* ctx, int n, const float* R, const float* betas, const int32_t* beta_index,
                    const float* transl, float* verts, float* joints, cudaStream_t stream) {
  EHB_CUDA(ctx->sc_A.ensure(static_cast<size_t>(n) * ehb::NJ * 12 * sizeof(float)));
  EHB_CUDA(ctx->sc_j24.ensure(static_cast<size_t>(n) * ehb::NJ * 3 * sizeof(float)));
  EHB_CUDA(ctx->sc_pf.ensure(static_cast<size_t>(n) * 207 * sizeof(float)));
  EHB_CUDA(ehb::launch_smpl_pose(ctx->smpl, R, betas, beta_index, ctx->sc_A.as<float>(), ctx->sc_j24.as<float>(),
                                 ctx->sc_pf.as<float>(), n, stream));
  ctx->launches += 1;
  if (verts && n >= 64 && ctx->smpl.NB <= 10 && ctx->smpl.skin_w4) {
    // pose blend on the tensor cores, body-major: Y[n_pad][3V pad] = pose_feature . posedirs (conv_umma.cu, fp16x3) —
    // A = the pose features of this call, "weights" = posedirs^T (constant) — then skinning
    const long long cols = static_cast<long long>(ctx->smpl.V) * 3;
    const int cols_pad = static_cast<int>((cols + 255) / 256 * 256);      // rows of the posedirs^T operand
    const int n_mtiles = static_cast<int>((n + 255) / 256) * 2;
    const size_t n_pad = static_cast<size_t>(n_mtiles) * 128;
    const float pf_scale = 4096.f;   // |R - I| <= 2  ->  operand magnitude <= 8192
    EHB_CUDA(ctx->sc_pf_hl.ensure(n_pad * 512 * sizeof(__half)));
    EHB_CUDA(ctx->sc_Y.ensure(n_pad * cols_pad * sizeof(float)));
    EHB_CUDA(ctx->s_zero_bias.ensure(static_cast<size_t>(cols_pad) * sizeof(float), true));
    EHB_CUDA(ehb::launch_smpl_pf_operand(ctx->sc_pf.as<float>(), ctx->sc_pf_hl.as<__half>(), n, static_cast<int>(n_pad),
                                         pf_scale, stream));
    const int bn = ehb::conv_gemm_tile_n(cols_pad, n, ctx->num_sms);
    CUtensorMap tA, tB;
    if (make_tmap_f16(&tA, ctx->sc_pf_hl.p, n_pad, 512, 128)) return 1;
    if (make_tmap_f16(&tB, ctx->s_pdT_hl.p, cols_pad, 512, bn / 2)) return 1;
    ehb::ConvGemmParams p{};
    p.bias = ctx->s_zero_bias.as<float>();
    p.out_f32 = ctx->sc_Y.as<float>();
    p.overflow_flag = ctx->overflow.as<int>();
    p.M = n;
    p.acc_scale_inv = 1.f / (ctx->s_pd_scale * pf_scale);
    p.act_scale = 1.f;
    p.K = 256;
    p.Cout = cols_pad;
    p.out_ld = 2 * cols_pad;
    p.n_mtiles = n_mtiles;
    p.n_ntiles = cols_pad / bn;
    EHB_CUDA(ehb::launch_conv_gemm(tA, tB, tA, tB, p, ctx->num_sms, stream));
    EHB_CUDA(ehb::launch_smpl_skin_tiled(ctx->smpl, betas, beta_index, ctx->sc_A.as<float>(), ctx->sc_Y.as<float>(),
                                         static_cast<size_t>(cols_pad), transl, verts, n, stream));
    ctx->launches += 3;
  } else if (verts) {
    EHB_CUDA(ehb::launch_smpl_skin(ctx->smpl, betas, beta_index, ctx->sc_A.as<float>(), ctx->sc_pf.as<float>(), transl,
                                   verts, n, stream));
    ctx->launches += 1;
  }
  if (joints) {
    if (!verts && ctx->smpl.n_extra > 0) return fail("joints need verts (vertex-picked extra joints)");
    EHB_CUDA(ehb::launch_smpl_joints(ctx->smpl, ctx->sc_j24.as<float>(), verts, transl, joints, n, stream));
    ctx->launches += 1;
  }
  return 0;
}

int ehb_smpl_forward(ehb_ctx* ctx, int n, const float* R, const float* betas, const float* transl, float* verts,
                     float* joints, void* stream_) {
  if (!ctx || !R || !betas) return fail("ehb_smpl_forward: null argument");
  if (!ctx->smpl_loaded) return fail("ehb_smpl_forward: call ehb_smpl_load first");
  if (n <= 0) return fail("ehb_smpl_forward: n must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  return smpl_run(ctx, n, R, betas, nullptr, transl, verts, joints, static_cast<cudaStream_t>(stream_));
}

__global__ void relu_copy_kernel(const float* in, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fmaxf(in[i], 0.f);
}

int ehb_gcn_load_nonlocal(ehb_ctx* ctx, const ehb_nonlocal_weights* w) {
  if (!ctx) return fail("ehb_gcn_load_nonlocal: null ctx");
  if (!w) {
    ctx->nl_loaded = false;
    return 0;
  }
  if (!ctx->gcn_loaded) return fail("ehb_gcn_load_nonlocal: call ehb_gcn_load first");
  const int C = ctx->hid, I = w->inter;
  if (I <= 0 || I % 64 != 0 || I * 2 != C) return fail("ehb_gcn_load_nonlocal: inter must be hid / 2 and a multiple of 64");
  EHB_CUDA(cudaSetDevice(ctx->device));
  // theta | phi | g stacked along N and padded to whole 256-column planes; B operand rows = output channels
  const int n_tpg = 3 * I, tpg_planes = (n_tpg + 255) / 256, w_planes = (C + 255) / 256;
  std::vector<float> wt(static_cast<size_t>(tpg_planes) * 256 * C, 0.f), bt(static_cast<size_t>(tpg_planes) * 256, 0.f);
  const float* srcw[3] = {w->theta_w, w->phi_w, w->g_w};
  const float* srcb[3] = {w->theta_b, w->phi_b, w->g_b};
  float m0 = 0.f, m1 = 0.f;
  for (int k = 0; k < 3; ++k)
    for (int r = 0; r < I; ++r) {
      for (int c = 0; c < C; ++c) {
        const float v = srcw[k][static_cast<size_t>(r) * C + c];
        wt[(static_cast<size_t>(k) * I + r) * C + c] = v;
        m0 = std::max(m0, std::fabs(v));
      }
      bt[k * I + r] = srcb[k][r];
    }
  std::vector<float> ww(static_cast<size_t>(w_planes) * 256 * I, 0.f);
  for (int r = 0; r < C; ++r)
    for (int c = 0; c < I; ++c) {
      const float v = w->W_w[static_cast<size_t>(r) * I + c];
      ww[static_cast<size_t>(r) * I + c] = v;
      m1 = std::max(m1, std::fabs(v));
    }
  if (!std::isfinite(m0) || !std::isfinite(m1)) return fail("ehb_gcn_load_nonlocal: non-finite weight");
  ctx->nl_wscale_tpg = pow2_scale(m0);
  ctx->nl_wscale_w = pow2_scale(m1);
  std::vector<__half> hl;
  split_hl(wt.data(), tpg_planes * 256, C, C, 0, ctx->nl_wscale_tpg, hl);
  EHB_CUDA(ctx->nl_wtpg_hl.upload(hl));
  split_hl(ww.data(), w_planes * 256, I, I, 0, ctx->nl_wscale_w, hl);
  EHB_CUDA(ctx->nl_ww_hl.upload(hl));
  EHB_CUDA(ctx->nl_btpg.upload(bt));
  std::vector<float> sc(C), sh(C);
  for (int c = 0; c < C; ++c) {
    const double s = double(w->bn_weight[c]) / std::sqrt(double(w->bn_var[c]) + double(w->bn_eps));
    sc[c] = float(s);
    sh[c] = float(double(w->bn_bias[c]) + (double(w->W_b[c]) - double(w->bn_mean[c])) * s);
  }
  EHB_CUDA(ctx->nl_scale.upload(sc));
  EHB_CUDA(ctx->nl_shift.upload(sh));
  ctx->nl_inter = I;
  ctx->nl_tpg_planes = tpg_planes;
  ctx->nl_w_planes = w_planes;
  ctx->nl_loaded = true;
  return 0;
}

static int run_nonlocal(ehb_ctx* ctx, cudaStream_t stream) {
  const int C = ctx->hid, I = ctx->nl_inter;
  const size_t rows = static_cast<size_t>(ctx->n_mtiles) * ehb::TILE_ROWS;
  const size_t plane = rows * 256;
  EHB_CUDA(ctx->nl_tpg.ensure(static_cast<size_t>(ctx->nl_tpg_planes) * plane * sizeof(float)));
  EHB_CUDA(ctx->nl_wy.ensure(static_cast<size_t>(ctx->nl_w_planes) * plane * sizeof(float)));
  EHB_CUDA(ctx->nl_y_hl.ensure(rows * 2 * I * sizeof(__half), true));
  const int L = static_cast<int>(ctx->hidden.size());
  const CUtensorMap& tA = ctx->tmA[L & 1];   // the last hidden layer wrote act_hl[L & 1]
  ehb::LinearParams lp{};
  lp.overflow_flag = ctx->overflow.as<int>();
  lp.M = static_cast<long long>(rows);
  lp.act_scale = ctx->act_scale;
  lp.n_mtiles = ctx->n_mtiles;
  lp.pts_per_cloud = static_cast<int>(rows);
  lp.K2 = 0;
  // theta | phi | g = x W^T + b
  for (int pl = 0; pl < ctx->nl_tpg_planes; ++pl) {
    CUtensorMap tB;
    if (make_tmap_f16(&tB, ctx->nl_wtpg_hl.as<__half>() + static_cast<size_t>(pl) * 256 * 2 * C, 256, 2 * C, 128)) return 1;
    ehb::LinearParams q = lp;
    q.bias = ctx->nl_btpg.as<float>() + pl * 256;
    q.out_f32 = ctx->nl_tpg.as<float>() + static_cast<size_t>(pl) * plane;
    q.acc_scale_inv = 1.f / (ctx->act_scale * ctx->nl_wscale_tpg);
    q.K1 = C;
    EHB_CUDA(ehb::launch_linear_umma(tA, tB, tA, tB, q, ctx->num_sms, stream));
  }
  ehb::NonLocalParams np{};
  np.tpg = ctx->nl_tpg.as<float>();
  np.wy = ctx->nl_wy.as<float>();
  np.plane_stride = plane;
  np.y_hl = ctx->nl_y_hl.as<__half>();
  np.res = ctx->res.as<float>();
  np.bn_scale = ctx->nl_scale.as<float>();
  np.bn_shift = ctx->nl_shift.as<float>();
  np.overflow_flag = ctx->overflow.as<int>();
  np.act_scale = ctx->act_scale;
  np.C = C;
  np.inter = I;
  np.n_slots = ctx->n_slots;
  EHB_CUDA(ehb::launch_nonlocal_attention(np, stream));
  // W y (bias and BatchNorm folded into the residual kernel's scale / shift)
  CUtensorMap tY;
  if (make_tmap_f16(&tY, ctx->nl_y_hl.p, rows, 2 * I, 128)) return 1;
  for (int pl = 0; pl < ctx->nl_w_planes; ++pl) {
    CUtensorMap tB;
    if (make_tmap_f16(&tB, ctx->nl_ww_hl.as<__half>() + static_cast<size_t>(pl) * 256 * 2 * I, 256, 2 * I, 128)) return 1;
    ehb::LinearParams q = lp;
    q.out_f32 = ctx->nl_wy.as<float>() + static_cast<size_t>(pl) * plane;
    q.acc_scale_inv = 1.f / (ctx->act_scale * ctx->nl_wscale_w);
    q.K1 = I;
    EHB_CUDA(ehb::launch_linear_umma(tY, tB, tY, tB, q, ctx->num_sms, stream));
  }
  EHB_CUDA(ehb::launch_nonlocal_residual(np, stream));
  ctx->launches += ctx->nl_tpg_planes + ctx->nl_w_planes + 2;
  return 0;
}

int ehb_pointnet_load(ehb_ctx* ctx, const ehb_pointnet_weights* w) {
  if (!ctx || !w) return fail("ehb_pointnet_load: null argument");
  if (w->hidden != 256) return fail("ehb_pointnet_load: hidden_dim must be 256");
  if (w->out_dim <= 0) return fail("ehb_pointnet_load: bad out_dim");
  EHB_CUDA(cudaSetDevice(ctx->device));
  const int H = w->hidden, H2 = 2 * H;
  auto up = [&](DevBuf& b, const float* src, size_t n) {
    std::vector<float> h(src, src + n);
    return b.upload(h);
  };
  {  // fc_pos_0 as [3][2H] for coalesced reads
    std::vector<float> t(static_cast<size_t>(3) * H2);
    for (int c = 0; c < H2; ++c)
      for (int d = 0; d < 3; ++d) t[static_cast<size_t>(d) * H2 + c] = w->fc_pos_w[static_cast<size_t>(c) * 3 + d];
    EHB_CUDA(ctx->pn_pos_w.upload(t));
    EHB_CUDA(up(ctx->pn_pos_b, w->fc_pos_b, H2));
    // block_0's shortcut acts on x = fc_pos_0(p), a K = 3 linear map, so shortcut(x) = p . (Wpos^T Ws^T) + bpos . Ws^T is
    // a K = 3 linear map too: it is added in the epilogue of the fc_1 GEMM straight from the point coordinates and the
    // un-activated operand of x is never written or read (respointnet.py:35-36, 88-97)
    std::vector<float> fw(3 * static_cast<size_t>(H)), fb(H);
    for (int r = 0; r < H; ++r) {
      double acc[3] = {0, 0, 0}, accb = 0;
      for (int c = 0; c < H2; ++c) {
        const double ws = w->shortcut_w[0][static_cast<size_t>(r) * H2 + c];
        for (int d = 0; d < 3; ++d) acc[d] += double(w->fc_pos_w[static_cast<size_t>(c) * 3 + d]) * ws;
        accb += double(w->fc_pos_b[c]) * ws;
      }
      for (int d = 0; d < 3; ++d) fw[static_cast<size_t>(d) * H + r] = float(acc[d]);
      fb[r] = float(accb + double(w->fc1_b[0][r]));
    }
    EHB_CUDA(ctx->pn_fold_w.upload(fw));
    EHB_CUDA(ctx->pn_fold_b.upload(fb));
  }
  for (int b = 0; b < 4; ++b) {
    const int kin = b == 0 ? H2 : H;   // per-point K of fc_0 / shortcut (the pooled half is a per-cloud row term)
    float m0 = 0.f, m1 = 0.f;
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < kin; ++c) {
        m0 = std::max(m0, std::fabs(w->fc0_w[b][static_cast<size_t>(r) * H2 + c]));
        m1 = std::max(m1, std::fabs(w->shortcut_w[b][static_cast<size_t>(r) * H2 + c]));
      }
    for (size_t i = 0; i < static_cast<size_t>(H) * H; ++i) m1 = std::max(m1, std::fabs(w->fc1_w[b][i]));
    ctx->pn_wscale[b][0] = pow2_scale(m0);
    ctx->pn_wscale[b][1] = pow2_scale(m1);   // fc_1 and shortcut accumulate into one TMEM accumulator: one scale
    std::vector<__half> hl;
    split_hl(w->fc0_w[b], H, kin, H2, 0, ctx->pn_wscale[b][0], hl);
    EHB_CUDA(ctx->pn_fc0[b].upload(hl));
    split_hl(w->fc1_w[b], H, H, H, 0, ctx->pn_wscale[b][1], hl);
    EHB_CUDA(ctx->pn_fc1[b].upload(hl));
    split_hl(w->shortcut_w[b], H, kin, H2, 0, ctx->pn_wscale[b][1], hl);
    EHB_CUDA(ctx->pn_sc[b].upload(hl));
    EHB_CUDA(up(ctx->pn_b0[b], w->fc0_b[b], H));
    EHB_CUDA(up(ctx->pn_b1[b], w->fc1_b[b], H));
    if (b > 0) {  // pooled halves, transposed to [K=H][N=H] for sgemm_nn
      std::vector<float> t0(static_cast<size_t>(H) * H), ts(static_cast<size_t>(H) * H);
      for (int r = 0; r < H; ++r)
        for (int c = 0; c < H; ++c) {
          t0[static_cast<size_t>(c) * H + r] = w->fc0_w[b][static_cast<size_t>(r) * H2 + H + c];
          ts[static_cast<size_t>(c) * H + r] = w->shortcut_w[b][static_cast<size_t>(r) * H2 + H + c];
        }
      EHB_CUDA(ctx->pn_w0b_t[b].upload(t0));
      EHB_CUDA(ctx->pn_wsb_t[b].upload(ts));
    }
  }
  {
    std::vector<float> t(static_cast<size_t>(H) * w->out_dim);
    for (int r = 0; r < w->out_dim; ++r)
      for (int c = 0; c < H; ++c) t[static_cast<size_t>(c) * w->out_dim + r] = w->fc_c_w[static_cast<size_t>(r) * H + c];
    EHB_CUDA(ctx->pn_fcc_t.upload(t));
    EHB_CUDA(up(ctx->pn_fcc_b, w->fc_c_b, w->out_dim));
  }
  ctx->pn_hidden = H;
  ctx->pn_out = w->out_dim;
  ctx->pn_loaded = true;
  return 0;
}

int ehb_pointnet_forward(ehb_ctx* ctx, const float* pts, int n_clouds, int n_pts, float* feats, void* stream_) {
  if (!ctx || !pts || !feats) return fail("ehb_pointnet_forward: null argument");
  if (!ctx->pn_loaded) return fail("ehb_pointnet_forward: call ehb_pointnet_load first");
  if (n_clouds <= 0 || n_pts <= 0) return fail("ehb_pointnet_forward: n_clouds and n_pts must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int H = ctx->pn_hidden, H2 = 2 * H;
  const long long M = static_cast<long long>(n_clouds) * n_pts;
  const int n_mtiles = static_cast<int>((M + 255) / 256) * 2;
  const size_t rows = static_cast<size_t>(n_mtiles) * 128;
  const float act_scale = ctx->act_scale;
  for (int i = 0; i < 2; ++i) {
    EHB_CUDA(ctx->pn_x[i].ensure(rows * 2 * H2 * sizeof(__half), true));
    EHB_CUDA(ctx->pn_y[i].ensure(rows * 2 * H * sizeof(__half), true));
  }
  EHB_CUDA(ctx->pn_h.ensure(rows * 2 * H * sizeof(__half), true));
  const size_t pc = static_cast<size_t>(n_clouds) * H;
  EHB_CUDA(ctx->pn_pool.ensure(pc * sizeof(int)));
  EHB_CUDA(ctx->pn_pooled.ensure(pc * sizeof(float)));
  EHB_CUDA(ctx->pn_pooled_relu.ensure(pc * sizeof(float)));
  EHB_CUDA(ctx->pn_row0.ensure(pc * sizeof(float)));
  EHB_CUDA(ctx->pn_rows.ensure(pc * sizeof(float)));
  int* ovf = ctx->overflow.as<int>();

  EHB_CUDA(ehb::launch_pointnet_pos(pts, ctx->pn_pos_w.as<float>(), ctx->pn_pos_b.as<float>(), nullptr,
                                    ctx->pn_x[1].as<__half>(), M, H2, act_scale, ovf, stream));
  ctx->launches += 1;
  // current block input: plain / relu'd fp16 operands with K = kin
  __half* x_plain = ctx->pn_x[0].as<__half>();
  __half* x_relu = ctx->pn_x[1].as<__half>();
  __half* y_plain = ctx->pn_y[0].as<__half>();
  __half* y_relu = ctx->pn_y[1].as<__half>();
  for (int b = 0; b < 4; ++b) {
    const int kin = b == 0 ? H2 : H;
    const bool last = b == 3;
    CUtensorMap tA_relu, tA_plain, tA_h, tB0, tB1, tBs;
    if (make_tmap_f16(&tA_relu, x_relu, rows, 2 * kin, 128)) return 1;
    if (make_tmap_f16(&tA_plain, x_plain, rows, 2 * kin, 128)) return 1;
    if (make_tmap_f16(&tA_h, ctx->pn_h.p, rows, 2 * H, 128)) return 1;
    if (make_tmap_f16(&tB0, ctx->pn_fc0[b].p, H, 2 * kin, 128)) return 1;
    if (make_tmap_f16(&tB1, ctx->pn_fc1[b].p, H, 2 * H, 128)) return 1;
    if (make_tmap_f16(&tBs, ctx->pn_sc[b].p, H, 2 * kin, 128)) return 1;
    const float* row0 = nullptr;
    const float* rows_s = nullptr;
    if (b > 0) {
      // per-cloud rows from the pooled half of the concatenated input (respointnet.py:38-40):
      //   row0 = relu(pooled) W0[:, H:]^T + b0      rows_s = pooled Ws[:, H:]^T + b1
      const size_t n = pc;
      fill_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(ctx->pn_row0.as<float>(), ctx->pn_b0[b].as<float>(), n_clouds, H);
      fill_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(ctx->pn_rows.as<float>(), ctx->pn_b1[b].as<float>(), n_clouds, H);
      EHB_CUDA(cudaGetLastError());
      if (ctx_sgemm(ctx, ctx->pn_pooled_relu.as<float>(), ctx->pn_w0b_t[b].as<float>(), ctx->pn_row0.as<float>(), n_clouds, H,
                    H, H, H, H, 1, stream))
        return 1;
      if (ctx_sgemm(ctx, ctx->pn_pooled.as<float>(), ctx->pn_wsb_t[b].as<float>(), ctx->pn_rows.as<float>(), n_clouds, H, H,
                    H, H, H, 1, stream))
        return 1;
      row0 = ctx->pn_row0.as<float>();
      rows_s = ctx->pn_rows.as<float>();
      ctx->launches += 2;
    }
    ehb::LinearParams p{};
    p.overflow_flag = ovf;
    p.M = M;
    p.act_scale = act_scale;
    p.n_mtiles = n_mtiles;
    p.pts_per_cloud = n_pts;
    // h = fc_0(relu(x))  -> only relu(h) is consumed (by fc_1)
    p.bias = b == 0 ? ctx->pn_b0[b].as<float>() : nullptr;
    p.rowvec = row0;
    p.out_hl_relu = ctx->pn_h.as<__half>();
    p.acc_scale_inv = 1.f / (act_scale * ctx->pn_wscale[b][0]);
    p.K1 = kin;
    p.K2 = 0;
    EHB_CUDA(ehb::launch_linear_umma(tA_relu, tB0, tA_relu, tB0, p, ctx->num_sms, stream));
    // net = shortcut(x) + fc_1(relu(h))  -> next block's operands (plain + relu) and the per-cloud max
    EHB_CUDA(ehb::launch_pool_init(ctx->pn_pool.as<int>(), static_cast<int>(pc), stream));
    ehb::LinearParams q{};
    q.overflow_flag = ovf;
    q.M = M;
    q.act_scale = act_scale;
    q.n_mtiles = n_mtiles;
    q.pts_per_cloud = n_pts;
    q.bias = b == 0 ? ctx->pn_fold_b.as<float>() : nullptr;   // block_0: fc_1 bias + the folded shortcut's constant
    q.rowvec = rows_s;
    q.pts = b == 0 ? pts : nullptr;
    q.ptw = b == 0 ? ctx->pn_fold_w.as<float>() : nullptr;
    q.out_hl = last ? nullptr : y_plain;
    q.out_hl_relu = last ? nullptr : y_relu;
    q.pool = ctx->pn_pool.as<int>();
    q.acc_scale_inv = 1.f / (act_scale * ctx->pn_wscale[b][1]);
    q.K1 = H;
    q.K2 = b == 0 ? 0 : kin;   // block_0's shortcut is the K = 3 epilogue term above
    EHB_CUDA(ehb::launch_linear_umma(tA_h, tB1, tA_plain, tBs, q, ctx->num_sms, stream));
    EHB_CUDA(ehb::launch_pool_decode(ctx->pn_pool.as<int>(), ctx->pn_pooled.as<float>(), static_cast<int>(pc), stream));
    relu_copy_kernel<<<static_cast<unsigned>((pc + 255) / 256), 256, 0, stream>>>(ctx->pn_pooled.as<float>(),
                                                                                 ctx->pn_pooled_relu.as<float>(),
                                                                                 static_cast<int>(pc));
    EHB_CUDA(cudaGetLastError());
    ctx->launches += 5;
    // the block's outputs are the next block's inputs (K = H from now on)
    x_plain = y_plain;
    x_relu = y_relu;
    y_plain = (y_plain == ctx->pn_y[0].as<__half>()) ? ctx->pn_x[0].as<__half>() : ctx->pn_y[0].as<__half>();
    y_relu = (y_relu == ctx->pn_y[1].as<__half>()) ? ctx->pn_x[1].as<__half>() : ctx->pn_y[1].as<__half>();
  }
  // c = fc_c(relu(max_pool(net)))  (respointnet.py:55-57)
  {
    const size_t n = static_cast<size_t>(n_clouds) * ctx->pn_out;
    fill_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(feats, ctx->pn_fcc_b.as<float>(), n_clouds, ctx->pn_out);
    EHB_CUDA(cudaGetLastError());
    if (ctx_sgemm(ctx, ctx->pn_pooled_relu.as<float>(), ctx->pn_fcc_t.as<float>(), feats, n_clouds, ctx->pn_out, H, H, ctx->pn_out,
                  ctx->pn_out, 1, stream))
      return 1;
    ctx->launches += 1;
  }
  return 0;
}

int ehb_maxpool3x3s2_nhwc(ehb_ctx* ctx, const float* in, int n, int h, int w, int c, float* out, void* stream_) {
  if (!ctx || !in || !out) return fail("ehb_maxpool3x3s2_nhwc: null argument");
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 4 != 0)
    return fail("ehb_maxpool3x3s2_nhwc: need positive n, h, w and a channel count that is a multiple of 4");
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail("ehb_maxpool3x3s2_nhwc: pointers must be 16-byte aligned");
  EHB_CUDA(cudaSetDevice(ctx->device));
  EHB_CUDA(ehb::launch_maxpool3x3s2_nhwc(in, out, n, h, w, c, static_cast<cudaStream_t>(stream_)));
  ctx->launches += 1;
  return 0;
}

int ehb_cond_inputs(ehb_ctx* ctx, const float* kp2d, const float* scene_feat, const float* transl_feat, const float* img_feat,
                    const float* fx, const float* box_center, const float* box_size, const float* cam_cx, const float* cam_cy,
                    int n_img, int scene_dim, int transl_dim, int img_dim, int with_focal_length, int with_bbox_info,
                    int with_cam_center, const int32_t* openpose_to_smpl, float fx_norm_coeff, uint8_t* vis, float* rest,
                    float* ctx_full, void* stream_) {
  if (!ctx || !kp2d || !scene_feat || !transl_feat || !img_feat || !openpose_to_smpl || !vis || !rest || !ctx_full)
    return fail("ehb_cond_inputs: null argument");
  if (n_img <= 0 || scene_dim <= 0 || transl_dim <= 0 || img_dim <= 0) return fail("ehb_cond_inputs: sizes must be positive");
  if ((with_focal_length || with_bbox_info || with_cam_center) && !fx) return fail("ehb_cond_inputs: fx is required by the camera flags");
  if (with_bbox_info && (!box_center || !box_size)) return fail("ehb_cond_inputs: box_center / box_size required (with_bbox_info)");
  if (with_cam_center && (!cam_cx || !cam_cy)) return fail("ehb_cond_inputs: cam_cx / cam_cy required (with_cam_center)");
  for (int j = 0; j < ehb::NJ; ++j)
    if (openpose_to_smpl[j] < 0 || openpose_to_smpl[j] >= 25) return fail("ehb_cond_inputs: openpose_to_smpl entries must be in [0, 25)");
  EHB_CUDA(cudaSetDevice(ctx->device));
  EHB_CUDA(ehb::launch_cond_inputs(kp2d, scene_feat, transl_feat, img_feat, fx, box_center, box_size, cam_cx, cam_cy, n_img,
                                   scene_dim, transl_dim, img_dim, with_focal_length, with_bbox_info, with_cam_center,
                                   openpose_to_smpl, fx_norm_coeff, vis, rest, ctx_full, static_cast<cudaStream_t>(stream_)));
  ctx->launches += 1;
  return 0;
}

int ehb_project_joints(ehb_ctx* ctx, const float* joints, const float* transl, const float* fx, const float* cam_cx,
                       const float* cam_cy, const int32_t* img_of_body, int n_bodies, int n_joints, float fx_norm_coeff,
                       float default_focal, float* kp3d_full, float* kp2d, float* focal_out, float* center_out, void* stream_) {
  if (!ctx || !joints || !transl || !kp3d_full || !kp2d || !focal_out || !center_out) return fail("ehb_project_joints: null argument");
  if (fx && (!cam_cx || !cam_cy)) return fail("ehb_project_joints: cam_cx / cam_cy required together with fx");
  if (n_bodies <= 0 || n_joints <= 0) return fail("ehb_project_joints: sizes must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  EHB_CUDA(ehb::launch_project_joints(joints, transl, fx, cam_cx, cam_cy, img_of_body, n_bodies, n_joints, fx_norm_coeff,
                                      default_focal, kp3d_full, kp2d, focal_out, center_out, static_cast<cudaStream_t>(stream_)));
  ctx->launches += 1;
  return 0;
}

int ehb_scene_crop(ehb_ctx* ctx, const float* verts, int n_bodies, int n_verts, const float* scene, int n_pts,
                   const int32_t* img_of_body, uint8_t* mask, int32_t* count, float* bbox, void* stream_) {
  if (!ctx || !verts || !scene || !mask || !count) return fail("ehb_scene_crop: null argument");
  if (n_bodies <= 0 || n_verts <= 0 || n_pts <= 0) return fail("ehb_scene_crop: sizes must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  EHB_CUDA(ehb::launch_scene_crop(verts, n_bodies, n_verts, scene, n_pts, img_of_body, mask, count, bbox,
                                  static_cast<cudaStream_t>(stream_)));
  ctx->launches += 1;
  return 0;
}

int ehb_procrustes_align(ehb_ctx* ctx, const float* s1, const float* s2, const float* mask, int n_problems, int n_points,
                         float* s1_hat, float* err, void* stream_) {
  if (!ctx || !s1 || !s2) return fail("ehb_procrustes_align: null argument");
  if (!s1_hat && !err) return fail("ehb_procrustes_align: at least one of s1_hat / err must be given");
  if (n_problems < 0 || n_points <= 0) return fail("ehb_procrustes_align: bad sizes");
  EHB_CUDA(cudaSetDevice(ctx->device));
  EHB_CUDA(ehb::launch_procrustes(s1, s2, mask, n_problems, n_points, s1_hat, err, static_cast<cudaStream_t>(stream_)));
  ctx->launches += n_problems > 0;
  return 0;
}

int ehb_eval_metrics(ehb_ctx* ctx, const float* pred_joints, const float* pred_verts, const float* transl,
                     const float* gt_joints, const float* gt_verts, const float* focal, const float* cam_cx,
                     const float* cam_cy, int n_img, int n_samples, int n_joints, int n_verts, uint8_t* joint_vis,
                     uint8_t* vert_vis, float* errors, float* diversity, void* stream_) {
  if (!ctx || !pred_joints || !pred_verts || !transl || !gt_joints || !gt_verts || !focal || !cam_cx || !cam_cy ||
      !joint_vis || !vert_vis || !errors || !diversity)
    return fail("ehb_eval_metrics: null argument");
  if (n_img <= 0 || n_samples <= 0 || n_joints <= 0 || n_joints > 256 || n_verts <= 0)
    return fail("ehb_eval_metrics: need n_img, n_samples, n_verts > 0 and 0 < n_joints <= 256");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EHB_CUDA(ehb::launch_vis_mask(gt_joints, focal, cam_cx, cam_cy, joint_vis, n_img, n_joints, 1920.f, 1080.f, stream));
  EHB_CUDA(ehb::launch_vis_mask(gt_verts, focal, cam_cx, cam_cy, vert_vis, n_img, n_verts, 1920.f, 1080.f, stream));
  EHB_CUDA(ehb::launch_pose_errors(pred_joints, pred_verts, transl, gt_joints, gt_verts, joint_vis, vert_vis, errors, n_img,
                                   n_samples, n_joints, n_verts, stream));
  EHB_CUDA(ehb::launch_diversity(pred_joints, joint_vis, diversity, n_img, n_samples, n_joints, stream));
  ctx->launches += 4;
  return 0;
}

int ehb_resnet_load(ehb_ctx* ctx, const ehb_resnet_weights* w) {
  if (!ctx || !w || !w->convs) return fail("ehb_resnet_load: null argument");
  int expect = 1;
  for (int s = 0; s < 4; ++s) {
    if (w->blocks[s] <= 0) return fail("ehb_resnet_load: blocks must be positive");
    expect += 3 * w->blocks[s] + 1;
  }
  if (w->n_convs != expect) return fail("ehb_resnet_load: n_convs does not match the block counts");
  EHB_CUDA(cudaSetDevice(ctx->device));
  for (auto* c : ctx->rn_convs) delete c;
  ctx->rn_convs.clear();
  ctx->rn_loaded = false;
  for (int i = 0; i < w->n_convs; ++i) {
    const ehb_conv_bn& c = w->convs[i];
    if (c.cout <= 0 || c.cout % 64 || c.cin <= 0 || c.kh <= 0 || c.kw <= 0 || c.stride <= 0 || c.pad < 0)
      return fail("ehb_resnet_load: bad convolution shape (cout must be a multiple of 64)");
    if (i > 0 && c.cin % 64) return fail("ehb_resnet_load: cin must be a multiple of 64 after the stem");
    auto* pl = new ehb_ctx::ConvPlan();
    ctx->rn_convs.push_back(pl);
    pl->cout = c.cout; pl->cin = c.cin; pl->kh = c.kh; pl->kw = c.kw; pl->stride = c.stride; pl->pad = c.pad;
    if (i == 0) {
      // The 7x7 / stride 2 / pad 3 stem on 3 channels (models/resnet.py:109) == a 4x4 / stride 1 / pad 2 convolution on the
      // space-to-depth image (16 channels: (dy*2 + dx)*3 + c, 12 used): input row 2*oy - 3 + ky = 2*(oy - 2 + ty) + dy with
      // ky = 2*ty + dy - 1 (taps with ky or kx outside 0..6 get zero weights).  The implicit-GEMM kernel then reads the image
      // through TMA boxes; no im2col matrix (617 MB written and re-read for 64 images) exists.
      if (c.kh != 7 || c.kw != 7 || c.cin != 3 || c.stride != 2 || c.pad != 3) return fail("ehb_resnet_load: unexpected stem");
      pl->Kp = 256;
      std::vector<float> wf(static_cast<size_t>(c.cout) * 256, 0.f), bias(c.cout);
      float maxabs = 0.f;
      for (int co = 0; co < c.cout; ++co) {
        const double s = double(c.bn_weight[co]) / std::sqrt(double(c.bn_var[co]) + double(c.bn_eps));
        bias[co] = float(double(c.bn_bias[co]) - double(c.bn_mean[co]) * s);
        for (int ty = 0; ty < 4; ++ty)
          for (int tx = 0; tx < 4; ++tx)
            for (int dy = 0; dy < 2; ++dy)
              for (int dx = 0; dx < 2; ++dx) {
                const int ky = 2 * ty + dy - 1, kx = 2 * tx + dx - 1;
                if (ky < 0 || ky > 6 || kx < 0 || kx > 6) continue;
                for (int ci = 0; ci < 3; ++ci) {
                  const float v = float(double(c.weight[((static_cast<size_t>(co) * 3 + ci) * 7 + ky) * 7 + kx]) * s);
                  wf[static_cast<size_t>(co) * 256 + (ty * 4 + tx) * 16 + (dy * 2 + dx) * 3 + ci] = v;
                  maxabs = std::max(maxabs, std::fabs(v));
                }
              }
      }
      if (!std::isfinite(maxabs)) return fail("ehb_resnet_load: non-finite weight");
      // what the kernel executes: 4 vertical taps over 64-half windows (4 horizontal taps x 16 channels), see stem_s2d_kernel
      pl->cin = 64; pl->kh = 4; pl->kw = 1; pl->stride = 1; pl->pad = 2;
      pl->stem = true;
      pl->w_scale = pow2_scale(maxabs);
      pl->wf.swap(wf);
      pl->bias_h.swap(bias);
      continue;
    }
    const int K = c.kh * c.kw * c.cin;
    pl->Kp = (K + 63) / 64 * 64;
    // fold BatchNorm (eval): y = conv(x; W) * s + (beta - mean * s),  s = gamma / sqrt(var + eps)
    // GEMM weight row co: k = (ky*kw + kx)*cin + ci  (the im2col order of resnet_ops.cu)
    std::vector<float> wf(static_cast<size_t>(c.cout) * pl->Kp, 0.f), bias(c.cout);
    float maxabs = 0.f;
    for (int co = 0; co < c.cout; ++co) {
      const double s = double(c.bn_weight[co]) / std::sqrt(double(c.bn_var[co]) + double(c.bn_eps));
      bias[co] = float(double(c.bn_bias[co]) - double(c.bn_mean[co]) * s);
      for (int ci = 0; ci < c.cin; ++ci)
        for (int t = 0; t < c.kh * c.kw; ++t) {
          const float v = float(double(c.weight[(static_cast<size_t>(co) * c.cin + ci) * c.kh * c.kw + t]) * s);
          wf[static_cast<size_t>(co) * pl->Kp + static_cast<size_t>(t) * c.cin + ci] = v;
          maxabs = std::max(maxabs, std::fabs(v));
        }
    }
    if (!std::isfinite(maxabs)) return fail("ehb_resnet_load: non-finite weight");
    pl->w_scale = pow2_scale(maxabs);
    pl->wf.swap(wf);
    pl->bias_h.swap(bias);
  }
  // conv3 and the projection shortcut of a stage's first block accumulate into one TMEM accumulator: one operand scale
  // for both, one summed bias (the projection's own bias buffer is unused)
  {
    int ci = 1;
    for (int s = 0; s < 4; ++s)
      for (int b = 0; b < w->blocks[s]; ++b) {
        if (b == 0) {
          auto& c3 = *ctx->rn_convs[ci + 2];
          auto& cd = *ctx->rn_convs[ci + 3];
          if (c3.cout != cd.cout) return fail("ehb_resnet_load: downsample / conv3 channel mismatch");
          const float sc = std::min(c3.w_scale, cd.w_scale);
          c3.w_scale = cd.w_scale = sc;
          for (int co = 0; co < c3.cout; ++co) c3.bias_h[co] += cd.bias_h[co];
        }
        ci += b == 0 ? 4 : 3;
      }
  }
  for (auto* pl : ctx->rn_convs) {
    std::vector<__half> hl;
    split_hl(pl->wf.data(), pl->cout, pl->Kp, pl->Kp, 0, pl->w_scale, hl);
    EHB_CUDA(pl->w_hl.upload(hl));
    EHB_CUDA(pl->bias.upload(pl->bias_h));
    std::vector<float>().swap(pl->wf);
  }
  for (int s = 0; s < 4; ++s) ctx->rn_blocks[s] = w->blocks[s];
  ctx->rn_loaded = true;
  return 0;
}

// one convolution as a GEMM: A [rows_pad][2*Kp] hi/lo (activation or im2col matrix) -> out [rows_pad][2*cout]
static int rn_gemm(ehb_ctx* ctx, const ehb_ctx::ConvPlan& c, const __half* A, long long rows, const __half* res, __half* out,
                   int relu, cudaStream_t stream, const ehb_ctx::ConvPlan* c2 = nullptr, const __half* A2 = nullptr) {
  const int n_mtiles = static_cast<int>((rows + 255) / 256) * 2;
  const size_t rows_pad = static_cast<size_t>(n_mtiles) * 128;
  const int kb_total = (c.Kp + (c2 ? c2->Kp : 0)) / 64;
  // 1x1 convolutions with up to `rn_kc_1x1` k-blocks stay in one accumulation (and may use the 256-wide tile): they are
  // epilogue-bound, and their chains are short (<= 16 main MMAs)
  const bool chunked = ctx->rn_kc > 0 && kb_total > std::max(ctx->rn_kc, c.kh == 1 ? ctx->rn_kc_1x1 : 0);
  const int bn = ehb::conv_gemm_tile_n(c.cout, rows, ctx->num_sms, chunked ? 128 : 256);
  CUtensorMap tA, tB;
  if (make_tmap_f16(&tA, A, rows_pad, 2 * static_cast<uint64_t>(c.Kp), 128)) return 1;
  if (make_tmap_f16(&tB, c.w_hl.p, c.cout, 2 * static_cast<uint64_t>(c.Kp), bn / 2)) return 1;
  CUtensorMap tA2 = tA, tB2 = tB;
  if (c2) {
    if (make_tmap_f16(&tA2, A2, rows_pad, 2 * static_cast<uint64_t>(c2->Kp), 128)) return 1;
    if (make_tmap_f16(&tB2, c2->w_hl.p, c2->cout, 2 * static_cast<uint64_t>(c2->Kp), bn / 2)) return 1;
  }
  ehb::ConvGemmParams p{};
  p.bias = c.bias.as<float>();
  p.res_hl = res;
  p.out_hl = out;
  p.out_f32 = nullptr;
  p.overflow_flag = ctx->overflow.as<int>();
  p.M = rows;
  p.acc_scale_inv = 1.f / (ctx->rn_act_scale * c.w_scale);
  p.act_scale = ctx->rn_act_scale;
  p.K = c.Kp;
  p.K2 = c2 ? c2->Kp : 0;
  p.Cout = c.cout;
  p.out_ld = 2 * c.cout;
  p.n_mtiles = n_mtiles;
  p.n_ntiles = c.cout / bn;
  p.relu = relu;
  p.kc = chunked ? ctx->rn_kc : 0;
  EHB_CUDA(ehb::launch_conv_gemm(tA, tB, tA2, tB2, p, ctx->num_sms, stream));
  ctx->launches += 1;
  return 0;
}

// a KHxKW convolution straight from the NHWC activation x [n][H][W][2*cin]: implicit GEMM, no im2col matrix
static int rn_gemm_implicit(ehb_ctx* ctx, const ehb_ctx::ConvPlan& c, const __half* x, int n, int H, int W, int Ho, int Wo,
                            __half* out, int relu, cudaStream_t stream) {
  if (Wo > 128 || c.cin % 64) return fail("rn_gemm_implicit: unsupported shape");
  int th = 128 / Wo, nb = 1, tpi;
  if (th >= Ho) {   // whole images per tile
    th = Ho;
    nb = std::max(1, 128 / (Ho * Wo));
    tpi = 1;
  } else {
    tpi = (Ho + th - 1) / th;
  }
  const long long rows = static_cast<long long>(n) * Ho * Wo;
  int n_tiles = ((n + nb - 1) / nb) * tpi;
  n_tiles = (n_tiles + 1) / 2 * 2;
  const bool chunked = ctx->rn_kc > 0 && c.Kp / 64 > ctx->rn_kc;
  const int bn = ehb::conv_gemm_tile_n(c.cout, static_cast<long long>(n_tiles) * 128, ctx->num_sms, chunked ? 128 : 256);
  CUtensorMap tA, tB;
  if (c.stem) {   // x = planes [2][n][H][W + 4][16]: both planes are "images" of one windowed map (lo = image n_img + i)
    if (make_tmap_f16_nhwc(&tA, x, 2 * static_cast<uint64_t>(n), H, Wo, 64, nb, th, Wo, 1, 16, W + 4)) return 1;
  } else if (make_tmap_f16_nhwc(&tA, x, n, H, W, 2 * static_cast<uint64_t>(c.cin), nb, th, Wo, c.stride)) return 1;
  if (make_tmap_f16(&tB, c.w_hl.p, c.cout, 2 * static_cast<uint64_t>(c.Kp), bn / 2)) return 1;
  ehb::ConvGemmParams p{};
  p.bias = c.bias.as<float>();
  p.out_hl = out;
  p.overflow_flag = ctx->overflow.as<int>();
  p.M = rows;
  p.acc_scale_inv = 1.f / (ctx->rn_act_scale * c.w_scale);
  p.act_scale = ctx->rn_act_scale;
  p.K = c.Kp;
  p.Cout = c.cout;
  p.out_ld = 2 * c.cout;
  p.n_mtiles = n_tiles;
  p.n_ntiles = c.cout / bn;
  p.relu = relu;
  p.kc = ctx->rn_kc;
  p.implicit = 1;
  p.pad_w = c.stem ? 0 : c.pad;
  p.lo_plane = c.stem ? n : 0;
  p.Cin = c.cin; p.kw = c.kw; p.pad = c.pad; p.stride = c.stride;
  p.Ho = Ho; p.Wo = Wo; p.th = th; p.nb = nb; p.tiles_per_img = tpi; p.n_img = n;
  EHB_CUDA(ehb::launch_conv_gemm(tA, tB, tA, tB, p, ctx->num_sms, stream));
  ctx->launches += 1;
  return 0;
}

int ehb_debug_set_resnet_mode(ehb_ctx* ctx, int implicit_gemm) {
  if (!ctx) return fail("null ctx");
  ctx->rn_implicit = implicit_gemm ? 1 : 0;
  return 0;
}

int ehb_debug_set_pdl(ehb_ctx* ctx, int on) {
  if (!ctx) return fail("null ctx");
  ctx->pdl = on ? 1 : 0;
  return 0;
}

int ehb_debug_set_k1_fused(ehb_ctx* ctx, int on) {
  if (!ctx) return fail("null ctx");
  ctx->k1_fused = on ? 1 : 0;
  return 0;
}

int ehb_debug_set_input_mode(ehb_ctx* ctx, int umma) {
  if (!ctx) return fail("null ctx");
  ctx->input_umma = umma ? 1 : 0;
  return 0;
}

int ehb_debug_set_conv_kc(ehb_ctx* ctx, int kc) {
  if (!ctx) return fail("null ctx");
  if (kc < 0) return fail("ehb_debug_set_conv_kc: kc must be >= 0");
  ctx->rn_kc = kc % 100;
  ctx->rn_kc_1x1 = kc >= 100 ? kc / 100 : ctx->rn_kc;   // bring-up: kc = 100 * (1x1 threshold) + chunk length
  return 0;
}

int ehb_debug_gemm_hl(ehb_ctx* ctx, const float* a, const float* w, int m, int n, int k, float a_scale, float w_scale,
                      int kc, float* out) {
  if (!ctx || !a || !w || !out) return fail("ehb_debug_gemm_hl: null argument");
  if (m <= 0 || n <= 0 || n % 64 || k <= 0 || k % 64) return fail("ehb_debug_gemm_hl: need m > 0, n % 64 == 0, k % 64 == 0");
  EHB_CUDA(cudaSetDevice(ctx->device));
  const int n_mtiles = (m + 255) / 256 * 2;
  const size_t rows_pad = static_cast<size_t>(n_mtiles) * 128;
  std::vector<__half> ah, wh;
  split_hl(a, m, k, k, 0, a_scale, ah);
  ah.resize(rows_pad * 2 * k, __half());
  split_hl(w, n, k, k, 0, w_scale, wh);
  std::vector<float> zero(n, 0.f);
  DevBuf dA, dW, dB, dO;
  EHB_CUDA(dA.upload(ah));
  EHB_CUDA(dW.upload(wh));
  EHB_CUDA(dB.upload(zero));
  EHB_CUDA(dO.ensure(static_cast<size_t>(m) * n * sizeof(float)));
  const bool chunked = kc > 0 && k / 64 > kc;
  const int bn = ehb::conv_gemm_tile_n(n, m, ctx->num_sms, chunked ? 128 : 256);
  CUtensorMap tA, tB;
  if (make_tmap_f16(&tA, dA.p, rows_pad, 2 * static_cast<uint64_t>(k), 128)) return 1;
  if (make_tmap_f16(&tB, dW.p, n, 2 * static_cast<uint64_t>(k), bn / 2)) return 1;
  ehb::ConvGemmParams p{};
  p.bias = dB.as<float>();
  p.out_f32 = dO.as<float>();
  p.overflow_flag = ctx->overflow.as<int>();
  p.M = m;
  p.acc_scale_inv = 1.f / (a_scale * w_scale);
  p.act_scale = a_scale;
  p.K = k;
  p.Cout = n;
  p.out_ld = 2 * n;
  p.n_mtiles = n_mtiles;
  p.n_ntiles = n / bn;
  p.kc = kc;
  EHB_CUDA(ehb::launch_conv_gemm(tA, tB, tA, tB, p, ctx->num_sms, nullptr));
  ctx->launches += 1;
  EHB_CUDA(cudaMemcpy(out, dO.p, static_cast<size_t>(m) * n * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int ehb_resnet_forward(ehb_ctx* ctx, const float* img, int n, int h, int w, float* feats, void* stream_) {
  if (!ctx || !img || !feats) return fail("ehb_resnet_forward: null argument");
  if (!ctx->rn_loaded) return fail("ehb_resnet_forward: call ehb_resnet_load first");
  if (n <= 0 || h < 32 || w < 32) return fail("ehb_resnet_forward: need n > 0 and an image of at least 32 x 32");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  auto pad_rows = [](long long r) { return static_cast<size_t>((r + 255) / 256) * 256; };
  auto out_dim = [](int x, int k, int s, int p) { return (x + 2 * p - k) / s + 1; };
  const auto& cv = ctx->rn_convs;
  // ---- buffer sizes for this batch
  const int H1 = out_dim(h, 7, 2, 3), W1 = out_dim(w, 7, 2, 3), H2 = out_dim(H1, 3, 2, 1), W2 = out_dim(W1, 3, 2, 1);
  // space-to-depth image: planes [hi | lo][n][H2][W2 + 4][16] (stem_s2d_kernel writes every element, pad columns included)
  const size_t s2d_bytes = 2 * static_cast<size_t>(n) * ((h + 1) / 2) * ((w + 1) / 2 + 4) * 16 * sizeof(__half) + 256;
  size_t col_bytes = 256;
  size_t x_bytes = pad_rows(static_cast<long long>(n) * H1 * W1) * 2 * cv[0]->cout * sizeof(__half);
  size_t y_bytes = 0;
  {
    int H = H2, W = W2, ci = 1;
    for (int s = 0; s < 4; ++s)
      for (int b = 0; b < ctx->rn_blocks[s]; ++b) {
        const auto& c1 = *cv[ci];
        const auto& c2 = *cv[ci + 1];
        const auto& c3 = *cv[ci + 2];
        const int Ho = out_dim(H, 3, c2.stride, 1), Wo = out_dim(W, 3, c2.stride, 1);
        const size_t rin = pad_rows(static_cast<long long>(n) * H * W), rout = pad_rows(static_cast<long long>(n) * Ho * Wo);
        y_bytes = std::max(y_bytes, std::max(rin * 2 * c1.cout, rout * 2 * c2.cout) * sizeof(__half));
        col_bytes = std::max(col_bytes, rout * 2 * c2.Kp * sizeof(__half));
        x_bytes = std::max(x_bytes, rout * 2 * c3.cout * sizeof(__half));
        if (b == 0) {
          const auto& cd = *cv[ci + 3];
          if (cd.stride != 1) col_bytes = std::max(col_bytes, rout * 2 * cd.Kp * sizeof(__half));
        }
        ci += b == 0 ? 4 : 3;
        H = Ho;
        W = Wo;
      }
  }
  EHB_CUDA(ctx->rn_s2d.ensure(s2d_bytes, true));
  EHB_CUDA(ctx->rn_col.ensure(col_bytes, true));
  EHB_CUDA(ctx->rn_x[0].ensure(x_bytes, true));
  EHB_CUDA(ctx->rn_x[1].ensure(x_bytes, true));
  EHB_CUDA(ctx->rn_y1.ensure(y_bytes, true));
  EHB_CUDA(ctx->rn_y2.ensure(y_bytes, true));
  __half* col = ctx->rn_col.as<__half>();
  // ---- stem: conv 7x7 / 2 + BN + ReLU, max-pool 3x3 / 2  (models/resnet.py:109-113, 140-143)
  {
    const auto& c0 = *cv[0];
    if (W1 > 128) return fail("ehb_resnet_forward: images wider than 256 pixels are not supported by the implicit-GEMM stem");
    // image -> space-to-depth hi/lo operand (51 MB for 64 images), then the stem as an implicit GEMM over it
    EHB_CUDA(ehb::launch_stem_s2d(img, ctx->rn_s2d.as<__half>(), n, h, w, ctx->rn_act_scale, stream));
    if (rn_gemm_implicit(ctx, c0, ctx->rn_s2d.as<__half>(), n, (h + 1) / 2, (w + 1) / 2, H1, W1, ctx->rn_x[1].as<__half>(), 1, stream))
      return 1;
    EHB_CUDA(ehb::launch_maxpool_hl(ctx->rn_x[1].as<__half>(), ctx->rn_x[0].as<__half>(), n, H1, W1, c0.cout, stream));
    ctx->launches += 2;
  }
  int cur = 0, H = H2, W = W2, C = cv[0]->cout, ci = 1;
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < ctx->rn_blocks[s]; ++b) {
      const auto& c1 = *cv[ci];
      const auto& c2 = *cv[ci + 1];
      const auto& c3 = *cv[ci + 2];
      if (c1.kh != 1 || c1.cin != C || c2.kh != 3 || c2.cin != c1.cout || c3.kh != 1 || c3.cin != c2.cout)
        return fail("ehb_resnet_forward: unexpected bottleneck layout");
      const __half* x = ctx->rn_x[cur].as<__half>();
      __half* xo = ctx->rn_x[cur ^ 1].as<__half>();
      const int Ho = out_dim(H, 3, c2.stride, 1), Wo = out_dim(W, 3, c2.stride, 1);
      const long long rin = static_cast<long long>(n) * H * W, rout = static_cast<long long>(n) * Ho * Wo;
      if (rn_gemm(ctx, c1, x, rin, nullptr, ctx->rn_y1.as<__half>(), 1, stream)) return 1;
      if (ctx->rn_implicit && Wo <= 128) {
        if (rn_gemm_implicit(ctx, c2, ctx->rn_y1.as<__half>(), n, H, W, Ho, Wo, ctx->rn_y2.as<__half>(), 1, stream)) return 1;
      } else {
        EHB_CUDA(ehb::launch_im2col_hl(ctx->rn_y1.as<__half>(), col, n, H, W, c1.cout, 3, 3, c2.stride, 1, Ho, Wo, stream));
        if (rn_gemm(ctx, c2, col, rout, nullptr, ctx->rn_y2.as<__half>(), 1, stream)) return 1;
        ctx->launches += 1;
      }
      if (b == 0) {
        // relu(conv3(y2) + downsample(x)): both GEMMs accumulate into the same tile (models/resnet.py:90-96)
        const auto& cd = *cv[ci + 3];
        if (cd.kh != 1 || cd.cin != C || cd.cout != c3.cout || cd.stride != c2.stride)
          return fail("ehb_resnet_forward: unexpected downsample layout");
        const __half* a = x;
        if (cd.stride != 1) {   // strided 1x1: gather the kept pixels (the 3x3 im2col matrix in `col` has been consumed)
          EHB_CUDA(ehb::launch_im2col_hl(x, col, n, H, W, C, 1, 1, cd.stride, 0, Ho, Wo, stream));
          ctx->launches += 1;
          a = col;
        }
        if (rn_gemm(ctx, c3, ctx->rn_y2.as<__half>(), rout, nullptr, xo, 1, stream, &cd, a)) return 1;
      } else {
        if (c3.cout != C || c2.stride != 1) return fail("ehb_resnet_forward: identity shortcut with a shape change");
        if (rn_gemm(ctx, c3, ctx->rn_y2.as<__half>(), rout, x, xo, 1, stream)) return 1;
      }
      ci += b == 0 ? 4 : 3;
      cur ^= 1;
      H = Ho;
      W = Wo;
      C = c3.cout;
    }
  EHB_CUDA(ehb::launch_avgpool_hl(ctx->rn_x[cur].as<__half>(), feats, n, H * W, C, ctx->rn_act_scale, stream));
  ctx->launches += 1;
  return 0;
}

__global__ void relu_inplace_kernel(float* x, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fmaxf(x[i], 0.f);
}

int ehb_linear_f32(ehb_ctx* ctx, const float* x, const float* w_t, const float* bias, int m, int n, int k, int relu,
                   float* y, void* stream_) {
  if (!ctx || !x || !w_t || !y) return fail("ehb_linear_f32: null argument");
  if (m <= 0 || n <= 0 || k <= 0) return fail("ehb_linear_f32: sizes must be positive");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t tot = static_cast<size_t>(m) * n;
  if (bias) {
    fill_rows_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, stream>>>(y, bias, m, n);
    EHB_CUDA(cudaGetLastError());
    ctx->launches += 1;
  }
  if (ctx_sgemm(ctx, x, w_t, y, m, n, k, k, n, n, bias ? 1 : 0, stream)) return 1;
  if (relu) {
    relu_inplace_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, stream>>>(y, tot);
    EHB_CUDA(cudaGetLastError());
    ctx->launches += 1;
  }
  return 0;
}

int ehb_nn_dist_sq(ehb_ctx* ctx, const float* q, const int32_t* q_index, int n_q, const float* r, const int32_t* r_index,
                   int n_r, int n_pairs, float* out, void* stream_) {
  if (!ctx || !q || !r || !out) return fail("ehb_nn_dist_sq: null argument");
  if (n_q <= 0 || n_r <= 0 || n_pairs < 0) return fail("ehb_nn_dist_sq: point counts must be positive");
  if (n_pairs > 65535) return fail("ehb_nn_dist_sq: at most 65535 cloud pairs per call");
  EHB_CUDA(cudaSetDevice(ctx->device));
  EHB_CUDA(ehb::launch_nn_dist_sq(q, q_index, n_q, r, r_index, n_r, n_pairs, out, static_cast<cudaStream_t>(stream_)));
  ctx->launches += n_pairs > 0;
  return 0;
}

int ehb_rotmat_to_angle_axis(ehb_ctx* ctx, const float* R, int n, float* aa, void* stream_) {
  if (!ctx || !R || !aa) return fail("ehb_rotmat_to_angle_axis: null argument");
  if (n < 0) return fail("ehb_rotmat_to_angle_axis: negative n");
  EHB_CUDA(cudaSetDevice(ctx->device));
  EHB_CUDA(ehb::launch_rotmat_to_aa(R, aa, n, static_cast<cudaStream_t>(stream_)));
  ctx->launches += n > 0;
  return 0;
}

int ehb_smpl_backward(ehb_ctx* ctx, const float* x_t, const float* betas, const float* g_verts, const float* g_joints,
                      const float* g_aa, float* grad_x, void* stream_) {
  if (!ctx || !x_t || !betas || !grad_x) return fail("ehb_smpl_backward: null argument");
  if (ctx->n_bodies <= 0) return fail("ehb_smpl_backward: call ehb_set_bodies first");
  if (!ctx->norm_set || !ctx->smpl_loaded) return fail("ehb_smpl_backward: call ehb_set_norm and ehb_smpl_load first");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int n = ctx->n_bodies;
  const size_t V = ctx->smpl.V;
  EHB_CUDA(ctx->sc_R.ensure(static_cast<size_t>(n) * ehb::NJ * 9 * sizeof(float)));
  EHB_CUDA(ctx->sc_A.ensure(static_cast<size_t>(n) * ehb::NJ * 12 * sizeof(float)));
  EHB_CUDA(ctx->sc_j24.ensure(static_cast<size_t>(n) * ehb::NJ * 3 * sizeof(float)));
  EHB_CUDA(ctx->sc_pf.ensure(static_cast<size_t>(n) * 207 * sizeof(float)));
  EHB_CUDA(ctx->sc_dvp.ensure(static_cast<size_t>(n) * V * 3 * sizeof(float)));
  EHB_CUDA(ctx->sc_dA.ensure(static_cast<size_t>(n) * ehb::NJ * 12 * sizeof(float)));
  EHB_CUDA(ctx->sc_dpf.ensure(static_cast<size_t>(n) * 207 * sizeof(float)));
  const int32_t* idx = ctx->img_of_body.as<int32_t>();
  // forward products for this x (rotations, skinning transforms, pose feature)
  EHB_CUDA(ehb::launch_rot6d(x_t, ctx->mean.as<float>(), ctx->std_.as<float>(), ctx->sc_R.as<float>(), n, stream));
  EHB_CUDA(ehb::launch_smpl_pose(ctx->smpl, ctx->sc_R.as<float>(), betas, idx, ctx->sc_A.as<float>(),
                                 ctx->sc_j24.as<float>(), ctx->sc_pf.as<float>(), n, stream));
  EHB_CUDA(ehb::launch_smpl_backward(ctx->smpl, x_t, ctx->mean.as<float>(), ctx->std_.as<float>(), betas, idx,
                                     ctx->sc_A.as<float>(), ctx->sc_pf.as<float>(), g_verts, g_joints, g_aa,
                                     ctx->sc_dvp.as<float>(), ctx->sc_dA.as<float>(), ctx->sc_dpf.as<float>(), grad_x, n,
                                     stream));
  ctx->launches += 5;
  return 0;
}

int ehb_decode(ehb_ctx* ctx, const float* x0, const float* betas, float* pose6d, float* R, float* verts, float* joints,
               void* stream_) {
  if (!ctx || !x0 || !R) return fail("ehb_decode: null argument");
  if (ctx->n_bodies <= 0) return fail("ehb_decode: call ehb_set_bodies first");
  if (!ctx->norm_set) return fail("ehb_decode: call ehb_set_norm first");
  EHB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int n = ctx->n_bodies;
  if (pose6d) {
    const size_t tot = static_cast<size_t>(n) * ehb::XDIM;
    denorm_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, stream>>>(x0, ctx->mean.as<float>(),
                                                                                ctx->std_.as<float>(), pose6d, tot);
    EHB_CUDA(cudaGetLastError());
    ctx->launches += 1;
  }
  EHB_CUDA(ehb::launch_rot6d(x0, ctx->mean.as<float>(), ctx->std_.as<float>(), R, n, stream));
  ctx->launches += 1;
  if (!verts && !joints) return 0;
  if (!ctx->smpl_loaded) return fail("ehb_decode: call ehb_smpl_load first");
  if (!betas) return fail("ehb_decode: betas is NULL");
  return smpl_run(ctx, n, R, betas, ctx->img_of_body.as<int32_t>(), nullptr, verts, joints, stream);
}

}  // extern "C"

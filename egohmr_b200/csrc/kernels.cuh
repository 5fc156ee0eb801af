// Internal (non-ABI) declarations shared by the CUDA translation units.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ehb {

constexpr int NJ = 24;          // SMPL joints == graph nodes
constexpr int XDIM = NJ * 6;    // 144-d rot6d body representation
constexpr int SLOTS_PER_TILE = 5;   // (body,pass) slots per 128-row GEMM tile (5*24 = 120 rows + 8 pad)
constexpr int TILE_ROWS = 128;

// sym(adj + adj2) of one ModulatedGraphConv, split the way its forward uses it
// (reference: models/egohmr/modulated_gcn/modulated_gcn_conv.py:42-46): `adj*E` (diagonal) and `adj*(1-E)`.
struct AdjMix {
  float diag[NJ];
  float off[NJ][NJ];  // off[j][i], zero on the diagonal
};

// Epilogue of one hidden ModulatedGraphConv + BatchNorm1d(eval) + ReLU (+ residual).
struct HiddenLayerParams {
  AdjMix adj;
  const float* mod;       // [24][C]  M[j][c] / (act_scale * w_scale)   (undoes the fp16 operand scaling, exact)
  const float* bn_scale;  // [C]      gamma / sqrt(var + eps)
  const float* bn_shift;  // [C]      beta + (bias - mean) * bn_scale
  float* res;             // [rows_pad][C] fp32 block-boundary activations (read if add_res, written if write_f32)
  __half* out_hl;         // [rows_pad][2C] next layer's A operand: [hi(C) | lo(C)] of act_scale * value
  int* overflow_flag;     // set to 1 if a scaled activation leaves the fp16 range
  float act_scale;
  int C;                  // channels (== K of the GEMM)
  int n_mtiles, n_ntiles; // n_ntiles = C / 128
  int n_slots;            // valid (body,pass) slots; slot s lives in tile s/5 at rows (s%5)*24 ..
  int add_res, write_f32, write_hl;
};

// launchers (gcn_umma.cu / gcn_simt.cu / smpl_lbs.cu)
// ctas = 1: one CTA per 128x256 tile (tmB box = 256 rows); ctas = 2: CTA pairs, tcgen05 cta_group::2 (tmB box = 128
// rows, n_mtiles even).
// pdl: launch with programmatic stream serialization (the kernel overlaps its set-up with the tail of the previous
// kernel in the stream and waits for it with griddepcontrol.wait before touching global memory)
cudaError_t launch_gcn_hidden_umma(const CUtensorMap& tmA, const CUtensorMap& tmB, const HiddenLayerParams& p,
                                   int num_sms, int ctas, bool pdl, cudaStream_t stream);
// transposed product (gcn_umma_t.cu): weights on the M side, N = 240 activation rows = 10 slots without pad rows.
// tmX: activation map with 120-row boxes; tmW: weight map with 64-row boxes.  n_mtiles even.
cudaError_t launch_gcn_hidden_umma_t(const CUtensorMap& tmX, const CUtensorMap& tmW, const HiddenLayerParams& p,
                                     int num_sms, bool pdl, cudaStream_t stream);
// all hidden layers of a reverse step as ONE persistent launch of the transposed kernel (gcn_umma_fused.cu): per-(layer, row
// group) completion counters order the layers unit by unit.  done: [n_layers][n_mtiles / 2] ints (zeroed by the launcher).
constexpr int MAX_FUSED_LAYERS = 8;
struct FusedHiddenMaps {
  CUtensorMap x[2];                    // the two activation operand buffers, 120-row boxes (layer l reads x[l & 1])
  CUtensorMap w[MAX_FUSED_LAYERS];     // per-layer weights, 64-row boxes
};
struct FusedHiddenParams {
  HiddenLayerParams layer[MAX_FUSED_LAYERS];
  int n_layers;
  int* done;
};
cudaError_t launch_gcn_hidden_fused(const FusedHiddenMaps& maps, const FusedHiddenParams& fp, int num_sms, bool pdl,
                                    cudaStream_t stream);
size_t gcn_hidden_umma_smem_bytes();
int gcn_hidden_umma_bk();   // fp16 elements per TMA box row (32: SWIZZLE_64B, 64: SWIZZLE_128B)
#ifndef EHB_UMMA_BK
#define EHB_UMMA_BK 64
#endif

// fp32 SIMT check path for the same layer: H = X * Wcat via sgemm, then the identical epilogue.
cudaError_t launch_gcn_hidden_simt(const float* x_f32, const float* wcat /*[K][2C] fp32*/, float* h_tmp /*[rows][2C]*/,
                                   const HiddenLayerParams& p, const float* mod_unscaled, float* out_f32,
                                   cudaStream_t stream);

// C[M][N] = A[M][K] * B[K][N]  (+ C if accumulate), all row-major fp32, plain FFMA.
cudaError_t launch_sgemm_nn(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                            int accumulate, cudaStream_t stream);

// Deterministic split-K variant for skinny problems (M = images / clouds): `splits` K chunks write partial planes into
// `scratch` (>= splits*M*N floats), a second kernel sums them in a fixed order.  splits <= 1 falls back to the above.
int sgemm_splitk_plan(int M, int N, int K, int num_sms);
cudaError_t launch_sgemm_nn_splitk(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb,
                                   int ldc, int accumulate, float* scratch, int splits, int* n_launches,
                                   cudaStream_t stream);

struct InputLayerParams {
  AdjMix adj;
  const float* a01;      // [n_img][2][C]   img_feat . W_k[0:img_dim]
  const float* be01;     // [n_img][2][C]   rest_feat . W_k[img_dim:cond_dim] + inproc_b . W_k[x rows]
  const float* cx01;     // [2][C]          inproc_b . W_k[x rows]  (the condition-free part of be01)
  const float* ct01;     // [n_steps][2][C] temb(step) . W_k[temb rows]
  const float* wx01;     // [2][6][C]       inproc_w^T . W_k[x rows]
  const float* mod;      // [24][C] (unscaled M)
  const float* bn_scale; // [C]
  const float* bn_shift; // [C]
  const uint8_t* vis;    // [n_img][24]
  const int32_t* slot_body;  // [n_slots]
  const uint8_t* slot_cond;  // [n_slots] 1 = image-conditioned pass, 0 = image-masked pass
  const int32_t* img_of_body;
  const float* x_t;      // [B][144]
  float* res;            // [rows_pad][C]
  __half* out_hl;        // [rows_pad][2C]
  int* overflow_flag;
  float act_scale;
  int C, n_slots, step;
  int mask_all;          // image-masked pass drops every condition (only_mask_img_cond=False, egohmr.py:157-158)
};
cudaError_t launch_gcn_input(const InputLayerParams& p, cudaStream_t stream);            // fp32 FFMA (check path)
cudaError_t launch_gcn_input_umma(const InputLayerParams& p, int num_sms, cudaStream_t stream);   // joint mix on tcgen05

// Non-local block between the residual blocks and the output layer (ModulatedGCN(nonlocal_layer=True)).
struct NonLocalParams {
  const float* tpg;       // theta|phi|g projections as 256-column fp32 planes [n_planes][rows_pad][256]
  const float* wy;        // W projection planes [C/256 (padded)][rows_pad][256]
  size_t plane_stride;    // rows_pad * 256
  __half* y_hl;           // [rows_pad][hi(inter) | lo(inter)]: A operand of the W projection
  float* res;             // [rows_pad][C] activations, updated in place
  const float* bn_scale;  // [C]  gamma / sqrt(var + eps)
  const float* bn_shift;  // [C]  beta + (conv bias - mean) * bn_scale
  int* overflow_flag;
  float act_scale;
  int C, inter, n_slots;
};
cudaError_t launch_nonlocal_attention(const NonLocalParams& p, cudaStream_t stream);
cudaError_t launch_nonlocal_residual(const NonLocalParams& p, cudaStream_t stream);

// Per-step sampler coefficients (host-computed in fp32 exactly as the reference's torch ops would):
//  DDIM:          c[0]=sqrt_recip_alphas_cumprod, c[1]=sqrt_recipm1_alphas_cumprod, c[2]=sqrt(alpha_bar_prev),
//                 c[3]=sqrt(1-alpha_bar_prev-sigma^2), c[4]=(t!=0)*sigma (eta), c[5]=sqrt(1-alpha_bar) on the steps where
//                 ddim_sample_with_grad applies the collision gradient (respaced t <= 3), else 0
//  DDPM:          c[0]=posterior_mean_coef1, c[1]=posterior_mean_coef2, c[2]=(t!=0)*exp(0.5*log_var),
//                 c[3]=gradient scale (cond_grad_weight*variance, or cond_grad_weight*0.01), 0 when unguided
struct StepCoef {
  float c[8];
};
enum SamplerKind { SAMPLER_DDIM = 0, SAMPLER_DDPM = 1 };

struct OutputLayerParams {
  AdjMix adj;
  const float* act;       // [rows_pad][C] fp32 activations of the last block
  const float* wout;      // [C][12]  (k*6+d)
  const float* mod;       // [24][6]
  const float* bias;      // [6]
  const uint8_t* vis;     // [n_img][24]
  const int32_t* img_of_body;
  const int32_t* body_slot;  // [B][2]: slot of the cond pass, slot of the uncond pass (-1 = not evaluated)
  const float* x_t;       // [B][144]
  const float* noise;     // [B][144] or null
  const float* grad;      // [B][144] or null
  float* x_prev;          // [B][144]
  float* x0;              // [B][144]
  float* out_cond;        // optional [B][144] raw image-conditioned denoiser output (tests)
  float* out_uncond;      // optional [B][144]
  float* x0_model;        // optional [B][144]: the fused model prediction BEFORE the sampler update touches it (the guided
                          // DDIM step re-derives pred_xstart; EgoHMR.forward's own output, 'other_outputs', keeps this one)
  StepCoef coef;
  int kind;
  int C, n_bodies, diffuse_fuse;
};
cudaError_t launch_gcn_output(const OutputLayerParams& p, int num_sms, cudaStream_t stream);
// x_prev = update(x_t, x0, noise, grad) elementwise over n floats (same arithmetic as the tail of launch_gcn_output);
// x0_out (optional) receives the pred_xstart the step ends with (the guided DDIM step re-derives it)
cudaError_t launch_sampler_update(const StepCoef& coef, int kind, const float* x_t, const float* x0,
                                  const float* noise, const float* grad, float* x_prev, float* x0_out, size_t n,
                                  cudaStream_t stream);

// ---- SMPL ----
struct SmplDevice {
  const float* v_template;   // [V][3]
  const float* shapedirs;    // [V][3][NB]
  const float* posedirs;     // [207][V*3]
  const float* lbs_weights;  // [V][24]
  const float* j_template;   // [24][3]    J_regressor . v_template
  const float* j_shapedirs;  // [24][3][NB] J_regressor . shapedirs
  const int32_t* extra_vids; // [n_extra]
  const float* shapedirs_t;  // [3 * NB][V]  shapedirs transposed (coalesced along the vertex axis)
  const float* skin_w4;      // [V][4] non-zero skinning weights in ascending joint order (0-padded), or null if a
  const uint8_t* skin_j4;    // [V][4] vertex has more than four non-zero weights (then only the dense route is used)
  int parents[NJ];
  int V, NB, n_extra;
};
// x (normalised rot6d, [B][144]) -> R [B][24][9]; mean/std may be null (identity).
cudaError_t launch_rot6d(const float* x, const float* mean, const float* std_, float* R, int n_bodies,
                         cudaStream_t stream);
// generic [n][6] -> [n][9] ('diffusion' column order), no normalisation stats
cudaError_t launch_rot6d_flat(const float* x6, float* R, int n, cudaStream_t stream);
// R [B][24][9], betas [nb][NB] indexed through beta_index (null = identity) -> A [B][24][12], posed joints [B][24][3],
// pose feature [B][207]
cudaError_t launch_smpl_pose(const SmplDevice& m, const float* R, const float* betas, const int32_t* beta_index,
                             float* A, float* joints24, float* posefeat, int n_bodies, cudaStream_t stream);
cudaError_t launch_smpl_skin(const SmplDevice& m, const float* betas, const int32_t* beta_index, const float* A,
                             const float* posefeat, const float* transl /*[B][3] or null*/, float* verts,
                             int n_bodies, cudaStream_t stream);
// tensor-core route of the pose blend: pose features -> GEMM A operand [n_pad][hi(256) | lo(256)] (K = 207 zero-padded),
// and skinning from the GEMM's body-major Y[n_pad][ldy] = pose_feature . posedirs (ldy = 3V padded to a tile multiple)
cudaError_t launch_smpl_pf_operand(const float* posefeat, __half* pf_hl, int n_bodies, int n_pad, float scale,
                                   cudaStream_t stream);
cudaError_t launch_smpl_skin_tiled(const SmplDevice& m, const float* betas, const int32_t* beta_index, const float* A,
                                   const float* Y, size_t ldy, const float* transl, float* verts, int n_bodies,
                                   cudaStream_t stream);
cudaError_t launch_smpl_joints(const SmplDevice& m, const float* joints24, const float* verts, const float* transl,
                               float* joints /*[B][24+n_extra][3]*/, int n_bodies, cudaStream_t stream);

// ---- ResPointNet linear layers (linear_umma.cu)
struct LinearParams {
  const float* bias;      // [256] or null
  const float* rowvec;    // [n_clouds][256] or null: per-cloud row added to every point of the cloud
  float* out_f32;         // [M][256] or null
  __half* out_hl;         // [M_pad][hi(256) | lo(256)] of act_scale * Y, or null
  __half* out_hl_relu;    // same for relu(Y), or null
  int* pool;              // [n_clouds][256] order-preserving-int column max over the cloud's points, or null
  const float* pts;       // [M][3] or null: adds pts[row] . ptw to every row (a K = 3 linear term folded into the epilogue)
  const float* ptw;       // [3][256]
  int* overflow_flag;
  long long M;            // valid rows
  float acc_scale_inv;    // 1 / (act_scale * w_scale)
  float act_scale;
  int K1, K2;             // K of the two accumulated operand pairs (K2 = 0: single)
  int n_mtiles;           // 128-row tiles, even
  int pts_per_cloud;
};
cudaError_t launch_linear_umma(const CUtensorMap& tmA1, const CUtensorMap& tmB1, const CUtensorMap& tmA2,
                               const CUtensorMap& tmB2, const LinearParams& p, int num_sms, cudaStream_t stream);
cudaError_t launch_pointnet_pos(const float* pts, const float* w, const float* b, __half* out_hl, __half* out_hl_relu,
                                long long M, int C, float act_scale, int* overflow_flag, cudaStream_t stream);
cudaError_t launch_pool_init(int* pool, int n, cudaStream_t stream);
cudaError_t launch_pool_decode(const int* pool, float* out, int n, cudaStream_t stream);

// ---- ResNet-50 convolutions as GEMMs (conv_umma.cu) and their data movers (resnet_ops.cu)
struct ConvGemmParams {
  const float* bias;      // [Cout] folded BatchNorm shift
  const __half* res_hl;   // identity branch, same layout as out_hl, or null
  __half* out_hl;         // [M_pad][hi(Cout) | lo(Cout)] of act_scale * Y, or null
  float* out_f32;         // [M][Cout] or null
  int* overflow_flag;
  long long M;            // valid rows (image pixels)
  float acc_scale_inv;    // 1 / (act_scale * w_scale)
  float act_scale;
  int K;                  // kh*kw*Cin padded to a multiple of 64
  int K2;                 // K of an optional second operand pair accumulated into the same tile (0 = none)
  int Cout, out_ld;       // out_ld = 2 * Cout (halves per row)
  int n_mtiles, n_ntiles; // 128-row tiles (even) and Cout / tile_n
  int relu;
  int kc;                 // k-blocks (of 64) chained into one TMEM accumulation before the epilogue warps take the partial
                          // sum over in fp32 registers; <= 0 or >= K/64: the whole K in one accumulator (tile_n <= 128 only)
  // Implicit-GEMM mode (implicit != 0): tmA is a 4-D map over the NHWC input [N][H][W][hi(Cin) | lo(Cin)] and the A tile
  // of k-block (tap, channel chunk) is the TMA box of the tile's output pixels shifted by the tap — the convolution's
  // zero padding is TMA's out-of-bounds fill, no im2col matrix exists.  An M tile is `nb` images x `th` full output
  // rows (th * Wo * nb <= 128); its output rows are contiguous, first row = (n0 * Ho + ho0) * Wo.
  int implicit;
  int Cin, kw, pad, stride;   // K = kh * kw * Cin, k = (ky * kw + kx) * Cin + c; `pad` rows above, `pad_w` columns left
  int pad_w;
  int lo_plane;               // 0: a pixel holds [hi(Cin) | lo(Cin)]; > 0: the lo halves are images n + lo_plane of the tensor
  int Ho, Wo, th, nb, tiles_per_img, n_img;
};
int conv_gemm_tile_n(int cout, long long rows, int num_sms, int max_bn = 256);   // 256, 128 or 64 output channels per tile
cudaError_t launch_conv_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2, const CUtensorMap& tmB2,
                             const ConvGemmParams& p, int num_sms, cudaStream_t stream);
// im2col of an NHWC fp16 hi/lo activation [N][H][W][hi(C) | lo(C)] for a KHxKW / stride / pad convolution:
// dst [N*Ho*Wo][hi(KH*KW*C) | lo(KH*KW*C)], k = (ky*KW + kx)*C + c, zero outside the image.  C % 8 == 0.
cudaError_t launch_im2col_hl(const __half* src, __half* dst, int N, int H, int W, int C, int KH, int KW, int stride,
                             int pad, int Ho, int Wo, cudaStream_t stream);
// space-to-depth of the fp32 NCHW image for the stem as a stride-1 4x4 convolution on 16 channels: two planes (hi, lo) of
// dst [2][N][H2][W2 + 4][16], channel (dy*2 + dx)*3 + c = act_scale * img[n][c][2*y2 + dy][2*(x - 2) + dx] (12 used, 4 zero);
// two zero columns on each side; H2 = (H+1)/2, W2 = (W+1)/2
cudaError_t launch_stem_s2d(const float* img, __half* dst, int N, int H, int W, float act_scale, cudaStream_t stream);
// MaxPool2d(3, 2, 1) on an NHWC hi/lo activation (the pair with the largest hi + lo wins; exact in fp32)
cudaError_t launch_maxpool_hl(const __half* src, __half* dst, int N, int H, int W, int C, cudaStream_t stream);
// global average pool: [N][HW][hi(C) | lo(C)] -> fp32 [N][C]
cudaError_t launch_avgpool_hl(const __half* src, float* dst, int N, int HW, int C, float act_scale, cudaStream_t stream);

// ---- image ops (image_ops.cu): MaxPool2d(3, 2, 1) on NHWC fp32, C % 4 == 0; out is [N][(H+1)/2][(W+1)/2][C]
cudaError_t launch_maxpool3x3s2_nhwc(const float* in, float* out, int N, int H, int W, int C, cudaStream_t stream);

// step-invariant glue of EgoHMR.forward (visibility mask, [scene | transl | camera] condition vector, beta-head input) and
// the final joints' translation + perspective projection; see image_ops.cu
cudaError_t launch_cond_inputs(const float* kp2d, const float* scene_feat, const float* transl_feat, const float* img_feat,
                               const float* fx, const float* box_center, const float* box_size, const float* cam_cx,
                               const float* cam_cy, int n, int sf, int tf, int img_dim, int with_focal, int with_bbox,
                               int with_center, const int32_t* o2s, float coeff, uint8_t* vis, float* rest, float* ctxfull,
                               cudaStream_t stream);
cudaError_t launch_project_joints(const float* joints, const float* transl, const float* fx, const float* cam_cx,
                                  const float* cam_cy, const int32_t* img_of_body, int n_bodies, int J, float coeff,
                                  float default_focal, float* kp3d_full, float* kp2d, float* focal_out, float* center_out,
                                  cudaStream_t stream);

// per-body bounding-box crop of the scene cloud (egohmr.py:550-554): mask [B][n_pts], count [B], bbox [B][6] (optional)
cudaError_t launch_scene_crop(const float* verts, int n_bodies, int V, const float* scene, int n_pts,
                              const int32_t* img_of_body, uint8_t* mask, int32_t* count, float* bbox,
                              cudaStream_t stream);

// ---- evaluation metrics (metrics.cu): batched Procrustes alignment (utils/pose_utils.py:11-105).  S1, S2 [P][N][3];
// mask [P][N][3] or null (the *_with_vis_mask variant); S1_hat [P][N][3] and err [P][N] = |S1_hat - S2| may be null.
cudaError_t launch_procrustes(const float* S1, const float* S2, const float* mask, int n_problems, int n_pts,
                              float* S1_hat, float* err, cudaStream_t stream);

// squared distance of every query point to its nearest reference point, per (query cloud, reference cloud) pair:
// pair i uses q[q_index ? q_index[i] : i] ([n_q][3]) and r[r_index ? r_index[i] : i] ([n_r][3]); out [n_pairs][n_q]
cudaError_t launch_nn_dist_sq(const float* q, const int32_t* q_index, int n_q, const float* r, const int32_t* r_index,
                              int n_r, int n_pairs, float* out, cudaStream_t stream);

// evaluation reductions of the reference driver (test_egohmr.py:373-494), see metrics.cu
cudaError_t launch_vis_mask(const float* pts, const float* focal, const float* cx, const float* cy, uint8_t* mask, int n_img,
                            int n_pts, float W, float H, cudaStream_t stream);
cudaError_t launch_pose_errors(const float* pj, const float* pv, const float* transl, const float* gj, const float* gv,
                               const uint8_t* jmask, const uint8_t* vmask, float* out, int n_img, int S, int J, int V,
                               cudaStream_t stream);
cudaError_t launch_diversity(const float* pj, const uint8_t* jmask, float* out, int n_img, int S, int J, cudaStream_t stream);

// ---- guidance backward (smpl_bwd.cu)
cudaError_t launch_rotmat_to_aa(const float* R, float* aa, int n, cudaStream_t stream);
// dL/dx [B][144] from dL/dverts [B][V][3], dL/djoints [B][24+E][3], dL/dfull_pose_aa [B][24][3] (each may be null).
// A / posefeat: forward products of launch_smpl_pose for the same x; d_vposed [B][V][3], dA [B][24][12], d_pf [B][207]
// are scratch.
cudaError_t launch_smpl_backward(const SmplDevice& m, const float* x, const float* mean, const float* std_,
                                 const float* betas, const int32_t* beta_index, const float* A, const float* posefeat,
                                 const float* g_verts, const float* g_joints, const float* g_aa, float* d_vposed,
                                 float* dA, float* d_pf, float* grad_x, int n_bodies, cudaStream_t stream);

}  // namespace ehb

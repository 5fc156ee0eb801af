/*
 * egohmr_b200 — C ABI of the B200-native EgoHMR diffusion-sampling hot path.
 *
 * Every entry point replaces one piece of the reference's Python hot path (citations are into the reference tree,
 * sanweiliti/EgoHMR @ dc4e0a0).  Plain pointers and sizes only: no torch / C++ types cross this boundary.
 * Unless marked HOST, pointers are DEVICE pointers on the context's device; `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).  All calls return 0 on success; on failure they return non-zero and
 * ehb_last_error() describes why.  No hot call allocates, synchronises the host, or touches the CPU for math.
 * There is no CPU fallback: without a CUDA device ehb_ctx_create fails.
 */
#ifndef EGOHMR_B200_H_
#define EGOHMR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EHB_NUM_JOINTS 24
#define EHB_XDIM 144 /* 24 joints x 6-D rotation */

typedef struct ehb_ctx ehb_ctx;

/* One ModulatedGraphConv (+ its BatchNorm1d when it is wrapped in a _GraphConv), parameters in the reference's
 * state_dict layout (models/egohmr/modulated_gcn/modulated_gcn_conv.py:16-36, modulated_gcn.py:9-19).  HOST. */
typedef struct {
  int32_t in_dim, out_dim;
  const float* W;    /* [2][in_dim][out_dim]   "...gconv.W"    */
  const float* M;    /* [24][out_dim]          "...gconv.M"    */
  const float* adj2; /* [24][24]               "...gconv.adj2" */
  const float* bias; /* [out_dim]              "...gconv.bias" */
  const float* bn_weight; /* [out_dim] or NULL (gconv_output has no BN) "...bn.weight"       */
  const float* bn_bias;   /* "...bn.bias"         */
  const float* bn_mean;   /* "...bn.running_mean" */
  const float* bn_var;    /* "...bn.running_var"  */
  float bn_eps;
} ehb_gconv;

/* The denoiser: ModulatedGCN (modulated_gcn.py:61-116) + InputProcess (egohmr.py:646-655) + the feature layout of
 * EgoHMR.forward's 3718-wide input (egohmr.py:220-236): [img(img_dim) | rest(cond_dim-img_dim) | x_feat | temb]. HOST. */
typedef struct {
  int32_t hid;       /* gcn_hid_dim, multiple of 128 */
  int32_t n_blocks;  /* diffusion_blk */
  int32_t img_dim;   /* 2048 */
  int32_t cond_dim;  /* 2694 = img + scene 512 + transl 128 + cam 6 */
  int32_t xfeat_dim; /* 512 */
  int32_t temb_dim;  /* 512 */
  int32_t diffuse_fuse; /* egohmr.py:239 */
  const float* adj;      /* [24][24] normalised skeleton adjacency built in egohmr.py:86-94 */
  const float* inproc_w; /* [xfeat_dim][6]  "input_process.poseEmbedding.weight" */
  const float* inproc_b; /* [xfeat_dim]     "input_process.poseEmbedding.bias"   */
  const ehb_gconv* layers; /* [0] gconv_input, [1 .. 2*n_blocks] gconv_layers.{b}.gconv{1,2}, [last] gconv_output */
  int32_t n_layers;        /* 2*n_blocks + 2 */
  int32_t mask_all_cond;   /* !only_mask_img_cond: the image-masked pass of diffuse_fuse zeroes EVERY condition
                              (mask_cond(force_mask=True), egohmr.py:150-158) instead of the image features only */
} ehb_gcn_weights;

/* NONLocalBlock2D(in_channels=hid, sub_sample=False) of ModulatedGCN(nonlocal_layer=True) (modulated_gcn.py:93-95,
 * nets/non_local_embedded_gaussian.py:6-59); 1x1 convolution weights [out][in] (the trailing 1x1 dims dropped).  HOST. */
typedef struct {
  int32_t inter;            /* hid / 2 */
  const float* theta_w;     /* [inter][hid]  "non_local.theta.weight" */
  const float* theta_b;     /* [inter] */
  const float* phi_w;       /* [inter][hid] */
  const float* phi_b;
  const float* g_w;         /* [inter][hid] */
  const float* g_b;
  const float* W_w;         /* [hid][inter]  "non_local.W.0.weight" */
  const float* W_b;         /* [hid] */
  const float* bn_weight;   /* [hid]  "non_local.W.1.*" (BatchNorm2d, eval) */
  const float* bn_bias;
  const float* bn_mean;
  const float* bn_var;
  float bn_eps;
} ehb_nonlocal_weights;

/* SMPL model tensors as smplx stores them (smplx/body_models.py::SMPL.__init__).  HOST. */
typedef struct {
  int32_t n_verts;       /* 6890 */
  int32_t n_betas;       /* 10   */
  int32_t n_extra;       /* 21 vertex-picked joints appended by VertexJointSelector */
  const float* v_template;  /* [V][3] */
  const float* shapedirs;   /* [V][3][n_betas] */
  const float* posedirs;    /* [207][V*3] */
  const float* J_regressor; /* [24][V] */
  const float* lbs_weights; /* [V][24] */
  const int32_t* parents;   /* [24], parents[0] = -1 */
  const int32_t* extra_vertex_ids; /* [n_extra] */
} ehb_smpl_model;

const char* ehb_last_error(void);
/* number of kernels this library has launched on `ctx` since creation (bench.py's gpu_launches) */
int64_t ehb_launch_count(const ehb_ctx* ctx);
/* Counter bumped whenever the library (re)allocates a device workspace.  Hot calls only allocate when a problem size
 * grows; a caller that captured hot calls into a CUDA graph must re-capture if this value changed since the capture. */
uint64_t ehb_alloc_epoch(void);

int ehb_ctx_create(int device, ehb_ctx** out);
void ehb_ctx_destroy(ehb_ctx* ctx);

/* models/egohmr/egohmr.py:95-99 (ModulatedGCN construction) + load_state_dict (test_egohmr.py:125-126):
 * ingest the denoiser weights and repack them into kernel layouts (fp16 hi/lo split, folded BN, folded input layer). */
int ehb_gcn_load(ehb_ctx* ctx, const ehb_gcn_weights* w);
/* Optional (gcn_nonlocal_layer=True, egohmr.py:37,99): the non-local block evaluated after the residual blocks
 * (modulated_gcn.py:103-109).  Call after ehb_gcn_load; w == NULL switches it off again. */
int ehb_gcn_load_nonlocal(ehb_ctx* ctx, const ehb_nonlocal_weights* w);
/* smplx.create('data/smpl', model_type='smpl', ...) (egohmr.py:105-107, test_egohmr.py:143-145) */
int ehb_smpl_load(ehb_ctx* ctx, const ehb_smpl_model* m);
/* body_rep_mean / body_rep_std (test_egohmr.py:109-111; used at egohmr.py:258,528).  HOST [144] each. */
int ehb_set_norm(ehb_ctx* ctx, const float* mean, const float* std);

/* GaussianDiffusion.__init__ tables (diffusion/gaussian_diffusion.py:122-169) reduced to what one sampler update
 * needs.  kind 0 = ddim_sample / ddim_sample_with_grad (:511-614, any eta), 1 = p_sample / p_sample_with_grad (:298-388).
 * coef is HOST [n_steps][8] floats, row i = coefficients of respaced timestep i (see DESIGN.md "sampler update"). */
int ehb_set_schedule(ehb_ctx* ctx, int kind, int n_steps, const float* coef);

/* Step-invariant conditioning of EgoHMR.forward (egohmr.py:181-223) for n_img images:
 *   img_feat  [n_img][img_dim]            backbone output (:183)
 *   rest_feat [n_img][cond_dim-img_dim]   [scene_feats | transl_feat | cam_feats] (:214-221)
 *   vis       [n_img][24] uint8           vis_mask_smpl (:186-189)
 * Folds them through the input ModulatedGraphConv's weight rows (one fp32 GEMM each). */
int ehb_set_cond(ehb_ctx* ctx, int n_img, const float* img_feat, const float* rest_feat, const uint8_t* vis,
                 void* stream);
/* Timestep embeddings (egohmr.py:178, TimestepEmbedder :642-643): temb [n_steps][temb_dim], row i =
 * embed_timestep(timestep_map[i]) for respaced step i; folded through the input layer's temb rows. */
int ehb_set_temb(ehb_ctx* ctx, int n_steps, const float* temb, void* stream);

/* The chains to sample: body b is conditioned on image img_of_body[b] (HOST int32 [n_bodies]).
 * The reference runs `num_samples` sequential chains per image (test_egohmr.py:251-255); here they are one batch. */
int ehb_set_bodies(ehb_ctx* ctx, int n_bodies, const int32_t* img_of_body);

/* One reverse-diffusion step = p_mean_variance (gaussian_diffusion.py:233-276: model call, i.e. the denoiser of
 * egohmr.py:232-257) + p_sample / p_sample_with_grad / ddim_sample.
 *   step   respaced timestep index i (the reference's `t`, uniform over the batch, :495)
 *   x_t    [n_bodies][144] in;  noise [n_bodies][144] or NULL (ignored by DDIM, eta=0);
 *   grad   [n_bodies][144] or NULL (model.guide_coll output, :379);
 *   x_prev [n_bodies][144] out ('sample');  x0 [n_bodies][144] out ('pred_xstart' == 'pred_x_start'). */
int ehb_denoise_step(ehb_ctx* ctx, int step, const float* x_t, const float* noise, const float* grad, float* x_prev,
                     float* x0, void* stream);
/* Same with one more output: x0_model [n_bodies][144] (may be NULL) = the model's own fused prediction, i.e.
 * 'other_outputs'['pred_x_start'].  It differs from x0 only on the guided DDIM steps, where ddim_sample_with_grad
 * (:579-592) re-derives 'pred_xstart' from the gradient-shifted eps. */
int ehb_denoise_step_ex(ehb_ctx* ctx, int step, const float* x_t, const float* noise, const float* grad, float* x_prev,
                        float* x0, float* x0_model, void* stream);
/* Same, but also returns the raw image-conditioned / image-masked denoiser outputs (egohmr.py:237,246). Tests only. */
int ehb_denoise_step_debug(ehb_ctx* ctx, int step, const float* x_t, const float* noise, const float* grad,
                           float* x_prev, float* x0, float* out_cond, float* out_uncond, void* stream);

/* The sampler update alone (gaussian_diffusion.py:331-336 / :373-387 / :537-555) for a caller-supplied pred_xstart:
 * the generic `p_sample(model, ...)` / `ddim_sample(model, ...)` entry points use it when `model` is not the fused
 * denoiser.  All pointers [n][144]; noise / grad may be NULL. */
int ehb_sampler_update(ehb_ctx* ctx, int step, int n, const float* x_t, const float* x0, const float* noise,
                       const float* grad, float* x_prev, void* stream);
/* Same; x0_out [n][144] (may be NULL) receives the 'pred_xstart' the step returns: ddim_sample_with_grad (:559-614)
 * re-derives it from the gradient-shifted eps on the guided steps, every other step returns x0 unchanged. */
int ehb_sampler_update_ex(ehb_ctx* ctx, int step, int n, const float* x_t, const float* x0, const float* noise,
                          const float* grad, float* x_prev, float* x0_out, void* stream);

/* utils/geometry.py:47-66 rot6d_to_rotmat(x, 'diffusion'): x6 [n][6] -> R [n][3][3]. */
int ehb_rot6d_to_rotmat(ehb_ctx* ctx, const float* x6, int n, float* R, void* stream);

/* egohmr.py:258-260 + :276: x0 (normalised) -> pose_6d = x0*std+mean -> R [n_bodies][24][3][3] -> SMPL.
 * betas [n_img][n_betas] indexed through img_of_body.  verts [n_bodies][V][3] (may be NULL to skip skinning, then
 * the 21 vertex-picked joints are skipped too and joints must be NULL), joints [n_bodies][45][3]. */
int ehb_decode(ehb_ctx* ctx, const float* x0, const float* betas, float* pose6d, float* R, float* verts, float* joints,
               void* stream);

/* smplx SMPL.forward(betas, body_pose, global_orient, transl, pose2rot=False) (call sites egohmr.py:276,492,537,
 * test_egohmr.py:291-292): R [n][24][3][3] (global_orient first), betas [n][n_betas], transl [n][3] or NULL. */
int ehb_smpl_forward(ehb_ctx* ctx, int n, const float* R, const float* betas, const float* transl, float* verts,
                     float* joints, void* stream);

/* ResnetPointnet scene encoder (models/respointnet.py:13-59), nn.Linear parameters in the reference layout ([out][in]).
 * hidden_dim must be 256 (EgoHMR constructs ResnetPointnet(out_dim=scene_feat_dim, hidden_dim=256), egohmr.py:69). HOST. */
typedef struct {
  int32_t hidden;   /* 256 */
  int32_t out_dim;  /* scene_feat_dim, 512 */
  const float* fc_pos_w; /* [2*hidden][3] */
  const float* fc_pos_b; /* [2*hidden]    */
  const float* fc0_w[4];      /* block_i.fc_0.weight     [hidden][2*hidden] */
  const float* fc0_b[4];      /* block_i.fc_0.bias       [hidden]           */
  const float* fc1_w[4];      /* block_i.fc_1.weight     [hidden][hidden]   */
  const float* fc1_b[4];      /* block_i.fc_1.bias       [hidden]           */
  const float* shortcut_w[4]; /* block_i.shortcut.weight [hidden][2*hidden] */
  const float* fc_c_w;   /* [out_dim][hidden] */
  const float* fc_c_b;   /* [out_dim]         */
} ehb_pointnet_weights;
int ehb_pointnet_load(ehb_ctx* ctx, const ehb_pointnet_weights* w);
/* scene_enc(scene_pcd_verts) (egohmr.py:214): pts [n_clouds][n_pts][3] -> feats [n_clouds][out_dim]. */
int ehb_pointnet_forward(ehb_ctx* ctx, const float* pts, int n_clouds, int n_pts, float* feats, void* stream);

/* nn.MaxPool2d(kernel_size=3, stride=2, padding=1) of the ResNet stem (models/resnet.py:113,142) on an NHWC fp32
 * tensor: in [n][h][w][c] -> out [n][(h+1)/2][(w+1)/2][c]; c % 4 == 0, 16-byte aligned pointers. */
int ehb_maxpool3x3s2_nhwc(ehb_ctx* ctx, const float* in, int n, int h, int w, int c, float* out, void* stream);

/* The step-invariant glue of EgoHMR.forward for n_img images in one launch (the reference: ~18 small torch kernels):
 *   vis  [n_img][24] uint8: vis_mask_smpl (egohmr.py:186-189) from orig_keypoints_2d kp2d [n_img][25][3] through the
 *        OpenPose -> SMPL table (HOST int32 [24], :110-114), OpenPose joint 8 forced visible;
 *   rest [n_img][scene_dim + transl_dim + cams] = [scene_feats | transl_feat | cam_cx/f, cam_cy/f | box_cx/f, box_cy/f,
 *        box_size/f | fx] with f = fx * fx_norm_coeff (:195-205, 220-223; camera parts per flag, in this order);
 *   ctx_full [n_img][img_dim + rest] = [img_feats | rest], the beta head's input (:262-263).
 * fx is the NORMALISED focal length of the batch dict; pointers a flag does not need may be NULL. */
int ehb_cond_inputs(ehb_ctx* ctx, const float* kp2d, const float* scene_feat, const float* transl_feat, const float* img_feat,
                    const float* fx, const float* box_center, const float* box_size, const float* cam_cx, const float* cam_cy,
                    int n_img, int scene_dim, int transl_dim, int img_dim, int with_focal_length, int with_bbox_info,
                    int with_cam_center, const int32_t* openpose_to_smpl, float fx_norm_coeff, uint8_t* vis, float* rest,
                    float* ctx_full, void* stream);

/* egohmr.py:277-301 for the final joints [n_bodies][n_joints][3]: kp3d_full = joints + transl[img], kp2d [n_bodies][n_joints][2]
 * = perspective_projection(joints, transl, focal, centre) (utils/geometry.py:78-116, identity rotation) mapped to
 * (u / 1920 - 0.5, v / 1080 - 0.5); focal_out / center_out [n_bodies][2] are the rows compute_loss reads (:361-366).
 * fx != NULL: focal = fx[img] * fx_norm_coeff, centre = (cam_cx, cam_cy)[img]; fx == NULL: default_focal and (960, 540).
 * img_of_body device int32 [n_bodies] or NULL (identity). */
int ehb_project_joints(ehb_ctx* ctx, const float* joints, const float* transl, const float* fx, const float* cam_cx,
                       const float* cam_cy, const int32_t* img_of_body, int n_bodies, int n_joints, float fx_norm_coeff,
                       float default_focal, float* kp3d_full, float* kp2d, float* focal_out, float* center_out, void* stream);

/* The scene crop of guide_coll / eval_coll (egohmr.py:550-554, 504-508) for every body in one launch: body b's
 * axis-aligned bounding box over verts [n_bodies][n_verts][3], then mask[b][i] = 1 iff point i of cloud
 * img_of_body[b] (device int32 [n_bodies], NULL = identity) lies inside it (inclusive on both sides, like the
 * reference's >= / <=).  scene [n_clouds][n_pts][3]; mask uint8 [n_bodies][n_pts]; count int32 [n_bodies] = points
 * kept (the reference's `inds.any()` / `inds.sum()`); bbox float [n_bodies][6] = min xyz, max xyz, or NULL. */
int ehb_scene_crop(ehb_ctx* ctx, const float* verts, int n_bodies, int n_verts, const float* scene, int n_pts,
                   const int32_t* img_of_body, uint8_t* mask, int32_t* count, float* bbox, void* stream);

/* utils/pose_utils.py:11-73 compute_similarity_transform_batch (mask == NULL) and :62-105 the *_with_vis_mask variant
 * (mask [n_problems][n_points][3], multiplied into both point sets before the fit, as the reference does): the
 * similarity transform (scale, R, t) taking s1 onto s2 by orthogonal Procrustes, one problem per (sample, body).
 * s1, s2 [n_problems][n_points][3]; s1_hat [n_problems][n_points][3] = scale R s1 + t; err [n_problems][n_points] =
 * |s1_hat - s2| (the per-joint PA-MPJPE of reconstruction_error, :108-115).  Either output may be NULL. */
int ehb_procrustes_align(ehb_ctx* ctx, const float* s1, const float* s2, const float* mask, int n_problems, int n_points,
                         float* s1_hat, float* err, void* stream);

/* The metric block of the evaluation driver (test_egohmr.py:373-494) for one batch, four launches instead of ~60 torch
 * launches, host copies and per-image Python loops:
 *   pred_joints [n_img][n_samples][n_joints][3], pred_verts [n_img][n_samples][n_verts][3]  SMPL outputs, NOT pelvis-aligned
 *   transl [n_img][3]; gt_joints [n_img][n_joints][3]; gt_verts [n_img][n_verts][3]; focal, cam_cx, cam_cy [n_img] (pixels)
 * outputs:
 *   joint_vis [n_img][n_joints], vert_vis [n_img][n_verts] uint8: the ground truth projects into the 1920 x 1080 frame (:375-388)
 *   errors [n_img][n_samples][9] = G-MPJPE mean, visible sum, invisible sum (:398-406) | MPJPE (pelvis-aligned) mean, vis, invis
 *          (:408-416) | V2V mean, vis, invis (:438-447)
 *   diversity [n_img][6] = per-joint std over the samples: all / visible / invisible joints (:449-468) | APD: all / vis / invis
 *          (:470-494); empty joint sets give NaN, as the reference's mean over nothing does.
 * PA-MPJPE (:418-436) is ehb_procrustes_align on the pelvis-aligned joints. */
int ehb_eval_metrics(ehb_ctx* ctx, const float* pred_joints, const float* pred_verts, const float* transl,
                     const float* gt_joints, const float* gt_verts, const float* focal, const float* cam_cx,
                     const float* cam_cy, int n_img, int n_samples, int n_joints, int n_verts, uint8_t* joint_vis,
                     uint8_t* vert_vis, float* errors, float* diversity, void* stream);

/* ResNet-50 image encoder (models/resnet.py:100-150: torchvision-style v1.5 bottlenecks, global average pool, no fc),
 * one entry per convolution with the BatchNorm2d that follows it, in network order: the stem, then for every
 * bottleneck conv1, conv2, conv3 and — in the first block of a stage — downsample.  HOST. */
typedef struct {
  int32_t cout, cin, kh, kw, stride, pad;
  const float* weight;     /* [cout][cin][kh][kw]  "...convN.weight" / "...downsample.0.weight" */
  const float* bn_weight;  /* [cout]  "...bnN.*" / "...downsample.1.*" */
  const float* bn_bias;
  const float* bn_mean;
  const float* bn_var;
  float bn_eps;
} ehb_conv_bn;
typedef struct {
  const ehb_conv_bn* convs;
  int32_t n_convs;         /* 1 + sum_b (3 + (b is the first block of its stage)) = 53 for ResNet-50 */
  int32_t blocks[4];       /* bottlenecks per stage: 3, 4, 6, 3 */
} ehb_resnet_weights;
int ehb_resnet_load(ehb_ctx* ctx, const ehb_resnet_weights* w);
/* backbone(img) (egohmr.py:183): img [n][3][h][w] fp32 NCHW (ImageNet-normalised crops, 224 x 224) -> feats [n][2048].
 * Every convolution is a tcgen05 GEMM with fp32-class accuracy (fp16 hi/lo operands, fp32 accumulate); BatchNorm is
 * folded, ReLU / residual adds are fused into the GEMM epilogues. */
int ehb_resnet_forward(ehb_ctx* ctx, const float* img, int n, int h, int w, float* feats, void* stream);

/* nn.Linear (+ ReLU) of the small step-invariant heads (FCHeadBeta egohmr.py:673-679, TranslEnc :685-691) in fp32 FFMA:
 * y[m][n] = x[m][k] . w_t[k][n] + bias[n]; w_t is the TRANSPOSED nn.Linear weight; bias may be NULL; relu != 0 clamps. */
int ehb_linear_f32(ehb_ctx* ctx, const float* x, const float* w_t, const float* bias, int m, int n, int k, int relu,
                   float* y, void* stream);

/* The 1-nearest-neighbour squared distances behind utils/pytorch3d_chamfer_distance.py::chamfer_distance (:160-164,
 * pytorch3d knn_points(K=1); contact score of test_egohmr.py:496-506): for pair i, every point of query cloud
 * q[q_index ? q_index[i] : i] ([n_q][3]) against reference cloud r[r_index ? r_index[i] : i] ([n_r][3]);
 * out [n_pairs][n_q].  Index arrays are device int32 [n_pairs] or NULL. */
int ehb_nn_dist_sq(ehb_ctx* ctx, const float* q, const int32_t* q_index, int n_q, const float* r, const int32_t* r_index,
                   int n_r, int n_pairs, float* out, void* stream);

/* utils/konia_transform.py:316-339 rotation_matrix_to_angle_axis: R [n][3][3] -> aa [n][3] (guide_coll / eval_coll feed
 * `full_pose` to the collision model as axis-angle, egohmr.py:495,540). */
int ehb_rotmat_to_angle_axis(ehb_ctx* ctx, const float* R, int n, float* aa, void* stream);

/* Backward of egohmr.py:528-540 (x_t*std+mean -> rot6d -> SMPL -> axis-angle), i.e. what torch.autograd.grad computes in
 * guide_coll (:562): given g_verts [n_bodies][V][3], g_joints [n_bodies][45][3], g_aa [n_bodies][24][3] (any may be
 * NULL = zero) for the bodies of ehb_set_bodies, writes grad_x [n_bodies][144].  Like the reference (which rebinds x_t to
 * x_t*std+mean before autograd.grad, :523-528,562) the gradient is w.r.t. the de-normalised pose.  betas [n_img][n_betas]. */
int ehb_smpl_backward(ehb_ctx* ctx, const float* x_t, const float* betas, const float* g_verts, const float* g_joints,
                      const float* g_aa, float* grad_x, void* stream);

/* Diagnostics.  gemm_mode 0 = tcgen05 fp16x3 kernel on CTA pairs (cta_group::2), transposed product: weights on the M
 * side, N = 240 activation rows = 10 (body, pass) slots, no pad rows (product path); 1 = fp32 FFMA check path (tests only);
 * 2 = tcgen05 fp16x3 row-major kernel on single CTAs (cta_group::1, bring-up comparison); 3 = the row-major CTA-pair
 * kernel (activations on the M side, 5 slots + 8 pad rows per 128-row tile: same bits as mode 0, 6 % more MMA time). */
int ehb_debug_set_gemm_mode(ehb_ctx* ctx, int gemm_mode);
/* ResNet 3x3 convolutions: 1 = implicit GEMM through 4-D TMA boxes (product path), 0 = explicit im2col matrix + the
 * same GEMM (bring-up comparison; both must give identical bits). */
int ehb_debug_set_resnet_mode(ehb_ctx* ctx, int implicit_gemm);
/* Hidden-layer launches with (1, default) or without (0) programmatic dependent launch (bring-up comparison). */
int ehb_debug_set_pdl(ehb_ctx* ctx, int on);
/* Input graph-convolution layer (K2): the fp32 FFMA kernel (0, default) or the variant with the 24x24 joint mix on the
 * tensor pipe (1; same result to fp32 rounding, same time: the kernel is bound by assembling the per-joint features). */
int ehb_debug_set_input_mode(ehb_ctx* ctx, int umma);
/* Hidden layers of a reverse step: one launch per layer (0, default) or ONE persistent launch ordered by per-(layer, row
 * group) completion counters (1: same bits, measured 7-8 % slower; kept as an experiment). */
int ehb_debug_set_k1_fused(ehb_ctx* ctx, int on);
/* ResNet convolution GEMMs: k-blocks (of 64 operand columns) chained into one tensor-memory accumulation before the
 * epilogue takes the partial sum over in fp32 registers (0 = the whole contraction in one accumulator). */
int ehb_debug_set_conv_kc(ehb_ctx* ctx, int kc);
/* The convolution GEMM primitive alone, for unit tests of its numerics: out[m][n] = a[m][k] . w[n][k]^T in the library's
 * error-compensated fp16 hi/lo scheme (operands scaled by a_scale / w_scale, powers of two).  All pointers HOST;
 * synchronous.  n % 64 == 0, k % 64 == 0. */
int ehb_debug_gemm_hl(ehb_ctx* ctx, const float* a, const float* w, int m, int n, int k, float a_scale, float w_scale,
                      int kc, float* out);
/* Returns 1 (and clears it) if any fp16 operand overflowed since the last call; synchronises `stream`. */
int ehb_check_overflow(ehb_ctx* ctx, void* stream);
/* Stream-ordered, non-blocking copy of the same flag into *host_flag (PINNED host memory): the caller reads it after
 * any later synchronisation point, so a sequence of sampling calls needs no host sync of its own; graph-capturable.
 * The device flag stays set until ehb_check_overflow clears it. */
int ehb_overflow_flag_async(ehb_ctx* ctx, int32_t* host_flag, void* stream);
/* Runs only hidden layer `layer` (1-based index into ehb_gcn_weights.layers) `iters` times on the current
 * activations and returns the average device time in ms through *ms (bench.py's roofline leg). */
int ehb_time_hidden_layer(ehb_ctx* ctx, int layer, int iters, float* ms, void* stream);

/* Same for any stage of one reverse step on the context's current activations: stage 0 = folded input layer,
 * 1 .. 2*n_blocks = hidden layers, 2*n_blocks+1 = output layer + fuse-select + sampler update (noise/grad = NULL).
 * x_t / x_prev / x0 as in ehb_denoise_step. */
int ehb_time_stage(ehb_ctx* ctx, int stage, int step, const float* x_t, float* x_prev, float* x0, int iters, float* ms,
                   void* stream);

/* Device pointer + size of an internal activation buffer (tests / bring-up only):
 * which 0 = fp32 block-boundary activations [rows_pad][hid], 1/2 = fp16 [hi|lo] operand ping/pong [rows_pad][2*hid]. */
int ehb_debug_get_buffer(ehb_ctx* ctx, int which, void** ptr, uint64_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* EGOHMR_B200_H_ */

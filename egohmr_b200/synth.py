"""Seeded synthetic weights, SMPL model and input batches (numpy only).

There is no checkpoint, SMPL ``.pkl`` or EgoBody data offline, so tests, ``bench.py`` and the golden-vector generator
all draw the same tensors from here.  Everything is keyed by the reference's own ``state_dict`` names
(``EgoHMR.state_dict()``, models/egohmr/egohmr.py:29-137) and batch-dict schema (dataloaders/egobody_dataset.py:241-277)
so a real checkpoint / real batch can be dropped in unchanged.  numpy's ``default_rng`` (PCG64) stream is stable across
platforms, which is what lets the committed golden vectors be regenerated bit-for-bit.
"""
import numpy as np

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
# smplx/vertex_ids.py['smplh'] in VertexJointSelector order: nose, reye, leye, rear, lear, LBigToe, LSmallToe, LHeel,
# RBigToe, RSmallToe, RHeel, l{thumb,index,middle,ring,pinky}, r{thumb,index,middle,ring,pinky}  (assumption recorded in
# oracle/smpl.py: restated from upstream memory, not verifiable offline).
SMPL_EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                         2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]
# utils/other_utils.py:86-108
SMPL_EDGES = [(0, 1), (0, 2), (0, 3), (1, 4), (2, 5), (3, 6), (4, 7), (5, 8), (6, 9), (7, 10), (8, 11), (9, 12),
              (9, 13), (9, 14), (12, 15), (13, 16), (14, 17), (16, 18), (17, 19), (18, 20), (19, 21), (20, 22), (21, 23)]

IMG_DIM, SCENE_DIM, TRANSL_DIM, CAM_DIM, XFEAT_DIM, TEMB_DIM = 2048, 512, 128, 6, 512, 512
COND_DIM = IMG_DIM + SCENE_DIM + TRANSL_DIM + CAM_DIM  # 2694
IN_DIM = COND_DIM + XFEAT_DIM + TEMB_DIM  # 3718


def skeleton_adjacency():
    """The fixed `adj` of ModulatedGCN exactly as egohmr.py:86-94 builds it (scipy-free restatement)."""
    a = np.zeros((24, 24), dtype=np.float32)
    for i, j in SMPL_EDGES:
        a[i, j] = 1.0
    a = a + a.T * (a.T > a) - a * (a.T > a)  # symmetrise
    rowsum = a.sum(1)
    r_inv = np.where(rowsum > 0, 1.0 / np.maximum(rowsum, 1e-30), 0.0).astype(np.float32)
    a = (r_inv[:, None] * a).astype(np.float32)
    eye = np.eye(24, dtype=np.float32)
    return (a * (1 - eye) + eye).astype(np.float32)


def make_smpl_model(seed=0, n_verts=6890, n_betas=10):
    """Synthetic SMPL with the real model's shapes, sparsity (<=4 skinning weights per vertex) and kinematic tree."""
    rng = np.random.default_rng(1000 + seed)
    V = n_verts
    v_template = (rng.uniform(-1, 1, (V, 3)) * np.array([0.35, 0.9, 0.15])).astype(np.float32)
    shapedirs = (rng.normal(0, 0.01, (V, 3, n_betas))).astype(np.float32)
    posedirs = (rng.normal(0, 0.003, (207, V * 3))).astype(np.float32)
    # joint regressor: each joint = convex combination of 32 random vertices
    J_regressor = np.zeros((24, V), dtype=np.float32)
    for j in range(24):
        idx = rng.choice(V, 32, replace=False)
        w = rng.uniform(0.1, 1.0, 32)
        J_regressor[j, idx] = (w / w.sum()).astype(np.float32)
    lbs_weights = np.zeros((V, 24), dtype=np.float32)
    for v in range(V):
        k = int(rng.integers(1, 5))
        idx = rng.choice(24, k, replace=False)
        w = rng.uniform(0.1, 1.0, k)
        lbs_weights[v, idx] = (w / w.sum()).astype(np.float32)
    extra = [vid % V for vid in SMPL_EXTRA_VERTEX_IDS]
    return {
        "v_template": v_template, "shapedirs": shapedirs, "posedirs": posedirs, "J_regressor": J_regressor,
        "lbs_weights": lbs_weights, "parents": np.array(SMPL_PARENTS, dtype=np.int32),
        "extra_vertex_ids": np.array(extra, dtype=np.int32),
        "init_betas": rng.normal(0, 0.5, (n_betas,)).astype(np.float32),
    }


def _linear(rng, sd, name, fan_in, fan_out, bias=True, w_scale=1.0):
    bound = 1.0 / np.sqrt(fan_in)
    sd[name + ".weight"] = (rng.uniform(-bound, bound, (fan_out, fan_in)) * w_scale).astype(np.float32)
    if bias:
        sd[name + ".bias"] = rng.uniform(-bound, bound, (fan_out,)).astype(np.float32)


def _bn(rng, sd, name, c, gamma=(0.5, 1.5)):
    sd[name + ".weight"] = rng.uniform(gamma[0], gamma[1], c).astype(np.float32)
    sd[name + ".bias"] = rng.normal(0, 0.1, c).astype(np.float32)
    sd[name + ".running_mean"] = rng.normal(0, 0.1, c).astype(np.float32)
    sd[name + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    sd[name + ".num_batches_tracked"] = np.array(1, dtype=np.int64)


def _gconv(rng, sd, name, fan_in, fan_out):
    """ModulatedGraphConv init (modulated_gcn_conv.py:21-36) but with adj2 ~ N(0, 0.05) so adjacency bugs show."""
    bw = 1.414 * np.sqrt(6.0 / (fan_in + fan_out))  # xavier_uniform over the trailing two dims (2 acts as a batch dim)
    sd[name + ".W"] = rng.uniform(-bw, bw, (2, fan_in, fan_out)).astype(np.float32)
    bm = 1.414 * np.sqrt(6.0 / (24 + fan_out))
    sd[name + ".M"] = rng.uniform(-bm, bm, (24, fan_out)).astype(np.float32)
    sd[name + ".adj2"] = rng.normal(0, 0.05, (24, 24)).astype(np.float32)
    stdv = 1.0 / np.sqrt(fan_out)
    sd[name + ".bias"] = rng.uniform(-stdv, stdv, (fan_out,)).astype(np.float32)


def positional_encoding_table(d_model=512, max_len=5000):
    """PositionalEncoding.pe (egohmr.py:614-621), fp32 op order preserved; shape [max_len, 1, d_model]."""
    pe = np.zeros((max_len, d_model), dtype=np.float32)
    position = np.arange(0, max_len, dtype=np.float32)[:, None]
    div_term = np.exp(np.arange(0, d_model, 2).astype(np.float32) * np.float32(-np.log(10000.0) / d_model))
    pe[:, 0::2] = np.sin(position * div_term)
    pe[:, 1::2] = np.cos(position * div_term)
    return pe[:, None, :]


RESNET50_LAYERS = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]


def make_state_dict(seed=0, hid=1024, n_blocks=4, with_encoders=True, n_betas=10, init_betas=None):
    """Random `EgoHMR.state_dict()` (reference key names).  BN running stats and adj2 are randomised."""
    rng = np.random.default_rng(2000 + seed)
    sd = {}
    # ---- denoiser
    _gconv(rng, sd, "diffusion_model.gconv_input.0.gconv", IN_DIM, hid)
    _bn(rng, sd, "diffusion_model.gconv_input.0.bn", hid)
    for b in range(n_blocks):
        for g in (1, 2):
            _gconv(rng, sd, f"diffusion_model.gconv_layers.{b}.gconv{g}.gconv", hid, hid)
            _bn(rng, sd, f"diffusion_model.gconv_layers.{b}.gconv{g}.bn", hid)
    _gconv(rng, sd, "diffusion_model.gconv_output", hid, 6)
    _linear(rng, sd, "input_process.poseEmbedding", 6, XFEAT_DIM)
    pe = positional_encoding_table(TEMB_DIM)
    sd["sequence_pos_encoder.pe"] = pe
    sd["embed_timestep.sequence_pos_encoder.pe"] = pe
    _linear(rng, sd, "embed_timestep.time_embed.0", TEMB_DIM, TEMB_DIM)
    _linear(rng, sd, "embed_timestep.time_embed.2", TEMB_DIM, TEMB_DIM)
    # ---- small heads
    _linear(rng, sd, "transl_enc.layers.0", 3, 64)
    _linear(rng, sd, "transl_enc.layers.2", 64, TRANSL_DIM)
    _linear(rng, sd, "beta_layer.layers.0", COND_DIM, 1024)
    _linear(rng, sd, "beta_layer.layers.2", 1024, n_betas, w_scale=0.2)
    sd["beta_layer.init_betas"] = (np.zeros((1, n_betas), np.float32) if init_betas is None
                                   else np.asarray(init_betas, np.float32).reshape(1, n_betas))
    if not with_encoders:
        return sd
    # ---- ResPointNet (models/respointnet.py:13-27)
    h = 256
    _linear(rng, sd, "scene_enc.fc_pos_0", 3, 2 * h)
    for b in range(4):
        _linear(rng, sd, f"scene_enc.block_{b}.fc_0", 2 * h, h)
        _linear(rng, sd, f"scene_enc.block_{b}.fc_1", h, h, w_scale=0.5)
        _linear(rng, sd, f"scene_enc.block_{b}.shortcut", 2 * h, h, bias=False)
    _linear(rng, sd, "scene_enc.fc_c", h, SCENE_DIM)
    # ---- ResNet-50 (models/resnet.py:100-150); the last BN of every bottleneck gets a small gamma so the
    # residual stream stays O(1) with random running stats
    def conv(name, cout, cin, k):
        n = k * k * cout
        sd[name + ".weight"] = rng.normal(0, np.sqrt(2.0 / n), (cout, cin, k, k)).astype(np.float32)

    conv("backbone.conv1", 64, 3, 7)
    _bn(rng, sd, "backbone.bn1", 64)
    inplanes = 64
    for li, (planes, blocks, stride) in enumerate(RESNET50_LAYERS, start=1):
        for bi in range(blocks):
            p = f"backbone.layer{li}.{bi}"
            conv(p + ".conv1", planes, inplanes, 1)
            _bn(rng, sd, p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3)
            _bn(rng, sd, p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1)
            _bn(rng, sd, p + ".bn3", planes * 4, gamma=(0.1, 0.3))
            if bi == 0:
                conv(p + ".downsample.0", planes * 4, inplanes, 1)
                _bn(rng, sd, p + ".downsample.1", planes * 4, gamma=(0.3, 0.6))
            inplanes = planes * 4
    return sd


def add_nonlocal(sd, seed=0, hid=1024):
    """NONLocalBlock2D parameters (`diffusion_model.non_local.*`, gcn_nonlocal_layer=True) from their own generator, so
    the rest of `make_state_dict` is unchanged.  The reference initialises W's BatchNorm to zero (an identity block);
    here it is random so the block does something."""
    rng = np.random.default_rng(2500 + seed)
    inter, p = hid // 2, "diffusion_model.non_local"
    for name, cin, cout in (("g", hid, inter), ("theta", hid, inter), ("phi", hid, inter), ("W.0", inter, hid)):
        sd[f"{p}.{name}.weight"] = rng.normal(0, np.sqrt(2.0 / cin) * (0.25 if name in ("theta", "phi") else 1.0),
                                               (cout, cin, 1, 1)).astype(np.float32)
        sd[f"{p}.{name}.bias"] = rng.normal(0, 0.05, cout).astype(np.float32)
    _bn(rng, sd, f"{p}.W.1", hid, gamma=(0.2, 0.6))
    return sd


def body_rep_stats(seed=0):
    """preprocess_stats.npz stand-in (test_egohmr.py:109-111): Xmean, Xstd of the 144-d rot6d representation."""
    rng = np.random.default_rng(3000 + seed)
    return rng.normal(0, 0.3, 144).astype(np.float32), rng.uniform(0.5, 1.5, 144).astype(np.float32)


def make_batch(seed=0, n_img=2, n_pts=1024):
    """Synthetic batch in the reference schema."""
    rng = np.random.default_rng(4000 + seed)
    f32 = np.float32
    kp = rng.uniform(0, 1000, (n_img, 25, 3)).astype(f32)
    kp[:, :, 2] = (rng.uniform(0, 1, (n_img, 25)) < 0.6).astype(f32) * rng.uniform(0.2, 1.0, (n_img, 25)).astype(f32)
    transl = (np.array([0, 0, 3.0]) + rng.normal(0, 0.1, (n_img, 3))).astype(f32)
    return {
        "img": rng.normal(0, 1, (n_img, 3, 224, 224)).astype(f32),
        "orig_keypoints_2d": kp,
        "fx": rng.uniform(0.9, 1.1, n_img).astype(f32),
        "cam_cx": np.full(n_img, 960.0, f32),
        "cam_cy": np.full(n_img, 540.0, f32),
        "box_center": rng.uniform(0, 1000, (n_img, 2)).astype(f32),
        "box_size": rng.uniform(100, 600, n_img).astype(f32),
        "smpl_params": {"transl": transl},
        "scene_pcd_verts_full": (rng.uniform(-1, 1, (n_img, n_pts, 3)) + np.array([0, 0, 3.0])).astype(f32),
    }


def make_gt(seed=0, n_img=2):
    """Ground-truth keys `model.compute_loss` reads (dataloaders/egobody_dataset.py:241-277, egohmr.py:320-325):
    merge into a `make_batch` dict (smpl_params gains global_orient / body_pose / betas next to transl)."""
    rng = np.random.default_rng(6000 + seed)
    f32 = np.float32
    kp3d = rng.normal(0, 0.3, (n_img, 24, 3)).astype(f32)
    return {
        "keypoints_3d": kp3d,
        "keypoints_3d_full": (kp3d + np.array([0, 0, 3.0])).astype(f32),
        "smpl_params": {"global_orient": rng.normal(0, 0.5, (n_img, 3)).astype(f32),
                        "body_pose": rng.normal(0, 0.3, (n_img, 69)).astype(f32),
                        "betas": rng.normal(0, 1.0, (n_img, 10)).astype(f32)},
        "smpl_params_is_axis_angle": {"global_orient": np.ones(n_img, bool), "body_pose": np.ones(n_img, bool),
                                      "betas": np.zeros(n_img, bool)},
        "gender": (rng.uniform(0, 1, n_img) < 0.5).astype(np.int64),
    }


def merge_gt(batch, gt):
    out = dict(batch)
    out["smpl_params"] = {**batch["smpl_params"], **gt["smpl_params"]}
    for k, v in gt.items():
        if k != "smpl_params":
            out[k] = v
    return out


def make_noise(seed, n_chains, n_bodies_per_chain, n_steps):
    """Pre-drawn noise in the reference's RNG consumption order: per chain, randn(bs,144) then one draw per step."""
    rng = np.random.default_rng(5000 + seed)
    return rng.normal(0, 1, (n_chains, n_steps + 1, n_bodies_per_chain, 144)).astype(np.float32)

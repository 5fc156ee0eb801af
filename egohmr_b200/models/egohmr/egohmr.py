"""`EgoHMR` with the reference's constructor / forward / state_dict interface (models/egohmr/egohmr.py:29-303),
re-architected for the sampling hot path:

* the step-invariant part of `forward` (image / scene / translation / camera encoders, visibility, beta head,
  egohmr.py:181-223,262-265) runs once per batch in PyTorch and is cached (`prepare`);
* everything that depends on x_t / t — the 10-layer Modulated GCN evaluated twice (image-conditioned and image-masked,
  :236-246), the fuse-select (:248-254), de-normalisation + rot6d (:258-260), SMPL (:276) — runs in
  libegohmr_b200.so through `self.engine`.

Parameters keep the reference's names, so `load_state_dict(torch.load(ckpt)['state_dict'], strict=False)`
(test_egohmr.py:125-126) ingests a reference checkpoint unchanged.
"""
import numpy as np
import torch
import torch.nn as nn

from ... import smpl as smpl_mod
from ... import synth
from ...engine import Engine
from ...utils.geometry import perspective_projection
from ..fast_encoders import FoldedResNet50, SplitPointNet
from ..resnet import ResNet50Features
from ..respointnet import ResnetPointnet

OPENPOSE_TO_SMPL = [8, 12, 9, 8, 13, 10, 8, 14, 11, 8, 14, 11, 0, 5, 2, 0, 5, 2, 6, 3, 7, 4, 7, 4]          # :111
OPENPOSE_TO_SMPL_LOOSEN = [8, 13, 10, 8, 13, 10, 8, 14, 11, 8, 14, 11, 1, 5, 2, 0, 5, 2, 6, 3, 7, 4, 7, 4]   # :114


class _GConvParams(nn.Module):
    """Parameter container for one ModulatedGraphConv (modulated_gcn_conv.py:16-36); the math lives in the kernels."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.W = nn.Parameter(torch.empty(2, in_dim, out_dim))
        self.M = nn.Parameter(torch.empty(24, out_dim))
        self.adj2 = nn.Parameter(torch.full((24, 24), 1e-6))
        self.bias = nn.Parameter(torch.empty(out_dim))
        nn.init.xavier_uniform_(self.W.data, gain=1.414)
        nn.init.xavier_uniform_(self.M.data, gain=1.414)
        stdv = 1.0 / np.sqrt(out_dim)
        self.bias.data.uniform_(-stdv, stdv)


class _GraphConvParams(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.gconv = _GConvParams(in_dim, out_dim)
        self.bn = nn.BatchNorm1d(out_dim)


class _ResGraphConvParams(nn.Module):
    def __init__(self, hid):
        super().__init__()
        self.gconv1 = _GraphConvParams(hid, hid)
        self.gconv2 = _GraphConvParams(hid, hid)


class _NonLocalParams(nn.Module):
    """Parameter container with NONLocalBlock2D's names and shapes (nets/non_local_embedded_gaussian.py:36-54)."""

    def __init__(self, hid):
        super().__init__()
        inter = hid // 2
        self.g = nn.Conv2d(hid, inter, 1)
        self.W = nn.Sequential(nn.Conv2d(inter, hid, 1), nn.BatchNorm2d(hid))
        nn.init.constant_(self.W[1].weight, 0)
        nn.init.constant_(self.W[1].bias, 0)
        self.theta = nn.Conv2d(hid, inter, 1)
        self.phi = nn.Conv2d(hid, inter, 1)


class ModulatedGCNParams(nn.Module):
    """state_dict-compatible skeleton of ModulatedGCN (modulated_gcn.py:61-97)."""

    def __init__(self, in_dim, hid_dim, out_dim, num_layers, nonlocal_layer=False):
        super().__init__()
        self.gconv_input = nn.Sequential(_GraphConvParams(in_dim, hid_dim))
        self.gconv_layers = nn.Sequential(*[_ResGraphConvParams(hid_dim) for _ in range(num_layers)])
        self.gconv_output = _GConvParams(hid_dim, out_dim)
        self.nonlocal_layer = nonlocal_layer
        if nonlocal_layer:
            self.non_local = _NonLocalParams(hid_dim)


class PositionalEncoding(nn.Module):
    def __init__(self, d_model, max_len=5000):
        super().__init__()
        self.register_buffer("pe", torch.from_numpy(synth.positional_encoding_table(d_model, max_len)))


class TimestepEmbedder(nn.Module):
    def __init__(self, latent_dim, sequence_pos_encoder):
        super().__init__()
        self.sequence_pos_encoder = sequence_pos_encoder
        self.time_embed = nn.Sequential(nn.Linear(latent_dim, latent_dim), nn.SiLU(), nn.Linear(latent_dim, latent_dim))

    def forward(self, timesteps):
        return self.time_embed(self.sequence_pos_encoder.pe[timesteps]).permute(1, 0, 2)


class InputProcess(nn.Module):
    def __init__(self, input_dim, latent_dim):
        super().__init__()
        self.poseEmbedding = nn.Linear(input_dim, latent_dim)


class FCHeadBeta(nn.Module):
    def __init__(self, in_dim, init_betas=None):
        super().__init__()
        self.layers = nn.Sequential(nn.Linear(in_dim, 1024), nn.ReLU(inplace=False), nn.Linear(1024, 10))
        nn.init.xavier_uniform_(self.layers[2].weight, gain=0.02)
        if init_betas is None:
            try:  # egohmr.py:669 reads the mean shape from CWD
                init_betas = np.load("data/smpl_mean_params.npz")["shape"].astype(np.float32)
            except (FileNotFoundError, OSError):
                init_betas = np.zeros(10, np.float32)
        self.register_buffer("init_betas", torch.from_numpy(np.asarray(init_betas, np.float32).reshape(1, 10)))

    def forward(self, feats, pred_pose=None):
        return self.layers(feats) + self.init_betas


class TranslEnc(nn.Module):
    def __init__(self, in_dim=3, out_dim=128):
        super().__init__()
        self.layers = nn.Sequential(nn.Linear(in_dim, 64), nn.ReLU(inplace=False), nn.Linear(64, out_dim))

    def forward(self, x):
        return self.layers(x)


class EgoHMR(nn.Module):
    def __init__(self, cfg, device=None, body_rep_mean=None, body_rep_std=None,
                 with_focal_length=False, with_bbox_info=False, with_cam_center=False,
                 scene_feat_dim=512, scene_type="whole_scene", scene_cano=False,
                 weight_loss_v2v=0, weight_loss_keypoints_3d=0, weight_loss_keypoints_3d_full=0,
                 weight_loss_keypoints_2d_full=0, weight_loss_betas=0, weight_loss_body_pose=0,
                 weight_loss_global_orient=0, weight_loss_pose_6d_ortho=0, weight_coap_penetration=0,
                 start_coap_epoch=0, cond_mask_prob=0, only_mask_img_cond=False, diffusion_blk=4, gcn_dropout=0.0,
                 gcn_nonlocal_layer=False, gcn_hid_dim=1024, pelvis_vis_loosen=False, diffuse_fuse=False,
                 smpl_model=None, collision_model=None):
        super().__init__()
        if gcn_nonlocal_layer and gcn_hid_dim % 128 != 0:
            raise ValueError("gcn_nonlocal_layer needs gcn_hid_dim to be a multiple of 128")
        self.cfg = cfg
        self.device = torch.device(device) if device is not None else torch.device("cuda", 0)
        self.with_focal_length, self.with_bbox_info, self.with_cam_center = with_focal_length, with_bbox_info, with_cam_center
        self.body_rep_mean, self.body_rep_std = body_rep_mean, body_rep_std
        self.diffuse_feat_dim = 6
        self.cond_mask_prob, self.only_mask_img_cond, self.diffuse_fuse = cond_mask_prob, only_mask_img_cond, diffuse_fuse
        self.scene_type, self.scene_cano = scene_type, scene_cano
        self.hid, self.n_blocks = gcn_hid_dim, diffusion_blk

        self.input_process_out_dim = 512
        self.input_process = InputProcess(6, 512)
        self.timestep_embed_dim = 512
        self.sequence_pos_encoder = PositionalEncoding(512)
        self.embed_timestep = TimestepEmbedder(512, self.sequence_pos_encoder)
        self.backbone = ResNet50Features()
        self.scene_enc = ResnetPointnet(out_dim=scene_feat_dim, hidden_dim=256)
        self.transl_enc = TranslEnc(3, 128)
        self.img_dim = cfg.MODEL.BACKBONE.OUT_CHANNELS
        ctx_dim = self.img_dim + (1 if with_focal_length else 0) + (3 if with_bbox_info else 0) + \
            (2 if with_cam_center else 0) + scene_feat_dim + 128  # egohmr.py:76-83
        self.cond_dim = ctx_dim
        self.register_buffer("adj", torch.from_numpy(synth.skeleton_adjacency()), persistent=False)  # :86-94
        self.diffusion_model = ModulatedGCNParams(ctx_dim + 512 + 512, gcn_hid_dim, 6, diffusion_blk, gcn_nonlocal_layer)
        init_betas = None if smpl_model is None else smpl_model.get("init_betas")
        self.beta_layer = FCHeadBeta(ctx_dim, init_betas)

        self.engine = Engine(self.device.index or 0)
        self.smpl = smpl_mod.create("data/smpl", model_type="smpl", gender="neutral", smpl_model=smpl_model,
                                    engine=self.engine)  # :105
        self.smpl_male = self.smpl if smpl_model is not None else smpl_mod.create("data/smpl", gender="male")
        self.smpl_female = self.smpl if smpl_model is not None else smpl_mod.create("data/smpl", gender="female")
        self.openpose_to_smpl = OPENPOSE_TO_SMPL_LOOSEN if pelvis_vis_loosen else OPENPOSE_TO_SMPL
        # collision guidance is a pluggable callable with COAP's signature: `attach_coap` upstream
        self.collision_model = collision_model
        self.weight_coap_penetration, self.start_coap_epoch = weight_coap_penetration, start_coap_epoch
        self.loss_weights = {"v2v": weight_loss_v2v, "keypoints_3d": weight_loss_keypoints_3d,
                             "keypoints_3d_full": weight_loss_keypoints_3d_full,
                             "keypoints_2d_full": weight_loss_keypoints_2d_full, "betas": weight_loss_betas,
                             "body_pose": weight_loss_body_pose, "global_orient": weight_loss_global_orient,
                             "pose_6d_ortho": weight_loss_pose_6d_ortho}
        self.to(self.device)
        self._weights_dirty = True
        self._synced_version = None
        self._cond_key = None
        self._cond = None
        self._cond_features = None
        self._temb_key = None
        self._bodies_key = None
        self._bodies_idx = None
        # The fp16 hi/lo operands have a finite range; a checkpoint that exceeds it must fail loudly (FloatingPointError),
        # not return Inf/NaN.  "sync": the flag is read back at the end of every eager sampling call (one host sync per
        # call).  "deferred" (default): every call enqueues a stream-ordered copy of the flag into pinned host memory and
        # the flag is examined without blocking at the start of the next call, in eval_coll / compute_loss (which the
        # reference driver runs right after each val_losses call) and by poll_overflow() — the ten per-sample calls of the
        # driver's loop (test_egohmr.py:251-255) then need no host sync of their own.
        self.overflow_check = "deferred"
        # val_losses draws the chains of several samples ahead and runs them as one batch (gaussian_diffusion.py::val_losses):
        # "auto" = as many as the previous batch received calls, an int = that many, 0 = off (one chain per call)
        self.samples_ahead = "auto"
        self._ovf_host = None
        self._ovf_event = None
        self.native_image_enc = True   # ResNet-50 on the tcgen05 convolution GEMMs (K9, fp32-class); False = cuDNN
        self.native_scene_enc = True   # ResPointNet on the tcgen05 linear kernel (K7); False = PyTorch/cuBLAS form

    # ------------------------------------------------------------------ weight ingestion
    def load_state_dict(self, state_dict, strict=True, **kw):
        own = self.state_dict()
        filtered = {k: v for k, v in state_dict.items() if k in own}  # smpl.* / coap.* buffers live elsewhere here
        res = super().load_state_dict(filtered, strict=False, **kw)
        self._weights_dirty = True
        self._cond_key = None
        if strict:
            missing = [k for k in own if k not in state_dict]
            if missing:
                raise RuntimeError(f"missing keys in state_dict: {missing[:5]}...")
        return res

    def _weights_version(self):
        """Sum of the autograd version counters of every parameter and buffer: any torch in-place update (load_state_dict on
        a sub-module, an optimizer step, .copy_) changes it."""
        return sum(t._version for t in self.parameters()) + sum(t._version for t in self.buffers())

    def _sync_engine(self):
        """(Re)pack the module's weights for the kernels when they changed since the last sampling call."""
        version = self._weights_version()
        if not self._weights_dirty and version == self._synced_version:
            return
        sd = {k: v for k, v in self.state_dict().items()
              if k.startswith("diffusion_model.") or k.startswith("input_process.")}
        bn_eps = self.diffusion_model.gconv_input[0].bn.eps
        self.engine.load_gcn(sd, self.adj, self.hid, self.n_blocks, self.diffuse_fuse, self.img_dim, self.cond_dim,
                             512, 512, bn_eps=bn_eps, mask_all_cond=not self.only_mask_img_cond)
        if self.diffusion_model.nonlocal_layer:
            self.engine.load_nonlocal(sd, bn_eps=self.diffusion_model.non_local.W[1].eps)
        if not self.engine.smpl_loaded:
            self.engine.load_smpl(self.smpl.model)
        mean = self.body_rep_mean if self.body_rep_mean is not None else torch.zeros(144)
        std = self.body_rep_std if self.body_rep_std is not None else torch.ones(144)
        self.engine.set_norm(torch.as_tensor(mean).detach().float().cpu().numpy(),
                             torch.as_tensor(std).detach().float().cpu().numpy())
        self._fast_backbone = FoldedResNet50(self.backbone, self.engine)
        self._fast_scene_enc = SplitPointNet(self.scene_enc)   # PyTorch form (kept for comparison / odd shapes)
        # the two small heads as (transposed weight, bias) pairs for the library's fp32 GEMM (no cuBLAS on the pass)
        lin = lambda m: (m.weight.detach().t().contiguous().float(), m.bias.detach().contiguous().float())
        self._heads = {"transl0": lin(self.transl_enc.layers[0]), "transl2": lin(self.transl_enc.layers[2]),
                       "beta0": lin(self.beta_layer.layers[0]), "beta2": lin(self.beta_layer.layers[2])}
        self.engine.load_pointnet({k: v for k, v in self.state_dict().items() if k.startswith("scene_enc.")})
        self.engine.load_resnet({k: v for k, v in self.state_dict().items() if k.startswith("backbone.")},
                                bn_eps=self.backbone.bn1.eps)
        self._weights_dirty = False
        self._synced_version = version
        self._cond_key = None
        self._temb_key = None
        self._bodies_key = None

    # ------------------------------------------------------------------ operand-range check
    def note_sampling_call_done(self):
        """Called by the samplers after the last kernel of a sampling call."""
        capturing = torch.cuda.is_current_stream_capturing()
        if self.overflow_check == "sync" and not capturing:
            if self.engine.check_overflow():
                raise FloatingPointError(self._OVF_MSG)
            return
        if self._ovf_host is None:
            if capturing:
                return
            self._ovf_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._ovf_event = torch.cuda.Event()
        # the flag copy is part of a captured pass too; the graph's owner records the event after each replay
        self.engine.overflow_flag_async(self._ovf_host)
        if not capturing:
            self._ovf_event.record()

    _OVF_MSG = ("fp16 operand overflow inside the tensor-core kernels: an activation left the representable range of the "
                "hi/lo operand format (see DESIGN.md, K1 numerics)")

    def poll_overflow(self, sync=False):
        """Raise FloatingPointError if an earlier sampling call overflowed the fp16 operand range.  Non-blocking unless
        `sync` (then it waits for the last enqueued flag copy)."""
        if self._ovf_event is None or torch.cuda.is_current_stream_capturing():
            return
        if sync:
            self._ovf_event.synchronize()
        elif not self._ovf_event.query():
            return
        if int(self._ovf_host[0]) != 0:
            self._ovf_host[0] = 0
            self.engine.check_overflow()    # clears the device flag
            raise FloatingPointError(self._OVF_MSG + " — raised by a sampling call issued earlier (overflow_check='deferred')")

    def validation_setup(self):  # egohmr.py:475-484
        self.training = False
        self.eval()

    # ------------------------------------------------------------------ step-invariant conditioning
    @staticmethod
    def _tkey(t):
        return (id(t), t.data_ptr(), tuple(t.shape), t._version)

    def invalidate(self):
        """Forget the cached step-invariant conditioning: the next `prepare` / sampling call re-runs the encoders.
        `sample_many` calls this itself (one call = one batch).  The per-sample `val_losses` calls of the reference
        driver's loop (test_egohmr.py:251-255) deliberately share the cache.  It is keyed on the IDENTITY of every batch
        tensor the conditioning reads (held through weak references: a new tensor that the caching allocator happens to
        place at a freed batch's address is a different object and misses) plus its data_ptr / shape / `_version`, so a new
        batch or any torch in-place write re-runs the encoders — but a writer that bypasses torch's version counter
        (`.data` writes, DLPack / raw-pointer writers, c10d collectives into a preallocated buffer) must call
        `invalidate()` (or pass `batch['batch_id']`, any hashable that changes per batch and joins the key)."""
        self._cond_key = None

    def batch_token(self, batch, transl, extra=()):
        """-> (key, weakrefs) identifying the content the step-invariant conditioning is computed from."""
        import weakref
        keys = ["img", "scene_pcd_verts_full", "orig_keypoints_2d"]
        if self.with_focal_length or self.with_bbox_info or self.with_cam_center:
            keys.append("fx")
        if self.with_bbox_info:
            keys += ["box_center", "box_size"]
        if self.with_cam_center:
            keys += ["cam_cx", "cam_cy"]
        tensors = [batch[k] for k in keys] + [transl]
        key = tuple(self._tkey(t) for t in tensors) + tuple(extra) + (batch.get("batch_id"),)
        return key, [weakref.ref(t) for t in tensors]

    @staticmethod
    def token_matches(old, new):
        """The cached token still describes the batch at hand: equal keys AND every tensor it was taken from is still alive
        (ids and addresses are recycled once a tensor is freed)."""
        return old is not None and old[0] == new[0] and all(r() is not None and r() is n() for r, n in zip(old[1], new[1]))

    @torch.no_grad()
    def prepare(self, batch, num_samples=1, features=None, force=False):
        """Compute (or reuse) everything in `forward` that does not depend on x_t / t and hand it to the engine.
        `features`: optional dict(img_feats, scene_feats, transl_feat) to bypass the encoders (tests);
        `force=True` ignores the cache (see `invalidate`)."""
        self._sync_engine()
        transl = batch["smpl_params"]["transl"]
        key = self.batch_token(batch, transl, (num_samples,))
        if not force and self.token_matches(self._cond_key, key) and self._cond_features is features:
            return self._cond
        self._cond_features = features
        was_training = self.training
        self.eval()
        bs = batch["img"].shape[0]
        pts = batch["scene_pcd_verts_full"] - transl.unsqueeze(1) if self.scene_cano else batch["scene_pcd_verts_full"]
        if features is None:
            img_feats = (self.engine.resnet_forward(batch["img"].float().contiguous()) if self.native_image_enc
                         else self._fast_backbone(batch["img"]))
            scene_feats = (self.engine.pointnet_forward(pts.float().contiguous()) if self.native_scene_enc
                           else self._fast_scene_enc(pts))
            h = self.engine.linear(transl.float().contiguous(), *self._heads["transl0"], relu=True)
            transl_feat = self.engine.linear(h, *self._heads["transl2"])
        else:
            img_feats, scene_feats, transl_feat = features["img_feats"], features["scene_feats"], features["transl_feat"]
        # visibility mask (:186-189), [scene | transl | camera] condition vector (:195-223) and the beta head's input (:263):
        # one kernel instead of ~18 small torch launches
        c = lambda k: batch[k].float().contiguous() if k in batch else None
        flags = (self.with_focal_length, self.with_bbox_info, self.with_cam_center)
        img_feats = img_feats.float().contiguous()
        vis_u8, rest, ctx_full = self.engine.cond_inputs(
            c("orig_keypoints_2d"), scene_feats.float().contiguous(), transl_feat.float().contiguous(), img_feats,
            c("fx") if any(flags) else None, c("box_center") if flags[1] else None, c("box_size") if flags[1] else None,
            c("cam_cx") if flags[2] else None, c("cam_cy") if flags[2] else None, flags, self.openpose_to_smpl,
            self.cfg.CAM.FX_NORM_COEFF)
        vis = vis_u8.bool()
        hb = self.engine.linear(ctx_full, *self._heads["beta0"], relu=True)
        betas = self.engine.linear(hb, *self._heads["beta2"]) + self.beta_layer.init_betas  # :263-265, :673-679
        self.engine.set_cond(img_feats, rest, vis_u8)
        iob = np.repeat(np.arange(bs, dtype=np.int32), num_samples)
        if self._bodies_key != (bs, num_samples) or self.engine.n_bodies != iob.shape[0]:
            self.engine.set_bodies(iob)      # uploads the slot tables (synchronising): only when the layout changes
            self._bodies_key = (bs, num_samples)
            self._bodies_idx = torch.from_numpy(iob.astype(np.int64)).to(img_feats.device)
            self._bodies_idx32 = self._bodies_idx.to(torch.int32)
        idx = self._bodies_idx
        self._cond = {"vis": vis, "betas_img": betas.float().contiguous(), "scene_pts": pts, "transl": transl,
                      "img_of_body": idx, "img_of_body_i32": self._bodies_idx32, "num_samples": num_samples, "bs": bs}
        self.scene_pcd_verts = pts
        self.input_transl = transl
        self._cond_key = key
        if was_training:
            self.train()
        return self._cond

    @torch.no_grad()
    def set_timesteps(self, t_orig_list):
        """Upload embed_timestep(t) for the original timesteps the sampler will visit (row i <-> respaced step i)."""
        self._sync_engine()
        key = tuple(int(t) for t in t_orig_list)
        if key == self._temb_key:
            return
        t = torch.tensor(key, device=self.device, dtype=torch.long)
        temb = self.embed_timestep(t).squeeze(0) if len(key) > 1 else self.embed_timestep(t).reshape(1, -1)
        self.engine.set_temb(temb.reshape(len(key), -1).float().contiguous())
        self._temb_key = key

    # ------------------------------------------------------------------ outputs
    def assemble_outputs(self, batch, x0, cond=None):
        """egohmr.py:256-303 for the final x0: de-normalise, rot6d, SMPL, global joints, 2-D projection."""
        cond = cond or self._cond
        idx = cond["img_of_body"]
        pose6d, R, verts, joints = self.engine.decode(x0, cond["betas_img"], want_smpl=True)
        betas = cond["betas_img"][idx]
        transl = cond["transl"][idx]
        out = {"pred_x_start": x0, "pred_pose_6d": pose6d,
               "pred_smpl_params": {"global_orient": R[:, 0:1].clone(), "body_pose": R[:, 1:].clone(), "betas": betas.clone()},
               "pred_keypoints_3d": joints, "pred_vertices": verts}
        # full-frame joints and their projection (:277-301) in one kernel; focal / centre rows are what compute_loss reads
        if self.with_focal_length:
            cam = (batch["fx"].float().contiguous(), batch["cam_cx"].float().contiguous(), batch["cam_cy"].float().contiguous())
        else:
            cam = (None, None, None)
        full, kp2d, focal, center = self.engine.project_joints(joints, cond["transl"].float().contiguous(), *cam,
                                                               cond["img_of_body_i32"], self.cfg.CAM.FX_NORM_COEFF,
                                                               self.cfg.EXTRA.FOCAL_LENGTH)
        self.camera_center_full, self.focal_length = center, focal
        out["pred_keypoints_3d_full"] = full
        out["pred_keypoints_2d_full"] = kp2d
        return out

    @torch.no_grad()
    def forward(self, batch, timesteps, eval_with_uncond=True):
        """EgoHMR.forward(batch, timesteps) (egohmr.py:173-303).  `timesteps` are ORIGINAL diffusion timesteps and must
        be uniform over the batch, as they always are while sampling (gaussian_diffusion.py:495)."""
        if self.diffuse_fuse and not eval_with_uncond:
            raise NotImplementedError("eval_with_uncond=False is the training-time path (egohmr.py:465)")
        t0 = int(timesteps[0])
        if not bool((timesteps == t0).all()):
            raise NotImplementedError("non-uniform timesteps only occur in training (out of scope)")
        cond = self.prepare(batch, num_samples=1)
        self.set_timesteps([t0])
        x_t = batch["x_t"].reshape(-1, 144).float().contiguous()
        batch["vis_mask_smpl"] = cond["vis"]  # egohmr.py:189 mutates the caller's dict
        x0 = torch.empty_like(x_t)
        scratch = torch.empty_like(x_t)
        self.engine.set_schedule(0, np.array([[1, 1, 1, 0, 0, 0, 0, 0]], np.float32))
        self.engine.denoise_step(0, x_t, None, None, scratch, x0)
        self._temb_key = None  # the sampler's table must be re-uploaded after a stand-alone forward
        return self.assemble_outputs(batch, x0, cond)

    # ------------------------------------------------------------------ collision guidance / evaluation
    UPPER_BODY = [0, 3, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23]  # egohmr.py:567

    def _smpl_state(self, x):
        """x (normalised rot6d) -> vertices, joints, axis-angle full pose, exactly what guide_coll / eval_coll hand to
        the collision model (egohmr.py:528-540, 492-495)."""
        cond = self._cond
        _, R, verts, joints = self.engine.decode(x, cond["betas_img"], want_smpl=True)
        aa = self.engine.rotmat_to_angle_axis(R.reshape(-1, 3, 3)).reshape(x.shape[0], -1)
        return verts, joints, aa

    def _crop_scene(self, verts, idx32):
        """Scene points inside each body's bounding box (egohmr.py:550-554, 504-508), all bodies in ONE launch:
        -> (mask bool [B,N], count int32 [B]).  The reference does two reductions, a compare and an `.any()` host sync
        per body."""
        pts = self.scene_pcd_verts
        if pts.dtype != torch.float32 or not pts.is_contiguous():
            pts = pts.float().contiguous()
        return self.engine.scene_crop(verts, pts, idx32)

    def _body_image_index(self, B):
        cond = self._cond
        if cond is not None and len(cond["img_of_body"]) == B:
            return cond["img_of_body"], cond["img_of_body_i32"]
        ar = torch.arange(B, device=self.device)
        return ar, ar.to(torch.int32)

    def guide_coll(self, batch, output, t, compute_grad="x_t"):
        """EgoHMR.guide_coll (egohmr.py:517-570): gradient of -mean(collision loss) w.r.t. x_t, leg joints only.
        Forward and backward through de-normalise / rot6d / SMPL / axis-angle run in the CUDA library, the per-body scene
        crop is one kernel; only the pluggable collision term itself is differentiated by autograd, w.r.t. the three
        tensors it receives.  A collision model that offers `collision_loss_batched(points[B,N,3], mask[B,N],
        SMPLOutput) -> [B]` is called once for all bodies; otherwise COAP's per-body `collision_loss(points[1,n,3],
        SMPLOutput, ret_collision_mask=None)` is called for the bodies that keep at least one point (:545-559)."""
        if self.collision_model is None:
            raise RuntimeError("guide_coll needs a collision model with COAP's collision_loss() (pass collision_model=...)")
        x_t = batch["x_t"] if compute_grad == "x_t" else output["pred_x_start"]
        x = x_t.detach().float().contiguous()
        B = x.shape[0]
        cond = self._cond
        if cond is None or self.engine.n_bodies != B:
            cond = self.prepare(batch, num_samples=1)
        verts, joints, aa = self._smpl_state(x)
        idx, idx32 = self._body_image_index(B)
        mask, count = self._crop_scene(verts, idx32)
        v = verts.detach().requires_grad_()
        j = joints.detach().requires_grad_()
        fp = aa.detach().requires_grad_()
        cm = self.collision_model
        with torch.enable_grad():
            if hasattr(cm, "collision_loss_batched"):
                losses = cm.collision_loss_batched(self.scene_pcd_verts[idx], mask,
                                                   smpl_mod.SMPLOutput(vertices=v, joints=j, full_pose=fp))
            else:
                counts = count.tolist()            # one host sync for the whole batch
                idx_l = idx.tolist()
                losses = torch.zeros(B, device=x.device)
                for i in range(B):  # COAP's interface is one body per call (:545-559)
                    if counts[i]:
                        so = smpl_mod.SMPLOutput(vertices=v[[i]], joints=j[[i]], full_pose=fp[[i]])
                        losses[i] = cm.collision_loss(self.scene_pcd_verts[idx_l[i]][mask[i]].unsqueeze(0), so,
                                                      ret_collision_mask=None)
            # No body collides -> zero gradient (:561-570).  The per-body path knows that on the host already; the batched
            # path needs no read-back: bodies with an empty crop have a masked, constant-zero loss whose gradient is zero.
            if not losses.requires_grad or (not hasattr(cm, "collision_loss_batched") and not any(counts)):
                return torch.zeros(B, 144, device=x.device)
            # The reference averages over the bodies of ONE val_losses call (:561), i.e. over the images of the batch.
            # When the samples of an image are flattened into this batch (sample_many) the gradient of a body must not
            # shrink with the number of samples: divide by the images, not by images x samples.
            n_call = cond["bs"] if cond["bs"] * cond["num_samples"] == B else B
            gv, gj, ga = torch.autograd.grad([-(losses.sum() / n_call)], [v, j, fp], allow_unused=True)
        c = lambda g: None if g is None else g.float().contiguous()
        grad = self.engine.smpl_backward(x, cond["betas_img"], c(gv), c(gj), c(ga)).reshape(-1, 24, 6)
        grad[:, 3:] = grad[:, 3:] * 2            # :564-565
        grad[:, self.UPPER_BODY] = 0             # :567
        return grad.reshape(-1, 144)

    @torch.no_grad()
    def eval_coll(self, output):
        """EgoHMR.eval_coll (egohmr.py:487-514): fraction of scene points the collision model marks as inside.
        `query_batched(points[B,N,3], mask[B,N], SMPLOutput) -> occupancy[B,N]` is used when the model offers it."""
        self.poll_overflow()
        if self.collision_model is None:
            raise RuntimeError("eval_coll needs a collision model with COAP's query() (pass collision_model=...)")
        p = output["pred_smpl_params"]
        B = p["body_pose"].shape[0]
        R = torch.cat([p["global_orient"].reshape(B, 1, 3, 3), p["body_pose"].reshape(B, 23, 3, 3)], dim=1).float().contiguous()
        verts, joints = self.engine.smpl_forward(R, p["betas"].float().contiguous())
        aa = self.engine.rotmat_to_angle_axis(R.reshape(-1, 3, 3)).reshape(B, -1)
        idx, idx32 = self._body_image_index(B)
        mask, count = self._crop_scene(verts, idx32)
        n_pts = self.scene_pcd_verts.shape[1]
        cm = self.collision_model
        if hasattr(cm, "query_batched"):
            occ = cm.query_batched(self.scene_pcd_verts[idx], mask, smpl_mod.SMPLOutput(vertices=verts, joints=joints, full_pose=aa))
            return (((occ > 0.5) & mask).sum(dim=1) / n_pts).tolist()
        counts, idx_l = count.tolist(), idx.tolist()
        ratios = []
        for i in range(B):
            if counts[i]:
                so = smpl_mod.SMPLOutput(vertices=verts[[i]].clone(), joints=joints[[i]].clone(), full_pose=aa[[i]].clone())
                occ = cm.query(self.scene_pcd_verts[idx_l[i]][mask[i]].unsqueeze(0), so)
                ratios.append(((occ > 0.5).sum() / n_pts).item())
            else:
                ratios.append(0.0)
        return ratios

    SMPL_TO_OPENPOSE = [24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34]  # :108

    @torch.no_grad()
    def compute_loss(self, batch, output, cur_epoch=0):
        """EgoHMR.compute_loss in validation mode (egohmr.py:305-445) — what `val_losses` runs by default
        (gaussian_diffusion.py:777-778; test_egohmr.py:252-255 does not pass compute_loss).  Adds `output['losses']` and
        `output['joint_vis_num_batch']`, returns the weighted total.  Small host-side tensor plumbing on the final
        outputs; the ground-truth bodies go through the CUDA SMPL (pose2rot=True)."""
        self.poll_overflow()
        if self.training:
            raise NotImplementedError("training-mode compute_loss (autograd through the denoiser) is out of scope")
        for k in ("keypoints_3d", "keypoints_3d_full", "smpl_params_is_axis_angle", "gender"):
            if k not in batch:
                raise KeyError(f"compute_loss needs batch['{k}'] (ground truth, egobody_dataset.py:241-277); pass "
                               f"compute_loss=False to val_losses for label-free sampling")
        p = output["pred_smpl_params"]
        kp3d = output["pred_keypoints_3d"][:, 0:24]
        kp3d_full = output["pred_keypoints_3d_full"][:, 0:24]
        kp2d = output["pred_keypoints_2d_full"][:, self.SMPL_TO_OPENPOSE, :]
        bs = p["body_pose"].shape[0]
        gt2d, gt3d, gt3d_full = batch["orig_keypoints_2d"], batch["keypoints_3d"], batch["keypoints_3d_full"]
        gp, is_aa, gender = batch["smpl_params"], batch["smpl_params_is_axis_angle"], batch["gender"]
        conf = gt2d[:, :, -1].unsqueeze(-1).clone()                               # losses.py:20-24
        conf[:, [1, 9, 12], :] = 0
        l_2d = (conf * (kp2d - gt2d[:, :, :-1]).abs()).sum(dim=(1, 2)).mean()
        l_3d = ((kp3d - kp3d[:, [0]]) - (gt3d - gt3d[:, [0]])).abs().sum(dim=(1, 2)).mean()      # pelvis_align=True
        l_3d_full = (kp3d_full - gt3d_full).abs().sum(dim=(1, 2)).mean()
        gt_args = {k: v.float() for k, v in gp.items()}
        gm, gf = self.smpl_male(**gt_args), self.smpl_female(**gt_args)            # :344-351
        gt_v, gt_j = gm.vertices, gm.joints
        fem = gender == 1
        gt_v[fem], gt_j[fem] = gf.vertices[fem], gf.joints[fem]
        l_v2v = ((output["pred_vertices"] - kp3d[:, [0]]) - (gt_v - gt_j[:, [0]])).abs().mean()
        g2d = perspective_projection(gt_j, torch.zeros(bs, 3, device=gt_j.device), self.focal_length,
                                     self.camera_center_full)[:, :24]              # :361-366
        vis = (g2d[:, :, 0] >= 0) * (g2d[:, :, 0] < 1920) * (g2d[:, :, 1] >= 0) * (g2d[:, :, 1] < 1080)
        l_vis = (torch.sqrt((((kp3d - kp3d[:, [0]]) - (gt3d[:, 0:24] - gt3d[:, [0]])) ** 2).sum(dim=-1)) * vis).sum()
        from ...utils.geometry import aa_to_rotmat
        lp = {}
        for k, pred in p.items():
            if k != "transl":
                gt = gp[k]
                if bool(torch.as_tensor(is_aa[k]).all()):
                    gt = aa_to_rotmat(gt.reshape(-1, 3)).view(bs, -1, 3, 3)
                lp[k] = ((pred - gt) ** 2).sum() / bs
        p6 = output["pred_pose_6d"].reshape(-1, 3, 2)
        l_ortho = ((torch.matmul(p6.permute(0, 2, 1), p6) - torch.eye(2, device=p6.device, dtype=p6.dtype).unsqueeze(0)) ** 2).mean()
        l_coap = torch.tensor(0.0, device=p6.device)
        if self.weight_coap_penetration > 0 and cur_epoch >= self.start_coap_epoch:
            raise NotImplementedError("the COAP penetration term of compute_loss is a training loss (its weight is 0 in "
                                      "test_egohmr.py:112-118)")
        w = self.loss_weights
        loss = (w["v2v"] * l_v2v + w["keypoints_3d"] * l_3d + w["keypoints_3d_full"] * l_3d_full
                + w["keypoints_2d_full"] * l_2d + w["betas"] * lp["betas"] + w["body_pose"] * lp["body_pose"]
                + w["global_orient"] * lp["global_orient"] + w["pose_6d_ortho"] * l_ortho
                + self.weight_coap_penetration * l_coap)
        output["losses"] = dict(loss=loss.detach(), loss_v2v=l_v2v.detach(), loss_keypoints_3d=l_3d.detach(),
                                loss_keypoints_3d_full=l_3d_full.detach(), loss_keypoints_2d_full=l_2d.detach(),
                                loss_betas=lp["betas"].detach(), loss_body_pose=lp["body_pose"].detach(),
                                loss_global_orient=lp["global_orient"].detach(), loss_pose_6d_ortho=l_ortho.detach(),
                                loss_coap_penetration=l_coap.detach(), loss_keypoints_3d_vis_batch_sum=l_vis.detach())
        output["joint_vis_num_batch"] = vis.sum()
        return loss

"""Run the UNMODIFIED reference (sanweiliti/EgoHMR) as a library: golden-vector generation and `bench.py`'s reference arms.

Where the reference tree comes from (first that exists): ``$EHB_REFERENCE_ROOT``; ``baseline/_ref`` (a verbatim copy made
by ``tools/install_ref.sh`` / ``__graft_entry__.build()`` — git-ignored like the built ``.so`` files, so no reference
source enters the history, but it travels to the GPU box with the snapshot); ``/root/reference`` (build container only).
Nothing in the product package (``egohmr_b200/``) imports this module: users are ``tests/golden/make_golden.py`` (fixture
generation) and ``bench.py --impl reference`` / ``--impl reference-cuda`` (the timed baselines).

The reference imports three packages that are not installed and cannot be fetched offline: ``smplx``
(0.1.28), ``coap`` (unpinned git), ``yacs``.  They are replaced by minimal ``sys.modules`` stand-ins so that
``models/egohmr/egohmr.py``, ``diffusion/*`` and ``utils/*`` run exactly as they are:

* ``smplx.create`` -> a torch restatement of ``smplx/lbs.py::lbs`` + ``SMPL.forward`` over a synthetic SMPL model
  (differentiable, because ``guide_coll`` back-propagates through it; plain torch ops, so it runs on CPU and CUDA like
  smplx itself);
* ``coap.attach_coap`` -> attaches an object exposing ``collision_loss`` / ``query`` with COAP's call signature, backed by
  a synthetic analytic penalty (see ``SyntheticCollision``);
* ``torch.utils.model_zoo.load_url`` -> ``{}`` (``models/resnet.py:211`` would download ImageNet weights);
* ``data/smpl_mean_params.npz`` is created in a scratch working directory (``egohmr.py:669`` reads it from CWD).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    """-> (path, kind) of the reference tree to import, or (None, None)."""
    env = os.environ.get("EHB_REFERENCE_ROOT")
    if os.environ.get("EHB_IGNORE_REFERENCE"):     # tests of the no-reference fallback
        return None, None
    for path, kind in ((env, "env"), (os.path.join(_HERE, "_ref"), "baseline/_ref"), ("/root/reference", "/root/reference")):
        if path and os.path.isfile(os.path.join(path, "diffusion", "gaussian_diffusion.py")):
            return path, kind
    return None, None


REFERENCE_ROOT = reference_root()[0]


class SMPLOutput:
    """smplx.utils.SMPLOutput stand-in: default-constructible, settable fields (egohmr.py:393-396,491)."""

    def __init__(self, vertices=None, joints=None, full_pose=None, betas=None, global_orient=None, body_pose=None):
        self.vertices = vertices
        self.joints = joints
        self.full_pose = full_pose
        self.betas = betas
        self.global_orient = global_orient
        self.body_pose = body_pose


def batch_rodrigues(rot_vecs):
    """smplx/lbs.py::batch_rodrigues restated from the published algorithm (angle = |v + 1e-8|, R = I + sin K + (1-cos) K^2)."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos, sin = torch.cos(angle).unsqueeze(1), torch.sin(angle).unsqueeze(1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros_like(rx)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(-1, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device).unsqueeze(0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


class TorchSMPL(nn.Module):
    """Torch restatement of smplx.SMPL.forward(pose2rot=False) / lbs() used as the `smplx.create` stand-in."""

    def __init__(self, model):
        super().__init__()
        f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
        self.register_buffer("v_template", f(model["v_template"]))
        self.register_buffer("shapedirs", f(model["shapedirs"]))
        self.register_buffer("posedirs", f(model["posedirs"]))
        self.register_buffer("J_regressor", f(model["J_regressor"]))
        self.register_buffer("lbs_weights", f(model["lbs_weights"]))
        self.parents = [int(p) for p in model["parents"]]
        self.extra = [int(v) for v in model["extra_vertex_ids"]]
        self.faces = np.zeros((1, 3), dtype=np.int64)
        # The reference casts SMPL inputs with .float() (egohmr.py:276), so in an fp64 run the SMPL outputs come back
        # as fp32 and its own perspective_projection then fails on mixed dtypes; `out_dtype` lets the fp64 golden run
        # get past that (the denoiser trace is what the fp64 run is for).
        self.out_dtype = None

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, return_full_pose=False,
                pose2rot=True, **kwargs):
        B = betas.shape[0]
        if pose2rot:   # compute_loss evaluates the ground-truth body from axis-angle parameters (egohmr.py:344-347)
            aa = torch.cat([global_orient.reshape(B, -1, 3), body_pose.reshape(B, -1, 3)], dim=1)
            full_pose = batch_rodrigues(aa.reshape(-1, 3)).reshape(B, 24, 3, 3)
        else:
            full_pose = torch.cat([global_orient.reshape(B, -1, 3, 3), body_pose.reshape(B, -1, 3, 3)], dim=1)
        dt = full_pose.dtype
        v_shaped = self.v_template.to(dt) + torch.einsum("bl,mkl->bmk", betas, self.shapedirs.to(dt))
        J = torch.einsum("bik,ji->bjk", v_shaped, self.J_regressor.to(dt))
        ident = torch.eye(3, dtype=dt, device=full_pose.device)
        pose_feature = (full_pose[:, 1:] - ident).reshape(B, -1)
        v_posed = torch.matmul(pose_feature, self.posedirs.to(dt)).view(B, -1, 3) + v_shaped
        rel = J.clone()
        rel[:, 1:] = J[:, 1:] - J[:, self.parents[1:]]
        T = torch.zeros(B, 24, 4, 4, dtype=dt, device=full_pose.device)
        T[:, :, :3, :3] = full_pose
        T[:, :, :3, 3] = rel
        T[:, :, 3, 3] = 1
        chain = [T[:, 0]]
        for i in range(1, 24):
            chain.append(torch.matmul(chain[self.parents[i]], T[:, i]))
        G = torch.stack(chain, dim=1)
        posed_joints = G[:, :, :3, 3]
        Jh = torch.cat([J, torch.zeros(B, 24, 1, dtype=dt, device=J.device)], dim=2).unsqueeze(-1)
        A = G - torch.nn.functional.pad(torch.matmul(G, Jh), [3, 0, 0, 0, 0, 0, 0, 0])
        Tv = torch.matmul(self.lbs_weights.to(dt).unsqueeze(0).expand(B, -1, -1), A.view(B, 24, 16)).view(B, -1, 4, 4)
        vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=dt, device=J.device)], dim=2)
        verts = torch.matmul(Tv, vh.unsqueeze(-1))[:, :, :3, 0]
        joints = torch.cat([posed_joints, verts[:, self.extra]], dim=1)
        if transl is not None:
            joints = joints + transl.unsqueeze(1)
            verts = verts + transl.unsqueeze(1)
        if self.out_dtype is not None:
            joints, verts = joints.to(self.out_dtype), verts.to(self.out_dtype)
        return SMPLOutput(vertices=verts, joints=joints, full_pose=full_pose if return_full_pose else None, betas=betas,
                          global_orient=global_orient, body_pose=body_pose)


from egohmr_b200.testing import SyntheticCollision  # noqa: E402  (the one collision stand-in, shared by both sides)


def make_cfg():
    """The only config fields the hot path reads."""
    ns = types.SimpleNamespace
    return ns(MODEL=ns(BACKBONE=ns(NUM_LAYERS=50, OUT_CHANNELS=2048)), CAM=ns(FX_NORM_COEFF=1500.0),
              EXTRA=ns(FOCAL_LENGTH=5000.0), TRAIN=ns(LR=1e-4, WEIGHT_DECAY=1e-4))


_installed = {}


def install(smpl_model, init_betas):
    """Install the stand-ins, chdir to a scratch dir holding data/smpl_mean_params.npz, put the reference on sys.path."""
    if _installed:
        return _installed["workdir"]
    smplx = types.ModuleType("smplx")
    smplx_utils = types.ModuleType("smplx.utils")
    smplx_utils.SMPLOutput = SMPLOutput
    smplx.utils = smplx_utils
    smplx.create = lambda *a, **k: TorchSMPL(smpl_model)
    sys.modules["smplx"] = smplx
    sys.modules["smplx.utils"] = smplx_utils
    coap = types.ModuleType("coap")

    def attach_coap(smpl, pretrained=True, device=None):
        object.__setattr__(smpl, "coap", SyntheticCollision())
        return smpl

    coap.attach_coap = attach_coap
    sys.modules["coap"] = coap
    import torch.utils.model_zoo as model_zoo
    model_zoo.load_url = lambda *a, **k: {}
    workdir = tempfile.mkdtemp(prefix="ehb_ref_")
    os.makedirs(os.path.join(workdir, "data"), exist_ok=True)
    np.savez(os.path.join(workdir, "data", "smpl_mean_params.npz"), shape=np.asarray(init_betas, dtype=np.float32))
    os.chdir(workdir)
    if REFERENCE_ROOT is None:
        raise RuntimeError("no reference tree: run tools/install_ref.sh in the build container (copies /root/reference "
                           "to baseline/_ref) or set EHB_REFERENCE_ROOT")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed["workdir"] = workdir
    return workdir


# ------------------------------------------------------------------------------------------------ driver restatement
def to_torch(batch, dtype=torch.float32, device="cpu"):
    """numpy batch dict (egohmr_b200.synth.make_batch, dataloaders/egobody_dataset.py:241-277 schema) -> torch."""
    out = {}
    for k, v in batch.items():
        if isinstance(v, dict):
            out[k] = to_torch(v, dtype, device)
        else:
            t = torch.from_numpy(np.asarray(v))
            out[k] = (t.to(dtype) if t.is_floating_point() else t).to(device)
    return out


def build_reference(hid=1024, n_blocks=4, dtype=torch.float32, seed=0, only_mask_img_cond=True, diffuse_fuse=True,
                    nonlocal_layer=False, device="cpu"):
    """The reference's EgoHMR exactly as test_egohmr.py:112-127 builds it (test-default flags, :53-78), holding the
    seeded synthetic weights of `egohmr_b200.synth` instead of a checkpoint.  -> (model, Xmean, Xstd)"""
    from egohmr_b200 import synth
    smpl_model = synth.make_smpl_model(seed)
    install(smpl_model, smpl_model["init_betas"])
    from models.egohmr.egohmr import EgoHMR
    mean, std = synth.body_rep_stats(seed)
    dev = torch.device(device)
    model = EgoHMR(cfg=make_cfg(), device=dev,
                   body_rep_mean=torch.from_numpy(mean).to(dtype).to(dev), body_rep_std=torch.from_numpy(std).to(dtype).to(dev),
                   with_focal_length=True, with_bbox_info=True, with_cam_center=True, scene_feat_dim=512,
                   scene_type="cube", scene_cano=True, cond_mask_prob=0.0, only_mask_img_cond=only_mask_img_cond,
                   pelvis_vis_loosen=True, diffuse_fuse=diffuse_fuse, diffusion_blk=n_blocks, gcn_hid_dim=hid,
                   gcn_nonlocal_layer=nonlocal_layer)
    sd = synth.make_state_dict(seed, hid=hid, n_blocks=n_blocks, init_betas=smpl_model["init_betas"])
    if nonlocal_layer:
        synth.add_nonlocal(sd, seed, hid)
    res = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys and all(k.startswith("smpl") for k in res.missing_keys), res
    model = model.to(dtype)
    # the reference keeps `adj` as a plain attribute (not a buffer): .to() does not move or cast it
    for m in model.modules():
        if hasattr(m, "adj") and isinstance(getattr(m, "adj"), torch.Tensor):
            m.adj = m.adj.to(dtype).to(dev)
    model.eval()
    if dtype == torch.float64:
        model.smpl.out_dtype = torch.float64
    return model, mean, std


def build_sampler(T, respacing, mean, std, dtype=torch.float32, device="cpu"):
    """create_gaussian_diffusion as test_egohmr.py:120-123 calls it."""
    from diffusion.model_util import create_gaussian_diffusion
    return create_gaussian_diffusion(num_diffusion_timesteps=T, timestep_respacing=respacing,
                                     body_rep_mean=torch.from_numpy(mean).to(dtype).to(device),
                                     body_rep_std=torch.from_numpy(std).to(dtype).to(device))


def make_driver_batch(seed, n_img, n_pts=1024, device="cpu", with_gt=True):
    """Synthetic batch in the dataloader's schema; with_gt adds the ground-truth keys `compute_loss` reads, because the
    driver's val_losses call leaves compute_loss at its default True (test_egohmr.py:252-255)."""
    from egohmr_b200 import synth
    b = synth.make_batch(seed, n_img, n_pts)
    if with_gt:
        b = synth.merge_gt(b, synth.make_gt(seed, n_img))
    return to_torch(b, torch.float32, device)


def driver_loop(model, diffusion_sample, batch, num_samples, respacing, with_coap_grad=False, cond_grad_weight=1.0,
                compute_loss=None, eval_coll=False):
    """test_egohmr.py:247-266 for one dataloader batch, verbatim in structure: `num_samples` sequential val_losses calls,
    the per-sample pred_smpl_params stacked to [bs, n_sample, ...].  `compute_loss=None` leaves val_losses' default
    (True), as the driver does."""
    curr_batch_size = batch["img"].shape[0]
    kw = {} if compute_loss is None else {"compute_loss": compute_loss}
    with torch.no_grad():
        shape = [curr_batch_size, 144]
        out_all = {"pred_smpl_params": {}}
        for _n in range(num_samples):
            out_cur = diffusion_sample.val_losses(model=model, batch=batch, shape=shape, progress=False,
                                                  clip_denoised=False, cur_epoch=0, timestep_respacing=respacing,
                                                  cond_fn_with_grad=with_coap_grad, cond_grad_weight=cond_grad_weight, **kw)
            if eval_coll:
                model.eval_coll(out_cur)
            for key, val in out_cur["pred_smpl_params"].items():
                out_all["pred_smpl_params"].setdefault(key, []).append(val.unsqueeze(1))
        for key in out_all["pred_smpl_params"]:
            out_all["pred_smpl_params"][key] = torch.cat(out_all["pred_smpl_params"][key], dim=1)
    return out_all


def flat_batch(batch, num_samples):
    """BASELINE.md 3 "B-gpu-flat": the reference's own code on a batch whose images are tiled `num_samples` times, i.e.
    one val_losses call of bs*num_samples bodies (the fairest batched use of the unmodified reference)."""
    rep = lambda t: t.repeat_interleave(num_samples, dim=0) if isinstance(t, torch.Tensor) and t.dim() > 0 else t
    return {k: ({kk: rep(vv) for kk, vv in v.items()} if isinstance(v, dict) else rep(v)) for k, v in batch.items()}

"""Import the UNMODIFIED reference (/root/reference) in the build container.

Only used by ``tests/golden/make_golden.py`` (fixture generation) — never at test/bench run time, because
``/root/reference`` does not exist on the GPU box.  The reference imports three packages that are not installed and
cannot be fetched offline: ``smplx`` (0.1.28), ``coap`` (unpinned git), ``yacs``.  They are replaced by
minimal ``sys.modules`` stand-ins so that ``models/egohmr/egohmr.py``, ``diffusion/*`` and ``utils/*`` run as they are:

* ``smplx.create`` -> a torch restatement of ``smplx/lbs.py::lbs`` + ``SMPL.forward`` over a synthetic SMPL model
  (differentiable, because ``guide_coll`` back-propagates through it);
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

REFERENCE_ROOT = "/root/reference"


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
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed["workdir"] = workdir
    return workdir

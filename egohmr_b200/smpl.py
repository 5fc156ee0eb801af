"""`SMPL()` with the call signature the reference uses (smplx.create / SMPL.forward), backed by the CUDA LBS kernels.

Reference call sites: models/egohmr/egohmr.py:105-107,276,492,537; test_egohmr.py:143-145,291-292,307-312.
smplx itself is not vendored in the reference and not installable offline, so the model tensors come
either from an SMPL `.pkl`/`.npz` under `model_path` (same file layout smplx reads) or from a dict of arrays.
"""
import os
import pickle

import numpy as np
import torch
import torch.nn as nn

from . import synth
from .engine import Engine


class SMPLOutput:
    """smplx.utils.SMPLOutput: default-constructible, fields settable (egohmr.py:393-396, 491-501)."""

    def __init__(self, vertices=None, joints=None, full_pose=None, betas=None, global_orient=None, body_pose=None,
                 transl=None):
        self.vertices = vertices
        self.joints = joints
        self.full_pose = full_pose
        self.betas = betas
        self.global_orient = global_orient
        self.body_pose = body_pose
        self.transl = transl


def _load_model_file(model_path, gender):
    cands = []
    if os.path.isdir(model_path):
        for sub in ("", "smpl"):
            for ext in ("pkl", "npz"):
                cands.append(os.path.join(model_path, sub, f"SMPL_{gender.upper()}.{ext}"))
    else:
        cands.append(model_path)
    for p in cands:
        if os.path.exists(p):
            if p.endswith(".npz"):
                d = dict(np.load(p, allow_pickle=True))
            else:
                with open(p, "rb") as f:
                    d = pickle.load(f, encoding="latin1")
            V = np.asarray(d["v_template"]).shape[0]
            posedirs = np.asarray(d["posedirs"], dtype=np.float32)
            if posedirs.ndim == 3:  # [V,3,207] on disk -> [207, V*3] as smplx registers it
                posedirs = posedirs.reshape(-1, posedirs.shape[-1]).T
            J_reg = d["J_regressor"]
            J_reg = np.asarray(J_reg.todense() if hasattr(J_reg, "todense") else J_reg, dtype=np.float32)
            parents = np.asarray(d["kintree_table"])[0].astype(np.int64).copy()
            parents[0] = -1
            return {
                "v_template": np.asarray(d["v_template"], np.float32),
                "shapedirs": np.asarray(d["shapedirs"], np.float32)[:, :, :10],
                "posedirs": posedirs, "J_regressor": J_reg,
                "lbs_weights": np.asarray(d["weights"], np.float32), "parents": parents.astype(np.int32),
                "extra_vertex_ids": np.array([v % V for v in synth.SMPL_EXTRA_VERTEX_IDS], np.int32),
                "faces": np.asarray(d["f"], np.int64) if "f" in d else np.zeros((0, 3), np.int64),
            }
    raise FileNotFoundError(f"no SMPL model file found under {model_path!r} (tried {cands})")


class SMPL(nn.Module):
    def __init__(self, model, engine=None, device=None, create_transl=True, batch_size=1):
        super().__init__()
        self.model = model
        self.faces = model.get("faces", np.zeros((0, 3), np.int64))
        self._engine = engine
        self.batch_size = batch_size
        self._dev = torch.device(device) if device is not None else None
        # smplx registers a zero `transl` parameter by default, which makes forward add zeros (a no-op)
        self.create_transl = create_transl
        self.register_buffer("_anchor", torch.zeros(1))

    def _eng(self, device):
        if self._engine is None:
            self._engine = Engine(device.index or 0)
        if not self._engine.smpl_loaded:
            self._engine.load_smpl(self.model)
        return self._engine

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, return_verts=True,
                return_full_pose=False, pose2rot=True, **kwargs):
        dev = betas.device
        n = betas.shape[0]
        eng = self._eng(dev)
        if pose2rot:
            from .utils.geometry import batch_rodrigues
            full = torch.cat([global_orient.reshape(n, -1, 3), body_pose.reshape(n, -1, 3)], dim=1)
            R = batch_rodrigues(full.reshape(-1, 3).float()).reshape(n, 24, 3, 3)
        else:
            R = torch.cat([global_orient.reshape(n, -1, 3, 3), body_pose.reshape(n, -1, 3, 3)], dim=1)
        R = R.float().contiguous()
        verts, joints = eng.smpl_forward(R, betas.float().contiguous(),
                                         None if transl is None else transl.float().contiguous())
        return SMPLOutput(vertices=verts, joints=joints, full_pose=R if return_full_pose else None, betas=betas,
                          global_orient=global_orient, body_pose=body_pose, transl=transl)


def create(model_path="data/smpl", model_type="smpl", gender="neutral", smpl_model=None, engine=None, **kwargs):
    """smplx.create(...) (egohmr.py:105-107).  `smpl_model` (dict of arrays) bypasses the file lookup."""
    if model_type != "smpl":
        raise ValueError("only model_type='smpl' is on the EgoHMR path")
    if smpl_model is None:
        smpl_model = _load_model_file(model_path, gender)
    return SMPL(smpl_model, engine=engine, create_transl=kwargs.get("create_transl", True),
                batch_size=kwargs.get("batch_size", 1))

"""SMPL forward pass (numpy).  TEST INFRASTRUCTURE (see oracle/__init__).

PARITY UNPINNED: the arithmetic lives in smplx==0.1.28 (reference environment.yml:197), which is neither vendored in
/root/reference nor installable offline, and the reference holds no vectors for it.  This restates the published
algorithm of smplx/lbs.py::lbs (blend_shapes, vertices2joints, batch_rigid_transform) and
smplx/body_models.py::SMPL.forward with pose2rot=False, anchored on the reference's call sites
(models/egohmr/egohmr.py:105-107,276,492,537; test_egohmr.py:143-145,291-292).  Explicit assumptions restated from
upstream memory: the 21 extra joints are vertices picked by VertexJointSelector in the order of
egohmr_b200.synth.SMPL_EXTRA_VERTEX_IDS; the module owns a zero `transl` parameter, so the translation add is a no-op
unless `transl` is passed.  Known-answer tests we author (tests/test_oracle_smpl.py): identity pose + zero betas gives
the template; a root-only rotation is a rigid rotation about joint 0.
"""
import numpy as np


def smpl_forward(model, R, betas, transl=None):
    """R [B,24,3,3] (global_orient first), betas [B,nb] -> dict(vertices [B,V,3], joints [B,24+n_extra,3], A, J)."""
    dt = R.dtype
    B = R.shape[0]
    v_t = model["v_template"].astype(dt)
    S = model["shapedirs"].astype(dt)
    P = model["posedirs"].astype(dt)
    Jr = model["J_regressor"].astype(dt)
    W = model["lbs_weights"].astype(dt)
    parents = [int(p) for p in model["parents"]]
    v_shaped = v_t[None] + np.einsum("bl,mkl->bmk", betas.astype(dt), S)          # blend_shapes
    J = np.einsum("bik,ji->bjk", v_shaped, Jr)                                     # vertices2joints
    pose_feature = (R[:, 1:] - np.eye(3, dtype=dt)).reshape(B, -1)                 # [B,207]
    v_posed = np.matmul(pose_feature, P).reshape(B, -1, 3) + v_shaped
    rel = J.copy()
    rel[:, 1:] = J[:, 1:] - J[:, parents[1:]]
    T = np.zeros((B, 24, 4, 4), dtype=dt)
    T[:, :, :3, :3] = R
    T[:, :, :3, 3] = rel
    T[:, :, 3, 3] = 1
    chain = [T[:, 0]]
    for i in range(1, 24):
        chain.append(np.matmul(chain[parents[i]], T[:, i]))
    G = np.stack(chain, axis=1)
    posed_joints = G[:, :, :3, 3]
    A = G.copy()
    A[:, :, :3, 3] = G[:, :, :3, 3] - np.einsum("bjrc,bjc->bjr", G[:, :, :3, :3], J)
    Tv = np.einsum("vj,bjrc->bvrc", W, A)
    verts = np.einsum("bvrc,bvc->bvr", Tv[:, :, :3, :3], v_posed) + Tv[:, :, :3, 3]
    joints = np.concatenate([posed_joints, verts[:, [int(v) for v in model["extra_vertex_ids"]]]], axis=1)
    if transl is not None:
        joints = joints + transl[:, None, :].astype(dt)
        verts = verts + transl[:, None, :].astype(dt)
    return {"vertices": verts, "joints": joints, "A": A[:, :, :3, :], "J": J, "pose_feature": pose_feature}

"""EgoHMR.compute_loss in validation mode (numpy).  TEST INFRASTRUCTURE (see oracle/__init__).

`val_losses(...)` calls it by default (diffusion/gaussian_diffusion.py:777-778, compute_loss=True is what
test_egohmr.py:252-255 gets), so it is part of the drop-in surface even though its result only lands in
output['losses'].  Restates models/egohmr/egohmr.py:305-445 and models/egohmr/losses.py:4-50,80-93; the ground-truth
body goes through SMPL with pose2rot=True, i.e. smplx/lbs.py::batch_rodrigues (parity unpinned, like oracle/smpl.py).
"""
import numpy as np

from . import geometry, smpl as o_smpl

SMPL_TO_OPENPOSE = [24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34]  # egohmr.py:108


def batch_rodrigues(rot_vecs):
    """smplx/lbs.py::batch_rodrigues: angle = |v + 1e-8|, R = I + sin(angle) K + (1 - cos(angle)) K K."""
    dt = rot_vecs.dtype
    angle = np.linalg.norm(rot_vecs + dt.type(1e-8), axis=1, keepdims=True)
    d = rot_vecs / angle
    K = np.zeros((rot_vecs.shape[0], 3, 3), dt)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -d[:, 2], d[:, 1], d[:, 2], -d[:, 0], -d[:, 1], d[:, 0]
    s, c = np.sin(angle)[:, :, None], np.cos(angle)[:, :, None]
    return np.eye(3, dtype=dt)[None] + s * K + (1 - c) * np.matmul(K, K)


def aa_to_rotmat(theta):
    """utils/geometry.py:5-43 (axis-angle -> quaternion -> matrix), used for the parameter losses (egohmr.py:381)."""
    dt = theta.dtype
    norm = np.linalg.norm(theta + dt.type(1e-8), axis=1)
    angle = norm[:, None]
    n = theta / angle
    angle = angle * dt.type(0.5)
    q = np.concatenate([np.cos(angle), np.sin(angle) * n], axis=1)
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return np.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz, 2 * wz + 2 * xy, w2 - x2 + y2 - z2,
                     2 * yz - 2 * wx, 2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], axis=1).reshape(-1, 3, 3)


def compute_loss(smpl_model, batch, out, weights=None, fx_norm_coeff=1500.0, dtype=np.float32):
    """-> (loss, losses dict, joint_vis_num_batch) for eval mode (self.training False, no COAP term: its weight is 0 in
    test_egohmr.py:112-118)."""
    dt = np.dtype(dtype)
    f = lambda a: np.asarray(a).astype(dt)
    B = out["pred_smpl_params"]["body_pose"].shape[0]
    kp3d = f(out["pred_keypoints_3d"])[:, :24]
    kp3d_full = f(out["pred_keypoints_3d_full"])[:, :24]
    kp2d = f(out["pred_keypoints_2d_full"])[:, SMPL_TO_OPENPOSE]
    gt2d = f(batch["orig_keypoints_2d"])
    conf = gt2d[:, :, -1:].copy()
    conf[:, [1, 9, 12]] = 0                                                        # losses.py:21-23
    l_2d = (conf * np.abs(kp2d - gt2d[:, :, :-1])).sum(axis=(1, 2)).mean()
    gt3d, gt3d_full = f(batch["keypoints_3d"]), f(batch["keypoints_3d_full"])
    l_3d = np.abs((kp3d - kp3d[:, [0]]) - (gt3d - gt3d[:, [0]])).sum(axis=(1, 2)).mean()   # pelvis_align=True
    l_3d_full = np.abs(kp3d_full - gt3d_full).sum(axis=(1, 2)).mean()
    gp = batch["smpl_params"]
    R = batch_rodrigues(np.concatenate([f(gp["global_orient"]).reshape(B, 1, 3), f(gp["body_pose"]).reshape(B, 23, 3)],
                                       axis=1).reshape(-1, 3)).reshape(B, 24, 3, 3)
    so = o_smpl.smpl_forward(smpl_model, R, f(gp["betas"]), transl=f(gp["transl"]))   # male == female model here
    gt_v, gt_j = so["vertices"], so["joints"]
    l_v2v = np.abs((f(out["pred_vertices"]) - kp3d[:, [0]]) - (gt_v - gt_j[:, [0]])).mean()
    fx = f(batch["fx"])
    focal = np.stack([fx, fx], axis=1) * dt.type(fx_norm_coeff)
    center = np.stack([f(batch["cam_cx"]), f(batch["cam_cy"])], axis=1)
    g2d = geometry.perspective_projection(gt_j, np.zeros((B, 3), dt), focal, center)[:, :24]
    vis = (g2d[:, :, 0] >= 0) & (g2d[:, :, 0] < 1920) & (g2d[:, :, 1] >= 0) & (g2d[:, :, 1] < 1080)
    l_vis = (np.sqrt((((kp3d - kp3d[:, [0]]) - (gt3d - gt3d[:, [0]])) ** 2).sum(-1)) * vis).sum()
    lp = {}
    for k, pred in out["pred_smpl_params"].items():
        gt = f(gp[k])
        if np.asarray(batch["smpl_params_is_axis_angle"][k]).all():
            gt = aa_to_rotmat(gt.reshape(-1, 3)).reshape(B, -1, 3, 3)
        lp[k] = ((f(pred) - gt) ** 2).sum() / B
    p6 = f(out["pred_pose_6d"]).reshape(-1, 3, 2)
    l_ortho = ((np.matmul(p6.transpose(0, 2, 1), p6) - np.eye(2, dtype=dt)[None]) ** 2).mean()
    w = weights or {}
    loss = (w.get("v2v", 0) * l_v2v + w.get("keypoints_3d", 0) * l_3d + w.get("keypoints_3d_full", 0) * l_3d_full
            + w.get("keypoints_2d_full", 0) * l_2d + w.get("betas", 0) * lp["betas"] + w.get("body_pose", 0) * lp["body_pose"]
            + w.get("global_orient", 0) * lp["global_orient"] + w.get("pose_6d_ortho", 0) * l_ortho)
    losses = dict(loss=loss, loss_v2v=l_v2v, loss_keypoints_3d=l_3d, loss_keypoints_3d_full=l_3d_full,
                  loss_keypoints_2d_full=l_2d, loss_betas=lp["betas"], loss_body_pose=lp["body_pose"],
                  loss_global_orient=lp["global_orient"], loss_pose_6d_ortho=l_ortho, loss_coap_penetration=0.0,
                  loss_keypoints_3d_vis_batch_sum=l_vis)
    return loss, losses, int(vis.sum())

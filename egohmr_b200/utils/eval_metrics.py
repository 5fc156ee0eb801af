"""The metric block of the reference's evaluation driver (test_egohmr.py:373-494) on the device.

The reference computes G-MPJPE / MPJPE / PA-MPJPE / V2V (each with visible / invisible-joint splits), the per-joint
standard deviation and the average pairwise distance (APD) of the samples with ~60 small torch launches, `.cpu().numpy()`
copies and Python loops over the images of every batch.  Here one call = four kernels (`ehb_eval_metrics`) + the batched
Procrustes kernel for PA-MPJPE (`ehb_procrustes_align`); results are device tensors named like the driver's accumulators."""
import torch

from .geometry import _engine_for


def evaluate_batch(pred_keypoints_3d, pred_vertices, transl, gt_keypoints_3d, gt_vertices, focal_length, cam_cx, cam_cy,
                   eval_with_vis_mask_pa=False, engine=None):
    """pred_keypoints_3d [bs,S,24,3] and pred_vertices [bs,S,V,3]: SMPL outputs of the S samples (test_egohmr.py:291-297, not
    pelvis-aligned, no translation); transl [bs,3] = batch['smpl_params']['transl']; gt_keypoints_3d [bs,24,3],
    gt_vertices [bs,V,3] in the camera frame (:306-318); focal_length, cam_cx, cam_cy [bs] in pixels (:239-241).
    -> dict of device tensors: *_mpjpe / v2v [bs,S] (means), *_vis / *_invis [bs,S] (sums over the visible / invisible
    joints or vertices, as the driver accumulates them), std_joints* / apd_joints* [bs], joint_vis_mask [bs,24],
    vertex_vis_mask [bs,V], joint_vis_num, vertex_vis_num (ints, :390-396)."""
    c = lambda t: t.float().contiguous()
    pj, pv = c(pred_keypoints_3d), c(pred_vertices)
    eng = engine or _engine_for(pj.device)
    bs, S, J = pj.shape[:3]
    jv, vv, err, div = eng.eval_metrics(pj, pv, c(transl), c(gt_keypoints_3d), c(gt_vertices), c(focal_length), c(cam_cx),
                                        c(cam_cy))
    out = {"joint_vis_mask": jv, "vertex_vis_mask": vv,
           "g_mpjpe": err[..., 0], "g_mpjpe_vis": err[..., 1], "g_mpjpe_invis": err[..., 2],
           "mpjpe": err[..., 3], "mpjpe_vis": err[..., 4], "mpjpe_invis": err[..., 5],
           "v2v": err[..., 6], "v2v_vis": err[..., 7], "v2v_invis": err[..., 8],
           "std_joints": div[:, 0], "std_joints_vis": div[:, 1], "std_joints_invis": div[:, 2],
           "apd_joints": div[:, 3], "apd_joints_vis": div[:, 4], "apd_joints_invis": div[:, 5]}
    # PA-MPJPE (:418-436): Procrustes-align every sample's pelvis-aligned joints to the ground truth
    pa = pj - pj[:, :, :1]
    ga = (gt_keypoints_3d - gt_keypoints_3d[:, :1]).float().unsqueeze(1).expand(bs, S, J, 3)
    mask = None
    if eval_with_vis_mask_pa:
        mask = jv.view(bs, 1, J, 1).expand(bs, S, J, 3).reshape(bs * S, J, 3).float().contiguous()
    _, pa_err = eng.procrustes(pa.reshape(bs * S, J, 3).contiguous(), ga.reshape(bs * S, J, 3).contiguous(), mask)
    pa_err = pa_err.view(bs, S, J)
    jm = jv.view(bs, 1, J).float()
    out["pa_mpjpe"] = pa_err.mean(dim=-1)
    out["pa_mpjpe_vis"] = (pa_err * jm).sum(dim=-1)
    out["pa_mpjpe_invis"] = (pa_err * (1 - jm)).sum(dim=-1)
    out["joint_vis_num"] = jv.sum()
    out["vertex_vis_num"] = vv.sum()
    return out

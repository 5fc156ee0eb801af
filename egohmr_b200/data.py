"""Stage-1 hand-off and scene-cloud ingestion: the two data formats either side of the sampling path
that the reference reads in `dataloaders/egobody_dataset.py`, and the batch dict the sampler consumes.

* stage-1 result: `results.pkl` written by test_prohmr_scene.py, key `pred_cam_full_list` [n_frames, 3]
  (egobody_dataset.py:94-98, :275-276) -> `batch['stage1_transl_full']`, which test_egohmr.py:243-245 puts into
  `batch['smpl_params']['transl']`;
* scene cloud: one `.npy` per frame, 20 000 points in scene coordinates (egobody_dataset.py:213-225), moved to the camera
  frame with a 4 x 4 transform (utils/geometry.py:137-141) and optionally subsampled (:270-273).
Host-side plumbing only; nothing here is on the timed path."""
import pickle

import numpy as np
import torch


def load_stage1_translations(results_pkl, spacing=1):
    """-> float32 [n, 3]: `pkl.load(fp)['pred_cam_full_list'].astype(float)[::spacing]` (egobody_dataset.py:94-98)."""
    with open(results_pkl, "rb") as fp:
        res = pickle.load(fp)
    if "pred_cam_full_list" not in res:
        raise KeyError(f"{results_pkl} has no 'pred_cam_full_list' (keys: {sorted(res)[:8]})")
    t = np.asarray(res["pred_cam_full_list"], dtype=np.float64)[::spacing]
    if t.ndim != 2 or t.shape[1] != 3:
        raise ValueError(f"{results_pkl}: pred_cam_full_list must be [n, 3], got {t.shape}")
    return t.astype(np.float32)


def points_coord_trans(xyz, trans):
    """utils/geometry.py:137-141: xyz [N, 3] in the source frame -> target frame, trans = source-to-target 4 x 4."""
    trans = np.asarray(trans)
    return xyz.dot(trans[:3, :3].transpose()) + trans[:3, 3].reshape(1, 3)


def load_scene_cloud(npy_path, trans=None, downsample_rate=1, n_points=None):
    """-> float32 [n_pts, 3] in the camera frame: np.load (egobody_dataset.py:217), `points_coord_trans` with the
    scene-to-camera transform (:225), `.astype(float32)[::downsample_rate]` (:271-272).  `n_points` checks the count the
    caller batches on (the released clouds hold 20 000 points)."""
    pts = np.load(npy_path)
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise ValueError(f"{npy_path}: expected an [n, 3] point cloud, got {pts.shape}")
    if trans is not None:
        pts = points_coord_trans(pts, trans)
    pts = pts.astype(np.float32)[::downsample_rate]
    if n_points is not None and pts.shape[0] != n_points:
        raise ValueError(f"{npy_path}: {pts.shape[0]} points after subsampling, expected {n_points}")
    return pts


def make_batch(img, orig_keypoints_2d, fx, cam_cx, cam_cy, box_center, box_size, transl, scene_clouds, device="cuda:0",
               fx_norm_coeff=1500.0, two_stage=True, gt_transl=None):
    """The batch dict `val_losses` / `EgoHMR.forward` read (producer: egobody_dataset.py:241-277; consumers:
    egohmr.py:175-213,232), from per-frame arrays:
      img [B,3,224,224] ImageNet-normalised crops; orig_keypoints_2d [B,25,3] OpenPose; fx [B] in PIXELS (stored divided by
      `fx_norm_coeff`, :262); cam_cx, cam_cy [B]; box_center [B,2]; box_size [B]; scene_clouds: list of [N,3] arrays or one
      [B,N,3] array in the camera frame; transl [B,3]: the stage-1 translations (`two_stage`, test_egohmr.py:243-245) or
      the ground truth."""
    f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(device)
    clouds = np.stack([np.asarray(c, dtype=np.float32) for c in scene_clouds]) if isinstance(scene_clouds, (list, tuple)) \
        else np.asarray(scene_clouds, dtype=np.float32)
    B = np.asarray(img).shape[0]
    if clouds.shape[0] != B or clouds.ndim != 3 or clouds.shape[2] != 3:
        raise ValueError(f"scene_clouds must be [B, N, 3] with B = {B}, got {clouds.shape}")
    batch = {"img": f(img), "orig_keypoints_2d": f(orig_keypoints_2d), "fx": f(np.asarray(fx, dtype=np.float64) / fx_norm_coeff),
             "cam_cx": f(cam_cx), "cam_cy": f(cam_cy), "box_center": f(box_center), "box_size": f(box_size),
             "scene_pcd_verts_full": f(clouds), "smpl_params": {"transl": f(transl if gt_transl is None or two_stage else gt_transl)}}
    if two_stage:
        batch["stage1_transl_full"] = batch["smpl_params"]["transl"]
    return batch

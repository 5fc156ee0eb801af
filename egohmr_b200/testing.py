"""Helpers shared by tests/, bench.py and smoke(): build the model + sampler from `synth` weights."""
import types

import numpy as np
import torch

from . import synth


def make_cfg():
    """The config fields the path reads (configs/prohmr.yaml:38-59 via yacs in the reference)."""
    ns = types.SimpleNamespace
    return ns(MODEL=ns(BACKBONE=ns(NUM_LAYERS=50, OUT_CHANNELS=2048)), CAM=ns(FX_NORM_COEFF=1500.0),
              EXTRA=ns(FOCAL_LENGTH=5000.0))


def torch_batch(batch_np, device):
    out = {}
    for k, v in batch_np.items():
        out[k] = torch_batch(v, device) if isinstance(v, dict) else torch.from_numpy(np.asarray(v)).to(device)
    return out


def build_model(hid=1024, n_blocks=4, T=50, respacing="ddim5", seed=0, device="cuda:0", diffuse_fuse=True,
                collision=True, only_mask_img_cond=True, nonlocal_layer=False):
    """-> (model, diffusion, state_dict(numpy), smpl_model, Xmean, Xstd) with the reference's test-default flags
    (test_egohmr.py:53-78, 112-118)."""
    from .diffusion.model_util import create_gaussian_diffusion
    from .models.egohmr.egohmr import EgoHMR
    smpl_model = synth.make_smpl_model(seed)
    sd = synth.make_state_dict(seed, hid=hid, n_blocks=n_blocks, init_betas=smpl_model["init_betas"])
    if nonlocal_layer:
        synth.add_nonlocal(sd, seed, hid)
    mean, std = synth.body_rep_stats(seed)
    dev = torch.device(device)
    model = EgoHMR(cfg=make_cfg(), device=dev, body_rep_mean=torch.from_numpy(mean).to(dev),
                   body_rep_std=torch.from_numpy(std).to(dev), with_focal_length=True, with_bbox_info=True,
                   with_cam_center=True, scene_feat_dim=512, scene_type="cube", scene_cano=True, cond_mask_prob=0.0,
                   only_mask_img_cond=only_mask_img_cond, pelvis_vis_loosen=True, diffuse_fuse=diffuse_fuse, diffusion_blk=n_blocks,
                   gcn_hid_dim=hid, gcn_nonlocal_layer=nonlocal_layer, smpl_model=smpl_model,
                   collision_model=SyntheticCollision() if collision else None)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    model.eval()
    diffusion = create_gaussian_diffusion(num_diffusion_timesteps=T, timestep_respacing=respacing,
                                          body_rep_mean=torch.from_numpy(mean).to(dev),
                                          body_rep_std=torch.from_numpy(std).to(dev))
    return model, diffusion, sd, smpl_model, mean, std


class SyntheticCollision:
    """Stand-in for COAP with its call signature (reference: models/egohmr/egohmr.py:117-120,509,555):
    `collision_loss(points[1,n,3], smpl_output, ret_collision_mask=None) -> scalar`, `query(points, smpl_output) ->
    occupancy[1,n]`.  COAP (unpinned git dependency + downloaded weights) cannot be reproduced offline, so BOTH the
    reference golden run (tests/golden/ref_standins.py) and this repo plug in this smooth analytic penalty.  It touches
    all three SMPLOutput fields the reference hands over — joints, vertices and the axis-angle full_pose — so every
    gradient path of guide_coll is exercised."""

    radius = 0.25
    vert_stride = 53

    def eval(self):
        return self

    def parameters(self):
        return []

    def _occ(self, points, smpl_output):
        j = smpl_output.joints[:, :24]
        d2 = ((points[:, :, None, :] - j[:, None, :, :]) ** 2).sum(-1)                 # [1,n,24]
        occ_j = torch.exp(-d2 / (2 * self.radius ** 2)).max(dim=-1).values             # [1,n]
        v = smpl_output.vertices[:, :: self.vert_stride]
        dv2 = ((points[:, :, None, :] - v[:, None, :, :]) ** 2).sum(-1)                # [1,n,V/stride]
        occ_v = torch.exp(-dv2 / (2 * (0.5 * self.radius) ** 2)).mean(dim=-1)
        return 0.5 * occ_j + 0.5 * occ_v

    def collision_loss(self, points, smpl_output, ret_collision_mask=None):
        reg = 1e-2 * (smpl_output.full_pose.reshape(1, -1) ** 2).mean()
        return self._occ(points, smpl_output).mean() + reg

    def query(self, points, smpl_output):
        return self._occ(points, smpl_output)


class BatchedSyntheticCollision(SyntheticCollision):
    """The same penalty behind the batched interface `EgoHMR.guide_coll` / `eval_coll` prefer when a collision model
    offers it: every body in one call, the per-body crop given as a mask instead of a compacted point list."""

    def collision_loss_batched(self, points, mask, smpl_output):
        m = mask.to(points.dtype)
        n = m.sum(dim=1)
        occ = self._occ(points, smpl_output)                                           # [B,N]
        reg = 1e-2 * (smpl_output.full_pose.reshape(points.shape[0], -1) ** 2).mean(dim=1)
        loss = (occ * m).sum(dim=1) / n.clamp_min(1) + reg
        return torch.where(n > 0, loss, torch.zeros_like(loss))

    def query_batched(self, points, mask, smpl_output):
        return self._occ(points, smpl_output)

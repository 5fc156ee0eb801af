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


def build_model(hid=1024, n_blocks=4, T=50, respacing="ddim5", seed=0, device="cuda:0", diffuse_fuse=True):
    """-> (model, diffusion, state_dict(numpy), smpl_model, Xmean, Xstd) with the reference's test-default flags
    (test_egohmr.py:53-78, 112-118)."""
    from .diffusion.model_util import create_gaussian_diffusion
    from .models.egohmr.egohmr import EgoHMR
    smpl_model = synth.make_smpl_model(seed)
    sd = synth.make_state_dict(seed, hid=hid, n_blocks=n_blocks, init_betas=smpl_model["init_betas"])
    mean, std = synth.body_rep_stats(seed)
    dev = torch.device(device)
    model = EgoHMR(cfg=make_cfg(), device=dev, body_rep_mean=torch.from_numpy(mean).to(dev),
                   body_rep_std=torch.from_numpy(std).to(dev), with_focal_length=True, with_bbox_info=True,
                   with_cam_center=True, scene_feat_dim=512, scene_type="cube", scene_cano=True, cond_mask_prob=0.0,
                   only_mask_img_cond=True, pelvis_vis_loosen=True, diffuse_fuse=diffuse_fuse, diffusion_blk=n_blocks,
                   gcn_hid_dim=hid, smpl_model=smpl_model)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    model.eval()
    diffusion = create_gaussian_diffusion(num_diffusion_timesteps=T, timestep_respacing=respacing,
                                          body_rep_mean=torch.from_numpy(mean).to(dev),
                                          body_rep_std=torch.from_numpy(std).to(dev))
    return model, diffusion, sd, smpl_model, mean, std

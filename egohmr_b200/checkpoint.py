"""Ingestion of the reference's released artefacts, mirroring test_egohmr.py:107-126:

    logdir = dirname(args.checkpoint)
    stats  = np.load(logdir + '/preprocess_stats/preprocess_stats.npz')          # Xmean, Xstd  [144]
    model  = EgoHMR(cfg, device, body_rep_mean=Xmean, body_rep_std=Xstd, ...)
    model.load_state_dict(torch.load(args.checkpoint)['state_dict'], strict=False)

The checkpoint is the reference's own `.pt` (keys `backbone.*`, `scene_enc.*`, `transl_enc.*`, `beta_layer.*`,
`diffusion_model.*`, `embed_timestep.*`, `input_process.*`, plus `smpl*.*` / `coap` buffers that live elsewhere here).
"""
import os

import numpy as np
import torch

OWNED_PREFIXES = ("backbone.", "scene_enc.", "transl_enc.", "beta_layer.", "diffusion_model.", "embed_timestep.",
                  "input_process.")
FOREIGN_PREFIXES = ("smpl.", "smpl_male.", "smpl_female.")   # SMPL / COAP buffers saved with the reference module


def load_preprocess_stats(checkpoint_path, device="cpu"):
    """-> (body_rep_mean, body_rep_std) float32 tensors [144] from `<logdir>/preprocess_stats/preprocess_stats.npz`
    (test_egohmr.py:108-111)."""
    logdir = os.path.dirname(checkpoint_path)
    path = os.path.join(logdir, "preprocess_stats", "preprocess_stats.npz")
    stats = np.load(path)
    for k in ("Xmean", "Xstd"):
        if k not in stats:
            raise KeyError(f"{path} has no '{k}' (keys: {list(stats.keys())})")
    mean = torch.from_numpy(np.asarray(stats["Xmean"])).float().reshape(-1).to(device)
    std = torch.from_numpy(np.asarray(stats["Xstd"])).float().reshape(-1).to(device)
    if mean.numel() != 144 or std.numel() != 144:
        raise ValueError(f"{path}: expected 144-d Xmean / Xstd, got {tuple(mean.shape)} / {tuple(std.shape)}")
    if not bool((std > 0).all()):
        raise ValueError(f"{path}: Xstd must be positive")
    return mean, std


def load_checkpoint(model, checkpoint_path, strict_owned=True):
    """`model.load_state_dict(torch.load(path)['state_dict'], strict=False)` (test_egohmr.py:125-126) with a report:
    -> dict(loaded=[...], ignored_foreign=[...], unexpected=[...], missing=[...]).  `strict_owned` raises if a
    parameter of the accelerated path (denoiser, encoders, heads) is absent from the file or has another shape."""
    weights = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    sd = weights["state_dict"] if isinstance(weights, dict) and "state_dict" in weights else weights
    own = model.state_dict()
    loaded, foreign, unexpected, bad_shape = [], [], [], []
    for k, v in sd.items():
        if k in own:
            if tuple(own[k].shape) != tuple(v.shape):
                bad_shape.append((k, tuple(v.shape), tuple(own[k].shape)))
            else:
                loaded.append(k)
        elif k.startswith(FOREIGN_PREFIXES) or ".coap." in k:
            foreign.append(k)
        else:
            unexpected.append(k)
    missing = [k for k in own if k not in sd and k.startswith(OWNED_PREFIXES) and not k.endswith("num_batches_tracked")]
    if bad_shape:
        raise RuntimeError(f"checkpoint shapes differ from the model's (gcn_hid_dim / diffusion_blk mismatch?): {bad_shape[:3]}")
    if strict_owned and missing:
        raise RuntimeError(f"checkpoint lacks {len(missing)} parameters of the sampling path, e.g. {missing[:5]}")
    model.load_state_dict({k: sd[k] for k in loaded}, strict=False)
    model.eval()
    return {"loaded": loaded, "ignored_foreign": foreign, "unexpected": unexpected, "missing": missing}

"""oracle/torch_eager.py (the eager-PyTorch restatement timed by `bench.py --impl reference-cuda`) against the vectors
the UNMODIFIED reference produced (tests/golden/make_golden.py): it must be the reference's dataflow AND its numbers."""
import os

import numpy as np
import torch

from egohmr_b200 import synth
from oracle import torch_eager


def _case(golden_dir, name, dtype):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    hid, nb, n_img, T, resp = int(g["hid"]), int(g["n_blocks"]), int(g["n_img"]), int(g["T"]), str(g["respacing"])
    model, sch = torch_eager.build(hid, nb, T, resp, "cpu", dtype=dtype)
    b = synth.make_batch(0, n_img)
    batch = {k: (torch.from_numpy(v) if not isinstance(v, dict) else {kk: torch.from_numpy(vv) for kk, vv in v.items()})
             for k, v in b.items()}
    noise = torch.from_numpy(synth.make_noise(0, 1, n_img, sch.num_timesteps)[0])
    out = torch_eager.val_losses(model, sch, batch, [n_img, 144], "ddim" if resp else "ddpm", noise=noise)
    return g, out


def test_eager_ddim5_fp64_matches_reference(golden_dir):
    g, out = _case(golden_dir, "ddim5_T50_hid1024_f64", torch.float64)
    assert np.abs(out["pred_x_start"].numpy() - g["pred_x_start"]).max() < 1e-11
    assert np.abs(out["pred_pose_6d"].numpy() - g["pred_pose_6d"]).max() < 1e-11
    assert np.abs(out["pred_smpl_params"]["betas"].numpy() - g["betas"]).max() < 1e-11
    assert np.abs(out["pred_vertices"].numpy() - g["pred_vertices"]).max() < 5e-6   # reference runs SMPL in fp32
    assert np.abs(out["pred_keypoints_2d_full"].numpy() - g["pred_keypoints_2d_full"]).max() < 5e-6


def test_eager_ddim5_fp32_matches_reference(golden_dir):
    g, out = _case(golden_dir, "ddim5_T50_hid1024_f32", torch.float32)
    assert np.abs(out["pred_x_start"].numpy() - g["pred_x_start"]).max() < 5e-6
    assert np.abs(out["pred_vertices"].numpy() - g["pred_vertices"]).max() < 2e-5


def test_eager_ddpm50_fp64_matches_reference(golden_dir):
    g, out = _case(golden_dir, "ddpm_T50_hid256_f64", torch.float64)
    assert np.abs(out["pred_x_start"].numpy() - g["pred_x_start"]).max() < 2e-6

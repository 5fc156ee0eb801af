"""EgoHMR.forward + the sampling loops, end to end (numpy).  TEST INFRASTRUCTURE (see oracle/__init__).

`sample` reproduces what test_egohmr.py:251-255 obtains from `diffusion_sample.val_losses(...)` for one chain,
including the reference's dataflow: by default the conditioning is computed once (`hoist=True`; exact in eval mode),
`hoist=False` recomputes the encoders and SMPL on every step exactly like the reference (CPU-baseline timing).
"""
import numpy as np

from . import encoders, gcn, geometry, schedule, smpl


def forward(sd, adj, n_blocks, smpl_model, batch, x_t, t_orig, mean, std, cond=None, dtype=np.float32,
            diffuse_fuse=True, with_smpl=True, only_mask_img_cond=True):
    """EgoHMR.forward (models/egohmr/egohmr.py:173-303) -> the reference's output dict (numpy)."""
    if cond is None:
        cond = encoders.conditioning(sd, batch, dtype)
    B = x_t.shape[0]
    x0, oc, ou = gcn.denoise(sd, adj, n_blocks, x_t.astype(dtype), t_orig, cond["img_feats"], cond["rest_feats"],
                             cond["vis"], diffuse_fuse, only_mask_img_cond)
    out = {"pred_x_start": x0, "out_cond": oc, "out_uncond": ou}
    pose6d = x0 * std.astype(dtype) + mean.astype(dtype)  # :258
    R = geometry.rot6d_to_rotmat(pose6d, "diffusion").reshape(B, 24, 3, 3)  # :260
    out["pred_pose_6d"] = pose6d
    out["pred_smpl_params"] = {"global_orient": R[:, [0]], "body_pose": R[:, 1:], "betas": cond["betas"]}
    if with_smpl:
        so = smpl.smpl_forward(smpl_model, R, cond["betas"])  # :276
        out["pred_keypoints_3d"] = so["joints"]
        out["pred_vertices"] = so["vertices"]
        transl = cond["transl"]
        out["pred_keypoints_3d_full"] = so["joints"] + transl[:, None, :]  # :294
        fx = batch["fx"].astype(dtype)
        focal = np.stack([fx, fx], axis=1) * np.dtype(dtype).type(1500.0)  # :283-285
        center = np.stack([batch["cam_cx"], batch["cam_cy"]], axis=1).astype(dtype)
        kp2d = geometry.perspective_projection(so["joints"], transl, focal, center)  # :295-298
        kp2d[:, :, 0] = kp2d[:, :, 0] / 1920 - 0.5
        kp2d[:, :, 1] = kp2d[:, :, 1] / 1080 - 0.5
        out["pred_keypoints_2d_full"] = kp2d
    return out


def sample(sd, adj, n_blocks, smpl_model, batch, sch, noise, mean, std, mode="ddim", dtype=np.float32, hoist=True,
           diffuse_fuse=True, grad_fn=None, cond_grad_weight=1.0, trace=None, only_mask_img_cond=True,
           skip_timesteps=0, init_data=None):
    """One chain of p_sample_loop / ddim_sample_loop (gaussian_diffusion.py:391-508, 618-718).

    noise: [n_steps+1, B, 144] in the reference's draw order — noise[0] is `th.randn(*shape)` (:478), noise[1+k] the
    `th.randn_like(x)` of the k-th executed step (:331/:547; DDIM draws it too but multiplies it by sigma = 0).
    Returns the last step's output dict ('other_outputs', :443,780) plus 'sample'."""
    cond = encoders.conditioning(sd, batch, dtype) if hoist else None
    x = noise[0].astype(dtype)
    if skip_timesteps and init_data is None:  # gaussian_diffusion.py:480-481
        init_data = np.zeros_like(x)
    first = sch.num_timesteps - skip_timesteps - 1  # :483
    if init_data is not None:  # :485-487 q_sample(init_data, t = first kept index, noise = the initial draw)
        c0, c1 = np.float32(sch.sqrt_alphas_cumprod[first]), np.float32(sch.sqrt_one_minus_alphas_cumprod[first])
        x = (dtype(c0) * init_data.astype(dtype) + dtype(c1) * x).astype(dtype)
    out = None
    for k, i in enumerate(range(first, -1, -1)):
        B = x.shape[0]
        t_orig = np.full(B, sch.timestep_map[i], dtype=np.int64)  # respace.py:124-126
        last = i == 0
        out = forward(sd, adj, n_blocks, smpl_model, batch, x, t_orig, mean, std, cond=cond, dtype=dtype,
                      diffuse_fuse=diffuse_fuse, with_smpl=(last or not hoist), only_mask_img_cond=only_mask_img_cond)
        x0 = out["pred_x_start"]
        if mode == "ddim":
            x_new = schedule.ddim_update(sch, x, x0, i, np.dtype(dtype).type)
        else:
            g = grad_fn(x, out, i) if (grad_fn is not None and i <= 10) else None
            x_new = schedule.ddpm_update(sch, x, x0, i, noise[1 + k].astype(dtype), g, cond_grad_weight,
                                         np.dtype(dtype).type)
        if trace is not None:
            trace.append({"t": i, "x_t": x.copy(), "pred_x_start": x0.copy(), "sample": x_new.copy()})
        x = x_new
    out["sample"] = x
    return out

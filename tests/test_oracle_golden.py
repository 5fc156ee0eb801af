"""The oracle (oracle/, numpy) against the vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py).
This is what pins the oracle: every later CUDA-vs-oracle comparison inherits its meaning from these checks."""
import os

import numpy as np
import pytest

from egohmr_b200 import synth
from oracle import egohmr as o_egohmr, encoders, gcn, geometry, schedule

TABLES = ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
          "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2")


@pytest.mark.parametrize("T", [50, 100, 1000])
@pytest.mark.parametrize("resp", ["", "ddim5"])
def test_schedule_tables_bit_identical(golden_dir, T, resp):
    tab = np.load(os.path.join(golden_dir, "schedule_tables.npz"))
    s = schedule.Schedule(T, resp)
    tag = f"T{T}_{resp or 'ddpm'}"
    assert list(tab[tag + "_timestep_map"]) == s.timestep_map
    for k in TABLES:
        assert np.array_equal(tab[f"{tag}_{k}"], getattr(s, k)), k  # float64, bit for bit


def test_ddim5_timestep_maps():
    # T=50/100/1000 'ddim5' -> strides 10/20/200
    assert schedule.Schedule(50, "ddim5").timestep_map == [0, 10, 20, 30, 40]
    assert schedule.Schedule(100, "ddim5").timestep_map == [0, 20, 40, 60, 80]
    assert schedule.Schedule(1000, "ddim5").timestep_map == [0, 200, 400, 600, 800]


def test_rot6d_golden(golden_dir):
    so = np.load(os.path.join(golden_dir, "small_ops.npz"))
    R64 = geometry.rot6d_to_rotmat(so["rot6d_x"].astype(np.float64))
    assert np.abs(R64 - so["rot6d_R64"]).max() < 1e-14
    assert np.abs(geometry.rot6d_to_rotmat(so["rot6d_x"]) - so["rot6d_R32"]).max() < 5e-7


def test_modulated_graph_conv_golden(golden_dir):
    so = np.load(os.path.join(golden_dir, "small_ops.npz"))
    y = gcn.modulated_graph_conv(so["gconv_x"].astype(np.float64), so["gconv_W"], so["gconv_M"],
                                 synth.skeleton_adjacency(), so["gconv_adj2"], so["gconv_bias"])
    assert np.abs(y - so["gconv_y64"]).max() < 1e-12
    y32 = gcn.modulated_graph_conv(so["gconv_x"], so["gconv_W"], so["gconv_M"], synth.skeleton_adjacency(),
                                   so["gconv_adj2"], so["gconv_bias"])
    assert np.abs(y32 - so["gconv_y32"]).max() < 2e-5


def _run_case(golden_dir, case, dtype, nonlocal_layer=False, **flags):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    hid, nb, n_img, T, resp = int(g["hid"]), int(g["n_blocks"]), int(g["n_img"]), int(g["T"]), str(g["respacing"])
    smpl = synth.make_smpl_model(0)
    sd = synth.make_state_dict(0, hid=hid, n_blocks=nb, init_betas=smpl["init_betas"])
    if nonlocal_layer:
        synth.add_nonlocal(sd, 0, hid)
    b = synth.make_batch(0, n_img)
    mean, std = synth.body_rep_stats(0)
    sch = schedule.Schedule(T, resp)
    assert sch.timestep_map == list(g["timestep_map"])
    noise = synth.make_noise(0, 1, n_img, sch.num_timesteps)[0]
    trace = []
    out = o_egohmr.sample(sd, synth.skeleton_adjacency(), nb, smpl, b, sch, noise, mean, std,
                          "ddim" if resp else "ddpm", dtype=dtype, trace=trace, **flags)
    return g, out, np.stack([t["pred_x_start"] for t in trace]), np.stack([t["x_t"] for t in trace])


def test_ddim5_full_size_fp64_trace(golden_dir):
    """T=50 ddim5, hid 1024 x 4 blocks, 2 images: every step's x_t and pred_x_start, float64."""
    g, out, x0s, xts = _run_case(golden_dir, "ddim5_T50_hid1024_f64", np.float64)
    assert list(g["trace_t_orig"]) == [40, 30, 20, 10, 0]
    assert np.abs(x0s - g["trace_x0"]).max() < 1e-12
    assert np.abs(xts - g["trace_x_t"]).max() < 1e-12
    assert np.abs(out["pred_pose_6d"] - g["pred_pose_6d"]).max() < 1e-12
    R = np.concatenate([g["global_orient"], g["body_pose"]], axis=1)
    Ro = np.concatenate([out["pred_smpl_params"]["global_orient"], out["pred_smpl_params"]["body_pose"]], axis=1)
    assert np.abs(R - Ro).max() < 1e-12
    assert np.abs(out["pred_smpl_params"]["betas"] - g["betas"]).max() < 1e-12
    # the reference casts SMPL inputs to fp32 (egohmr.py:276) even in a float64 run
    assert np.abs(out["pred_vertices"] - g["pred_vertices"]).max() < 5e-6
    assert np.abs(out["pred_keypoints_2d_full"] - g["pred_keypoints_2d_full"]).max() < 5e-6
    assert np.array_equal(encoders.vis_mask_smpl(synth.make_batch(0, 2)["orig_keypoints_2d"]),
                          g["vis_mask_smpl"].astype(bool))


def test_ddim5_full_size_fp32(golden_dir):
    g, out, x0s, xts = _run_case(golden_dir, "ddim5_T50_hid1024_f32", np.float32)
    assert np.abs(x0s - g["trace_x0"]).max() < 5e-6      # fp32 summation-order noise
    assert np.abs(out["pred_vertices"] - g["pred_vertices"]).max() < 2e-5
    assert np.abs(out["pred_keypoints_3d"] - g["pred_keypoints_3d"]).max() < 2e-5


def test_ddpm50_fp64_trace(golden_dir):
    """Full 50-step DDPM chain (p_sample, no guidance), hid 256 x 2 blocks, 3 images."""
    g, out, x0s, xts = _run_case(golden_dir, "ddpm_T50_hid256_f64", np.float64)
    assert list(g["trace_t_orig"]) == list(range(49, -1, -1))
    # exp(0.5*log_var) is evaluated in fp32 by torch and by numpy: 1-ulp libm differences times the noise
    assert np.abs(x0s - g["trace_x0"]).max() < 1e-6
    assert np.abs(xts - g["trace_x_t"]).max() < 2e-6


def test_ddim5_mask_all_conditions_fp64(golden_dir):
    """diffuse_fuse with only_mask_img_cond=False: the second pass zeroes every condition (egohmr.py:157-158)."""
    g, out, x0s, xts = _run_case(golden_dir, "ddim5_T50_hid256_maskall_f64", np.float64, only_mask_img_cond=False)
    assert np.abs(x0s - g["trace_x0"]).max() < 1e-12
    assert np.abs(out["pred_vertices"] - g["pred_vertices"]).max() < 5e-6
    # and it is a different computation from the default flag
    _, out2, x0d, _ = _run_case(golden_dir, "ddim5_T50_hid256_maskall_f64", np.float64)
    assert np.abs(x0d - g["trace_x0"]).max() > 1e-4


def test_ddim5_no_diffuse_fuse_fp64(golden_dir):
    """diffuse_fuse=False: a single image-conditioned pass, no fuse-select (egohmr.py:239)."""
    g, out, x0s, xts = _run_case(golden_dir, "ddim5_T50_hid256_nofuse_f64", np.float64, diffuse_fuse=False)
    assert np.abs(x0s - g["trace_x0"]).max() < 1e-12
    assert np.abs(out["pred_vertices"] - g["pred_vertices"]).max() < 5e-6


def test_ddim5_nonlocal_layer_fp64(golden_dir):
    """gcn_nonlocal_layer=True: NONLocalBlock2D after the residual blocks (modulated_gcn.py:103-109)."""
    g, out, x0s, xts = _run_case(golden_dir, "ddim5_T50_hid256_nonlocal_f64", np.float64, nonlocal_layer=True)
    assert np.abs(x0s - g["trace_x0"]).max() < 1e-12
    _, _, x0_plain, _ = _run_case(golden_dir, "ddim5_T50_hid256_nonlocal_f64", np.float64)
    assert np.abs(x0_plain - g["trace_x0"]).max() > 1e-3      # the block is not a no-op with these parameters


def test_encoders_match_reference_features(golden_dir):
    g = np.load(os.path.join(golden_dir, "ddim5_T50_hid1024_f64.npz"))
    smpl = synth.make_smpl_model(0)
    sd = synth.make_state_dict(0, init_betas=smpl["init_betas"])
    c = encoders.conditioning(sd, synth.make_batch(0, 2), np.float64)
    assert np.abs(c["img_feats"] - g["img_feats"]).max() < 1e-10
    assert np.abs(c["rest_feats"][:, :512] - g["scene_feats"]).max() < 1e-10
    assert np.abs(c["rest_feats"][:, 512:640] - g["transl_feat"]).max() < 1e-10


def test_guide_coll_gradient_golden(golden_dir):
    """EgoHMR.guide_coll with the shared synthetic collision term: the oracle's autograd restatement vs the reference."""
    from egohmr_b200.testing import SyntheticCollision
    from oracle import guidance
    g = np.load(os.path.join(golden_dir, "guide_grad_f64.npz"))
    sm = synth.make_smpl_model(0)
    b = synth.make_batch(0, 3)
    mean, std = synth.body_rep_stats(0)
    pts = b["scene_pcd_verts_full"] - b["smpl_params"]["transl"][:, None]
    gr = guidance.guide_coll(sm, SyntheticCollision(), g["x_t"], g["betas"], pts, mean, std)
    assert np.abs(g["grad"]).max() > 1e-3
    assert np.abs(gr - g["grad"]).max() < 2e-8        # the reference runs SMPL in fp32 (egohmr.py:537 casts .float())
    nz = sorted(set(np.nonzero(gr.reshape(3, 24, 6).any(axis=(0, 2)))[0].tolist()))
    assert nz == [1, 2, 4, 5, 7, 8, 10, 11]            # only leg joints survive (egohmr.py:567)


def test_rotation_matrix_to_angle_axis_branches():
    """All four quaternion branches + the small-angle path, against the closed form."""
    def rot(axis, ang):
        axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    cases = [([1, 0, 0], 0.3), ([0, 1, 0], 3.0), ([1, 0.1, 0], 3.1), ([0, 0, 1], 3.1), ([0.1, 1, 0.2], 3.0), ([1, 2, 3], 1e-4)]
    R = np.stack([rot(a, t) for a, t in cases])
    aa = geometry.rotation_matrix_to_angle_axis(R)
    for (a, t), v in zip(cases, aa):
        expect = np.asarray(a, np.float64) / np.linalg.norm(a) * t
        assert np.abs(v - expect).max() < 1e-5, (a, t, v, expect)


def test_guided_ddpm100_fp64_trace(golden_dir):
    """cfg-3 shape at small size: DDPM T=100 with the collision-guided mean shift for t <= 10 (gaussian_diffusion.py:378-385)."""
    from egohmr_b200.testing import SyntheticCollision
    from oracle import guidance
    g = np.load(os.path.join(golden_dir, "ddpm_guided_T100_hid256_f64.npz"))
    hid, nb, n_img = int(g["hid"]), int(g["n_blocks"]), int(g["n_img"])
    smpl = synth.make_smpl_model(0)
    sd = synth.make_state_dict(0, hid=hid, n_blocks=nb, init_betas=smpl["init_betas"])
    b = synth.make_batch(0, n_img)
    mean, std = synth.body_rep_stats(0)
    sch = schedule.Schedule(100, "")
    noise = synth.make_noise(0, 1, n_img, 100)[0]
    pts = b["scene_pcd_verts_full"] - b["smpl_params"]["transl"][:, None]
    col = SyntheticCollision()
    grad_fn = lambda x, out, i: guidance.guide_coll(smpl, col, x, out["pred_smpl_params"]["betas"], pts, mean, std)
    trace = []
    out = o_egohmr.sample(sd, synth.skeleton_adjacency(), nb, smpl, b, sch, noise, mean, std, "ddpm", dtype=np.float64,
                          grad_fn=grad_fn, cond_grad_weight=2.0, trace=trace)
    x0s = np.stack([t["pred_x_start"] for t in trace])
    assert np.abs(x0s - g["trace_x0"]).max() < 2e-6
    assert np.abs(out["pred_x_start"] - g["pred_x_start"]).max() < 2e-6


def test_compute_loss_golden(golden_dir):
    """val_losses' default compute_loss=True (what test_egohmr.py:252-255 runs): EgoHMR.compute_loss in eval mode."""
    from oracle import losses
    g = np.load(os.path.join(golden_dir, "val_losses_compute_loss_f32.npz"))
    smpl = synth.make_smpl_model(0)
    sd = synth.make_state_dict(0, hid=256, n_blocks=2, init_betas=smpl["init_betas"])
    b = synth.merge_gt(synth.make_batch(0, 3), synth.make_gt(0, 3))
    mean, std = synth.body_rep_stats(0)
    sch = schedule.Schedule(50, "ddim5")
    out = o_egohmr.sample(sd, synth.skeleton_adjacency(), 2, smpl, b, sch, synth.make_noise(0, 1, 3, 5)[0], mean, std, "ddim",
                          dtype=np.float32)
    assert np.abs(out["pred_x_start"] - g["pred_x_start"]).max() < 5e-6
    loss, ls, nvis = losses.compute_loss(smpl, b, out)
    assert nvis == int(g["joint_vis_num_batch"]) and float(g["loss"]) == 0.0 and loss == 0.0
    for k, v in ls.items():
        assert abs(float(v) - float(g[k])) <= 2e-5 * max(1.0, abs(float(g[k]))), (k, float(v), float(g[k]))


def test_procrustes_golden(golden_dir):
    """oracle/pose_utils.py against utils/pose_utils.py of the reference (PA-MPJPE alignment)."""
    from oracle import pose_utils
    g = np.load(os.path.join(golden_dir, "procrustes.npz"))
    re, hat = pose_utils.reconstruction_error(g["S1"], g["S2"], avg_joint=False)
    assert np.abs(hat - g["hat"]).max() < 2e-5 and np.abs(re - g["re"]).max() < 2e-5
    re_v, _ = pose_utils.reconstruction_error(g["S1"], g["S2"], mask=g["vis"], avg_joint=False)
    assert np.abs(re_v - g["re_vis"]).max() < 2e-5
    assert np.abs(pose_utils.reconstruction_error(g["S1"], g["S2"])[0] - g["re_avg"]).max() < 2e-5


def test_angle_axis_oracle_vs_reference_golden(golden_dir):
    """oracle/geometry.py::rotation_matrix_to_angle_axis pinned directly on the reference's own function
    (utils/konia_transform.py:316-339), including theta -> 0 and theta = pi."""
    from oracle import geometry
    g = np.load(os.path.join(golden_dir, "angle_axis.npz"))
    assert np.abs(geometry.rotation_matrix_to_angle_axis(g["R"].astype(np.float64)) - g["aa64"]).max() < 1e-14
    assert np.abs(geometry.rotation_matrix_to_angle_axis(g["R"]) - g["aa32"]).max() < 5e-7

"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (/root/reference).

Run once in the build container (`python tests/golden/make_golden.py`); the GPU box never runs this.  Inputs, weights
and the SMPL model come from `egohmr_b200.synth` (seeded, regenerable), so only the reference's OUTPUTS are stored.
The reference's RNG draws (`th.randn`, `th.randn_like`) are fed from `synth.make_noise` so every implementation sees
the same noise; nothing inside /root/reference is edited.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from egohmr_b200 import synth  # noqa: E402
import ref_standins  # noqa: E402


def to_torch(batch, dtype):
    return ref_standins.to_torch(batch, dtype)


class NoiseFeed:
    """Replaces torch.randn / torch.randn_like while the reference sampler runs."""

    def __init__(self, noise, dtype):
        self.noise = [torch.from_numpy(n).to(dtype) for n in noise]
        self.i = 0

    def _next(self, shape):
        n = self.noise[self.i]
        self.i += 1
        assert tuple(n.shape) == tuple(shape), (n.shape, shape)
        return n.clone()

    def randn(self, *shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return self._next(shape)

    def randn_like(self, x, **kw):
        return self._next(x.shape)


def build_reference(hid, n_blocks, dtype, seed=0, only_mask_img_cond=True, diffuse_fuse=True, nonlocal_layer=False):
    return ref_standins.build_reference(hid, n_blocks, dtype, seed, only_mask_img_cond, diffuse_fuse, nonlocal_layer)


def run_case(name, T, respacing, hid, n_blocks, n_img, dtype, seed=0, guided=False, only_mask_img_cond=True,
             diffuse_fuse=True, nonlocal_layer=False, trace_every=1, eta=None):
    """`trace_every`: keep every n-th step of the per-step trace (long chains); `eta`: call ddim_sample_loop directly with
    this eta (val_losses hard-codes eta=0.0, gaussian_diffusion.py:770)."""
    from diffusion.model_util import create_gaussian_diffusion
    import diffusion.gaussian_diffusion as gd
    model, mean, std = build_reference(hid, n_blocks, dtype, seed, only_mask_img_cond, diffuse_fuse, nonlocal_layer)
    diffusion = create_gaussian_diffusion(num_diffusion_timesteps=T, timestep_respacing=respacing,
                                          body_rep_mean=torch.from_numpy(mean).to(dtype),
                                          body_rep_std=torch.from_numpy(std).to(dtype))
    n_steps = diffusion.num_timesteps
    batch = to_torch(synth.make_batch(seed, n_img), dtype)
    noise = synth.make_noise(seed, 1, n_img, n_steps)[0]
    feed = NoiseFeed(noise, dtype)
    trace = []
    orig_pmv = diffusion.p_mean_variance.__func__ if hasattr(diffusion.p_mean_variance, "__func__") else None

    # record per-step x_t / pred_xstart by wrapping the model call (does not alter the reference's arithmetic)
    class Wrapped(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, b, t, **kw):
            o = self.m(b, t, **kw)
            trace.append((int(t[0]), b["x_t"].detach().clone().numpy(), o["pred_x_start"].detach().clone().numpy()))
            return o

        def __getattr__(self, k):
            try:
                return super().__getattr__(k)
            except AttributeError:
                return getattr(self.m, k)

    old_randn, old_randn_like = torch.randn, torch.randn_like
    torch.randn, torch.randn_like = feed.randn, feed.randn_like
    try:
        with torch.no_grad():   # the reference re-enables grad inside guide_coll (egohmr.py:518)
            if eta is None:
                out = diffusion.val_losses(model=Wrapped(model), batch=batch, shape=[n_img, 144], progress=False,
                                           clip_denoised=False, cur_epoch=0, timestep_respacing=respacing,
                                           cond_fn_with_grad=guided, cond_grad_weight=2.0, compute_loss=False)
            else:
                model.validation_setup()
                out = diffusion.ddim_sample_loop(model=Wrapped(model), batch=batch, shape=[n_img, 144], progress=False,
                                                 clip_denoised=False, eta=eta, cond_fn_with_grad=guided)["other_outputs"]
    finally:
        torch.randn, torch.randn_like = old_randn, old_randn_like
    g = lambda t: t.detach().numpy()
    if trace_every > 1:
        trace = trace[::trace_every] + ([trace[-1]] if (len(trace) - 1) % trace_every else [])
    rec = {
        "T": T, "respacing": respacing, "hid": hid, "n_blocks": n_blocks, "n_img": n_img, "seed": seed,
        "trace_every": trace_every, "eta": -1.0 if eta is None else float(eta),
        "timestep_map": np.array(diffusion.timestep_map), "trace_t_orig": np.array([t for t, _, _ in trace]),
        "trace_x_t": np.stack([x for _, x, _ in trace]), "trace_x0": np.stack([x for _, _, x in trace]),
        "pred_x_start": g(out["pred_x_start"]), "pred_pose_6d": g(out["pred_pose_6d"]),
        "global_orient": g(out["pred_smpl_params"]["global_orient"]), "body_pose": g(out["pred_smpl_params"]["body_pose"]),
        "betas": g(out["pred_smpl_params"]["betas"]), "pred_keypoints_3d": g(out["pred_keypoints_3d"]),
        "pred_vertices": g(out["pred_vertices"]), "pred_keypoints_2d_full": g(out["pred_keypoints_2d_full"]),
        "vis_mask_smpl": batch["vis_mask_smpl"].numpy(),
    }
    # step-invariant features of the reference's own encoders (inputs for feature-level parity tests)
    with torch.no_grad():
        rec["img_feats"] = g(model.backbone(batch["img"]))
        pts = batch["scene_pcd_verts_full"] - batch["smpl_params"]["transl"].unsqueeze(1)
        rec["scene_feats"] = g(model.scene_enc(pts))
        rec["transl_feat"] = g(model.transl_enc(batch["smpl_params"]["transl"]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in rec.items() if k in ("trace_x0", "pred_vertices")})


def guide_case(name, dtype, n_img=3, seed=0):
    """EgoHMR.guide_coll (egohmr.py:517-570) on a fixed x_t: the gradient the guided sampler adds to the mean."""
    model, mean, std = build_reference(256, 2, dtype, seed)
    batch = to_torch(synth.make_batch(seed, n_img), dtype)
    rng = np.random.default_rng(31)
    x_t = (0.7 * rng.normal(0, 1, (n_img, 144))).astype(np.float32)
    batch["x_t"] = torch.from_numpy(x_t).to(dtype)
    t = torch.tensor([8] * n_img)
    with torch.no_grad():
        out = model(batch, t)          # sets model.scene_pcd_verts and yields the betas guide_coll reads
        grad = model.guide_coll(batch, out, t, compute_grad="x_t")
        ratios = model.eval_coll(out)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x_t=x_t, grad=grad.detach().numpy(),
                        betas=out["pred_smpl_params"]["betas"].detach().numpy(), coll_ratio=np.array(ratios),
                        pred_x_start=out["pred_x_start"].detach().numpy())
    print("wrote", name, "max|grad| =", float(grad.abs().max()), "nonzero joints:",
          sorted(set(np.nonzero(grad.detach().numpy().reshape(n_img, 24, 6).any(axis=(0, 2)))[0].tolist())), "coll", ratios)


def loss_case(name, dtype, n_img=3, seed=0):
    """val_losses with its DEFAULT compute_loss=True, exactly how test_egohmr.py:252-255 calls it: the sampler's last
    output goes through EgoHMR.compute_loss (egohmr.py:305-445), which adds 'losses' and 'joint_vis_num_batch'."""
    from diffusion.model_util import create_gaussian_diffusion
    model, mean, std = build_reference(256, 2, dtype, seed)
    diffusion = create_gaussian_diffusion(num_diffusion_timesteps=50, timestep_respacing="ddim5",
                                          body_rep_mean=torch.from_numpy(mean).to(dtype),
                                          body_rep_std=torch.from_numpy(std).to(dtype))
    b = synth.merge_gt(synth.make_batch(seed, n_img), synth.make_gt(seed, n_img))
    batch = to_torch(b, dtype)
    batch["gender"] = torch.from_numpy(b["gender"])
    batch["smpl_params_is_axis_angle"] = {k: torch.from_numpy(v) for k, v in b["smpl_params_is_axis_angle"].items()}
    feed = NoiseFeed(synth.make_noise(seed, 1, n_img, 5)[0], dtype)
    old_randn, old_randn_like = torch.randn, torch.randn_like
    torch.randn, torch.randn_like = feed.randn, feed.randn_like
    try:
        with torch.no_grad():
            out = diffusion.val_losses(model=model, batch=batch, shape=[n_img, 144], progress=False, clip_denoised=False,
                                       cur_epoch=0, timestep_respacing="ddim5", cond_fn_with_grad=False,
                                       cond_grad_weight=1.0)
    finally:
        torch.randn, torch.randn_like = old_randn, old_randn_like
    rec = {"loss_" + k if not k.startswith("loss") else k: v.detach().numpy() for k, v in out["losses"].items()}
    rec["joint_vis_num_batch"] = out["joint_vis_num_batch"].detach().numpy()
    rec["pred_x_start"] = out["pred_x_start"].detach().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, {k: float(v) for k, v in rec.items() if v.ndim == 0})


def procrustes_case():
    """utils/pose_utils.py reconstruction_error / reconstruction_error_with_vis_mask straight from the reference, on
    PA-MPJPE-shaped inputs (24 joints) including a mirrored target (det < 0 branch) and a planar point set."""
    from utils.pose_utils import (compute_similarity_transform_batch, reconstruction_error,
                                  reconstruction_error_with_vis_mask)
    rng = np.random.default_rng(123)
    P, N = 12, 24
    S1 = rng.normal(0, 0.3, (P, N, 3)).astype(np.float32)
    Rm = np.linalg.qr(rng.normal(0, 1, (P, 3, 3)))[0]
    S2 = (1.7 * np.einsum("pij,pnj->pni", Rm, S1) + rng.normal(0, 1, (P, 1, 3)) + rng.normal(0, 0.02, (P, N, 3))).astype(np.float32)
    S2[3] = S2[3] * np.array([1, 1, -1], np.float32)          # mirrored: forces the det(U V^T) < 0 correction
    S1[5, :, 2] = 0                                           # planar source (rank-2 K)
    S2[6] = rng.normal(0, 0.3, (N, 3)).astype(np.float32)     # unrelated target
    vis = (rng.uniform(0, 1, (P, N, 1)) < 0.7).astype(np.float32).repeat(3, axis=2)
    rec = {"S1": S1, "S2": S2, "vis": vis, "hat": compute_similarity_transform_batch(S1, S2),
           "re": reconstruction_error(S1, S2, avg_joint=False), "re_avg": reconstruction_error(S1, S2),
           "re_vis": reconstruction_error_with_vis_mask(vis, S1, S2, avg_joint=False)}
    np.savez_compressed(os.path.join(HERE, "procrustes.npz"), **rec)
    print("wrote procrustes", rec["re_avg"])


def cfg2_case(name, dtype, n_img=8, S=10, seed=3):
    """configs[1] at its own denoiser size and sample count, through the reference DRIVER's loop (test_egohmr.py:247-266 via
    baseline/ref_harness.driver_loop): 8 distinct images x 10 sequential samples, DDIM-5 of T=50, hid 1024 / 4 blocks.
    The GPU test replicates the 8 images 8 times -> the full 64 x 10 = 640-body batch of the benchmark."""
    model, mean, std = build_reference(1024, 4, dtype, 0)
    diffusion = ref_standins.build_sampler(50, "ddim5", mean, std, dtype)
    batch = to_torch(synth.make_batch(seed, n_img), dtype)
    noise = synth.make_noise(seed, S, n_img, 5)                       # [S, 6, n_img, 144]
    feed = NoiseFeed([noise[n][k] for n in range(S) for k in range(6)], dtype)
    old_randn, old_randn_like = torch.randn, torch.randn_like
    torch.randn, torch.randn_like = feed.randn, feed.randn_like
    xs = []
    try:
        with torch.no_grad():
            for n in range(S):
                out = diffusion.val_losses(model=model, batch=batch, shape=[n_img, 144], progress=False, clip_denoised=False,
                                           cur_epoch=0, timestep_respacing="ddim5", cond_fn_with_grad=False,
                                           cond_grad_weight=1.0, compute_loss=False)
                xs.append({k: out[k].detach().clone().numpy() for k in ("pred_x_start", "pred_keypoints_3d")} |
                          {k: v.detach().clone().numpy() for k, v in out["pred_smpl_params"].items()})
    finally:
        torch.randn, torch.randn_like = old_randn, old_randn_like
    rec = {k: np.stack([x[k] for x in xs], axis=1) for k in xs[0]}    # [n_img, S, ...]
    rec.update(seed=seed, n_img=n_img, S=S)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in rec.items()})


def angle_axis_case():
    """utils/konia_transform.py:316-339 rotation_matrix_to_angle_axis straight from the reference: generic rotations,
    theta -> 0 (identity and tiny angles), theta = pi about several axes, and every branch of the quaternion conversion."""
    from utils.konia_transform import rotation_matrix_to_angle_axis

    def rodrigues(v):
        th_ = np.linalg.norm(v)
        if th_ == 0:
            return np.eye(3)
        k = v / th_
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return np.eye(3) + np.sin(th_) * K + (1 - np.cos(th_)) * K @ K

    rng = np.random.default_rng(5)
    vs = [rng.normal(0, 1, 3) for _ in range(64)]
    vs += [np.zeros(3), np.array([1e-7, 0, 0]), np.array([0, 1e-5, 1e-5]), np.array([1e-3, -1e-3, 1e-3])]
    for ax in (np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), np.array([0, 0, 1.0]), np.array([1.0, 1.0, 0]) / np.sqrt(2),
               np.array([1.0, -2.0, 3.0]) / np.sqrt(14)):
        vs += [np.pi * ax, (np.pi - 1e-4) * ax, (np.pi - 1e-2) * ax, 3.0 * ax, 0.5 * np.pi * ax]
    R = np.stack([rodrigues(v) for v in vs]).astype(np.float32)
    aa32 = rotation_matrix_to_angle_axis(torch.from_numpy(R)).numpy()
    aa64 = rotation_matrix_to_angle_axis(torch.from_numpy(R).double()).numpy()
    np.savez_compressed(os.path.join(HERE, "angle_axis.npz"), R=R, aa32=aa32, aa64=aa64)
    print("wrote angle_axis", R.shape, "max|aa| =", float(np.abs(aa32).max()))


def metrics_case():
    """The evaluation-metric block of the reference driver (test_egohmr.py:373-494: G-MPJPE / MPJPE / PA-MPJPE / V2V with
    visible / invisible splits, per-joint std and APD diversity) executed VERBATIM: the lines are read from the reference
    file at generation time and exec'd on synthetic predictions / ground truth (they are inline script code, not a
    function, and the file itself cannot be imported: pytorch3d / pyrender at its top)."""
    import textwrap
    import types as _t
    from utils.geometry import perspective_projection
    from utils.pose_utils import reconstruction_error, reconstruction_error_with_vis_mask
    src = open(os.path.join(ref_standins.REFERENCE_ROOT, "test_egohmr.py")).read().splitlines()
    first = next(i for i, l in enumerate(src) if "focal_length_proj = focal_length.unsqueeze(-1).repeat(1, 2)" in l)
    last = next(i for i, l in enumerate(src) if "apd_joints_invis_all[step * args.batch_size" in l)
    block = textwrap.dedent("\n".join(src[first:last + 1]))
    rng = np.random.default_rng(11)
    bs, S, V = 5, 4, 1000
    f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    gt_j = f(rng.normal(0, 0.35, (bs, 24, 3)) + np.array([0.0, 0.0, 3.0]))
    gt_j[1, :6, 0] += 4.0      # some joints project outside the 1920 x 1080 image -> invisible
    gt_j[3, 10:, 1] -= 3.0
    gt_v = f(rng.normal(0, 0.35, (bs, V, 3)) + np.array([0.0, 0.0, 3.0]))
    gt_v[1, :2000, 0] += 4.0
    pred_j = gt_j.unsqueeze(1) - gt_j[:, None, [0]] + f(rng.normal(0, 0.05, (bs, S, 24, 3)))      # [bs,S,24,3] around the GT
    pred_v = gt_v.unsqueeze(1) - gt_j[:, None, [0]] + f(rng.normal(0, 0.05, (bs, S, V, 3)))
    transl = f(rng.normal(0, 0.1, (bs, 3)) + np.array([0.0, 0.0, 3.0]))
    pred_pelvis = pred_j[:, :, [0]].clone()
    ns = {"torch": torch, "np": np, "perspective_projection": perspective_projection,
          "reconstruction_error": reconstruction_error, "reconstruction_error_with_vis_mask": reconstruction_error_with_vis_mask,
          "device": "cpu", "step": 0, "curr_batch_size": bs,
          "args": _t.SimpleNamespace(batch_size=bs, num_samples=S, eval_with_vis_mask_pa=False),
          "focal_length": f(rng.uniform(1400, 1600, bs)), "cam_cx": f(np.full(bs, 960.0)), "cam_cy": f(np.full(bs, 540.0)),
          "gt_keypoints_3d": gt_j, "gt_vertices": gt_v, "gt_keypoints_3d_align": gt_j - gt_j[:, [0]],
          "gt_vertices_align": gt_v - gt_j[:, [0]], "pred_keypoints_3d_align": pred_j - pred_pelvis,
          "pred_vertices_align": pred_v - pred_pelvis, "pred_keypoints_3d_full": pred_j + transl[:, None, None]}
    for k in ("joint_vis_num_list", "joint_invis_num_list", "vertex_vis_num_list", "vertex_invis_num_list"):
        ns[k] = []
    names = ["g_mpjpe_all", "g_mpjpe_vis_all_list", "g_mpjpe_invis_all_list", "mpjpe_all", "mpjpe_vis_all_list",
             "mpjpe_invis_all_list", "pa_mpjpe_all", "pa_mpjpe_vis_all_list", "pa_mpjpe_invis_all_list", "v2v_all",
             "v2v_vis_all_list", "v2v_invis_all_list"]
    for k in names:
        ns[k] = np.zeros((bs, S))
    names1 = ["std_joints_all", "std_joints_vis_all", "std_joints_invis_all", "apd_joints_all", "apd_joints_vis_all",
              "apd_joints_invis_all"]
    for k in names1:
        ns[k] = np.zeros(bs)
    exec(compile(block, "test_egohmr.py[metrics block]", "exec"), ns)
    rec = {k: ns[k] for k in names + names1}
    rec.update(pred_keypoints_3d=pred_j.numpy(), pred_vertices=pred_v.numpy(), transl=transl.numpy(), gt_keypoints_3d=gt_j.numpy(),
               gt_vertices=gt_v.numpy(), focal_length=ns["focal_length"].numpy(), cam_cx=ns["cam_cx"].numpy(),
               cam_cy=ns["cam_cy"].numpy(), joint_vis_mask=ns["joint_vis_mask"].numpy(),
               vertex_vis_num=np.array(ns["vertex_vis_num_list"]), joint_vis_num=np.array(ns["joint_vis_num_list"]))
    # store the inputs compactly: vertices as float16 offsets would change them, so keep float32 but only 5 x 4 bodies
    np.savez_compressed(os.path.join(HERE, "eval_metrics.npz"), **rec)
    print("wrote eval_metrics", {k: np.round(ns[k].mean(), 5) for k in ("mpjpe_all", "pa_mpjpe_all", "v2v_all", "apd_joints_all")},
          "vis joints", ns["joint_vis_num_list"])


def schedule_tables():
    from diffusion.model_util import create_gaussian_diffusion
    rec = {}
    for T in (50, 100, 1000):
        for resp in ("", "ddim5"):
            d = create_gaussian_diffusion(num_diffusion_timesteps=T, timestep_respacing=resp)
            tag = f"T{T}_{resp or 'ddpm'}"
            rec[tag + "_timestep_map"] = np.array(d.timestep_map)
            for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                      "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                      "posterior_mean_coef1", "posterior_mean_coef2"):
                rec[tag + "_" + k] = getattr(d, k)
    np.savez_compressed(os.path.join(HERE, "schedule_tables.npz"), **rec)
    print("wrote schedule_tables", len(rec))


def small_ops():
    """rot6d_to_rotmat and one ModulatedGraphConv straight from the reference's modules."""
    from utils.geometry import rot6d_to_rotmat
    from models.egohmr.modulated_gcn.modulated_gcn_conv import ModulatedGraphConv
    rng = np.random.default_rng(77)
    x = rng.normal(0, 1, (64, 144)).astype(np.float32)
    x[0, :6] = 0  # degenerate input exercises the eps clamp
    R32 = rot6d_to_rotmat(torch.from_numpy(x), "diffusion").numpy()
    R64 = rot6d_to_rotmat(torch.from_numpy(x).double(), "diffusion").numpy()
    adj = torch.from_numpy(synth.skeleton_adjacency())
    conv = ModulatedGraphConv(96, 128, adj)
    sdc = {k: rng.normal(0, 0.3, tuple(v.shape)).astype(np.float32) for k, v in conv.state_dict().items()}
    conv.load_state_dict({k: torch.from_numpy(v) for k, v in sdc.items()})
    xin = rng.normal(0, 1, (3, 24, 96)).astype(np.float32)
    with torch.no_grad():
        y32 = conv(torch.from_numpy(xin)).numpy()
        conv64 = conv.double()
        conv64.adj = adj.double()
        y64 = conv64(torch.from_numpy(xin).double()).numpy()
    np.savez_compressed(os.path.join(HERE, "small_ops.npz"), rot6d_x=x, rot6d_R32=R32, rot6d_R64=R64, gconv_x=xin,
                        gconv_y32=y32, gconv_y64=y64, **{"gconv_" + k: v for k, v in sdc.items()})
    print("wrote small_ops")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    # the stand-ins must be installed before importing reference modules
    smpl_model = synth.make_smpl_model(0)
    ref_standins.install(smpl_model, smpl_model["init_betas"])
    if len(sys.argv) > 1 and sys.argv[1] == "flags":   # the non-default model flags only (added later; small cases)
        run_case("ddim5_T50_hid256_maskall_f64", 50, "ddim5", 256, 2, 3, torch.float64, only_mask_img_cond=False)
        run_case("ddim5_T50_hid256_nofuse_f64", 50, "ddim5", 256, 2, 3, torch.float64, diffuse_fuse=False)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "nonlocal":
        run_case("ddim5_T50_hid256_nonlocal_f64", 50, "ddim5", 256, 2, 3, torch.float64, nonlocal_layer=True)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "procrustes":
        procrustes_case()
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cfg2":
        cfg2_case("cfg2_ddim5_8img_x10_hid1024_f64", torch.float64)
        cfg2_case("cfg2_ddim5_8img_x10_hid1024_f32", torch.float32)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "aa":
        angle_axis_case()
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "metrics":
        metrics_case()
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ddim_guided":   # ddim_sample_with_grad (:559-614) and eta != 0
        run_case("ddim5_guided_T50_hid256_f64", 50, "ddim5", 256, 2, 3, torch.float64, guided=True)
        run_case("ddim5_guided_T50_hid256_f32", 50, "ddim5", 256, 2, 3, torch.float32, guided=True)
        run_case("ddim5_eta05_T50_hid256_f64", 50, "ddim5", 256, 2, 3, torch.float64, eta=0.5)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cfg3":          # configs[2] at its real denoiser size: hid 1024, 4 blocks
        run_case("ddpm_guided_T100_hid1024_f64", 100, "", 1024, 4, 3, torch.float64, guided=True, trace_every=10)
        run_case("ddpm_guided_T100_hid1024_f32", 100, "", 1024, 4, 3, torch.float32, guided=True, trace_every=10)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cfg5":          # configs[4]'s chain length: T = 1000 DDPM steps
        run_case("ddpm_T1000_hid256_f64", 1000, "", 256, 2, 2, torch.float64, trace_every=100)
        run_case("ddpm_T1000_hid256_f32", 1000, "", 256, 2, 2, torch.float32, trace_every=100)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "loss":
        loss_case("val_losses_compute_loss_f32", torch.float32)
        raise SystemExit(0)
    schedule_tables()
    small_ops()
    run_case("ddim5_T50_hid1024_f32", 50, "ddim5", 1024, 4, 2, torch.float32)
    run_case("ddim5_T50_hid1024_f64", 50, "ddim5", 1024, 4, 2, torch.float64)
    run_case("ddpm_T50_hid256_f32", 50, "", 256, 2, 3, torch.float32)
    run_case("ddpm_T50_hid256_f64", 50, "", 256, 2, 3, torch.float64)
    guide_case("guide_grad_f32", torch.float32)
    guide_case("guide_grad_f64", torch.float64)
    run_case("ddpm_guided_T100_hid256_f32", 100, "", 256, 2, 3, torch.float32, guided=True)
    run_case("ddpm_guided_T100_hid256_f64", 100, "", 256, 2, 3, torch.float64, guided=True)
    run_case("ddim5_T50_hid256_maskall_f64", 50, "ddim5", 256, 2, 3, torch.float64, only_mask_img_cond=False)
    run_case("ddim5_T50_hid256_nofuse_f64", 50, "ddim5", 256, 2, 3, torch.float64, diffuse_fuse=False)
    loss_case("val_losses_compute_loss_f32", torch.float32)
    procrustes_case()
    run_case("ddim5_T50_hid256_nonlocal_f64", 50, "ddim5", 256, 2, 3, torch.float64, nonlocal_layer=True)

"""Times the BASELINE.json configurations other than the bench headline (cfg 2) on one GPU, one JSON line each:
  dropin   cfg 2 through the reference driver's own loop: num_samples sequential val_losses calls (test_egohmr.py:251-255)
  guided   cfg 3: DDPM-100 + collision-guided gradient, 32 images x 10 samples (per-body and batched collision interface)
  ddpm1000 cfg 5's per-GPU share: DDPM-1000, 64 images x 10 samples (512 x 10 over 8 GPUs)
  strong   cfg 4's per-GPU share: DDIM-5, 32 images x 10 samples (256 x 10 over 8 GPUs), graph replay
  realpts  the step-invariant encoders at the real dataset's shape: 64 images, 20 000-point scene clouds
           (dataloaders/egobody_dataset.py:213-225) instead of the 1 024 points of the benchmark configs
Usage: python tools/time_configs.py [dropin guided ddpm1000 strong realpts]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egohmr_b200 import synth  # noqa: E402
from egohmr_b200.diffusion.model_util import create_gaussian_diffusion  # noqa: E402
from egohmr_b200.testing import BatchedSyntheticCollision, SyntheticCollision, build_model, torch_batch  # noqa: E402

which = sys.argv[1:] or ["dropin", "guided", "guided_reference", "ddpm1000", "strong", "realpts", "metrics", "evalmetrics"]
dev = "cuda:0"
model, diffusion, sd, smpl_model, mean, std = build_model(1024, 4, T=50, respacing="ddim5")
mk = lambda T, r: create_gaussian_diffusion(num_diffusion_timesteps=T, timestep_respacing=r,
                                            body_rep_mean=torch.from_numpy(mean).to(dev), body_rep_std=torch.from_numpy(std).to(dev))


def timed(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) / reps * 1e3


def emit(name, bodies, ms, wall, **kw):
    print(json.dumps({"config": name, "bodies": bodies, "ms_per_pass": ms, "wall_ms_per_pass": wall,
                      "bodies_per_s": bodies / (ms * 1e-3), **kw}), flush=True)


if "dropin" in which:
    # a new batch object per pass, as a dataloader delivers (from the second batch on val_losses runs the samples ahead)
    batches = [torch_batch(synth.make_batch(100 + i, 64), dev) for i in range(2)]
    state = {"i": 0}

    def run():
        b = batches[state["i"] % 2]
        state["i"] += 1
        outs = []
        for _ in range(10):
            o = diffusion.val_losses(model=model, batch=b, shape=[64, 144], progress=False, clip_denoised=False,
                                     cur_epoch=0, timestep_respacing="ddim5", compute_loss=False)
            outs.append(o["pred_smpl_params"]["body_pose"].unsqueeze(1))
        return torch.cat(outs, dim=1)
    ms, wall = timed(run, 6, warm=3)
    emit("cfg2 drop-in loop: 10 sequential val_losses calls of 64 bodies (test_egohmr.py:251-255), DDIM-5, samples_ahead='auto'", 640, ms, wall)
    model.samples_ahead = 0
    ms, wall = timed(run, 6, warm=2)
    emit("cfg2 drop-in loop with samples_ahead=0 (one chain per call)", 640, ms, wall)
    model.samples_ahead = "auto"

if "guided" in which:
    d100 = mk(100, "")
    batch = torch_batch(synth.make_batch(101, 32), dev)
    for label, cm in (("per-body collision_loss calls (COAP's interface)", SyntheticCollision()),
                      ("collision_loss_batched (one call per guided step)", BatchedSyntheticCollision())):
        model.collision_model = cm

        def run():
            model._cond_key = None
            d100.sample_many(model, batch, 10, "", cond_fn_with_grad=True, cond_grad_weight=2.0)
        ms, wall = timed(run, 2)
        emit("cfg3 DDPM-100 + collision guidance (t <= 10), 32 images x 10 samples", 320, ms, wall, collision=label)

if "guided_reference" in which:
    # configs[2] on the UNMODIFIED reference (baseline/_ref through baseline/ref_harness.py), same GPU: the driver's loop, 10
    # sequential guided DDPM-100 chains of 32 images, per-body collision calls as the reference does them (one pass only)
    from baseline import ref_harness as rh
    if rh.reference_root()[0]:
        rmodel, rmean, rstd = rh.build_reference(1024, 4, device=dev)
        rsamp = rh.build_sampler(100, "", rmean, rstd, device=dev)
        rbatch = rh.make_driver_batch(101, 32, device=dev)
        rh.driver_loop(rmodel, rsamp, rh.make_driver_batch(101, 2, device=dev), 1, "", with_coap_grad=True, cond_grad_weight=2.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rh.driver_loop(rmodel, rsamp, rbatch, 10, "", with_coap_grad=True, cond_grad_weight=2.0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        emit("cfg3 on the UNMODIFIED reference (CUDA, torch defaults): DDPM-100 + collision guidance, 32 images x 10 sequential samples",
             320, dt * 1e3, dt * 1e3)
        del rmodel

if "ddpm1000" in which:
    d1000 = mk(1000, "")
    batch = torch_batch(synth.make_batch(102, 64), dev)

    def run():
        model._cond_key = None
        d1000.sample_many(model, batch, 10, "")
    ms, wall = timed(run, 1)
    emit("cfg5 per-GPU share: DDPM-1000, 64 images x 10 samples (eager loop)", 640, ms, wall)
    t0 = time.perf_counter()
    sampler = d1000.capture_sample_many(model, batch, 10, "")
    cap_s = time.perf_counter() - t0
    ms, wall = timed(lambda: sampler(batch), 2, warm=1)
    emit("cfg5 per-GPU share: DDPM-1000, 64 images x 10 samples (one CUDA graph of the 1000-step pass)", 640, ms, wall,
         capture_s=cap_s, launches_per_replay=sampler.launches_per_replay)
    del sampler

if "strong" in which:
    batch = torch_batch(synth.make_batch(103, 32), dev)
    sampler = diffusion.capture_sample_many(model, batch, 10, "ddim5")
    ms, wall = timed(lambda: sampler(batch), 20, warm=3)
    emit("cfg4 per-GPU share: DDIM-5, 32 images x 10 samples, CUDA-graph replay", 320, ms, wall)

    def run():
        model._cond_key = None
        diffusion.sample_many(model, batch, 10, "ddim5")
    ms, wall = timed(run, 20, warm=3)
    emit("cfg4 per-GPU share: DDIM-5, 32 images x 10 samples, eager launches", 320, ms, wall)

if "realpts" in which:
    import numpy as np
    b_np = synth.make_batch(104, 64, 20000)
    batch = torch_batch(b_np, dev)

    def run():
        model._cond_key = None
        model.prepare(batch, 10)
    ms, wall = timed(run, 5, warm=2)
    pts = batch["scene_pcd_verts_full"].float().contiguous()
    ms_pn, _ = timed(lambda: model.engine.pointnet_forward(pts), 5, warm=1)
    emit("encoders at the real data shape: 64 images + 64 x 20000-point clouds (prepare())", 64, ms, wall, pointnet_ms=ms_pn)

if "metrics" in which:
    from egohmr_b200.utils import pose_utils
    from egohmr_b200.utils.pytorch3d_chamfer_distance import chamfer_distance
    verts = torch.randn(640, 6890, 3, device=dev) * 0.4 + torch.tensor([0.0, 0.0, 3.0], device=dev)
    scene = torch_batch(synth.make_batch(105, 64, 20000), dev)["scene_pcd_verts_full"].float().contiguous()
    idx = torch.arange(64, device=dev).repeat_interleave(10)
    ms, wall = timed(lambda: chamfer_distance(verts, scene, y_index=idx, compute_y=False), 3, warm=1)
    emit("contact score: 640 bodies x 6890 vertices vs 20000-point scene clouds, vertex -> scene 1-NN", 640, ms, wall)
    ms, wall = timed(lambda: chamfer_distance(verts, scene, y_index=idx), 2, warm=0)
    emit("chamfer_distance both ways (as the reference's driver calls it)", 640, ms, wall)
    a, b = torch.randn(640, 24, 3, device=dev), torch.randn(640, 24, 3, device=dev)
    ms, wall = timed(lambda: pose_utils.reconstruction_error(a, b, avg_joint=False), 20, warm=2)
    emit("PA-MPJPE: Procrustes alignment of 640 x 24 joints", 640, ms, wall)

if "evalmetrics" in which:
    from egohmr_b200.utils.eval_metrics import evaluate_batch
    g = torch.Generator(device="cpu").manual_seed(0)
    bs, S = 64, 10
    pj = (torch.randn(bs, S, 24, 3, generator=g) * 0.3).to(dev)
    pv = (torch.randn(bs, S, 6890, 3, generator=g) * 0.3).to(dev)
    gj = (torch.randn(bs, 24, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 3.0])).to(dev)
    gv = (torch.randn(bs, 6890, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 3.0])).to(dev)
    tr = torch.tensor([[0.0, 0.0, 3.0]]).repeat(bs, 1).to(dev)
    f, cx, cy = torch.full((bs,), 1500.0, device=dev), torch.full((bs,), 960.0, device=dev), torch.full((bs,), 540.0, device=dev)
    ms, wall = timed(lambda: evaluate_batch(pj, pv, tr, gj, gv, f, cx, cy, engine=model.engine), 20, warm=2)
    emit("evaluation metrics of one batch (test_egohmr.py:373-494): 64 images x 10 samples, G-MPJPE / MPJPE / PA-MPJPE / V2V / std / APD",
         640, ms, wall)

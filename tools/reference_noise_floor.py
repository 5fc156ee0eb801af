"""How reproducible is the UNMODIFIED reference itself?  north_star asks for per-vertex error < 1e-3 mm "vs reference"; this
tool measures, on the GPU box, how far the reference's own fp32 runs are from its float64 run and from each other:

  ref CPU fp32 | ref CUDA fp32 (strict: cudnn.allow_tf32=False) | ref CUDA torch defaults (TF32 convolutions)
  | this repo (sm_100a kernels)          — each against the reference in float64 on the CPU, same inputs, same noise.

DDIM-5 of T=50, hid 1024 / 4 blocks, 8 images x 1 sample.  One JSON object on stdout (-> profiles/r02_reference_noise_floor.json)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_harness as rh  # noqa: E402
from egohmr_b200 import synth  # noqa: E402

N_IMG, SEED = 8, 3
noise = synth.make_noise(SEED, 1, N_IMG, 5)[0]


class Feed:
    def __init__(self, dtype, device):
        self.n = [torch.from_numpy(x).to(dtype).to(device) for x in noise]
        self.i = 0

    def randn(self, *a, **k):
        self.i += 1
        return self.n[self.i - 1].clone()

    def randn_like(self, x, **k):
        return self.randn()


def run_reference(dtype, device):
    model, mean, std = rh.build_reference(1024, 4, dtype, device=device)
    samp = rh.build_sampler(50, "ddim5", mean, std, dtype, device)
    batch = rh.to_torch(synth.make_batch(SEED, N_IMG), dtype, device)
    feed = Feed(dtype, device)
    old = torch.randn, torch.randn_like
    torch.randn, torch.randn_like = feed.randn, feed.randn_like
    try:
        with torch.no_grad():
            out = samp.val_losses(model=model, batch=batch, shape=[N_IMG, 144], progress=False, clip_denoised=False,
                                  cur_epoch=0, timestep_respacing="ddim5", compute_loss=False)
    finally:
        torch.randn, torch.randn_like = old
    return out["pred_x_start"].double().cpu().numpy(), out["pred_vertices"].double().cpu().numpy()


def run_ours():
    from egohmr_b200.testing import build_model, torch_batch
    model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5")
    batch = torch_batch(synth.make_batch(SEED, N_IMG), "cuda:0")
    out = diffusion.sample_many(model, batch, 1, "ddim5", noise=torch.from_numpy(noise).cuda())
    return out["pred_x_start"].double().cpu().numpy(), out["pred_vertices"].double().cpu().numpy()


torch.set_num_threads(os.cpu_count() or 8)
x64, v64 = run_reference(torch.float64, "cpu")
res = {"what": __doc__.split("\n\n")[0], "n_img": N_IMG, "unit": {"x0": "normalised rot6d units", "vertices": "mm"}}


def add(name, xv):
    x, v = xv
    res[name] = {"max_abs_x0_err_vs_ref_f64": float(np.abs(x - x64).max()),
                 "max_vertex_err_vs_ref_f64_mm": float(np.abs(v - v64).max() * 1e3)}
    return x, v


cpu32 = add("reference_cpu_fp32", run_reference(torch.float32, "cpu"))
torch.backends.cudnn.allow_tf32 = False
gpu32 = add("reference_cuda_fp32_strict", run_reference(torch.float32, "cuda:0"))
torch.backends.cudnn.allow_tf32 = True
gpu_tf32 = add("reference_cuda_torch_defaults_tf32_convs", run_reference(torch.float32, "cuda:0"))
ours = add("egohmr_b200", run_ours())
res["reference_cpu_fp32_vs_reference_cuda_fp32_strict"] = {
    "max_abs_x0_diff": float(np.abs(cpu32[0] - gpu32[0]).max()), "max_vertex_diff_mm": float(np.abs(cpu32[1] - gpu32[1]).max() * 1e3)}
res["egohmr_b200_vs_reference_cuda_fp32_strict"] = {
    "max_abs_x0_diff": float(np.abs(ours[0] - gpu32[0]).max()), "max_vertex_diff_mm": float(np.abs(ours[1] - gpu32[1]).max() * 1e3)}
print(json.dumps(res, indent=1))

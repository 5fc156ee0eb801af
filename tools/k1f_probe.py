"""A/B of the hidden layers of a reverse step: one persistent launch with per-(layer, row group) completion counters
(gcn_umma_fused.cu, default) vs one launch per layer (gcn_umma_t.cu).  Same bits expected; reverse-step and graphed-pass time."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egohmr_b200 import synth  # noqa: E402
from egohmr_b200.testing import build_model, torch_batch  # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5")
batch = torch_batch(synth.make_batch(100, n_img), "cuda:0")
B = n_img * 10
noise = torch.from_numpy(synth.make_noise(0, 1, B, diffusion.num_timesteps)[0]).cuda()
eng = model.engine
x = torch.randn(B, 144, device="cuda")
xp, x0 = torch.empty_like(x), torch.empty_like(x)
outs = {}
for fused in (0, 1):
    eng.set_k1_fused(fused)
    out = diffusion.sample_many(model, batch, 10, "ddim5", noise=noise)
    torch.cuda.synchronize()
    outs[fused] = out["pred_x_start"].clone()
    print(json.dumps({"fused": fused, "sampled": True, "overflow": bool(eng.check_overflow())}), flush=True)
d = (outs[0] - outs[1]).abs().max().item()
print(json.dumps({"max|x0(per-layer) - x0(fused)|": d, "finite": bool(torch.isfinite(outs[1]).all())}), flush=True)
for rep in range(3):
    for fused in (0, 1):
        eng.set_k1_fused(fused)
        for _ in range(5):
            eng.denoise_step(2, x, None, None, xp, x0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            eng.denoise_step(2, x, None, None, xp, x0)
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / 40
        torch.manual_seed(0)
        sampler = diffusion.capture_sample_many(model, batch, 10, "ddim5")
        for _ in range(3):
            sampler(batch)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            sampler(batch)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"fused": fused, "reverse_step_ms": step_ms, "graphed_pass_ms": e0.elapsed_time(e1) / 20}), flush=True)
        del sampler
eng.set_k1_fused(1)

"""A/B of programmatic dependent launch on the hidden GCN layers: device time of the 8 K1 launches of one reverse step,
issued back to back (as in a pass), with and without PDL; and the whole graphed pass."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egohmr_b200 import synth  # noqa: E402
from egohmr_b200.testing import build_model, torch_batch  # noqa: E402

model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5")
batch = torch_batch(synth.make_batch(100, 64), "cuda:0")
out = diffusion.sample_many(model, batch, 10, "ddim5")
ref = out["pred_x_start"].clone()
eng = model.engine
x = torch.randn(640, 144, device="cuda")
xp, x0 = torch.empty_like(x), torch.empty_like(x)


def steps(n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.denoise_step(2, x, None, None, xp, x0)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rep in range(2):
    for on in (0, 1):
        eng.set_pdl(on)
        steps(5)
        ms = steps(40)
        torch.manual_seed(0)
        sampler = diffusion.capture_sample_many(model, batch, 10, "ddim5")
        for _ in range(3):
            sampler(batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            sampler(batch)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"pdl": on, "eager_reverse_step_ms": ms, "graphed_pass_ms": e0.elapsed_time(e1) / 20}), flush=True)
        del sampler
eng.set_pdl(1)
noise_free = diffusion.sample_many(model, batch, 10, "ddim5")
print(json.dumps({"finite": bool(torch.isfinite(noise_free["pred_x_start"]).all())}))

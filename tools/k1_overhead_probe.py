"""Per-launch fixed cost of the GCN layer kernel: time back-to-back launches at shrinking problem sizes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egohmr_b200 import synth
from egohmr_b200.testing import build_model, torch_batch
model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5")
for n_img, S in ((64, 10), (16, 10), (4, 10), (1, 5), (1, 1)):
    batch = torch_batch(synth.make_batch(100, n_img), "cuda:0")
    model._cond_key = None
    diffusion.sample_many(model, batch, S, "ddim5")
    x = torch.randn(n_img * S, 144, device="cuda")
    ms = [model.engine.time_stage(l, 2, x, 200) for l in (1, 2)]
    slots = 2 * n_img * S
    units = ((slots + 4) // 5 + 1) // 2 * 8      # 256-row pair tiles x 8 n-tiles of 256 (h0 | h1 interleaved)
    print(f"bodies {n_img * S:4d}: {units:4d} units on 74 CTA pairs -> {ms[0] * 1e3:7.1f} / {ms[1] * 1e3:7.1f} us per launch")

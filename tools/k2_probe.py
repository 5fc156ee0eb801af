"""A/B of the input graph-convolution layer (K2): fp32 FFMA kernel vs joint mix on tcgen05.  Device time of the kernel
alone, of a whole reverse step, and the difference in the sampled x0 (same noise) between the two and vs float64."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egohmr_b200 import synth  # noqa: E402
from egohmr_b200.testing import build_model, torch_batch  # noqa: E402

model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5")
batch = torch_batch(synth.make_batch(100, 64), "cuda:0")
noise = torch.from_numpy(synth.make_noise(0, 1, 640, diffusion.num_timesteps)[0]).cuda()
eng = model.engine
x = torch.randn(640, 144, device="cuda")
xp, x0 = torch.empty_like(x), torch.empty_like(x)
outs = {}
for rep in range(3):
    for mode in (0, 1):
        eng.set_input_mode(mode)
        out = diffusion.sample_many(model, batch, 10, "ddim5", noise=noise)
        torch.cuda.synchronize()
        outs[mode] = out["pred_x_start"].clone()
        k2 = min(eng.time_stage(0, 2, x, 100) for _ in range(3))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            eng.denoise_step(2, x, None, None, xp, x0)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"input_umma": mode, "k2_ms": k2, "reverse_step_ms": e0.elapsed_time(e1) / 40,
                          "overflow": bool(eng.check_overflow())}), flush=True)
d = (outs[0] - outs[1]).abs().max().item()
print(json.dumps({"max|x0(ffma) - x0(umma)|": d, "finite": bool(torch.isfinite(outs[1]).all())}))

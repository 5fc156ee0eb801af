"""A/B of the hidden graph-convolution layer (K1): row-major product (gemm_mode 3: activations on the M side, 5 slots + 8 pad
rows per 128-row tile) vs transposed product (gemm_mode 0, the product path: weights on the M side, N = 240 = 10 slots, no pad rows).
Per-layer device time, a whole reverse step, and the sampled x0 (same noise) of the two."""
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
for rep in range(2):
    for mode in (3, 0):
        eng.set_gemm_mode(mode)
        out = diffusion.sample_many(model, batch, 10, "ddim5", noise=noise)
        torch.cuda.synchronize()
        outs[mode] = out["pred_x_start"].clone()
        layers = [min(eng.time_stage(l, 2, x, 20) for _ in range(3)) for l in range(1, 9)]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            eng.denoise_step(2, x, None, None, xp, x0)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"gemm_mode": mode, "layer_ms": [round(v, 4) for v in layers], "k1_mean_ms": sum(layers) / 8,
                          "reverse_step_ms": e0.elapsed_time(e1) / 40, "overflow": bool(eng.check_overflow())}), flush=True)
for rep in range(3):
    for mode in (3, 0):
        eng.set_gemm_mode(mode)
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
        print(json.dumps({"gemm_mode": mode, "graphed_pass_ms": e0.elapsed_time(e1) / 20}), flush=True)
        del sampler
eng.set_gemm_mode(0)
d = (outs[0] - outs[3]).abs().max().item()
print(json.dumps({"max|x0(row-major) - x0(transposed)|": d, "finite": bool(torch.isfinite(outs[3]).all())}))

#!/bin/bash
# A/B of the whole graphed sampling pass (cfg 2: 64 images x 10 samples, DDIM-5) between the in-tree library and
# build/libegohmr_b200_prev.so (a previous build), alternating on one box.  Prints ms per pass and a checksum of x0.
cp egohmr_b200/lib/libegohmr_b200.so /tmp/lib_new.so
for rep in 1 2 3; do
  for which in prev new; do
    if [ $which = prev ]; then cp build/libegohmr_b200_prev.so egohmr_b200/lib/libegohmr_b200.so; else cp /tmp/lib_new.so egohmr_b200/lib/libegohmr_b200.so; fi
    echo -n "$which: "; timeout 200 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from egohmr_b200 import synth
from egohmr_b200.testing import build_model, torch_batch
model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5")
batch = torch_batch(synth.make_batch(100, 64), "cuda:0")
noise = torch.from_numpy(synth.make_noise(0, 1, 640, diffusion.num_timesteps)[0]).cuda()
out = diffusion.sample_many(model, batch, 10, "ddim5", noise=noise)
chk = out["pred_x_start"].double().sum().item()
torch.manual_seed(0)
sampler = diffusion.capture_sample_many(model, batch, 10, "ddim5")
for _ in range(5): sampler(batch)
torch.cuda.synchronize()
best = 1e9
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): sampler(batch)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 10)
print(f"{best:.4f} ms per pass  x0 checksum {chk:.10e}")
PY
  done
done
cp /tmp/lib_new.so egohmr_b200/lib/libegohmr_b200.so

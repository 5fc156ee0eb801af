#!/bin/bash
# A/B of the native ResNet-50 between the in-tree library and build/libegohmr_b200_prev.so (the previous build), alternating
# on one box.  Prints the native time of each run and the feature checksum.
cp egohmr_b200/lib/libegohmr_b200.so /tmp/lib_new.so
for rep in 1 2 3; do
  for which in prev new; do
    if [ $which = prev ]; then cp build/libegohmr_b200_prev.so egohmr_b200/lib/libegohmr_b200.so; else cp /tmp/lib_new.so egohmr_b200/lib/libegohmr_b200.so; fi
    echo -n "$which: "; timeout 120 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from egohmr_b200 import synth
from egohmr_b200.testing import build_model, torch_batch
model, *_ = build_model(1024, 4, T=50, respacing="ddim5")
model._sync_engine()
img = torch_batch(synth.make_batch(100, 64), "cuda:0")["img"].contiguous()
for _ in range(5): f = model.engine.resnet_forward(img)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f = model.engine.resnet_forward(img)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 10)
print(f"{best:.4f} ms  checksum {f.double().sum().item():.10e} absmax {f.abs().max().item():.8e}")
PY
  done
done
cp /tmp/lib_new.so egohmr_b200/lib/libegohmr_b200.so

#!/bin/bash
# Times the graphed sampling pass (cfg 2, sustained: 3 x 15 replays, median) with each of the given library builds swapped into
# egohmr_b200/lib/ in turn (plus the in-tree one as "tree", with the fused hidden layers on and off), two rounds.
# Usage: bash tools/ab_libs.sh build/variants/lib_a.so ...
cp egohmr_b200/lib/libegohmr_b200.so /tmp/lib_tree.so
for rep in 1 2; do
  for lib in /tmp/lib_tree.so:0 /tmp/lib_tree.so:1 $(for l in "$@"; do echo $l:1; done); do
    cp ${lib%%:*} egohmr_b200/lib/libegohmr_b200.so
    echo -n "$(basename ${lib%%:*}) fused=${lib##*:}: "; EHB_FUSED=${lib##*:} timeout 120 python - <<'PY'
import os, sys, torch, statistics
sys.path.insert(0, os.getcwd())
from egohmr_b200 import synth
from egohmr_b200.testing import build_model, torch_batch
model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5")
batch = torch_batch(synth.make_batch(100, 64), "cuda:0")
diffusion.sample_many(model, batch, 10, "ddim5")
model.engine.set_k1_fused(int(os.environ["EHB_FUSED"]))
torch.manual_seed(0)
sampler = diffusion.capture_sample_many(model, batch, 10, "ddim5")
for _ in range(10): sampler(batch)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(15): sampler(batch)
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) / 15)
print(f"graphed pass {statistics.median(ts):.3f} ms  (runs {[round(t, 3) for t in ts]})")
PY
  done
done
cp /tmp/lib_tree.so egohmr_b200/lib/libegohmr_b200.so

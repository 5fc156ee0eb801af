"""One native ResNet-50 forward (K9) for 64 images bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
Optional argv: kc (accumulation chunk, default = library default)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egohmr_b200 import synth  # noqa: E402
from egohmr_b200.testing import build_model, torch_batch  # noqa: E402

model, *_ = build_model(1024, 4, T=50, respacing="ddim5")
model._sync_engine()
if len(sys.argv) > 1:
    model.engine.set_conv_kc(int(sys.argv[1]))
img = torch_batch(synth.make_batch(100, 64), "cuda:0")["img"].float().contiguous()
for _ in range(3):
    model.engine.resnet_forward(img)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.engine.resnet_forward(img)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

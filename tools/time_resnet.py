"""Times the three ResNet-50 feature providers on 64 images: native tcgen05 convolution GEMMs (K9), cuDNN with TF32
allowed (torch default) and cuDNN strict fp32; prints per-kernel-family shares of the native path when run under ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egohmr_b200 import synth  # noqa: E402
from egohmr_b200.testing import build_model, torch_batch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model, *_ = build_model(1024, 4, T=50, respacing="ddim5")
model._sync_engine()
img = torch_batch(synth.make_batch(100, n), "cuda:0")["img"].contiguous()


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


print(f"native tcgen05 (fp32-class): {timed(lambda: model.engine.resnet_forward(img)):.3f} ms for {n} images")
torch.backends.cudnn.allow_tf32 = True
print(f"cuDNN, TF32 allowed        : {timed(lambda: model._fast_backbone(img)):.3f} ms")
torch.backends.cudnn.allow_tf32 = False
print(f"cuDNN, strict fp32         : {timed(lambda: model._fast_backbone(img), 3):.3f} ms")
if os.environ.get("EHB_PROFILE"):
    torch.cuda.profiler.start()
    model.engine.resnet_forward(img)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()

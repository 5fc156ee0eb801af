"""CUDA-event timing of the two feature providers, with and without cudnn.benchmark."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egohmr_b200 import synth  # noqa: E402
from egohmr_b200.testing import build_model, torch_batch  # noqa: E402

model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5")
model._sync_engine()
batch = torch_batch(synth.make_batch(100, 64), "cuda:0")
pts = batch["scene_pcd_verts_full"] - batch["smpl_params"]["transl"].unsqueeze(1)


def timed(fn, it=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    print(f"cudnn.benchmark={bench}: resnet {timed(lambda: model._fast_backbone(batch['img'])):.3f} ms, "
          f"pointnet {timed(lambda: model._fast_scene_enc(pts)):.3f} ms, plain resnet {timed(lambda: model.backbone(batch['img'])):.3f} ms")
torch.backends.cudnn.benchmark = True
g = torch.cuda.CUDAGraph()
static_img = batch["img"].clone()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        model._fast_backbone(static_img)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = model._fast_backbone(static_img)
print(f"resnet under CUDA graph: {timed(g.replay):.3f} ms")

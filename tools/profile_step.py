"""One sampling pass of the bench workload bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egohmr_b200 import synth  # noqa: E402
from egohmr_b200.testing import build_model, torch_batch  # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5")
model._sync_engine()
model.engine.set_gemm_mode(int(os.environ.get("EHB_GEMM_MODE", "0")))
batch = torch_batch(synth.make_batch(100, n_img), "cuda:0")
for _ in range(3):
    model._cond_key = None
    diffusion.sample_many(model, batch, S, "ddim5")
torch.cuda.synchronize()
torch.cuda.profiler.start()
model._cond_key = None
diffusion.sample_many(model, batch, S, "ddim5")
torch.cuda.synchronize()
torch.cuda.profiler.stop()

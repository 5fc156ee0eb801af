"""How does the tensor core's fp32 accumulation (tcgen05.mma kind::f16, fp32 TMEM accumulator) round?

(1) GEMM probe through `ehb_debug_gemm_hl`: operands exactly representable in fp16 (lo = 0), so every product is exact and the
    only error is the accumulation; compared with the exact float64 sum, for contraction lengths K and chunk lengths kc
    (k-blocks of 64 chained into one TMEM accumulation before the epilogue's fp32 round-to-nearest sum takes over).
(2) ResNet-50 feature error vs float64 and time for 64 images as a function of kc.
Writes one JSON line per measurement to stdout."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egohmr_b200 import synth  # noqa: E402
from egohmr_b200.testing import build_model, torch_batch  # noqa: E402


def gemm_probe(eng):
    rng = np.random.default_rng(0)
    M, N = 256, 128
    for dist in ("positive", "signed"):
        for K in (64, 256, 1024, 4096):
            a = rng.uniform(0.5, 1.0, (M, K)).astype(np.float16).astype(np.float32)
            w = rng.uniform(0.5, 1.0, (N, K)).astype(np.float16).astype(np.float32)
            if dist == "signed":
                a *= rng.choice([-1.0, 1.0], a.shape).astype(np.float32)
            exact = a.astype(np.float64) @ w.astype(np.float64).T
            rn32 = exact.astype(np.float32).astype(np.float64)
            for kc in (0, 1, 2, 4, 8, 16):
                if kc and kc >= K // 64:
                    continue
                got = eng.debug_gemm_hl(a, w, 1.0, 1.0, kc).astype(np.float64)
                scale = np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64).T   # sum of |terms|
                rel = (got - exact) / scale
                print(json.dumps({"probe": "gemm", "dist": dist, "K": K, "kc": kc, "mma_chain": 3 * (K if not kc else kc * 64) // 16,
                                  "mean_rel_err": float(rel.mean()), "max_abs_rel_err": float(np.abs(rel).max()),
                                  "rn_fp32_max_rel": float(np.abs((rn32 - exact) / scale).max())}), flush=True)


def resnet_probe():
    model, diffusion, sd, smpl_model, mean, std = build_model(1024, 4, T=50, respacing="ddim5")
    model._sync_engine()
    eng = model.engine
    if len(sys.argv) == 1:
        gemm_probe(eng)
    img3 = torch_batch(synth.make_batch(4, 3), "cuda:0")["img"]
    img64 = torch_batch(synth.make_batch(100, 64), "cuda:0")["img"]
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        ref64 = model.backbone.double()(img3.double())
        model.backbone.float()
        ref32 = model.backbone(img3)
    print(json.dumps({"probe": "resnet", "impl": "torch fp32 (cuDNN strict)", "max_err_vs_f64": float((ref32.double() - ref64).abs().max()),
                      "max_feat": float(ref64.abs().max())}), flush=True)
    for kc in (int(a) for a in sys.argv[1:]) if len(sys.argv) > 1 else (0, 1, 2, 3, 4, 6, 9, 18):
        eng.set_conv_kc(kc)
        got = eng.resnet_forward(img3.contiguous())
        err = float((got.double() - ref64).abs().max())
        rel_mean = float(((got.double() - ref64) / ref64.abs().clamp_min(1e-3)).mean())
        for _ in range(3):
            eng.resnet_forward(img64.contiguous())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.resnet_forward(img64.contiguous())
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"probe": "resnet", "kc": kc, "max_err_vs_f64": err, "mean_signed_rel_err": rel_mean,
                          "ms_per_64_images": e0.elapsed_time(e1) / 10, "overflow": bool(eng.check_overflow())}), flush=True)
    eng.set_conv_kc(4)


if __name__ == "__main__":
    resnet_probe()

#!/bin/bash
# Per-warp timeline of one convolution GEMM launch of the native ResNet-50.  Build here:
#   nvcc -DEHB_CONV_TRACE ... -c egohmr_b200/csrc/conv_umma.cu -o build/conv_umma_trace.o ; link with the other objects into
#   build/libegohmr_b200_trace.so (see the end of this file's header in git history / DESIGN.md)
# Run on the GPU box: bash tools/conv_trace.sh "<launch ordinals>"   (ordinal = n-th conv_gemm launch of resnet_forward, from 0)
cp egohmr_b200/lib/libegohmr_b200.so /tmp/lib_product.so
cp build/libegohmr_b200_trace.so egohmr_b200/lib/libegohmr_b200.so
timeout 200 python - "$@" <<'PY'
import ctypes as C, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from egohmr_b200 import synth, _lib
from egohmr_b200.testing import build_model, torch_batch
model, *_ = build_model(1024, 4, T=50, respacing="ddim5")
model._sync_engine()
lib = _lib.load()
img = torch_batch(synth.make_batch(100, 64), "cuda:0")["img"].contiguous()
for _ in range(3): model.engine.resnet_forward(img)
torch.cuda.synchronize()
U, S = 12, 40
names = {0: "unit start", 1: "before final-accumulator wait", 2: "accumulator ready"}
for ordinal in [int(x) for x in (sys.argv[1].split() if len(sys.argv) > 1 else ["6"])]:
    assert lib.ehb_conv_trace_arm(ordinal, 1 << 30) == 0
    model.engine.resnet_forward(img)
    torch.cuda.synchronize()
    buf = np.zeros((U, S), np.int64)
    assert lib.ehb_conv_trace_read(buf.ctypes.data_as(C.POINTER(C.c_longlong))) == 0
    t0 = buf[0, 0]
    print(f"=== conv_gemm launch {ordinal}: cycles relative to the first epilogue stamp of CTA 0 (warp 0, lane 0 unless noted)")
    for u in range(U):
        r = buf[u]
        if r[0] == 0: break
        rel = lambda s: int(r[s] - t0) if r[s] else None
        chunks = []
        for ch in range(4):
            if r[4 + 4 * ch]: chunks.append({"ld": rel(4 + 4 * ch), "identity": rel(5 + 4 * ch), "math": rel(6 + 4 * ch), "stored": rel(7 + 4 * ch)})
        print(json.dumps({"unit": u, "start": rel(0), "wait_acc": rel(1), "acc_ready": rel(2), "chunks": chunks,
                          "mma": {"wait_tempty": rel(24), "tempty_ok": rel(25), "operands_landed": rel(26), "committed": rel(27)},
                          "tma": {"wait_empty": rel(30), "empty_ok": rel(31), "issued": rel(32)}}))
PY
cp /tmp/lib_product.so egohmr_b200/lib/libegohmr_b200.so

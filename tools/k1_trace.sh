#!/bin/bash
# Timeline of the hidden-layer kernel (K1, transposed product): build/libegohmr_b200_trace.so (tools/build_trace_lib.sh gcn_umma_t EHB_K1_TRACE) = the library with gcn_umma_t.cu
# compiled -DEHB_K1_TRACE.  Per unit of CTA 0 of one launch: where the MMA, TMA, tcgen05.ld and mix warps wait.
cp egohmr_b200/lib/libegohmr_b200.so /tmp/lib_product.so
cp build/libegohmr_b200_trace.so egohmr_b200/lib/libegohmr_b200.so
timeout 200 python - <<'PY'
import ctypes as C, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from egohmr_b200 import synth, _lib
from egohmr_b200.testing import build_model, torch_batch
model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5")
batch = torch_batch(synth.make_batch(100, 64), "cuda:0")
diffusion.sample_many(model, batch, 10, "ddim5")
torch.cuda.synchronize()
lib = _lib.load()
eng = model.engine
x = torch.randn(640, 144, device="cuda")
for layer in (1, 2):
    for _ in range(3): eng.time_stage(layer, 2, x, 5)
    assert lib.ehb_k1_trace_arm(1) == 0
    ms = eng.time_stage(layer, 2, x, 1)
    torch.cuda.synchronize()
    assert lib.ehb_k1_trace_arm(0) == 0
    buf = np.zeros((16, 16), np.int64)
    assert lib.ehb_k1_trace_read(buf.ctypes.data_as(C.POINTER(C.c_longlong))) == 0
    pairs = np.zeros((128, 2), np.uint64)
    assert lib.ehb_k1_trace_pairs(pairs.ctypes.data_as(C.POINTER(C.c_ulonglong))) == 0
    pt = pairs[:74].astype(np.int64)
    start0 = pt[:, 0].min()
    last = (pt[:, 1] - start0) / 1e3     # us from the first pair's start to each pair's last unit being issued
    n14 = last[:62]; n13 = last[62:]       # pairs 0-61 run 14 units, 62-73 run 13 (1024 = 13 * 74 + 62)
    print(f"=== hidden layer {layer}: per CTA pair, us from the earliest start to the last unit issued: pairs with 14 units "
          f"min {n14.min():.1f} median {np.median(n14):.1f} max {n14.max():.1f}; pairs with 13 units min {n13.min():.1f} "
          f"median {np.median(n13):.1f} max {n13.max():.1f}; start spread {(pt[:, 0].max() - start0) / 1e3:.1f} us")
    print("    per pair:", " ".join(f"{v:.0f}" for v in last))
    t0 = buf[0, 0]
    print(f"=== hidden layer {layer} (one launch, {ms:.3f} ms): cycles; unit = 192 MMAs of 256x240x16")
    for u in range(16):
        r = buf[u]
        if r[0] == 0: break
        print(json.dumps({"unit": u, "mma_start": int(r[0] - t0), "mma_wait_acc": int(r[1]), "mma_wait_operands": int(r[2]),
                          "mma_issued": int(r[3] - t0), "ld_wait_acc": int(r[4]), "acc_complete": int(r[5] - t0),
                          "staged": int(r[6] - t0), "stored": int(r[7] - t0), "tma_wait_stages": int(r[8]), "tma_issued": int(r[9] - t0)}))
PY
cp /tmp/lib_product.so egohmr_b200/lib/libegohmr_b200.so

"""Where does the end-to-end leg of bench.py lose its ~3 % against the device-resident leg?  Same graph, same batch; variants of
the per-step work around the replay, each timed with CUDA events on the compute stream AND the host clock."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from egohmr_b200 import sharding, synth  # noqa: E402
from egohmr_b200.testing import build_model, torch_batch  # noqa: E402

dev = torch.device("cuda", 0)
model, diffusion, *_ = build_model(bench.HID, bench.N_BLOCKS, T=bench.T, respacing=bench.RESPACING, device=str(dev))
batch_np = synth.make_batch(100, 64, bench.N_PTS)
batch_dev = torch_batch(batch_np, dev)
host, _ = bench._pinned_host_batch(batch_np)
S, B = 10, 640
for _ in range(3):
    model.invalidate()
    diffusion.sample_many(model, batch_dev, S, bench.RESPACING)
sampler = diffusion.capture_sample_many(model, batch_dev, S, bench.RESPACING)
res_host = {"R": torch.empty(B, 24, 3, 3).pin_memory(), "betas": torch.empty(B, 10).pin_memory(),
            "joints": torch.empty(B, 45, 3).pin_memory()}


def d2h(out):
    res_host["R"][:, :1].copy_(out["pred_smpl_params"]["global_orient"], non_blocking=True)
    res_host["R"][:, 1:].copy_(out["pred_smpl_params"]["body_pose"], non_blocking=True)
    res_host["betas"].copy_(out["pred_smpl_params"]["betas"], non_blocking=True)
    res_host["joints"].copy_(out["pred_keypoints_3d"], non_blocking=True)


res2 = {"params": torch.empty(B, 226).pin_memory(), "joints": torch.empty(B, 45, 3).pin_memory()}


def v_full_contig(i, n):
    out = sampler(staged=True)
    if i + 1 < n:
        sampler.stage(host)
    res2["params"].copy_(sharding.pack_results(out), non_blocking=True)
    res2["joints"].copy_(out["pred_keypoints_3d"].contiguous(), non_blocking=True)


def v_replay(i, n):
    sampler(batch_dev)


def v_replay_pack(i, n):
    sharding.pack_results(sampler(batch_dev))


def v_staged(i, n):
    out = sampler(staged=True)
    if i + 1 < n:
        sampler.stage(host)


def v_staged_d2h(i, n):
    out = sampler(staged=True)
    if i + 1 < n:
        sampler.stage(host)
    d2h(out)


def v_full(i, n):
    out = sampler(staged=True)
    if i + 1 < n:
        sampler.stage(host)
    sharding.pack_results(out)
    d2h(out)


def v_replay_d2h(i, n):
    d2h(sampler(batch_dev))


n = 20
for rep in range(2):
    for name, fn in (("replay", v_replay), ("replay+pack", v_replay_pack), ("replay+d2h", v_replay_d2h), ("staged", v_staged),
                     ("staged+d2h", v_staged_d2h), ("full e2e", v_full), ("full e2e, contiguous d2h", v_full_contig)):
        sampler.stage(host)
        fn(0, 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        sampler.stage(host)
        e0.record()
        for i in range(n):
            fn(i, n)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n * 1e3
        print(json.dumps({"variant": name, "event_ms_per_step": e0.elapsed_time(e1) / n, "wall_ms_per_step": wall}), flush=True)

#!/usr/bin/env python
"""Benchmark of the EgoHMR diffusion-sampling hot path (BASELINE.json: sampled bodies/sec, DDIM-5, batch 64 x 10 samples).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port, reference dataflow) on host cores

A "step" is one pass of the hot path over one synthetic batch: everything `diffusion.val_losses(...)` does for 64
images x 10 samples (step-invariant encoders once, 5 reverse-diffusion steps of the 10-layer GCN evaluated twice,
rot6d + SMPL for the final x0).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_IMG, N_SAMPLES, T, RESPACING, N_PTS = 64, 10, 50, "ddim5", 1024
HID, N_BLOCKS = 1024, 4
METRIC = "sampled bodies/sec (batch x num_samples), DDIM-5"
WORKLOAD = "configs[1]: DDIM-5 (T=50), batch=64 images x num_samples=10 = 640 bodies per step, 224x224 img + 1024 scene pts"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-img", type=int, default=N_IMG)
    ap.add_argument("--num-samples", type=int, default=N_SAMPLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-img", type=int, default=4, help="images in the bounded CPU-baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(n_img, n_samples, repeats=1):
    """The oracle port run with the REFERENCE's dataflow (hoist=False: encoders + SMPL recomputed on every step, one
    sequential chain per sample, test_egohmr.py:251-255), fp32, all host threads numpy/torch give it."""
    from egohmr_b200 import synth
    from oracle import egohmr as o_egohmr, schedule as o_schedule
    smpl_model = synth.make_smpl_model(0)
    sd = synth.make_state_dict(0, hid=HID, n_blocks=N_BLOCKS, init_betas=smpl_model["init_betas"])
    mean, std = synth.body_rep_stats(0)
    batch = synth.make_batch(0, n_img, N_PTS)
    sch = o_schedule.Schedule(T, RESPACING)
    adj = synth.skeleton_adjacency()
    noise = synth.make_noise(0, n_samples, n_img, sch.num_timesteps)
    o_egohmr.sample(sd, adj, N_BLOCKS, smpl_model, synth.make_batch(1, 1, N_PTS), sch, noise[0][:, :1], mean, std, "ddim",
                    dtype=np.float32, hoist=False)  # warm-up (thread pools, page-in)
    t0 = time.perf_counter()
    for _ in range(repeats):
        for n in range(n_samples):
            o_egohmr.sample(sd, adj, N_BLOCKS, smpl_model, batch, sch, noise[n], mean, std, "ddim", dtype=np.float32,
                            hoist=False)
    dt = time.perf_counter() - t0
    bodies = n_img * n_samples * repeats
    return bodies / dt, dt, bodies


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_img, ns = args.cpu_sample_img, 1
    vals = []
    for _ in range(max(1, args.warmup // 3)):
        cpu_baseline(1, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, dt, bodies = cpu_baseline(n_img, ns)
        vals.append(v)
        if time.perf_counter() - t0 > 150:   # keep the whole arm within a few minutes
            break
    value = float(np.mean(vals))
    sample = (f"{n_img} images x {ns} sample per step ({n_img * ns} bodies) of the 64x10 workload, reference dataflow "
              f"(ResNet-50 + PointNet + GCN x2 + SMPL on every DDIM step), numpy/torch-CPU fp32, {len(vals)} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "bodies/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * n_img * ns / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "bodies/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "bodies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from egohmr_b200 import synth
    from egohmr_b200.testing import build_model, torch_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run for N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    n_img, S = args.n_img, args.num_samples
    B = n_img * S
    model, diffusion, sd, smpl_model, mean, std = build_model(HID, N_BLOCKS, T=T, respacing=RESPACING, device=str(dev))
    # weak scaling: every rank samples its own 64 x 10 shard of a (64*N) x 10 job (seed = rank); the only collective is
    # the final gather of the packed results
    batch_np = synth.make_batch(100 + rank, n_img, N_PTS)
    batch_dev = torch_batch(batch_np, dev)
    host = {k: (v if isinstance(v, dict) else torch.from_numpy(np.asarray(v)).pin_memory()) for k, v in batch_np.items()}
    host["smpl_params"] = {"transl": torch.from_numpy(batch_np["smpl_params"]["transl"]).pin_memory()}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values() if isinstance(t, torch.Tensor)) + \
        host["smpl_params"]["transl"].numel() * 4

    def one_step(batch):
        model._cond_key = None   # a new batch every step: the encoders run every time
        return diffusion.sample_many(model, batch, S, RESPACING)

    from egohmr_b200 import sharding

    def gather_results(out):
        # 226 floats = 904 B per body; one all_gather per sampling pass, no per-step communication
        return sharding.gather_results(sharding.pack_results(out), n_img * world, S)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        gather_results(one_step(batch_dev))
    barrier()
    if model.engine.check_overflow():
        raise SystemExit("fp16 operand overflow in the GCN layers")

    # ---- device-resident leg (value)
    clocks = ClockSampler(local)
    clocks.start()
    l0 = model.engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        gather_results(one_step(batch_dev))
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = model.engine.launch_count() - l0
    clk = clocks.stop()

    # ---- end-to-end leg: host (pinned) inputs -> public API -> host results, copies inside the timed region
    res_host = {"R": torch.empty(B, 24, 3, 3).pin_memory(), "betas": torch.empty(B, 10).pin_memory(),
                "joints": torch.empty(B, 45, 3).pin_memory()}
    d2h_bytes = sum(t.numel() * 4 for t in res_host.values())

    def e2e_step():
        b = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
        b["smpl_params"] = {"transl": host["smpl_params"]["transl"].to(dev, non_blocking=True)}
        out = one_step(b)
        gather_results(out)
        res_host["R"][:, :1].copy_(out["pred_smpl_params"]["global_orient"], non_blocking=True)
        res_host["R"][:, 1:].copy_(out["pred_smpl_params"]["body_pose"], non_blocking=True)
        res_host["betas"].copy_(out["pred_smpl_params"]["betas"], non_blocking=True)
        res_host["joints"].copy_(out["pred_keypoints_3d"], non_blocking=True)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- stage breakdown + dominant-kernel roofline (timed alone, after the step timing)
    def timed(fn, iters=5):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            fn()
        b_.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / iters

    def enc():
        model._cond_key = None
        model.prepare(batch_dev, S)

    enc_ms = timed(enc)
    one_step(batch_dev)
    layer_ms = [model.engine.time_hidden_layer(l, 20) for l in range(1, 2 * N_BLOCKS + 1)]
    avg_layer_ms = float(np.mean(layer_ms))
    rows = 2 * B * 24
    flop = 2.0 * rows * HID * (2 * HID)          # fp32-equivalent FLOPs of one hidden layer's GEMM
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0))
    achieved = flop / (avg_layer_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["gcn_hidden_umma_dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {
        "kernel": "gcn_hidden_umma_kernel (8 launches per reverse step, 40 per sampling pass)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else "fallback 1590",
        "issued_tflops": 3 * achieved, "issued_frac": 3 * achieved / peak,
        "note": "achieved counts the algorithmic fp32-equivalent FLOPs (2*rows*K*2C); the kernel issues 3 fp16 "
                "tcgen05.mma per product (hi*hi + hi*lo + lo*hi) to reach fp32-class accuracy, so the tensor pipe "
                "itself runs at issued_frac of the measured fp16/bf16 peak",
        "avg_launch_ms": avg_layer_ms, "per_layer_ms": layer_ms,
    }

    # max over ranks of the device time
    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_bodies = B * world * args.steps
    line = {
        "metric": METRIC, "value": total_bodies / (ms * 1e-3), "unit": "bodies/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (GEMMs: fp16x3 error-compensated on tcgen05, fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "bodies_per_gpu_per_step": B, "hid": HID, "blocks": N_BLOCKS,
                   "diffuse_fuse": True, "sharding": "images split across ranks, one NCCL all_gather of 904 B/body",
                   "l2": "no explicit flush: each step streams ~0.5 GB of activations (>> 126 MB L2)",
                   "encoders": "ResNet-50 + ResPointNet run once per step in PyTorch (cuDNN, torch default TF32 convs)",
                   "stage_ms": {"encoders_once_per_batch": enc_ms, "gcn_hidden_layers_per_reverse_step": float(np.sum(layer_ms))}},
        "e2e": {"value": total_bodies / e2e_s, "unit": "bodies/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "timer": "host wall clock, synchronize on both sides"},
        "gpu_launches": int(launches),
        "clocks": clk, "roofline": roofline,
    }
    if not args.no_cpu_baseline and world == 1:
        v, dt, bodies = cpu_baseline(args.cpu_sample_img, 1)
        line["cpu_baseline"] = {"value": v, "unit": "bodies/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{bodies} bodies ({args.cpu_sample_img} images x 1 sample) of the 64x10 workload in "
                                          f"{dt:.1f} s; oracle port with the reference's dataflow (encoders + SMPL on every step)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)

#!/usr/bin/env python
"""Benchmark of the EgoHMR diffusion-sampling hot path (BASELINE.json: sampled bodies/sec, DDIM-5, batch 64 x 10 samples).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the UNMODIFIED reference (baseline/_ref) on the host cores, CPU fp32
    python bench.py --impl reference-cuda ... # the UNMODIFIED reference on cuda:0 (torch defaults / strict fp32 / flat batch)
    python bench.py --config 5 --gpus 8 ...   # configs[4]: DDPM-1000, 512 x 10 split over the ranks (strong scaling)

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


# `config` is identical in every arm (the driver compares the two arms' configs); arm-specific measurements live in `detail`
CONFIG = {"workload": WORKLOAD, "n_img": N_IMG, "num_samples": N_SAMPLES, "T": T, "respacing": RESPACING, "n_pts": N_PTS,
          "hid": HID, "blocks": N_BLOCKS, "diffuse_fuse": True, "weights": "random-init (seeded), synthetic SMPL model",
          "l2": "GPU arms: no explicit flush, each step streams ~0.5 GB of activations (>> 126 MB L2)"}
CFG4_IMG, CFG5_IMG = 256, 512   # configs[3] / configs[4]: images of the whole job, split over the ranks (strong scaling)

_REAL_STDOUT = None


def protect_stdout():
    """Libraries (NCCL's version banner, cuDNN warnings) print to fd 1; the driver wants ONE JSON line there.  Point
    fd 1 at stderr for the whole run and keep the original for `emit`."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--no-graph", action="store_true", help="issue the pass from Python instead of replaying the CUDA graph")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip the eager-PyTorch reference-dataflow leg")
    ap.add_argument("--n-img", type=int, default=N_IMG)
    ap.add_argument("--num-samples", type=int, default=N_SAMPLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-img", type=int, default=16, help="images per step of the bounded CPU sample (one val_losses call)")
    ap.add_argument("--cpu-sample-samples", type=int, default=4, help="samples per image in the oracle-port CPU sample (fallback only)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 5],
                    help="2 = configs[1] (headline; weak scaling over ranks + a configs[3] strong-scaling leg), "
                         "5 = configs[4]: DDPM-1000, 512 images x 10 samples split over the ranks")
    ap.add_argument("--no-strong", action="store_true", help="skip the configs[3] strong-scaling leg")
    ap.add_argument("--strong-img", type=int, default=CFG4_IMG)
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


def _ref_harness():
    """baseline/ref_harness.py if an unmodified reference tree is available (baseline/_ref travels with the snapshot)."""
    from baseline import ref_harness as rh
    return rh if rh.reference_root()[0] else None


def run_reference(args):
    """`--impl reference`: the UNMODIFIED reference on the host cores, driven exactly like test_egohmr.py:247-266
    (sequential val_losses calls with the driver's kwargs, compute_loss left at its default True), CPU fp32, all cores.
    A step is one val_losses call over `--cpu-sample-img` images (a bounded slice of the 64 x 10 workload; the reference's
    CPU throughput per body does not grow beyond ~16 images per call)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rh = _ref_harness()
    n_img = args.cpu_sample_img
    if rh is None:      # no reference tree on this machine: the numpy oracle port with the reference's dataflow
        n_img, kind = min(n_img, 4), "port"
        for _ in range(max(1, args.warmup // 3)):
            cpu_baseline(1, 1)
        times = []
        for _ in range(args.steps):
            v, dt, bodies = cpu_baseline(n_img, 1)
            times.append(dt)
        what = "oracle port (numpy) with the reference's dataflow — baseline/_ref absent"
        extra = {}
    else:
        kind = "reference"
        model, mean, std = rh.build_reference(HID, N_BLOCKS)
        samp = rh.build_sampler(T, RESPACING, mean, std)
        batch = rh.make_driver_batch(100, n_img, N_PTS)
        for _ in range(args.warmup):
            rh.driver_loop(model, samp, batch, 1, RESPACING)
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            rh.driver_loop(model, samp, batch, 1, RESPACING)
            times.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        rh.driver_loop(model, samp, batch, 1, RESPACING, compute_loss=False)
        extra = {"value_without_compute_loss": n_img / (time.perf_counter() - t0)}
        what = (f"unmodified reference ({rh.reference_root()[1]}; smplx / coap / yacs stand-ins, baseline/ref_harness.py), "
                "test_egohmr.py:247-266 loop, torch CPU fp32")
    value = n_img * len(times) / float(np.sum(times))
    sample = (f"{n_img} images x 1 sample per step = one val_losses call ({n_img} bodies) of the 64 x 10 workload; {what}; "
              f"{len(times)} steps, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "bodies/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(CONFIG),
        "cpu_baseline": {"value": value, "unit": "bodies/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "bodies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "detail": dict(extra, bodies_per_step=n_img),
    }
    emit(line)


def _cuda_timed(fn):
    import torch
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def reference_cuda_unmodified(device, n_img=N_IMG, n_samples=N_SAMPLES, repeats=1):
    """The UNMODIFIED reference on the GPU (SURVEY.md 8d / BASELINE.md 3: the denominator of the ">= 10x" target), same
    harness as the CPU arm: (a) the driver's loop — `n_samples` sequential val_losses calls of `n_img` images — with
    torch's default flags (cuDNN convolutions may use TF32, matmuls are fp32); (b) the same with cudnn.allow_tf32=False
    (strict fp32, the precision this repo's path delivers); (c) "flat": ONE val_losses call on the batch tiled to
    n_img*n_samples bodies, the most favourable batched use of the reference's own code.  Device-resident inputs, CUDA
    events.  -> dict of bodies/s, or None when no reference tree is available."""
    import torch
    rh = _ref_harness()
    if rh is None:
        return None
    model, mean, std = rh.build_reference(HID, N_BLOCKS, device=device)
    samp = rh.build_sampler(T, RESPACING, mean, std, device=device)
    batch = rh.make_driver_batch(100, n_img, N_PTS, device=device)
    bodies = n_img * n_samples
    out = {"what": f"unmodified reference ({rh.reference_root()[1]}) through baseline/ref_harness.py on {device}: "
                   f"test_egohmr.py:247-266 loop, {n_img} images x {n_samples} sequential samples, DDIM-5",
           "unit": "bodies/s"}
    kw = {}
    try:
        rh.driver_loop(model, samp, batch, 1, RESPACING)             # warm-up: cuDNN / cuBLAS plan selection
    except Exception as e:                                            # noqa: BLE001 - keep the arm alive, say what happened
        out["compute_loss_error"] = f"{type(e).__name__}: {e}"[:200]
        kw = {"compute_loss": False}
        rh.driver_loop(model, samp, batch, 1, RESPACING, **kw)
    old = torch.backends.cudnn.allow_tf32
    try:
        for label, tf32 in (("torch_defaults", True), ("cudnn_tf32_off", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            rh.driver_loop(model, samp, batch, 1, RESPACING, **kw)
            ms = min(_cuda_timed(lambda: rh.driver_loop(model, samp, batch, n_samples, RESPACING, **kw)) for _ in range(repeats))
            out[label] = bodies / (ms * 1e-3)
            out[label + "_ms"] = ms
        torch.backends.cudnn.allow_tf32 = True
        flat = rh.flat_batch(batch, n_samples)
        rh.driver_loop(model, samp, flat, 1, RESPACING, **kw)
        ms = min(_cuda_timed(lambda: rh.driver_loop(model, samp, flat, 1, RESPACING, **kw)) for _ in range(repeats))
        out["flat_torch_defaults"] = bodies / (ms * 1e-3)
        out["flat_torch_defaults_ms"] = ms
    finally:
        torch.backends.cudnn.allow_tf32 = old
    out["value"] = out["torch_defaults"]
    return out


def reference_cuda(n_img, n_samples, device, repeats=1):
    """Cross-check only: oracle/torch_eager.py, this repo's eager-PyTorch restatement of the reference's GPU dataflow
    (pinned on the reference's goldens).  The headline denominator is `reference_cuda_unmodified`."""
    import torch
    from egohmr_b200 import synth
    from egohmr_b200.testing import torch_batch
    from oracle import torch_eager
    model, sch = torch_eager.build(HID, N_BLOCKS, T, RESPACING, device)
    batch = torch_batch(synth.make_batch(100, n_img, N_PTS), device)
    torch_eager.val_losses(model, sch, batch, [n_img, 144], "ddim")   # warm-up: cuDNN/cuBLAS plan selection
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(repeats):
        for _n in range(n_samples):
            torch_eager.val_losses(model, sch, batch, [n_img, 144], "ddim")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return n_img * n_samples * repeats / (ms * 1e-3), ms


def run_reference_cuda(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    torch.cuda.set_device(0)
    res = reference_cuda_unmodified("cuda:0", args.n_img, args.num_samples, repeats=max(1, min(args.steps, 3)))
    kind = "reference"
    if res is None:
        v, ms = reference_cuda(args.n_img, args.num_samples, "cuda:0")
        res, kind = {"value": v, "torch_defaults_ms": ms, "what": "oracle/torch_eager.py port (baseline/_ref absent)"}, "port"
    emit({
        "impl": "reference-cuda", "metric": METRIC, "value": res["value"], "unit": "bodies/s", "n_gpus": 1,
        "steps": max(1, min(args.steps, 3)), "warmup": 1, "ms_per_step": res["torch_defaults_ms"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (torch defaults: TF32 cuDNN convolutions, fp32 matmuls)",
        "data": "synthetic", "config": dict(CONFIG), "detail": dict(res, kind=kind)})


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi polled every 20 ms in the background; rows are time-stamped on receipt so that `stop(t0, t1)` keeps the
    samples taken DURING the timed region (nvidia-smi needs ~1 s to start: start it well before)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 is None or (t0 <= ts <= t1 + 0.03)]
        window = "timed region"
        if not rows:     # region shorter than the polling period: fall back to everything seen (warm-up replays included)
            rows, window = [r for _, r in self.rows], "warm-up + timed region"
        sm, mx, reasons = [], [], set()
        for r in rows:
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
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------ GPU arm
def _pinned_host_batch(batch_np):
    import torch
    host = {k: (v if isinstance(v, dict) else torch.from_numpy(np.asarray(v)).pin_memory()) for k, v in batch_np.items()}
    host["smpl_params"] = {"transl": torch.from_numpy(batch_np["smpl_params"]["transl"]).pin_memory()}
    nbytes = sum(t.numel() * t.element_size() for t in host.values() if isinstance(t, torch.Tensor)) + \
        host["smpl_params"]["transl"].numel() * 4
    return host, nbytes


def strong_leg(args, model, diffusion, dev, rank, world, total_img, respacing, passes, label):
    """Strong scaling: ONE job of `total_img` images x 10 samples split over the ranks by image (sharding.shard_bounds),
    one CUDA graph per shard shape, one all_gather of the packed results per pass.  Per pass and rank the compute time
    (graph replay) and the gather are timed separately with CUDA events; the job time is the max over ranks."""
    import torch
    import torch.distributed as dist
    from egohmr_b200 import sharding, synth
    from egohmr_b200.testing import torch_batch
    S = args.num_samples
    lo, hi = sharding.shard_bounds(total_img, rank, world)
    n_loc = hi - lo
    full = synth.make_batch(300, total_img, N_PTS)      # the same global job on every rank; each keeps its image range
    batch = sharding.shard_batch(torch_batch(full, dev), rank, world)
    batch = {k: ({kk: vv.contiguous() for kk, vv in v.items()} if isinstance(v, dict) else
                 (v.contiguous() if isinstance(v, torch.Tensor) else v)) for k, v in batch.items()}
    sampler = diffusion.capture_sample_many(model, batch, S, respacing)
    gather = lambda out: sharding.gather_results(sharding.pack_results(out), total_img, S)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        full_res = gather(sampler(batch))
    barrier()
    # once: rank 0's slice of the gathered buffer is its own packed result (the collective moved the right rows)
    own = sharding.pack_results(sampler.static_out)
    assert torch.equal(full_res[lo * S: hi * S], own), "gathered buffer does not hold this rank's packed result"
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(passes)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(passes):
        ev[i][0].record()
        out = sampler(batch)
        ev[i][1].record()
        gather(out)
        ev[i][2].record()
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    comp = float(np.mean([ev[i][0].elapsed_time(ev[i][1]) for i in range(passes)]))
    gath = float(np.mean([ev[i][1].elapsed_time(ev[i][2]) for i in range(passes)]))
    stats = torch.tensor([total_ms, comp, gath], device=dev, dtype=torch.float64)
    if world > 1:
        allst = [torch.empty_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)
        allst = torch.stack(allst).cpu().numpy()
    else:
        allst = stats.cpu().numpy()[None]
    del sampler
    t_max = float(allst[:, 0].max())
    bodies = total_img * S
    return {"workload": label, "scaling": "strong", "bodies_per_pass": bodies, "passes": passes,
            "value": bodies * passes / (t_max * 1e-3), "unit": "bodies/s", "ms_per_pass": t_max / passes,
            "bodies_per_rank": [int((sharding.shard_bounds(total_img, r, world)[1] - sharding.shard_bounds(total_img, r, world)[0]) * S)
                                for r in range(world)],
            "per_rank_compute_ms": {"min": float(allst[:, 1].min()), "median": float(np.median(allst[:, 1])),
                                    "max": float(allst[:, 1].max())},
            "per_rank_gather_ms": {"min": float(allst[:, 2].min()), "median": float(np.median(allst[:, 2])),
                                   "max": float(allst[:, 2].max())},
            "note": "gather = NCCL all_gather of 904 B/body incl. waiting for the slowest rank; compute = graph replay"}


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

    if args.config == 5:
        return run_cfg5(args, dev, rank, world)

    n_img, S = args.n_img, args.num_samples
    B = n_img * S
    model, diffusion, sd, smpl_model, mean, std = build_model(HID, N_BLOCKS, T=T, respacing=RESPACING, device=str(dev))
    # weak scaling: every rank samples its own 64 x 10 shard of a (64*N) x 10 job (seed = rank); the only collective is
    # the final gather of the packed results
    batch_np = synth.make_batch(100 + rank, n_img, N_PTS)
    batch_dev = torch_batch(batch_np, dev)
    host, h2d_bytes = _pinned_host_batch(batch_np)

    def eager_step(batch):
        model.invalidate()   # a new batch every step: the encoders run every time
        return diffusion.sample_many(model, batch, S, RESPACING)

    from egohmr_b200 import sharding

    def gather_results(out):
        # 226 floats = 904 B per body; one all_gather per sampling pass, no per-step communication
        return sharding.gather_results(sharding.pack_results(out), n_img * world, S)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager: weight repacking, workspaces), then capture the pass as one CUDA graph
    for _ in range(max(args.warmup, 3)):
        gather_results(eager_step(batch_dev))
    barrier()
    if model.engine.check_overflow():
        raise SystemExit("fp16 operand overflow in the GCN layers")
    l0 = model.engine.launch_count()
    eager_step(batch_dev)
    launches_per_pass = model.engine.launch_count() - l0
    clocks = ClockSampler(local)
    clocks.start()
    sampler = None
    if not args.no_graph:
        sampler = diffusion.capture_sample_many(model, batch_dev, S, RESPACING)
        launches_per_pass = sampler.launches_per_replay
        for _ in range(2):
            full_res = gather_results(sampler(batch_dev))
        if world > 1:   # once: this rank's slice of the gathered buffer equals its own packed result
            torch.cuda.synchronize()
            assert torch.equal(full_res[rank * B: (rank + 1) * B], sharding.pack_results(sampler.static_out))
    one_step = sampler if sampler is not None else eager_step
    barrier()

    def timed_steps(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            gather_results(fn(batch_dev))
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    # ---- device-resident leg (value)
    clocks.wait_first()
    timed_steps(one_step, 2)
    w0 = time.time()
    ms = timed_steps(one_step, args.steps)
    clk = clocks.stop(w0, time.time())
    launches = launches_per_pass * args.steps
    eager_ms = timed_steps(eager_step, args.steps) if sampler is not None else ms
    # the same pass with the image encoder on cuDNN instead of the native tcgen05 convolution GEMMs, at torch's default
    # conv precision (TF32 allowed: what the reference's own CUDA path runs) — a library baseline for that stage
    model.native_image_enc = False
    try:
        eager_step(batch_dev)
        n_alt = max(3, args.steps // 4)
        cudnn_ms = timed_steps(eager_step, n_alt) / n_alt * args.steps
    finally:
        model.native_image_enc = True
        model.invalidate()

    # ---- end-to-end leg: host (pinned) inputs -> public API -> host results, copies inside the timed region
    # results: the packed [B, 226] rotation matrices + betas (what test_egohmr.py keeps per sample) and the 45 joints, as two
    # CONTIGUOUS device -> pinned-host copies (a copy into a strided slice of a host tensor goes through a pageable
    # temporary and blocks the host: 0.55 ms of GPU idle time per step, tools/e2e_probe.py)
    res_host = {"params": torch.empty(B, 226).pin_memory(), "joints": torch.empty(B, 45, 3).pin_memory()}
    d2h_bytes = sum(t.numel() * 4 for t in res_host.values())

    def e2e_step(last=False):
        if sampler is not None:
            # pinned host tensors -> device staging copy on the copy stream (overlaps the previous replay) -> the graph's
            # static inputs (device-to-device) -> one replay; the next step's upload starts as soon as this one is queued
            out = sampler(staged=True)
            if not last:
                sampler.stage(host)
        else:
            b = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
            b["smpl_params"] = {"transl": host["smpl_params"]["transl"].to(dev, non_blocking=True)}
            out = eager_step(b)
        full = gather_results(out)
        res_host["params"].copy_(full[rank * B:(rank + 1) * B], non_blocking=True)
        res_host["joints"].copy_(out["pred_keypoints_3d"].contiguous(), non_blocking=True)

    if sampler is not None:
        sampler.stage(host)
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    # exactly K uploads, K passes and K downloads inside the timed region: the first upload is issued here (the warm-up
    # step's look-ahead copy is simply overwritten), each step issues the next one, the last step issues none
    if sampler is not None:
        sampler.stage(host)
    for i in range(args.steps):
        e2e_step(last=(i == args.steps - 1))
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- the literally unchanged driver loop (test_egohmr.py:251-255): num_samples sequential val_losses calls
    # (a new batch object per pass, as a dataloader delivers; from the second batch on val_losses runs the samples of a
    # batch ahead as one pass — EgoHMR.samples_ahead — and the calls 2..S return the stored samples)
    drop_batches = [batch_dev, torch_batch(synth.make_batch(200 + rank, n_img, N_PTS), dev)]

    def dropin_pass(i):
        outs = []
        for _ in range(S):
            o = diffusion.val_losses(model=model, batch=drop_batches[i % 2], shape=[n_img, 144], progress=False,
                                     clip_denoised=False, cur_epoch=0, timestep_respacing=RESPACING, cond_fn_with_grad=False,
                                     cond_grad_weight=1.0, compute_loss=False)
            outs.append(o["pred_smpl_params"]["body_pose"].unsqueeze(1))
        return torch.cat(outs, dim=1)
    for i in range(3):
        dropin_pass(i)
    n_drop = max(4, args.steps // 2)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_drop):
        dropin_pass(3 + i)
    barrier()
    dropin_s = (time.perf_counter() - t0) / n_drop

    # ---- stage breakdown + per-kernel rooflines (each kernel timed alone with CUDA events, after the step timing)
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
        model.invalidate()
        model.prepare(batch_dev, S)

    enc_ms = timed(enc)
    img_c = batch_dev["img"].float().contiguous()
    resnet_ms = timed(lambda: model.engine.resnet_forward(img_c), 5)
    out_e = eager_step(batch_dev)     # leaves the context on this batch's activations / slot tables
    x_t = torch.randn(B, 144, device=dev)
    n_layers = 2 * N_BLOCKS
    k2_ms = model.engine.time_stage(0, 2, x_t, 20)
    layer_ms = [model.engine.time_stage(l, 2, x_t, 20) for l in range(1, n_layers + 1)]
    k3_ms = model.engine.time_stage(n_layers + 1, 2, x_t, 20)
    x0 = out_e["pred_x_start"].contiguous()
    betas_img = model._cond["betas_img"]
    dec_ms = timed(lambda: model.engine.decode(x0, betas_img, want_smpl=True), 10)
    avg_layer_ms = float(np.mean(layer_ms))
    rows = 2 * B * 24
    flop = 2.0 * rows * HID * (2 * HID)          # fp32-equivalent FLOPs of one hidden layer's GEMM
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0))
    hbm_peak = float(peaks.get("hbm_gbs", 6500.0))
    achieved = flop / (avg_layer_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["gcn_hidden_umma_dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {
        "kernel": "gcn_hidden_umma_t_kernel (8 launches per reverse step, 40 per sampling pass)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else "fallback 1590",
        "issued_tflops": 3 * achieved, "issued_frac": 3 * achieved / peak,
        "note": "achieved counts the algorithmic fp32-equivalent FLOPs (2*rows*K*2C); the kernel issues 3 fp16 "
                "tcgen05.mma per product (hi*hi + hi*lo + lo*hi) to reach fp32-class accuracy, so the tensor pipe "
                "itself runs at issued_frac of the measured fp16/bf16 peak",
        "avg_launch_ms": avg_layer_ms, "per_layer_ms": layer_ms,
    }
    # the HBM-bound kernels of the step (algorithmic bytes per body, constants once per launch)
    slot_bytes = 24 * HID * 4 + 24 * 2 * HID * 2            # fp32 activations + fp16 hi/lo operand of one (body, pass) slot
    k2_bytes = 2 * B * slot_bytes + B * 576
    k3_bytes = 2 * B * 24 * HID * 4 + B * 576 * 3
    V = model.engine.n_verts
    dec_bytes = B * (V * 12 + 45 * 12 + 576 * 2 + 864) + (207 * V * 3 + V * 3 * 11 + V * 24) * 4
    hbm = lambda name, nbytes, t_ms, note: {"kernel": name, "bound": "hbm", "achieved": nbytes / (t_ms * 1e-3) / 1e9,
                                            "peak": hbm_peak, "unit": "GB/s", "frac": nbytes / (t_ms * 1e-3) / 1e9 / hbm_peak,
                                            "algorithmic_bytes": nbytes, "avg_launch_ms": t_ms, "note": note}
    roofline_other = [
        hbm("gcn_input_kernel (K2)", k2_bytes, k2_ms, "writes fp32 + fp16 hi/lo activations of every slot: 393 KB/body"),
        hbm("gcn_output_kernel (K3: output layer + fuse-select + sampler update)", k3_bytes, k3_ms,
            "reads the fp32 activations of both passes: 196.6 KB/body"),
        hbm("ehb_decode (K4 rot6d + K5 SMPL: chain, pose-blend GEMM on tcgen05, tiled skinning, joints; 7 launches)",
            dec_bytes + 2 * B * V * 12, dec_ms,
            "84.1 KB/body out + the fp32 pose offsets written and re-read once (2 x 82.7 KB/body) + 19.3 MB constants"),
    ]
    # ResNet-50 (K9) against the tensor roof: 4.087 GMAC per image, x3 issued
    rn_tflops = 2 * 4.087e9 * n_img / (resnet_ms * 1e-3) / 1e12
    roofline_other.append({"kernel": "conv_gemm_kernel x 53 + movers (K9 ResNet-50, once per pass)", "bound": "tensor",
                           "achieved": rn_tflops, "peak": peak, "unit": "TFLOP/s", "frac": rn_tflops / peak,
                           "issued_frac": 3 * rn_tflops / peak, "avg_launch_ms": resnet_ms,
                           "note": f"whole encoder for {n_img} images; algorithmic 4.087 GMAC/image"})

    # ---- strong scaling: configs[3] (256 x 10) split over the ranks
    strong = None
    if not args.no_strong:
        del sampler
        sampler = None
        torch.cuda.empty_cache()
        strong = strong_leg(args, model, diffusion, dev, rank, world, args.strong_img, RESPACING, max(5, args.steps // 2),
                            f"configs[3]: DDIM-5 (T=50), {args.strong_img} images x {S} samples = {args.strong_img * S} bodies per "
                            f"pass, split by image over {world} rank(s)")

    ref_cuda = ref_port = None
    if rank == 0 and world == 1 and not args.no_reference_cuda:
        ref_cuda = reference_cuda_unmodified(str(dev))
        v, rms = reference_cuda(n_img, 2, str(dev))
        ref_port = {"value": v, "unit": "bodies/s", "sample": f"{n_img} images x 2 of the {S} samples in {rms:.0f} ms",
                    "what": "cross-check: oracle/torch_eager.py, this repo's eager-PyTorch restatement of the reference's GPU "
                            "dataflow (pinned on the reference's goldens)"}

    # max over ranks of the device time
    if world > 1:
        t = torch.tensor([ms, e2e_s, eager_ms, cudnn_ms, dropin_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, eager_ms, cudnn_ms, dropin_s = (float(x) for x in t)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_bodies = B * world * args.steps
    line = {
        "metric": METRIC, "value": total_bodies / (ms * 1e-3), "unit": "bodies/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (GEMMs: fp16x3 error-compensated on tcgen05, fp32 accumulate)",
        "data": "synthetic", "config": dict(CONFIG),
        "detail": {"bodies_per_gpu_per_step": B,
                   "sharding": "images split across ranks, one NCCL all_gather of 904 B/body",
                   "execution": ("whole pass replayed as ONE CUDA graph (diffusion/graphed.py)" if not args.no_graph
                                 else "eager: every launch issued from Python"),
                   "eager_value": total_bodies / (eager_ms * 1e-3),
                   "eager_cudnn_tf32_enc": total_bodies / (cudnn_ms * 1e-3),
                   "dropin_loop_value": B * world / dropin_s,
                   "dropin_loop": "the reference driver's own loop (test_egohmr.py:251-255): 10 sequential val_losses calls "
                                  "per batch, host clock",
                   "encoders": "once per pass, both native (K9 ResNet-50 conv GEMMs, K7 ResPointNet), fp16x3 = fp32-class; "
                               "eager_cudnn_tf32_enc = same pass with the cuDNN TF32 image encoder",
                   "stage_ms": {"encoders_once_per_pass": enc_ms, "resnet50_once_per_pass": resnet_ms,
                                "gcn_input_per_step": k2_ms,
                                "gcn_hidden_layers_per_step": float(np.sum(layer_ms)), "gcn_output_per_step": k3_ms,
                                "decode_once_per_pass": dec_ms}},
        "e2e": {"value": total_bodies / e2e_s, "unit": "bodies/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "timer": "host wall clock, synchronize on both sides",
                "pipelining": "inputs of step i+1 upload on a copy stream while step i computes (GraphedSampler.stage)"},
        "gpu_launches": int(launches),
        "clocks": clk, "roofline": roofline, "roofline_other": roofline_other,
    }
    if strong is not None:
        line["strong"] = strong
    if ref_cuda is not None:
        line["reference_cuda"] = ref_cuda
    if ref_port is not None:
        line["reference_cuda_port"] = ref_port
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = bounded_cpu_baseline(args)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def bounded_cpu_baseline(args):
    """`cpu_baseline` of the GPU arm's line: the unmodified reference on this box's host cores (the oracle port only when
    no reference tree travelled), a bounded sample of the 64 x 10 workload (~10-30 s)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rh = _ref_harness()
    n_img = args.cpu_sample_img
    if rh is None:
        v, dt, bodies = cpu_baseline(8, args.cpu_sample_samples)
        return {"value": v, "unit": "bodies/s", "cores": cores, "kind": "port",
                "sample": f"{bodies} bodies of the 64x10 workload in {dt:.1f} s; oracle port with the reference's dataflow"}
    model, mean, std = rh.build_reference(HID, N_BLOCKS)
    samp = rh.build_sampler(T, RESPACING, mean, std)
    batch = rh.make_driver_batch(100, n_img, N_PTS)
    rh.driver_loop(model, samp, rh.make_driver_batch(101, 2, N_PTS), 1, RESPACING)   # page-in / thread pools
    t0 = time.perf_counter()
    n_calls = 0
    while n_calls < 3 or (time.perf_counter() - t0 < 10 and n_calls < 10):
        rh.driver_loop(model, samp, batch, 1, RESPACING)
        n_calls += 1
    dt = time.perf_counter() - t0
    return {"value": n_img * n_calls / dt, "unit": "bodies/s", "cores": cores, "kind": "reference",
            "sample": f"{n_calls} val_losses calls x {n_img} images ({n_img * n_calls} bodies) of the 64x10 workload in {dt:.1f} s; "
                      f"unmodified reference ({rh.reference_root()[1]}), torch CPU fp32, {cores} threads"}


def run_cfg5(args, dev, rank, world):
    """configs[4]: DDPM-1000 full sampling, 512 images x 10 samples, split by image over the ranks; the 1000-step pass of a
    shard is ONE CUDA graph."""
    import torch
    import torch.distributed as dist
    from egohmr_b200.testing import build_model
    model, diffusion, sd, smpl_model, mean, std = build_model(HID, N_BLOCKS, T=1000, respacing="", device=str(dev))
    total_img = args.n_img if args.n_img != N_IMG else CFG5_IMG
    clocks = ClockSampler(dev.index)
    clocks.start()
    res = strong_leg(args, model, diffusion, dev, rank, world, total_img, "", max(1, args.steps),
                     f"configs[4]: DDPM-1000 full sampling, {total_img} images x {args.num_samples} samples, split by image "
                     f"over {world} rank(s); one CUDA graph per 1000-step pass")
    clk = clocks.stop()
    if rank == 0:
        cfg = dict(CONFIG, workload=res["workload"], n_img=total_img, T=1000, respacing="ddpm")
        emit({"metric": "sampled bodies/sec (batch x num_samples), DDPM-1000", "value": res["value"], "unit": "bodies/s",
              "n_gpus": world, "steps": res["passes"], "warmup": 2, "ms_per_step": res["ms_per_pass"],
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "f32 (GEMMs: fp16x3 error-compensated on tcgen05, fp32 accumulate)", "data": "synthetic",
              "config": cfg, "detail": res, "clocks": clk})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    protect_stdout()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "reference-cuda":
        run_reference_cuda(a)
    else:
        run_b200(a)

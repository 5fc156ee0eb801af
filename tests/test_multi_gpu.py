"""The N > 1 path on real GPUs (skipped with fewer than two): one process per GPU under torch.distributed.run, NCCL.
Every rank samples its image shard with its slice of a globally pre-drawn noise tensor; the NCCL all_gather of the packed
results must equal the unsharded single-GPU run bit for bit, for an even and a ragged split."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["EHB_ROOT"])
from egohmr_b200 import sharding, synth
from egohmr_b200.testing import build_model, torch_batch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model, diffusion, *_ = build_model(1024, 4, T=50, respacing="ddim5", device=str(dev), collision=False)
S = 3
for n_img in (6, 5):                                   # even and ragged split over the ranks
    batch = torch_batch(synth.make_batch(40 + n_img, n_img), dev)
    noise = torch.from_numpy(synth.make_noise(41, 1, n_img * S, 5)[0]).to(dev)
    mine = sharding.shard_batch(batch, rank, world)
    out = diffusion.sample_many(model, mine, S, "ddim5", noise=sharding.shard_noise(noise, n_img, S, rank, world))
    full = sharding.gather_results(sharding.pack_results(out), n_img, S)
    ref = sharding.pack_results(diffusion.sample_many(model, batch, S, "ddim5", noise=noise))   # every rank: the whole job
    assert full.shape == ref.shape == (n_img * S, sharding.PACKED_WIDTH), (full.shape, ref.shape)
    assert torch.equal(full, ref), f"rank {rank}: gathered result differs from the unsharded run (n_img={n_img})"
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("NCCL_GATHER_OK", world)
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_gather_of_sharded_ranks_equals_the_unsharded_run(tmp_path):
    world = 2
    script = tmp_path / "nccl_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, EHB_ROOT=ROOT, NCCL_DEBUG="WARN")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-4000:])
    assert f"NCCL_GATHER_OK {world}" in p.stdout

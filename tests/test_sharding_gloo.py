"""world_size-2 `gloo` test (CPU) of the multi-GPU plan: image sharding, packing, and the single result gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egohmr_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for n in (1, 2, 7, 64, 65, 256):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_noise_partitions_the_global_draw():
    """Concatenating every rank's noise slice in rank order gives back the global tensor, ragged shards included."""
    import torch
    from egohmr_b200 import sharding
    n_img, S = 7, 3
    noise = torch.arange(6 * n_img * S * 144, dtype=torch.float32).reshape(6, n_img * S, 144)
    for world in (1, 2, 3, 8):
        parts = [sharding.shard_noise(noise, n_img, S, r, world) for r in range(world)]
        assert torch.equal(torch.cat(parts, dim=1), noise)
        for r, p in enumerate(parts):
            lo, hi = sharding.shard_bounds(n_img, r, world)
            assert p.shape == (6, (hi - lo) * S, 144)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    out = {"pred_smpl_params": {"global_orient": torch.from_numpy(rng.normal(size=(6, 1, 3, 3)).astype(np.float32)),
                                "body_pose": torch.from_numpy(rng.normal(size=(6, 23, 3, 3)).astype(np.float32)),
                                "betas": torch.from_numpy(rng.normal(size=(6, 10)).astype(np.float32))}}
    packed = sharding.pack_results(out)
    assert packed.shape == (6, sharding.PACKED_WIDTH)
    back = sharding.unpack_results(packed)
    for k in back:
        assert torch.equal(back[k], out["pred_smpl_params"][k])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_img, S, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the "global" job: body b = img*S + n carries the value 1000*img + n in every packed column
        batch = {"img": torch.arange(n_img, dtype=torch.float32).reshape(n_img, 1, 1, 1),
                 "fx": torch.arange(n_img, dtype=torch.float32),
                 "smpl_params": {"transl": torch.arange(n_img, dtype=torch.float32).reshape(n_img, 1).repeat(1, 3)}}
        mine = sharding.shard_batch(batch, rank, world)
        lo, hi = sharding.shard_bounds(n_img, rank, world)
        assert mine["img"].shape[0] == hi - lo and torch.equal(mine["smpl_params"]["transl"][:, 0], mine["fx"])
        ids = (mine["fx"].repeat_interleave(S) * 1000 + torch.arange(S, dtype=torch.float32).repeat(hi - lo))
        packed = ids[:, None].repeat(1, sharding.PACKED_WIDTH).contiguous()
        full = sharding.gather_results(packed, n_img, S)
        expect = torch.arange(n_img, dtype=torch.float32).repeat_interleave(S) * 1000 + torch.arange(S, dtype=torch.float32).repeat(n_img)
        ok = full.shape == (n_img * S, sharding.PACKED_WIDTH) and torch.equal(full[:, 0], expect) and torch.equal(full[:, -1], expect)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_img", [8, 7])   # even split and ragged split
def test_gather_results_world2_gloo(n_img):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_img, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]

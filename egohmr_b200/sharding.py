"""Multi-GPU plan for the sampling path: bodies are independent, so images are split contiguously across
ranks (all samples of an image stay on one rank, its encoder features are computed once), weights are replicated and
the ONLY collective is one all_gather of the packed per-body result at the end.  Backend-agnostic (`nccl` on GPUs,
`gloo` in the CPU tests)."""
import torch
import torch.distributed as dist

PACKED_WIDTH = 9 + 207 + 10   # global_orient + body_pose (rotation matrices) + betas = 226 floats = 904 B per body


def shard_bounds(n_img, rank, world):
    """Contiguous [lo, hi) image range of `rank`; sizes differ by at most one, earlier ranks take the remainder."""
    base, extra = divmod(n_img, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    """Slice every per-image tensor of a reference-schema batch dict (dataloaders/egobody_dataset.py:241-277)."""
    n_img = batch["img"].shape[0]
    lo, hi = shard_bounds(n_img, rank, world)
    out = {}
    for k, v in batch.items():
        if isinstance(v, dict):
            out[k] = {kk: vv[lo:hi] for kk, vv in v.items()}
        elif isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n_img:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def shard_noise(noise, n_img_total, num_samples, rank, world):
    """Slice a globally pre-drawn noise tensor [n_steps+1, n_img_total*num_samples, 144] (reference draw order, body =
    image*num_samples + n) to the bodies of this rank's images, so that a sharded run reproduces the single-GPU chains
    bit for bit: feed the result to `sample_many(..., noise=...)` / `GraphedSampler(external_noise=True)`."""
    lo, hi = shard_bounds(n_img_total, rank, world)
    return noise[:, lo * num_samples: hi * num_samples].contiguous()


def pack_results(out):
    """[B, 226]: what test_egohmr.py keeps per sample (pred_smpl_params, :260-266)."""
    p = out["pred_smpl_params"]
    B = p["betas"].shape[0]
    return torch.cat([p["global_orient"].reshape(B, 9), p["body_pose"].reshape(B, 207), p["betas"].reshape(B, 10)],
                     dim=1).contiguous()


def unpack_results(packed):
    B = packed.shape[0]
    return {"global_orient": packed[:, :9].reshape(B, 1, 3, 3), "body_pose": packed[:, 9:216].reshape(B, 23, 3, 3),
            "betas": packed[:, 216:226]}


def gather_results(packed, n_img_total, num_samples, group=None):
    """all_gather of ragged shards -> [n_img_total*num_samples, 226] in global body order on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return packed
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [(shard_bounds(n_img_total, r, world)[1] - shard_bounds(n_img_total, r, world)[0]) * num_samples
             for r in range(world)]
    assert packed.shape[0] == sizes[rank], (packed.shape, sizes, rank)
    if len(set(sizes)) == 1:
        full = torch.empty(sum(sizes), packed.shape[1], device=packed.device, dtype=packed.dtype)
        dist.all_gather_into_tensor(full, packed, group=group)
        return full
    mx = max(sizes)
    pad = torch.zeros(mx, packed.shape[1], device=packed.device, dtype=packed.dtype)
    pad[: packed.shape[0]] = packed
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)

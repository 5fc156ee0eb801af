"""One CUDA graph for a whole sampling pass (encoders once, every reverse step, rot6d + SMPL + projection).

`GaussianDiffusion.sample_many` issues ~210 kernel launches per pass from Python; at cfg 2 the GPU waits on the host
for ~10 % of the pass, and at the strong-scaling shard sizes (32 images per GPU) for much more.  Every
launch of the pass is stream-ordered and allocation-free after the first call (the C ABI allocates its workspaces on
first use only, include/egohmr_b200.h), so the pass can be captured once and replayed:

    sampler = GraphedSampler(diffusion, model, batch, num_samples, "ddim5")
    out = sampler(batch)          # copies the batch into the graph's static inputs, replays, returns static outputs

The outputs are the graph's own buffers: they are overwritten by the next call (clone what must survive).  The torch
CUDA generator is registered with the graph, so every replay draws fresh noise exactly like `th.randn` /
`th.randn_like` in the eager loop (gaussian_diffusion.py:478,331,547 of the reference).  Collision-guided sampling is
not capturable (the collision callback is arbitrary Python with data-dependent control flow, egohmr.py:545-559) and
stays on the eager path.
"""
import numpy as np
import torch as th


def _flat_items(d, prefix=()):
    for k, v in d.items():
        if isinstance(v, dict):
            yield from _flat_items(v, prefix + (k,))
        elif isinstance(v, th.Tensor):
            yield prefix + (k,), v


def _get(d, path):
    for k in path:
        d = d[k]
    return d


class GraphedSampler:
    def __init__(self, diffusion, model, batch, num_samples, timestep_respacing="", warmup=2, external_noise=False):
        """`external_noise=True`: the pass reads its noise from a static [n_steps+1, bs*num_samples, 144] buffer filled
        by `__call__(batch, noise=...)` (reference draw order, see `sample_many`) instead of torch's generator —
        reproducible chains, e.g. when a global noise tensor is scattered over ranks."""
        if not th.cuda.is_available():
            raise RuntimeError("GraphedSampler needs a CUDA device (no CPU fallback)")
        self.diffusion, self.model = diffusion, model
        self.num_samples, self.respacing = num_samples, timestep_respacing
        self._clone = lambda d: {k: (self._clone(v) if isinstance(v, dict) else (v.clone() if isinstance(v, th.Tensor) else v))
                                 for k, v in d.items()}
        self.static_batch = self._clone(batch)
        # every tensor of the batch is a graph input, except the two keys the sampler itself writes into the dict
        # (gaussian_diffusion.py:256 'x_t', egohmr.py:189 'vis_mask_smpl')
        self._paths = [p for p, _ in _flat_items(self.static_batch) if p[0] not in ("x_t", "vis_mask_smpl")]
        self._staging = None
        self.static_noise = None
        if external_noise:
            n_bodies = batch["img"].shape[0] * num_samples
            self.static_noise = th.randn(diffusion.num_timesteps + 1, n_bodies, 144, device=batch["img"].device)
        # warm-up on a side stream: weight repacking, workspace allocation, cuDNN plan selection, slot tables
        s = th.cuda.Stream()
        s.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(s):
            for _ in range(max(1, warmup)):
                model._cond_key = None
                diffusion.sample_many(model, self.static_batch, num_samples, timestep_respacing, noise=self.static_noise)
        th.cuda.current_stream().wait_stream(s)
        th.cuda.synchronize()
        self._weights_version = self._version()
        self.graph = th.cuda.CUDAGraph()
        model._cond_key = None
        l0 = model.engine.launch_count()
        with th.cuda.graph(self.graph):
            self.static_out = diffusion.sample_many(model, self.static_batch, num_samples, timestep_respacing,
                                                    noise=self.static_noise)
        self.launches_per_replay = model.engine.launch_count() - l0
        # what the captured kernels read but the graph itself does not rewrite: workspace addresses, the slot tables of
        # this (images x samples) layout and the folded timestep embeddings of this schedule
        self._alloc_epoch = model.engine.alloc_epoch()
        self._bodies_key, self._temb_key = model._bodies_key, model._temb_key
        self._iob = np.repeat(np.arange(batch["img"].shape[0], dtype=np.int32), num_samples)
        model._cond_key = None   # the eager cache must not believe it has seen a later batch

    # ---------------------------------------------------------------- pipelined host -> device input staging
    def stage(self, batch):
        """Start copying `batch` (typically pinned host tensors) into a device-side staging copy of the inputs on a
        separate copy stream, so that the transfer of batch i+1 overlaps the replay of batch i.  Consume it with
        `sampler(staged=True)`."""
        if self._staging is None:
            self._staging = self._clone({k: v for k, v in self.static_batch.items() if k not in ("x_t", "vis_mask_smpl")})
            self._copy_stream = th.cuda.Stream()
            self._staged_evt, self._consumed_evt = th.cuda.Event(), th.cuda.Event()
            self._consumed_evt.record()
        self._copy_stream.wait_event(self._consumed_evt)     # the previous staging content has been moved on
        with th.cuda.stream(self._copy_stream):
            for path in self._paths:
                _get(self._staging, path).copy_(_get(batch, path), non_blocking=True)
            self._staged_evt.record()

    def check_overflow(self):
        """True if an fp16 operand overflowed in any replay since the last check (synchronises; replays themselves
        cannot read the flag back — in the default `overflow_check='deferred'` mode every replay ends with a copy of the
        flag to pinned memory, which `model.poll_overflow()` examines without blocking)."""
        return self.model.engine.check_overflow()

    def _version(self):
        return self.model._weights_version()

    def __call__(self, batch=None, noise=None, staged=False):
        """Replay the pass on `batch` (same shapes as at capture; None = reuse the static inputs as they are;
        `staged=True` = the batch handed to the last `stage()` call)."""
        if staged:
            if self._staging is None:
                raise RuntimeError("stage(batch) must be called before sampler(staged=True)")
            th.cuda.current_stream().wait_event(self._staged_evt)
            batch = self._staging
        if (noise is not None) != (self.static_noise is not None) and noise is not None:
            raise ValueError("this sampler was captured with torch's generator as the noise source (external_noise=False)")
        if noise is not None:
            self.static_noise.copy_(noise, non_blocking=True)
        if self.model._weights_dirty or self._version() != self._weights_version:
            raise RuntimeError("model weights changed after the graph was captured; build a new GraphedSampler")
        eng = self.model.engine
        if eng.alloc_epoch() != self._alloc_epoch:
            raise RuntimeError("the library reallocated a workspace (a larger problem ran on this context) after the "
                               "graph was captured; build a new GraphedSampler")
        if self.model._bodies_key != self._bodies_key or eng.n_bodies != self._iob.shape[0]:
            eng.set_bodies(self._iob)          # another layout ran eagerly in between: restore this one's slot tables
            self.model._bodies_key = self._bodies_key
            if eng.alloc_epoch() != self._alloc_epoch:
                raise RuntimeError("workspace reallocated while restoring the captured layout; build a new GraphedSampler")
        if self.model._temb_key != self._temb_key:
            self.model.set_timesteps(self.diffusion.timestep_map)
        if batch is not None and batch is not self.static_batch:
            for path in self._paths:
                src, dst = _get(batch, path), _get(self.static_batch, path)
                if src.shape != dst.shape:
                    raise ValueError(f"batch['{'/'.join(path)}'] has shape {tuple(src.shape)}, captured {tuple(dst.shape)}")
                dst.copy_(src, non_blocking=True)
            if staged:
                self._consumed_evt.record()
        self.graph.replay()
        if self.model._ovf_event is not None:
            self.model._ovf_event.record()    # the replay ended with a copy of the operand-overflow flag to pinned memory
        return self.static_out

"""diffusion/respace.py of the reference: timestep sub-sampling."""
import numpy as np

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """respace.py:8-61: which original steps a respaced sampler visits ("ddimN" = the first integer stride giving
    exactly N steps; otherwise evenly strided sections)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                steps = range(0, num_timesteps, stride)
                if len(steps) == want:
                    return set(steps)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    start, kept = 0, []
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0  # accumulated in floating point like the reference, so ties round identically
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """respace.py:64-114: keep `use_timesteps` of a base process, re-deriving betas from the alpha-bar ratios."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base_ac = np.cumprod(1.0 - np.array(kwargs["betas"], dtype=np.float64), axis=0)
        self.timestep_map, new_betas, last = [], [], 1.0
        for i, ac in enumerate(base_ac):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def _scale_timesteps(self, t):
        """respace.py:111-113: scaling is done by the wrapped model, after the index map."""
        return t

    def _map_timesteps(self, t):
        """_WrappedModel.__call__ (respace.py:124-129): respaced index -> original timestep, then (if rescale_timesteps)
        scaled by 1000 / original_num_steps."""
        import torch
        new_ts = torch.tensor(self.timestep_map, device=t.device, dtype=t.dtype)[t]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return new_ts

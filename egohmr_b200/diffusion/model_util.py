"""diffusion/model_util.py of the reference: the factory the drivers call (test_egohmr.py:122-123)."""
from . import gaussian_diffusion as gd
from .respace import SpacedDiffusion, space_timesteps


def create_gaussian_diffusion(num_diffusion_timesteps=1000, timestep_respacing="ddim5", body_rep_mean=None,
                              body_rep_std=None):
    """model_util.py:4-22: cosine schedule, no beta scaling, rescale_timesteps=False."""
    steps = num_diffusion_timesteps
    betas = gd.get_named_beta_schedule("cosine", steps, 1.0)
    if not timestep_respacing:
        timestep_respacing = [steps]
    return SpacedDiffusion(use_timesteps=space_timesteps(steps, timestep_respacing), betas=betas,
                           rescale_timesteps=False, body_rep_mean=body_rep_mean, body_rep_std=body_rep_std)

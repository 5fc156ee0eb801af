"""Noise schedule, timestep respacing and single-step sampler updates (numpy).  TEST INFRASTRUCTURE (see oracle/__init__).

Follows diffusion/gaussian_diffusion.py and diffusion/respace.py of the reference line by line; the float64 tables
are bit-identical to the reference's (same numpy/math calls in the same order) and the fp32 updates replay the
reference's torch op order.
"""
import math

import numpy as np


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """gaussian_diffusion.py:49-66."""
    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(betas)


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, scale_betas=1.0):
    """gaussian_diffusion.py:22-46."""
    if schedule_name == "linear":
        scale = scale_betas * 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def space_timesteps(num_timesteps, section_counts):
    """respace.py:8-61."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired_count = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired_count:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start_idx = 0
    all_steps = []
    for i, section_count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < section_count:
            raise ValueError(f"cannot divide section of {size} steps into {section_count}")
        frac_stride = 1 if section_count <= 1 else (size - 1) / (section_count - 1)
        cur_idx = 0.0
        taken = []
        for _ in range(section_count):
            taken.append(start_idx + round(cur_idx))
            cur_idx += frac_stride
        all_steps += taken
        start_idx += size
    return set(all_steps)


class Schedule:
    """The float64 tables of GaussianDiffusion.__init__ (gaussian_diffusion.py:122-169) after SpacedDiffusion
    re-derived the betas of the kept steps (respace.py:73-87)."""

    def __init__(self, num_diffusion_timesteps=1000, timestep_respacing="ddim5"):
        base_betas = get_named_beta_schedule("cosine", num_diffusion_timesteps, 1.0)  # model_util.py:8-12
        if not timestep_respacing:
            timestep_respacing = [num_diffusion_timesteps]
        use = space_timesteps(num_diffusion_timesteps, timestep_respacing)
        base_ac = np.cumprod(1.0 - np.array(base_betas, dtype=np.float64), axis=0)
        last = 1.0
        new_betas, self.timestep_map = [], []
        for i, ac in enumerate(base_ac):
            if i in use:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        betas = np.array(np.array(new_betas), dtype=np.float64)
        assert (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)


def _ex(arr, t):
    """_extract_into_tensor (gaussian_diffusion.py:784-797): the float64 table entry cast with .float().  The
    coefficient tensors are fp32 whatever dtype the sampler state has; torch promotes them only when they meet a
    float64 tensor, so every coefficient-only expression below is evaluated in fp32 first."""
    return np.float32(arr[t])


def ddim_coefficients(sch, t):
    """The four per-step scalars of ddim_sample with eta = 0 (gaussian_diffusion.py:537-551), fp32 like torch."""
    one = np.float32(1.0)
    alpha_bar = _ex(sch.alphas_cumprod, t)
    alpha_bar_prev = _ex(sch.alphas_cumprod_prev, t)
    sigma = np.float32(0.0) * np.sqrt((one - alpha_bar_prev) / (one - alpha_bar)) * np.sqrt(one - alpha_bar / alpha_bar_prev)
    return (_ex(sch.sqrt_recip_alphas_cumprod, t), _ex(sch.sqrt_recipm1_alphas_cumprod, t), np.sqrt(alpha_bar_prev),
            np.sqrt(one - alpha_bar_prev - sigma ** 2))


def ddim_update(sch, x, x0, t, dtype=np.float32):
    """ddim_sample with eta = 0 (gaussian_diffusion.py:511-556) given pred_xstart = x0; t is the respaced index."""
    c_recip, c_recipm1, s_abp, s_1mabp = (dtype(c) for c in ddim_coefficients(sch, t))
    eps = (c_recip * x - x0) / c_recipm1  # :286-290
    mean_pred = x0 * s_abp + s_1mabp * eps
    return mean_pred.astype(dtype)  # + nonzero_mask * sigma * noise == + 0


def ddpm_coefficients(sch, t, guided=False, cond_grad_weight=1.0):
    """coef1, coef2, (t != 0) * exp(0.5 * log_var), gradient scale — fp32 like torch (gaussian_diffusion.py:220-223,
    260-262, 336, 378-385)."""
    nonzero = np.float32(1.0 if t != 0 else 0.0)
    std = np.exp(np.float32(0.5) * _ex(sch.posterior_log_variance_clipped, t))
    gscale = np.float32(0.0)
    if guided and t <= 10:
        if t >= 5:
            gscale = np.float32(cond_grad_weight) * _ex(sch.posterior_variance, t)
        else:
            gscale = np.float32(cond_grad_weight * 0.01)
    return _ex(sch.posterior_mean_coef1, t), _ex(sch.posterior_mean_coef2, t), np.float32(nonzero * std), gscale


def ddpm_update(sch, x, x0, t, noise, grad=None, cond_grad_weight=1.0, dtype=np.float32):
    """p_sample / p_sample_with_grad (gaussian_diffusion.py:298-388) given pred_xstart = x0."""
    c1, c2, nz_std, gscale = (dtype(c) for c in ddpm_coefficients(sch, t, grad is not None, cond_grad_weight))
    mean = c1 * x0 + c2 * x  # :220-223
    if grad is not None and t <= 10:  # :378-385
        mean = mean + gscale * grad
    return (mean + nz_std * noise).astype(dtype)

"""Schedules and small host helpers of the sampling path.  Host code: these run once per
`sample()` call on the CPU and must reproduce the reference's fp32 values bit for bit
(ldm/modules/diffusionmodules/util.py:30-83, 160-180, 208-258)."""
import math

import numpy as np
import torch


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """util.py:30-52.  Only the schedules the configs can name are provided."""
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    elif schedule == "cosine":
        ts = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        alphas = torch.cos(ts / (1 + cosine_s) * np.pi / 2).pow(2)
        alphas = alphas / alphas[0]
        betas = torch.clamp(1 - alphas[1:] / alphas[:-1], min=0, max=0.999)
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=False):
    """util.py:55-70: t_i = 1 + i * (T // S)."""
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        steps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        steps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    return steps + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=False):
    """util.py:73-83.  alphacums: fp32 CPU tensor; returns (sigmas, alphas, alphas_prev)."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0].item()] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas.numpy()) * (1 - alphas.numpy() / alphas_prev))
    return sigmas, alphas, alphas_prev


def extract_into_tensor(a, t, x_shape):
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def timestep_embedding_host(timesteps, dim, max_period=10000):
    """fp32 restatement of util.py:160-180 for host-side use (tests); the device path is
    idiff_timestep_embedding."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class FourierEmbedder:
    """util.py:12-26 -- kept for API compatibility; holds only the frequency table.  The
    arithmetic is idiff_fourier_embed."""

    def __init__(self, num_freqs=64, temperature=100):
        self.num_freqs = num_freqs
        self.temperature = temperature
        self.freq_bands = temperature ** (torch.arange(num_freqs) / num_freqs)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module

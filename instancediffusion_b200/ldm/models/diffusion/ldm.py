"""LatentDiffusion -- the `target:` of configs/test_*.yaml:64 (reference: ldm/models/diffusion/ldm.py:10-19).

On the sampling path this object only carries the beta / alpha-cumprod buffers the PLMS samplers read
(`_plms_common.PLMSBase.make_schedule`); the forward-noising helper is kept for interface parity."""
import torch

from .ddpm import DDPM


class LatentDiffusion(DDPM):
    clip_denoised = False  # (the reference hard-codes it in __init__, ldm.py:14)

    def q_sample(self, x_start, t, noise=None):
        """x_t = sqrt(acp_t) x_0 + sqrt(1 - acp_t) eps, per-sample timesteps `t` (ldm.py:16-19)."""
        eps = torch.randn_like(x_start) if noise is None else noise
        bshape = (x_start.shape[0],) + (1,) * (x_start.dim() - 1)
        signal = self.sqrt_alphas_cumprod.gather(0, t).reshape(bshape)
        sigma = self.sqrt_one_minus_alphas_cumprod.gather(0, t).reshape(bshape)
        return signal * x_start + sigma * eps

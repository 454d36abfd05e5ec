"""First-stage model of the sampling pipeline: `AutoencoderKL.decode` turns the denoised latent into the
image right after the sampler loop (inference.py:96).  Same class name, constructor signature and
state_dict keys as the reference's ldm/models/autoencoder.py:12-37 (loaded strict by
utils/checkpoint.py:233), with the arithmetic in libidiff_b200.so.

decode(z) = Decoder(post_quant_conv(z / scale_factor)): the scale, the 1x1 post_quant_conv and the
NCHW fp32 -> NHWC fp16 conversion are one kernel (idiff_vae_latent_in); everything after it is the GEMM /
GroupNorm kernels of the UNet (ldm/modules/diffusionmodules/model.py in this package).
"""
from __future__ import annotations

import torch

from ... import ops
from ..modules._base import PackedModule, f32
from ..modules.diffusionmodules.model import Decoder, Encoder


class DiagonalGaussianDistribution(object):
    """ldm/modules/distributions/distributions.py:24-41 (the part `encode` uses)."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self):
        return self.mean + self.std * torch.randn(self.mean.shape).to(device=self.parameters.device)

    def mode(self):
        return self.mean


class AutoencoderKL(PackedModule):
    def __init__(self, ddconfig, embed_dim, scale_factor=1):
        super().__init__()
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        assert ddconfig["double_z"]
        self.quant_conv = torch.nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = torch.nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim
        self.scale_factor = scale_factor

    def _pack(self):
        zc = self.post_quant_conv.out_channels
        return {"wpq": f32(self.post_quant_conv.weight).reshape(zc, self.embed_dim).contiguous(),
                "bpq": f32(self.post_quant_conv.bias)}

    @torch.no_grad()
    def encode(self, x):
        """autoencoder.py:27-31 (not on the sampling path; provided for completeness of the module)."""
        h = self.encoder(x)
        moments = torch.nn.functional.conv2d(h, self.quant_conv.weight.float(), self.quant_conv.bias.float())
        return DiagonalGaussianDistribution(moments).sample() * self.scale_factor

    @torch.no_grad()
    def decode(self, z):
        """autoencoder.py:33-37: z (B, 4, h, w) -> image (B, 3, 8h, 8w) fp32 in [-1, 1] (nominally)."""
        if self.embed_dim != self.post_quant_conv.out_channels:
            raise NotImplementedError("post_quant_conv with embed_dim != z_channels is not used by the shipped configs")
        p = self.pk()
        B, _, H, W = z.shape
        z16 = ops.vae_latent_in(z.float().contiguous(), p["wpq"], p["bpq"], 1.0 / float(self.scale_factor))
        return self.decoder._decode_tokens(z16, B, H, W)

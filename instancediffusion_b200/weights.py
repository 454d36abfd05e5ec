"""Deterministic synthetic weights + model construction helpers.

The released InstanceDiffusion / SD1.5 checkpoints are network downloads (README.md:120 of the
reference) and are not available offline, so parity runs and the benchmark use seeded random
weights.  Every tensor is generated independently from (seed, crc32(key)) with the CPU generator,
so the same state_dict is reproduced bit for bit on any box -- the oracle (reference modules, CPU
fp32) and the CUDA path load identical values without shipping 4.9 GB.

Scales are chosen so activations stay O(1) through ~200 layers: N(0, 1/fan_in) for matrices,
1 + 0.1 N for norm gains, small biases; the 303 tensors the reference zero-initialises (proj_out,
ResBlock out conv, out.2, alpha_*, scaleu_*, null features -- SURVEY.md appendix B) get non-zero
values, otherwise eps == 0 identically and every parity test would be vacuous.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import torch

UNET_CONFIG = dict(
    image_size=64, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
    num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, transformer_depth=1, context_dim=768,
    fuser_type="gatedSA", use_checkpoint=True, sd_v1_5=True, efficient_attention=True,
)

# configs/test_*.yaml:36-39 of the reference: the test-time modality drops of each shipped config
UNIFUSION_FLAGS = {
    "box": dict(test_drop_scribbles=True, test_drop_masks=True),
    "point": dict(test_drop_boxes=True, test_drop_scribbles=True, test_drop_masks=True),
    "scribble": dict(test_drop_scribbles=False, test_drop_masks=False),
    "mask": dict(test_drop_scribbles=True, test_drop_masks=False),
}


def unet_config(flavor: str = "box", tokenizer_target: str | None = None) -> dict:
    cfg = dict(UNET_CONFIG)
    target = tokenizer_target or "instancediffusion_b200.ldm.modules.diffusionmodules.text_grounding_net.UniFusion"
    cfg["grounding_tokenizer"] = dict(target=target, params=dict(in_dim=768, out_dim=768, mid_dim=3072,
                                                                 **UNIFUSION_FLAGS[flavor]))
    return cfg


def synth_tensor(key: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    r = torch.randn(shape, generator=g, dtype=torch.float32) if len(shape) else torch.randn((), generator=g)
    if leaf.startswith("alpha_"):
        return r * 0.7                       # tanh(alpha) gates of the fusers
    if leaf.startswith("scaleu_"):
        return r * 0.4
    if leaf.startswith("null_"):
        return r                             # learned null features, comparable to the embeddings
    if leaf == "pos_embedding":
        return r * 0.5
    if leaf == "gamma":
        return r * 0.1
    if leaf == "bias":
        return r * 0.05
    if leaf == "weight":
        if len(shape) <= 1:
            return 1.0 + 0.1 * r             # GroupNorm / LayerNorm gains
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        return r * (1.0 / fan_in ** 0.5)
    return r * 0.1


def synth_state_dict(schema: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, tuple(s), seed) for k, s in schema}


def load_synthetic(module: torch.nn.Module, seed: int = 0, prefix: str = "") -> None:
    """Fill `module` (on any device) with the synthetic weights for its own state_dict keys."""
    sd = module.state_dict()
    new = {}
    for k, v in sd.items():
        new[k] = synth_tensor(prefix + k, tuple(v.shape), seed).to(v.dtype)
    module.load_state_dict(new, strict=True)


def build_unet(flavor: str = "box", device: str | torch.device = "cuda", seed: int | None = 0):
    """Construct the mirror UNetModel without the (slow, useless) default init, move it to
    `device` and, unless seed is None, fill it with the synthetic weights."""
    from .ldm.modules.diffusionmodules.openaimodel import UNetModel
    from .grounding_input.text_grounding_tokinzer_input import GroundingNetInput

    with torch.device("meta"):
        model = UNetModel(**unet_config(flavor))
    model = model.to_empty(device=device)
    model.eval()
    if seed is not None:
        sd = {}
        for k, v in model.state_dict().items():
            sd[k] = synth_tensor(k, tuple(v.shape), seed)
        model.load_state_dict(sd, strict=True)
    model.grounding_tokenizer_input = GroundingNetInput()
    return model

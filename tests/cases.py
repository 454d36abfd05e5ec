"""Seeded parity cases shared by (1) oracle/make_golden.py, which runs them through the reference's
own modules on the CPU and commits the outputs under tests/golden/, (2) the CPU tests of the
plain-torch restatement (oracle/torch_oracle.py) and (3) the `-m gpu` tests of the CUDA modules.

A case names a reference class (by its import path inside `ldm/...`), constructor arguments and
seeded inputs; because the mirror classes keep the reference's signatures, the *same* driver code
builds and calls either implementation.
"""
from __future__ import annotations

import zlib
from typing import Dict

import torch

WEIGHT_SEED = 1


def synth_input(case: str, key: str, shape, scale: float = 1.0) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(zlib.crc32(f"{case}/{key}".encode()) & 0x7FFFFFFF)
    return torch.randn(tuple(shape), generator=g) * scale


# name -> spec.  `module` is "<python module under ldm.modules>:<class>".
MODULE_CASES: Dict[str, dict] = {
    "gated_sa_L0": dict(module="attention:GatedSelfAttentionDense", args=(320, 768, 8, 40),
                        kwargs=dict(efficient_attention=True),
                        inputs=dict(x=(2, 256, 320), objs=(2, 184, 768))),
    "gated_sa_L2": dict(module="attention:GatedSelfAttentionDense", args=(1280, 768, 8, 160),
                        kwargs=dict(efficient_attention=True),
                        inputs=dict(x=(2, 64, 1280), objs=(2, 184, 768))),
    "self_attn_L1": dict(module="attention:SelfAttention", args=(640,), kwargs=dict(heads=8, dim_head=80, efficient_attention=True),
                         inputs=dict(x=(2, 256, 640))),
    "cross_attn_L1": dict(module="attention:CrossAttention", args=(640, 768, 768),
                          kwargs=dict(heads=8, dim_head=80, efficient_attention=True),
                          inputs=dict(x=(2, 256, 640), key=(2, 77, 768)), same_kv=True),
    "ff_L0": dict(module="attention:FeedForward", args=(320,), kwargs=dict(glu=True), inputs=dict(x=(2, 256, 320))),
    "btb_L1": dict(module="attention:BasicTransformerBlock", args=(640, 768, 768, 8, 80, "gatedSA"),
                   kwargs=dict(efficient_attention=True),
                   inputs=dict(x=(2, 256, 640), context=(2, 77, 768), objs=(2, 184, 768))),
    "st_L0": dict(module="attention:SpatialTransformer", args=(320, 768, 768, 8, 40),
                  kwargs=dict(depth=1, fuser_type="gatedSA", efficient_attention=True),
                  inputs=dict(x=(2, 320, 16, 16), context=(2, 77, 768), objs=(2, 184, 768))),
    "resblock_same": dict(module="diffusionmodules.openaimodel:ResBlock", args=(320, 1280, 0),
                          kwargs=dict(out_channels=320), inputs=dict(x=(2, 320, 16, 16), emb=(2, 1280))),
    "resblock_skip": dict(module="diffusionmodules.openaimodel:ResBlock", args=(960, 1280, 0),
                          kwargs=dict(out_channels=640), inputs=dict(x=(2, 960, 8, 8), emb=(2, 1280))),
    "upsample": dict(module="diffusionmodules.openaimodel:Upsample", args=(320, True),
                     kwargs=dict(out_channels=320), inputs=dict(x=(2, 320, 8, 8))),
    "downsample": dict(module="diffusionmodules.openaimodel:Downsample", args=(320, True),
                       kwargs=dict(out_channels=320), inputs=dict(x=(2, 320, 16, 16))),
}

# Fourier_filter(x, threshold=1, scale) cases (openaimodel.py:25-48): shape, scale
FOURIER_CASES = {
    "fourier_pow2": dict(shape=(2, 64, 16, 16), scale=1.3),
    "fourier_npow2": dict(shape=(1, 64, 12, 12), scale=0.6),
}

# UniFusion cases: flavor (test-time drop flags of configs/test_<flavor>.yaml), batch, instances
UNIFUSION_CASES = {
    "unifusion_box": dict(flavor="box", batch=2, n=3, seed=5),
    "unifusion_point": dict(flavor="point", batch=2, n=3, seed=6),
    "unifusion_scribble": dict(flavor="scribble", batch=2, n=3, seed=7),
}
# mask flavour: 256-point polygons + box-shaped binary `segs` through the ConvNeXt mask encoder
# (text_grounding_net.py:226-231, 277-287); fixture unifusion_mask.pt
UNIFUSION_MASK_CASES = {
    "unifusion_mask": dict(flavor="mask", batch=2, n=3, seed=8),
}
# ConvNeXt pieces (convnext.py:15-123); fixture convnext.pt.  `fn` names a module-level callable.
CONVNEXT_CASES = {
    "convnext_block96": dict(cls="Block", args=(96,), inputs=dict(x=(2, 96, 16, 16))),
    "convnext_block768": dict(cls="Block", args=(768,), inputs=dict(x=(1, 768, 8, 8))),
    "convnext_tiny": dict(cls="ConvNeXt", args=(), inputs=dict(x=(1, 3, 64, 64))),
}

# whole-UNet cases (B=1, 64x64 latent, box flavour, 2 instances)
UNET_CASE = dict(flavor="box", batch=1, n=2, seed=11, t=601, weight_seed=0)

# sampler cases (config 1 of BASELINE.json and a short plain-PLMS run)
SAMPLER_CASES = {
    "plms_S4": dict(S=4, n=1, mis=0.0, batch=1, seed=21, guidance=7.5, alpha_type=[0.8, 0.0, 0.2]),
    "mis_S10": dict(S=10, n=1, mis=0.36, batch=1, seed=22, guidance=7.5, alpha_type=[0.8, 0.0, 0.2]),
}

# round-2 additions (fixtures unet_extra.pt / samplers_extra.pt): the configuration bench.py runs
# (config 2: forward batch 4+4, 8 instances, 50 steps), a Multi-instance Sampler run that merges
# more than two latents, the point / scribble / mask flavours through the whole UNet, a 96x96 latent
# (768^2 images, config 4: non-power-of-two maps take the reference's fp32 FFT branch).
UNET_EXTRA_CASES = {
    "b4n8": dict(flavor="box", batch=4, n=8, seed=31, t=601, uncond=True),
    "point": dict(flavor="point", batch=1, n=2, seed=12, t=401),
    "scribble": dict(flavor="scribble", batch=1, n=2, seed=13, t=801),
    "mask": dict(flavor="mask", batch=1, n=2, seed=14, t=601),
    "lat96": dict(flavor="box", batch=1, n=2, seed=15, t=601, size=96),
}
SAMPLER_EXTRA_CASES = {
    "mis_S10_n3": dict(S=10, n=3, mis=0.36, batch=1, seed=23, guidance=7.5, alpha_type=[0.8, 0.0, 0.2]),
    "plms_S50_b4": dict(S=50, n=8, mis=0.0, batch=4, seed=24, guidance=7.5, alpha_type=[0.8, 0.0, 0.2]),
}


# instance-isolation attention mask (attention.py:187-255, efficient_attention=False, eval_local.py --use_masked_att):
# one gated self-attention block at the 64x64 level, att_masks rasterised from seeded boxes the way
# utils/input.py:34-37 does.  The fixture keeps every `stride`-th token row (the full output is 10 MB).
MASKED_CASES = {
    "gated_sa_masked": dict(args=(320, 768, 8, 40), batch=2, n=(3, 5), seed=51, stride=8),
}


def attmask_from_boxes(boxes, n, size=64, max_objs=30):
    """utils/input.py:34-37,79 restated in numpy: att_masks[k][x1:x2, y1:y2] = 1 with np.round (x on the first axis)."""
    import numpy as np
    att = np.zeros((max_objs, size, size), dtype=np.float32)
    for k in range(n):
        box = [float(v) for v in boxes[k]]
        x1, y1, x2, y2 = (int(np.round(box[0] * size)), int(np.round(box[1] * size)), int(np.round(box[2] * size)),
                          int(np.round(box[3] * size)))
        att[k][x1:x2, y1:y2] = 1
    return torch.from_numpy(att)


def masked_case_inputs(name: str, spec: dict):
    """x (B, 4096, 320), objs (B, 184, 768), boxes (B, 30, 4), counts, att_masks (B, 30, 64, 64)."""
    from instancediffusion_b200 import synthetic
    B = spec["batch"]
    x = synth_input(name, "x", (B, 64 * 64, spec["args"][0]))
    objs = synth_input(name, "objs", (B, 184, spec["args"][1]))
    boxes = torch.zeros((B, 30, 4))
    att = torch.zeros((B, 30, 64, 64))
    for b in range(B):
        lay = synthetic.make_layout(spec["n"][b], spec["seed"] + b, "box")
        boxes[b, :spec["n"][b]] = lay["boxes"]
        att[b] = attmask_from_boxes(lay["boxes"], spec["n"][b])
    return x, objs, boxes, torch.tensor(spec["n"], dtype=torch.int32), att


# first-stage model (ldm/models/autoencoder.py): latent std 0.9 is what a finished sampler run hands to decode
VAE_CASES = {
    "decode_32": dict(kind="decode", batch=1, size=32, seed=41, std=0.9),
    "decode_b2_16": dict(kind="decode", batch=2, size=16, seed=42, std=0.9),
    "encode_64": dict(kind="encode", batch=1, size=64, seed=43, std=0.5),
}


# CLIP text encoder (host prep): full-size text tower, synthetic weights, seeded token ids laid out
# like the tokenizer's output: BOS, `n` word tokens, EOS, then EOS padding (openai/clip pads with <|endoftext|>)
CLIP_CASES = {
    "clip_b3": dict(lengths=(5, 20, 75), seed=71),
}
CLIP_BOS, CLIP_EOS = 49406, 49407


def clip_token_ids(spec: dict) -> torch.Tensor:
    g = torch.Generator().manual_seed(spec["seed"])
    rows = []
    for n in spec["lengths"]:
        words = torch.randint(0, CLIP_BOS, (n,), generator=g)
        rows.append(torch.cat([torch.tensor([CLIP_BOS]), words, torch.full((77 - 1 - n,), CLIP_EOS)]))
    return torch.stack(rows).long()


def build_inputs(name: str, spec: dict) -> Dict[str, torch.Tensor]:
    ins = {k: synth_input(name, k, shp) for k, shp in spec["inputs"].items()}
    if spec.get("same_kv"):
        ins["value"] = ins["key"]
    return ins


def run_module_case(name: str, spec: dict, cls, device="cpu", dtype=torch.float32, prefix_loader=None):
    """Build `cls` (reference or mirror), load the case's synthetic weights, call forward."""
    from instancediffusion_b200.weights import load_synthetic
    mod = cls(*spec["args"], **spec.get("kwargs", {}))
    load_synthetic(mod, WEIGHT_SEED, prefix=name + ".")
    mod = mod.to(device).eval()
    ins = {k: v.to(device=device, dtype=dtype) for k, v in build_inputs(name, spec).items()}
    with torch.no_grad():
        if "key" in ins:
            kv = ins["key"]
            out = mod(ins["x"], kv, kv)
        else:
            out = mod(**ins)
    return out

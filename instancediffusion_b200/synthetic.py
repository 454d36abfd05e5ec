"""Synthetic sampling workloads: seeded layouts with the tensor layout of
the reference's `prepare_batch` (utils/input.py:41-125) -- every modality padded to 30 slots --
plus CLIP-shaped random context.  Generated on the CPU generator (bit-reproducible across boxes),
then moved to `device`.  Shared by bench.py, the tests and the oracle's golden script.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

MAX_OBJS = 30
TEXT_NORM = 28.7  # norm of CLIP pooled phrase features cited by the reference (utils/model.py:141)


def _gen(seed: int) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed(seed)


def make_layout(n: int, seed: int, flavor: str = "box", seg_size: int = 512):
    """One image's instances: boxes xyxy in [0,1], centre points, 20-point scribbles, 256-point
    polygons (mask flavour only), unit-norm*28.7 phrase features."""
    g = _gen(seed)
    xy0 = torch.rand((n, 2), generator=g) * 0.6
    wh = 0.1 + torch.rand((n, 2), generator=g) * 0.3
    xy1 = (xy0 + wh).clamp(max=1.0)
    boxes = torch.cat([xy0, xy1], -1)
    points = (xy0 + xy1) / 2
    text = torch.randn((n, 768), generator=g)
    text = text / text.norm(dim=-1, keepdim=True) * TEXT_NORM

    def pts_in_box(k):
        u = torch.rand((n, k, 2), generator=g)
        p = xy0[:, None, :] + u * (xy1 - xy0)[:, None, :]
        order = (p ** 2).sum(-1).argsort(dim=1)  # sorted by distance to the origin (decode_item.py:102-108)
        return torch.gather(p, 1, order[..., None].expand(-1, -1, 2)).reshape(n, 2 * k)

    scribbles = pts_in_box(20) if flavor in ("scribble",) else torch.zeros((n, 40))
    polygons = pts_in_box(256) if flavor == "mask" else torch.zeros((n, 512))
    return dict(boxes=boxes, points=points, text=text, scribbles=scribbles, polygons=polygons)


def make_grounding_batch(batch: int, n: int, seed: int, flavor: str = "box", device="cpu",
                         only_instance: int | None = None, seg_size: int = 512) -> Dict[str, torch.Tensor]:
    """The dict `GroundingNetInput.prepare` consumes (keys of utils/input.py:81-89), repeated
    `batch` times like the reference does for `num_images` samples of one layout.
    only_instance=k builds the single-instance input of the Multi-instance Sampler
    (prepare_instance_meta, utils/input.py:130-144): instance k alone in slot 0."""
    lay = make_layout(n, seed, flavor)
    idx = list(range(n)) if only_instance is None else [only_instance]
    m = len(idx)
    boxes = torch.zeros((MAX_OBJS, 4))
    masks = torch.zeros((MAX_OBJS,))
    text = torch.zeros((MAX_OBJS, 768))
    scribbles = torch.zeros((MAX_OBJS, 40))
    polygons = torch.zeros((MAX_OBJS, 512))
    points = torch.zeros((MAX_OBJS, 2))
    boxes[:m] = lay["boxes"][idx]
    masks[:m] = 1
    text[:m] = lay["text"][idx]
    scribbles[:m] = lay["scribbles"][idx]
    polygons[:m] = lay["polygons"][idx]
    points[:m] = lay["points"][idx]
    if flavor == "mask":
        # mask conditioning (configs/test_mask.yaml): box-shaped binary masks, one (S,S) plane per slot
        # (what decode_item.py hands to prepare_batch when the :249 mask discard is fixed)
        seg = torch.zeros((MAX_OBJS, seg_size, seg_size))
        for j, k in enumerate(idx):
            x0, y0, x1, y1 = [int(round(float(v) * seg_size)) for v in lay["boxes"][k]]
            seg[j, y0:max(y1, y0 + 1), x0:max(x1, x0 + 1)] = 1.0
        segs = seg.unsqueeze(0).repeat(batch, 1, 1, 1)
    else:
        # segs: all zero (inference.py:249 discards decoded masks, so this is what the CLI feeds);
        # a zero (1,1)-strided view keeps the (B,30,S,S) shape without 30 MiB per sample
        segs = torch.zeros((batch, MAX_OBJS, 1, 1)).expand(batch, MAX_OBJS, seg_size, seg_size)
    rep = lambda t: t.unsqueeze(0).repeat(batch, *([1] * t.dim())).to(device)
    return {
        "boxes": rep(boxes), "masks": rep(masks), "text_masks": rep(masks), "text_embeddings": rep(text),
        "scribbles": rep(scribbles), "polygons": rep(polygons), "points": rep(points),
        "segs": segs.to(device),
    }


def make_context(batch: int, seed: int, device="cpu", tokens: int = 77, dim: int = 768) -> torch.Tensor:
    """(B,77,768) stand-in for FrozenCLIPEmbedder output; one prompt repeated over the batch."""
    c = torch.randn((1, tokens, dim), generator=_gen(seed))
    return c.repeat(batch, 1, 1).to(device)


def make_noise(batch: int, seed: int, size: int = 64, device="cpu") -> torch.Tensor:
    """inference.py:300-301: torch.manual_seed(seed); randn(num_images,4,64,64) on CPU then .to(device)."""
    return torch.randn((batch, 4, size, size), generator=_gen(seed)).to(device)


def make_sampler_inputs(grounding_tokenizer_input, batch: int, n: int, seed: int, flavor: str = "box",
                        mis: bool = False, device="cpu", size: int = 64) -> Tuple[object, torch.Tensor]:
    """Build what inference.py:76-92 hands to `sampler.sample`: a single input dict (mis=False) or
    the list [global, inst_1..inst_n] (mis=True), and the unconditional context `uc`."""
    x = make_noise(batch, seed, size, device)
    uc = make_context(batch, seed + 7, device)

    def one(only, ctx_seed):
        gb = make_grounding_batch(batch, n, seed, flavor, device, only_instance=only)
        gi = grounding_tokenizer_input.prepare(gb)
        return dict(x=x, timesteps=None, context=make_context(batch, ctx_seed, device), grounding_input=gi)

    glob = one(None, seed + 11)
    if not mis:
        return glob, uc
    return [glob] + [one(k, seed + 100 + k) for k in range(n)], uc

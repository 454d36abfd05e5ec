"""Host prep of one sampling request, built directly on the device -- the tensor layout of
the reference's `prepare_batch` (utils/input.py:41-125) and `prepare_instance_meta` (:128-144):

  boxes (B,30,4)  masks (B,30)  text_masks (B,30)  text_embeddings (B,30,768)  polygons (B,30,512)
  scribbles (B,30,40)  segs (B,30,512,512)  points (B,30,2)  [att_masks (B,30,64,64)]  [instance_meta: the same
  per instance, instance i alone in slot 0]

Differences in mechanism, not in values: tensors are created on `device` (the reference builds them on the host and
copies ~30 MiB of `segs` per sample); all-zero `segs` are a stride-0 view; the attention masks are rasterised by
idiff_boxes_to_attmask.  The CLIP phrase features are outside the hot path: pass them in (`text_features`, one
(768,) tensor or None per phrase) or pass the reference's `get_clip_feature`-style callable as `encode_phrase`.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from .. import ops

N_SCRIBBLE_POINTS = 20   # utils/input.py:43
N_POLYGON_POINTS = 256   # utils/input.py:44
SEG_SIZE = 512


def complete_mask(has_mask, max_objs, device="cpu"):
    """utils/input.py:21-31."""
    mask = torch.ones(1, max_objs, device=device)
    if has_mask is None:
        return mask
    if type(has_mask) == int or type(has_mask) == float:
        return mask * has_mask
    for idx, value in enumerate(has_mask):
        mask[0, idx] = value
    return mask


def _rows(values: Sequence, width: int, max_objs: int, device) -> torch.Tensor:
    """One row per instance (None -> zeros), zero-padded to max_objs."""
    out = torch.zeros((max_objs, width), dtype=torch.float32)
    for i, v in enumerate(values):
        if v is not None:
            out[i] = torch.as_tensor(np.asarray(v, dtype=np.float32)).reshape(-1)
    return out.to(device)


def _segs(segs, n: int, batch: int, max_objs: int, device) -> torch.Tensor:
    if segs is None or len(segs) == 0 or all(s is None for s in segs) or not np.any(np.asarray(segs)):
        # all-zero masks (what inference.py:249 produces for every shipped demo): no 30 MiB per sample
        return torch.zeros((batch, max_objs, 1, 1), device=device).expand(batch, max_objs, SEG_SIZE, SEG_SIZE)
    out = torch.zeros((max_objs, SEG_SIZE, SEG_SIZE), dtype=torch.float32)
    arr = np.asarray(segs, dtype=np.float32).reshape(-1, SEG_SIZE, SEG_SIZE)
    out[:arr.shape[0]] = torch.from_numpy(arr)
    return out.to(device).unsqueeze(0).repeat(batch, 1, 1, 1)


def _one(locations, text_features, polygons, scribbles, segs, points, text_mask, batch, max_objs, device):
    n = len(locations)
    boxes = _rows(locations, 4, max_objs, device)
    masks = torch.zeros(max_objs, device=device)
    masks[:n] = 1
    text = torch.zeros((max_objs, 768), device=device)
    text_masks = torch.zeros(max_objs, device=device)
    for i, f in enumerate(text_features):
        if f is not None:
            text[i] = f.to(device=device, dtype=torch.float32).reshape(-1)
            text_masks[i] = 1
    rep = lambda t: t.unsqueeze(0).repeat(batch, *([1] * t.dim()))
    return {
        "boxes": rep(boxes),
        "masks": rep(masks),
        "text_masks": rep(text_masks) * complete_mask(text_mask, max_objs, device),
        "text_embeddings": rep(text),
        "polygons": rep(_rows(polygons, N_POLYGON_POINTS * 2, max_objs, device)),
        "scribbles": rep(_rows(scribbles, N_SCRIBBLE_POINTS * 2, max_objs, device)),
        "segs": _segs(segs, n, batch, max_objs, device),
        "points": rep(_rows(points, 2, max_objs, device)),
    }


@torch.no_grad()
def prepare_batch(meta: dict, batch: int = 1, max_objs: int = 30, model=None, processor=None, image_size: int = 64,
                  use_masked_att: bool = False, device="cuda", text_features: Optional[List] = None,
                  encode_phrase: Optional[Callable] = None) -> dict:
    """utils/input.py:41-125.  `model` / `processor` are accepted for signature compatibility and handed to
    `encode_phrase(model, processor, phrase)` when given; otherwise `text_features` supplies the phrase features."""
    phrases = meta.get("phrases")
    n = len(meta["locations"])
    phrases = [None] * n if phrases is None else phrases
    if text_features is None:
        text_features = [encode_phrase(model, processor, p) if (encode_phrase is not None and p is not None) else None
                         for p in phrases]
    none = [None] * n
    out = _one(meta["locations"], text_features, meta.get("polygons") or none, meta.get("scribbles") or none,
               meta.get("segs"), meta.get("points") or none, meta.get("text_mask"), batch, max_objs, device)
    att = None
    if use_masked_att:
        counts = torch.full((1,), n, dtype=torch.int32, device=device)
        att = ops.boxes_to_attmask(out["boxes"][:1].contiguous(), counts, image_size)[0]  # (max_objs, S, S)
        out["att_masks"] = att.unsqueeze(0).repeat(batch, 1, 1, 1)
    if "instance_meta" in meta:
        out["instance_meta"] = []
        for i, im in enumerate(meta["instance_meta"]):
            one = _one(im["locations"][:1], [text_features[i]], im["polygons"][:1], im["scribbles"][:1],
                       None if im.get("segs") is None else np.asarray(im["segs"])[:1], im["points"][:1], im.get("text_mask"),
                       batch, max_objs, device)
            if use_masked_att:
                a = torch.zeros_like(att)
                a[0] = att[i]
                one["att_masks"] = a.unsqueeze(0).repeat(batch, 1, 1, 1)
            out["instance_meta"].append(one)
    return out


def prepare_instance_meta(test_info: dict, i: int, file_name=None, save_folder_name=None, ckpt=None) -> dict:
    """utils/input.py:128-144: the single-instance request of the Multi-instance Sampler."""
    return {
        "ckpt": test_info.get("ckpt", None),
        "phrases": [test_info["phrases"][i]],
        "locations": [test_info["locations"][i]],
        "polygons": [test_info["polygons"][i]],
        "segs": [test_info["segs"][i]] if test_info.get("segs") is not None and len(test_info["segs"]) else None,
        "scribbles": [test_info["scribbles"][i]],
        "points": [test_info["points"][i]],
        "alpha_type": test_info["alpha_type"],
        "prompt": test_info["phrases"][i],
        "file_name": file_name,
        "save_folder_name": save_folder_name,
    }

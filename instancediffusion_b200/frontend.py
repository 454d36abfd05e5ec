"""Demo-JSON front end: the request format of the reference's CLI (inference.py:188-281,
demos/*.json) turned into the `meta` dict `prepare_batch` consumes -- boxes rescaled to [0,1] xyxy, centre points,
scribbles, polygons, per-instance captions, the alpha schedule -- and, for the Multi-instance Sampler, one
single-instance meta per annotation (inference.py:283-292).

As shipped, the reference discards decoded instance masks before use (inference.py:249 re-initialises the list), so
every demo runs with zero `segs`, zero polygons and all-zero scribbles unless scribbles are given explicitly; that
behaviour is reproduced (`keep_masks=False`).  `keep_masks=True` feeds masks through (COCO RLE needs pycocotools,
which is only imported then).
"""
from __future__ import annotations

import json
from typing import List, Optional

import numpy as np

from .utils.input import N_POLYGON_POINTS, N_SCRIBBLE_POINTS, prepare_instance_meta


def rescale_box(bbox, width, height):
    """inference.py:133-138: xywh in pixels -> xyxy in [0, 1]."""
    return [bbox[0] / width, bbox[1] / height, (bbox[0] + bbox[2]) / width, (bbox[1] + bbox[3]) / height]


def get_point_from_box(bbox):
    return [(bbox[0] + bbox[2]) / 2.0, (bbox[1] + bbox[3]) / 2.0]  # inference.py:140-142


def rescale_points(point, width, height):
    return [point[0] / float(width), point[1] / float(height)]  # inference.py:144-145


def rescale_scribbles(scribbles, width, height):
    return [[s[0] / float(width), s[1] / float(height)] for s in scribbles]  # inference.py:147-148


def equally_spaced_sampling_with_replacement(points_list, sample_size):
    """dataset/decode_item.py:79-100."""
    if sample_size <= len(points_list):
        gap = len(points_list) // sample_size
        return [points_list[i * gap] for i in range(sample_size)]
    return [points_list[int(i * len(points_list) / sample_size) % len(points_list)] for i in range(sample_size)]


def reorder_scribbles(scribbles):
    """dataset/decode_item.py:102-108: order by distance to the origin, resample to 20 points, order again."""
    key = lambda p: float(np.linalg.norm(np.array(p)))
    scribbles = sorted(scribbles, key=key)
    scribbles = equally_spaced_sampling_with_replacement(scribbles, N_SCRIBBLE_POINTS)
    return sorted(scribbles, key=key)


def read_request(path_or_dict, alpha: float = 0.8, ckpt: Optional[str] = None, save_folder_name: Optional[str] = None,
                 mis: float = 0.36, keep_masks: bool = False) -> List[dict]:
    """-> meta_list as inference.py:264-292 builds it: [global meta] with `instance_meta` attached when mis > 0."""
    data = path_or_dict if isinstance(path_or_dict, dict) else json.load(open(path_or_dict))
    W, H = data["width"], data["height"]
    annos = data["annos"]
    boxes = [a.get("bbox", [0, 0, 0, 0]) for a in annos]
    locations = [rescale_box(b, W, H) for b in boxes]
    phrases = [a["caption"] for a in annos]
    pts = [a["point"] for a in annos if "point" in a]
    points = [get_point_from_box(b) for b in locations] if len(pts) == 0 else [rescale_points(p, W, H) for p in pts]
    scr = [a["scribble"] for a in annos if "scribble" in a]
    masks = None
    if keep_masks and any(a.get("mask") for a in annos):
        from pycocotools import mask as coco_mask  # only for requests that carry RLE masks
        masks = [coco_mask.decode([a["mask"]]).astype(bool) if a.get("mask") else np.zeros((512, 512, 1), bool) for a in annos]
    if len(scr) == 0:
        # sample_random_points_from_mask on an all-zero mask returns 2k zeros (decode_item.py:117-119)
        scribbles = [[0.0] * (2 * N_SCRIBBLE_POINTS) for _ in annos]
    else:
        # inference.py:256-257 calls reorder_scribbles on the LIST of per-instance scribbles (a reference quirk:
        # it sorts / resamples instances, not points); kept as written
        scribbles = reorder_scribbles([rescale_scribbles(s, W, H) for s in scr])
        scribbles = [list(np.asarray(s, dtype=np.float32).reshape(-1)) for s in scribbles]
    if masks is None:
        polygons = [[0.0] * (2 * N_POLYGON_POINTS) for _ in annos]  # sample_sparse_points_from_mask(None) -> zeros (:267-268)
        segs = None
    else:
        raise NotImplementedError("mask-conditioned requests need the reference's polygon sampler "
                                  "(dataset/decode_item.py:217-260, skimage); pass polygons / segs explicitly")
    meta = dict(ckpt=ckpt, prompt=data["caption"], phrases=phrases, polygons=polygons, scribbles=scribbles, segs=segs,
                locations=locations, points=points, alpha_type=[alpha, 0.0, 1 - alpha], save_folder_name=save_folder_name)
    if mis > 0:
        meta["instance_meta"] = [prepare_instance_meta(meta, i, save_folder_name=save_folder_name) for i in range(len(annos))]
    return [meta]

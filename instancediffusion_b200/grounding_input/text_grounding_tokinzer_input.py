"""GroundingNetInput -- drop-in for grounding_input/text_grounding_tokinzer_input.py:13-94.

Pure dict plumbing on the host side of the boundary: `prepare` selects the tensors UniFusion
consumes and remembers their shapes; `get_null_input` returns the all-zero conditioning of the
classifier-free-guidance branch.  The null tensors are cached per (batch, device, dtype): the
reference re-allocates a 30 MiB-per-sample zero `segs` on every uncond forward (:73), here the
same zero tensors are handed back, which also lets UNetModel cache the null object tokens.
"""
import torch as th

_KEYS = ("boxes", "masks", "positive_embeddings", "scribbles", "polygons", "segs", "points")


class GroundingNetInput:
    def __init__(self):
        self.set = False
        self.return_att_masks = False
        self.image_size = 64
        self.return_att_masks32 = False
        self._null_cache = {}

    def prepare(self, batch, image_size=64, device=None, dtype=None, return_att_masks=False):
        self.set = True
        self.return_att_masks = return_att_masks
        pos = batch["text_embeddings"]
        self.dim_scribbles = batch["scribbles"].shape[-1]
        self.dim_polygons = batch["polygons"].shape[-1]
        self.dim_segs = batch["segs"].shape[-1]
        self.batch, self.max_box, self.in_dim = pos.shape
        self.device = pos.device
        self.dtype = pos.dtype
        self._null_cache.clear()
        out = {
            "boxes": batch["boxes"],
            "masks": batch["masks"],
            "positive_embeddings": pos,
            "scribbles": batch["scribbles"],
            "polygons": batch["polygons"],
            "segs": batch["segs"],
            "points": batch["points"],
        }
        if return_att_masks:
            assert "att_masks" in batch
            out["att_masks"] = batch["att_masks"]
        return out

    def get_null_input(self, batch=None, device=None, dtype=None):
        assert self.set, "not set yet, cannot call this funcion"
        batch = self.batch if batch is None else batch
        device = self.device if device is None else device
        dtype = self.dtype if dtype is None else dtype
        key = (batch, str(device), dtype, self.return_att_masks)
        hit = self._null_cache.get(key)
        if hit is not None:
            return hit
        z = lambda *shape: th.zeros(*shape, dtype=dtype, device=device)
        out = {
            "boxes": z(batch, self.max_box, 4),
            "masks": z(batch, self.max_box),
            "positive_embeddings": z(batch, self.max_box, self.in_dim),
            "scribbles": z(batch, self.max_box, self.dim_scribbles),
            "polygons": z(batch, self.max_box, self.dim_polygons),
            # all-zero segs select the learned null seg feature (text_grounding_net.py:279-283); a
            # (batch, max_box, 1, 1) zero view expanded to the full size keeps the semantics without
            # materialising 30 MiB per sample
            "segs": z(batch, self.max_box, 1, 1).expand(batch, self.max_box, self.dim_segs, self.dim_segs),
            "points": z(batch, self.max_box, 2),
        }
        if self.return_att_masks:
            out["att_masks"] = z(batch, self.max_box, self.image_size, self.image_size)
        self._null_cache[key] = out
        return out

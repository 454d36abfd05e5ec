"""Checkpoint pre-pack.  The reference ships one .pth of fp32 tensors per model
(utils/checkpoint.py:224-249 loads saved_ckpt["model"] / ["autoencoder"] / ... strict).  The kernels consume fp16
matrices, so a deployment packs once:

  pack_state_dict(sd)  -> {"index": [(key, shape, dtype, offset, numel), ...], "f16": one flat fp16 tensor holding every
                          matrix (dim >= 2), "f32": one flat fp32 tensor holding every vector / scalar}
  unpack_into(module, pack) restores fp32 masters (matrices upcast from fp16 -- the values the tensor cores see) and is
                          what a rank receives instead of 4.9 GB of fp32: the flat tensors are exactly the buffers
                          parallel.broadcast_pack ships (one NCCL broadcast each).

The pack is a plain dict of tensors: torch.save / torch.load move it; nothing here touches the GPU kernels.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist


def pack_state_dict(sd: Dict[str, torch.Tensor], dtype: torch.dtype = torch.float16) -> dict:
    """`dtype`: the 16-bit type the matrices are stored in (pass torch.bfloat16 for a bf16-storage deployment,
    ops.set_storage_dtype; the slot keeps the name "f16")."""
    index, mats, vecs = [], [], []
    off16 = off32 = 0
    for k, v in sd.items():
        v = v.detach()
        if v.is_floating_point() and v.dim() >= 2:
            index.append((k, tuple(v.shape), "f16", off16, v.numel()))
            mats.append(v.reshape(-1).to(dtype))
            off16 += v.numel()
        else:
            index.append((k, tuple(v.shape), "f32", off32, v.numel()))
            vecs.append(v.reshape(-1).to(torch.float32))
            off32 += v.numel()
    dev = next(iter(sd.values())).device
    return {"index": index,
            "f16": torch.cat(mats) if mats else torch.zeros(0, dtype=dtype, device=dev),
            "f32": torch.cat(vecs) if vecs else torch.zeros(0, dtype=torch.float32, device=dev)}


def unpack_state_dict(pack: dict) -> Dict[str, torch.Tensor]:
    out = {}
    for k, shape, kind, off, n in pack["index"]:
        flat = pack["f16"] if kind == "f16" else pack["f32"]
        out[k] = flat[off:off + n].view(shape).float()
    return out


@torch.no_grad()
def unpack_into(module: torch.nn.Module, pack: dict, strict: bool = True):
    """load_state_dict from a pack (fp32 masters; matrices carry fp16-rounded values)."""
    return module.load_state_dict(unpack_state_dict(pack), strict=strict)


def pack_bytes(pack: dict) -> int:
    return pack["f16"].numel() * 2 + pack["f32"].numel() * 4


@torch.no_grad()
def broadcast_pack(pack: dict, src: int = 0) -> int:
    """Broadcast the two flat buffers of a pack (every rank holds the same index; receivers pass a pack of empty
    buffers of the right sizes, e.g. from `empty_like_pack`).  Returns the bytes sent."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    dist.broadcast(pack["f16"], src=src)
    dist.broadcast(pack["f32"], src=src)
    return pack_bytes(pack)


def empty_like_pack(module: torch.nn.Module, device, dtype: torch.dtype = torch.float16) -> dict:
    """A pack with the index of `module`'s state_dict and uninitialised buffers (the receive side of broadcast_pack)."""
    index, off16, off32 = [], 0, 0
    for k, v in module.state_dict().items():
        if v.is_floating_point() and v.dim() >= 2:
            index.append((k, tuple(v.shape), "f16", off16, v.numel()))
            off16 += v.numel()
        else:
            index.append((k, tuple(v.shape), "f32", off32, v.numel()))
            off32 += v.numel()
    return {"index": index, "f16": torch.empty(off16, dtype=dtype, device=device),
            "f32": torch.empty(off32, dtype=torch.float32, device=device)}

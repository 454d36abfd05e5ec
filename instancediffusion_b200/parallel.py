"""Multi-GPU plumbing for the sampling path: one process per GPU, prompts sharded statically
(rank r takes images r::world, as eval_local.py:172-176 of the reference shards COCO ids), the
frozen weights broadcast ONCE at init over NCCL (NVLink 5 / NVSwitch), and no per-step collective
-- every denoising trajectory is independent.
"""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist


def env_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_distributed(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Static round-robin shard: item i goes to rank i % world."""
    return list(range(rank, n_items, world))


@torch.no_grad()
def broadcast_module_(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 1 << 30,
                      wire_dtype: torch.dtype | None = torch.float16) -> int:
    """Broadcast every parameter and buffer of `module` from rank `src`, coalesced into flat
    buckets (the collective cost on NVSwitch is latency- not link-bound, so a few ~1 GiB buckets
    amortise launch latency).  Returns the number of bytes sent.  After this call no collective is
    issued on the sampling path.

    wire_dtype=float16 (default): matrices (dim >= 2; 99.9 % of the bytes) travel in the precision the
    tensor cores consume them in -- the 2.46 GB pack of SURVEY.md section 8e instead of 4.9 GB of fp32
    masters; vectors (biases, norm gains, gates: used in fp32 by the epilogues) travel as they are.  The
    source rank rounds its own masters the same way, so every rank holds bit-identical parameters and
    samples bit-identical images for the same prompt.  wire_dtype=None ships the masters unchanged."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    total = 0
    groups = {}
    for t in tensors:
        wire = wire_dtype if (wire_dtype is not None and t.dtype == torch.float32 and t.dim() >= 2) else t.dtype
        groups.setdefault(wire, []).append(t)
    for wire, ts in groups.items():
        bucket: List[torch.Tensor] = []
        size = 0

        def flush():
            nonlocal bucket, size, total
            if not bucket:
                return
            flat = torch.cat([t.reshape(-1).to(wire) for t in bucket])
            dist.broadcast(flat, src=src)
            off = 0
            for t in bucket:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))  # (upcast into the master; the source rounds its own too)
                off += n
            total += flat.numel() * flat.element_size()
            bucket, size = [], 0

        esize = torch.empty((), dtype=wire).element_size()
        for t in ts:
            nbytes = t.numel() * esize
            if size + nbytes > bucket_bytes:
                flush()
            bucket.append(t)
            size += nbytes
        flush()
    if hasattr(module, "invalidate_pack"):
        module.invalidate_pack()
    return total


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()

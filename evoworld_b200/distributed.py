"""Multi-GPU sharding of independent clips: one process per GPU, no data-path collective
except one all-gather of the generated latents at each clip boundary.

Mirrors the reference's process-per-GPU launcher (inference_unity_curve_multi_gpu.sh:41-68: disjoint
`--start_idx` ranges, one process per CUDA_VISIBLE_DEVICES) with torch.distributed plumbing.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the process group when world > 1."""
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device(f"cuda:{local}")
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, local, world


def shard_range(num_items: int, rank: int, world: int, start_idx: int = 0) -> range:
    """Contiguous shard of [start_idx, start_idx + num_items) for `rank` — the reference's
    START_IDX + GPU * NUM_DATA_PER_GPU scheme, with the remainder spread over the first ranks."""
    base, rem = divmod(num_items, world)
    lo = start_idx + rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def gather_latents(latents: torch.Tensor) -> torch.Tensor:
    """All-gather of one rank's clip latents [T,4,h,w] (or [1,T,4,h,w]) -> [world, ...] on every rank."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return latents.unsqueeze(0)
    x = latents.contiguous()
    world = dist.get_world_size()
    out = torch.empty(world * x.numel(), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.view(-1))
    return out.view((world,) + tuple(x.shape))

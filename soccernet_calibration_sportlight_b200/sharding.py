"""Frame sharding across the GPUs of one box.  Every frame is independent
through all stages (the reference's driver already treats them so, make_submit.py:59-73), so
rank r takes a contiguous slice of the batch with a full weight replica and the ONLY exchange
is one all-gather of the fixed-size per-frame camera records (128 B each) at the end - NCCL on
the GPU box, gloo in the CPU test-suite."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced slices: the first ``n_frames % world`` ranks get one extra frame."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_records(local: torch.Tensor, n_frames: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """local: this rank's (n_local, ...) results for its ``shard_bounds`` slice -> the
    (n_frames, ...) results of the whole batch on every rank, in frame order.  One collective;
    ragged shards are padded to the largest shard."""
    if not dist.is_available() or not dist.is_initialized():
        if local.shape[0] != n_frames:
            raise ValueError("no process group: the local shard must be the whole batch")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_frames, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: shard has {local.shape[0]} frames, expected {hi - lo}")
    cap = -(-n_frames // world)
    tail = tuple(local.shape[1:])
    send = local.contiguous()
    if send.shape[0] != cap:
        pad = torch.zeros((cap - send.shape[0],) + tail, dtype=local.dtype, device=local.device)
        send = torch.cat([send, pad], dim=0)
    recv = torch.empty((world * cap,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if n_frames == world * cap:
        return recv
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_frames, r, world)
        parts.append(recv[r * cap:r * cap + (b - a)])
    return torch.cat(parts, dim=0)

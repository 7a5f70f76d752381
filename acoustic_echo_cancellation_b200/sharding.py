"""Utterance sharding across GPUs (one process per GPU, ``torch.distributed``).

The stage-1 path shards trivially: utterances are independent (the reference processes them one
at a time, Stage2_lhm/generate_h5files/train_wav2h5.py:13), so every rank runs the fused kernel on a
contiguous slice of the utterance index range and NOTHING on the data path communicates.  The only
collective is the final all-gather of per-utterance metrics (ERLE, a float each) -- NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n_items`` for ``rank``: sizes differ by at most one and the
    first ``n_items % world`` ranks take the extra item."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_metrics(local: torch.Tensor, n_items: int, group=None, async_op: bool = False):
    """All-gather per-utterance metrics of the contiguous shards back into utterance order.
    ``local`` is this rank's 1-D slice (length ``shard_range(n_items, rank, world)``).

    ``async_op=True`` returns ``(tensor, work)``: the collective runs on the backend's own stream and the
    caller's stream is not made to wait until ``work.wait()`` -- the next batch's kernel can start while the
    metrics of this one are still in flight (``tensor`` is valid after ``work.wait()``; ``work`` is None when
    there is nothing to wait for)."""
    if not dist.is_available() or not dist.is_initialized():
        if local.numel() != n_items:
            raise ValueError("single process must hold every item")
        return (local, None) if async_op else local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_items, world)
    if local.numel() != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.numel()} items, expected {sizes[rank]}")
    width = max(sizes) if sizes else 0
    if min(sizes) == width and dist.get_backend(group) == "nccl":
        # equal shards (the usual case): one collective straight into the result, no staging copies
        out = torch.empty(world * width, dtype=local.dtype, device=local.device)
        work = dist.all_gather_into_tensor(out, local.contiguous(), group=group, async_op=async_op)
        return (out, work) if async_op else out
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    work = dist.all_gather(parts, padded, group=group, async_op=async_op)
    if async_op:
        work.wait()      # ragged shards are re-assembled on the caller's stream
    out = torch.cat([p[:s] for p, s in zip(parts, sizes)])
    return (out, None) if async_op else out


def merge_filelists(local_paths: List[str], group=None) -> List[str]:
    """Every rank writes its own h5 shard; the file list (``tr_list.txt``,
    train_wav2h5.py:48-51) is the concatenation in rank order."""
    if not dist.is_available() or not dist.is_initialized():
        return list(local_paths)
    world = dist.get_world_size(group)
    out: List[List[str]] = [None] * world  # type: ignore[list-item]
    dist.all_gather_object(out, list(local_paths), group=group)
    return [p for part in out for p in part]

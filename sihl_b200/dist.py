"""Multi-GPU plumbing of the path: images are independent, so the batch is sharded across
ranks (one process per GPU) and the only exchange is the 8 fp64 loss partial sums
(``include/sihl_od.h``: sums layout) — one all-reduce per step over NCCL/NVLink
(SURVEY.md §8e).  Everything here also runs on the ``gloo`` backend for the CPU tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

NUM_SUMS = 8
TOTAL_WEIGHTS = (1.0, 10.0, 1.0, 1.0)


def shard_range(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous image range ``[start, end)`` of ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(int(global_batch), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_reduce_sums(sums: torch.Tensor, group=None, async_op: bool = False):
    """Global-batch normalisers: sum the 8 partial sums over the ranks (in place)."""
    assert sums.numel() == NUM_SUMS and sums.dtype == torch.float64, (sums.shape, sums.dtype)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def losses_from_sums(sums: torch.Tensor) -> torch.Tensor:
    """ref object_detection.py:163-172, :180, :197, :208, :210 on the 8 sums ->
    ``[location, box, class, iou, total]`` (fp64).  Eight scalars of host-side glue used by the
    distributed tests and for logging; on the GPU path ``sihl_od_loss_finalize`` does this."""
    s = sums.to(torch.float64)
    loc = s[0] / s[1]
    if float(s[6]) == 0.0:
        z = torch.zeros((), dtype=torch.float64, device=s.device)
        return torch.stack([loc, z, z, z, loc])
    iou, box, cls = s[2] / s[3], s[4] / s[3], s[5] / s[3]
    return torch.stack([loc, box, cls, iou, loc + 10.0 * box + cls + iou])


def ddp_mean_of_local_losses(local_losses: torch.Tensor, group=None) -> torch.Tensor:
    """The "DDP-faithful" semantics: every rank normalises by its own shard (what the reference
    does under Lightning DDP) and the logged value is the mean over ranks."""
    out = local_losses.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        out /= dist.get_world_size(group)
    return out

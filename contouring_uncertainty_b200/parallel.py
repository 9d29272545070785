"""Multi-GPU plumbing of the path: image sharding and the single end-of-step exchange.

The path shards by image: contiguous blocks of PATIENTS per rank, so the A2C/A4C views
and ED/ES instants of a patient stay on one GPU.  No collective sits on the per-sample critical path;
the only exchange is an all_gather of the per-image metric distributions (T_a values per image) and
an all_reduce of dataset-level accumulators, NCCL over NVLink on GPUs (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_images(n_images: int, rank: int, world: int, images_per_patient: int = 1) -> Tuple[int, int]:
    """[start, stop) of the contiguous image block of ``rank``; blocks are multiples of a patient.

    Patients are dealt as evenly as possible (the first ``n_patients % world`` ranks get one more).
    """
    if n_images % images_per_patient:
        raise ValueError("n_images must be a multiple of images_per_patient")
    n_pat = n_images // images_per_patient
    base, extra = divmod(n_pat, world)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return start * images_per_patient, stop * images_per_patient


def shard_sizes(n_images: int, world: int, images_per_patient: int = 1) -> List[int]:
    return [b - a for a, b in (shard_images(n_images, r, world, images_per_patient) for r in range(world))]


def gather_metric_distributions(local: torch.Tensor, sizes: Sequence[int], group=None) -> torch.Tensor:
    """all_gather of per-image distributions (n_local, T, ...) with ragged n_local -> (n_total, T, ...).

    ``sizes[r]`` = number of images of rank r (from ``shard_sizes``).  Ragged shards are padded to
    the largest block for the collective and trimmed afterwards.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def reduce_dataset_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """all_reduce(SUM) of dataset-level accumulators (counts, sums, sums of squares ...), in place."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats

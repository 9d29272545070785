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


class StepExchange:
    """The end-of-step exchange, overlapped with the next step's kernels.

    Step i's per-image distributions (equal shard sizes) and dataset-level sums are copied into slot
    ``i % depth`` and exchanged with ONE ``all_gather_into_tensor`` + one ``all_reduce`` issued with
    ``async_op=True``: NCCL's stream waits for the producing kernels, the compute stream does not wait for
    NCCL, so step i + 1 runs while step i's exchange is in flight.  A slot is reused only after its previous
    exchange has completed (``wait`` before the copy).  ``finish()`` joins everything onto the current stream.
    """

    def __init__(self, n_local: int, tail_shape: Sequence[int], dtype, n_stats: int, device, depth: int = 2, group=None,
                 stats_dtype=torch.float64):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.slots = []
        for _ in range(depth):
            self.slots.append({
                "local": torch.empty((n_local,) + tuple(tail_shape), dtype=dtype, device=device),
                "gathered": torch.empty((self.world * n_local,) + tuple(tail_shape), dtype=dtype, device=device),
                "stats": torch.empty((n_stats,), dtype=stats_dtype, device=device),
                "works": [],
            })
        self.last = None

    def submit(self, step: int, local: torch.Tensor, stats: torch.Tensor):
        slot = self.slots[step % len(self.slots)]
        for w in slot["works"]:
            w.wait()
        slot["local"].copy_(local, non_blocking=True)
        slot["stats"].copy_(stats, non_blocking=True)
        if self.world > 1:
            slot["works"] = [dist.all_gather_into_tensor(slot["gathered"], slot["local"], group=self.group, async_op=True),
                             dist.all_reduce(slot["stats"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)]
        else:
            slot["gathered"].copy_(slot["local"], non_blocking=True)
        self.last = slot
        return slot

    def finish(self):
        """Wait for every exchange in flight; returns (gathered, stats) of the last submitted step."""
        for slot in self.slots:
            for w in slot["works"]:
                w.wait()
            slot["works"] = []
        return (self.last["gathered"], self.last["stats"]) if self.last is not None else (None, None)


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to GPU ``device_index`` (so that pinned host
    buffers allocated afterwards are first-touched on the GPU's NUMA node and H2D / D2H copies do not cross sockets).
    Returns False (and changes nothing) when NVML is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False

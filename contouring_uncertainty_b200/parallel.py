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


PACK_PARTS = 16          # CUB200_METRIC_PACK_PARTS (include/cub200.h): partial int64 sums in a slot's tail


class MetricExchange:
    """The path's exchange as ONE collective per block of ``every`` steps (default: one per predict epoch).

    The reference gathers nothing per batch: its clinical metric processors run at the END of the predict epoch over all
    BatchResults (cu/results/clinical/clinical_metrics.py:51-66).  So a rank only buffers what those processors need --
    per step the per-image metric distributions (n_local, T) int32 and the exact int64 dataset sums -- in a device ring of
    ``every`` slots, and one ``all_gather_into_tensor`` ships a whole block: [every][n_local * T + 64] int32 words per rank, the
    int64 sums riding (as 16 partial pairs) in the tail of each slot (no separate all_reduce: every rank adds the gathered sums
    itself, exactly, in int64).  Blocks are double-buffered and the collective is asynchronous, so block b is on the wire
    while block b + 1 fills; ``finish()`` flushes a partial block and joins everything onto the current stream.
    """

    def __init__(self, n_local: int, n_per_image: int, every: int, device, group=None, n_blocks: int = 2):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.every = max(1, int(every))
        self.n_local, self.n_per_image = int(n_local), int(n_per_image)
        self.payload = self.n_local * self.n_per_image + 4 * PACK_PARTS
        # one ring of n_blocks * every slots: a step captured in a CUDA graph addresses it with its device-side step counter
        # (slot = step % (n_blocks * every), engine.GraphedStep); block b = slots [b * every, (b + 1) * every)
        self.ring = torch.zeros((n_blocks * self.every, self.payload), dtype=torch.int32, device=device)
        self.blocks = [{"local": self.ring[b * self.every:(b + 1) * self.every],
                        "gathered": torch.zeros((self.world * self.every, self.payload), dtype=torch.int32, device=device),
                        "work": None, "filled": 0} for b in range(n_blocks)]
        self.cur = 0
        self.collectives = 0
        self.last = None

    def prepare(self):
        """Call BEFORE the step that writes the next slot: a block that is being reused must have left the wire."""
        blk = self.blocks[self.cur]
        if blk["filled"] == 0 and blk["work"] is not None:
            blk["work"].wait()
            blk["work"] = None
        return blk

    def push(self, values: torch.Tensor, sums: torch.Tensor = None):
        """values (n_local, n_per_image) int32 [, sums (2,) int64] of this step (device tensors, no synchronisation).
        On CUDA tensors: ONE launch (cub200_metric_pack copies the values and forms the exact int64 sums itself)."""
        blk = self.prepare()
        slot = blk["local"][blk["filled"]]
        if values.is_cuda:
            from . import _native as nat
            with torch.cuda.device(values.device):
                nat.check(nat.lib().cub200_metric_pack(values.data_ptr(), values.numel(), slot.data_ptr(), 1, None,
                                                       nat.stream_ptr(values.device)), "cub200_metric_pack")
        else:                                               # host logic under gloo (CPU tests)
            if sums is None:
                v64 = values.reshape(-1).to(torch.int64)
                sums = torch.stack([v64.sum(), (v64 * v64).sum()])
            tail = self.payload - 4 * PACK_PARTS
            slot[:tail].copy_(values.reshape(-1), non_blocking=True)
            slot[tail:].zero_()                               # the sums ride as part 0, the other parts are zero
            slot[tail:tail + 4].copy_(sums.view(torch.int32), non_blocking=True)
        self.note_pushed()

    def slot_index(self) -> int:
        """Ring slot the next step writes (what a GraphedStep's counter must be congruent to, mod ring.shape[0])."""
        return self.cur * self.every + self.blocks[self.cur]["filled"]

    def note_pushed(self):
        """The slot of this step has been written (by push(), or by a graph replay that addresses the ring itself)."""
        blk = self.blocks[self.cur]
        blk["filled"] += 1
        if blk["filled"] == self.every:
            self._ship(blk)

    def _ship(self, blk):
        if self.world > 1:
            blk["work"] = dist.all_gather_into_tensor(blk["gathered"], blk["local"], group=self.group, async_op=True)
            self.collectives += 1
        else:
            blk["gathered"].copy_(blk["local"], non_blocking=True)
        blk["shipped"] = blk["filled"]
        blk["filled"] = 0
        self.last = blk
        self.cur = (self.cur + 1) % len(self.blocks)

    def finish(self):
        """Ship a partial block, wait for everything in flight.  Returns (values (world, steps, n_local, n_per_image) int32,
        sums (steps, 2) int64 summed over the ranks) of the LAST shipped block, or (None, None) if nothing was pushed."""
        blk = self.blocks[self.cur]
        if blk["filled"]:
            self._ship(blk)
        for b in self.blocks:
            if b["work"] is not None:
                b["work"].wait()
                b["work"] = None
        if self.last is None:
            return None, None
        k = self.last["shipped"]
        g = self.last["gathered"].view(self.world, self.every, self.payload)[:, :k]
        tail = self.payload - 4 * PACK_PARTS
        values = g[..., :tail].reshape(self.world, k, self.n_local, self.n_per_image)
        sums = g[..., tail:].contiguous().view(torch.int64).reshape(self.world, k, PACK_PARTS, 2).sum((0, 2))   # over ranks and parts
        return values, sums


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

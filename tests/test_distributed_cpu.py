"""world_size-2 gloo tests (CPU) of the N>1 host logic: sharding, ragged gather, dataset reduction."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from contouring_uncertainty_b200 import parallel as P


def test_shard_images_keeps_patients_together_and_covers_everything():
    for n_pat, world, ipp in ((10, 4, 4), (7, 2, 2), (64, 8, 1), (3, 4, 4)):
        n = n_pat * ipp
        blocks = [P.shard_images(n, r, world, ipp) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for (a, b), (c, d) in zip(blocks, blocks[1:]):
            assert b == c
        for a, b in blocks:
            assert a % ipp == 0 and b % ipp == 0
        sizes = P.shard_sizes(n, world, ipp)
        assert sum(sizes) == n and max(sizes) - min(sizes) <= ipp


def _worker(rank, world, port, n_images, T):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_images * T, dtype=torch.int32).reshape(n_images, T)     # "areas" of every image
        a, b = P.shard_images(n_images, rank, world, images_per_patient=2)
        sizes = P.shard_sizes(n_images, world, images_per_patient=2)
        got = P.gather_metric_distributions(full[a:b].clone(), sizes)
        assert torch.equal(got, full), (rank, got.shape)
        stats = torch.tensor([float(full[a:b].sum()), float(b - a)], dtype=torch.float64)
        P.reduce_dataset_stats(stats)
        assert stats[0].item() == float(full.sum()) and stats[1].item() == n_images
    finally:
        dist.destroy_process_group()


def test_gather_and_reduce_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    # 5 patients x 2 images over 2 ranks -> ragged shards (6 and 4 images)
    mp.spawn(_worker, args=(2, port, 10, 7), nprocs=2, join=True)


def _exchange_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_local, T = 3, 5
        ex = P.StepExchange(n_local, (T,), torch.int32, 2, "cpu", depth=2)
        for step in range(5):                                   # slots are reused: 5 steps over depth 2
            local = torch.full((n_local, T), 100 * step + rank, dtype=torch.int32)
            ex.submit(step, local, torch.tensor([float(step + rank), 1.0], dtype=torch.float64))
        gathered, stats = ex.finish()
        want = torch.cat([torch.full((n_local, T), 400 + r, dtype=torch.int32) for r in range(world)])
        assert torch.equal(gathered, want)
        assert stats.tolist() == [float(sum(4 + r for r in range(world))), float(world)]
    finally:
        dist.destroy_process_group()


def test_overlapped_step_exchange_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_exchange_worker, args=(2, port), nprocs=2, join=True)


def _metric_exchange_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_local, T, every = 3, 5, 4
        ex = P.MetricExchange(n_local, T, every, "cpu")
        big = 2 ** 40                                            # the sums are int64 riding in int32 words: must survive intact
        for step in range(10):                                   # 2 full blocks (one collective each) + a partial one of 2 steps
            vals = torch.full((n_local, T), 100 * step + rank, dtype=torch.int32)
            ex.push(vals, torch.tensor([big + step + rank, -7 * step], dtype=torch.int64))
        values, sums = ex.finish()
        assert ex.collectives == 3
        assert values.shape == (world, 2, n_local, T) and sums.shape == (2, 2)
        for r in range(world):
            assert torch.equal(values[r, 0], torch.full((n_local, T), 800 + r, dtype=torch.int32))
            assert torch.equal(values[r, 1], torch.full((n_local, T), 900 + r, dtype=torch.int32))
        assert sums.tolist() == [[sum(big + 8 + r for r in range(world)), -7 * 8 * world], [sum(big + 9 + r for r in range(world)), -7 * 9 * world]]
        # every = number of steps: ONE collective for the whole epoch
        ex = P.MetricExchange(n_local, T, 6, "cpu")
        for step in range(6):
            ex.push(torch.full((n_local, T), step * 10 + rank, dtype=torch.int32), torch.tensor([step, rank], dtype=torch.int64))
        values, sums = ex.finish()
        assert ex.collectives == 1 and values.shape == (world, 6, n_local, T)
        assert sums[:, 0].tolist() == [s * world for s in range(6)] and sums[:, 1].tolist() == [sum(range(world))] * 6
    finally:
        dist.destroy_process_group()


def test_metric_exchange_one_collective_per_block_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_metric_exchange_worker, args=(2, port), nprocs=2, join=True)


def test_metric_exchange_single_process():
    ex = P.MetricExchange(2, 3, 5, "cpu")
    assert ex.finish() == (None, None)
    for step in range(7):
        ex.push(torch.full((2, 3), step, dtype=torch.int32), torch.tensor([step, 2 * step], dtype=torch.int64))
    values, sums = ex.finish()
    assert values.shape == (1, 2, 2, 3) and values[0, 1, 0, 0].item() == 6 and sums.tolist() == [[5, 10], [6, 12]]

"""Reduced-size parity soak inside the GPU suite (the full-size one is tools/soak_parity.py): random, self-intersecting,
off-frame and half-integer contours through the CUDA rasteriser (fast span path AND the general flood it defers to) against
the oracle, bit for bit; the Simpson diameters of the same masks, fp64-identical.  The oracle runs on the host cores."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_rasteriser_and_simpson_soak_vs_oracle():
    import soak_parity as S
    from contouring_uncertainty_b200 import engine as E

    n, n_simpson = 2400, 160
    rng = np.random.default_rng(77)
    b = synth.LV_EXAMPLE_SHAPE
    kind = rng.integers(0, 5, n)
    c = b[None] + rng.normal(0, 3.0, (n, 21, 2)) + rng.normal(0, 6.0, (n, 1, 2))
    c[kind == 1] = b[None] + rng.normal(0, 11.0, ((kind == 1).sum(), 21, 2))                       # self-intersecting: general flood
    c[kind == 2] += rng.choice([-1, 1], ((kind == 2).sum(), 1, 2)) * rng.uniform(60, 130, ((kind == 2).sum(), 1, 2))  # off-frame
    c[kind == 3] = np.round(c[kind == 3]) + 0.5                                                     # rounding ties
    c = c.astype(np.float32)
    dev = torch.device("cuda:0")
    masks, area = E.rasterize_contours(torch.from_numpy(c).to(dev), return_masks=True)
    diff = torch.zeros((n // 100, 256, 256), dtype=torch.int32, device=dev)
    _, area2 = E.rasterize_contours(torch.from_numpy(c).to(dev), return_masks=False, diff=diff, per_image=100)
    got = masks.cpu().numpy().view(np.uint8).reshape(n, 256, 32)
    workers = min(os.cpu_count() or 1, 32)
    with mp.get_context("spawn").Pool(workers) as pool:
        want = np.concatenate(pool.map(S._raster_chunk, np.array_split(c, workers * 2)))
        sel = np.flatnonzero((kind == 0) | (kind == 1))[:n_simpson]
        m_np = np.unpackbits(want[sel], axis=-1, bitorder="little")
        wd = {vs: np.concatenate(pool.map(S._simpson_chunk, [(x, vs) for x in np.array_split(m_np, workers)])) for vs in ((0.5, 0.5), (0.5, 0.37))}
    assert int((got != want).any((1, 2)).sum()) == 0
    # the aggregated path (difference map + areas) agrees with the masks
    assert torch.equal(area, area2) and np.array_equal(area.cpu().numpy(), np.unpackbits(want, axis=-1, bitorder="little").sum((1, 2)))
    count = E.finalize_maps(diff, 100)[0].cpu().numpy()
    assert np.array_equal(count, np.unpackbits(want, axis=-1, bitorder="little").reshape(n // 100, 100, 256, 256).sum(1))
    for vs, w in wd.items():
        d, s = E.simpson_diameters(masks[torch.from_numpy(sel).to(dev)].contiguous(), vs)
        gd = np.concatenate([d.cpu().numpy(), s.cpu().numpy()[:, None]], 1)
        assert int((~((gd == w) | (np.isnan(gd) & np.isnan(w)))).any(1).sum()) == 0, vs

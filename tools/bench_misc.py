"""Timings of the small 'next row' kernels that have no bench of their own: general sample_entropy, LV + MYO masks,
generic-size rasteriser.   python tools/bench_misc.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from contouring_uncertainty_b200 import engine as E  # noqa: E402


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


dev = torch.device("cuda:0")
p = torch.softmax(torch.randn((16, 64, 3, 256, 256), device=dev), dim=2).contiguous()
t = timed(lambda: E.sample_entropy(p))
print(f"sample_entropy 16 img x 64 samples x 3 classes: {t:.3f} ms = {p.numel() * 4 / t / 1e6:.0f} GB/s read once")
lm = torch.from_numpy(np.tile(synth.endo_epi_landmarks(50, seed=1), (200, 1, 1)).astype(np.float32)).to(dev)
t = timed(lambda: E.us_myo_masks(lm))
print(f"us_myo_masks {len(lm)} landmark sets: {t:.3f} ms = {len(lm) / t * 1e3:.3e} sets/s")
c = torch.from_numpy(np.tile(synth.closed_contours(50, 44, seed=2), (400, 1, 1))).to(dev)
t = timed(lambda: E.rasterize_contours_generic(c))
print(f"rasterize_contours_generic {len(c)} contours of 44 points: {t:.3f} ms = {len(c) / t * 1e3:.3e} contours/s")

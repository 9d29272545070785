"""Time the mask-based LV longitudinal length (EchoMeasure.gls pieces) on one GPU.

    python tools/bench_gls.py [n_masks]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from contouring_uncertainty_b200 import engine as E  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    c = synth.LV_EXAMPLE_SHAPE[None] + rng.normal(0, 2.0, (n, 21, 2)) + rng.normal(0, 8, (n, 1, 2))
    masks, _ = E.rasterize_contours(torch.from_numpy(c.astype(np.float32)).to(dev))
    E.mask_lv_length(masks)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        length = E.mask_lv_length(masks)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ok = ~torch.isnan(length)
    print(f"{n} masks: {best:.3f} ms = {n / best * 1e3:.3e} masks/s; measurable {ok.float().mean().item():.3f}, "
          f"mean length {length[ok].mean().item():.2f} px")


if __name__ == "__main__":
    main()

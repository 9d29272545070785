"""Launch the kernels of the 'next' rows (segmentation post-processing, mask GLS, generic rasteriser) twice each at
representative sizes, for an ncu capture:

    ncu --set full --clock-control none -k regex:"seg_blob_kernel|seg_entropy_kernel|lv_length_kernel|raster_generic_kernel" \
        -o gpurun_out/prof_next python tools/profile_next_rows.py
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

dev = torch.device("cuda:0")
base = torch.from_numpy(synth.segmentation_samples(4, 128, seed=9, adversarial=False, speckle=0.0)).to(dev)
probs = base.repeat(16, 1, 1, 1).contiguous()                       # 64 images x 128 samples
rng = np.random.default_rng(0)
c = synth.LV_EXAMPLE_SHAPE[None] + rng.normal(0, 2.0, (20000, 21, 2)) + rng.normal(0, 8, (20000, 1, 2))
lv_masks, _ = E.rasterize_contours(torch.from_numpy(c.astype(np.float32)).to(dev))
lung = torch.from_numpy(synth.closed_contours(20000, 50, seed=3)).to(dev)
for _ in range(2):
    masks = E.seg_largest_blob(probs)
    E.seg_entropy(probs, masks, return_pred_samples=True)
    E.mask_lv_length(lv_masks)
    E.rasterize_contours_generic(lung, round_landmarks=True)
torch.cuda.synchronize()
print("done")

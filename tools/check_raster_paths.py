#!/usr/bin/env python
"""Self-consistency of the rasteriser's two outputs (GPU only, no oracle): the difference map / areas the fast and slow
fill paths accumulate must equal the sum of the bit-packed masks the same contours give.  Prints one JSON line."""
import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, synth
from contouring_uncertainty_b200 import engine as E

dev = torch.device("cuda:0")
rng = np.random.default_rng(7)
n, per = 8000, 500
b = synth.LV_EXAMPLE_SHAPE
kind = rng.integers(0, 5, n)
c = b[None] + rng.normal(0, 3.0, (n, 21, 2)) + rng.normal(0, 6.0, (n, 1, 2))
c[kind == 1] = b[None] + rng.normal(0, 11.0, ((kind == 1).sum(), 21, 2))
c[kind == 2] += rng.choice([-1, 1], ((kind == 2).sum(), 1, 2)) * rng.uniform(60, 130, ((kind == 2).sum(), 1, 2))
c[kind == 3] = np.round(c[kind == 3]) + 0.5
ct = torch.from_numpy(c.astype(np.float32)).to(dev)
masks, area = E.rasterize_contours(ct, return_masks=True)
diff = torch.zeros((n // per, 256, 256), dtype=torch.int32, device=dev)
_, area2 = E.rasterize_contours(ct, return_masks=False, diff=diff, per_image=per)
count, _, _, _ = E.finalize_maps(diff, per)
want = E.unpack_masks(masks).reshape(n // per, per, 256, 256).sum(1, dtype=torch.int32)
print(json.dumps({"check": "diff-map / area path == sum of masks", "contours": n,
                  "count_maps_equal": bool(torch.equal(count, want)), "areas_equal": bool(torch.equal(area, area2)),
                  "area_equals_popcount": bool(torch.equal(area.long(), E.unpack_masks(masks).sum((1, 2)).long()))}))

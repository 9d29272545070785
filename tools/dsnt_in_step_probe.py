"""Why the DSNT pass takes 68 us inside the step loop and 59 us alone: time it (CUDA events around the one launch) after
(a) another DSNT pass, (b) a whole step that leaves the full fp64 maps (count + mean + entropy + pred: 88 MB at 64 images)
dirty in L2, (c) a whole step with the compact outputs (uint16 count + pred: 12.6 MB).   python tools/dsnt_in_step_probe.py"""
import os, sys, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import bench
from contouring_uncertainty_b200 import UncertaintyEngine

dev = torch.device("cuda:0")
eng = UncertaintyEngine(device=dev)
n, T, iters = 64, 1000, 60
pool = [bench.make_logits_on_device(torch, dev, n, 1000 + i) for i in range(4)]
ws_full = eng.workspace(n, T)
ws_compact = eng.workspace(n, T, want_count=False, want_maps=False, want_count16=True)
out = {}
for name, ws in (("after_dsnt", None), ("after_step_full_maps", ws_full), ("after_step_compact", ws_compact)):
    w = ws or ws_full
    ts = []
    for i in range(iters + 5):
        if ws is None:
            eng.dsnt_into(w, pool[(i + 1) % 4], w.mu, w.cov)
        else:
            eng.run_into(w, pool[(i + 1) % 4], seed=i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.dsnt_into(w, pool[i % 4], w.mu, w.cov); e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
    out[name + "_us"] = round(float(np.median(ts)), 2)
bytes_dsnt = n * 21 * (256 * 256 * 4 + 44)
out["dsnt_GBps"] = {k: round(bytes_dsnt / (v * 1e-6) / 1e9, 1) for k, v in out.items()}
print(json.dumps(out))

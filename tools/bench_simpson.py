"""Micro-benchmark of the uncertainty-propagation path (config 4 shape): python tools/bench_simpson.py [patients] [T]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, synth
from contouring_uncertainty_b200 import engine as E
dev = torch.device("cuda:0")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
mu, cov = synth.point_distributions(P * 4, seed=123)
mu = mu.reshape(P, 4, 21, 2)
ctr = mu.mean(2, keepdims=True)
for v, s in enumerate((1.0, 0.8, 1.05, 0.84)):
    mu[:, v] = ctr[:, 0] + (mu[:, 0] - ctr[:, 0]) * s
mu = torch.from_numpy(mu.reshape(P * 4, 21, 2).copy()).to(dev); cov = torch.from_numpy(cov).to(dev)
model = E.ShapeModelConstants(device=dev)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), r
a_t = E.psm_factor(model, mu, cov)
t_s, contours = timed(lambda: E.sample_contours(mu, a_t, T, seed=1))
t_r, (masks, area) = timed(lambda: E.rasterize_contours(contours.reshape(-1, 21, 2), return_masks=True))
L = E.nat.lib(); n = P * 4 * T
ws = torch.empty((L.cub200_simpson_workspace_bytes(n),), dtype=torch.uint8, device=dev)
t_d, (diam, step, npts) = timed(lambda: E.simpson_diameters(masks.reshape(P, 4, T, 2048), 0.5, return_npts=True))
t_v, (edv, esv, ef) = timed(lambda: E.simpson_volumes(diam, step))
print(f"patients {P} T {T} ({n} masks): sample {t_s:.2f} ms, raster+masks {t_r:.2f} ms, simpson diameters {t_d:.2f} ms "
      f"({n/t_d*1e3:.3e} masks/s), volumes {t_v:.3f} ms; contour pts mean {npts.float().mean().item():.0f} max {npts.max().item()}; "
      f"EDV {edv.mean().item():.1f}+-{edv.std().item():.1f} ESV {esv.mean().item():.1f} EF {ef.nanmean().item():.3f} nan {ef.isnan().float().mean().item():.4f} invalid {(step!=step).float().mean().item():.5f}")

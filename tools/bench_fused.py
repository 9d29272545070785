"""Micro-benchmark of the fused kernel alone (config-2 size), for variant sweeps: CUB200_LIB=... python tools/bench_fused.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, synth
from contouring_uncertainty_b200 import engine as E
dev = torch.device("cuda:0")
n, T = 64, 1000
mu, cov = synth.point_distributions(n, seed=123)
mu, cov = torch.from_numpy(mu).to(dev), torch.from_numpy(cov).to(dev)
model = E.ShapeModelConstants(device=dev)
a_t = E.psm_factor(model, mu, cov)
ws = E.Workspace(n, T, 34, dev)
L = E.nat.lib(); st = torch.cuda.current_stream().cuda_stream
DIFF = None if os.environ.get('NODIFF') else ws.diff.data_ptr()   # NODIFF=1: no run-boundary atomics (how much of the time are they?)
def run(seed):
    ws.diff.zero_()
    E.nat.check(L.cub200_sample_raster_aggregate(mu.data_ptr(), a_t.data_ptr(), 34, None, seed, 0, n, T, 1, ws.contours.data_ptr(), ws.area.data_ptr(), DIFF, 1, None, ws.raster_ws.data_ptr(), None, st), "fused")
for i in range(5): run(i)
torch.cuda.synchronize()
ts = []
for i in range(30):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ws.diff.zero_(); e0.record()
    E.nat.check(L.cub200_sample_raster_aggregate(mu.data_ptr(), a_t.data_ptr(), 34, None, 100 + i, 0, n, T, 1, ws.contours.data_ptr(), ws.area.data_ptr(), DIFF, 1, None, ws.raster_ws.data_ptr(), None, st), "fused")
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(os.environ.get("CUB200_LIB", "default"), "fused median ms", float(np.median(ts)), "min", min(ts), "area sum", int(ws.area.sum()))

"""Per-shard weight of the config-2 step: the 8 shards bench.py gives the 8 ranks (logit seeds 1002..1009), on ONE GPU:
step time, samples deferred to the general flood, mean mask area.  Explains the spread of per-rank times at N > 1."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench as B
from contouring_uncertainty_b200 import engine as E
dev = torch.device("cuda:0")
n, T = 64, 1000
eng = E.UncertaintyEngine(device=dev)
ws = eng.workspace(n, T)
L = E.nat.lib(); st = torch.cuda.current_stream().cuda_stream
for r in range(8):
    logits = B.make_logits_on_device(torch, dev, n, seed=1002 + r)
    for i in range(3): eng.run_into(ws, logits, seed=i)
    torch.cuda.synchronize()
    ts, q = [], []
    for i in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run_into(ws, logits, seed=10 + i); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); q.append(int(ws.raster_ws.view(torch.int32)[0]))
    tr = []
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ws.diff.zero_(); e0.record()
        E.nat.check(L.cub200_rasterize_contours(ws.contours.data_ptr(), n * T, 1, None, ws.area.data_ptr(), ws.diff.data_ptr(), T, ws.raster_ws.data_ptr(), st), "r")
        e1.record(); torch.cuda.synchronize(); tr.append(e0.elapsed_time(e1))
    cov = ws.cov.reshape(n, 21, 4)
    print(json.dumps({"rank": r, "step_ms": round(float(np.median(ts)), 4), "raster_ms": round(float(np.median(tr)), 4), "deferred_per_step": float(np.mean(q)),
                      "mean_area": float(ws.area.float().mean()), "mean_trace_cov": float((cov[:, :, 0] + cov[:, :, 3]).mean())}))

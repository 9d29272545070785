"""Per-kernel device times of the headline step (config 2) for the library named by CUB200_LIB: median of 30 launches each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench as B
from contouring_uncertainty_b200 import engine as E
dev = torch.device("cuda:0")
n, T = 64, 1000
eng = E.UncertaintyEngine(device=dev)
logits = B.make_logits_on_device(torch, dev, n, seed=1002)
ws = eng.workspace(n, T)
L = E.nat.lib(); st = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=30):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3
eng.run_into(ws, logits, seed=1)
res = {
 "dsnt": timeit(lambda: E.nat.check(L.cub200_dsnt_moments(logits.data_ptr(), n * 21, 256, 256, 1, 1, None, ws.mu.data_ptr(), ws.cov.data_ptr(), ws.dsnt_ws.data_ptr(), st), "d")),
 "factor": timeit(lambda: E.nat.check(L.cub200_psm_factor(eng.model.pack.data_ptr(), ws.mu.data_ptr(), ws.cov.data_ptr(), n, ws.a_t.data_ptr(), st), "f")),
 "sample": timeit(lambda: E.nat.check(L.cub200_sample_raster_aggregate(ws.mu.data_ptr(), ws.a_t.data_ptr(), 34, None, 5, 0, n, T, 1, ws.contours.data_ptr(), None, None, 1, None, None, None, st), "s")),
 "raster(+general)": timeit(lambda: E.nat.check(L.cub200_rasterize_contours(ws.contours.data_ptr(), n * T, 1, None, ws.area.data_ptr(), ws.diff.data_ptr(), T, ws.raster_ws.data_ptr(), None, st), "r")),
 "finalize": timeit(lambda: E.nat.check(L.cub200_finalize_maps(ws.diff.data_ptr(), n, T, ws.count.data_ptr(), ws.mean.data_ptr(), ws.entropy.data_ptr(), ws.pred.data_ptr(), ws.lut.data_ptr(), 1, None, None, st), "z")),
 "step (run_into)": timeit(lambda: eng.run_into(ws, logits, seed=3)),
}
print(os.environ.get("CUB200_LIB", "default"), {k: round(v, 1) for k, v in res.items()}, "us")

"""Does the HBM-bound DSNT kernel of the NEXT batch overlap the issue-bound tail (factor, sampler, rasteriser, finalize) of the
current one when they are queued on two streams?  python tools/overlap_probe.py [iters]   (CUB200_LIB=<variant .so> to compare)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import synth
from contouring_uncertainty_b200 import UncertaintyEngine, _native as nat
from contouring_uncertainty_b200.engine import K, HW

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda:0")
eng = UncertaintyEngine(device=dev)
n, T = 64, 1000
ws = eng.workspace(n, T)
import bench
pool = [bench.make_logits_on_device(torch, dev, n, 1000 + i) for i in range(4)]
mu2 = torch.empty_like(ws.mu); cov2 = torch.empty_like(ws.cov)
L = nat.lib()
eng.run_into(ws, pool[0]); torch.cuda.synchronize()
sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=0)
hi = torch.cuda.Stream(dev, priority=-1)


def dsnt(i, st):
    nat.check(L.cub200_dsnt_moments(pool[i % 4].data_ptr(), n * K, HW, HW, 1, 1, None, mu2.data_ptr(), cov2.data_ptr(), ws.dsnt_ws.data_ptr(), st.cuda_stream), "dsnt")


def tail(i, st):
    s = st.cuda_stream
    nat.check(L.cub200_psm_factor(eng.model.pack.data_ptr(), ws.mu.data_ptr(), ws.cov.data_ptr(), n, ws.a_t.data_ptr(), s), "factor")
    with torch.cuda.stream(st):
        ws.diff.zero_()
    nat.check(L.cub200_sample_raster_aggregate(ws.mu.data_ptr(), ws.a_t.data_ptr(), ws.zdim, None, 7, i * n * T, n, T, 1, ws.contours.data_ptr(),
                                               ws.area.data_ptr(), ws.diff.data_ptr(), 1, None, ws.raster_ws.data_ptr(), None, s), "sra")
    nat.check(L.cub200_finalize_maps(ws.diff.data_ptr(), n, T, nat.ptr(ws.count), nat.ptr(ws.mean), nat.ptr(ws.entropy), ws.pred.data_ptr(),
                                     ws.lut.data_ptr(), 1, nat.ptr(ws.variance), nat.ptr(ws.count16), s), "fin")


def timed(body, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
        body()
        e1.record(torch.cuda.current_stream(dev)); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best


cur = torch.cuda.current_stream(dev)

def serial():
    for i in range(iters):
        dsnt(i, cur); tail(i, cur)

def overlapped(a, b):
    def f():
        a.wait_stream(cur); b.wait_stream(cur)
        for i in range(iters):
            # step i: dsnt of batch i + 1 beside the tail of batch i, joined at the end of the step like the branches of a graph
            dsnt(i, b); tail(i, a)
            a.wait_stream(b); b.wait_stream(a)
        cur.wait_stream(a); cur.wait_stream(b)
    return f

def only(fn):
    def f():
        for i in range(iters): fn(i, cur)
    return f

print({"lib": os.environ.get("CUB200_LIB", "default"), "dsnt_only_ms": round(timed(only(dsnt)), 4), "tail_only_ms": round(timed(only(tail)), 4),
       "serial_ms": round(timed(serial), 4), "two_streams_ms": round(timed(overlapped(sa, sb)), 4),
       "two_streams_tail_high_priority_ms": round(timed(overlapped(hi, sb)), 4)})

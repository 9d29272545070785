"""DSNT kernel alone at config-2 size: CUB200_DSNT_CLUSTER=1|2|4 python tools/bench_dsnt.py [n_img]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import bench as B
from contouring_uncertainty_b200 import engine as E
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
logits = B.make_logits_on_device(torch, dev, n, seed=1)
for _ in range(5): E.dsnt_moments(logits)
torch.cuda.synchronize()
ts = []
for _ in range(50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = E.dsnt_moments(logits); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
t = float(np.median(ts))
print(f"cluster={os.environ.get('CUB200_DSNT_CLUSTER', '4')} n_img={n}: dsnt {t*1e3:.1f} us, {n*21*(65536*4+44)/t/1e6:.0f} GB/s, mu[0,0]={r[1][0,0].tolist()}")

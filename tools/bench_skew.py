"""Micro-benchmark of the skew-normal path (config-3 per-GPU shape, scaled): python tools/bench_skew.py [n_img] [T]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, synth
from contouring_uncertainty_b200 import engine as E
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
mu, cov = synth.point_distributions(n, seed=123)
al = synth.skew_alphas(n, seed=124)
mu, cov, al = (torch.from_numpy(a).to(dev) for a in (mu, cov, al))
model = E.ShapeModelConstants(device=dev)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), r
t_tab, pdfs = timed(lambda: E.skew_pdf_grid(mu, cov, al, flip_alpha_y=True))
t_cond, _ = timed(lambda: E.psm_conditionals(model, mu))
t_smp, (out, idx) = timed(lambda: E.skew_sample(model, mu, cov, al, T, seed=3, pdfs=pdfs, return_indices=True))
diff = torch.zeros((n, 256, 256), dtype=torch.int32, device=dev)
t_ras, _ = timed(lambda: E.rasterize_contours(out.reshape(-1, 21, 2), return_masks=False, diff=diff, per_image=T))
print(f"n_img {n} T {T}: tables {t_tab:.3f} ms ({n*21*65536*4/t_tab/1e6:.1f} GB/s written), cond {t_cond:.3f} ms, "
      f"sampler(+cond) {t_smp:.3f} ms = {n*T/t_smp*1e3:.3e} samples/s, raster {t_ras:.3f} ms, fallbacks {(idx<0).sum().item()}")

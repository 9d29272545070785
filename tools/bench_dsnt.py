"""Micro-benchmark of K1 (dsnt_moments) at the headline size and beyond, with and without the row-slab split."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench as B
from contouring_uncertainty_b200 import engine as E
dev = torch.device("cuda:0")
L = E.nat.lib(); st = torch.cuda.current_stream().cuda_stream
for n in (1, 8, 64, 512):
    logits = B.make_logits_on_device(torch, dev, n, seed=3)
    mu = torch.empty((n, 21, 2), device=dev); sig = torch.empty((n, 21, 2, 2), device=dev)
    mu0 = torch.empty_like(mu); sig0 = torch.empty_like(sig)
    ws = E.dsnt_workspace(n * 21, dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    res = {}
    for name, w, (m, s) in (("one CTA per heatmap", None, (mu0, sig0)), ("persistent row ranges", ws.data_ptr(), (mu, sig))):
        ts = []
        for i in range(25):
            if n < 64: flush.zero_()                       # inputs smaller than L2: flush between iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            E.nat.check(L.cub200_dsnt_moments(logits.data_ptr(), n * 21, 256, 256, 1, 1, None, m.data_ptr(), s.data_ptr(), w, st), "dsnt")
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[name] = float(np.median(ts[5:]))
    gb = n * 21 * (256 * 256 * 4 + 44) / 1e9
    print(f"n_img {n:4d}: " + ", ".join(f"{k} {v * 1e3:7.1f} us = {gb / (v * 1e-3):6.0f} GB/s" for k, v in res.items()),
          "| max |d mu|", float((mu - mu0).abs().max()), "max rel d sigma", float(((sig - sig0).abs() / sig0.abs().clamp_min(1e-3)).max()),
          "| tickets zero again:", bool((ws.view(-1, 528)[:, 512:516] == 0).all()))

#!/usr/bin/env python
"""Secondary bench lines for BASELINE.json configs 1, 3, 4, 5 (config 2 is bench.py's headline).

  python tools/bench_configs.py [--configs 1,3,4,5] [--steps K] [--scale S]

One JSON line per config (device-resident inputs, CUDA-event timing, 2 warm-up steps).  `--scale` shrinks the
per-GPU image / patient / pair counts (default 1.0 = the sizes SURVEY.md 8(d) fixes per GPU).
"""
import argparse, json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, synth
import bench as B
from contouring_uncertainty_b200 import engine as E

import torch.distributed as dist

RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(LOCAL)
dev = torch.device("cuda", LOCAL)
if WORLD > 1:      # one process per GPU (torchrun); every rank runs its own shard of images / patients / pairs (weak scaling)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))


def timed(fn, steps, warmup=2):
    """ms per step: barrier + synchronize on both sides, CUDA events, max over ranks."""
    for i in range(warmup):
        fn(i)
    if WORLD > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    if WORLD > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if WORLD > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def exchange(local: torch.Tensor):
    """The path's only collective: all_gather of the per-sample metric distributions (end of a step)."""
    if WORLD == 1:
        return local
    out = torch.empty((WORLD * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def line(cfg, workload, units, unit, ms, steps, **extra):
    if RANK != 0:
        return
    print(json.dumps({"config_id": cfg, "metric": "mc_contour_samples_per_s", "value": WORLD * units / (ms * 1e-3), "unit": unit,
                      "n_gpus": WORLD, "steps": steps, "warmup": 2, "ms_per_step": ms, "scaling": "weak", "data": "synthetic",
                      "config": {"workload": workload + (f" (x {WORLD} GPUs, image-sharded)" if WORLD > 1 else "")}, **extra}), flush=True)


def config1(steps, scale):
    """dsnt-al, 1 image, 100 independent-Gaussian samples -> uncertainty map."""
    eng = E.UncertaintyEngine(device=dev, sampler="naive")
    logits = B.make_logits_on_device(torch, dev, 1, seed=1001)
    ws = eng.workspace(1, 100)
    ms = timed(lambda i: eng.run_into(ws, logits, seed=i), max(steps, 50))
    line(1, "dsnt-al: 1 image x 100 independent-Gaussian samples (latency case)", 100, "samples/s", ms, max(steps, 50))


def config3(steps, scale):
    """dsnt-skew: 512 images x 1000 skew-PSM samples per GPU."""
    n, T = max(8, int(512 * scale)), 1000
    model = E.ShapeModelConstants(device=dev)
    logits = B.make_logits_on_device(torch, dev, n, seed=1003 + RANK)
    alpha = torch.from_numpy(synth.skew_alphas(n, seed=1103)).to(dev)
    diff = torch.zeros((n, 256, 256), dtype=torch.int32, device=dev)
    parts = {}

    def step(i):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        _, mu, cov = E.dsnt_moments(logits)
        ev[1].record()
        tabs = E.skew_pdf_grid(mu, cov, alpha, flip_alpha_y=True, sampled_only=True)
        ev[2].record()
        s = E.skew_sample(model, mu, cov, alpha, T, seed=i, pdfs=tabs)
        ev[3].record()
        diff.zero_()
        _, area = E.rasterize_contours(s.reshape(-1, 21, 2), return_masks=False, diff=diff, per_image=T)
        ev[4].record()
        E.finalize_maps(diff, T, want_count=False)
        ev[5].record()
        exchange(area.reshape(n, T))
        parts["ev"] = ev

    ms = timed(step, steps)
    ev = parts["ev"]
    km = {k: ev[a].elapsed_time(ev[a + 1]) for a, k in enumerate(("dsnt", "pdf_tables", "skew_sampler", "raster_aggregate", "finalize"))}
    line(3, f"dsnt-skew: {n} images x {T} skew-normal PSM samples per GPU", n * T, "samples/s", ms, steps, kernel_ms=km,
         images_per_gpu=n)


def config4(steps, scale):
    """uncertainty propagation: patients x (A2C, A4C) x (ED, ES), 10 000 sequence-PSM samples -> Simpson EDV/ESV/EF."""
    P_, T = max(1, int(16 * scale)), 10000
    model, seq = E.ShapeModelConstants(device=dev), E.SequenceModelConstants(device=dev)
    mu, cov = synth.point_distributions(P_ * 4, seed=1004)
    mu = mu.reshape(P_, 2, 2, 21, 2)                        # [patient][view][instant]
    ctr = mu[:, :, :1].mean(3, keepdims=True)
    mu[:, :, 1] = (ctr + (mu[:, :, :1] - ctr) * 0.8)[:, :, 0]
    mu_t = torch.from_numpy(mu.reshape(P_ * 2, 2, 21, 2).copy()).to(dev)
    cov_t = torch.from_numpy(cov.reshape(P_ * 2, 2, 21, 2, 2)).to(dev)
    out, parts = {}, {}

    def step(i):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        s = E.sequence_sample(model, seq, mu_t, cov_t, T, seed=i)                 # (2P, T, 2, 21, 2): view-pairs sampled jointly
        ev[1].record()
        c = s.permute(0, 2, 1, 3, 4).reshape(P_, 4, T, 21, 2).contiguous()        # [A2C-ED, A2C-ES, A4C-ED, A4C-ES]
        masks, area = E.rasterize_contours(c.reshape(-1, 21, 2), return_masks=True)
        ev[2].record()
        diam, stp = E.simpson_diameters(masks.reshape(P_, 4, T, 2048), 0.5)
        ev[3].record()
        out["v"] = E.simpson_volumes(diam, stp)
        ev[4].record()
        exchange(torch.stack(out["v"], -1))
        parts["ev"] = ev

    ms = timed(step, steps)
    ev = parts["ev"]
    km = {k: ev[a].elapsed_time(ev[a + 1]) for a, k in enumerate(("sequence_sampler", "raster_masks", "simpson_diameters", "volumes"))}
    edv, esv, ef = out["v"]
    line(4, f"uncertainty propagation: {P_} patients x 4 images x {T} sequence-PSM samples -> Simpson biplane EDV/ESV/EF",
         P_ * T, "patient-samples/s", ms, steps, kernel_ms=km, masks_per_s=P_ * 4 * T / (ms * 1e-3),
         ef_mean=float(ef.nanmean()), ef_nan_frac=float(ef.isnan().float().mean()))


def config5(steps, scale):
    """camus-cont_sequence: (ED, ES) pairs x 1000 sequence-PSM samples, streaming aggregation."""
    npair, T = max(4, int(256 * scale)), 1000
    model, seq = E.ShapeModelConstants(device=dev), E.SequenceModelConstants(device=dev)
    logits = B.make_logits_on_device(torch, dev, 2 * npair, seed=1005 + RANK)
    diff = torch.zeros((2 * npair, 256, 256), dtype=torch.int32, device=dev)

    def step(i):
        _, mu, cov = E.dsnt_moments(logits)
        s = E.sequence_sample(model, seq, mu.reshape(npair, 2, 21, 2), cov.reshape(npair, 2, 21, 2, 2), T, seed=i)
        c = s.permute(0, 2, 1, 3, 4).reshape(-1, 21, 2).contiguous()
        diff.zero_()
        _, area = E.rasterize_contours(c, return_masks=False, diff=diff, per_image=T)
        E.finalize_maps(diff, T, want_count=False)
        fac = E.fac_from_areas(area.reshape(npair, 2, T)[:, 0].contiguous(), area.reshape(npair, 2, T)[:, 1].contiguous())
        exchange(fac)

    ms = timed(step, steps)
    line(5, f"camus-cont_sequence: {npair} (ED, ES) pairs x {T} sequence-PSM samples per GPU -> maps + FAC", 2 * npair * T,
         "samples/s", ms, steps, pairs_per_gpu=npair)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    for c in a.configs.split(","):
        globals()["config" + c.strip()](a.steps, a.scale)
        torch.cuda.empty_cache()
    if WORLD > 1:
        dist.destroy_process_group()

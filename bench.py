#!/usr/bin/env python
"""bench.py -- MC contour samples/s -> masks + uncertainty maps @256x256 (BASELINE.json metric).

Workload (config 2 of BASELINE.json): per GPU, 64 images x 1000 PSM-correlated samples
(camus-cont_psm_11_no_std): heatmap logits (64,21,256,256) fp32 -> DSNT moments -> PSM factor ->
fused sample/rasterise/aggregate -> mean / entropy / pred maps + per-sample LV areas.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun for N > 1); weak scaling (every rank processes its own 64 images);
for N > 1 each step ends with the path's only exchange: an NCCL all_gather of the per-image area
distributions and an all_reduce of dataset-level sums.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_IMG, T_A, K, HW = 64, 1000, 21, 256
METRIC = "mc_contour_samples_per_s"
UNIT = "samples/s"
WORKLOAD = "dsnt-al2: 64 images x 1000 PSM samples (camus-cont_psm_11_no_std) per GPU, 256x256, K=21"

# Algorithmic HBM bytes (DESIGN.md "Kernels"): K1 reads each logit once and writes 44 B per heatmap;
# the fused kernel writes the contour sample (168 B) + area (4 B) per sample, reads the factor;
# finalize reads the diff map and writes count(i32) + mean(f64) + entropy(f64) + pred(u8).
BYTES_DSNT_PER_IMG = K * (HW * HW * 4 + 44)
BYTES_FUSED_PER_IMG = lambda T: T * (168 + 4) + 34 * 42 * 4 + 168
BYTES_FINAL_PER_IMG = HW * HW * (4 + 4 + 8 + 8 + 1)


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (profiles/r01k_ncu_full_summary_dsnt_psm_finalize.csv, profiles/r01i_ncu_full_summary_fused.csv)
NCU_TRAFFIC = {"dsnt": 352339200 + 3981824, "fused": 5050112 + 688128, "finalize": 17550848 + 35057920, "psm_factor": 88576}
ROOF_NOTES = {
    "fused": "NOT HBM-bound by construction (masks never leave shared memory): binding resources are issue slots / ALU "
             "pipe -- ncu: IPC 2.2 of 4 (56 % of issue slots), ALU pipe 40 %, 4912 warp-instructions per sample, shared memory "
             "43.07 M wavefronts per launch = 86 KB / sample = 28 % of the 148 SM x 1 wavefront/clk peak "
             "(DESIGN.md section 3, profiles/r01i_*, r01d_*)",
    "dsnt": "HBM-bound: every logit read once; DRAM traffic == algorithmic bytes; 1344 CTAs = 3.03 waves of the 444 resident "
            "ones, so ~10 % of this 69 us launch is tail -- the same kernel streams 6.7 TB/s at 512 images (profiles/r01r_configs_1_3_4_5.jsonl)",
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n_samples = args
    import synth
    from oracle import port

    logits = synth.heatmap_logits(1, seed=seed)
    rng = np.random.default_rng(seed + 5)
    z = rng.standard_normal((n_samples, 17, 2))
    t0 = time.perf_counter()
    mu, cov = port.dsnt_predict_on_batch(logits.astype(np.float32), dtype=np.float32)
    model = _cpu_worker.model
    s = port.psm_sample(model, mu[0], cov[0], z).astype(np.float32)
    r = port.uncertainty_from_samples(s)
    return time.perf_counter() - t0, int(r["area"].sum())


def _cpu_init():
    from oracle import port
    from contouring_uncertainty_b200.engine import DEFAULT_PSM

    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass
    _cpu_worker.model = port.ShapeModel.load(DEFAULT_PSM)


def cpu_arm(n_images, n_samples, workers):
    """Oracle port (numpy/scipy restatement of the reference) over `n_images` images x `n_samples`
    samples with `workers` processes; returns (samples/s, wall seconds)."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_cpu_init) as pool:
        pool.map(_cpu_worker, [(900 + i, 2) for i in range(workers)])   # warm the workers (imports, model)
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(1000 + i, n_samples) for i in range(n_images)], chunksize=1)
        wall = time.perf_counter() - t0
    cpu_arm.single_core = n_samples / float(np.mean([r[0] for r in res]))     # samples/s of ONE process on one core
    return n_images * n_samples / wall, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    n_images, n_samples = 2 * workers, 250         # bounded sample of the same workload per step
    times = []
    for i in range(args.warmup + args.steps):
        v, wall = cpu_arm(n_images, n_samples, workers)
        if i >= args.warmup:
            times.append(wall)
    t = float(np.mean(times))
    value = n_images * n_samples / t
    sample = f"{n_images} images x {n_samples} PSM samples per step (of 64 x 1000), {workers} processes x 1 thread"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 spline / u8 masks (numpy, scipy)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                         "single_core_value": cpu_arm.single_core},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def make_logits_on_device(torch, dev, n_img, seed):
    """Synthetic heatmap logits of SURVEY.md 8(d), generated on the device in fp32."""
    g = torch.Generator(device=dev).manual_seed(seed)
    base = torch.tensor(np.asarray(__import__("synth").LV_EXAMPLE_SHAPE), dtype=torch.float32, device=dev)
    mu = base[None] + 4.0 * torch.randn((n_img, 1, 2), generator=g, device=dev) + 1.5 * torch.randn((n_img, K, 2), generator=g, device=dev)
    th = torch.rand((n_img, K), generator=g, device=dev) * np.pi
    l1 = 4 + 32 * torch.rand((n_img, K), generator=g, device=dev)
    l2 = 4 + 32 * torch.rand((n_img, K), generator=g, device=dev)
    c, s = torch.cos(th), torch.sin(th)
    # precision matrix of R diag(l1,l2) R^T
    p00 = c * c / l1 + s * s / l2
    p01 = c * s * (1 / l1 - 1 / l2)
    p11 = s * s / l1 + c * c / l2
    ys, xs = torch.meshgrid(torch.arange(HW, device=dev, dtype=torch.float32), torch.arange(HW, device=dev, dtype=torch.float32), indexing="ij")
    out = torch.empty((n_img, K, HW, HW), dtype=torch.float32, device=dev)
    for i in range(n_img):
        dx = xs[None] - mu[i, :, 0, None, None]
        dy = ys[None] - mu[i, :, 1, None, None]
        out[i] = -0.5 * (p00[i, :, None, None] * dx * dx + 2 * p01[i, :, None, None] * dx * dy + p11[i, :, None, None] * dy * dy)
        out[i] += 0.05 * torch.randn((K, HW, HW), generator=g, device=dev)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200 import parallel as PAR

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    if not os.path.exists(E.nat.LIB_PATH):
        if rank == 0:
            ge.build()
        if world > 1:
            dist.barrier()

    eng = E.UncertaintyEngine(device=dev)
    logits = make_logits_on_device(torch, dev, N_IMG, seed=1002 + rank)
    stream = torch.cuda.current_stream(dev)
    sizes = PAR.shard_sizes(N_IMG * world, world)          # weak scaling: 64 images per rank

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kernel_ms = {"dsnt": [], "psm_factor": [], "fused": [], "finalize": []}
    ws = eng.workspace(N_IMG, T_A)
    # the path's only exchange (NCCL): per-image area distributions (all_gather) + dataset-level sums (all_reduce),
    # issued asynchronously so that it overlaps the next step's kernels; joined before the timed region ends
    exch = PAR.StepExchange(N_IMG, (T_A,), torch.int32, 2, dev, depth=2, stats_dtype=torch.int64) if world > 1 else None
    sums = torch.empty((2,), dtype=torch.int64, device=dev)

    def step(i):
        """One pass of the hot path over the batch: 6 launches of libcub200 into preallocated buffers."""
        eng.run_into(ws, logits, seed=1000 + i)
        if world > 1:
            exch.submit(i, ws.area, E.metric_sums(ws.area, out=sums))

    def step_evented(i):
        """Same kernels through the allocating per-op API with an event between them (per-kernel times)."""
        e = [ev() for _ in range(5)]
        e[0].record(stream)
        _, mu, cov = E.dsnt_moments(logits, is_logit=True, covar=True)
        e[1].record(stream)
        a_t = E.psm_factor(eng.model, mu, cov)
        e[2].record(stream)
        r = E.sample_raster_aggregate(mu, a_t, T_A, seed=1000 + i, return_contours=True, diff=ws.diff.zero_())
        e[3].record(stream)
        E.finalize_maps(r["diff"], T_A)
        e[4].record(stream)
        return e

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # nvidia-smi is started BEFORE the warm-up and given time to reach its polling loop: its start-up
    # takes driver locks that stall launches for milliseconds, which must not land in the timed region
    clocks = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        clocks.start()
        t_wait = time.time()
        while len(clocks.lines) < 3 and time.time() - t_wait < 5.0 and clocks.proc is not None:
            eng.run_into(ws, logits, seed=0)          # local work only: no collective outside lock-step
            torch.cuda.synchronize(dev)
    for i in range(args.warmup):
        step(i)
    if exch is not None:
        exch.finish()
    sync_all()
    t0, t1 = ev(), ev()
    t0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    if exch is not None:
        gathered, totals = exch.finish()     # the compute stream joins the last exchanges: inside the timed region
    t1.record(stream)
    sync_all()
    ms_total = t0.elapsed_time(t1)
    if exch is not None:
        assert gathered.shape == (world * N_IMG, T_A) and torch.equal(gathered[rank * N_IMG:(rank + 1) * N_IMG], ws.area)
        assert int(totals[0]) == int(gathered.sum())
    clk = clocks.stop() if (rank == 0 and not args.no_clocks) else None
    evs = [step_evented(i) for i in range(min(args.steps, 20))]
    sync_all()
    for e in evs:
        for name, a, b in (("dsnt", 0, 1), ("psm_factor", 1, 2), ("fused", 2, 3), ("finalize", 3, 4)):
            kernel_ms[name].append(e[a].elapsed_time(e[b]))
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    value = world * N_IMG * T_A / (ms_step * 1e-3)
    r = {"contours": ws.contours, "area": ws.area}

    # ---- e2e: public host-buffer API (HostPipeline): pinned host logits in, pinned host results out;
    #      H2D + kernels + D2H all inside the timed region, every step
    numa_bound = PAR.bind_to_gpu_numa_node(local) if world > 1 else False   # pinned buffers on the GPU's NUMA node
    pipe = E.HostPipeline(eng, N_IMG, T_A, chunk=8, n_streams=2)
    h_logits = torch.empty(logits.shape, dtype=torch.float32, pin_memory=True)
    h_logits.copy_(logits)
    h_out = pipe.host_outputs()
    h2d = h_logits.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in h_out.values())
    for i in range(3):
        pipe.run(h_logits, h_out, seed=5000 + i)
    sync_all()
    t0, t1 = ev(), ev()
    t0.record(stream)
    n_e2e = max(3, min(args.steps, 20))
    for i in range(n_e2e):
        pipe.run(h_logits, h_out, seed=6000 + i)
    t1.record(stream)
    sync_all()
    te = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N_IMG * T_A / (float(te.item()) / n_e2e * 1e-3)
    e2e_h2d_gbs = h2d / (float(te.item()) / n_e2e * 1e-3) / 1e9      # per rank; a lone pinned H2D copy reaches ~55 GB/s here
    e2e_check = bool(torch.equal(h_out["area"][-8:], pipe.ws[(N_IMG // 8 - 1) % 2].area.cpu()))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- parity spot check of the timed configuration (not timed): 1 image x 8 samples vs the oracle
    from oracle import port
    contours = r["contours"][0, :8].cpu().numpy()
    o = port.uncertainty_from_samples(contours)
    parity_ok = bool(np.array_equal(r["area"][0, :8].cpu().numpy(), o["area"]))

    peak, peak_src = _peaks()
    km = {k: float(np.median(v)) for k, v in kernel_ms.items()}
    dominant = max(km, key=km.get)
    alg_bytes = {"dsnt": N_IMG * BYTES_DSNT_PER_IMG, "psm_factor": N_IMG * (34 * 42 * 4 + 42 * 4 + 84 * 4),
                 "fused": N_IMG * BYTES_FUSED_PER_IMG(T_A), "finalize": N_IMG * BYTES_FINAL_PER_IMG}

    def roof(k):
        ach = alg_bytes[k] / (km[k] * 1e-3) / 1e9
        return {"kernel": k, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": NCU_TRAFFIC.get(k), "ms_per_launch": km[k], "algorithmic_bytes": alg_bytes[k],
                "peak_source": peak_src, "note": ROOF_NOTES.get(k)}

    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        workers = max(1, min(cores, 64))
        n_images, n_samples = 2 * workers, 250
        v, wall = cpu_arm(n_images, n_samples, workers)
        cpu = {"value": v, "unit": UNIT, "cores": workers, "kind": "port", "single_core_value": cpu_arm.single_core,
               "sample": f"{n_images} images x {n_samples} PSM samples (of 64 x 1000), {workers} processes x 1 thread, {wall:.1f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 moments / f64 spline / u32 bit-masks", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_gpu": N_IMG, "samples_per_image": T_A,
                   "l2": "inputs larger than L2 (352 MB of logits per step)", "rng": "in-kernel Philox4x32-10"},
        "contour_pixels_per_s": value * HW * HW,
        "kernel_ms": km,
        "roofline": roof(dominant),
        "roofline_dsnt": roof("dsnt"),
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "numa_bound": numa_bound,
                "h2d_gb_per_s_per_gpu": e2e_h2d_gbs, "bound": "PCIe host-to-device copy of the fp32 logits"},
        "gpu_launches": (5 + (1 if world > 1 else 0)) * args.steps,   # dsnt, psm_factor, fused, entropy_lut, finalize (+ metric_sums) per step
        "clocks": clk,
        "parity_spot_check": parity_ok and e2e_check,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

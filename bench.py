#!/usr/bin/env python
"""bench.py -- MC contour samples/s -> masks + uncertainty maps @256x256 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config {1,2,3,4,5}] [--impl reference] [--exchange-every E]

--config picks the BASELINE.json configuration (default 2, the one the metric is quoted on):
  1  dsnt-al    1 image x 100 independent-Gaussian samples                         (latency case)
  2  dsnt-al2   64 images x 1000 PSM-correlated samples per GPU                    (headline)
  3  dsnt-skew  512 images x 1000 skew-normal PSM samples per GPU (4096 over 8 GPUs)
  4  uncertainty propagation: 256 patients x (A2C, A4C) x (ED, ES) per GPU, 10 000 sequence-PSM samples per image
     -> Simpson's biplane EDV / ESV / EF distributions + aleatoric / epistemic split
  5  camus-cont_sequence: 256 (ED, ES) pairs x 1000 sequence-PSM samples per GPU, streaming aggregation + FAC
A "step" is one pass of the path over one batch; `value` = sampled contours rasterised + aggregated per second, whole job.

One process per GPU (torchrun for N > 1); weak scaling (every rank processes its own images / patients / pairs).  The
path's only exchange -- per-image metric distributions + exact dataset sums, what the reference's clinical processors read at
the END of the predict epoch (cu/results/clinical/clinical_metrics.py:51-66) -- is ONE NCCL all_gather per
`--exchange-every` steps (default 0 = one for the whole timed run), issued asynchronously and joined inside the timed region.
Prints ONE JSON line (rank 0).
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

K, HW = 21, 256
METRIC = "mc_contour_samples_per_s"
UNIT = "samples/s"

WORKLOADS = {
    1: "dsnt-al: 1 image x 100 independent-Gaussian samples, 256x256, K=21 (latency case)",
    2: "dsnt-al2: 64 images x 1000 PSM samples (camus-cont_psm_11_no_std) per GPU, 256x256, K=21",
    3: "dsnt-skew: 512 images x 1000 skew-normal PSM samples per GPU (4096 images over 8 GPUs), 256x256, K=21",
    4: "uncertainty propagation: 256 patients x 4 images (A2C/A4C x ED/ES) x 10000 sequence-PSM samples per GPU -> Simpson "
       "biplane EDV/ESV/EF distributions + aleatoric/epistemic split",
    5: "camus-cont_sequence: 256 (ED,ES) pairs x 1000 sequence-PSM samples per GPU -> maps + FAC",
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_notes():
    """Per-kernel DRAM traffic / issue figures of the committed `ncu --set full` captures (profiles/r02_ncu_kernels.json:
    dram__bytes_read.sum + dram__bytes_write.sum per launch, and what binds the kernels that are not HBM-bound)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_kernels.json")
    return json.load(open(p)) if os.path.exists(p) else {}


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
# CPU arm on the host cores: the VERBATIM reference (oracle/_ref, a git-ignored copy made by oracle/make_ref.py and run under
# the import stubs of tests/golden/_refstub.py) for configs 1, 2, 3 and 5 when the copy is present, else -- and for config 4 --
# the oracle port (numpy / scipy restatement of the reference).  `cpu_baseline.kind` says which.
# ------------------------------------------------------------------------------------------------
REF_CONFIGS = (1, 2, 3, 5)        # config 4's Simpson needs skimage.measure.find_contours (absent here): the port's restatement


def _ref_kind(cfg):
    try:
        from oracle import ref_runner
        return "reference" if cfg in REF_CONFIGS and ref_runner.available() else "port"
    except Exception:
        return "port"


def _cpu_init(kind="port"):
    from oracle import port
    from contouring_uncertainty_b200.engine import DEFAULT_PSM, DEFAULT_SEQ_PSM
    _cpu_worker.kind = kind
    if kind == "reference":
        from oracle import ref_runner
        ref_runner._load()

    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass
    _cpu_worker.model = port.ShapeModel.load(DEFAULT_PSM)
    _cpu_worker.seq = port.SequenceShapeModel.load(DEFAULT_SEQ_PSM) if hasattr(port.SequenceShapeModel, "load") else None


def _cpu_worker(args):
    """One unit of the bounded CPU sample of config `cfg`: returns (seconds, contours processed)."""
    cfg, seed, n_samples = args
    import synth
    from oracle import port

    rng = np.random.default_rng(seed + 5)
    model = _cpu_worker.model
    t0 = time.perf_counter()
    if cfg in (1, 2) and getattr(_cpu_worker, "kind", "port") == "reference":
        from oracle import ref_runner
        from contouring_uncertainty_b200.engine import DEFAULT_PSM
        logits = synth.heatmap_logits(1, seed=seed).astype(np.float32)
        np.random.seed(seed); __import__("torch").manual_seed(seed)
        dt, _ = ref_runner.image_unit(logits, n_samples, "naive" if cfg == 1 else "psm", DEFAULT_PSM)
        return dt, n_samples
    if cfg == 3 and getattr(_cpu_worker, "kind", "port") == "reference":
        from oracle import ref_runner
        from contouring_uncertainty_b200.engine import DEFAULT_PSM
        logits = synth.heatmap_logits(1, seed=seed).astype(np.float32)
        al = synth.skew_alphas(1, seed=seed + 1)
        np.random.seed(seed); __import__("torch").manual_seed(seed)
        dt, _ = ref_runner.skew_image_unit(logits, al[0], n_samples, DEFAULT_PSM)
        return dt, n_samples
    if cfg == 5 and getattr(_cpu_worker, "kind", "port") == "reference":
        from oracle import ref_runner
        from contouring_uncertainty_b200.engine import DEFAULT_PSM, DEFAULT_SEQ_PSM
        mus, covs = synth.point_distributions(2, seed=seed)
        ctr = mus[:1].mean(1, keepdims=True)
        mus[1] = (ctr + (mus[:1] - ctr) * 0.8)[0]             # ES = ED shrunk about its centroid
        np.random.seed(seed); __import__("torch").manual_seed(seed); __import__("random").seed(seed)
        dt, _ = ref_runner.pair_unit(mus, covs, n_samples, DEFAULT_PSM, DEFAULT_SEQ_PSM)
        return dt, 2 * n_samples
    if cfg in (1, 2):
        logits = synth.heatmap_logits(1, seed=seed)
        t0 = time.perf_counter()                              # the synthetic input is not part of the path
        mu, cov = port.dsnt_predict_on_batch(logits.astype(np.float32), dtype=np.float32)
        if cfg == 1:
            s = port.naive_sample(mu[0], cov[0], rng.standard_normal((n_samples, K, 2))).astype(np.float32)
        else:
            s = port.psm_sample(model, mu[0], cov[0], rng.standard_normal((n_samples, 17, 2))).astype(np.float32)
        port.uncertainty_from_samples(s)
        return time.perf_counter() - t0, n_samples
    if cfg == 3:
        logits = synth.heatmap_logits(1, seed=seed)
        al = synth.skew_alphas(1, seed=seed + 1)
        t0 = time.perf_counter()
        mu, cov = port.dsnt_predict_on_batch(logits.astype(np.float32), dtype=np.float32)
        z = rng.standard_normal((n_samples, 3, 3)).astype(np.float32)
        s = port.skew_psm_sample(model, mu[0], cov[0], al[0], z, lambda n_, d: synth.exponential_field(seed, n_ * 14 + d))
        port.uncertainty_from_samples(np.asarray(s, dtype=np.float32))
        return time.perf_counter() - t0, n_samples
    # 4 / 5: sequence sampler on (ED, ES) pairs
    seq = _cpu_worker.seq
    mus, covs = synth.point_distributions(4, seed=seed)
    mus = mus.reshape(2, 2, K, 2).copy()
    ctr = mus[:, :1].mean(2, keepdims=True)
    mus[:, 1] = (ctr + (mus[:, :1] - ctr) * 0.8)[:, 0]
    covs = covs.reshape(2, 2, K, 2, 2)
    t0 = time.perf_counter()
    views = []
    for v in range(2 if cfg == 4 else 1):
        s = port.seq_psm_sample(model, seq, mus[v], covs[v], rng.integers(0, 2, n_samples), rng.standard_normal((n_samples, 2, 17, 2)))
        views.append(port.convert_to_mask(np.asarray(s, dtype=np.float32).reshape(-1, K, 2)).reshape(n_samples, 2, HW, HW))
    if cfg == 4:
        m2, m4 = views
        port.volumes_mc(m2[:, 0].astype(bool), m2[:, 1].astype(bool), m4[:, 0].astype(bool), m4[:, 1].astype(bool), (0.5, 0.5), (0.5, 0.5))
        return time.perf_counter() - t0, 4 * n_samples
    m = views[0]
    port.sample_entropy(m[:, 0][:, None].astype(float)); port.sample_entropy(m[:, 1][:, None].astype(float))
    return time.perf_counter() - t0, 2 * n_samples


CPU_SAMPLES = {1: 100, 2: 250, 3: 4, 4: 4, 5: 100}          # samples per work unit: ~10-30 s of CPU work per step in total
CPU_UNITS_PER_WORKER = {1: 4, 2: 2, 3: 1, 4: 1, 5: 1}


def cpu_arm(cfg, workers, kind="port"):
    """The reference (kind "reference") or the oracle port over a bounded sample of config `cfg` with `workers` processes
    (one thread each).  Returns (contours/s, wall seconds, description)."""
    import multiprocessing as mp

    n_units, n_samples = CPU_UNITS_PER_WORKER[cfg] * workers, CPU_SAMPLES[cfg]
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_cpu_init, initargs=(kind,)) as pool:
        # warm the workers (imports, model); two samples: the reference's samplers squeeze a single one away (sampler.py:40)
        pool.map(_cpu_worker, [(cfg, 900 + i, 2) for i in range(workers)])
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(cfg, 1000 + i, n_samples) for i in range(n_units)], chunksize=1)
        wall = time.perf_counter() - t0
    contours = sum(r[1] for r in res)
    cpu_arm.single_core = float(np.mean([r[1] / r[0] for r in res]))       # contours/s of ONE process on one core
    unit = {1: "images", 2: "images", 3: "images", 4: "patients", 5: "(ED,ES) pairs"}[cfg]
    what = "the verbatim reference (oracle/_ref)" if kind == "reference" else "the oracle port"
    return contours / wall, wall, (f"{what}: {n_units} {unit} x {n_samples} samples per step (bounded sample of: {WORKLOADS[cfg]}), "
                                   f"{workers} processes x 1 thread")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    times, vals = [], []
    kind = _ref_kind(args.config)
    for i in range(args.warmup + args.steps):
        v, wall, sample = cpu_arm(args.config, workers, kind)
        if i >= args.warmup:
            times.append(wall); vals.append(v)
    t, value = float(np.mean(times)), float(np.mean(vals))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 torch DSNT + sampler / f64 spline / i64 masks (torch, numpy, scipy)" if kind == "reference" else "f64 spline / u8 masks (numpy, scipy)",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "config_id": args.config, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample,
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


BYTES_DSNT_PER_IMG = K * (HW * HW * 4 + 44)                   # every logit read once, 44 B of moments per heatmap


class Timer:
    """CUDA events on the current stream around named phases of ONE evented step."""

    def __init__(self, torch, stream):
        self.torch, self.stream, self.marks = torch, stream, []

    def mark(self, name):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.stream)
        self.marks.append((name, e))

    def phases(self):
        return {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(self.marks, self.marks[1:])}


class PathConfig:
    """One BASELINE configuration: device-resident step, per-phase timing, host-buffer (e2e) step."""
    id = 0

    def __init__(self, torch, E, dev, rank):
        self.torch, self.E, self.dev, self.rank = torch, E, dev, rank

    def contours_per_step(self):
        raise NotImplementedError

    def metric_values(self):
        """(values (n_local, n_per_image) int32, sums (2,) int64) of the last step: what the epoch-end exchange ships."""
        raise NotImplementedError


class LogitsPSM(PathConfig):
    """configs 1 and 2: logits -> DSNT -> factor -> sample / rasterise / aggregate -> maps (engine.run_into)."""

    def __init__(self, torch, E, dev, rank, cfg):
        super().__init__(torch, E, dev, rank)
        self.id = cfg
        self.n, self.T, sampler = (1, 100, "naive") if cfg == 1 else (64, 1000, "psm")
        self.eng = E.UncertaintyEngine(device=dev, sampler=sampler)
        # A rank cycles through POOL distinct batches (a predict epoch never sees the same images twice in a row); every rank
        # holds the same pool and starts at its own phase, so over any POOL steps all ranks have done the same work.  (With one
        # fixed batch per rank the max-over-ranks time measures which rank drew the heaviest 64 images -- the rasteriser's time
        # varies by ~5 % between batches, tools/shard_weight.py -- not the system: profiles/r02d_scaling_attribution.txt.)
        self.pool = [make_logits_on_device(torch, dev, self.n, seed=1000 + cfg + b) for b in range(self.POOL)]
        self.logits = self.pool[0]
        self.ws = self.eng.workspace(self.n, self.T)
        self.sums = torch.empty((2,), dtype=torch.int64, device=dev)

    POOL = 8
    gs = None

    def contours_per_step(self):
        return self.n * self.T

    def enable_graph(self, exch):
        """The step (+ the metric buffering of the exchange) as ONE CUDA-graph launch per step: one engine.GraphedStep per
        input batch of the pool, all advancing the same device-side step counter."""
        ring = exch.ring if exch is not None else None
        if not self.pipelined:
            first = self.E.GraphedStep(self.eng, self.ws, self.pool[0], seed=1000, metric_ring=ring)
            self.gs = [first] + [self.E.GraphedStep(self.eng, self.ws, lg, seed=1000, metric_ring=ring, step=first.step) for lg in self.pool[1:]]
            return
        # software-pipelined: the graph of batch b also runs the DSNT pass of batch b + 1 (side branch, into the other pair of
        # moment buffers) beside the factor / sampler / rasteriser / finalize of batch b -- every replay does one whole step's
        # work, the HBM-bound kernel of one batch under the issue-bound ones of the other (POOL is even: the pairs alternate)
        torch = self.torch
        self.moments = [(self.ws.mu, self.ws.cov), (torch.empty_like(self.ws.mu), torch.empty_like(self.ws.cov))]
        for b in (0, 1):
            self.eng.dsnt_into(self.ws, self.pool[b], *self.moments[b])
        self.gs, step = [], None
        for b in range(self.POOL):
            g = self.E.GraphedStep(self.eng, self.ws, self.pool[b], seed=1000, metric_ring=ring, step=step,
                                   prefetch=(self.pool[(b + 1) % self.POOL], self.moments[b % 2], self.moments[(b + 1) % 2]))
            step = g.step
            self.gs.append(g)
        self.primed_for = None

    pipelined = False
    primed_for = None          # the batch whose moments the pipelined graphs will find in place at the next replay

    def step(self, i):
        b = (i + self.rank) % self.POOL
        if self.gs is None:
            self.eng.run_into(self.ws, self.pool[b], seed=1000 + i)
            return
        if self.pipelined:
            if self.primed_for != b:                     # first step of a run of consecutive steps: nobody prefetched batch b
                self.eng.dsnt_into(self.ws, self.pool[b], *self.moments[b % 2])
            self.primed_for = (b + 1) % self.POOL
        self.gs[b].replay()

    def metric_values(self):
        return self.ws.area, self.E.metric_sums(self.ws.area, out=self.sums)

    def step_evented(self, i, tm):
        """The launches of engine.run_into, one by one, with an event between the phases (same buffers, no allocation)."""
        E, ws, eng = self.E, self.ws, self.eng
        self.primed_for = None                                        # this leg overwrites ws.mu / ws.cov
        logits = self.pool[(i + self.rank) % self.POOL]
        L, n, T = E.nat.lib(), self.n, self.T
        st = self.torch.cuda.current_stream(self.dev).cuda_stream
        chk = E.nat.check
        tm.mark("start")
        chk(L.cub200_dsnt_moments(logits.data_ptr(), n * K, HW, HW, 1, 1, None, ws.mu.data_ptr(), ws.cov.data_ptr(), ws.dsnt_ws.data_ptr(), st), "dsnt")
        tm.mark("dsnt")
        if eng.sampler == "psm":
            chk(L.cub200_psm_factor(eng.model.pack.data_ptr(), ws.mu.data_ptr(), ws.cov.data_ptr(), n, ws.a_t.data_ptr(), st), "factor")
        else:
            chk(L.cub200_naive_factor(ws.cov.data_ptr(), n, ws.a_t.data_ptr(), st), "factor")
        tm.mark("factor")
        chk(L.cub200_sample_raster_aggregate(ws.mu.data_ptr(), ws.a_t.data_ptr(), ws.zdim, None, 1000 + i, 0, n, T, 1, ws.contours.data_ptr(),
                                             None, None, 1, None, None, None, st), "sample")
        tm.mark("sample")
        ws.diff.zero_()
        chk(L.cub200_rasterize_contours(ws.contours.data_ptr(), n * T, 1, None, ws.area.data_ptr(), ws.diff.data_ptr(), T,
                                        ws.raster_ws.data_ptr(), st), "raster")
        tm.mark("raster")
        chk(L.cub200_finalize_maps(ws.diff.data_ptr(), n, T, E.nat.ptr(ws.count), E.nat.ptr(ws.mean), E.nat.ptr(ws.entropy), ws.pred.data_ptr(),
                                   ws.lut.data_ptr(), 1, E.nat.ptr(ws.variance), E.nat.ptr(ws.count16), st), "finalize")
        tm.mark("finalize")

    def algorithmic_bytes(self):
        n, T = self.n, self.T
        zd = 34 if self.id == 2 else 42
        return {"dsnt": n * BYTES_DSNT_PER_IMG, "factor": n * (zd * 42 * 4 + 42 * 4 + 84 * 4),
                "sample": n * (T * 168 + zd * 42 * 4 + 168), "raster": n * T * (168 + 4),
                "finalize": n * HW * HW * (4 + 4 + 8 + 8 + 1)}

    # ---- e2e: HostPipeline (pinned host logits in, pinned host results out, compact count map)
    def e2e_setup(self):
        torch, E = self.torch, self.E
        chunk = 1 if self.n == 1 else 8
        self.pipe = E.HostPipeline(self.eng, self.n, self.T, chunk=chunk, n_streams=1 if self.n == 1 else 2, compact=True)
        self.h_in = torch.empty(self.logits.shape, dtype=torch.float32, pin_memory=True)
        self.h_in.copy_(self.logits)
        self.h_out = self.pipe.host_outputs()
        return self.h_in.numel() * 4, sum(t.numel() * t.element_size() for t in self.h_out.values())

    def e2e_step(self, i):
        self.pipe.run(self.h_in, self.h_out, seed=6000 + i)

    def e2e_check(self):
        last = self.pipe.ws[((self.n // self.pipe.chunk) - 1) % len(self.pipe.ws)]
        ok = bool(self.torch.equal(self.h_out["area"][-self.pipe.chunk:], last.area.cpu()))
        m = self.pipe.expand(self.h_out)
        return ok and m["mean"].shape == (self.n, HW, HW) and float(m["mean"].max()) <= 1.0

    def parity_spot_check(self):
        from oracle import port
        contours = self.ws.contours[0, :8].cpu().numpy()
        o = port.uncertainty_from_samples(contours)
        return bool(np.array_equal(self.ws.area[0, :8].cpu().numpy(), o["area"]))


class Skew(PathConfig):
    """config 3: logits + skewness -> DSNT -> pdf tables -> numerical skew-PSM sampler -> rasterise / aggregate -> maps."""
    id = 3

    def __init__(self, torch, E, dev, rank, n=512, T=1000):
        super().__init__(torch, E, dev, rank)
        import synth
        self.n, self.T = n, T
        self.model = E.ShapeModelConstants(device=dev)
        self.logits = make_logits_on_device(torch, dev, n, seed=1003 + rank)
        self.alpha = torch.from_numpy(synth.skew_alphas(n, seed=1103 + rank)).to(dev)
        self.diff = torch.zeros((n, HW, HW), dtype=torch.int32, device=dev)
        self.sums = torch.empty((2,), dtype=torch.int64, device=dev)
        self.area = None

    def contours_per_step(self):
        return self.n * self.T

    def _run(self, i, tm=None, logits=None, alpha=None):
        E = self.E
        logits = self.logits if logits is None else logits
        alpha = self.alpha if alpha is None else alpha
        mark = tm.mark if tm is not None else (lambda name: None)
        mark("start")
        _, mu, cov = E.dsnt_moments(logits)
        mark("dsnt")
        tabs = E.skew_pdf_grid(mu, cov, alpha, flip_alpha_y=True, sampled_only=True)
        mark("pdf_tables")
        s = E.skew_sample(self.model, mu, cov, alpha, self.T, seed=i, pdfs=tabs)
        mark("skew_sampler")
        n = logits.shape[0]
        diff = self.diff[:n]
        diff.zero_()
        _, self.area = E.rasterize_contours(s.reshape(-1, K, 2), return_masks=False, diff=diff, per_image=self.T)
        mark("raster")
        out = E.finalize_maps(diff, self.T, want_count=False, want_maps=tm is not None or logits is self.logits, want_count16=True)
        mark("finalize")
        return s, out

    def step(self, i):
        self._run(i)

    def step_evented(self, i, tm):
        self._run(i, tm)

    def metric_values(self):
        return self.area.reshape(self.n, self.T), self.E.metric_sums(self.area, out=self.sums)

    def algorithmic_bytes(self):
        n, T = self.n, self.T
        return {"dsnt": n * BYTES_DSNT_PER_IMG, "pdf_tables": n * 17 * HW * HW * 4,
                "skew_sampler": n * (T * 168 + 17 * HW * HW * 4), "raster": n * T * (168 + 4),
                "finalize": n * HW * HW * (4 + 8 + 8 + 1 + 2)}

    def e2e_setup(self):
        torch = self.torch
        self.chunk = 64
        self.h_in = torch.empty(self.logits.shape, dtype=torch.float32, pin_memory=True)
        self.h_in.copy_(self.logits)
        self.h_alpha = self.alpha.cpu().pin_memory()
        self.h_cont = torch.empty((self.n, self.T, K, 2), dtype=torch.float32, pin_memory=True)
        self.h_area = torch.empty((self.n, self.T), dtype=torch.int32, pin_memory=True)
        self.h_c16 = torch.empty((self.n, HW, HW), dtype=torch.uint16, pin_memory=True)
        self.d_in = torch.empty((self.chunk, K, HW, HW), dtype=torch.float32, device=self.dev)
        return (self.h_in.numel() + self.h_alpha.numel()) * 4, self.h_cont.numel() * 4 + self.h_area.numel() * 4 + self.h_c16.numel() * 2

    def e2e_step(self, i):
        for c0 in range(0, self.n, self.chunk):
            sl = slice(c0, c0 + self.chunk)
            self.d_in.copy_(self.h_in[sl], non_blocking=True)
            al = self.h_alpha[sl].to(self.dev, non_blocking=True)
            s, out = self._run(i, logits=self.d_in, alpha=al)
            self.h_cont[sl].copy_(s, non_blocking=True)
            self.h_area[sl].copy_(self.area.reshape(self.chunk, self.T), non_blocking=True)
            self.h_c16[sl].copy_(out[-1], non_blocking=True)

    def e2e_check(self):
        torch = self.torch
        return bool(int(self.h_area[-1].min()) > 0 and int(self.h_c16[-1].to(torch.int32).max()) <= self.T)

    def parity_spot_check(self):
        from oracle import port
        s, _ = self._run(0)
        o = port.uncertainty_from_samples(s[0, :4].cpu().numpy())
        return bool(np.array_equal(self.area.reshape(self.n, self.T)[0, :4].cpu().numpy(), o["area"]))


class Sequence(PathConfig):
    """configs 4 and 5: sequence-PSM sampling of (ED, ES) pairs.
    5: logits of the pairs -> DSNT -> joint samples -> rasterise / aggregate -> maps + per-sample FAC.
    4: predicted moments of P patients x 2 views -> 10 000 joint samples per view -> bit-packed masks -> Simpson's biplane
       EDV / ESV / EF per sample index -> aleatoric / epistemic split per patient (chunks of patients: 8 KiB per mask)."""

    def __init__(self, torch, E, dev, rank, cfg):
        super().__init__(torch, E, dev, rank)
        import synth
        self.id = cfg
        self.model, self.seq = E.ShapeModelConstants(device=dev), E.SequenceModelConstants(device=dev)
        self.sums = torch.empty((2,), dtype=torch.int64, device=dev)
        if cfg == 5:
            self.npair, self.T = 256, 1000
            self.logits = make_logits_on_device(torch, dev, 2 * self.npair, seed=1005 + rank)
            self.diff = torch.zeros((2 * self.npair, HW, HW), dtype=torch.int32, device=dev)
        else:
            self.P, self.T, self.pchunk = 256, 10000, 16
            mu, cov = synth.point_distributions(self.P * 4, seed=1004 + rank)
            mu = mu.reshape(self.P, 2, 2, K, 2)                     # [patient][view][instant]
            ctr = mu[:, :, :1].mean(3, keepdims=True)
            mu[:, :, 1] = (ctr + (mu[:, :, :1] - ctr) * 0.8)[:, :, 0]
            self.h_mu = torch.from_numpy(mu.reshape(self.P * 2, 2, K, 2).copy()).pin_memory()
            self.h_cov = torch.from_numpy(cov.reshape(self.P * 2, 2, K, 2, 2).copy()).pin_memory()
            self.mu_t, self.cov_t = self.h_mu.to(dev), self.h_cov.to(dev)
            self.vol = torch.empty((self.P, 3, self.T), dtype=torch.float64, device=dev)      # EDV, ESV, EF per sample
            self.stats = torch.empty((self.P, 3, 4), dtype=torch.float64, device=dev)
            self.area4 = torch.empty((self.P, 4, self.T), dtype=torch.int32, device=dev)

    def contours_per_step(self):
        return 2 * self.npair * self.T if self.id == 5 else self.P * 4 * self.T

    # ---- config 5
    def _run5(self, i, tm=None, logits=None):
        E, T = self.E, self.T
        logits = self.logits if logits is None else logits
        npair = logits.shape[0] // 2
        mark = tm.mark if tm is not None else (lambda name: None)
        mark("start")
        _, mu, cov = E.dsnt_moments(logits)
        mark("dsnt")
        s = E.sequence_sample(self.model, self.seq, mu.reshape(npair, 2, K, 2), cov.reshape(npair, 2, K, 2, 2), T, seed=i)
        mark("sequence_sampler")
        c = s.permute(0, 2, 1, 3, 4).reshape(-1, K, 2).contiguous()
        diff = self.diff[:2 * npair]
        diff.zero_()
        _, area = E.rasterize_contours(c, return_masks=False, diff=diff, per_image=T)
        mark("raster")
        out = E.finalize_maps(diff, T, want_count=False, want_maps=logits is self.logits, want_count16=True)
        a = area.reshape(npair, 2, T)
        self.fac = E.fac_from_areas(a[:, 0].contiguous(), a[:, 1].contiguous())
        mark("finalize_fac")
        self.area = area
        return c, out

    # ---- config 4
    def _run4(self, i, tm=None, mu=None, cov=None):
        E, T, pc = self.E, self.T, self.pchunk
        mu = self.mu_t if mu is None else mu
        cov = self.cov_t if cov is None else cov
        acc = {}
        for p0 in range(0, self.P, pc):
            sub = Timer(self.torch, self.torch.cuda.current_stream(self.dev)) if tm is not None else None
            mark = sub.mark if sub is not None else (lambda name: None)
            mark("start")
            s = E.sequence_sample(self.model, self.seq, mu[2 * p0:2 * (p0 + pc)], cov[2 * p0:2 * (p0 + pc)], T, seed=i * 977 + p0)
            mark("sequence_sampler")
            c = s.permute(0, 2, 1, 3, 4).reshape(pc, 4, T, K, 2).contiguous()      # [A2C-ED, A2C-ES, A4C-ED, A4C-ES]
            masks, area = E.rasterize_contours(c.reshape(-1, K, 2), return_masks=True)
            mark("raster_masks")
            diam, stp = E.simpson_diameters(masks.reshape(pc, 4, T, 2048), 0.5)
            mark("simpson_diameters")
            edv, esv, ef = E.simpson_volumes(diam, stp)
            self.vol[p0:p0 + pc, 0], self.vol[p0:p0 + pc, 1], self.vol[p0:p0 + pc, 2] = edv, esv, ef
            self.area4[p0:p0 + pc] = area.reshape(pc, 4, T)
            mark("volumes")
            if sub is not None:
                acc.setdefault("subs", []).append(sub)
        if tm is not None:
            tm.subs = acc["subs"]
        # aleatoric / epistemic split of the three distributions of every patient (T_e = 1 here: the 10 000 samples are aleatoric)
        r = E.metric_ae_split(self.vol.reshape(self.P * 3, 1, T))
        self.stats.copy_(r["stats"].reshape(self.P, 3, 4))

    def step(self, i):
        (self._run5 if self.id == 5 else self._run4)(i)

    def step_evented(self, i, tm):
        (self._run5 if self.id == 5 else self._run4)(i, tm)

    def metric_values(self):
        if self.id == 5:
            return self.area.reshape(2 * self.npair, self.T), self.E.metric_sums(self.area, out=self.sums)
        a = self.area4.reshape(self.P * 4, self.T)
        return a, self.E.metric_sums(a, out=self.sums)

    def algorithmic_bytes(self):
        T = self.T
        if self.id == 5:
            n = 2 * self.npair
            # per sampled pair: two factors of 34 x 42 fp32 written + read (the second instant's shape model is re-centred per sample)
            return {"dsnt": n * BYTES_DSNT_PER_IMG, "sequence_sampler": self.npair * T * (2 * 34 * 42 * 4 + 2 * 168 + 400),
                    "raster": n * T * (168 + 4), "finalize_fac": n * HW * HW * (4 + 8 + 8 + 1 + 2) + self.npair * T * 16}
        n = self.P * 4
        return {"sequence_sampler": 2 * self.P * T * (2 * 34 * 42 * 4 + 2 * 168 + 400), "raster_masks": n * T * (168 + 8192 + 4),
                "simpson_diameters": n * T * (8192 + 168), "volumes": n * T * 168 + self.P * T * 24}

    def e2e_setup(self):
        torch = self.torch
        if self.id == 5:
            self.chunk = 64                                        # images per chunk (32 pairs)
            self.h_in = torch.empty(self.logits.shape, dtype=torch.float32, pin_memory=True)
            self.h_in.copy_(self.logits)
            n = 2 * self.npair
            self.h_cont = torch.empty((n, self.T, K, 2), dtype=torch.float32, pin_memory=True)
            self.h_area = torch.empty((n, self.T), dtype=torch.int32, pin_memory=True)
            self.h_c16 = torch.empty((n, HW, HW), dtype=torch.uint16, pin_memory=True)
            self.h_fac = torch.empty((self.npair, self.T), dtype=torch.float64, pin_memory=True)
            self.d_in = torch.empty((self.chunk, K, HW, HW), dtype=torch.float32, device=self.dev)
            return self.h_in.numel() * 4, self.h_cont.numel() * 4 + self.h_area.numel() * 4 + self.h_c16.numel() * 2 + self.h_fac.numel() * 8
        self.h_vol = torch.empty(self.vol.shape, dtype=torch.float64, pin_memory=True)
        self.h_stats = torch.empty(self.stats.shape, dtype=torch.float64, pin_memory=True)
        self.h_area4 = torch.empty(self.area4.shape, dtype=torch.int32, pin_memory=True)
        return (self.h_mu.numel() + self.h_cov.numel()) * 4, self.h_vol.numel() * 8 + self.h_stats.numel() * 8 + self.h_area4.numel() * 4

    def e2e_step(self, i):
        if self.id == 5:
            for c0 in range(0, 2 * self.npair, self.chunk):
                sl = slice(c0, c0 + self.chunk)
                self.d_in.copy_(self.h_in[sl], non_blocking=True)
                c, out = self._run5(i, logits=self.d_in)
                self.h_cont[sl].copy_(c.reshape(self.chunk, self.T, K, 2), non_blocking=True)      # already image-major
                self.h_area[sl].copy_(self.area.reshape(self.chunk, self.T), non_blocking=True)
                self.h_c16[sl].copy_(out[-1], non_blocking=True)
                self.h_fac[c0 // 2:(c0 + self.chunk) // 2].copy_(self.fac, non_blocking=True)
            return
        mu, cov = self.h_mu.to(self.dev, non_blocking=True), self.h_cov.to(self.dev, non_blocking=True)
        self._run4(i, mu=mu, cov=cov)
        self.h_vol.copy_(self.vol, non_blocking=True)
        self.h_stats.copy_(self.stats, non_blocking=True)
        self.h_area4.copy_(self.area4, non_blocking=True)

    def e2e_check(self):
        if self.id == 5:
            return bool(int(self.h_area[-1].min()) > 0 and float(np.nanmax(self.h_fac.numpy())) <= 1.0)
        ef = self.h_vol[:, 2].numpy()
        return bool(np.nanmin(ef) >= 0.0 and np.nanmax(ef) <= 1.0 and np.isfinite(self.h_stats[:, :2].numpy()).all())

    def parity_spot_check(self):
        from oracle import port
        if self.id == 5:
            c, _ = self._run5(0)
            o = port.uncertainty_from_samples(c[:4].cpu().numpy())
            return bool(np.array_equal(self.area[:4].cpu().numpy(), o["area"]))
        # one patient-sample: the four masks of sample 0 of patient 0 -> Simpson volumes by the oracle
        E = self.E
        s = E.sequence_sample(self.model, self.seq, self.mu_t[:2], self.cov_t[:2], 2, seed=5)
        c = s.permute(0, 2, 1, 3, 4).reshape(1, 4, 2, K, 2).contiguous()
        edv, esv, ef, _ = E.biplane_volumes_from_contours(c, 0.5)
        mk = port.convert_to_mask(c.cpu().numpy()[0, :, 0]).astype(bool)
        want = port.compute_left_ventricle_volumes(mk[0], mk[1], (0.5, 0.5), mk[2], mk[3], (0.5, 0.5))
        return (float(edv[0, 0]), float(esv[0, 0])) == tuple(float(v) for v in want)


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
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    if not os.path.exists(E.nat.LIB_PATH):
        if rank == 0:
            ge.build()
        if world > 1:
            dist.barrier()

    cfg_id = args.config
    cfg = (LogitsPSM(torch, E, dev, rank, cfg_id) if cfg_id in (1, 2) else Skew(torch, E, dev, rank) if cfg_id == 3
           else Sequence(torch, E, dev, rank, cfg_id))
    stream = torch.cuda.current_stream(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    every = args.exchange_every if args.exchange_every > 0 else args.steps
    vals0, _ = (cfg.step(0), cfg.metric_values())[1]
    exch = PAR.MetricExchange(vals0.shape[0], vals0.shape[1], every, dev) if (world > 1 or args.force_exchange) else None
    graphed = cfg_id in (1, 2) and not args.no_graph      # sub-millisecond steps: one graph launch instead of 8-12 launches
    if graphed:
        cfg.pipelined = not args.no_pipeline
        cfg.enable_graph(exch)

    def step(i):
        if exch is not None and graphed:
            exch.prepare()                                # a ring block being reused must have left the wire
            cfg.step(i)                                   # the graph writes the step's slot itself (device-side step counter)
            exch.note_pushed()
        else:
            cfg.step(i)
            if exch is not None:
                exch.push(cfg.metric_values()[0])

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
            cfg.eng.run_into(cfg.ws, cfg.logits, seed=0) if cfg_id in (1, 2) else cfg.step(0)   # local work only: no collective outside lock-step
            torch.cuda.synchronize(dev)
    for i in range(args.warmup):
        step(i)
    collectives0 = 0
    if exch is not None:
        exch.finish()                                     # flush the warm-up block; the timed run starts on a fresh block
        collectives0 = exch.collectives
        if graphed:
            cfg.gs[0].resync(exch.slot_index(), exch.ring.shape[0])
    sync_all()
    t0, t1 = ev(), ev()
    launches0 = E.nat.lib().cub200_launch_count()
    t0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    t_mid = ev()
    t_mid.record(stream)                     # end of this rank's own steps; what follows is the join of the exchange
    if exch is not None:
        gathered, totals = exch.finish()     # the compute stream joins the exchange: inside the timed region
    t1.record(stream)
    gpu_launches = int(E.nat.lib().cub200_launch_count() - launches0)   # kernels of libcub200 launched in the timed region
    if graphed:
        gpu_launches += args.steps * cfg.gs[0].kernels_per_replay          # ... by the graph replays (counted at capture)
    sync_all()
    ms_total = t0.elapsed_time(t1)
    n_collectives = 0
    if exch is not None:
        n_collectives = exch.collectives - collectives0
        mine, sums = cfg.metric_values()
        assert torch.equal(gathered[rank if world > 1 else 0, -1], mine), "exchange: own block of the last step"
        assert int(totals[-1, 0]) == int(gathered[:, -1].sum()), "exchange: dataset sum"
    clk = clocks.stop() if (rank == 0 and not args.no_clocks) else None
    per_rank = [ms_total / args.steps]
    per_rank_own = [t0.elapsed_time(t_mid) / args.steps]      # the rank's own steps, before it joins the exchange
    if world > 1:
        tl = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(tl, torch.tensor([per_rank[0], per_rank_own[0]], dtype=torch.float64, device=dev))
        per_rank = [float(x[0].item()) for x in tl]
        per_rank_own = [float(x[1].item()) for x in tl]
    if args.device_only:
        if rank == 0:
            print(f"device-only run: {max(per_rank):.4f} ms per step (per rank: {[round(x, 4) for x in per_rank]}; own steps before the "
                  f"exchange join: {[round(x, 4) for x in per_rank_own]}), {gpu_launches} library launches in the timed region")
        if world > 1:
            dist.destroy_process_group()
        return

    # per-phase device times of the same step (allocating per-op API, an event between the phases)
    n_ev = max(1, min(args.steps, 20 if cfg_id in (1, 2) else 3))
    timers = []
    for i in range(n_ev):
        tm = Timer(torch, stream)
        cfg.step_evented(args.warmup + i, tm)
        timers.append(tm)
    sync_all()
    if cfg_id == 4:
        ph = [[s.phases() for s in tm.subs] for tm in timers]
        km = {k: float(np.median([sum(p[k] for p in one) for one in ph])) for k in ph[0][0]}
    else:
        ph = [tm.phases() for tm in timers]
        km = {k: float(np.median([p[k] for p in ph])) for k in ph[0]}
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    units = cfg.contours_per_step()
    value = world * units / (ms_step * 1e-3)

    # ---- e2e: the host-buffer entry: pinned host inputs -> device -> pinned host results, every step, inside the timed region
    numa_bound = PAR.bind_to_gpu_numa_node(local) if world > 1 else False   # pinned buffers on the GPU's NUMA node
    h2d, d2h = cfg.e2e_setup()
    n_e2e = max(3, min(args.steps, 20 if cfg_id in (1, 2) else 3))
    for i in range(3 if cfg_id in (1, 2) else 1):
        cfg.e2e_step(i)
    sync_all()
    t0, t1 = ev(), ev()
    t0.record(stream)
    for i in range(n_e2e):
        cfg.e2e_step(10 + i)
    t1.record(stream)
    sync_all()
    te = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item()) / n_e2e
    e2e_value = world * units / (e2e_ms * 1e-3)
    e2e_ok = cfg.e2e_check()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    parity_ok = cfg.parity_spot_check()          # not timed: a few samples of the timed configuration vs the oracle
    peak, peak_src = _peaks()
    notes = _ncu_notes()
    alg = cfg.algorithmic_bytes()
    dominant = max(km, key=km.get)

    def roof(k):
        ach = alg[k] / (km[k] * 1e-3) / 1e9
        n = notes.get(f"config{cfg_id}", {}).get(k, notes.get("kernels", {}).get(k, {}))
        return {"kernel": n.get("kernel", k), "phase": k, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": n.get("dram_bytes_per_launch"), "traffic_source": n.get("source"),
                "ms_per_launch": km[k], "algorithmic_bytes": alg[k], "peak_source": peak_src, "binding_resource": n.get("binding"),
                "issue_active_pct": n.get("issue_active_pct"), "ipc": n.get("ipc"), "warp_instructions_per_launch": n.get("warp_instructions"),
                "note": n.get("note") or ("a read-only stream runs a little above the measured peak, which is a copy (read + write): "
                                          "B200_PROFILING.md" if ach > peak else None)}

    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        workers = max(1, min(cores, 64))
        kind = _ref_kind(cfg_id)
        v, wall, sample = cpu_arm(cfg_id, workers, kind)
        cpu = {"value": v, "unit": UNIT, "cores": workers, "kind": kind, "single_core_value": cpu_arm.single_core,
               "sample": f"{sample}, {wall:.1f} s"}
        if kind == "reference":                       # the (faster) restatement beside it, for continuity with round 1
            vp, wallp, _ = cpu_arm(cfg_id, workers, "port")
            cpu["port_value"], cpu["port_single_core_value"] = vp, cpu_arm.single_core

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 moments / f64 spline / u32 bit-masks", "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg_id], "config_id": cfg_id, "contours_per_step_per_gpu": units,
                   "l2": ("inputs larger than L2 (%.0f MB of logits per step)" % (cfg.logits.numel() * 4 / 1e6)) if hasattr(cfg, "logits") and cfg.logits.numel() * 4 > 130e6
                         else "per-step working set (sampled contours / masks written and re-read, %.0f MB) larger than L2" % (units * 168 * 2 / 1e6) if units * 168 * 2 > 130e6
                         else "working set smaller than L2 (latency case): no flush",
                   "inputs": (f"{cfg.POOL} distinct batches cycled per rank (same pool on every rank, phase-shifted by rank)" if cfg_id in (1, 2) else "one batch per rank"),
                   "rng": "in-kernel Philox4x32-10", "launch": ("CUDA graph replay (one launch per step); the DSNT pass of batch i + 1 on a side branch beside the sampler / rasteriser of batch i" if getattr(cfg, "pipelined", False) else "CUDA graph replay (one launch per step)") if graphed else "eager (one launch per kernel)", "exchange": f"one all_gather per {every} steps ({n_collectives} in the timed run)" if world > 1 else "none (1 GPU)"},
        "contour_pixels_per_s": value * HW * HW,
        "kernel_ms": km,
        "kernel_ms_note": ("eager launches, one after the other, an event between the phases; in the timed graph the DSNT pass of the next batch "
                           "overlaps the other kernels, so the step is shorter than their sum") if (graphed and getattr(cfg, "pipelined", False)) else
                          "eager launches, one after the other, an event between the phases",
        "roofline": roof(dominant),
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "numa_bound": numa_bound, "h2d_gb_per_s_per_gpu": h2d / (e2e_ms * 1e-3) / 1e9,
                "result": "compact: uint16 count map (mean / entropy / pred are exact functions of it), contours, areas"},
        "gpu_launches": gpu_launches,
        "ms_per_step_per_rank": per_rank, "ms_per_step_per_rank_before_exchange_join": per_rank_own,
        "clocks": clk,
        "parity_spot_check": bool(parity_ok and e2e_ok),
    }
    if "dsnt" in km and dominant != "dsnt":
        line["roofline_dsnt"] = roof("dsnt")
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--exchange-every", type=int, default=0, help="steps per all_gather of the metric distributions (0 = one per run)")
    ap.add_argument("--no-graph", action="store_true", help="configs 1 / 2: launch the step's kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--no-pipeline", action="store_true", help="configs 1 / 2: every graph runs its own batch's DSNT pass first instead of finding it prefetched by the previous step's graph (engine.GraphedStep(prefetch=...): 1.1 %% on one B200, 2.8 %% with the exchange buffer at 2 GPUs)")
    ap.add_argument("--force-exchange", action="store_true", help="diagnostic: run the per-step metric push / block shipping on one GPU too")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi (profiling runs)")
    ap.add_argument("--device-only", action="store_true", help="profiling runs: only the device-resident steps (no per-phase, e2e, CPU or parity legs; no JSON line)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {1: 200, 2: 96, 3: 10, 4: 3, 5: 10}[args.config]      # config 2: a multiple of the batch pool (8) if args.impl == "ours" else 1
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

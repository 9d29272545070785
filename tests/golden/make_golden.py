"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE.

Run by hand in the build container (needs /root/reference):

    python tests/golden/make_golden.py

It imports the unmodified reference through ``_refstub`` (import stubs for absent plotting /
training packages, a restated ``skimage.draw.line``), feeds it seeded inputs and recorded random
draws and stores inputs + outputs as small ``.npz`` files.  The GPU box has no /root/reference:
tests there compare the oracle (``oracle/port.py``) and the CUDA path against these files.

Inputs that are large (heatmap logits) are NOT stored; ``synth.py`` regenerates them from the seed.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import _refstub  # noqa: E402

_refstub.install()

import torch  # noqa: E402

import synth  # noqa: E402  (tests/synth.py: seeded synthetic inputs shared by tests / bench)

PSM_PATH = os.path.join(_refstub.REFERENCE_ROOT, "camus-cont_psm_11_no_std.npy")


def golden_dsnt():
    """flat_softmax + dsnt + DSNTAleatoric.predict_on_batch on synthetic Gaussian-bump logits."""
    from contour_uncertainty.task.regression.dsnt.utils import flat_softmax, dsnt
    from contour_uncertainty.task.regression.dsnt.dsnt_al import DSNTAleatoric
    from contour_uncertainty.task.regression.aleatoric import AleatoricUncertaintyTask

    out = {}
    for name, (n, seed, kind) in {"bumps": (3, 11, "bumps"), "noise": (2, 12, "noise"),
                                  "bimodal": (2, 13, "bimodal")}.items():
        logits = synth.heatmap_logits(n, seed=seed, kind=kind)
        t = torch.from_numpy(logits)
        coords, var, covar = dsnt(flat_softmax(t))
        fake = SimpleNamespace(hparams=SimpleNamespace(covar=True), device="cpu")
        fake.get_cov_matrix = lambda vx, vy, c=0, _f=fake: AleatoricUncertaintyTask.get_cov_matrix(_f, vx, vy, c)
        px, sig = DSNTAleatoric.predict_on_batch(fake, t, lambda x: x)
        out[f"{name}_n"] = n
        out[f"{name}_seed"] = seed
        out[f"{name}_coords"] = coords.numpy()
        out[f"{name}_var"] = var.numpy()
        out[f"{name}_covar"] = covar.numpy()
        out[f"{name}_pixel_mu"] = px.numpy()
        out[f"{name}_pixel_sigma"] = sig.numpy()
    np.savez_compressed(os.path.join(HERE, "dsnt.npz"), **out)
    print("dsnt.npz", {k: np.asarray(v).shape for k, v in out.items() if k.endswith("coords")})


def golden_raster():
    """USContourToMask / reconstruction on random + adversarial contours; bit-packed masks."""
    from contour_uncertainty.data.camus.utils import USContourToMask
    from contour_uncertainty.utils.contour import reconstruction

    contours, tags = synth.raster_cases(seed=21, n_random=160)
    masks = np.zeros((len(contours), 256, 256), dtype=np.uint8)
    for i, c in enumerate(contours):
        masks[i] = USContourToMask.__call__(c.astype(np.float32), (256, 256), labels=[], apply_argmax=True)
    # the demo scripts call reconstruction() on UNROUNDED float contours (sampling_example3.py:127)
    unr = np.zeros((32, 256, 256), dtype=np.uint8)
    for i in range(32):
        unr[i] = reconstruction(contours[i].astype(np.float64), 256, 256)
    # reconstruction_type='linear' (validation-step dice of dsnt_al.py:80-84): polygon outline + hole fill
    lin = np.zeros((len(contours), 256, 256), dtype=np.uint8)
    for i, c in enumerate(contours):
        lin[i] = USContourToMask.__call__(c.astype(np.float32), (256, 256), labels=[], apply_argmax=True, reconstruction_type='linear')
    np.savez_compressed(os.path.join(HERE, "raster.npz"), contours=np.stack(contours).astype(np.float32),
                        tags=np.array(tags), masks=np.packbits(masks, axis=-1, bitorder="little"),
                        masks_unrounded=np.packbits(unr, axis=-1, bitorder="little"),
                        masks_linear=np.packbits(lin, axis=-1, bitorder="little"))
    print("raster.npz", masks.shape, "areas", masks.sum((1, 2))[:8], "tags", sorted(set(tags)))


def golden_psm():
    """PosteriorShapeModelSampler.__call__ with replayed standard-normal draws (fp32 reference)."""
    from contour_uncertainty.sampler.posterior_shape_model.psm import PosteriorShapeModelSampler

    sampler = PosteriorShapeModelSampler(PSM_PATH)
    rng = np.random.default_rng(31)
    n_img, n_s = 4, 12
    mus, covs = synth.point_distributions(n_img, seed=32)
    z = rng.standard_normal((n_img, n_s, 17, 2)).astype(np.float32)
    out = np.zeros((n_img, n_s, 21, 2), dtype=np.float32)
    for i in range(n_img):
        draws = [z[i, s, d][None] for s in range(n_s) for d in range(17)]
        with _refstub.replay_standard_normal(draws) as log:
            res = sampler(torch.from_numpy(mus[i]), torch.from_numpy(covs[i]), n=n_s)
        assert len(log) == n_s * 17 and all(sh == (1, 2) for sh in log), log[:3]
        out[i] = res.detach().numpy()
    # The same UNMODIFIED reference code executed in fp64 (default dtype + members cast to double):
    # removes the noise of its fp32 42x42 inverses (sigma2=1e-3 fill: up to ~3e-2 px) and is the
    # tight anchor for the oracle (agreement ~1e-10 px).
    torch.set_default_dtype(torch.float64)
    try:
        s64 = PosteriorShapeModelSampler(PSM_PATH)
        for a in ("mean", "scale", "X_train", "X_val", "mu", "Q"):
            setattr(s64, a, getattr(s64, a).double())
        out64 = np.zeros((n_img, n_s, 21, 2), dtype=np.float64)
        for i in range(n_img):
            draws = [z[i, s, d][None].astype(np.float64) for s in range(n_s) for d in range(17)]
            with _refstub.replay_standard_normal(draws):
                res = s64(torch.from_numpy(mus[i]).double(), torch.from_numpy(covs[i]).double(), n=n_s)
            assert res.dtype == torch.float64
            out64[i] = res.detach().numpy()
    finally:
        torch.set_default_dtype(torch.float32)
    print("psm fp32-vs-fp64 reference max diff (px)", np.abs(out64 - out).max())
    np.savez_compressed(os.path.join(HERE, "psm.npz"), mu=mus, cov=covs, z=z, samples=out, samples_f64=out64,
                        initial_points=np.array(sampler.initial_points),
                        order=np.array(sum(sampler.points_order, [])))
    print("psm.npz", out.shape, "max |s-mu|", np.abs(out - mus[:, None]).max())


def golden_psm_alpha():
    """PosteriorShapeModelSampler.__call__ WITH alpha (psm.py:231-237, 385-419): the three initial points come from
    BivariateSkewNormal.rvs_fast(mu_j, cov_j, alpha_j * [1, -1]) (one (1,3) normal draw each), the 14 level points from the
    Gaussian hierarchy ((1,2) draws).  fp32 as the reference runs it and the same code in fp64."""
    from contour_uncertainty.sampler.posterior_shape_model.psm import PosteriorShapeModelSampler

    rng = np.random.default_rng(36)
    n_img, n_s = 3, 8
    mus, covs = synth.point_distributions(n_img, seed=37)
    alphas = synth.skew_alphas(n_img, seed=38)
    z3 = rng.standard_normal((n_img, n_s, 3, 3)).astype(np.float32)
    z2 = rng.standard_normal((n_img, n_s, 14, 2)).astype(np.float32)
    outs = {}
    for mode in ("f32", "f64"):
        dt = np.float64 if mode == "f64" else np.float32
        if mode == "f64":
            torch.set_default_dtype(torch.float64)
        try:
            smp = PosteriorShapeModelSampler(PSM_PATH)
            if mode == "f64":
                for a in ("mean", "scale", "X_train", "X_val", "mu", "Q"):
                    setattr(smp, a, getattr(smp, a).double())
            res = np.zeros((n_img, n_s, 21, 2))
            for i in range(n_img):
                draws = []
                for s_ in range(n_s):
                    draws += [z3[i, s_, d][None].astype(dt) for d in range(3)] + [z2[i, s_, d][None].astype(dt) for d in range(14)]
                with _refstub.replay_standard_normal(draws) as log, torch.no_grad():
                    t = lambda a: torch.from_numpy(a.astype(dt))
                    r = smp(t(mus[i]), t(covs[i]), t(alphas[i]), n=n_s)
                assert len(log) == 17 * n_s and log[0] == (1, 3) and log[3] == (1, 2), log[:4]
                res[i] = r.detach().numpy()
            outs[mode] = res
        finally:
            torch.set_default_dtype(torch.float32)
    print("psm_alpha fp32-vs-fp64 max diff (px)", np.abs(outs["f32"] - outs["f64"]).max())
    np.savez_compressed(os.path.join(HERE, "psm_alpha.npz"), mu=mus, cov=covs, alpha=alphas, z3=z3, z2=z2,
                        samples=outs["f32"].astype(np.float32), samples_f64=outs["f64"])
    print("psm_alpha.npz", outs["f64"].shape)



def golden_seq():
    """SequencePSMSampler.__call__ with replayed draws and first-instant bits (fp32 and fp64 runs)."""
    import random

    from contour_uncertainty.sampler.posterior_shape_model.sequence_sampler import SequencePSMSampler

    SEQ_PATH = os.path.join(_refstub.REFERENCE_ROOT, "camus-cont_sequence_psm_11_no_std.npy")
    rng = np.random.default_rng(51)
    n_pair, n_s = 3, 8
    mus, covs = synth.point_distributions(2 * n_pair, seed=52)
    mus = mus.reshape(n_pair, 2, 21, 2).copy()
    ctr = mus[:, 1].mean(1, keepdims=True)
    mus[:, 1] = ctr + (mus[:, 1] - ctr) * 0.8                 # ES = ED shrunk about its centroid
    covs = covs.reshape(n_pair, 2, 21, 2, 2)
    z = rng.standard_normal((n_pair, n_s, 2, 17, 2)).astype(np.float32)
    bits = rng.integers(0, 2, (n_pair, n_s))
    outs = {}
    for mode in ("f32", "f64"):
        orig_float = torch.Tensor.float
        if mode == "f64":
            torch.set_default_dtype(torch.float64)
            torch.Tensor.float = lambda self, *a, **k: self.double()   # sequence_sampler.py:85 casts mu_c/cov_c to fp32
        try:
            smp = SequencePSMSampler(PSM_PATH, SEQ_PATH)
            if mode == "f64":
                for a in ("mean", "scale", "X_train", "X_val", "mu", "Q", "seq_mu", "seq_Q", "seq_mean", "seq_scale"):
                    setattr(smp, a, getattr(smp, a).double())
            res = np.zeros((n_pair, n_s, 2, 21, 2))
            for i in range(n_pair):
                draws = [z[i, s, c, d][None].astype(np.float64 if mode == "f64" else np.float32)
                         for s in range(n_s) for c in range(2) for d in range(17)]
                it = iter(bits[i].tolist())
                orig = random.randint
                random.randint = lambda a, b, _it=it: int(next(_it))
                try:
                    with _refstub.replay_standard_normal(draws) as log:
                        m = torch.from_numpy(mus[i]); c = torch.from_numpy(covs[i])
                        if mode == "f64":
                            m, c = m.double(), c.double()
                        r = smp(m, c, n=n_s)
                finally:
                    random.randint = orig
                assert len(log) == n_s * 34, len(log)
                res[i] = r.detach().numpy()
            outs[mode] = res
        finally:
            torch.set_default_dtype(torch.float32)
            torch.Tensor.float = orig_float
    print("seq fp32-vs-fp64 reference max diff (px)", np.abs(outs["f32"] - outs["f64"]).max())
    np.savez_compressed(os.path.join(HERE, "seq.npz"), mu=mus, cov=covs, z=z, bits=bits,
                        samples=outs["f32"].astype(np.float32), samples_f64=outs["f64"])
    print("seq.npz", outs["f64"].shape)


def _skew_inputs(n_s):
    """3 images x n_s samples.  Image 2 has tight covariances and its last sample replays huge normals
    for the initial points, so that the shape model's conditional means land far from the predicted
    landmarks: the product pdf then underflows to 0 everywhere in fp32, torch.multinomial raises and
    the reference falls back to mu_c (psm_skew.py:135-154)."""
    mus, covs = synth.point_distributions(3, seed=61)
    alphas = synth.skew_alphas(3, seed=62)
    covs[2] *= 0.15
    z = np.random.default_rng(64).standard_normal((3, n_s, 3, 3)).astype(np.float32)
    z[2, n_s - 1] = np.array([[1.0, 25.0, -30.0], [0.5, -28.0, 22.0], [2.0, 30.0, 30.0]], dtype=np.float32)
    return mus, covs, alphas, z


def golden_skew():
    """SkewPosteriorShapeModelSampler.__call__ (psm_skew.py) with the (1,3) normals of rvs_fast replayed
    and torch.multinomial's exponential race fed from synth.exponential_field; fp32 and fp64 runs."""
    from contour_uncertainty.sampler.posterior_shape_model.psm_skew import SkewPosteriorShapeModelSampler
    from contour_uncertainty.distributions.bivariateskewnormal import BivariateSkewNormal

    # torch.multinomial(p, 1) IS argmax(p / Exp(1)) on this torch build (ATen single-draw path)
    gen = torch.Generator().manual_seed(123)
    for trial in range(20):
        p = torch.rand(65536, generator=gen) ** 8
        torch.manual_seed(1000 + trial)
        want = torch.multinomial(p, 1)
        torch.manual_seed(1000 + trial)
        q = torch.empty_like(p).exponential_(1)
        assert int(torch.argmax(p / q)) == int(want), "multinomial != exponential race"

    Q_SEED = 63
    n_img, n_s = 3, 4
    mus, covs, alphas, z = _skew_inputs(n_s)
    real_multinomial = torch.multinomial
    outs, idxs, nfallback = {}, {}, {}
    for mode in ("f32", "f64"):
        if mode == "f64":
            torch.set_default_dtype(torch.float64)
        try:
            smp = SkewPosteriorShapeModelSampler(PSM_PATH)
            if mode == "f64":
                for a in ("mean", "scale", "X_train", "mu", "Q"):
                    setattr(smp, a, getattr(smp, a).double())
            res = np.zeros((n_img, n_s, 21, 2))
            index_log = []
            for i in range(n_img):
                state = {"draw": 0}

                def fake_multinomial(p, n, *a, _i=i, _st=state, **k):
                    d = _st["draw"]
                    _st["draw"] += 1
                    stream = (_i * n_s + d // 14) * 14 + d % 14
                    try:
                        real_multinomial(p, n)          # raises exactly when the reference's would
                    except Exception:
                        index_log.append(-1)
                        raise
                    q = torch.from_numpy(synth.exponential_field(Q_SEED, stream)).to(p.dtype)
                    idx = torch.argmax(p / q).reshape(1)
                    index_log.append(int(idx))
                    return idx

                draws = [z[i, s, d][None].astype(np.float64 if mode == "f64" else np.float32)
                         for s in range(n_s) for d in range(3)]
                torch.multinomial = fake_multinomial
                try:
                    with _refstub.replay_standard_normal(draws) as log:
                        m, c, al = (torch.from_numpy(a[i:i + 1]) for a in (mus, covs, alphas))
                        if mode == "f64":
                            m, c, al = m.double(), c.double(), al.double()
                        r = smp(m, c, al, n=n_s)
                finally:
                    torch.multinomial = real_multinomial
                assert len(log) == 3 * n_s and all(sh == (1, 3) for sh in log), log
                assert state["draw"] == 14 * n_s
                res[i] = r[0].detach().numpy()
            outs[mode] = res
            idxs[mode] = np.array(index_log).reshape(n_img, n_s, 14)
            nfallback[mode] = int((idxs[mode] < 0).sum())
            if mode == "f32":
                al_flip = torch.from_numpy(alphas) * torch.tensor([1., -1.])
                pdf_sub = np.stack([np.stack([BivariateSkewNormal.pdf(smp.grid_points, torch.from_numpy(mus[i, k]),
                                                                      torch.from_numpy(covs[i, k]), al_flip[i, k]).numpy()[3::16, 5::16]
                                              for k in range(21)]) for i in range(n_img)])
        finally:
            torch.set_default_dtype(torch.float32)
    same = (idxs["f32"] == idxs["f64"]).mean()
    print("skew: fallbacks", nfallback, "fp32-vs-fp64 identical draws", same,
          "max diff (px)", np.abs(outs["f32"] - outs["f64"]).max())
    np.savez_compressed(os.path.join(HERE, "skew.npz"), mu=mus, cov=covs, alpha=alphas, z=z, q_seed=Q_SEED,
                        samples=outs["f32"].astype(np.float32), samples_f64=outs["f64"],
                        indices=idxs["f32"], indices_f64=idxs["f64"], pdf_sub=pdf_sub.astype(np.float32))
    print("skew.npz", outs["f32"].shape)


def _simpson_contours(T):
    """(4, T, 21, 2) fp32 contour samples of one patient: [A2C-ED, A2C-ES, A4C-ED, A4C-ES] (ES = ED shrunk
    to 0.8 about its centroid, A4C = A2C x 1.05) + per-sample jitter; the last samples are
    adversarial: a tiny blob, an all-empty mask (degenerate contour far off-frame), a contour pushed
    against the right frame border (open marching-squares contour) and a self-intersecting one."""
    rng = np.random.default_rng(91)
    b = synth.LV_EXAMPLE_SHAPE
    ctr = b.mean(0)
    base = [b, ctr + (b - ctr) * 0.8, ctr + (b - ctr) * 1.05, ctr + (b - ctr) * 0.84]
    out = np.zeros((4, T, 21, 2))
    for v in range(4):
        for t in range(T):
            out[v, t] = base[v] + rng.normal(0, 2.5, (21, 2)) + rng.normal(0, 3.0, (1, 2))
    out[1, T - 4] = ctr + (b - ctr) * 0.001                     # one-pixel ES blob: no admissible chord -> exception -> zeros
    out[0, T - 3] = b + np.array([-700.0, 0.0])                 # IndexError path: nothing but the clipped closing line
    out[2, T - 2] = b + np.array([85.0, 0.0])                   # clipped at the right border
    out[3, T - 1] = b + rng.normal(0, 14.0, (21, 2))            # self-intersecting
    return out.astype(np.float32)


def golden_simpson():
    """compute_left_ventricle_volumes / _compute_diameters (vital heart_us.py) and the Monte-Carlo loop of
    Volume.compute (cu/results/clinical/patient.py:166-190) on rasterised contour samples.
    skimage.measure.find_contours is the restated one (oracle/port.py): parity unpinned for it."""
    from contour_uncertainty.data.camus.utils import USContourToMask
    from vital.metrics.evaluate.clinical.heart_us import compute_left_ventricle_volumes, _compute_diameters

    T = 10
    contours = _simpson_contours(T)
    vs = (0.5, 0.5)
    masks = np.array([[USContourToMask.__call__(contours[v, t], (256, 256), labels=[], apply_argmax=True) for t in range(T)]
                      for v in range(4)]).astype(bool)
    diam = np.full((4, T, 20), np.nan)
    step = np.full((4, T), np.nan)
    for v in range(4):
        for t in range(T):
            try:
                diam[v, t], step[v, t] = _compute_diameters(masks[v, t], vs)
            except Exception as e:  # the reference's callers swallow these
                print("  _compute_diameters raised", v, t, type(e).__name__)
    edv, esv, ef = np.zeros(T), np.zeros(T), np.zeros(T)
    for t in range(T):                                          # patient.py:166-187
        try:
            lv_edv, lv_esv = compute_left_ventricle_volumes(masks[0, t], masks[1, t], vs, masks[2, t], masks[3, t], vs)
            edv[t] = lv_edv
            esv[t] = lv_esv
            ef[t] = (lv_edv - lv_esv) / lv_edv
        except:  # noqa: E722  (as the reference)
            pass
    ef[np.logical_or(ef < 0, ef > 1)] = np.nan
    # anisotropic voxel spacing (the real CAMUS case): PIL nearest-neighbour stretch inside _compute_diameters
    vs2, vs4 = (0.5, 0.35), (0.31, 0.42)
    diam_a = np.full((4, T, 20), np.nan)
    step_a = np.full((4, T), np.nan)
    for v in range(4):
        for t in range(T):
            try:
                diam_a[v, t], step_a[v, t] = _compute_diameters(masks[v, t], vs2 if v < 2 else vs4)
            except Exception as e:
                print("  _compute_diameters (anisotropic) raised", v, t, type(e).__name__)
    edv_a, esv_a = np.zeros(T), np.zeros(T)
    for t in range(T):
        try:
            edv_a[t], esv_a[t] = compute_left_ventricle_volumes(masks[0, t], masks[1, t], vs2, masks[2, t], masks[3, t], vs4)
        except:  # noqa: E722
            pass
    np.savez_compressed(os.path.join(HERE, "simpson.npz"), contours=contours, voxelspacing=np.array(vs),
                        areas=masks.sum((2, 3)), diameters=diam, step=step, edv=edv, esv=esv, ef=ef,
                        vs_a2c=np.array(vs2), vs_a4c=np.array(vs4), diameters_aniso=diam_a, step_aniso=step_a,
                        edv_aniso=edv_a, esv_aniso=esv_a)
    print("simpson.npz edv", edv, "esv", esv, "ef", ef)


def golden_umap():
    """projected_uncertainty(_value), USUMap and USSkewUmap (cu/utils/{uncertainty_projection,umap,skew_umap}.py,
    cu/data/camus/utils.py:101-153) on 3 synthetic images.  Maps are stored sparsely (indices + values)."""
    from contour_uncertainty.data.camus.utils import USUMap, USSkewUmap
    from contour_uncertainty.utils.uncertainty_projection import projected_uncertainty, projected_uncertainty_value

    n = 3
    mus, covs = synth.point_distributions(n, seed=101)
    al = synth.skew_alphas(n, seed=102)
    out = dict(mu=mus, cov=covs, alpha=al)
    U, V, U0, VAL, U2, V2, AP, MODE = [], [], [], [], [], [], [], []
    for i in range(n):
        u, v = projected_uncertainty(mus[i], covs[i], all=True)
        u0, _ = projected_uncertainty(mus[i], covs[i])
        u2, v2, ap = projected_uncertainty(mus[i], covs[i], al[i].copy(), all=True, linear_close=True)
        U.append(u); V.append(v); U0.append(u0); VAL.append(projected_uncertainty_value(mus[i], covs[i]))
        U2.append(u2); V2.append(v2); AP.append(ap)
        um = USUMap.__call__(mus[i], covs[i], labels=[])
        mode, sk = USSkewUmap.__call__(mus[i], covs[i], al[i], labels=[])
        MODE.append(mode)
        for name, m in (("umap", um), ("skew_umap", sk)):
            idx = np.flatnonzero(m)
            out[f"{name}_idx_{i}"] = idx.astype(np.int32)
            out[f"{name}_val_{i}"] = m.reshape(-1)[idx]
    out.update(unc_all=np.array(U), proj_all=np.array(V), unc=np.array(U0), value=np.array(VAL), skew_unc=np.array(U2),
               skew_proj=np.array(V2), alpha_proj=np.array(AP), mode=np.array(MODE))
    np.savez_compressed(os.path.join(HERE, "umap.npz"), **out)
    print("umap.npz", {k: np.asarray(v).shape for k, v in out.items() if not k[-1].isdigit()}, "value", VAL)


def lung_point_laws(n, seed):
    """(mu (n,120,2), cov (n,120,2,2), alpha (n,120,2)) fp32 for the lung u-maps."""
    rng = np.random.default_rng(seed)
    mu = synth.lung_landmarks(n, seed=seed).astype(np.float32)
    cov = np.array([synth.random_spd(rng, 120, 2.0, 20.0) for _ in range(n)]).astype(np.float32)
    alpha = rng.normal(0, 2.0, (n, 120, 2)).astype(np.float32)
    return mu, cov, alpha


def golden_lung_umap():
    """LungUMap / LungSkewUmap (contour_uncertainty/data/lung/utils.py:72-114) executed as they are: skew_umap per structure
    (44 / 50 / 26 points), each / max, summed, clipped.  Maps are stored sparsely (indices + values)."""
    from contour_uncertainty.data.lung.utils import LungSkewUmap, LungUMap

    n = 2
    mu, cov, alpha = lung_point_laws(n, seed=131)
    out = dict(mu=mu, cov=cov, alpha=alpha)
    modes = []
    for i in range(n):
        um = LungUMap.__call__(mu[i], cov[i])
        mode, sk = LungSkewUmap.__call__(mu[i], cov[i], alpha[i])
        modes.append(mode)
        for name, m in (("umap", um), ("skew_umap", sk)):
            idx = np.flatnonzero(m)
            out[f"{name}_idx_{i}"] = idx.astype(np.int32)
            out[f"{name}_val_{i}"] = m.reshape(-1)[idx]
    out["mode"] = np.array(modes)
    np.savez_compressed(os.path.join(HERE, "lung_umap.npz"), **out)
    print("lung_umap.npz", out["mode"].shape, [len(out[f"umap_idx_{i}"]) for i in range(n)])


def golden_seqskew():
    """SequenceSkewPSMSampler.__call__ (psm_skew_sequence.py) with first-instant bits, rvs_fast normals and the
    31 multinomial races per sample replayed; fp32 and fp64 runs of the unmodified code."""
    import random

    from contour_uncertainty.sampler.posterior_shape_model.psm_skew_sequence import SequenceSkewPSMSampler

    SEQ_PATH = os.path.join(_refstub.REFERENCE_ROOT, "camus-cont_sequence_psm_11_no_std.npy")
    Q_SEED = 113
    n_pair, n_s = 2, 3
    mus, covs = synth.point_distributions(2 * n_pair, seed=111)
    mus = mus.reshape(n_pair, 2, 21, 2).copy()
    ctr = mus[:, 1].mean(1, keepdims=True)
    mus[:, 1] = ctr + (mus[:, 0] - mus[:, 0].mean(1, keepdims=True)) * 0.8 + np.random.default_rng(5).normal(0, 1.0, (n_pair, 21, 2)).astype(np.float32)
    covs = covs.reshape(n_pair, 2, 21, 2, 2)
    alphas = synth.skew_alphas(2 * n_pair, seed=112).reshape(n_pair, 2, 21, 2)
    rng = np.random.default_rng(114)
    z = rng.standard_normal((n_pair, n_s, 3, 3)).astype(np.float32)
    bits = rng.integers(0, 2, (n_pair, n_s))
    bits[0, 0], bits[0, 1] = 0, 1
    real_multinomial = torch.multinomial
    outs, idxs = {}, {}
    for mode in ("f32", "f64"):
        if mode == "f64":
            torch.set_default_dtype(torch.float64)
        try:
            smp = SequenceSkewPSMSampler(PSM_PATH, SEQ_PATH)
            if mode == "f64":
                for a in ("mean", "scale", "X_train", "mu", "Q", "seq_mu", "seq_Q", "seq_mean", "seq_scale", "seq_X_train"):
                    setattr(smp, a, getattr(smp, a).double())
            res = np.zeros((n_pair, 2, n_s, 21, 2))
            index_log = []
            for i in range(n_pair):
                state = {"draw": 0}

                def fake_multinomial(p, n, *a, _i=i, _st=state, **k):
                    d = _st["draw"]
                    _st["draw"] += 1
                    stream = (_i * n_s + d // 31) * 31 + d % 31
                    try:
                        real_multinomial(p, n)
                    except Exception:
                        index_log.append(-1)
                        raise
                    q = torch.from_numpy(synth.exponential_field(Q_SEED, stream)).to(p.dtype)
                    idx = torch.argmax(p / q).reshape(1)
                    index_log.append(int(idx))
                    return idx

                dt = np.float64 if mode == "f64" else np.float32
                draws = [z[i, s, d][None].astype(dt) for s in range(n_s) for d in range(3)]
                it = iter(bits[i].tolist())
                orig = random.randint
                random.randint = lambda a, b, _it=it: int(next(_it))
                torch.multinomial = fake_multinomial
                try:
                    with _refstub.replay_standard_normal(draws) as log:
                        m, c, al = (torch.from_numpy(a[i]) for a in (mus, covs, alphas))
                        if mode == "f64":
                            m, c, al = m.double(), c.double(), al.double()
                        r = smp(m, c, al, n=n_s)
                finally:
                    torch.multinomial = real_multinomial
                    random.randint = orig
                assert len(log) == 3 * n_s and state["draw"] == 31 * n_s, (len(log), state["draw"])
                res[i] = r.detach().numpy()
            outs[mode] = res
            idxs[mode] = np.array(index_log).reshape(n_pair, n_s, 31)
        finally:
            torch.set_default_dtype(torch.float32)
    print("seqskew: fp32-vs-fp64 identical draws", (idxs["f32"] == idxs["f64"]).mean(), "fallbacks", (idxs["f32"] < 0).sum(),
          "max diff (px)", np.abs(outs["f32"] - outs["f64"]).max())
    np.savez_compressed(os.path.join(HERE, "seqskew.npz"), mu=mus, cov=covs, alpha=alphas, z=z, bits=bits, q_seed=Q_SEED,
                        samples=outs["f32"].astype(np.float32), samples_f64=outs["f64"], indices=idxs["f32"], indices_f64=idxs["f64"])
    print("seqskew.npz", outs["f32"].shape)


def calibration_inputs(seed=3):
    """(umap (3,256,256) fp64 in [0,1] with runs exactly on bin boundaries, pred, gt uint8)."""
    rng = np.random.default_rng(seed)
    um = rng.uniform(0, 1, (3, 256, 256))
    um[0, :50] = 0.05 * 3
    um[1, :20] = 0.0
    um[2, 100:120] = 1.0
    pred = (rng.uniform(size=um.shape) < 0.4).astype(np.uint8)
    gt = (rng.uniform(size=um.shape) < 0.4).astype(np.uint8)
    return um, pred, gt


def golden_calibration():
    """Calibration.get_confidences_and_accuracies + Calibration.ece (cu/results/calibration.py) on synthetic maps."""
    from contour_uncertainty.results.calibration import Calibration

    um, pred, gt = calibration_inputs()
    res = [SimpleNamespace(pred=pred[i], gt=gt[i], uncertainty_map=um[i]) for i in range(3)]
    bb = np.linspace(0, 1, 21)
    fake = SimpleNamespace(nb_bins=20, bin_lowers=bb[:-1], bin_uppers=bb[1:],
                           get_pixel_confidences_and_accuracies=Calibration.get_pixel_confidences_and_accuracies)
    c, a = Calibration.get_confidences_and_accuracies(fake, res, "uncertainty_map")
    ece, mce = Calibration.ece(fake, c, a, name=None)
    np.savez_compressed(os.path.join(HERE, "calibration.npz"), seed=3, ece=float(np.asarray(ece).reshape(-1)[0]), mce=float(mce),
                        n_fg=len(c))
    print("calibration.npz ece", ece, "mce", mce, "foreground pixels", len(c))


def golden_gls():
    """perimeter / global_longitudinal_strain (cu/utils/clinical.py:30-71) on contour samples (incl. a zero-chord one)."""
    from contour_uncertainty.utils.clinical import perimeter, global_longitudinal_strain

    rng = np.random.default_rng(121)
    b = synth.LV_EXAMPLE_SHAPE
    ed = (b + rng.normal(0, 2.5, (12, 21, 2))).astype(np.float32)
    es = (b.mean(0) + (b - b.mean(0)) * 0.8 + rng.normal(0, 2.5, (12, 21, 2))).astype(np.float32)
    ed[3, 8] = ed[3, 7]                                     # zero chord: FITPACK raises -> raw points
    p_ed, p_es = perimeter(ed), perimeter(es)
    gls = np.array([global_longitudinal_strain(ed[i], es[i]) for i in range(12)])
    np.savez_compressed(os.path.join(HERE, "gls.npz"), ed=ed, es=es, perimeter_ed=p_ed, perimeter_es=p_es, gls=gls)
    print("gls.npz perimeter", p_ed[:4], "gls", gls[:4])


def golden_distributions():
    """BivariateNormal / BivariateSkewNormal classmethods (cu/distributions/*.py) on random points; rvs_fast replayed."""
    from contour_uncertainty.distributions.bivariatenormal import BivariateNormal
    from contour_uncertainty.distributions.bivariateskewnormal import BivariateSkewNormal

    rng = np.random.default_rng(131)
    mus, covs = synth.point_distributions(1, seed=132)
    mu, cov = torch.from_numpy(mus[0, 3]), torch.from_numpy(covs[0, 3])
    alpha = torch.tensor([1.7, -2.3])
    x = torch.from_numpy((mus[0, 3] + rng.normal(0, 6.0, (200, 2))).astype(np.float32))
    out = dict(mu=mu.numpy(), cov=cov.numpy(), alpha=alpha.numpy(), x=x.numpy(),
               n_logpdf=BivariateNormal.logpdf(x, mu, cov).numpy(), n_pdf=BivariateNormal.pdf(x, mu, cov).numpy(),
               s_logpdf=BivariateSkewNormal.logpdf(x, mu, cov, alpha).numpy(), s_pdf=BivariateSkewNormal.pdf(x, mu, cov, alpha).numpy())
    # paired laws (one per point)
    mu_n = torch.from_numpy(mus[0, :20]); cov_n = torch.from_numpy(covs[0, :20]); al_n = torch.from_numpy(synth.skew_alphas(1, 133)[0, :20])
    x_n = mu_n + torch.from_numpy(rng.normal(0, 4.0, (20, 2)).astype(np.float32))
    out.update(mu_n=mu_n.numpy(), cov_n=cov_n.numpy(), alpha_n=al_n.numpy(), x_n=x_n.numpy(),
               s_logpdf_n=BivariateSkewNormal.logpdf(x_n, mu_n, cov_n, al_n).numpy())
    z = rng.standard_normal((16, 3)).astype(np.float32)
    with _refstub.replay_standard_normal([z]) as log:
        r = BivariateSkewNormal.rvs_fast(mu, cov, alpha, size=(16,))
    assert log == [(16, 3)], log
    m0, v0, a0 = BivariateSkewNormal.marginal(mu, cov, alpha.numpy(), axis=0, angle=torch.tensor(0.4))
    out.update(z=z, rvs=r.numpy(), marginal=np.array([float(m0), float(v0), float(a0)]))
    # BivariateNormal.rvs (replayed draws), BivariateSkewNormal.mode / BivariateNormal.mode
    z2 = rng.standard_normal((16, 2)).astype(np.float32)
    with _refstub.replay_standard_normal([z2]) as log:
        r2 = BivariateNormal.rvs(mu, cov, size=(16,))
    assert log == [(16, 2)], log
    al_m = torch.from_numpy(synth.skew_alphas(1, 134)[0, :8])
    modes = np.stack([BivariateSkewNormal.mode(mu_n[k], cov_n[k], al_m[k]).numpy() for k in range(8)])
    out.update(z2=z2, n_rvs=r2.numpy(), alpha_mode=al_m.numpy(), s_mode=modes)
    np.savez_compressed(os.path.join(HERE, "distributions.npz"), **out)
    print("distributions.npz", {k: v.shape for k, v in out.items()})


def golden_naive():
    """NaiveSampler.__call__ under np.random.seed; the draws are re-derived from the same seed."""
    from contour_uncertainty.sampler.naive import NaiveSampler

    mus, covs = synth.point_distributions(2, seed=41)
    n_s = 16
    outs, zs = [], []
    for i in range(2):
        np.random.seed(410 + i)
        outs.append(NaiveSampler()(mus[i], covs[i], n_s))
        rs = np.random.RandomState(410 + i)
        zs.append(np.stack([rs.standard_normal((n_s, 2)) for _ in range(21)], 1))  # (n, K, 2)
    np.savez_compressed(os.path.join(HERE, "naive.npz"), mu=mus, cov=covs, z=np.stack(zs),
                        samples=np.stack(outs))
    print("naive.npz", np.stack(outs).shape, np.stack(outs).dtype)


def golden_maps():
    """samples -> masks -> mean / pred / sample_entropy / lv_area, as _predict_step + predict_step do."""
    from contour_uncertainty.data.camus.utils import USContourToMask
    from contour_uncertainty.task.uncertainty import UncertaintyTask
    from contour_uncertainty.utils.clinical import lv_area

    g = np.load(os.path.join(HERE, "psm.npz"))
    samples = g["samples"]  # (4, 12, 21, 2)
    n, T = samples.shape[:2]
    pred_samples = np.array([USContourToMask.__call__(samples[i, k], (256, 256), labels=[], apply_argmax=False)
                             for i in range(n) for k in range(T)]).reshape(n, 1, T, 1, 256, 256)
    pred = pred_samples.mean(axis=(1, 2)).squeeze().round().astype(int)  # aleatoric.py:111
    ent = np.array([UncertaintyTask.sample_entropy(pred_samples[i].astype(float)) for i in range(n)])
    area = np.array([[lv_area(pred_samples[i, 0, k, 0]) for k in range(T)] for i in range(n)])
    np.savez_compressed(os.path.join(HERE, "maps.npz"), count=pred_samples.sum((1, 2, 3)).astype(np.int32),
                        pred=pred.astype(np.uint8), entropy=ent, area=area.astype(np.int64),
                        mean=pred_samples.astype(float).mean(axis=(1, 2)).squeeze())
    print("maps.npz", ent.shape, ent.dtype, "entropy max", ent.max(), "areas", area[0, :4])


def golden_segpost():
    """SegmentationUncertaintyTask.predict_step executed verbatim on synthetic samples (a subclass supplies
    ``predict`` = the recorded samples; everything after it is the reference's own code)."""
    from contour_uncertainty.task.segmentation.segmentation_uncertainty import SegmentationUncertaintyTask
    from vital.data.config import Tags

    n, Te, Ta = 3, 2, 8
    samples = synth.segmentation_samples(n, Te * Ta, seed=41).reshape(n, Te, Ta, 1, 256, 256)

    class Recorded(SegmentationUncertaintyTask):
        def __init__(self):
            torch.nn.Module.__init__(self)

        def predict(self, img):
            s = torch.from_numpy(samples)
            return s.mean((1, 2)), np.zeros((n, 256, 256)), s

        @staticmethod
        def get_voxelspacing_and_instants(batch):
            return None, None

    task = Recorded()
    task.hparams = SimpleNamespace(data_params=SimpleNamespace(labels=["bg", "lv"]))
    batch = {Tags.img: torch.zeros((n, 1, 256, 256)), Tags.gt: torch.zeros((n, 256, 256)), Tags.id: [str(i) for i in range(n)]}
    res = task.predict_step(batch, 0)
    np.savez_compressed(os.path.join(HERE, "segpost.npz"), seed=41, shape=np.array([n, Te, Ta]),
                        entropy=res.entropy_map.astype(np.float32),
                        pred_samples=np.packbits(res.pred_samples.reshape(n, Te * Ta, 256, 256), axis=-1, bitorder="little"),
                        pred=res.pred.astype(np.uint8), entropy_mean=res.instant_uncertainty["entropy_mean"])
    print("segpost.npz", res.entropy_map.shape, res.entropy_map.dtype, "entropy max", res.entropy_map.max(),
          "pred_samples px", res.pred_samples.sum((-2, -1)).reshape(n, -1)[0])


def glsmask_extra_masks():
    """Masks that stress the polar peak search: disc, square, small blob, thin bar, empty, two blobs."""
    yy, xx = np.mgrid[0:256, 0:256]
    out = []
    out.append(((yy - 128) ** 2 + (xx - 128) ** 2 <= 50 ** 2))
    m = np.zeros((256, 256), bool); m[60:200, 80:180] = True; out.append(m)
    m = np.zeros((256, 256), bool); m[100:104, 100:103] = True; out.append(m)
    m = np.zeros((256, 256), bool); m[30:230, 126:130] = True; out.append(m)
    out.append(np.zeros((256, 256), bool))
    m = np.zeros((256, 256), bool); m[20:60, 20:60] = True; m[120:220, 100:180] = True; out.append(m)
    m = (((yy - 100) / 70.0) ** 2 + ((xx - 128) / 40.0) ** 2 <= 1) & (yy < 150); out.append(m)      # half ellipse, flat base
    return np.array(out)


def golden_glsmask():
    """EchoMeasure.gls (vital/vital/utils/image/us/measure.py:320-366, myo_labels=None) and its pieces on rasterised
    contour samples.  skimage.measure.find_contours is the restated one (oracle/port.py): parity unpinned for it."""
    from contour_uncertainty.data.camus.utils import USContourToMask
    from vital.data.camus.config import Label
    from vital.utils.image.us.measure import EchoMeasure
    import functools

    T = 10
    contours = _simpson_contours(T)
    masks = np.array([[USContourToMask.__call__(contours[v, t], (256, 256), labels=[], apply_argmax=True) for t in range(T)]
                      for v in range(4)]).astype(bool).reshape(4 * T, 256, 256)
    masks = np.concatenate([masks, glsmask_extra_masks()])
    base_fn = functools.partial(EchoMeasure._extract_landmarks_from_polar_contour, labels=Label.LV, polar_smoothing_factor=5e-3, apex=False)
    n = len(masks)
    corners = np.full((n, 2, 2), np.nan)
    npts = np.zeros(n, dtype=np.int64)
    length = np.full(n, np.nan)
    for i, m in enumerate(masks):
        try:
            corners[i] = np.array(base_fn(m.astype(int)))
            c = EchoMeasure._endo_epi_contour(m.astype(int), Label.LV, base_fn)
            npts[i] = len(c)
            length[i] = sum(np.linalg.norm((p1 - p0) * np.array((1, 1))) for p0, p1 in zip(c[:-1], c[1:]))   # measure.py:354
        except Exception as e:
            print("  mask", i, "raised", type(e).__name__, e)
    gls = np.full((2, T), np.nan)
    for v in range(2):
        for t in range(T):
            try:
                gls[v, t] = EchoMeasure.gls(np.stack([masks[(2 * v) * T + t], masks[(2 * v + 1) * T + t]]).astype(int), Label.LV, None)[1]
            except Exception as e:
                print("  gls", v, t, "raised", type(e).__name__)
    np.savez_compressed(os.path.join(HERE, "glsmask.npz"), masks=np.packbits(masks, axis=-1, bitorder="little"),
                        corners=corners, npts=npts, length=length, gls=gls)
    print("glsmask.npz", n, "masks; lengths", length[:4], "gls", gls[0, :4], "failed", int(np.isnan(length).sum()))


def glsmyo_extra_maps():
    """Hand-made label maps for the corner cases of _endo_base: a one-column frontier (the vertical branch, :209-214) on the
    left and on the right of the LV, a frontier in one place only, no myocardium at all (RuntimeError), one frontier pixel."""
    out = np.zeros((5, 256, 256), dtype=np.uint8)
    yy, xx = np.mgrid[:256, :256]
    lv = ((yy - 120) / 60.0) ** 2 + ((xx - 128) / 35.0) ** 2 <= 1
    out[0, 60:180, 94:160] = 1                     # LV block; MYO bar on its left edge: frontier = LV column 94 only
    out[0, 100:141, 90:94] = 2
    out[1][lv] = 1
    out[1][70:75, 126:131] = np.where(out[1][70:75, 126:131] == 1, 1, 2)         # myocardium touching the LV in one place only
    out[1][58:61, 128] = 2
    out[2][lv] = 1                                                               # no myocardium
    out[3, 60:180, 94:160] = 1                     # MYO bar on the right edge, off-centre: frontier = LV column 159 only
    out[3, 70:90, 160:163] = 2
    out[4, 60:180, 94:160] = 1                     # a myocardium pixel diagonal to the LV corner: one frontier pixel
    out[4, 59, 93] = 2
    return out


def golden_glsmyo():
    """EchoMeasure.gls with myo_labels (vital/vital/utils/image/us/measure.py:177-222 _endo_base, :153-175 _endo_epi_contour,
    :320-366 gls) on LV + MYO label maps, executed as it is.  skimage.measure.find_contours is the restated one."""
    import functools
    from vital.data.camus.config import Label
    from vital.utils.image.us.measure import EchoMeasure

    maps = np.concatenate([np.load(os.path.join(HERE, "myo.npz"))["argmax"], np.load(os.path.join(HERE, "myo_task.npz"))["pred_samples_argmax"].reshape(-1, 256, 256),
                           glsmyo_extra_maps()]).astype(np.uint8)
    base_fn = functools.partial(EchoMeasure._endo_base, lv_labels=Label.LV, myo_labels=Label.MYO)
    n = len(maps)
    corners = np.full((n, 2, 2), -1, dtype=np.int64)
    npts = np.zeros(n, dtype=np.int64)
    length = np.full(n, np.nan)
    for i, m in enumerate(maps):
        try:
            corners[i] = np.array(base_fn(m.astype(int)))
            c = EchoMeasure._endo_epi_contour(m.astype(int), Label.LV, base_fn)
            npts[i] = len(c)
            length[i] = sum(np.linalg.norm((p1 - p0) * np.array((0.7, 0.4))) for p0, p1 in zip(c[:-1], c[1:]))   # measure.py:354
        except Exception as e:
            print("  map", i, "raised", type(e).__name__, e)
    pairs = [(0, 1), (2, 3), (4, 5), (6, 7), (8, 9), (11, 13)]
    gls = np.array([EchoMeasure.gls(np.stack([maps[a], maps[b]]).astype(int), Label.LV, Label.MYO, voxelspacing=(0.7, 0.4))[1] for a, b in pairs])
    np.savez_compressed(os.path.join(HERE, "glsmyo.npz"), maps=maps, corners=corners, npts=npts, length=length, pairs=np.array(pairs), gls=gls)
    print("glsmyo.npz", n, "maps; lengths", length[:4], "gls", gls[:4], "failed", int(np.isnan(length).sum()))


def golden_generic():
    """reconstruction / linear_reconstruction for contours that are not 21 points long, and LungContourToMask
    (contour_uncertainty/data/lung/utils.py:25-69), executed as they are."""
    from contour_uncertainty.data.lung.utils import LungContourToMask
    from contour_uncertainty.utils.contour import linear_reconstruction, reconstruction

    out = {}
    for k in (4, 5, 8, 26, 44, 50, 64):
        c = np.concatenate([synth.closed_contours(6, k, seed=300 + k), synth.closed_contours(6, k, seed=400 + k, wild=True)])
        c[1] = np.round(c[1]) + 0.5                                  # rounding ties
        c[2, 3] = c[2, 2]                                            # zero chord -> FITPACK raises -> raw points
        spl = np.array([reconstruction(x, 256, 256) for x in c]).astype(np.uint8)
        lin = np.array([linear_reconstruction(x, 256, 256) for x in c]).astype(np.uint8)
        out[f"contours_{k}"] = c
        out[f"spline_{k}"] = np.packbits(spl, axis=-1, bitorder="little")
        out[f"linear_{k}"] = np.packbits(lin, axis=-1, bitorder="little")
    lm = synth.lung_landmarks(4, seed=77)
    out["lung_landmarks"] = lm
    out["lung_argmax"] = np.array([LungContourToMask.__call__(x) for x in lm]).astype(np.uint8)
    out["lung_stack_linear"] = np.array([LungContourToMask.__call__(x, apply_argmax=False, reconstruction_type="linear") for x in lm]).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "generic.npz"), **out)
    print("generic.npz", {k: v.shape for k, v in out.items() if k.startswith("spline")}, "lung areas",
          [(out["lung_argmax"] == l).sum() for l in (1, 2)])


def golden_dataset_calibration():
    """Dataset-level UCE / A-UCE / correlation / MAE as ClinicalMetrics.plot_calibration / plot_correlation form them
    (contour_uncertainty/results/clinical/clinical_metrics.py:156-226, results/utils/calibration.py:6-65)."""
    import json

    from contour_uncertainty.results.utils.calibration import compute_adaptive_calibration, compute_calibration
    from scipy.stats import pearsonr
    from sklearn.metrics import mean_absolute_error

    std, err, reject = synth.dataset_calibration_inputs()
    nas = np.isnan(std) | np.isnan(err)
    filt = (~reject)[~nas]
    g, p = np.where(reject, np.inf, err), std
    ok = ~(np.isnan(p) | np.isnan(g) | np.isinf(p) | np.isinf(g))
    want = {"uce": compute_calibration(err[20:], std[20:])[0], "a_uce": compute_adaptive_calibration(err[20:], std[20:], nb_bins=7)[0],
            "m_uce": compute_calibration(err[~nas], std[~nas], filters=filt)[0],
            "m_a_uce": compute_adaptive_calibration(err[~nas], std[~nas], filters=filt)[0],
            "correlation": float(pearsonr(g[ok], p[ok])[0]), "mae": float(mean_absolute_error(g[ok], p[ok]))}
    json.dump(want, open(os.path.join(HERE, "dataset_calibration.json"), "w"), indent=1)
    print("dataset_calibration.json", want)


def golden_myo():
    """USContourToMask with Label.MYO among the labels (cu/data/camus/utils.py:47-82), executed as it is;
    skimage.draw.polygon is the restated one (oracle/port.py): parity unpinned for it."""
    from contour_uncertainty.data.camus.utils import USContourToMask
    from vital.data.camus.config import Label

    lm = np.concatenate([synth.endo_epi_landmarks(10, seed=61), synth.endo_epi_landmarks(8, seed=62, wild=True)])
    labels = [Label.BG, Label.LV, Label.MYO]
    argmax = np.array([USContourToMask.__call__(x.astype(np.float32), (256, 256), labels=labels) for x in lm]).astype(np.uint8)
    stack_lin = np.array([USContourToMask.__call__(x.astype(np.float32), (256, 256), labels=labels, apply_argmax=False,
                                                   reconstruction_type="linear") for x in lm]).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "myo.npz"), landmarks=lm.astype(np.float32), argmax=argmax,
                        stack_linear=np.packbits(stack_lin, axis=-1, bitorder="little"))
    print("myo.npz", argmax.shape, "LV px", (argmax == 1).sum((1, 2))[:5], "MYO px", (argmax == 2).sum((1, 2))[:5])


# ------------------------------------------------------------------------------------------------
# The contour TASK end to end: the reference's own _predict_step + predict_step, executed verbatim
# ------------------------------------------------------------------------------------------------
TASK_N, TASK_TE, TASK_TA = 2, 2, 6          # N = 2 dodges the reference's .squeeze() bug at N = 1 (aleatoric.py:111)
TASK_SEED = 700


def _task_logits():
    """(T_e, N, 21, 256, 256) synthetic logits: pass e = the 'network' e of a two-member ensemble on the same 2 images."""
    base = synth.heatmap_logits(TASK_N, seed=TASK_SEED)
    other = synth.heatmap_logits(TASK_N, seed=TASK_SEED + 1)
    return np.stack([base, 0.7 * base + 0.3 * other]).astype(np.float32)


def _bare_task(cls, **hparams):
    """An instance of a reference task class without its Lightning / Hydra constructor (no checkpoint, no config tree):
    the attributes the predict path reads are set by hand, the functors are fetched by the reference's own
    on_predict_start() from a datamodule-like object (contour_uncertainty.py:59-63)."""
    from contour_uncertainty.data.camus.utils import USContourToMask, USSkewUmap, USUMap
    from vital.data.camus.config import Label

    task = cls.__new__(cls)
    torch.nn.Module.__init__(task)
    task.hparams = SimpleNamespace(data_params=SimpleNamespace(labels=[Label.BG, Label.LV], out_shape=(21, 2)), **hparams)
    task.trainer = SimpleNamespace(datamodule=SimpleNamespace(umap_fn=USUMap(), contour_to_mask_fn=USContourToMask(),
                                                              skew_umap_fn=USSkewUmap()))
    task.device = "cpu"
    task.on_predict_start()
    return task


def _task_batch():
    from contour_uncertainty.data.config import ContourTags
    from vital.data.config import Tags

    return {Tags.img: torch.zeros((TASK_N, 1, 256, 256)), Tags.gt: torch.zeros((TASK_N, 256, 256)),
            ContourTags.contour: torch.zeros((TASK_N, 21, 2)), Tags.id: [f"patient{i:04d}" for i in range(TASK_N)]}


def _store_result(out, tag, res):
    n = TASK_N
    for k in ("mu", "mode", "cov", "post_mu", "post_cov", "contour_samples", "pred", "uncertainty_map", "entropy_map", "alpha"):
        v = getattr(res, k)
        if v is not None:
            out[f"{tag}_{k}"] = np.asarray(v)
    out[f"{tag}_pred_samples"] = np.packbits(np.asarray(res.pred_samples).reshape(n, -1, 256, 256).astype(np.uint8), axis=-1, bitorder="little")
    out[f"{tag}_pred_samples_shape"] = np.array(res.pred_samples.shape)
    for k, v in res.point_uncertainty.items():
        out[f"{tag}_point_{k}"] = np.asarray(v)
    for k, v in res.instant_uncertainty.items():
        out[f"{tag}_instant_{k}"] = np.asarray(v)


def golden_task_al():
    """dsnt-al: DSNTAleatoric.predict -> AleatoricUncertaintyTask._predict_step -> ContourUncertaintyTask.predict_step
    (dsnt_al.py:118-151, aleatoric.py:54-135, contour_uncertainty.py:71-130), the reference's code from the network output
    to the BatchResult, on synthetic logits through a two-member stub 'ensemble' with replayed PSM draws.

    Two runs: 'f32' = everything as the reference runs it; 'f64' = the same code with the PSM sampler's members and inputs
    cast to double (its fp32 42x42 inverses at sigma2 = 1e-3 carry ~3e-2 px of noise, see golden_psm)."""
    from contour_uncertainty.sampler.posterior_shape_model.psm import PosteriorShapeModelSampler
    from contour_uncertainty.task.regression.dsnt.dsnt_al import DSNTAleatoric

    logits = _task_logits()
    z = np.random.default_rng(TASK_SEED + 2).standard_normal((TASK_N, TASK_TE, TASK_TA, 17, 2)).astype(np.float32)
    out = {"z": z, "shape": np.array([TASK_N, TASK_TE, TASK_TA]), "seed": TASK_SEED}

    class DoubleSampler(PosteriorShapeModelSampler):
        def __call__(self, mu, cov, *a, **k):
            return super().__call__(mu.double(), cov.double(), *a, **k)

    for mode in ("f32", "f64"):
        task = _bare_task(DSNTAleatoric, covar=True, t_e=TASK_TE, t_a=TASK_TA, sequence_sampler=False)
        task.ensembling = True
        task.model = [lambda img, _e=e: torch.from_numpy(logits[_e]) for e in range(TASK_TE)]
        if mode == "f64":
            torch.set_default_dtype(torch.float64)
        try:
            task.sampler = (DoubleSampler if mode == "f64" else PosteriorShapeModelSampler)(PSM_PATH)
            if mode == "f64":
                for a in ("mean", "scale", "X_train", "X_val", "mu", "Q"):
                    setattr(task.sampler, a, getattr(task.sampler, a).double())
            # AleatoricUncertaintyTask.sample: images outer, passes inner, t_a samples of 17 draws each (aleatoric.py:71-73)
            dt = np.float64 if mode == "f64" else np.float32
            draws = [z[i, e, s, d][None].astype(dt) for i in range(TASK_N) for e in range(TASK_TE) for s in range(TASK_TA) for d in range(17)]
            with _refstub.replay_standard_normal(draws) as log:
                if mode == "f64":
                    torch.set_default_dtype(torch.float32)      # the DSNT head stays fp32; only the sampler runs in double
                    orig_call = DoubleSampler.__call__

                    def in_double(self, mu, cov, *a, **k):
                        torch.set_default_dtype(torch.float64)
                        try:
                            return orig_call(self, mu, cov, *a, **k).float()
                        finally:
                            torch.set_default_dtype(torch.float32)

                    DoubleSampler.__call__ = in_double
                with torch.no_grad():                           # Lightning's predict loop runs predict_step without autograd
                    mu_te, cov_te = task.predict(_task_batch()["img"])
                    res = task.predict_step(_task_batch(), 0)
            assert len(log) == TASK_N * TASK_TE * TASK_TA * 17, len(log)
        finally:
            torch.set_default_dtype(torch.float32)
        if mode == "f32":
            out["mu_te"], out["cov_te"] = mu_te.numpy(), cov_te.numpy()
        _store_result(out, mode, res)
    d = np.abs(out["f32_contour_samples"] - out["f64_contour_samples"]).max()
    print("task_al: fp32-vs-fp64 sampler max diff (px)", d, "pred_samples equal", np.array_equal(out["f32_pred_samples"], out["f64_pred_samples"]))
    np.savez_compressed(os.path.join(HERE, "task_al.npz"), **out)
    print("task_al.npz", {k: v.shape for k, v in out.items() if k.startswith("f64_") and v.ndim > 1})


def golden_task_skew():
    """dsnt-skew: DSNTSkew.predict -> SkewUncertaintyTask._predict_step -> ContourUncertaintyTask.predict_step
    (dsnt_skew.py:153-199, aleatoric_skew.py:49-127, contour_uncertainty.py:71-130) verbatim: the stub networks return
    (logits, raw skewness) and skew_block is the identity; rvs_fast normals replayed, torch.multinomial fed from
    synth.exponential_field (checked equal to the real exponential race in golden_skew).  25 samples per pass are
    hard-coded in the reference (aleatoric_skew.py:63)."""
    from contour_uncertainty.sampler.posterior_shape_model.psm_skew import SkewPosteriorShapeModelSampler
    from contour_uncertainty.task.regression.dsnt.dsnt_skew import DSNTSkew

    T = 25
    Q_SEED = TASK_SEED + 5
    logits = _task_logits()
    raw_alpha = np.stack([synth.skew_alphas(TASK_N, seed=TASK_SEED + 3), synth.skew_alphas(TASK_N, seed=TASK_SEED + 4)]).astype(np.float32)
    z = np.random.default_rng(TASK_SEED + 6).standard_normal((TASK_N, TASK_TE, T, 3, 3)).astype(np.float32)
    out = {"z": z, "raw_alpha": raw_alpha, "shape": np.array([TASK_N, TASK_TE, T]), "seed": TASK_SEED, "q_seed": Q_SEED}

    task = _bare_task(DSNTSkew, covar=True, t_e=TASK_TE, t_a=T, sequence_sampler=False)
    task.ensembling = True
    task.model = [lambda img, _e=e: (torch.from_numpy(logits[_e]), torch.from_numpy(raw_alpha[_e])) for e in range(TASK_TE)]
    task.skew_block = lambda features: features
    task.skew_indices = list(range(21))
    task.sampler = SkewPosteriorShapeModelSampler(PSM_PATH)
    # SkewUncertaintyTask.sample: passes outer, images inner (aleatoric_skew.py:50-51, psm_skew.py:203-205), per sample 3
    # (1,3) normals then 14 multinomial draws
    order = [(i, e) for e in range(TASK_TE) for i in range(TASK_N)]
    draws = [z[i, e, s, d][None] for (i, e) in order for s in range(T) for d in range(3)]
    real_multinomial = torch.multinomial
    state = {"draw": 0}
    index_log = np.full((TASK_N, TASK_TE, T, 14), -2, dtype=np.int64)

    def fake_multinomial(p, n, *a, **k):
        c = state["draw"]
        state["draw"] += 1
        i, e = order[c // (14 * T)]
        s, d = (c // 14) % T, c % 14
        stream = ((i * TASK_TE + e) * T + s) * 14 + d
        try:
            real_multinomial(p, n)
        except Exception:
            index_log[i, e, s, d] = -1
            raise
        q = torch.from_numpy(synth.exponential_field(Q_SEED, stream)).to(p.dtype)
        idx = torch.argmax(p / q).reshape(1)
        index_log[i, e, s, d] = int(idx)
        return idx

    torch.multinomial = fake_multinomial
    try:
        with _refstub.replay_standard_normal(draws) as log:
            with torch.no_grad():                               # Lightning's predict loop runs predict_step without autograd
                mu_te, cov_te, alpha_te = task.predict(_task_batch()["img"])
                res = task.predict_step(_task_batch(), 0)
    finally:
        torch.multinomial = real_multinomial
    assert len(log) == TASK_N * TASK_TE * T * 3 and state["draw"] == TASK_N * TASK_TE * T * 14, (len(log), state["draw"])
    out["mu_te"], out["cov_te"], out["alpha_te"] = mu_te.numpy(), cov_te.numpy(), alpha_te.numpy()
    out["indices"] = index_log
    _store_result(out, "f32", res)
    np.savez_compressed(os.path.join(HERE, "task_skew.npz"), **out)
    print("task_skew.npz", {k: v.shape for k, v in out.items() if k.startswith("f32_") and v.ndim > 1}, "fallbacks", int((index_log == -1).sum()))



def golden_metrics():
    """The clinical metric processors on stored BatchResults: AreaError.compute (cu/results/clinical/instant.py:31-57),
    FAC.compute (view.py:35-78) and Volume.compute (patient.py:75-225), executed as they are on a synthetic patient
    (views 2CH / 4CH, instants ED / ES, T_e = 2, T_a = 6, anisotropic voxel spacing), plus
    aleatoric_epistemic_uncertainty on arrays with NaNs.  The sample masks come from USContourToMask on stored contours, so
    the test rebuilds the same BatchResults.  find_contours (inside Volume) is the restated one: parity unpinned for it."""
    from contour_uncertainty.data.camus.utils import USContourToMask
    from contour_uncertainty.data.config import BatchResult
    from contour_uncertainty.results.clinical.instant import AreaError
    from contour_uncertainty.results.clinical.patient import Volume
    from contour_uncertainty.results.clinical.utils import aleatoric_epistemic_uncertainty
    from contour_uncertainty.results.clinical.view import FAC
    from vital.data.camus.config import Label

    te, ta = 2, 6
    rng = np.random.default_rng(171)
    b = synth.LV_EXAMPLE_SHAPE
    ctr = b.mean(0)
    scale = {"2CH": (1.0, 0.8), "4CH": (1.05, 0.84)}
    vs = {"2CH": (0.5, 0.35), "4CH": (0.31, 0.42)}
    out = {"shape": np.array([te, ta])}
    views = {}
    for name in ("2CH", "4CH"):
        mu = np.stack([ctr + (b - ctr) * s for s in scale[name]]).astype(np.float32)                       # (2, 21, 2): ED, ES
        gt_c = (mu + rng.normal(0, 1.5, mu.shape)).astype(np.float32)
        cs = (mu[:, None, None] + rng.normal(0, 2.5, (2, te, ta, 21, 2)) + rng.normal(0, 2.0, (2, te, ta, 1, 2))).astype(np.float32)
        # adversarial samples: ES larger than ED (FAC < 0 -> rejected), an empty ED mask (FAC = nan), a tiny ES blob
        cs[1, 0, 1] = ctr + (b - ctr) * 1.2
        cs[0, 1, 2] = b + np.array([-700.0, 0.0])
        cs[1, 1, 4] = ctr + (b - ctr) * 0.02
        mask = lambda c, **k: USContourToMask.__call__(c, (256, 256), labels=[Label.BG, Label.LV], **k)
        pred = np.array([mask(mu[i]) for i in range(2)])
        gt = np.array([mask(gt_c[i]) for i in range(2)])
        ps = np.array([[[mask(cs[i, e, a]) for a in range(ta)] for e in range(te)] for i in range(2)])
        views[name] = BatchResult(id=f"patient0001/{name}", img=np.zeros((2, 256, 256)), gt=gt, pred=pred, labels=[Label.BG, Label.LV],
                                  uncertainty_map=np.zeros((2, 256, 256)), instants={"ED": 0, "ES": 1}, voxelspacing=vs[name],
                                  mu=mu, mode=mu, cov=np.tile(np.eye(2), (2, 21, 1, 1)), contour=gt_c, contour_samples=cs, pred_samples=ps)
        out.update({f"{name}_mu": mu, f"{name}_gt_contour": gt_c, f"{name}_contour_samples": cs, f"{name}_voxelspacing": np.array(vs[name])})
    scalars = ("pred", "gt", "error", "std", "mean", "aleatoric_std", "epistemic_std")
    for name, v in views.items():
        for key, inst in v.instants.items():
            r = AreaError().compute(v, key, inst)
            out[f"area_{name}_{key}"] = np.array([r[k] for k in scalars], dtype=np.float64)
            out[f"area_{name}_{key}_mc"] = np.array(r["mc"], dtype=np.float64)
        r = FAC().compute(v)
        out[f"fac_{name}"] = np.array([r[k] for k in scalars], dtype=np.float64)
        out[f"fac_{name}_mc"] = np.array(r["mc"], dtype=np.float64)
        out[f"fac_{name}_reject"] = np.array(r["reject"])
        out[f"fac_{name}_sample_reject"] = np.array(r["sample_reject"])
    r = Volume().compute(views)
    for m in ("EF", "EDV", "ESV"):
        out[f"vol_{m}"] = np.array([r[f"{m}_{k}"] for k in scalars], dtype=np.float64)
        out[f"vol_{m}_mc"] = np.array(r[f"{m}_mc"], dtype=np.float64)
    out["vol_EF_reject"] = np.array(r["EF_reject"])
    out["vol_EF_sample_reject"] = np.array(r["EF_sample_reject"])
    # the split itself on arrays with NaN rows / columns
    mcs = rng.normal(50, 5, (5, 3, 8))
    mcs[0, 1, :] = np.nan
    mcs[1, :, 2:5] = np.nan
    mcs[2] = np.nan
    mcs[3, 0, 0] = np.nan
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out["ae_in"] = mcs
            out["ae_out"] = np.array([aleatoric_epistemic_uncertainty(m) for m in mcs], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print("metrics.npz", {k: v.tolist() for k, v in out.items() if k.startswith(("fac_2CH", "vol_EF")) and v.size < 8})



def golden_myo_task():
    """Labels with Label.MYO: ContourUncertaintyTask.convert_to_mask (cu/task/regression/contour_uncertainty.py:26-57) on
    42-landmark contour samples -> one-hot (N,T_e,T_a,3,H,W) stacks, then predict_step's per-image sample_entropy (base 3,
    uncertainty.py:107-133) and the argmax of :81-82, executed as they are.  skimage.draw.polygon is the restated one."""
    from contour_uncertainty.task.regression.contour_uncertainty import ContourUncertaintyTask
    from vital.data.camus.config import Label

    n, te, ta = 2, 2, 5
    task = _bare_task(ContourUncertaintyTask)
    task.hparams.data_params.labels = [Label.BG, Label.LV, Label.MYO]
    mu = synth.endo_epi_landmarks(n, seed=91).astype(np.float32)
    rng = np.random.default_rng(92)
    cs = (mu[:, None, None] + rng.normal(0, 2.0, (n, te, ta, 42, 2)) + rng.normal(0, 2.0, (n, te, ta, 1, 2))).astype(np.float32)
    # image_shape[1:] is what the reference reshapes the stack to (:52): it only works with a 3-channel "image" shape
    pred, ps = task.convert_to_mask(mu, (n, 3, 256, 256), cs)
    ent = np.array([task.sample_entropy(ps[i].astype(float)) for i in range(n)])
    np.savez_compressed(os.path.join(HERE, "myo_task.npz"), mu=mu, contour_samples=cs, pred=pred.astype(np.uint8),
                        pred_samples_argmax=ps.argmax(3).astype(np.uint8), entropy=ent,
                        counts=ps.sum((1, 2)).astype(np.int32))
    print("myo_task.npz", ps.shape, ent.dtype, float(ent.max()), "class px", [(pred == c).sum() for c in range(3)])


from class_entropy_inputs import class_entropy_inputs  # noqa: E402


def golden_class_entropy():
    """UncertaintyTask.sample_entropy (cu/task/uncertainty.py:107-133) on probabilities: C = 3 (base 3) and C = 1."""
    from contour_uncertainty.task.uncertainty import UncertaintyTask

    p3, p1 = class_entropy_inputs()
    e3 = np.array([UncertaintyTask.sample_entropy(x) for x in p3])
    e1 = np.array([UncertaintyTask.sample_entropy(x.reshape(3, 4, 1, 64, 64)) for x in p1])
    np.savez_compressed(os.path.join(HERE, "class_entropy.npz"), e3=e3, e1=e1)
    print("class_entropy.npz", e3.dtype, e3.shape, float(e3.max()), float(e1.max()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["dsnt", "raster", "psm", "naive", "maps", "seq", "skew", "simpson", "umap", "seqskew", "calibration", "gls", "distributions", "segpost", "glsmask", "generic", "dataset_calibration", "myo", "class_entropy", "task_al", "task_skew", "metrics", "myo_task", "psm_alpha", "glsmyo", "lung_umap"]
    for w in which:
        globals()["golden_" + w]()

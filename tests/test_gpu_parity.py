"""GPU parity tests: every CUDA kernel, called through the C ABI (ctypes layer), against the oracle
and the committed golden vectors.  Tolerances are stated at each assert."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import synth
from oracle import port

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def eng(dev):
    from contouring_uncertainty_b200 import UncertaintyEngine

    return UncertaintyEngine(device=dev)


def _unpack_np(a):
    return np.unpackbits(a, axis=-1, bitorder="little")


# ------------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("kind", ["bumps", "noise", "bimodal"])
def test_dsnt_vs_oracle_and_golden(dev, kind):
    from contouring_uncertainty_b200.engine import dsnt_moments

    g = np.load(os.path.join(G, "dsnt.npz"))
    n, seed = int(g[f"{kind}_n"]), int(g[f"{kind}_seed"])
    logits = synth.heatmap_logits(n, seed=seed, kind=kind)
    norm, mu, sig = dsnt_moments(torch.from_numpy(logits).to(dev))
    norm, mu, sig = norm.cpu().numpy().astype(np.float64), mu.cpu().numpy().astype(np.float64), sig.cpu().numpy().astype(np.float64)
    coords, var, covar = port.dsnt(port.flat_softmax(logits.astype(np.float64)))   # fp64 oracle
    # coordinates: 1e-5 relative to the [-1,1] grid span (2)  -> 2e-5 abs; measured ~1e-7
    assert np.abs(norm[..., :2] - coords).max() < 2e-6
    # variances / covariance: <= 1e-5 relative (covariance relative to sqrt(var_x var_y))
    assert np.max(np.abs(norm[..., 2:4] - var) / var) < 1e-5
    assert np.max(np.abs(norm[..., 4] - covar) / np.sqrt(var[..., 0] * var[..., 1])) < 1e-5
    mu_o, sig_o = port.dsnt_predict_on_batch(logits)
    assert np.abs(mu - mu_o).max() < 2.6e-4            # = 2e-6 * 128
    assert np.max(np.abs(sig - sig_o) / np.sqrt(sig_o[..., 0, 0] * sig_o[..., 1, 1])[..., None, None]) < 1e-5
    # and against the verbatim fp32 reference (its own error vs fp64 is ~1.5e-6 rel, 3e-6 abs coords)
    assert np.abs(norm[..., :2] - g[f"{kind}_coords"]).max() < 6e-6
    assert np.max(np.abs(norm[..., 2:4] - g[f"{kind}_var"]) / var) < 2e-5
    assert np.abs(mu - g[f"{kind}_pixel_mu"]).max() < 8e-4


def test_dsnt_heatmap_mode_matches_dsnt_of_softmax(dev):
    """dsnt(heatmaps) on already-normalised heatmaps (the reference's two-call form)."""
    from contouring_uncertainty_b200.engine import dsnt_moments

    logits = synth.heatmap_logits(1, seed=5)
    heat = port.flat_softmax(logits.astype(np.float64))
    coords, var, covar = port.dsnt(heat)
    norm, _, _ = dsnt_moments(torch.from_numpy(heat.astype(np.float32)).to(dev), is_logit=False)
    norm = norm.cpu().numpy().astype(np.float64)
    assert np.abs(norm[..., :2] - coords).max() < 2e-6
    assert np.max(np.abs(norm[..., 2:4] - var) / var) < 1e-5


def test_dsnt_no_covar_flag(dev):
    from contouring_uncertainty_b200.engine import dsnt_moments

    logits = torch.from_numpy(synth.heatmap_logits(1, seed=6)).to(dev)
    _, _, s1 = dsnt_moments(logits, covar=True)
    _, _, s0 = dsnt_moments(logits, covar=False)
    assert torch.all(s0[..., 0, 1] == 0) and torch.all(s0[..., 1, 0] == 0)
    assert torch.equal(s0[..., 0, 0], s1[..., 0, 0]) and torch.equal(s0[..., 1, 1], s1[..., 1, 1])
    assert torch.any(s1[..., 0, 1] != 0)


# ------------------------------------------------------------------------------------------- raster
def test_raster_golden_bit_exact(dev):
    from contouring_uncertainty_b200.engine import rasterize_contours, unpack_masks

    g = np.load(os.path.join(G, "raster.npz"))
    contours = torch.from_numpy(g["contours"]).to(dev)
    masks, area = rasterize_contours(contours, round_landmarks=True)
    got = unpack_masks(masks).cpu().numpy()
    want = _unpack_np(g["masks"])
    bad = [(i, str(g["tags"][i]), int((got[i] != want[i]).sum())) for i in range(len(want)) if not np.array_equal(got[i], want[i])]
    assert not bad, bad
    assert np.array_equal(area.cpu().numpy(), want.reshape(len(want), -1).sum(1))
    # reconstruction() on the unrounded float contour
    masks, _ = rasterize_contours(contours[:32], round_landmarks=False)
    got = unpack_masks(masks).cpu().numpy()
    want = _unpack_np(g["masks_unrounded"])
    bad = [(i, int((got[i] != want[i]).sum())) for i in range(32) if not np.array_equal(got[i], want[i])]
    assert not bad, bad
    # reconstruction_type='linear': polygon outline + hole fill (incl. the off-frame / duplicate-point cases)
    masks, area = rasterize_contours(contours, reconstruction_type="linear")
    got = unpack_masks(masks).cpu().numpy()
    want = _unpack_np(g["masks_linear"])
    bad = [(i, str(g["tags"][i]), int((got[i] != want[i]).sum())) for i in range(len(want)) if not np.array_equal(got[i], want[i])]
    assert not bad, bad
    assert np.array_equal(area.cpu().numpy(), want.reshape(len(want), -1).sum(1))
    from contouring_uncertainty_b200.data.camus.utils import USContourToMask

    one = USContourToMask.__call__(g["contours"][3], (256, 256), labels=[], reconstruction_type="linear")
    assert np.array_equal(one, want[3])


def test_raster_random_vs_oracle_bit_exact(dev):
    """1500 random contours (sigma 2..10 px, incl. self-intersections) against the oracle."""
    from contouring_uncertainty_b200.engine import rasterize_contours, unpack_masks

    rng = np.random.default_rng(77)
    n = 1500
    sig = rng.uniform(2, 10, (n, 1, 1))
    c = (synth.LV_EXAMPLE_SHAPE[None] + rng.normal(0, 1, (n, 21, 2)) * sig + rng.normal(0, 20, (n, 1, 2))).astype(np.float32)
    masks, area = rasterize_contours(torch.from_numpy(c).to(dev))
    got = unpack_masks(masks).cpu().numpy()
    nbad = 0
    for i in range(n):
        want = port.us_contour_to_mask(c[i])
        nbad += int(not np.array_equal(got[i], want))
    assert nbad == 0, f"{nbad} of {n} masks differ"


# ------------------------------------------------------------------------------------------- samplers
def test_psm_samples_vs_golden(dev, eng):
    from contouring_uncertainty_b200.engine import psm_factor, sample_raster_aggregate

    g = np.load(os.path.join(G, "psm.npz"))
    mu, cov = torch.from_numpy(g["mu"]).to(dev), torch.from_numpy(g["cov"]).to(dev)
    z = torch.from_numpy(g["z"]).to(dev).reshape(g["z"].shape[0], g["z"].shape[1], 34)
    a_t = psm_factor(eng.model, mu, cov)
    r = sample_raster_aggregate(mu, a_t, z.shape[1], z=z)
    s = r["contours"].cpu().numpy().astype(np.float64)
    # tight: the reference's own code executed in fp64; 1e-5 relative to the coordinate scale (256 px)
    err = np.abs(s - g["samples_f64"]).max()
    assert err < 2.56e-3 * 0.1, err          # measured ~3e-5 px (fp32 storage of A, fp32 matvec)
    # loose: verbatim fp32 reference (its own noise up to ~3.4e-2 px on the filled points)
    assert np.abs(s - g["samples"]).max() < 6e-2


def test_psm_direct_single_samples_vs_golden_and_factor_path(dev, eng):
    """cub200_psm_sample_direct (one sample per law, no factor: the sequence sampler's second instant) on the golden's laws,
    every law repeated once per stored draw vector: against the reference executed in fp64, and against the factor path."""
    from contouring_uncertainty_b200.engine import psm_factor, psm_sample_direct, sample_contours

    g = np.load(os.path.join(G, "psm.npz"))
    n, T = g["z"].shape[:2]
    mu = torch.from_numpy(np.repeat(g["mu"], T, axis=0)).to(dev)                 # (n T, 21, 2): item = (law, draw)
    cov = torch.from_numpy(np.repeat(g["cov"], T, axis=0)).to(dev)
    z = torch.from_numpy(g["z"]).to(dev).reshape(n * T, 34)
    s = psm_sample_direct(eng.model, mu, cov, z=z).cpu().numpy().astype(np.float64).reshape(n, T, 21, 2)
    err = np.abs(s - g["samples_f64"]).max()
    assert err < 2.56e-3 * 0.1, err                                               # the bound of the factor path's test: 1e-5 of the coordinate scale
    assert np.abs(s - g["samples"]).max() < 6e-2                                  # verbatim fp32 reference (its own noise)
    # generated draws: the same Philox stream as the factor path with one sample per law
    m = min(64, n * T)
    a_t = psm_factor(eng.model, mu[:m], cov[:m])
    via_factor = sample_contours(mu[:m], a_t, 1, seed=77).reshape(m, 21, 2)
    direct = psm_sample_direct(eng.model, mu[:m], cov[:m], seed=77)
    gap = (via_factor - direct).abs().max().item()
    assert gap < 5e-4, gap                                                        # the factor path stores A in fp32 and sums in fp32


def test_naive_sampler_given_factor_vs_golden(dev):
    from contouring_uncertainty_b200.engine import sample_raster_aggregate, naive_factor

    g = np.load(os.path.join(G, "naive.npz"))
    n, T = g["z"].shape[:2]
    M = np.stack([port.naive_colouring(g["cov"][i]) for i in range(n)])      # x = z @ M[k]
    a_t = np.zeros((n, 42, 42), dtype=np.float32)
    for k in range(21):
        a_t[:, 2 * k:2 * k + 2, 2 * k:2 * k + 2] = M[:, k]                     # a_t[e][d] = M[e][d]
    mu = torch.from_numpy(g["mu"]).to(dev)
    z = torch.from_numpy(g["z"].reshape(n, T, 42).astype(np.float32)).to(dev)
    r = sample_raster_aggregate(mu, torch.from_numpy(a_t).to(dev), T, z=z)
    assert np.abs(r["contours"].cpu().numpy() - g["samples"]).max() < 1e-4
    # device Cholesky colouring: same distribution, A A^T = cov
    a = naive_factor(torch.from_numpy(g["cov"]).to(dev)).cpu().numpy().astype(np.float64)
    for k in range(21):
        blk = a[:, 2 * k:2 * k + 2, 2 * k:2 * k + 2]                           # [e][d]
        np.testing.assert_allclose(np.einsum("ned,nef->ndf", blk, blk), g["cov"][:, k], rtol=1e-5, atol=1e-5)


def test_philox_draws_are_standard_normal_and_seeded(dev):
    from contouring_uncertainty_b200.engine import sample_raster_aggregate

    n, T = 4, 4096
    mu = torch.zeros((n, 21, 2), device=dev) + 128.0
    a_t = torch.zeros((n, 42, 42), device=dev)
    a_t[:, torch.arange(42), torch.arange(42)] = 1.0          # contour = mu + z
    r1 = sample_raster_aggregate(mu, a_t, T, seed=11)
    r2 = sample_raster_aggregate(mu, a_t, T, seed=11)
    r3 = sample_raster_aggregate(mu, a_t, T, seed=12)
    assert torch.equal(r1["contours"], r2["contours"]) and torch.equal(r1["area"], r2["area"])
    assert not torch.equal(r1["contours"], r3["contours"])
    zz = (r1["contours"] - 128.0).reshape(-1, 42).double()
    m = zz.numel()
    assert abs(zz.mean().item()) < 5 / np.sqrt(m)
    assert abs(zz.var().item() - 1) < 5 * np.sqrt(2 / m)
    assert abs((zz ** 4).mean().item() - 3) < 0.05
    corr = torch.corrcoef(zz.T)
    assert (corr - torch.eye(42, device=dev, dtype=torch.float64)).abs().max() < 0.03


# ------------------------------------------------------------------------------------------- fused + maps
def test_fused_counts_areas_maps_vs_oracle(dev, eng):
    """masks / count map / area / pred bit-exact, mean & entropy <= 1e-12, given the kernel's samples."""
    from contouring_uncertainty_b200.engine import unpack_masks

    mu, cov = synth.point_distributions(3, seed=91)
    T = 40
    res = eng.from_moments(torch.from_numpy(mu).to(dev), torch.from_numpy(cov).to(dev), T, seed=5, return_masks=True)
    contours = res.contour_samples.cpu().numpy()
    masks = unpack_masks(res.masks).cpu().numpy()
    for i in range(3):
        o = port.uncertainty_from_samples(contours[i])
        want = port.convert_to_mask(contours[i])
        assert np.array_equal(masks[i], want)
        assert np.array_equal(res.count[i].cpu().numpy(), o["count"])
        assert np.array_equal(res.area[i].cpu().numpy(), o["area"])
        assert np.array_equal(res.pred[i].cpu().numpy(), o["pred"])
        assert np.abs(res.mean[i].cpu().numpy() - o["mean"]).max() <= 1e-15
        assert np.abs(res.entropy[i].cpu().numpy() - o["entropy"]).max() <= 1e-12


def test_maps_golden(dev):
    """Reference contour samples (fp32, from psm.npz) -> our rasteriser -> maps == reference maps."""
    from contouring_uncertainty_b200.engine import rasterize_contours, finalize_maps

    g = np.load(os.path.join(G, "maps.npz"))
    p = np.load(os.path.join(G, "psm.npz"))
    n, T = p["samples"].shape[:2]
    diff = torch.zeros((n, 256, 256), dtype=torch.int32, device=dev)
    _, area = rasterize_contours(torch.from_numpy(p["samples"].reshape(n * T, 21, 2)).to(dev), return_masks=False,
                                 diff=diff, per_image=T)
    count, mean, ent, pred = finalize_maps(diff, T)
    assert np.array_equal(count.cpu().numpy(), g["count"])
    assert np.array_equal(area.cpu().numpy().reshape(n, T), g["area"])
    assert np.array_equal(pred.cpu().numpy(), g["pred"])
    assert np.abs(ent.cpu().numpy() - g["entropy"]).max() <= 1e-12
    assert np.abs(mean.cpu().numpy() - g["mean"]).max() <= 1e-15


def test_full_size_properties(dev, eng):
    """BASELINE config 2 size (64 images x 1000 PSM samples): size-independent invariants."""
    n, T = 64, 1000
    mu, cov = synth.point_distributions(n, seed=123)
    res = eng.from_moments(torch.from_numpy(mu).to(dev), torch.from_numpy(cov).to(dev), T, seed=99)
    count = res.count.long()
    assert int(count.min()) >= 0 and int(count.max()) <= T
    # checksum of checksums: total mask pixels == sum of per-sample areas, per image
    assert torch.equal(count.sum((1, 2)), res.area.long().sum(1))
    # the difference map telescopes: row sums of diff are >= 0 and the prefix sum is the count
    assert torch.all(res.pred == (2 * count > T).to(torch.uint8))
    assert float(res.entropy.max()) <= 1.0 + 1e-12 and float(res.entropy.min()) >= 0.0
    assert torch.all((res.entropy == 0) == ((count == 0) | (count == T)))
    # the oracle at this size: all 1000 samples of the first and the last image rasterised by the CPU restatement give the
    # same count map pixel for pixel (and with it mean / entropy / pred, exact functions of the count) and the same areas
    cs = res.contour_samples.cpu().numpy()
    for i in (0, n - 1):
        want = port.convert_to_mask(cs[i])                                           # (T, 256, 256) uint8
        assert np.array_equal(want.sum(0, dtype=np.int64), count[i].cpu().numpy())
        assert np.array_equal(want.reshape(T, -1).sum(1, dtype=np.int64), res.area[i].cpu().numpy())
    # a sub-sample re-rasterised by the debug path gives the same areas (sample order / indexing)
    from contouring_uncertainty_b200.engine import rasterize_contours

    sel = res.contour_samples[::7, ::50].reshape(-1, 21, 2).contiguous()
    _, a2 = rasterize_contours(sel, return_masks=False)
    assert torch.equal(a2, res.area[::7, ::50].reshape(-1))
    # linearity of the aggregation: two half runs accumulated == one full run
    from contouring_uncertainty_b200.engine import psm_factor, sample_raster_aggregate

    m = torch.from_numpy(mu[:4]).to(dev)
    a_t = psm_factor(eng.model, m, torch.from_numpy(cov[:4]).to(dev))
    z = torch.randn((4, 64, 34), device=dev)
    full = sample_raster_aggregate(m, a_t, 64, z=z)["diff"]
    half = sample_raster_aggregate(m, a_t, 32, z=z[:, :32].contiguous())["diff"]
    half = sample_raster_aggregate(m, a_t, 32, z=z[:, 32:].contiguous(), diff=half)["diff"]
    assert torch.equal(full, half)


def test_fac_from_areas(dev):
    from contouring_uncertainty_b200.engine import fac_from_areas

    ed = torch.tensor([100, 100, 0, 50], dtype=torch.int32, device=dev)
    es = torch.tensor([60, 120, 0, 50], dtype=torch.int32, device=dev)
    f = fac_from_areas(ed, es).cpu().numpy()
    assert f[0] == (100 - 60) / 100 and np.isnan(f[1]) and np.isnan(f[2]) and f[3] == 0.0


def test_empty_and_bad_inputs(dev, eng):
    from contouring_uncertainty_b200.engine import dsnt_moments, rasterize_contours

    m, a = rasterize_contours(torch.zeros((0, 21, 2), device=dev))
    assert m.shape == (0, 2048) and a.shape == (0,)
    with pytest.raises(ValueError):
        dsnt_moments(torch.zeros((1, 21, 128, 128), device=dev))
    with pytest.raises(ValueError):
        dsnt_moments(torch.zeros((1, 21, 256, 256)))          # CPU tensor: no fallback
    with pytest.raises(ValueError):
        rasterize_contours(torch.zeros((3, 20, 2), device=dev))


def test_metric_sums_exact(dev):
    from contouring_uncertainty_b200 import engine as E

    a = torch.randint(0, 40000, (64, 1000), dtype=torch.int32, device=dev)
    s = E.metric_sums(a)
    assert s.tolist() == [int(a.sum()), int((a.long() * a.long()).sum())]
    assert E.metric_sums(a[:0].contiguous()).tolist() == [0, 0]


def test_calibration_bins_vs_oracle_and_reference(dev):
    """Dataset-level pixel calibration: bins are additive over images (what ranks all_reduce); ECE / MCE equal the
    executed reference's to 1e-12 (fp64 sums in a different order); boundary confidences belong to no bin."""
    from contouring_uncertainty_b200 import engine as E
    from test_oracle_golden import _calibration_inputs

    g = np.load(os.path.join(G, "calibration.npz"))
    um, pred, gt = _calibration_inputs(int(g["seed"]))
    t = lambda a: torch.from_numpy(a).to(dev)
    bins = E.calibration_bins(t(um), t(pred), t(gt), 20)
    ece, mce = E.ece_from_bins(bins)
    assert abs(ece - float(g["ece"])) < 1e-12 and abs(mce - float(g["mce"])) < 1e-12
    assert int(bins[-1, 0]) == int(g["n_fg"])
    conf, acc = port.pixel_confidences_and_accuracies(um, pred, gt)
    bounds = np.linspace(0, 1, 21)
    for b in range(20):
        inb = (conf > bounds[b]) & (conf < bounds[b + 1])
        assert int(bins[b, 0]) == int(inb.sum()) and int(bins[b, 2]) == int(acc[inb].sum())
        assert abs(float(bins[b, 1]) - conf[inb].sum()) < 1e-9
    assert int(bins[:-1, 0].sum()) < int(bins[-1, 0])          # the runs sitting exactly on boundaries are in no bin
    parts = sum(E.calibration_bins(t(um[i]), t(pred[i]), t(gt[i]), 20) for i in range(3))
    assert torch.equal(parts[:, [0, 2]], bins[:, [0, 2]]) and torch.allclose(parts[:, 1], bins[:, 1], rtol=1e-13, atol=0)


def test_contour_perimeter_and_gls_vs_reference(dev):
    """Spline arc length (perimeter) and per-sample GLS against the executed reference: 1e-10 relative (fp64 spline
    evaluated piecewise instead of through FITPACK's B-spline recurrence; different summation order)."""
    from contouring_uncertainty_b200 import engine as E

    g = np.load(os.path.join(G, "gls.npz"))
    ed, es = torch.from_numpy(g["ed"]).to(dev), torch.from_numpy(g["es"]).to(dev)
    got = E.contour_perimeter(ed).cpu().numpy()
    ok = np.arange(12) != 3          # contour 3 has a zero chord: the reference then sums its raw float32 points in fp32
    np.testing.assert_allclose(got[ok], g["perimeter_ed"][ok], rtol=1e-10)
    np.testing.assert_allclose(got[3], g["perimeter_ed"][3], rtol=2e-7)
    np.testing.assert_allclose(E.contour_perimeter(es).cpu().numpy(), g["perimeter_es"], rtol=1e-10)
    np.testing.assert_allclose(E.global_longitudinal_strain(ed, es).cpu().numpy()[ok], g["gls"][ok], rtol=1e-8, atol=1e-10)
    assert E.contour_perimeter(ed.reshape(3, 4, 21, 2)).shape == (3, 4)

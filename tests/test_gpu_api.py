"""GPU tests of the reference-shaped Python API (same names / argument meaning as the reference)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import synth
from oracle import port

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_us_contour_to_mask_functor():
    from contouring_uncertainty_b200.data.camus.utils import USContourToMask

    g = np.load(os.path.join(G, "raster.npz"))
    want = np.unpackbits(g["masks"], axis=-1, bitorder="little")
    m = USContourToMask.__call__(g["contours"][3], (256, 256), labels=[], apply_argmax=True)
    assert m.shape == (256, 256) and m.dtype == np.int64 and np.array_equal(m, want[3])
    m = USContourToMask.__call__(g["contours"][5][None], (256, 256), labels=[], apply_argmax=False)
    assert m.shape == (1, 256, 256) and np.array_equal(m[0], want[5])
    mm = USContourToMask.__call__(g["contours"][:16], (256, 256))
    assert mm.shape == (16, 256, 256) and np.array_equal(mm, want[:16])
    lin = np.unpackbits(g["masks_linear"], axis=-1, bitorder="little")
    assert np.array_equal(USContourToMask.__call__(g["contours"][0], reconstruction_type="linear"), lin[0])
    with pytest.raises(ValueError):
        USContourToMask.__call__(g["contours"][0], reconstruction_type="cubic")


def test_psm_sampler_call_signature_and_replay():
    from contouring_uncertainty_b200.sampler import PosteriorShapeModelSampler

    g = np.load(os.path.join(G, "psm.npz"))
    s = PosteriorShapeModelSampler()
    assert s.initial_points == [0, 10, 20] and s.points_order[0] == [5, 15]
    mu, cov = torch.from_numpy(g["mu"][1]), torch.from_numpy(g["cov"][1])       # CPU tensors in, like the reference
    out = s(mu, cov, n=g["z"].shape[1], z=g["z"][1])
    assert out.shape == (12, 21, 2) and out.dtype == torch.float32 and out.device.type == "cpu"
    assert np.abs(out.numpy() - g["samples_f64"][1]).max() < 2.6e-4
    out2 = s(torch.from_numpy(g["mu"]).cuda(), torch.from_numpy(g["cov"]).cuda(), n=5, seed=3)
    assert out2.shape == (4, 5, 21, 2) and out2.is_cuda
    # Philox samples follow the affine law: empirical mean ~ mu on the three free points
    big = s(mu, cov, n=20000, seed=1).double()
    assert (big.mean(0) - mu.double()).abs().max() < 0.15


def test_naive_sampler_matches_reference_with_numpy_colouring():
    from contouring_uncertainty_b200.sampler import NaiveSampler

    g = np.load(os.path.join(G, "naive.npz"))
    s = NaiveSampler()
    out = s(g["mu"][0], g["cov"][0], g["z"].shape[1], z=g["z"][0], colouring=port.naive_colouring(g["cov"][0]))
    assert out.shape == (16, 21, 2) and out.dtype == np.float64
    assert np.abs(out - g["samples"][0]).max() < 1e-4
    big = s(g["mu"][0], g["cov"][0], 20000, seed=2)
    emp = np.stack([np.cov(big[:, k].T) for k in range(21)])
    assert np.abs(emp - g["cov"][0]).max() / np.abs(g["cov"][0]).max() < 0.06


def test_sample_entropy_and_pipeline_batch_result():
    from contouring_uncertainty_b200.task import ContourUncertaintyPipeline, sample_entropy

    g = np.load(os.path.join(G, "maps.npz"))
    p = np.load(os.path.join(G, "psm.npz"))
    masks = port.convert_to_mask(p["samples"][0])                                # (T,H,W)
    e = sample_entropy(masks[:, None].astype(float))
    assert np.abs(e - g["entropy"][0]).max() <= 1e-12
    e5 = sample_entropy(masks.reshape(3, 4, 1, 256, 256).astype(float))
    assert np.abs(e5 - g["entropy"][0]).max() <= 1e-12

    pipe = ContourUncertaintyPipeline(t_a=20)
    logits = torch.from_numpy(synth.heatmap_logits(2, seed=8)).cuda()
    br = pipe.predict_step(logits, seed=4, return_pred_samples=True, id="patient0001/2CH")
    assert br.pred.shape == (2, 256, 256) and br.entropy_map.shape == (2, 256, 256)
    assert br.contour_samples.shape == (2, 1, 20, 21, 2) and br.pred_samples.shape == (2, 1, 20, 1, 256, 256)
    assert br.mu.shape == (2, 21, 2) and br.cov.shape == (2, 21, 2, 2) and br.sample_area.shape == (2, 1, 20)
    # the record is self-consistent with the reference's own definitions
    assert np.array_equal(br.pred, br.pred_samples.mean(axis=(1, 2)).squeeze().round().astype(int))
    assert np.array_equal(br.sample_area[:, 0], br.pred_samples[:, 0, :, 0].sum((-2, -1)))
    for i in range(2):
        o = port.uncertainty_from_samples(br.contour_samples[i, 0])
        assert np.array_equal(o["pred"], br.pred[i]) and np.abs(o["entropy"] - br.entropy_map[i]).max() <= 1e-12
    S, cov = pipe.predict([logits, logits])
    assert S.shape == (2, 2, 21, 2) and cov.shape == (2, 2, 21, 2, 2)


def test_dsnt_mirror_functions():
    from contouring_uncertainty_b200.task.regression.dsnt import utils as U

    logits = torch.from_numpy(synth.heatmap_logits(1, seed=9)).cuda()
    heat = U.flat_softmax(logits)
    coords, var, covar = U.dsnt(heat)
    (c2, v2, cv2), (px, sig) = U.dsnt_from_logits(logits)
    assert coords.shape == (1, 21, 2) and var.shape == (1, 21, 2) and covar.shape == (1, 21)
    assert (coords - c2).abs().max() < 2e-6 and ((var - v2).abs() / v2).max() < 2e-5
    assert (U.normalized_to_pixel_coordinates(c2, 256) - px).abs().max() < 3e-4
    assert (U.pixel_to_normalized_coordinates(px, 256) - c2).abs().max() < 3e-6
    o_mu, o_sig = port.dsnt_predict_on_batch(logits.cpu().numpy())
    assert np.abs(px.cpu().numpy() - o_mu).max() < 2.6e-4


def test_run_into_and_host_pipeline_match_allocating_path():
    from contouring_uncertainty_b200 import UncertaintyEngine
    from contouring_uncertainty_b200.engine import HostPipeline

    dev = torch.device("cuda:0")
    eng = UncertaintyEngine(device=dev)
    logits = torch.from_numpy(synth.heatmap_logits(4, seed=15)).to(dev)
    ref = eng(logits, 30, seed=77)
    ws = eng.run_into(eng.workspace(4, 30), logits, seed=77)
    torch.cuda.synchronize()
    for k in ("mu", "cov", "area", "count", "mean", "entropy", "pred"):
        assert torch.equal(getattr(ws, k), getattr(ref, k)), k
    assert torch.equal(ws.contours, ref.contour_samples)
    # host pipeline: the draws are keyed by (seed, global sample index), so the result does not depend on how the batch is
    # cut into chunks: chunks of 2 images on 2 streams and chunks of 1 image on 3 streams both reproduce the whole-batch call
    r = eng(logits, 30, seed=5)
    h_in = logits.cpu().pin_memory()
    for chunk, n_streams in ((2, 2), (1, 3)):
        pipe = HostPipeline(eng, 4, 30, chunk=chunk, n_streams=n_streams)
        h_out = pipe.run(h_in, pipe.host_outputs(), seed=5)
        torch.cuda.synchronize()
        assert torch.equal(h_out["area"], r.area.cpu())
        assert torch.equal(h_out["pred"], r.pred.cpu())
        assert torch.equal(h_out["entropy"], r.entropy.cpu())
        assert torch.equal(h_out["contours"], r.contour_samples.cpu())
    # a different seed gives different draws
    assert not torch.equal(eng(logits, 30, seed=6).area, r.area)
    with pytest.raises(ValueError):
        pipe.run(logits, h_out)                               # device tensor where host logits are expected


def test_predict_step_with_epistemic_passes_matches_oracle():
    """T_e = 3 passes: aggregation over T_e*T_a samples, epistemic+aleatoric covariance, posterior moments."""
    from contouring_uncertainty_b200.task import ContourUncertaintyPipeline

    pipe = ContourUncertaintyPipeline(t_a=15)
    passes = [torch.from_numpy(synth.heatmap_logits(2, seed=30 + e)).cuda() for e in range(3)]
    br = pipe.predict_step(passes, seed=9, return_pred_samples=True)
    assert br.contour_samples.shape == (2, 3, 15, 21, 2) and br.pred_samples.shape == (2, 3, 15, 1, 256, 256)
    mus, covs = zip(*[port.dsnt_predict_on_batch(p.cpu().numpy()) for p in passes])
    mu_o, cov_o = port.aleatoric_epistemic_moments(np.stack(mus, 1), np.stack(covs, 1))
    assert np.abs(br.mu - mu_o).max() < 3e-4 and np.abs(br.cov - cov_o).max() < 2e-3
    pm_o, pc_o = port.posterior_moments(br.contour_samples)
    assert np.abs(br.post_mu - pm_o).max() < 1e-4                     # reference means in fp32, ours in fp64
    # the reference's per-pass means are fp32 (numpy mean of a float32 array), which leaks ~1e-5 px^2 into
    # the epistemic term; ours are fp64
    assert np.abs(br.post_cov - pc_o).max() < 5e-5 * np.abs(pc_o).max()
    for i in range(2):
        masks = port.convert_to_mask(br.contour_samples[i].reshape(-1, 21, 2))
        assert np.array_equal(masks.reshape(3, 15, 256, 256), br.pred_samples[i, :, :, 0])
        o = port.maps_from_counts(masks.sum(0, dtype=np.int64), 45)
        assert np.array_equal(o[2], br.pred[i]) and np.abs(o[1] - br.entropy_map[i]).max() <= 1e-12
        assert np.array_equal(br.sample_area[i].reshape(-1), masks.reshape(45, -1).sum(1))


def test_sequence_sampler_vs_golden_and_pair_rasterisation():
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.sampler import SequencePSMSampler

    g = np.load(os.path.join(G, "seq.npz"))
    s = SequencePSMSampler()
    npair, T = g["z"].shape[:2]
    # single pair, CPU tensors in -> (n,2,K,2) CPU out, like the reference
    out = s(torch.from_numpy(g["mu"][0]), torch.from_numpy(g["cov"][0]), n=T, z=g["z"][0], first_instant=g["bits"][0])
    assert out.shape == (T, 2, 21, 2) and out.device.type == "cpu"
    # vs the reference's code executed in fp64: the second instant goes through two fp32 hand-overs
    # (mu_f, cov_f) -> 1e-5 of the 256-px scale
    assert np.abs(out.numpy() - g["samples_f64"][0]).max() < 2.6e-3
    assert np.abs(out.numpy() - g["samples"][0]).max() < 0.2                    # verbatim fp32 reference
    # batched
    outb = s(torch.from_numpy(g["mu"]).cuda(), torch.from_numpy(g["cov"]).cuda(), n=T, z=g["z"], first_instant=g["bits"])
    assert outb.shape == (npair, T, 2, 21, 2)
    assert np.abs(outb.cpu().numpy() - g["samples_f64"]).max() < 2.6e-3
    # ED/ES coupling: with random bits / Philox the areas of the two instants are positively correlated
    big = s(torch.from_numpy(g["mu"][:1]).cuda(), torch.from_numpy(g["cov"][:1]).cuda(), n=2000, seed=4)[0]   # (T,2,K,2)
    cont = big.permute(1, 0, 2, 3).reshape(-1, 21, 2).contiguous()
    _, area = E.rasterize_contours(cont, return_masks=False)
    a = area.reshape(2, 2000).double()
    assert torch.corrcoef(a)[0, 1].item() > 0.3
    fac = E.fac_from_areas(area[:2000].contiguous(), area[2000:].contiguous())
    assert torch.isfinite(fac).float().mean().item() > 0.95


def test_pipeline_sequence_sampler_view():
    """dsnt-al with task.sequence_sampler=True (tmi_scripts/test.sh): one view = (ED, ES)."""
    from contouring_uncertainty_b200.task import ContourUncertaintyPipeline
    from contouring_uncertainty_b200.utils.clinical import fac_mc, aleatoric_epistemic_uncertainty

    pipe = ContourUncertaintyPipeline(t_a=40, sequence_sampler=True)
    logits = torch.from_numpy(synth.heatmap_logits(2, seed=41)).cuda()
    br = pipe.predict_step(logits, seed=3, return_pred_samples=True)
    assert br.contour_samples.shape == (2, 1, 40, 21, 2) and br.sample_area.shape == (2, 1, 40)
    for i in range(2):
        masks = port.convert_to_mask(br.contour_samples[i, 0])
        assert np.array_equal(masks, br.pred_samples[i, 0, :, 0])
        assert np.array_equal(br.sample_area[i, 0], masks.reshape(40, -1).sum(1))
        o = port.maps_from_counts(masks.sum(0, dtype=np.int64), 40)
        assert np.array_equal(o[2], br.pred[i]) and np.abs(o[1] - br.entropy_map[i]).max() <= 1e-12
    f = fac_mc(br.sample_area[0], br.sample_area[1])
    want = np.array([[port.lv_fac(br.pred_samples[0, 0, k, 0], br.pred_samples[1, 0, k, 0]) for k in range(40)]])
    want[~((want >= 0) & (want <= 1))] = np.nan
    assert np.allclose(f, want, equal_nan=True)
    assert len(aleatoric_epistemic_uncertainty(f)) == 4


def test_predict_step_umap_and_scalar_uncertainties():
    """dsnt-al predict_step: uncertainty_map = USUMap(mu, cov), point / instant uncertainty dicts
    (contour_uncertainty.py:85-128) against the oracle."""
    from contouring_uncertainty_b200.task import ContourUncertaintyPipeline

    pipe = ContourUncertaintyPipeline(t_a=16)
    logits = torch.from_numpy(synth.heatmap_logits(2, seed=51)).cuda()
    br = pipe.predict_step(logits, seed=9)
    for i in range(2):
        want = port.us_umap(br.mu[i], br.cov[i])
        assert (br.uncertainty_map[i] != want).sum() <= 8
        assert abs(br.instant_uncertainty["cov_projection"][i] - port.projected_uncertainty_value(br.mu[i], br.cov[i])) < 1e-4
        mask = br.pred[i] != 0
        assert np.isclose(br.instant_uncertainty["umap_mean"][i], br.uncertainty_map[i].sum() / mask.sum())
        assert np.isclose(br.instant_uncertainty["entropy_mean"][i], br.entropy_map[i].sum() / mask.sum())
    assert np.allclose(br.point_uncertainty["cov_det"], np.linalg.det(br.cov) ** 0.25)
    assert br.point_uncertainty["post_cov_eigval_sum"].shape == (2, 21)
    assert set(br.instant_uncertainty) == {"cov_det_mean", "cov_eigenvalue_mean", "cov_projection", "umap_mean", "entropy_mean"}


def test_skew_pipeline_predict_step_vs_oracle():
    """dsnt-skew: SkewUncertaintyTask._predict_step + predict_step with the reference's random draws replayed."""
    from contouring_uncertainty_b200.task import SkewUncertaintyPipeline
    from contouring_uncertainty_b200.sampler import SkewPosteriorShapeModelSampler

    n, T = 2, 3
    pipe = SkewUncertaintyPipeline(t_a=T)
    logits = synth.heatmap_logits(n, seed=61)
    raw_alpha = synth.skew_alphas(n, seed=62)                       # skew_block output (before the y negation of dsnt_skew.py:164)
    z = np.random.default_rng(63).standard_normal((n, T, 3, 3)).astype(np.float32)
    ex = np.stack([np.stack([np.stack([synth.exponential_field(64, (i * T + s) * 14 + d) for d in range(14)])
                             for s in range(T)]) for i in range(n)])
    br = pipe.predict_step(torch.from_numpy(logits).cuda(), torch.from_numpy(raw_alpha).cuda(), z=z, expo=ex,
                           return_pred_samples=True)
    assert br.contour_samples.shape == (n, 1, T, 21, 2) and br.alpha.shape == (n, 21, 2) and br.mode.shape == (n, 21, 2)
    alpha = raw_alpha * np.array([1.0, -1.0], dtype=np.float32)
    assert np.array_equal(br.alpha, alpha)
    model = port.ShapeModel.load(pipe.model.path)
    for i in range(n):
        s = port.skew_psm_sample(model, br.mu[i], br.cov[i], alpha[i], z[i], lambda k, d, _i=i: ex[_i, k, d])
        assert np.abs(s - br.contour_samples[i, 0]).max() < 2e-4
        masks = port.convert_to_mask(br.contour_samples[i, 0])
        assert np.array_equal(masks, br.pred_samples[i, 0])
        o = port.maps_from_counts(masks.sum(0, dtype=np.int64), T)
        assert np.abs(o[1] - br.entropy_map[i]).max() <= 1e-12
        mode, um = port.us_skew_umap(br.mu[i], br.cov[i], alpha[i])
        assert np.abs(mode - br.mode[i]).max() < 1e-4 and (np.abs(um - br.uncertainty_map[i]) > 1e-12).sum() <= 8
        assert np.array_equal(br.pred[i], port.us_contour_to_mask(br.mode[i]))
        assert np.allclose(br.post_mu[i], br.contour_samples[i].reshape(-1, 21, 2).mean(0), atol=1e-4)
    # the sampler mirror gives the same contours through the reference's call signature
    smp = SkewPosteriorShapeModelSampler()
    out = smp(torch.from_numpy(br.mu), torch.from_numpy(br.cov), torch.from_numpy(alpha), n=T, z=z, expo=ex)
    assert out.shape == (n, T, 21, 2) and np.array_equal(out.numpy(), br.contour_samples[:, 0])


def test_new_paths_refuse_bad_inputs_and_handle_empty_batches():
    """Error behaviour of the skew / u-map / Simpson host layer: shape errors raise ValueError, CPU tensors are refused,
    empty batches return empty results."""
    from contouring_uncertainty_b200 import engine as E

    dev = torch.device("cuda:0")
    model = E.ShapeModelConstants(device=dev)
    mu = torch.zeros((0, 21, 2), device=dev)
    cov = torch.zeros((0, 21, 2, 2), device=dev)
    al = torch.zeros((0, 21, 2), device=dev)
    assert E.skew_sample(model, mu, cov, al, 4).shape == (0, 4, 21, 2)
    assert E.skew_pdf_grid(mu, cov, al).shape == (0, 21, 256, 256)
    assert E.umap_analytic(mu, cov).shape == (0, 256, 256)
    assert E.projected_uncertainty(mu, cov)[0].shape == (0, 21)
    assert E.contour_perimeter(torch.zeros((0, 21, 2), device=dev)).shape == (0,)
    d, s = E.simpson_diameters(torch.zeros((0, 2048), dtype=torch.int32, device=dev), 0.5)
    assert d.shape == (0, 20) and s.shape == (0,)
    with pytest.raises(ValueError):
        E.skew_sample(model, torch.zeros((2, 20, 2), device=dev), torch.zeros((2, 20, 2, 2), device=dev), torch.zeros((2, 20, 2), device=dev), 4)
    with pytest.raises(ValueError):
        E.simpson_volumes(torch.zeros((1, 3, 4, 20), dtype=torch.float64, device=dev), torch.zeros((1, 3, 4), dtype=torch.float64, device=dev))
    with pytest.raises(ValueError, match="no CPU implementation"):
        E.umap_analytic(torch.zeros((1, 21, 2)), torch.zeros((1, 21, 2, 2)))
    with pytest.raises(ValueError):
        E.simpson_diameters(torch.zeros((4, 100), dtype=torch.int32, device=dev), 0.5)
    with pytest.raises(ValueError):
        E.simpson_diameters(torch.zeros((4, 2048), dtype=torch.int32, device=dev), -1.0)


def test_distribution_mirrors_vs_reference():
    """contour_uncertainty.distributions classmethods: logpdf / pdf (kernel, fp64 inside) against the reference's fp32
    torch evaluation; rvs_fast with replayed normals; marginal."""
    from contouring_uncertainty_b200.distributions import BivariateNormal, BivariateSkewNormal

    g = np.load(os.path.join(G, "distributions.npz"))
    t = torch.from_numpy
    x, mu, cov, al = t(g["x"]), t(g["mu"]), t(g["cov"]), t(g["alpha"])
    # the reference evaluates the exponent in fp32: |d logpdf| ~ 1e-5 at |logpdf| ~ 50
    assert np.abs(BivariateNormal.logpdf(x, mu, cov).numpy() - g["n_logpdf"]).max() < 2e-5
    # skew factor: the reference forms 0.5 (1 + erf) in fp32, which cancels catastrophically once Phi < ~1e-4 (its values
    # are then quantised to 3e-8 under the + 1e-7 floor): compare logpdf where Phi is resolved, pdf (absolute) everywhere
    sl = BivariateSkewNormal.logpdf(x, mu, cov, al).numpy()
    resolved = (g["s_logpdf"] - g["n_logpdf"] - np.log(2)) > -9
    assert resolved.sum() > 100 and np.abs(sl - g["s_logpdf"])[resolved].max() < 2e-4
    np.testing.assert_allclose(BivariateNormal.pdf(x, mu, cov).numpy(), g["n_pdf"], rtol=3e-5, atol=1e-12)
    assert np.abs(BivariateSkewNormal.pdf(x, mu, cov, al).numpy() - g["s_pdf"]).max() < 2e-4 * g["s_pdf"].max()
    got = BivariateSkewNormal.logpdf(t(g["x_n"]), t(g["mu_n"]), t(g["cov_n"]), t(g["alpha_n"])).numpy()
    res_n = g["s_logpdf_n"] > -12
    assert np.abs(got - g["s_logpdf_n"])[res_n].max() < 2e-4
    # tighter: against the fp64 oracle
    want = port.skew_logpdf(g["x"].astype(np.float64), g["mu"], g["cov"], g["alpha"].astype(np.float64))
    assert np.abs(BivariateSkewNormal.logpdf(x, mu, cov, al).numpy() - want).max() < 5e-6
    r = BivariateSkewNormal.rvs_fast(mu, cov, al, size=(16,), z=g["z"])
    assert r.shape == (16, 2) and np.abs(r.numpy() - g["rvs"]).max() < 1e-4
    m0, v0, a0 = BivariateSkewNormal.marginal(mu, cov, g["alpha"], axis=0, angle=torch.tensor(0.4))
    np.testing.assert_allclose([float(m0), float(v0), float(a0)], g["marginal"], rtol=2e-6)
    # BivariateNormal.rvs with replayed normals, BivariateSkewNormal.mode / BivariateNormal.mode
    r2 = BivariateNormal.rvs(mu, cov, size=(16,), z=g["z2"])
    assert r2.shape == (16, 2) and np.abs(r2.numpy() - g["n_rvs"]).max() < 1e-4
    emp = BivariateNormal.rvs(mu, cov, size=(40000,), seed=9).numpy()
    assert np.abs(emp.mean(0) - g["mu"]).max() < 0.1 and np.abs(np.cov(emp.T) - g["cov"]).max() < 0.05 * np.abs(g["cov"]).max()
    for k in range(8):
        m = BivariateSkewNormal.mode(t(g["mu_n"][k]), t(g["cov_n"][k]), t(g["alpha_mode"][k])).numpy()
        np.testing.assert_allclose(m, g["s_mode"][k], rtol=1e-6)
    assert torch.equal(BivariateNormal.mode(mu, cov), mu)
    big = BivariateSkewNormal.rvs_fast(mu, cov, al, size=20000, seed=3).numpy()
    # mean of a skew normal: mu + sqrt(2/pi) delta, delta = cov alpha / sqrt(1 + alpha^T cov alpha)
    c, a = g["cov"].astype(np.float64), g["alpha"].astype(np.float64)
    delta = c @ a / np.sqrt(1 + a @ c @ a)
    assert np.abs(big.mean(0) - (g["mu"] + np.sqrt(2 / np.pi) * delta)).max() < 0.15


def test_numpy_mirrors_of_utils_and_task_classes():
    """utils.{umap, skew_umap, uncertainty_projection, contour, clinical} and the DSNT task classes: reference call
    signatures, numpy in / numpy out."""
    from contouring_uncertainty_b200.task.regression.contour_uncertainty import convert_to_mask
    from contouring_uncertainty_b200.task.regression.dsnt.dsnt_al import DSNTAleatoric
    from contouring_uncertainty_b200.task.regression.dsnt.dsnt_skew import DSNTSkew
    from contouring_uncertainty_b200.utils import clinical, contour, skew_umap, umap, uncertainty_projection

    g = np.load(os.path.join(G, "umap.npz"))
    mu, cov, al = g["mu"][0], g["cov"][0], g["alpha"][0]
    u, v = uncertainty_projection.projected_uncertainty(mu, cov, all=True)
    assert np.abs(u - g["unc_all"][0]).max() < 1e-5 and np.abs(v - g["proj_all"][0]).max() < 1e-11
    assert abs(uncertainty_projection.projected_uncertainty_value(mu, cov) - g["value"][0]) < 1e-3
    um = umap.uncertainty_map(mu, cov)
    assert (um != port.uncertainty_map(mu, cov)).sum() <= 8
    mode, sk = skew_umap.skew_umap(mu, cov, al, linear_close=True)
    assert np.abs(mode - g["mode"][0]).max() < 1e-4 and sk.shape == (256, 256)
    r = np.load(os.path.join(G, "raster.npz"))
    c = r["contours"][5].astype(np.float64)
    assert np.array_equal(contour.reconstruction(c, 256, 256), port.reconstruction(c, 256, 256))
    gl = np.load(os.path.join(G, "gls.npz"))
    assert abs(clinical.perimeter(gl["ed"][0]) - gl["perimeter_ed"][0]) < 1e-7
    assert abs(clinical.global_longitudinal_strain(gl["ed"][0], gl["es"][0]) - gl["gls"][0]) < 1e-8
    s = np.load(os.path.join(G, "simpson.npz"))
    masks = port.convert_to_mask(s["contours"][:, 0])
    edv, esv = clinical.compute_left_ventricle_volumes(masks[0], masks[1], tuple(s["voxelspacing"]), masks[2], masks[3], tuple(s["voxelspacing"]))
    assert (edv, esv) == (int(s["edv"][0]), int(s["esv"][0]))
    with pytest.raises(ValueError):
        clinical.compute_left_ventricle_volumes(np.zeros((256, 256)), masks[1], (0.5, 0.5), masks[2], masks[3], (0.5, 0.5))
    pred, ps = convert_to_mask(s["contours"][:2, 0], (2, 1, 256, 256), s["contours"][:2, :3].reshape(2, 1, 3, 21, 2))
    assert pred.shape == (2, 256, 256) and ps.shape == (2, 1, 3, 1, 256, 256) and np.array_equal(pred[0], masks[0])
    # task classes around a fake "network"
    logits = torch.from_numpy(synth.heatmap_logits(2, seed=71)).cuda()
    task = DSNTAleatoric(model=[lambda img: logits, lambda img: logits * 0.9])
    S, cv = task.predict(None)
    assert S.shape == (2, 2, 21, 2) and cv.shape == (2, 2, 21, 2, 2)
    mu_o, _ = port.dsnt_predict_on_batch(logits.cpu().numpy())
    assert np.abs(S[:, 0].cpu().numpy() - mu_o).max() < 1e-3
    raw = torch.from_numpy(synth.skew_alphas(2, seed=72)).cuda()
    sk_task = DSNTSkew(model=lambda img: (logits, raw), t_e=2)
    S2, c2, a2 = sk_task.predict(None)
    assert a2.shape == (2, 2, 21, 2) and torch.equal(a2[:, 0, :, 1], -raw[..., 1]) and torch.equal(a2[:, 0, :, 0], raw[..., 0])


def test_skew_pipeline_with_epistemic_passes():
    """dsnt-skew with T_e = 2 passes (ensemble): samples of both passes aggregate into one map per image."""
    from contouring_uncertainty_b200.task import SkewUncertaintyPipeline

    n, T = 2, 5
    pipe = SkewUncertaintyPipeline(t_a=T)
    l1 = torch.from_numpy(synth.heatmap_logits(n, seed=81)).cuda()
    l2 = torch.from_numpy(synth.heatmap_logits(n, seed=82)).cuda()
    a1 = torch.from_numpy(synth.skew_alphas(n, seed=83)).cuda()
    a2 = torch.from_numpy(synth.skew_alphas(n, seed=84)).cuda()
    br = pipe.predict_step([l1, l2], [a1, a2], seed=2, return_pred_samples=True)
    assert br.contour_samples.shape == (n, 2, T, 21, 2) and br.pred_samples.shape == (n, 2, T, 256, 256)
    assert br.sample_area.shape == (n, 2, T) and np.array_equal(br.sample_area, br.pred_samples.sum((-2, -1)))
    for i in range(n):
        masks = port.convert_to_mask(br.contour_samples[i].reshape(-1, 21, 2))
        assert np.array_equal(masks.reshape(2, T, 256, 256), br.pred_samples[i])
        o = port.maps_from_counts(masks.sum(0, dtype=np.int64), 2 * T)
        assert np.abs(o[1] - br.entropy_map[i]).max() <= 1e-12
    # epistemic + aleatoric covariance split (aleatoric_skew.py:65-73)
    mu, cov, al = pipe.predict([l1, l2], [a1, a2])
    want_mu, want_cov = port.aleatoric_epistemic_moments(mu.cpu().numpy(), cov.cpu().numpy())
    assert np.allclose(br.mu, want_mu, atol=1e-4) and np.allclose(br.cov, want_cov, rtol=1e-4, atol=1e-4)
    assert np.allclose(br.alpha, al.mean(1).cpu().numpy())


def test_psm_sampler_with_alpha_soft_masks_and_small_api_holes():
    """PosteriorShapeModelSampler(alpha=...) (psm.py:231-237) against the executed reference with replayed draws;
    convert_to_mask(soft_mask=True) against the oracle (scipy restatement of skimage.filters.gaussian: parity unpinned)."""
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.sampler.posterior_shape_model.psm import PosteriorShapeModelSampler
    from contouring_uncertainty_b200.task.regression.contour_uncertainty import convert_to_mask

    g = np.load(os.path.join(G, "psm_alpha.npz"))
    smp = PosteriorShapeModelSampler()
    t = torch.from_numpy
    n_img, n_s = g["samples"].shape[:2]
    out = smp(t(g["mu"]), t(g["cov"]), t(g["alpha"]), n=n_s, z=(g["z3"], g["z2"]))
    assert out.shape == (n_img, n_s, 21, 2)
    assert np.abs(out.numpy() - g["samples_f64"]).max() < 1e-3          # vs the reference executed in fp64
    assert np.abs(out.numpy() - g["samples"]).max() < 6e-2              # vs its fp32 run (its own noise: 1.8e-2 px)
    one = smp(t(g["mu"][0]), t(g["cov"][0]), t(g["alpha"][0]), n=4, seed=3)    # single image, generated draws
    assert one.shape == (4, 21, 2) and torch.isfinite(one).all()
    # skewness moves the mean of the initial points: E[s] - mu = sqrt(2/pi) delta, delta = cov a / sqrt(1 + a^T cov a), a = alpha * [1, -1]
    big = smp(t(g["mu"][0]), t(g["cov"][0]), t(g["alpha"][0]), n=20000, seed=5).numpy()
    a = g["alpha"][0, 10].astype(np.float64) * np.array([1.0, -1.0]); c = g["cov"][0, 10].astype(np.float64)
    want = np.sqrt(2 / np.pi) * c @ a / np.sqrt(1 + a @ c @ a)
    assert np.abs(big[:, 10].mean(0) - g["mu"][0, 10] - want).max() < 0.15
    # soft masks
    r = np.load(os.path.join(G, "raster.npz"))
    c4 = r["contours"][:4]
    pred, ps = convert_to_mask(c4[:1], (1, 1, 256, 256), c4.reshape(1, 1, 4, 21, 2), soft_mask=True)
    assert ps.shape == (1, 1, 4, 1, 256, 256) and ps.dtype == np.float64
    for k in range(4):
        want = port.soft_mask(port.us_contour_to_mask(c4[k]))
        assert np.abs(ps[0, 0, k, 0] - want).max() <= 1e-12
        assert ps[0, 0, k, 0].min() == 0.0 and ps[0, 0, k, 0].max() == 1.0


def test_graphed_step_replays_match_eager_steps_and_fill_the_metric_ring():
    """engine.GraphedStep: replay i of the captured step == the eager step with sample_offset = i * n_img * t_a (the device-side
    step counter shifts the Philox sample indices), and cub200_metric_pack leaves areas + exact int64 sums in ring slot i % n_slots."""
    import synth
    from contouring_uncertainty_b200 import engine as E

    dev = torch.device("cuda:0")
    n, T, slots = 3, 40, 4
    eng = E.UncertaintyEngine(device=dev)
    logits = torch.from_numpy(synth.heatmap_logits(n, seed=17)).to(dev)
    ws_g, ws_e = eng.workspace(n, T), eng.workspace(n, T)
    ring = torch.zeros((slots, n * T + 4 * E.METRIC_PACK_PARTS), dtype=torch.int32, device=dev)
    gs = E.GraphedStep(eng, ws_g, logits, seed=5, metric_ring=ring)
    assert gs.kernels_per_replay >= 7
    for i in range(6):
        gs.replay()
        eng.run_into(ws_e, logits, seed=5, sample_offset=i * n * T)
        torch.cuda.synchronize()
        assert torch.equal(ws_g.contours, ws_e.contours) and torch.equal(ws_g.area, ws_e.area)
        assert torch.equal(ws_g.count, ws_e.count) and torch.equal(ws_g.entropy, ws_e.entropy)
        slot = ring[i % slots]
        assert torch.equal(slot[:n * T], ws_e.area.reshape(-1))
        a = ws_e.area.reshape(-1).to(torch.int64)
        assert slot[n * T:].view(torch.int64).reshape(-1, 2).sum(0).tolist() == [int(a.sum()), int((a * a).sum())]
    assert int(gs.step.item()) == 6
    gs.resync(1, slots)                                   # next replay writes slot 1 with sample indices never used before
    assert int(gs.step.item()) == 9
    gs.replay()
    torch.cuda.synchronize()
    assert torch.equal(ring[1][:n * T], ws_g.area.reshape(-1))


def test_pipelined_graphed_steps_match_eager_steps():
    """engine.GraphedStep(prefetch=...): a cycle of batches where the graph of batch b also runs the DSNT pass of batch b + 1 on a
    side branch (two alternating pairs of moment buffers).  Replay i must leave exactly what the eager step i leaves: same
    kernels on the same inputs, only their order in time differs."""
    import synth
    from contouring_uncertainty_b200 import engine as E

    dev = torch.device("cuda:0")
    n, T, P = 3, 40, 4
    eng = E.UncertaintyEngine(device=dev)
    pool = [torch.from_numpy(synth.heatmap_logits(n, seed=20 + b)).to(dev) for b in range(P)]
    ws_g, ws_e = eng.workspace(n, T), eng.workspace(n, T)
    moments = [(ws_g.mu, ws_g.cov), (torch.empty_like(ws_g.mu), torch.empty_like(ws_g.cov))]
    for b in (0, 1):
        eng.dsnt_into(ws_g, pool[b], *moments[b])
    gs, step = [], None
    for b in range(P):
        g = E.GraphedStep(eng, ws_g, pool[b], seed=5, step=step, prefetch=(pool[(b + 1) % P], moments[b % 2], moments[(b + 1) % 2]))
        step = g.step
        gs.append(g)
    with pytest.raises(ValueError):
        E.GraphedStep(eng, ws_g, pool[0], seed=5, prefetch=(pool[1], moments[0], moments[0]))
    eng.dsnt_into(ws_g, pool[0], *moments[0])             # nobody prefetched the first batch
    for i in range(7):
        b = i % P
        gs[b].replay()
        # the eager step, with the prefetch branch's launch shape of the DSNT pass (one CTA per heatmap) for batches the graphs
        # prefetched; the persistent launch agrees with it to fp32 rounding (checked below)
        eng.dsnt_into(ws_e, pool[b], ws_e.mu, ws_e.cov, persistent=(i == 0))
        eng.run_tail_into(ws_e, ws_e.mu, ws_e.cov, seed=5, sample_offset=i * n * T)
        torch.cuda.synchronize()
        assert torch.equal(moments[b % 2][0], ws_e.mu) and torch.equal(moments[b % 2][1], ws_e.cov)
        mu_p, cov_p = torch.empty_like(ws_e.mu), torch.empty_like(ws_e.cov)
        eng.dsnt_into(ws_e, pool[b], mu_p, cov_p)
        assert float((mu_p - ws_e.mu).abs().max()) <= 1e-4 and float(((cov_p - ws_e.cov).abs() / (cov_p.abs() + 1e-2)).max()) <= 5e-5      # px, relative
        assert torch.equal(ws_g.contours, ws_e.contours) and torch.equal(ws_g.area, ws_e.area)
        assert torch.equal(ws_g.count, ws_e.count) and torch.equal(ws_g.entropy, ws_e.entropy)
    assert int(step.item()) == 7

"""The mirrored task classes against the reference's own predict_step, executed verbatim (tests/golden/task_al.npz,
task_skew.npz: DSNTAleatoric / DSNTSkew -> _predict_step -> ContourUncertaintyTask.predict_step on synthetic logits through a
two-member stub ensemble, random draws recorded).  Every BatchResult field is compared."""
import os

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _unpack(a):
    return np.unpackbits(a, axis=-1, bitorder="little")


def _logits(g):
    n, seed = int(g["shape"][0]), int(g["seed"])
    base, other = synth.heatmap_logits(n, seed=seed), synth.heatmap_logits(n, seed=seed + 1)
    return torch.from_numpy(np.stack([base, 0.7 * base + 0.3 * other]).astype(np.float32)).cuda()


def _batch(n):
    from contour_uncertainty.data.config import ContourTags              # through the opt-in alias
    from contour_uncertainty.data.config import Tags

    return {Tags.img: torch.zeros((n, 1, 256, 256), device="cuda"), Tags.gt: torch.zeros((n, 256, 256)),
            ContourTags.contour: torch.zeros((n, 21, 2)), Tags.id: [f"patient{i:04d}" for i in range(n)]}


@pytest.fixture()
def alias():
    import contouring_uncertainty_b200.compat as compat

    compat.install_alias()
    yield
    compat.uninstall_alias()


def _check_scalars(res, g, tag, post=True):
    for k in ("cov_xx", "cov_yy", "cov_det", "cov_eigval_sum"):
        np.testing.assert_allclose(res.point_uncertainty[k], g[f"{tag}_point_{k}"], rtol=2e-5, err_msg=k)
    if post:
        for k in ("post_cov_xx", "post_cov_yy", "post_cov_det", "post_cov_eigval_sum"):
            assert res.point_uncertainty[k].shape == g[f"{tag}_point_{k}"].shape
    for k in ("cov_det_mean", "cov_eigenvalue_mean", "cov_projection"):
        np.testing.assert_allclose(res.instant_uncertainty[k], g[f"{tag}_instant_{k}"], rtol=2e-5, err_msg=k)
    assert set(res.instant_uncertainty) == {"cov_det_mean", "cov_eigenvalue_mean", "cov_projection", "umap_mean", "entropy_mean"}


def test_dsnt_al_predict_step_matches_the_executed_reference(alias):
    """dsnt-al: the two-line swap of INTEGRATION.md -- the reference's import path, its constructor arguments, its
    predict_step(batch, batch_idx) -- reproduces the reference's BatchResult with the recorded draws."""
    from contour_uncertainty.task.regression.dsnt.dsnt_al import DSNTAleatoric

    g = np.load(os.path.join(G, "task_al.npz"))
    n, te, ta = (int(v) for v in g["shape"])
    logits = _logits(g)
    task = DSNTAleatoric(covar=True, psm_path=None, seq_psm_path=None, sequence_sampler=False, t_a=ta, t_e=te,
                         model=[lambda img, _e=e: logits[_e] for e in range(te)])
    task.on_predict_start()
    res = task.predict_step(_batch(n), 0, z=g["z"])
    # heads + bookkeeping (fp32 DSNT: 1e-5 relative)
    np.testing.assert_allclose(res.mu, g["f64_mu"], rtol=0, atol=8e-4)
    np.testing.assert_allclose(res.cov, g["f64_cov"], rtol=2e-5, atol=2e-4)
    assert np.array_equal(res.mode, res.mu)
    # samples: vs the reference's sampler executed in fp64 (tight) and as it runs in fp32 (its own noise ~4e-2 px)
    assert res.contour_samples.shape == (n, te, ta, 21, 2)
    assert np.abs(res.contour_samples - g["f64_contour_samples"]).max() < 1e-3
    assert np.abs(res.contour_samples - g["f32_contour_samples"]).max() < 6e-2
    np.testing.assert_allclose(res.post_mu, g["f64_post_mu"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(res.post_cov, g["f64_post_cov"], rtol=1e-3, atol=2e-3)
    # masks: bit-exact for every sample whose rounded landmarks agree with the reference's (all of them here)
    same = (np.round(res.contour_samples) == np.round(g["f64_contour_samples"])).all((-1, -2))
    assert same.all()
    want_ps = _unpack(g["f64_pred_samples"]).reshape(g["f64_pred_samples_shape"])
    assert res.pred_samples.shape == want_ps.shape == (n, te, ta, 256, 256)
    assert np.array_equal(res.pred_samples, want_ps)
    assert np.array_equal(res.pred, g["f64_pred"])
    assert np.abs(res.entropy_map - g["f64_entropy_map"]).max() <= 1e-12
    assert np.array_equal(res.sample_area, want_ps.sum((-2, -1)))
    # analytic u-map: 0 differing pixels
    assert (res.uncertainty_map != g["f64_uncertainty_map"]).sum() == 0
    _check_scalars(res, g, "f64")
    np.testing.assert_allclose(res.instant_uncertainty["umap_mean"], g["f64_instant_umap_mean"], rtol=1e-12)
    np.testing.assert_allclose(res.instant_uncertainty["entropy_mean"], g["f64_instant_entropy_mean"], rtol=1e-12)
    for k in ("post_cov_xx", "post_cov_yy", "post_cov_det", "post_cov_eigval_sum"):
        np.testing.assert_allclose(res.point_uncertainty[k], g[f"f64_point_{k}"], rtol=2e-3, err_msg=k)
    assert res.id == [f"patient{i:04d}" for i in range(n)] and res.gt.shape == (n, 256, 256) and res.contour.shape == (n, 21, 2)
    # convert_to_mask / sample_entropy of the task on the reference's own samples: the aggregation alone, bit-exact
    pred, ps = task.convert_to_mask(g["f64_mu"], (n, 1, 256, 256), g["f64_contour_samples"])
    assert np.array_equal(ps[:, :, :, 0], want_ps)
    ent = np.array([task.sample_entropy(ps[i].astype(float)) for i in range(n)])
    assert np.abs(ent - g["f64_entropy_map"]).max() <= 1e-12
    # and sample(): the batched sampler in the task's (image, pass) order
    cs = task.sample(torch.from_numpy(g["mu_te"]), torch.from_numpy(g["cov_te"]), ta, z=g["z"])
    assert cs.shape == (n, te, ta, 21, 2) and np.abs(cs - g["f64_contour_samples"]).max() < 1e-3


def test_dsnt_skew_predict_step_matches_the_executed_reference(alias):
    """dsnt-skew: 25 samples per pass (hard-coded in the reference), rvs_fast normals and the Exp(1) fields of every
    torch.multinomial replayed: grid indices, hence level points, masks, pred and entropy bit-exact."""
    from contour_uncertainty.task.regression.dsnt.dsnt_skew import DSNTSkew

    g = np.load(os.path.join(G, "task_skew.npz"))
    n, te, T = (int(v) for v in g["shape"])
    logits = _logits(g)
    raw = torch.from_numpy(g["raw_alpha"]).cuda()
    task = DSNTSkew(covar=True, psm_path=None, seq_psm_path=None, skew_indices=None, t_a=T, t_e=te,
                    model=[lambda img, _e=e: (logits[_e], raw[_e].reshape(n, 42)) for e in range(te)])
    qs = int(g["q_seed"])
    expo = np.stack([np.stack([np.stack([np.stack([synth.exponential_field(qs, ((i * te + e) * T + s) * 14 + d) for d in range(14)])
                                         for s in range(T)]) for e in range(te)]) for i in range(n)])
    res = task.predict_step(_batch(n), 0, z=g["z"], expo=expo)
    np.testing.assert_allclose(res.mu, g["f32_mu"], rtol=0, atol=8e-4)
    np.testing.assert_allclose(res.cov, g["f32_cov"], rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(res.alpha, g["f32_alpha"], rtol=1e-6)
    # level points are integer grid coordinates chosen by the replayed races: identical; the 4 filled-in points carry the
    # reference's fp32 noise (sigma2 = 1e-3 inverse)
    level = [p for p in range(21) if p not in (0, 10, 20, 4, 9, 11, 16)]
    assert np.array_equal(res.contour_samples[..., level, :], g["f32_contour_samples"][..., level, :])
    # (the reference's fp32 42x42 inverse at sigma2 = 1e-3 moves them by up to ~7e-2 px here; golden skew.npz holds an fp64
    # run of the same sampler, and the fp64 oracle is compared below)
    assert np.abs(res.contour_samples - g["f32_contour_samples"]).max() < 1e-1
    from oracle import port
    model = port.ShapeModel.load(task.sampler.model.path)
    for (i, e) in ((0, 1), (1, 0)):
        so = port.skew_psm_sample(model, g["mu_te"][i, e], g["cov_te"][i, e], g["alpha_te"][i, e], g["z"][i, e][:2],
                                  lambda n_, d, _i=i, _e=e: synth.exponential_field(qs, ((_i * te + _e) * T + n_) * 14 + d))
        assert np.abs(res.contour_samples[i, e, :2] - so).max() < 1e-3
    np.testing.assert_allclose(res.post_mu, g["f32_post_mu"], rtol=0, atol=2e-2)
    np.testing.assert_allclose(res.post_cov, g["f32_post_cov"], rtol=1e-2, atol=5e-2)     # the 4 filled-in points: fp32 noise x spread
    level_cov = np.asarray(res.post_cov)[:, level]                                          # level points are identical samples
    np.testing.assert_allclose(level_cov, g["f32_post_cov"][:, level], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(res.mode, g["f32_mode"], rtol=0, atol=1e-3)
    assert np.array_equal(res.pred, g["f32_pred"])
    # skew u-map: 200 level contours are rasterised per image; the task's own moments differ from the reference's fp32 DSNT
    # by ~1e-5 relative, which can move a level contour across a pixel centre in one of the 200 masks (that pixel's weighted
    # mean then changes by one weight).  On the task's moments: a handful of pixels; on the REFERENCE's moments: every pixel.
    d = np.abs(res.uncertainty_map - g["f32_uncertainty_map"])
    assert (d > 1e-12).sum() <= 32 and d.max() < 5e-2
    from contour_uncertainty.data.camus.utils import USSkewUmap
    for i in range(n):
        mode_i, um_i = USSkewUmap()(g["f32_mu"][i], g["f32_cov"][i], g["f32_alpha"][i], labels=[])
        assert np.abs(um_i - g["f32_uncertainty_map"][i]).max() <= 1e-12 and np.abs(mode_i - g["f32_mode"][i]).max() < 1e-4
    want_ps = _unpack(g["f32_pred_samples"]).reshape(g["f32_pred_samples_shape"])
    same = (np.round(res.contour_samples) == np.round(g["f32_contour_samples"])).all((-1, -2))
    assert same.mean() > 0.9                                   # the filled-in points may round differently (fp32 noise)
    assert np.array_equal(res.pred_samples[same], want_ps[same])
    if same.all():
        assert np.abs(res.entropy_map - g["f32_entropy_map"]).max() <= 1e-12
    else:
        cnt_diff = np.abs(res.pred_samples.sum((1, 2)) - want_ps.sum((1, 2))).max()
        assert cnt_diff <= (~same).sum()
    _check_scalars(res, g, "f32")
    # the aggregation alone on the reference's samples: exact
    pred, ps = task.convert_to_mask(g["f32_mode"], (n, 1, 256, 256), g["f32_contour_samples"])
    assert np.array_equal(ps[:, :, :, 0], want_ps) and np.array_equal(pred, g["f32_pred"])
    ent = np.array([task.sample_entropy(ps[i].astype(float)) for i in range(n)])
    assert np.abs(ent - g["f32_entropy_map"]).max() <= 1e-12


def test_task_constructor_and_functor_contract(alias):
    from contour_uncertainty.data.camus.datamodule import CamusContourDataModule
    from contour_uncertainty.data.config import BatchResult, Label
    from contour_uncertainty.task.regression.aleatoric import AleatoricUncertaintyTask
    from contour_uncertainty.task.regression.contour_uncertainty import ContourUncertaintyTask
    from contour_uncertainty.task.regression.dsnt.dsnt_al import DSNTAleatoric
    from contour_uncertainty.task.uncertainty import UncertaintyTask

    assert issubclass(DSNTAleatoric, AleatoricUncertaintyTask) and issubclass(AleatoricUncertaintyTask, ContourUncertaintyTask)
    assert issubclass(ContourUncertaintyTask, UncertaintyTask)
    dm = CamusContourDataModule(points_per_side=11)
    assert dm.data_params.out_shape == (21, 2) and dm.data_params.labels == (Label.BG, Label.LV)
    logits = torch.from_numpy(synth.heatmap_logits(2, seed=3)).cuda()
    calls = []

    class Functors:                                            # a datamodule-like object with its own u-map functor
        contour_to_mask_fn = dm.contour_to_mask_fn
        skew_umap_fn = dm.skew_umap_fn

        @staticmethod
        def umap_fn(mu, cov, labels=None):
            calls.append(mu.shape)
            return np.full((256, 256), 0.5)

    task = DSNTAleatoric(t_a=4, t_e=3, model=lambda img: logits, data_params=dm.data_params, return_pred_samples=False)
    from types import SimpleNamespace
    task.trainer = SimpleNamespace(datamodule=Functors)
    res = task.predict_step({"img": torch.zeros((2, 1, 256, 256)), "id": ["a", "b"]}, 0)
    assert isinstance(res, BatchResult) and calls == [(21, 2), (21, 2)] and (res.uncertainty_map == 0.5).all()
    assert res.contour_samples.shape == (2, 3, 4, 21, 2) and res.pred_samples is None and res.gt is None and res.contour is None
    assert res.entropy_map.shape == (2, 256, 256) and task.hparams.t_e == 3
    with pytest.raises(NotImplementedError):
        DSNTAleatoric(ensemble_ckpt="some/dir")

"""The clinical metric processors (results/clinical mirrors) against the reference's own AreaError / FAC / Volume.compute
executed on stored BatchResults (tests/golden/metrics.npz), the aleatoric / epistemic split kernel against numpy's
nanmean / nanstd, and the variance / uint16-count outputs of finalize_maps."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
SCALARS = ("pred", "gt", "error", "std", "mean", "aleatoric_std", "epistemic_std")


def _views(g, with_pred_samples=False):
    """The two BatchResults of the golden's patient, rebuilt with this package's rasteriser (sample_area + packed_samples)."""
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.data.camus.utils import USContourToMask
    from contouring_uncertainty_b200.data.config import BatchResult, Label

    te, ta = (int(v) for v in g["shape"])
    views = {}
    for name in ("2CH", "4CH"):
        mu, gt_c, cs = g[f"{name}_mu"], g[f"{name}_gt_contour"], g[f"{name}_contour_samples"]
        mask = USContourToMask()
        pred, gt = mask(mu, (256, 256), labels=[Label.BG, Label.LV]), mask(gt_c, (256, 256), labels=[Label.BG, Label.LV])
        masks, area = E.rasterize_contours(torch.from_numpy(cs.reshape(-1, 21, 2)).cuda(), return_masks=True)
        ps = E.unpack_masks(masks).cpu().numpy().astype(np.int64).reshape(2, te, ta, 256, 256) if with_pred_samples else None
        views[name] = BatchResult(id=f"patient0001/{name}", img=np.zeros((2, 256, 256)), gt=gt, pred=pred, labels=[Label.BG, Label.LV],
                                  uncertainty_map=np.zeros((2, 256, 256)), instants={"ED": 0, "ES": 1},
                                  voxelspacing=tuple(g[f"{name}_voxelspacing"]), mu=mu, mode=mu, cov=np.tile(np.eye(2), (2, 21, 1, 1)),
                                  contour=gt_c, contour_samples=cs, pred_samples=ps, sample_area=area.reshape(2, te, ta).cpu().numpy(),
                                  packed_samples=masks.reshape(2, te, ta, 2048).cpu().numpy())
    return views


def _eq(a, b, rtol=1e-12):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=0, equal_nan=True)


@pytest.mark.parametrize("with_pred_samples", [False, True])
def test_area_fac_volume_processors_match_reference(with_pred_samples):
    from contouring_uncertainty_b200.results.clinical.instant import AreaError
    from contouring_uncertainty_b200.results.clinical.patient import Volume
    from contouring_uncertainty_b200.results.clinical.view import FAC

    g = np.load(os.path.join(G, "metrics.npz"))
    views = _views(g, with_pred_samples)
    if with_pred_samples:                                         # a foreign BatchResult: only the reference's own fields
        for v in views.values():
            v.sample_area = v.packed_samples = None
    for name, v in views.items():
        for key, inst in v.instants.items():
            r = AreaError().compute(v, key, inst)
            _eq([r[k] for k in SCALARS], g[f"area_{name}_{key}"])
            _eq(r["mc"], g[f"area_{name}_{key}_mc"])
        r = FAC().compute(v)
        _eq([r[k] for k in SCALARS], g[f"fac_{name}"])
        _eq(r["mc"], g[f"fac_{name}_mc"])
        assert r["reject"] == bool(g[f"fac_{name}_reject"]) and np.array_equal(r["sample_reject"], g[f"fac_{name}_sample_reject"])
        assert r["sample_reject"].sum() >= 1 and np.isnan(np.asarray(r["mc"])).sum() >= 2     # a rejected sample and an empty-mask NaN
    r = Volume().compute(views)
    for m in ("EF", "EDV", "ESV"):
        _eq([r[f"{m}_{k}"] for k in SCALARS], g[f"vol_{m}"])
        _eq(r[f"{m}_mc"], g[f"vol_{m}_mc"])
    assert r["EF_reject"] == bool(g["vol_EF_reject"]) and np.array_equal(r["EF_sample_reject"], g["vol_EF_sample_reject"])
    # data frames as the reference's __call__ builds them
    df = FAC()(list(views.values()))
    assert list(df.index) == ["patient0001/2CH", "patient0001/4CH"] and "FAC_sample_reject" in df.columns
    adf = AreaError()(list(views.values()))
    assert list(adf.index) == ["patient0001/2CH/ED", "patient0001/2CH/ES", "patient0001/4CH/ED", "patient0001/4CH/ES"]
    pdf = Volume()({"patient0001": views, "patient0002": {"2CH": views["2CH"]}})
    assert list(pdf.index) == ["patient0001"] and float(pdf["EF_mean"].iloc[0]) == pytest.approx(float(g["vol_EF"][4]), rel=1e-12)


def test_clinical_metrics_epoch_end_numbers():
    from contouring_uncertainty_b200.results.clinical.clinical_metrics import ClinicalMetrics
    from contouring_uncertainty_b200.results.clinical.instant import AreaError
    from contouring_uncertainty_b200.results.clinical.patient import Volume
    from contouring_uncertainty_b200.results.clinical.view import FAC
    import copy

    g = np.load(os.path.join(G, "metrics.npz"))
    outs = []
    for p in range(3):                                            # three "patients": the same views under other ids
        for v in _views(g).values():
            w = copy.copy(v)
            w.id = v.id.replace("patient0001", f"patient{p:04d}")
            w.sample_area = v.sample_area + p                      # make the rows differ
            outs.append(w)
    cm = ClinicalMetrics(instant_metrics=(AreaError(),), view_metrics=(FAC(),), patient_metrics=(Volume(),))
    res = cm.on_predict_epoch_end(outputs=[outs])
    assert cm.instant_df.shape[0] == 12 and cm.view_df.shape[0] == 6 and cm.patient_df.shape[0] == 3 and cm.volume_df.shape[0] == 6
    for k in ("Area_gt_pred_correlation", "Area_gt_pred_mae", "Area_uce", "Area_a_uce", "FAC_reject", "FAC_uce", "EF_reject",
              "EF_gt_pred_mae", "Volume_gt_pred_correlation", "Volume_uce"):
        assert k in res, k
    assert res["FAC_reject"] == 0.0 and res["EF_reject"] == 0.0
    assert set(cm.volume_df.columns) >= {"Volume_pred", "Volume_gt", "Volume_std", "Volume_mean", "Volume_error"}


def test_ae_split_kernel_vs_numpy():
    import warnings

    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.utils.clinical import aleatoric_epistemic_uncertainty

    g = np.load(os.path.join(G, "metrics.npz"))
    r = E.metric_ae_split(torch.from_numpy(g["ae_in"]).cuda())
    _eq(r["stats"].cpu().numpy(), g["ae_out"])                    # rows / items that are all NaN give NaN, as numpy does
    rng = np.random.default_rng(5)
    for te, ta in ((1, 1), (1, 7), (3, 33), (40, 5), (2, 1000)):
        v = rng.normal(0.5, 0.4, (6, te, ta))
        v[rng.random(v.shape) < 0.1] = np.nan
        r = E.metric_ae_split(torch.from_numpy(v).cuda(), (0.0, 1.0), return_values=True)
        rej = (v < 0) | (v > 1)
        w = np.where(rej, np.nan, v)
        assert np.array_equal(r["reject"].cpu().numpy().astype(bool), rej) and np.array_equal(r["n_reject"].cpu().numpy(), rej.sum((1, 2)))
        assert np.array_equal(np.isnan(r["values"].cpu().numpy()), np.isnan(w))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = np.array([aleatoric_epistemic_uncertainty(x) for x in w])
        _eq(r["stats"].cpu().numpy(), want, rtol=1e-11)
    with pytest.raises(ValueError):
        E.metric_ae_split(torch.zeros((2, 3), dtype=torch.float64, device="cuda"))


def test_variance_and_count16_outputs_of_finalize_maps():
    from contouring_uncertainty_b200 import engine as E

    rng = np.random.default_rng(11)
    T = 37
    count = rng.integers(0, T + 1, (3, 256, 256)).astype(np.int32)
    diff = np.diff(count, axis=-1, prepend=0).astype(np.int32)
    c, mean, ent, pred, var, c16 = E.finalize_maps(torch.from_numpy(diff).cuda(), T, want_variance=True, want_count16=True)
    assert np.array_equal(c.cpu().numpy(), count) and np.array_equal(c16.cpu().numpy().astype(np.int32), count)
    p = count.astype(np.float64) / T
    assert np.array_equal(mean.cpu().numpy(), p)
    assert np.array_equal(var.cpu().numpy(), p * (1.0 - p))       # exact: the same two fp64 operations
    # the host expansion of the 2 B / pixel count map reproduces the device maps bit for bit
    h = E.maps_from_count(c16.cpu().numpy(), T)
    assert np.array_equal(h["mean"], mean.cpu().numpy()) and np.array_equal(h["variance"], var.cpu().numpy())
    assert np.array_equal(h["pred"], pred.cpu().numpy()) and np.abs(h["entropy"] - ent.cpu().numpy()).max() <= 1e-15
    out = E.finalize_maps(torch.from_numpy(diff).cuda(), T, want_count=False, want_maps=False, want_count16=True)
    assert out[0] is None and out[1] is None and out[2] is None and np.array_equal(out[4].cpu().numpy().astype(np.int32), count)
    with pytest.raises(ValueError):
        E.finalize_maps(torch.from_numpy(diff).cuda(), 70000, want_count16=True)

"""Mirrored helper (NOT an accelerated component): contour_uncertainty/results/utils/calibration.py, numbers only.

Dataset-level uncertainty calibration of a clinical metric from the per-image / per-patient (error, std) pairs every rank holds
after the end-of-step all_gather (``parallel.gather_metric_distributions``): a few thousand scalars, so this is host numpy by
design.  Same results as the reference's loops (fixture tests/golden/dataset_calibration.json), computed with one
searchsorted / bincount pass instead of a Python loop per bin.  The pixel-level calibration, which is the heavy reduction, is the
kernel ``engine.calibration_bins``."""
import numpy as np


def _filtered(error, uncertainty, filters):
    error, uncertainty = np.asarray(error, dtype=float), np.asarray(uncertainty, dtype=float)
    if filters is not None:
        keep = np.flatnonzero(np.asarray(filters).astype(bool))
        uncertainty, error = uncertainty[keep], error[keep]
    return error, uncertainty


def _binned_uce(error, uncertainty, bin_of, n_bins, weight_den):
    """sum over the non-empty bins of |mean uncertainty - mean error| * (bin size / weight_den); samples with bin_of < 0
    belong to no bin."""
    ok = bin_of >= 0
    size = np.bincount(bin_of[ok], minlength=n_bins)
    conf = np.bincount(bin_of[ok], weights=uncertainty[ok], minlength=n_bins)
    acc = np.bincount(bin_of[ok], weights=error[ok], minlength=n_bins)
    nz = size > 0
    conf, acc, size = conf[nz] / size[nz], acc[nz] / size[nz], size[nz]
    return float(np.sum(np.abs(conf - acc) * (size / weight_den))), list(conf), list(acc), list(size)


def compute_calibration(error, uncertainty, nb_bins: int = 10, filters=None):
    """calibration.py:6-34: equal-width bins over [min, max] of the uncertainty, OPEN at both ends like the reference (a value
    sitting exactly on a boundary -- the minimum and the maximum always do -- belongs to no bin, yet counts in the
    denominator of the bin weights) -> (uce, mean uncertainty, mean error, size per non-empty bin)."""
    error, uncertainty = _filtered(error, uncertainty, filters)
    bounds = np.linspace(uncertainty.min(), uncertainty.max(), nb_bins + 1)
    k = np.searchsorted(bounds, uncertainty, side="left") - 1           # bounds[k] < u <= bounds[k + 1]
    on_boundary = bounds[np.clip(k + 1, 0, nb_bins)] == uncertainty
    bin_of = np.where(on_boundary | (k < 0) | (k >= nb_bins), -1, k)
    return _binned_uce(error, uncertainty, bin_of, nb_bins, len(uncertainty))


def compute_adaptive_calibration(error, uncertainty, nb_bins: int = 10, filters=None):
    """calibration.py:37-65: equal-count bins (np.array_split of the samples sorted by uncertainty); the weight of a bin is
    len(bin) / nb_bins, exactly as the reference writes it."""
    error, uncertainty = _filtered(error, uncertainty, filters)
    order = np.argsort(uncertainty)
    n = len(order)
    big, small = n % nb_bins, n // nb_bins                             # array_split: the first n % nb_bins bins get one more
    sizes = np.array([small + 1] * big + [small] * (nb_bins - big))
    bin_sorted = np.repeat(np.arange(nb_bins), sizes)
    return _binned_uce(error[order], uncertainty[order], bin_sorted, nb_bins, nb_bins)


def metric_calibration(std, error, reject=None):
    """The numbers ClinicalMetrics.plot_calibration logs (contour_uncertainty/results/clinical/clinical_metrics.py:211-226):
    NaN pairs dropped, rejected samples filtered out -> dict(uce, a_uce)."""
    std, error = np.asarray(std, dtype=float), np.asarray(error, dtype=float)
    nas = np.logical_or(np.isnan(std), np.isnan(error))
    filters = None if reject is None else (~np.asarray(reject, dtype=bool))[~nas]
    std, error = std[~nas], error[~nas]
    return {"uce": compute_calibration(error, std, filters=filters)[0],
            "a_uce": compute_adaptive_calibration(error, std, filters=filters)[0]}


def metric_correlation(gt, pred):
    """clinical_metrics.py:163-190: Pearson correlation and mean absolute error over the finite pairs."""
    gt, pred = np.asarray(gt, dtype=float), np.asarray(pred, dtype=float)
    ok = ~(np.isnan(pred) | np.isnan(gt) | np.isinf(pred) | np.isinf(gt))
    g, p = gt[ok], pred[ok]
    gm, pm = g - g.mean(), p - p.mean()
    # scipy.stats.pearsonr: r = sum(xm/|xm| * ym/|ym|), clipped to [-1, 1]
    r = float(np.clip(np.dot(gm / np.linalg.norm(gm), pm / np.linalg.norm(pm)), -1.0, 1.0))
    return {"correlation": r, "mae": float(np.mean(np.abs(g - p)))}

"""Mirror of contour_uncertainty/results/utils/calibration.py (numbers only, no plotting): dataset-level uncertainty
calibration of a clinical metric from the per-image / per-patient (error, std) pairs every rank holds after the end-of-step
all_gather (``parallel.gather_metric_distributions``).  Host arithmetic on a few thousand scalars, numpy like the reference."""
import numpy as np


def _filtered(error, uncertainty, filters):
    error, uncertainty = np.asarray(error), np.asarray(uncertainty)
    if filters is not None:
        idx = np.where(np.asarray(filters).astype(bool))[0]
        uncertainty, error = uncertainty[idx], error[idx]
    return error, uncertainty


def compute_calibration(error, uncertainty, nb_bins: int = 10, filters=None):
    """calibration.py:6-34: equal-width bins over [min, max] of the uncertainty (open at both ends, like the reference:
    a value sitting exactly on a boundary belongs to no bin) -> (uce, mean uncertainty, mean error, size) per non-empty bin."""
    error, uncertainty = _filtered(error, uncertainty, filters)
    bounds = np.linspace(uncertainty.min(), uncertainty.max(), nb_bins + 1)
    uce = np.zeros(1)
    conf, acc, size = [], [], []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        in_bin = np.greater(uncertainty, lo) * np.less(uncertainty, hi)
        prop = in_bin.mean()
        if prop.item() > 0:
            a, c = error[in_bin].mean(), uncertainty[in_bin].mean()
            uce += np.abs(c - a) * prop
            conf.append(c)
            acc.append(a)
            size.append(in_bin.sum())
    return uce.item(), conf, acc, size


def compute_adaptive_calibration(error, uncertainty, nb_bins: int = 10, filters=None):
    """calibration.py:37-65: equal-count bins (np.array_split of the samples sorted by uncertainty); the weight of a bin is
    len(bin) / nb_bins, exactly as the reference writes it."""
    error, uncertainty = _filtered(error, uncertainty, filters)
    order = np.argsort(uncertainty)
    us, es = np.array_split(uncertainty[order], nb_bins), np.array_split(error[order], nb_bins)
    uce = np.zeros(1)
    conf, acc, size = [], [], []
    for u, e in zip(us, es):
        prop = len(u) / nb_bins
        if prop > 0:
            a, c = e.mean(), u.mean()
            uce += np.abs(c - a) * prop
            conf.append(c)
            acc.append(a)
            size.append(len(u))
    return uce.item(), conf, acc, size


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

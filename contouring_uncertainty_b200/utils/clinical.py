"""Mirror of contour_uncertainty/utils/clinical.py (lv_area, lv_FAC) and results/clinical/utils.py."""
import numpy as np


def lv_area(mask, voxelarea=None):
    """clinical.py:92-93 -> vital Measure.structure_area: number of truthy pixels x voxel area."""
    v = 1 if voxelarea is None else voxelarea
    return np.asarray(mask).astype(bool).sum((-2, -1)) * v


def lv_FAC(ed_mask, es_mask):
    """clinical.py:11-29."""
    ed, es = lv_area(ed_mask), lv_area(es_mask)
    return (ed - es) / ed


def aleatoric_epistemic_uncertainty(metric_mc):
    """results/clinical/utils.py:3-18: (mean, aleatoric, epistemic, total) with ddof = 0 stds."""
    assert metric_mc.ndim == 2, "metric_mc should have 2 dimensions, current shape is {}".format(metric_mc.shape)
    means = np.nanmean(metric_mc, axis=-1)
    stds = np.nanstd(metric_mc, axis=-1)
    mean, epi, alea = np.nanmean(means), np.nanstd(means), np.nanmean(stds)
    return mean, alea, epi, epi + alea


def fac_mc(sample_area_ed, sample_area_es):
    """FAC.compute, Monte-Carlo part (results/clinical/view.py:35-78): per sample index
    (A_ED - A_ES) / A_ED from exact pixel counts, values outside [0, 1] -> NaN."""
    ed = np.asarray(sample_area_ed, dtype=np.float64)
    es = np.asarray(sample_area_es, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = (ed - es) / ed
    f[~((f >= 0) & (f <= 1))] = np.nan
    return f

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


def _contours_t(c):
    import torch

    dev = torch.device("cuda", torch.cuda.current_device())
    return torch.as_tensor(np.asarray(c, dtype=np.float32)).to(dev).contiguous()


def perimeter(contours):
    """clinical.py:30-47: spline arc length of (K,2) or (N,K,2) contours (float / array), computed on the GPU."""
    from .. import engine as E

    c = np.asarray(contours)
    out = E.contour_perimeter(_contours_t(c.reshape(-1, E.K, 2))).cpu().numpy()
    return out[0] if c.ndim == 2 else out


def global_longitudinal_strain(ed_contour, es_contour, spline: bool = True):
    """clinical.py:50-71: (len_ED - len_ES) / len_ED on the spline arc lengths."""
    if not spline:
        raise NotImplementedError("spline=False is never used by the reference's callers")
    ed_len, es_len = perimeter(ed_contour), perimeter(es_contour)
    return (ed_len - es_len) / ed_len


def compute_left_ventricle_volumes(a2c_ed, a2c_es, a2c_voxelspacing, a4c_ed, a4c_es, a4c_voxelspacing):
    """vital/vital/metrics/evaluate/clinical/heart_us.py:13-48: (EDV, ESV) in mL from four binary (256,256) masks.
    Raises like the reference when a mask has no contour / no admissible base chord."""
    import torch

    from .. import engine as E

    dev = torch.device("cuda", torch.cuda.current_device())

    def pack(m):
        b = np.packbits(np.asarray(m).astype(bool).astype(np.uint8), axis=-1, bitorder="little")
        return torch.from_numpy(np.ascontiguousarray(b).view(np.int32).reshape(E.MASK_WORDS)).to(dev)

    d2, s2 = E.simpson_diameters(torch.stack([pack(a2c_ed), pack(a2c_es)]), a2c_voxelspacing)
    d4, s4 = E.simpson_diameters(torch.stack([pack(a4c_ed), pack(a4c_es)]), a4c_voxelspacing)
    diam, step = torch.cat([d2, d4])[None, :, None], torch.cat([s2, s4])[None, :, None]
    if bool(torch.isnan(step).any()):
        raise ValueError("no contour / no admissible base chord in one of the masks")
    edv, esv, _ = E.simpson_volumes(diam.contiguous(), step.contiguous())
    return int(edv.item()), int(esv.item())

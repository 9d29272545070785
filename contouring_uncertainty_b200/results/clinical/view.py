"""Mirror of contour_uncertainty/results/clinical/view.py (ViewMetric, FAC, GLS)."""
from typing import List

import numpy as np
import pandas as pd
import torch

from ... import engine as E
from ...data.config import BatchResult, Label
from ...utils.clinical import global_longitudinal_strain, lv_area
from .utils import ae_split, metric_error, packed_masks, prefix, sample_areas, view_sample_masks


class Instant:
    """vital/vital/data/camus/config.py:36-57."""
    ED: str = "ED"
    ES: str = "ES"


class ViewMetric:
    PREFIX: str = ""

    def compute(self, view: BatchResult):
        raise NotImplementedError

    def __call__(self, view_results: List[BatchResult]) -> pd.DataFrame:
        """view.py:22-27."""
        res = {}
        for view in view_results:
            res[view.id] = prefix(self.compute(view), self.PREFIX)
        return pd.DataFrame(res).T


def _fac(ed_area, es_area):
    """lv_FAC (cu/utils/clinical.py:11-29) on pixel counts."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return (np.asarray(ed_area, dtype=np.float64) - es_area) / ed_area


class FAC(ViewMetric):
    """view.py:30-78: fractional area change per sample index from the per-sample areas, the reject rules as they are
    (a sample outside [0, 1] is rejected and becomes NaN; the view is rejected when pred is outside (0, 1] or more than half of
    its samples are)."""
    PREFIX = 'FAC_'
    MIN_VALUE = 0
    MAX_VALUE = 1

    def compute(self, view: BatchResult):
        instants = view.instants
        ED, ES = instants[Instant.ED], instants[Instant.ES]
        pred = float(_fac(lv_area(view.pred[ED] == 1), lv_area(view.pred[ES] == 1)))
        gt = float(_fac(lv_area(view.gt[ED] == 1), lv_area(view.gt[ES] == 1)))
        area = sample_areas(view)
        s = ae_split(_fac(area[ED], area[ES]), (self.MIN_VALUE, self.MAX_VALUE))       # (T_e, T_a)
        reject = not (self.MIN_VALUE < pred <= self.MAX_VALUE)
        if s['n_reject'] / s['sample_reject'].size > 0.5:
            reject = True
        return {
            'pred': pred, 'gt': gt, 'error': metric_error(s['mean'], gt), 'mc': s['mc'].tolist(), 'std': s['std'], 'mean': s['mean'],
            'aleatoric_std': s['aleatoric_std'], 'epistemic_std': s['epistemic_std'], 'reject': reject,
            'sample_reject': s['sample_reject'],
        }


class GLS(ViewMetric):
    """view.py:81-189, the mask-based part (EchoMeasure.gls on the LV masks; base corners from the polar contour, or from
    EchoMeasure._endo_base when the view's labels hold Label.MYO, :90) and the contour GLS of the prediction / ground truth.
    Like the reference, the mask GLS is measured in pixels (view.py:94 passes no voxelspacing).  The reference's 'contour_mc' feeds MASKS to a function of contours (view.py:121-125), which has
    no meaning; here the Monte-Carlo contour GLS is taken from ``contour_samples`` (what its commented-out code intended)."""
    PREFIX = "GLS_"
    MIN_VALUE = 0
    MAX_VALUE = 1

    @classmethod
    def gls(cls, segmentation, view: BatchResult):
        """view.py:86-96: strain (percent) of the ES frame of one (N_frames, H, W) sequence; NaN when it cannot be measured."""
        seg = np.asarray(segmentation)
        myo = packed_masks(seg == Label.MYO) if Label.MYO in tuple(view.labels) else None
        return float(E.mask_gls(packed_masks(seg == Label.LV), myo_masks=myo)[view.instants[Instant.ES]].cpu())

    def compute(self, view: BatchResult):
        instants = view.instants
        ED, ES = instants[Instant.ED], instants[Instant.ES]
        output = {}
        if view.contour is not None:
            contour_pred = float(global_longitudinal_strain(view.mu[ED], view.mu[ES]))
            contour_gt = float(global_longitudinal_strain(view.contour[ED], view.contour[ES]))
            cs = np.asarray(view.contour_samples)                                      # (N, T_e, T_a, K, 2)
            mc = np.asarray(global_longitudinal_strain(cs[ED].reshape(-1, E.K, 2), cs[ES].reshape(-1, E.K, 2))).reshape(cs.shape[1:3])
            s = ae_split(mc)
            output = {
                'contour_pred': contour_pred, 'contour_gt': contour_gt, 'contour_error': metric_error(s['mean'], contour_gt),
                'contour_mc': mc.tolist(), 'contour_std': s['std'], 'contour_mean': s['mean'],
                'contour_aleatoric_std': s['aleatoric_std'], 'contour_epistemic_std': s['epistemic_std'],
            }
        pred = self.gls(view.pred, view) / -100
        gt = self.gls(view.gt, view) / -100
        if Label.MYO in tuple(view.labels):                                           # label maps: LV = 1, MYO = 2
            ps = np.asarray(view.pred_samples)
            sm, smyo = packed_masks(ps == Label.LV), packed_masks(ps == Label.MYO).permute(1, 2, 0, 3).contiguous()
        else:
            sm, smyo = view_sample_masks(view), None                                  # (N, T_e, T_a, 2048)
        g = E.mask_gls(sm.permute(1, 2, 0, 3).contiguous(), myo_masks=smyo)[..., ES].cpu().numpy()  # per (T_e, T_a): the sequence of N frames
        mc = g / -100
        mc[np.isinf(mc)] = np.nan
        s = ae_split(mc, (self.MIN_VALUE, self.MAX_VALUE))
        reject = not (self.MIN_VALUE < pred <= self.MAX_VALUE)
        if s['n_reject'] / s['sample_reject'].size > 0.5:
            reject = True
        with np.errstate(invalid="ignore"):
            output.update({
                'pred': pred, 'gt': gt, 'error': metric_error(s['mean'], gt), 'mc': s['mc'].tolist(), 'std': np.nanstd(s['mc']),
                'mean': np.nanmean(s['mc']), 'mc_nan_count': np.count_nonzero(np.isnan(s['mc'])),
                'aleatoric_std': s['aleatoric_std'], 'epistemic_std': s['epistemic_std'], 'reject': reject,
                'sample_reject': s['sample_reject'],
            })
        return output

"""Mirror of contour_uncertainty/results/clinical/patient.py (PatientMetric, Volume)."""
import numpy as np
import pandas as pd
import torch

from ... import engine as E
from ...utils.clinical import compute_left_ventricle_volumes
from .utils import ae_split, metric_error, prefix, view_sample_masks
from .view import Instant


def _has_both_views(r):
    return ('4CH' in r.keys() and '2CH' in r.keys()) or ('4C' in r.keys() and '2C' in r.keys())


class PatientMetric:
    PREFIX: str = ""

    def __init__(self):
        self.n_jobs = 0          # the reference forks multiprocessing.Pool(8) over patients (patient.py:33-46): the GPU does not need it

    def compute(self, patient_results: dict):
        raise NotImplementedError

    def __call__(self, results: dict) -> pd.DataFrame:
        """patient.py:25-66 (the n_jobs == 0 branch)."""
        res = {pid: prefix(self.compute(r), self.PREFIX) for pid, r in results.items() if _has_both_views(r)}
        return pd.DataFrame(res).T


class Volume(PatientMetric):
    """patient.py:69-225: Simpson's biplane EDV / ESV / EF of the prediction, the ground truth and of every sample index
    (A2C and A4C masks of the same index are paired), EF samples outside [0, 1] rejected, aleatoric / epistemic split."""
    PREFIX = ''
    MIN_VALUE = 0
    MAX_VALUE = 1

    def compute(self, patient_results: dict):
        try:
            a2c, a4c = patient_results['2CH'], patient_results['4CH']
        except KeyError:
            a2c, a4c = patient_results['2C'], patient_results['4C']
        i2, i4 = a2c.instants, a4c.instants
        a2c_ED, a2c_ES, a4c_ED, a4c_ES = i2[Instant.ED], i2[Instant.ES], i4[Instant.ED], i4[Instant.ES]
        lv = lambda m: np.asarray(m) == 1                                              # np.isin(mask, Label.LV)
        gt_edv, gt_esv = compute_left_ventricle_volumes(lv(a2c.gt[a2c_ED]), lv(a2c.gt[a2c_ES]), a2c.voxelspacing,
                                                        lv(a4c.gt[a4c_ED]), lv(a4c.gt[a4c_ES]), a4c.voxelspacing)
        gt_ef = (gt_edv - gt_esv) / gt_edv
        pred_edv, pred_esv = compute_left_ventricle_volumes(lv(a2c.pred[a2c_ED]), lv(a2c.pred[a2c_ES]), a2c.voxelspacing,
                                                            lv(a4c.pred[a4c_ED]), lv(a4c.pred[a4c_ES]), a4c.voxelspacing)
        pred_ef = (pred_edv - pred_esv) / pred_edv
        # Monte-Carlo volumes: every (T_e, T_a) index in one launch per view pair
        m2, m4 = view_sample_masks(a2c), view_sample_masks(a4c)                         # (N, T_e, T_a, 2048)
        te, ta = m2.shape[1], m2.shape[2]
        d2, s2 = E.simpson_diameters(torch.stack([m2[a2c_ED], m2[a2c_ES]]).reshape(2, te * ta, E.MASK_WORDS).contiguous(), tuple(a2c.voxelspacing))
        d4, s4 = E.simpson_diameters(torch.stack([m4[a4c_ED], m4[a4c_ES]]).reshape(2, te * ta, E.MASK_WORDS).contiguous(), tuple(a4c.voxelspacing))
        edv, esv, ef = E.simpson_volumes(torch.cat([d2, d4])[None].contiguous(), torch.cat([s2, s4])[None].contiguous())
        mc_edv, mc_esv, mc_ef = (t.reshape(te, ta).cpu().numpy() for t in (edv, esv, ef))
        sample_reject = np.isnan(mc_ef)                                                # the kernel NaNs exactly the EF outside [0, 1]
        s_edv, s_esv, s_ef = ae_split(mc_edv), ae_split(mc_esv), ae_split(mc_ef)
        reject = not (self.MIN_VALUE < s_ef['mean'] <= self.MAX_VALUE)
        if np.sum(sample_reject) / np.size(sample_reject) > 0.5:
            reject = True
        out = {'EF_pred': pred_ef, 'EF_gt': gt_ef, 'EF_error': metric_error(s_ef['mean'], gt_ef), 'EF_mc': mc_ef, 'EF_std': s_ef['std'],
               'EF_mean': s_ef['mean'], 'EF_aleatoric_std': s_ef['aleatoric_std'], 'EF_epistemic_std': s_ef['epistemic_std'],
               'EF_reject': reject, 'EF_sample_reject': sample_reject}
        for name, pred, gt, mc, s in (('EDV', pred_edv, gt_edv, mc_edv, s_edv), ('ESV', pred_esv, gt_esv, mc_esv, s_esv)):
            out.update({f'{name}_pred': pred, f'{name}_gt': gt, f'{name}_error': metric_error(s['mean'], gt), f'{name}_mc': mc,
                        f'{name}_std': s['std'], f'{name}_mean': s['mean'], f'{name}_aleatoric_std': s['aleatoric_std'],
                        f'{name}_epistemic_std': s['epistemic_std']})
        return out

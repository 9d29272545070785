"""Mirror of contour_uncertainty/results/clinical/instant.py (InstantMetric, AreaError)."""
from typing import List

import numpy as np
import pandas as pd

from ...data.config import BatchResult
from ...utils.clinical import lv_area
from .utils import ae_split, metric_error, prefix, sample_areas


class InstantMetric:
    PREFIX: str = ""

    def compute(self, view: BatchResult, instant_key: str, instant: int):
        raise NotImplementedError

    def __call__(self, view_results: List[BatchResult]) -> pd.DataFrame:
        """instant.py:19-25."""
        res = {}
        for view in view_results:
            for instant_key, instant in view.instants.items():
                res[f'{view.id}/{instant_key}'] = prefix(self.compute(view, instant_key, instant), self.PREFIX)
        return pd.DataFrame(res).T


class AreaError(InstantMetric):
    """instant.py:28-56.  The Monte-Carlo areas are the rasteriser's exact per-sample pixel counts."""
    PREFIX = 'Area_'

    def compute(self, view: BatchResult, instant_key: str, instant: int):
        voxelspacing = np.prod(view.voxelspacing)
        area_pred = lv_area(view.pred[instant] == 1) * voxelspacing
        area_gt = lv_area(view.gt[instant] == 1) * voxelspacing
        area_mc = sample_areas(view)[instant] * voxelspacing            # (T_e, T_a)
        s = ae_split(area_mc)
        return {
            'pred': area_pred, 'gt': area_gt, 'error': metric_error(s['mean'], area_gt), 'mc': np.asarray(area_mc).tolist(),
            'std': s['std'], 'mean': s['mean'], 'aleatoric_std': s['aleatoric_std'], 'epistemic_std': s['epistemic_std'],
        }

"""Mirror of contour_uncertainty/task/regression/dsnt/dsnt_skew.py::DSNTSkew (the predict part, no Lightning)."""
from typing import Tuple

import torch

from .... import engine as E
from .dsnt_al import DSNTAleatoric


class DSNTSkew(DSNTAleatoric):
    """dsnt_skew.py:153-199.  ``models`` return (heatmap logits, skewness (N,K,2)) -- the reference's ``skew_block`` on the
    bottleneck features stays torch; its y component is negated here as in the reference (:164)."""

    def predict_on_batch(self, img, model) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        heatmaps, alpha = model(img)
        alpha = alpha.clone()
        alpha[..., 1] = -alpha[..., 1]
        _, mu, sig = E.dsnt_moments(heatmaps.float().contiguous(), is_logit=True, covar=self.covar)
        return mu, sig, alpha

    def predict(self, img):
        S, cov, alpha = [], [], []
        for i in range(self.t_e):
            mu, sig, al = self.predict_on_batch(img, self.models[i] if self.ensembling else self.models[0])
            S.append(mu); cov.append(sig); alpha.append(al)
        return torch.stack(S).swapaxes(1, 0), torch.stack(cov).swapaxes(1, 0), torch.stack(alpha).swapaxes(1, 0)

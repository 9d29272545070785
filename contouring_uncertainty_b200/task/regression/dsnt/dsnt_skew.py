"""Mirror of contour_uncertainty/task/regression/dsnt/dsnt_skew.py::DSNTSkew (the predict part, no Lightning)."""
from typing import Tuple

import torch

from .... import engine as E
from ..aleatoric_skew import SkewUncertaintyTask


class DSNTSkew(SkewUncertaintyTask):
    """dsnt_skew.py:18-37, 153-199 (task=dsnt-skew).  ``model`` returns (heatmap logits, bottleneck features); the
    reference builds ``skew_block = model.confidence_net(2 * len(skew_indices))`` (:32) -- here it is taken from the model
    when it has one, or passed as ``skew_block=`` (any callable features -> (N, 2 K*)); both stay torch."""

    def __init__(self, covar: bool = True, mse_weight: float = 1, log_penalty_weight: float = 1, freeze_seg: bool = False,
                 *args, skew_block=None, **kwargs):
        super().__init__(*args, covar=covar, mse_weight=mse_weight, log_penalty_weight=log_penalty_weight, freeze_seg=freeze_seg,
                         **kwargs)
        if skew_block is None:
            first = self.model[0] if self.ensembling else self.model
            if hasattr(first, "confidence_net"):
                skew_block = first.confidence_net(len(self.skew_indices) * 2)
        self.skew_block = skew_block if skew_block is not None else (lambda features: features)

    def predict_on_batch(self, img, model) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """dsnt_skew.py:153-176: -> (pixel_coords, pixel_sigma, alpha (N,K,2)); the y component of the skewness is negated
        as in the reference (:164)."""
        heatmaps, features = model(img)
        a = self.skew_block(features).view((heatmaps.shape[0], len(self.skew_indices), 2))
        alpha = torch.zeros(heatmaps.shape[0], heatmaps.shape[1], 2, device=a.device, dtype=a.dtype)
        alpha[:, self.skew_indices, :] = a
        alpha[..., 1] = -alpha[..., 1]
        _, mu, sig = E.dsnt_moments(heatmaps.float().contiguous(), is_logit=True, covar=self.hparams.covar)
        return mu, sig, alpha.to(mu.device, torch.float32)

    def predict(self, img, scale=False):
        """dsnt_skew.py:178-199."""
        S, cov, alpha = [], [], []
        self.hparams.t_e = len(self.model) if self.ensembling else self.hparams.t_e
        for i in range(self.hparams.t_e):
            mu, sig, al = self.predict_on_batch(img, self.model[i] if self.ensembling else self.model)
            S.append(mu); cov.append(sig); alpha.append(al)
        return (torch.stack(S).swapaxes(1, 0).detach(), torch.stack(cov).swapaxes(1, 0).detach(),
                torch.stack(alpha).swapaxes(1, 0).detach())

"""Mirror of contour_uncertainty/task/regression/dsnt/dsnt_al.py::DSNTAleatoric (the predict part, no Lightning)."""
from typing import Sequence, Tuple

import torch

from .... import engine as E


class DSNTAleatoric:
    """dsnt_al.py:118-151.  ``models``: the U-Net(s) (cuDNN-backed torch, untouched) mapping an image batch to heatmap
    logits (N,21,256,256); one model = MC-dropout / single pass repeated ``t_e`` times, several = an ensemble."""

    def __init__(self, models, t_e: int = 1, covar: bool = True):
        self.models = list(models) if isinstance(models, (list, tuple)) else [models]
        self.ensembling = len(self.models) > 1
        self.t_e = len(self.models) if self.ensembling else t_e
        self.covar = covar

    def predict_on_batch(self, img, model) -> Tuple[torch.Tensor, torch.Tensor]:
        """dsnt_al.py:118-131: -> (pixel_coords (N,K,2), pixel_sigma (N,K,2,2)); softmax + DSNT + rescale in one kernel."""
        _, mu, sig = E.dsnt_moments(model(img).float().contiguous(), is_logit=True, covar=self.covar)
        return mu, sig

    def predict(self, img) -> Tuple[torch.Tensor, torch.Tensor]:
        """dsnt_al.py:133-151: -> S (N,T_e,K,2), cov (N,T_e,K,2,2).  Tensors stay on the device (the reference moves them
        to the CPU here); ``.cpu()`` by callers keeps working."""
        S, cov = [], []
        for i in range(self.t_e):
            mu, sig = self.predict_on_batch(img, self.models[i] if self.ensembling else self.models[0])
            S.append(mu)
            cov.append(sig)
        return torch.stack(S).swapaxes(1, 0), torch.stack(cov).swapaxes(1, 0)

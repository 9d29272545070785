"""Mirror of contour_uncertainty/task/regression/dsnt/dsnt_al.py::DSNTAleatoric (the predict part, no Lightning)."""
from typing import Tuple

import torch

from .... import engine as E
from ..aleatoric import AleatoricUncertaintyTask


class DSNTAleatoric(AleatoricUncertaintyTask):
    """dsnt_al.py:18-36, 118-151 (task=dsnt-al / dsnt-al2).  Constructor arguments as in the reference
    (``covar, mse_weight, log_penalty_weight, iterations`` here; ``psm_path, seq_psm_path, sequence_sampler`` of
    AleatoricUncertaintyTask; ``t_a, t_e, train_ensemble, ensemble_ckpt`` of UncertaintyTask) plus ``model=`` (the U-Net, or
    a list of U-Nets = an ensemble: cuDNN-backed torch, untouched) and ``data_params=``."""

    def __init__(self, covar: bool = True, mse_weight: float = 1, log_penalty_weight: float = 1, iterations: int = 1,
                 *args, **kwargs):
        super().__init__(*args, covar=covar, mse_weight=mse_weight, log_penalty_weight=log_penalty_weight, iterations=iterations,
                         **kwargs)

    def predict_on_batch(self, img, model) -> Tuple[torch.Tensor, torch.Tensor]:
        """dsnt_al.py:118-131: -> (pixel_coords (N,K,2), pixel_sigma (N,K,2,2)); softmax + DSNT + rescale in one kernel."""
        _, mu, sig = E.dsnt_moments(model(img).float().contiguous(), is_logit=True, covar=self.hparams.covar)
        return mu, sig

    def predict(self, img, scale=False) -> Tuple[torch.Tensor, torch.Tensor]:
        """dsnt_al.py:133-151: -> S (N,T_e,K,2), cov (N,T_e,K,2,2).  Tensors stay on the device (the reference moves them
        to the CPU here); ``.cpu()`` by callers keeps working."""
        S, cov = [], []
        self.hparams.t_e = len(self.model) if self.ensembling else self.hparams.t_e
        for i in range(self.hparams.t_e):
            mu, sig = self.predict_on_batch(img, self.model[i] if self.ensembling else self.model)
            S.append(mu)
            cov.append(sig)
        return torch.stack(S).swapaxes(1, 0).detach(), torch.stack(cov).swapaxes(1, 0).detach()

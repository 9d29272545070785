"""Mirror of contour_uncertainty/task/regression/aleatoric.py::AleatoricUncertaintyTask (:22-144)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Tuple

import numpy as np
import torch

from ... import engine as E
from ...data.camus.utils import USUMap
from ...data.config import ContourTags, Tags
from ...sampler.posterior_shape_model.psm import PosteriorShapeModelSampler
from ...sampler.posterior_shape_model.sequence_sampler import SequencePSMSampler
from .. import ContourUncertaintyPipeline
from .contour_uncertainty import ContourUncertaintyTask


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else (None if x is None else np.asarray(x))


class AleatoricUncertaintyTask(ContourUncertaintyTask):
    """aleatoric.py:22-49: ``psm_path`` / ``seq_psm_path`` name the shape-model files (``None`` = the files shipped with
    this package, the reference's camus-cont_psm_11_no_std.npy / camus-cont_sequence_psm_11_no_std.npy);
    ``sequence_sampler`` couples consecutive (ED, ES) images through the 84-dim sequence model."""

    def __init__(self, psm_path: str = None, seq_psm_path=None, sequence_sampler: bool = False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.hparams.psm_path, self.hparams.seq_psm_path, self.hparams.sequence_sampler = psm_path, seq_psm_path, sequence_sampler
        psm = str(Path(psm_path)) if psm_path is not None else None
        seq = str(Path(seq_psm_path)) if seq_psm_path is not None else None
        if sequence_sampler:
            self.sampler = SequencePSMSampler(psm_path=psm, sequence_psm_path=seq, device=self.device)
        else:
            self.sampler = PosteriorShapeModelSampler(psm_path=psm, device=self.device)
        self._pipe = ContourUncertaintyPipeline(psm_path=psm, t_a=self.hparams.t_a, covar=kwargs.get("covar", True),
                                                device=self.device, sequence_sampler=sequence_sampler, seq_psm_path=seq)

    def predict(self, img) -> Tuple:
        raise NotImplementedError

    def sample(self, mu: torch.Tensor, cov: torch.Tensor, t_a: int, z=None, first_instant=None, seed: int = 0):
        """aleatoric.py:54-78: mu (N,T_e,K,2), cov (N,T_e,K,2,2) -> contour samples (N,T_e,T_a,K,2) numpy.  One batched
        launch instead of the reference's per-image, per-pass sampler calls; ``z`` (N,T_e,T_a,17,2) [sequence sampler:
        (T_e,N/2,T_a,2,17,2)] replays the draws in the reference's order."""
        n, te = mu.shape[:2]
        mu, cov = mu.to(self.device, torch.float32), cov.to(self.device, torch.float32)
        if isinstance(self.sampler, SequencePSMSampler):
            out = []
            for e in range(te):                                  # (n,2,K,2) per pair -> (N,T_a,K,2)   (:66-69)
                zt = None if z is None else np.asarray(z).reshape(te, n // 2, t_a, 2, 17, 2)[e]
                fi = None if first_instant is None else np.asarray(first_instant).reshape(te, n // 2, t_a)[e]
                s = self.sampler(mu[:, e].reshape(n // 2, 2, E.K, 2), cov[:, e].reshape(n // 2, 2, E.K, 2, 2), n=t_a, z=zt,
                                 first_instant=fi, seed=seed + 7919 * e)
                out.append(s.permute(0, 2, 1, 3, 4).reshape(n, t_a, E.K, 2))
            return torch.stack(out, 1).cpu().numpy()
        zt = None if z is None else np.asarray(z).reshape(n * te, t_a, 17, 2)
        s = self.sampler(mu.reshape(n * te, E.K, 2), cov.reshape(n * te, E.K, 2, 2), n=t_a, z=zt, seed=seed)
        return s.reshape(n, te, t_a, E.K, 2).cpu().numpy()

    def _predict_step(self, batch: Any, z=None, first_instant=None, seed: int = 0):
        """aleatoric.py:80-135.  Returns (BatchResult, mu (N,K,2), cov (N,K,2,2) device tensors for predict_step)."""
        img = batch[Tags.img]
        contour = batch[ContourTags.contour] if ContourTags.contour in batch.keys() else None
        gt = _np(batch[Tags.gt]) if Tags.gt in batch.keys() else None
        mu, cov = self.predict(img)                               # (N, T_e, K, 2), (N, T_e, K, 2, 2)
        self._pipe.t_a = self.hparams.t_a
        voxelspacing, instants = self.get_voxelspacing_and_instants(batch)
        fn = getattr(self.umap_fn, "__func__", self.umap_fn)
        return self._pipe.moments_to_result(
            mu, cov, img=_np(img), seed=seed, z=z, first_instant=first_instant,
            return_pred_samples=self.hparams.return_pred_samples, id=batch[Tags.id] if Tags.id in batch.keys() else "",
            labels=self.hparams.data_params.labels, gt=gt, contour=_np(contour), instants=instants, voxelspacing=voxelspacing,
            umap_fn=None if isinstance(fn, USUMap) else fn)             # the stock functor = the batched kernel

    def get_cov_matrix(self, var_x, var_y, covar_xy=0):
        """aleatoric.py:138-144."""
        Sigma = torch.zeros((var_x.shape[0], var_x.shape[1], 2, 2), device=var_x.device)
        Sigma[:, :, 0, 0] = var_x
        Sigma[:, :, 0, 1] = covar_xy
        Sigma[:, :, 1, 0] = covar_xy
        Sigma[:, :, 1, 1] = var_y
        return Sigma

"""Mirror of contour_uncertainty/task/regression/aleatoric_skew.py::SkewUncertaintyTask (:25-127)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, List, Tuple

import numpy as np
import torch

from ... import engine as E
from ...data.camus.utils import USSkewUmap
from ...data.config import ContourTags, Tags
from ...sampler.posterior_shape_model.psm_skew import SkewPosteriorShapeModelSampler
from ...sampler.posterior_shape_model.psm_skew_sequence import SequenceSkewPSMSampler
from .. import SkewUncertaintyPipeline
from .aleatoric import AleatoricUncertaintyTask, _np


class SkewUncertaintyTask(AleatoricUncertaintyTask):
    """aleatoric_skew.py:27-46.  ``skew_indices``: landmarks that carry a skewness (None = all; the reference's other
    settings reach an undefined method, psm_skew.py:329)."""

    N_SAMPLES = 25                                                # aleatoric_skew.py:63 samples 25 contours per pass, whatever t_a says

    def __init__(self, psm_path: str = None, seq_psm_path=None, skew_indices: List[int] = None, *args, **kwargs):
        super().__init__(psm_path, seq_psm_path, *args, **kwargs)
        n_points = self.hparams.data_params.out_shape[0]
        self.skew_indices = list(range(n_points)) if skew_indices is None else skew_indices
        psm = str(Path(psm_path)) if psm_path is not None else None
        seq = str(Path(seq_psm_path)) if seq_psm_path is not None else None
        if self.hparams.sequence_sampler:
            self.sampler = SequenceSkewPSMSampler(psm_path=psm, sequence_psm_path=seq, device=self.device)
        else:
            self.sampler = SkewPosteriorShapeModelSampler(psm_path=psm, skew_indices=skew_indices, device=self.device)
        self._skew_pipe = SkewUncertaintyPipeline(psm_path=psm, t_a=self.N_SAMPLES, covar=kwargs.get("covar", True), device=self.device)

    def predict(self, img) -> Tuple:
        raise NotImplementedError

    def sample(self, mu, cov, alpha, T, z=None, expo=None, seed: int = 0):
        """aleatoric_skew.py:48-53: mu, cov, alpha (N,T_e,...) -> (N,T_e,T,K,2) numpy; one batched launch."""
        n, te = mu.shape[:2]
        to = lambda t: t.to(self.device, torch.float32).reshape((n * te,) + tuple(t.shape[2:]))
        s = self.sampler(to(mu), to(cov), to(alpha), n=T, z=z, expo=expo, seed=seed)
        return s.reshape(n, te, T, E.K, 2).cpu().numpy()

    def _predict_step(self, batch: Any, z=None, expo=None, seed: int = 0):
        """aleatoric_skew.py:55-127."""
        if self.hparams.sequence_sampler:
            raise NotImplementedError("the reference's SkewUncertaintyTask.sample cannot drive SequenceSkewPSMSampler either "
                                      "(its output is (2,n,K,2), aleatoric_skew.py:50-52); use engine.seq_skew_sample")
        img = batch[Tags.img]
        contour = batch[ContourTags.contour] if ContourTags.contour in batch.keys() else None
        gt = _np(batch[Tags.gt]) if Tags.gt in batch.keys() else None
        mu, cov, alpha = self.predict(img)
        voxelspacing, instants = self.get_voxelspacing_and_instants(batch)
        fn = getattr(self.skew_umap_fn, "__func__", self.skew_umap_fn)
        return self._skew_pipe.moments_to_result(
            mu, cov, alpha, img=_np(img), seed=seed, z=z, expo=expo, return_pred_samples=self.hparams.return_pred_samples,
            id=batch[Tags.id] if Tags.id in batch.keys() else "", labels=self.hparams.data_params.labels, gt=gt,
            contour=_np(contour), instants=instants, voxelspacing=voxelspacing,
            skew_umap_fn=None if isinstance(fn, USSkewUmap) else fn)

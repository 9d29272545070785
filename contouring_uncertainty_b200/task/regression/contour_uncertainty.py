"""Mirror of contour_uncertainty/task/regression/contour_uncertainty.py::ContourUncertaintyTask (:23-130)."""
from __future__ import annotations

from typing import Any, Tuple

import numpy as np
import torch

from ... import engine as E
from ...data.config import BatchResult, Label
from .. import add_point_and_instant_uncertainty
from ..uncertainty import UncertaintyTask


def convert_to_mask_myo(mu, image_shape, contour_samples=None):
    """contour_uncertainty.py:26-57 when Label.MYO is among the labels: mu (N,42,2) = 21 endocardial + 21 epicardial
    landmarks, contour_samples (N,T_e,T_a,42,2) -> (pred (N,H,W) int64 class map, pred_samples (N,T_e,T_a,3,H,W) int64 one-hot
    BG / LV / MYO | None), as USContourToMask gives them (cu/data/camus/utils.py:47-98)."""
    if tuple(image_shape[-2:]) != (E.HW, E.HW):
        raise ValueError("only 256x256 frames are supported")
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(dev).contiguous()

    def one_hot(lm):
        lv, myo, _ = E.us_myo_masks(lm.reshape(-1, 2 * E.K, 2))
        lv, myo = E.unpack_masks(lv).to(torch.int64), E.unpack_masks(myo).to(torch.int64)
        return torch.stack([((lv + myo) == 0).to(torch.int64), lv, myo], dim=1)           # (M, 3, H, W)

    n = image_shape[0]
    pred = one_hot(t(mu).reshape(n, 2 * E.K, 2)).argmax(1).cpu().numpy()
    pred_samples = None
    if contour_samples is not None:
        cs = t(contour_samples)
        pred_samples = one_hot(cs).cpu().numpy().reshape(tuple(cs.shape[:3]) + (3, E.HW, E.HW))
    return pred, pred_samples


def convert_to_mask(mu, image_shape, contour_samples=None, soft_mask: bool = False):
    """contour_uncertainty.py:26-57 for LV labels: mu (N,K,2), contour_samples (N,T_e,T_a,K,2) ->
    (pred (N,H,W) int64, pred_samples (N,T_e,T_a,1,H,W) int64 [fp64 soft masks with ``soft_mask``] | None).  One rasteriser
    launch for everything."""
    if tuple(image_shape[-2:]) != (E.HW, E.HW):
        raise ValueError("only 256x256 frames are supported")
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(dev).contiguous()
    n = image_shape[0]
    pm, _ = E.rasterize_contours(t(mu).reshape(n, E.K, 2), round_landmarks=True)
    pred = E.unpack_masks(pm).cpu().numpy().astype(np.int64)
    pred_samples = None
    if contour_samples is not None:
        cs = t(contour_samples)
        sm, _ = E.rasterize_contours(cs.reshape(-1, E.K, 2), round_landmarks=True)
        if soft_mask:                                         # :44-49: blurred, min-max normalised float masks
            pred_samples = E.soft_masks(sm).cpu().numpy().reshape(tuple(cs.shape[:3]) + (1, E.HW, E.HW))
        else:
            pred_samples = E.unpack_masks(sm).cpu().numpy().astype(np.int64).reshape(tuple(cs.shape[:3]) + (1, E.HW, E.HW))
    return pred, pred_samples


class ContourUncertaintyTask(UncertaintyTask):
    """contour_uncertainty.py:23-130.  ``return_pred_samples`` (extension, default True = the reference's behaviour):
    materialise ``BatchResult.pred_samples`` (N,T_e,T_a,H,W) int64; the maps never need it here (they come from the
    on-device count map), so large sample counts can switch it off."""

    def __init__(self, *args, return_pred_samples: bool = True, **kwargs):
        super().__init__(*args, **kwargs)
        self.hparams.return_pred_samples = return_pred_samples
        self.umap_fn = self.contour_to_mask_fn = self.skew_umap_fn = None

    def convert_to_mask(self, mu: np.ndarray, image_shape: Tuple, contour_samples=None, soft_mask: bool = False):
        """contour_uncertainty.py:26-57."""
        if Label.MYO in tuple(self.hparams.data_params.labels):
            if soft_mask:
                raise NotImplementedError("soft_mask is not provided")
            return convert_to_mask_myo(mu, image_shape, contour_samples)
        return convert_to_mask(mu, image_shape, contour_samples, soft_mask)

    def _fetch_functors(self):
        from ...data.camus.datamodule import CamusContourDataModule

        dm = getattr(self.trainer, "datamodule", None) if self.trainer is not None else None
        dm = dm if dm is not None else CamusContourDataModule
        self.umap_fn = staticmethod(dm.umap_fn)
        self.contour_to_mask_fn = staticmethod(dm.contour_to_mask_fn)
        self.skew_umap_fn = staticmethod(dm.skew_umap_fn)

    def on_predict_start(self):
        """contour_uncertainty.py:59-63: the functors come from the datamodule (``self.trainer.datamodule``; without a
        trainer: ``CamusContourDataModule``'s)."""
        self._fetch_functors()

    def on_fit_start(self):
        self._fetch_functors()

    def predict_step(self, batch: Any, batch_idx: int = 0, dataloader_idx: int = 0, **replay) -> BatchResult:
        """contour_uncertainty.py:71-130.  ``replay``: optional recorded draws handed to ``_predict_step`` (``z``,
        ``expo``, ``first_instant``, ``seed``) -- the reference draws from torch's / Python's global generators."""
        if self.umap_fn is None:
            self.on_predict_start()
        res, mu_t, cov_t = self._predict_step(batch, **replay)
        n = res.img.shape[0]
        if res.pred_samples is not None:
            if res.entropy_map is None:                             # :76-77 (here the entropy map comes from the count map)
                res.entropy_map = np.array([self.sample_entropy(res.pred_samples[i].astype(float)) for i in range(n)])
            if len(self.hparams.data_params.labels) <= 2:            # :79-82
                res.pred_samples = res.pred_samples.squeeze(3) if res.pred_samples.ndim == 6 else res.pred_samples
            else:
                res.pred_samples = res.pred_samples.argmax(3)
        return add_point_and_instant_uncertainty(res, mu_t, cov_t)     # :84-128

    def _predict_step(self, batch: Any, **replay):
        raise NotImplementedError

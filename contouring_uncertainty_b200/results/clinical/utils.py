"""Mirror of contour_uncertainty/results/clinical/utils.py + the device route the metric processors use."""
from typing import Optional, Sequence

import numpy as np
import torch

from ... import engine as E
from ...utils.clinical import aleatoric_epistemic_uncertainty  # noqa: F401  (results/clinical/utils.py:3-18, numpy, 4 scalars)


def prefix(map, prefix: str, exclude=None):
    """vital/vital/utils/format/native.py:5-24."""
    exclude = [] if exclude is None else [exclude] if isinstance(exclude, str) else exclude
    return {(k if k in exclude else f"{prefix}{k}"): v for k, v in map.items()}


def metric_error(prediction, gt, type: str = "absolute"):
    """cu/utils/clinical.py:96-110."""
    error = np.abs(prediction - gt)
    if type == "relative":
        error /= gt
    return error


def ae_split(mc: np.ndarray, value_range: Optional[Sequence[float]] = None):
    """(T_e, T_a) or (B, T_e, T_a) Monte-Carlo values -> dict(mean, aleatoric_std, epistemic_std, std [, mc, sample_reject,
    n_reject]) through ``cub200_metric_ae_split`` (one launch for the whole batch)."""
    a = np.asarray(mc, dtype=np.float64)
    single = a.ndim == 2
    a3 = a[None] if single else a
    dev = torch.device("cuda", torch.cuda.current_device())
    r = E.metric_ae_split(torch.from_numpy(np.ascontiguousarray(a3)).to(dev), value_range, return_values=True)
    stats = r["stats"].cpu().numpy()
    out = dict(mean=stats[:, 0], aleatoric_std=stats[:, 1], epistemic_std=stats[:, 2], std=stats[:, 3],
               mc=r["values"].cpu().numpy(), sample_reject=r["reject"].cpu().numpy().astype(bool), n_reject=r["n_reject"].cpu().numpy())
    return {k: v[0] for k, v in out.items()} if single else out


def sample_areas(view) -> np.ndarray:
    """Per-sample LV pixel counts (N, T_e, T_a) of a BatchResult: the exact counts the rasteriser wrote (``sample_area``), or
    -- for a BatchResult that was not produced by this package -- lv_area() of its pred_samples."""
    if getattr(view, "sample_area", None) is not None:
        return np.asarray(view.sample_area)
    ps = np.asarray(view.pred_samples)
    return (ps == 1).sum((-2, -1))


def packed_masks(masks: np.ndarray) -> torch.Tensor:
    """(..., 256, 256) 0/1 (any dtype; LV = label 1) -> (..., 2048) int32 bit-packed on the device."""
    b = np.packbits(np.asarray(masks) == 1, axis=-1, bitorder="little")
    w = np.ascontiguousarray(b).view(np.uint32).view(np.int32).reshape(b.shape[:-2] + (E.MASK_WORDS,))
    return torch.from_numpy(w).to(torch.device("cuda", torch.cuda.current_device()))


def view_sample_masks(view) -> torch.Tensor:
    """Bit-packed sample masks (N, T_e, T_a, 2048) of a BatchResult (``packed_samples`` when the task attached them)."""
    if getattr(view, "packed_samples", None) is not None:
        p = view.packed_samples
        return p if isinstance(p, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p)).to(torch.device("cuda", torch.cuda.current_device()))
    return packed_masks(view.pred_samples)

"""Mirror of contour_uncertainty/data/lung/utils.py: LungContourToMask (three structures of 44 / 50 / 26 points, each
rasterised by reconstruction() / linear_reconstruction() on the GPU, the generic-size rasteriser) and LungUMap /
LungSkewUmap (skew_umap per structure, each map divided by its maximum, summed and clipped to [0, 1])."""
from typing import Literal

import numpy as np
import torch

from ... import engine as E
from ...utils.skew_umap import skew_umap
from ..utils import ContourToMask, SkewUMap, UMap
from .config import HEART, LLUNG, RLUNG, Label


def split_landmarks(landmarks):
    """lung/utils.py:13-22 on the second-to-last axis."""
    p1, p2, p3 = RLUNG, RLUNG + LLUNG, RLUNG + LLUNG + HEART
    return landmarks[..., :p1, :], landmarks[..., p1:p2, :], landmarks[..., p2:p3, :]


class LungContourToMask(ContourToMask):
    @staticmethod
    def __call__(landmarks, shape=(256, 256), labels=None, apply_argmax: bool = True,
                 reconstruction_type: Literal['spline', 'linear'] = 'spline'):
        """landmarks (120, 2) [or (M, 120, 2)] -> argmax map (H, W) int64, or the (3, H, W) BG / LUNG / HEART stack when
        not apply_argmax (lung/utils.py:25-69).  The two lungs ADD into the LUNG channel, as in the reference."""
        if reconstruction_type not in ('spline', 'linear'):
            raise ValueError
        if tuple(shape) != (E.HW, E.HW):
            raise ValueError("only 256x256 frames are supported")
        lm = np.asarray(landmarks, dtype=np.float64).squeeze()
        single = lm.ndim == 2
        lm = np.round(lm.reshape(-1, RLUNG + LLUNG + HEART, 2))          # landmarks.round().astype(int) (:38)
        dev = torch.device("cuda", torch.cuda.current_device())
        parts = []
        for pts in split_landmarks(lm):
            masks, _ = E.rasterize_contours_generic(torch.from_numpy(np.ascontiguousarray(pts)).to(dev), round_landmarks=True,
                                                    reconstruction_type=reconstruction_type)
            parts.append(E.unpack_masks(masks).to(torch.int64))
        rl, ll, h = parts
        seg = torch.zeros((lm.shape[0], 3, E.HW, E.HW), dtype=torch.int64, device=dev)
        seg[:, Label.LUNG] = rl + ll
        seg[:, Label.HEART] = h
        seg[:, Label.BG] = (seg.sum(1) == 0).to(torch.int64)
        out = (seg.argmax(1) if apply_argmax else seg).cpu().numpy()
        return out[0] if single else out


def _structure_umaps(mu, cov, alpha):
    """skew_umap of the three structures (lung/utils.py:75-81, 96-102): [(projected_mode, umap)] * 3."""
    mu, cov, alpha = (np.asarray(a) for a in (mu, cov, alpha))
    return [skew_umap(m, c, a) for m, c, a in zip(split_landmarks(mu), _split_cov(cov), split_landmarks(alpha))]


def _split_cov(cov):
    p1, p2, p3 = RLUNG, RLUNG + LLUNG, RLUNG + LLUNG + HEART
    return cov[:p1], cov[p1:p2], cov[p2:p3]


def _combine(umaps):
    """np.clip(rl / rl.max() + ll / ll.max() + h / h.max(), 0, 1) (lung/utils.py:83, 104)."""
    return np.clip(umaps[0] / umaps[0].max() + umaps[1] / umaps[1].max() + umaps[2] / umaps[2].max(), a_min=0, a_max=1)


class LungUMap(UMap):
    @staticmethod
    def __call__(mu, cov, labels=None):
        """lung/utils.py:72-92: mu (120, 2), cov (120, 2, 2) -> (256, 256) float64."""
        mu = np.asarray(mu)
        return _combine([u for _, u in _structure_umaps(mu, cov, np.zeros_like(mu))])


class LungSkewUmap(SkewUMap):
    @staticmethod
    def __call__(mu, cov, alpha, labels=None):
        """lung/utils.py:95-114: -> (projected_mode (120, 2), umap (256, 256) float64)."""
        parts = _structure_umaps(mu, cov, alpha)
        return np.concatenate([m for m, _ in parts], axis=0), _combine([u for _, u in parts])

"""Mirror of contour_uncertainty/data/lung/utils.py::LungContourToMask: three structures of 44 / 50 / 26 points, each
rasterised by reconstruction() / linear_reconstruction() on the GPU (the generic-size rasteriser)."""
from typing import Literal

import numpy as np
import torch

from ... import engine as E
from ..utils import ContourToMask
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

"""Mirror of contour_uncertainty/data/camus/utils.py::USContourToMask (LV-only branch :43, :85-90; LV + MYO branch :47-82)."""
from typing import Literal

import numpy as np
import torch

from ... import engine as E
from ..utils import ContourToMask, UMap, SkewUMap


class USContourToMask(ContourToMask):
    @staticmethod
    def __call__(landmarks, shape=(256, 256), labels=None, apply_argmax: bool = True,
                 reconstruction_type: Literal['spline', 'linear'] = 'spline'):
        """landmarks (K,2) [or (M,K,2)] -> int64 mask (H,W) [(M,H,W)]; channel axis added when not apply_argmax.

        Rounds the landmarks first, then spline raster + closing line + hole fill (cu/utils/contour.py:28-40), or,
        with reconstruction_type='linear', the polygon outline + hole fill (cu/utils/contour.py:43-53).
        """
        if reconstruction_type not in ('spline', 'linear'):
            raise ValueError
        if tuple(shape) != (E.HW, E.HW):
            raise ValueError("only 256x256 frames are supported")
        if labels is not None and any(str(getattr(l, "name", l)).lower().endswith("myo") or l == 2 for l in labels):
            # two structures (:47-82): 42 landmarks = endocardium + epicardium -> BG / LV / MYO
            lm = np.asarray(landmarks, dtype=np.float32).squeeze()
            single = lm.ndim == 2
            lm = lm.reshape(-1, 2 * E.K, 2)
            dev = torch.device("cuda", torch.cuda.current_device())
            lv, myo, _ = E.us_myo_masks(torch.from_numpy(np.ascontiguousarray(lm)).to(dev), reconstruction_type=reconstruction_type)
            lv, myo = E.unpack_masks(lv).to(torch.int64), E.unpack_masks(myo).to(torch.int64)
            seg = torch.stack([((lv + myo) == 0).to(torch.int64), lv, myo], dim=1)       # Label.BG, Label.LV, Label.MYO
            out = (seg.argmax(1) if apply_argmax else seg).cpu().numpy()
            return out[0] if single else out
        lm = np.asarray(landmarks, dtype=np.float32).squeeze()
        single = lm.ndim == 2
        lm = lm.reshape(-1, E.K, 2)
        dev = torch.device("cuda", torch.cuda.current_device())
        masks, _ = E.rasterize_contours(torch.from_numpy(lm).to(dev), round_landmarks=True, reconstruction_type=reconstruction_type)
        out = E.unpack_masks(masks).cpu().numpy().astype(np.int64)
        if single:
            out = out[0]
            return out if apply_argmax else out[None]
        return out if apply_argmax else out[:, None]


class USUMap(UMap):
    """cu/data/camus/utils.py:101-124 (LV-only labels): uncertainty_map(mu, cov) / max -> (256,256) float64."""

    @staticmethod
    def __call__(mu, cov, labels=None):
        if labels is not None and any(str(l).lower().endswith("myo") for l in labels):
            raise NotImplementedError("the MYO (two-structure) branch is out of scope")
        dev = torch.device("cuda", torch.cuda.current_device())
        m = torch.as_tensor(np.asarray(mu, dtype=np.float32)).reshape(-1, E.K, 2).to(dev)
        c = torch.as_tensor(np.asarray(cov, dtype=np.float32)).reshape(-1, E.K, 2, 2).to(dev)
        out = E.umap_analytic(m, c, normalize=True).cpu().numpy()
        return out[0] if np.asarray(mu).ndim == 2 else out


class USSkewUmap(SkewUMap):
    """cu/data/camus/utils.py:127-153 (LV-only labels): skew_umap(mu, cov, alpha, linear_close=True), umap / max
    -> (projected_mode (21,2) float32, umap (256,256) float64)."""

    @staticmethod
    def __call__(mu, cov, alpha, labels=None):
        if labels is not None and any(str(l).lower().endswith("myo") for l in labels):
            raise NotImplementedError("the MYO (two-structure) branch is out of scope")
        dev = torch.device("cuda", torch.cuda.current_device())
        m = torch.as_tensor(np.asarray(mu, dtype=np.float32)).reshape(-1, E.K, 2).to(dev)
        c = torch.as_tensor(np.asarray(cov, dtype=np.float32)).reshape(-1, E.K, 2, 2).to(dev)
        a = torch.as_tensor(np.asarray(alpha, dtype=np.float32)).reshape(-1, E.K, 2).to(dev)
        mode, out = E.skew_umap(m, c, a, linear_close=True, normalize=True)
        mode, out = mode.cpu().numpy(), out.cpu().numpy()
        return (mode[0], out[0]) if np.asarray(mu).ndim == 2 else (mode, out)

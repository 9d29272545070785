"""Mirror of contour_uncertainty/task/regression/contour_uncertainty.py::ContourUncertaintyTask.convert_to_mask."""
import numpy as np
import torch

from ... import engine as E


def convert_to_mask(mu, image_shape, contour_samples=None, soft_mask: bool = False):
    """contour_uncertainty.py:26-57 for LV labels: mu (N,K,2), contour_samples (N,T_e,T_a,K,2) ->
    (pred (N,H,W) int64, pred_samples (N,T_e,T_a,1,H,W) int64 | None).  One rasteriser launch for everything."""
    if soft_mask:
        raise NotImplementedError("soft_mask (skimage.filters.gaussian) is off by default and not provided")
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
        pred_samples = E.unpack_masks(sm).cpu().numpy().astype(np.int64).reshape(tuple(cs.shape[:3]) + (1, E.HW, E.HW))
    return pred, pred_samples

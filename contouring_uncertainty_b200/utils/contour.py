"""Mirror of contour_uncertainty/utils/contour.py::reconstruction (numpy in / numpy out, rasterised on the GPU)."""
import numpy as np
import torch

from .. import engine as E


def reconstruction(points, height, width):
    """contour.py:28-40: spline raster of the UNROUNDED points + closing line + hole fill -> (256,256) int mask.
    Any number of points from 4 to 64 (the 21-point LV contour takes the hot-path rasteriser)."""
    if (height, width) != (E.HW, E.HW):
        raise NotImplementedError("256 x 256 frames only")
    dev = torch.device("cuda", torch.cuda.current_device())
    pts = np.asarray(points, dtype=np.float64)
    if pts.ndim != 2 or pts.shape[1] != 2:
        raise ValueError("points must be (K, 2)")
    c = torch.as_tensor(pts)[None].to(dev)
    if len(pts) == E.K:
        masks, _ = E.rasterize_contours(c, round_landmarks=False)
    else:
        masks, _ = E.rasterize_contours_generic(c, round_landmarks=False)
    return E.unpack_masks(masks)[0].cpu().numpy().astype(int)


def linear_reconstruction(points, height, width):
    """contour.py:43-53: the rounded points joined by skimage.draw.line segments + hole fill -> (256,256) int mask."""
    if (height, width) != (E.HW, E.HW):
        raise NotImplementedError("256 x 256 frames only")
    dev = torch.device("cuda", torch.cuda.current_device())
    pts = np.asarray(points, dtype=np.float64)
    if pts.ndim != 2 or pts.shape[1] != 2:
        raise ValueError("points must be (K, 2)")
    c = torch.as_tensor(pts)[None].to(dev)
    if len(pts) == E.K:
        masks, _ = E.rasterize_contours(c, reconstruction_type="linear")
    else:
        masks, _ = E.rasterize_contours_generic(c, reconstruction_type="linear")
    return E.unpack_masks(masks)[0].cpu().numpy().astype(int)

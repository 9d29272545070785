"""Mirror of contour_uncertainty/utils/contour.py::reconstruction (numpy in / numpy out, rasterised on the GPU)."""
import numpy as np
import torch

from .. import engine as E


def reconstruction(points, height, width):
    """contour.py:28-40: spline raster of the UNROUNDED points + closing line + hole fill -> (256,256) int mask."""
    if (height, width) != (E.HW, E.HW):
        raise NotImplementedError("256 x 256 frames only")
    dev = torch.device("cuda", torch.cuda.current_device())
    c = torch.as_tensor(np.asarray(points, dtype=np.float64)).reshape(1, E.K, 2).to(dev)
    masks, _ = E.rasterize_contours(c, round_landmarks=False)
    return E.unpack_masks(masks)[0].cpu().numpy().astype(int)


def linear_reconstruction(points, height, width):
    """contour.py:43-53: the rounded points joined by skimage.draw.line segments + hole fill -> (256,256) int mask."""
    if (height, width) != (E.HW, E.HW):
        raise NotImplementedError("256 x 256 frames only")
    dev = torch.device("cuda", torch.cuda.current_device())
    c = torch.as_tensor(np.asarray(points, dtype=np.float64)).reshape(1, E.K, 2).to(dev)
    masks, _ = E.rasterize_contours(c, reconstruction_type="linear")
    return E.unpack_masks(masks)[0].cpu().numpy().astype(int)

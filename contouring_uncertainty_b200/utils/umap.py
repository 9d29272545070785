"""Mirror of contour_uncertainty/utils/umap.py::uncertainty_map (numpy in / numpy out, computed on the GPU)."""
import numpy as np
import torch

from .. import engine as E


def uncertainty_map(mu_p, cov_p, shape=(256, 256), close=True):
    """umap.py:10-33: (256,256) float64 map, un-normalised (USUMap divides by its maximum)."""
    if tuple(shape) != (E.HW, E.HW) or not close:
        raise NotImplementedError("256 x 256 frames with the closing line (the reference's only use)")
    dev = torch.device("cuda", torch.cuda.current_device())
    m = torch.as_tensor(np.asarray(mu_p, dtype=np.float32)).reshape(1, E.K, 2).to(dev)
    c = torch.as_tensor(np.asarray(cov_p, dtype=np.float32)).reshape(1, E.K, 2, 2).to(dev)
    return E.umap_analytic(m, c, normalize=False)[0].cpu().numpy()

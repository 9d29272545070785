"""Mirror of contour_uncertainty/utils/skew_umap.py::skew_umap (numpy in / numpy out, computed on the GPU)."""
import numpy as np
import torch

from .. import engine as E


def skew_umap(mu, cov, alpha, shape=(256, 256), close=True, linear_close=False):
    """skew_umap.py:11-82: (projected_mode (K,2) float32, umap (256,256) float64, natural-log entropy, un-normalised)."""
    if tuple(shape) != (E.HW, E.HW):
        raise NotImplementedError("256 x 256 frames only")
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a, s: torch.as_tensor(np.asarray(a, dtype=np.float32)).reshape(s).to(dev)
    k = int(np.asarray(mu).shape[-2])                                 # 21 LV landmarks, or any 4..64 (lung / heart contours)
    mode, um = E.skew_umap(t(mu, (1, k, 2)), t(cov, (1, k, 2, 2)), t(alpha, (1, k, 2)), linear_close=linear_close, normalize=False)
    return mode[0].cpu().numpy(), um[0].cpu().numpy()

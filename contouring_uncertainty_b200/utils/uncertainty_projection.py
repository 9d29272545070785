"""Mirror of contour_uncertainty/utils/uncertainty_projection.py (numpy in / numpy out, computed on the GPU)."""
import numpy as np
import torch

from .. import engine as E


def _t(a, shape):
    dev = torch.device("cuda", torch.cuda.current_device())
    return torch.as_tensor(np.asarray(a, dtype=np.float32)).reshape(shape).to(dev)


def projected_uncertainty(mu, cov, alpha=None, use_eigenvalue: bool = True, all=False, linear_close=False):
    """uncertainty_projection.py:17-125: (uncertainties (K,) fp32, projections (K,2) fp64 [, alpha_proj (K,)])."""
    if not use_eigenvalue:
        raise NotImplementedError("use_eigenvalue=False is never used by the reference's callers")
    r = E.projected_uncertainty(_t(mu, (1, E.K, 2)), _t(cov, (1, E.K, 2, 2)), None if alpha is None else _t(alpha, (1, E.K, 2)),
                                all=all, linear_close=linear_close)
    return tuple(x[0].cpu().numpy() for x in r)


def projected_uncertainty_value(mu, cov, use_eigenvalue: bool = True):
    """uncertainty_projection.py:11-14."""
    return np.sum(projected_uncertainty(mu, cov, use_eigenvalue=use_eigenvalue)[0])

"""Mirror of contour_uncertainty/sampler/posterior_shape_model/psm_skew.py::SkewPosteriorShapeModelSampler."""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional

import torch

from ... import engine as E
from .psm import PosteriorShapeModelSampler


class SkewPosteriorShapeModelSampler(PosteriorShapeModelSampler):
    """psm_skew.py:162-208.  ``__call__(mu (N,K,2), cov (N,K,2,2), alpha (N,K,2), n=1) -> (N,n,K,2)``.

    Initial points by ``BivariateSkewNormal.rvs_fast``; every level point by a multinomial draw over
    the 256x256 product pdf (skew pdf of the prediction x shape-model conditional Gaussian), so level
    points are integer grid coordinates; the last four points are the posterior mean.  ``z`` (N,n,3,3)
    and ``expo`` (N,n,14,65536) replay the reference's random draws (rvs_fast normals and the Exp(1)
    field of each ``torch.multinomial``); otherwise Philox with ``seed``.  Only ``skew_indices=None``
    (all points) is supported -- the reference's other settings hit an undefined method (psm_skew.py:329).
    A 2-D ``mu`` is treated as one image (the reference's own 2-D branch is broken, psm_skew.py:199).
    """

    def __init__(self, psm_path: Optional[Path] = None, levels: int = 3, skew_indices: Optional[List[int]] = None, device=None):
        super().__init__(psm_path, levels, device)
        if skew_indices is not None and sorted(skew_indices) != list(range(E.K)):
            raise NotImplementedError("skew_indices must be None (all points)")
        self.skew_indices = list(range(E.K))
        self.grid_size = E.HW

    def __call__(self, mu: torch.Tensor, cov: torch.Tensor, alpha: torch.Tensor, n: int = 1, debug_img=None,
                 progress_bar: bool = False, z=None, expo=None, seed: int = 0, return_indices: bool = False):
        out_dev = mu.device
        to = lambda t, shp: torch.as_tensor(t, dtype=torch.float32).to(self.device).reshape(shp).contiguous()
        m, c, a = to(mu, (-1, E.K, 2)), to(cov, (-1, E.K, 2, 2)), to(alpha, (-1, E.K, 2))
        N = m.shape[0]
        zt = None if z is None else to(z, (N, n, 3, 3))
        et = None if expo is None else to(expo, (N, n, E.N_DRAW_SKEW, E.HW * E.HW))
        r = E.skew_sample(self.model, m, c, a, n, z=zt, expo=et, seed=seed, return_indices=return_indices)
        if return_indices:
            return r[0].to(out_dev), r[1].to(out_dev)
        return r.to(out_dev)

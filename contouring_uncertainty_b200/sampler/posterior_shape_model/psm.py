"""Mirror of contour_uncertainty/sampler/posterior_shape_model/psm.py::PosteriorShapeModelSampler."""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import torch

from ... import engine as E
from ..sampler import Sampler


class PosteriorShapeModelSampler(Sampler):
    """psm.py:23-93.  ``__call__(mu (K,2), cov (K,2,2), alpha=None, n=1) -> (n,K,2)`` torch fp32.

    Stateless per call (the reference overwrites ``self.mu / self.Q`` on every call and is not
    thread-safe).  Batched input (N,K,2) gives (N,n,K,2).  ``z`` (n,17,2) / (N,n,17,2) replays
    the standard-normal draws in the reference's order; otherwise Philox with ``seed``.
    Output lives on the input's device (CPU in -> CPU out), the computation is always on the GPU.
    """

    def __init__(self, psm_path: Optional[Path] = None, levels: int = 3, device=None):
        if levels != 3:
            raise NotImplementedError("levels=3 (the reference default and only configured value)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.model = E.ShapeModelConstants(str(psm_path) if psm_path is not None else E.DEFAULT_PSM, self.device)
        self.initial_points, self.points_order = self.get_points_order(E.K, levels=levels)

    def factor(self, mu: torch.Tensor, cov: torch.Tensor) -> torch.Tensor:
        return E.psm_factor(self.model, mu, cov)

    def __call__(self, mu: torch.Tensor, cov: torch.Tensor, alpha=None, n: int = 1, debug_img=None, z=None,
                 seed: int = 0) -> torch.Tensor:
        if alpha is not None:
            raise NotImplementedError("alpha (skew-normal initial points) belongs to the dsnt-skew row")
        out_dev = mu.device
        batched = mu.dim() == 3
        m = mu.to(self.device, torch.float32).reshape(-1, E.K, 2).contiguous()
        c = cov.to(self.device, torch.float32).reshape(-1, E.K, 2, 2).contiguous()
        zt = None
        if z is not None:
            zt = torch.as_tensor(z, dtype=torch.float32).to(self.device).reshape(m.shape[0], n, E.Z_PSM).contiguous()
        a_t = self.factor(m, c)
        r = E.sample_raster_aggregate(m, a_t, n, z=zt, seed=seed)
        s = r["contours"]
        return (s if batched else s[0]).to(out_dev)

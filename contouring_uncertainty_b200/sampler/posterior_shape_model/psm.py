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

    def _call_with_alpha(self, m, c, alpha, n, z, seed):
        """psm.py:231-237: with ``alpha`` the three initial points (0, 10, 20) are drawn from the bivariate skew-normal
        ``rvs_fast(mu_j, cov_j, alpha_j * [1, -1])``; everything after them is the Gaussian hierarchy.  The sampler is affine in
        the initial points' deviations, so the skew draws are turned into the equivalent first six normals
        z_k = chol(cov_k)^-1 (s_k - mu_k) and the rest runs through the same factor.  ``z`` = (z3 (N,n,3,3), z2 (N,n,14,2))
        replays the draws (rvs_fast normals, then the 14 level draws); otherwise Philox / torch's device generator."""
        N = m.shape[0]
        a = torch.as_tensor(alpha, dtype=torch.float32).to(self.device).reshape(N, E.K, 2) * torch.tensor([1.0, -1.0], device=self.device)
        z3 = z2 = None
        if z is not None:
            z3 = torch.as_tensor(z[0], dtype=torch.float32).to(self.device).reshape(N, n, 3, 3)
            z2 = torch.as_tensor(z[1], dtype=torch.float32).to(self.device).reshape(N, n, 28)
        zfull = torch.empty((N, n, E.Z_PSM), dtype=torch.float32, device=self.device)
        if z2 is None:
            g = torch.Generator(device=self.device).manual_seed(int(seed) + 0x5EED)
            zfull[:, :, 6:] = torch.randn((N, n, 28), generator=g, device=self.device)
        else:
            zfull[:, :, 6:] = z2
        for i in range(N):
            for d, k in enumerate(self.initial_points):
                s = E.skew_rvs_fast(m[i, k].contiguous(), c[i, k].contiguous(), a[i, k].contiguous(), n,
                                    z=None if z3 is None else z3[i, :, d].contiguous(), seed=seed + 7919 * (3 * i + d))
                l00 = torch.sqrt(c[i, k, 0, 0]); l10 = c[i, k, 1, 0] / l00; l11 = torch.sqrt(c[i, k, 1, 1] - l10 * l10)
                dx, dy = s[:, 0] - m[i, k, 0], s[:, 1] - m[i, k, 1]
                zfull[i, :, 2 * d] = dx / l00
                zfull[i, :, 2 * d + 1] = (dy - l10 * (dx / l00)) / l11
        return E.sample_contours(m, self.factor(m, c), n, z=zfull.contiguous())

    def __call__(self, mu: torch.Tensor, cov: torch.Tensor, alpha=None, n: int = 1, debug_img=None, z=None,
                 seed: int = 0) -> torch.Tensor:
        if alpha is not None:
            out_dev = mu.device
            batched = mu.dim() == 3
            m = mu.to(self.device, torch.float32).reshape(-1, E.K, 2).contiguous()
            c = cov.to(self.device, torch.float32).reshape(-1, E.K, 2, 2).contiguous()
            s = self._call_with_alpha(m, c, alpha, n, z, seed)
            return (s if batched else s[0]).to(out_dev)
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

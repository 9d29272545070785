"""Mirror of contour_uncertainty/sampler/sampler.py + naive.py (independent per-point Gaussians)."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

from .. import engine as E


def _device(dev=None):
    return torch.device(dev) if dev is not None else torch.device("cuda", torch.cuda.current_device())


class Sampler:
    """sampler.py:6-41.  ``__call__(mu (K,2), cov (K,2,2), n) -> (n,K,2)``."""

    def __call__(self, mu, cov, n: int):
        raise NotImplementedError

    @staticmethod
    def get_points_order(nb_points: int = 21, nb_initial_points: int = 3, levels: Optional[int] = None):
        """psm.py:43-71."""
        initial_points = np.round(np.linspace(0, nb_points - 1, nb_initial_points)).astype(int).tolist()
        levels = levels or int(math.log(nb_points, 2))
        all_points, point_order = list(initial_points), []
        for _ in range(levels):
            level_points = []
            for j in range(len(all_points) - 1):
                if all_points[j] + 1 != all_points[j + 1]:
                    point = (all_points[j] + all_points[j + 1]) / 2
                    point = math.ceil(point) if point > nb_points / 2 else math.floor(point)
                    level_points.append(int(point))
            if not level_points:
                break
            all_points.extend(level_points)
            all_points.sort()
            point_order.append(level_points)
        return initial_points, point_order


class NaiveSampler(Sampler):
    """naive.py:7-25: K independent bivariate normals.  numpy in -> numpy (n,K,2) float64 out.

    Production draws come from the in-kernel Philox generator (``seed``); ``z`` (n,K,2) replays
    given standard-normal draws, and ``colouring`` (K,2,2) overrides the Cholesky factor (numpy's
    ``multivariate_normal`` colours with an SVD factor: x = z @ colouring[k]).
    """

    def __init__(self, sample_indices: Optional[Sequence] = None, device=None):
        if sample_indices is not None and list(sample_indices) != list(range(E.K)):
            raise NotImplementedError("only the full 21-landmark contour is supported")
        self.device = device

    def __call__(self, mu, cov, n: int, z=None, colouring=None, seed: int = 0) -> np.ndarray:
        dev = _device(self.device)
        mu_t = torch.as_tensor(np.asarray(mu), dtype=torch.float32, device=dev).reshape(1, E.K, 2)
        cov_t = torch.as_tensor(np.asarray(cov), dtype=torch.float32, device=dev).reshape(1, E.K, 2, 2)
        if colouring is None:
            a_t = E.naive_factor(cov_t)
        else:
            a = torch.zeros((1, 42, 42), dtype=torch.float32, device=dev)
            m = torch.as_tensor(np.asarray(colouring), dtype=torch.float32, device=dev)
            for k in range(E.K):
                a[0, 2 * k:2 * k + 2, 2 * k:2 * k + 2] = m[k]
            a_t = a
        zt = None
        if z is not None:
            zt = torch.as_tensor(np.asarray(z), dtype=torch.float32, device=dev).reshape(1, n, 42)
        r = E.sample_raster_aggregate(mu_t, a_t, n, z=zt, seed=seed)
        return r["contours"][0].double().cpu().numpy()

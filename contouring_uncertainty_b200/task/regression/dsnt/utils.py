"""Mirror of contour_uncertainty/task/regression/dsnt/utils.py (same names and argument meaning).

``dsnt`` and ``dsnt_from_logits`` run the fused sm_100a kernel (K1); ``flat_softmax`` is kept for
callers that need the normalised heatmaps themselves (plots, training) and is plain torch -- it is
not on the predict path any more (the kernel fuses the softmax).
"""
from functools import reduce
from operator import mul

import torch

from ....engine import dsnt_moments


def normalized_linspace(length, dtype=None, device=None):
    """dsnt/utils.py:50-68."""
    if isinstance(length, torch.Tensor):
        length = length.to(device, dtype)
    first = -(length - 1.0) / length
    return torch.arange(length, dtype=dtype, device=device) * (2.0 / length) + first


def flat_softmax(inp):
    """dsnt/utils.py:71-77 (torch; training / plotting only)."""
    orig_size = inp.size()
    flat = inp.view(-1, reduce(mul, orig_size[2:]))
    flat = torch.nn.functional.softmax(flat, -1)
    return flat.view(*orig_size)


def dsnt(heatmaps, compute_skew=False):
    """dsnt/utils.py:7-47: (coords (N,K,2), var (N,K,2), covar (N,K)) of NORMALISED heatmaps."""
    if compute_skew:
        raise NotImplementedError("compute_skew=True is a debug branch of the reference (prints); not provided")
    norm, _, _ = dsnt_moments(heatmaps, is_logit=False)
    return norm[..., 0:2], norm[..., 2:4], norm[..., 4]


def dsnt_from_logits(logits, covar=True):
    """Fused flat_softmax + dsnt + pixel rescale + get_cov_matrix (dsnt_al.py:118-131).

    Returns (coords, var, covar) on the [-1,1] grid and (pixel_coords, pixel_sigma).
    """
    norm, mu, sig = dsnt_moments(logits, is_logit=True, covar=covar)
    return (norm[..., 0:2], norm[..., 2:4], norm[..., 4]), (mu, sig)


def normalized_to_pixel_coordinates(coords, size):
    """dsnt/utils.py:95-105."""
    if torch.is_tensor(coords):
        size = coords.new_tensor(size).flip(-1)
    return 0.5 * ((coords + 1) * size - 1)


def pixel_to_normalized_coordinates(coords, size):
    """dsnt/utils.py:108-118."""
    if torch.is_tensor(coords):
        size = coords.new_tensor(size).flip(-1)
    return ((2 * coords + 1) / size) - 1

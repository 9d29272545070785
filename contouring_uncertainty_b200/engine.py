"""Thin host layer over the C ABI: tensor checks, output allocation, and the end-to-end engine.

PyTorch is used for device memory and streams only; every computation on the path is a kernel of
``libcub200.so`` (``contouring_uncertainty_b200/csrc``).  Reference call sites are cited per op.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _native as nat

K = 21
P = 42
HW = 256
Z_PSM = 34
Z_NAIVE = 42
MASK_WORDS = 2048
PSM_PACK_DOUBLES = 3142     # CUB200_PSM_PACK_DOUBLES

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
DEFAULT_PSM = os.path.join(DATA_DIR, "camus-cont_psm_11_no_std.npy")
DEFAULT_SEQ_PSM = os.path.join(DATA_DIR, "camus-cont_sequence_psm_11_no_std.npy")


# --------------------------------------------------------------------------------------------------
# K1
# --------------------------------------------------------------------------------------------------
def dsnt_workspace(n_maps: int, device) -> torch.Tensor:
    """Zeroed workspace of ``cub200_dsnt_moments`` for ``n_maps`` heatmaps (row-slab partial moments + tickets)."""
    return torch.zeros((max(int(nat.lib().cub200_dsnt_workspace_bytes(int(n_maps))), 16),), dtype=torch.uint8, device=device)


def dsnt_moments(heat: torch.Tensor, is_logit: bool = True, covar: bool = True):
    """Fused flat_softmax + dsnt + pixel rescale (dsnt/utils.py:7-47,71-77,95-105; dsnt_al.py:118-131).

    heat (N, K, 256, 256) fp32 CUDA.  Returns (norm (N,K,5) [x, y, var_x, var_y, covar on the
    [-1,1] grid], pix_mu (N,K,2), pix_sigma (N,K,2,2)).
    """
    heat = nat.require_cuda(heat, torch.float32, "heat")
    if heat.dim() != 4 or heat.shape[-1] != HW or heat.shape[-2] != HW:
        raise ValueError(f"heat must be (N, K, {HW}, {HW}), got {tuple(heat.shape)}")
    n, k = heat.shape[:2]
    norm = torch.empty((n, k, 5), dtype=torch.float32, device=heat.device)
    mu = torch.empty((n, k, 2), dtype=torch.float32, device=heat.device)
    sig = torch.empty((n, k, 2, 2), dtype=torch.float32, device=heat.device)
    with torch.cuda.device(heat.device):
        rc = nat.lib().cub200_dsnt_moments(nat.ptr(heat), n * k, HW, HW, int(is_logit), int(covar), nat.ptr(norm),
                                           nat.ptr(mu), nat.ptr(sig), nat.ptr(dsnt_workspace(n * k, heat.device)),
                                           nat.stream_ptr(heat.device))
    nat.check(rc, "cub200_dsnt_moments")
    return norm, mu, sig


# --------------------------------------------------------------------------------------------------
# K2
# --------------------------------------------------------------------------------------------------
class ShapeModelConstants:
    """Per-model constants of the posterior shape model, uploaded once.

    Mirrors what ``PosteriorShapeModelSampler.__init__`` loads (psm.py:30-41): scaler_mean,
    scaler_scale, X_train, cast to fp32 like the reference; ``pca`` (posteriorshapemodel.py:9-46)
    of X_train about an arbitrary centre mu is C0 + (xbar - mu)(xbar - mu)^T, so only C0 (42x42)
    and xbar are needed on the device.
    """

    def __init__(self, path: str = DEFAULT_PSM, device="cuda"):
        d = np.load(str(path), allow_pickle=True).item()
        f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
        x = f32(d["X_train"])
        if x.shape[1] != P:
            raise ValueError(f"shape model must be {P}-dimensional, got {x.shape}")
        xbar = x.mean(0)
        c0 = (x - xbar).T @ (x - xbar) / x.shape[0]
        self.path = str(path)
        self.device = torch.device(device)
        self.c0 = torch.tensor(c0, dtype=torch.float64, device=self.device).contiguous()
        self.xbar = torch.tensor(xbar, dtype=torch.float64, device=self.device)
        self.mean = torch.tensor(f32(d["scaler_mean"]), dtype=torch.float64, device=self.device)
        self.scale = torch.tensor(f32(d["scaler_scale"]), dtype=torch.float64, device=self.device)
        # per-level constants of the Sherman-Morrison form (B^-1, K, ...), packed by the library's host helper
        h = [np.ascontiguousarray(a, dtype=np.float64) for a in (c0, xbar, f32(d["scaler_mean"]).reshape(-1), f32(d["scaler_scale"]).reshape(-1))]
        pack = np.empty((PSM_PACK_DOUBLES,), dtype=np.float64)
        rc = nat.lib().cub200_psm_model_pack_host(*(a.ctypes.data for a in h), pack.ctypes.data)
        nat.check(rc, "cub200_psm_model_pack_host")
        self.pack = torch.from_numpy(pack).to(self.device)


def psm_factor(model: ShapeModelConstants, mu: torch.Tensor, cov: torch.Tensor) -> torch.Tensor:
    """Affine map of the PSM sampler per image: (N,21,2),(N,21,2,2) -> A^T (N,34,42) fp32."""
    mu = nat.require_cuda(mu, torch.float32, "mu")
    cov = nat.require_cuda(cov, torch.float32, "cov")
    if mu.dim() < 2 or cov.dim() < 3 or mu.shape[-2:] != (K, 2) or cov.shape[-3:] != (K, 2, 2) or mu.shape[:-2] != cov.shape[:-3]:
        raise ValueError(f"mu (..,{K},2) / cov (..,{K},2,2) expected, got {tuple(mu.shape)} / {tuple(cov.shape)}")
    n = int(np.prod(mu.shape[:-2])) if mu.dim() > 2 else 1
    a = torch.empty(tuple(mu.shape[:-2]) + (Z_PSM, P), dtype=torch.float32, device=mu.device)
    with torch.cuda.device(mu.device):
        rc = nat.lib().cub200_psm_factor(nat.ptr(model.pack), nat.ptr(mu), nat.ptr(cov), n, nat.ptr(a), nat.stream_ptr(mu.device))
    nat.check(rc, "cub200_psm_factor")
    return a


def psm_sample_direct(model: ShapeModelConstants, mu: torch.Tensor, cov: torch.Tensor, z: Optional[torch.Tensor] = None,
                      seed: int = 0, sample_offset: int = 0) -> torch.Tensor:
    """ONE PosteriorShapeModelSampler sample per item, without the factor (what the sequence sampler's second instant needs):
    mu (M,21,2), cov (M,21,2,2) fp32 [, z (M,34)] -> contours (M,21,2) fp32.  Same draws as
    ``sample_contours(mu, psm_factor(model, mu, cov), 1, z, seed, sample_offset)``."""
    mu = nat.require_cuda(mu, torch.float32, "mu")
    cov = nat.require_cuda(cov, torch.float32, "cov")
    if mu.dim() != 3 or mu.shape[1:] != (K, 2) or tuple(cov.shape) != (mu.shape[0], K, 2, 2):
        raise ValueError("mu (M,21,2) / cov (M,21,2,2) expected")
    m, dev = mu.shape[0], mu.device
    if z is not None:
        z = nat.require_cuda(z, torch.float32, "z")
        if tuple(z.shape) != (m, Z_PSM):
            raise ValueError(f"z must be ({m},{Z_PSM})")
    out = torch.empty((m, K, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_psm_sample_direct(nat.ptr(model.pack), nat.ptr(mu), nat.ptr(cov), nat.ptr(z), int(seed) & (2**64 - 1),
                                                int(sample_offset) & (2**64 - 1), m, nat.ptr(out), nat.stream_ptr(dev))
    nat.check(rc, "cub200_psm_sample_direct")
    return out


def naive_factor(cov: torch.Tensor) -> torch.Tensor:
    """Block-diagonal Cholesky colouring of the independent sampler: (N,21,2,2) -> A^T (N,42,42)."""
    cov = nat.require_cuda(cov, torch.float32, "cov")
    if cov.shape[-3:] != (K, 2, 2):
        raise ValueError(f"cov (..,{K},2,2) expected, got {tuple(cov.shape)}")
    n = int(np.prod(cov.shape[:-3])) if cov.dim() > 3 else 1
    a = torch.empty(tuple(cov.shape[:-3]) + (Z_NAIVE, P), dtype=torch.float32, device=cov.device)
    with torch.cuda.device(cov.device):
        rc = nat.lib().cub200_naive_factor(nat.ptr(cov), n, nat.ptr(a), nat.stream_ptr(cov.device))
    nat.check(rc, "cub200_naive_factor")
    return a


# --------------------------------------------------------------------------------------------------
# K3
# --------------------------------------------------------------------------------------------------
def sample_raster_aggregate(mu: torch.Tensor, a_t: torch.Tensor, n_samples: int, z: Optional[torch.Tensor] = None,
                            seed: int = 0, round_landmarks: bool = True, return_contours: bool = True,
                            return_masks: bool = False, diff: Optional[torch.Tensor] = None, images_per_map: int = 1,
                            sample_offset: int = 0):
    """Sampler -> rasteriser -> aggregation.  mu (N,21,2); a_t (N,Z,42); z (N,T,Z) or None.

    Returns dict(contours (N,T,21,2), area (N,T) int32, diff (N,256,256) int32,
    masks (N,T,2048) uint32-as-int32 | None).  ``diff`` may be passed to keep accumulating.
    ``sample_offset``: index of this call's first sample in a larger batch (generated draws are keyed by
    (seed, sample_offset + image * T + sample), so chunked calls reproduce the unchunked one).
    """
    mu = nat.require_cuda(mu, torch.float32, "mu")
    a_t = nat.require_cuda(a_t, torch.float32, "a_t")
    if mu.dim() != 3 or mu.shape[1:] != (K, 2):
        raise ValueError(f"mu must be (N,{K},2), got {tuple(mu.shape)}")
    n = mu.shape[0]
    if a_t.dim() != 3 or a_t.shape[0] != n or a_t.shape[2] != P or not (1 <= a_t.shape[1] <= 42):
        raise ValueError(f"a_t must be (N,Z<=42,{P}), got {tuple(a_t.shape)}")
    zdim = a_t.shape[1]
    T = int(n_samples)
    dev = mu.device
    if z is not None:
        z = nat.require_cuda(z, torch.float32, "z")
        if tuple(z.shape) != (n, T, zdim):
            raise ValueError(f"z must be ({n},{T},{zdim}), got {tuple(z.shape)}")
    contours = torch.empty((n, T, K, 2), dtype=torch.float32, device=dev)   # also the hand-over between the two kernels
    area = torch.empty((n, T), dtype=torch.int32, device=dev)
    if images_per_map < 1 or n % images_per_map:
        raise ValueError("N must be a multiple of images_per_map")
    n_maps = n // images_per_map
    if diff is None:
        diff = torch.zeros((n_maps, HW, HW), dtype=torch.int32, device=dev)
    else:
        diff = nat.require_cuda(diff, torch.int32, "diff")
        if tuple(diff.shape) != (n_maps, HW, HW):
            raise ValueError("diff must be (N / images_per_map,256,256) int32")
    masks = torch.empty((n, T, MASK_WORDS), dtype=torch.int32, device=dev) if return_masks else None
    rws = raster_workspace(n * T, dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_sample_raster_aggregate(nat.ptr(mu), nat.ptr(a_t), zdim, nat.ptr(z), int(seed) & (2**64 - 1),
                                                      int(sample_offset) & (2**64 - 1), n, T, int(round_landmarks),
                                                      nat.ptr(contours), nat.ptr(area), nat.ptr(diff), int(images_per_map),
                                                      nat.ptr(masks), nat.ptr(rws), None, nat.stream_ptr(dev))
    nat.check(rc, "cub200_sample_raster_aggregate")
    return dict(contours=contours, area=area, diff=diff, masks=masks)


def sample_contours(mu: torch.Tensor, a_t: torch.Tensor, n_samples: int, z: Optional[torch.Tensor] = None,
                    seed: int = 0, sample_offset: int = 0) -> torch.Tensor:
    """Sampling only (K3 without rasterisation): contour = mu + A z -> (N,T,21,2) fp32."""
    mu = nat.require_cuda(mu, torch.float32, "mu")
    a_t = nat.require_cuda(a_t, torch.float32, "a_t")
    if mu.dim() != 3 or a_t.dim() != 3:
        raise ValueError("mu (N,21,2) / a_t (N,Z,42) expected")
    n, zdim, T = mu.shape[0], a_t.shape[1], int(n_samples)
    if mu.shape[1:] != (K, 2) or a_t.shape[0] != n or a_t.shape[2] != P or not (1 <= zdim <= 42):
        raise ValueError("mu (N,21,2) / a_t (N,1<=Z<=42,42) expected")
    if z is not None:
        z = nat.require_cuda(z, torch.float32, "z")
        if tuple(z.shape) != (n, T, zdim):
            raise ValueError(f"z must be ({n},{T},{zdim}), got {tuple(z.shape)}")
    out = torch.empty((n, T, K, 2), dtype=torch.float32, device=mu.device)
    with torch.cuda.device(mu.device):
        rc = nat.lib().cub200_sample_raster_aggregate(nat.ptr(mu), nat.ptr(a_t), zdim, nat.ptr(z), int(seed) & (2**64 - 1),
                                                      int(sample_offset) & (2**64 - 1), n, T, 1, nat.ptr(out), None, None, 1,
                                                      None, None, None, nat.stream_ptr(mu.device))
    nat.check(rc, "cub200_sample_raster_aggregate(sample-only)")
    return out


def raster_workspace(n: int, device) -> torch.Tensor:
    """Device workspace of the rasteriser for ``n`` contours (the queue of samples that need the general flood)."""
    return torch.empty((max(int(nat.lib().cub200_raster_workspace_bytes(int(n))), 4),), dtype=torch.uint8, device=device)


def _check_diff(diff: torch.Tensor, m: int, per_image: int, dev) -> torch.Tensor:
    """A caller-supplied difference map for ``m`` contours, ``per_image`` consecutive ones per map."""
    if not isinstance(diff, torch.Tensor) or not diff.is_cuda or diff.dtype != torch.int32 or not diff.is_contiguous():
        raise ValueError("diff must be a contiguous int32 CUDA tensor")
    if diff.device != dev:
        raise ValueError(f"diff lives on {diff.device}, the contours on {dev}")
    per_image = int(per_image)
    if per_image < 1 or m % per_image:
        raise ValueError(f"per_image must be >= 1 and divide the number of contours ({m}), got {per_image}")
    if diff.dim() != 3 or tuple(diff.shape) != (m // per_image, HW, HW):
        raise ValueError(f"diff must be ({m // per_image},{HW},{HW}) for {m} contours, {per_image} per image; got {tuple(diff.shape)}")
    return diff


def rasterize_contours(contours: torch.Tensor, round_landmarks: bool = True, return_masks: bool = True,
                       diff: Optional[torch.Tensor] = None, per_image: int = 0, reconstruction_type: str = "spline"):
    """Rasteriser alone: contours (M,21,2) fp32 or fp64 -> (masks (M,2048) int32 bit-packed | None, area (M) int32).
    reconstruction_type 'linear' = linear_reconstruction (cu/utils/contour.py:43-53): rounded landmarks joined by lines."""
    if reconstruction_type not in ("spline", "linear"):
        raise ValueError("reconstruction_type must be 'spline' or 'linear'")
    flags = 3 if reconstruction_type == "linear" else int(bool(round_landmarks))
    f64 = isinstance(contours, torch.Tensor) and contours.dtype == torch.float64
    contours = nat.require_cuda(contours, torch.float64 if f64 else torch.float32, "contours")
    if contours.dim() != 3 or contours.shape[1:] != (K, 2):
        raise ValueError(f"contours must be (M,{K},2), got {tuple(contours.shape)}")
    m = contours.shape[0]
    dev = contours.device
    if diff is not None:
        # the kernel adds into diff[(i // per_image)]: a wrong dtype / shape / device would be an out-of-bounds device write
        diff = _check_diff(diff, m, per_image, dev)
    masks = torch.empty((m, MASK_WORDS), dtype=torch.int32, device=dev) if return_masks else None
    area = torch.empty((m,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        fn = nat.lib().cub200_rasterize_contours_f64 if f64 else nat.lib().cub200_rasterize_contours
        rc = fn(nat.ptr(contours), m, flags, nat.ptr(masks), nat.ptr(area), nat.ptr(diff), int(per_image),
                nat.ptr(raster_workspace(m, dev)), nat.stream_ptr(dev))
    nat.check(rc, "cub200_rasterize_contours")
    return masks, area


RASTER_MAX_POINTS = 64


def rasterize_contours_generic(contours: torch.Tensor, round_landmarks: bool = True, reconstruction_type: str = "spline",
                               return_masks: bool = True):
    """reconstruction / linear_reconstruction (cu/utils/contour.py:28-53) for contours of any size: contours (M, K, 2) fp32 or
    fp64 with 4 <= K <= 64 -> (masks (M, 2048) int32 bit-packed | None, area (M) int32).  K == 21 may use rasterize_contours."""
    if reconstruction_type not in ("spline", "linear"):
        raise ValueError("reconstruction_type must be 'spline' or 'linear'")
    f64 = isinstance(contours, torch.Tensor) and contours.dtype == torch.float64
    contours = nat.require_cuda(contours, torch.float64 if f64 else torch.float32, "contours")
    if contours.dim() != 3 or contours.shape[2] != 2 or not (4 <= contours.shape[1] <= RASTER_MAX_POINTS):
        raise ValueError(f"contours must be (M, K, 2) with 4 <= K <= {RASTER_MAX_POINTS}, got {tuple(contours.shape)}")
    m, k = int(contours.shape[0]), int(contours.shape[1])
    dev = contours.device
    masks = torch.empty((m, MASK_WORDS), dtype=torch.int32, device=dev) if return_masks else None
    area = torch.empty((m,), dtype=torch.int32, device=dev)
    flags = 3 if reconstruction_type == "linear" else int(bool(round_landmarks))
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_rasterize_contours_generic(nat.ptr(contours), int(f64), m, k, flags, nat.ptr(masks), nat.ptr(area),
                                                         nat.stream_ptr(dev))
    nat.check(rc, "cub200_rasterize_contours_generic")
    return masks, area


def us_myo_masks(landmarks: torch.Tensor, reconstruction_type: str = "spline"):
    """USContourToMask's two-structure branch (cu/data/camus/utils.py:47-82): landmarks (M, 42, 2) fp32 = 21 endocardial +
    21 epicardial points -> (lv masks (M, 2048) int32, myo masks (M, 2048) int32, areas (M, 2) int32), bit-packed."""
    if reconstruction_type not in ("spline", "linear"):
        raise ValueError("reconstruction_type must be 'spline' or 'linear'")
    landmarks = nat.require_cuda(landmarks, torch.float32, "landmarks")
    if landmarks.dim() != 3 or tuple(landmarks.shape[1:]) != (2 * K, 2):
        raise ValueError(f"landmarks must be (M, {2 * K}, 2), got {tuple(landmarks.shape)}")
    m = int(landmarks.shape[0])
    dev = landmarks.device
    lv = torch.empty((m, MASK_WORDS), dtype=torch.int32, device=dev)
    myo = torch.empty((m, MASK_WORDS), dtype=torch.int32, device=dev)
    area = torch.empty((m, 2), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_us_myo_masks(nat.ptr(landmarks), m, 2 if reconstruction_type == "linear" else 0, nat.ptr(lv),
                                           nat.ptr(myo), nat.ptr(area), nat.stream_ptr(dev))
    nat.check(rc, "cub200_us_myo_masks")
    return lv, myo, area


def mask_class_entropy(lv_masks: torch.Tensor, myo_masks: torch.Tensor, want_counts: bool = False, want_pred: bool = False):
    """sample_entropy of LV + MYO one-hot sample stacks (cu/task/uncertainty.py:107-133, base 3) from the bit-packed masks of
    ``us_myo_masks``: lv_masks, myo_masks (N, T, 2048) int32 -> entropy (N,256,256) fp64 [, counts (N,3,256,256) int32]
    [, pred (N,256,256) uint8 = argmax of the class counts]."""
    lv_masks = nat.require_cuda(lv_masks, torch.int32, "lv_masks")
    myo_masks = nat.require_cuda(myo_masks, torch.int32, "myo_masks")
    if lv_masks.dim() != 3 or lv_masks.shape[2] != MASK_WORDS or lv_masks.shape != myo_masks.shape or lv_masks.shape[1] < 1:
        raise ValueError(f"lv_masks / myo_masks must both be (N, T >= 1, {MASK_WORDS}) int32")
    n, t = int(lv_masks.shape[0]), int(lv_masks.shape[1])
    dev = lv_masks.device
    ent = torch.empty((n, HW, HW), dtype=torch.float64, device=dev)
    cnt = torch.empty((n, 3, HW, HW), dtype=torch.int32, device=dev) if want_counts else None
    pred = torch.empty((n, HW, HW), dtype=torch.uint8, device=dev) if want_pred else None
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_mask_class_entropy(nat.ptr(lv_masks), nat.ptr(myo_masks), n, t, nat.ptr(ent), nat.ptr(cnt), nat.ptr(pred),
                                                 nat.stream_ptr(dev))
    nat.check(rc, "cub200_mask_class_entropy")
    out = (ent,)
    if want_counts:
        out += (cnt,)
    if want_pred:
        out += (pred,)
    return out if len(out) > 1 else ent


def sample_entropy(samples: torch.Tensor) -> torch.Tensor:
    """UncertaintyTask.sample_entropy (cu/task/uncertainty.py:107-133) for probabilities of any class count:
    samples (N, T, C, H, W) fp32 with 1 <= C <= 8 -> (N, H, W) fp32 (C == 1: bits; else base C)."""
    samples = nat.require_cuda(samples, torch.float32, "samples")
    if samples.dim() != 5 or not (1 <= samples.shape[2] <= 8) or samples.shape[1] < 1:
        raise ValueError(f"samples must be (N, T, C, H, W) with T >= 1 and 1 <= C <= 8, got {tuple(samples.shape)}")
    n, t, c, h, w = (int(v) for v in samples.shape)
    out = torch.empty((n, h, w), dtype=torch.float32, device=samples.device)
    with torch.cuda.device(samples.device):
        rc = nat.lib().cub200_sample_entropy(nat.ptr(samples), n, t, c, h * w, nat.ptr(out), nat.stream_ptr(samples.device))
    nat.check(rc, "cub200_sample_entropy")
    return out


def unpack_masks(masks: torch.Tensor) -> torch.Tensor:
    """(..., 2048) int32 bit-packed -> (..., 256, 256) uint8 (host or device; plain torch bit ops)."""
    w = masks.reshape(masks.shape[:-1] + (HW, 8, 1))
    bits = (w >> torch.arange(32, device=masks.device, dtype=torch.int32)) & 1
    return bits.reshape(masks.shape[:-1] + (HW, HW)).to(torch.uint8)


# --------------------------------------------------------------------------------------------------
# K4
# --------------------------------------------------------------------------------------------------
def finalize_maps(diff: torch.Tensor, n_samples: int, want_count: bool = True, want_variance: bool = False,
                  want_count16: bool = False, want_maps: bool = True):
    """difference map -> (count int32 | None, mean fp64, entropy fp64 [bits], pred uint8), all (N,256,256)
    [, variance fp64 = p (1 - p) when ``want_variance``] [, count uint16 when ``want_count16``, n_samples <= 65535].

    aleatoric.py:111 (pred) and uncertainty.py:107-133 (sample_entropy).  ``want_maps=False`` skips the fp64 mean / entropy
    maps (they are functions of the count map: ``maps_from_count`` expands them on the host from 2 B / pixel).
    """
    diff = nat.require_cuda(diff, torch.int32, "diff")
    if diff.dim() != 3 or diff.shape[1:] != (HW, HW):
        raise ValueError("diff must be (N,256,256) int32")
    if want_count16 and int(n_samples) > 65535:
        raise ValueError("a uint16 count map needs n_samples <= 65535")
    n = diff.shape[0]
    dev = diff.device
    count = torch.empty_like(diff) if want_count else None
    mean = torch.empty((n, HW, HW), dtype=torch.float64, device=dev) if want_maps else None
    ent = torch.empty((n, HW, HW), dtype=torch.float64, device=dev) if want_maps else None
    var = torch.empty((n, HW, HW), dtype=torch.float64, device=dev) if want_variance else None
    c16 = torch.empty((n, HW, HW), dtype=torch.uint16, device=dev) if want_count16 else None
    pred = torch.empty((n, HW, HW), dtype=torch.uint8, device=dev)
    lut = torch.empty((3 * (int(n_samples) + 1),), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_finalize_maps(nat.ptr(diff), n, int(n_samples), nat.ptr(count), nat.ptr(mean), nat.ptr(ent),
                                            nat.ptr(pred), nat.ptr(lut), 0, nat.ptr(var), nat.ptr(c16), nat.stream_ptr(dev))
    nat.check(rc, "cub200_finalize_maps")
    out = (count, mean, ent, pred)
    if want_variance:
        out += (var,)
    if want_count16:
        out += (c16,)
    return out


def maps_from_count(count: np.ndarray, n_samples: int) -> dict:
    """Host-side expansion of a count map (any integer dtype) into the maps the device would write: mean = c / T,
    entropy = scipy.stats.entropy([p, 1 - p], base=2) evaluated as the kernel's table does, pred = round-half-even(mean),
    variance = p (1 - p).  Bit-identical to ``finalize_maps`` (same fp64 operations on the same integers)."""
    T = int(n_samples)
    c = np.arange(T + 1, dtype=np.float64)
    p = c / T
    q = 1.0 - p
    s = p + q
    pn, qn = p / s, q / s
    with np.errstate(divide="ignore", invalid="ignore"):
        e = (np.where(pn > 0, -pn * np.log(pn), 0.0) + np.where(qn > 0, -qn * np.log(qn), 0.0)) / np.log(2.0)
    e[~np.isfinite(e)] = 0.0
    idx = np.asarray(count).astype(np.int64)
    return dict(mean=p[idx], entropy=e[np.clip(idx, 0, T)], pred=(2 * idx > T).astype(np.uint8), variance=(p * (1.0 - p))[idx])


def _gaussian_taps(sigma: float = 5.0, truncate: float = 1.0) -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius), centre tap first."""
    radius = int(truncate * float(sigma) + 0.5)
    xs = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * xs ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:])


def soft_masks(masks: torch.Tensor) -> torch.Tensor:
    """convert_to_mask(soft_mask=True) (cu/task/regression/contour_uncertainty.py:44-49): bit-packed masks (..., 2048) int32 ->
    (..., 256, 256) fp64 = min-max normalised Gaussian blur (sigma 5, truncate 1, edges repeated)."""
    masks = nat.require_cuda(masks, torch.int32, "masks")
    if masks.shape[-1] != MASK_WORDS:
        raise ValueError(f"masks must be (..., {MASK_WORDS}) int32 bit-packed")
    lead = tuple(masks.shape[:-1])
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    dev = masks.device
    taps = _gaussian_taps()
    if len(taps) != 6:
        raise RuntimeError("the kernel is built for 11 taps")
    w = torch.from_numpy(taps).to(dev)
    out = torch.empty(lead + (HW, HW), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_soft_masks(nat.ptr(masks), n, nat.ptr(w), nat.ptr(out), nat.stream_ptr(dev))
    nat.check(rc, "cub200_soft_masks")
    return out


def metric_ae_split(values: torch.Tensor, value_range=None, return_values: bool = False):
    """Aleatoric / epistemic split of Monte-Carlo metric distributions (results/clinical/utils.py:3-18), batched:
    values (B, T_e, T_a) fp64 CUDA -> dict(stats (B,4) [mean, aleatoric_std, epistemic_std, total], n_reject (B) int32
    [, values (B,T_e,T_a) with the rejected samples as NaN, reject (B,T_e,T_a) uint8]).  ``value_range`` = (lo, hi): the
    reject rule of FAC / GLS / EF (values outside become NaN before the statistics)."""
    values = nat.require_cuda(values, torch.float64, "values")
    if values.dim() != 3 or values.shape[1] < 1 or values.shape[2] < 1:
        raise ValueError(f"values must be (B, T_e >= 1, T_a >= 1), got {tuple(values.shape)}")
    b, te, ta = (int(v) for v in values.shape)
    dev = values.device
    stats = torch.empty((b, 4), dtype=torch.float64, device=dev)
    nrej = torch.empty((b,), dtype=torch.int32, device=dev)
    vout = torch.empty_like(values) if return_values else None
    rej = torch.empty((b, te, ta), dtype=torch.uint8, device=dev) if return_values else None
    lo, hi = (0.0, 0.0) if value_range is None else (float(value_range[0]), float(value_range[1]))
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_metric_ae_split(nat.ptr(values), b, te, ta, int(value_range is not None), lo, hi, nat.ptr(vout),
                                              nat.ptr(rej), nat.ptr(stats), nat.ptr(nrej), nat.stream_ptr(dev))
    nat.check(rc, "cub200_metric_ae_split")
    out = dict(stats=stats, n_reject=nrej)
    if return_values:
        out.update(values=vout, reject=rej)
    return out


def fac_from_areas(area_ed: torch.Tensor, area_es: torch.Tensor) -> torch.Tensor:
    """Per-sample FAC = (A_ED - A_ES)/A_ED, NaN outside [0,1] (clinical.py:11-29, view.py:35-78)."""
    area_ed = nat.require_cuda(area_ed, torch.int32, "area_ed")
    area_es = nat.require_cuda(area_es, torch.int32, "area_es")
    if area_ed.shape != area_es.shape:
        raise ValueError("area_ed / area_es shapes differ")
    out = torch.empty(area_ed.shape, dtype=torch.float64, device=area_ed.device)
    with torch.cuda.device(area_ed.device):
        rc = nat.lib().cub200_fac_from_areas(nat.ptr(area_ed), nat.ptr(area_es), area_ed.numel(), nat.ptr(out),
                                             nat.stream_ptr(area_ed.device))
    nat.check(rc, "cub200_fac_from_areas")
    return out


def metric_sums(values: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(sum, sum of squares) of an int32 per-sample metric as exact int64 (dataset-level accumulators)."""
    values = nat.require_cuda(values, torch.int32, "values")
    if out is None:
        out = torch.empty((2,), dtype=torch.int64, device=values.device)
    with torch.cuda.device(values.device):
        rc = nat.lib().cub200_metric_sums(nat.ptr(values), values.numel(), nat.ptr(out), nat.stream_ptr(values.device))
    nat.check(rc, "cub200_metric_sums")
    return out


class SequenceModelConstants:
    """Per-model constants of the 84-dim (ED, ES) sequence shape model (sequence_sampler.py:15-24).

    seq_mu / seq_Q are used as stored (the reference does not re-centre them), so for each choice b
    of the first-sampled instant the conditional-mean map M_b = S[o,g] (S[g,g] + I)^-1 and the 2x2
    conditional covariances are constants; they are computed here once, in fp64, on the host.
    """

    def __init__(self, path: str = DEFAULT_SEQ_PSM, device="cuda"):
        d = np.load(str(path), allow_pickle=True).item()
        f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
        q = f32(d["Q"])
        if q.shape != (2 * P, 2 * P):
            raise ValueError(f"sequence shape model must be {2 * P}-dimensional, got {q.shape}")
        S = q @ q.T
        scale = f32(d["scaler_scale"]).reshape(-1)
        m2t = np.zeros((2, P, P))
        covc = np.zeros((2, K, 2, 2))
        for b in (0, 1):
            g = np.arange(P) + P * b
            o = np.arange(P) + P * (1 - b)
            M = np.linalg.solve(S[np.ix_(g, g)] + np.eye(P), S[np.ix_(g, o)]).T          # (o, g)
            m2t[b] = M.T                                                                  # [e][row]
            C = (S[np.ix_(o, o)] - M @ S[np.ix_(g, o)]) * scale[o][None, :]               # cov_c *= seq_scale (columns)
            covc[b] = np.stack([C[2 * i:2 * i + 2, 2 * i:2 * i + 2] for i in range(K)])
        self.path = str(path)
        self.device = torch.device(device)
        t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=self.device)
        self.m2t, self.covc = t(m2t), t(covc)
        self.mu, self.mean, self.scale = t(f32(d["mu"]).reshape(-1)), t(f32(d["scaler_mean"]).reshape(-1)), t(scale)
        # for the skew sequence sampler, which re-centres the 84-dim PCA on every predicted pair
        # (psm_skew_sequence.py:66): pca() about any centre is C0 + d d^T
        x = f32(d["X_train"])
        xbar = x.mean(0)
        self.c0, self.xbar = t((x - xbar).T @ (x - xbar) / x.shape[0]), t(xbar)


def sequence_sample(model: ShapeModelConstants, seq: SequenceModelConstants, mu: torch.Tensor, cov: torch.Tensor,
                    n_samples: int, z: Optional[torch.Tensor] = None, first_instant: Optional[torch.Tensor] = None,
                    seed: int = 0) -> torch.Tensor:
    """SequencePSMSampler (sequence_sampler.py:26-144) for a batch of (ED, ES) pairs.

    mu (Np,2,21,2), cov (Np,2,21,2,2) fp32 CUDA; z (Np,T,2,34) optional draws (z[:,:,0] for the
    first-sampled instant, z[:,:,1] for the other); first_instant (Np,T) uint8 optional (else
    Bernoulli(1/2) from torch's generator).  Returns contours (Np, T, 2, 21, 2).

    Pipeline: K2 on both instants -> one PSM sample of every image (sample-only K3) -> seq_condition
    (84-dim conditioning + merge) -> ONE direct PSM sample per (pair, sample) from the merged distribution of the other
    instant (its PCA is re-centred per sample, psm.py:91: `psm_sample_direct`, no per-sample factor).
    """
    mu = nat.require_cuda(mu, torch.float32, "mu")
    cov = nat.require_cuda(cov, torch.float32, "cov")
    if mu.dim() != 4 or mu.shape[1:] != (2, K, 2) or cov.shape != mu.shape[:3] + (2, 2):
        raise ValueError("mu (Np,2,21,2) / cov (Np,2,21,2,2) expected")
    npair, T, dev = mu.shape[0], int(n_samples), mu.device
    if first_instant is None:
        first_instant = (torch.rand((npair, T), device=dev) < 0.5).to(torch.uint8)
    first_instant = nat.require_cuda(first_instant, torch.uint8, "first_instant")
    if tuple(first_instant.shape) != (npair, T):
        raise ValueError(f"first_instant must be ({npair},{T})")
    z1 = z2 = None
    if z is not None:
        z = nat.require_cuda(z, torch.float32, "z")
        if tuple(z.shape) != (npair, T, 2, Z_PSM):
            raise ValueError(f"z must be ({npair},{T},2,{Z_PSM})")
        z1 = z[:, :, 0].repeat_interleave(2, dim=0).contiguous()          # same draws whichever instant goes first
        z2 = z[:, :, 1].reshape(npair * T, 1, Z_PSM).contiguous()
    m2, c2 = mu.reshape(2 * npair, K, 2), cov.reshape(2 * npair, K, 2, 2)
    a_t = psm_factor(model, m2, c2)
    first_all = sample_contours(m2, a_t, T, z=z1, seed=seed)                 # (2Np,T,21,2)
    first = torch.empty((npair, T, K, 2), dtype=torch.float32, device=dev)
    mu_f = torch.empty((npair * T, K, 2), dtype=torch.float32, device=dev)
    cov_f = torch.empty((npair * T, K, 2, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_seq_condition(nat.ptr(mu), nat.ptr(cov), nat.ptr(first_all), nat.ptr(first_instant), npair, T,
                                            nat.ptr(seq.m2t), nat.ptr(seq.covc), nat.ptr(seq.mu), nat.ptr(seq.mean),
                                            nat.ptr(seq.scale), nat.ptr(first), nat.ptr(mu_f), nat.ptr(cov_f),
                                            nat.stream_ptr(dev))
    nat.check(rc, "cub200_seq_condition")
    # one sample from each per-(pair, sample) law: drawn directly, without its 34 x 42 factor (same draws as the factor path)
    second = psm_sample_direct(model, mu_f, cov_f, z=None if z2 is None else z2.reshape(npair * T, Z_PSM),
                               seed=seed + 0x9E3779B9).reshape(npair, T, K, 2)
    fi = first_instant.bool()[..., None, None]
    out = torch.empty((npair, T, 2, K, 2), dtype=torch.float32, device=dev)
    out[:, :, 0] = torch.where(fi, second, first)                            # instant 0 is "first" when bit == 0
    out[:, :, 1] = torch.where(fi, first, second)
    return out


def sample_moments(contours: torch.Tensor):
    """post_mu / post_cov (aleatoric.py:96-102): contours (..., T, 21, 2) -> mean (...,21,2), cov (...,21,2,2) fp64, ddof=1."""
    contours = nat.require_cuda(contours, torch.float32, "contours")
    if contours.dim() < 3 or contours.shape[-2:] != (K, 2):
        raise ValueError(f"contours must be (..., T, {K}, 2)")
    lead, T = tuple(contours.shape[:-3]), contours.shape[-3]
    n = int(np.prod(lead)) if lead else 1
    mean = torch.empty(lead + (K, 2), dtype=torch.float64, device=contours.device)
    cov = torch.empty(lead + (K, 2, 2), dtype=torch.float64, device=contours.device)
    with torch.cuda.device(contours.device):
        rc = nat.lib().cub200_sample_moments(nat.ptr(contours), n, T, nat.ptr(mean), nat.ptr(cov), nat.stream_ptr(contours.device))
    nat.check(rc, "cub200_sample_moments")
    return mean, cov


# --------------------------------------------------------------------------------------------------
# Skew-normal path (dsnt-skew)
# --------------------------------------------------------------------------------------------------
N_DRAW_SKEW = 14      # multinomial draws per sample (level points)


def skew_pdf_grid(mu: torch.Tensor, cov: torch.Tensor, alpha: torch.Tensor, flip_alpha_y: bool = False,
                  sampled_only: bool = False) -> torch.Tensor:
    """BivariateSkewNormal.pdf on the 256x256 integer grid (bivariateskewnormal.py:20-49; the ``pdfs``
    loop of psm_skew.py:234-238).  mu (...,2), cov (...,2,2), alpha (...,2) -> (...,256,256) fp32,
    table[x, y] (``indexing='ij'``).  ``sampled_only`` (inputs shaped (...,21,*)) leaves the tables of landmarks
    4, 9, 11, 16 unwritten: the samplers never read them."""
    mu = nat.require_cuda(mu, torch.float32, "mu")
    cov = nat.require_cuda(cov, torch.float32, "cov")
    alpha = nat.require_cuda(alpha, torch.float32, "alpha")
    lead = tuple(mu.shape[:-1])
    if mu.shape[-1] != 2 or tuple(cov.shape) != lead + (2, 2) or tuple(alpha.shape) != lead + (2,):
        raise ValueError("mu (...,2) / cov (...,2,2) / alpha (...,2) expected")
    n = int(np.prod(lead)) if lead else 1
    out = torch.empty(lead + (HW, HW), dtype=torch.float32, device=mu.device)
    with torch.cuda.device(mu.device):
        rc = nat.lib().cub200_skew_pdf_grid(nat.ptr(mu), nat.ptr(cov), nat.ptr(alpha), n, int(flip_alpha_y) | (2 if sampled_only else 0),
                                            nat.ptr(out), nat.stream_ptr(mu.device))
    nat.check(rc, "cub200_skew_pdf_grid")
    return out


def bivariate_pdf(x: torch.Tensor, mu: torch.Tensor, cov: torch.Tensor, alpha: Optional[torch.Tensor] = None, log: bool = False):
    """BivariateNormal / BivariateSkewNormal pdf or logpdf at points x (..., 2) (bivariatenormal.py:15-36,
    bivariateskewnormal.py:20-49); one law (mu (2,), cov (2,2), alpha (2,)) or one per point.  Returns x.shape[:-1] fp32."""
    x = nat.require_cuda(x, torch.float32, "x")
    mu = nat.require_cuda(mu, torch.float32, "mu")
    cov = nat.require_cuda(cov, torch.float32, "cov")
    if alpha is not None:
        alpha = nat.require_cuda(alpha, torch.float32, "alpha")
    if x.shape[-1] != 2:
        raise ValueError("last dim of x must be 2")
    lead = tuple(x.shape[:-1])
    n = int(np.prod(lead)) if lead else 1
    nd = mu.numel() // 2
    if cov.numel() != 4 * nd or (alpha is not None and alpha.numel() != 2 * nd) or nd not in (1, n):
        raise ValueError("mu / cov / alpha must describe one law or one law per point")
    out = torch.empty(lead, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = nat.lib().cub200_bivariate_pdf(nat.ptr(x), n, nat.ptr(mu), nat.ptr(cov), nat.ptr(alpha), nd, int(log), nat.ptr(out),
                                            nat.stream_ptr(x.device))
    nat.check(rc, "cub200_bivariate_pdf")
    return out


def skew_rvs_fast(mu: torch.Tensor, cov: torch.Tensor, alpha: torch.Tensor, n: int, z: Optional[torch.Tensor] = None, seed: int = 0):
    """BivariateSkewNormal.rvs_fast (bivariateskewnormal.py:160-191): n draws (n,2) of one law; z (n,3) replays the normals."""
    mu = nat.require_cuda(mu, torch.float32, "mu")
    cov = nat.require_cuda(cov, torch.float32, "cov")
    alpha = nat.require_cuda(alpha, torch.float32, "alpha")
    if mu.numel() != 2 or cov.numel() != 4 or alpha.numel() != 2:
        raise ValueError("mu (2,), cov (2,2), alpha (2,) expected")
    if z is not None:
        z = nat.require_cuda(z, torch.float32, "z")
        if tuple(z.shape) != (n, 3):
            raise ValueError(f"z must be ({n},3)")
    out = torch.empty((int(n), 2), dtype=torch.float32, device=mu.device)
    with torch.cuda.device(mu.device):
        rc = nat.lib().cub200_skew_rvs_fast(nat.ptr(mu), nat.ptr(cov), nat.ptr(alpha), nat.ptr(z), int(seed) & (2**64 - 1), int(n),
                                            nat.ptr(out), nat.stream_ptr(mu.device))
    nat.check(rc, "cub200_skew_rvs_fast")
    return out


def psm_conditionals(model: ShapeModelConstants, mu: torch.Tensor):
    """Per-image conditional maps of the shape model (compute_psm, psm_skew.py:469-491):
    mu (N,21,2) -> (mc_t (N,34,36) fp64, covc (N,14,3) fp64), see include/cub200.h."""
    mu = nat.require_cuda(mu, torch.float32, "mu")
    if mu.dim() != 3 or mu.shape[1:] != (K, 2):
        raise ValueError(f"mu must be (N,{K},2)")
    n = mu.shape[0]
    mc = torch.empty((n, Z_PSM, 36), dtype=torch.float64, device=mu.device)
    cc = torch.empty((n, N_DRAW_SKEW, 3), dtype=torch.float64, device=mu.device)
    with torch.cuda.device(mu.device):
        rc = nat.lib().cub200_psm_conditionals(nat.ptr(model.pack), nat.ptr(mu), n, nat.ptr(mc), nat.ptr(cc), nat.stream_ptr(mu.device))
    nat.check(rc, "cub200_psm_conditionals")
    return mc, cc


def skew_sample(model: ShapeModelConstants, mu: torch.Tensor, cov: torch.Tensor, alpha: torch.Tensor, n_samples: int,
                z: Optional[torch.Tensor] = None, expo: Optional[torch.Tensor] = None, seed: int = 0,
                return_indices: bool = False, pdfs: Optional[torch.Tensor] = None):
    """SkewPosteriorShapeModelSampler.__call__ (psm_skew.py:187-411) for a batch of images.

    mu (N,21,2), cov (N,21,2,2), alpha (N,21,2) fp32 CUDA as predicted.  z (N,T,3,3): normals of
    rvs_fast; expo (N,T,14,65536): Exp(1) fields of the multinomial races (both optional: Philox).
    Returns contours (N,T,21,2) [, indices (N,T,14) int32, -1 = fell back to mu_c]."""
    mu = nat.require_cuda(mu, torch.float32, "mu")
    cov = nat.require_cuda(cov, torch.float32, "cov")
    alpha = nat.require_cuda(alpha, torch.float32, "alpha")
    if mu.dim() != 3 or mu.shape[1:] != (K, 2) or tuple(cov.shape) != tuple(mu.shape) + (2,) or alpha.shape != mu.shape:
        raise ValueError("mu (N,21,2) / cov (N,21,2,2) / alpha (N,21,2) expected")
    n, T, dev = mu.shape[0], int(n_samples), mu.device
    if z is not None:
        z = nat.require_cuda(z, torch.float32, "z")
        if tuple(z.shape) != (n, T, 3, 3):
            raise ValueError(f"z must be ({n},{T},3,3)")
    if expo is not None:
        expo = nat.require_cuda(expo, torch.float32, "expo")
        if tuple(expo.shape) != (n, T, N_DRAW_SKEW, HW * HW):
            raise ValueError(f"expo must be ({n},{T},{N_DRAW_SKEW},{HW * HW})")
    if pdfs is None:
        pdfs = skew_pdf_grid(mu, cov, alpha, flip_alpha_y=True, sampled_only=True)
    else:
        pdfs = nat.require_cuda(pdfs, torch.float32, "pdfs")
        if tuple(pdfs.shape) != (n, K, HW, HW) or pdfs.device != dev:
            raise ValueError("pdfs must be (N,21,256,256) fp32 on the device of mu")
    mc, cc = psm_conditionals(model, mu)
    out = torch.empty((n, T, K, 2), dtype=torch.float32, device=dev)
    idx = torch.empty((n, T, N_DRAW_SKEW), dtype=torch.int32, device=dev) if return_indices else None
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_skew_sample(nat.ptr(mu), nat.ptr(cov), nat.ptr(alpha), nat.ptr(pdfs), nat.ptr(mc), nat.ptr(cc),
                                          nat.ptr(z), nat.ptr(expo), int(seed) & (2**64 - 1), n, T, nat.ptr(out), nat.ptr(idx),
                                          nat.stream_ptr(dev))
    nat.check(rc, "cub200_skew_sample")
    return (out, idx) if return_indices else out


def seq_skew_sample(model: ShapeModelConstants, seq: "SequenceModelConstants", mu: torch.Tensor, cov: torch.Tensor,
                    alpha: torch.Tensor, n_samples: int, first_instant: Optional[torch.Tensor] = None,
                    z: Optional[torch.Tensor] = None, expo: Optional[torch.Tensor] = None, seed: int = 0,
                    return_indices: bool = False):
    """SequenceSkewPSMSampler (psm_skew_sequence.py:34-104) for a batch of (ED, ES) pairs.

    mu (Np,2,21,2), cov (Np,2,21,2,2), alpha (Np,2,21,2) fp32 CUDA; first_instant (Np,T) uint8 (else Bernoulli(1/2));
    z (Np,T,3,3); expo (Np,T,31,65536).  Returns contours (Np, 2, T, 21, 2) [, indices (Np,T,31)]."""
    mu = nat.require_cuda(mu, torch.float32, "mu")
    cov = nat.require_cuda(cov, torch.float32, "cov")
    alpha = nat.require_cuda(alpha, torch.float32, "alpha")
    if mu.dim() != 4 or mu.shape[1:] != (2, K, 2) or tuple(cov.shape) != tuple(mu.shape) + (2,) or alpha.shape != mu.shape:
        raise ValueError("mu (Np,2,21,2) / cov (Np,2,21,2,2) / alpha (Np,2,21,2) expected")
    npair, T, dev = mu.shape[0], int(n_samples), mu.device
    if first_instant is None:
        first_instant = (torch.rand((npair, T), device=dev) < 0.5).to(torch.uint8)
    first_instant = nat.require_cuda(first_instant, torch.uint8, "first_instant")
    if tuple(first_instant.shape) != (npair, T):
        raise ValueError(f"first_instant must be ({npair},{T})")
    if z is not None:
        z = nat.require_cuda(z, torch.float32, "z")
        if tuple(z.shape) != (npair, T, 3, 3):
            raise ValueError(f"z must be ({npair},{T},3,3)")
    if expo is not None:
        expo = nat.require_cuda(expo, torch.float32, "expo")
        if tuple(expo.shape) != (npair, T, 31, HW * HW):
            raise ValueError(f"expo must be ({npair},{T},31,{HW * HW})")
    m2, c2, a2 = mu.reshape(2 * npair, K, 2), cov.reshape(2 * npair, K, 2, 2), alpha.reshape(2 * npair, K, 2)
    tab_f = skew_pdf_grid(m2, c2, a2, flip_alpha_y=True, sampled_only=True)
    tab_r = skew_pdf_grid(m2, c2, a2, flip_alpha_y=False, sampled_only=True)
    mc, cc = psm_conditionals(model, m2)
    m_seq = torch.empty((npair, 2, P, P), dtype=torch.float64, device=dev)
    c_seq = torch.empty((npair, 2, K, 4), dtype=torch.float64, device=dev)
    out = torch.empty((npair, 2, T, K, 2), dtype=torch.float32, device=dev)
    idx = torch.empty((npair, T, 31), dtype=torch.int32, device=dev) if return_indices else None
    L = nat.lib()
    with torch.cuda.device(dev):
        st = nat.stream_ptr(dev)
        nat.check(L.cub200_seq_skew_conditionals(nat.ptr(seq.c0), nat.ptr(seq.xbar), nat.ptr(seq.mean), nat.ptr(seq.scale),
                                                 nat.ptr(mu), npair, nat.ptr(m_seq), nat.ptr(c_seq), st), "cub200_seq_skew_conditionals")
        nat.check(L.cub200_seq_skew_sample(nat.ptr(mu), nat.ptr(cov), nat.ptr(alpha), nat.ptr(tab_f), nat.ptr(tab_r), nat.ptr(mc),
                                           nat.ptr(cc), nat.ptr(m_seq), nat.ptr(c_seq), nat.ptr(first_instant), nat.ptr(z),
                                           nat.ptr(expo), int(seed) & (2**64 - 1), npair, T, nat.ptr(out), nat.ptr(idx), st),
                  "cub200_seq_skew_sample")
    return (out, idx) if return_indices else out


# --------------------------------------------------------------------------------------------------
# Simpson's biplane volumes (uncertainty propagation)
# --------------------------------------------------------------------------------------------------
def pil_nearest_index(n_in: int, n_out: int) -> np.ndarray:
    """Source index of every output index of PIL's NEAREST resize (ImagingScaleAffine): (int)(0.5 a + k a), a = n_in / n_out,
    with the sum formed by repeated double additions exactly as Pillow does (it differs from (k + 0.5) a at exact ties)."""
    a = n_in / n_out
    xo = a * 0.5
    out = np.empty(n_out, dtype=np.uint16)
    for k in range(n_out):
        out[k] = min(int(xo), n_in - 1)
        xo += a
    return out


def isotropic_view(voxelspacing) -> tuple:
    """resize_image_to_isotropic (vital/vital/utils/image/transform.py:52-66) for a 256 x 256 mask: (height, width,
    row_index, col_index, isotropic spacing) of the nearest-neighbour stretch; (256, 256, None, None, s) when isotropic."""
    sy, sx = float(voxelspacing[0]), float(voxelspacing[1])
    iso = min(sy, sx)
    scaling = np.array((sy, sx)) / np.array((iso, iso))
    nh, nw = (np.array((HW, HW)) * scaling).round().astype(int)
    if nh == HW and nw == HW:
        return HW, HW, None, None, iso
    return int(nh), int(nw), pil_nearest_index(HW, int(nh)), pil_nearest_index(HW, int(nw)), iso


def simpson_diameters(masks: torch.Tensor, spacing, return_npts: bool = False):
    """_compute_diameters (vital heart_us.py:107-167) per bit-packed mask: masks (..., 2048) int32 ->
    (diameters (..., 20) fp64 mm, step (...) fp64); NaN where the reference raises.  ``spacing``: a float (isotropic) or
    (height, width) voxel sizes -- anisotropic masks are seen through their nearest-neighbour stretch to isotropic
    spacing, as the reference's PIL resize produces it."""
    masks = nat.require_cuda(masks, torch.int32, "masks")
    if masks.shape[-1] != MASK_WORDS:
        raise ValueError(f"masks must be (..., {MASK_WORDS}) int32 bit-packed")
    vs = (float(spacing), float(spacing)) if np.isscalar(spacing) else (float(spacing[0]), float(spacing[1]))
    if not (vs[0] > 0 and vs[1] > 0):
        raise ValueError("voxel spacing must be positive")
    nh, nw, ri, ci, iso = isotropic_view(vs)
    lead = tuple(masks.shape[:-1])
    n = int(np.prod(lead)) if lead else 1
    dev = masks.device
    diam = torch.empty(lead + (20,), dtype=torch.float64, device=dev)
    step = torch.empty(lead, dtype=torch.float64, device=dev)
    npts = torch.empty(lead, dtype=torch.int32, device=dev) if return_npts else None
    ws = torch.empty((max(int(nat.lib().cub200_simpson_workspace_bytes(n)), 1),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        if ri is None:
            rc = nat.lib().cub200_simpson_diameters(nat.ptr(masks), n, iso, nat.ptr(diam), nat.ptr(step), nat.ptr(npts),
                                                    nat.ptr(ws), nat.stream_ptr(dev))
        else:
            if max(nh, nw) > 1024:
                raise ValueError("voxel spacing ratio too large: the isotropic view would exceed 1024 pixels")
            rit = torch.from_numpy(ri.astype(np.int16)).to(dev)
            cit = torch.from_numpy(ci.astype(np.int16)).to(dev)
            rc = nat.lib().cub200_simpson_diameters_stretched(nat.ptr(masks), n, nat.ptr(rit), nat.ptr(cit), nh, nw, iso, nat.ptr(diam),
                                                              nat.ptr(step), nat.ptr(npts), nat.ptr(ws), nat.stream_ptr(dev))
    nat.check(rc, "cub200_simpson_diameters")
    return (diam, step, npts) if return_npts else (diam, step)


def simpson_volumes(diam: torch.Tensor, step: torch.Tensor):
    """compute_left_ventricle_volumes + the Monte-Carlo loop of Volume.compute (heart_us.py:13-71,
    patient.py:166-190): diam (P,4,T,20), step (P,4,T) [A2C-ED, A2C-ES, A4C-ED, A4C-ES] -> edv, esv, ef (P,T) fp64."""
    diam = nat.require_cuda(diam, torch.float64, "diam")
    step = nat.require_cuda(step, torch.float64, "step")
    if diam.dim() != 4 or diam.shape[1] != 4 or diam.shape[3] != 20 or tuple(step.shape) != tuple(diam.shape[:3]):
        raise ValueError("diam (P,4,T,20) / step (P,4,T) expected")
    P_, T = diam.shape[0], diam.shape[2]
    out = [torch.empty((P_, T), dtype=torch.float64, device=diam.device) for _ in range(3)]
    with torch.cuda.device(diam.device):
        rc = nat.lib().cub200_simpson_volumes(nat.ptr(diam), nat.ptr(step), P_, T, nat.ptr(out[0]), nat.ptr(out[1]), nat.ptr(out[2]),
                                              nat.stream_ptr(diam.device))
    nat.check(rc, "cub200_simpson_volumes")
    return tuple(out)


def biplane_volumes_from_contours(contours: torch.Tensor, spacing, spacing_a4c=None):
    """contours (P,4,T,21,2) fp32 [A2C-ED, A2C-ES, A4C-ED, A4C-ES] -> (edv, esv, ef (P,T), area (P,4,T)).
    Rasterise every sampled contour (bit-packed masks), then Simpson's biplane rule per sample.  ``spacing`` (float or
    (height, width)) is the A2C voxel spacing, ``spacing_a4c`` the A4C one (default: the same)."""
    contours = nat.require_cuda(contours, torch.float32, "contours")
    if contours.dim() != 5 or contours.shape[1] != 4 or contours.shape[3:] != (K, 2):
        raise ValueError("contours (P,4,T,21,2) expected")
    P_, _, T = contours.shape[:3]
    masks, area = rasterize_contours(contours.reshape(-1, K, 2), return_masks=True)
    masks = masks.reshape(P_, 4, T, MASK_WORDS)
    if spacing_a4c is None:
        diam, step = simpson_diameters(masks, spacing)
    else:
        d2, s2 = simpson_diameters(masks[:, :2].contiguous(), spacing)
        d4, s4 = simpson_diameters(masks[:, 2:].contiguous(), spacing_a4c)
        diam, step = torch.cat([d2, d4], 1), torch.cat([s2, s4], 1)
    edv, esv, ef = simpson_volumes(diam, step)
    return edv, esv, ef, area.reshape(P_, 4, T)


# --------------------------------------------------------------------------------------------------
# Analytic uncertainty maps
# --------------------------------------------------------------------------------------------------
_SQRT_2PI = np.sqrt(2 * np.pi)
# norm.pdf(np.linspace(-2, 2, 100))  (cu/utils/umap.py:21,29; scipy: exp(-x**2/2) / sqrt(2 pi))
UMAP_PDF_TABLE = np.exp(-np.linspace(-2, 2, 100) ** 2 / 2.0) / _SQRT_2PI
# norm.pdf(i, loc=0, scale=50) for i = 0..99, laid out as skew_umap's umap_weights (cu/utils/skew_umap.py:50-51)
_w = np.exp(-(np.arange(100) / 50.0) ** 2 / 2.0) / _SQRT_2PI / 50.0
SKEW_UMAP_WEIGHTS = np.concatenate([_w[::-1], _w])
# the denominator of np.average(rec, axis=0, weights=w): the weights broadcast to (200,1,1) and summed over axis 0
SKEW_UMAP_WEIGHT_SUM = float(np.broadcast_to(SKEW_UMAP_WEIGHTS, (1, 1, 200)).swapaxes(-1, 0).sum(axis=0)[0, 0])


def projected_uncertainty(mu: torch.Tensor, cov: torch.Tensor, alpha: Optional[torch.Tensor] = None, all: bool = False,
                          linear_close: bool = False):
    """cu/utils/uncertainty_projection.py:17-125 for a batch: mu (N,P,2), cov (N,P,2,2) [, alpha (N,P,2)] fp32 ->
    (uncertainties (N,P) fp32, projections (N,P,2) fp64 [, alpha_proj (N,P) fp32]).  P = 21 (the LV contour of the hot path) or
    any 4 <= P <= 64 (the lung / heart contours of cu/data/lung/utils.py)."""
    mu = nat.require_cuda(mu, torch.float32, "mu")
    cov = nat.require_cuda(cov, torch.float32, "cov")
    if mu.dim() != 3 or mu.shape[2] != 2 or not (4 <= mu.shape[1] <= RASTER_MAX_POINTS) or tuple(cov.shape) != tuple(mu.shape) + (2,):
        raise ValueError(f"mu (N,P,2) / cov (N,P,2,2) with 4 <= P <= {RASTER_MAX_POINTS} expected")
    n, p, dev = mu.shape[0], int(mu.shape[1]), mu.device
    ap = None
    if alpha is not None:
        alpha = nat.require_cuda(alpha, torch.float32, "alpha")
        if alpha.shape != mu.shape:
            raise ValueError("alpha must have the shape of mu")
        ap = torch.empty((n, p), dtype=torch.float32, device=dev)
    unc = torch.empty((n, p), dtype=torch.float32, device=dev)
    proj = torch.empty((n, p, 2), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        if p == K:
            rc = nat.lib().cub200_projected_uncertainty(nat.ptr(mu), nat.ptr(cov), nat.ptr(alpha), n, int(all), int(linear_close),
                                                        nat.ptr(unc), nat.ptr(proj), nat.ptr(ap), nat.stream_ptr(dev))
        else:
            rc = nat.lib().cub200_projected_uncertainty_generic(nat.ptr(mu), nat.ptr(cov), nat.ptr(alpha), n, p, int(all), int(linear_close),
                                                                nat.ptr(unc), nat.ptr(proj), nat.ptr(ap), nat.stream_ptr(dev))
    nat.check(rc, "cub200_projected_uncertainty")
    return (unc, proj) if alpha is None else (unc, proj, ap)


def umap_analytic(mu: torch.Tensor, cov: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """uncertainty_map (cu/utils/umap.py:10-33) [/ max = USUMap, cu/data/camus/utils.py:113-115]: (N,256,256) fp64."""
    unc, proj = projected_uncertainty(mu, cov, all=True)
    n, dev = mu.shape[0], mu.device
    table = torch.from_numpy(UMAP_PDF_TABLE).to(dev)
    kmap = torch.empty((n, HW, HW), dtype=torch.int32, device=dev)
    out = torch.empty((n, HW, HW), dtype=torch.float64, device=dev)
    mu = nat.require_cuda(mu, torch.float32, "mu")
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_umap_analytic(nat.ptr(mu), nat.ptr(unc), nat.ptr(proj), n, nat.ptr(table), int(normalize),
                                            nat.ptr(kmap), nat.ptr(out), nat.stream_ptr(dev))
    nat.check(rc, "cub200_umap_analytic")
    return out


def skew_umap(mu: torch.Tensor, cov: torch.Tensor, alpha: torch.Tensor, linear_close: bool = True, normalize: bool = True,
              return_contours: bool = False):
    """skew_umap (cu/utils/skew_umap.py:11-82) [USSkewUmap: linear_close=True, / max]: -> (projected_mode (N,P,2) fp32,
    umap (N,256,256) fp64).  200 level contours per image are reconstructed (unrounded, fp64) by the rasteriser.
    P = 21, or any 4 <= P <= 64 (LungUMap / LungSkewUmap call it per structure with 44, 50 and 26 points)."""
    unc, proj, ap = projected_uncertainty(mu, cov, alpha, all=True, linear_close=linear_close)
    mu = nat.require_cuda(mu, torch.float32, "mu")
    n, p, dev = mu.shape[0], int(mu.shape[1]), mu.device
    contours = torch.empty((n, 200, p, 2), dtype=torch.float64, device=dev)
    mode = torch.empty((n, p, 2), dtype=torch.float32, device=dev)
    L = nat.lib()
    with torch.cuda.device(dev):
        nat.check(L.cub200_skew_umap_contours(nat.ptr(mu), nat.ptr(unc), nat.ptr(proj), nat.ptr(ap), n, p, nat.ptr(contours),
                                              nat.ptr(mode), nat.stream_ptr(dev)), "cub200_skew_umap_contours")
    if p == K:
        masks, _ = rasterize_contours(contours.reshape(-1, K, 2), round_landmarks=False, return_masks=True)
    else:
        masks, _ = rasterize_contours_generic(contours.reshape(-1, p, 2), round_landmarks=False, return_masks=True)
    w = torch.from_numpy(SKEW_UMAP_WEIGHTS).to(dev)
    out = torch.empty((n, HW, HW), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        nat.check(L.cub200_weighted_mask_entropy(nat.ptr(masks), nat.ptr(w), n, 200, SKEW_UMAP_WEIGHT_SUM, int(normalize),
                                                 nat.ptr(out), nat.stream_ptr(dev)), "cub200_weighted_mask_entropy")
    return (mode, out, contours) if return_contours else (mode, out)


# --------------------------------------------------------------------------------------------------
# Dataset-level pixel calibration (results processor reduction)
# --------------------------------------------------------------------------------------------------
def contour_perimeter(contours: torch.Tensor) -> torch.Tensor:
    """perimeter() (cu/utils/clinical.py:30-47): contours (..., 21, 2) fp32 -> spline arc lengths (...) fp64."""
    contours = nat.require_cuda(contours, torch.float32, "contours")
    if contours.dim() < 2 or contours.shape[-2:] != (K, 2):
        raise ValueError(f"contours must be (..., {K}, 2)")
    lead = tuple(contours.shape[:-2])
    n = int(np.prod(lead)) if lead else 1
    out = torch.empty(lead, dtype=torch.float64, device=contours.device)
    with torch.cuda.device(contours.device):
        rc = nat.lib().cub200_contour_perimeter(nat.ptr(contours), n, nat.ptr(out), nat.stream_ptr(contours.device))
    nat.check(rc, "cub200_contour_perimeter")
    return out


def global_longitudinal_strain(ed_contours: torch.Tensor, es_contours: torch.Tensor) -> torch.Tensor:
    """global_longitudinal_strain (cu/utils/clinical.py:50-71) per sample: (len_ED - len_ES) / len_ED, fp64."""
    led, les = contour_perimeter(ed_contours), contour_perimeter(es_contours)
    return (led - les) / led


def calibration_bins(umap: torch.Tensor, pred: torch.Tensor, gt: torch.Tensor, nb_bins: int = 20) -> torch.Tensor:
    """Binned (count, sum confidence, sum accuracy) over the foreground pixels (cu/results/calibration.py:94-113,197-250):
    umap fp64, pred / gt uint8, any equal shapes -> (nb_bins + 1, 3) fp64; row nb_bins = (foreground pixel count, 0, 0).
    The rows are additive: ranks all_reduce them before ``ece_from_bins``."""
    umap = nat.require_cuda(umap, torch.float64, "umap")
    pred = nat.require_cuda(pred, torch.uint8, "pred")
    gt = nat.require_cuda(gt, torch.uint8, "gt")
    if pred.shape != umap.shape or gt.shape != umap.shape:
        raise ValueError("umap / pred / gt shapes differ")
    dev = umap.device
    bounds = torch.from_numpy(np.linspace(0, 1, nb_bins + 1)).to(dev)
    out = torch.empty((nb_bins + 1, 3), dtype=torch.float64, device=dev)
    ws = torch.empty((int(nat.lib().cub200_calibration_workspace_bytes()),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_calibration_bins(nat.ptr(umap), nat.ptr(pred), nat.ptr(gt), umap.numel(), nat.ptr(bounds), int(nb_bins),
                                               nat.ptr(out), nat.ptr(ws), nat.stream_ptr(dev))
    nat.check(rc, "cub200_calibration_bins")
    return out


def ece_from_bins(bins) -> tuple:
    """(ece, mce) from calibration bins as ``Calibration.ece`` forms them (cu/results/calibration.py:94-117)."""
    b = np.asarray(bins.cpu() if isinstance(bins, torch.Tensor) else bins, dtype=np.float64)
    total = b[-1, 0]
    ece, gaps = 0.0, []
    for count, s_conf, s_acc in b[:-1]:
        if count > 0 and count / total > 0:
            gap = abs(s_conf / count - s_acc / count)
            ece += gap * (count / total)
            gaps.append(gap)
    return ece, (max(gaps) if gaps else float("nan"))


_GLS_TABLES = {}
GLS_TABLE_STRIDE = 48
SIMPSON_MAX_POINTS = 2048


def gls_gauss_table(polar_smoothing_factor: float, device) -> torch.Tensor:
    """Row n = the normalised Gaussian kernel values w[0..lw] of scipy.ndimage.gaussian_filter1d(sigma = n * factor),
    formed exactly as scipy's _gaussian_kernel1d forms them (numpy exp, numpy sum); cached per (factor, device)."""
    key = (float(polar_smoothing_factor), str(device))
    if key not in _GLS_TABLES:
        tab = np.zeros((SIMPSON_MAX_POINTS + 1, GLS_TABLE_STRIDE))
        tab[:, 0] = 1.0
        for n in range(1, SIMPSON_MAX_POINTS + 1):
            sigma = n * float(polar_smoothing_factor)
            lw = int(4.0 * sigma + 0.5)
            if lw == 0 or sigma <= 0:
                continue
            x = np.arange(-lw, lw + 1)
            phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
            phi = phi / phi.sum()
            tab[n, :lw + 1] = phi[lw:]
        _GLS_TABLES[key] = torch.from_numpy(tab).to(device)
    return _GLS_TABLES[key]


def mask_lv_length(masks: torch.Tensor, voxelspacing=(1.0, 1.0), polar_smoothing_factor: float = 5e-3, return_corners: bool = False):
    """Longitudinal length of the LV endocardium of every mask (EchoMeasure._lv_longitudinal_length, myo_labels=None,
    vital/vital/utils/image/us/measure.py:339-354): bit-packed masks (..., 2048) int32 -> lengths (...) fp64, NaN where the
    reference raises [, base corners (..., 2, 2) as (row, col), points of the base-less contour (...) int32]."""
    masks = nat.require_cuda(masks, torch.int32, "masks")
    if masks.shape[-1] != MASK_WORDS:
        raise ValueError(f"masks must be (..., {MASK_WORDS}) bit-packed")
    lead = tuple(masks.shape[:-1])
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    dev = masks.device
    length = torch.empty(lead, dtype=torch.float64, device=dev)
    corners = torch.empty(lead + (2, 2), dtype=torch.float64, device=dev) if return_corners else None
    npts = torch.empty(lead, dtype=torch.int32, device=dev) if return_corners else None
    ws = torch.empty((max(int(nat.lib().cub200_simpson_workspace_bytes(n)), 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        table = gls_gauss_table(polar_smoothing_factor, dev) if polar_smoothing_factor > 0 else None
        rc = nat.lib().cub200_mask_lv_length(nat.ptr(masks), n, float(polar_smoothing_factor), nat.ptr(table), float(voxelspacing[0]),
                                             float(voxelspacing[1]), nat.ptr(length), nat.ptr(corners), nat.ptr(npts), nat.ptr(ws),
                                             nat.stream_ptr(dev))
    nat.check(rc, "cub200_mask_lv_length")
    return (length, corners, npts) if return_corners else length


def endo_base_corners(lv_masks: torch.Tensor, myo_masks: torch.Tensor) -> torch.Tensor:
    """EchoMeasure._endo_base (vital/vital/utils/image/us/measure.py:177-222): bit-packed LV and MYO masks (..., 2048) int32 ->
    (..., 2, 2) int32 = (left, right) base markers as (row, col); -1 where the reference raises (< 2 frontier pixels)."""
    lv_masks = nat.require_cuda(lv_masks, torch.int32, "lv_masks")
    myo_masks = nat.require_cuda(myo_masks, torch.int32, "myo_masks")
    if lv_masks.shape[-1] != MASK_WORDS or myo_masks.shape != lv_masks.shape or myo_masks.device != lv_masks.device:
        raise ValueError(f"lv_masks and myo_masks must be (..., {MASK_WORDS}) bit-packed, same shape and device")
    lead = tuple(lv_masks.shape[:-1])
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    dev = lv_masks.device
    corners = torch.empty(lead + (2, 2), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_endo_base_corners(nat.ptr(lv_masks), nat.ptr(myo_masks), n, nat.ptr(corners), nat.stream_ptr(dev))
    nat.check(rc, "cub200_endo_base_corners")
    return corners


def mask_lv_length_myo(lv_masks: torch.Tensor, myo_masks: torch.Tensor, voxelspacing=(1.0, 1.0), return_corners: bool = False):
    """EchoMeasure._lv_longitudinal_length with myo_labels (measure.py:339-354, base markers by _endo_base :177-222):
    bit-packed LV and MYO masks (..., 2048) int32 -> lengths (...) fp64, NaN where the reference raises
    [, base markers (..., 2, 2) int32 as (row, col), points of the base-less contour (...) int32]."""
    lv_masks = nat.require_cuda(lv_masks, torch.int32, "lv_masks")
    myo_masks = nat.require_cuda(myo_masks, torch.int32, "myo_masks")
    if lv_masks.shape[-1] != MASK_WORDS or myo_masks.shape != lv_masks.shape or myo_masks.device != lv_masks.device:
        raise ValueError(f"lv_masks and myo_masks must be (..., {MASK_WORDS}) bit-packed, same shape and device")
    lead = tuple(lv_masks.shape[:-1])
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    dev = lv_masks.device
    length = torch.empty(lead, dtype=torch.float64, device=dev)
    corners = torch.empty(lead + (2, 2), dtype=torch.int32, device=dev)
    npts = torch.empty(lead, dtype=torch.int32, device=dev) if return_corners else None
    ws = torch.empty((max(int(nat.lib().cub200_simpson_workspace_bytes(n)), 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_mask_lv_length_myo(nat.ptr(lv_masks), nat.ptr(myo_masks), n, float(voxelspacing[0]), float(voxelspacing[1]),
                                                 nat.ptr(length), nat.ptr(corners), nat.ptr(npts), nat.ptr(ws), nat.stream_ptr(dev))
    nat.check(rc, "cub200_mask_lv_length_myo")
    return (length, corners, npts) if return_corners else length


def mask_gls(masks: torch.Tensor, voxelspacing=(1.0, 1.0), myo_masks: Optional[torch.Tensor] = None) -> torch.Tensor:
    """EchoMeasure.gls (measure.py:320-366) on sequences of bit-packed masks (..., N_frames, 2048): strain in
    percent of every frame against the first one (the ED frame), (..., N_frames) fp64; a frame the reference cannot measure
    makes the whole sequence NaN (its exception leaves GLS.gls's try block, cu/results/clinical/view.py:86-96).
    myo_masks None: base corners from the polar contour (myo_labels=None); else EchoMeasure._endo_base on (LV, MYO)."""
    length = mask_lv_length(masks, voxelspacing) if myo_masks is None else mask_lv_length_myo(masks, myo_masks, voxelspacing)
    ed = length[..., :1]
    gls = (length - ed) / ed * 100
    return torch.where(torch.isnan(length).any(-1, keepdim=True), torch.full_like(gls, float("nan")), gls)


# --------------------------------------------------------------------------------------------------
# Segmentation-sample post-processing
# --------------------------------------------------------------------------------------------------
def seg_largest_blob(probs: torch.Tensor, return_area: bool = False):
    """Hole-filled largest blob of every rounded sample (segmentation_uncertainty.py:104-113, utils.py:12-32):
    probs (..., 256, 256) fp32 -> bit-packed masks (..., 2048) int32 [, areas (...) int32]."""
    probs = nat.require_cuda(probs, torch.float32, "probs")
    if probs.ndim < 2 or tuple(probs.shape[-2:]) != (256, 256):
        raise ValueError("probs must be (..., 256, 256)")
    lead = tuple(probs.shape[:-2])
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    dev = probs.device
    masks = torch.empty(lead + (MASK_WORDS,), dtype=torch.int32, device=dev)
    area = torch.empty(lead, dtype=torch.int32, device=dev) if return_area else None
    ws = torch.empty((int(nat.lib().cub200_seg_blob_workspace_bytes(n)),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_seg_largest_blob(nat.ptr(probs), n, nat.ptr(masks), nat.ptr(area), nat.ptr(ws), nat.stream_ptr(dev))
    nat.check(rc, "cub200_seg_largest_blob")
    return (masks, area) if return_area else masks


def seg_entropy(probs: torch.Tensor, masks: torch.Tensor, pad: int = 10, return_mean: bool = False,
                return_pred_samples: bool = False, return_samples: bool = False) -> dict:
    """Entropy map of the post-processed samples (segmentation_uncertainty.py:115-126, uncertainty.py:107-133):
    probs (N, T, 256, 256) fp32 + masks (N, T, 2048) -> dict(entropy (N,256,256) fp32 [, mean, pred_samples, samples])."""
    probs = nat.require_cuda(probs, torch.float32, "probs")
    masks = nat.require_cuda(masks, torch.int32, "masks")
    if probs.ndim != 4 or tuple(probs.shape[-2:]) != (256, 256):
        raise ValueError("probs must be (N, T, 256, 256)")
    n, T = int(probs.shape[0]), int(probs.shape[1])
    if tuple(masks.shape) != (n, T, MASK_WORDS):
        raise ValueError(f"masks must be ({n}, {T}, {MASK_WORDS})")
    if T == 0:
        raise ValueError("no samples")
    dev = probs.device
    out = {"entropy": torch.empty((n, 256, 256), dtype=torch.float32, device=dev)}
    if return_mean:
        out["mean"] = torch.empty((n, 256, 256), dtype=torch.float32, device=dev)
    if return_pred_samples:
        out["pred_samples"] = torch.empty((n, T, MASK_WORDS), dtype=torch.int32, device=dev)
    if return_samples:
        out["samples"] = torch.empty((n, T, 256, 256), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().cub200_seg_entropy(nat.ptr(probs), nat.ptr(masks), n, T, int(pad), nat.ptr(out["entropy"]),
                                          nat.ptr(out.get("mean")), nat.ptr(out.get("pred_samples")), nat.ptr(out.get("samples")),
                                          nat.stream_ptr(dev))
    nat.check(rc, "cub200_seg_entropy")
    return out


# --------------------------------------------------------------------------------------------------
# End-to-end engine
# --------------------------------------------------------------------------------------------------
@dataclass
class UncertaintyResult:
    mu: torch.Tensor              # (N,21,2) pixel coordinates
    cov: torch.Tensor             # (N,21,2,2)
    contour_samples: Optional[torch.Tensor]  # (N,T,21,2)
    area: torch.Tensor            # (N,T) int32 LV pixel counts
    count: Optional[torch.Tensor]  # (N,256,256) int32
    mean: torch.Tensor            # (N,256,256) fp64
    entropy: torch.Tensor         # (N,256,256) fp64, bits
    pred: torch.Tensor            # (N,256,256) uint8
    masks: Optional[torch.Tensor] = None  # (N,T,2048) int32 bit-packed


class Workspace:
    """Preallocated device buffers for ``UncertaintyEngine.run_into`` (no allocation per step)."""

    def __init__(self, n_img: int, t_a: int, zdim: int, device, want_count: bool = True, want_contours: bool = True,
                 want_maps: bool = True, want_variance: bool = False, want_count16: bool = False):
        dev = torch.device(device)
        f32, i32, f64 = torch.float32, torch.int32, torch.float64
        self.n_img, self.t_a, self.zdim, self.device = n_img, t_a, zdim, dev
        self.mu = torch.empty((n_img, K, 2), dtype=f32, device=dev)
        self.cov = torch.empty((n_img, K, 2, 2), dtype=f32, device=dev)
        self.a_t = torch.empty((n_img, zdim, P), dtype=f32, device=dev)
        self.contours = torch.empty((n_img, t_a, K, 2), dtype=f32, device=dev)   # always: the sampler -> rasteriser hand-over
        self.area = torch.empty((n_img, t_a), dtype=i32, device=dev)
        self.diff = torch.empty((n_img, HW, HW), dtype=i32, device=dev)
        self.count = torch.empty((n_img, HW, HW), dtype=i32, device=dev) if want_count else None
        self.mean = torch.empty((n_img, HW, HW), dtype=f64, device=dev) if want_maps else None
        self.entropy = torch.empty((n_img, HW, HW), dtype=f64, device=dev) if want_maps else None
        self.variance = torch.empty((n_img, HW, HW), dtype=f64, device=dev) if want_variance else None
        self.count16 = torch.empty((n_img, HW, HW), dtype=torch.uint16, device=dev) if want_count16 else None
        self.pred = torch.empty((n_img, HW, HW), dtype=torch.uint8, device=dev)
        self.lut = torch.empty((3 * (t_a + 1),), dtype=f64, device=dev)       # entropy | mean | variance of c / t_a: filled once, here
        with torch.cuda.device(dev):
            nat.check(nat.lib().cub200_count_tables(int(t_a), self.lut.data_ptr(), nat.stream_ptr(dev)), "cub200_count_tables")
        self.raster_ws = raster_workspace(n_img * t_a, dev)
        self.dsnt_ws = dsnt_workspace(n_img * K, dev)

    def result(self) -> "UncertaintyResult":
        return UncertaintyResult(self.mu, self.cov, self.contours, self.area, self.count, self.mean, self.entropy, self.pred)


class UncertaintyEngine:
    """heatmap logits -> DSNT -> sampler factor -> fused sample/raster/aggregate -> maps.

    The data-parallel hot path of ``AleatoricUncertaintyTask._predict_step`` + ``predict_step``
    (aleatoric.py:80-135, contour_uncertainty.py:71-82) for T_e = 1.
    """

    def __init__(self, psm_path: str = DEFAULT_PSM, device="cuda", sampler: str = "psm", covar: bool = True):
        if sampler not in ("psm", "naive"):
            raise ValueError("sampler must be 'psm' or 'naive'")
        self.device = torch.device(device)
        self.sampler = sampler
        self.covar = covar
        self.model = ShapeModelConstants(psm_path, self.device) if sampler == "psm" else None
        nat.lib()  # fail now if the extension is missing

    def factor(self, mu, cov):
        return psm_factor(self.model, mu, cov) if self.sampler == "psm" else naive_factor(cov)

    def from_moments(self, mu, cov, t_a: int, z=None, seed: int = 0, return_contours=True, return_masks=False,
                     want_count=True) -> UncertaintyResult:
        a_t = self.factor(mu, cov)
        r = sample_raster_aggregate(mu, a_t, t_a, z=z, seed=seed, return_contours=return_contours,
                                    return_masks=return_masks)
        count, mean, ent, pred = finalize_maps(r["diff"], t_a, want_count=want_count)
        return UncertaintyResult(mu, cov, r["contours"], r["area"], count, mean, ent, pred, r["masks"])

    def workspace(self, n_img: int, t_a: int, **kw) -> Workspace:
        return Workspace(n_img, t_a, Z_PSM if self.sampler == "psm" else Z_NAIVE, self.device, **kw)

    def run_into(self, ws: Workspace, logits: torch.Tensor, seed: int = 0, stream: Optional[torch.cuda.Stream] = None,
                 sample_offset: int = 0, step_dev: Optional[torch.Tensor] = None, after_raster=None):
        """The whole path into preallocated buffers: 7 launches (+1 memset), no allocation, no sync.

        ``logits`` must be (ws.n_img,21,256,256) fp32 CUDA contiguous.  ``sample_offset``: index of the first sample of this
        call in a larger batch (chunked callers), see ``sample_raster_aggregate``.  ``step_dev``: device step counter (one
        int64) that shifts the draws on the device -- for steps replayed from a CUDA graph (``GraphedStep``).
        ``after_raster``: callable invoked once the sampler / rasteriser launches are queued (contours and areas final), before
        finalize_maps: the place to fork work that only needs the per-sample outputs.  Returns ``ws``.
        """
        self.dsnt_into(ws, logits, ws.mu, ws.cov, stream=stream)
        return self.run_tail_into(ws, ws.mu, ws.cov, seed=seed, stream=stream, sample_offset=sample_offset, step_dev=step_dev,
                                  after_raster=after_raster)

    def dsnt_into(self, ws: Workspace, logits: torch.Tensor, mu: torch.Tensor, cov: torch.Tensor,
                  stream: Optional[torch.cuda.Stream] = None, persistent: bool = True):
        """First launch of ``run_into``: the DSNT moments of ``logits`` into ``mu`` (n_img,21,2) / ``cov`` (n_img,21,2,2) fp32
        (``ws.mu`` / ``ws.cov`` or buffers of the same shape: a pipelined caller keeps two pairs).  ``persistent=False``: one CTA
        per heatmap instead of one wave of equal row ranges (``cub200_dsnt_moments`` without a workspace) -- for a launch that
        shares the GPU with other kernels.  A heatmap that the persistent launch splits between two CTAs is summed in another
        order there: the two agree to fp32 rounding of the moments (~1e-7 relative), not bit for bit."""
        if logits.dtype != torch.float32 or not logits.is_cuda or not logits.is_contiguous() or \
                tuple(logits.shape) != (ws.n_img, K, HW, HW):
            raise ValueError(f"logits must be ({ws.n_img},{K},{HW},{HW}) fp32 CUDA contiguous")
        self._check_moments(ws, mu, cov)
        st = (stream or torch.cuda.current_stream(ws.device)).cuda_stream
        with torch.cuda.device(ws.device):
            nat.check(nat.lib().cub200_dsnt_moments(logits.data_ptr(), ws.n_img * K, HW, HW, 1, int(self.covar), None, mu.data_ptr(),
                                                    cov.data_ptr(), ws.dsnt_ws.data_ptr() if persistent else None, st), "cub200_dsnt_moments")

    @staticmethod
    def _check_moments(ws: Workspace, mu: torch.Tensor, cov: torch.Tensor):
        for t, ref, name in ((mu, ws.mu, "mu"), (cov, ws.cov, "cov")):
            if t.dtype != ref.dtype or t.device != ref.device or t.shape != ref.shape or not t.is_contiguous():
                raise ValueError(f"{name} must be {tuple(ref.shape)} {ref.dtype} contiguous on {ref.device}")

    def run_tail_into(self, ws: Workspace, mu: torch.Tensor, cov: torch.Tensor, seed: int = 0,
                      stream: Optional[torch.cuda.Stream] = None, sample_offset: int = 0,
                      step_dev: Optional[torch.Tensor] = None, after_raster=None):
        """``run_into`` after its DSNT launch: sampler factor, sampling + rasterisation + aggregation, maps, from the moments
        in ``mu`` / ``cov``."""
        self._check_moments(ws, mu, cov)
        L = nat.lib()
        st = (stream or torch.cuda.current_stream(ws.device)).cuda_stream
        n, T = ws.n_img, ws.t_a
        with torch.cuda.device(ws.device):
            if self.sampler == "psm":
                m = self.model
                nat.check(L.cub200_psm_factor(m.pack.data_ptr(), mu.data_ptr(), cov.data_ptr(), n, ws.a_t.data_ptr(), st),
                          "cub200_psm_factor")
            else:
                nat.check(L.cub200_naive_factor(cov.data_ptr(), n, ws.a_t.data_ptr(), st), "cub200_naive_factor")
            if stream is None:
                ws.diff.zero_()
            else:
                with torch.cuda.stream(stream):
                    ws.diff.zero_()
            nat.check(L.cub200_sample_raster_aggregate(mu.data_ptr(), ws.a_t.data_ptr(), ws.zdim, None,
                                                       int(seed) & (2**64 - 1), int(sample_offset) & (2**64 - 1), n, T, 1,
                                                       ws.contours.data_ptr(),
                                                       ws.area.data_ptr(), ws.diff.data_ptr(), 1, None, ws.raster_ws.data_ptr(),
                                                       nat.ptr(step_dev), st),
                      "cub200_sample_raster_aggregate")
            if after_raster is not None:
                after_raster()
            nat.check(L.cub200_finalize_maps(ws.diff.data_ptr(), n, T, nat.ptr(ws.count), nat.ptr(ws.mean),
                                             nat.ptr(ws.entropy), ws.pred.data_ptr(), ws.lut.data_ptr(), 1, nat.ptr(ws.variance),
                                             nat.ptr(ws.count16), st),
                      "cub200_finalize_maps")
        return ws

    def __call__(self, logits: torch.Tensor, t_a: int, z=None, seed: int = 0, return_contours=True,
                 return_masks=False, want_count=True) -> UncertaintyResult:
        _, mu, cov = dsnt_moments(logits, is_logit=True, covar=self.covar)
        return self.from_moments(mu, cov, t_a, z=z, seed=seed, return_contours=return_contours,
                                 return_masks=return_masks, want_count=want_count)


METRIC_PACK_PARTS = 16       # CUB200_METRIC_PACK_PARTS


class GraphedStep:
    """The whole device-resident step (``UncertaintyEngine.run_into`` [+ the per-step metric buffering of the epoch-end
    exchange]) captured ONCE in a CUDA graph and replayed with one launch per step.

    Eager, a step is 8 launches (+ 4 for the exchange buffer): on a box where several ranks share the host that is where the
    multi-GPU step loses time (profiles/r02d_*: kernel times identical at 1 and 8 GPUs, the step 0.44 -> 0.50 ms; replayed
    from a graph all ranks run at the single-GPU rate).  A captured launch's arguments are frozen, so what changes from step
    to step lives on the device: ``self.step`` (one int64) shifts the Philox sample indices of the sampler and selects the
    slot of the metric ring; the graph's last node increments it.  Step i of a GraphedStep draws the samples
    ``UncertaintyEngine.run_into(..., seed, sample_offset=i * n_img * t_a)`` draws.
    ``logits`` is read in place at every replay: refill the same buffer (``logits.copy_``) between replays.
    """

    def __init__(self, engine: "UncertaintyEngine", ws: Workspace, logits: torch.Tensor, seed: int = 0,
                 metric_ring: Optional[torch.Tensor] = None, step: Optional[torch.Tensor] = None, prefetch=None):
        """``step``: an existing device counter to share (several GraphedSteps over different input buffers that advance one
        sequence of steps); by default the object owns a fresh one.  Capturing leaves the counter at 0.

        ``prefetch`` = (next_logits, (mu, cov), (mu_next, cov_next)): SOFTWARE-PIPELINED step for a stream of batches.  The
        moments of THIS step's batch are expected in (mu, cov) already -- left there by the previous step's replay, or by
        ``engine.dsnt_into`` for the first one -- and the graph runs the DSNT pass of ``next_logits`` into (mu_next, cov_next)
        on a side branch beside this batch's factor / sampler / rasteriser / finalize: the HBM-bound kernel streams while the
        issue-bound ones compute (tools/overlap_probe.py).  ``logits`` is then only the record of which batch this step is."""
        self.engine, self.ws, self.logits, self.seed, self.ring = engine, ws, logits, int(seed), metric_ring
        self.prefetch = prefetch
        if prefetch is not None:
            nxt, (mu_c, cov_c), (mu_n, cov_n) = prefetch
            engine._check_moments(ws, mu_c, cov_c)
            engine._check_moments(ws, mu_n, cov_n)
            if mu_c.data_ptr() == mu_n.data_ptr() or cov_c.data_ptr() == cov_n.data_ptr():
                raise ValueError("prefetch: the next batch's moments need their own buffers")
        dev = ws.device
        if metric_ring is not None:
            metric_ring = nat.require_cuda(metric_ring, torch.int32, "metric_ring")
            if metric_ring.dim() != 2 or metric_ring.shape[1] != ws.area.numel() + 4 * METRIC_PACK_PARTS or metric_ring.device != dev:
                raise ValueError(f"metric_ring must be (n_slots, {ws.area.numel() + 4 * METRIC_PACK_PARTS}) int32 on {dev}")
        self.step = torch.zeros((1,), dtype=torch.int64, device=dev) if step is None else step
        if self.step.dtype != torch.int64 or self.step.numel() != 1 or self.step.device != dev:
            raise ValueError("step must be one int64 on the workspace's device")
        self.stream = torch.cuda.Stream(dev, priority=-1 if prefetch is not None else 0)
        self.side = torch.cuda.Stream(dev)
        self.ahead = torch.cuda.Stream(dev) if prefetch is not None else None
        self.graph = torch.cuda.CUDAGraph()
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            self._body()                                  # warm-up outside the capture (lazy attribute / module loads)
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        self.step.zero_()
        torch.cuda.synchronize(dev)
        n0 = nat.lib().cub200_launch_count()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self._body()
        self.kernels_per_replay = int(nat.lib().cub200_launch_count() - n0)   # library kernels one replay launches
        torch.cuda.synchronize(dev)

    def _body(self):
        L = nat.lib()
        st = self.stream.cuda_stream

        def fork_pack():
            # the metric buffer only needs the areas: it runs on a side branch of the graph, beside finalize_maps (a one-CTA
            # copy would otherwise hold the whole GPU for its ~8 us)
            if self.ring is None:
                return
            self.side.wait_stream(self.stream)
            with torch.cuda.device(self.ws.device):
                nat.check(L.cub200_metric_pack(self.ws.area.data_ptr(), self.ws.area.numel(), self.ring.data_ptr(),
                                               self.ring.shape[0], self.step.data_ptr(), self.side.cuda_stream), "cub200_metric_pack")

        if self.prefetch is None:
            self.engine.run_into(self.ws, self.logits, seed=self.seed, stream=self.stream, step_dev=self.step, after_raster=fork_pack)
        else:
            nxt, (mu_c, cov_c), (mu_n, cov_n) = self.prefetch
            self.ahead.wait_stream(self.stream)                                       # fork: the next batch's DSNT pass
            # one CTA per heatmap, not the persistent equal-range launch: that one assumes it owns every SM, and beside the
            # tail's kernels a third of its CTAs would start late (0.4325 vs 0.4230 ms per step)
            self.engine.dsnt_into(self.ws, nxt, mu_n, cov_n, stream=self.ahead, persistent=False)
            self.engine.run_tail_into(self.ws, mu_c, cov_c, seed=self.seed, stream=self.stream, step_dev=self.step,
                                      after_raster=fork_pack)
            self.stream.wait_stream(self.ahead)                                       # join
        if self.ring is not None:
            self.stream.wait_stream(self.side)
        with torch.cuda.device(self.ws.device):
            nat.check(L.cub200_step_counter_add(self.step.data_ptr(), 1, st), "cub200_step_counter_add")

    def reset(self, step: int = 0):
        """Set the device step counter."""
        self.step.fill_(int(step))

    def resync(self, slot_index: int, n_slots: int):
        """Make the next replay write ring slot ``slot_index`` without re-using sample indices: the counter jumps forward to
        the next value congruent to slot_index (mod n_slots).  Synchronises (reads the counter); for epoch boundaries."""
        cur = int(self.step.item())
        nxt = cur + ((int(slot_index) - cur) % int(n_slots))
        self.step.fill_(nxt)

    def replay(self):
        """One step on the CURRENT stream: one cudaGraphLaunch."""
        self.graph.replay()
        return self.ws


class HostPipeline:
    """Host-buffer entry of the path: pinned host logits in, pinned host results out.

    The batch is cut into chunks that flow through ``n_streams`` CUDA streams, so the host->device
    copy of chunk i+1 and the device->host copy of chunk i-1 overlap the kernels of chunk i (PCIe is
    the bottleneck of this entry: 5.5 MB of logits per image).  What the reference does at this
    boundary is the opposite copy (``S.cpu()``, dsnt_al.py:147-148) followed by seconds of Python.

    ``compact=True`` (needs t_a <= 65535) ships the uint16 COUNT map (128 KiB / image) instead of the fp64 mean and entropy
    maps and the uint8 pred (1.06 MiB / image): all three are functions of the count, ``expand()`` rebuilds them on the host
    bit for bit (``maps_from_count``).
    """

    def __init__(self, engine: UncertaintyEngine, n_img: int, t_a: int, chunk: int = 8, n_streams: int = 2, compact: bool = False):
        if n_img % chunk:
            raise ValueError("n_img must be a multiple of chunk")
        if compact and t_a > 65535:
            raise ValueError("compact results need t_a <= 65535")
        self.engine, self.n_img, self.t_a, self.chunk, self.compact = engine, n_img, t_a, chunk, compact
        self.out_keys = ("mu", "cov", "contours", "area", "count16") if compact else ("mu", "cov", "contours", "area", "mean", "entropy", "pred")
        dev = engine.device
        self.streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
        self.d_in = [torch.empty((chunk, K, HW, HW), dtype=torch.float32, device=dev) for _ in range(n_streams)]
        self.ws = [engine.workspace(chunk, t_a, want_count=False, want_maps=not compact, want_count16=compact) for _ in range(n_streams)]

    def host_outputs(self):
        n, T = self.n_img, self.t_a
        spec = {"mu": ((n, K, 2), torch.float32), "cov": ((n, K, 2, 2), torch.float32),
                "contours": ((n, T, K, 2), torch.float32), "area": ((n, T), torch.int32),
                "mean": ((n, HW, HW), torch.float64), "entropy": ((n, HW, HW), torch.float64),
                "pred": ((n, HW, HW), torch.uint8), "count16": ((n, HW, HW), torch.uint16)}
        return {k: torch.empty(*[spec[k][0]], dtype=spec[k][1], pin_memory=True) for k in self.out_keys}

    def expand(self, h_out: dict) -> dict:
        """mean / entropy / pred / variance maps of a compact result (host, exact)."""
        return maps_from_count(h_out["count16"].numpy(), self.t_a)

    def run(self, h_logits: torch.Tensor, h_out: dict, seed: int = 0):
        """Asynchronous: returns after enqueueing; the caller's current stream waits for the result."""
        if h_logits.is_cuda or tuple(h_logits.shape) != (self.n_img, K, HW, HW) or h_logits.dtype != torch.float32:
            raise ValueError(f"h_logits must be a host ({self.n_img},{K},{HW},{HW}) fp32 tensor (pinned for overlap)")
        cur = torch.cuda.current_stream(self.engine.device)
        for st in self.streams:
            st.wait_stream(cur)
        for ci, c0 in enumerate(range(0, self.n_img, self.chunk)):
            si = ci % len(self.streams)
            st, ws = self.streams[si], self.ws[si]
            with torch.cuda.stream(st):
                self.d_in[si].copy_(h_logits[c0:c0 + self.chunk], non_blocking=True)
                # one seed for the whole batch, the chunk's position as the Philox counter offset: the draws of an image
                # do not depend on the chunk size and no two (step, chunk) pairs share a stream
                self.engine.run_into(ws, self.d_in[si], seed=seed, stream=st, sample_offset=c0 * self.t_a)
                for k in self.out_keys:
                    h_out[k][c0:c0 + self.chunk].copy_(getattr(ws, k), non_blocking=True)
        for st in self.streams:
            cur.wait_stream(st)
        return h_out

"""Mirror of contour_uncertainty/data/config.py: BatchResult (:38-107), the output contract of the path, and the batch tags
the task classes read (ContourTags :11-35 < vital CamusTags < vital Tags, vital/vital/data/config.py:66-90,
vital/vital/data/camus/config.py:10-21,81-125)."""
from dataclasses import dataclass
from enum import IntEnum, unique
from typing import NamedTuple, Optional, Sequence, Tuple

import numpy as np


class LabelEnum(IntEnum):
    def __str__(self):  # vital/vital/data/config.py:12-16
        return self.name.lower()


@unique
class Label(LabelEnum):
    """vital/vital/data/camus/config.py:10-21."""
    BG = 0
    LV = 1
    MYO = 2
    ATRIUM = 3


@dataclass(frozen=True)
class Tags:
    id: str = "id"
    group: str = "group"
    neighbors: str = "neighbors"
    img: str = "img"
    gt: str = "gt"
    pred: str = "pred"
    post_pred: str = "post_pred"
    encoding: str = "z"


@dataclass(frozen=True)
class CamusTags(Tags):
    registered: str = "register"
    full_sequence: str = "sequence"
    instants: str = "instants"
    img_proc: str = "img_proc"
    gt_proc: str = "gt_proc"
    metadata: str = "metadata"
    voxelspacing: str = "voxelspacing"
    raw: str = "raw"
    post: str = "post"


@dataclass(frozen=True)
class ContourTags(CamusTags):
    contour: str = "contour"  # Contour GT (K,2)


class DataParameters(NamedTuple):
    """vital/vital/data/config.py:96-110."""
    in_shape: Optional[tuple] = None
    out_shape: Optional[tuple] = None
    labels: Optional[tuple] = None


@dataclass
class BatchResult:
    id: str
    img: np.ndarray  # [N (C), H, W]
    gt: np.ndarray  # [N, H, W]
    pred: np.ndarray  # [N, H, W]
    labels: Sequence
    uncertainty_map: np.ndarray  # [N, H, W]
    instants: dict = None
    voxelspacing: Tuple = None
    contour: np.ndarray = None
    mu: np.ndarray = None  # [N, K, 2]
    mode: np.ndarray = None  # [N, K, 2]
    cov: np.ndarray = None  # [N, K, 2, 2]
    alpha: np.ndarray = None  # [N, K, 2]
    pca_cov: np.ndarray = None
    post_mu: np.ndarray = None
    post_cov: np.ndarray = None
    contour_samples: np.ndarray = None  # [N, T_e, T_a, K, 2]
    pred_samples: np.ndarray = None  # [N, T_e, T_a, 1, H, W]
    entropy_map: np.ndarray = None  # [N, H, W]
    sample_weights: np.ndarray = None
    view_metrics: dict = None
    instant_metrics: dict = None
    view_uncertainty: dict = None
    instant_uncertainty: dict = None
    point_uncertainty: dict = None
    contour_validity: np.ndarray = None
    sample_validity: np.ndarray = None
    # extension: exact per-sample LV pixel counts (what lv_area(pred_samples[...]) returns)
    sample_area: np.ndarray = None  # [N, T_e, T_a]
    # extension: the sample masks bit-packed (256 rows x 8 uint32 words, bit x & 31 of word x >> 5), 8 KiB / sample instead of
    # the 512 KiB / sample of the int64 pred_samples; what the metric processors (Volume, GLS) read when present
    packed_samples: np.ndarray = None  # [N, T_e, T_a, 2048] int32
    # extension: p (1 - p) of the sample mean (the reference has no variance map)
    variance_map: np.ndarray = None  # [N, H, W]

    def __post_init__(self):
        assert self.img.ndim in [3, 4]
        n, w, h = self.img.shape[0], self.img.shape[-1], self.img.shape[-2]
        assert self.gt is None or self.gt.shape == (n, h, w)
        assert self.pred.shape == (n, h, w), f'Invalid shape for pred: got {self.pred.shape}'
        assert self.uncertainty_map.shape == (n, h, w), f'Invalid shape for u-map: got {self.uncertainty_map.shape}'
        if self.entropy_map is not None:
            assert self.entropy_map.shape == (n, h, w), f'Invalid shape for entropy: got {self.entropy_map.shape}'
        if self.instant_uncertainty is not None:
            for key, item in self.instant_uncertainty.items():
                assert item.ndim == 1 and len(item) == n, f'Invalid shape for instant_metrics {key}: {item.shape}'
        if self.mu is not None:
            assert self.mu.ndim == 3 and self.mu.shape[0] == n and self.mu.shape[-1] == 2
            k = self.mu.shape[1]
            assert self.cov.shape == (n, k, 2, 2)
            assert self.mode is None or self.mode.shape == (n, k, 2)
            assert self.alpha is None or self.alpha.shape == (n, k, 2)

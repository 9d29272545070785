"""Mirror of the sample post-processing of SegmentationUncertaintyTask.predict_step
(contour_uncertainty/task/segmentation/segmentation_uncertainty.py:98-152)."""
from dataclasses import dataclass
from typing import Optional

import torch

from ... import engine as E


@dataclass
class PostprocessedSamples:
    entropy: torch.Tensor                   # (N, 256, 256) fp32, 10-pixel frame zeroed            (:125-132)
    pred_samples: torch.Tensor              # (N, T_e, T_a, 2048) int32 bit-packed round(samples)  (:136, :152)
    masks: torch.Tensor                     # (N, T_e, T_a, 2048) int32 the kept blob (holes filled)
    area: torch.Tensor                      # (N, T_e, T_a) int32 pixels of the kept blob
    entropy_mean: torch.Tensor              # (N,) fp32                                            (:142)
    samples: Optional[torch.Tensor] = None  # (N, T_e, T_a, 1, 256, 256) fp32 raw * mask           (:115)

    def pred_samples_bool(self) -> torch.Tensor:
        """(N, T_e, T_a, 256, 256) bool, the layout of BatchResult.pred_samples."""
        return E.unpack_masks(self.pred_samples).bool()


def postprocess_samples(samples: torch.Tensor, pad: int = 10, return_samples: bool = False) -> PostprocessedSamples:
    """samples (N, T_e, T_a, 1, 256, 256) fp32 on the GPU: fill holes, keep the largest blob, multiply back, entropy."""
    if samples.ndim != 6 or samples.shape[3] != 1:
        raise ValueError("samples must be (N, T_e, T_a, 1, 256, 256): the post-processing is defined for binary segmentation")
    n, te, ta = (int(v) for v in samples.shape[:3])
    flat = samples.reshape(n, te * ta, E.HW, E.HW)
    masks, area = E.seg_largest_blob(flat, return_area=True)
    out = E.seg_entropy(flat, masks, pad=pad, return_pred_samples=True, return_samples=return_samples)
    return PostprocessedSamples(
        entropy=out["entropy"], pred_samples=out["pred_samples"].reshape(n, te, ta, E.MASK_WORDS),
        masks=masks.reshape(n, te, ta, E.MASK_WORDS), area=area.reshape(n, te, ta),
        entropy_mean=out["entropy"].mean((-2, -1)),
        samples=out["samples"].reshape(n, te, ta, 1, E.HW, E.HW) if return_samples else None)

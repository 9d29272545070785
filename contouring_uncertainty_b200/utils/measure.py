"""Mirror of vital/vital/utils/image/us/measure.py::EchoMeasure.gls and ._endo_base on the GPU."""
import numpy as np
import torch

from .. import engine as E


def pack_masks(masks) -> torch.Tensor:
    """(..., 256, 256) boolean-like numpy array -> (..., 2048) int32 bit-packed tensor on the current CUDA device."""
    m = np.asarray(masks).astype(bool).astype(np.uint8)
    if m.shape[-2:] != (E.HW, E.HW):
        raise ValueError("256 x 256 masks only")
    b = np.packbits(m, axis=-1, bitorder="little")
    t = torch.from_numpy(np.ascontiguousarray(b).view(np.int32).reshape(m.shape[:-2] + (E.MASK_WORDS,)))
    return t.to(torch.device("cuda", torch.cuda.current_device()))


class EchoMeasure:
    @staticmethod
    def _endo_base(segmentation, lv_labels, myo_labels):
        """measure.py:177-222: (256, 256) label map -> ((row, col), (row, col)) left / right markers at the base of the endocardium.
        Raises RuntimeError with fewer than two frontier pixels, like the reference."""
        seg = np.asarray(segmentation)
        c = E.endo_base_corners(pack_masks(np.isin(seg, lv_labels)), pack_masks(np.isin(seg, myo_labels))).cpu().numpy()
        if c[0, 0] < 0:
            raise RuntimeError("Identified fewer than 2 markers at the edges of the left ventricle/myocardium frontier. We need "
                               "to identify at least 2 such markers to determine the base of the left ventricle.")
        return c[0].astype(np.int64), c[1].astype(np.int64)

    @staticmethod
    def gls(segmentation, lv_labels=1, myo_labels=None, voxelspacing=(1, 1)):
        """measure.py:320-366: (N, 256, 256) segmentation of a sequence whose first frame is ED -> (N,) strain in percent.
        Raises RuntimeError where the reference raises (no contour / fewer than two base corners in a frame)."""
        seg = np.asarray(segmentation)
        if myo_labels is not None:
            length = E.mask_lv_length_myo(pack_masks(np.isin(seg, lv_labels)), pack_masks(np.isin(seg, myo_labels)), voxelspacing).cpu().numpy()
        else:
            length = E.mask_lv_length(pack_masks(np.isin(seg, lv_labels)), voxelspacing).cpu().numpy()
        if np.isnan(length).any():
            raise RuntimeError("unable to identify the base corners of the endocardium in frame(s) "
                               f"{np.nonzero(np.isnan(length))[0].tolist()}")
        return (length - length[0]) / length[0] * 100

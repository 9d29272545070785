"""Mirror of contour_uncertainty/data/camus/datamodule.py::CamusContourDataModule (:72-85), the part the predict path uses:
the three static functors the task fetches in ``on_predict_start`` (contour_uncertainty.py:59-63) and ``data_params``.
Dataset / HDF5 IO stays with the reference (out of scope)."""
from ..config import DataParameters, Label
from .utils import USContourToMask, USSkewUmap, USUMap


class CamusContourDataModule:
    contour_to_mask_fn = USContourToMask()
    umap_fn = USUMap()
    skew_umap_fn = USSkewUmap()

    def __init__(self, points_per_side: int = 11, labels=(Label.BG, Label.LV), in_shape=(1, 256, 256), **kwargs):
        labels = tuple(labels)
        remove_bg = int(Label.BG in labels)
        points_per_contour = points_per_side * 2 - 1
        nb_points = points_per_contour * (len(labels) - remove_bg)
        self.data_params = DataParameters(in_shape=tuple(in_shape), out_shape=(nb_points, 2), labels=labels)

"""Functor interfaces a datamodule hands to the task (the seam of contour_uncertainty/data/utils.py).

The reference binds three static callables on its datamodule (``contour_to_mask_fn``, ``umap_fn``, ``skew_umap_fn``,
camus/datamodule.py:72-74) and the task fetches them in ``on_predict_start``.  The GPU-backed implementations live in
``data/camus/utils.py``; these bases only fix the call signatures.
"""
import abc


class _StaticFunctor(abc.ABC):
    """A stateless callable used through ``staticmethod(datamodule.<name>)``: subclasses override ``__call__`` as a
    staticmethod with the signature documented on the concrete base below."""

    @staticmethod
    @abc.abstractmethod
    def __call__(*args, **kwargs):  # pragma: no cover - interface only
        raise NotImplementedError


class ContourToMask(_StaticFunctor):
    """``(landmarks (K,2), shape=(256,256), labels=None, apply_argmax=True, reconstruction_type='spline') -> int mask``;
    a leading channel axis is added when ``apply_argmax`` is False."""


class UMap(_StaticFunctor):
    """``(mu (K,2), cov (K,2,2), labels=None) -> (H,W) float64`` analytic uncertainty map."""


class SkewUMap(_StaticFunctor):
    """``(mu (K,2), cov (K,2,2), alpha (K,2), labels=None) -> (projected_mode (K,2), (H,W) float64 map)``."""

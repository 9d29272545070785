"""B200-native (sm_100a) uncertainty path of contouring-uncertainty.

heatmap logits -> DSNT moments -> Monte-Carlo contour samples -> rasterised masks -> pixel
uncertainty maps + clinical-metric distributions, behind the reference's Python call signatures.
Hand-written CUDA reached through a C ABI (include/cub200.h, lib/libcub200.so); no CPU fallback.
"""
from .engine import UncertaintyEngine, ShapeModelConstants, UncertaintyResult  # noqa: F401

__version__ = "0.1.0"

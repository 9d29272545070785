"""ctypes binding of libcub200.so (the C ABI declared in include/cub200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.  The
library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CUB200_LIB") or os.path.join(_HERE, "lib", "libcub200.so")

EINVAL, ECUDA = -1, -2
EXPECTED_ABI = 19          # the version of include/cub200.h the SIGNATURES table below was written for

_lib: Optional[C.CDLL] = None

_vp, _i32, _i64, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64

# name -> argtypes; every function returns int except the three queries
SIGNATURES = {
    "cub200_dsnt_moments": [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "cub200_psm_model_pack_host": [_vp, _vp, _vp, _vp, _vp],
    "cub200_psm_factor": [_vp, _vp, _vp, _i64, _vp, _vp],
    "cub200_naive_factor": [_vp, _i64, _vp, _vp],
    "cub200_sample_raster_aggregate": [_vp, _vp, _i32, _vp, _u64, _u64, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "cub200_step_counter_add": [_vp, _u64, _vp],
    "cub200_psm_sample_direct": [_vp, _vp, _vp, _vp, _u64, _u64, _i64, _vp, _vp],
    "cub200_metric_pack": [_vp, _i64, _vp, _i64, _vp, _vp],
    "cub200_rasterize_contours": [_vp, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp],
    "cub200_finalize_maps": [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp],
    "cub200_count_tables": [_i64, _vp, _vp],
    "cub200_soft_masks": [_vp, _i64, _vp, _vp, _vp],
    "cub200_metric_ae_split": [_vp, _i64, _i64, _i64, _i32, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp],
    "cub200_fac_from_areas": [_vp, _vp, _i64, _vp, _vp],
    "cub200_sample_moments": [_vp, _i64, _i64, _vp, _vp, _vp],
    "cub200_seq_condition": [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "cub200_skew_pdf_grid": [_vp, _vp, _vp, _i64, _i32, _vp, _vp],
    "cub200_psm_conditionals": [_vp, _vp, _i64, _vp, _vp, _vp],
    "cub200_skew_sample": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _i64, _i64, _vp, _vp, _vp],
    "cub200_simpson_diameters": [_vp, _i64, C.c_double, _vp, _vp, _vp, _vp, _vp],
    "cub200_simpson_diameters_stretched": [_vp, _i64, _vp, _vp, _i32, _i32, C.c_double, _vp, _vp, _vp, _vp, _vp],
    "cub200_simpson_volumes": [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp],
    "cub200_rasterize_contours_f64": [_vp, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp],
    "cub200_projected_uncertainty": [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp],
    "cub200_umap_analytic": [_vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp],
    "cub200_skew_umap_contours": [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp],
    "cub200_projected_uncertainty_generic": [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
    "cub200_weighted_mask_entropy": [_vp, _vp, _i64, _i32, C.c_double, _i32, _vp, _vp],
    "cub200_seq_skew_conditionals": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "cub200_seq_skew_sample": [_vp] * 12 + [_u64, _i64, _i64, _vp, _vp, _vp],
    "cub200_metric_sums": [_vp, _i64, _vp, _vp],
    "cub200_calibration_bins": [_vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp],
    "cub200_contour_perimeter": [_vp, _i64, _vp, _vp],
    "cub200_bivariate_pdf": [_vp, _i64, _vp, _vp, _vp, _i64, _i32, _vp, _vp],
    "cub200_skew_rvs_fast": [_vp, _vp, _vp, _vp, _u64, _i64, _vp, _vp],
    "cub200_seg_largest_blob": [_vp, _i64, _vp, _vp, _vp, _vp],
    "cub200_sample_entropy": [_vp, _i64, _i32, _i32, _i64, _vp, _vp],
    "cub200_us_myo_masks": [_vp, _i64, _i32, _vp, _vp, _vp, _vp],
    "cub200_mask_class_entropy": [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp],
    "cub200_rasterize_contours_generic": [_vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp],
    "cub200_mask_lv_length": [_vp, _i64, C.c_double, _vp, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp],
    "cub200_seg_entropy": [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "cub200_endo_base_corners": [_vp, _vp, _i64, _vp, _vp],
    "cub200_mask_lv_length_myo": [_vp, _vp, _i64, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp],
}
QUERIES = ("cub200_last_error", "cub200_abi_version", "cub200_built_for_sm", "cub200_simpson_workspace_bytes", "cub200_calibration_workspace_bytes",
           "cub200_seg_blob_workspace_bytes", "cub200_raster_workspace_bytes", "cub200_launch_count", "cub200_dsnt_workspace_bytes")


class NativeError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU / PyTorch fallback for this path.")
        l = C.CDLL(LIB_PATH)
        l.cub200_abi_version.restype = C.c_int
        have = l.cub200_abi_version()
        if have != EXPECTED_ABI:
            raise NativeError(f"{LIB_PATH} has ABI version {have}, this package binds version {EXPECTED_ABI}: rebuild it "
                              "(`python __graft_entry__.py --force`)")
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        l.cub200_last_error.restype = C.c_char_p
        l.cub200_last_error.argtypes = []
        l.cub200_simpson_workspace_bytes.restype = C.c_int64
        l.cub200_simpson_workspace_bytes.argtypes = [_i64]
        l.cub200_calibration_workspace_bytes.restype = C.c_int64
        l.cub200_calibration_workspace_bytes.argtypes = []
        l.cub200_seg_blob_workspace_bytes.restype = C.c_int64
        l.cub200_seg_blob_workspace_bytes.argtypes = [_i64]
        l.cub200_dsnt_workspace_bytes.restype = C.c_int64
        l.cub200_dsnt_workspace_bytes.argtypes = [_i64]
        l.cub200_launch_count.restype = C.c_int64
        l.cub200_launch_count.argtypes = []
        l.cub200_raster_workspace_bytes.restype = C.c_int64
        l.cub200_raster_workspace_bytes.argtypes = [_i64]
        l.cub200_abi_version.restype = C.c_int
        l.cub200_built_for_sm.restype = C.c_int
        _lib = l
    return _lib


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc == EINVAL:
        raise ValueError(f"{what}: invalid argument (shape / null pointer / unsupported size)")
    msg = lib().cub200_last_error().decode()
    raise NativeError(f"{what}: CUDA error: {msg}")


def require_cuda(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device: this path has no CPU implementation")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()

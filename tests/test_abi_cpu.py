"""CPU-only: the C-ABI library loads and exports every symbol include/cub200.h declares; the host
layer refuses CPU tensors (no fallback).  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cub200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cub200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    from contouring_uncertainty_b200 import _native

    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in cub200.h but not exported"
    assert set(names) == set(_native.SIGNATURES) | set(_native.QUERIES)
    assert _native.lib().cub200_abi_version() == _native.EXPECTED_ABI == 19
    assert _native.lib().cub200_built_for_sm() == 100


def test_sass_is_sm100a():
    import shutil
    import subprocess

    from contouring_uncertainty_b200 import _native

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_cpu_tensors_are_refused():
    import torch

    from contouring_uncertainty_b200 import engine

    with pytest.raises(ValueError, match="no CPU implementation"):
        engine.dsnt_moments(torch.zeros((1, 21, 256, 256)))
    with pytest.raises(ValueError, match="no CPU implementation"):
        engine.rasterize_contours(torch.zeros((1, 21, 2)))
    with pytest.raises(ValueError, match="no CPU implementation"):
        engine.finalize_maps(torch.zeros((1, 256, 256), dtype=torch.int32), 10)


def test_null_and_bad_sizes_return_einval():
    from contouring_uncertainty_b200 import _native

    l = _native.lib()
    assert l.cub200_dsnt_moments(None, 1, 256, 256, 1, 1, None, None, None, None, None) == _native.EINVAL
    assert l.cub200_dsnt_moments(None, 1, 128, 128, 1, 1, None, None, None, None, None) == _native.EINVAL
    assert l.cub200_sample_raster_aggregate(None, None, 34, None, 0, 0, 1, 1, 1, None, None, None, 1, None, None, None, None) == _native.EINVAL
    assert l.cub200_finalize_maps(None, 1, 10, None, None, None, None, None, 0, None, None, None) == _native.EINVAL
    # zero-sized work is a no-op success even without a device
    assert l.cub200_rasterize_contours(None, 0, 1, None, None, None, 0, None, None) == 0


def test_new_entry_points_validate_arguments():
    """Skew / sequence / u-map / Simpson entry points: null pointers and bad sizes -> EINVAL, empty work -> 0."""
    from contouring_uncertainty_b200 import _native

    l, E = _native.lib(), _native.EINVAL
    assert l.cub200_skew_pdf_grid(None, None, None, 1, 1, None, None) == E
    assert l.cub200_psm_conditionals(None, None, 1, None, None, None) == E
    assert l.cub200_psm_factor(None, None, None, 1, None, None) == E
    assert l.cub200_psm_model_pack_host(None, None, None, None, None) == E
    assert l.cub200_skew_sample(None, None, None, None, None, None, None, None, 0, 1, 1, None, None, None) == E
    assert l.cub200_seq_skew_conditionals(None, None, None, None, None, 1, None, None, None) == E
    assert l.cub200_seq_skew_sample(*([None] * 12), 0, 1, 1, None, None, None) == E
    assert l.cub200_projected_uncertainty(None, None, None, 1, 1, 0, None, None, None, None) == E
    assert l.cub200_umap_analytic(None, None, None, 1, None, 1, None, None, None) == E
    assert l.cub200_skew_umap_contours(None, None, None, None, 1, 21, None, None, None) == E
    assert l.cub200_projected_uncertainty_generic(None, None, None, 1, 44, 1, 0, None, None, None, None) == E
    assert l.cub200_weighted_mask_entropy(None, None, 1, 200, 1.0, 1, None, None) == E
    assert l.cub200_simpson_diameters(None, 1, 0.5, None, None, None, None, None) == E
    assert l.cub200_simpson_volumes(None, None, 1, 1, None, None, None, None) == E
    assert l.cub200_rasterize_contours_f64(None, 1, 0, None, None, None, 0, None, None) == E
    # empty work is a no-op
    assert l.cub200_skew_pdf_grid(None, None, None, 0, 1, None, None) == 0
    assert l.cub200_simpson_diameters(None, 0, 0.5, None, None, None, None, None) == 0
    assert l.cub200_simpson_volumes(None, None, 0, 10, None, None, None, None) == 0
    assert l.cub200_projected_uncertainty(None, None, None, 0, 1, 0, None, None, None, None) == 0
    assert l.cub200_rasterize_contours_f64(None, 0, 0, None, None, None, 0, None, None) == 0
    assert l.cub200_simpson_workspace_bytes(10) == 10 * (2048 * 4 + 4)
    assert l.cub200_seg_largest_blob(None, 1, None, None, None, None) == E
    assert l.cub200_sample_entropy(None, 1, 4, 3, 65536, None, None) == E
    assert l.cub200_sample_entropy(None, 0, 4, 3, 65536, None, None) == 0
    assert l.cub200_us_myo_masks(None, 1, 0, None, None, None, None) == E
    assert l.cub200_us_myo_masks(None, 0, 0, None, None, None, None) == 0
    assert l.cub200_rasterize_contours_generic(None, 0, 1, 26, 1, None, None, None) == E
    assert l.cub200_rasterize_contours_generic(None, 0, 0, 26, 1, None, None, None) == 0
    assert l.cub200_mask_lv_length(None, 1, 5e-3, None, 1.0, 1.0, None, None, None, None, None) == E
    assert l.cub200_mask_lv_length(None, 0, 5e-3, None, 1.0, 1.0, None, None, None, None, None) == 0
    assert l.cub200_endo_base_corners(None, None, 1, None, None) == E
    assert l.cub200_step_counter_add(None, 1, None) == E
    assert l.cub200_psm_sample_direct(None, None, None, None, 0, 0, 1, None, None) == E
    assert l.cub200_metric_pack(None, 1, None, 1, None, None) == E
    assert l.cub200_mask_lv_length_myo(None, None, 1, 1.0, 1.0, None, None, None, None, None) == E
    assert l.cub200_mask_lv_length_myo(None, None, 0, 1.0, 1.0, None, None, None, None, None) == 0
    assert l.cub200_seg_largest_blob(None, 0, None, None, None, None) == 0
    assert l.cub200_seg_entropy(None, None, 1, 1, 10, None, None, None, None, None) == E
    assert l.cub200_seg_entropy(None, None, 0, 1, 10, None, None, None, None, None) == 0
    assert l.cub200_seg_blob_workspace_bytes(3) == 16 + 3 * 3 * 32768 * 4 + 1 * 4 * 3 * 2048 * 4


def test_binding_tables_agree_with_the_header_and_the_integration_guide():
    """Every prototype of cub200.h has as many parameters as its ctypes signature, with pointers / integers / doubles in the
    same places; the stub printed in INTEGRATION.md is the same table."""
    import ctypes as C

    from contouring_uncertainty_b200 import _native

    src = open(os.path.join(ROOT, "include", "cub200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\bint\s+(cub200_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    seen = 0
    for name, params in protos:
        if name not in _native.SIGNATURES:
            continue
        seen += 1
        want = []
        for p in [q.strip() for q in params.split(",")]:
            if "*" in p:
                want.append(C.c_void_p)
            elif p.startswith("double"):
                want.append(C.c_double)
            elif p.startswith("uint64_t"):
                want.append(C.c_uint64)
            elif p.startswith("int64_t"):
                want.append(C.c_int64)
            elif p.startswith("int "):
                want.append(C.c_int)
            else:
                raise AssertionError(f"{name}: unrecognised parameter '{p}'")
        assert want == _native.SIGNATURES[name], name
    assert seen == len(_native.SIGNATURES)
    kinds = {"vp": C.c_void_p, "i32": C.c_int, "i64": C.c_int64, "u64": C.c_uint64, "C.c_double": C.c_double}
    guide = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name, args in re.findall(r"lib\.(cub200_\w+)\.argtypes = \[([^\]]*)\]", guide):
        assert [kinds[a.strip()] for a in args.split(",")] == _native.SIGNATURES[name], name

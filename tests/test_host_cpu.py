"""CPU tests of the host-side logic that needs no device."""
import os

import numpy as np

from oracle import port


def test_points_order_matches_oracle_and_reference_constants():
    from contouring_uncertainty_b200.sampler import Sampler

    assert Sampler.get_points_order(21, levels=3) == port.get_points_order(21, levels=3)
    ini, order = Sampler.get_points_order(21, levels=3)
    # sampling order == the permutation compiled into psm_factor.cu (c_order)
    flat = ini + sum(order, [])
    rest = [k for k in range(21) if k not in flat]
    assert flat + rest == [0, 10, 20, 5, 15, 2, 7, 13, 18, 1, 3, 6, 8, 12, 14, 17, 19, 4, 9, 11, 16]


def test_clinical_helpers_match_oracle():
    from contouring_uncertainty_b200.utils.clinical import aleatoric_epistemic_uncertainty, lv_FAC, lv_area

    rng = np.random.default_rng(0)
    ed = rng.random((5, 32, 32)) > 0.3
    es = rng.random((5, 32, 32)) > 0.6
    assert np.array_equal(lv_area(ed), port.lv_area(ed))
    assert np.allclose(lv_FAC(ed, es), port.lv_fac(ed, es))
    mc = rng.random((4, 25))
    mc[1, 3] = np.nan
    assert np.allclose(aleatoric_epistemic_uncertainty(mc), port.aleatoric_epistemic_uncertainty(mc))


def test_shape_model_constants_reproduce_pca_covariance():
    """C0 + d d^T == pca() covariance about an arbitrary centre (posteriorshapemodel.py:25-27)."""
    import torch

    from contouring_uncertainty_b200.engine import DEFAULT_PSM, ShapeModelConstants

    m = ShapeModelConstants(DEFAULT_PSM, device="cpu")
    om = port.ShapeModel.load(DEFAULT_PSM)
    mu_t = np.random.default_rng(1).normal(0, 5, 42)
    want = port.pca_cov(om.X_train, mu_t)
    d = m.xbar.numpy() - mu_t
    got = m.c0.numpy() + np.outer(d, d)
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()


def test_batch_result_shape_asserts():
    import pytest

    from contouring_uncertainty_b200.data.config import BatchResult

    z = np.zeros((2, 256, 256))
    BatchResult(id="x", img=z, gt=z, pred=z, labels=["lv"], uncertainty_map=z)
    with pytest.raises(AssertionError):
        BatchResult(id="x", img=z, gt=z, pred=z[0], labels=["lv"], uncertainty_map=z)


def test_gls_gauss_table_is_scipys_kernel():
    """The weight table handed to lv_length_kernel == scipy.ndimage's own Gaussian kernel (right half, centre first)."""
    from scipy.ndimage._filters import _gaussian_kernel1d

    from contouring_uncertainty_b200.engine import GLS_TABLE_STRIDE, SIMPSON_MAX_POINTS, gls_gauss_table

    tab = gls_gauss_table(5e-3, "cpu").numpy()
    assert tab.shape == (SIMPSON_MAX_POINTS + 1, GLS_TABLE_STRIDE)
    for n in (3, 24, 25, 100, 333, 547, 1000, 2048):
        sigma = n * 5e-3
        lw = int(4.0 * sigma + 0.5)
        assert lw < GLS_TABLE_STRIDE
        if lw == 0:
            assert tab[n, 0] == 1.0 and not tab[n, 1:].any()
            continue
        w = _gaussian_kernel1d(sigma, 0, lw)[::-1]
        assert np.array_equal(tab[n, :lw + 1], w[lw:]) and not tab[n, lw + 1:].any()


def test_seg_and_gls_wrappers_refuse_cpu_tensors():
    import pytest
    import torch

    from contouring_uncertainty_b200 import engine as E

    with pytest.raises(ValueError, match="no CPU implementation"):
        E.seg_largest_blob(torch.zeros((1, 256, 256)))
    with pytest.raises(ValueError, match="no CPU implementation"):
        E.mask_lv_length(torch.zeros((1, 2048), dtype=torch.int32))
    with pytest.raises(ValueError, match="no CPU implementation"):
        E.seg_entropy(torch.zeros((1, 1, 256, 256)), torch.zeros((1, 1, 2048), dtype=torch.int32))


def test_dataset_level_calibration_and_correlation_match_the_reference():
    """results/utils/calibration.py mirror vs the numbers the reference's own compute_calibration /
    compute_adaptive_calibration / pearsonr / mean_absolute_error gave (tests/golden/make_golden.py dataset_calibration)."""
    import json

    import synth
    from contouring_uncertainty_b200.results.utils import calibration as C

    std, err, reject = synth.dataset_calibration_inputs()
    got = {"uce": C.compute_calibration(err[20:], std[20:])[0], "a_uce": C.compute_adaptive_calibration(err[20:], std[20:], nb_bins=7)[0],
           **{f"m_{k}": v for k, v in C.metric_calibration(std, err, reject).items()},
           **C.metric_correlation(np.where(reject, np.inf, err), std)}
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_calibration.json")))
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-12 * max(1.0, abs(v)), (k, got[k], v)

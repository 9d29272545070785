"""CPU tests of the host-side logic that needs no device."""
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

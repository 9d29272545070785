"""GPU parity of the segmentation-sample post-processing against the oracle and the golden made by
executing SegmentationUncertaintyTask.predict_step."""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import port

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dev():
    return torch.device("cuda:0")


def test_golden_predict_step():
    from contouring_uncertainty_b200.task.segmentation.segmentation_uncertainty import postprocess_samples

    g = np.load(os.path.join(G, "segpost.npz"))
    n, Te, Ta = (int(v) for v in g["shape"])
    samples = synth.segmentation_samples(n, Te * Ta, seed=int(g["seed"])).reshape(n, Te, Ta, 1, 256, 256)
    res = postprocess_samples(torch.from_numpy(samples).to(_dev()), return_samples=True)
    want = np.unpackbits(g["pred_samples"], axis=-1, bitorder="little").astype(bool)
    got = res.pred_samples_bool().cpu().numpy().reshape(n, Te * Ta, 256, 256)
    assert np.array_equal(got, want)
    # fp32 tolerance: the reference averages the samples in fp32 (torch), the kernel in fp64 rounded once
    assert np.abs(res.entropy.cpu().numpy() - g["entropy"]).max() < 2e-6
    np.testing.assert_allclose(res.entropy_mean.cpu().numpy(), g["entropy_mean"], atol=1e-6)
    masks = port.seg_blob_masks(samples.reshape(n, Te * Ta, 256, 256))
    from contouring_uncertainty_b200.engine import unpack_masks

    assert np.array_equal(unpack_masks(res.masks).cpu().numpy().astype(bool).reshape(masks.shape), masks)
    assert np.array_equal(res.area.cpu().numpy().reshape(n, -1), masks.sum((-2, -1)))
    assert np.array_equal(res.samples.cpu().numpy().reshape(masks.shape), samples.reshape(masks.shape) * masks)


def test_random_samples_against_oracle():
    from contouring_uncertainty_b200 import engine as E

    n, T = 2, 24
    samples = synth.segmentation_samples(n, T, seed=77, adversarial=False)
    # a few pure-noise maps at several densities: percolation threshold, many equal-sized specks, overflow path
    rng = np.random.default_rng(5)
    for k, dens in enumerate((0.02, 0.3, 0.5, 0.593, 0.7, 0.97)):
        samples[1, k] = (rng.random((256, 256)) < dens).astype(np.float32) * 0.75
    t = torch.from_numpy(samples).to(_dev())
    masks, area = E.seg_largest_blob(t, return_area=True)
    want = port.seg_blob_masks(samples)
    got = E.unpack_masks(masks).cpu().numpy().astype(bool)
    for i in range(n):
        for k in range(T):
            assert np.array_equal(got[i, k], want[i, k]), (i, k)
    assert np.array_equal(area.cpu().numpy(), want.sum((-2, -1)))
    out = E.seg_entropy(t, masks, pad=10, return_mean=True, return_pred_samples=True)
    post, ent, pred_samples = port.seg_postprocess(samples.reshape(n, 1, T, 1, 256, 256))
    assert np.array_equal(E.unpack_masks(out["pred_samples"]).cpu().numpy().astype(bool), pred_samples.reshape(n, T, 256, 256))
    assert np.abs(out["entropy"].cpu().numpy() - ent).max() < 2e-6
    mean64 = post.reshape(n, T, 256, 256).astype(np.float64).mean(1)
    assert np.abs(out["mean"].cpu().numpy() - mean64).max() < 1e-7


@pytest.mark.parametrize("T", [1, 5, 33, 70])
def test_sample_counts_off_the_chunk_grid(T):
    """The entropy kernel walks the samples in chunks of 32 and groups of 4: counts that fit neither."""
    from contouring_uncertainty_b200 import engine as E

    samples = synth.segmentation_samples(2, T, seed=100 + T, adversarial=False)
    t = torch.from_numpy(samples).to(_dev())
    masks = E.seg_largest_blob(t)
    out = E.seg_entropy(t, masks, pad=10, return_pred_samples=True, return_samples=True, return_mean=True)
    post, ent, pred_samples = port.seg_postprocess(samples.reshape(2, 1, T, 1, 256, 256))
    assert np.array_equal(E.unpack_masks(out["pred_samples"]).cpu().numpy().astype(bool), pred_samples.reshape(2, T, 256, 256))
    assert np.array_equal(out["samples"].cpu().numpy(), post.reshape(2, T, 256, 256))
    assert np.abs(out["entropy"].cpu().numpy() - ent).max() < 2e-6


def test_big_blob_mirror_and_edge_cases():
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.task.segmentation.utils import big_blob

    seg = np.zeros((3, 256, 256), dtype=np.int64)
    seg[0, 10:20, 10:20] = 1
    seg[0, 12:14, 12:14] = 0                       # a hole: big_blob must not fill it
    seg[0, 100:105, 100:105] = 1
    seg[1, 5, 5] = 1
    seg[1, 200, 200] = 1                           # tie of two single pixels -> the first
    got = big_blob(seg)
    for i in range(3):
        assert np.array_equal(got[i], port.big_blob(seg[i])), i
    # empty batch
    empty = torch.zeros((0, 256, 256), device=_dev())
    assert E.seg_largest_blob(empty).shape == (0, 2048)
    with pytest.raises(ValueError):
        E.seg_largest_blob(torch.zeros((2, 128, 128), device=_dev()))


def test_full_size_properties():
    """64 images x 32 samples: idempotence of the blob filter, area = popcount, pred_samples inside the mask."""
    from contouring_uncertainty_b200 import engine as E

    n, T = 64, 32
    base = torch.from_numpy(synth.segmentation_samples(4, T, seed=9, adversarial=True)).to(_dev())
    g = torch.Generator(device=_dev()).manual_seed(1)
    probs = base.repeat(n // 4, 1, 1, 1)
    probs = (probs + 0.08 * torch.randn(probs.shape, device=_dev(), generator=g)).clamp_(0, 1).contiguous()
    masks, area = E.seg_largest_blob(probs, return_area=True)
    m = E.unpack_masks(masks)
    assert torch.equal(m.sum((-2, -1), dtype=torch.int32), area)
    again, area2 = E.seg_largest_blob(m.float(), return_area=True)
    assert torch.equal(again, masks) and torch.equal(area2, area)
    out = E.seg_entropy(probs, masks, return_pred_samples=True, return_samples=True, return_mean=True)
    ps = E.unpack_masks(out["pred_samples"]).bool()
    assert not (ps & ~m.bool()).any()
    assert torch.equal(ps, (probs * m).round() != 0)
    assert torch.equal(out["samples"], probs * m)
    ent = out["entropy"]
    assert (ent[:, :10] == 0).all() and (ent[:, -10:] == 0).all() and (ent[:, :, :10] == 0).all() and (ent[:, :, -10:] == 0).all()
    assert ent.min() >= 0 and ent.max() <= 1.0 + 1e-6
    y = (probs * m).double().mean(1)
    assert (out["mean"].double() - y).abs().max() < 1e-7

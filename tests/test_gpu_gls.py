"""GPU parity of the mask-based longitudinal strain (EchoMeasure.gls) against the golden made by executing
the reference (find_contours restated: parity unpinned for that third-party function) and against the oracle."""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import port

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _pack(masks):
    b = np.packbits(masks.astype(np.uint8), axis=-1, bitorder="little")
    return np.ascontiguousarray(b).view(np.uint32).reshape(masks.shape[:-2] + (2048,)).view(np.int32)


def test_golden_lengths_corners_gls():
    from contouring_uncertainty_b200 import engine as E

    g = np.load(os.path.join(G, "glsmask.npz"))
    dev = torch.device("cuda:0")
    packed = torch.from_numpy(np.ascontiguousarray(g["masks"]).view(np.int32).reshape(-1, 2048)).to(dev)
    length, corners, npts = E.mask_lv_length(packed, return_corners=True)
    length, corners, npts = length.cpu().numpy(), corners.cpu().numpy(), npts.cpu().numpy()
    want = g["length"]
    assert np.array_equal(np.isnan(length), np.isnan(want))                  # the reference raised for exactly these masks
    ok = ~np.isnan(want)
    # mirror-symmetric masks (the disc) have base peaks that tie to the last bit; which two the reference keeps is the tie
    # order of numpy's unstable argsort -- not part of the parity contract
    m_all = np.unpackbits(g["masks"], axis=-1, bitorder="little").astype(bool)
    ties = np.array([bool(ok[i]) and port.base_corner_tie(m_all[i]) for i in range(len(ok))])
    assert ties.sum() == 1 and ties[40]
    ok &= ~ties
    bad = [(int(i), corners[i].tolist(), g["corners"][i].tolist()) for i in np.nonzero(ok)[0] if not np.array_equal(corners[i], g["corners"][i])]
    assert not bad, bad                                                      # same contour points picked as base corners
    assert np.array_equal(npts[ok], g["npts"][ok])
    # the segment norms go through BLAS ddot in the reference (possibly fused): 1e-13 relative on the sum
    np.testing.assert_allclose(length[ok], want[ok], rtol=1e-13)
    T = g["gls"].shape[1]
    seq = packed[: 4 * T].reshape(4, T, 2048)
    for v in range(2):
        pair = torch.stack([seq[2 * v], seq[2 * v + 1]], dim=1)              # (T, 2 frames, 2048)
        gls = E.mask_gls(pair)[:, 1].cpu().numpy()
        assert np.array_equal(np.isnan(gls), np.isnan(g["gls"][v]))
        np.testing.assert_allclose(gls[~np.isnan(gls)], g["gls"][v][~np.isnan(gls)], rtol=1e-11)


def test_random_masks_against_oracle():
    """Sampled LV contours (PSM-like jitter), rasterised on the GPU, measured by both."""
    from contouring_uncertainty_b200 import engine as E

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    n = 96
    c = synth.LV_EXAMPLE_SHAPE[None] + rng.normal(0, 2.5, (n, 21, 2)) + rng.normal(0, 10, (n, 1, 2))
    c[n // 2:] = (c[n // 2:] - c[n // 2:].mean(1, keepdims=True)) * rng.uniform(0.5, 1.1, (n - n // 2, 1, 1)) + 128
    masks, _ = E.rasterize_contours(torch.from_numpy(c.astype(np.float32)).to(dev))
    length, corners, npts = E.mask_lv_length(masks, voxelspacing=(0.7, 0.4), return_corners=True)
    m = E.unpack_masks(masks).cpu().numpy().astype(bool)
    bad = []
    for i in range(n):
        try:
            left, right, _ = port.polar_base_corners(m[i])
            want = port.lv_longitudinal_length(m[i], (0.7, 0.4))
        except (RuntimeError, IndexError):
            if not np.isnan(length[i].item()):
                bad.append((i, "expected NaN"))
            continue
        if not (np.array_equal(corners[i].cpu().numpy(), np.array([left, right])) and abs(length[i].item() - want) <= 1e-12 * want):
            bad.append((i, length[i].item(), want))
    assert not bad, bad


def test_echo_measure_mirror_and_errors():
    from contouring_uncertainty_b200.utils.measure import EchoMeasure

    g = np.load(os.path.join(G, "glsmask.npz"))
    masks = np.unpackbits(g["masks"], axis=-1, bitorder="little").astype(np.int64)
    T = g["gls"].shape[1]
    got = EchoMeasure.gls(np.stack([masks[0], masks[T]]), 1, None)
    assert got[0] == 0 and abs(got[1] - g["gls"][0, 0]) < 1e-10
    with pytest.raises(RuntimeError):
        EchoMeasure.gls(np.stack([masks[0], np.zeros((256, 256), dtype=np.int64)]), 1, None)


def test_endo_base_golden_corners_lengths_gls():
    """EchoMeasure.gls with myo_labels (_endo_base markers, measure.py:177-222) against the executed reference: LV + MYO label
    maps incl. a one-column frontier (the vertical branch), a single frontier pixel and a map without myocardium."""
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.utils.measure import EchoMeasure

    g = np.load(os.path.join(G, "glsmyo.npz"))
    dev = torch.device("cuda:0")
    maps = g["maps"]
    lv = torch.from_numpy(_pack(maps == 1)).to(dev)
    myo = torch.from_numpy(_pack(maps == 2)).to(dev)
    corners = E.endo_base_corners(lv, myo).cpu().numpy()
    assert np.array_equal(corners, g["corners"])                             # -1 where the reference raised
    length, c2, npts = E.mask_lv_length_myo(lv, myo, voxelspacing=(0.7, 0.4), return_corners=True)
    length, npts = length.cpu().numpy(), npts.cpu().numpy()
    assert np.array_equal(c2.cpu().numpy(), g["corners"])
    want = g["length"]
    assert np.array_equal(np.isnan(length), np.isnan(want)) and np.isnan(want).sum() == 2
    ok = ~np.isnan(want)
    assert np.array_equal(npts[ok], g["npts"][ok])
    np.testing.assert_allclose(length[ok], want[ok], rtol=1e-13)             # the reference's norms go through BLAS ddot
    for (a, b), gl in zip(g["pairs"], g["gls"]):
        got = E.mask_gls(torch.stack([lv[a], lv[b]]), (0.7, 0.4), myo_masks=torch.stack([myo[a], myo[b]]))[1].item()
        assert abs(got - gl) <= 1e-11 * abs(gl)
        np.testing.assert_allclose(EchoMeasure.gls(np.stack([maps[a], maps[b]]), 1, 2, (0.7, 0.4))[1], gl, rtol=1e-11)
    l, r = EchoMeasure._endo_base(maps[0], 1, 2)
    assert l.tolist() == g["corners"][0, 0].tolist() and r.tolist() == g["corners"][0, 1].tolist()
    with pytest.raises(RuntimeError):
        EchoMeasure._endo_base(maps[40], 1, 2)
    with pytest.raises(RuntimeError):
        EchoMeasure.gls(np.stack([maps[0], maps[40]]), 1, 2)


def test_endo_base_random_landmark_sets_against_oracle():
    """LV + MYO masks rasterised on the GPU from jittered 42-landmark sets; markers, point counts and lengths vs the oracle."""
    from contouring_uncertainty_b200 import engine as E

    dev = torch.device("cuda:0")
    lm = np.concatenate([synth.endo_epi_landmarks(40, seed=5), synth.endo_epi_landmarks(8, seed=6, wild=True)]).astype(np.float32)
    lv, myo, _ = E.us_myo_masks(torch.from_numpy(lm).to(dev))
    length, corners, npts = E.mask_lv_length_myo(lv, myo, return_corners=True)
    seg = E.unpack_masks(lv).cpu().numpy().astype(np.uint8) + 2 * E.unpack_masks(myo).cpu().numpy().astype(np.uint8)
    bad = []
    for i in range(len(lm)):
        try:
            want_c = np.array(port.endo_base_corners(seg[i]))
            want_n = len(port.endo_contour_without_base_myo(seg[i]))
            want = port.lv_longitudinal_length_myo(seg[i])
        except (RuntimeError, IndexError):
            if not np.isnan(length[i].item()):
                bad.append((i, "expected NaN"))
            continue
        if not (np.array_equal(corners[i].cpu().numpy(), want_c) and npts[i].item() == want_n and abs(length[i].item() - want) <= 1e-12 * want):
            bad.append((i, corners[i].tolist(), want_c.tolist(), length[i].item(), want))
    assert not bad, bad

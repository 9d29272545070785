"""Pin the oracle (oracle/port.py) against outputs of the reference itself (tests/golden/*.npz,
made by tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

import synth
from oracle import port

G = os.path.join(os.path.dirname(__file__), "golden")


def _unpack(a):
    return np.unpackbits(a, axis=-1, bitorder="little")


@pytest.mark.parametrize("name", ["bumps", "noise", "bimodal"])
def test_dsnt_matches_reference(name):
    g = np.load(os.path.join(G, "dsnt.npz"))
    logits = synth.heatmap_logits(int(g[f"{name}_n"]), seed=int(g[f"{name}_seed"]), kind=name)
    coords, var, covar = port.dsnt(port.flat_softmax(logits.astype(np.float64)))
    # the reference runs in fp32 torch: its own error vs fp64 truth is ~1.5e-6 rel
    np.testing.assert_allclose(coords, g[f"{name}_coords"], rtol=0, atol=6e-6)  # fp32 sum of 65536 terms
    np.testing.assert_allclose(var, g[f"{name}_var"], rtol=2e-5, atol=1e-9)
    scale = np.sqrt(var[..., 0] * var[..., 1])
    assert np.max(np.abs(covar - g[f"{name}_covar"]) / scale) < 2e-5
    mu, sig = port.dsnt_predict_on_batch(logits)
    np.testing.assert_allclose(mu, g[f"{name}_pixel_mu"], rtol=0, atol=8e-4)  # 128 x the coords bound
    np.testing.assert_allclose(sig, g[f"{name}_pixel_sigma"], rtol=2e-5, atol=2e-4)


def test_raster_bit_exact():
    g = np.load(os.path.join(G, "raster.npz"))
    contours, masks = g["contours"], _unpack(g["masks"])
    bad = []
    for i, c in enumerate(contours):
        m = port.us_contour_to_mask(c)
        if not np.array_equal(m, masks[i]):
            bad.append((i, str(g["tags"][i])))
    assert not bad, bad
    unr = _unpack(g["masks_unrounded"])
    for i in range(unr.shape[0]):
        assert np.array_equal(port.reconstruction(contours[i].astype(np.float64), 256, 256), unr[i])
    lin = _unpack(g["masks_linear"])                 # reconstruction_type='linear'
    bad = [(i, str(g["tags"][i])) for i, c in enumerate(contours)
           if not np.array_equal(port.linear_reconstruction(c.round().astype(int), 256, 256), lin[i])]
    assert not bad, bad


def test_own_spline_matches_fitpack():
    """The collocation spline the CUDA kernel implements == splprep/splev to ~1e-12, same pixels."""
    g = np.load(os.path.join(G, "raster.npz"))
    worst = 0.0
    for i, c in enumerate(g["contours"][:120]):
        pts = c.round().astype(int).astype(np.float64)
        try:
            own = port.own_spline(pts, 1000)
        except ValueError:
            assert str(g["tags"][i]) in ("dup", "tiny"), g["tags"][i]
            continue
        ref = port.contour_spline(pts, n=1000)
        worst = max(worst, np.abs(own - ref).max())
        assert np.array_equal(own.round().astype(int), ref.round().astype(int)), i
    assert worst < 1e-10, worst


def test_psm_sampler_matches_reference():
    g = np.load(os.path.join(G, "psm.npz"))
    model = port.ShapeModel.load(os.path.join(os.path.dirname(G), "..", "contouring_uncertainty_b200", "data",
                                              "camus-cont_psm_11_no_std.npy"))
    assert model.initial_points == list(g["initial_points"])
    assert sum(model.points_order, []) == list(g["order"])
    for i in range(g["mu"].shape[0]):
        s = port.psm_sample(model, g["mu"][i], g["cov"][i], g["z"][i])
        # tight anchor: the reference's own code executed in fp64 (make_golden.py)
        assert np.abs(s - g["samples_f64"][i]).max() < 1e-8
        # loose anchor: the verbatim fp32 reference, whose ill-conditioned fp32 42x42 inverses
        # (sigma2 = 1e-3 fill of points 4, 9, 11, 16) carry up to ~3.4e-2 px of its own noise
        assert np.abs(s - g["samples"][i]).max() < 6e-2
        sampled = [k for k in range(21) if k not in (4, 9, 11, 16)]
        assert np.abs(s[:, sampled] - g["samples"][i][:, sampled]).max() < 6e-3


def test_naive_sampler_matches_reference():
    g = np.load(os.path.join(G, "naive.npz"))
    for i in range(g["mu"].shape[0]):
        s = port.naive_sample(g["mu"][i], g["cov"][i], g["z"][i])
        np.testing.assert_allclose(s, g["samples"][i], rtol=0, atol=1e-9)


def test_maps_match_reference():
    g = np.load(os.path.join(G, "maps.npz"))
    p = np.load(os.path.join(G, "psm.npz"))
    for i in range(p["samples"].shape[0]):
        r = port.uncertainty_from_samples(p["samples"][i])
        assert np.array_equal(r["count"], g["count"][i])
        assert np.array_equal(r["pred"], g["pred"][i])
        assert np.array_equal(r["area"], g["area"][i])
        np.testing.assert_allclose(r["entropy"], g["entropy"][i], rtol=0, atol=1e-12)
        np.testing.assert_allclose(r["mean"], g["mean"][i], rtol=0, atol=1e-15)
        m, e, pr = port.maps_from_counts(r["count"], r["T"])
        np.testing.assert_allclose(e, g["entropy"][i], rtol=0, atol=1e-12)
        assert np.array_equal(pr, g["pred"][i])


def test_points_order():
    ini, order = port.get_points_order(21, levels=3)
    assert ini == [0, 10, 20]
    assert order == [[5, 15], [2, 7, 13, 18], [1, 3, 6, 8, 12, 14, 17, 19]]


def test_sequence_sampler_matches_reference():
    g = np.load(os.path.join(G, "seq.npz"))
    data = os.path.join(os.path.dirname(G), "..", "contouring_uncertainty_b200", "data")
    model = port.ShapeModel.load(os.path.join(data, "camus-cont_psm_11_no_std.npy"))
    seq = port.SequenceShapeModel.load(os.path.join(data, "camus-cont_sequence_psm_11_no_std.npy"))
    for i in range(g["mu"].shape[0]):
        s = port.seq_psm_sample(model, seq, g["mu"][i], g["cov"][i], g["bits"][i], g["z"][i])
        assert np.abs(s - g["samples_f64"][i]).max() < 1e-7          # the reference's own code run in fp64
        assert np.abs(s - g["samples"][i]).max() < 0.2               # verbatim fp32 run (its own noise)


def _skew_q(g, i, n_s):
    return lambda n, d: synth.exponential_field(int(g["q_seed"]), (i * n_s + n) * 14 + d)


@pytest.mark.parametrize("img", [0, 2])
def test_skew_sampler_matches_reference(img):
    """psm_skew.py with the rvs_fast normals and the multinomial's exponential race replayed.  Grid indices
    are exact against both the verbatim fp32 run (incl. its 3 underflow fallbacks on image 2) and the fp64
    run of the same code; the fp64 run pins the float parts to 1e-9 px."""
    g = np.load(os.path.join(G, "skew.npz"))
    model = port.ShapeModel.load(os.path.join(os.path.dirname(G), "..", "contouring_uncertainty_b200", "data",
                                              "camus-cont_psm_11_no_std.npy"))
    n_s = g["z"].shape[1]
    s32, i32 = port.skew_psm_sample(model, g["mu"][img], g["cov"][img], g["alpha"][img], g["z"][img],
                                    _skew_q(g, img, n_s), store=np.float32, return_indices=True)
    assert np.array_equal(i32, g["indices"][img])
    assert (g["indices"][2] < 0).sum() == 3
    assert np.abs(s32 - g["samples"][img]).max() < 5e-2          # the reference's own fp32 PSM noise on the 4 filled points
    if img == 0:
        s64, i64 = port.skew_psm_sample(model, g["mu"][img], g["cov"][img], g["alpha"][img], g["z"][img],
                                        _skew_q(g, img, n_s), store=np.float64, return_indices=True)
        assert np.array_equal(i64, g["indices_f64"][img])
        assert np.abs(s64 - g["samples_f64"][img]).max() < 1e-9


def test_skew_pdf_table_matches_reference():
    g = np.load(os.path.join(G, "skew.npz"))
    al = g["alpha"] * np.array([1.0, -1.0])
    for i in range(3):
        tab = port.skew_pdf_grid(g["mu"][i], g["cov"][i], al[i])[:, 3::16, 5::16]
        ref = g["pdf_sub"][i]
        assert np.abs(tab - ref).max() < 2e-8
        big = ref > 1e-6
        assert (np.abs(tab - ref) / ref)[big].max() < 5e-5       # fp32 exp of a fp32 exponent of magnitude ~15


def test_simpson_matches_reference():
    """heart_us.py executed on rasterised contour samples (with the restated find_contours: parity unpinned for
    that third-party function).  The oracle restates _compute_diameters / the volume formula / the MC loop."""
    g = np.load(os.path.join(G, "simpson.npz"))
    contours = g["contours"]
    vs = tuple(g["voxelspacing"])
    T = contours.shape[1]
    masks = port.convert_to_mask(contours).astype(bool)
    assert np.array_equal(masks.sum((2, 3)), g["areas"])
    for v, t in ((0, 0), (1, T - 4), (0, T - 3), (2, T - 2), (3, T - 1)):      # ordinary + the adversarial ones
        d, s = port.compute_diameters(masks[v, t], vs)
        assert np.array_equal(d, g["diameters"][v, t]) and s == g["step"][v, t]
    sel = [0, T - 4, T - 3, T - 2, T - 1]
    edv, esv, ef = port.volumes_mc(masks[0][sel], masks[1][sel], masks[2][sel], masks[3][sel], vs, vs)
    assert np.array_equal(edv, g["edv"][sel]) and np.array_equal(esv, g["esv"][sel])
    assert np.array_equal(ef, g["ef"][sel], equal_nan=True)
    # an empty mask has no contour: the reference raises inside the bare except -> zeros
    e = np.zeros((1, 256, 256), dtype=bool)
    z = port.volumes_mc(e, masks[1][:1], masks[2][:1], masks[3][:1], vs, vs)
    assert z[0][0] == 0 and z[1][0] == 0 and z[2][0] == 0


def test_find_contours_closed_and_open():
    """Structure of the restated marching squares: a closed blob gives one closed contour starting at the to-point
    of its raster-last segment; a blob cut by the frame gives an open contour; creation order is raster order."""
    m = np.zeros((16, 16))
    m[3:8, 4:9] = 1
    c = port.find_contours(m, 0.5)
    assert len(c) == 1 and np.array_equal(c[0][0], c[0][-1]) and len(c[0]) == 21
    assert tuple(c[0][0]) == (7.5, 8.0)      # raster-last cell (7, 8) is case 1: top -> left
    m2 = np.zeros((16, 16))
    m2[5:9, 12:16] = 1                       # touches the right border
    m2[1:3, 1:3] = 1
    c2 = port.find_contours(m2, 0.5)
    assert len(c2) == 2 and np.array_equal(c2[0][0], c2[0][-1]) and c2[0][:, 0].max() < 3   # raster-first blob first
    assert not np.array_equal(c2[1][0], c2[1][-1]) and c2[1][0][1] == 15 and c2[1][-1][1] == 15


def _sparse_map(g, name, i):
    m = np.zeros(256 * 256)
    m[g[f"{name}_idx_{i}"]] = g[f"{name}_val_{i}"]
    return m.reshape(256, 256)


def test_umaps_match_reference():
    """projected_uncertainty / USUMap / USSkewUmap.  The reference rotates covariances in fp32 torch, so its projected
    standard deviations are reproduced to 1 fp32 ulp; maps are compared pixel by pixel."""
    g = np.load(os.path.join(G, "umap.npz"))
    for i in range(3):
        mu, cov, al = g["mu"][i], g["cov"][i], g["alpha"][i]
        u, v = port.projected_uncertainty(mu, cov, all=True)
        assert np.abs(u.view(np.int32) - g["unc_all"][i].view(np.int32)).max() <= 1 and np.abs(v - g["proj_all"][i]).max() < 1e-12
        u0, _ = port.projected_uncertainty(mu, cov)
        np.testing.assert_allclose(u0, g["unc"][i], rtol=3e-7)
        np.testing.assert_allclose(port.projected_uncertainty_value(mu, cov), g["value"][i], rtol=3e-7)
        u2, v2, ap = port.projected_uncertainty(mu, cov, al.copy(), all=True, linear_close=True)
        assert np.abs(u2.view(np.int32) - g["skew_unc"][i].view(np.int32)).max() <= 1
        assert np.abs(v2 - g["skew_proj"][i]).max() < 1e-12 and np.abs(ap - g["alpha_proj"][i]).max() < 2e-6
        if i == 0:
            um, ref = port.us_umap(mu, cov), _sparse_map(g, "umap", i)
            assert (um != ref).sum() == 0                       # pixel-identical to the executed reference
            mode, sk = port.us_skew_umap(mu, cov, al)
            ref = _sparse_map(g, "skew_umap", i)
            assert np.abs(mode - g["mode"][i]).max() < 1e-4 and (np.abs(sk - ref) > 1e-12).sum() == 0


def test_sequence_skew_sampler_matches_reference():
    """psm_skew_sequence.py with first-instant bits, rvs_fast normals and all 31 multinomial races replayed."""
    g = np.load(os.path.join(G, "seqskew.npz"))
    D = os.path.join(os.path.dirname(G), "..", "contouring_uncertainty_b200", "data")
    model = port.ShapeModel.load(os.path.join(D, "camus-cont_psm_11_no_std.npy"))
    seq = port.SequenceShapeModel.load(os.path.join(D, "camus-cont_sequence_psm_11_no_std.npy"))
    n_s = g["z"].shape[1]
    i = 0
    qf = lambda n, d: synth.exponential_field(int(g["q_seed"]), (i * n_s + n) * 31 + d)
    s, idx = port.seq_skew_psm_sample(model, seq, g["mu"][i], g["cov"][i], g["alpha"][i], g["bits"][i], g["z"][i], qf,
                                      store=np.float64, return_indices=True)
    assert np.array_equal(idx, g["indices_f64"][i]) and np.abs(s - g["samples_f64"][i]).max() < 1e-9
    assert np.array_equal(g["indices"], g["indices_f64"])          # the verbatim fp32 run drew the same cells


def _calibration_inputs(seed=3):
    rng = np.random.default_rng(seed)
    um = rng.uniform(0, 1, (3, 256, 256))
    um[0, :50] = 0.05 * 3
    um[1, :20] = 0.0
    um[2, 100:120] = 1.0
    pred = (rng.uniform(size=um.shape) < 0.4).astype(np.uint8)
    gt = (rng.uniform(size=um.shape) < 0.4).astype(np.uint8)
    return um, pred, gt


def test_calibration_matches_reference():
    g = np.load(os.path.join(G, "calibration.npz"))
    um, pred, gt = _calibration_inputs(int(g["seed"]))
    c, a = port.pixel_confidences_and_accuracies(um, pred, gt)
    assert len(c) == int(g["n_fg"])
    ece, mce = port.calibration_ece(c, a)
    assert ece == float(g["ece"]) and mce == float(g["mce"])


def test_perimeter_and_gls_match_reference():
    g = np.load(os.path.join(G, "gls.npz"))
    np.testing.assert_allclose(port.perimeter(g["ed"]), g["perimeter_ed"], rtol=1e-13)
    np.testing.assert_allclose(port.perimeter(g["es"]), g["perimeter_es"], rtol=1e-13)
    got = np.array([port.global_longitudinal_strain(g["ed"][i], g["es"][i]) for i in range(len(g["ed"]))])
    np.testing.assert_allclose(got, g["gls"], rtol=1e-11)


def test_segmentation_postprocessing_matches_reference():
    """oracle seg_postprocess vs SegmentationUncertaintyTask.predict_step executed on the same samples."""
    import synth

    g = np.load(os.path.join(G, "segpost.npz"))
    n, Te, Ta = (int(v) for v in g["shape"])
    samples = synth.segmentation_samples(n, Te * Ta, seed=int(g["seed"])).reshape(n, Te, Ta, 1, 256, 256)
    post, ent, pred_samples = port.seg_postprocess(samples)
    want = np.unpackbits(g["pred_samples"], axis=-1, bitorder="little").astype(bool)
    assert np.array_equal(pred_samples.reshape(n, Te * Ta, 256, 256), want)
    # torch's fp32 mean over the samples vs numpy's: a few ulp of the mean, amplified by |d entropy / d p| near 0 and 1
    assert np.abs(ent - g["entropy"]).max() < 2e-6
    assert ent.dtype == np.float32 and (ent[:, :10] == 0).all() and (ent[:, :, -10:] == 0).all()


def test_mask_gls_matches_reference():
    """oracle polar corners / base-less contour / EchoMeasure.gls vs the executed reference (find_contours restated)."""
    g = np.load(os.path.join(G, "glsmask.npz"))
    masks = _unpack(g["masks"]).astype(bool)
    for i, m in enumerate(masks):
        try:
            left, right, _ = port.polar_base_corners(m)
            c = port.endo_contour_without_base(m)
            length = port.lv_longitudinal_length(m)
        except (RuntimeError, IndexError):
            assert np.isnan(g["length"][i]), i
            continue
        assert np.array_equal(np.array([left, right]), g["corners"][i]), i
        assert len(c) == int(g["npts"][i]) and length == float(g["length"][i]), i
    T = g["gls"].shape[1]
    for v in range(2):
        for t in range(T):
            try:
                got = port.echo_gls(np.stack([masks[(2 * v) * T + t], masks[(2 * v + 1) * T + t]]))[1]
            except (RuntimeError, IndexError):
                got = np.nan
            assert (np.isnan(got) and np.isnan(g["gls"][v, t])) or got == g["gls"][v, t]


def test_generic_size_reconstruction_and_lung_masks_match_reference():
    g = np.load(os.path.join(G, "generic.npz"))
    for k in (4, 5, 8, 26, 44, 50, 64):
        c = g[f"contours_{k}"]
        spl, lin = _unpack(g[f"spline_{k}"]), _unpack(g[f"linear_{k}"])
        for i in range(len(c)):
            assert np.array_equal(port.reconstruction(c[i], 256, 256), spl[i]), (k, i)
            assert np.array_equal(port.linear_reconstruction(c[i], 256, 256), lin[i]), (k, i)
    for i, lm in enumerate(g["lung_landmarks"]):
        assert np.array_equal(port.lung_contour_to_mask(lm), g["lung_argmax"][i])
        assert np.array_equal(port.lung_contour_to_mask(lm, False, "linear"), g["lung_stack_linear"][i])


def test_two_structure_masks_match_reference():
    """oracle USContourToMask (LV + MYO) vs the executed reference (skimage.draw.polygon restated: parity unpinned)."""
    g = np.load(os.path.join(G, "myo.npz"))
    stack = _unpack(g["stack_linear"])
    for i, lm in enumerate(g["landmarks"]):
        assert np.array_equal(port.us_contour_to_mask_myo(lm), g["argmax"][i]), i
        if i % 3 == 0:
            assert np.array_equal(port.us_contour_to_mask_myo(lm, False, "linear"), stack[i]), i


def test_restated_skimage_functions_reproduce_their_published_docstring_examples():
    """scikit-image is absent here, so its three functions on the path are restated (parity unpinned).  The examples
    printed in their own docstrings -- quoted from memory, the only vectors of theirs available offline -- come out right."""
    img = np.zeros((10, 10), dtype=np.uint8)
    rr, cc = port.draw_polygon(np.array([1, 2, 8]), np.array([1, 7, 4]))          # skimage.draw.polygon docstring
    img[rr, cc] = 1
    want = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                     [0, 1, 0, 0, 0, 0, 0, 0, 0, 0],
                     [0, 0, 1, 1, 1, 1, 1, 1, 0, 0],
                     [0, 0, 1, 1, 1, 1, 1, 0, 0, 0],
                     [0, 0, 0, 1, 1, 1, 1, 0, 0, 0],
                     [0, 0, 0, 1, 1, 1, 0, 0, 0, 0],
                     [0, 0, 0, 0, 1, 1, 0, 0, 0, 0],
                     [0, 0, 0, 0, 1, 0, 0, 0, 0, 0],
                     [0, 0, 0, 0, 1, 0, 0, 0, 0, 0],
                     [0, 0, 0, 0, 0, 0, 0, 0, 0, 0]], dtype=np.uint8)
    assert np.array_equal(img, want)
    a = np.zeros((3, 3))
    a[0, 0] = 1
    c = port.find_contours(a, 0.5)                                                 # skimage.measure.find_contours docstring
    assert len(c) == 1 and np.array_equal(c[0], np.array([[0.0, 0.5], [0.5, 0.0]]))
    rr, cc = port.bresenham_line(1, 1, 8, 8)                                       # skimage.draw.line docstring
    assert np.array_equal(rr, np.arange(1, 9)) and np.array_equal(cc, np.arange(1, 9))


def test_class_entropy_matches_reference():
    """sample_entropy on probabilities (C = 3 in base 3, C = 1 in bits) vs the executed reference."""
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    import importlib.util

    spec = importlib.util.spec_from_file_location("_mk_inputs", os.path.join(sys_path, "class_entropy_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    p3, p1 = mod.class_entropy_inputs()
    g = np.load(os.path.join(G, "class_entropy.npz"))
    e3 = np.array([port.sample_entropy_general(x) for x in p3])
    e1 = np.array([port.sample_entropy_general(x.reshape(3, 4, 1, 64, 64)) for x in p1])
    # torch's fp32 mean vs numpy's: a few ulp of the mean
    assert np.abs(e3 - g["e3"]).max() < 2e-6 and np.abs(e1 - g["e1"]).max() < 2e-6


def _point_in_polygon_scalar(xp, yp, x, y):
    """The scalar loop draw_polygon's row-vectorised form stands for (0 outside, 1 inside, 2 vertex, 3 edge)."""
    n = len(xp)
    x1, y1 = xp[n - 1] - x, yp[n - 1] - y
    l_cross = r_cross = 0
    for i in range(n):
        x0, y0 = xp[i] - x, yp[i] - y
        if -1e-12 < x0 < 1e-12 and -1e-12 < y0 < 1e-12:
            return 2
        if (y0 > 0) != (y1 > 0) and (x0 * y1 - x1 * y0) / (y1 - y0) > 0:
            r_cross += 1
        if (y0 < 0) != (y1 < 0) and (x0 * y1 - x1 * y0) / (y1 - y0) < 0:
            l_cross += 1
        x1, y1 = x0, y0
    if (r_cross & 1) != (l_cross & 1):
        return 3
    return 1 if r_cross & 1 else 0


def test_polygon_fill_vectorisation_and_blob_filter_properties():
    rng = np.random.default_rng(8)
    for _ in range(60):
        k = int(rng.integers(3, 9))
        r, c = rng.integers(-5, 26, k).astype(float), rng.integers(-5, 26, k).astype(float)
        rr, cc = port.draw_polygon(r, c, (20, 20))
        got = np.zeros((20, 20), bool)
        got[rr, cc] = True
        want = np.zeros((20, 20), bool)
        for y in range(int(max(0, r.min())), min(19, int(np.ceil(r.max()))) + 1):
            for x in range(int(max(0, c.min())), min(19, int(np.ceil(c.max()))) + 1):
                want[y, x] = _point_in_polygon_scalar(c, r, float(x), float(y)) != 0
        assert np.array_equal(got, want)
    for dens in (0.0, 0.3, 0.6, 1.0):
        seg = (rng.random((1, 40, 40)) < dens).astype(float)
        out = port.big_blob(seg)
        assert np.array_equal(port.big_blob(out), out)                   # idempotent
        assert not (out.astype(bool) & ~seg.astype(bool)).any()          # never adds pixels
        if seg.any():
            from scipy import ndimage

            assert ndimage.label(out[0])[1] == 1                         # one component left


# ------------------------------------------------------------------------------------------------
# The contour task end to end: tests/golden/task_al.npz / task_skew.npz hold every BatchResult field of the reference's own
# predict_step (DSNTAleatoric / DSNTSkew -> _predict_step -> ContourUncertaintyTask.predict_step, executed verbatim)
# ------------------------------------------------------------------------------------------------
def _check_task_result(g, tag, r, umap_exact=True):
    ps = _unpack(g[f"{tag}_pred_samples"]).reshape(g[f"{tag}_pred_samples_shape"])
    assert np.array_equal(ps, r["pred_samples"])
    assert np.array_equal(r["pred"], g[f"{tag}_pred"])
    for k in ("mu", "mode", "cov", "post_mu", "post_cov"):
        np.testing.assert_allclose(r[k], g[f"{tag}_{k}"], rtol=1e-12, atol=1e-12, err_msg=k)
    assert np.abs(r["entropy_map"] - g[f"{tag}_entropy_map"]).max() <= 1e-15
    assert (r["uncertainty_map"] != g[f"{tag}_uncertainty_map"]).sum() == 0          # 0 differing pixels
    for k, v in r["point_uncertainty"].items():
        np.testing.assert_allclose(v, g[f"{tag}_point_{k}"], rtol=1e-12, atol=0, err_msg=k)
    for k, v in r["instant_uncertainty"].items():
        # cov_projection: the reference sums 21 fp32 values in fp32 (np.sum of a float32 array)
        np.testing.assert_allclose(v, g[f"{tag}_instant_{k}"], rtol=2e-7 if k == "cov_projection" else 1e-12, atol=0, err_msg=k)
    assert set(r["instant_uncertainty"]) == {"cov_det_mean", "cov_eigenvalue_mean", "cov_projection", "umap_mean", "entropy_mean"}


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_aleatoric_task_predict_step_matches_reference(tag):
    """aleatoric.py:90-135 + contour_uncertainty.py:71-130 restated == the executed reference, given its own per-pass moments
    and contour samples; pins port.aleatoric_epistemic_moments / posterior_moments / point_and_instant_uncertainty."""
    g = np.load(os.path.join(G, "task_al.npz"))
    mu, cov = port.aleatoric_epistemic_moments(g["mu_te"], g["cov_te"])
    assert np.array_equal(mu, g[f"{tag}_mu"]) and np.array_equal(cov, g[f"{tag}_cov"])
    pm, pc = port.posterior_moments(g[f"{tag}_contour_samples"])
    np.testing.assert_allclose(pm, g[f"{tag}_post_mu"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(pc, g[f"{tag}_post_cov"], rtol=1e-12, atol=1e-12)
    _check_task_result(g, tag, port.aleatoric_predict_step(g["mu_te"], g["cov_te"], g[f"{tag}_contour_samples"]))


def test_aleatoric_task_sampler_and_heads_match_reference():
    """The parts before the orchestration, on the task's own inputs: DSNT heads of both passes and the PSM samples with
    the task's draw order (images outer, passes inner, aleatoric.py:71-73)."""
    g = np.load(os.path.join(G, "task_al.npz"))
    n, te, ta = (int(v) for v in g["shape"])
    base, other = synth.heatmap_logits(n, seed=int(g["seed"])), synth.heatmap_logits(n, seed=int(g["seed"]) + 1)
    logits = np.stack([base, 0.7 * base + 0.3 * other]).astype(np.float32)
    for e in range(te):
        mu, sig = port.dsnt_predict_on_batch(logits[e])
        np.testing.assert_allclose(mu, g["mu_te"][:, e], rtol=0, atol=8e-4)
        np.testing.assert_allclose(sig, g["cov_te"][:, e], rtol=2e-5, atol=2e-4)
    model = port.ShapeModel.load(os.path.join(os.path.dirname(G), "..", "contouring_uncertainty_b200", "data", "camus-cont_psm_11_no_std.npy"))
    s = np.stack([[port.psm_sample(model, g["mu_te"][i, e], g["cov_te"][i, e], g["z"][i, e]) for e in range(te)] for i in range(n)])
    assert np.abs(s - g["f64_contour_samples"]).max() < 1e-4       # vs the reference's sampler executed in fp64
    assert np.abs(s - g["f32_contour_samples"]).max() < 6e-2       # vs its fp32 run (its own noise: 4.3e-2 px here)
    assert np.array_equal(np.round(s), np.round(g["f64_contour_samples"]))   # hence the same masks as the fp64 run


def test_skew_task_predict_step_matches_reference():
    """aleatoric_skew.py:65-127 + contour_uncertainty.py:71-130 restated == the executed reference."""
    g = np.load(os.path.join(G, "task_skew.npz"))
    r = port.skew_predict_step(g["mu_te"], g["cov_te"], g["alpha_te"], g["f32_contour_samples"])
    np.testing.assert_allclose(r["alpha"], g["f32_alpha"], rtol=1e-12, atol=0)
    _check_task_result(g, "f32", r)
    # DSNTSkew.predict_on_batch negates the y component of the raw skewness (dsnt_skew.py:164)
    assert np.array_equal(g["alpha_te"], np.swapaxes(g["raw_alpha"], 0, 1) * np.array([1.0, -1.0], dtype=np.float32))


def test_endo_base_gls_matches_reference():
    """EchoMeasure._endo_base / _endo_epi_contour / gls with myo_labels (measure.py:153-222, 320-366) executed verbatim on
    43 LV + MYO label maps (glsmyo.npz): markers and point counts identical, lengths and strains bit-identical."""
    g = np.load(os.path.join(G, "glsmyo.npz"))
    maps = g["maps"]
    raised = 0
    for i, m in enumerate(maps):
        try:
            left, right = port.endo_base_corners(m)
        except RuntimeError:
            raised += 1
            assert np.isnan(g["length"][i]) and (g["corners"][i] == -1).all()
            continue
        assert left.tolist() == g["corners"][i, 0].tolist() and right.tolist() == g["corners"][i, 1].tolist()
        assert len(port.endo_contour_without_base_myo(m)) == g["npts"][i]
        assert port.lv_longitudinal_length_myo(m, (0.7, 0.4)) == g["length"][i]
    assert raised == 2
    for (a, b), gl in zip(g["pairs"], g["gls"]):
        assert port.echo_gls_myo(np.stack([maps[a], maps[b]]), (0.7, 0.4))[1] == gl


def test_lung_umaps_match_reference():
    """LungUMap / LungSkewUmap (cu/data/lung/utils.py:72-114) executed verbatim: skew_umap on 44 / 50 / 26-point contours."""
    g = np.load(os.path.join(G, "lung_umap.npz"))
    i = 0
    mu, cov, al = g["mu"][i], g["cov"][i], g["alpha"][i]
    mode, sk = port.lung_skew_umap(mu, cov, al)
    assert np.abs(mode - g["mode"][i]).max() < 1e-4 and (np.abs(sk - _sparse_map(g, "skew_umap", i)) > 1e-12).sum() == 0
    _, um = port.lung_skew_umap(mu, cov, np.zeros_like(mu))
    # the reference rotates the covariances in fp32 torch; the oracle forms the fp64 value and rounds once: a 1-ulp (fp32)
    # difference of a projected deviation moves one level contour by ~1e-6 px, which can flip a pixel of ONE of the 200 masks
    d = np.abs(um - _sparse_map(g, "umap", i))
    assert (d > 1e-12).sum() <= 4 and d.max() < 1e-3

"""GPU parity tests of Simpson's biplane volumes (contour tracing, base chord, discs, EDV/ESV/EF) through the
C ABI, against the goldens made by executing the reference (tests/golden/simpson.npz) and the oracle."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import synth
from oracle import port

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _pack(masks):
    """(..., 256, 256) 0/1 -> (..., 2048) int32 bit-packed (bit x & 31 of word x >> 5)."""
    b = np.packbits(masks.astype(np.uint8), axis=-1, bitorder="little")
    return np.ascontiguousarray(b).view(np.uint32).reshape(masks.shape[:-2] + (2048,)).view(np.int32)


def test_simpson_golden_exact(dev):
    """Diameters / step are fp64-identical to the executed reference, EDV / ESV integer-identical, EF identical
    (same two integers divided), NaN where the reference rejects."""
    from contouring_uncertainty_b200 import engine as E

    g = np.load(os.path.join(G, "simpson.npz"))
    contours = torch.from_numpy(g["contours"]).to(dev)               # (4, T, 21, 2)
    T = contours.shape[1]
    spacing = float(g["voxelspacing"][0])
    masks, area = E.rasterize_contours(contours.reshape(-1, 21, 2), return_masks=True)
    assert np.array_equal(area.cpu().numpy().reshape(4, T), g["areas"])
    diam, step, npts = E.simpson_diameters(masks.reshape(4, T, 2048), spacing, return_npts=True)
    assert np.array_equal(diam.cpu().numpy(), g["diameters"])
    assert np.array_equal(step.cpu().numpy(), g["step"])
    edv, esv, ef = E.simpson_volumes(diam[None], step[None])
    assert np.array_equal(edv.cpu().numpy()[0], g["edv"]) and np.array_equal(esv.cpu().numpy()[0], g["esv"])
    assert np.array_equal(ef.cpu().numpy()[0], g["ef"], equal_nan=True)
    e2 = E.biplane_volumes_from_contours(contours[None].contiguous(), spacing)
    assert torch.equal(e2[0], edv) and torch.equal(e2[1], esv)
    # contour lengths = the oracle's marching squares
    m_np = E.unpack_masks(masks).cpu().numpy().reshape(4, T, 256, 256)
    for v, t in ((0, 0), (2, T - 2), (3, T - 1), (1, T - 4)):
        assert npts[v, t].item() == len(port.find_contours(m_np[v, t], 0.5)[0])


def test_simpson_random_masks_vs_oracle(dev):
    """PSM samples of a synthetic patient + hand-made masks (empty, full frame, two blobs, frame-cut blob)."""
    from contouring_uncertainty_b200 import engine as E

    rng = np.random.default_rng(17)
    b = synth.LV_EXAMPLE_SHAPE
    T = 6
    cont = np.stack([np.stack([b.mean(0) + (b - b.mean(0)) * s + rng.normal(0, 3.0, (21, 2)) for _ in range(T)])
                     for s in (1.0, 0.8, 1.05, 0.85)]).astype(np.float32)
    masks = port.convert_to_mask(cont)                                # (4, T, 256, 256)
    masks[0, 0] = 0                                                    # empty: no contour -> zeros
    masks[1, 1] = 1                                                    # full frame: no contour either
    masks[2, 2] = 0
    masks[2, 2, 40:90, 50:100] = 1
    masks[2, 2, 150:220, 120:200] = 1                                  # two blobs: [0] is the raster-first one
    masks[3, 3] = 0
    masks[3, 3, 200:256, 60:140] = 1                                   # cut by the bottom border: open contour
    masks[0, 4] = 0
    masks[0, 4, 100, 100] = 1                                          # single pixel
    pk = torch.from_numpy(_pack(masks)).to(dev)
    diam, step, npts = E.simpson_diameters(pk, 0.5, return_npts=True)
    edv, esv, ef = E.simpson_volumes(diam[None], step[None])
    o_edv, o_esv, o_ef = port.volumes_mc(masks[0], masks[1], masks[2], masks[3], (0.5, 0.5), (0.5, 0.5))
    assert np.array_equal(edv.cpu().numpy()[0], o_edv) and np.array_equal(esv.cpu().numpy()[0], o_esv)
    assert np.array_equal(ef.cpu().numpy()[0], o_ef, equal_nan=True)
    dn, sn, nn = diam.cpu().numpy(), step.cpu().numpy(), npts.cpu().numpy()
    for v in range(4):
        for t in range(T):
            try:
                d, s = port.compute_diameters(masks[v, t], (0.5, 0.5))
            except Exception:
                assert np.isnan(dn[v, t]).all() and np.isnan(sn[v, t])
                continue
            assert np.array_equal(dn[v, t], d) and sn[v, t] == s, (v, t)
            assert nn[v, t] == len(port.find_contours(masks[v, t], 0.5)[0])
    assert nn[0, 0] == 0 and nn[1, 1] == 0


def test_simpson_full_size_properties(dev):
    """BASELINE config 4 sample count (10 000 per image) for one patient: size-independent invariants."""
    from contouring_uncertainty_b200 import engine as E

    T = 10000
    mus, covs = synth.point_distributions(4, seed=401)
    ctr = mus[0].mean(0)
    for v, s in enumerate((1.0, 0.8, 1.05, 0.84)):
        mus[v] = ctr + (mus[0] - ctr) * s
    model = E.ShapeModelConstants(device=dev)
    m, c = torch.from_numpy(mus).to(dev), torch.from_numpy(covs).to(dev)
    contours = E.sample_contours(m, E.psm_factor(model, m, c), T, seed=3)               # (4, T, 21, 2)
    edv, esv, ef, area = E.biplane_volumes_from_contours(contours[None].contiguous(), 0.5)
    edv, esv, ef = edv[0], esv[0], ef[0]
    ok = ~ef.isnan()
    assert float(ok.float().mean()) > 0.99
    # volumes are integers (Python round), EF is the ratio of the two integers, NaN exactly outside [0, 1]
    assert torch.equal(edv, edv.round()) and torch.equal(esv, esv.round()) and float(edv.min()) > 0
    assert torch.equal(ef[ok], ((edv - esv) / edv)[ok])
    bad = ((edv - esv) / edv < 0) | ((edv - esv) / edv > 1)
    assert torch.equal(bad, ~ok)
    # larger masks give larger volumes: sample volumes correlate with the product of the two ED areas
    a = (area[0, 0].double() * area[0, 2].double())
    assert float(torch.corrcoef(torch.stack([a, edv]))[0, 1]) > 0.8
    # sub-sample through the mask entry points (order / indexing): same diameters
    sel = contours[:, ::1000].contiguous()
    masks, _ = E.rasterize_contours(sel.reshape(-1, 21, 2), return_masks=True)
    d1, s1 = E.simpson_diameters(masks.reshape(4, 10, 2048), 0.5)
    e1 = E.simpson_volumes(d1[None], s1[None])
    assert torch.equal(e1[0][0], edv[::1000]) and torch.equal(e1[1][0], esv[::1000])
    # spacing is a pure scale: diameters and step scale linearly
    d2, s2 = E.simpson_diameters(masks.reshape(4, 10, 2048), 1.0)
    assert torch.allclose(d2, 2 * d1, rtol=1e-15) and torch.allclose(s2, 2 * s1, rtol=1e-15)


def test_simpson_anisotropic_spacing_vs_reference(dev):
    """Anisotropic voxel spacing (the real CAMUS case): the reference stretches every mask with PIL's nearest-neighbour
    resize before measuring it; the kernels see the stretched mask through index tables.  Diameters / step fp64-identical
    to the executed reference, EDV / ESV identical."""
    from contouring_uncertainty_b200 import engine as E
    from PIL import Image

    g = np.load(os.path.join(G, "simpson.npz"))
    contours = torch.from_numpy(g["contours"]).to(dev)
    T = contours.shape[1]
    vs2, vs4 = tuple(g["vs_a2c"]), tuple(g["vs_a4c"])
    masks, _ = E.rasterize_contours(contours.reshape(-1, 21, 2), return_masks=True)
    masks = masks.reshape(4, T, 2048)
    d2, s2 = E.simpson_diameters(masks[:2].contiguous(), vs2)
    d4, s4 = E.simpson_diameters(masks[2:].contiguous(), vs4)
    diam, step = torch.cat([d2, d4]), torch.cat([s2, s4])
    assert np.array_equal(diam.cpu().numpy(), g["diameters_aniso"])
    assert np.array_equal(step.cpu().numpy(), g["step_aniso"])
    edv, esv, ef = E.simpson_volumes(diam[None], step[None])
    assert np.array_equal(edv.cpu().numpy()[0], g["edv_aniso"]) and np.array_equal(esv.cpu().numpy()[0], g["esv_aniso"])
    e2 = E.biplane_volumes_from_contours(contours[None].contiguous(), vs2, vs4)
    assert torch.equal(e2[0], edv) and torch.equal(e2[1], esv)
    # the index tables ARE PIL's nearest-neighbour resize
    rng = np.random.default_rng(0)
    img = rng.uniform(size=(256, 256)) < 0.5
    for nh, nw in ((366, 256), (256, 347), (300, 411)):
        want = np.array(Image.fromarray(img).resize((nw, nh), resample=Image.NEAREST))
        assert np.array_equal(img[np.ix_(E.pil_nearest_index(256, nh), E.pil_nearest_index(256, nw))], want)
    # isotropic spacing given as a pair takes the plain path
    d0, s0 = E.simpson_diameters(masks[:1].contiguous(), (0.5, 0.5))
    d1, s1 = E.simpson_diameters(masks[:1].contiguous(), 0.5)
    assert torch.equal(d0, d1) and torch.equal(s0, s1)

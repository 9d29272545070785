"""GPU parity tests of the analytic uncertainty maps (projected uncertainty, uncertainty_map, skew_umap) through the
C ABI, against the oracle and the goldens made by executing the reference (tests/golden/umap.npz)."""
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


def _sparse_map(g, name, i):
    m = np.zeros(256 * 256)
    m[g[f"{name}_idx_{i}"]] = g[f"{name}_val_{i}"]
    return m.reshape(256, 256)


def _inputs(dev, n, seed):
    mus, covs = synth.point_distributions(n, seed=seed)
    al = synth.skew_alphas(n, seed=seed + 1)
    t = lambda a: torch.from_numpy(a).to(dev)
    return mus, covs, al, t(mus), t(covs), t(al)


def test_projected_uncertainty_vs_oracle_and_reference(dev):
    from contouring_uncertainty_b200 import engine as E

    g = np.load(os.path.join(G, "umap.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    mu, cov, al = t(g["mu"]), t(g["cov"]), t(g["alpha"])
    u, v = E.projected_uncertainty(mu, cov, all=True)
    ulps = lambda a, b: np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32)).max()
    assert ulps(u.cpu().numpy(), g["unc_all"]) <= 1                      # fp32 rotation in the reference
    assert np.abs(v.cpu().numpy() - g["proj_all"]).max() < 1e-11
    u0, _ = E.projected_uncertainty(mu, cov)
    np.testing.assert_allclose(u0.cpu().numpy(), g["unc"], rtol=3e-7)
    np.testing.assert_allclose(u0.sum(1).cpu().numpy(), g["value"], rtol=1e-6)   # projected_uncertainty_value
    u2, v2, ap = E.projected_uncertainty(mu, cov, al, all=True, linear_close=True)
    assert ulps(u2.cpu().numpy(), g["skew_unc"]) <= 1
    assert np.abs(v2.cpu().numpy() - g["skew_proj"]).max() < 3e-7       # end points: fp32 linear_close vectors
    assert np.abs(v2.cpu().numpy()[:, 1:-1] - g["skew_proj"][:, 1:-1]).max() < 1e-11
    assert np.abs(ap.cpu().numpy() - g["alpha_proj"]).max() < 2e-6
    # oracle on other inputs: identical fp32 values
    mus, covs, als, tm, tc, ta = _inputs(dev, 6, 201)
    u, v, ap = E.projected_uncertainty(tm, tc, ta, all=True, linear_close=True)
    for i in range(6):
        uo, vo, apo = port.projected_uncertainty(mus[i], covs[i], als[i].copy(), all=True, linear_close=True)
        assert ulps(u[i].cpu().numpy(), uo) <= 1 and np.abs(v[i].cpu().numpy() - vo).max() < 3e-7
        assert np.abs(ap[i].cpu().numpy() - apo).max() < 2e-6


def test_umap_analytic_vs_reference_and_oracle(dev):
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.data.camus.utils import USUMap

    g = np.load(os.path.join(G, "umap.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    um = E.umap_analytic(t(g["mu"]), t(g["cov"])).cpu().numpy()
    for i in range(3):
        ref = _sparse_map(g, "umap", i)
        assert (um[i] != ref).sum() == 0                       # pixel-identical to the executed reference (golden)
        assert um[i].max() == 1.0
    assert np.array_equal(USUMap.__call__(g["mu"][0], g["cov"][0], labels=[]), um[0])
    mus, covs, _, tm, tc, _ = _inputs(dev, 3, 211)
    raw = E.umap_analytic(tm, tc, normalize=False).cpu().numpy()
    for i in range(3):
        # random inputs vs the oracle: the kernel forms u in fp64 and rounds once, the oracle (like the reference) in fp32
        # torch -- a 1-ulp (fp32) difference of u can move a rounded spline pixel of one offset contour
        assert (raw[i] != port.uncertainty_map(mus[i], covs[i])).sum() <= 8


def test_skew_umap_vs_reference_and_oracle(dev):
    from contouring_uncertainty_b200 import engine as E

    g = np.load(os.path.join(G, "umap.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    mode, sk, contours = E.skew_umap(t(g["mu"]), t(g["cov"]), t(g["alpha"]), return_contours=True)
    mode, sk = mode.cpu().numpy(), sk.cpu().numpy()
    assert np.abs(mode - g["mode"]).max() < 1e-4
    for i in range(3):
        ref = _sparse_map(g, "skew_umap", i)
        assert (np.abs(sk[i] - ref) > 1e-12).sum() == 0        # golden: every pixel within 1e-12
    # level contours against the oracle (fp64): identical level indices -> positions equal to fp32-ulp-of-u precision
    mo, co, w = port.skew_umap_contours(g["mu"][0], g["cov"][0], g["alpha"][0], linear_close=True)
    assert np.abs(contours[0].cpu().numpy() - co).max() < 2e-5
    assert np.allclose(w, E.SKEW_UMAP_WEIGHTS, rtol=0, atol=0)


def test_lung_umaps_vs_reference(dev):
    """LungUMap / LungSkewUmap (cu/data/lung/utils.py:72-114, executed reference): the general-size projected uncertainty
    (44 / 50 / 26-point splines), level contours, generic rasteriser and weighted entropy."""
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.data.lung.utils import LungSkewUmap, LungUMap

    g = np.load(os.path.join(G, "lung_umap.npz"))
    for i in range(2):
        mu, cov, al = g["mu"][i], g["cov"][i], g["alpha"][i]
        mode, sk = LungSkewUmap.__call__(mu, cov, al)
        assert mode.shape == (120, 2) and np.abs(mode - g["mode"][i]).max() < 1e-4
        # 1 fp32 ulp of a projected deviation (the reference rotates covariances in fp32 torch) can flip a pixel of ONE of the
        # 200 level masks: <= 4 pixels may differ, by one mask's weight (< 1e-3)
        d = np.abs(sk - _sparse_map(g, "skew_umap", i))
        assert (d > 1e-12).sum() <= 4 and d.max() < 1e-3
        um = LungUMap.__call__(mu, cov)
        d = np.abs(um - _sparse_map(g, "umap", i))
        assert (d > 1e-12).sum() <= 4 and d.max() < 1e-3 and um.max() <= 1.0
    # the general-size kernel on 21 points = the specialised one (different solver: cyclic reduction vs Thomas)
    mus, covs, als, tm, tc, ta = _inputs(dev, 3, 212)
    u, v, ap = E.projected_uncertainty(tm, tc, ta, all=True, linear_close=True)
    L = E.nat.lib()
    u2, v2, ap2 = torch.empty_like(u), torch.empty_like(v), torch.empty_like(ap)
    E.nat.check(L.cub200_projected_uncertainty_generic(tm.data_ptr(), tc.data_ptr(), ta.data_ptr(), 3, 21, 1, 1, u2.data_ptr(), v2.data_ptr(),
                                                       ap2.data_ptr(), None), "generic")
    torch.cuda.synchronize()
    assert (u - u2).abs().max().item() <= 1e-6 * u.abs().max().item() and (v - v2).abs().max().item() < 1e-9
    assert (ap - ap2).abs().max().item() < 1e-5

"""GPU parity tests of the skew-normal path (task dsnt-skew): pdf tables, shape-model conditionals and
the numerical (multinomial) PSM sampler, through the C ABI, against the oracle and the goldens made by
executing the reference (tests/golden/skew.npz)."""
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


@pytest.fixture(scope="module")
def model(dev):
    from contouring_uncertainty_b200 import engine as E

    return E.ShapeModelConstants(device=dev)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "skew.npz"))


def _expo(g, n_img, n_s):
    return np.stack([np.stack([np.stack([synth.exponential_field(int(g["q_seed"]), (i * n_s + s) * 14 + d)
                                         for d in range(14)]) for s in range(n_s)]) for i in range(n_img)])


def test_skew_pdf_tables_vs_oracle_and_reference(dev, gold):
    from contouring_uncertainty_b200 import engine as E

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tab = E.skew_pdf_grid(t(gold["mu"]), t(gold["cov"]), t(gold["alpha"]), flip_alpha_y=True).cpu().numpy()
    al = gold["alpha"] * np.array([1.0, -1.0])
    for i in range(3):
        want = port.skew_pdf_grid(gold["mu"][i], gold["cov"][i], al[i])
        # fp64 kernel rounded to fp32: half an fp32 ulp + denormal floor
        assert np.abs(tab[i] - want).max() <= 1e-7 * want.max()
        big = want > 1e-30
        assert (np.abs(tab[i] - want) / want)[big].max() < 1e-6
        ref = gold["pdf_sub"][i]                                # the verbatim fp32 reference
        assert np.abs(tab[i][:, 3::16, 5::16] - ref).max() < 2e-8
    # the un-flipped call is BivariateSkewNormal.pdf itself
    tab2 = E.skew_pdf_grid(t(gold["mu"][0]), t(gold["cov"][0]), t(gold["alpha"][0])).cpu().numpy()
    want2 = port.skew_pdf_grid(gold["mu"][0], gold["cov"][0], gold["alpha"][0].astype(np.float64))
    assert np.abs(tab2 - want2).max() <= 1e-7 * want2.max()


def test_psm_conditionals_vs_oracle(dev, model, gold):
    from contouring_uncertainty_b200 import engine as E

    order = [0, 10, 20, 5, 15, 2, 7, 13, 18, 1, 3, 6, 8, 12, 14, 17, 19, 4, 9, 11, 16]
    mu = gold["mu"]
    mc, cc = E.psm_conditionals(model, torch.from_numpy(mu).to(dev))
    mc, cc = mc.cpu().numpy(), cc.cpu().numpy()
    m = port.ShapeModel.load(model.path)
    rng = np.random.default_rng(5)
    for i in range(3):
        mu_t = (mu[i].astype(np.float64).reshape(-1) - m.mean) / m.scale
        S = port.pca_cov(m.X_train, mu_t)
        s = mu[i].astype(np.float64) + rng.normal(0, 3.0, (21, 2))
        s_t = (s.reshape(-1) - m.mean) / m.scale
        dev_perm = np.array([s[order[v // 2], v % 2] - mu[i][order[v // 2], v % 2] for v in range(34)], dtype=np.float64)
        for (p0, p1, sig) in ((3, 5, 1.0), (5, 9, 1.0), (9, 17, 1.0), (17, 21, 1e-3)):
            g = port.index_to_flat(sorted(order[:p0]))
            mu_c, cov_c = port.posterior_shape_model(s_t, g, mu_t, S, sig)
            mu_c = (mu_c * m.scale + m.mean).reshape(21, 2)
            cov_c = cov_c * m.scale[None, :]
            for pi in range(p0, p1):
                j = order[pi]
                for c in range(2):
                    v = 2 * pi + c
                    got = mu[i][j, c] + mc[i][:, v - 6] @ dev_perm
                    assert abs(got - mu_c[j, c]) < 1e-8, (i, pi, c, got, mu_c[j, c])
                if pi < 17:
                    blk = cov_c[2 * j:2 * j + 2, 2 * j:2 * j + 2]
                    # Sherman-Morrison form: Sig0 + d d^T - ... cancels terms of size |d|^2 ~ 1e3 -> ~1e-10 px^2 absolute
                    assert np.allclose(cc[i][pi - 3], [blk[0, 0], blk[1, 0], blk[1, 1]], rtol=1e-8, atol=1e-8)
        assert np.all(mc[i][6:, :4] == 0) and np.all(mc[i][10:, 4:12] == 0) and np.all(mc[i][18:, 12:28] == 0)


def test_skew_sampler_replay_vs_reference_and_oracle(dev, model, gold):
    """Random draws of the reference fed in: the 14 grid indices per sample must equal the reference's
    (fp64-executed run and the fp64 oracle) bit for bit; the float coordinates (3 rvs_fast points, 4 filled
    points) match to 2e-4 px (fp32 storage of 200-px coordinates).  Against the verbatim fp32 run, which
    carries its own fp32 shape-model noise (1e-4 relative on the race scores), indices may flip rarely:
    at least 95 % of the contours must agree exactly."""
    from contouring_uncertainty_b200 import engine as E

    n_img, n_s = gold["z"].shape[:2]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    expo = t(_expo(gold, n_img, n_s))
    out, idx = E.skew_sample(model, t(gold["mu"]), t(gold["cov"]), t(gold["alpha"]), n_s, z=t(gold["z"]), expo=expo,
                             return_indices=True)
    out, idx = out.cpu().numpy(), idx.cpu().numpy()
    ordinary = np.ones((n_img, n_s), dtype=bool)
    ordinary[2, n_s - 1] = False                      # the underflow sample: fp64 reference does not underflow
    assert np.array_equal(idx[ordinary], gold["indices_f64"][ordinary])
    assert np.abs(out[ordinary] - gold["samples_f64"][ordinary]).max() < 2e-4
    # the underflow sample follows the fp32 reference: same three fallbacks to mu_c, same indices elsewhere
    assert np.array_equal(idx[2, n_s - 1], gold["indices"][2, n_s - 1])
    assert (idx[2, n_s - 1] < 0).sum() == 3
    assert np.abs(out[2, n_s - 1] - gold["samples"][2, n_s - 1]).max() < 5e-2
    same = (idx == gold["indices"]).all(-1)
    assert same.mean() >= 0.95
    # oracle on other inputs (nothing stored): 2 images x 3 samples
    mus, covs = synth.point_distributions(2, seed=71)
    al = synth.skew_alphas(2, seed=72, scale=3.0)
    z = np.random.default_rng(73).standard_normal((2, 3, 3, 3)).astype(np.float32)
    ex = np.stack([np.stack([np.stack([synth.exponential_field(74, (i * 3 + s) * 14 + d) for d in range(14)])
                             for s in range(3)]) for i in range(2)])
    out2, idx2 = E.skew_sample(model, t(mus), t(covs), t(al), 3, z=t(z), expo=t(ex), return_indices=True)
    m = port.ShapeModel.load(model.path)
    for i in range(2):
        so, io = port.skew_psm_sample(m, mus[i], covs[i], al[i], z[i], lambda n, d, _i=i: ex[_i, n, d], store=np.float32,
                                      return_indices=True)
        assert np.array_equal(idx2[i].cpu().numpy(), io)
        assert np.abs(out2[i].cpu().numpy() - so).max() < 2e-4


def test_skew_sampler_philox_statistics(dev, model):
    """Production RNG: the in-kernel exponential race must draw from the product pdf.  For one image the
    empirical mean / covariance of a level-1 point over 20 000 samples is compared with the moments of the
    oracle's product pdf averaged over the initial-point draws (checked through the sampler's own indices)."""
    from contouring_uncertainty_b200 import engine as E

    mus, covs = synth.point_distributions(1, seed=81)
    al = synth.skew_alphas(1, seed=82)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    T = 20000
    out, idx = E.skew_sample(model, t(mus), t(covs), t(al), T, seed=5, return_indices=True)
    out2 = E.skew_sample(model, t(mus), t(covs), t(al), T, seed=5)
    out3 = E.skew_sample(model, t(mus), t(covs), t(al), T, seed=6)
    assert torch.equal(out, out2) and not torch.equal(out, out3)
    out, idx = out.cpu().numpy()[0], idx.cpu().numpy()[0]
    assert (idx >= 0).all()
    order = [5, 15, 2, 7, 13, 18, 1, 3, 6, 8, 12, 14, 17, 19]
    for d, j in enumerate(order):                       # indices and coordinates agree
        assert np.array_equal(out[:, j, 0], idx[:, d] // 256) and np.array_equal(out[:, j, 1], idx[:, d] % 256)
    # level-1 point 5 given the initial points: compare with exact conditional moments, sample by sample
    m = port.ShapeModel.load(model.path)
    mu, cov = mus[0].astype(np.float64), covs[0].astype(np.float64)
    alpha = al[0].astype(np.float64) * np.array([1.0, -1.0])
    mu_t = (mu.reshape(-1) - m.mean) / m.scale
    S = port.pca_cov(m.X_train, mu_t)
    p1 = port.skew_pdf_grid(mu, cov, alpha)[5]
    grid = port.skew_grid()
    zs = []
    for k in range(400):
        s = np.zeros((21, 2))
        s[[0, 10, 20]] = out[k, [0, 10, 20]]
        s_t = (s.reshape(-1) - m.mean) / m.scale
        mu_c, cov_c = port.posterior_shape_model(s_t, port.index_to_flat([0, 10, 20]), mu_t, S, 1.0)
        mu_c = (mu_c * m.scale + m.mean).reshape(21, 2)
        c2 = (cov_c * m.scale[None, :])[10:12, 10:12]
        p = p1 * np.exp(port.normal_logpdf(grid, mu_c[5], c2))
        p /= p.sum()
        mean = (p[..., None] * grid).sum((0, 1))
        var = (p[..., None] * (grid - mean) ** 2).sum((0, 1))
        zs.append((out[k, 5] - mean) / np.sqrt(var))
    zs = np.array(zs)                                   # standardised residuals: mean 0, variance 1
    assert np.abs(zs.mean(0)).max() < 4 / np.sqrt(400)
    assert np.abs(zs.var(0) - 1).max() < 0.25


def test_sequence_skew_sampler_replay_vs_reference(dev, model):
    """SequenceSkewPSMSampler with the reference's draws replayed: all 31 grid indices per sample equal the
    reference's (fp32 and fp64 runs agree on them), float coordinates to 2e-4 px."""
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.sampler import SequenceSkewPSMSampler

    g = np.load(os.path.join(G, "seqskew.npz"))
    n_pair, n_s = g["z"].shape[:2]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ex = np.stack([np.stack([np.stack([synth.exponential_field(int(g["q_seed"]), (i * n_s + s) * 31 + d) for d in range(31)])
                             for s in range(n_s)]) for i in range(n_pair)])
    seq = E.SequenceModelConstants(device=dev)
    out, idx = E.seq_skew_sample(model, seq, t(g["mu"]), t(g["cov"]), t(g["alpha"]), n_s, first_instant=t(g["bits"].astype(np.uint8)),
                                 z=t(g["z"]), expo=t(ex), return_indices=True)
    assert np.array_equal(idx.cpu().numpy(), g["indices_f64"])
    assert np.abs(out.cpu().numpy() - g["samples_f64"]).max() < 2e-4
    assert np.abs(out.cpu().numpy() - g["samples"]).max() < 5e-2          # the verbatim fp32 run's own noise on the filled points
    smp = SequenceSkewPSMSampler()
    o2 = smp(torch.from_numpy(g["mu"][0]), torch.from_numpy(g["cov"][0]), torch.from_numpy(g["alpha"][0]), n=n_s,
             first_instant=g["bits"][0], z=g["z"][0], expo=ex[0])
    assert o2.shape == (2, n_s, 21, 2) and np.array_equal(o2.numpy(), out[0].cpu().numpy())
    # production RNG: seeded, both instants integer-valued on the level points, ED/ES areas correlated
    o3 = E.seq_skew_sample(model, seq, t(g["mu"]), t(g["cov"]), t(g["alpha"]), 64, seed=3)
    o4 = E.seq_skew_sample(model, seq, t(g["mu"]), t(g["cov"]), t(g["alpha"]), 64, seed=3)
    assert o3.shape == (n_pair, 2, 64, 21, 2) and torch.isfinite(o3).all()
    lvl = [5, 15, 2, 7, 13, 18, 1, 3, 6, 8, 12, 14, 17, 19]
    assert torch.equal(o3[:, :, :, lvl], o3[:, :, :, lvl].round())


def test_skew_full_size_properties(dev, model):
    """BASELINE config 3 shape per GPU, scaled to 64 images x 1000 skew-PSM samples: size-independent invariants."""
    from contouring_uncertainty_b200 import engine as E

    n, T = 64, 1000
    mus, covs = synth.point_distributions(n, seed=301)
    al = synth.skew_alphas(n, seed=302)
    t = lambda a: torch.from_numpy(a).to(dev)
    out, idx = E.skew_sample(model, t(mus), t(covs), t(al), T, seed=17, return_indices=True)
    lvl = [5, 15, 2, 7, 13, 18, 1, 3, 6, 8, 12, 14, 17, 19]
    # every level point is a grid cell, equal to its recorded flat index; no fallback on ordinary inputs
    assert int(idx.min()) >= 0 and int(idx.max()) < 65536
    assert torch.equal(out[:, :, lvl, 0], (idx // 256).float()) and torch.equal(out[:, :, lvl, 1], (idx % 256).float())
    # idempotence / seeding: same seed -> same contours; the pdf tables can be passed in
    tabs = E.skew_pdf_grid(t(mus), t(covs), t(al), flip_alpha_y=True)
    assert torch.equal(E.skew_sample(model, t(mus), t(covs), t(al), T, seed=17, pdfs=tabs), out)
    # the sampled clouds sit on the predictions: a skew-normal's mean is at most sqrt(2/pi) sigma (<= 4.8 px here) from its
    # location parameter; spread bounded by the prediction's
    d = out.mean(1) - t(mus)
    assert float(d.abs().max()) < 6.0
    sd = out.std(1)
    lim = torch.sqrt(torch.stack([t(covs)[..., 0, 0], t(covs)[..., 1, 1]], -1)) * 1.6 + 0.6
    assert bool((sd[:, lvl] < lim[:, lvl]).all())
    # checksum of checksums through the rasteriser: per-image count-map sum == sum of per-sample areas
    diff = torch.zeros((n, 256, 256), dtype=torch.int32, device=dev)
    _, area = E.rasterize_contours(out.reshape(-1, 21, 2), return_masks=False, diff=diff, per_image=T)
    count, mean, ent, pred = E.finalize_maps(diff, T)
    assert torch.equal(count.long().sum((1, 2)), area.reshape(n, T).long().sum(1))
    assert torch.all(pred == (2 * count > T).to(torch.uint8))

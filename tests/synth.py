"""Seeded synthetic inputs shared by the golden generator and the tests (numpy, CPU, deterministic).

Shapes follow SURVEY.md section 8(d): K = 21 landmarks, 256x256 frames, landmark means around the
reference's ``LV_example_shape`` (contour_uncertainty/data/config.py:110-130).
"""
from __future__ import annotations

import numpy as np

K = 21
H = W = 256

LV_EXAMPLE_SHAPE = np.array(
    [[114., 205.], [105., 185.], [102., 167.], [102., 151.], [101., 134.], [98., 116.], [98., 100.],
     [99., 83.], [100., 67.], [106., 51.], [123., 35.], [139., 40.], [153., 55.], [168., 74.],
     [174., 92.], [179., 111.], [182., 128.], [184., 145.], [182., 163.], [179., 180.], [173., 200.]])


def random_spd(rng, n, lo=4.0, hi=36.0):
    """n random 2x2 SPD matrices R(theta) diag(l1, l2) R(theta)^T, l ~ U[lo, hi] px^2."""
    th = rng.uniform(0, np.pi, n)
    l1, l2 = rng.uniform(lo, hi, n), rng.uniform(lo, hi, n)
    c, s = np.cos(th), np.sin(th)
    out = np.zeros((n, 2, 2))
    out[:, 0, 0] = c * c * l1 + s * s * l2
    out[:, 0, 1] = out[:, 1, 0] = c * s * (l1 - l2)
    out[:, 1, 1] = s * s * l1 + c * c * l2
    return out


def point_distributions(n_img, seed):
    """(mu (n,K,2) fp32, cov (n,K,2,2) fp32): LV example shape + shift/jitter, random SPD covariances."""
    rng = np.random.default_rng(seed)
    mu = LV_EXAMPLE_SHAPE[None] + rng.normal(0, 4.0, (n_img, 1, 2)) + rng.normal(0, 1.5, (n_img, K, 2))
    cov = random_spd(rng, n_img * K).reshape(n_img, K, 2, 2)
    return mu.astype(np.float32), cov.astype(np.float32)


def heatmap_logits(n_img, seed, kind="bumps"):
    """(n_img, K, 256, 256) fp32 logits whose softmax is ~N(mu_ik, Sigma_ik) per landmark.

    kind: 'bumps'   logit = -1/2 d^T Sigma^-1 d + 0.05 N(0,1)           
          'noise'   a weak bump (x0.05) under N(0,1) logit noise          (broad, noisy heatmap)
          'bimodal' two equal bumps 30-60 px apart                        (re-centring stress)
    """
    rng = np.random.default_rng(seed)
    mu, cov = point_distributions(n_img, seed + 1000)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    out = np.empty((n_img, K, H, W), dtype=np.float32)
    for i in range(n_img):
        for k in range(K):
            P = np.linalg.inv(cov[i, k].astype(np.float64))

            def bump(m):
                dx, dy = xx - m[0], yy - m[1]
                return -0.5 * (P[0, 0] * dx * dx + 2 * P[0, 1] * dx * dy + P[1, 1] * dy * dy)

            if kind == "bumps":
                l = bump(mu[i, k]) + 0.05 * rng.standard_normal((H, W))
            elif kind == "noise":
                l = 0.05 * bump(mu[i, k]) + rng.standard_normal((H, W))
            elif kind == "bimodal":
                off = rng.uniform(30, 60) * np.array([np.cos(k), np.sin(k)])
                l = np.logaddexp(bump(mu[i, k]), bump(mu[i, k] + off)) + 0.05 * rng.standard_normal((H, W))
            else:
                raise ValueError(kind)
            out[i, k] = l.astype(np.float32)
    return out


def raster_cases(seed, n_random=160):
    """Contours (K,2) float64 exercising the reference's rasterisation rules.

    tags: 'example', 'random' (sigma 3 px), 'wild' (sigma 12 px: self-intersections), 'dup'
    (duplicate consecutive rounded landmarks -> splprep ValueError -> 21 isolated pixels, no
    spline), 'right'/'bottom' (clipped at 255), 'left'/'top' (negative indices wrap), 'far'
    (index < -256: IndexError -> spline not drawn), 'half' (x.5 ties, round-half-even), 'tiny'.
    """
    rng = np.random.default_rng(seed)
    base = LV_EXAMPLE_SHAPE
    cases, tags = [base.copy()], ["example"]
    for _ in range(n_random):
        cases.append(base + rng.normal(0, 3.0, (K, 2)) + rng.normal(0, 5.0, (1, 2)))
        tags.append("random")
    for _ in range(24):
        cases.append(base + rng.normal(0, 12.0, (K, 2)))
        tags.append("wild")
    for j in (0, 7, 19):
        c = base + rng.normal(0, 2.0, (K, 2))
        c[j + 1] = c[j] + 0.2
        cases.append(c)
        tags.append("dup")
    for dx, dy, t in ((90, 0, "right"), (0, 70, "bottom"), (-110, 0, "left"), (0, -60, "top"),
                      (-300, 0, "far"), (0, -340, "far"), (-200, -30, "left"), (75, 52, "right")):
        cases.append(base + np.array([dx, dy]) + rng.normal(0, 2.0, (K, 2)))
        tags.append(t)
    for _ in range(6):
        c = np.round(base + rng.normal(0, 3.0, (K, 2))) + 0.5
        cases.append(c)
        tags.append("half")
    ctr = base.mean(0)
    for s in (0.02, 0.1, 0.3):
        cases.append(ctr + (base - ctr) * s)
        tags.append("tiny")
    return cases, tags


def _splitmix64(x):
    """SplitMix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def exponential_field(seed, stream, n=H * W):
    """(n,) fp32 Exp(1) variates of replay stream ``stream`` -- the ``q`` of torch.multinomial's
    exponential race (argmax(p / q)).  Counter-based and numpy-version independent, so the 256 KB
    fields of the skew-sampler fixtures are regenerated from (seed, stream) instead of stored."""
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) + np.uint64(stream) * np.uint64(1 << 20)
        bits = _splitmix64(ctr ^ _splitmix64(np.full(n, seed, dtype=np.uint64)))
    u = ((bits >> np.uint64(40)).astype(np.float64) + 0.5) / float(1 << 24)
    return (-np.log(u)).astype(np.float32)


def skew_alphas(n_img, seed, scale=2.0):
    """(n_img, K, 2) fp32 skewness vectors ~ N(0, scale^2)."""
    return np.random.default_rng(seed).normal(0, scale, (n_img, K, 2)).astype(np.float32)


def segmentation_samples(n_img, T, seed, adversarial=True, speckle=0.002):
    """(n_img, T, 256, 256) float32 'network probabilities' of a binary segmentation: a soft blob per sample with
    dips (holes to fill), spurious islands (blobs to drop) and speckle.  With ``adversarial`` the first samples of
    image 0 are replaced by the corner cases of the post-processing (see the list below)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    out = np.empty((n_img, T, H, W), dtype=np.float32)
    for i in range(n_img):
        cy, cx = 128 + rng.normal(0, 12), 128 + rng.normal(0, 12)
        for t in range(T):
            a, b = rng.uniform(35, 70), rng.uniform(25, 50)
            th = rng.uniform(0, np.pi)
            u = (xx - cx - rng.normal(0, 3)) * np.cos(th) + (yy - cy - rng.normal(0, 3)) * np.sin(th)
            v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
            field = 6.0 * (1.0 - np.sqrt((u / a) ** 2 + (v / b) ** 2))
            for _ in range(rng.integers(0, 4)):                       # dips inside -> holes
                hy, hx, hs = cy + rng.normal(0, 15), cx + rng.normal(0, 15), rng.uniform(1.5, 6)
                field -= 9.0 * np.exp(-((yy - hy) ** 2 + (xx - hx) ** 2) / (2 * hs * hs))
            for _ in range(rng.integers(0, 5)):                       # islands outside -> extra blobs
                by, bx, bs = rng.uniform(5, 250), rng.uniform(5, 250), rng.uniform(1.0, 7)
                field += 9.0 * np.exp(-((yy - by) ** 2 + (xx - bx) ** 2) / (2 * bs * bs))
            p = 1.0 / (1.0 + np.exp(-field))
            speck = rng.random((H, W)) < speckle
            p = np.where(speck, 1.0 - p, p)
            out[i, t] = p.astype(np.float32)
    if adversarial and n_img > 0:
        cases = []
        cases.append(np.zeros((H, W)))                                  # nothing
        cases.append(np.ones((H, W)))                                   # everything
        c = np.zeros((H, W)); c[20:40, 30:50] = 0.9; c[100:120, 60:80] = 0.8    # two equal blobs -> the first one
        cases.append(c)
        c = np.zeros((H, W)); c[100:120, 60:80] = 0.8; c[20:40, 130:150] = 0.9; c[20, 150] = 0.7   # the later-met blob is larger
        cases.append(c)
        c = np.zeros((H, W)); c[50:100, 50:100] = 0.9; c[100:150, 100:150] = 0.9   # touching only diagonally
        cases.append(c)
        c = np.zeros((H, W)); c[60:200, 60:200] = 0.95; c[90:170, 90:170] = 0.1; c[120:140, 120:140] = 0.99   # ring with a core
        cases.append(c)
        c = np.zeros((H, W)); c[60:200, 60:200] = 0.95; c[90:170, 90:170] = 0.1; c[130, 170:200] = 0.2   # hole leaking through a channel
        cases.append(c)
        c = np.zeros((H, W)); c[60:200, 60:200] = 0.95; c[90:170, 90:170] = 0.1; c[89, 170] = 0.95; c[90, 170] = 0.1; c[89, 171] = 0.1
        cases.append(c)                                                 # diagonal-only leak: still a hole for the cross structure
        cases.append(rng.random((H, W)))                                # speckle: thousands of runs (overflow path)
        c = (rng.random((H, W)) < 0.55) * 0.9                           # percolating noise
        cases.append(c)
        c = np.zeros((H, W)); c[0:256, 0:5] = 0.9; c[0:5, :] = 0.9; c[251:, :] = 0.9; c[:, 251:] = 0.9; c[100:130, 100:130] = 0.7
        cases.append(c)                                                 # a frame along the border: the inside is a hole
        c = np.full((H, W), 0.5); c[40:60, 40:60] = np.nextafter(np.float32(0.5), np.float32(1))   # exactly 0.5 rounds to 0
        cases.append(c)
        c = np.zeros((H, W)); c[30:90, 30:90] = -0.8; c[150:170, 150:170] = 1.7   # out-of-range values round to -1 / 2
        cases.append(c)
        c = np.zeros((H, W)); c[::2, 100:140] = 0.9; c[:, 120] = 0.9      # comb: many runs, one component
        cases.append(c)
        for k, c in enumerate(cases[:T]):
            out[0, k] = c.astype(np.float32)
    return out


def closed_contours(n, k, seed, wild=False):
    """n open-ended contours of k points each ((n, k, 2) float64, (x, y)): a jittered arc of an ellipse from 200 to -20
    degrees (like a lung / heart outline whose last point is joined to the first by the closing line)."""
    rng = np.random.default_rng(seed)
    t = np.deg2rad(np.linspace(200.0, -20.0, k))[None]
    a, b = rng.uniform(25, 80, (n, 1)), rng.uniform(30, 100, (n, 1))
    cx, cy = rng.uniform(70, 186, (n, 1)), rng.uniform(70, 186, (n, 1))
    jit = rng.normal(0, 6.0 if wild else 1.2, (n, k, 2))
    c = np.stack([cx + a * np.cos(t), cy - b * np.sin(t)], -1) + jit
    if wild:
        c[::3] += rng.choice([-1, 1], (len(c[::3]), 1, 2)) * rng.uniform(40, 140, (len(c[::3]), 1, 2))   # partly off frame
    return c


def lung_landmarks(n, seed):
    """(n, 120, 2): right lung (44), left lung (50), heart (26) as LungContourToMask splits them."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 120, 2))
    for i in range(n):
        def arc(k, cx, cy, a, b):
            t = np.deg2rad(np.linspace(250.0, -70.0, k))
            return np.stack([cx + a * np.cos(t), cy - b * np.sin(t)], -1) + rng.normal(0, 1.0, (k, 2))
        out[i, :44] = arc(44, 75 + rng.normal(0, 4), 125 + rng.normal(0, 4), 38, 85)
        out[i, 44:94] = arc(50, 180 + rng.normal(0, 4), 130 + rng.normal(0, 4), 36, 80)
        out[i, 94:] = arc(26, 140 + rng.normal(0, 4), 165 + rng.normal(0, 4), 45, 35)
    return out


def dataset_calibration_inputs(seed=12, n=400):
    """(std, error, reject) of a clinical metric over a dataset: what the results processors calibrate."""
    rng = np.random.default_rng(seed)
    std = np.abs(rng.normal(0.05, 0.03, n))
    err = np.abs(rng.normal(0, 1, n)) * std * 1.3
    std[5] = np.nan
    err[17] = np.nan
    reject = rng.random(n) < 0.1
    return std, err, reject


def endo_epi_landmarks(n, seed, wild=False):
    """(n, 42, 2): the 21 LV landmarks (jittered) followed by 21 epicardial ones = the same shape grown about its centroid."""
    rng = np.random.default_rng(seed)
    lv = LV_EXAMPLE_SHAPE[None] + rng.normal(0, 5.0 if wild else 1.5, (n, 21, 2)) + rng.normal(0, 6.0, (n, 1, 2))
    cen = lv.mean(1, keepdims=True)
    epi = cen + (lv - cen) * rng.uniform(1.15, 1.35, (n, 1, 1)) + rng.normal(0, 4.0 if wild else 1.0, (n, 21, 2))
    out = np.concatenate([lv, epi], 1)
    if wild:
        out[::4] += rng.choice([-1, 1], (len(out[::4]), 1, 2)) * rng.uniform(50, 120, (len(out[::4]), 1, 2))     # partly off frame
        out[1, 25] = out[1, 24]                                                                                   # zero chord in the epicardium
    return out

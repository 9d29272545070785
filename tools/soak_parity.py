#!/usr/bin/env python
"""One-off parity soak (beyond what tests/ can afford on the CPU side): many random inputs through the CUDA path and
the oracle (multiprocessing on the GPU box's host cores).  Prints one JSON line per check.

  python tools/soak_parity.py [--contours 20000] [--simpson 1600] [--skew-images 24]
"""
import argparse, json, multiprocessing as mp, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np


def _raster_chunk(c):
    from oracle import port
    return np.stack([np.packbits(port.us_contour_to_mask(x).astype(np.uint8), axis=-1, bitorder="little") for x in c])


def _simpson_chunk(args):
    from oracle import port
    masks, vs = args
    out = []
    for m in masks:
        try:
            d, s = port.compute_diameters(m.astype(bool), vs)
        except Exception:
            d, s = np.full(20, np.nan), np.nan
        out.append(np.concatenate([d, [s]]))
    return np.stack(out)


def _skew_one(args):
    import synth
    from oracle import port
    mu, cov, al, z, seed, i = args
    model = port.ShapeModel.load(os.path.join(R, "contouring_uncertainty_b200", "data", "camus-cont_psm_11_no_std.npy"))
    qf = lambda n, d: synth.exponential_field(seed, (i * len(z) + n) * 14 + d)
    return port.skew_psm_sample(model, mu, cov, al, z, qf, store=np.float32, return_indices=True)


def _segblob_chunk(samples):
    from oracle import port
    return np.packbits(port.seg_blob_masks(samples), axis=-1, bitorder="little")


def _gls_chunk(masks):
    from oracle import port
    out = []
    for m in masks:
        try:
            left, right, _ = port.polar_base_corners(m)
            tie = port.base_corner_tie(m)
            out.append([port.lv_longitudinal_length(m), left[0], left[1], right[0], right[1], float(tie)])
        except (RuntimeError, IndexError):
            out.append([np.nan] * 5 + [0.0])
    return np.array(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contours", type=int, default=20000)
    ap.add_argument("--simpson", type=int, default=1600)
    ap.add_argument("--skew-images", type=int, default=24)
    ap.add_argument("--seg-samples", type=int, default=0)
    ap.add_argument("--gls-masks", type=int, default=0)
    a = ap.parse_args()
    import torch, synth
    from contouring_uncertainty_b200 import engine as E
    dev = torch.device("cuda:0")
    workers = min(os.cpu_count() or 1, 64)
    pool = mp.get_context("spawn").Pool(workers)
    rng = np.random.default_rng(2026)
    b = synth.LV_EXAMPLE_SHAPE

    # 4. segmentation samples: hole fill + largest blob bit-exact (speckled blobs, islands, dips; every 8th pure noise)
    if a.seg_samples:
        n = a.seg_samples
        t0 = time.time()
        seg = synth.segmentation_samples(n // 16, 16, seed=4242, adversarial=False)
        seg[:, ::8] = (rng.random(seg[:, ::8].shape) < rng.uniform(0.05, 0.95, (n // 16, 2, 1, 1))).astype(np.float32) * 0.8
        got = E.seg_largest_blob(torch.from_numpy(seg).to(dev)).cpu().numpy().view(np.uint8).reshape(-1, 256, 32)
        want = np.concatenate(pool.map(_segblob_chunk, np.array_split(seg.reshape(-1, 256, 256), workers * 4)))
        print(json.dumps({"check": "segmentation samples: largest blob of the hole-filled mask bit-exact vs oracle", "samples": int(len(want)),
                          "mismatching_masks": int((got != want).any((1, 2)).sum()), "seconds": round(time.time() - t0, 1)}), flush=True)

    # 5. mask-based LV length (EchoMeasure.gls pieces): same failures, same base corners, lengths to 1e-12
    if a.gls_masks:
        n = a.gls_masks
        t0 = time.time()
        c = b[None] + rng.normal(0, 3.0, (n, 21, 2)) + rng.normal(0, 8.0, (n, 1, 2))
        c[n // 2:] = (c[n // 2:] - c[n // 2:].mean(1, keepdims=True)) * rng.uniform(0.4, 1.15, (n - n // 2, 1, 1)) + 128
        masks, _ = E.rasterize_contours(torch.from_numpy(c.astype(np.float32)).to(dev))
        length, corners, _ = E.mask_lv_length(masks, return_corners=True)
        m_np = E.unpack_masks(masks).cpu().numpy().astype(bool)
        w = np.concatenate(pool.map(_gls_chunk, np.array_split(m_np, workers * 4)))
        g_len, g_cor = length.cpu().numpy(), corners.cpu().numpy().reshape(n, 4)
        tie = w[:, 5] > 0
        nan_mismatch = int((np.isnan(g_len) != np.isnan(w[:, 0])).sum())
        ok = ~np.isnan(w[:, 0]) & ~np.isnan(g_len) & ~tie
        print(json.dumps({"check": "mask LV length: failures / base corners identical, lengths <= 1e-12 rel vs oracle", "masks": n,
                          "reference_raises": int(np.isnan(w[:, 0]).sum()), "nan_mismatches": nan_mismatch, "exact_ties_skipped": int(tie.sum()),
                          "corner_mismatches": int((g_cor[ok] != w[ok, 1:5]).any(1).sum()),
                          "max_rel_length_diff": float((np.abs(g_len[ok] - w[ok, 0]) / w[ok, 0]).max()) if ok.any() else 0.0,
                          "seconds": round(time.time() - t0, 1)}), flush=True)

    # 1. rasteriser: random + wild + near-border + half-integer contours, masks bit-exact
    n = a.contours
    if n == 0:
        pool.close()
        return
    kind = rng.integers(0, 5, n)
    c = b[None] + rng.normal(0, 3.0, (n, 21, 2)) + rng.normal(0, 6.0, (n, 1, 2))
    c[kind == 1] = b[None] + rng.normal(0, 11.0, ((kind == 1).sum(), 21, 2))                       # self-intersecting
    c[kind == 2] += rng.choice([-1, 1], ((kind == 2).sum(), 1, 2)) * rng.uniform(60, 130, ((kind == 2).sum(), 1, 2))  # off-frame
    c[kind == 3] = np.round(c[kind == 3]) + 0.5                                                     # rounding ties
    c = c.astype(np.float32)
    t0 = time.time()
    masks, area = E.rasterize_contours(torch.from_numpy(c).to(dev), return_masks=True)
    got = masks.cpu().numpy().view(np.uint8).reshape(n, 256, 32)
    want = np.concatenate(pool.map(_raster_chunk, np.array_split(c, workers * 4)))
    bad = int((got != want).any((1, 2)).sum())
    print(json.dumps({"check": "rasterised masks bit-exact vs oracle", "contours": n, "mismatching_masks": bad,
                      "kinds": "random / self-intersecting / off-frame (wrap, clip, IndexError) / .5 ties", "seconds": round(time.time() - t0, 1)}), flush=True)

    # 2. Simpson diameters on sampled masks, isotropic and anisotropic spacing, fp64-identical
    ns = a.simpson
    sel = np.flatnonzero((kind == 0) | (kind == 1))[:ns]
    m_np = np.unpackbits(want[sel], axis=-1, bitorder="little")
    for vs in ((0.5, 0.5), (0.5, 0.37)):
        t0 = time.time()
        d, s = E.simpson_diameters(masks[torch.from_numpy(sel).to(dev)].contiguous(), vs)
        gd = np.concatenate([d.cpu().numpy(), s.cpu().numpy()[:, None]], 1)
        wd = np.concatenate(pool.map(_simpson_chunk, [(x, vs) for x in np.array_split(m_np, workers * 2)]))
        bad = int((~((gd == wd) | (np.isnan(gd) & np.isnan(wd)))).any(1).sum())
        print(json.dumps({"check": "Simpson diameters / step fp64-identical vs oracle", "masks": len(sel), "voxelspacing": vs,
                          "mismatching_masks": bad, "seconds": round(time.time() - t0, 1)}), flush=True)

    # 3. skew sampler with replayed draws: grid indices bit-exact
    ni, T = a.skew_images, 2
    mus, covs = synth.point_distributions(ni, seed=555)
    al = synth.skew_alphas(ni, seed=556, scale=2.5)
    z = rng.standard_normal((ni, T, 3, 3)).astype(np.float32)
    t0 = time.time()
    ex = np.stack([np.stack([np.stack([synth.exponential_field(557, (i * T + s) * 14 + d) for d in range(14)]) for s in range(T)]) for i in range(ni)])
    model = E.ShapeModelConstants(device=dev)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    out, idx = E.skew_sample(model, t(mus), t(covs), t(al), T, z=t(z), expo=t(ex), return_indices=True)
    res = pool.map(_skew_one, [(mus[i], covs[i], al[i], z[i], 557, i) for i in range(ni)])
    widx = np.stack([r[1] for r in res]); wout = np.stack([r[0] for r in res])
    print(json.dumps({"check": "skew sampler grid indices bit-exact vs oracle (replayed draws)", "draws": int(widx.size),
                      "mismatching_draws": int((idx.cpu().numpy() != widx).sum()),
                      "max_abs_diff_px": float(np.abs(out.cpu().numpy() - wout).max()), "seconds": round(time.time() - t0, 1)}), flush=True)
    pool.close()


if __name__ == "__main__":
    main()

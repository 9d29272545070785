"""The pruning rules of the base-chord search of csrc/simpson.cu (`chord_slim`, `chord_dead`, the per-block column / row test and
the per-point pre-pass), emulated step by step in numpy and compared with the exhaustive loop of the reference
(`_compute_diameters`, vital/vital/metrics/evaluate/clinical/heart_us.py:127-146, oracle/port.py::compute_diameters): the rules
may only ever drop pairs that cannot be selected, so the answer (length, i, j) must be identical on any contour.

`python tests/test_chord_search_cpu.py [count] [seed]` soaks more masks and prints the work model the stride of the coarse pass
was chosen with (warp-level iterations of each phase per mask)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import synth
from oracle import port

COARSE = 16


def contour_points(mask):
    cs = port.find_contours(mask, 0.5)
    return None if not cs else np.round(cs[0] * 2).astype(np.int64)       # (n, 2) row, col in half-pixel units


def exhaustive(P):
    """heart_us.py:127-146 in integer arithmetic: first longest valid chord in (i, j) loop order."""
    n = len(P)
    best = (0, None, None)
    for i in range(2, n):
        d = P[:i] - P[i]
        for j in np.nonzero((d[:, 1] > 0) & (np.abs(d[:, 0]) <= d[:, 1]))[0]:
            a = P[j] - P[i]
            l2 = int(a @ a)
            if l2 > best[0] and _valid(P, i, j, l2):
                best = (l2, i, int(j))
    return best


def _valid(P, i, j, l2):
    a = P[j] - P[i]
    b = P[i] - P[j + 1:i]
    cr = a[0] * b[:, 1] - a[1] * b[:, 0]
    return bool(np.all(cr * cr <= 256 * l2))


def chord_slim(w):
    if w <= 4:
        return 0
    s = int(0.5 * np.sqrt(float(w)))
    while 4 * s * s >= w:
        s -= 1
    while 4 * (s + 1) * (s + 1) < w:
        s += 1
    return s


def _steep(a, b):
    return abs(int(a[0]) - int(b[0])) - abs(int(a[1]) - int(b[1])) >= 23


def chord_dead(P, k, n, slim, direction):
    if slim <= 0:
        return False
    if (k if direction < 0 else n - 1 - k) <= slim:
        return True
    return any(_steep(P[k + direction * s], P[k]) for s in (slim, (3 * slim) >> 2, slim >> 1))


def _better(a, b):
    return a[0] > b[0] or (a[0] == b[0] and (a[1], a[2]) < (b[1], b[2]))


def kernel_search(P, coarse=COARSE, work=None):
    """The three steps of simpson_diameters_kernel's chord search.  `work` (dict) counts warp-level iterations."""
    n = len(P)
    best = (0, 1 << 30, 1 << 30)
    work = {} if work is None else work
    for k in ("coarse_i", "coarse_iters", "validations", "full_i", "live_blocks"):
        work.setdefault(k, 0)
    if n < 3:
        return best

    def near(i, j, k, l2):
        a, b = P[j] - P[i], P[i] - P[k]
        cr = int(a[0]) * int(b[1]) - int(a[1]) * int(b[0])
        return cr * cr <= 256 * l2

    def try_pairs(i, js):
        """One warp-wide call: the candidates js (increasing) of the same i; per-candidate filter (angle, selectable, middle and
        quarter points of the span), then validation longest first, first j among equals; the first valid one ends the call."""
        nonlocal best
        cands = []
        for j in js:
            a = P[j] - P[i]
            if not (a[1] > 0 and abs(a[0]) <= a[1]):
                continue
            c = (int(a @ a), int(i), int(j))
            if not _better(c, best):
                continue
            q = (i - j) >> 2
            if near(i, j, (i + j) >> 1, c[0]) and (q == 0 or (near(i, j, j + q, c[0]) and near(i, j, i - q, c[0]))):
                cands.append(c)
        for c in sorted(cands, key=lambda c: (-c[0], c[2])):
            work["validations"] += 1
            if _valid(P, c[1], c[2], c[0]):
                best = c
                return

    for i in range((n - 1) // coarse * coarse, coarse - 1, -coarse):
        work["coarse_i"] += 1
        if chord_dead(P, i, n, chord_slim(best[0]), -1):
            continue
        for j0 in range(0, i, 32 * coarse):
            work["coarse_iters"] += 1
            try_pairs(i, [j for j in range(j0, min(i, j0 + 32 * coarse), coarse)])
    if best[0] > 0:                                        # refinement around the coarse chord
        ci, cj = best[1], best[2]
        for i in range(min(n - 1, ci + coarse - 1), max(2, ci - (coarse - 1)) - 1, -1):
            try_pairs(i, [j for j in range(cj - 16, cj + 16) if 0 <= j < i])
    slim = chord_slim(best[0])
    dmin0 = max(1, int(np.ceil(np.sqrt(0.5 * best[0]) - 1e-9)))
    live_i = np.array([k >= 2 and not chord_dead(P, k, n, slim, -1) for k in range(n)])
    live_j = np.array([not chord_dead(P, k, n, slim, +1) for k in range(n)])
    nw = (n + 31) >> 5
    bmax, r0, r1 = np.zeros(nw, int), np.full(nw, 0xFFFF), np.zeros(nw, int)
    for t in range(nw):
        k = np.arange(32 * t, min(32 * t + 32, n))
        k = k[live_j[k]]
        if len(k):
            bmax[t], r0[t], r1[t] = P[k, 1].max(), P[k, 0].min(), P[k, 0].max()

    def block_ok(b, ir, ic, dmin):
        return bmax[b] - ic >= dmin and max(0, r0[b] - ir, ir - r1[b]) <= bmax[b] - ic

    for i in np.nonzero(live_i)[0]:
        ir, ic = P[i]
        if not any(block_ok(b, ir, ic, dmin0) for b in range((i + 31) >> 5)):
            continue
        work["full_i"] += 1
        dmin = max(1, int(np.ceil(np.sqrt(0.5 * best[0]) - 1e-9)))
        blocks = [b for b in range((i + 31) >> 5) if block_ok(b, ir, ic, dmin)]     # one ballot per 32 blocks, then block by block
        for b in blocks:
            work["live_blocks"] += 1
            try_pairs(i, [j for j in range(32 * b, min(32 * b + 32, i)) if live_j[j]])
    return best


def masks(rng, count):
    """LV-like contours (scaled, rotated, noisy, shifted off-frame), rectangles, speckled ellipses, smooth-noise blobs."""
    from scipy.ndimage import gaussian_filter

    b = synth.LV_EXAMPLE_SHAPE
    for q in range(count):
        kind = q % 6
        if kind <= 2:
            s = rng.uniform(0.5, 1.3)
            th = rng.uniform(-0.6, 0.6) if kind == 2 else rng.uniform(-0.15, 0.15)
            rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            c = b.mean(0) + ((b - b.mean(0)) * s) @ rot.T + rng.normal(0, rng.uniform(0.5, 6), (21, 2)) + rng.normal(0, 15, 2)
            yield port.convert_to_mask(c.astype(np.float32)[None])[0]
        elif kind == 3:
            m = np.zeros((256, 256), np.uint8)
            r, c = rng.integers(0, 200, 2)
            h, w = rng.integers(2, 120, 2)
            m[r:min(r + h, 256), c:min(c + w, 256)] = 1
            yield m
        elif kind == 4:
            yy, xx = np.mgrid[:256, :256]
            cy, cx = rng.uniform(20, 240, 2)
            ay, ax = rng.uniform(5, 120, 2)
            m = (((yy - cy) / ay) ** 2 + ((xx - cx) / ax) ** 2 <= 1).astype(np.uint8)
            yield m & (rng.random((256, 256)) > 0.02).astype(np.uint8)
        else:
            f = gaussian_filter(rng.normal(size=(256, 256)), rng.uniform(3, 12))
            yield (f > np.quantile(f, rng.uniform(0.5, 0.9))).astype(np.uint8)


def _check(count, seed, coarse=COARSE, work=None):
    rng = np.random.default_rng(seed)
    done = 0
    for m in masks(rng, count):
        P = contour_points(m)
        if P is None or len(P) > 900:
            continue
        if len(P) > 1:      # the fact chord_slim rests on: consecutive marching-squares points are <= 2 half-pixel units apart
            assert (np.diff(P, axis=0) ** 2).sum(1).max() <= 4
        want, got = exhaustive(P), kernel_search(P, coarse, work)
        assert (want[1] is None and got[0] == 0) or want == got, (len(P), want, got)
        done += 1
    return done


def test_chord_search_rules_are_exact():
    assert _check(36, seed=2024) >= 30


def test_chord_slim_is_the_largest_s():
    for w in list(range(0, 300)) + [4 * s * s + d for s in (10, 57, 300, 1000) for d in (-1, 0, 1)]:
        s = chord_slim(w)
        assert (s == 0 and w <= 4) or (4 * s * s < w <= 4 * (s + 1) * (s + 1))


if __name__ == "__main__":
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for coarse in (8, 16, 32):
        work = {}
        done = _check(count, seed, coarse, work)
        print(f"coarse stride {coarse}: {done} masks identical to the exhaustive loop; per mask:",
              {k: round(v / done, 1) for k, v in work.items()})

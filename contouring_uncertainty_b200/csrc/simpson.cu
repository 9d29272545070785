// Simpson's biplane volumes from sampled masks: marching-squares contour -> base chord -> apex ->
// 20 disc diameters per mask, then EDV / ESV / EF per Monte-Carlo sample.
//
// Replaces (reference paths):
//   skimage.measure.find_contours(mask, 0.5)[0]       (third party; restated, see oracle/port.py::find_contours)
//   _compute_diameters / _find_distance_to_edge / _distance_line_to_points / _get_angle_of_lines_to_point
//                                                     vital/vital/metrics/evaluate/clinical/heart_us.py:74-167
//   compute_left_ventricle_volumes / _compute_left_ventricle_volume_by_instant        heart_us.py:13-71
//   the Monte-Carlo loop of Volume.compute            contour_uncertainty/results/clinical/patient.py:166-190
//
// find_contours()[0] of a binary mask is the boundary component that owns the raster-first
// marching-squares segment; its point list follows the segment direction and starts at the to-point of
// the component's raster-LAST segment (that segment closes the loop in _assemble_contours) or, for a
// component cut by the frame (no padding), at the chain's open start.  All points lie on half-pixel
// positions, so they are stored exactly as (2 row, 2 col) uint16 pairs and every comparison of the
// chord search (angle <= 45 deg, distance-to-line <= 8 px, longest chord, farthest apex) is done in
// exact integer arithmetic; the ray marching that measures the discs repeats the reference's fp64
// operations one by one (no FMA contraction).
#include "contour_trace.cuh"

namespace cub200 {

constexpr int SP_WARPS = 4;
constexpr int CHORD_COARSE = 16;   // stride of the coarse pass of the chord search (8: 20 % more work, 32: 30 % more; tests/test_chord_search_cpu.py)

// ---------------------------------------------------------------- S2: base chord, apex, discs; one warp per mask
struct Chord {
    int len2;   // squared length in half-pixel units
    int i, j;
};
__device__ __forceinline__ bool chord_better(const Chord& a, const Chord& b) {   // is a better than b
    if (a.len2 != b.len2) return a.len2 > b.len2;
    if (a.i != b.i) return a.i < b.i;                 // first in (i, j) loop order wins ties (strict '<' in the reference)
    return a.j < b.j;
}


// Exact pruning of the chord search.  Consecutive contour points are at most 2 half-pixel units apart (a marching-squares
// segment joins two edge midpoints of one cell), so |P_i - P_j| <= 2 (i - j): with `wbest` the squared length of a valid
// chord already held, pairs less than slim + 1 points apart, 4 slim^2 < wbest, are too short to be selected.
__device__ __forceinline__ int chord_slim(int wbest) {          // largest s with 4 s^2 < wbest
    if (wbest <= 4) return 0;
    int s = (int)(0.5 * sqrt((double)wbest));
    while (4 * s * s >= wbest) --s;
    while (4 * (s + 1) * (s + 1) < wbest) ++s;
    return s;
}
// Can point k be ruled out as the later end i (DIR = -1: partners j < k) or the earlier end j (DIR = +1: partners i > k) of
// the answer?  Take a witness w = P[k + DIR s] - P[k], 1 <= s <= slim.  Partners up to the witness are too close (above).
// For the others the witness is an intermediate point of the chord, whose direction a is admissible (|a_row| <= a_col):
// |cross(a, w)| >= a_col (|w_row| - |w_col|) and |a|^2 <= 2 a_col^2, so (|w_row| - |w_col|)^2 > 512 puts the witness more
// than 8 px (16 units) from the chord's line whatever the partner: 23^2 = 529.  No partner at all beyond slim: ruled out too.
__device__ __forceinline__ bool chord_steep(uint32_t a, uint32_t b) {
    const int dr = (int)(a & 0xffff) - (int)(b & 0xffff), dc = (int)(a >> 16) - (int)(b >> 16);
    return abs(dr) - abs(dc) >= 23;
}
template <int DIR>
__device__ __forceinline__ bool chord_dead(const uint32_t* P, int k, int n, int slim) {
    if (slim <= 0) return false;
    if ((DIR < 0 ? k : n - 1 - k) <= slim) return true;
    const uint32_t p = P[k];
    return chord_steep(P[k + DIR * slim], p) || chord_steep(P[k + DIR * ((3 * slim) >> 2)], p) || chord_steep(P[k + DIR * (slim >> 1)], p);
}

__global__ void __launch_bounds__(SP_WARPS * 32)
simpson_diameters_kernel(const uint32_t* __restrict__ masks, const uint32_t* __restrict__ pts_g, const int32_t* __restrict__ npts,
                         int64_t n_masks, const uint16_t* __restrict__ row_index, const uint16_t* __restrict__ col_index, int vh,
                         int vw, double spacing, double* __restrict__ diam_out, double* __restrict__ step_out) {
    __shared__ uint32_t s_pts[SP_WARPS][SP_NMAX];
    __shared__ uint32_t s_bmax[SP_WARPS][SP_NMAX / 32];
    __shared__ uint32_t s_brow[SP_WARPS][SP_NMAX / 32];
    __shared__ uint32_t s_live[SP_WARPS][2 * (SP_NMAX / 32)];
    extern __shared__ uint16_t s_idx[];
    if (row_index) {
        for (int i = threadIdx.x; i < vh; i += blockDim.x) s_idx[i] = row_index[i];
        for (int i = threadIdx.x; i < vw; i += blockDim.x) s_idx[vh + i] = col_index[i];
        __syncthreads();
    }
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    const int64_t g = blockIdx.x * (int64_t)SP_WARPS + wid;
    if (g >= n_masks) return;
    const int n = npts[g];
    // the rays at the end read the mask itself: ask for its 64 lines now
    asm volatile("prefetch.global.L2 [%0];" ::"l"(masks + g * CUB200_MASK_WORDS + lane * 64));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(masks + g * CUB200_MASK_WORDS + lane * 64 + 32));
    double* dout = diam_out + g * 20;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (n < 3) {                                       // no contour / no chord: the reference raises
        if (lane < 20) dout[lane] = qnan;
        if (lane == 0) step_out[g] = qnan;
        return;
    }
    uint32_t* P = s_pts[wid];
    for (int k = lane; k < n; k += 32) P[k] = pts_g[g * SP_NMAX + k];
    __syncwarp();
    // ---- longest admissible chord (heart_us.py:127-146).  The reference keeps the first (i, j) in loop order among the
    //      longest valid chords, so pairs may be visited in any order as long as ties are broken by (i, j).  Three steps:
    //      (1) a coarse pass over every 16th point finds a long valid chord; (2) with its length as a bound most points are
    //      ruled out as an end of the answer by ONE intermediate point each (`chord_dead`, exact); (3) the full pass runs over
    //      the surviving i only and skips, 32 candidates at a time, the blocks of surviving j whose columns / rows cannot
    //      give a chord that long (|d_row| <= d_col  =>  len^2 <= 2 d_col^2).
    Chord best{0, 0x7fffffff, 0x7fffffff};               // warp-uniform
    // One candidate j per lane for the same i (on = the lane has one), j increasing with the lane.  Each lane applies the tests
    // that need no loop (admissible angle, can still be selected, the middle of the span -- where most far-away chords
    // fail); the survivors are then validated longest first (first j among equals) by the WHOLE warp, 32 intermediate points
    // at a time: as soon as one is valid no other candidate of this call can be selected.
    auto try_pairs = [&](int i, int ir, int ic, int j, bool on) {
        uint32_t key = 0u;                               // len2 << 5 | (31 - lane); len2 < 2^24 (coordinates < 2^11)
        if (on) {
            const int ar = (int)(P[j] & 0xffff) - ir, ac = (int)(P[j] >> 16) - ic;     // contour[j] - contour[i]
            // |degrees(arctan2(d_row, d_col))| <= 45  <=>  d_col > 0 and |d_row| <= d_col  (zero vector: NaN distances)
            const Chord cand{ar * ar + ac * ac, i, j};
            if (ac > 0 && abs(ar) <= ac && chord_better(cand, best)) {
                const int64_t lim = 256ll * cand.len2;
                auto near = [&](int k) {                 // is intermediate point k within 8 px of the chord's line?
                    const int br = ir - (int)(P[k] & 0xffff), bc = ic - (int)(P[k] >> 16);
                    const int64_t cr = (int64_t)ar * bc - (int64_t)ac * br;
                    return cr * cr <= lim;
                };
                const int q = (i - j) >> 2;              // the middle, then the quarter points (q >= 1: strictly inside)
                if (near((i + j) >> 1) && (q == 0 || (near(j + q) && near(i - q))))
                    key = ((uint32_t)cand.len2 << 5) | (uint32_t)(31 - lane);
            }
        }
        for (;;) {
            const uint32_t top = __reduce_max_sync(FULL, key);
            if (top == 0u) return;
            const int wl = 31 - (int)(top & 31u), wlen2 = (int)(top >> 5);
            const int wj = __shfl_sync(FULL, j, wl);
            const int ar = (int)(P[wj] & 0xffff) - ir, ac = (int)(P[wj] >> 16) - ic;
            const int64_t lim = 256ll * wlen2;           // (8 px)^2 * len2 in half-pixel units: cross^2 <= 256 len2
            bool ok = true;
            for (int k0 = wj + 1; ok && k0 < i; k0 += 32) {
                const int k = k0 + lane;
                bool good = true;
                if (k < i) {
                    const int br = ir - (int)(P[k] & 0xffff), bc = ic - (int)(P[k] >> 16);   // contour[i] - contour[k]
                    const int64_t cr = (int64_t)ar * bc - (int64_t)ac * br;
                    good = cr * cr <= lim;
                }
                ok = __all_sync(FULL, good);
            }
            if (ok) { best = Chord{wlen2, i, wj}; return; }
            if (lane == wl) key = 0u;
        }
    };
    // (1) coarse pass: i and j on every 16th point (validity is always checked on ALL intermediate points), i DESCENDING: a
    // straight stretch of the contour then meets its far end first, i.e. its longest chord, and the bound it sets prunes
    // the O(L^2) shorter chords of the same stretch that an ascending scan would have validated one after the other.  As
    // soon as a chord is held, the i that `chord_dead` rules out are skipped here as well.
    {
        int slim = 0, slim_for = 0;
        for (int i = (n - 1) & ~(CHORD_COARSE - 1); i >= CHORD_COARSE; i -= CHORD_COARSE) {
            if (best.len2 != slim_for) { slim = chord_slim(best.len2); slim_for = best.len2; }
            if (chord_dead<-1>(P, i, n, slim)) continue;                                // warp-uniform
            const int ir = P[i] & 0xffff, ic = P[i] >> 16;
            for (int j0 = 0; j0 < i; j0 += 32 * CHORD_COARSE) {
                const int j = j0 + CHORD_COARSE * lane;
                try_pairs(i, ir, ic, j < i ? j : 0, j < i);
            }
        }
    }
    // ... refined around the chord it found: every i within a coarse step of its i against the 32 j around its j, so that
    // the bound the next steps prune with is (nearly always) the final one
    if (best.len2 > 0) {
        const int ci = best.i, cj = best.j;
        for (int i = min(n - 1, ci + CHORD_COARSE - 1); i >= max(2, ci - (CHORD_COARSE - 1)); --i) {
            const int j = cj - 16 + lane;
            const bool on = j >= 0 && j < i;
            try_pairs(i, P[i] & 0xffff, P[i] >> 16, on ? j : 0, on);
        }
    }
    // (2) the points that cannot be the i / the j of the answer, and the column / row ranges of the others per block of 32
    uint32_t* bmax = s_bmax[wid];                        // max column (half-pixel units) of the block's surviving j
    uint32_t* brow = s_brow[wid];                        // their row range: min | max << 16
    uint32_t* live_i = s_live[wid];                      // bit k of word t: point 32 t + k may be the i of the answer
    uint32_t* live_j = s_live[wid] + SP_NMAX / 32;
    const int nwords = (n + 31) >> 5;
    {
        const int slim = chord_slim(best.len2);
        for (int t = 0; t < nwords; ++t) {
            const int k = 32 * t + lane;
            const bool inb = k < n;
            const uint32_t p = inb ? P[k] : 0u;
            const bool li = inb && k >= 2 && !chord_dead<-1>(P, k, n, slim);
            const bool lj = inb && !chord_dead<+1>(P, k, n, slim);
            const uint32_t wi = __ballot_sync(FULL, li), wj = __ballot_sync(FULL, lj);
            const uint32_t cm = __reduce_max_sync(FULL, lj ? (p >> 16) : 0u);
            const uint32_t r0 = __reduce_min_sync(FULL, lj ? (p & 0xffffu) : 0xffffu);
            const uint32_t r1 = __reduce_max_sync(FULL, lj ? (p & 0xffffu) : 0u);
            if (lane == 0) { live_i[t] = wi; live_j[t] = wj; bmax[t] = cm; brow[t] = r0 | (r1 << 16); }
        }
    }
    __syncwarp();
    // ... and the i none of whose blocks passes the block test of the full pass under the bound held now (the bound only
    // grows): lane = point, 32 i at a time, instead of one warp-wide iteration of the full pass each
    {
        const int dmin0 = max(1, (int)ceil(sqrt(0.5 * (double)best.len2) - 1e-9));
        for (int t = 0; t < nwords; ++t) {
            const uint32_t m = live_i[t];
            if (m == 0u) continue;                                                      // warp-uniform
            const int i = 32 * t + lane;
            bool keep = false;
            if ((m >> lane) & 1u) {
                const int ir = P[i] & 0xffff, ic = P[i] >> 16;
                const int nblk = (i + 31) >> 5;
                for (int b = 0; b < nblk && !keep; ++b) {
                    const int dcmax = (int)bmax[b] - ic;
                    const int rlo = (int)(brow[b] & 0xffff), rhi = (int)(brow[b] >> 16);
                    keep = dcmax >= dmin0 && max(0, max(rlo - ir, ir - rhi)) <= dcmax;
                }
            }
            const uint32_t m2 = __ballot_sync(FULL, keep);
            if (lane == 0) live_i[t] = m2;
        }
    }
    __syncwarp();
    // (3) full pass over the surviving i
    int dmin_for = -1, dmin = 1;
    for (int t = 0; t < nwords; ++t) {
        uint32_t todo = live_i[t];
        while (todo) {
            const int i = 32 * t + __ffs(todo) - 1;
            todo &= todo - 1;
            const int ir = P[i] & 0xffff, ic = P[i] >> 16;
            // a chord shorter than the best one held can never be selected (equal ones still can: ties go to the first
            // (i, j) in loop order)
            if (best.len2 != dmin_for) {                                                // d_col needed for len^2 >= best.len2
                dmin = max(1, (int)ceil(sqrt(0.5 * (double)best.len2) - 1e-9));
                dmin_for = best.len2;
            }
            const int nblk = (i + 31) >> 5;
            for (int jb0 = 0; jb0 < nblk; jb0 += 32) {
                const int jb = jb0 + lane, jc = jb < nblk ? jb : 0;
                // a block can hold a candidate only if its columns reach d_col >= dmin AND some row is within that d_col
                // of row i (|d_row| <= d_col); a block without a surviving j has column 0 and fails the first test
                const int dcmax = (int)bmax[jc] - ic;
                const int rlo = (int)(brow[jc] & 0xffff), rhi = (int)(brow[jc] >> 16);
                const int rdist = max(0, max(rlo - ir, ir - rhi));
                unsigned live = __ballot_sync(FULL, jb < nblk && dcmax >= dmin && rdist <= dcmax);
                while (live) {
                    const int b = __ffs(live) - 1;
                    live &= live - 1;
                    const int j = 32 * (jb0 + b) + lane;
                    try_pairs(i, ir, ic, j < i ? j : 0, j < i && ((live_j[jb0 + b] >> lane) & 1u));
                }
            }
        }
    }
    if (best.len2 <= 0) {                               // best_j never assigned -> UnboundLocalError
        if (lane < 20) dout[lane] = qnan;
        if (lane == 0) step_out[g] = qnan;
        return;
    }
    // mid_point = int(best_j + round((best_i - best_j) / 2))   (Python round: half to even)
    const int diff = best.i - best.j;
    int half = diff >> 1;
    if ((diff & 1) && (half & 1)) half += 1;
    const int mid = best.j + half;
    const int mr = P[mid] & 0xffff, mc = P[mid] >> 16;
    // apex = first farthest contour point from the mid point
    int a_len2 = 0, a_idx = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
        const int dr = (int)(P[i] & 0xffff) - mr, dc = (int)(P[i] >> 16) - mc;
        const int l2 = dr * dr + dc * dc;
        if (l2 > a_len2) { a_len2 = l2; a_idx = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int l2 = __shfl_xor_sync(FULL, a_len2, o), ix = __shfl_xor_sync(FULL, a_idx, o);
        if (l2 > a_len2 || (l2 == a_len2 && ix < a_idx)) { a_len2 = l2; a_idx = ix; }
    }
    if (a_len2 == 0) a_idx = 0;                          // all points coincide: apex stays 0, direction 0 -> NaN normal
    const double dir_r = 0.5 * (double)((int)(P[a_idx] & 0xffff) - mr), dir_c = 0.5 * (double)((int)(P[a_idx] >> 16) - mc);
    const double nrm = sqrt(__dadd_rn(__dmul_rn(dir_c, dir_c), __dmul_rn(dir_r, dir_r)));
    const double nr = -dir_c / nrm, nc = dir_r / nrm;    // normal_direction = (-direction[1], direction[0]) / |.|
    const double mid_r = 0.5 * (double)mr, mid_c = 0.5 * (double)mc;
    // the contour points are no longer needed: their shared-memory buffer now holds the mask, so that the ~50 dependent
    // steps of every ray read shared memory instead of global memory
    static_assert(SP_NMAX >= CUB200_MASK_WORDS, "the point buffer must be able to hold a mask");
    __syncwarp();
    {
        const uint4* src = reinterpret_cast<const uint4*>(masks + g * CUB200_MASK_WORDS);
        uint4* dst = reinterpret_cast<uint4*>(P);
        for (int k = lane; k < CUB200_MASK_WORDS / 4; k += 32) dst[k] = src[k];
    }
    __syncwarp();
    const MaskView mv{P, row_index ? s_idx : nullptr, row_index ? s_idx + vh : nullptr, vh, vw};
    // 20 discs x 2 directions (heart_us.py:158-164): 40 rays, each marched from 8 px in steps of 0.5 px until it leaves the
    // mask (or the frame; NaN too: the reference raises there).  The distances are 8 + m / 2, exact however they are formed.
    auto inside = [&](double pr, double pc, double vr, double vc, double dist) -> bool {
        const double y = rint(__dadd_rn(pr, __dmul_rn(dist, vr))), x = rint(__dadd_rn(pc, __dmul_rn(dist, vc)));
        if (!(y >= 0.0 && y < (double)vh && x >= 0.0 && x < (double)vw)) return false;
        return mv.bit((int)y, (int)x);
    };
    double d_first = 0.0, d_neg2 = 0.0;
    {   // rays 0..31, one per lane: lane k < 20 = disc k along +normal, lane 20 + k = disc k < 12 along -normal
        const int k = lane % 20;
        const double sgn = lane < 20 ? 1.0 : -1.0;
        const double frac = __dmul_rn((double)k, 0.05);                         // np.linspace(0, 1, 20, endpoint=False)
        const double pr = __dadd_rn(mid_r, __dmul_rn(dir_r, frac)), pc = __dadd_rn(mid_c, __dmul_rn(dir_c, frac));
        const double vr = sgn * nr, vc = sgn * nc;
        double dist = 8.0;
        while (inside(pr, pc, vr, vc, dist)) dist += 0.5;
        d_first = dist;
    }
    {   // rays 32..39 = discs 12..19 along -normal: four lanes per ray test four consecutive distances per iteration
        const int q = lane >> 2, ph = lane & 3;
        const double frac = __dmul_rn((double)(12 + q), 0.05);
        const double pr = __dadd_rn(mid_r, __dmul_rn(dir_r, frac)), pc = __dadd_rn(mid_c, __dmul_rn(dir_c, frac));
        const double vr = -nr, vc = -nc;
        double dist = 8.0 + 0.5 * (double)ph;
        bool found = false;
        for (;;) {
            const bool out = found || !inside(pr, pc, vr, vc, dist);
            const unsigned grp = (__ballot_sync(FULL, out) >> (4 * q)) & 0xfu;
            if (!found && grp) {                                                 // the first distance of this ray that is out
                found = true;
                d_neg2 = (dist - 0.5 * (double)ph) + 0.5 * (double)(__ffs(grp) - 1);
            }
            if (__all_sync(FULL, found)) break;
            dist += 2.0;
        }
    }
    // lane k < 20 holds d_pos of disc k; d_neg of disc k sits in lane 20 + k (k < 12) or in lanes 4 (k - 12) .. + 3
    const double neg_a = __shfl_sync(FULL, d_first, (lane + 20) & 31);
    const double neg_b = __shfl_sync(FULL, d_neg2, (4 * (lane - 12)) & 31);
    const double d_pos = d_first;
    if (lane < 20) dout[lane] = __dmul_rn(__dadd_rn(d_pos, lane < 12 ? neg_a : neg_b), spacing);
    if (lane == 0) step_out[g] = __dmul_rn(0.5 * sqrt((double)a_len2), spacing) / 20.0;
}

// ---------------------------------------------------------------- S3: volumes per Monte-Carlo sample
// np.sum of 20 doubles: pairwise_sum with 8 accumulators over the first 16, then the last 4 in order
__device__ __forceinline__ double np_sum20(const double* v) {
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(v[k], v[8 + k]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
#pragma unroll
    for (int k = 16; k < 20; ++k) res = __dadd_rn(res, v[k]);
    return res;
}

__device__ __forceinline__ double simpson_volume(const double* a2c, const double* a4c, double step) {
    double prod[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) prod[k] = __dmul_rn(a2c[k] / 1000.0, a4c[k] / 1000.0);
    const double v = __dmul_rn(__dmul_rn(np_sum20(prod), step / 1000.0), 3.141592653589793) / 4.0;
    return rint(__dmul_rn(v, 1e6));                               // Python round(): half to even
}

__global__ void simpson_volumes_kernel(const double* __restrict__ diam, const double* __restrict__ step, int64_t n_patients,
                                       int64_t n_samples, double* __restrict__ edv, double* __restrict__ esv,
                                       double* __restrict__ ef) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_patients * n_samples) return;
    const int64_t p = g / n_samples, t = g % n_samples;
    // masks of a patient: [A2C-ED, A2C-ES, A4C-ED, A4C-ES][sample]
    const double* d[4];
    double st[4];
    bool ok = true;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int64_t mi = (p * 4 + v) * n_samples + t;
        d[v] = diam + mi * 20;
        st[v] = step[mi];
        ok = ok && (st[v] == st[v]);
    }
    double v_ed = 0.0, v_es = 0.0, e = 0.0;
    if (ok) {                                                     // an exception anywhere leaves the sample at zero (patient.py:186)
        const double s = fmax(fmax(st[0], st[1]), fmax(st[2], st[3]));
        v_ed = simpson_volume(d[0], d[2], s);
        v_es = simpson_volume(d[1], d[3], s);
        if (!(v_ed == v_ed) || !(v_es == v_es)) { v_ed = 0.0; v_es = 0.0; }     // round(nan) raises
        else if (v_ed != 0.0) e = (v_ed - v_es) / v_ed;           // ZeroDivisionError keeps EF at zero
    }
    edv[g] = v_ed;
    esv[g] = v_es;
    ef[g] = (e < 0.0 || e > 1.0) ? __longlong_as_double(0x7ff8000000000000LL) : e;   // patient.py:189-190
}

}  // namespace cub200

extern "C" int64_t cub200_simpson_workspace_bytes(int64_t n_masks) {
    if (n_masks < 0) return CUB200_EINVAL;
    return n_masks * (int64_t)(cub200::SP_NMAX * sizeof(uint32_t) + sizeof(int32_t));
}

static int simpson_launch(const uint32_t* masks, int64_t n_masks, const uint16_t* row_index, const uint16_t* col_index, int vh, int vw,
                          double spacing, double* diam_out, double* step_out, int32_t* npts_out, void* workspace, cudaStream_t st) {
    using namespace cub200;
    if ((reinterpret_cast<uintptr_t>(masks) | reinterpret_cast<uintptr_t>(workspace)) & 15u) return CUB200_EINVAL;   // 16-byte loads / stores
    uint32_t* pts = reinterpret_cast<uint32_t*>(workspace);
    int32_t* npts = reinterpret_cast<int32_t*>(pts + n_masks * SP_NMAX);
    const size_t idx_smem = row_index ? sizeof(uint16_t) * (size_t)(vh + vw) : 0;
    contour_trace_kernel<<<(unsigned)((n_masks + 127) / 128), 128, idx_smem, st>>>(masks, n_masks, row_index, col_index, vh, vw, pts, npts);
    int rc = check_launch("contour_trace_kernel");
    if (rc) return rc;
    simpson_diameters_kernel<<<(unsigned)((n_masks + SP_WARPS - 1) / SP_WARPS), SP_WARPS * 32, idx_smem, st>>>(
        masks, pts, npts, n_masks, row_index, col_index, vh, vw, spacing, diam_out, step_out);
    rc = check_launch("simpson_diameters_kernel");
    if (rc) return rc;
    if (npts_out) CUB200_CUDA_OK(cudaMemcpyAsync(npts_out, npts, sizeof(int32_t) * n_masks, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int cub200_simpson_diameters(const uint32_t* masks, int64_t n_masks, double spacing, double* diam_out,
                                        double* step_out, int32_t* npts_out, void* workspace, void* stream) {
    if (n_masks == 0) return 0;
    if (!masks || !diam_out || !step_out || !workspace || n_masks < 0 || !(spacing > 0.0)) return CUB200_EINVAL;
    return simpson_launch(masks, n_masks, nullptr, nullptr, 256, 256, spacing, diam_out, step_out, npts_out, workspace, (cudaStream_t)stream);
}

extern "C" int cub200_simpson_diameters_stretched(const uint32_t* masks, int64_t n_masks, const uint16_t* row_index,
                                                  const uint16_t* col_index, int height, int width, double spacing,
                                                  double* diam_out, double* step_out, int32_t* npts_out, void* workspace,
                                                  void* stream) {
    if (n_masks == 0) return 0;
    if (!masks || !row_index || !col_index || !diam_out || !step_out || !workspace || n_masks < 0 || !(spacing > 0.0) ||
        height < 2 || width < 2 || height > CUB200_SIMPSON_MAX_SIDE || width > CUB200_SIMPSON_MAX_SIDE)
        return CUB200_EINVAL;
    return simpson_launch(masks, n_masks, row_index, col_index, height, width, spacing, diam_out, step_out, npts_out, workspace,
                          (cudaStream_t)stream);
}

extern "C" int cub200_simpson_volumes(const double* diam, const double* step, int64_t n_patients, int64_t n_samples,
                                      double* edv_out, double* esv_out, double* ef_out, void* stream) {
    using namespace cub200;
    if (n_patients == 0 || n_samples == 0) return 0;
    if (!diam || !step || !edv_out || !esv_out || !ef_out || n_patients < 0 || n_samples < 0) return CUB200_EINVAL;
    const int64_t n = n_patients * n_samples;
    simpson_volumes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(diam, step, n_patients, n_samples,
                                                                                          edv_out, esv_out, ef_out);
    return check_launch("simpson_volumes_kernel");
}

// K1  dsnt_moments: fused flat-softmax + DSNT expectation + central second moments.
//
// Replaces flat_softmax + dsnt + pixel rescale of the reference
// (contour_uncertainty/task/regression/dsnt/utils.py:7-47,71-77,95-105; dsnt_al.py:118-131).
//
// HBM-bound: every logit is read exactly once (262 144 B per heatmap) and 44 B are written.
// One CTA (8 warps) per heatmap; warp w streams rows [32w, 32w+32) with two coalesced 512-B
// float4 loads per row (lane l owns columns 4l..4l+3 and 128+4l..128+4l+3), 4 rows (8 float4 per
// thread, 32 KB per CTA) in flight.  Online softmax with a WARP-UNIFORM running maximum
// (redux.sync on an order-preserving integer image of the float) so the rescale branch never
// diverges.  Second moments are accumulated per owned column as (C, R, RR) = sum w, sum w dy,
// sum w dy^2 about a row centre that follows the running maximum (exact shift identities), which
// keeps sum w dy^2 ~ W var instead of W (var + offset^2): the fp32 cancellation that a raw-moment
// one-pass kernel suffers (1e-4 rel) never builds up.  The x moments are recovered
// at the end from the 8 column sums, and every cross-warp merge / final subtraction runs in fp64.
//
// With a workspace the launch is PERSISTENT: one wave of resident CTAs (3 per SM), each taking an equal, contiguous range
// of the batch's rows (n_maps x 256).  A heatmap that straddles two ranges is reduced in two parts: each CTA reduces its
// part to 6 fp64 moments about its own centre, publishes them, takes a ticket, and the CTA that arrives last re-centres and
// adds the parts in part order (deterministic) -- the same exact shift identities as the cross-warp merge.  Without it the
// headline's 1344 heatmaps are 3.03 waves of 444 CTAs and the last, nearly empty wave costs the time of a whole heatmap.
#include "common.cuh"

namespace cub200 {

#ifndef CUB200_DSNT_RESIDENT
#define CUB200_DSNT_RESIDENT 3   // CTAs per SM of the persistent launch (= what __launch_bounds__ below allows)
#endif
#ifndef CUB200_DSNT_WARPS
#define CUB200_DSNT_WARPS 8      // warps per CTA
#endif
#ifndef CUB200_DSNT_MINBLOCKS
#define CUB200_DSNT_MINBLOCKS 3
#endif
constexpr int DSNT_NW = CUB200_DSNT_WARPS;
constexpr int DSNT_THREADS = 32 * DSNT_NW;
constexpr int DSNT_W = 256;
constexpr int ROWS_PER_TILE = 4;
constexpr float LOG2E = 1.4426950408889634f;

// One slab of a split heatmap as published in the workspace: moments about (rc, cc), weights relative to exp(m).
struct __align__(16) DsntPartial {
    double a[6];
    float m;
    int rc, cc, pad;
};
constexpr int DSNT_MAX_SPLIT = 8;
struct DsntSlot {
    DsntPartial part[DSNT_MAX_SPLIT];
    unsigned ticket, pad[3];
};

template <bool IS_LOGIT>
__global__ void __launch_bounds__(DSNT_THREADS, CUB200_DSNT_MINBLOCKS)
dsnt_moments_kernel(const float* __restrict__ heat, int64_t n_maps, int use_covar,
                    float* __restrict__ norm_out, float* __restrict__ pix_mu, float* __restrict__ pix_sigma,
                    int64_t rows_per_cta, DsntSlot* __restrict__ slots) {
    __shared__ float s_m[DSNT_NW];
    __shared__ int s_rc[DSNT_NW], s_cc[DSNT_NW];
    __shared__ double s_red[DSNT_NW][6];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total_rows = n_maps * DSNT_W;
    const int64_t r_end = min((int64_t)(blockIdx.x + 1) * rows_per_cta, total_rows);
    for (int64_t r = (int64_t)blockIdx.x * rows_per_cta; r < r_end;) {
        // rows [lo, hi) of heatmap `map` (lo, hi multiples of ROWS_PER_TILE)
        const int64_t map = r / DSNT_W;
        const int lo = (int)(r - map * DSNT_W);
        const int hi = (int)min((int64_t)DSNT_W, lo + (r_end - r));
        r += hi - lo;
        const float4* base = reinterpret_cast<const float4*>(heat + map * (int64_t)(DSNT_W * DSNT_W));

        float C[8], R[8], RR[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) C[j] = R[j] = RR[j] = 0.f;
        float m = -INFINITY;                   // running max (logit, or heat value when !IS_LOGIT)
        int rc = lo;                           // row centre of the dy moments (warp-uniform)
        int cc = 0;                            // column of the running max (warp-uniform)

        // the 8 warps take the tiles of the part round robin: at any time the CTA streams 32 consecutive rows
        for (int r_tile = lo + warp * ROWS_PER_TILE; r_tile < hi; r_tile += DSNT_NW * ROWS_PER_TILE) {
            float4 v[ROWS_PER_TILE][2];
#pragma unroll
            for (int q = 0; q < ROWS_PER_TILE; ++q) {
                const float4* rowp = base + (int64_t)(r_tile + q) * (DSNT_W / 4);
                v[q][0] = ld_stream_f4(rowp + lane);
                v[q][1] = ld_stream_f4(rowp + 32 + lane);
            }
            {
                float tm = -INFINITY;
#pragma unroll
                for (int q = 0; q < ROWS_PER_TILE; ++q) {
                    tm = fmaxf(tm, fmaxf(fmaxf(v[q][0].x, v[q][0].y), fmaxf(v[q][0].z, v[q][0].w)));
                    tm = fmaxf(tm, fmaxf(fmaxf(v[q][1].x, v[q][1].y), fmaxf(v[q][1].z, v[q][1].w)));
                }
                const int wmax_i = __reduce_max_sync(FULL, float_ordered(tm));
                const float wmax = ordered_float(wmax_i);
                if (wmax > m) {  // warp-uniform
                    // which lane/column holds it (first lane wins); row centre moves to this tile
                    const unsigned who = __ballot_sync(FULL, tm == wmax);
                    const int src = __ffs(who) - 1;
                    int mycol = 0;
                    if (lane == src) {
#pragma unroll
                        for (int q = 0; q < ROWS_PER_TILE; ++q) {
                            const float a[8] = {v[q][0].x, v[q][0].y, v[q][0].z, v[q][0].w,
                                                v[q][1].x, v[q][1].y, v[q][1].z, v[q][1].w};
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (a[j] == wmax) mycol = (j < 4) ? 4 * lane + j : 128 + 4 * lane + (j - 4);
                        }
                    }
                    cc = __shfl_sync(FULL, mycol, src);
                    const float sc = IS_LOGIT ? ex2_approx((m - wmax) * LOG2E) : 1.f;  // m = -inf -> 0
                    const float d = (float)(r_tile - rc);             // new centre - old centre
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // moments about the new centre: dy' = dy - d
                        const float c = C[j] * sc, r1 = R[j] * sc, r2 = RR[j] * sc;
                        RR[j] = r2 - d * (2.f * r1 - d * c);
                        R[j] = r1 - d * c;
                        C[j] = c;
                    }
                    m = wmax;
                    rc = r_tile;
                }
            }
            const float mb = m * LOG2E;
#pragma unroll
            for (int q = 0; q < ROWS_PER_TILE; ++q) {
                const float dy = (float)(r_tile + q - rc);
                const float dy2 = dy * dy;
                const float a[8] = {v[q][0].x, v[q][0].y, v[q][0].z, v[q][0].w,
                                    v[q][1].x, v[q][1].y, v[q][1].z, v[q][1].w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float w = IS_LOGIT ? ex2_approx(fmaf(a[j], LOG2E, -mb)) : a[j];
                    C[j] += w;
                    R[j] = fmaf(w, dy, R[j]);
                    RR[j] = fmaf(w, dy2, RR[j]);
                }
            }
        }

        // ---- cross-warp merge in fp64 about the centre (rc*, cc*) of the warp holding the part's max (a warp that had no
        //      tile has m = -inf and zero moments: it contributes exactly 0)
        __syncthreads();                                      // the previous part's readers of s_* are done
        if (lane == 0) { s_m[warp] = m; s_rc[warp] = rc; s_cc[warp] = cc; }
        __syncthreads();
        float mstar = s_m[0];
        int best = 0;
#pragma unroll
        for (int w = 1; w < DSNT_NW; ++w)
            if (s_m[w] > mstar) { mstar = s_m[w]; best = w; }
        const int rcb = s_rc[best];
        const int ccb = s_cc[best];
        const double sc = IS_LOGIT ? (double)ex2_approx((m - mstar) * LOG2E) : 1.0;
        const double d = (double)(rcb - rc);
        double W = 0, Mx = 0, My = 0, Mxx = 0, Myy = 0, Mxy = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = (j < 4) ? 4 * lane + j : 128 + 4 * lane + (j - 4);
            const double dx = (double)(col - ccb);
            const double c = (double)C[j], r1 = (double)R[j], r2 = (double)RR[j];
            const double r1s = r1 - d * c;                     // about rc*
            const double r2s = r2 - d * (2.0 * r1 - d * c);
            W += c;
            Mx += c * dx;
            Mxx += c * dx * dx;
            My += r1s;
            Myy += r2s;
            Mxy += r1s * dx;
        }
        W = warp_sum(W * sc); Mx = warp_sum(Mx * sc); My = warp_sum(My * sc);
        Mxx = warp_sum(Mxx * sc); Myy = warp_sum(Myy * sc); Mxy = warp_sum(Mxy * sc);
        if (lane == 0) {
            s_red[warp][0] = W; s_red[warp][1] = Mx; s_red[warp][2] = My;
            s_red[warp][3] = Mxx; s_red[warp][4] = Myy; s_red[warp][5] = Mxy;
        }
        __syncthreads();
        if (threadIdx.x != 0) continue;

        double a[6] = {0, 0, 0, 0, 0, 0};
        for (int w = 0; w < DSNT_NW; ++w)
            for (int q = 0; q < 6; ++q) a[q] += s_red[w][q];
        int rcs = rcb, ccs = ccb;                              // centre of the moments (of the heatmap's after a merge)
        if (hi - lo != DSNT_W) {
            // a part of a heatmap: publish it, take a ticket; the last part to arrive merges all of them in part order
            const int64_t first_cta = (map * DSNT_W) / rows_per_cta;
            const int n_parts = (int)((map * DSNT_W + DSNT_W - 1) / rows_per_cta - first_cta) + 1;
            DsntSlot& slot = slots[map];
            DsntPartial& me = slot.part[blockIdx.x - first_cta];
#pragma unroll
            for (int q = 0; q < 6; ++q) me.a[q] = a[q];
            me.m = mstar; me.rc = rcs; me.cc = ccs;
            __threadfence();
            const unsigned ticket = atomicAdd(&slot.ticket, 1u);
            if (ticket != (unsigned)n_parts - 1u) continue;
            __threadfence();
            slot.ticket = 0u;                                 // the workspace is ready for the next launch
            const volatile DsntPartial* pp = slot.part;
            float mg = pp[0].m;
            int bp = 0;
            for (int p = 1; p < n_parts; ++p)
                if (pp[p].m > mg) { mg = pp[p].m; bp = p; }
            rcs = pp[bp].rc; ccs = pp[bp].cc;
#pragma unroll
            for (int q = 0; q < 6; ++q) a[q] = 0.0;
            for (int p = 0; p < n_parts; ++p) {
                const double scp = IS_LOGIT ? (double)ex2_approx((pp[p].m - mg) * LOG2E) : 1.0;
                const double dxc = (double)(pp[p].cc - ccs), dyc = (double)(pp[p].rc - rcs);
                const double w = pp[p].a[0], mx = pp[p].a[1], my = pp[p].a[2];
                a[0] += scp * w;
                a[1] += scp * (mx + dxc * w);
                a[2] += scp * (my + dyc * w);
                a[3] += scp * (pp[p].a[3] + dxc * (2.0 * mx + dxc * w));
                a[4] += scp * (pp[p].a[4] + dyc * (2.0 * my + dyc * w));
                a[5] += scp * (pp[p].a[5] + dxc * my + dyc * mx + dxc * dyc * w);
            }
        }
        const double iw = 1.0 / a[0];
        const double ex = a[1] * iw, ey = a[2] * iw;  // offsets from the centre, in pixels
        const double vx = a[3] * iw - ex * ex, vy = a[4] * iw - ey * ey;
        const double cxy = a[5] * iw - ex * ey;
        const double px = (double)ccs + ex, py = (double)rcs + ey;
        // For already-normalised heatmaps (IS_LOGIT = false) the reference does NOT divide by the
        // total mass: coords = sum h X etc.  Keep that: scale back by the mass W.
        if (norm_out) {
            const double g = 2.0 / DSNT_W;  // pixel -> normalised scale
            float* o = norm_out + map * 5;
            if (IS_LOGIT) {
                o[0] = (float)((2.0 * px + 1.0) / DSNT_W - 1.0);
                o[1] = (float)((2.0 * py + 1.0) / DSNT_W - 1.0);
                o[2] = (float)(vx * g * g);
                o[3] = (float)(vy * g * g);
                o[4] = (float)(cxy * g * g);
            } else {
                // x = sum h X ; var_x = sum h (X - x)^2 with total mass W (not assumed 1)
                const double xn = (2.0 * (double)ccs + 1.0) / DSNT_W - 1.0, yn = (2.0 * (double)rcs + 1.0) / DSNT_W - 1.0;
                const double Wt = a[0];
                const double sx = Wt * xn + a[1] * g, sy = Wt * yn + a[2] * g;  // sum h X, sum h Y
                const double ox = xn - sx, oy = yn - sy;                        // centre - mean
                o[0] = (float)sx;
                o[1] = (float)sy;
                o[2] = (float)(a[3] * g * g + 2.0 * ox * a[1] * g + ox * ox * Wt);
                o[3] = (float)(a[4] * g * g + 2.0 * oy * a[2] * g + oy * oy * Wt);
                o[4] = (float)(a[5] * g * g + ox * a[2] * g + oy * a[1] * g + ox * oy * Wt);
            }
        }
        if (pix_mu) {
            pix_mu[map * 2 + 0] = (float)px;
            pix_mu[map * 2 + 1] = (float)py;
        }
        if (pix_sigma) {
            float* o = pix_sigma + map * 4;
            o[0] = (float)vx;
            o[1] = o[2] = use_covar ? (float)cxy : 0.f;
            o[3] = (float)vy;
        }
    }
}

}  // namespace cub200

extern "C" int64_t cub200_dsnt_workspace_bytes(int64_t n_maps) {
    return n_maps < 0 ? 0 : n_maps * (int64_t)sizeof(cub200::DsntSlot);
}

extern "C" int cub200_dsnt_moments(const float* heat, int64_t n_maps, int h, int w, int is_logit, int use_covar,
                                   float* norm_out, float* pix_mu, float* pix_sigma, void* workspace, void* stream) {
    using namespace cub200;
    if (!heat || n_maps < 0 || h != DSNT_W || w != DSNT_W) return CUB200_EINVAL;
    if (n_maps == 0) return 0;
    if (n_maps > 0x7fffffff / DSNT_MAX_SPLIT) return CUB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    // with a workspace: one wave of resident CTAs over equal row ranges (multiples of the 4-row tile, >= 64 rows so that a
    // heatmap has at most 5 parts); without: one CTA per heatmap
    int64_t rows_per_cta = DSNT_W;
    if (workspace) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t total_rows = n_maps * DSNT_W, resident = (int64_t)CUB200_DSNT_RESIDENT * sms;
        // measured (tools/bench_dsnt.py): equal row ranges win when the one-CTA-per-heatmap grid would be 1..6 waves (the
        // headline: 1344 heatmaps, 66.0 -> 63.6 us) or a small fraction of one (21 heatmaps: 16.4 -> 14.4 us); with many
        // waves the dynamic CTA scheduler already balances, and staggers the per-heatmap epilogues better (512 images: 404 us
        // against 429 us)
        if ((n_maps >= resident && n_maps < 6 * resident) || 4 * n_maps <= resident) {
            rows_per_cta = ((total_rows + resident - 1) / resident + ROWS_PER_TILE - 1) / ROWS_PER_TILE * ROWS_PER_TILE;
            if (rows_per_cta < 64) rows_per_cta = 64;
        }
    }
    const unsigned grid = (unsigned)((n_maps * DSNT_W + rows_per_cta - 1) / rows_per_cta);
    DsntSlot* slots = reinterpret_cast<DsntSlot*>(workspace);
    if (is_logit)
        dsnt_moments_kernel<true><<<grid, DSNT_THREADS, 0, st>>>(heat, n_maps, use_covar, norm_out, pix_mu, pix_sigma, rows_per_cta, slots);
    else
        dsnt_moments_kernel<false><<<grid, DSNT_THREADS, 0, st>>>(heat, n_maps, use_covar, norm_out, pix_mu, pix_sigma, rows_per_cta, slots);
    return check_launch("dsnt_moments_kernel");
}

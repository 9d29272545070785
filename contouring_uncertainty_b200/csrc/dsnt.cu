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
// Small batches (the headline's 64 images = 1344 heatmaps = 3.03 waves of the 444 resident CTAs) lose the time of one whole
// heatmap to the last, nearly empty wave.  With a workspace the launch cuts every heatmap into SPLIT row slabs (2, 4 or 8,
// chosen so that the grid is >= ~10 waves): a CTA reduces its slab to 6 fp64 moments about its own centre, publishes them,
// takes a ticket, and the last CTA of a heatmap re-centres and adds the slabs in slab order (deterministic) -- the same
// exact shift identities as the cross-warp merge.
#include "common.cuh"

namespace cub200 {

constexpr int DSNT_THREADS = 256;
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
__global__ void __launch_bounds__(DSNT_THREADS, 3)
dsnt_moments_kernel(const float* __restrict__ heat, int64_t n_maps, int use_covar,
                    float* __restrict__ norm_out, float* __restrict__ pix_mu, float* __restrict__ pix_sigma,
                    int split, DsntSlot* __restrict__ slots) {
    __shared__ float s_m[8];
    __shared__ int s_rc[8], s_cc[8];
    __shared__ double s_red[8][6];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t map = blockIdx.x / split;
    const int part = blockIdx.x % split;
    const int rows_per_warp = 32 / split;                    // 32, 16, 8 or 4 (a multiple of ROWS_PER_TILE)
    const float4* base = reinterpret_cast<const float4*>(heat + map * (int64_t)(DSNT_W * DSNT_W));
    const int row0 = part * (DSNT_W / split) + warp * rows_per_warp;

    float C[8], R[8], RR[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) C[j] = R[j] = RR[j] = 0.f;
    float m = -INFINITY;                   // running max (logit, or heat value when !IS_LOGIT)
    int rc = row0;                         // row centre of the dy moments (warp-uniform)
    int cc = 0;                            // column of the running max (warp-uniform)

    for (int t = 0; t < rows_per_warp / ROWS_PER_TILE; ++t) {
        const int r_tile = row0 + t * ROWS_PER_TILE;
        float4 v[ROWS_PER_TILE][2];
#pragma unroll
        for (int r = 0; r < ROWS_PER_TILE; ++r) {
            const float4* rowp = base + (int64_t)(r_tile + r) * (DSNT_W / 4);
            v[r][0] = ld_stream_f4(rowp + lane);
            v[r][1] = ld_stream_f4(rowp + 32 + lane);
        }
        {
            float tm = -INFINITY;
#pragma unroll
            for (int r = 0; r < ROWS_PER_TILE; ++r) {
                tm = fmaxf(tm, fmaxf(fmaxf(v[r][0].x, v[r][0].y), fmaxf(v[r][0].z, v[r][0].w)));
                tm = fmaxf(tm, fmaxf(fmaxf(v[r][1].x, v[r][1].y), fmaxf(v[r][1].z, v[r][1].w)));
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
                    for (int r = 0; r < ROWS_PER_TILE; ++r) {
                        const float a[8] = {v[r][0].x, v[r][0].y, v[r][0].z, v[r][0].w,
                                            v[r][1].x, v[r][1].y, v[r][1].z, v[r][1].w};
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
        for (int r = 0; r < ROWS_PER_TILE; ++r) {
            const float dy = (float)(r_tile + r - rc);
            const float dy2 = dy * dy;
            const float a[8] = {v[r][0].x, v[r][0].y, v[r][0].z, v[r][0].w,
                                v[r][1].x, v[r][1].y, v[r][1].z, v[r][1].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float w = IS_LOGIT ? ex2_approx(fmaf(a[j], LOG2E, -mb)) : a[j];
                C[j] += w;
                R[j] = fmaf(w, dy, R[j]);
                RR[j] = fmaf(w, dy2, RR[j]);
            }
        }
    }

    // ---- cross-warp merge in fp64 about the centre (rc*, cc*) of the warp holding the global max
    if (lane == 0) { s_m[warp] = m; s_rc[warp] = rc; s_cc[warp] = cc; }
    __syncthreads();
    float mstar = s_m[0];
    int best = 0;
#pragma unroll
    for (int w = 1; w < 8; ++w)
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
    if (threadIdx.x == 0) {
        double a[6] = {0, 0, 0, 0, 0, 0};
        for (int w = 0; w < 8; ++w)
            for (int q = 0; q < 6; ++q) a[q] += s_red[w][q];
        int rcs = s_rc[best], ccs = s_cc[best];                // centre of this CTA's moments (of the heatmap's after a merge)
        if (split > 1) {
            // publish this slab, take a ticket; the last slab of the heatmap merges all of them in slab order
            DsntSlot& slot = slots[map];
            DsntPartial& me = slot.part[part];
#pragma unroll
            for (int q = 0; q < 6; ++q) me.a[q] = a[q];
            me.m = mstar; me.rc = rcs; me.cc = ccs;
            __threadfence();
            const unsigned ticket = atomicAdd(&slot.ticket, 1u);
            if (ticket != (unsigned)split - 1u) return;
            __threadfence();
            slot.ticket = 0u;                                 // the workspace is ready for the next launch
            const volatile DsntPartial* pp = slot.part;
            float mg = pp[0].m;
            int bp = 0;
            for (int p = 1; p < split; ++p)
                if (pp[p].m > mg) { mg = pp[p].m; bp = p; }
            rcs = pp[bp].rc; ccs = pp[bp].cc;
#pragma unroll
            for (int q = 0; q < 6; ++q) a[q] = 0.0;
            for (int p = 0; p < split; ++p) {
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
            float* s = pix_sigma + map * 4;
            s[0] = (float)vx;
            s[1] = s[2] = use_covar ? (float)cxy : 0.f;
            s[3] = (float)vy;
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
    // slabs per heatmap: the smallest power of two that makes the grid >= 10 waves of the resident CTAs (3 per SM)
    int split = 1;
    if (workspace) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        while (split < DSNT_MAX_SPLIT && n_maps * split < 10LL * 3 * sms) split *= 2;
    }
    const unsigned grid = (unsigned)(n_maps * split);
    DsntSlot* slots = reinterpret_cast<DsntSlot*>(workspace);
    if (is_logit)
        dsnt_moments_kernel<true><<<grid, DSNT_THREADS, 0, st>>>(heat, n_maps, use_covar, norm_out, pix_mu, pix_sigma, split, slots);
    else
        dsnt_moments_kernel<false><<<grid, DSNT_THREADS, 0, st>>>(heat, n_maps, use_covar, norm_out, pix_mu, pix_sigma, split, slots);
    return check_launch("dsnt_moments_kernel");
}

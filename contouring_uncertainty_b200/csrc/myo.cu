// Two-structure masks: USContourToMask when Label.MYO is among the labels.
//
// Replaces (reference paths, cu/ = contour_uncertainty/):
//   USContourToMask.__call__, MYO branch    cu/data/camus/utils.py:47-82
//   split_landmarks, poly2mask              cu/data/camus/utils.py:14-28
// Third party restated: skimage.draw.polygon (point-in-polygon over the bounding box; oracle/port.py::draw_polygon --
// scikit-image is absent from the image: PARITY UNPINNED for it).
//
// 42 rounded landmarks = endocardium (21) + epicardium (21).  One warp per landmark set:
//   LV mask   = reconstruction(lv) (or linear_reconstruction), exactly as fused.cu builds it;
//   polygon   = the 1000 rounded endocardial spline points followed by the 1000 epicardial ones in reverse order (the raw
//               21 points of a contour whose spline FITPACK refuses), NOT clipped to the frame;
//   MYO mask  = pixels the polygon covers (inside, on an edge or on a vertex) that are not LV.
// skimage classifies a pixel by two crossing counts along its row: r_cross (edges that straddle the row with the
// half-open rule y0 > 0 != y1 > 0 and cross it to the right of the pixel) and l_cross (y0 < 0 != y1 < 0, to the left);
// it is covered iff it is a vertex, r_cross is odd, or the parities differ -- i.e. iff vertex | r_odd | l_odd.  An edge
// crosses row r at the rational abscissa N / D (integers), so "to the right of column c" is c < N / D: each straddling
// (edge, row) pair toggles a prefix [0, ceil(N / D) - 1] of the row's r-parity bits and a suffix [floor(N / D) + 1, 255]
// of its l-parity bits, with shared-memory atomicXor.  All of it in exact integer arithmetic.
#include "flood.cuh"
#include "spline.cuh"

namespace cub200 {

constexpr int MY_WARPS = 2;
constexpr int MY_NPTS = 1000;
constexpr int MY_K = CUB200_K;
constexpr int MY_CLAMP = 8192;

struct __align__(16) MyoScratch {
    uint32_t bmp[BMP_WORDS];       // LV boundary pixels (obstacles of the flood)
    uint32_t pr[BMP_WORDS];        // parity of r_cross
    uint32_t pl[BMP_WORDS];        // parity of l_cross
    uint32_t pv[BMP_WORDS];        // polygon vertices
    SplineScratch sp;
    int2 verts[2 * MY_NPTS];       // polygon vertices (x, y)
};

// Evaluate the fitted spline at j / 999, j = 0..999 (32 consecutive points per lane), round half to even.
// store(j, ix, iy) receives every point (lane 31 repeats 24 points of lane 30).
template <typename F>
__device__ __forceinline__ void eval_spline_points(const SplineScratch& sp, int lane, F store) {
    const int j0 = min(lane * 32, MY_NPTS - 32);
    const double step = 1.0 / (double)(MY_NPTS - 1);
    int piece = 0;
    {
        const double us = (double)j0 * step;
#pragma unroll
        for (int k = 1; k <= MY_K - 2; ++k) piece = (us >= sp.u[k]) ? k : piece;
    }
    for (int jj = 0; jj < 32; ++jj) {
        const double uj = (double)(j0 + jj) * step;
        while (piece < MY_K - 2 && uj >= sp.piece[piece][0]) ++piece;
        const double* c = sp.piece[piece];
        const double sv = uj - c[1];
        const double xv = fma(fma(fma(c[5], sv, c[4]), sv, c[3]), sv, c[2]);
        const double yv = fma(fma(fma(c[9], sv, c[8]), sv, c[7]), sv, c[6]);
        store(j0 + jj, __double2int_rn(xv), __double2int_rn(yv));
    }
}

__device__ __forceinline__ void xor_range(uint32_t* bmp, int row, int a, int b) {      // columns a..b, 0 <= a <= b <= 255
    for (int w = a >> 5; w <= (b >> 5); ++w) {
        const int lo = max(a - 32 * w, 0), hi = min(b - 32 * w, 31);
        atomicXor(&bmp[bmp_addr(row, w)], (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo));
    }
}

__global__ void __launch_bounds__(MY_WARPS * 32)
us_myo_kernel(const float* __restrict__ landmarks, int64_t n, int flags, uint32_t* __restrict__ lv_out,
              uint32_t* __restrict__ myo_out, int32_t* __restrict__ area_out) {
    extern __shared__ __align__(16) unsigned char my_smem[];
    MyoScratch& S = reinterpret_cast<MyoScratch*>(my_smem)[threadIdx.x >> 5];
    const int lane = lane_id();
    const int64_t g = blockIdx.x * (int64_t)MY_WARPS + (threadIdx.x >> 5);
    if (g >= n) return;
    const bool linear = (flags & 2) != 0;
    {
        uint4* b4 = reinterpret_cast<uint4*>(S.bmp);           // bmp, pr, pl, pv are contiguous
        for (int i = lane; i < 4 * BMP_WORDS / 4; i += 32) b4[i] = make_uint4(0, 0, 0, 0);
    }
    // landmarks.round().astype(int): lane k < 21 holds endocardial point k and epicardial point k
    int lx = 0, ly = 0, ex = 0, ey = 0;
    if (lane < MY_K) {
        const float2 a = reinterpret_cast<const float2*>(landmarks)[g * 2 * MY_K + lane];
        const float2 b = reinterpret_cast<const float2*>(landmarks)[g * 2 * MY_K + MY_K + lane];
        const float cl = (float)MY_CLAMP;
        lx = __float2int_rn(fminf(fmaxf(a.x, -cl), cl)); ly = __float2int_rn(fminf(fmaxf(a.y, -cl), cl));
        ex = __float2int_rn(fminf(fmaxf(b.x, -cl), cl)); ey = __float2int_rn(fminf(fmaxf(b.y, -cl), cl));
    }
    __syncwarp();
    // ---- endocardium: spline points -> polygon vertices 0.. and (spline mode) the LV boundary pixels
    int n_lv = MY_NPTS;
    {
        const bool bad = fit_spline(S.sp, (double)lx, (double)ly, lane);
        bool oob = false;
        if (!bad) {
            int lowest = 0;
            eval_spline_points(S.sp, lane, [&](int j, int ix, int iy) {
                S.verts[j] = make_int2(ix, iy);
                if (!linear) {
                    const int cx = min(ix, 255), cy = min(iy, 255);
                    lowest = min(lowest, min(cx, cy));
                    atomicOr(&S.bmp[bmp_addr(cy & 255, (cx & 255) >> 5)], 1u << (cx & 31));
                }
            });
            oob = __any_sync(FULL, lowest < -256);
        } else {
            n_lv = MY_K;
            if (lane < MY_K) S.verts[lane] = make_int2(lx, ly);
            const int cx = min(lx, 255), cy = min(ly, 255);
            oob = __any_sync(FULL, lane < MY_K && (cx < -256 || cy < -256));
            if (!linear && !oob && lane < MY_K) atomicOr(&S.bmp[bmp_addr(cy & 255, (cx & 255) >> 5)], 1u << (cx & 31));
        }
        __syncwarp();
        if (oob && !linear) {                                   // IndexError in reconstruction(): nothing of the spline is drawn
            uint4* b4 = reinterpret_cast<uint4*>(S.bmp);
            for (int i = lane; i < BMP_WORDS / 4; i += 32) b4[i] = make_uint4(0, 0, 0, 0);
            __syncwarp();
        }
        for (int e = linear ? 0 : MY_K - 1; e < MY_K; ++e) {
            const int e1 = (e + 1 == MY_K) ? 0 : e + 1;
            draw_line_clipped(S.bmp, __shfl_sync(FULL, ly, e), __shfl_sync(FULL, lx, e), __shfl_sync(FULL, ly, e1),
                              __shfl_sync(FULL, lx, e1), lane);
        }
    }
    __syncwarp();
    // ---- epicardium: spline points -> polygon vertices in reverse order
    int n_ep = MY_NPTS;
    {
        const bool bad = fit_spline(S.sp, (double)ex, (double)ey, lane);
        if (!bad) {
            eval_spline_points(S.sp, lane, [&](int j, int ix, int iy) { S.verts[n_lv + (MY_NPTS - 1 - j)] = make_int2(ix, iy); });
        } else {
            n_ep = MY_K;
            if (lane < MY_K) S.verts[n_lv + (MY_K - 1 - lane)] = make_int2(ex, ey);
        }
    }
    __syncwarp();
    const int nv = n_lv + n_ep;
    // ---- polygon coverage: parity of the crossings right / left of every pixel, vertices
    for (int e = lane; e < nv; e += 32) {
        const int2 a = S.verts[e], b = S.verts[e == 0 ? nv - 1 : e - 1];        // edge (e - 1, e): point 0 = e, point 1 = e - 1
        if (a.x >= 0 && a.x <= 255 && a.y >= 0 && a.y <= 255) atomicOr(&S.pv[bmp_addr(a.y, a.x >> 5)], 1u << (a.x & 31));
        if (a.y == b.y) continue;
        const int ymin = min(a.y, b.y), ymax = max(a.y, b.y);
        const long long D = (long long)b.y - a.y;
        for (int r = max(ymin, 0); r <= min(ymax, 255); ++r) {
            const long long N = (long long)a.x * (b.y - r) - (long long)b.x * (a.y - r);
            long long q = N / D;                                                  // truncates: make it a floor
            if ((N % D != 0) && ((N < 0) != (D < 0))) --q;
            const bool exact = (q * D == N);
            if (r < ymax) {                                                       // (y0 > 0) != (y1 > 0): crossing to the right <=> c < N / D
                const long long hi = exact ? q - 1 : q;
                if (hi >= 0) xor_range(S.pr, r, 0, (int)min(hi, 255LL));
            }
            if (r > ymin) {                                                       // (y0 < 0) != (y1 < 0): crossing to the left <=> c > N / D
                const long long lo = q + 1;
                if (lo <= 255) xor_range(S.pl, r, (int)max(lo, 0LL), 255);
            }
        }
    }
    __syncwarp();
    // ---- LV = not reached by the background flood; MYO = covered by the polygon and not LV
    uint32_t R[8][8];
    const uint4* band4 = reinterpret_cast<const uint4*>(S.bmp + lane * BAND_STRIDE);
    flood_background(R, band4, lane);
    const uint4* pr4 = reinterpret_cast<const uint4*>(S.pr + lane * BAND_STRIDE);
    const uint4* pl4 = reinterpret_cast<const uint4*>(S.pl + lane * BAND_STRIDE);
    const uint4* pv4 = reinterpret_cast<const uint4*>(S.pv + lane * BAND_STRIDE);
    int a_lv = 0, a_myo = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t cov[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint4 x = pr4[2 * i + h], y = pl4[2 * i + h], z = pv4[2 * i + h];
            cov[4 * h] = x.x | y.x | z.x; cov[4 * h + 1] = x.y | y.y | z.y;
            cov[4 * h + 2] = x.z | y.z | z.z; cov[4 * h + 3] = x.w | y.w | z.w;
        }
        uint32_t lv[8], my[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            lv[w] = ~R[i][w];
            my[w] = cov[w] & R[i][w];
            a_lv += __popc(lv[w]);
            a_myo += __popc(my[w]);
        }
        const int64_t off = g * CUB200_MASK_WORDS + (lane * 8 + i) * 8;
        if (lv_out) {
            uint4* o = reinterpret_cast<uint4*>(lv_out + off);
            o[0] = make_uint4(lv[0], lv[1], lv[2], lv[3]);
            o[1] = make_uint4(lv[4], lv[5], lv[6], lv[7]);
        }
        if (myo_out) {
            uint4* o = reinterpret_cast<uint4*>(myo_out + off);
            o[0] = make_uint4(my[0], my[1], my[2], my[3]);
            o[1] = make_uint4(my[4], my[5], my[6], my[7]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a_lv += __shfl_xor_sync(FULL, a_lv, o);
        a_myo += __shfl_xor_sync(FULL, a_myo, o);
    }
    if (area_out && lane == 0) { area_out[2 * g] = a_lv; area_out[2 * g + 1] = a_myo; }
}

// Aggregation of LV + MYO sample stacks: per pixel the class counts over the T samples of an image (BG = neither bit) and
// UncertaintyTask.sample_entropy of the one-hot stack (cu/task/uncertainty.py:107-133, C = 3): y_hat = mean over the samples
// (counts / T, fp64), scipy.stats.entropy(y_hat, axis=0, base=3): pk = y_hat / sum(y_hat), S = sum(entr(pk)) / ln 3,
// non-finite -> 0.  One thread per pixel; the 32 lanes of a warp read the same mask word (a broadcast).
__global__ void __launch_bounds__(256)
mask_class_entropy_kernel(const uint32_t* __restrict__ lv, const uint32_t* __restrict__ myo, int64_t n_img, int64_t T,
                          double* __restrict__ entropy_out, int32_t* __restrict__ count_out, uint8_t* __restrict__ pred_out) {
    const int64_t px = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;       // image * 65536 + row * 256 + col
    if (px >= n_img * 65536) return;
    const int64_t img = px >> 16;
    const int word = (int)((px & 65535) >> 5), bit = (int)(px & 31);
    const uint32_t* pl = lv + img * T * CUB200_MASK_WORDS + word;
    const uint32_t* pm = myo + img * T * CUB200_MASK_WORDS + word;
    int c1 = 0, c2 = 0;
    for (int64_t t = 0; t < T; ++t) {
        c1 += (pl[t * CUB200_MASK_WORDS] >> bit) & 1u;
        c2 += (pm[t * CUB200_MASK_WORDS] >> bit) & 1u;
    }
    const int c0 = (int)T - c1 - c2;
    if (count_out) {
        count_out[(img * 3 + 0) * 65536 + (px & 65535)] = c0;
        count_out[(img * 3 + 1) * 65536 + (px & 65535)] = c1;
        count_out[(img * 3 + 2) * 65536 + (px & 65535)] = c2;
    }
    if (pred_out) {                                                           // argmax over (BG, LV, MYO), first maximum wins
        int best = 0, cb = c0;
        if (c1 > cb) { best = 1; cb = c1; }
        if (c2 > cb) best = 2;
        pred_out[px] = (uint8_t)best;
    }
    if (entropy_out) {
        const double dT = (double)T;
        const double p0 = (double)c0 / dT, p1 = (double)c1 / dT, p2 = (double)c2 / dT;
        const double s = (p0 + p1) + p2;
        const double q0 = p0 / s, q1 = p1 / s, q2 = p2 / s;
        const double e0 = q0 > 0.0 ? -q0 * log(q0) : 0.0, e1 = q1 > 0.0 ? -q1 * log(q1) : 0.0, e2 = q2 > 0.0 ? -q2 * log(q2) : 0.0;
        double e = ((e0 + e1) + e2) / log(3.0);
        if (!isfinite(e)) e = 0.0;
        entropy_out[px] = e;
    }
}

}  // namespace cub200

extern "C" int cub200_mask_class_entropy(const uint32_t* lv_masks, const uint32_t* myo_masks, int64_t n_img, int64_t n_samples,
                                         double* entropy_out, int32_t* count_out, uint8_t* pred_out, void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!lv_masks || !myo_masks || n_img < 0 || n_samples < 1 || n_samples > 0x3fffffff || (!entropy_out && !count_out && !pred_out))
        return CUB200_EINVAL;
    const int64_t n = n_img * 65536;
    mask_class_entropy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lv_masks, myo_masks, n_img, n_samples,
                                                                                            entropy_out, count_out, pred_out);
    return check_launch("mask_class_entropy_kernel");
}

extern "C" int cub200_us_myo_masks(const float* landmarks, int64_t n, int flags, uint32_t* lv_out, uint32_t* myo_out,
                                   int32_t* area_out, void* stream) {
    using namespace cub200;
    if (n == 0) return 0;
    if (!landmarks || (!lv_out && !myo_out && !area_out) || n < 0 || n > 0x7ffffff0 || (flags & ~2)) return CUB200_EINVAL;
    if ((reinterpret_cast<uintptr_t>(landmarks) & 7) || (reinterpret_cast<uintptr_t>(lv_out) & 15) || (reinterpret_cast<uintptr_t>(myo_out) & 15))
        return CUB200_EINVAL;
    const int smem = MY_WARPS * (int)sizeof(MyoScratch);
    CUB200_CUDA_OK(cudaFuncSetAttribute(us_myo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    us_myo_kernel<<<(unsigned)((n + MY_WARPS - 1) / MY_WARPS), MY_WARPS * 32, smem, (cudaStream_t)stream>>>(landmarks, n, flags, lv_out,
                                                                                                          myo_out, area_out);
    return check_launch("us_myo_kernel");
}

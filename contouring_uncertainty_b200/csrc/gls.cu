// Mask-based longitudinal length of the left ventricle: the quantity EchoMeasure.gls compares between frames
// (second half).
//
// Replaces (reference paths, vv/ = vital/vital/):
//   EchoMeasure.gls, _lv_longitudinal_length (myo_labels=None)      vv/utils/image/us/measure.py:320-366
//   EchoMeasure._extract_landmarks_from_polar_contour (apex=False)  measure.py:22-151
//   EchoMeasure._endo_epi_contour                                   measure.py:153-195
//   cart2pol(sort_by_theta=True)                                    vv/utils/image/coord.py:6-45
//   the per-sample loop of GLS.compute                              contour_uncertainty/results/clinical/view.py:98-189
// Third party restated: skimage.measure.find_contours(mask, 0.9)[0] (contour_trace.cuh; parity unpinned),
// scipy.ndimage.center_of_mass, scipy.ndimage.gaussian_filter1d (mode 'reflect', truncate 4: the symmetric
// correlate1d sum in its own order, no FMA), scipy.signal.find_peaks(height=0) (plateau mid-points).
//
// One CTA (128 threads) per mask:
//   1. centre of mass from exact integer sums of the bit-packed mask
//   2. contour points at level 0.9: c + (0.9 - f) / (t - f) along the crossed edge, minus the centre; rho = sqrt(x^2 + y^2)
//      with every operation rounded separately like numpy, theta = atan2(y, x)
//   3. bitonic sort of (theta, index) with rho carried along (np.argsort; ties -- the closing duplicate of a loop --
//      carry identical values, so their order cannot matter)
//   4. rho smoothed with sigma = n * 5e-3 in sorted order, local maxima, the two highest with theta > 0: the one at the
//      larger theta is the left corner of the base
//   5. the contour rolled to the first point equal to the left corner and cut at the first point equal to the right
//      corner; the segment lengths are summed sequentially in contour order (Python's sum)
// Any exception of the reference (no contour, fewer than two base peaks) gives NaN, as GLS.gls's try/except does.
#include "contour_trace.cuh"

namespace cub200 {

constexpr int GL_THREADS = 128;
constexpr int GL_MAXW = CUB200_GLS_TABLE_STRIDE;                          // smoothing half-width int(4 * 0.005 n + 0.5) <= 41 for n <= 2048

struct GlsShared {
    double key[SP_NMAX];                             // theta while sorting, then the smoothed rho, then the segment lengths
    double rho[SP_NMAX];
    uint16_t idx[SP_NMAX];
    double w[GL_MAXW];
    long long red[3][GL_THREADS / 32];
    int best[2][GL_THREADS / 32];
    double bestv[2][GL_THREADS / 32];
};

// coordinate of a contour point along one axis: `h` in half-pixel units; odd = interpolated between pixels a -> b
__device__ __forceinline__ double level_coord(int h, int fa) {
    if (!(h & 1)) return (double)(h >> 1);
    const double f = (double)fa, t = 1.0 - f;        // binary mask: the two pixels differ
    return __dadd_rn((double)(h >> 1), __ddiv_rn(__dsub_rn(0.9, f), __dsub_rn(t, f)));
}

__device__ __forceinline__ void point_rc(uint32_t p, const uint32_t* __restrict__ m, double& r, double& c) {
    const int r2 = (int)(p & 0xffffu), c2 = (int)(p >> 16);
    // the pixel on the low side of the interpolated axis decides on which side of the edge the point sits
    const int pr = r2 >> 1, pc = c2 >> 1;
    const int fa = (m[pr * 8 + (pc >> 5)] >> (pc & 31)) & 1u;
    r = level_coord(r2, fa);
    c = level_coord(c2, fa);
}

__global__ void __launch_bounds__(GL_THREADS)
lv_length_kernel(const uint32_t* __restrict__ masks, const uint32_t* __restrict__ pts_g, const int32_t* __restrict__ npts,
                 double smoothing_factor, const double* __restrict__ gauss_table, double vs_r, double vs_c,
                 const int32_t* __restrict__ corners_in, double* __restrict__ length_out,
                 double* __restrict__ corners_out, int32_t* __restrict__ npts_out) {
    extern __shared__ __align__(16) unsigned char gl_smem[];
    GlsShared& S = *reinterpret_cast<GlsShared*>(gl_smem);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t g = blockIdx.x;
    const uint32_t* m = masks + g * CUB200_MASK_WORDS;
    const uint32_t* pts = pts_g + g * SP_NMAX;
    const int n = npts[g];
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    auto fail = [&]() {
        if (tid == 0) {
            length_out[g] = qnan;
            if (npts_out) npts_out[g] = 0;
            if (corners_out) for (int k = 0; k < 4; ++k) corners_out[g * 4 + k] = qnan;
        }
    };
    if (n < 3) { fail(); return; }                   // find_contours() == [] -> IndexError; overflow (-1) is not supported

    int lk, mr;                                      // roll offset (left corner) and rolled index of the right corner
    uint32_t left_p = 0u, right_p = 0u;
    if (corners_in) {
        // ---- base markers given (EchoMeasure._endo_base, endo_base_corners_kernel): np.argmin of the distance to the contour
        const int ly = corners_in[g * 4], lx = corners_in[g * 4 + 1], ry = corners_in[g * 4 + 2], rx = corners_in[g * 4 + 3];
        if (ly < 0) { fail(); return; }              // fewer than two frontier pixels: RuntimeError (measure.py:203-207)
        // np.linalg.norm(corner - contour, axis=1) = sqrt(dy * dy + dx * dx), every operation rounded; first minimum wins
        auto block_argmin = [&](double cy, double cx, int roll) {
            double bd = __longlong_as_double(0x7ff0000000000000LL);
            int bq = 0x7fffffff;
            for (int k = tid; k < n; k += GL_THREADS) {
                double r, c;
                point_rc(pts[k], m, r, c);
                const double dy = __dsub_rn(cy, r), dx = __dsub_rn(cx, c);
                const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx)));
                int q = k - roll;
                if (q < 0) q += n;
                if (d < bd || (d == bd && q < bq)) { bd = d; bq = q; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(FULL, bd, o);
                const int oq = __shfl_xor_sync(FULL, bq, o);
                if (od < bd || (od == bd && oq < bq)) { bd = od; bq = oq; }
            }
            __syncthreads();                         // the previous call's readers are done
            if (lane == 0) { S.bestv[0][wid] = bd; S.best[0][wid] = bq; }
            __syncthreads();
            bd = S.bestv[0][0]; bq = S.best[0][0];
            for (int w = 1; w < GL_THREADS / 32; ++w) {
                const double od = S.bestv[0][w];
                const int oq = S.best[0][w];
                if (od < bd || (od == bd && oq < bq)) { bd = od; bq = oq; }
            }
            return bq;
        };
        lk = block_argmin((double)ly, (double)lx, 0);
        mr = block_argmin((double)ry, (double)rx, lk);
        __syncthreads();
    } else {
    // ---- 1. centre of mass (scipy.ndimage.center_of_mass: float sums of integers / count)
    long long cnt = 0, sr = 0, sc = 0;
    for (int w = tid; w < CUB200_MASK_WORDS; w += GL_THREADS) {
        const uint32_t v = m[w];
        const int pc = __popc(v);
        const int pos = __popc(v & 0xAAAAAAAAu) + 2 * __popc(v & 0xCCCCCCCCu) + 4 * __popc(v & 0xF0F0F0F0u) +
                        8 * __popc(v & 0xFF00FF00u) + 16 * __popc(v & 0xFFFF0000u);
        cnt += pc;
        sr += (long long)(w >> 3) * pc;
        sc += (long long)(32 * (w & 7)) * pc + pos;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(FULL, cnt, o);
        sr += __shfl_xor_sync(FULL, sr, o);
        sc += __shfl_xor_sync(FULL, sc, o);
    }
    if (lane == 0) { S.red[0][wid] = cnt; S.red[1][wid] = sr; S.red[2][wid] = sc; }
    __syncthreads();
    cnt = sr = sc = 0;
#pragma unroll
    for (int w = 0; w < GL_THREADS / 32; ++w) { cnt += S.red[0][w]; sr += S.red[1][w]; sc += S.red[2][w]; }
    const double com_r = __ddiv_rn((double)sr, (double)cnt), com_c = __ddiv_rn((double)sc, (double)cnt);

    // ---- 2. polar coordinates about the centre of mass
    int P = 1;
    while (P < n) P <<= 1;
    for (int k = tid; k < P; k += GL_THREADS) {
        if (k < n) {
            double r, c;
            point_rc(pts[k], m, r, c);
            const double y = __dsub_rn(r, com_r), x = __dsub_rn(c, com_c);
            S.rho[k] = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
            S.key[k] = atan2(y, x);
            S.idx[k] = (uint16_t)k;
        } else {
            S.key[k] = __longlong_as_double(0x7ff0000000000000LL);     // +inf pads sort to the end
            S.rho[k] = 0.0;
            S.idx[k] = 0xffffu;
        }
    }
    __syncthreads();
    // ---- 3. ascending bitonic sort by (theta, index)
    for (int kk = 2; kk <= P; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += GL_THREADS) {
                const int l = i ^ j;
                if (l > i) {
                    const double ki = S.key[i], kl = S.key[l];
                    const uint16_t ii = S.idx[i], il = S.idx[l];
                    const bool gt = (ki > kl) || (ki == kl && ii > il);
                    if (gt == ((i & kk) == 0)) {
                        S.key[i] = kl; S.key[l] = ki;
                        S.idx[i] = il; S.idx[l] = ii;
                        const double ri = S.rho[i];
                        S.rho[i] = S.rho[l]; S.rho[l] = ri;
                    }
                }
            }
            __syncthreads();
        }
    }
    // first sorted position with theta > 0 (base half of the polar signal)
    int npos0 = 0;
    for (int i = tid; i < n; i += GL_THREADS) npos0 += !(S.key[i] > 0.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) npos0 += __shfl_xor_sync(FULL, npos0, o);
    if (lane == 0) S.best[0][wid] = npos0;
    // Gaussian weights (scipy _gaussian_kernel1d): exp(-0.5 / sigma^2 * j^2) / sum, the sum formed like np.sum
    const double sigma = (double)n * smoothing_factor;
    const int lw = (smoothing_factor > 0.0) ? (int)(4.0 * sigma + 0.5) : 0;
    if (lw > GL_MAXW - 1) { fail(); return; }
    if (gauss_table) {
        // normalised kernel values w[0..lw] for this n, made by the host with numpy (bit-identical to scipy's weights)
        if (tid <= lw) S.w[tid] = gauss_table[(int64_t)n * GL_MAXW + tid];
        __syncthreads();
    } else {
        if (tid <= lw) S.w[tid] = exp(__dmul_rn(__ddiv_rn(-0.5, __dmul_rn(sigma, sigma)), (double)(tid * tid)));
        __syncthreads();
        double wsum;
        {
            // numpy's pairwise sum of the 2 lw + 1 kernel values phi[-lw..lw] (n < 8: a plain loop; else 8 accumulators)
            const int nw = 2 * lw + 1;
            auto phi = [&](int q) { const int j = q - lw; return S.w[j < 0 ? -j : j]; };
            if (nw < 8) {
                wsum = 0.0;
                for (int q = 0; q < nw; ++q) wsum = __dadd_rn(wsum, phi(q));
            } else {
                double r[8];
                for (int q = 0; q < 8; ++q) r[q] = phi(q);
                int q = 8;
                for (; q < nw - (nw % 8); q += 8)
                    for (int e = 0; e < 8; ++e) r[e] = __dadd_rn(r[e], phi(q + e));
                wsum = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                for (; q < nw; ++q) wsum = __dadd_rn(wsum, phi(q));
            }
        }
        __syncthreads();
        if (tid <= lw) S.w[tid] = __ddiv_rn(S.w[tid], wsum);
        __syncthreads();
    }
    const int first_pos = S.best[0][0] + S.best[0][1] + S.best[0][2] + S.best[0][3];
    // ---- 4. smoothed rho (correlate1d, symmetric kernel, mode 'reflect') into key[]
    for (int i = tid; i < n; i += GL_THREADS) {
        double tmp = __dmul_rn(S.rho[i], S.w[0]);
        for (int jj = lw; jj >= 1; --jj) {
            int a = i - jj, b = i + jj;
            if (a < 0) a = -a - 1;
            if (b >= n) b = 2 * n - b - 1;
            tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(S.rho[a], S.rho[b]), S.w[jj]));
        }
        S.key[i] = lw ? tmp : S.rho[i];
    }
    __syncthreads();
    // local maxima (scipy _local_maxima_1d) with theta > 0; keep the two highest (height, then position)
    double h1 = -1.0, h2 = -1.0;
    int p1 = -1, p2 = -1;
    for (int i = tid; i < n - 1; i += GL_THREADS) {
        if (i == 0) continue;                        // the first and the last sample cannot be peaks
        const double v = S.key[i];
        if (!(S.key[i - 1] < v)) continue;
        int ahead = i + 1;
        while (ahead < n - 1 && S.key[ahead] == v) ++ahead;
        if (!(S.key[ahead] < v)) continue;
        const int p = (i + ahead - 1) >> 1;
        if (p < first_pos || !(v >= 0.0)) continue;
        if (v > h1 || (v == h1 && p > p1)) { h2 = h1; p2 = p1; h1 = v; p1 = p; }
        else if (v > h2 || (v == h2 && p > p2)) { h2 = v; p2 = p; }
    }
    auto merge = [&](double oh, int op) {
        if (op < 0 || op == p1 || op == p2) return;
        if (oh > h1 || (oh == h1 && op > p1)) { h2 = h1; p2 = p1; h1 = oh; p1 = op; }
        else if (oh > h2 || (oh == h2 && op > p2)) { h2 = oh; p2 = op; }
    };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double a1 = __shfl_xor_sync(FULL, h1, o), a2 = __shfl_xor_sync(FULL, h2, o);
        const int q1 = __shfl_xor_sync(FULL, p1, o), q2 = __shfl_xor_sync(FULL, p2, o);
        merge(a1, q1);
        merge(a2, q2);
    }
    if (lane == 0) { S.bestv[0][wid] = h1; S.best[0][wid] = p1; S.bestv[1][wid] = h2; S.best[1][wid] = p2; }
    __syncthreads();
    h1 = h2 = -1.0; p1 = p2 = -1;
    for (int w = 0; w < GL_THREADS / 32; ++w) { merge(S.bestv[0][w], S.best[0][w]); merge(S.bestv[1][w], S.best[1][w]); }
    if (p2 < 0) { fail(); return; }                  // fewer than two base peaks: RuntimeError (measure.py:93-97)
    // descending theta: the left corner is the peak at the larger sorted position
    const int left_i = S.idx[max(p1, p2)], right_i = S.idx[min(p1, p2)];
    left_p = pts[left_i]; right_p = pts[right_i];
    __syncthreads();                                 // everyone has read bestv / best before they are reused
    // ---- 5. np.argmin of the distance = the first point equal to the corner; roll, cut, sum
    lk = 0x7fffffff;
    for (int k = tid; k < n; k += GL_THREADS) if (pts[k] == left_p) lk = min(lk, k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lk = min(lk, __shfl_xor_sync(FULL, lk, o));
    if (lane == 0) S.best[0][wid] = lk;
    __syncthreads();
    lk = min(min(S.best[0][0], S.best[0][1]), min(S.best[0][2], S.best[0][3]));
    mr = 0x7fffffff;
    for (int k = tid; k < n; k += GL_THREADS)
        if (pts[k] == right_p) { const int q = k - lk; mr = min(mr, q < 0 ? q + n : q); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mr = min(mr, __shfl_xor_sync(FULL, mr, o));
    if (lane == 0) S.best[1][wid] = mr;
    __syncthreads();
    mr = min(min(S.best[1][0], S.best[1][1]), min(S.best[1][2], S.best[1][3]));
    }
    // segment q joins rolled points q and q + 1: np.linalg.norm of a 2-vector = sqrt(x . x)
    for (int q = tid; q < mr; q += GL_THREADS) {
        int k0 = q + lk, k1 = q + 1 + lk;
        if (k0 >= n) k0 -= n;
        if (k1 >= n) k1 -= n;
        double r0, c0, r1, c1;
        point_rc(pts[k0], m, r0, c0);
        point_rc(pts[k1], m, r1, c1);
        const double dr = __dmul_rn(__dsub_rn(r1, r0), vs_r), dc = __dmul_rn(__dsub_rn(c1, c0), vs_c);   // (p1 - p0) * voxelspacing
        S.key[q] = __dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dc, dc)));
    }
    __syncthreads();
    if (tid == 0) {
        double total = 0.0;
        for (int q = 0; q < mr; ++q) total = __dadd_rn(total, S.key[q]);
        length_out[g] = total;
        if (npts_out) npts_out[g] = mr + 1;
        if (corners_out && corners_in) {
            for (int k = 0; k < 4; ++k) corners_out[g * 4 + k] = (double)corners_in[g * 4 + k];
        } else if (corners_out) {
            double r, c;
            point_rc(left_p, m, r, c);
            corners_out[g * 4] = r; corners_out[g * 4 + 1] = c;
            point_rc(right_p, m, r, c);
            corners_out[g * 4 + 2] = r; corners_out[g * 4 + 3] = c;
        }
    }
}

// ---------------------------------------------------------------- EchoMeasure._endo_base          measure.py:177-222
// One CTA per (LV mask, MYO mask) pair.  Frontier = LV & dilate8(MYO) & dilate8(~(LV | MYO)) (scipy.ndimage.binary_dilation,
// 3 x 3 structure, zeros beyond the frame), formed word by word in the bit-packed rows; its pixel count and coordinate sums
// are exact integers, so `x < x.mean()` is decided as x n < sum(x).  np.argmax / np.argmin over np.nonzero's raster order
// = extreme row, then the smallest column: one atomicMax of a packed key per half.
// corners_out (n, 4) int32 = (left row, left col, right row, right col); -1 when the reference raises (< 2 frontier pixels).
__device__ __forceinline__ uint32_t dilate_row_word(const uint32_t* __restrict__ m, int r, int w) {
    if (r < 0 || r > 255) return 0u;
    const uint32_t* row = m + r * 8;
    const uint32_t c = row[w], lo = w > 0 ? row[w - 1] : 0u, hi = w < 7 ? row[w + 1] : 0u;
    return c | (c << 1) | (lo >> 31) | (c >> 1) | (hi << 31);
}

__global__ void __launch_bounds__(GL_THREADS)
endo_base_corners_kernel(const uint32_t* __restrict__ lv_masks, const uint32_t* __restrict__ myo_masks,
                         int32_t* __restrict__ corners_out) {
    __shared__ uint32_t F[CUB200_MASK_WORDS];
    __shared__ unsigned long long sums[3];           // count, sum of columns, sum of rows
    __shared__ int ext[4];                           // min col, max col, best left key, best right key
    const int tid = threadIdx.x;
    const int64_t g = blockIdx.x;
    const uint32_t* lv = lv_masks + g * CUB200_MASK_WORDS;
    const uint32_t* my = myo_masks + g * CUB200_MASK_WORDS;
    if (tid < 3) sums[tid] = 0ull;
    if (tid == 0) { ext[0] = 256; ext[1] = -1; ext[2] = -1; ext[3] = -1; }
    __syncthreads();
    unsigned long long cnt = 0, sx = 0, sy = 0;
    int xmin = 256, xmax = -1;
    for (int i = tid; i < CUB200_MASK_WORDS; i += GL_THREADS) {
        const int r = i >> 3, w = i & 7;
        uint32_t f = lv[i];
        if (f) {
            uint32_t dm = 0u, dother = 0u;
#pragma unroll
            for (int dr = -1; dr <= 1; ++dr) {
                dm |= dilate_row_word(my, r + dr, w);
                // others = ~(LV | MYO) inside the frame: dilate the complement, rows beyond the frame contribute nothing
                const int rr = r + dr;
                if (rr >= 0 && rr <= 255) {
                    const uint32_t* a = lv + rr * 8;
                    const uint32_t* b = my + rr * 8;
                    const uint32_t c = ~(a[w] | b[w]);
                    const uint32_t lo = w > 0 ? ~(a[w - 1] | b[w - 1]) : 0u, hi = w < 7 ? ~(a[w + 1] | b[w + 1]) : 0u;
                    dother |= c | (c << 1) | (lo >> 31) | (c >> 1) | (hi << 31);
                }
            }
            f &= dm & dother;
        }
        F[i] = f;
        uint32_t v = f;
        while (v) {
            const int b = __ffs(v) - 1;
            v &= v - 1;
            const int x = 32 * w + b;
            ++cnt; sx += x; sy += r;
            xmin = min(xmin, x); xmax = max(xmax, x);
        }
    }
    if (cnt) {
        atomicAdd(&sums[0], cnt); atomicAdd(&sums[1], sx); atomicAdd(&sums[2], sy);
        atomicMin(&ext[0], xmin); atomicMax(&ext[1], xmax);
    }
    __syncthreads();
    const long long n = (long long)sums[0], SX = (long long)sums[1], SY = (long long)sums[2];
    if (n < 2) {
        if (tid < 4) corners_out[g * 4 + tid] = -1;
        return;
    }
    const bool one_column = ext[0] == ext[1];
    int kl = -1, kr = -1;
    for (int i = tid; i < CUB200_MASK_WORDS; i += GL_THREADS) {
        uint32_t v = F[i];
        const int r = i >> 3;
        while (v) {
            const int b = __ffs(v) - 1;
            v &= v - 1;
            const int x = 32 * (i & 7) + b;
            if (one_column) {
                // coord_mask = y > y.mean(): left = the highest pixel below the mean row, right = the lowest at or above it
                if ((long long)r * n > SY) kl = max(kl, ((255 - r) << 8) | (255 - x));
                else kr = max(kr, (r << 8) | (255 - x));
            } else {
                // coord_mask = x < x.mean(): the lowest pixel of each half, raster-first among equals
                if ((long long)x * n < SX) kl = max(kl, (r << 8) | (255 - x));
                else kr = max(kr, (r << 8) | (255 - x));
            }
        }
    }
    if (kl >= 0) atomicMax(&ext[2], kl);
    if (kr >= 0) atomicMax(&ext[3], kr);
    __syncthreads();
    if (tid == 0) {
        const int a = ext[2], b = ext[3];
        corners_out[g * 4 + 0] = one_column ? 255 - (a >> 8) : (a >> 8);
        corners_out[g * 4 + 1] = 255 - (a & 255);
        corners_out[g * 4 + 2] = b >> 8;
        corners_out[g * 4 + 3] = 255 - (b & 255);
    }
}

}  // namespace cub200

extern "C" int cub200_endo_base_corners(const uint32_t* lv_masks, const uint32_t* myo_masks, int64_t n_masks, int32_t* corners_out,
                                        void* stream) {
    using namespace cub200;
    if (n_masks == 0) return 0;
    if (!lv_masks || !myo_masks || !corners_out || n_masks < 0 || n_masks > 0x7fffffff) return CUB200_EINVAL;
    endo_base_corners_kernel<<<(unsigned)n_masks, GL_THREADS, 0, (cudaStream_t)stream>>>(lv_masks, myo_masks, corners_out);
    return check_launch("endo_base_corners_kernel");
}

extern "C" int cub200_mask_lv_length_myo(const uint32_t* lv_masks, const uint32_t* myo_masks, int64_t n_masks, double spacing_row,
                                         double spacing_col, double* length_out, int32_t* corners_out, int32_t* npts_out,
                                         void* workspace, void* stream) {
    using namespace cub200;
    if (n_masks == 0) return 0;
    if (!lv_masks || !myo_masks || !length_out || !corners_out || !workspace || n_masks < 0 || n_masks > 0x7fffffff) return CUB200_EINVAL;
    if ((reinterpret_cast<uintptr_t>(lv_masks) | reinterpret_cast<uintptr_t>(workspace)) & 15u) return CUB200_EINVAL;   // the tracer's 16-byte accesses
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* pts = reinterpret_cast<uint32_t*>(workspace);
    int32_t* npts = reinterpret_cast<int32_t*>(pts + n_masks * SP_NMAX);
    endo_base_corners_kernel<<<(unsigned)n_masks, GL_THREADS, 0, st>>>(lv_masks, myo_masks, corners_out);
    if (int rc = check_launch("endo_base_corners_kernel")) return rc;
    contour_trace_kernel<<<(unsigned)((n_masks + 127) / 128), 128, 0, st>>>(lv_masks, n_masks, nullptr, nullptr, 256, 256, pts, npts);
    if (int rc = check_launch("contour_trace_kernel")) return rc;
    CUB200_CUDA_OK(cudaFuncSetAttribute(lv_length_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GlsShared)));
    lv_length_kernel<<<(unsigned)n_masks, GL_THREADS, sizeof(GlsShared), st>>>(lv_masks, pts, npts, 0.0, nullptr, spacing_row, spacing_col,
                                                                              corners_out, length_out, nullptr, npts_out);
    return check_launch("lv_length_kernel");
}

extern "C" int cub200_mask_lv_length(const uint32_t* masks, int64_t n_masks, double smoothing_factor, const double* gauss_table,
                                     double spacing_row, double spacing_col, double* length_out, double* corners_out,
                                     int32_t* npts_out, void* workspace, void* stream) {
    using namespace cub200;
    if (n_masks == 0) return 0;
    if (!masks || !length_out || !workspace || n_masks < 0 || n_masks > 0x7fffffff || !(smoothing_factor >= 0.0) ||
        smoothing_factor > 5e-3)
        return CUB200_EINVAL;
    if ((reinterpret_cast<uintptr_t>(masks) | reinterpret_cast<uintptr_t>(workspace)) & 15u) return CUB200_EINVAL;   // the tracer's 16-byte accesses
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* pts = reinterpret_cast<uint32_t*>(workspace);
    int32_t* npts = reinterpret_cast<int32_t*>(pts + n_masks * SP_NMAX);
    contour_trace_kernel<<<(unsigned)((n_masks + 127) / 128), 128, 0, st>>>(masks, n_masks, nullptr, nullptr, 256, 256, pts, npts);
    if (int rc = check_launch("contour_trace_kernel")) return rc;
    CUB200_CUDA_OK(cudaFuncSetAttribute(lv_length_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GlsShared)));
    lv_length_kernel<<<(unsigned)n_masks, GL_THREADS, sizeof(GlsShared), st>>>(masks, pts, npts, smoothing_factor, gauss_table,
                                                                              spacing_row, spacing_col, nullptr, length_out, corners_out,
                                                                              npts_out);
    return check_launch("lv_length_kernel");
}

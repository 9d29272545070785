// Analytic (sampling-free) uncertainty maps and the projected per-point uncertainties.
//
// Replaces (reference paths, cu/ = contour_uncertainty/):
//   projected_uncertainty / projected_uncertainty_value   cu/utils/uncertainty_projection.py:11-125
//   BivariateNormal.marginal, BivariateSkewNormal.marginal, rotate_cov, rotate_alpha, cov2corr
//                                                         cu/distributions/bivariatenormal.py:67-86,
//                                                         bivariateskewnormal.py:92-135, utils.py:37-72,132-149
//   uncertainty_map                                       cu/utils/umap.py:10-33
//   USUMap.__call__ (LV labels)                           cu/data/camus/utils.py:101-124
//   skew_umap                                             cu/utils/skew_umap.py:11-82
//   USSkewUmap.__call__ (LV labels)                       cu/data/camus/utils.py:127-153
//
// The reference forms rotate_cov / rotate_alpha in fp32 torch: its projected variances carry fp32 rounding.
// Here the exact fp64 value is formed and rounded to fp32 once (oracle/port.py does the same), which agrees
// with the reference to 1 fp32 ulp.
#include "common.cuh"
#include "spline.cuh"
#include "spline_generic.cuh"

namespace cub200 {

constexpr int KP = CUB200_K;
constexpr int UM_WARPS = 4;

// derivative of the fitted spline at parameter x (splev(.., der=1), extrapolating with the end pieces)
__device__ __forceinline__ void spline_der(const SplineScratch& sp, double x, double& dx, double& dy) {
    int p = 0;
#pragma unroll
    for (int k = 1; k <= KP - 2; ++k) p = (x >= sp.u[k]) ? k : p;
    const double* c = sp.piece[p];
    const double s = x - c[1];
    dx = fma(fma(3.0 * c[5], s, 2.0 * c[4]), s, c[3]);
    dy = fma(fma(3.0 * c[9], s, 2.0 * c[8]), s, c[7]);
}

// ---------------------------------------------------------------- U1: projected uncertainty, one warp per image
// Parameter of the node of np.linspace(0, 1.01, 1000) nearest to u (first minimum of |u - unew|)   uncertainty_projection.py:32,47
__device__ __forceinline__ double nearest_node_param(double u) {
    const double step = 1.01 / 999.0;
    int j = (int)floor(u / step);
    j = max(0, min(j - 1, 999));
    double bestd = 1e300; int bi = 0;
    for (int k = j; k <= min(j + 3, 999); ++k) {
        const double un = (k == 999) ? 1.01 : (double)k * step;
        const double d = fabs(u - un);
        if (d < bestd) { bestd = d; bi = k; }
    }
    return (bi == 999) ? 1.01 : (double)bi * step;
}

// first, second, last-but-one and last point of the contour (the straight-line closure of linear_close)
struct EndPoints { double X0, Y0, X1, Y1, XM, YM, XL, YL; };

// One landmark `pt` of a contour of `npts` points given the spline derivative (dx, dy) at its node: projection vector,
// projected standard deviation (or the eigenvalue sum at the base / apex points), projected skewness.
// o = flat index of the landmark in unc / proj / alpha_proj / cov / alpha.               uncertainty_projection.py:46-123
__device__ __forceinline__ void project_point(int pt, int npts, int64_t o, double dx, double dy, const EndPoints& e,
                                              const float* __restrict__ cov, const float* __restrict__ alpha, int all_points,
                                              int linear_close, float* __restrict__ unc, double* __restrict__ proj,
                                              float* __restrict__ alpha_proj) {
    const double nrm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    // v = flip(der / |der|), v[1] = -v[1]
    double vx = dy / nrm, vy = -(dx / nrm);
    const double c = vx, s = vy;                             // (cos, sin) of angle = atan2(v_y, v_x)
    const float* cv = cov + o * 4;
    const double xx = cv[0], xy = cv[1], yx = cv[2], yy = cv[3];
    float out_u;
    if (!all_points && (pt == 0 || pt == npts / 2 || pt == npts - 1)) {
        // sum of sqrt(eigenvalues) of the 2x2 covariance
        const double tr = xx + yy, det = xx * yy - xy * yx;
        const double disc = sqrt(fmax(tr * tr * 0.25 - det, 0.0));
        out_u = (float)(sqrt(tr * 0.5 + disc) + sqrt(tr * 0.5 - disc));
        if (alpha_proj) alpha_proj[o] = __int_as_float(0x7fc00000);
    } else {
        if (linear_close && (pt == 0 || pt == npts - 1)) {
            // the reference does this on its fp32 numpy `mu`: fp32 arithmetic throughout (uncertainty_projection.py:55-68)
            const float x0 = (float)e.X0, y0 = (float)e.Y0, x1 = (float)e.X1, y1 = (float)e.Y1;
            const float xl = (float)e.XL, yl = (float)e.YL, xm = (float)e.XM, ym = (float)e.YM;
            float ax, ay, bx, by;
            if (pt == 0) { ax = __fsub_rn(x1, x0); ay = __fsub_rn(y1, y0); bx = __fsub_rn(xl, x0); by = __fsub_rn(yl, y0); }
            else { ax = __fsub_rn(xl, xm); ay = __fsub_rn(yl, ym); bx = __fsub_rn(xl, x0); by = __fsub_rn(yl, y0); }
            const float na = sqrtf(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)));
            const float nb = sqrtf(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)));
            ax = __fdiv_rn(ax, na); ay = __fdiv_rn(ay, na); bx = __fdiv_rn(bx, nb); by = __fdiv_rn(by, nb);
            const float wx = __fdiv_rn(__fadd_rn(ax, bx), 2.0f), wy = __fdiv_rn(__fadd_rn(ay, by), 2.0f);
            const float nw = sqrtf(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)));
            vx = (double)__fdiv_rn(wx, nw); vy = (double)__fdiv_rn(wy, nw);
        }
        // rotate_cov(cov, -angle) = R cov R^T, R = [[c, s], [-s, c]]
        const float r_xx = (float)(c * c * xx + c * s * (xy + yx) + s * s * yy);
        out_u = sqrtf(r_xx);
        if (alpha) {
            const float r_yy = (float)(s * s * xx - c * s * (xy + yx) + c * c * yy);
            const float r_xy = (float)(-c * s * xx + c * c * xy - s * s * yx + c * s * yy);
            const double a0 = alpha[2 * o], a1 = -(double)alpha[2 * o + 1];           // alpha[1] = -alpha[1]
            const float al0 = (float)(c * a0 + s * a1), al1 = (float)(-s * a0 + c * a1);   // rotate_alpha(alpha, -angle)
            const float std0 = sqrtf(r_xx), std1 = sqrtf(r_yy);
            const float corr_12 = __fdiv_rn(r_xy, __fmul_rn(std0, std1));
            const float corr_11 = __fdiv_rn(r_xx, __fmul_rn(std0, std0));
            const float corr_22 = __fdiv_rn(r_yy, __fmul_rn(std1, std1));
            const float corr_22_1 = __fsub_rn(corr_22, __fdiv_rn(__fmul_rn(corr_12, corr_12), corr_11));
            const float num = __fadd_rn(al0, __fmul_rn(__fmul_rn(__fdiv_rn(1.0f, corr_11), corr_12), al1));
            const float den = sqrtf(__fadd_rn(1.0f, __fmul_rn(__fmul_rn(al1, corr_22_1), al1)));
            alpha_proj[o] = __fdiv_rn(num, den);
        }
    }
    unc[o] = out_u;
    proj[2 * o] = vx;
    proj[2 * o + 1] = vy;
}

// unc (n,21) fp32, proj (n,21,2) fp64, alpha_proj (n,21) fp32 (only with alpha)
__global__ void __launch_bounds__(UM_WARPS * 32)
projected_uncertainty_kernel(const float* __restrict__ mu, const float* __restrict__ cov, const float* __restrict__ alpha,
                             int64_t n_img, int all_points, int linear_close, float* __restrict__ unc,
                             double* __restrict__ proj, float* __restrict__ alpha_proj) {
    __shared__ SplineScratch s_sp[UM_WARPS];
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    const int64_t img = blockIdx.x * (int64_t)UM_WARPS + wid;
    if (img >= n_img) return;
    SplineScratch& sp = s_sp[wid];
    double X = 0.0, Y = 0.0;
    if (lane < KP) { X = mu[(img * KP + lane) * 2]; Y = mu[(img * KP + lane) * 2 + 1]; }
    const bool bad = fit_spline(sp, X, Y, lane);            // splprep raises in the reference: outputs NaN
    EndPoints e;
    e.X0 = __shfl_sync(FULL, X, 0); e.Y0 = __shfl_sync(FULL, Y, 0);
    e.X1 = __shfl_sync(FULL, X, 1); e.Y1 = __shfl_sync(FULL, Y, 1);
    e.XL = __shfl_sync(FULL, X, KP - 1); e.YL = __shfl_sync(FULL, Y, KP - 1);
    e.XM = __shfl_sync(FULL, X, KP - 2); e.YM = __shfl_sync(FULL, Y, KP - 2);
    if (lane >= KP) return;
    const int64_t o = img * KP + lane;
    if (bad) {
        const float qn = __int_as_float(0x7fc00000);
        unc[o] = qn; proj[2 * o] = qn; proj[2 * o + 1] = qn;
        if (alpha_proj) alpha_proj[o] = qn;
        return;
    }
    double dx, dy;
    spline_der(sp, nearest_node_param(sp.u[lane]), dx, dy);
    project_point(lane, KP, o, dx, dy, e, cov, alpha, all_points, linear_close, unc, proj, alpha_proj);
}

// The same for contours of any size (4 <= npts <= 64: the lung / heart contours of LungUMap, cu/data/lung/utils.py:72-114):
// one warp per contour, spline_generic.cuh, lanes stride over the landmarks.
__global__ void __launch_bounds__(UM_WARPS * 32)
projected_uncertainty_generic_kernel(const float* __restrict__ mu, const float* __restrict__ cov, const float* __restrict__ alpha,
                                     int64_t n_img, int npts, int all_points, int linear_close, float* __restrict__ unc,
                                     double* __restrict__ proj, float* __restrict__ alpha_proj) {
    extern __shared__ __align__(16) unsigned char um_smem[];
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    const int64_t img = blockIdx.x * (int64_t)UM_WARPS + wid;
    if (img >= n_img) return;
    GenericSpline& S = reinterpret_cast<GenericSpline*>(um_smem)[wid];
    for (int k = lane; k < npts; k += 32) { S.X[k] = mu[(img * npts + k) * 2]; S.Y[k] = mu[(img * npts + k) * 2 + 1]; }
    __syncwarp();
    const bool bad = fit_spline_generic(S, npts, lane);
    const EndPoints e = {S.X[0], S.Y[0], S.X[1], S.Y[1], S.X[npts - 2], S.Y[npts - 2], S.X[npts - 1], S.Y[npts - 1]};
    for (int k = lane; k < npts; k += 32) {
        const int64_t o = img * npts + k;
        if (bad) {
            const float qn = __int_as_float(0x7fc00000);
            unc[o] = qn; proj[2 * o] = qn; proj[2 * o + 1] = qn;
            if (alpha_proj) alpha_proj[o] = qn;
            continue;
        }
        double dx, dy;
        spline_der_generic(S, npts, nearest_node_param(S.u[k]), dx, dy);
        project_point(k, npts, o, dx, dy, e, cov, alpha, all_points, linear_close, unc, proj, alpha_proj);
    }
}

// ---------------------------------------------------------------- U2: uncertainty_map, one warp per (image, offset)
// kmap (n,256,256) int32, zeroed by the caller: 1 + index of the LAST offset that touched the pixel
__global__ void __launch_bounds__(UM_WARPS * 32)
umap_offsets_kernel(const float* __restrict__ mu, const float* __restrict__ unc, const double* __restrict__ proj,
                    int64_t n_img, int32_t* __restrict__ kmap) {
    __shared__ SplineScratch s_sp[UM_WARPS];
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    const int64_t g = blockIdx.x * (int64_t)UM_WARPS + wid;
    if (g >= n_img * 100) return;
    const int64_t img = g / 100;
    const int k = (int)(g % 100);
    SplineScratch& sp = s_sp[wid];
    int32_t* km = kmap + img * 65536;
    // i = np.linspace(-2, 2, 100)[k];  m = mu + v * u * i
    const double off = (k == 99) ? 2.0 : __dadd_rn(__dmul_rn((double)k, 4.0 / 99.0), -2.0);
    double X = 0.0, Y = 0.0;
    if (lane < KP) {
        const int64_t o = img * KP + lane;
        const double u = unc[o];
        X = __dadd_rn((double)mu[2 * o], __dmul_rn(__dmul_rn(proj[2 * o], u), off));
        Y = __dadd_rn((double)mu[2 * o + 1], __dmul_rn(__dmul_rn(proj[2 * o + 1], u), off));
    }
    const bool bad = fit_spline(sp, X, Y, lane);
    if (!bad) {
        // contour_spline(m, n=1001): u = j / 1000, round half-even, clip to [0, 255]
        for (int j = lane; j < 1001; j += 32) {
            const double uj = (j == 1000) ? 1.0 : (double)j * 0.001;
            int p = 0;
#pragma unroll
            for (int q = 1; q <= KP - 2; ++q) p = (uj >= sp.u[q]) ? q : p;
            const double* c = sp.piece[p];
            const double sv = uj - c[1];
            const double xv = fma(fma(fma(c[5], sv, c[4]), sv, c[3]), sv, c[2]);
            const double yv = fma(fma(fma(c[9], sv, c[8]), sv, c[7]), sv, c[6]);
            const int ix = min(max(__double2int_rn(xv), 0), 255), iy = min(max(__double2int_rn(yv), 0), 255);
            atomicMax(km + iy * 256 + ix, k + 1);
        }
    } else if (lane < KP) {                                   // bare except: the raw points
        const int ix = min(max(__double2int_rn(X), 0), 255), iy = min(max(__double2int_rn(Y), 0), 255);
        atomicMax(km + iy * 256 + ix, k + 1);
    }
    // closing line between the TRUNCATED end points (m.astype(int)), clipped to the frame (umap.py:25-31)
    const int r0 = (int)__shfl_sync(FULL, Y, KP - 1), c0 = (int)__shfl_sync(FULL, X, KP - 1);
    const int r1 = (int)__shfl_sync(FULL, Y, 0), c1 = (int)__shfl_sync(FULL, X, 0);
    const int dr = abs(r1 - r0), dc = abs(c1 - c0);
    const int sc = (c1 - c0) > 0 ? 1 : -1, sr = (r1 - r0) > 0 ? 1 : -1;
    const bool steep = dr > dc;
    const int dmaj = steep ? dr : dc, dmin = steep ? dc : dr;
    for (int i = lane; i <= dmaj; i += 32) {
        int rr, cc;
        if (i == dmaj) { rr = r1; cc = c1; }
        else {
            const int nmin = (2 * dmin * i + dmaj) / (2 * dmaj);
            if (steep) { rr = r0 + i * sr; cc = c0 + nmin * sc; }
            else { cc = c0 + i * sc; rr = r0 + nmin * sr; }
        }
        rr = min(max(rr, 0), 255); cc = min(max(cc, 0), 255);
        atomicMax(km + rr * 256 + cc, k + 1);
    }
}

// U2b: values + normalisation by the map maximum, one CTA per image.  table[k] = norm.pdf(linspace(-2,2,100)[k])
__global__ void __launch_bounds__(256)
umap_finalize_kernel(const int32_t* __restrict__ kmap, const double* __restrict__ table, int normalize, double* __restrict__ out) {
    __shared__ double s_max[8];
    __shared__ double s_tab[100];
    const int64_t img = blockIdx.x;
    const int32_t* km = kmap + img * 65536;
    if (threadIdx.x < 100) s_tab[threadIdx.x] = table[threadIdx.x];
    __syncthreads();
    double mx = 0.0;
    for (int i = threadIdx.x; i < 65536; i += 256) { const int k = km[i]; if (k > 0) mx = fmax(mx, s_tab[k - 1]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = s_max[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmax(mx, s_max[w]);
    double* o = out + img * 65536;
    for (int i = threadIdx.x; i < 65536; i += 256) {
        const int k = km[i];
        const double v = k > 0 ? s_tab[k - 1] : 0.0;
        o[i] = normalize ? v / mx : v;
    }
}

// ---------------------------------------------------------------- U3: skew_umap level contours, one warp per (image, point)
// contours (n, 200, npts, 2) fp64, mode (n, npts, 2) fp32
__global__ void __launch_bounds__(UM_WARPS * 32)
skew_umap_contours_kernel(const float* __restrict__ mu, const float* __restrict__ unc, const double* __restrict__ proj,
                          const float* __restrict__ alpha_proj, int64_t n_img, int npts, double* __restrict__ contours,
                          float* __restrict__ mode) {
    __shared__ double s_y[UM_WARPS][1000];
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    const int64_t g = blockIdx.x * (int64_t)UM_WARPS + wid;
    if (g >= n_img * npts) return;
    const int64_t img = g / npts;
    const int pt = (int)(g % npts);
    double* y = s_y[wid];
    const float uf = unc[g];
    const double u = uf, a = alpha_proj[g];
    const double vx = proj[2 * g], vy = proj[2 * g + 1];
    const double mx = mu[2 * g], my = mu[2 * g + 1];
    // p1 = mu + v u cov_width, p2 = mu - v u cov_width   (cov_width = 2)
    const double ox = __dmul_rn(__dmul_rn(vx, u), 2.0), oy = __dmul_rn(__dmul_rn(vy, u), 2.0);
    const double p1x = __dadd_rn(mx, ox), p1y = __dadd_rn(my, oy), p2x = __dsub_rn(mx, ox), p2y = __dsub_rn(my, oy);
    // x = np.linspace(-3u, 3u, 1000) is a FLOAT32 array in the reference (u is a float32 numpy scalar): step = fl(fl(stop -
    // start) / 999), x_j = fl(fl(j step) + start), x_999 = stop.  scipy then forms (x - 0) / u in fp64.  The fp32 rounding of
    // x decides the arg-max between the two centre samples of a symmetric law (alpha = 0: LungUMap), hence every level index.
    const float start = __fmul_rn(-3.0f, uf), stop = __fmul_rn(3.0f, uf);
    const float step = __fdiv_rn(__fsub_rn(stop, start), 999.0f);
    double ymax = -1.0; int am = 0x7fffffff;
    for (int j = lane; j < 1000; j += 32) {
        const double x = (double)((j == 999) ? stop : __fadd_rn(__fmul_rn((float)j, step), start));
        const double t = x / u;
        const double pdf = exp(-(t * t) / 2.0) / 2.5066282746310002;
        const double v = (a == 0.0 ? pdf : 2.0 * pdf * normcdf(a * t)) / u;
        y[j] = v;
        if (v > ymax) { ymax = v; am = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v = __shfl_xor_sync(FULL, ymax, o);
        const int i = __shfl_xor_sync(FULL, am, o);
        if (v > ymax || (v == ymax && i < am)) { ymax = v; am = i; }
    }
    __syncwarp();
    for (int j = lane; j < 1000; j += 32) y[j] = y[j] / ymax;            // y = y / y.max()  (mode_y == 1.0)
    __syncwarp();
    if (lane == 0) {
        const double t = (double)am / 1000.0, ti = 1.0 - t;
        mode[2 * g] = (float)__dadd_rn(__dmul_rn(p1x, t), __dmul_rn(ti, p2x));
        mode[2 * g + 1] = (float)__dadd_rn(__dmul_rn(p1y, t), __dmul_rn(ti, p2y));
    }
    // 100 levels: nearest y on either side of the mode (first minimum of |y - val|)
    const int nr = 999 - am, nl = am;                         // right = y[am+1:], left = y[:am]
    for (int i = 0; i < 100; ++i) {
        const double lv = (i == 99) ? 0.95 : (double)i * (0.95 / 99.0);   // np.linspace(0, 0.95, 100)
        const double val = 1.0 - lv;
        double bd = 1e300; int bi = 0x7fffffff;
        for (int j = lane; j < nr; j += 32) { const double d = fabs(y[am + 1 + j] - val); if (d < bd) { bd = d; bi = j; } }
        double cd = 1e300; int ci = 0x7fffffff;
        for (int j = lane; j < nl; j += 32) { const double d = fabs(y[j] - val); if (d < cd) { cd = d; ci = j; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double d1 = __shfl_xor_sync(FULL, bd, o); const int i1 = __shfl_xor_sync(FULL, bi, o);
            if (d1 < bd || (d1 == bd && i1 < bi)) { bd = d1; bi = i1; }
            const double d2 = __shfl_xor_sync(FULL, cd, o); const int i2 = __shfl_xor_sync(FULL, ci, o);
            if (d2 < cd || (d2 == cd && i2 < ci)) { cd = d2; ci = i2; }
        }
        if (lane == 0) {
            // an empty side makes np.argmin raise in the reference: NaN here
            const double qn = __longlong_as_double(0x7ff8000000000000LL);
            const double tp = nr > 0 ? (double)(bi + am) / 1000.0 : qn, tm = nl > 0 ? (double)ci / 1000.0 : qn;
            double* cm = contours + ((img * 200 + (99 - i)) * npts + pt) * 2;      // umap_contours[N - i - 1]
            double* cp = contours + ((img * 200 + (100 + i)) * npts + pt) * 2;     // umap_contours[N + i]
            cm[0] = __dadd_rn(__dmul_rn(p1x, tm), __dmul_rn(1.0 - tm, p2x));
            cm[1] = __dadd_rn(__dmul_rn(p1y, tm), __dmul_rn(1.0 - tm, p2y));
            cp[0] = __dadd_rn(__dmul_rn(p1x, tp), __dmul_rn(1.0 - tp, p2x));
            cp[1] = __dadd_rn(__dmul_rn(p1y, tp), __dmul_rn(1.0 - tp, p2y));
        }
    }
}

// ---------------------------------------------------------------- U5: weighted mean of masks -> natural-log binary entropy
// masks (n, n_masks, 2048) bit-packed; weights (n_masks); out (n,256,256) fp64; one CTA per image
__global__ void __launch_bounds__(256)
weighted_entropy_kernel(const uint32_t* __restrict__ masks, const double* __restrict__ weights, int n_masks, double wsum,
                        int normalize, double* __restrict__ out) {
    __shared__ double s_max[8];
    extern __shared__ double s_w[];
    const int64_t img = blockIdx.x;
    for (int k = threadIdx.x; k < n_masks; k += 256) s_w[k] = weights[k];
    __syncthreads();
    const uint32_t* mk = masks + img * (int64_t)n_masks * CUB200_MASK_WORDS;
    double* o = out + img * 65536;
    double mx = 0.0;
    // thread t owns word (t + 256 i): 32 consecutive pixels; masks are summed in order k = 0..n-1 (np.average)
    for (int w = threadIdx.x; w < CUB200_MASK_WORDS; w += 256) {
        double acc[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) acc[b] = 0.0;
        for (int k = 0; k < n_masks; ++k) {
            const uint32_t bits = mk[(int64_t)k * CUB200_MASK_WORDS + w];
            const double wk = s_w[k];
#pragma unroll
            for (int b = 0; b < 32; ++b) acc[b] = __dadd_rn(acc[b], ((bits >> b) & 1u) ? wk : 0.0);
        }
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const double m = acc[b] / wsum, m1 = 1.0 - m;
            const double s = __dadd_rn(m, m1);                 // scipy.stats.entropy normalises pk
            const double pa = m / s, pb = m1 / s;
            // scipy.special.entr: -x log x, 0 at 0, -inf below 0 (m can exceed 1 by an ulp)
            const double ninf = __longlong_as_double(0xfff0000000000000LL);
            const double h = __dadd_rn(pa > 0.0 ? -pa * log(pa) : (pa == 0.0 ? 0.0 : ninf), pb > 0.0 ? -pb * log(pb) : (pb == 0.0 ? 0.0 : ninf));
            acc[b] = h;
            mx = fmax(mx, h);
        }
#pragma unroll
        for (int b = 0; b < 32; ++b) o[w * 32 + b] = acc[b];
    }
    if (!normalize) return;
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, of));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = s_max[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmax(mx, s_max[w]);
    for (int w = threadIdx.x; w < CUB200_MASK_WORDS; w += 256)
#pragma unroll
        for (int b = 0; b < 32; ++b) o[w * 32 + b] = o[w * 32 + b] / mx;
}

// ---------------------------------------------------------------- contour perimeter (spline arc length), one warp per contour
// perimeter() / global_longitudinal_strain() of cu/utils/clinical.py:30-71: contour_spline(contour) (n = 1001 points,
// unrounded landmarks) and the sum of its 1000 chord lengths; the 21 raw points when FITPACK rejects the input.
__global__ void __launch_bounds__(UM_WARPS * 32)
contour_perimeter_kernel(const float* __restrict__ contours, int64_t n, double* __restrict__ out) {
    __shared__ SplineScratch s_sp[UM_WARPS];
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    const int64_t g = blockIdx.x * (int64_t)UM_WARPS + wid;
    if (g >= n) return;
    SplineScratch& sp = s_sp[wid];
    double X = 0.0, Y = 0.0;
    if (lane < KP) { X = contours[(g * KP + lane) * 2]; Y = contours[(g * KP + lane) * 2 + 1]; }
    const bool bad = fit_spline(sp, X, Y, lane);
    double len = 0.0;
    if (!bad) {
        // lane L owns segments 32 L .. 32 L + 31 (of 1000): points 32 L .. 32 L + 32
        const int j0 = 32 * lane;
        double px = 0.0, py = 0.0;
        for (int jj = 0; jj <= 32; ++jj) {
            const int j = j0 + jj;
            if (j > 1000) break;
            const double uj = (j == 1000) ? 1.0 : (double)j * 0.001;
            int p = 0;
#pragma unroll
            for (int q = 1; q <= KP - 2; ++q) p = (uj >= sp.u[q]) ? q : p;
            const double* c = sp.piece[p];
            const double sv = uj - c[1];
            const double xv = fma(fma(fma(c[5], sv, c[4]), sv, c[3]), sv, c[2]);
            const double yv = fma(fma(fma(c[9], sv, c[8]), sv, c[7]), sv, c[6]);
            if (jj > 0) len += sqrt((xv - px) * (xv - px) + (yv - py) * (yv - py));
            px = xv; py = yv;
        }
    } else {
        const double Xn = __shfl_down_sync(FULL, X, 1), Yn = __shfl_down_sync(FULL, Y, 1);
        if (lane < KP - 1) len = sqrt((Xn - X) * (Xn - X) + (Yn - Y) * (Yn - Y));
    }
    len = warp_sum(len);
    if (lane == 0) out[g] = len;
}

}  // namespace cub200

extern "C" int cub200_contour_perimeter(const float* contours, int64_t n, double* perimeter_out, void* stream) {
    using namespace cub200;
    if (n == 0) return 0;
    if (!contours || !perimeter_out || n < 0) return CUB200_EINVAL;
    contour_perimeter_kernel<<<(unsigned)((n + UM_WARPS - 1) / UM_WARPS), UM_WARPS * 32, 0, (cudaStream_t)stream>>>(contours, n, perimeter_out);
    return check_launch("contour_perimeter_kernel");
}

extern "C" int cub200_projected_uncertainty(const float* mu, const float* cov, const float* alpha, int64_t n_img,
                                            int all_points, int linear_close, float* unc_out, double* proj_out,
                                            float* alpha_proj_out, void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!mu || !cov || !unc_out || !proj_out || n_img < 0 || (alpha && !alpha_proj_out)) return CUB200_EINVAL;
    projected_uncertainty_kernel<<<(unsigned)((n_img + UM_WARPS - 1) / UM_WARPS), UM_WARPS * 32, 0, (cudaStream_t)stream>>>(
        mu, cov, alpha, n_img, all_points, linear_close, unc_out, proj_out, alpha ? alpha_proj_out : nullptr);
    return check_launch("projected_uncertainty_kernel");
}

extern "C" int cub200_umap_analytic(const float* mu, const float* unc, const double* proj, int64_t n_img,
                                    const double* pdf_table, int normalize, int32_t* kmap_ws, double* umap_out, void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!mu || !unc || !proj || !pdf_table || !kmap_ws || !umap_out || n_img < 0 || n_img > 0x7fffffff / 100) return CUB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    CUB200_CUDA_OK(cudaMemsetAsync(kmap_ws, 0, sizeof(int32_t) * 65536 * n_img, st));
    umap_offsets_kernel<<<(unsigned)((n_img * 100 + UM_WARPS - 1) / UM_WARPS), UM_WARPS * 32, 0, st>>>(mu, unc, proj, n_img, kmap_ws);
    int rc = check_launch("umap_offsets_kernel");
    if (rc) return rc;
    umap_finalize_kernel<<<(unsigned)n_img, 256, 0, st>>>(kmap_ws, pdf_table, normalize, umap_out);
    return check_launch("umap_finalize_kernel");
}

extern "C" int cub200_projected_uncertainty_generic(const float* mu, const float* cov, const float* alpha, int64_t n_img,
                                                    int n_points, int all_points, int linear_close, float* unc_out,
                                                    double* proj_out, float* alpha_proj_out, void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!mu || !cov || !unc_out || !proj_out || n_img < 0 || (alpha && !alpha_proj_out) || n_points < 4 || n_points > GS_MAX)
        return CUB200_EINVAL;
    const int smem = UM_WARPS * (int)sizeof(GenericSpline);
    CUB200_CUDA_OK(cudaFuncSetAttribute(projected_uncertainty_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    projected_uncertainty_generic_kernel<<<(unsigned)((n_img + UM_WARPS - 1) / UM_WARPS), UM_WARPS * 32, smem, (cudaStream_t)stream>>>(
        mu, cov, alpha, n_img, n_points, all_points, linear_close, unc_out, proj_out, alpha ? alpha_proj_out : nullptr);
    return check_launch("projected_uncertainty_generic_kernel");
}

extern "C" int cub200_skew_umap_contours(const float* mu, const float* unc, const double* proj, const float* alpha_proj,
                                         int64_t n_img, int n_points, double* contours_out, float* mode_out, void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!mu || !unc || !proj || !alpha_proj || !contours_out || !mode_out || n_img < 0 || n_points < 1 || n_points > GS_MAX)
        return CUB200_EINVAL;
    const int64_t n = n_img * n_points;
    skew_umap_contours_kernel<<<(unsigned)((n + UM_WARPS - 1) / UM_WARPS), UM_WARPS * 32, 0, (cudaStream_t)stream>>>(
        mu, unc, proj, alpha_proj, n_img, n_points, contours_out, mode_out);
    return check_launch("skew_umap_contours_kernel");
}

extern "C" int cub200_weighted_mask_entropy(const uint32_t* masks, const double* weights, int64_t n_img, int n_masks,
                                            double weight_sum, int normalize, double* out, void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!masks || !weights || !out || n_img < 0 || n_masks < 1 || n_masks > 4096 || !(weight_sum > 0.0)) return CUB200_EINVAL;
    weighted_entropy_kernel<<<(unsigned)n_img, 256, sizeof(double) * n_masks, (cudaStream_t)stream>>>(
        masks, weights, n_masks, weight_sum, normalize, out);
    return check_launch("weighted_entropy_kernel");
}

// Interpolating cubic spline of FITPACK splprep(k=3, s=0) through 21 points, one warp, fp64.
//
// Replaces contour_spline's fit (contour_uncertainty/utils/contour.py:9-13) and the splprep calls of
// contour_uncertainty/utils/uncertainty_projection.py:31.  Chord-length parameter (parcur), knots at
// u_2..u_{m-3} == the not-a-knot cubic spline through (u_i, p_i): solved as a tridiagonal system for the second
// derivatives by parallel cyclic reduction in registers and stored as 20 polynomial pieces.
#pragma once
#include <climits>
#include <type_traits>
#include "common.cuh"

namespace cub200 {

// 1 / x to ~1 ulp: hardware seed (20 bits) + two Newton steps.  The IEEE division sequence costs ~3x as many issue slots
// and the solve is not bit-identical to FITPACK's anyway (different algorithm, agreement ~1e-13).
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

struct __align__(16) SplineScratch {
    double u[24];          // chord parameters u_0..u_20
    double piece[20][10];  // piece i on [u_i, u_{i+1}]: {u_{i+1}, u_i, x c0..c3, y c0..c3} (powers of u - u_i)
};

// Layout of the rasteriser's hot loop (fused.cu): the knots are replaced by the INDEX of the first of the 1000 evaluation
// points that belongs to a later piece, so the per-point piece test is an integer comparison.
struct __align__(16) SplineScratchIdx {
    double piece[20][10];  // {unused, u_i, x c0..c3, y c0..c3}
    int jnext[32];         // jnext[p] = first evaluation point of a piece after p (INT_MAX for the last piece, 19)
    int pnext[32];         // pnext[p] = the next piece after p that holds at least one evaluation point
};

constexpr int SPLINE_NPTS = 1000;   // contour_spline(points, n=1000)

// Smallest j >= 0 with (double)j * (1/999) >= u: the first of the 1000 evaluation points np.linspace(0, 1, 1000) at or
// after the knot u (the products are formed exactly as the evaluation loop forms them, so the piece a point falls into is
// the one a comparison of the parameters picks).
__device__ __forceinline__ int first_point_at_or_after(double u) {
    const double step = 1.0 / (double)(SPLINE_NPTS - 1);
    int j = max((int)ceil(u * (double)(SPLINE_NPTS - 1)), 0);
    if (j > 0 && (double)(j - 1) * step >= u) --j;
    if ((double)j * step < u) ++j;
    return j;
}

// Lane k < 21 passes point k as (X, Y).  Returns true when FITPACK would raise (zero chord / non-increasing
// parameter): the reference's bare except then falls back to the raw points.  On success sp is filled.
// `wide` (optional): set when some piece may leave (-2^30, 2^30) on its interval (|c0| + h(|c1| + h(|c2| + h |c3|)), NaN
// included) -- then `value + 1.5 * 2^52` no longer holds rint(value) in its low word and the caller must convert with F2I.
template <class Scratch>
__device__ __forceinline__ bool fit_spline(Scratch& sp, const double X, const double Y, const int lane, bool* wide = nullptr) {
    constexpr int KP = CUB200_K;
    constexpr bool INDEXED = std::is_same<Scratch, SplineScratchIdx>::value;
    // 3a. chord-length parameter (FITPACK parcur)
    const double Xp = __shfl_up_sync(FULL, X, 1), Yp = __shfl_up_sync(FULL, Y, 1);
    double dch = 0.0;
    if (lane >= 1 && lane < KP) dch = sqrt((X - Xp) * (X - Xp) + (Y - Yp) * (Y - Yp));
    double cum = dch;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(FULL, cum, o);
        if (lane >= o) cum += t;
    }
    const double total = __shfl_sync(FULL, cum, KP - 1);
    double u = (lane == KP - 1) ? 1.0 : cum * fast_rcp(total);
    if (lane == 0) u = 0.0;
    const double uprev = __shfl_up_sync(FULL, u, 1);
    // FITPACK rejects non-increasing parameters (zero chord) -> ValueError -> raw points drawn
    const bool bad = __any_sync(FULL, lane >= 1 && lane < KP && (dch == 0.0 || !(u > uprev))) || !(total > 0.0);
    __syncwarp();

    if (bad) return true;
    // 3b. not-a-knot cubic spline: tridiagonal system for the second derivatives M_1..M_19
    const double unext = __shfl_down_sync(FULL, u, 1);
    const double Xn = __shfl_down_sync(FULL, X, 1), Yn = __shfl_down_sync(FULL, Y, 1);
    const double h = unext - u;                         // h_k, k < 20
    const double ih = fast_rcp(h);
    const double dxk = (Xn - X) * ih, dyk = (Yn - Y) * ih;  // divided differences delta_k
    const double hm = __shfl_up_sync(FULL, h, 1);
    const double dxm = __shfl_up_sync(FULL, dxk, 1), dym = __shfl_up_sync(FULL, dyk, 1);
    int jnext = 0;
    if constexpr (INDEXED) {
        // evaluation points per piece: piece k holds the points j with jstart_k <= j < jstart_{k+1}; the last piece keeps the rest
        const int jstart = (lane == 0) ? 0 : first_point_at_or_after(u);
        jnext = __shfl_down_sync(FULL, jstart, 1);
        if (lane >= KP - 2) jnext = INT_MAX;
        sp.jnext[lane] = (lane < KP - 1) ? jnext : INT_MAX;   // padded: the binary search of the rasteriser reads all 32 entries
        // pieces without an evaluation point (two knots inside one step of 1/999) are skipped in ONE move
        const unsigned holds = __ballot_sync(FULL, lane < KP - 1 && jnext > jstart);
        sp.pnext[lane] = lane + __ffs((holds >> lane) >> 1);   // the last piece always holds point 999
    } else {
        if (lane < KP) sp.u[lane] = u;
    }
    // rows 1..19 (lane = row), normalised to unit diagonal; other lanes hold identity rows
    double ta = 0.0, tc = 0.0, tx = 0.0, ty = 0.0;
    if (lane >= 1 && lane <= KP - 2) {
        double a = hm, b = 2.0 * (hm + h), c = h;
        double r1 = 6.0 * (dxk - dxm), r2 = 6.0 * (dyk - dym);
        if (lane == 1) {                                 // M_0 eliminated (not-a-knot at u_1)
            const double f = h * fast_rcp(hm + h);
            a = 0.0; b = hm + 2.0 * h; c = h - hm; r1 *= f; r2 *= f;
        } else if (lane == KP - 2) {                     // M_20 eliminated (not-a-knot at u_19)
            const double f = hm * fast_rcp(hm + h);
            a = hm - h; b = h + 2.0 * hm; c = 0.0; r1 *= f; r2 *= f;
        }
        const double ib = fast_rcp(b);
        ta = a * ib; tc = c * ib; tx = r1 * ib; ty = r2 * ib;
    }
    // parallel cyclic reduction: 5 steps decouple the 19 equations (x and y share the matrix)
    #pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) {
        double am = __shfl_up_sync(FULL, ta, sft), cm = __shfl_up_sync(FULL, tc, sft);
        double xm = __shfl_up_sync(FULL, tx, sft), ym = __shfl_up_sync(FULL, ty, sft);
        double ap = __shfl_down_sync(FULL, ta, sft), cp = __shfl_down_sync(FULL, tc, sft);
        double xp = __shfl_down_sync(FULL, tx, sft), yp = __shfl_down_sync(FULL, ty, sft);
        // A shuffle from beyond the warp returns the lane's own (finite) values; no masking is needed: the coupling of row i
        // to a row that does not exist is exactly 0 (ta == 0 for lane < sft, tc == 0 for lane + sft > 31: rows 0 and 20..31
        // are identity rows, a_1 = c_19 = 0, and ta' = -ta am r, tc' = -tc cp r carry the zeros outwards step by step).
        const double r = fast_rcp(1.0 - ta * cm - tc * ap);
        tx = (tx - ta * xm - tc * xp) * r;
        ty = (ty - ta * ym - tc * yp) * r;
        ta = -ta * am * r;
        tc = -tc * cp * r;
    }
    // second derivatives: lane i holds M_i (i = 1..19); M_0, M_20 from the not-a-knot relations
    double Mx = tx, My = ty;
    {
        const double m1x = __shfl_down_sync(FULL, Mx, 1), m2x = __shfl_down_sync(FULL, Mx, 2);
        const double m1y = __shfl_down_sync(FULL, My, 1), m2y = __shfl_down_sync(FULL, My, 2);
        const double h1 = __shfl_down_sync(FULL, h, 1);
        const double p1x = __shfl_up_sync(FULL, Mx, 1), p2x = __shfl_up_sync(FULL, Mx, 2);
        const double p1y = __shfl_up_sync(FULL, My, 1), p2y = __shfl_up_sync(FULL, My, 2);
        const double hm2 = __shfl_up_sync(FULL, h, 2);
        if (lane == 0) { const double q = h * fast_rcp(h1); Mx = m1x - q * (m2x - m1x); My = m1y - q * (m2y - m1y); }
        if (lane == KP - 1) { const double q = hm * fast_rcp(hm2); Mx = p1x + q * (p1x - p2x); My = p1y + q * (p1y - p2y); }
    }
    {
        const double m1x = __shfl_down_sync(FULL, Mx, 1), m1y = __shfl_down_sync(FULL, My, 1);
        if (lane < KP - 1) {                             // piece coefficients
            double2* c = reinterpret_cast<double2*>(sp.piece[lane]);
            c[0] = make_double2(unext, u);
            c[1] = make_double2(X, dxk - h * (2.0 * Mx + m1x) * (1.0 / 6.0));
            c[2] = make_double2(0.5 * Mx, (m1x - Mx) * ih * (1.0 / 6.0));
            c[3] = make_double2(Y, dyk - h * (2.0 * My + m1y) * (1.0 / 6.0));
            c[4] = make_double2(0.5 * My, (m1y - My) * ih * (1.0 / 6.0));
        }
        if (wide) {
            const double bx = fabs(X) + h * (fabs(dxk - h * (2.0 * Mx + m1x) * (1.0 / 6.0)) + h * (fabs(0.5 * Mx) + h * fabs((m1x - Mx) * ih * (1.0 / 6.0))));
            const double by = fabs(Y) + h * (fabs(dyk - h * (2.0 * My + m1y) * (1.0 / 6.0)) + h * (fabs(0.5 * My) + h * fabs((m1y - My) * ih * (1.0 / 6.0))));
            *wide = __any_sync(FULL, lane < KP - 1 && !(fmax(bx, by) < 1073741824.0));
        }
    }
    __syncwarp();

    return false;
}

}  // namespace cub200

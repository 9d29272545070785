// Interpolating cubic spline of FITPACK splprep(k=3, s=0) through 4..64 points, one warp, fp64: the general-size companion of
// spline.cuh (which is specialised to the 21 LV landmarks and solves by parallel cyclic reduction in registers).
//
// Replaces contour_spline's fit (contour_uncertainty/utils/contour.py:9-13) and the splprep call of
// contour_uncertainty/utils/uncertainty_projection.py:31 for contours of any size (lung / heart contours of 44, 50, 26 points,
// contour_uncertainty/data/lung/utils.py).  Chord-length parameter (parcur), knots at u_2..u_{m-3} == the not-a-knot cubic
// spline through (u_i, p_i): the tridiagonal system for the second derivatives is solved by lane 0 (Thomas algorithm, <= 62
// unknowns, x and y together), the polynomial pieces are formed by all lanes.
#pragma once
#include "common.cuh"

namespace cub200 {

constexpr int GS_MAX = CUB200_RASTER_MAX_POINTS;     // 64

struct __align__(16) GenericSpline {
    double piece[GS_MAX][10];      // piece i on [u_i, u_{i+1}]: {u_{i+1}, u_i, x c0..c3, y c0..c3} (powers of u - u_i)
    double u[GS_MAX];
    double X[GS_MAX], Y[GS_MAX];   // the points the spline goes through (filled by the caller)
    double Mx[GS_MAX], My[GS_MAX]; // second derivatives
    double cp[GS_MAX];             // Thomas: modified super-diagonal
};

// S.X / S.Y[0..K) hold the points (visible to the whole warp).  Returns true when FITPACK would raise (zero chord /
// non-increasing parameter); on success S.u and S.piece[0..K-2] are filled and visible to the whole warp.
__device__ __forceinline__ bool fit_spline_generic(GenericSpline& S, const int K, const int lane) {
    bool bad;
    {
        // chord-length parameter (FITPACK parcur) and the not-a-knot system, by lane 0
        int badflag = 0;
        if (lane == 0) {
            double cum = 0.0;
            S.u[0] = 0.0;
            for (int k = 1; k < K; ++k) {
                const double dx = S.X[k] - S.X[k - 1], dy = S.Y[k] - S.Y[k - 1];
                const double d = sqrt(dx * dx + dy * dy);
                if (d == 0.0) badflag = 1;
                cum += d;
                S.u[k] = cum;
            }
            if (!(cum > 0.0)) badflag = 1;
            if (!badflag) {
                for (int k = 1; k < K - 1; ++k) {
                    S.u[k] = S.u[k] / cum;
                    if (!(S.u[k] > S.u[k - 1])) badflag = 1;
                }
                S.u[K - 1] = 1.0;
                if (!(S.u[K - 1] > S.u[K - 2])) badflag = 1;
            }
            if (!badflag) {
                // rows i = 1..K-2 of  h_{i-1} M_{i-1} + 2 (h_{i-1} + h_i) M_i + h_i M_{i+1} = 6 (delta_i - delta_{i-1}),
                // with M_0 and M_{K-1} eliminated through the not-a-knot conditions at u_1 and u_{K-2}
                const int last = K - 2;
                double prev_cp = 0.0, prev_dx = 0.0, prev_dy = 0.0;
                for (int i = 1; i <= last; ++i) {
                    const double hm = S.u[i] - S.u[i - 1], h = S.u[i + 1] - S.u[i];
                    const double dxk = (S.X[i + 1] - S.X[i]) / h, dyk = (S.Y[i + 1] - S.Y[i]) / h;
                    const double dxm = (S.X[i] - S.X[i - 1]) / hm, dym = (S.Y[i] - S.Y[i - 1]) / hm;
                    double a = hm, b = 2.0 * (hm + h), c = h;
                    double r1 = 6.0 * (dxk - dxm), r2 = 6.0 * (dyk - dym);
                    if (i == 1) {
                        const double f = h / (hm + h);
                        a = 0.0; b = hm + 2.0 * h; c = h - hm; r1 *= f; r2 *= f;
                    }
                    if (i == last) {
                        const double f = hm / (hm + h);
                        a = hm - h; b = h + 2.0 * hm; c = 0.0; r1 *= f; r2 *= f;
                    }
                    const double den = b - a * prev_cp;
                    prev_cp = c / den;
                    prev_dx = (r1 - a * prev_dx) / den;
                    prev_dy = (r2 - a * prev_dy) / den;
                    S.cp[i] = prev_cp; S.Mx[i] = prev_dx; S.My[i] = prev_dy;
                }
                for (int i = last - 1; i >= 1; --i) {
                    S.Mx[i] -= S.cp[i] * S.Mx[i + 1];
                    S.My[i] -= S.cp[i] * S.My[i + 1];
                }
                {
                    const double h0 = S.u[1] - S.u[0], h1 = S.u[2] - S.u[1];
                    const double q = h0 / h1;
                    S.Mx[0] = S.Mx[1] - q * (S.Mx[2] - S.Mx[1]);
                    S.My[0] = S.My[1] - q * (S.My[2] - S.My[1]);
                    const double hl = S.u[K - 1] - S.u[K - 2], hl2 = S.u[K - 2] - S.u[K - 3];
                    const double ql = hl / hl2;
                    S.Mx[K - 1] = S.Mx[K - 2] + ql * (S.Mx[K - 2] - S.Mx[K - 3]);
                    S.My[K - 1] = S.My[K - 2] + ql * (S.My[K - 2] - S.My[K - 3]);
                }
            }
        }
        bad = __shfl_sync(FULL, badflag, 0) != 0;
        __syncwarp();
        if (!bad) {
            for (int i = lane; i < K - 1; i += 32) {
                const double h = S.u[i + 1] - S.u[i], ih = 1.0 / h;
                const double dxk = (S.X[i + 1] - S.X[i]) * ih, dyk = (S.Y[i + 1] - S.Y[i]) * ih;
                double* c = S.piece[i];
                c[0] = S.u[i + 1]; c[1] = S.u[i];
                c[2] = S.X[i]; c[3] = dxk - h * (2.0 * S.Mx[i] + S.Mx[i + 1]) * (1.0 / 6.0);
                c[4] = 0.5 * S.Mx[i]; c[5] = (S.Mx[i + 1] - S.Mx[i]) * ih * (1.0 / 6.0);
                c[6] = S.Y[i]; c[7] = dyk - h * (2.0 * S.My[i] + S.My[i + 1]) * (1.0 / 6.0);
                c[8] = 0.5 * S.My[i]; c[9] = (S.My[i + 1] - S.My[i]) * ih * (1.0 / 6.0);
            }
            __syncwarp();
        }
    }
    return bad;
}

// derivative of the fitted spline at parameter x (splev(.., der=1), extrapolating with the end pieces)
__device__ __forceinline__ void spline_der_generic(const GenericSpline& S, const int K, double x, double& dx, double& dy) {
    int p = 0;
    for (int k = 1; k <= K - 2; ++k) p = (x >= S.u[k]) ? k : p;
    const double* c = S.piece[p];
    const double s = x - c[1];
    dx = fma(fma(3.0 * c[5], s, 2.0 * c[4]), s, c[3]);
    dy = fma(fma(3.0 * c[9], s, 2.0 * c[8]), s, c[7]);
}

}  // namespace cub200

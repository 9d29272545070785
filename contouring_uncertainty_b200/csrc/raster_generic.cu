// reconstruction() / linear_reconstruction() for contours with any number of points (4..64): the rasteriser of fused.cu is
// specialised to the 21 LV landmarks of the Monte-Carlo hot path; this one serves the functions' full domain -- e.g. the
// 44 / 50 / 26-point lung and heart contours of LungContourToMask.
//
// Replaces (reference paths, cu/ = contour_uncertainty/):
//   contour_spline, reconstruction, linear_reconstruction      cu/utils/contour.py:9-53
//   LungContourToMask.__call__ (per structure)                 cu/data/lung/utils.py:25-69
//
// One warp per contour.  FITPACK splprep(k=3, s=0) == the not-a-knot cubic spline through the points over the chord-
// length parameter: the tridiagonal system for the second derivatives is solved by lane 0 (Thomas algorithm, <= 62
// unknowns, x and y together), the polynomial pieces are formed by all lanes, then exactly the steps of fused.cu:
// 1000 evaluations at j / 999 -> round half even -> index rules (clip at 255, negative wrap, < -256 drops the whole
// spline, zero chord -> the raw points) -> closing line -> binary_fill_holes (flood.cuh) -> bit-packed mask + area.
#include "flood.cuh"
#include "spline_generic.cuh"

namespace cub200 {

constexpr int GK_MAX = CUB200_RASTER_MAX_POINTS;     // 64
constexpr int GK_WARPS = 4;
constexpr int GK_NPTS = 1000;
constexpr int GK_CLAMP = 8192;

struct __align__(16) GenericScratch {
    uint32_t bmp[BMP_WORDS];
    GenericSpline sp;              // the points the spline goes through (rounded or not), parameters, pieces
    int rix[GK_MAX], riy[GK_MAX];  // rounded points (closing line, linear reconstruction)
};

template <typename T>
__global__ void __launch_bounds__(GK_WARPS * 32)
raster_generic_kernel(const T* __restrict__ contours, int64_t n, int K, int flags, uint32_t* __restrict__ masks_out,
                      int32_t* __restrict__ area_out) {
    extern __shared__ __align__(16) unsigned char gk_smem[];
    GenericScratch& S = reinterpret_cast<GenericScratch*>(gk_smem)[threadIdx.x >> 5];
    const int lane = lane_id();
    const int64_t g = blockIdx.x * (int64_t)GK_WARPS + (threadIdx.x >> 5);
    if (g >= n) return;
    const bool round_landmarks = (flags & 1) != 0, linear = (flags & 2) != 0;
    uint32_t* bmp = S.bmp;
    {
        uint4* b4 = reinterpret_cast<uint4*>(bmp);
#pragma unroll
        for (int i = 0; i < BMP_WORDS / 4 / 32; ++i) b4[lane + 32 * i] = make_uint4(0, 0, 0, 0);
    }
    for (int k = lane; k < K; k += 32) {
        double fx = (double)contours[(g * K + k) * 2], fy = (double)contours[(g * K + k) * 2 + 1];
        fx = fmin(fmax(fx, -(double)GK_CLAMP), (double)GK_CLAMP);
        fy = fmin(fmax(fy, -(double)GK_CLAMP), (double)GK_CLAMP);
        const int rx = __double2int_rn(fx), ry = __double2int_rn(fy);
        S.rix[k] = rx; S.riy[k] = ry;
        S.sp.X[k] = round_landmarks ? (double)rx : fx;
        S.sp.Y[k] = round_landmarks ? (double)ry : fy;
    }
    __syncwarp();

    bool bad = false, oob = false;
    if (!linear) {
        bad = fit_spline_generic(S.sp, K, lane);               // spline_generic.cuh
        if (!bad) {
            // 1000 points, 32 consecutive per lane (lane 31 re-draws 24 points of lane 30)
            const int j0 = min(lane * 32, GK_NPTS - 32);
            const double step = 1.0 / (double)(GK_NPTS - 1);
            int piece = 0;
            {
                const double us = (double)j0 * step;
                for (int k = 1; k <= K - 2; ++k) piece = (us >= S.sp.u[k]) ? k : piece;
            }
            int lowest = 0;
            for (int jj = 0; jj < 32; ++jj) {
                const double uj = (double)(j0 + jj) * step;
                while (piece < K - 2 && uj >= S.sp.piece[piece][0]) ++piece;
                const double* c = S.sp.piece[piece];
                const double sv = uj - c[1];
                const double xv = fma(fma(fma(c[5], sv, c[4]), sv, c[3]), sv, c[2]);
                const double yv = fma(fma(fma(c[9], sv, c[8]), sv, c[7]), sv, c[6]);
                const int ix = min(__double2int_rn(xv), 255), iy = min(__double2int_rn(yv), 255);
                lowest = min(lowest, min(ix, iy));
                // [-256,-1] wraps like a numpy negative index
                atomicOr(&bmp[bmp_addr(iy & 255, (ix & 255) >> 5)], 1u << (ix & 31));
            }
            oob = __any_sync(FULL, lowest < -256);
            if (oob) {                                       // IndexError in the reference: nothing of the spline is drawn
                __syncwarp();
                uint4* b4 = reinterpret_cast<uint4*>(bmp);
                for (int i = 0; i < BMP_WORDS / 4 / 32; ++i) b4[lane + 32 * i] = make_uint4(0, 0, 0, 0);
            }
        } else {
            // fallback of contour_spline(): the raw points
            bool o = false;
            for (int k = lane; k < K; k += 32) {
                const int ix = min(__double2int_rn(S.sp.X[k]), 255), iy = min(__double2int_rn(S.sp.Y[k]), 255);
                o |= (ix < -256) | (iy < -256);
            }
            oob = __any_sync(FULL, o);
            if (!oob)
                for (int k = lane; k < K; k += 32) {
                    const int ix = min(__double2int_rn(S.sp.X[k]), 255), iy = min(__double2int_rn(S.sp.Y[k]), 255);
                    atomicOr(&bmp[bmp_addr(iy & 255, (ix & 255) >> 5)], 1u << (ix & 31));
                }
        }
        __syncwarp();
    }
    // closing line last point -> first point; linear_reconstruction: every edge k -> k + 1
    for (int e = linear ? 0 : K - 1; e < K; ++e) {
        const int e1 = (e + 1 == K) ? 0 : e + 1;
        draw_line_clipped(bmp, S.riy[e], S.rix[e], S.riy[e1], S.rix[e1], lane);
    }
    __syncwarp();
    uint32_t R[8][8];
    const uint4* band4 = reinterpret_cast<const uint4*>(bmp + lane * BAND_STRIDE);
    flood_background(R, band4, lane);
    int area = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int w = 0; w < 8; ++w) area += __popc(~R[i][w]);
        if (masks_out) {
            uint4* o = reinterpret_cast<uint4*>(masks_out + g * CUB200_MASK_WORDS + (lane * 8 + i) * 8);
            o[0] = make_uint4(~R[i][0], ~R[i][1], ~R[i][2], ~R[i][3]);
            o[1] = make_uint4(~R[i][4], ~R[i][5], ~R[i][6], ~R[i][7]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(FULL, area, o);
    if (area_out && lane == 0) area_out[g] = area;
}

template <typename T>
static int launch_generic(const T* contours, int64_t n, int n_points, int flags, uint32_t* masks_out, int32_t* area_out,
                          cudaStream_t st) {
    const int smem = GK_WARPS * (int)sizeof(GenericScratch);
    CUB200_CUDA_OK(cudaFuncSetAttribute(raster_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    raster_generic_kernel<T><<<(unsigned)((n + GK_WARPS - 1) / GK_WARPS), GK_WARPS * 32, smem, st>>>(contours, n, n_points, flags,
                                                                                                   masks_out, area_out);
    return check_launch("raster_generic_kernel");
}

}  // namespace cub200

extern "C" int cub200_rasterize_contours_generic(const void* contours, int is_f64, int64_t n, int n_points, int flags,
                                                 uint32_t* masks_out, int32_t* area_out, void* stream) {
    using namespace cub200;
    if (n == 0) return 0;
    if (!contours || (!masks_out && !area_out) || n < 0 || n > 0x7ffffff0 || n_points < 4 || n_points > GK_MAX || (flags & ~3))
        return CUB200_EINVAL;
    if (masks_out && (reinterpret_cast<uintptr_t>(masks_out) & 15)) return CUB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    return is_f64 ? launch_generic(reinterpret_cast<const double*>(contours), n, n_points, flags, masks_out, area_out, st)
                  : launch_generic(reinterpret_cast<const float*>(contours), n, n_points, flags, masks_out, area_out, st);
}

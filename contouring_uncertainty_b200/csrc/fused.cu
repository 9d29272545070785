// K3  sample_raster_aggregate: Monte-Carlo contour sampler -> spline rasteriser -> background
// flood -> streaming aggregation, ONE WARP PER SAMPLE, masks never leave shared memory.
//
// Replaces (reference paths, cu/ = contour_uncertainty/):
//   AleatoricUncertaintyTask.sample            cu/task/regression/aleatoric.py:54-78
//   ContourUncertaintyTask.convert_to_mask     cu/task/regression/contour_uncertainty.py:26-57
//   USContourToMask.__call__ (LV branch)       cu/data/camus/utils.py:31-98
//   reconstruction / contour_spline            cu/utils/contour.py:9-40
//   lv_area                                    cu/utils/clinical.py:92-93
//
// The mask rule is NOT point-in-polygon.  Per sample:
//   1. contour = mu + A z (z replayed from memory or Philox4x32-10 + Box-Muller), lane k = landmark k
//   2. landmarks -> round half-even -> int                         (camus/utils.py:43)
//   3. interpolating cubic B-spline of FITPACK splprep(k=3, s=0): chord-length parameter, knots at
//      u_2..u_{m-3} == the not-a-knot cubic spline through (u_i, p_i); solved in fp64 as a
//      tridiagonal system for the second derivatives and evaluated piecewise (Horner) at
//      u = j/999, j = 0..999; zero chord -> FITPACK raises -> the reference's bare except draws
//      the 21 raw points instead                                   (contour.py:9-25)
//   4. np.round half-even -> int; index clip(max=255) only: [-256,-1] wraps, < -256 raises and
//      NOTHING of the spline is drawn                              (contour.py:30-35)
//   5. Bresenham closing line landmark 20 -> 0 (skimage.draw.line), clipped to the frame (:36-38)
//   6. scipy.ndimage.binary_fill_holes == complement of the 4-connected background reachable from
//      outside the frame                                           (:39)
//   7. area = popcount; every run boundary of every mask row becomes a +1 / -1 red.global into the
//      image's difference map (the row prefix sum of which is the per-pixel count).
//
// Shared memory per warp: the 256x256 boundary bitmap (8 rows per lane, padded band stride 68
// words so the 128-bit band loads are bank-conflict free) + ~2 KB of spline scratch.
#include "common.cuh"
#include "flood.cuh"
#include "spline.cuh"

namespace cub200 {

#ifndef CUB200_FW_WARPS
#define CUB200_FW_WARPS 8
#endif
#ifndef CUB200_FW_MINBLOCKS
#define CUB200_FW_MINBLOCKS 2
#endif
constexpr int FW_WARPS = CUB200_FW_WARPS;
constexpr int FW_MINBLOCKS = CUB200_FW_MINBLOCKS;
constexpr int FW_THREADS = FW_WARPS * 32;
constexpr int NPTS = 1000;                      // contour_spline(points, n=1000)   contour.py:30
constexpr int KP = CUB200_K;                    // 21
constexpr int COORD_CLAMP = 8192;               // landmarks beyond +-8192 px behave as +-8192

struct __align__(16) WarpScratch {
    uint32_t bmp[BMP_WORDS];
    SplineScratch sp;
    float z[CUB200_ZMAX + 2];
};

// ---------------------------------------------------------------- bitmap helpers
__device__ __forceinline__ void set_pixel(uint32_t* bmp, int row, int col) {
    atomicOr(&bmp[bmp_addr(row, col >> 5)], 1u << (col & 31));
}

// Generic run extraction of one mask row (mask = ~R): every transition of R along the row (virtual
// R = 1 left of column 0) is +1 where the mask starts, -1 one past where it ends.  Returns the row's
// mask area.  Rare path (frame-border rows, rows of concave / self-intersecting samples).
__device__ __noinline__ int emit_row_generic(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t r4,
                                             uint32_t r5, uint32_t r6, uint32_t r7, int32_t* drow) {
    const uint32_t Rw[8] = {r0, r1, r2, r3, r4, r5, r6, r7};
    int area = (r7 >> 31) ? 0 : 256;                         // a run reaching column 255 ends at 256
    uint32_t prev = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t tw = Rw[w] ^ __funnelshift_l(prev, Rw[w], 1);
        prev = Rw[w];
        while (tw) {
            const int b = __ffs(tw) - 1;
            tw &= tw - 1;
            const int x = 32 * w + b;
            const bool start = !((Rw[w] >> b) & 1u);
            area += start ? -x : x;
            if (drow) atomicAdd(drow + x, start ? 1 : -1);
        }
    }
    return area;
}

// Take back the single run [first, last boundary pixel] that step 6a emitted for this row (rare).
__device__ __noinline__ int undo_row_span(const uint32_t* brow, int32_t* drow) {
    int xs = -1, xe = 0;
    for (int w = 0; w < 8; ++w) {
        const uint32_t b = brow[w];
        if (b) {
            if (xs < 0) xs = 32 * w + __ffs(b) - 1;
            xe = 32 * w + 32 - __clz(b);
        }
    }
    if (xs < 0) return 0;
    if (drow) {
        atomicAdd(drow + xs, -1);
        if (xe < 256) atomicAdd(drow + xe, 1);
    }
    return xe - xs;
}

// ---------------------------------------------------------------- one sample, one warp
// Lane k < 21 holds landmark k as (fx, fy).  Returns the mask area; accumulates run boundaries
// into diff_img (may be null) and stores the packed mask (may be null).
// flags: bit 0 = round the landmarks before the spline (USContourToMask), bit 1 = linear_reconstruction
// (cu/utils/contour.py:43-53): no spline, the 21 rounded landmarks are joined by skimage.draw.line segments.
__device__ int raster_sample(WarpScratch& ws, double fx, double fy, int flags,
                             int32_t* __restrict__ diff_img, uint32_t* __restrict__ mask_out) {
    const int lane = lane_id();
    const bool round_landmarks = (flags & 1) != 0, linear = (flags & 2) != 0;
    uint32_t* bmp = ws.bmp;

    // zero the boundary bitmap (uint4 stores)
    {
        uint4* b4 = reinterpret_cast<uint4*>(bmp);
        const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < BMP_WORDS / 4 / 32; ++i) b4[lane + 32 * i] = zero;
    }
    __syncwarp();

    // 2. rounded landmarks (closing line always uses them; the spline too when round_landmarks)
    fx = fmin(fmax(fx, -(double)COORD_CLAMP), (double)COORD_CLAMP);
    fy = fmin(fmax(fy, -(double)COORD_CLAMP), (double)COORD_CLAMP);
    const int rix = __double2int_rn(fx), riy = __double2int_rn(fy);
    const double X = round_landmarks ? (double)rix : fx;
    const double Y = round_landmarks ? (double)riy : fy;

    // 3a/3b. chord-length parameter + not-a-knot cubic spline (spline.cuh)
    const bool bad = linear ? false : fit_spline(ws.sp, X, Y, lane);

    bool oob = false;
    if (linear) {
        // polygon outline only: every edge is drawn by step 5
    } else if (!bad) {
        // 3c/4. evaluate 1000 points (32 consecutive per lane), round half-even, apply the index
        //       rules and set the pixels on the fly (consecutive points falling into the same
        //       32-pixel word are merged into one shared-memory atomic).  An index < -256 makes the
        //       reference's whole assignment raise: then the bitmap is simply cleared again (the
        //       closing line is drawn afterwards).
        const int j0 = min(lane * 32, NPTS - 32);           // lane 31 re-draws 24 points of lane 30
        const double step = 1.0 / (double)(NPTS - 1);       // j * step == np.linspace(0, 1, 1000)[j] exactly
        int piece = 0;
        {
            const double us = (double)j0 * step;
#pragma unroll
            for (int k = 1; k <= KP - 2; ++k) piece = (us >= ws.sp.u[k]) ? k : piece;
        }
        double un, ub, c0, c1, c2, c3, e0, e1, e2, e3;
#define CUB200_LOAD_PIECE()                                                              \
        {                                                                                \
            const double2* pc = reinterpret_cast<const double2*>(ws.sp.piece[piece]);       \
            const double2 q0 = pc[0], q1 = pc[1], q2 = pc[2], q3 = pc[3], q4 = pc[4];    \
            un = q0.x; ub = q0.y; c0 = q1.x; c1 = q1.y; c2 = q2.x; c3 = q2.y;            \
            e0 = q3.x; e1 = q3.y; e2 = q4.x; e3 = q4.y;                                  \
        }
        CUB200_LOAD_PIECE();
        uint32_t cur = 0xffffffffu, bits = 0;
        int lowest = 0;
#pragma unroll 4
        for (int jj = 0; jj < 32; ++jj) {
            const double uj = (double)(j0 + jj) * step;
            while (piece < KP - 2 && uj >= un) {
                ++piece;
                CUB200_LOAD_PIECE();
            }
            const double sv = uj - ub;
            const double xv = fma(fma(fma(c3, sv, c2), sv, c1), sv, c0);
            const double yv = fma(fma(fma(e3, sv, e2), sv, e1), sv, e0);
            const int ix = min(__double2int_rn(xv), 255), iy = min(__double2int_rn(yv), 255);
            lowest = min(lowest, min(ix, iy));
            // [-256,-1] wraps like a numpy negative index: (v + 256) == (v & 255) there
            const uint32_t cell = ((uint32_t)(iy & 255) << 3) | ((uint32_t)(ix & 255) >> 5);   // row * 8 + word
            if (cell != cur) {
                if (bits) atomicOr(&bmp[cur + (cur >> 6) * 4], bits);
                cur = cell; bits = 0;
            }
            bits |= 1u << (ix & 31);
        }
#undef CUB200_LOAD_PIECE
        if (bits) atomicOr(&bmp[cur + (cur >> 6) * 4], bits);
        oob = __any_sync(FULL, lowest < -256);
        if (oob) {                                           // IndexError in the reference: nothing drawn
            __syncwarp();
            uint4* b4 = reinterpret_cast<uint4*>(bmp);
            for (int i = 0; i < BMP_WORDS / 4 / 32; ++i) b4[lane + 32 * i] = make_uint4(0, 0, 0, 0);
        }
    } else {
        // fallback of contour_spline(): spline = the 21 raw points (already ints when rounded)
        int ix = 0, iy = 0;
        if (lane < KP) {
            ix = min(__double2int_rn(X), 255); iy = min(__double2int_rn(Y), 255);
            oob = (ix < -256) | (iy < -256);
        }
        oob = __any_sync(FULL, oob);
        if (!oob && lane < KP) set_pixel(bmp, iy & 255, ix & 255);
    }

    // 5. closing line (r0,c0) = landmark 20 -> (r1,c1) = landmark 0, skimage.draw.line closed form:
    //    pixel i < dmaj: major = m0 + i*smaj, minor = n0 + smin * floor((2*dmin*i + dmaj) / (2*dmaj));
    //    pixel dmaj = (r1, c1).  linear_reconstruction draws all 21 edges k -> k+1 this way.
#pragma unroll 1
    for (int e = linear ? 0 : KP - 1; e < KP; ++e) {
        const int e1 = (e + 1 == KP) ? 0 : e + 1;
        const int r0 = __shfl_sync(FULL, riy, e), c0 = __shfl_sync(FULL, rix, e);
        const int r1 = __shfl_sync(FULL, riy, e1), c1 = __shfl_sync(FULL, rix, e1);
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
            set_pixel(bmp, rr, cc);
        }
    }
    __syncwarp();

    // 6. background flood.  Lane owns rows 8*lane .. 8*lane+7; R = reached-from-outside.
    uint32_t R[8][8];
    uint32_t generic = 0;        // bit i: row i needs the generic run extraction (border rows, rows the
                                 // vertical flood advanced in)
    int area = 0;
    const uint4* band4 = reinterpret_cast<const uint4*>(bmp + lane * BAND_STRIDE);
    // 6a. rows are seeded from the left and right frame borders: everything left of the first and
    //     right of the last boundary pixel of the row; rows 0 and 255 are seeded everywhere.
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 b0 = band4[2 * i], b1 = band4[2 * i + 1];
        const uint32_t B[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const uint32_t rowor = (B[0] | B[1] | B[2]) | (B[3] | B[4] | B[5]) | (B[6] | B[7]);
        const bool border_row = (lane == 0 && i == 0) || (lane == 31 && i == 7);
        if (rowor == 0u) {
#pragma unroll
            for (int w = 0; w < 8; ++w) R[i][w] = 0xffffffffu;
        } else if (border_row) {
#pragma unroll
            for (int w = 0; w < 8; ++w) R[i][w] = ~B[w];
            generic |= 1u << i;
        } else {
            // left part: bits below the lowest set bit of the 256-bit row = (B - 1) & ~B
            uint32_t X[8];
            asm("sub.cc.u32 %0, %1, 1;" : "=r"(X[0]) : "r"(B[0]));
#pragma unroll
            for (int w = 1; w < 8; ++w) asm("subc.cc.u32 %0, %1, 0;" : "=r"(X[w]) : "r"(B[w]));
            // right part: bits above the highest set bit
            bool seen = false;
            int nlow = 0, nhigh = 0;
#pragma unroll
            for (int w = 7; w >= 0; --w) {
                const uint32_t smear = __funnelshift_rc(0xffffffffu, 0u, __clz(B[w]));  // 0 when B[w] == 0
                const uint32_t lo = X[w] & ~B[w], hi = seen ? 0u : ~smear;
                R[i][w] = lo | hi;
                nlow += __popc(lo);
                nhigh += __popc(hi);
                seen |= (B[w] != 0u);
            }
            // An ordinary row holds the single run [first, last boundary pixel] unless the vertical
            // flood later advances in it: emit it now (+1 at its start, -1 one past its end).
            const int xs = nlow, xe = 256 - nhigh;
            area += xe - xs;
            if (diff_img) {
                int32_t* drow = diff_img + (lane * 8 + i) * 256;
                atomicAdd(drow + xs, 1);
                if (xe < 256) atomicAdd(drow + xe, -1);
            }
        }
    }
    // 6b. vertical propagation to a fixed point: alternating Gauss-Seidel sweeps inside the band
    //     (neighbour lanes' edge rows by shuffle), stopping after the first sweep in which no lane
    //     advanced.  Convex-ish contours exit after one quiescent sweep.
    for (;;) {
        uint32_t touched = flood_sweep<0>(R, band4, lane);
        generic |= touched;
        if (!__any_sync(FULL, touched != 0u)) break;
        touched = flood_sweep<1>(R, band4, lane);
        generic |= touched;
        if (!__any_sync(FULL, touched != 0u)) break;
    }

    // 7. mask = not reached.  Ordinary rows were emitted in 6a.  Rows the vertical flood advanced in
    //    (concave / self-intersecting samples) take their 6a run back and go, like the frame-border
    //    rows, through the generic run extraction.
    const bool top = (lane == 0), bottom = (lane == 31);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = lane * 8 + i;
        if (mask_out) {
            uint4* o = reinterpret_cast<uint4*>(mask_out + row * 8);
            o[0] = make_uint4(~R[i][0], ~R[i][1], ~R[i][2], ~R[i][3]);
            o[1] = make_uint4(~R[i][4], ~R[i][5], ~R[i][6], ~R[i][7]);
        }
        if ((generic >> i) & 1u) {
            int32_t* drow = diff_img ? diff_img + row * 256 : nullptr;
            const bool border_row = (top && i == 0) || (bottom && i == 7);
            if (!border_row) area -= undo_row_span(bmp + lane * BAND_STRIDE + i * 8, drow);
            area += emit_row_generic(R[i][0], R[i][1], R[i][2], R[i][3], R[i][4], R[i][5], R[i][6], R[i][7], drow);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(FULL, area, o);
    __syncwarp();
    return area;
}

// ---------------------------------------------------------------- kernels
template <bool FROM_CONTOURS>
__global__ void __launch_bounds__(FW_THREADS, FW_MINBLOCKS)
sample_raster_kernel(const float* __restrict__ mu, const float* __restrict__ a_t, int zdim,
                     const float* __restrict__ z, uint64_t seed, int64_t n_img, int64_t n_samples,
                     int round_landmarks, const float* __restrict__ contours_in, const double* __restrict__ contours_in64,
                     float* __restrict__ contours_out, int32_t* __restrict__ area_out,
                     int32_t* __restrict__ diff, uint32_t* __restrict__ masks_out, int64_t per_image) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& ws = reinterpret_cast<WarpScratch*>(smem_raw)[threadIdx.x >> 5];
    const int lane = lane_id();
    const int64_t total = FROM_CONTOURS ? n_samples : n_img * n_samples;
    // each CTA walks a contiguous chunk of samples, so its 8 warps share the image's factor in L1
    const int64_t chunk = (total + gridDim.x - 1) / gridDim.x;
    const int64_t g_end = min(total, (int64_t)(blockIdx.x + 1) * chunk);
    for (int64_t g = (int64_t)blockIdx.x * chunk + (threadIdx.x >> 5); g < g_end; g += FW_WARPS) {
        float fx = 0.f, fy = 0.f;
        double dx = 0.0, dy = 0.0;
        int64_t img, map;
        if (FROM_CONTOURS) {
            img = per_image > 0 ? g / per_image : 0;
            map = img;
            if (lane < KP) {
                if (contours_in64) {
                    const double2 p = reinterpret_cast<const double2*>(contours_in64)[g * KP + lane];
                    dx = p.x; dy = p.y;
                } else {
                    const float2 p = reinterpret_cast<const float2*>(contours_in)[g * KP + lane];
                    dx = p.x; dy = p.y;
                }
            }
        } else {
            img = g / n_samples;
            map = img / per_image;                          // per_image = images_per_map here
            // 1. draws
            if (z) {
                for (int e = lane; e < zdim; e += 32) ws.z[e] = z[g * zdim + e];
            } else if (lane * 4 < zdim) {
                uint32_t r[4];
                philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)lane, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
                float n0, n1, n2, n3;
                box_muller(r[0], r[1], n0, n1);
                box_muller(r[2], r[3], n2, n3);
                ws.z[4 * lane] = n0; ws.z[4 * lane + 1] = n1; ws.z[4 * lane + 2] = n2; ws.z[4 * lane + 3] = n3;
            }
            __syncwarp();
            if (lane < KP) {
                const float2 m = reinterpret_cast<const float2*>(mu)[img * KP + lane];
                fx = m.x; fy = m.y;
                const float2* a = reinterpret_cast<const float2*>(a_t + img * (int64_t)zdim * CUB200_P) + lane;
                for (int e0 = 0; e0 < zdim; e0 += 8) {            // 8 factor rows in flight
                    float2 av[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) av[q] = (e0 + q < zdim) ? __ldg(a + (e0 + q) * KP) : make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float ze = (e0 + q < zdim) ? ws.z[e0 + q] : 0.f;
                        fx = fmaf(av[q].x, ze, fx);
                        fy = fmaf(av[q].y, ze, fy);
                    }
                }
                if (contours_out) reinterpret_cast<float2*>(contours_out)[g * KP + lane] = make_float2(fx, fy);
            }
            __syncwarp();
            if (!diff && !area_out && !masks_out) continue;     // sample-only call
        }
        if (!FROM_CONTOURS) { dx = fx; dy = fy; }
        const int area = raster_sample(ws, dx, dy, round_landmarks,
                                       diff ? diff + map * 65536 : nullptr,
                                       masks_out ? masks_out + g * CUB200_MASK_WORDS : nullptr);
        if (area_out && lane == 0) area_out[g] = area;
    }
}

static int launch_cfg(int64_t total, const void* kernel, int* grid_out) {
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        CUB200_CUDA_OK(cudaGetDevice(&dev));
        CUB200_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    CUB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(WarpScratch) * FW_WARPS)));
    const int64_t need = (total + FW_WARPS - 1) / FW_WARPS;
    const int64_t cap = (int64_t)num_sms * FW_MINBLOCKS;   // resident CTAs per SM; each walks a contiguous chunk
    *grid_out = (int)(need < cap ? need : cap);
    return 0;
}

}  // namespace cub200

extern "C" int cub200_sample_raster_aggregate(const float* mu, const float* a_t, int zdim, const float* z,
                                              uint64_t seed, int64_t n_img, int64_t n_samples,
                                              int round_landmarks, float* contours_out, int32_t* area_out,
                                              int32_t* diff, int64_t images_per_map, uint32_t* masks_out, void* stream) {
    using namespace cub200;
    if (!mu || !a_t || zdim < 1 || zdim > CUB200_ZMAX || n_img < 0 || n_samples < 0) return CUB200_EINVAL;
    if (images_per_map < 1 || n_img % images_per_map) return CUB200_EINVAL;
    if (!diff && !area_out && !masks_out && !contours_out) return CUB200_EINVAL;
    const int64_t total = n_img * n_samples;
    if (total == 0) return 0;
    int grid = 0;
    int rc = launch_cfg(total, (const void*)sample_raster_kernel<false>, &grid);
    if (rc) return rc;
    sample_raster_kernel<false><<<grid, FW_THREADS, sizeof(WarpScratch) * FW_WARPS, (cudaStream_t)stream>>>(
        mu, a_t, zdim, z, seed, n_img, n_samples, round_landmarks, nullptr, nullptr, contours_out, area_out, diff, masks_out, images_per_map);
    return check_launch("sample_raster_kernel");
}

extern "C" int cub200_rasterize_contours(const float* contours, int64_t n, int round_landmarks, uint32_t* masks_out,
                                         int32_t* area_out, int32_t* diff, int64_t per_image, void* stream) {
    using namespace cub200;
    if (n == 0) return 0;
    if (!contours || n < 0 || (diff && per_image <= 0)) return CUB200_EINVAL;
    int grid = 0;
    int rc = launch_cfg(n, (const void*)sample_raster_kernel<true>, &grid);
    if (rc) return rc;
    sample_raster_kernel<true><<<grid, FW_THREADS, sizeof(WarpScratch) * FW_WARPS, (cudaStream_t)stream>>>(
        nullptr, nullptr, 0, nullptr, 0, 0, n, round_landmarks, contours, nullptr, nullptr, area_out, diff, masks_out, per_image);
    return check_launch("sample_raster_kernel<contours>");
}

extern "C" int cub200_rasterize_contours_f64(const double* contours, int64_t n, int round_landmarks, uint32_t* masks_out,
                                             int32_t* area_out, int32_t* diff, int64_t per_image, void* stream) {
    using namespace cub200;
    if (n == 0) return 0;
    if (!contours || n < 0 || (diff && per_image <= 0)) return CUB200_EINVAL;
    int grid = 0;
    int rc = launch_cfg(n, (const void*)sample_raster_kernel<true>, &grid);
    if (rc) return rc;
    sample_raster_kernel<true><<<grid, FW_THREADS, sizeof(WarpScratch) * FW_WARPS, (cudaStream_t)stream>>>(
        nullptr, nullptr, 0, nullptr, 0, 0, n, round_landmarks, nullptr, contours, nullptr, area_out, diff, masks_out, per_image);
    return check_launch("sample_raster_kernel<contours f64>");
}

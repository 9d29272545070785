// K3  sample_raster_aggregate: Monte-Carlo contour sampler -> spline rasteriser -> background
// flood -> streaming aggregation.  Masks never leave shared memory.
//
// Replaces (reference paths, cu/ = contour_uncertainty/):
//   AleatoricUncertaintyTask.sample            cu/task/regression/aleatoric.py:54-78
//   ContourUncertaintyTask.convert_to_mask     cu/task/regression/contour_uncertainty.py:26-57
//   USContourToMask.__call__ (LV branch)       cu/data/camus/utils.py:31-98
//   reconstruction / contour_spline            cu/utils/contour.py:9-40
//   lv_area                                    cu/utils/clinical.py:92-93
//
// Two kernels:
//   sample_contours_kernel  ONE THREAD PER SAMPLE: contour = mu + A z (z replayed from memory or Philox4x32-10 +
//       Box-Muller), the image's factor staged in shared memory, contours written coalesced.  contour_samples is an
//       output of the API anyway (168 B / sample), so the rasteriser simply reads it back.
//   raster_kernel           ONE WARP PER SAMPLE, lane k = landmark k:
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
//      FAST PATH (every sample of a well-behaved contour): the flood seeded from the left and right frame borders
//      reaches, in every row, exactly the pixels outside [first, last boundary pixel]; it can only go further by
//      stepping VERTICALLY onto a free pixel inside a row's span, i.e. onto a pixel of the span that lies outside the
//      span of the row above or below.  So: per row the span [xs, xe) and the lengths of the boundary runs that start
//      at xs and end at xe - 1; if in every row the part of the span not covered by both neighbour spans is covered by
//      those runs, the seeded flood is the fixed point and the mask is the spans (lane = row, 8 x 32 rows, row groups
//      without a boundary pixel are skipped by the whole warp).  Exact, not a heuristic.
//      GENERAL PATH (concave / self-intersecting samples, contours touching rows 0 / 255): the bit-parallel flood with
//      Gauss-Seidel vertical sweeps (flood.cuh), lane = band of 8 rows.  It runs as a SECOND KERNEL over the samples the
//      first one queued (index list in the caller's workspace), so that its 64-register bitmap state does not set the
//      register budget -- and with it the number of resident warps -- of the common path.
//   7. area = mask pixels; every run boundary of every mask row becomes a +1 / -1 red.global into the
//      image's difference map (the row prefix sum of which is the per-pixel count).
//
// Shared memory per warp: the 256x256 boundary bitmap (band stride 68 words: the 128-bit row loads are bank-conflict
// free) + 1.7 KB of spline pieces.
#include <climits>
#include <type_traits>
#include "common.cuh"
#include "flood.cuh"
#include "spline.cuh"

namespace cub200 {

#ifndef CUB200_FW_WARPS
#define CUB200_FW_WARPS 10
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
constexpr int FG_WARPS = 8;                     // warps per CTA of the general-flood kernel (128 registers)

template <bool GENERAL>
struct __align__(16) WarpScratch {
    uint32_t bmp[GENERAL ? BMP_WORDS : 2048];   // general flood: padded bands (flood.cuh); fast path: plain rows of 8 words
    SplineScratchIdx sp;
};

// ---------------------------------------------------------------- bitmap helpers
template <bool PADDED>
__device__ __forceinline__ void set_pixel(uint32_t* bmp, int row, int col) {
    atomicOr(&bmp[bmp_addr_t<PADDED>(row, col >> 5)], 1u << (col & 31));
}

// Generic run extraction of one mask row (mask = ~R): every transition of R along the row (virtual
// R = 1 left of column 0) is +1 where the mask starts, -1 one past where it ends.  Returns the row's
// mask area.  Slow path only.
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

// General path of step 6/7: the bit-parallel flood of flood.cuh over the warp's bitmap, then the generic run extraction
// of every row.  Lane owns rows 8*lane .. 8*lane+7.  Returns the lane's share of the mask area.
__device__ __forceinline__ int fill_slow(const uint32_t* bmp, int lane, int32_t* __restrict__ diff_img,
                                      uint32_t* __restrict__ mask_out) {
    uint32_t R[8][8];
    const uint4* band4 = reinterpret_cast<const uint4*>(bmp + lane * BAND_STRIDE);
    flood_background(R, band4, lane);
    int area = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = lane * 8 + i;
        if (mask_out) {
            uint4* o = reinterpret_cast<uint4*>(mask_out + row * 8);
            o[0] = make_uint4(~R[i][0], ~R[i][1], ~R[i][2], ~R[i][3]);
            o[1] = make_uint4(~R[i][4], ~R[i][5], ~R[i][6], ~R[i][7]);
        }
        const uint32_t full = (R[i][0] & R[i][1] & R[i][2]) & (R[i][3] & R[i][4] & R[i][5]) & (R[i][6] & R[i][7]);
        if (full != 0xffffffffu)
            area += emit_row_generic(R[i][0], R[i][1], R[i][2], R[i][3], R[i][4], R[i][5], R[i][6], R[i][7],
                                     diff_img ? diff_img + row * 256 : nullptr);
    }
    return area;
}

// Row descriptor of the fast path: xs | xe << 16; an empty row is xs = 256, xe = 0, which makes every min / max of the
// neighbour test come out right without a special case.
constexpr uint32_t ROW_EMPTY = 256u;

// Span and end runs of one 256-bit row given as 8 words (bit b of word w = column 32 w + b), rowor != 0.
// Returns xs | xe << 16 and lr = Lrun | Rrun << 16: Lrun = length of the run of set bits that starts at xs, Rrun = length of
// the run that ends at xe - 1.  The first / last non-zero words are found through an 8-bit occupancy mask and re-read from
// shared memory (`brow`), through which a run that reaches the end of its 32-bit word is continued as well.
__device__ __forceinline__ uint32_t row_span(const uint4 b0, const uint4 b1, const uint32_t* brow, uint32_t& lr) {
    const uint32_t nz = (min(b0.x, 1u) | (min(b0.y, 1u) << 1)) | ((min(b0.z, 1u) << 2) | (min(b0.w, 1u) << 3)) |
                        ((min(b1.x, 1u) << 4) | (min(b1.y, 1u) << 5)) | ((min(b1.z, 1u) << 6) | (min(b1.w, 1u) << 7));
    const int wf = __ffs(nz) - 1, wl = 31 - __clz(nz);
    const uint32_t fw = brow[wf], lw = brow[wl];
    const int bf = __ffs(fw) - 1;                              // fw != 0
    int lrun = __ffs(~(fw >> bf)) - 1;                         // >> shifts zeros in: the complement is 0 only for bf == 0, fw == ~0
    if (lrun < 0) lrun = 32;
    const int cl = __clz(lw);
    int rrun = __clz(~(lw << cl));                             // clz(0) == 32: the whole word
    const int xs = 32 * wf + bf, xe = 32 * wl + 32 - cl;
    if (bf + lrun == 32) {                                     // the run goes on in the next word(s)
        for (int w = wf + 1; w < 8; ++w) {
            const uint32_t v = brow[w];
            if (v == 0xffffffffu) { lrun += 32; continue; }
            lrun += __ffs(~v) - 1;
            break;
        }
    }
    if (cl + rrun == 32) {
        for (int w = wl - 1; w >= 0; --w) {
            const uint32_t v = brow[w];
            if (v == 0xffffffffu) { rrun += 32; continue; }
            rrun += __clz(~v);
            break;
        }
    }
    lr = (uint32_t)lrun | ((uint32_t)rrun << 16);
    return (uint32_t)xs | ((uint32_t)xe << 16);
}

// 3c/4. The lane's 32 consecutive evaluation points of the fitted spline (np.linspace(0, 1, 1000)[j], j = 32 lane ..), each
// rounded half-even, clipped at 255, wrapped like a numpy negative index and set in the bitmap.  Returns the lowest index seen
// (< -256: the reference's assignment raises).
// MAGIC: rint() through `v + 1.5 * 2^52`, whose low word is the integer for |v| < 2^31 (fit_spline's `wide` flag guarantees
// 2^30): a DADD on the fp64 pipe instead of F2I.F64, which runs on the conversion unit (XU) at a fraction of the rate -- ncu had
// the XU pipe at 56 % with F2I, the busiest unit of the kernel (profiles/r02c_*).
template <bool GENERAL, bool MAGIC>
__device__ __forceinline__ int eval_draw(const SplineScratchIdx& sp, uint32_t* bmp, int lane) {
    const double step = 1.0 / (double)(NPTS - 1);       // j * step == np.linspace(0, 1, 1000)[j] exactly
    const int j0 = min(lane * 32, NPTS - 32);           // lane 31 re-draws 24 points of lane 30
    // number of pieces that end at or before j0: branch-free binary search (jnext is non-decreasing, entries 19..31 = INT_MAX)
    int piece = 0;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) piece += (sp.jnext[piece + sft - 1] <= j0) ? sft : 0;
    double ub, c0, c1, c2, c3, e0, e1, e2, e3;
    int jnext;
#define CUB200_LOAD_PIECE()                                                              \
    {                                                                                    \
        const double2* pc = reinterpret_cast<const double2*>(sp.piece[piece]);           \
        const double2 q0 = pc[0], q1 = pc[1], q2 = pc[2], q3 = pc[3], q4 = pc[4];        \
        jnext = sp.jnext[piece]; ub = q0.y;                                              \
        c0 = q1.x; c1 = q1.y; c2 = q2.x; c3 = q2.y;                                      \
        e0 = q3.x; e1 = q3.y; e2 = q4.x; e3 = q4.y;                                      \
    }
    CUB200_LOAD_PIECE();
    int lowest = 0;
    double jd = (double)j0;
#pragma unroll 4
    for (int jj = 0; jj < 32; ++jj) {
        if (j0 + jj >= jnext) {                          // jnext of the last piece is INT_MAX
            piece = sp.pnext[piece];
            CUB200_LOAD_PIECE();
        }
        const double sv = jd * step - ub;
        jd += 1.0;
        const double xv = fma(fma(fma(c3, sv, c2), sv, c1), sv, c0);
        const double yv = fma(fma(fma(e3, sv, e2), sv, e1), sv, e0);
        int ix, iy;
        if (MAGIC) {
            ix = __double2loint(xv + 6755399441055744.0); iy = __double2loint(yv + 6755399441055744.0);
        } else {
            ix = __double2int_rn(xv); iy = __double2int_rn(yv);
        }
        ix = min(ix, 255); iy = min(iy, 255);
        lowest = __vimin3_s32(lowest, ix, iy);
        // [-256,-1] wraps like a numpy negative index: (v + 256) == (v & 255) there
        const uint32_t cell = ((uint32_t)(iy & 255) << 3) | ((uint32_t)(ix & 255) >> 5);   // row * 8 + word
        // one shared-memory atomic per point: merging the bits of consecutive points that share a word saved nothing (some lane
        // of the warp changes word at nearly every point) and cost a compare, a select and a branch per point
        atomicOr(&bmp[GENERAL ? cell + ((cell >> 4) & ~3u) : cell], 1u << (ix & 31));       // + (row >> 3) * 4 when padded
    }
#undef CUB200_LOAD_PIECE
    return lowest;
}

// ---------------------------------------------------------------- one sample, one warp
// Lane k < 21 holds landmark k as (fx, fy).  Returns the mask area; accumulates run boundaries
// into diff_img (may be null) and stores the packed mask (may be null).
// flags: bit 0 = round the landmarks before the spline (USContourToMask), bit 1 = linear_reconstruction
// (cu/utils/contour.py:43-53): no spline, the 21 rounded landmarks are joined by skimage.draw.line segments.
// GENERAL = false: the fast fill; returns RASTER_DEFERRED (nothing written) when the sample needs the general flood.
constexpr int RASTER_DEFERRED = -1;
template <bool GENERAL>
__device__ __forceinline__ int raster_sample(WarpScratch<GENERAL>& ws, double fx, double fy, int flags,
                                             int32_t* __restrict__ diff_img, uint32_t* __restrict__ mask_out) {
    const int lane = lane_id();
    const bool round_landmarks = (flags & 1) != 0, linear = (flags & 2) != 0;
    uint32_t* bmp = ws.bmp;

    // zero the boundary bitmap (uint4 stores).  The general flood reads bands of 8 rows per lane (padded band stride, flood.cuh);
    // the fast path reads rows and keeps them plain row-major.
    constexpr int WORDS = GENERAL ? BMP_WORDS : 2048;
    {
        uint4* b4 = reinterpret_cast<uint4*>(bmp);
        const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < WORDS / 4 / 32; ++i) b4[lane + 32 * i] = zero;
    }
    __syncwarp();

    // 2. rounded landmarks (closing line always uses them; the spline too when round_landmarks)
    fx = fmin(fmax(fx, -(double)COORD_CLAMP), (double)COORD_CLAMP);
    fy = fmin(fmax(fy, -(double)COORD_CLAMP), (double)COORD_CLAMP);
    const int rix = __double2int_rn(fx), riy = __double2int_rn(fy);
    const double X = round_landmarks ? (double)rix : fx;
    const double Y = round_landmarks ? (double)riy : fy;

    // 3a/3b. chord-length parameter + not-a-knot cubic spline (spline.cuh)
    bool wide = false;
    const bool bad = linear ? false : fit_spline(ws.sp, X, Y, lane, &wide);

    bool oob = false;
    if (linear) {
        // polygon outline only: every edge is drawn by step 5
    } else if (!bad) {
        // 3c/4. evaluate 1000 points, round half-even, apply the index rules and set the pixels on the fly.
        //       An index < -256 makes the
        //       reference's whole assignment raise: then the bitmap is simply cleared again (the
        //       closing line is drawn afterwards).
        // (`wide`: a piece that may leave +-2^30 -- astronomically rare -- converts with F2I instead of the magic-number add)
        const int lowest = wide ? eval_draw<GENERAL, false>(ws.sp, bmp, lane) : eval_draw<GENERAL, true>(ws.sp, bmp, lane);
        oob = __any_sync(FULL, lowest < -256);
        if (oob) {                                           // IndexError in the reference: nothing drawn
            __syncwarp();
            uint4* b4 = reinterpret_cast<uint4*>(bmp);
            for (int i = 0; i < WORDS / 4 / 32; ++i) b4[lane + 32 * i] = make_uint4(0, 0, 0, 0);
        }
    } else {
        // fallback of contour_spline(): spline = the 21 raw points (already ints when rounded)
        int ix = 0, iy = 0;
        if (lane < KP) {
            ix = min(__double2int_rn(X), 255); iy = min(__double2int_rn(Y), 255);
            oob = (ix < -256) | (iy < -256);
        }
        oob = __any_sync(FULL, oob);
        if (!oob && lane < KP) set_pixel<GENERAL>(bmp, iy & 255, ix & 255);
    }
    __syncwarp();

    // 5. closing line (r0,c0) = landmark 20 -> (r1,c1) = landmark 0; linear_reconstruction draws all 21 edges k -> k+1
#pragma unroll 1
    for (int e = linear ? 0 : KP - 1; e < KP; ++e) {
        const int e1 = (e + 1 == KP) ? 0 : e + 1;
        draw_line_clipped<GENERAL>(bmp, __shfl_sync(FULL, riy, e), __shfl_sync(FULL, rix, e), __shfl_sync(FULL, riy, e1),
                          __shfl_sync(FULL, rix, e1), lane);
    }
    __syncwarp();

    int area = 0;
    if constexpr (GENERAL) {
        area = fill_slow(bmp, lane, diff_img, mask_out);
    } else {
        // 6. fast path: lane = row 32 i + lane.  Spans and end runs of the rows ...
        uint32_t d[8], lr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = 32 * i + lane;
            const uint32_t* brow = bmp + row * 8;
            const uint4 b0 = reinterpret_cast<const uint4*>(brow)[0], b1 = reinterpret_cast<const uint4*>(brow)[1];
            const uint32_t rowor = (b0.x | b0.y | b0.z) | (b0.w | b1.x | b1.y) | (b1.z | b1.w);
            d[i] = ROW_EMPTY; lr[i] = 0u;
            if (__any_sync(FULL, rowor != 0u)) {             // warp-uniform: row groups outside the contour cost 8 instructions
                if (rowor != 0u) d[i] = row_span(b0, b1, brow, lr[i]);
            }
        }
        // ... then the exact test that the border-seeded flood cannot advance vertically anywhere.
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (!__any_sync(FULL, d[i] != ROW_EMPTY)) continue;   // warp-uniform
            // spans of the rows above and below: neighbour lanes, across a row group lane 31 <-> lane 0 of the next group
            uint32_t up = __shfl_up_sync(FULL, d[i], 1), dn = __shfl_down_sync(FULL, d[i], 1);
            const uint32_t up_x = __shfl_sync(FULL, d[i > 0 ? i - 1 : 0], 31), dn_x = __shfl_sync(FULL, d[i < 7 ? i + 1 : 7], 0);
            if (lane == 0) up = (i > 0) ? up_x : ROW_EMPTY;
            if (lane == 31) dn = (i < 7) ? dn_x : ROW_EMPTY;
            const int xs = d[i] & 0xffff, xe = d[i] >> 16;
            const int need_l = min(max((int)(up & 0xffff), (int)(dn & 0xffff)), xe) - xs;
            const int need_r = xe - max(min((int)(up >> 16), (int)(dn >> 16)), xs);
            const bool row_ok = need_l <= (int)(lr[i] & 0xffff) && need_r <= (int)(lr[i] >> 16);
            ok = ok && (d[i] == ROW_EMPTY || row_ok);
        }
        // the frame-border rows 0 and 255 are seeded everywhere: a boundary pixel there needs the general flood
        if ((lane == 0 && d[0] != ROW_EMPTY) || (lane == 31 && d[7] != ROW_EMPTY)) ok = false;
        if (!__all_sync(FULL, ok)) return RASTER_DEFERRED;

        // 7. mask rows = spans
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool any = __any_sync(FULL, d[i] != ROW_EMPTY);
            const int row = 32 * i + lane;
            const int xs = d[i] & 0xffff, xe = d[i] >> 16;
            if (any && d[i] != ROW_EMPTY) {
                area += xe - xs;
                if (diff_img) {
                    int32_t* drow = diff_img + row * 256;
                    atomicAdd(drow + xs, 1);
                    if (xe < 256) atomicAdd(drow + xe, -1);
                }
            }
            if (mask_out) {
                uint32_t m[8];
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    const int lo = max(xs - 32 * w, 0), hi = min(xe - 32 * w, 32);
                    m[w] = (hi > lo) ? ((0xffffffffu >> (32 - hi)) & (0xffffffffu << lo)) : 0u;
                }
                uint4* o = reinterpret_cast<uint4*>(mask_out + row * 8);
                o[0] = make_uint4(m[0], m[1], m[2], m[3]);
                o[1] = make_uint4(m[4], m[5], m[6], m[7]);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(FULL, area, o);
    __syncwarp();
    return area;
}

// ---------------------------------------------------------------- kernels
// contour = mu + A z, one thread per sample, block = SC_THREADS samples of one image (blockIdx.x = image * chunks + chunk).
constexpr int SC_THREADS = 64;
constexpr int SC_PAD = CUB200_P + 1;            // staging row stride (floats): conflict-free column access


template <int ZD>
__global__ void __launch_bounds__(SC_THREADS)
sample_contours_kernel(const float* __restrict__ mu, const float* __restrict__ a_t, int zdim, const float* __restrict__ z,
                       uint64_t seed, uint64_t g_offset, const uint64_t* __restrict__ step_dev, uint64_t step_stride,
                       int64_t n_samples, int chunks, float* __restrict__ contours_out, uint32_t* __restrict__ queue) {
    __shared__ __align__(16) float a_s[CUB200_ZMAX * CUB200_P];
    if (queue && blockIdx.x == 0 && threadIdx.x == 0) queue[0] = 0u;   // the rasteriser that follows starts with an empty queue
    __shared__ float out_s[SC_THREADS * SC_PAD];
    const int64_t img = blockIdx.x / chunks;
    const int64_t s0 = (int64_t)(blockIdx.x % chunks) * SC_THREADS;
    const int tid = threadIdx.x;
    const int nz = ZD > 0 ? ZD : zdim;
    for (int e = tid; e < nz * CUB200_P; e += SC_THREADS) a_s[e] = a_t[img * (int64_t)nz * CUB200_P + e];
    __syncthreads();
    const int64_t s = s0 + tid;
    const int n_here = (int)min((int64_t)SC_THREADS, n_samples - s0);
    if (s < n_samples) {
        const int64_t g = img * n_samples + s;
        float zz[CUB200_ZMAX + 2];
        if (z) {
#pragma unroll
            for (int e = 0; e < CUB200_ZMAX; ++e) zz[e] = (e < nz) ? z[g * nz + e] : 0.f;
        } else {
            // draws keyed by (seed, global sample index); a device-side step counter (CUDA-graph replays) shifts the index range
            const uint64_t gc = (uint64_t)g + g_offset + (step_dev ? *step_dev * step_stride : 0ull);
#pragma unroll
            for (int q = 0; q < (CUB200_ZMAX + 3) / 4; ++q) {
                if (4 * q < nz) philox_normals4(gc, q, seed, zz[4 * q], zz[4 * q + 1], zz[4 * q + 2], zz[4 * q + 3]);
                else zz[4 * q] = zz[4 * q + 1] = zz[4 * q + 2] = zz[4 * q + 3] = 0.f;
            }
        }
        const float2* mu2 = reinterpret_cast<const float2*>(mu) + img * KP;
        const float2* a2 = reinterpret_cast<const float2*>(a_s);
#pragma unroll 3
        for (int k = 0; k < KP; ++k) {
            const float2 m = __ldg(mu2 + k);
            float fx = m.x, fy = m.y;
#pragma unroll
            for (int e = 0; e < CUB200_ZMAX; ++e) {
                if (e < nz) {                                 // same order as a sequential dot product over the draws
                    const float2 a = a2[e * KP + k];
                    fx = fmaf(a.x, zz[e], fx);
                    fy = fmaf(a.y, zz[e], fy);
                }
            }
            out_s[tid * SC_PAD + 2 * k] = fx;
            out_s[tid * SC_PAD + 2 * k + 1] = fy;
        }
    }
    __syncthreads();
    float* dst = contours_out + (img * n_samples + s0) * CUB200_P;
    for (int e = tid; e < n_here * CUB200_P; e += SC_THREADS) dst[e] = out_s[(e / CUB200_P) * SC_PAD + (e % CUB200_P)];
}

// Few samples per factor (the sequence sampler draws ONE contour from each of its per-sample factors): one warp per sample,
// lane k = landmark k, the factor streamed from global memory (coalesced rows).  Same draws, same summation order.
constexpr int SW_WARPS = 8;
__global__ void __launch_bounds__(SW_WARPS * 32)
sample_contours_warp_kernel(const float* __restrict__ mu, const float* __restrict__ a_t, int zdim, const float* __restrict__ z,
                            uint64_t seed, uint64_t g_offset, const uint64_t* __restrict__ step_dev, uint64_t step_stride,
                            int64_t n_img, int64_t n_samples, float* __restrict__ contours_out, uint32_t* __restrict__ queue) {
    __shared__ float z_s[SW_WARPS][CUB200_ZMAX + 2];
    if (queue && blockIdx.x == 0 && threadIdx.x == 0) queue[0] = 0u;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    float* zw = z_s[warp];
    const int64_t total = n_img * n_samples;
    if (step_dev) g_offset += *step_dev * step_stride;
    for (int64_t g = (int64_t)blockIdx.x * SW_WARPS + warp; g < total; g += (int64_t)gridDim.x * SW_WARPS) {
        const int64_t img = g / n_samples;
        if (z) {
            for (int e = lane; e < zdim; e += 32) zw[e] = z[g * zdim + e];
        } else if (lane * 4 < zdim) {
            philox_normals4((uint64_t)g + g_offset, lane, seed, zw[4 * lane], zw[4 * lane + 1], zw[4 * lane + 2], zw[4 * lane + 3]);
        }
        __syncwarp();
        if (lane < KP) {
            const float2 m = reinterpret_cast<const float2*>(mu)[img * KP + lane];
            float fx = m.x, fy = m.y;
            const float2* a = reinterpret_cast<const float2*>(a_t + img * (int64_t)zdim * CUB200_P) + lane;
            for (int e0 = 0; e0 < zdim; e0 += 8) {            // 8 factor rows in flight
                float2 av[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) av[q] = (e0 + q < zdim) ? __ldg(a + (e0 + q) * KP) : make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (e0 + q < zdim) {
                        fx = fmaf(av[q].x, zw[e0 + q], fx);
                        fy = fmaf(av[q].y, zw[e0 + q], fy);
                    }
                }
            }
            reinterpret_cast<float2*>(contours_out)[g * KP + lane] = make_float2(fx, fy);
        }
        __syncwarp();
    }
}

// Workspace of the rasteriser: ws[0] = number of queued samples, ws[1 + i] = index of the i-th sample that needs the general
// flood (written by raster_kernel, consumed by raster_general_kernel).
template <bool F64>
__device__ __forceinline__ void load_landmark(const void* __restrict__ contours_in, int64_t g, int lane, double& x, double& y) {
    x = 0.0; y = 0.0;
    if (lane < KP) {
        if (F64) {
            const double2 p = reinterpret_cast<const double2*>(contours_in)[g * KP + lane];
            x = p.x; y = p.y;
        } else {
            const float2 p = reinterpret_cast<const float2*>(contours_in)[g * KP + lane];
            x = p.x; y = p.y;
        }
    }
}

template <bool F64>
__global__ void __launch_bounds__(FW_THREADS, FW_MINBLOCKS)
raster_kernel(const void* __restrict__ contours_in, int64_t n, int flags, int32_t* __restrict__ area_out,
              int32_t* __restrict__ diff, uint32_t* __restrict__ masks_out, int64_t per_image, uint32_t* __restrict__ queue) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch<false>& ws = reinterpret_cast<WarpScratch<false>*>(smem_raw)[threadIdx.x >> 5];
    const int lane = lane_id();
    const int64_t stride = (int64_t)gridDim.x * FW_WARPS;
    int64_t g = (int64_t)blockIdx.x * FW_WARPS + (threadIdx.x >> 5);
    // the next contour is in flight (as raw bits) while this one is rasterised
    using Raw = typename std::conditional<F64, double2, float2>::type;
    const Raw* src = reinterpret_cast<const Raw*>(contours_in);
    Raw nxt = {};
    if (g < n && lane < KP) nxt = src[g * KP + lane];
    const bool small = n <= 0x7fffffffLL && per_image <= 0x7fffffffLL;   // 32-bit division for the map index
    for (; g < n; g += stride) {
        const Raw cur = nxt;
        if (g + stride < n && lane < KP) nxt = src[(g + stride) * KP + lane];
        const int64_t map = per_image <= 0 ? 0 : small ? (int64_t)((uint32_t)g / (uint32_t)per_image) : g / per_image;
        const int area = raster_sample<false>(ws, (double)cur.x, (double)cur.y, flags, diff ? diff + map * 65536 : nullptr,
                                              masks_out ? masks_out + g * CUB200_MASK_WORDS : nullptr);
        if (lane == 0) {
            if (area == RASTER_DEFERRED) queue[1 + atomicAdd(queue, 1u)] = (uint32_t)g;
            else if (area_out) area_out[g] = area;
        }
    }
}

template <bool F64>
__global__ void __launch_bounds__(FG_WARPS * 32, 2)
raster_general_kernel(const void* __restrict__ contours_in, int flags, int32_t* __restrict__ area_out,
                      int32_t* __restrict__ diff, uint32_t* __restrict__ masks_out, int64_t per_image, const uint32_t* __restrict__ queue) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch<true>& ws = reinterpret_cast<WarpScratch<true>*>(smem_raw)[threadIdx.x >> 5];
    const int lane = lane_id();
    const uint32_t count = queue[0];
    for (uint32_t q = blockIdx.x * FG_WARPS + (threadIdx.x >> 5); q < count; q += gridDim.x * FG_WARPS) {
        const int64_t g = queue[1 + q];
        double x, y;
        load_landmark<F64>(contours_in, g, lane, x, y);
        const int64_t map = per_image > 0 ? g / per_image : 0;
        const int area = raster_sample<true>(ws, x, y, flags, diff ? diff + map * 65536 : nullptr,
                                             masks_out ? masks_out + g * CUB200_MASK_WORDS : nullptr);
        if (area_out && lane == 0) area_out[g] = area;
    }
}

__global__ void zero_queue_kernel(uint32_t* queue) { queue[0] = 0u; }
__global__ void step_counter_add_kernel(uint64_t* counter, uint64_t inc) { *counter += inc; }

static int sm_count(int* out) {
    // per device (the library serves several GPUs of one process), benign race: every thread writes the same value
    static int cache[64] = {0};
    int dev = 0;
    CUB200_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !cache[dev]) {
        int v = 0;
        CUB200_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) cache[dev] = v;
        *out = v;
        return 0;
    }
    *out = cache[dev];
    return 0;
}

static int launch_raster(const void* contours, bool f64, int64_t n, int flags, int32_t* area_out, int32_t* diff,
                         uint32_t* masks_out, int64_t per_image, uint32_t* queue, bool queue_is_zero, cudaStream_t stream) {
    if (!queue || n > 0xfffffff0LL) return CUB200_EINVAL;
    int num_sms = 0;
    int rc = sm_count(&num_sms);
    if (rc) return rc;
    if (!queue_is_zero) zero_queue_kernel<<<1, 1, 0, stream>>>(queue);
    note_launches(queue_is_zero ? 1 : 2);                    // + raster_general_kernel (+ zero_queue_kernel); check_launch counts raster_kernel
    const void* kernel = f64 ? (const void*)raster_kernel<true> : (const void*)raster_kernel<false>;
    const void* general = f64 ? (const void*)raster_general_kernel<true> : (const void*)raster_general_kernel<false>;
    CUB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(WarpScratch<false>) * FW_WARPS)));
    CUB200_CUDA_OK(cudaFuncSetAttribute(general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(WarpScratch<true>) * FG_WARPS)));
    const int64_t need = (n + FW_WARPS - 1) / FW_WARPS;
    const int64_t cap = (int64_t)num_sms * FW_MINBLOCKS;   // resident CTAs; warps stride over the samples
    const int grid = (int)(need < cap ? need : cap);
    const int64_t need_g = (n + FG_WARPS - 1) / FG_WARPS;
    const int grid_g = (int)(need_g < 2 * num_sms ? need_g : 2 * num_sms);
    if (f64) {
        raster_kernel<true><<<grid, FW_THREADS, sizeof(WarpScratch<false>) * FW_WARPS, stream>>>(contours, n, flags, area_out, diff, masks_out, per_image, queue);
        raster_general_kernel<true><<<grid_g, FG_WARPS * 32, sizeof(WarpScratch<true>) * FG_WARPS, stream>>>(contours, flags, area_out, diff, masks_out, per_image, queue);
    } else {
        raster_kernel<false><<<grid, FW_THREADS, sizeof(WarpScratch<false>) * FW_WARPS, stream>>>(contours, n, flags, area_out, diff, masks_out, per_image, queue);
        raster_general_kernel<false><<<grid_g, FG_WARPS * 32, sizeof(WarpScratch<true>) * FG_WARPS, stream>>>(contours, flags, area_out, diff, masks_out, per_image, queue);
    }
    return check_launch("raster_kernel");
}

}  // namespace cub200

extern "C" int cub200_sample_raster_aggregate(const float* mu, const float* a_t, int zdim, const float* z,
                                              uint64_t seed, uint64_t sample_offset, int64_t n_img, int64_t n_samples,
                                              int round_landmarks, float* contours_out, int32_t* area_out,
                                              int32_t* diff, int64_t images_per_map, uint32_t* masks_out, void* workspace,
                                              const uint64_t* step_dev, void* stream) {
    using namespace cub200;
    if (!mu || !a_t || zdim < 1 || zdim > CUB200_ZMAX || n_img < 0 || n_samples < 0) return CUB200_EINVAL;
    if (images_per_map < 1 || n_img % images_per_map) return CUB200_EINVAL;
    if (!contours_out) return CUB200_EINVAL;                 // also the staging buffer between the two kernels
    const int64_t total = n_img * n_samples;
    if (total == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool raster = diff || area_out || masks_out;
    if (raster && !workspace) return CUB200_EINVAL;
    uint32_t* queue = raster ? (uint32_t*)workspace : nullptr;
    if (n_samples < SC_THREADS / 2) {
        const int64_t need = (total + SW_WARPS - 1) / SW_WARPS;
        const int grid = (int)(need < 148 * 16 ? need : 148 * 16);
        sample_contours_warp_kernel<<<grid, SW_WARPS * 32, 0, st>>>(mu, a_t, zdim, z, seed, sample_offset, step_dev, (uint64_t)total, n_img, n_samples, contours_out, queue);
    } else {
        const int64_t chunks = (n_samples + SC_THREADS - 1) / SC_THREADS;
        if (n_img * chunks > 0x7fffffffLL) return CUB200_EINVAL;
        const unsigned grid = (unsigned)(n_img * chunks);
        if (zdim == 34)
            sample_contours_kernel<34><<<grid, SC_THREADS, 0, st>>>(mu, a_t, zdim, z, seed, sample_offset, step_dev, (uint64_t)total, n_samples, (int)chunks, contours_out, queue);
        else if (zdim == 42)
            sample_contours_kernel<42><<<grid, SC_THREADS, 0, st>>>(mu, a_t, zdim, z, seed, sample_offset, step_dev, (uint64_t)total, n_samples, (int)chunks, contours_out, queue);
        else
            sample_contours_kernel<0><<<grid, SC_THREADS, 0, st>>>(mu, a_t, zdim, z, seed, sample_offset, step_dev, (uint64_t)total, n_samples, (int)chunks, contours_out, queue);
    }
    int rc = check_launch("sample_contours_kernel");
    if (rc || !raster) return rc;                              // sample-only call
    return launch_raster(contours_out, false, total, round_landmarks, area_out, diff, masks_out, n_samples * images_per_map,
                         (uint32_t*)workspace, true, st);
}

extern "C" int64_t cub200_raster_workspace_bytes(int64_t n) { return n < 0 ? 0 : 4 * (n + 1); }

extern "C" int cub200_step_counter_add(uint64_t* counter, uint64_t inc, void* stream) {
    using namespace cub200;
    if (!counter) return CUB200_EINVAL;
    step_counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, inc);
    return check_launch("step_counter_add_kernel");
}

extern "C" int cub200_rasterize_contours(const float* contours, int64_t n, int round_landmarks, uint32_t* masks_out,
                                         int32_t* area_out, int32_t* diff, int64_t per_image, void* workspace, void* stream) {
    using namespace cub200;
    if (n == 0) return 0;
    if (!contours || n < 0 || (diff && (per_image <= 0 || n % per_image))) return CUB200_EINVAL;
    return launch_raster(contours, false, n, round_landmarks, area_out, diff, masks_out, per_image, (uint32_t*)workspace, false,
                         (cudaStream_t)stream);
}

extern "C" int cub200_rasterize_contours_f64(const double* contours, int64_t n, int round_landmarks, uint32_t* masks_out,
                                             int32_t* area_out, int32_t* diff, int64_t per_image, void* workspace, void* stream) {
    using namespace cub200;
    if (n == 0) return 0;
    if (!contours || n < 0 || (diff && (per_image <= 0 || n % per_image))) return CUB200_EINVAL;
    return launch_raster(contours, true, n, round_landmarks, area_out, diff, masks_out, per_image, (uint32_t*)workspace, false,
                         (cudaStream_t)stream);
}

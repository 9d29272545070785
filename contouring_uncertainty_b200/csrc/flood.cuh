// Bit-packed 256x256 bitmaps owned by one warp (lane = a band of 8 rows) and the 4-connected flood over them.
// Shared by the contour rasteriser (fused.cu: binary_fill_holes of a drawn contour) and the segmentation
// post-processing (segpost.cu: binary_fill_holes of a rounded probability map).
#pragma once
#include "common.cuh"

namespace cub200 {

constexpr int BAND_STRIDE = 68;                 // 8 rows x 8 words + 4 pad words: 128-bit band loads are conflict free
constexpr int BMP_WORDS = 32 * BAND_STRIDE;     // 2176

__device__ __forceinline__ int bmp_addr(int row, int word) { return row * 8 + word + (row >> 3) * 4; }
// PADDED = false: plain row-major rows of 8 words (the rasteriser's fast path reads rows, not bands; one address add less per pixel)
template <bool PADDED>
__device__ __forceinline__ int bmp_addr_t(int row, int word) { return PADDED ? bmp_addr(row, word) : row * 8 + word; }

// 256-bit add / subtract as ONE asm statement each: the carry flag lives only inside the statement, so the compiler cannot
// schedule anything between the links of the chain (separate non-volatile asm statements have no dependence through CC.CF).
__device__ __forceinline__ void add256(uint32_t (&x)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    asm volatile("add.cc.u32 %0, %8, %16;\n\taddc.cc.u32 %1, %9, %17;\n\taddc.cc.u32 %2, %10, %18;\n\taddc.cc.u32 %3, %11, %19;\n\t"
                 "addc.cc.u32 %4, %12, %20;\n\taddc.cc.u32 %5, %13, %21;\n\taddc.cc.u32 %6, %14, %22;\n\taddc.u32 %7, %15, %23;"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
                   "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}
// x = a - 1 over 256 bits
__device__ __forceinline__ void dec256(uint32_t (&x)[8], const uint32_t (&a)[8]) {
    asm volatile("sub.cc.u32 %0, %8, 1;\n\tsubc.cc.u32 %1, %9, 0;\n\tsubc.cc.u32 %2, %10, 0;\n\tsubc.cc.u32 %3, %11, 0;\n\t"
                 "subc.cc.u32 %4, %12, 0;\n\tsubc.cc.u32 %5, %13, 0;\n\tsubc.cc.u32 %6, %14, 0;\n\tsubc.u32 %7, %15, 0;"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
}

// Fill complete runs of F that contain a seed (seeds subset of F).  256-bit rows as 8 words,
// bit b of word w = column 32w + b.  Add-with-carry trick, once upward and once on the
// bit-reversed row (rare path: only rows where the flood actually advances vertically).
static __device__ __noinline__ void run_fill(const uint32_t F[8], uint32_t r[8]) {
    uint32_t X[8];
    // upward: X = F + r
    {
        uint32_t Fc[8], rc[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) { Fc[w] = F[w]; rc[w] = r[w]; }
        add256(X, Fc, rc);
    }
    uint32_t up[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) up[w] = (F[w] & ~X[w]) | r[w];
    // downward: same on reversed bits (word w <-> 7-w, brev)
    uint32_t Fr[8], rr[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) { Fr[w] = __brev(F[7 - w]); rr[w] = __brev(r[7 - w]); }
    add256(X, Fr, rr);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t dn = Fr[7 - w] & ~X[7 - w];
        r[w] = up[w] | __brev(dn);
    }
}

// One Gauss-Seidel sweep of the vertical flood over the lane's 8-row band (UP = 0: rows 0->7,
// UP = 1: rows 7->0).  A row advances where a free, unreached pixel has a reached vertical
// neighbour; the whole run containing it is then filled.  Returns the bitmask of rows advanced.
template <int UP>
__device__ __forceinline__ uint32_t flood_sweep(uint32_t (&R)[8][8], const uint4* band4, int lane) {
    uint32_t changed = 0u;
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
        const int i = UP ? 7 - ii : ii;
        uint32_t nb[8];                                      // the row across the band edge
        if (i == 0) {
#pragma unroll
            for (int w = 0; w < 8; ++w) { nb[w] = __shfl_up_sync(FULL, R[7][w], 1); if (lane == 0) nb[w] = 0u; }
        }
        if (i == 7) {
#pragma unroll
            for (int w = 0; w < 8; ++w) { nb[w] = __shfl_down_sync(FULL, R[0][w], 1); if (lane == 31) nb[w] = 0u; }
        }
        const uint4 b0 = band4[2 * i], b1 = band4[2 * i + 1];
        const uint32_t B[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t cand[8], any = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t above = (i == 0) ? nb[w] : R[i > 0 ? i - 1 : 0][w];
            const uint32_t below = (i == 7) ? nb[w] : R[i < 7 ? i + 1 : 7][w];
            cand[w] = ~(B[w] | R[i][w]) & (above | below);
            any |= cand[w];
        }
        if (any) {
            uint32_t F[8], r[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) { F[w] = ~B[w]; r[w] = R[i][w] | cand[w]; }
            run_fill(F, r);
#pragma unroll
            for (int w = 0; w < 8; ++w) R[i][w] = r[w];
            changed |= 1u << i;
        }
    }
    return changed;
}


// binary_fill_holes' flood for the bitmap the warp holds (band4 = the lane's 8-row band of OBSTACLE pixels): R = the free
// pixels reachable from outside the frame through 4-connected free pixels.  Rows are seeded from the left and right
// frame borders (everything below the lowest / above the highest obstacle of the row; rows 0 and 255 everywhere), then
// alternating Gauss-Seidel sweeps propagate vertically until no lane advances.
__device__ __forceinline__ void flood_background(uint32_t (&R)[8][8], const uint4* band4, int lane) {
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
        } else {
            // seeded from the left and right frame borders: below the lowest / above the highest set bit of the row
            uint32_t X[8];
            dec256(X, B);
            bool seen = false;
#pragma unroll
            for (int w = 7; w >= 0; --w) {
                const uint32_t smear = __funnelshift_rc(0xffffffffu, 0u, __clz(B[w]));
                R[i][w] = (X[w] & ~B[w]) | (seen ? 0u : ~smear);
                seen |= (B[w] != 0u);
            }
        }
    }
    for (;;) {
        uint32_t touched = flood_sweep<0>(R, band4, lane);
        if (!__any_sync(FULL, touched != 0u)) break;
        touched = flood_sweep<1>(R, band4, lane);
        if (!__any_sync(FULL, touched != 0u)) break;
    }
}

// skimage.draw.line (Bresenham) from (r0, c0) to (r1, c1) in closed form, the pixels shared among the lanes and clipped to
// the frame as the reference's rr.clip / cc.clip do: pixel i < dmaj has major = m0 + i * smaj,
// minor = n0 + smin * floor((2 * dmin * i + dmaj) / (2 * dmaj)); pixel dmaj = (r1, c1).
template <bool PADDED = true>
__device__ __forceinline__ void draw_line_clipped(uint32_t* bmp, int r0, int c0, int r1, int c1, int lane) {
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
        atomicOr(&bmp[bmp_addr_t<PADDED>(rr, cc >> 5)], 1u << (cc & 31));
    }
}

}  // namespace cub200

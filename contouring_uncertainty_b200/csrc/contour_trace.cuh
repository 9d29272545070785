// Marching-squares contour tracing of a bit-packed mask: skimage.measure.find_contours(mask, level)[0] for a BINARY
// mask (third party, restated: see oracle/port.py::find_contours -- parity unpinned for it).  Shared by the Simpson
// kernels (simpson.cu, level 0.5) and the mask-based strain (gls.cu, level 0.9): for a 0/1 image every level in (0, 1)
// crosses the same cell edges in the same order, only the position of a point along its edge depends on the level, so
// the points are stored as half-pixel (2 row, 2 col) uint16 pairs naming the EDGE they lie on.
#pragma once
#include "common.cuh"

namespace cub200 {

constexpr int SP_NMAX = CUB200_SIMPSON_MAX_POINTS;   // contour points kept per mask

// to-edge of the segment that STARTS at edge f (to_of_from) / from-edge of the segment that ENDS at
// edge t (from_of_to), per square case (bit 0 ul, 1 ur, 2 ll, 3 lr; 'low' vertex connectivity).
// Edges: 0 top, 1 bottom, 2 left, 3 right; 9 = no such segment.
// The tables live in registers as 2-bit fields (a per-thread index into __constant__ memory is replayed once per distinct
// address of the warp, and every thread of a warp walks its own mask).
constexpr unsigned char k_to_of_from[16][4] = {
    {9, 9, 9, 9}, {2, 9, 9, 9}, {9, 9, 9, 0}, {9, 9, 9, 2}, {9, 9, 1, 9}, {1, 9, 9, 9}, {9, 9, 1, 0}, {9, 9, 9, 1},
    {9, 3, 9, 9}, {2, 3, 9, 9}, {9, 0, 9, 9}, {9, 2, 9, 9}, {9, 9, 3, 9}, {3, 9, 9, 9}, {9, 9, 0, 9}, {9, 9, 9, 9}};
constexpr unsigned char k_from_of_to[16][4] = {
    {9, 9, 9, 9}, {9, 9, 0, 9}, {3, 9, 9, 9}, {9, 9, 3, 9}, {9, 2, 9, 9}, {9, 0, 9, 9}, {3, 2, 9, 9}, {9, 3, 9, 9},
    {9, 9, 9, 1}, {9, 9, 0, 1}, {1, 9, 9, 9}, {9, 9, 1, 9}, {9, 9, 9, 2}, {9, 9, 9, 0}, {2, 9, 9, 9}, {9, 9, 9, 9}};
// from-edge of the FIRST segment a cell generates (cases 6 and 9 generate two)
constexpr unsigned char k_first_from[16] = {9, 0, 3, 3, 2, 0, 3, 3, 1, 0, 1, 1, 2, 0, 2, 9};
constexpr uint64_t pack_edge_table(const unsigned char (&t)[16][4], int half) {      // cases 8 half .. 8 half + 7; "none" -> 0
    uint64_t v = 0;
    for (int c = 0; c < 8; ++c)
        for (int e = 0; e < 4; ++e) v |= (uint64_t)(t[8 * half + c][e] > 3 ? 0 : t[8 * half + c][e]) << (2 * (4 * c + e));
    return v;
}
constexpr uint32_t pack_first_from() {
    uint32_t v = 0;
    for (int c = 0; c < 16; ++c) v |= (uint32_t)(k_first_from[c] > 3 ? 0 : k_first_from[c]) << (2 * c);
    return v;
}
constexpr uint64_t TO_OF_FROM_LO = pack_edge_table(k_to_of_from, 0), TO_OF_FROM_HI = pack_edge_table(k_to_of_from, 1);
constexpr uint64_t FROM_OF_TO_LO = pack_edge_table(k_from_of_to, 0), FROM_OF_TO_HI = pack_edge_table(k_from_of_to, 1);
constexpr uint32_t FIRST_FROM = pack_first_from();
__device__ __forceinline__ int to_of_from(int cs, int e) {
    return (int)(((cs & 8) ? TO_OF_FROM_HI : TO_OF_FROM_LO) >> (2 * (4 * (cs & 7) + e))) & 3;
}
__device__ __forceinline__ int from_of_to(int cs, int e) {
    return (int)(((cs & 8) ? FROM_OF_TO_HI : FROM_OF_TO_LO) >> (2 * (4 * (cs & 7) + e))) & 3;
}

// A mask as the Simpson code sees it: the 256 x 256 bit-packed mask itself, or (anisotropic voxel spacing) its
// nearest-neighbour stretch to H x W -- resize_image_to_isotropic (vital/vital/utils/image/transform.py:52-66) through
// PIL's NEAREST resize, whose source index per output row / column is precomputed by the host (ri, ci).
struct MaskView {
    const uint32_t* m;
    const uint16_t* ri;    // source row of each virtual row (nullptr: identity)
    const uint16_t* ci;    // source column of each virtual column
    int H, W;
    __device__ __forceinline__ int bit(int r, int c) const {
        if (ri) { r = ri[r]; c = ci[c]; }
        return (m[r * 8 + (c >> 5)] >> (c & 31)) & 1u;
    }
    __device__ __forceinline__ int cell(int r0, int c0) const {   // bit 0 ul, 1 ur, 2 ll, 3 lr of the cell at (r0, c0)
        if (ri) return bit(r0, c0) | (bit(r0, c0 + 1) << 1) | (bit(r0 + 1, c0) << 2) | (bit(r0 + 1, c0 + 1) << 3);
        // plain view: one word per row; the right-hand column sits in the next word only for c0 % 32 == 31 (c0 <= 254)
        const int w = c0 >> 5, b = c0 & 31;
        const uint32_t* p = m + r0 * 8 + w;
        uint32_t x0 = p[0] >> b, x1 = p[8] >> b;
        if (b == 31) { x0 |= p[1] << 1; x1 |= p[9] << 1; }
        return (int)((x0 & 3u) | ((x1 & 3u) << 2));
    }
};
// The tracer's view of a plain mask: the walk moves one cell at a time, so the two row words under the current cell stay in
// registers and a vertical step loads ONE word, a horizontal step inside a word none (every thread of a warp reads its own
// mask: each load instruction is 32 separate sectors, which is what bounds the kernel).
struct CellCache {
    const uint32_t* m;
    int rr, ww;            // cached: words ww of rows rr (wa) and rr + 1 (wb)
    uint32_t wa, wb;
    __device__ __forceinline__ int cell(int r0, int c0) {          // as MaskView::cell; r0 <= 254, c0 <= 254
        const int w = c0 >> 5, b = c0 & 31;
        if (w == ww && r0 == rr + 1) { wa = wb; wb = m[r0 * 8 + 8 + w]; }
        else if (w == ww && r0 == rr - 1) { wb = wa; wa = m[r0 * 8 + w]; }
        else if (w != ww || r0 != rr) { wa = m[r0 * 8 + w]; wb = m[r0 * 8 + 8 + w]; }
        rr = r0; ww = w;
        uint32_t x0 = wa >> b, x1 = wb >> b;
        if (b == 31) { x0 |= m[r0 * 8 + w + 1] << 1; x1 |= m[r0 * 8 + 8 + w + 1] << 1; }
        return (int)((x0 & 3u) | ((x1 & 3u) << 2));
    }
};
// a cell's second segment (saddle cases): the one starting at L (case 6) / B (case 9)
__device__ __forceinline__ int seg_rank(int cs, int from_edge) {
    return (cs == 6 && from_edge == 2) || (cs == 9 && from_edge == 1);
}
__device__ __forceinline__ uint32_t edge_point(int r0, int c0, int e) {
    // half-pixel units: top (2r0, 2c0+1), bottom (2r0+2, 2c0+1), left (2r0+1, 2c0), right (2r0+1, 2c0+2)
    const int r2 = 2 * r0 + (e == 0 ? 0 : e == 1 ? 2 : 1);
    const int c2 = 2 * c0 + (e == 2 ? 0 : e == 3 ? 2 : 1);
    return (uint32_t)r2 | ((uint32_t)c2 << 16);
}

// ---------------------------------------------------------------- S1: contour tracing, one thread per mask
static __global__ void __launch_bounds__(128)
contour_trace_kernel(const uint32_t* __restrict__ masks, int64_t n_masks, const uint16_t* __restrict__ row_index,
                     const uint16_t* __restrict__ col_index, int vh, int vw, uint32_t* __restrict__ pts, int32_t* __restrict__ npts) {
    extern __shared__ uint16_t s_idx[];                           // [vh rows][vw columns] source indices (stretched view)
    if (row_index) {
        for (int i = threadIdx.x; i < vh; i += blockDim.x) s_idx[i] = row_index[i];
        for (int i = threadIdx.x; i < vw; i += blockDim.x) s_idx[vh + i] = col_index[i];
        __syncthreads();
    }
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_masks) return;
    const uint32_t* m = masks + g * CUB200_MASK_WORDS;
    const MaskView mv{m, row_index ? s_idx : nullptr, row_index ? s_idx + vh : nullptr, vh, vw};
    const int RM = vh - 2, CM = vw - 2;                           // last cell row / column
    uint32_t* out = pts + g * SP_NMAX;
    // raster-first cell that generates a segment
    int fr = -1, fc = -1;
    if (!row_index) {
        uint32_t a[8], b[8];
        const uint4* m4 = reinterpret_cast<const uint4*>(m);      // two 16-byte loads per row (one sector per thread)
        {
            const uint4 u = m4[0], v = m4[1];
            a[0] = u.x; a[1] = u.y; a[2] = u.z; a[3] = u.w; a[4] = v.x; a[5] = v.y; a[6] = v.z; a[7] = v.w;
        }
        for (int r0 = 0; r0 < 255 && fr < 0; ++r0) {
            {
                const uint4 u = m4[2 * r0 + 2], v = m4[2 * r0 + 3];
                b[0] = u.x; b[1] = u.y; b[2] = u.z; b[3] = u.w; b[4] = v.x; b[5] = v.y; b[6] = v.z; b[7] = v.w;
            }
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const uint32_t a1 = (a[w] >> 1) | ((w < 7 ? a[w < 7 ? w + 1 : 7] : 0u) << 31);
                const uint32_t b1 = (b[w] >> 1) | ((w < 7 ? b[w < 7 ? w + 1 : 7] : 0u) << 31);
                uint32_t nu = (a[w] ^ b[w]) | (a[w] ^ a1) | (b[w] ^ b1);
                if (w == 7) nu &= 0x7fffffffu;                   // there is no cell at c0 = 255
                if (nu && fr < 0) { fr = r0; fc = 32 * w + __ffs(nu) - 1; }
            }
#pragma unroll
            for (int w = 0; w < 8; ++w) a[w] = b[w];
        }
    } else {
        // stretched view: skip the row pairs whose source rows are both empty, then scan the cells of the first other one
        for (int r0 = 0; r0 <= RM && fr < 0; ++r0) {
            const uint32_t* ra = m + mv.ri[r0] * 8;
            const uint32_t* rb = m + mv.ri[r0 + 1] * 8;
            uint32_t any = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) any |= ra[w] | rb[w];
            if (!any) continue;
            for (int c0 = 0; c0 <= CM; ++c0) {
                const int cs = mv.cell(r0, c0);
                if (cs != 0 && cs != 15) { fr = r0; fc = c0; break; }
            }
        }
    }
    if (fr < 0) { npts[g] = 0; return; }                          // find_contours() == [] -> IndexError
    // pass A: walk backwards from the first segment to the chain start, or around the loop
    CellCache cc{m, -4, -4, 0u, 0u};
    auto cell_at = [&](int r, int c) { return row_index ? mv.cell(r, c) : cc.cell(r, c); };
    const int cs0 = cell_at(fr, fc);
    const int f0 = (int)(FIRST_FROM >> (2 * cs0)) & 3;
    int r0 = fr, c0 = fc, fe = f0, cs = cs0;                      // current segment = (cell, from-edge); cs = the cell's case
    int sr = fr, sc = fc, sf = f0;                                // start segment of pass B
    bool closed = false;
    int last_key = -1, lr = 0, lc = 0, lt = 0;                    // raster-last segment of the loop
    for (int it = 0; it < 4 * SP_NMAX; ++it) {
        const int te = to_of_from(cs, fe);
        const int key = ((r0 * (CM + 1) + c0) << 1) | seg_rank(cs, fe);
        if (key > last_key) { last_key = key; lr = r0; lc = c0; lt = te; }
        // predecessor: the cell across the from-edge, whose segment ends on the shared edge
        int pr = r0, pc = c0, pe;
        if (fe == 0) { pr = r0 - 1; pe = 1; } else if (fe == 1) { pr = r0 + 1; pe = 0; }
        else if (fe == 2) { pc = c0 - 1; pe = 3; } else { pc = c0 + 1; pe = 2; }
        if (pr < 0 || pr > RM || pc < 0 || pc > CM) { sr = r0; sc = c0; sf = fe; break; }     // open chain start
        cs = cell_at(pr, pc);
        r0 = pr; c0 = pc; fe = from_of_to(cs, pe);
        if (r0 == fr && c0 == fc && fe == f0) { closed = true; break; }
    }
    if (closed) {
        // the loop starts at the to-point of its raster-last segment: first segment of pass B is the one after it
        int nr = lr, nc = lc, ne;
        if (lt == 0) { nr = lr - 1; ne = 1; } else if (lt == 1) { nr = lr + 1; ne = 0; }
        else if (lt == 2) { nc = lc - 1; ne = 3; } else { nc = lc + 1; ne = 2; }
        sr = nr; sc = nc; sf = ne;
    }
    // pass B: forward, recording points -- four at a time (one 16-byte store instead of four scattered 4-byte ones)
    int n = 0;
    uint32_t q0 = 0u, q1 = 0u, q2 = 0u, q3 = 0u;
    auto push = [&](uint32_t pt) {
        const int slot = n & 3;
        if (slot == 0) q0 = pt; else if (slot == 1) q1 = pt; else if (slot == 2) q2 = pt; else q3 = pt;
        ++n;
        if ((n & 3) == 0) reinterpret_cast<uint4*>(out)[(n >> 2) - 1] = make_uint4(q0, q1, q2, q3);
    };
    r0 = sr; c0 = sc; fe = sf;
    push(edge_point(r0, c0, fe));
    bool overflow = false;
    for (;;) {
        const int te = to_of_from(cell_at(r0, c0), fe);
        if (n >= SP_NMAX) { overflow = true; break; }
        push(edge_point(r0, c0, te));
        int nr = r0, nc = c0, ne;
        if (te == 0) { nr = r0 - 1; ne = 1; } else if (te == 1) { nr = r0 + 1; ne = 0; }
        else if (te == 2) { nc = c0 - 1; ne = 3; } else { nc = c0 + 1; ne = 2; }
        if (nr < 0 || nr > RM || nc < 0 || nc > CM) break;         // open chain end
        r0 = nr; c0 = nc; fe = ne;
        if (closed && r0 == sr && c0 == sc && fe == sf) break;
    }
    if (n & 3) {                                                   // the last, incomplete group
        const int base = n & ~3;
        out[base] = q0;
        if ((n & 3) > 1) out[base + 1] = q1;
        if ((n & 3) > 2) out[base + 2] = q2;
    }
    npts[g] = overflow ? -1 : n;
}

}  // namespace cub200

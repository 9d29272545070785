// Segmentation-sample post-processing: what SegmentationUncertaintyTask.predict_step
// does to every Monte-Carlo sample of a binary segmentation network before the entropy map is taken.
//
// Replaces (reference paths, cu/ = contour_uncertainty/):
//   predict_step, post-processing loop       cu/task/segmentation/segmentation_uncertainty.py:98-127
//   big_blob                                 cu/task/segmentation/utils.py:12-32
//   UncertaintyTask.sample_entropy (C == 1)  cu/task/uncertainty.py:107-133
//
// Per sample (256x256 fp32 probabilities):
//   1. samples.round(): numpy rounds half to even, so a pixel is foreground iff |p| > 0.5 (NaN counts as foreground,
//      as it does for ndimage's astype(bool))                                                        (:104)
//   2. scipy.ndimage.binary_fill_holes == complement of the 4-connected background reachable from outside the frame (:111)
//   3. big_blob: 4-connected components (ndimage.label, default structure), keep the largest; on ties np.argmax
//      keeps the lowest label == the component whose first pixel comes first in raster order    (utils.py:16-30)
//   4. samples = samples_raw * samples: the kept pixels carry their raw probability, everything else is 0   (:115)
// then per image: mean over the samples, binary entropy in bits (scipy.stats.entropy of [p, 1 - p], base 2),
// non-finite -> 0, a 10-pixel frame zeroed, pred_samples = round(samples) as bool                        (:125-140)
//
// Two kernels.  seg_blob_kernel: ONE WARP PER SAMPLE streams the sample once (HBM-bound: 256 KiB per sample),
// packs it into a 256x256 bitmap in shared memory, floods the background (flood.cuh), extracts the row runs of
// the filled mask and joins them with a lock-free union-find -- components of a segmentation output are a few
// hundred runs, not 65 536 pixels.  It writes the bit-packed blob mask (8 KiB per sample).  Run tables: 1088 runs in
// shared memory, up to 2048 in a per-warp slot of global memory, anything beyond (noise) is queued and redone by
// seg_blob_overflow_kernel with full-size tables.  Three CTAs of four warps per SM: the kernel needs ~168 registers
// and loses more to spills at 128 than to the fourth CTA it gives up (measured).  seg_entropy_kernel: one warp per (image, row) walks the samples, reads the
// probabilities only where the blob mask has pixels, sums them in fp64 in sample order and writes the entropy.
#include "flood.cuh"

#ifndef CUB200_SG_MINBLOCKS
#define CUB200_SG_MINBLOCKS 3
#endif

namespace cub200 {

constexpr int SG_WARPS = 4;
constexpr int SG_THREADS = SG_WARPS * 32;
constexpr int SG_CAP = BMP_WORDS / 2;            // 1088 runs per sample held in shared memory (see SegScratch)
constexpr int SG_WARP_SLOT = 2048;               // runs per sample held in the warp's global-memory slot
constexpr int SG_MAX_RUNS = 128 * 256;           // most runs a 256x256 mask can have
constexpr int SG_MAX_GRID = 1024;                // CTAs of seg_blob_kernel at most (it launches CUB200_SG_MINBLOCKS per SM)
constexpr int SG_OVERFLOW_CTAS = 64;             // warps (one per CTA) of the overflow kernel, one global slot each

struct __align__(16) SegScratch {
    uint32_t bmp[BMP_WORDS];                     // once the runs are extracted the bitmap is dead until the blob is drawn:
                                                 // parent[] and size[] of the union-find live in it meanwhile
    uint32_t run[SG_CAP];                        // row << 16 | first column << 8 | last column
    uint16_t row_base[264];                      // index of the first run of each row; [256] = number of runs
};

struct RunTables {
    uint32_t* run;
    uint32_t* parent;
    uint32_t* size;
    int cap;
};

// Root of x with path halving.  Links only ever point to smaller indices, so the plain stores of the halving can at
// worst re-shorten a path another lane has just shortened.
__device__ __forceinline__ uint32_t find_root(uint32_t* parent, uint32_t x) {
    volatile uint32_t* p = parent;
    for (;;) {
        const uint32_t q = p[x];
        if (q == x) return x;
        const uint32_t g = p[q];
        if (g == q) return q;
        p[x] = g;
        x = g;
    }
}

__device__ __forceinline__ uint32_t find_root_ro(const uint32_t* parent, uint32_t x) {
    const volatile uint32_t* p = parent;
    uint32_t q = p[x];
    while (q != x) { x = q; q = p[x]; }
    return x;
}

// Lock-free union by smaller index: the root of a component is its first run in raster order.
__device__ __forceinline__ void unite(uint32_t* parent, uint32_t a, uint32_t b) {
    for (;;) {
        a = find_root(parent, a);
        b = find_root(parent, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        const uint32_t old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;                                 // a was linked meanwhile: its former parent joins b as well
    }
}

// Stream one sample into the warp's bitmap.  Each row is two fully coalesced 512-byte loads: lane l holds pixels
// 4l..4l+3 and 128+4l..128+4l+3, i.e. one nibble of byte l/2 and one of byte 16 + l/2 of the packed row; a lane pair
// swaps nibbles so that even lanes own byte l/2 and odd lanes byte 16 + l/2.  The loads of the next four rows are in
// flight while the current four are packed.
__device__ __forceinline__ uint32_t nibble_of(const float4& v) {
    return ((fabsf(v.x) <= 0.5f) ? 0u : 1u) | ((fabsf(v.y) <= 0.5f) ? 0u : 2u) | ((fabsf(v.z) <= 0.5f) ? 0u : 4u) |
           ((fabsf(v.w) <= 0.5f) ? 0u : 8u);
}

__device__ __forceinline__ void load_rows(const float4* __restrict__ src, int row0, float4 (&v)[4][2]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i][0] = ld_stream_f4(src + (row0 + i) * 64);           // (plain __ldg, allocating in L1: 5 % slower)
        v[i][1] = ld_stream_f4(src + (row0 + i) * 64 + 32);
    }
}

__device__ __forceinline__ void pack_rows(const float4 (&v)[4][2], int row0, uint8_t* bytes, int lane) {
    const int byte_of_lane = ((lane & 1) << 4) + (lane >> 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t x = nibble_of(v[i][0]) | (nibble_of(v[i][1]) << 4);
        const uint32_t y = __shfl_xor_sync(FULL, x, 1);
        const uint32_t b = (lane & 1) ? ((y >> 4) | (x & 0xf0u)) : ((x & 15u) | ((y & 15u) << 4));
        bytes[4 * bmp_addr(row0 + i, 0) + byte_of_lane] = (uint8_t)b;
    }
}

__device__ __forceinline__ void pack_sample(const float* __restrict__ sample, uint32_t* bmp, int lane) {
    const float4* src = reinterpret_cast<const float4*>(sample) + lane;
    uint8_t* bytes = reinterpret_cast<uint8_t*>(bmp);
    float4 a[4][2], b[4][2];
    load_rows(src, 0, a);
#pragma unroll 1
    for (int row0 = 0; row0 < 256; row0 += 8) {
        load_rows(src, row0 + 4, b);
        pack_rows(a, row0, bytes, lane);
        if (row0 + 8 < 256) load_rows(src, row0 + 8, a);
        pack_rows(b, row0 + 4, bytes, lane);
    }
}

// Steps 2-3 on the bitmap the warp holds: bmp (foreground) -> bmp (largest blob of the hole-filled mask).
// The run tables are `small` when the sample's runs fit, else `big`.  Returns the blob area, or -1 when the sample has
// more runs than `big` holds (bmp then holds the hole-filled mask, on which the function can simply be run again).
__device__ int fill_and_keep_largest(uint32_t* bmp, uint16_t* row_base, const RunTables& small, const RunTables& big, int lane) {
    // ---- 2. background flood: R = reached from outside the frame
    uint32_t R[8][8];
    const uint4* band4 = reinterpret_cast<const uint4*>(bmp + lane * BAND_STRIDE);
    flood_background(R, band4, lane);
    // ---- the filled mask M = ~R replaces the foreground bitmap (frees the 64 registers of R)
    {
        uint4* band_w = reinterpret_cast<uint4*>(bmp + lane * BAND_STRIDE);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            band_w[2 * i] = make_uint4(~R[i][0], ~R[i][1], ~R[i][2], ~R[i][3]);
            band_w[2 * i + 1] = make_uint4(~R[i][4], ~R[i][5], ~R[i][6], ~R[i][7]);
        }
    }
    __syncwarp();
    // ---- 3a. runs of M: a run starts where M & ~(M << 1), ends where M & ~(M >> 1)
    int mine = 0;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const uint4 b0 = band4[2 * i], b1 = band4[2 * i + 1];
        const uint32_t M[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t prev = 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            mine += __popc(M[w] & ~__funnelshift_l(prev, M[w], 1));
            prev = M[w];
        }
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total > big.cap) return -1;
    const RunTables& tb = (total <= small.cap) ? small : big;
    if (total == 0) return 0;                    // M is empty: the bitmap already is the (empty) result
    int ks = incl - mine;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int row = lane * 8 + i;
        const uint4 b0 = band4[2 * i], b1 = band4[2 * i + 1];
        const uint32_t M[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        row_base[row] = (uint16_t)ks;
        int ke = ks;
        uint32_t prev = 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t m = M[w];
            const uint32_t next = (w < 7) ? M[w < 7 ? w + 1 : 7] : 0u;
            uint32_t st = m & ~__funnelshift_l(prev, m, 1);
            uint32_t en = m & ~__funnelshift_r(m, next, 1);
            prev = m;
            while (st) {                          // a run's start is met before (or with) its end: plain store, then OR
                const int b = __ffs(st) - 1;
                st &= st - 1;
                tb.run[ks] = ((uint32_t)row << 16) | ((uint32_t)(32 * w + b) << 8);
                ++ks;
            }
            while (en) {
                const int b = __ffs(en) - 1;
                en &= en - 1;
                tb.run[ke] |= (uint32_t)(32 * w + b);
                ++ke;
            }
        }
    }
    if (lane == 31) row_base[256] = (uint16_t)total;
    __syncwarp();                                // every lane is done reading the bitmap: parent / size may alias it
    for (int k = lane; k < total; k += 32) { tb.parent[k] = (uint32_t)k; tb.size[k] = 0u; }
    __syncwarp();
    // ---- 3b. join every run with the runs of the row above that share a column (4-connectivity)
    for (int k = lane; k < total; k += 32) {
        const uint32_t rk = tb.run[k];
        const int row = (int)(rk >> 16);
        if (row == 0) continue;
        const int s = (int)((rk >> 8) & 255u), e = (int)(rk & 255u);
        int lo = row_base[row - 1];
        const int hi = row_base[row];
        int n = hi - lo;
        while (n > 0) {                           // first run of the row above whose last column is >= s
            const int half = n >> 1;
            if ((int)(tb.run[lo + half] & 255u) < s) { lo += half + 1; n -= half + 1; }
            else n = half;
        }
        for (int j = lo; j < hi && (int)((tb.run[j] >> 8) & 255u) <= e; ++j) unite(tb.parent, (uint32_t)k, (uint32_t)j);
    }
    __syncwarp();
    // ---- 3c. component sizes at the roots, largest wins, ties to the lowest root
    for (int k = lane; k < total; k += 32) {
        const uint32_t rk = tb.run[k];
        const uint32_t root = find_root(tb.parent, (uint32_t)k);
        atomicAdd(&tb.size[root], (rk & 255u) - ((rk >> 8) & 255u) + 1u);
    }
    __syncwarp();
    uint32_t best_size = 0u, best_root = 0xffffffffu;
    for (int k = lane; k < total; k += 32) {
        const uint32_t sz = tb.size[k];              // non-roots hold 0
        if (sz > best_size) { best_size = sz; best_root = (uint32_t)k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t s2 = __shfl_xor_sync(FULL, best_size, o), r2 = __shfl_xor_sync(FULL, best_root, o);
        if (s2 > best_size || (s2 == best_size && r2 < best_root)) { best_size = s2; best_root = r2; }
    }
    // ---- the kept blob replaces the bitmap (its runs are flagged first: parent[] may live in the bitmap)
    for (int k = lane; k < total; k += 32)
        if (find_root_ro(tb.parent, (uint32_t)k) == best_root) tb.run[k] |= 0x80000000u;
    __syncwarp();
    {
        uint4* b4 = reinterpret_cast<uint4*>(bmp);
#pragma unroll
        for (int i = 0; i < BMP_WORDS / 4 / 32; ++i) b4[lane + 32 * i] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    for (int k = lane; k < total; k += 32) {
        const uint32_t rk = tb.run[k];
        if (!(rk >> 31)) continue;
        const int row = (int)((rk >> 16) & 255u), s = (int)((rk >> 8) & 255u), e = (int)(rk & 255u);
        for (int w = s >> 5; w <= (e >> 5); ++w) {
            const int lo = max(s - 32 * w, 0), hi = min(e - 32 * w, 31);
            const uint32_t bits = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
            atomicOr(&bmp[bmp_addr(row, w)], bits);
        }
    }
    __syncwarp();
    return (int)best_size;
}

__device__ __forceinline__ void store_mask(const uint32_t* bmp, uint32_t* __restrict__ out, int lane) {
#pragma unroll 4
    for (int i = 0; i < 64; ++i) {
        const int idx = lane + 32 * i;
        out[idx] = bmp[bmp_addr(idx >> 3, idx & 7)];
    }
}

__global__ void __launch_bounds__(SG_THREADS, CUB200_SG_MINBLOCKS)
seg_blob_kernel(const float* __restrict__ probs, int64_t n_samples, uint32_t* __restrict__ masks_out,
                int32_t* __restrict__ area_out, int32_t* __restrict__ overflow /* [0] = count, then sample ids */,
                uint32_t* __restrict__ warp_slots) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SegScratch& ws = reinterpret_cast<SegScratch*>(smem_raw)[threadIdx.x >> 5];
    const int lane = lane_id();
    const RunTables tb{ws.run, ws.bmp, ws.bmp + SG_CAP, SG_CAP};
    // second tier: a slot of global memory per warp (L2-resident in practice) for samples with up to SG_WARP_SLOT runs
    uint32_t* slot = warp_slots + ((int64_t)blockIdx.x * SG_WARPS + (threadIdx.x >> 5)) * (3 * SG_WARP_SLOT);
    const RunTables tb2{slot, slot + SG_WARP_SLOT, slot + 2 * SG_WARP_SLOT, SG_WARP_SLOT};
    const int64_t stride = (int64_t)gridDim.x * SG_WARPS;
    for (int64_t g = (int64_t)blockIdx.x * SG_WARPS + (threadIdx.x >> 5); g < n_samples; g += stride) {
        pack_sample(probs + g * 65536, ws.bmp, lane);
        __syncwarp();
        const int area = fill_and_keep_largest(ws.bmp, ws.row_base, tb, tb2, lane);
        if (area < 0) {
            if (lane == 0) overflow[1 + atomicAdd(&overflow[0], 1)] = (int32_t)g;
        } else {
            store_mask(ws.bmp, masks_out + g * CUB200_MASK_WORDS, lane);
            if (area_out && lane == 0) area_out[g] = area;
        }
        __syncwarp();
    }
}

// Samples with more than SG_CAP runs: same algorithm, run tables in the CTA's global-memory slot.
__global__ void __launch_bounds__(32)
seg_blob_overflow_kernel(const float* __restrict__ probs, uint32_t* __restrict__ masks_out, int32_t* __restrict__ area_out,
                         const int32_t* __restrict__ overflow, uint32_t* __restrict__ slots) {
    __shared__ __align__(16) uint32_t bmp[BMP_WORDS];
    __shared__ uint16_t row_base[264];
    const int lane = lane_id();
    uint32_t* slot = slots + (int64_t)blockIdx.x * 3 * SG_MAX_RUNS;
    const RunTables tb{slot, slot + SG_MAX_RUNS, slot + 2 * SG_MAX_RUNS, SG_MAX_RUNS};
    const int n = overflow[0];
    for (int q = blockIdx.x; q < n; q += gridDim.x) {
        const int64_t g = overflow[1 + q];
        pack_sample(probs + g * 65536, bmp, lane);
        __syncwarp();
        const int area = fill_and_keep_largest(bmp, row_base, tb, tb, lane);
        store_mask(bmp, masks_out + g * CUB200_MASK_WORDS, lane);
        if (area_out && lane == 0) area_out[g] = area;
        __syncwarp();
    }
}

// One warp per (image, row).  Per chunk of 32 samples the lanes first look at one sample each (the 32 bytes of its
// blob mask in this row): most (row, sample) pairs hold no blob pixel and are dismissed there, in one memory latency
// per chunk.  The remaining samples are taken in order, four at a time, lane = 8 consecutive pixels: the probabilities
// are read only under the lane's mask byte and summed in fp64.
__global__ void __launch_bounds__(256, 3)
seg_entropy_kernel(const float* __restrict__ probs, const uint32_t* __restrict__ masks, int64_t n_img, int n_samples, int pad,
                   float* __restrict__ entropy_out, float* __restrict__ mean_out, uint32_t* __restrict__ pred_samples_out,
                   float* __restrict__ samples_out) {
    const int lane = lane_id();
    const int64_t n_rows = n_img * 256;
    const int64_t stride = (int64_t)gridDim.x * 8;
    for (int64_t wr = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); wr < n_rows; wr += stride) {
        const int64_t img = wr >> 8;
        const int row = (int)(wr & 255);
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const uint32_t* mimg = masks + img * n_samples * (int64_t)CUB200_MASK_WORDS + row * 8;
        const uint8_t* mrow = reinterpret_cast<const uint8_t*>(mimg) + lane;
        const float* prow = probs + img * n_samples * 65536LL + row * 256 + lane * 8;
        uint32_t* oimg = pred_samples_out ? pred_samples_out + img * n_samples * (int64_t)CUB200_MASK_WORDS + row * 8 : nullptr;
        uint8_t* orow = oimg ? reinterpret_cast<uint8_t*>(oimg) + lane : nullptr;
        float* srow = samples_out ? samples_out + img * n_samples * 65536LL + row * 256 + lane * 8 : nullptr;
        for (int tc = 0; tc < n_samples; tc += 32) {
            const int ts = tc + lane;
            bool any = false;
            if (ts < n_samples) {
                const uint4* mr = reinterpret_cast<const uint4*>(mimg + (int64_t)ts * CUB200_MASK_WORDS);
                const uint4 a = __ldg(mr), b = __ldg(mr + 1);
                any = ((a.x | a.y | a.z | a.w) | (b.x | b.y | b.z | b.w)) != 0u;
                if (!any && oimg) {                   // nothing of this sample in the row: its pred_samples row is empty
                    uint4* o = reinterpret_cast<uint4*>(oimg + (int64_t)ts * CUB200_MASK_WORDS);
                    o[0] = o[1] = make_uint4(0, 0, 0, 0);
                }
            }
            unsigned act = __ballot_sync(FULL, any);
            if (srow) {
                unsigned idle = __ballot_sync(FULL, ts < n_samples && !any);
                while (idle) {
                    const int t = tc + __ffs(idle) - 1;
                    idle &= idle - 1;
                    float4* o = reinterpret_cast<float4*>(srow + (int64_t)t * 65536);
                    o[0] = o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            while (act) {
                int tq[4];
                uint32_t mb[4];
                float4 v[4][2];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tq[q] = act ? tc + __ffs(act) - 1 : -1;
                    act &= act - 1;                   // 0 & anything stays 0
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) mb[q] = (tq[q] >= 0) ? (uint32_t)__ldg(mrow + (int64_t)tq[q] * (CUB200_MASK_WORDS * 4)) : 0u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    v[q][0] = v[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (mb[q]) {
                        const float4* p = reinterpret_cast<const float4*>(prow + (int64_t)tq[q] * 65536);
                        v[q][0] = ld_stream_f4(p);
                        v[q][1] = ld_stream_f4(p + 1);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (tq[q] < 0) break;
                    uint32_t kept = 0u;
                    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (mb[q]) {
                        const float r[8] = {v[q][0].x, v[q][0].y, v[q][0].z, v[q][0].w, v[q][1].x, v[q][1].y, v[q][1].z, v[q][1].w};
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            x[k] = ((mb[q] >> k) & 1u) ? r[k] : 0.f;
                            acc[k] += (double)x[k];
                            kept |= (fabsf(x[k]) <= 0.5f) ? 0u : (1u << k);    // round(samples) != 0
                        }
                    }
                    if (orow) orow[(int64_t)tq[q] * (CUB200_MASK_WORDS * 4)] = (uint8_t)kept;
                    if (srow) {
                        float4* o = reinterpret_cast<float4*>(srow + (int64_t)tq[q] * 65536);
                        o[0] = make_float4(x[0], x[1], x[2], x[3]);
                        o[1] = make_float4(x[4], x[5], x[6], x[7]);
                    }
                }
            }
        }
        const bool framed = row < pad || row >= 256 - pad;
        float ent[8], mean[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int col = lane * 8 + k;
            // y_hat = mean (fp32 tensor), [y, 1 - y] normalised by their fp32 sum, entr() summed, / ln 2
            const float y = (float)(acc[k] / (double)n_samples);
            float h = 0.f;
            // most pixels are never inside a blob (y == 0: entropy 0) or lie in the zeroed frame: no fp64 logarithms there
            if (y != 0.f && !(framed || col < pad || col >= 256 - pad)) {
                const float y1 = 1.0f - y;
                const float sum = y + y1;
                const float pa = __fdiv_rn(y, sum), pb = __fdiv_rn(y1, sum);
                // scipy.special.entr in the array's own precision (fp32): -x log x for x > 0, 0 at 0, -inf below
                const float ea = pa > 0.f ? -pa * logf(pa) : (pa == 0.f ? 0.f : -INFINITY);
                const float eb = pb > 0.f ? -pb * logf(pb) : (pb == 0.f ? 0.f : -INFINITY);
                h = __fdiv_rn(ea + eb, 0.6931471805599453f);
                if (!isfinite(h)) h = 0.f;
            }
            ent[k] = h;
            mean[k] = y;
        }
        float4* eo = reinterpret_cast<float4*>(entropy_out + wr * 256 + lane * 8);
        eo[0] = make_float4(ent[0], ent[1], ent[2], ent[3]);
        eo[1] = make_float4(ent[4], ent[5], ent[6], ent[7]);
        if (mean_out) {
            float4* mo = reinterpret_cast<float4*>(mean_out + wr * 256 + lane * 8);
            mo[0] = make_float4(mean[0], mean[1], mean[2], mean[3]);
            mo[1] = make_float4(mean[4], mean[5], mean[6], mean[7]);
        }
    }
}

static int sm_count(int* out) {
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        CUB200_CUDA_OK(cudaGetDevice(&dev));
        CUB200_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    *out = num_sms;
    return 0;
}

}  // namespace cub200

// Workspace layout: overflow queue | run-table slots of the overflow kernel | per-warp slots of seg_blob_kernel, each
// region sized by the number of CTAs that can actually be launched for n_samples.
static int64_t seg_queue_bytes(int64_t n_samples) { return ((n_samples + 1) * 4 + 15) / 16 * 16; }
static int64_t seg_overflow_ctas(int64_t n_samples) { return n_samples < cub200::SG_OVERFLOW_CTAS ? n_samples : cub200::SG_OVERFLOW_CTAS; }
static int64_t seg_blob_ctas_max(int64_t n_samples) {
    const int64_t need = (n_samples + cub200::SG_WARPS - 1) / cub200::SG_WARPS;
    return need < cub200::SG_MAX_GRID ? need : cub200::SG_MAX_GRID;
}

extern "C" int64_t cub200_seg_blob_workspace_bytes(int64_t n_samples) {
    if (n_samples < 0) return CUB200_EINVAL;
    return seg_queue_bytes(n_samples) + seg_overflow_ctas(n_samples) * 3 * cub200::SG_MAX_RUNS * 4 +
           seg_blob_ctas_max(n_samples) * cub200::SG_WARPS * 3 * cub200::SG_WARP_SLOT * 4;
}

extern "C" int cub200_seg_largest_blob(const float* probs, int64_t n_samples, uint32_t* masks_out, int32_t* area_out,
                                       void* workspace, void* stream) {
    using namespace cub200;
    if (n_samples == 0) return 0;
    if (!probs || !masks_out || !workspace || n_samples < 0 || n_samples > 0x7ffffff0) return CUB200_EINVAL;
    if ((reinterpret_cast<uintptr_t>(probs) | reinterpret_cast<uintptr_t>(masks_out) | reinterpret_cast<uintptr_t>(workspace)) & 15)
        return CUB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = SG_WARPS * (int)sizeof(SegScratch);
    // per device, so set on every call (a process may drive several GPUs)
    CUB200_CUDA_OK(cudaFuncSetAttribute(seg_blob_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int sms = 0;
    if (int rc = sm_count(&sms)) return rc;
    int32_t* overflow = reinterpret_cast<int32_t*>(workspace);
    uint32_t* slots = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(workspace) + seg_queue_bytes(n_samples));
    uint32_t* warp_slots = slots + seg_overflow_ctas(n_samples) * 3 * SG_MAX_RUNS;
    CUB200_CUDA_OK(cudaMemsetAsync(overflow, 0, 4, st));
    const int64_t need = seg_blob_ctas_max(n_samples);
    const int64_t cap = (int64_t)sms * CUB200_SG_MINBLOCKS;
    seg_blob_kernel<<<(unsigned)(need < cap ? need : cap), SG_THREADS, smem, st>>>(probs, n_samples, masks_out, area_out, overflow,
                                                                                  warp_slots);
    if (int rc = check_launch("seg_blob_kernel")) return rc;
    seg_blob_overflow_kernel<<<(unsigned)seg_overflow_ctas(n_samples), 32, 0, st>>>(probs, masks_out, area_out, overflow, slots);
    return check_launch("seg_blob_overflow_kernel");
}

extern "C" int cub200_seg_entropy(const float* probs, const uint32_t* masks, int64_t n_img, int n_samples, int pad,
                                  float* entropy_out, float* mean_out, uint32_t* pred_samples_out, float* samples_out,
                                  void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!probs || !masks || !entropy_out || n_img < 0 || n_samples <= 0 || pad < 0 || pad > 128) return CUB200_EINVAL;
    if ((reinterpret_cast<uintptr_t>(probs) | reinterpret_cast<uintptr_t>(masks) | reinterpret_cast<uintptr_t>(entropy_out) |
         reinterpret_cast<uintptr_t>(mean_out) | reinterpret_cast<uintptr_t>(pred_samples_out) | reinterpret_cast<uintptr_t>(samples_out)) & 15)
        return CUB200_EINVAL;
    if (n_img > (int64_t)0x7fffffff / 32) return CUB200_EINVAL;
    int sms = 0;
    if (int rc = sm_count(&sms)) return rc;
    const int64_t need = n_img * 32;                 // 8 rows per CTA
    const int64_t cap = (int64_t)sms * 64;
    seg_entropy_kernel<<<(unsigned)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>(
        probs, masks, n_img, n_samples, pad, entropy_out, mean_out, pred_samples_out, samples_out);
    return check_launch("seg_entropy_kernel");
}

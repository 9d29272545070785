// K4  finalize_maps: difference map -> count -> mean / binary entropy / pred, plus FAC from areas.
//
// Replaces  pred = pred_samples.mean((1,2)).round()   contour_uncertainty/task/regression/aleatoric.py:111
//           UncertaintyTask.sample_entropy            contour_uncertainty/task/uncertainty.py:107-133
//           lv_FAC / FAC.compute per-sample part      contour_uncertainty/utils/clinical.py:11-29,
//                                                     contour_uncertainty/results/clinical/view.py:35-78
// The maps are functions of exact integer counts, so the entropy is a (T+1)-entry fp64 table.
// HBM-bound streaming kernel: one warp per row, 8 columns per lane, warp prefix sum.
#include "common.cuh"

namespace cub200 {

// scipy.stats.entropy([p, 1-p], base=2): pk /= sum(pk); S = sum(entr(pk)) / log(2); non-finite -> 0
__global__ void entropy_lut_kernel(double* __restrict__ lut, int64_t T) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c > T) return;
    const double p = (double)c / (double)T;
    const double q = 1.0 - p;
    const double s = p + q;
    const double pn = p / s, qn = q / s;
    const double ep = pn > 0.0 ? -pn * log(pn) : 0.0;
    const double eq = qn > 0.0 ? -qn * log(qn) : 0.0;
    double e = (ep + eq) / log(2.0);
    if (!isfinite(e)) e = 0.0;
    lut[c] = e;
    lut[(T + 1) + c] = p;                                  // mean = count / T
    lut[2 * (T + 1) + c] = p * (1.0 - p);                  // Bernoulli variance of the samples
}

// One warp per FM_ROWS consecutive rows, every access fully coalesced: in chunk k (64 columns) lane l owns columns
// 64 k + 2 l, + 1, so a warp-wide load / store covers one contiguous 256 B (int2), 512 B (double2) or 64 B (uchar2) segment.
// All row loads are issued before the first scan (FM_ROWS KiB in flight per warp); the row prefix sum is four 64-column
// warp scans chained by a running total.
constexpr int FM_ROWS = 4;
__global__ void __launch_bounds__(256)
finalize_maps_kernel(const int32_t* __restrict__ diff, int64_t n_rows, int64_t T, int32_t* __restrict__ count_out,
                     double* __restrict__ mean_out, double* __restrict__ entropy_out, uint8_t* __restrict__ pred_out,
                     const double* __restrict__ lut, double* __restrict__ variance_out, uint16_t* __restrict__ count16_out) {
    const int lane = threadIdx.x & 31;
    const int64_t row0 = (blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5)) * FM_ROWS;
    if (row0 >= n_rows) return;
    int2 v[FM_ROWS][4];
#pragma unroll
    for (int q = 0; q < FM_ROWS; ++q) {
        const int64_t row = min(row0 + q, n_rows - 1);
        const int2* d2 = reinterpret_cast<const int2*>(diff + row * 256) + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[q][k] = d2[32 * k];
    }
    const double* lm = lut + (T + 1);
    const double* lv = lut + 2 * (T + 1);
#pragma unroll
    for (int q = 0; q < FM_ROWS; ++q) {
        const int64_t row = row0 + q;
        if (row >= n_rows) break;
        int carry = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int a = v[q][k].x, b = a + v[q][k].y;      // inclusive within the lane's pair
            int incl = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const int excl = incl - b + carry;
            const int c0 = a + excl, c1 = b + excl;
            carry += __shfl_sync(FULL, incl, 31);
            const int64_t base = row * 256 + 64 * k + 2 * lane;
            if (count_out) *reinterpret_cast<int2*>(count_out + base) = make_int2(c0, c1);
            if (count16_out) *reinterpret_cast<uint32_t*>(count16_out + base) = ((uint32_t)c0 & 0xffffu) | ((uint32_t)c1 << 16);
            // the fp64 maps are functions of the count: three (T+1)-entry tables (entropy | mean | variance), L1-resident
            const int i0 = min(max(c0, 0), (int)T), i1 = min(max(c1, 0), (int)T);
            if (variance_out) *reinterpret_cast<double2*>(variance_out + base) = make_double2(lv[i0], lv[i1]);
            if (mean_out) *reinterpret_cast<double2*>(mean_out + base) = make_double2(lm[i0], lm[i1]);
            if (entropy_out) *reinterpret_cast<double2*>(entropy_out + base) = make_double2(lut[i0], lut[i1]);
            // round-half-even(count / T): 1 iff count / T > 0.5 (exactly 0.5 rounds to 0)
            if (pred_out) *reinterpret_cast<uchar2*>(pred_out + base) = make_uchar2((unsigned char)(2 * (int64_t)c0 > T), (unsigned char)(2 * (int64_t)c1 > T));
        }
    }
}

__global__ void fac_kernel(const int32_t* __restrict__ ed, const int32_t* __restrict__ es, int64_t n, double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = (double)ed[i], b = (double)es[i];
    double f = (a - b) / a;                              // clinical.py:29
    if (!(f >= 0.0 && f <= 1.0)) f = nan("");            // view.py: samples outside [0, 1] are rejected
    out[i] = f;
}

// soft masks of ContourUncertaintyTask.convert_to_mask(soft_mask=True) (cu/task/regression/contour_uncertainty.py:44-49):
// skimage.filters.gaussian(mask, sigma=(5, 5), truncate=1) -- which is scipy.ndimage.gaussian_filter(mode='nearest') on the
// float image: an 11-tap normalised Gaussian along axis 0, then along axis 1, each sum formed as scipy's symmetric
// correlate1d forms it (centre tap first, then the pairs from the outermost inwards) -- followed by the reference's
// (mask - min) / (max - min).  One CTA per mask; the weights come from the host (numpy's exp, as scipy computes them).
constexpr int SOFT_R = 5;
__global__ void __launch_bounds__(256)
soft_mask_kernel(const uint32_t* __restrict__ masks, const double* __restrict__ w /* w[0..5]: centre .. outermost */,
                 double* __restrict__ out) {
    __shared__ uint32_t bits[CUB200_MASK_WORDS];
    __shared__ double row[256];
    __shared__ double red[2][8];
    const int x = threadIdx.x;
    const uint32_t* m = masks + (int64_t)blockIdx.x * CUB200_MASK_WORDS;
    double* o = out + (int64_t)blockIdx.x * 65536;
    for (int i = x; i < CUB200_MASK_WORDS; i += 256) bits[i] = m[i];
    __syncthreads();
    double wk[SOFT_R + 1];
#pragma unroll
    for (int j = 0; j <= SOFT_R; ++j) wk[j] = w[j];
    auto px = [&](int r, int c) { return (double)((bits[r * 8 + (c >> 5)] >> (c & 31)) & 1u); };
    double vmin = 1e300, vmax = -1e300;
    for (int y = 0; y < 256; ++y) {
        // axis 0 (rows) for column x of row y, edge rows repeated (mode='nearest')
        double t = px(y, x) * wk[0];
#pragma unroll
        for (int j = SOFT_R; j >= 1; --j) t += (px(max(y - j, 0), x) + px(min(y + j, 255), x)) * wk[j];
        row[x] = t;
        __syncthreads();
        // axis 1 (columns) on the row just formed
        double v = row[x] * wk[0];
#pragma unroll
        for (int j = SOFT_R; j >= 1; --j) v += (row[max(x - j, 0)] + row[min(x + j, 255)]) * wk[j];
        o[y * 256 + x] = v;
        vmin = fmin(vmin, v); vmax = fmax(vmax, v);
        __syncthreads();
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { vmin = fmin(vmin, __shfl_xor_sync(FULL, vmin, s)); vmax = fmax(vmax, __shfl_xor_sync(FULL, vmax, s)); }
    if ((x & 31) == 0) { red[0][x >> 5] = vmin; red[1][x >> 5] = vmax; }
    __syncthreads();
    vmin = red[0][0]; vmax = red[1][0];
    for (int k = 1; k < 8; ++k) { vmin = fmin(vmin, red[0][k]); vmax = fmax(vmax, red[1][k]); }
    const double den = vmax - vmin;
    for (int y = 0; y < 256; ++y) o[y * 256 + x] = (o[y * 256 + x] - vmin) / den;   // 0 / 0 -> NaN for a constant mask, as numpy gives
}

// Aleatoric / epistemic split of Monte-Carlo metric distributions (results/clinical/utils.py:3-18 and the reject rule of
// FAC / GLS / Volume.compute, view.py:52-53, patient.py:192-193): one warp per item, values (T_e, T_a) fp64.
//   v < lo or v > hi (NaN compares false, as in numpy) -> rejected: the value becomes NaN
//   means_e = nanmean_a, stds_e = nanstd_a (ddof 0); mean = nanmean_e(means), epistemic = nanstd_e(means),
//   aleatoric = nanmean_e(stds), total = epistemic + aleatoric.  A row / item without a finite value gives NaN.
__device__ __forceinline__ double ae_value(const double* v, int64_t i, int apply_range, double lo, double hi, bool& rej) {
    double x = v[i];
    rej = apply_range && (x < lo || x > hi);
    return rej ? nan("") : x;
}

// nanmean over one pass (a row of t_a values) by the whole warp; cnt = number of finite values
__device__ __forceinline__ double ae_row_mean(const double* row, int t_a, int apply_range, double lo, double hi, int lane, int& cnt) {
    double s = 0.0;
    cnt = 0;
    for (int a = lane; a < t_a; a += 32) {
        bool rej;
        const double x = ae_value(row, a, apply_range, lo, hi, rej);
        if (x == x) { s += x; ++cnt; }
    }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
    return cnt ? s / (double)cnt : nan("");
}

__global__ void __launch_bounds__(128)
metric_ae_split_kernel(const double* __restrict__ values, int64_t n_items, int t_e, int t_a, int apply_range, double lo, double hi,
                       double* __restrict__ values_out, uint8_t* __restrict__ reject_out, double* __restrict__ stats_out,
                       int32_t* __restrict__ n_reject_out) {
    const int lane = threadIdx.x & 31;
    const int64_t item = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int64_t off = item * (int64_t)t_e * t_a;
    const double* v = values + off;
    double sum_m = 0.0, sum_s = 0.0;                     // over the passes that hold a finite value
    int n_m = 0, n_rej = 0;
    for (int e = 0; e < t_e; ++e) {
        const double* row = v + (int64_t)e * t_a;
        int cnt;
        const double m = ae_row_mean(row, t_a, apply_range, lo, hi, lane, cnt);
        double q = 0.0;
        for (int a = lane; a < t_a; a += 32) {
            bool rej;
            const double x = ae_value(row, a, apply_range, lo, hi, rej);
            n_rej += rej;
            if (values_out) values_out[off + (int64_t)e * t_a + a] = x;
            if (reject_out) reject_out[off + (int64_t)e * t_a + a] = rej ? 1 : 0;
            if (x == x) { const double d = x - m; q += d * d; }
        }
        q = warp_sum(q);
        if (cnt) { sum_m += m; sum_s += sqrt(q / (double)cnt); ++n_m; }
    }
    const double mean = n_m ? sum_m / (double)n_m : nan("");
    double q = 0.0;                                      // epistemic: std (ddof 0) of the per-pass means about their mean
    for (int e = 0; e < t_e; ++e) {
        int cnt;
        const double m = ae_row_mean(v + (int64_t)e * t_a, t_a, apply_range, lo, hi, lane, cnt);
        if (cnt) { const double d = m - mean; q += d * d; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_rej += __shfl_xor_sync(FULL, n_rej, o);
    if (lane == 0) {
        const double epi = n_m ? sqrt(q / (double)n_m) : nan("");
        const double ale = n_m ? sum_s / (double)n_m : nan("");
        stats_out[item * 4 + 0] = mean;
        stats_out[item * 4 + 1] = ale;
        stats_out[item * 4 + 2] = epi;
        stats_out[item * 4 + 3] = epi + ale;
        if (n_reject_out) n_reject_out[item] = n_rej;
    }
}

// Dataset-level accumulators of a per-sample integer metric (sum, sum of squares): exact in 64-bit integers, so the
// result does not depend on the order of the atomics; these are what the ranks all_reduce at the end of a step.
__global__ void __launch_bounds__(256)
metric_sums_kernel(const int32_t* __restrict__ v, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long s = 0, s2 = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long x = v[i];
        s += (unsigned long long)x;
        s2 += (unsigned long long)(x * x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(FULL, s, o); s2 += __shfl_xor_sync(FULL, s2, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, s); atomicAdd(out + 1, s2); }
}

// One launch for what a rank buffers per step for the epoch-end exchange: the per-sample metric values copied into a slot of
// a device ring + their exact int64 (sum, sum of squares).  CUB200_METRIC_PACK_PARTS CTAs, each copies one contiguous part and
// leaves ITS partial sums in its own four words of the slot's tail: no accumulator to zero, no atomics, no second pass -- the
// consumer adds the parts (it adds over the ranks anyway).  slot = *step_dev % n_slots (a device-side step counter: the launch
// is identical in every replay of a CUDA graph), or 0.
constexpr int MP_PARTS = CUB200_METRIC_PACK_PARTS;
__global__ void __launch_bounds__(1024)
metric_pack_kernel(const int32_t* __restrict__ v, int64_t n, int32_t* __restrict__ ring, int64_t n_slots,
                   const uint64_t* __restrict__ step_dev) {
    __shared__ unsigned long long red[2][32];
    const int64_t slot = step_dev ? (int64_t)(*step_dev % (uint64_t)n_slots) : 0;
    int32_t* out = ring + slot * (n + 4 * MP_PARTS);
    // part p = elements [p * per, min((p + 1) * per, n)), per a multiple of 4 so that 128-bit accesses stay aligned
    const int64_t per = (((n + MP_PARTS - 1) / MP_PARTS) + 3) & ~3LL;
    const int64_t lo = min((int64_t)blockIdx.x * per, n), hi = min(lo + per, n);
    unsigned long long s = 0, s2 = 0;
    auto acc = [&](int32_t x32) { const long long x = x32; s += (unsigned long long)x; s2 += (unsigned long long)(x * x); };
    int64_t i0 = lo;
    if (((reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const int4* v4 = reinterpret_cast<const int4*>(v);
        int4* o4 = reinterpret_cast<int4*>(out);
        const int64_t q_lo = lo >> 2, q_hi = hi >> 2;                 // whole quads of the part (lo is a multiple of 4)
        for (int64_t base = q_lo; base < q_hi; base += 4 * 1024) {
            int4 t[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int64_t i = base + q * 1024 + threadIdx.x; t[q] = i < q_hi ? v4[i] : make_int4(0, 0, 0, 0); }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t i = base + q * 1024 + threadIdx.x;
                if (i < q_hi) { o4[i] = t[q]; acc(t[q].x); acc(t[q].y); acc(t[q].z); acc(t[q].w); }
            }
        }
        i0 = q_hi << 2;
    }
    for (int64_t i = i0 + threadIdx.x; i < hi; i += 1024) {
        const int32_t x32 = v[i];
        out[i] = x32;
        acc(x32);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(FULL, s, o); s2 += __shfl_xor_sync(FULL, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x < 32) {
        s = red[0][threadIdx.x]; s2 = red[1][threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(FULL, s, o); s2 += __shfl_xor_sync(FULL, s2, o); }
        if (threadIdx.x == 0) {
            int32_t* tail = out + n + 4 * blockIdx.x;
            tail[0] = (int32_t)(uint32_t)s; tail[1] = (int32_t)(uint32_t)(s >> 32);
            tail[2] = (int32_t)(uint32_t)s2; tail[3] = (int32_t)(uint32_t)(s2 >> 32);
        }
    }
}

// post_mu / post_cov of aleatoric.py:96-102: thread (k, sg) streams the float2 of landmark k over the
// samples sg, sg+12, ...; moments about the first sample (no cancellation), fp64, block reduction.
constexpr int SM_GROUPS = 12;
__global__ void __launch_bounds__(CUB200_K * SM_GROUPS)
sample_moments_kernel(const float* __restrict__ contours, int64_t T, double* __restrict__ mean_out, double* __restrict__ cov_out) {
    __shared__ double red[SM_GROUPS][CUB200_K][5];
    const int k = threadIdx.x % CUB200_K, sg = threadIdx.x / CUB200_K;
    const float2* base = reinterpret_cast<const float2*>(contours) + (int64_t)blockIdx.x * T * CUB200_K + k;
    const float2 c = base[0];
    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
    for (int64_t t = sg; t < T; t += SM_GROUPS) {
        const float2 p = base[t * CUB200_K];
        const double dx = (double)p.x - (double)c.x, dy = (double)p.y - (double)c.y;
        sx += dx; sy += dy; sxx += dx * dx; syy += dy * dy; sxy += dx * dy;
    }
    red[sg][k][0] = sx; red[sg][k][1] = sy; red[sg][k][2] = sxx; red[sg][k][3] = syy; red[sg][k][4] = sxy;
    __syncthreads();
    if (sg == 0) {
        double a[5] = {0, 0, 0, 0, 0};
        for (int g = 0; g < SM_GROUPS; ++g)
            for (int q = 0; q < 5; ++q) a[q] += red[g][k][q];
        const double n = (double)T, mx = a[0] / n, my = a[1] / n;
        const int64_t o = (int64_t)blockIdx.x * CUB200_K + k;
        mean_out[o * 2 + 0] = (double)c.x + mx;
        mean_out[o * 2 + 1] = (double)c.y + my;
        const double den = n - 1.0;                       // np.cov default ddof = 1
        const double cxx = (a[2] - n * mx * mx) / den, cyy = (a[3] - n * my * my) / den, cxy = (a[4] - n * mx * my) / den;
        cov_out[o * 4 + 0] = cxx; cov_out[o * 4 + 1] = cxy; cov_out[o * 4 + 2] = cxy; cov_out[o * 4 + 3] = cyy;
    }
}

}  // namespace cub200

extern "C" int cub200_sample_moments(const float* contours, int64_t n_img, int64_t n_samples, double* mean_out,
                                     double* cov_out, void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!contours || !mean_out || !cov_out || n_img < 0 || n_samples < 1 || n_img > 0x7fffffff) return CUB200_EINVAL;
    sample_moments_kernel<<<(unsigned)n_img, CUB200_K * SM_GROUPS, 0, (cudaStream_t)stream>>>(contours, n_samples, mean_out, cov_out);
    return check_launch("sample_moments_kernel");
}

extern "C" int cub200_count_tables(int64_t n_samples, double* lut_ws, void* stream) {
    using namespace cub200;
    if (!lut_ws || n_samples < 1 || n_samples > 0x3fffffff) return CUB200_EINVAL;
    entropy_lut_kernel<<<(unsigned)((n_samples + 1 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lut_ws, n_samples);
    return check_launch("entropy_lut_kernel");
}

extern "C" int cub200_finalize_maps(const int32_t* diff, int64_t n_img, int64_t n_samples, int32_t* count_out,
                                    double* mean_out, double* entropy_out, uint8_t* pred_out, double* lut_ws, int lut_ready,
                                    double* variance_out, uint16_t* count16_out, void* stream) {
    using namespace cub200;
    if (!diff || n_img < 0 || n_samples < 1 || n_samples > 0x3fffffff) return CUB200_EINVAL;
    if (count16_out && n_samples > 65535) return CUB200_EINVAL;
    if ((entropy_out || mean_out || variance_out) && !lut_ws) return CUB200_EINVAL;
    if (n_img == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if ((entropy_out || mean_out || variance_out) && !lut_ready) {
        entropy_lut_kernel<<<(unsigned)((n_samples + 1 + 255) / 256), 256, 0, st>>>(lut_ws, n_samples);
        int rc = check_launch("entropy_lut_kernel");
        if (rc) return rc;
    }
    const int64_t rows = n_img * 256;
    finalize_maps_kernel<<<(unsigned)((rows + 8 * FM_ROWS - 1) / (8 * FM_ROWS)), 256, 0, st>>>(diff, rows, n_samples, count_out, mean_out,
                                                                    entropy_out, pred_out, lut_ws, variance_out, count16_out);
    return check_launch("finalize_maps_kernel");
}

extern "C" int cub200_soft_masks(const uint32_t* masks, int64_t n, const double* weights, double* out, void* stream) {
    using namespace cub200;
    if (n == 0) return 0;
    if (!masks || !weights || !out || n < 0 || n > 0x7fffffff) return CUB200_EINVAL;
    soft_mask_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(masks, weights, out);
    return check_launch("soft_mask_kernel");
}

extern "C" int cub200_metric_ae_split(const double* values, int64_t n_items, int64_t t_e, int64_t t_a, int apply_range, double lo,
                                      double hi, double* values_out, uint8_t* reject_out, double* stats_out, int32_t* n_reject_out,
                                      void* stream) {
    using namespace cub200;
    if (!values || !stats_out || n_items < 0 || t_e < 1 || t_a < 1 || t_e > 0x7fffffff || t_a > 0x7fffffff) return CUB200_EINVAL;
    if (n_items == 0) return 0;
    metric_ae_split_kernel<<<(unsigned)((n_items + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        values, n_items, (int)t_e, (int)t_a, apply_range, lo, hi, values_out, reject_out, stats_out, n_reject_out);
    return check_launch("metric_ae_split_kernel");
}

extern "C" int cub200_fac_from_areas(const int32_t* area_ed, const int32_t* area_es, int64_t n, double* fac_out,
                                     void* stream) {
    using namespace cub200;
    if (!area_ed || !area_es || !fac_out || n < 0) return CUB200_EINVAL;
    if (n == 0) return 0;
    fac_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(area_ed, area_es, n, fac_out);
    return check_launch("fac_kernel");
}

extern "C" int cub200_metric_sums(const int32_t* values, int64_t n, int64_t* sums_out, void* stream) {
    using namespace cub200;
    if (!sums_out || n < 0 || (n > 0 && !values)) return CUB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    CUB200_CUDA_OK(cudaMemsetAsync(sums_out, 0, 2 * sizeof(int64_t), st));
    if (n == 0) return 0;
    const int64_t blocks = (n + 255) / 256;
    metric_sums_kernel<<<(unsigned)(blocks < 592 ? blocks : 592), 256, 0, st>>>(values, n, reinterpret_cast<unsigned long long*>(sums_out));
    return check_launch("metric_sums_kernel");
}

extern "C" int cub200_metric_pack(const int32_t* values, int64_t n, int32_t* ring, int64_t n_slots, const uint64_t* step_dev,
                                  void* stream) {
    using namespace cub200;
    if (!values || !ring || n < 0 || n_slots < 1) return CUB200_EINVAL;
    metric_pack_kernel<<<MP_PARTS, 1024, 0, (cudaStream_t)stream>>>(values, n, ring, n_slots, step_dev);
    return check_launch("metric_pack_kernel");
}

// ---------------------------------------------------------------- sample_entropy for probabilities of any class count
// UncertaintyTask.sample_entropy (contour_uncertainty/task/uncertainty.py:107-133) in full: samples (T, C, P) fp32 per
// image -> y_hat = mean over T; C == 1: entropy of [y, 1 - y] in bits, else entropy over the C classes in base C;
// scipy.stats.entropy normalises by the class sum, sums entr() and divides by log(base), all in the array's fp32;
// non-finite -> 0.  HBM-bound: every probability is read once, thread = pixel, coalesced along the pixels.
namespace cub200 {
constexpr int SE_MAXC = 8;

// entropy of one pixel from the class sums over the samples
template <int CMAX>
__device__ __forceinline__ float class_entropy(const double (&acc)[CMAX], int T, int C) {
    constexpr int NY = CMAX < 2 ? 2 : CMAX;
    float y[NY];
#pragma unroll
    for (int c = 0; c < NY; ++c) y[c] = (c < CMAX) ? (float)(acc[c < CMAX ? c : 0] / (double)T) : 0.f;
    int nc = C;
    float lb = logf((float)C);
    if (C == 1) { y[1] = 1.0f - y[0]; nc = 2; lb = 0.6931471805599453f; }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NY; ++c) if (c < nc) sum += y[c];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NY; ++c)
        if (c < nc) {
            const float q = __fdiv_rn(y[c], sum);
            s += q > 0.f ? -q * logf(q) : (q == 0.f ? 0.f : -INFINITY);
        }
    const float h = __fdiv_rn(s, lb);
    return isfinite(h) ? h : 0.f;
}

// VEC pixels per thread (4: 128-bit loads; needs n_pixels % 4 == 0 and 16-byte aligned maps), classes held in CMAX
// register slots, UNR samples in flight per thread -- the bytes in flight are what a streaming kernel lives on.
template <int VEC, int CMAX, int UNR>
__global__ void __launch_bounds__(256)
sample_entropy_kernel(const float* __restrict__ samples, int64_t n_img, int T, int C, int64_t P, float* __restrict__ out) {
    const int64_t groups = P / VEC, total = n_img * groups;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = i / groups, p = (i - img * groups) * VEC;
        const float* base = samples + img * T * C * P + p;
        double acc[VEC][CMAX];
#pragma unroll
        for (int k = 0; k < VEC; ++k)
#pragma unroll
            for (int c = 0; c < CMAX; ++c) acc[k][c] = 0.0;
        auto load = [&](int t, int c, float (&v)[VEC]) {
            const float* q = base + ((int64_t)t * C + c) * P;
            if constexpr (VEC == 4) {
                const float4 f = ld_stream_f4(reinterpret_cast<const float4*>(q));
                v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
            } else {
                v[0] = __ldcs(q);
            }
        };
        int t = 0;
        for (; t + UNR <= T; t += UNR) {
            float v[UNR][CMAX][VEC];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if (c < C) load(t + u, c, v[u][c]);
                    else {
#pragma unroll
                        for (int k = 0; k < VEC; ++k) v[u][c][k] = 0.f;
                    }
                }
#pragma unroll
            for (int u = 0; u < UNR; ++u)                    // sample order kept
#pragma unroll
                for (int c = 0; c < CMAX; ++c)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k][c] += (double)v[u][c][k];
        }
        for (; t < T; ++t) {
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (c < C) {
                    float v[VEC];
                    load(t, c, v);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k][c] += (double)v[k];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) out[img * P + p + k] = class_entropy<CMAX>(acc[k], T, C);
    }
}

template <int VEC>
static void launch_sample_entropy(const float* samples, int64_t n_img, int T, int C, int64_t P, float* out, cudaStream_t st) {
    const int64_t total = n_img * (P / VEC);
    const int64_t blocks = (total + 255) / 256;
    const unsigned grid = (unsigned)(blocks < 148 * 32 ? blocks : 148 * 32);
    if (C == 1) sample_entropy_kernel<VEC, 1, 4><<<grid, 256, 0, st>>>(samples, n_img, T, C, P, out);
    else if (C == 2) sample_entropy_kernel<VEC, 2, 4><<<grid, 256, 0, st>>>(samples, n_img, T, C, P, out);
    else if (C == 3) sample_entropy_kernel<VEC, 3, 4><<<grid, 256, 0, st>>>(samples, n_img, T, C, P, out);
    else if (C == 4) sample_entropy_kernel<VEC, 4, 2><<<grid, 256, 0, st>>>(samples, n_img, T, C, P, out);
    else sample_entropy_kernel<VEC, SE_MAXC, 1><<<grid, 256, 0, st>>>(samples, n_img, T, C, P, out);
}
}  // namespace cub200

extern "C" int cub200_sample_entropy(const float* samples, int64_t n_img, int n_samples, int n_classes, int64_t n_pixels,
                                     float* entropy_out, void* stream) {
    using namespace cub200;
    if (n_img == 0 || n_pixels == 0) return 0;
    if (!samples || !entropy_out || n_img < 0 || n_pixels < 0 || n_samples < 1 || n_classes < 1 || n_classes > SE_MAXC)
        return CUB200_EINVAL;
    const bool vec = (n_pixels % 4 == 0) && !((reinterpret_cast<uintptr_t>(samples) | reinterpret_cast<uintptr_t>(entropy_out)) & 15);
    if (vec) launch_sample_entropy<4>(samples, n_img, n_samples, n_classes, n_pixels, entropy_out, (cudaStream_t)stream);
    else launch_sample_entropy<1>(samples, n_img, n_samples, n_classes, n_pixels, entropy_out, (cudaStream_t)stream);
    return check_launch("sample_entropy_kernel");
}

// Dataset-level pixel calibration: binned accumulators of (confidence, accuracy) over every foreground pixel.
//
// Replaces (reference, cu/ = contour_uncertainty/):
//   Calibration.get_pixel_confidences_and_accuracies / get_confidences_and_accuracies   cu/results/calibration.py:197-250
//   Calibration.ece (the bin loop)                                                       cu/results/calibration.py:94-113
// confidence = 1 - uncertainty map; accuracy = (pred == gt); pixels with pred + gt == 0 (background in both) are
// dropped; a pixel belongs to bin b iff  lower_b < confidence < upper_b  (both strict: a confidence that equals a
// boundary belongs to no bin, as in the reference).  HBM-bound streaming reduction: 10 B read per pixel.
// Counts and accuracy sums are integers (exact in fp64 whatever the order); the confidence sums go through shared-memory
// fp64 atomics per warp and a fixed-order reduction over warps and blocks, so they can differ from run to run in the
// last bits only (<= 1e-15 relative).
#include "common.cuh"

namespace cub200 {

constexpr int CAL_THREADS = 256;
constexpr int CAL_BLOCKS = 592;          // 4 x 148 SMs
constexpr int CAL_MAX_BINS = 64;

__global__ void __launch_bounds__(CAL_THREADS)
calibration_partial_kernel(const double* __restrict__ umap, const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                           int64_t n, const double* __restrict__ bounds, int nb, double* __restrict__ partial) {
    __shared__ double s_b[CAL_MAX_BINS + 1];
    __shared__ double s_acc[CAL_THREADS / 32][CAL_MAX_BINS + 1][3];
    for (int i = threadIdx.x; i <= nb; i += CAL_THREADS) s_b[i] = bounds[i];
    for (int i = threadIdx.x; i < (CAL_THREADS / 32) * (CAL_MAX_BINS + 1) * 3; i += CAL_THREADS) (&s_acc[0][0][0])[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double lo = s_b[0], hi = s_b[nb], inv = (double)nb / (hi - lo);
    // per-thread accumulation for the bin it is currently in (runs of equal bins are common), flushed on change
    int cur = -1;
    double c_n = 0.0, c_conf = 0.0, c_acc = 0.0, tot = 0.0;
    auto flush = [&]() {
        if (cur >= 0) {      // lanes of a warp may hit the same bin: shared-memory atomics
            atomicAdd(&s_acc[warp][cur][0], c_n);
            atomicAdd(&s_acc[warp][cur][1], c_conf);
            atomicAdd(&s_acc[warp][cur][2], c_acc);
        }
        c_n = c_conf = c_acc = 0.0;
    };
    (void)lane;
    for (int64_t i = (int64_t)blockIdx.x * CAL_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * CAL_THREADS) {
        const int p = pred[i], g = gt[i];
        if (p + g == 0) continue;
        tot += 1.0;
        const double conf = 1.0 - umap[i];
        int b = (int)floor((conf - lo) * inv);
        b = min(max(b, 0), nb - 1);
        // exact membership against the boundary values themselves
        if (b > 0 && !(conf > s_b[b])) --b;
        else if (b < nb - 1 && !(conf < s_b[b + 1])) ++b;
        if (!(conf > s_b[b] && conf < s_b[b + 1])) continue;
        if (b != cur) { flush(); cur = b; }
        c_n += 1.0;
        c_conf += conf;
        c_acc += (p == g) ? 1.0 : 0.0;
    }
    flush();
    atomicAdd(&s_acc[warp][nb][0], tot);
    __syncthreads();
    for (int i = threadIdx.x; i < (nb + 1) * 3; i += CAL_THREADS) {
        double a = 0.0;
        for (int w = 0; w < CAL_THREADS / 32; ++w) a += s_acc[w][i / 3][i % 3];
        partial[(int64_t)blockIdx.x * (CAL_MAX_BINS + 1) * 3 + i] = a;
    }
}

__global__ void calibration_final_kernel(const double* __restrict__ partial, int n_blocks, int nb, double* __restrict__ out) {
    const int i = threadIdx.x;
    if (i >= (nb + 1) * 3) return;
    double a = 0.0;
    for (int b = 0; b < n_blocks; ++b) a += partial[(int64_t)b * (CAL_MAX_BINS + 1) * 3 + i];
    out[i] = a;
}

}  // namespace cub200

extern "C" int64_t cub200_calibration_workspace_bytes(void) {
    return (int64_t)cub200::CAL_BLOCKS * (cub200::CAL_MAX_BINS + 1) * 3 * sizeof(double);
}

extern "C" int cub200_calibration_bins(const double* umap, const uint8_t* pred, const uint8_t* gt, int64_t n_pixels,
                                       const double* bounds, int nb_bins, double* bins_out, void* workspace, void* stream) {
    using namespace cub200;
    if (!bounds || !bins_out || !workspace || nb_bins < 1 || nb_bins > CAL_MAX_BINS || n_pixels < 0) return CUB200_EINVAL;
    if (n_pixels > 0 && (!umap || !pred || !gt)) return CUB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    double* partial = reinterpret_cast<double*>(workspace);
    calibration_partial_kernel<<<CAL_BLOCKS, CAL_THREADS, 0, st>>>(umap, pred, gt, n_pixels, bounds, nb_bins, partial);
    int rc = check_launch("calibration_partial_kernel");
    if (rc) return rc;
    calibration_final_kernel<<<1, 256, 0, st>>>(partial, CAL_BLOCKS, nb_bins, bins_out);
    return check_launch("calibration_final_kernel");
}

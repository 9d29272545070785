// Sequence (ED/ES-coupled) sampler: conditioning of the second instant on the sampled first one.
//
// Replaces SequencePSMSampler.sample_two_contours lines 68-91
// (contour_uncertainty/sampler/posterior_shape_model/sequence_sampler.py): the 84-dim sequence PSM is
// used AS STORED (seq_mu, seq_Q fixed; the per-call pca is commented out in the reference), so
//     mu_c(second) = seq_mu_o + M_b (transform(s_first) - seq_mu_g),   M_b = S[o,g] (S[g,g] + I)^-1
//     cov_c(second) = diag 2x2 blocks of S[o,o] - M_b S[g,o]   (times seq_scale, columns)
// are per-model constants for each choice b of the first instant (host, fp64, once per model).
// Per (pair, sample) this kernel does the 42x42 mat-vec and the product-of-Gaussians merge
// (psm.py:424-440) with the second instant's predicted distribution, and emits (mu_f, cov_f) of the
// second instant -- the input of a second PosteriorShapeModelSampler call (sequence_sampler.py:94).
// One warp per (pair, sample); lane k = landmark k.
#include "common.cuh"

namespace cub200 {

__global__ void __launch_bounds__(256)
seq_condition_kernel(const float* __restrict__ mu, const float* __restrict__ cov, const float* __restrict__ first,
                     const uint8_t* __restrict__ bits, int64_t n_pairs, int64_t T,
                     const double* __restrict__ m2t,      // (2, 42, 42): m2t[b][e][row]
                     const double* __restrict__ covc,     // (2, 21, 2, 2)
                     const double* __restrict__ seq_mu, const double* __restrict__ seq_mean,
                     const double* __restrict__ seq_scale,  // (84) each
                     float* __restrict__ first_out, float* __restrict__ mu_f, float* __restrict__ cov_f) {
    __shared__ double dsh[8][44];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * 8 + warp;      // (pair, sample)
    if (g >= n_pairs * T) return;
    const int64_t p = g / T, t = g % T;
    const int b = bits[g] ? 1 : 0, o = 1 - b;
    constexpr int K = CUB200_K;
    float2 s1 = make_float2(0.f, 0.f);
    if (lane < K) {
        // first-sampled instant: the contour drawn for image 2p + b (both instants were drawn upstream)
        s1 = reinterpret_cast<const float2*>(first)[((2 * p + b) * T + t) * K + lane];
        reinterpret_cast<float2*>(first_out)[g * K + lane] = s1;
        const int f0 = 42 * b + 2 * lane;                  // flat index in the 84-dim model
        dsh[warp][2 * lane] = ((double)s1.x - seq_mean[f0]) / seq_scale[f0] - seq_mu[f0];
        dsh[warp][2 * lane + 1] = ((double)s1.y - seq_mean[f0 + 1]) / seq_scale[f0 + 1] - seq_mu[f0 + 1];
    }
    __syncwarp();
    if (lane < K) {
        const int o0 = 42 * o + 2 * lane;
        double mx = seq_mu[o0], my = seq_mu[o0 + 1];
        const double2* mrow = reinterpret_cast<const double2*>(m2t + (int64_t)b * 42 * 42) + lane;
#pragma unroll 6
        for (int e = 0; e < 42; ++e) {
            const double2 m = mrow[e * 21];
            const double d = dsh[warp][e];
            mx = fma(m.x, d, mx);
            my = fma(m.y, d, my);
        }
        // inverse transform                                           sequence_sampler.py:81
        mx = mx * seq_scale[o0] + seq_mean[o0];
        my = my * seq_scale[o0 + 1] + seq_mean[o0 + 1];
        // merge_priors(mu, cov, mu_c, cov_c)                          psm.py:436-437
        const float* cp = cov + ((2 * p + o) * K + lane) * 4;
        const double p00 = cp[0], p01 = cp[1], p10 = cp[2], p11 = cp[3];
        const double* cc = covc + ((int64_t)b * K + lane) * 4;
        const double c00 = cc[0], c01 = cc[1], c10 = cc[2], c11 = cc[3];
        const double t00 = p00 + c00, t01 = p01 + c01, t10 = p10 + c10, t11 = p11 + c11;
        const double idet = 1.0 / (t00 * t11 - t01 * t10);
        const double i00 = t11 * idet, i01 = -t01 * idet, i10 = -t10 * idet, i11 = t00 * idet;
        const double g00 = p00 * i00 + p01 * i10, g01 = p00 * i01 + p01 * i11;   // cov_p inv
        const double g10 = p10 * i00 + p11 * i10, g11 = p10 * i01 + p11 * i11;
        const double h00 = c00 * i00 + c01 * i10, h01 = c00 * i01 + c01 * i11;   // cov_c inv
        const double h10 = c10 * i00 + c11 * i10, h11 = c10 * i01 + c11 * i11;
        const float2 mp = reinterpret_cast<const float2*>(mu)[(2 * p + o) * K + lane];
        const double fx = g00 * mx + g01 * my + h00 * (double)mp.x + h01 * (double)mp.y;
        const double fy = g10 * mx + g11 * my + h10 * (double)mp.x + h11 * (double)mp.y;
        reinterpret_cast<float2*>(mu_f)[g * K + lane] = make_float2((float)fx, (float)fy);
        float4 cf;                                          // sigma_f = cov_p inv cov_c
        cf.x = (float)(g00 * c00 + g01 * c10); cf.y = (float)(g00 * c01 + g01 * c11);
        cf.z = (float)(g10 * c00 + g11 * c10); cf.w = (float)(g10 * c01 + g11 * c11);
        reinterpret_cast<float4*>(cov_f)[g * K + lane] = cf;
    }
}

}  // namespace cub200

extern "C" int cub200_seq_condition(const float* mu, const float* cov, const float* first_samples, const uint8_t* first_instant,
                                    int64_t n_pairs, int64_t n_samples, const double* m2t, const double* covc,
                                    const double* seq_mu, const double* seq_mean, const double* seq_scale,
                                    float* first_out, float* mu_f_out, float* cov_f_out, void* stream) {
    using namespace cub200;
    const int64_t total = n_pairs * n_samples;
    if (total == 0) return 0;
    if (!mu || !cov || !first_samples || !first_instant || !m2t || !covc || !seq_mu || !seq_mean || !seq_scale ||
        !first_out || !mu_f_out || !cov_f_out || n_pairs < 0 || n_samples < 0)
        return CUB200_EINVAL;
    seq_condition_kernel<<<(unsigned)((total + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        mu, cov, first_samples, first_instant, n_pairs, n_samples, m2t, covc, seq_mu, seq_mean, seq_scale, first_out,
        mu_f_out, cov_f_out);
    return check_launch("seq_condition_kernel");
}

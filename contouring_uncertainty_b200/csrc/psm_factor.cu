// K2  psm_factor: the affine map of the hierarchical posterior-shape-model sampler, per image.
//
// The reference re-runs a 42x42 eigendecomposition per image and two 42x42 fp32 inverses per level
// per SAMPLE (contour_uncertainty/sampler/posterior_shape_model/psm.py:91,265-352;
// posteriorshapemodel.py:9-81).  The sampler is exactly affine in its standard-normal draws:
//     contour = mu + A z,   z in R^34,  A in R^(42x34)            
// so everything is hoisted to once per image and done eig-free with S = Q Q^T = C0 + d d^T,
// d = xbar - transform(mu):
//     mu_c  = mu + S[:,g] (S[g,g] + s2 I)^-1 (s_g - mu_g)
//     Sig_c = S - S[:,g] (S[g,g] + s2 I)^-1 S[g,:]                (only 2x2 diagonal blocks used)
//     merge_priors: G = cov_p (cov_p + cov_c)^-1, cov_f = G cov_c, mu_f - mu = G (mu_c - mu)
//     draw: mu_f + chol(cov_f) z_slot
// One CTA per image, fp64 throughout (34x34 Cholesky at most), result stored fp32 transposed
// (draw-major) for the coalesced reads of the fused sampler.
#include "common.cuh"

namespace cub200 {

constexpr int PF_THREADS = 128;
constexpr int P = CUB200_P;   // 42
constexpr int Z = 34;         // draws of the PSM sampler

__constant__ int c_initial[3] = {0, 10, 20};
__constant__ int c_level_pts[3][8] = {{5, 15, 0, 0, 0, 0, 0, 0}, {2, 7, 13, 18, 0, 0, 0, 0}, {1, 3, 6, 8, 12, 14, 17, 19}};
__constant__ int c_level_n[3] = {2, 4, 8};
__constant__ int c_rest[4] = {4, 9, 11, 16};
// draw slot of each landmark (order 0,10,20,5,15,2,7,13,18,1,3,6,8,12,14,17,19); -1 = filled
__constant__ int c_slot[21] = {0, 9, 5, 10, -1, 3, 11, 6, 12, -1, 1, -1, 13, 7, 14, 4, -1, 15, 8, 16, 2};

struct PFShared {
    double S[P][P];
    double A[P][Z];
    double G[Z][Z];
    double Mt[Z][16];   // solution columns: Mt[q][c], c = 2*(point index in level) + component
    double scale[P];
    double d[P];
    int g[Z];
};

// In-place lower Cholesky of s.G[0..n)[0..n)
__device__ void cholesky(PFShared& s, int n) {
    for (int k = 0; k < n; ++k) {
        if (threadIdx.x == 0) {
            double v = s.G[k][k];
            for (int p = 0; p < k; ++p) v -= s.G[k][p] * s.G[k][p];
            s.G[k][k] = sqrt(v);
        }
        __syncthreads();
        for (int i = k + 1 + threadIdx.x; i < n; i += PF_THREADS) {
            double v = s.G[i][k];
            for (int p = 0; p < k; ++p) v -= s.G[i][p] * s.G[k][p];
            s.G[i][k] = v / s.G[k][k];
        }
        __syncthreads();
    }
}

// One level: condition on the ng observed flat indices s.g, produce the rows of A for `npts` points.
__device__ void psm_level(PFShared& s, int ng, const int* pts, int npts, double sigma2, bool merge,
                          const float* __restrict__ cov_img) {
    const int tid = threadIdx.x;
    for (int i = tid; i < ng * ng; i += PF_THREADS) {
        const int a = i / ng, b = i % ng;
        s.G[a][b] = s.S[s.g[a]][s.g[b]] + (a == b ? sigma2 : 0.0);
    }
    __syncthreads();
    cholesky(s, ng);
    const int ncol = 2 * npts;
    if (tid < ncol) {  // G x = S[g, row]  ->  x = M[row, :]
        const int row = 2 * pts[tid >> 1] + (tid & 1);
        for (int i = 0; i < ng; ++i) {
            double v = s.S[s.g[i]][row];
            for (int p = 0; p < i; ++p) v -= s.G[i][p] * s.Mt[p][tid];
            s.Mt[i][tid] = v / s.G[i][i];
        }
        for (int i = ng - 1; i >= 0; --i) {
            double v = s.Mt[i][tid];
            for (int p = i + 1; p < ng; ++p) v -= s.G[p][i] * s.Mt[p][tid];
            s.Mt[i][tid] = v / s.G[i][i];
        }
    }
    __syncthreads();
    // B[c][e] = scale[row_c] * sum_q Mt[q][c] / scale[g[q]] * A[g[q]][e]   -> staged in G (free now)
    double (*B)[Z] = reinterpret_cast<double (*)[Z]>(&s.G[0][0]);  // needs ncol <= 34 rows: ok (<=16)
    for (int i = tid; i < ncol * Z; i += PF_THREADS) {
        const int c = i / Z, e = i % Z;
        const int row = 2 * pts[c >> 1] + (c & 1);
        double v = 0.0;
        for (int q = 0; q < ng; ++q) v += s.Mt[q][c] / s.scale[s.g[q]] * s.A[s.g[q]][e];
        B[c][e] = v * s.scale[row];
    }
    __syncthreads();
    // per point: merge with the predicted distribution and add the coloured draw
    for (int i = tid; i < npts * Z; i += PF_THREADS) {
        const int pi = i / Z, e = i % Z;
        const int j = pts[pi];
        if (!merge) {
            s.A[2 * j][e] = B[2 * pi][e];
            s.A[2 * j + 1][e] = B[2 * pi + 1][e];
            continue;
        }
        // Sig_c diagonal block, then cov_c = Sig_c * scale (columns)        psm.py:269-272
        double cc[2][2];
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                double v = s.S[2 * j + a][2 * j + b];
                for (int q = 0; q < ng; ++q) v -= s.Mt[q][2 * pi + a] * s.S[s.g[q]][2 * j + b];
                cc[a][b] = v * s.scale[2 * j + b];
            }
        const double p00 = cov_img[4 * j + 0], p01 = cov_img[4 * j + 1], p10 = cov_img[4 * j + 2], p11 = cov_img[4 * j + 3];
        const double t00 = p00 + cc[0][0], t01 = p01 + cc[0][1], t10 = p10 + cc[1][0], t11 = p11 + cc[1][1];
        const double idet = 1.0 / (t00 * t11 - t01 * t10);
        const double i00 = t11 * idet, i01 = -t01 * idet, i10 = -t10 * idet, i11 = t00 * idet;
        // Gm = cov_p inv(T)                                                psm.py:436-437
        const double g00 = p00 * i00 + p01 * i10, g01 = p00 * i01 + p01 * i11;
        const double g10 = p10 * i00 + p11 * i10, g11 = p10 * i01 + p11 * i11;
        double a0 = g00 * B[2 * pi][e] + g01 * B[2 * pi + 1][e];
        double a1 = g10 * B[2 * pi][e] + g11 * B[2 * pi + 1][e];
        const int sl = c_slot[j];
        if (e == 2 * sl || e == 2 * sl + 1) {
            // cov_f = Gm cov_c ; lower Cholesky reading the lower triangle (torch.linalg.cholesky)
            const double f00 = g00 * cc[0][0] + g01 * cc[1][0];
            const double f10 = g10 * cc[0][0] + g11 * cc[1][0];
            const double f11 = g10 * cc[0][1] + g11 * cc[1][1];
            const double l00 = sqrt(f00), l10 = f10 / l00, l11 = sqrt(f11 - l10 * l10);
            if (e == 2 * sl) { a0 += l00; a1 += l10; }
            else { a1 += l11; }
        }
        s.A[2 * j][e] = a0;
        s.A[2 * j + 1][e] = a1;
    }
    __syncthreads();
    if (tid < npts) { s.g[ng + 2 * tid] = 2 * pts[tid]; s.g[ng + 2 * tid + 1] = 2 * pts[tid] + 1; }
    __syncthreads();
}

__global__ void __launch_bounds__(PF_THREADS)
psm_factor_kernel(const double* __restrict__ c0, const double* __restrict__ xbar, const double* __restrict__ smean,
                  const double* __restrict__ sscale, const float* __restrict__ mu, const float* __restrict__ cov,
                  float* __restrict__ a_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PFShared& s = *reinterpret_cast<PFShared*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t img = blockIdx.x;
    const float* mu_i = mu + img * P;
    const float* cov_i = cov + img * (CUB200_K * 4);

    if (tid < P) {
        s.scale[tid] = sscale[tid];
        s.d[tid] = xbar[tid] - ((double)mu_i[tid] - smean[tid]) / sscale[tid];  // xbar - transform(mu)
    }
    for (int i = tid; i < P * Z; i += PF_THREADS) s.A[i / Z][i % Z] = 0.0;
    __syncthreads();
    for (int i = tid; i < P * P; i += PF_THREADS) {
        const int a = i / P, b = i % P;
        s.S[a][b] = c0[i] + s.d[a] * s.d[b];  // pca() covariance about transform(mu)   posteriorshapemodel.py:25-27
    }
    // initial points: independent draws mu_j + chol(cov_j) z                       psm.py:239
    if (tid < 3) {
        const int j = c_initial[tid], sl = c_slot[j];
        const double p00 = cov_i[4 * j], p10 = cov_i[4 * j + 2], p11 = cov_i[4 * j + 3];
        const double l00 = sqrt(p00), l10 = p10 / l00, l11 = sqrt(p11 - l10 * l10);
        s.A[2 * j][2 * sl] = l00;
        s.A[2 * j + 1][2 * sl] = l10;
        s.A[2 * j + 1][2 * sl + 1] = l11;
        s.g[2 * tid] = 2 * j;
        s.g[2 * tid + 1] = 2 * j + 1;
    }
    __syncthreads();
    int ng = 6;
    for (int lvl = 0; lvl < 3; ++lvl) {  // sigmas = [1, 1, 1, 1]                     psm.py:225
        psm_level(s, ng, c_level_pts[lvl], c_level_n[lvl], 1.0, true, cov_i);
        ng += 2 * c_level_n[lvl];
    }
    psm_level(s, ng, c_rest, 4, 0.001, false, cov_i);  // posterior mean fill         psm.py:355-368
    float* out = a_out + img * (Z * P);
    for (int i = tid; i < Z * P; i += PF_THREADS) out[i] = (float)s.A[i % P][i / P];  // out[e][d]
}

__global__ void naive_factor_kernel(const float* __restrict__ cov, int64_t n_img, float* __restrict__ a_out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // one (image, landmark)
    if (i >= n_img * CUB200_K) return;
    const int64_t img = i / CUB200_K;
    const int k = (int)(i % CUB200_K);
    const float* c = cov + i * 4;
    const double l00 = sqrt((double)c[0]), l10 = (double)c[2] / l00, l11 = sqrt((double)c[3] - l10 * l10);
    float* out = a_out + img * (P * P);  // out[e][d], zeroed by the caller-side memset below
    out[(2 * k) * P + 2 * k] = (float)l00;
    out[(2 * k) * P + 2 * k + 1] = (float)l10;
    out[(2 * k + 1) * P + 2 * k + 1] = (float)l11;
}

}  // namespace cub200

extern "C" int cub200_psm_factor(const double* model_c0, const double* model_xbar, const double* scaler_mean,
                                 const double* scaler_scale, const float* mu, const float* cov, int64_t n_img,
                                 float* a_out, void* stream) {
    using namespace cub200;
    if (!model_c0 || !model_xbar || !scaler_mean || !scaler_scale || !mu || !cov || !a_out || n_img < 0) return CUB200_EINVAL;
    if (n_img == 0) return 0;
    if (n_img > 0x7fffffff) return CUB200_EINVAL;
    static bool attr_set = false;
    if (!attr_set) {
        CUB200_CUDA_OK(cudaFuncSetAttribute(psm_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PFShared)));
        attr_set = true;
    }
    psm_factor_kernel<<<(unsigned)n_img, PF_THREADS, sizeof(PFShared), (cudaStream_t)stream>>>(
        model_c0, model_xbar, scaler_mean, scaler_scale, mu, cov, a_out);
    return check_launch("psm_factor_kernel");
}

extern "C" int cub200_naive_factor(const float* cov, int64_t n_img, float* a_out, void* stream) {
    using namespace cub200;
    if (!cov || !a_out || n_img < 0) return CUB200_EINVAL;
    if (n_img == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CUB200_CUDA_OK(cudaMemsetAsync(a_out, 0, sizeof(float) * P * P * n_img, st));
    const int64_t n = n_img * CUB200_K;
    naive_factor_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(cov, n_img, a_out);
    return check_launch("naive_factor_kernel");
}

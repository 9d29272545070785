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

constexpr int PF_THREADS = 256;
constexpr int P = CUB200_P;   // 42
constexpr int Z = 34;         // draws of the PSM sampler

// Landmarks in sampling order: initial [0,10,20], levels [5,15], [2,7,13,18], [1,3,6,8,12,14,17,19]
// (get_points_order, psm.py:43-71), then the four posterior-mean-filled points.  The draw slot of
// the landmark at position pi < 17 is pi.
__constant__ int c_order[21] = {0, 10, 20, 5, 15, 2, 7, 13, 18, 1, 3, 6, 8, 12, 14, 17, 19, 4, 9, 11, 16};

// Everything is done in the permuted variable order v = 2*pi + c, so every conditioning set is a
// prefix [0, n) with n = 6, 10, 18, 34, and ONE right-looking Cholesky sweep serves all levels:
// after eliminating the first n variables of T = S + s2 I_[0,n) the trailing block holds the
// posterior covariance Sig_c = S22 - S21 G^-1 S12, and Acc = L21 (L11^-1 At_g) holds the posterior
// mean map M At_g (At = A / scale, the sampled contour in the scaler's space).
struct PFShared {
    double T[P][P + 1];     // working matrix (lower part: L; trailing block: Schur complement)
    double A[P][Z];         // pixel-space map of every variable (row v = permuted variable)
    double Acc[P][Z];       // sum_k L[v][k] y_k
    double y[Z];
    double col[P];
    double scale[P];
    double d[P];
    int perm[P];
};

__device__ __forceinline__ void pf_init_T(PFShared& s, const double* __restrict__ c0, double sigma2, int n_noise) {
    for (int i = threadIdx.x; i < P * P; i += PF_THREADS) {
        const int a = i / P, b = i % P;
        // pca() covariance about transform(mu): C0 + d d^T          posteriorshapemodel.py:25-27
        s.T[a][b] = c0[s.perm[a] * P + s.perm[b]] + s.d[a] * s.d[b] + ((a == b && a < n_noise) ? sigma2 : 0.0);
    }
}

// Eliminate variable k (right-looking): column of L, y_k, rank-1 updates of the trailing block of T
// and of Acc.  `upto`: rows >= this bound are the only ones whose Acc is still needed (all rows > k
// are updated in T).
__device__ __forceinline__ void pf_eliminate(PFShared& s, int k) {
    const int tid = threadIdx.x;
    const double lkk = sqrt(s.T[k][k]);
    const double ilkk = 1.0 / lkk;
    if (tid < P) s.col[tid] = (tid > k) ? s.T[tid][k] * ilkk : 0.0;
    else if (tid >= 64 && tid < 64 + Z) {
        const int e = tid - 64;
        s.y[e] = (s.A[k][e] / s.scale[k] - s.Acc[k][e]) * ilkk;
    }
    __syncthreads();
    // 2-D thread map (no integer division): lane -> column(s), warp -> rows k+1+w, k+1+w+8, ...
    const int tx = tid & 31, ty = tid >> 5;
    for (int a = k + 1 + ty; a < P; a += PF_THREADS / 32) {
        const double ca = s.col[a];
        const int b0 = k + 1 + tx, b1 = b0 + 32;
        if (b0 < P) s.T[a][b0] -= ca * s.col[b0];
        if (b1 < P) s.T[a][b1] -= ca * s.col[b1];
        s.Acc[a][tx] += ca * s.y[tx];
        if (tx + 32 < Z) s.Acc[a][tx + 32] += ca * s.y[tx + 32];
    }
    __syncthreads();
}

// Level: the points at positions [p0, p1) get  mu_f = mu + G (mu_c - mu) + chol(cov_f) z_slot  with
// mu_c - mu = scale * Acc, cov_c = Sig_c * scale (columns), G = cov_p (cov_p + cov_c)^-1,
// cov_f = G cov_c                                                   psm.py:265-285, 424-440
__device__ __forceinline__ void pf_level(PFShared& s, int p0, int p1, const float* __restrict__ cov_img) {
    const int n = (p1 - p0) * Z;
    for (int i = threadIdx.x; i < n; i += PF_THREADS) {
        const int pi = p0 + i / Z, e = i % Z;
        const int v0 = 2 * pi, v1 = v0 + 1;
        const int j = c_order[pi];
        const double c00 = s.T[v0][v0] * s.scale[v0], c01 = s.T[v0][v1] * s.scale[v1];
        const double c10 = s.T[v1][v0] * s.scale[v0], c11 = s.T[v1][v1] * s.scale[v1];
        const double p00 = cov_img[4 * j + 0], p01 = cov_img[4 * j + 1], p10 = cov_img[4 * j + 2], p11 = cov_img[4 * j + 3];
        const double t00 = p00 + c00, t01 = p01 + c01, t10 = p10 + c10, t11 = p11 + c11;
        const double idet = 1.0 / (t00 * t11 - t01 * t10);
        const double i00 = t11 * idet, i01 = -t01 * idet, i10 = -t10 * idet, i11 = t00 * idet;
        const double g00 = p00 * i00 + p01 * i10, g01 = p00 * i01 + p01 * i11;
        const double g10 = p10 * i00 + p11 * i10, g11 = p10 * i01 + p11 * i11;
        const double b0 = s.scale[v0] * s.Acc[v0][e], b1 = s.scale[v1] * s.Acc[v1][e];
        double a0 = g00 * b0 + g01 * b1, a1 = g10 * b0 + g11 * b1;
        if (e == v0 || e == v1) {          // draw slot of position pi is columns (2 pi, 2 pi + 1)
            // cov_f = G cov_c ; lower Cholesky reading the lower triangle (torch.linalg.cholesky)
            const double f00 = g00 * c00 + g01 * c10;
            const double f10 = g10 * c00 + g11 * c10;
            const double f11 = g10 * c01 + g11 * c11;
            const double l00 = sqrt(f00), l10 = f10 / l00, l11 = sqrt(f11 - l10 * l10);
            if (e == v0) { a0 += l00; a1 += l10; }
            else { a1 += l11; }
        }
        s.A[v0][e] = a0;
        s.A[v1][e] = a1;
    }
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
        const int o = 2 * c_order[tid >> 1] + (tid & 1);      // original flat index of permuted variable
        s.perm[tid] = o;
        s.scale[tid] = sscale[o];
        s.d[tid] = xbar[o] - ((double)mu_i[o] - smean[o]) / sscale[o];  // xbar - transform(mu)
    }
    for (int i = tid; i < P * Z; i += PF_THREADS) { s.A[i / Z][i % Z] = 0.0; s.Acc[i / Z][i % Z] = 0.0; }
    __syncthreads();
    pf_init_T(s, c0, 1.0, 6);                                 // sigmas = [1, 1, 1, 1]   psm.py:225
    // initial points: independent draws mu_j + chol(cov_j) z                       psm.py:239
    if (tid < 3) {
        const int j = c_order[tid];
        const double p00 = cov_i[4 * j], p10 = cov_i[4 * j + 2], p11 = cov_i[4 * j + 3];
        const double l00 = sqrt(p00), l10 = p10 / l00, l11 = sqrt(p11 - l10 * l10);
        s.A[2 * tid][2 * tid] = l00;
        s.A[2 * tid + 1][2 * tid] = l10;
        s.A[2 * tid + 1][2 * tid + 1] = l11;
    }
    __syncthreads();
    // sweep 1: sigma2 = 1; levels at n = 6, 10, 18
    for (int k = 0; k < 18; ++k) {
        if (k == 6 || k == 10) {
            const int p0 = k >> 1, p1 = (k == 6) ? 5 : 9;
            pf_level(s, p0, p1, cov_i);
            if (tid >= 2 * p0 && tid < 2 * p1) s.T[tid][tid] += 1.0;   // the new points join the conditioning set
            __syncthreads();
        }
        pf_eliminate(s, k);
    }
    pf_level(s, 9, 17, cov_i);
    // sweep 2: posterior mean of the remaining points given all 17 sampled ones, sigma2 = 0.001   psm.py:355-368
    pf_init_T(s, c0, 0.001, Z);
    for (int i = tid; i < P * Z; i += PF_THREADS) s.Acc[i / Z][i % Z] = 0.0;
    __syncthreads();
    for (int k = 0; k < Z; ++k) pf_eliminate(s, k);
    for (int i = tid; i < 8 * Z; i += PF_THREADS) {
        const int v = Z + i / Z, e = i % Z;
        s.A[v][e] = s.scale[v] * s.Acc[v][e];
    }
    __syncthreads();
    float* out = a_out + img * (Z * P);
    for (int i = tid; i < Z * P; i += PF_THREADS) {
        const int e = i / P, v = i % P;
        out[e * P + s.perm[v]] = (float)s.A[v][e];            // out[e][d], d = original flat index
    }
}

// ---------------------------------------------------------------- conditional maps (skew path)
// The skew-normal sampler (psm_skew.py:247-411) draws every level point from a GRID pdf, so its contour
// is not affine in Gaussian draws; what stays constant per image is the shape model's conditional law
// of each level given the already-sampled points (compute_psm, psm_skew.py:469-491):
//     mu_c[v] = mu[v] + sum_{e < n} Mc[v][e] (s[e] - mu[e]),      cov_c = 2x2 blocks of Sig_c * scale
// The same right-looking sweep as above with A = identity (the sampled values themselves are the
// free variables) yields Mc; rows/columns are in the permuted (sampling) order.
//   mc_t (n_img, 34, 36) fp64: [e][v - 6], v = 6..41 (level points at positions 3..16, then the four
//        posterior-mean-filled points with sigma2 = 1e-3); zero for e >= the level's prefix
//   covc (n_img, 14, 3) fp64: (c00, c10, c11) of positions 3..16 (MultivariateNormal reads the lower
//        triangle, psm_skew.py:78)
__device__ __forceinline__ void pc_emit(PFShared& s, int p0, int p1, int n, double* __restrict__ mc, double* __restrict__ cc) {
    const int rows = 2 * (p1 - p0);
    for (int i = threadIdx.x; i < rows * Z; i += PF_THREADS) {
        const int v = 2 * p0 + i / Z, e = i % Z;
        mc[e * 36 + (v - 6)] = (e < n) ? s.scale[v] * s.Acc[v][e] : 0.0;
    }
    if (cc) {
        for (int pi = p0 + threadIdx.x; pi < p1; pi += PF_THREADS) {
            const int v0 = 2 * pi, v1 = v0 + 1;
            cc[(pi - 3) * 3 + 0] = s.T[v0][v0] * s.scale[v0];
            cc[(pi - 3) * 3 + 1] = s.T[v1][v0] * s.scale[v0];
            cc[(pi - 3) * 3 + 2] = s.T[v1][v1] * s.scale[v1];
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PF_THREADS)
psm_cond_kernel(const double* __restrict__ c0, const double* __restrict__ xbar, const double* __restrict__ smean,
                const double* __restrict__ sscale, const float* __restrict__ mu, double* __restrict__ mc_t,
                double* __restrict__ covc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PFShared& s = *reinterpret_cast<PFShared*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t img = blockIdx.x;
    const float* mu_i = mu + img * P;
    double* mc = mc_t + img * (Z * 36);
    double* cc = covc + img * (14 * 3);
    if (tid < P) {
        const int o = 2 * c_order[tid >> 1] + (tid & 1);
        s.perm[tid] = o;
        s.scale[tid] = sscale[o];
        s.d[tid] = xbar[o] - ((double)mu_i[o] - smean[o]) / sscale[o];
    }
    for (int i = tid; i < P * Z; i += PF_THREADS) { s.A[i / Z][i % Z] = (i / Z == i % Z) ? 1.0 : 0.0; s.Acc[i / Z][i % Z] = 0.0; }
    __syncthreads();
    pf_init_T(s, c0, 1.0, 6);                                 // sigmas = [1, 1, 1, 1]   psm_skew.py:270
    __syncthreads();
    for (int k = 0; k < 18; ++k) {
        if (k == 6 || k == 10) {
            const int p0 = k >> 1, p1 = (k == 6) ? 5 : 9;
            pc_emit(s, p0, p1, k, mc, cc);
            if (tid >= 2 * p0 && tid < 2 * p1) s.T[tid][tid] += 1.0;
            __syncthreads();
        }
        pf_eliminate(s, k);
    }
    pc_emit(s, 9, 17, 18, mc, cc);
    pf_init_T(s, c0, 0.001, Z);                               // psm_skew.py:388
    for (int i = tid; i < P * Z; i += PF_THREADS) s.Acc[i / Z][i % Z] = 0.0;
    __syncthreads();
    for (int k = 0; k < Z; ++k) pf_eliminate(s, k);
    pc_emit(s, 17, 21, Z, mc, nullptr);
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

extern "C" int cub200_psm_conditionals(const double* model_c0, const double* model_xbar, const double* scaler_mean,
                                       const double* scaler_scale, const float* mu, int64_t n_img, double* mc_t,
                                       double* covc, void* stream) {
    using namespace cub200;
    if (!model_c0 || !model_xbar || !scaler_mean || !scaler_scale || !mu || !mc_t || !covc || n_img < 0) return CUB200_EINVAL;
    if (n_img == 0) return 0;
    if (n_img > 0x7fffffff) return CUB200_EINVAL;
    static bool attr_set = false;
    if (!attr_set) {
        CUB200_CUDA_OK(cudaFuncSetAttribute(psm_cond_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PFShared)));
        attr_set = true;
    }
    psm_cond_kernel<<<(unsigned)n_img, PF_THREADS, sizeof(PFShared), (cudaStream_t)stream>>>(
        model_c0, model_xbar, scaler_mean, scaler_scale, mu, mc_t, covc);
    return check_launch("psm_cond_kernel");
}

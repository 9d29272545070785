// K2  psm_factor / psm_conditionals: the hierarchical posterior-shape-model sampler reduced to per-image
// (or per-sample) constants, ONE WARP per image, no eigendecomposition, no factorisation at run time.
//
// The reference re-runs a 42x42 eigendecomposition per image and two 42x42 fp32 inverses per level per
// SAMPLE (contour_uncertainty/sampler/posterior_shape_model/psm.py:91,265-352; posteriorshapemodel.py:9-81).
// Facts used here:
//   * pca() about the centre transform(mu) only enters through S = Q Q^T = C0 + d d^T, C0 = covariance of
//     X_train about its own mean (a constant of the model), d = xbar - transform(mu);
//   * the conditioning sets of the three levels and of the final fill are PREFIXES of the sampling order
//     (points 0,10,20 | 5,15 | 2,7,13,18 | 1,3,6,8,12,14,17,19 | 4,9,11,16), with noise sigma2 = 1,1,1,1e-3;
//   * so with B = C0[g,g] + sigma2 I (constant) Sherman-Morrison gives, for u = B^-1 d_g, gamma = 1/(1 + d_g^T u):
//         M     = S[r,g] (S[g,g] + sigma2 I)^-1 = K + gamma (d_r - C0[r,g] u) u^T,          K = C0[r,g] B^-1
//         Sig_c = S[r,r] - M S[g,r]             = Sig0 + d_r d_r^T - (K d_g) d_r^T - c w^T - (c s) d_r^T
//     (c = gamma (d_r - w), w = C0[r,g] u, s = d_g^T u, Sig0 = C0[r,r] - K C0[g,r]); only 2x2 diagonal blocks are used.
//   B^-1, K, C0[r,g] and Sig0 of the four levels are packed once per shape model (cub200_psm_model_pack_host);
//   per image the work is a few mat-vecs and the (rows x n)(n x 34) product K At_g.
//   * the Gaussian sampler is affine in its draws, contour = mu + A z (A: 42 x 34): per level
//         mu_c - mu = scale M At_g z,  merge_priors: G = cov_p (cov_p + cov_c)^-1, cov_f = G cov_c,
//         draw: mu_f + chol(cov_f) z_slot                                            psm.py:265-285, 424-440
//     the skew sampler needs M itself (its level points are grid draws): psm_conditionals.
#include <cmath>
#include <utility>
#include <vector>
#include "common.cuh"

namespace cub200 {

constexpr int P = CUB200_P;   // 42
constexpr int Z = 34;         // draws of the PSM sampler = sampled coordinates
#ifndef CUB200_PF_WARPS
#define CUB200_PF_WARPS 4
#endif
constexpr int PF_WARPS = CUB200_PF_WARPS;

// Landmarks in sampling order (get_points_order, psm.py:43-71), then the four posterior-mean-filled points.
// Everything below works on the permuted variables v = 2 * position + coordinate.
__constant__ int c_order[21] = {0, 10, 20, 5, 15, 2, 7, 13, 18, 1, 3, 6, 8, 12, 14, 17, 19, 4, 9, 11, 16};
static const int h_order[21] = {0, 10, 20, 5, 15, 2, 7, 13, 18, 1, 3, 6, 8, 12, 14, 17, 19, 4, 9, 11, 16};

// level l: positions [p0, p1), observed prefix n = 2 p0 variables, rows = 2 (p1 - p0)
struct LevelDef { int p0, p1, n, rows; double sigma2; int off_binv, off_k, off_crg, off_sig0; };
// pack layout (doubles): [xbar_p 42][mean_p 42][scale_p 42] then per level Binv (n x n), K (rows x n), Crg (rows x n),
// Sig0 (rows/2 x 4: the 2x2 diagonal blocks)
constexpr int PK_XBAR = 0, PK_MEAN = 42, PK_SCALE = 84, PK_LEVELS = 126;
constexpr int lvl_size(int n, int rows) { return n * n + 2 * rows * n + 2 * rows; }
constexpr int L0 = PK_LEVELS, L1 = L0 + lvl_size(6, 4), L2 = L1 + lvl_size(10, 8), L3 = L2 + lvl_size(18, 16);
constexpr int PACK_DOUBLES = L3 + lvl_size(34, 8);
static_assert(PACK_DOUBLES == CUB200_PSM_PACK_DOUBLES, "include/cub200.h: CUB200_PSM_PACK_DOUBLES");
__host__ __device__ constexpr LevelDef level_def(int l) {
    return l == 0 ? LevelDef{3, 5, 6, 4, 1.0, L0, L0 + 36, L0 + 36 + 24, L0 + 36 + 48}
         : l == 1 ? LevelDef{5, 9, 10, 8, 1.0, L1, L1 + 100, L1 + 100 + 80, L1 + 100 + 160}
         : l == 2 ? LevelDef{9, 17, 18, 16, 1.0, L2, L2 + 324, L2 + 324 + 288, L2 + 324 + 576}
                  : LevelDef{17, 21, 34, 8, 1e-3, L3, L3 + 1156, L3 + 1156 + 272, L3 + 1156 + 544};
}

struct __align__(16) PFWarp {
    double A[Z][Z];        // map of the sampled variables IN THE SCALER'S SPACE, At = A / scale (row = permuted variable, col = draw)
    double Acc[16][Z];     // M At_g of the current level
    double d[P], u[Z], t[Z];
    double w[16], kd[16], corr[16];
    double sc[8][3];       // conditional covariance blocks of the current level: (c00, c10, c11)
};

// An image is processed by one warp (large batches: per-sample factors of the sequence sampler) or by the whole
// 4-warp CTA (small batches: a 64-image step would otherwise run on 64 warps of the GPU).
template <int WPI> __device__ __forceinline__ void group_sync() { if (WPI == 1) __syncwarp(); else __syncthreads(); }

// The per-level scalars and vectors of one image: ws.u, ws.w, ws.kd, ws.corr, ws.sc.  `lane` = thread index in the
// image's group of NT = 32 WPI threads.
template <int WPI, class Scratch>
__device__ __forceinline__ void level_common(Scratch& ws, const double* __restrict__ pk, const LevelDef L, int lane) {
    constexpr int NT = 32 * WPI;
    const double* Binv = pk + L.off_binv;
    const double* K = pk + L.off_k;
    const double* Crg = pk + L.off_crg;
    const int n = L.n, rows = L.rows;
    // u = B^-1 d_g
    for (int i = lane; i < n; i += NT) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc = fma(Binv[i * n + k], ws.d[k], acc);
        ws.u[i] = acc;
    }
    group_sync<WPI>();
    double s = 0.0;                                      // every warp forms the same dot product
    for (int k = lane & 31; k < n; k += 32) s = fma(ws.d[k], ws.u[k], s);
    s = warp_sum(s);
    const double gamma = 1.0 / (1.0 + s);
    if (lane < rows) {
        double w = 0.0, kd = 0.0;
        for (int k = 0; k < n; ++k) {
            w = fma(Crg[lane * n + k], ws.u[k], w);
            kd = fma(K[lane * n + k], ws.d[k], kd);
        }
        ws.w[lane] = w;
        ws.kd[lane] = kd;
        ws.corr[lane] = gamma * (ws.d[2 * L.p0 + lane] - w);
    }
    group_sync<WPI>();
    // conditional covariance blocks (levels with a draw): Sig_c * scale (columns)
    if (L.sigma2 == 1.0 && lane < rows / 2) {
        const double* sig0 = pk + L.off_sig0 + lane * 4;
        const int a0 = 2 * lane, a1 = a0 + 1;
        const double d0 = ws.d[2 * L.p0 + a0], d1 = ws.d[2 * L.p0 + a1];
        const double s0 = pk[PK_SCALE + 2 * L.p0 + a0], s1 = pk[PK_SCALE + 2 * L.p0 + a1];
        auto blk = [&](int a, int b, double da, double db, double s0ab) {
            return s0ab + da * db - ws.kd[a] * db - ws.corr[a] * ws.w[b] - ws.corr[a] * s * db;
        };
        ws.sc[lane][0] = blk(a0, a0, d0, d0, sig0[0]) * s0;     // c00
        ws.sc[lane][1] = blk(a1, a0, d1, d0, sig0[2]) * s0;     // c10 (the lower triangle is what Cholesky reads)
        ws.sc[lane][2] = blk(a1, a1, d1, d1, sig0[3]) * s1;     // c11
    }
    group_sync<WPI>();
}

// ---------------------------------------------------------------- Gaussian sampler: affine factor
template <int WPI>
__global__ void __launch_bounds__(PF_WARPS * 32)
psm_factor_kernel(const double* __restrict__ pack, const float* __restrict__ mu, const float* __restrict__ cov, int64_t n_img,
                  float* __restrict__ a_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* pk = reinterpret_cast<double*>(smem_raw);
    PFWarp* warps = reinterpret_cast<PFWarp*>(pk + PACK_DOUBLES);
    for (int i = threadIdx.x; i < PACK_DOUBLES; i += PF_WARPS * 32) pk[i] = pack[i];
    __syncthreads();
    constexpr int NT = 32 * WPI, IPC = PF_WARPS / WPI;       // threads per image, images per CTA pass
    const int lane = threadIdx.x % NT, wid = threadIdx.x / NT;
    PFWarp& ws = warps[wid];
    // every group of a CTA runs the same number of passes (the group syncs of WPI > 1 are CTA barriers)
    for (int64_t img0 = (int64_t)blockIdx.x * IPC; img0 < n_img; img0 += (int64_t)gridDim.x * IPC) {
        const int64_t img = img0 + wid;
        if (WPI == 1 && img >= n_img) break;
        const float* mu_i = mu + img * P;
        const float* cov_i = cov + img * (CUB200_K * 4);
        for (int v = lane; v < P; v += NT) {
            const int o = 2 * c_order[v >> 1] + (v & 1);
            ws.d[v] = pk[PK_XBAR + v] - ((double)mu_i[o] - pk[PK_MEAN + v]) / pk[PK_SCALE + v];   // xbar - transform(mu)
        }
        for (int i = lane; i < Z * Z; i += NT) (&ws.A[0][0])[i] = 0.0;
        group_sync<WPI>();
        // initial points: independent draws mu_j + chol(cov_j) z                       psm.py:239
        if (lane < 3) {
            const int j = c_order[lane];
            const double p00 = cov_i[4 * j], p10 = cov_i[4 * j + 2], p11 = cov_i[4 * j + 3];
            const double l00 = sqrt(p00), l10 = p10 / l00, l11 = sqrt(p11 - l10 * l10);
            ws.A[2 * lane][2 * lane] = l00 / pk[PK_SCALE + 2 * lane];
            ws.A[2 * lane + 1][2 * lane] = l10 / pk[PK_SCALE + 2 * lane + 1];
            ws.A[2 * lane + 1][2 * lane + 1] = l11 / pk[PK_SCALE + 2 * lane + 1];
        }
        group_sync<WPI>();
        float* out = a_out + img * (Z * P);
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
            const LevelDef L = level_def(l);
            level_common<WPI>(ws, pk, L, lane);
            const int n = L.n, rows = L.rows;
            const double* K = pk + L.off_k;
            // t = u^T At_g ; Acc = K At_g + corr t^T      (At = A / scale: the sampled contour in the scaler's space)
            for (int e = lane; e < Z; e += NT) {
                double acc = 0.0;
                for (int k = 0; k < n; ++k) acc = fma(ws.u[k], ws.A[k][e], acc);
                ws.t[e] = acc;
            }
            group_sync<WPI>();
            for (int idx = lane; idx < rows * Z; idx += NT) {
                const int a = idx / Z, e = idx - a * Z;
                double acc = ws.corr[a] * ws.t[e];
                for (int k = 0; k < n; ++k) acc = fma(K[a * n + k], ws.A[k][e], acc);
                ws.Acc[a][e] = acc;
            }
            group_sync<WPI>();
            if (l < 3) {
                // merge with the predicted law and colour: rows of A for the level's points
                for (int idx = lane; idx < (rows / 2) * Z; idx += NT) {
                    const int q = idx / Z, e = idx - q * Z;
                    const int pi = L.p0 + q, v0 = 2 * pi, v1 = v0 + 1, j = c_order[pi];
                    const double c00 = ws.sc[q][0], c10 = ws.sc[q][1], c11 = ws.sc[q][2];
                    // c01 = Sig_c[v0][v1] * scale[v1] and Sig_c is symmetric
                    const double c01 = c10 / pk[PK_SCALE + v0] * pk[PK_SCALE + v1];
                    const double p00 = cov_i[4 * j + 0], p01 = cov_i[4 * j + 1], p10 = cov_i[4 * j + 2], p11 = cov_i[4 * j + 3];
                    const double t00 = p00 + c00, t01 = p01 + c01, t10 = p10 + c10, t11 = p11 + c11;
                    const double idet = 1.0 / (t00 * t11 - t01 * t10);
                    const double i00 = t11 * idet, i01 = -t01 * idet, i10 = -t10 * idet, i11 = t00 * idet;
                    const double g00 = p00 * i00 + p01 * i10, g01 = p00 * i01 + p01 * i11;
                    const double g10 = p10 * i00 + p11 * i10, g11 = p10 * i01 + p11 * i11;
                    const double b0 = pk[PK_SCALE + v0] * ws.Acc[2 * q][e], b1 = pk[PK_SCALE + v1] * ws.Acc[2 * q + 1][e];
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
                    ws.A[v0][e] = a0 / pk[PK_SCALE + v0];
                    ws.A[v1][e] = a1 / pk[PK_SCALE + v1];
                }
                group_sync<WPI>();
            } else {
                // the remaining points: posterior mean given all 17 sampled ones, sigma2 = 1e-3      psm.py:355-368
                for (int idx = lane; idx < rows * Z; idx += NT) {
                    const int a = idx / Z, e = idx - a * Z, v = Z + a;
                    const int o = 2 * c_order[v >> 1] + (v & 1);
                    out[e * P + o] = (float)(pk[PK_SCALE + v] * ws.Acc[a][e]);
                }
            }
        }
        for (int idx = lane; idx < Z * Z; idx += NT) {
            const int v = idx / Z, e = idx - v * Z;
            const int o = 2 * c_order[v >> 1] + (v & 1);
            out[e * P + o] = (float)(ws.A[v][e] * pk[PK_SCALE + v]);   // out[e][d], d = original flat index
        }
        group_sync<WPI>();
    }
}


// ---------------------------------------------------------------- Gaussian sampler: ONE sample per item, no factor
// The sequence sampler draws exactly one contour from each of its per-(pair, sample) merged laws (sequence_sampler.py:94:
// the second instant's PosteriorShapeModelSampler call, whose pca is re-centred per sample, psm.py:91).  Building the
// 34 x 42 affine factor for that (psm_factor_kernel<1>) and applying it to one draw vector is 34 x the arithmetic and 5.7 KB of
// traffic per sample for nothing: every formula of the factor is linear in its columns, so the same recursion runs on the
// single column x = A z -- the deviation of the sampled variables from their predicted means in the scaler's space.
// One warp per item; draws z (n_items, 34) replayed, or Philox keyed by (seed, item + sample_offset) exactly like
// sample_contours_warp_kernel (same draws as the factor path; the result differs from it only by that path's fp32
// rounding of the factor).  contours_out (n_items, 21, 2) fp32.
struct __align__(16) PDWarp {
    double x[Z], acc[16];
    double d[P], u[Z];
    double w[16], kd[16], corr[16];
    double sc[8][3];
    float z[Z + 2];
};
constexpr int PD_WARPS = 8;

__global__ void __launch_bounds__(PD_WARPS * 32)
psm_sample_direct_kernel(const double* __restrict__ pack, const float* __restrict__ mu, const float* __restrict__ cov,
                         const float* __restrict__ z, uint64_t seed, uint64_t g_offset, int64_t n_items,
                         float* __restrict__ contours_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* pk = reinterpret_cast<double*>(smem_raw);
    PDWarp* warps = reinterpret_cast<PDWarp*>(pk + PACK_DOUBLES);
    for (int i = threadIdx.x; i < PACK_DOUBLES; i += PD_WARPS * 32) pk[i] = pack[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    PDWarp& ws = warps[wid];
    for (int64_t it = (int64_t)blockIdx.x * PD_WARPS + wid; it < n_items; it += (int64_t)gridDim.x * PD_WARPS) {
        const float* mu_i = mu + it * P;
        const float* cov_i = cov + it * (CUB200_K * 4);
        for (int v = lane; v < P; v += 32) {
            const int o = 2 * c_order[v >> 1] + (v & 1);
            ws.d[v] = pk[PK_XBAR + v] - ((double)mu_i[o] - pk[PK_MEAN + v]) / pk[PK_SCALE + v];   // xbar - transform(mu)
        }
        if (z) {
            for (int e = lane; e < Z; e += 32) ws.z[e] = z[it * Z + e];
        } else if (lane * 4 < Z) {
            philox_normals4((uint64_t)it + g_offset, lane, seed, ws.z[4 * lane], ws.z[4 * lane + 1], ws.z[4 * lane + 2], ws.z[4 * lane + 3]);
        }
        __syncwarp();
        // initial points: independent draws mu_j + chol(cov_j) z                       psm.py:239
        if (lane < 3) {
            const int j = c_order[lane];
            const double p00 = cov_i[4 * j], p10 = cov_i[4 * j + 2], p11 = cov_i[4 * j + 3];
            const double l00 = sqrt(p00), l10 = p10 / l00, l11 = sqrt(p11 - l10 * l10);
            const double z0 = ws.z[2 * lane], z1 = ws.z[2 * lane + 1];
            ws.x[2 * lane] = l00 * z0 / pk[PK_SCALE + 2 * lane];
            ws.x[2 * lane + 1] = (l10 * z0 + l11 * z1) / pk[PK_SCALE + 2 * lane + 1];
        }
        __syncwarp();
        float* out = contours_out + it * P;
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
            const LevelDef L = level_def(l);
            level_common<1>(ws, pk, L, lane);
            const int n = L.n, rows = L.rows;
            const double* K = pk + L.off_k;
            // t = u^T x_g ; acc = K x_g + corr t
            double t = 0.0;
            for (int k = lane; k < n; k += 32) t = fma(ws.u[k], ws.x[k], t);
            t = warp_sum(t);
            // rows x n mat-vec: two lanes per row, each half of the columns (rows <= 16)
            {
                const int a = lane >> 1, h = lane & 1;
                double acc = 0.0;
                if (a < rows) {
                    const int k0 = h ? (n >> 1) : 0, k1 = h ? n : (n >> 1);
                    for (int k = k0; k < k1; ++k) acc = fma(K[a * n + k], ws.x[k], acc);
                }
                acc += __shfl_xor_sync(FULL, acc, 1);
                if (a < rows && h == 0) ws.acc[a] = fma(ws.corr[a], t, acc);
            }
            __syncwarp();
            if (l < 3) {
                // merge with the predicted law and colour with the position's own two draws
                if (lane < rows / 2) {
                    const int q = lane;
                    const int pi = L.p0 + q, v0 = 2 * pi, v1 = v0 + 1, j = c_order[pi];
                    const double c00 = ws.sc[q][0], c10 = ws.sc[q][1], c11 = ws.sc[q][2];
                    const double c01 = c10 / pk[PK_SCALE + v0] * pk[PK_SCALE + v1];
                    const double p00 = cov_i[4 * j + 0], p01 = cov_i[4 * j + 1], p10 = cov_i[4 * j + 2], p11 = cov_i[4 * j + 3];
                    const double t00 = p00 + c00, t01 = p01 + c01, t10 = p10 + c10, t11 = p11 + c11;
                    const double idet = 1.0 / (t00 * t11 - t01 * t10);
                    const double i00 = t11 * idet, i01 = -t01 * idet, i10 = -t10 * idet, i11 = t00 * idet;
                    const double g00 = p00 * i00 + p01 * i10, g01 = p00 * i01 + p01 * i11;
                    const double g10 = p10 * i00 + p11 * i10, g11 = p10 * i01 + p11 * i11;
                    const double b0 = pk[PK_SCALE + v0] * ws.acc[2 * q], b1 = pk[PK_SCALE + v1] * ws.acc[2 * q + 1];
                    // cov_f = G cov_c ; lower Cholesky reading the lower triangle (torch.linalg.cholesky)
                    const double f00 = g00 * c00 + g01 * c10;
                    const double f10 = g10 * c00 + g11 * c10;
                    const double f11 = g10 * c01 + g11 * c11;
                    const double l00 = sqrt(f00), l10 = f10 / l00, l11 = sqrt(f11 - l10 * l10);
                    const double z0 = ws.z[v0], z1 = ws.z[v1];
                    const double a0 = g00 * b0 + g01 * b1 + l00 * z0;
                    const double a1 = g10 * b0 + g11 * b1 + l10 * z0 + l11 * z1;
                    ws.x[v0] = a0 / pk[PK_SCALE + v0];
                    ws.x[v1] = a1 / pk[PK_SCALE + v1];
                }
                __syncwarp();
            } else if (lane < rows) {
                // the remaining points: posterior mean given all 17 sampled ones, sigma2 = 1e-3      psm.py:355-368
                const int v = Z + lane;
                const int o = 2 * c_order[v >> 1] + (v & 1);
                out[o] = (float)((double)mu_i[o] + pk[PK_SCALE + v] * ws.acc[lane]);
            }
        }
        for (int v = lane; v < Z; v += 32) {
            const int o = 2 * c_order[v >> 1] + (v & 1);
            out[o] = (float)((double)mu_i[o] + ws.x[v] * pk[PK_SCALE + v]);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- conditional maps (skew path)
//   mc_t (n_img, 34, 36) fp64: [e][v - 6], v = 6..41; zero for e >= the level's prefix
//   covc (n_img, 14, 3) fp64: (c00, c10, c11) of positions 3..16 (MultivariateNormal reads the lower triangle)
template <int WPI>
__global__ void __launch_bounds__(PF_WARPS * 32)
psm_cond_kernel(const double* __restrict__ pack, const float* __restrict__ mu, int64_t n_img, double* __restrict__ mc_t,
                double* __restrict__ covc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* pk = reinterpret_cast<double*>(smem_raw);
    PFWarp* warps = reinterpret_cast<PFWarp*>(pk + PACK_DOUBLES);
    for (int i = threadIdx.x; i < PACK_DOUBLES; i += PF_WARPS * 32) pk[i] = pack[i];
    __syncthreads();
    constexpr int NT = 32 * WPI, IPC = PF_WARPS / WPI;       // threads per image, images per CTA pass
    const int lane = threadIdx.x % NT, wid = threadIdx.x / NT;
    PFWarp& ws = warps[wid];
    // every group of a CTA runs the same number of passes (the group syncs of WPI > 1 are CTA barriers)
    for (int64_t img0 = (int64_t)blockIdx.x * IPC; img0 < n_img; img0 += (int64_t)gridDim.x * IPC) {
        const int64_t img = img0 + wid;
        if (WPI == 1 && img >= n_img) break;
        const float* mu_i = mu + img * P;
        double* mc = mc_t + img * (Z * 36);
        double* cc = covc + img * (14 * 3);
        for (int v = lane; v < P; v += NT) {
            const int o = 2 * c_order[v >> 1] + (v & 1);
            ws.d[v] = pk[PK_XBAR + v] - ((double)mu_i[o] - pk[PK_MEAN + v]) / pk[PK_SCALE + v];
        }
        group_sync<WPI>();
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
            const LevelDef L = level_def(l);
            level_common<WPI>(ws, pk, L, lane);
            const int n = L.n, rows = L.rows;
            const double* K = pk + L.off_k;
            // mu_c[v] - mu[v] = scale_v sum_e M[v][e] (s_e - mu_e) / scale_e      (compute_psm, psm_skew.py:469-491)
            for (int idx = lane; idx < rows * Z; idx += NT) {
                const int a = idx / Z, e = idx - a * Z, v = 2 * L.p0 + a;
                mc[e * 36 + (v - 6)] = (e < n) ? pk[PK_SCALE + v] * (K[a * n + e] + ws.corr[a] * ws.u[e]) / pk[PK_SCALE + e] : 0.0;
            }
            if (l < 3 && lane < rows / 2) {
                const int pi = L.p0 + lane;
                cc[(pi - 3) * 3 + 0] = ws.sc[lane][0];
                cc[(pi - 3) * 3 + 1] = ws.sc[lane][1];
                cc[(pi - 3) * 3 + 2] = ws.sc[lane][2];
            }
            group_sync<WPI>();
        }
    }
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

static int pf_launch_cfg(int64_t n_img, const void* kernel, int images_per_cta, int* grid, size_t* smem) {
    const bool warp_per_image = images_per_cta > 1;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        CUB200_CUDA_OK(cudaGetDevice(&dev));
        CUB200_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // (reading the constants through L1 instead of staging them, to fit a third CTA per SM, was slower: 9.3 vs 7.9 ms)
    *smem = sizeof(double) * PACK_DOUBLES + sizeof(PFWarp) * (warp_per_image ? PF_WARPS : 1);
    CUB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
    const int64_t need = (n_img + images_per_cta - 1) / images_per_cta;
    const int64_t cap = (int64_t)num_sms * (warp_per_image ? 2 : 4);
    *grid = (int)(need < cap ? need : cap);
    return 0;
}
// one CTA per image while that still fills the GPU, one warp per image beyond
static bool pf_cta_per_image(int64_t n_img) { return n_img <= 4096; }

// Gauss-Jordan inverse of an n x n symmetric positive-definite matrix (host, fp64, partial pivoting).
static bool invert_spd(std::vector<double>& a, int n) {
    std::vector<double> inv((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (std::fabs(a[(size_t)r * n + c]) > std::fabs(a[(size_t)piv * n + c])) piv = r;
        if (a[(size_t)piv * n + c] == 0.0) return false;
        if (piv != c)
            for (int k = 0; k < n; ++k) {
                std::swap(a[(size_t)piv * n + k], a[(size_t)c * n + k]);
                std::swap(inv[(size_t)piv * n + k], inv[(size_t)c * n + k]);
            }
        const double ip = 1.0 / a[(size_t)c * n + c];
        for (int k = 0; k < n; ++k) { a[(size_t)c * n + k] *= ip; inv[(size_t)c * n + k] *= ip; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[(size_t)r * n + c];
            if (f == 0.0) continue;
            for (int k = 0; k < n; ++k) { a[(size_t)r * n + k] -= f * a[(size_t)c * n + k]; inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k]; }
        }
    }
    a.swap(inv);
    return true;
}

}  // namespace cub200

extern "C" int cub200_psm_model_pack_host(const double* c0_host, const double* xbar_host, const double* mean_host,
                                          const double* scale_host, double* pack_out_host) {
    using namespace cub200;
    if (!c0_host || !xbar_host || !mean_host || !scale_host || !pack_out_host) return CUB200_EINVAL;
    int perm[P];
    for (int v = 0; v < P; ++v) perm[v] = 2 * h_order[v >> 1] + (v & 1);
    for (int v = 0; v < P; ++v) {
        pack_out_host[PK_XBAR + v] = xbar_host[perm[v]];
        pack_out_host[PK_MEAN + v] = mean_host[perm[v]];
        pack_out_host[PK_SCALE + v] = scale_host[perm[v]];
    }
    auto C = [&](int a, int b) { return c0_host[perm[a] * P + perm[b]]; };
    for (int l = 0; l < 4; ++l) {
        const LevelDef L = level_def(l);
        const int n = L.n, rows = L.rows, r0 = 2 * L.p0;
        std::vector<double> B((size_t)n * n);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < n; ++k) B[(size_t)i * n + k] = C(i, k) + (i == k ? L.sigma2 : 0.0);
        if (!invert_spd(B, n)) return CUB200_EINVAL;
        double* binv = pack_out_host + L.off_binv;
        double* K = pack_out_host + L.off_k;
        double* crg = pack_out_host + L.off_crg;
        double* sig0 = pack_out_host + L.off_sig0;
        for (int i = 0; i < n * n; ++i) binv[i] = B[i];
        for (int a = 0; a < rows; ++a)
            for (int k = 0; k < n; ++k) {
                crg[a * n + k] = C(r0 + a, k);
                double acc = 0.0;
                for (int m = 0; m < n; ++m) acc += C(r0 + a, m) * B[(size_t)m * n + k];
                K[a * n + k] = acc;
            }
        for (int q = 0; q < rows / 2; ++q)
            for (int i = 0; i < 2; ++i)
                for (int k = 0; k < 2; ++k) {
                    const int a = 2 * q + i, b = 2 * q + k;
                    double acc = C(r0 + a, r0 + b);
                    for (int m = 0; m < n; ++m) acc -= K[a * n + m] * C(m, r0 + b);
                    sig0[q * 4 + i * 2 + k] = acc;
                }
    }
    return 0;
}

extern "C" int cub200_psm_factor(const double* model_pack, const float* mu, const float* cov, int64_t n_img, float* a_out,
                                 void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!model_pack || !mu || !cov || !a_out || n_img < 0) return CUB200_EINVAL;
    int grid = 0;
    size_t smem = 0;
    const bool wide = pf_cta_per_image(n_img);
    int rc = wide ? pf_launch_cfg(n_img, (const void*)psm_factor_kernel<PF_WARPS>, 1, &grid, &smem)
                  : pf_launch_cfg(n_img, (const void*)psm_factor_kernel<1>, PF_WARPS, &grid, &smem);
    if (rc) return rc;
    if (wide) psm_factor_kernel<PF_WARPS><<<grid, PF_WARPS * 32, smem, (cudaStream_t)stream>>>(model_pack, mu, cov, n_img, a_out);
    else psm_factor_kernel<1><<<grid, PF_WARPS * 32, smem, (cudaStream_t)stream>>>(model_pack, mu, cov, n_img, a_out);
    return check_launch("psm_factor_kernel");
}

extern "C" int cub200_psm_sample_direct(const double* model_pack, const float* mu, const float* cov, const float* z, uint64_t seed,
                                        uint64_t sample_offset, int64_t n_items, float* contours_out, void* stream) {
    using namespace cub200;
    if (n_items == 0) return 0;
    if (!model_pack || !mu || !cov || !contours_out || n_items < 0) return CUB200_EINVAL;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        CUB200_CUDA_OK(cudaGetDevice(&dev));
        CUB200_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const size_t smem = sizeof(double) * PACK_DOUBLES + sizeof(PDWarp) * PD_WARPS;
    CUB200_CUDA_OK(cudaFuncSetAttribute(psm_sample_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t need = (n_items + PD_WARPS - 1) / PD_WARPS;
    const int64_t cap = (int64_t)num_sms * 6;                  // 6 CTAs x 8 warps per SM (32 KB of shared memory each)
    psm_sample_direct_kernel<<<(int)(need < cap ? need : cap), PD_WARPS * 32, smem, (cudaStream_t)stream>>>(
        model_pack, mu, cov, z, seed, sample_offset, n_items, contours_out);
    return check_launch("psm_sample_direct_kernel");
}

extern "C" int cub200_psm_conditionals(const double* model_pack, const float* mu, int64_t n_img, double* mc_t, double* covc,
                                       void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!model_pack || !mu || !mc_t || !covc || n_img < 0) return CUB200_EINVAL;
    int grid = 0;
    size_t smem = 0;
    const bool wide = pf_cta_per_image(n_img);
    int rc = wide ? pf_launch_cfg(n_img, (const void*)psm_cond_kernel<PF_WARPS>, 1, &grid, &smem)
                  : pf_launch_cfg(n_img, (const void*)psm_cond_kernel<1>, PF_WARPS, &grid, &smem);
    if (rc) return rc;
    if (wide) psm_cond_kernel<PF_WARPS><<<grid, PF_WARPS * 32, smem, (cudaStream_t)stream>>>(model_pack, mu, n_img, mc_t, covc);
    else psm_cond_kernel<1><<<grid, PF_WARPS * 32, smem, (cudaStream_t)stream>>>(model_pack, mu, n_img, mc_t, covc);
    return check_launch("psm_cond_kernel");
}

extern "C" int cub200_naive_factor(const float* cov, int64_t n_img, float* a_out, void* stream) {
    using namespace cub200;
    if (n_img == 0) return 0;
    if (!cov || !a_out || n_img < 0) return CUB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    CUB200_CUDA_OK(cudaMemsetAsync(a_out, 0, sizeof(float) * P * P * n_img, st));
    const int64_t n = n_img * CUB200_K;
    naive_factor_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(cov, n_img, a_out);
    return check_launch("naive_factor_kernel");
}

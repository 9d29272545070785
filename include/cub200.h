/*
 * cub200.h -- C ABI of the B200-native uncertainty path of contouring-uncertainty.
 *
 * One shared library (libcub200.so, sm_100a).  Plain pointers and sizes only: no torch / C++ types
 * cross this boundary.  All data pointers are DEVICE pointers unless the name ends in `_host`.
 * Every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 * stream), allocates nothing, and returns 0 or a negative error code:
 *     CUB200_EINVAL (-1)  bad shape / null pointer / unsupported size
 *     CUB200_ECUDA  (-2)  a CUDA runtime error (text: cub200_last_error())
 * Buffers are owned by the caller.  Calls on distinct streams are thread-safe.
 *
 * The reference (ThierryJudge/contouring-uncertainty) has no FFI layer -- it is pure Python -- so
 * each entry point cites the Python function(s) it replaces (paths relative to the reference root,
 * cu/ = contour_uncertainty/).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Fixed geometry of this path: K = 21 landmarks (config/data/camus-cont.yaml:15), frames of
 * 256 x 256 (config/task/model/unet2.yaml:8).
 */
#ifndef CUB200_H
#define CUB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUB200_K 21          /* landmarks per contour                        */
#define CUB200_P 42          /* 2K: flattened contour dimension              */
#define CUB200_HW 256        /* frame height = width                         */
#define CUB200_ZMAX 42       /* max number of standard-normal draws / sample */
#define CUB200_MASK_WORDS 2048 /* 256 rows x 8 uint32 (bit x&31 of word x>>5) */
#define CUB200_SIMPSON_MAX_POINTS 2048 /* marching-squares contour points kept per mask */
#define CUB200_SIMPSON_MAX_SIDE 1024    /* largest side of a stretched (isotropic) mask view */

#define CUB200_EINVAL (-1)
#define CUB200_ECUDA (-2)

/* Text of the last CUDA error seen by the calling thread ("" if none). */
const char* cub200_last_error(void);
/* ABI version (bumped on any signature change) and the SM the kernels were built for (100). */
int cub200_abi_version(void);
int cub200_built_for_sm(void);
/* Number of kernels this library has launched in this process so far (all threads, all streams): what a caller subtracts
 * around a region to report how many of the library's kernels ran in it. */
int64_t cub200_launch_count(void);

/*
 * K1  DSNT heads: fused flat-softmax + coordinate expectation + 2x2 central second moments, one
 * pass over each heatmap.
 *   replaces  flat_softmax            cu/task/regression/dsnt/utils.py:71-77
 *             dsnt                    cu/task/regression/dsnt/utils.py:7-47
 *             normalized_to_pixel_coordinates  cu/task/regression/dsnt/utils.py:95-105
 *             DSNTAleatoric.predict_on_batch   cu/task/regression/dsnt/dsnt_al.py:118-131
 *             get_cov_matrix          cu/task/regression/aleatoric.py:138-144
 *   heat      (n_maps, h, w) fp32 contiguous, h = w = 256.  Logits when is_logit != 0 (softmax is
 *             fused), already-normalised heatmaps otherwise (what dsnt() receives).
 *   norm_out  optional (n_maps, 5) fp32: x, y, var_x, var_y, covar on the [-1,1] grid (= dsnt()).
 *   pix_mu    optional (n_maps, 2) fp32 pixel coordinates (x, y).
 *   pix_sigma optional (n_maps, 2, 2) fp32 pixel covariance; off-diagonals 0 when use_covar == 0
 *             (hparams.covar, dsnt_al.py:124).
 *   workspace optional: cub200_dsnt_workspace_bytes(n_maps) bytes of device memory, ZERO before its first use (the call
 *             leaves it zeroed again).  With it, small batches are cut into 2 / 4 / 8 row slabs per heatmap whose partial
 *             moments the last-arriving CTA merges, so that the launch is >= 10 waves of resident CTAs instead of ending on
 *             a nearly empty one; NULL = one CTA per heatmap.  A workspace must not be shared by concurrent calls.
 */
int cub200_dsnt_moments(const float* heat, int64_t n_maps, int h, int w, int is_logit, int use_covar,
                        float* norm_out, float* pix_mu, float* pix_sigma, void* workspace, void* stream);
int64_t cub200_dsnt_workspace_bytes(int64_t n_maps);

/*
 * Shape-model constants, packed once per model ON THE HOST (plain C++, no device work): the sampling-order
 * permutation of xbar / scaler mean / scaler scale and, for the four conditioning prefixes of the hierarchical sampler
 * (6, 10, 18 coordinates with sigma2 = 1; 34 with sigma2 = 1e-3), B^-1 = (C0[g,g] + sigma2 I)^-1, K = C0[r,g] B^-1,
 * C0[r,g] and the 2x2 diagonal blocks of C0[r,r] - K C0[g,r].
 *   c0_host (42,42) fp64: covariance of X_train about its own mean (biased, /N) -- pca() about any centre is C0 + d d^T
 *   xbar_host (42): mean of X_train; mean_host, scale_host (42): the StandardScaler of the model file
 *   pack_out_host: CUB200_PSM_PACK_DOUBLES doubles; copy them to the device and pass as `model_pack` below.
 */
#define CUB200_PSM_PACK_DOUBLES 3142
int cub200_psm_model_pack_host(const double* c0_host, const double* xbar_host, const double* mean_host,
                               const double* scale_host, double* pack_out_host);

/*
 * K2  Posterior-shape-model factor: per image, the affine map of the hierarchical PSM sampler,
 *     contour = mu + A z,  z in R^34 (17 draws of shape (1,2): points 0,10,20 | 5,15 | 2,7,13,18 |
 *     1,3,6,8,12,14,17,19),  A in R^(42 x 34).  One warp per image; Sherman-Morrison on the packed constants.
 *   replaces  pca                     cu/sampler/posterior_shape_model/posteriorshapemodel.py:9-46
 *             posterior_shape_model   .../posteriorshapemodel.py:49-81
 *             PosteriorShapeModelSampler.__call__ / sample_endo_contour / merge_priors / transform
 *                                     cu/sampler/posterior_shape_model/psm.py:73-93,199-384,424-450
 *   model_pack (CUB200_PSM_PACK_DOUBLES) fp64 device copy of cub200_psm_model_pack_host's output
 *   mu (n_img, 21, 2) fp32, cov (n_img, 21, 2, 2) fp32
 *   a_out (n_img, 34, 42) fp32, stored TRANSPOSED (draw-major) so that a sample is
 *         contour[d] = mu[d] + sum_e a_out[e][d] * z[e].
 */
int cub200_psm_factor(const double* model_pack, const float* mu, const float* cov, int64_t n_img, float* a_out, void* stream);

/*
 * Independent-Gaussian (naive) sampler factor: block-diagonal A (42 draws), 2x2 Cholesky colouring.
 *   replaces  Sampler.sample_points / NaiveSampler.__call__   cu/sampler/sampler.py:22-41, naive.py:13-25
 *   a_out (n_img, 42, 42) fp32 transposed as above; draw slot of landmark k = (2k, 2k+1).
 *   (numpy colours with an SVD factor; for draw-replay parity the host passes its own A.)
 */
int cub200_naive_factor(const float* cov, int64_t n_img, float* a_out, void* stream);

/*
 * K3  Monte-Carlo sampler -> rasteriser -> streaming aggregation.  contour = mu + A z (one thread per sample), then
 *     one warp per sample:  round  ->  cubic spline (fp64, 1000 points)  ->  pixel set + closing
 *     line  ->  4-connected background flood from the frame border  ->  per-sample area and
 *     run-boundary increments into the image's difference map.  Masks never leave shared memory.
 *   replaces  AleatoricUncertaintyTask.sample          cu/task/regression/aleatoric.py:54-78
 *             ContourUncertaintyTask.convert_to_mask   cu/task/regression/contour_uncertainty.py:26-57
 *             USContourToMask.__call__ (LV branch)     cu/data/camus/utils.py:31-98
 *             reconstruction / contour_spline          cu/utils/contour.py:9-40
 *             lv_area                                  cu/utils/clinical.py:92-93
 *   mu (n_img, 42) fp32; a_t (n_img, zdim, 42) fp32 (transposed factor), zdim <= 42
 *   z  optional (n_img, n_samples, zdim) fp32 standard-normal draws to replay; NULL -> Philox4x32-10
 *      + Box-Muller keyed by (seed, sample_offset + image * n_samples + sample): a caller that cuts a batch into
 *      chunks passes the index of the chunk's first sample, so the draws do not depend on the chunking
 *   contours_out (n_img, n_samples, 21, 2) fp32, REQUIRED: the API's contour_samples, and the hand-over between the
 *                sampling kernel (one thread per sample) and the rasteriser (one warp per sample).  When area_out,
 *                diff and masks_out are all NULL the call only samples (no rasterisation).
 *   area_out     optional (n_img, n_samples) int32  mask pixel counts
 *   diff         (n_img, 256, 256) int32, ACCUMULATED into (caller zeroes): +1 where a mask run
 *                starts, -1 one past where it ends; the row prefix sum is the count map
 *   masks_out    optional (n_img, n_samples, 2048) uint32 bit-packed masks (debug / return_masks)
 *   round_landmarks: 1 = USContourToMask semantics (round landmarks first), 0 = reconstruction()
 *                on the raw float contour (cu/sampler/sampling_example3.py:127); 3 = linear_reconstruction
 *                (cu/utils/contour.py:43-53, reconstruction_type='linear'): the rounded landmarks joined by
 *                skimage.draw.line segments instead of the spline.  Bit 0 = round, bit 1 = linear.
 *   images_per_map: consecutive (mu, a_t) entries that aggregate into ONE difference map (the T_e
 *                epistemic passes of an image, aleatoric.py:71-73); diff is (n_img / images_per_map, 256, 256)
 *   workspace:   cub200_raster_workspace_bytes(n_img * n_samples) bytes of device memory (not needed by a sample-only
 *                call): the queue of samples whose fill needs the general flood (second kernel of the rasteriser)
 *   step_dev:    optional DEVICE pointer to a step counter: *step_dev * n_img * n_samples is added to sample_offset on the
 *                device, so a step captured in a CUDA graph draws fresh samples at every replay (the argument list of a
 *                captured launch is frozen; the counter is advanced inside the graph with cub200_step_counter_add).
 */
int64_t cub200_raster_workspace_bytes(int64_t n_contours);
int cub200_sample_raster_aggregate(const float* mu, const float* a_t, int zdim, const float* z,
                                   uint64_t seed, uint64_t sample_offset, int64_t n_img, int64_t n_samples,
                                   int round_landmarks, float* contours_out, int32_t* area_out,
                                   int32_t* diff, int64_t images_per_map, uint32_t* masks_out, void* workspace,
                                   const uint64_t* step_dev, void* stream);
/* *counter += inc, as a (capturable) kernel launch on `stream`. */
int cub200_step_counter_add(uint64_t* counter, uint64_t inc, void* stream);

/*
 * Rasteriser alone (the same device code as K3, contours given): bit-packed masks out.
 *   contours (n, 21, 2) fp32; masks_out (n, 2048) uint32 optional; area_out (n) int32 optional;
 *   diff optional (n_img, 256, 256) int32 accumulated, image of contour i = i / per_image.
 *   workspace: cub200_raster_workspace_bytes(n) bytes of device memory.
 */
int cub200_rasterize_contours(const float* contours, int64_t n, int round_landmarks, uint32_t* masks_out,
                              int32_t* area_out, int32_t* diff, int64_t per_image, void* workspace, void* stream);

/* Same as cub200_rasterize_contours with fp64 contours (n, 21, 2): the analytic skew u-map reconstructs unrounded
 * fp64 level contours (cu/utils/skew_umap.py:62-64). */
int cub200_rasterize_contours_f64(const double* contours, int64_t n, int round_landmarks, uint32_t* masks_out,
                                  int32_t* area_out, int32_t* diff, int64_t per_image, void* workspace, void* stream);

/*
 * K4  Finalise the per-image maps from the difference map.
 *   replaces  pred = pred_samples.mean((1,2)).round()   cu/task/regression/aleatoric.py:111
 *             UncertaintyTask.sample_entropy            cu/task/uncertainty.py:107-133
 *   diff (n_img,256,256) int32; n_samples = T (samples aggregated per image)
 *   count_out optional (n_img,256,256) int32; mean_out optional fp64 = count/T;
 *   entropy_out optional fp64 binary entropy in bits (scipy.stats.entropy([p,1-p], base=2));
 *   pred_out optional uint8 = round-half-even(count/T)
 *   lut_ws: device workspace of 3 * (n_samples + 1) doubles (entropy, mean and variance of c/T, c = 0..T), written by the call
 *           unless lut_ready != 0 (then it must hold the tables of the same n_samples: cub200_count_tables);
 *           needed when any of mean_out / entropy_out / variance_out is given.
 *   variance_out optional fp64 = p (1 - p), p = count/T: the per-pixel variance of the 0/1 samples (the reference has no
 *                variance map; this is the Bernoulli variance its mean map implies)
 *   count16_out optional uint16 count map (n_samples <= 65535): mean / entropy / pred / variance are functions of the count,
 *                so a host that wants them can fetch 2 B / pixel and index a (T+1)-entry table
 */
int cub200_finalize_maps(const int32_t* diff, int64_t n_img, int64_t n_samples, int32_t* count_out,
                         double* mean_out, double* entropy_out, uint8_t* pred_out, double* lut_ws, int lut_ready,
                         double* variance_out, uint16_t* count16_out, void* stream);
/* The tables alone (one small launch): a caller that finalises many batches with the same n_samples fills lut_ws once and
 * passes lut_ready = 1 to cub200_finalize_maps afterwards. */
int cub200_count_tables(int64_t n_samples, double* lut_ws, void* stream);

/*
 * Soft masks of convert_to_mask(soft_mask=True): Gaussian blur (sigma 5, truncate 1: 11 taps, edges repeated) of every
 * bit-packed mask, then (m - min) / (max - min).
 *   replaces  skimage.filters.gaussian(mask, sigma=(5,5), truncate=1) + the min-max normalisation
 *             cu/task/regression/contour_uncertainty.py:44-49   (scikit-image is absent from the build image: restated as the
 *             scipy.ndimage.gaussian_filter call it wraps -- parity unpinned for that third-party function)
 *   masks (n, 2048) u32; weights (6) fp64: the normalised Gaussian taps w[0] (centre) .. w[5], formed on the host as scipy
 *   forms them; out (n, 256, 256) fp64.
 */
int cub200_soft_masks(const uint32_t* masks, int64_t n, const double* weights, double* out, void* stream);

/*
 * Aleatoric / epistemic split of Monte-Carlo metric distributions, batched: one result per item (instant, view or patient).
 *   replaces  aleatoric_epistemic_uncertainty            cu/results/clinical/utils.py:3-18
 *             the sample_reject rule of FAC / GLS / Volume.compute   cu/results/clinical/view.py:52-53,165-166,
 *                                                                     cu/results/clinical/patient.py:192-193
 *   values (n_items, t_e, t_a) fp64, NaN allowed.  apply_range != 0: v < lo or v > hi -> rejected (becomes NaN; a NaN input
 *   is not "rejected", as numpy's comparisons are false for it).
 *   values_out optional (n_items, t_e, t_a) fp64: the values with the rejected ones as NaN ('mc' of the reference's dicts)
 *   reject_out optional (n_items, t_e, t_a) uint8: 'sample_reject';  n_reject_out optional (n_items) int32: their number
 *   stats_out (n_items, 4) fp64: mean = nanmean_e(nanmean_a), aleatoric = nanmean_e(nanstd_a), epistemic = nanstd_e(nanmean_a),
 *   total = epistemic + aleatoric (standard deviations, ddof 0, whatever the reference's variable names say).
 */
int cub200_metric_ae_split(const double* values, int64_t n_items, int64_t t_e, int64_t t_a, int apply_range, double lo, double hi,
                           double* values_out, uint8_t* reject_out, double* stats_out, int32_t* n_reject_out, void* stream);

/*
 * Per-view clinical reductions from per-sample areas.
 *   replaces  lv_FAC / FAC.compute (per-sample part)    cu/utils/clinical.py:11-29,
 *                                                      cu/results/clinical/view.py:35-78
 *   area_ed, area_es (n) int32 -> fac_out (n) fp64 = (ed - es)/ed, NaN when outside [0,1].
 */
int cub200_fac_from_areas(const int32_t* area_ed, const int32_t* area_es, int64_t n, double* fac_out,
                          void* stream);

/*
 * Dataset-level accumulators of a per-sample integer metric (LV areas): sums_out[0] = sum v, sums_out[1] = sum v^2,
 * exact 64-bit integers (what the ranks all_reduce at the end of a step; mean / std of the pooled distribution follow).
 *   replaces  the np.mean / np.std reductions over gathered per-sample metrics   cu/results/clinical/utils.py:3-18
 */
int cub200_metric_sums(const int32_t* values, int64_t n, int64_t* sums_out, void* stream);
/*
 * What a rank buffers per step for the epoch-end exchange (cu/results/clinical/clinical_metrics.py:51-66 reads every
 * BatchResult at the END of the predict epoch), in one launch: slot = ring + s * (n + 4 * CUB200_METRIC_PACK_PARTS) receives
 * the n values and, in its tail, CUB200_METRIC_PACK_PARTS partial (sum, sum of squares) pairs as exact int64 (two int32 words
 * each, low word first): the consumer adds the parts.  s = *step_dev % n_slots (device-side step counter, see
 * cub200_sample_raster_aggregate) or 0 when step_dev is NULL.  ring (n_slots, n + 4 * CUB200_METRIC_PACK_PARTS) int32.
 */
#define CUB200_METRIC_PACK_PARTS 16
int cub200_metric_pack(const int32_t* values, int64_t n, int32_t* ring, int64_t n_slots, const uint64_t* step_dev,
                       void* stream);

/*
 * Posterior moments of the sampled contours: per (image, landmark) the sample mean and the unbiased
 * 2x2 sample covariance over the T samples.
 *   replaces  post_mu = contour_samples.mean(axis=2); post_cov = np.cov(...)   cu/task/regression/aleatoric.py:96-102
 *   contours (n_img, n_samples, 21, 2) fp32 -> mean_out (n_img, 21, 2) fp64, cov_out (n_img, 21, 2, 2) fp64 (ddof = 1)
 */
int cub200_sample_moments(const float* contours, int64_t n_img, int64_t n_samples, double* mean_out,
                          double* cov_out, void* stream);

/*
 * Sequence (ED/ES-coupled) sampler: condition the second instant of every pair on the sampled first one.
 *   replaces  SequencePSMSampler.sample_two_contours, lines 68-91
 *             cu/sampler/posterior_shape_model/sequence_sampler.py  (+ merge_priors psm.py:424-440)
 *   mu (n_pairs, 2, 21, 2), cov (n_pairs, 2, 21, 2, 2) fp32: predicted distributions of (instant 0, instant 1)
 *   first_samples (n_pairs * 2, n_samples, 21, 2) fp32: a PSM sample of EVERY image (K2 + K3 sample-only)
 *   first_instant (n_pairs, n_samples) uint8: which instant is sampled first (random.randint(0,1), :47)
 *   m2t (2, 42, 42) fp64: [b][e][row] = (S[o,g] (S[g,g] + I)^-1)[row][e] of the 84-dim model, g = dims of
 *        instant b, o = the other instant;  covc (2, 21, 2, 2) fp64: 2x2 blocks of S[o,o] - M S[g,o], columns
 *        scaled by seq_scale;  seq_mu, seq_mean, seq_scale (84) fp64
 *   first_out (n_pairs, n_samples, 21, 2): the first-sampled contour;  mu_f_out (n_pairs, n_samples, 21, 2),
 *   cov_f_out (n_pairs, n_samples, 21, 2, 2): merged distribution of the OTHER instant, to be sampled by a
 *   second pass of K2 + K3 (sequence_sampler.py:94).
 */
int cub200_seq_condition(const float* mu, const float* cov, const float* first_samples, const uint8_t* first_instant,
                         int64_t n_pairs, int64_t n_samples, const double* m2t, const double* covc,
                         const double* seq_mu, const double* seq_mean, const double* seq_scale,
                         float* first_out, float* mu_f_out, float* cov_f_out, void* stream);

/*
 * Skew-normal grid pdf tables: table[d][x][y] = 2 N((x,y); mu_d, cov_d) (Phi(alpha_d^T cov_d^-1/2 ((x,y) - mu_d)) + 1e-7)
 * on the integer grid x, y = 0..255 (torch.meshgrid(indexing='ij'), psm_skew.py:181-185).
 *   replaces  BivariateSkewNormal.pdf / logpdf / affine / unit_normal_logcdf  cu/distributions/bivariateskewnormal.py:20-49
 *             BivariateNormal.logpdf                                          cu/distributions/bivariatenormal.py:15-32
 *             batch_matrix_pow(cov, -0.5)                                     cu/distributions/utils.py:100-129
 *             the `pdfs` loop of sample_one_instant                           cu/sampler/posterior_shape_model/psm_skew.py:234-238
 *   mu (n_dist,2), cov (n_dist,2,2), alpha (n_dist,2) fp32; pdf_out (n_dist,256,256) fp32.
 *   flags: bit 0 applies the sampler's alpha * [1,-1] (psm_skew.py:232); bit 1 (n_dist a multiple of 21: whole images)
 *   leaves the tables of landmarks 4, 9, 11, 16 unwritten -- the samplers never read them (posterior-mean-filled points).
 */
int cub200_skew_pdf_grid(const float* mu, const float* cov, const float* alpha, int64_t n_dist, int flags,
                         float* pdf_out, void* stream);

/*
 * Bivariate normal / skew-normal pdf or logpdf at arbitrary points.
 *   replaces  BivariateNormal.logpdf / pdf            cu/distributions/bivariatenormal.py:15-36
 *             BivariateSkewNormal.logpdf / pdf        cu/distributions/bivariateskewnormal.py:20-49
 *   x (n_points,2) fp32; mu (n_dist,2), cov (n_dist,2,2), alpha (n_dist,2) fp32, alpha NULL = plain normal;
 *   n_dist = 1 (one law for all points) or n_points (paired); out (n_points) fp32.
 */
int cub200_bivariate_pdf(const float* x, int64_t n_points, const float* mu, const float* cov, const float* alpha,
                         int64_t n_dist, int want_log, float* out, void* stream);

/*
 * n draws of one bivariate skew-normal law.
 *   replaces  BivariateSkewNormal.rvs_fast            cu/distributions/bivariateskewnormal.py:160-191
 *   mu (2), cov (2,2), alpha (2) fp32; z optional (n,3) fp32 standard normals to replay (MultivariateNormal.sample of the
 *   3x3 augmented law), else Philox with `seed`; out (n,2) fp32.
 */
int cub200_skew_rvs_fast(const float* mu, const float* cov, const float* alpha, const float* z, uint64_t seed, int64_t n,
                         float* out, void* stream);

/*
 * ONE PSM sample per item without building its factor: the second instant of the sequence sampler draws a single contour
 * from every per-(pair, sample) merged law           cu/sampler/posterior_shape_model/sequence_sampler.py:94, psm.py:73-93
 * (equivalent to cub200_psm_factor + cub200_sample_raster_aggregate(n_samples = 1, sample-only) with the same seed and
 * sample_offset -- same draws -- at 1/34 of the arithmetic and without the 5.7 KB factor per item).
 *   mu (n_items,21,2), cov (n_items,21,2,2) fp32; z (n_items,34) fp32 replayed draws or NULL (Philox keyed by
 *   (seed, item + sample_offset)); contours_out (n_items,21,2) fp32.
 */
int cub200_psm_sample_direct(const double* model_pack, const float* mu, const float* cov, const float* z, uint64_t seed,
                             uint64_t sample_offset, int64_t n_items, float* contours_out, void* stream);

/*
 * Conditional maps of the posterior shape model per image (the part of compute_psm that does not depend
 * on the sampled values): mu_c = mu + Mc (s - mu) over the already-sampled coordinates, and the 2x2
 * conditional covariances of the 14 level points.
 *   replaces  pca + posterior_shape_model + compute_psm   cu/sampler/posterior_shape_model/posteriorshapemodel.py:9-81,
 *                                                         psm_skew.py:228,469-491 (and the sigma2 = 1e-3 fill, :384-395)
 *   mu (n_img,21,2) fp32;  mc_t (n_img,34,36) fp64 = [e][v-6] in sampling order (points 0,10,20 | 5,15 | 2,7,13,18 |
 *   1,3,6,8,12,14,17,19 | 4,9,11,16; variable v = 2*position + coordinate);  covc (n_img,14,3) fp64 = (c00,c10,c11).
 */
int cub200_psm_conditionals(const double* model_pack, const float* mu, int64_t n_img, double* mc_t, double* covc, void* stream);

/*
 * Skew-normal posterior-shape-model sampler, one warp per sample.
 *   replaces  SkewPosteriorShapeModelSampler.__call__ / sample_one_instant / sample_contour   psm_skew.py:187-411
 *             numerical_sampling (torch.multinomial == exponential race argmax(p / q))       psm_skew.py:45-158
 *             BivariateSkewNormal.rvs_fast                                                   bivariateskewnormal.py:160-191
 *   mu, cov, alpha (n_img,21,..) fp32 as predicted (the kernel applies alpha * [1,-1]); pdfs (n_img,21,256,256)
 *   from cub200_skew_pdf_grid(flip_alpha_y = 1); mc_t, covc from cub200_psm_conditionals.
 *   z    optional (n_img, n_samples, 3, 3) fp32: the (1,3) normals of rvs_fast for points 0, 10, 20
 *   expo optional (n_img, n_samples, 14, 65536) fp32 Exp(1) fields of the 14 multinomial draws (level order);
 *        NULL -> Philox4x32-10 keyed by (seed, sample, draw, cell)
 *   contours_out (n_img, n_samples, 21, 2) fp32; indices_out optional (n_img, n_samples, 14) int32 flat grid
 *   index x*256 + y of every draw, -1 where the product pdf vanished and mu_c was taken (psm_skew.py:135-154).
 */
int cub200_skew_sample(const float* mu, const float* cov, const float* alpha, const float* pdfs, const double* mc_t,
                       const double* covc, const float* z, const float* expo, uint64_t seed, int64_t n_img,
                       int64_t n_samples, float* contours_out, int32_t* indices_out, void* stream);

/*
 * Skew-normal SEQUENCE sampler (ED/ES pairs): per (pair, first instant) the conditional law of the other instant under
 * the 84-dim sequence shape model re-centred on the predicted pair.
 *   replaces  pca(seq_X_train, sequence_transform(mu)) + posterior_shape_model(.., sigma2=1) + rescaling
 *             cu/sampler/posterior_shape_model/psm_skew_sequence.py:66,72-86
 *   seq_c0 (84,84), seq_xbar, seq_mean, seq_scale (84) fp64; mu (n_pairs,2,21,2) fp32
 *   m_seq (n_pairs,2,42,42) fp64 [pair][b][e][row]; cov_seq (n_pairs,2,21,4) fp64 [pair][b][landmark][2x2]
 */
int cub200_seq_skew_conditionals(const double* seq_c0, const double* seq_xbar, const double* seq_mean,
                                 const double* seq_scale, const float* mu, int64_t n_pairs, double* m_seq, double* cov_seq,
                                 void* stream);

/*
 * Skew-normal SEQUENCE sampler: first a skew-PSM sample of the randomly chosen instant, then the other instant from
 * pdfs = normalise(skew pdf x N(grid; sequence conditional)) -- its three initial points by a multinomial over pdfs, its
 * level points over pdfs x N(grid; shape-model conditional).
 *   replaces  SequenceSkewPSMSampler.__call__ / sample_two_contours   psm_skew_sequence.py:34-104
 *   mu, cov, alpha (n_pairs,2,21,..) fp32; pdfs_flipped / pdfs_raw (2 n_pairs,21,256,256): cub200_skew_pdf_grid with
 *   flip_alpha_y = 1 / 0 (the second instant's tables use alpha as given, :91); mc_t, covc per image
 *   (cub200_psm_conditionals on the 2 n_pairs images); m_seq, cov_seq from cub200_seq_skew_conditionals
 *   first_instant (n_pairs, n_samples) uint8 (random.randint(0,1), :45); z optional (n_pairs, n_samples, 3, 3);
 *   expo optional (n_pairs, n_samples, 31, 65536): 14 races of the first instant, then 3 + 14 of the second
 *   contours_out (n_pairs, 2, n_samples, 21, 2) fp32; indices_out optional (n_pairs, n_samples, 31) int32
 */
int cub200_seq_skew_sample(const float* mu, const float* cov, const float* alpha, const float* pdfs_flipped,
                           const float* pdfs_raw, const double* mc_t, const double* covc, const double* m_seq,
                           const double* cov_seq, const uint8_t* first_instant, const float* z, const float* expo,
                           uint64_t seed, int64_t n_pairs, int64_t n_samples, float* contours_out, int32_t* indices_out,
                           void* stream);

/*
 * Simpson's biplane rule, step 1: per mask, the marching-squares contour (find_contours(mask, 0.5)[0]), the longest
 * admissible base chord, the apex and the 20 disc diameters.
 *   replaces  _compute_diameters (+ helpers)      vital/vital/metrics/evaluate/clinical/heart_us.py:74-167
 *             skimage.measure.find_contours        (third party, restated; oracle/port.py::find_contours)
 *   masks (n_masks, 2048) uint32 bit-packed (cub200_rasterize_contours / cub200_sample_raster_aggregate), 16-byte aligned
 *   spacing: isotropic pixel size in mm (anisotropic spacing = PIL nearest-neighbour stretch first: not done here)
 *   diam_out (n_masks, 20) fp64 mm; step_out (n_masks) fp64 = mid-line length * spacing / 20; both NaN when the
 *   reference would raise (empty mask, no admissible chord, contour longer than CUB200_SIMPSON_MAX_POINTS)
 *   npts_out optional (n_masks) int32 contour length (0 none, -1 too long)
 *   workspace: cub200_simpson_workspace_bytes(n_masks) bytes of device memory, 16-byte aligned
 */
int64_t cub200_simpson_workspace_bytes(int64_t n_masks);
int cub200_simpson_diameters(const uint32_t* masks, int64_t n_masks, double spacing, double* diam_out, double* step_out,
                             int32_t* npts_out, void* workspace, void* stream);

/*
 * The same on masks with ANISOTROPIC voxel spacing: the reference first stretches the mask to isotropic spacing with PIL's
 * nearest-neighbour resize (resize_image_to_isotropic, vital/vital/utils/image/transform.py:52-66; heart_us.py:117).  The
 * stretched mask is never materialised: row_index (height) / col_index (width) uint16 device arrays give the source row /
 * column of every row / column of the stretched view (PIL's ImagingScaleAffine rule: index = (int)(0.5 a + k a) with the
 * sum formed by repeated double additions, a = 256 / size); height, width <= CUB200_SIMPSON_MAX_SIDE; spacing = the
 * isotropic (smaller) spacing.
 */
int cub200_simpson_diameters_stretched(const uint32_t* masks, int64_t n_masks, const uint16_t* row_index,
                                       const uint16_t* col_index, int height, int width, double spacing, double* diam_out,
                                       double* step_out, int32_t* npts_out, void* workspace, void* stream);

/*
 * Simpson's biplane rule, step 2: EDV / ESV (mL, integers stored as fp64) and EF per Monte-Carlo sample.
 *   replaces  compute_left_ventricle_volumes / _compute_left_ventricle_volume_by_instant   heart_us.py:13-71
 *             the sample loop of Volume.compute                cu/results/clinical/patient.py:166-190
 *   diam (n_patients, 4, n_samples, 20), step (n_patients, 4, n_samples): masks ordered [A2C-ED, A2C-ES, A4C-ED, A4C-ES]
 *   edv_out, esv_out, ef_out (n_patients, n_samples) fp64; a sample with any NaN input stays (0, 0, 0) like the
 *   reference's bare except; EF outside [0, 1] -> NaN.
 */
int cub200_simpson_volumes(const double* diam, const double* step, int64_t n_patients, int64_t n_samples,
                           double* edv_out, double* esv_out, double* ef_out, void* stream);

/*
 * Projected per-point uncertainty: spline normal at every landmark and the standard deviation of the point's law
 * along it (skew-normal: also the projected skewness).
 *   replaces  projected_uncertainty / projected_uncertainty_value     cu/utils/uncertainty_projection.py:11-125
 *             BivariateNormal.marginal, BivariateSkewNormal.marginal  cu/distributions/bivariatenormal.py:67-86,
 *                                                                     bivariateskewnormal.py:92-135
 *   mu (n_img,21,2), cov (n_img,21,2,2) fp32; alpha (n_img,21,2) fp32 or NULL
 *   all_points: `all=True` (else base / apex points 0, 10, 20 get sum(sqrt(eig(cov))));  linear_close as the reference
 *   unc_out (n_img,21) fp32; proj_out (n_img,21,2) fp64; alpha_proj_out (n_img,21) fp32 (required with alpha)
 *   projected_uncertainty_value = the sum of unc_out over the 21 points with all_points = 0.
 */
int cub200_projected_uncertainty(const float* mu, const float* cov, const float* alpha, int64_t n_img, int all_points,
                                 int linear_close, float* unc_out, double* proj_out, float* alpha_proj_out, void* stream);
/* The same for contours of n_points landmarks, 4 <= n_points <= CUB200_RASTER_MAX_POINTS (the 44 / 50 / 26-point lung and
 * heart contours of LungUMap / LungSkewUmap, cu/data/lung/utils.py:72-114; base / apex points 0, n_points / 2, n_points - 1). */
int cub200_projected_uncertainty_generic(const float* mu, const float* cov, const float* alpha, int64_t n_img, int n_points,
                                         int all_points, int linear_close, float* unc_out, double* proj_out,
                                         float* alpha_proj_out, void* stream);

/*
 * Analytic uncertainty map: 100 offset contours mu + v u i, i = linspace(-2, 2, 100); a pixel takes the pdf value
 * of the LAST offset whose spline / closing line touches it.
 *   replaces  uncertainty_map        cu/utils/umap.py:10-33;   USUMap.__call__ (LV)   cu/data/camus/utils.py:101-124
 *   unc, proj from cub200_projected_uncertainty(all_points = 1); pdf_table (100) fp64 = norm.pdf(linspace(-2,2,100));
 *   normalize != 0 divides by the map maximum (USUMap); kmap_ws (n_img,256,256) int32 workspace;
 *   umap_out (n_img,256,256) fp64.
 */
int cub200_umap_analytic(const float* mu, const float* unc, const double* proj, int64_t n_img, const double* pdf_table,
                         int normalize, int32_t* kmap_ws, double* umap_out, void* stream);

/*
 * Skew u-map, step 1: per landmark the projected skew-normal on a 1000-point axis, its mode and the 2 x 100 level
 * offsets -> 200 level contours per image.
 *   replaces  skew_umap lines 13-58   cu/utils/skew_umap.py
 *   unc, proj, alpha_proj from cub200_projected_uncertainty[_generic](alpha, all_points = 1, linear_close as wanted)
 *   n_points landmarks per contour (21 for the LV; 1..CUB200_RASTER_MAX_POINTS)
 *   contours_out (n_img,200,n_points,2) fp64; mode_out (n_img,n_points,2) fp32 (projected_mode)
 */
int cub200_skew_umap_contours(const float* mu, const float* unc, const double* proj, const float* alpha_proj,
                              int64_t n_img, int n_points, double* contours_out, float* mode_out, void* stream);

/*
 * Skew u-map, step 2: weighted mean of bit-packed masks -> natural-log binary entropy (scipy.stats.entropy), optionally
 * divided by its maximum.
 *   replaces  skew_umap lines 76-80   cu/utils/skew_umap.py;   USSkewUmap.__call__ (LV)   cu/data/camus/utils.py:127-153
 *   masks (n_img, n_masks, 2048) uint32; weights (n_masks) fp64; weight_sum = their sum as np.average forms it
 */
int cub200_weighted_mask_entropy(const uint32_t* masks, const double* weights, int64_t n_img, int n_masks,
                                 double weight_sum, int normalize, double* out, void* stream);

/*
 * Spline arc length of sampled contours (longitudinal length of the endocardium): contour_spline(contour) with 1001
 * points on the unrounded landmarks, sum of the 1000 chord lengths.  GLS of a sample = (len_ED - len_ES) / len_ED.
 *   replaces  perimeter, global_longitudinal_strain, compute_gls   cu/utils/clinical.py:30-78
 *   contours (n, 21, 2) fp32 -> perimeter_out (n) fp64 (pixels)
 */
int cub200_contour_perimeter(const float* contours, int64_t n, double* perimeter_out, void* stream);

/*
 * Dataset-level pixel calibration bins (the accumulators the ranks all_reduce).
 *   replaces  Calibration.get_pixel_confidences_and_accuracies / get_confidences_and_accuracies / ece (bin loop)
 *             cu/results/calibration.py:94-113,197-250
 *   umap (n_pixels) fp64 uncertainty (confidence = 1 - umap); pred, gt (n_pixels) uint8 labels; pixels with pred + gt == 0
 *   are dropped; bin b holds the pixels with bounds[b] < confidence < bounds[b+1] (strict, as the reference).
 *   bounds (nb_bins + 1) fp64 (np.linspace(0, 1, nb_bins + 1)); nb_bins <= 64
 *   bins_out ((nb_bins + 1) * 3) fp64: per bin (count, sum confidence, sum accuracy); the last triple's first entry is the
 *   number of foreground pixels (the denominator of prop_in_bin).  workspace: cub200_calibration_workspace_bytes().
 */
int64_t cub200_calibration_workspace_bytes(void);
int cub200_calibration_bins(const double* umap, const uint8_t* pred, const uint8_t* gt, int64_t n_pixels, const double* bounds,
                            int nb_bins, double* bins_out, void* workspace, void* stream);

/*
 * Segmentation-sample post-processing: fill holes, keep the largest blob, entropy of the masked mean.
 * Replaces the per-sample loop of SegmentationUncertaintyTask.predict_step
 *             cu/task/segmentation/segmentation_uncertainty.py:98-140  (ndimage.binary_fill_holes + big_blob,
 *             cu/task/segmentation/utils.py:12-32) and UncertaintyTask.sample_entropy (C == 1) cu/task/uncertainty.py:107-133.
 *
 * cub200_seg_largest_blob: probs (n_samples, 256, 256) fp32, 16-byte aligned.  A pixel is foreground iff round(p) != 0,
 *   i.e. |p| > 0.5.  masks_out (n_samples, 2048) u32 = the largest 4-connected component of the hole-filled foreground,
 *   bit b of word 8 r + w = pixel (r, 32 w + b); ties go to the component met first in raster order (np.argmax).
 *   area_out (n_samples) i32 = its pixel count, may be NULL.  workspace: cub200_seg_blob_workspace_bytes(n_samples).
 * cub200_seg_entropy: probs (n_img, n_samples, 256, 256) fp32 and the masks of cub200_seg_largest_blob ->
 *   entropy_out (n_img, 256, 256) fp32 = binary entropy (bits) of mean_t(probs * mask), non-finite -> 0, a frame of
 *   `pad` pixels zeroed (the reference uses 10).  Optional (NULL to skip): mean_out (n_img, 256, 256) fp32,
 *   pred_samples_out (n_img * n_samples, 2048) u32 = round(probs * mask) != 0 (BatchResult.pred_samples),
 *   samples_out (n_img, n_samples, 256, 256) fp32 = probs * mask.
 */
int64_t cub200_seg_blob_workspace_bytes(int64_t n_samples);
int cub200_seg_largest_blob(const float* probs, int64_t n_samples, uint32_t* masks_out, int32_t* area_out, void* workspace,
                            void* stream);
int cub200_seg_entropy(const float* probs, const uint32_t* masks, int64_t n_img, int n_samples, int pad, float* entropy_out,
                       float* mean_out, uint32_t* pred_samples_out, float* samples_out, void* stream);

/*
 * Mask-based longitudinal length of the LV endocardium: what EchoMeasure.gls compares between the
 * frames of a sequence.  Replaces EchoMeasure.gls / _lv_longitudinal_length (myo_labels=None),
 * _extract_landmarks_from_polar_contour (apex=False), _endo_epi_contour   vital/vital/utils/image/us/measure.py:22-195,320-366
 * and the per-sample loop of GLS.compute   contour_uncertainty/results/clinical/view.py:146.
 *   masks (n_masks, 2048) u32 bit-packed 256x256; smoothing_factor = polar_smoothing_factor (the reference uses 5e-3; 0 = none);
 *   spacing_row / spacing_col = voxelspacing (the reference's callers leave it at (1, 1)).
 *   gauss_table: optional device table (CUB200_SIMPSON_MAX_POINTS + 1, CUB200_GLS_TABLE_STRIDE) f64; row n holds the
 *   normalised Gaussian kernel values w[0..lw] scipy.ndimage.gaussian_filter1d uses for sigma = n * smoothing_factor
 *   (lw = int(4 sigma + 0.5)), made by the host with the same exp as scipy so that the smoothed signal, hence every
 *   peak comparison, is bit-identical.  NULL: the kernel evaluates the weights itself (its exp may differ by an ulp,
 *   which can only matter for peaks that tie to the last bit, e.g. on mirror-symmetric masks).
 *   length_out (n_masks) f64: sum of the segment lengths of the base-less contour; NaN where the reference raises
 *   (no contour, fewer than two base corners, more than CUB200_SIMPSON_MAX_POINTS contour points).
 *   Optional: corners_out (n_masks, 2, 2) f64 = (left, right) base corners as (row, col); npts_out (n_masks) i32 = points
 *   of the base-less contour.  workspace: cub200_simpson_workspace_bytes(n_masks).  masks and workspace 16-byte aligned.
 * GLS of frame k vs frame 0 = (L_k - L_0) / L_0 * 100.
 */
#define CUB200_GLS_TABLE_STRIDE 48
int cub200_mask_lv_length(const uint32_t* masks, int64_t n_masks, double smoothing_factor, const double* gauss_table,
                          double spacing_row, double spacing_col, double* length_out, double* corners_out, int32_t* npts_out,
                          void* workspace, void* stream);

/*
 * The same length with the base markers of EchoMeasure._endo_base            vv/utils/image/us/measure.py:177-222
 * (EchoMeasure.gls with myo_labels, measure.py:339-352; GLS.gls passes Label.MYO when the view's labels hold it,
 * cu/results/clinical/view.py:90-94): the LV pixels whose 8-neighbourhood meets both the myocardium and the rest, split at
 * their mean column (mean row when they lie in one column), the lowest pixel of each half.
 *   lv_masks, myo_masks (n_masks, 2048) u32 bit-packed (e.g. the two outputs of cub200_us_myo_masks).
 *   cub200_endo_base_corners: corners_out (n_masks, 4) i32 = (left row, left col, right row, right col); -1 where the
 *   reference raises (fewer than two frontier pixels).
 *   cub200_mask_lv_length_myo: corners_out as above (required: also the staging buffer), length_out (n_masks) f64 = length
 *   of the LV contour (level 0.9) from the point closest to the left marker to the point closest to the right one (first
 *   minimum, np.argmin), NaN where the reference raises; optional npts_out (n_masks) i32.
 *   workspace: cub200_simpson_workspace_bytes(n_masks).
 */
int cub200_endo_base_corners(const uint32_t* lv_masks, const uint32_t* myo_masks, int64_t n_masks, int32_t* corners_out,
                             void* stream);
int cub200_mask_lv_length_myo(const uint32_t* lv_masks, const uint32_t* myo_masks, int64_t n_masks, double spacing_row,
                              double spacing_col, double* length_out, int32_t* corners_out, int32_t* npts_out,
                              void* workspace, void* stream);

/*
 * Rasteriser for contours with any number of points: reconstruction / linear_reconstruction   cu/utils/contour.py:28-53
 * as LungContourToMask (cu/data/lung/utils.py:25-69) uses them per structure (44 / 50 / 26 points).
 *   contours (n, n_points, 2) (x, y), fp32 (is_f64 = 0) or fp64 (is_f64 = 1); 4 <= n_points <= CUB200_RASTER_MAX_POINTS;
 *   flags as round_landmarks of cub200_rasterize_contours (bit 0 = round the points first, bit 1 = linear).
 *   masks_out (n, 2048) u32 bit-packed and / or area_out (n) i32.
 */
#define CUB200_RASTER_MAX_POINTS 64
int cub200_rasterize_contours_generic(const void* contours, int is_f64, int64_t n, int n_points, int flags, uint32_t* masks_out,
                                      int32_t* area_out, void* stream);

/*
 * Two-structure (LV + MYO) masks: USContourToMask when Label.MYO is among the labels   cu/data/camus/utils.py:47-82
 * (split_landmarks, contour_spline x 2, poly2mask -> skimage.draw.polygon [restated, parity unpinned], reconstruction).
 *   landmarks (n, 42, 2) fp32 (x, y): 21 endocardial then 21 epicardial points, rounded half to even by the call;
 *   flags: bit 1 = reconstruction_type 'linear' for the LV mask (the polygon always follows the splines).
 *   lv_out / myo_out (n, 2048) u32 bit-packed: seg_map[LV] and seg_map[MYO] = clip(myo_mask - lv_mask, 0, 1);
 *   area_out (n, 2) i32 = their pixel counts.  Any of the three may be NULL (not all).
 */
/*
 * Aggregation of LV + MYO sample stacks (three classes: BG = neither mask, LV, MYO).
 *   replaces  UncertaintyTask.sample_entropy for one-hot (T, 3, H, W) stacks   cu/task/uncertainty.py:107-133 (base = C = 3)
 *             the class-count / argmax bookkeeping around it                    cu/task/regression/contour_uncertainty.py:76-82
 *   lv_masks, myo_masks (n_img, n_samples, 2048) u32 bit-packed (cub200_us_myo_masks)
 *   entropy_out optional (n_img, 256, 256) fp64; count_out optional (n_img, 3, 256, 256) int32 class counts;
 *   pred_out optional (n_img, 256, 256) uint8 = argmax of the counts (first maximum wins, like numpy).
 */
int cub200_mask_class_entropy(const uint32_t* lv_masks, const uint32_t* myo_masks, int64_t n_img, int64_t n_samples,
                              double* entropy_out, int32_t* count_out, uint8_t* pred_out, void* stream);

int cub200_us_myo_masks(const float* landmarks, int64_t n, int flags, uint32_t* lv_out, uint32_t* myo_out, int32_t* area_out,
                        void* stream);

/*
 * UncertaintyTask.sample_entropy in full (cu/task/uncertainty.py:107-133): probabilities of any class count.
 *   samples (n_img, n_samples, n_classes, n_pixels) fp32, 1 <= n_classes <= 8 -> entropy_out (n_img, n_pixels) fp32:
 *   mean over the samples, then C == 1: entropy of [p, 1 - p] in bits, else entropy over the classes in base C
 *   (scipy.stats.entropy: normalised by the class sum), non-finite -> 0.  For 0/1 masks of the contour path the count-map
 *   route (cub200_finalize_maps) gives the same map in fp64.
 */
int cub200_sample_entropy(const float* samples, int64_t n_img, int n_samples, int n_classes, int64_t n_pixels,
                          float* entropy_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CUB200_H */

/* Plain-C consumer of libcub200.so: no torch, no C++ -- only include/cub200.h, cudart and dlopen-free linking.
 * Rasterises the reference's LV example contour (contour_uncertainty/data/config.py:110-130), aggregates 3 copies into a
 * difference map, finalises the maps and prints what a parity test can check. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../include/cub200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %s at %s\n", cudaGetErrorString(e_), #x); return 2; } } while (0)

static const float LV[21][2] = {{114, 205}, {105, 185}, {102, 167}, {102, 151}, {101, 134}, {98, 116}, {98, 100}, {99, 83},
                                {100, 67},  {106, 51},  {123, 35},  {139, 40},  {153, 55},  {168, 74}, {174, 92}, {179, 111},
                                {182, 128}, {184, 145}, {182, 163}, {179, 180}, {173, 200}};

int main(void) {
    const int n = 3;
    float h_c[3][21][2];
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 21; ++k) { h_c[i][k][0] = LV[k][0] + (float)i; h_c[i][k][1] = LV[k][1]; }
    float* d_c; uint32_t* d_m; int32_t *d_area, *d_diff, *d_count; double *d_mean, *d_ent, *d_lut; uint8_t* d_pred;
    CK(cudaMalloc((void**)&d_c, sizeof(h_c)));
    CK(cudaMalloc((void**)&d_m, (size_t)n * CUB200_MASK_WORDS * 4));
    CK(cudaMalloc((void**)&d_area, n * 4));
    CK(cudaMalloc((void**)&d_diff, 65536 * 4));
    CK(cudaMalloc((void**)&d_count, 65536 * 4));
    CK(cudaMalloc((void**)&d_mean, 65536 * 8));
    CK(cudaMalloc((void**)&d_ent, 65536 * 8));
    CK(cudaMalloc((void**)&d_lut, 3 * (n + 1) * 8));
    CK(cudaMalloc((void**)&d_pred, 65536));
    CK(cudaMemcpy(d_c, h_c, sizeof(h_c), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_diff, 0, 65536 * 4));
    void* d_rws;
    CK(cudaMalloc(&d_rws, (size_t)cub200_raster_workspace_bytes(n)));
    int rc = cub200_rasterize_contours(d_c, n, 1, d_m, d_area, d_diff, n, d_rws, NULL);
    if (rc) { printf("rasterize rc %d %s\n", rc, cub200_last_error()); return 3; }
    rc = cub200_finalize_maps(d_diff, 1, n, d_count, d_mean, d_ent, d_pred, d_lut, 0, NULL, NULL, NULL);
    if (rc) { printf("finalize rc %d %s\n", rc, cub200_last_error()); return 3; }
    CK(cudaDeviceSynchronize());
    int32_t h_area[3];
    static int32_t h_count[65536];
    static uint8_t h_pred[65536];
    CK(cudaMemcpy(h_area, d_area, sizeof(h_area), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_pred, d_pred, sizeof(h_pred), cudaMemcpyDeviceToHost));
    long total = 0, pred = 0;
    for (int i = 0; i < 65536; ++i) { total += h_count[i]; pred += h_pred[i]; }
    printf("abi %d sm %d area %d %d %d count_sum %ld pred_sum %ld\n", cub200_abi_version(), cub200_built_for_sm(), h_area[0], h_area[1],
           h_area[2], total, pred);
    /* mask-based LV length of the same three masks (weights evaluated in the kernel: no host table), then the largest
     * blob of the masks fed back as 0/1 probabilities (must reproduce their areas) */
    double* d_len; void* d_ws; float* d_prob; uint32_t* d_blob; int32_t* d_barea; void* d_ws2;
    CK(cudaMalloc((void**)&d_len, n * 8));
    CK(cudaMalloc(&d_ws, (size_t)cub200_simpson_workspace_bytes(n)));
    rc = cub200_mask_lv_length(d_m, n, 5e-3, NULL, 1.0, 1.0, d_len, NULL, NULL, d_ws, NULL);
    if (rc) { printf("lv_length rc %d %s\n", rc, cub200_last_error()); return 3; }
    double h_len[3];
    CK(cudaMemcpy(h_len, d_len, sizeof(h_len), cudaMemcpyDeviceToHost));
    static uint32_t h_m[3 * CUB200_MASK_WORDS];
    static float h_prob[3 * 65536];
    CK(cudaMemcpy(h_m, d_m, sizeof(h_m), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i)
        for (int p = 0; p < 65536; ++p) h_prob[i * 65536 + p] = (float)((h_m[i * CUB200_MASK_WORDS + (p >> 5)] >> (p & 31)) & 1u);
    CK(cudaMalloc((void**)&d_prob, sizeof(h_prob)));
    CK(cudaMalloc((void**)&d_blob, sizeof(h_m)));
    CK(cudaMalloc((void**)&d_barea, n * 4));
    CK(cudaMalloc(&d_ws2, (size_t)cub200_seg_blob_workspace_bytes(n)));
    CK(cudaMemcpy(d_prob, h_prob, sizeof(h_prob), cudaMemcpyHostToDevice));
    rc = cub200_seg_largest_blob(d_prob, n, d_blob, d_barea, d_ws2, NULL);
    if (rc) { printf("largest_blob rc %d %s\n", rc, cub200_last_error()); return 3; }
    int32_t h_barea[3];
    CK(cudaMemcpy(h_barea, d_barea, sizeof(h_barea), cudaMemcpyDeviceToHost));
    printf("lv_length %.12f %.12f %.12f blob_area %d %d %d\n", h_len[0], h_len[1], h_len[2], h_barea[0], h_barea[1], h_barea[2]);
    /* error behaviour: invalid arguments come back as codes, never as exceptions */
    printf("einval %d empty %d\n", cub200_rasterize_contours(NULL, 1, 1, NULL, NULL, NULL, 0, NULL, NULL),
           cub200_rasterize_contours(NULL, 0, 1, NULL, NULL, NULL, 0, NULL, NULL));
    return 0;
}

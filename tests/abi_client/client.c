/* A plain C client of libcnrma_b200.so: no Python, no torch -- cudaMalloc'd buffers, the C ABI of include/cnrma_b200.h,
 * and the C oracle as the checker (test infrastructure).  Stage A on a small scene: NCHW maps are converted with
 * cnrma_to_channels_last, aggregated with cnrma_aggregate_views, and the volume / counts must equal the oracle's bit
 * for bit.  Without a usable device it reports the library's status string and exits 0 (the CPU test only checks that
 * the client builds, links and that the library fails cleanly).  Built and run by tests/test_abi_c_client.py. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cnrma_b200.h"

void cnrma_oracle_aggregate_views(int V, int C, int H, int W, int nx, int ny, int nz, float voxel_size, const float *origin,
                                  float stride, const float *projections, const float *features, int mean, float *volume,
                                  int64_t *count);

static uint32_t rng_state = 12345u;
static float frand(void) {
    rng_state = rng_state * 1664525u + 1013904223u;
    return (float)(rng_state >> 8) / 16777216.0f - 0.5f;
}

#define CK(call)                                                                     \
    do {                                                                             \
        int s_ = (call);                                                             \
        if (s_ != CNRMA_OK) {                                                        \
            fprintf(stderr, "%s -> %s\n", #call, cnrma_status_string(s_));           \
            return 2;                                                                \
        }                                                                            \
    } while (0)

int main(void) {
    printf("abi version %d\n", cnrma_abi_version());
    const int dev_status = cnrma_check_device();
    if (dev_status != CNRMA_OK) {
        printf("no usable device: %s\n", cnrma_status_string(dev_status));
        return 0;
    }
    enum { V = 5, C = 16, H = 12, W = 16, NX = 10, NY = 9, NZ = 5 };
    const float vs = 0.4f, stride = 4.0f, origin[3] = {0.1f, -0.2f, 0.0f};
    const int nvox = NX * NY * NZ;
    /* cameras on a ring looking at the centre: K [R|t] with rows 0-1 in full-resolution pixels */
    float P[V][12];
    for (int v = 0; v < V; ++v) {
        const double ang = 6.283185307179586 * v / V, cx = 2.0 + 1.2 * cos(ang), cy = 1.8 + 1.2 * sin(ang), cz = 1.0;
        const double fx = cos(ang + 3.141592653589793), fy = sin(ang + 3.141592653589793);   /* forward */
        const double R[3][3] = {{-fy, fx, 0.0}, {0.0, 0.0, -1.0}, {fx, fy, 0.0}};           /* right, down, forward */
        const double f = 0.9 * W * 4, u0 = 0.5 * W * 4, v0 = 0.5 * H * 4;
        double Rt[3][4];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) Rt[r][c] = R[r][c];
            Rt[r][3] = -(R[r][0] * cx + R[r][1] * cy + R[r][2] * cz);
        }
        for (int c = 0; c < 4; ++c) {
            P[v][c] = (float)(f * Rt[0][c] + u0 * Rt[2][c]);
            P[v][4 + c] = (float)(f * Rt[1][c] + v0 * Rt[2][c]);
            P[v][8 + c] = (float)Rt[2][c];
        }
    }
    const size_t map = (size_t)C * H * W;
    float *h_feat = (float *)malloc(sizeof(float) * V * map);
    for (size_t i = 0; i < V * map; ++i) h_feat[i] = frand();
    float *ref_vol = (float *)malloc(sizeof(float) * C * nvox);
    int64_t *ref_cnt = (int64_t *)malloc(sizeof(int64_t) * nvox);
    cnrma_oracle_aggregate_views(V, C, H, W, NX, NY, NZ, vs, origin, stride, &P[0][0], h_feat, 1, ref_vol, ref_cnt);

    float *d_nchw, *d_cl, *d_P, *d_vol;
    int32_t *d_cnt;
    uint8_t *d_valid;
    if (cudaMalloc((void **)&d_nchw, sizeof(float) * V * map) || cudaMalloc((void **)&d_cl, sizeof(float) * V * map) ||
        cudaMalloc((void **)&d_P, sizeof(P)) || cudaMalloc((void **)&d_vol, sizeof(float) * C * nvox) ||
        cudaMalloc((void **)&d_cnt, sizeof(int32_t) * nvox) || cudaMalloc((void **)&d_valid, nvox))
        return 3;
    cudaMemcpy(d_nchw, h_feat, sizeof(float) * V * map, cudaMemcpyHostToDevice);
    cudaMemcpy(d_P, P, sizeof(P), cudaMemcpyHostToDevice);
    cudaMemset(d_vol, 0xFF, sizeof(float) * C * nvox);

    const cnrma_grid grid = {NX, NY, NZ, vs, {origin[0], origin[1], origin[2]}};
    const void *ptrs[V];
    for (int v = 0; v < V; ++v) ptrs[v] = d_nchw + v * map;
    const cnrma_features nchw = {V, C, H, W, CNRMA_F32, (int64_t)H * W, W, 1, ptrs};
    CK(cnrma_to_channels_last(&nchw, d_cl, NULL));
    for (int v = 0; v < V; ++v) ptrs[v] = d_cl + v * map;
    const cnrma_features cl = {V, C, H, W, CNRMA_F32, 1, (int64_t)W * C, C, ptrs};
    CK(cnrma_aggregate_views(&grid, &cl, d_P, 12, stride, CNRMA_AGG_MEAN, d_vol, C, 1, d_cnt, d_valid, NULL));
    if (cudaDeviceSynchronize() != cudaSuccess) return 4;

    float *vol = (float *)malloc(sizeof(float) * C * nvox);
    int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * nvox);
    cudaMemcpy(vol, d_vol, sizeof(float) * C * nvox, cudaMemcpyDeviceToHost);
    cudaMemcpy(cnt, d_cnt, sizeof(int32_t) * nvox, cudaMemcpyDeviceToHost);
    long bad = 0, seen = 0;
    for (int i = 0; i < nvox; ++i) {
        if (cnt[i] != (int32_t)ref_cnt[i]) ++bad;
        seen += cnt[i] > 0;
        for (int c = 0; c < C; ++c)                        /* library: [nvox, C]; oracle: [C, nvox] */
            if (memcmp(&vol[(size_t)i * C + c], &ref_vol[(size_t)c * nvox + i], 4) != 0) ++bad;
    }
    printf("voxels seen %ld of %d, mismatches %ld\n", seen, nvox, bad);
    return (bad == 0 && seen > nvox / 4) ? 0 : 1;
}

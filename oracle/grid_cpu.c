/* TEST / BASELINE INFRASTRUCTURE - not part of the product (ucnerf_b200/ never loads this).
 *
 * Plain-C restatement of the forward of the reference's CUDA-only hash-grid kernel,
 *   kernel_grid            nerf/gridencoder/src/gridencoder.cu:L87-197   (no dy_dx branch)
 *   fast_hash              gridencoder.cu:L50-64
 *   get_grid_index         gridencoder.cu:L66-84
 *   launch / entry         gridencoder.cu:L373-383, L448-471
 * so that the reference's own Python (internal/models.py + gridencoder/grid.py, which has no CPU kernel) can be timed on
 * the host cores: oracle/ref_shim.py binds it as the `_gridencoder` backend of the CPU reference arm (bench.py
 * --impl reference / cpu_baseline).  One OpenMP thread team over the batch, fp32 embeddings.
 *
 * Arithmetic follows the CUDA code as nvcc compiles it: `inputs*scale + 0.5` and `w*emb + acc` are fused (fmaf), every
 * other operation is a separately rounded fp32 operation (build with -ffp-contract=off).  Pinned bit-for-bit against
 * the numpy restatement oracle/ucnerf_oracle.py::grid_encode_forward (tests/test_oracle_grid_c.py), which in turn is
 * pinned against the reference kernel's own outputs (tests/golden/gridref_*.npz, tests/test_gpu_grid.py).
 *
 * build:  gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC oracle/grid_cpu.c -o oracle/grid_cpu.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAXD 5

static const uint32_t PRIMES[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};

static inline uint32_t grid_index(const uint32_t* pg, int D, uint32_t hashmap_size, uint32_t resolution, int gridtype,
                                  int align_corners) {
    uint32_t stride = 1, index = 0;
    int d = 0;
    for (; d < D && stride <= hashmap_size; d++) {          /* L72: early-exit loop */
        index += pg[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) {           /* L79-81 */
        index = 0;
        for (int i = 0; i < D; i++) index ^= pg[i] * PRIMES[i];
    }
    return index % hashmap_size;
}

/* outputs [L,B,C] fp32 (caller-allocated); returns 0, or -1 for unsupported D */
int ucnerf_oracle_grid_forward_f32(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs,
                                   uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int gridtype,
                                   int align_corners, int interp) {
    if (D < 1 || D > MAXD) return -1;
    for (uint32_t level = 0; level < L; level++) {
        const float* grid = embeddings + (size_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        const float scale = exp2f((float)level * S) * (float)H - 1.0f;       /* L138 */
        const uint32_t resolution = (uint32_t)ceilf(scale) + 1;               /* L139 */
        float* out_l = outputs + (size_t)level * B * C;
#pragma omp parallel for schedule(static)
        for (int64_t b = 0; b < (int64_t)B; b++) {
            const float* x = inputs + (size_t)b * D;
            float* out = out_l + (size_t)b * C;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++)
                if (x[d] < 0.0f || x[d] > 1.0f) oob = 1;                       /* L110-135 */
            if (oob) {
                memset(out, 0, sizeof(float) * C);
                continue;
            }
            float pos[MAXD];
            uint32_t pos_grid[MAXD];
            for (uint32_t d = 0; d < D; d++) {                                /* L146-159 */
                float p = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
                float fl = floorf(p);
                pos_grid[d] = (uint32_t)fl;
                p = p - fl;
                if (interp == 1) p = p * p * (3.0f - 2.0f * p);
                pos[d] = p;
            }
            float acc[8];
            float* res = acc;
            float big[64];
            if (C > 8) res = big;
            for (uint32_t c = 0; c < C; c++) res[c] = 0.0f;
            for (uint32_t idx = 0; idx < (1u << D); idx++) {                   /* L166-191 */
                float w = 1.0f;
                uint32_t pgl[MAXD];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) {
                        w = w * (1.0f - pos[d]);
                        pgl[d] = pos_grid[d];
                    } else {
                        w = w * pos[d];
                        pgl[d] = pos_grid[d] + 1;
                    }
                }
                const float* e = grid + (size_t)grid_index(pgl, (int)D, hashmap_size, resolution, gridtype, align_corners) * C;
                for (uint32_t c = 0; c < C; c++) res[c] = fmaf(w, e[c], res[c]);
            }
            for (uint32_t c = 0; c < C; c++) out[c] = res[c];
        }
    }
    return 0;
}

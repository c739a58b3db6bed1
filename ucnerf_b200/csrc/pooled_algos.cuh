// Pooled hash-grid features of one sampled interval, forward and backward - the front end of MLP.predict_density
// written once as host+device templates (CUDA instantiation: pooled_encode.cu; serial CPU instantiation for the
// GPU-less test-suite: tests/cpu_harness.cpp).
//
// Reference (under /root/reference/nerf/): internal/models.py:L485-496 (contract, / bound, encoder, erf down-weighting,
// mean over the multisample points), internal/coord.py:L60-72,L75-116 (contract_mean_std under no_grad),
// gridencoder/grid.py:L158-174 ((x + 1) / 2), gridencoder/src/gridencoder.cu:L87-197 (kernel_grid),
// L248-340 (kernel_grid_backward).  For one interval b with M multisample points (mean_j, std_j):
//
//   z_j, s_j  = contract(mean_j, std_j);  x_j = z_j / 2;  sigma_j = s_j / 2;  g_j = (x_j + 1) / 2
//   om_jl     = erf(1 / sqrt(8 sigma_j^2 G_l^2))                       G_l = grid_sizes[l]
//   F[b,l,:]  = 1/M sum_j om_jl * trilerp_l(g_j)[:]                     (zero contribution if g_j is outside [0,1]^3)
//   coord[b]  = 1/M sum_j x_j                                           (models.py:L512)
//
// backward (the means / stds carry no gradient: track_linearize is @torch.no_grad, coord.py:L75):
//   dE[idx_l(corner k of g_j)] += w_k(g_j) * om_jl / M * dF[b,l,:]
#pragma once
#include "ray_algos.cuh"

namespace ucnerf {

// coord.py:L60-72 + models.py:L489-493 + grid.py:L162 for one point: unit-cube coordinate g, contracted std / 2 and
// the contracted mean / 2 (xh).  Plain fp32 IEEE operations in the reference's order.
UC_HD void pooled_point(const float* mean, float std, bool contract, float (&g)[3], float& sigma, float (&xh)[3]) {
    float x[3] = {mean[0], mean[1], mean[2]};
    float sd = std;
    if (contract) {
        const float m2 = fmaxf(fa(fa(fm(x[0], x[0]), fm(x[1], x[1])), fm(x[2], x[2])), kEps);
        if (!(m2 <= 1.f)) {
            const float mag = fsqrt(m2);
            const float k = fd(fs(fm(2.f, mag), 1.f), m2);
#pragma unroll
            for (int i = 0; i < 3; ++i) x[i] = fm(k, x[i]);
            const float c = fd(cbrtf(fs(fm(2.f, mag), 1.f)), mag);   // torch.pow(., 1/3): within 1 ulp of cbrtf
            sd = fm(fm(c, c), sd);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = fm(x[i], 0.5f);           // means / bound, bound = 2
        sd = fm(sd, 0.5f);
    }
    sigma = sd;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        xh[i] = x[i];
        g[i] = fm(fa(x[i], 1.f), 0.5f);
    }
}

UC_HD bool in_unit_cube(const float (&g)[3]) {   // gridencoder.cu:L110-135: out-of-range input -> zeros
    return !(g[0] < 0.f || g[0] > 1.f || g[1] < 0.f || g[1] > 1.f || g[2] < 0.f || g[2] > 1.f);
}

// models.py:L495 erf(1 / sqrt(8 * std^2 * grid_sizes^2)); g2 = float(grid_sizes[l]^2)
UC_HD float pooled_erf_weight(float sigma, float g2) {
    return erff(fd(1.f, fsqrt(fm(fm(8.f, fm(sigma, sigma)), g2))));
}

// trilinear corner weight, corner k: bit0 -> x, bit1 -> y, bit2 -> z (gridencoder.cu:L166-185)
UC_HD float corner_weight(const CellCoords& c, int k) {
    return ((k & 1) ? c.fx : 1.f - c.fx) * ((k & 2) ? c.fy : 1.f - c.fy) * ((k & 4) ? c.fz : 1.f - c.fz);
}

// F[4] of one (interval, level).  Load: float4 operator()(size_t entry) - read-only gather of a table entry.
template <class Load>
UC_HD void pooled_level_forward(const GridLevel& lv, float g2, const float* means, const float* stds, int M,
                                bool contract, const Load& load, float (&F)[4]) {
    F[0] = F[1] = F[2] = F[3] = 0.f;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        float g[3], sigma, xh[3];
        pooled_point(means + 3 * j, stds[j], contract, g, sigma, xh);
        if (!in_unit_cube(g)) continue;
        const CellCoords c = cell_of(lv, g);
        float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t idx = level_index(lv, c.ix + (k & 1), c.iy + ((k >> 1) & 1), c.iz + ((k >> 2) & 1));
            const float w = corner_weight(c, k);
            const float4 v = load((size_t)lv.offset + idx);
            r[0] = fmaf(w, v.x, r[0]); r[1] = fmaf(w, v.y, r[1]); r[2] = fmaf(w, v.z, r[2]); r[3] = fmaf(w, v.w, r[3]);
        }
        const float om = pooled_erf_weight(sigma, g2);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) F[ch] = fmaf(om, r[ch], F[ch]);
    }
    const float inv = 1.f / (float)M;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) F[ch] *= inv;
}

// scatter of one (interval, level).  Add: void operator()(size_t entry, float a, float b, float c, float d).
template <class Add>
UC_HD void pooled_level_backward(const GridLevel& lv, float g2, const float* means, const float* stds, int M,
                                 bool contract, const float (&dF)[4], const Add& add) {
    const float inv = 1.f / (float)M;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        float g[3], sigma, xh[3];
        pooled_point(means + 3 * j, stds[j], contract, g, sigma, xh);
        if (!in_unit_cube(g)) continue;       // gridencoder.cu:L276-281
        const CellCoords c = cell_of(lv, g);
        const float coef = pooled_erf_weight(sigma, g2) * inv;
        const float d0 = coef * dF[0], d1 = coef * dF[1], d2 = coef * dF[2], d3 = coef * dF[3];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t idx = level_index(lv, c.ix + (k & 1), c.iy + ((k >> 1) & 1), c.iz + ((k >> 2) & 1));
            const float w = corner_weight(c, k);
            add((size_t)lv.offset + idx, w * d0, w * d1, w * d2, w * d3);
        }
    }
}

// Same scatter with the corner coefficients of CONSECUTIVE points that share a grid cell summed before they are
// reduced: the M points of an interval are ordered along the ray and mostly stay in one cell on the coarse levels
// (93 / 86 / 75 / 63 / 48 / 30 % of the proposal intervals on levels 0..5), so a run costs 8 reductions instead of
// 8 per point.  dE[corner k of the run's cell] += (sum_j w_k(g_j) om_jl / M) * dF: the same value up to the order of
// the fp32 sums.  The atomics are the limiter of the backward, not the arithmetic.
template <class Add>
UC_HD void pooled_level_backward_runs(const GridLevel& lv, float g2, const float* means, const float* stds, int M,
                                      bool contract, const float (&dF)[4], const Add& add) {
    const float inv = 1.f / (float)M;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t cx = 0, cy = 0, cz = 0;
    bool open = false;
#pragma unroll
    for (int j = 0; j <= M; ++j) {
        bool inside = false;
        CellCoords c{};
        float coef = 0.f;
        if (j < M) {
            float g[3], sigma, xh[3];
            pooled_point(means + 3 * j, stds[j], contract, g, sigma, xh);
            inside = in_unit_cube(g);          // gridencoder.cu:L276-281
            if (inside) {
                c = cell_of(lv, g);
                coef = pooled_erf_weight(sigma, g2) * inv;
            }
        }
        // close the current run when the cell changes or after the last point
        if (open && (j == M || (inside && (c.ix != cx || c.iy != cy || c.iz != cz)))) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t idx = level_index(lv, cx + (k & 1), cy + ((k >> 1) & 1), cz + ((k >> 2) & 1));
                add((size_t)lv.offset + idx, a[k] * dF[0], a[k] * dF[1], a[k] * dF[2], a[k] * dF[3]);
                a[k] = 0.f;
            }
            open = false;
        }
        if (inside) {
            cx = c.ix; cy = c.iy; cz = c.iz;
            open = true;
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fmaf(corner_weight(c, k), coef, a[k]);
        }
    }
}

// Run merging across K CONSECUTIVE intervals (neighbouring samples of one ray, K * M points ordered along the ray): on
// the coarse levels a thread then issues one set of 8 reductions for the whole stretch instead of one per interval.
// The intervals have different dF, so the 8 x 4 products are accumulated (not the 8 weights).  grad_features points at
// the first interval's row; rows are `row_stride` floats apart; `level_off` = 4 * level.
template <class Add>
UC_HD void pooled_level_backward_ray_runs(const GridLevel& lv, float g2, const float* means, const float* stds, int M, int K,
                                          bool contract, const float* grad_features, int row_stride, int level_off,
                                          const Add& add) {
    const float inv = 1.f / (float)M;
    float a[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k][0] = a[k][1] = a[k][2] = a[k][3] = 0.f;
    uint32_t cx = 0, cy = 0, cz = 0;
    bool open = false;
    for (int i = 0; i <= K; ++i) {
        float dF[4] = {0.f, 0.f, 0.f, 0.f};
        if (i < K) {
            const float* gp = grad_features + (size_t)i * row_stride + level_off;
            dF[0] = gp[0]; dF[1] = gp[1]; dF[2] = gp[2]; dF[3] = gp[3];
        }
        const int npts = i < K ? M : 1;            // one extra pass closes the last run
        for (int j = 0; j < npts; ++j) {
            bool inside = false;
            CellCoords c{};
            float coef = 0.f;
            if (i < K) {
                float g[3], sigma, xh[3];
                pooled_point(means + 3 * ((size_t)i * M + j), stds[(size_t)i * M + j], contract, g, sigma, xh);
                inside = in_unit_cube(g);
                if (inside) {
                    c = cell_of(lv, g);
                    coef = pooled_erf_weight(sigma, g2) * inv;
                }
            }
            if (open && (i == K || (inside && (c.ix != cx || c.iy != cy || c.iz != cz)))) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t idx = level_index(lv, cx + (k & 1), cy + ((k >> 1) & 1), cz + ((k >> 2) & 1));
                    add((size_t)lv.offset + idx, a[k][0], a[k][1], a[k][2], a[k][3]);
                    a[k][0] = a[k][1] = a[k][2] = a[k][3] = 0.f;
                }
                open = false;
            }
            if (inside) {
                cx = c.ix; cy = c.iy; cz = c.iz;
                open = true;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float w = corner_weight(c, k) * coef;
                    a[k][0] = fmaf(w, dF[0], a[k][0]); a[k][1] = fmaf(w, dF[1], a[k][1]);
                    a[k][2] = fmaf(w, dF[2], a[k][2]); a[k][3] = fmaf(w, dF[3], a[k][3]);
                }
            }
        }
    }
}

// models.py:L512 means.mean(dim=-2) of the contracted means / 2
UC_HD void pooled_coord(const float* means, const float* stds, int M, bool contract, float (&out)[3]) {
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < M; ++j) {
        float g[3], sigma, xh[3];
        pooled_point(means + 3 * j, stds[j], contract, g, sigma, xh);
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[i] = j == 0 ? xh[i] : fa(acc[i], xh[i]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = fd(acc[i], (float)M);
}

}  // namespace ucnerf

// Fused optimiser step for a hash-grid table ("hash-decay + Adam"): one pass over the
// embeddings does what the reference spreads over the hash-decay loss term and its autograd
// (internal/models.py:L297-306 segment_coo(param ** 2, idx, reduce='mean').mean(), train_utils.py:L301-305
// hash_decay_mults), `param.grad.nan_to_num_()` (train_utils.py:L344-345), torch.optim.Adam
// (train_utils.py:L347-366: betas, eps, no weight decay / amsgrad) and optimizer.zero_grad() (train.py:L164):
//
//   g    = nan_to_num(grad + mult * 2 p / (T_level * L * C))        d/dp of mult * mean_{l,c} mean_{e in l} p[e,c]^2
//   m    = m + (g - m) (1 - beta1)                                   exp_avg.lerp_(g, 1 - beta1)
//   v    = v beta2 + (1 - beta2) g g                                 exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
//   p   -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)              bc_i = 1 - beta_i ^ step
//   grad = 0                                                         (optional)
//
// HBM-bound streaming kernel: 16 B read (p, g, m, v) + 16 B written (p, m, v, g) per float, 128-bit accesses,
// grid = multiple of 148 SMs.  The reference reads / writes the tables > 10 times per step for the same result.
#include "../../include/ucnerf_b200.h"
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace ucnerf {

struct AdamLevels {
    uint32_t n_levels;
    uint32_t end[UCNERF_MAX_GRID_LEVELS];     // entry offset where each level ends
    float coef[UCNERF_MAX_GRID_LEVELS];       // mult * 2 / (T_level * L * C)
};

__device__ __forceinline__ float nan_to_num_f(float x) {   // torch.nan_to_num defaults: nan -> 0, +-inf -> +-FLT_MAX
    if (x != x) return 0.f;
    return fminf(fmaxf(x, -3.402823466e+38f), 3.402823466e+38f);
}

// train_utils.clip_gradients (L335-345) on one gradient value: global-norm clip coefficient (computed by the caller over
// ALL parameters, see table_stats_kernel), then value clipping to +-max_val (max_val <= 0: off), then nan_to_num_()
__device__ __forceinline__ float clip_grad_f(float g, float scale, float max_val) {
    g *= scale;
    if (max_val > 0.f) g = g != g ? g : fminf(fmaxf(g, -max_val), max_val);   // clamp_ keeps NaN, nan_to_num_ then zeroes it
    return nan_to_num_f(g);
}

// Per-level sums of a table in ONE read pass: out[2 l] += sum p^2 (the hash-decay loss, models.py:L297-306) and
// out[2 l + 1] += sum (g + c_l p)^2 (this table's share of the global gradient norm that clip_grad_norm_ takes over all
// parameters, train_utils.py:L336-337, with the hash-decay gradient the fused optimiser folds in).  fp64 accumulation.
__global__ void __launch_bounds__(256)
table_stats_kernel(const float4* __restrict__ p, const float4* __restrict__ g, uint32_t entries,
                   const __grid_constant__ AdamLevels lv, double* __restrict__ out) {
    __shared__ double sh[2 * UCNERF_MAX_GRID_LEVELS];
    for (int i = threadIdx.x; i < 2 * UCNERF_MAX_GRID_LEVELS; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const uint32_t per = div_up(entries, gridDim.x);
    const uint32_t e0 = blockIdx.x * per, e1 = min(e0 + per, entries);   // contiguous slice per CTA: at most 2-3 levels
    uint32_t l = 0;
    double sp = 0.0, sg = 0.0;
    for (uint32_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        uint32_t le = l;
        while (le + 1 < lv.n_levels && e >= lv.end[le]) ++le;
        if (le != l) {   // this thread moved to the next level: bank what it has
            atomicAdd(&sh[2 * l], sp);
            atomicAdd(&sh[2 * l + 1], sg);
            sp = sg = 0.0;
            l = le;
        }
        const float c = lv.coef[l];
        const float4 pp = p[e];
        sp += (double)pp.x * pp.x + (double)pp.y * pp.y + (double)pp.z * pp.z + (double)pp.w * pp.w;
        if (g) {
            const float4 gg = g[e];
            const float a = fmaf(c, pp.x, gg.x), b = fmaf(c, pp.y, gg.y), cc = fmaf(c, pp.z, gg.z), d = fmaf(c, pp.w, gg.w);
            sg += (double)a * a + (double)b * b + (double)cc * cc + (double)d * d;
        }
    }
    atomicAdd(&sh[2 * l], sp);
    atomicAdd(&sh[2 * l + 1], sg);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * (int)lv.n_levels; i += blockDim.x)
        if (sh[i] != 0.0) atomicAdd(&out[i], sh[i]);
}

// one thread = one table entry of C = 4 floats (float4); entries = sum T
__global__ void __launch_bounds__(256)
grid_adam_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                 uint32_t entries, const __grid_constant__ AdamLevels lv, float one_minus_b1, float b2, float one_minus_b2,
                 float step_size, float inv_sqrt_bc2, float eps, int zero_grad, float grad_scale, float max_val) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < entries; e += gridDim.x * blockDim.x) {
        uint32_t l = 0;
        while (l + 1 < lv.n_levels && e >= lv.end[l]) ++l;
        const float c = lv.coef[l];
        const float4 pp = p[e], gg = g[e], mm = m[e], vv = v[e];
        float4 po, mo, vo;
        {
            const float gr = clip_grad_f(fmaf(c, pp.x, gg.x), grad_scale, max_val);
            const float mn = fmaf(gr - mm.x, one_minus_b1, mm.x);
            const float vn = fmaf(one_minus_b2 * gr, gr, vv.x * b2);
            const float den = sqrtf(vn) * inv_sqrt_bc2 + eps;
            po.x = pp.x - step_size * (mn / den);
            mo.x = mn; vo.x = vn;
        }
        {
            const float gr = clip_grad_f(fmaf(c, pp.y, gg.y), grad_scale, max_val);
            const float mn = fmaf(gr - mm.y, one_minus_b1, mm.y);
            const float vn = fmaf(one_minus_b2 * gr, gr, vv.y * b2);
            const float den = sqrtf(vn) * inv_sqrt_bc2 + eps;
            po.y = pp.y - step_size * (mn / den);
            mo.y = mn; vo.y = vn;
        }
        {
            const float gr = clip_grad_f(fmaf(c, pp.z, gg.z), grad_scale, max_val);
            const float mn = fmaf(gr - mm.z, one_minus_b1, mm.z);
            const float vn = fmaf(one_minus_b2 * gr, gr, vv.z * b2);
            const float den = sqrtf(vn) * inv_sqrt_bc2 + eps;
            po.z = pp.z - step_size * (mn / den);
            mo.z = mn; vo.z = vn;
        }
        {
            const float gr = clip_grad_f(fmaf(c, pp.w, gg.w), grad_scale, max_val);
            const float mn = fmaf(gr - mm.w, one_minus_b1, mm.w);
            const float vn = fmaf(one_minus_b2 * gr, gr, vv.w * b2);
            const float den = sqrtf(vn) * inv_sqrt_bc2 + eps;
            po.w = pp.w - step_size * (mn / den);
            mo.w = mn; vo.w = vn;
        }
        p[e] = po; m[e] = mo; v[e] = vo;
        if (zero_grad) g[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

}  // namespace ucnerf

using namespace ucnerf;

static int make_levels(AdamLevels& lv, const int32_t* offsets_host, uint32_t L, uint32_t C, double hash_decay_mult) {
    lv.n_levels = L;
    for (uint32_t l = 0; l < L; ++l) {
        lv.end[l] = (uint32_t)offsets_host[l + 1];
        const double T = (double)(offsets_host[l + 1] - offsets_host[l]);
        lv.coef[l] = (float)(hash_decay_mult * 2.0 / (T * (double)L * (double)C));
    }
    return 0;
}

extern "C" int ucnerf_grid_adam_step_clipped(float* embeddings, float* grad, float* exp_avg, float* exp_avg_sq,
                                             const int32_t* offsets_host, uint32_t L, uint32_t C, double lr, double beta1,
                                             double beta2, double eps, uint64_t step, double hash_decay_mult, int zero_grad,
                                             double grad_scale, double grad_max_val, void* stream) {
    UC_REQUIRE(embeddings && grad && exp_avg && exp_avg_sq && offsets_host, "grid_adam_step: null argument");
    UC_REQUIRE(C == 4, "grid_adam_step: level_dim must be 4");
    UC_REQUIRE(L >= 1 && L <= UCNERF_MAX_GRID_LEVELS, "grid_adam_step: levels must be in [1,16]");
    UC_REQUIRE(step >= 1, "grid_adam_step: step counts from 1 (torch.optim.Adam)");
    const uint32_t entries = (uint32_t)offsets_host[L];
    if (entries == 0) return 0;
    AdamLevels lv{};
    make_levels(lv, offsets_host, L, C, hash_decay_mult);
    const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
    const float step_size = (float)(lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
    const uint32_t blocks = std::min<uint32_t>(div_up(entries, 256u), (uint32_t)kNumSMs * 16u);
    grid_adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(embeddings), reinterpret_cast<float4*>(grad), reinterpret_cast<float4*>(exp_avg),
        reinterpret_cast<float4*>(exp_avg_sq), entries, lv, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
        step_size, inv_sqrt_bc2, (float)eps, zero_grad, (float)grad_scale, (float)grad_max_val);
    UC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ucnerf_grid_adam_step(float* embeddings, float* grad, float* exp_avg, float* exp_avg_sq,
                                     const int32_t* offsets_host, uint32_t L, uint32_t C, double lr, double beta1,
                                     double beta2, double eps, uint64_t step, double hash_decay_mult, int zero_grad,
                                     void* stream) {
    return ucnerf_grid_adam_step_clipped(embeddings, grad, exp_avg, exp_avg_sq, offsets_host, L, C, lr, beta1, beta2, eps,
                                         step, hash_decay_mult, zero_grad, 1.0, 0.0, stream);
}

extern "C" int ucnerf_grid_table_stats(const float* embeddings, const float* grad, const int32_t* offsets_host, uint32_t L,
                                       uint32_t C, double hash_decay_mult, double* out_sums, void* stream) {
    UC_REQUIRE(embeddings && offsets_host && out_sums, "grid_table_stats: null argument");
    UC_REQUIRE(C == 4, "grid_table_stats: level_dim must be 4");
    UC_REQUIRE(L >= 1 && L <= UCNERF_MAX_GRID_LEVELS, "grid_table_stats: levels must be in [1,16]");
    const uint32_t entries = (uint32_t)offsets_host[L];
    UC_CUDA_OK(cudaMemsetAsync(out_sums, 0, sizeof(double) * 2 * L, (cudaStream_t)stream));
    if (entries == 0) return 0;
    AdamLevels lv{};
    make_levels(lv, offsets_host, L, C, hash_decay_mult);
    const uint32_t blocks = std::min<uint32_t>(div_up(entries, 256u), (uint32_t)kNumSMs * 8u);
    table_stats_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(embeddings),
                                                                 reinterpret_cast<const float4*>(grad), entries, lv, out_sums);
    UC_LAUNCH_CHECK();
    return 0;
}

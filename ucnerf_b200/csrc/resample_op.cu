// Stand-alone resampling op: one pass of Model.forward's level loop between the previous level's (sdist, weights) and
// the new interval fenceposts - max-dilation, slice, annealed logits, softmax, CDF, inverse CDF, midpoints
// (internal/models.py:L156-205; stepfun.py max_dilate_weights L75-105, sample_intervals L251-294, sample L175-218,
// invert_cdf L154-161; math.py sorted_interp L88-107) - for callers that keep the rest of the level in PyTorch, i.e.
// the reference's TRAINING step (`rand=True`: the u grid is jittered per ray or per sample, stepfun.py:L206-212; the
// result is detached by `stop_level_grad`, models.py:L203-204, so no backward exists).  The reference materialises
// O(S^2) temporaries per ray here ([N,385,128] fp32 = 2.96 GB at N = 15,000 for the dilation, four [N,383,32] masks
// for the inverse CDF); this kernel is the hot path's warp-per-ray algorithm (ray_algos.cuh::resample_ray: ranked
// 3-way merge, range-max over contiguous bins, fp64 warp scan, galloping searches) with the jitter added to u.
// Own kernel instantiation: the eval path's resample_kernel (ray_march.cu) is left untouched.
#include "../../include/ucnerf_b200.h"
#include "ray_march.cuh"

namespace ucnerf {

constexpr int kResampleOpWarps = 4;

struct ResampleOpParams {
    uint32_t n_rays;
    int n_prev;
    const float* t_prev;   // [N, n_prev+1] or NULL (first level: [0, 1])
    const float* w_prev;   // [N, n_prev]   or NULL (first level: [1])
    int dilate;
    float dilation, anneal, padding;
    int S;
    const float* u;        // [S] base grid
    const float* jitter;   // [N, jitter_cols] or NULL, already scaled by max_jitter
    int jitter_cols;       // 1 (single_jitter) or S
    float* out_sdist;      // [N, S+1]
};

__global__ void __launch_bounds__(32 * kResampleOpWarps)
resample_op_kernel(const ResampleOpParams p) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5;
    const uint32_t ray = blockIdx.x * kResampleOpWarps + warp;
    if (ray >= p.n_rays) return;
    WarpExec ex{(int)(threadIdx.x & 31)};
    ResampleScratch sc;
    sc.carve(smem + (size_t)warp * ResampleScratch::floats(p.n_prev, p.S), p.n_prev, p.S);
    const float* tprev = p.t_prev ? p.t_prev + (size_t)ray * (p.n_prev + 1) : nullptr;
    const float* wprev = p.w_prev ? p.w_prev + (size_t)ray * p.n_prev : nullptr;
    const float* jit = p.jitter ? p.jitter + (size_t)ray * p.jitter_cols : nullptr;
    resample_ray(ex, p.n_prev, tprev, wprev, p.dilate != 0, p.dilation, p.anneal, p.padding, p.S, p.u, sc,
                 p.out_sdist + (size_t)ray * (p.S + 1), jit, p.jitter_cols > 1 ? 1 : 0);
}

}  // namespace ucnerf

using namespace ucnerf;

extern "C" int ucnerf_resample_intervals(const float* t_prev, const float* w_prev, uint32_t n_rays, int32_t n_prev,
                                         int dilate, float dilation, float anneal, float padding, int32_t S,
                                         const float* u, const float* jitter, int32_t jitter_cols, float* out_sdist,
                                         void* stream) {
    if (n_rays == 0) return 0;
    UC_REQUIRE(S >= 2 && n_prev >= 1 && n_prev < 65536, "resample_intervals: need S >= 2 and 1 <= n_prev < 65536");
    UC_REQUIRE(u && out_sdist, "resample_intervals: null u / out_sdist");
    UC_REQUIRE((t_prev == nullptr) == (w_prev == nullptr), "resample_intervals: t_prev and w_prev go together");
    UC_REQUIRE(t_prev || n_prev == 1, "resample_intervals: the first level has one bin");
    UC_REQUIRE(!jitter || jitter_cols == 1 || jitter_cols == S, "resample_intervals: jitter_cols must be 1 or S");
    ResampleOpParams p{};
    p.n_rays = n_rays; p.n_prev = n_prev; p.t_prev = t_prev; p.w_prev = w_prev; p.dilate = dilate ? 1 : 0;
    p.dilation = dilation; p.anneal = anneal; p.padding = padding; p.S = S; p.u = u; p.jitter = jitter;
    p.jitter_cols = jitter ? jitter_cols : 0; p.out_sdist = out_sdist;
    const size_t smem = kResampleOpWarps * ResampleScratch::floats(n_prev, S) * sizeof(float);
    UC_REQUIRE(smem <= 227 * 1024, "resample_intervals: too many bins / samples per ray for shared memory");
    UC_ENSURE_SMEM(smem, resample_op_kernel);
    resample_op_kernel<<<div_up(n_rays, (uint32_t)kResampleOpWarps), 32 * kResampleOpWarps, smem, (cudaStream_t)stream>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

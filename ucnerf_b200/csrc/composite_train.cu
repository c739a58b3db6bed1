// Differentiable alpha compositing for the training step (SURVEY.md section 8a rows R7 + R8 under autograd, config 5):
// render.compute_alpha_weights + the acc / rgb part of render.volumetric_rendering (render.py:L155-216) as one forward
// and one backward kernel.  The reference spends ~12 elementwise / cumsum / reduction launches per level here and the
// same again (plus saved [N,S] tensors) in the backward.  One thread per ray: a train batch is 15,000 rays x <= 128
// samples (7.7 MB of densities), so the kernels are launch- and latency-bound, not bandwidth-bound; the per-ray loops are
// the host+device functions of train_algos.cuh, i.e. exactly the code the CPU test-suite runs.
#include "../../include/ucnerf_b200.h"
#include "train_algos.cuh"

namespace ucnerf {

constexpr int kCompositeTrainThreads = 128;

__global__ void __launch_bounds__(kCompositeTrainThreads)
composite_train_forward_kernel(const float* __restrict__ tdist, const float* __restrict__ density,
                               const float* __restrict__ rgbs, const float* __restrict__ dirs, uint32_t N, int S, float bg,
                               float* __restrict__ weights, float* __restrict__ rgb, float* __restrict__ acc) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const float dn = norm3(dirs[3 * (size_t)r], dirs[3 * (size_t)r + 1], dirs[3 * (size_t)r + 2]);   // render.py:L158
    float c[3], a;
    composite_train_forward_ray(S, tdist + (size_t)r * (S + 1), density + (size_t)r * S,
                                rgbs ? rgbs + (size_t)r * S * 3 : nullptr, dn, bg, weights + (size_t)r * S, c, a);
    rgb[3 * (size_t)r] = c[0]; rgb[3 * (size_t)r + 1] = c[1]; rgb[3 * (size_t)r + 2] = c[2];
    acc[r] = a;
}

__global__ void __launch_bounds__(kCompositeTrainThreads)
composite_train_backward_kernel(const float* __restrict__ tdist, const float* __restrict__ density,
                                const float* __restrict__ rgbs, const float* __restrict__ dirs,
                                const float* __restrict__ weights, const float* __restrict__ acc,
                                const float* __restrict__ g_weights, const float* __restrict__ g_rgb,
                                const float* __restrict__ g_acc, uint32_t N, int S, float bg,
                                float* __restrict__ d_density, float* __restrict__ d_rgbs) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const float dn = norm3(dirs[3 * (size_t)r], dirs[3 * (size_t)r + 1], dirs[3 * (size_t)r + 2]);
    composite_train_backward_ray(S, tdist + (size_t)r * (S + 1), density + (size_t)r * S,
                                 rgbs ? rgbs + (size_t)r * S * 3 : nullptr, dn, bg, weights + (size_t)r * S, acc[r],
                                 g_weights ? g_weights + (size_t)r * S : nullptr, g_rgb ? g_rgb + 3 * (size_t)r : nullptr,
                                 g_acc ? g_acc + r : nullptr, d_density + (size_t)r * S,
                                 (rgbs && d_rgbs) ? d_rgbs + (size_t)r * S * 3 : nullptr);
}

}  // namespace ucnerf

using namespace ucnerf;

extern "C" int ucnerf_composite_train_forward(const float* tdist, const float* density, const float* rgbs, const float* dirs,
                                              uint32_t N, int32_t S, float bg, float* weights, float* rgb, float* acc,
                                              void* stream) {
    if (N == 0) return 0;
    UC_REQUIRE(S >= 1, "composite_train: S must be >= 1");
    UC_REQUIRE(tdist && density && dirs && weights && rgb && acc, "composite_train_forward: null pointer");
    composite_train_forward_kernel<<<div_up(N, (uint32_t)kCompositeTrainThreads), kCompositeTrainThreads, 0,
                                     (cudaStream_t)stream>>>(tdist, density, rgbs, dirs, N, S, bg, weights, rgb, acc);
    UC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ucnerf_composite_train_backward(const float* tdist, const float* density, const float* rgbs, const float* dirs,
                                               const float* weights, const float* acc, const float* g_weights,
                                               const float* g_rgb, const float* g_acc, uint32_t N, int32_t S, float bg,
                                               float* d_density, float* d_rgbs, void* stream) {
    if (N == 0) return 0;
    UC_REQUIRE(S >= 1, "composite_train: S must be >= 1");
    UC_REQUIRE(tdist && density && dirs && weights && acc && d_density, "composite_train_backward: null pointer");
    composite_train_backward_kernel<<<div_up(N, (uint32_t)kCompositeTrainThreads), kCompositeTrainThreads, 0,
                                      (cudaStream_t)stream>>>(tdist, density, rgbs, dirs, weights, acc, g_weights, g_rgb,
                                                              g_acc, N, S, bg, d_density, d_rgbs);
    UC_LAUNCH_CHECK();
    return 0;
}

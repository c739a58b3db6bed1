// Stand-alone render.cast_rays (internal/render.py:L94-152) for callers that keep the MLPs in PyTorch (the training
// step): tdist [N,S+1] + rays -> means [N,S,6,3], stds [N,S,6], ts [N,S,6] in ONE kernel, deterministic pattern or the
// random rotation / flip of rand=True (the uniform draws are inputs, so a seeded torch generator reproduces the
// reference's stream).  The reference spends ~30 elementwise ATen launches on [N,S,6,.] tensors here.  Thread = one
// (ray, interval); the per-interval code is train_algos.cuh::cast_interval (host+device, run serially by the CPU tests).
// No gradient: tdist is detached (models.py:L203-204) and the rays are data.
#include "../../include/ucnerf_b200.h"
#include "train_algos.cuh"

namespace ucnerf {

struct CastParams {
    uint32_t n_rays;
    int S;
    const float *tdist, *origins, *directions, *cam_dirs, *radii, *rand_vec;
    const float *rot01, *flip01;   // [N,S] each, NULL -> deterministic pattern
    float std_scale;
    ConeTable cone;
    float *means, *stds, *ts;
};

__global__ void __launch_bounds__(128)
cast_rays_kernel(const __grid_constant__ CastParams p) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (size_t)p.n_rays * p.S) return;
    const uint32_t ray = (uint32_t)(q / p.S);
    const int s = (int)(q - (size_t)ray * p.S);
    RayGeom rg;
    make_ray_geom(rg, p.origins + 3 * (size_t)ray, p.directions + 3 * (size_t)ray, p.cam_dirs + 3 * (size_t)ray,
                  p.rand_vec + 3 * (size_t)ray, p.radii[ray], 0.f, 1.f);
    const float t0 = p.tdist[(size_t)ray * (p.S + 1) + s], t1 = p.tdist[(size_t)ray * (p.S + 1) + s + 1];
    const bool rand = p.rot01 != nullptr;
    float m[18], sd[6], ts[6];
    cast_interval(rg, t0, t1, p.cone, s, rand, rand ? p.rot01[q] : 0.f, rand ? p.flip01[q] : 1.f, p.std_scale, m, sd, ts);
#pragma unroll
    for (int i = 0; i < 18; ++i) p.means[q * 18 + i] = m[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        p.stds[q * 6 + i] = sd[i];
        if (p.ts) p.ts[q * 6 + i] = ts[i];
    }
}

}  // namespace ucnerf

using namespace ucnerf;

extern "C" int ucnerf_cast_rays(const float* tdist, const float* origins, const float* directions, const float* cam_dirs,
                                const float* radii, const float* rand_vec, const float* rot01, const float* flip01,
                                uint32_t n_rays, int32_t S, float std_scale, float* means, float* stds, float* ts,
                                void* stream) {
    if (n_rays == 0) return 0;
    UC_REQUIRE(S >= 1, "cast_rays: S must be >= 1");
    UC_REQUIRE(tdist && origins && directions && cam_dirs && radii && rand_vec && means && stds, "cast_rays: null pointer");
    UC_REQUIRE((rot01 == nullptr) == (flip01 == nullptr), "cast_rays: rot01 and flip01 go together");
    CastParams p{};
    p.n_rays = n_rays; p.S = S; p.tdist = tdist; p.origins = origins; p.directions = directions; p.cam_dirs = cam_dirs;
    p.radii = radii; p.rand_vec = rand_vec; p.rot01 = rot01; p.flip01 = flip01; p.std_scale = std_scale;
    make_cone_table(p.cone);
    p.means = means; p.stds = stds; p.ts = ts;
    const size_t total = (size_t)n_rays * S;
    cast_rays_kernel<<<(unsigned)div_up(total, (size_t)128), 128, 0, (cudaStream_t)stream>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

// Pooled hash-grid encode for TRAINING (SURVEY.md section 8a rows R4 + R10, config 5): the front end of
// MLP.predict_density as ONE forward and ONE backward kernel at the seam where the reference holds
// `means [B,M,3]`, `stds [B,M]` (internal/models.py:L485-496).  The reference runs, per MLP call:
//   contract_mean_std (8 elementwise ATen kernels, coord.py:L60-72) -> (x+1)/2 -> kernel_grid writing [L, B*M, C]
//   (gridencoder.cu:L87-245) -> permute copy (grid.py:L57) -> erf weights [B,M,L] -> multiply -> mean over M
// and the mirror image of all of it under autograd before kernel_grid_backward (gridencoder.cu:L248-340) reads the
// permuted [L, B*M, C] gradient.  Here a thread owns one (interval, level): it contracts the M points in registers,
// gathers 8 corners per point with 128-bit read-only loads, applies the erf down-weighting and writes the pooled
// 4 features once; the backward thread recomputes the same cell / weights and scatters with one
// `red.global.add.v4.f32` per corner.  Algorithmic traffic per interval-level (M = 6): 48 x 16 B gathered (forward)
// or reduced (backward) + 16 B of features, against 6 x (16 B out + 16 B permuted + 16 B weighted) + the same again
// in the backward for the reference chain.  blockIdx.y = level keeps one level's table (<= 32 MiB) L2-resident.
//
// The per-point arithmetic lives in pooled_algos.cuh (host+device templates, also instantiated by the CPU test harness).
#include "../../include/ucnerf_b200.h"
#include "pooled_algos.cuh"

namespace ucnerf {

struct PooledLevels {
    int num_levels;
    GridLevel lv[UCNERF_MAX_GRID_LEVELS];
    float g2[UCNERF_MAX_GRID_LEVELS];   // float(grid_sizes[l]^2): torch squares the int32 buffer, then promotes
};

constexpr int kPooledThreads = 256;
constexpr int kPooledMaxM = 8;

struct TableLoad {
    const float4* table;
    __device__ __forceinline__ float4 operator()(size_t entry) const { return ldg_f4(table + entry); }
};
struct TableRedAdd {
    float4* grad;
    __device__ __forceinline__ void operator()(size_t entry, float a, float b, float c, float d) const {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(grad + entry), "f"(a), "f"(b), "f"(c), "f"(d)
                     : "memory");
    }
};

// the M points of interval b, staged in registers / local memory (M <= 8)
__device__ __forceinline__ void load_interval(const float* __restrict__ means, const float* __restrict__ stds, uint32_t b,
                                              int M, float (&mj)[3 * kPooledMaxM], float (&sj)[kPooledMaxM]) {
    const float* mp = means + (size_t)b * M * 3;
    const float* sp = stds + (size_t)b * M;
#pragma unroll
    for (int i = 0; i < 3 * kPooledMaxM; ++i)
        if (i < 3 * M) mj[i] = __ldg(mp + i);
#pragma unroll
    for (int i = 0; i < kPooledMaxM; ++i)
        if (i < M) sj[i] = __ldg(sp + i);
}

// MT: compile-time number of multisample points (6 = render.cast_rays' hexagonal pattern; 0 = runtime M)
template <int MT>
__global__ void __launch_bounds__(kPooledThreads)
pooled_forward_kernel(const float* __restrict__ means, const float* __restrict__ stds, uint32_t B, int M, int contract,
                      const float4* __restrict__ table, const __grid_constant__ PooledLevels pl,
                      float* __restrict__ features, float* __restrict__ coord) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int l = blockIdx.y;
    if (MT) M = MT;
    float mj[3 * kPooledMaxM], sj[kPooledMaxM];
    load_interval(means, stds, b, M, mj, sj);
    float F[4];
    pooled_level_forward(pl.lv[l], pl.g2[l], mj, sj, M, contract != 0, TableLoad{table}, F);
    *reinterpret_cast<float4*>(features + ((size_t)b * pl.num_levels + l) * 4) = make_float4(F[0], F[1], F[2], F[3]);
    if (coord && l == 0) {
        float c[3];
        pooled_coord(mj, sj, M, contract != 0, c);
        coord[3 * (size_t)b] = c[0]; coord[3 * (size_t)b + 1] = c[1]; coord[3 * (size_t)b + 2] = c[2];
    }
}

// RUNS: sum the corner coefficients of consecutive points in one cell before reducing (pooled_level_backward_runs)
template <int MT, bool RUNS>
__global__ void __launch_bounds__(kPooledThreads)
pooled_backward_kernel(const float* __restrict__ grad_features, const float* __restrict__ means,
                       const float* __restrict__ stds, uint32_t B, int M, int contract,
                       const __grid_constant__ PooledLevels pl, float4* __restrict__ grad_table) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int l = blockIdx.y;
    const float4 d = __ldg(reinterpret_cast<const float4*>(grad_features + ((size_t)b * pl.num_levels + l) * 4));
    if (d.x == 0.f && d.y == 0.f && d.z == 0.f && d.w == 0.f) return;   // nothing to add (also skips padded rows)
    if (MT) M = MT;
    float mj[3 * kPooledMaxM], sj[kPooledMaxM];
    load_interval(means, stds, b, M, mj, sj);
    const float dF[4] = {d.x, d.y, d.z, d.w};
    if constexpr (RUNS) pooled_level_backward_runs(pl.lv[l], pl.g2[l], mj, sj, M, contract != 0, dF, TableRedAdd{grad_table});
    else pooled_level_backward(pl.lv[l], pl.g2[l], mj, sj, M, contract != 0, dF, TableRedAdd{grad_table});
}

// K consecutive intervals per thread with run merging across them (pooled_level_backward_ray_runs); B % K == 0
template <int MT, int K>
__global__ void __launch_bounds__(kPooledThreads)
pooled_backward_ray_runs_kernel(const float* __restrict__ grad_features, const float* __restrict__ means,
                                const float* __restrict__ stds, uint32_t groups, int M, int contract,
                                const __grid_constant__ PooledLevels pl, float4* __restrict__ grad_table) {
    const uint32_t gidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gidx >= groups) return;
    const int l = blockIdx.y;
    if (MT) M = MT;
    const size_t b0 = (size_t)gidx * K;
    pooled_level_backward_ray_runs(pl.lv[l], pl.g2[l], means + b0 * M * 3, stds + b0 * M, M, K, contract != 0,
                                   grad_features + b0 * pl.num_levels * 4, pl.num_levels * 4, 4 * l, TableRedAdd{grad_table});
}

static int make_levels(PooledLevels& pl, const int32_t* offsets_host, const int32_t* grid_sizes_host, uint32_t L,
                       float S, uint32_t H) {
    UC_REQUIRE(offsets_host && grid_sizes_host, "pooled_encode: null offsets / grid_sizes");
    UC_REQUIRE(L >= 1 && L <= UCNERF_MAX_GRID_LEVELS, "pooled_encode: levels must be in [1,16]");
    pl.num_levels = (int)L;
    for (uint32_t l = 0; l < L; ++l) {
        UC_REQUIRE(offsets_host[l + 1] > offsets_host[l], "pooled_encode: offsets must increase");
        const int64_t gs = grid_sizes_host[l];
        make_grid_level(pl.lv[l], (int)l, offsets_host[l], offsets_host[l + 1], S, H, gs);
        pl.g2[l] = (float)(int32_t)(gs * gs);
    }
    return 0;
}

}  // namespace ucnerf

using namespace ucnerf;

extern "C" int ucnerf_pooled_encode_forward(const float* means, const float* stds, uint32_t B, uint32_t M, int flags,
                                            const float* embeddings, const int32_t* offsets_host,
                                            const int32_t* grid_sizes_host, uint32_t L, uint32_t C, float S, uint32_t H,
                                            float* features, float* coord, void* stream) {
    UC_REQUIRE(C == 4, "pooled_encode: level_dim must be 4");
    UC_REQUIRE(M >= 1 && M <= (uint32_t)kPooledMaxM, "pooled_encode: 1 <= multisample points <= 8");
    if (B == 0) return 0;
    UC_REQUIRE(means && stds && embeddings && features, "pooled_encode_forward: null pointer");
    PooledLevels pl;
    if (int e = make_levels(pl, offsets_host, grid_sizes_host, L, S, H)) return e;
    const dim3 grid(div_up(B, (uint32_t)kPooledThreads), L, 1);
    const int contract = flags & UCNERF_POOLED_CONTRACT;
    if (M == 6)
        pooled_forward_kernel<6><<<grid, kPooledThreads, 0, (cudaStream_t)stream>>>(
            means, stds, B, (int)M, contract, reinterpret_cast<const float4*>(embeddings), pl, features, coord);
    else
        pooled_forward_kernel<0><<<grid, kPooledThreads, 0, (cudaStream_t)stream>>>(
            means, stds, B, (int)M, contract, reinterpret_cast<const float4*>(embeddings), pl, features, coord);
    UC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ucnerf_pooled_encode_backward(const float* grad_features, const float* means, const float* stds, uint32_t B,
                                             uint32_t M, int flags, const int32_t* offsets_host,
                                             const int32_t* grid_sizes_host, uint32_t L, uint32_t C, float S, uint32_t H,
                                             float* grad_embeddings, void* stream) {
    UC_REQUIRE(C == 4, "pooled_encode: level_dim must be 4");
    UC_REQUIRE(M >= 1 && M <= (uint32_t)kPooledMaxM, "pooled_encode: 1 <= multisample points <= 8");
    if (B == 0) return 0;
    UC_REQUIRE(grad_features && means && stds && grad_embeddings, "pooled_encode_backward: null pointer");
    PooledLevels pl;
    if (int e = make_levels(pl, offsets_host, grid_sizes_host, L, S, H)) return e;
    const dim3 grid(div_up(B, (uint32_t)kPooledThreads), L, 1);
    const int contract = flags & UCNERF_POOLED_CONTRACT;
    const bool runs = (flags & UCNERF_POOLED_MERGE_RUNS) != 0;
    float4* ge = reinterpret_cast<float4*>(grad_embeddings);
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int kRayRun = 4;     // consecutive intervals per thread of the ray-run variant
    if ((flags & UCNERF_POOLED_MERGE_RAY_RUNS) && B % kRayRun == 0) {
        const uint32_t groups = B / kRayRun;
        const dim3 g2(div_up(groups, (uint32_t)kPooledThreads), L, 1);
        if (M == 6)
            pooled_backward_ray_runs_kernel<6, kRayRun><<<g2, kPooledThreads, 0, st>>>(grad_features, means, stds, groups, (int)M, contract, pl, ge);
        else
            pooled_backward_ray_runs_kernel<0, kRayRun><<<g2, kPooledThreads, 0, st>>>(grad_features, means, stds, groups, (int)M, contract, pl, ge);
        UC_LAUNCH_CHECK();
        return 0;
    }
    if (M == 6 && runs)
        pooled_backward_kernel<6, true><<<grid, kPooledThreads, 0, st>>>(grad_features, means, stds, B, (int)M, contract, pl, ge);
    else if (M == 6)
        pooled_backward_kernel<6, false><<<grid, kPooledThreads, 0, st>>>(grad_features, means, stds, B, (int)M, contract, pl, ge);
    else if (runs)
        pooled_backward_kernel<0, true><<<grid, kPooledThreads, 0, st>>>(grad_features, means, stds, B, (int)M, contract, pl, ge);
    else
        pooled_backward_kernel<0, false><<<grid, kPooledThreads, 0, st>>>(grad_features, means, stds, B, (int)M, contract, pl, ge);
    UC_LAUNCH_CHECK();
    return 0;
}

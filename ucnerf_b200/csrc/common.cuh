// Shared helpers for the ucnerf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include <atomic>
#include <string>

namespace ucnerf {

// ---- error plumbing for the C ABI ------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launch_count;

#define UC_CUDA_OK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::ucnerf::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
            return 2;                                                                             \
        }                                                                                         \
    } while (0)

#define UC_REQUIRE(cond, msg)                                                                     \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            ::ucnerf::set_error(msg);                                                             \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

// every kernel launch of the library goes through this so gpu_launches can be reported
#define UC_LAUNCH_CHECK()                                                                         \
    do {                                                                                          \
        ::ucnerf::g_launch_count.fetch_add(1, std::memory_order_relaxed);                         \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            ::ucnerf::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e));  \
            return 3;                                                                             \
        }                                                                                         \
    } while (0)

// Opt a kernel into > 48 KB of dynamic shared memory.  cudaFuncSetAttribute is per DEVICE (and context), so what has been
// configured is remembered per device, not in one process-wide flag: a process that renders on several GPUs (or from
// several threads; the call is idempotent) never launches an unconfigured kernel.  Usage: UC_ENSURE_SMEM(bytes, kernel<..>)
#define UC_ENSURE_SMEM(bytes, ...)                                                                               \
    do {                                                                                                         \
        static std::atomic<size_t> _uc_cfg[64];                                                                  \
        const size_t _uc_b = (size_t)(bytes);                                                                    \
        if (_uc_b > 48 * 1024) {                                                                                 \
            int _uc_dev = 0;                                                                                     \
            UC_CUDA_OK(cudaGetDevice(&_uc_dev));                                                                 \
            const int _uc_i = (_uc_dev >= 0 && _uc_dev < 64) ? _uc_dev : 63;                                     \
            if (_uc_dev != _uc_i || _uc_cfg[_uc_i].load(std::memory_order_relaxed) < _uc_b) {                    \
                UC_CUDA_OK(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)_uc_b)); \
                _uc_cfg[_uc_i].store(_uc_b, std::memory_order_relaxed);                                          \
            }                                                                                                    \
        }                                                                                                        \
    } while (0)

template <typename T>
__host__ __device__ inline T div_up(T a, T b) {
    return (a + b - 1) / b;
}

constexpr int kNumSMs = 148;  // B200

// ---- 128-bit read-only gather (hash-table entries are read-only during a render) -------------
#if defined(__CUDACC__)
__device__ __forceinline__ float4 ldg_f4(const float4* p) {
    return __ldg(p);
}
#endif

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2): two IEEE fp32 operations per issue slot -------------
#if defined(__CUDACC__)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a);
    rb = *reinterpret_cast<unsigned long long*>(&b);
    rc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a);
    rb = *reinterpret_cast<unsigned long long*>(&b);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
#endif

// ---- per-level geometry of a hash grid, precomputed on the host for the fused path -----------
struct GridLevel {
    uint32_t offset;        // entry offset of the level inside `embeddings`
    uint32_t hashmap_size;  // entries in the level
    uint32_t stride1;       // resolution + 1 (align_corners=False): dense index stride
    uint32_t stride2;       // stride1 * stride1
    uint32_t hashed;        // 1: XOR-prime hash, 0: dense index
    uint32_t pow2_mask;     // hashmap_size-1 if power of two else 0 (then use %)
    float scale;            // exp2f(l*S)*H - 1
    float grid_size;        // python-side grid_sizes[l] (erf down-weighting), as float
    uint32_t mod_mode;      // 0: index needs no reduction, 1: & pow2_mask, 2: % hashmap_size
};

struct GridDesc {
    const float4* table;  // C == 4 on the fused path
    int num_levels;
    GridLevel lv[16];
};

// Level l of a GridEncoder with offsets[l..l+1], log2(per_level_scale) and base resolution (D = 3, gridtype hash,
// align_corners = False): the per-level constants of kernel_grid (gridencoder.cu:L137-143, L50-84) precomputed on the host.
inline void make_grid_level(GridLevel& g, int l, int32_t off0, int32_t off1, float log2_per_level_scale,
                            uint32_t base_resolution, int64_t grid_size) {
    g.offset = (uint32_t)off0;
    g.hashmap_size = (uint32_t)(off1 - off0);
    const float scale = exp2f((float)l * log2_per_level_scale) * (float)base_resolution - 1.0f;
    const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
    g.scale = scale;
    g.stride1 = resolution + 1;
    g.stride2 = g.stride1 * g.stride1;
    uint32_t stride = 1;
    for (int d = 0; d < 3 && stride <= g.hashmap_size; ++d) stride *= (resolution + 1);
    g.hashed = stride > g.hashmap_size ? 1u : 0u;
    g.pow2_mask = (g.hashmap_size & (g.hashmap_size - 1)) == 0 ? g.hashmap_size - 1 : 0u;
    if (g.hashmap_size == 1) g.pow2_mask = 0;
    g.grid_size = (float)grid_size;
    // dense index of an in-range point: max = (res+1)^3 - 1 < hashmap_size  => no modulo needed.  The stride product
    // above is 32-bit like the reference's (gridencoder.cu:L72-77): for resolution + 1 >= 65537 it wraps, a level can
    // then look dense although (res+1)^3 exceeds the table; the reference indexes such a level with the wrapped
    // strides and reduces modulo the table size - so do we (same stride1 / stride2 arithmetic, mod_mode != 0).
    bool fits = true;
    uint64_t full = 1;
    for (int d = 0; d < 3 && fits; ++d) {
        full *= (uint64_t)(resolution + 1);
        fits = full <= (uint64_t)g.hashmap_size;
    }
    if (!g.hashed && fits) g.mod_mode = 0;
    else g.mod_mode = g.pow2_mask ? 1u : 2u;
}

}  // namespace ucnerf

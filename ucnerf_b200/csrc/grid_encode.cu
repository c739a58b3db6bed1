// Stand-alone multi-resolution hash-grid encoder: the `_gridencoder` drop-in kernels.
//
// Replaces (reference /root/reference/nerf/gridencoder/src/gridencoder.cu):
//   kernel_grid            L87-245   -> grid_forward_kernel
//   kernel_grid_backward   L248-340  -> grid_backward_kernel
//   kernel_input_backward  L343-369  -> input_backward_kernel
//   kernel_grad_tv         L506-610  -> grad_tv_kernel
//
// B200 design notes
//  * one thread = one (point, level), level = blockIdx.y exactly like the reference so that blocks
//    of one level are scheduled together and a level's table (<= 32 MiB at T=2^21, C=4, fp32) stays
//    resident in the 126 MB L2 while it is being gathered; grids are sized in multiples of 148 SMs
//    worth of 256-thread CTAs by the caller's B, nothing else to tune: the kernel is a random
//    16-byte gather and is bound by L2/HBM sector throughput.
//  * a feature vector (C scalars) is fetched with ONE vector load (ld.global.nc.v2/.v4) instead
//    of C scalar loads, and written back with one vector store.
//  * backward accumulates a whole C=4 fp32 feature vector with one `red.global.add.v4.f32`
//    (sm_90+) instead of four scalar atomics; fp16 uses red.add.noftz.f16x2 like the reference.
//  * launches go to the caller's stream (the reference launches on the legacy default stream).
// The arithmetic is written in the same expression forms as the reference so that nvcc makes the
// same FMA-contraction decisions: forward outputs are bit-identical to the reference kernel for
// fp32/fp64 (checked on the GPU box by tests/test_gpu_grid_vs_ref.py).
#include "common.cuh"
#include "../../include/ucnerf_b200.h"

namespace ucnerf {

template <typename T> struct Vec;  // C-wide vector load/store helpers

template <typename T, uint32_t C>
__device__ __forceinline__ void load_feat(const T* __restrict__ p, T (&v)[C]) {
    if constexpr (sizeof(T) * C == 16) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        *reinterpret_cast<uint4*>(v) = r;
    } else if constexpr (sizeof(T) * C == 8) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        *reinterpret_cast<uint2*>(v) = r;
    } else if constexpr (sizeof(T) * C == 4) {
        const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(p));
        *reinterpret_cast<uint32_t*>(v) = r;
    } else if constexpr (sizeof(T) * C == 32) {
        const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(p));
        const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        reinterpret_cast<uint4*>(v)[0] = r0;
        reinterpret_cast<uint4*>(v)[1] = r1;
    } else if constexpr (sizeof(T) * C == 64) {
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(v)[i] = __ldg(reinterpret_cast<const uint4*>(p) + i);
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) v[c] = p[c];
    }
}

template <typename T, uint32_t C>
__device__ __forceinline__ void store_feat(T* __restrict__ p, const T (&v)[C]) {
    if constexpr (sizeof(T) * C == 16) {
        *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v);
    } else if constexpr (sizeof(T) * C == 8) {
        *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(v);
    } else if constexpr (sizeof(T) * C == 32) {
        reinterpret_cast<uint4*>(p)[0] = reinterpret_cast<const uint4*>(v)[0];
        reinterpret_cast<uint4*>(p)[1] = reinterpret_cast<const uint4*>(v)[1];
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) p[c] = v[c];
    }
}

// fp16 arithmetic of the reference goes through c10::Half: every binary op is evaluated in fp32 and
// the result rounded back to half (c10/util/Half-inl.h).  These helpers restate that.
template <typename T> struct Arith {
    using acc_t = T;
    static __device__ __forceinline__ T zero() { return T(0); }
    // results += w * g            (gridencoder.cu:L187)
    static __device__ __forceinline__ void mac(T& acc, float w, T g) { acc += w * g; }
    // results_grad += w * (gr - gl) * deriv     (gridencoder.cu:L235)
    static __device__ __forceinline__ void mac_diff(T& acc, float w, T gr, T gl, float deriv) {
        acc += w * (gr - gl) * deriv;
    }
};
template <> struct Arith<__half> {
    static __device__ __forceinline__ __half zero() { return __float2half(0.f); }
    static __device__ __forceinline__ void mac(__half& acc, float w, __half g) {
        const __half prod = __float2half(w * __half2float(g));
        acc = __float2half(__half2float(acc) + __half2float(prod));
    }
    static __device__ __forceinline__ void mac_diff(__half& acc, float w, __half gr, __half gl, float deriv) {
        const __half diff = __float2half(__half2float(gr) - __half2float(gl));
        const __half prod = __float2half(w * __half2float(diff) * deriv);
        acc = __float2half(__half2float(acc) + __half2float(prod));
    }
};

__device__ __forceinline__ float smoothstep_f(float v) { return v * v * (3.0f - 2.0f * v); }
__device__ __forceinline__ float smoothstep_d(float v) { return 6 * v * (1.0f - v); }

// fast_hash + get_grid_index, gridencoder.cu:L50-84 (entry index, before `* C`).
template <uint32_t D>
__device__ __forceinline__ uint32_t grid_index(const uint32_t gridtype, const bool align_corners,
                                               const uint32_t hashmap_size, const uint32_t resolution,
                                               const uint32_t (&pg)[D]) {
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pg[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) {
        uint32_t h = 0;
#pragma unroll
        for (uint32_t i = 0; i < D; ++i) h ^= pg[i] * primes[i];
        index = h;
    }
    return index % hashmap_size;
}

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256)
grid_forward_kernel(const float* __restrict__ inputs, const T* __restrict__ grid, const int* __restrict__ offsets,
                    T* __restrict__ outputs, const uint32_t B, const uint32_t L, const float S, const uint32_t H,
                    T* __restrict__ dy_dx, const uint32_t gridtype, const bool align_corners, const uint32_t interp) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;

    grid += (size_t)(uint32_t)offsets[level] * C;
    inputs += (size_t)b * D;
    outputs += ((size_t)level * B + b) * C;

    float x[D];
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        x[d] = inputs[d];
        if (x[d] < 0 || x[d] > 1) oob = true;
    }
    if (oob) {  // gridencoder.cu:L118-135
        T z[C];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) z[c] = Arith<T>::zero();
        store_feat<T, C>(outputs, z);
        if (dy_dx) {
            T* o = dy_dx + ((size_t)b * L + level) * D * C;
#pragma unroll
            for (uint32_t d = 0; d < D; d++) store_feat<T, C>(o + d * C, z);
        }
        return;
    }

    const uint32_t hashmap_size = offsets[level + 1] - offsets[level];
    const float scale = exp2f(level * S) * H - 1.0f;
    const uint32_t resolution = (uint32_t)ceil(scale) + 1;

    float pos[D], pos_deriv[D];
    uint32_t pg[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        pos[d] = x[d] * scale + (align_corners ? 0.0f : 0.5f);
        pg[d] = floorf(pos[d]);
        pos[d] -= (float)pg[d];
        if (interp == 1) {
            pos_deriv[d] = smoothstep_d(pos[d]);
            pos[d] = smoothstep_f(pos[d]);
        } else {
            pos_deriv[d] = 1.0f;
        }
    }

    // issue all 2^D gathers first (independent), then reduce in the reference's corner order
    T res[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) res[c] = Arith<T>::zero();
    constexpr uint32_t NC = 1u << D;
    constexpr uint32_t BATCH = (NC > 8) ? 8 : NC;
#pragma unroll
    for (uint32_t base = 0; base < NC; base += BATCH) {
        T feat[BATCH][C];
        float w[BATCH];
#pragma unroll
        for (uint32_t i = 0; i < BATCH; i++) {
            const uint32_t idx = base + i;
            float wi = 1;
            uint32_t pl[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) {
                if ((idx & (1u << d)) == 0) {
                    wi *= 1 - pos[d];
                    pl[d] = pg[d];
                } else {
                    wi *= pos[d];
                    pl[d] = pg[d] + 1;
                }
            }
            w[i] = wi;
            const uint32_t index = grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pl);
            load_feat<T, C>(grid + (size_t)index * C, feat[i]);
        }
#pragma unroll
        for (uint32_t i = 0; i < BATCH; i++) {
#pragma unroll
            for (uint32_t c = 0; c < C; c++) Arith<T>::mac(res[c], w[i], feat[i][c]);
        }
    }
    store_feat<T, C>(outputs, res);

    if (dy_dx) {  // gridencoder.cu:L199-244
        T* o = dy_dx + ((size_t)b * L + level) * D * C;
#pragma unroll
        for (uint32_t gd = 0; gd < D; gd++) {
            T rg[C];
#pragma unroll
            for (uint32_t c = 0; c < C; c++) rg[c] = Arith<T>::zero();
#pragma unroll
            for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                float w = scale;
                uint32_t pl[D];
#pragma unroll
                for (uint32_t nd = 0; nd < D - 1; nd++) {
                    const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                    if ((idx & (1u << nd)) == 0) {
                        w *= 1 - pos[d];
                        pl[d] = pg[d];
                    } else {
                        w *= pos[d];
                        pl[d] = pg[d] + 1;
                    }
                }
                pl[gd] = pg[gd];
                const uint32_t il = grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pl);
                pl[gd] = pg[gd] + 1;
                const uint32_t ir = grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pl);
                T fl[C], fr[C];
                load_feat<T, C>(grid + (size_t)il * C, fl);
                load_feat<T, C>(grid + (size_t)ir * C, fr);
#pragma unroll
                for (uint32_t c = 0; c < C; c++) Arith<T>::mac_diff(rg[c], w, fr[c], fl[c], pos_deriv[gd]);
            }
            store_feat<T, C>(o + gd * C, rg);
        }
    }
}

// ---- backward -------------------------------------------------------------------------------
template <typename T, uint32_t C>
__device__ __forceinline__ void red_add_feat(T* __restrict__ dst, float w, const T (&g)[C]) {
    if constexpr (std::is_same<T, float>::value && C == 4) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(w * g[0]), "f"(w * g[1]),
                     "f"(w * g[2]), "f"(w * g[3])
                     : "memory");
    } else if constexpr (std::is_same<T, float>::value && C == 2) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(w * g[0]), "f"(w * g[1]) : "memory");
    } else if constexpr (std::is_same<T, float>::value && C == 8) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(w * g[0]), "f"(w * g[1]),
                     "f"(w * g[2]), "f"(w * g[3])
                     : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(w * g[4]), "f"(w * g[5]),
                     "f"(w * g[6]), "f"(w * g[7])
                     : "memory");
    } else if constexpr (std::is_same<T, __half>::value && (C % 2 == 0)) {
#pragma unroll
        for (uint32_t c = 0; c < C; c += 2) {  // gridencoder.cu:L325-331
            __half2 v = __halves2half2(__float2half(w * __half2float(g[c])), __float2half(w * __half2float(g[c + 1])));
            atomicAdd(reinterpret_cast<__half2*>(dst + c), v);
        }
    } else if constexpr (std::is_same<T, __half>::value) {
#pragma unroll
        for (uint32_t c = 0; c < C; c++) atomicAdd(dst + c, __float2half(w * __half2float(g[c])));
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; c++) atomicAdd(dst + c, (T)(w * g[c]));
    }
}

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256)
grid_backward_kernel(const T* __restrict__ grad, const float* __restrict__ inputs, const int* __restrict__ offsets,
                     T* __restrict__ grad_grid, const uint32_t B, const uint32_t L, const float S, const uint32_t H,
                     const uint32_t gridtype, const bool align_corners, const uint32_t interp) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    grad_grid += (size_t)(uint32_t)offsets[level] * C;
    inputs += (size_t)b * D;
    grad += ((size_t)level * B + b) * C;

    const uint32_t hashmap_size = offsets[level + 1] - offsets[level];
    const float scale = exp2f(level * S) * H - 1.0f;
    const uint32_t resolution = (uint32_t)ceil(scale) + 1;

    float pos[D];
    uint32_t pg[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const float x = inputs[d];
        if (x < 0 || x > 1) return;  // gridencoder.cu:L276-281
        pos[d] = x * scale + (align_corners ? 0.0f : 0.5f);
        pg[d] = floorf(pos[d]);
        pos[d] -= (float)pg[d];
        if (interp == 1) pos[d] = smoothstep_f(pos[d]);
    }
    T g[C];
    load_feat<T, C>(grad, g);
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); idx++) {
        float w = 1;
        uint32_t pl[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if ((idx & (1u << d)) == 0) {
                w *= 1 - pos[d];
                pl[d] = pg[d];
            } else {
                w *= pos[d];
                pl[d] = pg[d] + 1;
            }
        }
        const uint32_t index = grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pl);
        red_add_feat<T, C>(grad_grid + (size_t)index * C, w, g);
    }
}

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256)
input_backward_kernel(const T* __restrict__ grad, const T* __restrict__ dy_dx, T* __restrict__ grad_inputs,
                      uint32_t B, uint32_t L) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t >= B * D) return;
    const uint32_t b = t / D;
    const uint32_t d = t - b * D;
    dy_dx += (size_t)b * L * D * C;
    T result = Arith<T>::zero();
    for (uint32_t l = 0; l < L; l++) {
        T gv[C], dv[C];
        load_feat<T, C>(grad + ((size_t)l * B + b) * C, gv);
        load_feat<T, C>(dy_dx + ((size_t)l * D + d) * C, dv);
#pragma unroll
        for (uint32_t ch = 0; ch < C; ch++) {
            if constexpr (std::is_same<T, __half>::value) {
                const __half prod = __float2half(__half2float(gv[ch]) * __half2float(dv[ch]));
                result = __float2half(__half2float(result) + __half2float(prod));
            } else {
                result += gv[ch] * dv[ch];
            }
        }
    }
    grad_inputs[t] = result;
}

// ---- total-variation gradient ---------------------------------------------------------------
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256)
grad_tv_kernel(const T* __restrict__ inputs, const T* __restrict__ grid, T* __restrict__ grad,
               const int* __restrict__ offsets, const float weight, const uint32_t B, const uint32_t L, const float S,
               const uint32_t H, const uint32_t gridtype, const bool align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    inputs += (size_t)b * D;
    grid += (size_t)(uint32_t)offsets[level] * C;
    grad += (size_t)(uint32_t)offsets[level] * C;

    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        x[d] = (float)inputs[d];
        if (x[d] < 0 || x[d] > 1) return;
    }
    const uint32_t hashmap_size = offsets[level + 1] - offsets[level];
    const float scale = exp2f(level * S) * H - 1.0f;
    const uint32_t resolution = (uint32_t)ceil(scale) + 1;

    uint32_t pg[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const float p = x[d] * scale + (align_corners ? 0.0f : 0.5f);
        pg[d] = floorf(p);
    }
    float results[C], idelta[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) results[c] = idelta[c] = 0.f;
    const uint32_t index = grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pg);
    T center[C];
    load_feat<T, C>(grid + (size_t)index * C, center);
    const float w = weight / (2 * D);
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const uint32_t cur = pg[d];
        if (cur < resolution) {
            pg[d] = cur + 1;
            const uint32_t ir = grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pg);
            T f[C];
            load_feat<T, C>(grid + (size_t)ir * C, f);
#pragma unroll
            for (uint32_t c = 0; c < C; c++) {
                const float gv = (float)center[c] - (float)f[c];
                results[c] += gv;
                idelta[c] += gv * gv;
            }
        }
        if (cur > 0) {
            pg[d] = cur - 1;
            const uint32_t il = grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pg);
            T f[C];
            load_feat<T, C>(grid + (size_t)il * C, f);
#pragma unroll
            for (uint32_t c = 0; c < C; c++) {
                const float gv = (float)center[c] - (float)f[c];
                results[c] += gv;
                idelta[c] += gv * gv;
            }
        }
        pg[d] = cur;
    }
#pragma unroll
    for (uint32_t c = 0; c < C; c++) {
        const float v = w * results[c] * rsqrtf(idelta[c] + 1e-9f);
        if constexpr (std::is_same<T, __half>::value)
            atomicAdd(grad + (size_t)index * C + c, __float2half(v));
        else
            atomicAdd(grad + (size_t)index * C + c, (T)v);
    }
}

// ---- dispatch -------------------------------------------------------------------------------
template <typename T, uint32_t D, uint32_t C>
static int launch_forward(const float* inputs, const void* emb, const int* offsets, void* out, uint32_t B, uint32_t L,
                          float S, uint32_t H, void* dy_dx, uint32_t gridtype, bool ac, uint32_t interp,
                          cudaStream_t st) {
    if (B == 0 || L == 0) return 0;
    dim3 grid(div_up(B, 256u), L, 1);
    grid_forward_kernel<T, D, C><<<grid, 256, 0, st>>>(inputs, (const T*)emb, offsets, (T*)out, B, L, S, H, (T*)dy_dx,
                                                       gridtype, ac, interp);
    UC_LAUNCH_CHECK();
    return 0;
}

template <typename T, uint32_t D, uint32_t C>
static int launch_backward(const void* grad, const float* inputs, const int* offsets, void* gemb, uint32_t B,
                           uint32_t L, float S, uint32_t H, const void* dy_dx, void* ginp, uint32_t gridtype, bool ac,
                           uint32_t interp, cudaStream_t st) {
    if (B == 0 || L == 0) return 0;
    dim3 grid(div_up(B, 256u), L, 1);
    grid_backward_kernel<T, D, C><<<grid, 256, 0, st>>>((const T*)grad, inputs, offsets, (T*)gemb, B, L, S, H,
                                                        gridtype, ac, interp);
    UC_LAUNCH_CHECK();
    if (dy_dx) {
        input_backward_kernel<T, D, C><<<div_up(B * D, 256u), 256, 0, st>>>((const T*)grad, (const T*)dy_dx,
                                                                            (T*)ginp, B, L);
        UC_LAUNCH_CHECK();
    }
    return 0;
}

template <typename T, uint32_t D, uint32_t C>
static int launch_tv(const void* inputs, const void* emb, void* grad, const int* offsets, float weight, uint32_t B,
                     uint32_t L, float S, uint32_t H, uint32_t gridtype, bool ac, cudaStream_t st) {
    if (B == 0 || L == 0) return 0;
    dim3 grid(div_up(B, 256u), L, 1);
    grad_tv_kernel<T, D, C><<<grid, 256, 0, st>>>((const T*)inputs, (const T*)emb, (T*)grad, offsets, weight, B, L, S,
                                                  H, gridtype, ac);
    UC_LAUNCH_CHECK();
    return 0;
}

#define UC_DISPATCH_DC(T, FN, ...)                                                       \
    switch (D * 16 + C) {                                                                \
        case 2 * 16 + 1: return FN<T, 2, 1>(__VA_ARGS__);                                \
        case 2 * 16 + 2: return FN<T, 2, 2>(__VA_ARGS__);                                \
        case 2 * 16 + 4: return FN<T, 2, 4>(__VA_ARGS__);                                \
        case 2 * 16 + 8: return FN<T, 2, 8>(__VA_ARGS__);                                \
        case 3 * 16 + 1: return FN<T, 3, 1>(__VA_ARGS__);                                \
        case 3 * 16 + 2: return FN<T, 3, 2>(__VA_ARGS__);                                \
        case 3 * 16 + 4: return FN<T, 3, 4>(__VA_ARGS__);                                \
        case 3 * 16 + 8: return FN<T, 3, 8>(__VA_ARGS__);                                \
        case 4 * 16 + 1: return FN<T, 4, 1>(__VA_ARGS__);                                \
        case 4 * 16 + 2: return FN<T, 4, 2>(__VA_ARGS__);                                \
        case 4 * 16 + 4: return FN<T, 4, 4>(__VA_ARGS__);                                \
        case 4 * 16 + 8: return FN<T, 4, 8>(__VA_ARGS__);                                \
        case 5 * 16 + 1: return FN<T, 5, 1>(__VA_ARGS__);                                \
        case 5 * 16 + 2: return FN<T, 5, 2>(__VA_ARGS__);                                \
        case 5 * 16 + 4: return FN<T, 5, 4>(__VA_ARGS__);                                \
        case 5 * 16 + 8: return FN<T, 5, 8>(__VA_ARGS__);                                \
        default: break;                                                                  \
    }

#define UC_DISPATCH_T(FN, ...)                                                           \
    if (dtype == UCNERF_F32) { UC_DISPATCH_DC(float, FN, __VA_ARGS__) }                  \
    else if (dtype == UCNERF_F16) { UC_DISPATCH_DC(__half, FN, __VA_ARGS__) }            \
    else if (dtype == UCNERF_F64) { UC_DISPATCH_DC(double, FN, __VA_ARGS__) }            \
    else { set_error("GridEncoding: dtype must be f32, f16 or f64"); return 1; }

static int check_dc(uint32_t D, uint32_t C) {
    UC_REQUIRE(D >= 2 && D <= 5, "GridEncoding: D must be 2, 3, 4, or 5.");
    UC_REQUIRE(C == 1 || C == 2 || C == 4 || C == 8, "GridEncoding: C must be 1, 2, 4, or 8.");
    return 0;
}

}  // namespace ucnerf

using namespace ucnerf;

extern "C" int ucnerf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                                          void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                          uint32_t H, void* dy_dx, uint32_t gridtype, int align_corners,
                                          uint32_t interp, int dtype, void* stream) {
    if (int e = check_dc(D, C)) return e;
    UC_REQUIRE(B == 0 || (inputs && embeddings && offsets && outputs), "grid_encode_forward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const bool ac = align_corners != 0;
    UC_DISPATCH_T(launch_forward, inputs, embeddings, offsets, outputs, B, L, S, H, dy_dx, gridtype, ac, interp, st)
    set_error("grid_encode_forward: unsupported D/C");
    return 1;
}

extern "C" int ucnerf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                                           const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D,
                                           uint32_t C, uint32_t L, float S, uint32_t H, const void* dy_dx,
                                           void* grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                                           int dtype, void* stream) {
    (void)embeddings;
    if (int e = check_dc(D, C)) return e;
    UC_REQUIRE(B == 0 || (grad && inputs && offsets && grad_embeddings), "grid_encode_backward: null pointer");
    UC_REQUIRE((dy_dx == nullptr) == (grad_inputs == nullptr), "grid_encode_backward: dy_dx and grad_inputs go together");
    cudaStream_t st = (cudaStream_t)stream;
    const bool ac = align_corners != 0;
    UC_DISPATCH_T(launch_backward, grad, inputs, offsets, grad_embeddings, B, L, S, H, dy_dx, grad_inputs, gridtype,
                  ac, interp, st)
    set_error("grid_encode_backward: unsupported D/C");
    return 1;
}

extern "C" int ucnerf_grad_total_variation(const void* inputs, const void* embeddings, void* grad,
                                           const int32_t* offsets, float weight, uint32_t B, uint32_t D, uint32_t C,
                                           uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                           int dtype, void* stream) {
    if (int e = check_dc(D, C)) return e;
    UC_REQUIRE(B == 0 || (inputs && embeddings && grad && offsets), "grad_total_variation: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const bool ac = align_corners != 0;
    UC_DISPATCH_T(launch_tv, inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, ac, st)
    set_error("grad_total_variation: unsupported D/C");
    return 1;
}

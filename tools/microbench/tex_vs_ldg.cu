// Can texture fetches (TEX data pipe) run beside LDG.128 gathers (LSU data pipe)?  float4 gathers from an L2-resident
// table, lanes of a warp clustered in a few 128-byte lines like the hash-grid gathers of sample_encode_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_vs_ldg tex_vs_ldg.cu && ./tex_vs_ldg
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 256;

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// MODE 0: 8 LDG per iteration; 1: 8 TEX; 2: 5 LDG + 3 TEX; 3: 6 LDG + 2 TEX; 4: 4 + 4
template <int MODE>
__global__ void __launch_bounds__(128, 5) gather(const float4* __restrict__ tab, cudaTextureObject_t tex, uint32_t mask, float* out, int spread) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int NT = MODE == 0 ? 0 : MODE == 1 ? 8 : MODE == 2 ? 3 : MODE == 3 ? 2 : 4;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        const uint32_t base = mix(warp * 977u + it) & mask;
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // corner k: a different line per corner, lanes spread over `spread` entries around it
            const uint32_t idx = (base + 4099u * k + (lane % spread)) & mask;
            if (k < 8 - NT) v[k] = __ldg(tab + idx);
            else v[k] = tex1Dfetch<float4>(tex, (int)idx);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.f) out[0] = acc.x;
}

template <int MODE>
float run(const float4* tab, cudaTextureObject_t tex, uint32_t mask, float* out, int spread) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = 148 * 5 * 8;
    gather<MODE><<<blocks, 128>>>(tab, tex, mask, out, spread);
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) gather<MODE><<<blocks, 128>>>(tab, tex, mask, out, spread);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double gathers = 5.0 * blocks * 128.0 * kIters * 8.0;
    return (float)(gathers / (ms * 1e-3) / 1e9);   // G lane-gathers / s
}

int main() {
    const uint32_t n = 1u << 22;   // 4 M entries x 16 B = 64 MB (L2 resident)
    float4* tab; float* out;
    cudaMalloc(&tab, (size_t)n * 16); cudaMalloc(&out, 16);
    cudaMemset(tab, 0, (size_t)n * 16);
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = (size_t)n * 16;
    cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) { printf("texture object failed\n"); return 1; }
    for (int spread : {1, 4, 8, 32}) {
        printf("spread %2d entries/warp-corner: LDG8 %.1f  TEX8 %.1f  LDG5+TEX3 %.1f  LDG6+TEX2 %.1f  LDG4+TEX4 %.1f  G gathers/s\n", spread,
               run<0>(tab, tex, n - 1, out, spread), run<1>(tab, tex, n - 1, out, spread), run<2>(tab, tex, n - 1, out, spread),
               run<3>(tab, tex, n - 1, out, spread), run<4>(tab, tex, n - 1, out, spread));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

// Microbenchmark: throughput of the legacy warp-level tensor-core path (mma.sync) on B200 - is it worth moving the
// 24/40 -> 64 density layer of sample_encode_kernel onto it?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

template <int KIND>   // 0: f16 m16n8k16, 1: tf32 m16n8k8
__global__ void __launch_bounds__(256) rate_kernel(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x ^ 5u, 11u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {   // 8 independent accumulators per warp
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int kind = 0; kind < 2; ++kind) {
        for (int warps_per_sm : {4, 8, 16, 32}) {
            const int blocks = 148 * (warps_per_sm * 32 / 256 > 0 ? warps_per_sm * 32 / 256 : 1);
            const int threads = warps_per_sm * 32 < 256 ? warps_per_sm * 32 : 256;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (kind == 0) rate_kernel<0><<<blocks, threads>>>(out, iters);
                else rate_kernel<1><<<blocks, threads>>>(out, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double mmas = (double)blocks * (threads / 32) * 8.0 * iters;
            const double macs = mmas * (kind == 0 ? 16 * 8 * 16 : 16 * 8 * 8);
            printf("%s warps/SM=%2d: %.3f ms, %.1f MAC/clk/SM (at 1.965 GHz), %.1f TFLOP/s\n", kind == 0 ? "f16  m16n8k16" : "tf32 m16n8k8 ",
                   warps_per_sm, ms, macs / (ms * 1e-3) / 148 / 1.965e9, 2 * macs / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}

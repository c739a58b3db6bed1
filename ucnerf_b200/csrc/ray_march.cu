// Fused forward-render kernels for sm_100a: resample -> sample/encode/density-MLP -> colour MLP ->
// composite.  One launch of each per sampling level and chunk; see DESIGN.md for the data flow and
// roofline of each kernel.  Reference path replaced: internal/models.py:L152-311 (level loop).
#include "ray_march.cuh"

namespace ucnerf {

// =================================================================================================
// 1. resample: one warp per ray (stepfun.max_dilate_weights + sample_intervals, rand=False)
// =================================================================================================
constexpr int kWarpsPerBlockRay = 4;

__global__ void __launch_bounds__(32 * kWarpsPerBlockRay)
resample_kernel(const ResampleParams p) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5;
    const uint32_t ray = blockIdx.x * kWarpsPerBlockRay + warp;
    if (ray >= p.n_rays) return;
    WarpExec ex{(int)(threadIdx.x & 31)};
    ResampleScratch sc;
    sc.carve(smem + (size_t)warp * ResampleScratch::floats(p.n_prev, p.S), p.n_prev, p.S);
    const float* tprev = p.t_prev ? p.t_prev + (size_t)ray * p.t_prev_stride : nullptr;
    const float* wprev = p.w_prev ? p.w_prev + (size_t)ray * p.n_prev : nullptr;
    resample_ray(ex, p.n_prev, tprev, wprev, p.dilate != 0, p.dilation, p.anneal, p.padding, p.S, p.u, sc,
                 p.out_sdist + (size_t)ray * (p.S + 1));
    if (p.dbg_scratch && ray == p.dbg_ray) {
        __syncwarp();
        const float* base = smem + (size_t)warp * ResampleScratch::floats(p.n_prev, p.S);
        for (int i = ex.lane; i < (int)ResampleScratch::floats(p.n_prev, p.S); i += 32) p.dbg_scratch[i] = base[i];
    }
}

int launch_resample(const ResampleParams& p, cudaStream_t st) {
    if (p.n_rays == 0) return 0;
    const size_t smem = kWarpsPerBlockRay * ResampleScratch::floats(p.n_prev, p.S) * sizeof(float);
    UC_REQUIRE(smem <= 227 * 1024, "resample: too many samples per ray for shared memory");
    UC_ENSURE_SMEM(smem, resample_kernel);
    resample_kernel<<<div_up(p.n_rays, (uint32_t)kWarpsPerBlockRay), 32 * kWarpsPerBlockRay, smem, st>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

// =================================================================================================
// 2b. sample coordinates ("coord" of ray_history, models.py:L677): only launched when the caller asks for them
//     (render_image(return_weights=True) for extract.py); thread = (ray, sample)
// =================================================================================================
__global__ void __launch_bounds__(256)
sample_coord_kernel(const __grid_constant__ SampleParams p, float* __restrict__ out) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (size_t)p.n_rays * p.S) return;
    const uint32_t ray = (uint32_t)(q / p.S);
    const int s = (int)(q - (size_t)ray * p.S);
    RayGeom rg;
    make_ray_geom(rg, p.rays.origins + 3 * (size_t)ray, p.rays.directions + 3 * (size_t)ray, p.rays.cam_dirs + 3 * (size_t)ray,
                  p.rays.rand_vec + 3 * (size_t)ray, p.rays.radii[ray], p.rays.near[ray], p.rays.far[ray]);
    const float s0 = p.sdist[(size_t)ray * p.sdist_stride + s], s1 = p.sdist[(size_t)ray * p.sdist_stride + s + 1];
    const float t0 = fa(fm(s0, rg.far), fm(fs(1.f, s0), rg.near));
    const float t1 = fa(fm(s1, rg.far), fm(fs(1.f, s1), rg.near));
    float c[3];
    interval_coord(rg, t0, t1, p.cone, s & 1, p.std_scale, c);
    out[3 * q] = c[0]; out[3 * q + 1] = c[1]; out[3 * q + 2] = c[2];
}

int launch_sample_coord(const SampleParams& p, float* out, cudaStream_t st) {
    const size_t total = (size_t)p.n_rays * p.S;
    if (total == 0) return 0;
    sample_coord_kernel<<<(unsigned)div_up(total, (size_t)256), 256, 0, st>>>(p, out);
    UC_LAUNCH_CHECK();
    return 0;
}

// =================================================================================================
// 3. colour MLP, fp32 SIMT path: 64-row tiles, register-tiled GEMM chain in shared memory.
//    Reference (models.py:L587-674): x = W2 h1 + b2 ; in = [x, direnc] ; a = relu(V0 in + c0) ;
//    a2 = relu(V1 [a, in] + c1) ; rgb = sigmoid(R a2 + r0) * (1 + 2 pad) - pad.  The bottleneck x has no
//    activation, so W2 is folded into V0 / V1 on the host (ColorParams): 2.3x fewer MACs per sample.
// =================================================================================================
constexpr int kColorThreads = 256;
constexpr int kTileRows = 64;
constexpr int kKC = 16;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// acc[r][0..7] += A[rows of this thread][0..K) * Wt[0..K)[cols of this thread];  K % 16 == 0
template <int NP>
__device__ __forceinline__ void gemm_acc(float (&acc)[NP / 32][8], const float* __restrict__ A, int lda, int K,
                                         const float* __restrict__ Wt, float* wbuf) {
    constexpr int TX = NP / 8;
    constexpr int RPT = NP / 32;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    constexpr int CHUNK_F4 = kKC * NP / 4;
    const int nchunks = K / kKC;
    auto prefetch = [&](int c, int buf) {
        const float4* src = reinterpret_cast<const float4*>(Wt + (size_t)c * kKC * NP);
        float4* dst = reinterpret_cast<float4*>(wbuf + buf * kKC * NP);
        for (int i = threadIdx.x; i < CHUNK_F4; i += kColorThreads) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    prefetch(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) {
            prefetch(c + 1, (c + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* wb = wbuf + (c & 1) * kKC * NP;
        const float* a0 = A + (size_t)(ty * RPT) * lda + c * kKC;
#pragma unroll
        for (int k4 = 0; k4 < kKC / 4; ++k4) {
            float4 av[RPT];
#pragma unroll
            for (int r = 0; r < RPT; ++r) av[r] = *reinterpret_cast<const float4*>(a0 + (size_t)r * lda + k4 * 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 b0 = *reinterpret_cast<const float4*>(wb + (k4 * 4 + kk) * NP + tx * 4);
                const float4 b1 = *reinterpret_cast<const float4*>(wb + (k4 * 4 + kk) * NP + NP / 2 + tx * 4);
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const float a = kk == 0 ? av[r].x : kk == 1 ? av[r].y : kk == 2 ? av[r].z : av[r].w;
                    acc[r][0] = fmaf(a, b0.x, acc[r][0]);
                    acc[r][1] = fmaf(a, b0.y, acc[r][1]);
                    acc[r][2] = fmaf(a, b0.z, acc[r][2]);
                    acc[r][3] = fmaf(a, b0.w, acc[r][3]);
                    acc[r][4] = fmaf(a, b1.x, acc[r][4]);
                    acc[r][5] = fmaf(a, b1.y, acc[r][5]);
                    acc[r][6] = fmaf(a, b1.z, acc[r][6]);
                    acc[r][7] = fmaf(a, b1.w, acc[r][7]);
                }
            }
        }
        __syncthreads();  // wbuf[(c&1)] may be overwritten by the prefetch of chunk c+2
    }
}

template <int NP, bool RELU>
__device__ __forceinline__ void epilogue_store(const float (&acc)[NP / 32][8], const float* __restrict__ bias,
                                               float* dst, int ldd) {
    constexpr int TX = NP / 8;
    constexpr int RPT = NP / 32;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const float4 bi0 = __ldg(reinterpret_cast<const float4*>(bias + tx * 4));
    const float4 bi1 = __ldg(reinterpret_cast<const float4*>(bias + NP / 2 + tx * 4));
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        float4 o0 = make_float4(acc[r][0] + bi0.x, acc[r][1] + bi0.y, acc[r][2] + bi0.z, acc[r][3] + bi0.w);
        float4 o1 = make_float4(acc[r][4] + bi1.x, acc[r][5] + bi1.y, acc[r][6] + bi1.z, acc[r][7] + bi1.w);
        if (RELU) {
            o0.x = fmaxf(o0.x, 0.f); o0.y = fmaxf(o0.y, 0.f); o0.z = fmaxf(o0.z, 0.f); o0.w = fmaxf(o0.w, 0.f);
            o1.x = fmaxf(o1.x, 0.f); o1.y = fmaxf(o1.y, 0.f); o1.z = fmaxf(o1.z, 0.f); o1.w = fmaxf(o1.w, 0.f);
        }
        float* row = dst + (size_t)(ty * RPT + r) * ldd;
        *reinterpret_cast<float4*>(row + tx * 4) = o0;
        *reinterpret_cast<float4*>(row + NP / 2 + tx * 4) = o1;
    }
}

template <int NP>
struct ColorSmem {
    static constexpr int KA = 96;            // [h1 (64) | direnc (32)]
    static constexpr int LDA = KA + 4;       // +4 floats: bank skew
    static constexpr int LDB = NP + 4;
    static constexpr size_t bytes = sizeof(float) * ((size_t)kTileRows * LDA + (size_t)kTileRows * LDB + 2 * kKC * NP);
};

template <int NP>
__global__ void __launch_bounds__(kColorThreads, 1)
color_mlp_simt_kernel(const __grid_constant__ ColorParams p) {
    extern __shared__ __align__(16) float smem[];
    using SM = ColorSmem<NP>;
    float* actA = smem;
    float* actB = actA + kTileRows * SM::LDA;
    float* wbuf = actB + kTileRows * SM::LDB;
    constexpr int RPT = NP / 32;
    const uint32_t ntiles = div_up(p.n_rows, (uint32_t)kTileRows);
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t row0 = tile * kTileRows;
        // stage h1 tile -> actA[:, 0:64], direnc -> actA[:, 64:96]
        for (int i = threadIdx.x; i < kTileRows * 16; i += kColorThreads) {
            const int r = i >> 4, c4 = i & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < p.n_rows) v = __ldg(reinterpret_cast<const float4*>(p.h1 + (size_t)(row0 + r) * 64) + c4);
            *reinterpret_cast<float4*>(actA + r * SM::LDA + c4 * 4) = v;
        }
        for (int i = threadIdx.x; i < kTileRows * 32; i += kColorThreads) {
            const int r = i >> 5, c = i & 31;
            float val = 0.f;
            const int ndir = 3 + 6 * p.deg_view;
            if (row0 + r < p.n_rows && c < ndir) {
                const uint32_t ray = (row0 + r) / (uint32_t)p.S;
                if (c < 3) {
                    val = p.viewdirs[3 * (size_t)ray + c];
                } else {  // coord.py:L214-225 pos_enc
                    const int q = c - 3;
                    const int half = 3 * p.deg_view;
                    const int qq = q < half ? q : q - half;
                    const int deg = qq / 3, ax = qq - 3 * deg;
                    float x = fm(p.viewdirs[3 * (size_t)ray + ax], (float)(1 << deg));
                    if (q >= half) x = fa(x, 1.57079637f);
                    val = sinf(x);
                }
            }
            actA[r * SM::LDA + 64 + c] = val;
        }
        __syncthreads();
        float acc[RPT][8];
        // a = relu(P0 [h1, direnc] + c0')
#pragma unroll
        for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
        gemm_acc<NP>(acc, actA, SM::LDA, SM::KA, p.p0t, wbuf);
        epilogue_store<NP, true>(acc, p.c0, actB, SM::LDB);
        __syncthreads();
        // a2 = relu(V1a a + P1 [h1, direnc] + c1')
#pragma unroll
        for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
        gemm_acc<NP>(acc, actB, SM::LDB, NP, p.v1t, wbuf);
        gemm_acc<NP>(acc, actA, SM::LDA, SM::KA, p.v1t + (size_t)NP * NP, wbuf);
        epilogue_store<NP, true>(acc, p.c1, actB, SM::LDB);
        __syncthreads();
        // rgb layer + sigmoid + padding
        if (threadIdx.x < kTileRows * 3) {
            const int r = threadIdx.x / 3, c = threadIdx.x - 3 * r;
            if (row0 + r < p.n_rows) {
                float v = __ldg(p.r0 + c);
                const float* a = actB + r * SM::LDB;
                for (int k = 0; k < NP; ++k) v = fmaf(a[k], __ldg(p.rt + 4 * k + c), v);
                const float sgm = sigmoid_f(v);
                p.rgb[(size_t)(row0 + r) * 3 + c] = fs(fm(sgm, p.rgb_scale), p.rgb_padding);
            }
        }
        __syncthreads();
    }
}

template <int NP>
static int launch_color_t(const ColorParams& p, cudaStream_t st) {
    UC_ENSURE_SMEM(ColorSmem<NP>::bytes, color_mlp_simt_kernel<NP>);
    const uint32_t ntiles = div_up(p.n_rows, (uint32_t)kTileRows);
    const uint32_t blocks = ntiles < (uint32_t)kNumSMs ? ntiles : (uint32_t)kNumSMs;
    color_mlp_simt_kernel<NP><<<blocks, kColorThreads, ColorSmem<NP>::bytes, st>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

int launch_color_mlp_simt(const ColorParams& p, int np, cudaStream_t st) {
    if (p.n_rows == 0) return 0;
    switch (np) {
        case 32: return launch_color_t<32>(p, st);
        case 64: return launch_color_t<64>(p, st);
        case 128: return launch_color_t<128>(p, st);
        case 256: return launch_color_t<256>(p, st);
        default: set_error("color_mlp: padded width must be 32, 64, 128 or 256"); return 1;
    }
}

// =================================================================================================
// 4. composite: one warp per ray (render.compute_alpha_weights + volumetric_rendering)
// =================================================================================================
__global__ void __launch_bounds__(32 * kWarpsPerBlockRay)
composite_kernel(const CompositeParams p) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5;
    const uint32_t ray = blockIdx.x * kWarpsPerBlockRay + warp;
    if (ray >= p.n_rays) return;
    WarpExec ex{(int)(threadIdx.x & 31)};
    CompositeScratch sc;
    sc.carve(smem + (size_t)warp * CompositeScratch::floats(p.S), p.S);
    RayOutputs ro;
    const float dir[3] = {p.rays.directions[3 * (size_t)ray], p.rays.directions[3 * (size_t)ray + 1],
                          p.rays.directions[3 * (size_t)ray + 2]};
    composite_ray(ex, p.S, p.sdist + (size_t)ray * p.sdist_stride, p.density + (size_t)ray * p.S,
                  p.rgb ? p.rgb + (size_t)ray * p.S * 3 : nullptr, dir, p.rays.near[ray], p.rays.far[ray], p.bg,
                  p.extras != 0, sc, p.weights + (size_t)ray * p.S, ro);
    if (ex.lane == 0) {
        if (p.use_affine) {
            // brightness correction (models.py:L349): rgb <- A[:3,:3] rgb + A[:3,3], one affine per image at eval
            // (extrinsic_optimizer.py:L15-25 evaluated once on the host instead of once per ray)
            const float r0 = ro.rgb[0], r1 = ro.rgb[1], r2 = ro.rgb[2];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                ro.rgb[i] = fa(fmaf(p.affine[4 * i + 2], r2, fmaf(p.affine[4 * i + 1], r1, fm(p.affine[4 * i], r0))), p.affine[4 * i + 3]);
        }
        if (p.o_rgb) { p.o_rgb[3 * (size_t)ray] = ro.rgb[0]; p.o_rgb[3 * (size_t)ray + 1] = ro.rgb[1]; p.o_rgb[3 * (size_t)ray + 2] = ro.rgb[2]; }
        if (p.o_depth) p.o_depth[ray] = ro.depth;
        if (p.o_depth_raw) p.o_depth_raw[ray] = ro.depth_raw;
        if (p.o_acc) p.o_acc[ray] = ro.acc;
        if (p.o_mean) p.o_mean[ray] = ro.dist_mean;
        if (p.o_median) p.o_median[ray] = ro.dist_median;
        if (p.o_p5) p.o_p5[ray] = ro.dist_p5;
        if (p.o_p95) p.o_p95[ray] = ro.dist_p95;
        if (p.o_packed) {
            float4* o = reinterpret_cast<float4*>(p.o_packed + 12 * (size_t)ray);
            o[0] = make_float4(ro.rgb[0], ro.rgb[1], ro.rgb[2], ro.depth);
            o[1] = make_float4(ro.acc, ro.dist_mean, ro.dist_median, ro.dist_p5);
            o[2] = make_float4(ro.dist_p95, ro.depth_raw, 0.f, 0.f);
        }
    }
    // fused tile exchange: the finished row goes to every rank's image - lane k stores it to image k (this rank's own with a
    // local store, the peers' with posted stores over NVLink), so the n_peers x 48 bytes leave in one store instruction
    // per float4 instead of serially from one lane
    if (p.n_peers > 0) {
        float4 a = make_float4(ro.rgb[0], ro.rgb[1], ro.rgb[2], ro.depth);
        float4 b = make_float4(ro.acc, ro.dist_mean, ro.dist_median, ro.dist_p5);
        float4 c = make_float4(ro.dist_p95, ro.depth_raw, 0.f, 0.f);
        a.x = __shfl_sync(0xffffffffu, a.x, 0); a.y = __shfl_sync(0xffffffffu, a.y, 0); a.z = __shfl_sync(0xffffffffu, a.z, 0);
        a.w = __shfl_sync(0xffffffffu, a.w, 0);
        b.x = __shfl_sync(0xffffffffu, b.x, 0); b.y = __shfl_sync(0xffffffffu, b.y, 0); b.z = __shfl_sync(0xffffffffu, b.z, 0);
        b.w = __shfl_sync(0xffffffffu, b.w, 0);
        c.x = __shfl_sync(0xffffffffu, c.x, 0); c.y = __shfl_sync(0xffffffffu, c.y, 0);
        if (ex.lane < p.n_peers) {
            float4* o = reinterpret_cast<float4*>(p.peer_packed[ex.lane] + 12 * ((size_t)p.peer_row0 + ray));
            o[0] = a; o[1] = b; o[2] = c;
        }
    }
}

int launch_composite(const CompositeParams& p, cudaStream_t st) {
    if (p.n_rays == 0) return 0;
    const size_t smem = kWarpsPerBlockRay * CompositeScratch::floats(p.S) * sizeof(float);
    UC_REQUIRE(smem <= 227 * 1024, "composite: too many samples per ray for shared memory");
    UC_ENSURE_SMEM(smem, composite_kernel);
    composite_kernel<<<div_up(p.n_rays, (uint32_t)kWarpsPerBlockRay), 32 * kWarpsPerBlockRay, smem, st>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

}  // namespace ucnerf

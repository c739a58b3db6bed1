// sample + encode + density MLP: one thread per ray-sample (the hash-gather kernel of the path).
//   render.cast_rays -> coord.contract -> GridEncoder (6 points x L levels x 8 corners) -> erf pooling ->
//   density_layer -> softplus          (models.py:L208-230, L485-512, L581; gridencoder.cu:L87-197)
//
// B200 notes (profiles/r1_summary.md): the proposal level is bound by instruction issue / the FP32 pipe (about 115
// instructions per point-level, 8 of them loads), the NeRF level by L1 wavefronts (distinct 128-byte lines per gather
// instruction) and L1-miss latency; DRAM traffic is far below the algorithmic gather bytes because the proposal table
// (101 MB) stays L2-resident and neighbouring rays share cells on the coarse levels.  Hence:
//   * per-level code is specialised at compile time (dense index without modulo vs. XOR-prime hash + mask),
//     corner indices are built from shared partial terms, all 8 gathers of a point-level are issued before use;
//   * per-level constants come from the constant bank (__grid_constant__ params), weights of the 24/40 -> 64
//     layer from shared memory as broadcast LDS.128;
//   * a warp = 32 neighbouring rays at the SAME sample index (see sample_pos), so the lanes of one gather fall into
//     a few lines on most levels.
#include "ray_march.cuh"

namespace ucnerf {

constexpr int kSampleThreads = 128;
// resident CTAs per SM the register allocator is asked to allow (profiles/: occupancy vs. spills trade-off)
#ifndef UC_MINB_SMALL
#define UC_MINB_SMALL 5   // LMAX <= 6  -> 96 registers (6 -> 80 registers spills in the MLP phase)
#endif
#ifndef UC_MMA_ONE_PASS_MINB
#define UC_MMA_ONE_PASS_MINB 4
#endif
#ifndef UC_H1_STORE_256
#define UC_H1_STORE_256 1   // 32-byte stores of h1 (measured: NeRF kernel 11.27 -> 10.97 ms)
#endif
#ifndef UC_REMAP_PROP
#define UC_REMAP_PROP 1
#endif
#ifndef UC_MINB_LARGE
#define UC_MINB_LARGE 4   // LMAX <= 10 -> 128 registers (5 -> 96 registers spills with FFMA2 register pairs; measured slower)
#endif

// entry address = level base + 16 * index as ONE IMAD.WIDE.U32 (the compiler otherwise re-associates the level offset
// into a 64-bit add + LEA pair per corner: 4 instructions per address in an instruction-issue-bound kernel)
__device__ __forceinline__ const float4* entry_ptr(uint64_t tab, uint32_t idx) {
    uint64_t a;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(idx), "l"(tab));
    return reinterpret_cast<const float4*>(a);
}

// 1/sqrt(x) as the single MUFU.RSQ that rsqrtf is built on, without its denormal-input rescaling (3 extra issue slots
// per point-level): the argument 8 sigma^2 G^2 is >= 1e-12 for any representable cone, and a flushed denormal gives
// +inf -> erf weight exactly 1, the same value the rescaled form leads to.
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// x rounded to TF32 (10-bit mantissa, round to nearest) as an fp32 value; x - tf32_hi(x) is exact in fp32
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// fp32 -> bf16 (round to nearest even) bit pattern; finite inputs
__device__ __forceinline__ uint32_t bf16_bits(float x) {
    const uint32_t b = __float_as_uint(x);
    return (b + 0x7fffu + ((b >> 16) & 1u)) >> 16;
}
// D += A B, m16n8k8, A row-major [16][8], B column-major [8][8] (legacy warp-level tensor-core path: at 64 outputs and
// K = 24 / 40 per sample the layer is far too small for a tcgen05 tile pipeline to pay its TMEM round trip)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// MODE: 0 dense (index < table size, no reduction), 1 hashed with power-of-two table, 2 generic (runtime flags)
template <int MODE>
__device__ __forceinline__ void gather8(const GridLevel& lv, const float4* __restrict__ tab_, const CellCoords& c,
                                        float4 (&v)[8]) {
    const uint64_t tab = reinterpret_cast<uint64_t>(tab_);
    if constexpr (MODE == 0) {
        const uint32_t b00 = c.ix + c.iy * lv.stride1 + c.iz * lv.stride2;
        const uint32_t b10 = b00 + lv.stride1, b01 = b00 + lv.stride2, b11 = b10 + lv.stride2;
        const float4 *p00 = entry_ptr(tab, b00), *p10 = entry_ptr(tab, b10), *p01 = entry_ptr(tab, b01), *p11 = entry_ptr(tab, b11);
        v[0] = ldg_f4(p00); v[1] = ldg_f4(p00 + 1);
        v[2] = ldg_f4(p10); v[3] = ldg_f4(p10 + 1);
        v[4] = ldg_f4(p01); v[5] = ldg_f4(p01 + 1);
        v[6] = ldg_f4(p11); v[7] = ldg_f4(p11 + 1);
    } else if constexpr (MODE == 1) {
        const uint32_t hy0 = c.iy * 2654435761u, hy1 = hy0 + 2654435761u;
        const uint32_t hz0 = c.iz * 805459861u, hz1 = hz0 + 805459861u;
        const uint32_t a00 = hy0 ^ hz0, a10 = hy1 ^ hz0, a01 = hy0 ^ hz1, a11 = hy1 ^ hz1;
        const uint32_t x0 = c.ix, x1 = c.ix + 1, m = lv.pow2_mask;
        v[0] = ldg_f4(entry_ptr(tab, (x0 ^ a00) & m)); v[1] = ldg_f4(entry_ptr(tab, (x1 ^ a00) & m));
        v[2] = ldg_f4(entry_ptr(tab, (x0 ^ a10) & m)); v[3] = ldg_f4(entry_ptr(tab, (x1 ^ a10) & m));
        v[4] = ldg_f4(entry_ptr(tab, (x0 ^ a01) & m)); v[5] = ldg_f4(entry_ptr(tab, (x1 ^ a01) & m));
        v[6] = ldg_f4(entry_ptr(tab, (x0 ^ a11) & m)); v[7] = ldg_f4(entry_ptr(tab, (x1 ^ a11) & m));
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            v[k] = ldg_f4(entry_ptr(tab, level_index(lv, c.ix + (k & 1), c.iy + ((k >> 1) & 1), c.iz + ((k >> 2) & 1))));
    }
}

// trilinear weights in the reference's association ((wx * wy) * wz), corner k: bit0 -> x, bit1 -> y, bit2 -> z.
// Packed fp32x2 math (FFMA2/FMUL2): every weight is computed as a duplicated pair {w, w} so the four channels of a
// corner take two fused multiply-adds; element-wise results are bit-identical to the scalar sequence.
struct Feat4 {
    float2 xy, zw;
};
__device__ __forceinline__ Feat4 interp8(const CellCoords& c, const float4 (&v)[8]) {
    const float2 fx = make_float2(c.fx, c.fx), fy = make_float2(c.fy, c.fy), fz = make_float2(c.fz, c.fz);
    const float2 gx = make_float2(1 - c.fx, 1 - c.fx), gy = make_float2(1 - c.fy, 1 - c.fy), gz = make_float2(1 - c.fz, 1 - c.fz);
    const float2 w00 = fmul2(gx, gy), w10 = fmul2(fx, gy), w01 = fmul2(gx, fy), w11 = fmul2(fx, fy);
    const float2 w[8] = {fmul2(w00, gz), fmul2(w10, gz), fmul2(w01, gz), fmul2(w11, gz),
                         fmul2(w00, fz), fmul2(w10, fz), fmul2(w01, fz), fmul2(w11, fz)};
    Feat4 r;
    r.xy = make_float2(0.f, 0.f);
    r.zw = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        r.xy = ffma2(w[k], make_float2(v[k].x, v[k].y), r.xy);
        r.zw = ffma2(w[k], make_float2(v[k].z, v[k].w), r.zw);
    }
    return r;
}

// Work mapping: a warp = 32 CONSECUTIVE RAYS at the SAME sample index (position q -> ray group q / (32 S), sample
// (q / 32) % S, ray-in-group q % 32).  Neighbouring pixels' rays are a fraction of a cell apart on most levels, so
// the 32 lanes of one gather instruction fall into a handful of 128-byte lines instead of 32 different ones (the
// kernel is bound by L1 wavefronts = distinct lines per instruction).  For unordered ray batches this is merely
// as slow as any other mapping.  Row index of (ray, s) in the [N*S] buffers stays ray * S + s.
struct SamplePos {
    uint32_t ray;
    int s;
    bool valid;
};
// The mapping for the positions q = block0 + i of ONE CTA, with the 64-bit division done once.
// rw_log2 = 5: lane -> ray, warp -> sample.  rw_log2 < 5 (experiment, option "warp_rays"): a warp covers 2^rw_log2
// rays x 2^(5 - rw_log2) consecutive samples; the warps of a group enumerate (sample block, ray block).
//
// tile_w > 0 (the rays are whole rows of a row-major image of that width, SampleParams::tile_w): the 32 rays of a group
// are a patch of pw x 32 / pw pixels (lane -> dx = lane % pw, dy = lane / pw; pw = 4, 8 or 16) instead of 32 pixels of one row.  Same rays, same
// results, another assignment of rays to warps: the lanes of a gather are closer together in the scene, so they share
// more cells and 128-byte lines on the middle levels (measured with permuted ray batches in round 1:
// profiles/r1_summary.md; here without moving any data).  Bands of 8 rows that do not fit keep the row mapping.
struct BlockSamples {
    uint32_t group, rem0, span, n_rays, tile_w, pw_log2, tiles_per_band, tiled_groups;
    int S, rw_log2;
    __device__ __forceinline__ BlockSamples(size_t block0, uint32_t n_rays_, int S_, int rw_log2_, uint32_t tile_w_, uint32_t pw_log2_)
        : span(32u * (uint32_t)S_), n_rays(n_rays_), tile_w(rw_log2_ == 5 ? tile_w_ : 0u), pw_log2(pw_log2_), S(S_), rw_log2(rw_log2_) {
        const size_t g = block0 / span;
        group = (uint32_t)g;
        rem0 = (uint32_t)(block0 - g * span);
        tiles_per_band = tile_w >> pw_log2;
        const uint32_t band_rays = (32u >> pw_log2) * tile_w;       // rows of a band x width
        tiled_groups = tile_w ? (n_rays / band_rays) * tiles_per_band : 0u;
    }
    __device__ __forceinline__ SamplePos at(uint32_t i) const {
        uint32_t rem = rem0 + i, grp = group;
        while (rem >= span) { rem -= span; ++grp; }
        const uint32_t lane = rem & 31u, w = rem >> 5, sw_log2 = 5u - (uint32_t)rw_log2;
        const uint32_t ray_block = w & ((1u << sw_log2) - 1u), sample_block = w >> sw_log2;
        SamplePos sp;
        if (grp < tiled_groups) {
            const uint32_t band = grp / tiles_per_band, col = grp - band * tiles_per_band;
            sp.ray = ((32u >> pw_log2) * band + (lane >> pw_log2)) * tile_w + (col << pw_log2) + (lane & ((1u << pw_log2) - 1u));
        } else {
            sp.ray = grp * 32u + (ray_block << rw_log2) + (lane & ((1u << rw_log2) - 1u));
        }
        sp.s = (int)((sample_block << sw_log2) + (lane >> rw_log2));
        sp.valid = sp.ray < n_rays;
        return sp;
    }
};

// The basis of a ray's cones (render.py:L139-146: two cross products and normalisations) is the same for all its samples:
// ray_geom_kernel evaluates make_ray_geom once per ray into a [15][ld] array that the sample threads read coalesced
// (same arithmetic, same bits; 15 strided loads and about 100 instructions less per sample).
__global__ void __launch_bounds__(256) ray_geom_kernel(RayPtrs r, uint32_t n_rays, float* __restrict__ geom, uint32_t ld) {
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    RayGeom g;
    make_ray_geom(g, r.origins + 3 * (size_t)ray, r.directions + 3 * (size_t)ray, r.cam_dirs + 3 * (size_t)ray,
                  r.rand_vec + 3 * (size_t)ray, r.radii[ray], r.near[ray], r.far[ray]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        geom[(size_t)i * ld + ray] = g.o[i];
        geom[(size_t)(3 + i) * ld + ray] = g.d[i];
        geom[(size_t)(6 + i) * ld + ray] = g.e1[i];
        geom[(size_t)(9 + i) * ld + ray] = g.e2[i];
    }
    geom[(size_t)12 * ld + ray] = g.radius;
    geom[(size_t)13 * ld + ray] = g.near;
    geom[(size_t)14 * ld + ray] = g.far;
}

int launch_ray_geom(const RayPtrs& rays, uint32_t n_rays, float* geom, uint32_t geom_ld, cudaStream_t st) {
    if (n_rays == 0) return 0;
    ray_geom_kernel<<<div_up(n_rays, 256u), 256, 0, st>>>(rays, n_rays, geom, geom_ld);
    UC_LAUNCH_CHECK();
    return 0;
}

__device__ __forceinline__ void load_ray_geom(RayGeom& rg, const SampleParams& p, uint32_t ray) {
    if (p.geom) {
        const float* g = p.geom + ray;
        const size_t ld = p.geom_ld;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            rg.o[i] = g[(size_t)i * ld];
            rg.d[i] = g[(size_t)(3 + i) * ld];
            rg.e1[i] = g[(size_t)(6 + i) * ld];
            rg.e2[i] = g[(size_t)(9 + i) * ld];
        }
        rg.radius = g[(size_t)12 * ld];
        rg.near = g[(size_t)13 * ld];
        rg.far = g[(size_t)14 * ld];
    } else {
        make_ray_geom(rg, p.rays.origins + 3 * (size_t)ray, p.rays.directions + 3 * (size_t)ray,
                      p.rays.cam_dirs + 3 * (size_t)ray, p.rays.rand_vec + 3 * (size_t)ray, p.rays.radii[ray],
                      p.rays.near[ray], p.rays.far[ray]);
    }
}

// ND >= 0: levels [0, ND) are dense, levels >= ND hashed with power-of-two tables (compile-time specialisation);
// ND < 0: decide per level at run time.
//
// Two phases per 128-sample CTA.  (1) gather: thread = sample, pooled features F[L*4] in registers, then parked in
// shared memory.  (2) density layer: thread = (4 samples, 16 hidden units), so every weight fetched from shared
// memory feeds 16 FMAs instead of 4 - the kernel is LSU-wavefront bound and the weight loads of a
// thread-per-sample MLP were a quarter of all wavefronts.  Mappings are chosen bank-conflict free:
// thread t: hg = t % 4 -> hidden units {hg + 4 jj}, sg = t / 4 -> samples {sg + 32 s}.  h1 is an internal buffer, so
// it is stored in that permuted column order (column 16 hg + jj <-> hidden unit hg + 4 jj, see kH1Perm in
// ray_march.cuh) and the consumer's weight rows are permuted to match on the host.
//
// RUNS = true ("cell-run reuse"): the six multisample points of an interval are ordered along the ray and, on the
// coarse levels, mostly fall into the SAME grid cell (measured on the waymo.gin frame: 93 / 86 / 75 / 63 / 48 / 30 %
// of the proposal samples have all six points in one cell on levels 0..5; 9.9 distinct cells per sample instead of
// 36 point-levels).  The loop nest is therefore level-outer / point-inner: the six grid coordinates stay in
// registers, and the 8 corner entries of a level are re-gathered only when a point enters a different cell than its
// predecessor.  Interpolation, erf pooling and the summation order over the six points are unchanged, so results
// are bit-identical to the RUNS = false form; only redundant gathers (L1 wavefronts, the limiter) disappear.  The
// level loop is rolled (per-level constants by dynamic constant-bank index), pooled features go straight to the
// shared-memory rows the density layer reads.
//
// MLP = 2 (default where it pays): the density layer runs on the tensor cores.  A warp multiplies its own 32 feature
// rows (two m16 tiles) with the 24/40 x 64 weight as mma.sync.m16n8k8 TF32 with the 3-term split (x = hi + lo,
// lo*hi + hi*lo + hi*hi accumulated in fp32: products exact to about 2^-20, no scaling and no range limit, unlike an
// FP16 split).  Features never leave the warp: each k-tile (two levels) goes from the registers through a 1.5 KB
// warp-private transpose buffer into A fragments.  Weights sit in shared memory already split and in fragment order:
// per (unit, k-tile, t) a float2 {hi[k = t], hi[k = t + 4]} and one word of two bf16 {lo[t], lo[t + 4]} (a bf16 is an
// exact TF32 operand, and 8 mantissa bits of the 2^-11-sized rest keep the product at 2^-20).  Row strides are padded so
// that the 64-bit loads of a half warp and the 32-bit loads of a warp are bank-conflict free.  Compared with the FFMA2
// forms (MLP = 1 / 0, kept as options) the layer's shared-memory wavefronts drop from about 17 to 4 per proposal sample,
// its 3 K FMAs leave the FP32 pipe, and - measured to matter as much - the CTA needs 16 KB instead of 22 KB of shared
// memory, which leaves the L1 more room for the gathers (profiles/r2_mma_density.md).
template <int LMAX>
struct MmaLayout {
    static constexpr int KT = LMAX / 2;                                                                // k-tiles of 8 feature columns
    static constexpr int SH = KT * 8 + (((KT * 8) % 32 == 8 || (KT * 8) % 32 == 24) ? 0 : 8);          // hi row stride (words) = 8 (mod 16)
    static constexpr int SL = KT * 4 + (((KT * 4) % 8 == 4) ? 0 : 4);                                  // lo row stride (words) = 4 (mod 8)
    static constexpr int SA = 12;                                                                      // transpose buffer row stride (words)
    static constexpr size_t smem_bytes = sizeof(float) * (64 * (SH + SL) + 128 + (kSampleThreads / 32) * 32 * SA);
};

template <int LMAX, bool NERF, int ND, int MINB, int MLP, bool RUNS>
__global__ void __launch_bounds__(kSampleThreads, MINB)
sample_encode_kernel(const __grid_constant__ SampleParams p) {
    constexpr bool REMAP = MLP == 1, MMA = MLP == 2;
    static_assert(!(MMA && RUNS), "the tensor-core density layer takes the features from registers");
    static_assert(LMAX % 2 == 0, "k-tiles of two levels");
    using ML = MmaLayout<LMAX>;
    constexpr int LC = LMAX * 4;
    constexpr int LDS = LC + 4;  // row stride (floats): 16-byte slots of 8 consecutive rows fall in distinct banks
    extern __shared__ __align__(16) float smem_dyn[];
    float* sW1 = smem_dyn;                                   // [64][LDS]   (MMA: hi pairs [64][SH], then lo words [64][SL])
    float* sB1 = sW1 + 64 * (MMA ? ML::SH + ML::SL : LDS);   // [64]
    float* sW2 = sB1 + 64;                                   // [64]
    float* sF = sW2 + 64;                                    // [128][LDS] for the re-mapped phase / RUNS; MMA: [4 warps][32][SA]
    (void)sF;
    if constexpr (MMA) {
        uint32_t* sWl = reinterpret_cast<uint32_t*>(sW1 + 64 * ML::SH);
        for (int i = threadIdx.x; i < 64 * ML::KT * 4; i += kSampleThreads) {
            const int n = i / (ML::KT * 4), r = i - n * (ML::KT * 4), kt = r >> 2, t = r & 3;
            const float wa = __ldg(p.w1p + n * LC + 8 * kt + t), wb = __ldg(p.w1p + n * LC + 8 * kt + t + 4);
            const float ha = tf32_hi(wa), hb = tf32_hi(wb);
            *reinterpret_cast<float2*>(sW1 + n * ML::SH + 8 * kt + 2 * t) = make_float2(ha, hb);
            sWl[n * ML::SL + 4 * kt + t] = bf16_bits(wa - ha) | (bf16_bits(wb - hb) << 16);
        }
    } else {
        for (int i = threadIdx.x; i < 64 * LMAX; i += kSampleThreads) {
            const int j = i / LMAX, k4 = i - j * LMAX;
            *reinterpret_cast<float4*>(sW1 + j * LDS + 4 * k4) = __ldg(reinterpret_cast<const float4*>(p.w1p) + i);
        }
    }
    if (threadIdx.x < 64) {
        sB1[threadIdx.x] = p.b1[threadIdx.x];
        sW2[threadIdx.x] = p.w2[threadIdx.x];
    }

    if constexpr (MMA) __syncthreads();  // weights staged; the warps run independently from here on
    const size_t block0 = (size_t)blockIdx.x * kSampleThreads;
    const BlockSamples samples(block0, p.n_rays, p.S, p.rw_log2, p.tile_w, p.tile_pw_log2);
    const SamplePos me = samples.at(threadIdx.x);
    const size_t idx = (size_t)me.ray * p.S + me.s;  // row of this sample in the [N*S] buffers
    float2 F2[LMAX * 2];  // pooled features, (x,y) / (z,w) pairs per level
#pragma unroll
    for (int i = 0; i < LMAX * 2; ++i) F2[i] = make_float2(0.f, 0.f);

    if constexpr (RUNS) {
        float* myF = sF + threadIdx.x * LDS;
        int L = 0;
        if (me.valid) {
            const uint32_t ray = me.ray;
            const int s = me.s;
            float gq[6][3], s8q[6];
            uint32_t inmask = 0;
            {
                RayGeom rg;
                load_ray_geom(rg, p, ray);
                const float s0 = p.sdist[(size_t)ray * p.sdist_stride + s];
                const float s1 = p.sdist[(size_t)ray * p.sdist_stride + s + 1];
                const float t0 = fa(fm(s0, rg.far), fm(fs(1.f, s0), rg.near));
                const float t1 = fa(fm(s1, rg.far), fm(fs(1.f, s1), rg.near));
                const ConeInterval ci = make_cone_interval(t0, t1);
                const int odd = s & 1;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    float sigma;
                    cone_point(rg, ci, p.cone, j, odd, p.std_scale, gq[j], sigma);
                    s8q[j] = fm(8.f, fm(sigma, sigma));
                    // gridencoder.cu:L110-135: out-of-range input -> zero features for every level
                    const bool oob = gq[j][0] < 0.f || gq[j][0] > 1.f || gq[j][1] < 0.f || gq[j][1] > 1.f || gq[j][2] < 0.f || gq[j][2] > 1.f;
                    if (!oob) inmask |= 1u << j;
                }
            }
            L = p.grid.num_levels;
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                const GridLevel& lv = p.grid.lv[l];
                const float4* tab = p.grid.table + lv.offset;
                const float g2l = p.g2[l];
                float4 v[8];
                uint32_t pix = 0xffffffffu, piy = 0, piz = 0;   // no cell has ix = 2^32 - 1: the first in-range point always gathers
                float2 Fxy = make_float2(0.f, 0.f), Fzw = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    if ((inmask >> j) & 1u) {
                        const CellCoords c = cell_of(lv, gq[j]);
                        if (c.ix != pix || c.iy != piy || c.iz != piz) {
                            if (!lv.hashed && lv.mod_mode == 0) gather8<0>(lv, tab, c, v);
                            else if (lv.hashed && lv.mod_mode == 1) gather8<1>(lv, tab, c, v);
                            else gather8<2>(lv, tab, c, v);
                            pix = c.ix; piy = c.iy; piz = c.iz;
                        }
                        const Feat4 r = interp8(c, v);
                        const float ea = rsqrt_fast(fm(s8q[j], g2l));       // models.py:L495, see the RUNS = false form below
                        const float om = ea >= 4.f ? 1.f : erff(ea);
                        const float2 om2 = make_float2(om, om);
                        Fxy = ffma2(om2, r.xy, Fxy);
                        Fzw = ffma2(om2, r.zw, Fzw);
                    }
                }
                // .mean(dim=-3) over the 6 points, models.py:L496
                *reinterpret_cast<float4*>(myF + 4 * l) =
                    make_float4(Fxy.x * 0.16666667f, Fxy.y * 0.16666667f, Fzw.x * 0.16666667f, Fzw.y * 0.16666667f);
            }
        }
        for (int l = L; l < LMAX; ++l) *reinterpret_cast<float4*>(myF + 4 * l) = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (MLP == 0) {
#pragma unroll
            for (int l = 0; l < LMAX; ++l) {
                const float4 f = *reinterpret_cast<const float4*>(myF + 4 * l);
                F2[2 * l] = make_float2(f.x, f.y);   // own row: already the mean over the six points
                F2[2 * l + 1] = make_float2(f.z, f.w);
            }
        }
    } else
    if (me.valid) {
        const uint32_t ray = me.ray;
        const int s = me.s;
        RayGeom rg;
        load_ray_geom(rg, p, ray);
        const float s0 = p.sdist[(size_t)ray * p.sdist_stride + s];
        const float s1 = p.sdist[(size_t)ray * p.sdist_stride + s + 1];
        const float t0 = fa(fm(s0, rg.far), fm(fs(1.f, s0), rg.near));
        const float t1 = fa(fm(s1, rg.far), fm(fs(1.f, s1), rg.near));
        const ConeInterval ci = make_cone_interval(t0, t1);
        const int odd = s & 1;
        const int L = p.grid.num_levels;
#pragma unroll 1
        for (int j = 0; j < 6; ++j) {
            float g[3], sigma;
            cone_point(rg, ci, p.cone, j, odd, p.std_scale, g, sigma);
            // gridencoder.cu:L110-135: out-of-range input -> zero features for every level
            if (g[0] < 0.f || g[0] > 1.f || g[1] < 0.f || g[1] > 1.f || g[2] < 0.f || g[2] > 1.f) continue;
            const float s8 = fm(8.f, fm(sigma, sigma));
#pragma unroll
            for (int l = 0; l < LMAX; ++l) {
                if (l < L) {
                    const GridLevel& lv = p.grid.lv[l];
                    const CellCoords c = cell_of(lv, g);
                    const float4* tab = p.grid.table + lv.offset;
                    float4 v[8];
                    if constexpr (ND >= 0) {
                        if (l < ND) gather8<0>(lv, tab, c, v);
                        else gather8<1>(lv, tab, c, v);
                    } else {
                        gather8<2>(lv, tab, c, v);
                    }
                    const Feat4 r = interp8(c, v);
                    // models.py:L495 scale-aware down-weighting erf(1/sqrt(8 std^2 G^2));
                    // erf(x) rounds to exactly 1.0f for x >= 4, so coarse levels skip the evaluation
                    const float ea = rsqrt_fast(fm(s8, p.g2[l]));
                    const float om = ea >= 4.f ? 1.f : erff(ea);
                    const float2 om2 = make_float2(om, om);
                    F2[2 * l] = ffma2(om2, r.xy, F2[2 * l]);
                    F2[2 * l + 1] = ffma2(om2, r.zw, F2[2 * l + 1]);
                }
            }
        }
    }
    if constexpr (MMA) {
        // density_layer: Linear(L*C,64) -> ReLU -> Linear(64, .)[0]   (models.py:L438-441, L507-508)
        constexpr int KT = ML::KT, SH = ML::SH, SL = ML::SL, SA = ML::SA;
        // hidden units per pass: all 64 (64 accumulator registers) where the register budget is 128, else 2 x 32
        constexpr int NHALF = (MINB <= UC_MMA_ONE_PASS_MINB) ? 1 : 2, NT = 8 / NHALF;
        const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
        float* sAw = sF + wq * (32 * SA);                        // this warp's transpose buffer [32 rows][8 columns]
        const float* Ar = sAw + g * SA + t;                      // A fragment element (row g + 8 i, column t + 4 j)
        const float2* Wh = reinterpret_cast<const float2*>(sW1 + g * SH) + t;                     // unit 8 nt + g: {hi[t], hi[t + 4]}
        const uint32_t* Wl = reinterpret_cast<const uint32_t*>(sW1 + 64 * SH) + g * SL + t;       // {lo[t], lo[t + 4]} as bf16 x 2
        const float2 sixth = make_float2(0.16666667f, 0.16666667f);
#pragma unroll
        for (int i = 0; i < LMAX * 2; ++i) F2[i] = fmul2(F2[i], sixth);  // .mean(dim=-3) over the 6 points, models.py:L496
        float raw[4] = {0.f, 0.f, 0.f, 0.f};                     // rows g, g + 8, 16 + g, 24 + g of the warp's block
#pragma unroll 1
        for (int half = 0; half < NHALF; ++half) {
            float acc[2][NT][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[mt][n][e] = 0.f;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                __syncwarp();                                    // the previous k-tile has been read
                *reinterpret_cast<float4*>(sAw + lane * SA) = make_float4(F2[4 * kt].x, F2[4 * kt].y, F2[4 * kt + 1].x, F2[4 * kt + 1].y);
                *reinterpret_cast<float4*>(sAw + lane * SA + 4) =
                    make_float4(F2[4 * kt + 2].x, F2[4 * kt + 2].y, F2[4 * kt + 3].x, F2[4 * kt + 3].y);
                __syncwarp();
                uint32_t ah[2][4], al[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // the tensor core reads the upper 19 bits of an operand: hi = those bits, lo = the exact rest
                        const float x = Ar[(16 * mt + 8 * (e & 1)) * SA + 4 * (e >> 1)];
                        ah[mt][e] = __float_as_uint(x) & 0xffffe000u;
                        al[mt][e] = __float_as_uint(x - __uint_as_float(ah[mt][e]));
                    }
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const int nt = NT * half + n;
                    const float2 wh = Wh[(8 * nt * SH + 8 * kt) / 2];
                    const uint32_t wl = Wl[8 * nt * SL + 4 * kt];
                    const uint32_t bh0 = __float_as_uint(wh.x), bh1 = __float_as_uint(wh.y);
                    const uint32_t bl0 = wl << 16, bl1 = wl & 0xffff0000u;
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {             // small terms first
                        mma_tf32(acc[mt][n], al[mt], bh0, bh1);
                        mma_tf32(acc[mt][n], ah[mt], bl0, bl1);
                        mma_tf32(acc[mt][n], ah[mt], bh0, bh1);
                    }
                }
            }
            // C fragment: acc[mt][n][2 rh + e] = (row 16 mt + g + 8 rh, unit 8 (NT half + n) + 2 t + e)
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int u = 8 * (NT * half + n) + 2 * t;
                const float2 b = *reinterpret_cast<const float2*>(sB1 + u), w2 = *reinterpret_cast<const float2*>(sW2 + u);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
                        const float a0 = fmaxf(acc[mt][n][2 * rh] + b.x, 0.f), a1 = fmaxf(acc[mt][n][2 * rh + 1] + b.y, 0.f);
                        acc[mt][n][2 * rh] = a0;
                        acc[mt][n][2 * rh + 1] = a1;
                        raw[2 * mt + rh] = fmaf(w2.y, a1, fmaf(w2.x, a0, raw[2 * mt + rh]));
                    }
            }
            if constexpr (NERF) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
                        const SamplePos o = samples.at(32 * wq + 16 * mt + 8 * rh + g);
                        if (o.valid) {  // columns 32 (nt / 4) + 8 t + 2 (nt % 4) + e (h1_col), nt = NT half + n
                            float* hrow = p.h1 + ((size_t)o.ray * p.S + o.s) * 64 + 8 * t + 8 * NT * half;
#if UC_H1_STORE_256
                            // 32-byte stores (sm_100): half the store instructions for the same lines
#pragma unroll
                            for (int n = 0; n < NT; n += 4)
                                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(hrow + 8 * n),
                                             "f"(acc[mt][n][2 * rh]), "f"(acc[mt][n][2 * rh + 1]), "f"(acc[mt][n + 1][2 * rh]),
                                             "f"(acc[mt][n + 1][2 * rh + 1]), "f"(acc[mt][n + 2][2 * rh]), "f"(acc[mt][n + 2][2 * rh + 1]),
                                             "f"(acc[mt][n + 3][2 * rh]), "f"(acc[mt][n + 3][2 * rh + 1])
                                             : "memory");
#else
#pragma unroll
                            for (int n = 0; n < NT; n += 2)
                                *reinterpret_cast<float4*>(hrow + 32 * (n / 4) + 2 * (n % 4)) = make_float4(
                                    acc[mt][n][2 * rh], acc[mt][n][2 * rh + 1], acc[mt][n + 1][2 * rh], acc[mt][n + 1][2 * rh + 1]);
#endif
                        }
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            raw[i] += __shfl_xor_sync(0xffffffffu, raw[i], 1);
            raw[i] += __shfl_xor_sync(0xffffffffu, raw[i], 2);
        }
        // lane t of the quad writes the quad's row 16 (t / 2) + 8 (t % 2) + g
        const float mine = t == 0 ? raw[0] : t == 1 ? raw[1] : t == 2 ? raw[2] : raw[3];
        const SamplePos o = samples.at(32 * wq + 16 * (t >> 1) + 8 * (t & 1) + g);
        if (o.valid) p.density[(size_t)o.ray * p.S + o.s] = softplus_f(mine + p.b2 + p.density_bias);  // models.py:L581
        return;
    }
    if constexpr (MLP == 0) {
        // thread-per-sample density layer (weights as broadcast LDS.128); used where the re-mapped phase does not pay
        __syncthreads();  // weights staged
        if (!me.valid) return;
        if constexpr (!RUNS) {
            const float2 sixth = make_float2(0.16666667f, 0.16666667f);
#pragma unroll
            for (int i = 0; i < LMAX * 2; ++i) F2[i] = fmul2(F2[i], sixth);  // .mean(dim=-3) over the 6 points, models.py:L496
        }
        float raw = p.b2;
        float* hrow = NERF ? p.h1 + idx * 64 : nullptr;
#pragma unroll 4
        for (int c = 0; c < 64; c += 4) {  // h1 column c holds hidden unit h1_perm(c) (same layout as the re-mapped path)
            float hv[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int j = h1_perm(c + cc);
                float2 a2 = make_float2(sB1[j], 0.f);  // even / odd partial sums (FFMA2)
                const float4* wr = reinterpret_cast<const float4*>(sW1 + j * LDS);
#pragma unroll
                for (int l = 0; l < LMAX; ++l) {
                    const float4 w = wr[l];
                    a2 = ffma2(make_float2(w.x, w.y), F2[2 * l], a2);
                    a2 = ffma2(make_float2(w.z, w.w), F2[2 * l + 1], a2);
                }
                const float a = fmaxf(a2.x + a2.y, 0.f);
                raw = fmaf(sW2[j], a, raw);
                hv[cc] = a;
            }
            if (NERF) *reinterpret_cast<float4*>(hrow + c) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        }
        p.density[idx] = softplus_f(raw + p.density_bias);  // models.py:L581
        return;
    }
    // .mean(dim=-3) over the 6 points (models.py:L496), parked in shared memory for the re-mapped MLP phase
    if constexpr (!RUNS) {
#pragma unroll
        for (int l = 0; l < LMAX; ++l)
            *reinterpret_cast<float4*>(sF + threadIdx.x * LDS + 4 * l) =
                make_float4(F2[2 * l].x * 0.16666667f, F2[2 * l].y * 0.16666667f, F2[2 * l + 1].x * 0.16666667f,
                            F2[2 * l + 1].y * 0.16666667f);
    }
    __syncthreads();

    // density_layer: Linear(L*C,64) -> ReLU -> Linear(64, .)[0]   (models.py:L438-441, L507-508)
    const int hg = threadIdx.x & 3, sg = threadIdx.x >> 2;
    float raw[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {  // 2 x 8 hidden units per thread keeps the accumulators at 32 registers
        float2 acc2[4][8];  // (even-k, odd-k) partial sums per (sample, hidden unit): FFMA2
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const float b = sB1[hg + 4 * (8 * half + jj)];
#pragma unroll
            for (int s = 0; s < 4; ++s) acc2[s][jj] = make_float2(b, 0.f);
        }
#pragma unroll 2
        for (int k4 = 0; k4 < LMAX; ++k4) {
            float4 f[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) f[s] = *reinterpret_cast<const float4*>(sF + (sg + 32 * s) * LDS + 4 * k4);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float4 w = *reinterpret_cast<const float4*>(sW1 + (hg + 4 * (8 * half + jj)) * LDS + 4 * k4);
                const float2 wxy = make_float2(w.x, w.y), wzw = make_float2(w.z, w.w);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    acc2[s][jj] = ffma2(wxy, make_float2(f[s].x, f[s].y), acc2[s][jj]);
                    acc2[s][jj] = ffma2(wzw, make_float2(f[s].z, f[s].w), acc2[s][jj]);
                }
            }
        }
        float acc[4][8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const float w2 = sW2[hg + 4 * (8 * half + jj)];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                acc[s][jj] = fmaxf(acc2[s][jj].x + acc2[s][jj].y, 0.f);
                raw[s] = fmaf(w2, acc[s][jj], raw[s]);
            }
        }
        if (NERF) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const SamplePos o = samples.at(sg + 32 * s);
                if (o.valid) {
                    float* hrow = p.h1 + ((size_t)o.ray * p.S + o.s) * 64;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) hrow[h1_col(hg + 4 * (8 * half + jj))] = acc[s][jj];  // permuted columns
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        raw[s] += __shfl_xor_sync(0xffffffffu, raw[s], 1);
        raw[s] += __shfl_xor_sync(0xffffffffu, raw[s], 2);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const SamplePos o = samples.at(sg + 32 * s);
        if (o.valid && hg == s) p.density[(size_t)o.ray * p.S + o.s] = softplus_f(raw[s] + p.b2 + p.density_bias);  // models.py:L581
    }
}

int sample_encode_lmax(int L) {
    const int opts[] = {4, 6, 8, 10, 16};
    for (int o : opts)
        if (L <= o) return o;
    return 0;
}

// number of leading dense levels if the remaining ones are hashed with power-of-two tables, else -1
static int dense_prefix(const GridDesc& g) {
    int nd = 0;
    while (nd < g.num_levels && !g.lv[nd].hashed && g.lv[nd].mod_mode == 0) ++nd;
    for (int l = nd; l < g.num_levels; ++l)
        if (!g.lv[l].hashed || g.lv[l].mod_mode != 1) return -1;
    return nd;
}

template <int LMAX, bool NERF, int ND, int MINB, bool RUNS>
static int launch_one_impl(const SampleParams& p, cudaStream_t st) {
    const size_t total = (size_t)div_up(p.n_rays, 32u) * 32u * (size_t)p.S;  // ray groups of 32, see sample_pos
    const unsigned blocks = (unsigned)div_up(total, (size_t)kSampleThreads);
    // measured on B200 (profiles/r1_summary.md): the re-mapped density layer pays on the proposal level only
    constexpr int kFfma = (UC_REMAP_PROP ? !NERF : false) ? 1 : 0;
    if constexpr (!RUNS) {
        if (p.mlp_mma) {
            constexpr size_t smem = MmaLayout<LMAX>::smem_bytes;
            UC_ENSURE_SMEM(smem, sample_encode_kernel<LMAX, NERF, ND, MINB, 2, RUNS>);
            sample_encode_kernel<LMAX, NERF, ND, MINB, 2, RUNS><<<blocks, kSampleThreads, smem, st>>>(p);
            UC_LAUNCH_CHECK();
            return 0;
        }
    }
    constexpr size_t smem = sizeof(float) * ((64 + ((kFfma || RUNS) ? kSampleThreads : 0)) * (LMAX * 4 + 4) + 128);
    UC_ENSURE_SMEM(smem, sample_encode_kernel<LMAX, NERF, ND, MINB, kFfma, RUNS>);
    sample_encode_kernel<LMAX, NERF, ND, MINB, kFfma, RUNS><<<blocks, kSampleThreads, smem, st>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

template <int LMAX, bool NERF, int ND, int MINB>
static int launch_one(const SampleParams& p, cudaStream_t st) { return launch_one_impl<LMAX, NERF, ND, MINB, false>(p, st); }

template <int LMAX, int MINB>
static int launch_lmax(const SampleParams& p, bool nerf, int nd, cudaStream_t st) {
    if (p.cell_runs)  // level loop rolled: one instantiation per (LMAX, level kind)
        return nerf ? launch_one_impl<LMAX, true, -1, MINB, true>(p, st) : launch_one_impl<LMAX, false, -1, MINB, true>(p, st);
    if (nd == 3) return nerf ? launch_one<LMAX, true, 3, MINB>(p, st) : launch_one<LMAX, false, 3, MINB>(p, st);
    if constexpr (LMAX == 4) {
        if (nd == 1) return nerf ? launch_one<LMAX, true, 1, MINB>(p, st) : launch_one<LMAX, false, 1, MINB>(p, st);
    }
    return nerf ? launch_one<LMAX, true, -1, MINB>(p, st) : launch_one<LMAX, false, -1, MINB>(p, st);
}

int launch_sample_encode(const SampleParams& p, bool nerf, cudaStream_t st) {
    if (p.n_rays == 0) return 0;
    const int nd = dense_prefix(p.grid);
    switch (sample_encode_lmax(p.grid.num_levels)) {
        case 4: return launch_lmax<4, UC_MINB_SMALL>(p, nerf, nd, st);
        case 6: return launch_lmax<6, UC_MINB_SMALL>(p, nerf, nd, st);
        case 8: return launch_lmax<8, UC_MINB_LARGE>(p, nerf, nd, st);
        case 10: return launch_lmax<10, UC_MINB_LARGE>(p, nerf, nd, st);
        case 16: return launch_lmax<16, 3>(p, nerf, nd, st);
        default: set_error("sample_encode: grid levels must be <= 16"); return 1;
    }
}

}  // namespace ucnerf

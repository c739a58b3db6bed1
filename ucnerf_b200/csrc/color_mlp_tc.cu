// Colour MLP on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate through the 3xTF32 split.
//
//   a   = relu(P0 [h1, direnc] + c0')            (K = 96,  N = 256)      -> TMEM accumulator "acc3"
//   a2  = relu(P1 [h1, direnc] + V1a a + c1')    (K = 352, N = 256)      -> TMEM accumulator "acc4"
//   rgb = sigmoid(R a2 + r0) * (1 + 2 pad) - pad (N = 3, CUDA cores in the epilogue)
// (the activation-free bottleneck layer of models.py:L438-441/L601 is folded into P0 / P1 on the host.)
//
// Why 3xTF32: the parity bar is rgb L-inf < 1e-4 against an fp32 reference.  One TF32 pass (10-bit mantissa)
// leaves ~1e-3 on the pre-activations.  Splitting x = hi + lo (both TF32-representable) and accumulating
// hi*hi + lo*hi + hi*lo in the fp32 TMEM accumulator keeps ~2^-21 relative error per product at 3 MMAs per
// k-step, still ~5x the fp32 CUDA-core rate.
//
// CTA = one 128-row tile at a time (persistent over tiles), 10 warps:
//   warps 0-7  two producer / epilogue warpgroups (thread t of a group owns row t; group g fills A slot g, i.e.
//              the chunks with c % 2 == g).  They build the A operand chunk by chunk in shared memory
//              in the UMMA canonical K-major SWIZZLE_128B layout (hi tile + lo tile): from global h1 / the
//              computed view-direction encoding, or from acc3 in TMEM (tcgen05.ld -> +bias -> relu -> split).
//              Finally drains acc4, applies the rgb layer + sigmoid and writes the sample colours.
//   warp 8     one elected thread issues tcgen05.mma (M=128, N=256, K=8, kind::tf32) and tcgen05.commit.
//   warp 9     one elected thread streams the pre-swizzled weight chunks (hi|lo, 64 KB each) from L2 with
//              cp.async.bulk (TMA bulk copy, completes on an mbarrier).
// Pipelines: A ring (2 x 32 KB) and B ring (2 x 64 KB) with full/empty mbarriers; acc3/acc4 full/empty
// mbarriers order MMA vs. TMEM drains.  TMEM: all 512 columns (acc3 = [0,256), acc4 = [256,512)).
#include <cstring>

#include "ray_march.cuh"

namespace ucnerf {

namespace tc {

constexpr int kTileM = 128;
constexpr int kN = 256;
constexpr int kKC = 32;                       // K elements per chunk = one 128-byte swizzle row
constexpr int kSteps = 12;                    // chunks per tile (see the step table below)
constexpr uint32_t kATileBytes = kTileM * kKC * 4;   // 16 KB (one of hi / lo)
constexpr uint32_t kASlotBytes = 2 * kATileBytes;    // 32 KB
constexpr uint32_t kBTileBytes = kN * kKC * 4;       // 32 KB
constexpr uint32_t kBSlotBytes = 2 * kBTileBytes;    // 64 KB
constexpr int kStages = 2;
constexpr uint32_t kSmemA = 0;
constexpr uint32_t kSmemB = kSmemA + kStages * kASlotBytes;            // 65536
constexpr uint32_t kSmemMisc = kSmemB + kStages * kBSlotBytes;         // 196608
// misc region: barriers (16 x 8 B), tmem ptr, then c0[256], c1[256], R[256] float4, r0[4]
constexpr uint32_t kOffBar = 0;
constexpr uint32_t kOffTmem = 128;
constexpr uint32_t kOffC0 = 256;
constexpr uint32_t kOffC1 = kOffC0 + 1024;
constexpr uint32_t kOffR = kOffC1 + 1024;
constexpr uint32_t kOffR0 = kOffR + 4096;
constexpr uint32_t kOffPart = kOffR0 + 16;          // [128] float4: rgb partial sums handed from group 1 to group 0
constexpr uint32_t kMiscBytes = kOffPart + 2048;
constexpr uint32_t kSmemTotal = kSmemMisc + kMiscBytes + 1024;  // +1024: manual 1 KB alignment slack
constexpr int kThreads = 320;                     // warps 0-3 / 4-7: producer groups, 8: MMA, 9: weight loader
constexpr int kMmaWarp = 8;

// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=TF32 [7,10), b=TF32 [10,13),
// a/b K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

enum Bar { A_FULL0 = 0, A_FULL1, A_EMPTY0, A_EMPTY1, B_FULL0, B_FULL1, B_EMPTY0, B_EMPTY1, ACC3_FULL, ACC4_FULL,
           ACC3_EMPTY, ACC4_EMPTY, PART_FULL, PART_EMPTY, NUM_BARS };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug or a faulted copy must surface as an error code, never as a hung GPU.  The first
// thread whose wait exceeds kWatchdogNs records where it was stuck in p.dbg and raises the abort flag; every other
// wait loop polls the flag and bails out, so the kernel drains and the host reports the record.
constexpr unsigned long long kWatchdogNs = 400ull * 1000 * 1000;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, uint32_t* dbg, uint32_t tag, uint32_t bar_id,
                                            uint32_t it, uint32_t step) {
    const unsigned long long t0 = gtimer();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 255u) == 0) {
            if (*reinterpret_cast<volatile uint32_t*>(dbg) != 0u) return false;
            if (gtimer() - t0 > kWatchdogNs) {
                if (atomicCAS(dbg, 0u, tag) == 0u) {
                    dbg[1] = blockIdx.x; dbg[2] = threadIdx.x; dbg[3] = bar_id; dbg[4] = parity; dbg[5] = it; dbg[6] = step;
                    __threadfence();
                }
                return false;
            }
        }
    }
    return true;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, uint32_t* dbg, uint32_t tag, uint32_t bar_id,
                                          uint32_t it, uint32_t step) {
    if (mbar_try_wait(bar, parity)) return true;
    return mbar_wait_slow(bar, parity, dbg, tag, bar_id, it, step);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_128B, 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// round-to-nearest (ties away) to TF32 = cvt.rna.tf32.f32, done with two integer ops: the conversion pipe issues
// at a quarter of the ALU rate and 64 conversions per row and chunk made it the producers' bottleneck
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

// write one row (32 fp32 values) of an A chunk as hi / lo TF32 tiles in the SWIZZLE_128B K-major layout
__device__ __forceinline__ void store_a_row(uint8_t* slot, int row, const float (&v)[32]) {
    uint8_t* base = slot + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 hi, lo;
        uint32_t h;
        h = to_tf32(v[4 * c + 0]); hi.x = h; lo.x = to_tf32(v[4 * c + 0] - __uint_as_float(h));
        h = to_tf32(v[4 * c + 1]); hi.y = h; lo.y = to_tf32(v[4 * c + 1] - __uint_as_float(h));
        h = to_tf32(v[4 * c + 2]); hi.z = h; lo.z = to_tf32(v[4 * c + 2] - __uint_as_float(h));
        h = to_tf32(v[4 * c + 3]); hi.w = h; lo.w = to_tf32(v[4 * c + 3] - __uint_as_float(h));
        const int pc = (c ^ (row & 7)) * 16;
        *reinterpret_cast<uint4*>(base + pc) = hi;
        *reinterpret_cast<uint4*>(base + kATileBytes + pc) = lo;
    }
}

}  // namespace tc

using namespace tc;

// Step table of one tile (A chunk source x weight chunk = step index in the blob -> accumulator):
//   0: h1[0:32]  x P0 -> acc3 (init)     2: h1[0:32]  x P1 -> acc4 (init)     4..11: a[32j:32j+32] x V1a -> acc4
//   1: h1[32:64] x P0 -> acc3, commit    3: h1[32:64] x P1 -> acc4                   11: commit acc4
// The view-direction encoding is constant per ray, so its contribution (P0d direnc + c0', P1d direnc + c1') is
// evaluated once per ray by dir_bias_kernel (fp32 FMAs) and added as a per-ray bias when the accumulators are
// drained: two of fourteen MMA steps and all sinf() evaluations leave this kernel.
// Producer warpgroup g (0/1) builds the chunks with (c & 1) == g into A slot g, so the two groups alternate and
// each chunk's production overlaps the MMAs of the previous one.  The final epilogue of tile i (drain acc4, rgb
// layer) is split by accumulator columns between the groups and is executed *after* each group has produced its
// first chunk of tile i+1, so the tensor pipe already works on the next tile while acc4 is drained.
__global__ void __launch_bounds__(kThreads, 1)
color_mlp_tc_kernel(const __grid_constant__ ColorTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* misc = smem + kSmemMisc;
    const uint32_t bar0 = smem_u32(misc + kOffBar);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(misc + kOffTmem);
    float4* sR = reinterpret_cast<float4*>(misc + kOffR);
    float* sR0 = reinterpret_cast<float*>(misc + kOffR0);
    float4* sPart = reinterpret_cast<float4*>(misc + kOffPart);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < 256; i += kThreads) {
        sR[i] = reinterpret_cast<const float4*>(p.rt)[i];
    }
    if (threadIdx.x < 4) sR0[threadIdx.x] = p.r0[threadIdx.x];
    if (threadIdx.x == 0) {
        mbar_init(BAR(A_FULL0), 128); mbar_init(BAR(A_FULL1), 128);
        mbar_init(BAR(A_EMPTY0), 1); mbar_init(BAR(A_EMPTY1), 1);
        mbar_init(BAR(B_FULL0), 1); mbar_init(BAR(B_FULL1), 1);
        mbar_init(BAR(B_EMPTY0), 1); mbar_init(BAR(B_EMPTY1), 1);
        mbar_init(BAR(ACC3_FULL), 1); mbar_init(BAR(ACC4_FULL), 1);
        mbar_init(BAR(ACC3_EMPTY), 256); mbar_init(BAR(ACC4_EMPTY), 256);
        mbar_init(BAR(PART_FULL), 128); mbar_init(BAR(PART_EMPTY), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {  // whole warp: allocate all 512 TMEM columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(misc + kOffTmem)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const uint32_t ntiles = (p.n_rows + kTileM - 1) / kTileM;

    if (warp < 8) {
        // ================= producer / epilogue warpgroups: thread <-> row t of the tile ================
        const int g = warp >> 2;                 // warpgroup 0 / 1  == A slot it fills
        const int t = threadIdx.x & 127;         // row inside the tile
        const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint8_t* my_slot = smem + kSmemA + g * kASlotBytes;
        uint32_t k = 0;                          // chunks produced by this group so far (A_EMPTY phase)
        uint32_t it = 0, prev_tile = 0;
        // profiling (debug_flags bit 2): thread 0 of each group in CTA 0 -> dbg[16 + 8 g ...]: total, waits on
        // A_EMPTY / ACC3_FULL / ACC4_FULL+PART, time in h1 chunks / dir chunks / tmem chunks (kilo-cycles)
        const bool prof = (p.debug_flags & 4u) != 0 && t == 0 && blockIdx.x == 0;
        long long pw_aempty = 0, pw_acc3 = 0, pw_epi = 0, pt_h1 = 0, pt_dir = 0, pt_tm = 0, pt_store = 0, pt_fence = 0;
        const long long pt_begin = clock64();

        // drain this group's half of acc4 for tile `tl` (iteration `itp`), rgb layer, sigmoid, store
        auto final_epilogue = [&](uint32_t tl, uint32_t itp) -> bool {
            if (!mbar_wait(BAR(ACC4_FULL), itp & 1, p.dbg, 3, ACC4_FULL, itp, 99)) return false;
            tc_fence_after();
            float o0 = 0.f, o1 = 0.f, o2 = 0.f;
            const uint32_t erow = tl * kTileM + t;
            const float4* bias1 = reinterpret_cast<const float4*>(
                p.dir_bias + (size_t)((erow < p.n_rows ? erow : 0u) / (uint32_t)p.S) * 512 + 256 + g * 128);
            // software pipeline: the TMEM load of the next 32 columns is in flight while the current ones are consumed
            uint32_t ra[32], rb[32];
            tmem_ld32_issue(lane_taddr + (uint32_t)(256 + g * 128), ra);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float4 bb[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) bb[q] = __ldg(bias1 + 8 * jj + q);
                tmem_ld_wait();
                uint32_t (&cur)[32] = (jj & 1) ? rb : ra;
                uint32_t (&nxt)[32] = (jj & 1) ? ra : rb;
                if (jj < 3) tmem_ld32_issue(lane_taddr + (uint32_t)(256 + g * 128 + 32 * (jj + 1)), nxt);
                const int col = g * 128 + 32 * jj;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float bq[4] = {bb[q].x, bb[q].y, bb[q].z, bb[q].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float a2 = fmaxf(__uint_as_float(cur[4 * q + e]) + bq[e], 0.f);
                        const float4 w = sR[col + 4 * q + e];
                        o0 = fmaf(a2, w.x, o0);
                        o1 = fmaf(a2, w.y, o1);
                        o2 = fmaf(a2, w.z, o2);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(BAR(ACC4_EMPTY));
            if (g == 1) {
                if (!mbar_wait(BAR(PART_EMPTY), (itp & 1) ^ 1, p.dbg, 9, PART_EMPTY, itp, 99)) return false;
                sPart[t] = make_float4(o0, o1, o2, 0.f);
                mbar_arrive(BAR(PART_FULL));
            } else {
                if (!mbar_wait(BAR(PART_FULL), itp & 1, p.dbg, 10, PART_FULL, itp, 99)) return false;
                const float4 q = sPart[t];
                mbar_arrive(BAR(PART_EMPTY));
                const uint32_t row = tl * kTileM + t;
                if (row < p.n_rows) {
                    float* o = p.rgb + (size_t)row * 3;
                    o[0] = fs(fm(sigmoid_f(o0 + q.x + sR0[0]), p.rgb_scale), p.rgb_padding);
                    o[1] = fs(fm(sigmoid_f(o1 + q.y + sR0[1]), p.rgb_scale), p.rgb_padding);
                    o[2] = fs(fm(sigmoid_f(o2 + q.z + sR0[2]), p.rgb_scale), p.rgb_padding);
                }
            }
            return true;
        };

        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t row = tile * kTileM + t;
            const bool valid = row < p.n_rows;
            const float* h1row = p.h1 + (size_t)(valid ? row : 0) * 64;
            const float* bias0 = p.dir_bias + (size_t)((valid ? row : 0u) / (uint32_t)p.S) * 512;  // per-ray [c0' | c1'] rows
            float hkeep[32];
#pragma unroll 1
            for (int c = g; c < kSteps; c += 2) {
                float v[32];
                const long long tc0 = prof ? clock64() : 0;
                long long tw = 0;
                if (c < 2) {  // this group's half of the h1 row: loaded once, kept in registers for chunk c + 2
                    const float4* src = reinterpret_cast<const float4*>(h1row + 32 * g);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 x = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                        hkeep[4 * q] = x.x; hkeep[4 * q + 1] = x.y; hkeep[4 * q + 2] = x.z; hkeep[4 * q + 3] = x.w;
                    }
                }
                if (c < 4) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = hkeep[i];
                } else {
                    const int j = c - 4;
                    if (j < 2) {  // first TMEM chunk of this group for this tile
                        const long long ta = prof ? clock64() : 0;
                        if (!mbar_wait(BAR(ACC3_FULL), it & 1, p.dbg, 1, ACC3_FULL, it, c)) goto teardown;
                        if (prof) { tw = clock64() - ta; pw_acc3 += tw; }
                        tc_fence_after();
                    }
                    uint32_t r[32];
                    const long long tl0 = prof ? clock64() : 0;
                    tmem_ld32_issue(lane_taddr + (uint32_t)(32 * j), r);
                    const float4* b0 = reinterpret_cast<const float4*>(bias0 + 32 * j);
                    float4 bb[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) bb[q] = __ldg(b0 + q);  // overlaps the TMEM load
                    tmem_ld_wait();
                    if (prof) pt_dir += clock64() - tl0;  // profiler slot "dir" = TMEM load + bias load latency
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        v[4 * q] = fmaxf(__uint_as_float(r[4 * q]) + bb[q].x, 0.f);
                        v[4 * q + 1] = fmaxf(__uint_as_float(r[4 * q + 1]) + bb[q].y, 0.f);
                        v[4 * q + 2] = fmaxf(__uint_as_float(r[4 * q + 2]) + bb[q].z, 0.f);
                        v[4 * q + 3] = fmaxf(__uint_as_float(r[4 * q + 3]) + bb[q].w, 0.f);
                    }
                    if (j >= 6) {  // this group's part of acc3 is drained
                        tc_fence_before();
                        mbar_arrive(BAR(ACC3_EMPTY));
                    }
                }
                const long long tb = prof ? clock64() : 0;
                if (!mbar_wait(BAR(A_EMPTY0 + g), (k & 1) ^ 1, p.dbg, 2, A_EMPTY0 + g, it, c)) goto teardown;
                const long long tb2 = prof ? clock64() : 0;
                if (prof) { pw_aempty += tb2 - tb; tw += tb2 - tb; }
                if (!(p.debug_flags & 2u)) store_a_row(my_slot, t, v);  // bit 1: profiling experiment, skip A stores
                const long long ts1 = prof ? clock64() : 0;
                fence_proxy_async();
                const long long ts2 = prof ? clock64() : 0;
                mbar_arrive(BAR(A_FULL0 + g));
                if (prof) { pt_store += ts1 - tb2; pt_fence += ts2 - ts1; }
                ++k;
                if (prof) {
                    const long long work = clock64() - tc0 - tw;
                    if (c < 4) pt_h1 += work; else pt_tm += work;
                }
                if (c == g && it > 0) {
                    const long long te = prof ? clock64() : 0;
                    if (!final_epilogue(prev_tile, it - 1)) goto teardown;
                    if (prof) pw_epi += clock64() - te;
                }
            }
            prev_tile = tile;
        }
        if (it > 0) {
            if (!final_epilogue(prev_tile, it - 1)) goto teardown;
        }
        if (prof) {
            uint32_t* d = p.dbg + 16 + 8 * g;
            d[0] = (uint32_t)((clock64() - pt_begin) >> 10); d[1] = (uint32_t)(pw_aempty >> 10); d[2] = (uint32_t)(pw_acc3 >> 10);
            d[3] = (uint32_t)(pw_epi >> 10); d[4] = (uint32_t)(pt_h1 >> 10); d[5] = (uint32_t)(pt_dir >> 10); d[6] = (uint32_t)(pt_tm >> 10);
            d[7] = (uint32_t)((pt_store >> 10) | ((pt_fence >> 10) << 16));
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (one thread) ==================================================
        if (lane == 0) {
            uint32_t slot = 0, phase = 0, it = 0;
            // profiling (debug_flags bit 2): cycles this thread spent waiting per barrier kind, CTA 0 -> dbg[8..12]
            const bool prof = (p.debug_flags & 4u) != 0;
            long long w_acc3 = 0, w_acc4 = 0, w_b = 0, w_a = 0;
            const long long t_begin = clock64();
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
#pragma unroll 1
                for (int s = 0; s < kSteps; ++s) {
                    long long t0 = prof ? clock64() : 0;
                    if (s == 0 && !mbar_wait(BAR(ACC3_EMPTY), (it & 1) ^ 1, p.dbg, 4, ACC3_EMPTY, it, s)) goto teardown;
                    if (prof) { const long long t1 = clock64(); w_acc3 += t1 - t0; t0 = t1; }
                    if (s == 2 && !mbar_wait(BAR(ACC4_EMPTY), (it & 1) ^ 1, p.dbg, 5, ACC4_EMPTY, it, s)) goto teardown;
                    if (prof) { const long long t1 = clock64(); w_acc4 += t1 - t0; t0 = t1; }
                    if (!mbar_wait(BAR(B_FULL0 + slot), phase, p.dbg, 6, B_FULL0 + slot, it, s)) goto teardown;
                    if (prof) { const long long t1 = clock64(); w_b += t1 - t0; t0 = t1; }
                    if (!mbar_wait(BAR(A_FULL0 + slot), phase, p.dbg, 7, A_FULL0 + slot, it, s)) goto teardown;
                    if (prof) { const long long t1 = clock64(); w_a += t1 - t0; t0 = t1; }
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + kSmemA + slot * kASlotBytes);
                    const uint32_t a_lo = a_hi + kATileBytes;
                    const uint32_t b_hi = smem_u32(smem + kSmemB + slot * kBSlotBytes);
                    const uint32_t b_lo = b_hi + kBTileBytes;
                    const uint32_t acc = tmem_base + (s < 2 ? 0u : 256u);
                    const bool init = (s == 0 || s == 2);
#pragma unroll
                    for (int ks = 0; ks < kKC / 8; ++ks) {
                        const uint64_t dah = make_desc(a_hi + ks * 32), dal = make_desc(a_lo + ks * 32);
                        const uint64_t dbh = make_desc(b_hi + ks * 32), dbl = make_desc(b_lo + ks * 32);
                        umma_tf32(acc, dah, dbh, (init && ks == 0) ? 0u : 1u);
                        umma_tf32(acc, dal, dbh, 1u);
                        umma_tf32(acc, dah, dbl, 1u);
                    }
                    umma_commit(BAR(A_EMPTY0 + slot));
                    umma_commit(BAR(B_EMPTY0 + slot));
                    if (s == 1) umma_commit(BAR(ACC3_FULL));
                    if (s == kSteps - 1) umma_commit(BAR(ACC4_FULL));
                    slot ^= 1;
                    phase ^= (slot == 0);
                }
            }
            if (prof && blockIdx.x == 0) {
                p.dbg[8] = (uint32_t)((clock64() - t_begin) >> 10);
                p.dbg[9] = (uint32_t)(w_acc3 >> 10); p.dbg[10] = (uint32_t)(w_acc4 >> 10);
                p.dbg[11] = (uint32_t)(w_b >> 10); p.dbg[12] = (uint32_t)(w_a >> 10); p.dbg[13] = it;
            }
        }
    } else {
        // ================= weight loader (one thread): pre-swizzled hi|lo chunks, L2 -> smem ==========
        if (lane == 0) {
            uint32_t slot = 0, phase = 0, it = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
#pragma unroll 1
                for (int s = 0; s < kSteps; ++s) {
                    if (!mbar_wait(BAR(B_EMPTY0 + slot), phase ^ 1, p.dbg, 8, B_EMPTY0 + slot, it, s)) goto teardown;
                    if (p.debug_flags & 1u) {  // profiling experiment: no weight traffic (results are garbage)
                        mbar_arrive(BAR(B_FULL0 + slot));
                    } else {
                        mbar_expect_tx(BAR(B_FULL0 + slot), kBSlotBytes);
                        bulk_g2s(smem_u32(smem + kSmemB + slot * kBSlotBytes), p.wblob + (size_t)s * kBSlotBytes,
                                 kBSlotBytes, BAR(B_FULL0 + slot));
                    }
                    slot ^= 1;
                    phase ^= (slot == 0);
                }
            }
        }
    }

teardown:
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// Per-ray constant part of both colour layers: out[ray][0:256] = c0' + P0d direnc(viewdir), out[ray][256:512] =
// c1' + P1d direnc(viewdir), direnc = pos_enc(viewdirs, 0, 4) (coord.py:L214-225, 27 values).  fp32 FMAs.
__global__ void __launch_bounds__(256)
dir_bias_kernel(const float* __restrict__ viewdirs, const float* __restrict__ wdir /* [2][32][256] */,
                const float* __restrict__ c0, const float* __restrict__ c1, float* __restrict__ out, uint32_t n_rays) {
    __shared__ float enc[8][32];
    const uint32_t ray0 = blockIdx.x * 8;
    if (threadIdx.x < 8 * 32) {
        const int r = threadIdx.x >> 5, c = threadIdx.x & 31;
        float val = 0.f;
        if (ray0 + r < n_rays && c < 27) {
            const float* vd = viewdirs + 3 * (size_t)(ray0 + r);
            if (c < 3) {
                val = vd[c];
            } else {
                const int q = c - 3, qq = q < 12 ? q : q - 12, deg = qq / 3, ax = qq - 3 * deg;
                float x = fm(vd[ax], (float)(1 << deg));
                if (q >= 12) x = fa(x, 1.57079637f);
                val = sinf(x);
            }
        }
        enc[r][c] = val;
    }
    __syncthreads();
    const int n = threadIdx.x;
    float w0[27], w1[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        w0[k] = __ldg(wdir + (size_t)k * 256 + n);
        w1[k] = __ldg(wdir + (size_t)(32 + k) * 256 + n);
    }
    const float b0 = c0[n], b1 = c1[n];
    for (int r = 0; r < 8 && ray0 + r < n_rays; ++r) {
        float a0 = b0, a1 = b1;
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            a0 = fmaf(w0[k], enc[r][k], a0);
            a1 = fmaf(w1[k], enc[r][k], a1);
        }
        out[(size_t)(ray0 + r) * 512 + n] = a0;
        out[(size_t)(ray0 + r) * 512 + 256 + n] = a1;
    }
}

int launch_dir_bias(const float* viewdirs, const float* wdir, const float* c0, const float* c1, float* out,
                    uint32_t n_rays, cudaStream_t st) {
    if (n_rays == 0) return 0;
    dir_bias_kernel<<<div_up(n_rays, 8u), 256, 0, st>>>(viewdirs, wdir, c0, c1, out, n_rays);
    UC_LAUNCH_CHECK();
    return 0;
}

static uint32_t* g_tc_dbg = nullptr;   // [32] words: watchdog record of color_mlp_tc_kernel (0 = healthy)

int color_tc_status(uint32_t* out16) {
    for (int i = 0; i < 32; ++i) out16[i] = 0;
    if (!g_tc_dbg) return 0;
    UC_CUDA_OK(cudaDeviceSynchronize());
    UC_CUDA_OK(cudaMemcpy(out16, g_tc_dbg, 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

int launch_color_mlp_tc(const ColorTcParams& p_in, cudaStream_t st) {
    if (p_in.n_rows == 0) return 0;
    if (!g_tc_dbg) {
        UC_CUDA_OK(cudaMalloc(&g_tc_dbg, 32 * sizeof(uint32_t)));
        UC_CUDA_OK(cudaMemset(g_tc_dbg, 0, 32 * sizeof(uint32_t)));
    }
    ColorTcParams p = p_in;
    p.dbg = g_tc_dbg;
    static bool configured = false;
    if (!configured) {
        UC_CUDA_OK(cudaFuncSetAttribute(color_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal));
        configured = true;
    }
    const uint32_t ntiles = (p.n_rows + kTileM - 1) / kTileM;
    const uint32_t blocks = ntiles < (uint32_t)kNumSMs ? ntiles : (uint32_t)kNumSMs;
    color_mlp_tc_kernel<<<blocks, kThreads, kSmemTotal, st>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

uint32_t color_tc_blob_bytes() { return kSteps * kBSlotBytes; }

// Host: lay the folded weights out as the kernel consumes them.  Wt chunks are K-major [k][256] fp32 arrays of 32
// rows each (step order of the kernel); every chunk becomes hi tile | lo tile, each in the UMMA K-major
// SWIZZLE_128B image: element (n, k) at (n/8)*1024 + (n%8)*128 + ((k/4) ^ (n%8))*16 + (k%4)*4.
static inline uint32_t tf32_rna_bits(float x) {
    uint32_t b;
    memcpy(&b, &x, 4);
    if ((b & 0x7F800000u) == 0x7F800000u) return b;
    b += 0x1000u;
    return b & 0xFFFFE000u;
}
void color_tc_pack_chunk(const float* wt_rows /* [32][256] */, uint8_t* dst /* 64 KB */) {
    for (int n = 0; n < kN; ++n)
        for (int k = 0; k < kKC; ++k) {
            const float w = wt_rows[(size_t)k * kN + n];
            const uint32_t hb = tf32_rna_bits(w);
            float hf;
            memcpy(&hf, &hb, 4);
            const uint32_t lb = tf32_rna_bits(w - hf);
            const size_t off = (size_t)(n >> 3) * 1024 + (n & 7) * 128 + (((k >> 2) ^ (n & 7)) * 16) + (k & 3) * 4;
            memcpy(dst + off, &hb, 4);
            memcpy(dst + kBTileBytes + off, &lb, 4);
        }
}

}  // namespace ucnerf

// Colour MLP on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate through a 3-term FP16 split.
//
//   a   = relu(P0 h1 + [P0d direnc + c0'])            (K = 64,  N = 256)      -> TMEM accumulator "acc3"
//   a2  = relu(P1 h1 + V1a a + [P1d direnc + c1'])    (K = 320, N = 256)      -> TMEM accumulator "acc4"
//   rgb = sigmoid(R a2 + r0) * (1 + 2 pad) - pad      (N = 3, CUDA cores in the epilogue)
// (the activation-free bottleneck layer of models.py:L438-441/L601 is folded into P0 / P1 on the host; the bracketed
//  view-direction terms are per-ray biases from dir_bias_kernel.)
//
// Why a split: the parity bar is rgb L-inf < 1e-4 against an fp32 reference and one low-precision pass (TF32/BF16/FP16)
// leaves ~1e-3 on the pre-activations.  Every operand is written as x * s = hi + lo with hi, lo in FP16 (s a power
// of two chosen so that hi uses the upper part of the FP16 range: 22 significant bits per operand) and
// hi*hi + lo*hi + hi*lo is accumulated in the fp32 TMEM accumulator: ~2^-22 relative error per product, the same as
// the 3xTF32 split this kernel used first, at HALF the tensor-pipe time (kind::f16 retires K = 16 per instruction at
// the rate kind::tf32 retires K = 8) and half the shared-memory / L2 operand traffic (2-byte elements).
// Range: activations are scaled by kActScale = 8, so |h1|, |a| must stay below 65504 / 8 = 8188; larger values become
// inf in the FP16 hi part and surface as NaN colours (never as silently wrong ones) - use color_mlp = 0 for such nets.
//
// CTA = one 128-row tile at a time (persistent over tiles), 10 warps:
//   warps 0-7  two producer / epilogue warpgroups (thread t of a group owns row t).  They build the A operand chunk
//              by chunk (64 K-elements = one 128-byte swizzle row of FP16) in shared memory in the UMMA canonical
//              K-major SWIZZLE_128B layout (hi tile + lo tile): the h1 tile ONCE per tile with coalesced loads by all
//              eight warps (it feeds steps 0 and 1), the four chunks of `a` from acc3 in TMEM (tcgen05.ld -> scale +
//              per-ray bias -> relu -> split), group g converting columns [32 g, 32 g + 32) of every chunk.
//              Finally they drain acc4 (group g: columns [128 g, 128 g + 128)), apply the rgb layer + sigmoid and
//              write the sample colours.
//   warp 8     one elected thread issues tcgen05.mma (M=128, N=256, K=16, kind::f16) and tcgen05.commit.
//   warp 9     one elected thread streams the pre-swizzled weight chunks (hi|lo, 64 KB each) and, per tile, the per-ray
//              bias rows of the rays the tile touches from L2 with cp.async.bulk (completes on an mbarrier).
// Pipelines: A ring (2 x 32 KB, slot = running chunk counter & 1, 5 chunks per tile) and B ring (2 x 64 KB) with
// full/empty mbarriers; acc3/acc4 full/empty mbarriers order MMA vs. TMEM drains; a 2-stage ring for the bias rows.
// TMEM: all 512 columns (acc3 = [0,256), acc4 = [256,512)).
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>

#include "ray_march.cuh"
#include "tc_common.cuh"

namespace ucnerf {

namespace tc {

constexpr int kN = 256;
constexpr int kSteps = 6;                     // chunks per tile (see the step table below)
constexpr float kActScale = 8.f;              // power-of-two scale of the A operand (h1, a) before the FP16 split
constexpr uint32_t kBTileBytes = kN * kKC * 2;       // 32 KB
constexpr uint32_t kBSlotBytes = 2 * kBTileBytes;    // 64 KB
constexpr int kStages = 2;
constexpr uint32_t kSmemA = 0;
constexpr uint32_t kSmemB = kSmemA + kStages * kASlotBytes;            // 65536
constexpr uint32_t kSmemMisc = kSmemB + kStages * kBSlotBytes;         // 196608
// misc region: barriers (NUM_BARS x 8 B), tmem ptr, then c0[256], c1[256], R[256] float4, r0[4]
constexpr uint32_t kOffBar = 0;
constexpr uint32_t kOffTmem = 192;
constexpr uint32_t kOffC0 = 256;
constexpr uint32_t kOffC1 = kOffC0 + 1024;
constexpr uint32_t kOffR = kOffC1 + 1024;
constexpr uint32_t kOffR0 = kOffR + 4096;
constexpr uint32_t kOffPart = kOffR0 + 16;          // [128] float4: rgb partial sums handed from group 1 to group 0
constexpr uint32_t kMiscBytes = kOffPart + 2048;
// per-tile staging of the per-ray bias rows (dir_bias, 2 KB per ray): a 128-row tile touches at most kBiasRays rays
// when S >= 32; two stages, filled by the loader thread with one bulk copy per tile (the rows of consecutive rays are
// contiguous), so the producer groups read their biases from shared memory instead of waiting on L2 / HBM latency
// six times per tile
constexpr int kBiasRays = 5;
constexpr uint32_t kBiasStageBytes = kBiasRays * 2048;
constexpr uint32_t kSmemBias = kSmemMisc + ((kMiscBytes + 127) & ~127u);
constexpr uint32_t kSmemTotal = kSmemBias + 2 * kBiasStageBytes + 1024;  // +1024: manual 1 KB alignment slack
static_assert(kSmemTotal <= 232448, "shared memory budget");
constexpr int kThreads = 320;                     // warps 0-3 / 4-7: producer groups, 8: MMA, 9: weight loader
constexpr int kMmaWarp = 8;

constexpr uint32_t kIdesc = make_idesc(kN);

enum Bar { A_FULL0 = 0, A_FULL1, A_EMPTY0, A_EMPTY1, B_FULL0, B_FULL1, B_EMPTY0, B_EMPTY1, ACC3_FULL, ACC4_FULL,
           ACC3_EMPTY, ACC4_EMPTY, PART_FULL, PART_EMPTY, BIAS_FULL0, BIAS_FULL1, BIAS_EMPTY0, BIAS_EMPTY1, NUM_BARS };
static_assert(NUM_BARS * 8 <= kOffTmem, "barrier block overlaps the TMEM pointer slot");

}  // namespace tc

using namespace tc;

// Step table of one tile (A chunk source x weight chunk = step index in the blob -> accumulator):
//   0: h1[0:64] x P0 -> acc3 (init), commit acc3      2..5: a[64j:64j+64] x V1a -> acc4 (j = step - 2)
//   1: h1[0:64] x P1 -> acc4 (init)                      5: commit acc4
// The view-direction encoding is constant per ray, so its contribution (P0d direnc + c0', P1d direnc + c1') is
// evaluated once per ray by dir_bias_kernel (fp32 FMAs) and added as a per-ray bias when the accumulators are
// drained: no MMA step and no sinf() evaluation for it in this kernel.
// Producer warpgroup g (0/1) builds the chunks with (c & 1) == g into A slot g, so the two groups alternate and
// each chunk's production overlaps the MMAs of the previous one.  The final epilogue of tile i (drain acc4, rgb
// layer) is split by accumulator columns between the groups and is executed *after* each group has produced its
// first chunk of tile i+1, so the tensor pipe already works on the next tile while acc4 is drained.
// Scales: A operands carry kActScale, the weights of acc3 / acc4 carry the power-of-two factors chosen by the host
// (color_tc_weight_scale); acc3 * k0 + bias0 (bias0 pre-multiplied by kActScale) is directly the scaled `a`,
// acc4 * k1 + bias1 the unscaled pre-activation of the last hidden layer.
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

    // rgb-layer weights for the packed epilogue: entry 2 q = {x_k, x_k+1, y_k, y_k+1}, 2 q + 1 = {z_k, z_k+1, 0, 0}, k = 2 q
    for (int i = threadIdx.x; i < 128; i += kThreads) {
        const float4 w0 = reinterpret_cast<const float4*>(p.rt)[2 * i], w1 = reinterpret_cast<const float4*>(p.rt)[2 * i + 1];
        sR[2 * i] = make_float4(w0.x, w1.x, w0.y, w1.y);
        sR[2 * i + 1] = make_float4(w0.z, w1.z, 0.f, 0.f);
    }
    if (threadIdx.x < 4) sR0[threadIdx.x] = p.r0[threadIdx.x];
    if (threadIdx.x == 0) {
        mbar_init(BAR(A_FULL0), 256); mbar_init(BAR(A_FULL1), 256);
        mbar_init(BAR(A_EMPTY0), 1); mbar_init(BAR(A_EMPTY1), 1);
        mbar_init(BAR(B_FULL0), 1); mbar_init(BAR(B_FULL1), 1);
        mbar_init(BAR(B_EMPTY0), 1); mbar_init(BAR(B_EMPTY1), 1);
        mbar_init(BAR(ACC3_FULL), 1); mbar_init(BAR(ACC4_FULL), 1);
        mbar_init(BAR(ACC3_EMPTY), 256); mbar_init(BAR(ACC4_EMPTY), 256);
        mbar_init(BAR(PART_FULL), 128); mbar_init(BAR(PART_EMPTY), 128);
        mbar_init(BAR(BIAS_FULL0), 1); mbar_init(BAR(BIAS_FULL1), 1);
        mbar_init(BAR(BIAS_EMPTY0), 256); mbar_init(BAR(BIAS_EMPTY1), 256);
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
    // staged per-ray biases need a tile to span <= kBiasRays rays: S >= 32 (else the global-memory path is used)
    const bool smem_bias = p.S >= 32 && !(p.debug_flags & 16u);
    const uint32_t n_rays_total = p.n_rows / (uint32_t)p.S;
    const uint8_t* sBias = smem + kSmemBias;

    if (warp < 8) {
        // ================= producer / epilogue warpgroups: thread <-> row t of the tile ================
        const int g = warp >> 2;                 // warpgroup 0 / 1  == A slot it fills
        const int t = threadIdx.x & 127;         // row inside the tile
        const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t it = 0, prev_tile = 0;
        // profiling (debug_flags bit 2): thread 0 of each group in CTA 0 -> dbg[16 + 8 g ...]: total, waits on
        // A_EMPTY / ACC3_FULL / ACC4_FULL+PART, time in h1 chunks / dir chunks / tmem chunks (kilo-cycles)
        const bool prof = (p.debug_flags & 4u) != 0 && t == 0 && blockIdx.x == 0;
        long long pw_aempty = 0, pw_acc3 = 0, pw_epi = 0, pt_h1 = 0, pt_tm = 0;
        const long long pt_begin = clock64();

        // drain this group's half of acc4 for tile `tl` (iteration `itp`), rgb layer, sigmoid, store
        auto final_epilogue = [&](uint32_t tl, uint32_t itp) -> bool {
            if (!mbar_wait(BAR(ACC4_FULL), itp & 1, p.dbg, 3, ACC4_FULL, itp, 99)) return false;
            tc_fence_after();
            // packed fp32x2 math: two neighbouring columns per FFMA2 (even / odd partial sums of the three outputs)
            float2 o0 = make_float2(0.f, 0.f), o1 = make_float2(0.f, 0.f), o2 = make_float2(0.f, 0.f);
            const float2 k1 = make_float2(p.k1, p.k1);
            const uint32_t erow = tl * kTileM + t;
            const uint32_t eray = (erow < p.n_rows ? erow : 0u) / (uint32_t)p.S;
            const float4* bias1;
            if (smem_bias) {
                if (!mbar_wait(BAR(BIAS_FULL0 + (itp & 1)), (itp >> 1) & 1, p.dbg, 11, BIAS_FULL0 + (itp & 1), itp, 99)) return false;
                const uint32_t f = (tl * kTileM) / (uint32_t)p.S;
                const uint32_t lr = eray >= f ? min(eray - f, (uint32_t)kBiasRays - 1) : 0u;
                bias1 = reinterpret_cast<const float4*>(sBias + (itp & 1) * kBiasStageBytes + lr * 2048 + 1024 + g * 512);
            } else {
                bias1 = reinterpret_cast<const float4*>(p.dir_bias + (size_t)eray * 512 + 256 + g * 128);
            }
            // software pipeline: the TMEM load of the next 32 columns is in flight while the current ones are consumed
            uint32_t ra[32], rb[32];
            tmem_ld32_issue(lane_taddr + (uint32_t)(256 + g * 128), ra);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float4 bb[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) bb[q] = smem_bias ? bias1[8 * jj + q] : __ldg(bias1 + 8 * jj + q);
                tmem_ld_wait();
                uint32_t (&cur)[32] = (jj & 1) ? rb : ra;
                uint32_t (&nxt)[32] = (jj & 1) ? ra : rb;
                if (jj < 3) {
                    tmem_ld32_issue(lane_taddr + (uint32_t)(256 + g * 128 + 32 * (jj + 1)), nxt);
                } else {
                    // every TMEM read of this thread has completed: hand acc4 back before the last block's arithmetic
                    tc_fence_before();
                    mbar_arrive(BAR(ACC4_EMPTY));
                }
                const int pair0 = (g * 128 + 32 * jj) / 2;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float2 a01 = ffma2(make_float2(__uint_as_float(cur[4 * q]), __uint_as_float(cur[4 * q + 1])), k1, make_float2(bb[q].x, bb[q].y));
                    float2 a23 = ffma2(make_float2(__uint_as_float(cur[4 * q + 2]), __uint_as_float(cur[4 * q + 3])), k1, make_float2(bb[q].z, bb[q].w));
                    a01.x = fmaxf(a01.x, 0.f); a01.y = fmaxf(a01.y, 0.f);
                    a23.x = fmaxf(a23.x, 0.f); a23.y = fmaxf(a23.y, 0.f);
                    const float4 wa = sR[2 * (pair0 + 2 * q)], wb = sR[2 * (pair0 + 2 * q) + 1];
                    const float4 wc = sR[2 * (pair0 + 2 * q + 1)], wd = sR[2 * (pair0 + 2 * q + 1) + 1];
                    o0 = ffma2(a01, make_float2(wa.x, wa.y), o0);
                    o1 = ffma2(a01, make_float2(wa.z, wa.w), o1);
                    o2 = ffma2(a01, make_float2(wb.x, wb.y), o2);
                    o0 = ffma2(a23, make_float2(wc.x, wc.y), o0);
                    o1 = ffma2(a23, make_float2(wc.z, wc.w), o1);
                    o2 = ffma2(a23, make_float2(wd.x, wd.y), o2);
                }
            }
            const float s0 = o0.x + o0.y, s1 = o1.x + o1.y, s2 = o2.x + o2.y;
            if (smem_bias) mbar_arrive(BAR(BIAS_EMPTY0 + (itp & 1)));   // last read of this tile's staged biases
            if (g == 1) {
                if (!mbar_wait(BAR(PART_EMPTY), (itp & 1) ^ 1, p.dbg, 9, PART_EMPTY, itp, 99)) return false;
                sPart[t] = make_float4(s0, s1, s2, 0.f);
                mbar_arrive(BAR(PART_FULL));
            } else {
                if (!mbar_wait(BAR(PART_FULL), itp & 1, p.dbg, 10, PART_FULL, itp, 99)) return false;
                const float4 q = sPart[t];
                mbar_arrive(BAR(PART_EMPTY));
                const uint32_t row = tl * kTileM + t;
                if (row < p.n_rows) {
                    float* o = p.rgb + (size_t)row * 3;
                    o[0] = fs(fm(sigmoid_f(s0 + q.x + sR0[0]), p.rgb_scale), p.rgb_padding);
                    o[1] = fs(fm(sigmoid_f(s1 + q.y + sR0[1]), p.rgb_scale), p.rgb_padding);
                    o[2] = fs(fm(sigmoid_f(s2 + q.z + sR0[2]), p.rgb_scale), p.rgb_padding);
                }
            }
            return true;
        };

        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t row = tile * kTileM + t;
            const bool valid = row < p.n_rows;
            const uint32_t ray = (valid ? row : 0u) / (uint32_t)p.S;
            const float* bias0 = p.dir_bias + (size_t)ray * 512;  // per-ray [c0' | c1'] rows
            if (smem_bias) {
                const uint32_t f = (tile * kTileM) / (uint32_t)p.S;
                const uint32_t lr = ray >= f ? min(ray - f, (uint32_t)kBiasRays - 1) : 0u;
                bias0 = reinterpret_cast<const float*>(sBias + (it & 1) * kBiasStageBytes + lr * 2048);
            }
            // ---- chunk H: the h1 tile (128 rows x 64), produced ONCE by all 8 warps and used by steps 0 and 1.
            // Coalesced: warp w covers rows [16 w, 16 w + 16), one instruction = 2 rows x 256 B (4 lines instead of the
            // 32 a thread-per-row load touches); lane -> (row parity, 16-byte piece).
            {
                const long long tc0 = prof ? clock64() : 0;
                const uint32_t n = 5u * it, slot = n & 1u, u = n >> 1;
                const int piece = lane & 15;
                float4 x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t r = (uint32_t)(16 * warp + 2 * i + (lane >> 4));
                    const uint32_t grow = tile * kTileM + r;
                    x[i] = grow < p.n_rows ? __ldg(reinterpret_cast<const float4*>(p.h1 + (size_t)grow * 64) + piece)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                // h1 streams from HBM (1 GB per chunk of rays): pull this CTA's NEXT tile into L2 one tile time ahead
                if (!(p.debug_flags & 8u)) {
                    const uint32_t nrow = (tile + gridDim.x) * kTileM + (uint32_t)(threadIdx.x >> 1);
                    if (nrow < p.n_rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.h1 + (size_t)nrow * 64 + 32 * (threadIdx.x & 1)));
                }
                const long long tb = prof ? clock64() : 0;
                if (!mbar_wait(BAR(A_EMPTY0 + slot), (u & 1) ^ 1, p.dbg, 2, A_EMPTY0 + slot, it, 0)) goto teardown;
                long long tw = 0;
                if (prof) { tw = clock64() - tb; pw_aempty += tw; }
                uint8_t* sl = smem + kSmemA + slot * kASlotBytes;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = 16 * warp + 2 * i + (lane >> 4);
                    uint2 hi, lo;
                    split2(x[i].x * kActScale, x[i].y * kActScale, hi.x, lo.x);
                    split2(x[i].z * kActScale, x[i].w * kActScale, hi.y, lo.y);
                    // K elements [4 piece, 4 piece + 4): 16-byte unit piece / 2 (swizzled with the row), half piece % 2
                    uint8_t* dst = sl + (r >> 3) * 1024 + (r & 7) * 128 + (((piece >> 1) ^ (r & 7)) * 16) + (piece & 1) * 8;
                    *reinterpret_cast<uint2*>(dst) = hi;
                    *reinterpret_cast<uint2*>(dst + kATileBytes) = lo;
                }
                fence_proxy_async();
                mbar_arrive(BAR(A_FULL0 + slot));
                if (prof) pt_h1 += clock64() - tc0 - tw;
            }
            // the previous tile's accumulator drain runs now: the tensor pipe already has this tile's first step
            if (it > 0) {
                const long long te = prof ? clock64() : 0;
                if (!final_epilogue(prev_tile, it - 1)) goto teardown;
                if (prof) pw_epi += clock64() - te;
            }
            // ---- chunks a_j = relu(acc3[:, 64 j : 64 j + 64] k0 + bias0) for steps 2..5: both groups work on every
            // chunk, group g converts columns [64 j + 32 g, + 32) of its row (half the latency per chunk)
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const long long tc0 = prof ? clock64() : 0;
                long long tw = 0;
                const uint32_t n = 5u * it + 1u + (uint32_t)j, slot = n & 1u, u = n >> 1;
                if (j == 0) {
                    const long long ta = prof ? clock64() : 0;
                    if (!mbar_wait(BAR(ACC3_FULL), it & 1, p.dbg, 1, ACC3_FULL, it, j)) goto teardown;
                    if (smem_bias && !mbar_wait(BAR(BIAS_FULL0 + (it & 1)), (it >> 1) & 1, p.dbg, 12, BIAS_FULL0 + (it & 1), it, j)) goto teardown;
                    if (prof) { tw = clock64() - ta; pw_acc3 += tw; }
                    tc_fence_after();
                }
                uint32_t r0[32];
                tmem_ld32_issue(lane_taddr + (uint32_t)(64 * j + 32 * g), r0);
                const float4* b0 = reinterpret_cast<const float4*>(bias0 + 64 * j + 32 * g);
                float4 bb[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) bb[q] = smem_bias ? b0[q] : __ldg(b0 + q);  // overlaps the TMEM load
                tmem_ld_wait();
                if (j == 3) {  // this thread's last read of acc3
                    tc_fence_before();
                    mbar_arrive(BAR(ACC3_EMPTY));
                }
                const long long tb = prof ? clock64() : 0;
                if (!mbar_wait(BAR(A_EMPTY0 + slot), (u & 1) ^ 1, p.dbg, 2, A_EMPTY0 + slot, it, 2 + j)) goto teardown;
                if (prof) { const long long d = clock64() - tb; pw_aempty += d; tw += d; }
                const float k0 = p.k0;
                float v[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 bq = bb[q];
                    v[4 * q] = fmaxf(fmaf(__uint_as_float(r0[4 * q]), k0, bq.x), 0.f);
                    v[4 * q + 1] = fmaxf(fmaf(__uint_as_float(r0[4 * q + 1]), k0, bq.y), 0.f);
                    v[4 * q + 2] = fmaxf(fmaf(__uint_as_float(r0[4 * q + 2]), k0, bq.z), 0.f);
                    v[4 * q + 3] = fmaxf(fmaf(__uint_as_float(r0[4 * q + 3]), k0, bq.w), 0.f);
                }
                store_a_half(smem + kSmemA + slot * kASlotBytes, t, g, v);
                fence_proxy_async();
                mbar_arrive(BAR(A_FULL0 + slot));
                if (prof) pt_tm += clock64() - tc0 - tw;
            }
            prev_tile = tile;
        }
        if (it > 0) {
            if (!final_epilogue(prev_tile, it - 1)) goto teardown;
        }
        if (prof) {
            uint32_t* d = p.dbg + 16 + 8 * g;
            d[0] = (uint32_t)((clock64() - pt_begin) >> 10); d[1] = (uint32_t)(pw_aempty >> 10); d[2] = (uint32_t)(pw_acc3 >> 10);
            d[3] = (uint32_t)(pw_epi >> 10); d[4] = (uint32_t)(pt_h1 >> 10); d[5] = 0; d[6] = (uint32_t)(pt_tm >> 10);
            d[7] = 0;
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
                    if (s == 1 && !mbar_wait(BAR(ACC4_EMPTY), (it & 1) ^ 1, p.dbg, 5, ACC4_EMPTY, it, s)) goto teardown;
                    if (prof) { const long long t1 = clock64(); w_acc4 += t1 - t0; t0 = t1; }
                    if (!mbar_wait(BAR(B_FULL0 + slot), phase, p.dbg, 6, B_FULL0 + slot, it, s)) goto teardown;
                    if (prof) { const long long t1 = clock64(); w_b += t1 - t0; t0 = t1; }
                    // A chunks of a tile: n = 5 it (the h1 tile, steps 0 AND 1), 5 it + 1 .. 5 it + 4 (a_0..a_3, steps 2..5);
                    // chunk n lives in A slot n & 1 and is its (n >> 1)-th use
                    const uint32_t an = 5u * it + (s < 2 ? 0u : (uint32_t)(s - 1)), aslot = an & 1u;
                    if (s != 1 && !mbar_wait(BAR(A_FULL0 + aslot), (an >> 1) & 1, p.dbg, 7, A_FULL0 + aslot, it, s)) goto teardown;
                    if (prof) { const long long t1 = clock64(); w_a += t1 - t0; t0 = t1; }
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + kSmemA + aslot * kASlotBytes);
                    const uint32_t a_lo = a_hi + kATileBytes;
                    const uint32_t b_hi = smem_u32(smem + kSmemB + slot * kBSlotBytes);
                    const uint32_t b_lo = b_hi + kBTileBytes;
                    const uint32_t acc = tmem_base + (s == 0 ? 0u : 256u);
                    const bool init = s < 2;
#pragma unroll
                    for (int ks = 0; ks < kKC / 16; ++ks) {  // K = 16 per instruction = 32 bytes of the swizzled row
                        const uint64_t dah = make_desc(a_hi + ks * 32), dal = make_desc(a_lo + ks * 32);
                        const uint64_t dbh = make_desc(b_hi + ks * 32), dbl = make_desc(b_lo + ks * 32);
                        umma_f16(acc, dah, dbh, kIdesc, (init && ks == 0) ? 0u : 1u);
                        umma_f16(acc, dal, dbh, kIdesc, 1u);
                        umma_f16(acc, dah, dbl, kIdesc, 1u);
                    }
                    if (s != 0) umma_commit(BAR(A_EMPTY0 + aslot));   // the h1 tile is released after step 1
                    umma_commit(BAR(B_EMPTY0 + slot));
                    if (s == 0) umma_commit(BAR(ACC3_FULL));
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
                if (smem_bias) {
                    // bias rows of the rays this tile touches: one contiguous bulk copy into stage it & 1
                    const uint32_t st2 = it & 1;
                    if (!mbar_wait(BAR(BIAS_EMPTY0 + st2), ((it >> 1) & 1) ^ 1, p.dbg, 13, BIAS_EMPTY0 + st2, it, 0)) goto teardown;
                    const uint32_t f = (tile * kTileM) / (uint32_t)p.S;
                    uint32_t l = (tile * kTileM + kTileM - 1) / (uint32_t)p.S;
                    if (l >= n_rays_total) l = n_rays_total - 1;
                    uint32_t cnt = l - f + 1;
                    if (cnt > (uint32_t)kBiasRays) cnt = kBiasRays;
                    mbar_expect_tx(BAR(BIAS_FULL0 + st2), cnt * 2048);
                    bulk_g2s(smem_u32(sBias + st2 * kBiasStageBytes), p.dir_bias + (size_t)f * 512, cnt * 2048,
                             BAR(BIAS_FULL0 + st2));
                }
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

// Per-ray constant part of both colour layers: out[ray][0:256] = kActScale * (c0' + P0d direnc(viewdir)) (the scale of
// the FP16 A operand, see color_mlp_tc_kernel), out[ray][256:512] = c1' + P1d direnc(viewdir),
// direnc = pos_enc(viewdirs, 0, 4) (coord.py:L214-225, 27 values).  fp32 FMAs.
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
        out[(size_t)(ray0 + r) * 512 + n] = a0 * kActScale;  // exact (power of two)
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
    UC_ENSURE_SMEM(kSmemTotal, color_mlp_tc_kernel);
    const uint32_t ntiles = (p.n_rows + kTileM - 1) / kTileM;
    const uint32_t blocks = ntiles < (uint32_t)kNumSMs ? ntiles : (uint32_t)kNumSMs;
    color_mlp_tc_kernel<<<blocks, kThreads, kSmemTotal, st>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

uint32_t color_tc_blob_bytes() { return kSteps * kBSlotBytes; }
float color_tc_act_scale() { return kActScale; }

// Power-of-two factor that moves the largest |w| of a weight block into [2^13, 2^14): FP16 hi then carries 11
// significant bits and lo (|lo| <= 2^-11 |w s|) stays a normal FP16 number down to 2^-24 of the largest weight.
float color_tc_weight_scale(const float* w, size_t n) {
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(w[i]));
    if (!(mx > 0.f) || !std::isfinite(mx)) return 1.f;
    int e;
    frexpf(mx, &e);          // mx = f * 2^e, f in [0.5, 1)
    return ldexpf(1.f, 14 - e);
}

// Host: lay the folded weights out as the kernel consumes them.  Wt chunks are K-major [k][256] fp32 arrays of 64
// rows each (step order of the kernel); every chunk becomes hi tile | lo tile of (w * scale), each in the UMMA
// K-major SWIZZLE_128B image: element (n, k) at (n/8)*1024 + (n%8)*128 + ((k/8) ^ (n%8))*16 + (k%8)*2.
void color_tc_pack_chunk(const float* wt_rows /* [64][256] */, float scale, uint8_t* dst /* 64 KB */) {
    for (int n = 0; n < kN; ++n)
        for (int k = 0; k < kKC; ++k) {
            const float w = wt_rows[(size_t)k * kN + n] * scale;
            const __half h = __float2half_rn(w);
            const __half l = __float2half_rn(w - __half2float(h));
            const size_t off = (size_t)(n >> 3) * 1024 + (n & 7) * 128 + (((k >> 3) ^ (n & 7)) * 16) + (k & 7) * 2;
            memcpy(dst + off, &h, 2);
            memcpy(dst + kBTileBytes + off, &l, 2);
        }
}

}  // namespace ucnerf

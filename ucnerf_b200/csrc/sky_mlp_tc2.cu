// EXPERIMENT, not the default (env UCNERF_SKY_PIPELINE=2): it removes the layer-boundary bubble as designed (the MMA thread's
// wait for the first operand chunk drops from 18.6k to 5.3k cycles per tile) and is still 5 % SLOWER (161 vs 153 ms per
// 800x600 frame), because the N = 128 MMAs it needs retire in ~100 cycles instead of the 64 their FLOPs would take: an
// M128 x N128 x K16 instruction re-reads the 4 KB A block for every 128 output columns, 8 KB of operands per 64 cycles =
// the 128 B/clk shared-memory port, shared with the converters' stores (N = 256: 12 KB per 128 cycles).  Numbers:
// profiles/r2_sky_halfpass_experiment.txt.
//
// Sky head on tcgen05, second pipeline ("half passes"): same arithmetic as sky_mlp_tc.cu (3-term FP16 split, fp32 TMEM
// accumulators, the reference's NeRF.forward at 120 samples per ray, models.py:L797-820), reorganised so that the tensor pipe
// does not idle at the layer boundaries (sky_mlp_tc.cu: 24 % of the MMA thread's time, profiles/r1_sky_pipeline_waits.txt).
//
// In sky_mlp_tc.cu layer l+1 cannot start before ALL of layer l has retired and the first 64 columns have gone through
// tcgen05.ld -> scale / bias / relu -> FP16 split -> shared memory.  Here every 256-wide layer runs as TWO passes over its
// four K chunks: pass h accumulates output columns [128 h, 128 h + 128) (N = 128 MMAs, weight half-chunks of 32 KB).  While
// pass 1 of layer l occupies the tensor pipe the converters drain pass 0's columns into the first two K chunks of layer
// l + 1, so layer l + 1 starts the moment pass 1 retires, and its chunks 2 / 3 are converted under its own first two
// chunk-MMAs.  For that the WHOLE input of a layer stays resident: four A slots (128 KB), and the B ring shrinks to
// 2 x 32 KB - the same 192 KB.  A slot j holds K chunk j of the current layer; it is rewritten for the next layer as soon
// as pass 1 has consumed it.
//
// The two K = 3 steps (xyz x W0, and the xyz part of the skip layer 5) leave the tensor pipe: layer 0 is evaluated by the
// converters on CUDA cores in fp32 (3 FMAs per column) straight into layer 1's operand, and W5[:, 0:3] xyz is added as a
// per-row term when layer 5's accumulator is drained.  Steps per tile: 7 layers x 2 passes x 4 chunks + 4 (view layer,
// N = 128) = 60 half-steps of 12 MMAs (M128 x N128 x K16).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ray_march.cuh"
#include "tc_common.cuh"

namespace ucnerf {

namespace sky2 {

using namespace tc;

constexpr float kActScale = 8.f;
constexpr uint32_t kBHalfTile = 128 * kKC * 2;        // 16 KB: 128 output columns x 64 K (one of hi / lo)
constexpr uint32_t kBSlotBytes = 2 * kBHalfTile;      // 32 KB
constexpr uint32_t kSmemA = 0;                        // 4 slots x 32 KB
constexpr uint32_t kSmemB = kSmemA + 4 * kASlotBytes; // 131072
constexpr uint32_t kSmemMisc = kSmemB + 2 * kBSlotBytes;   // 196608
constexpr uint32_t kOffBar = 0;
constexpr uint32_t kOffTmem = 192;
constexpr uint32_t kOffBias = 256;                    // [8][256] floats (x kActScale)
constexpr uint32_t kOffWa = kOffBias + 8 * 1024;      // [256] alpha weights
constexpr uint32_t kOffRgbW = kOffWa + 1024;          // [128] float4
constexpr uint32_t kOffPart = kOffRgbW + 2048;        // [128] float4
constexpr uint32_t kOffW0 = kOffPart + 2048;          // [256] float4: kActScale * (W0[c][0..2], b0[c])
constexpr uint32_t kOffW5p = kOffW0 + 4096;           // [256] float4: kActScale * (W5[c][0..2], 0)
constexpr uint32_t kMiscBytes = kOffW5p + 4096;
constexpr uint32_t kSmemTotal = kSmemMisc + kMiscBytes + 1024;
static_assert(kSmemTotal <= 232448, "shared memory budget");
constexpr int kThreads = 320;
constexpr int kMmaWarp = 8;
constexpr int kHalfSteps = 60;                        // weight half-chunks per tile
constexpr uint32_t kIdesc128 = make_idesc(128);

enum Bar { A_FULL0 = 0, A_EMPTY0 = 4, B_FULL0 = 8, B_EMPTY0 = 10, ACC_FULL00 = 12 /* [acc][half] */, EPI_DONE = 16,
           PART_FULL, PART_EMPTY, NUM_BARS };
static_assert(NUM_BARS * 8 <= kOffTmem, "barrier block overlaps the TMEM pointer slot");

}  // namespace sky2

using namespace sky2;

__global__ void __launch_bounds__(kThreads, 1)
sky_mlp_tc2_kernel(const __grid_constant__ SkyTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* misc = smem + kSmemMisc;
    const uint32_t bar0 = smem_u32(misc + kOffBar);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(misc + kOffTmem);
    float* sBias = reinterpret_cast<float*>(misc + kOffBias);
    float* sWa = reinterpret_cast<float*>(misc + kOffWa);
    float4* sRgbW = reinterpret_cast<float4*>(misc + kOffRgbW);
    float4* sPart = reinterpret_cast<float4*>(misc + kOffPart);
    float4* sW0 = reinterpret_cast<float4*>(misc + kOffW0);
    float4* sW5p = reinterpret_cast<float4*>(misc + kOffW5p);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * 256; i += kThreads) sBias[i] = p.bias8[i];
    for (int i = threadIdx.x; i < 256; i += kThreads) {
        sWa[i] = p.w_alpha[i];
        sW0[i] = reinterpret_cast<const float4*>(p.w0x)[i];
        sW5p[i] = reinterpret_cast<const float4*>(p.w5x)[i];
    }
    for (int i = threadIdx.x; i < 128; i += kThreads) sRgbW[i] = reinterpret_cast<const float4*>(p.rgb_w)[i];
    if (threadIdx.x == 0) {
        for (int j = 0; j < 4; ++j) { mbar_init(BAR(A_FULL0 + j), 256); mbar_init(BAR(A_EMPTY0 + j), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(BAR(B_FULL0 + s), 1); mbar_init(BAR(B_EMPTY0 + s), 1); }
        for (int a = 0; a < 4; ++a) mbar_init(BAR(ACC_FULL00 + a), 1);
        mbar_init(BAR(EPI_DONE), 256);
        mbar_init(BAR(PART_FULL), 128); mbar_init(BAR(PART_EMPTY), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
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
    (void)lane;

    if (warp < 8) {
        // ================= converters / epilogue: thread <-> row t of the tile, columns 32 g .. of every chunk ==========
        const int g = warp >> 2;
        const int t = threadIdx.x & 127;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t it = 0, prev_tile = 0;
        float alpha_part = 0.f, alpha_prev = 0.f;
        uint32_t use[2][2] = {{0u, 0u}, {0u, 0u}};   // completed fills of each accumulator half this thread has consumed
        int epi_acc = 0;
        uint32_t epi_use = 0;
        float px = 0.f, py = 0.f, pz = 0.f;          // this row's sample position (kept for the skip layer)

        // A slot j is used once per layer: its u-th use (u = 8 it + layer index of the consumer - 1)
        auto wait_slot = [&](int j, uint32_t u, uint32_t step) -> bool {
            return mbar_wait(BAR(A_EMPTY0 + j), (u & 1u) ^ 1u, p.dbg, 2, A_EMPTY0 + j, it, step);
        };
        auto publish = [&](int j) {
            fence_proxy_async();
            tc_fence_before();   // this thread's TMEM reads precede the MMAs the arrival releases
            mbar_arrive(BAR(A_FULL0 + j));
        };
        // layer 0 on CUDA cores: h0 = relu(W0 p + b0) (x kActScale) -> the four K chunks of layer 1
        auto produce_h0 = [&](uint32_t tile) -> bool {
            const uint32_t row = tile * kTileM + t;
            px = py = pz = 0.f;
            if (row < p.n_rows) {
                const uint32_t ray = row / (uint32_t)p.n_samples, s = row - ray * (uint32_t)p.n_samples;
                const float tv = p.t_vals[s];
                // models.py:L872 (bug-compatible): z = near (1 - t) + (1 / far) t, near = the batch's far
                const float z = fa(fm(p.far[ray], fs(1.f, tv)), fm(fd(1.f, p.sky_far), tv));
                px = fa(p.origins[3 * (size_t)ray], fm(p.directions[3 * (size_t)ray], z));
                py = fa(p.origins[3 * (size_t)ray + 1], fm(p.directions[3 * (size_t)ray + 1], z));
                pz = fa(p.origins[3 * (size_t)ray + 2], fm(p.directions[3 * (size_t)ray + 2], z));
            }
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                float v[32];
                const float4* w = sW0 + 64 * j + 32 * g;
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float4 ww = w[q];
                    v[q] = fmaxf(fmaf(px, ww.x, fmaf(py, ww.y, fmaf(pz, ww.z, ww.w))), 0.f);
                }
                if (!wait_slot(j, 8u * it, (uint32_t)(100 + j))) return false;
                store_a_half(smem + kSmemA + j * kASlotBytes, t, g, v);
                publish(j);
            }
            return true;
        };
        // drain layer `layer`'s accumulator (two 128-column halves) into the four K chunks of layer `layer + 1`
        auto convert = [&](int layer, bool with_alpha, bool with_p) -> bool {
            const int acc = (layer + (int)it) & 1;
            const float k = p.k[layer];
            const float* bias = sBias + layer * 256;
            const uint32_t u = 8u * it + (uint32_t)layer;      // the consumer (layer + 1) is use index `layer` of this tile
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int h = j >> 1;
                if ((j & 1) == 0) {
                    if (!mbar_wait(BAR(ACC_FULL00 + 2 * acc + h), use[acc][h] & 1u, p.dbg, 1, ACC_FULL00 + 2 * acc + h, it,
                                   (uint32_t)layer))
                        return false;
                    use[acc][h] += 1;
                    tc_fence_after();
                }
                uint32_t r[32];
                tmem_ld32_issue(lane_taddr + (uint32_t)(256 * acc + 64 * j + 32 * g), r);
                const float4* b0 = reinterpret_cast<const float4*>(bias + 64 * j + 32 * g);
                float4 bb[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) bb[q] = b0[q];
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    v[4 * q] = fmaf(__uint_as_float(r[4 * q]), k, bb[q].x);
                    v[4 * q + 1] = fmaf(__uint_as_float(r[4 * q + 1]), k, bb[q].y);
                    v[4 * q + 2] = fmaf(__uint_as_float(r[4 * q + 2]), k, bb[q].z);
                    v[4 * q + 3] = fmaf(__uint_as_float(r[4 * q + 3]), k, bb[q].w);
                }
                if (with_p) {   // skip layer: + W5[:, 0:3] xyz (fp32, CUDA cores)
                    const float4* w = sW5p + 64 * j + 32 * g;
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float4 ww = w[q];
                        v[q] = fmaf(px, ww.x, fmaf(py, ww.y, fmaf(pz, ww.z, v[q])));
                    }
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
                if (with_alpha) {
                    const float* wa = sWa + 64 * j + 32 * g;
#pragma unroll
                    for (int q = 0; q < 32; ++q) alpha_part = fmaf(v[q], wa[q], alpha_part);
                }
                if (!wait_slot(j, u, (uint32_t)(10 * layer + j))) return false;
                store_a_half(smem + kSmemA + j * kASlotBytes, t, g, v);
                publish(j);
            }
            return true;
        };
        // drain the view layer (128 columns) of tile `tl`, rgb layer, hand the raw outputs out
        auto final_epilogue = [&](uint32_t tl, uint32_t itp) -> bool {
            if (!mbar_wait(BAR(ACC_FULL00 + 2 * epi_acc), epi_use & 1u, p.dbg, 3, ACC_FULL00 + 2 * epi_acc, itp, 99)) return false;
            tc_fence_after();
            uint32_t ra[32], rb[32];
            tmem_ld32_issue(lane_taddr + (uint32_t)(256 * epi_acc + 64 * g), ra);
            tmem_ld32_issue(lane_taddr + (uint32_t)(256 * epi_acc + 64 * g + 32), rb);
            const uint32_t row = tl * kTileM + t;
            const uint32_t ray = (row < p.n_rows ? row : 0u) / (uint32_t)p.n_samples;
            const float4* vb = reinterpret_cast<const float4*>(p.view_bias + (size_t)ray * 128 + 64 * g);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(BAR(EPI_DONE));
            float o0 = 0.f, o1 = 0.f, o2 = 0.f;
            const float k9 = p.k[8];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 b = __ldg(vb + q);
                const uint32_t* rr = q < 8 ? ra : rb;
                const int c = 4 * (q & 7);
                const float a0 = fmaxf(fmaf(__uint_as_float(rr[c]), k9, b.x), 0.f);
                const float a1 = fmaxf(fmaf(__uint_as_float(rr[c + 1]), k9, b.y), 0.f);
                const float a2 = fmaxf(fmaf(__uint_as_float(rr[c + 2]), k9, b.z), 0.f);
                const float a3 = fmaxf(fmaf(__uint_as_float(rr[c + 3]), k9, b.w), 0.f);
                const float4 w0 = sRgbW[64 * g + 4 * q], w1 = sRgbW[64 * g + 4 * q + 1], w2 = sRgbW[64 * g + 4 * q + 2],
                             w3 = sRgbW[64 * g + 4 * q + 3];
                o0 = fmaf(a3, w3.x, fmaf(a2, w2.x, fmaf(a1, w1.x, fmaf(a0, w0.x, o0))));
                o1 = fmaf(a3, w3.y, fmaf(a2, w2.y, fmaf(a1, w1.y, fmaf(a0, w0.y, o1))));
                o2 = fmaf(a3, w3.z, fmaf(a2, w2.z, fmaf(a1, w1.z, fmaf(a0, w0.z, o2))));
            }
            if (g == 1) {
                if (!mbar_wait(BAR(PART_EMPTY), (itp & 1) ^ 1, p.dbg, 9, PART_EMPTY, itp, 99)) return false;
                sPart[t] = make_float4(o0, o1, o2, alpha_prev);
                mbar_arrive(BAR(PART_FULL));
            } else {
                if (!mbar_wait(BAR(PART_FULL), itp & 1, p.dbg, 10, PART_FULL, itp, 99)) return false;
                const float4 q = sPart[t];
                mbar_arrive(BAR(PART_EMPTY));
                if (row < p.n_rows)
                    reinterpret_cast<float4*>(p.raw)[row] =
                        make_float4(o0 + q.x + p.rgb_b[0], o1 + q.y + p.rgb_b[1], o2 + q.z + p.rgb_b[2],
                                    (alpha_prev + q.w) * (1.f / kActScale) + p.b_alpha);
            }
            return true;
        };

        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            if (!produce_h0(tile)) goto teardown;              // layer 0 -> input of layer 1
            if (it > 0) {
                if (!final_epilogue(prev_tile, it - 1)) goto teardown;
            }
            if (!convert(1, false, false)) goto teardown;
            if (!convert(2, false, false)) goto teardown;
            if (!convert(3, false, false)) goto teardown;
            if (!convert(4, false, false)) goto teardown;      // h4 -> input of the skip layer (its xyz part is added below)
            if (!convert(5, false, true)) goto teardown;       // layer 5 = W5h h4 (tensor cores) + W5p xyz (here)
            if (!convert(6, false, false)) goto teardown;
            alpha_part = 0.f;
            if (!convert(7, true, false)) goto teardown;       // h7 -> view layer (+ alpha)
            alpha_prev = alpha_part;
            epi_acc = (8 + (int)it) & 1;                       // the view layer's accumulator, drained during the next tile
            epi_use = use[epi_acc][0];
            use[epi_acc][0] += 1;
            prev_tile = tile;
        }
        if (it > 0) {
            if (!final_epilogue(prev_tile, it - 1)) goto teardown;
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (one thread) ==================================================
        if (lane == 0) {
            uint32_t it = 0, m = 0;     // m: running weight half-chunk index (B ring)
            // profiling (debug_flags bit 2): cycles this thread waited per barrier kind, CTA 0 -> dbg[8..12]
            const bool prof = (p.debug_flags & 4u) != 0 && blockIdx.x == 0;
            long long w_epi = 0, w_b = 0, w_a = 0;
            const long long t_begin = clock64();
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
#pragma unroll 1
                for (int layer = 1; layer <= 8; ++layer) {
                    const uint32_t acc_i = (uint32_t)((layer + (int)it) & 1);
                    const int passes = layer == 8 ? 1 : 2;
                    const uint32_t u = 8u * it + (uint32_t)(layer - 1);      // use index of the A slots for this layer
                    // layer 1 writes the accumulator the previous tile's view layer is drained from
                    long long t0 = prof ? clock64() : 0;
                    if (layer == 1 && !mbar_wait(BAR(EPI_DONE), (it & 1) ^ 1, p.dbg, 5, EPI_DONE, it, 0)) goto teardown;
                    if (prof) w_epi += clock64() - t0;
#pragma unroll 1
                    for (int h = 0; h < passes; ++h) {
                        const uint32_t acc = tmem_base + 256u * acc_i + 128u * (uint32_t)h;
#pragma unroll 1
                        for (int j = 0; j < 4; ++j, ++m) {
                            const uint32_t slot = m & 1u, ph = (m >> 1) & 1u;
                            t0 = prof ? clock64() : 0;
                            if (!mbar_wait(BAR(B_FULL0 + slot), ph, p.dbg, 6, B_FULL0 + slot, it, (uint32_t)(10 * layer + j))) goto teardown;
                            if (prof) { const long long t1 = clock64(); w_b += t1 - t0; t0 = t1; }
                            if (h == 0 && !mbar_wait(BAR(A_FULL0 + j), u & 1u, p.dbg, 7, A_FULL0 + j, it, (uint32_t)(10 * layer + j)))
                                goto teardown;
                            if (prof) w_a += clock64() - t0;
                            tc_fence_after();
                            const uint32_t a_hi = smem_u32(smem + kSmemA + j * kASlotBytes);
                            const uint32_t a_lo = a_hi + kATileBytes;
                            const uint32_t b_hi = smem_u32(smem + kSmemB + slot * kBSlotBytes);
                            const uint32_t b_lo = b_hi + kBHalfTile;
#pragma unroll
                            for (int ks = 0; ks < kKC / 16; ++ks) {
                                const uint64_t dah = make_desc(a_hi + ks * 32), dal = make_desc(a_lo + ks * 32);
                                const uint64_t dbh = make_desc(b_hi + ks * 32), dbl = make_desc(b_lo + ks * 32);
                                umma_f16(acc, dah, dbh, kIdesc128, (j == 0 && ks == 0) ? 0u : 1u);
                                umma_f16(acc, dal, dbh, kIdesc128, 1u);
                                umma_f16(acc, dah, dbl, kIdesc128, 1u);
                            }
                            umma_commit(BAR(B_EMPTY0 + slot));
                            if (h == passes - 1) umma_commit(BAR(A_EMPTY0 + j));
                        }
                        umma_commit(BAR(ACC_FULL00 + 2 * (int)acc_i + h));
                    }
                }
            }
            if (prof) {
                p.dbg[8] = (uint32_t)((clock64() - t_begin) >> 10); p.dbg[9] = (uint32_t)(w_epi >> 10);
                p.dbg[10] = (uint32_t)(w_b >> 10); p.dbg[11] = (uint32_t)(w_a >> 10); p.dbg[12] = it;
            }
        }
    } else {
        // ================= weight loader (one thread) ==============================================
        if (lane == 0) {
            uint32_t it = 0, m = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
#pragma unroll 1
                for (int s = 0; s < kHalfSteps; ++s, ++m) {
                    const uint32_t slot = m & 1u, ph = (m >> 1) & 1u;
                    if (!mbar_wait(BAR(B_EMPTY0 + slot), ph ^ 1u, p.dbg, 8, B_EMPTY0 + slot, it, (uint32_t)s)) goto teardown;
                    mbar_expect_tx(BAR(B_FULL0 + slot), kBSlotBytes);
                    bulk_g2s(smem_u32(smem + kSmemB + slot * kBSlotBytes), p.wblob2 + (size_t)s * kBSlotBytes, kBSlotBytes,
                             BAR(B_FULL0 + slot));
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

int launch_sky_mlp_tc2(const SkyTcParams& p_in, uint32_t* dbg, cudaStream_t st) {
    if (p_in.n_rows == 0) return 0;
    SkyTcParams p = p_in;
    p.dbg = dbg;
    if (const char* e = getenv("UCNERF_SKY_DEBUG")) p.debug_flags = (uint32_t)atoi(e);   // profiling experiments only
    UC_ENSURE_SMEM(kSmemTotal, sky_mlp_tc2_kernel);
    const uint32_t ntiles = (p.n_rows + kTileM - 1) / kTileM;
    const uint32_t blocks = ntiles < (uint32_t)kNumSMs ? ntiles : (uint32_t)kNumSMs;
    sky_mlp_tc2_kernel<<<blocks, kThreads, kSmemTotal, st>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

uint32_t sky_tc2_blob_bytes() { return (uint32_t)kHalfSteps * kBSlotBytes; }
int sky_tc2_half_steps() { return kHalfSteps; }

}  // namespace ucnerf

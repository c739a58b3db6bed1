// Sky head on the 5th-generation tensor cores: the 8x256 NeRF MLP the reference evaluates
// at 120 samples per ray behind every pixel (models.py:L84-92 skynerf, L743-820 NeRF.forward, L849-904 render_rays,
// L822-847 raw2outputs): 562,688 MAC per sample, 67.5 M MAC per ray - 9x the whole fused foreground path.
//
//   h0 = relu(W0 p + b0)                  p = o + d z (raw xyz, K = 3)
//   h_l = relu(W_l h_{l-1} + b_l)         l = 1..7, 256 wide; layer 5 takes [p, h4] (skip connection, K = 259)
//   alpha = w_a . h7 + b_a                (CUDA cores, folded into the conversion of h7)
//   f = W_f h7 + b_f                      (no activation)            } folded on the host (fp64): v = relu(W' h7 + b'(ray)),
//   v = relu(W_v [f, emb(view)] + b_v)    128 wide                   } W' = W_vf W_f, b' = b_v + W_vf b_f + W_vv emb(view)
//   rgb = W_r v + b_r                     (CUDA cores in the epilogue)
//
// Same machinery as color_mlp_tc.cu (tc_common.cuh): 128-row tiles, persistent CTAs, two producer / epilogue
// warpgroups + one MMA issuer + one weight loader, tcgen05.mma kind::f16 with fp32 accumulators in TMEM, every operand
// as an FP16 hi + lo pair (3-term split: fp32-level accuracy), weights pre-scaled / pre-split / pre-swizzled on the
// host and streamed L2 -> smem with cp.async.bulk.  34 K-chunks (= MMA steps = weight chunks) per tile:
//   step 0        p      x W0          -> layer 0         steps 17..21  [p, h4] x W5   -> layer 5
//   steps 1..4    h0     x W1          -> layer 1         steps 22..25  h5 x W6        -> layer 6
//   steps 5..8    h1     x W2          -> layer 2         steps 26..29  h6 x W7        -> layer 7
//   steps 9..12   h2     x W3          -> layer 3         steps 30..33  h7 x W' (N=128)-> layer 8 (view layer)
//   steps 13..16  h3     x W4          -> layer 4
// Layer l of the it-th tile of a CTA accumulates into TMEM accumulator (l + it) & 1 (nine layers per tile, so the
// accumulator the view layer of one tile is drained from is not the one layer 0 of the next tile writes).
// The two TMEM accumulators (256 columns each) ping-pong: while layer l accumulates into one, both warpgroups drain
// the other (tcgen05.ld -> scale + bias -> relu -> FP16 split -> swizzled smem) chunk by chunk into the A ring that
// feeds layer l.  The per-sample outputs (rgb_raw, alpha_raw) go to HBM; sky_composite_kernel integrates them per ray
// exactly as raw2outputs does (including the reference's decreasing sample depths, models.py:L872).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ray_march.cuh"
#include "tc_common.cuh"

namespace ucnerf {

namespace sky {

using namespace tc;

constexpr int kN = 256;
constexpr float kActScale = 8.f;
constexpr uint32_t kBTileBytes = kN * kKC * 2;       // 32 KB
constexpr uint32_t kBSlotBytes = 2 * kBTileBytes;    // 64 KB
constexpr uint32_t kSmemA = 0;
constexpr uint32_t kSmemB = kSmemA + 2 * kASlotBytes;            // 65536
constexpr uint32_t kSmemMisc = kSmemB + 2 * kBSlotBytes;         // 196608
constexpr uint32_t kOffBar = 0;
constexpr uint32_t kOffTmem = 192;
constexpr int kStepsPerTile = 34;
constexpr uint32_t kOffBias = 256;                                // [8][256] floats (already x kActScale)
constexpr uint32_t kOffWa = kOffBias + 8 * 1024;                  // [256] alpha weights
constexpr uint32_t kOffRgbW = kOffWa + 1024;                      // [128] float4 (rgb weights, w = 0)
constexpr uint32_t kOffPart = kOffRgbW + 2048;                    // [128] float4 partial sums group 1 -> group 0
constexpr uint32_t kMiscBytes = kOffPart + 2048;
constexpr uint32_t kSmemTotal = kSmemMisc + kMiscBytes + 1024;    // +1024: manual 1 KB alignment slack
static_assert(kSmemTotal <= 232448, "shared memory budget");
constexpr int kThreads = 320;
constexpr int kMmaWarp = 8;
constexpr uint32_t kIdesc256 = make_idesc(256), kIdesc128 = make_idesc(128);

enum Bar { A_FULL0 = 0, A_FULL1, A_EMPTY0, A_EMPTY1, B_FULL0, B_FULL1, B_EMPTY0, B_EMPTY1, ACC_FULL0, ACC_FULL1,
           EPI_DONE, PART_FULL, PART_EMPTY, NUM_BARS };
static_assert(NUM_BARS * 8 <= kOffTmem, "barrier block overlaps the TMEM pointer slot");

}  // namespace sky

using namespace sky;

__global__ void __launch_bounds__(kThreads, 1)
sky_mlp_tc_kernel(const __grid_constant__ SkyTcParams p) {
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * 256; i += kThreads) sBias[i] = p.bias8[i];
    for (int i = threadIdx.x; i < 256; i += kThreads) sWa[i] = p.w_alpha[i];
    for (int i = threadIdx.x; i < 128; i += kThreads) sRgbW[i] = reinterpret_cast<const float4*>(p.rgb_w)[i];
    if (threadIdx.x == 0) {
        mbar_init(BAR(A_FULL0), 256); mbar_init(BAR(A_FULL1), 256);
        mbar_init(BAR(A_EMPTY0), 1); mbar_init(BAR(A_EMPTY1), 1);
        mbar_init(BAR(B_FULL0), 1); mbar_init(BAR(B_FULL1), 1);
        mbar_init(BAR(B_EMPTY0), 1); mbar_init(BAR(B_EMPTY1), 1);
        mbar_init(BAR(ACC_FULL0), 1); mbar_init(BAR(ACC_FULL1), 1);
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
        // ================= producer / epilogue warpgroups: thread <-> row t of the tile ================
        const int g = warp >> 2;
        const int t = threadIdx.x & 127;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t it = 0, prev_tile = 0;
        float alpha_part = 0.f, alpha_prev = 0.f;
        // profiling (debug_flags bit 2): thread 0 of group 0 in CTA 0 -> dbg[16..]: total, wait ACC_FULL, wait A_EMPTY, epilogue
        const bool prof = (p.debug_flags & 4u) != 0 && t == 0 && g == 0 && blockIdx.x == 0;
        long long pw_acc = 0, pw_slot = 0, pw_epi = 0;
        const long long pt_begin = clock64();
        uint32_t use[2] = {0u, 0u};          // completed fills of each accumulator this thread has consumed
        int epi_acc = 0;
        uint32_t epi_use = 0;                // accumulator / fill index of the previous tile's view layer

        // A chunk n of the running sequence lives in slot n & 1 and is its (n >> 1)-th use
        auto wait_slot = [&](uint32_t n, uint32_t step) -> bool {
            const long long t0 = prof ? clock64() : 0;
            const bool ok = mbar_wait(BAR(A_EMPTY0 + (n & 1u)), ((n >> 1) & 1u) ^ 1u, p.dbg, 2, A_EMPTY0 + (n & 1u), it, step);
            if (prof) pw_slot += clock64() - t0;
            return ok;
        };
        auto publish = [&](uint32_t n) {
            fence_proxy_async();
            tc_fence_before();   // this thread's TMEM reads precede the MMAs the arrival releases
            mbar_arrive(BAR(A_FULL0 + (n & 1u)));
        };
        // the xyz chunk (K = 16 used: [8 x, 8 y, 8 z, 0 ...]): only group 0 has something to write
        auto produce_p = [&](uint32_t tile, uint32_t n) -> bool {
            float px = 0.f, py = 0.f, pz = 0.f;
            const uint32_t row = tile * kTileM + t;
            if (g == 0 && row < p.n_rows) {
                const uint32_t ray = row / (uint32_t)p.n_samples, s = row - ray * (uint32_t)p.n_samples;
                const float tv = p.t_vals[s];
                // models.py:L872 (bug-compatible): z = near (1 - t) + (1 / far) t, near = the batch's far
                const float z = fa(fm(p.far[ray], fs(1.f, tv)), fm(fd(1.f, p.sky_far), tv));
                px = fa(p.origins[3 * (size_t)ray], fm(p.directions[3 * (size_t)ray], z));
                py = fa(p.origins[3 * (size_t)ray + 1], fm(p.directions[3 * (size_t)ray + 1], z));
                pz = fa(p.origins[3 * (size_t)ray + 2], fm(p.directions[3 * (size_t)ray + 2], z));
            }
            if (!wait_slot(n, 100)) return false;
            if (g == 0) {
                uint8_t* sl = smem + kSmemA + (n & 1u) * kASlotBytes;
                uint8_t* base = sl + (t >> 3) * 1024 + (t & 7) * 128;
                uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
                split2(px * kActScale, py * kActScale, hi.x, lo.x);
                split2(pz * kActScale, 0.f, hi.y, lo.y);
                const uint4 zero = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(base + ((0 ^ (t & 7)) * 16)) = hi;            // K elements 0..7
                *reinterpret_cast<uint4*>(base + ((1 ^ (t & 7)) * 16)) = zero;          // K elements 8..15
                *reinterpret_cast<uint4*>(base + kATileBytes + ((0 ^ (t & 7)) * 16)) = lo;
                *reinterpret_cast<uint4*>(base + kATileBytes + ((1 ^ (t & 7)) * 16)) = zero;
            }
            publish(n);
            return true;
        };
        // drain one accumulator into four A chunks: v = act(acc k + bias8[col]); thread = (row t, 32 columns of each chunk)
        auto convert = [&](int layer, bool with_alpha, uint32_t n_first) -> bool {
            const int acc = (layer + (int)it) & 1;
            const long long tw0 = prof ? clock64() : 0;
            if (!mbar_wait(BAR(ACC_FULL0 + acc), use[acc] & 1u, p.dbg, 1, ACC_FULL0 + acc, it, (uint32_t)layer)) return false;
            if (prof) pw_acc += clock64() - tw0;
            use[acc] += 1;
            tc_fence_after();
            constexpr bool relu = true;
            const float k = p.k[layer];
            const float* bias = sBias + layer * 256;
            // software pipeline over the four chunks: the TMEM load of chunk j + 1 is in flight while chunk j is converted
            uint32_t ra[32], rb[32];
            tmem_ld32_issue(lane_taddr + (uint32_t)(256 * acc + 32 * g), ra);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t (&r0)[32] = (j & 1) ? rb : ra;
                uint32_t (&rn)[32] = (j & 1) ? ra : rb;
                const float4* b0 = reinterpret_cast<const float4*>(bias + 64 * j + 32 * g);
                float4 bb[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) bb[q] = b0[q];
                tmem_ld_wait();
                if (j < 3) tmem_ld32_issue(lane_taddr + (uint32_t)(256 * acc + 64 * (j + 1) + 32 * g), rn);
                if (!wait_slot(n_first + (uint32_t)j, (uint32_t)(10 * layer + j))) return false;
                float v[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    v[4 * q] = fmaf(__uint_as_float(r0[4 * q]), k, bb[q].x);
                    v[4 * q + 1] = fmaf(__uint_as_float(r0[4 * q + 1]), k, bb[q].y);
                    v[4 * q + 2] = fmaf(__uint_as_float(r0[4 * q + 2]), k, bb[q].z);
                    v[4 * q + 3] = fmaf(__uint_as_float(r0[4 * q + 3]), k, bb[q].w);
                }
                if (relu) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
                }
                if (with_alpha) {
                    const float* wa = sWa + 64 * j + 32 * g;
#pragma unroll
                    for (int q = 0; q < 32; ++q) alpha_part = fmaf(v[q], wa[q], alpha_part);
                }
                store_a_half(smem + kSmemA + ((n_first + (uint32_t)j) & 1u) * kASlotBytes, t, g, v);
                publish(n_first + (uint32_t)j);
            }
            return true;
        };
        // drain the view layer (acc1, 128 columns) of tile `tl`, rgb layer, hand the raw outputs out
        auto final_epilogue = [&](uint32_t tl, uint32_t itp) -> bool {
            if (!mbar_wait(BAR(ACC_FULL0 + epi_acc), epi_use & 1u, p.dbg, 3, ACC_FULL0 + epi_acc, itp, 99)) return false;
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
            const uint32_t n0 = (uint32_t)kStepsPerTile * it;
            if (!produce_p(tile, n0)) goto teardown;
            if (it > 0) {
                const long long te = prof ? clock64() : 0;
                if (!final_epilogue(prev_tile, it - 1)) goto teardown;
                if (prof) pw_epi += clock64() - te;
            }
            if (!convert(0, false, n0 + 1)) goto teardown;    // h0 -> layer 1
            if (!convert(1, false, n0 + 5)) goto teardown;    // h1 -> layer 2
            if (!convert(2, false, n0 + 9)) goto teardown;    // h2 -> layer 3
            if (!convert(3, false, n0 + 13)) goto teardown;   // h3 -> layer 4
            if (!produce_p(tile, n0 + 17)) goto teardown;     // xyz again: skip connection
            if (!convert(4, false, n0 + 18)) goto teardown;   // h4 -> layer 5
            if (!convert(5, false, n0 + 22)) goto teardown;   // h5 -> layer 6
            if (!convert(6, false, n0 + 26)) goto teardown;   // h6 -> layer 7
            alpha_part = 0.f;
            if (!convert(7, true, n0 + 30)) goto teardown;    // h7 -> view layer (+ alpha)
            alpha_prev = alpha_part;
            epi_acc = (8 + (int)it) & 1;                      // the view layer's accumulator, drained during the next tile
            epi_use = use[epi_acc];
            use[epi_acc] += 1;
            prev_tile = tile;
        }
        if (it > 0) {
            if (!final_epilogue(prev_tile, it - 1)) goto teardown;
        }
        if (prof) {
            p.dbg[16] = (uint32_t)((clock64() - pt_begin) >> 10); p.dbg[17] = (uint32_t)(pw_acc >> 10);
            p.dbg[18] = (uint32_t)(pw_slot >> 10); p.dbg[19] = (uint32_t)(pw_epi >> 10); p.dbg[20] = it;
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (one thread) ==================================================
        if (lane == 0) {
            uint32_t it = 0;
            // profiling (debug_flags bit 2): cycles this thread waited per barrier kind, CTA 0 -> dbg[8..13]
            const bool prof = (p.debug_flags & 4u) != 0 && blockIdx.x == 0;
            long long w_epi = 0, w_b = 0, w_a = 0;
            const long long t_begin = clock64();
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
#pragma unroll 1
                for (int s = 0; s < kStepsPerTile; ++s) {
                    // layer of this step, whether it opens / closes the layer, and its accumulator
                    int layer, first, last;
                    if (s == 0) { layer = 0; first = 1; last = 1; }
                    else if (s <= 16) { layer = 1 + (s - 1) / 4; first = ((s - 1) & 3) == 0; last = ((s - 1) & 3) == 3; }
                    else if (s <= 21) { layer = 5; first = s == 17; last = s == 21; }
                    else { layer = 6 + (s - 22) / 4; first = ((s - 22) & 3) == 0; last = ((s - 22) & 3) == 3; }
                    const uint32_t acc_i = (uint32_t)((layer + (int)it) & 1);
                    const bool p_step = (s == 0 || s == 17);
                    // layer 1 writes the accumulator the previous tile's view layer is drained from
                    long long t0 = prof ? clock64() : 0;
                    if (s == 1 && !mbar_wait(BAR(EPI_DONE), (it & 1) ^ 1, p.dbg, 5, EPI_DONE, it, s)) goto teardown;
                    if (prof) { const long long t1 = clock64(); w_epi += t1 - t0; t0 = t1; }
                    const uint32_t m = (uint32_t)kStepsPerTile * it + (uint32_t)s, slot = m & 1u, ph = (m >> 1) & 1u;
                    if (!mbar_wait(BAR(B_FULL0 + slot), ph, p.dbg, 6, B_FULL0 + slot, it, s)) goto teardown;
                    if (prof) { const long long t1 = clock64(); w_b += t1 - t0; t0 = t1; }
                    if (!mbar_wait(BAR(A_FULL0 + slot), ph, p.dbg, 7, A_FULL0 + slot, it, s)) goto teardown;
                    if (prof) { const long long t1 = clock64(); w_a += t1 - t0; t0 = t1; }
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + kSmemA + slot * kASlotBytes);
                    const uint32_t a_lo = a_hi + kATileBytes;
                    const uint32_t b_hi = smem_u32(smem + kSmemB + slot * kBSlotBytes);
                    const uint32_t b_lo = b_hi + (layer == 8 ? kBTileBytes / 2 : kBTileBytes);
                    const uint32_t acc = tmem_base + 256u * acc_i;
                    const uint32_t idesc = layer == 8 ? kIdesc128 : kIdesc256;
                    const int nks = p_step ? 1 : kKC / 16;
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint64_t dah = make_desc(a_hi + ks * 32), dal = make_desc(a_lo + ks * 32);
                        const uint64_t dbh = make_desc(b_hi + ks * 32), dbl = make_desc(b_lo + ks * 32);
                        umma_f16(acc, dah, dbh, idesc, (first && ks == 0) ? 0u : 1u);
                        umma_f16(acc, dal, dbh, idesc, 1u);
                        umma_f16(acc, dah, dbl, idesc, 1u);
                    }
                    umma_commit(BAR(A_EMPTY0 + slot));
                    umma_commit(BAR(B_EMPTY0 + slot));
                    if (last) umma_commit(BAR(ACC_FULL0 + acc_i));
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
            uint32_t it = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
#pragma unroll 1
                for (int s = 0; s < kStepsPerTile; ++s) {
                    const uint32_t m = (uint32_t)kStepsPerTile * it + (uint32_t)s, slot = m & 1u, ph = (m >> 1) & 1u;
                    if (!mbar_wait(BAR(B_EMPTY0 + slot), ph ^ 1u, p.dbg, 8, B_EMPTY0 + slot, it, s)) goto teardown;
                    const uint32_t bytes = s >= 30 ? kBSlotBytes / 2 : kBSlotBytes;
                    mbar_expect_tx(BAR(B_FULL0 + slot), bytes);
                    bulk_g2s(smem_u32(smem + kSmemB + slot * kBSlotBytes), p.wblob + (size_t)s * kBSlotBytes, bytes,
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

// Per-ray constant part of the view layer: out[ray][0:128] = b_v + W_v[:, 256:283] emb(view), emb = [x, sin(x f), cos(x f)]
// for f = 1, 2, 4, 8 (models.py:L689-727 Embedder, get_embedder(4)); `view` is the ray's cam_dirs (models.py:L331).
__global__ void __launch_bounds__(128)
sky_view_bias_kernel(const float* __restrict__ views, const float* __restrict__ wv_view /* [27][128] */,
                     const float* __restrict__ bv, float* __restrict__ out, uint32_t n_rays) {
    __shared__ float enc[8][28];
    const uint32_t ray0 = blockIdx.x * 8;
    for (int i = threadIdx.x; i < 8 * 27; i += 128) {
        const int r = i / 27, c = i - 27 * r;
        float val = 0.f;
        if (ray0 + r < n_rays) {
            const float* vd = views + 3 * (size_t)(ray0 + r);
            if (c < 3) val = vd[c];
            else {
                const int q = c - 3, f = q / 6, w = q - 6 * f;      // per frequency: sin(x f) [3], cos(x f) [3]
                const float x = fm(vd[w % 3], (float)(1 << f));
                val = w < 3 ? sinf(x) : cosf(x);
            }
        }
        enc[r][c] = val;
    }
    __syncthreads();
    const int n = threadIdx.x;
    float w[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) w[k] = __ldg(wv_view + (size_t)k * 128 + n);
    const float b = bv[n];
    for (int r = 0; r < 8 && ray0 + r < n_rays; ++r) {
        float a = b;
#pragma unroll
        for (int k = 0; k < 27; ++k) a = fmaf(w[k], enc[r][k], a);
        out[(size_t)(ray0 + r) * 128 + n] = a;
    }
}

// raw2outputs (models.py:L822-847) for one ray per thread: sample depths z (L868-875), dists (last = 1e10) x |d|,
// rgb = sigmoid(raw), alpha = 1 - exp(-relu(raw_a) dists), w = alpha cumprod(1 - alpha + 1e-10), rgb_map = sum w rgb.
__global__ void __launch_bounds__(128)
sky_composite_kernel(const float4* __restrict__ raw, const float* __restrict__ directions, const float* __restrict__ far,
                     const float* __restrict__ t_vals, float sky_far, int n_samples, float* __restrict__ out, uint32_t n_rays) {
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const float dn = norm3(directions[3 * (size_t)ray], directions[3 * (size_t)ray + 1], directions[3 * (size_t)ray + 2]);
    const float near = far[ray], inv_far = fd(1.f, sky_far);
    double T = 1.0, r = 0.0, g = 0.0, b = 0.0;   // torch's CPU cumprod / sum accumulate fp32 inputs in fp64
    float z = fa(fm(near, fs(1.f, t_vals[0])), fm(inv_far, t_vals[0]));
    for (int s = 0; s < n_samples; ++s) {
        float dist;
        if (s + 1 < n_samples) {
            const float zn = fa(fm(near, fs(1.f, t_vals[s + 1])), fm(inv_far, t_vals[s + 1]));
            dist = fs(zn, z);
            z = zn;
        } else {
            dist = 1e10f;
        }
        dist = fm(dist, dn);
        const float4 v = raw[(size_t)ray * n_samples + s];
        const float alpha = fs(1.f, expf(-fm(fmaxf(v.w, 0.f), dist)));
        const float w = fm(alpha, (float)T);
        r += fm(w, sigmoid_f(v.x)); g += fm(w, sigmoid_f(v.y)); b += fm(w, sigmoid_f(v.z));
        T *= (double)fa(fs(1.f, alpha), 1e-10f);
    }
    out[3 * (size_t)ray] = (float)r; out[3 * (size_t)ray + 1] = (float)g; out[3 * (size_t)ray + 2] = (float)b;
}

static uint32_t* g_sky_dbg = nullptr;

int sky_tc_status(uint32_t* out32) {
    for (int i = 0; i < 32; ++i) out32[i] = 0;
    if (!g_sky_dbg) return 0;
    UC_CUDA_OK(cudaDeviceSynchronize());
    UC_CUDA_OK(cudaMemcpy(out32, g_sky_dbg, 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

uint32_t* sky_tc_dbg_buffer() {
    if (!g_sky_dbg) {
        if (cudaMalloc(&g_sky_dbg, 32 * sizeof(uint32_t)) != cudaSuccess) return nullptr;
        cudaMemset(g_sky_dbg, 0, 32 * sizeof(uint32_t));
    }
    return g_sky_dbg;
}

int launch_sky_mlp_tc(const SkyTcParams& p_in, cudaStream_t st) {
    if (p_in.n_rows == 0) return 0;
    UC_REQUIRE(sky_tc_dbg_buffer() != nullptr, "sky: cannot allocate the watchdog record");
    SkyTcParams p = p_in;
    p.dbg = g_sky_dbg;
    if (const char* e = getenv("UCNERF_SKY_DEBUG")) p.debug_flags = (uint32_t)atoi(e);   // profiling experiments only
    UC_ENSURE_SMEM(kSmemTotal, sky_mlp_tc_kernel);
    const uint32_t ntiles = (p.n_rows + kTileM - 1) / kTileM;
    const uint32_t blocks = ntiles < (uint32_t)kNumSMs ? ntiles : (uint32_t)kNumSMs;
    sky_mlp_tc_kernel<<<blocks, kThreads, kSmemTotal, st>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

int launch_sky_view_bias(const float* views, const float* wv_view, const float* bv, float* out, uint32_t n_rays, cudaStream_t st) {
    if (n_rays == 0) return 0;
    sky_view_bias_kernel<<<div_up(n_rays, 8u), 128, 0, st>>>(views, wv_view, bv, out, n_rays);
    UC_LAUNCH_CHECK();
    return 0;
}

int launch_sky_composite(const float* raw, const float* directions, const float* far, const float* t_vals, float sky_far,
                         int n_samples, float* out, uint32_t n_rays, cudaStream_t st) {
    if (n_rays == 0) return 0;
    sky_composite_kernel<<<div_up(n_rays, 128u), 128, 0, st>>>(reinterpret_cast<const float4*>(raw), directions, far, t_vals,
                                                                sky_far, n_samples, out, n_rays);
    UC_LAUNCH_CHECK();
    return 0;
}

uint32_t sky_tc_blob_bytes() { return (uint32_t)kStepsPerTile * kBSlotBytes; }
int sky_tc_steps() { return kStepsPerTile; }
float sky_tc_act_scale() { return kActScale; }

// Host: one K-major [64][n_cols] fp32 block -> hi tile | lo tile of (w * scale) in the UMMA K-major SWIZZLE_128B image
// (element (n, k) at (n/8)*1024 + (n%8)*128 + ((k/8) ^ (n%8))*16 + (k%8)*2); lo tile follows the hi tile.
void sky_tc_pack_chunk(const float* wt_rows, int n_cols, float scale, uint8_t* dst) {
    const size_t tile = (size_t)n_cols * kKC * 2;
    for (int n = 0; n < n_cols; ++n)
        for (int k = 0; k < kKC; ++k) {
            const float w = wt_rows[(size_t)k * n_cols + n] * scale;
            const __half h = __float2half_rn(w);
            const __half l = __float2half_rn(w - __half2float(h));
            const size_t off = (size_t)(n >> 3) * 1024 + (n & 7) * 128 + (((k >> 3) ^ (n & 7)) * 16) + (k & 7) * 2;
            memcpy(dst + off, &h, 2);
            memcpy(dst + tile + off, &l, 2);
        }
}

}  // namespace ucnerf

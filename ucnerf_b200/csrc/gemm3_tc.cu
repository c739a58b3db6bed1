// fp32-accurate GEMMs on the 5th-generation tensor cores (tcgen05 + TMEM) for the TRAINING step's dense layers
// (internal/models.py:L438-441 density_layer, L475-483 / L643-652 view-dependent colour layers, and their backward):
//
//   NT   C[M,N]   = sum_s A_s[M,K_s] B_s[N,K_s]^T (+ bias[N]) (relu)      forward  y = x W^T + b   and   dx = dy W
//   TN   C[N1,N2] += A[M,N1]^T B[M,N2]                                     weight gradient  dW = dy^T x   (reduction over the rows)
//
// replacing the cuBLAS fp32 SIMT SGEMMs that `nn.Linear` runs for the reference (no TF32 there, so a single low-precision
// pass would not be the same computation).  fp32 accuracy comes from the 3xTF32 split: every operand is written as
// x = hi + lo with hi, lo representable in TF32 (fp32 exponent range, so gradients of any magnitude need no scaling) and
// hi*hi + lo*hi + hi*lo is accumulated in the fp32 TMEM accumulator: ~2^-21 relative error per product.
//
// One CTA per SM, 13 warps:
//   warps 0-7   loaders: read fp32 operand tiles from global memory (coalesced; the TN form reads 4 rows x 32 B per
//               instruction and transposes on the way), split them and store hi / lo tiles in shared memory in the UMMA
//               canonical K-major SWIZZLE_128B layout (32 TF32 per 128-byte row) - bank-conflict-free mappings
//   warps 8-11  epilogue: tcgen05.ld the accumulator (warp w owns TMEM lanes 32 (w % 4) ..), + bias, relu, store / red.add
//   warp 12     one elected thread issues tcgen05.mma (M = 128, N <= 256, K = 8, kind::tf32) and tcgen05.commit
//   warp 13     NT form: one elected thread streams the B operand (the layer's weights, split and swizzled ONCE per call
//               by pack_b_kernel instead of once per 128-row tile) from L2 with cp.async.bulk onto the stage's mbarrier
// 2-stage operand ring (96 KB per stage) with full / empty mbarriers; NT double-buffers the accumulator (2 x 256 TMEM
// columns) so the drain of tile i overlaps the MMAs of tile i + 1.  Every mbarrier wait is bounded (watchdog ->
// error code), as in the other tensor-core kernels.
#include "../../include/ucnerf_b200.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "tc_common.cuh"

namespace ucnerf {
namespace g3 {

using namespace tc;

constexpr int kKC32 = 32;                                  // K elements per chunk = one 128-byte swizzle row of TF32
constexpr uint32_t kATile = 128 * 128;                     // 16 KB (one of hi / lo): 128 rows x 128 B
constexpr uint32_t kBTile = 256 * 128;                     // 32 KB
constexpr uint32_t kStageBytes = 2 * kATile + 2 * kBTile;  // 96 KB: A hi | A lo | B hi | B lo
constexpr int kStages = 2;
constexpr uint32_t kSmemMisc = kStages * kStageBytes;      // 196608
constexpr uint32_t kOffBar = 0, kOffTmem = 128, kOffBias = 256;
constexpr uint32_t kOffStage = kOffBias + 1024;                       // epilogue staging: 4 warps x 32 rows x 144 B
constexpr uint32_t kEpiRowBytes = 144, kEpiWarpBytes = 32 * kEpiRowBytes;
constexpr uint32_t kSmemTotal = kSmemMisc + kOffStage + 4 * kEpiWarpBytes + 1024;   // + manual 1 KB alignment slack
static_assert(kSmemTotal <= 232448, "shared memory budget");
constexpr int kLoaderThreads = 256, kEpiWarp0 = 8, kMmaWarp = 12, kCopyWarp = 13, kThreads = 14 * 32;
constexpr uint32_t kBlobChunk = 2 * kBTile;                // packed weights: hi tile | lo tile per K chunk (64 KB)
constexpr int kMaxChunks = 24;

enum Bar { FULL0 = 0, FULL1, EMPTY0, EMPTY1, ACC_FULL0, ACC_FULL1, ACC_EMPTY0, ACC_EMPTY1, NUM_BARS };

// instruction descriptor (cute::UMMA::InstrDescriptor): c = F32 [4,6), a = b = TF32 (2) at [7,10) / [10,13), K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// round to TF32 (nearest, ties away) with two integer ops; lo = x - hi is exact in fp32
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}

struct Chunk {             // one K chunk (32 columns) of the NT form: operand pointers at the chunk's first column
    const float* a;        // A_s + k0, rows lda apart
    const float* b;        // B_s + k0, rows ldb apart
    uint32_t lda, ldb;
    uint32_t kvalid;       // columns of this chunk that exist (1..32); the rest reads as zero
    uint32_t veca, vecb;   // 1: every row start of a / b is 16-byte aligned (float4 loads allowed)
};

struct NtParams {
    uint32_t M, N, npad;   // npad = N rounded up to a multiple of 16 (MMA N)
    uint32_t nchunks;
    Chunk ch[kMaxChunks];
    const float* bias;     // [N] or NULL
    int relu;
    float* C;
    uint32_t ldc;
    uint32_t cvec;         // 1: rows of C are 16-byte aligned (float4 stores)
    const uint8_t* bblob;  // packed B operand (pack_b_kernel): chunk c at c * kBlobChunk; NULL: the loaders split B per tile
    uint32_t* dbg;
};

struct TnParams {
    uint32_t M, N1, N2;    // C[N1,N2] += A[M,N1]^T B[M,N2]
    const float *A, *B;
    uint32_t lda, ldb;
    float* C;
    uint32_t ldc;
    uint32_t chunks_per_cta;   // row chunks (32 rows) per CTA along grid.x
    uint32_t mn_major;         // 1: MN-major operand tiles (no transposition), 0: transposing loader + K-major tiles
    uint32_t veca, vecb;       // float4 loads allowed along the rows of A / B
    uint32_t* dbg;
};

// 12 MMAs of one operand stage: 4 k-steps of 8 x (hi*hi + lo*hi + hi*lo)
__device__ __forceinline__ void issue_stage(uint32_t stage_addr, uint32_t acc_taddr, uint32_t idesc, bool first) {
    const uint32_t a_hi = stage_addr, a_lo = a_hi + kATile, b_hi = a_lo + kATile, b_lo = b_hi + kBTile;
#pragma unroll
    for (int ks = 0; ks < kKC32 / 8; ++ks) {
        const uint64_t dah = make_desc(a_hi + ks * 32), dal = make_desc(a_lo + ks * 32);
        const uint64_t dbh = make_desc(b_hi + ks * 32), dbl = make_desc(b_lo + ks * 32);
        umma_tf32(acc_taddr, dah, dbh, idesc, (first && ks == 0) ? 0u : 1u);
        umma_tf32(acc_taddr, dal, dbh, idesc, 1u);
        umma_tf32(acc_taddr, dah, dbl, idesc, 1u);
    }
}

// rows x 32 fp32 tile with K contiguous in global memory -> hi / lo swizzled tiles.  256 threads; 16-byte piece p of the
// tile: row p / 8, piece p % 8; 8 consecutive threads cover one 128-byte row (coalesced, conflict-free stores).  All loads
// of a thread are issued before the first split / store (PIECES = ceil(rows_total * 8 / 256) register-resident pieces).
template <int PIECES>
__device__ __forceinline__ void fetch_rowmajor(float4 (&x)[PIECES], const float* src, uint32_t ld, uint32_t rows_total,
                                               uint32_t rows_valid, uint32_t kvalid, bool vec, int t) {
#pragma unroll
    for (int i = 0; i < PIECES; ++i) {
        const uint32_t p = t + kLoaderThreads * i, row = p >> 3, pc = p & 7;
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows_valid && row < rows_total) {
            const float* g = src + (size_t)row * ld + 4 * pc;
            if (vec && 4 * pc + 3 < kvalid) {
                x[i] = __ldg(reinterpret_cast<const float4*>(g));
            } else {
                if (4 * pc + 0 < kvalid) x[i].x = __ldg(g + 0);
                if (4 * pc + 1 < kvalid) x[i].y = __ldg(g + 1);
                if (4 * pc + 2 < kvalid) x[i].z = __ldg(g + 2);
                if (4 * pc + 3 < kvalid) x[i].w = __ldg(g + 3);
            }
        }
    }
}
template <int PIECES>
__device__ __forceinline__ void store_kmajor(const float4 (&x)[PIECES], uint8_t* hi_tile, uint32_t tile_bytes, uint32_t rows_total,
                                             int t) {
#pragma unroll
    for (int i = 0; i < PIECES; ++i) {
        const uint32_t p = t + kLoaderThreads * i, row = p >> 3, pc = p & 7;
        if (row < rows_total) {
            uint4 hi, lo;
            split_tf32(x[i].x, hi.x, lo.x); split_tf32(x[i].y, hi.y, lo.y); split_tf32(x[i].z, hi.z, lo.z); split_tf32(x[i].w, hi.w, lo.w);
            uint8_t* dst = hi_tile + (row >> 3) * 1024 + (row & 7) * 128 + ((pc ^ (row & 7)) * 16);
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + tile_bytes) = lo;
        }
    }
}
template <int PIECES>
__device__ __forceinline__ void load_rowmajor_tile(uint8_t* hi_tile, uint32_t tile_bytes, const float* src, uint32_t ld,
                                                   uint32_t rows_total, uint32_t rows_valid, uint32_t kvalid, bool vec, int t) {
    float4 x[PIECES];
    fetch_rowmajor<PIECES>(x, src, ld, rows_total, rows_valid, kvalid, vec, t);
    store_kmajor<PIECES>(x, hi_tile, tile_bytes, rows_total, t);
}

// TRANSPOSED load for the TN form: the tile's row index is a COLUMN n of the source, its K index a source ROW m.
// Tile rows [0, rows_total) <-> source columns col0 + row (valid below cols_valid), K = 32 source rows from m0 (valid below
// m_valid).  One warp instruction covers 8 columns x 4 rows: lane = a + 8 b, a = column % 8, b = row % 4 - four 32-byte
// sectors from global memory, and 32 distinct shared-memory banks on the way out (bank = ((q ^ a) * 4 + b)).  A warp owns
// the (j = tile row / 8, q = k / 4) pairs pi = q + 8 j with pi % 8 == warp, U of them in flight at a time.
template <int U>
__device__ __forceinline__ void load_transposed_tile(uint8_t* hi_tile, uint32_t tile_bytes, const float* src, uint32_t ld,
                                                     uint32_t rows_total, uint32_t col0, uint32_t cols_valid, uint32_t m0,
                                                     uint32_t m_valid, int warp, int lane) {
    const uint32_t a = lane & 7, b = lane >> 3;
    const uint32_t npairs = rows_total;   // (rows_total / 8) values of j x 8 values of q
    for (uint32_t base = warp; base < npairs; base += 8 * U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t pi = base + 8 * u;
            const uint32_t q = pi & 7, j = pi >> 3;
            const uint32_t col = col0 + 8 * j + a, m = m0 + 4 * q + b;
            v[u] = (pi < npairs && col < cols_valid && m < m_valid) ? __ldg(src + (size_t)m * ld + col) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t pi = base + 8 * u;
            if (pi < npairs) {
                const uint32_t q = pi & 7, j = pi >> 3;
                uint32_t hi, lo;
                split_tf32(v[u], hi, lo);
                uint8_t* dst = hi_tile + j * 1024 + a * 128 + ((q ^ a) * 16) + b * 4;
                *reinterpret_cast<uint32_t*>(dst) = hi;
                *reinterpret_cast<uint32_t*>(dst + tile_bytes) = lo;
            }
        }
    }
}

// MN-MAJOR tile for the TN form without any transposition: the reduction index K runs over source ROWS and the tile's
// M / N index over source COLUMNS, which are contiguous in memory - the canonical MN-major layout of UMMA (instruction-
// descriptor bits 15 / 16).  For 32-bit operands the only MN-major shared-memory layout is SWIZZLE_128B_BASE32B (layout
// type 1; CUTLASS: Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> o (32 MN x 4 K):(1, 32)): atoms of 4 K-rows x 128 bytes
// (32 MN elements), the 32-byte chunk index of a row XOR-ed with k % 4.  Atom (kb = k / 4, mb = mn / 32) sits at
// (kb * (mn_total / 32) + mb) * 512: leading byte offset (next MN atom) 512, stride byte offset (next K atom)
// mn_total / 32 * 512; a K = 8 MMA reads two K atoms.  Loads are float4 along the source row: as cheap as the K-major
// loader of the NT form.
template <int PIECES>
__device__ __forceinline__ void fetch_mnmajor(float4 (&x)[PIECES], const float* src, uint32_t ld, uint32_t mn_total, uint32_t col0,
                                              uint32_t cols_valid, uint32_t m0, uint32_t m_valid, bool vec, int t) {
    const uint32_t ppr = mn_total >> 2;            // 16-byte pieces per K row
#pragma unroll
    for (int i = 0; i < PIECES; ++i) {
        const uint32_t p = t + kLoaderThreads * i, k = p / ppr, mp = p - k * ppr;
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t col = col0 + 4 * mp, m = m0 + k;
        if (k < 32 && m < m_valid && col < cols_valid) {
            const float* g = src + (size_t)m * ld + col;
            if (vec && col + 3 < cols_valid) {
                x[i] = __ldg(reinterpret_cast<const float4*>(g));
            } else {
                x[i].x = __ldg(g);
                if (col + 1 < cols_valid) x[i].y = __ldg(g + 1);
                if (col + 2 < cols_valid) x[i].z = __ldg(g + 2);
                if (col + 3 < cols_valid) x[i].w = __ldg(g + 3);
            }
        }
    }
}
template <int PIECES>
__device__ __forceinline__ void store_mnmajor(const float4 (&x)[PIECES], uint8_t* hi_tile, uint32_t tile_bytes, uint32_t mn_total,
                                              int t) {
    const uint32_t ppr = mn_total >> 2, atoms = mn_total >> 5;
#pragma unroll
    for (int i = 0; i < PIECES; ++i) {
        const uint32_t p = t + kLoaderThreads * i, k = p / ppr, mp = p - k * ppr;
        if (k < 32) {
            uint4 hi, lo;
            split_tf32(x[i].x, hi.x, lo.x); split_tf32(x[i].y, hi.y, lo.y); split_tf32(x[i].z, hi.z, lo.z); split_tf32(x[i].w, hi.w, lo.w);
            uint8_t* dst = hi_tile + ((k >> 2) * atoms + (mp >> 3)) * 512 + (k & 3) * 128 + (((((mp & 7) >> 1) ^ (k & 3))) * 32) +
                           (mp & 1) * 16;
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + tile_bytes) = lo;
        }
    }
}
// shared-memory matrix descriptor, MN-major SWIZZLE_128B_BASE32B (layout type 1): LBO / SBO in 16-byte units
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)(sbo16 & 0x3FFFu) << 32) |
           (1ull << 46) | (1ull << 61);
}
// 12 MMAs of one MN-major operand stage (A: 128 wide = 4 atoms, B: nb atoms)
__device__ __forceinline__ void issue_stage_mn(uint32_t stage_addr, uint32_t acc_taddr, uint32_t idesc, bool first, uint32_t nb) {
    const uint32_t a_hi = stage_addr, a_lo = a_hi + kATile, b_hi = a_lo + kATile, b_lo = b_hi + kBTile;
#pragma unroll
    for (int ks = 0; ks < kKC32 / 8; ++ks) {
        const uint64_t dah = make_desc_mn(a_hi + ks * 4096, 32, 128), dal = make_desc_mn(a_lo + ks * 4096, 32, 128);
        const uint64_t dbh = make_desc_mn(b_hi + ks * nb * 1024, 32, nb * 32), dbl = make_desc_mn(b_lo + ks * nb * 1024, 32, nb * 32);
        umma_tf32(acc_taddr, dah, dbh, idesc, (first && ks == 0) ? 0u : 1u);
        umma_tf32(acc_taddr, dal, dbh, idesc, 1u);
        umma_tf32(acc_taddr, dah, dbl, idesc, 1u);
    }
}

struct Shared {
    uint8_t* smem;
    uint32_t bar0;
    __device__ uint32_t bar(int i) const { return bar0 + 8u * (uint32_t)i; }
    __device__ uint8_t* stage(int s) const { return smem + (size_t)s * kStageBytes; }
};

__device__ __forceinline__ Shared setup(uint8_t* smem_raw, int warp, uint32_t& tmem_base, uint32_t full_count = kLoaderThreads,
                                        int mma_warp = kMmaWarp) {
    Shared sh;
    sh.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* misc = sh.smem + kSmemMisc;
    sh.bar0 = smem_u32(misc + kOffBar);
    if (threadIdx.x == 0) {
        mbar_init(sh.bar(FULL0), full_count); mbar_init(sh.bar(FULL1), full_count);
        mbar_init(sh.bar(EMPTY0), 1); mbar_init(sh.bar(EMPTY1), 1);
        mbar_init(sh.bar(ACC_FULL0), 1); mbar_init(sh.bar(ACC_FULL1), 1);
        mbar_init(sh.bar(ACC_EMPTY0), 128); mbar_init(sh.bar(ACC_EMPTY1), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == mma_warp) {   // whole warp: allocate all 512 TMEM columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(misc + kOffTmem)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem_base = *reinterpret_cast<volatile uint32_t*>(misc + kOffTmem);
    return sh;
}

__device__ __forceinline__ void teardown(int warp, uint32_t tmem_base, int mma_warp = kMmaWarp) {
    tc_fence_before();
    __syncthreads();
    if (warp == mma_warp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) gemm3_nt_kernel(const __grid_constant__ NtParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t tmem_base;
    const Shared sh = setup(smem_raw, warp, tmem_base, kLoaderThreads + (p.bblob ? 1u : 0u));
    float* sBias = reinterpret_cast<float*>(sh.smem + kSmemMisc + kOffBias);
    for (int i = threadIdx.x; i < 256; i += kThreads) sBias[i] = (p.bias && i < (int)p.N) ? p.bias[i] : 0.f;
    __syncthreads();
    const uint32_t ntiles = (p.M + 127) / 128;
    const uint32_t idesc = idesc_tf32(p.npad);

    if (warp < 8) {
        // ================= loaders =================
        // Software pipeline over the flattened (tile, chunk) sequence: the global loads of step i + 1 are issued before the
        // split / store of step i, and neither needs the stage to be free - the loads are in flight while the MMAs of
        // earlier chunks still read it.
        const int t = threadIdx.x;
        const uint32_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
        const uint32_t nsteps = my_tiles * p.nchunks;
        auto fetch = [&](float4 (&x)[4], uint32_t i) {
            const uint32_t tile = blockIdx.x + (i / p.nchunks) * gridDim.x, c = i % p.nchunks;
            const Chunk& ch = p.ch[c];
            const uint32_t row0 = tile * 128, rows_valid = min(128u, p.M - row0);
            fetch_rowmajor<4>(x, ch.a + (size_t)row0 * ch.lda, ch.lda, 128, rows_valid, ch.kvalid, ch.veca != 0, t);
        };
        auto commit = [&](const float4 (&x)[4], uint32_t i) -> bool {
            const int s = i & 1;
            if (!mbar_wait(sh.bar(EMPTY0 + s), ((i >> 1) & 1) ^ 1, p.dbg, 1, EMPTY0 + s, i, 0)) return false;
            uint8_t* st = sh.stage(s);
            store_kmajor<4>(x, st, kATile, 128, t);
            if (!p.bblob) {
                const Chunk& ch = p.ch[i % p.nchunks];
                load_rowmajor_tile<8>(st + 2 * kATile, kBTile, ch.b, ch.ldb, p.npad, p.N, ch.kvalid, ch.vecb != 0, t);
            }
            fence_proxy_async();
            mbar_arrive(sh.bar(FULL0 + s));
            return true;
        };
        float4 x0[4], x1[4];
        if (nsteps > 0) fetch(x0, 0);
        for (uint32_t i = 0; i < nsteps; i += 2) {
            if (i + 1 < nsteps) fetch(x1, i + 1);
            if (!commit(x0, i)) goto done;
            if (i + 1 < nsteps) {
                if (i + 2 < nsteps) fetch(x0, i + 2);
                if (!commit(x1, i + 1)) goto done;
            }
        }
    } else if (warp < kMmaWarp) {
        // ================= epilogue: thread e <-> row e of the tile =================
        const int e = threadIdx.x - 32 * kEpiWarp0;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint8_t* stage_w = sh.smem + kSmemMisc + kOffStage + (warp & 3) * kEpiWarpBytes;
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            if (!mbar_wait(sh.bar(ACC_FULL0 + acc), (it >> 1) & 1, p.dbg, 2, ACC_FULL0 + acc, tile, 0)) goto done;
            tc_fence_after();
            const uint32_t row = tile * 128 + e;
            float* crow = p.C + (size_t)row * p.ldc;
            for (uint32_t j = 0; j * 32 < p.npad; ++j) {
                uint32_t r[32];
                tmem_ld32(lane_taddr + (uint32_t)(acc * 256) + 32 * j, r);
                if (row < p.M) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        v[i] = __uint_as_float(r[i]) + sBias[(32 * j + i) & 255];
                        if (p.relu) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (!(p.cvec && 32 * j + 32 <= p.N)) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (32 * j + i < p.N) crow[32 * j + i] = v[i];
                    } else {
                        // thread = row would store 32 different 128-byte lines per instruction: stage the warp's 32 x 32 block
                        // in shared memory and write it out 4 full rows (4 x 128 B) per instruction instead
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4*>(stage_w + lane * kEpiRowBytes + 16 * q) =
                                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                }
                if (p.cvec && 32 * j + 32 <= p.N) {      // (warp-uniform)
                    __syncwarp();
                    const uint32_t wrow0 = tile * 128 + (uint32_t)(warp & 3) * 32;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t rr = 4 * i + (lane >> 3), pc = lane & 7;
                        const float4 x = *reinterpret_cast<const float4*>(stage_w + rr * kEpiRowBytes + 16 * pc);
                        if (wrow0 + rr < p.M) *reinterpret_cast<float4*>(p.C + (size_t)(wrow0 + rr) * p.ldc + 32 * j + 4 * pc) = x;
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive(sh.bar(ACC_EMPTY0 + acc));
        }
    } else if (warp == kCopyWarp) {
        // ================= B operand: bulk copies of the pre-packed weight chunks =================
        if (lane == 0 && p.bblob) {
            uint32_t k = 0;
            const uint32_t bytes = p.npad * 128;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (uint32_t c = 0; c < p.nchunks; ++c, ++k) {
                    const int s = k & 1;
                    if (!mbar_wait(sh.bar(EMPTY0 + s), ((k >> 1) & 1) ^ 1, p.dbg, 5, EMPTY0 + s, tile, c)) goto done;
                    const uint32_t b_hi = smem_u32(sh.stage(s) + 2 * kATile);
                    mbar_expect_tx(sh.bar(FULL0 + s), 2 * bytes);
                    bulk_g2s(b_hi, p.bblob + (size_t)c * kBlobChunk, bytes, sh.bar(FULL0 + s));
                    bulk_g2s(b_hi + kBTile, p.bblob + (size_t)c * kBlobChunk + kBTile, bytes, sh.bar(FULL0 + s));
                }
            }
        }
    } else if (lane == 0) {
        // ================= MMA issuer =================
        uint32_t k = 0, it = 0;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            if (!mbar_wait(sh.bar(ACC_EMPTY0 + acc), ((it >> 1) & 1) ^ 1, p.dbg, 3, ACC_EMPTY0 + acc, tile, 0)) goto done;
            tc_fence_after();
            for (uint32_t c = 0; c < p.nchunks; ++c, ++k) {
                const int s = k & 1;
                if (!mbar_wait(sh.bar(FULL0 + s), (k >> 1) & 1, p.dbg, 4, FULL0 + s, tile, c)) goto done;
                tc_fence_after();
                issue_stage(smem_u32(sh.stage(s)), tmem_base + (uint32_t)(acc * 256), idesc, c == 0);
                umma_commit(sh.bar(EMPTY0 + s));
            }
            umma_commit(sh.bar(ACC_FULL0 + acc));
        }
    }
done:
    teardown(warp, tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------------
// grid = (row splits, N1 tiles of 128, N2 blocks of 256)
__global__ void __launch_bounds__(kThreads, 1) gemm3_tn_kernel(const __grid_constant__ TnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t tmem_base;
    const Shared sh = setup(smem_raw, warp, tmem_base);
    const uint32_t n1_0 = blockIdx.y * 128, n2_0 = blockIdx.z * 256;
    const uint32_t n2 = min(256u, p.N2 - n2_0), npad2 = p.mn_major ? ((n2 + 31) & ~31u) : ((n2 + 15) & ~15u);
    const uint32_t total_chunks = (p.M + 31) / 32;
    const uint32_t c_begin = blockIdx.x * p.chunks_per_cta, c_end = min(c_begin + p.chunks_per_cta, total_chunks);
    const uint32_t idesc = idesc_tf32(npad2) | (p.mn_major ? ((1u << 15) | (1u << 16)) : 0u);
    if (c_begin >= c_end) goto done;   // (uniform per CTA)

    if (warp < 8) {
        uint32_t k = 0;
        if (p.mn_major) {
            // the global loads do not need the stage: they are in flight while the MMAs of chunk k - 2 still read it.
            // (Issuing the loads of chunk c + 1 before the stores of chunk c as well - 24 float4 in flight per thread - was
            // measured 2.8x SLOWER: 128 registers + 184 bytes of spills in the loader warps.)
            const int t = threadIdx.x;
            for (uint32_t c = c_begin; c < c_end; ++c, ++k) {
                const int s = k & 1;
                float4 xa[4], xb[8];
                fetch_mnmajor<4>(xa, p.A, p.lda, 128, n1_0, p.N1, c * 32, p.M, p.veca != 0, t);
                fetch_mnmajor<8>(xb, p.B, p.ldb, npad2, n2_0, p.N2, c * 32, p.M, p.vecb != 0, t);
                if (!mbar_wait(sh.bar(EMPTY0 + s), ((k >> 1) & 1) ^ 1, p.dbg, 11, EMPTY0 + s, c, 0)) goto done;
                uint8_t* st = sh.stage(s);
                store_mnmajor<4>(xa, st, kATile, 128, t);
                store_mnmajor<8>(xb, st + 2 * kATile, kBTile, npad2, t);
                fence_proxy_async();
                mbar_arrive(sh.bar(FULL0 + s));
            }
        } else {   // transposing loader + K-major tiles (the first version; env UCNERF_GEMM_TN_TRANSPOSE=1)
            for (uint32_t c = c_begin; c < c_end; ++c, ++k) {
                const int s = k & 1;
                uint8_t* st = sh.stage(s);
                if (!mbar_wait(sh.bar(EMPTY0 + s), ((k >> 1) & 1) ^ 1, p.dbg, 11, EMPTY0 + s, c, 0)) goto done;
                load_transposed_tile<16>(st, kATile, p.A, p.lda, 128, n1_0, p.N1, c * 32, p.M, warp, lane);
                load_transposed_tile<16>(st + 2 * kATile, kBTile, p.B, p.ldb, npad2, n2_0, p.N2, c * 32, p.M, warp, lane);
                fence_proxy_async();
                mbar_arrive(sh.bar(FULL0 + s));
            }
        }
    } else if (warp < kMmaWarp) {
        const int e = threadIdx.x - 32 * kEpiWarp0;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        if (!mbar_wait(sh.bar(ACC_FULL0), 0, p.dbg, 12, ACC_FULL0, 0, 0)) goto done;
        tc_fence_after();
        const uint32_t row = n1_0 + e;
        for (uint32_t j = 0; j * 32 < npad2; ++j) {
            uint32_t r[32];
            tmem_ld32(lane_taddr + 32 * j, r);
            if (row < p.N1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const uint32_t col = 32 * j + i;
                    if (col < n2) atomicAdd(p.C + (size_t)row * p.ldc + n2_0 + col, __uint_as_float(r[i]));
                }
            }
        }
        tc_fence_before();
    } else if (warp == kMmaWarp && lane == 0) {
        uint32_t k = 0;
        for (uint32_t c = c_begin; c < c_end; ++c, ++k) {
            const int s = k & 1;
            if (!mbar_wait(sh.bar(FULL0 + s), (k >> 1) & 1, p.dbg, 13, FULL0 + s, c, 0)) goto done;
            tc_fence_after();
            if (p.mn_major) issue_stage_mn(smem_u32(sh.stage(s)), tmem_base, idesc, c == c_begin, npad2 >> 5);
            else issue_stage(smem_u32(sh.stage(s)), tmem_base, idesc, c == c_begin);
            umma_commit(sh.bar(EMPTY0 + s));
        }
        umma_commit(sh.bar(ACC_FULL0));
    }
done:
    teardown(warp, tmem_base);
}

// B operand of the NT form (a layer's weights or their transpose), split into hi / lo TF32 and swizzled once per call:
// block c writes K chunk c as [hi tile | lo tile] in the layout the MMA reads from shared memory.
__global__ void __launch_bounds__(kLoaderThreads) pack_b_kernel(const __grid_constant__ NtParams p, uint8_t* __restrict__ blob) {
    const Chunk& ch = p.ch[blockIdx.x];
    load_rowmajor_tile<8>(blob + (size_t)blockIdx.x * kBlobChunk, kBTile, ch.b, ch.ldb, p.npad, p.N, ch.kvalid, ch.vecb != 0,
                          threadIdx.x);
}

// Backward glue of a dense layer in ONE pass over the rows: g = gy * (y > 0) (the ReLU mask, relu'(pre) == (out > 0)) and
// colsum[n] = sum_m g[m, n] (the bias gradient) - what autograd otherwise spreads over a compare, a multiply and a column
// reduction (three reads + one write of the [M, N] tensor instead of two reads + one write).  y == NULL: no mask;
// g == NULL: column sums only.  Thread = 4 consecutive columns (float4), rows strided over the grid.
__global__ void __launch_bounds__(256) relu_mask_colsum_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                                               float* __restrict__ g, float* __restrict__ colsum, uint32_t M,
                                                               uint32_t N, uint32_t rows_per_cta) {
    const uint32_t n4 = (N + 3) >> 2;                      // float4 groups per row (N % 4 == 0 on this path)
    const uint32_t tpr = n4;                               // threads per row
    const uint32_t rpi = 256 / tpr;                        // rows per iteration
    const uint32_t c4 = threadIdx.x % tpr, rsub = threadIdx.x / tpr;
    const uint32_t r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, M);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rsub < rpi) {
        for (uint32_t r = r0 + rsub; r < r1; r += rpi) {
            const size_t o = (size_t)r * N + 4 * c4;
            float4 v = *reinterpret_cast<const float4*>(gy + o);
            if (y) {
                const float4 yy = *reinterpret_cast<const float4*>(y + o);
                v.x = yy.x > 0.f ? v.x : 0.f; v.y = yy.y > 0.f ? v.y : 0.f; v.z = yy.z > 0.f ? v.z : 0.f; v.w = yy.w > 0.f ? v.w : 0.f;
            }
            if (g) *reinterpret_cast<float4*>(g + o) = v;
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (colsum) {
        __shared__ float4 sh[256];
        sh[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x < tpr) {
            float4 t = sh[threadIdx.x];
            for (uint32_t k = 1; k < rpi; ++k) {
                const float4 u = sh[threadIdx.x + k * tpr];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            atomicAdd(colsum + 4 * threadIdx.x + 0, t.x); atomicAdd(colsum + 4 * threadIdx.x + 1, t.y);
            atomicAdd(colsum + 4 * threadIdx.x + 2, t.z); atomicAdd(colsum + 4 * threadIdx.x + 3, t.w);
        }
    }
}

static uint32_t* g_dbg = nullptr;   // [32] words: watchdog record (0 = healthy)

// packed-weight workspace per (device, stream): launches on one stream are ordered, different streams get their own
static std::mutex g_blob_mu;
static std::map<std::pair<int, void*>, uint8_t*> g_blobs;
static int blob_for(void* stream, uint8_t** out) {
    int dev = 0;
    UC_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_blob_mu);
    auto key = std::make_pair(dev, stream);
    auto it = g_blobs.find(key);
    if (it == g_blobs.end()) {
        uint8_t* b = nullptr;
        UC_CUDA_OK(cudaMalloc(&b, (size_t)kMaxChunks * kBlobChunk));
        it = g_blobs.emplace(key, b).first;
    }
    *out = it->second;
    return 0;
}

static int ensure_dbg() {
    if (!g_dbg) {
        UC_CUDA_OK(cudaMalloc(&g_dbg, 32 * sizeof(uint32_t)));
        UC_CUDA_OK(cudaMemset(g_dbg, 0, 32 * sizeof(uint32_t)));
    }
    return 0;
}

}  // namespace g3
}  // namespace ucnerf

using namespace ucnerf;
using namespace ucnerf::g3;

extern "C" int ucnerf_gemm_status(uint32_t* out32) {
    UC_REQUIRE(out32, "gemm_status: null argument");
    std::memset(out32, 0, 32 * sizeof(uint32_t));
    if (!g_dbg) return 0;
    UC_CUDA_OK(cudaMemcpy(out32, g_dbg, 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (out32[0] != 0) {
        set_error("gemm3: pipeline watchdog fired (tag " + std::to_string(out32[0]) + ", block " + std::to_string(out32[1]) +
                  ", barrier " + std::to_string(out32[3]) + ")");
        UC_CUDA_OK(cudaMemset(g_dbg, 0, 32 * sizeof(uint32_t)));
        return 5;
    }
    return 0;
}

extern "C" int ucnerf_gemm_nt(uint32_t M, uint32_t N, uint32_t nseg, const ucnerf_gemm_seg* segs, const float* bias, int relu,
                              float* C, uint32_t ldc, void* stream) {
    UC_REQUIRE(segs && C && nseg >= 1, "gemm_nt: null argument");
    UC_REQUIRE(N >= 1 && N <= 256, "gemm_nt: N must be in [1, 256]");
    UC_REQUIRE(ldc >= N, "gemm_nt: ldc < N");
    if (M == 0) return 0;
    if (int e = ensure_dbg()) return e;
    NtParams p{};
    p.M = M; p.N = N; p.npad = (N + 15) & ~15u; p.bias = bias; p.relu = relu; p.C = C; p.ldc = ldc; p.dbg = g_dbg;
    p.cvec = (reinterpret_cast<uintptr_t>(C) % 16 == 0 && ldc % 4 == 0) ? 1u : 0u;
    uint32_t nc = 0;
    for (uint32_t s = 0; s < nseg; ++s) {
        const ucnerf_gemm_seg& g = segs[s];
        UC_REQUIRE(g.a && g.b && g.k >= 1 && g.lda >= g.k && g.ldb >= g.k, "gemm_nt: bad segment");
        const bool veca = reinterpret_cast<uintptr_t>(g.a) % 16 == 0 && g.lda % 4 == 0;
        const bool vecb = reinterpret_cast<uintptr_t>(g.b) % 16 == 0 && g.ldb % 4 == 0;
        for (uint32_t k0 = 0; k0 < g.k; k0 += kKC32) {
            UC_REQUIRE(nc < (uint32_t)kMaxChunks, "gemm_nt: sum of K over the segments exceeds 768");
            Chunk& c = p.ch[nc++];
            c.a = g.a + k0; c.b = g.b + k0; c.lda = g.lda; c.ldb = g.ldb; c.kvalid = std::min<uint32_t>(kKC32, g.k - k0);
            c.veca = veca ? 1u : 0u; c.vecb = vecb ? 1u : 0u;
        }
    }
    p.nchunks = nc;
    const uint32_t ntiles = (M + 127) / 128;
    static const int no_pack = getenv("UCNERF_GEMM_NO_PACK") ? atoi(getenv("UCNERF_GEMM_NO_PACK")) : 0;
    if (!no_pack && ntiles > 1) {   // weights split + swizzled once per call, streamed by bulk copies
        uint8_t* blob = nullptr;
        if (int e = blob_for(stream, &blob)) return e;
        pack_b_kernel<<<nc, kLoaderThreads, 0, (cudaStream_t)stream>>>(p, blob);
        UC_LAUNCH_CHECK();
        p.bblob = blob;
    }
    UC_ENSURE_SMEM(kSmemTotal, gemm3_nt_kernel);
    gemm3_nt_kernel<<<std::min<uint32_t>(ntiles, (uint32_t)kNumSMs), kThreads, kSmemTotal, (cudaStream_t)stream>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ucnerf_gemm_tn(uint32_t M, uint32_t N1, uint32_t N2, const float* A, uint32_t lda, const float* B, uint32_t ldb,
                              float* C, uint32_t ldc, void* stream) {
    UC_REQUIRE(A && B && C, "gemm_tn: null argument");
    UC_REQUIRE(N1 >= 1 && N2 >= 1 && lda >= N1 && ldb >= N2 && ldc >= N2, "gemm_tn: bad shape");
    if (M == 0) return 0;
    if (int e = ensure_dbg()) return e;
    TnParams p{};
    p.M = M; p.N1 = N1; p.N2 = N2; p.A = A; p.B = B; p.lda = lda; p.ldb = ldb; p.C = C; p.ldc = ldc; p.dbg = g_dbg;
    {
        static const int force_transpose = getenv("UCNERF_GEMM_TN_TRANSPOSE") ? atoi(getenv("UCNERF_GEMM_TN_TRANSPOSE")) : 0;
        p.mn_major = force_transpose ? 0u : 1u;
    }
    p.veca = (reinterpret_cast<uintptr_t>(A) % 16 == 0 && lda % 4 == 0) ? 1u : 0u;
    p.vecb = (reinterpret_cast<uintptr_t>(B) % 16 == 0 && ldb % 4 == 0) ? 1u : 0u;
    const uint32_t t1 = (N1 + 127) / 128, t2 = (N2 + 255) / 256;
    const uint32_t total_chunks = (M + 31) / 32;
    uint32_t gx = std::max<uint32_t>(1u, (uint32_t)kNumSMs / (t1 * t2));
    gx = std::min(gx, total_chunks);
    p.chunks_per_cta = (total_chunks + gx - 1) / gx;
    gx = (total_chunks + p.chunks_per_cta - 1) / p.chunks_per_cta;
    UC_ENSURE_SMEM(kSmemTotal, gemm3_tn_kernel);
    gemm3_tn_kernel<<<dim3(gx, t1, t2), kThreads, kSmemTotal, (cudaStream_t)stream>>>(p);
    UC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ucnerf_relu_mask_colsum(const float* gy, const float* y, float* g, float* colsum, uint32_t M, uint32_t N,
                                       void* stream) {
    UC_REQUIRE(gy && (g || colsum), "relu_mask_colsum: null argument");
    UC_REQUIRE(N >= 4 && N <= 1024 && N % 4 == 0, "relu_mask_colsum: N must be a multiple of 4 in [4, 1024]");
    UC_REQUIRE(reinterpret_cast<uintptr_t>(gy) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(g) % 16 == 0, "relu_mask_colsum: tensors must be 16-byte aligned");
    if (colsum) UC_CUDA_OK(cudaMemsetAsync(colsum, 0, sizeof(float) * N, (cudaStream_t)stream));
    if (M == 0) return 0;
    const uint32_t ctas = std::min<uint32_t>((uint32_t)kNumSMs * 8u, (M + 63) / 64);
    const uint32_t rows_per_cta = (M + ctas - 1) / ctas;
    relu_mask_colsum_kernel<<<(M + rows_per_cta - 1) / rows_per_cta, 256, 0, (cudaStream_t)stream>>>(gy, y, g, colsum, M, N,
                                                                                                 rows_per_cta);
    UC_LAUNCH_CHECK();
    return 0;
}

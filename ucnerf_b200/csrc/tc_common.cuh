// tcgen05 / TMEM / mbarrier / bulk-copy helpers shared by the tensor-core MLP kernels (color_mlp_tc.cu, sky_mlp_tc.cu).
// sm_100a only.  Geometry common to both kernels: 128-row tiles, K chunks of 64 FP16 elements (one 128-byte swizzle
// row), A operand as hi tile + lo tile (16 KB each) in the UMMA canonical K-major SWIZZLE_128B layout.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace ucnerf {
namespace tc {

constexpr int kTileM = 128;
constexpr int kKC = 64;                              // K elements per chunk = one 128-byte swizzle row of FP16
constexpr uint32_t kATileBytes = kTileM * kKC * 2;   // 16 KB (one of hi / lo)
constexpr uint32_t kASlotBytes = 2 * kATileBytes;    // 32 KB

// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=F16 (0) [7,10), b=F16 (0) [10,13),
// a/b K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24); }

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
static __device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, uint32_t* dbg, uint32_t tag, uint32_t bar_id,
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
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
// x (already scaled) -> FP16 hi / lo pair for two neighbouring K elements: hi = rn(x), lo = rn(x - hi)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// write half `hf` (32 of the 64 K elements) of one row of an A chunk as hi / lo FP16 tiles in the SWIZZLE_128B
// K-major layout: row r at (r/8)*1024 + (r%8)*128, 16-byte piece pc (8 elements) at ((pc ^ (r%8)) * 16)
__device__ __forceinline__ void store_a_half(uint8_t* slot, int row, int hf, const float (&v)[32]) {
    uint8_t* base = slot + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint4 hi, lo;
        split2(v[8 * q + 0], v[8 * q + 1], hi.x, lo.x);
        split2(v[8 * q + 2], v[8 * q + 3], hi.y, lo.y);
        split2(v[8 * q + 4], v[8 * q + 5], hi.z, lo.z);
        split2(v[8 * q + 6], v[8 * q + 7], hi.w, lo.w);
        const int pc = ((4 * hf + q) ^ (row & 7)) * 16;
        *reinterpret_cast<uint4*>(base + pc) = hi;
        *reinterpret_cast<uint4*>(base + kATileBytes + pc) = lo;
    }
}


}  // namespace tc
}  // namespace ucnerf

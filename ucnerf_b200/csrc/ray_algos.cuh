// Per-ray and per-sample algorithms of the UC-NeRF eval path, written once as host+device
// templates.  The CUDA kernels (ray_march.cu) instantiate them with WarpExec (32 lanes cooperate
// on one ray); tests/cpu_harness.cpp instantiates the very same code with SerialExec so the tie
// semantics of the resampler can be checked against the oracle in a GPU-less container.  The
// harness is test infrastructure: the product library contains only the CUDA instantiation.
//
// Reference (under /root/reference/nerf/internal/): stepfun.py:L63-128,L154-218,L251-294,L329-339;
// math.py:L88-107; render.py:L94-244; coord.py:L60-72; models.py:L143-208,L485-496.
//
// Wherever the reference evaluates a chain of separate torch elementwise ops (each rounded to fp32)
// the code below uses the explicitly-rounded helpers fm/fa/fs/fd so nvcc cannot contract them into
// FMAs: sample positions feed a hash grid whose finest level has 8192 cells per unit, so one ulp of
// position is ~5e-4 of a cell; keeping these chains bit-identical removes that error source.
#pragma once
#include <math.h>
#include <string.h>
#include <stdint.h>
#include "common.cuh"

#if defined(__CUDACC__)
#define UC_HD __host__ __device__ __forceinline__
#else
#define UC_HD inline
#endif

namespace ucnerf {

constexpr float kEps = 1.1920929e-07f;  // torch.finfo(float32).eps
constexpr float kFltMax = 3.402823466e+38f;

UC_HD float fm(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
UC_HD float fa(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
UC_HD float fs(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b;
    return r;
#endif
}
UC_HD float fd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b;
    return r;
#endif
}
UC_HD float fsqrt(float a) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}

// ---- execution policies ---------------------------------------------------------------------
struct SerialExec {
    static constexpr int kStride = 1;
    int lane = 0;
    UC_HD float sum(double v) const { return (float)v; }
    UC_HD float max(float v) const { return v; }
    UC_HD void sync() const {}
};

#if defined(__CUDACC__)
struct WarpExec {
    static constexpr int kStride = 32;
    int lane;
    // partial sums are carried in fp64 and rounded once: torch's fp32 reductions are cascade sums whose
    // error (~1 ulp) is far below a naive sequential fp32 accumulation over several hundred terms
    __device__ __forceinline__ float sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return (float)v;
    }
    __device__ __forceinline__ float max(float v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};
#endif

// cw[0] = 0; cw[k+1] = min(float(sum_{i<=k} pw[i]), 1) for k < nb-1; cw[nb] = 1.
// stepfun.py:L108-128 integrate_weights.  torch's CPU cumsum accumulates fp32 inputs in fp64 and
// rounds every output to fp32 (ATen cumsum_cpu_kernel, acc_type<float,false> = double); the scan
// below is a fp64 warp scan rounded the same way.
template <class X>
UC_HD void cdf_scan(const X& ex, const float* pw, int nb, float* cw) {
#if defined(__CUDA_ARCH__)
    if constexpr (X::kStride == 32) {
        double carry = 0.0;
        if (ex.lane == 0) cw[0] = 0.f;
        for (int base = 0; base < nb - 1; base += 32) {
            const int k = base + ex.lane;
            double v = (k < nb - 1) ? (double)pw[k] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double n = __shfl_up_sync(0xffffffffu, v, o);
                if (ex.lane >= o) v += n;
            }
            v += carry;
            if (k < nb - 1) cw[k + 1] = fminf((float)v, 1.f);
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
        if (ex.lane == 0) cw[nb] = 1.f;
        return;
    }
#endif
    if (ex.lane == 0) {
        double acc = 0.0;
        cw[0] = 0.f;
        for (int k = 0; k < nb - 1; ++k) {
            acc += (double)pw[k];
            cw[k + 1] = fminf((float)acc, 1.f);
        }
        cw[nb] = 1.f;
    }
}

// first index in [0,n) with a[idx] > x (n if none); `a` non-decreasing
UC_HD int upper_bound_f(const float* a, int n, float x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Partition point of a monotone predicate (true ... true, false ... false) over [0,n): the first index where it is
// false (n if none), found by galloping out from `guess` and bisecting the bracket.  Same result as a plain binary
// search; with a guess that is off by e it takes ~2 log2(e) + 2 probes instead of log2(n) + 1 - the resampler's
// merged / dilated fenceposts are shifted copies of one sorted list, so good guesses are free.
// Predicates: a[i] + off < v  /  a[i] + off <= v  (x - d is evaluated as x + (-d): the same IEEE result).
struct ShiftedLess {
    const float* a; float off, v;
    UC_HD bool operator()(int i) const { return fa(a[i], off) < v; }
};
struct ShiftedLeq {
    const float* a; float off, v;
    UC_HD bool operator()(int i) const { return fa(a[i], off) <= v; }
};
// first index >= start where the predicate is false (ties / short runs: the caller knows pred holds before `start`)
template <class Pred>
UC_HD int advance_while(int n, int start, const Pred pred) {
    int i = start;
    while (i < n && pred(i)) ++i;
    return i;
}
UC_HD float bits_to_float(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
UC_HD uint32_t float_to_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

template <class Pred>
UC_HD int partition_from(int n, int guess, const Pred pred) {
    if (n <= 0) return 0;
    const int g = guess < 0 ? 0 : (guess > n - 1 ? n - 1 : guess);
    int lo = 0, hi = n;                    // the answer stays in [lo, hi]
    const bool up = pred(g);
    if (up) lo = g + 1; else hi = g;
    int step = 1;
    while (lo < hi) {                      // gallop away from the guess until the predicate flips
        const int q = up ? lo + step - 1 : hi - step;
        if (q < lo || q >= hi) break;      // ran out of range on that side: [lo, hi] is the bracket
        const bool pq = pred(q);
        if (pq) lo = q + 1; else hi = q;
        if (pq != up) break;
        step <<= 1;
    }
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pred(mid)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// math.py:L88-107 sorted_interp for one query: "last xp <= x" / "first xp > x" semantics, max/min over
// values, nan_to_num(.,0) then clip to [0,1].  xp, fp have n entries, both non-decreasing.
UC_HD float sorted_interp_one(float x, const float* xp, const float* fp, int n) {
    const int ub = upper_bound_f(xp, n, x);
    const int lo = ub > 0 ? ub - 1 : 0;
    const int hi = ub < n ? ub : n - 1;
    const float xp0 = xp[lo], xp1 = xp[hi], fp0 = fp[lo], fp1 = fp[hi];
    float off = fd(fs(x, xp0), fs(xp1, xp0));
    if (off != off) off = 0.f;                // nan -> 0
    else if (off > kFltMax) off = kFltMax;    // +inf -> max
    else if (off < -kFltMax) off = -kFltMax;  // -inf -> lowest
    off = fminf(fmaxf(off, 0.f), 1.f);
    return fa(fp0, fm(off, fs(fp1, fp0)));
}

// scratch of one ray for resample_ray (all float arrays)
struct ResampleScratch {
    float* tp;  // [n+1]
    float* pp;  // [n]
    float* T;   // [3n+1]
    float* W;   // [3n+1]
    float* CW;  // [3n+2] (>= nb+1) ; also [>= 2] for level 0
    float* C;   // [S]
    float* J;   // [3n+1] packed (jhi << 16 | jlo) bin ranges of the dilation, as raw bits; aliases W: entry k is read
                // by the lane that then writes W[k]
    UC_HD static size_t floats(int n, int S) { return (size_t)(n + 1) + n + 2 * (3 * n + 1) + (3 * n + 2) + S; }
    UC_HD void carve(float* base, int n, int S) {
        tp = base; pp = tp + (n + 1); T = pp + n; W = T + (3 * n + 1); CW = W + (3 * n + 1); C = CW + (3 * n + 2);
        J = W;
        (void)S;
    }
};

// One sampling level for one ray: [optional max-dilation of the previous histogram] -> logits ->
// softmax -> CDF -> inverse-CDF at the deterministic u grid -> interval fenceposts.
// models.py:L156-205 + stepfun.py max_dilate_weights / sample_intervals (rand=False).
//   t_prev [n+1], w_prev [n]: previous level's sdist / weights (NULL,NULL,n=1 for the first level:
//   sdist=[0,1], weights=[1], models.py:L143-147).  u [S]: linspace(1/2S, 1-1/2S-eps, S).
//   out_sdist [S+1].
//   jit (optional, training: stepfun.py:L206-212 rand=True): jitter added to the u grid, already scaled by max_jitter;
//   jit_stride 0 = one value for the whole ray (single_jitter), 1 = one per sample.
template <class X>
UC_HD void resample_ray(const X& ex, int n, const float* t_prev, const float* w_prev, bool dilate, float dilation,
                        float anneal, float padding, int S, const float* u, ResampleScratch sc, float* out_sdist,
                        const float* jit = nullptr, int jit_stride = 0) {
    const int lane = ex.lane;
    constexpr int st = X::kStride;
    for (int k = lane; k <= n; k += st) sc.tp[k] = t_prev ? t_prev[k] : (k == 0 ? 0.f : 1.f);
    ex.sync();
    for (int k = lane; k < n; k += st) {
        const float w = w_prev ? w_prev[k] : 1.f;
        // stepfun.py:L63-66 weight_to_pdf
        sc.pp[k] = dilate ? fd(w, fmaxf(fs(sc.tp[k + 1], sc.tp[k]), kEps)) : w;
    }
    ex.sync();
    const float* tq;
    float* wq;
    int nb;
    if (dilate) {
        const int m = 3 * n + 1;
        const float* tp = sc.tp;
        // stepfun.py:L77-80: sort(cat[t, t-d, t+d]) then clip to the domain [0,1]; a 3-way merge of
        // three sorted lists done by ranking (ties ordered A<B<C, values equal so order is moot)
        // The same searches also give, for every merged fencepost T, the contiguous range of dilated bins that cover it
        // (stepfun.py:L81-87: t0_j <= T < t1_j  <=>  j in [jlo, jhi), jhi = #{j : t0_j <= T}, jlo = #{j : t1_j <= T}):
        // the strict counts are extended over ties, fenceposts clipped to the domain ends fall back to a search.
        for (int k = lane; k < m; k += st) {
            float v;
            int rank, jhi, jlo;
            if (k <= n) {  // A: t  (always inside [0,1]: T = v)
                v = tp[k];
                const int nb = partition_from(n, k, ShiftedLess{tp, -dilation, v});            // #B < v
                const int nc = partition_from(n, k - 2, ShiftedLess{tp + 1, dilation, v});     // #C < v
                rank = k + nb + nc;
                jhi = advance_while(n, nb, ShiftedLeq{tp, -dilation, v});
                jlo = advance_while(n, nc, ShiftedLeq{tp + 1, dilation, v});
            } else if (k < 2 * n + 1) {  // B: t[:-1] - d  (may fall below 0)
                const int i0 = k - (n + 1);
                v = fs(tp[i0], dilation);
                const int na = partition_from(n + 1, i0 - 1, ShiftedLeq{tp, 0.f, v});          // #A <= v
                const int nc = partition_from(n, i0 - 3, ShiftedLess{tp + 1, dilation, v});    // #C < v
                rank = i0 + na + nc;
                if (v >= 0.f) {
                    jhi = advance_while(n, i0 + 1, ShiftedLeq{tp, -dilation, v});
                    jlo = advance_while(n, nc, ShiftedLeq{tp + 1, dilation, v});
                } else {
                    jhi = partition_from(n, i0 + 1, ShiftedLeq{tp, -dilation, 0.f});
                    jlo = partition_from(n, 0, ShiftedLeq{tp + 1, dilation, 0.f});
                }
            } else {  // C: t[1:] + d  (may exceed 1)
                const int i0 = k - (2 * n + 1);
                v = fa(tp[i0 + 1], dilation);
                const int na = partition_from(n + 1, i0 + 2, ShiftedLeq{tp, 0.f, v});          // #A <= v
                const int nb = partition_from(n, i0 + 3, ShiftedLeq{tp, -dilation, v});        // #B <= v
                rank = i0 + na + nb;
                if (v <= 1.f) {
                    jhi = nb;
                    jlo = advance_while(n, i0 + 1, ShiftedLeq{tp + 1, dilation, v});
                } else {
                    jhi = partition_from(n, n, ShiftedLeq{tp, -dilation, 1.f});
                    jlo = partition_from(n, i0, ShiftedLeq{tp + 1, dilation, 1.f});
                }
            }
            sc.T[rank] = fminf(fmaxf(v, 0.f), 1.f);
            sc.J[rank] = bits_to_float(((uint32_t)jhi << 16) | (uint32_t)jlo);
        }
        ex.sync();
        // stepfun.py:L81-87: w_dilate[k] = max_j { p_j : t0_j <= T_k < t1_j } (0 if none), k < m-1
        double part = 0.0;
        for (int k = lane; k < m - 1; k += st) {
            const float Tk = sc.T[k];
            const uint32_t jj = float_to_bits(sc.J[k]);
            const int jhi = (int)(jj >> 16), jlo = (int)(jj & 0xffffu);
            float pm = 0.f;
            for (int j = jlo; j < jhi; ++j) pm = fmaxf(pm, sc.pp[j]);
            const float w = fm(pm, fs(sc.T[k + 1], Tk));  // pdf_to_weight, stepfun.py:L69-72
            sc.W[k] = w;
            part += w;
        }
        const float denom = fmaxf(ex.sum(part), kEps);  // stepfun.py:L102-103 renormalize
        ex.sync();
        for (int k = lane; k < m - 1; k += st) sc.W[k] = fd(sc.W[k], denom);
        ex.sync();
        tq = sc.T + 1;  // models.py:L175-176: drop first / last fencepost and weight
        wq = sc.W + 1;
        nb = m - 3;
    } else {
        tq = sc.tp;
        wq = sc.pp;
        nb = n;
    }
    // models.py:L188-191 logits, stepfun.py:L157 softmax
    float mx = -INFINITY;
    for (int k = lane; k < nb; k += st) {
        const float lg = (tq[k + 1] > tq[k]) ? fm(anneal, logf(fa(wq[k], padding))) : -INFINITY;
        sc.CW[k] = lg;
        mx = fmaxf(mx, lg);
    }
    mx = ex.max(mx);
    ex.sync();
    double part = 0.0;
    for (int k = lane; k < nb; k += st) {
        const float e = expf(sc.CW[k] - mx);
        sc.CW[k] = e;
        part += e;
    }
    const float tot = ex.sum(part);
    ex.sync();
    for (int k = lane; k < nb; k += st) wq[k] = fd(sc.CW[k], tot);
    ex.sync();
    cdf_scan(ex, wq, nb, sc.CW);
    ex.sync();
    // stepfun.py:L158-160 + math.py sorted_interp: centers
    for (int i = lane; i < S; i += st)
        sc.C[i] = sorted_interp_one(jit ? fa(u[i], jit[i * jit_stride]) : u[i], sc.CW, tq, nb + 1);
    ex.sync();
    // stepfun.py:L281-293: midpoints, reflected + clamped end fenceposts (domain [0,1])
    for (int k = lane; k <= S; k += st) {
        float v;
        if (k == 0) {
            const float mid0 = fd(fa(sc.C[1], sc.C[0]), 2.f);
            v = fmaxf(fs(fm(2.f, sc.C[0]), mid0), 0.f);
        } else if (k == S) {
            const float midl = fd(fa(sc.C[S - 1], sc.C[S - 2]), 2.f);
            v = fminf(fs(fm(2.f, sc.C[S - 1]), midl), 1.f);
        } else {
            v = fd(fa(sc.C[k], sc.C[k - 1]), 2.f);
        }
        out_sdist[k] = v;
    }
    ex.sync();
}

// ---- ray geometry / cone multisampling --------------------------------------------------------
struct RayGeom {
    float o[3], d[3], e1[3], e2[3];
    float radius, near, far;
};

// torch.linalg.vector_norm (used by torch.norm and F.normalize) accumulates x*x with a fused
// multiply-add chain on CPU (checked bit-exact in tests/test_device_algos_cpu.py)
UC_HD float norm3(float x, float y, float z) { return fsqrt(fmaf(z, z, fmaf(y, y, fm(x, x)))); }

UC_HD void normalize3(float (&v)[3]) {  // F.normalize(dim=-1), eps 1e-12
    const float n = norm3(v[0], v[1], v[2]);
    const float d = fmaxf(n, 1e-12f);
    v[0] = fd(v[0], d); v[1] = fd(v[1], d); v[2] = fd(v[2], d);
}
UC_HD void cross3(const float (&a)[3], const float (&b)[3], float (&c)[3]) {
    c[0] = fs(fm(a[1], b[2]), fm(a[2], b[1]));
    c[1] = fs(fm(a[2], b[0]), fm(a[0], b[2]));
    c[2] = fs(fm(a[0], b[1]), fm(a[1], b[0]));
}

// render.py:L139-146: basis = [normalize(cam x rand), normalize(cam x ortho1), directions]
UC_HD void make_ray_geom(RayGeom& g, const float* o, const float* d, const float* cam, const float* rv,
                         float radius, float near, float far) {
    float c[3] = {cam[0], cam[1], cam[2]}, q[3] = {rv[0], rv[1], rv[2]};
#pragma unroll
    for (int i = 0; i < 3; ++i) { g.o[i] = o[i]; g.d[i] = d[i]; }
    cross3(c, q, g.e1);
    normalize3(g.e1);
    cross3(c, g.e1, g.e2);
    normalize3(g.e2);
    g.radius = radius; g.near = near; g.far = far;
}

// constants of the deterministic hexagonal pattern, computed once on the host (render.py:L119-131)
struct ConeTable {
    float cosv[2][6];  // [odd sample?][j]
    float sinv[2][6];
    float tcoef[6];    // 3/sqrt(7) * (2j/5 - 1)
};

// Constants of the deterministic pattern as torch (CPU, fp32) evaluates them: cos/sin of
// pi/3*[0,2,4,3,5,1] (even samples) and of 5pi/3 - (that + pi/6) (odd samples), render.py:L119-131, and
// 3/sqrt(7)*(2j/5-1), L116.  Embedded as hex-float literals so every host produces identical tables
// (libm and torch's vectorised sin differ by one ulp on one entry); tests/test_host_logic.py re-derives
// them with torch and checks equality.
inline void make_cone_table(ConeTable& ct) {
    static const float kCos[2][6] = {{0x1.0000000000000p+0f, -0x1.0000020000000p-1f, -0x1.fffffa0000000p-2f, -0x1.0000000000000p+0f, 0x1.fffffa0000000p-2f, 0x1.fffffe0000000p-2f}, {0x1.99bc5c0000000p-27f, -0x1.bb67ae0000000p-1f, 0x1.bb67b00000000p-1f, 0x1.5110b40000000p-22f, 0x1.bb67b00000000p-1f, -0x1.bb67b20000000p-1f}};
    static const float kSin[2][6] = {{0x0.0p+0f, 0x1.bb67ae0000000p-1f, -0x1.bb67b00000000p-1f, -0x1.777a5c0000000p-24f, -0x1.bb67b00000000p-1f, 0x1.bb67ae0000000p-1f}, {-0x1.0000000000000p+0f, 0x1.0000020000000p-1f, 0x1.fffffa0000000p-2f, 0x1.0000000000000p+0f, -0x1.fffffa0000000p-2f, -0x1.fffff20000000p-2f}};
    static const float kT[6] = {-0x1.2246d60000000p+0f, -0x1.5c55020000000p-1f, -0x1.d071540000000p-3f, 0x1.d0715e0000000p-3f, 0x1.5c55020000000p-1f, 0x1.2246d60000000p+0f};
    for (int j = 0; j < 6; ++j) {
        ct.cosv[0][j] = kCos[0][j]; ct.cosv[1][j] = kCos[1][j];
        ct.sinv[0][j] = kSin[0][j]; ct.sinv[1][j] = kSin[1][j];
        ct.tcoef[j] = kT[j];
    }
}

// quantities of one interval [t0,t1] shared by its six points (render.py:L108-117)
struct ConeInterval {
    float t0, tdA, B, Cq;  // t = t0 + tdA * (B + tcoef_j * Cq)
};
UC_HD ConeInterval make_cone_interval(float t0, float t1) {
    ConeInterval ci;
    const float tm = fm(fa(t0, t1), 0.5f);  // x/2 == x*0.5 exactly
    const float td = fm(fs(t1, t0), 0.5f);
    const float td2 = fm(td, td), tm2 = fm(tm, tm);
    const float A = fa(td2, fm(3.f, tm2));
    // torch evaluates t_m ** 4 with a <=1-ulp pow; rounding the exact fp64 product reproduces it
    const double tmd = (double)tm * (double)tm;
    const float tm4 = (float)(tmd * tmd);
    const float dd = fs(td2, tm2);
    ci.Cq = fsqrt(fa(fm(dd, dd), fm(4.f, tm4)));
    ci.B = fa(fm(t1, t1), fm(2.f, tm2));
    ci.tdA = fd(td, A);
    ci.t0 = t0;
    return ci;
}

// One multisample point -> unit-cube grid coordinate g in [0,1]^3 and contracted std (already /2).
// render.py:L116-148 (point), coord.py:L60-72 (contract), models.py:L489-493 (/2), grid.py:L162 ((x+1)/2).
// xh (optional): the contracted mean / 2 itself (models.py:L491 `means / bound`), the `coord` the reference reports.
UC_HD void cone_point(const RayGeom& rg, const ConeInterval& ci, const ConeTable& ct, int j, int odd, float std_scale,
                      float (&g)[3], float& sigma, float* xh = nullptr) {
    const float t = fa(ci.t0, fm(ci.tdA, fa(ci.B, fm(ct.tcoef[j], ci.Cq))));
    const float rt = fm(rg.radius, t);
    // the lateral offsets (|px|,|py| ~ 1e-4 t) and the std only need relative accuracy ~1e-7: multiply by 1/sqrt(2)
    // instead of the reference's IEEE division (changes the rounding of x by < 1e-4 ulp on average)
    const float px = fm(fm(rt, ct.cosv[odd][j]), 0.70710678118f);
    const float py = fm(fm(rt, ct.sinv[odd][j]), 0.70710678118f);
    float sd = fm(fm(fm(std_scale, rg.radius), t), 0.70710678118f);
    float x[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        x[i] = fa(fa(fa(fm(px, rg.e1[i]), fm(py, rg.e2[i])), fm(t, rg.d[i])), rg.o[i]);
    const float m2 = fmaxf(fa(fa(fm(x[0], x[0]), fm(x[1], x[1])), fm(x[2], x[2])), kEps);
    if (!(m2 <= 1.f)) {
        const float mag = fsqrt(m2);
        const float k = fd(fs(fm(2.f, mag), 1.f), m2);
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = fm(k, x[i]);
        // coord.py:L68 (pow(2|x|-1, 1/3) / |x|)^2: cbrtf + fast division are within 2 ulp of the torch result, and the
        // std only enters through erf(1/(sqrt(8) std G))
#if defined(__CUDA_ARCH__)
        const float c = __fdividef(cbrtf(fs(fm(2.f, mag), 1.f)), mag);
#else
        const float c = fd(cbrtf(fs(fm(2.f, mag), 1.f)), mag);
#endif
        sd = fm(fm(c, c), sd);
    }
    sigma = fm(sd, 0.5f);
#pragma unroll
    for (int i = 0; i < 3; ++i) g[i] = fm(fa(fm(x[i], 0.5f), 1.f), 0.5f);
    if (xh) {
#pragma unroll
        for (int i = 0; i < 3; ++i) xh[i] = fm(x[i], 0.5f);
    }
}

// models.py:L512 `means.mean(dim=-2)` of one interval: the six contracted multisample means / 2, averaged ("coord",
// models.py:L677, consumed by extract.py through render_image(return_weights=True))
UC_HD void interval_coord(const RayGeom& rg, float t0, float t1, const ConeTable& ct, int odd, float std_scale, float (&c)[3]) {
    const ConeInterval ci = make_cone_interval(t0, t1);
    float acc[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < 6; ++j) {
        float g[3], sg, xh[3];
        cone_point(rg, ci, ct, j, odd, std_scale, g, sg, xh);
        for (int i = 0; i < 3; ++i) acc[i] = j == 0 ? xh[i] : fa(acc[i], xh[i]);
    }
    for (int i = 0; i < 3; ++i) c[i] = fd(acc[i], 6.f);
}

// ---- hash-grid lookup on the fused path (D=3, C=4, gridtype=hash, align_corners=False, linear) --
// gridencoder.cu:L50-84,L137-197 restated for precomputed per-level constants.
UC_HD uint32_t level_index(const GridLevel& lv, uint32_t x, uint32_t y, uint32_t z) {
    uint32_t idx;
    if (lv.hashed) idx = x ^ (y * 2654435761u) ^ (z * 805459861u);
    else idx = x + y * lv.stride1 + z * lv.stride2;
    // `index % hashmap_size` (gridencoder.cu:L83): hashed levels have power-of-two tables (AND mask); a dense
    // index of an in-range point is already < (res+1)^3 <= hashmap_size, so no reduction is needed there.
    // mod_mode: 0 = none, 1 = mask, 2 = generic modulo (non power-of-two hashed table).
    if (lv.mod_mode == 1) return idx & lv.pow2_mask;
    if (lv.mod_mode == 2) return idx % lv.hashmap_size;
    return idx;
}

struct CellCoords {
    uint32_t ix, iy, iz;
    float fx, fy, fz;
};
UC_HD CellCoords cell_of(const GridLevel& lv, const float (&g)[3]) {
    CellCoords c;
    const float px = fmaf(g[0], lv.scale, 0.5f), py = fmaf(g[1], lv.scale, 0.5f), pz = fmaf(g[2], lv.scale, 0.5f);
    const float flx = floorf(px), fly = floorf(py), flz = floorf(pz);
    c.ix = (uint32_t)flx; c.iy = (uint32_t)fly; c.iz = (uint32_t)flz;
    c.fx = px - flx; c.fy = py - fly; c.fz = pz - flz;
    return c;
}

// ---- activations ------------------------------------------------------------------------------
UC_HD float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // F.softplus, beta=1, threshold=20
UC_HD float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// ---- alpha compositing for one ray --------------------------------------------------------------
// render.py:L155-174 compute_alpha_weights + L177-244 volumetric_rendering (+ stepfun.weighted_percentile).
struct CompositeScratch {
    float* t;    // [S+2] metric fenceposts + far
    float* dd;   // [S+1] density*delta, then weights (+bg_w)
    float* cw;   // [S+2]
    UC_HD static size_t floats(int S) { return (size_t)3 * (S + 2); }
    UC_HD void carve(float* base, int S) { t = base; dd = t + (S + 2); cw = dd + (S + 2); }
};

struct RayOutputs {
    float rgb[3];
    float depth, depth_raw, acc, dist_mean, dist_median, dist_p5, dist_p95;
};

// sdist [S+1] normalised fenceposts, density [S], rgb [S*3] or NULL (proposal level: zeros),
// out_w [S] written.  `extras` selects the compute_extras branch (render.py:L218-242).
template <class X>
UC_HD void composite_ray(const X& ex, int S, const float* sdist, const float* density, const float* rgb,
                         const float* dir, float near, float far, float bg, bool extras, CompositeScratch sc,
                         float* out_w, RayOutputs& ro) {
    const int lane = ex.lane;
    constexpr int st = X::kStride;
    const float dn = norm3(dir[0], dir[1], dir[2]);  // torch.norm(dirs, dim=-1), render.py:L158
    for (int k = lane; k <= S; k += st) {
        const float s = sdist[k];
        sc.t[k] = fa(fm(s, far), fm(fs(1.f, s), near));  // coord.py:L176 s_to_t with fn=None
    }
    if (lane == 0) sc.t[S + 1] = far;
    ex.sync();
    for (int k = lane; k < S; k += st) sc.dd[k] = fm(density[k], fm(fs(sc.t[k + 1], sc.t[k]), dn));
    ex.sync();
    // exclusive fp64 cumsum of dd (torch CPU cumsum accumulates in double), stored as cw[k] = cum_{<k}
    {
#if defined(__CUDA_ARCH__)
        if constexpr (X::kStride == 32) {
            double carry = 0.0;
            for (int base = 0; base < S; base += 32) {
                const int k = base + lane;
                double v = (k < S) ? (double)sc.dd[k] : 0.0;
                const double own = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double n = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v += n;
                }
                v += carry;
                if (k < S) sc.cw[k] = (k == 0) ? 0.f : (float)(v - own);
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
        } else
#endif
        {
            if (lane == 0) {
                double acc = 0.0;
                for (int k = 0; k < S; ++k) { sc.cw[k] = (k == 0) ? 0.f : (float)acc; acc += (double)sc.dd[k]; }
            }
        }
    }
    ex.sync();
    double p_acc = 0.0, p_r = 0.0, p_g = 0.0, p_b = 0.0, p_d = 0.0, p_l = 0.0;
    for (int k = lane; k < S; k += st) {
        const float alpha = fs(1.f, expf(-sc.dd[k]));
        const float trans = expf(-sc.cw[k]);
        const float w = fm(alpha, trans);
        out_w[k] = w;
        p_acc += w;
        const float tmid = fm(0.5f, fa(sc.t[k], sc.t[k + 1]));
        p_d += fm(w, tmid);
        if (extras) p_l += fm(w, logf(tmid));
        if (rgb) { p_r += fm(w, rgb[3 * k]); p_g += fm(w, rgb[3 * k + 1]); p_b += fm(w, rgb[3 * k + 2]); }
    }
    ex.sync();
    // the weights are re-read from out_w (same lane wrote what it reads back below in the scan)
    const float acc = ex.sum(p_acc);
    const float bgw = fmaxf(fs(1.f, acc), 0.f);
    ro.acc = acc;
    ro.rgb[0] = fa(ex.sum(p_r), fm(bgw, bg));
    ro.rgb[1] = fa(ex.sum(p_g), fm(bgw, bg));
    ro.rgb[2] = fa(ex.sum(p_b), fm(bgw, bg));
    const float accc = fmaxf(acc, kEps);
    const float t_first = sc.t[0], t_last = sc.t[S];
    {
        float d = fd(ex.sum(p_d), accc);
        if (d != d) d = INFINITY;  // nan_to_num(x, inf): nan -> inf (+-inf -> +-FLT_MAX, then clipped anyway)
        d = fminf(fmaxf(d, t_first), t_last);
        ro.depth_raw = d;
        ro.depth = (acc < 0.6f) ? 300.f : d;  // render.py:L208,L213
    }
    ro.dist_mean = ro.dist_median = ro.dist_p5 = ro.dist_p95 = 0.f;
    if (extras) {
        float e = expf(fd(ex.sum(p_l), accc));
        if (e != e) e = INFINITY;
        else if (e > kFltMax) e = kFltMax;
        ro.dist_mean = fminf(fmaxf(e, t_first), t_last);
        // weighted percentiles over [tdist, far] with weights [w, bg_w] (stepfun.py:L329-339)
        ex.sync();
        for (int k = lane; k < S; k += st) sc.dd[k] = out_w[k];
        if (lane == 0) sc.dd[S] = bgw;
        ex.sync();
        cdf_scan(ex, sc.dd, S + 1, sc.cw);
        ex.sync();
        ro.dist_p5 = sorted_interp_one(0.05f, sc.cw, sc.t, S + 2);
        ro.dist_median = sorted_interp_one(0.5f, sc.cw, sc.t, S + 2);
        ro.dist_p95 = sorted_interp_one(0.95f, sc.cw, sc.t, S + 2);
    }
    ex.sync();
}

// ---- camera -> ray -----------------------------------------------------
// camera_utils.pixels_to_rays (internal/camera_utils.py:L448-557; perspective, no distortion, no NDC) evaluated as
// numpy does: float64 arithmetic on float32-valued matrices, results cast to float32 (datasets.py:L476).
UC_HD double dm(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
UC_HD double da(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

struct CameraConst {
    double pixtocam[9];    // inverse intrinsics, row-major 3x3 (float32 values in the reference's Waymo loader)
    double rot[9];         // camtoworld[:3,:3], row-major
    float origin[3];       // camtoworld[:3,3]
    float cam_dir[3];      // -camtoworld[:3,2]  (datasets.py:L446)
    float near, far;
    uint32_t width, height;
    uint64_t rand_seed;
};

struct PixelRay {
    float dir[3], view[3], plane[2], radius;
};

// direction through pixel centre (x + 0.5, y + 0.5): pixtocam @ [x, y, 1], OpenCV -> OpenGL flip (diag(1,-1,-1)),
// then camtoworld rotation; sums left to right like the reference's matmuls
UC_HD void pixel_dir(const CameraConst& c, double px, double py, double (&cam)[3], double (&d)[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) cam[i] = da(da(dm(c.pixtocam[3 * i], px), dm(c.pixtocam[3 * i + 1], py)), c.pixtocam[3 * i + 2]);
    cam[1] = -cam[1];
    cam[2] = -cam[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) d[i] = da(da(dm(c.rot[3 * i], cam[0]), dm(c.rot[3 * i + 1], cam[1])), dm(c.rot[3 * i + 2], cam[2]));
}
UC_HD double dnorm3(const double (&v)[3]) { return sqrt(da(da(dm(v[0], v[0]), dm(v[1], v[1])), dm(v[2], v[2]))); }

UC_HD void pixel_to_ray(const CameraConst& c, int x, int y, PixelRay& r) {
    double cam[3], d[3], cx[3], dx[3], cy[3], dy[3];
    pixel_dir(c, (double)x + 0.5, (double)y + 0.5, cam, d);
    pixel_dir(c, (double)(x + 1) + 0.5, (double)y + 0.5, cx, dx);
    pixel_dir(c, (double)x + 0.5, (double)(y + 1) + 0.5, cy, dy);
    const double n = dnorm3(d);
    const double ex[3] = {da(dx[0], -d[0]), da(dx[1], -d[1]), da(dx[2], -d[2])};
    const double ey[3] = {da(dy[0], -d[0]), da(dy[1], -d[1]), da(dy[2], -d[2])};
    // radii = (0.5 * (|dx - d| + |dy - d|)) * 2 / sqrt(12)      camera_utils.py:L548-556
    const double rad = dm(dm(0.5, da(dnorm3(ex), dnorm3(ey))), 2.0) / 3.4641016151377544;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.dir[i] = (float)d[i];
        r.view[i] = (float)(d[i] / n);
    }
    r.plane[0] = (float)cam[0];
    r.plane[1] = (float)cam[1];
    r.radius = (float)rad;
}

// counter-based standard normals (4 per call) for the cone-basis vector: splitmix64 -> two Box-Muller pairs
UC_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
UC_HD void normal4(uint64_t seed, uint64_t counter, float (&n)[4]) {
    const uint64_t a = splitmix64(seed ^ splitmix64(counter)), b = splitmix64(a);
    const float u0 = ((float)(uint32_t)(a >> 40) + 0.5f) * (1.0f / 16777216.0f), u1 = ((float)(uint32_t)((a >> 8) & 0xffffffu)) * (1.0f / 16777216.0f);
    const float u2 = ((float)(uint32_t)(b >> 40) + 0.5f) * (1.0f / 16777216.0f), u3 = ((float)(uint32_t)((b >> 8) & 0xffffffu)) * (1.0f / 16777216.0f);
    const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
    n[0] = r0 * cosf(6.2831853f * u1); n[1] = r0 * sinf(6.2831853f * u1);
    n[2] = r1 * cosf(6.2831853f * u3); n[3] = r1 * sinf(6.2831853f * u3);
}

}  // namespace ucnerf

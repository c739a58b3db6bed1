// Host side of the sky head: handle, weight re-layout for sky_mlp_tc.cu, chunked launch
// sequence.  Replaces models.py:L326-337 (ray_batch assembly + render_rays(network_fn=skynerf)) behind the C ABI
// ucnerf_sky_* (include/ucnerf_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/ucnerf_b200.h"
#include "ray_march.cuh"

namespace ucnerf {

struct SkyBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        UC_CUDA_OK(cudaMalloc(&p, bytes));
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

static int sky_fetch(std::vector<float>& dst, const float* dev, size_t n) {
    dst.resize(n);
    UC_REQUIRE(dev != nullptr, "sky: null weight pointer");
    UC_CUDA_OK(cudaMemcpy(dst.data(), dev, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
static int sky_upload(SkyBuf& b, const void* src, size_t bytes) {
    if (int e = b.ensure(bytes)) return e;
    UC_CUDA_OK(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}

// power-of-two factor that moves the largest |w| into [2^13, 2^14) (see color_tc_weight_scale)
static float pow2_scale(const std::vector<float>& w) {
    float mx = 0.f;
    for (float v : w) mx = fmaxf(mx, fabsf(v));
    if (!(mx > 0.f) || !std::isfinite(mx)) return 1.f;
    int e;
    frexpf(mx, &e);
    return ldexpf(1.f, 14 - e);
}

}  // namespace ucnerf

using namespace ucnerf;

struct ucnerf_sky {
    int n_samples = 120;
    SkyBuf wblob, wblob2, w0x, w5x, bias8, w_alpha, rgb_w, wv_view, bv, t_vals, view_bias, raw;
    int pipeline = 1;   // 1: sky_mlp_tc.cu (default), 2: half-pass pipeline (sky_mlp_tc2.cu: measured 5 % slower, kept as an
                        // experiment; env UCNERF_SKY_PIPELINE=2)
    float k[10];
    float b_alpha = 0.f, rgb_b[3] = {0, 0, 0};
    int64_t chunk_rays = 262144;
    std::mutex mu;
};

extern "C" int ucnerf_sky_destroy(ucnerf_sky* s) {
    if (!s) return 0;
    for (SkyBuf* b : {&s->wblob, &s->wblob2, &s->w0x, &s->w5x, &s->bias8, &s->w_alpha, &s->rgb_w, &s->wv_view, &s->bv, &s->t_vals, &s->view_bias, &s->raw})
        b->release();
    delete s;
    return 0;
}

extern "C" int ucnerf_sky_create(const ucnerf_sky_desc* d, ucnerf_sky** out) {
    UC_REQUIRE(d && out, "sky_create: null argument");
    UC_REQUIRE(d->n_samples >= 2 && d->n_samples <= 4096, "sky_create: n_samples out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("ucnerf_b200 requires a CUDA device (sm_100a); none is visible - there is no CPU fallback");
        return 4;
    }
    ucnerf_sky* s = new ucnerf_sky();
    s->n_samples = d->n_samples;
    auto fail = [&](int e) { ucnerf_sky_destroy(s); return e; };
    const float act = sky_tc_act_scale();
    // ---- fetch (nn.Linear layouts [out, in]) ----
    std::vector<float> W[8], B[8], Wf, Bf, Wa, Ba, Wv, Bv, Wr, Br;
    for (int l = 0; l < 8; ++l) {
        const int fin = l == 0 ? 3 : (l == 5 ? 259 : 256);
        if (int e = sky_fetch(W[l], d->pts_w[l], (size_t)256 * fin)) return fail(e);
        if (int e = sky_fetch(B[l], d->pts_b[l], 256)) return fail(e);
    }
    if (int e = sky_fetch(Wf, d->feature_w, 256 * 256)) return fail(e);
    if (int e = sky_fetch(Bf, d->feature_b, 256)) return fail(e);
    if (int e = sky_fetch(Wa, d->alpha_w, 256)) return fail(e);
    if (int e = sky_fetch(Ba, d->alpha_b, 1)) return fail(e);
    if (int e = sky_fetch(Wv, d->views_w, (size_t)128 * 283)) return fail(e);
    if (int e = sky_fetch(Bv, d->views_b, 128)) return fail(e);
    if (int e = sky_fetch(Wr, d->rgb_w, 3 * 128)) return fail(e);
    if (int e = sky_fetch(Br, d->rgb_b, 3)) return fail(e);
    // ---- fold the activation-free feature layer into the view layer (fp64): W' = W_vf W_f, b' = b_v + W_vf b_f ----
    std::vector<float> wfold((size_t)128 * 256), bfold(128);
    for (int n = 0; n < 128; ++n) {
        double bacc = Bv[n];
        for (int j = 0; j < 256; ++j) bacc += (double)Wv[(size_t)n * 283 + j] * (double)Bf[j];
        bfold[n] = (float)bacc;
        for (int k = 0; k < 256; ++k) {
            double acc = 0.0;
            for (int j = 0; j < 256; ++j) acc += (double)Wv[(size_t)n * 283 + j] * (double)Wf[(size_t)j * 256 + k];
            wfold[(size_t)n * 256 + k] = (float)acc;
        }
    }
    // ---- 34 weight chunks in step order (sky_mlp_tc.cu header) ----
    const int nsteps = sky_tc_steps();
    std::vector<uint8_t> blob(sky_tc_blob_bytes(), 0);
    const size_t stride = blob.size() / nsteps;
    std::vector<float> wt((size_t)64 * 256);
    int step = 0;
    float sw[9];
    for (int l = 0; l < 8; ++l) sw[l] = pow2_scale(W[l]);
    sw[8] = pow2_scale(wfold);
    auto pack_block = [&](const std::vector<float>& w, int fin, int col0, int kcount, int n_cols, float scale) {
        // Wt[k][n] = w[n][col0 + k] for k < kcount, zero beyond
        std::fill(wt.begin(), wt.end(), 0.f);
        for (int k = 0; k < kcount; ++k)
            for (int n = 0; n < n_cols; ++n) wt[(size_t)k * n_cols + n] = w[(size_t)n * fin + col0 + k];
        sky_tc_pack_chunk(wt.data(), n_cols, scale, blob.data() + stride * step);
        ++step;
    };
    pack_block(W[0], 3, 0, 3, 256, sw[0]);                                         // step 0: xyz
    for (int l = 1; l <= 4; ++l)
        for (int j = 0; j < 4; ++j) pack_block(W[l], 256, 64 * j, 64, 256, sw[l]);  // steps 1..16
    pack_block(W[5], 259, 0, 3, 256, sw[5]);                                        // step 17: xyz part of the skip layer
    for (int j = 0; j < 4; ++j) pack_block(W[5], 259, 3 + 64 * j, 64, 256, sw[5]);  // steps 18..21
    for (int l = 6; l <= 7; ++l)
        for (int j = 0; j < 4; ++j) pack_block(W[l], 256, 64 * j, 64, 256, sw[l]);  // steps 22..29
    for (int j = 0; j < 4; ++j) pack_block(wfold, 256, 64 * j, 64, 128, sw[8]);     // steps 30..33: folded view layer (N = 128)
    if (step != nsteps) { set_error("sky_create: internal chunk count"); return fail(1); }
    // ---- second pipeline: 60 half-chunks (128 output columns x 64 K) in the order its MMA thread consumes them:
    //      layers 1..7: pass h (columns 128 h ..) x K chunk j; then the folded view layer (N = 128) ----
    std::vector<uint8_t> blob2(sky_tc2_blob_bytes(), 0);
    {
        const size_t stride2 = blob2.size() / sky_tc2_half_steps();
        std::vector<float> wt2((size_t)64 * 128);
        int hs = 0;
        auto pack_half = [&](const std::vector<float>& w, int fin, int col0, int n0, float scale) {
            for (int k = 0; k < 64; ++k)
                for (int n = 0; n < 128; ++n) wt2[(size_t)k * 128 + n] = w[(size_t)(n0 + n) * fin + col0 + k];
            sky_tc_pack_chunk(wt2.data(), 128, scale, blob2.data() + stride2 * hs);
            ++hs;
        };
        for (int l = 1; l <= 7; ++l)
            for (int h = 0; h < 2; ++h)
                for (int j = 0; j < 4; ++j) pack_half(W[l], l == 5 ? 259 : 256, (l == 5 ? 3 : 0) + 64 * j, 128 * h, sw[l]);
        for (int j = 0; j < 4; ++j) pack_half(wfold, 256, 64 * j, 0, sw[8]);
        if (hs != sky_tc2_half_steps()) { set_error("sky_create: internal half-chunk count"); return fail(1); }
    }
    std::vector<float> w0x((size_t)256 * 4), w5x((size_t)256 * 4);
    for (int c = 0; c < 256; ++c) {
        for (int i = 0; i < 3; ++i) {
            w0x[(size_t)4 * c + i] = act * W[0][(size_t)c * 3 + i];
            w5x[(size_t)4 * c + i] = act * W[5][(size_t)c * 259 + i];
        }
        w0x[(size_t)4 * c + 3] = act * B[0][c];
        w5x[(size_t)4 * c + 3] = 0.f;
    }
    if (const char* e = getenv("UCNERF_SKY_PIPELINE")) s->pipeline = atoi(e) == 2 ? 2 : 1;
    for (int l = 0; l < 8; ++l) s->k[l] = 1.f / sw[l];
    s->k[8] = 1.f / (act * sw[8]);
    s->k[9] = 0.f;
    std::vector<float> bias8((size_t)8 * 256);
    for (int l = 0; l < 8; ++l)
        for (int c = 0; c < 256; ++c) bias8[(size_t)l * 256 + c] = B[l][c] * act;
    std::vector<float> rgbw((size_t)128 * 4, 0.f), wvv((size_t)27 * 128);
    for (int k = 0; k < 128; ++k)
        for (int c = 0; c < 3; ++c) rgbw[(size_t)k * 4 + c] = Wr[(size_t)c * 128 + k];
    for (int k = 0; k < 27; ++k)
        for (int n = 0; n < 128; ++n) wvv[(size_t)k * 128 + n] = Wv[(size_t)n * 283 + 256 + k];
    s->b_alpha = Ba[0];
    for (int c = 0; c < 3; ++c) s->rgb_b[c] = Br[c];
    // torch.linspace(0., 1., n_samples) in float32 (symmetric evaluation, see torch_linspace_f32 in model.cu)
    std::vector<float> tv(s->n_samples);
    {
        const int steps = s->n_samples;
        const float st = (1.f - 0.f) / (float)(steps - 1);
        for (int i = 0; i < steps; ++i) tv[i] = i < steps / 2 ? std::fmaf(st, (float)i, 0.f) : std::fmaf(-st, (float)(steps - i - 1), 1.f);
    }
    if (int e = sky_upload(s->wblob, blob.data(), blob.size())) return fail(e);
    if (int e = sky_upload(s->wblob2, blob2.data(), blob2.size())) return fail(e);
    if (int e = sky_upload(s->w0x, w0x.data(), w0x.size() * 4)) return fail(e);
    if (int e = sky_upload(s->w5x, w5x.data(), w5x.size() * 4)) return fail(e);
    if (int e = sky_upload(s->bias8, bias8.data(), bias8.size() * 4)) return fail(e);
    if (int e = sky_upload(s->w_alpha, Wa.data(), 256 * 4)) return fail(e);
    if (int e = sky_upload(s->rgb_w, rgbw.data(), rgbw.size() * 4)) return fail(e);
    if (int e = sky_upload(s->wv_view, wvv.data(), wvv.size() * 4)) return fail(e);
    if (int e = sky_upload(s->bv, bfold.data(), 128 * 4)) return fail(e);
    if (int e = sky_upload(s->t_vals, tv.data(), tv.size() * 4)) return fail(e);
    *out = s;
    return 0;
}

extern "C" int ucnerf_sky_render(ucnerf_sky* s, uint64_t n_rays, const float* origins, const float* directions,
                                 const float* far, const float* views, double sky_far, float* sky_rgb, void* stream) {
    UC_REQUIRE(s && origins && directions && far && views && sky_rgb, "sky_render: null argument");
    if (n_rays == 0) return 0;
    std::lock_guard<std::mutex> lk(s->mu);
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t chunk = (uint64_t)s->chunk_rays;
    for (uint64_t r0 = 0; r0 < n_rays; r0 += chunk) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(chunk, n_rays - r0);
        if (int e = s->view_bias.ensure((size_t)n * 128 * sizeof(float))) return e;
        if (int e = s->raw.ensure((size_t)n * s->n_samples * 4 * sizeof(float))) return e;
        if (int e = launch_sky_view_bias(views + 3 * r0, s->wv_view.as<float>(), s->bv.as<float>(), s->view_bias.as<float>(), n, st)) return e;
        SkyTcParams p{};
        p.n_rows = n * (uint32_t)s->n_samples; p.n_samples = s->n_samples;
        p.origins = origins + 3 * r0; p.directions = directions + 3 * r0; p.far = far + r0;
        p.t_vals = s->t_vals.as<float>(); p.sky_far = (float)sky_far; p.view_bias = s->view_bias.as<float>();
        p.wblob = s->wblob.as<uint8_t>(); p.bias8 = s->bias8.as<float>();
        std::memcpy(p.k, s->k, sizeof(p.k));
        p.w_alpha = s->w_alpha.as<float>(); p.b_alpha = s->b_alpha; p.rgb_w = s->rgb_w.as<float>();
        std::memcpy(p.rgb_b, s->rgb_b, sizeof(p.rgb_b));
        p.raw = s->raw.as<float>();
        p.wblob2 = s->wblob2.as<uint8_t>(); p.w0x = s->w0x.as<float>(); p.w5x = s->w5x.as<float>();
        if (s->pipeline == 2) {
            uint32_t* dbg = sky_tc_dbg_buffer();
            UC_REQUIRE(dbg != nullptr, "sky: cannot allocate the watchdog record");
            if (int e = launch_sky_mlp_tc2(p, dbg, st)) return e;
        } else if (int e = launch_sky_mlp_tc(p, st)) return e;
        if (int e = launch_sky_composite(s->raw.as<float>(), directions + 3 * r0, far + r0, s->t_vals.as<float>(), (float)sky_far,
                                         s->n_samples, sky_rgb + 3 * r0, n, st)) return e;
    }
    return 0;
}

// Watchdog record of the sky tensor-core kernel (synchronises the device): out32[0] != 0 means a pipeline wait timed out.
extern "C" int ucnerf_debug_sky_status(uint32_t* out32) {
    UC_REQUIRE(out32, "debug_sky_status: null");
    return sky_tc_status(out32);
}

// Host side of the fused render path: model handle, weight re-layout, chunked launch sequence and
// the HOST-buffer entry used for end-to-end timing.  C ABI in include/ucnerf_b200.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/ucnerf_b200.h"
#include "ray_march.cuh"

namespace ucnerf {

static thread_local std::string g_err;
std::atomic<uint64_t> g_launch_count{0};
void set_error(const std::string& msg) { g_err = msg; }

// torch.linspace(start, end, steps) for float32 on CPU (ATen RangeFactoriesKernel): step in fp32,
// symmetric evaluation from both ends, fused multiply-add (checked bit-exact in tests/test_host_logic.py).  stepfun.py:L203-204 deterministic_center u grid.
static void torch_linspace_f32(float start, float end, int steps, float* out) {
    if (steps == 1) { out[0] = start; return; }
    const float step = (end - start) / (float)(steps - 1);
    const int halfway = steps / 2;
    for (int i = 0; i < steps; ++i) {
        // the vectorised ATen kernel evaluates both branches with a fused multiply-add
        out[i] = (i < halfway) ? std::fmaf(step, (float)i, start) : std::fmaf(-step, (float)(steps - i - 1), end);
    }
}

static void deterministic_u(int S, float* out) {
    const double pad = 1.0 / (2.0 * S);
    const double eps = (double)kEps;
    torch_linspace_f32((float)pad, (float)(1.0 - pad - eps), S, out);
}

struct DevBuf {
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

struct LevelState {
    GridDesc grid;
    float g2[16];
    int lmax = 0;
    int S = 0;
    DevBuf w1p, b1, w2, u;
    float b2 = 0.f;
    DevBuf sdist, weights;  // workspace when the caller does not ask for them
};

}  // namespace ucnerf

using namespace ucnerf;

struct ucnerf_model {
    ucnerf_model_desc d;
    std::vector<int32_t> offsets[UCNERF_MAX_PROP_LEVELS + 1], grid_sizes[UCNERF_MAX_PROP_LEVELS + 1];
    int num_levels = 0;  // sampling levels = prop + 1
    LevelState lv[UCNERF_MAX_PROP_LEVELS + 1];
    ConeTable cone;
    int np = 0;  // padded colour-MLP width
    DevBuf w2t, b2, v0t, c0, v1t, c1, rt, r0, wblob, wdir, dir_bias;
    uint32_t tc_debug = 0;
    float tc_k0 = 1.f, tc_k1 = 1.f;   // accumulator scales of the FP16-split tensor-core colour MLP
    bool tc_ok = false;   // tensor-core colour MLP available for these shapes (W = 256, deg_view = 4)
    DevBuf density, h1, rgb_s, geom;
    // host-entry staging
    DevBuf stage_in, stage_out, cam_rays;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // host entry: H2D / D2H overlap the render stream chunk by chunk
    std::vector<cudaEvent_t> ev_in, ev_done;
    int64_t chunk_rays = 131072;
    int64_t ray_tile_width = 0;  // > 0: ray batches are whole rows of a row-major image of this width (see SampleParams::tile_w)
    int ray_tile_patch[2] = {0, 4};
    int ray_geom = 1;  // per-ray cone basis precomputed once per chunk (ray_geom_kernel) instead of per sample  // patch width per level kind {proposal, NeRF}: 0 = rows of 32 pixels, 4 / 8 / 16 = 4x8 / 8x4 / 16x2
    int color_mode = 2;   // 0 = fp32 SIMT, 1 = tcgen05 FP16 split (error if shapes unsupported), 2 = auto
    bool use_affine = false;
    float affine[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};   // BrightnessCorrection affine of the current image (row-major 3x4)
    uint32_t n_peers = 0;                 // fused tile exchange (ucnerf_set_peer_targets)
    float* peer_images[UCNERF_MAX_PEERS] = {};
    uint64_t peer_row0 = 0;
    int encode_runs = 0;  // cell-run reuse in sample_encode_kernel (bit 0 = proposal levels, bit 1 = NeRF level): measured
                          // slower on B200 (profiles/r1_summary.md), kept as an option
    int encode_mlp_mma = 3;  // density layer of sample_encode_kernel on mma.sync 3xTF32 (bit 0 = proposal levels, bit 1 = NeRF level)
    int warp_rays_log2[2] = {5, 5};  // sample_encode_kernel warp shape {proposal levels, NeRF level}: 2^k rays x 2^(5-k) samples
    bool timing = false;
    float ms[5] = {0, 0, 0, 0, 0};
    uint32_t nlaunch[5] = {0, 0, 0, 0, 0};
    struct Timed { cudaEvent_t a, b; int slot; };
    std::vector<Timed> pending;            // recorded, not yet resolved (no sync on the launch path)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
    cudaEvent_t cur_a = nullptr, cur_b = nullptr;
    std::mutex mu;
};

namespace ucnerf {

static int fetch_host(std::vector<float>& dst, const float* dev, size_t n) {
    dst.resize(n);
    UC_REQUIRE(dev != nullptr, "model: null weight pointer");
    UC_CUDA_OK(cudaMemcpy(dst.data(), dev, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
static int upload(DevBuf& b, const std::vector<float>& src) {
    if (int e = b.ensure(src.size() * sizeof(float))) return e;
    UC_CUDA_OK(cudaMemcpy(b.p, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

static int pad_width(int w) {
    for (int o : {32, 64, 128, 256})
        if (w <= o) return o;
    return 0;
}

// Per-level constants exactly as the reference kernel derives them (gridencoder.cu:L66-84,L137-139).
static int build_grid(ucnerf_model* m, int li, const ucnerf_mlp_desc& md) {
    LevelState& ls = m->lv[li];
    const int L = md.grid_levels;
    UC_REQUIRE(L >= 1 && L <= UCNERF_MAX_GRID_LEVELS, "model: grid_levels must be in [1,16]");
    UC_REQUIRE(md.level_dim == 4, "model: the fused path requires level_dim == 4");
    UC_REQUIRE(md.embeddings && md.offsets_host && md.grid_sizes_host, "model: null grid pointer");
    m->offsets[li].assign(md.offsets_host, md.offsets_host + L + 1);
    m->grid_sizes[li].assign(md.grid_sizes_host, md.grid_sizes_host + L);
    ls.grid.table = reinterpret_cast<const float4*>(md.embeddings);
    ls.grid.num_levels = L;
    for (int l = 0; l < L; ++l) {
        const int64_t gs = m->grid_sizes[li][l];
        make_grid_level(ls.grid.lv[l], l, m->offsets[li][l], m->offsets[li][l + 1], md.log2_per_level_scale,
                        (uint32_t)md.base_resolution, gs);
        ls.g2[l] = (float)(int32_t)(gs * gs);  // torch: int32 grid_sizes ** 2, promoted to fp32 in the product
    }
    ls.lmax = sample_encode_lmax(L);
    UC_REQUIRE(ls.lmax > 0, "model: unsupported number of grid levels");
    return 0;
}

static int build_density_layer(ucnerf_model* m, int li, const ucnerf_mlp_desc& md, int out_dim) {
    LevelState& ls = m->lv[li];
    const int L = md.grid_levels, LC = L * 4, LP = ls.lmax * 4;
    std::vector<float> w0, b0, w2, b2;
    if (int e = fetch_host(w0, md.density0_w, (size_t)64 * LC)) return e;
    if (int e = fetch_host(b0, md.density0_b, 64)) return e;
    if (int e = fetch_host(w2, md.density2_w, (size_t)out_dim * 64)) return e;
    if (int e = fetch_host(b2, md.density2_b, out_dim)) return e;
    std::vector<float> w1p((size_t)64 * LP, 0.f);
    for (int j = 0; j < 64; ++j)
        for (int k = 0; k < LC; ++k) w1p[(size_t)j * LP + k] = w0[(size_t)j * LC + k];
    if (int e = upload(ls.w1p, w1p)) return e;
    if (int e = upload(ls.b1, b0)) return e;
    std::vector<float> row0(w2.begin(), w2.begin() + 64);
    if (int e = upload(ls.w2, row0)) return e;
    ls.b2 = b2[0];
    return 0;
}

static int build_color(ucnerf_model* m) {
    const ucnerf_model_desc& d = m->d;
    const int BW = d.bottleneck_width, W = d.net_width_viewdirs, ND = 3 + 6 * d.deg_view;
    UC_REQUIRE(d.deg_view >= 0 && ND <= 32, "model: deg_view must be <= 4");
    UC_REQUIRE(BW >= 1 && W >= 1, "model: bad MLP widths");
    const int NP = pad_width(std::max(BW, W));
    UC_REQUIRE(NP > 0, "model: bottleneck_width / net_width_viewdirs must be <= 256");
    m->np = NP;
    const int DIN = BW + ND;
    std::vector<float> w2, b2, v0, c0, v1, c1, r, r0;
    if (int e = fetch_host(w2, d.nerf.density2_w, (size_t)BW * 64)) return e;
    if (int e = fetch_host(b2, d.nerf.density2_b, BW)) return e;
    if (int e = fetch_host(v0, d.view0_w, (size_t)W * DIN)) return e;
    if (int e = fetch_host(c0, d.view0_b, W)) return e;
    if (int e = fetch_host(v1, d.view1_w, (size_t)W * (W + DIN))) return e;
    if (int e = fetch_host(c1, d.view1_b, W)) return e;
    if (int e = fetch_host(r, d.rgb_w, (size_t)3 * W)) return e;
    if (int e = fetch_host(r0, d.rgb_b, 3)) return e;
    // Fold the activation-free bottleneck layer x = W2 h1 + b2 into its consumers (fp64 on the host):
    //   V0 [x, dir] + c0 = (V0x W2) h1 + V0d dir + (c0 + V0x b2)        (likewise for V1's x block)
    // K-major ("transposed") zero-padded layouts Wt[k][n]; K rows: [h1 (64) | direnc (32)].
    std::vector<float> p0t((size_t)96 * NP, 0.f), c0p(NP, 0.f);
    for (int n = 0; n < W; ++n) {
        const float* vr = &v0[(size_t)n * DIN];
        double bacc = c0[n];
        for (int j = 0; j < BW; ++j) bacc += (double)vr[j] * (double)b2[j];
        c0p[n] = (float)bacc;
        for (int k = 0; k < 64; ++k) {  // h1 column k holds hidden unit h1_perm(k)
            const int hu = h1_perm(k);
            double acc = 0.0;
            for (int j = 0; j < BW; ++j) acc += (double)vr[j] * (double)w2[(size_t)j * 64 + hu];
            p0t[(size_t)k * NP + n] = (float)acc;
        }
        for (int k = 0; k < ND; ++k) p0t[(size_t)(64 + k) * NP + n] = vr[BW + k];
    }
    std::vector<float> v1t((size_t)(NP + 96) * NP, 0.f), c1p(NP, 0.f);
    for (int n = 0; n < W; ++n) {
        const float* src = &v1[(size_t)n * (W + DIN)];
        double bacc = c1[n];
        for (int j = 0; j < BW; ++j) bacc += (double)src[W + j] * (double)b2[j];
        c1p[n] = (float)bacc;
        for (int k = 0; k < W; ++k) v1t[(size_t)k * NP + n] = src[k];
        for (int k = 0; k < 64; ++k) {
            const int hu = h1_perm(k);
            double acc = 0.0;
            for (int j = 0; j < BW; ++j) acc += (double)src[W + j] * (double)w2[(size_t)j * 64 + hu];
            v1t[(size_t)(NP + k) * NP + n] = (float)acc;
        }
        for (int k = 0; k < ND; ++k) v1t[(size_t)(NP + 64 + k) * NP + n] = src[W + BW + k];
    }
    std::vector<float> rt((size_t)NP * 4, 0.f), r0p(4, 0.f);
    for (int c = 0; c < 3; ++c) {
        r0p[c] = r0[c];
        for (int k = 0; k < W; ++k) rt[(size_t)k * 4 + c] = r[(size_t)c * W + k];
    }
    if (int e = upload(m->v0t, p0t)) return e;
    m->tc_ok = (NP == 256 && d.deg_view == 4);
    if (m->tc_ok) {
        // step order of color_mlp_tc_kernel: P0 (h1 rows) -> acc3; P1 (h1 rows), V1a x4 -> acc4.  Both operand blocks
        // of one accumulator share one power-of-two weight scale; activations carry color_tc_act_scale().
        std::vector<uint8_t> blob(color_tc_blob_bytes());
        const size_t cb = blob.size() / 6;
        const float sw0 = color_tc_weight_scale(&p0t[0], (size_t)64 * NP);
        const float sw1 = color_tc_weight_scale(&v1t[0], (size_t)(NP + 64) * NP);  // rows [a (NP) | h1 (64)]
        color_tc_pack_chunk(&p0t[0], sw0, blob.data());
        color_tc_pack_chunk(&v1t[(size_t)NP * NP], sw1, blob.data() + cb);
        for (int j = 0; j < 4; ++j) color_tc_pack_chunk(&v1t[(size_t)(64 * j) * NP], sw1, blob.data() + cb * (2 + j));
        m->tc_k0 = 1.f / sw0;
        m->tc_k1 = 1.f / (color_tc_act_scale() * sw1);
        // view-direction rows of both layers for dir_bias_kernel: [2][32][256]
        std::vector<float> wd((size_t)2 * 32 * 256, 0.f);
        for (int k = 0; k < 32; ++k)
            for (int n = 0; n < 256; ++n) {
                wd[(size_t)k * 256 + n] = p0t[(size_t)(64 + k) * NP + n];
                wd[(size_t)(32 + k) * 256 + n] = v1t[(size_t)(NP + 64 + k) * NP + n];
            }
        if (int e = upload(m->wdir, wd)) return e;
        if (int e = m->wblob.ensure(blob.size())) return e;
        UC_CUDA_OK(cudaMemcpy(m->wblob.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    }
    if (int e = upload(m->c0, c0p)) return e;
    if (int e = upload(m->v1t, v1t)) return e;
    if (int e = upload(m->c1, c1p)) return e;
    if (int e = upload(m->rt, rt)) return e;
    if (int e = upload(m->r0, r0p)) return e;
    return 0;
}

static int build_model(ucnerf_model* m, const ucnerf_model_desc* desc) {
    m->d = *desc;
    const ucnerf_model_desc& d = m->d;
    UC_REQUIRE(d.num_prop_levels >= 1 && d.num_prop_levels <= UCNERF_MAX_PROP_LEVELS, "model: num_prop_levels must be in [1,4]");
    UC_REQUIRE(d.num_prop_samples >= 2 && d.num_nerf_samples >= 2, "model: num_samples must be > 1");  // stepfun.py:L268
    m->num_levels = d.num_prop_levels + 1;
    make_cone_table(m->cone);
    for (int li = 0; li < m->num_levels; ++li) {
        const bool nerf = li == m->num_levels - 1;
        const ucnerf_mlp_desc& md = nerf ? d.nerf : d.prop[li];
        LevelState& ls = m->lv[li];
        ls.S = nerf ? d.num_nerf_samples : d.num_prop_samples;
        if (int e = build_grid(m, li, md)) return e;
        if (int e = build_density_layer(m, li, md, nerf ? d.bottleneck_width : 1)) return e;
        std::vector<float> u(ls.S);
        deterministic_u(ls.S, u.data());
        if (int e = upload(ls.u, u)) return e;
    }
    if (int e = build_color(m)) return e;
    return 0;
}

// Per-kernel-family timing: an event pair is recorded around each launch on the launch stream and resolved
// later in ucnerf_get_timing, so enabling it adds no synchronisation to the timed region.
static int time_begin(ucnerf_model* m, cudaStream_t st) {
    if (!m->timing) return 0;
    if (m->pool.empty()) {
        cudaEvent_t a, b;
        UC_CUDA_OK(cudaEventCreate(&a));
        UC_CUDA_OK(cudaEventCreate(&b));
        m->pool.emplace_back(a, b);
    }
    m->cur_a = m->pool.back().first;
    m->cur_b = m->pool.back().second;
    m->pool.pop_back();
    UC_CUDA_OK(cudaEventRecord(m->cur_a, st));
    return 0;
}
static int time_end(ucnerf_model* m, cudaStream_t st, int slot) {
    if (!m->timing) return 0;
    UC_CUDA_OK(cudaEventRecord(m->cur_b, st));
    m->pending.push_back({m->cur_a, m->cur_b, slot});
    return 0;
}
static int resolve_timing(ucnerf_model* m) {
    for (auto& t : m->pending) {
        UC_CUDA_OK(cudaEventSynchronize(t.b));
        float ms = 0.f;
        UC_CUDA_OK(cudaEventElapsedTime(&ms, t.a, t.b));
        m->ms[t.slot] += ms;
        m->nlaunch[t.slot] += 1;
        m->pool.emplace_back(t.a, t.b);
    }
    m->pending.clear();
    return 0;
}

static int render_chunk(ucnerf_model* m, uint32_t n, const ucnerf_rays& r, size_t ray0, double train_frac,
                        const ucnerf_outputs& o, cudaStream_t st, uint64_t tile_w) {
    const ucnerf_model_desc& d = m->d;
    RayPtrs rp{r.origins + 3 * ray0, r.directions + 3 * ray0, r.viewdirs + 3 * ray0, r.cam_dirs + 3 * ray0,
               r.radii + ray0, r.near + ray0, r.far + ray0, r.rand_vec + 3 * ray0};
    // models.py:L179-184 Schlick bias anneal
    float anneal = 1.f;
    if (d.anneal_slope > 0.0) {
        const double s = d.anneal_slope, x = train_frac;
        anneal = (float)((s * x) / ((s - 1.0) * x + 1.0));
    }
    int smax = 0;
    for (int li = 0; li < m->num_levels; ++li) smax = std::max(smax, m->lv[li].S);
    if (int e = m->density.ensure((size_t)n * smax * sizeof(float))) return e;
    // the cone basis of every ray of the chunk, once (all levels and samples read it)
    const uint32_t geom_ld = (n + 31u) & ~31u;
    const float* geom = nullptr;
    if (m->ray_geom) {
        if (int e = m->geom.ensure((size_t)15 * geom_ld * sizeof(float))) return e;
        if (int e = launch_ray_geom(rp, n, m->geom.as<float>(), geom_ld, st)) return e;
        geom = m->geom.as<float>();
    }

    const float* t_prev = nullptr;
    const float* w_prev = nullptr;
    uint32_t t_prev_stride = 0;
    int n_prev = 1;
    double prod = 1.0;
    for (int li = 0; li < m->num_levels; ++li) {
        LevelState& ls = m->lv[li];
        const bool nerf = li == m->num_levels - 1;
        const int S = ls.S;
        float* sdist = o.sdist[li] ? o.sdist[li] + ray0 * (S + 1) : nullptr;
        float* weights = o.weights[li] ? o.weights[li] + ray0 * S : nullptr;
        // The first level resamples sdist=[0,1], w=[1] (models.py:L143-147): its fenceposts are identical for every
        // ray, so unless the caller wants them written out they are computed for ONE ray and shared (stride 0).
        const bool shared_row = (li == 0 && sdist == nullptr);
        uint32_t sdist_stride = (uint32_t)(S + 1);
        if (!sdist) {
            if (int e = ls.sdist.ensure((size_t)(shared_row ? 1 : n) * (S + 1) * sizeof(float))) return e;
            sdist = ls.sdist.as<float>();
            if (shared_row) sdist_stride = 0;
        }
        if (!weights) {
            if (int e = ls.weights.ensure((size_t)n * S * sizeof(float))) return e;
            weights = ls.weights.as<float>();
        }
        // models.py:L158-162
        const float dilation = (float)(d.dilation_bias + d.dilation_multiplier * (1.0 - 0.0) / prod);
        prod *= S;
        const bool use_dil = (d.dilation_bias > 0.0 || d.dilation_multiplier > 0.0) && li > 0;

        ResampleParams rs{};
        rs.n_rays = shared_row ? 1u : n; rs.n_prev = n_prev; rs.t_prev = t_prev; rs.t_prev_stride = t_prev_stride;
        rs.w_prev = w_prev; rs.dilate = use_dil ? 1 : 0;
        rs.dilation = dilation; rs.anneal = anneal; rs.padding = (float)d.resample_padding; rs.S = S;
        rs.u = ls.u.as<float>(); rs.out_sdist = sdist;
        if (int e = time_begin(m, st)) return e;
        if (int e = launch_resample(rs, st)) return e;
        if (int e = time_end(m, st, 0)) return e;

        SampleParams sp{};
        sp.n_rays = n; sp.S = S; sp.rays = rp; sp.sdist = sdist; sp.sdist_stride = sdist_stride; sp.grid = ls.grid;
        sp.cone = m->cone;
        sp.std_scale = (float)d.std_scale; sp.density_bias = (float)d.density_bias;
        sp.w1p = ls.w1p.as<float>(); sp.b1 = ls.b1.as<float>(); sp.w2 = ls.w2.as<float>(); sp.b2 = ls.b2;
        sp.density = (nerf && o.sample_density) ? o.sample_density + ray0 * S : m->density.as<float>();
        std::memcpy(sp.g2, ls.g2, sizeof(sp.g2));
        sp.cell_runs = (m->encode_runs >> (nerf ? 1 : 0)) & 1;
        sp.mlp_mma = (m->encode_mlp_mma >> (nerf ? 1 : 0)) & 1;
        sp.geom = geom; sp.geom_ld = geom_ld;
        const int patch = m->ray_tile_patch[nerf ? 1 : 0];
        sp.tile_w = (patch > 0 && tile_w > 0 && tile_w % (uint64_t)patch == 0 && ray0 % (size_t)tile_w == 0) ? (uint32_t)tile_w : 0u;
        sp.tile_pw_log2 = patch == 16 ? 4u : patch == 8 ? 3u : 2u;
        sp.rw_log2 = m->warp_rays_log2[nerf ? 1 : 0];
        if (S % (32 >> sp.rw_log2) != 0) sp.rw_log2 = 5;   // sample blocks must tile S
        float* rgb_s = nullptr;
        if (nerf) {
            if (int e = m->h1.ensure((size_t)n * S * 64 * sizeof(float))) return e;
            sp.h1 = m->h1.as<float>();
            if (o.sample_rgb) rgb_s = o.sample_rgb + ray0 * S * 3;
            else {
                if (int e = m->rgb_s.ensure((size_t)n * S * 3 * sizeof(float))) return e;
                rgb_s = m->rgb_s.as<float>();
            }
        }
        if (int e = time_begin(m, st)) return e;
        if (int e = launch_sample_encode(sp, nerf, st)) return e;
        if (int e = time_end(m, st, nerf ? 2 : 1)) return e;
        if (nerf && o.sample_coord)
            if (int e = launch_sample_coord(sp, o.sample_coord + ray0 * S * 3, st)) return e;

        if (nerf) {
            ColorParams cp{};
            cp.n_rows = n * (uint32_t)S; cp.S = S; cp.deg_view = d.deg_view; cp.h1 = sp.h1; cp.viewdirs = rp.viewdirs;
            cp.p0t = m->v0t.as<float>(); cp.c0 = m->c0.as<float>();
            cp.v1t = m->v1t.as<float>(); cp.c1 = m->c1.as<float>(); cp.rt = m->rt.as<float>(); cp.r0 = m->r0.as<float>();
            cp.rgb_scale = (float)(1.0 + 2.0 * d.rgb_padding); cp.rgb_padding = (float)d.rgb_padding; cp.rgb = rgb_s;
            if (int e = time_begin(m, st)) return e;
            const bool use_tc = m->color_mode == 1 || (m->color_mode == 2 && m->tc_ok);
            if (use_tc) {
                UC_REQUIRE(m->tc_ok, "color_mlp=1 (tensor core) needs net_width_viewdirs/bottleneck <= 256 padded to 256 and deg_view == 4");
                ColorTcParams tp{};
                tp.n_rows = cp.n_rows; tp.S = S; tp.h1 = cp.h1; tp.viewdirs = cp.viewdirs;
                if (int e = m->dir_bias.ensure((size_t)n * 512 * sizeof(float))) return e;
                if (int e = launch_dir_bias(rp.viewdirs, m->wdir.as<float>(), cp.c0, cp.c1, m->dir_bias.as<float>(), n, st)) return e;
                tp.wblob = m->wblob.as<uint8_t>(); tp.dir_bias = m->dir_bias.as<float>(); tp.rt = cp.rt; tp.r0 = cp.r0;
                tp.rgb_scale = cp.rgb_scale; tp.rgb_padding = cp.rgb_padding; tp.rgb = cp.rgb; tp.debug_flags = m->tc_debug;
                tp.k0 = m->tc_k0; tp.k1 = m->tc_k1;
                if (int e = launch_color_mlp_tc(tp, st)) return e;
            } else {
                if (int e = launch_color_mlp_simt(cp, m->np, st)) return e;
            }
            if (int e = time_end(m, st, 3)) return e;
        }

        CompositeParams cq{};
        cq.n_rays = n; cq.S = S; cq.sdist = sdist; cq.sdist_stride = sdist_stride; cq.density = sp.density; cq.rgb = rgb_s;
        cq.rays = rp;
        cq.bg = (float)d.bg_intensity; cq.extras = 1; cq.weights = weights;
        if (nerf) {
            cq.o_rgb = o.rgb ? o.rgb + 3 * ray0 : nullptr;
            cq.o_depth = o.depth ? o.depth + ray0 : nullptr;
            cq.o_depth_raw = o.depth_raw ? o.depth_raw + ray0 : nullptr;
            cq.o_acc = o.acc ? o.acc + ray0 : nullptr;
            cq.o_mean = o.distance_mean ? o.distance_mean + ray0 : nullptr;
            cq.o_median = o.distance_median ? o.distance_median + ray0 : nullptr;
            cq.o_p5 = o.distance_percentile_5 ? o.distance_percentile_5 + ray0 : nullptr;
            cq.o_p95 = o.distance_percentile_95 ? o.distance_percentile_95 + ray0 : nullptr;
            cq.o_packed = o.packed ? o.packed + 12 * ray0 : nullptr;
            cq.n_peers = (int)m->n_peers;
            for (uint32_t k = 0; k < m->n_peers; ++k) cq.peer_packed[k] = m->peer_images[k];
            cq.peer_row0 = m->peer_row0 + ray0;
            cq.extras = (cq.o_mean || cq.o_median || cq.o_p5 || cq.o_p95 || cq.o_packed || cq.n_peers) ? 1 : 0;
            cq.use_affine = m->use_affine ? 1 : 0;
            std::memcpy(cq.affine, m->affine, sizeof(cq.affine));
        } else {
            cq.extras = 0;
        }
        if (int e = time_begin(m, st)) return e;
        if (int e = launch_composite(cq, st)) return e;
        if (int e = time_end(m, st, 4)) return e;

        t_prev = sdist; t_prev_stride = sdist_stride; w_prev = weights; n_prev = S;
    }
    return 0;
}

}  // namespace ucnerf

extern "C" int ucnerf_abi_version(void) { return UCNERF_ABI_VERSION; }
extern "C" const char* ucnerf_last_error(void) { return g_err.c_str(); }
extern "C" uint64_t ucnerf_launch_count(void) { return g_launch_count.load(); }

extern "C" int ucnerf_model_create(const ucnerf_model_desc* desc, ucnerf_model** out) {
    UC_REQUIRE(desc && out, "model_create: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("ucnerf_b200 requires a CUDA device (sm_100a); none is visible - there is no CPU fallback");
        return 4;
    }
    ucnerf_model* m = new ucnerf_model();
    if (int e = build_model(m, desc)) {
        ucnerf_model_destroy(m);
        return e;
    }
    *out = m;
    return 0;
}

extern "C" int ucnerf_model_refresh(ucnerf_model* m, const ucnerf_model_desc* desc, void* stream) {
    UC_REQUIRE(m && desc, "model_refresh: null argument");
    std::lock_guard<std::mutex> lk(m->mu);
    UC_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return build_model(m, desc);
}

extern "C" int ucnerf_model_destroy(ucnerf_model* m) {
    if (!m) return 0;
    for (auto& ls : m->lv) { ls.w1p.release(); ls.b1.release(); ls.w2.release(); ls.u.release(); ls.sdist.release(); ls.weights.release(); }
    for (DevBuf* b : {&m->w2t, &m->b2, &m->v0t, &m->c0, &m->v1t, &m->c1, &m->rt, &m->r0, &m->density, &m->h1, &m->rgb_s,
                      &m->stage_in, &m->stage_out, &m->cam_rays, &m->wblob, &m->wdir, &m->dir_bias})
        b->release();
    resolve_timing(m);
    for (auto& e : m->pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto e : m->ev_in) cudaEventDestroy(e);
    for (auto e : m->ev_done) cudaEventDestroy(e);
    if (m->copy_in) cudaStreamDestroy(m->copy_in);
    if (m->copy_out) cudaStreamDestroy(m->copy_out);
    delete m;
    return 0;
}

extern "C" int ucnerf_set_option(ucnerf_model* m, const char* key, int64_t value) {
    UC_REQUIRE(m && key, "set_option: null argument");
    const std::string k(key);
    if (k == "chunk_rays") { UC_REQUIRE(value >= 1, "chunk_rays must be >= 1"); m->chunk_rays = value; }
    else if (k == "ray_geom") m->ray_geom = value != 0;
    else if (k == "ray_tile_prop" || k == "ray_tile_nerf") {
        UC_REQUIRE(value == 0 || value == 4 || value == 8 || value == 16, "ray_tile_*: patch width 0 (rows), 4, 8 or 16");
        m->ray_tile_patch[k == "ray_tile_nerf" ? 1 : 0] = (int)value;
    }
    else if (k == "ray_tile_width") {
        UC_REQUIRE(value >= 0 && value % 4 == 0 && value <= (1 << 20), "ray_tile_width: image width, a multiple of 4 (0 = off)");
        m->ray_tile_width = value;
    }
    else if (k == "color_mlp") { UC_REQUIRE(value >= 0 && value <= 2, "color_mlp: 0 = fp32 SIMT, 1 = tensor core, 2 = auto"); m->color_mode = (int)value; }
    else if (k == "encode_runs") { UC_REQUIRE(value >= 0 && value <= 3, "encode_runs: bit 0 = proposal levels, bit 1 = NeRF level"); m->encode_runs = (int)value; }
    else if (k == "encode_mlp_mma") { UC_REQUIRE(value >= 0 && value <= 3, "encode_mlp_mma: bit 0 = proposal levels, bit 1 = NeRF level"); m->encode_mlp_mma = (int)value; }
    else if (k == "warp_rays_prop" || k == "warp_rays_nerf") {
        UC_REQUIRE(value == 32 || value == 16 || value == 8 || value == 4, "warp_rays_*: 32, 16, 8 or 4 rays per warp");
        m->warp_rays_log2[k == "warp_rays_nerf" ? 1 : 0] = value == 32 ? 5 : value == 16 ? 4 : value == 8 ? 3 : 2;
    }
    else if (k == "timing") {
        m->timing = value != 0;
        // events are created here, not inside the region the caller is about to time (cudaEventCreate occasionally takes
        // milliseconds); 512 pairs cover 18 frames of 4 chunks between two ucnerf_get_timing calls
        while (m->timing && m->pool.size() < 512) {
            cudaEvent_t a, b;
            UC_CUDA_OK(cudaEventCreate(&a));
            UC_CUDA_OK(cudaEventCreate(&b));
            m->pool.emplace_back(a, b);
        }
    }
    else if (k == "tc_debug") m->tc_debug = (uint32_t)value;  // profiling experiments (results invalid when != 0)
    else { set_error("set_option: unknown key " + k); return 1; }
    return 0;
}

extern "C" int ucnerf_set_rgb_affine(ucnerf_model* m, const float* affine12_host) {
    UC_REQUIRE(m, "set_rgb_affine: null model");
    std::lock_guard<std::mutex> lk(m->mu);
    m->use_affine = affine12_host != nullptr;
    if (affine12_host) std::memcpy(m->affine, affine12_host, sizeof(m->affine));
    return 0;
}

extern "C" int ucnerf_set_peer_targets(ucnerf_model* m, uint32_t n_peers, void* const* peer_images, uint64_t row0) {
    UC_REQUIRE(m, "set_peer_targets: null model");
    UC_REQUIRE(n_peers <= UCNERF_MAX_PEERS, "set_peer_targets: too many peers");
    UC_REQUIRE(n_peers == 0 || peer_images, "set_peer_targets: null peer list");
    std::lock_guard<std::mutex> lk(m->mu);
    for (uint32_t k = 0; k < n_peers; ++k) {
        UC_REQUIRE(peer_images[k], "set_peer_targets: null peer image");
        m->peer_images[k] = static_cast<float*>(peer_images[k]);
    }
    m->n_peers = n_peers;
    m->peer_row0 = row0;
    return 0;
}

extern "C" int ucnerf_get_timing(ucnerf_model* m, float* ms_out5, uint32_t* launches_out5, int reset) {
    UC_REQUIRE(m && ms_out5, "get_timing: null argument");
    std::lock_guard<std::mutex> lk(m->mu);
    if (int e = resolve_timing(m)) return e;
    for (int i = 0; i < 5; ++i) {
        ms_out5[i] = m->ms[i];
        if (launches_out5) launches_out5[i] = m->nlaunch[i];
    }
    if (reset) for (int i = 0; i < 5; ++i) { m->ms[i] = 0.f; m->nlaunch[i] = 0; }
    return 0;
}

namespace ucnerf {
// rays per internal chunk: with ray_tile_width set, a whole number of 8-row bands, so that every chunk starts on an image row
static uint64_t effective_chunk(const ucnerf_model* m, uint64_t tile_w) {
    const uint64_t c = (uint64_t)m->chunk_rays, band = 8u * tile_w;
    if (band == 0 || c < band) return c;
    return c / band * band;
}
}  // namespace ucnerf

static int render_rays_tiled(ucnerf_model* m, uint64_t n_rays, const ucnerf_rays* rays, double train_frac,
                             const ucnerf_outputs* out, void* stream, int64_t tile_w_or_option);

extern "C" int ucnerf_render_rays(ucnerf_model* m, uint64_t n_rays, const ucnerf_rays* rays, double train_frac,
                                  const ucnerf_outputs* out, void* stream) {
    return render_rays_tiled(m, n_rays, rays, train_frac, out, stream, -1);
}

// tile_w_or_option: image width of the row-major ray batch (0 = unknown), -1 = the model's "ray_tile_width" option
static int render_rays_tiled(ucnerf_model* m, uint64_t n_rays, const ucnerf_rays* rays, double train_frac,
                             const ucnerf_outputs* out, void* stream, int64_t tile_w_or_option) {
    UC_REQUIRE(m && rays && out, "render_rays: null argument");
    if (n_rays == 0) return 0;
    UC_REQUIRE(rays->origins && rays->directions && rays->viewdirs && rays->cam_dirs && rays->radii && rays->near &&
                   rays->far && rays->rand_vec,
               "render_rays: every ray array (incl. rand_vec) must be provided");
    std::lock_guard<std::mutex> lk(m->mu);
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t tile_w = (uint64_t)(tile_w_or_option < 0 ? m->ray_tile_width : tile_w_or_option);
    const uint64_t chunk = effective_chunk(m, tile_w);
    for (uint64_t r0 = 0; r0 < n_rays; r0 += chunk) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(chunk, n_rays - r0);
        if (int e = render_chunk(m, n, *rays, (size_t)r0, train_frac, *out, st, tile_w)) return e;
    }
    return 0;
}

namespace ucnerf {

// Device staging of the outputs a host caller asked for: fills `od` with device pointers into m->stage_out and
// returns the (host, device, floats) slots to copy back after the render.
struct OutSlot { float* host; size_t floats; float** dev_field; };
static int stage_outputs(ucnerf_model* m, size_t N, const ucnerf_outputs* oh, ucnerf_outputs& od, std::vector<OutSlot>& slots) {
    const int nl = m->num_levels;
    auto add = [&](float* host, size_t per_ray, float** field) { if (host) slots.push_back({host, N * per_ray, field}); };
    add(oh->rgb, 3, &od.rgb); add(oh->depth, 1, &od.depth); add(oh->depth_raw, 1, &od.depth_raw); add(oh->acc, 1, &od.acc);
    add(oh->distance_mean, 1, &od.distance_mean); add(oh->distance_median, 1, &od.distance_median);
    add(oh->distance_percentile_5, 1, &od.distance_percentile_5); add(oh->distance_percentile_95, 1, &od.distance_percentile_95);
    for (int l = 0; l < nl; ++l) {
        add(oh->sdist[l], (size_t)m->lv[l].S + 1, &od.sdist[l]);
        add(oh->weights[l], (size_t)m->lv[l].S, &od.weights[l]);
    }
    add(oh->sample_rgb, (size_t)m->lv[nl - 1].S * 3, &od.sample_rgb);
    add(oh->sample_density, (size_t)m->lv[nl - 1].S, &od.sample_density);
    add(oh->packed, 12, &od.packed);
    add(oh->sample_coord, (size_t)m->lv[nl - 1].S * 3, &od.sample_coord);
    size_t tot = 0;
    for (auto& s : slots) tot += (s.floats + 3) & ~size_t(3);
    if (int e = m->stage_out.ensure(std::max<size_t>(tot, 4) * sizeof(float))) return e;
    size_t o2 = 0;
    for (auto& s : slots) { *s.dev_field = m->stage_out.as<float>() + o2; o2 += (s.floats + 3) & ~size_t(3); }
    return 0;
}

static int copy_back_and_check(ucnerf_model* m, std::vector<OutSlot>& slots, cudaStream_t st) {
    (void)m;
    for (auto& s : slots)
        UC_CUDA_OK(cudaMemcpyAsync(s.host, *s.dev_field, s.floats * sizeof(float), cudaMemcpyDeviceToHost, st));
    UC_CUDA_OK(cudaStreamSynchronize(st));
    uint32_t wd[32];
    if (int e = color_tc_status(wd)) return e;
    if (wd[0] != 0) {
        set_error("color_mlp_tc: pipeline watchdog fired (tag " + std::to_string(wd[0]) + ", barrier " +
                  std::to_string(wd[3]) + ", step " + std::to_string(wd[6]) + ")");
        return 5;
    }
    return 0;
}

static int camera_const(const ucnerf_camera* cam, CameraConst& c) {
    UC_REQUIRE(cam->width >= 1 && cam->height >= 1, "camera: empty image");
    for (int i = 0; i < 9; ++i) c.pixtocam[i] = cam->pixtocam[i];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) c.rot[3 * i + j] = cam->camtoworld[4 * i + j];
        c.origin[i] = (float)cam->camtoworld[4 * i + 3];
        c.cam_dir[i] = (float)(-cam->camtoworld[4 * i + 2]);   // datasets.py:L446
    }
    c.near = cam->near; c.far = cam->far; c.width = cam->width; c.height = cam->height; c.rand_seed = cam->rand_seed;
    return 0;
}

// rays of image rows [row0, row0 + n_rows) into the model-owned buffer; rd receives the device pointers
static int camera_rays(ucnerf_model* m, const ucnerf_camera* cam, uint32_t row0, uint32_t n_rows, ucnerf_rays& rd, cudaStream_t st) {
    UC_REQUIRE((uint64_t)row0 + n_rows <= cam->height, "camera: row range outside the image");
    const size_t N = (size_t)n_rows * cam->width;
    if (int e = m->cam_rays.ensure(N * 18 * sizeof(float))) return e;
    float* b = m->cam_rays.as<float>();
    RayOutPtrs o{};
    o.origins = b; o.directions = b + 3 * N; o.viewdirs = b + 6 * N; o.cam_dirs = b + 9 * N; o.rand_vec = b + 12 * N;
    o.radii = b + 15 * N; o.near = b + 16 * N; o.far = b + 17 * N; o.imageplane = nullptr;
    CameraConst c;
    if (int e = camera_const(cam, c)) return e;
    if (int e = launch_generate_rays(c, row0, n_rows, o, st)) return e;
    rd.origins = o.origins; rd.directions = o.directions; rd.viewdirs = o.viewdirs; rd.cam_dirs = o.cam_dirs;
    rd.rand_vec = o.rand_vec; rd.radii = o.radii; rd.near = o.near; rd.far = o.far;
    return 0;
}

}  // namespace ucnerf

// Chunk pipeline of the host entries: the H2D copies of chunk c+1 (when `srcs` is given) and the D2H copies of chunk
// c-1 run on their own streams under chunk c's kernels.  The caller holds m->mu.
static int render_pipelined(ucnerf_model* m, size_t N, const ucnerf_rays& rd, const float* const* srcs, float* const* dsts,
                            const size_t* widths, double train_frac, const ucnerf_outputs& od, std::vector<OutSlot>& slots,
                            cudaStream_t st, uint64_t tile_w) {
    // ---- chunk pipeline: the copies of chunk c+1 (in) and c-1 (out) run on their own streams under chunk c's kernels ----
    if (!m->copy_in) {
        UC_CUDA_OK(cudaStreamCreateWithFlags(&m->copy_in, cudaStreamNonBlocking));
        UC_CUDA_OK(cudaStreamCreateWithFlags(&m->copy_out, cudaStreamNonBlocking));
    }
    const size_t chunk = (size_t)effective_chunk(m, tile_w);
    const size_t nchunks = (N + chunk - 1) / chunk;
    while (m->ev_in.size() < nchunks + 1) {
        cudaEvent_t a, b;
        UC_CUDA_OK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        UC_CUDA_OK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        m->ev_in.push_back(a);
        m->ev_done.push_back(b);
    }
    // the copy streams start after whatever the caller queued on `st` before this call
    UC_CUDA_OK(cudaEventRecord(m->ev_in[nchunks], st));
    UC_CUDA_OK(cudaStreamWaitEvent(m->copy_in, m->ev_in[nchunks], 0));
    UC_CUDA_OK(cudaStreamWaitEvent(m->copy_out, m->ev_in[nchunks], 0));
    for (size_t c = 0; c < nchunks; ++c) {
        const size_t r0 = c * chunk, n = std::min(chunk, N - r0);
        if (srcs)
            for (int i = 0; i < 8; ++i)
                UC_CUDA_OK(cudaMemcpyAsync(dsts[i] + r0 * widths[i], srcs[i] + r0 * widths[i], n * widths[i] * sizeof(float),
                                           cudaMemcpyHostToDevice, m->copy_in));
        UC_CUDA_OK(cudaEventRecord(m->ev_in[c], m->copy_in));
    }
    for (size_t c = 0; c < nchunks; ++c) {
        const size_t r0 = c * chunk, n = std::min(chunk, N - r0);
        UC_CUDA_OK(cudaStreamWaitEvent(st, m->ev_in[c], 0));
        if (int e = render_chunk(m, (uint32_t)n, rd, r0, train_frac, od, st, tile_w)) return e;
        UC_CUDA_OK(cudaEventRecord(m->ev_done[c], st));
        UC_CUDA_OK(cudaStreamWaitEvent(m->copy_out, m->ev_done[c], 0));
        for (auto& sl : slots) {
            const size_t per = sl.floats / N;
            UC_CUDA_OK(cudaMemcpyAsync(sl.host + r0 * per, *sl.dev_field + r0 * per, n * per * sizeof(float),
                                       cudaMemcpyDeviceToHost, m->copy_out));
        }
    }
    UC_CUDA_OK(cudaStreamSynchronize(m->copy_out));
    UC_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int ucnerf_generate_rays(const ucnerf_camera* cam, uint32_t row0, uint32_t n_rows, const ucnerf_ray_buffers* out,
                                    void* stream) {
    UC_REQUIRE(cam && out, "generate_rays: null argument");
    UC_REQUIRE(out->directions && out->viewdirs && out->radii, "generate_rays: directions, viewdirs and radii are required");
    UC_REQUIRE((uint64_t)row0 + n_rows <= cam->height, "generate_rays: row range outside the image");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("ucnerf_b200 requires a CUDA device (sm_100a); none is visible - there is no CPU fallback");
        return 4;
    }
    CameraConst c;
    if (int e = camera_const(cam, c)) return e;
    RayOutPtrs o{out->origins, out->directions, out->viewdirs, out->cam_dirs, out->radii, out->near, out->far,
                 out->imageplane, out->rand_vec};
    return launch_generate_rays(c, row0, n_rows, o, (cudaStream_t)stream);
}

extern "C" int ucnerf_render_camera(ucnerf_model* m, const ucnerf_camera* cam, uint32_t row0, uint32_t n_rows,
                                    double train_frac, const ucnerf_outputs* out, void* stream) {
    UC_REQUIRE(m && cam && out, "render_camera: null argument");
    if (n_rows == 0) return 0;
    ucnerf_rays rd{};
    {
        std::lock_guard<std::mutex> lk(m->mu);
        if (int e = camera_rays(m, cam, row0, n_rows, rd, (cudaStream_t)stream)) return e;
    }
    return render_rays_tiled(m, (uint64_t)n_rows * cam->width, &rd, train_frac, out, stream, cam->width % 4 == 0 ? (int64_t)cam->width : 0);
}

extern "C" int ucnerf_render_camera_host(ucnerf_model* m, const ucnerf_camera* cam, uint32_t row0, uint32_t n_rows,
                                         double train_frac, const ucnerf_outputs* oh, void* stream) {
    UC_REQUIRE(m && cam && oh, "render_camera_host: null argument");
    if (n_rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lk(m->mu);
    const size_t N = (size_t)n_rows * cam->width;
    ucnerf_outputs od{};
    std::vector<OutSlot> slots;
    if (int e = stage_outputs(m, N, oh, od, slots)) return e;
    ucnerf_rays rd{};
    if (int e = camera_rays(m, cam, row0, n_rows, rd, st)) return e;      // one kernel on the render stream
    if (int e = render_pipelined(m, N, rd, nullptr, nullptr, nullptr, train_frac, od, slots, st, cam->width % 4 == 0 ? cam->width : 0)) return e;
    slots.clear();
    return copy_back_and_check(m, slots, st);
}

extern "C" int ucnerf_render_rays_host(ucnerf_model* m, uint64_t n_rays, const ucnerf_rays* rh, double train_frac,
                                       const ucnerf_outputs* oh, void* stream) {
    UC_REQUIRE(m && rh && oh, "render_rays_host: null argument");
    if (n_rays == 0) return 0;
    UC_REQUIRE(rh->origins && rh->directions && rh->viewdirs && rh->cam_dirs && rh->radii && rh->near && rh->far &&
                   rh->rand_vec,
               "render_rays_host: every ray array (incl. rand_vec) must be provided");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = (size_t)n_rays;
    std::lock_guard<std::mutex> lk(m->mu);
    // ---- staging: 5 x [N,3] + 3 x [N] floats in, the requested outputs out ----
    if (int e = m->stage_in.ensure(N * 18 * sizeof(float))) return e;
    float* base = m->stage_in.as<float>();
    ucnerf_rays rd{};
    const float* srcs[8] = {rh->origins, rh->directions, rh->viewdirs, rh->cam_dirs, rh->rand_vec, rh->radii, rh->near, rh->far};
    const size_t widths[8] = {3, 3, 3, 3, 3, 1, 1, 1};
    float* dsts[8];
    size_t off = 0;
    for (int i = 0; i < 8; ++i) { dsts[i] = base + off; off += N * widths[i]; }
    rd.origins = dsts[0]; rd.directions = dsts[1]; rd.viewdirs = dsts[2]; rd.cam_dirs = dsts[3]; rd.rand_vec = dsts[4];
    rd.radii = dsts[5]; rd.near = dsts[6]; rd.far = dsts[7];
    ucnerf_outputs od{};
    std::vector<OutSlot> slots;
    if (int e = stage_outputs(m, N, oh, od, slots)) return e;
    if (int e = render_pipelined(m, N, rd, srcs, dsts, widths, train_frac, od, slots, st, (uint64_t)m->ray_tile_width)) return e;
    slots.clear();
    return copy_back_and_check(m, slots, st);
}

// Watchdog record of the tensor-core colour MLP (synchronises the device): out16[0] != 0 means a pipeline wait
// timed out; [1..6] = block, thread, barrier id, parity, tile iteration, step.
extern "C" int ucnerf_debug_tc_status(uint32_t* out16) {
    UC_REQUIRE(out16, "debug_tc_status: null");
    return color_tc_status(out16);
}

// Debug entry: one resampling level on device arrays (the kernel of the render path); scratch_out (device, may be NULL)
// receives the shared-memory scratch of ray `dbg_ray`: [tp (n+1) | pp (n) | T (3n+1) | W (3n+1) | CW (3n+2) | C (S)].
extern "C" int ucnerf_debug_resample(uint32_t n_rays, int n_prev, const float* t_prev, const float* w_prev, int dilate,
                                     float dilation, float anneal, float padding, int S, float* out_sdist,
                                     float* scratch_out, uint32_t dbg_ray, void* stream) {
    std::vector<float> u(S);
    deterministic_u(S, u.data());
    float* du = nullptr;
    UC_CUDA_OK(cudaMalloc(&du, S * sizeof(float)));
    UC_CUDA_OK(cudaMemcpy(du, u.data(), S * sizeof(float), cudaMemcpyHostToDevice));
    ResampleParams rs{};
    rs.n_rays = n_rays; rs.n_prev = n_prev; rs.t_prev = t_prev; rs.t_prev_stride = (uint32_t)(n_prev + 1); rs.w_prev = w_prev;
    rs.dilate = dilate; rs.dilation = dilation; rs.anneal = anneal; rs.padding = padding; rs.S = S; rs.u = du;
    rs.out_sdist = out_sdist; rs.dbg_scratch = scratch_out; rs.dbg_ray = dbg_ray;
    int e = launch_resample(rs, (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(du);
    return e;
}

// ---- small host-only helpers exported for the CPU test-suite (no GPU needed) ---------------------
extern "C" int ucnerf_debug_u_grid(int S, float* out_host) {
    UC_REQUIRE(S >= 1 && out_host, "debug_u_grid: bad argument");
    deterministic_u(S, out_host);
    return 0;
}
extern "C" int ucnerf_debug_cone_table(float* out30) {
    UC_REQUIRE(out30, "debug_cone_table: null");
    ConeTable ct;
    make_cone_table(ct);
    std::memcpy(out30, &ct, sizeof(float) * 30);
    return 0;
}

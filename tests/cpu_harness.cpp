// TEST INFRASTRUCTURE: instantiates the host+device algorithm templates of
// ucnerf_b200/csrc/ray_algos.cuh with SerialExec so the not-gpu test-suite can check the device
// logic (tie semantics of the resampler, cone sampling op order, hash indexing, compositing)
// against the oracle in a container without a GPU.  Never linked into libucnerf_b200.so.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC -I/usr/local/cuda/include tests/cpu_harness.cpp
#include <cstring>
#include <vector>
#include "../ucnerf_b200/csrc/ray_algos.cuh"
#include "../ucnerf_b200/csrc/pooled_algos.cuh"
#include "../ucnerf_b200/csrc/train_algos.cuh"

using namespace ucnerf;

extern "C" {

void h_resample(int n_rays, int n_prev, const float* t_prev, const float* w_prev, int dilate, float dilation,
                float anneal, float padding, int S, const float* u, float* out) {
    std::vector<float> scratch(ResampleScratch::floats(n_prev, S));
    SerialExec ex;
    for (int r = 0; r < n_rays; ++r) {
        ResampleScratch sc;
        sc.carve(scratch.data(), n_prev, S);
        resample_ray(ex, n_prev, t_prev ? t_prev + (size_t)r * (n_prev + 1) : nullptr,
                     w_prev ? w_prev + (size_t)r * n_prev : nullptr, dilate != 0, dilation, anneal, padding, S, u, sc,
                     out + (size_t)r * (S + 1));
    }
}

// training variant: jitter [N, jitter_cols] (already scaled by max_jitter) added to the base u grid, stepfun.py:L206-212
void h_resample_jitter(int n_rays, int n_prev, const float* t_prev, const float* w_prev, int dilate, float dilation,
                       float anneal, float padding, int S, const float* u, const float* jitter, int jitter_cols, float* out) {
    std::vector<float> scratch(ResampleScratch::floats(n_prev, S));
    SerialExec ex;
    for (int r = 0; r < n_rays; ++r) {
        ResampleScratch sc;
        sc.carve(scratch.data(), n_prev, S);
        resample_ray(ex, n_prev, t_prev ? t_prev + (size_t)r * (n_prev + 1) : nullptr,
                     w_prev ? w_prev + (size_t)r * n_prev : nullptr, dilate != 0, dilation, anneal, padding, S, u, sc,
                     out + (size_t)r * (S + 1), jitter ? jitter + (size_t)r * jitter_cols : nullptr, jitter_cols > 1 ? 1 : 0);
    }
}

// out_ray: [N, 10] = rgb[3], depth, depth_raw, acc, mean, median, p5, p95
void h_composite(int n_rays, int S, const float* sdist, const float* density, const float* rgb, const float* dirs,
                 const float* near, const float* far, float bg, int extras, float* out_w, float* out_ray) {
    std::vector<float> scratch(CompositeScratch::floats(S));
    SerialExec ex;
    for (int r = 0; r < n_rays; ++r) {
        CompositeScratch sc;
        sc.carve(scratch.data(), S);
        RayOutputs ro;
        composite_ray(ex, S, sdist + (size_t)r * (S + 1), density + (size_t)r * S,
                      rgb ? rgb + (size_t)r * S * 3 : nullptr, dirs + 3 * (size_t)r, near[r], far[r], bg, extras != 0, sc,
                      out_w + (size_t)r * S, ro);
        float* o = out_ray + 10 * (size_t)r;
        o[0] = ro.rgb[0]; o[1] = ro.rgb[1]; o[2] = ro.rgb[2]; o[3] = ro.depth; o[4] = ro.depth_raw; o[5] = ro.acc;
        o[6] = ro.dist_mean; o[7] = ro.dist_median; o[8] = ro.dist_p5; o[9] = ro.dist_p95;
    }
}

// out_g [N,S,6,3] unit-cube coordinates, out_sigma [N,S,6]
void h_cone_points(int n_rays, int S, const float* origins, const float* directions, const float* cam_dirs,
                   const float* rand_vec, const float* radii, const float* near, const float* far, const float* sdist,
                   float std_scale, float* out_g, float* out_sigma) {
    ConeTable ct;
    make_cone_table(ct);
    for (int r = 0; r < n_rays; ++r) {
        RayGeom rg;
        make_ray_geom(rg, origins + 3 * r, directions + 3 * r, cam_dirs + 3 * r, rand_vec + 3 * r, radii[r], near[r], far[r]);
        for (int s = 0; s < S; ++s) {
            const float s0 = sdist[(size_t)r * (S + 1) + s], s1 = sdist[(size_t)r * (S + 1) + s + 1];
            const float t0 = fa(fm(s0, rg.far), fm(fs(1.f, s0), rg.near));
            const float t1 = fa(fm(s1, rg.far), fm(fs(1.f, s1), rg.near));
            const ConeInterval ci = make_cone_interval(t0, t1);
            for (int j = 0; j < 6; ++j) {
                float g[3], sg;
                cone_point(rg, ci, ct, j, s & 1, std_scale, g, sg);
                float* og = out_g + (((size_t)r * S + s) * 6 + j) * 3;
                og[0] = g[0]; og[1] = g[1]; og[2] = g[2];
                out_sigma[((size_t)r * S + s) * 6 + j] = sg;
            }
        }
    }
}

// out_c [N,S,3]: `coord` of every interval (mean of the six contracted multisample means / 2)
void h_sample_coord(int n_rays, int S, const float* origins, const float* directions, const float* cam_dirs,
                    const float* rand_vec, const float* radii, const float* near, const float* far, const float* sdist,
                    float std_scale, float* out_c) {
    ConeTable ct;
    make_cone_table(ct);
    for (int r = 0; r < n_rays; ++r) {
        RayGeom rg;
        make_ray_geom(rg, origins + 3 * r, directions + 3 * r, cam_dirs + 3 * r, rand_vec + 3 * r, radii[r], near[r], far[r]);
        for (int s = 0; s < S; ++s) {
            const float s0 = sdist[(size_t)r * (S + 1) + s], s1 = sdist[(size_t)r * (S + 1) + s + 1];
            const float t0 = fa(fm(s0, rg.far), fm(fs(1.f, s0), rg.near));
            const float t1 = fa(fm(s1, rg.far), fm(fs(1.f, s1), rg.near));
            float c[3];
            interval_coord(rg, t0, t1, ct, s & 1, std_scale, c);
            for (int i = 0; i < 3; ++i) out_c[((size_t)r * S + s) * 3 + i] = c[i];
        }
    }
}

// camera -> rays for rows [row0, row0 + n_rows): out_dir / out_view [n,3], out_plane [n,2], out_radius [n]; normals [n,4]
void h_pixel_rays(const double* pixtocam, const double* camtoworld, int width, int height, int row0, int n_rows,
                  uint64_t seed, float* out_dir, float* out_view, float* out_plane, float* out_radius, float* normals) {
    CameraConst c;
    for (int i = 0; i < 9; ++i) c.pixtocam[i] = pixtocam[i];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c.rot[3 * i + j] = camtoworld[4 * i + j];
    c.width = (uint32_t)width; c.height = (uint32_t)height; c.rand_seed = seed;
    size_t i = 0;
    for (int y = row0; y < row0 + n_rows; ++y)
        for (int x = 0; x < width; ++x, ++i) {
            PixelRay pr;
            pixel_to_ray(c, x, y, pr);
            for (int k = 0; k < 3; ++k) { out_dir[3 * i + k] = pr.dir[k]; out_view[3 * i + k] = pr.view[k]; }
            out_plane[2 * i] = pr.plane[0]; out_plane[2 * i + 1] = pr.plane[1];
            out_radius[i] = pr.radius;
            if (normals) {
                float n4[4];
                normal4(seed, (uint64_t)y * width + x, n4);
                for (int k = 0; k < 4; ++k) normals[4 * i + k] = n4[k];
            }
        }
}

// fused-path hash lookup for B points: lv = L x {offset, hashmap_size, stride1, hashed, pow2_mask, scale_bits}
void h_grid_features(int B, int L, const uint32_t* lvdesc, const float* table, const float* g, float* out) {
    for (int b = 0; b < B; ++b) {
        const float gg[3] = {g[3 * b], g[3 * b + 1], g[3 * b + 2]};
        for (int l = 0; l < L; ++l) {
            GridLevel lv;
            lv.offset = lvdesc[6 * l]; lv.hashmap_size = lvdesc[6 * l + 1]; lv.stride1 = lvdesc[6 * l + 2];
            lv.stride2 = lv.stride1 * lv.stride1;
            lv.hashed = lvdesc[6 * l + 3]; lv.pow2_mask = lvdesc[6 * l + 4];
            lv.mod_mode = !lv.hashed ? 0u : (lv.pow2_mask ? 1u : 2u);
            std::memcpy(&lv.scale, &lvdesc[6 * l + 5], 4);
            const CellCoords c = cell_of(lv, gg);
            float r[4] = {0, 0, 0, 0};
            for (int k = 0; k < 8; ++k) {
                const uint32_t ii = level_index(lv, c.ix + (k & 1), c.iy + ((k >> 1) & 1), c.iz + ((k >> 2) & 1));
                const float w = ((k & 1) ? c.fx : 1 - c.fx) * ((k & 2) ? c.fy : 1 - c.fy) * ((k & 4) ? c.fz : 1 - c.fz);
                const float* v = table + 4 * ((size_t)lv.offset + ii);
                for (int ch = 0; ch < 4; ++ch) r[ch] = fmaf(w, v[ch], r[ch]);
            }
            std::memcpy(out + ((size_t)b * L + l) * 4, r, 16);
        }
    }
}
}

// pooled hash-grid encode (pooled_algos.cuh) for B intervals of M points, levels built by the product's own
// make_grid_level from offsets / grid_sizes: features [B, L*4], coord [B,3] (optional)
extern "C" void h_pooled_forward(int B, int M, int contract, int L, const int32_t* offsets, const int32_t* grid_sizes, float S,
                                 uint32_t H, const float* table, const float* means, const float* stds, float* features,
                                 float* coord) {
    struct HostLoad {
        const float* t;
        float4 operator()(size_t e) const { return make_float4(t[4 * e], t[4 * e + 1], t[4 * e + 2], t[4 * e + 3]); }
    };
    for (int l = 0; l < L; ++l) {
        GridLevel lv;
        make_grid_level(lv, l, offsets[l], offsets[l + 1], S, H, grid_sizes[l]);
        const float g2 = (float)(int32_t)((int64_t)grid_sizes[l] * grid_sizes[l]);
        for (int b = 0; b < B; ++b) {
            float F[4];
            pooled_level_forward(lv, g2, means + (size_t)b * M * 3, stds + (size_t)b * M, M, contract != 0, HostLoad{table}, F);
            std::memcpy(features + ((size_t)b * L + l) * 4, F, 16);
        }
    }
    if (coord)
        for (int b = 0; b < B; ++b) {
            float c[3];
            pooled_coord(means + (size_t)b * M * 3, stds + (size_t)b * M, M, contract != 0, c);
            std::memcpy(coord + 3 * (size_t)b, c, 12);
        }
}

// grad_table [sum T, 4] accumulated in fp64 (the CUDA instantiation uses fp32 red.add in nondeterministic order)
extern "C" void h_pooled_backward(int B, int M, int flags, int L, const int32_t* offsets, const int32_t* grid_sizes, float S,
                                  uint32_t H, const float* grad_features, const float* means, const float* stds,
                                  double* grad_table) {
    struct HostAdd {
        double* t;
        void operator()(size_t e, float a, float b, float c, float d) const {
            t[4 * e] += a; t[4 * e + 1] += b; t[4 * e + 2] += c; t[4 * e + 3] += d;
        }
    };
    for (int l = 0; l < L; ++l) {
        GridLevel lv;
        make_grid_level(lv, l, offsets[l], offsets[l + 1], S, H, grid_sizes[l]);
        const float g2 = (float)(int32_t)((int64_t)grid_sizes[l] * grid_sizes[l]);
        if ((flags & 4) && B % 4 == 0) {      // ray-run variant: 4 consecutive intervals per (serial) thread
            for (int b = 0; b < B; b += 4)
                pooled_level_backward_ray_runs(lv, g2, means + (size_t)b * M * 3, stds + (size_t)b * M, M, 4, (flags & 1) != 0,
                                               grad_features + (size_t)b * L * 4, L * 4, 4 * l, HostAdd{grad_table});
            continue;
        }
        for (int b = 0; b < B; ++b) {
            float dF[4];
            std::memcpy(dF, grad_features + ((size_t)b * L + l) * 4, 16);
            const bool contract = (flags & 1) != 0;
            if (flags & 2) pooled_level_backward_runs(lv, g2, means + (size_t)b * M * 3, stds + (size_t)b * M, M, contract, dF, HostAdd{grad_table});
            else pooled_level_backward(lv, g2, means + (size_t)b * M * 3, stds + (size_t)b * M, M, contract, dF, HostAdd{grad_table});
        }
    }
}

// differentiable compositing (train_algos.cuh), the loops the CUDA kernels run one thread per ray
extern "C" void h_composite_train_forward(int N, int S, const float* tdist, const float* density, const float* rgbs,
                                          const float* dirs, float bg, float* weights, float* rgb, float* acc) {
    for (int r = 0; r < N; ++r) {
        const float dn = norm3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
        float c[3], a;
        composite_train_forward_ray(S, tdist + (size_t)r * (S + 1), density + (size_t)r * S,
                                    rgbs ? rgbs + (size_t)r * S * 3 : nullptr, dn, bg, weights + (size_t)r * S, c, a);
        rgb[3 * r] = c[0]; rgb[3 * r + 1] = c[1]; rgb[3 * r + 2] = c[2];
        acc[r] = a;
    }
}
extern "C" void h_composite_train_backward(int N, int S, const float* tdist, const float* density, const float* rgbs,
                                           const float* dirs, float bg, const float* weights, const float* acc,
                                           const float* g_w, const float* g_rgb, const float* g_acc, float* d_density,
                                           float* d_rgbs) {
    for (int r = 0; r < N; ++r) {
        const float dn = norm3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
        composite_train_backward_ray(S, tdist + (size_t)r * (S + 1), density + (size_t)r * S,
                                     rgbs ? rgbs + (size_t)r * S * 3 : nullptr, dn, bg, weights + (size_t)r * S, acc[r],
                                     g_w ? g_w + (size_t)r * S : nullptr, g_rgb ? g_rgb + 3 * r : nullptr,
                                     g_acc ? g_acc + r : nullptr, d_density + (size_t)r * S,
                                     (rgbs && d_rgbs) ? d_rgbs + (size_t)r * S * 3 : nullptr);
    }
}

// render.cast_rays (train_algos.cuh::cast_interval): means [N,S,6,3], stds [N,S,6], ts [N,S,6]; rot01 / flip01 NULL = rand=False
extern "C" void h_cast_rays(int N, int S, const float* tdist, const float* origins, const float* directions,
                            const float* cam_dirs, const float* radii, const float* rand_vec, const float* rot01,
                            const float* flip01, float std_scale, float* means, float* stds, float* ts) {
    ConeTable ct;
    make_cone_table(ct);
    for (int r = 0; r < N; ++r) {
        RayGeom rg;
        make_ray_geom(rg, origins + 3 * r, directions + 3 * r, cam_dirs + 3 * r, rand_vec + 3 * r, radii[r], 0.f, 1.f);
        for (int s = 0; s < S; ++s) {
            const size_t q = (size_t)r * S + s;
            cast_interval(rg, tdist[(size_t)r * (S + 1) + s], tdist[(size_t)r * (S + 1) + s + 1], ct, s, rot01 != nullptr,
                          rot01 ? rot01[q] : 0.f, flip01 ? flip01[q] : 1.f, std_scale, means + q * 18, stds + q * 6, ts + q * 6);
        }
    }
}

// debug variant: also returns the scratch arrays of the last ray processed (T, W(=probabilities), CW, C)
extern "C" void h_resample_debug(int n_prev, const float* t_prev, const float* w_prev, int dilate, float dilation,
                                 float anneal, float padding, int S, const float* u, float* out, float* T, float* W,
                                 float* CW, float* C) {
    std::vector<float> scratch(ResampleScratch::floats(n_prev, S));
    SerialExec ex;
    ResampleScratch sc;
    sc.carve(scratch.data(), n_prev, S);
    resample_ray(ex, n_prev, t_prev, w_prev, dilate != 0, dilation, anneal, padding, S, u, sc, out);
    std::memcpy(T, sc.T, sizeof(float) * (3 * n_prev + 1));
    std::memcpy(W, sc.W, sizeof(float) * (3 * n_prev + 1));
    std::memcpy(CW, sc.CW, sizeof(float) * (3 * n_prev + 2));
    std::memcpy(C, sc.C, sizeof(float) * S);
}

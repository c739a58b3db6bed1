// Alpha compositing of one ray for TRAINING, forward and backward, as host+device code (CUDA instantiation:
// composite_train.cu, one thread per ray; the very same functions run serially in tests/cpu_harness.cpp).
//
// Reference (under /root/reference/nerf/internal/): render.py:L155-174 compute_alpha_weights (opaque_background = False),
// L177-216 volumetric_rendering (acc, bg_w, rgb) - the part of a level the losses differentiate:
//
//   dd_k  = density_k * (t_{k+1} - t_k) * |d|        T_k = exp(-sum_{j<k} dd_j)        w_k = (1 - exp(-dd_k)) * T_k
//   acc   = sum_k w_k                                  rgb = sum_k w_k c_k + max(1 - acc, 0) * bg
//
// backward, with G_w [S], G_rgb [3], G_acc the incoming gradients (any of them may be absent):
//   g_k          = G_w[k] + G_rgb . c_k + G_acc - [1 - acc >= 0] * bg * (G_rgb[0] + G_rgb[1] + G_rgb[2])
//   d c_k        = w_k * G_rgb
//   d density_k  = (t_{k+1} - t_k) |d| * ( g_k * T_k * exp(-dd_k)  -  sum_{j>k} g_j w_j )
// which is what autograd derives through alpha (first term) and through the exclusive cumsum in T (second term).
// tdist carries no gradient (sdist is detached, models.py:L203-204).  Running sums are kept in fp64 and rounded where
// torch holds an fp32 tensor.
#pragma once
#include "ray_algos.cuh"

namespace ucnerf {

// t [S+1] metric fenceposts, density [S], rgb [S*3] or NULL (proposal levels render no colour), dnorm = |d|.
// w_out [S]; rgb_out [3] (bg only when rgb == NULL); acc_out.
UC_HD void composite_train_forward_ray(int S, const float* t, const float* density, const float* rgb, float dnorm, float bg,
                                       float* w_out, float (&rgb_out)[3], float& acc_out) {
    double cum = 0.0, acc = 0.0, r0 = 0.0, r1 = 0.0, r2 = 0.0;
    for (int k = 0; k < S; ++k) {
        const float dd = fm(density[k], fm(fs(t[k + 1], t[k]), dnorm));
        const float T = expf(-(float)cum);
        const float w = fm(fs(1.f, expf(-dd)), T);
        w_out[k] = w;
        acc += (double)w;
        if (rgb) {
            r0 += (double)fm(w, rgb[3 * k]);
            r1 += (double)fm(w, rgb[3 * k + 1]);
            r2 += (double)fm(w, rgb[3 * k + 2]);
        }
        cum += (double)dd;
    }
    const float accf = (float)acc;
    const float bgw = fmaxf(fs(1.f, accf), 0.f);        // render.py:L204 clamp_min(0)
    rgb_out[0] = fa((float)r0, fm(bgw, bg));
    rgb_out[1] = fa((float)r1, fm(bgw, bg));
    rgb_out[2] = fa((float)r2, fm(bgw, bg));
    acc_out = accf;
}

// w [S] = the forward's weights, acc its sum.  g_w [S] / g_rgb [3] / g_acc: incoming gradients, NULL = none.
// d_density [S]; d_rgb [S*3] or NULL.
UC_HD void composite_train_backward_ray(int S, const float* t, const float* density, const float* rgb, float dnorm, float bg,
                                        const float* w, float acc, const float* g_w, const float* g_rgb, const float* g_acc,
                                        float* d_density, float* d_rgb) {
    const float G0 = g_rgb ? g_rgb[0] : 0.f, G1 = g_rgb ? g_rgb[1] : 0.f, G2 = g_rgb ? g_rgb[2] : 0.f;
    // d rgb_i / d acc = -bg where clamp_min passes the gradient (1 - acc >= 0)
    const float gshared = (g_acc ? g_acc[0] : 0.f) - ((fs(1.f, acc) >= 0.f) ? bg * (G0 + G1 + G2) : 0.f);
    double cum = 0.0;
    for (int k = 0; k < S; ++k) cum += (double)fm(density[k], fm(fs(t[k + 1], t[k]), dnorm));
    double suffix = 0.0;                                  // sum_{j>k} g_j w_j
    for (int k = S - 1; k >= 0; --k) {
        const float delta = fm(fs(t[k + 1], t[k]), dnorm);
        const float dd = fm(density[k], delta);
        cum -= (double)dd;                                // exclusive prefix of dd at k
        float g = gshared + (g_w ? g_w[k] : 0.f);
        if (rgb) g += G0 * rgb[3 * k] + G1 * rgb[3 * k + 1] + G2 * rgb[3 * k + 2];
        const float T = expf(-(float)cum);
        const double d_dd = (double)g * (double)(T * expf(-dd)) - suffix;
        d_density[k] = (float)(d_dd * (double)delta);
        if (d_rgb) {
            d_rgb[3 * k] = w[k] * G0; d_rgb[3 * k + 1] = w[k] * G1; d_rgb[3 * k + 2] = w[k] * G2;
        }
        suffix += (double)g * (double)w[k];
    }
}

}  // namespace ucnerf

namespace ucnerf {

// render.cast_rays for one multisample point of one interval, BEFORE any warp / contraction: world-space mean, std and
// the point's distance t (render.py:L108-148).  `cosv`, `sinv`: cos / sin of the point's angle in the cone cross
// section; the lateral offsets and the std are divided by sqrt(2) with an IEEE division exactly as the reference does
// (the eval kernel's cone_point multiplies by the reciprocal instead and fuses the contraction).
UC_HD void cast_point(const RayGeom& rg, const ConeInterval& ci, float tcoef, float cosv, float sinv, float std_scale,
                      float (&mean)[3], float& std, float& t_out) {
    const float t = fa(ci.t0, fm(ci.tdA, fa(ci.B, fm(tcoef, ci.Cq))));
    const float rt = fm(rg.radius, t);
    const float px = fd(fm(rt, cosv), 1.41421356237f);
    const float py = fd(fm(rt, sinv), 1.41421356237f);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        mean[i] = fa(fa(fa(fm(px, rg.e1[i]), fm(py, rg.e2[i])), fm(t, rg.d[i])), rg.o[i]);
    std = fd(fm(fm(std_scale, rg.radius), t), 1.41421356237f);
    t_out = t;
}

// The six points of interval s of one ray.  rand = false: the deterministic pattern (30-degree rotation and flip of every
// other interval, render.py:L125-131, cos / sin from the torch-pinned table).  rand = true (training, L119-124):
// rot01, flip01 = the ray's uniform draws for this interval: angle_j = pi/3 k_j + 2 pi rot01, kept if flip01 > 0.5, else
// 5 pi / 3 - angle_j.  means [6*3], stds [6], ts [6].
UC_HD void cast_interval(const RayGeom& rg, float t0, float t1, const ConeTable& ct, int s, bool rand, float rot01,
                         float flip01, float std_scale, float* means, float* stds, float* ts) {
    const ConeInterval ci = make_cone_interval(t0, t1);
    const float kk[6] = {0.f, 2.f, 4.f, 3.f, 5.f, 1.f};
    for (int j = 0; j < 6; ++j) {
        float cv, sv;
        if (rand) {
            float deg = fa(fm(1.0471975511965976f, kk[j]), fm(6.283185307179586f, rot01));
            if (!(flip01 > 0.5f)) deg = fs(5.235987755982989f, deg);
            cv = cosf(deg);
            sv = sinf(deg);
        } else {
            cv = ct.cosv[s & 1][j];
            sv = ct.sinv[s & 1][j];
        }
        float m[3], sd, t;
        cast_point(rg, ci, ct.tcoef[j], cv, sv, std_scale, m, sd, t);
        means[3 * j] = m[0]; means[3 * j + 1] = m[1]; means[3 * j + 2] = m[2];
        stds[j] = sd;
        ts[j] = t;
    }
}

}  // namespace ucnerf

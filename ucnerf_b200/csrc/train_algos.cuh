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

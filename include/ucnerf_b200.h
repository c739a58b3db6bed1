/*
 * ucnerf_b200 - C ABI of the B200-native UC-NeRF forward-render hot path.
 *
 * Plain C: pointers, sizes and PODs only (no torch / ATen types).  All `const void*` / `void*`
 * data pointers are DEVICE pointers unless the function name ends in `_host`.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Every entry point returns 0 on
 * success and a non-zero code on failure; ucnerf_last_error() then returns a message (thread
 * local).  Nothing is retained across calls except what lives inside a ucnerf_model handle.
 *
 * Reference interface each entry point replaces (paths under /root/reference/nerf/):
 *   ucnerf_grid_encode_forward   <- grid_encode_forward   gridencoder/src/gridencoder.h:L12, gridencoder.cu:L448-471
 *   ucnerf_grid_encode_backward  <- grid_encode_backward  gridencoder/src/gridencoder.h:L13, gridencoder.cu:L473-503
 *   ucnerf_grad_total_variation  <- grad_total_variation  gridencoder/src/gridencoder.h:L15, gridencoder.cu:L639-645
 *   ucnerf_model_* / ucnerf_render_rays[_host]
 *                                <- Model.forward eval path internal/models.py:L97-324 (level loop,
 *                                   stepfun.py resampling, render.py cast_rays / compute_alpha_weights /
 *                                   volumetric_rendering, MLP.forward models.py:L514-685), called per chunk
 *                                   from render_image models.py:L907-1007.
 *   ucnerf_generate_rays / ucnerf_render_camera[_host]
 *                                <- camera_utils.pixels_to_rays internal/camera_utils.py:L448-557 (perspective, no
 *                                   distortion / NDC), cast_pinhole_rays L611-632 and Dataset._make_ray_batch
 *                                   internal/datasets.py:L386-476 (near / far broadcast, cam_dirs, float32 cast): the
 *                                   numpy ray generation of the eval loader, SURVEY.md section 8f N3.
 * The Python-side binding a reference maintainer would add is shown in INTEGRATION.md.
 */
#ifndef UCNERF_B200_H
#define UCNERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UCNERF_ABI_VERSION 1

/* embeddings / outputs / grad dtype codes (reference: AT_DISPATCH_FLOATING_TYPES_AND_HALF) */
#define UCNERF_F32 0
#define UCNERF_F16 1
#define UCNERF_F64 2

#define UCNERF_MAX_GRID_LEVELS 16
#define UCNERF_MAX_PROP_LEVELS 4

int ucnerf_abi_version(void);
const char* ucnerf_last_error(void);

/* ---- gridencoder drop-in (gridencoder.h:L12-15). Argument meaning identical to the reference:
 * inputs [B,D] f32 in [0,1]; embeddings [sum T,C]; offsets [L+1] int32; outputs [L,B,C] (caller
 * allocated, embeddings dtype); S = log2(per_level_scale); H = base resolution; dy_dx [B,L*D*C] or
 * NULL; gridtype 0 hash / 1 tiled; interp 0 linear / 1 smoothstep.  D in {2,3,4,5}, C in {1,2,4,8}
 * else error (reference: std::runtime_error, gridencoder.cu:L381,L398). */
int ucnerf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                               void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                               uint32_t H, void* dy_dx, uint32_t gridtype, int align_corners,
                               uint32_t interp, int dtype, void* stream);

/* grad [L,B,C]; grad_embeddings [sum T,C] caller-zeroed, accumulated with atomics;
 * dy_dx / grad_inputs [B,D] optional (both NULL or both set; grad_inputs is overwritten). */
int ucnerf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                                const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D,
                                uint32_t C, uint32_t L, float S, uint32_t H, const void* dy_dx,
                                void* grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                                int dtype, void* stream);

/* inputs [B,D] (embeddings dtype, as in the reference), grad accumulated in place. */
int ucnerf_grad_total_variation(const void* inputs, const void* embeddings, void* grad,
                                const int32_t* offsets, float weight, uint32_t B, uint32_t D, uint32_t C,
                                uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                int dtype, void* stream);

/* Fused optimiser step for one GridEncoder table: hash-decay gradient
 * (models.py:L297-306 with train_utils.py:L301-305 `hash_decay_mults`), grad.nan_to_num_() (train_utils.py:L344-345),
 * torch.optim.Adam as train_utils.create_optimizer builds it (L347-366: betas, eps, no weight decay, no amsgrad) and
 * zero_grad (train.py:L164) in ONE pass over embeddings / grad / exp_avg / exp_avg_sq [sum T, C] (device, fp32, C = 4).
 * offsets_host [L+1] as in GridEncoder.offsets; `step` counts from 1; hash_decay_mult = 0 gives plain Adam. */
int ucnerf_grid_adam_step(float* embeddings, float* grad, float* exp_avg, float* exp_avg_sq,
                          const int32_t* offsets_host, uint32_t L, uint32_t C, double lr, double beta1, double beta2,
                          double eps, uint64_t step, double hash_decay_mult, int zero_grad, void* stream);

/* Same with the reference's gradient clipping (train_utils.clip_gradients, L335-345) applied in the pass, before
 * nan_to_num_(): g <- g * grad_scale (the global-norm coefficient min(1, max_norm / (norm + 1e-6)) of
 * torch.nn.utils.clip_grad_norm_, computed by the caller over ALL parameters - this table's share of the squared norm
 * comes from ucnerf_grid_table_stats), then clamp to +-grad_max_val (<= 0: off). */
int ucnerf_grid_adam_step_clipped(float* embeddings, float* grad, float* exp_avg, float* exp_avg_sq,
                                  const int32_t* offsets_host, uint32_t L, uint32_t C, double lr, double beta1,
                                  double beta2, double eps, uint64_t step, double hash_decay_mult, int zero_grad,
                                  double grad_scale, double grad_max_val, void* stream);

/* One read pass over a table: out_sums[2 l] = sum over level l of p^2 and out_sums[2 l + 1] = sum of (g + c_l p)^2 with
 * c_l = hash_decay_mult * 2 / (T_l L C) (grad may be NULL: zeros), device doubles [2 L].  From them:
 *   loss_hash_decay (models.py:L297-306)           = sum_l out[2 l] / (T_l L C)
 *   this table's squared gradient norm for clipping = sum_l out[2 l + 1]  (hash-decay gradient included). */
int ucnerf_grid_table_stats(const float* embeddings, const float* grad, const int32_t* offsets_host, uint32_t L,
                            uint32_t C, double hash_decay_mult, double* out_sums, void* stream);

/* Pooled hash-grid encode for training (SURVEY.md section 8a rows R4 + R10): the front end of MLP.predict_density
 * (internal/models.py:L485-496) in one kernel each way.  means [B,M,3] / stds [B,M] are the multisample Gaussians
 * render.cast_rays returns (M = 6), device fp32; flags & UCNERF_POOLED_CONTRACT applies coord.contract_mean_std (coord.py:L60-72) and the
 * division by bound = 2 (models.py:L489-493), then GridEncoder.forward's (x + 1) / 2 (grid.py:L162), kernel_grid
 * (gridencoder.cu:L87-197; D = 3, hash grid, align_corners = False, linear), the erf down-weighting with
 * grid_sizes (models.py:L495) and the mean over the M points (L496):
 *   features [B, L*C] fp32, coord [B,3] or NULL = means.mean(dim=-2) after contraction (models.py:L512).
 * offsets_host [L+1] / grid_sizes_host [L] are HOST copies of GridEncoder.offsets / .grid_sizes; S = log2(per_level_scale),
 * H = base_resolution as in ucnerf_grid_encode_forward.  C must be 4. */
#define UCNERF_POOLED_CONTRACT 1     /* warp_fn = 'contract' and bound = 2 (models.py:L487-493) */
#define UCNERF_POOLED_MERGE_RUNS 2   /* backward only: sum the corner weights of consecutive points that share a cell before reducing */
#define UCNERF_POOLED_MERGE_RAY_RUNS 4   /* backward only: the same across 4 consecutive intervals (neighbouring samples of a ray) per thread; needs B % 4 == 0, else ignored */
int ucnerf_pooled_encode_forward(const float* means, const float* stds, uint32_t B, uint32_t M, int flags,
                                 const float* embeddings, const int32_t* offsets_host, const int32_t* grid_sizes_host,
                                 uint32_t L, uint32_t C, float S, uint32_t H, float* features, float* coord, void* stream);

/* Backward of the above with respect to the embeddings (means / stds carry no gradient: coord.track_linearize is
 * @torch.no_grad, coord.py:L75): replaces the autograd chain mean -> mul -> permute -> kernel_grid_backward
 * (gridencoder.cu:L248-340, grid.py:L65-89).  grad_features [B, L*C]; grad_embeddings [sum T, C] is ACCUMULATED into
 * (caller-zeroed, like grid.py:L76). */
int ucnerf_pooled_encode_backward(const float* grad_features, const float* means, const float* stds, uint32_t B, uint32_t M,
                                  int flags, const int32_t* offsets_host, const int32_t* grid_sizes_host, uint32_t L,
                                  uint32_t C, float S, uint32_t H, float* grad_embeddings, void* stream);

/* Stand-alone resampling pass of Model.forward's level loop (internal/models.py:L156-205: stepfun.max_dilate_weights
 * L75-105 when dilate != 0, slice of the first/last bin, logits = where(dt > 0, anneal * log(w + padding), -inf),
 * stepfun.sample_intervals L251-294 on the domain [0, 1]) for callers that keep the rest of the level in PyTorch
 * (the training step; the result carries no gradient, models.py:L203-204).  Device fp32 pointers:
 *   t_prev [N, n_prev+1], w_prev [N, n_prev] - the previous level's sdist / weights (both NULL with n_prev = 1 for the
 *   first level, models.py:L143-147); u [S] - the base grid of stepfun.sample (L198-211: linspace(1/2S, 1-1/2S-eps, S) for
 *   rand=False, linspace(0, 1-u_max, S) for rand=True); jitter [N, jitter_cols] or NULL - torch.rand(...) * max_jitter
 *   (L212), jitter_cols = 1 (single_jitter) or S; out_sdist [N, S+1]. */
int ucnerf_resample_intervals(const float* t_prev, const float* w_prev, uint32_t n_rays, int32_t n_prev, int dilate,
                              float dilation, float anneal, float padding, int32_t S, const float* u, const float* jitter,
                              int32_t jitter_cols, float* out_sdist, void* stream);

/* Differentiable alpha compositing for the training step (render.py:L155-174 compute_alpha_weights with
 * opaque_background = False, and the acc / bg_w / rgb lines of volumetric_rendering, L202-205).  Device fp32:
 * tdist [N, S+1] metric fenceposts (no gradient: sdist is detached, models.py:L203-204), density [N, S], rgbs [N, S, 3] or
 * NULL (proposal levels), dirs [N, 3] (NOT unit norm), bg = the constant background intensity.
 * forward  -> weights [N, S], rgb [N, 3], acc [N].
 * backward: incoming g_weights [N, S] / g_rgb [N, 3] / g_acc [N] (each may be NULL = no gradient), the forward's
 *           weights / acc -> d_density [N, S], d_rgbs [N, S, 3] (NULL allowed; ignored when rgbs is NULL). */
int ucnerf_composite_train_forward(const float* tdist, const float* density, const float* rgbs, const float* dirs, uint32_t N,
                                   int32_t S, float bg, float* weights, float* rgb, float* acc, void* stream);
int ucnerf_composite_train_backward(const float* tdist, const float* density, const float* rgbs, const float* dirs,
                                    const float* weights, const float* acc, const float* g_weights, const float* g_rgb,
                                    const float* g_acc, uint32_t N, int32_t S, float bg, float* d_density, float* d_rgbs,
                                    void* stream);

/* Stand-alone render.cast_rays (internal/render.py:L94-152) for the training step.  Device fp32: tdist [N, S+1] metric
 * fenceposts, origins / directions / cam_dirs [N,3], radii [N], rand_vec [N,3] = the torch.randn_like(cam_dirs) draw
 * of L140; rot01 / flip01 [N,S] = the two torch.rand_like draws of rand=True (L121-122: rotation, flip mask), both NULL
 * for the deterministic pattern (L125-131).  Outputs means [N,S,6,3], stds [N,S,6], ts [N,S,6] (ts may be NULL). */
int ucnerf_cast_rays(const float* tdist, const float* origins, const float* directions, const float* cam_dirs,
                     const float* radii, const float* rand_vec, const float* rot01, const float* flip01, uint32_t n_rays,
                     int32_t S, float std_scale, float* means, float* stds, float* ts, void* stream);

/* ---- fused forward render (eval path, rand=False) ---- */

/* One MLP's GridEncoder + density_layer (models.py:L425-441).  Pointers are device pointers to the
 * tensors of the reference state_dict, in their native layouts (nn.Linear weight = [out,in]). */
typedef struct ucnerf_mlp_desc {
    const float* embeddings;      /* <prefix>.encoder.embeddings [sum T, C] (NOT copied: read at render time) */
    const int32_t* offsets_host;  /* <prefix>.encoder.offsets  [L+1]  HOST pointer (copied)             */
    const int32_t* grid_sizes_host; /* <prefix>.encoder.grid_sizes [L] HOST pointer (copied)             */
    int32_t grid_levels;          /* L */
    int32_t level_dim;            /* C (must be 4 on the fused path)                                     */
    int32_t base_resolution;      /* H */
    float log2_per_level_scale;   /* S */
    const float* density0_w;      /* density_layer.0.weight [64, L*C] */
    const float* density0_b;      /* density_layer.0.bias   [64]      */
    const float* density2_w;      /* density_layer.2.weight [1 or bottleneck, 64] */
    const float* density2_b;      /* density_layer.2.bias */
} ucnerf_mlp_desc;

typedef struct ucnerf_model_desc {
    int32_t num_prop_levels;      /* Model.num_levels - 1 (>=1, <= UCNERF_MAX_PROP_LEVELS)  */
    int32_t num_prop_samples;     /* Model.num_prop_samples                                  */
    int32_t num_nerf_samples;     /* Model.num_nerf_samples                                  */
    int32_t bottleneck_width;     /* MLP.bottleneck_width                                    */
    int32_t net_width_viewdirs;   /* MLP.net_width_viewdirs (net_depth_viewdirs=2, skip_layer_dir=0 fixed) */
    int32_t deg_view;             /* MLP.deg_view (4)                                        */
    /* python floats of the reference are carried as doubles so derived fp32 constants round identically */
    double dilation_multiplier;   /* Model.dilation_multiplier */
    double dilation_bias;         /* Model.dilation_bias       */
    double anneal_slope;          /* Model.anneal_slope        */
    double resample_padding;      /* Model.resample_padding    */
    double std_scale;             /* Model.std_scale           */
    double bg_intensity;          /* Model.bg_intensity_range (min==max)                      */
    double density_bias;          /* MLP.density_bias          */
    double rgb_padding;           /* MLP.rgb_padding           */
    ucnerf_mlp_desc prop[UCNERF_MAX_PROP_LEVELS];
    ucnerf_mlp_desc nerf;
    const float* view0_w;         /* nerf_mlp.lin_second_stage_0.weight [W, bottleneck+dir]   */
    const float* view0_b;
    const float* view1_w;         /* nerf_mlp.lin_second_stage_1.weight [W, W+bottleneck+dir] */
    const float* view1_b;
    const float* rgb_w;           /* nerf_mlp.rgb_layer.weight [3, W] */
    const float* rgb_b;
} ucnerf_model_desc;

typedef struct ucnerf_model ucnerf_model;  /* opaque */

/* Copies / re-lays-out the (small) MLP weights into library-owned device memory; keeps the
 * embeddings pointers.  Call ucnerf_model_refresh after the source weights changed. */
int ucnerf_model_create(const ucnerf_model_desc* desc, ucnerf_model** out);
int ucnerf_model_refresh(ucnerf_model* m, const ucnerf_model_desc* desc, void* stream);
int ucnerf_model_destroy(ucnerf_model* m);

/* Ray batch (datasets.py:L386-476 keys).  All [N,3] / [N] f32, contiguous.  rand_vec is the
 * cone-basis vector the reference draws with torch.randn_like(cam_dirs) (render.py:L140). */
typedef struct ucnerf_rays {
    const float* origins;
    const float* directions;
    const float* viewdirs;
    const float* cam_dirs;
    const float* radii;
    const float* near;
    const float* far;
    const float* rand_vec;
} ucnerf_rays;

/* Outputs; any pointer may be NULL (not produced).  Level index l = 0..num_prop_levels (last = NeRF). */
typedef struct ucnerf_outputs {
    float* rgb;                   /* [N,3] final level */
    float* depth;                 /* [N]   with the reference's acc<0.6 -> 300 override (render.py:L208,L213) */
    float* depth_raw;             /* [N]   before that override */
    float* acc;                   /* [N] */
    float* distance_mean;         /* [N] */
    float* distance_median;       /* [N] */
    float* distance_percentile_5; /* [N] */
    float* distance_percentile_95;/* [N] */
    float* sdist[UCNERF_MAX_PROP_LEVELS + 1];   /* [N, S_l+1] normalised fenceposts of level l */
    float* weights[UCNERF_MAX_PROP_LEVELS + 1]; /* [N, S_l] */
    float* sample_rgb;            /* [N, S_nerf, 3] per-sample colours of the NeRF level */
    float* sample_density;        /* [N, S_nerf] */
    float* packed;                /* [N,12] (rgb[3], depth, acc, distance_mean, distance_median, distance_percentile_5,
                                     distance_percentile_95, depth_raw, 0, 0): the one buffer the multi-GPU
                                     tile all-gather moves */
    float* sample_coord;          /* [N, S_nerf, 3] `coord` of the NeRF level (models.py:L512,L677): mean of the six
                                     contracted multisample positions / 2 */
} ucnerf_outputs;

/* Rays and outputs in device memory.  train_frac only enters through the anneal (models.py:L179-184). */
int ucnerf_render_rays(ucnerf_model* m, uint64_t n_rays, const ucnerf_rays* rays, double train_frac,
                       const ucnerf_outputs* out, void* stream);

/* Same, with HOST buffers (pinned or pageable): copies the ray batch in, renders, copies every
 * non-NULL output back and synchronises the stream.  This is the end-to-end entry bench.py times. */
int ucnerf_render_rays_host(ucnerf_model* m, uint64_t n_rays, const ucnerf_rays* rays_host, double train_frac,
                            const ucnerf_outputs* out_host, void* stream);

/* ---- camera -> rays on the GPU (eval loader's numpy ray generation, camera_utils.py:L448-557) ----
 * One pinhole camera (perspective, no lens distortion, no NDC).  Matrices are doubles holding the values the
 * reference holds (its Waymo loader keeps float32 intrinsics / poses; numpy promotes to float64 for the arithmetic,
 * the results are cast to float32 - the kernel does the same). */
typedef struct ucnerf_camera {
    double pixtocam[9];     /* inverse intrinsic matrix, row-major 3x3 (Dataset.pixtocams[i])       */
    double camtoworld[12];  /* pose, row-major 3x4, OpenGL convention (Dataset.camtoworlds[i])      */
    uint32_t width, height;
    float near, far;        /* Dataset.near / .far, broadcast to every ray                          */
    uint64_t rand_seed;     /* seed of the counter-based N(0,1) cone-basis vectors (render.py:L140) */
} ucnerf_camera;

/* Writable ray batch for rows [row0, row0 + n_rows) of the image, n = n_rows * width rays in row-major pixel
 * order (pixel_coordinates, camera_utils.py:L368-370).  Any of origins, cam_dirs, near, far, imageplane [n,2],
 * rand_vec may be NULL.  rand_vec: the reference draws torch.randn_like(cam_dirs) at every call; here a
 * reproducible counter-based draw keyed by (rand_seed, pixel index). */
typedef struct ucnerf_ray_buffers {
    float* origins;
    float* directions;
    float* viewdirs;
    float* cam_dirs;
    float* radii;
    float* near;
    float* far;
    float* rand_vec;
    float* imageplane;
} ucnerf_ray_buffers;

int ucnerf_generate_rays(const ucnerf_camera* cam, uint32_t row0, uint32_t n_rows, const ucnerf_ray_buffers* out,
                         void* stream);

/* generate_rays + render_rays for image rows [row0, row0 + n_rows): the frame is rendered from the ~200 bytes of
 * camera parameters; the ray batch lives in library-owned device memory.  Outputs as in ucnerf_render_rays
 * (device pointers) / ucnerf_render_rays_host (host pointers, copies back and synchronises). */
int ucnerf_render_camera(ucnerf_model* m, const ucnerf_camera* cam, uint32_t row0, uint32_t n_rows, double train_frac,
                         const ucnerf_outputs* out, void* stream);
int ucnerf_render_camera_host(ucnerf_model* m, const ucnerf_camera* cam, uint32_t row0, uint32_t n_rows,
                              double train_frac, const ucnerf_outputs* out_host, void* stream);

/* Number of kernels launched by this library in this process so far (bench.py's gpu_launches). */
uint64_t ucnerf_launch_count(void);

/* Tunables: "chunk_rays" (rays per internal chunk), "color_mlp" (0 = fp32 SIMT, 1 = tcgen05 FP16 3-term split, 2 = auto:
 * tensor cores whenever the MLP widths allow, the default), "timing" (see ucnerf_get_timing), "encode_mlp_mma" (density
 * layer of the sample/encode kernel as mma.sync 3xTF32: bit 0 = proposal levels, bit 1 = NeRF level, default 3; 0 = the
 * fp32 FMA forms), "encode_runs" (cell-run reuse of gathered corners, bit-identical, default 0; implies the FMA forms),
 * "warp_rays_prop" / "warp_rays_nerf" (rays per warp of that kernel: 32, 16, 8 or 4), "ray_tile_width" (W > 0, a multiple of
 * 4: the ray batches of ucnerf_render_rays[_host] are whole rows of a row-major image W pixels wide, so the warps of that
 * kernel may take pixel patches instead of 32 pixels of one row - results are bit-identical for ANY batch, the hint only
 * changes which rays share a warp; the camera entries know W and do this by themselves; 0 = off, the default),
 * "ray_tile_prop" / "ray_tile_nerf" (patch width per level kind: 0 = rows, 4 = 4x8, 8 = 8x4, 16 = 16x2; default 0 / 4). */
int ucnerf_set_option(ucnerf_model* m, const char* key, int64_t value);

/* Brightness-correction head folded into the compositing epilogue.  The reference evaluates
 * BrightnessCorrection (extrinsic_optimizer.py:L4-25: latent code -> 3x256 MLP -> 3x4 affine) once PER RAY for the
 * single camera index of an eval image and applies rgb <- A[:3,:3] rgb + A[:3,3] to the rendered colour
 * (models.py:L339-363).  Here the caller evaluates the head once per image and passes the affine (row-major [3][4],
 * HOST pointer); every following render applies it to the final level's rgb (also inside `packed`).  NULL switches it
 * off.  The sky branch of that code (model_sky=True) is not part of the fused path. */
int ucnerf_set_rgb_affine(ucnerf_model* m, const float* affine12_host);

/* ---- fp32-accurate tensor-core GEMMs for the training step's dense layers (csrc/gemm3_tc.cu) ----
 * Replace the cuBLAS fp32 SGEMMs that nn.Linear runs for the reference's MLPs (internal/models.py:L438-441, L475-483,
 * L643-652) and their autograd: tcgen05 kind::tf32 with the 3xTF32 split (hi*hi + lo*hi + hi*lo, fp32 accumulate in
 * TMEM), ~2^-21 relative error per product, no operand scaling.  All matrices are device fp32, row-major.
 *   ucnerf_gemm_nt: C[M,N] = sum_s A_s[M,k_s] B_s[N,k_s]^T (+ bias[N]) (relu)   N <= 256, sum of ceil(k_s / 32) <= 24
 *                   forward y = [x_0 | x_1 | ..] W^T + b with B_s = W[:, cols of segment s] (ldb = W's row length),
 *                   input gradient dx_s = dy W_s with B = W_s^T
 *   ucnerf_gemm_tn: C[N1,N2] += A[M,N1]^T B[M,N2]   (reduction over the M rows; C initialised by the caller)
 *                   weight gradient dW_s = dy^T x_s
 * ucnerf_gemm_status: 0, or 5 (+ ucnerf_last_error) when a kernel's pipeline watchdog fired. */
typedef struct ucnerf_gemm_seg {
    const float* a;   /* [M, k]  rows lda apart */
    const float* b;   /* [N, k]  rows ldb apart */
    uint32_t lda, ldb, k;
} ucnerf_gemm_seg;
int ucnerf_gemm_nt(uint32_t M, uint32_t N, uint32_t nseg, const ucnerf_gemm_seg* segs, const float* bias, int relu, float* C,
                   uint32_t ldc, void* stream);
int ucnerf_gemm_tn(uint32_t M, uint32_t N1, uint32_t N2, const float* A, uint32_t lda, const float* B, uint32_t ldb, float* C,
                   uint32_t ldc, void* stream);
int ucnerf_gemm_status(uint32_t* out32);
/* Backward glue of a dense layer in one pass: g[M,N] = gy * (y > 0) (y NULL: g = gy; g NULL: not written) and
 * colsum[N] = column sums of g (the bias gradient; NULL: skipped).  Contiguous fp32 [M,N], N a multiple of 4 <= 1024. */
int ucnerf_relu_mask_colsum(const float* gy, const float* y, float* g, float* colsum, uint32_t M, uint32_t N, void* stream);

/* ---- fused tile exchange over NVLink peer memory ----
 * Multi-GPU render without a trailing all-gather: while peer targets are set, the compositing kernel of the final level
 * stores every finished packed row [12 floats, layout of ucnerf_outputs.packed] of ray i of the call into
 * peer_images[k] + 12 * (row0 + i) for k < n_peers - the image buffers of all ranks (this rank's own included), peer
 * ones mapped with ucnerf_peer_open.  n_peers = 0 switches it off.  The caller orders frame completion across ranks
 * (one tiny all-reduce after the render, ucnerf_b200/peer.py).  The reference instead gathers every leaf of every
 * 15k-ray chunk with accelerate.gather (models.py:L965-968). */
#define UCNERF_MAX_PEERS 16
int ucnerf_set_peer_targets(ucnerf_model* m, uint32_t n_peers, void* const* peer_images, uint64_t row0);

/* Image buffers for that exchange: plain device allocations with a 64-byte IPC handle the owner sends to its peers
 * (alloc / free on the owner, open / close on a peer; zero-initialised). */
int ucnerf_peer_alloc(uint64_t bytes, void** dptr_out, uint8_t* handle64_out);
int ucnerf_peer_open(const uint8_t* handle64, void** dptr_out);
int ucnerf_peer_close(void* dptr);
int ucnerf_peer_free(void* dptr);

/* ---- sky head on tensor cores ----
 * Replaces models.py:L326-337: ray_batch = [origins, directions, near = far, far = 1.5 far[0], cam_dirs] ->
 * render_rays(network_fn = skynerf) -> rgb_map (models.py:L743-904: NeRF D=8 W=256 raw-xyz input, skip after layer 4,
 * 4-frequency view embedding; 120 samples; raw2outputs), bug-compatible with the reference's sample depths
 * z = near (1 - t) + (1 / far) t.  Weight pointers are DEVICE pointers to the reference state_dict tensors
 * (nn.Linear layout [out, in]); they are copied / re-laid-out at creation. */
typedef struct ucnerf_sky_desc {
    const float* pts_w[8];   /* skynerf.pts_linears.{0..7}.weight  [256,3] [256,256]x4 [256,259] [256,256]x2 */
    const float* pts_b[8];   /* skynerf.pts_linears.{0..7}.bias    [256] */
    const float* feature_w;  /* skynerf.feature_linear.weight [256,256] */
    const float* feature_b;
    const float* alpha_w;    /* skynerf.alpha_linear.weight [1,256] */
    const float* alpha_b;
    const float* views_w;    /* skynerf.views_linears.0.weight [128,283] */
    const float* views_b;
    const float* rgb_w;      /* skynerf.rgb_linear.weight [3,128] */
    const float* rgb_b;
    int32_t n_samples;       /* render_rays N_samples (120) */
} ucnerf_sky_desc;

typedef struct ucnerf_sky ucnerf_sky;  /* opaque */
int ucnerf_sky_create(const ucnerf_sky_desc* desc, ucnerf_sky** out);
int ucnerf_sky_destroy(ucnerf_sky* sky);
/* origins, directions, views (= the batch's cam_dirs) [N,3], far [N] device arrays; sky_far = 1.5 * far[0] as the
 * reference computes it on the host (models.py:L329); sky_rgb [N,3] device output = ret['rgb_map']. */
int ucnerf_sky_render(ucnerf_sky* sky, uint64_t n_rays, const float* origins, const float* directions, const float* far,
                      const float* views, double sky_far, float* sky_rgb, void* stream);

/* Timing probe: with option "timing" != 0, ucnerf_render_rays records a CUDA event pair around every kernel
 * launch on the launch stream (no synchronisation is added).  ucnerf_get_timing waits for the recorded events
 * and returns accumulated device milliseconds and launch counts per kernel family:
 * [resample, encode_prop, encode_nerf, color_mlp, composite].  launches_out5 may be NULL. */
int ucnerf_get_timing(ucnerf_model* m, float* ms_out5, uint32_t* launches_out5, int reset);

#ifdef __cplusplus
}
#endif
#endif /* UCNERF_B200_H */

// Kernel parameter blocks + launch wrappers of the fused forward-render path (ray_march.cu).
#pragma once
#include "common.cuh"
#include "ray_algos.cuh"

namespace ucnerf {

struct RayPtrs {
    const float *origins, *directions, *viewdirs, *cam_dirs, *radii, *near, *far, *rand_vec;
};

// outputs of generate_rays_kernel (any of origins / cam_dirs / near / far / imageplane / rand_vec may be NULL)
struct RayOutPtrs {
    float *origins, *directions, *viewdirs, *cam_dirs, *radii, *near, *far, *imageplane, *rand_vec;
};

struct ResampleParams {
    uint32_t n_rays;
    int n_prev;              // bins of the previous level (1 for the first level)
    const float* t_prev;     // [N, n_prev+1] or NULL
    uint32_t t_prev_stride;  // floats between rays (0: one row shared by all rays, e.g. the first level)
    const float* w_prev;     // [N, n_prev]   or NULL
    int dilate;
    float dilation, anneal, padding;
    int S;
    const float* u;          // [S]
    float* out_sdist;        // [N, S+1]
    float* dbg_scratch;      // debugging: scratch arrays of ray dbg_ray after the call (NULL in production)
    uint32_t dbg_ray;
};

struct SampleParams {
    uint32_t n_rays;
    int S;
    RayPtrs rays;
    const float* sdist;      // [N, S+1]
    uint32_t sdist_stride;   // floats between rays (0: shared row)
    GridDesc grid;
    ConeTable cone;
    float std_scale, density_bias;
    const float* w1p;        // [64][LMAX*4] zero padded
    const float* b1;         // [64]
    const float* w2;         // [64] (row 0 of density_layer.2)
    float b2;
    float* density;          // [N, S]
    float* h1;               // [N*S, 64] (NeRF level only)
    float g2[16];            // float(grid_sizes[l]^2)
    int rw_log2;             // warp shape of sample_encode_kernel: 2^rw_log2 rays x 2^(5 - rw_log2) samples (5 = 32 rays x 1)
    int cell_runs;           // 1: level-outer loop with cell-run reuse of the gathered corners (sample_encode.cu)
    uint32_t tile_w;         // > 0: the rays of this launch are whole rows of an image of this width (row-major): a warp of
                             // sample_encode_kernel then takes a patch of 2^tile_pw_log2 x 32 / 2^tile_pw_log2 pixels
                             // instead of 32 pixels of one row
    uint32_t tile_pw_log2;   // 2, 3 or 4 (patch 4 x 8, 8 x 4, 16 x 2)
    int mlp_mma;             // 1: density layer on the tensor cores (mma.sync 3xTF32), 0: FFMA2 forms (sample_encode.cu)
    const float* geom;       // [15][geom_ld] per-ray RayGeom (o, d, e1, e2, radius, near, far) from ray_geom_kernel, or NULL:
    uint32_t geom_ld;        // every sample of a ray needs the same basis - built once per ray, read coalesced
};

// Colour MLP with the linear bottleneck layer folded into its two consumers (exact algebra, see model.cu):
//   a   = relu(P0 [h1, direnc] + c0')          P0  = [V0x W2 | V0d]          c0' = c0 + V0x b2
//   a2  = relu(V1a a + P1 [h1, direnc] + c1')  P1  = [V1x W2 | V1d]          c1' = c1 + V1x b2
//   rgb = sigmoid(R a2 + r0) * (1 + 2 pad) - pad
struct ColorParams {
    uint32_t n_rows;         // N * S
    int S;
    int deg_view;
    const float* h1;         // [rows, 64]
    const float* viewdirs;   // [N, 3]
    const float *p0t, *c0;   // [96][NP], [NP]          rows: h1 (64), direnc (32, zero padded)
    const float *v1t, *c1;   // [NP + 96][NP], [NP]     rows: a (NP), h1 (64), direnc (32)
    const float *rt, *r0;    // [NP][4], [4]
    float rgb_scale, rgb_padding;   // (float)(1 + 2 pad), (float)pad
    float* rgb;              // [rows, 3]
};

// tensor-core variant (color_mlp_tc.cu): W = 256, deg_view = 4 only; weights pre-scaled / pre-split / pre-swizzled (wblob)
struct ColorTcParams {
    uint32_t n_rows;
    int S;
    const float* h1;         // [rows, 64]
    const float* viewdirs;   // [N, 3]
    const uint8_t* wblob;    // 6 chunks x (hi tile | lo tile) FP16 in UMMA K-major SWIZZLE_128B layout
    const float* dir_bias;   // [N rays][512]: per-ray biases of both layers (dir_bias_kernel; first half x act scale)
    float k0, k1;            // accumulator -> value factors: 1 / scale(P0), 1 / (act scale * scale(P1, V1a))
    const float *rt, *r0;    // [256][4], [4]
    float rgb_scale, rgb_padding;
    float* rgb;              // [rows, 3]
    uint32_t* dbg;           // watchdog record (set by the launcher)
    uint32_t debug_flags;    // profiling experiments only (ucnerf_set_option "tc_debug"); 0 in production
};

// sky head (sky_mlp_tc.cu): 8x256 NeRF MLP at n_samples depths per ray, tensor cores
struct SkyTcParams {
    uint32_t n_rows;          // n_rays * n_samples
    int n_samples;
    const float *origins, *directions, *far;   // per ray
    const float* t_vals;      // [n_samples] torch.linspace(0, 1, n_samples)
    float sky_far;            // 1.5 x far[0] (models.py:L329)
    const float* view_bias;   // [n_rays][128]: b_v + W_vf b_f + W_v[:, 256:283] emb(view)
    const uint8_t* wblob;     // 34 weight chunks (hi | lo FP16, UMMA K-major SWIZZLE_128B), 64 KB stride
    const float* bias8;       // [8][256]: biases of layers 0..7, x activation scale
    float k[10];              // accumulator -> value factors: layers 0..7, [8] = folded view layer
    const float* w_alpha;     // [256]
    float b_alpha;
    const float* rgb_w;       // [128][4]
    float rgb_b[3];
    float* raw;               // [n_rows][4] = rgb_raw, alpha_raw
    uint32_t* dbg;
    uint32_t debug_flags;     // bit 2: in-kernel wait profiler (env UCNERF_SKY_DEBUG, development only)
    // second pipeline (sky_mlp_tc2.cu): 60 weight half-chunks (128 output columns x 64 K, hi | lo FP16, 32 KB stride) and
    // the two K = 3 blocks for the CUDA cores: [256] float4 = act scale x (W0[c][0..2], b0[c]) / (W5[c][0..2], 0)
    const uint8_t* wblob2;
    const float *w0x, *w5x;
};

struct CompositeParams {
    uint32_t n_rays;
    int S;
    const float* sdist;      // [N, S+1]
    uint32_t sdist_stride;   // floats between rays (0: shared row)
    const float* density;    // [N, S]
    const float* rgb;        // [N, S, 3] or NULL
    RayPtrs rays;
    float bg;
    int extras;
    float* weights;          // [N, S]
    // optional per-ray outputs (NULL = skip)
    float *o_rgb, *o_depth, *o_depth_raw, *o_acc, *o_mean, *o_median, *o_p5, *o_p95, *o_packed;
    int n_peers;             // fused tile exchange: the packed row also goes to peer_packed[k] + 12 * (peer_row0 + ray)
    float* peer_packed[16];  // image buffers of all ranks (NVLink peer mappings + this rank's own), see peer.cu
    uint64_t peer_row0;
    int use_affine;          // final level only: rgb <- A rgb + t (BrightnessCorrection, models.py:L339-363)
    float affine[12];        // row-major [3][4]
};

int launch_generate_rays(const CameraConst& cam, uint32_t row0, uint32_t n_rows, const RayOutPtrs& o, cudaStream_t st);
int launch_resample(const ResampleParams& p, cudaStream_t st);
int launch_sample_encode(const SampleParams& p, bool nerf, cudaStream_t st);
int launch_sample_coord(const SampleParams& p, float* out, cudaStream_t st);
int launch_color_mlp_simt(const ColorParams& p, int np, cudaStream_t st);
int launch_composite(const CompositeParams& p, cudaStream_t st);
int launch_color_mlp_tc(const ColorTcParams& p, cudaStream_t st);
uint32_t color_tc_blob_bytes();
int color_tc_status(uint32_t* out16);
int launch_dir_bias(const float* viewdirs, const float* wdir, const float* c0, const float* c1, float* out,
                    uint32_t n_rays, cudaStream_t st);
void color_tc_pack_chunk(const float* wt_rows, float scale, uint8_t* dst);
float color_tc_weight_scale(const float* w, size_t n);
float color_tc_act_scale();
int launch_sky_mlp_tc(const SkyTcParams& p, cudaStream_t st);
int launch_sky_view_bias(const float* views, const float* wv_view, const float* bv, float* out, uint32_t n_rays, cudaStream_t st);
int launch_sky_composite(const float* raw, const float* directions, const float* far, const float* t_vals, float sky_far,
                         int n_samples, float* out, uint32_t n_rays, cudaStream_t st);
int launch_sky_mlp_tc2(const SkyTcParams& p, uint32_t* dbg, cudaStream_t st);
uint32_t sky_tc2_blob_bytes();
int sky_tc2_half_steps();
uint32_t* sky_tc_dbg_buffer();
int sky_tc_status(uint32_t* out32);
uint32_t sky_tc_blob_bytes();
int sky_tc_steps();
float sky_tc_act_scale();
void sky_tc_pack_chunk(const float* wt_rows, int n_cols, float scale, uint8_t* dst);
int sample_encode_lmax(int L);
int launch_ray_geom(const RayPtrs& rays, uint32_t n_rays, float* geom, uint32_t geom_ld, cudaStream_t st);
// h1 column c holds hidden unit h1_perm(c) of density_layer.0 (layout written by sample_encode_kernel): lane t of a
// quad owns the accumulator columns {8 nt + 2 t + e} of the mma.sync C fragments.  They are stored as two 32-byte pieces
// (nt < 4 / nt >= 4) at columns 32 (nt / 4) + 8 t + 2 (nt % 4) + e, so that one 256-bit store instruction of a quad covers
// one whole 128-byte line of the row.  h1_col is the inverse.
__host__ __device__ inline int h1_perm(int c) { return 8 * (4 * (c / 32) + (c % 8) / 2) + 2 * ((c % 32) / 8) + (c % 2); }
__host__ __device__ inline int h1_col(int u) { return 32 * (u / 32) + 8 * ((u % 8) / 2) + 2 * ((u / 8) % 4) + (u % 2); }

}  // namespace ucnerf

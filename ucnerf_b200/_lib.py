"""ctypes binding of libucnerf_b200.so (C ABI in include/ucnerf_b200.h).

There is NO CPU fallback: if the shared library is missing or a CUDA device is absent the calls
raise.  The library is built in-tree by `python -m ucnerf_b200.build` (nvcc, sm_100a)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UCNERF_B200_LIB") or os.path.join(_HERE, "libucnerf_b200.so")  # env: A/B builds

MAX_GRID_LEVELS = 16
MAX_PROP_LEVELS = 4
F32, F16, F64 = 0, 1, 2

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)


class MlpDesc(C.Structure):
    _fields_ = [
        ("embeddings", C.c_void_p),
        ("offsets_host", C.c_void_p),
        ("grid_sizes_host", C.c_void_p),
        ("grid_levels", C.c_int32),
        ("level_dim", C.c_int32),
        ("base_resolution", C.c_int32),
        ("log2_per_level_scale", C.c_float),
        ("density0_w", C.c_void_p),
        ("density0_b", C.c_void_p),
        ("density2_w", C.c_void_p),
        ("density2_b", C.c_void_p),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("num_prop_levels", C.c_int32),
        ("num_prop_samples", C.c_int32),
        ("num_nerf_samples", C.c_int32),
        ("bottleneck_width", C.c_int32),
        ("net_width_viewdirs", C.c_int32),
        ("deg_view", C.c_int32),
        ("dilation_multiplier", C.c_double),
        ("dilation_bias", C.c_double),
        ("anneal_slope", C.c_double),
        ("resample_padding", C.c_double),
        ("std_scale", C.c_double),
        ("bg_intensity", C.c_double),
        ("density_bias", C.c_double),
        ("rgb_padding", C.c_double),
        ("prop", MlpDesc * MAX_PROP_LEVELS),
        ("nerf", MlpDesc),
        ("view0_w", C.c_void_p),
        ("view0_b", C.c_void_p),
        ("view1_w", C.c_void_p),
        ("view1_b", C.c_void_p),
        ("rgb_w", C.c_void_p),
        ("rgb_b", C.c_void_p),
    ]


class Rays(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("origins", "directions", "viewdirs", "cam_dirs", "radii", "near", "far", "rand_vec")]


class Outputs(C.Structure):
    _fields_ = [
        ("rgb", C.c_void_p), ("depth", C.c_void_p), ("depth_raw", C.c_void_p), ("acc", C.c_void_p),
        ("distance_mean", C.c_void_p), ("distance_median", C.c_void_p),
        ("distance_percentile_5", C.c_void_p), ("distance_percentile_95", C.c_void_p),
        ("sdist", C.c_void_p * (MAX_PROP_LEVELS + 1)),
        ("weights", C.c_void_p * (MAX_PROP_LEVELS + 1)),
        ("sample_rgb", C.c_void_p), ("sample_density", C.c_void_p), ("packed", C.c_void_p),
        ("sample_coord", C.c_void_p),
    ]


class GemmSeg(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("lda", C.c_uint32), ("ldb", C.c_uint32), ("k", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("pixtocam", C.c_double * 9), ("camtoworld", C.c_double * 12), ("width", C.c_uint32),
                ("height", C.c_uint32), ("near", C.c_float), ("far", C.c_float), ("rand_seed", C.c_uint64)]


class RayBuffers(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("origins", "directions", "viewdirs", "cam_dirs", "radii", "near", "far", "rand_vec", "imageplane")]


class SkyDesc(C.Structure):
    _fields_ = [("pts_w", C.c_void_p * 8), ("pts_b", C.c_void_p * 8), ("feature_w", C.c_void_p), ("feature_b", C.c_void_p),
                ("alpha_w", C.c_void_p), ("alpha_b", C.c_void_p), ("views_w", C.c_void_p), ("views_b", C.c_void_p),
                ("rgb_w", C.c_void_p), ("rgb_b", C.c_void_p), ("n_samples", C.c_int32)]


# every symbol include/ucnerf_b200.h declares (tests/test_abi_symbols.py checks the .so exports them)
EXPORTS = [
    "ucnerf_abi_version", "ucnerf_last_error", "ucnerf_grid_encode_forward", "ucnerf_grid_encode_backward",
    "ucnerf_grad_total_variation", "ucnerf_model_create", "ucnerf_model_refresh", "ucnerf_model_destroy",
    "ucnerf_render_rays", "ucnerf_render_rays_host", "ucnerf_launch_count", "ucnerf_set_option",
    "ucnerf_get_timing", "ucnerf_generate_rays", "ucnerf_render_camera", "ucnerf_render_camera_host",
    "ucnerf_set_rgb_affine", "ucnerf_set_peer_targets", "ucnerf_peer_alloc", "ucnerf_peer_open", "ucnerf_peer_close",
    "ucnerf_peer_free", "ucnerf_gemm_nt", "ucnerf_gemm_tn", "ucnerf_gemm_status", "ucnerf_relu_mask_colsum", "ucnerf_sky_create", "ucnerf_sky_destroy", "ucnerf_sky_render",
    "ucnerf_grid_adam_step", "ucnerf_grid_adam_step_clipped", "ucnerf_grid_table_stats",
    "ucnerf_pooled_encode_forward",
    "ucnerf_pooled_encode_backward",
    "ucnerf_resample_intervals",
    "ucnerf_composite_train_forward",
    "ucnerf_composite_train_backward",
    "ucnerf_cast_rays",
]

_lib = None


class UcnerfError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built - never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UcnerfError(
            f"{LIB_PATH} not found: build it with `python -m ucnerf_b200.build` (nvcc, sm_100a). "
            "ucnerf_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    u32, i32, f32, vp = C.c_uint32, C.c_int32, C.c_float, C.c_void_p
    lib.ucnerf_abi_version.restype = C.c_int
    lib.ucnerf_last_error.restype = C.c_char_p
    lib.ucnerf_launch_count.restype = C.c_uint64
    lib.ucnerf_grid_encode_forward.argtypes = [vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, vp, u32, C.c_int, u32,
                                               C.c_int, vp]
    lib.ucnerf_grid_encode_backward.argtypes = [vp, vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, vp, vp, u32,
                                                C.c_int, u32, C.c_int, vp]
    lib.ucnerf_grad_total_variation.argtypes = [vp, vp, vp, vp, f32, u32, u32, u32, u32, f32, u32, u32, C.c_int,
                                                C.c_int, vp]
    lib.ucnerf_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(vp)]
    lib.ucnerf_model_refresh.argtypes = [vp, C.POINTER(ModelDesc), vp]
    lib.ucnerf_model_destroy.argtypes = [vp]
    lib.ucnerf_render_rays.argtypes = [vp, C.c_uint64, C.POINTER(Rays), C.c_double, C.POINTER(Outputs), vp]
    lib.ucnerf_render_rays_host.argtypes = [vp, C.c_uint64, C.POINTER(Rays), C.c_double, C.POINTER(Outputs), vp]
    lib.ucnerf_generate_rays.argtypes = [C.POINTER(Camera), u32, u32, C.POINTER(RayBuffers), vp]
    lib.ucnerf_render_camera.argtypes = [vp, C.POINTER(Camera), u32, u32, C.c_double, C.POINTER(Outputs), vp]
    lib.ucnerf_render_camera_host.argtypes = [vp, C.POINTER(Camera), u32, u32, C.c_double, C.POINTER(Outputs), vp]
    lib.ucnerf_set_rgb_affine.argtypes = [vp, vp]
    lib.ucnerf_set_peer_targets.argtypes = [vp, u32, C.POINTER(vp), C.c_uint64]
    lib.ucnerf_peer_alloc.argtypes = [C.c_uint64, C.POINTER(vp), C.c_char_p]
    lib.ucnerf_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.ucnerf_peer_close.argtypes = [vp]
    lib.ucnerf_peer_free.argtypes = [vp]
    lib.ucnerf_gemm_nt.argtypes = [u32, u32, u32, C.POINTER(GemmSeg), vp, C.c_int, vp, u32, vp]
    lib.ucnerf_gemm_tn.argtypes = [u32, u32, u32, vp, u32, vp, u32, vp, u32, vp]
    lib.ucnerf_gemm_status.argtypes = [C.POINTER(C.c_uint32)]
    lib.ucnerf_relu_mask_colsum.argtypes = [vp, vp, vp, vp, u32, u32, vp]
    lib.ucnerf_grid_adam_step.argtypes = [vp, vp, vp, vp, vp, u32, u32, C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_uint64, C.c_double, C.c_int, vp]
    lib.ucnerf_grid_adam_step_clipped.argtypes = [vp, vp, vp, vp, vp, u32, u32, C.c_double, C.c_double, C.c_double,
                                                  C.c_double, C.c_uint64, C.c_double, C.c_int, C.c_double, C.c_double, vp]
    lib.ucnerf_grid_table_stats.argtypes = [vp, vp, vp, u32, u32, C.c_double, vp, vp]
    lib.ucnerf_pooled_encode_forward.argtypes = [vp, vp, u32, u32, C.c_int, vp, vp, vp, u32, u32, C.c_float, u32, vp, vp, vp]
    lib.ucnerf_pooled_encode_backward.argtypes = [vp, vp, vp, u32, u32, C.c_int, vp, vp, u32, u32, C.c_float, u32, vp, vp]
    lib.ucnerf_resample_intervals.argtypes = [vp, vp, u32, i32, C.c_int, f32, f32, f32, i32, vp, vp, i32, vp, vp]
    lib.ucnerf_composite_train_forward.argtypes = [vp, vp, vp, vp, u32, i32, f32, vp, vp, vp, vp]
    lib.ucnerf_composite_train_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, u32, i32, f32, vp, vp, vp]
    lib.ucnerf_cast_rays.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, u32, i32, f32, vp, vp, vp, vp]
    lib.ucnerf_sky_create.argtypes = [C.POINTER(SkyDesc), C.POINTER(vp)]
    lib.ucnerf_sky_destroy.argtypes = [vp]
    lib.ucnerf_sky_render.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, C.c_double, vp, vp]
    lib.ucnerf_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.ucnerf_get_timing.argtypes = [vp, c_float_p, C.POINTER(C.c_uint32), C.c_int]
    lib.ucnerf_debug_u_grid.argtypes = [C.c_int, vp]
    lib.ucnerf_debug_cone_table.argtypes = [vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("ucnerf_abi_version",):
            pass
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().ucnerf_last_error()
        raise UcnerfError(f"{what}: {msg.decode() if msg else 'error'} (code {rc})")


def launch_count():
    return int(load().ucnerf_launch_count())

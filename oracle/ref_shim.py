"""TEST INFRASTRUCTURE - import the *unmodified* reference Python hot path on CPU.

Resolves the reference tree from /root/reference (build container) or from the byte-identical staged copy
baseline/_ref/nerf (baseline/stage_ref.py; git-ignored, shipped to the GPU box).  Used by oracle/make_golden.py to
pin oracle/ucnerf_oracle.py and to generate tests/golden/*.npz, by tests/test_gpu_reference_modules.py (the reference's
own modules on the drop-in kernels) and by bench.py's reference legs.  The product package never imports it.

The reference needs 12 third-party modules that are absent here and a native `_gridencoder`
backend that has no CPU implementation (gridencoder.cu:L15,L449-452).  We install inert stubs
for the former in sys.modules and, for the latter, a stand-in whose forward is the oracle's
restatement of kernel_grid (that restatement is pinned separately against the reference CUDA
kernel on the GPU box, see oracle/build_ref.py).  Everything else - internal/models.py,
render.py, stepfun.py, coord.py, math.py, gridencoder/grid.py - is the reference's own code.
"""
import contextlib
import importlib
import os
import sys
import types

import numpy as np
import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = "/root/reference/nerf"
if not os.path.isdir(os.path.join(REF_ROOT, "internal")):
    REF_ROOT = os.path.join(_REPO, "baseline", "_ref", "nerf")
REF_CUDA_SO = os.path.join(_REPO, "oracle", "_ref", "_gridencoder_ref.so")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "internal"))


class _StubModule(types.ModuleType):
    """Module whose missing attributes resolve to inert objects (only defined names matter)."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything()


def _stub(name, **attrs):
    m = _StubModule(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, k):
        return _Anything()


def _install_stubs():
    def configurable(*args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return args[0]
        return lambda f: f

    gin = _stub("gin", configurable=configurable, add_config_file_search_path=lambda *a, **k: None,
                parse_config_files_and_bindings=lambda *a, **k: None, config_scope=contextlib.nullcontext)
    gin.config = _stub("gin.config", external_configurable=lambda f, module=None: f)

    class Accelerator:
        process_index = 0
        num_processes = 1
        is_main_process = True

        def autocast(self):
            return contextlib.nullcontext()

        def gather(self, v):
            return v

    _stub("accelerate", Accelerator=Accelerator)

    def segment_coo(src, index, out=None, dim_size=None, reduce="sum"):
        n = out.shape[0] if out is not None else int(index.max()) + 1
        res = torch.zeros((n,) + src.shape[1:], dtype=src.dtype).index_add_(0, index, src)
        if reduce == "mean":
            cnt = torch.zeros(n, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
            res = res / cnt.clamp_min(1).reshape((n,) + (1,) * (src.dim() - 1))
        return res

    _stub("torch_scatter", segment_coo=segment_coo)
    mpl = _stub("matplotlib", cm=_Anything(), use=lambda *a, **k: None)
    sys.modules["matplotlib.cm"] = _stub("matplotlib.cm", get_cmap=_Anything())
    sys.modules["matplotlib.pyplot"] = _stub("matplotlib.pyplot")
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    _stub("mpl_toolkits")
    _stub("mpl_toolkits.mplot3d", Axes3D=_Anything)
    sk = _stub("skimage")
    sk.metrics = _stub("skimage.metrics", structural_similarity=_Anything())
    _stub("lpips", LPIPS=_Anything)
    ns = _stub("nuscenes")
    ns.nuscenes = _stub("nuscenes.nuscenes", NuScenes=_Anything)
    _stub("pyquaternion", Quaternion=_Anything)
    _stub("rawpy")
    _stub("tensorboardX", SummaryWriter=_Anything)
    _stub("imageio")
    _stub("mediapy")
    _stub("trimesh")
    _stub("pycolmap", SceneManager=_Anything)
    for name in ("cv2", "PIL", "scipy"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)


def _install_grid_backend():
    from oracle import ucnerf_oracle as O

    def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype,
                            align_corners, interp):
        out, dd = O.grid_encode_forward(inputs.detach().numpy(), embeddings.detach().numpy(), offsets.numpy(),
                                        B, D, C, L, S, H, dy_dx is not None, gridtype, align_corners, interp)
        outputs.copy_(torch.from_numpy(out))
        if dy_dx is not None:
            dy_dx.copy_(torch.from_numpy(dd))

    def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx,
                             grad_inputs, gridtype, align_corners, interp):
        ge, gi = O.grid_encode_backward(grad.numpy(), inputs.numpy(), embeddings.detach().numpy(), offsets.numpy(),
                                        B, D, C, L, S, H, None if dy_dx is None else dy_dx.numpy(),
                                        gridtype, align_corners, interp)
        grad_embeddings.copy_(torch.from_numpy(ge))
        if grad_inputs is not None:
            grad_inputs.copy_(torch.from_numpy(gi))

    def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
        out = O.grad_total_variation(inputs.numpy(), embeddings.detach().numpy(), grad.numpy(), offsets.numpy(),
                                     weight, B, D, C, L, S, H, gridtype, align_corners)
        grad.copy_(torch.from_numpy(out))

    return _stub("_gridencoder", grid_encode_forward=grid_encode_forward, grid_encode_backward=grid_encode_backward,
                 grad_total_variation=grad_total_variation)


_BACKENDS = {}


def grid_backend(kind):
    """The module the reference's grid.py binds as `_backend` (gridencoder/grid.py:L9-12):
      "oracle"   - CPU stand-in (the oracle's numpy restatement of kernel_grid; the reference has no CPU kernel),
      "oracle_c" - CPU stand-in for timing: forward through the C + OpenMP restatement oracle/grid_cpu.c (bit-identical
                   to "oracle", tests/test_oracle_grid_c.py), backward / TV / dy_dx through the numpy one,
      "ref_cuda" - the reference's own gridencoder.cu compiled for sm_100a (oracle/build_ref.py),
      "dropin"   - ucnerf_b200/dropin/_gridencoder.py, i.e. the product kernels behind the reference's module name."""
    if kind in _BACKENDS:
        return _BACKENDS[kind]
    if kind == "oracle":
        prev = sys.modules.get("_gridencoder")
        m = _install_grid_backend()
        if prev is not None:
            sys.modules["_gridencoder"] = prev
    elif kind == "oracle_c":
        from oracle import build_c
        base = grid_backend("oracle")

        def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype,
                                align_corners, interp):
            if dy_dx is not None or embeddings.dtype != torch.float32:
                return base.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype,
                                                align_corners, interp)
            x, e, o = inputs.detach().contiguous(), embeddings.detach().contiguous(), offsets.to(torch.int32).contiguous()
            assert outputs.is_contiguous() and outputs.dtype == torch.float32
            rc = build_c.load().ucnerf_oracle_grid_forward_f32(x.data_ptr(), e.data_ptr(), o.data_ptr(), outputs.data_ptr(),
                                                               B, D, C, L, float(S), int(H), int(gridtype),
                                                               int(bool(align_corners)), int(interp))
            if rc != 0:
                raise RuntimeError("grid_cpu.c: unsupported input dimension")

        m = _StubModule("_gridencoder_oracle_c")
        m.__dict__.update(grid_encode_forward=grid_encode_forward, grid_encode_backward=base.grid_encode_backward,
                          grad_total_variation=base.grad_total_variation)
    elif kind == "ref_cuda":
        import importlib.util
        if not os.path.exists(REF_CUDA_SO):
            raise RuntimeError("oracle/_ref/_gridencoder_ref.so not built (oracle/build_ref.py)")
        spec = importlib.util.spec_from_file_location("_gridencoder_ref", REF_CUDA_SO)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    elif kind == "dropin":
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            "_gridencoder_dropin", os.path.join(_REPO, "ucnerf_b200", "dropin", "_gridencoder.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    else:
        raise KeyError(kind)
    _BACKENDS[kind] = m
    return m


def use_grid_backend(kind):
    """Re-bind the (already imported, unmodified) reference grid.py to another native backend."""
    R = load_reference()
    R.grid._backend = grid_backend(kind)
    return R


_MODELS = None


def load_reference(backend="oracle"):
    """Returns the reference modules (models, configs, render, stepfun, coord, math, grid); `backend` = what the
    reference's `import _gridencoder` resolves to at first import (see grid_backend / use_grid_backend)."""
    global _MODELS
    if _MODELS is not None:
        return _MODELS
    if not available():
        raise RuntimeError("reference tree not present")
    _install_stubs()
    sys.modules["_gridencoder"] = grid_backend(backend)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from internal import models, configs, render, stepfun, coord, math as rmath
        from gridencoder import grid
    _MODELS = types.SimpleNamespace(models=models, configs=configs, render=render, stepfun=stepfun,
                                    coord=coord, math=rmath, grid=grid)
    return _MODELS


def build_reference_model(cfg, params):
    """Instantiate the reference Model with `cfg`'s hyper-parameters (set as class attributes, which is
    what gin does through ctor kwargs, models.py:L25-27) and load `params` (reference state_dict names)."""
    R = load_reference()
    M = R.models
    M.Model.num_levels = cfg.num_levels
    M.Model.num_prop_samples = cfg.num_prop_samples
    M.Model.num_nerf_samples = cfg.num_nerf_samples
    M.Model.opaque_background = False
    M.Model.prop_desired_grid_size = [g.desired_resolution for g in cfg.prop_grids]
    M.Model.dilation_multiplier = cfg.dilation_multiplier
    M.Model.dilation_bias = cfg.dilation_bias
    for cls, gs in ((M.PropMLP, cfg.prop_grids[0]), (M.NerfMLP, cfg.nerf_grid)):
        cls.disable_density_normals = True
        cls.grid_log2_hashmap_size = gs.log2_hashmap_size
        cls.grid_level_dim = gs.level_dim
        cls.grid_base_resolution = gs.base_resolution
        cls.bottleneck_width = cfg.bottleneck_width
        cls.net_width_viewdirs = cfg.net_width_viewdirs
    M.PropMLP.disable_rgb = True
    M.NerfMLP.disable_rgb = False
    M.NerfMLP.grid_disired_resolution = cfg.nerf_grid.desired_resolution
    conf = R.configs.Config()
    model = M.Model(config=conf)
    sd = model.state_dict()
    new = {}
    for k, v in sd.items():
        if k in params:
            assert tuple(params[k].shape) == tuple(v.shape), (k, params[k].shape, v.shape)
            new[k] = params[k].to(v.dtype)
        else:
            assert k.endswith('.idx'), k  # level-id buffer, only used by training losses
            new[k] = v
    model.load_state_dict(new)
    model.eval()
    return model, conf


@contextlib.contextmanager
def inject_rand_vec(rand_vec):
    """Make the reference's `torch.randn_like(cam_dirs)` (render.py:L140) return `rand_vec`."""
    orig = torch.randn_like

    def patched(t, *a, **k):
        if tuple(t.shape) == tuple(rand_vec.shape):
            return rand_vec.clone()
        return orig(t, *a, **k)

    torch.randn_like = patched
    try:
        yield
    finally:
        torch.randn_like = orig


@contextlib.contextmanager
def inject_rand_vec_rows(cam_dirs_full, rand_vec_full):
    """Same for a chunked render (`render_image` slices the flat batch into views, models.py:L939-953): a call
    `randn_like(chunk_of_cam_dirs)` gets the rows of `rand_vec_full` that the chunk occupies in `cam_dirs_full`."""
    orig = torch.randn_like
    base, nbytes = cam_dirs_full.data_ptr(), cam_dirs_full.numel() * cam_dirs_full.element_size()
    row = cam_dirs_full.shape[-1] * cam_dirs_full.element_size()
    flat = rand_vec_full.reshape(-1, rand_vec_full.shape[-1])

    def patched(t, *a, **k):
        off = t.data_ptr() - base
        if t.dim() == 2 and t.shape[-1] == flat.shape[-1] and 0 <= off < nbytes and off % row == 0 and t.is_contiguous():
            r0 = off // row
            return flat[r0:r0 + t.shape[0]].to(t.device).clone()
        return orig(t, *a, **k)

    torch.randn_like = patched
    try:
        yield
    finally:
        torch.randn_like = orig


@torch.no_grad()
def reference_forward(cfg, params, batch, compute_extras=True):
    model, conf = build_reference_model(cfg, params)
    b = {k: v for k, v in batch.items() if k != 'rand_vec'}
    with inject_rand_vec(batch['rand_vec']):
        renderings, ray_history = model(False, b, train_frac=1.0, compute_extras=compute_extras, zero_glo=True)
    return renderings, ray_history, model, conf

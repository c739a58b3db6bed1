"""Host-side mirror of the reference's render surface for the eval hot path.

  * `HotPathModel`    - owns a ucnerf_model handle built from a reference state_dict (same key names,
                        SURVEY.md section 5 "checkpoint") and exposes
                          .forward(rand, batch, train_frac, compute_extras, ...) -> (renderings, ray_history)
                        with the contract of `Model.forward` (internal/models.py:L97-365, eval path), and
                          .render_rays / .render_rays_host  (flat tensors in, dict of tensors out).
  * `render_image`    - drop-in for `models.render_image` (internal/models.py:L907-1007): same signature, same
                        keys in the returned dict.  Instead of slicing every 15k-ray chunk by rank and calling
                        `accelerator.gather` on every leaf, each rank renders ONE contiguous tile of the image and
                        the ranks exchange ONE packed [rays, 12] buffer with a single NCCL all-gather.

PyTorch is used for device memory, streams and torch.distributed only; all arithmetic of the path runs in
libucnerf_b200.so.  There is no fallback: without the library or without a CUDA device these raise."""
import ctypes as C
import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib

PACKED_WIDTH = 12
PACKED_FIELDS = ("rgb", "depth", "acc", "distance_mean", "distance_median", "distance_percentile_5",
                 "distance_percentile_95", "depth_raw")
_RAY_KEYS = ("origins", "directions", "viewdirs", "cam_dirs", "radii", "near", "far")


def _grid_num_levels(desired, base=16, interval=2):
    # internal/models.py:L425-426
    return int(np.log(desired / base) / np.log(interval)) + 1


def make_camera(pixtocam, camtoworld, width: int, height: int, near: float, far: float, rand_seed: int = 0):
    """ucnerf_camera from the arrays the reference's Dataset holds (`pixtocams[i]` [3,3], `camtoworlds[i]` [3,4] or
    [4,4], `width`, `height`, `near`, `far`; internal/datasets.py:L296-348)."""
    cam = _lib.Camera()
    p = np.asarray(pixtocam, dtype=np.float64).reshape(3, 3)
    c = np.asarray(camtoworld, dtype=np.float64)[:3, :4]
    cam.pixtocam[:] = p.reshape(-1).tolist()
    cam.camtoworld[:] = np.ascontiguousarray(c).reshape(-1).tolist()
    cam.width, cam.height, cam.near, cam.far, cam.rand_seed = int(width), int(height), float(near), float(far), int(rand_seed)
    return cam


def generate_rays(pixtocam, camtoworld, width, height, near, far, rows=None, rand_seed=0, device=None,
                  with_rand_vec=True) -> Dict[str, torch.Tensor]:
    """GPU replacement of the eval loader's numpy ray generation (camera_utils.pixels_to_rays / cast_pinhole_rays,
    internal/camera_utils.py:L448-557,L611-632 + Dataset._make_ray_batch, datasets.py:L386-476) for one perspective
    camera without lens distortion: returns the reference's ray-dict keys as flat CUDA tensors for image rows
    `rows=(row0, n_rows)` (default: the whole image), row-major pixel order."""
    if not torch.cuda.is_available():
        raise _lib.UcnerfError("ucnerf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    lib = _lib.load()
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    row0, n_rows = rows if rows is not None else (0, height)
    n = n_rows * width
    cam = make_camera(pixtocam, camtoworld, width, height, near, far, rand_seed)
    shapes = {"origins": 3, "directions": 3, "viewdirs": 3, "cam_dirs": 3, "radii": 1, "near": 1, "far": 1,
              "imageplane": 2}
    if with_rand_vec:
        shapes["rand_vec"] = 3
    out = {k: torch.empty((n, w), device=dev, dtype=torch.float32) for k, w in shapes.items()}
    rb = _lib.RayBuffers()
    for k, t in out.items():
        setattr(rb, k, t.data_ptr())
    with torch.cuda.device(dev):
        _lib.check(lib.ucnerf_generate_rays(C.byref(cam), row0, n_rows, C.byref(rb),
                                            torch.cuda.current_stream().cuda_stream), "generate_rays")
    return out


class HotPathModel:
    """B200 renderer for one UC-NeRF `Model` (proposal MLPs + NeRF MLP, waymo.gin-style)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], *, num_prop_samples: int, num_nerf_samples: int,
                 num_prop_levels: Optional[int] = None, bottleneck_width: int = 256, net_width_viewdirs: int = 256,
                 deg_view: int = 4, base_resolution: int = 16, dilation_multiplier: float = 0.5,
                 dilation_bias: float = 0.0025, anneal_slope: float = 10.0, resample_padding: float = 0.0,
                 std_scale: float = 0.5, bg_intensity: float = 1.0, density_bias: float = -1.0,
                 rgb_padding: float = 0.001, vis_num_rays: int = 16, per_level_scales: Optional[Dict[str, float]] = None,
                 device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.UcnerfError("ucnerf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if num_prop_levels is None:
            num_prop_levels = len({k.split('.')[0] for k in state_dict if k.startswith('prop_mlp_')})
        self.num_prop_levels = num_prop_levels
        self.num_levels = num_prop_levels + 1
        self.num_prop_samples, self.num_nerf_samples = num_prop_samples, num_nerf_samples
        self.samples = [num_prop_samples] * num_prop_levels + [num_nerf_samples]
        self.vis_num_rays = vis_num_rays
        self._pls = dict(per_level_scales or {})
        self._hyper = dict(bottleneck_width=bottleneck_width, net_width_viewdirs=net_width_viewdirs, deg_view=deg_view,
                           base_resolution=base_resolution, dilation_multiplier=dilation_multiplier,
                           dilation_bias=dilation_bias, anneal_slope=anneal_slope, resample_padding=resample_padding,
                           std_scale=std_scale, bg_intensity=bg_intensity, density_bias=density_bias,
                           rgb_padding=rgb_padding)
        self._keep = {}      # device tensors the handle points into (embeddings) or was built from
        self._handle = C.c_void_p()
        self._build(state_dict, create=True)

    # ---- construction -------------------------------------------------------------------------
    @classmethod
    def from_reference_model(cls, model, config=None, device=None):
        """Build from a live reference `internal.models.Model` (reads its attributes + state_dict)."""
        m = model.module if hasattr(model, "module") else model
        nerf = m.nerf_mlp
        if getattr(m, "raydist_fn", None) is not None:
            raise NotImplementedError("fused path implements raydist_fn=None (waymo.gin) only")
        if getattr(m, "num_glo_features", 0) > 0:
            raise NotImplementedError("fused path does not implement GLO features")
        if not (nerf.disable_density_normals and not nerf.disable_rgb):
            raise NotImplementedError("fused path needs NerfMLP.disable_density_normals=True, disable_rgb=False")
        bg = m.bg_intensity_range
        if bg[0] != bg[1]:
            raise NotImplementedError("fused path needs a constant background (bg_intensity_range min == max)")
        sd = {k: v for k, v in m.state_dict().items()}
        pls = {'nerf_mlp': float(nerf.encoder.per_level_scale)}
        for i in range(m.num_levels - 1):
            pls[f'prop_mlp_{i}'] = float(m.get_submodule(f'prop_mlp_{i}').encoder.per_level_scale)
        return cls(sd, num_prop_samples=m.num_prop_samples, num_nerf_samples=m.num_nerf_samples,
                   num_prop_levels=m.num_levels - 1, bottleneck_width=nerf.bottleneck_width,
                   net_width_viewdirs=nerf.net_width_viewdirs, deg_view=nerf.deg_view,
                   base_resolution=nerf.grid_base_resolution, dilation_multiplier=m.dilation_multiplier,
                   dilation_bias=m.dilation_bias, anneal_slope=m.anneal_slope, resample_padding=m.resample_padding,
                   std_scale=m.std_scale, bg_intensity=float(bg[0]), density_bias=nerf.density_bias,
                   rgb_padding=nerf.rgb_padding,
                   vis_num_rays=getattr(config, "vis_num_rays", 16) if config is not None else 16,
                   per_level_scales=pls, device=device)

    def _dev(self, t):
        return t.detach().to(self.device, torch.float32).contiguous()

    def _mlp_desc(self, sd, prefix, desc):
        emb = self._dev(sd[prefix + '.encoder.embeddings'])
        offsets = sd[prefix + '.encoder.offsets'].detach().cpu().to(torch.int32).contiguous()
        grid_sizes = sd[prefix + '.encoder.grid_sizes'].detach().cpu().to(torch.int32).contiguous()
        L = offsets.numel() - 1
        C_ = emb.shape[1]
        h = self._hyper
        # per_level_scale as GridEncoder derives it (gridencoder/grid.py:L103-104) from the finest python-side
        # resolution: grid_sizes[-1] - 1 == ceil(H * s^(L-1)); for the power-of-two grids of models.py this is exact.
        desired = int(grid_sizes[-1].item()) - 1
        pls = np.exp2(np.log2(desired / h['base_resolution']) / (L - 1)) if L > 1 else 2.0
        pls = self._pls.get(prefix, pls)  # exact value when built from a live GridEncoder
        w0, b0 = self._dev(sd[prefix + '.density_layer.0.weight']), self._dev(sd[prefix + '.density_layer.0.bias'])
        w2, b2 = self._dev(sd[prefix + '.density_layer.2.weight']), self._dev(sd[prefix + '.density_layer.2.bias'])
        self._keep[prefix] = (emb, offsets, grid_sizes, w0, b0, w2, b2)
        desc.embeddings = emb.data_ptr()
        desc.offsets_host = offsets.data_ptr()
        desc.grid_sizes_host = grid_sizes.data_ptr()
        desc.grid_levels, desc.level_dim = L, C_
        desc.base_resolution = h['base_resolution']
        desc.log2_per_level_scale = float(np.log2(pls))
        desc.density0_w, desc.density0_b = w0.data_ptr(), b0.data_ptr()
        desc.density2_w, desc.density2_b = w2.data_ptr(), b2.data_ptr()
        if w0.shape != (64, L * C_):
            raise ValueError(f"{prefix}.density_layer.0.weight has shape {tuple(w0.shape)}, expected (64, {L * C_})")

    def _build(self, sd, create):
        h = self._hyper
        d = _lib.ModelDesc()
        d.num_prop_levels = self.num_prop_levels
        d.num_prop_samples, d.num_nerf_samples = self.num_prop_samples, self.num_nerf_samples
        d.bottleneck_width, d.net_width_viewdirs, d.deg_view = h['bottleneck_width'], h['net_width_viewdirs'], h['deg_view']
        for k in ("dilation_multiplier", "dilation_bias", "anneal_slope", "resample_padding", "std_scale",
                  "bg_intensity", "density_bias", "rgb_padding"):
            setattr(d, k, float(h[k]))
        for i in range(self.num_prop_levels):
            self._mlp_desc(sd, f'prop_mlp_{i}', d.prop[i])
        self._mlp_desc(sd, 'nerf_mlp', d.nerf)
        names = ('lin_second_stage_0.weight', 'lin_second_stage_0.bias', 'lin_second_stage_1.weight',
                 'lin_second_stage_1.bias', 'rgb_layer.weight', 'rgb_layer.bias')
        ts = [self._dev(sd['nerf_mlp.' + n]) for n in names]
        self._keep['view'] = ts
        (d.view0_w, d.view0_b, d.view1_w, d.view1_b, d.rgb_w, d.rgb_b) = [t.data_ptr() for t in ts]
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if create:
                _lib.check(self.lib.ucnerf_model_create(C.byref(d), C.byref(self._handle)), "model_create")
            else:
                _lib.check(self.lib.ucnerf_model_refresh(self._handle, C.byref(d),
                                                         torch.cuda.current_stream().cuda_stream), "model_refresh")

    def refresh(self, state_dict):
        """Re-read the weights (e.g. after an optimiser step)."""
        self._build(state_dict, create=False)

    def set_option(self, key: str, value: int):
        _lib.check(self.lib.ucnerf_set_option(self._handle, key.encode(), int(value)), "set_option")

    def set_rgb_affine(self, affine):
        """Brightness-correction affine of the image being rendered ([3,4] or [1,3,4] tensor / array, e.g.
        `model.brightness_corr(indices=cam_idx)[0]`, extrinsic_optimizer.py:L15-25), applied to the final rgb inside the
        compositing kernel as models.py:L349 does; None switches it off."""
        if affine is None:
            _lib.check(self.lib.ucnerf_set_rgb_affine(self._handle, None), "set_rgb_affine")
            return
        a = np.ascontiguousarray(torch.as_tensor(affine).detach().cpu().to(torch.float32).reshape(3, 4).numpy())
        _lib.check(self.lib.ucnerf_set_rgb_affine(self._handle, a.ctypes.data), "set_rgb_affine")

    def timing(self, reset=True):
        """Per kernel family: (device milliseconds, launches) accumulated since the last reset; needs
        set_option("timing", 1)."""
        buf, cnt = (C.c_float * 5)(), (C.c_uint32 * 5)()
        _lib.check(self.lib.ucnerf_get_timing(self._handle, buf, cnt, int(reset)), "get_timing")
        names = ("resample", "encode_prop", "encode_nerf", "color_mlp", "composite")
        return {n: (float(buf[i]), int(cnt[i])) for i, n in enumerate(names)}

    def close(self):
        if self._handle:
            self.lib.ucnerf_model_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- rendering ----------------------------------------------------------------------------
    def _out_spec(self, n, want):
        S = self.samples
        spec = {"rgb": (n, 3), "depth": (n,), "depth_raw": (n,), "acc": (n,), "distance_mean": (n,),
                "distance_median": (n,), "distance_percentile_5": (n,), "distance_percentile_95": (n,),
                "sample_rgb": (n, S[-1], 3), "sample_density": (n, S[-1]), "packed": (n, PACKED_WIDTH),
                "sample_coord": (n, S[-1], 3)}
        for l in range(self.num_levels):
            spec[f"sdist_{l}"] = (n, S[l] + 1)
            spec[f"weights_{l}"] = (n, S[l])
        unknown = set(want) - set(spec)
        if unknown:
            raise KeyError(f"unknown outputs {sorted(unknown)}")
        return {k: spec[k] for k in want}

    @staticmethod
    def _fill_struct(o, bufs):
        for k, t in bufs.items():
            if k.startswith("sdist_"):
                o.sdist[int(k[6:])] = t.data_ptr()
            elif k.startswith("weights_"):
                o.weights[int(k[8:])] = t.data_ptr()
            else:
                setattr(o, k, t.data_ptr())

    def render_rays(self, batch: Dict[str, torch.Tensor], train_frac: float = 1.0, rand_vec=None,
                    want=("rgb", "depth", "acc")) -> Dict[str, torch.Tensor]:
        """batch: flat CUDA tensors origins/directions/viewdirs/cam_dirs [N,3], radii/near/far [N,1] or [N]."""
        n = batch['origins'].shape[0]
        rays = _lib.Rays()
        keep = []
        for k in _RAY_KEYS:
            t = batch[k]
            if t.device != self.device:
                raise ValueError(f"batch['{k}'] is on {t.device}, renderer on {self.device}")
            t = t.detach().to(torch.float32).contiguous()
            keep.append(t)
            setattr(rays, k, t.data_ptr())
        if rand_vec is None:
            rand_vec = batch.get('rand_vec')
        if rand_vec is None:  # render.py:L140 draws it at every call
            rand_vec = torch.randn_like(keep[3])
        rand_vec = rand_vec.detach().to(self.device, torch.float32).contiguous()
        rays.rand_vec = rand_vec.data_ptr()
        spec = self._out_spec(n, want)
        bufs = {k: torch.empty(shape, device=self.device, dtype=torch.float32) for k, shape in spec.items()}
        o = _lib.Outputs()
        self._fill_struct(o, bufs)
        with torch.cuda.device(self.device):
            rc = self.lib.ucnerf_render_rays(self._handle, n, C.byref(rays), float(train_frac), C.byref(o),
                                             torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "render_rays")
        return bufs

    def render_rays_host(self, batch: Dict[str, torch.Tensor], train_frac: float = 1.0,
                         want=("rgb", "depth", "acc"), out: Optional[Dict[str, torch.Tensor]] = None):
        """Same with HOST tensors (pinned recommended): H2D, render, D2H inside the library call."""
        n = batch['origins'].shape[0]
        rays = _lib.Rays()
        keep = []
        for k in _RAY_KEYS + ("rand_vec",):
            t = batch[k]
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"batch['{k}'] must be a contiguous float32 CPU tensor")
            keep.append(t)
            setattr(rays, k, t.data_ptr())
        spec = self._out_spec(n, want)
        if out is None:
            out = {k: torch.empty(shape, dtype=torch.float32, pin_memory=True) for k, shape in spec.items()}
        o = _lib.Outputs()
        self._fill_struct(o, out)
        with torch.cuda.device(self.device):
            rc = self.lib.ucnerf_render_rays_host(self._handle, n, C.byref(rays), float(train_frac), C.byref(o),
                                                  torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "render_rays_host")
        return out

    def render_camera(self, pixtocam, camtoworld, width, height, near, far, rows=None, train_frac: float = 1.0,
                      rand_seed: int = 0, want=("rgb", "depth", "acc"), host_out: Optional[Dict[str, torch.Tensor]] = None):
        """Render image rows `rows=(row0, n_rows)` (default all) of one pinhole camera: the ray batch is generated on
        the GPU from the camera parameters (ucnerf_render_camera).  With `host_out` (dict of pinned CPU tensors, or
        True to allocate them) the results are copied back inside the call (ucnerf_render_camera_host)."""
        row0, n_rows = rows if rows is not None else (0, height)
        n = n_rows * width
        cam = make_camera(pixtocam, camtoworld, width, height, near, far, rand_seed)
        spec = self._out_spec(n, want)
        o = _lib.Outputs()
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            if host_out is not None and host_out is not False:
                if host_out is True:
                    host_out = {k: torch.empty(shape, dtype=torch.float32, pin_memory=True) for k, shape in spec.items()}
                self._fill_struct(o, host_out)
                rc = self.lib.ucnerf_render_camera_host(self._handle, C.byref(cam), row0, n_rows, float(train_frac),
                                                        C.byref(o), st)
                _lib.check(rc, "render_camera_host")
                return host_out
            bufs = {k: torch.empty(shape, device=self.device, dtype=torch.float32) for k, shape in spec.items()}
            self._fill_struct(o, bufs)
            rc = self.lib.ucnerf_render_camera(self._handle, C.byref(cam), row0, n_rows, float(train_frac), C.byref(o), st)
        _lib.check(rc, "render_camera")
        return bufs

    def forward(self, rand, batch, train_frac, compute_extras, zero_glo=True, eval_camidx=None, rand_vec=None):
        """`Model.forward` contract for the eval path (internal/models.py:L97-365, heads excluded):
        returns (renderings, ray_history), one entry per sampling level."""
        if rand:
            raise NotImplementedError("the fused path implements the deterministic eval path (rand=False); for the training "
                                      "step use ucnerf_b200.train_forward.level_loop on the torch model")
        lead = batch['origins'].shape[:-1]
        flat = {k: batch[k].reshape(-1, batch[k].shape[-1]) for k in _RAY_KEYS}
        if rand_vec is None and batch.get('rand_vec') is not None:
            rand_vec = batch['rand_vec'].reshape(-1, 3)
        want = ["rgb", "depth", "acc", "sample_rgb", "sample_density", "sample_coord"]
        if compute_extras:
            want += ["distance_mean", "distance_median", "distance_percentile_5", "distance_percentile_95"]
        for l in range(self.num_levels):
            want += [f"sdist_{l}", f"weights_{l}"]
        out = self.render_rays(flat, train_frac, rand_vec, want)
        renderings, history = [], []
        nv = self.vis_num_rays
        for l in range(self.num_levels):
            last = l == self.num_levels - 1
            r = {"weights": out[f"weights_{l}"].reshape(lead + (-1,))}
            if last:
                for k in ("rgb", "depth", "acc", "distance_mean", "distance_median", "distance_percentile_5",
                          "distance_percentile_95"):
                    if k in out:
                        r[k] = out[k].reshape(lead + out[k].shape[1:])
            if compute_extras:
                r['ray_sdist'] = out[f"sdist_{l}"][:nv]
                r['ray_weights'] = out[f"weights_{l}"][:nv]
                if last:
                    r['ray_rgbs'] = out["sample_rgb"][:nv]
            renderings.append(r)
            h = {"sdist": out[f"sdist_{l}"].reshape(lead + (-1,)), "weights": r["weights"]}
            if last:
                h["rgb"] = out["sample_rgb"].reshape(lead + out["sample_rgb"].shape[1:])
                h["density"] = out["sample_density"].reshape(lead + (-1,))
                h["coord"] = out["sample_coord"].reshape(lead + out["sample_coord"].shape[1:])
            history.append(h)
        if compute_extras:  # models.py:L313-324: proposal levels show the final average colour
            final_rgb = torch.sum(renderings[-1]['ray_rgbs'] * renderings[-1]['ray_weights'][..., None], dim=-2)
            for l in range(self.num_levels - 1):
                S = self.samples[l]
                renderings[l]['ray_rgbs'] = final_rgb[:, None, :].expand(-1, S, -1)
        return renderings, history


class SkyHead:
    """Tensor-core sky head (ucnerf_sky_*): the reference's `render_rays(ray_batch, network_fn=model.skynerf)`
    (internal/models.py:L326-337, L743-904) for the architecture `Model` builds (D=8, W=256, raw-xyz input, skip after
    layer 4, 4-frequency view embedding, 120 samples).  Built from a state_dict with the reference key names
    (`skynerf.pts_linears.N.weight` ...)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], prefix: str = "skynerf", n_samples: int = 120, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.UcnerfError("ucnerf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        g = lambda k: state_dict[f"{prefix}.{k}"].detach().to(self.device, torch.float32).contiguous()
        shapes = {"pts_linears.0.weight": (256, 3), "pts_linears.5.weight": (256, 259), "views_linears.0.weight": (128, 283),
                  "feature_linear.weight": (256, 256), "alpha_linear.weight": (1, 256), "rgb_linear.weight": (3, 128)}
        for k, shp in shapes.items():
            if tuple(state_dict[f"{prefix}.{k}"].shape) != shp:
                raise NotImplementedError(f"sky head: {prefix}.{k} has shape {tuple(state_dict[f'{prefix}.{k}'].shape)}, "
                                          f"the tensor-core kernel implements the reference architecture {shp}")
        d = _lib.SkyDesc()
        keep = []
        for i in range(8):
            w, b = g(f"pts_linears.{i}.weight"), g(f"pts_linears.{i}.bias")
            keep += [w, b]
            d.pts_w[i], d.pts_b[i] = w.data_ptr(), b.data_ptr()
        for name, fw, fb in (("feature_linear", "feature_w", "feature_b"), ("alpha_linear", "alpha_w", "alpha_b"),
                             ("views_linears.0", "views_w", "views_b"), ("rgb_linear", "rgb_w", "rgb_b")):
            w, b = g(name + ".weight"), g(name + ".bias")
            keep += [w, b]
            setattr(d, fw, w.data_ptr())
            setattr(d, fb, b.data_ptr())
        d.n_samples = n_samples
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            _lib.check(self.lib.ucnerf_sky_create(C.byref(d), C.byref(self._handle)), "sky_create")
        del keep

    def render(self, origins, directions, far, views) -> torch.Tensor:
        """[N,3] origins / directions / views (= the batch's cam_dirs), far [N] or [N,1] -> sky rgb_map [N,3]."""
        n = origins.shape[0]
        t = [x.detach().to(self.device, torch.float32).contiguous() for x in (origins, directions, far.reshape(-1), views)]
        out = torch.empty((n, 3), device=self.device, dtype=torch.float32)
        if n == 0:
            return out
        sky_far = float(t[2][0]) * 1.5        # models.py:L329 (the same host read the reference does)
        with torch.cuda.device(self.device):
            rc = self.lib.ucnerf_sky_render(self._handle, n, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(),
                                            sky_far, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "sky_render")
        return out

    def close(self):
        if self._handle:
            self.lib.ucnerf_sky_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_bounds(num_rays: int, world: int, rank: int):
    """Contiguous tile [start, stop) of rank `rank`; every rank renders ceil(num_rays / world) rays (the last
    tiles are padded by re-rendering the final ray) so one fixed-size all-gather moves the image."""
    per = (num_rays + world - 1) // world
    start = min(rank * per, num_rays)
    stop = min(start + per, num_rays)
    return per, start, stop


def gather_tiles(tile: torch.Tensor, world: int, num_rows: Optional[int] = None) -> torch.Tensor:
    """The ONE collective of a multi-GPU render: every rank contributes its `[per, W]` tile,
    every rank receives the `[world * per, W]` image (trimmed to `num_rows`, which drops the padding of the last tile).
    `render_image` and `bench.py --gpus N` both call this."""
    if world == 1:
        return tile if num_rows is None else tile[:num_rows]
    import torch.distributed as dist
    tile = tile.contiguous()
    full = torch.empty((world * tile.shape[0],) + tuple(tile.shape[1:]), device=tile.device, dtype=tile.dtype)
    dist.all_gather_into_tensor(full, tile)
    return full if num_rows is None else full[:num_rows]


def _peer_image(r, world, per, config):
    """The PeerImage (fused NVLink tile exchange, ucnerf_b200/peer.py) cached on the renderer, or None when it is switched
    off / not applicable (CPU stand-in renderer, non-NCCL backend) / not available (no peer access): then the plain
    all-gather path runs.  Creating it is collective, and so is the fallback decision."""
    if not getattr(config, "ucnerf_peer_exchange", True) or not hasattr(r, "_handle") or r.device.type != "cuda":
        return None
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_backend() != "nccl":
        return None
    cached = getattr(r, "_peer_image", None)
    if cached is not None and cached[0] == (world, per):
        return cached[1]
    from .peer import PeerImage
    if cached is not None and cached[1] is not None:
        cached[1].close()
    try:
        pi = PeerImage(world * per, device=r.device)
    except _lib.UcnerfError:
        pi = None
    r._peer_image = ((world, per), pi)
    return pi


def _param_versions(module):
    """Fingerprint of a module's parameters: in-place updates (optimiser steps, load_state_dict) bump `_version`,
    re-assigned storage changes `data_ptr`."""
    return tuple((p._version, p.data_ptr()) for p in module.parameters())


def _get_renderer(model, config):
    """The HotPathModel cached on the reference module.  ucnerf_model_create copies the MLP weights into library-owned
    memory, so a cached handle goes stale when training continues between two test renders (train.py:L330 calls
    render_image periodically) or after load_state_dict: the parameter versions are compared on every call and the
    handle is refreshed when any moved."""
    m = model.module if hasattr(model, "module") else model
    mods = [m.nerf_mlp] + [m.get_submodule(f'prop_mlp_{i}') for i in range(m.num_levels - 1)]
    versions = tuple(_param_versions(x) for x in mods)
    r = getattr(m, "_ucnerf_b200_renderer", None)
    if r is None:
        r = HotPathModel.from_reference_model(m, config)
        object.__setattr__(m, "_ucnerf_b200_renderer", r)
    elif versions != getattr(m, "_ucnerf_b200_renderer_versions", None):
        r.refresh({k: v for k, v in m.state_dict().items()})
    object.__setattr__(m, "_ucnerf_b200_renderer_versions", versions)
    return r


@torch.no_grad()
def render_image(model, accelerator, batch, rand, train_frac, config, verbose=True, return_weights=False,
                 eval_camidx=0, rand_vec=None, renderer: Optional[HotPathModel] = None):
    """Drop-in for internal/models.py:L907-1007 `render_image`: fused forward path; with `config.brightness_correction`
    the affine is applied in the compositing kernel, with `config.model_sky` the reference's own sky head runs on top
    (keys `sky_rgbs`, `affine_trans`, `affine_trans_sky` as models.py:L336-363).

    model: a reference `Model` (a HotPathModel is built from it, cached on the module and refreshed whenever the
    module's parameters changed since the previous call) or None when `renderer` is given.  accelerator: only process_index / num_processes are read."""
    if rand:
        raise NotImplementedError("render_image on the fused path is the deterministic eval path (rand=False)")
    r = renderer if renderer is not None else _get_renderer(model, config)
    height, width = batch['origins'].shape[:2]
    num_rays = height * width
    flat = {k: batch[k].reshape(num_rays, -1).to(r.device, non_blocking=True) for k in _RAY_KEYS}
    if rand_vec is None:
        rand_vec = batch.get('rand_vec')
    if rand_vec is not None:
        rand_vec = rand_vec.reshape(num_rays, 3).to(r.device, non_blocking=True)
    world = getattr(accelerator, "num_processes", 1) if accelerator is not None else 1
    rank = getattr(accelerator, "process_index", 0) if accelerator is not None else 0
    per, start, stop = shard_bounds(num_rays, world, rank)
    idx = torch.arange(start, start + per, device=r.device).clamp_(max=num_rays - 1)
    local = {k: v[idx] if (world > 1) else v for k, v in flat.items()}
    if rand_vec is None:
        # one draw for the whole image, identical on every rank would need a shared seed; the reference draws
        # per rank (render.py:L140 under set_seed(device_specific=True)), so a per-rank draw matches it.
        lrv = torch.randn_like(local['cam_dirs'])
    else:
        lrv = rand_vec[idx] if world > 1 else rand_vec
    # heads of the shipped Waymo configuration (scripts/train_waymo.sh: model_sky + brightness_correction).
    # Brightness (models.py:L339-363): ONE evaluation of the reference module per image, the affine is applied in the
    # compositing epilogue.  Sky (models.py:L326-337): the reference's own `render_rays` + `skynerf` run unchanged on
    # this rank's tile ("run unchanged on top"); its tensor-core kernel is §8f N1.
    affine = affine_sky = None
    m = (model.module if hasattr(model, "module") else model) if model is not None else None
    use_sky = m is not None and getattr(config, "model_sky", False)
    if m is not None and getattr(config, "brightness_correction", False):
        idx_t = torch.as_tensor(eval_camidx).reshape(-1)[:1].to(next(m.brightness_corr.parameters()).device)
        res = m.brightness_corr(indices=idx_t.repeat(2))          # .squeeze() in the reference needs >= 2 indices
        affine, affine_sky = (res[0][0], res[1][0]) if use_sky else (res[0], None)
        r.set_rgb_affine(affine)    # reset right after this image's render_rays
    nl = r.num_levels
    # Whole-image buffers: the packed pixels and - where the caller gets them or the sky blend reads them - the NeRF
    # level's weights.  The per-sample bundles (`ray_sdist` / `ray_weights` / `ray_rgbs`) are wanted for vis_num_rays
    # random rays only (models.py:L996-1005), so they come from a second render of just those rays below instead of
    # materialising [rays, S+1] / [rays, S] / [rays, S, 3] for every level of every ray.
    need_weights = world == 1 or return_weights or use_sky
    want = ["packed"] + ([f"weights_{nl - 1}"] if need_weights else [])
    if return_weights:
        want.append("sample_coord")
    nv = r.vis_num_rays
    n_local = stop - start
    # The reference keeps the first vis_num_rays rays of every render_chunk_size chunk (models.py:L286-295), concatenates
    # them over the chunks and then draws `randperm(total)[:vis_num_rays]` (L996-1005): the same candidates and the same
    # draw from torch's CPU generator here, so a seeded run shows the same bundle.  (With several ranks the reference's
    # candidates are spread over its per-chunk rank slices; here they are this rank's tile - a random bundle either way.)
    ref_chunk = int(getattr(config, "render_chunk_size", 16384) or 16384)
    cand = (torch.arange(0, max(n_local, 1), ref_chunk)[:, None] + torch.arange(nv)[None, :])
    cand = cand[(cand < max(n_local, 1)) & (torch.arange(nv)[None, :] < ref_chunk)]
    pick = cand[torch.randperm(cand.numel())[:nv]].to(r.device)
    vis_want = [f"sdist_{l}" for l in range(nl)] + [f"weights_{l}" for l in range(nl)] + ["sample_rgb"]
    # N > 1 on NVLink-connected GPUs: the tile exchange is fused into the compositing kernel (peer.PeerImage) unless the
    # sky head still has to blend into the pixels afterwards or `config.ucnerf_peer_exchange = False`
    peer = _peer_image(r, world, per, config) if (world > 1 and not use_sky) else None
    peer_img = None
    # this rank's tile is a run of whole image rows when it starts on a row: tell the library the row length, so that the
    # gather kernel's warps take pixel patches instead of 32 pixels of one row (bit-identical, see ray_tile_width)
    tiled = width % 4 == 0 and start % width == 0 and hasattr(r, "set_option")
    try:
        if tiled:
            r.set_option("ray_tile_width", width)
        if peer is not None:
            peer_img, out = peer.render(r, local, train_frac, lrv, rank * per, [w for w in want if w != "packed"])
        else:
            out = r.render_rays(local, train_frac, lrv, want)
        if tiled:
            r.set_option("ray_tile_width", 0)     # the picked rays below are no image
        vis = r.render_rays({k: v[pick] for k, v in local.items()}, train_frac, lrv[pick], vis_want)
    finally:
        if tiled:
            r.set_option("ray_tile_width", 0)
        if affine is not None:
            r.set_rgb_affine(None)   # the affine belongs to this image only, also when the render raises
    packed = out["packed"] if peer_img is None else None
    sky_rgbs = None
    if use_sky:
        sky_rgbs = _sky_head(m, local, config)
        if affine_sky is not None:  # models.py:L353-354
            sky_opacity = 1 - torch.sum(out[f"weights_{nl - 1}"], dim=-1, keepdim=True)
            packed[:, 0:3] += sky_opacity * (sky_rgbs @ affine_sky[:3, :3].T + affine_sky[:3, 3])
    if peer_img is not None:
        packed = peer_img[:num_rays].clone()                # every rank's tile is already here (the buffer is reused 2 frames on)
    elif world > 1:
        packed = gather_tiles(packed, world, num_rays)      # the ONE collective per image
        if sky_rgbs is not None:
            sky_rgbs = gather_tiles(sky_rgbs, world, num_rays)
    rendering = {
        "rgb": packed[:, 0:3].reshape(height, width, 3),
        "depth": packed[:, 3].reshape(height, width),
        "acc": packed[:, 4].reshape(height, width),
        "distance_mean": packed[:, 5].reshape(height, width),
        "distance_median": packed[:, 6].reshape(height, width),
        "distance_percentile_5": packed[:, 7].reshape(height, width),
        "distance_percentile_95": packed[:, 8].reshape(height, width),
    }
    if need_weights and (world == 1 or return_weights):
        rendering["weights"] = gather_tiles(out[f"weights_{nl - 1}"], world, num_rays).reshape(height, width, -1)
    if return_weights:   # models.py:L976-978: the NeRF level's sample coordinates for extract.py
        co = out["sample_coord"].reshape(out["sample_coord"].shape[0], -1)
        rendering["coord"] = gather_tiles(co, world, num_rays).reshape(height, width, -1, 3)
    # ray bundles for vis.visualize_suite: a random subset of vis_num_rays of this rank's rays per level
    final_rgb = torch.sum(vis["sample_rgb"] * vis[f"weights_{nl - 1}"][..., None], dim=-2)
    rendering["ray_sdist"] = [vis[f"sdist_{l}"] for l in range(nl)]
    rendering["ray_weights"] = [vis[f"weights_{l}"] for l in range(nl)]
    rendering["ray_rgbs"] = [final_rgb[:, None, :].expand(-1, r.samples[l], -1) for l in range(nl - 1)] + \
                            [vis["sample_rgb"]]
    if sky_rgbs is not None:
        rendering["sky_rgbs"] = sky_rgbs.reshape(height, width, 3)       # models.py:L336-337
    if affine is not None:
        rendering["affine_trans"] = affine[None, None].expand(height, width, 3, 4)   # models.py:L361-363, L987-990
        if affine_sky is not None:
            rendering["affine_trans_sky"] = affine_sky[None, None].expand(height, width, 3, 4)
    return rendering


def _sky_head(m, rays, config):
    """Sky colours of a flat ray dict: the tensor-core kernel when `model.skynerf` has the reference architecture (a
    SkyHead is built once and cached on the module), else - or with `config.ucnerf_reference_sky = True` - the
    reference's own torch head."""
    if not getattr(config, "ucnerf_reference_sky", False):
        head = getattr(m, "_ucnerf_b200_sky", None)
        versions = _param_versions(m.skynerf)
        if head and versions != getattr(m, "_ucnerf_b200_sky_versions", None):
            head.close()         # sky weights moved (training step / load_state_dict): re-pack them
            head = None
        if head is None:
            try:
                head = SkyHead({"skynerf." + k: v for k, v in m.skynerf.state_dict().items()}, device=rays["origins"].device)
            except (NotImplementedError, KeyError):
                head = False
            object.__setattr__(m, "_ucnerf_b200_sky", head)
            object.__setattr__(m, "_ucnerf_b200_sky_versions", versions)
        if head:
            return head.render(rays["origins"], rays["directions"], rays["far"], rays["cam_dirs"])
    return _reference_sky_head(m, rays, getattr(config, "render_chunk_size", 16384))


def _reference_sky_head(m, rays, chunk):
    """models.py:L326-337 on a flat ray dict: the reference's `render_rays(ray_batch, network_fn=model.skynerf)`
    (resolved from the module the live model class comes from), chunked so the 120-sample activations fit."""
    import sys
    render_rays = getattr(sys.modules[type(m).__module__], "render_rays")
    o, d, far, cam = rays["origins"], rays["directions"], rays["far"].reshape(-1, 1), rays["cam_dirs"]
    outs = []
    for a in range(0, o.shape[0], chunk):
        sky_near = far[a:a + chunk]
        sky_far = torch.full_like(sky_near * 1.2, float(far[0]) * 1.5)   # L329: the first ray's far for the whole batch
        ray_batch = torch.concat([o[a:a + chunk], d[a:a + chunk], sky_near, sky_far, cam[a:a + chunk]], dim=-1)
        outs.append(render_rays(ray_batch=ray_batch, network_fn=m.skynerf)["rgb_map"])
    return torch.cat(outs)

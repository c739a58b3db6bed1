"""Host-side mirror of the reference's model classes (internal/models.py: `Model` L28-95, `MLP` L367-483, `NerfMLP`,
`PropMLP`) for the configuration this package supports: the same class attributes (set as keyword arguments, the way gin
sets them), the same sub-module and parameter names - a reference checkpoint loads with `load_state_dict`
(`strict=False` when it also carries the heads' `skynerf.*` / `brightness_corr.*` entries) - and the same
`forward(rand, batch, train_frac, compute_extras, zero_glo=True, eval_camidx=None) -> (renderings, ray_history)`.
The compute is not here: training mode runs `train_forward.level_loop` (native resampling, cast_rays, pooled hash-grid
encode, compositing around this module's `nn.Linear` layers), eval mode the fused render path (`render.HotPathModel`).
The sky / brightness heads of the reference (models.py:L84-92, L326-363) are separate modules and not mirrored here."""
import numpy as np
import torch
import torch.nn as nn

from .gridencoder import GridEncoder


def _set_kwargs(obj, kwargs):
    for k, v in kwargs.items():
        if not hasattr(type(obj), k):
            raise TypeError(f"{type(obj).__name__} has no hyper-parameter {k!r}")
        setattr(obj, k, v)


class MLP(nn.Module):
    """models.py:L367-483 (the attributes the supported configuration reads; defaults as the reference, except
    `disable_density_normals`, whose reference default cannot run - SURVEY.md section 8c (vi))."""
    bottleneck_width: int = 256
    net_depth_viewdirs: int = 2
    net_width_viewdirs: int = 256
    skip_layer_dir: int = 0
    num_rgb_channels: int = 3
    deg_view: int = 4
    bottleneck_noise: float = 0.0
    density_bias: float = -1.
    density_noise: float = 0.
    rgb_premultiplier: float = 1.
    rgb_bias: float = 0.
    rgb_padding: float = 0.001
    disable_density_normals: bool = True
    disable_rgb: bool = False
    warp_fn = 'contract'
    num_glo_features: int = 0
    scale_featurization: bool = False
    grid_level_interval: int = 2
    grid_level_dim: int = 4
    grid_base_resolution: int = 16
    grid_disired_resolution: int = 8192      # (sic) the reference's spelling, kept for gin / kwargs compatibility
    grid_log2_hashmap_size: int = 21

    def __init__(self, **kwargs):
        super().__init__()
        _set_kwargs(self, kwargs)
        if not self.disable_density_normals or self.num_glo_features > 0 or self.scale_featurization:
            raise NotImplementedError("ucnerf_b200.models.MLP: normals / GLO / scale_featurization are not supported")
        self.grid_num_levels = int(np.log(self.grid_disired_resolution / self.grid_base_resolution)
                                   / np.log(self.grid_level_interval)) + 1                           # L425-426
        self.encoder = GridEncoder(input_dim=3, num_levels=self.grid_num_levels, level_dim=self.grid_level_dim,
                                   base_resolution=self.grid_base_resolution,
                                   desired_resolution=self.grid_disired_resolution,
                                   log2_hashmap_size=self.grid_log2_hashmap_size, gridtype='hash', align_corners=False)
        self.density_layer = nn.Sequential(nn.Linear(self.encoder.output_dim, 64), nn.ReLU(),
                                           nn.Linear(64, 1 if self.disable_rgb else self.bottleneck_width))   # L438-441
        if not self.disable_rgb:
            dim_dir_enc = 3 + 2 * 3 * self.deg_view                      # pos_enc with append_identity (coord.py:L214-225)
            last_dim_rgb = input_dim_rgb = self.bottleneck_width + dim_dir_enc
            for i in range(self.net_depth_viewdirs):                     # L475-483
                lin = nn.Linear(last_dim_rgb, self.net_width_viewdirs)
                torch.nn.init.kaiming_uniform_(lin.weight)
                self.register_module(f"lin_second_stage_{i}", lin)
                last_dim_rgb = self.net_width_viewdirs
                if i == self.skip_layer_dir:
                    last_dim_rgb += input_dim_rgb
            self.rgb_layer = nn.Linear(last_dim_rgb, self.num_rgb_channels)


class NerfMLP(MLP):
    pass


class PropMLP(MLP):
    disable_rgb: bool = True       # configs/*.gin: PropMLP.disable_rgb = True


class Model(nn.Module):
    """models.py:L28-95 + forward L97-365 without the heads."""
    num_prop_samples: int = 64
    num_nerf_samples: int = 32
    num_levels: int = 3
    bg_intensity_range = (1., 1.)
    anneal_slope: float = 10
    stop_level_grad: bool = True
    use_viewdirs: bool = True
    raydist_fn = None
    single_jitter: bool = True
    dilation_multiplier: float = 0.5
    dilation_bias: float = 0.0025
    num_glo_features: int = 0
    learned_exposure_scaling: bool = False
    near_anneal_rate = None
    single_mlp: bool = False
    distinct_prop: bool = True
    resample_padding: float = 0.0
    opaque_background: bool = False
    std_scale: float = 0.5
    prop_desired_grid_size = [512, 2048]

    def __init__(self, config=None, nerf_mlp_kwargs=None, prop_mlp_kwargs=None, **kwargs):
        super().__init__()
        _set_kwargs(self, kwargs)
        if self.single_mlp or not self.distinct_prop or self.num_glo_features > 0:
            raise NotImplementedError("ucnerf_b200.models.Model: single_mlp / shared proposal MLP / GLO are not supported")
        self.config = config
        self.nerf_mlp = NerfMLP(**(nerf_mlp_kwargs or {}))                                           # L62-63
        for i in range(self.num_levels - 1):                                                         # L69-70
            self.register_module(f'prop_mlp_{i}', PropMLP(grid_disired_resolution=self.prop_desired_grid_size[i],
                                                          **(prop_mlp_kwargs or {})))
        self._renderer = None
        self._renderer_versions = None

    def _eval_renderer(self):
        from .render import HotPathModel
        versions = tuple(p._version for p in self.parameters())
        if self._renderer is None:
            self._renderer = HotPathModel.from_reference_model(self, self.config)
        elif versions != self._renderer_versions:
            self._renderer.refresh(self.state_dict())        # the weights moved since the last eval call
        self._renderer_versions = versions
        return self._renderer

    def forward(self, rand, batch, train_frac, compute_extras, zero_glo=True, eval_camidx=None):
        if rand or self.training:
            from .train_forward import level_loop
            return level_loop(self, rand, batch, train_frac, compute_extras)
        with torch.no_grad():
            return self._eval_renderer().forward(False, batch, train_frac, compute_extras, zero_glo, eval_camidx)

"""TEST INFRASTRUCTURE - the named parity cases (config + seeds) shared by oracle/make_golden.py, tests/ and
bench.py.  Config labels follow SURVEY.md section 8(d)."""
from oracle import ucnerf_oracle as O


def three_level_config():
    """The reference's *class defaults* for sampling (models.py:L33-35,L55): 2 proposal levels x 64 samples with
    512 / 2048 grids (L=6 / L=8) + 32 NeRF samples; MLP shapes as waymo.gin."""
    return O.HotPathConfig(num_prop_samples=64, num_nerf_samples=32, prop_grids=[O.GridSpec(512), O.GridSpec(2048)])


CASES = {
    # name: (config factory, n_rays, weight seed, ray seed)
    "config1": (O.config1, 512, 0, 0),
    "waymo": (O.waymo_config, 128, 0, 1),
    "three_level": (three_level_config, 64, 2, 3),
    "target1024": (O.target_config, 16, 0, 4),
}


def make_case(name, n_rays=None):
    factory, n, wseed, rseed = CASES[name]
    cfg = factory()
    params = O.init_params(cfg, seed=wseed)
    batch = O.synthetic_rays(n_rays or n, seed=rseed)
    return cfg, params, batch


def param_checksums(params):
    import torch
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in params.items()
            if v.dtype == torch.float32}


def make_heads(seed=0, n_views=9):
    """Seeded weights of the sky head (`skynerf.*`, models.py:L84-92,L743-795) and the brightness-correction head
    (`brightness_corr.*`, extrinsic_optimizer.py:L4-39) under the reference state_dict names, nn.Linear-style init."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    out = {}

    def linear(name, fin, fout, scale=1.0):
        b = scale / math.sqrt(fin)
        out[name + '.weight'] = (torch.rand((fout, fin), generator=g) * 2 - 1) * b
        out[name + '.bias'] = (torch.rand((fout,), generator=g) * 2 - 1) * b

    for i in range(8):
        linear(f'skynerf.pts_linears.{i}', 3 if i == 0 else (259 if i == 5 else 256), 256, 1.7)
    linear('skynerf.views_linears.0', 283, 128, 1.7)
    linear('skynerf.feature_linear', 256, 256)
    linear('skynerf.alpha_linear', 256, 1, 8.0)
    linear('skynerf.rgb_linear', 128, 3, 4.0)
    out['brightness_corr.latent_code'] = torch.randn((n_views, 4), generator=g) * 0.5
    out['brightness_corr.sky_latent_code'] = torch.randn((n_views, 4), generator=g) * 0.5
    linear('brightness_corr.brightness_MLP.pts_linears.0', 4, 256)
    linear('brightness_corr.brightness_MLP.pts_linears.1', 256, 256)
    linear('brightness_corr.brightness_MLP.pts_linears.2', 256, 256)
    linear('brightness_corr.brightness_MLP.output_linear', 256, 12, 4.0)
    return out
